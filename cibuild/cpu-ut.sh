#!/usr/bin/env bash
# CPU unit / integration tier (the reference's cibuild/cpu-ut/*.sh): build, lint-free import, every non-GPU test.
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests/ -x -q -m "not gpu" "$@"
# kernel tier without a GPU: the SIMT .cu sources compiled for the host (csrc/cuda/emu/cuda_emu.h) -- every op-program model through the GPU
# Processor's interpreter under ASAN, the multi-rank flag protocol under TSAN (the default test run covers a subset of this)
DEEPREC_EMU_SANITIZE_FULL=1 python -m pytest tests/test_cuda_emu_sanitizers.py -x -q

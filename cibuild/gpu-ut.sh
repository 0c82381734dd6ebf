#!/usr/bin/env bash
# GPU tier (cibuild/gpu-ut/*.sh): needs one B200; `GPUS=8 cibuild/gpu-ut.sh` adds the multi-GPU equivalence check.
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build(); g.smoke()"
python -m pytest tests/ -x -q -m gpu "$@"
if [ "${GPUS:-1}" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port 29519 tests/mp_check.py
fi

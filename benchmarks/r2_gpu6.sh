#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2i; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_tier.py -q -m gpu --timeout 180 -x > "$OUT/test_tier.txt" 2>&1; echo "tier tests rc=$?" | tee -a "$OUT/log.txt"; tail -12 "$OUT/test_tier.txt"
timeout 400 python -m pytest tests/test_gpu_rec_engine.py tests/test_gpu_table_engine.py tests/test_gpu_serving.py tests/test_gpu_zz_serving_proto.py -q -m gpu --timeout 180 -x 2>&1 | tail -3

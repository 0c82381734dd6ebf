#!/bin/bash
for m in 0 1 3 15; do
  echo "== DEEPREC_FUSED_DBG=$m"
  DEEPREC_FUSED_DBG=$m timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dlrm_inter_gemm -s 2 -c 2 python bench.py --steps 2 --warmup 1 --prefill 2 --no-graph --skip-e2e 2>&1 | grep -E "gpu__time" | head -3
done
timeout 300 python -m pytest tests/test_gpu_sparse_pipeline.py -q -m gpu --timeout 120 -x -k fused 2>&1 | tail -4
timeout 200 python bench.py --steps 20 --warmup 5 --skip-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['config']['final_loss'])"

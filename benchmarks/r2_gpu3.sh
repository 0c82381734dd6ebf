#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2d; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_sparse_pipeline.py -q -m gpu --timeout 120 -x > "$OUT/test_sp.txt" 2>&1; echo "sp tests rc=$?" | tee -a "$OUT/log.txt"; tail -5 "$OUT/test_sp.txt"
timeout 300 python -m pytest tests/test_gpu_table_engine.py -q -m gpu --timeout 120 -x -k oracle > "$OUT/test_eng.txt" 2>&1; echo "oracle rc=$?" | tee -a "$OUT/log.txt"; tail -3 "$OUT/test_eng.txt"
for extra in "" "--stream warm"; do
  timeout 200 python bench.py --steps 20 --warmup 5 --skip-e2e $extra 2>>"$OUT/bench.err" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench $extra', d['ms_per_step'], d['config']['final_loss'])"
done
DEEPREC_GEMM_BRES=1 timeout 200 python bench.py --steps 20 --warmup 5 --skip-e2e 2>>"$OUT/bench.err" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench BRES=1', d['ms_per_step'])"
timeout 200 python benchmarks/step_timing.py 2>&1 | tail -1 | cut -c1-600
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_dlrm_inter_gemm|k_dot_fwd_tc" -c 3 python bench.py --steps 2 --warmup 1 --prefill 2 --no-graph --skip-e2e 2>&1 | grep -E "k_dlrm|gpu__time" | head -8

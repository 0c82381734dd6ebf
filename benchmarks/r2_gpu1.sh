#!/usr/bin/env bash
# Round-2 single-GPU validation pass: new sparse pipeline tests first, then every GPU test file on its own (one failure must not
# hide the others; a hang is bounded by `timeout`), then the headline bench in both stream modes and a launch list.
set -u
OUT=gpurun_out/r2a; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
for t in tests/test_gpu_sparse_pipeline.py tests/test_gpu_table_engine.py tests/test_gpu_kernels.py tests/test_gpu_nn.py tests/test_gpu_serving.py tests/test_gpu_fp8.py tests/test_gpu_fused_ops.py tests/test_gpu_zz_*.py; do
  timeout 420 python -m pytest "$t" -q -m gpu --timeout 180 > "$OUT/$(basename "$t" .py).txt" 2>&1; echo "$t rc=$?" | tee -a "$OUT/log.txt"
done
timeout 300 python bench.py --steps 20 --warmup 5 > "$OUT/bench_fresh.json" 2> "$OUT/bench_fresh.err"; echo "bench fresh rc=$?" | tee -a "$OUT/log.txt"
timeout 300 python bench.py --steps 20 --warmup 5 --stream warm > "$OUT/bench_warm.json" 2> "$OUT/bench_warm.err"; echo "bench warm rc=$?" | tee -a "$OUT/log.txt"
timeout 300 python bench.py --steps 20 --warmup 5 --impl nccl_comm > "$OUT/bench_nccl_comm.json" 2> "$OUT/bench_nccl_comm.err"; echo "bench nccl_comm rc=$?" | tee -a "$OUT/log.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --prefill 2 --no-graph --skip-e2e > "$OUT/ncu_launch.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/log.txt"
tail -c 1500 "$OUT"/bench_*.json

#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2b; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
for t in tests/test_gpu_sparse_pipeline.py tests/test_gpu_table_engine.py; do
  timeout 420 python -m pytest "$t" -q -m gpu --timeout 180 -x > "$OUT/$(basename "$t" .py).txt" 2>&1; echo "$t rc=$?" | tee -a "$OUT/log.txt"
done
timeout 300 python bench.py --steps 20 --warmup 5 > "$OUT/bench_fresh.json" 2> "$OUT/bench_fresh.err"; echo "bench fresh rc=$?" | tee -a "$OUT/log.txt"
timeout 300 python bench.py --steps 20 --warmup 5 --stream warm --skip-e2e > "$OUT/bench_warm.json" 2> "$OUT/bench_warm.err"; echo "bench warm rc=$?" | tee -a "$OUT/log.txt"
timeout 300 python benchmarks/step_timing.py > "$OUT/step_timing.txt" 2>&1; echo "timing rc=$?" | tee -a "$OUT/log.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --prefill 2 --no-graph --skip-e2e > "$OUT/ncu_launch.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/log.txt"
tail -c 1200 "$OUT"/bench_*.json; cat "$OUT/step_timing.txt" | tail -3

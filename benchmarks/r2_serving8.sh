#!/usr/bin/env bash
# ProcessorGroup on N GPUs of one box: bf16 and fp8 dense paths
set -u
N=${1:-8}
OUT=gpurun_out/r2s$N; mkdir -p "$OUT"
for dt in bf16 fp8; do
  timeout 200 python benchmarks/serving_bench.py --gpus $N --sessions 4 --threads $((4 * N)) --batch 2048 --requests 8000 --dtype $dt 2>>"$OUT/serving.err" | tail -1 > "$OUT/serving_$dt.json"; cut -c1-600 "$OUT/serving_$dt.json"
done
tail -5 "$OUT/serving.err"

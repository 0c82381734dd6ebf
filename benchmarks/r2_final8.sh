#!/usr/bin/env bash
# 8-GPU final pass: correctness (mp_check), headline bench with NVLS on / off, FusedRecEngine DeepFM + DIN, serving on 8 GPUs, all-reduce micro-bench
set -u
N=${1:-8}
OUT=gpurun_out/r2f$N; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DEEPREC_NVLS_2PHASE_BYTES=0 timeout 300 $TR --master-port 29611 tests/mp_check.py > "$OUT/mp_check.txt" 2>&1; echo "mp_check rc=$?" | tee -a "$OUT/log.txt"; tail -1 "$OUT/mp_check.txt"
for nv in 1 0; do
  DEEPREC_NVLS=$nv timeout 300 $TR --master-port 2962$nv bench.py --gpus $N --steps 20 --warmup 5 2> "$OUT/bench_nvls$nv.err" | tail -1 > "$OUT/bench_nvls$nv.json"
  python -c "import json; d=json.loads(open('$OUT/bench_nvls$nv.json').read()); print('bench NVLS=$nv', round(d['value']/1e6,2), 'M', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" 2>/dev/null || tail -5 "$OUT/bench_nvls$nv.err"
done
timeout 300 $TR --master-port 29630 benchmarks/nvls_bench.py > "$OUT/nvls_bench.txt" 2>&1; tail -1 "$OUT/nvls_bench.txt" | cut -c1-600
for m in deepfm din; do
  b=65536; [ "$m" = din ] && b=16384
  timeout 300 $TR --master-port 2964$((RANDOM%9)) benchmarks/zoo_bench.py --engine --model $m --batch $b --steps 20 --warmup 5 2> "$OUT/zoo_${m}.err" | tail -1 > "$OUT/zoo_${m}.json"
  python -c "import json; d=json.loads(open('$OUT/zoo_${m}.json').read()); print('$m n$N', round(d['value']/1e6,2), 'M', d['ms_per_step'], d['final_loss'])" 2>/dev/null || tail -5 "$OUT/zoo_${m}.err"
done
if [ "${2:-}" = "serving" ]; then
  for dt in bf16; do
    timeout 300 python benchmarks/serving_bench.py --gpus $N --sessions 4 --threads $((4 * N)) --batch 2048 --requests 8000 --dtype $dt 2>>"$OUT/serving.err" | tail -1 > "$OUT/serving_$dt.json"; cut -c1-500 "$OUT/serving_$dt.json"
  done
fi

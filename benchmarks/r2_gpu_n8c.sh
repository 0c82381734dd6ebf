#!/usr/bin/env bash
set -u
N=${1:-8}
OUT=gpurun_out/r2n${N}c; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/mp_check.py > "$OUT/mp_check.txt" 2>&1; echo "mp_check rc=$?" | tee -a "$OUT/log.txt"
for sb in 4 8; do
  timeout 300 $TR --master-port 2962$sb bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e --sparse-blocks $sb > "$OUT/bench_sb$sb.json" 2> "$OUT/bench_sb$sb.err"; echo "bench sb$sb rc=$?" | tee -a "$OUT/log.txt"
done
timeout 300 $TR --master-port 29614 benchmarks/step_timing.py > "$OUT/step_timing.txt" 2>&1; echo "timing rc=$?" | tee -a "$OUT/log.txt"
grep "rank" "$OUT/step_timing.txt" | sort | cut -c1-330
tail -1 "$OUT/mp_check.txt"
for f in "$OUT"/bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],4))"; done

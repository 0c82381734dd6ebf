#!/usr/bin/env bash
set -u
OUT=gpurun_out/r2g; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_rec_engine.py -q -m gpu --timeout 180 -x > "$OUT/test_rec.txt" 2>&1; echo "rec tests rc=$?" | tee -a "$OUT/log.txt"; tail -15 "$OUT/test_rec.txt"
for m in deepfm din; do
  [ -s "$OUT/stop" ] && break
  for b in 8192 65536; do
    [ "$m" = din ] && [ "$b" = 65536 ] && b=16384
    timeout 300 python benchmarks/zoo_bench.py --engine --model $m --batch $b --steps 20 --warmup 5 2> "$OUT/zoo_${m}_$b.err" | tail -1 > "$OUT/zoo_${m}_$b.json"; echo "zoo $m $b rc=$?" | tee -a "$OUT/log.txt"
    python -c "import json; d=json.loads(open('$OUT/zoo_${m}_$b.json').read()); print('$m', $b, round(d['value']/1e6,2), 'M', d['ms_per_step'], d['final_loss'])" 2>/dev/null || tail -5 "$OUT/zoo_${m}_$b.err"
  done
done

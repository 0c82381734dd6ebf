#!/usr/bin/env bash
# 2-GPU pass: checkpoint re-shard test + mp_check + FusedRecEngine at 2 ranks
set -u
OUT=gpurun_out/r2h; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests/test_gpu_rec_engine.py tests/test_gpu_multi.py -q -m gpu --timeout 300 -x > "$OUT/tests.txt" 2>&1; echo "tests rc=$?" | tee -a "$OUT/log.txt"; tail -6 "$OUT/tests.txt"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for m in deepfm din; do
  b=65536; [ "$m" = din ] && b=16384
  timeout 300 $TR --master-port 2964$((RANDOM%9)) benchmarks/zoo_bench.py --engine --model $m --batch $b --steps 20 --warmup 5 2> "$OUT/zoo_${m}_n2.err" | tail -1 > "$OUT/zoo_${m}_n2.json"
  python -c "import json; d=json.loads(open('$OUT/zoo_${m}_n2.json').read()); print('$m n2', round(d['value']/1e6,2), 'M', d['ms_per_step'], d['final_loss'])" 2>/dev/null || tail -5 "$OUT/zoo_${m}_n2.err"
done

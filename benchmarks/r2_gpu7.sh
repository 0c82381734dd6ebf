#!/usr/bin/env bash
set -u
N=${1:-2}
OUT=gpurun_out/r2j; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/mp_check.py > "$OUT/mp_check_nvls.txt" 2>&1; echo "mp_check(nvls) rc=$?" | tee -a "$OUT/log.txt"; tail -2 "$OUT/mp_check_nvls.txt"
DEEPREC_NVLS=0 timeout 300 $TR --master-port 29612 tests/mp_check.py > "$OUT/mp_check_p2p.txt" 2>&1; echo "mp_check(p2p) rc=$?" | tee -a "$OUT/log.txt"; tail -1 "$OUT/mp_check_p2p.txt"
for nv in 1 0; do
  DEEPREC_NVLS=$nv timeout 300 $TR --master-port 2962$nv bench.py --gpus $N --steps 20 --warmup 5 --skip-e2e 2> "$OUT/bench_nvls$nv.err" | tail -1 > "$OUT/bench_nvls$nv.json"
  python -c "import json; d=json.loads(open('$OUT/bench_nvls$nv.json').read()); print('bench NVLS=$nv', round(d['value']/1e6,2), 'M', d['ms_per_step'])" 2>/dev/null || tail -5 "$OUT/bench_nvls$nv.err"
done
timeout 300 $TR --master-port 29630 benchmarks/nvls_bench.py > "$OUT/nvls_bench.txt" 2>&1; tail -4 "$OUT/nvls_bench.txt"

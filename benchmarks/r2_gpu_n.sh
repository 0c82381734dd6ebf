#!/usr/bin/env bash
# multi-GPU validation: mp_check (P2P unique-first pipeline == NCCL SOK dataflow), bench at N, phase timing.   usage: r2_gpu_n.sh N
set -u
N=${1:-2}
OUT=gpurun_out/r2n$N; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/mp_check.py > "$OUT/mp_check.txt" 2>&1; echo "mp_check rc=$?" | tee -a "$OUT/log.txt"
timeout 300 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > "$OUT/bench_fresh.json" 2> "$OUT/bench_fresh.err"; echo "bench fresh rc=$?" | tee -a "$OUT/log.txt"
timeout 300 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --stream warm --skip-e2e > "$OUT/bench_warm.json" 2> "$OUT/bench_warm.err"; echo "bench warm rc=$?" | tee -a "$OUT/log.txt"
timeout 300 $TR --master-port 29614 benchmarks/step_timing.py > "$OUT/step_timing.txt" 2>&1; echo "timing rc=$?" | tee -a "$OUT/log.txt"
if [ "${2:-}" = "nccl" ]; then
  timeout 300 $TR --master-port 29615 bench.py --gpus $N --steps 20 --warmup 5 --impl nccl_comm --skip-e2e > "$OUT/bench_nccl_comm.json" 2> "$OUT/bench_nccl_comm.err"; echo "bench nccl_comm rc=$?" | tee -a "$OUT/log.txt"
fi
tail -5 "$OUT/mp_check.txt"; tail -c 900 "$OUT"/bench_*.json; grep "rank" "$OUT/step_timing.txt" | tail -8

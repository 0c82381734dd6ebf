#!/usr/bin/env bash
# First GPU call of a round: validate everything that was written without a GPU (tests/test_gpu_zz_*.py) and A/B-measure the
# opt-in paths against the defaults.  One box, one GPU unless noted; everything lands under gpurun_out/ab/.
#
#   gpurun --timeout 1500 -- 'bash benchmarks/ab_validate.sh'                 # 1 GPU part
#   gpurun --gpus 8 --timeout 900 -- 'bash benchmarks/ab_validate.sh multi'   # 8 GPU part
set -u
OUT=gpurun_out/ab; mkdir -p "$OUT"
run() { echo "=== $*" | tee -a "$OUT/log.txt"; "$@" >>"$OUT/log.txt" 2>&1; echo "rc=$?" | tee -a "$OUT/log.txt"; }

if [ "${1:-single}" = "multi" ]; then
  N=${2:-8}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
  for bres in 0 1; do
    DEEPREC_GEMM_BRES=$bres $TR bench.py --gpus $N --steps 30 --warmup 5 2>>"$OUT/log.txt" | tail -1 > "$OUT/bench_n${N}_bres${bres}.json"
  done
  $TR benchmarks/zoo_bench.py --model deepfm --steps 20 --warmup 5 2>>"$OUT/log.txt" | tail -1 > "$OUT/zoo_deepfm_n${N}.json"        # BASELINE config #3
  python benchmarks/serving_bench.py --gpus $N --sessions 4 --threads $((4 * N)) --batch 2048 --requests 8000 --dtype fp8 2>>"$OUT/log.txt" | tail -1 > "$OUT/serving_n${N}_fp8.json"   # config #5
  python benchmarks/serving_bench.py --gpus $N --sessions 4 --threads $((4 * N)) --batch 2048 --requests 8000 --dtype bf16 2>>"$OUT/log.txt" | tail -1 > "$OUT/serving_n${N}_bf16.json"
  exit 0
fi

# ---- 1. the unvalidated tests, each file on its own so that one failure does not hide the others
for t in tests/test_gpu_zz_*.py; do
  DEEPREC_RUN_UNVALIDATED=1 timeout 600 python -m pytest "$t" -q -m gpu -x > "$OUT/$(basename "$t" .py).txt" 2>&1; echo "$t rc=$?" | tee -a "$OUT/log.txt"
done
# ---- 1b. kernels written in the last session of round 2 (CTA-pair GEMM, block-scaled fp8): tests, then the GEMM A/B table, then the step A/B
for t in tests/test_gpu_zzzzzz_gemm_2cta.py tests/test_gpu_zzzzzzz_mxfp8.py; do
  timeout 900 python -m pytest "$t" -q -m gpu > "$OUT/$(basename "$t" .py).txt" 2>&1; echo "$t rc=$?" | tee -a "$OUT/log.txt"
done
timeout 600 python benchmarks/gemm_ab.py > "$OUT/gemm_ab.jsonl" 2>>"$OUT/log.txt"
for c in 0 1; do
  DEEPREC_GEMM_2CTA=$c timeout 600 python bench.py --steps 30 --warmup 5 2>>"$OUT/log.txt" | tail -1 > "$OUT/bench_n1_2cta${c}.json"
done
DEEPREC_GEMM_2CTA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tn_2cta -c 1 -s 10 -o "$OUT/prof_gemm_2cta" -f python bench.py --steps 2 --warmup 1 >>"$OUT/log.txt" 2>&1
# ---- 2. A/B: B-resident GEMM
for bres in 0 1; do
  DEEPREC_GEMM_BRES=$bres timeout 600 python bench.py --steps 30 --warmup 5 2>>"$OUT/log.txt" | tail -1 > "$OUT/bench_n1_bres${bres}.json"
done
# ---- 3. A/B: one-hot group lookup fast path on the framework-API models
for fast in 0 1; do
  DEEPREC_FAST_ONEHOT=$fast python benchmarks/zoo_bench.py --model deepfm 2>>"$OUT/log.txt" | tail -1 > "$OUT/zoo_deepfm_fast${fast}.json"
done
python benchmarks/zoo_bench.py --model din 2>>"$OUT/log.txt" | tail -1 > "$OUT/zoo_din.json"
# serving with more callers than sessions: RR now takes the first idle session (compare with profiles/serving_bench_n1.jsonl)
python benchmarks/serving_bench.py --sessions 4 --threads 16 --batch 256 --requests 8000 --dtype bf16 2>>"$OUT/log.txt" | tail -1 > "$OUT/serving_n1_s4_t16_b256.json"
# ---- 4. ncu of the two new kernels (one capture each)
DEEPREC_GEMM_BRES=1 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tn_v2 -c 1 -s 40 -o "$OUT/prof_gemm_bres" -f python bench.py --steps 2 --warmup 1 >>"$OUT/log.txt" 2>&1
DEEPREC_RUN_UNVALIDATED=1 ncu --set full --clock-control none --import-source on -k regex:k_din_attention_fwd -c 1 -o "$OUT/prof_din_attention" -f python -m pytest tests/test_gpu_zz_attention.py -q -m gpu -k "257" >>"$OUT/log.txt" 2>&1
ls -la "$OUT" >>"$OUT/log.txt"

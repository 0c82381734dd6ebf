#!/bin/bash
# Serving latency/QPS sweep + one `ncu --set full` capture per hot kernel (single GPU).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
for dt in bf16 fp8; do
for cfg in "1 1 1" "4 8 256" "4 16 2048"; do
  set -- $cfg
  timeout 200 python benchmarks/serving_bench.py --sessions $1 --threads $2 --batch $3 --requests 2000 --dtype $dt 2>/dev/null | tail -1 | cut -c1-420
done
done > gpurun_out/serving_bench.jsonl
cat gpurun_out/serving_bench.jsonl
for k in k_gemm_tn_v2 k_gemm_nt_splitk k_bn_bwd_apply_v2; do
  [ -f gpurun_out/prof_$k.ncu-rep ] && continue
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/prof_$k python benchmarks/profile_kernels.py > /dev/null 2>&1
done
# one launch per step: skip the first (cold) step, capture the second
for k in k_dot_fwd_tc k_dot_bwd_tc k_lookup k_accumulate k_apply k_gather; do
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$k python benchmarks/profile_kernels.py > /dev/null 2>&1
done
ls gpurun_out/*.ncu-rep | wc -l

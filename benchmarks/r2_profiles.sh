#!/bin/bash
# One `ncu --set full` capture per hot kernel of the round-2 step (single GPU; never a multi-rank command).  The eager (no-graph) bench
# runs 2 prefill + 1 warm-up + 2 timed steps; launch-skip counts pick a steady-state instance.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --prefill 2 --no-graph --skip-e2e"
cap() {  # name regex skip count
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_$1 $CMD > /dev/null 2>&1; echo "$1 rc=$?"
}
cap sp_dedup k_sp_dedup 3 1
cap sp_lookup k_sp_lookup 3 1
cap sp_segsum k_sp_segsum 3 1
cap sp_grad k_sp_grad 3 1
cap apply "k_apply" 3 1
cap dot_fwd k_dot_fwd_tc 3 1
cap dot_bwd k_dot_bwd_tc 3 1
cap gemm_tn_v2 k_gemm_tn_v2 36 4
cap gemm_dw k_gemm_nt_splitk 24 3
cap bn_bwd_apply k_bn_bwd_apply_v2 12 2
ls -la gpurun_out/*.ncu-rep

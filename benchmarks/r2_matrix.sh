#!/usr/bin/env bash
# 1-GPU measurement matrix the round-1 verdict asked for: sustained (>= 5 s) run with its clock record, alpha in {0, 1.05},
# per-GPU batch in {8192, 65536}, admission on.  One JSON line each -> gpurun_out/r2m/*.json
set -u
OUT=gpurun_out/r2m; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 400 python bench.py "$@" 2> "$OUT/$name.err" | tail -1 > "$OUT/$name.json"; python -c "
import json; d=json.loads(open('$OUT/$name.json').read()); c=d['config']; print('$name', round(d['value']/1e6,2),'M', round(d['ms_per_step'],4),'ms', 'new_keys', c['new_keys_per_step'], 'uniq', round(c['unique_ratio'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'W', d['clocks'].get('power_w_max'), 'e2e', d.get('e2e',{}).get('ms_per_step'))" 2>/dev/null || tail -3 "$OUT/$name.err"; }
run fresh_b65536 --steps 20 --warmup 5
run sustained_5s --steps 4500 --warmup 5 --pool 64 --skip-e2e
run alpha0_b65536 --steps 20 --warmup 5 --alpha 0 --skip-e2e
run fresh_b8192 --steps 40 --warmup 5 --batch 8192
run alpha0_b8192 --steps 40 --warmup 5 --batch 8192 --alpha 0 --skip-e2e
run admission_ff2 --steps 20 --warmup 5 --filter-freq 2 --skip-e2e
run e2e_100 --steps 100 --warmup 5

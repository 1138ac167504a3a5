#!/bin/bash
mkdir -p gpurun_out
leg() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; }
leg j_attn 300 python scripts/gpu_diag.py --group attn
grep -E "^(FAIL|EXC)" gpurun_out/j_attn.log | head
leg j_tuned 300 python scripts/gpu_diag.py --group tuned
grep -E "^(FAIL|EXC)" gpurun_out/j_tuned.log | head
leg j_mb_attn 200 python scripts/gpu_microbench.py attn
cat gpurun_out/j_mb_attn.log | tail -n 14
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e"
timeout -k 10 400 python bench.py $Q > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
echo "bench rc=$? $(python -c "
import json; d=json.load(open('gpurun_out/j_bench.json')); b=d.get('batch8',{})
print('B=1', round(d['value'],1), round(d['ms_per_step'],3), 'steady', round(d['steady_state']['ms_per_step'],3), 'fp', d['x_final_fingerprint'],
      '| B=8', round(b['value'],1), round(b['ms_per_step'],3), 'steady', round(b['steady_state']['ms_per_step'],3), 'fp', b['x_final_fingerprint'])
" 2>&1 | tail -n 2)"

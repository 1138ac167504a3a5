#!/bin/bash
# Round-2 GPU call I: attention with the softmax denominator from the PV MMA (ones row in V^T), timestep tables hoisted out
# of the step graph, GroupNorm path rule — numerics, parity through the graphs, timing.
mkdir -p gpurun_out
leg() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; }
leg i_attn 300 python scripts/gpu_diag.py --group attn
grep -E "^(FAIL|EXC)" gpurun_out/i_attn.log | head
leg i_tuned 300 python scripts/gpu_diag.py --group tuned
grep -E "^(FAIL|EXC)" gpurun_out/i_tuned.log | head
leg i_misc 300 python scripts/gpu_diag.py --group misc
grep -E "^(FAIL|EXC)" gpurun_out/i_misc.log | head
leg i_mb_attn 200 python scripts/gpu_microbench.py attn
cat gpurun_out/i_mb_attn.log | tail -n 12
leg i_parity 900 python -m pytest tests/test_parity_r2_gpu.py tests/test_parity_gpu.py tests/test_dropin_gpu.py -x -q -m gpu
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-roofline"
timeout -k 10 400 python bench.py $Q > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench rc=$? $(python -c "
import json; d=json.load(open('gpurun_out/i_bench.json')); b=d.get('batch8',{})
print('B=1', round(d['value'],1), round(d['ms_per_step'],3), 'steady', round(d['steady_state']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['launches_per_step'], 'fp', d['x_final_fingerprint'],
      '| B=8', round(b['value'],1), round(b['ms_per_step'],3), 'steady', round(b['steady_state']['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'fp', b['x_final_fingerprint'])
" 2>&1 | tail -n 2)"

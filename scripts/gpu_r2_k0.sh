#!/bin/bash
mkdir -p gpurun_out
leg() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; }
leg k0_gemm 300 python scripts/gpu_diag.py --group gemm
grep -E "^(FAIL|EXC)|folded" gpurun_out/k0_gemm.log | head
leg k0_parity 900 python -m pytest tests/test_parity_r2_gpu.py tests/test_parity_gpu.py -x -q -m gpu
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e"
for i in 1 2; do
timeout -k 10 400 python bench.py $Q > gpurun_out/k0_bench$i.json 2> gpurun_out/k0_bench$i.err
echo "bench$i rc=$? $(python -c "
import json; d=json.load(open('gpurun_out/k0_bench$i.json')); b=d.get('batch8',{})
print('B=1', round(d['value'],1), round(d['ms_per_step'],3), 'steady', round(d['steady_state']['ms_per_step'],3), 'launches', d['launches_per_step'], 'fp', d['x_final_fingerprint'],
      '| B=8', round(b['value'],1), round(b['ms_per_step'],3), 'steady', round(b['steady_state']['ms_per_step'],3), 'fp', b['x_final_fingerprint'])
" 2>&1 | tail -n 2)"
done

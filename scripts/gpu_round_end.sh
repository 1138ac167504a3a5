#!/bin/bash
# Round-2 final GPU call: what the driver runs at round end (smoke, -m gpu suite, bench both arms) plus the committed
# profiles of the final state (NVTX-scoped launch lists at one and eight frames, one --set full pass over the dominant kernels).
mkdir -p gpurun_out
leg() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; }
leg k_smoke 600 python __graft_entry__.py smoke
leg k_pytest 1500 python -m pytest tests -q -m gpu -x
grep -E "passed|failed|Error|FAILED" gpurun_out/k_pytest.log | tail -n 4
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/k_full.json 2> gpurun_out/k_full.err
echo "bench rc=$? $(python -c "
import json; d=json.load(open('gpurun_out/k_full.json')); b=d.get('batch8',{})
print('B=1', round(d['value'],1), round(d['ms_per_step'],3), 'steady', round(d['steady_state']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'launches', d['launches_per_step'],
      '| B=8', round(b['value'],1), round(b['ms_per_step'],3), 'steady', round(b['steady_state']['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'roof', round(b['roofline']['frac'],3))
print('eager', [(r['frames'], round(r['value'],1), round(r['ms_per_step'],2)) for r in d['gpu_eager_baseline'].get('runs', [])], 'cpu', round(d['cpu_baseline']['value'],4), 'clocks', d['clocks'])
" 2>&1 | tail -n 3)"
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/k_ref.json 2> gpurun_out/k_ref.err; echo "ref arm rc=$? $(cut -c1-160 gpurun_out/k_ref.json)"
NQ="--steps 3 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e --nvtx"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout -k 10 400 ncu --nvtx --nvtx-include "mdb_step/" --metrics $M --clock-control none --csv --log-file gpurun_out/k_launches_b1.csv \
  python bench.py $NQ --no-batch8 > gpurun_out/k_ncu_b1.log 2>&1; echo "ncu b1 rc=$? $(wc -l < gpurun_out/k_launches_b1.csv) lines"
timeout -k 10 400 ncu --nvtx --nvtx-include "mdb_step/" --metrics $M --clock-control none --csv --log-file gpurun_out/k_launches_b8.csv \
  python bench.py $NQ --batch 8 > gpurun_out/k_ncu_b8.log 2>&1; echo "ncu b8 rc=$? $(wc -l < gpurun_out/k_launches_b8.csv) lines"
KF='regex:^(gemm_|attn|gn_|layernorm)'
timeout -k 10 600 ncu -k "$KF" --set full --clock-control none -o /tmp/k_targets -f python scripts/gpu_ncu_targets.py 1 > gpurun_out/k_ncu_targets.log 2>&1
echo "ncu targets rc=$?"
ncu -i /tmp/k_targets.ncu-rep --page raw --csv > gpurun_out/k_targets_raw.csv 2> /dev/null; echo "raw csv $(wc -c < gpurun_out/k_targets_raw.csv) bytes"

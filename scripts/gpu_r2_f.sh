#!/bin/bash
# Round-2 GPU call F: the final default bench line (kept under profiles/), ncu launch lists of one captured step at
# one and at eight frames (only this library's kernels are profiled: model set-up runs at native speed), and one
# `--set full` pass over the dominant kernels at their real shapes (raw CSV only: the .ncu-rep stays on the box).
mkdir -p gpurun_out
timeout -k 10 900 python bench.py --steps 20 --warmup 3 > gpurun_out/f_full.json 2> gpurun_out/f_full.err
echo "bench rc=$? $(python -c "
import json; d=json.load(open('gpurun_out/f_full.json')); b=d.get('batch8',{})
print('B=1', round(d['value'],1), round(d['ms_per_step'],3), 'steady', round(d['steady_state']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3),
      '| B=8', round(b['value'],1), round(b['ms_per_step'],3), 'steady', round(b['steady_state']['ms_per_step'],3), 'e2e', round(b['e2e']['value'],1), 'roof', round(b['roofline']['frac'],3))
print('eager', [(r['frames'], round(r['value'],1), round(r['ms_per_step'],2)) for r in d['gpu_eager_baseline'].get('runs', [])], 'cpu', round(d['cpu_baseline']['value'],4))
" 2>&1 | tail -n 3)"
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref arm rc=$? $(cut -c1-200 gpurun_out/f_ref.json)"
KF='regex:^(gemm_|attn|gn_|layernorm|skinny_|conv3x3|direct_conv|add_kernel|im2col|upsample|nchw_|nhwc_|timestep_|cfg_ddim|splitk_|softmax_)'
NQ="--steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e"
t0=$(date +%s)
timeout -k 10 500 ncu -k "$KF" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/f_launches_b1.csv python bench.py $NQ --no-batch8 > gpurun_out/f_ncu_b1.log 2>&1
echo "ncu b1 rc=$? $(wc -l < gpurun_out/f_launches_b1.csv) lines ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout -k 10 500 ncu -k "$KF" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/f_launches_b8.csv python bench.py $NQ --batch 8 > gpurun_out/f_ncu_b8.log 2>&1
echo "ncu b8 rc=$? $(wc -l < gpurun_out/f_launches_b8.csv) lines ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout -k 10 600 ncu -k "$KF" --set full --clock-control none -o /tmp/f_targets -f python scripts/gpu_ncu_targets.py 1 > gpurun_out/f_ncu_targets.log 2>&1
echo "ncu targets rc=$? ($(( $(date +%s) - t0 )) s)"
ncu -i /tmp/f_targets.ncu-rep --page raw --csv > gpurun_out/f_targets_raw.csv 2> /dev/null; echo "raw csv $(wc -c < gpurun_out/f_targets_raw.csv) bytes"
du -sh gpurun_out

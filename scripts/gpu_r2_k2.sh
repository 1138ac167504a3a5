#!/bin/bash
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e"
run() { local name=$1; shift
timeout -k 10 400 python bench.py $Q "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
echo "$name rc=$? $(python -c "
import json; d=json.load(open('gpurun_out/$name.json')); b=d.get('batch8',{})
print('B=1', round(d['value'],1), 'steady', round(d['steady_state']['ms_per_step'],3), 'bank', round(d['bank_build_ms'],1), 'launches', d['launches_per_step'],
      '| B=8', round(b['value'],1), 'steady', round(b['steady_state']['ms_per_step'],3))
" 2>&1 | tail -n 1)"; }
run k2_nofuse_p128 --ln-fuse-max-rows 0
run k2_nofuse_p256 --ln-fuse-max-rows 0 --tune pair_min_tiles=256
run k2_nofuse_p512 --ln-fuse-max-rows 0 --tune pair_min_tiles=512
run k2_fuse_p256 --tune pair_min_tiles=256
run k2_fuseall_p128 --ln-fuse-max-rows 1000000000
run k2_fuseall_p256 --ln-fuse-max-rows 1000000000 --tune pair_min_tiles=256

#!/bin/bash
# Round-2 GPU call G: clean one-step ncu launch lists (NVTX range around ONE graph replay), new conv cases (rows wider
# than the tile), VAE timing with the implicit-GEMM path at every level.
mkdir -p gpurun_out
leg() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; }
leg g_conv 300 python scripts/gpu_diag.py --group conv
grep -E "^(FAIL|EXC)" gpurun_out/g_conv.log | head
leg g_tuned 300 python scripts/gpu_diag.py --group tuned
grep -E "^(FAIL|EXC)" gpurun_out/g_tuned.log | head
leg g_vae 400 python scripts/gpu_vae_parity.py
grep -E "decode|encode|OK|FAIL" gpurun_out/g_vae.log | tail -n 8
leg g_vaetest 400 python -m pytest tests/test_vae_gpu.py -q -m gpu -x
NQ="--steps 3 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e --nvtx"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
t0=$(date +%s)
timeout -k 10 500 ncu --nvtx --nvtx-include "mdb_step/" --metrics $M --clock-control none --csv --log-file gpurun_out/g_launches_b1.csv \
  python bench.py $NQ --no-batch8 > gpurun_out/g_ncu_b1.log 2>&1
echo "ncu b1 rc=$? $(wc -l < gpurun_out/g_launches_b1.csv) lines ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout -k 10 500 ncu --nvtx --nvtx-include "mdb_step/" --metrics $M --clock-control none --csv --log-file gpurun_out/g_launches_b8.csv \
  python bench.py $NQ --batch 8 > gpurun_out/g_ncu_b8.log 2>&1
echo "ncu b8 rc=$? $(wc -l < gpurun_out/g_launches_b8.csv) lines ($(( $(date +%s) - t0 )) s)"

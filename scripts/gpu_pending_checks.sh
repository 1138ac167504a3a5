#!/bin/bash
# Everything that was written after the round-1 GPU budget ran out and is therefore still opt-in.
# One gpurun call (~3-4 min of box time); each leg is bounded by `timeout` and writes under gpurun_out/.
#   gpurun --timeout 600 -- 'bash scripts/gpu_pending_checks.sh'
mkdir -p gpurun_out
echo "== persistent CTA-pair GEMM: numerics (14 cases), then all three GEMM variants shape by shape"
MDB_TEST_PAIR_MODE=2 timeout 120 python scripts/gpu_diag.py --group pair > gpurun_out/pending_pairp.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/pending_pairp.log
timeout 200 python scripts/gpu_microbench.py pair 0,1,2 > gpurun_out/pending_microbench_pair.log 2>&1; echo "rc=$?"
echo "== VAE decoder (softmax kernel, GroupNorm with 4 channels/group, 128-pixel-row implicit GEMM, im2col at 256/512)"
timeout 200 python scripts/gpu_vae_parity.py > gpurun_out/pending_vae.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/pending_vae.log
echo "== full step with the persistent pair kernel (eight frames)"
MDB_GEMM_PAIR=2 timeout 90 python bench.py --batch 8 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b8_pair2.json 2> gpurun_out/pending_b8_pair2.err; echo "rc=$? (124 = hung)"
echo "== round-1 late additions: persistent pair GEMM with TMA-store epilogue (MDB_GEMM_PAIR=3), pair tiles + cluster split-K (MDB_GEMM_PAIR_SPLITK=1)"
timeout 150 python scripts/gpu_diag.py --group pending --pending-filter MDB_GEMM_PAIR_SPLITK > gpurun_out/pending_pairs.log 2>&1; echo "rc=$? (124 = hung)"; tail -n 3 gpurun_out/pending_pairs.log
timeout 150 python scripts/gpu_diag.py --group pending --pending-filter "'MDB_GEMM_PAIR', '3'" > gpurun_out/pending_pairq.log 2>&1; echo "rc=$? (124 = hung)"; tail -n 3 gpurun_out/pending_pairq.log
timeout 200 python scripts/gpu_microbench.py pair 0,3 > gpurun_out/pending_microbench_pairq.log 2>&1; echo "rc=$?"
timeout 200 python scripts/gpu_microbench.py pairs > gpurun_out/pending_microbench_pairs.log 2>&1; echo "rc=$?"
MDB_GEMM_PAIR=3 timeout 90 python bench.py --batch 8 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b8_pair3.json 2> gpurun_out/pending_b8_pair3.err; echo "rc=$? (124 = hung)"
MDB_GEMM_PAIR_SPLITK=1 timeout 90 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b1_pairs.json 2> gpurun_out/pending_b1_pairs.err; echo "rc=$? (124 = hung)"
MDB_GEMM_PAIR=3 MDB_GEMM_PAIR_SPLITK=1 timeout 90 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b1_pair3_pairs.json 2> gpurun_out/pending_b1_pair3_pairs.err; echo "rc=$? (124 = hung)"
echo "== attention d=40 on the two-Q-tile kernel at two CTAs per SM (MDB_ATTN=4): numerics, then per-shape time against the default (v3)"
MDB_ATTN=4 timeout 120 python scripts/gpu_diag.py --group attn > gpurun_out/pending_attn4.log 2>&1; echo "rc=$? (124 = hung)"; tail -n 2 gpurun_out/pending_attn4.log
timeout 100 python scripts/gpu_microbench.py attn > gpurun_out/pending_microbench_attn3.log 2>&1; echo "rc=$?"
MDB_ATTN=4 timeout 100 python scripts/gpu_microbench.py attn > gpurun_out/pending_microbench_attn4.log 2>&1; echo "rc=$?"
MDB_ATTN=4 timeout 90 python bench.py --batch 8 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b8_attn4.json 2> gpurun_out/pending_b8_attn4.err; echo "rc=$? (124 = hung)"
echo "== secondary baseline: the reference's path as eager PyTorch (cuDNN/cuBLAS/SDPA, fp16 autocast) on this GPU"
timeout 150 python tests/torch_gpu_baseline.py --batch 1 --steps 10 --warmup 2 > gpurun_out/torch_gpu_baseline_b1.json 2> gpurun_out/torch_gpu_baseline_b1.err; echo "rc=$?"; cat gpurun_out/torch_gpu_baseline_b1.json
timeout 150 python tests/torch_gpu_baseline.py --batch 1 --steps 10 --warmup 2 --algorithmic > gpurun_out/torch_gpu_baseline_b1_alg.json 2>> gpurun_out/torch_gpu_baseline_b1.err; echo "rc=$?"; cat gpurun_out/torch_gpu_baseline_b1_alg.json
timeout 150 python tests/torch_gpu_baseline.py --batch 8 --steps 5 --warmup 1 --algorithmic > gpurun_out/torch_gpu_baseline_b8_alg.json 2>> gpurun_out/torch_gpu_baseline_b1.err; echo "rc=$?"; cat gpurun_out/torch_gpu_baseline_b8_alg.json
echo "== single-launch GroupNorm for small batches (MDB_GN_FUSED=1): numerics, per-shape time, one-frame step"
timeout 100 python scripts/gpu_diag.py --group pending --pending-filter MDB_GN_FUSED > gpurun_out/pending_gnfused.log 2>&1; echo "rc=$? (124 = hung)"; tail -n 3 gpurun_out/pending_gnfused.log
timeout 100 python scripts/gpu_microbench.py misc > gpurun_out/pending_microbench_misc.log 2>&1; echo "rc=$?"
MDB_GN_FUSED=1 timeout 100 python scripts/gpu_microbench.py misc > gpurun_out/pending_microbench_misc_gnfused.log 2>&1; echo "rc=$?"
MDB_GN_FUSED=1 timeout 90 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b1_gnfused.json 2> gpurun_out/pending_b1_gnfused.err; echo "rc=$? (124 = hung)"
echo "== bank build overlapped with the first steps (MDB_BANK_OVERLAP=1, one frame); compare value AND the final latent checksum with the default run"
timeout 120 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b1_default.json 2> gpurun_out/pending_b1_default.err; echo "rc=$?"
MDB_BANK_OVERLAP=1 timeout 120 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b1_overlap.json 2> gpurun_out/pending_b1_overlap.err; echo "rc=$? (124 = hung)"
MDB_BANK_OVERLAP=1 MDB_BANK_CHUNK=5 timeout 120 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b1_overlap5.json 2> gpurun_out/pending_b1_overlap5.err; echo "rc=$? (124 = hung)"
echo "== default GEMM tiles with the TMA-store epilogue (MDB_GEMM_TMAST=1): numerics, per-shape time at eight frames, full step"
timeout 120 python scripts/gpu_diag.py --group pending --pending-filter MDB_GEMM_TMAST > gpurun_out/pending_tmast.log 2>&1; echo "rc=$? (124 = hung)"; tail -n 3 gpurun_out/pending_tmast.log
MDB_GEMM_TMAST=1 timeout 200 python scripts/gpu_microbench.py pair 0 > gpurun_out/pending_microbench_tmast.log 2>&1; echo "rc=$?"
MDB_GEMM_TMAST=1 timeout 90 python bench.py --batch 8 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/pending_b8_tmast.json 2> gpurun_out/pending_b8_tmast.err; echo "rc=$? (124 = hung)"

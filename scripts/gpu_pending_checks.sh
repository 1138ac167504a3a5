#!/bin/bash
# Everything that was written after the round-1 GPU budget ran out and is therefore still opt-in.
# Every leg is bounded by `timeout` (124 = hung: the kernel never ran to completion) and writes under gpurun_out/.
#
#   gpurun --timeout 900 -- 'bash scripts/gpu_pending_checks.sh numerics'          # first: ~4 min, decides the rest
#   gpurun --timeout 900 -- 'bash scripts/gpu_pending_checks.sh gemm8 gemm1'       # then the timing legs that matter
#   sections: numerics gemm8 gemm1 attn gn overlap baseline vae tune dropin all
#
# Reading the results: a feature becomes the default when (1) its numerics leg is all "ok", (2) its bench line
# has "finite": true and an "x_final_fingerprint" equal (to ~1e-3) to pending_b1_default.json / pending_b8_default.json,
# and (3) its "value" is higher — `python scripts/decide_defaults.py [--write]` applies exactly these rules to
# gpurun_out/ and writes magicdance_b200/switch_defaults.json.  Then move the green cases from PENDING_CASES to
# ALL_CASES (tests/kernel_cases.py).
mkdir -p gpurun_out
want() { for s in $SECTIONS; do [ "$s" = "$1" ] || [ "$s" = all ] && return 0; done; return 1; }
SECTIONS="${*:-all}"
B1="--steps 50 --warmup 3 --no-cpu-baseline --no-roofline"
B8="--batch 8 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline"
bench() {  # bench <name> <timeout> <bench args...>   (environment switches are inherited from the caller)
  local name=$1 t=$2; shift 2
  timeout -k 10 "$t" python bench.py "$@" > "gpurun_out/$name.json" 2> "gpurun_out/$name.err"
  echo "$name rc=$? $(python - "$name" <<'EOF'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    print(f"value={d['value']:.1f} ms/step={d['ms_per_step']:.3f} finite={d['finite']} fp={d.get('x_final_fingerprint')}")
except Exception as e:
    print("no result:", e)
EOF
)"
}
diag() {  # diag <log name> <timeout> <gpu_diag args...>
  local name=$1 t=$2; shift 2
  timeout -k 10 "$t" python scripts/gpu_diag.py "$@" > "gpurun_out/$name.log" 2>&1
  echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log)"; grep -E "^(FAIL|EXC)" "gpurun_out/$name.log" | head -n 8
}

if want numerics; then
  echo "== numerics of every pending kernel (tests/kernel_cases.py PENDING_CASES), one process per feature"
  diag pending_pairs 150 --group pending --pending-filter MDB_GEMM_PAIR_SPLITK   # pair tiles + split-K in the cluster
  diag pending_pairq 150 --group pending --pending-filter "'MDB_GEMM_PAIR', '3'" # persistent pair, TMA-store epilogue
  diag pending_tmast 120 --group pending --pending-filter MDB_GEMM_TMAST         # default tiles, TMA-store epilogue
  diag pending_gnfused 100 --group pending --pending-filter MDB_GN_FUSED         # single-launch GroupNorm
  MDB_ATTN=4 diag pending_attn4 120 --group attn                                 # d=40 attention, 2 Q tiles x 2 CTAs/SM
  MDB_TEST_PAIR_MODE=2 diag pending_pairp 120 --group pair                       # persistent pair, first cut (was green)
fi

if want gemm8; then
  echo "== GEMM variants at eight frames: shape by shape (warm L2, graph replay), then the full step"
  timeout -k 10 250 python scripts/gpu_microbench.py pair 0,1,2,3 > gpurun_out/pending_microbench_pair.log 2>&1; echo "microbench pair rc=$?"
  MDB_GEMM_TMAST=1 timeout -k 10 200 python scripts/gpu_microbench.py pair 0 > gpurun_out/pending_microbench_tmast.log 2>&1; echo "microbench tmast rc=$?"
  bench pending_b8_default 120 $B8
  MDB_GEMM_PAIR=2 bench pending_b8_pair2 90 $B8
  MDB_GEMM_PAIR=3 bench pending_b8_pair3 90 $B8
  MDB_GEMM_TMAST=1 bench pending_b8_tmast 90 $B8
  MDB_GEMM_PAIR=3 MDB_GEMM_PAIR_SPLITK=1 bench pending_b8_pair3_pairs 90 $B8   # + pair split-K for the one-wave layers
fi

if want gemm1; then
  echo "== one frame: pair tiles + cluster split-K on the weight-streaming layers"
  timeout -k 10 200 python scripts/gpu_microbench.py pairs > gpurun_out/pending_microbench_pairs.log 2>&1; echo "microbench pairs rc=$?"
  bench pending_b1_default 120 $B1
  MDB_GEMM_PAIR_SPLITK=1 bench pending_b1_pairs 90 $B1
  MDB_GEMM_PAIR=3 MDB_GEMM_PAIR_SPLITK=1 bench pending_b1_pair3_pairs 90 $B1   # + persistent pair for the bank build
fi

if want attn; then
  echo "== attention d=40 on the two-Q-tile kernel at two CTAs per SM (MDB_ATTN=4) against the default (v3)"
  timeout -k 10 100 python scripts/gpu_microbench.py attn > gpurun_out/pending_microbench_attn3.log 2>&1; echo "microbench attn v3 rc=$?"
  MDB_ATTN=4 timeout -k 10 100 python scripts/gpu_microbench.py attn > gpurun_out/pending_microbench_attn4.log 2>&1; echo "microbench attn v4 rc=$?"
  MDB_ATTN=4 bench pending_b8_attn4 90 $B8
  MDB_ATTN=4 bench pending_b1_attn4 90 $B1
fi

if want gn; then
  echo "== single-launch GroupNorm for small batches (MDB_GN_FUSED=1)"
  timeout -k 10 100 python scripts/gpu_microbench.py misc > gpurun_out/pending_microbench_misc.log 2>&1; echo "microbench misc rc=$?"
  MDB_GN_FUSED=1 timeout -k 10 100 python scripts/gpu_microbench.py misc > gpurun_out/pending_microbench_misc_gnfused.log 2>&1; echo "microbench misc fused rc=$?"
  MDB_GN_FUSED=1 bench pending_b1_gnfused 90 $B1
  # does the single launch also pay at eight frames (cond+uncond batch 16, bank build batch 25)?
  MDB_GN_FUSED=1 MDB_GN_FUSED_MAX_BATCH=64 timeout -k 10 100 python scripts/gpu_microbench.py misc > gpurun_out/pending_microbench_misc_gnfused64.log 2>&1; echo "microbench misc fused(all batches) rc=$?"
  [ -f gpurun_out/pending_b8_default.json ] || bench pending_b8_default 120 $B8
  MDB_GN_FUSED=1 MDB_GN_FUSED_MAX_BATCH=64 bench pending_b8_gnfused64 90 $B8
fi

if want overlap; then
  echo "== bank build overlapped with the first steps (MDB_BANK_OVERLAP=1, one frame)"
  [ -f gpurun_out/pending_b1_default.json ] || bench pending_b1_default 120 $B1
  MDB_BANK_OVERLAP=1 bench pending_b1_overlap 120 $B1
  MDB_BANK_OVERLAP=1 MDB_BANK_CHUNK=5 bench pending_b1_overlap5 120 $B1
  echo "== everything for one frame together"
  MDB_BANK_OVERLAP=1 MDB_GN_FUSED=1 MDB_GEMM_PAIR_SPLITK=1 bench pending_b1_all 120 $B1
fi

if want baseline; then
  echo "== secondary baseline: the reference's path as eager PyTorch (cuDNN/cuBLAS/SDPA, fp16 autocast) on this GPU"
  for cfg in "b1 --batch 1 --steps 10 --warmup 2" "b1_alg --batch 1 --steps 10 --warmup 2 --algorithmic" "b8_alg --batch 8 --steps 5 --warmup 1 --algorithmic"; do
    set -- $cfg; name=$1; shift
    timeout -k 10 150 python tests/torch_gpu_baseline.py "$@" > "gpurun_out/torch_gpu_baseline_$name.json" 2> "gpurun_out/torch_gpu_baseline_$name.err"
    echo "torch baseline $name rc=$?"; cat "gpurun_out/torch_gpu_baseline_$name.json"
  done
fi

if want vae; then
  echo "== VAE decoder + encoder (softmax kernel, br-padded im2col, GroupNorm with 4 channels/group, 128-pixel-row implicit GEMM, im2col at 256/512)"
  timeout -k 10 400 python scripts/gpu_vae_parity.py > gpurun_out/pending_vae.log 2>&1; echo "vae rc=$?"; tail -n 12 gpurun_out/pending_vae.log
fi

if want tune; then
  echo "== per-shape GEMM launch plan (scripts/gpu_tune_gemm.py): dry run with the validated kernels only; add --variants for the green opt-ins"
  timeout -k 10 600 python scripts/gpu_tune_gemm.py --frames 1,8 --dry-run > gpurun_out/tune_gemm.log 2>&1; echo "tune rc=$?"; tail -n 6 gpurun_out/tune_gemm.log
fi

if want dropin; then
  echo "== drop-in sampler through CUDA graphs (MDB_DROPIN_GRAPH=1): same result as the eager loop, time per step of both"
  timeout -k 10 300 python scripts/gpu_dropin_graph_check.py > gpurun_out/pending_dropin_graph.log 2>&1; echo "dropin rc=$?"; tail -n 6 gpurun_out/pending_dropin_graph.log
fi

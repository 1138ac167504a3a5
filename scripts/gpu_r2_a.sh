#!/bin/bash
# Round-2 GPU call A: new GroupNorm kernels, round-2 parity gates, drop-in graph path, VAE, GPU eager baseline,
# never-run round-1 kernels (decide: promote or delete).  Every leg is bounded by `timeout`; logs under gpurun_out/.
mkdir -p gpurun_out
leg() {  # leg <name> <timeout> <cmd...>
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv,noheader
leg a_misc 240 python scripts/gpu_diag.py --group misc
grep -E "^(FAIL|EXC)" gpurun_out/a_misc.log | head -n 20
leg a_parity_r2 600 python -m pytest tests/test_parity_r2_gpu.py -x -q -s -m gpu
grep -E "rel-L2|b8_64|traj50|sample_log|Error|assert" gpurun_out/a_parity_r2.log | head -n 30
leg a_pytest_rest 600 python -m pytest tests -q -m gpu --deselect tests/test_parity_r2_gpu.py -x
leg a_vae 400 python scripts/gpu_vae_parity.py
tail -n 14 gpurun_out/a_vae.log
leg a_bench 600 python bench.py --steps 20 --warmup 3
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/a_bench.log") if l.startswith("{")][-1])
    json.dump(d, open("gpurun_out/a_bench.json", "w"), indent=1)
    pick = lambda r: {k: r.get(k) for k in ("value", "ms_per_step", "bank_build_ms", "step_launches", "finite")}
    print("bench B=1:", {k: d.get(k) for k in ("value", "ms_per_step", "launches_per_step", "bank_build_ms")},
          "steady", d.get("steady_state", {}).get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"),
          "roof", d.get("roofline", {}).get("frac"))
    b8 = d.get("batch8") or {}
    print("bench B=8:", pick(b8), "steady", b8.get("steady_state", {}).get("ms_per_step"), "e2e", b8.get("e2e", {}).get("value"),
          "roof", (b8.get("roofline") or {}).get("frac"))
    print("gpu eager:", json.dumps(d.get("gpu_eager_baseline"))[:600])
    print("cpu:", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("no bench result:", e)
PY
echo "== never-run round-1 kernels (numerics only)"
for f in MDB_GEMM_PAIR_SPLITK "'MDB_GEMM_PAIR', '3'" MDB_GEMM_TMAST; do
  n=$(echo "$f" | tr -dc 'A-Z0-9_')
  leg "a_pending_$n" 150 python scripts/gpu_diag.py --group pending --pending-filter "$f"
  grep -E "^(FAIL|EXC)" "gpurun_out/a_pending_$n.log" | head -n 6
done
MDB_ATTN=4 leg a_pending_attn4 120 python scripts/gpu_diag.py --group attn

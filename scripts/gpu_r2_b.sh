#!/bin/bash
# Round-2 GPU call B: the full default bench line, bank-build overlap, and the timing of the round-1 kernels whose
# numerics were green in call A (promote or delete), plus the ncu launch list of one B=1 step.
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e"
bench() {  # bench <name> <timeout> <bench args...>   (environment switches are inherited from the caller)
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout -k 10 "$t" python bench.py "$@" > "gpurun_out/$name.json" 2> "gpurun_out/$name.err"
  echo "$name rc=$? ($(( $(date +%s) - t0 )) s) $(python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    ss = d.get("steady_state", {})
    print(f"value={d['value']:.1f} ms/step={d['ms_per_step']:.3f} steady={ss.get('ms_per_step', 0):.3f} bank_ms={d.get('bank_build_ms', 0):.1f} "
          f"launches/step={d.get('launches_per_step')} finite={d['finite']} fp={d.get('x_final_fingerprint')}")
except Exception as e:
    print("no result:", e)
PY
)"
}
bench b_full 900 --steps 20 --warmup 3
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/b_full.json"))
    print("  e2e", d.get("e2e", {}).get("value"), "roof", d.get("roofline", {}).get("frac"))
    b8 = d.get("batch8") or {}
    print("  B=8: value", b8.get("value"), "ms/step", b8.get("ms_per_step"), "steady", b8.get("steady_state", {}).get("ms_per_step"),
          "e2e", b8.get("e2e", {}).get("value"), "roof", (b8.get("roofline") or {}).get("frac"), "fp", b8.get("x_final_fingerprint"))
    print("  gpu eager:", json.dumps(d.get("gpu_eager_baseline"))[:700])
    print("  cpu:", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("no bench result:", e)
PY
bench b_b1_default 300 $Q --no-batch8
bench b_b1_overlap 300 $Q --no-batch8 --bank-overlap on
MDB_GEMM_PAIR_SPLITK=1 bench b_b1_pairs 300 $Q --no-batch8
MDB_ATTN=4 bench b_b1_attn4 300 $Q --no-batch8
bench b_b8_default 300 $Q --batch 8
MDB_GEMM_PAIR=3 bench b_b8_pair3 300 $Q --batch 8
MDB_GEMM_TMAST=1 bench b_b8_tmast 300 $Q --batch 8
MDB_ATTN=4 bench b_b8_attn4 300 $Q --batch 8
MDB_GEMM_PAIR=3 MDB_ATTN=4 bench b_b8_pair3_attn4 300 $Q --batch 8
echo "== ncu launch list of the B=1 run (one captured step)"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_b1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e --no-batch8 > gpurun_out/b_ncu.log 2>&1
echo "ncu rc=$? $(wc -l < gpurun_out/b_launches_b1.csv) lines"

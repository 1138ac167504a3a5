#!/bin/bash
# Round-2 GPU call C: the cleaned-up library (pair kernel / two-Q-tile attention by heuristics, no env switches) —
# full -m gpu suite, cold-weight deep-K microbench, bench with and without the TMA-store epilogue, GPU eager baseline.
mkdir -p gpurun_out
leg() {  # leg <name> <timeout> <cmd...>
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"
}
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e"
bench() {
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout -k 10 "$t" python bench.py "$@" > "gpurun_out/$name.json" 2> "gpurun_out/$name.err"
  echo "$name rc=$? ($(( $(date +%s) - t0 )) s) $(python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}.json"))
    def show(tag, r):
        ss = r.get("steady_state", {})
        return (f"{tag}: value={r['value']:.1f} ms/step={r['ms_per_step']:.3f} steady={ss.get('ms_per_step', 0):.3f} "
                f"bank_ms={r.get('bank_build_ms', 0):.1f} fp={r.get('x_final_fingerprint')}")
    print(show("B=%d" % d["config"]["frames_per_gpu"], d), "|", show("B=8", d["batch8"]) if d.get("batch8") else "")
except Exception as e:
    print("no result:", e)
PY
)"
}
leg c_pytest 900 python -m pytest tests -q -m gpu -x
grep -E "passed|failed|Error" gpurun_out/c_pytest.log | tail -n 5
leg c_deepk 400 python scripts/gpu_microbench.py deepk
cat gpurun_out/c_deepk.log | tail -n 100
bench c_default 400 $Q
bench c_tmast 400 $Q --tune tma_store=1
bench c_nopair 400 $Q --tune pair_min_tiles=1073741824
leg c_eager 400 python tests/torch_gpu_baseline.py --batch 1,8 --steps 5 --warmup 2 --algorithmic
tail -n 12 gpurun_out/c_eager.log | cut -c1-900

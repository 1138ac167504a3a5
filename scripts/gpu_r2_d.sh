#!/bin/bash
# Round-2 GPU call D: cleaned library with automatic tile / split-K choice — full -m gpu suite, the full bench line,
# ncu launch lists (B=1, B=8) and one `--set full` capture of the dominant kernels.
mkdir -p gpurun_out
leg() {  # leg <name> <timeout> <cmd...>
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s): $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"
}
leg d_pytest 1200 python -m pytest tests -q -m gpu -x
grep -E "passed|failed|Error|FAILED" gpurun_out/d_pytest.log | tail -n 6
timeout -k 10 900 python bench.py --steps 20 --warmup 3 > gpurun_out/d_full.json 2> gpurun_out/d_full.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/d_full.json"))
    def show(tag, r):
        ss = r.get("steady_state", {})
        print(f"  {tag}: value={r['value']:.1f} ms/step={r['ms_per_step']:.3f} steady={ss.get('ms_per_step', 0):.3f} bank_ms={r.get('bank_build_ms', 0):.1f} "
              f"e2e={r.get('e2e', {}).get('value')} roof={(r.get('roofline') or {}).get('frac')} launches/step={r.get('step_launches', r.get('launches_per_step'))} fp={r.get('x_final_fingerprint')}")
    show("B=1", d)
    if d.get("batch8"):
        show("B=8", d["batch8"])
    print("  gpu eager:", json.dumps(d.get("gpu_eager_baseline"))[:900])
    print("  cpu:", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("no bench result:", e)
    print(open("gpurun_out/d_full.err").read()[-1500:])
PY
NQ="--steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_b1.csv \
  python bench.py $NQ --no-batch8 > gpurun_out/d_ncu_b1.log 2>&1
echo "ncu b1 rc=$? $(wc -l < gpurun_out/d_launches_b1.csv) lines"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_b8.csv \
  python bench.py $NQ --batch 8 > gpurun_out/d_ncu_b8.log 2>&1
echo "ncu b8 rc=$? $(wc -l < gpurun_out/d_launches_b8.csv) lines"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -o gpurun_out/d_targets -f \
  python scripts/gpu_ncu_targets.py 1 > gpurun_out/d_ncu_targets.log 2>&1
echo "ncu targets rc=$? $(ls -la gpurun_out/d_targets.ncu-rep 2>/dev/null | awk '{print $5}') bytes"

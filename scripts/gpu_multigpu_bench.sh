#!/bin/bash
# Round-2 GPU call E (2 GPUs): the multi-GPU bench line — frames of one sequence sharded over the ranks, bank timesteps dealt
# round-robin + one all-gather per slot row, probe-frame equality check, configs[3] sub-record (8 frames per GPU).
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/e_n$N.json 2> gpurun_out/e_n$N.err
echo "bench N=$N rc=$?"
tail -n 5 gpurun_out/e_n$N.err | cut -c1-400
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/e_n{n}.json"))
    def show(tag, r):
        ss = r.get("steady_state", {})
        print(f"  {tag}: value={r['value']:.1f} ms/step={r['ms_per_step']:.3f} steady={ss.get('ms_per_step', 0):.3f} ({ss.get('value', 0):.1f}/s) "
              f"bank_ms={r.get('bank_build_ms', 0):.1f} allgather_ms={r.get('allgather_ms')} e2e={r.get('e2e', {}).get('value')} fp={r.get('x_final_fingerprint')}")
    show("main (1 frame/GPU)", d)
    if d.get("config4"):
        show("config4 (8 frames/GPU)", d["config4"])
    print("  check:", d.get("multi_gpu_check"))
except Exception as e:
    print("no result:", e)
PY

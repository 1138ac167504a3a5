for c in 10 17 25 50; do
MDB_BANK_CHUNK=$c timeout 120 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/chunk_$c.json 2> gpurun_out/chunk_$c.err
python - $c <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/chunk_{c}.json").read().strip().splitlines()[-1])
    print("chunk",c,"value %.2f ms %.3f steady %.2f e2e %.2f"%(d["value"],d["ms_per_step"],d["steady_state"]["value"],d["e2e"]["value"]))
except Exception as e:
    print("chunk",c,"failed",e)
PY
done

"""profiles/traffic.json from an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum
per launch: sums the gemm_tc_kernel launches of ONE captured DDIM step (between two cfg_ddim_update launches)."""
import collections
import csv
import json
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1)
    d[r["Metric Name"]] = v * scale
launches = list(by_id.values())
ends = [i for i, l in enumerate(launches) if "cfg_ddim_update" in l["name"]]
step = launches[ends[-2] + 1: ends[-1] + 1]
out = {}
for fam in ("gemm_tc_kernel", "attn"):
    sel = [l for l in step if fam in l["name"]]
    rd = sum(l.get("dram__bytes_read.sum", 0) for l in sel)
    wr = sum(l.get("dram__bytes_write.sum", 0) for l in sel)
    t = sum(l.get("gpu__time_duration.sum", 0) for l in sel)
    out[fam if fam != "attn" else "attention"] = {
        "launches_per_step": len(sel), "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
        "dram_bytes_per_launch_avg": (rd + wr) / max(len(sel), 1), "kernel_seconds_per_step_ncu": t,
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                  "on `bench.py --steps 2 --warmup 1` (B=1 frame, paired cond/uncond batch), one captured step"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))

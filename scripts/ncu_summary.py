"""Condenses ncu outputs into the small text summaries committed under profiles/.

  python scripts/ncu_summary.py rep   gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_top_kernels.md
  python scripts/ncu_summary.py list  gpurun_out/launches.csv      > profiles/rNN_launch_list_step.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
]


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full --clock-control none: per-launch metrics ({path})\n")
    print("| kernel | grid | " + " | ".join(k for _, k in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for r in data:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void mdb::", "").replace("mdb::", "")
        cells = []
        for full, _ in KEYS:
            if full in hdr:
                i = hdr.index(full)
                v = r[i]
                try:
                    v = f"{float(v):.2f}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("n/a")
        print(f"| {name} | {r[hdr.index('Grid Size')]} | " + " | ".join(cells) + " |")


def launch_list(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    ends = [i for i, r in enumerate(rows) if "cfg_ddim_update" in r["Kernel Name"]]
    step = rows[ends[-2] + 1: ends[-1] + 1] if len(ends) >= 2 else rows
    tot, cnt = collections.Counter(), collections.Counter()
    for row in step:
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else v
        nm = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void mdb::", "").replace("mdb::", "")
        tot[nm] += v
        cnt[nm] += 1
    total = sum(tot.values())
    print(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache and serialised:\n"
          f"# compare SHARES, not absolutes) — the kernels of ONE captured DDIM step ({path})\n")
    print(f"one step: {sum(cnt.values())} kernel launches, {total / 1e3:.2f} ms summed kernel time\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---|---|---|---|")
    for k, v in tot.most_common():
        print(f"| {k[:80]} | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% | {v / cnt[k]:.2f} |")


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2])

"""Turns the output of scripts/gpu_pending_checks.sh (gpurun_out/pending_*.log, pending_*.json) into a decision per
opt-in switch and, with --write, into magicdance_b200/switch_defaults.json.  Runs on the CPU box after the gpurun
call has merged gpurun_out/ back.

A switch is enabled only if ALL of:
  * its numerics log ends with "<group>: 0 failing of N" (N > 0) and holds no FAIL / EXC line;
  * every bench line measured with it is finite and its final-latent fingerprint agrees with the default run of
    the same batch size to --fp-tol (relative; default 2e-3 — different kernels round differently, a wrong kernel
    is off by orders of magnitude more);
  * it is faster than the default run of the same batch size by more than --min-gain (default 1 %) in at least one
    configuration and not slower by more than --min-gain in any.

    python scripts/decide_defaults.py [--dir gpurun_out] [--write]
"""
import argparse
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# switch -> (numerics log, [(bench json, the default run it is compared with)], the environment it stands for)
FEATURES = {
    "MDB_GEMM_PAIR=3": ("pending_pairq.log", [("pending_b8_pair3.json", "pending_b8_default.json")], {"MDB_GEMM_PAIR": "3"}),
    "MDB_GEMM_PAIR=2": ("pending_pairp.log", [("pending_b8_pair2.json", "pending_b8_default.json")], {"MDB_GEMM_PAIR": "2"}),
    "MDB_GEMM_TMAST=1": ("pending_tmast.log", [("pending_b8_tmast.json", "pending_b8_default.json")], {"MDB_GEMM_TMAST": "1"}),
    "MDB_GEMM_PAIR_SPLITK=1": ("pending_pairs.log", [("pending_b1_pairs.json", "pending_b1_default.json")],
                               {"MDB_GEMM_PAIR_SPLITK": "1"}),
    "MDB_ATTN=4": ("pending_attn4.log", [("pending_b8_attn4.json", "pending_b8_default.json"),
                                         ("pending_b1_attn4.json", "pending_b1_default.json")], {"MDB_ATTN": "4"}),
    "MDB_GN_FUSED=1": ("pending_gnfused.log", [("pending_b1_gnfused.json", "pending_b1_default.json")], {"MDB_GN_FUSED": "1"}),
    "MDB_BANK_OVERLAP=1": (None, [("pending_b1_overlap.json", "pending_b1_default.json")], {"MDB_BANK_OVERLAP": "1"}),
}
# switches that claim the same dispatch slot: keep the fastest
EXCLUSIVE = [["MDB_GEMM_PAIR=3", "MDB_GEMM_PAIR=2"]]


def numerics_ok(path):
    """(ok, detail) from a scripts/gpu_diag.py log"""
    if path is None:
        return True, "no kernel of its own (host-side scheduling)"
    if not os.path.isfile(path):
        return False, "numerics log missing"
    with open(path) as f:
        lines = f.read().splitlines()
    bad = [ln for ln in lines if ln.startswith(("FAIL", "EXC"))]
    tail = next((ln for ln in reversed(lines) if re.search(r": \d+ failing of \d+", ln)), None)
    if tail is None:
        return False, "log has no summary line (hung or crashed)"
    m = re.search(r": (\d+) failing of (\d+)", tail)
    failing, total = int(m.group(1)), int(m.group(2))
    if bad or failing or total == 0:
        return False, f"{failing} failing of {total}" + (f"; first: {bad[0][:90]}" if bad else "")
    return True, f"{total} cases ok"


def load(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


def fp_close(a, b, tol):
    if not a or not b or len(a) != len(b):
        return False
    return all(abs(x - y) <= tol * max(abs(x), abs(y), 1e-6) for x, y in zip(a, b))


def judge(dirname, fp_tol, min_gain):
    report, enabled = [], {}
    gains = {}
    for name, (log, runs, env) in FEATURES.items():
        ok, why = numerics_ok(os.path.join(dirname, log) if log else None)
        notes, best, worst = [f"numerics: {why}"], None, None
        for run, base in runs:
            d, b = load(os.path.join(dirname, run)), load(os.path.join(dirname, base))
            if d is None or b is None:
                notes.append(f"{run}: missing ({'run' if d is None else 'default run'})")
                ok = False
                continue
            if not d.get("finite", False):
                notes.append(f"{run}: not finite")
                ok = False
                continue
            if not fp_close(d.get("x_final_fingerprint"), b.get("x_final_fingerprint"), fp_tol):
                notes.append(f"{run}: final latent differs from the default run {d.get('x_final_fingerprint')} vs "
                             f"{b.get('x_final_fingerprint')}")
                ok = False
                continue
            g = d["value"] / b["value"] - 1.0
            notes.append(f"{run}: {d['value']:.1f} vs {b['value']:.1f} {d.get('unit', '')} ({g * 100:+.1f} %)")
            best = g if best is None else max(best, g)
            worst = g if worst is None else min(worst, g)
        take = bool(ok and best is not None and best > min_gain and worst > -min_gain)
        gains[name] = best if best is not None else -1.0
        report.append((name, take, notes))
        if take:
            enabled[name] = env
    # the drop-in sampler's graph path has its own gate script (result check + both timings in one log)
    name, path = "MDB_DROPIN_GRAPH=1", os.path.join(dirname, "pending_dropin_graph.log")
    if os.path.isfile(path):
        with open(path) as f:
            text = f.read()
        t = {tag: float(v) for tag, v in re.findall(r"^(eager loop|graph replay)\s*: ([0-9.]+) ms/step", text, re.M)}
        ok = text.strip().endswith("OK") and "eager loop" in t and "graph replay" in t
        gain = (t["eager loop"] / t["graph replay"] - 1.0) if ok else -1.0
        take = ok and gain > min_gain
        report.append((name, take, [f"gate script: {'OK' if ok else 'not OK'}"] +
                       ([f"eager {t['eager loop']:.2f} ms/step, graph {t['graph replay']:.2f} ms/step ({gain * 100:+.1f} %)"] if ok else [])))
        gains[name] = gain
        if take:
            enabled[name] = {"MDB_DROPIN_GRAPH": "1"}
    else:
        report.append((name, False, ["gate log missing"]))
    for group in EXCLUSIVE:
        live = [n for n in group if n in enabled]
        for n in sorted(live, key=lambda x: -gains[x])[1:]:
            del enabled[n]
            report.append((n, False, [f"superseded by {sorted(live, key=lambda x: -gains[x])[0]}"]))
    return report, enabled


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=os.path.join(REPO, "gpurun_out"))
    ap.add_argument("--fp-tol", type=float, default=2e-3)
    ap.add_argument("--min-gain", type=float, default=0.01)
    ap.add_argument("--write", action="store_true", help="write magicdance_b200/switch_defaults.json")
    args = ap.parse_args()
    report, enabled = judge(args.dir, args.fp_tol, args.min_gain)
    for name, take, notes in report:
        print(f"{'ENABLE ' if take else 'keep off'} {name}")
        for n in notes:
            print(f"           {n}")
    cfg = {}
    for env in enabled.values():
        cfg.update(env)
    print("switch_defaults:", json.dumps(cfg, sort_keys=True))
    if args.write:
        path = os.path.join(REPO, "magicdance_b200", "switch_defaults.json")
        if cfg:
            with open(path, "w") as f:
                json.dump(cfg, f, indent=1, sort_keys=True)
            print("wrote", path)
        elif os.path.exists(path):
            os.remove(path)
            print("removed", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())

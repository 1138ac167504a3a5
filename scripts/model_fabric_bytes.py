"""MODEL, not a measurement: the bytes one SM pulls through the L2 -> SM fabric for every tensor-core GEMM / conv of one
denoise step, under the default launch choice and under the opt-in pair kernels, and the time that traffic alone would
take at the measured per-SM ingest rate (46 B/clk, B300_MICROARCH 'TMA service/SM').  CPU only.

The model was checked against three measured points of the default kernel: conv 32x32 640->640
at one frame 27 us predicted / 30 measured, conv 64x64 320->320 18.6 / 20-28, conv 64x64 at eight frames 128 / 117.
It ignores everything else that bounds a kernel (launch + prologue, epilogue, DRAM, the MMA itself), so it is a LOWER
bound per layer and only says where the operand traffic is the binding term and what the pair tiles would buy there.

    python scripts/model_fabric_bytes.py [--frames 1] [--md profiles/r01_model_fabric_bytes.md]
"""
import argparse
import os
import sys
from collections import Counter

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from magicdance_b200.engine import NetConfig, _auto_splits, block_plan  # noqa: E402

SM, BYTES_PER_CLK, GHZ = 148, 46.0, 1.9
A_TILE = 128 * 64 * 2  # one 128-row K chunk of A


def step_gemms(frames, latent=64):
    """Counter of (m, n, k, conv, geglu) over one DDIM step: UNet on the cond+uncond pair (2*frames samples), pose
    ControlNet on `frames` samples (engine.py: _res, _transformer, _run_block, controlnet)."""
    cfg = NetConfig()
    inp, mid, out = block_plan(cfg)
    cnt = Counter()

    def walk(blocks, batch, hw0, zero_convs):
        hw = hw0
        for blk in blocks:
            for kind, _, cin, cout in blk:
                m = batch * hw
                if kind == "res":
                    cnt[(m, cout, 9 * cin, True, False)] += 1
                    cnt[(m, cout, 9 * cout, True, False)] += 1
                    if cin != cout:
                        cnt[(m, cout, cin, False, False)] += 1
                elif kind == "attn":
                    c = cout
                    for n, k in ((c, c), (2 * c, c), (c, c), (c, c), (c, c), (c, 4 * c), (c, c)):  # pin, qk, o, q2, o2, ff2, pout
                        cnt[(m, n, k, False, False)] += 1
                    cnt[(c, m, c, False, False)] += 1            # V^T = Wv . X^T
                    cnt[(m, 8 * c, c, False, True)] += 1         # GEGLU
                elif kind == "down":
                    hw //= 4
                    cnt[(batch * hw, cout, 9 * cin, False, False)] += 1   # im2col + GEMM
                elif kind == "up":
                    hw *= 4
                    cnt[(batch * hw, cout, 9 * cin, True, False)] += 1
            if zero_convs and blk[-1][0] != "up":
                c = blk[-1][3]
                cnt[(batch * hw, c, c, False, False)] += 1
        return hw

    hw = walk(inp, 2 * frames, latent * latent, False)
    hw = walk([mid], 2 * frames, hw, False)
    walk(out, 2 * frames, hw, False)
    hw = walk(inp[1:], frames, latent * latent, True)   # ControlNet: conv_in is a direct conv
    walk([mid], frames, hw, True)
    return cnt


def default_choice(m, n, k, geglu):
    """(CTAs, bytes per CTA) of the default single-CTA kernel (mdb_gemm_f16 + engine._auto_splits)"""
    splits = 1 if geglu else _auto_splits(m, n, k)
    mt, chunks = -(-m // 128), k // 64
    if geglu:
        bn = 128
    elif n % 160 == 0:
        bn = 80 if mt * (n // 160) * splits < 100 else 160
    else:
        bn = 128
    ctas = mt * -(-n // bn) * splits
    return ctas, -(-chunks // splits) * (A_TILE + bn * 128), f"bn{bn} s{splits}"


def pair_choice(m, n, k, geglu):
    """the opt-in pair kernels: pairs (<= 148 CTAs, long K) else pairq (>= 128 tile-equivalents); None if neither takes it"""
    mt, chunks = -(-m // 128), k // 64
    if mt < 2:
        return None
    m_pairs = (mt + 1) // 2
    best = None
    if not geglu and chunks >= 16:                                   # gemm_pairs_kernel
        for w in (320, 160, 128):
            if (w != 128 and n % w) or (w == 128 and n % 160 == 0):
                continue
            c = 2 * m_pairs * -(-n // w)
            if c > 148:
                continue
            s = 4 if (c * 4 <= 128 and chunks >= 16) else (2 if (c * 2 <= 132 and chunks >= 8) else 1)
            cost = -(-chunks // s) * (A_TILE + 64 * w)
            if best is None or cost < best[1]:
                best = (c * s, cost, f"pairs bn{w} S{s}")
    if best is None:                                                  # gemm_pairq_kernel (persistent)
        if geglu:
            w = 256 if n % 256 == 0 else 0
        elif n % 320 == 0 and chunks >= 16:
            w = 320
        elif n % 256 == 0:
            w = 256
        elif n % 160 == 0:
            w = 160
        else:
            w = 128
        if w and 2 * m_pairs * -(-n // w) >= 128:
            tiles = m_pairs * -(-n // w)
            rounds = -(-tiles // 74)
            best = (148, rounds * chunks * (A_TILE + 64 * w), f"pairq bn{w} x{rounds}")
    return best


def us(ctas, bytes_per_cta):
    per_sm = -(-ctas // SM) * bytes_per_cta
    return per_sm / BYTES_PER_CLK / (GHZ * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--md", default=None)
    args = ap.parse_args()
    cnt = step_gemms(args.frames)
    rows, tot_d, tot_p = [], 0.0, 0.0
    for (m, n, k, conv, geglu), c in sorted(cnt.items(), key=lambda kv: -kv[1] * us(*default_choice(*kv[0][:3], kv[0][4])[:2])):
        dc = default_choice(m, n, k, geglu)
        pc = pair_choice(m, n, k, geglu)
        td = us(dc[0], dc[1])
        tp = us(pc[0], pc[1]) if pc else td
        tot_d += c * td
        tot_p += c * min(tp, td)
        rows.append((f"{m}x{n}x{k}{' conv' if conv else ''}{' geglu' if geglu else ''}", c, dc[2], td, pc[2] if pc else "-", tp))
    lines = [f"# MODEL (not measured): L2->SM operand traffic per SM, one DDIM step, {args.frames} frame(s) + CFG pair",
             "", "Generated by `scripts/model_fabric_bytes.py`; see its header for what the model is and is not.", "",
             f"{sum(cnt.values())} tensor-core launches per step.  Fabric-only time, default kernel: **{tot_d / 1e3:.2f} ms**; "
             f"with the opt-in pair kernels where they apply: **{tot_p / 1e3:.2f} ms**.", "",
             "| shape (M x N x K) | launches | default | fabric us | pair kernel | fabric us |", "|---|---|---|---|---|---|"]
    for name, c, dl, td, pl, tp in rows[:28]:
        lines.append(f"| {name} | {c} | {dl} | {td:.1f} | {pl} | {tp:.1f} |")
    text = "\n".join(lines) + "\n"
    print(text)
    if args.md:
        with open(args.md, "w") as f:
            f.write(text)


if __name__ == "__main__":
    main()

"""Per-kernel GPU time without profiler overhead: each case is captured `reps` times back to back
in a CUDA graph and the replay is timed with CUDA events (warm L2; launch latency hidden)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from magicdance_b200 import ops  # noqa: E402
from magicdance_b200.engine import pack_geglu  # noqa: E402

D = "cuda"
h = lambda *s: torch.randn(*s, device=D).half()
f = lambda *s: torch.randn(*s, device=D)


def timeit(name, fn, reps=20, flops=None, bytes_=None):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    extra = ""
    if flops:
        extra += f"  {flops / best / 1e6:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / best / 1e3:8.1f} GB/s"
    print(f"{name:58s} {best:9.2f} us{extra}", flush=True)


def gemm_case(m, n, k, splits=1, bias=True, res=True):
    a, w = h(m, k), h(n, k)
    b = f(n) if bias else None
    r = h(m, n) if res else None
    timeit(f"gemm m={m} n={n} k={k} splits={splits}", lambda: ops.gemm(a, w, bias=b, residual=r, splits=splits),
           flops=2.0 * m * n * k)


def conv_case(b, hh, ww, cin, cout, splits=1):
    x, w = h(b * hh * ww, cin), h(cout, 9 * cin)
    bias = f(cout)
    timeit(f"conv B={b} {hh}x{ww} {cin}->{cout} splits={splits}",
           lambda: ops.gemm(x, w, bias=bias, conv=(b, hh, ww, cin), splits=splits), flops=2.0 * b * hh * ww * cout * 9 * cin)


def attn_case(b, d, nq, n0, n1=0):
    c = 8 * d
    q, k0, vt0 = h(b * nq, c), h(b * n0, c), h(c, b * ((n0 + 7) // 8 * 8))
    kw = {}
    if n1:
        kw = dict(k1=h(n1, c), vt1=h(c, n1), n1=n1, kv1_batches=1, bank_batches=b)
    timeit(f"attention B={b} d={d} nq={nq} n0={n0} n1={n1}",
           lambda: ops.attention(q, k0, vt0, n0, heads=8, d=d, batch=b, nq=nq, ldv0_batch=(n0 + 7) // 8 * 8, **kw),
           flops=4.0 * b * 8 * nq * (n0 + n1) * d)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ops.ensure_device()
    if which in ("all", "gemm"):
        for m, n, k in ((128, 160, 64), (128, 160, 640), (128, 160, 2560), (128, 1280, 1280)):
            gemm_case(m, n, k, res=False, bias=False)
        for s in (1, 2):
            gemm_case(4096, 320, 320, s)
        for s in (1, 2, 4):
            gemm_case(1024, 640, 640, s)
        for s in (1, 5, 10):
            gemm_case(256, 1280, 1280, s)
        for s in (1, 4, 8, 16):
            gemm_case(64, 1280, 1280, s)
        gemm_case(4096, 320, 1280)
        gemm_case(8192, 320, 320)
        gemm_case(32768, 320, 320)
        x, (wp, bp) = h(4096, 320), pack_geglu(torch.randn(2560, 320), torch.randn(2560), D)
        timeit("geglu m=4096 c=320", lambda: ops.gemm(x, wp, bias=bp, epilogue=ops.EPI_GEGLU), flops=2.0 * 4096 * 2560 * 320)
    if which in ("all", "conv"):
        for s in (1, 2, 4):
            conv_case(1, 64, 64, 320, 320, s)
        conv_case(8, 64, 64, 320, 320)
        conv_case(1, 64, 64, 640, 320, 2)
        for s in (1, 2, 4):
            conv_case(1, 32, 32, 640, 640, s)
        conv_case(8, 32, 32, 640, 640)
        for s in (1, 4, 9):
            conv_case(1, 16, 16, 1280, 1280, s)
        for s in (1, 8, 16):
            conv_case(1, 8, 8, 1280, 1280, s)
        conv_case(2, 8, 8, 2560, 1280, 16)
        conv_case(8, 8, 8, 1280, 1280, 4)
    if which == "pair":
        # single-CTA 128 x BN tiles vs the persistent CTA-pair kernel (256 x BN, cta_group::2), shape by shape, at the cond+uncond batch of one frame (2) and of eight frames (16)
        big = 1 << 30
        for label, tune in (("single-CTA tiles", dict(pair_min_tiles=big)), ("pair kernel (forced)", dict(pair_min_tiles=1))):
            print(f"--- {label}", flush=True)
            with ops.tuning(**tune):
                for b in (2, 16):
                    conv_case(b, 64, 64, 320, 320)
                    conv_case(b, 64, 64, 640, 320)
                    conv_case(b, 32, 32, 640, 640)
                    conv_case(b, 32, 32, 1280, 640)
                    conv_case(b, 16, 16, 1280, 1280)
                    conv_case(b, 8, 8, 1280, 1280)
                    gemm_case(b * 4096, 320, 320)
                    gemm_case(b * 4096, 320, 1280)
                    gemm_case(b * 1024, 640, 640)
                    gemm_case(b * 1024, 640, 2560)
                    gemm_case(b * 256, 1280, 1280)
                    gemm_case(b * 256, 1280, 5120)
                    for hw, c in ((4096, 320), (1024, 640), (256, 1280)):
                        x, (wp, bp) = h(b * hw, c), pack_geglu(torch.randn(8 * c, c), torch.randn(8 * c), D)
                        timeit(f"geglu m={b * hw} c={c}", lambda: ops.gemm(x, wp, bias=bp, epilogue=ops.EPI_GEGLU),
                               flops=2.0 * b * hw * 8 * c * c)
        return
    if which == "gn":
        # GroupNorm paths by batch: mode 1 = stats (+ last-CTA fold) -> apply, mode 2 = single-launch cluster kernel
        for b in (2, 4, 8, 16, 25):
            for (hw, c1, c2) in ((4096, 320, 0), (4096, 640, 320), (1024, 640, 0), (1024, 1280, 640), (256, 1280, 0),
                                 (256, 1280, 1280), (64, 1280, 1280)):
                x1 = h(b * hw, c1)
                x2 = h(b * hw, c2) if c2 else None
                g_, b_ = f(c1 + c2), f(c1 + c2)
                for mode in (1, 2):
                    timeit(f"groupnorm B={b} hw={hw} c={c1}+{c2} mode={mode}",
                           lambda: ops.groupnorm(x1, g_, b_, batch=b, hw=hw, eps=1e-5, silu=True, x2=x2, mode=mode),
                           bytes_=3.0 * b * hw * (c1 + c2) * 2)
        return
    if which == "deepk":
        # the single-frame weight-streaming layers (cond+uncond batch of 2) with COLD weights, as inside a step (each
        # step streams 2.4 GB of weights through a 126 MB L2): eight weight copies are cycled; 80- vs 160-wide tiles
        # and the split-K factor
        def cold(name, m, n, k, conv, flops):
            ws = [h(n, k) for _ in range(max(2, int(300e6 / (2.0 * n * k)) + 1))]
            x = h(m, conv[3] if conv else k)
            bias, res = f(n), h(m, n)
            for bn_below, bnl in ((1 << 30, 80), (0, 160)):
                for sp in (1, 2, 4, 8):
                    if k // 64 < 4 * sp:
                        continue
                    state = {"i": 0}

                    def fn():
                        w = ws[state["i"] % len(ws)]
                        state["i"] += 1
                        ops.gemm(x, w, bias=bias, residual=res, splits=sp, **({"conv": conv} if conv else {}))
                    with ops.tuning(bn80_below=bn_below, pair_min_tiles=1 << 30):
                        timeit(f"{name} bn={bnl} splits={sp}", fn, reps=len(ws) * 2, flops=flops)
        for (hh, cin, cout) in ((16, 1280, 1280), (16, 2560, 1280), (8, 1280, 1280), (8, 2560, 1280), (32, 640, 640),
                                (32, 1280, 640), (64, 320, 320)):
            m = 2 * hh * hh
            cold(f"conv B=2 {hh}x{hh} {cin}->{cout}", m, cout, 9 * cin, (2, hh, hh, cin), 2.0 * m * cout * 9 * cin)
        for (m, n, k) in ((512, 1280, 5120), (512, 1280, 1280), (2048, 640, 2560), (2048, 640, 640), (8192, 320, 1280)):
            cold(f"gemm m={m} n={n} k={k}", m, n, k, None, 2.0 * m * n * k)
        return
    if which in ("all", "attn"):
        attn_case(1, 40, 4096, 4096)
        attn_case(1, 40, 4096, 4096, 4096)
        attn_case(8, 40, 4096, 4096, 4096)
        attn_case(1, 40, 4096, 77)
        attn_case(16, 80, 1024, 1024, 1024)
        with ops.tuning(attn40_2q_min_ctas=1 << 30):
            attn_case(16, 80, 1024, 1024, 1024)   # one-Q-tile kernel for comparison
            attn_case(8, 40, 4096, 4096, 4096)
        attn_case(1, 80, 1024, 1024, 1024)
        attn_case(1, 80, 1024, 77)
        attn_case(1, 160, 256, 256, 256)
        attn_case(1, 160, 64, 64, 64)
        attn_case(1, 160, 256, 77)
    if which in ("all", "misc"):
        for (b, hw, c1, c2) in ((1, 4096, 320, 0), (1, 4096, 640, 320), (1, 1024, 640, 0), (1, 256, 1280, 1280), (1, 64, 1280, 1280), (8, 4096, 320, 0)):
            x1 = h(b * hw, c1)
            x2 = h(b * hw, c2) if c2 else None
            g_, b_ = f(c1 + c2), f(c1 + c2)
            timeit(f"groupnorm B={b} hw={hw} c={c1}+{c2}", lambda: ops.groupnorm(x1, g_, b_, batch=b, hw=hw, eps=1e-5, silu=True, x2=x2),
                   bytes_=2.0 * b * hw * (c1 + c2) * 3)
        for rows, c in ((4096, 320), (1024, 640), (256, 1280)):
            x, g_, b_ = h(rows, c), f(c), f(c)
            timeit(f"layernorm rows={rows} c={c}", lambda: ops.layernorm(x, g_, b_), bytes_=4.0 * rows * c)
        a, b2 = h(4096, 320), h(4096, 320)
        timeit("add 4096x320", lambda: ops.add(a, b2, batch=1), bytes_=6.0 * 4096 * 320)
        x = h(1024, 640)
        timeit("upsample 32x32x640", lambda: ops.upsample2x(x, batch=1, h=32, w=32, c=640))
        x = h(4096, 320)
        timeit("im2col 64x64x320", lambda: ops.im2col3x3(x, batch=1, h=64, w=64, c=320, stride=2))
        e, w, bb = f(1, 1280), h(20160, 1280), f(20160)
        timeit("skinny_linear 1x20160x1280", lambda: ops.skinny_linear(e, w, bb, silu_in=True), bytes_=2.0 * 20160 * 1280)
        for (hh, cin, cout, s) in ((512, 3, 16, 1), (512, 16, 16, 1), (512, 16, 32, 2), (256, 32, 32, 1), (256, 32, 96, 2),
                                   (128, 96, 96, 1), (128, 96, 256, 2)):
            x, w, bb = h(hh * hh, cin), h(cout, 9 * cin), f(cout)
            ho = (hh - 1) // s + 1
            timeit(f"direct conv {hh}x{hh} {cin}->{cout} s={s}",
                   lambda: ops.conv3x3_direct(x, w, bb, batch=1, h=hh, w=hh, cin=cin, cout=cout, stride=s, silu=True),
                   reps=5, flops=2.0 * ho * ho * cout * 9 * cin)
        x = h(4096, 4)
        w, bb = h(320, 36), f(320)
        timeit("direct conv_in 64x64 4->320", lambda: ops.conv3x3_direct(x, w, bb, batch=1, h=64, w=64, cin=4, cout=320))
        x = h(4096, 320)
        w, bb = h(4, 2880), f(4)
        timeit("direct conv_out 64x64 320->4", lambda: ops.conv3x3_direct(x, w, bb, batch=1, h=64, w=64, cin=320, cout=4))


if __name__ == "__main__":
    main()

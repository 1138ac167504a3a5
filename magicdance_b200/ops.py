"""Torch-tensor front end of the C ABI (include/magicdance_b200.h).

PyTorch is plumbing here: it owns device memory and the current CUDA stream; every function
below launches OUR kernels through ctypes.  No function has a PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

EPI_NONE, EPI_GEGLU = 0, 1
SKINNY_MAX_ROWS = 16  # mdb_skinny_linear_f32: rows per launch
_GEMM_DEBUG = os.environ.get("MDB_GEMM_DEBUG", "0") == "1"
TRACE = None  # set to a list to record (m, n, k, conv, epilogue, splits, k2) of every gemm() call (bench.py)


class tuning:
    """`with ops.tuning(pair_min_tiles=1):` — the library's launch heuristics (include/magicdance_b200.h
    mdb_set_tuning) for the duration of the block; tests use it to force a kernel variant onto small problems.
    Launches captured into CUDA graphs keep the variant they were captured with."""
    _KEYS = {"pair_min_tiles": _lib.TUNE_GEMM_PAIR_MIN_TILES,
             "attn40_2q_min_ctas": _lib.TUNE_ATTN40_2Q_MIN_CTAS, "bn80_below": _lib.TUNE_GEMM_BN80_BELOW}

    def __init__(self, **kw):
        self.want = {self._KEYS[k]: int(v) for k, v in kw.items() if v is not None}

    def __enter__(self):
        lib = _lib.load()
        self.old = {k: int(lib.mdb_get_tuning(k)) for k in self.want}
        for k, v in self.want.items():
            _lib.check(lib.mdb_set_tuning(k, v), "set_tuning")
        return self

    def __exit__(self, *a):
        lib = _lib.load()
        for k, v in self.old.items():
            lib.mdb_set_tuning(k, v)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name):
    if not t.is_cuda:
        raise RuntimeError(f"magicdance_b200: {name} must be a CUDA tensor (no CPU fallback exists for the hot path)")
    if t.dtype != dtype:
        raise RuntimeError(f"magicdance_b200: {name} must be {dtype}, got {t.dtype}")


def require_cuda(device):
    """The single place that decides where the networks may live: an sm_100 CUDA device, nothing else."""
    if torch.device(device).type != "cuda":
        raise RuntimeError("magicdance_b200: the networks run only on an sm_100 CUDA device — move the model to "
                           "the GPU first (there is no CPU/PyTorch fallback for the hot path)")


def ensure_device():
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("magicdance_b200: no CUDA device visible; the hot path has no CPU fallback")
    _lib.check(lib.mdb_device_check(), "device_check")
    return lib


def launch_count() -> int:
    return int(_lib.load().mdb_launch_count())


_ws_cache = {}


_ws_tag = 0  # scratch buffers are per "lane": concurrent streams must not share split-K / GroupNorm scratch


class workspace_lane:
    """`with ops.workspace_lane(1):` — kernels issued inside use their own scratch buffers, so a second
    CUDA stream can run another branch of the step concurrently (pipeline.GraphedDenoiser)."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        global _ws_tag
        self.prev, _ws_tag = _ws_tag, self.tag

    def __exit__(self, *a):
        global _ws_tag
        _ws_tag = self.prev


def current_lane():
    return _ws_tag


_ws_retired = []  # outgrown buffers stay alive: captured CUDA graphs hold raw pointers into them


def _workspace(key, numel, dtype, device, zero=False):
    """Per-(key, device, lane) scratch buffer, grown geometrically.  A buffer that is outgrown is RETIRED, never
    freed: launches captured into CUDA graphs keep its address, and a replay must not touch memory the caching
    allocator has handed to somebody else."""
    k = (key, device, _ws_tag)
    t = _ws_cache.get(k)
    if t is None or t.numel() < numel or t.dtype != dtype:
        if t is not None:
            _ws_retired.append(t)
            numel = max(numel, 2 * t.numel())
        alloc = torch.zeros if zero else torch.empty
        t = alloc(max(numel, 1), dtype=dtype, device=device)
        _ws_cache[k] = t
    return t


def gemm(a, w, *, out=None, bias=None, bias_batch_stride=0, rows_per_batch=0, residual=None, epilogue=EPI_NONE,
         a2=None, conv=None, conv_stride=1, splits=0, m=None, ln_u=None, ln_eps=1e-5):
    """D = epilogue(A @ W^T).  a: [M, K1] fp16 (last dim contiguous, row stride arbitrary) or, with
    conv=(nb, h, w, c), an NHWC activation; w: [N, K] fp16; a2: optional second K-range source.
    splits: 0 = the library picks tile width and split-K (1/2/4/8, reduced inside a thread-block cluster),
    1 = no split, n = exactly n splits (a count other than 2, 4, 8 goes through an fp32 workspace).
    ln_u: LayerNorm over a's rows folded into the GEMM — w must be W diag(gamma), bias W beta (+ b), ln_u the row sums
    of w (engine.fold_layernorm); D = rstd_r (a w^T - mean_r ln_u) + bias.  Small grids only."""
    lib = _lib.load()
    _chk(a, torch.float16, "a")
    _chk(w, torch.float16, "w")
    g = _lib.GemmDesc()
    n, k = w.shape
    if conv is not None:
        nb, h, wd, c = conv
        assert conv_stride in (1, 2) and a.is_contiguous() and a.numel() == nb * h * wd * c
        m = nb * ((h - 1) // conv_stride + 1) * ((wd - 1) // conv_stride + 1)  # output pixels (3x3, pad 1)
        g.conv, g.nb, g.h, g.w, g.c = conv_stride, nb, h, wd, c  # descriptor: conv = 1 (stride 1) | 2 (stride 2)
        g.a, g.lda, g.k1 = a.data_ptr(), c, k
    else:
        assert a.dim() == 2 and a.stride(1) == 1
        m = a.shape[0] if m is None else m
        g.a, g.lda, g.k1 = a.data_ptr(), a.stride(0), a.shape[1]
        if a2 is not None:
            _chk(a2, torch.float16, "a2")
            assert a2.dim() == 2 and a2.stride(1) == 1 and a2.shape[0] == a.shape[0]
            g.a2, g.lda2 = a2.data_ptr(), a2.stride(0)
            assert a.shape[1] + a2.shape[1] == k
        else:
            assert a.shape[1] == k, (a.shape, w.shape)
    n_out = n // 2 if epilogue == EPI_GEGLU else n
    if out is None:
        out = torch.empty((m, n_out), dtype=torch.float16, device=a.device)
    _chk(out, torch.float16, "out")
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape[0] >= m and out.shape[1] == n_out, (out.shape, m, n_out)
    assert w.stride(1) == 1
    g.b, g.ldb = w.data_ptr(), w.stride(0)
    g.d, g.ldd = out.data_ptr(), out.stride(0)
    if bias is not None:
        _chk(bias, torch.float32, "bias")
        g.bias, g.bias_batch_stride, g.rows_per_batch = bias.data_ptr(), bias_batch_stride, rows_per_batch
    if residual is not None:
        _chk(residual, torch.float16, "residual")
        assert residual.dim() == 2 and residual.stride(1) == 1
        g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
    g.m, g.n, g.k, g.epilogue = m, n, k, epilogue
    if ln_u is not None:
        _chk(ln_u, torch.float32, "ln_u")
        assert conv is None and a2 is None and ln_u.numel() == n and ln_u.is_contiguous()
        g.ln_u, g.ln_eps = ln_u.data_ptr(), float(ln_eps)
        splits = 1
    if TRACE is not None:
        TRACE.append((m, n, k, (tuple(conv) + (conv_stride,)) if conv is not None else None, epilogue, splits,
                      a2.shape[1] if a2 is not None else 0))
    if splits > 1 and splits not in (2, 4, 8):
        ws = _workspace("splitk", splits * m * n, torch.float32, a.device)
        g.splits, g.splitk_ws = splits, ws.data_ptr()
    else:
        g.splits = splits
    if _GEMM_DEBUG:  # MDB_GEMM_DEBUG=1: name every GEMM before it runs and wait for it (pins down a hanging shape)
        import sys
        print(f"gemm m={m} n={n} k={k} conv={conv} epi={epilogue} splits={g.splits} a2={a2 is not None} lda={g.lda} "
              f"ldd={g.ldd} bias={bias is not None} bbs={g.bias_batch_stride} res={residual is not None}",
              file=sys.stderr, flush=True)
    _lib.check(lib.mdb_gemm_f16(C.byref(g), _stream()), "gemm_f16")
    if _GEMM_DEBUG:
        torch.cuda.synchronize()
    return out


def attention(q, k0, vt0, n0, *, heads, d, batch, nq, out=None, kv0_batches=None, ldv0_batch=None,
              k1=None, vt1=None, n1=0, kv1_batches=1, ldv1_batch=None, bank_batches=0, scale=None):
    """softmax([q k0^T | q k1^T] * scale) [v0 ; v1]; k*: [rows, heads*d] (row stride free), vt*: [heads*d, cols]."""
    lib = _lib.load()
    for t, nm in ((q, "q"), (k0, "k0"), (vt0, "vt0")):
        _chk(t, torch.float16, nm)
    a = _lib.AttnDesc()
    hd = heads * d
    if out is None:
        out = torch.empty((batch * nq, hd), dtype=torch.float16, device=q.device)
    a.q, a.ldq = q.data_ptr(), q.stride(0)
    a.k0, a.ldk0, a.vt0, a.ldvt0 = k0.data_ptr(), k0.stride(0), vt0.data_ptr(), vt0.stride(0)
    a.n0 = n0
    a.kv0_batches = batch if kv0_batches is None else kv0_batches
    a.ldv0_batch = n0 if ldv0_batch is None else ldv0_batch
    if n1 > 0:
        _chk(k1, torch.float16, "k1")
        _chk(vt1, torch.float16, "vt1")
        a.k1, a.ldk1, a.vt1, a.ldvt1 = k1.data_ptr(), k1.stride(0), vt1.data_ptr(), vt1.stride(0)
        a.n1, a.kv1_batches = n1, kv1_batches
        a.ldv1_batch = n1 if ldv1_batch is None else ldv1_batch
        a.bank_batches = bank_batches
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    a.batch, a.heads, a.d, a.nq = batch, heads, d, nq
    a.scale = float(d) ** -0.5 if scale is None else scale
    _lib.check(lib.mdb_attention_f16(C.byref(a), _stream()), "attention_f16")
    return out


GN_AUTO, GN_TWO_KERNELS, GN_CLUSTER = 0, 1, 2  # mdb_groupnorm_f16 `mode`
GN_DEFAULT_MODE = GN_AUTO


def groupnorm(x1, gamma, beta, *, batch, hw, eps, silu, x2=None, out=None, mode=None):
    """GroupNorm(32) [+SiLU] over [x1 | x2] channels; deterministic, pivot-shifted statistics (csrc/norm.cu)."""
    lib = _lib.load()
    _chk(x1, torch.float16, "x1")
    mode = GN_DEFAULT_MODE if mode is None else mode
    c1 = x1.shape[-1]
    c2 = 0 if x2 is None else x2.shape[-1]
    assert x1.is_contiguous() and (x2 is None or x2.is_contiguous())
    if out is None:
        out = torch.empty((batch * hw, c1 + c2), dtype=torch.float16, device=x1.device)
    # workspace of the two-kernel path (tickets must be zero at first use: _workspace(zero=True)); the single-launch
    # cluster path ignores it
    need = int(lib.mdb_groupnorm_ws_floats(c1 + c2, batch, hw))
    ws = _workspace("gn", need, torch.float32, x1.device, zero=True)
    _lib.check(lib.mdb_groupnorm_f16(x1.data_ptr(), c1, _ptr(x2), c2, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                     ws.data_ptr(), batch, hw, eps, int(silu), int(mode), _stream()), "groupnorm_f16")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    lib = _lib.load()
    _chk(x, torch.float16, "x")
    assert x.is_contiguous()
    rows, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.check(lib.mdb_layernorm_f16(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), rows, c, eps,
                                     _stream()), "layernorm_f16")
    return out


def conv3x3_direct(x, wt, bias, *, batch, h, w, cin, cout, stride=1, silu=False, residual=None, out=None):
    lib = _lib.load()
    _chk(x, torch.float16, "x")
    ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
    if out is None:
        out = torch.empty((batch * ho * wo, cout), dtype=torch.float16, device=x.device)
    _lib.check(lib.mdb_conv3x3_direct_f16(x.data_ptr(), wt.data_ptr(), _ptr(bias), _ptr(residual), out.data_ptr(), batch,
                                          h, w, cin, cout, stride, int(silu), _stream()), "conv3x3_direct_f16")
    return out


def im2col3x3(x, *, batch, h, w, c, stride, pad="same"):
    """pad='same': one halo pixel on every side (the UNet's convolutions); pad='br': one padding row / column
    at the bottom / right only (the VAE encoder's Downsample, model.py:82-84)"""
    lib = _lib.load()
    _chk(x, torch.float16, "x")
    if pad == "br":
        ho, wo = (h + 1 - 3) // stride + 1, (w + 1 - 3) // stride + 1
        col = torch.empty((batch * ho * wo, 9 * c), dtype=torch.float16, device=x.device)
        _lib.check(lib.mdb_im2col3x3_br_f16(x.data_ptr(), col.data_ptr(), batch, h, w, c, stride, _stream()),
                   "im2col3x3_br_f16")
        return col
    assert pad == "same"
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    col = torch.empty((batch * ho * wo, 9 * c), dtype=torch.float16, device=x.device)
    _lib.check(lib.mdb_im2col3x3_f16(x.data_ptr(), col.data_ptr(), batch, h, w, c, stride, _stream()), "im2col3x3_f16")
    return col


def upsample2x(x, *, batch, h, w, c):
    lib = _lib.load()
    _chk(x, torch.float16, "x")
    y = torch.empty((batch * 4 * h * w, c), dtype=torch.float16, device=x.device)
    _lib.check(lib.mdb_upsample2x_f16(x.data_ptr(), y.data_ptr(), batch, h, w, c, _stream()), "upsample2x_f16")
    return y


def add(a, b, *, batch, b_batches=None, out=None):
    """out = a + b (out may alias a)."""
    lib = _lib.load()
    _chk(a, torch.float16, "a")
    _chk(b, torch.float16, "b")
    assert a.is_contiguous() and b.is_contiguous()
    if out is None:
        out = torch.empty_like(a)
    n_per = a.numel() // batch
    bb = batch if b_batches is None else b_batches
    assert b.numel() == n_per * bb
    _lib.check(lib.mdb_add_f16(a.data_ptr(), b.data_ptr(), out.data_ptr(), n_per, batch, bb, _stream()), "add_f16")
    return out


def timestep_embedding(t, dim, rows=None):
    """rows > len(t): row b uses t[b % len(t)] (one timestep for the whole batch, or the cond | uncond pair)"""
    lib = _lib.load()
    _chk(t, torch.int64, "t")
    assert t.dim() == 1 and t.is_contiguous()
    rows = t.shape[0] if rows is None else rows
    assert rows % t.shape[0] == 0
    out = torch.empty((rows, dim), dtype=torch.float32, device=t.device)
    _lib.check(lib.mdb_timestep_embedding_f32(t.data_ptr(), t.shape[0], out.data_ptr(), rows, dim, _stream()),
               "timestep_embedding_f32")
    return out


def skinny_linear(x, w, bias, *, silu_in=False, silu_out=False):
    lib = _lib.load()
    _chk(x, torch.float32, "x")
    _chk(w, torch.float16, "w")
    rows, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k and x.is_contiguous() and w.is_contiguous()
    out = torch.empty((rows, n), dtype=torch.float32, device=x.device)
    # the kernel keeps at most 16 activation rows in registers per weight row; taller inputs (more than eight
    # frames per batch, or more than 16 timesteps per bank-build chunk) go through it 16 rows at a time
    for r0 in range(0, rows, SKINNY_MAX_ROWS):
        r = min(SKINNY_MAX_ROWS, rows - r0)
        _lib.check(lib.mdb_skinny_linear_f32(x[r0:].data_ptr(), w.data_ptr(), _ptr(bias), out[r0:].data_ptr(), r, n, k,
                                             int(silu_in), int(silu_out), _stream()), "skinny_linear_f32")
    return out


def softmax_rows(x, scale=1.0):
    """in place: x[r, :] <- softmax(scale * x[r, :]) (fp16 storage, fp32 arithmetic)"""
    lib = _lib.load()
    _chk(x, torch.float16, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    _lib.check(lib.mdb_softmax_rows_f16(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], float(scale), _stream()),
               "softmax_rows_f16")
    return x


def nchw_f32_to_nhwc_f16(x, out=None, copies=1):
    """copies > 1: the result holds the batch `copies` times over ([copies*B*H*W, C])"""
    lib = _lib.load()
    _chk(x, torch.float32, "x")
    x = x.contiguous()
    b, c, h, w = x.shape
    y = torch.empty((copies * b * h * w, c), dtype=torch.float16, device=x.device) if out is None else out
    _lib.check(lib.mdb_nchw_f32_to_nhwc_f16(x.data_ptr(), y.data_ptr(), b, c, h, w, copies, _stream()),
               "nchw_f32_to_nhwc_f16")
    return y


def nhwc_f16_to_nchw_f32(x, *, batch, c, h, w, out=None):
    lib = _lib.load()
    _chk(x, torch.float16, "x")
    if out is None:
        out = torch.empty((batch, c, h, w), dtype=torch.float32, device=x.device)
    _lib.check(lib.mdb_nhwc_f16_to_nchw_f32(x.data_ptr(), out.data_ptr(), batch, c, h, w, _stream()),
               "nhwc_f16_to_nchw_f32")
    return out


def cfg_ddim_update(x, eps_c, eps_u, coef, noise=None, x_prev=None, pred_x0=None, update_x=False):
    """coef: device fp32[6] = {scale, sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma, sqrt(1-a_t)};
    update_x: x is overwritten with x_prev as well (the chain advances in place)"""
    lib = _lib.load()
    for t, nm in ((x, "x"), (eps_c, "eps_c"), (eps_u, "eps_u"), (coef, "coef")):
        _chk(t, torch.float32, nm)
    if x_prev is None:
        x_prev = torch.empty_like(x)
    if pred_x0 is None:
        pred_x0 = torch.empty_like(x)
    _lib.check(lib.mdb_cfg_ddim_update_f32(x.data_ptr(), eps_c.data_ptr(), eps_u.data_ptr(), _ptr(noise),
                                           x_prev.data_ptr(), pred_x0.data_ptr(), x.numel(), coef.data_ptr(),
                                           int(update_x), _stream()),
               "cfg_ddim_update_f32")
    return x_prev, pred_x0


"""TEST DOUBLES — plain PyTorch (CPU) stand-ins for the few `magicdance_b200.ops` entry points the VAE decoder
calls, interpreting the SAME packed layouts the kernels read (NHWC fp16 activations [B*H*W, C]; conv weights
[O][kh][kw][I]; 1x1 / linear weights [O, I]; V^T by swapped operands).  They exist so that the ORCHESTRATION in
magicdance_b200/vae.py (operand order, folds, shapes, strides) can be checked against the oracle without a GPU;
the kernels themselves are validated by tests/test_kernels_gpu.py.  Never imported by the product."""
import torch
import torch.nn.functional as F


def _h(t):
    return t.to(torch.float16)


def gemm(a, w, *, out=None, bias=None, bias_batch_stride=0, rows_per_batch=0, residual=None, epilogue=0, a2=None,
         conv=None, conv_stride=1, splits=1, m=None, ln_u=None, ln_eps=1e-5):
    """D = epilogue(A @ W^T) exactly as include/magicdance_b200.h describes mdb_gemm_f16 (dual-source A, conv mode,
    per-batch bias rows, residual, GEGLU over [value | gate] blocks of 32 interleaved columns)."""
    assert a.dtype == torch.float16 and w.dtype == torch.float16
    n, k = w.shape
    assert k % 64 == 0, "K must be a multiple of 64"
    if conv is not None:
        b, h, ww, cin = conv
        assert a2 is None and k == 9 * cin and a.numel() == b * h * ww * cin
        x = a.float().reshape(b, h, ww, cin).permute(0, 3, 1, 2)
        wt = w.float().reshape(n, 3, 3, cin).permute(0, 3, 1, 2)
        y = F.conv2d(x, wt, None, padding=1, stride=conv_stride).permute(0, 2, 3, 1).reshape(-1, n)
    else:
        af = a.float() if a2 is None else torch.cat([a.float(), a2.float()], dim=1)
        if m is not None:
            af = af[:m]
        assert af.shape[1] == k, (af.shape, w.shape)
        y = af @ w.float().t()
        if ln_u is not None:  # LayerNorm folded into the GEMM: rstd_r (A W'^T - mean_r u); the caller's bias carries W beta
            mean = af.mean(dim=1, keepdim=True)
            rstd = torch.rsqrt(af.var(dim=1, unbiased=False, keepdim=True) + ln_eps)
            y = rstd * (y - mean * ln_u.reshape(1, n))
    rows = y.shape[0]
    if bias is not None:
        assert bias.dtype == torch.float32
        if bias_batch_stride:
            assert rows_per_batch > 0 and rows % rows_per_batch == 0 and bias.dim() == 2 and bias.stride(0) == bias_batch_stride
            y = (y.reshape(rows // rows_per_batch, rows_per_batch, n) + bias[:, None, :]).reshape(rows, n)
        else:
            y = y + bias.reshape(1, n)
    if epilogue == 1:  # GEGLU on interleaved [32 value | 32 gate] column blocks (engine.pack_geglu)
        assert residual is None and n % 64 == 0
        yb = y.reshape(rows, n // 64, 2, 32)
        y = (yb[:, :, 0] * F.gelu(yb[:, :, 1])).reshape(rows, n // 2)
    if residual is not None:
        y = y + residual.float()[:rows]
    y = _h(y)
    if out is not None:
        assert out.dtype == torch.float16 and out.shape[0] >= rows and out.shape[1] == y.shape[1]
        out[:rows].copy_(y)
        return out
    return y


def attention(q, k0, vt0, n0, *, heads, d, batch, nq, out=None, kv0_batches=None, ldv0_batch=None,
              k1=None, vt1=None, n1=0, kv1_batches=1, ldv1_batch=None, bank_batches=0, scale=None):
    """softmax([q k0^T | q k1^T] * scale) [v0 ; v1] with V given transposed (mdb_attention_f16)."""
    c = heads * d
    kv0_batches = batch if kv0_batches is None else kv0_batches
    ldv0 = n0 if ldv0_batch is None else ldv0_batch
    ldv1 = n1 if ldv1_batch is None else ldv1_batch
    scale = float(d) ** -0.5 if scale is None else scale
    res = []
    for b in range(batch):
        qq = q[b * nq:(b + 1) * nq, :c].float().reshape(nq, heads, d).transpose(0, 1)
        kb = b if kv0_batches > 1 else 0
        kk = k0[kb * n0:(kb + 1) * n0, :c].float()
        vv = vt0[:, kb * ldv0:kb * ldv0 + n0].float().t()
        if n1 and b < bank_batches:
            bb = b if kv1_batches > 1 else 0
            kk = torch.cat([kk, k1[bb * n1:(bb + 1) * n1, :c].float()], 0)
            vv = torch.cat([vv, vt1[:, bb * ldv1:bb * ldv1 + n1].float().t()], 0)
        kk = kk.reshape(-1, heads, d).transpose(0, 1)
        vv = vv.reshape(-1, heads, d).transpose(0, 1)
        o = torch.softmax((qq @ kk.transpose(1, 2)) * scale, dim=-1) @ vv
        res.append(o.transpose(0, 1).reshape(nq, c))
    y = _h(torch.cat(res, 0))
    if out is not None:
        out.copy_(y)
        return out
    return y


def conv3x3_direct(x, wt, bias, *, batch, h, w, cin, cout, stride=1, silu=False, residual=None, out=None):
    assert x.dtype == torch.float16 and tuple(x.shape) == (batch * h * w, cin) and tuple(wt.shape) == (cout, 9 * cin)
    xi = x.float().reshape(batch, h, w, cin).permute(0, 3, 1, 2)
    wk = wt.float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    y = F.conv2d(xi, wk, bias, padding=1, stride=stride)
    if silu:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1).reshape(-1, cout)
    if residual is not None:
        y = y + residual.float()
    return _h(y)


def groupnorm(x1, gamma, beta, *, batch, hw, eps, silu, x2=None, out=None, mode=0):
    assert x1.dtype == torch.float16 and gamma.dtype == torch.float32
    if x2 is not None:  # fused channel concat (cldm.py:104)
        x1 = torch.cat([x1, x2], dim=1)
    c = x1.shape[1]
    assert c % 32 == 0 and (c // 32 >= 8 or c // 32 == 4)
    y = F.group_norm(x1.float().reshape(batch, hw, c).permute(0, 2, 1), 32, gamma, beta, eps=eps)
    if silu:
        y = F.silu(y)
    return _h(y.permute(0, 2, 1).reshape(batch * hw, c))


def upsample2x(x, *, batch, h, w, c):
    y = x.reshape(batch, h, w, c).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    return y.reshape(batch * 4 * h * w, c).contiguous()


def softmax_rows(x, scale=1.0):
    assert x.dtype == torch.float16 and x.shape[1] % 8 == 0
    x.copy_(_h(torch.softmax(x.float() * scale, dim=1)))
    return x


def nchw_f32_to_nhwc_f16(x, out=None, copies=1):
    b, c, h, w = x.shape
    y = _h(x.permute(0, 2, 3, 1).reshape(b * h * w, c)).contiguous()
    return y if copies == 1 else y.repeat(copies, 1)


def nhwc_f16_to_nchw_f32(x, *, batch, c, h, w, out=None):
    return x.float().reshape(batch, h, w, c).permute(0, 3, 1, 2).contiguous()


def im2col3x3(x, *, batch, h, w, c, stride, pad="same"):
    xi = x.float().reshape(batch, h, w, c).permute(0, 3, 1, 2)
    if pad == "br":  # one padding row / column at the bottom / right only
        cols = F.unfold(F.pad(xi, (0, 1, 0, 1)), 3, padding=0, stride=stride)
        ho, wo = (h + 1 - 3) // stride + 1, (w + 1 - 3) // stride + 1
    else:
        cols = F.unfold(xi, 3, padding=1, stride=stride)                  # [b, c*9, L], channel-major
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    cols = cols.reshape(batch, c, 9, ho * wo).permute(0, 3, 2, 1)         # -> tap-major, channel-minor
    return _h(cols.reshape(batch * ho * wo, 9 * c)).contiguous()


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    return _h(F.layer_norm(x.float(), (x.shape[1],), gamma, beta, eps))


def add(a, b, *, batch, b_batches=None, out=None):
    n_per = a.numel() // batch
    bb = batch if b_batches is None else b_batches
    assert b.numel() == n_per * bb and bb in (1, batch)
    y = _h((a.float().reshape(batch, n_per) + b.float().reshape(bb, n_per)).reshape(a.shape))
    if out is not None:
        out.copy_(y)
        return out
    return y


def timestep_embedding(t, dim, rows=None):
    import math
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    if rows is not None and rows != t.shape[0]:
        t = t.repeat(rows // t.shape[0])  # row b uses t[b % len(t)]
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def skinny_linear(x, w, bias, *, silu_in=False, silu_out=False):
    assert x.dtype == torch.float32 and w.dtype == torch.float16
    y = (F.silu(x) if silu_in else x) @ w.float().t()
    if bias is not None:
        y = y + bias
    return F.silu(y) if silu_out else y


def ensure_device():
    return None


def cfg_ddim_update(x, eps_c, eps_u, coef, noise=None, x_prev=None, pred_x0=None, update_x=False):
    """ddim.py:605,617-645 with coef = {scale, sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma, sqrt(1-a_t)}"""
    scale, sa, sap, sdir, sigma, s1a = (float(v) for v in coef[:6])
    eps = eps_u + scale * (eps_c - eps_u)
    p0 = (x - s1a * eps) / sa
    xp = sap * p0 + sdir * eps
    if noise is not None:
        xp = xp + sigma * noise
    return xp, p0


def require_cuda(device):
    return None  # the stand-ins run wherever the tensors are

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
         conv=None, splits=1):
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and epilogue == 0 and a2 is None
    if conv is not None:
        b, h, ww, cin = conv
        o = w.shape[0]
        x = a.float().reshape(b, h, ww, cin).permute(0, 3, 1, 2)
        wt = w.float().reshape(o, 3, 3, cin).permute(0, 3, 1, 2)
        y = F.conv2d(x, wt, None, padding=1).permute(0, 2, 3, 1).reshape(b * h * ww, o)
    else:
        assert a.shape[1] == w.shape[1] and a.shape[1] % 64 == 0, "K must match and be a multiple of 64"
        y = a.float() @ w.float().t()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias_batch_stride == 0
        y = y + bias
    if residual is not None:
        y = y + residual.float()
    y = _h(y)
    if out is not None:
        assert out.shape == y.shape and out.dtype == torch.float16
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


def groupnorm(x1, gamma, beta, *, batch, hw, eps, silu, x2=None, out=None):
    assert x2 is None and x1.dtype == torch.float16 and gamma.dtype == torch.float32
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


def nchw_f32_to_nhwc_f16(x, out=None):
    b, c, h, w = x.shape
    return _h(x.permute(0, 2, 3, 1).reshape(b * h * w, c)).contiguous()


def nhwc_f16_to_nchw_f32(x, *, batch, c, h, w, out=None):
    return x.float().reshape(batch, h, w, c).permute(0, 3, 1, 2).contiguous()


def im2col3x3(x, *, batch, h, w, c, stride):
    xi = x.float().reshape(batch, h, w, c).permute(0, 3, 1, 2)
    cols = F.unfold(xi, 3, padding=1, stride=stride)                      # [b, c*9, L], channel-major
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    cols = cols.reshape(batch, c, 9, ho * wo).permute(0, 3, 2, 1)         # -> tap-major, channel-minor
    return _h(cols.reshape(batch * ho * wo, 9 * c)).contiguous()

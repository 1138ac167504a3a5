"""Numerics of every C-ABI kernel against a plain PyTorch fp32 reference of the same op (GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    from tests import kernel_cases as K
    return K.ALL_CASES


def _ids():
    from tests import kernel_cases as K
    return [f"{f.__name__}-{'-'.join(str(getattr(x, '__name__', x)) for x in a)}" for f, a in K.ALL_CASES]


@pytest.mark.parametrize("fn,args", _cases(), ids=_ids())
def test_kernel_matches_torch_fp32(fn, args):
    err, tol, desc = fn(*args)
    torch.cuda.synchronize()
    assert err <= tol, f"{desc}: error {err:.3e} > {tol:.1e}"


def test_native_library_is_loaded_and_counts_launches():
    from magicdance_b200 import ops, _lib
    import os
    ops.ensure_device()
    n0 = ops.launch_count()
    x = torch.randn(64, 320, device="cuda").half()
    g = torch.ones(320, device="cuda")
    ops.layernorm(x, g, torch.zeros_like(g))
    assert ops.launch_count() == n0 + 1
    with open(f"/proc/{os.getpid()}/maps") as f:
        assert "libmagicdance_b200.so" in f.read()


def test_cpu_tensors_are_rejected_loudly():
    from magicdance_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.layernorm(torch.randn(4, 320).half(), torch.ones(320), torch.zeros(320))

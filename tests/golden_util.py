"""Helpers shared by the oracle (CPU) and CUDA-parity (GPU) tests for reading tests/golden/*.npz."""
import os

import numpy as np
import torch

from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().reshape(-1).cpu()
    b = b.detach().double().reshape(-1).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def check_summary(gold, key, t: torch.Tensor, tol: float, what=""):
    """Compare tensor t with a golden stored as subsample + moments (oracle/synth.py:summarize)."""
    shape = tuple(int(v) for v in gold[key + "/shape"])
    assert tuple(t.shape) == shape, f"{what}{key}: shape {tuple(t.shape)} != golden {shape}"
    f = t.detach().float().reshape(-1).cpu()
    idx = synth.sample_indices(f.numel())
    got = f[idx]
    ref = torch.from_numpy(gold[key + "/sample"])
    err = rel_l2(got, ref)
    assert err <= tol, f"{what}{key}: rel-L2 of subsample {err:.3e} > {tol:.1e}"
    l2 = float(gold[key + "/moments"][2])
    got_l2 = float(f.double().norm())
    assert abs(got_l2 - l2) <= max(4 * tol, 1e-4) * l2 + 1e-6, f"{what}{key}: L2 norm {got_l2} vs golden {l2}"
    return err


def small32_inputs():
    inp = synth.synth_inputs(2, 32, seed=0, shared_reference=False)
    inp["t"] = torch.tensor([981, 441], dtype=torch.long)
    return inp


def full64_inputs():
    return synth.synth_inputs(1, 64, seed=0, shared_reference=True)


def grad16_inputs():
    """inputs of tests/golden/grad16.npz (oracle/make_golden_grad.py: grad_inputs)"""
    inp = synth.synth_inputs(2, 16, seed=5, shared_reference=False)
    g = torch.Generator().manual_seed(17)
    inp["x0"] = 0.9 * torch.randn(2, 4, 16, 16, generator=g)
    inp["noise"] = torch.randn(2, 4, 16, 16, generator=g)
    inp["t_train"] = torch.tensor([812, 97], dtype=torch.long)
    return inp


def grad_sample_positions(numel, n=16):
    if numel <= n:
        return np.arange(numel)
    return (np.arange(n, dtype=np.int64) * (numel - 1)) // (n - 1)

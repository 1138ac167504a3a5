"""Deterministic synthetic weights and inputs (benchmarks, smoke test, parity fixtures).

No checkpoint and no dataset ship with the reference (pretrained_weights/ and TikTok-v4/ are
placeholders), and every zero_module tensor (openaimodel.py:249-252,749; attention.py:357-361;
cldm.py:733-734) makes a freshly built model output exactly 0.  The parity fixtures therefore
use weights generated per state-dict KEY from (seed, crc32(key)), so the very same tensors can
be rebuilt on any box from the committed key->shape manifest without the reference.
"""
from __future__ import annotations

import json
import os
import zlib

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
MANIFEST = os.path.join(HERE, "state_manifest.json")  # key -> shape of the reference LDM state_dict

SCHEDULE_KEYS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2", "logvar",
)


def _gen(seed: int, key: str, device="cpu") -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    return g


def synth_tensor(key: str, shape, seed: int = 0, device="cpu") -> torch.Tensor:
    """device='cpu' is the reproducible stream the golden fixtures were made with; a CUDA device uses the
    GPU generator (same distributions, different values, ~100x faster) for benchmarks."""
    shape = tuple(shape)
    g = _gen(seed, key, device)
    rn = lambda: torch.randn(shape, generator=g, device=device)
    if len(shape) == 4:  # conv weight OIHW
        fan_in = shape[1] * shape[2] * shape[3]
        return rn() * (1.0 / fan_in) ** 0.5
    if len(shape) == 2:  # linear weight (out, in)
        return rn() * (1.0 / shape[1]) ** 0.5
    if len(shape) == 1:
        if key.endswith(".weight"):  # GroupNorm / LayerNorm scale
            return 1.0 + 0.1 * rn()
        return 0.05 * rn()  # any bias
    raise ValueError(f"unexpected parameter rank for {key}: {shape}")


def load_manifest(path: str = MANIFEST) -> dict:
    with open(path) as f:
        return json.load(f)


def synth_state_dict(manifest: dict | None = None, seed: int = 0, prefixes=None, device="cpu") -> dict:
    """name -> fp32 tensor for every network parameter in the manifest (schedule buffers are
    derived, not synthesised: see restatement.make_schedule)."""
    manifest = manifest or load_manifest()
    out = {}
    for key in sorted(manifest):
        if key in SCHEDULE_KEYS:
            continue
        if prefixes is not None and not key.startswith(tuple(prefixes)):
            continue
        out[key] = synth_tensor(key, manifest[key], seed, device)
    return out


def synth_inputs(batch: int, latent: int, seed: int = 0, shared_reference: bool = True, t_value: int = 981) -> dict:
    """Inputs of one apply_model call: x ~ N(0,1), reference latent
    ~ 0.8 N(0,1), sparse pose map in [0,1] at 8x the latent size, context ~ N(0,1) standing in
    for CLIP(""), one timestep for the whole batch."""
    g = _gen(seed, f"inputs/{batch}/{latent}")
    x = torch.randn(batch, 4, latent, latent, generator=g)
    if shared_reference:
        ref = (0.8 * torch.randn(1, 4, latent, latent, generator=g)).expand(batch, -1, -1, -1).contiguous()
    else:
        ref = 0.8 * torch.randn(batch, 4, latent, latent, generator=g)
    u = torch.rand(batch, 3, latent * 8, latent * 8, generator=g)
    v = torch.rand(batch, 3, latent * 8, latent * 8, generator=g)
    pose = torch.where(u > 0.97, v, torch.zeros_like(v))
    ctx = torch.randn(1, 77, 768, generator=g).expand(batch, -1, -1).contiguous()
    t = torch.full((batch,), t_value, dtype=torch.long)
    return {"x": x, "ref": ref, "pose": pose, "context": ctx, "t": t}


def sample_indices(numel: int, n: int = 4096) -> torch.Tensor:
    """Deterministic subsample positions used to store large tensors compactly."""
    if numel <= n:
        return torch.arange(numel)
    return torch.linspace(0, numel - 1, n).round().long()


def summarize(t: torch.Tensor, n: int = 4096) -> dict:
    f = t.detach().float().reshape(-1)
    return {
        "shape": list(t.shape),
        "sample": f[sample_indices(f.numel(), n)].clone(),
        "mean": float(f.mean()),
        "std": float(f.std()),
        "l2": float(f.norm()),
        "absmax": float(f.abs().max()),
    }

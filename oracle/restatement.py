"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's DDIM denoising hot path.

This file is the ORACLE for the CUDA path.  It is imported only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; the product
(magicdance_b200/, model_lib/) never imports it.

Parity status: PINNED against the reference's own code — tests/golden/*.npz were produced by
oracle/make_golden.py importing /root/reference (unmodified) under oracle/ref_shim.py, and
tests/test_oracle.py checks this restatement against them.  The reference ships no tests or
golden vectors of its own, so "the reference itself, run on CPU in fp32 with
seeded synthetic weights" is the strongest pin available.

The restatement is functional (flat state dict + key prefix, no nn.Module tree); every
function cites the reference lines it follows, paths relative to
/root/reference/model_lib/ControlNet/.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_NET_CFG = dict(  # models/cldm_v15_reference_only_pose.yaml:21-72
    in_channels=4, out_channels=4, hint_channels=3, model_channels=320,
    attention_resolutions=(4, 2, 1), num_res_blocks=2, channel_mult=(1, 2, 4, 4),
    num_heads=8, context_dim=768,
)


# ----------------------------------------------------------------------------- schedule
def make_schedule(timesteps=1000, linear_start=0.00085, linear_end=0.0120):
    """ldm/modules/diffusionmodules/util.py:21-28 ('linear') + ldm/models/diffusion/ddpm.py:120-133."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
    return {"betas": betas, "alphas_cumprod": alphas_cumprod}


def ddim_schedule(alphas_cumprod, num_ddim_steps=50, num_ddpm_steps=1000, eta=0.0):
    """util.py:45-73 (make_ddim_timesteps 'uniform', make_ddim_sampling_parameters)."""
    c = num_ddpm_steps // num_ddim_steps
    ts = np.asarray(list(range(0, num_ddpm_steps, c))) + 1
    a = alphas_cumprod[ts]
    a_prev = np.asarray([alphas_cumprod[0]] + alphas_cumprod[ts[:-1]].tolist())
    sig = eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
    return {"timesteps": ts, "alphas": a, "alphas_prev": a_prev, "sigmas": sig,
            "sqrt_one_minus_alphas": np.sqrt(1.0 - a)}


# ----------------------------------------------------------------------------- small ops
def timestep_embedding(t, dim, max_period=10000):
    """util.py:189-209: [cos | sin], freqs = exp(-ln(max_period) * i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _gn(sd, p, x, eps):
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def time_embed(sd, p, t, model_channels):
    """openaimodel.py:547-551 applied as in cldm.py:67-68."""
    e = timestep_embedding(t, model_channels)
    return _lin(sd, p + "time_embed.2", F.silu(_lin(sd, p + "time_embed.0", e)))


# ----------------------------------------------------------------------------- blocks
def resblock(sd, p, x, emb):
    """openaimodel.py:275-295 (no up/down, no scale-shift norm): GroupNorm32 eps 1e-5."""
    h = _conv(sd, p + "in_layers.2", F.silu(_gn(sd, p + "in_layers.0", x, 1e-5)))
    h = h + _lin(sd, p + "emb_layers.1", F.silu(emb))[:, :, None, None]
    h = _conv(sd, p + "out_layers.3", F.silu(_gn(sd, p + "out_layers.0", h, 1e-5)))
    if p + "skip_connection.weight" in sd:
        x = _conv(sd, p + "skip_connection", x, padding=0)
    return x + h


def attention(sd, p, x, context, heads):
    """attention.py:168-199 (vanilla CrossAttention; fp32 throughout here)."""
    q = _lin(sd, p + "to_q", x)
    k = _lin(sd, p + "to_k", context)
    v = _lin(sd, p + "to_v", context)
    b, n, c = q.shape
    d = c // heads
    q, k, v = (t.reshape(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    out = out.transpose(1, 2).reshape(b, n, c)
    return _lin(sd, p + "to_out.0", out)


def feed_forward(sd, p, x):
    """attention.py:50-77: GEGLU (exact erf GELU) then Linear."""
    a, gate = _lin(sd, p + "net.0.proj", x).chunk(2, dim=-1)
    return _lin(sd, p + "net.2", a * F.gelu(gate))


def transformer_block(sd, p, x, context, heads, mode, bank, bank_entry):
    """attention.py:278-320.  mode: 'plain' (uc / attention_mode None), 'write', 'read'."""
    n1 = _ln(sd, p + "norm1", x)
    if mode == "write":
        bank.append([n1])
        x = attention(sd, p + "attn1.", n1, n1, heads) + x
    elif mode == "read" and bank_entry is not None and len(bank_entry) > 0:
        x = attention(sd, p + "attn1.", n1, torch.cat([n1] + list(bank_entry), dim=1), heads) + x
    else:
        x = attention(sd, p + "attn1.", n1, n1, heads) + x
    x = attention(sd, p + "attn2.", _ln(sd, p + "norm2", x), context, heads) + x
    x = feed_forward(sd, p + "ff.", _ln(sd, p + "norm3", x)) + x
    return x


def spatial_transformer(sd, p, x, context, heads, mode, bank, bank_entry):
    """attention.py:366-385: GroupNorm eps 1e-6 (attention.py:89-90), 1x1 proj_in/out."""
    b, c, h, w = x.shape
    y = _conv(sd, p + "proj_in", _gn(sd, p + "norm", x, 1e-6), padding=0)
    y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
    y = transformer_block(sd, p + "transformer_blocks.0.", y, context, heads, mode, bank, bank_entry)
    y = y.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return _conv(sd, p + "proj_out", y, padding=0) + x


# ----------------------------------------------------------------------------- block map
def block_plan(cfg):
    """Layer list built the way openaimodel.py:562-750 / cldm.py:256-462 build it.
    Returns (input_blocks, middle, output_blocks); each block is a list of
    ('conv_in'|'res'|'attn'|'down'|'up', sub_index) in TimestepEmbedSequential order."""
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    inp = [[("conv_in", 0)]]
    ds = 1
    for level in range(len(mult)):
        for _ in range(nrb):
            blk = [("res", 0)]
            if ds in cfg["attention_resolutions"]:
                blk.append(("attn", 1))
            inp.append(blk)
        if level != len(mult) - 1:
            inp.append([("down", 0)])
            ds *= 2
    mid = [("res", 0), ("attn", 1), ("res", 2)]
    out = []
    for level in reversed(range(len(mult))):
        for i in range(nrb + 1):
            blk = [("res", 0)]
            if ds in cfg["attention_resolutions"]:
                blk.append(("attn", len(blk)))
            if level and i == nrb:
                blk.append(("up", len(blk)))
                ds //= 2
            out.append(blk)
    return inp, mid, out


def _run_block(sd, p, blk, h, emb, context, heads, mode, bank, attn_index):
    """openaimodel.py:79-108 (TimestepEmbedSequential dispatch)."""
    for kind, j in blk:
        q = f"{p}{j}."
        if kind == "conv_in":
            h = _conv(sd, q[:-1], h)
        elif kind == "res":
            h = resblock(sd, q, h, emb)
        elif kind == "attn":
            entry = bank[attn_index[0]] if mode == "read" else None
            h = spatial_transformer(sd, q, h, context, heads, mode, bank, entry)
            if mode == "read":
                attn_index[0] += 1
        elif kind == "down":  # openaimodel.py:178-180
            h = _conv(sd, q + "op", h, stride=2)
        elif kind == "up":  # openaimodel.py:129-139 (nearest x2 then conv)
            h = _conv(sd, q + "conv", F.interpolate(h, scale_factor=2, mode="nearest"))
    return h


# ----------------------------------------------------------------------------- the three nets
def unet_forward(sd, p, x, t, context, cfg=DEFAULT_NET_CFG, bank=None, pose_control=None, uc=False,
                 taps=None):
    """ControlledUnetModelAttnPose.forward, cldm.py:59-112.
    uc=True: plain SD UNet (no bank, no pose residuals, cldm.py:70-84); else 'read' mode."""
    inp, mid, out = block_plan(cfg)
    heads = cfg["num_heads"]
    emb = time_embed(sd, p, t, cfg["model_channels"])
    mode = "plain" if uc else "read"
    bank = None if uc else bank
    pose = None if (uc or pose_control is None) else list(pose_control)
    idx = [0]
    hs, h = [], x.float()
    for i, blk in enumerate(inp):
        h = _run_block(sd, f"{p}input_blocks.{i}.", blk, h, emb, context, heads, mode, bank, idx)
        hs.append(h)
        if taps is not None:
            taps.append(h)
    h = _run_block(sd, f"{p}middle_block.", mid, h, emb, context, heads, mode, bank, idx)
    if taps is not None:
        taps.append(h)
    if pose is not None:
        h = h + pose.pop()
    for i, blk in enumerate(out):
        skip = hs.pop()
        if pose is not None:
            skip = skip + pose.pop()
        h = torch.cat([h, skip], dim=1)
        h = _run_block(sd, f"{p}output_blocks.{i}.", blk, h, emb, context, heads, mode, bank, idx)
        if taps is not None:
            taps.append(h)
    h = _conv(sd, p + "out.2", F.silu(_gn(sd, p + "out.0", h, 1e-5)))
    return h


def appearance_forward(sd, p, x, t, context, cfg=DEFAULT_NET_CFG):
    """ControlNetReferenceOnly.forward in 'write' mode, cldm.py:469-497: a full UNet twin whose
    only product is the bank of norm1(x) hidden states (attention.py:287-298)."""
    inp, mid, out = block_plan(cfg)
    heads = cfg["num_heads"]
    emb = time_embed(sd, p, t, cfg["model_channels"])
    bank, idx = [], [0]
    hs, h = [], x.float()
    for i, blk in enumerate(inp):
        h = _run_block(sd, f"{p}input_blocks.{i}.", blk, h, emb, context, heads, "write", bank, idx)
        hs.append(h)
    h = _run_block(sd, f"{p}middle_block.", mid, h, emb, context, heads, "write", bank, idx)
    for i, blk in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, f"{p}output_blocks.{i}.", blk, h, emb, context, heads, "write", bank, idx)
    return bank


def hint_block(sd, p, hint):
    """cldm.py:599-615: 8 conv3x3 (+SiLU except after the last), strides 1,1,2,1,2,1,2,1."""
    h = hint.float()
    strides = (1, 1, 2, 1, 2, 1, 2, 1)
    for i, s in enumerate(strides):
        h = _conv(sd, f"{p}input_hint_block.{2 * i}", h, stride=s)
        if i != len(strides) - 1:
            h = F.silu(h)
    return h


def controlnet_forward(sd, p, x, hint, t, context, cfg=DEFAULT_NET_CFG):
    """ControlNet.forward, cldm.py:736-757: hint added once after input block 0; 13 zero convs."""
    inp, mid, _ = block_plan(cfg)
    heads = cfg["num_heads"]
    emb = time_embed(sd, p, t, cfg["model_channels"])
    guided = hint_block(sd, p, hint)
    outs, idx = [], [0]
    h = x.float()
    for i, blk in enumerate(inp):
        h = _run_block(sd, f"{p}input_blocks.{i}.", blk, h, emb, context, heads, "plain", None, idx)
        if guided is not None:
            h = h + guided
            guided = None
        outs.append(_conv(sd, f"{p}zero_convs.{i}.0", h, padding=0))
    h = _run_block(sd, f"{p}middle_block.", mid, h, emb, context, heads, "plain", None, idx)
    outs.append(_conv(sd, f"{p}middle_block_out.0", h, padding=0))
    return outs


# ----------------------------------------------------------------------------- glue + sampler
UNET = "model.diffusion_model."
APPEARANCE = "appearance_control_model."
POSE = "pose_control_model."


def apply_model(sd, x_noisy, t, context, pose_map, reference_image_noisy, uc=False, cfg=DEFAULT_NET_CFG,
                return_parts=False):
    """ControlLDMReferenceOnlyPose.apply_model, cldm.py:1099-1117 (c_crossattn_void absent)."""
    bank = []
    if reference_image_noisy is not None:
        bank = appearance_forward(sd, APPEARANCE, reference_image_noisy, t, context, cfg)
    pose_control = controlnet_forward(sd, POSE, x_noisy, pose_map, t, context, cfg)
    taps = [] if return_parts else None
    eps = unet_forward(sd, UNET, x_noisy, t, context, cfg, bank=bank, pose_control=pose_control, uc=uc, taps=taps)
    if return_parts:
        return eps, bank, pose_control, taps
    return eps


def p_sample_ddim(sd, x, t, index, context, pose_map, reference_latent, sched, scale=7.0, cfg=DEFAULT_NET_CFG):
    """DDIMSampler_ReferenceOnly.p_sample_ddim, ddim.py:518-645, 'controlnet is more important'
    branch (ddim.py:598-605) with wonoise=True (ddim.py:532-533), eta=0: the unconditional call
    gets the SAME cond (c, not uc) with reference None and uc=True."""
    e_c = apply_model(sd, x, t, context, pose_map, reference_latent, uc=False, cfg=cfg)
    e_u = apply_model(sd, x, t, context, pose_map, None, uc=True, cfg=cfg)
    e_t = e_u + scale * (e_c - e_u)
    return ddim_update(x, e_t, index, sched) + (e_c, e_u)


def ddim_update(x, e_t, index, sched):
    """ddim.py:617-645 (eps-parameterisation, sigma=0 => noise term vanishes)."""
    a_t = float(sched["alphas"][index])
    a_prev = float(sched["alphas_prev"][index])
    sigma = float(sched["sigmas"][index])
    s1m = float(sched["sqrt_one_minus_alphas"][index])
    pred_x0 = (x - s1m * e_t) / math.sqrt(a_t)
    dir_xt = math.sqrt(1.0 - a_prev - sigma ** 2) * e_t
    x_prev = math.sqrt(a_prev) * pred_x0 + dir_xt
    return x_prev, pred_x0


# ----------------------------------------------------------------------------- training caller (rows a15 / a16)
def q_sample(x0, t, noise, alphas_cumprod):
    """ddpm.py:356-359: sqrt(acp_t) x0 + sqrt(1 - acp_t) eps, per-sample t."""
    a = torch.as_tensor(alphas_cumprod[t.cpu().numpy()], dtype=x0.dtype, device=x0.device).reshape(-1, 1, 1, 1)
    return a.sqrt() * x0 + (1 - a).sqrt() * noise


def p_losses(sd, x0, t, noise, context, pose_map, reference_latent, cfg=DEFAULT_NET_CFG, x_noisy=None):
    """LatentDiffusionReferenceOnly.p_losses, ddpm.py:2165-2212, as train_tiktok.py:1212-1214 reaches it: 'eps'
    parameterisation, wonoise (the reference latent stays clean), l_simple_weight 1, logvar == 0 (not learned),
    original_elbo_weight 0 -> loss = mean_b mean_chw (eps_pred - eps)^2.  Differentiable: tensors of `sd` that require
    grad receive the gradients loss.backward() gives the reference's parameters (CheckpointFunction, util.py:118-187,
    recomputes the same values — tests/golden/grad16.npz holds both and their difference)."""
    if x_noisy is None:
        x_noisy = q_sample(x0, t, noise, make_schedule()["alphas_cumprod"])
    eps = apply_model(sd, x_noisy, t, context, pose_map, reference_latent, uc=False, cfg=cfg)
    loss_simple = ((eps - noise) ** 2).mean(dim=(1, 2, 3))
    return loss_simple.mean(), loss_simple, eps

"""TEST INFRASTRUCTURE — CPU restatement of the reference's first-stage autoencoder.

The step right after the hot path: `decode_first_stage` turns the final latent of every frame into an
image (2514.5 GFLOP per 512x512 frame), `encode_first_stage` turns the reference image into the latent the
appearance network sees (1116.7 GFLOP).  This file restates, as plain functional fp32 PyTorch over a
state-dict, exactly what the reference's `AutoencoderKL` does for the yaml's ddconfig
(models/cldm_v15_reference_only_pose.yaml:73-93: ch 128, ch_mult [1,2,4,4], 2 res blocks, no attention
except the middle block, z_channels 4, double_z, embed_dim 4).  It is the checker for the CUDA VAE path
that comes next; nothing in the product may import it.

Pinned: tests/test_oracle_vae.py compares it with tests/golden/vae*.npz, which
oracle/make_golden_vae.py produced by running the UNMODIFIED reference modules on the same synthetic
weights (magicdance_b200/vae_manifest.json lists their keys and shapes).

Reference lines (model_lib/ControlNet/ldm/...):
  modules/diffusionmodules/model.py:40-47    nonlinearity (swish), Normalize = GroupNorm(32, eps=1e-6)
  modules/diffusionmodules/model.py:50-66    Upsample: nearest x2 then conv3x3
  modules/diffusionmodules/model.py:68-88    Downsample: pad (0,1,0,1) then conv3x3 stride 2, no padding
  modules/diffusionmodules/model.py:90-149   ResnetBlock (temb is None here: temb_ch = 0)
  modules/diffusionmodules/model.py:152-203  AttnBlock: single head over all channels, scale c^-0.5
  modules/diffusionmodules/model.py:452-543  Encoder
  modules/diffusionmodules/model.py:546-652  Decoder
  models/autoencoder.py:82-91                encode = quant_conv(encoder(x)) -> moments; decode = decoder(post_quant_conv(z))
  modules/distributions/distributions.py:24-37  DiagonalGaussianDistribution: mean/logvar split, logvar clamp, sample
  models/diffusion/ddpm.py:1935-1942, 2100-2112  get_first_stage_encoding (x scale_factor), decode_first_stage (z / scale_factor)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

PREFIX = "first_stage_model."
CH, CH_MULT, NUM_RES_BLOCKS, Z_CHANNELS = 128, (1, 2, 4, 4), 2, 4
SCALE_FACTOR = 0.18215  # yaml:9 scale_factor


def _p(sd, name):
    return sd[PREFIX + name].float()


def swish(x):  # model.py:40-42
    return x * torch.sigmoid(x)


def group_norm(sd, name, x):  # model.py:45-46: 32 groups, eps 1e-6, affine
    return F.group_norm(x, 32, _p(sd, name + ".weight"), _p(sd, name + ".bias"), eps=1e-6)


def conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, _p(sd, name + ".weight"), _p(sd, name + ".bias"), stride=stride, padding=padding)


def resnet_block(sd, name, x):
    """model.py:129-149 with temb=None, dropout 0: conv2(swish(norm2(conv1(swish(norm1(x)))))) + shortcut(x);
    the shortcut is a 1x1 `nin_shortcut` conv when the channel count changes (model.py:119-127)."""
    h = conv(sd, name + ".conv1", swish(group_norm(sd, name + ".norm1", x)))
    h = conv(sd, name + ".conv2", swish(group_norm(sd, name + ".norm2", h)))
    if PREFIX + name + ".nin_shortcut.weight" in sd:
        x = conv(sd, name + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd, name, x):
    """model.py:179-203: q,k,v = 1x1 convs of GroupNorm(x); softmax over keys of q.k * c^-0.5 with ONE head
    spanning all c channels; proj_out 1x1; residual."""
    h = group_norm(sd, name + ".norm", x)
    q = conv(sd, name + ".q", h, padding=0)
    k = conv(sd, name + ".k", h, padding=0)
    v = conv(sd, name + ".v", h, padding=0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)          # b, hw, c
    k = k.reshape(b, c, hh * ww)                           # b, c, hw
    w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + conv(sd, name + ".proj_out", h, padding=0)


def decoder(sd, z, taps=None):
    """model.py:619-652 (give_pre_end False, tanh_out False)."""
    n_res = len(CH_MULT)
    h = conv(sd, "decoder.conv_in", z)
    h = resnet_block(sd, "decoder.mid.block_1", h)
    h = attn_block(sd, "decoder.mid.attn_1", h)
    h = resnet_block(sd, "decoder.mid.block_2", h)
    if taps is not None:
        taps["mid"] = h
    for i_level in reversed(range(n_res)):
        for i_block in range(NUM_RES_BLOCKS + 1):
            h = resnet_block(sd, f"decoder.up.{i_level}.block.{i_block}", h)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # model.py:62
            h = conv(sd, f"decoder.up.{i_level}.upsample.conv", h)
        if taps is not None:
            taps[f"up{i_level}"] = h
    h = swish(group_norm(sd, "decoder.norm_out", h))
    return conv(sd, "decoder.conv_out", h)


def encoder(sd, x, taps=None):
    """model.py:518-543."""
    n_res = len(CH_MULT)
    h = conv(sd, "encoder.conv_in", x)
    for i_level in range(n_res):
        for i_block in range(NUM_RES_BLOCKS):
            h = resnet_block(sd, f"encoder.down.{i_level}.block.{i_block}", h)
        if i_level != n_res - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)  # model.py:82-83: asymmetric padding
            h = conv(sd, f"encoder.down.{i_level}.downsample.conv", h, stride=2, padding=0)
        if taps is not None:
            taps[f"down{i_level}"] = h
    h = resnet_block(sd, "encoder.mid.block_1", h)
    h = attn_block(sd, "encoder.mid.attn_1", h)
    h = resnet_block(sd, "encoder.mid.block_2", h)
    h = swish(group_norm(sd, "encoder.norm_out", h))
    return conv(sd, "encoder.conv_out", h)


def vae_decode(sd, z, taps=None):
    """autoencoder.py:88-91."""
    return decoder(sd, conv(sd, "post_quant_conv", z, padding=0), taps)


def vae_encode_moments(sd, x, taps=None):
    """autoencoder.py:82-86: the 8-channel moments; DiagonalGaussianDistribution splits them into mean and
    logvar (clamped to [-30, 20]), distributions.py:27-28."""
    return conv(sd, "quant_conv", encoder(sd, x, taps), padding=0)


def posterior_sample(moments, noise=None):
    """distributions.py:27-37: mean + exp(0.5 * clamp(logvar)) * noise (noise None -> the mode)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def decode_first_stage(sd, z, taps=None):
    """ddpm.py:2107-2108: the latent is divided by scale_factor first."""
    return vae_decode(sd, z / SCALE_FACTOR, taps)


def get_first_stage_encoding(moments, noise=None):
    """ddpm.py:1936-1942: scale_factor * posterior.sample()."""
    return SCALE_FACTOR * posterior_sample(moments, noise)


def vae_inputs(batch: int, latent: int, seed: int = 0):
    """Seeded inputs shared by oracle/make_golden_vae.py and the tests: a latent as the sampler hands it to
    decode_first_stage (N(0,1) scaled to the latent's natural range), an image in [-1, 1] at 8x the latent
    size, and the posterior noise."""
    g = torch.Generator(device="cpu").manual_seed(7919 * (seed + 1) + 100 * batch + latent)
    z = torch.randn(batch, 4, latent, latent, generator=g) * SCALE_FACTOR * 5.0
    img = torch.rand(batch, 3, latent * 8, latent * 8, generator=g) * 2.0 - 1.0
    noise = torch.randn(batch, 4, latent, latent, generator=g)
    return z, img, noise

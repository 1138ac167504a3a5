"""model_lib.ControlNet.ldm.modules.diffusionmodules.util (reference: util.py:21-73,189-209): the host-side
schedule helpers; timestep_embedding runs on the GPU kernel."""
import numpy as np
import torch

from magicdance_b200 import ops
from magicdance_b200.dropin.ddpm import extract_into_tensor  # noqa: F401
from magicdance_b200.pipeline import ddim_parameters, ddim_timesteps_uniform, linear_beta_schedule


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    if schedule != "linear":
        raise NotImplementedError("MagicPose uses the 'linear' schedule")
    return linear_beta_schedule(n_timestep, linear_start, linear_end)


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    if ddim_discr_method != "uniform":
        raise NotImplementedError("MagicPose uses the 'uniform' discretisation")
    return ddim_timesteps_uniform(num_ddim_timesteps, num_ddpm_timesteps)


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    return ddim_parameters(np.asarray(alphacums), ddim_timesteps, eta)


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    assert max_period == 10000 and not repeat_only
    return ops.timestep_embedding(timesteps.to(torch.int64), dim)


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)

"""Pins the CPU oracle (oracle/restatement.py) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py, run in the build container).  CPU only, fp32."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import synth
from tests import golden_util as G

TOL = 2e-4  # fp32 vs fp32, different op ordering only


@pytest.fixture(scope="module")
def weights():
    torch.set_grad_enabled(False)
    return synth.synth_state_dict(seed=0)


def test_manifest_matches_survey_counts():
    m = synth.load_manifest()
    assert sum(k.startswith("model.diffusion_model.") for k in m) == 686
    assert sum(k.startswith("appearance_control_model.") for k in m) == 698
    assert sum(k.startswith("pose_control_model.") for k in m) == 340
    for k in synth.SCHEDULE_KEYS:
        assert m[k] == [1000]


def test_schedule_matches_reference_buffers():
    g = G.load("full64")
    sched = R.make_schedule()
    np.testing.assert_allclose(sched["alphas_cumprod"], g["full64/alphas_cumprod"], rtol=2e-6)
    d = R.ddim_schedule(sched["alphas_cumprod"].astype(np.float32).astype(np.float64))
    assert list(d["timesteps"]) == list(g["full64/ddim_timesteps"])
    np.testing.assert_allclose(d["alphas"], g["full64/ddim_alphas"], rtol=1e-6)
    np.testing.assert_allclose(d["alphas_prev"], g["full64/ddim_alphas_prev"], rtol=1e-6)


def test_block_plan_shape():
    inp, mid, out = R.block_plan(R.DEFAULT_NET_CFG)
    assert len(inp) == 12 and len(out) == 12 and len(mid) == 3
    n_attn = sum(k == "attn" for blk in inp + [mid] + out for k, _ in blk)
    assert n_attn == 16  # 16 banks
    assert [k for k, _ in out[2]] == ["res", "up"] and [k for k, _ in out[5]] == ["res", "attn", "up"]


def test_small32_apply_model_matches_reference(weights):
    g = G.load("small32")
    inp = G.small32_inputs()
    eps_c, bank, pose, taps = R.apply_model(weights, inp["x"], inp["t"], inp["context"], inp["pose"], inp["ref"],
                                            uc=False, return_parts=True)
    assert len(bank) == int(g["small32/n_bank"]) == 16
    assert len(pose) == int(g["small32/n_pose"]) == 13
    assert len(taps) == int(g["small32/n_tap"]) == 25
    for i, b in enumerate(bank):
        G.check_summary(g, f"small32/bank{i}", b[0], TOL)
    for i, p in enumerate(pose):
        G.check_summary(g, f"small32/pose{i}", p, TOL)
    for i, a in enumerate(taps):
        G.check_summary(g, f"small32/tap{i}", a, TOL)
    assert G.rel_l2(eps_c, torch.from_numpy(g["small32/eps_c"])) <= TOL
    eps_u = R.apply_model(weights, inp["x"], inp["t"], inp["context"], inp["pose"], None, uc=True)
    assert G.rel_l2(eps_u, torch.from_numpy(g["small32/eps_u"])) <= TOL
    # parity must not be vacuous: the two branches differ and neither is ~0
    assert float(eps_c.abs().mean()) > 1e-2 and G.rel_l2(eps_c, eps_u) > 1e-2


def test_full64_sampler_step_matches_reference(weights):
    g = G.load("full64")
    inp = G.full64_inputs()
    sched = R.ddim_schedule(R.make_schedule()["alphas_cumprod"].astype(np.float32).astype(np.float64))
    x_prev, pred_x0, e_c, e_u = R.p_sample_ddim(weights, inp["x"], inp["t"], 49, inp["context"], inp["pose"],
                                                inp["ref"], sched, scale=7.0)
    assert G.rel_l2(e_c, torch.from_numpy(g["full64/eps_c"])) <= TOL
    assert G.rel_l2(e_u, torch.from_numpy(g["full64/eps_u"])) <= TOL
    assert G.rel_l2(x_prev, torch.from_numpy(g["full64/x_prev"])) <= TOL
    assert G.rel_l2(pred_x0, torch.from_numpy(g["full64/pred_x0"])) <= TOL

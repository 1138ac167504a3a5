"""The drop-in boundary: same dotted paths, constructor kwargs and state-dict keys as the
reference; no CPU fallback."""
import os

import pytest
import torch


def _full_manifest():
    """key -> shape of the reference LDM's state dict: the three networks + 13 schedule buffers recorded by
    oracle/make_golden.py, plus the 248 first_stage_model.* tensors recorded by oracle/make_golden_vae.py"""
    import json
    from magicdance_b200 import synth
    manifest = dict(synth.load_manifest())
    with open(os.path.join(os.path.dirname(synth.MANIFEST), "vae_manifest.json")) as f:
        manifest.update(json.load(f))
    return manifest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML = os.path.join(REPO, "model_lib", "ControlNet", "models", "cldm_v15_reference_only_pose.yaml")


@pytest.fixture(scope="module")
def model():
    from model_lib.ControlNet.cldm.model import create_model
    return create_model(YAML)


def test_yaml_targets_resolve_to_dropin_classes(model):
    import yaml
    cfg = yaml.safe_load(open(YAML))["model"]
    assert cfg["target"] == "model_lib.ControlNet.cldm.cldm.ControlLDMReferenceOnlyPose"
    from model_lib.ControlNet.cldm import cldm
    assert type(model) is cldm.ControlLDMReferenceOnlyPose
    assert type(model.model.diffusion_model) is cldm.ControlledUnetModelAttnPose
    assert type(model.appearance_control_model) is cldm.ControlNetReferenceOnly
    assert type(model.pose_control_model) is cldm.ControlNet
    # attributes the reference scripts set / read (test_tiktok.py:374-375,224; train_tiktok.py:798-822)
    model.sd_locked, model.only_mid_control = True, False
    assert model.channels == 4 and model.image_size == 64 and model.num_timesteps == 1000
    for name in ("input_blocks", "middle_block", "output_blocks", "out"):
        assert hasattr(model.model.diffusion_model, name)
    for name in ("input_blocks", "middle_block", "input_hint_block", "middle_block_out", "zero_convs"):
        assert hasattr(model.pose_control_model, name)


def test_state_dict_keys_and_shapes_equal_the_reference(model):
    from magicdance_b200 import synth
    manifest = _full_manifest()  # recorded from the unmodified reference
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert ours == manifest


def test_schedule_buffers_match_reference(model):
    import numpy as np
    from tests import golden_util as G
    g = G.load("full64")
    np.testing.assert_allclose(model.alphas_cumprod.numpy(), g["full64/alphas_cumprod"], rtol=2e-6)
    from model_lib.ControlNet.ldm.models.diffusion.ddim import DDIMSampler_ReferenceOnly
    s = DDIMSampler_ReferenceOnly(model)
    s.make_schedule(50, ddim_eta=0.0, verbose=False)
    assert list(s.ddim_timesteps) == list(g["full64/ddim_timesteps"])
    np.testing.assert_allclose(s.ddim_alphas, g["full64/ddim_alphas"], rtol=1e-6)


def test_cpu_model_fails_loudly(model):
    x = torch.zeros(1, 4, 64, 64)
    cond = {"c_concat": [torch.zeros(1, 3, 512, 512)], "c_crossattn": [torch.zeros(1, 77, 768)]}
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        model.apply_model(x, torch.zeros(1, dtype=torch.long), cond, x)


def test_training_entry_refuses_to_fake_gradients(model):
    x = torch.zeros(1, 4, 64, 64)
    cond = {"c_concat": [torch.zeros(1, 3, 512, 512)], "c_crossattn": [torch.zeros(1, 77, 768)],
            "image_control": [x], "wonoise": True}
    with torch.enable_grad(), pytest.raises(NotImplementedError, match="forward"):  # (other modules disable grad globally)
        model(x, cond)


def test_strict_load_of_a_reference_checkpoint_layout(model):
    from magicdance_b200 import synth
    manifest = _full_manifest()
    sd = {k: torch.zeros(v) for k, v in manifest.items()}  # a checkpoint with exactly the reference's keys
    schedule = {k: v.clone() for k, v in model.state_dict().items() if k in synth.SCHEDULE_KEYS}
    try:
        missing, unexpected = model.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
        assert model.model.diffusion_model._packed is None  # packed fp16 copies are invalidated by a load
    finally:  # the module-scoped model keeps its derived schedule buffers for the tests below
        model.load_state_dict(schedule, strict=False)
        model._test_loaded_seed = None


def test_dropin_apply_model_orchestration_matches_reference_small32(model, monkeypatch):
    """The drop-in module tree end to end on the CPU: reference-layout state dict -> strict load -> lazily packed
    networks -> engine -> apply_model(x, t, cond dict, reference latent), with the kernels replaced by the
    layout-faithful PyTorch stand-ins of tests/fake_ops.py.  Result vs the UNMODIFIED reference's golden eps."""
    from magicdance_b200 import ops, synth
    from tests import fake_ops, golden_util as G
    from tests.test_engine_cpu import _PATCHED
    for name in _PATCHED:
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    torch.set_grad_enabled(False)
    sd = synth.synth_state_dict(seed=0)
    own = model.state_dict()
    sd.update({k: own[k] for k in synth.SCHEDULE_KEYS})  # the 13 schedule buffers are derived, not synthesised
    sd.update({k: own[k] for k in own if k.startswith("first_stage_model.")})  # the VAE is not on this test's path
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    model._test_loaded_seed = 0
    g = G.load("small32")
    inp = G.small32_inputs()
    cond = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]]}
    eps_c = model.apply_model(inp["x"], inp["t"], cond, inp["ref"])
    eps_u = model.apply_model(inp["x"], inp["t"], cond, None, uc=True)
    e_c = G.rel_l2(eps_c, torch.from_numpy(g["small32/eps_c"]))
    e_u = G.rel_l2(eps_u, torch.from_numpy(g["small32/eps_u"]))
    assert e_c <= 5e-3 and e_u <= 5e-3, (e_c, e_u)
    # the sub-networks called on their own, as the reference's apply_model does (cldm.py:1105-1117)
    bank = []
    assert model.appearance_control_model(x=inp["ref"], hint=None, timesteps=inp["t"], context=inp["context"],
                                          attention_bank=bank, attention_mode="write", uc=False) == []
    assert len(bank) == 16 and all(isinstance(b, list) and b[0].dim() == 3 for b in bank)
    res = model.pose_control_model(x=inp["x"], hint=inp["pose"], timesteps=inp["t"], context=inp["context"])
    assert len(res) == 13
    for i, r in enumerate(res):
        G.check_summary(g, f"small32/pose{i}", r, 5e-3)


def test_sample_log_four_step_chain_matches_the_oracle(model, monkeypatch):
    """The sampler's host loop end to end on the CPU (kernels = layout-faithful test doubles): sample_log ->
    DDIMSampler_ReferenceOnly.sample -> ddim_sampling -> p_sample_ddim for a 4-step DDIM schedule (timesteps 751, 501, 251, 1)
    from a given x_T — schedule construction, index order, the per-timestep bank cache, the cached hint features, CFG
    with the 'controlnet is more important' branch — against the oracle's chain of p_sample_ddim (ddim.py:460-645)."""
    import numpy as np
    from magicdance_b200 import ops, synth
    from oracle import restatement as R
    from tests import fake_ops, golden_util as G
    from tests.test_engine_cpu import _PATCHED
    for name in _PATCHED + ("cfg_ddim_update",):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    torch.set_grad_enabled(False)
    try:
        sd = synth.synth_state_dict(seed=0)
        if getattr(model, "_test_loaded_seed", None) != 0:  # the apply_model test above leaves these weights loaded and packed
            own = model.state_dict()
            sd.update({k: own[k] for k in synth.SCHEDULE_KEYS})
            sd.update({k: own[k] for k in own if k.startswith("first_stage_model.")})
            model.load_state_dict(sd, strict=True)
            model._test_loaded_seed = 0
        inp = synth.synth_inputs(1, 16, seed=5, shared_reference=True)  # 128x128 image: the host loop is size-independent
        uc_ctx = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(9))  # must be ignored (ddim.py:599-604)
        cond = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "image_control": [inp["ref"]], "wonoise": True}
        uc = {"c_concat": [inp["pose"]], "c_crossattn": [uc_ctx]}
        seen = []
        x0, inter = model.sample_log(cond, 1, ddim=True, ddim_steps=4, eta=0.0, unconditional_guidance_scale=7.0,
                                     unconditional_conditioning=uc, x_T=inp["x"], callback=seen.append)
        assert seen == [0, 1, 2, 3] and len(inter["x_inter"]) >= 2
        # the caller's tensors reach the engine as they are: one text-K/V entry per network pass kind for the whole
        # chain (a fresh torch.cat copy per step would miss every step), one hint entry, one bank entry per timestep
        pipe = next(iter(model._mdb_pipelines.values()))
        mine = [k for k in model.engine()._ctx_cache if tuple(k[3][:3]) == (1, 77, 768)]  # (net, storage, version, shape)
        assert 1 <= len(mine) <= 4 and len(pipe._hint_cache) == 1 and len(pipe._bank_cache) == 4
        sched = R.ddim_schedule(R.make_schedule()["alphas_cumprod"].astype(np.float32).astype(np.float64), num_ddim_steps=4)
        assert [int(t) for t in sched["timesteps"]] == [1, 251, 501, 751]
        x = inp["x"]
        for index in (3, 2, 1, 0):
            t = torch.full((1,), int(sched["timesteps"][index]), dtype=torch.long)
            x = R.p_sample_ddim(sd, x, t, index, inp["context"], inp["pose"], inp["ref"], sched, scale=7.0)[0]
        err = G.rel_l2(x0, x)
        assert tuple(x0.shape) == (1, 4, 16, 16) and err <= 3e-2, err
    finally:
        model.__dict__.pop("_mdb_pipelines", None)
        torch.set_grad_enabled(True)


def test_batched_cfg_branch_matches_the_reference_golden(model, monkeypatch):
    """p_sample_ddim when the unconditional conditioning carries image_control too (ddim.py:539-566, every control_mode
    other than 'controlnet_important'): one apply_model over [unconditional ; conditional], both halves with bank and
    pose residuals and their own prompt — against the UNMODIFIED reference's golden (oracle/make_golden_r2.py cfgb)."""
    from magicdance_b200 import ops, synth
    from magicdance_b200.dropin.ddim import DDIMSampler_ReferenceOnly
    from tests import fake_ops, golden_util as G
    from tests.test_engine_cpu import _PATCHED
    for name in _PATCHED + ("cfg_ddim_update",):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    torch.set_grad_enabled(False)
    try:
        if getattr(model, "_test_loaded_seed", None) != 0:
            sd = synth.synth_state_dict(seed=0)
            own = model.state_dict()
            sd.update({k: own[k] for k in synth.SCHEDULE_KEYS})
            sd.update({k: own[k] for k in own if k.startswith("first_stage_model.")})
            model.load_state_dict(sd, strict=True)
            model._test_loaded_seed = 0
        g = G.load("cfgb32")
        inp = synth.synth_inputs(1, 32, seed=0, shared_reference=True)
        uc_ctx = torch.from_numpy(g["uc_context"])
        c = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "image_control": [inp["ref"]], "wonoise": True,
             "overlap_sampling": False}
        uc = {"c_concat": [inp["pose"]], "c_crossattn": [uc_ctx], "image_control": [inp["ref"]], "wonoise": True,
              "overlap_sampling": False}
        sampler = DDIMSampler_ReferenceOnly(model)
        sampler.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
        ts = torch.full((1,), int(sampler.ddim_timesteps[30]), dtype=torch.long)
        x_prev, pred_x0 = sampler.p_sample_ddim(inp["x"], c, ts, index=30, unconditional_guidance_scale=7.0,
                                                unconditional_conditioning=uc)
        assert G.rel_l2(x_prev, torch.from_numpy(g["x_prev"])) <= 2e-2
        assert G.rel_l2(pred_x0, torch.from_numpy(g["pred_x0"])) <= 2e-2
    finally:
        model.__dict__.pop("_mdb_pipelines", None)
        torch.set_grad_enabled(True)


def test_p_losses_forward_value_matches_the_oracle(model, monkeypatch):
    """The training caller's forward (ddpm.py:2165-2212, q_sample ddpm.py:356-359) under no_grad:
    x_noisy = sqrt(acp_t) x0 + sqrt(1 - acp_t) eps, apply_model with the CLEAN reference latent (wonoise), and
    loss = mean((eps_pred - eps)^2) (l_simple_weight 1, logvar 0, original_elbo_weight 0) — against the same arithmetic
    on the oracle's apply_model.  Gradients stay refused (test above); this pins the value and the loss_dict keys."""
    from magicdance_b200 import ops, synth
    from oracle import restatement as R
    from tests import fake_ops
    from tests.test_engine_cpu import _PATCHED
    for name in _PATCHED:
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    torch.set_grad_enabled(False)
    try:
        sd = synth.synth_state_dict(seed=0)
        if getattr(model, "_test_loaded_seed", None) != 0:
            own = model.state_dict()
            sd.update({k: own[k] for k in synth.SCHEDULE_KEYS})
            model.load_state_dict(sd, strict=True)
            model._test_loaded_seed = 0
        inp = synth.synth_inputs(2, 16, seed=11, shared_reference=False)
        g = torch.Generator().manual_seed(3)
        x0 = torch.randn(2, 4, 16, 16, generator=g) * 0.8
        noise = torch.randn(2, 4, 16, 16, generator=g)
        t = torch.tensor([700, 35], dtype=torch.long)
        cond = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "image_control": [inp["ref"]], "wonoise": True}
        model.eval()
        loss, ld = model.p_losses(x0, cond, t, noise=noise)
        assert set(ld) == {"val/loss_simple", "val/loss_vlb", "val/loss"}
        acp = R.make_schedule()["alphas_cumprod"]
        a = torch.tensor(acp[t.numpy()], dtype=torch.float32).reshape(2, 1, 1, 1)
        x_noisy = a.sqrt() * x0 + (1 - a).sqrt() * noise
        eps = R.apply_model(sd, x_noisy, t, inp["context"], inp["pose"], inp["ref"], uc=False)
        want = ((eps - noise) ** 2).mean(dim=(1, 2, 3)).mean()
        assert abs(float(loss) - float(want)) <= 2e-2 * float(want), (float(loss), float(want))
        assert abs(float(ld["val/loss_simple"]) - float(want)) <= 2e-2 * float(want)
    finally:
        torch.set_grad_enabled(True)


def test_autoencoder_dropin_has_the_reference_keys_and_decodes_through_the_test_doubles(monkeypatch):
    """magicdance_b200.dropin.autoencoder.AutoencoderKL (opt-in first_stage target): the reference's constructor
    kwargs (yaml:93-114), exactly its 248 state-dict keys / shapes (manifest recorded from the unmodified reference),
    strict load, loud failure on the CPU, and — through the CPU test doubles — decode(z / scale_factor) equal to the
    reference's golden image and encode(x).mode() equal to its golden moments' mean."""
    import json
    import os
    import numpy as np
    import pytest
    from magicdance_b200 import ops, synth, vae
    from magicdance_b200.dropin.autoencoder import AutoencoderKL, DiagonalGaussianDistribution
    from oracle import vae_restatement as V
    from tests import fake_ops
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "magicdance_b200", "vae_manifest.json")) as f:
        manifest = json.load(f)
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4, monitor="val/rec_loss")
    want = {k[len(vae.PREFIX):]: tuple(v) for k, v in manifest.items()}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == want and len(want) == 248
    sd = synth.synth_state_dict(manifest, seed=0)
    m.load_state_dict({k[len(vae.PREFIX):]: v for k, v in sd.items()}, strict=True)
    with pytest.raises(NotImplementedError):
        AutoencoderKL(ddconfig=dict(dd, ch=64), lossconfig=None, embed_dim=4)
    with pytest.raises(RuntimeError, match="no CPU"):
        m.decode(torch.zeros(1, 4, 16, 16))
    for name in ("gemm", "conv3x3_direct", "groupnorm", "upsample2x", "softmax_rows", "nchw_f32_to_nhwc_f16",
                 "nhwc_f16_to_nchw_f32", "im2col3x3"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    torch.set_grad_enabled(False)
    try:
        dec = vae.VaeDecoder.__new__(vae.VaeDecoder)
        dec.p = vae.PackedVaeDecoder(m._prefixed_state(), "cpu", scale_factor=1.0)   # what decoder_engine() packs
        enc = vae.VaeEncoder.__new__(vae.VaeEncoder)
        enc.p = vae.PackedVaeEncoder(m._prefixed_state(), "cpu")
        z, img, noise = V.vae_inputs(2, 16)
        gold = np.load(os.path.join(here, "tests", "golden", "vae16.npz"))
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
        assert rel(dec._decode(z / vae.SCALE_FACTOR), torch.from_numpy(gold["vae16/decoded"])) <= 5e-3
        post = DiagonalGaussianDistribution(enc._encode(img))
        gm = torch.from_numpy(gold["vae16/moments"])
        assert rel(post.mode(), gm[:, :4]) <= 5e-3 and post.sample().shape == (2, 4, 16, 16)
        assert torch.allclose(post.std, torch.exp(0.5 * torch.clamp(post.parameters[:, 4:], -30, 20)))
    finally:
        torch.set_grad_enabled(True)


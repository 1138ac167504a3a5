"""Host logic of the denoiser (magicdance_b200/engine.py + pipeline.py) on the CPU.

Every `magicdance_b200.ops` entry point the engine calls is replaced by a PyTorch stand-in from tests/fake_ops.py
that reads the SAME packed layouts as the CUDA kernels (NHWC fp16 activations, tap-major conv weights, interleaved
GEGLU rows, transposed V, per-batch bias rows, dual-source A ...).  What is checked here is the orchestration — the
block plan walk, weight packing, bank projection and its two-source attention, pose-residual bookkeeping, the
cond/uncond paired batch, the DDIM update — against the golden vectors of the UNMODIFIED reference.  The kernels
themselves are checked on the GPU (tests/test_kernels_gpu.py, tests/test_parity_gpu.py)."""
import pytest
import torch

from oracle import synth
from tests import fake_ops
from tests import golden_util as G

TOL = 5e-3  # fp16 storage (emulated by the stand-ins) vs the fp32 reference

_PATCHED = ("gemm", "attention", "conv3x3_direct", "groupnorm", "layernorm", "upsample2x", "add", "im2col3x3",
            "timestep_embedding", "skinny_linear", "nchw_f32_to_nhwc_f16", "nhwc_f16_to_nchw_f32", "softmax_rows",
            "ensure_device", "require_cuda")


@pytest.fixture(scope="module")
def engine():
    """One engine (2.1 G parameters repacked to fp16 on the CPU) for the whole module, with the kernel entry points
    patched for exactly as long as it lives."""
    from magicdance_b200 import ops
    from magicdance_b200.engine import DenoiseEngine
    mp = pytest.MonkeyPatch()
    for name in _PATCHED + ("cfg_ddim_update",):
        mp.setattr(ops, name, getattr(fake_ops, name))
    torch.set_grad_enabled(False)
    try:
        yield DenoiseEngine(synth.synth_state_dict(seed=0), device="cpu")
    finally:
        mp.undo()


def test_apply_model_orchestration_matches_reference_small32(engine):
    g = G.load("small32")
    inp = G.small32_inputs()
    eps_c, bank, pose, _ = engine.apply_model(inp["x"], inp["t"], inp["context"], inp["pose"], inp["ref"], uc=False,
                                              return_parts=True)
    assert len(bank) == 16 and len(pose) == 13
    for i, b in enumerate(bank):  # the appearance net's norm1 states: ours [B*N, C], the reference's (B, N, C)
        shape = tuple(int(v) for v in g[f"small32/bank{i}/shape"])
        G.check_summary(g, f"small32/bank{i}", b.reshape(shape), TOL)
    for i, p_ in enumerate(pose):  # 13 ControlNet residuals: ours NHWC [B*H*W, C], the reference's NCHW
        bsz, c, h, w = (int(v) for v in g[f"small32/pose{i}/shape"])
        G.check_summary(g, f"small32/pose{i}", p_.reshape(bsz, h, w, c).permute(0, 3, 1, 2), TOL)
    e_c = G.rel_l2(eps_c, torch.from_numpy(g["small32/eps_c"]))
    eps_u = engine.apply_model(inp["x"], inp["t"], inp["context"], inp["pose"], None, uc=True)
    e_u = G.rel_l2(eps_u, torch.from_numpy(g["small32/eps_u"]))
    assert e_c <= TOL and e_u <= TOL, (e_c, e_u)
    assert G.rel_l2(eps_c, eps_u) > 1e-2  # the two branches really differ


def test_sampler_step_orchestration_matches_reference_full64(engine):
    """One full p_sample_ddim (ddim.py:518-645; index 49, t = 981, CFG 7) at the headline size through
    pipeline.DenoisePipeline.step — bank build, K/V projection, cached hint features, the paired cond/uncond
    UNet batch and the fused update — against the unmodified reference's x_prev / pred_x0 / eps."""
    from magicdance_b200.pipeline import DenoisePipeline
    g = G.load("full64")
    inp = G.full64_inputs()
    pipe = DenoisePipeline(engine, ddim_steps=50, scale=7.0, eta=0.0)
    assert int(pipe.t_dev[49]) == int(inp["t"][0]) == 981
    bank_kv = pipe.reference_bank(inp["ref"], inp["context"], 49)
    hint = pipe.hint(inp["pose"])
    x_prev, pred_x0, eps_c, eps_u = pipe.step(inp["x"], 49, inp["context"], hint, bank_kv)
    errs = {k: G.rel_l2(v, torch.from_numpy(g["full64/" + k]))
            for k, v in (("eps_c", eps_c), ("eps_u", eps_u), ("x_prev", x_prev), ("pred_x0", pred_x0))}
    assert errs["eps_c"] <= TOL and errs["eps_u"] <= TOL, errs
    # CFG 7 amplifies the difference of two nearly equal fp16-rounded fields (ddim.py:605)
    assert errs["x_prev"] <= 2 * TOL and errs["pred_x0"] <= 4 * TOL, errs


def test_timestep_batched_bank_slots_feed_the_step_like_the_direct_bank(engine):
    """The multi-GPU / multi-frame data path: the appearance pass batched over TIMESTEPS
    (pipeline.build_bank_slots) writes each timestep's projected K / V^T into one flat slot (parallel.BankLayout —
    what the NCCL all-gather moves); a step that reads its bank through views of that slot must equal a step that
    builds the bank for its own timestep directly."""
    from magicdance_b200 import parallel
    from magicdance_b200.pipeline import DenoisePipeline, build_bank_slots
    inp = synth.synth_inputs(1, 32, seed=3, shared_reference=True)
    pipe = DenoisePipeline(engine, ddim_steps=50, scale=7.0, eta=0.0)
    geo = engine.attn_geometry(32, 32)
    tokens = [n for n, _ in geo]
    layout = parallel.BankLayout([(n, c) for n, c in geo])
    indices = [49, 20, 3]                                   # three timesteps in ONE appearance pass
    slots = torch.zeros((len(indices), layout.numel), dtype=torch.float16)
    build_bank_slots(engine, inp["ref"], pipe.t_dev[torch.as_tensor(indices)], inp["context"], layout, tokens, slots)
    hint = pipe.hint(inp["pose"])
    for j, ix in enumerate(indices[:2]):
        via_slot = layout.views(slots[j], tokens, 1)
        direct = pipe.reference_bank(inp["ref"], inp["context"], ix)
        assert len(via_slot) == len(direct) == 16
        for (k_s, vt_s, n_s, b_s), (k_d, vt_d, n_d, b_d) in zip(via_slot, direct):
            assert (n_s, b_s) == (n_d, b_d)
            # batched over 3 timesteps vs alone: the same arithmetic per sample up to fp16 rounding noise
            assert G.rel_l2(k_s, k_d) <= 3e-3 and G.rel_l2(vt_s, vt_d[:, :n_d]) <= 3e-3
        x1, _, ec1, _ = pipe.step(inp["x"], ix, inp["context"], hint, via_slot)
        x2, _, ec2, _ = pipe.step(inp["x"], ix, inp["context"], hint, direct)
        assert G.rel_l2(ec1, ec2) <= 3e-3 and G.rel_l2(x1, x2) <= 3e-3
    # the bank really depends on the timestep (SURVEY §8a semantics 4)
    assert G.rel_l2(slots[0].float(), slots[1].float()) > 1e-2


def test_shared_reference_bank_and_paired_batch_match_the_oracle(engine):
    """The bench / video configuration: several frames share ONE reference image, so the bank is built for a
    single sample (kv1_batches = 1) and read by the conditional half of the paired cond/uncond batch only
    (bank_batches = frames).  Checked against the CPU oracle (itself pinned to the reference) on fresh inputs."""
    from oracle import restatement as R
    inp = synth.synth_inputs(2, 32, seed=5, shared_reference=True)
    inp["t"] = torch.full((2,), 441, dtype=torch.long)
    weights = synth.synth_state_dict(seed=0)
    ref_c = R.apply_model(weights, inp["x"], inp["t"], inp["context"], inp["pose"], inp["ref"], uc=False)
    ref_u = R.apply_model(weights, inp["x"], inp["t"], inp["context"], inp["pose"], None, uc=True)
    del weights
    bank = engine.appearance_write(inp["ref"][:1], inp["t"][:1], inp["context"][:1])
    bank_kv = engine.project_bank(bank, 1)
    pose = engine.controlnet(inp["x"], engine.hint_features(inp["pose"]), inp["t"], inp["context"])
    eps_c, eps_u = engine.unet_forward(inp["x"], inp["t"], inp["context"], bank_kv=bank_kv, pose=pose, cfg_pair=True)
    e_c, e_u = G.rel_l2(eps_c, ref_c), G.rel_l2(eps_u, ref_u)
    assert e_c <= TOL and e_u <= TOL, (e_c, e_u)

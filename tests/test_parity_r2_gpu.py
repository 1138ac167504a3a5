"""Round-2 parity gates on the EXACT code bench.py times (pipeline.GraphedDenoiser) and on the reference-facing
sample_log / p_losses API, against goldens made by the unmodified reference (oracle/make_golden_r2.py):

  tests/golden/b8_64.npz    BASELINE configs[2]: eight frames (one x_T / reference / prompt, eight pose maps), one
                            full p_sample_ddim at latent 64 — the paired cond/uncond batch of 16
  tests/golden/traj50.npz   BASELINE configs[1]: the whole 50-step chain, x after ddim index 49/40/25/10/0
  tests/golden/ploss32.npz  LatentDiffusionReferenceOnly.p_losses forward (ddpm.py:2165-2212)

Tolerances (fp16 storage / fp32 accumulation against the fp32 reference; SYNTHETIC weights — no checkpoint ships):
CFG-combined step outputs <= 5e-3 (measured 0.9e-3 x_prev / 2.4e-3 pred_x0: the x7 guidance amplifies the cond/uncond
difference), 50-step trajectory <= 1e-2 (measured 1.7e-3 at every checkpoint: the chain does not amplify the error),
eps <= 5e-3.  SURVEY §8c proposed 2e-2 / 3e-2; the measured margins allow the tighter gates.
"""
import os


import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML = os.path.join(REPO, "model_lib", "ControlNet", "models", "cldm_v15_reference_only_pose.yaml")

TOL_EPS = 5e-3
TOL_STEP = 5e-3
TOL_TRAJ = 1e-2
TRAJ_KEEP = (49, 40, 25, 10, 0)


@pytest.fixture(scope="module")
def engine():
    from magicdance_b200 import synth
    from magicdance_b200.engine import DenoiseEngine
    torch.set_grad_enabled(False)
    return DenoiseEngine(synth.synth_state_dict(seed=0), device="cuda")


def b8_inputs():
    from magicdance_b200 import synth
    inp = synth.synth_inputs(8, 64, seed=0, shared_reference=True)
    inp["x"] = inp["x"][:1].expand(8, -1, -1, -1).contiguous()
    return inp


def test_b8_latent64_step_through_the_step_graph(engine):
    """configs[2]'s shape through the captured graphs: bank of ddim index 49 built by the bank graph, one replay of
    the step graph with the paired batch of 16, against the reference sampler's own x_prev / pred_x0."""
    from tests import golden_util as G
    from magicdance_b200.pipeline import DenoisePipeline, GraphedDenoiser
    g = G.load("b8_64")
    inp = {k: v.cuda() for k, v in b8_inputs().items()}
    pipe = DenoisePipeline(engine, ddim_steps=50, scale=7.0, eta=0.0)
    gd = GraphedDenoiser(pipe, 8, (64, 64), inp["context"][:1].contiguous(), bank_chunk=2).capture()
    slots = torch.zeros((1, gd.layout.numel), dtype=torch.float16, device="cuda")
    gd.build_bank([49], inp["ref"][:1], slots)
    gd.hint.copy_(pipe.hint(inp["pose"]))
    gd.x.copy_(inp["x"])
    gd.step(49, slots[0])
    torch.cuda.synchronize()
    e_x = G.rel_l2(gd.x_prev, torch.from_numpy(g["x_prev"]))
    e_p = G.rel_l2(gd.pred_x0, torch.from_numpy(g["pred_x0"]))
    # per frame too: a mis-routed pose map or bank would show up as ONE bad frame inside a good average
    per = [G.rel_l2(gd.pred_x0[i], torch.from_numpy(g["pred_x0"][i])) for i in range(8)]
    print(f"b8_64: x_prev {e_x:.3e} pred_x0 {e_p:.3e} per-frame max {max(per):.3e}")
    assert e_x <= TOL_STEP and e_p <= TOL_STEP and max(per) <= TOL_STEP
    # eager pipeline on the same inputs: eps of the conditional / unconditional halves agree with the graph's result
    bank_kv = pipe.reference_bank(inp["ref"], inp["context"], 49, first_only=True)
    x_e, p_e, _, _ = pipe.step(inp["x"], 49, inp["context"][:1].contiguous(), pipe.hint(inp["pose"]), bank_kv)
    assert G.rel_l2(x_e, gd.x_prev) <= 2e-3


def test_traj50_chain_through_the_graphs(engine):
    """configs[1]: the 50-step chain exactly as bench.py runs it (bank built 25 timesteps per appearance pass, one
    step-graph replay per step) against the reference sampler's ddim_sampling trajectory."""
    from tests import golden_util as G
    from magicdance_b200 import parallel
    from magicdance_b200.pipeline import DenoisePipeline, GraphedDenoiser, plan_bank_chunks
    g = G.load("traj50")
    inp = {k: v.cuda() for k, v in G.full64_inputs().items()}
    pipe = DenoisePipeline(engine, ddim_steps=50, scale=7.0, eta=0.0)
    gd = GraphedDenoiser(pipe, 1, (64, 64), inp["context"], bank_chunk=parallel.bank_chunk_size(50, 1)).capture()
    slots = torch.zeros((50, gd.layout.numel), dtype=torch.float16, device="cuda")
    order = list(range(49, -1, -1))
    for s0, part in plan_bank_chunks(order, gd.bank_chunk):
        gd.build_bank(part, inp["ref"], slots[s0:s0 + len(part)])
    gd.hint.copy_(pipe.hint(inp["pose"]))
    gd.x.copy_(inp["x"])
    errs = {}
    for s, ix in enumerate(order):
        gd.step(ix, slots[s])
        if ix in TRAJ_KEEP:
            errs[ix] = (G.rel_l2(gd.x_prev, torch.from_numpy(g[f"x_prev/{ix}"])),
                        G.rel_l2(gd.pred_x0, torch.from_numpy(g[f"pred_x0/{ix}"])))
    print("traj50 rel-L2 (x_prev, pred_x0) by ddim index:", {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in errs.items()})
    assert torch.isfinite(gd.x_prev).all()
    assert errs[49][0] <= TOL_STEP
    for ix in TRAJ_KEEP:
        assert errs[ix][0] <= TOL_TRAJ, (ix, errs[ix])
    assert G.rel_l2(gd.x_prev, torch.from_numpy(g["x_final"])) <= TOL_TRAJ


@pytest.fixture(scope="module")
def model():
    from magicdance_b200 import synth
    from model_lib.ControlNet.cldm.model import create_model
    torch.set_grad_enabled(False)
    m = create_model(YAML)
    missing, unexpected = m.load_state_dict(synth.synth_state_dict(seed=0), strict=False)
    assert not unexpected and all(k in synth.SCHEDULE_KEYS or k.startswith("first_stage_model.") for k in missing)
    return m.cuda().eval()


def test_sample_log_full_chain_matches_reference_and_is_bit_reproducible(model):
    """The call test_tiktok.py:261-268 makes — model.sample_log(cond, ..., x_T) — for the 50-step chain of configs[1]
    (CUDA-graph replay is the default path of the drop-in sampler), from HOST tensors; a second frame of the same
    reference reuses the bank and must reproduce the first BIT FOR BIT (no atomics anywhere on the path)."""
    from tests import golden_util as G
    from magicdance_b200 import ops
    g = G.load("traj50")
    inp = G.full64_inputs()  # host tensors: the H2D copies happen inside the call
    gen = torch.Generator().manual_seed(123)
    uc_ctx = torch.randn(1, 77, 768, generator=gen)  # ignored by the reference (ddim.py:599-604)
    c = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "image_control": [inp["ref"]], "wonoise": True,
         "overlap_sampling": False}
    uc = {"c_concat": [inp["pose"]], "c_crossattn": [uc_ctx], "wonoise": True, "overlap_sampling": False}
    seen = []
    n0 = ops.launch_count()
    s1, inter = model.sample_log(cond=c, batch_size=1, ddim=True, ddim_steps=50, eta=0.0, unconditional_guidance_scale=7,
                                 unconditional_conditioning=uc, inpaint=None, x_T=inp["x"],
                                 img_callback=lambda p0, i: seen.append(i))
    n1 = ops.launch_count()
    s2, _ = model.sample_log(cond=c, batch_size=1, ddim=True, ddim_steps=50, eta=0.0, unconditional_guidance_scale=7,
                             unconditional_conditioning=uc, inpaint=None, x_T=inp["x"])
    torch.cuda.synchronize()
    assert seen == list(range(50)) and s1.shape == (1, 4, 64, 64) and s1.is_cuda
    assert len(inter["x_inter"]) >= 2
    err = G.rel_l2(s1, torch.from_numpy(g["x_final"]))
    print(f"sample_log 50-step x_0 rel-L2 vs reference {err:.3e}")
    assert err <= TOL_TRAJ
    assert torch.equal(s1, s2), "two runs of the same frame differ: something on the path is not deterministic"
    assert n1 - n0 < 5000  # graphs: the eager loop would be ~650 launches x 50 steps


def test_sample_log_eager_loop_equals_graph_replay(model):
    """the eager per-step loop (use_graphs=False) and the graph replay give the same chain (4 steps, latent 32)"""
    from tests import golden_util as G
    from magicdance_b200 import synth
    from model_lib.ControlNet.ldm.models.diffusion.ddim import DDIMSampler_ReferenceOnly
    inp = {k: v.cuda() for k, v in synth.synth_inputs(2, 32, seed=5, shared_reference=True).items()}
    inp["x"] = inp["x"][:1].expand(2, -1, -1, -1).contiguous()
    c = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "image_control": [inp["ref"]], "wonoise": True,
         "overlap_sampling": False}
    uc = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "wonoise": True, "overlap_sampling": False}
    model.image_size = 32
    try:
        kw = dict(cond=c, batch_size=2, ddim=True, ddim_steps=4, eta=0.0, unconditional_guidance_scale=7.0,
                  unconditional_conditioning=uc, x_T=inp["x"])
        a, _ = model.sample_log(**kw)
        DDIMSampler_ReferenceOnly.use_graphs = False
        b, _ = model.sample_log(**kw)
    finally:
        DDIMSampler_ReferenceOnly.use_graphs = True
        model.image_size = 64
    assert torch.isfinite(a).all()
    assert G.rel_l2(a, b) <= 5e-3  # same kernels; split-K choices differ between the batch-1 bank pass and the batched one


def test_p_losses_forward_matches_reference(model):
    """ddpm.py:2165-2212 forward value: q_sample, apply_model on the noised latent with per-sample t, the eps-loss."""
    from tests import golden_util as G
    g = G.load("ploss32")
    inp = {k: v.cuda() for k, v in G.small32_inputs().items()}
    x0, noise = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["noise"]).cuda()
    t = torch.from_numpy(g["t"]).cuda()
    cond = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "image_control": [inp["ref"]], "wonoise": True}
    got = {}
    fwd = model.apply_model

    def rec(*a, **k):
        got["eps"] = fwd(*a, **k)
        return got["eps"]

    model.apply_model = rec
    try:
        with torch.no_grad():
            loss, ld = model.p_losses(x0, cond, t, noise=noise)
    finally:
        del model.apply_model
    assert G.rel_l2(got["eps"], torch.from_numpy(g["eps"])) <= TOL_EPS
    assert abs(float(loss) - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))
    assert abs(float(ld["val/loss_simple"]) - float(g["loss_simple"])) <= 1e-2 * abs(float(g["loss_simple"]))
    assert abs(float(ld["val/loss_vlb"]) - float(g["loss_vlb"])) <= 1e-2 * abs(float(g["loss_vlb"])) + 1e-9

"""Oracle for the TRAINING row: the restatement's p_losses + autograd against the
gradients the UNMODIFIED reference produced with its CheckpointFunction active and the stage-2 freeze policy
(tests/golden/grad16.npz, oracle/make_golden_grad.py).  The product has no backward kernels; this pins the
checker they will be held to, exactly as test_oracle.py pins the forward."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import synth
from tests import golden_util as G

TRAINED = ("appearance_control_model.", "pose_control_model.")


@pytest.fixture(scope="module")
def grads():
    gold = G.load("grad16")
    names = [str(n) for n in gold["names"]]
    sd = synth.synth_state_dict(seed=0)
    # train_tiktok.py:798-822 (--finetune_control): appearance net + pose ControlNet trained, SD UNet frozen
    assert sorted(names) == sorted(k for k in sd if k.startswith(TRAINED))
    for k in names:
        sd[k] = sd[k].clone().requires_grad_(True)
    inp = G.grad16_inputs()
    x_noisy = R.q_sample(inp["x0"], inp["t_train"], inp["noise"], R.make_schedule()["alphas_cumprod"]).requires_grad_(True)
    with torch.enable_grad():
        loss, loss_simple, _ = R.p_losses(sd, inp["x0"], inp["t_train"], inp["noise"], inp["context"], inp["pose"],
                                          inp["ref"], x_noisy=x_noisy)
        loss.backward()
    return gold, names, sd, float(loss.detach()), x_noisy.grad


def test_loss_and_input_gradient_match_the_reference(grads):
    gold, names, sd, loss, dx = grads
    assert abs(loss - float(gold["loss"])) <= 1e-5 * float(gold["loss"])
    # through the frozen UNet (dgrad only) and both trained branches
    assert G.rel_l2(dx, torch.from_numpy(gold["d_x_noisy"])) <= 2e-4


def test_the_same_parameters_are_left_without_a_gradient(grads):
    """the tail of the appearance net after its last norm1 and its `out` never reach the loss: the reference wraps the
    model in DDP(find_unused_parameters=True) for these (train_tiktok.py:1002-1009)"""
    gold, names, sd, _, _ = grads
    want = {n for n, h in zip(names, gold["has_grad"]) if not h}
    got = {n for n in names if sd[n].grad is None or float(sd[n].grad.abs().max()) == 0.0}
    assert want <= got  # the restatement skips the dead tail; nothing the reference trains may be missing
    reached = {n for n in names if n not in got}
    assert reached == {n for n, h in zip(names, gold["has_grad"]) if h and float(gold["gnorm"][names.index(n)]) > 0}
    assert len(want) == 36


def test_every_parameter_gradient_matches_the_reference(grads):
    gold, names, sd, _, _ = grads
    worst = 0.0
    for i, n in enumerate(names):
        if not gold["has_grad"][i] or float(gold["gnorm"][i]) == 0.0:
            continue
        g = sd[n].grad.detach().double().flatten()
        norm = float(gold["gnorm"][i])
        assert abs(float(g.norm()) - norm) <= 5e-4 * norm, n
        pos = G.grad_sample_positions(g.numel())
        err = float((g[torch.from_numpy(pos)] - torch.from_numpy(gold["gsample"][i, :len(pos)])).norm()) / (norm / np.sqrt(g.numel()) * np.sqrt(len(pos)))
        worst = max(worst, err)
        assert err <= 5e-3, (n, err)
        assert abs(float(g.sum()) - float(gold["gsum"][i])) <= 5e-4 * norm * np.sqrt(g.numel()) + 1e-9, n
    for key in gold.files:
        if key.startswith("full/"):
            assert G.rel_l2(sd[key[5:]].grad, torch.from_numpy(gold[key])) <= 2e-4, key
    assert float(gold["ckpt_vs_plain_max_rel"]) <= 1e-4  # CheckpointFunction (util.py:118-187) changes memory, not values

"""GPU gate of the opt-in CUDA-graph path of the drop-in sampler (MDB_DROPIN_GRAPH=1, magicdance_b200/dropin/ddim.py
_ddim_sampling_graphed): sample_log through the reference-facing API, eager loop vs graph replay — same result
(to the run-to-run tolerance of the eager path) and the time per step of both, for a second frame of the same
reference (bank and graphs cached) as in a video.

    python scripts/gpu_dropin_graph_check.py [--latent 64] [--steps 50]
"""
import argparse
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    from magicdance_b200 import synth
    from model_lib.ControlNet.cldm.model import create_model
    torch.set_grad_enabled(False)
    m = create_model(os.path.join(REPO, "model_lib", "ControlNet", "models", "cldm_v15_reference_only_pose.yaml"))
    missing, unexpected = m.load_state_dict(synth.synth_state_dict(seed=0), strict=False)
    assert not unexpected
    m = m.cuda().eval()
    m.image_size = args.latent
    inp = {k: v.cuda() for k, v in synth.synth_inputs(1, args.latent, seed=5, shared_reference=True).items()}
    c = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "image_control": [inp["ref"]], "wonoise": True,
         "overlap_sampling": False}
    uc = {"c_concat": [inp["pose"]], "c_crossattn": [inp["context"]], "wonoise": True, "overlap_sampling": False}

    def run(tag, reps=2):
        out, best = None, 1e30
        for _ in range(reps):  # the first call builds bank / graphs, the second is a later frame of the video
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out, _ = m.sample_log(cond=c, batch_size=1, ddim=True, ddim_steps=args.steps, eta=0.0,
                                  unconditional_guidance_scale=7.0, unconditional_conditioning=uc, x_T=inp["x"])
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"{tag}: {best * 1e3 / args.steps:.3f} ms/step ({args.steps / best:.1f} steps/s), finite={bool(torch.isfinite(out).all())}",
              flush=True)
        return out

    os.environ["MDB_DROPIN_GRAPH"] = "0"
    eager = run("eager loop ")
    eager2 = run("eager again", reps=1)
    os.environ["MDB_DROPIN_GRAPH"] = "1"
    graph = run("graph replay")
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    noise, err = rel(eager2, eager), rel(graph, eager)
    print(f"eager vs eager (atomics order) {noise:.3e}; graph vs eager {err:.3e}")
    ok = err <= max(5e-3, 3 * noise)
    print("OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

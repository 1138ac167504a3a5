"""Runs the per-kernel numerics cases (tests/kernel_cases.py) one by one and prints every error
instead of stopping at the first failure.  Used through gpurun while bringing kernels up:

    python scripts/gpu_diag.py --group gemm
"""
import argparse
import os
import sys
import time
import traceback

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from tests import kernel_cases as K  # noqa: E402

GROUPS = {
    "misc": ("case_layout", "case_add", "case_upsample", "case_time_path", "case_cfg_ddim", "case_layernorm",
             "case_groupnorm", "case_conv_direct"),
    "gemm": ("case_gemm", "case_gemm_ln", "case_gemm_batch_bias", "case_gemm_dual", "case_gemm_strided_out", "case_geglu"),
    "conv": ("case_conv", "case_conv_s2", "case_down", "case_conv_im2col"),
    "attn": ("case_attention",),
    "tuned": ("case_tuned",),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="all")
    ap.add_argument("--first", type=int, default=0, help="only the first N cases of the group")
    args = ap.parse_args()
    names = sum(GROUPS.values(), ()) if args.group == "all" else GROUPS[args.group]
    cases = [(f, a) for f, a in K.ALL_CASES if f.__name__ in names]
    if args.first:
        cases = cases[:args.first]
    print(f"device: {torch.cuda.get_device_name(0)}; {len(cases)} cases in group {args.group}", flush=True)
    bad = 0
    for f, a in cases:
        t0 = time.time()
        try:
            err, tol, desc = f(*a)
            torch.cuda.synchronize()
            ok = err <= tol
            bad += not ok
            print(f"{'ok  ' if ok else 'FAIL'} {desc}: err={err:.3e} tol={tol:.1e} ({time.time() - t0:.2f}s)", flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print(f"EXC  {f.__name__}{a}: {type(e).__name__}: {e}", flush=True)
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e2:  # noqa: BLE001
                print("context is dead:", e2, flush=True)
                break
    print(f"group {args.group}: {bad} failing of {len(cases)}", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

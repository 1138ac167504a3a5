"""CPU-only checks of the host logic and of the C-ABI library's exported surface."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from magicdance_b200 import build, _lib
    path = build.build()
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "magicdance_b200.h")).read()
    declared = set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    nm = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (mdb_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported
    assert lib.mdb_abi_version() == 1 and lib.mdb_launch_count() == 0


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    from magicdance_b200 import build
    path = build.build()
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "UTMASTG.2D"):
        assert mnem in sass, f"{mnem} (tcgen05 / TMA) missing from the compiled kernels"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path must not be present"


def test_no_cuda_means_loud_failure():
    from magicdance_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.ensure_device()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.layernorm(torch.zeros(2, 320).half(), torch.ones(320), torch.zeros(320))


def test_block_plan_and_packing_consume_the_reference_state_dict():
    """Every key of the reference state dict (manifest recorded from the unmodified reference) is
    consumed by the weight packer, except the appearance twin's dead hint block; shapes line up."""
    from magicdance_b200 import synth
    from magicdance_b200.engine import NetConfig, PackedNet, UNET, APPEARANCE, POSE
    manifest = synth.load_manifest()
    seen = set()

    class Rec(dict):
        def __getitem__(self, k):
            seen.add(k)
            return torch.empty(manifest[k], device="meta")

        def __contains__(self, k):
            return k in manifest

    sd = Rec()
    for prefix, kind in ((UNET, "unet"), (APPEARANCE, "appearance"), (POSE, "controlnet")):
        PackedNet(sd, prefix, NetConfig(), kind, "meta")
    nets = [k for k in manifest if k.startswith((UNET, APPEARANCE, POSE))]
    unused = sorted(set(nets) - seen)
    assert all(k.startswith(APPEARANCE + "input_hint_block.") for k in unused), unused[:5]
    assert len(unused) == 16


def test_geglu_packing_interleaves_value_and_gate_rows():
    from magicdance_b200.engine import pack_geglu
    c = 64
    w = torch.arange(8 * c, dtype=torch.float32)[:, None].expand(8 * c, 4).contiguous()
    b = torch.arange(8 * c, dtype=torch.float32)
    wp, bp = pack_geglu(w, b, "cpu")
    assert bp[:32].tolist() == list(range(32)) and bp[32:64].tolist() == list(range(4 * c, 4 * c + 32))
    assert bp[64:96].tolist() == list(range(32, 64))
    assert torch.equal(wp[:, 0].float(), bp.half().float())


def test_schedule_matches_reference_buffers():
    from magicdance_b200 import pipeline as P
    from tests import golden_util as G
    g = G.load("full64")
    acp = P.alphas_cumprod_f32()
    np.testing.assert_allclose(acp, g["full64/alphas_cumprod"], rtol=2e-6)
    ts = P.ddim_timesteps_uniform(50)
    assert list(ts) == list(g["full64/ddim_timesteps"]) and ts[0] == 1 and ts[-1] == 981
    sig, a, ap = P.ddim_parameters(acp, ts, 0.0)
    np.testing.assert_allclose(a, g["full64/ddim_alphas"], rtol=1e-6)
    np.testing.assert_allclose(ap, g["full64/ddim_alphas_prev"], rtol=1e-6)
    assert float(np.abs(sig).max()) == 0.0


def test_sharding_helpers():
    from magicdance_b200 import parallel as P
    assert [len(P.shard_frames(64, 8, r)) for r in range(8)] == [8] * 8
    assert [len(P.shard_frames(10, 4, r)) for r in range(4)] == [3, 3, 2, 2]
    covered = sorted(i for r in range(4) for i in P.shard_frames(10, 4, r))
    assert covered == list(range(10))
    idx = list(range(49, -1, -1))
    shares = [P.shard_timesteps(idx, 8, r) for r in range(8)]
    assert [len(s) for s in shares] == [7, 7, 6, 6, 6, 6, 6, 6]
    assert sorted(sum(shares, [])) == list(range(50))
    table = P.owner_slot(idx, 8)
    for r, sh in enumerate(shares):
        for s, ix in enumerate(sh):
            assert table[ix] == (r, s)
    # bank-build batches: equal chunks of at most 25 of the rank's share, never more launches than needed
    assert [P.bank_chunk_size(50, w) for w in (1, 2, 4, 8)] == [25, 25, 13, 7]
    assert P.bank_chunk_size(3, 1) == 3 and P.bank_chunk_size(1, 8) == 1 and P.bank_chunk_size(51, 1) == 17
    for n in range(1, 60):
        for w in (1, 2, 3, 8):
            c, share = P.bank_chunk_size(n, w), (n + w - 1) // w
            assert 1 <= c <= 25 and -(-share // c) == -(-share // 25)
    # overlapped bank build: chunks follow the order in which the steps consume the timesteps, slots are contiguous
    from magicdance_b200.pipeline import plan_bank_chunks
    plan = plan_bank_chunks(idx, 10)
    assert [s0 for s0, _ in plan] == [0, 10, 20, 30, 40] and plan[0][1] == list(range(49, 39, -1))
    assert sum((part for _, part in plan), []) == idx
    assert plan_bank_chunks([49, 48, 49, 47], 2) == [(0, [49, 48]), (2, [47])]  # repeated steps share one slot


def test_header_is_plain_c_and_matches_the_ctypes_structs(tmp_path):
    """include/magicdance_b200.h must compile as C99 (it is what a cgo/JNI/ctypes host binds) and the
    descriptor structs must have exactly the layout magicdance_b200/_lib.py declares."""
    import ctypes as C
    import os
    from magicdance_b200 import _lib
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "magicdance_b200.h"', 'int main(void) {']
    for cname, cls in (("mdb_gemm_desc", _lib.GemmDesc), ("mdb_attn_desc", _lib.AttnDesc)):
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in (("mdb_gemm_desc", _lib.GemmDesc), ("mdb_attn_desc", _lib.AttnDesc)):
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_vae_decoder_packing_consumes_every_decoder_tensor():
    """host logic of the (GPU-unvalidated, opt-in) VAE decoder: the repack reads every
    first_stage_model.{post_quant_conv,decoder}.* tensor of the reference state dict exactly once, and the folds
    (1/scale_factor into post_quant_conv, c^-0.5 into q, the v bias into proj_out) are the ones the oracle implies."""
    import json
    import os
    import torch
    from magicdance_b200 import synth
    from magicdance_b200.vae import PackedVaeDecoder, PREFIX, SCALE_FACTOR
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "magicdance_b200", "vae_manifest.json")) as f:
        manifest = json.load(f)
    sd = synth.synth_state_dict(manifest, seed=0)
    p = PackedVaeDecoder(sd, "cpu")
    want = {k for k in manifest if k.startswith(PREFIX + "decoder.") or k.startswith(PREFIX + "post_quant_conv.")}
    assert sorted(p.consumed) == sorted(want) and len(set(p.consumed)) == len(p.consumed)
    assert tuple(p.in_w.shape) == (512, 36) and tuple(p.out_w.shape) == (3, 9 * 128) and p.c_mid == 512
    assert [r.cout for lvl in (3, 2, 1, 0) for r in p.up[lvl][0]] == [512] * 6 + [256] * 3 + [128] * 3
    assert p.up[0][1] is None and all(p.up[lvl][1] is not None for lvl in (1, 2, 3))
    # centre-tap 1x1 with the 1/scale fold
    w = sd[PREFIX + "post_quant_conv.weight"][:, :, 0, 0] / SCALE_FACTOR
    got = p.pq_w.float().reshape(4, 3, 3, 4)
    assert torch.allclose(got[:, 1, 1, :], w.half().float()) and float(got.abs().sum() - got[:, 1, 1, :].abs().sum()) == 0.0
    # proj_out(P (V + 1 bv^T)) == proj_out(P V) + Wp bv + bp
    c = 512
    wp = sd[PREFIX + "decoder.mid.attn_1.proj_out.weight"].reshape(c, c)
    bias = sd[PREFIX + "decoder.mid.attn_1.proj_out.bias"] + wp @ sd[PREFIX + "decoder.mid.attn_1.v.bias"]
    assert torch.allclose(p.bp, bias, atol=1e-6)
    assert torch.allclose(p.wq.float(), (sd[PREFIX + "decoder.mid.attn_1.q.weight"].reshape(c, c) * c ** -0.5).half().float())


def test_vae_decoder_orchestration_matches_the_oracle_with_cpu_test_doubles(monkeypatch):
    """The VAE decoder's host logic (magicdance_b200/vae.py: operand order, layouts, the three folds, the
    GEMM -> softmax -> GEMM attention) run on tests/fake_ops.py — PyTorch stand-ins that read the same packed
    layouts as the kernels — must reproduce the pinned oracle / the reference golden at latent 16.  The CUDA
    kernels are not exercised here (tests/test_kernels_gpu.py, scripts/gpu_vae_parity.py)."""
    import json
    import os
    import numpy as np
    import torch
    from magicdance_b200 import ops, synth, vae
    from oracle import vae_restatement as V
    from tests import fake_ops
    for name in ("gemm", "conv3x3_direct", "groupnorm", "upsample2x", "softmax_rows", "nchw_f32_to_nhwc_f16",
                 "nhwc_f16_to_nchw_f32", "im2col3x3"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "magicdance_b200", "vae_manifest.json")) as f:
        manifest = json.load(f)
    torch.set_grad_enabled(False)
    sd = synth.synth_state_dict(manifest, seed=0)
    dec = vae.VaeDecoder.__new__(vae.VaeDecoder)  # no device check: the test doubles run on the CPU
    dec.p = vae.PackedVaeDecoder(sd, "cpu")
    z, _, _ = V.vae_inputs(2, 16)
    img = dec._decode(z)
    gold = torch.from_numpy(np.load(os.path.join(here, "tests", "golden", "vae16.npz"))["vae16/decoded"])
    err = float((img.double() - gold.double()).norm() / gold.double().norm())
    assert tuple(img.shape) == (2, 3, 128, 128) and err <= 5e-3, err
    with __import__("pytest").raises(RuntimeError, match="no CPU fallback"):
        dec.decode(z)  # the public entry refuses CPU tensors


def test_vae_encoder_orchestration_matches_the_reference_golden_with_cpu_test_doubles(monkeypatch):
    """The VAE encoder's host logic (magicdance_b200/vae.py: the bottom/right-padded stride-2 downsample as
    im2col(pad="br") + GEMM, ResnetBlocks, the single-head attention, quant_conv as a centre-tap conv) on the CPU test
    doubles must reproduce the moments and the scaled posterior sample the UNMODIFIED reference produced
    (tests/golden/vae16.npz), and its repack must read every encoder / quant_conv tensor exactly once."""
    import json
    import os
    import numpy as np
    import torch
    from magicdance_b200 import ops, synth, vae
    from oracle import vae_restatement as V
    from tests import fake_ops
    for name in ("gemm", "conv3x3_direct", "groupnorm", "softmax_rows", "nchw_f32_to_nhwc_f16", "nhwc_f16_to_nchw_f32",
                 "im2col3x3"):
        monkeypatch.setattr(ops, name, getattr(fake_ops, name))
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "magicdance_b200", "vae_manifest.json")) as f:
        manifest = json.load(f)
    torch.set_grad_enabled(False)
    sd = synth.synth_state_dict(manifest, seed=0)
    enc = vae.VaeEncoder.__new__(vae.VaeEncoder)  # no device check: the test doubles run on the CPU
    enc.p = vae.PackedVaeEncoder(sd, "cpu")
    want = {k for k in manifest if k.startswith(vae.PREFIX + "encoder.") or k.startswith(vae.PREFIX + "quant_conv.")}
    assert sorted(enc.p.consumed) == sorted(want) and len(set(enc.p.consumed)) == len(enc.p.consumed)
    # decoder + encoder + post_quant_conv + quant_conv = the whole first_stage_model
    dec = vae.PackedVaeDecoder(sd, "cpu")
    assert set(enc.p.consumed) | set(dec.consumed) == {k for k in manifest if k.startswith(vae.PREFIX)}
    _, img, noise = V.vae_inputs(2, 16)
    mom = enc._encode(img)
    gold = np.load(os.path.join(here, "tests", "golden", "vae16.npz"))
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    assert tuple(mom.shape) == (2, 8, 16, 16) and rel(mom, torch.from_numpy(gold["vae16/moments"])) <= 5e-3
    z = vae.SCALE_FACTOR * vae.posterior_sample(mom, noise)      # get_first_stage_encoding, ddpm.py:1936-1942
    assert rel(z, torch.from_numpy(gold["vae16/encoding"])) <= 5e-3
    assert torch.equal(vae.posterior_sample(mom), mom[:, :4])    # .mode()
    # the double of the br-padded gather is the reference's own F.pad(x, (0,1,0,1)) + unfold(padding=0)
    x = torch.randn(2 * 6 * 6, 8).half()
    col = fake_ops.im2col3x3(x, batch=2, h=6, w=6, c=8, stride=2, pad="br")
    assert tuple(col.shape) == (2 * 3 * 3, 72)
    assert torch.equal(col[0, :8], x[0]) and torch.equal(col[2, 2 * 8:3 * 8], torch.zeros(8).half())  # right edge pad
    with __import__("pytest").raises(RuntimeError, match="no CPU fallback"):
        enc.encode(img)


def test_gemm_plan_plumbing_and_tuner_candidates(tmp_path, monkeypatch):
    """The per-shape launch plan (scripts/gpu_tune_gemm.py -> magicdance_b200/gemm_plan.json -> ops.gemm): no plan file
    is committed, so the library starts with an empty plan; a plan file is keyed by shape, can be disabled with
    MDB_GEMM_PLAN=0; the switches it carries are set only around one launch; the tuner's candidate list starts with the
    engine's own choice and never proposes more splits than K chunks allow."""
    import importlib.util
    import json
    import os
    from magicdance_b200 import ops
    assert not os.path.exists(ops._PLAN_PATH) and ops.load_gemm_plan() == 0 and ops.GEMM_PLAN == {}
    key = ops.gemm_plan_key(512, 1280, 11520, (2, 16, 16, 1280), 0, 0)
    assert key == "512x1280x11520|conv1|epi0|a2_0" and key != ops.gemm_plan_key(512, 1280, 11520, None, 0, 0)
    path = tmp_path / "plan.json"
    path.write_text(json.dumps({"plan": {key: {"splits": 4, "env": {"MDB_GEMM_BN": 160}, "us": 1.0}}}))
    try:
        assert ops.load_gemm_plan(str(path)) == 1 and ops.GEMM_PLAN[key] == {"splits": 4, "env": {"MDB_GEMM_BN": "160"}}
        monkeypatch.setenv("MDB_GEMM_PLAN", "0")
        assert ops.load_gemm_plan(str(path)) == 0
    finally:
        monkeypatch.delenv("MDB_GEMM_PLAN", raising=False)
        ops.load_gemm_plan()
    monkeypatch.delenv("MDB_X_TEST", raising=False)
    with ops._env_switches({"MDB_X_TEST": "1"}):
        assert os.environ["MDB_X_TEST"] == "1"
    assert "MDB_X_TEST" not in os.environ
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gpu_tune_gemm", os.path.join(here, "scripts", "gpu_tune_gemm.py"))
    tune = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tune)
    cands = tune.candidates(8192, 320, 320, None, 0, 1, ["tmast", "pairs"])
    assert cands[0] == ("base", 1, {}) and all(s == 1 for _, s, _ in cands)          # 5 K chunks: no split is proposed
    assert {"tmast", "pairs"} <= {label for label, _, _ in cands}
    cands = tune.candidates(128, 1280, 11520, (2, 8, 8, 1280), 0, 8, ["pairs"])
    assert cands[0] == ("base", 8, {}) and "pairs" not in {label for label, _, _ in cands}   # one M tile: no pair
    assert {s for _, s, _ in cands} == {1, 2, 4, 8}
    assert all(env == {} or set(env) <= {"MDB_GEMM_BN", "MDB_GEMM_DEEP"} for _, _, env in cands)
    assert [c for c in tune.candidates(4096, 2560, 320, None, ops.EPI_GEGLU, 1, [])] == [("base", 1, {})]


def test_gpu_case_lists_and_scripts_are_well_formed():
    """The GPU-side case lists are data that only runs on the GPU box: check here that every (function, args) pair of
    ALL_CASES and PENDING_CASES binds to its function's signature (wrappers are followed to the wrapped case), that the
    pending switches are ones the library reads, and that every GPU script at least compiles — a typo must not cost
    GPU minutes."""
    import inspect
    import os
    import py_compile
    from tests import kernel_cases as K
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def bind(fn, args):
        if fn is K.case_pair:
            return bind(args[0], args[1:])
        if fn is K.case_env:
            env, inner = args[0], args[1]
            assert all(isinstance(a, str) and isinstance(b, str) for a, b in env)
            return bind(inner, args[2:])
        inspect.signature(fn).bind(*args)
        return fn

    for fn, args in K.ALL_CASES + K.PENDING_CASES:
        assert callable(bind(fn, args))
    src = ""
    for f in ("gemm.cu", "attention.cu", "norm.cu", "misc.cu"):
        with open(os.path.join(here, "magicdance_b200", "csrc", f)) as fh:
            src += fh.read()
    with open(os.path.join(here, "magicdance_b200", "ops.py")) as fh:
        src += fh.read()
    switches = {a for fn, args in K.PENDING_CASES if fn is K.case_env for a, _ in args[0]}
    assert switches and all(f'"{sw}"' in src for sw in switches), switches
    assert not any(c in K.ALL_CASES for c in K.PENDING_CASES)  # unvalidated kernels stay out of the -m gpu suite
    for f in sorted(os.listdir(os.path.join(here, "scripts"))):
        if f.endswith(".py"):
            py_compile.compile(os.path.join(here, "scripts", f), doraise=True)
    py_compile.compile(os.path.join(here, "tests", "torch_gpu_baseline.py"), doraise=True)
    py_compile.compile(os.path.join(here, "bench.py"), doraise=True)


def test_switch_defaults_file_is_absent_and_would_not_override_the_environment(tmp_path, monkeypatch):
    """No opt-in is the default yet: magicdance_b200/switch_defaults.json does not exist.  When it does, its MDB_*
    entries are applied with setdefault (an explicit environment variable wins, other keys are ignored)."""
    import json
    import os
    import magicdance_b200 as M
    pkg = os.path.dirname(os.path.abspath(M.__file__))
    assert not os.path.exists(os.path.join(pkg, "switch_defaults.json")) and M.SWITCH_DEFAULTS == {}
    fake_pkg = tmp_path / "pkg"
    fake_pkg.mkdir()
    (fake_pkg / "switch_defaults.json").write_text(json.dumps({"MDB_T_A": 1, "MDB_T_B": "x", "PATH": "/nope"}))
    monkeypatch.setenv("MDB_T_B", "explicit")
    monkeypatch.delenv("MDB_T_A", raising=False)
    monkeypatch.setattr(M._os.path, "abspath", lambda p: str(fake_pkg / "__init__.py"))
    cfg = M._apply_switch_defaults()
    assert cfg == {"MDB_T_A": "1", "MDB_T_B": "x"}
    assert os.environ["MDB_T_A"] == "1" and os.environ["MDB_T_B"] == "explicit" and os.environ["PATH"] != "/nope"
    monkeypatch.delenv("MDB_T_A", raising=False)


def test_decide_defaults_rules(tmp_path):
    """scripts/decide_defaults.py: a switch is enabled only with green numerics AND a finite bench line whose final
    latent agrees with the default run AND a gain; a hung numerics log, a diverging latent or a slowdown keep it off;
    of two variants claiming the same slot the faster survives."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("decide_defaults", os.path.join(here, "scripts", "decide_defaults.py"))
    dd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dd)
    d = tmp_path

    def bench(name, value, fp=(0.5, 12.0), finite=True):
        (d / name).write_text(json.dumps({"value": value, "unit": "frame-steps/s", "finite": finite,
                                          "x_final_fingerprint": list(fp)}))

    def log(name, failing, total, extra=""):
        (d / name).write_text(f"device: B200\nok   case: err=1e-4\n{extra}group pending: {failing} failing of {total}\n")

    bench("pending_b1_default.json", 117.0)
    bench("pending_b8_default.json", 240.0)
    log("pending_pairq.log", 0, 17); bench("pending_b8_pair3.json", 262.0, fp=(0.5002, 12.004))     # green, +9 %
    log("pending_pairp.log", 0, 14); bench("pending_b8_pair2.json", 250.0)                          # green, +4 %: superseded
    log("pending_tmast.log", 1, 10, "FAIL gemm m=8192: err=3e-1\n"); bench("pending_b8_tmast.json", 300.0)  # wrong numerics
    (d / "pending_pairs.log").write_text("device: B200\nok   case\n")                            # hung: no summary line
    bench("pending_b1_pairs.json", 130.0)
    log("pending_attn4.log", 0, 12); bench("pending_b8_attn4.json", 250.0); bench("pending_b1_attn4.json", 110.0)  # slower at B=1
    log("pending_gnfused.log", 0, 7); bench("pending_b1_gnfused.json", 124.0, fp=(0.9, 30.0))      # latent differs
    bench("pending_b1_overlap.json", 126.0)                                                          # host-side, +7.7 %
    report, enabled = dd.judge(str(d), 2e-3, 0.01)
    assert set(enabled) == {"MDB_GEMM_PAIR=3", "MDB_BANK_OVERLAP=1"}, (enabled, report)
    (d / "pending_dropin_graph.log").write_text("eager loop : 21.500 ms/step (46.5 steps/s), finite=True\n"
                                                "graph replay: 7.100 ms/step (140.8 steps/s), finite=True\n"
                                                "eager vs eager (atomics order) 1.0e-03; graph vs eager 1.2e-03\nOK\n")
    assert "MDB_DROPIN_GRAPH=1" in dd.judge(str(d), 2e-3, 0.01)[1]
    (d / "pending_dropin_graph.log").write_text("eager loop : 21.5 ms/step\ngraph replay: 7.1 ms/step\nFAILED\n")
    assert "MDB_DROPIN_GRAPH=1" not in dd.judge(str(d), 2e-3, 0.01)[1]
    (d / "pending_dropin_graph.log").unlink()
    cfg = {}
    for env in enabled.values():
        cfg.update(env)
    assert cfg == {"MDB_GEMM_PAIR": "3", "MDB_BANK_OVERLAP": "1"}
    assert dd.numerics_ok(str(d / "missing.log")) == (False, "numerics log missing")
    assert not dd.fp_close([1.0, 2.0], [1.0, 2.1], 2e-3) and dd.fp_close([1.0, 2.0], [1.001, 2.001], 2e-3)


def test_polynomial_gelu_constants_in_the_kernel_source_are_accurate():
    """gelu_erf_poly_f (common.cuh, used by the opt-in GEGLU epilogues): the constants in the source, evaluated in
    float32 Horner arithmetic exactly as the kernel does, reproduce exact-erf GELU (attention.py:57) to 1e-4 absolute
    over [-12, 12] — far below the fp16 rounding of the result — and the fit script regenerates the same constants."""
    import importlib.util
    import os
    import re
    import numpy as np
    from scipy.special import erf
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "magicdance_b200", "csrc", "common.cuh")) as f:
        src = f.read()
    body = src[src.index("float gelu_erf_poly_f(float x)"):]
    body = body[:body.index("return fmaf(hx")]
    first = float(re.search(r"float p = ([-0-9.e+]+)f;", body).group(1))
    rest = [float(m) for m in re.findall(r"p = fmaf\(p, u, ([-0-9.e+]+)f\);", body)]
    assert len(rest) == 8
    x = np.linspace(-12, 12, 200001).astype(np.float32)
    z = np.clip(x * np.float32(0.70710678118654752), np.float32(-3), np.float32(3)).astype(np.float32)
    u = (z * z).astype(np.float32)
    p = np.full_like(u, np.float32(first))
    for c in rest:
        p = (p * u + np.float32(c)).astype(np.float32)
    hx = (np.float32(0.5) * x).astype(np.float32)
    got = (hx * (p * z).astype(np.float32) + hx).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) * 0.70710678118654752))
    assert np.abs(got - ref).max() < 1e-4
    spec = importlib.util.spec_from_file_location("fit_erf_poly", os.path.join(here, "scripts", "fit_erf_poly.py"))
    fit = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fit)
    coef = fit.fit(3.0, 8)
    assert np.allclose(coef[::-1], [first] + rest, rtol=1e-6, atol=1e-12)


def test_pipeline_caches_are_keyed_on_tensor_identity_and_version():
    """DenoisePipeline's per-sequence caches (a stale appearance bank would silently render the previous reference
    image): the bank is cached per (reference tensor, its in-place version, ddim index) and dropped as a whole when a
    different reference shows up; hint features are cached per caller-supplied frame key only."""
    import torch
    from magicdance_b200.pipeline import DenoisePipeline

    class Eng:
        device = torch.device("cpu")
        calls = {"app": 0, "proj": 0, "hint": 0}

        def appearance_write(self, ref, t, ctx):
            self.calls["app"] += 1
            return [ref.clone(), t.clone()]

        def project_bank(self, bank, batches):
            self.calls["proj"] += 1
            return ("kv", float(bank[0].sum()), int(bank[1][0]), batches)

        def hint_features(self, pose):
            self.calls["hint"] += 1
            return pose * 2

    eng = Eng()
    pipe = DenoisePipeline(eng, ddim_steps=50, scale=7.0, eta=0.0)
    ref, ctx = torch.ones(1, 4, 8, 8), torch.zeros(1, 77, 768)
    a = pipe.reference_bank(ref, ctx, 49)
    assert pipe.reference_bank(ref, ctx, 49) is a and eng.calls["app"] == 1          # hit
    b = pipe.reference_bank(ref, ctx, 48)
    assert eng.calls["app"] == 2 and b[2] == int(pipe.timesteps[48]) and a[2] == int(pipe.timesteps[49]) == 981
    ref.add_(1.0)                                                                    # same storage, new content
    c = pipe.reference_bank(ref, ctx, 49)
    assert eng.calls["app"] == 3 and c[1] == float(ref.sum()) != a[1]
    assert len(pipe._bank_cache) == 1                                                # the old sequence's banks are gone
    other = torch.ones(1, 4, 8, 8)
    pipe.reference_bank(other, ctx, 49)
    assert eng.calls["app"] == 4 and len(pipe._bank_cache) == 1
    two = torch.stack([other[0], other[0]])                                          # all rows the same image: row 0 only
    d = pipe.reference_bank(two, ctx, 49, first_only=True)
    assert d[3] == 1 and eng.calls["app"] == 5
    pose = torch.ones(1, 3, 16, 16)
    h1 = pipe.hint(pose, frame_key="f0")
    assert pipe.hint(pose, frame_key="f0") is h1 and eng.calls["hint"] == 1
    pipe.hint(pose, frame_key="f1")
    pipe.hint(pose)                                                                  # no key: never cached
    pipe.hint(pose)
    assert eng.calls["hint"] == 4
    pipe.clear_caches()
    assert not pipe._bank_cache and not pipe._hint_cache
    # keyed on a tensor's address (the drop-in sampler): the entry keeps that tensor alive, so the next frame's pose map
    # cannot be allocated at the same address and hit the previous frame's features
    pose_a = torch.ones(1, 3, 16, 16)
    ptr = pose_a.data_ptr()
    fa = pipe.hint(pose_a, frame_key=(ptr, pose_a._version, tuple(pose_a.shape)), keep_alive=pose_a)
    del pose_a
    pose_b = torch.zeros(1, 3, 16, 16)
    assert pose_b.data_ptr() != ptr
    fb = pipe.hint(pose_b, frame_key=(pose_b.data_ptr(), pose_b._version, tuple(pose_b.shape)), keep_alive=pose_b)
    assert float(fa.sum()) != float(fb.sum())
    for i in range(20):  # bounded: the oldest frames leave
        pipe.hint(pose_b, frame_key=("k", i))
    assert len(pipe._hint_cache) == pipe.HINT_CACHE_FRAMES


"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through the public operators, i.e.
through the C ABI of libxfscan.so, and is compared with
  (1) the committed golden fixtures produced by the reference's own Python (tests/golden/), and
  (2) the CPU oracle (oracle/) on seeded inputs following the reference's test distributions
      (models/selective_scan/test_selective_scan.py:153-179; shapes with H != W as models/csm_triton.py:524).
Contract (BASELINE.json): index routes bit exact; scan outputs and gradients within 1e-4 relative in fp32 and
2e-2 with bf16 inputs, relative = max|a-b| / max|b| per tensor (conftest.rel_err).
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import bf16_bits_to_f32, f16_bits_to_f32, rel_err

pytestmark = pytest.mark.gpu

TOL32 = 1e-4
TOL16 = 2e-2


def dev():
    return torch.device("cuda:0")


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    return x if dtype is None else x.to(dtype)


def bits(a, dtype):
    """uint16 bit patterns -> torch 16-bit tensor on the GPU"""
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).to(dev()).view(dtype)


def n(x):
    return x.detach().float().cpu().numpy()


@pytest.fixture(scope="module")
def xf():
    import xfmamba_b200
    from xfmamba_b200 import _lib
    assert _lib.lib().xfs_device_ok(0) == 0, "not an sm_100 device"
    return xfmamba_b200


# ================================================================================================ routes
@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("scans", [0, 1, 2])
def test_cross_scan_merge_golden(xf, golden, tag, scans):
    g = golden("csm")
    x = t(g[f"{tag}_x"]).requires_grad_(True)
    xs = xf.cross_scan_fn(x, True, True, False, scans)
    assert np.array_equal(n(xs), g[f"{tag}_s{scans}_xs"])
    xs.backward(t(g[f"{tag}_gx"]))
    assert np.array_equal(n(x.grad), g[f"{tag}_s{scans}_dx"])
    ys = t(g[f"{tag}_ys"]).requires_grad_(True)
    y = xf.cross_merge_fn(ys, True, True, False, scans)
    assert np.array_equal(n(y), g[f"{tag}_s{scans}_y"]), "merge add order must match the reference bit for bit"
    y.backward(t(g[f"{tag}_gy"]))
    assert np.array_equal(n(ys.grad), g[f"{tag}_s{scans}_dys"])
    # one_by_one
    B, _, C, H, W = g[f"{tag}_ys"].shape
    a = xf.cross_scan_fn(t(g[f"{tag}_ys"]), True, True, True, scans)
    assert np.array_equal(n(a).reshape(-1), g[f"{tag}_s{scans}_xs1b1"].reshape(-1))
    m = xf.cross_merge_fn(t(g[f"{tag}_ys"]), True, True, True, scans)
    assert np.array_equal(n(m).reshape(-1), g[f"{tag}_s{scans}_y1b1"].reshape(-1))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cross_scan_merge_bf16_golden(xf, golden, tag):
    g = golden("csm")
    xb = t(g[f"{tag}_x"]).to(torch.bfloat16)
    xs = xf.cross_scan_fn(xb)
    assert np.array_equal(xs.view(torch.int16).cpu().numpy().view(np.uint16), g[f"{tag}_bf16_xs"])
    yb = xf.cross_merge_fn(t(g[f"{tag}_ys"]).to(torch.bfloat16))
    assert np.array_equal(yb.view(torch.int16).cpu().numpy().view(np.uint16), g[f"{tag}_bf16_y"])


@pytest.mark.parametrize("shape", [(2, 5, 56, 57), (1, 3, 57, 58), (3, 2, 1, 7), (1, 1, 33, 1), (2, 7, 14, 14), (1, 2, 64, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_cross_scan_merge_oracle(xf, shape, dtype):
    B, C, H, W = shape
    rng = np.random.default_rng(hash(shape) % 2**32)
    x = rng.standard_normal(shape, dtype=np.float32)
    ys = rng.standard_normal((B, 4, C, H, W), dtype=np.float32)
    if dtype == torch.float16:
        x, ys = x.astype(np.float16), ys.astype(np.float16)
    for scans in (0, 1, 2):
        xs = xf.cross_scan_fn(t(x), scans=scans)
        assert np.array_equal(xs.cpu().numpy(), oracle.cross_scan(x, scans))
        if dtype == torch.float32:
            y = xf.cross_merge_fn(t(ys), scans=scans)
            assert np.array_equal(n(y), oracle.cross_merge(ys.reshape(B, 4, C, H * W), H, W, scans))


def test_cross_scan_channel_last_layouts(xf):
    """layout flags of models/csm_triton.py:22-85 (not used by XFMamba, kept for API parity): compare with permutes"""
    torch.manual_seed(0)
    x = torch.randn(2, 3, 5, 4, device=dev())
    ref = xf.cross_scan_fn(x)                                    # (B,4,C,L)
    assert torch.equal(xf.cross_scan_fn(x.permute(0, 2, 3, 1).contiguous(), False, True), ref)
    assert torch.equal(xf.cross_scan_fn(x, True, False), ref.permute(0, 3, 1, 2))
    ys = torch.randn(2, 4, 3, 5, 4, device=dev())
    refm = xf.cross_merge_fn(ys)                                 # (B,C,L)
    assert torch.equal(xf.cross_merge_fn(ys.permute(0, 3, 4, 1, 2).contiguous(), True, False), refm)
    assert torch.equal(xf.cross_merge_fn(ys, False, True), refm.permute(0, 2, 1))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_swap_golden(xf, golden, tag):
    g = golden("swap")
    x, x2 = t(g[f"{tag}_x"]).requires_grad_(True), t(g[f"{tag}_x2"]).requires_grad_(True)
    xs = xf.SwappingScan_multiview.apply(x, x2)
    assert np.array_equal(n(xs), g[f"{tag}_xs"])
    xs.backward(t(g[f"{tag}_gxs"]))
    assert np.array_equal(n(x.grad), g[f"{tag}_dx"]) and np.array_equal(n(x2.grad), g[f"{tag}_dx2"])
    ys = t(g[f"{tag}_ys"]).requires_grad_(True)
    y, y2 = xf.SwappingMerge_multiview.apply(ys)
    assert np.array_equal(n(y), g[f"{tag}_y"]) and np.array_equal(n(y2), g[f"{tag}_y2"])
    (y * t(g[f"{tag}_gy"])).sum().add((y2 * t(g[f"{tag}_gy2"])).sum()).backward()
    assert np.array_equal(n(ys.grad), g[f"{tag}_dys"])
    # the exact adjoint (opt-in) differs from the reference's as-written backward on even channels only
    xa, xb = t(g[f"{tag}_x"]).requires_grad_(True), t(g[f"{tag}_x2"]).requires_grad_(True)
    xf.swapping_scan(xa, xb, exact_adjoint=True).backward(t(g[f"{tag}_gxs"]))
    gx = g[f"{tag}_gxs"]
    B, C, H, W = g[f"{tag}_x"].shape
    exp = np.where((np.arange(C) % 2 == 0)[None, :, None], gx[:, 1], gx[:, 0]).reshape(B, C, H, W)
    assert np.array_equal(n(xa.grad), exp)


# ================================================================================================ selective scan
def _golden_case(g, name):
    meta = [int(v) for v in g[f"{name}_meta"]]
    Bsz, K, Cd, N, L, has_D, has_bias, softplus, oflex, dt = meta
    tdt = {0: torch.float32, 1: torch.bfloat16, 2: torch.float16}[dt]
    mk = (lambda a: t(a)) if dt == 0 else (lambda a: bits(a, tdt))
    ten = {k: mk(g[f"{name}_{k}"]) for k in ("u", "delta", "B", "C")}
    ten["A"] = t(g[f"{name}_A"])
    ten["D"] = t(g[f"{name}_D"]) if has_D else None
    ten["delta_bias"] = t(g[f"{name}_delta_bias"]) if has_bias else None
    return ten, bool(softplus), bool(oflex), dt


GRAD_KEYS = ["u", "delta", "A", "B", "C", "D", "delta_bias"]


@pytest.mark.parametrize("name", ["s1", "s2", "s3", "s4", "s5", "s6", "h1", "h2", "h3"])
def test_selective_scan_golden(xf, golden, name):
    g = golden("scan")
    ten, softplus, oflex, dt = _golden_case(g, name)
    leaves = {k: (v.clone().requires_grad_(True) if v is not None else None) for k, v in ten.items()}
    out = xf.selective_scan_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                               leaves["delta_bias"], softplus, oflex)
    conv = {0: lambda a: a, 1: bf16_bits_to_f32, 2: f16_bits_to_f32}[dt]
    ref_out = g[f"{name}_out"] if (oflex or dt == 0) else conv(g[f"{name}_out"])
    assert out.dtype == (torch.float32 if oflex else ten["u"].dtype)          # models/csms6s.py:68
    tol = TOL32 if dt == 0 else TOL16
    assert rel_err(n(out), ref_out) < tol
    out.backward(t(g[f"{name}_dout"]).to(out.dtype))
    for k in GRAD_KEYS:
        if leaves[k] is None:
            continue
        ref = g[f"{name}_d{k}"]
        ref = ref if ref.dtype == np.float32 else conv(ref)
        assert leaves[k].grad.dtype == leaves[k].dtype
        assert rel_err(n(leaves[k].grad), ref) < tol, f"{name}: d{k}"


def _rand_scan(rng, Bsz, K, Cd, N, L):
    """input distributions of models/selective_scan/test_selective_scan.py:157-179"""
    KD = K * Cd
    f = lambda *s: rng.standard_normal(s, dtype=np.float32)
    r = lambda *s: rng.random(s, dtype=np.float32)
    return dict(u=f(Bsz, KD, L), delta=0.5 * r(Bsz, KD, L), A=-0.5 * r(KD, N), B=f(Bsz, K, N, L), C=f(Bsz, K, N, L),
                D=f(KD), delta_bias=0.5 * r(KD), dout=f(Bsz, KD, L))


@pytest.mark.parametrize("seqlen", [64, 128, 256, 372, 512, 784, 1024, 1134, 2048, 4096])
@pytest.mark.parametrize("groups", [1, 2])
def test_selective_scan_reference_grid_fp32(xf, seqlen, groups):
    """the reference's own pytest grid (test_selective_scan.py:137-156): batch 2, dim 24, dstate 8, seed 0"""
    rng = np.random.default_rng(seqlen * 10 + groups)
    c = _rand_scan(rng, 2, groups, 24 // groups, 8, seqlen)
    has_D, has_bias, softplus = True, True, True
    leaves = {k: t(c[k]).requires_grad_(True) for k in GRAD_KEYS}
    out = xf.selective_scan_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                               leaves["delta_bias"], softplus, True)
    ref = oracle.selective_scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["delta_bias"], softplus, "f64")
    assert rel_err(n(out), ref) < TOL32
    out.backward(t(c["dout"]))
    grads = oracle.selective_scan_bwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["delta_bias"], c["dout"], softplus, "f64")
    for k, gr in zip(GRAD_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL32, f"d{k}"


@pytest.mark.parametrize("seqlen", [72, 256, 264, 784, 3136, 4104])
@pytest.mark.parametrize("groups", [1, 4])
@pytest.mark.parametrize("softplus", [True, False])
def test_selective_scan_dstate1_vs_oracle(xf, seqlen, groups, softplus):
    """d_state = 1, fp32, L % 8 == 0: what `selective_scan_fn` sees in every VMamba-style SS2D block (K = 4 groups,
    models/fusion_vmamba.py:1170-1174), served by the 256-bit / packed backward kernel (csrc/selective_scan.cu sscan_n1_bwd_kernel).
    Short chunk last (264, 4104), one chunk (72, 256), softplus outliers (series below 2^-6, identity above 20) mixed in."""
    rng = np.random.default_rng(seqlen * 100 + groups * 2 + int(softplus))
    c = _rand_scan(rng, 2, groups, 12 // groups, 1, seqlen)
    if softplus:
        pick = rng.integers(0, 8, size=c["delta"].shape)
        c["delta"] = np.where(pick == 0, -11.0, np.where(pick == 1, 21.0, c["delta"])).astype(np.float32)
        c["A"] = (0.05 * c["A"]).astype(np.float32)
    leaves = {k: t(c[k]).requires_grad_(True) for k in GRAD_KEYS}
    out = xf.selective_scan_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                               leaves["delta_bias"], softplus, True)
    ref = oracle.selective_scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["delta_bias"], softplus, "f64")
    assert rel_err(n(out), ref) < TOL32
    out.backward(t(c["dout"]))
    grads = oracle.selective_scan_bwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["delta_bias"], c["dout"], softplus, "f64")
    for k, gr in zip(GRAD_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL32, f"d{k}"


@pytest.mark.parametrize("flags", [(False, False, False), (True, False, True), (False, True, False), (False, True, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_selective_scan_flags_and_dtypes(xf, flags, dtype):
    has_D, has_bias, softplus = flags
    rng = np.random.default_rng(7)
    c = _rand_scan(rng, 2, 2, 5, 3, 300)
    if dtype != torch.float32:   # quantise the 16-bit tensors first so that oracle and kernel see identical inputs
        for k in ("u", "delta", "B", "C"):
            c[k] = t(c[k]).to(dtype).float().cpu().numpy()
    D = c["D"] if has_D else None
    bias = c["delta_bias"] if has_bias else None
    leaves = {k: t(c[k]).to(dtype if k in ("u", "delta", "B", "C") else torch.float32).requires_grad_(True) for k in GRAD_KEYS}
    out = xf.selective_scan_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"],
                               leaves["D"] if has_D else None, leaves["delta_bias"] if has_bias else None, softplus, True)
    assert out.dtype == torch.float32
    ref = oracle.selective_scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], D, bias, softplus, "f64")
    assert rel_err(n(out), ref) < TOL32          # fp32 arithmetic inside: 16-bit inputs only quantise the inputs
    out.backward(t(c["dout"]))
    grads = oracle.selective_scan_bwd(c["u"], c["delta"], c["A"], c["B"], c["C"], D, bias, c["dout"], softplus, "f64")
    tol = TOL32 if dtype == torch.float32 else TOL16       # 16-bit grads are rounded to the input dtype on store
    for k, gr in zip(GRAD_KEYS, grads):
        if gr is None:
            assert leaves[k].grad is None
            continue
        assert rel_err(n(leaves[k].grad), gr) < tol, f"d{k}"


def test_selective_scan_softplus_threshold_and_extremes(xf):
    """x > 20 passes through; very negative x underflows to 0 without NaN (models/csms6s.py:49-50)"""
    rng = np.random.default_rng(3)
    c = _rand_scan(rng, 1, 1, 4, 2, 96)
    c["delta"] = (rng.random((1, 4, 96), dtype=np.float32) * 140.0 - 70.0)
    out = xf.selective_scan_fn(t(c["u"]), t(c["delta"]), t(c["A"]), t(c["B"]), t(c["C"]), t(c["D"]), t(c["delta_bias"]), True, True)
    ref = oracle.selective_scan_fwd(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["delta_bias"], True, "f64")
    assert torch.isfinite(out).all()
    assert rel_err(n(out), ref) < TOL32


def test_selective_scan_argument_errors(xf):
    d = dev()
    u = torch.randn(2, 6, 16, device=d)
    A = torch.randn(6, 4, device=d)
    Bm = torch.randn(2, 2, 4, 16, device=d)
    with pytest.raises(RuntimeError, match="float32"):
        xf.selective_scan_fn(u, u, A.half(), Bm, Bm)
    with pytest.raises(RuntimeError, match="share one dtype"):
        xf.selective_scan_fn(u, u.half(), A, Bm, Bm)
    with pytest.raises(RuntimeError, match="dividable"):
        xf.selective_scan_fn(u, u, A, torch.randn(2, 4, 4, 16, device=d), torch.randn(2, 4, 4, 16, device=d))
    with pytest.raises(RuntimeError, match="delta"):
        xf.selective_scan_fn(u, u[:, :, :8], A, Bm, Bm)
    # empty batch is a no-op, as in torch
    e = xf.selective_scan_fn(u[:0], u[:0], A, Bm[:0], Bm[:0])
    assert e.shape == (0, 6, 16)


# ================================================================================================ fused SS2D core
def _rand_ss2d(rng, Bsz, D, N, H, W, model_like=False):
    L = H * W
    f = lambda *s: rng.standard_normal(s, dtype=np.float32)
    r = lambda *s: rng.random(s, dtype=np.float32)
    c = dict(x=f(Bsz, D, H, W), delta=0.5 * r(Bsz, 4 * D, L), A=-0.5 * r(4 * D, N), Bs=f(Bsz, 4, N, L), Cs=f(Bsz, 4, N, L),
             Ds=f(4 * D), delta_bias=0.5 * r(4 * D), dy=f(Bsz, D, L))
    if model_like:    # A = -(1..N), dt in [1e-3, 1e-1] (models/fusion_vmamba.py:303-328)
        c["A"] = -np.tile(np.arange(1, N + 1, dtype=np.float32), (4 * D, 1))
        dt = np.exp(r(4 * D) * (np.log(0.1) - np.log(0.001)) + np.log(0.001)).astype(np.float32)
        c["delta_bias"] = (dt + np.log(-np.expm1(-dt))).astype(np.float32)
        c["delta"] = 0.3 * f(Bsz, 4 * D, L)
    return c


SS2D_KEYS = ["x", "delta", "A", "Bs", "Cs", "Ds", "delta_bias"]


def _run_ss2d(xf, c, dtype=torch.float32, oflex=True, has_D=True, has_bias=True):
    leaves = {k: t(c[k]).to(dtype if k in ("x", "delta", "Bs", "Cs") else torch.float32).requires_grad_(True) for k in SS2D_KEYS}
    y = xf.ss2d_scan(leaves["x"], leaves["delta"], leaves["A"], leaves["Bs"], leaves["Cs"],
                     leaves["Ds"] if has_D else None, leaves["delta_bias"] if has_bias else None, True, oflex)
    y.backward(t(c["dy"]).to(y.dtype))
    return y, leaves


def test_ss2d_golden_core(xf, golden):
    """operator-level tensors recorded inside the reference's SS2Dv2.forward_corev2 (models/fusion_vmamba.py:1176-1181)"""
    g = golden("cores")
    x = t(g["ss2d_x"])
    y = xf.ss2d_scan(x, t(g["ss2d_dts"]), t(g["ss2d_As"]), t(g["ss2d_Bs"]), t(g["ss2d_Cs"]), t(g["ss2d_Ds"]), t(g["ss2d_delta_bias"]))
    assert rel_err(n(y), g["ss2d_ymerged"]) < TOL32
    # and the unfused drop-in chain gives the same tensors the reference recorded
    B, D, H, W = x.shape
    xs = xf.cross_scan_fn(x)
    assert np.array_equal(n(xs).reshape(B, 4 * D, H * W), g["ss2d_us"])
    ys = xf.selective_scan_fn(xs.view(B, -1, H * W), t(g["ss2d_dts"]), t(g["ss2d_As"]), t(g["ss2d_Bs"]), t(g["ss2d_Cs"]),
                              t(g["ss2d_Ds"]), t(g["ss2d_delta_bias"]), True, True)
    assert rel_err(n(ys), g["ss2d_ys"].reshape(B, 4 * D, H * W)) < TOL32
    assert rel_err(n(xf.cross_merge_fn(ys.view(B, 4, D, H, W))), g["ss2d_ymerged"]) < TOL32


@pytest.mark.parametrize("shape", [
    (2, 4, 1, 6, 5), (1, 3, 1, 14, 14), (2, 6, 1, 28, 28), (1, 5, 1, 17, 19), (1, 2, 1, 56, 57), (2, 2, 1, 7, 7),
    (1, 2, 4, 7, 7), (2, 3, 16, 7, 7), (1, 2, 3, 16, 17), (1, 1, 1, 1, 300), (1, 2, 1, 32, 8),
    # one-chunk shapes (64 < L <= 256, L % 4 == 0): the warp-per-channel kernels of csrc/ss2d_mid.cu
    (2, 5, 1, 14, 14), (1, 4, 1, 16, 16), (2, 3, 1, 9, 8), (1, 2, 1, 10, 20), (1, 6, 1, 12, 11), (3, 1, 1, 4, 17), (1, 9, 1, 13, 20),
])
@pytest.mark.parametrize("model_like", [False, True])
def test_ss2d_fused_oracle_fp32(xf, shape, model_like):
    Bsz, D, N, H, W = shape
    rng = np.random.default_rng(abs(hash(shape)) % 2**32)
    c = _rand_ss2d(rng, Bsz, D, N, H, W, model_like)
    assert xf.ss2d_fused_supported(D, N, H, W, torch.float32, True)
    y, leaves = _run_ss2d(xf, c)
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < TOL32
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    for k, gr in zip(SS2D_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL32, f"d{k}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("oflex", [True, False])
@pytest.mark.parametrize("hw", [(16, 24), (14, 14), (9, 12), (7, 7)])      # general / one-chunk (L % 8 != 0 and == 0) / short
def test_ss2d_fused_16bit(xf, dtype, oflex, hw):
    rng = np.random.default_rng(11)
    c = _rand_ss2d(rng, 2, 5, 1, hw[0], hw[1])
    for k in ("x", "delta", "Bs", "Cs"):
        c[k] = t(c[k]).to(dtype).float().cpu().numpy()
    y, leaves = _run_ss2d(xf, c, dtype, oflex)
    assert y.dtype == (torch.float32 if oflex else dtype)
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < (TOL32 if oflex else TOL16)
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    for k, gr in zip(SS2D_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL16, f"d{k}"


def test_ss2d_fused_optional_params(xf):
    rng = np.random.default_rng(5)
    c = _rand_ss2d(rng, 1, 3, 1, 9, 11)
    y, leaves = _run_ss2d(xf, c, has_D=False, has_bias=False)
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], None, None, True, "f64")
    assert rel_err(n(y), ref) < TOL32
    assert leaves["Ds"].grad is None and leaves["delta_bias"].grad is None


def test_ss2d_full_size_properties(xf):
    """config-2 shape (56x56, D=192, N=1, K=4) at a size the CPU oracle cannot reach: size-independent properties.
    (1) fused == composition of the three stand-alone CUDA operators, forward and all gradients;
    (2) route symmetry: rotating the image by 180 degrees and exchanging the parameters of routes (0,2) and (1,3)
        rotates the output (route 2/3 are the flips of routes 0/1, models/csm_triton.py:29)."""
    torch.manual_seed(0)
    d = dev()
    Bsz, D, H, W, N = 8, 192, 56, 56, 1
    L = H * W
    x = torch.randn(Bsz, D, H, W, device=d)
    delta = 0.5 * torch.rand(Bsz, 4 * D, L, device=d)
    A = -0.5 * torch.rand(4 * D, N, device=d)
    Bs, Cs = torch.randn(Bsz, 4, N, L, device=d), torch.randn(Bsz, 4, N, L, device=d)
    Ds, bias = torch.randn(4 * D, device=d), 0.5 * torch.rand(4 * D, device=d)
    dy = torch.randn(Bsz, D, L, device=d)

    def run(fused):
        lv = [v.clone().requires_grad_(True) for v in (x, delta, A, Bs, Cs, Ds, bias)]
        if fused:
            y = xf.ss2d_scan(*lv)
        else:
            xs = xf.cross_scan_fn(lv[0]).view(Bsz, -1, L)
            ys = xf.selective_scan_fn(xs, *lv[1:], True, True)
            y = xf.cross_merge_fn(ys.view(Bsz, 4, D, H, W))
        y.backward(dy)
        return y, [v.grad for v in lv]

    yf, gf = run(True)
    yu, gu = run(False)
    assert rel_err(n(yf), n(yu)) < TOL32
    for k, a, b in zip(SS2D_KEYS, gf, gu):
        assert rel_err(n(a), n(b)) < TOL32, f"d{k}"

    def swap_routes(v, lead):          # exchange route blocks (0<->2, 1<->3) along the dim that holds 4 routes
        shp = v.shape
        v4 = v.reshape(*shp[:lead], 4, -1)
        return v4[(slice(None),) * lead + ([2, 3, 0, 1],)].reshape(shp)

    y_rot = xf.ss2d_scan(x.flip(2, 3), swap_routes(delta, 1), swap_routes(A, 0), swap_routes(Bs, 1), swap_routes(Cs, 1),
                         swap_routes(Ds, 0), swap_routes(bias, 0))
    assert rel_err(n(y_rot), n(yf.flip(2))) < TOL32


@pytest.mark.parametrize("shape", [(1, 6, 1, 128, 128), (1, 3, 1, 120, 130), (2, 2, 1, 127, 128)])
def test_ss2d_hires_inference_is_one_fused_launch(xf, shape):
    """BASELINE config 5 (512^2 input: stage-1 L = 16384): without checkpoints the forward is ONE fused launch even though four
    64 KB image buffers exceed shared memory (routes 0 / 2 take x from their ring slots, csrc/ss2d_ring_fwd.cu kBig); ragged
    sizes put the short chunk first on the flipped routes.  With gradients the three stand-alone operators still run."""
    from xfmamba_b200 import _lib, fusion_ops
    Bsz, D, N, H, W = shape
    rng = np.random.default_rng(abs(hash(shape)) % 2**32)
    c = _rand_ss2d(rng, Bsz, D, N, H, W, model_like=True)
    assert fusion_ops.ss2d_fused_supported(D, N, H, W, torch.float32, backward=False)
    assert not fusion_ops.ss2d_fused_supported(D, N, H, W, torch.float32, backward=True)
    args = [t(c[k]) for k in SS2D_KEYS]
    before = _lib.launch_count()
    with torch.no_grad():
        y = xf.ss2d_scan(*args)
    assert _lib.launch_count() - before == 1
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < TOL32


@pytest.mark.parametrize("shape", [(1, 3, 1, 20, 24), (2, 2, 1, 13, 20), (1, 4, 1, 14, 14)])
def test_ss2d_softplus_outliers_vs_oracle(xf, shape):
    """softplus regimes the fast lg2(1 + e) form does not cover, mixed inside single chunks: delta + bias far below zero (the
    e < 2^-6 series; models/csms6s.py:49-50 computes log1p(exp(x))), above the threshold 20 (identity, sigmoid = 1) and
    ordinary values -- the multi-chunk lane-checkpoint kernels (L = 480, 260: 256-bit and 128-bit rows) and the one-chunk
    kernel repair those under a warp vote, the backward after re-reading B."""
    Bsz, D, N, H, W = shape
    L = H * W
    rng = np.random.default_rng(abs(hash(shape)) % 2**32)
    c = _rand_ss2d(rng, Bsz, D, N, H, W)
    pick = rng.integers(0, 4, size=c["delta"].shape)
    c["delta"] = np.where(pick == 0, -12.0 + 2.0 * rng.random(c["delta"].shape),
                          np.where(pick == 1, 20.5 + 3.0 * rng.random(c["delta"].shape), c["delta"])).astype(np.float32)
    c["A"] = (0.02 * c["A"]).astype(np.float32)                       # keep exp(dt A) away from underflow at dt ~ 23
    y, leaves = _run_ss2d(xf, c)
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < TOL32
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    for k, gr in zip(SS2D_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL32, f"d{k}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ss2d_config2_shape_vs_oracle(xf, dtype):
    """BASELINE config 2 at its exact per-image shape (D = 192, 56x56, N = 1, K = 4; 8 images) against the CPU oracle:
    forward and all seven gradients, fp32 and bf16 inputs (fp32 out).  Also per-row relative checks of the delta-scale
    quantities (a tensor-scale metric hides errors of small dt / ddelta values)."""
    rng = np.random.default_rng(2)
    Bsz, D, N, H, W = 8, 192, 1, 56, 56
    c = _rand_ss2d(rng, Bsz, D, N, H, W)
    if dtype != torch.float32:
        for k in ("x", "delta", "Bs", "Cs"):
            c[k] = t(c[k]).to(dtype).float().cpu().numpy()
    tol = TOL32 if dtype == torch.float32 else TOL16
    y, leaves = _run_ss2d(xf, c, dtype, True)
    assert y.dtype == torch.float32
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < TOL32
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    for k, gr in zip(SS2D_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < tol, f"d{k}"
    if dtype == torch.float32:      # row-wise: every (b, k*D + d) row of ddelta and every (b, d) row of y on its own scale
        dd, ddr = n(leaves["delta"].grad), grads[1]
        row = np.max(np.abs(dd - ddr), axis=-1) / np.maximum(np.max(np.abs(ddr), axis=-1), 1e-30)
        assert float(row.max()) < 2e-4, float(row.max())
        yy = n(y)
        row = np.max(np.abs(yy - ref), axis=-1) / np.maximum(np.max(np.abs(ref), axis=-1), 1e-30)
        assert float(row.max()) < 2e-4, float(row.max())


@pytest.mark.parametrize("shape", [(1, 64, 1, 64, 64), (1, 64, 1, 128, 128), (2, 3, 1, 40, 44), (1, 2, 1, 20, 13), (1, 3, 1, 2, 130)])
@pytest.mark.parametrize("model_like", [False, True])
def test_ss2d_long_rows_vs_oracle(xf, shape, model_like):
    """long-sequence path (512^2 inputs: stage-1 L = 16384, stage-2 L = 4096) and ragged multi-chunk shapes, whichever kernel
    or composition serves them: forward and all gradients against the oracle.  model_like puts dt in [1e-3, 1e-1], where
    softplus needs its small-argument branch."""
    Bsz, D, N, H, W = shape
    rng = np.random.default_rng(abs(hash(shape)) % 2**32)
    c = _rand_ss2d(rng, Bsz, D, N, H, W, model_like)
    y, leaves = _run_ss2d(xf, c)
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < TOL32
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    for k, gr in zip(SS2D_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL32, f"d{k}"


def test_ss2d_softplus_branches_mixed_in_one_chunk(xf):
    """the TMA-fed kernels take a chunk-uniform shortcut when no element of a chunk needs the small-argument series or the
    x > 20 identity of softplus (csrc/ss2d_ring.cuh); mix all three regimes inside single chunks and across chunks."""
    rng = np.random.default_rng(17)
    Bsz, D, N, H, W = 2, 4, 1, 24, 28
    c = _rand_ss2d(rng, Bsz, D, N, H, W)
    L = H * W
    d = c["delta"]
    d[:, :, 5::37] = -9.0 + rng.random(d[:, :, 5::37].shape, dtype=np.float32)         # e^x ~ 1e-4: series branch
    d[:, :, 11::53] = 21.0 + 3 * rng.random(d[:, :, 11::53].shape, dtype=np.float32)   # x > 20: identity branch
    d[0, 1, 300:560] = -12.0                                                            # a whole chunk of tiny dt
    d[1, 2, :] = 0.25                                                                   # a whole row on the fast path
    c["A"] = -0.01 * rng.random((4 * D, N), dtype=np.float32)                          # keep exp(dt A) away from 0 for dt ~ 24
    y, leaves = _run_ss2d(xf, c)
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < TOL32
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    for k, gr in zip(SS2D_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL32, f"d{k}"


def test_ss2d_unfused_fallback_large_L(xf):
    """L too large for the fused working set -> the operator composes the stand-alone kernels (still CUDA)"""
    rng = np.random.default_rng(9)
    H, W = 128, 128
    assert not xf.ss2d_fused_supported(2, 1, H, W, torch.float32, True)
    c = _rand_ss2d(rng, 1, 2, 1, H, W)
    y, leaves = _run_ss2d(xf, c)
    ref = oracle.ss2d_fwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    assert rel_err(n(y), ref) < TOL32
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    for k, gr in zip(SS2D_KEYS, grads):
        assert rel_err(n(leaves[k].grad), gr) < TOL32, f"d{k}"


def test_ss2d_scan_4d_delta_and_no_grad(xf):
    """delta given as (B, 4, D, L) gets its gradient back in that shape; under torch.no_grad() nothing is recorded even though
    the parameters require grad"""
    rng = np.random.default_rng(3)
    Bsz, D, N, H, W = 2, 3, 1, 20, 16
    c = _rand_ss2d(rng, Bsz, D, N, H, W)
    lv = {k: t(c[k]).requires_grad_(True) for k in SS2D_KEYS}
    d4 = lv["delta"].detach().view(Bsz, 4, D, H * W).clone().requires_grad_(True)
    y = xf.ss2d_scan(lv["x"], d4, lv["A"], lv["Bs"], lv["Cs"], lv["Ds"], lv["delta_bias"])
    y.backward(t(c["dy"]))
    assert d4.grad.shape == d4.shape
    grads = oracle.ss2d_bwd(c["x"], c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], c["dy"], True, "f64")
    assert rel_err(n(d4.grad).reshape(Bsz, 4 * D, -1), grads[1]) < TOL32
    with torch.no_grad():
        y2 = xf.ss2d_scan(lv["x"], lv["delta"], lv["A"], lv["Bs"], lv["Cs"], lv["Ds"], lv["delta_bias"])
        ys = xf.selective_scan_fn(xf.cross_scan_fn(lv["x"]).view(Bsz, -1, H * W), lv["delta"], lv["A"], lv["Bs"], lv["Cs"], lv["Ds"],
                                  lv["delta_bias"], True, True)
    assert y2.grad_fn is None and not y2.requires_grad and ys.grad_fn is None
    assert rel_err(n(y2), n(y)) < 1e-6


# ================================================================================================ fusion cores
def _core_module(g, prefix, device):
    """stands in for the reference module object: exactly the attributes the fused cores read (models/fusion_vmamba.py
    :470-478, :801-808), filled with the parameters the reference recorded in tests/golden/cores.npz"""
    import types
    torch.backends.cudnn.allow_tf32 = False          # x_proj is a library 1x1 convolution: keep it in fp32 for the 1e-4 bar
    torch.backends.cuda.matmul.allow_tf32 = False
    D = g[prefix + "out_norm.weight"].shape[0]
    norm = torch.nn.LayerNorm(D).to(device)
    with torch.no_grad():
        norm.weight.copy_(torch.from_numpy(g[prefix + "out_norm.weight"]))
        norm.bias.copy_(torch.from_numpy(g[prefix + "out_norm.bias"]))
    mod = types.SimpleNamespace(out_norm=norm, channel_first=bool(int(g[prefix + "channel_first"])), x_proj_bias=None)
    for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds"):
        setattr(mod, k, torch.from_numpy(g[prefix + k]).to(device).requires_grad_(True))
    return mod


def _check_core_grads(g, prefix, mod, inputs, tol):
    for name, leaf in inputs.items():
        assert rel_err(n(leaf.grad), g[prefix + "d" + name]) < tol, "d" + name
    for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds"):
        assert rel_err(n(getattr(mod, k).grad), g[prefix + "grad_" + k]) < tol, k
    assert rel_err(n(mod.out_norm.weight.grad), g[prefix + "grad_out_norm.weight"]) < tol
    assert rel_err(n(mod.out_norm.bias.grad), g[prefix + "grad_out_norm.bias"]) < tol


@pytest.mark.parametrize("via_patch", [False, True])
def test_shallow_fusion_core_golden(xf, golden, via_patch):
    """ShallowFuse_SS2Dv4.forward_corev2 (models/fusion_vmamba.py:777-845) as the reference recorded it: both outputs, the
    gradients of both inputs and of every parameter -- through model.shallow_fuse_core and through the function
    patch.install(fused=True) puts in the reference class."""
    import xfmamba_b200.patch as xfpatch
    from xfmamba_b200 import model as M
    g = golden("cores")
    mod = _core_module(g, "shallow_", dev())
    x, x2 = t(g["shallow_x"]).requires_grad_(True), t(g["shallow_x2"]).requires_grad_(True)
    if via_patch:
        y, y2 = xfpatch._fused_shallow_core(mod, x, x2, force_fp32=False, no_einsum=True)
    else:
        ys = M.shallow_fuse_core(x, x2, mod.x_proj_weight, mod.dt_projs_weight, mod.dt_projs_bias, mod.A_logs, mod.Ds)
        y, y2 = (xfpatch._finish(mod, v, x) for v in ys)
    assert rel_err(n(y), g["shallow_y"]) < TOL32 and rel_err(n(y2), g["shallow_y2"]) < TOL32
    ((y * t(g["shallow_gy"])).sum() + (y2 * t(g["shallow_gy2"])).sum()).backward()
    _check_core_grads(g, "shallow_", mod, dict(x=x, x2=x2), 2e-4)


@pytest.mark.parametrize("via_patch", [False, True])
def test_deep_fusion_core_golden(xf, golden, via_patch):
    """Cross_SS2Dv5.forward_corev2 (models/fusion_vmamba.py:446-578): three streams, one parameter set, the view streams
    reading Cs of the fused stream (:536-538, 567-569) -- outputs, input gradients and every parameter gradient."""
    import xfmamba_b200.patch as xfpatch
    from xfmamba_b200 import model as M
    g = golden("cores")
    mod = _core_module(g, "deep_", dev())
    x, x2, xfu = (t(g["deep_" + k]).requires_grad_(True) for k in ("x", "x2", "xf"))
    if via_patch:
        y, y2, yf = xfpatch._fused_cross_core(mod, x, x2, xfu, force_fp32=False, no_einsum=True)
    else:
        ys = M.cross_fuse_core(x, x2, xfu, mod.x_proj_weight, mod.dt_projs_weight, mod.dt_projs_bias, mod.A_logs, mod.Ds)
        y, y2, yf = (xfpatch._finish(mod, v, x) for v in ys)
    for got, key in ((y, "deep_y"), (y2, "deep_y2"), (yf, "deep_yf")):
        assert rel_err(n(got), g[key]) < TOL32, key
    ((y * t(g["deep_gy"])).sum() + (y2 * t(g["deep_gy2"])).sum() + (yf * t(g["deep_gyf"])).sum()).backward()
    _check_core_grads(g, "deep_", mod, dict(x=x, x2=x2, xf=xfu), 2e-4)


@pytest.mark.parametrize("shape", [(2, 5, 16, 7, 7), (1, 8, 4, 5, 6), (2, 3, 1, 8, 8), (1, 33, 16, 7, 7)])
def test_cross_ss2d_x3_single_launch_vs_oracle(xf, shape):
    """xfs_cross_ss2d_x3: three SS2D streams (one parameter set, own x / delta / Bs, ONE shared Cs) in one forward and one
    backward launch, against the oracle run stream by stream; the shared-Cs and parameter gradients sum over the streams
    (models/fusion_vmamba.py:536-538, 567-569)."""
    from xfmamba_b200 import _lib, fusion_ops
    Bsz, D, N, H, W = shape
    L = H * W
    rng = np.random.default_rng(abs(hash(shape)) % 2**32)
    cs = [_rand_ss2d(rng, Bsz, D, N, H, W) for _ in range(3)]
    shared = {k: cs[0][k] for k in ("A", "Cs", "Ds", "delta_bias")}
    assert fusion_ops.cross_ss2d_x3_supported(N, H, W)
    lv = [{k: t(c[k]).requires_grad_(True) for k in ("x", "delta", "Bs")} for c in cs]
    sh = {k: t(v).requires_grad_(True) for k, v in shared.items()}
    before = _lib.launch_count()
    ys = fusion_ops.cross_ss2d_x3([v["x"] for v in lv], [v["delta"] for v in lv], [v["Bs"] for v in lv], sh["Cs"], sh["A"], sh["Ds"],
                                  sh["delta_bias"])
    sum((y * t(c["dy"])).sum() for y, c in zip(ys, cs)).backward()
    assert _lib.launch_count() - before == 2           # one forward kernel, one backward kernel
    tot = {k: 0.0 for k in ("A", "Cs", "Ds", "delta_bias")}
    for y, c, v in zip(ys, cs, lv):
        args = (c["x"], c["delta"], shared["A"], c["Bs"], shared["Cs"], shared["Ds"], shared["delta_bias"])
        assert rel_err(n(y), oracle.ss2d_fwd(*args, True, "f64")) < TOL32
        gx, gd, gA, gB, gC, gD, gb = oracle.ss2d_bwd(*args, c["dy"], True, "f64")
        assert rel_err(n(v["x"].grad), gx) < TOL32 and rel_err(n(v["delta"].grad), gd) < TOL32 and rel_err(n(v["Bs"].grad), gB) < TOL32
        for k, gr in (("A", gA), ("Cs", gC), ("Ds", gD), ("delta_bias", gb)):
            tot[k] = tot[k] + gr
    for k in tot:
        assert rel_err(n(sh[k].grad), tot[k]) < TOL32, k


@pytest.mark.parametrize("shape", [(2, 6, 16, 49), (1, 5, 4, 12), (2, 130, 16, 49), (1, 3, 1, 64)])
def test_swap_scan_fused_vs_oracle(xf, shape):
    """xfs_swap_scan_fused: SwappingScan + S6 (K = 2) + SwappingMerge in one kernel, forward and the as-written backward
    (models/fusion_vmamba.py:189-241, 812, 831-835), against the oracle's three-step composition."""
    from xfmamba_b200 import _lib, fusion_ops
    Bsz, D, N, L = shape
    rng = np.random.default_rng(abs(hash(shape)) % 2**32)
    f = lambda *s_: rng.standard_normal(s_, dtype=np.float32)
    r = lambda *s_: rng.random(s_, dtype=np.float32)
    c = dict(x=f(Bsz, D, L), x2=f(Bsz, D, L), delta=0.5 * r(Bsz, 2 * D, L), A=-0.5 * r(2 * D, N), Bs=f(Bsz, 2, N, L), Cs=f(Bsz, 2, N, L),
             Ds=f(2 * D), delta_bias=0.5 * r(2 * D), dy=f(Bsz, D, L), dy2=f(Bsz, D, L))
    lv = {k: t(v).requires_grad_(True) for k, v in c.items() if k not in ("dy", "dy2")}
    before = _lib.launch_count()
    y, y2 = fusion_ops.swap_scan_fused(lv["x"], lv["x2"], lv["delta"], lv["A"], lv["Bs"], lv["Cs"], lv["Ds"], lv["delta_bias"])
    ((y * t(c["dy"])).sum() + (y2 * t(c["dy2"])).sum()).backward()
    assert _lib.launch_count() - before == 2
    xs = oracle.swap_scan(c["x"], c["x2"]).reshape(Bsz, 2 * D, L)
    ys = oracle.selective_scan_fwd(xs, c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], True, "f64")
    ry, ry2 = oracle.swap_merge(ys.reshape(Bsz, 2, D, L))
    assert rel_err(n(y), ry) < TOL32 and rel_err(n(y2), ry2) < TOL32
    dout = np.stack([c["dy"], c["dy2"]], axis=1).reshape(Bsz, 2 * D, L)              # SwappingMerge.backward: stack
    du, dd, dA, dB, dC, dD, db = oracle.selective_scan_bwd(xs, c["delta"], c["A"], c["Bs"], c["Cs"], c["Ds"], c["delta_bias"], dout, True, "f64")
    du = du.reshape(Bsz, 2, D, L)                                                    # SwappingScan.backward as written: halves
    for k, gr in (("x", du[:, 0]), ("x2", du[:, 1]), ("delta", dd), ("A", dA), ("Bs", dB), ("Cs", dC), ("Ds", dD), ("delta_bias", db)):
        assert rel_err(n(lv[k].grad), gr) < TOL32, k


def test_native_library_was_used(xf):
    """the driver checks which .so the test process loaded; make the launch counter prove it too"""
    from xfmamba_b200 import _lib
    before = _lib.launch_count()
    xf.cross_scan_fn(torch.randn(1, 1, 4, 4, device=dev()))
    assert _lib.launch_count() == before + 1


# ================================================================================================ LayerNorm2d (consumer)
@pytest.mark.parametrize("shape", [(2, 192, 56, 56), (3, 16, 7, 7), (1, 2048, 7, 7), (2, 5, 3, 11), (1, 96, 1, 33)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm2d_matches_reference_layernorm2d(xf, shape, dtype):
    """reference LayerNorm2d = permute -> F.layer_norm -> permute (models/fusion_vmamba.py:52-57)"""
    from xfmamba_b200.norm import layer_norm_2d
    torch.manual_seed(0)
    B, C, H, W = shape
    x = (torch.randn(shape, device=dev()) * 2 + 0.5).to(dtype).requires_grad_(True)
    w = (torch.rand(C, device=dev()) + 0.5).requires_grad_(True)
    b = torch.randn(C, device=dev()).requires_grad_(True)
    g = torch.randn(shape, device=dev()).to(dtype)
    y = layer_norm_2d(x, w, b, 1e-5)
    y.backward(g)
    xr = x.detach().float().clone().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr.permute(0, 2, 3, 1), (C,), wr, br, 1e-5).permute(0, 3, 1, 2)
    yr.backward(g.float())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(n(y), oracle.layernorm2d(n(x), n(w), n(b), 1e-5)) < tol      # CPU oracle (float64)
    assert rel_err(n(y), n(yr)) < tol
    assert rel_err(n(x.grad), n(xr.grad)) < (1e-4 if dtype == torch.float32 else 2e-2)
    assert rel_err(n(w.grad), n(wr.grad)) < (1e-4 if dtype == torch.float32 else 2e-2)
    assert rel_err(n(b.grad), n(br.grad)) < (1e-4 if dtype == torch.float32 else 2e-2)


# ---- depthwise 3x3 conv + SiLU: the producer of the scan input (reference models/fusion_vmamba.py:405-413,1199-1200) ----
@pytest.mark.parametrize("shape,bias,act", [
    ((2, 5, 7, 7), True, True), ((3, 20, 14, 14), False, True), ((2, 11, 28, 28), True, True), ((2, 3, 56, 56), False, True),
    ((1, 2, 57, 58), True, True), ((1, 1, 128, 128), True, True), ((2, 4, 16, 16), True, False), ((1, 70, 7, 7), False, True),
    ((2, 9, 13, 5), True, True),
])
def test_dwconv3x3_silu_fwd_bwd_vs_oracle(shape, bias, act):
    from xfmamba_b200.conv import dwconv3x3_silu
    rng = np.random.default_rng(sum(shape))
    B, C, H, W = shape
    x = rng.standard_normal(shape).astype(np.float32)
    w = (rng.standard_normal((C, 1, 3, 3)) * 0.4).astype(np.float32)
    b = rng.standard_normal(C).astype(np.float32) if bias else None
    dy = rng.standard_normal(shape).astype(np.float32)
    xt, wt = t(x).requires_grad_(), t(w).requires_grad_()
    bt = t(b).requires_grad_() if bias else None
    y = dwconv3x3_silu(xt, wt, bt, act)
    y.backward(t(dy))
    assert rel_err(n(y), oracle.dwconv3x3_silu(x, w, b, act)) < TOL32
    dx, dw, db = oracle.dwconv3x3_silu_bwd(x, w, b, dy, act)
    assert rel_err(n(xt.grad), dx) < TOL32
    assert rel_err(n(wt.grad), dw) < TOL32
    if bias:
        assert rel_err(n(bt.grad), db) < TOL32


def test_dwconv3x3_silu_bf16_and_torch_agreement():
    import torch.nn.functional as F
    from xfmamba_b200.conv import dwconv3x3_silu
    torch.manual_seed(3)
    x = torch.randn(4, 24, 28, 28, device=dev())
    w = torch.randn(24, 1, 3, 3, device=dev()) * 0.3
    b = torch.randn(24, device=dev())
    ref = oracle.dwconv3x3_silu(n(x.bfloat16()), n(w), n(b))
    got = dwconv3x3_silu(x.bfloat16(), w, b)
    assert got.dtype == torch.bfloat16
    assert rel_err(n(got), ref) < TOL16
    torch.backends.cudnn.allow_tf32 = False
    assert rel_err(n(dwconv3x3_silu(x, w, b)), n(F.silu(F.conv2d(x, w, b, padding=1, groups=24)))) < TOL32


def test_dwconv3x3_rejects_cpu_tensors():
    from xfmamba_b200.conv import dwconv3x3_silu
    with pytest.raises(RuntimeError):
        dwconv3x3_silu(torch.randn(1, 2, 4, 4), torch.randn(2, 1, 3, 3))


# ---- low-rank delta projection (reference F.conv1d(dts_r, dt_projs_weight, groups=K), models/fusion_vmamba.py:1155-1157) ----
@pytest.mark.parametrize("B,K,R,D,L,sliced", [
    (2, 4, 6, 192, 3136, True), (3, 4, 24, 70, 196, True), (2, 4, 48, 33, 49, True), (2, 2, 64, 40, 256, False),
    (1, 4, 1, 5, 7, False), (2, 4, 12, 384, 784, True), (1, 4, 8, 16, 1023, False),
    (3, 4, 8, 20, 49, True), (5, 2, 16, 9, 100, False),      # short rows: several images per CTA, last group partial
])
def test_dt_proj_fwd_bwd_vs_oracle(B, K, R, D, L, sliced):
    from xfmamba_b200.proj import dt_proj
    rng = np.random.default_rng(B * 1000 + R + L)
    N = 1
    full = rng.standard_normal((B, K, R + 2 * N, L)).astype(np.float32)
    w = (rng.standard_normal((K, D, R)) * R ** -0.5).astype(np.float32)
    g = rng.standard_normal((B, K * D, L)).astype(np.float32)
    ft = t(full).requires_grad_()
    zt = torch.split(ft, [R, N, N], dim=2)[0] if sliced else ft[:, :, :R].contiguous()   # the model passes the split slice as is
    wt = t(w).requires_grad_()
    out = dt_proj(zt, wt)
    assert out.shape == (B, K * D, L)
    z = full[:, :, :R]
    assert rel_err(n(out), oracle.dt_proj(z, w)) < TOL32
    out.backward(t(g))
    dz, dw = oracle.dt_proj_bwd(z, w, g)
    assert rel_err(n(ft.grad)[:, :, :R], dz) < TOL32
    assert np.all(n(ft.grad)[:, :, R:] == 0)
    assert rel_err(n(wt.grad), dw) < TOL32


@pytest.mark.parametrize("which", ["z", "w"])
def test_dt_proj_bwd_single_gradient(which):
    """only dz or only dW requested: the other output pointer of xfs_dt_proj_bwd is NULL"""
    from xfmamba_b200.proj import dt_proj
    rng = np.random.default_rng(11)
    B, K, R, D, L = 3, 4, 12, 100, 196
    z = rng.standard_normal((B, K, R, L)).astype(np.float32)
    w = (rng.standard_normal((K, D, R)) * R ** -0.5).astype(np.float32)
    g = rng.standard_normal((B, K * D, L)).astype(np.float32)
    zt, wt = t(z).requires_grad_(which == "z"), t(w).requires_grad_(which == "w")
    dt_proj(zt, wt).backward(t(g))
    dz, dw = oracle.dt_proj_bwd(z, w, g)
    if which == "z":
        assert wt.grad is None and rel_err(n(zt.grad), dz) < TOL32
    else:
        assert zt.grad is None and rel_err(n(wt.grad), dw) < TOL32


def test_dt_proj_bf16():
    from xfmamba_b200.proj import dt_proj
    torch.manual_seed(2)
    z = torch.randn(2, 4, 12, 784, device=dev()).bfloat16()
    w = torch.randn(4, 50, 12, device=dev()) * 12 ** -0.5
    out = dt_proj(z, w)
    assert out.dtype == torch.bfloat16
    assert rel_err(n(out), oracle.dt_proj(n(z), n(w))) < TOL16


@pytest.mark.parametrize("shape", [(2, 6, 1, 28, 28), (2, 5, 1, 14, 14), (1, 8, 16, 7, 7)])
def test_ss2d_bwd_accumulator_replicas_agree(xf, shape):
    """xfs_ss2d_bwd_args.acc_replicas: dBs/dCs spread over R copies (channel d -> copy d % R) must sum to the plain result"""
    from xfmamba_b200 import fusion_ops
    Bsz, D, N, H, W = shape
    rng = np.random.default_rng(11)
    c = _rand_ss2d(rng, Bsz, D, N, H, W, False)
    args = [t(c[k]) for k in ("x", "delta", "A", "Bs", "Cs", "Ds", "delta_bias")]
    dy = t(rng.standard_normal((Bsz, D, H * W)).astype(np.float32))
    y, st = fusion_ops.ss2d_fwd_raw(*args, True, torch.float32, True)
    def run(R):
        shp = tuple(args[3].shape) if R == 1 else (R,) + tuple(args[3].shape)
        out = (torch.empty_like(args[0]), torch.empty_like(args[1]), torch.empty_like(args[2]),
               torch.empty(shp, device=dev()), torch.empty(shp, device=dev()), torch.empty_like(args[5]), torch.empty_like(args[6]))
        return [n(v) for v in fusion_ops.ss2d_bwd_raw(*args, dy, st, True, out)]
    base = run(1)
    for R in (3, 8):
        got = run(R)
        assert got[3].shape == base[3].shape
        for a, b in zip(got, base):
            assert rel_err(a, b) < 1e-5


# ---- the reference's own CUDA kernel on the same box (oracle/_ref, built from /root/reference by oracle/build_ref.py) ----
@pytest.mark.parametrize("B,D,H,W", [(4, 48, 56, 56), (2, 24, 28, 28), (3, 16, 14, 14), (2, 8, 20, 19)])
def test_ss2d_vs_reference_cuda_core(B, D, H, W):
    """fused SS2D core fwd + bwd against selective_scan_cuda_core.fwd / .bwd (models/csms6s.py:83, 101) between torch
    CrossScan / CrossMerge (models/csm_triton.py:22-30, 56-62): both fp32 on the GPU, fast-math on both sides"""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/selective_scan_cuda_core.so not built")
    from xfmamba_b200 import ss2d_scan
    m = ref_gpu.load()
    torch.manual_seed(B * 100 + H)
    L, KD = H * W, 4 * D
    x = torch.randn(B, D, H, W, device=dev())
    delta = 0.5 * torch.rand(B, KD, L, device=dev())
    A = -0.5 * torch.rand(KD, 1, device=dev())
    Bs, Cs = torch.randn(B, 4, 1, L, device=dev()), torch.randn(B, 4, 1, L, device=dev())
    Ds, bias = torch.randn(KD, device=dev()), 0.5 * torch.rand(KD, device=dev())
    dy = torch.randn(B, D, L, device=dev())
    y_ref, saved = ref_gpu.ss2d_fwd(m, x, delta, A, Bs, Cs, Ds, bias)
    g_ref = ref_gpu.ss2d_bwd(m, saved, dy, x.shape, delta, A, Bs, Cs, Ds, bias)
    leaves = [t_.clone().requires_grad_() for t_ in (x, delta, A, Bs, Cs, Ds, bias)]
    y = ss2d_scan(*leaves, True, True)
    g = torch.autograd.grad(y, leaves, dy)
    assert rel_err(n(y), n(y_ref)) < TOL32
    for name, a, b in zip(("dx", "ddelta", "dA", "dBs", "dCs", "dDs", "ddelta_bias"), g, g_ref):
        assert rel_err(n(a).reshape(-1), n(b).reshape(-1)) < TOL32, name


@pytest.mark.parametrize("B,K,C,N,L,dtype,softplus", [
    (2, 2, 96, 16, 49, "f32", True),        # the 7x7 fusion blocks: K = 2 (swap scan), N = 16
    (2, 4, 24, 1, 3136, "f32", True),       # stage-1 rows, N = 1 (the d_state = 1 kernels)
    (1, 4, 8, 4, 1500, "f32", False),       # no softplus, L crossing chunk boundaries, N = 4
    (2, 4, 16, 1, 784, "bf16", True),
    (2, 2, 32, 16, 49, "f16", True),
])
def test_selective_scan_fn_vs_reference_cuda_core(B, K, C, N, L, dtype, softplus):
    """stand-alone selective_scan_fn fwd + all 7 gradients against the reference's kernel (models/csms6s.py:83, 101)"""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/selective_scan_cuda_core.so not built")
    from xfmamba_b200 import csms6s
    m = ref_gpu.load()
    torch.manual_seed(L + N)
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    tol = TOL32 if dtype == "f32" else TOL16
    KD = K * C
    u = torch.randn(B, KD, L, device=dev()).to(tdt)
    delta = (0.5 * torch.rand(B, KD, L, device=dev())).to(tdt)
    A = -0.5 * torch.rand(KD, N, device=dev())
    Bs, Cs = torch.randn(B, K, N, L, device=dev()).to(tdt), torch.randn(B, K, N, L, device=dev()).to(tdt)
    Ds, bias = torch.randn(KD, device=dev()), 0.5 * torch.rand(KD, device=dev())
    dy = torch.randn(B, KD, L, device=dev()).to(tdt)
    out, st, *_ = m.fwd(u, delta, A, Bs, Cs, Ds, bias, softplus, 1)
    g_ref = m.bwd(u, delta, A, Bs, Cs, Ds, bias, dy, st, softplus, 1)[:7]
    leaves = [t_.clone().requires_grad_() for t_ in (u, delta, A, Bs, Cs, Ds, bias)]
    y = csms6s.selective_scan_fn(*leaves, softplus, False)          # oflex=False: output in the input dtype, as the core backend
    assert y.dtype == out.dtype
    g = torch.autograd.grad(y, leaves, dy)
    assert rel_err(n(y.float()), n(out.float())) < tol
    for name, a, b in zip(("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"), g, g_ref):
        assert rel_err(n(a.float()).reshape(-1), n(b.float()).reshape(-1)) < tol, name

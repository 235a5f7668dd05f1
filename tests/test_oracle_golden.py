"""Pins the CPU oracle (oracle/) against fixtures produced by the reference's own Python.

The fixtures come from tests/golden/make_golden.py (reference imported unmodified).  Index routes must be
bit exact; fp32 scan outputs must agree to ~1e-6 (both are fp32 evaluations of the same formula, differing only
in libm vs ATen exp/log1p rounding and einsum association); gradients are checked with the f64 oracle against
the reference's fp32 autograd.
"""
import numpy as np
import pytest

import oracle
from conftest import bf16_bits_to_f32, f16_bits_to_f32, rel_err


# ------------------------------------------------------------------------------------------------ routes
def test_route_known_answer():
    # H=2, W=3
    assert oracle.route_table(2, 3).tolist() == [[0, 1, 2, 3, 4, 5], [0, 3, 1, 4, 2, 5], [5, 4, 3, 2, 1, 0],
                                                 [5, 2, 4, 1, 3, 0]]
    lib = oracle.lib()
    for scans in (0, 1, 2):
        r = oracle.route_table(5, 7, scans)
        for k in range(4):
            assert [lib.xfo_route_index(k, l, 5, 7, scans) for l in range(35)] == r[k].tolist()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("scans", [0, 1, 2])
def test_cross_scan_merge_golden(golden, tag, scans):
    g = golden("csm")
    x, ys = g[f"{tag}_x"], g[f"{tag}_ys"]
    B, C, H, W = x.shape
    xs = oracle.cross_scan(x, scans)
    assert np.array_equal(xs, g[f"{tag}_s{scans}_xs"])
    assert np.array_equal(xs, oracle.np_cross_scan(x, scans))
    y = oracle.cross_merge(ys.reshape(B, 4, C, H * W), H, W, scans)
    assert np.array_equal(y, g[f"{tag}_s{scans}_y"]), "merge must follow the reference's add order bit for bit"
    # backward of scan = merge of the grad; backward of merge = scan of the grad (models/csm_triton.py:208-273)
    dx = oracle.cross_merge(g[f"{tag}_gx"], H, W, scans).reshape(B, C, H, W)
    assert np.array_equal(dx, g[f"{tag}_s{scans}_dx"])
    dys = oracle.cross_scan(g[f"{tag}_gy"].reshape(B, C, H, W), scans).reshape(B, 4, C, H, W)
    assert np.array_equal(dys, g[f"{tag}_s{scans}_dys"])
    # one by one
    # (for scans=1 the reference's cross_scan1b1_fwd flattens dims (2,3) of a 5-D tensor, models/csm_triton.py:100,
    #  so its result has shape (B,4,C*H,W): same memory, odd shape -> compare flattened)
    assert np.array_equal(oracle.cross_scan(ys, scans, one_by_one=True).reshape(-1), g[f"{tag}_s{scans}_xs1b1"].reshape(-1))
    assert np.array_equal(oracle.cross_merge_1b1(ys.reshape(B, 4, C, H * W), H, W, scans), g[f"{tag}_s{scans}_y1b1"].reshape(B, 4, C, H * W))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cross_scan_bf16_bit_exact(golden, tag):
    g = golden("csm")
    x = g[f"{tag}_x"]
    xb = (oracle.bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)
    assert np.array_equal(oracle.cross_scan(xb, 0), g[f"{tag}_bf16_xs"])
    # merge in bf16: every add is rounded to bf16 (ATen computes a+b in fp32 and rounds the result)
    ys = oracle.bf16_round(g[f"{tag}_ys"])
    B, _, C, H, W = ys.shape
    L = H * W
    r = oracle.route_table(H, W, 0)
    ysf = ys.reshape(B, 4, C, L)
    t0 = oracle.bf16_round(ysf[:, 0] + ysf[:, 2][..., ::-1])
    t1 = oracle.bf16_round(ysf[:, 1] + ysf[:, 3][..., ::-1])
    t1n = np.empty_like(t1)
    t1n[..., r[1]] = t1
    y = oracle.bf16_round(t0 + t1n)
    assert np.array_equal(y, bf16_bits_to_f32(g[f"{tag}_bf16_y"]))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_swap_golden(golden, tag):
    g = golden("swap")
    xs = oracle.swap_scan(g[f"{tag}_x"], g[f"{tag}_x2"])
    assert np.array_equal(xs, g[f"{tag}_xs"])
    y, y2 = oracle.swap_merge(g[f"{tag}_ys"])
    assert np.array_equal(y, g[f"{tag}_y"]) and np.array_equal(y2, g[f"{tag}_y2"])
    # backward AS WRITTEN in the reference: plain split / stack, no un-swap (models/fusion_vmamba.py:217-241)
    B, C, H, W = g[f"{tag}_x"].shape
    assert np.array_equal(g[f"{tag}_gxs"][:, 0].reshape(B, C, H, W), g[f"{tag}_dx"])
    assert np.array_equal(g[f"{tag}_gxs"][:, 1].reshape(B, C, H, W), g[f"{tag}_dx2"])
    assert np.array_equal(np.stack([g[f"{tag}_gy"], g[f"{tag}_gy2"]], 1), g[f"{tag}_dys"])


# ------------------------------------------------------------------------------------------------ scan
def _case(g, name):
    meta = g[f"{name}_meta"]
    Bsz, K, Cd, N, L, has_D, has_bias, softplus, oflex, dt = [int(v) for v in meta]
    conv = {0: lambda a: a, 1: bf16_bits_to_f32, 2: f16_bits_to_f32}[dt]
    t = {k: conv(g[f"{name}_{k}"]) for k in ("u", "delta", "B", "C")}
    t["A"] = g[f"{name}_A"]
    t["D"] = g[f"{name}_D"] if has_D else None
    t["delta_bias"] = g[f"{name}_delta_bias"] if has_bias else None
    return t, dict(softplus=bool(softplus), oflex=bool(oflex), dt=dt, conv=conv)


SCAN_CASES = ["s1", "s2", "s3", "s4", "s5", "s6", "h1", "h2", "h3"]


@pytest.mark.parametrize("name", SCAN_CASES)
def test_selective_scan_fwd_golden(golden, name):
    g = golden("scan")
    t, m = _case(g, name)
    out32 = oracle.selective_scan_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"], m["softplus"], "f32")
    out64 = oracle.selective_scan_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"], m["softplus"], "f64")
    ref = g[f"{name}_out"]
    if m["dt"] != 0 and not m["oflex"]:
        ref = m["conv"](ref)                       # output cast to the 16-bit input dtype (models/csms6s.py:68)
        assert rel_err(oracle.bf16_round(out32), ref) < 8e-3
        return
    if m["dt"] != 0:
        # 16-bit inputs: the reference applies bias + softplus in the 16-bit dtype before .float() (:47-52);
        # the oracle follows the fp32 contract the native kernels implement -> agreement to 16-bit rounding of delta
        assert rel_err(out32, ref) < (2e-2 if m["dt"] == 1 else 3e-3)
        return
    assert rel_err(out32, ref) < 2e-6, "fp32 oracle vs reference fp32"
    assert rel_err(out64, ref) < 2e-5, "f64 truth vs reference fp32"


@pytest.mark.parametrize("name", ["s1", "s2", "s3", "s4", "s5", "s6"])
def test_selective_scan_bwd_golden(golden, name):
    g = golden("scan")
    t, m = _case(g, name)
    grads = oracle.selective_scan_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"],
                                      g[f"{name}_dout"], m["softplus"], "f64")
    for key, got in zip(["u", "delta", "A", "B", "C", "D", "delta_bias"], grads):
        if got is None:
            assert f"{name}_d{key}" not in g
            continue
        assert rel_err(got, g[f"{name}_d{key}"]) < 3e-5, key
    grads32 = oracle.selective_scan_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["delta_bias"],
                                        g[f"{name}_dout"], m["softplus"], "f32")
    for key, got in zip(["u", "delta", "A", "B", "C", "D", "delta_bias"], grads32):
        if got is not None:
            assert rel_err(got, g[f"{name}_d{key}"]) < 1e-4, key


def test_ss2d_core_golden(golden):
    """composition cross_scan -> scan -> cross_merge against SS2Dv2.forward_corev2's own intermediates"""
    g = golden("cores")
    x = g["ss2d_x"]
    B, D, H, W = x.shape
    assert np.array_equal(oracle.cross_scan(x, 0).reshape(B, 4 * D, H * W), g["ss2d_us"])
    y = oracle.ss2d_fwd(x, g["ss2d_dts"], g["ss2d_As"], g["ss2d_Bs"], g["ss2d_Cs"], g["ss2d_Ds"], g["ss2d_delta_bias"])
    assert rel_err(y, g["ss2d_ymerged"]) < 2e-6
    ys = oracle.selective_scan_fwd(g["ss2d_us"], g["ss2d_dts"], g["ss2d_As"], g["ss2d_Bs"], g["ss2d_Cs"], g["ss2d_Ds"],
                                   g["ss2d_delta_bias"])
    assert rel_err(ys.reshape(g["ss2d_ys"].shape), g["ss2d_ys"]) < 2e-6

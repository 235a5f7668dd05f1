"""CPU: pins oracle.dwconv3x3_silu / _bwd to torch's own depthwise convolution + SiLU (the operators the reference calls,
models/fusion_vmamba.py:405-413,1199-1200) including autograd gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from conftest import rel_err


@pytest.mark.parametrize("shape,bias", [((2, 3, 7, 7), True), ((1, 4, 14, 13), False), ((2, 2, 5, 9), True)])
def test_oracle_dwconv_matches_torch(shape, bias):
    torch.manual_seed(sum(shape))
    C = shape[1]
    x = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    w = torch.randn(C, 1, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.randn(C, dtype=torch.float64, requires_grad=True) if bias else None
    dy = torch.randn(shape, dtype=torch.float64)
    y = F.silu(F.conv2d(x, w, b, padding=1, groups=C))
    y.backward(dy)
    bn = None if b is None else b.detach().numpy()
    assert rel_err(oracle.dwconv3x3_silu(x.detach().numpy(), w.detach().numpy(), bn), y.detach().numpy()) < 1e-12
    dx, dw, db = oracle.dwconv3x3_silu_bwd(x.detach().numpy(), w.detach().numpy(), bn, dy.numpy())
    assert rel_err(dx, x.grad.numpy()) < 1e-12
    assert rel_err(dw, w.grad.numpy()) < 1e-12
    if bias:
        assert rel_err(db, b.grad.numpy()) < 1e-12


def test_oracle_dt_proj_matches_torch_grouped_conv1d():
    """the reference's formulation: F.conv1d(dts_r.view(B, K*R, L), W.view(K*D, R, 1), groups=K) (models/fusion_vmamba.py:1155-1157)"""
    torch.manual_seed(5)
    B, K, R, D, L = 2, 4, 3, 10, 21
    z = torch.randn(B, K, R, L, dtype=torch.float64, requires_grad=True)
    w = torch.randn(K, D, R, dtype=torch.float64, requires_grad=True)
    g = torch.randn(B, K * D, L, dtype=torch.float64)
    out = F.conv1d(z.view(B, K * R, L), w.view(K * D, R, 1), groups=K)
    out.backward(g)
    assert rel_err(oracle.dt_proj(z.detach().numpy(), w.detach().numpy()), out.detach().numpy()) < 1e-12
    dz, dw = oracle.dt_proj_bwd(z.detach().numpy(), w.detach().numpy(), g.numpy())
    assert rel_err(dz, z.grad.numpy()) < 1e-12
    assert rel_err(dw, w.grad.numpy()) < 1e-12

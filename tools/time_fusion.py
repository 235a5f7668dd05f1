"""time the fusion-block scan kernels at their XFMamba-B shapes (GPU box): three deep-fusion streams and the shallow swap scan"""
import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200 import fusion_ops, _lib
dev = torch.device("cuda:0")
B, D, H, W, N = 32, 2048, 7, 7, 16
L = H * W
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda *s: torch.randn(*s, device=dev, generator=g)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / iters * 1e3
xs = [mk(B, D, H, W) for _ in range(3)]; ds = [0.5 * torch.rand(B, 4 * D, L, device=dev, generator=g) for _ in range(3)]
Bs = [mk(B, 4, N, L) for _ in range(3)]; Cs = mk(B, 4, N, L)
A = -0.5 * torch.rand(4 * D, N, device=dev, generator=g); Ds = mk(4 * D); bias = 0.5 * torch.rand(4 * D, device=dev, generator=g)
dys = [mk(B, D, L) for _ in range(3)]
# (a) three separate fused SS2D launches (round-1 path)
def three_fwd():
    return [fusion_ops.ss2d_fwd_raw(xs[i], ds[i], A, Bs[i], Cs, Ds, bias, True, torch.float32, True) for i in range(3)]
st = three_fwd()
def three_bwd():
    for i in range(3): fusion_ops.ss2d_bwd_raw(xs[i], ds[i], A, Bs[i], Cs, Ds, bias, dys[i], st[i][1], True)
t3f, t3b = timeit(three_fwd), timeit(three_bwd)
# (b) one x3 launch
lv = [t.clone().requires_grad_(True) for t in xs]
def x3_fwd():
    with torch.no_grad(): return fusion_ops.cross_ss2d_x3(xs, ds, Bs, Cs, A, Ds, bias)
tx3f = timeit(x3_fwd)
def x3_fb():
    ys = fusion_ops.cross_ss2d_x3(lv, ds, Bs, Cs, A, Ds, bias)
    torch.autograd.backward(ys, dys)
tx3fb = timeit(x3_fb)
elems = 3 * B * 4 * D * L
mufu_us = lambda per: elems * per / (148 * 16 * 1.965e9) * 1e6
print(f"deep fusion (3 streams, B={B} D={D} 7x7 N={N}): 3 launches fwd {t3f:.1f} us bwd(+memsets) {t3b:.1f} us | x3 fwd {tx3f:.1f} us, fwd+bwd(autograd, +allocs) {tx3fb:.1f} us"
      f" | MUFU floor fwd {mufu_us(N + 2):.1f} us ({mufu_us(N + 2) / tx3f:.2f} of it)")
# shallow: K = 2
x, x2 = mk(B, D, L), mk(B, D, L); d2 = 0.5 * torch.rand(B, 2 * D, L, device=dev, generator=g)
B2, C2 = mk(B, 2, N, L), mk(B, 2, N, L); A2 = -0.5 * torch.rand(2 * D, N, device=dev, generator=g); D2 = mk(2 * D); b2 = 0.5 * torch.rand(2 * D, device=dev, generator=g)
from xfmamba_b200 import selective_scan_fn
def comp():
    with torch.no_grad():
        xs_ = fusion_ops.SwappingScan_multiview.apply(x.view(B, D, H, W), x2.view(B, D, H, W))
        ys_ = selective_scan_fn(xs_.view(B, -1, L), d2, A2, B2, C2, D2, b2, True, True)
        return fusion_ops.SwappingMerge_multiview.apply(ys_.view(B, 2, D, L))
def fused():
    with torch.no_grad(): return fusion_ops.swap_scan_fused(x, x2, d2, A2, B2, C2, D2, b2)
print(f"shallow fusion (B={B} D={D} L={L} N={N}): swap + scan + split {timeit(comp):.1f} us | fused {timeit(fused):.1f} us")

"""stand-alone operators: time and algorithmic GB/s at the config-2 shape (GPU box)"""
import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200 import csm, csms6s, fusion_ops
dev = torch.device("cuda:0")
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
def line(name, ms, nbytes):
    print(f"{name:46s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.0f} GB/s")
B, D, H, W, N, K = 64, 192, 56, 56, 1, 4
L, KD, s = H * W, K * D, 4
x = torch.randn(B, D, H, W, device=dev)
ys = torch.randn(B, K, D, H, W, device=dev)
line("cross_scan  (B,192,56,56) -> (B,4,192,L)", timeit(lambda: csm.cross_scan_fn(x)), B * D * L * 5 * s)
line("cross_merge (B,4,192,56,56) -> (B,192,L)", timeit(lambda: csm.cross_merge_fn(ys)), B * D * L * 5 * s)
u = torch.randn(B, KD, L, device=dev); delta = 0.5 * torch.rand(B, KD, L, device=dev); A = -0.5 * torch.rand(KD, N, device=dev)
Bs = torch.randn(B, K, N, L, device=dev); Cs = torch.randn(B, K, N, L, device=dev); Ds = torch.randn(KD, device=dev); bias = 0.5 * torch.rand(KD, device=dev)
line("selective_scan_fn fwd (B,768,3136) N=1", timeit(lambda: csms6s.selective_scan_fn(u, delta, A, Bs, Cs, Ds, bias, True, True)), B * L * (KD * 3 * s + 2 * K * N * s))
ur, dr, Ar, Br, Cr, Dr, br = [t.clone().requires_grad_() for t in (u, delta, A, Bs, Cs, Ds, bias)]
dy = torch.randn(B, KD, L, device=dev)
def fb():
    y = csms6s.selective_scan_fn(ur, dr, Ar, Br, Cr, Dr, br, True, True)
    torch.autograd.grad(y, (ur, dr, Ar, Br, Cr, Dr, br), dy)
tf = timeit(lambda: csms6s.selective_scan_fn(ur, dr, Ar, Br, Cr, Dr, br, True, True))
tfb = timeit(fb)
line("selective_scan_fn bwd (autograd, incl. memsets)", tfb - tf, B * L * (KD * 3 * s + KD * 2 * s + 4 * K * N * s))
x1 = torch.randn(32, 1536, 7, 7, device=dev); x2 = torch.randn(32, 1536, 7, 7, device=dev)
line("swapping_scan (32,1536,7,7) x2 -> (32,2,1536,49)", timeit(lambda: fusion_ops.SwappingScan_multiview.apply(x1, x2)), 32 * 1536 * 49 * 4 * s)
u16 = torch.randn(32, 3072, 49, device=dev); d16 = 0.5 * torch.rand(32, 3072, 49, device=dev); A16 = -0.5 * torch.rand(3072, 16, device=dev)
B16 = torch.randn(32, 2, 16, 49, device=dev); C16 = torch.randn(32, 2, 16, 49, device=dev); D16 = torch.randn(3072, device=dev); b16 = 0.5 * torch.rand(3072, device=dev)
line("selective_scan_fn fwd (32,3072,49) N=16 K=2", timeit(lambda: csms6s.selective_scan_fn(u16, d16, A16, B16, C16, D16, b16, True, True)), 32 * 49 * (3072 * 3 * s + 2 * 2 * 16 * s))
# ---- producer / consumer kernels around the core
from xfmamba_b200.conv import dwconv3x3_silu
from xfmamba_b200.norm import layer_norm_2d
from xfmamba_b200.proj import dt_proj
for (Bn, Cn, Hn) in ((128, 192, 56), (128, 384, 28), (128, 768, 14), (128, 1536, 7)):
    xx = torch.randn(Bn, Cn, Hn, Hn, device=dev, requires_grad=True); ww = torch.randn(Cn, 1, 3, 3, device=dev, requires_grad=True)
    bb = torch.randn(Cn, device=dev, requires_grad=True); gg = torch.randn(Bn, Cn, Hn, Hn, device=dev)
    n = xx.numel()
    tfw = timeit(lambda: dwconv3x3_silu(xx.detach(), ww.detach(), bb.detach()))
    def fbc():
        yy = dwconv3x3_silu(xx, ww, bb); torch.autograd.grad(yy, (xx, ww, bb), gg)
    line(f"dwconv3x3+SiLU fwd ({Bn},{Cn},{Hn},{Hn})", tfw, n * 2 * s)
    line(f"dwconv3x3+SiLU bwd ({Bn},{Cn},{Hn},{Hn})", timeit(fbc) - tfw, n * 3 * s)
    lw = torch.randn(Cn, device=dev, requires_grad=True); lb = torch.randn(Cn, device=dev, requires_grad=True)
    tlf = timeit(lambda: layer_norm_2d(xx.detach(), lw.detach(), lb.detach()))
    def fbl():
        yy = layer_norm_2d(xx, lw, lb); torch.autograd.grad(yy, (xx, lw, lb), gg)
    line(f"LayerNorm2d fwd ({Bn},{Cn},{Hn},{Hn})", tlf, n * 2 * s)
    line(f"LayerNorm2d bwd ({Bn},{Cn},{Hn},{Hn})", timeit(fbl) - tlf, n * 3 * s)
    Rn = Cn // 32 if Cn >= 192 else 6
    zz = torch.randn(Bn, 4, Rn + 2, Hn * Hn, device=dev)[:, :, :Rn]; wd = torch.randn(4, Cn, Rn, device=dev)
    line(f"dt_proj fwd (B={Bn}, R={Rn}, D={Cn}, L={Hn*Hn})", timeit(lambda: dt_proj(zz, wd)), Bn * 4 * Cn * Hn * Hn * s)

"""time ss2d_scan (inference, whichever path the library takes) at the 512^2-input stage-1 / stage-2 shapes of XFMamba-B (GPU box)"""
import sys, torch
sys.path.insert(0, ".")
import xfmamba_b200 as xf
from xfmamba_b200 import _lib
dev = torch.device("cuda:0")
def run(B, D, H, W):
    L = H * W
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, D, H, W, device=dev, generator=g)
    delta = 0.5 * torch.rand(B, 4 * D, L, device=dev, generator=g)
    A = -0.5 * torch.rand(4 * D, 1, device=dev, generator=g)
    Bs = torch.randn(B, 4, 1, L, device=dev, generator=g); Cs = torch.randn(B, 4, 1, L, device=dev, generator=g)
    Ds = torch.randn(4 * D, device=dev, generator=g); bias = 0.5 * torch.rand(4 * D, device=dev, generator=g)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(3): xf.ss2d_scan(x, delta, A, Bs, Cs, Ds, bias)
        n0 = _lib.launch_count()
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(10): y = xf.ss2d_scan(x, delta, A, Bs, Cs, Ds, bias)
        e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    fb = B * L * (D * 4 + 4 * D * 4 + 8 * 4 + D * 4)
    print(f"B={B} D={D} {H}x{W}: ss2d_scan forward {t*1e3:8.1f} us  {fb/t/1e6:7.0f} GB/s algorithmic (24 B per (b,d,l)), {(_lib.launch_count()-n0)//10} launches")
run(16, 256, 128, 128); run(16, 512, 64, 64); run(16, 1024, 32, 32)

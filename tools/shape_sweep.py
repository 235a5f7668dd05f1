"""fused SS2D kernel time / algorithmic GB/s per XFMamba stage shape (GPU box)"""
import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200 import fusion_ops, _lib
dev = torch.device("cuda:0")
def run(B, D, H, W, N, dtype=torch.float32, iters=10):
    L = H * W
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, D, H, W, device=dev, generator=g).to(dtype)
    delta = (0.5 * torch.rand(B, 4 * D, L, device=dev, generator=g)).to(dtype)
    A = -0.5 * torch.rand(4 * D, N, device=dev, generator=g)
    Bs = torch.randn(B, 4, N, L, device=dev, generator=g).to(dtype); Cs = torch.randn(B, 4, N, L, device=dev, generator=g).to(dtype)
    Ds = torch.randn(4 * D, device=dev, generator=g); bias = 0.5 * torch.rand(4 * D, device=dev, generator=g)
    dy = torch.randn(B, D, L, device=dev, generator=g)
    s = 4 if dtype == torch.float32 else 2
    fb = B * L * (D * s + 4 * D * s + 8 * N * s + D * 4); bb = B * L * (2 * D * s + 8 * D * s + 16 * N * s + D * 4)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf = tb = 0.0
    for it in range(iters + 3):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(); y, st = fusion_ops.ss2d_fwd_raw(x, delta, A, Bs, Cs, Ds, bias, True, torch.float32, True)
        e1.record(); out = fusion_ops.ss2d_bwd_raw(x, delta, A, Bs, Cs, Ds, bias, dy, st, True)
        e2.record(); torch.cuda.synchronize()
        if it >= 3: tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    tf /= iters; tb /= iters
    print(f"B={B:3d} D={D:5d} {H:3d}x{W:<3d} N={N:2d} {str(dtype)[6:]:8s} fwd {tf*1e3:8.1f} us {fb/tf/1e6:7.0f} GB/s | bwd(+memsets) {tb*1e3:8.1f} us {bb/tb/1e6:7.0f} GB/s | elements {B*4*D*L/1e6:7.1f} M  fwd {tf*1e6/(B*4*D*L)*1e3:.2f} ps/el bwd {tb*1e6/(B*4*D*L)*1e3:.2f}")
for shp in [(1, 192, 56, 56, 1), (8, 192, 56, 56, 1), (64, 192, 56, 56, 1), (128, 192, 56, 56, 1), (64, 96, 56, 56, 1), (64, 256, 56, 56, 1),
            (64, 192, 56, 57, 1), (64, 384, 28, 28, 1), (64, 512, 28, 28, 1), (64, 1024, 14, 14, 1), (64, 2048, 7, 7, 1), (32, 2048, 7, 7, 16), (8, 256, 128, 128, 1), (8, 1024, 32, 32, 1)]:
    try: run(*shp)
    except Exception as e: print(shp, "ERR", e)
run(64, 192, 56, 56, 1, torch.bfloat16)
run(64, 1024, 14, 14, 1, torch.bfloat16)

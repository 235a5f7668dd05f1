"""time the fused SS2D forward / backward launches of one shape (GPU box):  python tools/time_shape.py B D H W N [bf16]"""
import os, sys, torch
sys.path.insert(0, ".")
from xfmamba_b200 import fusion_ops
dev = torch.device("cuda:0")
B, D, H, W, N = [int(v) for v in sys.argv[1:6]]
dtype = torch.bfloat16 if len(sys.argv) > 6 and sys.argv[6] == "bf16" else torch.float32
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
iters = 20
for it in range(iters + 5):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); y, st = fusion_ops.ss2d_fwd_raw(x, delta, A, Bs, Cs, Ds, bias, True, torch.float32, True)
    e1.record(); out = fusion_ops.ss2d_bwd_raw(x, delta, A, Bs, Cs, Ds, bias, dy, st, True)
    e2.record(); torch.cuda.synchronize()
    if it >= 5: tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
tf /= iters; tb /= iters
print(f"cfg fwd={os.environ.get('XFS_RING_FWD','-')} bwd={os.environ.get('XFS_RING_BWD','-')} B={B} D={D} {H}x{W} N={N} {str(dtype)[6:]} fwd {tf*1e3:8.1f} us {fb/tf/1e6:7.0f} GB/s | bwd(+memsets) {tb*1e3:8.1f} us {bb/tb/1e6:7.0f} GB/s")

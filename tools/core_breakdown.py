"""kernel-time breakdown of one ss2d_core call (the part of an SS2D block between the depthwise conv and out_norm) at a
given stage shape, plus candidate formulations of the dt projection (GPU box)
usage: python tools/core_breakdown.py B D H W R"""
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from xfmamba_b200 import model as M
from torch.profiler import profile, ProfilerActivity
B, D, H, W, R = [int(v) for v in sys.argv[1:6]]
N, K, L = 1, 4, H * W
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, D, H, W, device=dev)
xw = torch.randn(K, R + 2 * N, D, device=dev) * D ** -0.5
dw = torch.randn(K, D, R, device=dev) * R ** -0.5
db = torch.rand(K, D, device=dev)
Al = torch.zeros(K * D, N, device=dev)
Ds = torch.ones(K * D, device=dev)

def run(fn, tag):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    ev = prof.key_averages(); tot = sum(e.device_time_total for e in ev)
    print(f"-- {tag}: total {tot/1e3:.3f} ms")
    for e in sorted(ev, key=lambda e: -e.device_time_total)[:8]:
        print(f"   {e.device_time_total/1e3:7.3f} ms x{e.count:<3} {e.key[:100]}")

with torch.no_grad():
    run(lambda: M.ss2d_core(x, xw, dw, db, Al, Ds), f"ss2d_core B={B} D={D} {H}x{W} R={R}")
    dts_r = torch.randn(B, K, R, L, device=dev)
    run(lambda: F.conv1d(dts_r.reshape(B, K * R, L), dw.reshape(K * D, R, 1), groups=K), "dt_proj: grouped conv1d")
    run(lambda: torch.matmul(dw.unsqueeze(0), dts_r), "dt_proj: broadcast matmul (K,D,R)@(B,K,R,L)")
    run(lambda: torch.einsum("bkrl,kdr->bkdl", dts_r, dw), "dt_proj: einsum")
    z = torch.randn(B, K * (R + 2 * N), H, W, device=dev)
    run(lambda: F.conv2d(x, xw.reshape(K * (R + 2 * N), D, 1, 1)), "x_proj: 1x1 conv2d")
    run(lambda: torch.matmul(xw.reshape(1, K * (R + 2 * N), D), x.view(B, D, L)), "x_proj: matmul")

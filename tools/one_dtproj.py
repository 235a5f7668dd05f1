"""dt_proj forward at one shape: time and delta GB/s (GPU box).  usage: one_dtproj.py B K R D L [B K R D L ...]"""
import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200.proj import dt_proj
dev = torch.device("cuda:0")
a = [int(v) for v in sys.argv[1:]]
for i in range(0, len(a), 5):
    B, K, R, D, L = a[i:i + 5]
    z = torch.randn(B, K, R + 2, L, device=dev)[:, :, :R]
    w = torch.randn(K, D, R, device=dev)
    for _ in range(3):
        out = dt_proj(z, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = dt_proj(z, w)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = torch.einsum("bkrl,kdr->bkdl", z.double(), w.double()).reshape(B, K * D, L)
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"dt_proj fwd (B={B}, R={R}, D={D}, L={L}) {ms*1e3:8.1f} us {out.numel()*4/ms/1e6:7.0f} GB/s  rel err {err:.2e}")

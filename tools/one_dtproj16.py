"""dt_proj forward with bf16 rows: time and error against float64 (GPU box)"""
import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200.proj import dt_proj
dev = torch.device("cuda:0")
for (B, K, R, D, L) in ((128, 4, 24, 768, 196), (128, 4, 12, 384, 784)):
    z = torch.randn(B, K, R, L, device=dev).bfloat16(); w = torch.randn(K, D, R, device=dev)
    for _ in range(3): out = dt_proj(z, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): out = dt_proj(z, w)
    e1.record(); torch.cuda.synchronize()
    ref = torch.einsum("bkrl,kdr->bkdl", z.double(), w.double()).reshape(B, K * D, L)
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"dt_proj fwd bf16 (B={B}, R={R}, D={D}, L={L}) {e0.elapsed_time(e1)/20*1e3:8.1f} us  rel err {err:.2e} (bf16 output rounding 4e-3)")

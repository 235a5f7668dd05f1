"""dt_proj backward (dz, dW) per stage shape: time of the autograd backward (GPU box).  usage: time_dtproj_bwd.py [base|small]"""
import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200.proj import dt_proj
dev = torch.device("cuda:0")
shapes = {"base": [(64, 8, 256, 3136), (64, 16, 512, 784), (64, 32, 1024, 196), (64, 64, 2048, 49)],
          "small": [(64, 6, 192, 3136), (64, 12, 384, 784), (64, 24, 768, 196), (64, 48, 1536, 49)]}[sys.argv[1] if len(sys.argv) > 1 else "base"]
for B, R, D, L in shapes:
    full = torch.randn(B, 4, R + 2, L, device=dev, requires_grad=True)
    w = torch.randn(4, D, R, device=dev, requires_grad=True)
    g = torch.randn(B, 4 * D, L, device=dev)
    out = dt_proj(full[:, :, :R], w)
    def bwd():
        torch.autograd.grad(out, (full, w), g, retain_graph=True)
    for _ in range(3): bwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): bwd()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"dt_proj bwd (B={B}, R={R}, D={D}, L={L}) {ms*1e3:8.1f} us   g read once = {g.numel()*4/6.55e6:6.1f} us")

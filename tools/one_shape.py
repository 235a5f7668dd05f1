import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200 import fusion_ops
dev = torch.device("cuda:0")
B, D, H, W, N = [int(v) for v in sys.argv[1:6]]
L = H * W
x = torch.randn(B, D, H, W, device=dev); delta = 0.5 * torch.rand(B, 4 * D, L, device=dev); A = -0.5 * torch.rand(4 * D, N, device=dev)
Bs = torch.randn(B, 4, N, L, device=dev); Cs = torch.randn(B, 4, N, L, device=dev); Ds = torch.randn(4 * D, device=dev); bias = 0.5 * torch.rand(4 * D, device=dev)
dy = torch.randn(B, D, L, device=dev)
for _ in range(3):
    y, st = fusion_ops.ss2d_fwd_raw(x, delta, A, Bs, Cs, Ds, bias, True, torch.float32, True)
    fusion_ops.ss2d_bwd_raw(x, delta, A, Bs, Cs, Ds, bias, dy, st, True)
torch.cuda.synchronize()

import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200 import fusion_ops
dev = torch.device("cuda:0")
B, D, H, W, N = 32, 2048, 7, 7, 16
L = H * W
mk = lambda *s: torch.randn(*s, device=dev)
xs = [mk(B, D, H, W) for _ in range(3)]; ds = [0.5 * torch.rand(B, 4 * D, L, device=dev) for _ in range(3)]
Bs = [mk(B, 4, N, L) for _ in range(3)]; Cs = mk(B, 4, N, L)
A = -0.5 * torch.rand(4 * D, N, device=dev); Ds = mk(4 * D); bias = 0.5 * torch.rand(4 * D, device=dev)
lv = [t.clone().requires_grad_(True) for t in xs]
for _ in range(2):
    ys = fusion_ops.cross_ss2d_x3(lv, ds, Bs, Cs, A, Ds, bias)
    torch.autograd.backward(ys, [mk(B, D, L) for _ in range(3)])
torch.cuda.synchronize()

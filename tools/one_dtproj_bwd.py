import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200.proj import dt_proj
dev = torch.device("cuda:0")
B, R, D, L = [int(v) for v in sys.argv[1:5]]
full = torch.randn(B, 4, R + 2, L, device=dev, requires_grad=True)
w = torch.randn(4, D, R, device=dev, requires_grad=True)
g = torch.randn(B, 4 * D, L, device=dev)
out = dt_proj(full[:, :, :R], w)
for _ in range(3):
    torch.autograd.grad(out, (full, w), g, retain_graph=True)
torch.cuda.synchronize()

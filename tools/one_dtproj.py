import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200.proj import dt_proj
B, K, R, D, L = [int(v) for v in sys.argv[1:6]]
dev = torch.device("cuda:0")
z = torch.randn(B, K, R + 2, L, device=dev)[:, :, :R]
w = torch.randn(K, D, R, device=dev)
for _ in range(3):
    out = dt_proj(z, w)
torch.cuda.synchronize()

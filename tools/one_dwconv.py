import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200 import conv
dev = torch.device("cuda:0")
x = torch.randn(128, 192, 56, 56, device=dev, requires_grad=True)
w = torch.randn(192, 1, 3, 3, device=dev, requires_grad=True); b = torch.randn(192, device=dev, requires_grad=True)
for _ in range(3):
    y = conv.dwconv3x3_silu(x, w, b)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()

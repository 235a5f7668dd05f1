"""kernel-time breakdown of one XFMamba step with torch.profiler (GPU box)"""
import sys, torch
sys.path.insert(0, ".")
from xfmamba_b200.model import TwoViewXFMamba
from torch.profiler import profile, ProfilerActivity
variant, batch, train, amp = sys.argv[1], int(sys.argv[2]), sys.argv[3] == "train", sys.argv[4] == "bf16"
dev = torch.device("cuda:0")
m = TwoViewXFMamba(outputs=2, type=variant).to(dev)
xa, xb = torch.randn(batch, 1, 224, 224, device=dev), torch.randn(batch, 1, 224, 224, device=dev)
y = torch.randint(0, 2, (batch,), device=dev)
m.train(train)
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        if train:
            loss = torch.nn.functional.cross_entropy(m(xa, xb).float(), y); loss.backward()
        else:
            with torch.no_grad(): m(xa, xb)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev)
print(f"{variant} batch {batch} {'train' if train else 'infer'} {'bf16' if amp else 'fp32'}: total device time {tot/1e3:.2f} ms")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:45]:
    print(f"  {100*e.device_time_total/tot:5.1f}%  {e.device_time_total/1e3:8.2f} ms  x{e.count:<5} {e.key[:90]}")

"""who launches the strided-copy kernels in one inference step (GPU box)"""
import sys, torch, collections
sys.path.insert(0, ".")
from xfmamba_b200.model import TwoViewXFMamba
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
m = TwoViewXFMamba(outputs=13, type="small").to(dev).eval()
xa, xb = torch.randn(64, 1, 224, 224, device=dev), torch.randn(64, 1, 224, 224, device=dev)
with torch.no_grad():
    for _ in range(2): m(xa, xb)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        m(xa, xb); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.name in ("aten::copy_", "aten::contiguous", "aten::clone") and e.device_time_total > 0 and e.name == "aten::copy_":
        st = [s for s in (e.stack or []) if "xfmamba_b200" in s or "tools/" in s]
        key = (st[0] if st else "?", str(e.input_shapes[:1]))
        agg[key][0] += e.device_time_total; agg[key][1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{v[0]/1e3:7.3f} ms x{v[1]:<3} {k[0][-90:]}  {k[1]}")

"""Where a training step's time goes (CUDA-event timing per phase + per-kernel launch list via torch profiler)."""
import dataclasses, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from instant_nvr_b200.config import PathConfig
from instant_nvr_b200.network import Network
from instant_nvr_b200.renderer import Renderer
from instant_nvr_b200.synthetic import make_rays

cfg = PathConfig.inb_377(N_samples=64).with_(perturb=1.0, use_reg_distortion=True)
frame, _ = bench.build_views(1, 64, 64)
with torch.device("cuda"):
    net = Network(cfg)
net = net.cuda()
bench.device_weights(net, frame)
gframe = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in frame.items()}
rays = make_rays(frame, 32, 32)
batch = {**gframe, **{k: v.cuda() for k, v in rays.items()}}
target = torch.rand(1, 1024, 3, device="cuda")
params = [p for p in net.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=5e-4, eps=1e-15)
r = Renderer(net)
net.train()
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
tot = {"fwd": 0, "bwd": 0, "opt": 0}
for it in range(8):
    opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a = ev()
    ret = r.render(dict(batch))
    loss = ((ret["rgb_map"] - target) ** 2).mean() + 0.1 * ret["reg_distortion_loss"].mean() + 0.1 * torch.norm(ret["resd"], dim=2).mean()
    if ret["oresd"].numel():
        loss = loss + 0.01 * (ret["oresd"] ** 2).mean()
    b = ev(); loss.backward(); c = ev(); opt.step(); d = ev(); torch.cuda.synchronize(); t1 = time.perf_counter()
    if it >= 3:
        tot["fwd"] += a.elapsed_time(b); tot["bwd"] += b.elapsed_time(c); tot["opt"] += c.elapsed_time(d)
    print(it, "fwd %.2f bwd %.2f opt %.2f wall %.2f ms" % (a.elapsed_time(b), b.elapsed_time(c), c.elapsed_time(d), 1e3 * (t1 - t0)), "surv", ret["resd"].shape[1])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    opt.zero_grad(set_to_none=True)
    ret = r.render(dict(batch))
    loss = ((ret["rgb_map"] - target) ** 2).mean() + 0.1 * torch.norm(ret["resd"], dim=2).mean()
    loss.backward(); opt.step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

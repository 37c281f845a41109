"""Where a training step's time goes: CUDA-event time and HOST time per phase of the bench's training step (fused Adam),
then the torch profiler's CPU-side and CUDA-side top lists of one step.  usage: python tools/train_breakdown.py [torch]"""
import dataclasses, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from instant_nvr_b200.config import PathConfig
from instant_nvr_b200.network import Network
from instant_nvr_b200.optimizer import FusedAdam
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
use_torch = len(sys.argv) > 1 and sys.argv[1] == "torch"
opt = torch.optim.Adam(params, lr=5e-4, eps=1e-15) if use_torch else FusedAdam(params, lr=5e-4, eps=1e-15)
r = Renderer(net)
net.train()


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def loss_of(ret):
    loss = ((ret["rgb_map"] - target) ** 2).mean() + 0.1 * ret["reg_distortion_loss"].mean() + 0.1 * torch.norm(ret["resd"], dim=2).mean()
    if ret["oresd"].numel():
        loss = loss + 0.01 * (ret["oresd"] ** 2).mean()
    return loss


N = 20
gpu = {"render": 0.0, "loss": 0.0, "bwd": 0.0, "opt": 0.0}
host = {"zero": 0.0, "render": 0.0, "loss": 0.0, "bwd": 0.0, "opt": 0.0}
wall = 0.0
for it in range(N + 3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    t1 = time.perf_counter(); a = ev()
    ret = r.render(dict(batch))
    t2 = time.perf_counter(); b = ev()
    loss = loss_of(ret)
    t3 = time.perf_counter(); c = ev()
    loss.backward()
    t4 = time.perf_counter(); d = ev()
    opt.step()
    t5 = time.perf_counter(); e = ev()
    torch.cuda.synchronize(); t6 = time.perf_counter()
    if it >= 3:
        for k, (x, y) in {"render": (a, b), "loss": (b, c), "bwd": (c, d), "opt": (d, e)}.items():
            gpu[k] += x.elapsed_time(y) / N
        for k, (x, y) in {"zero": (t0, t1), "render": (t1, t2), "loss": (t2, t3), "bwd": (t3, t4), "opt": (t4, t5)}.items():
            host[k] += 1e3 * (y - x) / N
        wall += 1e3 * (t6 - t0) / N
print("optimizer:", type(opt).__name__, " survivors", ret["resd"].shape[1] // 5)
print("event ms between phase marks :", {k: round(v, 3) for k, v in gpu.items()}, "sum", round(sum(gpu.values()), 3))
print("host ms per phase (no sync)  :", {k: round(v, 3) for k, v in host.items()}, "sum", round(sum(host.values()), 3))
print("wall ms per step (sync both sides): %.3f" % wall)

# back-to-back steps as the bench times them
torch.cuda.synchronize(); e0 = ev()
for _ in range(N):
    opt.zero_grad(set_to_none=True); loss_of(r.render(dict(batch))).backward(); opt.step()
e1 = ev(); torch.cuda.synchronize()
print("back-to-back ms per step: %.3f" % (e0.elapsed_time(e1) / N), " optimizer plan builds so far:", getattr(opt, "plan_builds", None))
for _ in range(2):
    torch.cuda.synchronize(); e0 = ev()
    for _ in range(N):
        opt.zero_grad(set_to_none=True); loss_of(r.render(dict(batch))).backward(); opt.step()
    e1 = ev(); torch.cuda.synchronize()
    print("back-to-back ms per step (again): %.3f" % (e0.elapsed_time(e1) / N), " plan builds:", getattr(opt, "plan_builds", None))

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss_of(r.render(dict(batch))).backward()
        opt.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=70))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))

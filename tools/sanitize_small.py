"""Small end-to-end exercise of every kernel family for compute-sanitizer: eval render (merged launches, tensor-core MLP and
deformer, TMA staging), the profiled per-part path, a multi-pass render, the peer-frame path at world 1, one training step with
the fused optimizer, rays / image / SSIM helpers.  usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from instant_nvr_b200.config import PathConfig
from instant_nvr_b200.network import Network
from instant_nvr_b200.optimizer import FusedAdam
from instant_nvr_b200.rays import assemble_image, psnr_metric, ssim_metric
from instant_nvr_b200.renderer import Renderer
from instant_nvr_b200.sharding import PeerFrame
from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays

cfg = PathConfig.inb_377(N_samples=16, log2_T_cap=12).with_(use_reg_distortion=True, perturb=1.0)
frame = make_frame(seed=1)
rays = make_rays(frame, 24, 24, drop_missing=True)
net = Network(cfg, device="cpu")
fill_weights(net.state_dict(), seed=1, table_gain=50.0, bounds=frame["bounds"][0])
net = net.cuda().eval()
gb = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**frame, **rays}.items()}
eng = net.engine()
o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
rgb, acc, raw = eng.render_rays(o, d, n, f, 16, batch=gb, want_raw=True)
eng.profile(True)
rgb2, acc2 = eng.render_rays(o, d, n, f, 16)
eng.profile_read(); eng.profile(False)
assert torch.equal(rgb, rgb2)
eng.max_points_per_pass = 37 * 16; eng._ws = None
rgb3, _ = eng.render_rays(o, d, n, f, 16)
eng.max_points_per_pass = 32 << 20; eng._ws = None
assert torch.equal(rgb, rgb3)
pf = PeerFrame(eng, o.shape[0], 0, 1, tile=64)
fr = pf.render(o, d, n, f, 16)
assert torch.equal(fr[:, :3], rgb)
pf.close()
# the same frame as a two-lane render (two passes on two engine streams, half a workspace each), device and host entry points
os.environ["NVR_TWO_LANE_MIN_SAMPLES"] = "1024"
from instant_nvr_b200.engine import Engine
eng2 = Engine(cfg)
del os.environ["NVR_TWO_LANE_MIN_SAMPLES"]
eng2.bind_params(net)
rgb4, acc4, raw4 = eng2.render_rays(o, d, n, f, 16, batch=gb, want_raw=True)
assert eng2.counters()["n_passes"] == 2 and torch.equal(rgb4, rgb) and torch.equal(raw4, raw)
print("two-lane footprint", eng2.gather_footprint())
host = [t.cpu().contiguous().pin_memory() for t in (o, d, n, f)]
rgb_h, acc_h = torch.empty(o.shape[0], 3).pin_memory(), torch.empty(o.shape[0]).pin_memory()
eng2.render_rays_host(*host, 16, rgb_h, acc_h)
assert torch.equal(rgb_h, rgb.cpu()) and torch.equal(acc_h, acc.cpu())
del eng2
print("footprint", eng.gather_footprint() if eng._ws is not None and eng.render_rays(o, d, n, f, 16) is not None else None)
net.train()
params = [p for p in net.parameters() if p.requires_grad]
opt = FusedAdam(params, lr=1e-3, eps=1e-15)
r = Renderer(net)
tgt = torch.rand(1, o.shape[0], 3, device="cuda")
for _ in range(2):
    opt.zero_grad(set_to_none=True)
    ret = r.render(dict(gb))
    loss = ((ret["rgb_map"] - tgt) ** 2).mean() + 0.1 * ret["reg_distortion_loss"].mean() + 0.1 * torch.norm(ret["resd"], dim=2).mean()
    if ret["oresd"].numel():
        loss = loss + 0.01 * (ret["oresd"] ** 2).mean()
    loss.backward()
    opt.step()
net.eval()
m = torch.zeros(24, 24, dtype=torch.bool, device="cuda")
m.view(-1)[rays["coord"][0].cuda()] = True
img = assemble_image(rgb, m)
print("psnr", psnr_metric(img, img * 0.9), "ssim", ssim_metric(img, img * 0.9, m))
torch.cuda.synchronize()
print("sanitize_small ok")

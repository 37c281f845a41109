#!/usr/bin/env python
"""Benchmark of the instant-nvr per-ray hot path on B200 (BASELINE.json metric: ray-samples/s at 512x512x128).

    python bench.py [--gpus N --steps K --warmup W]           our CUDA path (default: config c2, strong scaling)
    python bench.py --config c4|c5 ...                         BASELINE.json configs[3] / configs[4] as the main workload
    python bench.py --impl reference [...]                     the reference algorithm (oracle port) on the host cores
    python bench.py --impl reference-gpu [...]                 the same op-by-op PyTorch pipeline on the B200 ("existing GPU path")
    torchrun --nproc-per-node N bench.py --gpus N ...          one rank per GPU (the driver launches it so)

A step = one forward-only render of ONE frame through the shipped inb_377 network (1.14 GB of grid tables, reference
init, synthetic pseudo-SMPL frame): 512x512 pixels x 128 samples per ray for c2.  Like the reference (mask_at_box,
if_nerf_data_utils.py:92-107) only the rays that hit the subject's bounding box are rendered (90 % of the pixels);
`value` counts those rays only.  N GPUs ("scaling": "strong", the default): the frame's rays are dealt to the ranks in
interleaved 1024-ray tiles (instant_nvr_b200/sharding.py), every rank renders its tiles and the frame is assembled on
every rank -- total work is the same at every N.  ``--scaling weak`` renders N different views, one per rank's worth.

Prints ONE JSON line (rank 0).  value = whole-job ray-samples/s with the rays resident in HBM (CUDA events, max over
ranks); e2e = the same through the host-buffer C-ABI call (pinned host rays -> H2D -> render -> D2H pixels).  The default
run also measures, after the timed region, the supplementary lines the JSON carries: per-stage times (serialised,
profiled pass), the gather's reuse-aware roofline, c4 / c5 (every N), and at N = 1 the dense a = 1 variant, the
pre-summed inference tables, the training step, the reference pipeline on the GPU and the CPU baseline.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

CONFIGS = {   # BASELINE.json configs[1], [3], [4]
    "c2": dict(H=512, W=512, S=128, name="inb_377 512x512 px x 128 samples/ray, forward-only render (BASELINE.json configs[1])"),
    "c4": dict(H=1024, W=1024, S=64, name="1024x1024 px x 64 samples/ray novel-view render, ray tiles over the ranks (BASELINE.json configs[3])"),
    "c5": dict(H=2160, W=3840, S=256, name="synthetic 3840x2160 px x 256 samples/ray throughput sweep (BASELINE.json configs[4])"),
}
CPU_SAMPLE_SIDE = 64           # the CPU arm renders a 64x64-ray strided sub-grid of the same view
GPU_REF_SIDE = 128             # the reference-on-GPU arm renders a 128x128-ray strided sub-grid
BYTES_PER_PAIR = 16 * 8 * 16 * 4   # 16 levels x 8 corners x 16 fp32 features = 8192 B (no reuse assumed)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference", "reference-gpu"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="strong")
    ap.add_argument("--assemble", choices=["peer", "nccl"], default="peer",
                    help="N > 1: frame assembly by peer-memory stores from the compositing kernel + flag barrier (nvr_render_rays_frame) "
                         "or by the round-1 path (pad + NCCL all_gather + un-permute)")
    ap.add_argument("--emulate-shard-of", type=int, default=0, metavar="N",
                    help="development aid (not a bench line): on ONE GPU render only rank 0's shard of an N-rank run, i.e. the "
                         "per-rank work of `--gpus N` without the other ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the supplementary training-step measurement")
    ap.add_argument("--no-extras", action="store_true", help="only the main workload (value, e2e, stage times, roofline)")
    ap.add_argument("--steps-only", action="store_true",
                    help="only the warm-up + timed render steps (what an ncu launch list / --set full capture should see)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def csrc_hash():
    """Identity of the CUDA sources the loaded library was built from (ties an ncu capture under profiles/ to a binary)."""
    h = hashlib.sha256()
    d = os.path.join(REPO, "instant_nvr_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    h.update(open(os.path.join(REPO, "include", "nvr_b200.h"), "rb").read())
    return h.hexdigest()[:16]


_NVML = {}


def nvml_handle(index):
    """NVML initialised ONCE per process (nvmlInit takes ~100 ms: done inside the sampling thread it outlasted a 20-step
    timed region and the line carried no clock samples)."""
    if index not in _NVML:
        try:
            import pynvml as nv
            nv.nvmlInit()
            _NVML[index] = (nv, nv.nvmlDeviceGetHandleByIndex(index))
        except Exception as e:
            _NVML[index] = (None, f"nvml_unavailable:{type(e).__name__}")
    return _NVML[index]


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML while the timed region runs (every 5 ms, first sample at once)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        self.nv, self.h = nvml_handle(index)

    def run(self):
        nv, h = self.nv, self.h
        if nv is None:                  # NVML missing: report that rather than fail the bench
            self.reasons.add(h)
            return
        try:
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while True:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                if self.stop_flag:
                    break
                time.sleep(0.005)
        except Exception as e:
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def build_views(n_views, H, W, seed=0):
    """frame + the bbox-hitting rays of n_views views of it (views orbit the subject), concatenated."""
    from instant_nvr_b200.synthetic import make_frame, make_rays
    frame = make_frame(seed=seed)
    views = [make_rays(frame, H, W, azimuth_deg=(360.0 / max(n_views, 1)) * v, drop_missing=True) for v in range(n_views)]
    cat = lambda k: torch.cat([v[k][0] for v in views])
    return frame, {k: cat(k).contiguous() for k in ("ray_o", "ray_d", "near", "far")}


def device_weights(net, frame, seed=0):
    """Reference-init magnitudes (kaiming-normal tables / U(+-1/sqrt(fan_in)) linears), generated on the device."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            leaf = name.rsplit(".", 1)[-1]
            if leaf in ("dense", "hash"):
                emb = dict(net.named_modules())[name.rsplit(".", 1)[0]]
                std = (2.0 / (emb.spec.T * emb.spec.n_feat)) ** 0.5
                p.normal_(0.0, std, generator=g)
            elif leaf == "rgb_latent":
                p.normal_(0.0, (2.0 / p.shape[1]) ** 0.5, generator=g)
            elif leaf in ("weight", "bias") and p.requires_grad:
                fan_in = p.shape[1] if p.dim() == 2 else dict(net.named_parameters())[name[:-4] + "weight"].shape[1]
                p.uniform_(-1.0 / fan_in ** 0.5, 1.0 / fan_in ** 0.5, generator=g)
        for pid, part in enumerate(net.tpose_human.part_networks):
            part.embedder.bounds.copy_(frame["bounds"][0][pid])


def _oracle():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import nvr_oracle as O
    return O


def strided_rays(frame, H, W, side):
    """A side x side strided sub-grid of the H x W view's rays (bbox-hitting ones only)."""
    from instant_nvr_b200.synthetic import make_rays
    full = make_rays(frame, H, W)
    ys = torch.arange(side) * (H // side) + (H // side) // 2
    xs = torch.arange(side) * (W // side) + (W // side) // 2
    idx = (ys[:, None] * W + xs[None]).reshape(-1)
    idx = idx[full["mask_at_box"][0][idx]]
    return {k: full[k][:, idx].contiguous() for k in ("ray_o", "ray_d", "near", "far")}


def cpu_reference_run(sd, frame, cfgd, steps, warmup):
    """The reference algorithm (oracle port) on the host threads, on a strided sub-grid of the same view."""
    O = _oracle()
    from instant_nvr_b200.synthetic import make_rays
    S = cfgd["S"]
    # "all the host threads it can use": torch's CPU ops stop scaling (and regress badly) on many-core hosts,
    # so time a small probe at a few thread counts and keep the fastest
    probe = {**frame, **make_rays(frame, 24, 24)}
    best = (None, float("inf"))
    for nt in sorted({min(os.cpu_count() or 1, c) for c in (8, 16, 32, 64, 10**6)}):
        torch.set_num_threads(nt)
        with torch.no_grad():
            O.render(sd, probe, S, 0.05, want_raw=False)
            t0 = time.perf_counter()
            O.render(sd, probe, S, 0.05, want_raw=False)
            dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (nt, dt)
    torch.set_num_threads(best[0])
    rays = strided_rays(frame, cfgd["H"], cfgd["W"], CPU_SAMPLE_SIDE)
    batch = {**frame, **rays}
    n = rays["ray_o"].shape[1] * S
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.render(sd, batch, S, 0.05, want_raw=False)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    tot = sum(times)
    return {"value": n * len(times) / tot, "unit": "ray-samples/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": "port",
            "sample": f"{CPU_SAMPLE_SIDE}x{CPU_SAMPLE_SIDE}-ray strided sub-grid of the {cfgd['H']}x{cfgd['W']} view (bbox-hitting rays: "
                      f"{rays['ray_o'].shape[1]}) x {S} samples = {n} ray-samples per step, {len(times)} step(s); KNN by brute-force torch top-k (exact)",
            "ms_per_step": 1e3 * tot / len(times)}, n


def gpu_reference_run(sd, frame, cfgd, steps, warmup):
    """The reference's op-by-op PyTorch pipeline (the oracle port: same ops, 4096-ray chunks, brute-force KNN stand-in for
    pytorch3d) executed on the B200 -- the "existing GPU path" of SURVEY.md section 8(d)."""
    O = _oracle()
    S = cfgd["S"]
    rays = strided_rays(frame, cfgd["H"], cfgd["W"], GPU_REF_SIDE)
    to = lambda d: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}
    sd_g, batch = to(sd), to({**frame, **rays})
    n = rays["ray_o"].shape[1] * S
    ms = []
    with torch.no_grad(), torch.device("cuda"):
        for i in range(warmup + steps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            O.render(sd_g, batch, S, 0.05, want_raw=False)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ms.append(e0.elapsed_time(e1))
    tot = sum(ms)
    return {"value": n * len(ms) / (tot * 1e-3), "unit": "ray-samples/s", "ms_per_step": tot / len(ms), "kind": "port on cuda",
            "sample": f"{GPU_REF_SIDE}x{GPU_REF_SIDE}-ray strided sub-grid of the {cfgd['H']}x{cfgd['W']} view ({rays['ray_o'].shape[1]} bbox-hitting rays) x {S} "
                      f"samples = {n} ray-samples per step, {len(ms)} step(s), 4096-ray chunks like inb_renderer.py:217-237",
            "note": "plain torch ops on cuda:0 (eager, fp32, one nonzero()/gather chain per chunk); none of this repo's kernels"}


def ncu_traffic(kernel):
    """DRAM bytes (read + write) of `kernel`, summed over ONE step, from the newest ncu --set full summary under profiles/
    together with whether that capture was taken from the sources this library was built from."""
    import csv
    import glob
    # the capture taken from THIS library's sources first (meta.json carries the hash), then the newest by name (r2h > r2a > r1l:
    # file times mean nothing after a checkout)
    def rank(path):
        meta_path = path.replace(".csv", ".meta.json")
        try:
            same = json.load(open(meta_path)).get("csrc_hash") == csrc_hash()
        except Exception:
            same = False
        return (same, os.path.basename(path))
    files = sorted(glob.glob(os.path.join(REPO, "profiles", "*ncu_full_summary.csv")), key=rank)
    for path in reversed(files):
        try:
            rows = list(csv.reader(open(path)))
            h = rows[0]
            ri = next(i for i, c in enumerate(h) if c.startswith("dram__bytes_read.sum"))
            wi = next(i for i, c in enumerate(h) if c.startswith("dram__bytes_write.sum"))
            scale = lambda col: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[col.split("[")[1].rstrip("]")]
            vals = [float(r[ri]) * scale(h[ri]) + float(r[wi]) * scale(h[wi]) for r in rows[1:] if r[0].strip() == kernel]
            if not vals:
                continue
            meta_path = path.replace(".csv", ".meta.json")
            meta = json.load(open(meta_path)) if os.path.exists(meta_path) else {}
            steps = max(int(meta.get("steps_captured", 1)), 1)
            return {"bytes_per_step": sum(vals) / steps, "launches_per_step": len(vals) / steps, "source": os.path.relpath(path, REPO),
                    "same_sources_as_this_library": meta.get("csrc_hash") == csrc_hash(), "capture_csrc_hash": meta.get("csrc_hash")}
        except Exception:
            continue
    return None


def gather_uniform_roofline(eng, net, peak):
    """k_embed on 4 Mi points drawn uniformly in the body part's bounding box (no two points share fine-level
    rows): the no-reuse case the 8192 B/pair algorithmic figure describes."""
    n = 4 << 20
    b = net.tpose_human.part_networks[0].embedder.bounds.detach()
    x = (b[0] + (b[1] - b[0]) * torch.rand(n, 3, device="cuda")).contiguous()
    for _ in range(3):
        eng.embed_part(0, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        eng.embed_part(0, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    achieved = n * BYTES_PER_PAIR / (ms * 1e-3) / 1e9
    return {"kernel": "k_embed", "points": n, "part": "body", "ms": ms, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "note": "uniform random points in the body bbox, all 16 levels in one launch; 6 dense levels "
            "(52.7 MB) stay cache resident, the 10 hashed levels (671 MB) do not; includes the (n,19) fp32 output write"}


def mlp_tensor_report(prof, steps, mode):
    """Tensor-core work of the part-MLP kernel in the profiled steps.  mlp_mode 3 (k_mlp_f16): three kind::f16 MMAs
    (M=128, N=64, K=16) per K-step of the fp16-split product; modes 1 / 2 (k_mlp_tc): three kind::tf32 MMAs (K=8)."""
    pairs = prof["pairs"]
    k_elems = [32 + 48 + 64 + 64, 32 + 48 + 64, 32 + 48 + 64 + 64, 32 + 48 + 64, 32 + 48 + 64] if mode == 3 else \
        [24 + 48 + 64 + 64, 24 + 48 + 64, 24 + 48 + 64 + 64, 24 + 48 + 64, 24 + 48 + 64]
    kk = 16 if mode == 3 else 8
    mmas = sum(((p + 127) // 128) * 3 * (k // kk) for p, k in zip(pairs, k_elems))
    flops = mmas * 2 * 128 * 64 * kk
    ms = prof["ms"]["mlp"]
    n_sm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    # one tcgen05.mma of M=128, N=64 occupies the tensor pipe for max(M,128) * N / 256 = 32 cycles whatever the kind
    pipe_ms = mmas * 32 / n_sm / 1.965e9 * 1e3
    # MUFU floor: 2 MUFU (ex2 + lg2) per hidden activation, 16 lanes per clock and SM
    acts = sum(p * (192 if i in (0, 2) else 128) for i, p in enumerate(pairs))
    mufu_ms = acts * 2 / 16 / n_sm / 1.965e9 * 1e3
    return {"kernel": "k_mlp_f16" if mode == 3 else "k_mlp_tc", "operands": "fp16-split (hi+lo), kind::f16" if mode == 3 else "3xTF32, kind::tf32",
            "mma_flops_per_step": flops / steps, "ms": ms / steps, "achieved_tflops": flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0,
            "tensor_pipe_busy_estimate": pipe_ms / ms if ms > 0 else 0.0, "mufu_floor_frac": mufu_ms / ms if ms > 0 else 0.0,
            "note": "tensor_pipe_busy_estimate = issued tcgen05.mma x 32 cycles / (SMs x 1.965 GHz x kernel time); mufu_floor_frac = "
                    "(2 MUFU per softplus x activations / 16 per clock per SM) / kernel time: the unit that bounds this kernel (ncu: "
                    "MIO-throttle stalls on the ex2 / lg2 lines); the measured sm__pipe_tensor_cycles_active is in profiles/*ncu_full_summary.csv"}


def train_step_report(net, gframe, frame, n_rays=1024, n_samples=64, steps=10):
    """BASELINE.json configs[2]: one training step = Renderer.render in training mode on 1024 rays x 64 samples
    (stratified jitter, pair + distortion regularisers) + image loss + backward + Adam step, full-size tables.
    Timed twice: with the fused optimizer step (nvr_adam_step) and with torch.optim.Adam (what the reference builds)."""
    import dataclasses
    from instant_nvr_b200.optimizer import FusedAdam
    from instant_nvr_b200.renderer import Renderer
    from instant_nvr_b200.synthetic import make_rays
    cfg0 = net.cfg
    net.cfg = dataclasses.replace(cfg0, N_samples=n_samples, perturb=1.0, use_reg_distortion=True)
    try:
        rays = make_rays(frame, 32, 32)
        batch = {**gframe, **{k: v.cuda() for k, v in rays.items()}}
        target = torch.rand(1, n_rays, 3, device="cuda")
        params = [p for p in net.parameters() if p.requires_grad]
        r = Renderer(net)
        net.train()
        out = {}
        for name, make in (("fused_adam", lambda: FusedAdam(params, lr=5e-4, eps=1e-15)),
                           ("torch_adam", lambda: torch.optim.Adam(params, lr=5e-4, eps=1e-15))):
            opt = make()

            def step():
                opt.zero_grad(set_to_none=True)
                ret = r.render(dict(batch))
                loss = ((ret["rgb_map"] - target) ** 2).mean() + 0.1 * ret["reg_distortion_loss"].mean() \
                    + 0.1 * torch.norm(ret["resd"], dim=2).mean()
                if ret["oresd"].numel():
                    loss = loss + 0.01 * (ret["oresd"] ** 2).mean()
                loss.backward()
                opt.step()
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) / steps
            # the optimizer step alone (gradients left in place from the last step)
            e0.record()
            for _ in range(steps):
                opt.step()
            e1.record()
            torch.cuda.synchronize()
            out[name + "_opt_only"] = e0.elapsed_time(e1) / steps
            del opt
            torch.cuda.empty_cache()
        ms = out["fused_adam"]
        n_par = sum(p.numel() for p in params)
        return {"workload": f"training step, {n_rays} rays x {n_samples} samples, fwd + bwd + Adam on {n_par / 1e6:.0f} M parameters",
                "ms_per_step": ms, "ray_samples_per_sec": n_rays * n_samples / (ms * 1e-3),
                "ms_per_step_torch_adam": out["torch_adam"],
                "optimizer_ms": {"nvr_adam_step": out["fused_adam_opt_only"], "torch.optim.Adam": out["torch_adam_opt_only"]},
                "optimizer_gbs": n_par * 28 / (out["fused_adam_opt_only"] * 1e-3) / 1e9,
                "note": "backward returns dense gradients like the reference's autograd (1.14 GB zero-fill + scatter); the "
                        "optimizer step is nvr_adam_step (28 B per parameter, one pass); everything runs in libnvr_b200.so"}
    finally:
        net.eval()
        net.cfg = cfg0


def grad_allreduce_check(rank, world):
    """tests/test_gpu_multi.py::test_two_rank_gradient_allreduce_equals_full_batch inside the multi-GPU bench run (the driver's
    GPU test box has one GPU and skips it): every rank back-propagates its 1/world of the rays of one small training batch, the
    gradients are all-reduced over NCCL and compared with the single-GPU gradient of the whole batch (the loss is a sum over
    rays) -- what the reference's DDP wrapper relies on."""
    import torch.distributed as dist
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.renderer import Renderer
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    cfg = PathConfig.inb_377(N_samples=16, log2_T_cap=12).with_(use_pair_reg=False)
    frame = make_frame(seed=5)
    rays = make_rays(frame, 24, 24)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=5, table_gain=100.0, bounds=frame["bounds"][0])
    net = net.cuda().train()
    gb = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**frame, **rays}.items()}
    R = gb["ray_o"].shape[1]
    tgt = torch.rand(1, R, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    r = Renderer(net)

    def grads_for(sel):
        for p in net.parameters():
            p.grad = None
        b = dict(gb)
        for k in ("ray_o", "ray_d", "near", "far", "occupancy"):
            b[k] = gb[k][:, sel]
        ret = r.render(b)
        (((ret["rgb_map"] - tgt[:, sel]) ** 2).sum() + ret["resd"].pow(2).sum()).backward()
        return {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in net.named_parameters() if p.requires_grad}
    full = grads_for(torch.arange(R, device="cuda"))
    mine = grads_for(torch.arange(rank, R, world, device="cuda"))
    worst = torch.zeros(1, device="cuda")
    for n, g in mine.items():
        dist.all_reduce(g)
        worst = torch.maximum(worst, (g - full[n]).abs().max() / (full[n].abs().max() + 1e-12))
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    w = float(worst.item())
    return {"worst_relative_difference": w, "tensors": len(mine), "rays": int(R), "ok": w < 1e-3,
            "what": "all-reduced per-rank gradients (each rank: 1/world of the rays) vs the single-GPU gradient of the whole batch"}


def cpu_state_dict(S, seed=0):
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import make_frame
    cfg = PathConfig.inb_377(N_samples=S)
    frame = make_frame(seed=seed)
    torch.manual_seed(seed)
    net = Network(cfg, device="cpu")       # constructors give the reference init (kaiming tables, default Linear)
    with torch.no_grad():
        for pid, part in enumerate(net.tpose_human.part_networks):
            part.embedder.bounds.copy_(frame["bounds"][0][pid])
    return net.state_dict(), frame


class Workload:
    """One config's rays sharded over the ranks + the two step functions (device-resident and end-to-end)."""

    def __init__(self, eng, gframe, frame_cpu, cfgd, n_views, rank, world, assemble_mode="peer", shard_of=0):
        from instant_nvr_b200.sharding import PeerFrame, shard_indices
        from instant_nvr_b200.synthetic import make_rays
        self.eng, self.gframe, self.cfgd, self.rank, self.world = eng, gframe, cfgd, rank, world
        H, W = cfgd["H"], cfgd["W"]
        views = [make_rays(frame_cpu, H, W, azimuth_deg=(360.0 / max(n_views, 1)) * v, drop_missing=True) for v in range(n_views)]
        rays = {k: torch.cat([v[k][0] for v in views]).contiguous() for k in ("ray_o", "ray_d", "near", "far")}
        self.n_views, self.n_pixels = n_views, n_views * H * W
        self.n_total = rays["ray_o"].shape[0]
        idx = shard_indices(self.n_total, rank, world) if not shard_of else shard_indices(self.n_total, 0, shard_of)
        self.host = {k: v[idx].contiguous().pin_memory() for k, v in rays.items()}
        self.dev = {k: v.cuda() for k, v in self.host.items()}
        self.n_local = idx.numel()
        self.rgb_h, self.acc_h = torch.empty(self.n_local, 3).pin_memory(), torch.empty(self.n_local).pin_memory()
        self.samples_per_step = self.n_total * cfgd["S"]
        self.pf = PeerFrame(eng, self.n_total, rank, world) if world > 1 and assemble_mode == "peer" else None

    def close(self):
        if self.pf is not None:
            self.pf.close()
            self.pf = None

    def step_device(self):
        from instant_nvr_b200.sharding import assemble
        eng, dev = self.eng, self.dev
        eng.bind_frame(self.gframe, force=True)      # every step is a new frame: per-frame preparation is inside the timed region
        if self.pf is not None:                      # the compositing kernel stores into every rank's frame + one flag barrier
            return self.pf.render(dev["ray_o"], dev["ray_d"], dev["near"], dev["far"], self.cfgd["S"])
        rgb, acc = eng.render_rays(dev["ray_o"], dev["ray_d"], dev["near"], dev["far"], self.cfgd["S"])
        if self.world > 1:
            return assemble(torch.cat([rgb, acc[:, None]], 1), self.n_total, self.rank, self.world)
        return rgb

    def step_e2e(self):
        from instant_nvr_b200.sharding import assemble
        eng, host = self.eng, self.host
        eng.bind_frame(self.gframe, force=True)
        if self.world == 1:
            eng.render_rays_host(host["ray_o"], host["ray_d"], host["near"], host["far"], self.cfgd["S"], self.rgb_h, self.acc_h)
        else:
            # every rank: its own tiles in from pinned host memory, the frame assembled on every GPU, its own tiles back out
            if self.pf is not None:              # one library call: copies on the lanes' streams, render, flag barrier, sync
                self.pf.render_host(host["ray_o"], host["ray_d"], host["near"], host["far"], self.cfgd["S"], self.rgb_h, self.acc_h)
                return
            d = {k: v.cuda(non_blocking=True) for k, v in host.items()}
            rgb, acc = eng.render_rays(d["ray_o"], d["ray_d"], d["near"], d["far"], self.cfgd["S"])
            assemble(torch.cat([rgb, acc[:, None]], 1), self.n_total, self.rank, self.world)
            self.rgb_h.copy_(rgb, non_blocking=True)
            self.acc_h.copy_(acc, non_blocking=True)
            torch.cuda.current_stream().synchronize()


def main():
    args = parse()
    rank, local_rank, world = dist_env()
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version there) go to stderr
    out_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(out_fd, (json.dumps(line) + "\n").encode())

    cfgd = CONFIGS[args.config]
    workload = cfgd["name"]

    # ------------------------------------------------------------------ reference arms
    if args.impl in ("reference", "reference-gpu"):
        if rank != 0:
            return
        sd, frame = cpu_state_dict(cfgd["S"])
        steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
        if args.impl == "reference":
            base, n = cpu_reference_run(sd, frame, cfgd, steps, warmup)
        else:
            if not torch.cuda.is_available():
                raise SystemExit("bench.py --impl reference-gpu: no CUDA device")
            base = gpu_reference_run(sd, frame, cfgd, steps, warmup)
            base["cores"] = 0
        line = {"impl": args.impl, "metric": "ray_samples_per_sec", "value": base["value"], "unit": "ray-samples/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "sample": base["sample"]},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.engine import Engine
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import make_frame

    cfg = PathConfig.inb_377(N_samples=cfgd["S"])
    n_views = world if args.scaling == "weak" else 1
    frame = make_frame(seed=0)
    with torch.device("cuda"):
        net = Network(cfg)
    net = net.cuda().eval()
    device_weights(net, frame)
    gframe = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in frame.items()}
    eng = net.engine()
    eng.bind_frame(gframe)
    wl = Workload(eng, gframe, frame, cfgd, n_views, rank, world, args.assemble, args.emulate_shard_of)

    def timed(e, fn, steps, warmup, profile=False, mark=False):
        # mark: cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off sees exactly them)
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = e.counters()["kernel_launches"]
        if profile:
            e.profile(True)
        if mark:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if mark:
            torch.cuda.profiler.stop()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sampler.stop_flag = True
        sampler.join(2.0)
        prof = e.profile_read() if profile else None
        if profile:
            e.profile(False)
        launches = e.counters()["kernel_launches"] - launches0
        return float(ms.item()), sampler.summary(), prof, launches

    W = max(args.warmup, 3)
    if args.steps_only:
        ms_plain, _, _, _ = timed(eng, wl.step_device, args.steps, args.warmup, mark=bool(os.environ.get("NVR_MARK_PLAIN")))
        ms, clocks, prof, launches = timed(eng, wl.step_device, args.steps, args.warmup, profile=True, mark=not os.environ.get("NVR_MARK_PLAIN"))
        if rank == 0:
            emit({"metric": "ray_samples_per_sec", "value": wl.samples_per_step * args.steps / (ms * 1e-3), "unit": "ray-samples/s",
                  "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                  "ms_per_step_unprofiled": ms_plain / args.steps,
                  "gpu_launches": launches, "stage_ms_per_step": {k: v / args.steps for k, v in prof["ms"].items()},
                  "mlp_mode": eng.mlp_mode, "tune": eng.tune, "csrc_hash": csrc_hash(), "rays_per_gpu": wl.n_local,
                  "note": "--steps-only: profiler-facing run, not a bench line"})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the headline: K steps, rays resident in HBM; then the same through host buffers
    ms, clocks, _, launches = timed(eng, wl.step_device, args.steps, W)
    value = wl.samples_per_step * args.steps / (ms * 1e-3)
    ms_e2e, clocks_e2e, _, _ = timed(eng, wl.step_e2e, args.steps, W)
    e2e_value = wl.samples_per_step * args.steps / (ms_e2e * 1e-3)

    # ---- the same steps once more with one CUDA-event pair around every launch (nvr_profile): per-stage / per-part
    #      kernel durations and the work counters of rank 0
    p_steps = max(3, min(args.steps, 10))
    ms_prof, _, prof, _ = timed(eng, wl.step_device, p_steps, 1, profile=True)
    per = lambda x: x / p_steps
    pairs = sum(prof["pairs"])
    stage_ms = {k: per(v) for k, v in prof["ms"].items()}
    stage_share = {k: (v / sum(stage_ms.values()) if sum(stage_ms.values()) else 0.0) for k, v in stage_ms.items()}

    # ---- roofline of the dominant kernel (the grid gather), rank 0's launches of the profiled steps.
    # Algorithmic bytes = the COMPULSORY table traffic of a launch: the distinct 32-byte sectors its pair list touches
    # (counted on the device with the gather's own index arithmetic, nvr_gather_footprint) x 32 B, plus the pair records it
    # reads (32 B) and the embedding rows it writes (80 B).  SURVEY.md 8(d)'s 8192 B/pair assumes no reuse between pairs,
    # which neighbouring samples violate by 5-20x; it is kept as `no_reuse` for reference and measured for real by
    # `roofline_uniform`.
    peak, peak_src, peaks = measured_peaks()
    embed_ms = stage_ms["embed"]
    roofline = None
    if rank == 0:
        single_pass = wl.n_local * cfgd["S"] <= eng.max_points_per_pass - 64
        n_embed = max(prof["launches"]["embed"], 1) / p_steps
        per_part = []
        uniq_total = 0.0
        if single_pass:
            eng.render_rays(wl.dev["ray_o"], wl.dev["ray_d"], wl.dev["near"], wl.dev["far"], cfgd["S"])   # this rank's pass, no collective
            uniq = eng.gather_footprint()
            names = ("body", "leg", "head", "larm", "rarm")
            for p in range(5):
                pp = per(prof["pairs"][p])
                t = per(prof["embed_part_ms"][p])
                b = uniq[p] * 32.0 + pp * (32 + 80)
                uniq_total += b
                per_part.append({"part": names[p], "pairs": pp, "ms": t, "unique_table_bytes": uniq[p] * 32.0,
                                 "compulsory_bytes": b, "compulsory_gbs": b / (t * 1e-3) / 1e9 if t > 0 else 0.0,
                                 "frac": b / (t * 1e-3) / 1e9 / peak if t > 0 else 0.0,
                                 "no_reuse_bytes": pp * BYTES_PER_PAIR, "reuse_factor": pp * BYTES_PER_PAIR / max(uniq[p] * 32.0, 1.0)})
        achieved = uniq_total / (embed_ms * 1e-3) / 1e9 if embed_ms > 0 and uniq_total else 0.0
        no_reuse = per(pairs) * BYTES_PER_PAIR / (embed_ms * 1e-3) / 1e9 if embed_ms > 0 else 0.0
        tr = ncu_traffic("k_embed")
        roofline = {"bound": "hbm", "kernel": "k_embed", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (tr["bytes_per_step"] / max(tr["launches_per_step"], 1)) if tr else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": uniq_total / n_embed if n_embed else None,
                    "avg_launch_ms": embed_ms / n_embed if n_embed else None, "launches_per_step": n_embed,
                    "definition": "achieved = compulsory bytes of one step's gather launches (distinct 32-B table sectors touched, counted on "
                                  "device per part, + 32 B pair record read + 80 B embedding row written per pair) / summed CUDA-event "
                                  "duration of those launches in the profiled (serialised) steps, rank 0",
                    "per_part": per_part,
                    "no_reuse": {"bytes_per_pair": BYTES_PER_PAIR, "achieved": no_reuse, "frac": no_reuse / peak,
                                 "note": "SURVEY.md 8(d)'s figure (16 levels x 8 corners x 64 B, no reuse between pairs): > 1 because "
                                         "neighbouring samples share rows, i.e. not a roofline fraction; see per_part.reuse_factor"},
                    "dram": None if not tr else {
                        "bytes_per_step": tr["bytes_per_step"], "gbs": tr["bytes_per_step"] / (embed_ms * 1e-3) / 1e9 if embed_ms > 0 else None,
                        "frac_of_peak": tr["bytes_per_step"] / (embed_ms * 1e-3) / 1e9 / peak if embed_ms > 0 else None,
                        "over_compulsory": tr["bytes_per_step"] / uniq_total if uniq_total else None,
                        "source": tr["source"], "same_sources_as_this_library": tr["same_sources_as_this_library"],
                        "note": "dram__bytes_read+write of every k_embed launch of one step (ncu --set full, cold-cache replays) over the "
                                "live event time of this run: what the kernel actually moved"}}
    # in situ the gather of the cache-resident parts is bound by the L1 request path: a 64-byte table row is half of a
    # 128-byte line, every corner row of every pair costs one L1 wavefront, and an SM retires one wavefront per clock
    sm_clock = (clocks.get("sm_mhz") or 1965) * 1e6
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    wavefronts = per(pairs) * 16 * 8
    l1_floor_ms = wavefronts / (n_sm * sm_clock) * 1e3
    roofline_l1 = {"kernel": "k_embed", "bound": "l1 wavefronts", "wavefronts_per_step": wavefronts, "floor_ms": l1_floor_ms, "measured_ms": embed_ms,
                   "frac": l1_floor_ms / embed_ms if embed_ms > 0 else 0.0,
                   "note": "16 levels x 8 corners x 1 wavefront (one 64 B row = half a 128 B line) per pair at 1 wavefront/clk/SM "
                           f"({n_sm} SMs x {sm_clock / 1e6:.0f} MHz, B300_MICROARCH.md 'L1tex wavefront queue'); the in-situ limit of "
                           "this data layout whatever the cache hit rate"}
    # FP32 ceilings of the two ALU-bound kernels: instructions are not counted live, so the statement is time-based --
    # the stage's share of the step and what the FP32 pipes could do in that time
    fp32_peak_tflops = n_sm * 128 * 2 * sm_clock / 1e12
    alu = {"fp32_peak_tflops": fp32_peak_tflops,
           "knn": {"ms": stage_ms["knn"], "survivors": per(prof["survivors"]),
                   "ns_per_survivor_part": stage_ms["knn"] * 1e6 / max(per(prof["survivors"]) * 5, 1)},
           "warp": {"ms": stage_ms["warp"], "pairs": per(pairs),
                    "deformer_mlp_flops": per(pairs) * 3456, "lbs_flops_min": per(pairs) * (4 * 24 * 2 + 24 * 2 * 2 + 120),
                    "fp32_frac_deformer_only": per(pairs) * 3456 / (stage_ms["warp"] * 1e-3) / 1e12 / fp32_peak_tflops if stage_ms["warp"] > 0 else 0.0},
           "note": "k_knn / k_warp are FP32-ALU + issue bound (ncu: 68 % / 59 % issue-active); the deformer MLP is 3456 FLOP per pair "
                   "(SURVEY.md 8(d)); fractions are of 148 SMs x 128 lanes x 2 FLOP x SM clock"}

    # ---- N > 1: the frame every rank assembled must BE the single-GPU frame (rays are independent): rank 0 renders the whole
    #      frame alone and every rank compares its assembled copy bit for bit (the check tests/test_gpu_multi.py makes, here
    #      inside the scaling run itself)
    frame_check = None
    if world > 1:
        try:
            from instant_nvr_b200.synthetic import make_rays
            assembled = wl.step_device().clone()
            full = make_rays(frame, cfgd["H"], cfgd["W"], drop_missing=True)
            ref = torch.empty(wl.n_total, 4, device="cuda")
            if rank == 0:
                fr = {k: full[k][0].cuda() for k in ("ray_o", "ray_d", "near", "far")}
                r_rgb, r_acc = eng.render_rays(fr["ray_o"], fr["ray_d"], fr["near"], fr["far"], cfgd["S"])
                ref = torch.cat([r_rgb, r_acc[:, None]], 1).contiguous()
            dist.broadcast(ref, 0)
            ok = torch.tensor([int(torch.equal(assembled[:, :4], ref))], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            frame_check = {"assembled_frame_equals_single_gpu_render_bitwise_on_every_rank": bool(ok.item()),
                           "max_abs_diff_rank0": float((assembled[:, :4] - ref).abs().max().item()), "rays": wl.n_total}
        except Exception as ex:
            frame_check = {"error": f"{type(ex).__name__}: {ex}"}
    grad_check = None
    if world > 1 and not args.no_extras:
        try:
            grad_check = grad_allreduce_check(rank, world)
        except Exception as ex:
            grad_check = {"error": f"{type(ex).__name__}: {ex}"}
    extras = {}
    wl.close()                                       # one peer frame buffer per engine: the supplementary workloads bring their own
    if not args.no_extras:
        # ---- c4 / c5 (BASELINE.json configs[3], [4]) at this N: one frame split over the ranks, a few steps each
        for name in ("c4", "c5"):
            if name == args.config:
                continue
            try:
                cd = CONFIGS[name]
                w2 = Workload(eng, gframe, frame, cd, 1, rank, world, args.assemble)
                st = 5 if name == "c4" else 2
                m, _, _, _ = timed(eng, w2.step_device, st, 1)
                m2, _, _, _ = timed(eng, w2.step_e2e, st, 1)
                extras[name] = {"workload": cd["name"], "value": w2.samples_per_step * st / (m * 1e-3), "unit": "ray-samples/s",
                                "ms_per_frame": m / st, "steps": st, "warmup": 1, "rays_rendered": w2.n_total, "pixels": w2.n_pixels,
                                "rays_per_gpu": w2.n_local, "scaling": "strong",
                                "e2e": {"value": w2.samples_per_step * st / (m2 * 1e-3), "ms_per_frame": m2 / st,
                                        "h2d_bytes_per_step": w2.n_local * 32 * world, "d2h_bytes_per_step": w2.n_local * 16 * world}}
                w2.close()
                del w2
                torch.cuda.empty_cache()
            except Exception as ex:     # a supplementary line must never cost the headline
                extras[name] = {"error": f"{type(ex).__name__}: {ex}"}
        if rank == 0:
            try:
                extras["roofline_uniform"] = gather_uniform_roofline(eng, net, peak)
            except Exception as ex:
                extras["roofline_uniform"] = {"error": f"{type(ex).__name__}: {ex}"}
    if world == 1 and not args.no_extras:
        # ---- SURVEY.md 8(d) dense variant: cull off, every sample flagged in exactly one part (a = 1)
        try:
            engd = Engine(cfg, tune=16)
            engd.bind_params(net)
            wd = Workload(engd, gframe, frame, cfgd, 1, rank, world)
            md, _, profd, _ = timed(engd, wd.step_device, 3, 1, profile=True)
            pd = sum(profd["pairs"]) / 3
            extras["dense_a1"] = {"value": wd.samples_per_step * 3 / (md * 1e-3), "unit": "ray-samples/s", "ms_per_step": md / 3,
                                  "active_pairs_per_sample": pd / wd.samples_per_step, "pairs_per_part": [p / 3 for p in profd["pairs"]],
                                  "stage_ms_per_step": {k: v / 3 for k, v in profd["ms"].items()},
                                  "gather_no_reuse_gbs": pd * BYTES_PER_PAIR / (profd["ms"]["embed"] / 3 * 1e-3) / 1e9,
                                  "target": "SURVEY.md 8(d): >= 0.68 G ray-samples/s (70 % of the a = 1 HBM roofline)",
                                  "note": "NVR_TUNE_DENSE_A1: every sample survives the cull and is flagged in exactly one part (nearest "
                                          "by weighted neighbour distance); samples far from the subject clamp onto the parts' boundary "
                                          "cells, so their rows are cache hits -- the no-reuse bandwidth figure again exceeds the HBM peak"}
            del engd, wd
            torch.cuda.empty_cache()
        except Exception as ex:
            extras["dense_a1"] = {"error": f"{type(ex).__name__}: {ex}"}
        # ---- the same render with the opt-in pre-summed inference tables (nvr_prepare_inference)
        try:
            eng.inference_tables = True
            ms_p, _, prof_p, _ = timed(eng, wl.step_device, 5, 2, profile=True)
            eng.inference_tables = False
            eng._check(eng.lib.nvr_prepare_inference(eng._h, 0, None), "nvr_prepare_inference")
            eng._tables_key = None
            extras["inference_tables"] = {"value": wl.samples_per_step * 5 / (ms_p * 1e-3), "unit": "ray-samples/s", "ms_per_step": ms_p / 5,
                                          "stage_ms_per_step": {k: v / 5 for k, v in prof_p["ms"].items()},
                                          "note": "opt-in: sum_f of every table row taken once per weight update (286 MB of sums), 4 B per "
                                                  "corner instead of 64 B; NOT the headline, whose gather reads the reference's full "
                                                  "tables; timed with per-launch events (serialised)"}
        except Exception as ex:
            extras["inference_tables"] = {"error": f"{type(ex).__name__}: {ex}"}
        if not args.no_train:
            try:
                extras["train_step"] = train_step_report(net, gframe, frame)
            except Exception as ex:
                extras["train_step"] = {"error": f"{type(ex).__name__}: {ex}"}
        sd_cpu = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        try:
            extras["reference_gpu"] = gpu_reference_run(sd_cpu, frame, cfgd, 2, 1)
        except Exception as ex:
            extras["reference_gpu"] = {"error": f"{type(ex).__name__}: {ex}"}
        if not args.no_cpu_baseline:
            cpu_base, _ = cpu_reference_run(sd_cpu, frame, cfgd, 3, 1)
            extras["cpu_baseline"] = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        S = cfgd["S"]
        line = {
            "metric": "ray_samples_per_sec", "value": value, "unit": "ray-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "views_per_step": n_views, "pixels_per_view": cfgd["H"] * cfgd["W"],
                       "rays_rendered": wl.n_total, "rays_per_gpu": wl.n_local, "samples_per_ray": S,
                       "rays_note": "rays = the pixels whose ray hits the subject's bounding box (near < far), as the reference's "
                                    "mask_at_box selects them; the remaining pixels are background and are not rendered or counted",
                       "sharding": ("single GPU" if world == 1 else "interleaved 1024-ray tiles; frame assembled by peer-memory stores from the "
                                    "compositing kernel + 1 flag barrier (nvr_render_rays_frame)" if args.assemble == "peer" else
                                    "interleaved 1024-ray tiles + pad + NCCL all_gather + un-permute"),
                       "l2": "no flush: each step streams 1.14 GB of tables + ~GBs of workspace, far above the 126 MB L2",
                       "survivor_fraction": per(prof["survivors"]) / (wl.n_local * S),
                       "active_pairs_per_sample": per(pairs) / (wl.n_local * S),
                       "pairs_per_step_per_part": [per(p) for p in prof["pairs"]],
                       "far_pairs_per_step_per_part": [per(p) for p in prof.get("far_pairs", [0] * 5)],
                       "pairs_note": "pairs = (sample, part) pairs evaluated (gather + MLPs); far pairs = flagged pairs of parts "
                                     "farther than ~0.73 m (Gaussian weights sum < 1e-20), all answered by ONE shared evaluation per "
                                     "part (NVR_TUNE=8 evaluates each on its own; results agree to fp32 rounding, "
                                     "tests/test_gpu_parity.py::test_far_field_pairs_share_one_evaluation)"},
            "e2e": {"value": e2e_value, "unit": "ray-samples/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": wl.n_local * 32 * world, "d2h_bytes_per_step": wl.n_local * 16 * world,
                    "api": "nvr_render_rays_host (pinned host rays -> H2D -> render -> D2H rgb/acc)" if world == 1 else
                           "per rank: nvr_render_rays_frame_host (pinned host rays of its tiles -> H2D -> render -> frame assembled on every GPU -> D2H of its tiles)"},
            "gpu_launches": launches,
            "clocks": clocks, "clocks_e2e": clocks_e2e,
            "roofline": roofline,
            "roofline_l1_insitu": roofline_l1,
            "roofline_alu": alu,
            "mlp_tensor": mlp_tensor_report(prof, p_steps, eng.mlp_mode),
            "stage_ms_per_step": stage_ms, "stage_share": stage_share,
            "stage_note": f"per-launch CUDA events over {p_steps} extra steps with every launch serialised on one stream "
                          f"({ms_prof / p_steps:.3f} ms/step in that mode); the headline steps run without the events",
            "embed_part_ms": [per(v) for v in prof["embed_part_ms"]], "mlp_part_ms": [per(v) for v in prof["mlp_part_ms"]],
            "csrc_hash": csrc_hash(),
            "frame_check": frame_check, "grad_allreduce_check": grad_check,
            "cpu_baseline": extras.pop("cpu_baseline", None),
        }
        line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

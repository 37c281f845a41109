"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle and the
reference-generated golden fixtures.  Run on the B200 box with ``pytest -m gpu``.

Tolerances (north_star: per-sample occ/rgb within 1e-4, PSNR within 0.1 dB):
  * raw = [r,g,b,occ] lives in [0,1]; "within 1e-4" is applied as |ours - ref| <= 1e-4 * max(1, |ref|).
  * The reference extrapolates the grid outside a part's bbox with corner weights ~(1+|o|)^3 of mixed
    sign, so its own fp32 result there is summation-order noise (tests/test_host_emul.py shows the
    effect against fp64).  Samples whose arg-max part was evaluated outside its bbox are therefore
    compared at a looser, stated bound and their count is reported; with the reference's own init
    magnitudes (table_gain 1) every sample must pass the strict bound.
  * Decision flips: a sample whose cull distance / part distance is within 2e-6 of smpl_thresh may
    legitimately land on the other side in fp32 (FMA contraction); such samples are excluded and counted.
Every test appends its measured errors to gpurun_out/diag.jsonl.
"""
import json
import os
import sys
import time

import numpy as np
import pytest
import torch

from conftest import REPO, load_golden

sys.path.insert(0, os.path.join(REPO, "oracle"))
import nvr_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
DIAG = os.path.join(REPO, "gpurun_out", "diag.jsonl")


def diag(name, **kw):
    os.makedirs(os.path.dirname(DIAG), exist_ok=True)
    rec = {"test": name}
    rec.update({k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in kw.items()})
    with open(DIAG, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print("[diag]", rec)


def to_cuda(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


@pytest.fixture(scope="module")
def gpu(golden_setup):
    """Golden config on the GPU: weights gain 200 (stress) with a gain-1 twin."""
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import fill_weights
    cfg, frame = golden_setup["cfg"], golden_setup["frame"]
    nets, sds = {}, {}
    for gain in (1.0, 200.0):
        net = Network(cfg, device="cpu")
        fill_weights(net.state_dict(), seed=golden_setup["seed"], table_gain=gain, bounds=frame["bounds"][0])
        sds[gain] = {k: v.clone() for k, v in net.state_dict().items()}
        nets[gain] = net.cuda().eval()
    return dict(cfg=cfg, frame=frame, rays=golden_setup["rays"], batch=golden_setup["batch"],
                gbatch=to_cuda(golden_setup["batch"]), nets=nets, sds=sds)


def test_native_library_is_loaded(gpu):
    eng = gpu["nets"][1.0].engine()
    with open("/proc/self/maps") as f:
        assert "libnvr_b200.so" in f.read()
    from instant_nvr_b200 import cabi
    assert eng.lib.nvr_abi_version() == cabi.ABI_VERSION


def _inside(x, bounds):
    u = (x - bounds[0]) / (bounds[1] - bounds[0])
    return ((u >= 0) & (u <= 1)).all(-1)


def test_embed_part(gpu):
    g = torch.Generator().manual_seed(11)
    net, sd = gpu["nets"][200.0], gpu["sds"][200.0]
    eng = net.engine()
    for pid in range(5):
        b = sd[f"tpose_human.part_networks.{pid}.embedder.bounds"]
        n = 5003                                             # not a multiple of 8 / 64
        x = (b[0] + (b[1] - b[0]) * (torch.rand(n, 3, generator=g) * 1.3 - 0.15)).contiguous()
        ours = eng.embed_part(pid, x.cuda()).cpu()
        ref = O.grid_embed(sd, f"tpose_human.part_networks.{pid}.embedder.", x, True)
        ins = _inside(x, b)
        err_in = (ours - ref)[ins].abs().max().item()
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items() if f"part_networks.{pid}.embedder." in k}
        exact = O.grid_embed(sd64, f"tpose_human.part_networks.{pid}.embedder.", x.double(), True)
        e_ours = (ours.double() - exact)[~ins].pow(2).mean().sqrt().item()
        e_ref = (ref.double() - exact)[~ins].pow(2).mean().sqrt().item()
        diag("embed_part", part=pid, n_inside=int(ins.sum()), max_err_inside=err_in, rms_out_ours=e_ours, rms_out_oracle=e_ref)
        assert err_in < 3e-6 + 1e-5 * ref[ins].abs().max().item(), (pid, err_in)
        assert e_ours <= 3 * e_ref + 1e-6, (pid, e_ours, e_ref)


def test_deformer(gpu):
    g = torch.Generator().manual_seed(12)
    frame = gpu["frame"]
    tb = frame["tbounds"][0]
    x0 = (tb[0] + (tb[1] - tb[0]) * (torch.rand(3001, 3, generator=g) * 1.2 - 0.1)).contiguous()
    for gain in (1.0, 200.0):
        net, sd = gpu["nets"][gain], gpu["sds"][gain]
        ours = net.resd(x0[None].cuda(), gpu["gbatch"])[0].cpu()
        ref = O.deformer(sd, x0, frame["tuv"][0], tb, frame["frame_dim"])
        err = (ours - ref).abs().max().item()
        diag("deformer", gain=gain, max_err=err)
        assert err < 2e-6, err


def test_part_mlp(gpu):
    g = torch.Generator().manual_seed(13)
    net, sd, frame = gpu["nets"][200.0], gpu["sds"][200.0], gpu["frame"]
    eng = net.engine()
    lat = int(frame["latent_index"][0])
    for pid in range(5):
        n = 1000 + 37 * pid
        e = torch.cat([torch.rand(n, 3, generator=g), torch.randn(n, 16, generator=g)], -1)
        v = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        ours = eng.part_mlp(pid, e.cuda(), v.cuda(), gpu["gbatch"]).cpu()
        pre = f"tpose_human.part_networks.{pid}."
        h = O.mlp_softplus(e, sd, pre + "occ.linears.")
        occ = 1 - torch.exp(-torch.nn.functional.softplus(h[..., :1]))
        inp = torch.cat([e, O.posenc(v), h[..., 1:], sd[pre + "rgb_latent"][lat][None].expand(n, -1)], -1)
        rgb = O.mlp_softplus(inp, sd, pre + "rgb.linears.").sigmoid()
        ref = torch.cat([rgb, occ], -1)
        err = (ours - ref).abs().max().item()
        diag("part_mlp", part=pid, max_err=err)
        assert err < 5e-6, (pid, err)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_part_mlp_tensor_core(gpu, mode):
    """The tcgen05 3xTF32 MLP kernel (mode 1: one epilogue warpgroup per tile slot, mode 2: two) against the fp32
    oracle (and against our fp32 FFMA kernel)."""
    from instant_nvr_b200.engine import Engine
    g = torch.Generator().manual_seed(14)
    net, sd, frame = gpu["nets"][200.0], gpu["sds"][200.0], gpu["frame"]
    eng_tc = Engine(gpu["cfg"], mlp_mode=mode)
    eng_tc.bind_params(net)
    eng_ff = Engine(gpu["cfg"], mlp_mode=0)
    eng_ff.bind_params(net)
    lat = int(frame["latent_index"][0])
    for pid in range(5):
        for n in (1, 127, 128, 129, 20000 + 37 * pid):
            e = torch.cat([torch.rand(n, 3, generator=g), torch.randn(n, 16, generator=g)], -1)
            v = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
            ours = eng_tc.part_mlp(pid, e.cuda(), v.cuda(), gpu["gbatch"]).cpu()
            ffma = eng_ff.part_mlp(pid, e.cuda(), v.cuda(), gpu["gbatch"]).cpu()
            pre = f"tpose_human.part_networks.{pid}."
            h = O.mlp_softplus(e, sd, pre + "occ.linears.")
            occ = 1 - torch.exp(-torch.nn.functional.softplus(h[..., :1]))
            inp = torch.cat([e, O.posenc(v), h[..., 1:], sd[pre + "rgb_latent"][lat][None].expand(n, -1)], -1)
            rgb = O.mlp_softplus(inp, sd, pre + "rgb.linears.").sigmoid()
            ref = torch.cat([rgb, occ], -1)
            err = (ours - ref).abs().max().item()
            err_ff = (ours - ffma).abs().max().item()
            diag("part_mlp_tc", mode=mode, part=pid, n=n, max_err_vs_oracle=err, max_err_vs_ffma=err_ff)
            assert err < 2e-5, (pid, n, err)          # north-star bound is 1e-4; the fp32 FFMA kernel holds 5e-6


def _points(gpu, n_rays_side=24, S=24):
    from instant_nvr_b200.synthetic import make_rays
    rays = make_rays(gpu["frame"], n_rays_side, n_rays_side)
    b = O.strip_batch({**gpu["frame"], **rays})
    pts, _ = O.sample_along_rays(b["ray_o"], b["ray_d"], b["near"], b["far"], S)
    vd = b["ray_d"][:, None].expand(-1, S, 3).reshape(-1, 3).contiguous()
    return pts.reshape(-1, 3).contiguous(), vd, b


def test_warp_stage(gpu):
    """cull + KNN + LBS + deformer taps against the oracle's stages."""
    wpts, vd, b = _points(gpu)
    net, sd, cfg = gpu["nets"][200.0], gpu["sds"][200.0], gpu["cfg"]
    raw, surv, warp = net.engine().query_points_debug(wpts.cuda(), vd.cuda(), gpu["gbatch"])
    surv, warp = surv.cpu(), warp.cpu()
    _, _, st = O.network_forward(sd, wpts, vd, b, cfg.smpl_thresh, want_stages=True)
    keep_ref = torch.zeros(wpts.shape[0], dtype=torch.bool)
    keep_ref[st["pind"]] = True
    border = (st["pnorm"] - cfg.smpl_thresh).abs() < 2e-6
    cull_mismatch = ((surv >= 0) != keep_ref) & ~border
    pind = st["pind"]
    w = warp[pind]                                          # (M,5,8)
    flag_ref = st["flag"]
    fborder = (st["pdist"] - cfg.smpl_thresh).abs() < 2e-6
    flag_mismatch = ((w[..., 0] > 0.5) != flag_ref) & ~fborder
    both = (w[..., 0] > 0.5) & flag_ref
    pd_err = (w[..., 7] - st["pdist"]).abs().max().item()
    scale = st["tpose"].abs().clamp(min=1.0)
    x_err = ((w[..., 1:4] - st["tpose"]).abs() / scale)[both].max().item()
    v_err = ((w[..., 4:7] - st["tdirs"]).abs() / st["tdirs"].abs().clamp(min=1.0))[both].max().item()
    diag("warp_stage", n=int(wpts.shape[0]), survivors=int(keep_ref.sum()), flagged=int(flag_ref.sum()),
         cull_mismatch=int(cull_mismatch.sum()), flag_mismatch=int(flag_mismatch.sum()), pdist_err=pd_err, x_relerr=x_err,
         v_relerr=v_err)
    assert cull_mismatch.sum() == 0 and flag_mismatch.sum() == 0
    assert pd_err < 1e-6 and x_err < 2e-5 and v_err < 2e-5


def _compare_raw(name, ours, ref, st, sd, strict_all):
    """ours/ref (N,4).  Returns stats; asserts the tolerance policy in the module docstring."""
    err = (ours - ref).abs() / ref.abs().clamp(min=1.0)
    e = err.max(-1).values
    pind = st["pind"]
    # conditioning: was any flagged part of the sample evaluated outside its bbox?
    outside = torch.zeros(ours.shape[0], dtype=torch.bool)
    M = pind.shape[0]
    out_part = torch.zeros(M, 5, dtype=torch.bool)
    for pid in range(5):
        b = sd[f"tpose_human.part_networks.{pid}.embedder.bounds"]
        out_part[:, pid] = st["flag"][:, pid] & ~_inside(st["tpose"][:, pid], b)
    outside[pind] = out_part.any(-1)
    n_bad_strict = int((e[~outside] > 1e-4).sum())
    stats = dict(n=int(ours.shape[0]), active=int(M), outside=int(outside.sum()), max_err_inside=float(e[~outside].max()),
                 max_err_outside=float(e[outside].max()) if outside.any() else 0.0, frac_1e4=float((e <= 1e-4).float().mean()),
                 n_bad_inside=n_bad_strict)
    diag(name, **stats)
    assert n_bad_strict == 0, stats
    if strict_all:
        assert float(e.max()) <= 1e-4, stats
    return stats


@pytest.mark.parametrize("gain", [1.0, 200.0])
def test_query_points_vs_oracle(gpu, gain):
    wpts, vd, b = _points(gpu)
    net, sd, cfg = gpu["nets"][gain], gpu["sds"][gain], gpu["cfg"]
    out = net(wpts.cuda(), vd.cuda(), None, gpu["gbatch"])
    raw, occ = out["raw"][0].cpu(), out["occ"][0].cpu()
    rraw, rocc, st = O.network_forward(sd, wpts, vd, b, cfg.smpl_thresh, want_stages=True)
    assert raw.shape == rraw.shape and occ.shape == rocc.shape
    assert torch.equal(occ[:, 0], raw[:, 3])
    _compare_raw(f"query_points_gain{int(gain)}", raw, rraw, st, sd, strict_all=(gain == 1.0))


@pytest.mark.parametrize("gain", [1.0, 200.0])
def test_render_vs_reference_golden(gpu, gain):
    """Renderer.render against the fixtures recorded from the reference's own code."""
    from instant_nvr_b200.renderer import Renderer
    gold = load_golden(f"e2e_gain{int(gain)}.npz")
    net, sd, cfg = gpu["nets"][gain], gpu["sds"][gain], gpu["cfg"]
    ret = Renderer(net, return_raw=True).render(dict(gpu["gbatch"]))
    assert ret["rgb_map"].device.type == "cpu" and ret["rgb_map"].shape == gold["rgb_map"].shape
    graw = torch.from_numpy(gold["raw"])[0]
    b = O.strip_batch(gpu["batch"])
    pts, _ = O.sample_along_rays(b["ray_o"], b["ray_d"], b["near"], b["far"], cfg.N_samples)
    vd = b["ray_d"][:, None].expand(-1, cfg.N_samples, 3).reshape(-1, 3)
    _, _, st = O.network_forward(sd, pts.reshape(-1, 3), vd, b, cfg.smpl_thresh, want_stages=True)
    _compare_raw(f"render_raw_gain{int(gain)}", ret["raw"][0], graw, st, sd, strict_all=(gain == 1.0))
    g_rgb, g_acc = torch.from_numpy(gold["rgb_map"]), torch.from_numpy(gold["acc_map"])
    rgb_err = (ret["rgb_map"] - g_rgb).abs().max().item()
    acc_err = (ret["acc_map"] - g_acc).abs().max().item()
    psnr_vs_ref = O.psnr(ret["rgb_map"], g_rgb)
    # PSNR-within-0.1-dB against an arbitrary ground truth: use reference + noise as the stand-in image
    gt = (g_rgb + 0.05 * torch.randn(g_rgb.shape, generator=torch.Generator().manual_seed(0))).clamp(0, 1)
    d_psnr = abs(O.psnr(ret["rgb_map"], gt) - O.psnr(g_rgb, gt))
    diag(f"render_maps_gain{int(gain)}", rgb_err=rgb_err, acc_err=acc_err, psnr_vs_reference=psnr_vs_ref, dpsnr=d_psnr)
    assert d_psnr < 0.1
    if gain == 1.0:
        assert rgb_err < 1e-4 and acc_err < 1e-4 and psnr_vs_ref > 80.0


def test_render_host_and_multipass_match(gpu):
    """Host-buffer entry point == device entry point, and splitting into passes changes nothing."""
    cfg, net = gpu["cfg"], gpu["nets"][200.0]
    eng = net.engine()
    gb = gpu["gbatch"]
    S = cfg.N_samples
    rgb, acc, raw = eng.render_rays(gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0], S, batch=gb, want_raw=True)
    R = rgb.shape[0]
    o, d, n, f = (gpu["rays"][k][0].contiguous().pin_memory() for k in ("ray_o", "ray_d", "near", "far"))
    rgb_h, acc_h = torch.empty(R, 3).pin_memory(), torch.empty(R).pin_memory()
    eng.render_rays_host(o, d, n, f, S, rgb_h, acc_h)
    assert torch.equal(rgb_h, rgb.cpu()) and torch.equal(acc_h, acc.cpu())
    keep = eng.max_points_per_pass
    try:
        eng.max_points_per_pass = 37 * S                    # 37 rays per pass -> 28 passes, ragged tail
        eng._ws = None
        rgb2, acc2, raw2 = eng.render_rays(gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0], S, want_raw=True)
    finally:
        eng.max_points_per_pass = keep
        eng._ws = None
    assert torch.equal(raw2, raw) and torch.equal(rgb2, rgb) and torch.equal(acc2, acc)
    c = eng.counters()
    diag("counters_last_pass", **{k: v for k, v in c.items()})


def test_two_lane_render_matches_one_lane(gpu, monkeypatch):
    """A render call of >= 2^19 samples runs as two lanes (two passes on two streams of the engine, half a workspace each); here
    the threshold is lowered so the small fixture takes that path: same bits as the one-lane render, device and host entry points,
    the counters are the two passes' sums and the gather footprint is the union over both halves."""
    from instant_nvr_b200.engine import Engine
    cfg, net = gpu["cfg"], gpu["nets"][200.0]
    gb = gpu["gbatch"]
    S = cfg.N_samples
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    eng1 = Engine(cfg, tune=2048)                                # NVR_TUNE_ONE_LANE
    eng1.bind_params(net)
    rgb1, acc1, raw1 = eng1.render_rays(o, d, n, f, S, batch=gb, want_raw=True)
    c1, fp1 = eng1.counters(), eng1.gather_footprint()
    monkeypatch.setenv("NVR_TWO_LANE_MIN_SAMPLES", "1024")
    eng2 = Engine(cfg)
    monkeypatch.delenv("NVR_TWO_LANE_MIN_SAMPLES")
    eng2.bind_params(net)
    rgb2, acc2, raw2 = eng2.render_rays(o, d, n, f, S, batch=gb, want_raw=True)
    c2, fp2 = eng2.counters(), eng2.gather_footprint()
    assert c1["n_passes"] == 1 and c2["n_passes"] == 2
    assert torch.equal(raw2, raw1) and torch.equal(rgb2, rgb1) and torch.equal(acc2, acc1)
    assert c2["n_survivors"] == c1["n_survivors"] and c2["n_far_pairs"] == c1["n_far_pairs"] and fp2 == fp1
    assert [a - 2 for a in c2["n_pairs"]] == [a - 1 for a in c1["n_pairs"]]
    host = [t.cpu().contiguous().pin_memory() for t in (o, d, n, f)]
    rgb_h, acc_h = torch.empty(o.shape[0], 3).pin_memory(), torch.empty(o.shape[0]).pin_memory()
    eng2.render_rays_host(*host, S, rgb_h, acc_h)
    assert torch.equal(rgb_h, rgb1.cpu()) and torch.equal(acc_h, acc1.cpu())
    # odd ray counts: the second lane takes the shorter half, a single ray stays one pass
    for k in (1, 2, 3, 65):
        a, b = eng2.render_rays(o[:k], d[:k], n[:k], f[:k], S)
        assert torch.equal(a, rgb1[:k]) and torch.equal(b, acc1[:k])


@pytest.mark.parametrize("gain,thresh", [(1.0, None), (200.0, None), (200.0, 0.1), (200.0, 0.02)])
def test_far_field_pairs_share_one_evaluation(gpu, gain, thresh):
    """Default mode answers every flagged pair whose Gaussian weights sum to < 1e-20 (part farther than ~0.73 m) with ONE
    shared zero-weight evaluation per part (csrc/nvr_kernels.cuh NVR_FAR_WSUM); NVR_TUNE_NO_FAR_COLLAPSE evaluates each
    pair on its own, as the reference does.  The canonical points differ by < 5e-13 m, which fp32 absorbs: the two
    modes must agree to 1e-6 on every sample and bit for bit on (nearly) all of them; the pair accounting must add up."""
    from instant_nvr_b200.engine import Engine
    cfg, gb, net = gpu["cfg"], gpu["gbatch"], gpu["nets"][gain]
    if thresh is not None:                                    # cfg.smpl_thresh of other configs (0.1) and a tight one: the short cuts scale with it
        cfg = cfg.with_(smpl_thresh=thresh)
    S = cfg.N_samples
    out = {}
    for tune in (0, 8, 12):                                   # 12: also no quick / early-out cull (every sample through the exact lookup)
        eng = Engine(cfg, tune=tune)
        eng.bind_params(net)
        rgb, acc, raw = eng.render_rays(gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0], S, want_raw=True, batch=gb)
        out[tune] = (rgb, acc, raw, eng.counters())
    (rgb0, acc0, raw0, c0), (rgb8, acc8, raw8, c8) = out[0], out[8]
    assert torch.equal(out[12][2], raw8) and out[12][3]["n_survivors"] == c8["n_survivors"] and out[12][3]["n_pairs"] == c8["n_pairs"]
    assert sum(c8["n_far_pairs"]) == 0 and sum(c0["n_far_pairs"]) > 0
    assert c0["n_survivors"] == c8["n_survivors"] > 0
    for p in range(5):
        # one shared far-field pair per part and PASS (a render of this size runs as two lanes = two passes)
        assert c8["n_pairs"][p] == c0["n_pairs"][p] - c0["n_passes"] + c0["n_far_pairs"][p], p
    d = (raw0 - raw8).abs()
    differ = int((d.max(dim=-1).values > 0).sum())
    active = int((raw8[..., 3] > 0).sum())
    diag("far_collapse", gain=gain, thresh=cfg.smpl_thresh, far_pairs=c0["n_far_pairs"], evaluated=c0["n_pairs"], flagged=c8["n_pairs"],
         max_abs=d.max().item(), samples_not_bitwise_equal=differ, active=active,
         rgb_max_abs=(rgb0 - rgb8).abs().max().item())
    assert d.max().item() <= 1e-6 and differ <= max(2, active // 1000)
    assert (rgb0 - rgb8).abs().max().item() <= 1e-6 and (acc0 - acc8).abs().max().item() <= 1e-6


def test_edge_cases(gpu):
    cfg, net, gb = gpu["cfg"], gpu["nets"][1.0], gpu["gbatch"]
    eng = net.engine()
    eng.bind_frame(gb)
    z3, z1 = torch.zeros(0, 3, device="cuda"), torch.zeros(0, device="cuda")
    rgb, acc = eng.render_rays(z3, z3, z1, z1, cfg.N_samples)
    assert rgb.shape == (0, 3) and acc.shape == (0,)
    raw, occ = eng.query_points(z3, z3, gb)
    assert raw.shape == (0, 4)
    # rays that miss the body entirely -> exact zeros
    far_o = torch.tensor([[50.0, 50.0, 50.0]] * 33, device="cuda")
    dirs = torch.nn.functional.normalize(torch.tensor([[1.0, 0.2, 0.1]] * 33, device="cuda"), dim=-1)
    rgb, acc, raw = eng.render_rays(far_o, dirs, torch.full((33,), 1.0, device="cuda"), torch.full((33,), 2.0, device="cuda"),
                                    cfg.N_samples, want_raw=True)
    assert rgb.abs().max() == 0 and acc.abs().max() == 0 and raw.abs().max() == 0
    # single ray, single sample
    r1 = eng.render_rays(gb["ray_o"][0][500:501], gb["ray_d"][0][500:501], gb["near"][0][500:501], gb["far"][0][500:501], 1)
    assert r1[0].shape == (1, 3)
    # NaN coordinates must not crash or poison neighbours
    pts = torch.cat([torch.full((1, 3), float("nan"), device="cuda"), gb["wpts"][0][:7] if "wpts" in gb else torch.zeros(7, 3, device="cuda")])
    raw, _ = eng.query_points(pts, torch.nn.functional.normalize(torch.ones_like(pts), dim=-1), gb)
    torch.cuda.synchronize()
    assert torch.isfinite(raw[1:]).all()


def test_no_cpu_path(gpu):
    from instant_nvr_b200.network import Network
    cpu_net = Network(gpu["cfg"], device="cpu").eval()
    with pytest.raises(RuntimeError):
        cpu_net(torch.zeros(4, 3), torch.zeros(4, 3), None, gpu["batch"])     # no CPU path


@pytest.fixture(scope="module")
def full(golden_setup):
    """The shipped inb_377 config (1.14 GB of tables), reference-init magnitudes x 50."""
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import fill_weights
    cfg = PathConfig.inb_377(N_samples=32)
    frame = golden_setup["frame"]
    t0 = time.time()
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=3, table_gain=50.0, bounds=frame["bounds"][0])
    sd = net.state_dict()
    gnet = Network(cfg, device="cpu")
    gnet.load_state_dict(sd)
    gnet = gnet.cuda().eval()
    print(f"[full] built + filled in {time.time() - t0:.1f}s")
    return dict(cfg=cfg, frame=frame, sd=sd, net=gnet)


def test_full_config_c1(full):
    """Config 1 shape (64x64 rays x 32 samples) with the shipped table sizes (T = 1048583, res 2005:
    int64 hash products past 2^32) against the oracle."""
    from instant_nvr_b200.renderer import Renderer
    from instant_nvr_b200.synthetic import make_rays
    cfg, frame, sd, net = full["cfg"], full["frame"], full["sd"], full["net"]
    rays = make_rays(frame, 64, 64)
    batch = {**frame, **rays}
    ret = Renderer(net, return_raw=True).render(to_cuda(batch))
    ref = O.render(sd, batch, cfg.N_samples, cfg.smpl_thresh)
    b = O.strip_batch(batch)
    pts, _ = O.sample_along_rays(b["ray_o"], b["ray_d"], b["near"], b["far"], cfg.N_samples)
    vd = b["ray_d"][:, None].expand(-1, cfg.N_samples, 3).reshape(-1, 3)
    _, _, st = O.network_forward(sd, pts.reshape(-1, 3), vd, b, cfg.smpl_thresh, want_stages=True)
    _compare_raw("full_c1_raw", ret["raw"][0], ref["raw"][0], st, sd, strict_all=False)
    psnr = O.psnr(ret["rgb_map"], ref["rgb_map"])
    diag("full_c1_maps", rgb_err=(ret["rgb_map"] - ref["rgb_map"]).abs().max().item(), psnr_vs_oracle=psnr)
    assert psnr > 50.0


def test_full_size_properties_c2(full):
    """Config 2 size (512x512 rays x 128 samples = 33.5 M ray-samples), too big for the oracle:
    size-independent properties instead."""
    from instant_nvr_b200.synthetic import make_rays
    cfg, frame, net = full["cfg"], full["frame"], full["net"]
    S = 128
    eng = net.engine()
    rays = make_rays(frame, 512, 512)
    gb = to_cuda({**frame, **rays})
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    torch.cuda.synchronize()
    t0 = time.time()
    rgb, acc = eng.render_rays(o, d, n, f, S, batch=gb)
    torch.cuda.synchronize()
    dt = time.time() - t0
    c = eng.counters()
    diag("c2_render", seconds=dt, ray_samples=512 * 512 * S, **c)
    assert torch.isfinite(rgb).all() and torch.isfinite(acc).all()
    assert acc.min() >= 0 and acc.max() <= 1 + 1e-5 and rgb.min() >= 0 and rgb.max() <= 1 + 1e-5
    # (0) the frame went through the pipeline as two lanes (two passes on two streams, half a workspace each): the one-lane
    #     render (NVR_TUNE_ONE_LANE) and the host entry point (each lane copies its own rays in and pixels out) give the same bits
    from instant_nvr_b200.engine import Engine
    assert c["n_passes"] == 2
    fp2 = eng.gather_footprint()                              # distinct table sectors over BOTH lanes' pair lists
    eng1 = Engine(cfg, tune=2048)
    eng1.bind_params(net)
    rgb_1, acc_1 = eng1.render_rays(o, d, n, f, S, batch=gb)
    assert eng1.counters()["n_passes"] == 1
    assert eng1.gather_footprint() == fp2 and min(fp2) > 0
    assert torch.equal(rgb_1, rgb) and torch.equal(acc_1, acc)
    c1 = eng1.counters()
    assert c1["n_survivors"] == c["n_survivors"] and c1["n_far_pairs"] == c["n_far_pairs"]
    assert [a - 1 for a in c1["n_pairs"]] == [a - 2 for a in c["n_pairs"]]        # one shared far-field pair per part and pass
    del eng1
    host = [t.cpu().contiguous().pin_memory() for t in (o, d, n, f)]
    rgb_h, acc_h = torch.empty(o.shape[0], 3).pin_memory(), torch.empty(o.shape[0]).pin_memory()
    eng.render_rays_host(*host, S, rgb_h, acc_h)
    assert torch.equal(rgb_h, rgb.cpu()) and torch.equal(acc_h, acc.cpu())
    # (1) ray-permutation equivariance: rendering a shuffled subset gives the same pixels (bit-exact:
    #     per-ray results do not depend on neighbours or on compaction order)
    perm = torch.randperm(512 * 512, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))[:20000]
    rgb_p, acc_p, raw_p = eng.render_rays(o[perm], d[perm], n[perm], f[perm], S, want_raw=True)
    c_perm = eng.counters()
    assert torch.equal(rgb_p, rgb[perm]) and torch.equal(acc_p, acc[perm])
    # (2) the two entry points agree: querying the same sample points gives the same raw
    pts = (o[perm][:512, None] + d[perm][:512, None] * (n[perm][:512, None, None] * (1 - torch.linspace(0, 1, S, device="cuda"))[None, :, None]
                                                       + f[perm][:512, None, None] * torch.linspace(0, 1, S, device="cuda")[None, :, None]))
    vd = d[perm][:512, None].expand(-1, S, 3).reshape(-1, 3)
    raw_q, _ = eng.query_points(pts.reshape(-1, 3), vd, gb)
    diff = (raw_q.view(512, S, 4) - raw_p.view(-1, S, 4)[:512]).abs()
    diag("c2_entrypoint_consistency", max_diff=diff.max().item(), frac_exact=float((diff == 0).float().mean()))
    assert float((diff.max(-1).values < 1e-3).float().mean()) > 0.999
    # (3) compositing identity: acc == 1 - prod(1 - alpha) along each ray
    alpha = raw_p.view(-1, S, 4)[..., 3].double()
    acc_ref = 1 - torch.prod(1 - alpha, dim=-1)
    assert (acc_p.double() - acc_ref).abs().max() < 1e-5
    # (4) the KNN short cuts (far-field pairs share one evaluation, certainly-unflagged parts skip the search) against
    #     the every-pair evaluation (NVR_TUNE_NO_FAR_COLLAPSE) on the same 2.56 M samples, full-size tables
    from instant_nvr_b200.engine import Engine
    c0 = c_perm
    eng8 = Engine(cfg, tune=12)                               # every pair on its own, every sample through the exact cull lookup
    eng8.bind_params(net)
    rgb_8, acc_8, raw_8 = eng8.render_rays(o[perm], d[perm], n[perm], f[perm], S, want_raw=True, batch=gb)
    c8 = eng8.counters()
    dr = (raw_8 - raw_p).abs()
    differ = int((dr.max(dim=-1).values > 0).sum())
    diag("c2_far_collapse", far_pairs=c0["n_far_pairs"], evaluated=c0["n_pairs"], flagged=c8["n_pairs"], max_abs=dr.max().item(),
         samples_not_bitwise_equal=differ, survivors=c8["n_survivors"])
    assert c0["n_survivors"] == c8["n_survivors"]
    for p in range(5):
        # one shared far-field pair per part and PASS (a render of this size runs as two lanes = two passes)
        assert c8["n_pairs"][p] == c0["n_pairs"][p] - c0["n_passes"] + c0["n_far_pairs"][p], p
    assert dr.max().item() <= 1e-6 and differ <= max(2, c8["n_survivors"] // 10000)
    assert (rgb_8 - rgb_p).abs().max().item() <= 1e-6


def test_full_size_properties_c4_c5(full):
    """BASELINE configs[3] (1024x1024 rays x 64 samples, ray tiles dealt to 8 ranks) and configs[4]
    (3840x2160 rays x 256 samples = 2.1 G ray-samples, hundreds of passes): size-independent properties."""
    from instant_nvr_b200.sharding import shard_indices
    from instant_nvr_b200.synthetic import make_rays
    cfg, frame, net = full["cfg"], full["frame"], full["net"]
    eng = net.engine()
    gen = torch.Generator(device="cuda").manual_seed(1)
    # ---- C4: the union of the 8 interleaved shards reproduces the single-pass frame bit for bit
    rays = make_rays(frame, 1024, 1024)
    gb = to_cuda({**frame, **rays})
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    rgb, acc = eng.render_rays(o, d, n, f, 64, batch=gb)
    assert torch.isfinite(rgb).all() and acc.max() > 0.01
    out = torch.empty_like(rgb)
    for r in range(8):
        idx = shard_indices(1024 * 1024, r, 8).cuda()
        out[idx] = eng.render_rays(o[idx], d[idx], n[idx], f[idx], 64)[0]
    assert torch.equal(out, rgb)
    # ---- C5: 4K x 256 samples
    rays = make_rays(frame, 2160, 3840)
    gb = to_cuda({**frame, **rays})
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    torch.cuda.synchronize()
    t0 = time.time()
    rgb, acc = eng.render_rays(o, d, n, f, 256, batch=gb)
    torch.cuda.synchronize()
    dt = time.time() - t0
    diag("c5_render", seconds=dt, ray_samples=2160 * 3840 * 256, g_ray_samples_per_s=2160 * 3840 * 256 / dt / 1e9)
    assert torch.isfinite(rgb).all() and torch.isfinite(acc).all() and acc.min() >= 0 and acc.max() <= 1 + 1e-5
    perm = torch.randperm(2160 * 3840, device="cuda", generator=gen)[:4096]
    rgb_p, acc_p, raw_p = eng.render_rays(o[perm], d[perm], n[perm], f[perm], 256, want_raw=True)
    assert torch.equal(rgb_p, rgb[perm]) and torch.equal(acc_p, acc[perm])
    alpha = raw_p.view(-1, 256, 4)[..., 3].double()
    assert (acc_p.double() - (1 - torch.prod(1 - alpha, dim=-1))).abs().max() < 1e-5


@pytest.fixture(scope="module")
def full_gain1(full):
    """The same shipped-size network with the reference's OWN init magnitudes (table gain 1): the strict 1e-4 bound applies to
    every sample."""
    from instant_nvr_b200.network import Network
    sd1 = {k: (v / 50.0 if k.endswith((".embedder.dense", ".embedder.hash")) and "part_networks" in k else v.clone())
           for k, v in full["sd"].items()}
    gnet = Network(full["cfg"], device="cpu")
    gnet.load_state_dict(sd1)
    return dict(cfg=full["cfg"], frame=full["frame"], sd=sd1, net=gnet.cuda().eval())


CONFIG_SHAPES = {"c2": (512, 512, 128), "c4": (1024, 1024, 64), "c5": (2160, 3840, 256)}


@pytest.mark.parametrize("name", ["c2", "c4", "c5"])
@pytest.mark.parametrize("gain", [50.0, 1.0])
def test_baseline_configs_ray_subset_vs_oracle(full, full_gain1, name, gain):
    """BASELINE.json configs[1] / [3] / [4] at their FULL sizes and the shipped table sizes: a random 4096-ray subset of the
    frame's (bbox-hitting) rays, rendered (a) inside the whole-frame render and (b) on its own, against the oracle --
    per-sample raw, rgb_map, acc_map within 1e-4 (strict for every sample at the reference's init magnitudes, gain 1; at
    gain 50 samples evaluated outside a part's bbox carry the reference's own extrapolation noise and are counted)."""
    from instant_nvr_b200.synthetic import make_rays
    fx = full if gain == 50.0 else full_gain1
    cfg, frame, sd, net = fx["cfg"], fx["frame"], fx["sd"], fx["net"]
    H, W, S = CONFIG_SHAPES[name]
    eng = net.engine()
    rays = make_rays(frame, H, W, drop_missing=True)
    gb = to_cuda({**frame, **rays})
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    R = o.shape[0]
    sel = torch.randperm(R, generator=torch.Generator().manual_seed(7))[:4096].sort().values
    t0 = time.time()
    rgb_full, acc_full = eng.render_rays(o, d, n, f, S, batch=gb)                       # the whole frame
    torch.cuda.synchronize()
    t_full = time.time() - t0
    sg = sel.cuda()
    rgb_s, acc_s, raw_s = eng.render_rays(o[sg], d[sg], n[sg], f[sg], S, want_raw=True)
    assert torch.equal(rgb_s, rgb_full[sg]) and torch.equal(acc_s, acc_full[sg])       # the subset IS the frame's pixels
    sub = {**frame, **{k: rays[k][:, sel] for k in ("ray_o", "ray_d", "near", "far")}}
    b = O.strip_batch(sub)
    pts, _ = O.sample_along_rays(b["ray_o"], b["ray_d"], b["near"], b["far"], S)
    vd = b["ray_d"][:, None].expand(-1, S, 3).reshape(-1, 3)
    t0 = time.time()
    with torch.no_grad():
        raw_o, _, st = O.network_forward(sd, pts.reshape(-1, 3), vd, b, cfg.smpl_thresh, want_stages=True)
        _, rgb_o, acc_o = O.composite(raw_o.reshape(-1, S, 4))
    t_oracle = time.time() - t0
    # decision flips (module docstring): a sample whose cull distance / a part distance sits within 2e-6 of smpl_thresh may land
    # on the other side under FMA contraction; such samples (and their rays, for the maps) are excluded and counted
    edge = (st["pnorm"].reshape(-1) - cfg.smpl_thresh).abs() < 2e-6
    edge[st["pind"]] |= ((st["pdist"] - cfg.smpl_thresh).abs() < 2e-6).any(-1)
    n_edge = int(edge.sum())
    assert n_edge <= max(8, edge.numel() // 50000), n_edge
    ours_raw, ref_raw = raw_s.cpu().clone(), raw_o.clone()
    ours_raw[edge] = 0
    ref_raw[edge] = 0
    stats = _compare_raw(f"{name}_subset_raw_gain{int(gain)}", ours_raw, ref_raw, st, sd, strict_all=(gain == 1.0))
    ray_ok = ~edge.reshape(-1, S).any(-1)
    rgb_err = (rgb_s.cpu() - rgb_o)[ray_ok].abs().max().item()
    acc_err = (acc_s.cpu() - acc_o)[ray_ok].abs().max().item()
    psnr = O.psnr(rgb_s.cpu()[ray_ok], rgb_o[ray_ok])
    diag(f"{name}_subset_maps_gain{int(gain)}", rays=R, subset=int(sel.numel()), samples=int(sel.numel()) * S, threshold_edge_samples=n_edge, rgb_err=rgb_err, acc_err=acc_err,
         psnr_vs_oracle=psnr, frame_seconds=t_full, oracle_seconds=t_oracle, active=stats["active"])
    assert stats["active"] > 1000
    if gain == 1.0:
        assert rgb_err < 1e-4 and acc_err < 1e-4
    assert psnr > 60.0


def test_dense_a1_variant_is_dense(gpu):
    """NVR_TUNE_DENSE_A1 (bench.py's `dense_a1` line): every sample survives the cull and is flagged in exactly one
    part, so evaluated pairs == samples."""
    from instant_nvr_b200.engine import Engine
    cfg, gb, net = gpu["cfg"], gpu["gbatch"], gpu["nets"][1.0]
    eng = Engine(cfg, tune=16)
    eng.bind_params(net)
    S = cfg.N_samples
    rgb, acc, raw = eng.render_rays(gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0], S, want_raw=True, batch=gb)
    c = eng.counters()
    n = gb["ray_o"].shape[1] * S
    assert c["n_survivors"] == n and sum(c["n_pairs"]) == n and sum(c["n_far_pairs"]) == 0
    assert torch.isfinite(raw).all() and torch.isfinite(rgb).all()
    uniq = eng.gather_footprint()
    assert all(u >= 0 for u in uniq) and sum(uniq) > 0


def test_gather_footprint_counts_distinct_sectors(gpu):
    """nvr_gather_footprint against a direct count with the oracle's row arithmetic: distinct (row, half) sectors the flagged
    pairs of a render touch, per part."""
    from instant_nvr_b200.engine import Engine
    cfg, gb, net, sd = gpu["cfg"], gpu["gbatch"], gpu["nets"][200.0], gpu["sds"][200.0]
    eng = Engine(cfg, tune=8)                                 # every pair on its own: the pair lists are the reference's flagged pairs
    eng.bind_params(net)
    wpts, vd, b = _points(gpu)
    eng.query_points(wpts.cuda(), vd.cuda(), gb)
    uniq = eng.gather_footprint()
    _, _, st = O.network_forward(sd, wpts, vd, b, cfg.smpl_thresh, want_stages=True)
    for pid in range(5):
        x = st["tpose"][:, pid][st["flag"][:, pid]]
        pre = f"tpose_human.part_networks.{pid}.embedder."
        rows = O.grid_rows(sd, pre, x)
        n_ref = 2 * int(torch.unique(rows).numel())
        diag("gather_footprint", part=pid, pairs=int(x.shape[0]), unique_sectors=int(uniq[pid]), oracle=n_ref)
        assert abs(uniq[pid] - n_ref) <= max(4, n_ref // 500), (pid, uniq[pid], n_ref)    # threshold-edge pairs may differ


def test_peer_frame_single_rank(gpu):
    """nvr_render_rays_frame at world == 1 (the fused compositing-kernel store + barrier code path on a one-GPU box): the frame is
    the plain render, for ragged ray counts, several passes and consecutive frames (alternating slots)."""
    from instant_nvr_b200.sharding import PeerFrame
    cfg, net, gb = gpu["cfg"], gpu["nets"][200.0], gpu["gbatch"]
    eng = net.engine()
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    S = cfg.N_samples
    rgb, acc = eng.render_rays(o, d, n, f, S, batch=gb)
    R = o.shape[0]
    pf = PeerFrame(eng, R, 0, 1, tile=64)
    try:
        first = pf.render(o, d, n, f, S, batch=gb)
        a = first.clone()
        second, l_rgb, l_acc = pf.render(o, d, n, f, S, want_local=True)
        assert second.data_ptr() != first.data_ptr()
        for fr in (a, second):
            assert torch.equal(fr[:, :3], rgb) and torch.equal(fr[:, 3], acc)
        assert torch.equal(l_rgb, rgb) and torch.equal(l_acc, acc)
        keep = eng.max_points_per_pass
        try:
            eng.max_points_per_pass = 37 * S
            eng._ws = None
            third = pf.render(o, d, n, f, S)
            assert torch.equal(third[:, :3], rgb) and torch.equal(third[:, 3], acc)
        finally:
            eng.max_points_per_pass = keep
            eng._ws = None
        g = pf.allgather(rgb, acc)
        assert torch.equal(g[:, :3], rgb) and torch.equal(g[:, 3], acc)
        # host-buffer form (nvr_render_rays_frame_host)
        host = [t.cpu().contiguous().pin_memory() for t in (o, d, n, f)]
        rgb_h, acc_h = torch.empty(R, 3).pin_memory(), torch.empty(R).pin_memory()
        fourth = pf.render_host(*host, S, rgb_h, acc_h)
        assert torch.equal(fourth[:, :3], rgb) and torch.equal(fourth[:, 3], acc)
        assert torch.equal(rgb_h, rgb.cpu()) and torch.equal(acc_h, acc.cpu())
    finally:
        pf.close()


def test_inference_tables_match_full_tables(gpu):
    """Opt-in pre-summed inference tables: same render within fp32 re-association, and they follow in-place updates."""
    from instant_nvr_b200.engine import Engine
    cfg, net, gb = gpu["cfg"], gpu["nets"][200.0], gpu["gbatch"]
    eng_full = Engine(cfg, inference_tables=False)
    eng_full.bind_params(net)
    eng_sum = Engine(cfg, inference_tables=True)
    eng_sum.bind_params(net)
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    rgb_a, acc_a, raw_a = eng_full.render_rays(o, d, n, f, cfg.N_samples, batch=gb, want_raw=True)
    rgb_b, acc_b, raw_b = eng_sum.render_rays(o, d, n, f, cfg.N_samples, batch=gb, want_raw=True)
    err = (raw_a - raw_b).abs().max().item()
    diag("inference_tables", max_raw_diff=err, psnr=O.psnr(rgb_b.cpu(), rgb_a.cpu()))
    assert err < 2e-4 and O.psnr(rgb_b.cpu(), rgb_a.cpu()) > 80
    # an in-place weight update must be picked up (the sums are a snapshot keyed on tensor versions)
    tab = net.tpose_human.part_networks[0].embedder.dense
    with torch.no_grad():
        tab.mul_(0.5)
    try:
        rgb_c = eng_sum.render_rays(o, d, n, f, cfg.N_samples, batch=gb)[0]
        rgb_d = eng_full.render_rays(o, d, n, f, cfg.N_samples, batch=gb)[0]
        assert (rgb_c - rgb_d).abs().max() < 1e-4 and (rgb_c - rgb_b).abs().max() > 1e-4
    finally:
        with torch.no_grad():
            tab.mul_(2.0)


def test_eval_render_returns_raw_and_occ_lazily(gpu):
    """Renderer.render (eval) hands out the reference's `raw` / `occ` keys (inb_renderer.py:111-115) on first access."""
    from instant_nvr_b200.renderer import Renderer
    net = gpu["nets"][1.0]
    eager = Renderer(net, return_raw=True).render(dict(gpu["gbatch"]))
    lazy = Renderer(net).render(dict(gpu["gbatch"]))
    assert set(lazy.keys()) == {"rgb_map", "acc_map"} and "raw" in lazy and "occ" in lazy and "nope" not in lazy
    assert torch.equal(lazy["rgb_map"], eager["rgb_map"])
    assert torch.equal(lazy["occ"], eager["occ"]) and torch.equal(lazy["raw"], eager["raw"])
    assert set(lazy.keys()) == {"rgb_map", "acc_map", "raw", "occ"} and lazy["raw"].device.type == "cpu"

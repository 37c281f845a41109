"""Training path on the GPU against the oracle's autograd: the oracle is plain torch
math, so ``loss.backward()`` on it IS the reference's gradient flow (tables, MLPs, latent row, deformer)."""
import os
import sys

import pytest
import torch

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
import nvr_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    cfg = PathConfig.inb_377(N_samples=24, log2_T_cap=12).with_(use_reg_distortion=True)
    frame = make_frame(seed=4)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=4, table_gain=200.0, bounds=frame["bounds"][0])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    gnet = Network(cfg, device="cpu")
    gnet.load_state_dict(sd)
    gnet = gnet.cuda()
    rays = make_rays(frame, 20, 20)
    gbatch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**frame, **rays}.items()}
    return dict(cfg=cfg, frame=frame, rays=rays, sd=sd, net=gnet, gbatch=gbatch)


def _trainable_names(net):
    from instant_nvr_b200.training import trainable
    return [n for n, _ in trainable(net)]


def _rel(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def test_network_forward_backward_matches_oracle_autograd(setup):
    cfg, net, sd, frame, rays = setup["cfg"], setup["net"], setup["sd"], setup["frame"], setup["rays"]
    b = O.strip_batch({**frame, **rays})
    pts, _ = O.sample_along_rays(b["ray_o"], b["ray_d"], b["near"], b["far"], cfg.N_samples)
    wpts = pts.reshape(-1, 3).contiguous()
    vd = b["ray_d"][:, None].expand(-1, cfg.N_samples, 3).reshape(-1, 3).contiguous()
    N = wpts.shape[0]
    g = torch.Generator().manual_seed(21)
    Wr = torch.randn(N, 4, generator=g)
    names = _trainable_names(net)

    # ---- oracle with autograd
    sdr = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    raw_o, occ_o, st = O.network_forward(sdr, wpts, vd, b, cfg.smpl_thresh, want_stages=True)
    M = st["pind"].shape[0]
    Wd = torch.randn(M, 5, 3, generator=g)
    loss_o = (raw_o * Wr).sum() + 3.0 * (st["resd"] * Wd).sum() + 0.5 * occ_o.sum()
    loss_o.backward()

    # ---- ours
    net.train()
    for p in net.parameters():
        p.grad = None
    ret = net(wpts.cuda(), vd.cuda(), None, setup["gbatch"])
    assert ret["resd"].shape == (1, M, 5, 3) and ret["tpts"].shape == (1, 5 * M, 3) and ret["tocc"].shape == (1, 5 * M, 1)
    # (samples whose parts are evaluated outside their bbox carry the reference's own extrapolation noise,
    #  DESIGN.md section 2, hence 1e-3 here; the strict per-sample bound is tests/test_gpu_parity.py's job)
    print("[train] fwd max err raw", (ret["raw"][0].cpu() - raw_o.detach()).abs().max().item(),
          "resd", (ret["resd"][0].cpu() - st["resd"].detach()).abs().max().item())
    assert (ret["raw"][0].cpu() - raw_o.detach()).abs().max() < 1e-3
    assert (ret["resd"][0].cpu() - st["resd"].detach()).abs().max() < 1e-5
    assert (ret["tocc"].reshape(M, 5).cpu() - st["raws"][..., 3].detach()).abs().max() < 1e-3
    flag = st["flag"]
    assert (ret["tpts"].reshape(M, 5, 3).cpu() - st["bigpose"].detach())[flag].abs().max() < 1e-4
    loss = (ret["raw"][0] * Wr.cuda()).sum() + 3.0 * (ret["resd"][0] * Wd.cuda()).sum() + 0.5 * ret["occ"].sum()
    loss.backward()
    net.eval()
    assert abs(loss.item() - loss_o.item()) < 1e-3 * max(1.0, abs(loss_o.item()))
    worst = {}
    params = dict(net.named_parameters())
    for n in names:
        ref = sdr[n].grad
        ours = params[n].grad
        assert ours is not None, n
        if ref is None:
            ref = torch.zeros_like(sdr[n])
        worst[n] = _rel(ours.cpu(), ref)
    bad = {k: v for k, v in worst.items() if v > 2e-3}
    print("[train] worst relative gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    assert not bad, bad
    # every gradient that the reference produces non-zero must be non-zero here too
    for n in names:
        if sdr[n].grad is not None and sdr[n].grad.abs().max() > 0:
            assert params[n].grad.abs().max() > 0, n


def test_composite_matches_oracle_autograd(setup):
    eng = setup["net"].engine()
    from instant_nvr_b200.training import composite
    g = torch.Generator().manual_seed(22)
    raw = torch.rand(300, 37, 4, generator=g)
    raw[5, :, 3] = 0.0
    raw[6, 3, 3] = 1.0                                  # a fully opaque sample: no division by (1 - alpha)
    ro = raw.clone().requires_grad_(True)
    w_o, rgb_o, acc_o = O.composite(ro)
    Ww, Wc, Wa = torch.randn(300, 37, generator=g), torch.randn(300, 3, generator=g), torch.randn(300, generator=g)
    ((w_o * Ww).sum() + (rgb_o * Wc).sum() + (acc_o * Wa).sum()).backward()
    rg = raw.cuda().requires_grad_(True)
    w, rgb, acc = composite(eng, rg)
    ((w * Ww.cuda()).sum() + (rgb * Wc.cuda()).sum() + (acc * Wa.cuda()).sum()).backward()
    assert (w.detach().cpu() - w_o.detach()).abs().max() < 1e-6 and (rgb.detach().cpu() - rgb_o.detach()).abs().max() < 1e-5
    ok = torch.ones(300, dtype=torch.bool)
    ok[6] = False                                       # torch's cumprod backward is itself ill-defined at alpha == 1
    assert (rg.grad.cpu() - ro.grad)[ok].abs().max() < 2e-5
    assert torch.isfinite(rg.grad).all()


def test_deformer_backward_matches_oracle_autograd(setup):
    net, sd, frame = setup["net"], setup["sd"], setup["frame"]
    g = torch.Generator().manual_seed(23)
    tb = frame["tbounds"][0]
    x = (tb[0] + (tb[1] - tb[0]) * torch.rand(1, 777, 3, generator=g)).contiguous()
    W = torch.randn(1, 777, 3, generator=g)
    names = [n for n in _trainable_names(net) if n.startswith("tpose_deformer.")]
    sdr = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    (O.deformer(sdr, x[0], frame["tuv"][0], tb, frame["frame_dim"]) * W[0]).sum().backward()
    net.train()
    for p in net.parameters():
        p.grad = None
    (net.resd(x.cuda(), setup["gbatch"]) * W.cuda()).sum().backward()
    net.eval()
    params = dict(net.named_parameters())
    for n in names:
        assert _rel(params[n].grad.cpu(), sdr[n].grad) < 1e-3, n


def test_render_train_step(setup):
    """Renderer.render in training mode: the reference's training-branch outputs, finite gradients everywhere,
    and gradient steps on a fixed target reduce the image loss (step length from the first-order prediction)."""
    from instant_nvr_b200.renderer import Renderer
    net, gb = setup["net"], setup["gbatch"]
    saved = {k: v.clone() for k, v in net.state_dict().items()}
    net.train()
    r = Renderer(net)
    g = torch.Generator(device="cuda").manual_seed(24)
    target = torch.rand(1, gb["ray_o"].shape[1], 3, device="cuda", generator=g)
    params = [p for p in net.parameters() if p.requires_grad]
    losses = []
    for it in range(4):
        for p in params:
            p.grad = None
        ret = r.render(dict(gb))
        for k in ("rgb_map", "acc_map", "raw", "resd", "tpts", "tocc", "oresd", "reg_distortion_loss"):
            assert k in ret, k
        assert ret["rgb_map"].shape == target.shape and ret["reg_distortion_loss"].shape == (1, target.shape[1])
        # the reference's Renderer flattens resd to (1, 5N', 3) (inb_renderer.py:134-136); Network.forward keeps (1, N', 5, 3)
        assert ret["resd"].shape == ret["tpts"].shape and ret["tocc"].shape == ret["tpts"].shape[:2] + (1,)
        img = ((ret["rgb_map"] - target) ** 2).mean()
        reg = 0.1 * ret["reg_distortion_loss"].mean() + 0.1 * torch.norm(ret["resd"], dim=2).mean()
        if ret["oresd"].numel():
            reg = reg + 0.01 * (ret["oresd"] ** 2).mean()
        losses.append(img.item())
        if it == 0:
            (img + reg).backward()                      # every loss term of the reference's wrapper has a gradient path
            for n, p in net.named_parameters():
                if p.requires_grad and p.grad is not None:
                    assert torch.isfinite(p.grad).all(), n
            for p in params:
                p.grad = None
            ret = r.render(dict(gb))
            img = ((ret["rgb_map"] - target) ** 2).mean()
        img.backward()
        gn2 = sum(float((p.grad ** 2).sum()) for p in params if p.grad is not None)
        assert gn2 > 0
        # backtracking along -grad: the hash tables (gain 200) make the loss very non-linear, so only a descent
        # direction is asserted, not a step length
        c, ok = 0.05, False
        for _ in range(10):
            lr = c * img.item() / gn2
            with torch.no_grad():
                for p in params:
                    if p.grad is not None:
                        p -= lr * p.grad
                new = ((r.render(dict(gb))["rgb_map"] - target) ** 2).mean().item()
                if new < img.item():
                    ok = True
                    break
                for p in params:
                    if p.grad is not None:
                        p += lr * p.grad
            c /= 4
        assert ok, "no step length along -grad reduced the image loss"
    net.eval()
    net.load_state_dict(saved)
    print("[train] image loss over 4 gradient steps:", losses)
    assert all(b < a for a, b in zip(losses, losses[1:]))


def test_render_train_outputs_match_oracle(setup):
    """Renderer.render in training mode (perturb 0) against the oracle's render_train, which tests/test_oracle_train_golden.py
    pins to the reference's own training forward: every output that does not depend on the pair regulariser's random
    neighbour draw (the first half of `oresd` is the gathered `resd`, the second half the deformer at jittered points)."""
    from instant_nvr_b200.renderer import Renderer
    cfg, net, sd, frame, rays, gb = setup["cfg"], setup["net"], setup["sd"], setup["frame"], setup["rays"], setup["gbatch"]
    ref = O.render_train(sd, {**frame, **rays}, cfg.N_samples, cfg.smpl_thresh, use_pair_reg=True, use_reg_distortion=True)
    net.train()
    try:
        with torch.no_grad():
            ret = Renderer(net).render(dict(gb))
    finally:
        net.eval()
    M5 = ref["tpts"].shape[1]
    assert ret["resd"].shape == (1, M5, 3) and ret["tpts"].shape == (1, M5, 3) and ret["tocc"].shape == (1, M5, 1)
    assert (ret["resd"].cpu() - ref["resd"].reshape(1, -1, 3)).abs().max() < 1e-5
    assert (ret["tocc"].cpu() - ref["tocc"]).abs().max() < 1e-3
    flag = ref["tocc"].reshape(-1) > 0
    assert (ret["tpts"].cpu() - ref["tpts"])[0][flag].abs().max() < 1e-4
    assert (ret["raw"].cpu() - ref["raw"]).abs().max() < 1e-3
    assert (ret["rgb_map"].cpu() - ref["rgb_map"]).abs().max() < 1e-3 and (ret["acc_map"].cpu() - ref["acc_map"]).abs().max() < 1e-3
    d, dr = ret["reg_distortion_loss"].cpu(), ref["reg_distortion_loss"]
    assert d.shape == dr.shape and (d - dr).abs().max() <= 1e-3 * max(dr.abs().max().item(), 1e-6)
    K = ref["oresd"].shape[1] // 2
    Kg = ret["oresd"].shape[1] // 2
    assert ret["oresd"].shape == (1, 2 * Kg, 3) and abs(Kg - K) <= 1   # same pairs selected (|tocc - 0.5| < 0.02; a pair whose
    if K and Kg == K:                                                  # tocc sits on the edge to 1e-6 may fall either way) ...
        assert (ret["oresd"][:, :K].cpu() - ref["oresd"][:, :K]).abs().max() < 1e-5     # ... and the same gathered residuals


@pytest.fixture(scope="module")
def setup_full():
    """BASELINE.json configs[2]: 1024 rays x 64 samples per step with the SHIPPED table sizes (1.14 GB), reference init (gain 1)."""
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    cfg = PathConfig.inb_377(N_samples=64).with_(use_reg_distortion=True)
    frame = make_frame(seed=4)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=5, table_gain=1.0, bounds=frame["bounds"][0])
    sd = net.state_dict()
    gnet = Network(cfg, device="cpu")
    gnet.load_state_dict(sd)
    gnet = gnet.cuda()
    rays = make_rays(frame, 32, 32)
    gbatch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**frame, **rays}.items()}
    return dict(cfg=cfg, frame=frame, rays=rays, sd=sd, net=gnet, gbatch=gbatch)


def test_train_config3_full_tables_vs_oracle_autograd(setup_full):
    """The training step's forward + backward at configs[2] size with the full-size tables against the oracle's autograd:
    forward outputs within 1e-4 (strict: reference-init magnitudes), every trainable tensor's gradient within 2e-3 of its
    scale, the big tables included (1M-row hash levels, int64 hash products past 2^32 in the backward's index arithmetic)."""
    import json
    s = setup_full
    cfg, net, sd, frame, rays = s["cfg"], s["net"], s["sd"], s["frame"], s["rays"]
    b = O.strip_batch({**frame, **rays})
    pts, _ = O.sample_along_rays(b["ray_o"], b["ray_d"], b["near"], b["far"], cfg.N_samples)
    wpts = pts.reshape(-1, 3).contiguous()
    vd = b["ray_d"][:, None].expand(-1, cfg.N_samples, 3).reshape(-1, 3).contiguous()
    N = wpts.shape[0]
    assert N == 1024 * 64
    g = torch.Generator().manual_seed(31)
    Wr = torch.randn(N, 4, generator=g)
    names = _trainable_names(net)
    sdr = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    raw_o, occ_o, st = O.network_forward(sdr, wpts, vd, b, cfg.smpl_thresh, want_stages=True)
    M = st["pind"].shape[0]
    Wd = torch.randn(M, 5, 3, generator=g)
    loss_o = (raw_o * Wr).sum() + 3.0 * (st["resd"] * Wd).sum() + 0.5 * occ_o.sum()
    loss_o.backward()

    net.train()
    for p in net.parameters():
        p.grad = None
    ret = net(wpts.cuda(), vd.cuda(), None, s["gbatch"])
    raw_err = (ret["raw"][0].cpu() - raw_o.detach()).abs().max().item()
    resd_err = (ret["resd"][0].cpu() - st["resd"].detach()).abs().max().item()
    tocc_err = (ret["tocc"].reshape(M, 5).cpu() - st["raws"][..., 3].detach()).abs().max().item()
    assert ret["resd"].shape == (1, M, 5, 3)
    assert raw_err < 1e-4 and tocc_err < 1e-4 and resd_err < 1e-5, (raw_err, tocc_err, resd_err)
    loss = (ret["raw"][0] * Wr.cuda()).sum() + 3.0 * (ret["resd"][0] * Wd.cuda()).sum() + 0.5 * ret["occ"].sum()
    loss.backward()
    net.eval()
    params = dict(net.named_parameters())
    worst = {}
    for n in names:
        ref = sdr[n].grad if sdr[n].grad is not None else torch.zeros_like(sdr[n])
        worst[n] = _rel(params[n].grad.cpu(), ref)
        if ref.abs().max() > 0:
            assert params[n].grad.abs().max() > 0, n
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:6]
    print("[train c3] survivors", M, "fwd err raw/tocc/resd", raw_err, tocc_err, resd_err, "worst grads", top)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "diag.jsonl"), "a") as f:
        f.write(json.dumps({"test": "train_config3_full", "survivors": int(M), "raw_err": raw_err, "tocc_err": tocc_err, "resd_err": resd_err,
                            "worst_grad_rel": top[0][1], "worst_grad_name": top[0][0]}) + "\n")
    bad = {k: v for k, v in worst.items() if v > 2e-3}
    assert not bad, bad


def test_training_forward_strict_at_reference_init():
    """Same small config as `setup`, but with the reference's own init magnitudes (table gain 1): raw / tocc within the
    north-star's 1e-4 (the loose 1e-3 above is extrapolation noise at gain 200)."""
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.renderer import Renderer
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    cfg = PathConfig.inb_377(N_samples=24, log2_T_cap=12).with_(use_reg_distortion=True)
    frame = make_frame(seed=4)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=4, table_gain=1.0, bounds=frame["bounds"][0])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    rays = make_rays(frame, 20, 20)
    gb = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**frame, **rays}.items()}
    ref = O.render_train(sd, {**frame, **rays}, cfg.N_samples, cfg.smpl_thresh, use_pair_reg=True, use_reg_distortion=True)
    net.train()
    try:
        with torch.no_grad():
            ret = Renderer(net).render(dict(gb))
    finally:
        net.eval()
    assert (ret["raw"].cpu() - ref["raw"]).abs().max() < 1e-4
    assert (ret["tocc"].cpu() - ref["tocc"]).abs().max() < 1e-4
    assert (ret["resd"].cpu() - ref["resd"].reshape(1, -1, 3)).abs().max() < 1e-5
    assert (ret["rgb_map"].cpu() - ref["rgb_map"]).abs().max() < 1e-4 and (ret["acc_map"].cpu() - ref["acc_map"]).abs().max() < 1e-4
    d, dr = ret["reg_distortion_loss"].cpu(), ref["reg_distortion_loss"]
    assert (d - dr).abs().max() <= 1e-4 * max(dr.abs().max().item(), 1e-6) + 1e-7


def test_fused_adam_step_refreshes_inference_tables():
    """FusedAdam writes parameters through raw pointers; the engine's pre-summed inference tables are keyed on tensor
    versions, so the optimizer must bump them -- an eval render after a step must match the full-table render."""
    from instant_nvr_b200.config import PathConfig
    from instant_nvr_b200.engine import Engine
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.optimizer import FusedAdam
    from instant_nvr_b200.synthetic import fill_weights, make_frame, make_rays
    cfg = PathConfig.inb_377(N_samples=16, log2_T_cap=12)
    frame = make_frame(seed=2)
    net = Network(cfg, device="cpu")
    fill_weights(net.state_dict(), seed=2, table_gain=200.0, bounds=frame["bounds"][0])
    net = net.cuda().eval()
    rays = make_rays(frame, 16, 16)
    gb = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in {**frame, **rays}.items()}
    o, d, n, f = gb["ray_o"][0], gb["ray_d"][0], gb["near"][0], gb["far"][0]
    eng_sum = Engine(cfg, inference_tables=True)
    eng_sum.bind_params(net)
    eng_full = Engine(cfg, inference_tables=False)
    eng_full.bind_params(net)
    before = eng_sum.render_rays(o, d, n, f, 16, batch=gb)[0].clone()
    params = [p for p in net.parameters() if p.requires_grad]
    opt = FusedAdam(params, lr=5e-2, eps=1e-15)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for p in params:
        p.grad = torch.randn(p.shape, device="cuda", generator=gen)
    v0 = params[0]._version
    opt.step()
    assert params[0]._version > v0
    after_sum = eng_sum.render_rays(o, d, n, f, 16, batch=gb)[0]
    after_full = eng_full.render_rays(o, d, n, f, 16, batch=gb)[0]
    assert (after_sum - after_full).abs().max() < 2e-4
    assert (after_sum - before).abs().max() > 1e-3


def test_train_sample_and_distortion_match_reference_formulas(setup):
    """nvr_train_sample / nvr_distortion_* against the reference's own torch statements (inb_renderer.py:15-31, 96-103) run
    eagerly on the same device: z_vals / wpts bit for bit (same operation order), the regulariser and its gradient to fp32
    summation-order rounding."""
    from instant_nvr_b200.training import _DistortionFn
    eng = setup["net"].engine()
    gb = setup["gbatch"]
    ray_o, ray_d, near, far = gb["ray_o"], gb["ray_d"], gb["near"], gb["far"]
    R = ray_o.shape[1]
    for S, perturb in ((24, True), (64, True), (64, False), (1, False), (2, True)):
        u = torch.rand(1, R, S, device="cuda", generator=torch.Generator(device="cuda").manual_seed(S)) if perturb else None
        t_vals = torch.linspace(0.0, 1.0, steps=S, device="cuda")
        z_ref = near[..., None] * (1.0 - t_vals) + far[..., None] * t_vals
        if perturb:
            mids = 0.5 * (z_ref[..., 1:] + z_ref[..., :-1])
            upper = torch.cat([mids, z_ref[..., -1:]], -1)
            lower = torch.cat([z_ref[..., :1], mids], -1)
            z_ref = lower + (upper - lower) * u
        w_ref = ray_o[:, :, None] + ray_d[:, :, None] * z_ref[..., None]
        z, wpts, vd = eng.train_sample(ray_o[0], ray_d[0], near[0], far[0], S, None if u is None else u[0])
        assert torch.equal(z, z_ref[0]), (S, perturb, (z - z_ref[0]).abs().max().item())
        assert torch.equal(wpts.view(R, S, 3), w_ref[0])
        assert torch.equal(vd.view(R, S, 3), ray_d[0][:, None].expand(R, S, 3))
    S = 64
    g = torch.Generator(device="cuda").manual_seed(3)
    z = torch.sort(torch.rand(R, S, device="cuda", generator=g) * 2 + 2, dim=-1).values
    w = torch.rand(R, S, device="cuda", generator=g) * 0.1
    wr = w.clone().requires_grad_(True)
    ww = wr.reshape(R, S, 1) * wr.reshape(R, 1, S)
    nxt = torch.cat([z[:, 1:], z[:, -1:]], dim=-1)
    mid = (z + nxt) / 2
    ref = (ww * torch.abs(mid.reshape(R, S, 1) - mid.reshape(R, 1, S))).sum(dim=-1).sum(dim=-1)
    G = torch.randn(R, device="cuda", generator=g)
    (ref * G).sum().backward()
    wo = w.clone().requires_grad_(True)
    ours = _DistortionFn.apply(eng, wo, z)
    (ours * G).sum().backward()
    assert (ours - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert (wo.grad - wr.grad).abs().max() <= 1e-5 * wr.grad.abs().max()

"""The steps either side of the path on the GPU, through the C-ABI: fused Adam against
torch.optim.Adam (the optimizer the reference constructs), camera rays against the reference-generated golden vectors
and the numpy oracle, image assembly + PSNR against the evaluator's formula."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
import rays_oracle as RO  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(REPO, "tests", "golden", "rays.npz"))


# ---------------------------------------------------------------------------------------------- Adam
SIZES = [1, 3, 4, 5, 8191, 8192, 8193, 100003] + [257] * 26 + [(300, 16), (64, 19)]      # > 24 tensors: two launches


def _make_params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s if isinstance(s, tuple) else (s,), generator=g) for s in SIZES]


def _grads(params, step, seed=100):
    g = torch.Generator().manual_seed(seed + step)
    out = []
    for p in params:
        x = torch.randn(p.shape, generator=g) * (10.0 ** torch.randint(-5, 2, p.shape, generator=g).float())
        x.view(-1)[::5] = 0.0                              # untouched hash rows
        out.append(x)
    return out


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(wd):
    from instant_nvr_b200.optimizer import FusedAdam
    init = _make_params(0)
    ref = [torch.nn.Parameter(p.clone().cuda()) for p in init]
    ours = [torch.nn.Parameter(p.clone().cuda()) for p in init]
    groups = lambda ps: [{"params": [p], "lr": 5e-4 * (1.0 if i % 2 else 0.1)} for i, p in enumerate(ps)]
    o_ref = torch.optim.Adam(groups(ref), lr=5e-4, eps=1e-15, weight_decay=wd)
    o_our = FusedAdam(groups(ours), lr=5e-4, eps=1e-15, weight_decay=wd)
    gmax = [torch.zeros_like(p) for p in ours]
    for step in range(1, 5):
        gs = _grads(init, step)
        for p, q, g, gm in zip(ref, ours, gs, gmax):
            p.grad, q.grad = g.cuda(), g.cuda()
            torch.maximum(gm, g.cuda().abs() + wd * q.detach().abs(), out=gm)
        o_ref.step()
        o_our.step()
        for i, (p, q, gm) in enumerate(zip(ref, ours, gmax)):
            sr, so = o_ref.state[p], o_our.state[q]
            assert float(so["step"]) == float(sr["step"]) == step
            assert torch.all((so["exp_avg"] - sr["exp_avg"]).abs() <= 2e-6 * sr["exp_avg"].abs() + 2e-7 * gm), i
            assert torch.all((so["exp_avg_sq"] - sr["exp_avg_sq"]).abs() <= 2e-6 * sr["exp_avg_sq"].abs() + 4e-10 * gm ** 2), i
            torch.testing.assert_close(q.detach(), p.detach(), rtol=1.2e-7, atol=4e-9)
    # checkpoints interchange: continue the fused run in torch.optim.Adam and vice versa
    import copy
    sd = copy.deepcopy(o_our.state_dict())               # load_state_dict may alias tensors that already fit
    cont = [torch.nn.Parameter(q.detach().clone()) for q in ours]
    o_cont = torch.optim.Adam(groups(cont), lr=5e-4, eps=1e-15, weight_decay=wd)
    o_cont.load_state_dict(sd)
    o_back = FusedAdam(groups([torch.nn.Parameter(p.detach().clone()) for p in ref]), lr=5e-4, eps=1e-15, weight_decay=wd)
    o_back.load_state_dict(copy.deepcopy(o_ref.state_dict()))
    gs = _grads(init, 9)
    back = [p for g in o_back.param_groups for p in g["params"]]
    for p, q, r, g in zip(ref, cont, back, gs):
        p.grad, q.grad, r.grad = g.cuda(), g.cuda(), g.cuda()
    o_ref.step(); o_cont.step(); o_back.step()
    for p, q, r in zip(ref, cont, back):
        torch.testing.assert_close(q.detach(), p.detach(), rtol=1.2e-7, atol=4e-9)
        torch.testing.assert_close(r.detach(), p.detach(), rtol=1.2e-7, atol=4e-9)


def test_fused_adam_zero_grad_and_errors():
    from instant_nvr_b200.optimizer import FusedAdam
    p = torch.nn.Parameter(torch.randn(10001).cuda())
    q = torch.nn.Parameter(torch.randn(7).cuda())          # no gradient: skipped like torch does
    opt = FusedAdam([p, q], lr=1e-3, zero_grad_in_step=True)
    p.grad = torch.randn(10001).cuda()
    buf = p.grad.data_ptr()
    q0 = q.detach().clone()
    opt.step()
    assert p.grad.data_ptr() == buf and not p.grad.any()
    assert torch.equal(q.detach(), q0) and len(opt.state[q]) == 0
    with pytest.raises(RuntimeError):
        c = torch.nn.Parameter(torch.randn(4))
        c.grad = torch.randn(4)
        FusedAdam([c]).step()                                # CPU parameter: no CPU path


def test_fused_adam_full_size_tables():
    """One step over a 671 MB table (the body part's hash levels): equality with torch.optim.Adam and the pass stays
    a single streaming sweep (timing printed for the log, not asserted)."""
    from instant_nvr_b200.optimizer import FusedAdam
    n = 10 * 1048583 * 16
    g = torch.Generator(device="cuda").manual_seed(3)
    p0 = torch.randn(n, device="cuda", generator=g) * 1e-3
    grad = torch.randn(n, device="cuda", generator=g) * 1e-4
    grad[::3] = 0
    a, b = torch.nn.Parameter(p0.clone()), torch.nn.Parameter(p0.clone())
    oa, ob = torch.optim.Adam([a], lr=5e-4, eps=1e-15), FusedAdam([b], lr=5e-4, eps=1e-15)
    times = {}
    for name, prm, opt in (("torch", a, oa), ("fused", b, ob)):
        prm.grad = grad
        opt.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        times[name] = e0.elapsed_time(e1) / 3
    print(f"adam over {n * 4 / 1e6:.0f} MB: torch {times['torch']:.3f} ms, fused {times['fused']:.3f} ms "
          f"({n * 28 / times['fused'] / 1e6:.0f} GB/s)")
    torch.testing.assert_close(b.detach(), a.detach(), rtol=2e-7, atol=1e-8)
    torch.testing.assert_close(ob.state[b]["exp_avg_sq"], oa.state[a]["exp_avg_sq"], rtol=4e-6, atol=1e-20)


# ---------------------------------------------------------------------------------------------- camera rays
def _cam(n):
    H, W = (int(v) for v in GOLD[f"c{n}_HW"])
    return H, W, GOLD[f"c{n}_K"], GOLD[f"c{n}_R"], GOLD[f"c{n}_T"], GOLD[f"c{n}_bounds"]


@pytest.mark.parametrize("n", range(3))
def test_rays_match_reference_golden(n):
    """get_rays_within_bounds_coord on the device against the reference's own output: directions within one fp32 ulp
    (float64 dot products may round differently from BLAS), near / far 2e-6 relative -- bit-identical wherever the
    direction bits are --, the same rays in the same order."""
    from instant_nvr_b200.rays import get_rays_within_bounds_coord
    H, W, K, R, T, bounds = _cam(n)
    ray_o, ray_d, near, far, mask, coord = get_rays_within_bounds_coord(H, W, K, R, T, bounds)
    ref = {k: GOLD[f"c{n}_{k}"] for k in ("ray_o", "ray_d", "near", "far", "mask")}
    mask_c = mask.cpu().numpy()
    differ = mask_c != ref["mask"]
    assert differ.sum() <= 2                               # near == far to rounding on a bbox silhouette pixel
    if differ.sum() == 0:
        assert np.array_equal(ray_o.cpu().numpy(), ref["ray_o"])
        d = ray_d.cpu().numpy()
        assert np.abs(d - ref["ray_d"]).max() <= 6e-8
        np.testing.assert_allclose(near.cpu().numpy(), ref["near"], rtol=2e-6)
        np.testing.assert_allclose(far.cpu().numpy(), ref["far"], rtol=2e-6)
        same = np.all(d == ref["ray_d"], axis=1)
        assert same.mean() > 0.9
        assert np.array_equal(near.cpu().numpy()[same], ref["near"][same]) and np.array_equal(far.cpu().numpy()[same], ref["far"][same])
    jj, ii = np.nonzero(mask_c)
    assert np.array_equal(coord.cpu().numpy(), np.stack([ii, jj], 1))          # (col, row), row-major order


@pytest.mark.parametrize("H,W", [(512, 512), (1024, 1024), (2160, 3840), (7, 5), (1, 1)])
def test_rays_size_properties(H, W):
    from instant_nvr_b200.rays import get_rays_within_bounds_coord
    _, _, K, R, T, bounds = _cam(0)
    K = K.copy()
    K[:2] *= W / 96.0                                      # same view at the new resolution
    ray_o, ray_d, near, far, mask, coord = get_rays_within_bounds_coord(H, W, K, R, T, bounds)
    n = int(mask.sum())
    assert ray_o.shape == (n, 3) and ray_d.shape == (n, 3) and near.shape == (n,) and far.shape == (n,)
    if n == 0:
        return
    assert torch.all(near < far) and torch.all(torch.isfinite(near)) and torch.all(torch.isfinite(far))
    pix = coord[:, 1] * W + coord[:, 0]
    assert torch.all(pix[1:] > pix[:-1]) and torch.all(mask.reshape(-1)[pix])
    assert torch.allclose(ray_d.norm(dim=1), torch.ones(n, device="cuda"), atol=1e-6)
    if H * W <= 1 << 20:                                   # the oracle on the same camera (seconds at 1 Mpx)
        ref = RO.get_rays_within_bounds(H, W, K, R, T, bounds)
        assert abs(n - int(ref[4].sum())) <= 4
        if n == int(ref[4].sum()) and np.array_equal(mask.cpu().numpy(), ref[4]):
            assert np.abs(ray_d.cpu().numpy() - ref[1]).max() <= 6e-8
            np.testing.assert_allclose(near.cpu().numpy(), ref[2], rtol=2e-6)


def test_rays_feed_the_renderer(golden_setup):
    """End to end without a host round trip: device-generated rays -> Renderer.render -> assembled image -> PSNR
    against the oracle rendering the numpy-generated rays."""
    import nvr_oracle as O
    from instant_nvr_b200.network import Network
    from instant_nvr_b200.rays import assemble_image, get_rays_within_bounds, psnr_metric
    from instant_nvr_b200.renderer import Renderer
    s = golden_setup
    frame = s["frame"]
    H = W = 40
    ppts = frame["ppts"][0].numpy()
    wb = np.stack([ppts.min(0) - 0.05, ppts.max(0) + 0.05]).astype(np.float32)
    K = np.array([[55.0, 0, W / 2 - 0.5], [0, 55.0, H / 2 - 0.5], [0, 0, 1.0]])
    center, cam_pos = wb.mean(0).astype(np.float64), wb.mean(0).astype(np.float64) + np.array([0.3, 0.1, -3.0])
    z = (center - cam_pos) / np.linalg.norm(center - cam_pos)
    x = np.cross(np.array([0.0, -1.0, 0.0]), z); x /= np.linalg.norm(x)
    R = np.stack([x, np.cross(z, x), z]); T = (-R @ cam_pos).reshape(3, 1)
    ro, rd, near, far, mask = get_rays_within_bounds(H, W, K, R, T, wb)
    ref_rays = RO.get_rays_within_bounds(H, W, K, R, T, wb)
    assert np.array_equal(mask.cpu().numpy(), ref_rays[4])
    net = Network(s["cfg"], device="cpu")
    net.load_state_dict(s["sd"])
    net = net.cuda().eval()
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in frame.items()}
    batch.update(ray_o=ro[None], ray_d=rd[None], near=near[None], far=far[None])
    out = Renderer(net, output_device=None).render(batch)
    cb = {**frame, "ray_o": torch.from_numpy(ref_rays[0])[None], "ray_d": torch.from_numpy(ref_rays[1])[None],
          "near": torch.from_numpy(ref_rays[2])[None], "far": torch.from_numpy(ref_rays[3])[None]}
    ref = O.render(s["sd"], cb, s["cfg"].N_samples, s["cfg"].smpl_thresh)
    img = assemble_image(out["rgb_map"][0], mask)
    img_ref = RO.assemble_image(ref["rgb_map"][0].numpy(), ref_rays[4])
    assert img.shape == (H, W, 3) and not img[~mask].any()
    assert np.abs(img.cpu().numpy() - img_ref).max() < 1e-4
    gt = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(1))
    ours = psnr_metric(img, gt.cuda())
    want = RO.psnr_metric(img.cpu().numpy(), gt.numpy())
    assert ours == pytest.approx(want, rel=1e-9)
    assert abs(ours - RO.psnr_metric(img_ref, gt.numpy())) < 0.1          # north_star: PSNR within 0.1 dB


def test_ssim_metric_matches_oracle():
    """rays.ssim_metric (nvr_ssim_sums) against the numpy oracle of the evaluator's SSIM (skimage 0.19.3 restated): assembled
    images of a render-sized frame, ragged crops, identical images."""
    import numpy as np
    import rays_oracle as RO
    from instant_nvr_b200.rays import ssim_metric
    g = torch.Generator().manual_seed(5)
    for (H, W, box) in ((64, 80, (5, 60, 7, 71)), (512, 512, (37, 480, 100, 401)), (9, 9, (0, 9, 0, 9))):
        a = torch.rand(H, W, 3, generator=g)
        b = (a + 0.05 * torch.randn(H, W, 3, generator=g)).clamp(0, 1)
        m = torch.zeros(H, W, dtype=torch.bool)
        m[box[0]:box[1], box[2]:box[3]] = True
        a[~m] = 0
        b[~m] = 0
        ref = RO.ssim(a.numpy(), b.numpy(), m.numpy())
        ours = ssim_metric(a.cuda(), b.cuda(), m.cuda())
        assert abs(ours - ref) < 1e-10, (H, W, ours, ref)
        assert abs(ssim_metric(a.cuda(), a.cuda(), m.cuda()) - 1.0) < 1e-12
    with pytest.raises(ValueError):
        m = torch.zeros(32, 32, dtype=torch.bool)
        m[3:8, 3:20] = True                                  # 5 rows: smaller than the 7 x 7 window
        ssim_metric(torch.rand(32, 32, 3).cuda(), torch.rand(32, 32, 3).cuda(), m.cuda())

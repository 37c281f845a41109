"""The camera-ray / metric restatement (oracle/rays_oracle.py) against the reference's own outputs
(tests/golden/rays.npz, made by tests/golden/make_golden_rays.py), and the per-pixel device arithmetic
(csrc/nvr_math.cuh: nvr_pixel_ray, nvr_near_far, nvr_adam_update) compiled for the host against the oracle /
torch.optim.Adam.  Runs without a GPU."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
import rays_oracle as RO  # noqa: E402

GOLD = np.load(os.path.join(REPO, "tests", "golden", "rays.npz"))
N_CAMS = 3


def cam(n):
    H, W = (int(v) for v in GOLD[f"c{n}_HW"])
    return H, W, GOLD[f"c{n}_K"], GOLD[f"c{n}_R"], GOLD[f"c{n}_T"], GOLD[f"c{n}_bounds"]


@pytest.mark.parametrize("n", range(N_CAMS))
def test_oracle_matches_reference_bitwise(n):
    H, W, K, R, T, bounds = cam(n)
    ray_o, ray_d, near, far, mask = RO.get_rays_within_bounds(H, W, K, R, T, bounds)
    assert mask.sum() > 100
    for name, got in (("ray_o", ray_o), ("ray_d", ray_d), ("near", near), ("far", far), ("mask", mask)):
        ref = GOLD[f"c{n}_{name}"]
        assert got.dtype == ref.dtype and got.shape == ref.shape, name
        assert np.array_equal(got, ref), name


def test_psnr_formula():
    rng = np.random.default_rng(0)
    a, b = rng.random((40, 30, 3)).astype(np.float32), rng.random((40, 30, 3)).astype(np.float32)
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    assert RO.psnr_metric(a, b) == pytest.approx(10 * np.log10(1.0 / mse), rel=1e-12)
    m = rng.random((40, 30)) < 0.4
    img = RO.assemble_image(a[m], m)
    assert np.array_equal(img[m], a[m].astype(np.float64)) and not img[~m].any()


# ---- the device arithmetic, compiled for the host (tests/host_emul) ---------------------------------------------
@pytest.fixture(scope="module")
def emul():
    from test_host_emul import build_emul
    return build_emul()


@pytest.mark.parametrize("n", range(N_CAMS))
def test_device_ray_math_matches_oracle(emul, n):
    """Per pixel: ray_d within 1 fp32 ulp of the reference (float64 dot products may round differently from BLAS),
    near/far within 2e-6 relative, mask_at_box identical except where |near - far| is at rounding level."""
    H, W, K, R, T, bounds = cam(n)
    Kinv = np.ascontiguousarray(np.linalg.inv(K))
    Rm, Tv = np.ascontiguousarray(R.reshape(9)), np.ascontiguousarray(T.reshape(3))
    ray_d = np.zeros((H * W, 3), np.float32)
    near, far = np.zeros(H * W, np.float32), np.zeros(H * W, np.float32)
    mask = np.zeros(H * W, np.uint8)
    o = np.zeros(3, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    b32 = np.ascontiguousarray(bounds, dtype=np.float32)
    emul.emul_rays(C.c_int(H), C.c_int(W), p(Kinv), p(Rm), p(Tv), p(b32), p(o), p(ray_d), p(near), p(far), p(mask))
    ro, rd = RO.get_rays(H, W, K, R, T)
    ro32, rd32 = ro.reshape(-1, 3).astype(np.float32), rd.reshape(-1, 3).astype(np.float32)
    assert np.array_equal(o, ro32[0])
    assert np.abs(ray_d - rd32).max() <= 6e-8
    ref = RO.get_rays_within_bounds(H, W, K, R, T, bounds)
    ref_mask = ref[4].reshape(-1)
    differ = mask.astype(bool) != ref_mask
    assert differ.sum() <= 2
    both = mask.astype(bool) & ref_mask
    sel_ref = both[ref_mask]
    np.testing.assert_allclose(near[both], ref[2][sel_ref], rtol=2e-6)
    np.testing.assert_allclose(far[both], ref[3][sel_ref], rtol=2e-6)
    same_d = np.all(ray_d == rd32, axis=1) & both               # identical direction bits => identical near / far bits
    assert same_d.sum() > 0.9 * both.sum()
    assert np.array_equal(near[same_d], ref[2][same_d[ref_mask]]) and np.array_equal(far[same_d], ref[3][same_d[ref_mask]])


@pytest.mark.parametrize("wd,steps", [(0.0, 5), (0.01, 3)])
def test_device_adam_math_matches_torch(emul, wd, steps):
    """nvr_adam_update (+ the host-side bias corrections of nvr_adam_step) against torch.optim.Adam itself, the
    optimizer lib/train/optimizer.py:27 constructs (eps 1e-15)."""
    torch.manual_seed(0)
    n = 4099
    p0 = torch.randn(n)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=5e-4, eps=1e-15, weight_decay=wd)
    p, m, v = p0.clone().numpy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    gmax = np.zeros(n, np.float32)
    for step in range(1, steps + 1):
        g = torch.randn(n) * (10.0 ** torch.randint(-6, 2, (n,)).float())
        g[::7] = 0.0                                     # rows of the hash tables no sample touched
        p_ref.grad = g.clone()
        opt.step()
        gn = g.numpy().copy()
        emul.emul_adam(ptr(p), ptr(gn), ptr(m), ptr(v), C.c_longlong(n), C.c_longlong(step), C.c_double(5e-4), C.c_double(wd),
                       C.c_double(0.9), C.c_double(0.999), C.c_double(1e-15))
        st = opt.state[p_ref]
        # g + wd*p and the lerp cancel when terms change scale or sign: the error is an ulp of the largest term seen
        gmax = np.maximum(gmax, np.abs(gn) + wd * np.abs(p))
        m_ref, v_ref = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
        assert np.all(np.abs(m - m_ref) <= 2e-6 * np.abs(m_ref) + 2e-7 * gmax)
        assert np.all(np.abs(v - v_ref) <= 2e-6 * np.abs(v_ref) + 4e-7 * (1 - 0.999) * gmax ** 2)
        np.testing.assert_allclose(p, p_ref.detach().numpy(), rtol=1.2e-7, atol=4e-9)           # one ulp of p + the update's own rounding


def test_ssim_oracle_matches_direct_windowed_evaluation():
    """oracle ssim (skimage 0.19.3 structural_similarity restated over scipy's uniform_filter, as skimage computes it) against a
    direct per-window evaluation of the published formula; plus the identities SSIM(a, a) = 1 and symmetry."""
    import numpy as np
    import rays_oracle as RO
    rng = np.random.default_rng(0)
    H, W = 40, 50
    a = rng.random((H, W, 3)).astype(np.float32).astype(np.float64)
    b = np.clip(a + 0.1 * rng.standard_normal((H, W, 3)), 0, 1).astype(np.float32).astype(np.float64)
    m = np.zeros((H, W), bool)
    m[5:33, 8:41] = True
    s = RO.ssim(a, b, m)
    aa, bb = a[5:33, 8:41], b[5:33, 8:41]
    vals = []
    for c in range(3):
        tot, cnt = 0.0, 0
        for y in range(3, aa.shape[0] - 3):
            for x in range(3, aa.shape[1] - 3):
                wa, wb = aa[y - 3:y + 4, x - 3:x + 4, c], bb[y - 3:y + 4, x - 3:x + 4, c]
                ux, uy = wa.mean(), wb.mean()
                k = 49 / 48
                vx, vy, vxy = k * ((wa * wa).mean() - ux * ux), k * ((wb * wb).mean() - uy * uy), k * ((wa * wb).mean() - ux * uy)
                tot += ((2 * ux * uy + 4e-4) * (2 * vxy + 3.6e-3)) / ((ux * ux + uy * uy + 4e-4) * (vx + vy + 3.6e-3))
                cnt += 1
        vals.append(tot / cnt)
    assert abs(s - np.mean(vals)) < 1e-12
    assert abs(RO.ssim(a, a, m) - 1.0) < 1e-12 and abs(RO.ssim(b, a, m) - s) < 1e-12

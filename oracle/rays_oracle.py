"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's camera-ray generation and evaluator metric, the
steps either side of the per-ray path.  Only tests/ may import this file.

Pinned: tests/golden/rays.npz holds the outputs of the reference's own functions
(lib/utils/if_nerf/if_nerf_data_utils.py, imported unmodified by tests/golden/make_golden_rays.py);
tests/test_rays_oracle.py checks this restatement against them bit for bit.
"""
import numpy as np


def get_rays(H, W, K, R, T):
    """if_nerf_data_utils.py:24-38.  float64 in, float64 out (the callers cast to float32)."""
    K, R, T = (np.asarray(a, dtype=np.float64) for a in (K, R, T))
    rays_o = -np.dot(R.T, T).ravel()                                                   # :26
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")   # :28-30
    xy1 = np.stack([i, j, np.ones_like(i)], axis=2)                                    # :31
    pixel_camera = np.dot(xy1, np.linalg.inv(K).T)                                     # :32
    pixel_world = np.dot(pixel_camera - T.ravel(), R)                                  # :33
    rays_d = pixel_world - rays_o[None, None]                                          # :35
    rays_d = rays_d / np.linalg.norm(rays_d, axis=2, keepdims=True)                    # :36
    rays_o = np.broadcast_to(rays_o, rays_d.shape)                                     # :37
    return rays_o, rays_d


def get_near_far(bounds, ray_o, ray_d):
    """if_nerf_data_utils.py:92-107 (float32 arrays; note ray_o[:1]: one origin for all rays)."""
    norm_d = np.linalg.norm(ray_d, axis=-1, keepdims=True)
    viewdir = ray_d / norm_d
    viewdir[(viewdir < 1e-5) & (viewdir > -1e-10)] = 1e-5
    viewdir[(viewdir > -1e-5) & (viewdir < 1e-10)] = -1e-5
    tmin = (bounds[:1] - ray_o[:1]) / viewdir
    tmax = (bounds[1:2] - ray_o[:1]) / viewdir
    t1 = np.minimum(tmin, tmax)
    t2 = np.maximum(tmin, tmax)
    near = np.max(t1, axis=-1)
    far = np.min(t2, axis=-1)
    mask_at_box = near < far
    near = near[mask_at_box] / norm_d[mask_at_box, 0]
    far = far[mask_at_box] / norm_d[mask_at_box, 0]
    return near, far, mask_at_box


def get_rays_within_bounds(H, W, K, R, T, bounds):
    """if_nerf_data_utils.py:329-343."""
    ray_o, ray_d = get_rays(H, W, K, R, T)
    ray_o = ray_o.reshape(-1, 3).astype(np.float32)
    ray_d = ray_d.reshape(-1, 3).astype(np.float32)
    near, far, mask_at_box = get_near_far(np.asarray(bounds, dtype=np.float32), ray_o, ray_d)
    return (ray_o[mask_at_box], ray_d[mask_at_box], near.astype(np.float32), far.astype(np.float32),
            mask_at_box.reshape(H, W))


def psnr_metric(img_pred, img_gt):
    """evaluators/if_nerf.py:28-31 (float64 images, as `np.zeros((H, W, 3))` makes them, :84-90)."""
    mse = np.mean((np.asarray(img_pred, dtype=np.float64) - np.asarray(img_gt, dtype=np.float64)) ** 2)
    return -10 * np.log(mse) / np.log(10)


def assemble_image(rgb, mask_at_box):
    """evaluators/if_nerf.py:84-86."""
    img = np.zeros(mask_at_box.shape + (3,))
    img[mask_at_box] = rgb
    return img


def ssim(img_pred, img_gt, mask_at_box):
    """Evaluator.ssim_metric (lib/evaluators/if_nerf.py:33-74) minus its file output: crop both (H,W,3) float64 images to
    cv2.boundingRect(mask_at_box) (:68-71) and call skimage's compare_ssim(multichannel=True) (:73).

    skimage is a third-party dependency absent from /root/reference (requirements.txt pins scikit-image==0.19.3) and from
    this image, so ``structural_similarity`` is RESTATED here from its published source (skimage/metrics/_structural_similarity.py
    @ v0.19.3): win_size 7, uniform_filter means (scipy.ndimage, as skimage does), sample covariance NP/(NP-1), data_range =
    dtype_range[float64] = 2, K1 0.01, K2 0.03, mean of S cropped by (win_size-1)//2, mean over channels.  PARITY UNPINNED against
    skimage itself (it cannot be imported here); the restatement is checked against a direct windowed evaluation."""
    import numpy as np
    from scipy.ndimage import uniform_filter
    m = np.asarray(mask_at_box, dtype=bool)
    rows, cols = np.nonzero(m.any(1))[0], np.nonzero(m.any(0))[0]
    y, x, h, w = rows[0], cols[0], rows[-1] - rows[0] + 1, cols[-1] - cols[0] + 1
    a = np.asarray(img_pred, dtype=np.float64)[y:y + h, x:x + w]
    b = np.asarray(img_gt, dtype=np.float64)[y:y + h, x:x + w]
    win, K1, K2, R = 7, 0.01, 0.03, 2.0
    NP = win ** 2
    cov_norm = NP / (NP - 1)
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    pad = (win - 1) // 2
    vals = []
    for c in range(a.shape[2]):
        x1, y1 = a[..., c], b[..., c]
        ux, uy = uniform_filter(x1, size=win), uniform_filter(y1, size=win)
        uxx, uyy, uxy = uniform_filter(x1 * x1, size=win), uniform_filter(y1 * y1, size=win), uniform_filter(x1 * y1, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        vals.append(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return float(np.mean(vals))

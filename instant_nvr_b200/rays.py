"""Device-side versions of the steps either side of the per-ray path, with the
reference's names and argument meaning:

* ``get_rays_within_bounds`` / ``get_rays_within_bounds_coord`` -- ``lib/utils/if_nerf/if_nerf_data_utils.py:329-362``
  (``get_rays`` :24-38 + ``get_near_far`` :92-107 + the ``mask_at_box`` compaction), numpy on DataLoader workers in
  the reference, one call per rendered frame (``tpose_dataset.py:438``, ``tpose_novel_view_dataset.py:206``);
* ``assemble_image`` / ``psnr_metric`` / ``ssim_metric`` -- ``lib/evaluators/if_nerf.py:28-31, 33-74, 84-113``.

All arithmetic runs in ``libnvr_b200.so`` (``nvr_generate_rays``, ``nvr_assemble_image``, ``nvr_sq_diff_sum``, ``nvr_ssim_sums``); outputs
are CUDA tensors, so a novel-view render needs no host round trip between ray generation, the render and the metric.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Tuple

import numpy as np
import torch

from .optimizer import aux_handle, check


def _dbl(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


def _generate(H: int, W: int, K, R, T, bounds, device, want_coord: bool):
    device = torch.device(device)
    lib, h = aux_handle(device)
    Kinv = _dbl(np.linalg.inv(np.asarray(K, dtype=np.float64)), 9)        # the reference's own np.linalg.inv(K) (:32)
    Rm, Tv = _dbl(R, 9), _dbl(T, 3)
    b = torch.as_tensor(np.asarray(bounds, dtype=np.float32) if not torch.is_tensor(bounds) else bounds,
                        dtype=torch.float32).to(device).contiguous()
    if b.shape != (2, 3):
        raise ValueError("bounds must be (2,3)")
    n_pix = H * W
    f32 = dict(dtype=torch.float32, device=device)
    ray_o, ray_d = torch.empty(n_pix, 3, **f32), torch.empty(n_pix, 3, **f32)
    near, far = torch.empty(n_pix, **f32), torch.empty(n_pix, **f32)
    coord = torch.empty(n_pix, dtype=torch.int32, device=device) if want_coord else None
    mask = torch.empty(n_pix, dtype=torch.uint8, device=device)
    count = torch.empty(1, dtype=torch.int32, device=device)
    ws = torch.empty(int(lib.nvr_rays_workspace_bytes(H, W)), dtype=torch.uint8, device=device)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    with torch.cuda.device(device):
        check(lib, h, lib.nvr_generate_rays(h, H, W, dp(Kinv), dp(Rm), dp(Tv), b.data_ptr(), ray_o.data_ptr(), ray_d.data_ptr(),
                                            near.data_ptr(), far.data_ptr(), coord.data_ptr() if want_coord else None,
                                            mask.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream(device).cuda_stream), "nvr_generate_rays")
    n = int(count.item())                                  # the one host sync: output sizes are data-dependent
    return ray_o[:n], ray_d[:n], near[:n], far[:n], mask.view(H, W).bool(), (coord[:n] if want_coord else None)


def get_rays_within_bounds(H: int, W: int, K, R, T, bounds, device="cuda"):
    """-> ray_o (n,3), ray_d (n,3), near (n), far (n), mask_at_box (H,W) bool -- rays of the pixels whose ray hits the
    bbox, row-major order (if_nerf_data_utils.py:329-343)."""
    return _generate(H, W, K, R, T, bounds, device, False)[:5]


def get_rays_within_bounds_coord(H: int, W: int, K, R, T, bounds, device="cuda"):
    """... plus coord (n,2) int64 = (col, row) of every ray (if_nerf_data_utils.py:346-362; the reference's uint8 cast
    of the coordinates, a wrap-around bug for images wider than 255 px, is not reproduced)."""
    ray_o, ray_d, near, far, mask, c = _generate(H, W, K, R, T, bounds, device, True)
    c = c.long()
    return ray_o, ray_d, near, far, mask, torch.stack([c % W, c // W], dim=1)


def assemble_image(rgb: torch.Tensor, mask_at_box: torch.Tensor) -> torch.Tensor:
    """``img = zeros((H,W,3)); img[mask_at_box] = rgb`` (evaluators/if_nerf.py:84-91) on the device."""
    H, W = mask_at_box.shape
    device = rgb.device
    lib, h = aux_handle(device)
    rgb = rgb.reshape(-1, 3).to(torch.float32).contiguous()
    coord = mask_at_box.reshape(-1).nonzero(as_tuple=True)[0].to(torch.int32)
    if coord.numel() != rgb.shape[0]:
        raise ValueError("rgb must have one row per True pixel of mask_at_box")
    img = torch.empty(H * W, 3, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib, h, lib.nvr_assemble_image(h, rgb.data_ptr(), coord.data_ptr(), rgb.shape[0], H * W, img.data_ptr(),
                                             torch.cuda.current_stream(device).cuda_stream), "nvr_assemble_image")
    return img.view(H, W, 3)


def mse_metric(pred: torch.Tensor, gt: torch.Tensor) -> float:
    """``np.mean((pred - gt) ** 2)`` in float64 (evaluators/if_nerf.py:28, 108)."""
    if pred.shape != gt.shape:
        raise ValueError("shape mismatch")
    device = pred.device
    lib, h = aux_handle(device)
    a, b = pred.to(torch.float32).contiguous(), gt.to(device=device, dtype=torch.float32).contiguous()
    out = torch.empty(1, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib, h, lib.nvr_sq_diff_sum(h, a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(),
                                          torch.cuda.current_stream(device).cuda_stream), "nvr_sq_diff_sum")
    return float(out.item()) / max(a.numel(), 1)


def psnr_metric(img_pred: torch.Tensor, img_gt: torch.Tensor) -> float:
    """``-10 * log(mse) / log(10)`` (evaluators/if_nerf.py:28-31)."""
    mse = mse_metric(img_pred, img_gt)
    return -10.0 * math.log(mse) / math.log(10.0) if mse > 0 else float("inf")


def ssim_metric(img_pred: torch.Tensor, img_gt: torch.Tensor, mask_at_box: torch.Tensor) -> float:
    """``Evaluator.ssim_metric`` (evaluators/if_nerf.py:33-74) without the numpy / cv2 round trip: the two assembled (H,W,3)
    images are cropped to ``cv2.boundingRect(mask_at_box)`` and compared with skimage 0.19.3's
    ``structural_similarity(multichannel=True)`` (uniform 7x7 windows, sample covariance, data_range 2, interior mean)."""
    H, W = mask_at_box.shape
    device = img_pred.device
    lib, h = aux_handle(device)
    a = img_pred.to(torch.float32).contiguous()
    b = img_gt.to(device=device, dtype=torch.float32).contiguous()
    if a.shape != (H, W, 3) or b.shape != (H, W, 3):
        raise ValueError("images must be (H, W, 3) like mask_at_box")
    m = mask_at_box.to(device)
    rows, cols = m.any(dim=1).nonzero(as_tuple=True)[0], m.any(dim=0).nonzero(as_tuple=True)[0]
    if rows.numel() == 0:
        raise ValueError("empty mask_at_box")
    y0, y1, x0, x1 = int(rows[0]), int(rows[-1]), int(cols[0]), int(cols[-1])      # cv2.boundingRect of the mask
    w, hh = x1 - x0 + 1, y1 - y0 + 1
    if w < 7 or hh < 7:
        raise ValueError("win_size exceeds image extent")                            # what skimage raises
    out = torch.empty(3, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib, h, lib.nvr_ssim_sums(h, a.data_ptr(), b.data_ptr(), H, W, x0, y0, w, hh, out.data_ptr(),
                                        torch.cuda.current_stream(device).cuda_stream), "nvr_ssim_sums")
    return float((out / float((w - 6) * (hh - 6))).mean().item())

"""Seeded synthetic inputs for the hot path: a pseudo-SMPL frame, a pinhole ray grid and weights.

The licensed ZJU-MoCap / MonoCap data is not in the container, so every test and bench line uses
data of the reference's *shapes and conventions* built here:

* ``make_frame``  -- 6890 vertices on a capsule skeleton with 24 SMPL-ordered joints, Gaussian
  skinning weights, forward-kinematics ``A`` (T-pose -> posed) and ``big_A`` (T-pose -> big pose),
  the per-part vertex tables, a 2.5 cm ``pbw`` distance volume and a ``tuv`` volume: the keys
  ``tpose_dataset.py:454-600`` puts in ``batch``.
* ``make_rays``   -- H x W pinhole rays through the world bbox, ``near/far`` by the slab test
  (``if_nerf_data_utils.py:92-107``).
* ``fill_weights`` -- overwrite every trainable parameter of a reference-layout ``state_dict``
  from a counter-based stream keyed by parameter name, so the *reference* network (golden
  generation), the oracle and the CUDA path can be given bit-identical weights from a seed.

Everything is numpy on the CPU and deterministic; tensors come back in reference layout
(leading batch dim 1).
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np
import torch

from .config import NUM_JOINTS, NUM_PARTS, PART_NAMES

# joint -> part, lib/utils/blend_utils.py:10-16
PART_JOINTS = {
    "body": [14, 13, 9, 6, 3, 0],
    "leg": [1, 2, 4, 5, 7, 8, 10, 11],
    "head": [12, 15],
    "larm": [16, 18, 20, 22],
    "rarm": [17, 19, 21, 23],
}
PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

# T-pose joints (metres, y up) laid out so the big-pose body sits inside the inb_377 part bboxes
_J = np.array([
    [0.00, -0.25, 0.00],   # 0 pelvis
    [0.09, -0.33, 0.00],   # 1 L hip
    [-0.09, -0.33, 0.00],  # 2 R hip
    [0.00, -0.13, 0.00],   # 3 spine1
    [0.10, -0.72, 0.00],   # 4 L knee
    [-0.10, -0.72, 0.00],  # 5 R knee
    [0.00, 0.00, 0.00],    # 6 spine2
    [0.10, -1.10, -0.02],  # 7 L ankle
    [-0.10, -1.10, -0.02], # 8 R ankle
    [0.00, 0.10, 0.00],    # 9 spine3
    [0.11, -1.15, 0.10],   # 10 L foot
    [-0.11, -1.15, 0.10],  # 11 R foot
    [0.00, 0.31, 0.00],    # 12 neck
    [0.07, 0.22, 0.00],    # 13 L collar
    [-0.07, 0.22, 0.00],   # 14 R collar
    [0.00, 0.45, 0.00],    # 15 head
    [0.20, 0.23, 0.00],    # 16 L shoulder
    [-0.20, 0.23, 0.00],   # 17 R shoulder
    [0.46, 0.23, 0.00],    # 18 L elbow
    [-0.46, 0.23, 0.00],   # 19 R elbow
    [0.70, 0.23, 0.00],    # 20 L wrist
    [-0.70, 0.23, 0.00],   # 21 R wrist
    [0.80, 0.23, 0.00],    # 22 L hand
    [-0.80, 0.23, 0.00],   # 23 R hand
], dtype=np.float64)

# capsules (joint a, joint b, radius); a sphere is a capsule with a == b
_BONES = [
    (0, 3, 0.125), (3, 6, 0.125), (6, 9, 0.13), (9, 12, 0.10), (13, 16, 0.07), (14, 17, 0.07),
    (12, 15, 0.055), (15, 15, 0.105),
    (1, 4, 0.075), (4, 7, 0.052), (7, 10, 0.04), (2, 5, 0.075), (5, 8, 0.052), (8, 11, 0.04),
    (16, 18, 0.046), (18, 20, 0.038), (20, 22, 0.032), (17, 19, 0.046), (19, 21, 0.038), (21, 23, 0.032),
]


def _rodrigues(rvec: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _rigid_chain(poses: np.ndarray) -> np.ndarray:
    """SMPL forward kinematics with the rest pose removed (what
    ``if_nerf_dutils.get_rigid_transformation`` returns): (24,3) axis-angles -> (24,4,4)."""
    G = np.zeros((NUM_JOINTS, 4, 4))
    for j in range(NUM_JOINTS):
        T = np.eye(4)
        T[:3, :3] = _rodrigues(poses[j])
        T[:3, 3] = _J[j] - (_J[PARENTS[j]] if PARENTS[j] >= 0 else 0.0)
        G[j] = T if PARENTS[j] < 0 else G[PARENTS[j]] @ T
    A = G.copy()
    for j in range(NUM_JOINTS):
        A[j, :3, 3] = G[j, :3, 3] - G[j, :3, :3] @ _J[j]
    return A


def _capsule_surface(rng: np.random.Generator, a, b, r, n) -> np.ndarray:
    axis = b - a
    length = np.linalg.norm(axis)
    pts = rng.standard_normal((n, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    if length < 1e-9:
        return a + r * pts
    ax = axis / length
    # area split between the cylinder wall and the two hemispherical caps
    wall = rng.random(n) < (length / (length + 2 * r))
    t = rng.random(n) * length
    radial = pts - (pts @ ax)[:, None] * ax
    radial /= np.linalg.norm(radial, axis=1, keepdims=True) + 1e-12
    out = np.where(wall[:, None], a + t[:, None] * ax + r * radial,
                   np.where(((pts @ ax) > 0)[:, None], b + r * pts, a + r * pts))
    return out


def _lbs(W: np.ndarray, A: np.ndarray, V: np.ndarray) -> np.ndarray:
    M = np.einsum("vj,jab->vab", W, A)
    return np.einsum("vab,vb->va", M[:, :3, :3], V) + M[:, :3, 3]


def _subject(rng: np.random.Generator, n_verts: int, pose_scale: float) -> Dict[str, np.ndarray]:
    """The pseudo-SMPL subject and its pose: everything the reference reads from ``lbs/*.npy``, ``smpl-meta/*.npy``,
    ``new_params/{i}.npy`` and ``new_vertices/{i}.npy`` (float64; callers cast)."""
    # ---- T-pose vertices on capsules, count proportional to surface area -------------------
    areas = np.array([2 * np.pi * r * np.linalg.norm(_J[b] - _J[a]) + 4 * np.pi * r * r for a, b, r in _BONES])
    counts = np.floor(areas / areas.sum() * n_verts).astype(int)
    counts[0] += n_verts - counts.sum()
    V_T = np.concatenate([_capsule_surface(rng, _J[a], _J[b], r, n) for (a, b, r), n in zip(_BONES, counts)])
    V_T = V_T[rng.permutation(n_verts)]
    # ---- Gaussian skinning weights on joint distance, top-4, normalised --------------------
    d2 = ((V_T[:, None] - _J[None]) ** 2).sum(-1)
    W = np.exp(-d2 / (2 * 0.09 ** 2))
    kth = np.sort(W, axis=1)[:, -4][:, None]
    W = np.where(W >= kth, W, 0.0)
    W /= W.sum(1, keepdims=True)
    joint2part = np.zeros(NUM_JOINTS, dtype=int)
    for pid, name in enumerate(PART_NAMES):
        joint2part[PART_JOINTS[name]] = pid
    parts = joint2part[W.argmax(1)]
    # ---- poses --------------------------------------------------------------------------
    big_poses = np.zeros((NUM_JOINTS, 3))
    big_poses[1] = [0, 0, 7 / 180 * np.pi]          # tpose_dataset.py:283-286 (arms lowered less so the
    big_poses[2] = [0, 0, -7 / 180 * np.pi]         # synthetic body stays inside the yaml arm bboxes)
    big_poses[16] = [0, 0, -20 / 180 * np.pi]
    big_poses[17] = [0, 0, 20 / 180 * np.pi]
    big_A = _rigid_chain(big_poses)
    poses = pose_scale * rng.standard_normal((NUM_JOINTS, 3))
    poses[0] = 0
    A = _rigid_chain(poses)
    V_big = _lbs(W, big_A, V_T)
    V_pose = _lbs(W, A, V_T)
    Rh = np.array([0.1, 0.35, -0.05])
    R = _rodrigues(Rh)
    Th = np.array([[0.12, 0.04, -0.2]])
    V_world = V_pose @ R.T + Th                      # blend_utils.py:385-392 (inverse of world->pose)
    return {"V_T": V_T, "W": W, "parts": parts, "big_poses": big_poses, "big_A": big_A, "poses": poses, "A": A,
            "V_big": V_big, "V_pose": V_pose, "Rh": Rh, "R": R, "Th": Th, "V_world": V_world}


def make_subject(seed: int = 0, n_verts: int = 6890, pose_scale: float = 0.2) -> Dict[str, np.ndarray]:
    """The raw per-subject / per-frame SMPL arrays of ``make_frame(seed)`` in the dtypes the reference's dataset loads
    them with (``tpose_dataset.py:83-110, 247-265, 363-366``): what the frame preprocessing step (SURVEY.md 8(f)
    rank 3, ``instant_nvr_b200.smpl_frame``) starts from."""
    s = _subject(np.random.default_rng(seed), n_verts, pose_scale)
    return {
        "joints": _J.astype(np.float32), "parents": np.array([0] + PARENTS[1:], dtype=np.int64),
        "weights": s["W"].astype(np.float32), "tpose": s["V_big"].astype(np.float32),
        "wxyz": s["V_world"].astype(np.float32), "Rh": s["Rh"].astype(np.float32)[None], "Th": s["Th"].astype(np.float32),
        "poses": s["poses"].astype(np.float32).reshape(-1), "big_poses": s["big_poses"].astype(np.float32),
    }


def make_frame(seed: int = 0, n_verts: int = 6890, pose_scale: float = 0.2, voxel: float = 0.025,
               latent_index: int = 25, num_train_frame: int = 100) -> Dict[str, torch.Tensor]:
    """One reference-style per-frame ``batch`` (without rays).  All tensors CPU, batch dim 1."""
    rng = np.random.default_rng(seed)
    s = _subject(rng, n_verts, pose_scale)
    W, parts, A, big_A = s["W"], s["parts"], s["A"], s["big_A"]
    V_big, V_pose, V_world, R, Th = s["V_big"], s["V_pose"], s["V_world"], s["R"], s["Th"]

    def bounds_of(x, pad):
        return np.stack([x.min(0) - pad, x.max(0) + pad]).astype(np.float32)
    pbounds, wbounds, tbounds = bounds_of(V_pose, 0.05), bounds_of(V_world, 0.05), bounds_of(V_big, 0.05)

    # ---- per-part vertex tables (tpose_dataset.py:570-600) -------------------------------
    lengths2 = np.array([(parts == p).sum() for p in range(NUM_PARTS)])
    maxlen = int(lengths2.max())
    part_pts = np.zeros((NUM_PARTS, maxlen, 3), np.float32)
    part_pbw = np.zeros((NUM_PARTS, maxlen, NUM_JOINTS), np.float32)
    bounds = np.zeros((NUM_PARTS, 2, 3), np.float32)
    for p in range(NUM_PARTS):
        m = parts == p
        part_pts[p, :lengths2[p]] = V_pose[m]
        part_pbw[p, :lengths2[p]] = W[m]
        bounds[p, 0] = V_big[m].min(0) - 0.2          # cfg.bbox_overlap, config.py:27
        bounds[p, 1] = V_big[m].max(0) + 0.2

    # ---- pbw volume: 24 blend weights of the nearest vertex + distance (tools/prepare_zjumocap.py:152-165)
    def grid_axes(b):
        return [np.arange(b[0, k], b[1, k] + voxel, voxel) for k in range(3)]
    gx, gy, gz = grid_axes(pbounds.astype(np.float64))
    grid = np.stack(np.meshgrid(gx, gy, gz, indexing="ij"), -1).reshape(-1, 3)
    Vp = torch.from_numpy(V_pose)
    dist = np.empty(len(grid))
    near_idx = np.empty(len(grid), dtype=np.int64)
    for s in range(0, len(grid), 16384):
        dd = torch.cdist(torch.from_numpy(grid[s:s + 16384]), Vp)
        dmin, imin = dd.min(1)
        dist[s:s + 16384] = dmin.numpy()
        near_idx[s:s + 16384] = imin.numpy()
    pbw = np.concatenate([W[near_idx], dist[:, None]], -1).reshape(len(gx), len(gy), len(gz), 25).astype(np.float32)

    # ---- tuv volume over the big-pose bounds: smooth (u,v) + a little noise ------------------
    tx, ty, tz = grid_axes(tbounds.astype(np.float64))
    tg = np.stack(np.meshgrid(tx, ty, tz, indexing="ij"), -1)
    u = (np.arctan2(tg[..., 2], tg[..., 0]) / (2 * np.pi) + 0.5)
    v = (tg[..., 1] - tbounds[0, 1]) / (tbounds[1, 1] - tbounds[0, 1])
    tuv = np.stack([u, v], -1) * 0.9 + 0.05 + 0.02 * rng.standard_normal(tg.shape[:3] + (2,))
    tuv = np.clip(tuv, 0.0, 1.0).astype(np.float32)

    f32 = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))[None]
    return {
        "R": f32(R), "Th": f32(Th),
        "pbw": f32(pbw), "pbounds": f32(pbounds), "wbounds": f32(wbounds),
        "tuv": f32(tuv), "tbounds": f32(tbounds),
        "part_pts": f32(part_pts), "part_pbw": f32(part_pbw),
        "lengths2": torch.from_numpy(lengths2.astype(np.int64))[None],
        "bounds": f32(bounds),
        "ppts": f32(V_pose), "wpts": f32(V_world), "tpts": f32(V_big),
        "A": f32(A), "big_A": f32(big_A),
        "frame_dim": torch.tensor([latent_index / num_train_frame], dtype=torch.float32),
        "latent_index": torch.tensor([latent_index], dtype=torch.int64),
    }


def make_rays(frame: Dict[str, torch.Tensor], H: int, W: int, cam_dist: float = 3.0,
              tile: Tuple[int, int, int, int] | None = None, azimuth_deg: float = 0.0,
              drop_missing: bool = False) -> Dict[str, torch.Tensor]:
    """Pinhole rays, one per pixel of an H x W image whose frustum just covers the world bbox.
    ``tile=(r0, r1, c0, c1)`` keeps only that pixel window (bounded CPU-baseline samples).
    ``azimuth_deg`` orbits the camera around the vertical axis (multi-view batches).
    Returns ray_o, ray_d (1,R,3), near, far, occupancy (1,R); rays that miss the bbox get a
    degenerate near == far interval at the bbox centre depth, or -- ``drop_missing`` -- are dropped as the
    reference's ``mask_at_box`` does (if_nerf_data_utils.py:92-107, 329-343: only rays with near < far are
    rendered); ``coord`` (1,R) then holds the row-major pixel index of every kept ray."""
    wb = frame["wbounds"][0].double().numpy()
    centre = wb.mean(0)
    half = (wb[1] - wb[0]) / 2
    az = np.deg2rad(azimuth_deg)
    fwd = np.array([np.sin(az), 0.0, np.cos(az)])           # from the bbox centre towards the camera
    right = np.array([np.cos(az), 0.0, -np.sin(az)])
    eye = centre + cam_dist * fwd
    # image plane through the bbox centre spans the bbox's horizontal / vertical extent with a 5 % margin
    span = 1.05 * max(abs(half[0] * right[0]) + abs(half[2] * right[2]), half[1])
    ys = (np.arange(H) + 0.5) / H * 2 - 1
    xs = (np.arange(W) + 0.5) / W * 2 - 1
    if tile is not None:
        ys, xs = ys[tile[0]:tile[1]], xs[tile[2]:tile[3]]
    px, py = np.meshgrid(xs * span, -ys * span)       # row 0 is the top of the image
    target = (centre[None, None] + px[..., None] * right + py[..., None] * np.array([0.0, 1.0, 0.0])).reshape(-1, 3)
    d = target - eye
    d /= np.linalg.norm(d, axis=1, keepdims=True)     # normalised as in if_nerf_data_utils.py:36
    o = np.broadcast_to(eye, d.shape)
    # slab test against the world bbox (if_nerf_data_utils.py:92-107)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (wb[0] - o) / d
        t1 = (wb[1] - o) / d
    tmin = np.nanmax(np.minimum(t0, t1), axis=1)
    tmax = np.nanmin(np.maximum(t0, t1), axis=1)
    hit = tmax > tmin
    mid = np.full(len(d), cam_dist)
    near = np.where(hit, tmin, mid)
    far = np.where(hit, tmax, mid)
    f32 = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))[None]
    if drop_missing:
        keep = np.nonzero(hit)[0]
        return {"ray_o": f32(o[keep]), "ray_d": f32(d[keep]), "near": f32(near[keep]), "far": f32(far[keep]),
                "occupancy": f32(np.ones(len(keep))), "mask_at_box": torch.from_numpy(hit)[None],
                "coord": torch.from_numpy(keep.astype(np.int64))[None]}
    return {"ray_o": f32(o), "ray_d": f32(d), "near": f32(near), "far": f32(far),
            "occupancy": f32(hit.astype(np.float32)), "mask_at_box": torch.from_numpy(hit)[None]}


# ------------------------------------------------------------------------------------------
# seeded weights
# ------------------------------------------------------------------------------------------
def _stream(seed: int, name: str) -> np.random.Generator:
    # Philox keyed by (seed, crc32(name)): independent of iteration order and of other parameters
    return np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFF, zlib.crc32(name.encode())]))


def fill_weights(sd: Dict[str, torch.Tensor], seed: int = 0, table_gain: float = 1.0,
                 mlp_gain: float = 1.0, bounds: torch.Tensor | None = None) -> None:
    """Overwrite, in place, every trainable tensor of a reference-layout ``state_dict``.

    Distributions follow the reference init (kaiming-normal tables and rgb_latent,
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) Linear weights/biases) so magnitudes are realistic;
    ``table_gain`` scales the grid features (the reference init makes them ~1e-3, which barely
    exercises the gather in a parity test).  ``bounds`` (P,2,3), if given, is written to the five
    part embedders' ``bounds`` (what the reference does at training iter 1,
    part_base_embedder.py:107-109)."""
    for name, t in sd.items():
        leaf = name.rsplit(".", 1)[-1]
        g = _stream(seed, name)
        if leaf in ("dense", "hash"):
            hname = name.rsplit(".", 1)[0] + ".hash"
            T, Fdim = sd[hname].shape[1], sd[hname].shape[2]
            std = table_gain * np.sqrt(2.0 / (T * Fdim))
            arr = g.standard_normal(t.numel(), dtype=np.float32) * np.float32(std)
        elif leaf == "rgb_latent":
            arr = g.standard_normal(t.numel(), dtype=np.float32) * np.float32(np.sqrt(2.0 / t.shape[1]))
        elif leaf == "weight" and t.dim() == 2:
            bound = mlp_gain / np.sqrt(t.shape[1])
            arr = (g.random(t.numel(), dtype=np.float32) * 2 - 1) * np.float32(bound)
        elif leaf == "bias" and t.dim() == 1:
            wname = name.rsplit(".", 1)[0] + ".weight"
            bound = 1.0 / np.sqrt(sd[wname].shape[1])
            arr = (g.random(t.numel(), dtype=np.float32) * 2 - 1) * np.float32(bound)
        else:
            continue
        with torch.no_grad():
            t.copy_(torch.from_numpy(arr).reshape(t.shape))
    if bounds is not None:
        with torch.no_grad():
            for pid in range(NUM_PARTS):
                sd[f"tpose_human.part_networks.{pid}.embedder.bounds"].copy_(bounds[pid])

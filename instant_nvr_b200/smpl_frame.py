"""Per-frame SMPL preprocessing on the device: from one frame's pose parameters and world
vertices to the frame tensors ``Network.forward`` consumes, with the reference's names and argument meaning:

* ``get_rigid_transformation``  -- ``lib/utils/if_nerf/if_nerf_data_utils.py:545-577`` (+ ``batch_rodrigues`` :523-542)
* ``SmplSubject``               -- the per-subject constants ``Dataset.__init__`` / ``load_smpl`` set up
  (``lib/datasets/h36m/tpose_dataset.py:82-110``) and the static half of the ``use_knn`` block (:570-600): part labels,
  ``lengths2``, ``part_pbw``, per-part big-pose ``bounds``, ``tbounds``
* ``prepare_frame``             -- ``Dataset.prepare_input`` (:247-293), the per-frame half of the ``use_knn`` block,
  ``get_bounds`` (``if_nerf_data_utils.py:689-696``) and the volume ``tools/prepare_zjumocap.py:474-508`` (``get_bweights``)
  pre-bakes to ``lbs/bweights/{frame}.npy``

All per-frame arithmetic runs in ``libnvr_b200.so`` (``nvr_smpl_pose_frame``, ``nvr_smpl_volume_dims``,
``nvr_smpl_bweights``); the per-subject constants are index bookkeeping done once on the host.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import cabi
from .config import NUM_JOINTS, NUM_PARTS, PART_NAMES
from .optimizer import aux_handle, check
from .synthetic import PART_JOINTS


def big_poses_default(tpose_geometry: bool = True) -> np.ndarray:
    """The canonical 'big pose' of ``tpose_dataset.py:276-287`` (``cfg.tpose_geometry`` is True, ``lib/config/config.py:237``)."""
    bp = np.zeros(NUM_JOINTS * 3)
    if tpose_geometry:
        bp[5], bp[8] = np.deg2rad(30), np.deg2rad(-30)
        return bp.reshape(-1, 3)
    bp = bp.reshape(-1, 3)
    bp[1], bp[2] = [0, 0, 7. / 180. * np.pi], [0, 0, -7. / 180. * np.pi]
    bp[16], bp[17] = [0, 0, -55. / 180. * np.pi], [0, 0, 55. / 180. * np.pi]
    return bp


def _pose_struct(Rh, Th, poses, big_poses, joints, parents) -> cabi.NvrSmplPose:
    p = cabi.NvrSmplPose()
    for name, val, n in (("Rh", Rh, 3), ("Th", Th, 3), ("poses", poses, 72), ("big_poses", big_poses, 72)):
        a = np.asarray(val, dtype=np.float64).reshape(-1)
        if a.size != n:
            raise ValueError(f"{name}: expected {n} values, got {a.size}")
        getattr(p, name)[:] = a.tolist()
    j = np.asarray(joints, dtype=np.float32).reshape(-1)
    par = np.asarray(parents).astype(np.int64).reshape(-1)
    if j.size != 72 or par.size != NUM_JOINTS:
        raise ValueError("joints must be (24,3) and parents (24,)")
    p.joints[:] = j.tolist()
    p.parents[:] = [0] + [int(v) for v in par[1:]]
    return p


class SmplSubject:
    """Per-subject constants on ``device``.  ``weights`` (V,24) = smpl-meta/weights.npy, ``tpose`` (V,3) =
    lbs/bigpose_vertices.npy, ``joints`` (24,3) = lbs/joints.npy, ``parents`` (24) = lbs/parents.npy."""

    def __init__(self, joints, parents, weights, tpose, device="cuda", bbox_overlap: float = 0.2, box_padding: float = 0.05):
        self.device = torch.device(device)
        self.joints = np.asarray(joints, dtype=np.float32)                     # tpose_dataset.py:83-84
        self.parents = np.asarray(parents).astype(np.int64)
        w = np.ascontiguousarray(weights, dtype=np.float32)
        t = np.ascontiguousarray(tpose, dtype=np.float32)
        if w.ndim != 2 or w.shape[1] != NUM_JOINTS or t.shape != (w.shape[0], 3):
            raise ValueError("weights must be (V,24) and tpose (V,3)")
        V = w.shape[0]
        # load_smpl (:96-110): the part of a vertex is the part of its dominant joint
        parts = np.zeros(V, dtype=np.int64)
        wmax = w.argmax(axis=-1)
        for pid, name in enumerate(PART_NAMES):
            for bwid in PART_JOINTS[name]:
                parts[wmax == bwid] = pid
        self.parts = parts
        # static half of tpose_dataset.py:570-600
        lengths2 = np.array([(parts == pid).sum() for pid in range(NUM_PARTS)], dtype=np.int64)
        self.maxlen = int(lengths2.max())
        vert_slot = np.full(V, -1, dtype=np.int32)
        part_pbw = np.zeros((NUM_PARTS, self.maxlen, NUM_JOINTS), dtype=np.float32)
        bounds = np.zeros((NUM_PARTS, 2, 3), dtype=np.float32)
        for pid in range(NUM_PARTS):
            idx = np.nonzero(parts == pid)[0]
            vert_slot[idx] = pid * self.maxlen + np.arange(len(idx), dtype=np.int32)
            part_pbw[pid, :len(idx)] = w[idx]
            if len(idx):
                bounds[pid, 0] = t[idx].min(axis=0) - np.float32(bbox_overlap)    # float32 arithmetic, as numpy does on a float32 array
                bounds[pid, 1] = t[idx].max(axis=0) + np.float32(bbox_overlap)
        tb = np.stack([t.min(axis=0) - np.float32(box_padding), t.max(axis=0) + np.float32(box_padding)]).astype(np.float32)
        dev = self.device
        self.n_verts = V
        self.box_padding = float(box_padding)
        self.weights = torch.from_numpy(w).to(dev)
        self.tpose = torch.from_numpy(t).to(dev)
        self.vert_slot = torch.from_numpy(vert_slot).to(dev)
        self.part_pbw = torch.from_numpy(part_pbw).to(dev)
        self.lengths2 = torch.from_numpy(lengths2).to(dev)
        self.bounds = torch.from_numpy(bounds).to(dev)
        self.tbounds = torch.from_numpy(tb).to(dev)


def get_rigid_transformation(poses, joints, parents, device="cuda") -> torch.Tensor:
    """``if_nerf_data_utils.get_rigid_transformation`` -> (24,4,4) float32 CUDA tensor."""
    device = torch.device(device)
    lib, h = aux_handle(device)
    pose = _pose_struct(np.zeros(3), np.zeros(3), poses, np.zeros(72), joints, parents)
    f32 = dict(dtype=torch.float32, device=device)
    A = torch.empty(NUM_JOINTS, 4, 4, **f32)
    R, Th, ppts, wxyz = torch.empty(9, **f32), torch.empty(3, **f32), torch.empty(1, 3, **f32), torch.zeros(1, 3, **f32)
    ws = torch.empty(int(lib.nvr_smpl_workspace_bytes(1)) + 256, dtype=torch.uint8, device=device)
    off = (-ws.data_ptr()) % 256
    out = cabi.NvrSmplOut(R.data_ptr(), Th.data_ptr(), A.data_ptr(), None, ppts.data_ptr(), None, None, None)
    with torch.cuda.device(device):
        check(lib, h, lib.nvr_smpl_pose_frame(h, C.byref(pose), wxyz.data_ptr(), 1, None, 0, 0.05, C.byref(out), ws.data_ptr() + off,
                                              ws.numel() - off, torch.cuda.current_stream(device).cuda_stream), "nvr_smpl_pose_frame")
    return A


def prepare_frame(subject: SmplSubject, wxyz, Rh, Th, poses, big_poses=None, volume: bool = True,
                  tuv: Optional[torch.Tensor] = None, latent_index: int = 0, num_train_frame: int = 1) -> Dict[str, torch.Tensor]:
    """One frame's ``batch`` entries (leading batch dim 1, on ``subject.device``), keyed as ``tpose_dataset.py:470-600`` keys
    them: R, Th, A, big_A, ppts, wpts, pbounds, wbounds, tbounds, part_pts, part_pbw, lengths2, bounds, tpts, pbw (when
    ``volume``), plus tuv / frame_dim / latent_index when ``tuv`` is given.  ``wxyz`` (V,3) may be a CUDA tensor (no copy)
    or a numpy array; Rh (3), Th (3), poses (72) are host values.  One host sync (the volume's data-dependent dims)."""
    dev = subject.device
    lib, h = aux_handle(dev)
    V = subject.n_verts
    if torch.is_tensor(wxyz):
        w = wxyz.to(device=dev, dtype=torch.float32).reshape(V, 3).contiguous()
    else:
        w = torch.from_numpy(np.ascontiguousarray(np.asarray(wxyz, dtype=np.float32).reshape(V, 3))).to(dev)
    pose = _pose_struct(Rh, Th, poses, big_poses_default() if big_poses is None else big_poses, subject.joints, subject.parents)
    f32 = dict(dtype=torch.float32, device=dev)
    R, Th_d = torch.empty(3, 3, **f32), torch.empty(1, 3, **f32)
    A, big_A = torch.empty(NUM_JOINTS, 4, 4, **f32), torch.empty(NUM_JOINTS, 4, 4, **f32)
    ppts = torch.empty(V, 3, **f32)
    part_pts = torch.zeros(NUM_PARTS, subject.maxlen, 3, **f32)
    pbounds, wbounds = torch.empty(2, 3, **f32), torch.empty(2, 3, **f32)
    ws = torch.empty(int(lib.nvr_smpl_workspace_bytes(V)) + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    wsp, wsn = ws.data_ptr() + off, ws.numel() - off
    out = cabi.NvrSmplOut(R.data_ptr(), Th_d.data_ptr(), A.data_ptr(), big_A.data_ptr(), ppts.data_ptr(), part_pts.data_ptr(),
                          pbounds.data_ptr(), wbounds.data_ptr())
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        check(lib, h, lib.nvr_smpl_pose_frame(h, C.byref(pose), w.data_ptr(), V, subject.vert_slot.data_ptr(), subject.maxlen,
                                              subject.box_padding, C.byref(out), wsp, wsn, stream), "nvr_smpl_pose_frame")
        ret = {"R": R[None], "Th": Th_d[None], "A": A[None], "big_A": big_A[None], "ppts": ppts[None], "wpts": w[None],
               "tpts": subject.tpose[None], "pbounds": pbounds[None], "wbounds": wbounds[None], "tbounds": subject.tbounds[None],
               "part_pts": part_pts[None], "part_pbw": subject.part_pbw[None], "lengths2": subject.lengths2[None],
               "bounds": subject.bounds[None]}
        if volume:
            dims, origin = (C.c_int32 * 3)(), (C.c_double * 3)()
            check(lib, h, lib.nvr_smpl_volume_dims(h, wsp, dims, origin, stream), "nvr_smpl_volume_dims")
            pbw = torch.empty(dims[0], dims[1], dims[2], 25, **f32)
            check(lib, h, lib.nvr_smpl_bweights(h, wsp, V, subject.weights.data_ptr(), dims, origin, pbw.data_ptr(), stream),
                  "nvr_smpl_bweights")
            ret["pbw"] = pbw[None]
    if tuv is not None:
        ret["tuv"] = tuv.to(device=dev, dtype=torch.float32).reshape((1,) + tuple(tuv.shape[-4:]))
        ret["frame_dim"] = torch.tensor([latent_index / num_train_frame], **f32)          # tpose_dataset.py:497
        ret["latent_index"] = torch.tensor([latent_index], dtype=torch.int64, device=dev)
    return ret

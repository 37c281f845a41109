"""Render-time multi-GPU: interleaved ray-tile sharding and frame assembly.

The reference renders on one GPU (``run.py`` is single-device; DDP exists only for training,
SURVEY.md section 2.2).  Rays are independent given the replicated weights and frame tensors, so the
path shards with no data-path collective; the only exchange is assembling the frame
(16 B/ray: rgb + acc), one ``all_gather_into_tensor`` over NCCL/NVLink.

Tiles are dealt round-robin (``tile_id % world``) rather than as contiguous bands: the body sits in
the image centre, so bands would be badly load-imbalanced.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_rays: int, rank: int, world: int, tile: int = 1024) -> torch.Tensor:
    """Ray indices owned by ``rank``: tiles of ``tile`` consecutive rays, dealt round-robin."""
    n_tiles = (n_rays + tile - 1) // tile
    if rank >= n_tiles:
        return torch.zeros(0, dtype=torch.long)
    mine = torch.arange(rank, n_tiles, world)
    idx = (mine[:, None] * tile + torch.arange(tile)[None]).reshape(-1)
    return idx[idx < n_rays]


def shard_capacity(n_rays: int, world: int, tile: int = 1024) -> int:
    """Largest shard size over ranks (shards are padded to it for the fixed-size all-gather)."""
    n_tiles = (n_rays + tile - 1) // tile
    return ((n_tiles + world - 1) // world) * tile


def assemble(local: torch.Tensor, n_rays: int, rank: int, world: int, tile: int = 1024, group=None) -> torch.Tensor:
    """All-gather per-rank shard results ``local`` (n_local, C) into the full (n_rays, C) frame on
    every rank.  One collective; shards are padded to equal length."""
    C = local.shape[1]
    cap = shard_capacity(n_rays, world, tile)
    buf = local.new_zeros(cap, C)
    buf[: local.shape[0]] = local
    if world == 1:
        gathered = buf[None]
    else:
        out = local.new_empty(world * cap, C)
        dist.all_gather_into_tensor(out, buf, group=group)
        gathered = out.view(world, cap, C)
    # tile k*world + r of the frame is tile k of rank r's shard: one permute, no per-rank scatter
    # (only the last global tile can be partial, and it is the last tile of its rank, so the zero
    # padding always falls past n_rays)
    tiles = gathered.view(world, cap // tile, tile, C).permute(1, 0, 2, 3).reshape(-1, C)
    return tiles[:n_rays].contiguous()


def render_sharded(render_fn, ray_o, ray_d, near, far, rank: int, world: int, tile: int = 1024, group=None
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``render_fn(o, d, near, far) -> (rgb (n,3), acc (n,))`` on this rank's shard, then assemble."""
    n = ray_o.shape[0]
    idx = shard_indices(n, rank, world, tile).to(ray_o.device)
    rgb, acc = render_fn(ray_o[idx], ray_d[idx], near[idx], far[idx])
    full = assemble(torch.cat([rgb, acc[:, None]], dim=1), n, rank, world, tile, group)
    return full[:, :3].contiguous(), full[:, 3].contiguous()


class _DeviceArray:
    """A borrowed device pointer as a __cuda_array_interface__ object (zero-copy torch.as_tensor)."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], owner):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None}
        self._owner = owner


class PeerFrame:
    """Frame assembly over NVLink peer memory (``include/nvr_b200.h``: nvr_frame_* / nvr_render_rays_frame).

    Every rank owns a two-slot frame buffer inside the CUDA library; the buffers are mapped into each other through CUDA IPC
    handles exchanged with one ``all_gather_object``.  ``render`` then renders this rank's shard with the compositing kernel
    storing each finished ray straight into every rank's frame, plus one flag barrier -- no padding copy, no collective
    launch, no un-permute.  The returned (n_rays, 4) = [r, g, b, acc] tensor aliases library memory and is valid until the
    call after the next one.  Every rank must make the same sequence of calls."""

    def __init__(self, engine, n_rays: int, rank: int, world: int, tile: int = 1024, group=None):
        from . import cabi
        self.eng, self.n_rays, self.rank, self.world, self.tile, self.group = engine, int(n_rays), rank, world, tile, group
        mine = cabi.NvrIpcHandle()
        engine._check(engine.lib.nvr_frame_create(engine._h, self.n_rays, rank, world, tile, C.byref(mine)), "nvr_frame_create")
        arr = (cabi.NvrIpcHandle * world)()
        if world > 1:
            blobs = [None] * world
            dist.all_gather_object(blobs, bytes(mine.bytes), group=group)
            for r, b in enumerate(blobs):
                C.memmove(arr[r].bytes, b, 64)
        engine._check(engine.lib.nvr_frame_connect(engine._h, arr), "nvr_frame_connect")
        if world > 1:
            dist.barrier(group=group)            # every rank has mapped every buffer before anyone stores into one
        self._open = True

    def _view(self, ptr: int) -> torch.Tensor:
        if self.n_rays == 0:
            return torch.empty(0, 4, device=self.eng.device)
        return torch.as_tensor(_DeviceArray(ptr, (self.n_rays, 4), self), device=self.eng.device)

    def render(self, ray_o, ray_d, near, far, n_samples: int, batch=None, want_local: bool = False):
        """This rank's shard (``shard_indices`` order) -> the assembled frame (n_rays, 4) [, the shard's own rgb (n,3), acc (n)]."""
        from .engine import _dev_f32, _stream_ptr
        eng = self.eng
        if batch is not None:
            eng.bind_frame(batch)
        eng._refresh_inference_tables()
        dev = eng.device
        ray_o, ray_d, near, far = (_dev_f32(t, dev) for t in (ray_o, ray_d, near, far))
        R = ray_o.shape[0]
        rgb = torch.empty(R, 3, dtype=torch.float32, device=dev) if want_local else None
        acc = torch.empty(R, dtype=torch.float32, device=dev) if want_local else None
        ws, ws_bytes = eng._workspace(R * n_samples)
        out = C.c_void_p()
        eng._check(eng.lib.nvr_render_rays_frame(eng._h, ray_o.data_ptr(), ray_d.data_ptr(), near.data_ptr(), far.data_ptr(), R,
                                                 int(n_samples), rgb.data_ptr() if want_local else None,
                                                 acc.data_ptr() if want_local else None, ws, ws_bytes, _stream_ptr(dev),
                                                 C.byref(out)), "nvr_render_rays_frame")
        frame = self._view(out.value)
        return (frame, rgb, acc) if want_local else frame

    def render_host(self, ray_o, ray_d, near, far, n_samples: int, rgb_out, acc_out, batch=None) -> torch.Tensor:
        """``render`` with this rank's shard in HOST (pinned) buffers and its own pixels delivered to host buffers; the copies
        ride on the render's two lanes.  Returns the assembled frame (device); the call has synchronised the stream."""
        from .engine import _stream_ptr
        eng = self.eng
        if batch is not None:
            eng.bind_frame(batch)
        eng._refresh_inference_tables()
        for t in (ray_o, ray_d, near, far, rgb_out, acc_out):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("render_host takes contiguous fp32 CPU tensors")
        R = ray_o.shape[0]
        if eng._io is None or eng._io.numel() < 12 * R:
            eng._io = torch.empty(max(12 * R, 12), dtype=torch.float32, device=eng.device)
        ws, ws_bytes = eng._workspace(R * n_samples)
        out = C.c_void_p()
        eng._check(eng.lib.nvr_render_rays_frame_host(eng._h, ray_o.data_ptr(), ray_d.data_ptr(), near.data_ptr(), far.data_ptr(), R,
                                                      int(n_samples), rgb_out.data_ptr(), acc_out.data_ptr(), eng._io.data_ptr(),
                                                      ws, ws_bytes, _stream_ptr(eng.device), C.byref(out)), "nvr_render_rays_frame_host")
        return self._view(out.value)

    def allgather(self, rgb: torch.Tensor, acc: torch.Tensor) -> torch.Tensor:
        """Unfused form: scatter already rendered shard pixels + barrier."""
        from .engine import _dev_f32, _stream_ptr
        eng = self.eng
        rgb, acc = _dev_f32(rgb, eng.device), _dev_f32(acc, eng.device)
        out = C.c_void_p()
        eng._check(eng.lib.nvr_allgather_frame(eng._h, rgb.data_ptr(), acc.data_ptr(), rgb.shape[0], _stream_ptr(eng.device),
                                               C.byref(out)), "nvr_allgather_frame")
        return self._view(out.value)

    def close(self) -> None:
        if not self._open:
            return
        self._open = False
        eng = self.eng
        eng._check(eng.lib.nvr_frame_disconnect(eng._h), "nvr_frame_disconnect")
        if self.world > 1:
            dist.barrier(group=self.group)       # nobody frees a buffer a peer still maps
        eng._check(eng.lib.nvr_frame_destroy(eng._h), "nvr_frame_destroy")

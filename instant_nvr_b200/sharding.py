"""Render-time multi-GPU: interleaved ray-tile sharding and frame assembly.

The reference renders on one GPU (``run.py`` is single-device; DDP exists only for training,
SURVEY.md section 2.2).  Rays are independent given the replicated weights and frame tensors, so the
path shards with no data-path collective; the only exchange is assembling the frame
(16 B/ray: rgb + acc), one ``all_gather_into_tensor`` over NCCL/NVLink.

Tiles are dealt round-robin (``tile_id % world``) rather than as contiguous bands: the body sits in
the image centre, so bands would be badly load-imbalanced.
"""
from __future__ import annotations

from typing import Tuple

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

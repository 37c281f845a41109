"""Parameter containers with the reference's exact ``state_dict`` layout.

These modules own ``nn.Parameter`` storage only -- names, shapes and dtypes match what
``make_network(cfg)`` builds (``inb_part_network_multiassign.py:68-75``,
``part_base_network.py:31-42``, ``part_base_embedder.py:48-88``, ``uv_deformer.py:12-21``,
``freq_embedder.py:6-10``) so ``load_network`` / ``save_model`` / ``make_optimizer``
(``lib/utils/net_utils.py:423-528``, ``lib/train/optimizer.py:13-31``) work unchanged.
No forward math lives here: the CUDA path borrows the device pointers (``cabi.py``).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .config import GridSpec, PartSpec, PathConfig


def _frozen(t: torch.Tensor) -> nn.Parameter:
    # the reference stores its buffers as requires_grad=False Parameters so they land in state_dict
    return nn.Parameter(t, requires_grad=False)


class GridParams(nn.Module):
    """Storage of one multi-resolution grid; key names follow part_base_embedder.py:48-88."""

    def __init__(self, spec: GridSpec):
        super().__init__()
        spec.check_supported()
        self.spec = spec
        res, cnt = spec.res, spec.cnt
        self.bounds = _frozen(torch.tensor(spec.bbox, dtype=torch.float32).reshape(2, 3))
        # python-float 1/(res-1) rounded to fp32 by torch.tensor(), as in the reference (:54,:57)
        self.entries_size = _frozen(torch.tensor([1 / (r - 1) for r in res]))
        self.entries_num = _frozen(torch.tensor(res))
        self.entries_min = _frozen(torch.tensor([0 for _ in res]))
        self.entries_cnt = _frozen(torch.tensor(cnt))
        self.entries_sum = _frozen(torch.tensor(cnt).cumsum(dim=-1))
        T, F = spec.T, spec.n_feat
        # kaiming_normal_ on a (L, T, F) tensor: fan_in = T*F, gain sqrt(2)  (:71-72)
        std = math.sqrt(2.0 / (T * F))
        self.dense = nn.Parameter(torch.empty(spec.dense_rows, F).normal_(0.0, std))
        self.hash = nn.Parameter(torch.empty(spec.n_hash_levels, T, F).normal_(0.0, std))
        self.offsets = _frozen(torch.tensor(
            [[float((c >> 2) & 1), float((c >> 1) & 1), float(c & 1)] for c in range(8)]))

    @property
    def out_dim(self) -> int:
        return self.spec.out_dim


class LinearStack(nn.Module):
    """``MLP`` of part_base_network.py:11-24: key prefix ``linears.{i}``."""

    def __init__(self, indim: int, outdim: int, d_hidden: int, n_layers: int):
        super().__init__()
        dims = [indim] + [d_hidden] * n_layers + [outdim]
        self.linears = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))


class _FreqBands(nn.Module):
    def __init__(self, multires: int):
        super().__init__()
        fb = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)
        self.freq_bands = _frozen(fb[:, None, None].expand(multires, 2, 1).clone())


class ViewDirParams(nn.Module):
    """freq_embedder.Embedder: key ``embedder.freq_bands`` (4,2,1)."""

    def __init__(self, multires: int):
        super().__init__()
        self.embedder = _FreqBands(multires)
        self.out_dim = 3 + 3 * 2 * multires


class PartParams(nn.Module):
    def __init__(self, cfg: PathConfig, spec: PartSpec, pid: int):
        super().__init__()
        self.pid, self.partname = pid, spec.name
        self.embedder = GridParams(spec.grid)
        self.embedder_dir = ViewDirParams(cfg.view_res)
        self.occ = LinearStack(self.embedder.out_dim, 1 + cfg.geo_feature_dim, cfg.d_hidden, 1)
        self.rgb_latent = nn.Parameter(torch.zeros(cfg.num_latent_code, cfg.latent_code_dim))
        nn.init.kaiming_normal_(self.rgb_latent)
        self.rgb = LinearStack(cfg.rgb_in_dim, 3, cfg.d_hidden, spec.rgb_hidden_layers)


class DeformerParams(nn.Module):
    """uv_deformer.Deformer: ``embedder.*`` + ``mlp.{0,2,4}``."""

    def __init__(self, cfg: PathConfig):
        super().__init__()
        self.embedder = GridParams(cfg.deformer_grid)
        h = cfg.deformer_hidden
        # indices 1 and 3 are the (parameter-free) Softplus slots of the reference's nn.Sequential
        self.mlp = nn.Sequential(nn.Linear(self.embedder.out_dim, h), nn.Identity(),
                                 nn.Linear(h, h), nn.Identity(), nn.Linear(h, 3))


class PartSet(nn.Module):
    """TPoseHuman: key prefix ``part_networks.{i}``."""

    def __init__(self, cfg: PathConfig):
        super().__init__()
        self.part_networks = nn.ModuleList(PartParams(cfg, p, i) for i, p in enumerate(cfg.parts))

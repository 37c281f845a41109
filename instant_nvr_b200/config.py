"""Static description of the per-ray hot path (what the reference reads from its global ``cfg``).

The reference resolves everything through a yacs singleton parsed at import
(``lib/config/config.py:386-397``).  The hot path only needs a handful of those keys
("Config / flags"); they are collected here in plain dataclasses so the
CUDA path, the oracle and the tests share one source of truth and nothing on the GPU box
needs the reference tree.

``PathConfig.inb_377()`` restates ``configs/inb/inb_377.yaml`` (+ ``config.py`` defaults);
``PathConfig.from_reference_cfg(cfg)`` reads a live reference ``cfg`` (drop-in use inside
``train_net.py`` / ``run.py``).
"""
from __future__ import annotations

from functools import cached_property
from dataclasses import dataclass, field, replace
from typing import List, Sequence, Tuple

# Part order is fixed by the reference: lib/utils/blend_utils.py:17
PART_NAMES = ("body", "leg", "head", "larm", "rarm")
NUM_PARTS = 5
NUM_JOINTS = 24
KNN_K = 4                     # blend_utils.py:741 (default K; cfg.knn_k is unused on the batch path)
KNN_RADIUS = 0.075            # blend_utils.py:741
KNN_EPS = 1e-8                # blend_utils.py:741
MAX_LEVELS = 16


def next_prime(n: int) -> int:
    """Smallest prime strictly greater than ``n`` (what ``sympy.nextprime`` returns,
    part_base_embedder.py:42).  Trial division is plenty for n <= 2**24."""
    def is_prime(v: int) -> bool:
        if v < 2:
            return False
        if v % 2 == 0:
            return v == 2
        f = 3
        while f * f <= v:
            if v % f == 0:
                return False
            f += 2
        return True
    c = n + 1
    while not is_prime(c):
        c += 1
    return c


@dataclass(frozen=True)
class GridSpec:
    """One multi-resolution dense+hashed grid (part_base_embedder.py:13-104).  Derived quantities are cached
    (the spec is frozen): next_prime by trial division must not run on every pointer-table rebuild."""
    n_levels: int = 16
    n_feat: int = 16
    log2_T: int = 18
    base_res: int = 2
    b: float = 1.38
    sum_features: bool = True          # sum=True & sum_over_features=True -> one scalar per level
    bbox: Tuple[Tuple[float, float, float], Tuple[float, float, float]] = ((0., 0., 0.), (1., 1., 1.))
    use_batch_bounds: bool = True      # bounds replaced from batch['bounds'] at iter_step == 1

    @cached_property
    def T(self) -> int:
        return next_prime(2 ** self.log2_T)

    @cached_property
    def res(self) -> List[int]:
        # int(base * b**i) in Python float64, exactly as part_base_embedder.py:52
        return [int(self.base_res * self.b ** i) for i in range(self.n_levels)]

    @cached_property
    def cnt(self) -> List[int]:
        return [r ** 3 for r in self.res]

    @cached_property
    def start_hash(self) -> int:
        # first level whose dense entry count exceeds the table size (part_base_embedder.py:63-67)
        T = self.T
        for i, c in enumerate(self.cnt):
            if c > T:
                return i
        return self.n_levels

    @cached_property
    def n_hash_levels(self) -> int:
        return self.n_levels - self.start_hash

    @cached_property
    def dense_rows(self) -> int:
        return sum(self.cnt[: self.start_hash])

    @cached_property
    def dense_offsets(self) -> List[int]:
        """Row offset of level l inside the packed ``dense`` table (entries_sum[l-1], :129)."""
        out, acc = [], 0
        for l in range(self.n_levels):
            out.append(acc if l < self.start_hash else 0)
            if l < self.start_hash:
                acc += self.cnt[l]
        return out

    @cached_property
    def out_dim(self) -> int:
        return 3 + (self.n_levels if self.sum_features else self.n_levels * self.n_feat)

    def check_supported(self) -> None:
        if not (1 <= self.n_levels <= MAX_LEVELS):
            raise ValueError(f"n_levels={self.n_levels} outside 1..{MAX_LEVELS}")
        if self.start_hash == 0:
            # separate_dense would be disabled by the reference (:68); the shipped configs never hit it
            raise NotImplementedError("grids with no dense level (separate_dense off) are not supported")


@dataclass(frozen=True)
class PartSpec:
    name: str
    grid: GridSpec
    rgb_hidden_layers: int = 2         # MLP(n_layers): body/head 2, leg/arms 1 (inb_377.yaml:103-107,...)


@dataclass(frozen=True)
class PathConfig:
    parts: Tuple[PartSpec, ...]
    deformer_grid: GridSpec
    N_samples: int = 64
    smpl_thresh: float = 0.05
    num_latent_code: int = 100
    latent_code_dim: int = 8
    geo_feature_dim: int = 16
    view_res: int = 4                  # viewdir_embedder.kwargs.res -> 3 + 3*2*4 = 27
    d_hidden: int = 64                 # network.occ.d_hidden == network.color.d_hidden
    deformer_hidden: int = 32          # uv_deformer.py:15-21
    ps: Tuple[int, int, int] = (1, 19349663, 83492791)   # config.py:17
    chunk: int = 4096
    render_chunk: int = 4096
    tpose_viewdir: bool = True
    aggr: str = ""
    perturb: float = 0.0
    use_pair_reg: bool = True          # config.py defaults of the reference (lib/config/config.py)
    use_reg_distortion: bool = False
    use_freespace_loss: bool = False

    # ---- derived -----------------------------------------------------------------------
    @property
    def embed_dim(self) -> int:       # 19
        return self.parts[0].grid.out_dim

    @property
    def view_dim(self) -> int:        # 27
        return 3 + 3 * 2 * self.view_res

    @property
    def rgb_in_dim(self) -> int:      # 19 + 27 + 16 + 8 = 70  (part_base_network.py:37)
        return self.embed_dim + self.view_dim + self.geo_feature_dim + self.latent_code_dim

    def check_supported(self) -> None:
        if len(self.parts) != NUM_PARTS:
            raise NotImplementedError("only the 5-part layout (cfg.part3/part6 off) is supported")
        if self.aggr != "":
            raise NotImplementedError(f"cfg.aggr={self.aggr!r}: only arg-max aggregation ('') is implemented")
        if not self.tpose_viewdir:
            raise NotImplementedError("cfg.tpose_viewdir=False is not implemented")
        for p in self.parts:
            p.grid.check_supported()
            if not p.grid.sum_features or p.grid.n_feat != 16:
                raise NotImplementedError("part grids must be F=16 with per-level feature sum")
            if p.grid.out_dim != 19:
                raise NotImplementedError("part embedding must be 3+16 wide")
            if p.rgb_hidden_layers not in (1, 2):
                raise NotImplementedError("rgb MLP must have 1 or 2 hidden layers")
        g = self.deformer_grid
        g.check_supported()
        if g.sum_features or g.out_dim != 19:
            raise NotImplementedError("deformer grid must be concat mode with 3 + L*F == 19")
        if self.d_hidden != 64 or self.deformer_hidden != 32 or self.geo_feature_dim != 16 \
                or self.latent_code_dim != 8 or self.view_res != 4:
            raise NotImplementedError("MLP widths are compiled in (64/32/16/8, view_res 4)")

    def with_(self, **kw) -> "PathConfig":
        return replace(self, **kw)

    # ---- factories ---------------------------------------------------------------------
    @classmethod
    def inb_377(cls, N_samples: int = 64, log2_T_cap: int | None = None) -> "PathConfig":
        """configs/inb/inb_377.yaml.  ``log2_T_cap`` shrinks the hash tables (tests only)."""
        def cap(v):
            return v if log2_T_cap is None else min(v, log2_T_cap)
        mk = lambda log2T, base, bbox: GridSpec(16, 16, cap(log2T), base, 1.38, True, bbox, True)
        parts = (
            PartSpec("body", mk(20, 16, ((-1., -1.2, -0.34), (0.8, 0.7, 0.5))), 2),
            PartSpec("leg", mk(20, 2, ((-1., -1.2, -0.34), (0.8, -0.3, 0.5))), 1),
            PartSpec("head", mk(18, 2, ((-0.3, 0.3, -0.3), (0.3, 0.7, 0.3))), 2),
            PartSpec("larm", mk(15, 2, ((0.2, 0., -0.2), (0.9, 0.35, 0.2))), 1),
            PartSpec("rarm", mk(15, 2, ((-0.9, 0., -0.2), (-0.2, 0.35, 0.2))), 1),
        )
        deformer = GridSpec(8, 2, 14, 4, 1.38, False, ((0., 0., 0.), (1., 1., 1.)), False)
        return cls(parts=parts, deformer_grid=deformer, N_samples=N_samples, smpl_thresh=0.05)

    @classmethod
    def from_reference_cfg(cls, cfg) -> "PathConfig":
        """Build from a live reference ``cfg`` (yacs CfgNode); mirrors what
        make_part_embedder / make_part_color_network / make_deformer read
        (lib/networks/make_network.py:26-88)."""
        def grid_from(kwargs, bbox, default_bounds_flag):
            return GridSpec(
                n_levels=int(kwargs.get("n_levels", 16)),
                n_feat=int(kwargs.get("n_features_per_level", 16)),
                log2_T=int(kwargs.get("log2_hashmap_size", 18)),
                base_res=int(kwargs.get("base_resolution", 2)),
                b=float(kwargs.get("b", 1.38)),
                sum_features=bool(kwargs.get("sum", True)) and bool(kwargs.get("sum_over_features", True)),
                bbox=tuple(tuple(float(v) for v in row) for row in bbox),
                use_batch_bounds=bool(kwargs.get("use_batch_bounds", default_bounds_flag)),
            )
        parts = []
        for name in PART_NAMES:
            pc = getattr(cfg.partnet, name)
            n_layers = int(cfg.network.color.n_layers)
            if "color_network" in pc and "kwargs" in pc.color_network:
                n_layers = int(pc.color_network.kwargs.get("n_layers", n_layers))
            parts.append(PartSpec(name, grid_from(dict(pc.embedder.kwargs), pc.bbox, cfg.use_batch_bounds), n_layers))
        dk = dict(cfg.tpose_deformer.embedder.kwargs)
        deformer = grid_from(dk, ((0., 0., 0.), (1., 1., 1.)), cfg.use_batch_bounds)
        out = cls(
            parts=tuple(parts), deformer_grid=deformer, N_samples=int(cfg.N_samples),
            smpl_thresh=float(cfg.smpl_thresh), num_latent_code=int(cfg.num_latent_code),
            latent_code_dim=int(cfg.latent_code_dim), geo_feature_dim=int(cfg.geo_feature_dim),
            view_res=int(cfg.viewdir_embedder.kwargs.res), d_hidden=int(cfg.network.occ.d_hidden),
            ps=tuple(int(v) for v in cfg.ps), chunk=int(cfg.chunk), render_chunk=int(cfg.render_chunk),
            tpose_viewdir=bool(cfg.tpose_viewdir), aggr=str(cfg.aggr), perturb=float(cfg.perturb),
            use_pair_reg=bool(getattr(cfg, "use_pair_reg", True)), use_reg_distortion=bool(getattr(cfg, "use_reg_distortion", False)),
            use_freespace_loss=bool(getattr(cfg, "use_freespace_loss", False)),
        )
        if cfg.part3 or cfg.part6 or cfg.part_deform:
            raise NotImplementedError("cfg.part3 / part6 / part_deform are outside the inb hot path")
        if bool(getattr(cfg, "use_occ_loss", False)):
            # inb_renderer.py:122-129 indexes `obj_occ.shape[1]` of a 1-D tensor: the reference itself raises with this flag
            raise NotImplementedError("cfg.use_occ_loss ('obj_occupancy', inb_renderer.py:122-129) is not implemented")
        return out

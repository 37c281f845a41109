"""Drop-in for ``lib.networks.bw_deform.inb_part_network_multiassign.Network`` (reference
``inb_part_network_multiassign.py:67-168``), selected with ``network_module:
instant_nvr_b200.network`` (``lib/networks/make_network.py:5-8``).

Same constructor contract (no arguments -> reads the reference's global ``cfg``), same
``state_dict`` layout (Appendix D of SURVEY.md), same ``forward(wpts, viewdir, dists, batch)``
return dict.  The math runs in the CUDA library behind the C-ABI (``include/nvr_b200.h``); there
is no PyTorch / CPU fallback -- calling ``forward`` without the built library or on CPU tensors
raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .config import PathConfig
from .params import DeformerParams, PartSet


def _cfg_from_reference() -> PathConfig:
    try:
        from lib.config import cfg as ref_cfg      # the reference's yacs singleton
    except Exception as e:                         # pragma: no cover - only inside the reference tree
        raise RuntimeError(
            "Network() without arguments reads the reference's global cfg (lib.config.cfg); "
            "outside the reference tree pass a PathConfig explicitly") from e
    return PathConfig.from_reference_cfg(ref_cfg)


class Network(nn.Module):
    def __init__(self, cfg: Optional[PathConfig] = None, device: Optional[str] = None):
        super().__init__()
        self.cfg = cfg if cfg is not None else _cfg_from_reference()
        self.cfg.check_supported()
        self.tpose_deformer = DeformerParams(self.cfg)
        self.tpose_human = PartSet(self.cfg)
        self._engine = None
        if device is not None:
            self.to(device)

    # ---- engine plumbing -----------------------------------------------------------------
    def engine(self):
        """The CUDA engine bound to this module's parameter storages (created lazily, re-bound
        when parameters moved / were replaced, e.g. after load_state_dict or .cuda())."""
        from .engine import Engine
        dev = next(self.parameters()).device
        if self._engine is not None and self._engine.device != dev and dev.type == "cuda":
            self._engine = None                      # the module moved to another GPU: a new handle on that device
        if self._engine is None:
            self._engine = Engine(self.cfg, device=dev if dev.type == "cuda" else None)
        if self._engine._params_key is None:         # bound until _apply() (.cuda() / .to() / ...) replaces the storages
            self._engine.bind_params(self)
        return self._engine

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._trainable_cache = None                 # training.trainable(): parameters may have been replaced
        if getattr(self, "_engine", None) is not None:
            self._engine.invalidate_params()
        return out

    # ---- reference surface -----------------------------------------------------------------
    def forward(self, wpts: torch.Tensor, viewdir: torch.Tensor, dists: torch.Tensor, batch: Dict):
        """wpts, viewdir (N,3) f32 world-space sample points / unit view directions; ``dists`` is
        accepted and ignored exactly as the reference does (part_base_network.py:44-63 never reads
        it).  Returns {'raw': (1,N,4) = [r,g,b,occ], 'occ': (1,N,1)}; in training mode also resd / tpts / tocc."""
        self._maybe_update_bounds(batch)
        if self.training:
            # + 'resd' (1,N',5,3), 'tpts' (1,5N',3), 'tocc' (1,5N',1), all attached to autograd (:161-165)
            from .training import network_train_forward
            return network_train_forward(self, wpts, viewdir, batch)
        raw, occ = self.engine().query_points(wpts, viewdir, batch)
        return {"raw": raw[None], "occ": occ[None]}

    def resd(self, tpts: torch.Tensor, batch: Dict):
        """Deformer residual at canonical points (B,N,3) -> (B,N,3)
        (inb_part_network_multiassign.py:122-124)."""
        B, N, D = tpts.shape
        if self.training and torch.is_grad_enabled():
            from .training import deformer_train
            return deformer_train(self, tpts, batch)
        return self.engine().deformer_residual(tpts.reshape(-1, 3), batch).view(B, N, D)

    def _maybe_update_bounds(self, batch: Dict) -> None:
        # part_base_embedder.py:107-109: at training iter 1 the part bboxes are replaced by the
        # data-derived ones.
        if "iter_step" in batch and batch["iter_step"] == 1 and "bounds" in batch:
            with torch.no_grad():
                for pid, part in enumerate(self.tpose_human.part_networks):
                    if part.embedder.spec.use_batch_bounds:
                        part.embedder.bounds.copy_(batch["bounds"][0][pid])

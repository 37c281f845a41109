"""Drop-in for ``lib.networks.renderer.inb_renderer.Renderer`` (reference
``inb_renderer.py:11-239``), selected with ``renderer_module: instant_nvr_b200.renderer``
(``lib/networks/renderer/make_renderer.py:5-16``).

``render(batch)`` runs sampling, the network and the compositing in the CUDA library in one
call over *all* rays (the reference's 4096-ray Python loop only existed to bound its 8 KiB/point
intermediate, ``inb_renderer.py:217-237``; the library splits passes itself by workspace size).
"""
from __future__ import annotations

from typing import Dict

import torch

from .network import Network


class RenderOutput(dict):
    """The eval render's result dict.  The reference returns the per-sample ``raw (1,R*S,4)`` / ``occ (1,R*S,1)`` with every
    render (inb_renderer.py:111-115, 199-200: 16 B per sample, copied to the CPU); almost nobody reads them (the mesh
    visualiser does, lib/visualizers/if_nerf.py:138-169), so here they are produced on first access -- ``ret['raw']``,
    ``ret['occ']``, ``'raw' in ret`` work as with the reference, at the price of one more render when they are asked for.
    They are not part of ``keys()`` / iteration until then."""

    _LAZY = ("raw", "occ")

    def __init__(self, data, producer=None):
        super().__init__(data)
        self._producer = producer

    def __missing__(self, key):
        if key in self._LAZY and self._producer is not None:
            self.update(self._producer())
            self._producer = None
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (key in self._LAZY and self._producer is not None)

    def get(self, key, default=None):
        return self[key] if key in self else default


class Renderer:
    def __init__(self, net: Network, return_raw: bool = False, output_device: str = "cpu"):
        self.net = net
        self.return_raw = return_raw          # 'raw'/'occ' per sample (16 B/sample) eagerly; otherwise lazily (RenderOutput)
        self.output_device = output_device    # the reference hands every eval output back on the CPU (:199-200)

    def render(self, batch: Dict, test: bool = False, epoch: int = -1) -> Dict[str, torch.Tensor]:
        """batch: reference layout, leading batch dim 1 -- ray_o, ray_d (1,R,3), near, far (1,R) plus the
        frame tensors.  Returns rgb_map (1,R,3), acc_map (1,R) and raw (1,R*S,4), occ (1,R*S,1) --
        eagerly with ``return_raw``, on first access otherwise (RenderOutput)."""
        net = self.net
        if net.training:
            from .training import render_train
            net._maybe_update_bounds(batch)
            return render_train(self, batch, epoch)
        if epoch != -1:
            batch["epoch"] = epoch
        ray_o, ray_d, near, far = batch["ray_o"], batch["ray_d"], batch["near"], batch["far"]
        if ray_o.shape[0] != 1:
            raise ValueError("n_batch must be 1 (the reference asserts it, inb_part_network_multiassign.py:84)")
        net._maybe_update_bounds(batch)
        S = net.cfg.N_samples
        out = net.engine().render_rays(ray_o[0], ray_d[0], near[0], far[0], S, batch=batch, want_raw=self.return_raw)
        ret = {"rgb_map": out[0][None], "acc_map": out[1][None]}
        if self.return_raw:
            ret["raw"] = out[2][None]
            ret["occ"] = out[2][None, :, 3:4].contiguous()
        if self.output_device is not None:
            ret = {k: v.detach().to(self.output_device) for k, v in ret.items()}
        if self.return_raw:
            return ret

        def produce():
            o = net.engine().render_rays(ray_o[0], ray_d[0], near[0], far[0], S, batch=batch, want_raw=True)
            extra = {"raw": o[2][None], "occ": o[2][None, :, 3:4].contiguous()}
            if self.output_device is not None:
                extra = {k: v.detach().to(self.output_device) for k, v in extra.items()}
            return extra
        return RenderOutput(ret, produce)

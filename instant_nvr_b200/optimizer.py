"""Drop-in for the optimizer ``lib/train/optimizer.py:13-31`` builds: ``torch.optim.Adam`` with one param group per
tensor, ``eps = cfg.train.eps`` (1e-15 in the shipped configs), amsgrad off -- as ONE pass over HBM
(``nvr_adam_step``: 28 B per element) instead of the library's multi-kernel foreach update.

``FusedAdam`` keeps torch's ``param_groups`` / ``state`` layout (``step`` as a CPU fp32 scalar tensor, ``exp_avg``,
``exp_avg_sq``), so ``state_dict()`` / ``load_state_dict()`` interchange with ``torch.optim.Adam`` checkpoints
(the reference's ``save_model`` / ``load_model`` store ``optim.state_dict()``, ``lib/utils/net_utils.py:423-460``)
and ``lr_scheduler``s work unchanged.  There is no CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch

from . import cabi


_HANDLES: Dict[int, "C.c_void_p"] = {}


def aux_handle(device: torch.device):
    """A library handle for the entry points that need no bound network (optimizer step, camera rays, image
    metrics); one per device, kept for the life of the process."""
    if not torch.cuda.is_available():
        raise RuntimeError("instant_nvr_b200 needs a CUDA device: there is no CPU implementation")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _HANDLES:
        lib = cabi.load()
        conf = cabi.NvrConfig(cabi.ABI_VERSION, idx, 0.05, 1)
        h = C.c_void_p()
        rc = lib.nvr_create(C.byref(conf), C.byref(h))
        if rc != 0 or not h.value:
            raise RuntimeError(f"nvr_create failed with code {rc}")
        _HANDLES[idx] = h
    return cabi.load(), _HANDLES[idx]


def check(lib, h, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.nvr_last_error(h)
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics (L2 weight decay added to the gradient,
    bias-corrected, amsgrad=False, maximize=False), fp32 contiguous CUDA parameters only."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 zero_grad_in_step: bool = False):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        # the keys torch.optim.Adam puts into a param group, so state_dicts interchange
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.zero_grad_in_step = bool(zero_grad_in_step)   # clear .grad in the same pass (keeps the buffers)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # one NvrAdamTensor per parameter with a gradient, grouped by (beta1, beta2, eps).  The ctypes arrays are cached between
        # steps and only re-built when a pointer moved (a new gradient buffer, a re-loaded state): per step the host then just
        # bumps `step` / `lr` in place instead of constructing 67 structs
        entries, touched = [], []
        device = None
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize") or group.get("decoupled_weight_decay"):
                raise RuntimeError("FusedAdam implements plain Adam only (amsgrad / maximize / AdamW are not on the reference's path)")
            hyper = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors (there is no CPU path)")
                if g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                    raise RuntimeError("FusedAdam: gradients must be dense contiguous fp32 on the parameter's device")
                if device is None:
                    device = p.device
                elif device != p.device:
                    raise RuntimeError("FusedAdam: all parameters must be on one device")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                entries.append((hyper, p, g, st, float(group["lr"]), float(group["weight_decay"])))
                touched += [p, st["exp_avg"], st["exp_avg_sq"]] + ([g] if self.zero_grad_in_step else [])
        if device is None:
            return loss
        key = tuple((e[0], e[1].data_ptr(), e[2].data_ptr(), e[3]["exp_avg"].data_ptr(), e[3]["exp_avg_sq"].data_ptr(), e[1].numel())
                    for e in entries)
        cache = getattr(self, "_pack", None)
        if cache is None or cache["key"] != key:
            by_hyper: Dict[tuple, List[int]] = {}
            for i, e in enumerate(entries):
                by_hyper.setdefault(e[0], []).append(i)
            arrays = []
            for hyper, idxs in by_hyper.items():
                arr = (cabi.NvrAdamTensor * len(idxs))()
                for j, i in enumerate(idxs):
                    _, p, g, st, _, _ = entries[i]
                    arr[j].param, arr[j].grad = p.data_ptr(), g.data_ptr()
                    arr[j].exp_avg, arr[j].exp_avg_sq, arr[j].numel = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
                arrays.append((hyper, idxs, arr))
            cache = self._pack = {"key": key, "arrays": arrays}
        lib, h = aux_handle(device)
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            for (beta1, beta2, eps), idxs, arr in cache["arrays"]:
                for j, i in enumerate(idxs):
                    _, _, _, st, lr, wd = entries[i]
                    arr[j].step, arr[j].lr, arr[j].weight_decay = int(st["step"]), lr, wd
                check(lib, h, lib.nvr_adam_step(h, arr, len(idxs), beta1, beta2, eps, int(self.zero_grad_in_step), stream),
                      "nvr_adam_step")
        # the library wrote through raw pointers: tell autograd / anything keyed on tensor versions (the engine's
        # pre-summed inference tables) that these tensors changed in place, as torch.optim.Adam's in-place ops would
        torch.autograd.graph.increment_version(touched)
        return loss


def make_optimizer(cfg, net, lr=None, weight_decay=None) -> FusedAdam:
    """``lib/train/optimizer.py:13-31`` with Adam replaced by the fused step: the same per-tensor param groups (names
    containing 'data' get ``lr``, all others ``lr * cfg.mlp_weight_decay``), same eps / weight decay."""
    lr = cfg.train.lr if lr is None else lr
    weight_decay = cfg.train.weight_decay if weight_decay is None else weight_decay
    if "adam" not in cfg.train.optim or cfg.train.optim != "adam":
        raise ValueError("instant_nvr_b200.optimizer covers cfg.train.optim == 'adam' (the shipped configs); "
                         "radam / sgd stay the reference's")
    groups = []
    for key, value in net.named_parameters():
        if not value.requires_grad:
            continue
        scale = 1.0 if "data" in key else cfg.mlp_weight_decay
        groups.append({"params": [value], "lr": lr * scale, "weight_decay": weight_decay})
    return FusedAdam(groups, lr, weight_decay=weight_decay, eps=cfg.train.eps)

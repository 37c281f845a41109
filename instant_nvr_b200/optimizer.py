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

    # ---- per-step host work ------------------------------------------------------------------------
    # A step of the reference's optimizer touches 67 tensors.  Walking them the obvious way (a CPU tensor `step += 1` and
    # `int(step)` per tensor, 67 ctypes structs rebuilt, a 335-pointer cache key) cost 0.55 ms of host time per step, all of
    # it on the training step's critical path (the Adam launch waits for it).  The plan below is built once and revalidated
    # per step with two pointer reads per tensor; `step` is bumped through a numpy view of the state's own 0-d tensor.
    class _Entry:
        __slots__ = ("p", "group", "st", "m", "v", "step_t", "step_np", "pptr", "gptr", "a", "lr", "wd", "hyper")

    def _step_view(self, st):
        t = st["step"]
        try:
            return t, (t.numpy() if (t.device.type == "cpu" and t.dtype == torch.float32 and t.dim() == 0) else None)
        except Exception:
            return t, None

    def _build_plan(self):
        entries, device = [], None
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize") or group.get("decoupled_weight_decay"):
                raise RuntimeError("FusedAdam implements plain Adam only (amsgrad / maximize / AdamW are not on the reference's path)")
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
                e = FusedAdam._Entry()
                e.p, e.group, e.st, e.m, e.v = p, group, st, st["exp_avg"], st["exp_avg_sq"]
                e.step_t, e.step_np = self._step_view(st)
                e.pptr, e.gptr = p.data_ptr(), g.data_ptr()
                e.lr = e.wd = None
                entries.append(e)
        by_hyper: Dict[tuple, List] = {}
        for e in entries:
            e.hyper = (float(e.group["betas"][0]), float(e.group["betas"][1]), float(e.group["eps"]))
            by_hyper.setdefault(e.hyper, []).append(e)
        arrays = []
        for hyper, es in by_hyper.items():
            arr = (cabi.NvrAdamTensor * len(es))()
            for j, e in enumerate(es):
                e.a = arr[j]                     # a view of the array element: field writes go straight into `arr`
                arr[j].param, arr[j].grad = e.pptr, e.gptr
                arr[j].exp_avg, arr[j].exp_avg_sq, arr[j].numel = e.m.data_ptr(), e.v.data_ptr(), e.p.numel()
            arrays.append((hyper, arr, len(es)))
        n_params = sum(len(g["params"]) for g in self.param_groups)
        versioned = [t for e in entries for t in (e.p, e.m, e.v)]
        # (no reference to the gradient tensors is kept: zero_grad(set_to_none=True) must be able to free them)
        return {"entries": entries, "arrays": arrays, "device": device, "n_params": n_params, "versioned": versioned}

    def _plan_valid(self, plan) -> bool:
        n = 0
        for group in self.param_groups:
            n += len(group["params"])
        if n != plan["n_params"]:
            return False
        seen = 0
        for e in plan["entries"]:
            g = e.p.grad
            st = e.st
            b = e.group["betas"]
            if (b[0], b[1], e.group["eps"]) != e.hyper:
                return False
            if g is None:
                return False
            gp = g.data_ptr()
            if gp != e.gptr:             # a new gradient buffer (the allocator handed the backward another block): patch the pointer
                if g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous() or g.device != e.p.device:
                    return False         # rebuilt -> _build_plan raises the proper error
                e.a.grad = e.gptr = gp
            if (e.p.data_ptr() != e.pptr or st.get("exp_avg") is not e.m
                    or st.get("exp_avg_sq") is not e.v or st.get("step") is not e.step_t or self.state.get(e.p) is not st):
                return False
            seen += 1
        if seen != n:                    # a parameter without a gradient last time may have one now
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is not None:
                        seen -= 1
            return seen == 0
        return True

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plan = getattr(self, "_plan", None)
        if plan is None or not self._plan_valid(plan):
            plan = self._plan = self._build_plan()
            self.plan_builds = getattr(self, "plan_builds", 0) + 1
        if plan["device"] is None:
            return loss
        for e in plan["entries"]:
            if e.step_np is not None:
                e.step_np[()] += 1.0
                cnt = int(e.step_np)
            else:                        # a step tensor that is not a 0-d fp32 CPU tensor (e.g. a capturable checkpoint)
                e.step_t += 1
                cnt = int(e.step_t)
            a = e.a
            a.step = cnt
            lr, wd = e.group["lr"], e.group["weight_decay"]
            if lr != e.lr:
                a.lr = e.lr = float(lr)
            if wd != e.wd:
                a.weight_decay = e.wd = float(wd)
        device = plan["device"]
        lib, h = aux_handle(device)
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            for (beta1, beta2, eps), arr, n in plan["arrays"]:
                check(lib, h, lib.nvr_adam_step(h, arr, n, beta1, beta2, eps, int(self.zero_grad_in_step), stream), "nvr_adam_step")
        # the library wrote through raw pointers: tell autograd / anything keyed on tensor versions (the engine's
        # pre-summed inference tables) that these tensors changed in place, as torch.optim.Adam's in-place ops would
        torch.autograd.graph.increment_version(plan["versioned"] + ([e.p.grad for e in plan["entries"]] if self.zero_grad_in_step else []))
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

"""Training-mode surfaces of the drop-in Network / Renderer: autograd glue over the C-ABI training entry
points (``nvr_train_forward`` / ``nvr_train_backward`` / ``nvr_deformer_backward`` / ``nvr_composite_*``).

Mirrors what the reference's autograd gives for ``Network.forward`` with ``self.training``
(``inb_part_network_multiassign.py:126-168``), ``Network.resd`` (``:122-124``) and ``volume_rendering``
(``lib/utils/net_utils.py:12-44``): gradients reach the part grids, the part MLPs, the frame's latent row, the
deformer MLP and the deformer grid -- and nothing else (KNN weights / blended transforms / view directions are
constants of the frame, ``:85-90``).  All arithmetic is in the CUDA library; this file only moves pointers.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


def trainable(net) -> List[Tuple[str, torch.nn.Parameter]]:
    """The reference's trainable tensors in a fixed order (dense, hash, MLP weights/biases,
    rgb_latent; every other state_dict entry is a frozen buffer).  Cached on the module (a walk over the module tree costs
    ~0.1 ms and a training step asked for it six times); ``Network._apply`` drops the cache when parameters are replaced."""
    cached = getattr(net, "_trainable_cache", None)
    if cached is None:
        cached = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        try:
            net._trainable_cache = cached
        except Exception:
            pass
    return cached


_ZERO_PLANS: Dict[tuple, tuple] = {}


def zero_grads(params: Dict[str, torch.Tensor], names: List[str]) -> Dict[str, torch.Tensor]:
    """Fresh zero-filled gradient buffers for ``names``: ONE allocation and ONE fill launch (the 67 trainable tensors
    are views of it at 256-byte-aligned offsets) instead of one of each per tensor.  A new buffer every call: autograd
    takes ownership of what a backward returns (``.grad`` may alias it), so it must not be reused.  The returned dict
    carries the flat buffer's address under ``"_base"`` (engine: the gradient descriptor is cached per address)."""
    key = (id(params), len(names))
    plan = _ZERO_PLANS.get(key, None)
    if plan is None or plan[0] != tuple(names):
        offs, total = [], 0
        for name in names:
            offs.append(total)
            total += (params[name].numel() + 63) & ~63
        plan = (tuple(names), total, [(name, tuple(params[name].shape), params[name].stride(), o) for name, o in zip(names, offs)])
        if len(_ZERO_PLANS) > 8:
            _ZERO_PLANS.clear()
        _ZERO_PLANS[key] = plan
    total = plan[1]
    ref = params[names[0]]
    flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
    out = {name: flat.as_strided(shape, stride, off) for name, shape, stride, off in plan[2]}     # one view op per tensor
    out["_base"] = (flat.data_ptr(), total, len(names))
    return out


def _param_dict(net) -> Dict[str, torch.Tensor]:
    cached = getattr(net, "_trainable_dict", None)
    if cached is None or getattr(net, "_trainable_dict_src", None) is not trainable(net):
        cached = dict(trainable(net))
        net._trainable_dict, net._trainable_dict_src = cached, trainable(net)
    return cached


class _NetworkTrainFn(torch.autograd.Function):
    """raw (N,4), occ (N,1), resd (Ns,5,3), tocc (Ns,5) [differentiable]; x0 (Ns,5,3) [constant].  Survivor rows come out of
    the library in the reference's order already (ascending sample index, nvr_train_forward)."""

    @staticmethod
    def forward(ctx, net, batch, wpts, viewdir, *params):
        eng = net.engine()
        state = eng.train_forward(wpts, viewdir, batch)
        ctx.net, ctx.state = net, state
        ctx.names = [n for n, _ in trainable(net)]
        ns = state["n_surv"]
        outs = (state["raw"], state["occ"][:, None], state["resd"][:ns], state["tocc"][:ns], state["x0"][:ns])
        ctx.mark_non_differentiable(outs[4])
        return outs

    @staticmethod
    def backward(ctx, d_raw, d_occ, d_resd, d_tocc, _dx0):
        net, state = ctx.net, ctx.state
        eng = net.engine()
        n = state["raw"].shape[0]
        g_raw = torch.zeros(n, 4, dtype=torch.float32, device=state["raw"].device) if d_raw is None else d_raw.contiguous()
        if d_occ is not None:
            g_raw = g_raw.clone() if d_raw is not None else g_raw
            g_raw[:, 3] += d_occ[:, 0]                       # 'occ' is the 4th column of the fused raw (:254-255)
        grads = zero_grads(_param_dict(net), ctx.names)
        eng.train_backward(state, g_raw, None if d_resd is None else d_resd.contiguous(),
                           None if d_tocc is None else d_tocc.contiguous(), net, grads)
        ctx.state = None
        return (None, None, None, None) + tuple(grads[name] for name in ctx.names)


def network_train_forward(net, wpts: torch.Tensor, viewdir: torch.Tensor, batch: Dict) -> Dict[str, torch.Tensor]:
    """``Network.forward`` with ``self.training``: {'raw' (1,N,4), 'occ' (1,N,1), 'resd' (1,N',5,3),
    'tpts' (1,5N',3), 'tocc' (1,5N',1)} with survivors in the reference's order (ascending sample index)."""
    params = [p for _, p in trainable(net)]
    raw, occ, resd, tocc, x0 = _NetworkTrainFn.apply(net, batch, wpts, viewdir, *params)
    return {"raw": raw[None], "occ": occ[None], "resd": resd[None], "tpts": x0.reshape(1, -1, 3),
            "tocc": tocc.reshape(1, -1, 1)}


class _DeformerFn(torch.autograd.Function):
    """Network.resd on explicit canonical points; gradient to the deformer's parameters only (the points are
    constants of the step: the pair regulariser's jittered neighbours, inb_renderer.py:78-94)."""

    @staticmethod
    def forward(ctx, net, batch, tpts, *params):
        eng = net.engine()
        ctx.net, ctx.batch = net, batch
        ctx.names = [n for n, _ in trainable(net) if n.startswith("tpose_deformer.")]
        pts = tpts.detach().reshape(-1, 3).contiguous()
        ctx.save_for_backward(pts)
        return eng.deformer_residual(pts, batch)

    @staticmethod
    def backward(ctx, d_resd):
        (pts,) = ctx.saved_tensors
        net = ctx.net
        grads = zero_grads(_param_dict(net), ctx.names)
        net.engine().deformer_backward(pts, d_resd.contiguous(), ctx.batch, net, grads)
        return (None, None, None) + tuple(grads[name] for name in ctx.names)


def deformer_train(net, tpts: torch.Tensor, batch: Dict) -> torch.Tensor:
    B, N, D = tpts.shape
    params = [p for n, p in trainable(net) if n.startswith("tpose_deformer.")]
    return _DeformerFn.apply(net, batch, tpts, *params).view(B, N, D)


class _CompositeFn(torch.autograd.Function):
    """volume_rendering(rgb, occ, epsilon=0): raw (R,S,4) -> weights (R,S), rgb_map (R,3), acc_map (R)."""

    @staticmethod
    def forward(ctx, eng, raw):
        raw = raw.contiguous()
        ctx.eng = eng
        ctx.save_for_backward(raw)
        return eng.composite_forward(raw)

    @staticmethod
    def backward(ctx, d_weights, d_rgb_map, d_acc_map):
        (raw,) = ctx.saved_tensors
        return None, ctx.eng.composite_backward(raw, d_weights, d_rgb_map, d_acc_map)


def composite(eng, raw: torch.Tensor):
    return _CompositeFn.apply(eng, raw)


class _DistortionFn(torch.autograd.Function):
    """reg_distortion_loss (inb_renderer.py:96-103): weights (R,S) [differentiable], z_vals (R,S) [constant] -> loss (R)."""

    @staticmethod
    def forward(ctx, eng, weights, z_vals):
        weights, z_vals = weights.contiguous(), z_vals.contiguous()
        ctx.eng = eng
        ctx.save_for_backward(weights, z_vals)
        return eng.distortion_forward(weights, z_vals)

    @staticmethod
    def backward(ctx, d_loss):
        weights, z_vals = ctx.saved_tensors
        return None, ctx.eng.distortion_backward(weights, z_vals, d_loss.contiguous()), None


def render_train(renderer, batch: Dict, epoch: int = -1) -> Dict[str, torch.Tensor]:
    """``Renderer.render`` with ``net.training`` (inb_renderer.py:53-239): stratified jitter, the network in
    training mode, compositing, pair and distortion regularisers.  Sampling-distance bookkeeping (linspace, the
    jitter draw) is the reference's own torch code; everything per sample runs in the CUDA library."""
    net = renderer.net
    cfg = net.cfg
    if epoch != -1:
        batch["epoch"] = epoch
    ray_o, ray_d, near, far = batch["ray_o"], batch["ray_d"], batch["near"], batch["far"]
    n_batch, n_pixel = ray_o.shape[:2]
    if n_batch != 1:
        raise ValueError("n_batch must be 1 (the reference asserts it, inb_part_network_multiassign.py:84)")
    S = cfg.N_samples
    eng = net.engine()
    # :15-31 in one launch; the stratified jitter keeps the reference's torch.rand draw (same generator, same shape)
    u = torch.rand(n_batch, n_pixel, S, device=near.device, dtype=near.dtype) if cfg.perturb > 0.0 else None
    z_vals, wpts, viewdir = eng.train_sample(ray_o[0], ray_d[0], near[0], far[0], S, None if u is None else u[0])
    ret = net(wpts, viewdir, None, batch)                        # `dists` (:45-47) is dead in the reference's networks

    raw = ret["raw"].reshape(n_pixel, S, 4)
    weights, rgb_map, acc_map = composite(eng, raw)                                             # :72
    if cfg.use_pair_reg:                                                                        # :78-94
        tocc = ret["tocc"].view(-1)
        reg_inds = ((tocc - 0.5).abs() < 0.02).nonzero(as_tuple=True)[0]
        if reg_inds.numel():
            reg_tpts = ret["tpts"].view(-1, 3)[reg_inds][None]
            reg_resd = ret["resd"].view(-1, 3)[reg_inds][None]
            neighbor = reg_tpts + (torch.rand_like(reg_tpts) - 0.5) * 0.01                     # compute_val_pair_around_range :40
            ret["oresd"] = torch.cat([reg_resd, net.resd(neighbor, batch)], dim=1)             # :45-46
        else:
            ret["oresd"] = torch.zeros(1, 0, 3, device=raw.device)
    if cfg.use_reg_distortion:                                                                  # :96-103
        ret["reg_distortion_loss"] = _DistortionFn.apply(eng, weights, z_vals)[None]
    ret.update({"rgb_map": rgb_map[None], "acc_map": acc_map[None], "raw": raw.reshape(1, -1, 4)})
    ret["resd"] = ret["resd"].reshape(n_batch, ret["tpts"].shape[1], 3)                         # :134-136: (1, 5N', 3) out of the renderer
    if cfg.use_freespace_loss:                                                                  # :118-121
        # the reference's local `occ` is, by then, the PREDICTED occupancy raw[..., 3] (it overwrites the argument, :71), so
        # the selection is "samples whose predicted occupancy is exactly 0" and the result is (1, K)
        occ_pred = raw[..., 3]
        ret["freespace_occupancy"] = occ_pred[occ_pred == 0][None]
    return ret

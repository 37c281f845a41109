"""Host-side engine: marshals PyTorch tensors into the C-ABI of ``libnvr_b200.so``.

PyTorch is plumbing here (device memory, the current CUDA stream); all arithmetic of the path runs
in the CUDA library.  Parameter and frame tensors are *borrowed*: the engine keeps Python
references so the storages outlive the binding, and re-binds when a storage pointer changes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import cabi
from .config import NUM_PARTS, PathConfig

_FRAME_KEYS = ("R", "Th", "pbw", "pbounds", "part_pts", "part_pbw", "lengths2", "A", "big_A", "tuv", "tbounds",
               "frame_dim", "latent_index")


def _stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


DEFAULT_MLP_MODE = 3
DEFAULT_TUNE = 0


def _params_struct(net, grads: Optional[Dict[str, torch.Tensor]] = None) -> "cabi.NvrParams":
    """NvrParams over the module tree.  With ``grads`` (parameter name -> gradient buffer) the pointer fields hold the
    gradient buffers instead (0 where absent), which is what nvr_train_backward / nvr_deformer_backward take.
    A gradient descriptor is cached per flat gradient buffer (``grads["_base"]``, training.zero_grads): the allocator hands
    the same block back step after step, so the ~1 ms walk over the module tree happens once."""
    base = None if grads is None else grads.get("_base")
    if base is not None:
        cache = net.__dict__.setdefault("_grad_struct_cache", {})
        hit = cache.get(base)
        if hit is not None and hit[0] is getattr(net, "_trainable_cache", None):
            return hit[1]
    names = {id(p): n for n, p in net.named_parameters()}

    def ptr(t):
        if grads is None:
            return t.data_ptr()
        g = grads.get(names[id(t)])
        return g.data_ptr() if g is not None else 0

    def lin(m):
        d = cabi.linear_desc(m)
        d.weight, d.bias = ptr(m.weight), ptr(m.bias)
        return d

    def grid(e):
        g = cabi.grid_desc(e)
        g.dense, g.hash = ptr(e.dense), ptr(e.hash)
        return g
    P = cabi.NvrParams()
    for i, part in enumerate(net.tpose_human.part_networks):
        d = P.part[i]
        d.grid = grid(part.embedder)
        for k, m in enumerate(part.occ.linears):
            d.occ[k] = lin(m)
        for k, m in enumerate(part.rgb.linears):
            d.rgb[k] = lin(m)
        d.n_rgb = len(part.rgb.linears)
        d.n_latent = part.rgb_latent.shape[0]
        d.rgb_latent = ptr(part.rgb_latent)
    P.deformer_grid = grid(net.tpose_deformer.embedder)
    for k, idx in enumerate((0, 2, 4)):
        P.deformer_mlp[k] = lin(net.tpose_deformer.mlp[idx])
    if base is not None:
        if len(cache) > 8:
            cache.clear()
        cache[base] = (getattr(net, "_trainable_cache", None), P)
    return P


class Engine:
    def __init__(self, cfg: PathConfig, device: Optional[torch.device] = None, max_points_per_pass: int = 32 << 20,
                 mlp_mode: Optional[int] = None, inference_tables: Optional[bool] = None, tune: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("instant_nvr_b200 needs a CUDA device: the hot path has no CPU implementation")
        cfg.check_supported()
        if mlp_mode is None:        # 1: tcgen05 3xTF32 tensor-core tiles (default); 0: fp32 FFMA tiles
            mlp_mode = int(os.environ.get("NVR_MLP_MODE", DEFAULT_MLP_MODE))
        self.mlp_mode = int(mlp_mode)
        # pre-summed inference tables (nvr_prepare_inference): opt-in, NVR_INFERENCE_TABLES=1 or the constructor flag
        self.inference_tables = bool(int(os.environ.get("NVR_INFERENCE_TABLES", "0"))) if inference_tables is None else bool(inference_tables)
        self._tables_key = None
        self.cfg = cfg
        self.lib = cabi.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.max_points_per_pass = int(os.environ.get("NVR_PASS_POINTS", max_points_per_pass))
        self.tune = int(os.environ.get("NVR_TUNE", DEFAULT_TUNE)) if tune is None else int(tune)   # NVR_TUNE_* bits of include/nvr_b200.h
        conf = cabi.NvrConfig(cabi.ABI_VERSION, self.device.index or 0, float(cfg.smpl_thresh), int(mlp_mode), self.tune, 0)
        h = C.c_void_p()
        rc = self.lib.nvr_create(C.byref(conf), C.byref(h))
        if rc != 0 or not h.value:
            raise RuntimeError(f"nvr_create failed with code {rc}")
        self._h = h
        self._params_key = None
        self._params_keep = None
        self._params_net = None
        self._frame_key = None
        self._frame_keep = None
        self._topo_src = None
        self._topo_key = 0
        self._ws: Optional[torch.Tensor] = None
        self._io: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.nvr_destroy(h)
            except Exception:
                pass
            self._h = None

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.nvr_last_error(self._h)
            raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")

    # ---- parameters --------------------------------------------------------------------------
    def invalidate_params(self) -> None:
        self._params_key = None

    def _refresh_inference_tables(self) -> None:
        """Re-sum the tables when they changed in place (optimizer steps bump ``_version``); eval entry points only."""
        if not self.inference_tables:
            return
        key = tuple((t.data_ptr(), t._version) for t in self._net_tables)
        if key != self._tables_key:
            self._check(self.lib.nvr_prepare_inference(self._h, 1, _stream_ptr(self.device)), "nvr_prepare_inference")
            self._tables_key = key

    def bind_params(self, net) -> None:
        """``net``: instant_nvr_b200.network.Network (or any module with the same tree)."""
        if self._params_key is not None and self._params_net is net:
            return                                   # bound; Network._apply() / invalidate_params() un-binds
        tensors = [p for p in net.parameters()]
        key = tuple(p.data_ptr() for p in tensors)
        if key == self._params_key:
            self._params_net = net
            return
        for name, p in net.named_parameters():
            if p.device != self.device:
                raise RuntimeError(f"parameter {name} is on {p.device}, engine is on {self.device} "
                                   "(move the module with .cuda(); there is no CPU path)")
            if not p.is_contiguous():
                raise RuntimeError(f"parameter {name} is not contiguous")
        P = _params_struct(net)
        self._check(self.lib.nvr_bind_params(self._h, C.byref(P)), "nvr_bind_params")
        self._params_key, self._params_keep, self._params_net = key, tensors, net
        self._tables_key = None
        self._net_tables = [t for part in net.tpose_human.part_networks for t in (part.embedder.dense, part.embedder.hash)]

    # ---- frame --------------------------------------------------------------------------------
    def bind_frame(self, batch: Dict, force: bool = False) -> None:
        """Per-frame preparation (distance volume, KD vertex clusters).  Cached on the identity of the frame
        tensors; ``force`` re-runs it (what a new frame costs)."""
        src = [batch[k] for k in _FRAME_KEYS]
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in src)
        if key == self._frame_key and not force:
            return
        dev = self.device
        t = {k: batch[k] for k in _FRAME_KEYS}
        for k in ("pbw", "part_pts", "part_pbw", "tuv", "A", "big_A", "R", "Th", "pbounds", "tbounds", "lengths2"):
            if t[k].shape[0] != 1:
                raise ValueError(f"batch['{k}'] must have a leading batch dim of 1 (reference asserts n_batch == 1)")
        keep = {
            "R": _dev_f32(t["R"][0], dev), "Th": _dev_f32(t["Th"][0].reshape(3), dev),
            "pbw": _dev_f32(t["pbw"][0], dev), "pbounds": _dev_f32(t["pbounds"][0], dev),
            "part_pts": _dev_f32(t["part_pts"][0], dev), "part_pbw": _dev_f32(t["part_pbw"][0], dev),
            "lengths2": t["lengths2"][0].to(device=dev, dtype=torch.int64).contiguous(),
            "A": _dev_f32(t["A"][0], dev), "big_A": _dev_f32(t["big_A"][0], dev),
            "tuv": _dev_f32(t["tuv"][0], dev), "tbounds": _dev_f32(t["tbounds"][0], dev),
            "frame_dim": _dev_f32(t["frame_dim"].reshape(-1)[:1], dev),
            "latent_index": t["latent_index"].reshape(-1)[:1].to(device=dev, dtype=torch.int64).contiguous(),
        }
        if keep["part_pts"].shape[0] != NUM_PARTS or keep["part_pbw"].shape[-1] != 24 or keep["tuv"].shape[-1] != 2:
            raise ValueError("part_pts must be (1,5,maxlen,3), part_pbw (1,5,maxlen,24), tuv (1,D,H,W,2)")
        F = cabi.NvrFrame()
        F.R, F.Th = keep["R"].data_ptr(), keep["Th"].data_ptr()
        F.pbw, F.pbounds = keep["pbw"].data_ptr(), keep["pbounds"].data_ptr()
        for a in range(3):
            F.pbw_dims[a] = keep["pbw"].shape[a]
            F.tuv_dims[a] = keep["tuv"].shape[a]
        F.pbw_channels = keep["pbw"].shape[3]
        F.part_pts, F.part_pbw = keep["part_pts"].data_ptr(), keep["part_pbw"].data_ptr()
        F.lengths2, F.maxlen = keep["lengths2"].data_ptr(), keep["part_pts"].shape[1]
        F.A, F.big_A = keep["A"].data_ptr(), keep["big_A"].data_ptr()
        F.tuv, F.tbounds = keep["tuv"].data_ptr(), keep["tbounds"].data_ptr()
        F.frame_dim, F.latent_index = keep["frame_dim"].data_ptr(), keep["latent_index"].data_ptr()
        # same subject / same part split => the KD vertex partition of the previous frame is reused
        lkey = (t["lengths2"].data_ptr(), t["lengths2"]._version)
        if lkey != self._topo_src:
            self._topo_src, self._topo_keep = lkey, t["lengths2"]
            self._topo_key = (hash((int(F.maxlen),) + tuple(int(v) for v in t["lengths2"][0].tolist())) & 0x7FFFFFFFFFFFFFFF) | 1
        F.topology_key = self._topo_key
        self._check(self.lib.nvr_bind_frame(self._h, C.byref(F), _stream_ptr(self.device)), "nvr_bind_frame")
        keep["_src"] = src          # the key holds the SOURCE tensors' addresses: keep them alive so no new tensor can reuse one
        self._frame_key, self._frame_keep = key, keep

    # ---- scratch --------------------------------------------------------------------------------
    def _workspace(self, n_points: int) -> Tuple[int, int]:
        n = max(64, min(int(n_points), self.max_points_per_pass))
        need = int(self.lib.nvr_workspace_bytes(self._h, n))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws.data_ptr(), self._ws.numel()

    # ---- entry points ----------------------------------------------------------------------------
    def query_points(self, wpts: torch.Tensor, viewdir: torch.Tensor, batch: Dict):
        """Network.forward (eval): (N,3),(N,3) -> raw (N,4), occ (N,1) on the device."""
        self.bind_frame(batch)
        self._refresh_inference_tables()
        wpts, viewdir = _dev_f32(wpts, self.device), _dev_f32(viewdir, self.device)
        n = wpts.shape[0]
        raw = torch.empty(n, 4, dtype=torch.float32, device=self.device)
        occ = torch.empty(n, 1, dtype=torch.float32, device=self.device)
        ws, ws_bytes = self._workspace(n)
        self._check(self.lib.nvr_query_points(self._h, wpts.data_ptr(), viewdir.data_ptr(), n, raw.data_ptr(),
                                              occ.data_ptr(), ws, ws_bytes, _stream_ptr(self.device)), "nvr_query_points")
        return raw, occ

    def render_rays(self, ray_o, ray_d, near, far, n_samples: int, batch: Optional[Dict] = None, want_raw: bool = False):
        """Renderer.render (eval): rays (R,3),(R,3),(R,),(R,) -> rgb_map (R,3), acc_map (R,) [, raw (R*S,4)]."""
        if batch is not None:
            self.bind_frame(batch)
        self._refresh_inference_tables()
        ray_o, ray_d = _dev_f32(ray_o, self.device), _dev_f32(ray_d, self.device)
        near, far = _dev_f32(near, self.device), _dev_f32(far, self.device)
        R = ray_o.shape[0]
        rgb = torch.empty(R, 3, dtype=torch.float32, device=self.device)
        acc = torch.empty(R, dtype=torch.float32, device=self.device)
        raw = torch.empty(R * n_samples, 4, dtype=torch.float32, device=self.device) if want_raw else None
        ws, ws_bytes = self._workspace(R * n_samples)
        self._check(self.lib.nvr_render_rays(self._h, ray_o.data_ptr(), ray_d.data_ptr(), near.data_ptr(), far.data_ptr(),
                                             R, int(n_samples), rgb.data_ptr(), acc.data_ptr(),
                                             raw.data_ptr() if want_raw else None, ws, ws_bytes, _stream_ptr(self.device)),
                    "nvr_render_rays")
        return (rgb, acc, raw) if want_raw else (rgb, acc)

    def render_rays_host(self, ray_o, ray_d, near, far, n_samples: int, rgb_out, acc_out) -> None:
        """Host (pinned) ray buffers in, host rgb_map / acc_map out; copies are part of the call."""
        self._refresh_inference_tables()
        for t in (ray_o, ray_d, near, far, rgb_out, acc_out):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("render_rays_host takes contiguous fp32 CPU tensors")
        R = ray_o.shape[0]
        if self._io is None or self._io.numel() < 12 * R:
            self._io = torch.empty(12 * R, dtype=torch.float32, device=self.device)
        ws, ws_bytes = self._workspace(R * n_samples)
        self._check(self.lib.nvr_render_rays_host(self._h, ray_o.data_ptr(), ray_d.data_ptr(), near.data_ptr(),
                                                  far.data_ptr(), R, int(n_samples), rgb_out.data_ptr(), acc_out.data_ptr(),
                                                  self._io.data_ptr(), ws, ws_bytes, _stream_ptr(self.device)), "nvr_render_rays_host")

    def deformer_residual(self, tpts: torch.Tensor, batch: Dict) -> torch.Tensor:
        self.bind_frame(batch)
        tpts = _dev_f32(tpts, self.device)
        out = torch.empty_like(tpts)
        self._check(self.lib.nvr_deformer_residual(self._h, tpts.data_ptr(), tpts.shape[0], out.data_ptr(), _stream_ptr(self.device)),
                    "nvr_deformer_residual")
        return out

    def embed_part(self, part: int, xyz: torch.Tensor) -> torch.Tensor:
        xyz = _dev_f32(xyz, self.device)
        out = torch.empty(xyz.shape[0], 19, dtype=torch.float32, device=self.device)
        self._check(self.lib.nvr_embed_part(self._h, int(part), xyz.data_ptr(), xyz.shape[0], out.data_ptr(), _stream_ptr(self.device)),
                    "nvr_embed_part")
        return out

    def part_mlp(self, part: int, emb: torch.Tensor, dirs: torch.Tensor, batch: Dict) -> torch.Tensor:
        """occ + rgb MLPs of one part on explicit (n,19) embeddings and (n,3) canonical view dirs -> raw (n,4)."""
        self.bind_frame(batch)
        n = emb.shape[0]
        e20 = torch.zeros(n, 20, dtype=torch.float32, device=self.device)
        e20[:, :19] = _dev_f32(emb, self.device)
        dirs = _dev_f32(dirs, self.device)
        raw = torch.empty(n, 4, dtype=torch.float32, device=self.device)
        ws, ws_bytes = self._workspace(n)
        self._check(self.lib.nvr_part_mlp(self._h, int(part), e20.data_ptr(), dirs.data_ptr(), n, raw.data_ptr(), ws, ws_bytes,
                                          _stream_ptr(self.device)), "nvr_part_mlp")
        return raw

    # ---- training ---------------------------------------------------------------------------------
    # The forward's workspace (43 MB at 1024 rays x 64) must live until its backward, the backward's scratch only for the call.
    # Allocating both per step made the caching allocator split and re-merge its big blocks around the 1.14 GB gradient buffer:
    # the gradient pointer moved every 2-3 steps (optimizer descriptors rebuilt) and now and then a step paid a cudaMalloc.
    # They are pooled here instead; a forward whose backward never runs just lets its workspace go with its state dict.
    def _pool_take(self, kind: str, nbytes: int) -> torch.Tensor:
        pool = self.__dict__.setdefault("_train_pool", {"ws": [], "scratch": []})[kind]
        stream = _stream_ptr(self.device)
        for i, (t, s) in enumerate(pool):
            if t.numel() == nbytes and s == stream:      # exact size: the backward derives the record stride from it
                pool.pop(i)
                return t
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def _pool_give(self, kind: str, t: torch.Tensor) -> None:
        pool = self.__dict__.setdefault("_train_pool", {"ws": [], "scratch": []})[kind]
        pool.append((t, _stream_ptr(self.device)))
        if len(pool) > 2:
            pool.pop(0)

    def train_forward(self, wpts: torch.Tensor, viewdir: torch.Tensor, batch: Dict) -> Dict:
        """nvr_train_forward on all points in one pass.  Returns the outputs plus the private workspace the
        backward needs (kept alive by the returned dict)."""
        self.bind_frame(batch)
        wpts, viewdir = _dev_f32(wpts, self.device), _dev_f32(viewdir, self.device)
        n = wpts.shape[0]
        dev, f32 = self.device, torch.float32
        st = {"raw": torch.empty(n, 4, dtype=f32, device=dev), "occ": torch.empty(n, dtype=f32, device=dev),
              "x0": torch.empty(n, 5, 3, dtype=f32, device=dev), "resd": torch.empty(n, 5, 3, dtype=f32, device=dev),
              "tocc": torch.empty(n, 5, dtype=f32, device=dev), "rank_of_slot": torch.empty(n, dtype=torch.int32, device=dev),
              "ws": self._pool_take("ws", int(self.lib.nvr_workspace_bytes(self._h, max(n, 64)))),
              "n": n, "batch": batch}
        self._check(self.lib.nvr_train_forward(self._h, wpts.data_ptr(), viewdir.data_ptr(), n, st["raw"].data_ptr(),
                                               st["occ"].data_ptr(), st["x0"].data_ptr(), st["resd"].data_ptr(),
                                               st["tocc"].data_ptr(), st["rank_of_slot"].data_ptr(), st["ws"].data_ptr(),
                                               st["ws"].numel(), _stream_ptr(self.device)), "nvr_train_forward")
        st["n_surv"] = int(self.counters()["n_survivors"]) if n else 0      # one host sync (the reference has seven)
        return st

    def train_backward(self, st: Dict, d_raw: torch.Tensor, d_resd, d_tocc, net, grads: Dict[str, torch.Tensor]) -> None:
        self.bind_params(net)
        self.bind_frame(st["batch"])
        n = st["n"]
        scratch = self._pool_take("scratch", int(self.lib.nvr_train_scratch_bytes(self._h, max(n, 64))))
        G = _params_struct(net, grads)
        d_raw = _dev_f32(d_raw, self.device)
        d_resd = None if d_resd is None else _dev_f32(d_resd, self.device)      # converted ONCE: the pointers below are theirs
        d_tocc = None if d_tocc is None else _dev_f32(d_tocc, self.device)
        ptr = lambda t: 0 if t is None else t.data_ptr()
        self._check(self.lib.nvr_train_backward(self._h, d_raw.data_ptr(), ptr(d_resd), ptr(d_tocc), st["x0"].data_ptr(),
                                                st["rank_of_slot"].data_ptr(), n,
                                                C.byref(G), st["ws"].data_ptr(), st["ws"].numel(), scratch.data_ptr(),
                                                scratch.numel(), _stream_ptr(self.device)), "nvr_train_backward")
        # both buffers are dead once the backward's kernels have run; later users are ordered behind them on the same stream
        self._pool_give("scratch", scratch)
        self._pool_give("ws", st.pop("ws"))

    def deformer_backward(self, tpts: torch.Tensor, d_resd: torch.Tensor, batch: Dict, net, grads: Dict[str, torch.Tensor]) -> None:
        self.bind_params(net)
        self.bind_frame(batch)
        tpts, d_resd = _dev_f32(tpts.reshape(-1, 3), self.device), _dev_f32(d_resd.reshape(-1, 3), self.device)
        G = _params_struct(net, grads)
        self._check(self.lib.nvr_deformer_backward(self._h, tpts.data_ptr(), d_resd.data_ptr(), tpts.shape[0], C.byref(G),
                                                   _stream_ptr(self.device)), "nvr_deformer_backward")

    def train_sample(self, ray_o, ray_d, near, far, n_samples: int, u: Optional[torch.Tensor]):
        """Renderer.get_wsampling_points in one launch: rays (R,3),(R,3),(R,),(R,) [+ the jitter draw u (R,S)] ->
        z_vals (R,S), wpts (R*S,3), viewdir (R*S,3)."""
        dev = self.device
        ray_o, ray_d, near, far = (_dev_f32(t, dev) for t in (ray_o, ray_d, near, far))
        u = None if u is None else _dev_f32(u, dev)
        R = ray_o.shape[0]
        z = torch.empty(R, n_samples, dtype=torch.float32, device=dev)
        wpts = torch.empty(R * n_samples, 3, dtype=torch.float32, device=dev)
        vd = torch.empty(R * n_samples, 3, dtype=torch.float32, device=dev)
        self._check(self.lib.nvr_train_sample(self._h, ray_o.data_ptr(), ray_d.data_ptr(), near.data_ptr(), far.data_ptr(),
                                              None if u is None else u.data_ptr(), R, int(n_samples), z.data_ptr(), wpts.data_ptr(),
                                              vd.data_ptr(), _stream_ptr(dev)), "nvr_train_sample")
        return z, wpts, vd

    def distortion_forward(self, weights: torch.Tensor, z_vals: torch.Tensor) -> torch.Tensor:
        R, S = weights.shape
        loss = torch.empty(R, dtype=torch.float32, device=self.device)
        self._check(self.lib.nvr_distortion_forward(self._h, weights.data_ptr(), z_vals.data_ptr(), R, S, loss.data_ptr(),
                                                    _stream_ptr(self.device)), "nvr_distortion_forward")
        return loss

    def distortion_backward(self, weights: torch.Tensor, z_vals: torch.Tensor, d_loss: torch.Tensor) -> torch.Tensor:
        R, S = weights.shape
        d_w = torch.empty(R, S, dtype=torch.float32, device=self.device)
        d_loss = _dev_f32(d_loss, self.device)
        self._check(self.lib.nvr_distortion_backward(self._h, weights.data_ptr(), z_vals.data_ptr(), d_loss.data_ptr(), R, S,
                                                     d_w.data_ptr(), _stream_ptr(self.device)), "nvr_distortion_backward")
        return d_w

    def composite_forward(self, raw: torch.Tensor):
        raw = _dev_f32(raw, self.device)
        R, S = raw.shape[:2]
        w = torch.empty(R, S, dtype=torch.float32, device=self.device)
        rgb = torch.empty(R, 3, dtype=torch.float32, device=self.device)
        acc = torch.empty(R, dtype=torch.float32, device=self.device)
        self._check(self.lib.nvr_composite_forward(self._h, raw.data_ptr(), R, S, w.data_ptr(), rgb.data_ptr(), acc.data_ptr(),
                                                   _stream_ptr(self.device)), "nvr_composite_forward")
        return w, rgb, acc

    def composite_backward(self, raw: torch.Tensor, d_w, d_rgb, d_acc) -> torch.Tensor:
        raw = _dev_f32(raw, self.device)
        R, S = raw.shape[:2]
        d_raw = torch.empty(R, S, 4, dtype=torch.float32, device=self.device)
        ts = [None if t is None else _dev_f32(t, self.device) for t in (d_w, d_rgb, d_acc)]
        self._check(self.lib.nvr_composite_backward(self._h, raw.data_ptr(), R, S, *[0 if t is None else t.data_ptr() for t in ts],
                                                    d_raw.data_ptr(), _stream_ptr(self.device)), "nvr_composite_backward")
        return d_raw

    def query_points_debug(self, wpts: torch.Tensor, viewdir: torch.Tensor, batch: Dict):
        """query_points + per-stage taps: raw (N,4), surv_of_sample (N,), warp (N,5,8)=[flag,x,y,z,vx,vy,vz,pdist]."""
        self.bind_frame(batch)
        wpts, viewdir = _dev_f32(wpts, self.device), _dev_f32(viewdir, self.device)
        n = wpts.shape[0]
        raw = torch.empty(n, 4, dtype=torch.float32, device=self.device)
        surv = torch.empty(n, dtype=torch.int32, device=self.device)
        warp = torch.empty(n, 5, 8, dtype=torch.float32, device=self.device)
        keep, self.max_points_per_pass = self.max_points_per_pass, max(self.max_points_per_pass, n)
        ws, ws_bytes = self._workspace(n)
        self.max_points_per_pass = keep
        self._check(self.lib.nvr_query_points_debug(self._h, wpts.data_ptr(), viewdir.data_ptr(), n, raw.data_ptr(),
                                                    surv.data_ptr(), warp.data_ptr(), ws, ws_bytes, _stream_ptr(self.device)),
                    "nvr_query_points_debug")
        return raw, surv, warp

    def profile(self, enable: bool) -> None:
        self._check(self.lib.nvr_profile(self._h, int(enable)), "nvr_profile")

    def profile_read(self) -> Dict:
        p = cabi.NvrStageProfile()
        self._check(self.lib.nvr_profile_read(self._h, C.byref(p)), "nvr_profile_read")
        return {"ms": dict(zip(cabi.STAGE_NAMES, list(p.ms))), "launches": dict(zip(cabi.STAGE_NAMES, list(p.launches))),
                "passes": p.passes, "survivors": p.survivors, "pairs": list(p.pairs), "far_pairs": list(p.far_pairs),
                "embed_part_ms": list(p.embed_part_ms), "mlp_part_ms": list(p.mlp_part_ms)}

    def gather_footprint(self):
        """Distinct 32-byte table sectors the last (single-pass) render's pair lists touch, per part (measurement aid)."""
        out = (C.c_int64 * NUM_PARTS)()
        if self._ws is None:
            raise RuntimeError("gather_footprint: render something first")
        self._check(self.lib.nvr_gather_footprint(self._h, self._ws.data_ptr(), self._ws.numel(), out, _stream_ptr(self.device)),
                    "nvr_gather_footprint")
        return list(out)

    def counters(self) -> Dict[str, int]:
        c = cabi.NvrCounters()
        self._check(self.lib.nvr_read_counters(self._h, C.byref(c), _stream_ptr(self.device)), "nvr_read_counters")
        return {"n_points": c.n_points, "n_survivors": c.n_survivors, "n_pairs": list(c.n_pairs),
                "n_far_pairs": list(c.n_far_pairs), "kernel_launches": c.kernel_launches, "n_passes": c.n_passes}

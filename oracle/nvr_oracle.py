"""CPU ORACLE for the instant-nvr per-ray hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product (``instant_nvr_b200``) never does:
it has no CPU path and fails loudly when the CUDA library is missing.

What it is: a from-scratch restatement, in plain fp32 torch-CPU tensor math, of the reference
algorithm (zju3dv/instant-nvr @ a6f4d68) for

    ray samples -> world->pose -> distance cull -> per-part K=4 NN blend weights -> LBS warp to
    canonical/big-pose -> UV-time deformer -> per-part dense+hashed grid -> occ / rgb MLPs ->
    arg-max part fusion -> alpha compositing.

Each function cites the reference file:line it follows.  The reference has NO tests / golden
vectors for this path; the oracle is pinned instead against outputs of the
reference's own modules executed in the build container under the stub harness
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``).

Third-party arithmetic not in the reference tree: ``pytorch3d.ops.knn.knn_points``
(pytorch3d == 0.7.2, docs/install.md:10) -- exact K nearest by squared L2 among the first
``lengths2[b]`` points; restated in ``knn_sq`` below (brute force).

All inputs are torch CPU tensors.  ``sd`` is a reference-layout ``state_dict`` (SURVEY.md
Appendix D); the grid geometry (levels, resolutions, table size) is read from the tensors in
it, so the oracle does not depend on the product's config code.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

PART_NAMES = ("body", "leg", "head", "larm", "rarm")   # lib/utils/blend_utils.py:17
NUM_PARTS = 5
PRIMES = (1, 19349663, 83492791)                        # lib/config/config.py:17
FP32_EPS = torch.finfo(torch.float32).eps


# --------------------------------------------------------------------------------------
# sampling + compositing                                         inb_renderer.py / net_utils.py
# --------------------------------------------------------------------------------------
def sample_along_rays(ray_o, ray_d, near, far, n_samples: int):
    """inb_renderer.py:15-31 (eval branch: no stratified jitter).
    ray_o, ray_d (R,3); near, far (R,) -> pts (R,S,3), z (R,S)."""
    t = torch.linspace(0.0, 1.0, steps=n_samples, dtype=near.dtype)
    z = near[..., None] * (1.0 - t) + far[..., None] * t
    pts = ray_o[:, None] + ray_d[:, None] * z[..., None]
    return pts, z


def composite(raw):
    """net_utils.py:12-44 as called from inb_renderer.py:69-72: epsilon == cfg.random_bg == 0.
    raw (R,S,4)=[rgb,alpha] -> weights (R,S), rgb_map (R,3), acc_map (R,)."""
    rgb, alpha = raw[..., :3], raw[..., 3]
    ones = alpha.new_ones(alpha.shape[0], 1)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 0.0], dim=-1), dim=-1)[..., :-1]
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, dim=-2)
    acc_map = torch.sum(weights, -1)
    return weights, rgb_map, acc_map


# --------------------------------------------------------------------------------------
# trilinear volume lookup == F.grid_sample(trilinear, border, align_corners=True)
# --------------------------------------------------------------------------------------
def sample_volume(vol, pts, bounds):
    """blend_utils.py:501-525 / 528-555.  vol (D,H,W,C) indexed by (x,y,z) of the point,
    pts (N,3), bounds (2,3) -> (N,C).

    Restates ATen's grid_sampler_3d (bilinear, padding border, align_corners=True): the point
    is normalised to [-1,1], un-normalised to voxel units ((g+1)/2*(size-1)), clipped to
    [0,size-1], and the 8 neighbours are blended with the corner products ATen uses."""
    D, H, W, C = vol.shape
    g = (pts - bounds[0]) / (bounds[1] - bounds[0])
    g = g * 2 - 1
    out = pts.new_zeros(pts.shape[0], C)
    # ATen's x <-> last grid coord <-> W; the reference flips xyz->zyx so point.x indexes D.
    coord = []
    for axis, size in ((0, D), (1, H), (2, W)):
        c = ((g[:, axis] + 1) / 2) * (size - 1)
        c = torch.clamp(c, min=0.0, max=float(size - 1))
        coord.append(c)
    ix, iy, iz = coord[2], coord[1], coord[0]          # ATen naming: x->W, y->H, z->D
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    # weights in ATen's order tnw, tne, tsw, tse, bnw, bne, bsw, bse
    corners = (
        (x0, y0, z0, (x1 - ix) * (y1 - iy) * (z1 - iz)),
        (x1, y0, z0, (ix - x0) * (y1 - iy) * (z1 - iz)),
        (x0, y1, z0, (x1 - ix) * (iy - y0) * (z1 - iz)),
        (x1, y1, z0, (ix - x0) * (iy - y0) * (z1 - iz)),
        (x0, y0, z1, (x1 - ix) * (y1 - iy) * (iz - z0)),
        (x1, y0, z1, (ix - x0) * (y1 - iy) * (iz - z0)),
        (x0, y1, z1, (x1 - ix) * (iy - y0) * (iz - z0)),
        (x1, y1, z1, (ix - x0) * (iy - y0) * (iz - z0)),
    )
    flat = vol.reshape(D * H * W, C)
    for cx, cy, cz, w in corners:
        inb = (cx <= W - 1) & (cy <= H - 1) & (cz <= D - 1)       # lower side is >= 0 after the clip
        lin = (cz.clamp(max=D - 1).long() * H + cy.clamp(max=H - 1).long()) * W + cx.clamp(max=W - 1).long()
        v = flat[lin]
        out = out + torch.where(inb[:, None], v * w[:, None], torch.zeros_like(v))
    return out


# --------------------------------------------------------------------------------------
# K=4 nearest posed vertices -> blend weights                     blend_utils.py:732-763,817-825
# --------------------------------------------------------------------------------------
def knn_sq(src, ref, length: int, K: int = 4, chunk: int = 2048):
    """pytorch3d knn_points semantics (squared L2, first ``length`` reference points only).
    src (N,3), ref (M,3) -> d2 (N,K) ascending, idx (N,K)."""
    ref = ref[:length]
    d2s, idxs = [], []
    for s in range(0, src.shape[0], chunk):
        p = src[s:s + chunk]
        d2 = ((p[:, None] - ref[None]) ** 2).sum(-1)
        d, i = d2.topk(K, dim=-1, largest=False)
        d2s.append(d)
        idxs.append(i)
    if not d2s:
        return src.new_zeros(0, K), torch.zeros(0, K, dtype=torch.long)
    return torch.cat(d2s), torch.cat(idxs)


def knn_blend_weights(pts, part_pts, part_pbw, lengths2, K: int = 4, eps: float = 1e-8, radius: float = 0.075):
    """blend_utils.py:741-763 + 817-825.  pts (N,3); part_pts (P,maxlen,3); part_pbw (P,maxlen,24);
    lengths2 (P,) -> bw (N,P,24), pdist (N,P)."""
    P = part_pts.shape[0]
    bws, pds = [], []
    for p in range(P):
        d2, idx = knn_sq(pts, part_pts[p], int(lengths2[p]), K)
        d = d2.sqrt()                                               # cast_knn_points :736
        w = (-d ** 2 / (2 * radius ** 2)).exp()                     # :746
        w = w / (w.sum(dim=-1, keepdim=True) + eps)                 # :747
        pd = (d * w).sum(-1)                                        # :748
        vals = part_pbw[p][idx]                                     # (N,K,24) :761
        bw = (vals * w[..., None]).sum(1)                           # :762
        bws.append(bw)
        pds.append(pd)
    return torch.stack(bws, 1), torch.stack(pds, 1)


# --------------------------------------------------------------------------------------
# LBS                                                              blend_utils.py:293-317,395-487
# --------------------------------------------------------------------------------------
def inverse_3x3(R, eps: float = FP32_EPS):
    """blend_utils.py:293-317: transposed cofactors / (det + fp32 eps).  R (...,3,3)."""
    def minor(i, j):
        r = [k for k in range(3) if k != i]
        c = [k for k in range(3) if k != j]
        return R[..., r[0], c[0]] * R[..., r[1], c[1]] - R[..., r[0], c[1]] * R[..., r[1], c[0]]
    m = [[minor(i, j) for j in range(3)] for i in range(3)]
    sign = [[1.0, -1.0, 1.0], [-1.0, 1.0, -1.0], [1.0, -1.0, 1.0]]
    det = R[..., 0, 0] * m[0][0] - R[..., 0, 1] * m[0][1] + R[..., 0, 2] * m[0][2]
    rows = []
    for a in range(3):
        rows.append(torch.stack([(m[b][a] * sign[b][a]) / (det + eps) for b in range(3)], -1))
    return torch.stack(rows, -2)


def lbs_to_bigpose(pts, dirs, bw, A, big_A):
    """inb_part_network_multiassign.py:92-106 with blend_utils.py:395-487.
    pts, dirs (M,3); bw (M,24); A, big_A (24,4,4) -> bigpose pts (M,3), bigpose dirs (M,3)."""
    A_bw = (bw @ A.reshape(24, 16)).reshape(-1, 4, 4)                      # :415
    R_inv = inverse_3x3(A_bw[:, :3, :3])                                   # :418
    big = (bw @ big_A.reshape(24, 16)).reshape(-1, 4, 4)                   # :402
    t = torch.sum(R_inv * (pts - A_bw[:, :3, 3])[:, None], dim=2)          # :433-436
    x0 = torch.sum(big[:, :3, :3] * t[:, None], dim=2) + big[:, :3, 3]     # :468-470
    td = torch.sum(R_inv * dirs[:, None], dim=2)                           # :453
    v = torch.sum(big[:, :3, :3] * td[:, None], dim=2)                     # :485-486
    return x0, v


# --------------------------------------------------------------------------------------
# multi-resolution dense + hashed grid                             part_base_embedder.py:106-174
# --------------------------------------------------------------------------------------
def grid_geometry(sd: Dict[str, torch.Tensor], prefix: str):
    res = [int(v) for v in sd[prefix + "entries_num"]]
    T = int(sd[prefix + "hash"].shape[1])
    start_hash = len(res)
    for i, r in enumerate(res):
        if r ** 3 > T:
            start_hash = i
            break
    return res, T, start_hash


def grid_embed(sd, prefix: str, xyz, sum_features: bool):
    """part_base_embedder.py:106-174.  xyz (N,3) -> (N, 3+L) (sum mode) or (N, 3+L*F) (concat)."""
    res, T, sh = grid_geometry(sd, prefix)
    L = len(res)
    bounds = sd[prefix + "bounds"]
    size = sd[prefix + "entries_size"]
    esum = sd[prefix + "entries_sum"]
    dense, hsh = sd[prefix + "dense"], sd[prefix + "hash"]
    offs = sd[prefix + "offsets"]
    N = xyz.shape[0]
    u = (xyz - bounds[0]) / (bounds[1] - bounds[0])                         # :112
    outs = []
    for l in range(L):
        f = u / size[l]                                                     # :115
        i = (f[:, None] + offs[None]).long()                                # :116 trunc toward zero
        i = i.clip(0, res[l] - 1)                                           # :117
        o = f - i[:, 0]                                                     # :118 (clamped corner 0)
        if l < sh:
            idx = i[..., 0] * (res[l] ** 2) + i[..., 1] * res[l] + i[..., 2]   # :124-127
            if l > 0:
                idx = idx + int(esum[l - 1])                                # :129
            val = dense[idx]                                                # (N,8,F)
        else:
            idx = (i[..., 0] * PRIMES[0] ^ i[..., 1] * PRIMES[1] ^ i[..., 2] * PRIMES[2]) % T   # :132-136
            val = hsh[l - sh][idx]
        mul = (1 - offs[None]) + (2 * offs[None] - 1.0) * o[:, None]        # :158
        mul = mul[..., 0] * mul[..., 1] * mul[..., 2]                       # :159
        outs.append((mul[..., None] * val).sum(dim=-2))                     # :160  (N,F)
    val = torch.stack(outs, 1)                                              # (N,L,F)
    if sum_features:
        val = val.sum(dim=-1)                                               # :165
    else:
        val = val.reshape(N, -1)                                            # :169
    return torch.cat([u, val], dim=-1)                                      # :173


def grid_rows(sd, prefix: str, xyz):
    """The table rows grid_embed reads for xyz (N,3): (N, L, 8) int64, dense rows numbered [0, D) and the rows of hashed level
    l as D + (l - start_hash) * T + idx (D = total dense rows).  Same index arithmetic as grid_embed (:112-136); used by the
    tests of the gather-footprint measurement aid."""
    res, T, sh = grid_geometry(sd, prefix)
    bounds, size, esum, offs = sd[prefix + "bounds"], sd[prefix + "entries_size"], sd[prefix + "entries_sum"], sd[prefix + "offsets"]
    D = int(sd[prefix + "dense"].shape[0])
    u = (xyz - bounds[0]) / (bounds[1] - bounds[0])
    rows = []
    for l in range(len(res)):
        f = u / size[l]
        i = (f[:, None] + offs[None]).long().clip(0, res[l] - 1)
        if l < sh:
            idx = i[..., 0] * (res[l] ** 2) + i[..., 1] * res[l] + i[..., 2]
            if l > 0:
                idx = idx + int(esum[l - 1])
        else:
            idx = (i[..., 0] * PRIMES[0] ^ i[..., 1] * PRIMES[1] ^ i[..., 2] * PRIMES[2]) % T + D + (l - sh) * T
        rows.append(idx)
    return torch.stack(rows, 1)


# --------------------------------------------------------------------------------------
# MLPs, view-direction encoding, deformer, part network
# --------------------------------------------------------------------------------------
def _linear(x, sd, prefix):
    return F.linear(x, sd[prefix + "weight"], sd[prefix + "bias"])


def mlp_softplus(x, sd, prefix: str):
    """part_base_network.py:11-24: Linear -> Softplus ... -> Linear (prefix 'a.b.linears.')."""
    n = 0
    while f"{prefix}{n}.weight" in sd:
        n += 1
    for i in range(n - 1):
        x = F.softplus(_linear(x, sd, f"{prefix}{i}."))
    return _linear(x, sd, f"{prefix}{n - 1}.")


def posenc(x, multires: int = 4):
    """freq_embedder.py:20-31: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^3 x), cos(2^3 x)]."""
    out = [x]
    for k in range(multires):
        fx = x * (2.0 ** k)
        out += [torch.sin(fx), torch.cos(fx)]
    return torch.cat(out, dim=-1)


def deformer(sd, x0, tuv, tbounds, frame_dim):
    """uv_deformer.py:23-45 on already-selected points x0 (M,3) -> residual (M,3)."""
    uv = sample_volume(tuv, x0, tbounds)                                    # :32 (M,2)
    t = frame_dim.reshape(1, 1).expand(uv.shape[0], 1).float()              # :35
    uvt = torch.cat([uv, t], dim=-1)
    feat = grid_embed(sd, "tpose_deformer.embedder.", uvt, sum_features=False)
    h = F.softplus(_linear(feat, sd, "tpose_deformer.mlp.0."))
    h = F.softplus(_linear(h, sd, "tpose_deformer.mlp.2."))
    return 0.05 * torch.tanh(_linear(h, sd, "tpose_deformer.mlp.4."))       # :39


def part_network(sd, pid: int, x, v, latent_index: int):
    """part_base_network.py:44-63.  x, v (M,3) -> raw (M,4) = [rgb, occ]."""
    pre = f"tpose_human.part_networks.{pid}."
    e = grid_embed(sd, pre + "embedder.", x, sum_features=True)             # :49
    h = mlp_softplus(e, sd, pre + "occ.linears.")                           # :50
    occ = 1 - torch.exp(-F.softplus(h[..., :1]))                            # :51
    feat = h[..., 1:]
    ed = posenc(v)                                                          # :54
    lat = sd[pre + "rgb_latent"][latent_index][None].expand(x.shape[0], -1)   # :55
    inp = torch.cat([e, ed, feat, lat], dim=-1)                             # :56
    rgb = mlp_softplus(inp, sd, pre + "rgb.linears.").sigmoid()             # :57-58
    return torch.cat([rgb, occ], dim=-1)


# --------------------------------------------------------------------------------------
# Network.forward (eval)                                inb_part_network_multiassign.py:126-168
# --------------------------------------------------------------------------------------
def network_forward(sd, wpts, viewdir, batch, smpl_thresh: float, want_stages: bool = False):
    """wpts, viewdir (N,3) -> raw (N,4), occ (N,1)  [+ dict of per-stage tensors].

    ``batch`` holds the reference batch tensors WITHOUT the leading batch dim of 1:
    R (3,3), Th (1,3), pbw (D,H,W,25), pbounds (2,3), part_pts (P,maxlen,3),
    part_pbw (P,maxlen,24), lengths2 (P,), A, big_A (24,4,4), tuv (D,H,W,2), tbounds (2,3),
    frame_dim (1,), latent_index (1,)."""
    N = wpts.shape[0]
    pose_pts = torch.matmul(wpts - batch["Th"], batch["R"])                 # blend_utils.py:372
    pose_dirs = torch.matmul(viewdir, batch["R"])                           # :381
    pnorm = sample_volume(batch["pbw"][..., -1:], pose_pts, batch["pbounds"])[:, 0]   # :134-135
    pind = (pnorm < smpl_thresh).nonzero(as_tuple=True)[0]                  # :136-137
    p, d = pose_pts[pind], pose_dirs[pind]
    M = p.shape[0]

    bw, pdist = knn_blend_weights(p, batch["part_pts"], batch["part_pbw"], batch["lengths2"])   # :88
    flag = pdist < smpl_thresh                                              # :90  (M,P)
    pe = p[:, None].expand(M, NUM_PARTS, 3).reshape(-1, 3)
    de = d[:, None].expand(M, NUM_PARTS, 3).reshape(-1, 3)
    x0, vdir = lbs_to_bigpose(pe, de, bw.reshape(-1, 24), batch["A"], batch["big_A"])   # :98-106
    fl = flag.reshape(-1)
    resd = torch.zeros_like(x0)
    sel = fl.nonzero(as_tuple=True)[0]
    resd[sel] = deformer(sd, x0[sel], batch["tuv"], batch["tbounds"], batch["frame_dim"])   # :112
    x = (x0 + resd).reshape(M, NUM_PARTS, 3)                                # :113
    vdir = vdir.reshape(M, NUM_PARTS, 3)

    raws = torch.zeros(M, NUM_PARTS, 4)                                     # :201-202
    lat = int(batch["latent_index"].reshape(-1)[0])
    for pid in range(NUM_PARTS):
        idx = flag[:, pid].nonzero(as_tuple=True)[0]                        # :206-209
        if idx.numel():
            raws[idx, pid] = part_network(sd, pid, x[idx, pid], vdir[idx, pid], lat)
    occs = raws[..., 3:]
    best = occs.argmax(dim=1).reshape(M, 1, 1).expand(M, 1, 4)              # :253
    raw_s = torch.gather(raws, 1, best)[:, 0]                               # :254
    occ_s = occs.max(dim=1)[0]                                              # :255

    raw = torch.zeros(N, 4)
    occ = torch.zeros(N, 1)
    raw[pind] = raw_s                                                       # :156-157
    occ[pind] = occ_s
    if not want_stages:
        return raw, occ
    stages = dict(pose_pts=pose_pts, pose_dirs=pose_dirs, pnorm=pnorm, pind=pind, bw=bw, pdist=pdist,
                  flag=flag, bigpose=x0.reshape(M, NUM_PARTS, 3), resd=resd.reshape(M, NUM_PARTS, 3),
                  tpose=x, tdirs=vdir, raws=raws)
    return raw, occ, stages


def strip_batch(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Drop the leading batch dim of 1 from the reference-style batch dict."""
    keys = ("R", "Th", "pbw", "pbounds", "part_pts", "part_pbw", "lengths2", "A", "big_A", "tuv",
            "tbounds", "ray_o", "ray_d", "near", "far")
    out = {k: batch[k][0] for k in keys if k in batch}
    out["frame_dim"] = batch["frame_dim"].reshape(-1).float()
    out["latent_index"] = batch["latent_index"].reshape(-1)
    return out


def render(sd, batch, n_samples: int, smpl_thresh: float, chunk: int = 4096, want_raw: bool = True):
    """Renderer.render, eval branch (inb_renderer.py:204-239): 4096-ray chunks.
    ``batch`` in reference layout (leading batch dim 1).  Returns dict of CPU tensors with the
    reference's shapes: rgb_map (1,R,3), acc_map (1,R), raw (1,R*S,4), occ (1,R*S,1)."""
    b = strip_batch(batch)
    R = b["ray_o"].shape[0]
    outs = {"rgb_map": [], "acc_map": [], "raw": [], "occ": []}
    for s in range(0, R, chunk):
        sl = slice(s, s + chunk)
        pts, _ = sample_along_rays(b["ray_o"][sl], b["ray_d"][sl], b["near"][sl], b["far"][sl], n_samples)
        r, S = pts.shape[:2]
        vd = b["ray_d"][sl][:, None].expand(r, S, 3).reshape(-1, 3)
        raw, occ = network_forward(sd, pts.reshape(-1, 3), vd, b, smpl_thresh)
        _, rgb_map, acc_map = composite(raw.reshape(r, S, 4))
        outs["rgb_map"].append(rgb_map)
        outs["acc_map"].append(acc_map)
        if want_raw:
            outs["raw"].append(raw)
            outs["occ"].append(occ)
    ret = {"rgb_map": torch.cat(outs["rgb_map"])[None], "acc_map": torch.cat(outs["acc_map"])[None]}
    if want_raw:
        ret["raw"] = torch.cat(outs["raw"])[None]
        ret["occ"] = torch.cat(outs["occ"])[None]
    return ret


def render_train(sd, batch, n_samples: int, smpl_thresh: float, use_pair_reg: bool = True,
                 use_reg_distortion: bool = True, pair_noise: Optional[torch.Tensor] = None):
    """Renderer.render with ``net.training`` and cfg.perturb == 0 (inb_renderer.py:53-125, 204-239; one chunk) on top of
    Network.forward's training outputs (inb_part_network_multiassign.py:161-165).  Plain torch math: ``backward()`` on the
    result is the reference's gradient flow when ``sd`` holds leaf tensors with requires_grad.

    Returns the reference's keys: rgb_map (1,R,3), acc_map (1,R), raw (1,R*S,4), occ (1,R*S,1), resd (1,M,5,3),
    tpts (1,5M,3), tocc (1,5M,1), oresd (1,2K,3) [pair regulariser, :78-94], reg_distortion_loss (1,R) [:96-103].
    ``pair_noise`` (K,3) stands for the ``torch.rand_like`` draw of compute_val_pair_around_range
    (inb_part_network_multiassign.py:40); when None it is drawn here from the global generator, as the reference does."""
    b = strip_batch(batch)
    pts, z = sample_along_rays(b["ray_o"], b["ray_d"], b["near"], b["far"], n_samples)           # :15-31, perturb 0
    R, S = pts.shape[:2]
    vd = b["ray_d"][:, None].expand(R, S, 3).reshape(-1, 3)
    raw, occ, st = network_forward(sd, pts.reshape(-1, 3), vd, b, smpl_thresh, want_stages=True)
    M = st["pind"].shape[0]
    ret = {"raw": raw[None], "occ": occ[None], "resd": st["resd"][None], "tpts": st["bigpose"].reshape(1, -1, 3),
           "tocc": st["raws"][..., 3].reshape(1, -1, 1)}                                        # multiassign :161-165
    weights, rgb_map, acc_map = composite(raw.reshape(R, S, 4))                                  # :69-72
    if use_pair_reg:                                                                             # :78-94
        tocc = ret["tocc"].view(-1)
        reg = ((tocc - 0.5).abs() < 0.02).nonzero(as_tuple=True)[0]
        if reg.numel():
            reg_tpts = ret["tpts"].view(-1, 3)[reg]
            reg_resd = ret["resd"].view(-1, 3)[reg]
            noise = torch.rand_like(reg_tpts) if pair_noise is None else pair_noise
            neighbor = reg_tpts + (noise - 0.5) * 0.01                                           # multiassign :40
            nei = deformer(sd, neighbor, b["tuv"], b["tbounds"], b["frame_dim"])                  # Network.resd :122-124
            ret["oresd"] = torch.cat([reg_resd, nei], dim=0)[None]                                # :45-46
        else:
            ret["oresd"] = raw.new_zeros(1, 0, 3)
    if use_reg_distortion:                                                                       # :96-103
        ww = weights.reshape(R, S, 1) * weights.reshape(R, 1, S)
        nxt = torch.cat([z[:, 1:], z[:, -1:]], dim=-1)
        mid = (z + nxt) / 2
        diff = torch.abs(mid.reshape(R, S, 1) - mid.reshape(R, 1, S))
        ret["reg_distortion_loss"] = (ww * diff).sum(dim=-1).sum(dim=-1)[None]
    ret.update({"rgb_map": rgb_map[None], "acc_map": acc_map[None]})
    return ret


def psnr(img_pred, img_gt) -> float:
    """lib/evaluators/if_nerf.py:27-30: -10*ln(mse)/ln(10)."""
    mse = torch.mean((img_pred - img_gt) ** 2)
    return float(-10.0 * torch.log(mse) / torch.log(torch.tensor(10.0)))

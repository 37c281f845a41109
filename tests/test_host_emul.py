"""Per-sample device arithmetic (csrc/nvr_math.cuh) compiled for the host and checked against
the oracle -- runs without a GPU.

tests/host_emul/emul.cpp is a TEST-ONLY harness: it compiles the same ``__host__ __device__``
statements the CUDA kernels execute per sample with g++ (-ffp-contract=off) so index arithmetic,
trilinear / KNN / LBS / deformer math can be debugged in the CPU container.  It is not shipped,
not loaded by the product and not a fallback.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
import nvr_oracle as O  # noqa: E402
from instant_nvr_b200 import cabi  # noqa: E402

EMUL_DIR = os.path.join(REPO, "tests", "host_emul")


def build_emul():
    so = os.path.join(EMUL_DIR, "libnvr_emul.so")
    src = os.path.join(EMUL_DIR, "emul.cpp")
    deps = [src, os.path.join(REPO, "instant_nvr_b200", "csrc", "nvr_math.cuh"), os.path.join(REPO, "instant_nvr_b200", "csrc", "nvr_smpl.cuh"),
            os.path.join(REPO, "include", "nvr_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def emul():
    return build_emul()


def fp(t):
    return C.c_void_p(t.data_ptr())


def test_struct_sizes_match_header(emul):
    for which, cls in enumerate((cabi.NvrGrid, cabi.NvrLinear, cabi.NvrPart, cabi.NvrParams, cabi.NvrFrame,
                                 cabi.NvrConfig, cabi.NvrCounters, cabi.NvrStageProfile, cabi.NvrAdamTensor, cabi.NvrSmplPose,
                                 cabi.NvrSmplOut, cabi.NvrSmplPose)):
        assert emul.emul_sizeof(which) == C.sizeof(cls), cls.__name__


def test_barrett_mod(emul):
    rng = np.random.default_rng(0)
    for T in (257, 4099, 16411, 32771, 65537, 262147, 1048583, 2**31 - 1):
        ix, iy, iz = (rng.integers(0, 8192, 400000) for _ in range(3))
        h = (ix * 1) ^ (iy * 19349663) ^ (iz * 83492791)
        k = np.arange(1, 4000, dtype=np.int64)
        edges = np.concatenate([k * T - 1, k * T, k * T + 1, (2**40 // T - k) * T - 1, (2**40 // T - k) * T])
        h = np.concatenate([h, edges, [0, T - 1, T, T + 1, 2 * T - 1, 2**40 - 1, 2**40, 2**62 + 12345]]).astype(np.int64)
        emul.emul_check_mod.restype = C.c_longlong
        bad = emul.emul_check_mod(C.c_longlong(T), h.ctypes.data_as(C.c_void_p), C.c_longlong(len(h)))
        assert bad == 0, T


def _embed(emul, gp, x):
    g = cabi.grid_desc(gp)
    out = torch.empty(x.shape[0], gp.out_dim)
    emul.emul_embed(C.byref(g), fp(x), C.c_longlong(x.shape[0]), fp(out), C.c_int(gp.out_dim))
    return out


def _embed_f64(sd, prefix, x, sum_features):
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items() if k.startswith(prefix)}
    return O.grid_embed(sd64, prefix, x.double(), sum_features)


def test_grid_embed_matches_oracle(emul, golden_setup):
    """Inside the bbox the emulated device arithmetic and the oracle agree to ~1 ulp of the result.
    Outside it the reference extrapolates with corner weights of magnitude (1+|o|)^3 and opposite
    signs, so its own fp32 result is summation-order noise; there we only require that our error
    against exact (fp64) arithmetic is of the same size as the oracle's."""
    net, sd = golden_setup["net"], golden_setup["sd"]
    g = torch.Generator().manual_seed(3)
    for pid in range(5):
        gp = net.tpose_human.part_networks[pid].embedder
        pre = f"tpose_human.part_networks.{pid}.embedder."
        lo, hi = gp.bounds[0], gp.bounds[1]
        x = (lo + (hi - lo) * (torch.rand(700, 3, generator=g) * 1.4 - 0.2)).contiguous()
        x[:8] = torch.stack([lo, hi, lo - 1, hi + 1, (lo + hi) / 2, lo + 1e-7, hi - 1e-7, hi * 0 + 5.0])   # edges / far out
        ours = _embed(emul, gp, x)
        ref = O.grid_embed(sd, pre, x, True)
        u = (x - lo) / (hi - lo)
        inside = ((u >= 0) & (u <= 1)).all(1)
        assert inside.sum() > 100
        assert torch.allclose(ours[inside], ref[inside], atol=2e-6, rtol=1e-5), (pid, (ours - ref)[inside].abs().max())
        assert torch.equal(ours[:, :3], ref[:, :3])
        exact = _embed_f64(sd, pre, x, True)
        e_ours = (ours.double() - exact)[~inside].pow(2).mean().sqrt()
        e_ref = (ref.double() - exact)[~inside].pow(2).mean().sqrt()
        assert e_ours <= 3 * e_ref + 1e-6, (pid, float(e_ours), float(e_ref))
    gp = net.tpose_deformer.embedder
    x = (torch.rand(500, 3, generator=g) * 1.3 - 0.15).contiguous()
    ours = _embed(emul, gp, x)
    ref = O.grid_embed(sd, "tpose_deformer.embedder.", x, False)
    inside = ((x >= 0) & (x <= 1)).all(1)
    assert torch.allclose(ours[inside], ref[inside], atol=1e-6, rtol=1e-5), (ours - ref)[inside].abs().max()
    exact = _embed_f64(sd, "tpose_deformer.embedder.", x, False)
    e_ours = (ours.double() - exact)[~inside].pow(2).mean().sqrt()
    e_ref = (ref.double() - exact)[~inside].pow(2).mean().sqrt()
    assert e_ours <= 3 * e_ref + 1e-6


def test_full_size_hash_indices(emul):
    """The shipped body grid (T = 1048583, res up to 2005): int64 hash without wrap-around.  A small
    table cannot be built at that T, so the table is an index-revealing ramp: every row holds
    row_index / 16 in all 16 features, hence the level sum at a corner with weight 1 is the row."""
    from instant_nvr_b200.config import PathConfig
    spec = PathConfig.inb_377().parts[0].grid
    T, sh = spec.T, spec.start_hash
    res = spec.res
    # emulate one hashed level with a private table: rows = T, value = row id (exact in fp32 < 2^24)
    l = 15
    table = (torch.arange(T, dtype=torch.float32) / 16.0)[:, None].expand(T, 16).contiguous()
    g = cabi.NvrGrid()
    bounds = torch.tensor([[0., 0., 0.], [1., 1., 1.]])
    dense = torch.zeros(res[0] ** 3, 16)
    g.dense, g.hash, g.bounds = dense.data_ptr(), table.data_ptr(), bounds.data_ptr()
    g.n_levels, g.n_feat, g.start_hash, g.sum_features, g.table_size = 2, 16, 1, 1, T
    g.res[0], g.size[0], g.dense_off[0] = res[0], float(np.float32(1 / (res[0] - 1))), 0
    g.res[1], g.size[1], g.dense_off[1] = res[l], float(np.float32(1 / (res[l] - 1))), 0
    rng = np.random.default_rng(1)
    ijk = rng.integers(0, res[l] - 1, (4000, 3))
    size = np.float32(1 / (res[l] - 1))
    x = torch.from_numpy((ijk.astype(np.float32) + np.float32(0.25)) * size).contiguous()   # interior of the cell
    out = torch.empty(x.shape[0], 5)
    emul.emul_embed(C.byref(g), fp(x), C.c_longlong(x.shape[0]), fp(out), C.c_int(5))
    # oracle with the same synthetic state dict
    sd = {"g.bounds": bounds, "g.entries_size": torch.tensor([g.size[0], g.size[1]]),
          "g.entries_num": torch.tensor([res[0], res[l]]), "g.entries_sum": torch.tensor([res[0] ** 3, 0]),
          "g.dense": dense, "g.hash": table[None], "g.offsets": torch.tensor(
              [[float((c >> 2) & 1), float((c >> 1) & 1), float(c & 1)] for c in range(8)])}
    ref = O.grid_embed(sd, "g.", x, True)
    assert torch.allclose(out, ref, rtol=2e-6, atol=1e-2), (out - ref).abs().max()
    # and the defining property: a uint32-wrapping hash would land elsewhere for most of these
    i = torch.from_numpy(ijk)
    h64 = (i[:, 0] * 1 ^ i[:, 1] * 19349663 ^ i[:, 2] * 83492791) % T
    h32 = ((i[:, 0] * 1 ^ i[:, 1] * 19349663 ^ i[:, 2] * 83492791) & 0xFFFFFFFF) % T
    assert (h64 != h32).float().mean() > 0.5


def test_volume_sampling(emul, golden_setup):
    frame = golden_setup["frame"]
    g = torch.Generator().manual_seed(4)
    pb = frame["pbounds"][0]
    pts = (pb[0] + (pb[1] - pb[0]) * (torch.rand(2000, 3, generator=g) * 1.4 - 0.2)).contiguous()
    pts[:2] = torch.stack([pb[0], pb[1]])
    vol = frame["pbw"][0].contiguous()
    D, H, W, Cc = vol.shape
    out = torch.empty(2000, 1)
    emul.emul_sample_volume(fp(vol), D, H, W, Cc, fp(pb.contiguous()), Cc - 1, 1, fp(pts), C.c_longlong(2000), fp(out))
    ref = O.sample_volume(vol[..., -1:], pts, pb)
    assert torch.allclose(out, ref, atol=1e-6, rtol=1e-5), (out - ref).abs().max()
    tb = frame["tbounds"][0]
    tuv = frame["tuv"][0].contiguous()
    pts = (tb[0] + (tb[1] - tb[0]) * (torch.rand(1000, 3, generator=g) * 1.2 - 0.1)).contiguous()
    out = torch.empty(1000, 2)
    emul.emul_sample_volume(fp(tuv), *tuv.shape, fp(tb.contiguous()), 0, 2, fp(pts), C.c_longlong(1000), fp(out))
    assert torch.allclose(out, O.sample_volume(tuv, pts, tb), atol=1e-6, rtol=1e-5)


def test_ray_points(emul, golden_setup):
    frame, rays = golden_setup["frame"], golden_setup["rays"]
    S = 32
    o, d, n, f = (rays[k][0].contiguous() for k in ("ray_o", "ray_d", "near", "far"))
    R = o.shape[0]
    w, p = torch.empty(R * S, 3), torch.empty(R * S, 3)
    Rm, Th = frame["R"][0].contiguous(), frame["Th"][0].reshape(3).contiguous()
    emul.emul_ray_points(fp(o), fp(d), fp(n), fp(f), C.c_longlong(R), S, fp(Rm), fp(Th), fp(w), fp(p))
    pts, _ = O.sample_along_rays(o, d, n, f, S)
    assert torch.equal(w, pts.reshape(-1, 3))                      # bit-exact: same fp32 statements
    ref_p = torch.matmul(pts.reshape(-1, 3) - frame["Th"][0], Rm)
    assert torch.allclose(p, ref_p, atol=5e-7, rtol=1e-6)


def test_knn_lbs(emul, golden_setup):
    frame = golden_setup["frame"]
    g = torch.Generator().manual_seed(5)
    n = 300
    q = frame["ppts"][0][torch.randperm(6890, generator=g)[:n - 50]] + 0.04 * torch.randn(n - 50, 3, generator=g)
    pb = frame["pbounds"][0]
    q = torch.cat([q, pb[0] + (pb[1] - pb[0]) * torch.rand(50, 3, generator=g)]).contiguous()
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).contiguous()
    pp, pw = frame["part_pts"][0].contiguous(), frame["part_pbw"][0].contiguous()
    ln = frame["lengths2"][0].contiguous()
    A, bA = frame["A"][0].contiguous(), frame["big_A"][0].contiguous()
    bw, pd = torch.empty(n, 5, 24), torch.empty(n, 5)
    x0, v = torch.empty(n, 5, 3), torch.empty(n, 5, 3)
    idx_bf = torch.empty(n, 5, 4, dtype=torch.int32)
    emul.emul_knn_lbs(fp(pp), fp(pw), fp(ln), pp.shape[1], fp(A), fp(bA), fp(q), fp(dirs), C.c_longlong(n),
                      fp(bw), fp(pd), fp(x0), fp(v), C.c_int(0), fp(idx_bf), None)
    # the production search (Morton clusters + AABB pruning) selects exactly the brute-force neighbours
    bw_c, pd_c, x0_c, v_c = (torch.empty_like(t) for t in (bw, pd, x0, v))
    idx_cl = torch.empty_like(idx_bf)
    scanned = C.c_longlong(0)
    emul.emul_knn_lbs(fp(pp), fp(pw), fp(ln), pp.shape[1], fp(A), fp(bA), fp(q), fp(dirs), C.c_longlong(n),
                      fp(bw_c), fp(pd_c), fp(x0_c), fp(v_c), C.c_int(1), fp(idx_cl), C.byref(scanned))
    assert torch.equal(idx_bf, idx_cl)
    assert torch.equal(bw, bw_c) and torch.equal(pd, pd_c) and torch.equal(x0, x0_c)
    total_clusters = int(sum((int(c) + 15) // 16 for c in ln)) * n
    print(f"[knn] clusters scanned {scanned.value} of {total_clusters} ({scanned.value / total_clusters:.3f})")
    assert scanned.value < 0.25 * total_clusters
    rbw, rpd = O.knn_blend_weights(q, pp, pw, ln)
    assert torch.allclose(bw, rbw, atol=1e-6, rtol=1e-5), (bw - rbw).abs().max()
    assert torch.allclose(pd, rpd, atol=1e-6, rtol=1e-5)
    qe = q[:, None].expand(n, 5, 3).reshape(-1, 3)
    de = dirs[:, None].expand(n, 5, 3).reshape(-1, 3)
    rx0, rv = O.lbs_to_bigpose(qe, de, rbw.reshape(-1, 24), A, bA)
    assert torch.allclose(x0.reshape(-1, 3), rx0, atol=1e-5, rtol=1e-4), (x0.reshape(-1, 3) - rx0).abs().max()
    assert torch.allclose(v.reshape(-1, 3), rv, atol=1e-5, rtol=1e-4)


def test_deformer(emul, golden_setup):
    net, sd, frame = golden_setup["net"], golden_setup["sd"], golden_setup["frame"]
    g = torch.Generator().manual_seed(6)
    tb = frame["tbounds"][0]
    x0 = (tb[0] + (tb[1] - tb[0]) * (torch.rand(400, 3, generator=g) * 1.2 - 0.1)).contiguous()
    gd = cabi.grid_desc(net.tpose_deformer.embedder)
    mlp = (cabi.NvrLinear * 3)(*[cabi.linear_desc(net.tpose_deformer.mlp[i]) for i in (0, 2, 4)])
    tuv = frame["tuv"][0].contiguous()
    out = torch.empty(400, 3)
    emul.emul_deformer(C.byref(gd), mlp, fp(tuv), *tuv.shape[:3], fp(tb.contiguous()), C.c_float(float(frame["frame_dim"][0])),
                       fp(x0), C.c_longlong(400), fp(out))
    ref = O.deformer(sd, x0, tuv, tb, frame["frame_dim"])
    assert torch.allclose(out, ref, atol=1e-6, rtol=1e-4), (out - ref).abs().max()


def test_posenc_and_activations(emul):
    g = torch.Generator().manual_seed(7)
    v = torch.nn.functional.normalize(torch.randn(256, 3, generator=g), dim=-1).contiguous()
    out = torch.empty(256, 27)
    emul.emul_posenc(fp(v), C.c_longlong(256), fp(out))
    assert torch.allclose(out, O.posenc(v), atol=1e-6)
    x = torch.cat([torch.linspace(-30, 30, 999), torch.tensor([19.999, 20.0, 20.001])]).contiguous()
    sp, sg = torch.empty_like(x), torch.empty_like(x)
    emul.emul_activations(fp(x), C.c_longlong(x.numel()), fp(sp), fp(sg))
    assert torch.allclose(sp, torch.nn.functional.softplus(x), rtol=2e-6, atol=1e-30)
    assert torch.allclose(sg, torch.sigmoid(x), rtol=2e-6, atol=1e-30)


def test_cull_early_out_is_conservative(emul):
    """The coarse-minimum early-out of k_cull may only fire on samples the exact 8-tap lookup culls too; on the synthetic
    frame it should remove most of the work (samples far from the body)."""
    from instant_nvr_b200.synthetic import make_frame
    frame = make_frame(seed=3)
    dist = frame["pbw"][0, ..., -1].contiguous()
    D, H, W = dist.shape
    b = frame["pbounds"][0].contiguous()
    g = torch.Generator().manual_seed(0)
    lo, hi = b[0], b[1]
    pts = (lo + (hi - lo) * (torch.rand(400000, 3, generator=g) * 1.3 - 0.15)).contiguous()     # inside and outside the bbox
    special = torch.tensor([[float("nan"), 0.0, 0.0], [float("inf"), 0.0, 0.0], [-float("inf"), 1.0, 1.0]])
    pts = torch.cat([pts, special, lo[None], hi[None]]).contiguous()
    n = pts.shape[0]
    for thresh in (0.05, 0.1):
        keep, early = torch.zeros(n, dtype=torch.uint8), torch.zeros(n, dtype=torch.uint8)
        emul.emul_cull(fp(dist), C.c_int(D), C.c_int(H), C.c_int(W), fp(b), fp(pts), C.c_longlong(n), C.c_float(thresh),
                       fp(keep), fp(early))
        assert not (early.bool() & keep.bool()).any()
        culled = ~keep.bool()
        assert keep.sum() > 1000
        assert early.sum() > 0.6 * culled.sum(), (int(early.sum()), int(culled.sum()))
    # adversarial volume: values straddling the threshold by a few ulps everywhere
    vol = (0.05 * (1.0 + 3e-6 * torch.randn(D, H, W, generator=g))).contiguous()
    keep, early = torch.zeros(n, dtype=torch.uint8), torch.zeros(n, dtype=torch.uint8)
    emul.emul_cull(fp(vol), C.c_int(D), C.c_int(H), C.c_int(W), fp(b), fp(pts), C.c_longlong(n), C.c_float(0.05), fp(keep), fp(early))
    assert not (early.bool() & keep.bool()).any() and keep.any()


def test_cull_quick_world_space_is_conservative(emul):
    """nvr_cull_quick (one affine map from WORLD coordinates to voxel coordinates, coarse-minimum grid with a one-voxel
    margin) may only cull samples the exact world -> pose -> 8-tap lookup culls too, and should remove most of them."""
    from instant_nvr_b200.synthetic import make_frame
    g = torch.Generator().manual_seed(1)
    for seed in (3, 5):
        frame = make_frame(seed=seed)
        dist = frame["pbw"][0, ..., -1].contiguous()
        D, H, W = dist.shape
        b, R, Th = frame["pbounds"][0].contiguous(), frame["R"][0].contiguous(), frame["Th"][0].contiguous()
        lo, hi = b[0], b[1]
        pose = lo + (hi - lo) * (torch.rand(400000, 3, generator=g) * 1.3 - 0.15)
        # points exactly on voxel planes and coarse-cell boundaries: where an approximate coordinate may change cell
        grid = lo + (hi - lo) * (torch.randint(0, max(D, H, W), (50000, 3), generator=g).float()
                                 / torch.tensor([D - 1.0, H - 1.0, W - 1.0])).clamp(max=1.0)
        pose = torch.cat([pose, grid, lo[None], hi[None]])
        wpts = (pose @ R.T + Th).contiguous()                         # world points whose pose-space image is `pose`
        special = torch.tensor([[float("nan"), 0.0, 0.0], [float("inf"), 0.0, 0.0], [-float("inf"), 1.0, 1.0], [1e30, -1e30, 0.0]])
        wpts = torch.cat([wpts, special]).contiguous()
        n = wpts.shape[0]
        for vol, thresh in ((dist, 0.05), (dist, 0.1), ((0.05 * (1.0 + 3e-6 * torch.randn(D, H, W, generator=g))).contiguous(), 0.05)):
            keep, quick = torch.zeros(n, dtype=torch.uint8), torch.zeros(n, dtype=torch.uint8)
            emul.emul_cull_quick(fp(vol), C.c_int(D), C.c_int(H), C.c_int(W), fp(b), fp(R), fp(Th), fp(wpts), C.c_longlong(n),
                                 C.c_float(thresh), fp(keep), fp(quick))
            assert not (quick.bool() & keep.bool()).any()
            if vol is dist:
                assert keep.sum() > 1000 and quick.sum() > 0.55 * (~keep.bool()).sum(), (int(quick.sum()), int((~keep.bool()).sum()))


@pytest.mark.parametrize("n_rays,S", [(1, 1), (1, 7), (31, 16), (32, 64), (33, 64), (100, 128), (1000, 3), (4097, 32), (20000, 128),
                                      (65, 256), (7, 1000)])
def test_cull_walk_visits_every_sample_once(emul, n_rays, S):
    """k_cull's depth-major walk (groups of 32 rays in chunks of 16 rays x 2 steps, 2048-position spans) is a bijection onto the
    n_rays x S samples for ragged sizes: ray counts that are not multiples of 32, sample counts that do not divide a span."""
    visits = np.zeros(n_rays * S, dtype=np.int32)
    npos = C.c_longlong(0)
    for grid in (1, 3, 148 * 8):
        visits[:] = 0
        emul.emul_cull_walk(C.c_longlong(n_rays), C.c_int(S), C.c_int(grid), visits.ctypes.data_as(C.c_void_p), C.byref(npos))
        assert visits.min() == 1 and visits.max() == 1, (grid, int(visits.min()), int(visits.max()))
        group = 64 * ((((S + 1) // 2) + 3) // 4 * 4)           # positions per 32 rays: depth pairs padded to a multiple of 4
        assert npos.value >= n_rays * S and npos.value < ((n_rays + 31) // 32) * group + 2048


@pytest.mark.parametrize("n_rays,S", [(32, 64), (100, 128), (4097, 32), (65, 256), (33, 7)])
def test_cull_walk_chunks_are_compact(emul, n_rays, S):
    """32 consecutive positions of the walk -- one warp iteration of k_cull, hence the granule of the survivor order and of the
    KNN units -- cover at most 16 consecutive rays and 2 consecutive depth steps (csrc/nvr_math.cuh cull_locate)."""
    mr, ms = C.c_int(0), C.c_int(0)
    emul.emul_cull_chunk_extent(C.c_longlong(n_rays), C.c_int(S), C.byref(mr), C.byref(ms))
    assert 1 <= mr.value <= 16 and 1 <= ms.value <= 2, (mr.value, ms.value)

// TEST-ONLY host emulation of the per-sample device arithmetic (instant_nvr_b200/csrc/nvr_math.cuh).
//
// Compiled with g++ by tests/test_host_emul.py and driven through ctypes so that the exact
// statements the CUDA kernels execute per sample can be checked against the oracle in a container
// without a GPU.  Nothing in the product links or loads this file; it is not a CPU fallback.
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/nvr_b200.h"
#include "../../instant_nvr_b200/csrc/nvr_math.cuh"
#include "../../instant_nvr_b200/csrc/nvr_smpl.cuh"

static GridDev to_dev(const NvrGrid& g) {
    GridDev d;
    memset(&d, 0, sizeof(d));
    d.dense = g.dense; d.hash = g.hash; d.bounds = g.bounds;
    d.n_levels = g.n_levels; d.n_feat = g.n_feat; d.start_hash = g.start_hash; d.sum_features = g.sum_features;
    d.T = (unsigned long long)g.table_size;
    d.T_magic = (unsigned long long)((((unsigned __int128)1) << 64) / (unsigned __int128)g.table_size);
    for (int l = 0; l < NVR_MAX_LEVELS; ++l) { d.res[l] = g.res[l]; d.size[l] = g.size[l]; d.dense_off[l] = g.dense_off[l]; }
    return d;
}

extern "C" {

void emul_embed(const NvrGrid* g, const float* xyz, long long n, float* out, int out_dim) {
    GridDev d = to_dev(*g);
    for (long long i = 0; i < n; ++i) {
        if (d.n_feat == 16) nvr_embed_point<16>(d, xyz + i * 3, out + i * out_dim);
        else nvr_embed_point<2>(d, xyz + i * 3, out + i * out_dim);
    }
}

// Barrett reduction against the plain modulo, over hash values the path can produce.
long long emul_check_mod(long long T, const long long* h, long long n) {
    unsigned long long magic = (unsigned long long)((((unsigned __int128)1) << 64) / (unsigned __int128)T);
    long long bad = 0;
    for (long long i = 0; i < n; ++i)
        if (nvr_mod_T((unsigned long long)h[i], (unsigned long long)T, magic) != (unsigned long long)h[i] % (unsigned long long)T) ++bad;
    // the 32-bit reduction used by the gather kernel, on the values inside its domain (h < 2^40)
    const unsigned int magic40 = (unsigned int)((1ull << 40) / (unsigned long long)T);
    for (long long i = 0; i < n; ++i)
        if ((unsigned long long)h[i] < (1ull << 40) &&
            nvr_mod_T40((unsigned long long)h[i], (unsigned int)T, magic40) != (unsigned long long)h[i] % (unsigned long long)T) ++bad;
    return bad;
}

void emul_sample_volume(const float* vol, int D, int H, int W, int Cc, const float* bounds, int ch0, int nch,
                        const float* pts, long long n, float* out) {
    VolumeDev v{vol, D, H, W, Cc, bounds};
    for (long long i = 0; i < n; ++i) nvr_sample_volume(v, pts + i * 3, ch0, nch, out + i * nch);
}

void emul_ray_points(const float* ray_o, const float* ray_d, const float* near_, const float* far_, long long n_rays, int S,
                     const float* R, const float* Th, float* wpts, float* ppts) {
    for (long long r = 0; r < n_rays; ++r)
        for (int k = 0; k < S; ++k) {
            float w[3], p[3];
            nvr_ray_sample(ray_o + r * 3, ray_d + r * 3, near_[r], far_[r], k, S, w);
            nvr_world_to_pose(R, Th, w, p);
            for (int a = 0; a < 3; ++a) { wpts[(r * S + k) * 3 + a] = w[a]; ppts[(r * S + k) * 3 + a] = p[a]; }
        }
}

// Host restatement of k_cluster_verts: balanced KD partition into clusters of NVR_CL (median split along
// the longest axis, left half rounded to whole clusters; ties by vertex index), one AABB per cluster.
struct HostClusters { std::vector<float4> verts, lo, hi; };
static float idx_bits(int j) { float f; memcpy(&f, &j, 4); return f; }
static void kd_split(const float* src, std::vector<int>& ids, int a, int b) {     // [a, b) in clusters
    if (b - a <= 1) return;
    const int i0 = a * NVR_CL, i1 = std::min<int>((int)ids.size(), b * NVR_CL);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = i0; i < i1; ++i)
        for (int x = 0; x < 3; ++x) { lo[x] = fminf(lo[x], src[ids[i] * 3 + x]); hi[x] = fmaxf(hi[x], src[ids[i] * 3 + x]); }
    int ax = 0;
    float best = -1.0f;
    for (int x = 0; x < 3; ++x) if (hi[x] - lo[x] > best) { best = hi[x] - lo[x]; ax = x; }
    std::sort(ids.begin() + i0, ids.begin() + i1, [&](int p, int q) {
        const float fp = src[p * 3 + ax], fq = src[q * 3 + ax];
        return fp < fq || (fp == fq && p < q);
    });
    const int mid = a + (b - a) / 2;
    kd_split(src, ids, a, mid);
    kd_split(src, ids, mid, b);
}
static HostClusters build_clusters(const float* src, int n) {
    HostClusters hc;
    const int ncl = (n + NVR_CL - 1) / NVR_CL;
    std::vector<int> ids(n);
    for (int j = 0; j < n; ++j) ids[j] = j;
    kd_split(src, ids, 0, ncl);
    // one structure-of-arrays block per cluster: x[16] | y[16] | z[16] | index bits[16]  (k_cluster_apply)
    hc.verts.assign((size_t)ncl * NVR_CL, float4{INFINITY, INFINITY, INFINITY, INFINITY});
    for (int c = 0; c < ncl; ++c) for (int s = 0; s < NVR_CL; ++s) ((float*)&hc.verts[(size_t)c * NVR_CL])[48 + s] = idx_bits(0);
    hc.lo.assign(ncl, float4{INFINITY, INFINITY, INFINITY, 0.f});
    hc.hi.assign(ncl, float4{-INFINITY, -INFINITY, -INFINITY, 0.f});
    for (int i = 0; i < n; ++i) {
        const int j = ids[i], c = i / NVR_CL;
        float* blk = (float*)&hc.verts[(size_t)c * NVR_CL];
        const int slot = i % NVR_CL;
        blk[slot] = src[j * 3]; blk[16 + slot] = src[j * 3 + 1]; blk[32 + slot] = src[j * 3 + 2]; blk[48 + slot] = idx_bits(j);
        hc.lo[c].x = fminf(hc.lo[c].x, src[j * 3]); hc.hi[c].x = fmaxf(hc.hi[c].x, src[j * 3]);
        hc.lo[c].y = fminf(hc.lo[c].y, src[j * 3 + 1]); hc.hi[c].y = fmaxf(hc.hi[c].y, src[j * 3 + 1]);
        hc.lo[c].z = fminf(hc.lo[c].z, src[j * 3 + 2]); hc.hi[c].z = fmaxf(hc.hi[c].z, src[j * 3 + 2]);
    }
    return hc;
}

// per point: for each of the 5 parts, K=4 NN blend weights (bw 24, pdist) and the LBS warp (x0, v).
// clustered = 0: brute force over the vertices in their original order.
// clustered = 1: the pruned search of k_warp (seed cluster, then AABB lower bounds), scalar form: the
//                seed is chosen for a DIFFERENT point (the previous one), as lane 0's query is on the GPU.
// idx_out (optional): the 4 selected vertex indices per (point, part), sorted ascending by (d2, index).
void emul_knn_lbs(const float* part_pts, const float* part_pbw, const long long* lengths2, int maxlen, const float* A,
                  const float* bigA, const float* pts, const float* dirs, long long n, float* bw_out, float* pdist_out,
                  float* x0_out, float* v_out, int clustered, int* idx_out, long long* scanned_out) {
    long long scanned = 0;
    for (int part = 0; part < NVR_PARTS; ++part) {
        const int cnt = (int)lengths2[part];
        const float* src = part_pts + (long long)part * maxlen * 3;
        // brute force: the vertices in their original order, cut into blocks of 16 (+inf padding)
        const int nblk = (cnt + NVR_CL - 1) / NVR_CL;
        std::vector<float4> verts((size_t)nblk * NVR_CL, float4{INFINITY, INFINITY, INFINITY, INFINITY});
        for (int j = 0; j < nblk * NVR_CL; ++j) {
            float* blk = (float*)&verts[(size_t)(j / NVR_CL) * NVR_CL];
            const int slot = j % NVR_CL;
            if (j < cnt) { blk[slot] = src[j * 3]; blk[16 + slot] = src[j * 3 + 1]; blk[32 + slot] = src[j * 3 + 2]; }
            blk[48 + slot] = idx_bits(j < cnt ? j : 0);
        }
        HostClusters hc = build_clusters(src, cnt);
        const int ncl = (int)hc.lo.size();
        for (long long i = 0; i < n; ++i) {
            Knn4 k;
            nvr_knn_init(k);
            const float* p = pts + i * 3;
            if (!clustered) {
                for (int b = 0; b < nblk; ++b) nvr_knn_scan(verts.data() + (size_t)b * NVR_CL, p, k);
            } else if (ncl > 0) {
                const float* rep = clustered == 2 ? p : pts + (i > 0 ? i - 1 : 0) * 3;   // 2: own point (best case)
                int seed = 0;
                float best = INFINITY;
                for (int c = 0; c < ncl; ++c) {
                    const float lb = nvr_aabb_lb(hc.lo[c], hc.hi[c], rep);
                    if (lb < best) { best = lb; seed = c; }
                }
                nvr_knn_scan(hc.verts.data() + (size_t)seed * NVR_CL, p, k);
                ++scanned;
                for (int c = 0; c < ncl; ++c) {
                    if (c == seed) continue;
                    const float lb = nvr_aabb_lb(hc.lo[c], hc.hi[c], p);
                    if (!(lb * NVR_PRUNE_SLACK > nvr_knn_d2(k, 3))) { nvr_knn_scan(hc.verts.data() + (size_t)c * NVR_CL, p, k); ++scanned; }
                }
            }
            float bw[NVR_JOINTS];
            float pd = nvr_knn_blend(k, part_pbw + (long long)part * maxlen * NVR_JOINTS, bw);
            float x0[3], v[3];
            nvr_lbs_to_bigpose(bw, A, bigA, pts + i * 3, dirs + i * 3, x0, v);
            const long long o = i * NVR_PARTS + part;
            for (int j = 0; j < NVR_JOINTS; ++j) bw_out[o * NVR_JOINTS + j] = bw[j];
            pdist_out[o] = pd;
            for (int a = 0; a < 3; ++a) { x0_out[o * 3 + a] = x0[a]; v_out[o * 3 + a] = v[a]; }
            if (idx_out) for (int q = 0; q < NVR_KNN; ++q) idx_out[o * NVR_KNN + q] = nvr_knn_idx(k, q);
        }
    }
    if (scanned_out) *scanned_out = scanned;
}

void emul_deformer(const NvrGrid* g, const NvrLinear* mlp, const float* tuv, int D, int H, int W, const float* tbounds,
                   float frame_dim, const float* x0, long long n, float* out) {
    GridDev d = to_dev(*g);
    DeformerMlp m{mlp[0].weight, mlp[0].bias, mlp[1].weight, mlp[1].bias, mlp[2].weight, mlp[2].bias};
    VolumeDev v{tuv, D, H, W, 2, tbounds};
    float sc[32];
    alignas(16) float pk[NVR_DEF_PACKED_FLOATS];                  // the layout stage_deformer builds in shared memory
    nvr_pack_deformer(m, pk, 0, 1);
    for (long long i = 0; i < n; ++i) nvr_deformer_point(d, pk, v, frame_dim, x0 + i * 3, out + i * 3, sc, 1);
}

void emul_posenc(const float* v, long long n, float* out) {
    for (long long i = 0; i < n; ++i) nvr_posenc27(v + i * 3, out + i * 27);
}

void emul_activations(const float* x, long long n, float* softplus, float* sigmoid) {
    for (long long i = 0; i < n; ++i) { softplus[i] = nvr_softplus(x[i]); sigmoid[i] = nvr_sigmoid(x[i]); }
}

int emul_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(NvrGrid);
        case 1: return (int)sizeof(NvrLinear);
        case 2: return (int)sizeof(NvrPart);
        case 3: return (int)sizeof(NvrParams);
        case 4: return (int)sizeof(NvrFrame);
        case 5: return (int)sizeof(NvrConfig);
        case 6: return (int)sizeof(NvrCounters);
        case 7: return (int)sizeof(NvrStageProfile);
        case 8: return (int)sizeof(NvrAdamTensor);
        case 9: return (int)sizeof(NvrSmplPose);
        case 10: return (int)sizeof(NvrSmplOut);
        case 11: return (int)sizeof(SmplPoseDev);
    }
    return -1;
}
// per-frame SMPL preprocessing arithmetic (csrc/nvr_smpl.cuh)
void emul_smpl_chain(const double* poses, const float* joints, const int* parents, float* out) {
    double G[NVR_JOINTS * 16];
    nvr_smpl_chain(poses, joints, parents, G, out);
}
void emul_rodrigues_cv(const double* r, double* R) { nvr_rodrigues_cv(r, R); }
int emul_arange_len(double start, double stop, double step) { return nvr_arange_len(start, stop, step); }
void emul_arange_fill(double start, double step, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = nvr_arange_val(start, step, i);
}
// get_rays + get_near_far per pixel, row-major, no compaction (the kernels k_rays_mask / k_rays_emit run exactly
// these two calls per pixel); o = float32 camera origin.
void emul_rays(int H, int W, const double* Kinv, const double* R, const double* T, const float* bounds, float* o,
               float* ray_d, float* near_, float* far_, unsigned char* mask) {
    CameraDev cam;
    for (int i = 0; i < 9; ++i) { cam.Kinv[i] = Kinv[i]; cam.R[i] = R[i]; }
    for (int a = 0; a < 3; ++a) {
        cam.T[a] = T[a];
        cam.o[a] = -((R[0 * 3 + a] * T[0] + R[1 * 3 + a] * T[1]) + R[2 * 3 + a] * T[2]);
        o[a] = (float)cam.o[a];
    }
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            const long long pix = (long long)j * W + i;
            nvr_pixel_ray(cam, i, j, ray_d + pix * 3);
            mask[pix] = nvr_near_far(bounds, o, ray_d + pix * 3, near_ + pix, far_ + pix) ? 1 : 0;
        }
}

// one Adam step over n elements with the scalars nvr_adam_step derives on the host
void emul_adam(float* p, const float* g, float* m, float* v, long long n, long long step, double lr, double wd, double beta1,
               double beta2, double eps) {
    AdamScalars s;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    s.beta1 = (float)beta1; s.beta2 = (float)beta2;
    s.one_minus_beta1 = (float)(1.0 - beta1); s.one_minus_beta2 = (float)(1.0 - beta2);
    s.eps = (float)eps; s.weight_decay = (float)wd;
    s.neg_step_size = (float)(-(lr / bc1));
    s.bc2_sqrt = (float)sqrt(bc2);
    for (long long i = 0; i < n; ++i) nvr_adam_update(s, p[i], g[i], m[i], v[i]);
}
// distance cull with and without the coarse-minimum early-out (k_frame_coarse + k_cull): keep[i] = exact decision,
// early[i] = 1 when the early-out fires (it must then agree with a culled exact decision).
// k_cull's walk, position by position exactly as the kernel's CTAs / warps / lanes take them: visits[sample id] += 1
void emul_cull_walk(long long n_rays, int S, int grid, int* visits, long long* n_positions) {
    const int SPAN = 2048, T = 8;
    CullWalk cw;
    cw.S = S; cw.group = cull_group_positions(S); cw.n_rays = n_rays;
    const long long n_map = ((n_rays + 31) / 32) * (long long)cw.group;
    long long pos = 0;
    for (int block = 0; block < grid; ++block)
        for (long long sbase = (long long)block * SPAN; sbase < n_map; sbase += (long long)grid * SPAN) {
            cw.g0 = sbase / (long long)cw.group;
            cw.w0 = (unsigned)(sbase - cw.g0 * (long long)cw.group);
            for (int wid = 0; wid < 8; ++wid)
                for (int t = 0; t < T; ++t)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int local = (wid * T + t) * 32 + lane;
                        long long r, i; int k;
                        const bool valid = cull_locate(cw, local, r, k, i) && sbase + local < n_map;
                        ++pos;
                        if (valid) { if (k < 0 || k >= S || i != r * S + k) visits[0] = -1000000; else visits[i] += 1; }
                    }
        }
    *n_positions = pos;
}
// extent of every aligned 32-position chunk of the walk (= one warp iteration of k_cull, the granule of the survivor order):
// the largest number of distinct rays and of distinct depth steps a chunk covers, over the whole walk
void emul_cull_chunk_extent(long long n_rays, int S, int* max_rays, int* max_steps) {
    CullWalk cw;
    cw.S = S; cw.group = cull_group_positions(S); cw.n_rays = n_rays;
    const long long n_map = ((n_rays + 31) / 32) * (long long)cw.group;
    *max_rays = 0; *max_steps = 0;
    for (long long sbase = 0; sbase < n_map; sbase += 32) {
        cw.g0 = sbase / (long long)cw.group;
        cw.w0 = (unsigned)(sbase - cw.g0 * (long long)cw.group);
        long long rlo = 1ll << 60, rhi = -1; int klo = 1 << 30, khi = -1;
        for (int lane = 0; lane < 32; ++lane) {
            long long r, i; int k;
            if (!cull_locate(cw, lane, r, k, i)) continue;
            rlo = r < rlo ? r : rlo; rhi = r > rhi ? r : rhi; klo = k < klo ? k : klo; khi = k > khi ? k : khi;
        }
        if (rhi >= 0) {
            if ((int)(rhi - rlo + 1) > *max_rays) *max_rays = (int)(rhi - rlo + 1);
            if (khi - klo + 1 > *max_steps) *max_steps = khi - klo + 1;
        }
    }
}
// quick world-space cull (nvr_cull_quick) next to the exact decision for world points: keep = exact lookup < thresh
void emul_cull_quick(const float* dist, int D, int H, int W, const float* bounds, const float* R, const float* Th, const float* wpts,
                     long long n, float thresh, unsigned char* keep, unsigned char* quick) {
    VolumeDev v{dist, D, H, W, 1, bounds};
    const int cD = nvr_coarse_dim(D), cH = nvr_coarse_dim(H), cW = nvr_coarse_dim(W);
    std::vector<float> cmin((size_t)cD * cH * cW);
    for (int i = 0; i < cD * cH * cW; ++i) cmin[i] = nvr_coarse_min(dist, D, H, W, i / (cH * cW), (i / cW) % cH, i % cW);
    CullQuick q;
    nvr_cull_quick_setup(v, R, Th, q);
    for (long long i = 0; i < n; ++i) {
        const float* w = wpts + i * 3;
        float cq[3], p[3], c[3], pn;
        for (int a = 0; a < 3; ++a) cq[a] = ((w[0] * q.M[a] + w[1] * q.M[3 + a]) + w[2] * q.M[6 + a]) + q.t[a];
        quick[i] = nvr_cull_quick(v, q, cmin.data(), cq, thresh) ? 1 : 0;
        nvr_world_to_pose(R, Th, w, p);
        nvr_volume_coords(v, p, c);
        nvr_sample_volume_at(v, c, 0, 1, &pn);
        keep[i] = pn < thresh ? 1 : 0;
    }
}
void emul_cull(const float* dist, int D, int H, int W, const float* bounds, const float* pts, long long n, float thresh,
               unsigned char* keep, unsigned char* early) {
    VolumeDev v{dist, D, H, W, 1, bounds};
    const int cD = nvr_coarse_dim(D), cH = nvr_coarse_dim(H), cW = nvr_coarse_dim(W);
    std::vector<float> cmin((size_t)cD * cH * cW);
    for (int i = 0; i < cD * cH * cW; ++i) cmin[i] = nvr_coarse_min(dist, D, H, W, i / (cH * cW), (i / cW) % cH, i % cW);
    for (long long i = 0; i < n; ++i) {
        float c[3], pn;
        nvr_volume_coords(v, pts + i * 3, c);
        early[i] = nvr_cull_early_out(v, cmin.data(), c, thresh) ? 1 : 0;
        nvr_sample_volume_at(v, c, 0, 1, &pn);
        keep[i] = pn < thresh ? 1 : 0;
    }
}
}

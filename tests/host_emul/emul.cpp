// TEST-ONLY host emulation of the per-sample device arithmetic (instant_nvr_b200/csrc/nvr_math.cuh).
//
// Compiled with g++ by tests/test_host_emul.py and driven through ctypes so that the exact
// statements the CUDA kernels execute per sample can be checked against the oracle in a container
// without a GPU.  Nothing in the product links or loads this file; it is not a CPU fallback.
#include <string.h>

#include <vector>

#include "../../include/nvr_b200.h"
#include "../../instant_nvr_b200/csrc/nvr_math.cuh"

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

// per point: for each of the 5 parts, K=4 NN blend weights (bw 24, pdist) and the LBS warp (x0, v).
void emul_knn_lbs(const float* part_pts, const float* part_pbw, const long long* lengths2, int maxlen, const float* A,
                  const float* bigA, const float* pts, const float* dirs, long long n, float* bw_out, float* pdist_out,
                  float* x0_out, float* v_out) {
    for (int part = 0; part < NVR_PARTS; ++part) {
        std::vector<float4> verts(lengths2[part]);
        for (long long j = 0; j < lengths2[part]; ++j) {
            const float* s = part_pts + ((long long)part * maxlen + j) * 3;
            verts[j] = float4{s[0], s[1], s[2], 0.f};
        }
        for (long long i = 0; i < n; ++i) {
            Knn4 k;
            nvr_knn_init(k);
            nvr_knn_scan(verts.data(), (int)verts.size(), pts + i * 3, k);
            float bw[NVR_JOINTS];
            float pd = nvr_knn_blend(k, part_pbw + (long long)part * maxlen * NVR_JOINTS, bw);
            float x0[3], v[3];
            nvr_lbs_to_bigpose(bw, A, bigA, pts + i * 3, dirs + i * 3, x0, v);
            const long long o = i * NVR_PARTS + part;
            for (int j = 0; j < NVR_JOINTS; ++j) bw_out[o * NVR_JOINTS + j] = bw[j];
            pdist_out[o] = pd;
            for (int a = 0; a < 3; ++a) { x0_out[o * 3 + a] = x0[a]; v_out[o * 3 + a] = v[a]; }
        }
    }
}

void emul_deformer(const NvrGrid* g, const NvrLinear* mlp, const float* tuv, int D, int H, int W, const float* tbounds,
                   float frame_dim, const float* x0, long long n, float* out) {
    GridDev d = to_dev(*g);
    DeformerMlp m{mlp[0].weight, mlp[0].bias, mlp[1].weight, mlp[1].bias, mlp[2].weight, mlp[2].bias};
    VolumeDev v{tuv, D, H, W, 2, tbounds};
    for (long long i = 0; i < n; ++i) nvr_deformer_point(d, m, v, frame_dim, x0 + i * 3, out + i * 3);
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
    }
    return -1;
}
}

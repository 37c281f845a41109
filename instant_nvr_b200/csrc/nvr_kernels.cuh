// nvr_kernels.cuh -- the sm_100a kernels of the per-ray hot path.
//
//   k_frame_prep   per frame: distance channel of pbw -> compact volume, part vertices -> packed float4
//   k_cull         sample gen (ray mode) / point fetch, world->pose, distance cull, block compaction
//   k_warp         per survivor: 5x exact K=4 NN, Gaussian blend weights, LBS to big pose, deformer,
//                  per-part append of flagged (sample, part) pairs
//   k_embed        THE gather: quad-lane 64-byte row loads of the dense+hashed grids, per-level sums
//   k_mlp          occ + rgb MLPs on 128-pair tiles (fp32 FFMA register tiles)
//   k_resolve      arg-max part fusion; per-sample raw/occ and/or per-ray alpha compositing
//
// No host synchronisation anywhere: list lengths live in device counters, every consumer kernel is
// a persistent grid-stride loop that reads its trip count from them.
#pragma once
#include <cuda_runtime.h>

#include "nvr_math.cuh"

#define NVR_EMB_STRIDE 20          // 19 used
#define NVR_CTR_SURV 0
#define NVR_CTR_PAIR 1             // [1..5]
#define NVR_CTR_WORDS 16

struct __align__(16) PairRec {     // one flagged (sample, part) pair: 32 B
    float x, y, z;                 // canonical (big pose + residual) point
    float vx, vy, vz;              // canonical view direction
    int surv;                      // survivor slot
    int _pad;
};

struct FrameDev {                  // per-frame tensors as the kernels see them
    const float* R;
    const float* Th;
    VolumeDev dist;                // compact (D,H,W,1) distance volume
    VolumeDev tuv;                 // (D',H',W',2)
    const float4* verts;           // packed part vertices
    const int* part_off;           // [6] offsets into verts (device)
    const float* part_pbw;         // (P, maxlen, 24)
    int maxlen;
    const float* A;
    const float* bigA;
    const float* frame_dim;
    const long long* latent_index;
};

struct LinearDev { const float* w; const float* b; int in, out; };
struct PartMlpDev {
    LinearDev occ[2];
    LinearDev rgb[3];
    int n_rgb;
    int n_latent;
    const float* latent;
};

// -----------------------------------------------------------------------------------------
// per-frame preparation
// -----------------------------------------------------------------------------------------
__global__ void k_frame_prep(const float* __restrict__ pbw, int n_vox, int C, float* __restrict__ dist,
                             const float* __restrict__ part_pts, const long long* __restrict__ lengths2,
                             int maxlen, float4* __restrict__ verts, int* __restrict__ part_off) {
    int off[NVR_PARTS + 1];
    off[0] = 0;
#pragma unroll
    for (int p = 0; p < NVR_PARTS; ++p) {
        long long len = lengths2[p];
        len = len < 0 ? 0 : (len > maxlen ? maxlen : len);
        off[p + 1] = off[p] + (int)len;
    }
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (tid <= NVR_PARTS) part_off[tid] = off[tid];
    for (int i = tid; i < n_vox; i += nth) dist[i] = pbw[(long long)i * C + (C - 1)];
    for (int i = tid; i < NVR_PARTS * maxlen; i += nth) {
        const int p = i / maxlen, j = i - p * maxlen;
        if (j < off[p + 1] - off[p]) {
            const float* s = part_pts + (long long)i * 3;
            verts[off[p] + j] = make_float4(s[0], s[1], s[2], 0.0f);
        }
    }
}

// -----------------------------------------------------------------------------------------
// cull: samples -> survivors
// -----------------------------------------------------------------------------------------
// ray mode  (n_samples > 0): sample i = ray i / n_samples, step i % n_samples; pts = ray_o, aux = ray_d
// point mode (n_samples == 0): pts = wpts (n,3)
__global__ void __launch_bounds__(256)
k_cull(FrameDev fr, const float* __restrict__ pts, const float* __restrict__ ray_d,
       const float* __restrict__ near_, const float* __restrict__ far_, long long n, int n_samples,
       float thresh, int* __restrict__ counters, int* __restrict__ surv_of_sample, float4* __restrict__ surv) {
    __shared__ int warp_cnt[8];
    __shared__ int block_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long base = (long long)blockIdx.x * 256; base < n; base += (long long)gridDim.x * 256) {
        const long long i = base + threadIdx.x;
        bool keep = false;
        float p[3] = {0.f, 0.f, 0.f};
        if (i < n) {
            float w[3];
            if (n_samples > 0) {
                const long long r = i / n_samples;
                const int k = (int)(i - r * n_samples);
                const float o[3] = {pts[r * 3], pts[r * 3 + 1], pts[r * 3 + 2]};
                const float d[3] = {ray_d[r * 3], ray_d[r * 3 + 1], ray_d[r * 3 + 2]};
                nvr_ray_sample(o, d, near_[r], far_[r], k, n_samples, w);
            } else {
                w[0] = pts[i * 3]; w[1] = pts[i * 3 + 1]; w[2] = pts[i * 3 + 2];
            }
            nvr_world_to_pose(fr.R, fr.Th, w, p);
            float pn;
            nvr_sample_volume(fr.dist, p, 0, 1, &pn);
            keep = pn < thresh;                                   // inb_part_network_multiassign.py:136
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(ballot);
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
            block_base = tot ? atomicAdd(&counters[NVR_CTR_SURV], tot) : 0;
        }
        __syncthreads();
        if (i < n) {
            int slot = -1;
            if (keep) {
                slot = block_base + warp_cnt[wid] + __popc(ballot & ((1u << lane) - 1u));
                surv[slot] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));   // passes are < 2^31 samples
            }
            surv_of_sample[i] = slot;
        }
        __syncthreads();
    }
}

// -----------------------------------------------------------------------------------------
// warp: survivors -> flagged (sample, part) pairs in canonical space
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_warp(FrameDev fr, GridDev dg, DeformerMlp dm_g, const float* __restrict__ dirs, int dir_div, float thresh,
       int* __restrict__ counters, const float4* __restrict__ surv, PairRec* __restrict__ pairs, int cap,
       float4* __restrict__ raws, float* __restrict__ dbg) {
    // dbg (optional, per SAMPLE): [n][5][8] = flag, x, y, z, vx, vy, vz, pdist -- per-stage parity tests
    // deformer MLP weights + both joint transform sets staged once per CTA
    __shared__ float s_w[32 * 19 + 32 + 32 * 32 + 32 + 3 * 32 + 3 + 1];
    __shared__ float s_A[NVR_JOINTS * 16], s_bigA[NVR_JOINTS * 16];
    float* sw0 = s_w; float* sb0 = sw0 + 32 * 19; float* sw1 = sb0 + 32; float* sb1 = sw1 + 32 * 32;
    float* sw2 = sb1 + 32; float* sb2 = sw2 + 3 * 32;
    for (int i = threadIdx.x; i < 32 * 19; i += blockDim.x) sw0[i] = dm_g.w0[i];
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) sw1[i] = dm_g.w1[i];
    for (int i = threadIdx.x; i < 3 * 32; i += blockDim.x) sw2[i] = dm_g.w2[i];
    if (threadIdx.x < 32) { sb0[threadIdx.x] = dm_g.b0[threadIdx.x]; sb1[threadIdx.x] = dm_g.b1[threadIdx.x]; }
    if (threadIdx.x < 3) sb2[threadIdx.x] = dm_g.b2[threadIdx.x];
    for (int i = threadIdx.x; i < NVR_JOINTS * 16; i += blockDim.x) { s_A[i] = fr.A[i]; s_bigA[i] = fr.bigA[i]; }
    __syncthreads();
    DeformerMlp dm = {sw0, sb0, sw1, sb1, sw2, sb2};
    const float frame_dim = fr.frame_dim[0];
    const int n_surv = counters[NVR_CTR_SURV];
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count so the ballots below see full warps
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n_surv; base += gridDim.x * blockDim.x) {
        const int s = base + lane;
        const bool live = s < n_surv;
        float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
        if (live) {
            const float4 sv = surv[s];
            p[0] = sv.x; p[1] = sv.y; p[2] = sv.z;
            const long long di = (long long)(__float_as_int(sv.w) / dir_div) * 3;
            const float wd[3] = {dirs[di], dirs[di + 1], dirs[di + 2]};
            nvr_dir_to_pose(fr.R, wd, d);
        }
#pragma unroll 1
        for (int part = 0; part < NVR_PARTS; ++part) {
            const int off = fr.part_off[part], cnt = fr.part_off[part + 1] - off;
            bool flag = false;
            PairRec rec;
            if (live) {
                Knn4 k;
                nvr_knn_init(k);
                nvr_knn_scan(fr.verts + off, cnt, p, k);
                float bw[NVR_JOINTS];
                const float pdist = nvr_knn_blend(k, fr.part_pbw + (long long)part * fr.maxlen * NVR_JOINTS, bw);
                flag = pdist < thresh;                             // inb_part_network_multiassign.py:90
                float* dr = dbg ? dbg + ((long long)__float_as_int(surv[s].w) * NVR_PARTS + part) * 8 : nullptr;
                if (dr) { dr[0] = flag ? 1.0f : 0.0f; dr[7] = pdist; }
                if (flag) {
                    float x0[3], v[3], r[3];
                    nvr_lbs_to_bigpose(bw, s_A, s_bigA, p, d, x0, v);
                    nvr_deformer_point(dg, dm, fr.tuv, frame_dim, x0, r);
                    rec.x = x0[0] + r[0]; rec.y = x0[1] + r[1]; rec.z = x0[2] + r[2];   // :113
                    rec.vx = v[0]; rec.vy = v[1]; rec.vz = v[2];
                    rec.surv = s; rec._pad = 0;
                    if (dr) { dr[1] = rec.x; dr[2] = rec.y; dr[3] = rec.z; dr[4] = v[0]; dr[5] = v[1]; dr[6] = v[2]; }
                } else {
                    raws[(long long)s * NVR_PARTS + part] = make_float4(0.f, 0.f, 0.f, 0.f);   // :201-202
                }
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, flag);
            if (ballot) {
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&counters[NVR_CTR_PAIR + part], __popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (flag) pairs[(long long)part * cap + wbase + __popc(ballot & ((1u << lane) - 1u))] = rec;
            }
        }
    }
}

// Deformer on explicit canonical points (Network.resd).
__global__ void k_deformer(FrameDev fr, GridDev dg, DeformerMlp dm, const float* __restrict__ x, long long n,
                           float* __restrict__ out) {
    const float frame_dim = fr.frame_dim[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x0[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
        float r[3];
        nvr_deformer_point(dg, dm, fr.tuv, frame_dim, x0, r);
        out[i * 3] = r[0]; out[i * 3 + 1] = r[1]; out[i * 3 + 2] = r[2];
    }
}

// -----------------------------------------------------------------------------------------
// THE gather: part grid embedding, 4 lanes per 64-byte row
// -----------------------------------------------------------------------------------------
// A warp works on 8 points at a time: lane = 4*point + quarter.  For every level the 8 corner rows
// (16 fp32 = 64 B each) are fetched as one 16-byte vector per lane, so each warp-wide load
// instruction covers 8 complete rows = 16 fully-used 32-byte sectors.  Per-feature trilinear sums
// are kept per lane, folded over the lane's 4 features and then over the 4 quarter-lanes.
// Input points: float x[3] at `xbase + i * xstride` (PairRec lists: stride 8; plain xyz: stride 3).
// The point count is *count_dev when non-null (device-side list length), else n_imm.  One launch per part.
__device__ __forceinline__ float4 ld_row_quarter(const float* tab, long long row, int q) {
    return __ldg(reinterpret_cast<const float4*>(tab + row * 16) + q);
}

__global__ void __launch_bounds__(256)
k_embed(GridDev g, const float* __restrict__ xb, int xstride, const int* __restrict__ count_dev, int n_imm,
        float* __restrict__ eb, int emb_stride) {
    const int n = count_dev ? *count_dev : n_imm;
    const int lane = threadIdx.x & 31, q = lane & 3;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 8; base < n; base += n_warps * 8) {
        const int pt = base + (lane >> 2);
        const bool live = pt < n;
        const int pi = live ? pt : n - 1;
        const float* xp = xb + (long long)pi * xstride;
        const float x[3] = {xp[0], xp[1], xp[2]};
        float u[3];
        nvr_normalise(g, x, u);
        float lev[NVR_LEVELS];
#pragma unroll
        for (int l = 0; l < NVR_LEVELS; ++l) {
            lev[l] = 0.0f;
            if (l < g.n_levels) {
                LevelCoord lc;
                nvr_level_coord(u, g.size[l], g.res[l], lc);
                const float* tab = nvr_level_table(g, l);
                float4 v[8];
                float w[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    v[c] = ld_row_quarter(tab, nvr_corner_row(g, l, lc, c), q);
                    w[c] = nvr_corner_weight(lc, c);
                }
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    a.x += w[c] * v[c].x; a.y += w[c] * v[c].y; a.z += w[c] * v[c].z; a.w += w[c] * v[c].w;
                }
                float s = (a.x + a.y) + (a.z + a.w);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                lev[l] = s;
            }
        }
        if (live) {
            float* o = eb + (long long)pt * emb_stride;
            // element e is written by quarter-lane e % 4
#pragma unroll
            for (int e = 0; e < 19; ++e) {
                if ((e & 3) == q) o[e] = e < 3 ? u[e] : lev[e - 3];
            }
        }
    }
}

// -----------------------------------------------------------------------------------------
// occ + rgb MLPs, fp32 register tiles                      part_base_network.py:44-63
// -----------------------------------------------------------------------------------------
#define MLP_TILE 128
#define MLP_LDA 72                 // row stride of the wide activation buffer (70 used)
#define MLP_LDH 68                 // row stride of the hidden buffer (64 used)

// Dense layer on a 128-row tile: out[r][o] = epi(bias[o] + sum_k in[r][k] * Wt[k][o]).
// Wt is the transposed weight in shared memory, row stride NP = 4*NQ (zero padded).
// Work item = (row group, output quad); a thread's rows are rg + i*G so neighbouring lanes touch
// neighbouring rows (conflict-free LDS).
template <int K, int NQ, int PPT, class Epi>
__device__ __forceinline__ void dense_layer(const float* __restrict__ in, int ldi, const float* __restrict__ Wt,
                                            const float* __restrict__ bias, Epi epi) {
    constexpr int G = MLP_TILE / PPT, NP = NQ * 4, ITEMS = G * NQ;
    for (int it = threadIdx.x; it < ITEMS; it += blockDim.x) {
        const int oq = it % NQ, rg = it / NQ;
        float acc[PPT][4];
        const float4 b4 = *reinterpret_cast<const float4*>(bias + oq * 4);
#pragma unroll
        for (int i = 0; i < PPT; ++i) { acc[i][0] = b4.x; acc[i][1] = b4.y; acc[i][2] = b4.z; acc[i][3] = b4.w; }
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const float4 w4 = *reinterpret_cast<const float4*>(Wt + k * NP + oq * 4);
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                const float a = in[(rg + i * G) * ldi + k];
                acc[i][0] += a * w4.x; acc[i][1] += a * w4.y; acc[i][2] += a * w4.z; acc[i][3] += a * w4.w;
            }
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) epi(rg + i * G, oq * 4 + j, acc[i][j]);
    }
}

// shared-memory plan (floats)
#define MLP_S_A 0                                   // [128][72] wide activations / rgb input
#define MLP_S_H (MLP_S_A + MLP_TILE * MLP_LDA)      // [128][68] hidden
#define MLP_S_W0 (MLP_S_H + MLP_TILE * MLP_LDH)     // occ0^T  [19][64]
#define MLP_S_W1 (MLP_S_W0 + 19 * 64)               // occ1^T  [64][20]
#define MLP_S_W2 (MLP_S_W1 + 64 * 20)               // rgb0^T  [70][64]
#define MLP_S_W3 (MLP_S_W2 + 70 * 64)               // rgb1^T  [64][64]  (3-linear parts)
#define MLP_S_W4 (MLP_S_W3 + 64 * 64)               // rgb last^T [64][4]
#define MLP_S_B (MLP_S_W4 + 64 * 4)                 // biases: 64 + 20 + 64 + 64 + 4
#define MLP_S_OCC (MLP_S_B + 64 + 20 + 64 + 64 + 4) // [128] occupancy
#define MLP_S_SURV (MLP_S_OCC + MLP_TILE)           // [128] int survivor slots
#define MLP_S_END (MLP_S_SURV + MLP_TILE)
#define MLP_SMEM_BYTES (MLP_S_END * 4)

__device__ __forceinline__ void stage_transposed(float* dst, int np, const LinearDev& L) {
    // dst[k][o] = W[o][k], zero padding for o >= out
    for (int i = threadIdx.x; i < L.in * np; i += blockDim.x) {
        const int k = i / np, o = i - k * np;
        dst[i] = o < L.out ? L.w[o * L.in + k] : 0.0f;
    }
}
__device__ __forceinline__ void stage_bias(float* dst, int np, const LinearDev& L) {
    for (int i = threadIdx.x; i < np; i += blockDim.x) dst[i] = i < L.out ? L.b[i] : 0.0f;
}

// One launch per part; `pl` / `el` are that part's pair list and embedding rows, *count_dev its length.
__global__ void __launch_bounds__(256)
k_mlp(PartMlpDev pm, int part, const long long* __restrict__ latent_index, const int* __restrict__ count_dev,
      const PairRec* __restrict__ pl, const float* __restrict__ el, float4* __restrict__ raws, int out_stride) {
    extern __shared__ __align__(16) float sm[];
    float* sA = sm + MLP_S_A; float* sH = sm + MLP_S_H;
    float* sB = sm + MLP_S_B;
    float* b_occ0 = sB; float* b_occ1 = sB + 64; float* b_rgb0 = sB + 84; float* b_rgb1 = sB + 148; float* b_rgbL = sB + 212;
    float* sOcc = sm + MLP_S_OCC;
    int* sSurv = reinterpret_cast<int*>(sm + MLP_S_SURV);
    {
        const int n = *count_dev;
        const int n_tiles = (n + MLP_TILE - 1) / MLP_TILE;
        if ((int)blockIdx.x >= n_tiles) return;                   // block-uniform
        const bool three = pm.n_rgb == 3;
        stage_transposed(sm + MLP_S_W0, 64, pm.occ[0]);
        stage_transposed(sm + MLP_S_W1, 20, pm.occ[1]);
        stage_transposed(sm + MLP_S_W2, 64, pm.rgb[0]);
        if (three) stage_transposed(sm + MLP_S_W3, 64, pm.rgb[1]);
        stage_transposed(sm + MLP_S_W4, 4, pm.rgb[pm.n_rgb - 1]);
        stage_bias(b_occ0, 64, pm.occ[0]);
        stage_bias(b_occ1, 20, pm.occ[1]);
        stage_bias(b_rgb0, 64, pm.rgb[0]);
        if (three) stage_bias(b_rgb1, 64, pm.rgb[1]);
        stage_bias(b_rgbL, 4, pm.rgb[pm.n_rgb - 1]);
        long long li = latent_index[0];
        li = li < 0 ? 0 : (li >= pm.n_latent ? pm.n_latent - 1 : li);
        const float* lat = pm.latent + li * 8;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int t0 = tile * MLP_TILE;
            __syncthreads();
            // ---- build the 70-wide input rows: [embed 19 | posenc 27 | (feat 16, later) | latent 8]
            for (int i = threadIdx.x; i < MLP_TILE * 19; i += blockDim.x) {
                const int r = i / 19, c = i - r * 19;
                const int pr = min(t0 + r, n - 1);
                sA[r * MLP_LDA + c] = el[(long long)pr * NVR_EMB_STRIDE + c];
            }
            if (threadIdx.x < MLP_TILE) {
                const int r = threadIdx.x, pr = min(t0 + r, n - 1);
                const PairRec rec = pl[pr];
                const float v[3] = {rec.vx, rec.vy, rec.vz};
                nvr_posenc27(v, sA + r * MLP_LDA + 19);           // part_base_network.py:54
#pragma unroll
                for (int c = 0; c < 8; ++c) sA[r * MLP_LDA + 62 + c] = lat[c];   // :55
                sSurv[r] = rec.surv;
            }
            __syncthreads();
            // ---- occ MLP: 19 -> 64 (softplus) -> 17
            dense_layer<19, 16, 8>(sA, MLP_LDA, sm + MLP_S_W0, b_occ0,
                                   [&](int r, int o, float v) { sH[r * MLP_LDH + o] = nvr_softplus(v); });
            __syncthreads();
            dense_layer<64, 5, 2>(sH, MLP_LDH, sm + MLP_S_W1, b_occ1, [&](int r, int o, float v) {
                if (o == 0) sOcc[r] = 1.0f - expf(-nvr_softplus(v));         // :51
                else if (o < 17) sA[r * MLP_LDA + 45 + o] = v;              // feature = hidden[1:], :52
            });
            __syncthreads();
            // ---- rgb MLP: 70 -> 64 [-> 64] -> 3 (sigmoid)
            dense_layer<70, 16, 8>(sA, MLP_LDA, sm + MLP_S_W2, b_rgb0,
                                   [&](int r, int o, float v) { sH[r * MLP_LDH + o] = nvr_softplus(v); });
            __syncthreads();
            const float* last_in = sH;
            int last_ld = MLP_LDH;
            if (three) {
                dense_layer<64, 16, 8>(sH, MLP_LDH, sm + MLP_S_W3, b_rgb1,
                                       [&](int r, int o, float v) { sA[r * MLP_LDA + o] = nvr_softplus(v); });
                __syncthreads();
                last_in = sA; last_ld = MLP_LDA;
            }
            float* sOut = three ? sH : sA;                         // [128][4] rgb staging in the free buffer
            dense_layer<64, 1, 1>(last_in, last_ld, sm + MLP_S_W4, b_rgbL,
                                  [&](int r, int o, float v) { sOut[r * 4 + o] = nvr_sigmoid(v); });   // :58
            __syncthreads();
            if (threadIdx.x < MLP_TILE && t0 + threadIdx.x < n) {
                const int r = threadIdx.x;
                raws[(long long)sSurv[r] * out_stride + part] =
                    make_float4(sOut[r * 4], sOut[r * 4 + 1], sOut[r * 4 + 2], sOcc[r]);   // raw = [rgb, occ] :60
            }
        }
    }
}

// Stand-alone MLP entry (nvr_part_mlp): wrap explicit view directions into a pair list.
__global__ void k_make_pairs(const float* __restrict__ dirs, int n, PairRec* __restrict__ pl, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = n;
    if (i < n) {
        PairRec r;
        r.x = r.y = r.z = 0.f;
        r.vx = dirs[i * 3]; r.vy = dirs[i * 3 + 1]; r.vz = dirs[i * 3 + 2];
        r.surv = i; r._pad = 0;
        pl[i] = r;
    }
}

// -----------------------------------------------------------------------------------------
// resolve: arg-max over parts, scatter to samples, composite along rays
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ float4 fuse_parts(const float4* __restrict__ raws, int slot) {
    // inb_part_network_multiassign.py:253-255: raw of the part with the largest occupancy (first on ties)
    float4 best = raws[(long long)slot * NVR_PARTS];
#pragma unroll
    for (int p = 1; p < NVR_PARTS; ++p) {
        const float4 r = raws[(long long)slot * NVR_PARTS + p];
        if (r.w > best.w) best = r;
    }
    return best;
}

__global__ void k_resolve_points(const int* __restrict__ surv_of_sample, const float4* __restrict__ raws, long long n,
                                 float4* __restrict__ raw_out, float* __restrict__ occ_out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int slot = surv_of_sample[i];
        const float4 r = slot >= 0 ? fuse_parts(raws, slot) : make_float4(0.f, 0.f, 0.f, 0.f);
        raw_out[i] = r;
        if (occ_out) occ_out[i] = r.w;
    }
}

// One warp per ray; samples are walked 32 at a time with a shuffle product-scan of (1 - alpha)
// (net_utils.py:12-15 with epsilon = 0; :39-41).
__global__ void __launch_bounds__(256)
k_resolve_rays(const int* __restrict__ surv_of_sample, const float4* __restrict__ raws, long long n_rays, int S,
               float* __restrict__ rgb_map, float* __restrict__ acc_map, float4* __restrict__ raw_out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long ray = warp; ray < n_rays; ray += n_warps) {
        float carry = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, ca = 0.f;
        for (int k0 = 0; k0 < S; k0 += 32) {
            const int k = k0 + lane;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < S) {
                const int slot = surv_of_sample[ray * S + k];
                if (slot >= 0) r = fuse_parts(raws, slot);
                if (raw_out) raw_out[ray * S + k] = r;
            }
            float incl = 1.0f - r.w;                              // inclusive product of (1 - alpha)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl *= up;
            }
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float w = r.w * (carry * excl);                 // alpha_i * prod_{j<i} (1 - alpha_j)
            cr += w * r.x; cg += w * r.y; cb += w * r.z; ca += w;
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            cr += __shfl_xor_sync(0xffffffffu, cr, d); cg += __shfl_xor_sync(0xffffffffu, cg, d);
            cb += __shfl_xor_sync(0xffffffffu, cb, d); ca += __shfl_xor_sync(0xffffffffu, ca, d);
        }
        if (lane == 0) {
            rgb_map[ray * 3] = cr; rgb_map[ray * 3 + 1] = cg; rgb_map[ray * 3 + 2] = cb;
            acc_map[ray] = ca;
        }
    }
}

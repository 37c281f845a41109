// nvr_train.cuh -- backward of the hot path: the gradients the reference gets
// from autograd through
//     raw (N,4) -> arg-max part fusion -> part MLPs -> part grid (tables + canonical xyz) -> xyz = x0 + resd
//     resd -> deformer MLP + deformer grid
// (inb_part_network_multiassign.py:126-168, 194-256; part_base_network.py:44-63;
//  part_base_embedder.py:106-174; uv_deformer.py:23-45).  KNN weights, blended transforms and x0 carry no
// gradient (`torch.no_grad()` block :85-90 and constant inputs), the canonical view direction neither.
//
// A training step is small (the shipped configs: 1024-4096 rays x 64 samples, a few 10^4 pairs with a non-zero
// gradient), so these kernels favour clarity: 64-pair tiles, fp32 FFMA with weights, activations and the
// weight-gradient accumulators in shared memory; one atomicAdd per weight per CTA at the end.
//
//   k_bwd_select     per flagged pair: gradient of its raw=[rgb,occ] (only the arg-max part of a survivor
//                    receives d raw; d tocc optional) -> compact per-part lists of pairs with a gradient
//   k_mlp_bwd        tile: recompute the part MLPs forward, backprop; dW/db accumulate, d emb out
//   k_embed_bwd      pair x level: table gradients (red.global.add) and d canonical xyz
//   k_deformer_bwd   tile: recompute the deformer forward, backprop; dW/db, deformer grid gradients
//   k_composite_fwd / k_composite_bwd   net_utils.py:12-44 on explicit raw (training uses jittered samples)
#pragma once
#include <cuda_runtime.h>

#include "nvr_kernels.cuh"

struct __align__(16) GradRec {     // a pair with a non-zero output gradient
    int pair;                      // index into the part's pair list
    int surv;
    float d[4];                    // d loss / d [r, g, b, occ]
    int _pad[2];
};

struct LinearGrad { float* w; float* b; };
struct PartMlpGrad {
    LinearGrad occ[2];
    LinearGrad rgb[3];
    float* latent;                 // (n_latent, 8)
};
struct GridGrad { float* dense; float* hash; };

#define BT 64                      // pairs per backward tile
#define BTHREADS 256

// ---- generic 64-row tile pieces (shared memory operands) -------------------------------------------
// out[r][o] = b[o] + sum_k in[r][k] W[o][k]      (W row-major (N,K) as nn.Linear stores it)
__device__ __forceinline__ void tile_linear(const float* in, int ldi, int K, const float* W, const float* b, int N,
                                            float* out, int ldo) {
    for (int it = threadIdx.x; it < BT * N; it += blockDim.x) {
        const int r = it & (BT - 1), o = it / BT;
        float acc = b[o];
        const float* x = in + r * ldi;
        const float* w = W + o * K;
        for (int k = 0; k < K; ++k) acc += x[k] * w[k];
        out[r * ldo + o] = acc;
    }
}
// din[r][k] (+)= sum_o dout[r][o] W[o][k]
__device__ __forceinline__ void tile_input_grad(const float* dout, int ldo, int N, const float* W, int K, float* din, int ldi,
                                                bool accumulate) {
    for (int it = threadIdx.x; it < BT * K; it += blockDim.x) {
        const int r = it & (BT - 1), k = it / BT;
        float acc = accumulate ? din[r * ldi + k] : 0.0f;
        const float* d = dout + r * ldo;
        for (int o = 0; o < N; ++o) acc += d[o] * W[o * K + k];
        din[r * ldi + k] = acc;
    }
}
// dW[o][k] += sum_r dout[r][o] in[r][k];  db[o] += sum_r dout[r][o]      (rows >= n_rows are skipped)
__device__ __forceinline__ void tile_weight_grad(const float* dout, int ldo, int N, const float* in, int ldi, int K, int n_rows,
                                                 float* dW, float* db) {
    for (int it = threadIdx.x; it < N * K; it += blockDim.x) {
        const int o = it / K, k = it - o * K;
        float acc = 0.0f;
        for (int r = 0; r < n_rows; ++r) acc += dout[r * ldo + o] * in[r * ldi + k];
        dW[it] += acc;
    }
    for (int o = threadIdx.x; o < N; o += blockDim.x) {
        float acc = 0.0f;
        for (int r = 0; r < n_rows; ++r) acc += dout[r * ldo + o];
        db[o] += acc;
    }
}
__device__ __forceinline__ void tile_copy_in(float* dst, const float* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void tile_zero(float* dst, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = 0.0f;
}
__device__ __forceinline__ void tile_flush(float* gdst, const float* acc, int n) {
    if (!gdst) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = acc[i];
        if (v != 0.0f) atomicAdd(gdst + i, v);
    }
}
// softplus'(z) = sigmoid(z) = 1 - exp(-softplus(z)), from the stored activation
__device__ __forceinline__ float softplus_grad_from_out(float sp) { return 1.0f - expf(-sp); }

// ---------------------------------------------------------------------------------------------------
// which pairs receive a gradient                                   inb_part_network_multiassign.py:253-255
// ---------------------------------------------------------------------------------------------------
// d_raw: (n_samples, 4) gradient of the scattered raw; d_tocc: optional (survivors, 5) gradient of every pair's
// occupancy, row rank_of_slot[slot] (ascending sample order; the slot itself when rank_of_slot is null).  blockIdx.y = part.
__global__ void __launch_bounds__(256)
k_bwd_select(const int* __restrict__ counters, const PairRec* __restrict__ pairs, int cap, const float4* __restrict__ surv,
             const float4* __restrict__ raws, const float4* __restrict__ d_raw, const float* __restrict__ d_tocc,
             const int* __restrict__ rank_of_slot, GradRec* __restrict__ glist, int* __restrict__ gcount) {
    const int part = blockIdx.y;
    const int n = counters[NVR_CTR_PAIR + part];
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        bool has = false;
        GradRec g;
        if (i < n) {
            const int s = pairs[(long long)part * cap + i].surv;
            // arg-max part of this survivor, first index on ties (fuse_parts)
            int best = 0;
            float bo = raws[(long long)s * NVR_PARTS].w;
#pragma unroll
            for (int p = 1; p < NVR_PARTS; ++p) {
                const float o = raws[(long long)s * NVR_PARTS + p].w;
                if (o > bo) { bo = o; best = p; }
            }
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            if (best == part) {
                const float4 dr = d_raw[__float_as_int(surv[s].w)];
                d[0] = dr.x; d[1] = dr.y; d[2] = dr.z; d[3] = dr.w;
            }
            if (d_tocc) d[3] += d_tocc[(long long)(rank_of_slot ? rank_of_slot[s] : s) * NVR_PARTS + part];
            has = d[0] != 0.f || d[1] != 0.f || d[2] != 0.f || d[3] != 0.f;
            g.pair = i; g.surv = s; g.d[0] = d[0]; g.d[1] = d[1]; g.d[2] = d[2]; g.d[3] = d[3]; g._pad[0] = g._pad[1] = 0;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, has);
        if (ballot) {
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(&gcount[part], __popc(ballot));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (has) glist[(long long)part * cap + wbase + __popc(ballot & ((1u << lane) - 1u))] = g;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// part MLP backward                                                          part_base_network.py:44-63
// ---------------------------------------------------------------------------------------------------
// shared memory plan (floats); row strides are odd so that lanes walking rows hit different banks
#define MB_LDE 21
#define MB_LDH 65
#define MB_LDO 17
#define MB_LDX 71
#define MB_W0 0                               // (64,19)
#define MB_W1 (MB_W0 + 64 * 19)               // (17,64)
#define MB_W2 (MB_W1 + 17 * 64)               // (64,70)
#define MB_W3 (MB_W2 + 64 * 70)               // (64,64)
#define MB_W4 (MB_W3 + 64 * 64)               // (3,64)
#define MB_B (MB_W4 + 3 * 64)                 // b0 64 | b1 17 | b2 64 | b3 64 | b4 3  (+ pad) = 216
#define MB_WEND (MB_B + 216)
#define MB_DW (MB_WEND)                       // gradient accumulators, same layout as [MB_W0, MB_WEND)
#define MB_DLAT (MB_DW + MB_WEND)             // 8
#define MB_E (MB_DLAT + 8)                    // [64][21] embedding, later d emb
#define MB_H (MB_E + BT * MB_LDE)             // [64][65] h
#define MB_O (MB_H + BT * MB_LDH)             // [64][17] o, later d o
#define MB_X (MB_O + BT * MB_LDO)             // [64][71] rgb input, later d input
#define MB_G (MB_X + BT * MB_LDX)             // [64][65] g
#define MB_G3 (MB_G + BT * MB_LDH)            // [64][65] second rgb hidden
#define MB_D (MB_G3 + BT * MB_LDH)            // [64][65] gradient scratch
#define MB_Y (MB_D + BT * MB_LDH)             // [64][4] rgb logits -> d logits
#define MB_END (MB_Y + BT * 4)
#define MB_SMEM_BYTES (MB_END * 4)

// One launch per part.  d_emb: [gcap][NVR_EMB_STRIDE] gradient of each listed pair's 19-D embedding (in list order).
__global__ void __launch_bounds__(BTHREADS, 1)
k_mlp_bwd(PartMlpDev pm, PartMlpGrad pg, const long long* __restrict__ latent_index, const int* __restrict__ gcount_dev,
          const GradRec* __restrict__ gl, const PairRec* __restrict__ pl, const float* __restrict__ el,
          float* __restrict__ d_emb) {
    extern __shared__ __align__(128) float sm[];
    const int n = *gcount_dev;
    const int n_tiles = (n + BT - 1) / BT;
    if ((int)blockIdx.x >= n_tiles) return;
    const bool three = pm.n_rgb == 3;
    float* W0 = sm + MB_W0; float* W1 = sm + MB_W1; float* W2 = sm + MB_W2; float* W3 = sm + MB_W3; float* W4 = sm + MB_W4;
    float* b0 = sm + MB_B; float* b1 = b0 + 64; float* b2 = b1 + 17; float* b3 = b2 + 64; float* b4 = b3 + 64;
    float* dW0 = sm + MB_DW + MB_W0; float* dW1 = sm + MB_DW + MB_W1; float* dW2 = sm + MB_DW + MB_W2;
    float* dW3 = sm + MB_DW + MB_W3; float* dW4 = sm + MB_DW + MB_W4;
    float* db0 = sm + MB_DW + MB_B; float* db1 = db0 + 64; float* db2 = db1 + 17; float* db3 = db2 + 64; float* db4 = db3 + 64;
    float* dlat = sm + MB_DLAT;
    float* E = sm + MB_E; float* H = sm + MB_H; float* O = sm + MB_O; float* X = sm + MB_X; float* G = sm + MB_G;
    float* G3 = sm + MB_G3; float* D = sm + MB_D; float* Y = sm + MB_Y;
    tile_copy_in(W0, pm.occ[0].w, 64 * 19); tile_copy_in(W1, pm.occ[1].w, 17 * 64); tile_copy_in(W2, pm.rgb[0].w, 64 * 70);
    if (three) tile_copy_in(W3, pm.rgb[1].w, 64 * 64);
    tile_copy_in(W4, pm.rgb[pm.n_rgb - 1].w, 3 * 64);
    tile_copy_in(b0, pm.occ[0].b, 64); tile_copy_in(b1, pm.occ[1].b, 17); tile_copy_in(b2, pm.rgb[0].b, 64);
    if (three) tile_copy_in(b3, pm.rgb[1].b, 64);
    tile_copy_in(b4, pm.rgb[pm.n_rgb - 1].b, 3);
    tile_zero(sm + MB_DW, MB_WEND + 8);
    long long li = latent_index[0];
    li = li < 0 ? 0 : (li >= pm.n_latent ? pm.n_latent - 1 : li);
    const float* lat = pm.latent + li * 8;
    __syncthreads();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int t0 = tile * BT, rows = min(BT, n - t0);
        // ---- inputs: embedding, posenc, latent; output gradient
        for (int i = threadIdx.x; i < BT * 19; i += blockDim.x) {
            const int r = i / 19, c = i - r * 19;
            const GradRec g = gl[min(t0 + r, n - 1)];
            const float v = el[(long long)g.pair * NVR_EMB_STRIDE + c];
            E[r * MB_LDE + c] = v;
            X[r * MB_LDX + c] = v;
        }
        if (threadIdx.x < BT) {
            const int r = threadIdx.x;
            const GradRec g = gl[min(t0 + r, n - 1)];
            const PairRec rec = pl[g.pair];
            const float v[3] = {rec.vx, rec.vy, rec.vz};
            nvr_posenc27(v, X + r * MB_LDX + 19);
#pragma unroll
            for (int c = 0; c < 8; ++c) X[r * MB_LDX + 62 + c] = lat[c];
        }
        __syncthreads();
        // ---- forward recompute
        tile_linear(E, MB_LDE, 19, W0, b0, 64, H, MB_LDH);
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 64; i += blockDim.x) { float& h = H[(i & (BT - 1)) * MB_LDH + i / BT]; h = nvr_softplus(h); }
        __syncthreads();
        tile_linear(H, MB_LDH, 64, W1, b1, 17, O, MB_LDO);
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 16; i += blockDim.x) {
            const int r = i & (BT - 1), c = i / BT;
            X[r * MB_LDX + 46 + c] = O[r * MB_LDO + 1 + c];              // feature = hidden[1:]   (:52)
        }
        __syncthreads();
        tile_linear(X, MB_LDX, 70, W2, b2, 64, G, MB_LDH);
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 64; i += blockDim.x) { float& g = G[(i & (BT - 1)) * MB_LDH + i / BT]; g = nvr_softplus(g); }
        __syncthreads();
        float* GL = G;                                                   // last hidden activation
        if (three) {
            tile_linear(G, MB_LDH, 64, W3, b3, 64, G3, MB_LDH);
            __syncthreads();
            for (int i = threadIdx.x; i < BT * 64; i += blockDim.x) { float& g = G3[(i & (BT - 1)) * MB_LDH + i / BT]; g = nvr_softplus(g); }
            __syncthreads();
            GL = G3;
        }
        tile_linear(GL, MB_LDH, 64, W4, b4, 3, Y, 4);
        __syncthreads();
        // ---- backward.  d logits = d rgb * rgb (1 - rgb); rows beyond the list get zero gradient
        if (threadIdx.x < BT) {
            const int r = threadIdx.x;
            const GradRec g = gl[min(t0 + r, n - 1)];
            const float live = r < rows ? 1.0f : 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s = nvr_sigmoid(Y[r * 4 + c]);
                Y[r * 4 + c] = live * g.d[c] * s * (1.0f - s);
            }
            // occ = 1 - exp(-softplus(o0))  ->  d o0 = d occ * exp(-softplus(o0)) * sigmoid(o0)
            const float o0 = O[r * MB_LDO];
            const float sp = nvr_softplus(o0);
            Y[r * 4 + 3] = live * g.d[3] * expf(-sp) * nvr_sigmoid(o0);
        }
        __syncthreads();
        tile_weight_grad(Y, 4, 3, GL, MB_LDH, 64, rows, dW4, db4);
        tile_input_grad(Y, 4, 3, W4, 64, D, MB_LDH, false);             // d g_last
        __syncthreads();
        if (three) {
            for (int i = threadIdx.x; i < BT * 64; i += blockDim.x) {
                const int idx = (i & (BT - 1)) * MB_LDH + i / BT;
                D[idx] *= softplus_grad_from_out(G3[idx]);               // d z3
            }
            __syncthreads();
            tile_weight_grad(D, MB_LDH, 64, G, MB_LDH, 64, rows, dW3, db3);
            tile_input_grad(D, MB_LDH, 64, W3, 64, G3, MB_LDH, false);  // d g   (G3 is free now)
            __syncthreads();
            for (int i = threadIdx.x; i < BT * 64; i += blockDim.x) {
                const int idx = (i & (BT - 1)) * MB_LDH + i / BT;
                D[idx] = G3[idx] * softplus_grad_from_out(G[idx]);       // d z2
            }
        } else {
            for (int i = threadIdx.x; i < BT * 64; i += blockDim.x) {
                const int idx = (i & (BT - 1)) * MB_LDH + i / BT;
                D[idx] *= softplus_grad_from_out(G[idx]);                // d z2
            }
        }
        __syncthreads();
        tile_weight_grad(D, MB_LDH, 64, X, MB_LDX, 70, rows, dW2, db2);
        __syncthreads();
        tile_input_grad(D, MB_LDH, 64, W2, 70, X, MB_LDX, false);       // d [e | pe | feat | lat]   (X overwritten)
        __syncthreads();
        // d o = [d o0 | d feat];  d latent row
        for (int i = threadIdx.x; i < BT * 17; i += blockDim.x) {
            const int r = i & (BT - 1), c = i / BT;
            O[r * MB_LDO + c] = c == 0 ? Y[r * 4 + 3] : X[r * MB_LDX + 45 + c];
        }
        if (threadIdx.x < 8) {
            float acc = 0.0f;
            for (int r = 0; r < rows; ++r) acc += X[r * MB_LDX + 62 + threadIdx.x];
            dlat[threadIdx.x] += acc;
        }
        __syncthreads();
        tile_weight_grad(O, MB_LDO, 17, H, MB_LDH, 64, rows, dW1, db1);
        tile_input_grad(O, MB_LDO, 17, W1, 64, D, MB_LDH, false);       // d h
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 64; i += blockDim.x) {
            const int idx = (i & (BT - 1)) * MB_LDH + i / BT;
            D[idx] *= softplus_grad_from_out(H[idx]);                    // d z0
        }
        __syncthreads();
        tile_weight_grad(D, MB_LDH, 64, E, MB_LDE, 19, rows, dW0, db0);
        __syncthreads();
        tile_input_grad(D, MB_LDH, 64, W0, 19, X, MB_LDX, true);        // d e += occ branch  (X[:, 0:19] holds the rgb branch)
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 19; i += blockDim.x) {
            const int r = i / 19, c = i - r * 19;
            d_emb[(long long)(t0 + r) * NVR_EMB_STRIDE + c] = X[r * MB_LDX + c];
        }
        __syncthreads();
    }
    tile_flush(pg.occ[0].w, dW0, 64 * 19); tile_flush(pg.occ[0].b, db0, 64);
    tile_flush(pg.occ[1].w, dW1, 17 * 64); tile_flush(pg.occ[1].b, db1, 17);
    tile_flush(pg.rgb[0].w, dW2, 64 * 70); tile_flush(pg.rgb[0].b, db2, 64);
    if (three) { tile_flush(pg.rgb[1].w, dW3, 64 * 64); tile_flush(pg.rgb[1].b, db3, 64); }
    tile_flush(pg.rgb[pm.n_rgb - 1].w, dW4, 3 * 64); tile_flush(pg.rgb[pm.n_rgb - 1].b, db4, 3);
    if (pg.latent) tile_flush(pg.latent + li * 8, dlat, 8);
}

// ---------------------------------------------------------------------------------------------------
// part grid backward                                                     part_base_embedder.py:106-174
// ---------------------------------------------------------------------------------------------------
// 16 lanes per listed pair, lane = level.  out_l = sum_f sum_c w_c t[c][f]:
//   d t[c][f] = w_c * d out_l   (same value for the 16 features of a row: four red.global.add.v4.f32)
//   d o_a     = d out_l * sum_c (d w_c / d o_a) * S_c,  S_c = sum_f t[c][f];  o = u/size - i0, u = (x - b0)/(b1 - b0)
// The xyz passthrough adds d emb[0:3] / (b1 - b0).  d_x: [gcap][3] in list order.
__device__ __forceinline__ void red_add_v4(float* p, float v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_embed_bwd(GridDev g, GridGrad gg, const int* __restrict__ gcount_dev, const GradRec* __restrict__ gl,
            const PairRec* __restrict__ pl, const float* __restrict__ d_emb, float* __restrict__ d_x) {
    const int n = *gcount_dev;
    const int l = threadIdx.x & 15;
    const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, n_grp = (gridDim.x * blockDim.x) >> 4;
    for (int base = grp - ((threadIdx.x >> 4) & 1); base < n; base += n_grp) {    // both half-warps iterate together
        const int i = base + ((threadIdx.x >> 4) & 1);
        const bool live = i < n;
        const int ii = live ? i : n - 1;
        const PairRec rec = pl[gl[ii].pair];
        const float x[3] = {rec.x, rec.y, rec.z};
        float u[3];
        nvr_normalise(g, x, u);
        float du[3] = {0.f, 0.f, 0.f};
        if (l < g.n_levels) {
            const float dl = live ? d_emb[(long long)ii * NVR_EMB_STRIDE + 3 + l] : 0.0f;
            LevelCoord lc;
            nvr_level_coord(u, g.size[l], g.res[l], lc);
            float* gtab = l < g.start_hash ? gg.dense : gg.hash;
            const float* tab = nvr_level_table(g, l);
            const float wx[2] = {1.0f - lc.o[0], lc.o[0]}, wy[2] = {1.0f - lc.o[1], lc.o[1]}, wz[2] = {1.0f - lc.o[2], lc.o[2]};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const long long row = nvr_corner_row(g, l, lc, c);
                const int bx = (c >> 2) & 1, by = (c >> 1) & 1, bz = c & 1;
                const float w = (wx[bx] * wy[by]) * wz[bz];
                if (dl != 0.0f && gtab) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) red_add_v4(gtab + row * 16 + q * 4, w * dl);
                }
                const float4* t4 = reinterpret_cast<const float4*>(tab + row * 16);
                float S = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float4 t = __ldg(t4 + q); S += (t.x + t.y) + (t.z + t.w); }
                du[0] += (bx ? 1.0f : -1.0f) * wy[by] * wz[bz] * S;
                du[1] += (by ? 1.0f : -1.0f) * wx[bx] * wz[bz] * S;
                du[2] += (bz ? 1.0f : -1.0f) * wx[bx] * wy[by] * S;
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) du[a] = du[a] * dl / g.size[l];
        }
#pragma unroll
        for (int d = 8; d > 0; d >>= 1)
#pragma unroll
            for (int a = 0; a < 3; ++a) du[a] += __shfl_xor_sync(0xffffffffu, du[a], d);
        if (live && l < 3)
            d_x[(long long)i * 3 + l] = (du[l] + d_emb[(long long)i * NVR_EMB_STRIDE + l]) / (g.bounds[3 + l] - g.bounds[l]);
    }
}

// ---------------------------------------------------------------------------------------------------
// deformer backward                                                              uv_deformer.py:23-45
// ---------------------------------------------------------------------------------------------------
// Work list: n points x0 (n,3) with gradient d_r (n,3) of their residual.  No gradient flows into x0.
#define DB_LDE 21
#define DB_LDH 33
#define DB_W0 0                               // (32,19)
#define DB_W1 (DB_W0 + 32 * 19)               // (32,32)
#define DB_W2 (DB_W1 + 32 * 32)               // (3,32)
#define DB_B (DB_W2 + 3 * 32)                 // 32 | 32 | 3 (+1)
#define DB_WEND (DB_B + 68)
#define DB_DW DB_WEND
#define DB_E (DB_DW + DB_WEND)                // [64][21]
#define DB_H1 (DB_E + BT * DB_LDE)            // [64][33]
#define DB_H2 (DB_H1 + BT * DB_LDH)
#define DB_D (DB_H2 + BT * DB_LDH)            // [64][33] gradient scratch
#define DB_Y (DB_D + BT * DB_LDH)             // [64][4]
#define DB_END (DB_Y + BT * 4)
#define DB_SMEM_BYTES (DB_END * 4)

struct DeformerGrad { float *w0, *b0, *w1, *b1, *w2, *b2; };

__global__ void __launch_bounds__(BTHREADS)
k_deformer_bwd(FrameDev fr, GridDev dg, DeformerMlp dm, DeformerGrad dgr, GridGrad gg, const float* __restrict__ x0,
               const float* __restrict__ d_r, const int* __restrict__ count_dev, int n_imm) {
    extern __shared__ __align__(128) float sm[];
    const int n = count_dev ? *count_dev : n_imm;
    const int n_tiles = (n + BT - 1) / BT;
    if ((int)blockIdx.x >= n_tiles) return;
    float* W0 = sm + DB_W0; float* W1 = sm + DB_W1; float* W2 = sm + DB_W2;
    float* b0 = sm + DB_B; float* b1 = b0 + 32; float* b2 = b1 + 32;
    float* dW0 = sm + DB_DW + DB_W0; float* dW1 = sm + DB_DW + DB_W1; float* dW2 = sm + DB_DW + DB_W2;
    float* db0 = sm + DB_DW + DB_B; float* db1 = db0 + 32; float* db2 = db1 + 32;
    float* E = sm + DB_E; float* H1 = sm + DB_H1; float* H2 = sm + DB_H2; float* D = sm + DB_D; float* Y = sm + DB_Y;
    tile_copy_in(W0, dm.w0, 32 * 19); tile_copy_in(W1, dm.w1, 32 * 32); tile_copy_in(W2, dm.w2, 3 * 32);
    tile_copy_in(b0, dm.b0, 32); tile_copy_in(b1, dm.b1, 32); tile_copy_in(b2, dm.b2, 3);
    tile_zero(sm + DB_DW, DB_WEND);
    const float frame_dim = fr.frame_dim[0];
    __syncthreads();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int t0 = tile * BT, rows = min(BT, n - t0);
        if (threadIdx.x < BT) {
            const int r = threadIdx.x, i = min(t0 + r, n - 1);
            const float p[3] = {x0[(long long)i * 3], x0[(long long)i * 3 + 1], x0[(long long)i * 3 + 2]};
            float uvt[3];
            nvr_sample_volume(fr.tuv, p, 0, 2, uvt);
            uvt[2] = frame_dim;
            nvr_embed_point<2>(dg, uvt, E + r * DB_LDE, 1);
        }
        __syncthreads();
        tile_linear(E, DB_LDE, 19, W0, b0, 32, H1, DB_LDH);
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 32; i += blockDim.x) { float& h = H1[(i & (BT - 1)) * DB_LDH + i / BT]; h = nvr_softplus(h); }
        __syncthreads();
        tile_linear(H1, DB_LDH, 32, W1, b1, 32, H2, DB_LDH);
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 32; i += blockDim.x) { float& h = H2[(i & (BT - 1)) * DB_LDH + i / BT]; h = nvr_softplus(h); }
        __syncthreads();
        tile_linear(H2, DB_LDH, 32, W2, b2, 3, Y, 4);
        __syncthreads();
        if (threadIdx.x < BT) {                                         // r = 0.05 tanh(y)  ->  d y = d r * 0.05 (1 - tanh^2 y)
            const int r = threadIdx.x, i = min(t0 + r, n - 1);
            const float live = r < rows ? 1.0f : 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float t = tanhf(Y[r * 4 + c]);
                Y[r * 4 + c] = live * d_r[(long long)i * 3 + c] * 0.05f * (1.0f - t * t);
            }
        }
        __syncthreads();
        tile_weight_grad(Y, 4, 3, H2, DB_LDH, 32, rows, dW2, db2);
        tile_input_grad(Y, 4, 3, W2, 32, D, DB_LDH, false);
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 32; i += blockDim.x) {
            const int idx = (i & (BT - 1)) * DB_LDH + i / BT;
            D[idx] *= softplus_grad_from_out(H2[idx]);
        }
        __syncthreads();
        tile_weight_grad(D, DB_LDH, 32, H1, DB_LDH, 32, rows, dW1, db1);
        tile_input_grad(D, DB_LDH, 32, W1, 32, H2, DB_LDH, false);      // d h1  (H2 is free now)
        __syncthreads();
        for (int i = threadIdx.x; i < BT * 32; i += blockDim.x) {
            const int idx = (i & (BT - 1)) * DB_LDH + i / BT;
            D[idx] = H2[idx] * softplus_grad_from_out(H1[idx]);          // d z0
        }
        __syncthreads();
        tile_weight_grad(D, DB_LDH, 32, E, DB_LDE, 19, rows, dW0, db0);
        __syncthreads();
        tile_input_grad(D, DB_LDH, 32, W0, 19, E, DB_LDE, false);       // d e  (E overwritten)
        __syncthreads();
        // deformer grid: concat mode, out[3 + 2 l + f] = sum_c w_c t[c][f]  ->  d t[c][f] = w_c * d out[3 + 2 l + f]
        for (int it = threadIdx.x; it < rows * dg.n_levels; it += blockDim.x) {
            const int r = it / dg.n_levels, l = it - r * dg.n_levels, i = t0 + r;
            const float p[3] = {x0[(long long)i * 3], x0[(long long)i * 3 + 1], x0[(long long)i * 3 + 2]};
            float uvt[3], u[3];
            nvr_sample_volume(fr.tuv, p, 0, 2, uvt);
            uvt[2] = frame_dim;
            nvr_normalise(dg, uvt, u);
            LevelCoord lc;
            nvr_level_coord(u, dg.size[l], dg.res[l], lc);
            float* gtab = l < dg.start_hash ? gg.dense : gg.hash;
            const float d0 = E[r * DB_LDE + 3 + 2 * l], d1 = E[r * DB_LDE + 4 + 2 * l];
            if (gtab && (d0 != 0.0f || d1 != 0.0f)) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float w = nvr_corner_weight(lc, c);
                    float* row = gtab + nvr_corner_row(dg, l, lc, c) * 2;
                    atomicAdd(row, w * d0);
                    atomicAdd(row + 1, w * d1);
                }
            }
        }
        __syncthreads();
    }
    tile_flush(dgr.w0, dW0, 32 * 19); tile_flush(dgr.b0, db0, 32);
    tile_flush(dgr.w1, dW1, 32 * 32); tile_flush(dgr.b1, db1, 32);
    tile_flush(dgr.w2, dW2, 3 * 32); tile_flush(dgr.b2, db2, 3);
}

// d residual of every listed pair = d loss / d resd[slot][part] (offset / pair regularisers, may be null)
//                                 + d loss / d canonical xyz (from the part grid: x = x0 + resd).
// Builds the deformer work list for one part: x0 and d_r per flagged pair; pairs without any gradient are kept
// (their d_r is zero) so the list is simply the part's pair list.  blockIdx.y = part.
__global__ void k_bwd_resd_list(const int* __restrict__ counters, const PairRec* __restrict__ pairs, int cap,
                                const float* __restrict__ x0_slots, const float* __restrict__ d_resd_slots,
                                const int* __restrict__ rank_of_slot, float* __restrict__ wl_x0, float* __restrict__ wl_dr) {
    const int part = blockIdx.y;
    const int n = counters[NVR_CTR_PAIR + part];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int s = pairs[(long long)part * cap + i].surv;
        const long long src = ((long long)(rank_of_slot ? rank_of_slot[s] : s) * NVR_PARTS + part) * 3, dst = ((long long)part * cap + i) * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            wl_x0[dst + a] = x0_slots[src + a];
            wl_dr[dst + a] = d_resd_slots ? d_resd_slots[src + a] : 0.0f;
        }
    }
}
// adds the grid's d xyz of the listed (gradient-carrying) pairs into the deformer work list
__global__ void k_bwd_add_dx(const int* __restrict__ gcount_dev, const GradRec* __restrict__ gl, const float* __restrict__ d_x,
                             float* __restrict__ wl_dr) {
    const int n = *gcount_dev;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pair = gl[i].pair;
#pragma unroll
        for (int a = 0; a < 3; ++a) wl_dr[(long long)pair * 3 + a] += d_x[(long long)i * 3 + a];
    }
}

// ---------------------------------------------------------------------------------------------------
// alpha compositing on explicit raw, forward and backward                          net_utils.py:12-44
// ---------------------------------------------------------------------------------------------------
// raw (R,S,4) -> weights (R,S), rgb_map (R,3), acc_map (R).  One thread per ray (training batches are
// a few thousand rays of 64 samples).
__global__ void k_composite_fwd(const float4* __restrict__ raw, long long n_rays, int S, float* __restrict__ weights,
                                float* __restrict__ rgb_map, float* __restrict__ acc_map) {
    for (long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x; ray < n_rays; ray += (long long)gridDim.x * blockDim.x) {
        float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, ca = 0.f;
        for (int k = 0; k < S; ++k) {
            const float4 r = raw[ray * S + k];
            const float w = r.w * T;                                    // alpha_i * prod_{j<i} (1 - alpha_j), epsilon = 0
            weights[ray * S + k] = w;
            cr += w * r.x; cg += w * r.y; cb += w * r.z; ca += w;
            T *= 1.0f - r.w;
        }
        rgb_map[ray * 3] = cr; rgb_map[ray * 3 + 1] = cg; rgb_map[ray * 3 + 2] = cb;
        acc_map[ray] = ca;
    }
}
// d w_i = d_rgb_map . c_i + d_acc + d_weights_i;  d c_i = w_i d_rgb_map;
// d alpha_i = T_i (d w_i - Q_i),  Q_i = sum_{k>i} d w_k alpha_k prod_{i<j<k} (1 - alpha_j)
//           = alpha_{i+1} d w_{i+1} + (1 - alpha_{i+1}) Q_{i+1}          (no division by 1 - alpha)
__global__ void k_composite_bwd(const float4* __restrict__ raw, long long n_rays, int S, const float* __restrict__ d_weights,
                                const float* __restrict__ d_rgb_map, const float* __restrict__ d_acc_map, float4* __restrict__ d_raw) {
    for (long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x; ray < n_rays; ray += (long long)gridDim.x * blockDim.x) {
        const float gr = d_rgb_map ? d_rgb_map[ray * 3] : 0.f, gg = d_rgb_map ? d_rgb_map[ray * 3 + 1] : 0.f,
                    gb = d_rgb_map ? d_rgb_map[ray * 3 + 2] : 0.f, ga = d_acc_map ? d_acc_map[ray] : 0.f;
        float T = 1.0f;
        for (int k = 0; k < S; ++k) {                                   // forward sweep: store T_i in d_raw.w for the reverse sweep
            const float a = raw[ray * S + k].w;
            d_raw[ray * S + k].w = T;
            T *= 1.0f - a;
        }
        float Q = 0.0f;
        for (int k = S - 1; k >= 0; --k) {
            const float4 r = raw[ray * S + k];
            const float Ti = d_raw[ray * S + k].w;
            const float dw = (gr * r.x + gg * r.y + gb * r.z) + ga + (d_weights ? d_weights[ray * S + k] : 0.0f);
            const float w = r.w * Ti;
            d_raw[ray * S + k] = make_float4(w * gr, w * gg, w * gb, Ti * (dw - Q));
            Q = r.w * dw + (1.0f - r.w) * Q;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// training-time sampling along the rays                                         inb_renderer.py:15-31
// ---------------------------------------------------------------------------------------------------
// z = near (1 - t) + far t (t = linspace(0, 1, S)); with `u` (the reference's torch.rand draw, (R,S)) the stratified jitter
// z' = lower + (upper - lower) u between the midpoints; wpts = o + d z', viewdir = d.  Separate multiplies and adds like the
// reference's elementwise torch ops (no FMA contraction): z_vals feeds the distortion regulariser bit for bit.
__global__ void k_train_sample(const float* __restrict__ ray_o, const float* __restrict__ ray_d, const float* __restrict__ near_,
                               const float* __restrict__ far_, const float* __restrict__ u, long long n_rays, int S,
                               float* __restrict__ z_vals, float* __restrict__ wpts, float* __restrict__ viewdir) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rays * S; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / S;
        const int k = (int)(i - r * S);
        const float nr = near_[r], fa = far_[r];
        auto zk = [&](int kk) {
            const float t = nvr_linspace01(kk, S);
            return __fadd_rn(__fmul_rn(nr, 1.0f - t), __fmul_rn(fa, t));                        // :18
        };
        float z = zk(k);
        if (u) {                                                                               // :20-27
            const float lower = k > 0 ? __fmul_rn(0.5f, __fadd_rn(z, zk(k - 1))) : z;
            const float upper = k < S - 1 ? __fmul_rn(0.5f, __fadd_rn(zk(k + 1), z)) : z;
            z = __fadd_rn(lower, __fmul_rn(upper - lower, u[i]));
        }
        z_vals[i] = z;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float d = ray_d[r * 3 + a];
            wpts[i * 3 + a] = __fadd_rn(ray_o[r * 3 + a], __fmul_rn(d, z));                    // :29
            viewdir[i * 3 + a] = d;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// distortion regulariser                                                        inb_renderer.py:96-103
// ---------------------------------------------------------------------------------------------------
// loss_r = sum_ij w_i w_j |m_i - m_j|, m_k = (z_k + z_{k+1}) / 2 (the last interval has zero length: m_{S-1} = z_{S-1}).
// One warp per ray; backward: d w_i = 2 g_r sum_j w_j |m_i - m_j| (z carries no gradient: the sample depths are constants).
__global__ void __launch_bounds__(256)
k_distortion(const float* __restrict__ weights, const float* __restrict__ z_vals, const float* __restrict__ d_loss, long long n_rays,
             int S, float* __restrict__ loss, float* __restrict__ d_weights) {
    extern __shared__ float sm_d[];                               // [warps][2][S]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* sw = sm_d + (size_t)wid * 2 * S;
    float* smid = sw + S;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rays; r += ((long long)gridDim.x * blockDim.x) >> 5) {
        for (int k = lane; k < S; k += 32) {
            const float z = z_vals[r * S + k], zn = z_vals[r * S + (k + 1 < S ? k + 1 : k)];
            sw[k] = weights[r * S + k];
            smid[k] = (z + zn) / 2;
        }
        __syncwarp();
        float acc = 0.0f;
        const float g = d_loss ? d_loss[r] : 0.0f;
        for (int i = lane; i < S; i += 32) {
            float row = 0.0f;
            const float mi = smid[i];
            for (int j = 0; j < S; ++j) row += sw[j] * fabsf(mi - smid[j]);
            acc += sw[i] * row;
            if (d_weights) d_weights[r * S + i] = 2.0f * g * row;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (loss && lane == 0) loss[r] = acc;
        __syncwarp();
    }
}

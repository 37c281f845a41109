// nvr_warp_tc.cuh -- k_warp with the deformer MLP (uv_deformer.py:23-45: 19 -> 32 -> 32 -> 3, softplus) on tcgen05.
//
// k_warp evaluates the MLP per thread on the CUDA cores: 1 728 multiply-adds per pair fed by broadcast shared-memory loads,
// half of the kernel's time (ncu r2f: 34 % of the stall samples on the FFMA2 lines waiting for LDS, another 15 % around them).
// Here a CTA of 128 threads owns a tile of 128 pairs: every thread still does its pair's blend / LBS / tuv lookup / 8-level
// F=2 grid embedding exactly as k_warp does, but writes the 19 embedding values as fp16-split halves straight into the tile's
// X panel, and the three layers run as 3xFP16 tensor-core GEMMs (nvr_mlp_f16.cuh: hi + lo operands, fp32 accumulation in
// tensor memory, fp32-equivalent results) with the thread reading back its own row for bias + softplus between them.
//
//   shared memory (29 KB per CTA, 7 CTAs per SM): weight panels hi/lo for W0 (K 32 x N 32), W1 (32 x 32), W2 (32 x 16, three
//   real rows), biases, the frame's A / big_A, the tile's X panels hi/lo ([4 chunks][128 rows][16 B] each)
//   tensor memory (64 columns per CTA): [0,32) accumulator of layer 0, then the hidden activations as the next A operand (hi in
//   [0,16), lo in [16,32), two K elements per column); [32,64) accumulators of layers 1 and 2
#pragma once
#include "nvr_mlp_f16.cuh"

#define WT_THREADS 128
#define WT_P0 0                                            // byte offsets; a panel is [K/8][N][8 halves], hi then lo
#define WT_P1 (WT_P0 + 2 * 32 * 32 * 2)
#define WT_P2 (WT_P1 + 2 * 32 * 32 * 2)
#define WT_BIAS (WT_P2 + 2 * 32 * 16 * 2)                  // fp32: b0 32 | b1 32 | b2 4
#define WT_A (WT_BIAS + 68 * 4)                            // fp32: A 24x16 | big_A 24x16
#define WT_X (((WT_A + 2 * NVR_JOINTS * 16 * 4) + 127) & ~127)   // X_hi | X_lo, 8192 B each
#define WT_BAR (WT_X + 2 * 8192)
#define WT_SMEM_BYTES (WT_BAR + 16)
#define WT_TMEM_COLS 64

__device__ __forceinline__ void wt_put(unsigned char* sm, int off, int N, int n, int k, float w) {   // element (n, k) of a panel pair
    float h, l;
    split11(w, h, l);
    __half* hb = reinterpret_cast<__half*>(sm + off);
    const int idx = ((k >> 3) * N + n) * 8 + (k & 7);
    hb[idx] = __float2half_rn(h);
    hb[32 * N + idx] = __float2half_rn(l);                  // the lo panel follows the hi panel (K = 32 -> 32 N halves)
}

// 3xFP16 GEMM with K = 32: D = A B^T; A from shared memory (X panels) or tensor memory (packed activations), B = a panel pair
template <int N>
__device__ __forceinline__ void wt_gemm_ss(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t idesc) {
    constexpr uint32_t lbo_b = N * 16, lbo_a = 128 * 16, sbo = 128;
    const uint32_t b_lo = b_hi + 32 * N * 2;
    uint32_t acc = 0u;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const uint64_t bh = umma_desc(b_hi + ks * 2 * lbo_b, lbo_b, sbo), bl = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, sbo);
        const uint64_t ah = umma_desc(a_hi + ks * 2 * lbo_a, lbo_a, sbo), al = umma_desc(a_lo + ks * 2 * lbo_a, lbo_a, sbo);
        umma_f16_ss(d, al, bh, idesc, acc);
        umma_f16_ss(d, ah, bl, idesc, 1u);
        umma_f16_ss(d, ah, bh, idesc, 1u);
        acc = 1u;
    }
}
template <int N>
__device__ __forceinline__ void wt_gemm_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t idesc) {
    constexpr uint32_t lbo_b = N * 16, sbo = 128;
    const uint32_t b_lo = b_hi + 32 * N * 2;
    uint32_t acc = 0u;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const uint64_t bh = umma_desc(b_hi + ks * 2 * lbo_b, lbo_b, sbo), bl = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, sbo);
        umma_f16_ts(d, a_lo + ks * 8, bh, idesc, acc);
        umma_f16_ts(d, a_hi + ks * 8, bl, idesc, 1u);
        umma_f16_ts(d, a_hi + ks * 8, bh, idesc, 1u);
        acc = 1u;
    }
}

// 32 activations of one row -> hi / lo halves of the next A operand (16 + 16 packed columns)
__device__ __forceinline__ void wt_store_split(uint32_t t_hi, uint32_t t_lo, const float* a) {
    float hv[16], lv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float h0, l0, h1, l1;
        split11(a[2 * i], h0, l0);
        split11(a[2 * i + 1], h1, l1);
        hv[i] = __uint_as_float(pack_h2(h0, h1));
        lv[i] = __uint_as_float(pack_h2(l0, l1));
    }
    tmem_st16(t_hi, hv);
    tmem_st16(t_lo, lv);
}

// Same contract as k_warp (nvr_kernels.cuh): blockIdx.y = part; one tile = 128 consecutive records of the part's list.
__global__ void __launch_bounds__(WT_THREADS, 6)
k_warp_tc(FrameDev fr, GridDev dg, DeformerMlp dm, const float* __restrict__ dirs, int dir_div,
          const int* __restrict__ counters, const float4* __restrict__ surv, const KnnRec* __restrict__ recs,
          PairRec* __restrict__ pairs, int cap, float* __restrict__ dbg, float* __restrict__ out_x0, float* __restrict__ out_resd,
          const int* __restrict__ out_rank) {
    extern __shared__ __align__(128) unsigned char smw[];
    const int part = blockIdx.y;
    const int n = counters[NVR_CTR_PAIR + part];
    const int n_tiles = (n + 127) / 128;
    if ((int)blockIdx.x >= n_tiles) return;                       // block-uniform, before any allocation
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smw + WT_BAR);
    uint32_t* tbase_slot = reinterpret_cast<uint32_t*>(smw + WT_BAR + 8);
    float* sb = reinterpret_cast<float*>(smw + WT_BIAS);
    float* sA = reinterpret_cast<float*>(smw + WT_A);
    float* sBig = sA + NVR_JOINTS * 16;
    // ---- per CTA: weight panels (fp16 split), biases, the frame's joint transforms
    for (int i = tid; i < 32 * 32; i += WT_THREADS) {
        const int nn = i >> 5, k = i & 31;
        wt_put(smw, WT_P0, 32, nn, k, k < 19 ? dm.w0[nn * 19 + k] : 0.0f);
        wt_put(smw, WT_P1, 32, nn, k, dm.w1[nn * 32 + k]);
        if (nn < 16) wt_put(smw, WT_P2, 16, nn, k, nn < 3 ? dm.w2[nn * 32 + k] : 0.0f);
    }
    for (int i = tid; i < 32; i += WT_THREADS) { sb[i] = dm.b0[i]; sb[32 + i] = dm.b1[i]; }
    if (tid < 4) sb[64 + tid] = tid < 3 ? dm.b2[tid] : 0.0f;
    for (int i = tid; i < NVR_JOINTS * 16; i += WT_THREADS) { sA[i] = fr.A[i]; sBig[i] = fr.bigA[i]; }
    if (warp == 0) tmem_alloc(smem_u32(tbase_slot), WT_TMEM_COLS);
    if (tid == 0) mbar_init(smem_u32(bar), 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tbase_slot;
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar_a = smem_u32(bar);
    unsigned char* x_hi = smw + WT_X;
    unsigned char* x_lo = x_hi + 8192;
    const uint32_t s_xhi = smem_u32(x_hi), s_xlo = smem_u32(x_lo);
    const uint32_t s_p0 = smem_u32(smw + WT_P0), s_p1 = smem_u32(smw + WT_P1), s_p2 = smem_u32(smw + WT_P2);
    const uint32_t idesc32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);   // D F32, A/B F16 K-major, N 32, M 128
    const uint32_t idesc16 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
    const float frame_dim = fr.frame_dim[0];
    const float* pbw_part = fr.part_pbw + (long long)part * fr.maxlen * NVR_JOINTS;
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int i = tile * 128 + tid;
        const bool live = i < n;
        const KnnRec rec = recs[(long long)part * cap + (live ? i : n - 1)];
        const float4 sv = surv[rec.surv];
        const float p[3] = {sv.x, sv.y, sv.z};
        const int sample = __float_as_int(sv.w);
        const long long di = (long long)(sample / dir_div) * 3;
        const float wd[3] = {dirs[di], dirs[di + 1], dirs[di + 2]};
        float d[3], x0[3], v[3];
        nvr_dir_to_pose(fr.R, wd, d);
        nvr_blend_lbs(rec.idx, rec.w, pbw_part, sA, sBig, p, d, x0, v);
        // ---- deformer input: (u, v) from the tuv volume, frame_dim, 8-level F=2 grid -> 19 values, written as fp16-split
        //      halves into the tile's X panels (row = tid; columns 19..31 zero)
        {
            float uvt[3];
            nvr_sample_volume(fr.tuv, x0, 0, 2, uvt);              // pts_sample_uv :32
            uvt[2] = frame_dim;                                    // :35
            __half* xh = reinterpret_cast<__half*>(x_hi);
            __half* xl = reinterpret_cast<__half*>(x_lo);
            auto emit = [&](int k, float val) {
                float h, l;
                split11(val, h, l);
                const int idx = ((k >> 3) * 128 + tid) * 8 + (k & 7);
                xh[idx] = __float2half_rn(h);
                xl[idx] = __float2half_rn(l);
            };
            nvr_embed_point_f2_emit(dg, uvt, emit);                // :37
#pragma unroll
            for (int k = 19; k < 32; ++k) emit(k, 0.0f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        float a[32];
        // ---- layer 0
        if (tid == 0) {
            tc_fence_after();
            wt_gemm_ss<32>(tbase, s_xhi, s_xlo, s_p0, idesc32);
            umma_commit(bar_a);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        tc_fence_after();
        tmem_ld16(trow, a); tmem_ld16(trow + 16, a + 16);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = nvr_softplus_hidden(a[k] + sb[k]);
        wt_store_split(trow, trow + 16, a);
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        // ---- layer 1
        if (tid == 0) {
            tc_fence_after();
            wt_gemm_ts<32>(tbase + 32, tbase, tbase + 16, s_p1, idesc32);
            umma_commit(bar_a);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        tc_fence_after();
        tmem_ld16(trow + 32, a); tmem_ld16(trow + 48, a + 16);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = nvr_softplus_hidden(a[k] + sb[32 + k]);
        wt_store_split(trow, trow + 16, a);
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        // ---- layer 2 (three real outputs in a 16-wide tile) + 0.05 tanh
        if (tid == 0) {
            tc_fence_after();
            wt_gemm_ts<16>(tbase + 32, tbase, tbase + 16, s_p2, idesc16);
            umma_commit(bar_a);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        tc_fence_after();
        tmem_ld16(trow + 32, a);
        tmem_wait_ld();
        float r[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) r[o] = 0.05f * tanhf(a[o] + sb[64 + o]);                     // :39
        tc_fence_before();                                         // this tile's TMEM reads precede the next tile's MMAs
        if (live) {
            PairRec out;
            out.x = x0[0] + r[0]; out.y = x0[1] + r[1]; out.z = x0[2] + r[2];                    // :113
            out.vx = v[0]; out.vy = v[1]; out.vz = v[2];
            out.surv = rec.surv; out._pad = 0;
            pairs[(long long)part * cap + i] = out;
            if (out_x0) {
                const long long o3 = ((long long)(out_rank ? out_rank[rec.surv] : rec.surv) * NVR_PARTS + part) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) { out_x0[o3 + c] = x0[c]; out_resd[o3 + c] = r[c]; }
            }
            if (dbg) {
                float* dr = dbg + ((long long)sample * NVR_PARTS + part) * 8;
                dr[1] = out.x; dr[2] = out.y; dr[3] = out.z; dr[4] = v[0]; dr[5] = v[1]; dr[6] = v[2];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tbase_slot, WT_TMEM_COLS);
}

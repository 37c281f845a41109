// nvr_mlp_tc.cuh -- the part MLPs (part_base_network.py:44-63) on the 5th-generation tensor cores.
//
//   k_mlp_prep   per render, one CTA per part: repack the part's Linear weights into the tcgen05 operand
//                panels (3xTF32 split: hi = top 19 bits, lo = w - hi) and fold what can be folded
//   k_mlp_tc<H>  persistent, one CTA (two 128-pair tile slots x H epilogue warpgroups of 128 threads) per SM;
//                weight panels and the input tile in shared memory, hidden activations + accumulators in
//                TMEM, tcgen05.mma kind::tf32 issued by one thread per slot.  H = 2 (512 threads): two
//                warps share each 32-lane TMEM quarter of a slot and split the 64 activation columns
//
// Algebra.  Reference, per pair (e = grid embedding 19, pe = PosEnc(dir) 27, lat = latent row 8):
//     h    = softplus(W0 e + b0)                       64
//     o    = W1 h + b1                                 17      (no activation)
//     occ  = 1 - exp(-softplus(o[0]));  feat = o[1:17]
//     g    = softplus(W2 [e | pe | feat | lat] + b2)   64
//     g    = softplus(W3 g + b3)                       64      (body, head only)
//     rgb  = sigmoid(W4 g + b4)                        3
// `feat` is linear in h and `lat` is the same row for every pair of a frame, so
//     W2 [e|pe|feat|lat] + b2 = W2x [e|pe] + (W2f W1f) h + (b2 + W2f b1f + W2l lat) = W2x x + M h + b2'
// which removes the 17-wide GEMM and the 70-wide concatenation: three 64-wide GEMMs (K = 24, 48+64, 64)
// run on the tensor cores; the two skinny products (o[0], rgb: 4 outputs x 64) stay on the CUDA cores in
// full fp32 inside the epilogue that already holds the 64 activations in registers.
//
// Precision.  kind::tf32 keeps 10 mantissa bits per operand; per-sample outputs must match the fp32
// reference to 1e-4, so every GEMM is the 3xTF32 sum  A_hi B_hi + A_lo B_hi + A_hi B_lo  (error ~2^-21
// relative, fp32 accumulation in TMEM).  hi/lo are split explicitly (mask / subtract), not left to the
// hardware's operand truncation.
//
// Operand placement: see "shared memory and TMEM plan" below (x in shared memory, hidden activations and
// accumulators in TMEM, two tile slots per CTA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nvr_kernels.cuh"

// ---- packed per-part parameter block (floats), produced by k_mlp_prep ------------------------------
// B panels are K-major "interleaved" (no swizzle) UMMA operands: [K/4 chunks][N = 64 rows][4 floats],
// i.e. 8x(16 B) core matrices 128 B apart along N (SBO) and 64*16 B apart along K (LBO).
#define TC_N 64
#define TC_K0 24                   // e, padded (columns 19..23 of x hold pe values; their weights are zero)
#define TC_KX 48                   // [e | pe | 0 0]
#define TC_KH 64
#define TC_PANEL(K) ((K) * TC_N)
#define TC_OFF_P0 0
#define TC_OFF_PX (TC_OFF_P0 + 2 * TC_PANEL(TC_K0))
#define TC_OFF_PM (TC_OFF_PX + 2 * TC_PANEL(TC_KX))
#define TC_OFF_P3 (TC_OFF_PM + 2 * TC_PANEL(TC_KH))
#define TC_OFF_B0 (TC_OFF_P3 + 2 * TC_PANEL(TC_KH))   // b0 64
#define TC_OFF_B2 (TC_OFF_B0 + 64)                     // b2' 64
#define TC_OFF_B3 (TC_OFF_B2 + 64)                     // b3 64
#define TC_OFF_W1 (TC_OFF_B3 + 64)                     // W1[0,:] 64
#define TC_OFF_W4 (TC_OFF_W1 + 64)                     // W4 3x64
#define TC_OFF_SC (TC_OFF_W4 + 192)                    // b1[0], b4[0..2]
#define TC_BLOCK_FLOATS (TC_OFF_SC + 4)
#define TC_SMEM_BYTES (TC_BLOCK_FLOATS * 4 + 64)       // + mbarrier, tmem base

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// element (n, k) of a [K/4][64][4] panel
__device__ __forceinline__ int panel_idx(int n, int k) { return ((k >> 2) * TC_N + n) * 4 + (k & 3); }

__global__ void __launch_bounds__(256)
k_mlp_prep(const PartMlpDev* __restrict__ parts, const long long* __restrict__ latent_index, float* __restrict__ blocks) {
    const PartMlpDev pm = parts[blockIdx.x];
    float* blk = blocks + (size_t)blockIdx.x * TC_BLOCK_FLOATS;
    const float* W0 = pm.occ[0].w;  const float* W1 = pm.occ[1].w;    // (64,19), (17,64)
    const float* W2 = pm.rgb[0].w;                                      // (64,70)
    const bool three = pm.n_rgb == 3;
    const float* W3 = pm.rgb[1].w;                                      // (64,64) when three
    const float* W4 = pm.rgb[pm.n_rgb - 1].w;                           // (3,64)
    long long li = latent_index[0];
    li = li < 0 ? 0 : (li >= pm.n_latent ? pm.n_latent - 1 : li);
    const float* lat = pm.latent + li * 8;
    auto put = [&](int off, int K, int n, int k, float w) {
        const float h = tf32_hi(w);
        blk[off + panel_idx(n, k)] = h;
        blk[off + TC_PANEL(K) + panel_idx(n, k)] = w - h;
    };
    for (int i = threadIdx.x; i < 64 * TC_K0; i += blockDim.x) {
        const int n = i / TC_K0, k = i - n * TC_K0;
        put(TC_OFF_P0, TC_K0, n, k, k < 19 ? W0[n * 19 + k] : 0.0f);
    }
    for (int i = threadIdx.x; i < 64 * TC_KX; i += blockDim.x) {
        const int n = i / TC_KX, k = i - n * TC_KX;
        put(TC_OFF_PX, TC_KX, n, k, k < 46 ? W2[n * 70 + k] : 0.0f);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i >> 6, k = i & 63;
        float m = 0.0f;                                                 // M = W2f W1f
        for (int j = 0; j < 16; ++j) m += W2[n * 70 + 46 + j] * W1[(1 + j) * 64 + k];
        put(TC_OFF_PM, TC_KH, n, k, m);
        put(TC_OFF_P3, TC_KH, n, k, three ? W3[n * 64 + k] : 0.0f);
    }
    for (int n = threadIdx.x; n < 64; n += blockDim.x) {
        blk[TC_OFF_B0 + n] = pm.occ[0].b[n];
        float b = pm.rgb[0].b[n];
        for (int j = 0; j < 16; ++j) b += W2[n * 70 + 46 + j] * pm.occ[1].b[1 + j];
        for (int c = 0; c < 8; ++c) b += W2[n * 70 + 62 + c] * lat[c];
        blk[TC_OFF_B2 + n] = b;
        blk[TC_OFF_B3 + n] = three ? pm.rgb[1].b[n] : 0.0f;
        blk[TC_OFF_W1 + n] = W1[n];
        for (int j = 0; j < 3; ++j) blk[TC_OFF_W4 + j * 64 + n] = W4[j * 64 + n];
    }
    if (threadIdx.x == 0) {
        blk[TC_OFF_SC] = pm.occ[1].b[0];
        for (int j = 0; j < 3; ++j) blk[TC_OFF_SC + 1 + j] = pm.rgb[pm.n_rgb - 1].b[j];
    }
}

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
// D[tmem] (+)= A[tmem] . B[smem]^T, M = 128, N = 64, K = 8 (tf32)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, no swizzle: start address, LBO (K-chunk stride), SBO (8-row group stride), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                 "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
                 "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                 "r"(__float_as_uint(v[15])) : "memory");
}
// write v[0..16) as the hi and lo halves of a 3xTF32 A operand
__device__ __forceinline__ void tmem_st16_split(uint32_t t_hi, uint32_t t_lo, const float* v) {
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { h[i] = tf32_hi(v[i]); l[i] = v[i] - h[i]; }
    tmem_st16(t_hi, h);
    tmem_st16(t_lo, l);
}

// ---- shared memory and TMEM plan ------------------------------------------------------------------
// A CTA runs TWO tile slots, each with its own 128-pair tile in flight, so one slot's tensor-core GEMMs
// overlap the other slot's activation epilogue.  Warp w works for slot (w >> 2) & 1 on TMEM lanes
// 32 (w & 3) .. +31 (the hardware's lane window of a warp) and, with H = 2 warpgroups per slot, on activation
// columns 32 (w >> 3) .. +31: the epilogue (tcgen05.ld -> bias + softplus -> hi/lo split -> tcgen05.st) is
// latency-bound at 2 warps per scheduler (ncu r1f: issue 37 %, tensor pipe 26 %), so the second warpgroup
// per slot nearly doubles what the tensor pipe is fed.  The partial dot products of the two skinny layers
// (o[0], rgb) cross between the halves through 4 floats per row of shared memory.  The slots share the weight
// panels; each owns an X operand buffer in shared memory and 192 TMEM columns:
//     smem  X_hi / X_lo   [48/4 chunks][128 rows][4] floats each (the same K-major interleaved layout as B)
//     TMEM  [0,64) H_hi  [64,128) H_lo  (hidden activations, A operand of the next GEMM)  [128,192) D
// GEMM 0 accumulates into the H_hi columns (H is dead then); GEMM 3 into D (consumed by then).
#define TC_THREADS(H) (256 * (H))
#define TC_XPANEL (TC_KX * 128)                                  // floats per X panel
#define TC_SM_X (TC_BLOCK_FLOATS)                                // [slot][hi|lo][TC_XPANEL]
#define TC_SM_BAR (TC_SM_X + 4 * TC_XPANEL)                      // 2 mbarriers (4 floats), tmem base (1)
#define TC_SM_EX (TC_SM_BAR + 8)                                 // [slot][128 rows][4]: o0 / rgb partials of the 2nd half
#undef TC_SMEM_BYTES
#define TC_SMEM_BYTES ((TC_SM_EX + 2 * 128 * 4) * 4)
#define TC_SLOT_COLS 192
#define TC_COL_HHI 0
#define TC_COL_HLO 64
#define TC_COL_D 128
#define TC_TMEM_COLS 512

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// 3xTF32 GEMM: D (+)= A[:, 0:K] B^T, small terms first.  A from TMEM (TS) ...
__device__ __forceinline__ void gemm3_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_smem, int K, uint32_t idesc, bool first) {
    const uint32_t lbo = TC_N * 16, sbo = 128;
    const uint32_t b_lo_smem = b_hi_smem + (uint32_t)TC_PANEL(K) * 4;
    uint32_t acc = first ? 0u : 1u;
#pragma unroll 1
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t boff = (uint32_t)ks * 2 * lbo;
        const uint64_t bh = umma_desc(b_hi_smem + boff, lbo, sbo), bl = umma_desc(b_lo_smem + boff, lbo, sbo);
        umma_tf32_ts(d, a_lo + ks * 8, bh, idesc, acc);
        umma_tf32_ts(d, a_hi + ks * 8, bl, idesc, 1u);
        umma_tf32_ts(d, a_hi + ks * 8, bh, idesc, 1u);
        acc = 1u;
    }
}
// ... or from shared memory (SS): a_*_smem are [K/4][128][4] panels
__device__ __forceinline__ void gemm3_ss(uint32_t d, uint32_t a_hi_smem, uint32_t a_lo_smem, uint32_t b_hi_smem, int K, int KB,
                                         uint32_t idesc, bool first) {
    const uint32_t lbo_b = TC_N * 16, lbo_a = 128 * 16, sbo = 128;
    const uint32_t b_lo_smem = b_hi_smem + (uint32_t)TC_PANEL(KB) * 4;
    uint32_t acc = first ? 0u : 1u;
#pragma unroll 1
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t boff = (uint32_t)ks * 2 * lbo_b, aoff = (uint32_t)ks * 2 * lbo_a;
        const uint64_t bh = umma_desc(b_hi_smem + boff, lbo_b, sbo), bl = umma_desc(b_lo_smem + boff, lbo_b, sbo);
        const uint64_t ah = umma_desc(a_hi_smem + aoff, lbo_a, sbo), al = umma_desc(a_lo_smem + aoff, lbo_a, sbo);
        umma_tf32_ss(d, al, bh, idesc, acc);
        umma_tf32_ss(d, ah, bl, idesc, 1u);
        umma_tf32_ss(d, ah, bh, idesc, 1u);
        acc = 1u;
    }
}

// PosEnc (freq_embedder.py:20-31) with one sincosf per axis and double-angle steps for 2v, 4v, 8v
// (absolute error < 1e-6, against 24 separate sinf / cosf calls).
// sin and cos of a moderate argument without sincosf's slow path (its Payne-Hanek branch costs a stack frame and local
// memory traffic even when never taken; ncu r2a: 10 % of k_mlp_tc's stall samples): three-term Cody-Waite reduction by pi/2
// and the classic degree-7 / degree-8 minimax kernels on [-pi/4, pi/4] (error < 2 ulp for |x| < 1e4; canonical view
// directions are O(1)).  Larger or non-finite arguments take the library routine.
__device__ __forceinline__ void sincos_small(float x, float* s, float* c) {
    if (!(fabsf(x) < 1.0e4f)) { sincosf(x, s, c); return; }
    const float k = rintf(x * 0.6366197723675814f);
    float r = fmaf(-k, 1.5707962512969971f, x);
    r = fmaf(-k, 7.5497894158615964e-08f, r);
    r = fmaf(-k, 5.3903029534742384e-15f, r);
    const float r2 = r * r;
    const float sr = fmaf(r * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
    const float cr = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                          fmaf(r2, -0.5f, 1.0f));
    const int q = (int)k & 3;
    const float a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    *s = (q & 2) ? -a : a;
    *c = ((q + 1) & 2) ? -b : b;
}
__device__ __forceinline__ void posenc27_doubling(const float v[3], float* out) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        out[a] = v[a];
        float s, c;
        sincos_small(v[a], &s, &c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            out[3 + k * 6 + a] = s;
            out[3 + k * 6 + 3 + a] = c;
            const float s2 = 2.0f * s * c, c2 = 1.0f - 2.0f * s * s;
            s = s2; c = c2;
        }
    }
}
template <int H>
__device__ __forceinline__ void slot_sync(int slot) { asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(128 * H) : "memory"); }

// One launch per part.  blk = that part's packed block; pl / el = its pair list and embedding rows.
template <int H>
__global__ void __launch_bounds__(TC_THREADS(H), 1)
k_mlp_tc(const float* __restrict__ blk, int n_rgb, int part, const int* __restrict__ count_dev, const PairRec* __restrict__ pl,
         const float* __restrict__ el, float4* __restrict__ raws, int out_stride) {
    extern __shared__ __align__(128) float sm[];
    const int n = *count_dev;
    const int n_tiles = (n + 127) / 128;
    if ((int)blockIdx.x * 2 >= n_tiles) return;                     // block-uniform, before any allocation
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TC_SM_BAR);
    uint32_t* tbase_slot = reinterpret_cast<uint32_t*>(sm + TC_SM_BAR + 4);
    const int tid = threadIdx.x, warp = tid >> 5, slot = (warp >> 2) & 1, half = warp >> 3, stid = tid & 127;
    constexpr int CPH = 4 / H;                                      // 16-column chunks per half
    const bool lead = half == 0;
    float* ex = sm + TC_SM_EX + (slot * 128 + stid) * 4;
    {   // weights -> shared memory (float4 copies; the block is 16-byte aligned by construction)
        const float4* src = reinterpret_cast<const float4*>(blk);
        float4* dst = reinterpret_cast<float4*>(sm);
        for (int i = tid; i < TC_BLOCK_FLOATS / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (warp == 0) tmem_alloc(smem_u32(tbase_slot), TC_TMEM_COLS);
    if (tid == 0) { mbar_init(smem_u32(bars), 1); mbar_init(smem_u32(bars + 1), 1); }
    // generic-proxy writes of the weight panels must be visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tbase_slot + (uint32_t)slot * TC_SLOT_COLS;
    const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 TMEM lanes
    const uint32_t bar_a = smem_u32(bars + slot);
    float* x_hi = sm + TC_SM_X + slot * 2 * TC_XPANEL;
    float* x_lo = x_hi + TC_XPANEL;
    const uint32_t s_xhi = smem_u32(x_hi), s_xlo = smem_u32(x_lo);
    const uint32_t s_p0 = smem_u32(sm + TC_OFF_P0), s_px = smem_u32(sm + TC_OFF_PX), s_pm = smem_u32(sm + TC_OFF_PM),
                   s_p3 = smem_u32(sm + TC_OFF_P3);
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((128u >> 4) << 24);
    const float* b0 = sm + TC_OFF_B0; const float* b2 = sm + TC_OFF_B2; const float* b3 = sm + TC_OFF_B3;
    const float* w1 = sm + TC_OFF_W1; const float* w4 = sm + TC_OFF_W4; const float* sc = sm + TC_OFF_SC;
    const bool three = n_rgb == 3;
    uint32_t phase = 0;

    for (int tile = blockIdx.x * 2 + slot; tile < n_tiles; tile += gridDim.x * 2) {
        const int row = tile * 128 + stid;
        const int pr = min(row, n - 1);
        int surv;
        // ---- x = [e 19 | pe 27 | 0 0] -> X_hi / X_lo panels (row = stid)
        {
            float x[48];
            const float4* e4 = reinterpret_cast<const float4*>(el + (size_t)pr * NVR_EMB_STRIDE);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float4 t = __ldg(e4 + q);
                x[q * 4] = t.x; x[q * 4 + 1] = t.y; x[q * 4 + 2] = t.z; x[q * 4 + 3] = t.w;   // x[19] is overwritten below
            }
            const PairRec rec = pl[pr];
            surv = rec.surv;
            const float v[3] = {rec.vx, rec.vy, rec.vz};
            posenc27_doubling(v, x + 19);                           // part_base_network.py:54
            x[46] = 0.0f; x[47] = 0.0f;
#pragma unroll
            for (int c = 0; c < TC_KX / 4; ++c) {
                if (H == 2 && (c & 1) != half) continue;            // the halves interleave the 16-byte chunks
                float4 h4, l4;
                h4.x = tf32_hi(x[c * 4]); h4.y = tf32_hi(x[c * 4 + 1]); h4.z = tf32_hi(x[c * 4 + 2]); h4.w = tf32_hi(x[c * 4 + 3]);
                l4.x = x[c * 4] - h4.x; l4.y = x[c * 4 + 1] - h4.y; l4.z = x[c * 4 + 2] - h4.z; l4.w = x[c * 4 + 3] - h4.w;
                reinterpret_cast<float4*>(x_hi)[c * 128 + stid] = h4;
                reinterpret_cast<float4*>(x_lo)[c * 128 + stid] = l4;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();                                           // also orders the previous tile's TMEM reads
        slot_sync<H>(slot);
        // ---- GEMM 0: h_pre = W0 e     (accumulator in the H_hi columns)
        if (stid == 0 && lead) {
            tc_fence_after();
            gemm3_ss(tbase + TC_COL_HHI, s_xhi, s_xlo, s_p0, TC_K0, TC_K0, idesc, true);
            umma_commit(bar_a);
            // the x half of GEMM 2 needs nothing from the first epilogue: queue it now (D is free: the previous
            // tile's last reads were fenced before the slot barrier above) so it runs under that epilogue;
            // it is covered by GEMM 2's commit below, not by the one just issued
            gemm3_ss(tbase + TC_COL_D, s_xhi, s_xlo, s_px, TC_KX, TC_KX, idesc, true);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        tc_fence_after();
        float o0 = lead ? sc[0] : 0.0f;
#pragma unroll 1
        for (int c = half * CPH; c < half * CPH + CPH; ++c) {
            float h[16];
            tmem_ld16(trow + TC_COL_HHI + c * 16, h);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                h[k] = nvr_softplus_hidden(h[k] + b0[c * 16 + k]);        // MLP.forward :20-22
                o0 += w1[c * 16 + k] * h[k];
            }
            tmem_st16_split(trow + TC_COL_HHI + c * 16, trow + TC_COL_HLO + c * 16, h);
        }
        if (H == 2 && !lead) ex[0] = o0;
        tmem_wait_st();
        tc_fence_before();
        slot_sync<H>(slot);
        if (H == 2 && lead) o0 += ex[0];
        const float occ = 1.0f - expf(-nvr_softplus(o0));           // :51 (used by the lead half only)
        // ---- GEMM 2: g_pre = W2x x (queued with GEMM 0) + M h
        if (stid == 0 && lead) {
            tc_fence_after();
            gemm3_ts(tbase + TC_COL_D, tbase + TC_COL_HHI, tbase + TC_COL_HLO, s_pm, TC_KH, idesc, false);
            umma_commit(bar_a);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        tc_fence_after();
        float r[3] = {lead ? sc[1] : 0.0f, lead ? sc[2] : 0.0f, lead ? sc[3] : 0.0f};
        if (three) {                                                // block-uniform
#pragma unroll 1
            for (int c = half * CPH; c < half * CPH + CPH; ++c) {
                float g[16];
                tmem_ld16(trow + TC_COL_D + c * 16, g);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 16; ++k) g[k] = nvr_softplus_hidden(g[k] + b2[c * 16 + k]);
                tmem_st16_split(trow + TC_COL_HHI + c * 16, trow + TC_COL_HLO + c * 16, g);
            }
            tmem_wait_st();
            tc_fence_before();
            slot_sync<H>(slot);
            // ---- GEMM 3: W3 g   (accumulator back in D: GEMM 2's result has been consumed)
            if (stid == 0 && lead) {
                tc_fence_after();
                gemm3_ts(tbase + TC_COL_D, tbase + TC_COL_HHI, tbase + TC_COL_HLO, s_p3, TC_KH, idesc, true);
                umma_commit(bar_a);
            }
            mbar_wait(bar_a, phase); phase ^= 1;
            tc_fence_after();
        }
        const float* bl = three ? b3 : b2;
        // ---- last hidden activation + rgb = sigmoid(W4 g + b4)   (:58), raw = [rgb, occ] (:60)
#pragma unroll 1
        for (int c = half * CPH; c < half * CPH + CPH; ++c) {
            float g[16];
            tmem_ld16(trow + TC_COL_D + c * 16, g);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float a = nvr_softplus_hidden(g[k] + bl[c * 16 + k]);
                r[0] += w4[c * 16 + k] * a; r[1] += w4[64 + c * 16 + k] * a; r[2] += w4[128 + c * 16 + k] * a;
            }
        }
        if (H == 2) {                                               // second half hands its partial rgb sums over
            if (!lead) { ex[1] = r[0]; ex[2] = r[1]; ex[3] = r[2]; }
            tc_fence_before();                                      // its TMEM reads precede the next tile's MMAs too
            slot_sync<H>(slot);
            if (lead) { r[0] += ex[1]; r[1] += ex[2]; r[2] += ex[3]; }
        }
        if (lead && row < n)
            raws[(long long)surv * out_stride + part] = make_float4(nvr_sigmoid(r[0]), nvr_sigmoid(r[1]), nvr_sigmoid(r[2]), occ);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tbase_slot, TC_TMEM_COLS);
}

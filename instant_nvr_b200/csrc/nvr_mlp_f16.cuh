// nvr_mlp_f16.cuh -- the part MLPs (part_base_network.py:44-63) on tcgen05 with fp16-split operands, FOUR tile slots per SM.
//
// Same algebra as nvr_mlp_tc.cuh (three 64-wide GEMMs per pair: K = 19, 46 + 64, 64; the two skinny products on the CUDA
// cores inside the epilogue), different operand format:
//
//   3xFP16.  Every operand x is split x = hi + lo with hi = x rounded to an 11-bit significand and lo = x - hi (exact in fp32),
//   both stored as fp16; a GEMM is  A_lo B_hi + A_hi B_lo + A_hi B_hi  with fp32 accumulation in TMEM.  fp16 x fp16 products
//   are exact in fp32, so the result carries ~22 significand bits like 3xTF32 does (measured against fp64: the two agree to
//   4e-8 on the network's outputs), as long as the values sit in fp16's range: |x| < 65504, and a lo part below 6e-5 keeps
//   only 2^-24 absolute -- 3e-8 of a pre-activation term, far below the 1e-4 budget.  Activations (softplus outputs),
//   weights (|w| < 1), grid embeddings and the positional encoding are all O(1).
//
//   Why: kind::f16 runs at twice the rate of kind::tf32 (K = 16 per 32-cycle instruction instead of 8), and a packed fp16
//   A operand takes HALF the tensor-memory columns -- a tile slot shrinks from 192 to 128 columns, so FOUR 128-pair tiles are
//   in flight per SM instead of two.  ncu (profiles/r2c) showed the two-slot kernel idle a quarter of the time waiting for
//   its own GEMMs and MUFU-queue-bound the rest: the chain GEMM -> tcgen05.ld -> softplus -> tcgen05.st -> GEMM of one slot is
//   serial, and only more independent chains fill the tensor pipe, the MUFU pipe and the issue slots at the same time.
//
// Shared memory: weight panels hi/lo (53 KB) + per slot one X panel pair (24 KB).  Tensor memory per slot (128 columns):
//     [0,64)   GEMM 0 accumulator (fp32), then the hidden activations as the next A operand: hi in [0,32), lo in [32,64),
//              two consecutive K elements per 32-bit column
//     [64,128) accumulator of GEMM 2 (x part queued together with GEMM 0) and GEMM 3
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nvr_mlp_tc.cuh"

#define F16_K0 32                   // e (19) padded to a multiple of the MMA K (16); the pad columns' weights are zero
#define F16_KX 48                   // [e 19 | pe 27 | 0 0]
#define F16_KH 64
#define F16_PANEL_HALVES(K) ((K) * 64)                               // one hi OR lo panel: [K/8 chunks][64 rows][8 halves]
#define F16_OFF_P0 0                                                 // offsets in HALVES
#define F16_OFF_PX (F16_OFF_P0 + 2 * F16_PANEL_HALVES(F16_K0))
#define F16_OFF_PM (F16_OFF_PX + 2 * F16_PANEL_HALVES(F16_KX))
#define F16_OFF_P3 (F16_OFF_PM + 2 * F16_PANEL_HALVES(F16_KH))
#define F16_W_HALVES (F16_OFF_P3 + 2 * F16_PANEL_HALVES(F16_KH))      // 26624 halves = 53248 B
#define F16_OFF_F32 (F16_W_HALVES / 2)                               // fp32 tail, in FLOATS from the block start
#define F16_F_B0 (F16_OFF_F32)                                       // b0 64
#define F16_F_B2 (F16_F_B0 + 64)                                     // b2' 64
#define F16_F_B3 (F16_F_B2 + 64)                                     // b3 64
#define F16_F_W1 (F16_F_B3 + 64)                                     // W1[0,:] 64
#define F16_F_W4 (F16_F_W1 + 64)                                     // W4 3x64
#define F16_F_SC (F16_F_W4 + 192)                                    // b1[0], b4[0..2]
#define F16_BLOCK_FLOATS (F16_F_SC + 4)                              // 13764 floats = 55056 B
#define F16_SLOTS 4
#define F16_XPANEL_BYTES (F16_KX * 128 * 2)                          // one hi OR lo X panel: [6 chunks][128 rows][16 B]
#define F16_SM_X (F16_BLOCK_FLOATS * 4)                              // byte offset of the X panels: [slot][hi|lo]
#define F16_STAGE_E_BYTES (128 * NVR_EMB_STRIDE * 4)                 // one tile's embedding rows (128 x 80 B, contiguous in HBM)
#define F16_STAGE_BYTES (F16_STAGE_E_BYTES + 128 * 32)               // + its pair records (128 x 32 B)
#define F16_SM_STAGE (F16_SM_X + F16_SLOTS * 2 * F16_XPANEL_BYTES)   // [slot] raw input rows of the slot's NEXT tile (TMA bulk copies)
#define F16_SM_BAR (F16_SM_STAGE + F16_SLOTS * F16_STAGE_BYTES)      // mbarriers: 4 MMA + 4 stage-full + 1 parameter block; tmem base
#define F16_SMEM_BYTES (F16_SM_BAR + 96)
#define F16_SLOT_COLS 128
#define F16_COL_H 0
#define F16_COL_D 64
#define F16_THREADS (128 * F16_SLOTS)

// fp32 -> (hi, lo): hi keeps the top 11 significand bits (a value fp16 holds exactly inside its normal range), lo the rest
__device__ __forceinline__ void split11(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {      // a -> bits [0,16) (the lower K index), b -> [16,32)
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// grid (5 parts, F16_PREP_SPLIT): blockIdx.y takes every F16_PREP_SPLIT-th output row n, so the packing (the 64 x 64 x 16
// product M = W2f W1f above all) is spread over 80 CTAs instead of 5 -- it sits on the critical path of every pass
#define F16_PREP_SPLIT 16
__global__ void __launch_bounds__(256)
k_mlp_prep16(const PartMlpDev* __restrict__ parts, const long long* __restrict__ latent_index, float* __restrict__ blocks) {
    const PartMlpDev pm = parts[blockIdx.x];
    float* blk = blocks + (size_t)blockIdx.x * F16_BLOCK_FLOATS;
    __half* hb = reinterpret_cast<__half*>(blk);
    const float* W0 = pm.occ[0].w;  const float* W1 = pm.occ[1].w;    // (64,19), (17,64)
    const float* W2 = pm.rgb[0].w;                                      // (64,70)
    const bool three = pm.n_rgb == 3;
    const float* W3 = pm.rgb[1].w;                                      // (64,64) when three
    const float* W4 = pm.rgb[pm.n_rgb - 1].w;                           // (3,64)
    long long li = latent_index[0];
    li = li < 0 ? 0 : (li >= pm.n_latent ? pm.n_latent - 1 : li);
    const float* lat = pm.latent + li * 8;
    const int ny = gridDim.y, y = blockIdx.y, rows = (64 - y + ny - 1) / ny;   // this CTA's rows: n = y + j ny, j < rows
    auto put = [&](int off, int K, int n, int k, float w) {              // element (n, k) of a [K/8][64][8] panel pair
        float h, l;
        split11(w, h, l);
        const int idx = ((k >> 3) * 64 + n) * 8 + (k & 7);
        hb[off + idx] = __float2half_rn(h);
        hb[off + F16_PANEL_HALVES(K) + idx] = __float2half_rn(l);
    };
    for (int i = threadIdx.x; i < rows * F16_K0; i += blockDim.x) {
        const int n = y + (i / F16_K0) * ny, k = i % F16_K0;
        put(F16_OFF_P0, F16_K0, n, k, k < 19 ? W0[n * 19 + k] : 0.0f);
    }
    for (int i = threadIdx.x; i < rows * F16_KX; i += blockDim.x) {
        const int n = y + (i / F16_KX) * ny, k = i % F16_KX;
        put(F16_OFF_PX, F16_KX, n, k, k < 46 ? W2[n * 70 + k] : 0.0f);
    }
    for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
        const int n = y + (i >> 6) * ny, k = i & 63;
        float m = 0.0f;                                                 // M = W2f W1f
        for (int j = 0; j < 16; ++j) m += W2[n * 70 + 46 + j] * W1[(1 + j) * 64 + k];
        put(F16_OFF_PM, F16_KH, n, k, m);
        put(F16_OFF_P3, F16_KH, n, k, three ? W3[n * 64 + k] : 0.0f);
    }
    for (int j = threadIdx.x; j < rows; j += blockDim.x) {
        const int n = y + j * ny;
        blk[F16_F_B0 + n] = pm.occ[0].b[n];
        float b = pm.rgb[0].b[n];
        for (int q = 0; q < 16; ++q) b += W2[n * 70 + 46 + q] * pm.occ[1].b[1 + q];
        for (int c = 0; c < 8; ++c) b += W2[n * 70 + 62 + c] * lat[c];
        blk[F16_F_B2 + n] = b;
        blk[F16_F_B3 + n] = three ? pm.rgb[1].b[n] : 0.0f;
        blk[F16_F_W1 + n] = W1[n];
        for (int q = 0; q < 3; ++q) blk[F16_F_W4 + q * 64 + n] = W4[q * 64 + n];
    }
    if (threadIdx.x == 0 && y == 0) {
        blk[F16_F_SC] = pm.occ[1].b[0];
        for (int j = 0; j < 3; ++j) blk[F16_F_SC + 1 + j] = pm.rgb[pm.n_rgb - 1].b[j];
    }
}

__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// 3xFP16 GEMM over K (a multiple of 16): D (+)= A[:, 0:K] B^T, small terms first.  Panels are [K/8][rows][8 halves]: the byte
// geometry of the tf32 panels (16-byte chunks, 128 B between 8-row groups), one MMA (K = 16) spans two chunks.
// A from shared memory ([K/8][128][8] panels) ...
template <int K, int KB>
__device__ __forceinline__ void gemm3h_ss(uint32_t d, uint32_t a_hi_smem, uint32_t a_lo_smem, uint32_t b_hi_smem, uint32_t idesc, bool first) {
    constexpr uint32_t lbo_b = 64 * 16, lbo_a = 128 * 16, sbo = 128;
    const uint32_t b_lo_smem = b_hi_smem + (uint32_t)F16_PANEL_HALVES(KB) * 2;
    uint32_t acc = first ? 0u : 1u;
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t boff = (uint32_t)ks * 2 * lbo_b, aoff = (uint32_t)ks * 2 * lbo_a;
        const uint64_t bh = umma_desc(b_hi_smem + boff, lbo_b, sbo), bl = umma_desc(b_lo_smem + boff, lbo_b, sbo);
        const uint64_t ah = umma_desc(a_hi_smem + aoff, lbo_a, sbo), al = umma_desc(a_lo_smem + aoff, lbo_a, sbo);
        umma_f16_ss(d, al, bh, idesc, acc);
        umma_f16_ss(d, ah, bl, idesc, 1u);
        umma_f16_ss(d, ah, bh, idesc, 1u);
        acc = 1u;
    }
}
// ... or from tensor memory: K elements packed two per column, so one MMA advances 8 columns
template <int K>
__device__ __forceinline__ void gemm3h_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_smem, uint32_t idesc, bool first) {
    constexpr uint32_t lbo = 64 * 16, sbo = 128;
    const uint32_t b_lo_smem = b_hi_smem + (uint32_t)F16_PANEL_HALVES(K) * 2;
    uint32_t acc = first ? 0u : 1u;
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t boff = (uint32_t)ks * 2 * lbo;
        const uint64_t bh = umma_desc(b_hi_smem + boff, lbo, sbo), bl = umma_desc(b_lo_smem + boff, lbo, sbo);
        umma_f16_ts(d, a_lo + ks * 8, bh, idesc, acc);
        umma_f16_ts(d, a_hi + ks * 8, bl, idesc, 1u);
        umma_f16_ts(d, a_hi + ks * 8, bh, idesc, 1u);
        acc = 1u;
    }
}

// TMA bulk copy (cp.async.bulk, SASS UBLKCP): `bytes` (a multiple of 16) from global to shared memory, completion counted on an
// mbarrier in transaction bytes.  One thread issues it; nobody's registers or LSU slots are tied up while the data is in flight.
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void slot_sync128(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory"); }

// 64 activations of one row -> the hi / lo halves of the next A operand in tensor memory (32 + 32 packed columns)
__device__ __forceinline__ void store_split_h(uint32_t t_hi, uint32_t t_lo, const float* a) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float hv[16], lv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float h0, l0, h1, l1;
            split11(a[q * 32 + 2 * i], h0, l0);
            split11(a[q * 32 + 2 * i + 1], h1, l1);
            hv[i] = __uint_as_float(pack_h2(h0, h1));
            lv[i] = __uint_as_float(pack_h2(l0, l1));
        }
        tmem_st16(t_hi + q * 16, hv);
        tmem_st16(t_lo + q * 16, lv);
    }
}

// Work of one launch: the pair lists of n_parts parts (one part: the stand-alone entry point; five: a render pass).  All tiles
// of all parts form ONE index space dealt round-robin to the (CTA, slot) pairs, so a launch is balanced whatever the parts'
// sizes, and a CTA walks its tiles part by part, reloading the 54 KB parameter block at each part boundary.
struct MlpBatch {
    const float* blk[NVR_PARTS];          // packed parameter block of each part (k_mlp_prep16)
    const int* count[NVR_PARTS];          // device-side list lengths
    const PairRec* pl[NVR_PARTS];
    const float* el[NVR_PARTS];
    int n_rgb[NVR_PARTS];
    int out_part[NVR_PARTS];              // column of `raws` the part's results go to
    int n_parts;
};

__global__ void __launch_bounds__(F16_THREADS, 1)
k_mlp_f16(const MlpBatch mb_param, float4* __restrict__ raws, int out_stride) {
    extern __shared__ __align__(128) unsigned char smb[];
    __shared__ MlpBatch mb;                                           // indexed by a run-time part below: a copy in shared
    if (threadIdx.x == 0) mb = mb_param;                              // memory, not a stack copy of the parameter
    __syncthreads();
    float* smf = reinterpret_cast<float*>(smb);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smb + F16_SM_BAR);
    uint32_t* tbase_slot = reinterpret_cast<uint32_t*>(smb + F16_SM_BAR + 80);
    const int tid = threadIdx.x, warp = tid >> 5, slot = warp >> 2, stid = tid & 127;
    int tiles_before = 0, total_tiles = 0;
#pragma unroll
    for (int p = 0; p < NVR_PARTS; ++p)
        if (p < mb.n_parts) total_tiles += (*mb.count[p] + 127) / 128;
    if ((int)blockIdx.x * F16_SLOTS >= total_tiles) return;           // block-uniform, before any allocation
    if (warp == 0) tmem_alloc(smem_u32(tbase_slot), 512);
    if (tid < 2 * F16_SLOTS + 1) mbar_init(smem_u32(bars + tid), 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tbase_slot + (uint32_t)slot * F16_SLOT_COLS;
    const uint32_t bar_full = smem_u32(bars + F16_SLOTS + slot), bar_w = smem_u32(bars + 2 * F16_SLOTS);
    unsigned char* stage = smb + F16_SM_STAGE + slot * F16_STAGE_BYTES;
    const uint32_t s_stage = smem_u32(stage);
    uint32_t phase_full = 0, phase_w = 0;
    const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 TMEM lanes
    const uint32_t bar_a = smem_u32(bars + slot);
    unsigned char* x_hi = smb + F16_SM_X + slot * 2 * F16_XPANEL_BYTES;
    unsigned char* x_lo = x_hi + F16_XPANEL_BYTES;
    const uint32_t s_xhi = smem_u32(x_hi), s_xlo = smem_u32(x_lo);
    const uint32_t s_w = smem_u32(smb);
    const uint32_t s_p0 = s_w + F16_OFF_P0 * 2, s_px = s_w + F16_OFF_PX * 2, s_pm = s_w + F16_OFF_PM * 2, s_p3 = s_w + F16_OFF_P3 * 2;
    // instruction descriptor: D = F32, A = B = F16, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    const float* b0 = smf + F16_F_B0; const float* b2 = smf + F16_F_B2; const float* b3 = smf + F16_F_B3;
    const float* w1 = smf + F16_F_W1; const float* w4 = smf + F16_F_W4; const float* sc = smf + F16_F_SC;
    uint32_t phase = 0;
    const int stride = gridDim.x * F16_SLOTS, mine = blockIdx.x * F16_SLOTS + slot;   // global tile g belongs to slot g % stride

#pragma unroll 1
    for (int pi = 0; pi < mb.n_parts; ++pi) {
    const int n = *mb.count[pi];
    const int n_tiles = (n + 127) / 128;
    // global tile tiles_before + t belongs to slot (tiles_before + t) % stride: the first tile of this part owned by
    // (this CTA, slot s) is t0(s); the CTA takes part in the part iff one of its four slots owns a tile
    bool cta_has = false;
#pragma unroll
    for (int sl = 0; sl < F16_SLOTS; ++sl)
        cta_has |= (((int)blockIdx.x * F16_SLOTS + sl - tiles_before) % stride + stride) % stride < n_tiles;
    const int t0 = ((mine - tiles_before) % stride + stride) % stride;
    tiles_before += n_tiles;
    if (!cta_has) continue;                                           // block-uniform
    __syncthreads();                                                  // every slot is done with the previous part's panels
    const PairRec* __restrict__ pl = mb.pl[pi];
    const float* __restrict__ el = mb.el[pi];
    // tile t's input rows -> the slot's staging buffer: two bulk copies (embedding rows, pair records) on the slot's "full" barrier
    auto prefetch = [&](int t) {
        const uint32_t rows = (uint32_t)min(128, n - t * 128);
        mbar_expect_tx(bar_full, rows * (NVR_EMB_STRIDE * 4 + 32));
        bulk_g2s(s_stage, el + (size_t)t * 128 * NVR_EMB_STRIDE, rows * NVR_EMB_STRIDE * 4, bar_full);
        bulk_g2s(s_stage + F16_STAGE_E_BYTES, pl + (size_t)t * 128, rows * 32, bar_full);
    };
    if (tid == 0) {   // parameter block -> shared memory: ONE 54 KB bulk copy
        mbar_expect_tx(bar_w, F16_BLOCK_FLOATS * 4);
        bulk_g2s(s_w, mb.blk[pi], F16_BLOCK_FLOATS * 4, bar_w);
    }
    if (stid == 0 && t0 < n_tiles) prefetch(t0);
    mbar_wait(bar_w, phase_w); phase_w ^= 1;
    const bool three = mb.n_rgb[pi] == 3;
    const int part = mb.out_part[pi];

    for (int tile = t0; tile < n_tiles; tile += stride) {
        const int row = tile * 128 + stid;
        const int sr = min(stid, n - 1 - tile * 128);               // rows past the list's end re-read its last row
        int surv;
        // ---- x = [e 19 | pe 27 | 0 0] -> X_hi / X_lo panels (row = stid): six 16-byte chunks of 8 halves each
        mbar_wait(bar_full, phase_full); phase_full ^= 1;            // the tile's rows have landed in the staging buffer
        {
            float x[48];
            const float4* e4 = reinterpret_cast<const float4*>(stage + (size_t)sr * NVR_EMB_STRIDE * 4);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float4 t = e4[q];
                x[q * 4] = t.x; x[q * 4 + 1] = t.y; x[q * 4 + 2] = t.z; x[q * 4 + 3] = t.w;   // x[19] is overwritten below
            }
            const float4* r4 = reinterpret_cast<const float4*>(stage + F16_STAGE_E_BYTES + (size_t)sr * 32);
            const float4 ra = r4[0], rb = r4[1];                     // PairRec: x y z vx | vy vz surv pad
            surv = __float_as_int(rb.z);
            const float v[3] = {ra.w, rb.x, rb.y};
            posenc27_doubling(v, x + 19);                           // part_base_network.py:54
            x[46] = 0.0f; x[47] = 0.0f;
#pragma unroll
            for (int c = 0; c < F16_KX / 8; ++c) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float h0, l0, h1, l1;
                    split11(x[c * 8 + 2 * i], h0, l0);
                    split11(x[c * 8 + 2 * i + 1], h1, l1);
                    hw[i] = pack_h2(h0, h1);
                    lw[i] = pack_h2(l0, l1);
                }
                reinterpret_cast<uint4*>(x_hi)[c * 128 + stid] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                reinterpret_cast<uint4*>(x_lo)[c * 128 + stid] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();                                           // also orders the previous tile's TMEM reads
        slot_sync128(slot);
        // ---- GEMM 0: h_pre = W0 e (accumulator in the H columns), and the x half of GEMM 2 queued right behind it into D:
        //      it needs nothing from the first epilogue and runs under it (covered by GEMM 2's commit, not by this one)
        if (stid == 0) {
            if (tile + stride < n_tiles) prefetch(tile + stride);    // every thread of the slot has read its staged row (barrier above)
            tc_fence_after();
            gemm3h_ss<F16_K0, F16_K0>(tbase + F16_COL_H, s_xhi, s_xlo, s_p0, idesc, true);
            umma_commit(bar_a);
            gemm3h_ss<F16_KX, F16_KX>(tbase + F16_COL_D, s_xhi, s_xlo, s_px, idesc, true);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        tc_fence_after();
        float a[64];
        // ---- epilogue 1: the whole accumulator row first (its columns are about to be overwritten by the packed operand)
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16(trow + F16_COL_H + c * 16, a + c * 16);
        tmem_wait_ld();
        float o0 = sc[0];
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            a[k] = nvr_softplus_hidden(a[k] + b0[k]);               // MLP.forward :20-22
            o0 += w1[k] * a[k];
        }
        store_split_h(trow + F16_COL_H, trow + F16_COL_H + 32, a);
        tmem_wait_st();
        tc_fence_before();
        slot_sync128(slot);
        const float occ = 1.0f - expf(-nvr_softplus(o0));           // :51
        // ---- GEMM 2: g_pre = W2x x (queued with GEMM 0) + M h
        if (stid == 0) {
            tc_fence_after();
            gemm3h_ts<F16_KH>(tbase + F16_COL_D, tbase + F16_COL_H, tbase + F16_COL_H + 32, s_pm, idesc, false);
            umma_commit(bar_a);
        }
        mbar_wait(bar_a, phase); phase ^= 1;
        tc_fence_after();
        if (three) {                                                // block-uniform
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld16(trow + F16_COL_D + c * 16, a + c * 16);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 64; ++k) a[k] = nvr_softplus_hidden(a[k] + b2[k]);
            store_split_h(trow + F16_COL_H, trow + F16_COL_H + 32, a);
            tmem_wait_st();
            tc_fence_before();
            slot_sync128(slot);
            // ---- GEMM 3: W3 g   (accumulator back in D: GEMM 2's result has been consumed)
            if (stid == 0) {
                tc_fence_after();
                gemm3h_ts<F16_KH>(tbase + F16_COL_D, tbase + F16_COL_H, tbase + F16_COL_H + 32, s_p3, idesc, true);
                umma_commit(bar_a);
            }
            mbar_wait(bar_a, phase); phase ^= 1;
            tc_fence_after();
        }
        const float* bl = three ? b3 : b2;
        // ---- last hidden activation + rgb = sigmoid(W4 g + b4)   (:58), raw = [rgb, occ] (:60)
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16(trow + F16_COL_D + c * 16, a + c * 16);
        tmem_wait_ld();
        float r[3] = {sc[1], sc[2], sc[3]};
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const float g = nvr_softplus_hidden(a[k] + bl[k]);
            r[0] += w4[k] * g; r[1] += w4[64 + k] * g; r[2] += w4[128 + k] * g;
        }
        if (row < n)
            raws[(long long)surv * out_stride + part] = make_float4(nvr_sigmoid(r[0]), nvr_sigmoid(r[1]), nvr_sigmoid(r[2]), occ);
    }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tbase_slot, 512);
}

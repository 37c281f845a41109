// nvr_kernels.cuh -- the sm_100a kernels of the per-ray hot path.
//
//   k_frame_all    per frame, one launch: distance channel of pbw -> compact volume | its coarse-minimum grid | cluster re-posing
//   k_cull         sample gen (ray mode, depth-major walk) / point fetch, quick world-space cull, world->pose,
//                  exact distance cull, per-warp compaction in shared memory, one atomic per 2048 positions
//   k_cluster_verts per frame: balanced KD partition of each part's vertices into clusters + AABBs
//                  (the KNN acceleration structure); k_frame_all's cluster blocks re-pose it every frame
//   k_knn          per survivor: 5x exact K=4 NN (group search over the clusters; whole-warp short cuts for far-field
//                  and certainly-unflagged parts), Gaussian weights, per-part append of flagged (sample, part)
//                  neighbour records; far-field pairs are answered by one shared pair per part
//   k_warp         per evaluated pair: blend weights, LBS to big pose, deformer -> canonical point + dir
//   k_embed        THE gather: two lanes per 64-byte row of the dense+hashed grids (one 256-bit load each), per-level sums
//   k_mlp          occ + rgb MLPs on 128-pair tiles (fp32 FFMA register tiles; the tcgen05 version is nvr_mlp_tc.cuh)
//   k_resolve      far-field marker substitution, arg-max part fusion; per-sample raw/occ and/or per-ray alpha compositing
//
// No host synchronisation anywhere: list lengths live in device counters, every consumer kernel is
// a persistent grid-stride loop that reads its trip count from them.
#pragma once
#include <cuda_runtime.h>

#include "nvr_math.cuh"
#include "nvr_frame.cuh"

#define NVR_EMB_STRIDE 20          // 19 used
#define NVR_CTR_SURV 0
#define NVR_CTR_PAIR 1             // [1..5]
#define NVR_CTR_FAR 6              // [6..10] flagged pairs answered by the part's shared far-field pair
#define NVR_CTR_WORK 11             // k_knn's dynamic work-unit counter
#define NVR_CTR_EMBED_WORK 12       // [12..16] k_embed_parts' work-unit counters, one per part
#define NVR_CTR_UNITS 17            // k_cull's count of KNN unit descriptors (groups of <= 32 neighbouring survivors)
#define NVR_CTR_WORDS 32
// Far-field pairs.  A part farther than ~0.73 m from a sample still gets flagged (its Gaussian weights sum to far less
// than the 1e-8 in the normalisation, so pdist -> 0 < smpl_thresh: DESIGN.md section 1).  Once sum(w) < NVR_FAR_WSUM the
// normalised weights are < 1e-12, the blended transforms are ~1e-12, and the canonical point and direction the part
// network sees are the origin to within 5e-13 m / 1e-29 -- far below one fp32 ulp of everything they are added to
// (the residual, the bbox corner).  All such pairs of a part therefore share ONE evaluation: k_knn answers them with a
// marker (occ = -1) and appends a single zero-weight pair per part whose result every marker resolves to.
#define NVR_FAR_WSUM 1e-20f
// 4 exp(-d2 / 0.01125) < 1e-20  <=>  d2 > 0.01125 (ln 4 + 20 ln 10) = 0.5337: when even the nearest cluster box of a part is
// farther than this from the box of a warp's queries, all its lanes are far-field for the part and the 4-NN search is
// skipped (1 % margin on the exponent for expf / lower-bound rounding).
#define NVR_FAR_D2 0.54f
// Certainly-unflagged parts.  pdist = sum(w d) / (sum(w) + 1e-8) >= dmin * sum(w) / (sum(w) + 1e-8).  When every vertex of
// the part is at least dmin = 1.04 smpl_thresh from every query of the warp (5.2 cm) and the 4th-nearest distance of
// every query is at most sqrt(NVR_REACH_D2) = 0.41 m (then sum(w) >= 4 exp(-0.17 / 0.01125) = 1.1e-6 and the ratio is
// >= 0.991), pdist >= 1.03 smpl_thresh: the part is NOT flagged for any lane, whatever its exact neighbours are, and
// the search is skipped.
#define NVR_GAP_FACTOR2 1.0816f    // 1.04^2
#define NVR_REACH_D2 0.17f

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
    const float* dist_cmin;        // per coarse cell minimum of dist (nvr_cull_early_out); null = no early-out
    VolumeDev tuv;                 // (D',H',W',2)
    const float4* verts;           // part vertices, spatially sorted: one 16-float4 SoA block per cluster (x | y | z | orig index)
    const float4* cl_lo;           // per-cluster AABB
    const float4* cl_hi;
    const int* cl_off;             // [6] cluster offsets per part (device)
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
__device__ __forceinline__ void frame_prep_part(const float* __restrict__ pbw, int n_vox, int C, float* __restrict__ dist, int bid, int nb) {
    const int tid = bid * blockDim.x + threadIdx.x, nth = nb * blockDim.x;
    for (int i = tid; i < n_vox; i += nth) dist[i] = pbw[(long long)i * C + (C - 1)];
}

// coarse minimum grid of the distance channel: one WARP per coarse cell, the lanes share out the up to 6^3 voxels
// nvr_coarse_min visits (same set, same NaN rule: a NaN voxel poisons the cell), then a shuffle minimum.  Reads the channel
// straight from pbw (stride C), so it does not depend on the compact copy made next to it.
__device__ __forceinline__ void frame_coarse_part(const float* __restrict__ pbw, int C, int D, int H, int W, float* __restrict__ cmin, int bid, int nb) {
    const int cD = nvr_coarse_dim(D), cH = nvr_coarse_dim(H), cW = nvr_coarse_dim(W);
    const int lane = threadIdx.x & 31;
    const int n_cells = cD * cH * cW;
    for (int i = (bid * blockDim.x + threadIdx.x) >> 5; i < n_cells; i += (nb * blockDim.x) >> 5) {
        const int cz = i / (cH * cW), cy = (i / cW) % cH, cx = i % cW;
        const int z0 = cz * NVR_CULL_B > 0 ? cz * NVR_CULL_B - 1 : 0, z1 = min(cz * NVR_CULL_B + NVR_CULL_B + 1, D - 1);
        const int y0 = cy * NVR_CULL_B > 0 ? cy * NVR_CULL_B - 1 : 0, y1 = min(cy * NVR_CULL_B + NVR_CULL_B + 1, H - 1);
        const int x0 = cx * NVR_CULL_B > 0 ? cx * NVR_CULL_B - 1 : 0, x1 = min(cx * NVR_CULL_B + NVR_CULL_B + 1, W - 1);
        const int nz = z1 - z0 + 1, ny = y1 - y0 + 1, nx = x1 - x0 + 1;
        float m = INFINITY;
        bool nan = false;
        for (int t = lane; t < nz * ny * nx; t += 32) {
            const int z = z0 + t / (ny * nx), y = y0 + (t / nx) % ny, x = x0 + t % nx;
            const float d = pbw[(((long long)z * H + y) * W + x) * C + (C - 1)];
            nan |= d != d;
            m = fminf(m, d);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, s));
        nan = __any_sync(0xffffffffu, nan);
        if (lane == 0) cmin[i] = nan ? __int_as_float(0x7fc00000) : m;
    }
}

// Per-frame KNN acceleration structure, one CTA per part: a balanced KD partition of the part's posed
// vertices into clusters of NVR_CL (median splits along the longest axis of each segment's bounding
// box, left half rounded to whole clusters), plus each cluster's AABB.  Every level is ONE bitonic
// sort in shared memory of 64-bit keys (segment | order-preserving coordinate bits | vertex index),
// so segments are sorted side by side; ~log2(#clusters) levels, tens of microseconds per frame.
// (Tighter boxes than a Morton curve: ~1.7x fewer clusters survive the lower-bound test.)
// The search in k_knn stays EXACT whatever the partition: clusters are only skipped when their AABB
// lower bound exceeds the current 4th-best distance.  Padding vertices are +inf (never selected).
#define NVR_SORT_MAX 8192
#define NVR_CL_MAX (NVR_SORT_MAX / NVR_CL)
#define NVR_CLUSTER_SMEM (NVR_SORT_MAX * 8 + NVR_CL_MAX * (2 * 2 + 6 * 4))

__device__ __forceinline__ unsigned int float_order_bits(float f) {   // monotone float -> uint
    const unsigned int u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

__global__ void __launch_bounds__(1024)
k_cluster_verts(const float* __restrict__ part_pts, const long long* __restrict__ lengths2, int maxlen,
                int* __restrict__ perm, int* __restrict__ cl_off) {
    extern __shared__ unsigned long long s_key[];                  // [NVR_SORT_MAX]
    unsigned short* s_ra = reinterpret_cast<unsigned short*>(s_key + NVR_SORT_MAX);   // segment [ra, rb) of each cluster
    unsigned short* s_rb = s_ra + NVR_CL_MAX;
    unsigned int* s_box = reinterpret_cast<unsigned int*>(s_rb + NVR_CL_MAX);         // [6][NVR_CL_MAX] order bits, by segment start
    const int part = blockIdx.x, tid = threadIdx.x;
    int coff = 0, n = 0;
    for (int p = 0; p < NVR_PARTS; ++p) {
        long long len = lengths2[p];
        len = len < 0 ? 0 : (len > maxlen ? maxlen : len);
        if (p == part) n = (int)len;
        if (p < part) coff += (int)((len + NVR_CL - 1) / NVR_CL);
    }
    const int ncl = (n + NVR_CL - 1) / NVR_CL;
    if (tid == 0) cl_off[part + 1] = coff + ncl;
    if (part == 0 && tid == 0) cl_off[0] = 0;
    const float* src = part_pts + (long long)part * maxlen * 3;
    int M = 32;
    while (M < n) M <<= 1;
    for (int j = tid; j < M; j += blockDim.x) s_key[j] = j < n ? (unsigned long long)j : ~0ull;
    for (int c = tid; c < ncl; c += blockDim.x) { s_ra[c] = 0; s_rb[c] = (unsigned short)ncl; }
    __syncthreads();
    int span = ncl;                                                // largest segment, in clusters
    while (span > 1) {
        // 1. bounding box of every segment (keyed by its first cluster)
        for (int c = tid; c < ncl; c += blockDim.x)
            if (s_ra[c] == c)
#pragma unroll
                for (int a = 0; a < 3; ++a) { s_box[a * NVR_CL_MAX + c] = 0xffffffffu; s_box[(3 + a) * NVR_CL_MAX + c] = 0u; }
        __syncthreads();
        for (int i0 = tid - (tid & 31); i0 < n; i0 += blockDim.x) {   // warp-uniform trip count
            const int i = i0 + (tid & 31);
            const bool valid = i < n;
            const int j = valid ? (int)(s_key[i] & 0x1fffull) : 0, seg = valid ? (int)s_ra[i / NVR_CL] : -1;
            const bool whole_warp = __match_any_sync(0xffffffffu, seg) == 0xffffffffu;   // one segment, all valid
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const unsigned int u = float_order_bits(src[j * 3 + a]);
                if (whole_warp) {
                    const unsigned int lo = __reduce_min_sync(0xffffffffu, u), hi = __reduce_max_sync(0xffffffffu, u);
                    if ((tid & 31) == 0) { atomicMin(&s_box[a * NVR_CL_MAX + seg], lo); atomicMax(&s_box[(3 + a) * NVR_CL_MAX + seg], hi); }
                } else if (valid) {
                    atomicMin(&s_box[a * NVR_CL_MAX + seg], u);
                    atomicMax(&s_box[(3 + a) * NVR_CL_MAX + seg], u);
                }
            }
        }
        __syncthreads();
        // 2. keys: segment | coordinate along the segment's longest axis | vertex
        for (int i = tid; i < n; i += blockDim.x) {
            const int j = (int)(s_key[i] & 0x1fffull), seg = s_ra[i / NVR_CL];
            int ax = 0;
            float best = -1.0f;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                // order bits back to floats: extent = hi - lo
                unsigned int ul = s_box[a * NVR_CL_MAX + seg], uh = s_box[(3 + a) * NVR_CL_MAX + seg];
                ul ^= (ul >> 31) ? 0x80000000u : 0xffffffffu;
                uh ^= (uh >> 31) ? 0x80000000u : 0xffffffffu;
                const float e = __uint_as_float(uh) - __uint_as_float(ul);
                if (e > best) { best = e; ax = a; }
            }
            s_key[i] = ((unsigned long long)seg << 45) | ((unsigned long long)float_order_bits(src[j * 3 + ax]) << 13) | (unsigned)j;
        }
        __syncthreads();
        // 3. one bitonic sort orders every segment along its own axis
        for (int k = 2; k <= M; k <<= 1)
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int i = tid; i < M; i += blockDim.x) {
                    const int ixj = i ^ jj;
                    if (ixj > i) {
                        const unsigned long long x = s_key[i], y = s_key[ixj];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { s_key[i] = y; s_key[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        // 4. split every segment of more than one cluster at its middle cluster
        for (int c = tid; c < ncl; c += blockDim.x) {
            const int ra = s_ra[c], rb = s_rb[c];
            if (rb - ra > 1) {
                const int mid = ra + (rb - ra) / 2;
                if (c < mid) s_rb[c] = (unsigned short)mid; else s_ra[c] = (unsigned short)mid;
            }
        }
        span = (span + 1) / 2;
        __syncthreads();
    }
    // the partition: slot -> vertex index within the part (-1 = padding); positions are filled by k_cluster_apply
    for (int i = tid; i < ncl * NVR_CL; i += blockDim.x) perm[(long long)coff * NVR_CL + i] = i < n ? (int)(s_key[i] & 0x1fffull) : -1;
}

// Every frame: gather the posed vertices into cluster order and recompute the cluster AABBs.
// 16 lanes per cluster (NVR_CL == 16): lane = slot, min/max by shuffles inside the half-warp.
__device__ __forceinline__ void cluster_apply_part(const float* __restrict__ part_pts, int maxlen, const int* __restrict__ perm, const int* __restrict__ cl_off,
                float4* __restrict__ verts, float4* __restrict__ cl_lo, float4* __restrict__ cl_hi, int bid, int nb) {
    static_assert(NVR_CL == 16, "cluster_apply maps one half-warp to one cluster");
    const int total = cl_off[NVR_PARTS];
    const int lane16 = threadIdx.x & 15;
    for (int c = (bid * blockDim.x + threadIdx.x) >> 4; c < ((total + 1) & ~1); c += (nb * blockDim.x) >> 4) {
        const bool in = c < total;                                 // clusters are handled in warp-wide pairs
        int part = 0;
#pragma unroll
        for (int p = 1; p < NVR_PARTS; ++p) part += (in && c >= cl_off[p]) ? 1 : 0;
        const int j = in ? perm[(long long)c * NVR_CL + lane16] : -1;
        float x = INFINITY, y = INFINITY, z = INFINITY;
        if (j >= 0) {
            const float* s = part_pts + ((long long)part * maxlen + j) * 3;
            x = s[0]; y = s[1]; z = s[2];
        }
        if (in) {                                                  // structure-of-arrays cluster block (nvr_knn_scan)
            float* blk = reinterpret_cast<float*>(verts + (long long)c * NVR_CL);
            blk[lane16] = x; blk[16 + lane16] = y; blk[32 + lane16] = z; blk[48 + lane16] = __int_as_float(j >= 0 ? j : 0);
        }
        float lo[3] = {x, y, z}, hi[3] = {j >= 0 ? x : -INFINITY, j >= 0 ? y : -INFINITY, j >= 0 ? z : -INFINITY};
        int cnt = j >= 0 ? 1 : 0;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
            }
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        }
        if (in && lane16 == 0) {
            cl_lo[c] = make_float4(lo[0], lo[1], lo[2], __int_as_float(cnt));
            cl_hi[c] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
    }
}

// The three per-frame preparation steps are independent of each other (the coarse grid reads pbw directly), so they are ONE
// launch: blocks [0, nb_prep) copy the distance channel, [nb_prep, nb_prep + nb_coarse) build the coarse minimum grid, the rest
// re-pose the vertex clusters.  128 threads per block.
__global__ void __launch_bounds__(128)
k_frame_all(const float* __restrict__ pbw, int C, int D, int H, int W, float* __restrict__ dist, float* __restrict__ cmin,
            const float* __restrict__ part_pts, int maxlen, const int* __restrict__ perm, const int* __restrict__ cl_off,
            float4* __restrict__ verts, float4* __restrict__ cl_lo, float4* __restrict__ cl_hi, int nb_prep, int nb_coarse) {
    const int b = blockIdx.x;
    if (b < nb_prep) frame_prep_part(pbw, D * H * W, C, dist, b, nb_prep);
    else if (b < nb_prep + nb_coarse) frame_coarse_part(pbw, C, D, H, W, cmin, b - nb_prep, nb_coarse);
    else cluster_apply_part(part_pts, maxlen, perm, cl_off, verts, cl_lo, cl_hi, b - nb_prep - nb_coarse, gridDim.x - nb_prep - nb_coarse);
}

// Exact K=4 nearest vertices of one part for the 32 queries of a warp (one per lane), as a GROUP
// search: the warp's survivors are neighbouring samples of one or two rays, so they share most of
// their candidate set.  qlo/qhi = AABB of the warp's live queries (warp-uniform).
//   1. lanes share out the part's clusters: U = min over clusters (>= 4 vertices) of the largest
//      possible query-to-vertex distance bounds EVERY lane's 4th-nearest distance; the cluster whose
//      box is nearest to the query box is scanned first by all lanes (tight per-lane bounds early);
//   2. clusters whose box-to-box lower bound exceeds U are dropped for the whole warp (ballot);
//   3. each remaining cluster is scanned (whole warp, broadcast loads) only if ANY lane's own AABB
//      lower bound does not exceed its current 4th-best distance.
// Nothing is approximated: a cluster is skipped only when it cannot contain a better neighbour.
__device__ __forceinline__ float box_gap2(const float4& lo, const float4& hi, const float qlo[3], const float qhi[3]) {
    const float dx = fmaxf(fmaxf(lo.x - qhi[0], qlo[0] - hi.x), 0.0f);
    const float dy = fmaxf(fmaxf(lo.y - qhi[1], qlo[1] - hi.y), 0.0f);
    const float dz = fmaxf(fmaxf(lo.z - qhi[2], qlo[2] - hi.z), 0.0f);
    return dx * dx + dy * dy + dz * dz;
}
__device__ __forceinline__ float box_reach2(const float4& lo, const float4& hi, const float qlo[3], const float qhi[3]) {
    const float dx = fmaxf(fabsf(qhi[0] - lo.x), fabsf(hi.x - qlo[0]));
    const float dy = fmaxf(fabsf(qhi[1] - lo.y), fabsf(hi.y - qlo[1]));
    const float dz = fmaxf(fabsf(qhi[2] - lo.z), fabsf(hi.z - qlo[2]));
    return dx * dx + dy * dy + dz * dz;
}

// Returns KNN_FAR / KNN_UNFLAGGED (and leaves k untouched) when allow_skip and the whole part is far-field / certainly
// unflagged for every query of the warp; KNN_SEARCHED otherwise.
#define KNN_SEARCHED 0
#define KNN_FAR 1
#define KNN_UNFLAGGED 2
__device__ __forceinline__ int knn_part_group(const FrameDev& fr, int part, const float p[3], bool live,
                                              const float qlo[3], const float qhi[3], Knn4& k, bool allow_skip, float thresh,
                                              bool allow_unflagged_skip = true) {
    const int lane = threadIdx.x & 31;
    const int c0 = fr.cl_off[part], ncl = fr.cl_off[part + 1] - c0;
    if (ncl <= 0) return KNN_SEARCHED;
    float U = INFINITY, best = INFINITY;
    int seed = 0, nv = 0;
    for (int c = lane; c < ncl; c += 32) {
        const float4 lo = __ldg(fr.cl_lo + c0 + c), hi = __ldg(fr.cl_hi + c0 + c);
        nv += __float_as_int(lo.w);
        if (__float_as_int(lo.w) >= NVR_KNN) U = fminf(U, box_reach2(lo, hi, qlo, qhi));
        const float g = box_gap2(lo, hi, qlo, qhi);
        if (g < best) { best = g; seed = c; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        U = fminf(U, __shfl_xor_sync(0xffffffffu, U, d));
        const float ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int os = __shfl_xor_sync(0xffffffffu, seed, d);
        if (ob < best || (ob == best && os < seed)) { best = ob; seed = os; }
    }
    // every vertex of the part is farther than sqrt(best) from every query: with >= 4 real vertices all four
    // neighbours are, so sum(w) < NVR_FAR_WSUM for every lane whatever the search would return
    if (allow_skip && best > NVR_FAR_D2 && __reduce_add_sync(0xffffffffu, nv) >= NVR_KNN) return KNN_FAR;
    // U is finite only if some cluster holds >= 4 real vertices, so the four neighbours are real here too
    if (allow_skip && allow_unflagged_skip && best > NVR_GAP_FACTOR2 * thresh * thresh && U <= NVR_REACH_D2) return KNN_UNFLAGGED;
    U *= 1.00001f;
    // the seed first: every live lane's 4th-best is still +inf, so all of them take it
    nvr_knn_scan(fr.verts + (long long)(c0 + seed) * NVR_CL, p, k);
    // then rounds of 32 clusters, one per lane.  Inside a round the clusters are taken BEST-FIRST (smallest box-to-box
    // gap) and the group bound -- no lane needs a cluster whose gap exceeds the LARGEST current 4th-best distance of the
    // warp (non-negative floats order like their bits) -- is refreshed after every scan, so the few nearest clusters
    // tighten it and most of the round's other clusters are dropped by one vote without being looked at per lane.
    for (int cb = 0; cb < ncl; cb += 32) {
        const int c = cb + lane;
        bool avail = c < ncl && c != seed;                        // this lane's cluster has not been taken yet
        float gapc = 0.0f;
        if (avail) gapc = box_gap2(__ldg(fr.cl_lo + c0 + c), __ldg(fr.cl_hi + c0 + c), qlo, qhi) * NVR_PRUNE_SLACK;
        while (true) {
            const float worst = __uint_as_float(__reduce_max_sync(0xffffffffu, live ? (unsigned int)(k.key[3] >> 32) : 0u));
            const float bound = fminf(U, worst * 1.00001f);
            const bool cand = avail && !(gapc > bound);
            // smallest gap among the remaining candidates; the lane index rides in the 5 low mantissa bits
            const unsigned int sel = __reduce_min_sync(0xffffffffu, cand ? ((__float_as_uint(gapc) & ~31u) | (unsigned)lane) : 0xffffffffu);
            if (sel == 0xffffffffu) break;
            const int cc = cb + (int)(sel & 31u);
            if (lane == (int)(sel & 31u)) avail = false;
            const float lb = nvr_aabb_lb(__ldg(fr.cl_lo + c0 + cc), __ldg(fr.cl_hi + c0 + cc), p);
            const bool need = live && !(lb * NVR_PRUNE_SLACK > nvr_knn_d2(k, 3));
            if (__any_sync(0xffffffffu, need)) nvr_knn_scan(fr.verts + (long long)(c0 + cc) * NVR_CL, p, k);
        }
    }
    return KNN_SEARCHED;
}

// -----------------------------------------------------------------------------------------
// cull: samples -> survivors
// -----------------------------------------------------------------------------------------
// ray mode  (n_samples > 0): sample i = ray i / n_samples, step i % n_samples; pts = ray_o, aux = ray_d
// point mode (n_samples == 0): pts = wpts (n,3)
// Ray mode walks the samples DEPTH-MAJOR inside groups of 32 consecutive rays, in chunks of 16 rays x 2 depth steps
// (cull_locate, nvr_math.cuh), so the survivor list -- and with it every later kernel's warps -- holds neighbouring
// pixels at neighbouring depths (a few cm across) instead of the entry and exit shells of one ray (tens of cm apart):
// tighter KNN query boxes, shared neighbour rows and grid cells.  A CTA compacts CULL_T chunks of 256 positions (64 depth
// steps of one ray group) in shared memory and appends them with ONE atomic, so the runs of neighbouring survivors are
// hundreds long and the survivor records leave as coalesced 16-byte stores.  Results are per sample and do not depend on
// the order.
#define CULL_T 8
#define CULL_SPAN (256 * CULL_T)
// surv_of_sample must be pre-filled with -1 (cudaMemsetAsync 0xFF): only survivors' entries are written here.
__global__ void __launch_bounds__(256, 4)
k_cull(FrameDev fr, const float* __restrict__ pts, const float* __restrict__ ray_d,
       const float* __restrict__ near_, const float* __restrict__ far_, long long n, int n_samples,
       float thresh, int* __restrict__ counters, int* __restrict__ surv_of_sample, float4* __restrict__ surv, int keep_all,
       int2* __restrict__ units) {
    // keep_all (NVR_TUNE_DENSE_A1, measurement variant): every valid sample survives, whatever its distance
    // units: KNN work units (first survivor slot, 1..32 survivors).  A unit never crosses a span -- two spans are appended in
    // completion order and can lie anywhere in the frame -- nor an empty warp segment inside a span (8 depth steps = the gap
    // between a ray group's entry and exit shells), so the 32 queries a k_knn warp searches for together are neighbours.
    // With units cut every 32 slots of the survivor list, ~2 in 5 of them straddled two regions tens of cm apart and the
    // group search pruned against a query box spanning both.
    __shared__ float4 s_surv[CULL_SPAN];
    __shared__ int warp_cnt[8];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool rays = n_samples > 0;
    CullWalk cw;
    cw.S = rays ? n_samples : 1;
    cw.group = cull_group_positions(cw.S);
    cw.n_rays = rays ? n / n_samples : 0;
    const long long n_map = rays ? ((cw.n_rays + 31) / 32) * (long long)cw.group : n;
    __shared__ CullQuick cq;                                      // uniform: one copy per CTA, broadcast reads
    const bool quick = fr.dist_cmin != nullptr && !keep_all;
    if (quick && threadIdx.x == 0) nvr_cull_quick_setup(fr.dist, fr.R, fr.Th, cq);
    __syncthreads();
    for (long long sbase = (long long)blockIdx.x * CULL_SPAN; sbase < n_map; sbase += (long long)gridDim.x * CULL_SPAN) {
        cw.g0 = sbase / (long long)cw.group;
        cw.w0 = (unsigned)(sbase - cw.g0 * (long long)cw.group);
        // warp `wid` owns positions [wid * 256, wid * 256 + 256) of the span -- 8 consecutive depth steps of the 32 rays, as
        // 4 chunks (16 rays x 2 steps) of one half of the rays and then 4 of the other half --
        // and compacts them into its own 256-record segment of s_surv: no CTA barrier inside the loop
        int run = 0;                                              // survivors of this warp so far (warp-uniform)
        long long r_have = -1;                                    // the ray whose data the registers below hold
        float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f}, nr = 0.f, fa = 0.f, qa[3] = {0.f, 0.f, 0.f}, qb[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int t = 0; t < CULL_T; ++t) {
            const int local = (wid * CULL_T + t) * 32 + lane;
            long long i = sbase + local, r = 0;                   // i = sample id
            int k = 0;
            bool valid = i < n;
            if (rays) valid = cull_locate(cw, local, r, k, i) && sbase + local < n_map;
            bool keep = false;
            float p[3] = {0.f, 0.f, 0.f};
            if (valid) {
                float w[3], c[3];
                bool culled = false;
                if (rays) {
                    if (r != r_have) {                            // a lane changes ray once per span (the group's other 16 rays)
                        r_have = r;
#pragma unroll
                        for (int a = 0; a < 3; ++a) { o[a] = pts[r * 3 + a]; d[a] = ray_d[r * 3 + a]; }
                        nr = near_[r]; fa = far_[r];
                        if (quick) nvr_cull_quick_ray(cq, o, d, qa, qb);
                    }
                    if (quick) {
                        const float tk = nvr_linspace01(k, n_samples);
                        const float z = nr * (1.0f - tk) + fa * tk;
#pragma unroll
                        for (int a = 0; a < 3; ++a) c[a] = qa[a] + z * qb[a];
                        culled = nvr_cull_quick(fr.dist, cq, fr.dist_cmin, c, thresh);
                    }
                    if (!culled) nvr_ray_sample(o, d, nr, fa, k, n_samples, w);
                } else {
                    w[0] = pts[i * 3]; w[1] = pts[i * 3 + 1]; w[2] = pts[i * 3 + 2];
                    if (quick) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) c[a] = ((w[0] * cq.M[a] + w[1] * cq.M[3 + a]) + w[2] * cq.M[6 + a]) + cq.t[a];
                        culled = nvr_cull_quick(fr.dist, cq, fr.dist_cmin, c, thresh);
                    }
                }
                if (!culled) {                                    // the reference's arithmetic, bit for bit
                    nvr_world_to_pose(fr.R, fr.Th, w, p);
                    nvr_volume_coords(fr.dist, p, c);
                    if (keep_all) keep = true;
                    else if (!(quick && nvr_cull_early_out(fr.dist, fr.dist_cmin, c, thresh))) {
                        float pn;
                        nvr_sample_volume_at(fr.dist, c, 0, 1, &pn);
                        keep = pn < thresh;                       // inb_part_network_multiassign.py:136
                    }
                }
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (keep) s_surv[wid * 256 + run + __popc(ballot & ((1u << lane) - 1u))] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));   // passes are < 2^31 samples
            run += __popc(ballot);
        }
        if (lane == 0) warp_cnt[wid] = run;
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; off += w < wid ? c : 0; total += c; }
        if (threadIdx.x == 0) {
            s_base = total ? atomicAdd(&counters[NVR_CTR_SURV], total) : 0;
            if (total && units) {                                  // runs of consecutive non-empty warp segments -> units of <= 32
                int nu = 0, runlen = 0;
#pragma unroll
                for (int w = 0; w <= 8; ++w) {
                    const int c = w < 8 ? warp_cnt[w] : 0;
                    if (c) runlen += c;
                    else { nu += (runlen + 31) >> 5; runlen = 0; }
                }
                int u = atomicAdd(&counters[NVR_CTR_UNITS], nu);
                int start = s_base;
                runlen = 0;
#pragma unroll
                for (int w = 0; w <= 8; ++w) {
                    const int c = w < 8 ? warp_cnt[w] : 0;
                    if (c) runlen += c;
                    else {
                        for (int done = 0; done < runlen; done += 32) units[u++] = make_int2(start + done, min(32, runlen - done));
                        start += runlen; runlen = 0;
                    }
                }
            }
        }
        __syncthreads();
        const int gbase = s_base + off;
        for (int x = lane; x < run; x += 32) {                    // each warp appends its own segment
            const float4 sv = s_surv[wid * 256 + x];
            surv[gbase + x] = sv;
            surv_of_sample[__float_as_int(sv.w)] = gbase + x;
        }
        __syncwarp();                                             // the segment is rewritten by this warp only
    }
}

// Training: rank_of_slot[slot] = position of the survivor in ASCENDING SAMPLE ORDER (the order of the reference's
// nonzero(), inb_part_network_multiassign.py:137), from the sample -> slot map.  One CTA: every thread counts the survivors of
// its contiguous chunk of samples, a block scan turns the counts into offsets, a second sweep hands out the ranks.
__global__ void __launch_bounds__(1024)
k_rank_slots(const int* __restrict__ surv_of_sample, long long n, int* __restrict__ rank_of_slot) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long chunk = (n + blockDim.x - 1) / blockDim.x, lo = tid * chunk, hi = lo + chunk < n ? lo + chunk : n;
    int cnt = 0;
    for (long long i = lo; i < hi; ++i) cnt += surv_of_sample[i] >= 0;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += up;
        }
        s_warp[lane] = v;
    }
    __syncthreads();
    int rank = incl - cnt + (wid ? s_warp[wid - 1] : 0);
    for (long long i = lo; i < hi; ++i) {
        const int slot = surv_of_sample[i];
        if (slot >= 0) rank_of_slot[slot] = rank++;
    }
}

// -----------------------------------------------------------------------------------------
// knn: survivors -> flagged (sample, part) neighbour records
// -----------------------------------------------------------------------------------------
struct __align__(16) KnnRec {      // a flagged (sample, part) pair before the warp: 48 B
    float w[NVR_KNN];              // normalised Gaussian weights of the 4 neighbours
    int idx[NVR_KNN];              // their rows in part_pbw
    int surv;                      // survivor slot
    int _pad[3];
};

// DENSE (NVR_TUNE_DENSE_A1, measurement variant): every survivor is flagged in exactly ONE part -- the one with the smallest
// weighted neighbour distance (first on ties) -- instead of every part with pdist < thresh; no far-field sharing.
template <int MINB, bool DENSE = false>
__global__ void __launch_bounds__(256, MINB)
k_knn(FrameDev fr, float thresh, int* __restrict__ counters, float4* __restrict__ surv,
      KnnRec* __restrict__ recs, int cap, float4* __restrict__ raws, float* __restrict__ dbg, int far_slot,
      const int2* __restrict__ units) {
    // dbg (optional, per SAMPLE): [n][5][8] = flag, x, y, z, vx, vy, vz, pdist -- per-stage parity tests
    // far_slot >= 0: survivor slot reserved for the shared far-field pairs (NVR_FAR_WSUM); -1 = evaluate every pair
    const int n_surv = counters[NVR_CTR_SURV];
    const int lane = threadIdx.x & 31;
    if (far_slot >= 0 && blockIdx.x == 0 && threadIdx.x < NVR_PARTS && n_surv > 0) {
        const int part = threadIdx.x;
        if (part == 0) surv[far_slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
        KnnRec rec;                                               // zero weights: blended transforms, point and direction are 0
#pragma unroll
        for (int i = 0; i < NVR_KNN; ++i) { rec.w[i] = 0.0f; rec.idx[i] = 0; }
        rec.surv = far_slot; rec._pad[0] = rec._pad[1] = rec._pad[2] = 0;
        recs[(long long)part * cap + atomicAdd(&counters[NVR_CTR_PAIR + part], 1)] = rec;
    }
    // Work unit = (32 consecutive survivors, ONE part), handed out through a device counter: a part search costs anything
    // from a box test (far-field / certainly unflagged) to a dozen cluster scans, so a static grid-stride split leaves the
    // launch waiting for its unluckiest warp -- visibly so when a pass holds only a few survivors per warp slot (an 8-GPU
    // shard of a 512 x 512 frame: 310 k survivors over 4736 warp slots).  DENSE keeps all five parts in one warp (arg-min).
    // Units are numbered part-major, body first: the body part has the most clusters and the most survivors next to it, the
    // arms are far-field / certainly unflagged for most of the frame, so the expensive searches are handed out first and the
    // launch ends on cheap units (ncu on an 8-GPU shard: SM active cycles avg / max 0.71 with the parts interleaved).
    const int n_groups = counters[NVR_CTR_UNITS];                 // k_cull's unit descriptors (groups of neighbouring survivors)
    const int n_units = n_groups * (DENSE ? 1 : NVR_PARTS);
    // the counter fetch for the NEXT unit is issued before this unit is searched, so its round trip (11 % of the stall samples
    // when it sat at the top of the loop, profiles/r2h_line_stalls_k_knn.txt) runs under the search
    int unit = 0;
    if (lane == 0) unit = atomicAdd(&counters[NVR_CTR_WORK], 1);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    for (int next = 0; unit < n_units; unit = __shfl_sync(0xffffffffu, next, 0)) {
        if (lane == 0) next = atomicAdd(&counters[NVR_CTR_WORK], 1);
        const int unit_part = DENSE ? 0 : unit / n_groups;
        const int2 ud = units[DENSE ? unit : unit - unit_part * n_groups];
        const int s = ud.x + lane;
        const bool live = lane < ud.y;
        float p[3] = {0.f, 0.f, 0.f};
        int sample = 0;
        if (live) {
            const float4 sv = surv[s];
            p[0] = sv.x; p[1] = sv.y; p[2] = sv.z;
            sample = __float_as_int(sv.w);
        }
        float qlo[3], qhi[3];                                     // AABB of the warp's live queries
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            qlo[a] = live ? p[a] : INFINITY;
            qhi[a] = live ? p[a] : -INFINITY;
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) {
                qlo[a] = fminf(qlo[a], __shfl_xor_sync(0xffffffffu, qlo[a], dd));
                qhi[a] = fmaxf(qhi[a], __shfl_xor_sync(0xffffffffu, qhi[a], dd));
            }
        }
        if constexpr (DENSE) {
            float best_d = INFINITY;
            int best_part = 0;
            KnnRec best;
#pragma unroll
            for (int i = 0; i < NVR_KNN; ++i) { best.w[i] = 0.0f; best.idx[i] = 0; }
#pragma unroll 1
            for (int part = 0; part < NVR_PARTS; ++part) {
                Knn4 k;
                nvr_knn_init(k);
                // a part that is far-field for the whole warp (every vertex > 0.73 m away) cannot be the nearest one unless all
                // five are; it then keeps the zero-weight record of part 0
                const int group = knn_part_group(fr, part, p, live, qlo, qhi, k, true, thresh, false);
                float w[NVR_KNN];
                const float pdist = group == KNN_FAR ? INFINITY : nvr_knn_weights(k, w);
                if (live) raws[(long long)s * NVR_PARTS + part] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && pdist < best_d) {
                    best_d = pdist; best_part = part;
#pragma unroll
                    for (int i = 0; i < NVR_KNN; ++i) { best.w[i] = w[i]; best.idx[i] = nvr_knn_idx(k, i); }
                }
            }
            best.surv = s; best._pad[0] = best._pad[1] = best._pad[2] = 0;
#pragma unroll 1
            for (int part = 0; part < NVR_PARTS; ++part) {
                const bool flag = live && best_part == part;
                const unsigned ballot = __ballot_sync(0xffffffffu, flag);
                if (!ballot) continue;
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&counters[NVR_CTR_PAIR + part], __popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (flag) recs[(long long)part * cap + wbase + __popc(ballot & ((1u << lane) - 1u))] = best;
            }
        } else {
        {
            const int part = unit_part;
            Knn4 k;
            nvr_knn_init(k);
            const int group = knn_part_group(fr, part, p, live, qlo, qhi, k, far_slot >= 0, thresh);
            bool flag = false;
            KnnRec rec;
            bool far = live && group == KNN_FAR;
            if (far) raws[(long long)s * NVR_PARTS + part] = make_float4(0.f, 0.f, 0.f, -1.0f);
            if (live && group == KNN_UNFLAGGED) raws[(long long)s * NVR_PARTS + part] = make_float4(0.f, 0.f, 0.f, 0.f);   // :201-202
            if (live && group == KNN_SEARCHED) {
                float wsum;
                const float pdist = nvr_knn_weights(k, rec.w, &wsum);
                flag = pdist < thresh;                             // inb_part_network_multiassign.py:90
                far = flag && far_slot >= 0 && wsum < NVR_FAR_WSUM;
                if (far) {                                         // answered by the part's shared far-field pair
                    raws[(long long)s * NVR_PARTS + part] = make_float4(0.f, 0.f, 0.f, -1.0f);
                    flag = false;
                }
                if (dbg) {
                    float* dr = dbg + ((long long)sample * NVR_PARTS + part) * 8;
                    dr[0] = flag ? 1.0f : 0.0f; dr[7] = pdist;
                }
                if (!flag && !far) raws[(long long)s * NVR_PARTS + part] = make_float4(0.f, 0.f, 0.f, 0.f);   // :201-202
            }
            const unsigned fballot = __ballot_sync(0xffffffffu, far);
            if (fballot && lane == 0) atomicAdd(&counters[NVR_CTR_FAR + part], __popc(fballot));
            const unsigned ballot = __ballot_sync(0xffffffffu, flag);
            if (ballot) {
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&counters[NVR_CTR_PAIR + part], __popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (flag) {
#pragma unroll
                    for (int i = 0; i < NVR_KNN; ++i) rec.idx[i] = nvr_knn_idx(k, i);
                    rec.surv = s; rec._pad[0] = rec._pad[1] = rec._pad[2] = 0;
                    recs[(long long)part * cap + wbase + __popc(ballot & ((1u << lane) - 1u))] = rec;
                }
            }
        }
        }
    }
}

// -----------------------------------------------------------------------------------------
// warp: neighbour records -> canonical-space pairs (blend, LBS, deformer); blockIdx.y = part
// -----------------------------------------------------------------------------------------
#define WARP_THREADS 128
#define WARP_SMEM_FLOATS (NVR_DEF_PACKED_FLOATS + 2 * NVR_JOINTS * 16 + 32 * WARP_THREADS)
struct DeformerSmem { const float* pk; float* A; float* bigA; float* scratch; };
__device__ __forceinline__ DeformerSmem stage_deformer(float* sm, const DeformerMlp& g, const float* A, const float* bigA) {
    float* sA = sm + NVR_DEF_PACKED_FLOATS; float* sB = sA + NVR_JOINTS * 16;
    nvr_pack_deformer(g, sm, threadIdx.x, blockDim.x);
    if (A) for (int i = threadIdx.x; i < NVR_JOINTS * 16; i += blockDim.x) { sA[i] = A[i]; sB[i] = bigA[i]; }
    __syncthreads();
    DeformerSmem d;
    d.pk = sm; d.A = sA; d.bigA = sB; d.scratch = sB + NVR_JOINTS * 16;
    return d;
}

template <int MINB>
__global__ void __launch_bounds__(WARP_THREADS, MINB)
k_warp(FrameDev fr, GridDev dg, DeformerMlp dm_g, const float* __restrict__ dirs, int dir_div,
       const int* __restrict__ counters, const float4* __restrict__ surv, const KnnRec* __restrict__ recs,
       PairRec* __restrict__ pairs, int cap, float* __restrict__ dbg, float* __restrict__ out_x0, float* __restrict__ out_resd,
       const int* __restrict__ out_rank) {
    // out_x0 / out_resd (training): (survivors, 5, 3) big-pose point and residual of every flagged (survivor, part), row =
    // out_rank[slot] (the survivor's position in ascending sample order, k_rank_slots) or the slot itself when out_rank is null
    __shared__ __align__(16) float sm[WARP_SMEM_FLOATS];
    const int part = blockIdx.y;
    const int n = counters[NVR_CTR_PAIR + part];
    if ((int)(blockIdx.x * blockDim.x) >= n) return;              // block-uniform
    const DeformerSmem ds = stage_deformer(sm, dm_g, fr.A, fr.bigA);
    float* sc = ds.scratch + threadIdx.x;
    const float frame_dim = fr.frame_dim[0];
    const float* pbw_part = fr.part_pbw + (long long)part * fr.maxlen * NVR_JOINTS;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const KnnRec rec = recs[(long long)part * cap + i];
        const float4 sv = surv[rec.surv];
        const float p[3] = {sv.x, sv.y, sv.z};
        const int sample = __float_as_int(sv.w);
        const long long di = (long long)(sample / dir_div) * 3;
        const float wd[3] = {dirs[di], dirs[di + 1], dirs[di + 2]};
        float d[3], x0[3], v[3], r[3];
        nvr_dir_to_pose(fr.R, wd, d);
        nvr_blend_lbs(rec.idx, rec.w, pbw_part, ds.A, ds.bigA, p, d, x0, v);
        nvr_deformer_point(dg, ds.pk, fr.tuv, frame_dim, x0, r, sc, WARP_THREADS);
        PairRec out;
        out.x = x0[0] + r[0]; out.y = x0[1] + r[1]; out.z = x0[2] + r[2];   // :113
        out.vx = v[0]; out.vy = v[1]; out.vz = v[2];
        out.surv = rec.surv; out._pad = 0;
        pairs[(long long)part * cap + i] = out;
        if (out_x0) {
            const long long o3 = ((long long)(out_rank ? out_rank[rec.surv] : rec.surv) * NVR_PARTS + part) * 3;
#pragma unroll
            for (int a = 0; a < 3; ++a) { out_x0[o3 + a] = x0[a]; out_resd[o3 + a] = r[a]; }
        }
        if (dbg) {
            float* dr = dbg + ((long long)sample * NVR_PARTS + part) * 8;
            dr[1] = out.x; dr[2] = out.y; dr[3] = out.z; dr[4] = v[0]; dr[5] = v[1]; dr[6] = v[2];
        }
    }
}

// Deformer on explicit canonical points (Network.resd).
__global__ void __launch_bounds__(WARP_THREADS)
k_deformer(FrameDev fr, GridDev dg, DeformerMlp dm_g, const float* __restrict__ x, long long n, float* __restrict__ out) {
    __shared__ __align__(16) float sm[WARP_SMEM_FLOATS];
    const DeformerSmem ds = stage_deformer(sm, dm_g, nullptr, nullptr);
    float* sc = ds.scratch + threadIdx.x;
    const float frame_dim = fr.frame_dim[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x0[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
        float r[3];
        nvr_deformer_point(dg, ds.pk, fr.tuv, frame_dim, x0, r, sc, WARP_THREADS);
        out[i * 3] = r[0]; out[i * 3 + 1] = r[1]; out[i * 3 + 2] = r[2];
    }
}

// -----------------------------------------------------------------------------------------
// THE gather: part grid embedding, one 32-byte sector per lane
// -----------------------------------------------------------------------------------------
// A warp works on 16 points at a time: lane = 2*point + half.  A grid entry is one 64-byte row of 16
// fp32 features = two 32-byte sectors; each lane fetches ONE whole sector with a single 256-bit load
// (LDG.E.256), so every byte of every sector that moves is used and the per-point index arithmetic is
// shared by only two lanes.  The 8 corner loads of a level are independent and issued back to back.
// The level loop is deliberately NOT unrolled: unrolled, the kernel is ~10k instructions and spends
// most of its time waiting for instruction fetch (profiles/r1a: 62 % stall_no_inst); rolled it is a
// few hundred instructions that stay in the instruction cache.
// Input points: float x[3] at `xbase + i * xstride` (PairRec lists: stride 8; plain xyz: stride 3).
// The point count is *count_dev when non-null (device-side list length), else n_imm.  One launch per part.
struct __align__(32) Sector { float v[8]; };
__device__ __forceinline__ Sector ld_sector(const float* p) {
    Sector r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}

// Clamped corner coordinates of one axis, reference semantics (part_base_embedder.py:115-118):
// `.long()` truncates toward zero; cvt.rzi.s32 saturates instead of wrapping, which is the same after
// the clamp to [0, res-1] (NaN -> 0 -> clamps to 0, as LLONG_MIN does).
__device__ __forceinline__ void axis_coord(float u, float size, int res, int& i0, int& i1, float& o) {
    const float f = u / size;                                     // IEEE fp32 divide
    i0 = min(max(__float2int_rz(f + 0.0f), 0), res - 1);
    i1 = min(max(__float2int_rz(f + 1.0f), 0), res - 1);
    o = f - (float)i0;
}

// The 8 corner rows of level l (x = bit 2, y = bit 1, z = bit 0 of the corner index) from the clamped corner coordinates;
// returns the table the rows index (dense or hash).  Shared by k_embed, k_embed_presum and k_embed_footprint.
__device__ __forceinline__ const float* level_rows(const GridDev& g, int l, int res, const int i0[3], const int i1[3],
                                                   bool fast_mod, unsigned int T32, unsigned int row[8]) {
    if (l < g.start_hash) {                                       // dense level  (:124-129)
        const unsigned int off = (unsigned int)g.dense_off[l];
        const unsigned int ax[2] = {(unsigned int)(i0[0] * res * res) + off, (unsigned int)(i1[0] * res * res) + off};
        const unsigned int ay[2] = {(unsigned int)(i0[1] * res), (unsigned int)(i1[1] * res)};
        const unsigned int az[2] = {(unsigned int)i0[2], (unsigned int)i1[2]};
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = ax[(c >> 2) & 1] + ay[(c >> 1) & 1] + az[c & 1];
        return g.dense;
    }
    // hashed level (:132-136), int64 products
    const unsigned int off = (unsigned int)(l - g.start_hash) * T32;
    const unsigned long long hx[2] = {(unsigned long long)i0[0], (unsigned long long)i1[0]};
    const unsigned long long hy[2] = {(unsigned long long)i0[1] * 19349663ull, (unsigned long long)i1[1] * 19349663ull};
    const unsigned long long hz[2] = {(unsigned long long)i0[2] * 83492791ull, (unsigned long long)i1[2] * 83492791ull};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const unsigned long long h = hx[(c >> 2) & 1] ^ hy[(c >> 1) & 1] ^ hz[c & 1];
        row[c] = (fast_mod ? nvr_mod_T40(h, T32, g.T_magic40) : (unsigned int)nvr_mod_T(h, g.T, g.T_magic)) + off;
    }
    return g.hash;
}

// DYN: the 16-pair units are handed out through a device counter (`work`) instead of a static grid stride, so a launch is
// balanced when a warp only gets two or three units (a shard of a frame on one of eight GPUs)
template <bool DYN>
__device__ __forceinline__ void embed_body(const GridDev& g, const float* __restrict__ xb, int xstride, int n,
                                           float* __restrict__ eb, int emb_stride, int l_begin, int l_end, int* work) {
    const int lane = threadIdx.x & 31, half = lane & 1;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    int pending = 0;                                   // DYN: lane 0's fetch of the unit after the current one, issued a unit ahead
    auto next_unit = [&](int prev) {
        if (!DYN) return prev + n_warps * 16;
        const int u = __shfl_sync(0xffffffffu, pending, 0);
        if (lane == 0) pending = atomicAdd(work, 1);
        return u * 16;
    };
    if (DYN && lane == 0) pending = atomicAdd(work, 1);
    const bool fast_mod = g.T_magic40 != 0;
    const unsigned int T32 = (unsigned int)g.T;
    for (int base = DYN ? next_unit(0) : warp * 16; base < n; base = next_unit(base)) {
        const int pt = base + (lane >> 1);
        const bool live = pt < n;
        const float* xp = xb + (long long)(live ? pt : n - 1) * xstride;
        const float x[3] = {xp[0], xp[1], xp[2]};
        float u[3];
        nvr_normalise(g, x, u);
        float* o = eb + (long long)pt * emb_stride;
        if (live && half == 0 && l_begin == 0) { o[0] = u[0]; o[1] = u[1]; o[2] = u[2]; }
#pragma unroll 1
        for (int l = l_begin; l < l_end; ++l) {
            const int res = g.res[l];
            const float size = g.size[l];
            int i0[3], i1[3];
            float of[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) axis_coord(u[a], size, res, i0[a], i1[a], of[a]);
            unsigned int row[8];
            const float* tab = level_rows(g, l, res, i0, i1, fast_mod, T32, row);
            Sector v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = ld_sector(tab + (unsigned long long)row[c] * 16 + half * 8);
            const float wx[2] = {1.0f - of[0], of[0]}, wy[2] = {1.0f - of[1], of[1]}, wz[2] = {1.0f - of[2], of[2]};
            F2 acc[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[f] = F2{0.0f, 0.0f};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = (wx[(c >> 2) & 1] * wy[(c >> 1) & 1]) * wz[c & 1];      // :158-159
                const F2 w2 = {w, w};
#pragma unroll
                for (int f = 0; f < 4; ++f) nvr_fma2(acc[f], w2, F2{v[c].v[2 * f], v[c].v[2 * f + 1]});   // :160, two features per FFMA2
            }
            float sfeat = ((acc[0].x + acc[0].y) + (acc[1].x + acc[1].y)) + ((acc[2].x + acc[2].y) + (acc[3].x + acc[3].y));
            sfeat += __shfl_xor_sync(0xffffffffu, sfeat, 1);                            // :165 sum over the 16 features
            if (live && half == (l & 1)) o[3 + l] = sfeat;
        }
    }
}

__global__ void __launch_bounds__(256, 2)
k_embed(GridDev g, const float* __restrict__ xb, int xstride, const int* __restrict__ count_dev, int n_imm,
        float* __restrict__ eb, int emb_stride, int l_begin, int l_end) {
    // levels [l_begin, l_end) only (the normalised coordinates are written by the launch with l_begin == 0): the level-major
    // experiment gathers a part one L2-sized slice of its tables per launch (nvr_cabi.cu, embed_plan)
    embed_body<false>(g, xb, xstride, count_dev ? *count_dev : n_imm, eb, emb_stride, l_begin, l_end, nullptr);
}

// All five parts of a pass as ONE grid: blockIdx.y = part.  Same per-CTA work as five k_embed launches (a CTA gathers one part's
// pairs, so consecutive units still share coarse-level rows in L1), but the launches' tails overlap: CTAs of part p + 1 start as
// the last CTAs of part p drain.  The part's grid description is staged in shared memory (a run-time index into a kernel
// parameter array would be copied to the stack).
struct EmbedBatch { const float* x[NVR_PARTS]; const int* count[NVR_PARTS]; float* out[NVR_PARTS]; int* work[NVR_PARTS]; };
__global__ void __launch_bounds__(256, 2)
k_embed_parts(const GridDev* __restrict__ grids, EmbedBatch b, int xstride, int emb_stride) {
    __shared__ GridDev sg;
    __shared__ const float* s_x;
    __shared__ const int* s_count;
    __shared__ float* s_out;
    __shared__ int* s_work;
    const int part = blockIdx.y;
    {
        const int* src = reinterpret_cast<const int*>(grids + part);
        int* dst = reinterpret_cast<int*>(&sg);
        for (int i = threadIdx.x; i < (int)(sizeof(GridDev) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) {
#pragma unroll
            for (int p = 0; p < NVR_PARTS; ++p)
                if (p == part) { s_x = b.x[p]; s_count = b.count[p]; s_out = b.out[p]; s_work = b.work[p]; }
        }
    }
    __syncthreads();
    embed_body<true>(sg, s_x, xstride, *s_count, s_out, emb_stride, 0, sg.n_levels, s_work);
}

// -----------------------------------------------------------------------------------------
// inference tables: the part networks only ever consume sum_f of an entry (part_base_embedder.py:165), so for
// forward-only rendering the 16 features of every row can be summed ONCE per weight update: 4 B per corner
// instead of 64 B (16x less gather traffic; SURVEY.md section 7 "legal algorithmic win").  Opt-in
// (nvr_prepare_inference); training and the roofline-defined run use the full tables.
// -----------------------------------------------------------------------------------------
__global__ void k_presum_rows(const float* __restrict__ tab, long long n_rows, float* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        const float4* t4 = reinterpret_cast<const float4*>(tab + r * 16);
        const float4 a = __ldg(t4), b = __ldg(t4 + 1), c = __ldg(t4 + 2), d = __ldg(t4 + 3);
        out[r] = (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) + (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
    }
}

// Same contract as k_embed, on the pre-summed tables (sd = dense sums, sh = hash sums): one lane = one point.
__global__ void __launch_bounds__(256)
k_embed_presum(GridDev g, const float* __restrict__ sd, const float* __restrict__ sh, const float* __restrict__ xb, int xstride,
               const int* __restrict__ count_dev, int n_imm, float* __restrict__ eb, int emb_stride) {
    const int n = count_dev ? *count_dev : n_imm;
    const bool fast_mod = g.T_magic40 != 0;
    const unsigned int T32 = (unsigned int)g.T;
    for (int pt = blockIdx.x * blockDim.x + threadIdx.x; pt < n; pt += gridDim.x * blockDim.x) {
        const float* xp = xb + (long long)pt * xstride;
        const float x[3] = {xp[0], xp[1], xp[2]};
        float u[3];
        nvr_normalise(g, x, u);
        float* o = eb + (long long)pt * emb_stride;
        o[0] = u[0]; o[1] = u[1]; o[2] = u[2];
#pragma unroll 1
        for (int l = 0; l < g.n_levels; ++l) {
            const int res = g.res[l];
            int i0[3], i1[3];
            float of[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) axis_coord(u[a], g.size[l], res, i0[a], i1[a], of[a]);
            unsigned int row[8];
            const float* tab = level_rows(g, l, res, i0, i1, fast_mod, T32, row) == g.dense ? sd : sh;
            float v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + row[c]);
            const float wx[2] = {1.0f - of[0], of[0]}, wy[2] = {1.0f - of[1], of[1]}, wz[2] = {1.0f - of[2], of[2]};
            float lev = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) lev += ((wx[(c >> 2) & 1] * wy[(c >> 1) & 1]) * wz[c & 1]) * v[c];
            o[3 + l] = lev;
        }
    }
}

// -----------------------------------------------------------------------------------------
// gather footprint (measurement aid, nvr_gather_footprint): the DISTINCT 32-byte sectors one part's pair list touches.
// Same walk and index arithmetic as k_embed (lane = 2 * point + half, one sector per lane and corner); instead of
// loading the sector the lane sets its bit in a bitmap over [dense sectors | hash sectors].
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_embed_footprint(GridDev g, const float* __restrict__ xb, int xstride, const int* __restrict__ count_dev,
                  unsigned int* __restrict__ bitmap, unsigned long long dense_sectors) {
    const int n = *count_dev;
    const int lane = threadIdx.x & 31, half = lane & 1;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool fast_mod = g.T_magic40 != 0;
    const unsigned int T32 = (unsigned int)g.T;
    for (int base = warp * 16; base < n; base += n_warps * 16) {
        const int pt = base + (lane >> 1);
        if (pt >= n) continue;
        const float* xp = xb + (long long)pt * xstride;
        const float x[3] = {xp[0], xp[1], xp[2]};
        float u[3];
        nvr_normalise(g, x, u);
#pragma unroll 1
        for (int l = 0; l < g.n_levels; ++l) {
            const int res = g.res[l];
            int i0[3], i1[3];
            float of[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) axis_coord(u[a], g.size[l], res, i0[a], i1[a], of[a]);
            unsigned int row[8];
            const bool dense = level_rows(g, l, res, i0, i1, fast_mod, T32, row) == g.dense;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned long long sec = (dense ? 0ull : dense_sectors) + (unsigned long long)row[c] * 2 + half;
                const unsigned int bit = 1u << (sec & 31);
                unsigned int* w = bitmap + (sec >> 5);
                if (!(*(volatile unsigned int*)w & bit)) atomicOr(w, bit);
            }
        }
    }
}

// Two-lane render: a pass's counters added into the call's totals (both lanes, any order).
__global__ void k_add_counters(const int* __restrict__ src, int* __restrict__ dst) {
    if (threadIdx.x < NVR_CTR_WORDS && src[threadIdx.x]) atomicAdd(dst + threadIdx.x, src[threadIdx.x]);
}

__global__ void k_popcount_words(const unsigned int* __restrict__ words, long long n_words, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (long long)gridDim.x * blockDim.x)
        acc += __popc(words[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
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
    extern __shared__ __align__(128) float sm[];
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
__device__ __forceinline__ float4 fuse_parts(const float4* __restrict__ raws, int slot, const float4* __restrict__ far_raws) {
    // inb_part_network_multiassign.py:253-255: raw of the part with the largest occupancy (first on ties).
    // occ = -1 marks a pair answered by the part's shared far-field evaluation far_raws[p] (NVR_FAR_WSUM).
    float4 best = raws[(long long)slot * NVR_PARTS];
    if (best.w < 0.0f) best = far_raws[0];
#pragma unroll
    for (int p = 1; p < NVR_PARTS; ++p) {
        float4 r = raws[(long long)slot * NVR_PARTS + p];
        if (r.w < 0.0f) r = far_raws[p];
        if (r.w > best.w) best = r;
    }
    return best;
}

__global__ void k_resolve_points(const int* __restrict__ surv_of_sample, const float4* __restrict__ raws, const float4* __restrict__ far_raws,
                                 long long n, float4* __restrict__ raw_out, float* __restrict__ occ_out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int slot = surv_of_sample[i];
        const float4 r = slot >= 0 ? fuse_parts(raws, slot, far_raws) : make_float4(0.f, 0.f, 0.f, 0.f);
        raw_out[i] = r;
        if (occ_out) occ_out[i] = r.w;
    }
}

// One warp per ray; samples are walked 32 at a time with a shuffle product-scan of (1 - alpha)
// (net_utils.py:12-15 with epsilon = 0; :39-41).
// fo.world > 0: the composited ray also goes to its final position in every rank's frame slot (nvr_frame.cuh), lanes 0..world-1
// storing to one rank each; ray_base = index of this pass's first ray inside the rank's shard.  rgb_map / acc_map may then be null.
__global__ void __launch_bounds__(256)
k_resolve_rays(const int* __restrict__ surv_of_sample, const float4* __restrict__ raws, const float4* __restrict__ far_raws,
               long long n_rays, int S, float* __restrict__ rgb_map, float* __restrict__ acc_map, float4* __restrict__ raw_out,
               FrameOut fo, long long ray_base) {
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
                if (slot >= 0) r = fuse_parts(raws, slot, far_raws);
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
        if (lane == 0 && rgb_map) {
            rgb_map[ray * 3] = cr; rgb_map[ray * 3 + 1] = cg; rgb_map[ray * 3 + 2] = cb;
            acc_map[ray] = ca;
        }
        if (lane < fo.world) {
            const long long gi = frame_index(fo, ray_base + ray);
            float4* dst = nullptr;                                  // static indices only: a dynamic one would copy the
#pragma unroll                                                      // parameter array to the stack
            for (int r = 0; r < NVR_MAX_RANKS; ++r) dst = lane == r ? fo.slot[r] : dst;
            if (gi < fo.n_total) dst[gi] = make_float4(cr, cg, cb, ca);
        }
    }
}

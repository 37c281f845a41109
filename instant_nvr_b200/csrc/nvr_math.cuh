// nvr_math.cuh -- per-sample arithmetic of the hot path, shared by every kernel.
//
// Everything here is `__host__ __device__` so the exact same statements can be compiled with g++
// by tests/host_emul (a TEST-ONLY harness that checks this arithmetic against the oracle without
// a GPU).  The product never runs it on the CPU: the kernels in nvr_kernels.cu are the only callers
// in libnvr_b200.so.
//
// Reference citations are to zju3dv/instant-nvr @ a6f4d68.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define NVR_HD __host__ __device__ __forceinline__
#else
#define NVR_HD inline
struct float4 { float x, y, z, w; };      // host emulation only
#endif

#define NVR_LEVELS 16
#define NVR_PARTS 5
#define NVR_JOINTS 24
#define NVR_KNN 4

// Two fp32 fused multiply-adds in one instruction (Blackwell FFMA2, PTX fma.rn.f32x2): c = a * b + c per half,
// each half rounded once like fmaf.  Halves the issue slots of the FMA-heavy inner loops; the host build (tests)
// evaluates the same two products separately.
struct F2 { float x, y; };
NVR_HD void nvr_fma2(F2& c, const F2& a, const F2& b) {
#ifdef __CUDA_ARCH__
    unsigned long long ua, ub, uc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(uc) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(uc));
#else
    c.x = a.x * b.x + c.x;
    c.y = a.y * b.y + c.y;
#endif
}
NVR_HD F2 nvr_sub2(const F2& a, const F2& b) {                    // FADD2
#ifdef __CUDA_ARCH__
    unsigned long long ua, ub, ur;
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(ur) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ur));
    return r;
#else
    return F2{a.x - b.x, a.y - b.y};
#endif
}
NVR_HD F2 nvr_mul2(const F2& a, const F2& b) {                    // FMUL2
#ifdef __CUDA_ARCH__
    unsigned long long ua, ub, ur;
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ur) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ur));
    return r;
#else
    return F2{a.x * b.x, a.y * b.y};
#endif
}
// warp vote on the device (callers are warp-converged), identity in the scalar host build
#ifdef __CUDA_ARCH__
#define NVR_ANY(x) __any_sync(0xffffffffu, (x))
#define NVR_ANY_ACTIVE(x) __any_sync(__activemask(), (x))   // for loops whose tail leaves a warp partially active
#else
#define NVR_ANY(x) (x)
#define NVR_ANY_ACTIVE(x) (x)
#endif

// Device-side view of one grid (mirrors NvrGrid, plus the Barrett constant for `% T`).
struct GridDev {
    const float* dense;
    const float* hash;
    const float* bounds;
    int n_levels, n_feat, start_hash, sum_features;
    unsigned long long T;
    unsigned long long T_magic;   // floor(2^64 / T)
    unsigned int T_magic40;       // floor(2^40 / T); 0 when the 32-bit reduction is not applicable (see nvr_mod_T40)
    int res[NVR_LEVELS];
    float size[NVR_LEVELS];
    long long dense_off[NVR_LEVELS];
};

struct VolumeDev {                // a (D,H,W,C) fp32 volume indexed by the point's (x,y,z)
    const float* data;
    int D, H, W, C;
    const float* bounds;          // (2,3) device
};

NVR_HD float nvr_softplus(float x) {          // torch.nn.Softplus(beta=1, threshold=20)
    return x > 20.0f ? x : log1pf(expf(x));
}
NVR_HD float nvr_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
// Softplus of HIDDEN activations on the device: max(x,0) + log1p(exp(-|x|)) on the MUFU units (ex2 / lg2),
// absolute error ~1e-7 (the next layer only sees it through a dot product), ~5x fewer instructions than the
// precise form, and equal to x above torch's threshold of 20.  Output heads (occupancy) keep nvr_softplus.
NVR_HD float nvr_softplus_hidden(float x) {
#ifdef __CUDA_ARCH__
    // ex2.approx / lg2.approx with .ftz: ONE MUFU each (the __expf / __logf intrinsics wrap them in denormal scaling code,
    // 9 instructions per activation against 6 here).  t = exp(-|x|) in (0, 1]: flushing a denormal t to 0 changes nothing below
    // 1e-38, and 1 + t in [1, 2] is never denormal.
    float t, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(x) * -1.4426950408889634f));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + t));
    return fmaf(l, 0.6931471805599453f, fmaxf(x, 0.0f));
#else
    return nvr_softplus(x);
#endif
}

// ---------------------------------------------------------------------------------------
// hash-grid index arithmetic                          part_base_embedder.py:112-136, 158-159
// ---------------------------------------------------------------------------------------
NVR_HD unsigned long long nvr_umulhi64(unsigned long long a, unsigned long long b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#endif
}

// h mod T for h < 2^63 via Barrett reduction (T is one of four primes, never a power of two)
NVR_HD unsigned long long nvr_mod_T(unsigned long long h, unsigned long long T, unsigned long long magic) {
    unsigned long long q = nvr_umulhi64(h, magic);
    unsigned long long r = h - q * T;
    return r >= T ? r - T : r;
}

// h mod T for h < 2^40 and 2^8 < T < 2^31 with 32-bit arithmetic: q = umulhi(h >> 8, floor(2^40 / T))
// under-estimates floor(h / T) by at most 1 (h / 2^40 + 2^8 / T < 1), so two conditional subtractions
// are more than enough.  The part grids satisfy the bounds: coordinates < 2^13 make the int64 hash
// (x*1 ^ y*19349663 ^ z*83492791) < 2^40.  tests/test_host_emul.py::test_barrett_mod checks it against %.
NVR_HD unsigned int nvr_umulhi32(unsigned int a, unsigned int b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (unsigned int)(((unsigned long long)a * b) >> 32);
#endif
}
NVR_HD unsigned int nvr_mod_T40(unsigned long long h, unsigned int T, unsigned int magic40) {
    const unsigned int q = nvr_umulhi32((unsigned int)(h >> 8), magic40);
    unsigned int r = (unsigned int)h - q * T;                    // exact modulo 2^32, true value < 3T
    unsigned int t = r - T;
    r = t < r ? t : r;                                           // r >= T  <=>  r - T does not wrap
    t = r - T;
    r = t < r ? t : r;
    return r;
}

struct LevelCoord {
    int i0[3], i1[3];   // clamped integer corner coordinates for offset 0 / 1 along each axis
    float o[3];         // f - float(i0): trilinear offset measured from the CLAMPED corner (:118)
};

// u: normalised coordinate (xyz - b0) / (b1 - b0), NOT clamped (:112)
NVR_HD void nvr_level_coord(const float u[3], float size, int res, LevelCoord& lc) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float f = u[a] / size;                               // :115  (IEEE fp32 divide)
        long long t0 = (long long)(f + 0.0f);                // :116  .long() truncates toward zero
        long long t1 = (long long)(f + 1.0f);
        long long hi = (long long)res - 1;
        t0 = t0 < 0 ? 0 : (t0 > hi ? hi : t0);               // :117
        t1 = t1 < 0 ? 0 : (t1 > hi ? hi : t1);
        lc.i0[a] = (int)t0;
        lc.i1[a] = (int)t1;
        lc.o[a] = f - (float)t0;                             // :118
    }
}

// corner c: x = bit 2, y = bit 1, z = bit 0 (offsets table, :81-88)
NVR_HD float nvr_corner_weight(const LevelCoord& lc, int c) {
    float wx = (c & 4) ? lc.o[0] : 1.0f - lc.o[0];           // (1-off) + (2*off-1)*o  (:158)
    float wy = (c & 2) ? lc.o[1] : 1.0f - lc.o[1];
    float wz = (c & 1) ? lc.o[2] : 1.0f - lc.o[2];
    return (wx * wy) * wz;                                   // :159
}

// Row of corner c of level l inside `dense` (l < start_hash) or `hash` flattened to (H*T, F).
NVR_HD long long nvr_corner_row(const GridDev& g, int l, const LevelCoord& lc, int c) {
    long long ix = (c & 4) ? lc.i1[0] : lc.i0[0];
    long long iy = (c & 2) ? lc.i1[1] : lc.i0[1];
    long long iz = (c & 1) ? lc.i1[2] : lc.i0[2];
    if (l < g.start_hash) {
        long long r = g.res[l];
        return ix * r * r + iy * r + iz + g.dense_off[l];    // :124-129
    }
    unsigned long long h = ((unsigned long long)ix * 1ull) ^ ((unsigned long long)iy * 19349663ull) ^
                           ((unsigned long long)iz * 83492791ull);          // :132-135 (int64, no wrap)
    return (long long)(nvr_mod_T(h, g.T, g.T_magic) + (unsigned long long)(l - g.start_hash) * g.T);   // :136
}

NVR_HD const float* nvr_level_table(const GridDev& g, int l) { return l < g.start_hash ? g.dense : g.hash; }

NVR_HD void nvr_normalise(const GridDev& g, const float x[3], float u[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) u[a] = (x[a] - g.bounds[a]) / (g.bounds[3 + a] - g.bounds[a]);   // :112
}

// Scalar (one thread = one point) embedding with F features per entry.  Used by the deformer (F=2,
// concat) and by the host emulation; the part grids' production path is the quad-lane gather kernel,
// which shares nvr_level_coord / nvr_corner_row / nvr_corner_weight with this function.
//   sum_features: out[3 + l] = sum_f sum_c w_c * t[row_c][f];  concat: out[3 + l*F + f]
template <int F>
NVR_HD void nvr_embed_point(const GridDev& g, const float x[3], float* out, int os = 1) {   // os: output element stride
    float u[3];
    nvr_normalise(g, x, u);
    out[0] = u[0]; out[os] = u[1]; out[2 * os] = u[2];
    for (int l = 0; l < g.n_levels; ++l) {
        LevelCoord lc;
        nvr_level_coord(u, g.size[l], g.res[l], lc);
        const float* tab = nvr_level_table(g, l);
        float acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float w = nvr_corner_weight(lc, c);
            const float* row = tab + nvr_corner_row(g, l, lc, c) * F;
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] += w * row[f];                   // :160
        }
        if (g.sum_features) {
            float s = 0.0f;
#pragma unroll
            for (int f = 0; f < F; ++f) s += acc[f];                            // :165
            out[(3 + l) * os] = s;
        } else {
#pragma unroll
            for (int f = 0; f < F; ++f) out[(3 + l * F + f) * os] = acc[f];     // :169
        }
    }
}


#ifdef __CUDACC__
// Device form of nvr_embed_point<2> in concat mode (the deformer grid: 8 levels x 2 features, 19 outputs): the same values
// in the same order -- clamped corners, weights, accumulation over the corners c = 0..7 -- with 32-bit index arithmetic
// (cvt.rzi + clamp == .long() + clamp, products < 2^40 reduced by nvr_mod_T40) and the 8 corner rows of a level fetched as
// independent 8-byte loads before the first multiply (ncu r2a: the scalar form spent 17 % of k_warp's stall samples waiting
// on one dependent row load after another and 14 % of its instructions on 64-bit row arithmetic).
template <class Emit>
__device__ __forceinline__ void nvr_embed_point_f2_emit(const GridDev& g, const float x[3], Emit emit) {   // emit(k, value), k < 19
    float u[3];
    nvr_normalise(g, x, u);
    emit(0, u[0]); emit(1, u[1]); emit(2, u[2]);
    const bool fast_mod = g.T_magic40 != 0;
    const unsigned int T32 = (unsigned int)g.T;
#pragma unroll 1
    for (int l = 0; l < g.n_levels; ++l) {
        const int res = g.res[l];
        const float size = g.size[l];
        int i0[3], i1[3];
        float o[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float f = u[a] / size;                             // :115 (IEEE fp32 divide)
            i0[a] = min(max(__float2int_rz(f + 0.0f), 0), res - 1);  // :116-117
            i1[a] = min(max(__float2int_rz(f + 1.0f), 0), res - 1);
            o[a] = f - (float)i0[a];                                 // :118
        }
        unsigned int row[8];
        const float2* tab;
        if (l < g.start_hash) {
            tab = reinterpret_cast<const float2*>(g.dense);
            const unsigned int off = (unsigned int)g.dense_off[l];
            const unsigned int ax[2] = {(unsigned int)(i0[0] * res * res) + off, (unsigned int)(i1[0] * res * res) + off};
            const unsigned int ay[2] = {(unsigned int)(i0[1] * res), (unsigned int)(i1[1] * res)};
            const unsigned int az[2] = {(unsigned int)i0[2], (unsigned int)i1[2]};
#pragma unroll
            for (int c = 0; c < 8; ++c) row[c] = ax[(c >> 2) & 1] + ay[(c >> 1) & 1] + az[c & 1];
        } else {
            tab = reinterpret_cast<const float2*>(g.hash);
            const unsigned int off = (unsigned int)(l - g.start_hash) * T32;
            const unsigned long long hx[2] = {(unsigned long long)i0[0], (unsigned long long)i1[0]};
            const unsigned long long hy[2] = {(unsigned long long)i0[1] * 19349663ull, (unsigned long long)i1[1] * 19349663ull};
            const unsigned long long hz[2] = {(unsigned long long)i0[2] * 83492791ull, (unsigned long long)i1[2] * 83492791ull};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned long long h = hx[(c >> 2) & 1] ^ hy[(c >> 1) & 1] ^ hz[c & 1];
                row[c] = (fast_mod ? nvr_mod_T40(h, T32, g.T_magic40) : (unsigned int)nvr_mod_T(h, g.T, g.T_magic)) + off;
            }
        }
        float2 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + row[c]);
        const float wx[2] = {1.0f - o[0], o[0]}, wy[2] = {1.0f - o[1], o[1]}, wz[2] = {1.0f - o[2], o[2]};
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float w = (wx[(c >> 2) & 1] * wy[(c >> 1) & 1]) * wz[c & 1];   // :158-159
            a0 += w * v[c].x; a1 += w * v[c].y;                                    // :160
        }
        emit(3 + l * 2, a0); emit(3 + l * 2 + 1, a1);                            // :169
    }
}
__device__ __forceinline__ void nvr_embed_point_f2_dev(const GridDev& g, const float x[3], float* out, int os) {
    nvr_embed_point_f2_emit(g, x, [&](int k, float v) { out[k * os] = v; });
}
#endif

// ---------------------------------------------------------------------------------------
// trilinear volume lookup == F.grid_sample(bilinear, border, align_corners=True)
//                                                     blend_utils.py:501-525, 528-555
// ---------------------------------------------------------------------------------------
// channels [ch0, ch0+nch) of the C-channel volume; point axis x->D, y->H, z->W.
// clipped continuous voxel coordinates of p along the volume's (D, H, W) axes
NVR_HD void nvr_volume_coords(const VolumeDev& v, const float p[3], float c[3]) {
    const int dims[3] = {v.D, v.H, v.W};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float g = (p[a] - v.bounds[a]) / (v.bounds[3 + a] - v.bounds[a]);
        g = g * 2.0f - 1.0f;
        float t = ((g + 1.0f) / 2.0f) * (float)(dims[a] - 1);     // grid_sampler_unnormalize, align_corners
        t = fminf(fmaxf(t, 0.0f), (float)(dims[a] - 1));          // clip_coordinates (border)
        c[a] = t;
    }
}
NVR_HD void nvr_sample_volume_at(const VolumeDev& v, const float c[3], int ch0, int nch, float* out);
NVR_HD void nvr_sample_volume(const VolumeDev& v, const float p[3], int ch0, int nch, float* out) {
    float c[3];
    nvr_volume_coords(v, p, c);
    nvr_sample_volume_at(v, c, ch0, nch, out);
}
NVR_HD void nvr_sample_volume_at(const VolumeDev& v, const float c[3], int ch0, int nch, float* out) {
    // ATen names: x <-> W (our z), y <-> H (our y), z <-> D (our x)
    const float ix = c[2], iy = c[1], iz = c[0];
    const float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
    const float x1 = x0 + 1.0f, y1 = y0 + 1.0f, z1 = z0 + 1.0f;
    for (int k = 0; k < nch; ++k) out[k] = 0.0f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {                       // tnw, tne, tsw, tse, bnw, bne, bsw, bse
        const bool hx = n & 1, hy = n & 2, hz = n & 4;
        const float wx = hx ? (ix - x0) : (x1 - ix);
        const float wy = hy ? (iy - y0) : (y1 - iy);
        const float wz = hz ? (iz - z0) : (z1 - iz);
        const float w = (wx * wy) * wz;
        const int cx = (int)(hx ? x1 : x0), cy = (int)(hy ? y1 : y0), cz = (int)(hz ? z1 : z0);
        if (cx <= v.W - 1 && cy <= v.H - 1 && cz <= v.D - 1) {    // within_bounds_3d (>= 0 holds after the clip)
            const float* src = v.data + (((long long)cz * v.H + cy) * v.W + cx) * v.C + ch0;
            for (int k = 0; k < nch; ++k) out[k] += src[k] * w;
        }
    }
}

// Conservative early-out for the distance cull.  A trilinear lookup is a convex combination of the 8 voxels around
// (floor(c), floor(c) + 1), so it cannot fall below their minimum.  cmin holds, per coarse cell of NVR_CULL_B^3 fine
// cells, the minimum of the distance volume over fine indices [cB, cB + B] per axis (one past the cell, for the +1
// corners).  If that minimum exceeds the threshold by more than the rounding of the 8-term sum, the sample is culled
// without touching the fine volume; everything else takes the exact path, so the survivor set is unchanged.
#define NVR_CULL_B 4
#define NVR_CULL_MARGIN 1.00001f
NVR_HD int nvr_coarse_dim(int d) { return (d + NVR_CULL_B - 1) / NVR_CULL_B; }
NVR_HD bool nvr_cull_early_out(const VolumeDev& v, const float* cmin, const float c[3], float thresh) {
    const int cz = (int)floorf(c[0]) / NVR_CULL_B, cy = (int)floorf(c[1]) / NVR_CULL_B, cx = (int)floorf(c[2]) / NVR_CULL_B;
    const float m = cmin[((long long)cz * nvr_coarse_dim(v.H) + cy) * nvr_coarse_dim(v.W) + cx];
    return m > thresh * NVR_CULL_MARGIN;
}
// minimum of a 1-channel volume over the fine indices a coarse cell's lookups can touch
// (one extra voxel on every side, so that a cell index taken from coordinates that are off by less than a voxel --
// nvr_cull_quick -- is covered too)
NVR_HD float nvr_coarse_min(const float* dist, int D, int H, int W, int cz, int cy, int cx) {
    float m = INFINITY;
    for (int z = (cz * NVR_CULL_B > 0 ? cz * NVR_CULL_B - 1 : 0); z <= cz * NVR_CULL_B + NVR_CULL_B + 1 && z < D; ++z)
        for (int y = (cy * NVR_CULL_B > 0 ? cy * NVR_CULL_B - 1 : 0); y <= cy * NVR_CULL_B + NVR_CULL_B + 1 && y < H; ++y)
            for (int x = (cx * NVR_CULL_B > 0 ? cx * NVR_CULL_B - 1 : 0); x <= cx * NVR_CULL_B + NVR_CULL_B + 1 && x < W; ++x) {
                const float d = dist[((long long)z * H + y) * W + x];
                m = (d < m || d != d) ? d : m;                  // a NaN voxel poisons the cell: never early-out on it
            }
    return m;
}

// Quick conservative cull straight from WORLD coordinates: the chain world -> pose -> normalised -> voxel coordinates
// is affine, c_a = sum_b w_b M[b][a] + t_a, so one 3x4 map (12 FMAs, no divisions) gives the voxel coordinates to
// ~1e-4 voxel -- not the reference's bits, but the coarse-minimum grid carries a one-voxel margin, so a sample this
// test culls is one the exact lookup would cull as well; everything else takes the exact path.  Along a ray the map
// collapses further to c = A + z B with A = o.M + t, B = d.M.
struct CullQuick { float M[9]; float t[3]; float cmax[3]; };
NVR_HD void nvr_cull_quick_setup(const VolumeDev& v, const float* R, const float* Th, CullQuick& q) {
    const int dims[3] = {v.D, v.H, v.W};
    for (int a = 0; a < 3; ++a) {
        const float s = (float)(dims[a] - 1) / (v.bounds[3 + a] - v.bounds[a]);
        for (int b = 0; b < 3; ++b) q.M[b * 3 + a] = R[b * 3 + a] * s;
        q.t[a] = -(((Th[0] * R[0 * 3 + a] + Th[1] * R[1 * 3 + a]) + Th[2] * R[2 * 3 + a]) + v.bounds[a]) * s;
        q.cmax[a] = (float)(dims[a] - 1);
    }
}
NVR_HD void nvr_cull_quick_ray(const CullQuick& q, const float o[3], const float d[3], float A[3], float B[3]) {
    for (int a = 0; a < 3; ++a) {
        A[a] = ((o[0] * q.M[0 * 3 + a] + o[1] * q.M[1 * 3 + a]) + o[2] * q.M[2 * 3 + a]) + q.t[a];
        B[a] = (d[0] * q.M[0 * 3 + a] + d[1] * q.M[1 * 3 + a]) + d[2] * q.M[2 * 3 + a];
    }
}
// c: approximate voxel coordinates (unclamped).  true = certainly culled.  NaN coordinates clamp to 0 here; the exact
// path culls them too (a NaN lookup is never < thresh).
NVR_HD bool nvr_cull_quick(const VolumeDev& v, const CullQuick& q, const float* cmin, const float c[3], float thresh) {
    const int cz = (int)fminf(fmaxf(c[0], 0.0f), q.cmax[0]) / NVR_CULL_B;
    const int cy = (int)fminf(fmaxf(c[1], 0.0f), q.cmax[1]) / NVR_CULL_B;
    const int cx = (int)fminf(fmaxf(c[2], 0.0f), q.cmax[2]) / NVR_CULL_B;
    const float m = cmin[(cz * nvr_coarse_dim(v.H) + cy) * nvr_coarse_dim(v.W) + cx];
    return m > thresh * NVR_CULL_MARGIN;
}

// k_cull's walk over the samples of a pass (csrc/nvr_kernels.cuh): position inside the walk -> (ray, step, sample id).
// Rays are taken in groups of 32 and a group is walked in CHUNKS of 32 positions = 16 rays x 2 depth steps, so 32 consecutive
// positions (one warp iteration, and later one KNN unit / one warp of every other kernel) are a patch ~6 cm x 3 cm instead of
// a 13 cm x 1.5 cm strip of 32 rays at one depth.  Chunk c of a group: half = (c >> 2) & 1 (rays 0-15 or 16-31 of the group),
// depth pair = ((c >> 3) << 2) | (c & 3) -- i.e. a warp's 8 chunks are 8 depth steps of one half, then the same 8 steps of the
// other half, so a lane changes ray once per span.  The pairs are padded to a multiple of 4 (steps >= S are invalid
// positions), which makes chunk -> (half, pair) a bijection; for S = 64 / 128 / 256 nothing is padded.
// g0 / w0: group index and offset of the CTA's first position (one 64-bit division per 2048 positions); everything per
// position is 32-bit.
struct CullWalk { long long g0; unsigned w0, group; int S; long long n_rays; };
NVR_HD unsigned cull_group_positions(int S) { return 64u * ((((unsigned)S + 1u) / 2u + 3u) & ~3u); }
NVR_HD bool cull_locate(const CullWalk& cw, int local, long long& r, int& k, long long& i) {
    const unsigned wl = cw.w0 + (unsigned)local;
    const unsigned q = wl / cw.group, w = wl - q * cw.group;
    const unsigned c = w >> 5, sub = w & 31u;
    k = (int)(((((c >> 3) << 2) | (c & 3u)) << 1) | (sub >> 4));
    r = (cw.g0 + q) * 32 + (((c >> 2) & 1u) << 4) + (sub & 15u);
    i = r * cw.S + k;
    return r < cw.n_rays && k < cw.S;
}


// ---------------------------------------------------------------------------------------
// K=4 nearest vertices -> Gaussian blend weights                 blend_utils.py:732-763
// ---------------------------------------------------------------------------------------
// The running 4 best neighbours, ascending.  A candidate is the 64-bit key (bits(d2) << 32 | vertex
// index): squared distances are non-negative floats, whose bit patterns order like their values, so
// one unsigned compare orders candidates by (d2, vertex index).  The result therefore does not depend
// on the order in which vertices are visited and equals a stable top-k over the original vertex order
// (ties keep the lower index) -- which is what lets the production scan walk spatial clusters.
// NaN distances (bits above +inf) never enter.
struct Knn4 {
    unsigned long long key[NVR_KNN];
};
NVR_HD unsigned int nvr_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    unsigned int u; memcpy(&u, &f, 4); return u;
#endif
}
NVR_HD float nvr_u2f(unsigned int u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
NVR_HD float nvr_knn_d2(const Knn4& k, int i) { return nvr_u2f((unsigned int)(k.key[i] >> 32)); }
NVR_HD int nvr_knn_idx(const Knn4& k, int i) { return (int)(unsigned int)k.key[i]; }
NVR_HD void nvr_knn_init(Knn4& k) {
#pragma unroll
    for (int i = 0; i < NVR_KNN; ++i) k.key[i] = 0x7f800000ull << 32;          // (+inf, vertex 0)
}
// branch-free sorted insertion: bubble the candidate through the 4 slots
NVR_HD void nvr_knn_insert(Knn4& k, float d2, int j) {
    unsigned long long nk = ((unsigned long long)nvr_f2u(d2) << 32) | (unsigned int)j;
#pragma unroll
    for (int i = 0; i < NVR_KNN; ++i) {
        const unsigned long long lo = k.key[i] < nk ? k.key[i] : nk, hi = k.key[i] < nk ? nk : k.key[i];
        k.key[i] = lo;
        nk = hi;
    }
}
// could (d2, any index) still enter?  (d2 <= current 4th-best distance; false for NaN)
NVR_HD bool nvr_knn_admits(const Knn4& k, float d2) { return nvr_f2u(d2) <= (unsigned int)(k.key[NVR_KNN - 1] >> 32); }

NVR_HD float nvr_dist2(const float p[3], const float4& v) {
    const float dx = p[0] - v.x, dy = p[1] - v.y, dz = p[2] - v.z;
    return dx * dx + dy * dy + dz * dz;                           // squared L2, as knn_points returns it
}

NVR_HD float4 nvr_ld_vert(const float4* v) {
#ifdef __CUDA_ARCH__
    return __ldg(v);
#else
    return *v;
#endif
}
NVR_HD int nvr_f2i(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int id; memcpy(&id, &f, 4); return id;
#endif
}

// One cluster = NVR_CL (16) vertices stored as a 16-float4 structure-of-arrays block:
//     blk[0..3] = x of vertices 0..15, blk[4..7] = y, blk[8..11] = z, blk[12..15] = bit patterns of the ORIGINAL
//     vertex indices; padding slots hold x = y = z = +inf (distance +inf, never admitted ahead of a real vertex).
// Vertices are taken four at a time with the x/y/z differences, squares and sums of TWO vertices per packed
// instruction (FADD2 / FMUL2 / FFMA2).  On the device all 32 lanes of the warp must call this together: a group
// of four costs the distance arithmetic plus one vote when no lane can use it, and each candidate is inserted
// (branch-free, per lane) only when some lane admits it.  Exact: the result is the 4 smallest (d2, index) keys.
NVR_HD void nvr_knn_scan(const float4* blk, const float p[3], Knn4& k) {
    const F2 px = {p[0], p[0]}, py = {p[1], p[1]}, pz = {p[2], p[2]};
#pragma unroll 2
    for (int q = 0; q < 4; ++q) {
        const float4 x = nvr_ld_vert(blk + q), y = nvr_ld_vert(blk + 4 + q), z = nvr_ld_vert(blk + 8 + q), id = nvr_ld_vert(blk + 12 + q);
        const F2 dx0 = nvr_sub2(px, F2{x.x, x.y}), dx1 = nvr_sub2(px, F2{x.z, x.w});
        const F2 dy0 = nvr_sub2(py, F2{y.x, y.y}), dy1 = nvr_sub2(py, F2{y.z, y.w});
        const F2 dz0 = nvr_sub2(pz, F2{z.x, z.y}), dz1 = nvr_sub2(pz, F2{z.z, z.w});
        F2 d0 = nvr_mul2(dx0, dx0), d1 = nvr_mul2(dx1, dx1);
        nvr_fma2(d0, dy0, dy0); nvr_fma2(d1, dy1, dy1);
        nvr_fma2(d0, dz0, dz0); nvr_fma2(d1, dz1, dz1);           // squared L2, as knn_points returns it
        if (NVR_ANY(nvr_knn_admits(k, fminf(fminf(d0.x, d0.y), fminf(d1.x, d1.y))))) {
            if (NVR_ANY(nvr_knn_admits(k, d0.x))) { if (nvr_knn_admits(k, d0.x)) nvr_knn_insert(k, d0.x, nvr_f2i(id.x)); }
            if (NVR_ANY(nvr_knn_admits(k, d0.y))) { if (nvr_knn_admits(k, d0.y)) nvr_knn_insert(k, d0.y, nvr_f2i(id.y)); }
            if (NVR_ANY(nvr_knn_admits(k, d1.x))) { if (nvr_knn_admits(k, d1.x)) nvr_knn_insert(k, d1.x, nvr_f2i(id.z)); }
            if (NVR_ANY(nvr_knn_admits(k, d1.y))) { if (nvr_knn_admits(k, d1.y)) nvr_knn_insert(k, d1.y, nvr_f2i(id.w)); }
        }
    }
}

// Lower bound of the scan's distance over every v inside the box [lo, hi].  Each fp32 operation of the
// distance is monotone in |p - v| per axis, so the same expression on the per-axis gap is a lower
// bound up to the contraction (fma vs mul+add) the compiler picks; callers prune with a 1e-6
// relative slack for that.
NVR_HD float nvr_aabb_lb(const float4& lo, const float4& hi, const float p[3]) {
    const float dx = fmaxf(fmaxf(lo.x - p[0], p[0] - hi.x), 0.0f);
    const float dy = fmaxf(fmaxf(lo.y - p[1], p[1] - hi.y), 0.0f);
    const float dz = fmaxf(fmaxf(lo.z - p[2], p[2] - hi.z), 0.0f);
    return dx * dx + dy * dy + dz * dz;
}
#ifdef NVR_CL_OVERRIDE
#define NVR_CL NVR_CL_OVERRIDE
#else
#define NVR_CL 16
#endif
// vertices per spatial cluster (one AABB each)
#define NVR_PRUNE_SLACK 0.999999f

// From the 4 neighbours to their normalised Gaussian weights and the weighted distance:
// sample_blend_closest_points, :741-748.
NVR_HD float nvr_knn_weights(const Knn4& k, float w[NVR_KNN], float* wsum_out = nullptr) {
    float d[NVR_KNN], wsum = 0.0f;
#pragma unroll
    for (int i = 0; i < NVR_KNN; ++i) {
        d[i] = sqrtf(nvr_knn_d2(k, i));                          // cast_knn_points :736
        w[i] = expf(-(d[i] * d[i]) / 0.01125f);                  // :746, 2*radius^2 = 2*0.075^2
        wsum += w[i];
    }
    if (wsum_out) *wsum_out = wsum;
    const float denom = wsum + 1e-8f;                            // :747
    float pdist = 0.0f;
#pragma unroll
    for (int i = 0; i < NVR_KNN; ++i) {
        w[i] = w[i] / denom;
        pdist += d[i] * w[i];                                    // :748
    }
    return pdist;
}

// Blend weight of joint j from the 4 neighbour rows (:762).  pbw_part: (maxlen, 24) rows of this part.
NVR_HD float nvr_blend_joint(const int idx[NVR_KNN], const float w[NVR_KNN], const float* pbw_part, int j) {
    float b = 0.0f;
#pragma unroll
    for (int i = 0; i < NVR_KNN; ++i) b += pbw_part[(long long)idx[i] * NVR_JOINTS + j] * w[i];
    return b;
}

// (bw[24], pdist) in one call -- host emulation and per-stage tests.
NVR_HD float nvr_knn_blend(const Knn4& k, const float* pbw_part, float bw[NVR_JOINTS]) {
    float w[NVR_KNN];
    int idx[NVR_KNN];
    const float pdist = nvr_knn_weights(k, w);
    for (int i = 0; i < NVR_KNN; ++i) idx[i] = nvr_knn_idx(k, i);
    for (int j = 0; j < NVR_JOINTS; ++j) bw[j] = nvr_blend_joint(idx, w, pbw_part, j);
    return pdist;
}

// ---------------------------------------------------------------------------------------
// LBS: pose space -> T pose -> big pose             blend_utils.py:293-317, 395-487
// ---------------------------------------------------------------------------------------
// M, B: rows 0..2 of the blended pose / big-pose 4x4s.  In: pose-space point p, direction d.
// Out: big-pose point x0 and direction v.
NVR_HD void nvr_lbs_apply(const float M[12], const float B[12], const float p[3], const float d[3], float x0[3], float v[3]) {
    // 3x3 inverse by transposed cofactors / (det + fp32 eps)      torch_inverse_3x3 :293-317
    const float a = M[0], b = M[1], c = M[2], dd = M[4], e = M[5], f = M[6], g = M[8], h = M[9], i = M[10];
    const float m00 = e * i - f * h, m01 = dd * i - f * g, m02 = dd * h - e * g;
    const float m10 = b * i - c * h, m11 = a * i - c * g, m12 = a * h - b * g;
    const float m20 = b * f - c * e, m21 = a * f - c * dd, m22 = a * e - b * dd;
    const float det = (a * m00 - b * m01) + c * m02;
    const float den = det + 1.1920928955078125e-07f;
    float Ri[9];
    Ri[0] = m00 / den;  Ri[1] = -m10 / den; Ri[2] = m20 / den;
    Ri[3] = -m01 / den; Ri[4] = m11 / den;  Ri[5] = -m21 / den;
    Ri[6] = m02 / den;  Ri[7] = -m12 / den; Ri[8] = m22 / den;
    const float q[3] = {p[0] - M[3], p[1] - M[7], p[2] - M[11]};  // pose_points_to_tpose_points :433
    float t[3], td[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        t[r] = (Ri[r * 3 + 0] * q[0] + Ri[r * 3 + 1] * q[1]) + Ri[r * 3 + 2] * q[2];      // :436
        td[r] = (Ri[r * 3 + 0] * d[0] + Ri[r * 3 + 1] * d[1]) + Ri[r * 3 + 2] * d[2];     // :453
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        x0[r] = ((B[r * 4 + 0] * t[0] + B[r * 4 + 1] * t[1]) + B[r * 4 + 2] * t[2]) + B[r * 4 + 3];   // :469-470
        v[r] = (B[r * 4 + 0] * td[0] + B[r * 4 + 1] * td[1]) + B[r * 4 + 2] * td[2];                  // :486
    }
}

// A, bigA: (24,4,4) row-major.  Blend weights given explicitly (host emulation, tests).
NVR_HD void nvr_lbs_to_bigpose(const float bw[NVR_JOINTS], const float* A, const float* bigA,
                               const float p[3], const float d[3], float x0[3], float v[3]) {
    float M[12], B[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) { M[e] = 0.0f; B[e] = 0.0f; }
    for (int j = 0; j < NVR_JOINTS; ++j) {
        const float w = bw[j];
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            M[e] += w * A[j * 16 + e];                            // get_inverse_blend_params :415
            B[e] += w * bigA[j * 16 + e];                         // get_blend_params :402
        }
    }
    nvr_lbs_apply(M, B, p, d, x0, v);
}

// The production form: blend weights are produced joint by joint from the neighbour rows and folded
// straight into the two blended transforms (same operation order as the two-step form above), so the
// 24 weights never exist as an array and the joint loop stays rolled (small code, no local memory).
NVR_HD void nvr_blend_lbs(const int idx[NVR_KNN], const float w[NVR_KNN], const float* pbw_part, const float* A,
                          const float* bigA, const float p[3], const float d[3], float x0[3], float v[3]) {
    F2 M2[6], B2[6];                                              // element pairs (e, e+1): two FMAs per FFMA2
#pragma unroll
    for (int e = 0; e < 6; ++e) { M2[e] = F2{0.0f, 0.0f}; B2[e] = F2{0.0f, 0.0f}; }
    // joints four at a time: the neighbour rows are 24 contiguous floats (96 B, 16-byte aligned), so a quad of joints is ONE
    // 16-byte load per neighbour instead of four 4-byte ones; per joint the sum over the neighbours keeps nvr_blend_joint's order
#pragma unroll 1
    for (int jq = 0; jq < NVR_JOINTS / 4; ++jq) {
        float bq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < NVR_KNN; ++i) {
            const float4 r = nvr_ld_vert(reinterpret_cast<const float4*>(pbw_part + (long long)idx[i] * NVR_JOINTS) + jq);
            bq[0] += r.x * w[i]; bq[1] += r.y * w[i]; bq[2] += r.z * w[i]; bq[3] += r.w * w[i];      // :762
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float bj = bq[jj];
            const int j = jq * 4 + jj;
            // skinning rows are sparse (<= 4 joints per vertex) and a warp's pairs are neighbours: most joints have zero
            // weight for every lane, and adding 0 * A_j changes nothing
            if (!NVR_ANY_ACTIVE(bj != 0.0f)) continue;
            const F2 b2 = {bj, bj};
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                nvr_fma2(M2[e], b2, F2{A[j * 16 + 2 * e], A[j * 16 + 2 * e + 1]});
                nvr_fma2(B2[e], b2, F2{bigA[j * 16 + 2 * e], bigA[j * 16 + 2 * e + 1]});
            }
        }
    }
    float M[12], B[12];
#pragma unroll
    for (int e = 0; e < 6; ++e) { M[2 * e] = M2[e].x; M[2 * e + 1] = M2[e].y; B[2 * e] = B2[e].x; B[2 * e + 1] = B2[e].y; }
    nvr_lbs_apply(M, B, p, d, x0, v);
}

// ---------------------------------------------------------------------------------------
// view-direction encoding                                       freq_embedder.py:20-31
// ---------------------------------------------------------------------------------------
NVR_HD void nvr_posenc27(const float v[3], float* out) {
    out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float fr = (float)(1 << k);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float x = v[a] * fr;
            out[3 + k * 6 + a] = sinf(x);
            out[3 + k * 6 + 3 + a] = cosf(x);
        }
    }
}

// ---------------------------------------------------------------------------------------
// UV-time deformer                                              uv_deformer.py:23-45
// ---------------------------------------------------------------------------------------
// DeformerMlp: the nn.Linear tensors as they are stored (row-major (out,in)).  The forward reads a PACKED copy
// (shared memory on the device) laid out for 16-byte loads and paired FMAs:
//     w0p [32][20] = [W0 row (19) | b0]   (the input row gets a constant 1 in column 19)
//     w1  [32][32], b1 [32], w2 [3][32], b2 [3] (+1 pad)
struct DeformerMlp {
    const float *w0, *b0, *w1, *b1, *w2, *b2;    // 32x19, 32, 32x32, 32, 3x32, 3
};
#define NVR_DEF_W0P 0
#define NVR_DEF_W1 (32 * 20)
#define NVR_DEF_B1 (NVR_DEF_W1 + 32 * 32)
#define NVR_DEF_W2 (NVR_DEF_B1 + 32)
#define NVR_DEF_B2 (NVR_DEF_W2 + 3 * 32)
#define NVR_DEF_PACKED_FLOATS (NVR_DEF_B2 + 4)
// thread `tid` of `nthreads` writes its share of the packed block (host: tid 0 of 1)
NVR_HD void nvr_pack_deformer(const DeformerMlp& m, float* out, int tid, int nthreads) {
    for (int i = tid; i < 32 * 20; i += nthreads) {
        const int o = i / 20, k = i - o * 20;
        out[NVR_DEF_W0P + i] = k < 19 ? m.w0[o * 19 + k] : m.b0[o];
    }
    for (int i = tid; i < 32 * 32; i += nthreads) out[NVR_DEF_W1 + i] = m.w1[i];
    for (int i = tid; i < 32; i += nthreads) out[NVR_DEF_B1 + i] = m.b1[i];
    for (int i = tid; i < 3 * 32; i += nthreads) out[NVR_DEF_W2 + i] = m.w2[i];
    for (int i = tid; i < 4; i += nthreads) out[NVR_DEF_B2 + i] = i < 3 ? m.b2[i] : 0.0f;
}

// sum_k w[k] x[k] over K (a multiple of 4) inputs held in registers, weights by 16-byte loads; even / odd k
// accumulate in the two halves of one FFMA2 accumulator (fp32 re-association against a sequential sum: the
// reference's own sgemm does not promise an order either).
template <int K>
NVR_HD float nvr_dot_packed(const float* w, const float* x, float init) {
    F2 acc = {init, 0.0f};
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(w + 4 * q);
        nvr_fma2(acc, F2{t.x, t.y}, F2{x[4 * q], x[4 * q + 1]});
        nvr_fma2(acc, F2{t.z, t.w}, F2{x[4 * q + 2], x[4 * q + 3]});
    }
    return acc.x + acc.y;
}

// pk: the packed block (see above).  sc: 32 scratch floats of THIS thread, element stride ss (shared memory on the
// device: sc = base + tid, ss = block size, so neighbouring lanes hit neighbouring banks; a plain array with
// ss = 1 on the host).  The output loops stay rolled -- the activations live in the scratch row instead of 64
// registers -- which keeps the kernel's code inside the instruction cache.
NVR_HD void nvr_deformer_point(const GridDev& g, const float* pk, const VolumeDev& tuv, float frame_dim,
                               const float x0[3], float resd[3], float* sc, int ss) {
    float uvt[3];
    nvr_sample_volume(tuv, x0, 0, 2, uvt);                        // pts_sample_uv :32
    uvt[2] = frame_dim;                                           // :35
#ifdef __CUDA_ARCH__
    nvr_embed_point_f2_dev(g, uvt, sc, ss);                       // :37 (8 levels x 2 features, concat) -> sc[0..18]
#else
    nvr_embed_point<2>(g, uvt, sc, ss);
#endif
    {
        float e[20];
#pragma unroll
        for (int i = 0; i < 19; ++i) e[i] = sc[i * ss];
        e[19] = 1.0f;                                             // picks up b0 from column 19 of w0p
#pragma unroll 1
        for (int o = 0; o < 32; ++o) sc[o * ss] = nvr_softplus_hidden(nvr_dot_packed<20>(pk + NVR_DEF_W0P + o * 20, e, 0.0f));
    }
    float h[32];
    {
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i] = sc[i * ss];
#pragma unroll 1
        for (int o = 0; o < 32; ++o) sc[o * ss] = nvr_softplus_hidden(nvr_dot_packed<32>(pk + NVR_DEF_W1 + o * 32, h, pk[NVR_DEF_B1 + o]));
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) h[i] = sc[i * ss];
#pragma unroll
    for (int o = 0; o < 3; ++o) resd[o] = 0.05f * tanhf(nvr_dot_packed<32>(pk + NVR_DEF_W2 + o * 32, h, pk[NVR_DEF_B2 + o]));   // :39
}

// ---------------------------------------------------------------------------------------
// sampling along the ray + world -> pose              inb_renderer.py:15-31, blend_utils.py:366-382
// ---------------------------------------------------------------------------------------
// torch.linspace(0, 1, S)[k] in fp32: step = 1/(S-1); first half start + k*step, second half
// end - (S-1-k)*step (ATen's symmetric formula).
NVR_HD float nvr_linspace01(int k, int S) {
    if (S == 1) return 0.0f;
    const float step = 1.0f / (float)(S - 1);
    return (k < S / 2) ? (0.0f + step * (float)k) : (1.0f - step * (float)(S - 1 - k));
}

NVR_HD void nvr_ray_sample(const float o[3], const float d[3], float near_, float far_, int k, int S, float wp[3]) {
    const float t = nvr_linspace01(k, S);
    const float z = near_ * (1.0f - t) + far_ * t;               // :18
#pragma unroll
    for (int a = 0; a < 3; ++a) wp[a] = o[a] + d[a] * z;          // :29
}

// p = (w - Th) . R   (row vector times matrix), v = d . R
NVR_HD void nvr_world_to_pose(const float* R, const float* Th, const float w[3], float p[3]) {
    const float q[3] = {w[0] - Th[0], w[1] - Th[1], w[2] - Th[2]};
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = (q[0] * R[0 * 3 + c] + q[1] * R[1 * 3 + c]) + q[2] * R[2 * 3 + c];
}
NVR_HD void nvr_dir_to_pose(const float* R, const float d[3], float v[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (d[0] * R[0 * 3 + c] + d[1] * R[1 * 3 + c]) + d[2] * R[2 * 3 + c];
}

// ---------------------------------------------------------------------------------------
// camera rays, the step before the path        lib/utils/if_nerf/if_nerf_data_utils.py:24-38 (get_rays),
//                                               :92-107 (get_near_far), :329-343 (get_rays_within_bounds)
// ---------------------------------------------------------------------------------------
// get_rays runs in float64 in the reference (K, R, T come from the annotation files as float64; only the
// result is cast to float32, :332-333), so the pixel -> ray arithmetic here is double as well.
struct CameraDev {
    double Kinv[9];                // np.linalg.inv(K), row-major (taken on the host with the same LAPACK call)
    double R[9];                   // world -> camera rotation, row-major
    double T[3];
    double o[3];                   // camera origin -R^T T (:26)
};

// pixel (col i, row j) -> unit ray direction (float32, as after `.astype(np.float32)`)
NVR_HD void nvr_pixel_ray(const CameraDev& c, int i, int j, float ray_d[3]) {
    const double xy1[3] = {(double)(float)i, (double)(float)j, 1.0};
    double pc[3], pw[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)                                  // xy1 . inv(K)^T  (:32)
        pc[a] = (xy1[0] * c.Kinv[a * 3 + 0] + xy1[1] * c.Kinv[a * 3 + 1]) + xy1[2] * c.Kinv[a * 3 + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a)                                  // (pixel_camera - T) . R  (:33)
        pw[a] = ((pc[0] - c.T[0]) * c.R[0 * 3 + a] + (pc[1] - c.T[1]) * c.R[1 * 3 + a]) + (pc[2] - c.T[2]) * c.R[2 * 3 + a];
    const double d[3] = {pw[0] - c.o[0], pw[1] - c.o[1], pw[2] - c.o[2]};   // :35
    const double nrm = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);     // :36
#pragma unroll
    for (int a = 0; a < 3; ++a) ray_d[a] = (float)(d[a] / nrm);
}

#ifdef __CUDA_ARCH__
#define NVR_FMUL(a, b) __fmul_rn((a), (b))
#define NVR_FADD(a, b) __fadd_rn((a), (b))
#else
#define NVR_FMUL(a, b) ((a) * (b))
#define NVR_FADD(a, b) ((a) + (b))
#endif

// get_near_far (:92-107) in float32 on one ray; `o0` is the FIRST ray's origin (the reference indexes
// ray_o[:1], i.e. it assumes one camera origin for the whole batch).  Returns mask_at_box = near < far;
// near / far are already divided by |ray_d| (:105-106).  No fused multiply-adds: near < far is a bit-level
// decision in the reference.
NVR_HD bool nvr_near_far(const float* bounds /* (2,3) */, const float o0[3], const float ray_d[3], float* near_, float* far_) {
    const float norm_d = sqrtf(NVR_FADD(NVR_FADD(NVR_FMUL(ray_d[0], ray_d[0]), NVR_FMUL(ray_d[1], ray_d[1])), NVR_FMUL(ray_d[2], ray_d[2])));
    float nr = -INFINITY, fr = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = ray_d[a] / norm_d;
        if (v < 1e-5f && v > -1e-10f) v = 1e-5f;                 // :96
        if (v > -1e-5f && v < 1e-10f) v = -1e-5f;                // :97
        const float tmin = (bounds[a] - o0[a]) / v;              // :98
        const float tmax = (bounds[3 + a] - o0[a]) / v;          // :99
        nr = fmaxf(nr, fminf(tmin, tmax));                       // :100,102
        fr = fminf(fr, fmaxf(tmin, tmax));                       // :101,103
    }
    *near_ = nr / norm_d;
    *far_ = fr / norm_d;
    return nr < fr;                                              // :104
}

// ---------------------------------------------------------------------------------------
// Adam, one element                torch.optim.Adam (single-tensor form, amsgrad=False, maximize=False) as
//                                  lib/train/optimizer.py:27 constructs it (eps = cfg.train.eps = 1e-15)
// ---------------------------------------------------------------------------------------
struct AdamScalars {
    float beta1, beta2, one_minus_beta1, one_minus_beta2;
    float eps, weight_decay;
    float neg_step_size;           // -lr / (1 - beta1^t)
    float bc2_sqrt;                // sqrt(1 - beta2^t)
};
NVR_HD void nvr_adam_update(const AdamScalars& s, float& p, float g, float& m, float& v) {
    if (s.weight_decay != 0.0f) g = g + s.weight_decay * p;      // grad.add(param, alpha=weight_decay)
    m = m + s.one_minus_beta1 * (g - m);                         // exp_avg.lerp_(grad, 1 - beta1)
    v = v * s.beta2 + s.one_minus_beta2 * (g * g);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;           // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p + s.neg_step_size * (m / denom);                       // param.addcdiv_(exp_avg, denom, value=-step_size)
}

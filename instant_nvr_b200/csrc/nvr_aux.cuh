// nvr_aux.cuh -- the steps either side of the per-ray path:
//
//   rank 1  k_adam                      dense Adam over every trainable tensor in one launch per <=24 tensors
//                                       (torch.optim.Adam as lib/train/optimizer.py:27 builds it, eps 1e-15)
//   rank 2  k_rays_mask / k_rays_scan / k_rays_emit
//                                       get_rays_within_bounds[_coord]: pixel -> ray, bbox near/far, mask_at_box and
//                                       the ORDERED compaction of the surviving rays
//                                       (lib/utils/if_nerf/if_nerf_data_utils.py:24-38, 92-107, 329-343)
//   rank 4  k_assemble_image, k_sq_diff image assembly `img[mask_at_box] = rgb` and the evaluator's MSE
//                                       (lib/evaluators/if_nerf.py:28-31, 84-113)
//
// All of it is streaming work: one pass over the data, 16-byte accesses, no reuse.
#pragma once
#include <cuda_runtime.h>

#include "nvr_math.cuh"

// ---- Adam ------------------------------------------------------------------------------------------
#define ADAM_MAX_TENSORS 24
#define ADAM_CHUNK 8192            // elements per CTA: 256 threads x 8 float4
struct AdamTensorDev {
    float* p; float* g; float* m; float* v;
    long long n;
    AdamScalars s;
    int vec;                       // all four pointers 16-byte aligned
};
struct AdamBatch {
    AdamTensorDev t[ADAM_MAX_TENSORS];
    int chunk_begin[ADAM_MAX_TENSORS + 1];
    int n_tensors;
    int zero_grad;
};

// Traffic per element: read p, g, m, v, write p, m, v (+ g when zero_grad) = 28 (32) bytes; nothing is reused, so
// every access is a streaming (evict-first) 16-byte one.
__global__ void __launch_bounds__(256) k_adam(const __grid_constant__ AdamBatch b) {
    int t = 0;
    const int c = blockIdx.x;
    while (t + 1 < b.n_tensors && c >= b.chunk_begin[t + 1]) ++t;
    const AdamTensorDev& T = b.t[t];
    const AdamScalars s = T.s;
    const long long base = (long long)(c - b.chunk_begin[t]) * ADAM_CHUNK;
    const long long end = base + ADAM_CHUNK < T.n ? base + ADAM_CHUNK : T.n;
    long long i = base + (long long)threadIdx.x * 4;
    if (T.vec) {
        for (; i + 3 < end; i += 256 * 4) {
            float4 p = __ldcs((const float4*)(T.p + i)), g = __ldcs((const float4*)(T.g + i));
            float4 m = __ldcs((const float4*)(T.m + i)), v = __ldcs((const float4*)(T.v + i));
            nvr_adam_update(s, p.x, g.x, m.x, v.x);
            nvr_adam_update(s, p.y, g.y, m.y, v.y);
            nvr_adam_update(s, p.z, g.z, m.z, v.z);
            nvr_adam_update(s, p.w, g.w, m.w, v.w);
            __stcs((float4*)(T.p + i), p);
            __stcs((float4*)(T.m + i), m);
            __stcs((float4*)(T.v + i), v);
            if (b.zero_grad) __stcs((float4*)(T.g + i), make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
    // tail of an aligned tensor (< 4 elements, one thread) or the whole chunk of an unaligned one
    for (long long j = T.vec ? i : base + threadIdx.x; j < end; j += T.vec ? 1 : 256) {
        if (T.vec && j >= i + 4) break;
        float p = T.p[j], m = T.m[j], v = T.v[j];
        nvr_adam_update(s, p, T.g[j], m, v);
        T.p[j] = p; T.m[j] = m; T.v[j] = v;
        if (b.zero_grad) T.g[j] = 0.0f;
    }
}

// ---- camera rays --------------------------------------------------------------------------------------
#define RAYS_BLOCK 1024            // pixels per CTA (also the compaction tile)

// mask_at_box per pixel (row-major, pixel = j*W + i) and the number of surviving pixels per 1024-pixel tile
__global__ void __launch_bounds__(RAYS_BLOCK) k_rays_mask(const __grid_constant__ CameraDev cam, int H, int W,
                                                          const float* __restrict__ bounds, unsigned char* __restrict__ mask,
                                                          int* __restrict__ tile_count) {
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const long long pix = (long long)blockIdx.x * RAYS_BLOCK + threadIdx.x;
    bool in = false;
    if (pix < (long long)H * W) {
        float d[3], nr, fr;
        nvr_pixel_ray(cam, (int)(pix % W), (int)(pix / W), d);
        const float o0[3] = {(float)cam.o[0], (float)cam.o[1], (float)cam.o[2]};
        in = nvr_near_far(bounds, o0, d, &nr, &fr);
        mask[pix] = in ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&s_count, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_count;
}

// exclusive scan of the tile counts (one CTA; a 4K frame has 8 100 tiles); tile_off[n_tiles] = total
__global__ void __launch_bounds__(1024) k_rays_scan(const int* __restrict__ tile_count, int n_tiles, int* __restrict__ tile_off,
                                                    int* __restrict__ n_rays_out) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < n_tiles ? tile_count[idx] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_excl = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int carry = s_carry;
        if (idx < n_tiles) tile_off[idx] = carry + warp_excl + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) { tile_off[n_tiles] = s_carry; *n_rays_out = s_carry; }
}

// surviving pixels in row-major order -> ray_o, ray_d (n,3), near, far (n), coord (n) = pixel index j*W + i
__global__ void __launch_bounds__(RAYS_BLOCK) k_rays_emit(const __grid_constant__ CameraDev cam, int H, int W,
                                                          const float* __restrict__ bounds, const int* __restrict__ tile_off,
                                                          float* __restrict__ ray_o, float* __restrict__ ray_d,
                                                          float* __restrict__ near_, float* __restrict__ far_, int* __restrict__ coord) {
    __shared__ int s_warp[32];
    const long long pix = (long long)blockIdx.x * RAYS_BLOCK + threadIdx.x;
    bool in = false;
    float d[3] = {0.f, 0.f, 0.f}, nr = 0.f, fr = 0.f;
    const float o0[3] = {(float)cam.o[0], (float)cam.o[1], (float)cam.o[2]};
    if (pix < (long long)H * W) {
        nvr_pixel_ray(cam, (int)(pix % W), (int)(pix / W), d);
        in = nvr_near_far(bounds, o0, d, &nr, &fr);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int v = s_warp[threadIdx.x];
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (threadIdx.x >= o) x += y;
        }
        s_warp[threadIdx.x] = x - v;
    }
    __syncthreads();
    if (!in) return;
    const long long r = (long long)tile_off[blockIdx.x] + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
    for (int a = 0; a < 3; ++a) { ray_o[r * 3 + a] = o0[a]; ray_d[r * 3 + a] = d[a]; }
    near_[r] = nr; far_[r] = fr;
    if (coord) coord[r] = (int)pix;
}

// ---- image assembly + MSE ----------------------------------------------------------------------------
// img (n_pix,3) must be zero-filled by the caller's memset; img[coord[r]] = rgb[r]
__global__ void k_assemble_image(const float* __restrict__ rgb, const int* __restrict__ coord, long long n, float* __restrict__ img) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const long long p = coord[r];
#pragma unroll
        for (int a = 0; a < 3; ++a) img[p * 3 + a] = rgb[r * 3 + a];
    }
}

// out[0] += sum (a - b)^2 in float64 (the evaluator works on float64 images holding float32 values)
__global__ void __launch_bounds__(256) k_sq_diff(const float* __restrict__ a, const float* __restrict__ b, long long n, double* __restrict__ out) {
    __shared__ double s_part[8];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = (double)a[i] - (double)b[i];
        acc += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        atomicAdd(out, t);
    }
}

// ---- SSIM ------------------------------------------------------------------------------------------------------------
// The evaluator's structural similarity (lib/evaluators/if_nerf.py:33-74): skimage 0.19.3 `structural_similarity(a, b,
// multichannel=True)` on the float64 crop [y0, y0+h) x [x0, x0+w) of two (H, W, 3) images holding float32 values -- uniform
// 7 x 7 windows, sample covariance (49/48), data_range 2 (skimage's dtype range of float images: C1 = (0.01 * 2)^2,
// C2 = (0.03 * 2)^2), mean of S over the interior (3 pixels cropped on every side) and over the channels.
// One thread per interior pixel and channel, window sums in float64; out[c] += sum of S over the channel's interior.
__global__ void __launch_bounds__(256)
k_ssim(const float* __restrict__ a, const float* __restrict__ b, int W, int x0, int y0, int w, int h, double* __restrict__ out) {
    __shared__ double s_part[8][3];
    const int iw = w - 6, ih = h - 6;
    const long long n = (long long)iw * ih * 3;
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % 3);
        const long long px = i / 3;
        const int cx = x0 + 3 + (int)(px % iw), cy = y0 + 3 + (int)(px / iw);
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        for (int dy = -3; dy <= 3; ++dy)
            for (int dx = -3; dx <= 3; ++dx) {
                const long long o = ((long long)(cy + dy) * W + (cx + dx)) * 3 + c;
                const double x = a[o], y = b[o];
                sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
            }
        const double ux = sx / 49.0, uy = sy / 49.0, uxx = sxx / 49.0, uyy = syy / 49.0, uxy = sxy / 49.0;
        const double cov = 49.0 / 48.0;
        const double vx = cov * (uxx - ux * ux), vy = cov * (uyy - uy * uy), vxy = cov * (uxy - ux * uy);
        const double C1 = 0.0004, C2 = 0.0036;
        const double S = ((2.0 * ux * uy + C1) * (2.0 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
        acc[c] += S;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][c] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int wp = 0; wp < 8; ++wp) t += s_part[wp][threadIdx.x];
        atomicAdd(out + threadIdx.x, t);
    }
}

// nvr_smpl.cuh -- per-frame SMPL preprocessing on the device: everything the reference's
// dataset and its offline tool compute per frame to produce the tensors the per-ray path consumes.
//
//   k_smpl_transforms   Rodrigues + kinematic chain for `poses` and `big_poses` -> A, big_A (24,4,4); R = Rodrigues(Rh)
//                       (lib/utils/if_nerf/if_nerf_data_utils.py:523-577; lib/datasets/h36m/tpose_dataset.py:257-289)
//   k_smpl_pose_verts   ppts = (wxyz - Th) . R in float32 (tpose_dataset.py:269), scattered into the per-part vertex table
//                       (tpose_dataset.py:579-584), pbounds / wbounds (if_nerf_data_utils.py:689-696), and the float64
//                       posed vertices + their bbox the volume tool works on (tools/prepare_zjumocap.py:485, 152-156)
//   k_smpl_bounds       finishes the two float32 bboxes (+- cfg.box_padding)
//   k_smpl_bweights     the (D,H,W,25) volume: per 2.5 cm voxel the 24 skinning weights of the nearest posed vertex +
//                       the distance to it (tools/prepare_zjumocap.py:474-508; psbody closest_vertices = exact 1-NN),
//                       float64 brute force over shared-memory vertex tiles -- replaces the pre-baked
//                       lbs/bweights/{frame}.npy files
//
// The `__host__ __device__` arithmetic is compiled for the host by tests/host_emul (test-only).
#pragma once
#include "nvr_math.cuh"

struct SmplPoseDev {               // per frame, passed by value (kernel parameter space)
    double Rh[3], Th[3];
    double poses[NVR_JOINTS * 3];
    double big_poses[NVR_JOINTS * 3];
    float joints[NVR_JOINTS * 3];
    int parents[NVR_JOINTS];
};

// cv2.Rodrigues for a rotation vector (OpenCV calib3d): double arithmetic; identity below DBL_EPSILON.
NVR_HD void nvr_rodrigues_cv(const double r[3], double R[9]) {
    double x = r[0], y = r[1], z = r[2];
    const double theta = sqrt((x * x + y * y) + z * z);
    if (theta < 2.220446049250313e-16) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    x *= it; y *= it; z *= it;
    const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
    const double rx[9] = {0.0, -z, y, z, 0.0, -x, -y, x, 0.0};
    for (int i = 0; i < 9; ++i) R[i] = (c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[i]) + s * rx[i];
}

// batch_rodrigues (if_nerf_data_utils.py:523-542) of ONE joint, on float64 poses: angle = |p + 1e-8|, K = [p/angle]_x,
// R = I + sin K + (1 - cos) K.K
NVR_HD void nvr_rodrigues_smpl(const double p[3], double R[9]) {
    const double q[3] = {p[0] + 1e-8, p[1] + 1e-8, p[2] + 1e-8};
    const double angle = sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    const double rx = p[0] / angle, ry = p[1] / angle, rz = p[2] / angle;
    const double c = cos(angle), s = sin(angle);
    const double K[9] = {0.0, -rz, ry, rz, 0.0, -rx, -ry, rx, 0.0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double kk = (K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j]) + K[i * 3 + 2] * K[2 * 3 + j];
            R[i * 3 + j] = ((i == j ? 1.0 : 0.0) + s * K[i * 3 + j]) + (1.0 - c) * kk;
        }
}

// get_rigid_transformation (if_nerf_data_utils.py:545-577): `G` is scratch for the 24 chained 4x4 (double); out (24,4,4) f32.
// Serial over joints (parents precede children); 24 x 64 double FMAs.
NVR_HD void nvr_smpl_chain(const double* poses, const float* joints, const int* parents, double* G, float* out) {
    for (int j = 0; j < NVR_JOINTS; ++j) {
        double L[16], R[9];
        nvr_rodrigues_smpl(poses + j * 3, R);
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) L[a * 4 + b] = R[a * 3 + b];
            // rel_joints[1:] -= joints[parents[1:]] in the dtype of `joints`: float32 (:554-555)
            L[a * 4 + 3] = (double)(j == 0 ? joints[a] : joints[j * 3 + a] - joints[parents[j] * 3 + a]);
        }
        L[12] = L[13] = L[14] = 0.0; L[15] = 1.0;
        double* g = G + j * 16;
        if (j == 0) {
            for (int i = 0; i < 16; ++i) g[i] = L[i];
        } else {
            const double* pg = G + parents[j] * 16;                                   // np.dot(chain[parent], local) :566
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b)
                    g[a * 4 + b] = ((pg[a * 4 + 0] * L[0 * 4 + b] + pg[a * 4 + 1] * L[1 * 4 + b]) + pg[a * 4 + 2] * L[2 * 4 + b]) + pg[a * 4 + 3] * L[3 * 4 + b];
        }
    }
    for (int j = 0; j < NVR_JOINTS; ++j) {
        const double* g = G + j * 16;
        for (int a = 0; a < 4; ++a) {
            // transforms[..., 3] -= sum(transforms * [joint, 0], axis=2)  (:571-574)
            const double rel = ((g[a * 4 + 0] * (double)joints[j * 3 + 0] + g[a * 4 + 1] * (double)joints[j * 3 + 1]) + g[a * 4 + 2] * (double)joints[j * 3 + 2]) + g[a * 4 + 3] * 0.0;
            for (int b = 0; b < 3; ++b) out[j * 16 + a * 4 + b] = (float)g[a * 4 + b];
            out[j * 16 + a * 4 + 3] = (float)(g[a * 4 + 3] - rel);
        }
    }
}

// np.arange(start, stop, step) in float64: length ceil((stop - start) / step); element i = start + i * delta with
// delta = (start + step) - start  (numpy fills from its first two elements).  No fused multiply-add: the voxel
// centres are the reference's bit for bit.
NVR_HD int nvr_arange_len(double start, double stop, double step) {
    const double n = ceil((stop - start) / step);
    return n <= 0.0 ? 0 : (n > 2147483647.0 ? 2147483647 : (int)n);
}
NVR_HD double nvr_arange_val(double start, double step, int i) {
    const double delta = (start + step) - start;
#ifdef __CUDA_ARCH__
    return __dadd_rn(start, __dmul_rn((double)i, delta));
#else
    return start + (double)i * delta;
#endif
}

#ifdef __CUDACC__
// Workspace layout (nvr_smpl_workspace_bytes): [0] float min/max of ppts and wxyz as order-preserving uint32 (12 words),
// [64 B] double min/max of the float64 posed vertices as order-preserving uint64 (6 words), [128 B] R64 (9 doubles) +
// Th64 (3 doubles), [256 B] float64 posed vertices (V,3).
#define SMPL_WS_F32BOX 0
#define SMPL_WS_F64BOX 64
#define SMPL_WS_RT 128
#define SMPL_WS_PXYZ 256

__device__ __forceinline__ unsigned int ord32(float f) {
    const unsigned int u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float unord32(unsigned int u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}
__device__ __forceinline__ unsigned long long ord64(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return u ^ ((u >> 63) ? ~0ull : 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double unord64(unsigned long long u) {
    u ^= (u >> 63) ? 0x8000000000000000ull : ~0ull;
    double d;
    memcpy(&d, &u, 8);
    return d;
}

// One CTA, 64 threads: warp 0 lane 0 chains `poses`, warp 1 lane 0 chains `big_poses`, thread 1 takes R.  The chain is
// serial by construction (a joint needs its parent) and 1.5 k double FMAs long: microseconds, off the critical path.
__global__ void __launch_bounds__(64) k_smpl_transforms(const __grid_constant__ SmplPoseDev P, float* __restrict__ A,
                                                        float* __restrict__ bigA, float* __restrict__ R_out,
                                                        float* __restrict__ Th_out, unsigned char* __restrict__ ws) {
    __shared__ double G[2][NVR_JOINTS * 16];
    const int t = threadIdx.x;
    if (t == 0 && A) nvr_smpl_chain(P.poses, P.joints, P.parents, G[0], A);
    if (t == 32 && bigA) nvr_smpl_chain(P.big_poses, P.joints, P.parents, G[1], bigA);
    if (t == 1) {
        // dataset: Rh, Th cast to float32 first, R = cv2.Rodrigues(Rh).astype(float32)  (tpose_dataset.py:257-259)
        const double rh32[3] = {(double)(float)P.Rh[0], (double)(float)P.Rh[1], (double)(float)P.Rh[2]};
        double R[9];
        nvr_rodrigues_cv(rh32, R);
        for (int i = 0; i < 9; ++i) R_out[i] = (float)R[i];
        for (int a = 0; a < 3; ++a) Th_out[a] = (float)P.Th[a];
        // tool: the stored values as they are, float64  (tools/prepare_zjumocap.py:144-146)
        double* rt = reinterpret_cast<double*>(ws + SMPL_WS_RT);
        nvr_rodrigues_cv(P.Rh, rt);
        for (int a = 0; a < 3; ++a) rt[9 + a] = P.Th[a];
    }
    if (t >= 2 && t < 14) reinterpret_cast<unsigned int*>(ws + SMPL_WS_F32BOX)[t - 2] = ((t - 2) % 6 < 3) ? 0xffffffffu : 0u;
    if (t >= 14 && t < 20) reinterpret_cast<unsigned long long*>(ws + SMPL_WS_F64BOX)[t - 14] = (t - 14 < 3) ? ~0ull : 0ull;
}

// vert_slot[v] = part * maxlen + rank of v among its part's vertices (static per subject), or -1.
__global__ void __launch_bounds__(256) k_smpl_pose_verts(const float* __restrict__ wxyz, int n_verts, const float* __restrict__ R32,
                                                         const float* __restrict__ Th32, const int* __restrict__ vert_slot,
                                                         float* __restrict__ ppts, float* __restrict__ part_pts,
                                                         unsigned char* __restrict__ ws) {
    __shared__ unsigned int s32[12];
    __shared__ unsigned long long s64[6];
    if (threadIdx.x < 12) s32[threadIdx.x] = (threadIdx.x % 6 < 3) ? 0xffffffffu : 0u;
    if (threadIdx.x < 6) s64[threadIdx.x] = (threadIdx.x < 3) ? ~0ull : 0ull;
    __syncthreads();
    const double* rt = reinterpret_cast<const double*>(ws + SMPL_WS_RT);
    double* pxyz64 = reinterpret_cast<double*>(ws + SMPL_WS_PXYZ);
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_verts) {
        const float w[3] = {wxyz[v * 3], wxyz[v * 3 + 1], wxyz[v * 3 + 2]};
        float p[3];
        nvr_world_to_pose(R32, Th32, w, p);                                           // tpose_dataset.py:269
        const int slot = vert_slot ? vert_slot[v] : -1;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            ppts[v * 3 + a] = p[a];
            if (slot >= 0) part_pts[(long long)slot * 3 + a] = p[a];                  // tpose_dataset.py:582
            const double q = (((double)w[0] - rt[9]) * rt[0 * 3 + a] + ((double)w[1] - rt[10]) * rt[1 * 3 + a]) + ((double)w[2] - rt[11]) * rt[2 * 3 + a];
            pxyz64[(long long)v * 3 + a] = q;                                         // prepare_zjumocap.py:485
            atomicMin(&s32[a], ord32(p[a])); atomicMax(&s32[3 + a], ord32(p[a]));
            atomicMin(&s32[6 + a], ord32(w[a])); atomicMax(&s32[9 + a], ord32(w[a]));
            atomicMin(&s64[a], ord64(q)); atomicMax(&s64[3 + a], ord64(q));
        }
    }
    __syncthreads();
    unsigned int* g32 = reinterpret_cast<unsigned int*>(ws + SMPL_WS_F32BOX);
    unsigned long long* g64 = reinterpret_cast<unsigned long long*>(ws + SMPL_WS_F64BOX);
    if (threadIdx.x < 12) { if (threadIdx.x % 6 < 3) atomicMin(&g32[threadIdx.x], s32[threadIdx.x]); else atomicMax(&g32[threadIdx.x], s32[threadIdx.x]); }
    if (threadIdx.x < 6) { if (threadIdx.x < 3) atomicMin(&g64[threadIdx.x], s64[threadIdx.x]); else atomicMax(&g64[threadIdx.x], s64[threadIdx.x]); }
}

// get_bounds (if_nerf_data_utils.py:689-696): float32 min - 0.05f / max + 0.05f
__global__ void k_smpl_bounds(const unsigned char* __restrict__ ws, float padding, float* __restrict__ pbounds, float* __restrict__ wbounds) {
    const int t = threadIdx.x;
    if (t >= 12) return;
    const unsigned int* g32 = reinterpret_cast<const unsigned int*>(ws + SMPL_WS_F32BOX);
    const float v = unord32(g32[t]);
    float* dst = t < 6 ? pbounds : wbounds;
    if (dst) dst[t % 6] = (t % 6 < 3) ? v - padding : v + padding;
}

// One thread per voxel, vertices staged through shared memory 1024 at a time (24 KB of doubles per tile), float64
// distances: 2.4e8 - 6e8 pair evaluations per frame, well under a millisecond of FP64 work on B200.
#define SMPL_VTILE 1024
__global__ void __launch_bounds__(256) k_smpl_bweights(const unsigned char* __restrict__ ws, int n_verts, const float* __restrict__ weights,
                                                       int D, int H, int W, double x0, double y0, double z0, double step,
                                                       float* __restrict__ pbw) {
    __shared__ double sv[SMPL_VTILE * 3];
    const double* pxyz64 = reinterpret_cast<const double*>(ws + SMPL_WS_PXYZ);
    const long long n_vox = (long long)D * H * W;
    const long long vox = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = vox < n_vox;
    const int iz = live ? (int)(vox % W) : 0, iy = live ? (int)((vox / W) % H) : 0, ix = live ? (int)(vox / ((long long)W * H)) : 0;
    const double px = nvr_arange_val(x0, step, ix), py = nvr_arange_val(y0, step, iy), pz = nvr_arange_val(z0, step, iz);
    double best = INFINITY;
    int best_i = 0;
    for (int base = 0; base < n_verts; base += SMPL_VTILE) {
        const int cnt = min(SMPL_VTILE, n_verts - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 3; i += blockDim.x) sv[i] = pxyz64[(long long)base * 3 + i];
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < cnt; ++i) {
            const double dx = px - sv[i * 3], dy = py - sv[i * 3 + 1], dz = pz - sv[i * 3 + 2];
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (d2 < best) { best = d2; best_i = base + i; }                          // strict: lowest index on ties
        }
    }
    if (!live) return;
    float* o = pbw + vox * 25;
    const float* wr = weights + (long long)best_i * NVR_JOINTS;
#pragma unroll
    for (int j = 0; j < NVR_JOINTS; ++j) o[j] = wr[j];                                // prepare_zjumocap.py:494
    o[24] = (float)sqrt(best);                                                        // :493, 505-506
}
#endif  // __CUDACC__

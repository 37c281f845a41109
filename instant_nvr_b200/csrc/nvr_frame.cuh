// nvr_frame.cuh -- multi-GPU frame assembly over NVLink peer memory.
//
// Rays shard over the ranks in interleaved tiles (tile t of the frame belongs to rank t % world); the only exchange of a
// render is assembling the frame, 16 B per ray.  Instead of a library all-gather (pad to equal shards, all-gather, un-permute:
// three extra passes over the frame plus a collective launch), the kernel that finishes a ray -- k_resolve_rays -- stores
// [r, g, b, acc] straight to the ray's FINAL position in every rank's frame buffer: a local store for its own copy, NVLink
// peer stores (posted writes) for the others.  One flag barrier per frame (k_frame_barrier) then makes the frame complete
// on every rank.
//
// Frame buffers are cudaMalloc'd by the library and shared between the ranks' processes with CUDA IPC handles (exchanged by the
// host side through whatever control plane it has: torch.distributed's store in the Python mirror).  Two frame slots alternate,
// so a rank may start writing frame k + 1 into its peers while they still read frame k; a rank reaches frame k + 2 (same slot)
// only through barrier k + 1, which a peer signals after everything it enqueued before it -- including its reads of frame k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NVR_MAX_RANKS 8
#define NVR_FRAME_FLAG_STRIDE 32          // uint32 words per flag (one 128-byte line each)
#define NVR_FRAME_HEADER_BYTES (NVR_MAX_RANKS * NVR_FRAME_FLAG_STRIDE * 4)

struct FrameOut {                         // kernel parameter: where a finished ray goes
    float4* slot[NVR_MAX_RANKS];          // this frame's slot in every rank's buffer (peer-mapped), [n_total] float4
    int world, rank, tile;
    long long n_total;
};

// local ray i of this rank's shard -> its index in the frame (tile k of the shard is tile k * world + rank of the frame)
__device__ __forceinline__ long long frame_index(const FrameOut& f, long long i) {
    return ((i / f.tile) * f.world + f.rank) * (long long)f.tile + (i % f.tile);
}

// Flag barrier over peer memory: rank r writes its epoch into flags[r] of every peer, then waits until every peer's epoch has
// arrived in its own flags.  One warp; lane = peer.  Everything this rank stored to its peers in earlier kernels of the stream
// is ordered before the flag by the kernel boundary plus the system-scope fence.
struct FrameFlags { unsigned int* peer[NVR_MAX_RANKS]; };   // every rank's flag block (peer-mapped); peer[rank] is the local one
__global__ void k_frame_barrier(FrameFlags ff, int world, int rank, unsigned int epoch) {
    unsigned int* local_flags = nullptr;
#pragma unroll
    for (int r = 0; r < NVR_MAX_RANKS; ++r) local_flags = rank == r ? ff.peer[r] : local_flags;
    const int lane = threadIdx.x;
    __threadfence_system();
    if (lane < world) {
        unsigned int* peer = nullptr;
#pragma unroll
        for (int r = 0; r < NVR_MAX_RANKS; ++r) peer = lane == r ? ff.peer[r] : peer;
        unsigned int* dst = peer + rank * NVR_FRAME_FLAG_STRIDE;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
        const unsigned int* src = local_flags + lane * NVR_FRAME_FLAG_STRIDE;
        unsigned int v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        } while ((int)(v - epoch) < 0);
    }
    __syncwarp();
    __threadfence_system();
}

// Stand-alone scatter of already composited rays (rgb (n,3), acc (n)) into every rank's frame slot: the unfused form, for callers
// that rendered through the plain nvr_render_rays.
__global__ void k_frame_scatter(FrameOut f, const float* __restrict__ rgb, const float* __restrict__ acc, long long n_local) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (long long)gridDim.x * blockDim.x) {
        const long long gi = frame_index(f, i);
        if (gi >= f.n_total) continue;
        const float4 v = make_float4(rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2], acc[i]);
#pragma unroll
        for (int r = 0; r < NVR_MAX_RANKS; ++r)
            if (r < f.world) f.slot[r][gi] = v;
    }
}

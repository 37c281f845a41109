/* nvr_b200.h -- C-ABI of the B200-native instant-nvr per-ray hot path.
 *
 * The reference (zju3dv/instant-nvr @ a6f4d68) has no FFI for this path: it is PyTorch ops
 * plus pytorch3d's KNN behind two Python seams,
 *     Network.forward(wpts, viewdir, dists, batch)   lib/networks/bw_deform/inb_part_network_multiassign.py:126-168
 *     Renderer.render(batch)                         lib/networks/renderer/inb_renderer.py:204-239
 * The entry points below are what a ctypes binding for those two seams binds (INTEGRATION.md
 * shows the stub); each cites the reference code it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *   - every pointer in NvrParams / NvrFrame and every array argument without the `_host` suffix is
 *     a DEVICE pointer, fp32 row-major unless noted, BORROWED for the duration of the binding /
 *     call (the caller -- PyTorch -- owns parameter and batch storage, so optimizers and
 *     checkpoints keep working on the same memory).
 *   - all work is enqueued on the caller's stream; no entry point synchronises the device except
 *     the `_host` variants (which must, to hand back host results) and nvr_read_counters.
 *   - return value: 0 = ok, non-zero = error; nvr_last_error(h) gives the message.  The library
 *     never calls exit()/abort().
 *   - one handle per device per host thread (thread-compatible, not thread-safe).
 */
#ifndef NVR_B200_H
#define NVR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVR_ABI_VERSION 10
#define NVR_MAX_LEVELS 16
#define NVR_NUM_PARTS 5    /* body, leg, head, larm, rarm -- lib/utils/blend_utils.py:17 */
#define NVR_NUM_JOINTS 24

typedef struct NvrEngine* NvrHandle;

/* One multi-resolution dense+hashed grid: lib/networks/embedders/part_base_embedder.py:13-104.
 * `dense` = Embedder.dense (sum_l<start_hash res_l^3, F); `hash` = Embedder.hash (H, T, F);
 * `bounds` = Embedder.bounds (2,3) -- read on the device at every launch, so the iter-1 bbox
 * overwrite (:107-109) needs no re-binding.  res / size / dense_off are the host-side values of
 * entries_num / entries_size (fp32) / entries_sum[l-1]. */
typedef struct NvrGrid {
    const float* dense;
    const float* hash;
    const float* bounds;
    int32_t n_levels;
    int32_t n_feat;
    int32_t start_hash;
    int32_t sum_features;          /* 1: per-level feature sum (part grids); 0: concat (deformer) */
    int64_t table_size;            /* T = nextprime(2**log2_hashmap_size), :42 */
    int32_t res[NVR_MAX_LEVELS];
    float size[NVR_MAX_LEVELS];
    int64_t dense_off[NVR_MAX_LEVELS];
} NvrGrid;

/* nn.Linear: weight (out, in) row-major, bias (out). */
typedef struct NvrLinear {
    const float* weight;
    const float* bias;
    int32_t in_dim;
    int32_t out_dim;
} NvrLinear;

/* One part network: lib/networks/bw_deform/part_base_network.py:27-63. */
typedef struct NvrPart {
    NvrGrid grid;
    NvrLinear occ[2];              /* 19 -> 64 -> 17 */
    NvrLinear rgb[3];              /* 70 -> 64 [-> 64] -> 3 */
    int32_t n_rgb;                 /* 2 (leg, arms) or 3 (body, head) linears */
    int32_t n_latent;              /* rows of rgb_latent */
    const float* rgb_latent;       /* (n_latent, 8) */
} NvrPart;

typedef struct NvrParams {
    NvrPart part[NVR_NUM_PARTS];
    NvrGrid deformer_grid;         /* lib/networks/deformers/uv_deformer.py:12-21 */
    NvrLinear deformer_mlp[3];     /* 19 -> 32 -> 32 -> 3 */
} NvrParams;

/* Per-frame SMPL tensors, the `batch` keys the path consumes (produced by
 * lib/datasets/h36m/tpose_dataset.py:454-600), without the leading batch dim of 1. */
typedef struct NvrFrame {
    const float* R;                /* (3,3)  world->pose rotation, blend_utils.py:366-382 */
    const float* Th;               /* (3)    */
    const float* pbw;              /* (D,H,W,C) blend-weight volume; channel C-1 = distance to SMPL */
    int32_t pbw_dims[3];
    int32_t pbw_channels;          /* 25 */
    const float* pbounds;          /* (2,3) */
    const float* part_pts;         /* (P, maxlen, 3) posed vertices per part */
    const float* part_pbw;         /* (P, maxlen, 24) */
    const int64_t* lengths2;       /* (P) valid vertices per part */
    int32_t maxlen;
    int32_t _pad0;
    const float* A;                /* (24,4,4) T-pose -> posed */
    const float* big_A;            /* (24,4,4) T-pose -> big pose */
    const float* tuv;              /* (D',H',W',2) */
    int32_t tuv_dims[3];
    int32_t _pad1;
    const float* tbounds;          /* (2,3) */
    const float* frame_dim;        /* (1) f32 */
    const int64_t* latent_index;   /* (1) int64 */
    int64_t topology_key;          /* 0: rebuild the KNN vertex partition for this frame.  Non-zero: the caller
                                      vouches that frames with equal keys list the same vertices in the same
                                      order in part_pts (same subject, same part split), so the partition of the
                                      previous frame is reused and only vertex positions / boxes are refreshed.
                                      The search is exact for ANY partition; the key only trades build time
                                      against box tightness. */
} NvrFrame;

typedef struct NvrConfig {
    int32_t abi_version;           /* NVR_ABI_VERSION */
    int32_t device;                /* CUDA device ordinal */
    float smpl_thresh;             /* cfg.smpl_thresh */
    int32_t mlp_mode;              /* part MLPs: 0 fp32 FFMA tiles; 1 tcgen05 3xTF32 tiles, two tile slots, 256-thread CTAs; 2 the same with two
                                      epilogue warpgroups per tile slot (512-thread CTAs); 3 tcgen05 kind::f16 with fp16-split operands
                                      (hi + lo, three MMAs per product like 3xTF32, fp32-equivalent results), FOUR tile slots per SM */
    uint32_t tune;                 /* NVR_TUNE_* bits: occupancy variants of the same kernels (identical results) */
    uint32_t _pad;
} NvrConfig;
#define NVR_TUNE_WARP_OCC4 1u      /* k_warp without a register cap (127 registers, 4 CTAs/SM) instead of 8 CTAs/SM (<= 64) */
#define NVR_TUNE_KNN_OCC5 2u       /* k_knn at 5 CTAs/SM (<= 48 registers, spills) instead of 4 CTAs/SM (<= 64) */
#define NVR_TUNE_NO_CULL_EARLY_OUT 4u  /* cull: always take the 8-tap lookup (disable the coarse-minimum early-out) */
#define NVR_TUNE_NO_FAR_COLLAPSE 8u    /* evaluate every flagged pair on its own: by default the pairs of a part whose Gaussian
                                          weights sum to < 1e-20 (part farther than ~0.73 m: blended transforms ~1e-12, canonical
                                          point = origin to 5e-13 m) share ONE evaluation per part and frame */
#define NVR_TUNE_DENSE_A1 16u          /* MEASUREMENT variant of SURVEY.md section 8(d) ("a = 1"): the distance cull keeps every sample
                                          and every sample is flagged in exactly ONE part (the part with the smallest weighted
                                          neighbour distance, first on ties), so (sample, part) pairs == samples whatever the input.
                                          Not the reference's semantics; bench.py's `dense_a1` line only */
#define NVR_TUNE_LEVEL_MAJOR 64u        /* experiment: gather a part whose tables are several times the L2 (body: 724 MB) level-major,
                                          one launch per <= 72 MB slice of its tables (identical results; 2.4x less DRAM traffic, no
                                          faster: L2 -> SM row delivery is the same ~6.5 TB/s bound) */
#define NVR_TUNE_WARP_FFMA 128u         /* k_warp with the deformer MLP on the CUDA cores (FFMA2 from broadcast shared-memory loads)
                                          instead of k_warp_tc (tcgen05, fp16-split operands); results agree to ~1e-7 */
#define NVR_TUNE_ONE_LANE 2048u         /* render calls as one chain of passes on the caller's stream (round-2f behaviour) instead of two
                                          lanes: two (or more) passes on two streams of the engine, each on half of the workspace, whose
                                          launch ramps and tails overlap; identical results */
#define NVR_TUNE_SERIAL 32u            /* one stream, no CUDA graph: every launch of a pass back to back on the caller's stream
                                          (what the per-stage CUDA-event timing of nvr_profile needs; nvr_profile(h, 1) implies it) */

/* Device-side work counters of the most recent call's passes (diagnostics / benchmark accounting). */
typedef struct NvrCounters {
    int64_t n_points;              /* samples submitted */
    int64_t n_survivors;           /* samples with pnorm < smpl_thresh */
    int64_t n_pairs[NVR_NUM_PARTS];/* (sample, part) pairs EVALUATED (incl. the one shared far-field pair per part) */
    int64_t n_far_pairs[NVR_NUM_PARTS]; /* flagged pairs answered by the shared far-field pair (NVR_TUNE_NO_FAR_COLLAPSE: 0);
                                      flagged pairs of the reference = n_pairs + n_far_pairs - (n_far_pairs or collapse on ? 1 : 0) */
    int64_t kernel_launches;       /* kernels this handle has launched since creation */
    int64_t n_passes;              /* passes the most recent call ran (a two-lane render: >= 2; the counters above are their
                                      sums, so n_pairs holds one shared far-field pair per part PER PASS) */
} NvrCounters;

int nvr_abi_version(void);
int nvr_create(const NvrConfig* cfg, NvrHandle* out);
int nvr_destroy(NvrHandle h);
const char* nvr_last_error(NvrHandle h);

/* Borrow the nn.Parameter storages (Appendix D of SURVEY.md).  Cheap; call again whenever a
 * storage moved (load_state_dict keeps storages, .cuda()/.to() do not). */
int nvr_bind_params(NvrHandle h, const NvrParams* p);
/* Borrow the per-frame tensors and run the per-frame preparation (distance-channel extraction,
 * vertex packing) on `stream`. */
int nvr_bind_frame(NvrHandle h, const NvrFrame* f, void* stream);

/* Bytes of scratch needed to process up to `max_points` samples in one pass. */
size_t nvr_workspace_bytes(NvrHandle h, int64_t max_points);

/* == Network.forward, eval branch (inb_part_network_multiassign.py:126-168) ==
 * wpts, viewdir (n,3) -> raw (n,4) = [r,g,b,occ], occ (n).  Samples are processed in passes of at
 * most `ws_bytes`-worth of points. */
int nvr_query_points(NvrHandle h, const float* wpts, const float* viewdir, int64_t n,
                     float* raw, float* occ, void* workspace, size_t ws_bytes, void* stream);

/* == Renderer.render, eval branch (inb_renderer.py:15-76,204-239 + net_utils.py:12-44) ==
 * ray_o, ray_d (n_rays,3), near, far (n_rays) -> rgb_map (n_rays,3), acc_map (n_rays);
 * raw (n_rays*n_samples,4) is written only when non-NULL (parity checks). */
int nvr_render_rays(NvrHandle h, const float* ray_o, const float* ray_d, const float* near_,
                    const float* far_, int64_t n_rays, int32_t n_samples, float* rgb_map,
                    float* acc_map, float* raw, void* workspace, size_t ws_bytes, void* stream);

/* Same call with HOST ray buffers and HOST outputs (pinned memory recommended): copies the rays
 * in, renders, copies rgb_map / acc_map back and synchronises `stream`.  `dev_io` is a device
 * scratch of at least n_rays * 48 bytes. */
int nvr_render_rays_host(NvrHandle h, const float* ray_o_host, const float* ray_d_host,
                         const float* near_host, const float* far_host, int64_t n_rays,
                         int32_t n_samples, float* rgb_map_host, float* acc_map_host, void* dev_io,
                         void* workspace, size_t ws_bytes, void* stream);

/* == Multi-GPU frame assembly over NVLink peer memory (the reference renders on one device,
 *    run.py / inb_renderer.py:204-239, so there is no reference counterpart) ==
 * Rays shard over `world` ranks (one process per GPU) in interleaved tiles of `tile` rays: tile t of the frame's n_rays_total
 * rays belongs to rank t % world, and a rank's shard lists its tiles in ascending order.  nvr_frame_create allocates this
 * rank's frame buffer (two slots of n_rays_total x [r, g, b, acc] + barrier flags) and returns its CUDA IPC handle; the host
 * side exchanges the handles over its own control plane (instant_nvr_b200/sharding.py: torch.distributed.all_gather_object)
 * and passes all `world` of them, in rank order, to nvr_frame_connect (entry `rank` is ignored).  world == 1 needs no handles.
 * nvr_render_rays_frame == nvr_render_rays on this rank's shard whose compositing kernel ALSO stores every finished ray to its
 * final position in every rank's frame (NVLink peer stores), followed by one flag barrier: when the work enqueued by the call
 * has run, *frame_out (device pointer, n_rays_total x 4 floats) holds the complete frame on every rank.  It stays valid until
 * the call after the next one (the two slots alternate).  rgb_map / acc_map (the shard's own pixels) may both be NULL.
 * nvr_allgather_frame is the unfused form for pixels that already exist: scatter rgb_map (n,3) / acc_map (n) + barrier.
 * Every rank must make the same sequence of frame calls (the barrier waits for all of them). */
typedef struct NvrIpcHandle { unsigned char bytes[64]; } NvrIpcHandle;
int nvr_frame_create(NvrHandle h, int64_t n_rays_total, int32_t rank, int32_t world, int32_t tile, NvrIpcHandle* handle_out);
int nvr_frame_connect(NvrHandle h, const NvrIpcHandle* handles);
int nvr_frame_disconnect(NvrHandle h);   /* unmap the peers' buffers; EVERY rank must have done this before any rank ... */
int nvr_frame_destroy(NvrHandle h);      /* ... frees its own (the host side puts a barrier between the two) */
int nvr_render_rays_frame(NvrHandle h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                          int64_t n_rays_local, int32_t n_samples, float* rgb_map, float* acc_map,
                          void* workspace, size_t ws_bytes, void* stream, const float** frame_out);
int nvr_allgather_frame(NvrHandle h, const float* rgb_map, const float* acc_map, int64_t n_rays_local, void* stream,
                        const float** frame_out);
/* nvr_render_rays_frame with HOST ray buffers of this rank's shard and HOST outputs for the shard's own pixels (pinned memory
 * recommended), like nvr_render_rays_host: the copies ride on the lanes' streams (the second lane's rays arrive under the first
 * lane's kernels, the first lane's pixels leave under the second's), then the flag barrier and a synchronise of `stream`.
 * `dev_io`: device scratch of at least n_rays_local * 48 bytes. */
int nvr_render_rays_frame_host(NvrHandle h, const float* ray_o_host, const float* ray_d_host, const float* near_host,
                               const float* far_host, int64_t n_rays_local, int32_t n_samples, float* rgb_map_host,
                               float* acc_map_host, void* dev_io, void* workspace, size_t ws_bytes, void* stream,
                               const float** frame_out);

/* == Network.resd (inb_part_network_multiassign.py:122-124; uv_deformer.py:23-45, flag=None) ==
 * canonical points (n,3) -> 0.05*tanh(MLP(grid(u,v,t))) (n,3). */
int nvr_deformer_residual(NvrHandle h, const float* tpts, int64_t n, float* resd, void* stream);

/* Grid embedding of one part (part_base_embedder.py:106-174): xyz (n,3) -> (n,19).  Exposed for
 * per-stage parity tests and the gather roofline measurement. */
int nvr_embed_part(NvrHandle h, int32_t part, const float* xyz, int64_t n, float* out, void* stream);

/* Occ + rgb MLPs of one part on explicit inputs (part_base_network.py:50-60): emb (n, 20-float rows,
 * 19 used), canonical view dirs (n,3) -> raw (n,4) = [rgb, occ].  Per-stage parity tests and the
 * isolated MLP measurement.  Uses batch['latent_index'] of the bound frame. */
int nvr_part_mlp(NvrHandle h, int32_t part, const float* emb, const float* dirs, int64_t n, float* raw,
                 void* workspace, size_t ws_bytes, void* stream);

/* nvr_query_points plus per-stage taps for the parity tests: surv_of_sample (n) = survivor slot or -1,
 * warp_dbg (n,5,8) = [flag, x, y, z, vx, vy, vz, pdist] of every (sample, part) after the warp stage.
 * All n points must fit one pass. */
int nvr_query_points_debug(NvrHandle h, const float* wpts, const float* viewdir, int64_t n, float* raw,
                           int32_t* surv_of_sample, float* warp_dbg, void* workspace, size_t ws_bytes, void* stream);

/* Inference tables (opt-in).  The part networks consume only the per-level SUM of an entry's 16 features
 * (part_base_embedder.py:165), so for forward-only rendering the sums can be taken once per weight update:
 * nvr_prepare_inference(h, 1, stream) builds (dense rows + hash rows) fp32 sums per part (286 MB for inb_377) and
 * makes nvr_query_points / nvr_render_rays* gather 4 B per corner instead of 64 B.  Results differ from the full
 * tables only by fp32 re-association (sum_c w_c sum_f t vs sum_f sum_c w_c t).  The sums are a SNAPSHOT: call
 * again after the tables changed; nvr_bind_params and nvr_prepare_inference(h, 0, ...) drop them.  Training entry
 * points always read the full tables. */
int nvr_prepare_inference(NvrHandle h, int32_t enable, void* stream);

/* ==================== training ====================
 * Network.forward in training mode (inb_part_network_multiassign.py:126-168 with self.training) and its
 * backward.  All n points are processed in ONE pass (the workspace must hold n points) and the workspace
 * must stay untouched between nvr_train_forward and the matching nvr_train_backward: it holds the
 * survivor / pair lists and embeddings the backward needs.
 *
 * Outputs besides raw (n,4) / occ (n): per SURVIVOR, in ascending sample order -- the order of the reference's nonzero()
 * (inb_part_network_multiassign.py:137), rows r < n_survivors (later rows are zero):
 *   x0 (n,5,3)   big-pose point before the residual (= ret['tpts'] rows r*5+p; zeros where the part is unflagged)
 *   resd (n,5,3) deformer residual (= ret['resd']; zeros where unflagged)
 *   tocc (n,5)   per-part occupancy (= ret['tocc']; zeros where unflagged)
 * rank_of_slot (n) int32 receives the survivor-slot -> row map the backward needs (opaque to the caller).
 * The survivor count is read with nvr_read_counters after the call. */
int nvr_train_forward(NvrHandle h, const float* wpts, const float* viewdir, int64_t n, float* raw, float* occ,
                      float* x0, float* resd, float* tocc, int32_t* rank_of_slot,
                      void* workspace, size_t ws_bytes, void* stream);

/* Bytes of extra scratch nvr_train_backward needs for a forward of n points. */
size_t nvr_train_scratch_bytes(NvrHandle h, int64_t n);

/* Backward of nvr_train_forward.  d_raw (n,4): gradient of the scattered raw (add the gradient of `occ` into
 * its 4th column: it is the same number).  d_resd (n_survivors,5,3) and d_tocc (n_survivors,5), the forward's row order, may
 * be NULL.  `x0` and `rank_of_slot` are the forward's outputs.  `grads` has the layout of NvrParams but every pointer is the gradient
 * buffer of that tensor (same shape, fp32, ACCUMULATED into; NULL = not wanted); non-pointer fields are ignored.
 * Gradient flow is the reference's: tables, MLPs, latent row, deformer; none through KNN / LBS / view dirs. */
int nvr_train_backward(NvrHandle h, const float* d_raw, const float* d_resd, const float* d_tocc, const float* x0,
                       const int32_t* rank_of_slot, int64_t n, const NvrParams* grads, void* workspace, size_t ws_bytes,
                       void* scratch, size_t scratch_bytes, void* stream);

/* Training-time sampling along the rays (Renderer.get_wsampling_points, inb_renderer.py:15-31) in one launch:
 * z = near (1 - t) + far t, t = linspace(0, 1, S); with `u` (n_rays, S) -- the reference's torch.rand draw, NULL when
 * cfg.perturb == 0 -- the stratified jitter between the midpoints (:20-27); wpts = ray_o + ray_d z, viewdir = ray_d.
 * Outputs z_vals (n_rays,S), wpts / viewdir (n_rays*S,3); same operation order as the reference's elementwise ops. */
int nvr_train_sample(NvrHandle h, const float* ray_o, const float* ray_d, const float* near_, const float* far_, const float* u,
                     int64_t n_rays, int32_t n_samples, float* z_vals, float* wpts, float* viewdir, void* stream);

/* Distortion regulariser (inb_renderer.py:96-103): loss[r] = sum_ij w_i w_j |m_i - m_j| with the interval midpoints
 * m_k = (z_k + z_{k+1}) / 2, and its backward d_weights[r,i] = 2 d_loss[r] sum_j w_j |m_i - m_j| (z_vals are constants). */
int nvr_distortion_forward(NvrHandle h, const float* weights, const float* z_vals, int64_t n_rays, int32_t n_samples,
                           float* loss, void* stream);
int nvr_distortion_backward(NvrHandle h, const float* weights, const float* z_vals, const float* d_loss, int64_t n_rays,
                            int32_t n_samples, float* d_weights, void* stream);

/* Backward of nvr_deformer_residual (Network.resd on explicit points, used by the pair regulariser):
 * tpts (n,3), d_resd (n,3) -> accumulates into grads->deformer_grid / deformer_mlp. */
int nvr_deformer_backward(NvrHandle h, const float* tpts, const float* d_resd, int64_t n, const NvrParams* grads, void* stream);

/* Alpha compositing on explicit raw (net_utils.py:12-44, epsilon = 0) and its backward: raw (n_rays,S,4) ->
 * weights (n_rays,S), rgb_map (n_rays,3), acc_map (n_rays).  d_weights / d_rgb_map / d_acc_map may be NULL. */
int nvr_composite_forward(NvrHandle h, const float* raw, int64_t n_rays, int32_t n_samples, float* weights,
                          float* rgb_map, float* acc_map, void* stream);
int nvr_composite_backward(NvrHandle h, const float* raw, int64_t n_rays, int32_t n_samples, const float* d_weights,
                           const float* d_rgb_map, const float* d_acc_map, float* d_raw, void* stream);

/* ==================== the steps either side of the path ====================
 *
 * -- Optimizer step (8(f) rank 1).  torch.optim.Adam as lib/train/optimizer.py:13-31 builds it (one param group
 * per tensor, amsgrad off, eps = cfg.train.eps = 1e-15), applied to every listed tensor in ONE pass over HBM:
 * 28 B per element (read p, g, m, v; write p, m, v) against the ~70 B of the library's multi-kernel foreach form.
 * `tensors` is a HOST array; every pointer in it is a device pointer to `numel` contiguous fp32 values.
 * `step` is the 1-based number of THIS update of that tensor (torch keeps one counter per parameter).  The bias
 * corrections are taken in double on the host exactly as torch does (1 - beta**step).  zero_grad != 0 also clears
 * the gradient buffers in the same pass (optimizer.zero_grad(set_to_none=False) for free). */
typedef struct NvrAdamTensor {
    float* param;
    float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t numel;
    int64_t step;
    double lr;
    double weight_decay;
} NvrAdamTensor;
int nvr_adam_step(NvrHandle h, const NvrAdamTensor* tensors, int32_t n_tensors, double beta1, double beta2, double eps,
                  int32_t zero_grad, void* stream);

/* -- Camera rays (8(f) rank 2): get_rays + get_near_far + the mask_at_box compaction, i.e.
 * get_rays_within_bounds / get_rays_within_bounds_coord (lib/utils/if_nerf/if_nerf_data_utils.py:24-38, 92-107,
 * 329-362), which the reference runs in numpy on DataLoader workers for every frame (tpose_dataset.py:438,
 * tpose_novel_view_dataset.py:206).  Host inputs: K_inv = inv(K) (3,3), R (3,3), T (3) row-major doubles (the
 * reference's get_rays works in float64).  `bounds` (2,3) fp32 is a DEVICE pointer (world-space bbox).
 * Outputs (device, capacity H*W rays): ray_o, ray_d (n,3), near, far (n), coord (n) int32 = pixel index row*W + col of
 * each surviving ray in row-major order (may be NULL), mask_at_box (H*W) uint8, n_rays (1) int32.
 * `workspace` needs nvr_rays_workspace_bytes(H, W) bytes. */
size_t nvr_rays_workspace_bytes(int32_t H, int32_t W);
int nvr_generate_rays(NvrHandle h, int32_t H, int32_t W, const double* K_inv_host, const double* R_host, const double* T_host,
                      const float* bounds, float* ray_o, float* ray_d, float* near_, float* far_, int32_t* coord,
                      uint8_t* mask_at_box, int32_t* n_rays, void* workspace, size_t ws_bytes, void* stream);

/* -- Image assembly and the evaluator's MSE (8(f) rank 4; lib/evaluators/if_nerf.py:28-31, 84-113):
 * img (n_pixels,3) = 0, img[coord[r]] = rgb[r];   sq_sum[0] = sum_i ((double)a[i] - (double)b[i])^2  (the caller divides
 * by n and takes -10 log10 for the PSNR).  sq_sum is a device double, overwritten. */
int nvr_assemble_image(NvrHandle h, const float* rgb, const int32_t* coord, int64_t n_rays, int64_t n_pixels, float* img, void* stream);
int nvr_sq_diff_sum(NvrHandle h, const float* a, const float* b, int64_t n, double* sq_sum, void* stream);
/* The evaluator's SSIM (lib/evaluators/if_nerf.py:33-74): skimage 0.19.3 structural_similarity(multichannel=True) on the
 * float64 crop rows [y0, y0+h) x columns [x0, x0+w) (cv2.boundingRect of mask_at_box) of two (H, W, 3) fp32 images -- uniform
 * 7 x 7 windows, sample covariance, data_range 2 (skimage's dtype range of float images), interior only.  sums (3 device
 * doubles, overwritten) = per-channel sums of S over the (h-6) x (w-6) interior; ssim = mean_c sums[c] / ((h-6)(w-6)). */
int nvr_ssim_sums(NvrHandle h, const float* img_a, const float* img_b, int32_t H, int32_t W, int32_t x0, int32_t y0,
                  int32_t w, int32_t hh, double* sums, void* stream);

/* -- Per-frame SMPL preprocessing (8(f) rank 3): what the reference's dataset computes in numpy for every frame
 * (lib/datasets/h36m/tpose_dataset.py:247-293 prepare_input, :570-600 part tables; get_rigid_transformation,
 * lib/utils/if_nerf/if_nerf_data_utils.py:523-577; get_bounds :689-696) and what its offline tool pre-bakes to
 * lbs/bweights/{frame}.npy (tools/prepare_zjumocap.py:474-508 get_bweights, :152-165 get_grid_points), on the device, so a
 * frame goes from (pose parameters, world vertices) to a bound NvrFrame without touching the host or the disk.
 *
 * NvrSmplPose is HOST data (the numbers of one new_params/{frame}.npy + the subject's joints / parents); they travel as
 * kernel parameters.  poses / big_poses are axis-angles (24,3); joints is lbs/joints.npy as float32 (the reference does
 * the parent-relative subtraction in float32, :554-555); parents[0] is ignored.
 * Device inputs: wxyz (V,3) fp32 world vertices (new_vertices/{frame}.npy); vert_slot (V) int32 = part * maxlen + rank of
 * the vertex inside its part (static per subject; -1 = in no part) -- may be NULL together with part_pts.
 * Device outputs (any may be NULL): R (3,3), Th (3), A / big_A (24,4,4), ppts (V,3), part_pts (P,maxlen,3) -- rows
 * beyond a part's length are NOT written (zero them once), pbounds / wbounds (2,3) = bbox(ppts / wxyz) -+ box_padding.
 * The workspace (nvr_smpl_workspace_bytes(V) bytes, 256-byte aligned) also keeps the float64 posed vertices the volume is
 * built from; pass the same one to the two calls below.
 *
 * nvr_smpl_volume_dims SYNCHRONISES the stream: it reads the float64 bbox back and returns the volume's dims
 * (D,H,W) = lengths of np.arange(min - 0.05, max + 0.05 + 0.025, 0.025) per axis and their first elements origin[3].
 * nvr_smpl_bweights fills pbw (D,H,W,25) fp32: the 24 skinning weights (`weights` (V,24), device) of the posed vertex
 * nearest to each voxel centre and the distance to it, both from float64 arithmetic (psbody closest_vertices(use_cgal=True)
 * = exact nearest vertex; lowest index on exact ties). */
typedef struct NvrSmplPose {
    double Rh[3];
    double Th[3];
    double poses[72];
    double big_poses[72];
    float joints[72];
    int32_t parents[24];
} NvrSmplPose;
typedef struct NvrSmplOut {
    float* R;
    float* Th;
    float* A;
    float* big_A;
    float* ppts;
    float* part_pts;
    float* pbounds;
    float* wbounds;
} NvrSmplOut;
size_t nvr_smpl_workspace_bytes(int32_t n_verts);
int nvr_smpl_pose_frame(NvrHandle h, const NvrSmplPose* pose_host, const float* wxyz, int32_t n_verts, const int32_t* vert_slot,
                        int32_t maxlen, float box_padding, const NvrSmplOut* out, void* workspace, size_t ws_bytes, void* stream);
int nvr_smpl_volume_dims(NvrHandle h, const void* workspace, int32_t dims[3], double origin[3], void* stream);
int nvr_smpl_bweights(NvrHandle h, const void* workspace, int32_t n_verts, const float* weights, const int32_t dims[3],
                      const double origin[3], float* pbw, void* stream);

/* Per-stage device timing.  nvr_profile(h, 1) clears the accumulators and makes every later launch
 * record a CUDA-event pair on its stream; nvr_profile_read synchronises the device and sums them. */
#define NVR_STAGE_PREP 0
#define NVR_STAGE_CULL 1
#define NVR_STAGE_KNN 2       /* 4-NN search + blend-weight flags */
#define NVR_STAGE_WARP 3      /* LBS + deformer */
#define NVR_STAGE_EMBED 4     /* the grid gather */
#define NVR_STAGE_MLP 5
#define NVR_STAGE_RESOLVE 6
#define NVR_NUM_STAGES 7
typedef struct NvrStageProfile {
    double ms[NVR_NUM_STAGES];            /* summed launch durations */
    int64_t launches[NVR_NUM_STAGES];
    int64_t passes;
    int64_t survivors;                    /* summed over the profiled passes */
    int64_t pairs[NVR_NUM_PARTS];         /* evaluated pairs (what the gather and the MLPs processed) */
    int64_t far_pairs[NVR_NUM_PARTS];     /* pairs answered by the shared far-field pair */
    double embed_part_ms[NVR_NUM_PARTS];  /* NVR_STAGE_EMBED / NVR_STAGE_MLP split by part (one launch per part and pass) */
    double mlp_part_ms[NVR_NUM_PARTS];
} NvrStageProfile;
int nvr_profile(NvrHandle h, int32_t enable);
int nvr_profile_read(NvrHandle h, NvrStageProfile* out);

/* Measurement aid for the gather's roofline (bench.py): how many DISTINCT 32-byte sectors of each part's dense + hash
 * tables the pair lists of the most recent pass touch (16 levels x 8 corners x 2 sectors per pair, duplicates counted
 * once) -- the compulsory table traffic of that pass, against SURVEY.md 8(d)'s no-reuse figure of 8192 B per pair.
 * Re-walks the pair lists left in `workspace` by the last single-pass nvr_query_points / nvr_render_rays call with the
 * gather's own index arithmetic, marking a bitmap; synchronises `stream`.  unique_sectors_host: 5 int64. */
int nvr_gather_footprint(NvrHandle h, void* workspace, size_t ws_bytes, int64_t* unique_sectors_host, void* stream);

/* Copies the device counters of the last pass to the host (synchronises `stream`). */
int nvr_read_counters(NvrHandle h, NvrCounters* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NVR_B200_H */

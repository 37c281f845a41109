// nvr_cabi.cu -- extern "C" surface of libnvr_b200.so (declared in include/nvr_b200.h).
//
// Host-side orchestration only: pointer bookkeeping, pass splitting, launches.  There is no CPU
// compute path here; every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/nvr_b200.h"
#include "nvr_kernels.cuh"
#include "nvr_smpl.cuh"
#include "nvr_mlp_tc.cuh"
#include "nvr_mlp_f16.cuh"
#include "nvr_warp_tc.cuh"
#include "nvr_train.cuh"
#include "nvr_aux.cuh"

struct NvrEngine {
    NvrConfig cfg;
    std::string err;
    int sm_count = 148;
    bool have_params = false, have_frame = false;
    NvrParams params;
    GridDev part_grid[NVR_PARTS];
    GridDev def_grid;
    PartMlpDev part_mlp[NVR_PARTS];
    DeformerMlp def_mlp;
    NvrFrame frame;
    FrameDev fdev;
    // per-frame device buffers owned by the engine (grow-only)
    float* d_dist = nullptr; size_t dist_cap = 0;        // compact distance volume, followed by its coarse minimum grid
    float4* d_verts = nullptr; size_t verts_cap = 0;      // clustered vertices + 2 AABB corners per cluster
    int* d_cl_off = nullptr;
    int* d_perm = nullptr; size_t perm_cap = 0;            // KD partition: cluster slot -> vertex index in its part
    long long perm_key = 0; int perm_maxlen = 0;           // topology the partition was built for
    PartMlpDev* d_part_mlp = nullptr;       // device copy of part_mlp[] for k_mlp_prep
    GridDev* d_part_grid = nullptr;         // device copy of part_grid[] for k_embed_parts
    float* d_mlp_blocks = nullptr;          // NVR_PARTS packed tcgen05 parameter blocks (mlp_mode 1)
    float* d_presum = nullptr; size_t presum_cap = 0;     // inference tables: per part [dense rows | hash rows] sums
    size_t presum_off[NVR_PARTS + 1] = {0};
    bool presum_valid = false;
    int* d_counters_snapshot = nullptr;     // last pass's counters (two-lane render: the call's passes summed), for nvr_read_counters
    bool snapshot_accumulate = false;
    cudaStream_t part_stream[NVR_PARTS] = {nullptr};   // training backward: the five parts' chains run side by side
    cudaEvent_t ev_fork = nullptr, ev_join[NVR_PARTS] = {nullptr};
    long long launches = 0;
    long long last_points = 0;
    long long last_passes = 1;
    long long two_lane_min = 1ll << 19;     // render calls of at least this many samples run as two lanes (NVR_TWO_LANE_MIN_SAMPLES)
    int last_lanes = 1;
    int last_lanes_used = 1;                // how many of the two half workspaces a pass ran in (a one-ray call: only the first)                     // 2: the most recent call carved its workspace into two halves (two-lane render)
    // multi-GPU frame assembly (nvr_frame.cuh): the local buffer [flags | slot 0 | slot 1] and the peers' mappings
    struct PeerFrame {
        void* local = nullptr; void* peer[NVR_MAX_RANKS] = {nullptr};
        long long n_total = 0; int rank = 0, world = 0, tile = 1024; unsigned int epoch = 0; bool connected = false;
    } pf;
    // optional per-stage timing (nvr_profile): event pairs around every launch + per-pass counter snapshots
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int stage; int part; size_t e0, e1; };
    std::vector<Span> spans;
    int* h_pass_counters = nullptr;         // pinned [PROF_MAX_PASSES][NVR_CTR_WORDS]
    int n_pass_snap = 0;
};
static const int PROF_MAX_PASSES = 8192;

struct StageTimer {                         // RAII: records an event pair on `st` when profiling is on
    NvrEngine* h; cudaStream_t st; int stage; int part; size_t e0 = 0; bool on;
    StageTimer(NvrEngine* h_, cudaStream_t st_, int stage_, int part_ = -1) : h(h_), st(st_), stage(stage_), part(part_), on(h_->profiling) {
        if (!on) return;
        if (h->ev_used + 2 > h->ev_pool.size()) {
            for (int i = 0; i < 64; ++i) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
        }
        e0 = h->ev_used; h->ev_used += 2;
        cudaEventRecord(h->ev_pool[e0], st);
    }
    ~StageTimer() {
        if (!on) return;
        cudaEventRecord(h->ev_pool[e0 + 1], st);
        h->spans.push_back({stage, part, e0, e0 + 1});
    }
};

#define NVR_CHECK(h, expr)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                        \
            return 2;                                                                             \
        }                                                                                         \
    } while (0)

static int fail(NvrEngine* h, const char* msg) {
    if (h) h->err = msg;
    return 1;
}

static GridDev to_dev(const NvrGrid& g) {
    GridDev d;
    memset(&d, 0, sizeof(d));
    d.dense = g.dense; d.hash = g.hash; d.bounds = g.bounds;
    d.n_levels = g.n_levels; d.n_feat = g.n_feat; d.start_hash = g.start_hash; d.sum_features = g.sum_features;
    d.T = (unsigned long long)g.table_size;
    d.T_magic = g.table_size > 1 ? (unsigned long long)((((unsigned __int128)1) << 64) / (unsigned __int128)g.table_size) : 0ull;
    int max_res = 0;
    for (int l = 0; l < NVR_MAX_LEVELS; ++l) {
        d.res[l] = g.res[l]; d.size[l] = g.size[l]; d.dense_off[l] = g.dense_off[l];
        if (l < g.n_levels) max_res = std::max(max_res, g.res[l]);
    }
    // 32-bit modulo (nvr_mod_T40) needs hash values < 2^40 (coordinates < 2^13) and 2^8 < T < 2^31
    d.T_magic40 = (max_res <= 8192 && g.table_size > 256 && g.table_size < (1ll << 31))
                      ? (unsigned int)((1ull << 40) / (unsigned long long)g.table_size) : 0u;
    return d;
}
static LinearDev to_dev(const NvrLinear& l) { return LinearDev{l.weight, l.bias, l.in_dim, l.out_dim}; }

static void frame_release(NvrEngine* h);

extern "C" int nvr_abi_version(void) { return NVR_ABI_VERSION; }

extern "C" int nvr_create(const NvrConfig* cfg, NvrHandle* out) {
    if (!cfg || !out) return 1;
    *out = nullptr;
    if (cfg->abi_version != NVR_ABI_VERSION) return 3;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return 4;                                        // no CUDA device: there is no CPU fallback
    }
    NvrEngine* h = new NvrEngine();
    h->cfg = *cfg;
    if (const char* e = getenv("NVR_TWO_LANE_MIN_SAMPLES")) {          // tests / sanitizer runs: two lanes on small renders
        const long long v = atoll(e);
        if (v >= 128) h->two_lane_min = v;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(cfg->device) != cudaSuccess || cudaMalloc(&h->d_cl_off, 8 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h->d_counters_snapshot, NVR_CTR_WORDS * sizeof(int)) != cudaSuccess ||
        cudaMemset(h->d_counters_snapshot, 0, NVR_CTR_WORDS * sizeof(int)) != cudaSuccess ||
        cudaFuncSetAttribute(k_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, MLP_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_cluster_verts, cudaFuncAttributeMaxDynamicSharedMemorySize, NVR_CLUSTER_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_mlp_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_mlp_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_mlp_f16, cudaFuncAttributeMaxDynamicSharedMemorySize, F16_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_deformer_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_SMEM_BYTES) != cudaSuccess ||
        cudaMalloc(&h->d_part_mlp, NVR_PARTS * sizeof(PartMlpDev)) != cudaSuccess ||
        cudaMalloc(&h->d_part_grid, NVR_PARTS * sizeof(GridDev)) != cudaSuccess ||
        cudaMalloc(&h->d_mlp_blocks, (size_t)NVR_PARTS * TC_BLOCK_FLOATS * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        delete h;
        return 5;
    }
    for (int p = 0; p < NVR_PARTS; ++p)
        if (cudaStreamCreateWithFlags(&h->part_stream[p], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_join[p], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); delete h; return 5; }
    if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); delete h; return 5; }
    *out = h;
    return 0;
}

extern "C" int nvr_destroy(NvrHandle h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->d_dist); cudaFree(h->d_verts); cudaFree(h->d_cl_off); cudaFree(h->d_perm); cudaFree(h->d_part_mlp); cudaFree(h->d_part_grid); cudaFree(h->d_mlp_blocks); cudaFree(h->d_presum); cudaFree(h->d_counters_snapshot);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (int p = 0; p < NVR_PARTS; ++p) { if (h->part_stream[p]) cudaStreamDestroy(h->part_stream[p]); if (h->ev_join[p]) cudaEventDestroy(h->ev_join[p]); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->h_pass_counters) cudaFreeHost(h->h_pass_counters);
    frame_release(h);
    delete h;
    return 0;
}

extern "C" const char* nvr_last_error(NvrHandle h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int nvr_bind_params(NvrHandle h, const NvrParams* p) {
    if (!h || !p) return fail(h, "nvr_bind_params: null argument");
    for (int i = 0; i < NVR_NUM_PARTS; ++i) {
        const NvrPart& pt = p->part[i];
        const NvrGrid& g = pt.grid;
        if (g.n_levels < 1 || g.n_levels > NVR_MAX_LEVELS || g.n_feat != 16 || !g.sum_features || g.start_hash < 1)
            return fail(h, "nvr_bind_params: part grid must be F=16, per-level feature sum, 1..16 levels, >=1 dense level");
        if (pt.occ[0].in_dim != 19 || pt.occ[0].out_dim != 64 || pt.occ[1].in_dim != 64 || pt.occ[1].out_dim != 17)
            return fail(h, "nvr_bind_params: occ MLP must be 19->64->17");
        if (pt.n_rgb != 2 && pt.n_rgb != 3) return fail(h, "nvr_bind_params: rgb MLP must have 2 or 3 linears");
        if (pt.rgb[0].in_dim != 70 || pt.rgb[0].out_dim != 64 || pt.rgb[pt.n_rgb - 1].in_dim != 64 ||
            pt.rgb[pt.n_rgb - 1].out_dim != 3 || (pt.n_rgb == 3 && (pt.rgb[1].in_dim != 64 || pt.rgb[1].out_dim != 64)))
            return fail(h, "nvr_bind_params: rgb MLP must be 70->64[->64]->3");
        if (!g.dense || !g.hash || !g.bounds || !pt.rgb_latent) return fail(h, "nvr_bind_params: null table pointer");
        if (g.table_size < 2 || (long long)(g.n_levels - g.start_hash) * g.table_size >= (1ll << 32))
            return fail(h, "nvr_bind_params: hashed levels must have fewer than 2^32 rows in total");
        for (int l = 0; l < g.start_hash && l < g.n_levels; ++l)
            if (g.dense_off[l] < 0 || g.dense_off[l] + (long long)g.res[l] * g.res[l] * g.res[l] >= (1ll << 31))
                return fail(h, "nvr_bind_params: dense levels must have fewer than 2^31 rows in total");
    }
    const NvrGrid& dg = p->deformer_grid;
    if (dg.n_feat != 2 || dg.sum_features || dg.n_levels != 8 || dg.start_hash < 1)
        return fail(h, "nvr_bind_params: deformer grid must be 8 levels x F=2, concat");
    if (p->deformer_mlp[0].in_dim != 19 || p->deformer_mlp[0].out_dim != 32 || p->deformer_mlp[1].in_dim != 32 ||
        p->deformer_mlp[1].out_dim != 32 || p->deformer_mlp[2].in_dim != 32 || p->deformer_mlp[2].out_dim != 3)
        return fail(h, "nvr_bind_params: deformer MLP must be 19->32->32->3");
    h->params = *p;
    for (int i = 0; i < NVR_NUM_PARTS; ++i) {
        const NvrPart& pt = p->part[i];
        h->part_grid[i] = to_dev(pt.grid);
        PartMlpDev& m = h->part_mlp[i];
        m.occ[0] = to_dev(pt.occ[0]); m.occ[1] = to_dev(pt.occ[1]);
        for (int k = 0; k < 3; ++k) m.rgb[k] = to_dev(pt.rgb[k < pt.n_rgb ? k : 0]);
        m.n_rgb = pt.n_rgb; m.n_latent = pt.n_latent; m.latent = pt.rgb_latent;
    }
    h->def_grid = to_dev(dg);
    h->def_mlp = DeformerMlp{p->deformer_mlp[0].weight, p->deformer_mlp[0].bias, p->deformer_mlp[1].weight,
                             p->deformer_mlp[1].bias, p->deformer_mlp[2].weight, p->deformer_mlp[2].bias};
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    NVR_CHECK(h, cudaMemcpy(h->d_part_mlp, h->part_mlp, sizeof(h->part_mlp), cudaMemcpyHostToDevice));
    NVR_CHECK(h, cudaMemcpy(h->d_part_grid, h->part_grid, sizeof(h->part_grid), cudaMemcpyHostToDevice));
    h->have_params = true;
    h->presum_valid = false;                 // new storages: the inference tables must be rebuilt
    return 0;
}

extern "C" int nvr_bind_frame(NvrHandle h, const NvrFrame* f, void* stream_) {
    if (!h || !f) return fail(h, "nvr_bind_frame: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!f->R || !f->Th || !f->pbw || !f->pbounds || !f->part_pts || !f->part_pbw || !f->lengths2 || !f->A ||
        !f->big_A || !f->tuv || !f->tbounds || !f->frame_dim || !f->latent_index)
        return fail(h, "nvr_bind_frame: null tensor pointer");
    if (f->pbw_channels < 1 || f->maxlen < 1) return fail(h, "nvr_bind_frame: bad pbw_channels / maxlen");
    if (f->maxlen > NVR_SORT_MAX) return fail(h, "nvr_bind_frame: a part with more than 8192 vertices is not supported");
    for (int a = 0; a < 3; ++a)
        if (f->pbw_dims[a] < 1 || f->tuv_dims[a] < 1) return fail(h, "nvr_bind_frame: bad volume dims");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    const size_t n_vox = (size_t)f->pbw_dims[0] * f->pbw_dims[1] * f->pbw_dims[2];
    const size_t max_cl = (size_t)NVR_NUM_PARTS * ((f->maxlen + NVR_CL - 1) / NVR_CL);
    const size_t n_verts = max_cl * NVR_CL + 2 * max_cl;   // float4 slots: vertices, then cl_lo, then cl_hi
    const size_t n_coarse = (size_t)nvr_coarse_dim(f->pbw_dims[0]) * nvr_coarse_dim(f->pbw_dims[1]) * nvr_coarse_dim(f->pbw_dims[2]);
    if (n_vox + n_coarse > h->dist_cap) {
        NVR_CHECK(h, cudaFree(h->d_dist));
        h->d_dist = nullptr; h->dist_cap = 0;
        NVR_CHECK(h, cudaMalloc(&h->d_dist, (n_vox + n_coarse) * sizeof(float)));
        h->dist_cap = n_vox + n_coarse;
    }
    if (n_verts > h->verts_cap) {
        NVR_CHECK(h, cudaFree(h->d_verts));
        h->d_verts = nullptr; h->verts_cap = 0;
        NVR_CHECK(h, cudaMalloc(&h->d_verts, n_verts * sizeof(float4)));
        h->verts_cap = n_verts;
    }
    if (max_cl * NVR_CL > h->perm_cap) {
        NVR_CHECK(h, cudaFree(h->d_perm));
        h->d_perm = nullptr; h->perm_cap = 0; h->perm_key = 0;
        NVR_CHECK(h, cudaMalloc(&h->d_perm, max_cl * NVR_CL * sizeof(int)));
        h->perm_cap = max_cl * NVR_CL;
    }
    StageTimer tm_(h, stream, NVR_STAGE_PREP);
    float* d_cmin = h->d_dist + n_vox;
    float4* cl_lo = h->d_verts + max_cl * NVR_CL;
    float4* cl_hi = cl_lo + max_cl;
    if (f->topology_key == 0 || f->topology_key != h->perm_key || f->maxlen != h->perm_maxlen) {
        k_cluster_verts<<<NVR_NUM_PARTS, 1024, NVR_CLUSTER_SMEM, stream>>>(f->part_pts, (const long long*)f->lengths2, f->maxlen,
                                                                          h->d_perm, h->d_cl_off);
        h->perm_key = f->topology_key; h->perm_maxlen = f->maxlen;
        h->launches++;
    }
    {   // distance channel copy | coarse minimum grid | cluster re-posing: one launch (k_frame_all)
        const int nb_prep = std::max<int>(1, std::min<int>(h->sm_count * 2, (int)((n_vox + 511) / 512)));
        const int nb_coarse = std::max<int>(1, std::min<int>(h->sm_count * 4, (int)((n_coarse + 3) / 4)));
        const int nb_apply = std::max<int>(1, (int)((max_cl * NVR_CL + 127) / 128));
        k_frame_all<<<nb_prep + nb_coarse + nb_apply, 128, 0, stream>>>(f->pbw, f->pbw_channels, f->pbw_dims[0], f->pbw_dims[1], f->pbw_dims[2],
                                                                       h->d_dist, d_cmin, f->part_pts, f->maxlen, h->d_perm, h->d_cl_off,
                                                                       h->d_verts, cl_lo, cl_hi, nb_prep, nb_coarse);
    }
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 1;
    h->frame = *f;
    FrameDev& d = h->fdev;
    d.R = f->R; d.Th = f->Th;
    d.dist = VolumeDev{h->d_dist, f->pbw_dims[0], f->pbw_dims[1], f->pbw_dims[2], 1, f->pbounds};
    d.dist_cmin = (h->cfg.tune & NVR_TUNE_NO_CULL_EARLY_OUT) ? nullptr : d_cmin;
    d.tuv = VolumeDev{f->tuv, f->tuv_dims[0], f->tuv_dims[1], f->tuv_dims[2], 2, f->tbounds};
    d.verts = h->d_verts; d.cl_lo = cl_lo; d.cl_hi = cl_hi; d.cl_off = h->d_cl_off; d.part_pbw = f->part_pbw; d.maxlen = f->maxlen;
    d.A = f->A; d.bigA = f->big_A; d.frame_dim = f->frame_dim; d.latent_index = (const long long*)f->latent_index;
    h->have_frame = true;
    return 0;
}

// ---- workspace ---------------------------------------------------------------------------
static const size_t WS_HEADER = 256;
static const size_t WS_PER_POINT = sizeof(int) + sizeof(float4) + NVR_NUM_PARTS * sizeof(PairRec) +
                                   NVR_NUM_PARTS * NVR_EMB_STRIDE * sizeof(float) + NVR_NUM_PARTS * sizeof(float4);

extern "C" size_t nvr_workspace_bytes(NvrHandle, int64_t max_points) {
    if (max_points < 1) max_points = 1;
    const size_t pts = ((size_t)max_points + 63) & ~(size_t)63;
    // + 128: the reserved far-field slots and rounding; + 1024: a two-lane render (render_rays_impl) carves the buffer into two
    // halves, each with its own header, reserved slots and up to one ray (<= 256 samples) more than half of the points
    return WS_HEADER + (pts + 128 + 1024) * WS_PER_POINT;
}

struct Workspace {
    int* counters; int* surv_of_sample; float4* surv; PairRec* pairs; float* emb; float4* raws;
    long long cap;                          // stride of every per-sample / per-pair array
    long long pts;                          // samples one pass may hold: cap - 64 (the last survivor slot and one
                                            // pair record per part belong to the shared far-field pairs, NVR_FAR_WSUM)
};
static bool carve(void* ws, size_t bytes, Workspace& w) {
    if (!ws || bytes < WS_HEADER + 192 * WS_PER_POINT || ((uintptr_t)ws & 255)) return false;
    long long cap = (long long)((bytes - WS_HEADER) / WS_PER_POINT) - 64;
    cap = std::min<long long>(cap & ~63ll, 1ll << 30);
    if (cap < 128) return false;
    char* p = (char*)ws;
    w.counters = (int*)p; p += WS_HEADER;
    w.surv = (float4*)p; p += cap * sizeof(float4);
    w.pairs = (PairRec*)p; p += cap * NVR_NUM_PARTS * sizeof(PairRec);
    w.raws = (float4*)p; p += cap * NVR_NUM_PARTS * sizeof(float4);
    w.emb = (float*)p; p += cap * NVR_NUM_PARTS * NVR_EMB_STRIDE * sizeof(float);
    w.surv_of_sample = (int*)p;
    w.cap = cap;
    w.pts = cap - 64;
    return true;
}

static long long dense_rows(const NvrGrid& g) {
    long long r = 0;
    for (int l = 0; l < g.start_hash && l < g.n_levels; ++l) r += (long long)g.res[l] * g.res[l] * g.res[l];
    return r;
}
static long long hash_rows(const NvrGrid& g) { return (long long)(g.n_levels - g.start_hash) * g.table_size; }

static int grid_for(long long items, int per_block, int max_blocks) {
    long long b = (items + per_block - 1) / per_block;
    return (int)std::max<long long>(1, std::min<long long>(b, max_blocks));
}

// Level ranges one part's gather is launched in: ONE launch over all levels, unless NVR_TUNE_LEVEL_MAJOR asks for the
// level-major experiment -- a part whose tables are several times the L2 (body: 724 MB) gathered in slices of consecutive
// levels of at most L2_SLICE_BYTES, one sweep over the pair list per slice.  Measured (profiles/r2a): the body part's DRAM
// reads drop from 4.5 GB to 1.8 GB per frame but its time does not (0.84 against 0.79 ms): a 67 MB hashed level still misses
// the L2 half of the time, and L2 -> SM delivery of 64-byte rows tops out at the same ~6.5 TB/s as HBM does.  Kept opt-in.
static const long long L2_SLICE_BYTES = 72ll << 20;
static const long long L2_SPLIT_ABOVE_BYTES = 384ll << 20;   // ~3x the 126 MB L2 (leg: 290 MB, DRAM traffic already ~1.5x its tables)
static int embed_plan(const NvrEngine* h, int p, int begin[NVR_MAX_LEVELS], int end[NVR_MAX_LEVELS]) {
    const NvrGrid& g = h->params.part[p].grid;
    const long long total = (dense_rows(g) + hash_rows(g)) * 64;
    if (!(h->cfg.tune & NVR_TUNE_LEVEL_MAJOR) || total <= L2_SPLIT_ABOVE_BYTES) { begin[0] = 0; end[0] = g.n_levels; return 1; }
    int n = 0, lb = 0;
    long long acc = 0;
    for (int l = 0; l < g.n_levels; ++l) {
        const long long bytes = (l < g.start_hash ? (long long)g.res[l] * g.res[l] * g.res[l] : g.table_size) * 64;
        if (l > lb && acc + bytes > L2_SLICE_BYTES) { begin[n] = lb; end[n] = l; ++n; lb = l; acc = 0; }
        acc += bytes;
    }
    begin[n] = lb; end[n] = g.n_levels;
    return n + 1;
}

// tensor-core part MLPs over one part's pair list; mlp_mode 3: fp16-split operands, four tile slots (nvr_mlp_f16.cuh);
// 1 / 2: 3xTF32, two tile slots with one / two epilogue warpgroups each (nvr_mlp_tc.cuh)
static size_t mlp_block_floats(const NvrEngine* h) { return h->cfg.mlp_mode == 3 ? (size_t)F16_BLOCK_FLOATS : (size_t)TC_BLOCK_FLOATS; }
static void launch_mlp_prep(NvrEngine* h, cudaStream_t st) {
    if (h->cfg.mlp_mode == 3) k_mlp_prep16<<<dim3(NVR_NUM_PARTS, F16_PREP_SPLIT), 256, 0, st>>>(h->d_part_mlp, h->fdev.latent_index, h->d_mlp_blocks);
    else k_mlp_prep<<<NVR_NUM_PARTS, 256, 0, st>>>(h->d_part_mlp, h->fdev.latent_index, h->d_mlp_blocks);
}
static void launch_mlp_tc(NvrEngine* h, int grid, const float* blk, int n_rgb, int part, const int* count, const PairRec* pl,
                          const float* el, float4* raws, int out_stride, cudaStream_t st) {
    if (h->cfg.mlp_mode == 3) {
        MlpBatch mb;
        memset(&mb, 0, sizeof(mb));
        mb.blk[0] = blk; mb.count[0] = count; mb.pl[0] = pl; mb.el[0] = el; mb.n_rgb[0] = n_rgb; mb.out_part[0] = part; mb.n_parts = 1;
        k_mlp_f16<<<grid, F16_THREADS, F16_SMEM_BYTES, st>>>(mb, raws, out_stride);
    } else if (h->cfg.mlp_mode == 2)
        k_mlp_tc<2><<<grid, TC_THREADS(2), TC_SMEM_BYTES, st>>>(blk, n_rgb, part, count, pl, el, raws, out_stride);
    else
        k_mlp_tc<1><<<grid, TC_THREADS(1), TC_SMEM_BYTES, st>>>(blk, n_rgb, part, count, pl, el, raws, out_stride);
}

// One pass over `n` samples (n <= ws.pts): cull -> warp -> 5x(embed, mlp).  The caller resolves.
// Far-field pairs share one evaluation per part (NVR_FAR_WSUM) unless the caller needs every pair's own record
// (per-stage debug output, training) or NVR_TUNE_NO_FAR_COLLAPSE is set.
static bool far_collapse(const NvrEngine* h, const float* dbg, const float* out_x0) {
    return !(h->cfg.tune & (NVR_TUNE_NO_FAR_COLLAPSE | NVR_TUNE_DENSE_A1)) && !dbg && !out_x0;
}
static const float4* far_raws(const NvrEngine* h, const Workspace& w, bool on) {
    return on ? w.raws + (w.cap - 1) * NVR_NUM_PARTS : nullptr;
}
static int run_pass(NvrEngine* h, const Workspace& w, const float* pts, const float* ray_d, const float* near_,
                    const float* far_, long long n, int n_samples, const float* dirs, int dir_div, cudaStream_t st,
                    float* dbg = nullptr, float* out_x0 = nullptr, float* out_resd = nullptr, bool full_tables = false,
                    int* rank_of_slot = nullptr, bool mlp_prep_done = false) {
    const int sm = h->sm_count;
    const bool dense_a1 = (h->cfg.tune & NVR_TUNE_DENSE_A1) && !dbg && !out_x0;   // measurement variant (a = 1)
    NVR_CHECK(h, cudaMemsetAsync(w.counters, 0, NVR_CTR_WORDS * sizeof(int), st));
    // KNN unit descriptors (k_cull -> k_knn) borrow the head of the pair lists, which k_warp only writes after k_knn is done:
    // at most one unit per 32 survivors + 4 per 2048-position span, 8 B each, against 160 B of pair records per sample
    int2* units = (int2*)w.pairs;
    { StageTimer t(h, st, NVR_STAGE_CULL);
    NVR_CHECK(h, cudaMemsetAsync(w.surv_of_sample, 0xFF, (size_t)n * sizeof(int), st));   // -1 = culled; k_cull fills in the survivors
    k_cull<<<grid_for(n, 256 * CULL_T, sm * 8), 256, 0, st>>>(h->fdev, pts, ray_d, near_, far_, n, n_samples, h->cfg.smpl_thresh,
                                                      w.counters, w.surv_of_sample, w.surv, dense_a1 ? 1 : 0, units); }
    if (rank_of_slot) {      // training: survivors' positions in ascending sample order (the order the reference returns them in)
        k_rank_slots<<<1, 1024, 0, st>>>(w.surv_of_sample, n, rank_of_slot);
        h->launches++;
    }
    // neighbour records alias the embedding buffer: they are consumed by k_warp before k_embed writes it
    KnnRec* recs = (KnnRec*)w.emb;
    const int far_slot = far_collapse(h, dbg, out_x0) ? (int)w.cap - 1 : -1;
    { StageTimer t(h, st, NVR_STAGE_KNN);
    if (dense_a1)
        k_knn<4, true><<<grid_for(n, 256, sm * 8), 256, 0, st>>>(h->fdev, h->cfg.smpl_thresh, w.counters, w.surv, recs, (int)w.cap, w.raws, nullptr, -1, units);
    else if (h->cfg.tune & NVR_TUNE_KNN_OCC5)   // <= 48 registers: 5 CTAs (40 warps) per SM instead of 4
        k_knn<5><<<grid_for(n, 256, sm * 10), 256, 0, st>>>(h->fdev, h->cfg.smpl_thresh, w.counters, w.surv, recs, (int)w.cap, w.raws, dbg, far_slot, units);
    else
        k_knn<4><<<grid_for(n, 256, sm * 4), 256, 0, st>>>(h->fdev, h->cfg.smpl_thresh, w.counters, w.surv, recs, (int)w.cap, w.raws, dbg, far_slot, units); }   // work-counter loop: one wave of resident CTAs
    { StageTimer t(h, st, NVR_STAGE_WARP);
    if (!(h->cfg.tune & (NVR_TUNE_WARP_FFMA | NVR_TUNE_WARP_OCC4))) {
        // default: deformer MLP on tcgen05, one 128-pair tile per CTA iteration (nvr_warp_tc.cuh)
        const dim3 wg(grid_for(n, WT_THREADS, sm * 3 / 2), NVR_NUM_PARTS);
        k_warp_tc<<<wg, WT_THREADS, WT_SMEM_BYTES, st>>>(h->fdev, h->def_grid, h->def_mlp, dirs, dir_div, w.counters, w.surv, recs, w.pairs, (int)w.cap,
                                                         dbg, out_x0, out_resd, rank_of_slot);
    } else {
    const dim3 wg(grid_for(n, WARP_THREADS, sm * 3), NVR_NUM_PARTS);
    if (!(h->cfg.tune & NVR_TUNE_WARP_OCC4))    // <= 64 registers: 8 CTAs (32 warps) per SM instead of 4
        k_warp<8><<<wg, WARP_THREADS, 0, st>>>(h->fdev, h->def_grid, h->def_mlp, dirs, dir_div, w.counters, w.surv, recs, w.pairs, (int)w.cap, dbg, out_x0, out_resd, rank_of_slot);
    else
        k_warp<1><<<wg, WARP_THREADS, 0, st>>>(h->fdev, h->def_grid, h->def_mlp, dirs, dir_div, w.counters, w.surv, recs, w.pairs, (int)w.cap, dbg, out_x0, out_resd, rank_of_slot); } }
    const bool tc = h->cfg.mlp_mode >= 1;
    if (tc && !mlp_prep_done) {   // weights may have changed since the last call (training): repack every pass, 5 small CTAs
        StageTimer t(h, st, NVR_STAGE_MLP);
        launch_mlp_prep(h, st);
        h->launches++;
    }
    // Stage timing (nvr_profile) keeps one gather + one MLP launch per part, so the per-part event times exist; otherwise the five
    // gathers go out as one grid (blockIdx.y = part: the parts' launch tails overlap) and the five MLPs as one balanced launch
    const bool merged = !h->profiling && !(h->cfg.tune & NVR_TUNE_SERIAL);
    const bool presum = h->presum_valid && !full_tables;
    if (merged && !presum && !(h->cfg.tune & NVR_TUNE_LEVEL_MAJOR)) {
        EmbedBatch eb;
        for (int p = 0; p < NVR_NUM_PARTS; ++p) {
            eb.x[p] = (const float*)(w.pairs + (long long)p * w.cap);
            eb.count[p] = w.counters + NVR_CTR_PAIR + p;
            eb.out[p] = w.emb + (long long)p * w.cap * NVR_EMB_STRIDE;
            eb.work[p] = w.counters + NVR_CTR_EMBED_WORK + p;
        }
        k_embed_parts<<<dim3(grid_for(n, 128, sm * 2), NVR_NUM_PARTS), 256, 0, st>>>(h->d_part_grid, eb, 8, NVR_EMB_STRIDE);
        h->launches -= NVR_NUM_PARTS - 1;
    }
    for (int p = 0; p < NVR_NUM_PARTS; ++p) {
        const PairRec* pl = w.pairs + (long long)p * w.cap;
        float* el = w.emb + (long long)p * w.cap * NVR_EMB_STRIDE;
        if (!(merged && !presum && !(h->cfg.tune & NVR_TUNE_LEVEL_MAJOR))) {
        StageTimer t(h, st, NVR_STAGE_EMBED, p);
        if (presum) {
            const float* sd = h->d_presum + h->presum_off[p];
            k_embed_presum<<<grid_for(n, 256, sm * 4), 256, 0, st>>>(h->part_grid[p], sd, sd + dense_rows(h->params.part[p].grid), (const float*)pl, 8,
                                                                    w.counters + NVR_CTR_PAIR + p, 0, el, NVR_EMB_STRIDE);
        } else {
            int lb[NVR_MAX_LEVELS], le[NVR_MAX_LEVELS];
            const int n_slices = embed_plan(h, p, lb, le);
            for (int i = 0; i < n_slices; ++i)
                k_embed<<<grid_for(n, 128, sm * 2), 256, 0, st>>>(h->part_grid[p], (const float*)pl, 8, w.counters + NVR_CTR_PAIR + p, 0,
                                                                 el, NVR_EMB_STRIDE, lb[i], le[i]);
            h->launches += n_slices - 1;
        } }
        if (merged && h->cfg.mlp_mode == 3) continue;
        StageTimer t(h, st, NVR_STAGE_MLP, p);
        if (tc)
            launch_mlp_tc(h, grid_for(n, h->cfg.mlp_mode == 3 ? 128 * F16_SLOTS : 256, sm), h->d_mlp_blocks + (size_t)p * mlp_block_floats(h), h->part_mlp[p].n_rgb, p,
                          w.counters + NVR_CTR_PAIR + p, pl, el, w.raws, NVR_NUM_PARTS, st);
        else
            k_mlp<<<grid_for(n, MLP_TILE, sm), 256, MLP_SMEM_BYTES, st>>>(h->part_mlp[p], p, h->fdev.latent_index,
                                                                          w.counters + NVR_CTR_PAIR + p, pl, el, w.raws, NVR_NUM_PARTS);
    }
    if (merged && h->cfg.mlp_mode == 3) {
        MlpBatch mb;
        memset(&mb, 0, sizeof(mb));
        for (int p = 0; p < NVR_NUM_PARTS; ++p) {
            mb.blk[p] = h->d_mlp_blocks + (size_t)p * F16_BLOCK_FLOATS; mb.count[p] = w.counters + NVR_CTR_PAIR + p;
            mb.pl[p] = w.pairs + (long long)p * w.cap; mb.el[p] = w.emb + (long long)p * w.cap * NVR_EMB_STRIDE;
            mb.n_rgb[p] = h->part_mlp[p].n_rgb; mb.out_part[p] = p;
        }
        mb.n_parts = NVR_NUM_PARTS;
        k_mlp_f16<<<grid_for(n, 128 * F16_SLOTS / 2, sm), F16_THREADS, F16_SMEM_BYTES, st>>>(mb, w.raws, NVR_NUM_PARTS);
        h->launches -= NVR_NUM_PARTS - 1;
    }
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 3 + 2 * NVR_NUM_PARTS;
    h->last_points = n;
    h->last_passes = 1;
    h->last_lanes = 1;
    h->last_lanes_used = 1;
    return 0;
}

static int snapshot_counters(NvrEngine* h, const Workspace& w, cudaStream_t st) {
    if (h->profiling && h->n_pass_snap < PROF_MAX_PASSES) {
        NVR_CHECK(h, cudaMemcpyAsync(h->h_pass_counters + (size_t)h->n_pass_snap * NVR_CTR_WORDS, w.counters,
                                     NVR_CTR_WORDS * sizeof(int), cudaMemcpyDeviceToHost, st));
        h->n_pass_snap++;
    }
    if (h->snapshot_accumulate) {        // two-lane render: the call's passes add up (both lanes, atomics)
        k_add_counters<<<1, 32, 0, st>>>(w.counters, h->d_counters_snapshot);
        NVR_CHECK(h, cudaGetLastError());
    } else
    NVR_CHECK(h, cudaMemcpyAsync(h->d_counters_snapshot, w.counters, NVR_CTR_WORDS * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int ready(NvrEngine* h, const char* who) {
    if (!h) return 1;
    if (!h->have_params || !h->have_frame) {
        h->err = std::string(who) + ": bind_params and bind_frame first";
        return 1;
    }
    cudaError_t e = cudaSetDevice(h->cfg.device);
    if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return 2; }
    return 0;
}

extern "C" int nvr_query_points(NvrHandle h, const float* wpts, const float* viewdir, int64_t n, float* raw, float* occ,
                                void* workspace, size_t ws_bytes, void* stream_) {
    if (int rc = ready(h, "nvr_query_points")) return rc;
    if (n < 0 || (n > 0 && (!wpts || !viewdir || !raw))) return fail(h, "nvr_query_points: null argument");
    Workspace w;
    if (!carve(workspace, ws_bytes, w)) return fail(h, "nvr_query_points: workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream_;
    for (long long s = 0; s < n; s += w.pts) {
        const long long m = std::min<long long>(w.pts, n - s);
        if (int rc = run_pass(h, w, wpts + s * 3, nullptr, nullptr, nullptr, m, 0, viewdir + s * 3, 1, st)) return rc;
        { StageTimer t(h, st, NVR_STAGE_RESOLVE);
        k_resolve_points<<<grid_for(m, 256, h->sm_count * 16), 256, 0, st>>>(w.surv_of_sample, w.raws, far_raws(h, w, far_collapse(h, nullptr, nullptr)), m, (float4*)raw + s,
                                                                           occ ? occ + s : nullptr); }
        NVR_CHECK(h, cudaGetLastError());
        h->launches++;
        if (int rc = snapshot_counters(h, w, st)) return rc;
    }
    return 0;
}

// nvr_render_rays_host: the pinned host buffers behind the device arrays.  A two-lane render copies each pass's rays in and its
// pixels out on the pass's own stream, so the second lane's H2D and the first lane's D2H run under the other lane's kernels.
struct HostIO { const float* ray_o; const float* ray_d; const float* near_; const float* far_; float* rgb; float* acc; };

static int render_rays_impl(NvrHandle h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                            int64_t n_rays, int32_t n_samples, float* rgb_map, float* acc_map, float* raw,
                            void* workspace, size_t ws_bytes, void* stream_, const FrameOut& fo, const HostIO* hio = nullptr) {
    if (int rc = ready(h, "nvr_render_rays")) return rc;
    if (n_rays < 0 || n_samples < 1) return fail(h, "nvr_render_rays: bad n_rays / n_samples");
    if (n_rays > 0 && (!ray_o || !ray_d || !near_ || !far_ || ((!rgb_map || !acc_map) && fo.world == 0) || (!rgb_map != !acc_map)))
        return fail(h, "nvr_render_rays: null argument");
    Workspace w;
    if (!carve(workspace, ws_bytes, w)) return fail(h, "nvr_render_rays: workspace too small or not 256-byte aligned");
    long long rays_per_pass = w.pts / n_samples;
    if (rays_per_pass < 1) return fail(h, "nvr_render_rays: workspace smaller than one ray");
    cudaStream_t st = (cudaStream_t)stream_;
    // Two lanes: the call's rays go through the pipeline as (at least) two passes on two streams of the engine, each on its own
    // half of the workspace, so one pass's kernels fill the SMs the other pass's launch ramps and tails leave idle (a pass is
    // ~9 dependent launches; at an 8-GPU shard of a 512 x 512 frame ramps + tails were 0.2 of 0.86 ms).  Per-ray results do
    // not depend on which rays share a pass, so the output is bit-identical to the one-lane render.  Not in the serialised
    // profiling mode, not for small calls.
    Workspace lane_w[2];
    const size_t half_bytes = (ws_bytes / 2) & ~(size_t)255;
    const bool two = !h->profiling && !(h->cfg.tune & (NVR_TUNE_SERIAL | NVR_TUNE_ONE_LANE | NVR_TUNE_LEVEL_MAJOR)) &&
                     n_rays * (long long)n_samples >= h->two_lane_min &&
                     carve(workspace, half_bytes, lane_w[0]) && carve((char*)workspace + half_bytes, half_bytes, lane_w[1]) &&
                     lane_w[1].pts / n_samples >= 1;
    if (two) {
        rays_per_pass = std::min<long long>(lane_w[1].pts / n_samples, (n_rays + 1) / 2);
        NVR_CHECK(h, cudaMemsetAsync(h->d_counters_snapshot, 0, NVR_CTR_WORDS * sizeof(int), st));
        if (h->cfg.mlp_mode >= 1) { launch_mlp_prep(h, st); h->launches++; }        // once per call, before the lanes fork
        NVR_CHECK(h, cudaEventRecord(h->ev_fork, st));
        for (int l = 0; l < 2; ++l) NVR_CHECK(h, cudaStreamWaitEvent(h->part_stream[l], h->ev_fork, 0));
        h->snapshot_accumulate = true;
        int rc = 0, pass = 0;
        for (long long r = 0; r < n_rays && !rc; r += rays_per_pass, ++pass) {
            const Workspace& lw = lane_w[pass & 1];
            cudaStream_t ls = h->part_stream[pass & 1];
            const long long nr = std::min<long long>(rays_per_pass, n_rays - r);
            if (hio) {
                cudaMemcpyAsync((float*)ray_o + r * 3, hio->ray_o + r * 3, nr * 3 * sizeof(float), cudaMemcpyHostToDevice, ls);
                cudaMemcpyAsync((float*)ray_d + r * 3, hio->ray_d + r * 3, nr * 3 * sizeof(float), cudaMemcpyHostToDevice, ls);
                cudaMemcpyAsync((float*)near_ + r, hio->near_ + r, nr * sizeof(float), cudaMemcpyHostToDevice, ls);
                if (cudaMemcpyAsync((float*)far_ + r, hio->far_ + r, nr * sizeof(float), cudaMemcpyHostToDevice, ls) != cudaSuccess) {
                    h->err = std::string("nvr_render_rays_host: H2D copy: ") + cudaGetErrorString(cudaGetLastError()); rc = 2; break;
                }
            }
            rc = run_pass(h, lw, ray_o + r * 3, ray_d + r * 3, near_ + r, far_ + r, nr * n_samples, n_samples, ray_d + r * 3, n_samples, ls,
                          nullptr, nullptr, nullptr, false, nullptr, true);
            if (rc) break;
            k_resolve_rays<<<grid_for(nr, 8, h->sm_count * 8), 256, 0, ls>>>(lw.surv_of_sample, lw.raws, far_raws(h, lw, far_collapse(h, nullptr, nullptr)), nr, n_samples,
                                                                           rgb_map ? rgb_map + r * 3 : nullptr, acc_map ? acc_map + r : nullptr,
                                                                           raw ? (float4*)raw + r * n_samples : nullptr, fo, r);
            if (cudaGetLastError() != cudaSuccess) { h->err = "k_resolve_rays launch failed"; rc = 2; break; }
            h->launches++;
            if (hio) {
                cudaMemcpyAsync(hio->rgb + r * 3, rgb_map + r * 3, nr * 3 * sizeof(float), cudaMemcpyDeviceToHost, ls);
                if (cudaMemcpyAsync(hio->acc + r, acc_map + r, nr * sizeof(float), cudaMemcpyDeviceToHost, ls) != cudaSuccess) {
                    h->err = std::string("nvr_render_rays_host: D2H copy: ") + cudaGetErrorString(cudaGetLastError()); rc = 2; break;
                }
            }
            rc = snapshot_counters(h, lw, ls);
        }
        h->snapshot_accumulate = false;
        for (int l = 0; l < 2; ++l) {                 // join even after an error: the caller's stream must not run ahead of the lanes
            cudaEventRecord(h->ev_join[l], h->part_stream[l]);
            cudaStreamWaitEvent(st, h->ev_join[l], 0);
        }
        h->last_points = n_rays * (long long)n_samples;
        h->last_passes = pass;
        h->last_lanes = 2;
        h->last_lanes_used = pass >= 2 ? 2 : 1;
        return rc;
    }
    if (hio) {
        NVR_CHECK(h, cudaMemcpyAsync((float*)ray_o, hio->ray_o, n_rays * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
        NVR_CHECK(h, cudaMemcpyAsync((float*)ray_d, hio->ray_d, n_rays * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
        NVR_CHECK(h, cudaMemcpyAsync((float*)near_, hio->near_, n_rays * sizeof(float), cudaMemcpyHostToDevice, st));
        NVR_CHECK(h, cudaMemcpyAsync((float*)far_, hio->far_, n_rays * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    for (long long r = 0; r < n_rays; r += rays_per_pass) {
        const long long nr = std::min<long long>(rays_per_pass, n_rays - r);
        const long long m = nr * n_samples;
        if (int rc = run_pass(h, w, ray_o + r * 3, ray_d + r * 3, near_ + r, far_ + r, m, n_samples, ray_d + r * 3, n_samples, st))
            return rc;
        { StageTimer t(h, st, NVR_STAGE_RESOLVE);
        k_resolve_rays<<<grid_for(nr, 8, h->sm_count * 8), 256, 0, st>>>(w.surv_of_sample, w.raws, far_raws(h, w, far_collapse(h, nullptr, nullptr)), nr, n_samples,
                                                                       rgb_map ? rgb_map + r * 3 : nullptr, acc_map ? acc_map + r : nullptr,
                                                                       raw ? (float4*)raw + r * n_samples : nullptr, fo, r); }
        NVR_CHECK(h, cudaGetLastError());
        h->launches++;
        if (int rc = snapshot_counters(h, w, st)) return rc;
    }
    if (hio) {
        NVR_CHECK(h, cudaMemcpyAsync(hio->rgb, rgb_map, n_rays * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
        NVR_CHECK(h, cudaMemcpyAsync(hio->acc, acc_map, n_rays * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

extern "C" int nvr_render_rays(NvrHandle h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                               int64_t n_rays, int32_t n_samples, float* rgb_map, float* acc_map, float* raw,
                               void* workspace, size_t ws_bytes, void* stream_) {
    FrameOut none;
    memset(&none, 0, sizeof(none));
    return render_rays_impl(h, ray_o, ray_d, near_, far_, n_rays, n_samples, rgb_map, acc_map, raw, workspace, ws_bytes, stream_, none);
}

// ---- multi-GPU frame assembly over peer memory (nvr_frame.cuh) -----------------------------------------------------
static void frame_release(NvrEngine* h) {
    NvrEngine::PeerFrame& pf = h->pf;
    for (int r = 0; r < NVR_MAX_RANKS; ++r)
        if (pf.peer[r] && pf.peer[r] != pf.local) cudaIpcCloseMemHandle(pf.peer[r]);
    if (pf.local) cudaFree(pf.local);
    pf = NvrEngine::PeerFrame();
}

extern "C" int nvr_frame_create(NvrHandle h, int64_t n_rays_total, int32_t rank, int32_t world, int32_t tile, NvrIpcHandle* handle_out) {
    if (!h || !handle_out) return fail(h, "nvr_frame_create: null argument");
    if (world < 1 || world > NVR_MAX_RANKS || rank < 0 || rank >= world || tile < 1 || n_rays_total < 0)
        return fail(h, "nvr_frame_create: need 1 <= world <= 8, 0 <= rank < world, tile >= 1");
    if (h->pf.local) return fail(h, "nvr_frame_create: a frame buffer exists (nvr_frame_disconnect on every rank, then nvr_frame_destroy)");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    NvrEngine::PeerFrame& pf = h->pf;
    const size_t bytes = NVR_FRAME_HEADER_BYTES + 2 * (size_t)std::max<int64_t>(n_rays_total, 1) * sizeof(float4);
    NVR_CHECK(h, cudaMalloc(&pf.local, bytes));
    NVR_CHECK(h, cudaMemset(pf.local, 0, bytes));
    pf.n_total = n_rays_total; pf.rank = rank; pf.world = world; pf.tile = tile; pf.epoch = 0; pf.connected = false;
    memset(handle_out, 0, sizeof(*handle_out));
    if (world > 1) {
        static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(NvrIpcHandle), "NvrIpcHandle holds a cudaIpcMemHandle_t");
        cudaIpcMemHandle_t ih;
        NVR_CHECK(h, cudaIpcGetMemHandle(&ih, pf.local));
        memcpy(handle_out->bytes, &ih, sizeof(ih));
    }
    return 0;
}

extern "C" int nvr_frame_connect(NvrHandle h, const NvrIpcHandle* handles) {
    if (!h) return 1;
    NvrEngine::PeerFrame& pf = h->pf;
    if (!pf.local) return fail(h, "nvr_frame_connect: nvr_frame_create first");
    if (pf.world > 1 && !handles) return fail(h, "nvr_frame_connect: null argument");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    for (int r = 0; r < pf.world; ++r) {
        if (r == pf.rank) { pf.peer[r] = pf.local; continue; }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, handles[r].bytes, sizeof(ih));
        NVR_CHECK(h, cudaIpcOpenMemHandle(&pf.peer[r], ih, cudaIpcMemLazyEnablePeerAccess));
    }
    pf.connected = true;
    return 0;
}

extern "C" int nvr_frame_disconnect(NvrHandle h) {
    if (!h) return 0;
    NvrEngine::PeerFrame& pf = h->pf;
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    NVR_CHECK(h, cudaDeviceSynchronize());
    for (int r = 0; r < NVR_MAX_RANKS; ++r) {
        if (pf.peer[r] && pf.peer[r] != pf.local) NVR_CHECK(h, cudaIpcCloseMemHandle(pf.peer[r]));
        pf.peer[r] = nullptr;
    }
    pf.connected = false;
    return 0;
}

extern "C" int nvr_frame_destroy(NvrHandle h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    frame_release(h);
    return 0;
}

// the slot the NEXT frame goes to, as every rank sees it, + the barrier that completes it
static FrameOut frame_next(NvrEngine* h) {
    NvrEngine::PeerFrame& pf = h->pf;
    FrameOut fo;
    memset(&fo, 0, sizeof(fo));
    const size_t slot_off = NVR_FRAME_HEADER_BYTES + (size_t)((pf.epoch + 1) & 1) * (size_t)std::max<long long>(pf.n_total, 1) * sizeof(float4);
    for (int r = 0; r < pf.world; ++r) fo.slot[r] = (float4*)((char*)pf.peer[r] + slot_off);
    fo.world = pf.world; fo.rank = pf.rank; fo.tile = pf.tile; fo.n_total = pf.n_total;
    return fo;
}
static int frame_finish(NvrEngine* h, const FrameOut& fo, cudaStream_t st, const float** frame_out) {
    NvrEngine::PeerFrame& pf = h->pf;
    pf.epoch++;
    FrameFlags ff;
    for (int r = 0; r < NVR_MAX_RANKS; ++r) ff.peer[r] = r < pf.world ? (unsigned int*)pf.peer[r] : nullptr;
    k_frame_barrier<<<1, 32, 0, st>>>(ff, pf.world, pf.rank, pf.epoch);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    if (frame_out) *frame_out = (const float*)fo.slot[pf.rank];
    return 0;
}

extern "C" int nvr_render_rays_frame(NvrHandle h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                                     int64_t n_rays_local, int32_t n_samples, float* rgb_map, float* acc_map,
                                     void* workspace, size_t ws_bytes, void* stream_, const float** frame_out) {
    if (!h) return 1;
    if (!h->pf.connected) return fail(h, "nvr_render_rays_frame: nvr_frame_create + nvr_frame_connect first");
    const FrameOut fo = frame_next(h);
    if (int rc = render_rays_impl(h, ray_o, ray_d, near_, far_, n_rays_local, n_samples, rgb_map, acc_map, nullptr, workspace, ws_bytes, stream_, fo))
        return rc;
    return frame_finish(h, fo, (cudaStream_t)stream_, frame_out);
}

extern "C" int nvr_render_rays_frame_host(NvrHandle h, const float* ray_o_host, const float* ray_d_host, const float* near_host,
                                          const float* far_host, int64_t n_rays, int32_t n_samples, float* rgb_map_host,
                                          float* acc_map_host, void* dev_io, void* workspace, size_t ws_bytes, void* stream_,
                                          const float** frame_out) {
    if (!h) return 1;
    if (!h->pf.connected) return fail(h, "nvr_render_rays_frame_host: nvr_frame_create + nvr_frame_connect first");
    if (n_rays > 0 && (!dev_io || !ray_o_host || !ray_d_host || !near_host || !far_host || !rgb_map_host || !acc_map_host))
        return fail(h, "nvr_render_rays_frame_host: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    float* d = (float*)dev_io;                      // [o 3n | d 3n | near n | far n | rgb 3n | acc n] = 12n floats
    float *d_o = d, *d_d = d + 3 * n_rays, *d_n = d + 6 * n_rays, *d_f = d + 7 * n_rays, *d_rgb = d + 8 * n_rays,
          *d_acc = d + 11 * n_rays;
    const HostIO hio{ray_o_host, ray_d_host, near_host, far_host, rgb_map_host, acc_map_host};
    const FrameOut fo = frame_next(h);
    if (int rc = render_rays_impl(h, d_o, d_d, d_n, d_f, n_rays, n_samples, d_rgb, d_acc, nullptr, workspace, ws_bytes, stream_, fo, &hio))
        return rc;
    if (int rc = frame_finish(h, fo, st, frame_out)) return rc;
    NVR_CHECK(h, cudaStreamSynchronize(st));
    return 0;
}

extern "C" int nvr_allgather_frame(NvrHandle h, const float* rgb_map, const float* acc_map, int64_t n_rays_local, void* stream_,
                                   const float** frame_out) {
    if (!h) return 1;
    if (!h->pf.connected) return fail(h, "nvr_allgather_frame: nvr_frame_create + nvr_frame_connect first");
    if (n_rays_local < 0 || (n_rays_local > 0 && (!rgb_map || !acc_map))) return fail(h, "nvr_allgather_frame: null argument");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    const FrameOut fo = frame_next(h);
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_rays_local > 0) {
        k_frame_scatter<<<grid_for(n_rays_local, 256, h->sm_count * 4), 256, 0, st>>>(fo, rgb_map, acc_map, n_rays_local);
        h->launches++;
    }
    return frame_finish(h, fo, st, frame_out);
}

extern "C" int nvr_render_rays_host(NvrHandle h, const float* ray_o_host, const float* ray_d_host, const float* near_host,
                                    const float* far_host, int64_t n_rays, int32_t n_samples, float* rgb_map_host,
                                    float* acc_map_host, void* dev_io, void* workspace, size_t ws_bytes, void* stream_) {
    if (int rc = ready(h, "nvr_render_rays_host")) return rc;
    if (!dev_io || !ray_o_host || !ray_d_host || !near_host || !far_host || !rgb_map_host || !acc_map_host)
        return fail(h, "nvr_render_rays_host: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    float* d = (float*)dev_io;                      // [o 3n | d 3n | near n | far n | rgb 3n | acc n] = 12n floats
    float *d_o = d, *d_d = d + 3 * n_rays, *d_n = d + 6 * n_rays, *d_f = d + 7 * n_rays, *d_rgb = d + 8 * n_rays,
          *d_acc = d + 11 * n_rays;
    const HostIO hio{ray_o_host, ray_d_host, near_host, far_host, rgb_map_host, acc_map_host};
    FrameOut none;
    memset(&none, 0, sizeof(none));
    if (int rc = render_rays_impl(h, d_o, d_d, d_n, d_f, n_rays, n_samples, d_rgb, d_acc, nullptr, workspace, ws_bytes, stream_, none, &hio))
        return rc;
    NVR_CHECK(h, cudaStreamSynchronize(st));
    return 0;
}

extern "C" int nvr_deformer_residual(NvrHandle h, const float* tpts, int64_t n, float* resd, void* stream_) {
    if (int rc = ready(h, "nvr_deformer_residual")) return rc;
    if (n > 0 && (!tpts || !resd)) return fail(h, "nvr_deformer_residual: null argument");
    if (n == 0) return 0;
    k_deformer<<<grid_for(n, WARP_THREADS, h->sm_count * 4), WARP_THREADS, 0, (cudaStream_t)stream_>>>(h->fdev, h->def_grid, h->def_mlp, tpts, n, resd);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int nvr_embed_part(NvrHandle h, int32_t part, const float* xyz, int64_t n, float* out, void* stream_) {
    if (!h || !h->have_params) return fail(h, "nvr_embed_part: bind_params first");
    if (part < 0 || part >= NVR_NUM_PARTS) return fail(h, "nvr_embed_part: bad part");
    if (n > 0 && (!xyz || !out)) return fail(h, "nvr_embed_part: null argument");
    if (n == 0) return 0;
    if (n >= (1ll << 31)) return fail(h, "nvr_embed_part: n must be < 2^31");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    k_embed<<<grid_for(n, 128, h->sm_count * 2), 256, 0, (cudaStream_t)stream_>>>(h->part_grid[part], xyz, 3, nullptr, (int)n, out, 19, 0,
                                                                                  h->params.part[part].grid.n_levels);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int nvr_part_mlp(NvrHandle h, int32_t part, const float* emb, const float* dirs, int64_t n, float* raw,
                            void* workspace, size_t ws_bytes, void* stream_) {
    if (int rc = ready(h, "nvr_part_mlp")) return rc;
    if (part < 0 || part >= NVR_NUM_PARTS) return fail(h, "nvr_part_mlp: bad part");
    if (n > 0 && (!emb || !dirs || !raw)) return fail(h, "nvr_part_mlp: null argument");
    if (n == 0) return 0;
    Workspace w;
    if (!carve(workspace, ws_bytes, w) || w.pts < n) return fail(h, "nvr_part_mlp: workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    k_make_pairs<<<(int)((n + 255) / 256), 256, 0, st>>>(dirs, (int)n, w.pairs, w.counters);
    if (h->cfg.mlp_mode >= 1) {
        if (((uintptr_t)emb & 15) != 0) return fail(h, "nvr_part_mlp: emb must be 16-byte aligned (rows of 20 floats)");
        launch_mlp_prep(h, st);
        launch_mlp_tc(h, grid_for(n, h->cfg.mlp_mode == 3 ? 128 * F16_SLOTS : 256, h->sm_count), h->d_mlp_blocks + (size_t)part * mlp_block_floats(h), h->part_mlp[part].n_rgb, 0,
                      w.counters, w.pairs, emb, (float4*)raw, 1, st);
        h->launches++;
    } else {
        k_mlp<<<grid_for(n, MLP_TILE, h->sm_count), 256, MLP_SMEM_BYTES, st>>>(h->part_mlp[part], 0, h->fdev.latent_index, w.counters,
                                                                              w.pairs, emb, (float4*)raw, 1);
    }
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 2;
    return 0;
}

extern "C" int nvr_query_points_debug(NvrHandle h, const float* wpts, const float* viewdir, int64_t n, float* raw,
                                      int32_t* surv_of_sample, float* warp_dbg, void* workspace, size_t ws_bytes, void* stream_) {
    if (int rc = ready(h, "nvr_query_points_debug")) return rc;
    Workspace w;
    if (!carve(workspace, ws_bytes, w) || w.pts < n) return fail(h, "nvr_query_points_debug: workspace must hold all points in one pass");
    if (!wpts || !viewdir || !raw || !surv_of_sample || !warp_dbg) return fail(h, "nvr_query_points_debug: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    NVR_CHECK(h, cudaMemsetAsync(warp_dbg, 0, (size_t)n * NVR_NUM_PARTS * 8 * sizeof(float), st));
    if (int rc = run_pass(h, w, wpts, nullptr, nullptr, nullptr, n, 0, viewdir, 1, st, warp_dbg)) return rc;
    k_resolve_points<<<grid_for(n, 256, h->sm_count * 16), 256, 0, st>>>(w.surv_of_sample, w.raws, nullptr, n, (float4*)raw, nullptr);
    NVR_CHECK(h, cudaMemcpyAsync(surv_of_sample, w.surv_of_sample, n * sizeof(int), cudaMemcpyDeviceToDevice, st));
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return snapshot_counters(h, w, st);
}

// ---- inference tables ----------------------------------------------------------------------------
extern "C" int nvr_prepare_inference(NvrHandle h, int32_t enable, void* stream_) {
    if (!h || !h->have_params) return fail(h, "nvr_prepare_inference: bind_params first");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    if (!enable) { h->presum_valid = false; return 0; }
    size_t total = 0;
    for (int p = 0; p < NVR_NUM_PARTS; ++p) {
        h->presum_off[p] = total;
        total += (size_t)(dense_rows(h->params.part[p].grid) + hash_rows(h->params.part[p].grid));
    }
    h->presum_off[NVR_NUM_PARTS] = total;
    if (total > h->presum_cap) {
        NVR_CHECK(h, cudaFree(h->d_presum));
        h->d_presum = nullptr; h->presum_cap = 0;
        NVR_CHECK(h, cudaMalloc(&h->d_presum, total * sizeof(float)));
        h->presum_cap = total;
    }
    cudaStream_t st = (cudaStream_t)stream_;
    for (int p = 0; p < NVR_NUM_PARTS; ++p) {
        const NvrGrid& g = h->params.part[p].grid;
        const long long nd = dense_rows(g), nh = hash_rows(g);
        float* out = h->d_presum + h->presum_off[p];
        k_presum_rows<<<grid_for(nd, 256, h->sm_count * 16), 256, 0, st>>>(g.dense, nd, out);
        k_presum_rows<<<grid_for(nh, 256, h->sm_count * 16), 256, 0, st>>>(g.hash, nh, out + nd);
    }
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 2 * NVR_NUM_PARTS;
    h->presum_valid = true;
    return 0;
}

// ---- training ----------------------------------------------------------------------------------
__global__ void k_export_slots(const int* __restrict__ counters, const float4* __restrict__ raws, const int* __restrict__ rank_of_slot,
                               float* __restrict__ tocc) {
    const int n = counters[NVR_CTR_SURV];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const long long row = rank_of_slot[s];
#pragma unroll
        for (int p = 0; p < NVR_PARTS; ++p) tocc[row * NVR_PARTS + p] = raws[(long long)s * NVR_PARTS + p].w;
    }
}

extern "C" int nvr_train_forward(NvrHandle h, const float* wpts, const float* viewdir, int64_t n, float* raw, float* occ,
                                 float* x0, float* resd, float* tocc, int32_t* rank_of_slot, void* workspace, size_t ws_bytes,
                                 void* stream_) {
    if (int rc = ready(h, "nvr_train_forward")) return rc;
    if (n < 0 || (n > 0 && (!wpts || !viewdir || !raw || !x0 || !resd || !tocc || !rank_of_slot)))
        return fail(h, "nvr_train_forward: null argument");
    Workspace w;
    if (!carve(workspace, ws_bytes, w) || w.pts < n) return fail(h, "nvr_train_forward: workspace must hold all points in one pass");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    NVR_CHECK(h, cudaMemsetAsync(x0, 0, (size_t)n * NVR_NUM_PARTS * 3 * sizeof(float), st));
    NVR_CHECK(h, cudaMemsetAsync(resd, 0, (size_t)n * NVR_NUM_PARTS * 3 * sizeof(float), st));
    NVR_CHECK(h, cudaMemsetAsync(tocc, 0, (size_t)n * NVR_NUM_PARTS * sizeof(float), st));
    if (int rc = run_pass(h, w, wpts, nullptr, nullptr, nullptr, n, 0, viewdir, 1, st, nullptr, x0, resd, true, rank_of_slot)) return rc;
    k_resolve_points<<<grid_for(n, 256, h->sm_count * 16), 256, 0, st>>>(w.surv_of_sample, w.raws, nullptr, n, (float4*)raw, occ);
    k_export_slots<<<grid_for(n, 256, h->sm_count * 8), 256, 0, st>>>(w.counters, w.raws, rank_of_slot, tocc);
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 2;
    return snapshot_counters(h, w, st);
}

// scratch of the backward: [gcount 64 B][glist 5*cap GradRec][d_emb 5*cap*20][d_x 5*cap*3][wl_x0 5*cap*3][wl_dr 5*cap*3]
static size_t train_scratch_bytes(long long cap) {
    return 256 + (size_t)cap * NVR_NUM_PARTS * (sizeof(GradRec) + (NVR_EMB_STRIDE + 9) * sizeof(float));
}
// sized for the stride carve() gives a workspace of nvr_workspace_bytes(n) (the same arithmetic, so the two cannot drift apart)
extern "C" size_t nvr_train_scratch_bytes(NvrHandle h, int64_t n) {
    const size_t bytes = nvr_workspace_bytes(h, n);
    long long cap = (long long)((bytes - WS_HEADER) / WS_PER_POINT) - 64;
    cap = std::min<long long>(cap & ~63ll, 1ll << 30);
    return train_scratch_bytes(cap);
}

static PartMlpGrad mlp_grad(const NvrPart& g, int n_rgb) {
    PartMlpGrad o;
    for (int i = 0; i < 2; ++i) o.occ[i] = LinearGrad{(float*)g.occ[i].weight, (float*)g.occ[i].bias};
    for (int i = 0; i < 3; ++i) o.rgb[i] = i < n_rgb ? LinearGrad{(float*)g.rgb[i].weight, (float*)g.rgb[i].bias} : LinearGrad{nullptr, nullptr};
    o.latent = (float*)g.rgb_latent;
    return o;
}
static DeformerGrad deformer_grad(const NvrParams* g) {
    return DeformerGrad{(float*)g->deformer_mlp[0].weight, (float*)g->deformer_mlp[0].bias, (float*)g->deformer_mlp[1].weight,
                        (float*)g->deformer_mlp[1].bias, (float*)g->deformer_mlp[2].weight, (float*)g->deformer_mlp[2].bias};
}

extern "C" int nvr_train_backward(NvrHandle h, const float* d_raw, const float* d_resd, const float* d_tocc, const float* x0,
                                  const int32_t* rank_of_slot, int64_t n, const NvrParams* grads, void* workspace, size_t ws_bytes,
                                  void* scratch, size_t scratch_bytes, void* stream_) {
    if (int rc = ready(h, "nvr_train_backward")) return rc;
    if (n < 0 || !grads || (n > 0 && (!d_raw || !x0 || !rank_of_slot))) return fail(h, "nvr_train_backward: null argument");
    Workspace w;
    if (!carve(workspace, ws_bytes, w) || w.pts < n) return fail(h, "nvr_train_backward: not the forward's workspace");
    if (!scratch || ((uintptr_t)scratch & 255) || scratch_bytes < train_scratch_bytes(w.cap))
        return fail(h, "nvr_train_backward: scratch too small (nvr_train_scratch_bytes) or not 256-byte aligned");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    const long long cap = w.cap;
    char* p = (char*)scratch;
    int* gcount = (int*)p; p += 256;
    GradRec* glist = (GradRec*)p; p += (size_t)cap * NVR_NUM_PARTS * sizeof(GradRec);
    float* d_emb = (float*)p; p += (size_t)cap * NVR_NUM_PARTS * NVR_EMB_STRIDE * sizeof(float);
    float* d_x = (float*)p; p += (size_t)cap * NVR_NUM_PARTS * 3 * sizeof(float);
    float* wl_x0 = (float*)p; p += (size_t)cap * NVR_NUM_PARTS * 3 * sizeof(float);
    float* wl_dr = (float*)p;
    const int sm = h->sm_count;
    NVR_CHECK(h, cudaMemsetAsync(gcount, 0, 64, st));
    k_bwd_select<<<dim3(grid_for(n, 256, sm * 4), NVR_NUM_PARTS), 256, 0, st>>>(w.counters, w.pairs, (int)cap, w.surv, w.raws,
                                                                              (const float4*)d_raw, d_tocc, rank_of_slot, glist, gcount);
    k_bwd_resd_list<<<dim3(grid_for(n, 256, sm * 4), NVR_NUM_PARTS), 256, 0, st>>>(w.counters, w.pairs, (int)cap, x0, d_resd, rank_of_slot, wl_x0, wl_dr);
    // the five parts' chains are independent (they only meet in atomic adds on the deformer's gradients) and, at training sizes,
    // each kernel is a handful of CTAs: run them side by side on the engine's part streams instead of back to back
    NVR_CHECK(h, cudaEventRecord(h->ev_fork, st));
    cudaStream_t caller = st;
    for (int pt = 0; pt < NVR_NUM_PARTS; ++pt) {
        st = h->part_stream[pt];
        NVR_CHECK(h, cudaStreamWaitEvent(st, h->ev_fork, 0));
        const NvrPart& gp = grads->part[pt];
        const GradRec* gl = glist + (size_t)pt * cap;
        const PairRec* pl = w.pairs + (size_t)pt * cap;
        const float* el = w.emb + (size_t)pt * cap * NVR_EMB_STRIDE;
        float* de = d_emb + (size_t)pt * cap * NVR_EMB_STRIDE;
        float* dx = d_x + (size_t)pt * cap * 3;
        k_mlp_bwd<<<grid_for(n, BT, sm), BTHREADS, MB_SMEM_BYTES, st>>>(h->part_mlp[pt], mlp_grad(gp, h->part_mlp[pt].n_rgb),
                                                                       h->fdev.latent_index, gcount + pt, gl, pl, el, de);
        k_embed_bwd<<<grid_for(n, 16, sm * 8), 256, 0, st>>>(h->part_grid[pt], GridGrad{(float*)gp.grid.dense, (float*)gp.grid.hash},
                                                           gcount + pt, gl, pl, de, dx);
        k_bwd_add_dx<<<grid_for(n, 256, sm * 2), 256, 0, st>>>(gcount + pt, gl, dx, wl_dr + (size_t)pt * cap * 3);
        k_deformer_bwd<<<grid_for(n, BT, sm), BTHREADS, DB_SMEM_BYTES, st>>>(
            h->fdev, h->def_grid, h->def_mlp, deformer_grad(grads), GridGrad{(float*)grads->deformer_grid.dense, (float*)grads->deformer_grid.hash},
            wl_x0 + (size_t)pt * cap * 3, wl_dr + (size_t)pt * cap * 3, w.counters + NVR_CTR_PAIR + pt, 0);
        NVR_CHECK(h, cudaEventRecord(h->ev_join[pt], st));
        NVR_CHECK(h, cudaStreamWaitEvent(caller, h->ev_join[pt], 0));
    }
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 2 + 4 * NVR_NUM_PARTS;
    return 0;
}

// ---- training-time sampling and the distortion regulariser (inb_renderer.py:15-31, 96-103) ------------------------------
extern "C" int nvr_train_sample(NvrHandle h, const float* ray_o, const float* ray_d, const float* near_, const float* far_,
                                const float* u, int64_t n_rays, int32_t n_samples, float* z_vals, float* wpts, float* viewdir,
                                void* stream_) {
    if (!h) return 1;
    if (n_rays < 0 || n_samples < 1 || (n_rays > 0 && (!ray_o || !ray_d || !near_ || !far_ || !z_vals || !wpts || !viewdir)))
        return fail(h, "nvr_train_sample: bad argument");
    if (n_rays == 0) return 0;
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    k_train_sample<<<grid_for(n_rays * n_samples, 256, h->sm_count * 8), 256, 0, (cudaStream_t)stream_>>>(ray_o, ray_d, near_, far_, u, n_rays, n_samples,
                                                                                                    z_vals, wpts, viewdir);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

static int distortion_launch(NvrHandle h, const float* weights, const float* z_vals, const float* d_loss, int64_t n_rays, int32_t S,
                             float* loss, float* d_weights, void* stream_) {
    if (n_rays == 0) return 0;
    if (S > 4096) return fail(h, "nvr_distortion: more than 4096 samples per ray");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    k_distortion<<<grid_for(n_rays, 8, h->sm_count * 8), 256, 8 * 2 * S * sizeof(float), (cudaStream_t)stream_>>>(weights, z_vals, d_loss, n_rays, S, loss, d_weights);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}
extern "C" int nvr_distortion_forward(NvrHandle h, const float* weights, const float* z_vals, int64_t n_rays, int32_t n_samples,
                                      float* loss, void* stream_) {
    if (!h) return 1;
    if (n_rays < 0 || n_samples < 1 || (n_rays > 0 && (!weights || !z_vals || !loss))) return fail(h, "nvr_distortion_forward: bad argument");
    return distortion_launch(h, weights, z_vals, nullptr, n_rays, n_samples, loss, nullptr, stream_);
}
extern "C" int nvr_distortion_backward(NvrHandle h, const float* weights, const float* z_vals, const float* d_loss, int64_t n_rays,
                                       int32_t n_samples, float* d_weights, void* stream_) {
    if (!h) return 1;
    if (n_rays < 0 || n_samples < 1 || (n_rays > 0 && (!weights || !z_vals || !d_loss || !d_weights)))
        return fail(h, "nvr_distortion_backward: bad argument");
    return distortion_launch(h, weights, z_vals, d_loss, n_rays, n_samples, nullptr, d_weights, stream_);
}

extern "C" int nvr_deformer_backward(NvrHandle h, const float* tpts, const float* d_resd, int64_t n, const NvrParams* grads, void* stream_) {
    if (int rc = ready(h, "nvr_deformer_backward")) return rc;
    if (n < 0 || !grads || (n > 0 && (!tpts || !d_resd))) return fail(h, "nvr_deformer_backward: null argument");
    if (n == 0) return 0;
    if (n >= (1ll << 31)) return fail(h, "nvr_deformer_backward: n must be < 2^31");
    k_deformer_bwd<<<grid_for(n, BT, h->sm_count), BTHREADS, DB_SMEM_BYTES, (cudaStream_t)stream_>>>(
        h->fdev, h->def_grid, h->def_mlp, deformer_grad(grads), GridGrad{(float*)grads->deformer_grid.dense, (float*)grads->deformer_grid.hash},
        tpts, d_resd, nullptr, (int)n);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int nvr_composite_forward(NvrHandle h, const float* raw, int64_t n_rays, int32_t n_samples, float* weights, float* rgb_map,
                                     float* acc_map, void* stream_) {
    if (!h) return 1;
    if (n_rays < 0 || n_samples < 1 || (n_rays > 0 && (!raw || !weights || !rgb_map || !acc_map))) return fail(h, "nvr_composite_forward: bad argument");
    if (n_rays == 0) return 0;
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    k_composite_fwd<<<grid_for(n_rays, 128, h->sm_count * 8), 128, 0, (cudaStream_t)stream_>>>((const float4*)raw, n_rays, n_samples, weights, rgb_map, acc_map);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int nvr_composite_backward(NvrHandle h, const float* raw, int64_t n_rays, int32_t n_samples, const float* d_weights,
                                      const float* d_rgb_map, const float* d_acc_map, float* d_raw, void* stream_) {
    if (!h) return 1;
    if (n_rays < 0 || n_samples < 1 || (n_rays > 0 && (!raw || !d_raw))) return fail(h, "nvr_composite_backward: bad argument");
    if (n_rays == 0) return 0;
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    k_composite_bwd<<<grid_for(n_rays, 128, h->sm_count * 8), 128, 0, (cudaStream_t)stream_>>>((const float4*)raw, n_rays, n_samples, d_weights,
                                                                                             d_rgb_map, d_acc_map, (float4*)d_raw);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// ---- optimizer step / camera rays / image metrics --------------------------------
extern "C" int nvr_adam_step(NvrHandle h, const NvrAdamTensor* tensors, int32_t n_tensors, double beta1, double beta2, double eps,
                             int32_t zero_grad, void* stream_) {
    if (!h) return 1;
    if (n_tensors < 0 || (n_tensors > 0 && !tensors)) return fail(h, "nvr_adam_step: null argument");
    if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0)) return fail(h, "nvr_adam_step: bad betas / eps");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream_;
    AdamBatch b;
    memset(&b, 0, sizeof(b));
    b.zero_grad = zero_grad != 0;
    auto flush = [&]() -> cudaError_t {
        if (b.n_tensors == 0) return cudaSuccess;
        k_adam<<<b.chunk_begin[b.n_tensors], 256, 0, st>>>(b);
        h->launches++;
        b.n_tensors = 0;
        return cudaGetLastError();
    };
    for (int i = 0; i < n_tensors; ++i) {
        const NvrAdamTensor& t = tensors[i];
        if (t.numel == 0) continue;
        if (t.numel < 0 || !t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq || t.step < 1)
            return fail(h, "nvr_adam_step: tensor with null pointer, negative numel or step < 1");
        const long long chunks = (t.numel + ADAM_CHUNK - 1) / ADAM_CHUNK;
        if (chunks >= (1ll << 30)) return fail(h, "nvr_adam_step: tensor too large");
        if (b.n_tensors == ADAM_MAX_TENSORS || (long long)b.chunk_begin[b.n_tensors] + chunks >= (1ll << 31)) NVR_CHECK(h, flush());
        AdamTensorDev& d = b.t[b.n_tensors];
        d.p = t.param; d.g = t.grad; d.m = t.exp_avg; d.v = t.exp_avg_sq; d.n = t.numel;
        d.vec = (((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) & 15) == 0;
        const double bc1 = 1.0 - pow(beta1, (double)t.step), bc2 = 1.0 - pow(beta2, (double)t.step);
        d.s.beta1 = (float)beta1; d.s.beta2 = (float)beta2;
        d.s.one_minus_beta1 = (float)(1.0 - beta1); d.s.one_minus_beta2 = (float)(1.0 - beta2);
        d.s.eps = (float)eps; d.s.weight_decay = (float)t.weight_decay;
        d.s.neg_step_size = (float)(-(t.lr / bc1));
        d.s.bc2_sqrt = (float)sqrt(bc2);
        if (b.n_tensors == 0) b.chunk_begin[0] = 0;
        b.chunk_begin[b.n_tensors + 1] = b.chunk_begin[b.n_tensors] + (int)chunks;
        b.n_tensors++;
    }
    NVR_CHECK(h, flush());
    return 0;
}

extern "C" size_t nvr_rays_workspace_bytes(int32_t H, int32_t W) {
    const long long tiles = ((long long)(H > 0 ? H : 0) * (W > 0 ? W : 0) + RAYS_BLOCK - 1) / RAYS_BLOCK;
    return (size_t)(2 * tiles + 2) * sizeof(int);
}

extern "C" int nvr_generate_rays(NvrHandle h, int32_t H, int32_t W, const double* K_inv, const double* R, const double* T,
                                 const float* bounds, float* ray_o, float* ray_d, float* near_, float* far_, int32_t* coord,
                                 uint8_t* mask_at_box, int32_t* n_rays, void* workspace, size_t ws_bytes, void* stream_) {
    if (!h) return 1;
    if (H < 1 || W < 1 || (long long)H * W >= (1ll << 31)) return fail(h, "nvr_generate_rays: bad image size");
    if (!K_inv || !R || !T || !bounds || !ray_o || !ray_d || !near_ || !far_ || !mask_at_box || !n_rays)
        return fail(h, "nvr_generate_rays: null argument");
    if (!workspace || ws_bytes < nvr_rays_workspace_bytes(H, W) || ((uintptr_t)workspace & 3))
        return fail(h, "nvr_generate_rays: workspace too small (nvr_rays_workspace_bytes)");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    CameraDev cam;
    for (int i = 0; i < 9; ++i) { cam.Kinv[i] = K_inv[i]; cam.R[i] = R[i]; }
    for (int a = 0; a < 3; ++a) {
        cam.T[a] = T[a];
        cam.o[a] = -((R[0 * 3 + a] * T[0] + R[1 * 3 + a] * T[1]) + R[2 * 3 + a] * T[2]);     // -R^T T (:26)
    }
    const int tiles = (int)(((long long)H * W + RAYS_BLOCK - 1) / RAYS_BLOCK);
    int* tile_count = (int*)workspace;
    int* tile_off = tile_count + tiles;
    cudaStream_t st = (cudaStream_t)stream_;
    k_rays_mask<<<tiles, RAYS_BLOCK, 0, st>>>(cam, H, W, bounds, mask_at_box, tile_count);
    k_rays_scan<<<1, 1024, 0, st>>>(tile_count, tiles, tile_off, n_rays);
    k_rays_emit<<<tiles, RAYS_BLOCK, 0, st>>>(cam, H, W, bounds, tile_off, ray_o, ray_d, near_, far_, coord);
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 3;
    return 0;
}

extern "C" int nvr_assemble_image(NvrHandle h, const float* rgb, const int32_t* coord, int64_t n_rays, int64_t n_pixels, float* img,
                                  void* stream_) {
    if (!h) return 1;
    if (n_rays < 0 || n_pixels < 0 || n_rays > n_pixels || (n_pixels > 0 && !img) || (n_rays > 0 && (!rgb || !coord)))
        return fail(h, "nvr_assemble_image: bad argument");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream_;
    NVR_CHECK(h, cudaMemsetAsync(img, 0, (size_t)n_pixels * 3 * sizeof(float), st));
    if (n_rays == 0) return 0;
    k_assemble_image<<<grid_for(n_rays, 256, h->sm_count * 8), 256, 0, st>>>(rgb, coord, n_rays, img);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int nvr_sq_diff_sum(NvrHandle h, const float* a, const float* b, int64_t n, double* sq_sum, void* stream_) {
    if (!h) return 1;
    if (n < 0 || !sq_sum || (n > 0 && (!a || !b))) return fail(h, "nvr_sq_diff_sum: bad argument");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream_;
    NVR_CHECK(h, cudaMemsetAsync(sq_sum, 0, sizeof(double), st));
    if (n == 0) return 0;
    k_sq_diff<<<grid_for(n, 256 * 8, h->sm_count * 8), 256, 0, st>>>(a, b, n, sq_sum);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int nvr_ssim_sums(NvrHandle h, const float* img_a, const float* img_b, int32_t H, int32_t W, int32_t x0, int32_t y0,
                             int32_t w, int32_t hh, double* sums, void* stream_) {
    if (!h) return 1;
    if (!img_a || !img_b || !sums || H < 1 || W < 1 || x0 < 0 || y0 < 0 || w < 1 || hh < 1 || x0 + w > W || y0 + hh > H)
        return fail(h, "nvr_ssim_sums: bad argument");
    if (w < 7 || hh < 7) return fail(h, "nvr_ssim_sums: the crop must be at least 7 x 7 (win_size exceeds image extent)");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream_;
    NVR_CHECK(h, cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));
    k_ssim<<<grid_for((long long)(w - 6) * (hh - 6) * 3, 256, h->sm_count * 8), 256, 0, st>>>(img_a, img_b, W, x0, y0, w, hh, sums);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// ---- per-frame SMPL preprocessing ---------------------------------------------------------
extern "C" size_t nvr_smpl_workspace_bytes(int32_t n_verts) {
    return n_verts < 0 ? 0 : (size_t)SMPL_WS_PXYZ + (size_t)n_verts * 3 * sizeof(double);
}

extern "C" int nvr_smpl_pose_frame(NvrHandle h, const NvrSmplPose* pose, const float* wxyz, int32_t n_verts, const int32_t* vert_slot,
                                   int32_t maxlen, float box_padding, const NvrSmplOut* out, void* workspace, size_t ws_bytes,
                                   void* stream_) {
    if (!h) return 1;
    if (!pose || !wxyz || !out || n_verts < 1) return fail(h, "nvr_smpl_pose_frame: null argument / no vertices");
    if (!out->R || !out->Th || !out->ppts) return fail(h, "nvr_smpl_pose_frame: R, Th and ppts outputs are required");
    if ((vert_slot != nullptr) != (out->part_pts != nullptr) || (vert_slot && maxlen < 1))
        return fail(h, "nvr_smpl_pose_frame: vert_slot, maxlen and part_pts go together");
    if (!workspace || ((uintptr_t)workspace & 255) || ws_bytes < nvr_smpl_workspace_bytes(n_verts))
        return fail(h, "nvr_smpl_pose_frame: workspace too small or not 256-byte aligned (nvr_smpl_workspace_bytes)");
    for (int j = 1; j < NVR_JOINTS; ++j)
        if (pose->parents[j] < 0 || pose->parents[j] >= j) return fail(h, "nvr_smpl_pose_frame: parents must precede their children");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    static_assert(sizeof(SmplPoseDev) == sizeof(NvrSmplPose), "SmplPoseDev mirrors NvrSmplPose");
    SmplPoseDev P;
    memcpy(&P, pose, sizeof(P));
    P.parents[0] = 0;
    cudaStream_t st = (cudaStream_t)stream_;
    unsigned char* ws = (unsigned char*)workspace;
    k_smpl_transforms<<<1, 64, 0, st>>>(P, out->A, out->big_A, out->R, out->Th, ws);
    k_smpl_pose_verts<<<(n_verts + 255) / 256, 256, 0, st>>>(wxyz, n_verts, out->R, out->Th, vert_slot, out->ppts, out->part_pts, ws);
    k_smpl_bounds<<<1, 32, 0, st>>>(ws, box_padding, out->pbounds, out->wbounds);
    NVR_CHECK(h, cudaGetLastError());
    h->launches += 3;
    return 0;
}

extern "C" int nvr_smpl_volume_dims(NvrHandle h, const void* workspace, int32_t dims[3], double origin[3], void* stream_) {
    if (!h) return 1;
    if (!workspace || !dims || !origin) return fail(h, "nvr_smpl_volume_dims: null argument");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    unsigned long long box[6];
    cudaStream_t st = (cudaStream_t)stream_;
    NVR_CHECK(h, cudaMemcpyAsync(box, (const unsigned char*)workspace + SMPL_WS_F64BOX, sizeof(box), cudaMemcpyDeviceToHost, st));
    NVR_CHECK(h, cudaStreamSynchronize(st));
    for (int a = 0; a < 3; ++a) {
        // get_grid_points (tools/prepare_zjumocap.py:152-165): min -= 0.05, max += 0.05, arange(min, max + vsize, vsize)
        const double lo = unord64(box[a]) - 0.05, hi = unord64(box[3 + a]) + 0.05;
        if (!(lo <= hi)) return fail(h, "nvr_smpl_volume_dims: empty / non-finite vertex bbox (run nvr_smpl_pose_frame first)");
        origin[a] = lo;
        dims[a] = nvr_arange_len(lo, hi + 0.025, 0.025);
    }
    return 0;
}

extern "C" int nvr_smpl_bweights(NvrHandle h, const void* workspace, int32_t n_verts, const float* weights, const int32_t dims[3],
                                 const double origin[3], float* pbw, void* stream_) {
    if (!h) return 1;
    if (!workspace || !weights || !dims || !origin || !pbw || n_verts < 1) return fail(h, "nvr_smpl_bweights: null argument");
    const long long n_vox = (long long)dims[0] * dims[1] * dims[2];
    if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || n_vox > (1ll << 28)) return fail(h, "nvr_smpl_bweights: bad volume dims");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream_;
    k_smpl_bweights<<<(unsigned)((n_vox + 255) / 256), 256, 0, st>>>((const unsigned char*)workspace, n_verts, weights, dims[0], dims[1],
                                                                     dims[2], origin[0], origin[1], origin[2], 0.025, pbw);
    NVR_CHECK(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// ---- gather footprint (measurement aid) ----------------------------------------------------------------------
extern "C" int nvr_gather_footprint(NvrHandle h, void* workspace, size_t ws_bytes, int64_t* unique_sectors_host, void* stream_) {
    if (int rc = ready(h, "nvr_gather_footprint")) return rc;
    if (!unique_sectors_host) return fail(h, "nvr_gather_footprint: null argument");
    // the layout the most recent call used: one workspace, or the two halves of a two-lane render (one pass each; the union
    // of both halves' pair lists is the frame's footprint)
    Workspace lw[2];
    const int lanes = h->last_lanes;
    const size_t half_bytes = (ws_bytes / 2) & ~(size_t)255;
    if (lanes == 2 ? !(carve(workspace, half_bytes, lw[0]) && carve((char*)workspace + half_bytes, half_bytes, lw[1]))
                   : !carve(workspace, ws_bytes, lw[0]))
        return fail(h, "nvr_gather_footprint: not a pass workspace");
    if (h->last_passes > lanes) return fail(h, "nvr_gather_footprint: the last call ran more than one pass per lane");
    cudaStream_t st = (cudaStream_t)stream_;
    size_t max_words = 0;
    for (int p = 0; p < NVR_NUM_PARTS; ++p) {
        const NvrGrid& g = h->params.part[p].grid;
        max_words = std::max(max_words, (size_t)((2 * (dense_rows(g) + hash_rows(g)) + 31) / 32));
    }
    unsigned int* bitmap = nullptr;
    unsigned long long* d_out = nullptr;
    NVR_CHECK(h, cudaMalloc(&bitmap, max_words * sizeof(unsigned int)));
    if (cudaMalloc(&d_out, NVR_NUM_PARTS * sizeof(unsigned long long)) != cudaSuccess) { cudaFree(bitmap); return fail(h, "nvr_gather_footprint: out of memory"); }
    cudaMemsetAsync(d_out, 0, NVR_NUM_PARTS * sizeof(unsigned long long), st);
    for (int p = 0; p < NVR_NUM_PARTS; ++p) {
        const NvrGrid& g = h->params.part[p].grid;
        const size_t words = (size_t)((2 * (dense_rows(g) + hash_rows(g)) + 31) / 32);
        cudaMemsetAsync(bitmap, 0, words * sizeof(unsigned int), st);
        for (int l = 0; l < h->last_lanes_used; ++l) {
            const Workspace& w = lw[l];
            k_embed_footprint<<<h->sm_count * 8, 256, 0, st>>>(h->part_grid[p], (const float*)(w.pairs + (long long)p * w.cap), 8,
                                                               w.counters + NVR_CTR_PAIR + p, bitmap, 2ull * (unsigned long long)dense_rows(g));
        }
        k_popcount_words<<<h->sm_count * 4, 256, 0, st>>>(bitmap, (long long)words, d_out + p);
    }
    unsigned long long host[NVR_NUM_PARTS];
    cudaError_t e = cudaMemcpyAsync(host, d_out, sizeof(host), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(bitmap); cudaFree(d_out);
    if (e != cudaSuccess) { h->err = std::string("nvr_gather_footprint: ") + cudaGetErrorString(e); return 2; }
    for (int p = 0; p < NVR_NUM_PARTS; ++p) unique_sectors_host[p] = (int64_t)host[p];
    h->launches += 2 * NVR_NUM_PARTS;
    return 0;
}

extern "C" int nvr_profile(NvrHandle h, int32_t enable) {
    if (!h) return 1;
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    if (enable && !h->h_pass_counters)
        NVR_CHECK(h, cudaMallocHost(&h->h_pass_counters, (size_t)PROF_MAX_PASSES * NVR_CTR_WORDS * sizeof(int)));
    h->profiling = enable != 0;
    h->ev_used = 0; h->spans.clear(); h->n_pass_snap = 0;
    return 0;
}

extern "C" int nvr_profile_read(NvrHandle h, NvrStageProfile* out) {
    if (!h || !out) return fail(h, "nvr_profile_read: null argument");
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    NVR_CHECK(h, cudaDeviceSynchronize());
    memset(out, 0, sizeof(*out));
    for (const auto& sp : h->spans) {
        float ms = 0.f;
        NVR_CHECK(h, cudaEventElapsedTime(&ms, h->ev_pool[sp.e0], h->ev_pool[sp.e1]));
        out->ms[sp.stage] += ms;
        out->launches[sp.stage]++;
        if (sp.part >= 0 && sp.stage == NVR_STAGE_EMBED) out->embed_part_ms[sp.part] += ms;
        if (sp.part >= 0 && sp.stage == NVR_STAGE_MLP) out->mlp_part_ms[sp.part] += ms;
    }
    out->passes = h->n_pass_snap;
    for (int i = 0; i < h->n_pass_snap; ++i) {
        const int* c = h->h_pass_counters + (size_t)i * NVR_CTR_WORDS;
        out->survivors += c[NVR_CTR_SURV];
        for (int p = 0; p < NVR_NUM_PARTS; ++p) { out->pairs[p] += c[NVR_CTR_PAIR + p]; out->far_pairs[p] += c[NVR_CTR_FAR + p]; }
    }
    return 0;
}

extern "C" int nvr_read_counters(NvrHandle h, NvrCounters* out, void* stream_) {
    if (!h || !out) return fail(h, "nvr_read_counters: null argument");
    int host[NVR_CTR_WORDS];
    NVR_CHECK(h, cudaSetDevice(h->cfg.device));
    NVR_CHECK(h, cudaMemcpyAsync(host, h->d_counters_snapshot, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    NVR_CHECK(h, cudaStreamSynchronize((cudaStream_t)stream_));
    out->n_points = h->last_points;
    out->n_survivors = host[NVR_CTR_SURV];
    for (int p = 0; p < NVR_NUM_PARTS; ++p) { out->n_pairs[p] = host[NVR_CTR_PAIR + p]; out->n_far_pairs[p] = host[NVR_CTR_FAR + p]; }
    out->kernel_launches = h->launches;
    out->n_passes = h->last_passes;
    return 0;
}

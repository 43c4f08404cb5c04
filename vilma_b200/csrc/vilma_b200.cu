// libvilma_b200.so -- host side of the C ABI declared in include/vilma_b200.h.
// Builds the HBM layout of the LD store, owns the fit state and sequences the kernels
// in ld_kernels.cuh / snp_kernels.cuh on one CUDA stream.  sm_100a only; no CPU path.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vilma_b200.h"
#include "ld_kernels.cuh"
#include "snp_kernels.cuh"
#include "snp_tile_kernel.cuh"
#include "setup_kernels.cuh"

static_assert(VB_MAX_POPS <= VB_MAXP, "header / kernel cohort limits disagree");

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int vb_fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return vb_fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                     \
    } while (0)
#define CK_LAUNCH(ctx)                                                                    \
    do {                                                                                  \
        (ctx)->launches++;                                                                \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess)                                                           \
            return vb_fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),   \
                           __FILE__, __LINE__);                                           \
    } while (0)

static inline int64_t even_up(int64_t v) { return (v + 1) & ~int64_t(1); }
// Measured on C2 (tools/ld_bench.py): claiming groups in natural (memory) order beats largest-first
// (0.696 vs 0.736 ms) and capping the groups per block does not pay either.
#ifndef VB_SYM_MAX_GROUPS
#define VB_SYM_MAX_GROUPS 1000000
#endif
#ifndef VB_SYM_LPT
#define VB_SYM_LPT 0
#endif
static bool g_disable_sym = false;
static int g_tile_mode = -1;         // vb_set_option("snp_tile", ...): see tile_plan()
static bool g_ann_slots = true;      // vb_set_option("snp_ann_slots", 0): fused annotation sums by warp shuffles only
static bool g_snp3_park = true;      // vb_set_option("snp3_park", 0): three-pass kernel parks logits in the output buffers
static int g_fused_finish = -1;      // vb_set_option("ld_fused_finish", v): -1 automatic, 0 separate finish kernel, 1 always fused
static bool g_shard_zero_copy = true;   // vb_set_option("shard_zero_copy", 0): sharded uploads always through host staging
static bool g_factor_once = true;    // vb_set_option("ld_factor_once", 0): factor blocks always in the two-pass form
static bool g_three_pass = true;     // vb_set_option("snp_three_pass", 0): always the online single-pass kernel   // vb_set_option("ld_symmetric", 0): store dense blocks in full

// ------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------
struct LdSlab {          // a row-major matrix rows x ld (ld even), the unit the items cut up
    size_t off;          // offset in doubles into LdPop::mat
    int64_t rows, cols, ld;
    int64_t x_off;       // offset (doubles, even) into xall of the x segment
    int64_t y_off;       // output offset (doubles) of row 0
};
struct SymSlab {                   // one column slab of a symmetric-packed block (see ld_kernels.cuh)
    int64_t J0, w;                 // first column within the block, width of its diagonal tile
    size_t off;                    // offset (doubles) of its panels in LdPop::mat
    uint32_t g0, ng;               // its groups [g0, g0 + ng): the tile's first, then those below the tile
    uint32_t gb0;                  // first group below the tile
    int64_t Rg;                    // rows per group below the tile
};
struct LdBlock {
    int64_t n, r;                  // r < 0: dense
    int64_t xpos, tpos;            // padded block-order / rank-space offsets
    std::vector<int> slabs1;       // indices into LdPop::slabs (phase 1: V' = diag(s) U^T)
    std::vector<int> slabs2;       // phase 2: dense R or U
    bool filled = false;
    bool sym = false;              // dense block stored symmetric-packed
    bool fac1 = false;             // factor block stored read-once (U sqrt(s), chunked column-major)
    size_t fac_off = 0;            // its offset (doubles) in LdPop::mat
    int64_t n_pad = 0;
    int fac_c = 0;                 // columns per chunk
    uint32_t fg0 = 0, fng = 0;     // its groups in LdPop::fgroups (and, after gout_fbase, in gout)
    std::vector<SymSlab> sslabs;   // its column slabs (one unless n > VB_SYM_NMAX)
};
struct LdPop {
    bool begun = false, finalized = false;
    int64_t M = 0, xb_len = 0, tb_len = 0;
    int nslab1 = 1, nslab2 = 1;    // max column slabs per phase
    std::vector<LdBlock> blocks;
    std::vector<LdSlab> slabs;
    double* mat = nullptr;
    size_t mat_len = 0;
    double* xall = nullptr;        // [xb | tb]
    double* yb = nullptr;          // [nslab2][xb_len]
    double* tbs = nullptr;         // [nslab1][tb_len] when nslab1 > 1
    VbLdItem *items1 = nullptr, *items2 = nullptr;
    uint32_t* sched = nullptr;     // [2][2] dynamic-scheduling counters per phase
    int64_t n_items1 = 0, n_items2 = 0;
    int32_t *pos = nullptr, *snp = nullptr;
    int32_t* xbpos = nullptr;      // [M] block-order position of each SNP (-1: not in this LD)
    uint32_t* fin_counter = nullptr;
    int64_t nreal = 0;
    // symmetric-packed blocks
    VbSymItem* sitems = nullptr;
    VbSymGroup* sgroups = nullptr;
    VbSymGroupOut* gout = nullptr;
    VbSymBlockRef* bref = nullptr;
    int64_t n_sgroups = 0;
    // read-once factor blocks (vb_ld_fac_kernel); their partial vectors follow the symmetric groups' in gout
    VbFacItem* fitems = nullptr;
    VbSymGroup* fgroups = nullptr;
    int64_t n_fgroups = 0;
    uint32_t gout_fbase = 0;
    double* ypart = nullptr;
    VbFinRec* finrec = nullptr;    // per block-order position: finish-kernel record
    uint32_t *xstart = nullptr, *xoffs = nullptr;   // wide blocks: row sums a position takes from slabs to its left
    // finish fused into the symmetric mat-vec (all blocks symmetric-packed)
    bool all_sym = false;
    VbSymBlockFin* bfin = nullptr;
    uint32_t *block_cnt = nullptr, *done_cnt = nullptr;
    double* part_blk = nullptr;
    std::vector<uint32_t> block_ng;
    std::vector<VbSymGroupOut> gout_host;
    int64_t bytes = 0;             // algorithmic bytes per mat-vec
};

struct Fit {
    bool created = false;
    int K = 0, P = 0, A = 0;
    int64_t M = 0;
    double *adj = nullptr, *se = nullptr, *sld = nullptr, *scal = nullptr;
    int32_t* ann = nullptr;
    double *prec = nullptr, *logdet = nullptr, *logh = nullptr, *gfull = nullptr, *inv_tau_dev = nullptr;
    double inv_tau[VB_MAXP];
    double* mu[2] = {nullptr, nullptr};
    double* delta[2] = {nullptr, nullptr};
    double *pm[2] = {nullptr, nullptr}, *linked[2] = {nullptr, nullptr};
    int cur_mu = 0, cur_delta = 0, cur_vec = 0;
    int trial_kind = -1;           // -1 none, 0 beta trial (new mu+delta), 1 refresh (new delta)
    double* scratch3 = nullptr;    // [3][P][M]
    double *pm_prev = nullptr, *pm_ckpt = nullptr, *pm_next = nullptr;
    int akf = 0, nsp = 0;          // fused annotation sums per evaluation; partial row stride
    int64_t mutations = 0;         // bumped by every public vb_fit_* call (guards speculative work)
    double* part_snp = nullptr;
    int grid_snp = 0;
    int snp_grid_used = 0;         // CTAs of the last per-SNP launch (= rows of part_snp it wrote)
    double* part_fin = nullptr;
    int grid_fin = 0;
    double* part_ann = nullptr;
    int grid_ann = 0;
    double* part_diff = nullptr;
    int grid_diff = 0;
    int64_t* shard_idx = nullptr;  // global SNP index of each local SNP (vb_fit_set_shard)
    int64_t shard_total = 0;
    std::vector<int64_t> shard_runs;   // (global start, local start, length) of runs of consecutive SNPs
    double* shard_stage = nullptr;     // page-locked staging: K (P+1) M doubles
};

#define VB_PROF_CATS 4
struct vb_ld;
struct vb_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    std::vector<vb_ld*> fit_ld;   // the P operators the fit state uses (not owned)
    Fit fit;
    // optional per-kernel timing with CUDA events on `stream` (bench.py roofline leg)
    bool profiling = false;
    // category 0: LD mat-vec, 1: per-SNP kernel, 2: mat-vec finish (+ final reduction / rank exchange),
    // 3: bookkeeping kernels (annotation sums, convergence partials); start/stop pairs
    std::vector<cudaEvent_t> ev[VB_PROF_CATS];
    size_t ev_used[VB_PROF_CATS] = {0, 0, 0, 0};
};
#define VB_PROF_PAIRS 32768
static inline void prof_begin(vb_ctx* c, int cat) {
    if (c->profiling && c->ev_used[cat] + 2 <= c->ev[cat].size())
        cudaEventRecord(c->ev[cat][c->ev_used[cat]], c->stream);
}
static inline void prof_end(vb_ctx* c, int cat) {
    if (c->profiling && c->ev_used[cat] + 2 <= c->ev[cat].size()) {
        cudaEventRecord(c->ev[cat][c->ev_used[cat] + 1], c->stream);
        c->ev_used[cat] += 2;
    }
}
struct vb_ld {
    vb_ctx* ctx = nullptr;
    LdPop L;
};

#include <dlfcn.h>
#include <atomic>
#include <chrono>

#define VB_L_MAX 1e12
#define VB_REL_TOL 1e-6
#define VB_ABS_TOL 1e-6
#define VB_EM_TOL 10.0
#define VB_MAX_NUM_ITERS 20

namespace {
typedef struct { char internal[128]; } vbNcclUniqueId;
typedef void* vbNcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(vbNcclUniqueId*) = nullptr;
    int (*CommInitRank)(vbNcclComm*, int, vbNcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, vbNcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(vbNcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const int kNcclDouble = 8, kNcclSum = 0, kNcclMax = 2;   // ncclFloat64 / ncclSum / ncclMax

int load_nccl() {
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return vb_fail("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(vbNcclUniqueId*))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(vbNcclComm*, int, vbNcclUniqueId, int))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, vbNcclComm, cudaStream_t))dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(vbNcclComm))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce)
        return vb_fail("NCCL symbols missing in the loaded libnccl");
    return 0;
}
}  // namespace

struct NativeLoop {
    bool ready = false;
    std::vector<double> chi, ranks, counts, logdet, tau, hyper;
    int scale_se = 0;
    double* pinned = nullptr;     // host staging
    size_t pinned_len = 0;
    double *stats_dev = nullptr, *ann_dev = nullptr, *diff_dev = nullptr;
    vbNcclComm comm = nullptr;
    int nranks = 1;
    // cross-rank mailbox exchange (vb_common.cuh VbXrank)
    bool xr_ready = false;        // mailboxes exchanged (or single rank): host-flag signalling usable
    bool xr_active = false;       // set by vb_fit_iteration while it drives the evaluations
    bool want_diff = false;       // next evaluation also reduces the convergence partials
    double diff_atol = 0.0, diff_rtol = 0.0;
    int xr_nranks = 1, xr_rank = 0;
    unsigned char* xr_box = nullptr;                 // my mailbox (device): data + flags
    double* xr_peer_box[VB_XR_MAXRANKS] = {nullptr};
    uint32_t* xr_peer_flag[VB_XR_MAXRANKS] = {nullptr};
    double *xr_host_out = nullptr, *xr_host_out_dev = nullptr;
    uint32_t *xr_host_flag = nullptr, *xr_host_flag_dev = nullptr;
    uint32_t* xr_dev_err = nullptr;
    uint32_t epoch = 0;
    // a trial evaluation queued speculatively behind the last refresh of the previous iteration
    bool spec_pending = false;
    uint32_t spec_epoch = 0;
    double spec_step = 0.0;
    int64_t spec_mutations = 0;
    int64_t spec_used = 0, spec_wasted = 0;
    // host-side timing of the native loop (seconds): enqueueing evaluations / waiting for results
    double t_enqueue = 0.0, t_wait = 0.0;
    int64_t n_wait = 0;
};
#define VB_XR_DATA_BYTES (2 * VB_XR_MAXRANKS * VB_XR_MAXVALS * sizeof(double))
#define VB_XR_BOX_BYTES (VB_XR_DATA_BYTES + 2 * VB_XR_MAXRANKS * sizeof(uint32_t) + 64)
static std::vector<std::pair<vb_ctx*, NativeLoop*>> g_loops;
static NativeLoop* loop_of(vb_ctx* ctx, bool create) {
    for (auto& pr : g_loops)
        if (pr.first == ctx) return pr.second;
    if (!create) return nullptr;
    NativeLoop* nl = new NativeLoop();
    g_loops.emplace_back(ctx, nl);
    return nl;
}



// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
extern "C" int vb_abi_version(void) { return VB_ABI_VERSION; }
#ifndef VB_SOURCE_HASH
#define VB_SOURCE_HASH "unknown"
#endif
static const char g_source_hash[] = "VB_SOURCE_HASH=" VB_SOURCE_HASH;
extern "C" const char* vb_source_hash(void) { return g_source_hash + 15; }
// Process-wide options read when an LD operator is created.
//   "ld_symmetric" (default 1): store dense blocks with n <= VB_SYM_NMAX symmetric-packed.
//   "snp_three_pass" (default 1): P <= 2 updates use the exact-max three-pass softmax kernel.
//   "snp3_park" (default 1): the three-pass kernel keeps logits / weights / mu' in shared memory when
//       K (P+1) KB fits 32 KB per CTA.
//   "snp_tile" (default -1 = automatic): the K-split tile kernel (snp_tile_kernel.cuh) with W warps per
//       32-SNP tile; 0 = never, W > 0 = always with that many warps.
extern "C" int64_t vb_ld_sym_nmax(void) { return VB_SYM_BLOCK_MAX; }
extern "C" int64_t vb_ld_fac_nmax(void) { return g_factor_once ? VB_SYM_NMAX : 0; }
extern "C" int vb_set_option(const char* name, int64_t value) {
    if (name && std::strcmp(name, "ld_symmetric") == 0) {
        g_disable_sym = (value == 0);
        return 0;
    }
    if (name && std::strcmp(name, "shard_zero_copy") == 0) {
        g_shard_zero_copy = (value != 0);
        return 0;
    }
    if (name && std::strcmp(name, "ld_factor_once") == 0) {
        g_factor_once = (value != 0);
        return 0;
    }
    if (name && std::strcmp(name, "ld_fused_finish") == 0) {
        g_fused_finish = (int)value;
        return 0;
    }
    if (name && std::strcmp(name, "snp_three_pass") == 0) {
        g_three_pass = (value != 0);
        return 0;
    }
    if (name && std::strcmp(name, "snp_ann_slots") == 0) {
        g_ann_slots = (value != 0);
        return 0;
    }
    if (name && std::strcmp(name, "snp3_park") == 0) {
        g_snp3_park = (value != 0);
        return 0;
    }
    if (name && std::strcmp(name, "snp_tile") == 0) {       // -1 auto, 0 never, W = 1,2,4,8,16 forced
        if (value > VB_TILE_MAXW) return vb_fail("vb_set_option: snp_tile takes -1, 0 or W <= %d", VB_TILE_MAXW);
        g_tile_mode = (int)value;
        return 0;
    }
    return vb_fail("vb_set_option: unknown option '%s'", name ? name : "(null)");
}
extern "C" const char* vb_last_error(void) { return g_err.c_str(); }

extern "C" int vb_ctx_create(int device, void* stream, vb_ctx** out) {
    if (!out) return vb_fail("vb_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return vb_fail("vb_ctx_create: no CUDA device available (%s); vilma_b200 has no CPU path",
                       cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return vb_fail("vb_ctx_create: bad device %d", device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return vb_fail("vb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
    vb_ctx* c = new vb_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    CK(cudaFuncSetAttribute(vb_ld_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            VB_LD_SMEM));
    CK(cudaFuncSetAttribute(vb_ld_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            VB_SYM_SMEM));
    CK(cudaFuncSetAttribute(vb_ld_sym_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(vb_ld_fac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VB_FAC_SMEM));
    CK(cudaFuncSetAttribute(vb_ld_fac_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    *out = c;
    return 0;
}

static void free_ld(LdPop& L) {
    cudaFree(L.mat); cudaFree(L.xall); cudaFree(L.yb); cudaFree(L.tbs);
    cudaFree(L.items1); cudaFree(L.items2); cudaFree(L.sched);
    cudaFree(L.pos); cudaFree(L.snp); cudaFree(L.xbpos); cudaFree(L.fin_counter);
    cudaFree(L.sitems); cudaFree(L.sgroups); cudaFree(L.gout); cudaFree(L.bref); cudaFree(L.ypart);
    cudaFree(L.fitems); cudaFree(L.fgroups);
    cudaFree(L.finrec); cudaFree(L.xstart); cudaFree(L.xoffs);
    cudaFree(L.bfin); cudaFree(L.block_cnt); cudaFree(L.done_cnt); cudaFree(L.part_blk);
    L = LdPop();
}
static void free_fit(Fit& f) {
    cudaFree(f.adj); cudaFree(f.se); cudaFree(f.sld); cudaFree(f.scal); cudaFree(f.ann);
    cudaFree(f.prec); cudaFree(f.logdet); cudaFree(f.logh); cudaFree(f.inv_tau_dev);   // gfull lives in logh's allocation
    cudaFree(f.shard_idx);
    if (f.shard_stage) cudaFreeHost(f.shard_stage);
    for (int s = 0; s < 2; ++s) {
        cudaFree(f.mu[s]); cudaFree(f.delta[s]); cudaFree(f.pm[s]);
        cudaFree(f.linked[s]);
    }
    cudaFree(f.scratch3); cudaFree(f.pm_prev); cudaFree(f.pm_ckpt); cudaFree(f.pm_next);
    cudaFree(f.part_snp); cudaFree(f.part_fin); cudaFree(f.part_ann); cudaFree(f.part_diff);
    f = Fit();
}

extern "C" int vb_ctx_destroy(vb_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_fit(ctx->fit);
    for (int cat = 0; cat < 2; ++cat)
        for (auto& e : ctx->ev[cat]) cudaEventDestroy(e);
    for (size_t i = 0; i < g_loops.size(); ++i) {
        if (g_loops[i].first != ctx) continue;
        NativeLoop* nl = g_loops[i].second;
        if (nl->pinned) cudaFreeHost(nl->pinned);
        cudaFree(nl->stats_dev); cudaFree(nl->ann_dev); cudaFree(nl->diff_dev);
        for (int r = 0; r < nl->xr_nranks; ++r)
            if (nl->xr_ready && r != nl->xr_rank && nl->xr_peer_box[r]) cudaIpcCloseMemHandle(nl->xr_peer_box[r]);
        cudaFree(nl->xr_box);
        if (nl->xr_host_out) cudaFreeHost(nl->xr_host_out);
        if (nl->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(nl->comm);
        delete nl;
        g_loops.erase(g_loops.begin() + i);
        break;
    }
    delete ctx;
    return 0;
}
extern "C" int vb_ctx_profile(vb_ctx* ctx, int enable) {
    if (!ctx) return vb_fail("null ctx");
    CK(cudaSetDevice(ctx->device));
    if (enable && ctx->ev[0].empty()) {
        for (int cat = 0; cat < VB_PROF_CATS; ++cat) {
            ctx->ev[cat].resize(2 * VB_PROF_PAIRS);
            for (auto& e : ctx->ev[cat]) CK(cudaEventCreate(&e));
        }
    }
    ctx->profiling = enable != 0;
    for (int cat = 0; cat < VB_PROF_CATS; ++cat) ctx->ev_used[cat] = 0;
    return 0;
}
// total_ms[cat], count[cat] of the launches timed since profiling was (re-)enabled; resets.
extern "C" int vb_ctx_profile_read(vb_ctx* ctx, double* total_ms, int64_t* count) {
    if (!ctx) return vb_fail("null ctx");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int cat = 0; cat < VB_PROF_CATS; ++cat) {
        double tot = 0.0;
        for (size_t i = 0; i + 1 < ctx->ev_used[cat]; i += 2) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ctx->ev[cat][i], ctx->ev[cat][i + 1]));
            tot += ms;
        }
        total_ms[cat] = tot;
        count[cat] = (int64_t)(ctx->ev_used[cat] / 2);
        ctx->ev_used[cat] = 0;
    }
    return 0;
}
extern "C" int vb_ctx_sync(vb_ctx* ctx) {
    if (!ctx) return vb_fail("null ctx");
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int64_t vb_ctx_launch_count(const vb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------
// LD store
// ------------------------------------------------------------------------------------
static void add_slabs(LdPop& L, std::vector<int>& into, int64_t rows, int64_t cols, int64_t x_base,
                      int64_t y_base, int64_t y_slab_stride, size_t& cursor, int& nslab_max) {
    const int nsl = (int)((cols + VB_LD_CMAX - 1) / VB_LD_CMAX);
    int64_t width = even_up((cols + nsl - 1) / nsl);
    for (int s = 0; s < nsl; ++s) {
        const int64_t c0 = s * width;
        const int64_t c1 = std::min(cols, c0 + width);
        LdSlab sl;
        sl.off = cursor;
        sl.rows = rows;
        sl.cols = c1 - c0;
        sl.ld = even_up(c1 - c0);
        sl.x_off = x_base + c0;
        sl.y_off = y_base + (int64_t)s * y_slab_stride;
        cursor += (size_t)rows * sl.ld;
        into.push_back((int)L.slabs.size());
        L.slabs.push_back(sl);
    }
    nslab_max = std::max(nslab_max, nsl);
}

static int build_items(vb_ctx* ctx, LdPop& L, int phase, VbLdItem** d_items, int64_t* n_items) {
    std::vector<VbLdItem> items;
    for (auto& b : L.blocks) {
        const std::vector<int>& sl = phase == 1 ? b.slabs1 : b.slabs2;
        for (int si : sl) {
            const LdSlab& s = L.slabs[si];
            int64_t rows_per = std::max<int64_t>(1, VB_LD_STAGE_A / (s.ld * 8));
            rows_per = std::min<int64_t>(rows_per, 65535);
            for (int64_t r0 = 0; r0 < s.rows; r0 += rows_per) {
                const int64_t nr = std::min(rows_per, s.rows - r0);
                VbLdItem it;
                const size_t a_off = s.off + (size_t)r0 * s.ld;          // doubles, even
                if ((a_off >> 1) > 0xffffffffull)
                    return vb_fail("LD store of one cohort exceeds 64 GiB on this rank");
                it.a_off16 = (uint32_t)(a_off >> 1);
                it.x_off2 = (uint32_t)(s.x_off >> 1);
                it.y_off = (uint32_t)(s.y_off + r0);
                it.nrows = (uint16_t)nr;
                it.ld2 = (uint16_t)(s.ld >> 1);
                items.push_back(it);
            }
        }
    }
    *n_items = (int64_t)items.size();
    if (items.size() > 0xfffffff0ull) return vb_fail("too many LD work items");
    if (!items.empty()) {
        CK(cudaMalloc(d_items, items.size() * sizeof(VbLdItem)));
        CK(cudaMemcpy(*d_items, items.data(), items.size() * sizeof(VbLdItem), cudaMemcpyHostToDevice));
    }
    return 0;
}

static int ld_begin(vb_ctx* ctx, LdPop& L, int64_t M, int64_t nblocks, const int64_t* n,
                    const int64_t* rank);

extern "C" int vb_ld_create(vb_ctx* ctx, int64_t M, int64_t nblocks, const int64_t* n,
                            const int64_t* rank, vb_ld** out) {
    if (!ctx || !out) return vb_fail("vb_ld_create: null argument");
    if (M < 1 || nblocks < 0) return vb_fail("vb_ld_create: bad sizes");
    CK(cudaSetDevice(ctx->device));
    vb_ld* h = new vb_ld();
    h->ctx = ctx;
    if (ld_begin(ctx, h->L, M, nblocks, n, rank)) {
        free_ld(h->L);
        delete h;
        return 1;
    }
    *out = h;
    return 0;
}
extern "C" int vb_ld_destroy(vb_ld* ld) {
    if (!ld) return 0;
    cudaSetDevice(ld->ctx->device);
    cudaStreamSynchronize(ld->ctx->stream);
    free_ld(ld->L);
    delete ld;
    return 0;
}

static int ld_begin(vb_ctx* ctx, LdPop& L, int64_t M, int64_t nblocks, const int64_t* n,
                    const int64_t* rank) {
    L.begun = true;
    L.M = M;
    L.blocks.resize(nblocks);
    int64_t xpos = 0, tpos = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        if (n[b] <= 0) return vb_fail("vb_ld_create: block %lld has n=%lld", (long long)b, (long long)n[b]);
        L.blocks[b].n = n[b];
        L.blocks[b].r = rank[b];
        L.blocks[b].xpos = xpos;
        xpos += even_up(n[b]);
        if (rank[b] >= 0) {
            if (rank[b] == 0) return vb_fail("vb_ld_create: factor block %lld has rank 0", (long long)b);
            L.blocks[b].tpos = tpos;
            tpos += even_up(rank[b]);
        }
    }
    L.xb_len = std::max<int64_t>(xpos, 2);
    L.tb_len = tpos;
    if (L.xb_len + L.tb_len > 0x7fffffffll) return vb_fail("vb_ld_create: too many SNPs for 32-bit positions");
    // first pass to learn the slab counts (y strides depend on them)
    size_t cursor = 0;
    int ns1 = 1, ns2 = 1;
    for (auto& b : L.blocks) {
        b.fac1 = b.r >= 0 && g_factor_once && b.n <= VB_SYM_NMAX;
        if (b.fac1) continue;
        if (b.r < 0) ns2 = std::max(ns2, (int)((b.n + VB_LD_CMAX - 1) / VB_LD_CMAX));
        else {
            ns1 = std::max(ns1, (int)((b.n + VB_LD_CMAX - 1) / VB_LD_CMAX));
            ns2 = std::max(ns2, (int)((b.r + VB_LD_CMAX - 1) / VB_LD_CMAX));
        }
    }
    L.nslab1 = ns1;
    L.nslab2 = ns2;
    int dummy1 = 1, dummy2 = 1;
    L.bytes = 0;
    // groups of ~0.5 MB: smaller groups for small (multi-GPU) shards were measured slower (flush cost)
    const size_t group_bytes_base = VB_SYM_GROUP_BYTES;
    std::vector<VbSymItem> sitems;
    std::vector<VbSymGroup> sgroups;
    std::vector<size_t> sgroup_bytes;
    std::vector<VbSymGroupOut> gout;
    std::vector<VbSymBlockRef> bref(L.blocks.size());
    std::vector<VbFacItem> fitems;
    std::vector<VbSymGroup> fgroups;
    std::vector<VbSymGroupOut> fgout;
    size_t ypart_len = 0;
    for (size_t bi = 0; bi < L.blocks.size(); ++bi) {
        LdBlock& b = L.blocks[bi];
        bref[bi].g0 = bref[bi].ng = 0;
        if (b.r < 0 && b.n <= VB_SYM_BLOCK_MAX && !g_disable_sym) {
            // symmetric-packed: column slabs of <= VB_SYM_NMAX columns; within a slab panels of 8 rows,
            // chunks of <= 512 columns, groups of ~0.5 MB (1 MB below the slab's diagonal tile)
            b.sym = true;
            const int64_t n = b.n;
            const int64_t wmax = VB_SYM_NMAX & ~int64_t(7);
            const int64_t T = (n + wmax - 1) / wmax;
            const int64_t wt = T == 1 ? n : ((((n + T - 1) / T) + 7) & ~int64_t(7));
            bref[bi].g0 = (uint32_t)sgroups.size();
            for (int64_t J0 = 0; J0 < n; J0 += wt) {
                SymSlab sl;
                sl.J0 = J0;
                sl.w = std::min<int64_t>(wt, n - J0);
                sl.off = cursor;
                sl.g0 = (uint32_t)sgroups.size();
                const int64_t nr = n - J0, w = sl.w;                   // rows of the slab, tile width
                const int64_t npan = (nr + VB_SYM_R - 1) / VB_SYM_R;
                const int64_t qt = (w + VB_SYM_R - 1) / VB_SYM_R;      // panels of the tile
                if (nr > w && (w % VB_SYM_R)) return vb_fail("internal: ragged tile above full-width panels");
                sl.Rg = std::max<int64_t>(VB_SYM_R, std::min<int64_t>(
                    VB_SYM_GROUP_ROWS, (int64_t)(VB_SYM_BELOW_BYTES / (8 * VB_SYM_R * std::max<int64_t>(w, 1))) * VB_SYM_R));
                const size_t tile_doubles = vb_sym_tile_doubles(w);
                size_t group_bytes = 0;
                int64_t grow0 = 0;
                const size_t group_target = std::max<size_t>(group_bytes_base, (size_t)4 * w * (w + 1) / VB_SYM_MAX_GROUPS);
                VbSymGroup cur;
                cur.first_item = (uint32_t)sitems.size();
                cur.n_items = 0;
                sl.gb0 = 0;
                bool gb0_set = false;
                for (int64_t q = 0; q < npan; ++q) {
                    const int64_t r0 = q * VB_SYM_R;
                    const bool below = q >= qt;
                    const int64_t W = below ? w : even_up(std::min<int64_t>(r0 + VB_SYM_R, w));
                    const size_t poff = below ? sl.off + tile_doubles + (size_t)(q - qt) * VB_SYM_R * w
                                              : sl.off + (size_t)32 * q * (q + 1);
                    if (below && !gb0_set) { sl.gb0 = (uint32_t)sgroups.size(); gb0_set = true; }
                    for (int64_t c0 = 0; c0 < W; c0 += VB_SYM_CC) {
                        const int64_t wc = std::min<int64_t>(VB_SYM_CC, W - c0);
                        VbSymItem it;
                        const size_t a_off = poff + (size_t)c0 * VB_SYM_R;
                        if ((a_off >> 1) > 0xffffffffull)
                            return vb_fail("LD store of one cohort exceeds 64 GiB on this rank");
                        it.a_off16 = (uint32_t)(a_off >> 1);
                        it.x_off2 = (uint32_t)((b.xpos + J0 + c0) >> 1);
                        it.xr_off2 = (uint32_t)((b.xpos + J0 + r0) >> 1);
                        it.wc2 = (uint16_t)(wc >> 1);
                        it.c0_2 = (uint16_t)(c0 >> 1);
                        const int64_t elig = below ? wc : std::max<int64_t>(0, std::min<int64_t>(wc, r0 - c0));
                        it.elig2 = (uint16_t)(elig >> 1);
                        it.flags = VB_SYM_VALID;
                        if (c0 == 0) it.flags |= VB_SYM_FIRST;
                        if (c0 + VB_SYM_CC >= W) it.flags |= VB_SYM_LASTPANEL;
                        it.r0 = (uint16_t)r0;
                        it.grow0 = (uint16_t)grow0;
                        it.out_off = 0;
                        it.out_len = 0;
                        it.nrows_g = 0;
                        it.blk = (uint32_t)bi;
                        it.pad_[0] = it.pad_[1] = it.pad_[2] = 0;
                        sitems.push_back(it);
                        cur.n_items++;
                        group_bytes += (size_t)wc * VB_SYM_R * 8;
                    }
                    const int64_t rows_in_group = r0 + VB_SYM_R - grow0;
                    const bool end_group = q == npan - 1 || q == qt - 1 ||
                        (below ? rows_in_group >= sl.Rg
                               : (group_bytes >= group_target || rows_in_group + VB_SYM_R > VB_SYM_GROUP_ROWS));
                    if (end_group) {
                        VbSymItem& last = sitems.back();
                        last.flags |= VB_SYM_LASTGROUP;
                        last.out_off = (uint32_t)ypart_len;
                        VbSymGroupOut go;
                        go.off = last.out_off;
                        if (below) {
                            last.flags |= VB_SYM_BELOW;
                            last.out_len = (uint16_t)w;
                            last.nrows_g = (uint16_t)(std::min<int64_t>(r0 + VB_SYM_R, nr) - grow0);
                        } else {
                            last.out_len = (uint16_t)std::min<int64_t>(r0 + VB_SYM_R, w);
                        }
                        go.len = last.out_len;
                        if (ypart_len + go.len + last.nrows_g > 0xffffffffull) return vb_fail("LD partial buffer too large");
                        ypart_len += go.len + last.nrows_g;
                        gout.push_back(go);
                        sgroups.push_back(cur);
                        sgroup_bytes.push_back(group_bytes);
                        cur.first_item = (uint32_t)sitems.size();
                        cur.n_items = 0;
                        group_bytes = 0;
                        grow0 = r0 + VB_SYM_R;
                    }
                }
                sl.ng = (uint32_t)sgroups.size() - sl.g0;
                if (!gb0_set) sl.gb0 = sl.g0 + sl.ng;
                cursor += vb_sym_slab_doubles(nr, w);
                b.sslabs.push_back(sl);
            }
            bref[bi].ng = (uint32_t)sgroups.size() - bref[bi].g0;
            L.bytes += 4 * n * (n + 1);
        } else if (b.r < 0) {
            add_slabs(L, b.slabs2, b.n, b.n, b.xpos, b.xpos, L.xb_len, cursor, dummy2);
            L.bytes += 8 * b.n * b.n;
        } else if (b.fac1) {
            // read-once factor: U' = U sqrt(s) in chunks of c columns (c x n_pad column-major, contiguous),
            // groups of consecutive chunks (~0.5 MB), one partial y vector (n_pad) per group
            b.n_pad = even_up(b.n);
            b.fac_c = vb_fac_chunk_cols(b.n_pad);
            b.fac_off = cursor;
            b.fg0 = (uint32_t)fgroups.size();
            VbSymGroup cur;
            cur.first_item = (uint32_t)fitems.size();
            cur.n_items = 0;
            size_t group_bytes = 0;
            for (int64_t j0 = 0; j0 < b.r; j0 += b.fac_c) {
                const int64_t cq = std::min<int64_t>(b.fac_c, b.r - j0);
                const size_t a_off = b.fac_off + (size_t)j0 * b.n_pad;
                if ((a_off >> 1) > 0xffffffffull) return vb_fail("LD store of one cohort exceeds 64 GiB on this rank");
                VbFacItem it;
                it.a_off16 = (uint32_t)(a_off >> 1);
                it.x_off2 = (uint32_t)(b.xpos >> 1);
                it.n2 = (uint16_t)(b.n_pad >> 1);
                it.c = (uint16_t)cq;
                it.flags = VB_FAC_VALID | (j0 == 0 ? VB_FAC_FIRST : 0);
                it.pad = 0;
                it.out_off = 0;
                fitems.push_back(it);
                cur.n_items++;
                group_bytes += (size_t)cq * b.n_pad * 8;
                if (j0 + cq >= b.r || group_bytes >= group_bytes_base) {
                    ypart_len = (size_t)even_up((int64_t)ypart_len);          // double2 stores
                    VbFacItem& last = fitems.back();
                    last.flags |= VB_FAC_LASTGROUP;
                    last.out_off = (uint32_t)ypart_len;
                    VbSymGroupOut go;
                    go.off = last.out_off;
                    go.len = (uint32_t)b.n_pad;
                    if (ypart_len + b.n_pad > 0xffffffffull) return vb_fail("LD partial buffer too large");
                    ypart_len += b.n_pad;
                    fgout.push_back(go);
                    fgroups.push_back(cur);
                    cur.first_item = (uint32_t)fitems.size();
                    cur.n_items = 0;
                    group_bytes = 0;
                }
            }
            b.fng = (uint32_t)fgroups.size() - b.fg0;
            cursor += (size_t)b.r * b.n_pad;
            L.bytes += 8 * b.n_pad * b.r;
        } else {
            // phase 1: t = V' x, V' = diag(s) U^T  (r x n), x from xb, out to tb
            add_slabs(L, b.slabs1, b.r, b.n, b.xpos, b.tpos, L.tb_len, cursor, dummy1);
            // phase 2: y = U t  (n x r), x from tb (stored behind xb in xall), out to yb
            add_slabs(L, b.slabs2, b.n, b.r, L.xb_len + b.tpos, b.xpos, L.xb_len, cursor, dummy2);
            L.bytes += 16 * b.n * b.r;
        }
    }
    L.all_sym = !L.blocks.empty();
    for (auto& b : L.blocks) L.all_sym = L.all_sym && b.sym;
    L.block_ng.resize(L.blocks.size());
    for (size_t bi = 0; bi < L.blocks.size(); ++bi) L.block_ng[bi] = bref[bi].ng;
    L.mat_len = std::max<size_t>(cursor, 2);
    CK(cudaMalloc(&L.mat, L.mat_len * sizeof(double)));
    CK(cudaMemsetAsync(L.mat, 0, L.mat_len * sizeof(double), ctx->stream));
    // + slack: the symmetric kernel reads 8 x values per panel even for a ragged last panel
    CK(cudaMalloc(&L.xall, (size_t)(L.xb_len + L.tb_len + 16) * sizeof(double)));
    CK(cudaMemsetAsync(L.xall, 0, (size_t)(L.xb_len + L.tb_len + 16) * sizeof(double), ctx->stream));
    L.n_sgroups = (int64_t)sgroups.size();
    L.n_fgroups = (int64_t)fgroups.size();
    L.gout_fbase = (uint32_t)gout.size();
    gout.insert(gout.end(), fgout.begin(), fgout.end());
    if (L.n_fgroups > 0) {
        CK(cudaMalloc(&L.fitems, fitems.size() * sizeof(VbFacItem)));
        CK(cudaMemcpy(L.fitems, fitems.data(), fitems.size() * sizeof(VbFacItem), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&L.fgroups, fgroups.size() * sizeof(VbSymGroup)));
        CK(cudaMemcpy(L.fgroups, fgroups.data(), fgroups.size() * sizeof(VbSymGroup), cudaMemcpyHostToDevice));
    }
    if (L.n_sgroups > 0) {
        CK(cudaMalloc(&L.sitems, sitems.size() * sizeof(VbSymItem)));
        CK(cudaMemcpy(L.sitems, sitems.data(), sitems.size() * sizeof(VbSymItem), cudaMemcpyHostToDevice));
        // claim order: largest groups first (LPT), so the tail of the dynamic schedule is made of
        // the smallest groups; group *indices* (gout / bref) keep the block order
        std::vector<uint32_t> order(sgroups.size());
        for (size_t g = 0; g < order.size(); ++g) order[g] = (uint32_t)g;
        if (VB_SYM_LPT)
            std::stable_sort(order.begin(), order.end(),
                             [&](uint32_t x, uint32_t y) { return sgroup_bytes[x] > sgroup_bytes[y]; });
        std::vector<VbSymGroup> sched(sgroups.size());
        for (size_t g = 0; g < order.size(); ++g) sched[g] = sgroups[order[g]];
        CK(cudaMalloc(&L.sgroups, sched.size() * sizeof(VbSymGroup)));
        CK(cudaMemcpy(L.sgroups, sched.data(), sched.size() * sizeof(VbSymGroup), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&L.bref, bref.size() * sizeof(VbSymBlockRef)));
        CK(cudaMemcpy(L.bref, bref.data(), bref.size() * sizeof(VbSymBlockRef), cudaMemcpyHostToDevice));
    }
    if (L.n_sgroups + L.n_fgroups > 0) {
        L.gout_host = gout;
        CK(cudaMalloc(&L.gout, gout.size() * sizeof(VbSymGroupOut)));
        CK(cudaMemcpy(L.gout, gout.data(), gout.size() * sizeof(VbSymGroupOut), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&L.ypart, (std::max<size_t>(ypart_len, 1) + 2) * sizeof(double)));
        CK(cudaMemsetAsync(L.ypart, 0, (std::max<size_t>(ypart_len, 1) + 2) * sizeof(double), ctx->stream));
    }
    CK(cudaMalloc(&L.yb, (size_t)L.nslab2 * L.xb_len * sizeof(double)));
    CK(cudaMemsetAsync(L.yb, 0, (size_t)L.nslab2 * L.xb_len * sizeof(double), ctx->stream));
    if (L.nslab1 > 1) {
        CK(cudaMalloc(&L.tbs, (size_t)L.nslab1 * L.tb_len * sizeof(double)));
        CK(cudaMemsetAsync(L.tbs, 0, (size_t)L.nslab1 * L.tb_len * sizeof(double), ctx->stream));
    }
    if (build_items(ctx, L, 1, &L.items1, &L.n_items1)) return 1;
    if (build_items(ctx, L, 2, &L.items2, &L.n_items2)) return 1;
    CK(cudaMalloc(&L.sched, 8 * sizeof(uint32_t)));
    CK(cudaMemsetAsync(L.sched, 0, 8 * sizeof(uint32_t), ctx->stream));
    return 0;
}

extern "C" int vb_ld_set_dense(vb_ld* h, int64_t b, const double* R, int64_t ld, int on_device) {
    if (!h) return vb_fail("vb_ld_set_dense: null handle");
    vb_ctx* ctx = h->ctx;
    LdPop& L = h->L;
    if (b < 0 || b >= (int64_t)L.blocks.size()) return vb_fail("vb_ld_set_dense: bad block");
    LdBlock& B = L.blocks[b];
    if (B.r >= 0) return vb_fail("vb_ld_set_dense: block %lld was declared as a factor", (long long)b);
    CK(cudaSetDevice(ctx->device));
    if (B.sym) {
        const double* dR = R;
        double* tmp = nullptr;
        int64_t dld = ld;
        if (!on_device) {
            CK(cudaMalloc(&tmp, (size_t)B.n * B.n * sizeof(double)));
            CK(cudaMemcpy2DAsync(tmp, B.n * sizeof(double), R, ld * sizeof(double), B.n * sizeof(double),
                                 B.n, cudaMemcpyHostToDevice, ctx->stream));
            dR = tmp;
            dld = B.n;
        }
        for (const SymSlab& sl : B.sslabs) {
            const int64_t nr = B.n - sl.J0;
            const int npan = (int)((nr + VB_SYM_R - 1) / VB_SYM_R);
            vb_pack_sym_kernel<<<npan, 256, 0, ctx->stream>>>(dR + (size_t)sl.J0 * dld + sl.J0, dld, (int)nr,
                                                              (int)sl.w, L.mat + sl.off);
            CK_LAUNCH(ctx);
        }
        if (!on_device) {
            CK(cudaStreamSynchronize(ctx->stream));
            cudaFree(tmp);
        }
        B.filled = true;
        return 0;
    }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int64_t c0 = 0;
    for (int si : B.slabs2) {
        const LdSlab& s = L.slabs[si];
        CK(cudaMemcpy2DAsync(L.mat + s.off, s.ld * sizeof(double), R + c0, ld * sizeof(double),
                             s.cols * sizeof(double), s.rows, kind, ctx->stream));
        c0 += s.cols;
    }
    if (!on_device) CK(cudaStreamSynchronize(ctx->stream));   // caller may reuse its buffer
    B.filled = true;
    return 0;
}

// out[c][j] = s[c] * U[j][c]   (V' = diag(s) U^T), written into a slab with leading dim ld
__global__ void vb_pack_vprime_kernel(const double* __restrict__ U, const double* __restrict__ s,
                                      int64_t n, int64_t r, int64_t c0, int64_t cols, int64_t ld,
                                      double* __restrict__ out) {
    // out is r x ld; column j of out = SNP (c0 + j)
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < r * cols;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = idx / cols, j = idx % cols;
        out[k * ld + j] = s[k] * U[(c0 + j) * r + k];
    }
}

extern "C" int vb_ld_set_factor(vb_ld* h, int64_t b, const double* U, const double* s,
                                int on_device) {
    if (!h) return vb_fail("vb_ld_set_factor: null handle");
    vb_ctx* ctx = h->ctx;
    LdPop& L = h->L;
    if (b < 0 || b >= (int64_t)L.blocks.size()) return vb_fail("vb_ld_set_factor: bad block");
    LdBlock& B = L.blocks[b];
    if (B.r < 0) return vb_fail("vb_ld_set_factor: block %lld was declared dense", (long long)b);
    CK(cudaSetDevice(ctx->device));
    const int64_t n = B.n, r = B.r;
    const double *dU = U, *ds = s;
    double *tmpU = nullptr, *tmps = nullptr;
    if (!on_device) {
        CK(cudaMalloc(&tmpU, (size_t)n * r * sizeof(double)));
        CK(cudaMalloc(&tmps, (size_t)r * sizeof(double)));
        CK(cudaMemcpyAsync(tmpU, U, (size_t)n * r * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(tmps, s, (size_t)r * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dU = tmpU;
        ds = tmps;
    }
    if (B.fac1) {
        std::vector<double> hs((size_t)r);
        if (on_device) CK(cudaMemcpy(hs.data(), s, (size_t)r * sizeof(double), cudaMemcpyDeviceToHost));
        else std::memcpy(hs.data(), s, (size_t)r * sizeof(double));
        for (int64_t k = 0; k < r; ++k)
            if (!(hs[k] >= 0.0)) {
                if (!on_device) { cudaStreamSynchronize(ctx->stream); cudaFree(tmpU); cudaFree(tmps); }
                return vb_fail("vb_ld_set_factor: block %lld has a negative (or NaN) factor weight; the read-once "
                               "form needs s >= 0 -- declare the block dense or set option ld_factor_once=0",
                               (long long)b);
            }
        const int nchunks = (int)((r + B.fac_c - 1) / B.fac_c);
        vb_pack_fac_kernel<<<nchunks, 256, 0, ctx->stream>>>(dU, ds, (int)n, (int)r, (int)B.n_pad, B.fac_c,
                                                             L.mat + B.fac_off);
        CK_LAUNCH(ctx);
    }
    int64_t c0 = 0;
    for (int si : B.slabs1) {
        const LdSlab& sl = L.slabs[si];
        const int64_t tot = r * sl.cols;
        const int grid = (int)std::min<int64_t>((tot + 255) / 256, 4096);
        vb_pack_vprime_kernel<<<grid, 256, 0, ctx->stream>>>(dU, ds, n, r, c0, sl.cols, sl.ld,
                                                              L.mat + sl.off);
        CK_LAUNCH(ctx);
        c0 += sl.cols;
    }
    c0 = 0;
    for (int si : B.slabs2) {
        const LdSlab& sl = L.slabs[si];
        CK(cudaMemcpy2DAsync(L.mat + sl.off, sl.ld * sizeof(double), dU + c0, r * sizeof(double),
                             sl.cols * sizeof(double), sl.rows, cudaMemcpyDeviceToDevice, ctx->stream));
        c0 += sl.cols;
    }
    if (!on_device) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(tmpU);
        cudaFree(tmps);
    }
    B.filled = true;
    return 0;
}

extern "C" int vb_ld_finalize(vb_ld* h, const int64_t* perm_host, int64_t nperm) {
    if (!h) return vb_fail("vb_ld_finalize: null handle");
    vb_ctx* ctx = h->ctx;
    LdPop& L = h->L;
    CK(cudaSetDevice(ctx->device));
    int64_t tot = 0;
    for (auto& b : L.blocks) {
        if (!b.filled) return vb_fail("vb_ld_finalize: a block was never set");
        tot += b.n;
    }
    if (tot != nperm) return vb_fail("vb_ld_finalize: perm has %lld entries, blocks hold %lld SNPs",
                                     (long long)nperm, (long long)tot);
    std::vector<int32_t> pos(std::max<int64_t>(nperm, 1)), snp(std::max<int64_t>(nperm, 1));
    std::vector<char> seen(L.M, 0);
    std::vector<VbFinRec> rec(std::max<int64_t>(nperm, 1));
    std::vector<uint32_t> xstart, xoffs;            // wide blocks only
    bool any_wide = false;
    for (auto& b : L.blocks) any_wide = any_wide || b.sslabs.size() > 1;
    if (any_wide) xstart.assign(std::max<int64_t>(nperm, 1), 0);
    int64_t j = 0;
    for (size_t bi = 0; bi < L.blocks.size(); ++bi) {
        LdBlock& b = L.blocks[bi];
        for (int64_t t = 0; t < b.n; ++t, ++j) {
            rec[j].gfirst = -1;
            rec[j].loc_ncover = 0;
            if (b.sym) {
                // the slab whose columns contain t, and the first of its groups whose partial vector
                // covers column l = t - J0 (lengths increase through the tile, then stay at w)
                size_t si = 0;
                while (si + 1 < b.sslabs.size() && t >= b.sslabs[si].J0 + b.sslabs[si].w) ++si;
                const SymSlab& sl = b.sslabs[si];
                const uint32_t l = (uint32_t)(t - sl.J0);
                uint32_t g = sl.g0;
                while (g + 1 < sl.g0 + sl.ng && L.gout_host[g].len <= l) ++g;
                const uint32_t ncover = sl.g0 + sl.ng - g;
                if (l > 0xfff || ncover > 0xfff || si > 0xff) return vb_fail("internal: finish record overflow");
                rec[j].gfirst = (int32_t)g;
                rec[j].loc_ncover = l | (ncover << 12) | ((uint32_t)si << 24);
                if (si > 0) {
                    // its row sums from the slabs to the left: one entry per slab, in slab order
                    if (xoffs.size() + si > 0xffffffffull) return vb_fail("LD finish table too large");
                    xstart[j] = (uint32_t)xoffs.size();
                    for (size_t s2 = 0; s2 < si; ++s2) {
                        const SymSlab& le = b.sslabs[s2];
                        const int64_t below = t - (le.J0 + le.w);           // row index below that slab's tile
                        const uint32_t gb = le.gb0 + (uint32_t)(below / le.Rg);
                        xoffs.push_back(L.gout_host[gb].off + L.gout_host[gb].len + (uint32_t)(below % le.Rg));
                    }
                }
            } else if (b.fac1) {
                // every group of the block emits a full-length partial vector
                if (t > 0xfff || b.fng > 0xfff) return vb_fail("internal: finish record overflow");
                rec[j].gfirst = (int32_t)(L.gout_fbase + b.fg0);
                rec[j].loc_ncover = (uint32_t)t | (b.fng << 12);
            }
            const int64_t i = perm_host[j];
            if (i < 0 || i >= L.M) return vb_fail("vb_ld_finalize: perm[%lld]=%lld out of range", (long long)j, (long long)i);
            if (seen[i]) return vb_fail("vb_ld_finalize: SNP %lld appears twice in perm", (long long)i);
            seen[i] = 1;
            pos[j] = (int32_t)(b.xpos + t);
            snp[j] = (int32_t)i;
            rec[j].pos = pos[j];
            rec[j].snp = snp[j];
        }
    }
    L.nreal = nperm;
    {
        std::vector<int32_t> xbpos(L.M, -1);
        for (int64_t t = 0; t < nperm; ++t) xbpos[snp[t]] = pos[t];
        CK(cudaMalloc(&L.xbpos, (size_t)L.M * sizeof(int32_t)));
        CK(cudaMemcpy(L.xbpos, xbpos.data(), (size_t)L.M * sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&L.fin_counter, sizeof(uint32_t)));
        CK(cudaMemset(L.fin_counter, 0, sizeof(uint32_t)));
    }
    if (L.n_sgroups + L.n_fgroups > 0) {
        CK(cudaMalloc(&L.finrec, rec.size() * sizeof(VbFinRec)));
        CK(cudaMemcpy(L.finrec, rec.data(), rec.size() * sizeof(VbFinRec), cudaMemcpyHostToDevice));
        if (L.all_sym) {
            std::vector<VbSymBlockFin> bf(L.blocks.size());
            uint32_t p0 = 0;
            for (size_t bi = 0; bi < L.blocks.size(); ++bi) {
                bf[bi].pos0 = p0;
                bf[bi].n = (uint32_t)L.blocks[bi].n;
                bf[bi].ng = L.block_ng[bi];
                bf[bi].pad = 0;
                p0 += (uint32_t)L.blocks[bi].n;
            }
            CK(cudaMalloc(&L.bfin, bf.size() * sizeof(VbSymBlockFin)));
            CK(cudaMemcpy(L.bfin, bf.data(), bf.size() * sizeof(VbSymBlockFin), cudaMemcpyHostToDevice));
            CK(cudaMalloc(&L.block_cnt, bf.size() * sizeof(uint32_t)));
            CK(cudaMemset(L.block_cnt, 0, bf.size() * sizeof(uint32_t)));
            CK(cudaMalloc(&L.done_cnt, 2 * sizeof(uint32_t)));
            CK(cudaMemset(L.done_cnt, 0, 2 * sizeof(uint32_t)));
            CK(cudaMalloc(&L.part_blk, bf.size() * sizeof(double)));
        }
        if (any_wide) {
            if (xoffs.empty()) xoffs.push_back(0);
            CK(cudaMalloc(&L.xstart, xstart.size() * sizeof(uint32_t)));
            CK(cudaMemcpy(L.xstart, xstart.data(), xstart.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            CK(cudaMalloc(&L.xoffs, xoffs.size() * sizeof(uint32_t)));
            CK(cudaMemcpy(L.xoffs, xoffs.data(), xoffs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    CK(cudaMalloc(&L.pos, pos.size() * sizeof(int32_t)));
    CK(cudaMalloc(&L.snp, snp.size() * sizeof(int32_t)));
    CK(cudaMemcpy(L.pos, pos.data(), pos.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.snp, snp.data(), snp.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    L.finalized = true;
    return 0;
}

extern "C" int64_t vb_ld_bytes(const vb_ld* h) { return h ? h->L.bytes : -1; }

// x (SNP order, device) -> linked (SNP order, device), optional partial sums of x.y
static int ld_apply(vb_ctx* ctx, LdPop& L, const double* x_snp, double* y_snp, double* partial,
                    int grid_fin, const VbFinalArgs* final_args = nullptr) {
    if (!L.finalized) return vb_fail("LD operator used before vb_ld_finalize");
    cudaStream_t st = ctx->stream;
    VbFinalArgs fa;
    std::memset(&fa, 0, sizeof(fa));
    if (final_args) fa = *final_args;
    fa.counter = L.fin_counter;
    if (!fa.do_final) fa.xr.enabled = 0;
    if (L.nreal > 0 && x_snp) {
        const int g = (int)std::min<int64_t>((L.nreal + 255) / 256, 1184);
        vb_ld_gather_kernel<<<g, 256, 0, st>>>(x_snp, L.pos, L.snp, L.nreal, L.xall);
        CK_LAUNCH(ctx);
    }
    if (L.n_items1 > 0) {
        double* out1 = L.nslab1 > 1 ? L.tbs : L.xall + L.xb_len;
        prof_begin(ctx, 0);
        vb_ld_matvec_kernel<<<ctx->num_sms, VB_LD_THREADS, VB_LD_SMEM, st>>>(
            L.mat, L.items1, (uint32_t)L.n_items1, L.sched, L.xall, out1);
        prof_end(ctx, 0);
        CK_LAUNCH(ctx);
        if (L.nslab1 > 1) {
            const int g = (int)std::min<int64_t>((L.tb_len + 255) / 256, 1184);
            vb_ld_slab_sum_kernel<<<g, 256, 0, st>>>(L.tbs, L.tb_len, L.nslab1, L.xall + L.xb_len);
            CK_LAUNCH(ctx);
        }
    }
    if (L.n_items2 > 0) {
        prof_begin(ctx, 0);
        vb_ld_matvec_kernel<<<ctx->num_sms, VB_LD_THREADS, VB_LD_SMEM, st>>>(
            L.mat, L.items2, (uint32_t)L.n_items2, L.sched + 2, L.xall, L.yb);
        prof_end(ctx, 0);
        CK_LAUNCH(ctx);
    }
    VbFuseFin ff;
    std::memset(&ff, 0, sizeof(ff));
    // the finish fused into the mat-vec needs its partial-sum slot: callers without one (vb_ld_dot) and
    // operators that mix block forms keep the separate finish kernel
    // (a block's finish stalls the CTA that flushed its last group for a few microseconds: measured on C2,
    // 1700 blocks over 296 CTAs, 0.806 ms fused vs 0.706 + 0.045 ms separate; and on 1/8 and 1/4 shards --
    // 212 / 425 blocks -- 0.164 vs 0.107 + 0.021 ms and 0.260 vs 0.194 + 0.024 ms: the separate kernel's
    // massively parallel gather wins at every size, so automatic = off; the fused path stays selectable)
    const bool want_fused = g_fused_finish > 0;
    const bool fused = want_fused && L.all_sym && L.n_sgroups > 0 && partial != nullptr;
    if (fused) {
        ff.enabled = 1;
        ff.nblocks = (uint32_t)L.blocks.size();
        ff.rec = L.finrec; ff.gout = L.gout; ff.xstart = L.xstart; ff.xoffs = L.xoffs;
        ff.bfin = L.bfin; ff.block_cnt = L.block_cnt; ff.done_cnt = L.done_cnt; ff.part_blk = L.part_blk;
        ff.y_snp = y_snp;
        ff.part_fin = partial;
        ff.fa = fa;
    }
    if (L.n_fgroups > 0) {
        prof_begin(ctx, 0);
        vb_ld_fac_kernel<<<ctx->num_sms * 2, VB_LD_THREADS, VB_FAC_SMEM, st>>>(
            L.mat, L.fitems, L.fgroups, (uint32_t)L.n_fgroups, L.sched + 6, L.xall, L.ypart);
        prof_end(ctx, 0);
        CK_LAUNCH(ctx);
    }
    if (L.n_sgroups > 0) {
        prof_begin(ctx, 0);
        vb_ld_sym_kernel<<<ctx->num_sms * VB_SYM_CTAS_PER_SM, VB_LD_THREADS, VB_SYM_SMEM, st>>>(
            L.mat, L.sitems, L.sgroups, (uint32_t)L.n_sgroups, L.sched + 4, L.xall, L.ypart, ff);
        prof_end(ctx, 0);
        CK_LAUNCH(ctx);
        if (fused) return 0;
    }
    if (L.n_sgroups + L.n_fgroups > 0) {
        prof_begin(ctx, 2);
        vb_ld_finish_sym_kernel<<<grid_fin, 256, 0, st>>>(L.yb, L.xb_len, L.nslab2, L.ypart, L.finrec,
                                                          L.gout, L.xstart, L.xoffs, L.xall, L.nreal, y_snp,
                                                          partial, fa);
        prof_end(ctx, 2);
        CK_LAUNCH(ctx);
    } else {
        prof_begin(ctx, 2);
        vb_ld_finish_kernel<<<grid_fin, 256, 0, st>>>(L.yb, L.xb_len, L.nslab2, L.xall, L.pos, L.snp,
                                                      L.nreal, y_snp, partial, fa);
        prof_end(ctx, 2);
        CK_LAUNCH(ctx);
    }
    return 0;
}

extern "C" int vb_ld_dot(vb_ld* h, const double* x_dev, double* y_dev) {
    if (!h) return vb_fail("vb_ld_dot: null handle");
    vb_ctx* ctx = h->ctx;
    CK(cudaSetDevice(ctx->device));
    LdPop& L = h->L;
    CK(cudaMemsetAsync(y_dev, 0, (size_t)L.M * sizeof(double), ctx->stream));
    return ld_apply(ctx, L, x_dev, y_dev, nullptr, 296);
}

// ------------------------------------------------------------------------------------
// set-up on the device (dense, numerically full-rank blocks)
// ------------------------------------------------------------------------------------
extern "C" int64_t vb_setup_nmax(void) { return 4096; }
extern "C" int vb_setup_dense(vb_ctx* ctx, int64_t nblocks, const int64_t* n_host, const double* R_dev,
                              double* W_dev, const double* z_dev, const double* reg_dev, double* mle_dev,
                              double* rmle_dev, double* ridge_dev, double* chi_dev, double* lam_dev,
                              int32_t* status_dev) {
    if (!ctx || !n_host || nblocks < 1) return vb_fail("vb_setup_dense: bad argument");
    CK(cudaSetDevice(ctx->device));
    std::vector<VbSetupBlock> blk(nblocks);
    std::vector<int64_t> order(nblocks);
    int64_t moff = 0, voff = 0, nmax = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        if (n_host[b] < 1 || n_host[b] > vb_setup_nmax())
            return vb_fail("vb_setup_dense: block %lld has n=%lld (1..%lld supported)", (long long)b,
                           (long long)n_host[b], (long long)vb_setup_nmax());
        blk[b].mat_off = moff; blk[b].vec_off = voff; blk[b].n = (int32_t)n_host[b]; blk[b].index = (int32_t)b;
        moff += n_host[b] * n_host[b];
        voff += n_host[b];
        nmax = std::max(nmax, n_host[b]);
        order[b] = b;
    }
    // largest blocks first: the tail of the dynamic schedule is made of the cheapest ones
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t c) { return n_host[a] > n_host[c]; });
    std::vector<VbSetupBlock> sched(nblocks);
    for (int64_t b = 0; b < nblocks; ++b) sched[b] = blk[order[b]];
    VbSetupBlock* d_blk = nullptr;
    uint32_t* d_cnt = nullptr;
    CK(cudaMalloc(&d_blk, nblocks * sizeof(VbSetupBlock)));
    CK(cudaMalloc(&d_cnt, sizeof(uint32_t)));
    CK(cudaMemcpyAsync(d_blk, sched.data(), nblocks * sizeof(VbSetupBlock), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t), ctx->stream));
    const size_t smem = (2 * (size_t)nmax + 32 * 33 + 2 * 64 * 33) * sizeof(double);
    CK(cudaFuncSetAttribute(vb_setup_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(200 * 1024) / (smem + 2048)));
    const int grid = (int)std::min<int64_t>(nblocks, (int64_t)ctx->num_sms * per_sm);
    vb_setup_dense_kernel<<<grid, VB_SETUP_THREADS, smem, ctx->stream>>>(
        d_blk, (int)nblocks, d_cnt, R_dev, W_dev, z_dev, reg_dev, mle_dev, rmle_dev, ridge_dev, chi_dev,
        lam_dev, status_dev, (int)nmax);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_blk);
    cudaFree(d_cnt);
    if (e != cudaSuccess) return vb_fail("vb_setup_dense: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------
// fit state
// ------------------------------------------------------------------------------------
extern "C" int vb_fit_create(vb_ctx* ctx, int K, int P, int64_t M, int A, vb_ld* const* lds) {
    if (!ctx || !lds) return vb_fail("vb_fit_create: null argument");
    if (P < 1 || P > VB_MAX_POPS)
        return vb_fail("vb_fit_create: %d cohorts requested; kernels are compiled for 1..%d", P, VB_MAX_POPS);
    if (K < 1 || A < 1 || M < 1) return vb_fail("vb_fit_create: bad sizes K=%d A=%d M=%lld", K, A, (long long)M);
    CK(cudaSetDevice(ctx->device));
    Fit& f = ctx->fit;
    if (f.created) free_fit(f);
    f.created = true;
    f.K = K; f.P = P; f.A = A; f.M = M;
    const size_t PM = (size_t)P * M, KM = (size_t)K * M;
    CK(cudaMalloc(&f.adj, PM * 8)); CK(cudaMalloc(&f.se, PM * 8)); CK(cudaMalloc(&f.sld, PM * 8));
    CK(cudaMalloc(&f.scal, PM * 8)); CK(cudaMalloc(&f.ann, (size_t)M * 4));
    CK(cudaMalloc(&f.prec, (size_t)K * P * P * 8)); CK(cudaMalloc(&f.logdet, (size_t)K * 8));
    // log hyper_delta and the delta-gradient table side by side: the hyper step uploads both at once
    CK(cudaMalloc(&f.logh, (size_t)2 * A * K * 8)); CK(cudaMalloc(&f.inv_tau_dev, VB_MAXP * 8));
    f.gfull = f.logh + (size_t)A * K;
    CK(cudaMemsetAsync(f.logh, 0, (size_t)2 * A * K * 8, ctx->stream));
    for (int s = 0; s < 2; ++s) {
        CK(cudaMalloc(&f.mu[s], KM * P * 8));
        CK(cudaMalloc(&f.delta[s], KM * 8));
        CK(cudaMalloc(&f.pm[s], PM * 8));
        CK(cudaMalloc(&f.linked[s], PM * 8));
        CK(cudaMemsetAsync(f.linked[s], 0, PM * 8, ctx->stream));
        CK(cudaMemsetAsync(f.pm[s], 0, PM * 8, ctx->stream));
    }
    CK(cudaMalloc(&f.scratch3, 3 * PM * 8));
    CK(cudaMalloc(&f.pm_prev, PM * 8)); CK(cudaMalloc(&f.pm_ckpt, PM * 8));
    CK(cudaMalloc(&f.pm_next, PM * 8));
    CK(cudaMemsetAsync(f.pm_next, 0, PM * 8, ctx->stream));
    f.akf = 0;                                        // see vb_fit_set_fusion
    f.nsp = VB_NSNPSTAT(P) + VB_FUSE_ANN_MAX;
    CK(cudaMemsetAsync(f.pm_prev, 0, PM * 8, ctx->stream));
    CK(cudaMemsetAsync(f.pm_ckpt, 0, PM * 8, ctx->stream));
    f.grid_snp = (int)std::min<int64_t>((M + 127) / 128, (int64_t)ctx->num_sms * 16);
    f.grid_fin = 8 * ctx->num_sms;
    f.grid_ann = (int)std::min<int64_t>((M + 255) / 256, (int64_t)ctx->num_sms);
    f.grid_diff = (int)std::min<int64_t>((int64_t)(PM + 255) / 256, (int64_t)ctx->num_sms * 4);
    CK(cudaMalloc(&f.part_snp, (size_t)f.grid_snp * f.nsp * 8));
    CK(cudaMalloc(&f.part_fin, (size_t)P * f.grid_fin * 8));
    CK(cudaMemsetAsync(f.part_fin, 0, (size_t)P * f.grid_fin * 8, ctx->stream));
    CK(cudaMalloc(&f.part_ann, (size_t)f.grid_ann * K * A * 8));
    CK(cudaMalloc(&f.part_diff, (size_t)f.grid_diff * 10 * 8));
    for (int p = 0; p < VB_MAXP; ++p) f.inv_tau[p] = 1.0;
    CK(cudaMemcpyAsync(f.inv_tau_dev, f.inv_tau, VB_MAXP * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    f.cur_mu = f.cur_delta = f.cur_vec = 0;
    f.trial_kind = -1;
    ctx->fit_ld.assign(lds, lds + P);
    for (int p = 0; p < P; ++p) {
        if (!lds[p] || lds[p]->ctx != ctx) return vb_fail("vb_fit_create: LD operator %d belongs to another context", p);
        if (!lds[p]->L.finalized) return vb_fail("vb_fit_create: LD operator %d is not finalized", p);
        if (lds[p]->L.M != M) return vb_fail("vb_fit_create: LD operator %d has M=%lld, fit has M=%lld", p, (long long)lds[p]->L.M, (long long)M);
    }
    return 0;
}
// fuse_ann != 0: every evaluation also accumulates the per-annotation sums of that state's delta
// (needs A*K <= 48 and A <= 4, silently off otherwise).  Worth it when a separate pass + reduction
// per hyper step costs more than ~5 % extra per-SNP kernel time, i.e. on multi-GPU runs.
struct TilePlan { int W, grid; size_t smem; };
static TilePlan tile_plan(const vb_ctx* ctx, const Fit& f, int akf);
// fuse_ann == 2: only where the sums are nearly free -- the three-pass kernel's per-thread
// shared-memory slots (P <= 2, A*K <= 16), e.g. the single-cohort default grid on one GPU.
extern "C" int vb_fit_set_fusion(vb_ctx* ctx, int fuse_ann) {
    if (!ctx || !ctx->fit.created) return vb_fail("fit state not created");
    Fit& f = ctx->fit;
    bool on = fuse_ann && f.A * f.K <= VB_FUSE_ANN_MAX && f.A <= 4;
    if (fuse_ann == 2)
        on = on && f.P <= 2 && f.A * f.K <= VB_FUSE_ANN_SLOTS && g_three_pass && tile_plan(ctx, f, 0).W == 0;
    f.akf = on ? f.A * f.K : 0;
    return 0;
}
extern "C" int vb_fit_destroy(vb_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_fit(ctx->fit);
    ctx->fit_ld.clear();
    return 0;
}

#define NEED_FIT(ctx)                                                      \
    if (!(ctx) || !(ctx)->fit.created) return vb_fail("fit state not created"); \
    CK(cudaSetDevice((ctx)->device));                                      \
    Fit& f = (ctx)->fit;                                                   \
    f.mutations++;

extern "C" int vb_fit_set_snp_data(vb_ctx* ctx, const double* adj, const double* se, const double* sld,
                                   const double* scal, const int32_t* ann) {
    NEED_FIT(ctx);
    const size_t PM = (size_t)f.P * f.M;
    CK(cudaMemcpyAsync(f.adj, adj, PM * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.se, se, PM * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.sld, sld, PM * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.scal, scal, PM * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.ann, ann, (size_t)f.M * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int vb_fit_set_mixture(vb_ctx* ctx, const double* prec, const double* logdet) {
    NEED_FIT(ctx);
    CK(cudaMemcpyAsync(f.prec, prec, (size_t)f.K * f.P * f.P * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f.logdet, logdet, (size_t)f.K * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int vb_fit_set_hyper(vb_ctx* ctx, const double* hyper) {
    NEED_FIT(ctx);
    std::vector<double> lh((size_t)f.A * f.K);
    for (size_t t = 0; t < lh.size(); ++t) lh[t] = std::log(hyper[t]);
    CK(cudaMemcpyAsync(f.logh, lh.data(), lh.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    // pageable source: cudaMemcpyAsync returns once the data is staged, so no stream sync is needed
    return 0;
}
extern "C" int vb_fit_set_delta_grad(vb_ctx* ctx, const double* g) {
    NEED_FIT(ctx);
    std::vector<double> full((size_t)f.A * f.K, 0.0);
    for (int a = 0; a < f.A; ++a)
        for (int k = 0; k + 1 < f.K; ++k) full[(size_t)a * f.K + k] = g[(size_t)a * (f.K - 1) + k];
    CK(cudaMemcpyAsync(f.gfull, full.data(), full.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    // pageable source: cudaMemcpyAsync returns once the data is staged, so no stream sync is needed
    return 0;
}
// hyper_delta [A][K] and its gradient table g [A][K-1] in ONE host->device copy (native loop)
static int fit_set_hyper_tables(vb_ctx* ctx, const double* hyper, const double* g) {
    NEED_FIT(ctx);
    const size_t AK = (size_t)f.A * f.K;
    std::vector<double> both(2 * AK, 0.0);
    for (size_t t = 0; t < AK; ++t) both[t] = std::log(hyper[t]);
    for (int a = 0; a < f.A; ++a)
        for (int k = 0; k + 1 < f.K; ++k) both[AK + (size_t)a * f.K + k] = g[(size_t)a * (f.K - 1) + k];
    CK(cudaMemcpyAsync(f.logh, both.data(), both.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    // pageable source: cudaMemcpyAsync returns once the data is staged, so no stream sync is needed
    return 0;
}
extern "C" int vb_fit_set_tau(vb_ctx* ctx, const double* tau) {
    NEED_FIT(ctx);
    for (int p = 0; p < f.P; ++p) f.inv_tau[p] = 1.0 / tau[p];
    CK(cudaMemcpyAsync(f.inv_tau_dev, f.inv_tau, VB_MAXP * 8, cudaMemcpyHostToDevice, ctx->stream));
    // pageable source: cudaMemcpyAsync returns once the data is staged, so no stream sync is needed
    return 0;
}
static int fit_set_params(vb_ctx* ctx, const double* mu, const double* delta_mk, cudaMemcpyKind kind) {
    NEED_FIT(ctx);
    const size_t KM = (size_t)f.K * f.M;
    if (mu) CK(cudaMemcpyAsync(f.mu[f.cur_mu], mu, KM * f.P * 8, kind, ctx->stream));     // (null: already queued)
    // vi_delta arrives in the reference layout [M][K]; stage it in the trial slot, transpose on device
    double* stage = f.delta[1 - f.cur_delta];
    CK(cudaMemcpyAsync(stage, delta_mk, KM * 8, kind, ctx->stream));
    const int grid = (int)std::min<int64_t>((f.M + 255) / 256, 4096);
    vb_mk_to_km_kernel<<<grid, 256, 0, ctx->stream>>>(stage, f.M, f.K, f.delta[f.cur_delta]);
    CK_LAUNCH(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    f.trial_kind = -1;
    return 0;
}
extern "C" int vb_fit_set_params(vb_ctx* ctx, const double* mu, const double* delta_mk) {
    return fit_set_params(ctx, mu, delta_mk, cudaMemcpyHostToDevice);
}
// same, from device buffers (this rank's shard already gathered on the device)
extern "C" int vb_fit_set_params_dev(vb_ctx* ctx, const double* mu_dev, const double* delta_mk_dev) {
    return fit_set_params(ctx, mu_dev, delta_mk_dev, cudaMemcpyDeviceToDevice);
}
static int fit_get_params(vb_ctx* ctx, double* mu, double* delta_mk, cudaMemcpyKind kind) {
    NEED_FIT(ctx);
    const size_t KM = (size_t)f.K * f.M;
    if (mu) CK(cudaMemcpyAsync(mu, f.mu[f.cur_mu], KM * f.P * 8, kind, ctx->stream));
    if (delta_mk) {
        double* stage = f.delta[1 - f.cur_delta];      // any pending trial is discarded
        const int grid = (int)std::min<int64_t>((f.M + 255) / 256, 4096);
        vb_km_to_mk_kernel<<<grid, 256, 0, ctx->stream>>>(f.delta[f.cur_delta], f.M, f.K, stage);
        CK_LAUNCH(ctx);
        CK(cudaMemcpyAsync(delta_mk, stage, KM * 8, kind, ctx->stream));
        f.trial_kind = -1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int vb_fit_get_params(vb_ctx* ctx, double* mu, double* delta_mk) {
    return fit_get_params(ctx, mu, delta_mk, cudaMemcpyDeviceToHost);
}
// same, into device buffers (for a device-side gather of the ranks' shards)
extern "C" int vb_fit_get_params_dev(vb_ctx* ctx, double* mu_dev, double* delta_mk_dev) {
    return fit_get_params(ctx, mu_dev, delta_mk_dev, cudaMemcpyDeviceToDevice);
}

// ---- sharded transfers: this rank's SNPs <-> GLOBAL host arrays.  Every rank moves only its own 1/N of
// the bytes.  Device -> host: scatter kernels write straight into host memory the device can address
// (page-locked or registered; posted PCIe writes, runs of consecutive SNPs keep them coalesced).
// rows: global[r*M_tot + idx[j]] = local[r*M_loc + j]   (vi_mu [K][P][M])
__global__ void vb_shard_rows_kernel(double* __restrict__ glob, const double* __restrict__ loc,
                                     const int64_t* __restrict__ idx, int64_t M_loc, int64_t M_tot) {
    const size_t r = blockIdx.y;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < M_loc; j += (int64_t)gridDim.x * blockDim.x) {
        glob[r * M_tot + idx[j]] = loc[r * M_loc + j];
    }
}
// vi_delta: global [M_tot][K] (reference layout) = local [K][M_loc]; k fastest so that a SNP's K values
// (and a run of SNPs) are one contiguous range on the host side
__global__ void vb_shard_delta_kernel(double* __restrict__ glob, const double* __restrict__ loc,
                                      const int64_t* __restrict__ idx, int64_t M_loc, int K) {
    const int64_t n = M_loc * K;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / K;
        const int k = (int)(t - j * K);
        glob[idx[j] * K + k] = loc[(size_t)k * M_loc + j];
    }
}
static double* device_view(const void* host);
// the reverse direction (upload from host memory the device can address): zero-copy reads over PCIe
__global__ void vb_shard_rows_gather_kernel(const double* __restrict__ glob, double* __restrict__ loc,
                                            const int64_t* __restrict__ idx, int64_t M_loc, int64_t M_tot) {
    const size_t r = blockIdx.y;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < M_loc; j += (int64_t)gridDim.x * blockDim.x)
        loc[r * M_loc + j] = glob[r * M_tot + idx[j]];
}
__global__ void vb_shard_delta_gather_kernel(const double* __restrict__ glob, double* __restrict__ loc,
                                             const int64_t* __restrict__ idx, int64_t M_loc, int K) {
    const int64_t n = M_loc * K;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / K;
        const int k = (int)(t - j * K);
        loc[(size_t)k * M_loc + j] = glob[idx[j] * K + k];
    }
}
extern "C" int vb_fit_set_shard(vb_ctx* ctx, const int64_t* snps_host, int64_t M_total) {
    NEED_FIT(ctx);
    for (int64_t j = 0; j < f.M; ++j)
        if (snps_host[j] < 0 || snps_host[j] >= M_total) return vb_fail("vb_fit_set_shard: index out of range");
    if (!f.shard_idx) CK(cudaMalloc(&f.shard_idx, (size_t)f.M * sizeof(int64_t)));
    CK(cudaMemcpy(f.shard_idx, snps_host, (size_t)f.M * sizeof(int64_t), cudaMemcpyHostToDevice));
    f.shard_total = M_total;
    // page-locked staging for uploads, allocated here so that no transfer pays for cudaHostAlloc (~4 ms)
    if (!f.shard_stage)
        CK(cudaHostAlloc(&f.shard_stage, ((size_t)f.K * f.P + f.K) * (size_t)f.M * 8, cudaHostAllocDefault));
    f.shard_runs.clear();
    for (int64_t j = 0; j < f.M;) {
        int64_t e = j + 1;
        while (e < f.M && snps_host[e] == snps_host[e - 1] + 1) ++e;
        f.shard_runs.push_back(snps_host[j]);
        f.shard_runs.push_back(j);
        f.shard_runs.push_back(e - j);
        j = e;
    }
    return 0;
}
// Upload of this rank's SNPs from GLOBAL host arrays.  Page-locked / registered sources (option
// "shard_zero_copy", default on): the GPU gathers its SNPs itself with zero-copy reads -- no host pass.  Any other
// memory (pageable included): a few host threads cut the runs out into page-locked staging (memcpy per
// run and row), then two full-speed DMA copies, the first overlapping the second cut.
extern "C" int vb_fit_set_params_shard(vb_ctx* ctx, const double* mu_g, const double* dl_g) {
    NEED_FIT(ctx);
    if (!f.shard_idx) return vb_fail("sharded transfer: call vb_fit_set_shard first");
    if (g_shard_zero_copy) {
        const double* gmu = device_view(mu_g);
        const double* gdl = device_view(dl_g);
        if (gmu && gdl) {
            const int gx = (int)std::min<int64_t>((f.M + 255) / 256, 2048);
            vb_shard_rows_gather_kernel<<<dim3(gx, f.K * f.P), 256, 0, ctx->stream>>>(gmu, f.mu[f.cur_mu], f.shard_idx,
                                                                                     f.M, f.shard_total);
            CK_LAUNCH(ctx);
            const int gd = (int)std::min<int64_t>((f.M * f.K + 255) / 256, 148 * 64);
            vb_shard_delta_gather_kernel<<<gd, 256, 0, ctx->stream>>>(gdl, f.delta[f.cur_delta], f.shard_idx, f.M, f.K);
            CK_LAUNCH(ctx);
            CK(cudaStreamSynchronize(ctx->stream));
            f.trial_kind = -1;
            return 0;
        }
    }
    const size_t KM = (size_t)f.K * f.M, rows = (size_t)f.K * f.P;
    double* st_mu = f.shard_stage;
    double* st_dl = f.shard_stage + rows * f.M;
    const int64_t* runs = f.shard_runs.data();
    const size_t nruns = f.shard_runs.size() / 3;
    const int64_t M = f.M, Mt = f.shard_total;
    const int K = f.K;
    const int T = (int)std::max<size_t>(1, std::min<size_t>(4, KM * (f.P + 1) / (1u << 20)));
    auto run_threads = [&](auto&& work) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    };
    // mu first; its DMA runs while the threads cut out delta
    run_threads([&](int t) {
        for (size_t r = t; r < rows; r += T)
            for (size_t q = 0; q < nruns; ++q)
                std::memcpy(st_mu + r * M + runs[3 * q + 1], mu_g + r * Mt + runs[3 * q], (size_t)runs[3 * q + 2] * 8);
    });
    CK(cudaMemcpyAsync(f.mu[f.cur_mu], st_mu, KM * f.P * 8, cudaMemcpyHostToDevice, ctx->stream));
    run_threads([&](int t) {
        for (size_t q = t; q < nruns; q += T)
            std::memcpy(st_dl + (size_t)runs[3 * q + 1] * K, dl_g + (size_t)runs[3 * q] * K, (size_t)runs[3 * q + 2] * K * 8);
    });
    return fit_set_params(ctx, nullptr, st_dl, cudaMemcpyHostToDevice);
}
// device address of a host pointer the GPU can access directly, or nullptr
static double* device_view(const void* host) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return reinterpret_cast<double*>(at.devicePointer);
}
extern "C" int vb_host_accessible(const void* host) { return device_view(host) != nullptr; }
extern "C" int vb_host_register(void* host, int64_t bytes) {
    if (device_view(host)) return 0;
    const cudaError_t e = cudaHostRegister(host, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return vb_fail("cudaHostRegister: %s", cudaGetErrorString(e)); }
    return 0;
}
extern "C" int vb_host_unregister(void* host) {
    const cudaError_t e = cudaHostUnregister(host);
    if (e != cudaSuccess) { cudaGetLastError(); return vb_fail("cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return 0;
}
// Download: the GPU scatters this rank's SNPs straight into the (registered) global host arrays
extern "C" int vb_fit_get_params_shard(vb_ctx* ctx, double* mu_host, double* delta_host) {
    NEED_FIT(ctx);
    if (!f.shard_idx) return vb_fail("sharded transfer: call vb_fit_set_shard first");
    double* gmu = device_view(mu_host);
    double* gdl = device_view(delta_host);
    if (!gmu || !gdl) return vb_fail("sharded transfer: the host arrays must be page-locked or registered");
    const int gx = (int)std::min<int64_t>((f.M + 255) / 256, 2048);
    vb_shard_rows_kernel<<<dim3(gx, f.K * f.P), 256, 0, ctx->stream>>>(gmu, f.mu[f.cur_mu], f.shard_idx, f.M, f.shard_total);
    CK_LAUNCH(ctx);
    const int gd = (int)std::min<int64_t>((f.M * f.K + 255) / 256, 148 * 64);
    vb_shard_delta_kernel<<<gd, 256, 0, ctx->stream>>>(gdl, f.delta[f.cur_delta], f.shard_idx, f.M, f.K);
    CK_LAUNCH(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Launch geometry of the tile kernel: W warps per 32-SNP tile and CTAs per SM, chosen to maximise
// resident threads under the shared-memory (logits + merge scratch) and register limits; ties go to
// the smaller W (shorter merge).  Returns W = 0 when the thread-per-SNP kernels should run.
static TilePlan tile_plan(const vb_ctx* ctx, const Fit& f, int akf) {
    TilePlan best{0, 0, 0};
    if (g_tile_mode == 0) return best;
    // automatic: large grids and P >= 3 (where thread-per-SNP parks K logits per SNP in HBM and pays
    // three logs and square roots per component); the tuned three-pass kernel keeps small P <= 2 grids
    if (g_tile_mode < 0 && f.P <= 2 && f.K < 32) return best;
    const int target = f.P == 1 ? VbTileCfg<1>::THREADS_PER_SM
                                : (f.P <= 3 ? VbTileCfg<3>::THREADS_PER_SM : VbTileCfg<5>::THREADS_PER_SM);
    const int maxw = std::min(VB_TILE_MAXW, (f.P == 1 ? VbTileCfg<1>::MAXT : (f.P <= 3 ? VbTileCfg<3>::MAXT : VbTileCfg<5>::MAXT)) / 32);
    const size_t cap = 227 * 1024;
    int best_threads = 0;
    static const int kWidths[] = {1, 2, 3, 4, 6, 8, 12, 16};     // warps per tile (k is split round-robin: any W works)
    for (int W : kWidths) {
        if (W > maxw) break;
        if (g_tile_mode > 0 && W != g_tile_mode) continue;
        const size_t sm = vb_tile_smem(f.K, f.P, W, akf);
        if (sm > cap) continue;
        int ctas = std::min<int>(std::min<int>(target / (32 * W), (int)(cap / (sm + 1024))), 16);
        if (ctas < 1) continue;
        const int threads = ctas * 32 * W;
        if (threads > best_threads) {
            best_threads = threads;
            best.W = W;
            best.smem = sm;
            const int64_t tiles = (f.M + VB_TILE_SNPS - 1) / VB_TILE_SNPS;
            best.grid = (int)std::min<int64_t>(std::min<int64_t>((int64_t)ctx->num_sms * ctas, tiles), f.grid_snp);
        }
    }
    return best;
}
// The plan for a problem shape without a device (host logic only; tests/test_host_logic.py).
extern "C" int vb_debug_tile_plan(int P, int K, int64_t M, int akf, int num_sms, int* W, int* grid,
                                  int64_t* smem_bytes) {
    if (P < 1 || P > VB_MAX_POPS || K < 1 || M < 1 || num_sms < 1 || !W || !grid || !smem_bytes)
        return vb_fail("vb_debug_tile_plan: bad argument");
    vb_ctx ctx;
    ctx.num_sms = num_sms;
    Fit f;
    f.P = P; f.K = K; f.M = M;
    f.grid_snp = (int)std::min<int64_t>((M + 127) / 128, (int64_t)num_sms * 16);
    const TilePlan tp = tile_plan(&ctx, f, akf);
    *W = tp.W; *grid = tp.grid; *smem_bytes = (int64_t)tp.smem;
    return 0;
}
template <int P, int MODE>
static void launch_tile_one(const VbSnpArgs& a, const TilePlan& tp, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(vb_snp_tile_kernel<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    vb_snp_tile_kernel<P, MODE><<<tp.grid, 32 * tp.W, tp.smem, st>>>(a);
}
template <int MODE>
static int launch_snp(vb_ctx* ctx, const VbSnpArgs& a, int P, int grid) {
    cudaStream_t st = ctx->stream;
    ctx->fit.snp_grid_used = grid;
    if constexpr (MODE != VB_MODE_EVAL) {
        const TilePlan tp = tile_plan(ctx, ctx->fit, a.fuse_ann ? a.A * a.K : 0);
        if (tp.W > 0) {
            ctx->fit.snp_grid_used = tp.grid;
            prof_begin(ctx, 1);
            switch (P) {
                case 1: launch_tile_one<1, MODE>(a, tp, st); break;
                case 2: launch_tile_one<2, MODE>(a, tp, st); break;
                case 3: launch_tile_one<3, MODE>(a, tp, st); break;
                case 4: launch_tile_one<4, MODE>(a, tp, st); break;
                case 5: launch_tile_one<5, MODE>(a, tp, st); break;
                case 6: launch_tile_one<6, MODE>(a, tp, st); break;
                default: return vb_fail("unsupported cohort count %d", P);
            }
            prof_end(ctx, 1);
            CK_LAUNCH(ctx);
            return 0;
        }
    }
    const size_t sm = a.fuse_ann ? (size_t)a.A * a.K * (VB_SNP_THREADS / 32) * sizeof(double) : 0;
    prof_begin(ctx, 1);
    if constexpr (MODE != VB_MODE_EVAL) {
        if (P <= 2 && g_three_pass) {       // exact-max softmax, one exp per (k, SNP)
            VbSnpArgs a3 = a;
            size_t sm3 = sm;
            const bool slots = a.fuse_ann && a.A * a.K <= VB_FUSE_ANN_SLOTS && g_ann_slots;
            if (slots) {
                a3.fuse_ann = 2;
                sm3 = (size_t)a.A * a.K * VB_SNP_THREADS * sizeof(double);
            }
            const size_t park = (size_t)a.K * (P + 1) * VB_SNP_THREADS * sizeof(double);
            // (not together with the fused annotation slots: 43 KB per CTA -> 5 CTAs / SM, measured 12 % slower;
            // with the sums by warp shuffles -- option snp_ann_slots=0 -- parking stays on)
            if (g_snp3_park && park <= VB_SNP3_PARK_MAX_BYTES && !slots) {
                static bool carve_set = false;
                if (!carve_set) {       // all shared memory, no L1 needed: the kernel streams
                    cudaFuncSetAttribute(vb_snp3_kernel<1, MODE, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    cudaFuncSetAttribute(vb_snp3_kernel<2, MODE, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    cudaFuncSetAttribute(vb_snp3_kernel<1, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                    cudaFuncSetAttribute(vb_snp3_kernel<2, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                    carve_set = true;
                }
                if (P == 1) vb_snp3_kernel<1, MODE, true><<<grid, VB_SNP_THREADS, park + sm3, st>>>(a3);
                else vb_snp3_kernel<2, MODE, true><<<grid, VB_SNP_THREADS, park + sm3, st>>>(a3);
            } else if (P == 1) vb_snp3_kernel<1, MODE, false><<<grid, VB_SNP_THREADS, sm3, st>>>(a3);
            else vb_snp3_kernel<2, MODE, false><<<grid, VB_SNP_THREADS, sm3, st>>>(a3);
            prof_end(ctx, 1);
            CK_LAUNCH(ctx);
            return 0;
        }
    }
    switch (P) {
        case 1: vb_snp_kernel<1, MODE><<<grid, VB_SNP_THREADS, sm, st>>>(a); break;
        case 2: vb_snp_kernel<2, MODE><<<grid, VB_SNP_THREADS, sm, st>>>(a); break;
        case 3: vb_snp_kernel<3, MODE><<<grid, VB_SNP_THREADS, sm, st>>>(a); break;
        case 4: vb_snp_kernel<4, MODE><<<grid, VB_SNP_THREADS, sm, st>>>(a); break;
        case 5: vb_snp_kernel<5, MODE><<<grid, VB_SNP_THREADS, sm, st>>>(a); break;
        case 6: vb_snp_kernel<6, MODE><<<grid, VB_SNP_THREADS, sm, st>>>(a); break;
        default: return vb_fail("unsupported cohort count %d", P);
    }
    prof_end(ctx, 1);
    CK_LAUNCH(ctx);
    return 0;
}

static void base_args(const Fit& f, VbSnpArgs& a) {
    std::memset(&a, 0, sizeof(a));
    a.K = f.K; a.A = f.A; a.M = f.M;
    a.adj = f.adj; a.se = f.se; a.sld = f.sld; a.ann = f.ann;
    a.prec = f.prec; a.logdet = f.logdet; a.logh = f.logh; a.gfull = f.gfull;
    for (int p = 0; p < VB_MAXP; ++p) a.inv_tau[p] = f.inv_tau[p];
    a.partial = f.part_snp;
    a.fuse_ann = f.akf > 0;
    a.nsp = f.nsp;
}
// route z = pm/se of every cohort into its LD operator's block-order input
static void route_z(const vb_ctx* ctx, VbSnpArgs& a) {
    const Fit& f = ctx->fit;
    for (int p = 0; p < f.P; ++p) {
        a.xbpos[p] = ctx->fit_ld[p]->L.xbpos;
        a.xb[p] = ctx->fit_ld[p]->L.xall;
    }
}

// mat-vecs of every cohort (their inputs were written by the per-SNP kernel straight into each
// operator's block-order buffer) for state slot v; the last cohort's finish kernel also does
// the fixed-order final reduction into stats_dev.
static int finish_eval(vb_ctx* ctx, int v, double* stats_dev) {
    Fit& f = ctx->fit;
    NativeLoop* nl = loop_of(ctx, false);
    VbFinalArgs fa;
    std::memset(&fa, 0, sizeof(fa));
    fa.part_snp = f.part_snp;
    fa.part_fin = f.part_fin;
    fa.stats = stats_dev;
    fa.n_part_snp = f.snp_grid_used;
    fa.n_part_fin = f.grid_fin;
    fa.P = f.P;
    fa.nsp = f.nsp;
    fa.akf = f.akf;
    const bool native = nl && nl->xr_active;
    if (native && nl->want_diff) {
        // convergence partials of the state being evaluated, reduced together with everything else
        const int64_t n = (int64_t)f.P * f.M;
        prof_begin(ctx, 3);
        vb_pm_diff_kernel<<<f.grid_diff, 256, 0, ctx->stream>>>(f.pm[v], f.scal, f.pm_prev, f.pm_ckpt,
                                                                f.pm_next, n, nl->diff_atol,
                                                                nl->diff_rtol, f.part_diff);
        prof_end(ctx, 3);
        CK_LAUNCH(ctx);
        fa.part_diff = f.part_diff;
        fa.n_part_diff = f.grid_diff;
    }
    if (native && nl->xr_ready) {
        VbXrank& xr = fa.xr;
        xr.enabled = 1;
        xr.nranks = nl->xr_nranks;
        xr.rank = nl->xr_rank;
        xr.epoch = nl->epoch;
        for (int r = 0; r < nl->xr_nranks; ++r) {
            xr.peer_box[r] = nl->xr_peer_box[r];
            xr.peer_flag[r] = nl->xr_peer_flag[r];
        }
        xr.host_out = nl->xr_host_out_dev;
        xr.host_flag = nl->xr_host_flag_dev;
        xr.dev_err = nl->xr_dev_err;
        xr.n_sum = 3 * f.P + 3 + f.akf + (fa.part_diff ? 5 : 0);
        xr.n_max = fa.part_diff ? 5 : 0;
        if (xr.n_sum + xr.n_max > VB_XR_MAXVALS)
            return vb_fail("statistics vector of %d values exceeds the exchange mailbox (%d)",
                           xr.n_sum + xr.n_max, VB_XR_MAXVALS);
    }
    for (int p = 0; p < f.P; ++p) {
        LdPop& L = ctx->fit_ld[p]->L;
        fa.do_final = (p == f.P - 1);
        if (ld_apply(ctx, L, nullptr, f.linked[v] + (size_t)p * f.M,
                     f.part_fin + (size_t)p * f.grid_fin, f.grid_fin, &fa))
            return 1;
    }
    return 0;
}

extern "C" int vb_fit_eval(vb_ctx* ctx, double* stats_dev) {
    NEED_FIT(ctx);
    VbSnpArgs a;
    base_args(f, a);
    a.mu_in = f.mu[f.cur_mu];
    a.delta_in = f.delta[f.cur_delta];
    a.pm_out = f.pm[f.cur_vec];
    route_z(ctx, a);
    if (launch_snp<VB_MODE_EVAL>(ctx, a, f.P, f.grid_snp)) return 1;
    f.trial_kind = -1;
    return finish_eval(ctx, f.cur_vec, stats_dev);
}

extern "C" int vb_fit_beta_trial(vb_ctx* ctx, double step, double* stats_dev) {
    NEED_FIT(ctx);
    VbSnpArgs a;
    base_args(f, a);
    const int tv = 1 - f.cur_vec;
    a.mu_in = f.mu[f.cur_mu];
    a.pm_in = f.pm[f.cur_vec];
    a.linked_in = f.linked[f.cur_vec];
    a.step = step;
    a.mu_out = f.mu[1 - f.cur_mu];
    a.delta_out = f.delta[1 - f.cur_delta];
    a.pm_out = f.pm[tv];
    route_z(ctx, a);
    if (launch_snp<VB_MODE_TRIAL>(ctx, a, f.P, f.grid_snp)) return 1;
    f.trial_kind = 0;
    return finish_eval(ctx, tv, stats_dev);
}

extern "C" int vb_fit_refresh_delta(vb_ctx* ctx, double* stats_dev) {
    NEED_FIT(ctx);
    VbSnpArgs a;
    base_args(f, a);
    const int tv = 1 - f.cur_vec;
    a.mu_in = f.mu[f.cur_mu];
    a.delta_out = f.delta[1 - f.cur_delta];
    a.pm_out = f.pm[tv];
    route_z(ctx, a);
    if (launch_snp<VB_MODE_REFRESH>(ctx, a, f.P, f.grid_snp)) return 1;
    f.trial_kind = 1;
    return finish_eval(ctx, tv, stats_dev);
}

extern "C" int vb_fit_accept(vb_ctx* ctx) {
    NEED_FIT(ctx);
    if (f.trial_kind < 0) return vb_fail("vb_fit_accept: no trial state pending");
    if (f.trial_kind == 0) f.cur_mu = 1 - f.cur_mu;
    f.cur_delta = 1 - f.cur_delta;
    f.cur_vec = 1 - f.cur_vec;
    f.trial_kind = -1;
    return 0;
}

extern "C" int vb_fit_sum_annotations(vb_ctx* ctx, double* out_dev) {
    NEED_FIT(ctx);
    dim3 grid(f.grid_ann, f.K);
    prof_begin(ctx, 3);
    if (f.A == 1)
        vb_sum_columns_kernel<<<grid, 256, 0, ctx->stream>>>(f.delta[f.cur_delta], f.M, f.K, f.part_ann);
    else
        vb_sum_annotations_kernel<<<grid, 256, 0, ctx->stream>>>(f.delta[f.cur_delta], f.ann, f.M, f.K,
                                                                 f.A, f.part_ann);
    CK_LAUNCH(ctx);
    const int tot = f.K * f.A;
    vb_sum_annotations_final_kernel<<<(tot * 32 + 127) / 128, 128, 0, ctx->stream>>>(f.part_ann, f.grid_ann,
                                                                                     f.K, f.A, out_dev);
    prof_end(ctx, 3);
    CK_LAUNCH(ctx);
    return 0;
}

extern "C" int vb_fit_posterior(vb_ctx* ctx, double* pm_host, double* pv_host) {
    NEED_FIT(ctx);
    const size_t PM = (size_t)f.P * f.M;
    VbSnpArgs a;
    base_args(f, a);
    a.mu_in = f.mu[f.cur_mu];
    a.delta_in = f.delta[f.cur_delta];
    a.pm_out = f.scratch3;
    a.pv_out = f.scratch3 + 2 * PM;
    a.fuse_ann = 0;
    if (launch_snp<VB_MODE_EVAL>(ctx, a, f.P, f.grid_snp)) return 1;
    if (pm_host) CK(cudaMemcpyAsync(pm_host, f.scratch3, PM * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (pv_host) CK(cudaMemcpyAsync(pv_host, f.scratch3 + 2 * PM, PM * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int vb_fit_pm_diff(vb_ctx* ctx, double atol, double rtol, double* out_dev) {
    NEED_FIT(ctx);
    const int64_t n = (int64_t)f.P * f.M;
    vb_pm_diff_kernel<<<f.grid_diff, 256, 0, ctx->stream>>>(f.pm[f.cur_vec], f.scal, f.pm_prev,
                                                            f.pm_ckpt, f.pm_prev, n, atol, rtol,
                                                            f.part_diff);
    CK_LAUNCH(ctx);
    vb_pm_diff_final_kernel<<<1, 256, 0, ctx->stream>>>(f.part_diff, f.grid_diff, out_dev);
    CK_LAUNCH(ctx);
    return 0;
}
extern "C" int vb_fit_pm_mark(vb_ctx* ctx, int which) {
    NEED_FIT(ctx);
    const int64_t n = (int64_t)f.P * f.M;
    vb_scale_copy_kernel<<<f.grid_diff, 256, 0, ctx->stream>>>(f.pm[f.cur_vec], f.scal,
                                                               which == 0 ? f.pm_prev : f.pm_ckpt, n);
    CK_LAUNCH(ctx);
    return 0;
}

extern "C" int vb_fit_vi_sigma(vb_ctx* ctx, int k0, int k1, double* out_host) {
    NEED_FIT(ctx);
    if (k0 < 0 || k1 > f.K || k0 >= k1) return vb_fail("vb_fit_vi_sigma: bad slice [%d,%d)", k0, k1);
    const size_t len = (size_t)(k1 - k0) * f.P * f.P * f.M;
    double* buf = nullptr;
    CK(cudaMalloc(&buf, len * 8));
    dim3 grid((unsigned)std::min<int64_t>((f.M + 127) / 128, 2048), k1 - k0);
    switch (f.P) {
        case 1: vb_vi_sigma_kernel<1><<<grid, 128, 0, ctx->stream>>>(f.prec, f.sld, f.inv_tau_dev, f.M, k0, k1, buf); break;
        case 2: vb_vi_sigma_kernel<2><<<grid, 128, 0, ctx->stream>>>(f.prec, f.sld, f.inv_tau_dev, f.M, k0, k1, buf); break;
        case 3: vb_vi_sigma_kernel<3><<<grid, 128, 0, ctx->stream>>>(f.prec, f.sld, f.inv_tau_dev, f.M, k0, k1, buf); break;
        case 4: vb_vi_sigma_kernel<4><<<grid, 128, 0, ctx->stream>>>(f.prec, f.sld, f.inv_tau_dev, f.M, k0, k1, buf); break;
        case 5: vb_vi_sigma_kernel<5><<<grid, 128, 0, ctx->stream>>>(f.prec, f.sld, f.inv_tau_dev, f.M, k0, k1, buf); break;
        case 6: vb_vi_sigma_kernel<6><<<grid, 128, 0, ctx->stream>>>(f.prec, f.sld, f.inv_tau_dev, f.M, k0, k1, buf); break;
        default: cudaFree(buf); return vb_fail("unsupported cohort count %d", f.P);
    }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, buf, len * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(buf);
    if (e != cudaSuccess) return vb_fail("vb_fit_vi_sigma: %s", cudaGetErrorString(e));
    return 0;
}

// Device initialisation (MultiPopVI._initialize, reference :643-700), in two steps around the host's
// hyper_delta computation.  fake_mu_host: [P][M] jittered ridge start.
extern "C" int vb_fit_init_delta(vb_ctx* ctx, const double* fake_mu_host) {
    NEED_FIT(ctx);
    if (!fake_mu_host) return vb_fail("vb_fit_init_delta: null argument");
    const size_t PM = (size_t)f.P * f.M;
    double* fm = f.scratch3 + PM;          // middle third: vb_fit_posterior uses the outer two
    CK(cudaMemcpyAsync(fm, fake_mu_host, PM * 8, cudaMemcpyHostToDevice, ctx->stream));
    double* d = f.delta[f.cur_delta];
    const int g = f.grid_snp;
    switch (f.P) {
        case 1: vb_init_delta_kernel<1><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.logdet, f.sld, f.inv_tau_dev, fm, d); break;
        case 2: vb_init_delta_kernel<2><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.logdet, f.sld, f.inv_tau_dev, fm, d); break;
        case 3: vb_init_delta_kernel<3><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.logdet, f.sld, f.inv_tau_dev, fm, d); break;
        case 4: vb_init_delta_kernel<4><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.logdet, f.sld, f.inv_tau_dev, fm, d); break;
        case 5: vb_init_delta_kernel<5><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.logdet, f.sld, f.inv_tau_dev, fm, d); break;
        case 6: vb_init_delta_kernel<6><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.logdet, f.sld, f.inv_tau_dev, fm, d); break;
        default: return vb_fail("unsupported cohort count %d", f.P);
    }
    CK_LAUNCH(ctx);
    f.trial_kind = -1;
    CK(cudaStreamSynchronize(ctx->stream));      // fake_mu_host may be pageable and reused by the caller
    return 0;
}
// mu of the current state from delta (as left by vb_fit_init_delta) and the fake_mu kept in scratch.
extern "C" int vb_fit_init_mu(vb_ctx* ctx) {
    NEED_FIT(ctx);
    const double* d = f.delta[f.cur_delta];
    const double* fm = f.scratch3 + (size_t)f.P * f.M;
    double* mu = f.mu[f.cur_mu];
    const int g = f.grid_snp;
    switch (f.P) {
        case 1: vb_init_mu_kernel<1><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.sld, f.inv_tau_dev, fm, d, mu); break;
        case 2: vb_init_mu_kernel<2><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.sld, f.inv_tau_dev, fm, d, mu); break;
        case 3: vb_init_mu_kernel<3><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.sld, f.inv_tau_dev, fm, d, mu); break;
        case 4: vb_init_mu_kernel<4><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.sld, f.inv_tau_dev, fm, d, mu); break;
        case 5: vb_init_mu_kernel<5><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.sld, f.inv_tau_dev, fm, d, mu); break;
        case 6: vb_init_mu_kernel<6><<<g, VB_SNP_THREADS, 0, ctx->stream>>>(f.K, f.M, f.prec, f.sld, f.inv_tau_dev, fm, d, mu); break;
        default: return vb_fail("unsupported cohort count %d", f.P);
    }
    CK_LAUNCH(ctx);
    f.trial_kind = -1;
    return 0;
}

// ====================================================================================
// Native control loop: one outer iteration of the fit without leaving C++.
//
// Restates VIScheme._nat_grad_step (variational_inference.py:419-450), MultiPopVI._update_beta
// (:762-802), _update_hyper_delta (:825-860), _update_error_scaling (:472-486, :735-738) and the
// convergence test of optimize() (:376-377) with the same thresholds, on host scalars reduced from
// the device statistics.  The Python class keeps an identical loop (used when INFO logging is on,
// and by the CPU tests); this one removes ~0.2 ms of interpreter / framework latency per evaluated
// state, which is what limits strong scaling once a rank's kernels take ~0.2 ms.
// Multi-GPU: statistics are summed with ncclAllReduce on the context's stream (NCCL is resolved
// at run time from the libnccl already loaded in the process).
// ====================================================================================
extern "C" int vb_nccl_unique_id(char* out128) {
    if (load_nccl()) return 1;
    vbNcclUniqueId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc) return vb_fail("ncclGetUniqueId failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    std::memcpy(out128, id.internal, 128);
    return 0;
}

extern "C" int vb_comm_init(vb_ctx* ctx, int nranks, int rank, const char* id128) {
    if (!ctx) return vb_fail("null ctx");
    if (nranks <= 1) return 0;
    if (load_nccl()) return 1;
    CK(cudaSetDevice(ctx->device));
    NativeLoop* nl = loop_of(ctx, true);
    vbNcclUniqueId id;
    std::memcpy(id.internal, id128, 128);
    int rc = g_nccl.CommInitRank(&nl->comm, nranks, id, rank);
    if (rc) return vb_fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    nl->nranks = nranks;
    return 0;
}

// Allocate this rank's mailbox and return its CUDA IPC handle (64 bytes).
extern "C" int vb_xr_create(vb_ctx* ctx, char* handle_out64) {
    if (!ctx) return vb_fail("null ctx");
    CK(cudaSetDevice(ctx->device));
    NativeLoop* nl = loop_of(ctx, true);
    if (!nl->xr_box) {
        CK(cudaMalloc(&nl->xr_box, VB_XR_BOX_BYTES + sizeof(uint32_t) * 16));
        CK(cudaMemset(nl->xr_box, 0, VB_XR_BOX_BYTES + sizeof(uint32_t) * 16));
        nl->xr_dev_err = reinterpret_cast<uint32_t*>(nl->xr_box + VB_XR_BOX_BYTES);
        CK(cudaHostAlloc(&nl->xr_host_out, 2 * VB_XR_MAXVALS * sizeof(double) + 64, cudaHostAllocMapped));
        std::memset(nl->xr_host_out, 0, 2 * VB_XR_MAXVALS * sizeof(double) + 64);
        nl->xr_host_flag = reinterpret_cast<uint32_t*>(nl->xr_host_out + 2 * VB_XR_MAXVALS);
        CK(cudaHostGetDevicePointer(&nl->xr_host_out_dev, nl->xr_host_out, 0));
        nl->xr_host_flag_dev = reinterpret_cast<uint32_t*>(nl->xr_host_out_dev + 2 * VB_XR_MAXVALS);
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, nl->xr_box));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle_out64, &h, 64);
    return 0;
}
// handles: nranks x 64 bytes, in rank order (every rank passes the same array).
extern "C" int vb_xr_open(vb_ctx* ctx, int nranks, int rank, const char* handles) {
    if (!ctx) return vb_fail("null ctx");
    if (nranks < 1 || nranks > VB_XR_MAXRANKS || rank < 0 || rank >= nranks)
        return vb_fail("vb_xr_open: %d ranks unsupported (1..%d)", nranks, VB_XR_MAXRANKS);
    CK(cudaSetDevice(ctx->device));
    NativeLoop* nl = loop_of(ctx, true);
    if (!nl->xr_box) return vb_fail("vb_xr_open: call vb_xr_create first");
    for (int r = 0; r < nranks; ++r) {
        void* base = nl->xr_box;
        if (r != rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + (size_t)r * 64, 64);
            CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        }
        nl->xr_peer_box[r] = reinterpret_cast<double*>(base);
        nl->xr_peer_flag[r] = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(base) + VB_XR_DATA_BYTES);
    }
    nl->xr_nranks = nranks;
    nl->xr_rank = rank;
    nl->xr_ready = true;
    return 0;
}
// host side of the rendezvous: poll the mapped flag until the evaluation `epoch` has been published
static int xr_wait(NativeLoop* nl, uint32_t epoch, double* out, int n) {
    volatile uint32_t* hf = nl->xr_host_flag + (epoch & 1);
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (true) {
        const uint32_t v = *hf;
        if (v == epoch) break;
        if (v == 0xffffffffu) return vb_fail("cross-rank exchange timed out waiting for a peer");
        if ((++spins & 0xfffff) == 0) {
            if (cudaPeekAtLastError() != cudaSuccess) return vb_fail("CUDA error while waiting: %s", cudaGetErrorString(cudaGetLastError()));
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0)
                return vb_fail("timed out waiting for an evaluation to finish");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    std::memcpy(out, nl->xr_host_out + (size_t)(epoch & 1) * VB_XR_MAXVALS, n * sizeof(double));
    return 0;
}

extern "C" int vb_fit_set_constants(vb_ctx* ctx, const double* chi_stat, const double* ld_ranks,
                                    const double* annotation_counts, const double* log_det,
                                    int scale_se) {
    NEED_FIT(ctx);
    NativeLoop* nl = loop_of(ctx, true);
    nl->chi.assign(chi_stat, chi_stat + f.P);
    nl->ranks.assign(ld_ranks, ld_ranks + f.P);
    nl->counts.assign(annotation_counts, annotation_counts + f.A);
    nl->logdet.assign(log_det, log_det + f.K);
    nl->scale_se = scale_se;
    const size_t need = std::max<size_t>(3 * f.P + 3 + VB_FUSE_ANN_MAX + 10, (size_t)f.A * f.K) + 16;
    if (nl->pinned_len < need) {
        if (nl->pinned) cudaFreeHost(nl->pinned);
        CK(cudaHostAlloc(&nl->pinned, need * sizeof(double), cudaHostAllocDefault));
        nl->pinned_len = need;
    }
    cudaFree(nl->stats_dev); cudaFree(nl->ann_dev); cudaFree(nl->diff_dev);
    CK(cudaMalloc(&nl->stats_dev, (3 * f.P + 3 + VB_FUSE_ANN_MAX + 10) * sizeof(double)));
    CK(cudaMalloc(&nl->ann_dev, (size_t)f.A * f.K * sizeof(double)));
    CK(cudaMalloc(&nl->diff_dev, 16 * sizeof(double)));
    nl->ready = true;
    return 0;
}

// device vector -> (all-reduced) host values
// n values are summed over ranks; `tail` further values are copied rank-local (maxima, logging only)
static int reduce_to_host(vb_ctx* ctx, NativeLoop* nl, double* dev, int n, double* out, int tail = 0) {
    if (nl->comm) {
        int rc = g_nccl.AllReduce(dev, dev, (size_t)n, kNcclDouble, kNcclSum, nl->comm, ctx->stream);
        if (rc) return vb_fail("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    }
    CK(cudaMemcpyAsync(nl->pinned, dev, (n + tail) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, nl->pinned, (n + tail) * sizeof(double));
    return 0;
}

static double objective_of(const NativeLoop* nl, int P, const double* s, const double* tau) {
    double loglik = 0.0;
    for (int p = 0; p < P; ++p) {
        const double per_pop = (-0.5 * (s[P + p] + s[2 * P + p]) + s[p]) - 0.5 * nl->chi[p];
        loglik += per_pop / tau[p] - 0.5 * nl->ranks[p] * std::log(tau[p]);
    }
    return loglik - (s[3 * P] + s[3 * P + 1] + s[3 * P + 2]);
}
static inline bool close_to_zero(double a, double atol) { return std::fabs(a) <= atol; }     // np.isclose(a, 0, atol, rtol=0)
static inline bool np_isclose(double a, double b) { return std::fabs(a - b) <= 1e-8 + 1e-5 * std::fabs(b); }

// Bring the statistics of the evaluation just queued to the host (summed over ranks): through the
// mailbox exchange fused into the evaluation's last CTA when available, else NCCL + copy + sync.
static int fetch_stats(vb_ctx* ctx, NativeLoop* nl, int n_sum, int n_tail, double* out, uint32_t epoch = 0) {
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = nl->xr_ready ? xr_wait(nl, epoch ? epoch : nl->epoch, out, n_sum + n_tail)
                                : reduce_to_host(ctx, nl, nl->stats_dev, n_sum, out, n_tail);
    nl->t_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    nl->n_wait++;
    return rc;
}
// {seconds enqueueing, seconds waiting, rendezvous count, reserved} of the native loop so far
extern "C" int vb_fit_timing(vb_ctx* ctx, double* out4) {
    NativeLoop* nl = loop_of(ctx, false);
    if (!nl) return vb_fail("no native loop state");
    out4[0] = nl->t_enqueue; out4[1] = nl->t_wait; out4[2] = (double)nl->n_wait;
    out4[3] = (double)nl->spec_used + 1e-6 * (double)nl->spec_wasted;
    return 0;
}
// delta refresh + evaluation of the resulting state, with the convergence bookkeeping of the new
// state (against prev / ckpt, written to pm_next) reduced in the same rendezvous.
// stats: [3P+3 | akf annotation sums | 10 diff statistics]
static int native_refresh(vb_ctx* ctx, NativeLoop* nl, Fit& f, double* stats, const double* tau,
                          double* obj, vb_step_io* io, double speculate_step = 0.0) {
    const int nbase = 3 * f.P + 3 + f.akf;
    nl->want_diff = nl->xr_ready;
    nl->diff_atol = io->atol;
    nl->diff_rtol = io->rtol;
    const uint32_t e_refresh = ++nl->epoch;
    const auto te0 = std::chrono::steady_clock::now();
    const int rc = vb_fit_refresh_delta(ctx, nl->stats_dev);
    nl->want_diff = false;
    if (rc) return 1;
    if (!nl->xr_ready) {
        // fallback: queue the convergence kernels behind the evaluation, then one NCCL rendezvous
        const int64_t n = (int64_t)f.P * f.M;
        vb_pm_diff_kernel<<<f.grid_diff, 256, 0, ctx->stream>>>(f.pm[1 - f.cur_vec], f.scal, f.pm_prev,
                                                                f.pm_ckpt, f.pm_next, n, io->atol,
                                                                io->rtol, f.part_diff);
        CK_LAUNCH(ctx);
        vb_pm_diff_final_kernel<<<1, 256, 0, ctx->stream>>>(f.part_diff, f.grid_diff, nl->stats_dev + nbase);
        CK_LAUNCH(ctx);
    }
    // the refreshed state is accepted unconditionally (reference :859): flip now
    if (vb_fit_accept(ctx)) return 1;
    nl->spec_pending = false;
    if (speculate_step > 0.0 && nl->xr_ready) {
        // Speculation: the next outer iteration starts with a beta trial at a step size that is
        // already known (L[0] / 1.25).  Queue it right behind the refresh so the GPU never idles
        // while the host reads the refresh statistics; the next call consumes it if nothing touched
        // the state in between, otherwise it is simply discarded.
        nl->spec_epoch = ++nl->epoch;
        if (vb_fit_beta_trial(ctx, speculate_step, nl->stats_dev)) return 1;
        nl->spec_pending = true;
        nl->spec_step = speculate_step;
    }
    nl->t_enqueue += std::chrono::duration<double>(std::chrono::steady_clock::now() - te0).count();
    if (fetch_stats(ctx, nl, nbase + 5, 5, stats, e_refresh)) return 1;
    io->evals++;
    *obj = objective_of(nl, f.P, stats, tau);
    std::memcpy(io->diff, stats + nbase, 10 * sizeof(double));
    return 0;
}

// returns 0 ok, 1 error, 2 "Encountered a numerical error." (reference :793/:797)
extern "C" int vb_fit_iteration(vb_ctx* ctx, vb_step_io* io, double* tau_io, double* hyper_io,
                                double* stats_io) {
    NEED_FIT(ctx);
    NativeLoop* nl = loop_of(ctx, false);
    if (!nl || !nl->ready) return vb_fail("vb_fit_iteration: call vb_fit_set_constants first");
    const int64_t mut_entry = f.mutations;      // (already counts this call)
    const int P = f.P, K = f.K, A = f.A, NS = 3 * P + 3, NSX = NS + f.akf + 10;
    bool spec_usable = nl->spec_pending && nl->spec_mutations + 1 == mut_entry && f.trial_kind == 0;
    if (nl->spec_pending && !spec_usable) nl->spec_wasted++;
    nl->spec_pending = false;
    const double conv_tol = io->has_running ? 0.1 * io->running_elbo_delta : INFINITY;
    double new_elbo_delta = 0.0;
    double obj = io->obj;
    io->trials = 0;
    io->evals = 0;
    io->rejects = 0;
    std::vector<double> stats(NSX, 0.0), trial(NSX, 0.0);
    std::memcpy(stats.data(), stats_io, NS * sizeof(double));
    struct ActiveGuard {          // evaluations queued from here use the fused rendezvous
        NativeLoop* n;
        explicit ActiveGuard(NativeLoop* p) : n(p) { n->xr_active = true; }
        ~ActiveGuard() { n->xr_active = false; }
    } guard(nl);
    bool ann_valid = false;          // stats[NS..NS+akf) hold the accepted state's annotation sums
    double* L = io->L;

    // ---- idx 0: beta (natural-gradient step with backtracking on 1/L)
    {
        bool have_orig = false;
        double orig_obj = 0.0;
        for (int it = 0; it < VB_MAX_NUM_ITERS; ++it) {
            L[0] = std::max(1.0, L[0] / 1.25);
            if (!have_orig) { orig_obj = obj; have_orig = true; }
            double new_obj = orig_obj;
            bool accepted = false, bail = false;
            while (true) {
                const double step = 1.0 / L[0];
                if (spec_usable && step == nl->spec_step) {
                    // the trial queued behind the previous iteration's refresh
                    spec_usable = false;
                    nl->spec_used++;
                    if (fetch_stats(ctx, nl, NS + f.akf, 0, trial.data(), nl->spec_epoch)) return 1;
                } else {
                    if (spec_usable) { spec_usable = false; nl->spec_wasted++; }
                    nl->epoch++;
                    const auto te0 = std::chrono::steady_clock::now();
                    if (vb_fit_beta_trial(ctx, step, nl->stats_dev)) return 1;
                    nl->t_enqueue += std::chrono::duration<double>(std::chrono::steady_clock::now() - te0).count();
                    if (fetch_stats(ctx, nl, NS + f.akf, 0, trial.data())) return 1;
                }
                io->trials++;
                io->evals++;
                new_obj = objective_of(nl, P, trial.data(), tau_io);
                if (new_obj >= orig_obj - VB_REL_TOL * std::fabs(orig_obj) - VB_ABS_TOL) {
                    if (L[0] > VB_L_MAX && !np_isclose(orig_obj, new_obj)) return 2;
                    accepted = true;
                    break;
                }
                if (L[0] > VB_L_MAX) {
                    if (!np_isclose(orig_obj, new_obj)) return 2;
                    bail = true;
                    break;
                }
                L[0] *= io->line_search_rate;
                io->rejects++;
            }
            if (accepted) {
                if (vb_fit_accept(ctx)) return 1;
                stats = trial;
                ann_valid = f.akf > 0;
                obj = new_obj;
            } else if (bail) {
                new_obj = orig_obj;
            }
            new_elbo_delta += new_obj - orig_obj;
            if (close_to_zero(new_obj - orig_obj, conv_tol) || L[0] == 1.0 || L[0] > VB_L_MAX) break;
            orig_obj = new_obj;
        }
    }
    // ---- idx 1: hyper_delta (closed form), L[1] stays 1 => a single pass
    {
        for (int it = 0; it < VB_MAX_NUM_ITERS; ++it) {
            L[1] = std::max(1.0, L[1] / 1.25);
            const double orig_obj = obj;
            std::vector<double> sums((size_t)A * K);
            if (ann_valid) {
                // rode along with the evaluation of the accepted state (index a*K + k)
                std::memcpy(sums.data(), stats.data() + NS, (size_t)A * K * sizeof(double));
            } else {
                if (vb_fit_sum_annotations(ctx, nl->ann_dev)) return 1;
                if (reduce_to_host(ctx, nl, nl->ann_dev, A * K, sums.data())) return 1;
            }
            std::vector<double> g((size_t)A * std::max(K - 1, 1));
            for (int a = 0; a < A; ++a) {
                double tot = 0.0;
                for (int k = 0; k < K; ++k) {
                    double v = sums[(size_t)a * K + k] / (nl->counts[a] + VB_EPSILON);
                    v = std::max(v, VB_EPSILON);
                    hyper_io[(size_t)a * K + k] = v;
                    tot += v;
                }
                for (int k = 0; k < K; ++k) hyper_io[(size_t)a * K + k] /= tot;
                const double last = std::log(hyper_io[(size_t)a * K + K - 1]) - 0.5 * nl->logdet[K - 1];
                for (int k = 0; k + 1 < K; ++k)
                    g[(size_t)a * (K - 1) + k] =
                        (std::log(hyper_io[(size_t)a * K + k]) - 0.5 * nl->logdet[k]) - last;
            }
            if (fit_set_hyper_tables(ctx, hyper_io, g.data())) return 1;
            double new_obj;
            // L[1] is always 1 here, so this is the iteration's last evaluation unless the error
            // scaling is being learned: speculate the next iteration's first beta trial behind it
            const double spec_step = (!nl->scale_se && L[1] == 1.0 && io->speculate)
                                         ? 1.0 / std::max(1.0, L[0] / 1.25) : 0.0;
            if (native_refresh(ctx, nl, f, stats.data(), tau_io, &new_obj, io, spec_step)) return 1;
            ann_valid = f.akf > 0;
            obj = new_obj;
            new_elbo_delta += new_obj - orig_obj;
            if (close_to_zero(new_obj - orig_obj, conv_tol) || L[1] == 1.0 || L[1] > VB_L_MAX) break;
        }
    }
    // ---- idx 2: annotation update is a no-op in this scheme
    L[2] = std::max(1.0, L[2] / 1.25);

    // ---- error scaling (only with --learn-scaling and a small ELBO change)
    if (nl->scale_se && new_elbo_delta < VB_EM_TOL) {
        const double orig_obj = obj;
        for (int p = 0; p < P; ++p)
            tau_io[p] = (nl->chi[p] - 2 * stats[p] + stats[2 * P + p] + stats[P + p]) / nl->ranks[p];
        if (vb_fit_set_tau(ctx, tau_io)) return 1;
        double new_obj;
        if (native_refresh(ctx, nl, f, stats.data(), tau_io, &new_obj, io)) return 1;
        obj = new_obj;
        new_elbo_delta += new_obj - orig_obj;
    }
    io->obj = obj;
    io->elbo_delta = new_elbo_delta;
    std::memcpy(stats_io, stats.data(), NS * sizeof(double));
    nl->spec_mutations = f.mutations;

    // the last refresh of the iteration compared the new posterior mean with prev / ckpt and wrote
    // it to pm_next: it becomes prev now
    std::swap(f.pm_prev, f.pm_next);
    return 0;
}

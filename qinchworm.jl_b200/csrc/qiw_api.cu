// qiw_api.cu — C-ABI entry points of libqinchworm_cuda.so (include/qinchworm.h) and the runtime
// behind them: device-resident model tables, compiled entries, launch planning, NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/qinchworm.h"
#include "qiw_device.cuh"
#include "qiw_host.hpp"

namespace qiw {
cudaError_t launch_scalar_step(bool real_mode, bool dual, const StepParams& p, dim3 grid, int threads, size_t smem, cudaStream_t st);
cudaError_t launch_scalar_run(bool real_mode, bool dual, const RunParams& rp, int n_ctas, int threads, size_t smem, cudaStream_t st);
int scalar_run_max_ctas(bool real_mode, bool dual, int threads, size_t smem, int n_sm);
cudaError_t launch_reduce(const DevEntryDyn* dyn, const DevEntry* entries, const double2* partials, int pitch, int S,
                          double t_i, double t_w, double t_f, double2* out, int n_entries, cudaStream_t st);
cudaError_t launch_finish_step(double2* P, int n_tau, int bsize, const int* diag, int n_diag, double h, int k_f,
                               const double2* contribs, int n_contrib, int do_normalize, double2* hist, cudaStream_t st);
cudaError_t launch_scale_P(double2* P, int n_tau, int bsize, double h, double lambda, cudaStream_t st);
cudaError_t launch_sobol_points(int D, const uint32_t* m, const uint32_t* x0, unsigned long long start,
                                unsigned long long count, uint32_t* out, cudaStream_t st);
cudaError_t launch_dfma_peak(double* out, int blocks, int iters, cudaStream_t st);
cudaError_t launch_block_step(const StepParams& p, const BlockParams& bp, dim3 grid, int threads, size_t smem, cudaStream_t st);
cudaError_t launch_block_mma(const StepParams& p, const BlockParams& bp, const BlockWalkParams& wp, dim3 grid, int threads,
                             size_t smem, cudaStream_t st);
cudaError_t launch_block_walk(const StepParams& p, const BlockParams& bp, const BlockWalkParams& wp, dim3 grid, int threads,
                              size_t smem, cudaStream_t st);
}  // namespace qiw

using namespace qiw;

// ---- NCCL through dlopen (no link-time dependency: the library must load on a CPU-only box) ----
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
        GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
        CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
        GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) { err = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclDouble = 8;  // ncclFloat64
constexpr int kNcclSum = 0;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)16);
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    cudaError_t upload(const T* src, size_t n, cudaStream_t st) {
        cudaError_t e = reserve(n);
        if (e != cudaSuccess || n == 0) return e;
        return cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct EntryDev {
    EntryProgram prog;
    bool valid = false;
    DevBuf<uint4> lane_items;            // scalar models: records of the lane program (+ padding for the prefetch)
    DevBuf<uint4> lane_segdef4;          // packed definitions of the segment-product table
    bool imag_coefs = false;     // every folded coefficient is purely imaginary (real-mode precondition)
    DevBuf<uint64_t> words;      // block models: tree word stream
    DevBuf<uint4> xwords;        // block models: expanded words of the real-arithmetic walker
    std::vector<double> walk_cost;   // block models: estimated cost of every tree for the walker (load balancing)
    DevBuf<uint32_t> tree_off;
    DevBuf<uint32_t> xtree_off;              // block models: [n_units + 1] word offsets of the walker's units in xwords
    std::vector<uint32_t> xtree_off_h;
    std::vector<uint4> xunits_h;             // block models: the walk units' words (host copy)
    DevBuf<double2> coefs;
    DevBuf<int4> dslots;
    DevBuf<uint32_t> sobol;   // m[D][32] + x0[D] of the current call
    std::vector<uint32_t> default_sobol;
};

struct Plan {   // launch plan of one qiw_eval call shape, cached
    std::vector<int> ids;
    uint64_t count = 0;
    bool explicit_mode = false;
    struct Group {
        int item0, n_items; int max_slots; int max_dslots; int max_nodes1 = 2;
        size_t smem[2]; int spb[2], warps[2], aux_off[2], red_off[2];   // [0] complex arithmetic, [1] real arithmetic
    };
    std::vector<Group> groups;
    std::vector<WorkItem> items;
    std::vector<int> item0, n_items;       // per call entry
    std::vector<DevEntryDyn> h_dyn;        // per call entry
    DevBuf<DevEntryDyn> d_dyn;
    DevBuf<double2> d_partials, d_out;
    bool dyn_resident = false;
    DevBuf<double> d_ucache;               // simplex roots cached across calls with the same sequence
    std::vector<size_t> ucache_off;
    bool ucache_enabled = false, ucache_valid = false;
    size_t partial_rows = 0;
    uint64_t max_sb = 1;
    int pitch = 1;
    int block_threads = 0;                 // block models: threads per CTA
    size_t scratch_per_thread = 0;
    bool lane_dual = false;                // some entry's lane program has records shared by two initial sectors: the kernels' DUAL instantiations run
    bool block_real = false;               // block models: planned for the real-arithmetic tree replay (blocks up to 4x4)
    bool block_mma = false;                // block models with blocks of 5 to 8 rows, real arithmetic: FP64 tensor-core kernel
    DevBuf<int> d_bounds;                  // [n_items][warps + 1] tree ranges (block_walk_kernel)
    int bw_warps = 0, bw_max_sp = 0, bw_nI = 0, bw_nD = 0, bw_pool_n = 0;
    DevBuf<uint32_t> d_sobol;
    std::vector<size_t> sobol_off;         // per call entry, offset into d_sobol
    std::vector<uint32_t> h_sobol;
    bool default_sobol_resident = false;
    DevBuf<WorkItem> d_items;
    DevBuf<uint32_t> d_chunk_off;          // scalar models: [chunks + 1] first run of every chunk of every CTA job
    DevBuf<LaneRun> d_runs;                // runs of records (equal shape and initial sector) the chunks consist of
    // persistent run kernel (qiw_inchworm_run on small steps): one job per CTA for the whole run
    struct RunPlan {
        int state = 0;                     // 0 = not built, 1 = usable, -1 = this plan does not fit the run kernel
        int real = -1;
        int n_jobs = 0, n_ctas = 0, threads = 0;
        size_t smem = 0;
        int ok_off = 0, pw_off = 0, red_off = 0, ds_off = 0, P_off = 0, D_off = 0, out_off = 0, rows_staged = 0;
        bool by_sm = false; int ctas_per_sm = 1;
        DevBuf<RunJob> d_jobs;
        DevBuf<int> d_entry_job0, d_cta_job0;
        DevBuf<WorkItem> d_items;
        DevBuf<uint32_t> d_chunk_off;
        DevBuf<LaneRun> d_runs;
        DevBuf<double2> d_partials;
        DevBuf<unsigned int> d_barrier;
    } run;
};
}  // namespace

struct qiw_context {
    int device = 0;
    bool no_device = false;   // QIW_DEVICE_NONE: planning only
    int warps = 8;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    HostModel model;
    bool have_model = false;
    int n_tau = 0;
    double beta = 0;
    std::vector<cplx> hostP;
    DevBuf<double2> dP;
    DevBuf<double> dE;
    struct Table { int kind = 0, n = 0; double beta = 0; bool imag_only = false; DevBuf<double2> y, M; };
    std::vector<uint8_t> p_row_complex;   // per grid point: the stored P row has a non-zero real part
    int last_real_mode = 0;
    std::vector<Table> tables;
    DevBuf<DevDelta> dDeltas;
    std::vector<DevDelta> hDeltas;
    bool deltas_dirty = true;
    std::vector<std::unique_ptr<EntryDev>> entries;
    DevBuf<DevEntry> dEntries;
    bool entries_dirty = true;
    DevBuf<double2> dPerSample, dHist;
    DevBuf<double> dTimes;
    DevBuf<int> dDiag;
    // block models
    DevBuf<int> dDim, dBoff, dEoff, dOpTarget;
    DevBuf<long long> dOpOff;
    DevBuf<double2> dPool, dScratch;
    DevBuf<double> dPoolRe;
    bool pool_real = false;
    DevBuf<const uint64_t*> dWordsPtr;
    DevBuf<const uint32_t*> dTreeOffPtr;
    DevBuf<const uint4*> dXWordsPtr;
    DevBuf<const uint32_t*> dXTreeOffPtr;
    DevBuf<int> dNTrees;
    DevBuf<unsigned long long> dTrace;
    DevBuf<unsigned int> dCounter;   // arrival counters of the step kernel's fused tail (self-resetting), one per time triple
    DevBuf<double> dTimes3;          // batched evaluation: (t_i, t_w, t_f) triples
    DevBuf<double2> dBatchPartials, dBatchOut;
    DevBuf<uint32_t> dSeqSobol;      // randomised qMC: per-sequence Sobol parameters
    DevBuf<DevEntryDyn> dSeqDyn;
    double2* hOut = nullptr;  // pinned
    size_t hOutCap = 0;
    std::vector<std::unique_ptr<Plan>> plans;
    double last_ms = 0;
    int64_t launches = 0;
    // optional per-kernel profiling
    bool profiling = false;
    struct ProfEv { int cls; cudaEvent_t a, b; };
    std::vector<ProfEv> prof_events;
    double prof_ms[QIW_PROFILE_CLASSES] = {0};
    int64_t prof_n[QIW_PROFILE_CLASSES] = {0};
    // NCCL
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    // peer-memory all-reduce
    unsigned char* peerLocal = nullptr;            // this rank's mailbox
    std::vector<unsigned char*> peerPtrs;          // every rank's mailbox as mapped here (own = peerLocal)
    DevBuf<unsigned char*> dPeerPtrs;
    DevBuf<int> dPeerStatus;
    unsigned long long peer_seq = 0;
    bool peer_ready = false;
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            return QIW_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

struct ProfScope {   // records an event pair around one launch when profiling is on
    qiw_context* ctx; int cls; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(qiw_context* c, int k) : ctx(c), cls(k) {
        if (!ctx->profiling) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, ctx->stream);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->prof_events.push_back({cls, a, b});
    }
};

static int ensure_host_out_impl(qiw_context* ctx, size_t n);
static inline int ensure_host_out(qiw_context* ctx, size_t n) { return ensure_host_out_impl(ctx, n); }
static void release_plan(Plan& pl);
static void drop_plans(qiw_context* ctx) {
    for (auto& pl : ctx->plans) release_plan(*pl);
    ctx->plans.clear();
}

static int fail(qiw_context* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

extern "C" {

const char* qiw_version(void) { return "0.1.0"; }

int qiw_create(const qiw_options* opts, qiw_context** out) {
    if (!out) return QIW_ERR_BAD_ARG;
    *out = nullptr;
    if (opts && opts->device == QIW_DEVICE_NONE) {
        qiw_context* c = new qiw_context();
        c->no_device = true;
        c->device = QIW_DEVICE_NONE;
        *out = c;
        return QIW_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        // no CPU fallback by design
        fprintf(stderr, "libqinchworm_cuda: no CUDA device available (%s); this library has no CPU path\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return QIW_ERR_CUDA;
    }
    qiw_context* ctx = new qiw_context();
    int dev = -1;
    if (opts) { dev = opts->device; if (opts->warps_per_block > 0) { int w = std::min(8, opts->warps_per_block); ctx->warps = 1; while (ctx->warps * 2 <= w) ctx->warps *= 2; } }   // power of two
    if (dev < 0) cudaGetDevice(&dev);
    ctx->device = dev;
    if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return QIW_ERR_CUDA;
    }
    *out = ctx;
    return QIW_OK;
}

int qiw_destroy(qiw_context* ctx) {
    if (!ctx) return QIW_OK;
    if (ctx->no_device) { delete ctx; return QIW_OK; }
    cudaSetDevice(ctx->device);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    cudaStreamSynchronize(ctx->stream);
    for (size_t q = 0; q < ctx->peerPtrs.size(); ++q)
        if (ctx->peerPtrs[q] && ctx->peerPtrs[q] != ctx->peerLocal) cudaIpcCloseMemHandle(ctx->peerPtrs[q]);
    if (ctx->peerLocal) cudaFree(ctx->peerLocal);
    ctx->dPeerPtrs.release(); ctx->dPeerStatus.release();
    ctx->dP.release(); ctx->dE.release(); ctx->dDeltas.release(); ctx->dEntries.release();
    ctx->dPerSample.release(); ctx->dTimes.release(); ctx->dHist.release(); ctx->dDiag.release(); ctx->dTrace.release(); ctx->dCounter.release();
    ctx->dTimes3.release(); ctx->dBatchPartials.release(); ctx->dBatchOut.release(); ctx->dSeqSobol.release(); ctx->dSeqDyn.release();
    ctx->dDim.release(); ctx->dBoff.release(); ctx->dEoff.release(); ctx->dOpTarget.release(); ctx->dOpOff.release();
    ctx->dPool.release(); ctx->dPoolRe.release(); ctx->dScratch.release(); ctx->dWordsPtr.release(); ctx->dTreeOffPtr.release(); ctx->dXWordsPtr.release(); ctx->dXTreeOffPtr.release(); ctx->dNTrees.release();
    for (auto& t : ctx->tables) { t.y.release(); t.M.release(); }
    for (auto& e : ctx->entries)
        if (e) { e->lane_items.release(); e->lane_segdef4.release(); e->words.release(); e->xwords.release(); e->xtree_off.release(); e->tree_off.release(); e->coefs.release(); e->dslots.release(); e->sobol.release(); }
    for (auto& pl : ctx->plans) release_plan(*pl);
    ctx->plans.clear();
    if (ctx->hOut) cudaFreeHost(ctx->hOut);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return QIW_OK;
}

const char* qiw_last_error(const qiw_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int qiw_set_model(qiw_context* ctx, int32_t S, const int32_t* dims, const double* energies, int32_t n_ops,
                  const int32_t* op_target, const int64_t* op_mat_off, const double* op_pool, int32_t n_pairs,
                  const int32_t* pair_op_i, const int32_t* pair_op_f, const int32_t* pair_table, int32_t n_corr,
                  const int32_t* corr_A, const int32_t* corr_B) {
    if (!ctx || S <= 0 || !dims || !energies || n_ops < 0 || n_pairs < 0) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_model: bad argument");
    HostModel& m = ctx->model;
    m = HostModel();
    m.S = S;
    m.dim.assign(dims, dims + S);
    m.boff.resize(S); m.eoff.resize(S);
    int off = 0, eoff = 0;
    m.scalar = true; m.maxdim = 0;
    for (int s = 0; s < S; ++s) {
        if (dims[s] <= 0) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_model: non-positive sector dimension");
        m.boff[s] = off; off += dims[s] * dims[s];
        m.eoff[s] = eoff; eoff += dims[s];
        if (dims[s] != 1) m.scalar = false;
        m.maxdim = std::max(m.maxdim, (int)dims[s]);
    }
    m.bsize = off;
    m.energies.assign(energies, energies + eoff);
    m.n_ops = n_ops;
    m.op_target.assign(op_target, op_target + (size_t)n_ops * S);
    m.op_off.assign(op_mat_off, op_mat_off + (size_t)n_ops * S);
    size_t pool_n = 0;
    for (int o = 0; o < n_ops; ++o)
        for (int s = 0; s < S; ++s) {
            int t = m.op_target[(size_t)o * S + s];
            if (t >= S) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_model: operator target out of range");
            if (t >= 0) pool_n = std::max(pool_n, (size_t)m.op_off[(size_t)o * S + s] + (size_t)dims[t] * dims[s]);
        }
    m.pool.resize(pool_n);
    for (size_t k = 0; k < pool_n; ++k) m.pool[k] = cplx(op_pool[2 * k], op_pool[2 * k + 1]);
    m.pair_op_i.assign(pair_op_i, pair_op_i + n_pairs);
    m.pair_op_f.assign(pair_op_f, pair_op_f + n_pairs);
    m.pair_table.assign(pair_table, pair_table + n_pairs);
    for (int p = 0; p < n_pairs; ++p)
        if (pair_op_i[p] < 0 || pair_op_i[p] >= n_ops || pair_op_f[p] < 0 || pair_op_f[p] >= n_ops || pair_table[p] < 0 ||
            pair_table[p] >= kMaxTables)
            return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_model: bad interaction pair");
    m.attachable.assign(S, {});
    for (int s = 0; s < S; ++s)  // findall(op -> haskey(op[1], s), pair_operator_mat), src/expansion.jl:180-183
        for (int p = 0; p < n_pairs; ++p)
            if (m.target(pair_op_i[p], s) >= 0) m.attachable[s].push_back(p);
    m.corr_A.assign(corr_A, corr_A + n_corr);
    m.corr_B.assign(corr_B, corr_B + n_corr);
    ctx->have_model = true;
    if (!ctx->no_device) {
        cudaSetDevice(ctx->device);
        CK(ctx->dE.upload(m.energies.data(), m.energies.size(), ctx->stream));
        CK(ctx->dDim.upload(m.dim.data(), m.dim.size(), ctx->stream));
        CK(ctx->dBoff.upload(m.boff.data(), m.boff.size(), ctx->stream));
        CK(ctx->dEoff.upload(m.eoff.data(), m.eoff.size(), ctx->stream));
        CK(ctx->dOpTarget.upload(m.op_target.data(), m.op_target.size(), ctx->stream));
        std::vector<long long> off64(m.op_off.begin(), m.op_off.end());
        CK(ctx->dOpOff.upload(off64.data(), off64.size(), ctx->stream));
        CK(ctx->dPool.upload((const double2*)m.pool.data(), m.pool.size(), ctx->stream));
        std::vector<double> pre(std::max<size_t>(m.pool.size(), 1), 0.0);
        ctx->pool_real = true;
        for (size_t k = 0; k < m.pool.size(); ++k) { pre[k] = m.pool[k].real(); if (m.pool[k].imag() != 0.0) ctx->pool_real = false; }
        CK(ctx->dPoolRe.upload(pre.data(), pre.size(), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (m.maxdim > 8) return fail(ctx, QIW_ERR_UNSUPPORTED, "qiw_set_model: sector blocks larger than 8x8 are not supported");
    for (auto& e : ctx->entries) if (e) e->valid = false;   // programs depend on the model
    ctx->entries_dirty = true;
    drop_plans(ctx);
    return QIW_OK;
}

int qiw_set_grid(qiw_context* ctx, int32_t n_tau, double beta) {
    if (!ctx || !ctx->have_model || n_tau < 2 || !(beta > 0)) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_grid: bad argument (model first)");
    if (ctx->no_device) { ctx->n_tau = n_tau; ctx->beta = beta; return QIW_OK; }
    cudaSetDevice(ctx->device);
    ctx->n_tau = n_tau; ctx->beta = beta;
    ctx->hostP.assign((size_t)n_tau * ctx->model.bsize, cplx(0));
    ctx->p_row_complex.assign(n_tau, 0);
    CK(ctx->dP.reserve((size_t)n_tau * ctx->model.bsize));
    CK(cudaMemsetAsync(ctx->dP.p, 0, (size_t)n_tau * ctx->model.bsize * sizeof(double2), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return QIW_OK;
}

int qiw_set_delta(qiw_context* ctx, int32_t id, int32_t kind, int32_t n, double beta, const double* values) {
    if (!ctx || id < 0 || id >= kMaxTables || n < 2 || !values || (kind != 0 && kind != 1))
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_delta: bad argument");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context has no device tables");
    cudaSetDevice(ctx->device);
    if ((int)ctx->tables.size() <= id) ctx->tables.resize(id + 1);
    auto& t = ctx->tables[id];
    t.kind = kind; t.n = n; t.beta = beta;
    std::vector<cplx> y(n), M(n);
    for (int k = 0; k < n; ++k) y[k] = cplx(values[2 * k], values[2 * k + 1]);
    natural_spline_second_derivatives(n, beta / (n - 1), y.data(), M.data());
    t.imag_only = true;
    for (int k = 0; k < n; ++k) if (y[k].real() != 0.0 || M[k].real() != 0.0) t.imag_only = false;
    CK(t.y.upload((const double2*)y.data(), n, ctx->stream));
    CK(t.M.upload((const double2*)M.data(), n, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_plans(ctx);      // shared-memory layouts of cached plans depend on the tables' sizes and kinds
    ctx->deltas_dirty = true;
    return QIW_OK;
}

int qiw_set_P(qiw_context* ctx, int32_t first, int32_t count, const double* rows) {
    if (!ctx || ctx->n_tau == 0 || first < 0 || count < 0 || first + count > ctx->n_tau || !rows)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_P: bad argument (grid first)");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context has no device tables");
    cudaSetDevice(ctx->device);
    const size_t bs = ctx->model.bsize;
    for (int k = 0; k < count; ++k) {
        uint8_t c = 0;
        for (size_t el = 0; el < bs; ++el) if (rows[2 * ((size_t)k * bs + el)] != 0.0) { c = 1; break; }
        ctx->p_row_complex[first + k] = c;
    }
    CK(cudaMemcpyAsync(ctx->dP.p + (size_t)first * bs, rows, (size_t)count * bs * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return QIW_OK;
}

int qiw_scale_P(qiw_context* ctx, int32_t k_f, const double* row, double lambda) {
    if (!ctx || ctx->n_tau == 0 || k_f < 0 || k_f >= ctx->n_tau || !std::isfinite(lambda))
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_scale_P: bad argument (grid first)");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context has no device tables");
    cudaSetDevice(ctx->device);
    const size_t bs = ctx->model.bsize;
    if (row) {
        uint8_t c = 0;
        for (size_t el = 0; el < bs; ++el) if (row[2 * el] != 0.0) { c = 1; break; }
        ctx->p_row_complex[k_f] = c;
        int rc = ensure_host_out(ctx, bs);      // pinned staging: the copy is asynchronous, the caller's buffer is free at return
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream)); // the staging buffer may still feed the previous call's copy
        memcpy(ctx->hOut, row, bs * sizeof(double2));
        CK(cudaMemcpyAsync(ctx->dP.p + (size_t)k_f * bs, ctx->hOut, bs * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (lambda != 0.0) {
        CK(launch_scale_P(ctx->dP.p, ctx->n_tau, (int)bs, ctx->beta / (ctx->n_tau - 1), lambda, ctx->stream));
        ctx->launches++;
    }
    return QIW_OK;
}

int qiw_get_P(qiw_context* ctx, int32_t first, int32_t count, double* rows) {
    if (!ctx || ctx->n_tau == 0 || first < 0 || count < 0 || first + count > ctx->n_tau || !rows)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_get_P: bad argument");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context has no device tables");
    cudaSetDevice(ctx->device);
    const size_t bs = ctx->model.bsize;
    CK(cudaMemcpyAsync(rows, ctx->dP.p + (size_t)first * bs, (size_t)count * bs * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return QIW_OK;
}

// Walk units of a block-model entry (host only; also built by planning-only contexts so that the CPU tests
// can replay them): expanded words + the cut into sub-trees and column groups described below.
static int build_walk_units(const HostModel& hm, const EntryProgram& pr, EntryDev& ed, std::string& err) {
    // expanded words: everything the walker needs about an edge, resolved against the model here
    std::vector<uint4> xw(pr.words.size(), make_uint4(0, 0, 0, 0));
    std::vector<char> is_root(pr.words.size(), 0);
    for (size_t t = 0; t + 1 < pr.tree_off.size(); ++t) is_root[pr.tree_off[t]] = 1;
    const size_t n_real = pr.tree_off.empty() ? 0 : pr.tree_off.back();
    for (size_t k = 0; k < n_real; ++k) {
        const uint64_t w = pr.words[k];
        const uint32_t sA = (uint32_t)w & 0xFFFu, sB = ((uint32_t)w >> 12) & 0xFFFu, nc = ((uint32_t)w >> 24) & 0xFFu;
        const uint32_t aux = (uint32_t)(w >> 32) & 0xFFFFu;
        const int op = (int)((w >> 48) & 0xFFF) - 1;
        uint4 x;
        if (is_root[k]) {   // slotA = sector after the node at position 1, aux = initial sector
            const int s_init = (int)aux, s_next = (int)sA;
            x.x = (uint32_t)hm.dim[s_init] | ((uint32_t)hm.dim[s_next] << 4) | ((op >= 0 ? 1u : 0u) << 8) | (nc << 16);
            x.y = 0;
            x.z = op >= 0 ? (uint32_t)hm.op_off[(size_t)op * hm.S + s_init] : 0u;
            x.w = (uint32_t)hm.boff[s_init];
        } else {
            const int s = (int)sA;
            const int tgt = op >= 0 ? hm.target(op, s) : s;
            x.x = (uint32_t)hm.dim[s] | ((uint32_t)hm.dim[tgt] << 4) | ((op >= 0 ? 1u : 0u) << 8) | (nc << 16);
            x.y = (uint32_t)hm.boff[s] | ((sB ? (sB - (uint32_t)pr.nP + 1u) : 0u) << 16);
            x.z = op >= 0 ? (uint32_t)hm.op_off[(size_t)op * hm.S + s] : 0u;
            x.w = aux;
        }
        xw[k] = x;
    }
    // Walk units.  A tree (one topology, one initial sector) is too coarse a unit of work for a warp
    // — an entry has a few dozen trees of very different cost — so trees are cut into sub-trees of
    // bounded cost: a unit = root word, the path from the root down to the sub-tree's root (words
    // copied with one child each, replayed per unit), the sub-tree's pre-order words verbatim.
    // The table offset of an edge's interval is folded into the word here (y, low 16 bits), so the
    // device loop does not track the depth.
    auto edge_cost = [&](const uint4& x, uint32_t d0) {
        const uint32_t ds = x.x & 0xFu, dr = (x.x >> 4) & 0xFu;
        return (double)(ds * ds + ((x.x >> 8) & 1u) * dr * ds) * d0 + 40.0;
    };
    std::vector<size_t> sub_end(xw.size(), 0);
    std::vector<double> sub_cost(xw.size(), 0.0);
    // Trees are independent: big entries are scanned and cut on several host threads, each over a contiguous range
    // of trees; sums and the unit stream are put together in tree order, so the result does not depend on the
    // number of threads (QIW_COMPILE_THREADS=1: sequential; tests compare both bit for bit).
    const size_t n_trees = pr.tree_off.empty() ? 0 : pr.tree_off.size() - 1;
    int n_threads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    if (n_real < 100000) n_threads = 1;   // small entries: a thread costs more than the work
    if (const char* ev = getenv("QIW_COMPILE_THREADS")) n_threads = std::max(1, std::min(atoi(ev), 64));   // an explicit request wins
    n_threads = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(n_trees, 1));
    auto for_tree_ranges = [&](const std::function<void(int, size_t, size_t)>& fn) {
        if (n_threads <= 1) { fn(0, 0, n_trees); return; }
        auto work = [&](int w) {   // ranges of equal word count, not equal tree count
            auto cut = [&](int i) {
                const size_t want = n_real * (size_t)i / (size_t)n_threads;
                return (size_t)(std::lower_bound(pr.tree_off.begin(), pr.tree_off.begin() + n_trees, (uint32_t)want) - pr.tree_off.begin());
            };
            fn(w, w == 0 ? 0 : cut(w), w == n_threads - 1 ? n_trees : cut(w + 1));
        };
        std::vector<std::thread> pool;
        int started = 0;
        try {
            for (; started < n_threads - 1; ++started) pool.emplace_back(work, started);
        } catch (const std::system_error&) {}   // no more threads to be had: the caller does the rest itself
        for (int w = started; w < n_threads; ++w) work(w);
        for (auto& t : pool) t.join();
    };
    std::vector<char> fits_w(n_threads, 1);
    std::vector<double> tree_cost_sum(n_trees, 0.0);
    for_tree_ranges([&](int w_idx, size_t t_begin, size_t t_end) {
    for (size_t t = t_begin; t < t_end; ++t) {
        bool fits = true;
        const size_t r = pr.tree_off[t];
        const uint32_t d0 = xw[r].x & 0xFu;
        struct Frame { size_t k; uint32_t left; };
        std::vector<Frame> st;
        st.push_back({r, xw[r].x >> 16});
        size_t k = r + 1;
        // iterative pre-order scan: depth of an edge = stack size when it is read
        while (!st.empty()) {
            if (st.back().left == 0) {
                const Frame f = st.back(); st.pop_back();
                sub_end[f.k] = k;
                if (!st.empty()) sub_cost[st.back().k] += sub_cost[f.k];
                continue;
            }
            --st.back().left;
            const size_t interval = st.size() - 1;
            const uint32_t off = (uint32_t)(interval * (size_t)hm.bsize) + (xw[k].y & 0xFFFFu);
            if (off > 0xFFFFu) fits = false;
            xw[k].y = (xw[k].y & 0xFFFF0000u) | (off & 0xFFFFu);
            sub_cost[k] = edge_cost(xw[k], std::min(d0, 2u));   // per column group (at most two columns)
            st.push_back({k, xw[k].x >> 16});
            ++k;
        }
        tree_cost_sum[t] = sub_cost[r] * (double)((d0 + 1) / 2);
        if (!fits) fits_w[w_idx] = 0;
    }
    });
    bool fits = true;
    double entry_cost = 0;
    for (int w = 0; w < n_threads; ++w) fits = fits && fits_w[w];
    for (size_t t = 0; t < n_trees; ++t) entry_cost += tree_cost_sum[t];
    if (!fits) { err = "block tables too large for the walker's word format"; return QIW_ERR_UNSUPPORTED; }
    double target = std::max(4000.0, entry_cost / 1024.0);
    if (const char* ev = getenv("QIW_WALK_UNIT_COST")) target = std::max(1.0, atof(ev));   // tests: force deep cuts
    struct Part { std::vector<uint4> units; std::vector<uint32_t> off; std::vector<double> cost; };
    std::vector<Part> parts(n_threads);
    ed.xtree_off_h.clear(); ed.walk_cost.clear();
    for_tree_ranges([&](int w_idx, size_t t_begin, size_t t_end) {
    std::vector<uint4>& units = parts[w_idx].units;
    std::vector<uint32_t>& xtree_off_part = parts[w_idx].off;
    std::vector<double>& walk_cost_part = parts[w_idx].cost;
    for (size_t t = t_begin; t < t_end; ++t) {
        const size_t r = pr.tree_off[t];
        const uint32_t d0 = xw[r].x & 0xFu;
        std::vector<size_t> path;
        // one unit per group of at most two columns of the running product (kernel: block_walk_tree)
        auto emit = [&](size_t k) {
            for (uint32_t c0 = 0; c0 < d0; c0 += 2) {
                const uint32_t nc = std::min(2u, d0 - c0);
                xtree_off_part.push_back((uint32_t)units.size());
                uint4 rw = xw[r];
                rw.x |= (nc << 9) | (c0 << 11);
                double c = sub_cost[k];
                if (k == r) { units.push_back(rw); units.insert(units.end(), xw.begin() + r + 1, xw.begin() + sub_end[r]); }
                else {
                    rw.x = (rw.x & 0xFFFFu) | (1u << 16); units.push_back(rw);
                    for (size_t q : path) { uint4 w = xw[q]; w.x = (w.x & 0xFFFFu) | (1u << 16); units.push_back(w); c += edge_cost(w, nc); }
                    units.insert(units.end(), xw.begin() + k, xw.begin() + sub_end[k]);
                }
                walk_cost_part.push_back(c);
            }
        };
        std::function<void(size_t)> split = [&](size_t k) {
            const uint32_t nc = xw[k].x >> 16;
            if (sub_cost[k] <= target || nc == 0) { emit(k); return; }
            if (k != r) path.push_back(k);
            size_t c = k + 1;
            for (uint32_t i = 0; i < nc; ++i) { split(c); c = sub_end[c]; }
            if (k != r) path.pop_back();
        };
        if (sub_end[r] > r) split(r);
    }
    });
    std::vector<uint4> units;
    {
        size_t total = 0;
        for (const Part& pt : parts) total += pt.units.size();
        units.reserve(total + (size_t)kWalkPrefetch + 1);
        for (Part& pt : parts) {
            const uint32_t base = (uint32_t)units.size();
            for (uint32_t o : pt.off) ed.xtree_off_h.push_back(base + o);
            ed.walk_cost.insert(ed.walk_cost.end(), pt.cost.begin(), pt.cost.end());
            units.insert(units.end(), pt.units.begin(), pt.units.end());
            std::vector<uint4>().swap(pt.units);
        }
    }
    ed.xtree_off_h.push_back((uint32_t)units.size());
    units.insert(units.end(), (size_t)kWalkPrefetch + 1, make_uint4(0, 0, 0, 0));   // the walker reads one word ahead and prefetches kWalkPrefetch ahead
    ed.xunits_h.swap(units);
    return QIW_OK;
}

int qiw_set_topologies(qiw_context* ctx, int32_t entry_id, int32_t mode, int32_t order, int32_t n_pts_after,
                       int32_t corr_idx, int32_t n_top, const int32_t* pairs, const int32_t* parity) {
    if (!ctx || !ctx->have_model || entry_id < 0 || entry_id > 4095 || n_top < 0 || (n_top > 0 && (!parity || (order > 0 && !pairs))))
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_set_topologies: bad argument (model first)");
    if (!ctx->no_device) cudaSetDevice(ctx->device);
    if ((int)ctx->entries.size() <= entry_id) ctx->entries.resize(entry_id + 1);
    if (!ctx->entries[entry_id]) ctx->entries[entry_id].reset(new EntryDev());
    EntryDev& ed = *ctx->entries[entry_id];
    ed.valid = false;
    std::string err;
    int rc = compile_entry(ctx->model, mode, order, n_pts_after, corr_idx, n_top, pairs, parity, ed.prog, err);
    if (rc) return fail(ctx, rc, "qiw_set_topologies: " + err);
    const EntryProgram& pr = ed.prog;
    if (!ctx->model.scalar && ctx->model.maxdim <= 4) {
        rc = build_walk_units(ctx->model, pr, ed, err);
        if (rc) return fail(ctx, rc, "qiw_set_topologies: " + err);
    }
    if (ctx->no_device) { ed.valid = true; return QIW_OK; }
    if (ctx->model.scalar) {   // the lane program: records (padded by 8 words: the walk fetches one record ahead) and
        // the segment-product table's definitions
        std::vector<uint4> items(pr.lane_items.size() / 8 + 8, make_uint4(0u, 0u, 0u, 0u));
        memcpy(items.data(), pr.lane_items.data(), pr.lane_items.size() * sizeof(uint16_t));
        CK(ed.lane_items.upload(items.data(), items.size(), ctx->stream));
        const int st = pr.seg_stride, nw4 = st > 7 ? 2 : 1;
        std::vector<uint16_t> defs((size_t)std::max(pr.nSegL, 1) * nw4 * 8, (uint16_t)0xFFFF);
        for (int j = 0; j < pr.nSegL; ++j) {
            uint16_t* d = defs.data() + (size_t)j * nw4 * 8;
            d[0] = pr.lane_seg_coef[j];
            for (int i = 0; i < st; ++i) d[1 + i] = pr.lane_segdef[(size_t)j * st + i];
        }
        CK(ed.lane_segdef4.upload(reinterpret_cast<const uint4*>(defs.data()), defs.size() / 8, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));   // the staging vectors go out of scope
    }
    if (!ctx->model.scalar) {
        CK(ed.words.upload(pr.words.data(), pr.words.size(), ctx->stream));
        CK(ed.xwords.upload(ed.xunits_h.data(), ed.xunits_h.size(), ctx->stream));
        CK(ed.xtree_off.upload(ed.xtree_off_h.data(), ed.xtree_off_h.size(), ctx->stream));
        CK(ed.tree_off.upload(pr.tree_off.data(), pr.tree_off.size(), ctx->stream));
    }
    std::vector<double2> cf(std::max<size_t>(pr.coefs.size(), 1));
    ed.imag_coefs = true;
    for (size_t k = 0; k < pr.coefs.size(); ++k) {
        cf[k] = make_double2(pr.coefs[k].real(), pr.coefs[k].imag());
        if (pr.coefs[k].real() != 0.0) ed.imag_coefs = false;
    }
    CK(ed.coefs.upload(cf.data(), cf.size(), ctx->stream));
    std::vector<int4> ds(std::max<size_t>(pr.dslots.size(), 1));
    for (size_t k = 0; k < pr.dslots.size(); ++k) ds[k] = make_int4(pr.dslots[k].pos_tail, pr.dslots[k].pos_head, pr.dslots[k].table, 0);
    CK(ed.dslots.upload(ds.data(), ds.size(), ctx->stream));
    ed.default_sobol.assign((size_t)pr.D * 33 + 1, 0u);
    if (pr.D > 0 && sobol_direction_numbers(pr.D, ed.default_sobol.data())) return fail(ctx, QIW_ERR_BAD_ARG, "Sobol dimension too large");
    CK(ed.sobol.reserve((size_t)pr.D * 33 + 1));
    CK(cudaStreamSynchronize(ctx->stream));
    ed.valid = true;
    ctx->entries_dirty = true;
    drop_plans(ctx);
    return QIW_OK;
}

int qiw_entry_stats(qiw_context* ctx, int32_t id, int64_t* n_top, int64_t* n_leaves, int64_t* n_edges, double* flops) {
    if (!ctx || id < 0 || id >= (int)ctx->entries.size() || !ctx->entries[id] || !ctx->entries[id]->valid)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_entry_stats: unknown entry");
    const EntryProgram& p = ctx->entries[id]->prog;
    if (n_top) *n_top = p.n_top;
    if (n_leaves) *n_leaves = p.n_leaves;
    if (n_edges) *n_edges = p.n_edges;
    if (flops) *flops = p.flops_per_sample;
    return QIW_OK;
}

int qiw_entry_program(qiw_context* ctx, int32_t id, int64_t* n_words, uint64_t* words, int64_t* n_trees,
                      uint32_t* tree_off, int64_t* n_coefs, double* coefs, int64_t* n_dslots, int32_t* dslots,
                      int32_t* pos_src, int32_t* info) {
    if (!ctx || id < 0 || id >= (int)ctx->entries.size() || !ctx->entries[id] || !ctx->entries[id]->valid)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_entry_program: unknown entry");
    const EntryProgram& p = ctx->entries[id]->prog;
    if (n_words) *n_words = (int64_t)p.words.size();
    if (n_trees) *n_trees = (int64_t)p.tree_off.size() - 1;
    if (n_coefs) *n_coefs = (int64_t)p.coefs.size();
    if (n_dslots) *n_dslots = (int64_t)p.dslots.size();
    if (words) memcpy(words, p.words.data(), p.words.size() * sizeof(uint64_t));
    if (tree_off) memcpy(tree_off, p.tree_off.data(), p.tree_off.size() * sizeof(uint32_t));
    if (coefs) for (size_t k = 0; k < p.coefs.size(); ++k) { coefs[2 * k] = p.coefs[k].real(); coefs[2 * k + 1] = p.coefs[k].imag(); }
    if (dslots) for (size_t k = 0; k < p.dslots.size(); ++k) { dslots[3 * k] = p.dslots[k].pos_tail; dslots[3 * k + 1] = p.dslots[k].pos_head; dslots[3 * k + 2] = p.dslots[k].table; }
    if (pos_src) for (int k = 0; k <= kMaxNodes; ++k) pos_src[k] = p.pos_src[k];
    if (info) { info[0] = p.n_nodes; info[1] = p.nP; info[2] = ctx->model.S; info[3] = p.scalar ? 1 : 0; }
    return QIW_OK;
}

int qiw_entry_records(qiw_context* ctx, int32_t id, int32_t* info, uint32_t* rec2, uint16_t* segdef) {
    if (!ctx || id < 0 || id >= (int)ctx->entries.size() || !ctx->entries[id] || !ctx->entries[id]->valid)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_entry_records: unknown entry");
    const EntryProgram& p = ctx->entries[id]->prog;
    if (!p.scalar) return fail(ctx, QIW_ERR_UNSUPPORTED, "qiw_entry_records: only 1x1-block models have configuration records");
    if (info) {
        info[0] = p.K; info[1] = p.L2; info[2] = (int32_t)p.n_leaves; info[3] = p.nSeg; info[4] = p.seg_stride;
        info[5] = p.nP; info[6] = (int32_t)p.dslots.size(); info[7] = 0;
    }
    if (rec2) memcpy(rec2, p.rec2.data(), p.rec2.size() * sizeof(uint32_t));
    if (segdef) memcpy(segdef, p.segdef.data(), (size_t)p.nSeg * p.seg_stride * sizeof(uint16_t));
    return QIW_OK;
}

int qiw_entry_walk_units(qiw_context* ctx, int32_t id, int64_t* n_units, int64_t* n_words, uint32_t* unit_off, uint32_t* words) {
    if (!ctx || id < 0 || id >= (int)ctx->entries.size() || !ctx->entries[id] || !ctx->entries[id]->valid)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_entry_walk_units: unknown entry");
    const EntryDev& ed = *ctx->entries[id];
    if (ed.prog.scalar) return fail(ctx, QIW_ERR_UNSUPPORTED, "qiw_entry_walk_units: only sector-block models have walk units");
    if (n_units) *n_units = (int64_t)ed.xtree_off_h.size() - 1;
    if (n_words) *n_words = ed.xtree_off_h.empty() ? 0 : (int64_t)ed.xtree_off_h.back();
    if (unit_off) memcpy(unit_off, ed.xtree_off_h.data(), ed.xtree_off_h.size() * sizeof(uint32_t));
    if (words && !ed.xtree_off_h.empty()) memcpy(words, ed.xunits_h.data(), (size_t)ed.xtree_off_h.back() * sizeof(uint4));
    return QIW_OK;
}

int qiw_entry_lane_program(qiw_context* ctx, int32_t id, int32_t* info, int32_t* sections, uint32_t* items, uint16_t* segdef,
                           uint16_t* seg_coef) {
    if (!ctx || id < 0 || id >= (int)ctx->entries.size() || !ctx->entries[id] || !ctx->entries[id]->valid)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_entry_lane_program: unknown entry");
    const EntryProgram& p = ctx->entries[id]->prog;
    if (!p.scalar) return fail(ctx, QIW_ERR_UNSUPPORTED, "qiw_entry_lane_program: only 1x1-block models have a lane program");
    if (info) {
        info[0] = (int32_t)p.lane_sections.size(); info[1] = (int32_t)p.lane_items.size(); info[2] = p.nSegL; info[3] = p.seg_stride;
        info[4] = p.K; info[5] = p.order; info[6] = p.nP + (int32_t)p.dslots.size(); info[7] = (int32_t)p.lane_cost;
    }
    if (sections)
        for (size_t k = 0; k < p.lane_sections.size(); ++k) {
            const auto& sc = p.lane_sections[k];
            sections[4 * k] = sc.s_i | ((sc.s_b + 1) << 16); sections[4 * k + 1] = sc.M; sections[4 * k + 2] = (int32_t)sc.n_rec; sections[4 * k + 3] = (int32_t)(sc.chunk0 * 8u);
        }
    if (items) for (size_t k = 0; k < p.lane_items.size(); ++k) items[k] = p.lane_items[k];
    if (segdef) memcpy(segdef, p.lane_segdef.data(), (size_t)p.nSegL * p.seg_stride * sizeof(uint16_t));
    if (seg_coef) memcpy(seg_coef, p.lane_seg_coef.data(), (size_t)p.nSegL * sizeof(uint16_t));
    return QIW_OK;
}

}  // extern "C"

// ---- launch planning ---------------------------------------------------------------------------


static int sync_static_tables(qiw_context* ctx) {
    if (ctx->deltas_dirty) {
        std::vector<DevDelta> dd(std::max<size_t>(ctx->tables.size(), 1));
        for (size_t t = 0; t < ctx->tables.size(); ++t) {
            auto& tb = ctx->tables[t];
            dd[t].y = tb.y.p; dd[t].M = tb.M.p; dd[t].kind = tb.kind; dd[t].n = tb.n;
            dd[t].h = tb.n > 1 ? tb.beta / (tb.n - 1) : 1.0;
            dd[t].inv_h = 1.0 / dd[t].h;
        }
        CK(ctx->dDeltas.upload(dd.data(), dd.size(), ctx->stream));
        ctx->hDeltas = dd;
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->deltas_dirty = false;
    }
    if (ctx->entries_dirty) {
        std::vector<DevEntry> de(std::max<size_t>(ctx->entries.size(), 1));
        memset(de.data(), 0, de.size() * sizeof(DevEntry));
        for (size_t i = 0; i < ctx->entries.size(); ++i) {
            if (!ctx->entries[i] || !ctx->entries[i]->valid) continue;
            EntryDev& ed = *ctx->entries[i];
            const EntryProgram& p = ed.prog;
            DevEntry& d = de[i];
            d.mode = p.mode; d.order = p.order; d.n_nodes = p.n_nodes; d.D = p.D;
            d.d_after = (p.mode == 0) ? p.D : p.n_pts_after;
            d.d_before = p.D - d.d_after;
            d.nP = p.nP; d.nD = (int)p.dslots.size();
            d.n_coefs = (int)p.coefs.size();
            d.K = p.K; d.nSegL = p.nSegL; d.seg_stride = p.seg_stride;
            d.lane_items = ed.lane_items.p; d.lane_segdef4 = ed.lane_segdef4.p;
            d.exact = (p.order == 0);
            for (int k = 0; k <= kDevMaxNodes; ++k) d.pos_src[k] = p.pos_src[k];
            d.coefs = ed.coefs.p; d.dslots = ed.dslots.p;
        }
        CK(ctx->dEntries.upload(de.data(), de.size(), ctx->stream));
        if (!ctx->model.scalar) {
            std::vector<const uint64_t*> wp(de.size(), nullptr);
            std::vector<const uint32_t*> tp(de.size(), nullptr);
            std::vector<const uint4*> xp(de.size(), nullptr);
            std::vector<const uint32_t*> xtp(de.size(), nullptr);
            std::vector<int> nt(de.size(), 0);
            for (size_t i = 0; i < ctx->entries.size(); ++i) {
                if (!ctx->entries[i] || !ctx->entries[i]->valid) continue;
                wp[i] = ctx->entries[i]->words.p; tp[i] = ctx->entries[i]->tree_off.p; xp[i] = ctx->entries[i]->xwords.p;
                xtp[i] = ctx->entries[i]->xtree_off.p;
                nt[i] = (int)ctx->entries[i]->prog.tree_off.size() - 1;
            }
            CK(ctx->dWordsPtr.upload(wp.data(), wp.size(), ctx->stream));
            CK(ctx->dTreeOffPtr.upload(tp.data(), tp.size(), ctx->stream));
            CK(ctx->dXWordsPtr.upload(xp.data(), xp.size(), ctx->stream));
            CK(ctx->dXTreeOffPtr.upload(xtp.data(), xtp.size(), ctx->stream));
            CK(ctx->dNTrees.upload(nt.data(), nt.size(), ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->entries_dirty = false;
    }
    return QIW_OK;
}

static void release_plan(Plan& pl) {
    pl.d_items.release(); pl.d_sobol.release(); pl.d_ucache.release(); pl.d_bounds.release();
    pl.d_dyn.release(); pl.d_partials.release(); pl.d_out.release(); pl.d_chunk_off.release(); pl.d_runs.release();
    pl.run.d_jobs.release(); pl.run.d_entry_job0.release(); pl.run.d_cta_job0.release(); pl.run.d_items.release(); pl.run.d_chunk_off.release();
    pl.run.d_runs.release(); pl.run.d_partials.release(); pl.run.d_barrier.release();
}

// Real arithmetic is exact when every operand is (real number) * i: the stored P rows and Delta tables
// purely imaginary, the folded coefficients purely imaginary.  QIW_FORCE_COMPLEX=1
// disables it (tests compare the two modes).
static bool real_mode_possible(qiw_context* ctx, int n_entries, const int32_t* ids) {
    if (!ctx->model.scalar && !ctx->pool_real) return false;
    if (const char* env = getenv("QIW_FORCE_COMPLEX")) if (env[0] == '1') return false;
    for (auto& t : ctx->tables) if (t.n > 0 && !t.imag_only) return false;
    for (uint8_t c : ctx->p_row_complex) if (c) return false;
    for (int i = 0; i < n_entries; ++i) if (!ctx->entries[ids[i]]->imag_coefs) return false;
    return true;
}

// Cuts an entry's lane program into `n_chunks` chunks of equal cumulative cost and appends them: one entry of
// `chunk_off` per chunk (the caller terminates the array), a chunk being a list of runs — one per section it touches.
static void append_entry_chunks(const EntryProgram& p, int n_chunks, std::vector<uint32_t>& chunk_off, std::vector<LaneRun>& runs) {
    size_t si = 0; uint32_t r_in = 0;   // cursor: section, record inside it
    int64_t done = 0;
    for (int c = 0; c < n_chunks; ++c) {
        chunk_off.push_back((uint32_t)runs.size());
        const int64_t target = (int64_t)((double)p.lane_cost * (double)(c + 1) / (double)n_chunks + 0.5);
        const bool last = (c + 1 == n_chunks);
        while (si < p.lane_sections.size() && (last || done < target)) {
            const auto& sec = p.lane_sections[si];
            uint32_t take = sec.n_rec - r_in;
            if (!last) take = (uint32_t)std::min<int64_t>(take, std::max<int64_t>(1, (target - done + sec.cost - 1) / sec.cost));
            const int ni = lane_record_items(p.order, p.K, sec.M) / 8;
            const uint32_t mc = sec.M == 1 ? 0u : (sec.M == 2 ? 1u : 2u);
            runs.push_back(make_uint4(sec.chunk0 + r_in * (uint32_t)ni, take, (uint32_t)sec.s_i | ((uint32_t)(sec.s_b + 1) << 16),
                                      (uint32_t)(p.order * 16 + (p.K - 1) * 4) + mc));
            done += (int64_t)take * sec.cost;
            r_in += take;
            if (r_in == sec.n_rec) { ++si; r_in = 0; }
        }
    }
}

// Splits every entry's trees into chunks of similar cost and groups chunks into CTA jobs.
// Plans are cached per (entry list, sample count, mode).
static int get_plan(qiw_context* ctx, int n_entries, const int32_t* ids, uint64_t count, bool explicit_mode, Plan** out) {
    for (auto& up : ctx->plans) {
        Plan& c = *up;
        if (c.count == count && c.explicit_mode == explicit_mode && (int)c.ids.size() == n_entries &&
            std::equal(ids, ids + n_entries, c.ids.begin()) &&
            (ctx->model.scalar || (c.block_real || c.block_mma) == (!explicit_mode && real_mode_possible(ctx, n_entries, ids)))) { *out = &c; return QIW_OK; }
    }
    if (ctx->plans.size() >= 6) { release_plan(*ctx->plans.front()); ctx->plans.erase(ctx->plans.begin()); }
    std::unique_ptr<Plan> pl(new Plan());
    pl->ids.assign(ids, ids + n_entries);
    pl->count = count; pl->explicit_mode = explicit_mode;
    const int S = ctx->model.S;
    int ndev_sm = 148;
    cudaDeviceGetAttribute(&ndev_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    if (!ctx->model.scalar && !explicit_mode && ctx->model.maxdim <= 4 && real_mode_possible(ctx, n_entries, ids)) {
        // Block models, real arithmetic: CTA = (entry, 32 samples, group of tree chunks), W warps per CTA,
        // every warp replays a contiguous range of the entry's trees of about equal cost (live edges).
        pl->block_real = true;
        const int bs = ctx->model.bsize;
        int nI_max = 1, nD_max = 1, max_order = 0;
        double total_cost = 0;
        std::vector<double> entry_cost;
        uint64_t n_sb = 1;
        for (int i = 0; i < n_entries; ++i) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            nI_max = std::max(nI_max, p.n_nodes - 1);
            nD_max = std::max(nD_max, (int)p.dslots.size());
            max_order = std::max(max_order, p.order);
            { double c = 0; for (double x : ctx->entries[ids[i]]->walk_cost) c += x; total_cost += c; entry_cost.push_back(c); }
            n_sb = std::max<uint64_t>(n_sb, ((p.order == 0 ? 1 : count) + 31) / 32);
        }
        const int max_sp = max_order + 2;
        if (max_sp > kWalkMaxSp) return fail(ctx, QIW_ERR_UNSUPPORTED, "expansion order too high for the block walker's branch stack");
        // shared memory: the per-sample tables of 32 samples + per-warp block sums.  The branch-point stack lives in
        // per-thread local memory, and ONE CTA of 24 warps runs per SM: its 24 warps share one set of tables, which
        // leaves most of the SM's 256 KB to L1 — where the stack frames and the word streams then stay (measured:
        // two CTAs of 12 warps with two sets of tables 41.6 ms, one CTA of 24 warps 35.3 ms on the two-band model)
        // the operator pool (real parts) rides along in shared memory when it is small
        const size_t pool_stage = ctx->model.pool.size() <= 2048 ? ctx->model.pool.size() : 0;
        pl->bw_pool_n = (int)pool_stage;
        auto smem_of = [&](int Wn) {
            return ((size_t)nI_max * bs * 32 + (size_t)nD_max * 32 + (size_t)Wn * bs + (size_t)(kDevMaxNodes + 1) * 32 +
                    (size_t)kDevMaxDim * 32) * sizeof(double) + 32 * sizeof(int) + pool_stage * sizeof(double) + 64;
        };
        int Wn = 24;
        if (const char* ev = getenv("QIW_WALK_WARPS")) { const int v = atoi(ev); if (v >= 1 && v <= 24) Wn = v; }
        if (smem_of(Wn) > (size_t)226 * 1024) return fail(ctx, QIW_ERR_UNSUPPORTED, "per-sample block tables exceed shared memory");
        // CTA jobs: enough to fill the machine about four times over, proportional to the entries' cost
        double want_ctas = 4.0 * ndev_sm;   // one 24-warp CTA is resident per SM; about four rounds of jobs
        if (const char* ev = getenv("QIW_WALK_CTAS_PER_SM")) want_ctas = std::max(1.0, atof(ev)) * ndev_sm;
        std::vector<int> bounds;
        pl->item0.resize(n_entries); pl->n_items.resize(n_entries);
        for (int i = 0; i < n_entries; ++i) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            const int n_trees = (int)ctx->entries[ids[i]]->xtree_off_h.size() - 1;   // walk units
            int jobs = (int)std::ceil(want_ctas * entry_cost[i] / std::max(total_cost, 1.0) / (double)n_sb);
            jobs = std::max(1, std::min(jobs, (n_trees + Wn - 1) / Wn));
            const int n_chunks = jobs * Wn;
            // chunk boundaries at equal cumulative cost
            std::vector<double> cum(n_trees + 1, 0.0);
            const std::vector<double>& wc = ctx->entries[ids[i]]->walk_cost;
            for (int t = 0; t < n_trees; ++t) cum[t + 1] = cum[t] + wc[t] + 60.0;
            pl->item0[i] = (int)pl->items.size();
            int t_prev = 0;
            for (int j = 0; j < jobs; ++j) {
                WorkItem it;
                it.entry = ids[i]; it.slot = i; it.chunk0 = j * Wn; it.n_chunks = Wn; it.n_chunks_total = n_chunks;
                it.partial0 = (int)pl->items.size();
                pl->items.push_back(it);
                bounds.push_back(t_prev);
                for (int w2 = 1; w2 <= Wn; ++w2) {
                    const double target = cum[n_trees] * (double)(j * Wn + w2) / (double)n_chunks;
                    int t_end = (int)(std::lower_bound(cum.begin(), cum.end(), target - 1e-9) - cum.begin());
                    t_end = std::max(t_prev, std::min(t_end, n_trees));
                    if (j == jobs - 1 && w2 == Wn) t_end = n_trees;
                    bounds.push_back(t_end);
                    t_prev = t_end;
                }
            }
            pl->n_items[i] = jobs;
        }
        CK(pl->d_bounds.upload(bounds.data(), bounds.size(), ctx->stream));
        Plan::Group g;
        g.item0 = 0; g.max_slots = 1; g.max_dslots = 1;
        g.n_items = (int)pl->items.size();
        g.smem[0] = g.smem[1] = smem_of(Wn);
        g.spb[0] = g.spb[1] = 32;
        pl->groups.push_back(g);
        pl->max_sb = n_sb;
        pl->block_threads = 0;
        pl->bw_warps = Wn; pl->bw_max_sp = max_sp; pl->bw_nI = nI_max; pl->bw_nD = nD_max;
    } else if (!ctx->model.scalar && !explicit_mode && real_mode_possible(ctx, n_entries, ids)) {
        // Blocks of 5 to 8 rows, real arithmetic: FP64 tensor cores (block_mma_kernel), warp = sample.
        // CTA = (entry, chunk of its trees, Wn consecutive samples).
        pl->block_mma = true;
        const int bs = ctx->model.bsize;
        int nI_max = 1, nD_max = 1;
        for (int i = 0; i < n_entries; ++i) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            nI_max = std::max(nI_max, p.n_nodes - 1);
            nD_max = std::max(nD_max, (int)p.dslots.size());
        }
        auto smem_of = [&](int Wn) { return (size_t)Wn * ((size_t)nI_max * bs + nD_max + (kDevMaxNodes + 1) + bs) * sizeof(double); };
        int Wn = 8;
        while (Wn > 1 && smem_of(Wn) > (size_t)100 * 1024) Wn >>= 1;      // two CTAs per SM
        if (smem_of(Wn) > (size_t)226 * 1024) return fail(ctx, QIW_ERR_UNSUPPORTED, "per-sample block tables exceed shared memory");
        uint64_t n_sb_max = 1;
        for (int i = 0; i < n_entries; ++i) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            n_sb_max = std::max<uint64_t>(n_sb_max, ((p.order == 0 ? 1 : count) + Wn - 1) / Wn);
        }
        const int split = (int)std::max<double>(1.0, std::ceil(8.0 * ndev_sm / ((double)n_entries * (double)n_sb_max)));
        pl->item0.resize(n_entries); pl->n_items.resize(n_entries);
        for (int i = 0; i < n_entries; ++i) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            const int n_trees = (int)p.tree_off.size() - 1;
            const int n_chunks = std::max(1, std::min(n_trees, split));
            pl->item0[i] = (int)pl->items.size();
            for (int c0 = 0; c0 < n_chunks; ++c0) {
                WorkItem it;
                it.entry = ids[i]; it.slot = i; it.chunk0 = c0; it.n_chunks = 1; it.n_chunks_total = n_chunks;
                it.partial0 = (int)pl->items.size();
                pl->items.push_back(it);
            }
            pl->n_items[i] = n_chunks;
        }
        Plan::Group g;
        g.item0 = 0; g.max_slots = 1; g.max_dslots = 1;
        g.n_items = (int)pl->items.size();
        g.smem[0] = g.smem[1] = smem_of(Wn);
        g.spb[0] = g.spb[1] = Wn;
        pl->groups.push_back(g);
        pl->max_sb = n_sb_max;
        pl->block_threads = 0;
        pl->bw_warps = Wn; pl->bw_nI = nI_max; pl->bw_nD = nD_max;
    } else if (!ctx->model.scalar) {
        // Block models, complex arithmetic (general fallback on the device) and per-sample evaluation:
        // one thread per (sample, chunk of trees); 64 samples per CTA.
        const int TPB = 64;
        uint64_t n_sb_max = 1;
        for (int i = 0; i < n_entries; ++i) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            n_sb_max = std::max<uint64_t>(n_sb_max, ((p.order == 0 ? 1 : count) + TPB - 1) / TPB);
        }
        const int split = explicit_mode ? 1 : (int)std::max<double>(1.0, std::ceil(4.0 * ndev_sm / ((double)n_entries * (double)n_sb_max)));
        pl->item0.resize(n_entries); pl->n_items.resize(n_entries);
        Plan::Group g;
        g.item0 = 0; g.max_slots = 1; g.max_dslots = 1;
        size_t spt = 1;
        for (int i = 0; i < n_entries; ++i) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            const int n_trees = (int)p.tree_off.size() - 1;
            const int n_chunks = std::max(1, std::min(n_trees, split));
            pl->item0[i] = (int)pl->items.size();
            for (int c0 = 0; c0 < n_chunks; ++c0) {
                WorkItem it;
                it.entry = ids[i]; it.slot = i; it.chunk0 = c0; it.n_chunks = 1; it.n_chunks_total = n_chunks;
                it.partial0 = (int)pl->items.size();
                pl->items.push_back(it);
            }
            pl->n_items[i] = n_chunks;
            spt = std::max(spt, (size_t)(p.n_nodes - 1) * ctx->model.bsize + p.dslots.size() + ctx->model.bsize);
        }
        g.n_items = (int)pl->items.size();
        g.smem[0] = g.smem[1] = (size_t)(TPB / 32) * ctx->model.bsize * sizeof(double2);
        g.spb[0] = g.spb[1] = TPB;
        pl->groups.push_back(g);
        pl->max_sb = n_sb_max;
        pl->block_threads = TPB;
        pl->scratch_per_thread = spt;
    } else {
    // Chunking.  A CTA owns (entry, 32 samples) and builds that sample block's tables once, so the
    // fewer CTAs share an entry the less set-up work is repeated: by default an entry's lane program
    // is split into exactly W chunks of equal cost (one per warp, one CTA job).  Only entries whose chunks
    // would exceed `chunk_cap` operand loads are split further (their set-up cost is then amortised
    // anyway), and when the call is too small to fill the machine the cap is lowered to expose more CTAs.
    double max_cost = 1;
    uint64_t n_sb_all = 1;
    int max_pd = 1;
    Plan::Group g;
    g.item0 = 0; g.max_slots = 1; g.max_dslots = 1;
    for (int i = 0; i < n_entries; ++i) {
        const EntryProgram& p = ctx->entries[ids[i]]->prog;
        max_cost = std::max(max_cost, (double)p.lane_cost);
        max_pd = std::max(max_pd, p.nP + (int)p.dslots.size());
        g.max_slots = std::max(g.max_slots, p.nP + (int)p.dslots.size() + p.nSegL);
        g.max_dslots = std::max(g.max_dslots, (int)p.dslots.size());
        g.max_nodes1 = std::max(g.max_nodes1, p.n_nodes + 1);
    }
    // samples per CTA pass, warps per CTA and shared memory, for complex (16-byte operands) and real (8-byte)
    // arithmetic: 32 samples (one per lane) unless the tables of the largest entry do not fit, then the largest power of
    // two that does; 24 warps per SM as 3 CTAs of 8, 2 of 12 or 1 of 24, whichever the table size allows (all warps of
    // a CTA share one table, and the walk needs the warps: its operand loads are latency-bound below ~6 per scheduler)
    const size_t aux_bytes = (size_t)g.max_nodes1 * 32 * (2 * sizeof(double) + sizeof(int)) + 32 * sizeof(int);
    for (int real = 0; real < 2; ++real) {
        const size_t opsz = real ? sizeof(double) : sizeof(double2);
        auto aux_off_of = [&](int spb) {   // after the propagator / interaction rows and the roots, inside the table if it is long enough
            size_t lo = std::max((size_t)max_pd * spb * opsz, (size_t)kDevMaxDim * 32 * sizeof(double));
            size_t tb = (size_t)g.max_slots * spb * opsz;
            size_t off = std::max(lo, tb > aux_bytes ? tb - aux_bytes : (size_t)0);
            return (off + 15) & ~(size_t)15;
        };
        auto red_off_of = [&](int spb) {
            return (std::max((size_t)g.max_slots * spb * opsz, aux_off_of(spb) + aux_bytes) + 15) & ~(size_t)15;
        };
        auto smem_of = [&](int spb, int Wn) {
            return red_off_of(spb) + (size_t)S * Wn * sizeof(double2) + (size_t)g.max_dslots * sizeof(uint32_t) + 16;
        };
        const size_t cap = (size_t)226 * 1024;
        int spb = 32;
        while (spb > 1 && smem_of(spb, 24) > cap) spb >>= 1;
        if (smem_of(spb, 24) > cap) return fail(ctx, QIW_ERR_UNSUPPORTED, "per-sample tables exceed shared memory (too many sectors for the scalar kernel)");
        if (const char* env = getenv("QIW_SPB")) {   // tuning override (power of two <= 32)
            const int v = atoi(env);
            if (v >= 1 && v <= 32 && (v & (v - 1)) == 0 && smem_of(v, 24) <= cap) spb = v;
        }
        const size_t per_sm = (size_t)227 * 1024;    // 1 KB per resident CTA is reserved by the system
        int Wn = 24;
        if (3 * (smem_of(spb, 8) + 1024) <= per_sm) Wn = 8;
        else if (2 * (smem_of(spb, 12) + 1024) <= per_sm) Wn = 12;
        if (const char* env = getenv("QIW_STEP_WARPS")) { const int v = atoi(env); if (v >= 1 && v <= 24 && smem_of(spb, v) <= cap) Wn = v; }
        g.spb[real] = spb; g.warps[real] = Wn;
        g.smem[real] = smem_of(spb, Wn);
        g.aux_off[real] = (int)aux_off_of(spb); g.red_off[real] = (int)red_off_of(spb);
    }
    const int real_plan = real_mode_possible(ctx, n_entries, ids) ? 1 : 0;
    const int W = g.warps[real_plan];
    const int spb_plan = g.spb[real_plan];
    for (int i = 0; i < n_entries; ++i) {
        const EntryProgram& p = ctx->entries[ids[i]]->prog;
        const uint64_t c = p.order == 0 ? 1 : count;
        n_sb_all = std::max<uint64_t>(n_sb_all, (c + spb_plan - 1) / spb_plan);
    }
    double chunk_cap = 3072.0;   // operand loads per warp and sample block
    if (const char* env = getenv("QIW_CHUNK_CAP")) chunk_cap = std::max(16.0, atof(env));
    {
        // CTAs available if every entry is one job; lower the cap until ~3 CTAs per SM exist
        const double ctas = (double)n_entries * (double)n_sb_all;
        const double want = 3.0 * ndev_sm;
        if (ctas < want) chunk_cap = std::max(16.0, std::min(chunk_cap, std::floor(max_cost / W / std::ceil(want / ctas))));
    }
    pl->item0.resize(n_entries); pl->n_items.resize(n_entries);
    // heavy entries first: their CTAs are the critical path of the launch
    std::vector<int> order_idx(n_entries);
    for (int i = 0; i < n_entries; ++i) order_idx[i] = i;
    std::stable_sort(order_idx.begin(), order_idx.end(), [&](int a2, int b2) {
        return ctx->entries[ids[a2]]->prog.lane_cost > ctx->entries[ids[b2]]->prog.lane_cost; });
    {
        std::vector<uint32_t> chunk_off;
        std::vector<LaneRun> runs;
        for (int i : order_idx) {
            const EntryProgram& p = ctx->entries[ids[i]]->prog;
            const uint64_t c = p.order == 0 ? 1 : count;
            pl->max_sb = std::max<uint64_t>(pl->max_sb, (c + spb_plan - 1) / spb_plan);
            int64_t n_rec = 0;
            for (const auto& sec : p.lane_sections) { n_rec += sec.n_rec; if (sec.s_b >= 0) pl->lane_dual = true; }
            const int jobs = (int)std::max(1.0, std::ceil((double)p.lane_cost / (W * chunk_cap)));
            const int n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(n_rec, (int64_t)W * jobs));
            pl->item0[i] = (int)pl->items.size();
            for (int c0 = 0; c0 < n_chunks; c0 += W) {
                WorkItem it;
                it.entry = ids[i]; it.slot = i; it.chunk0 = (int)chunk_off.size() + c0; it.n_chunks = std::min(W, n_chunks - c0);
                it.n_chunks_total = n_chunks;
                it.partial0 = (int)pl->items.size();
                pl->items.push_back(it);
            }
            append_entry_chunks(p, n_chunks, chunk_off, runs);
            pl->n_items[i] = (int)pl->items.size() - pl->item0[i];
        }
        chunk_off.push_back((uint32_t)runs.size());
        if (runs.empty()) runs.push_back(make_uint4(0u, 0u, 0u, 0u));
        CK(pl->d_chunk_off.upload(chunk_off.data(), chunk_off.size(), ctx->stream));
        CK(pl->d_runs.upload(runs.data(), runs.size(), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        g.n_items = (int)pl->items.size() - g.item0;
        pl->groups.push_back(g);
    }
    }
    // grid.x (common to all groups = row pitch of the partials buffer): enough sample-block
    // columns to fill the machine a few times over; the kernel strides over the remaining blocks
    {
        const uint64_t want_ctas = (uint64_t)ndev_sm * 8;
        uint64_t gx = std::max<uint64_t>(1, want_ctas / std::max<size_t>(1, pl->items.size()));
        pl->pitch = (int)std::min<uint64_t>(pl->max_sb, std::min<uint64_t>(gx, 65535));
    }
    pl->partial_rows = pl->items.size() * (size_t)pl->pitch;
    // static part of the per-entry call data
    pl->h_dyn.assign(n_entries, DevEntryDyn());
    size_t tot = 0;
    for (int i = 0; i < n_entries; ++i) { pl->sobol_off.push_back(tot); tot += (size_t)ctx->entries[ids[i]]->prog.D * 33 + 1; }
    pl->h_sobol.assign(tot, 0u);
    CK(pl->d_sobol.reserve(tot));
    CK(pl->d_dyn.reserve(n_entries));
    {   // cache of the unit-simplex roots: 8 bytes per (entry dimension, sample); skipped when large
        size_t ud = 0;
        for (int i = 0; i < n_entries; ++i) { pl->ucache_off.push_back(ud); ud += (size_t)ctx->entries[ids[i]]->prog.D * count; }
        pl->ucache_enabled = !explicit_mode && ud > 0 && ud * sizeof(double) <= ((size_t)2 << 30);
        if (pl->ucache_enabled) CK(pl->d_ucache.reserve(ud));
    }
    for (int i = 0; i < n_entries; ++i) {
        DevEntryDyn& dy = pl->h_dyn[i];
        dy.sobol = pl->d_sobol.p + pl->sobol_off[i];
        dy.entry = ids[i]; dy.out_index = i; dy.item0 = pl->item0[i]; dy.n_items = pl->n_items[i];
    }
    CK(pl->d_items.upload(pl->items.data(), pl->items.size(), ctx->stream));
    CK(pl->d_partials.reserve(pl->partial_rows * ctx->model.bsize));
    if (pl->block_threads) CK(ctx->dScratch.reserve((size_t)pl->pitch * pl->items.size() * pl->block_threads * pl->scratch_per_thread));
    CK(pl->d_out.reserve((size_t)n_entries * ctx->model.bsize));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->plans.push_back(std::move(pl));
    *out = ctx->plans.back().get();
    return QIW_OK;
}

// Per-call data of a plan: Sobol parameters, sample range, weights.
static int stage_call(qiw_context* ctx, Plan& pl, const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t start,
                      uint64_t count, uint64_t N_total, bool allreduce) {
    const int n_entries = (int)pl.ids.size();
    const bool default_sobol = (sobol_m == nullptr && sobol_x0 == nullptr);
    if (!(default_sobol && pl.default_sobol_resident)) {
        size_t moff = 0, xoff = 0;
        for (int i = 0; i < n_entries; ++i) {
            EntryDev& ed = *ctx->entries[pl.ids[i]];
            const int D = ed.prog.D;
            uint32_t* sb = pl.h_sobol.data() + pl.sobol_off[i];
            memcpy(sb, sobol_m ? sobol_m + moff : ed.default_sobol.data(), (size_t)D * 32 * sizeof(uint32_t));
            if (sobol_x0) memcpy(sb + (size_t)D * 32, sobol_x0 + xoff, (size_t)D * sizeof(uint32_t));
            else memset(sb + (size_t)D * 32, 0, (size_t)D * sizeof(uint32_t));
            moff += (size_t)D * 32; xoff += D;
        }
        CK(cudaMemcpyAsync(pl.d_sobol.p, pl.h_sobol.data(), pl.h_sobol.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        pl.default_sobol_resident = default_sobol;
        pl.ucache_valid = false;   // new sequence parameters: the cached roots are stale
    }
    bool changed = !pl.dyn_resident;
    for (int i = 0; i < n_entries; ++i) {
        const EntryProgram& p = ctx->entries[pl.ids[i]]->prog;
        DevEntryDyn& dy = pl.h_dyn[i];
        unsigned long long s2, c2;
        double w;
        if (p.order == 0) {
            // exact entries are identical on every rank: only rank 0 contributes to the sum
            s2 = 0; c2 = 1; w = (allreduce && ctx->rank != 0) ? 0.0 : 1.0;
        } else {
            s2 = start; c2 = count; w = pl.explicit_mode ? 1.0 : 1.0 / (double)N_total;
        }
        if (dy.start != s2 || dy.count != c2) pl.ucache_valid = false;
        double* uc = (pl.ucache_enabled && p.order > 0) ? pl.d_ucache.p + pl.ucache_off[i] : nullptr;
        const int uv = (uc && pl.ucache_valid) ? 1 : 0;
        if (dy.start != s2 || dy.count != c2 || dy.weight != w || dy.ucache != uc || dy.ucache_valid != uv) changed = true;
        dy.start = s2; dy.count = c2; dy.weight = w; dy.ucache = uc; dy.ucache_valid = uv;
    }
    if (changed) {
        CK(cudaMemcpyAsync(pl.d_dyn.p, pl.h_dyn.data(), pl.h_dyn.size() * sizeof(DevEntryDyn), cudaMemcpyHostToDevice, ctx->stream));
        pl.dyn_resident = true;
    }
    return QIW_OK;
}

// How long a rank waits inside the kernel for its peers' block sums before it gives up (poisons the sums, raises the
// status flag): ranks reach a collective seconds apart when one of them compiles a big entry or checks results on the host.
static unsigned long long peer_timeout_ns() {
    double s = 60.0;
    if (const char* env = getenv("QIW_PEER_TIMEOUT_S")) { const double v = atof(env); if (v > 0) s = v; }
    return (unsigned long long)(s * 1e9);
}

// What the step kernel's last CTA does after the reduction in the device-resident loop.
struct FinishArgs { int k_f = -1; int normalize = 0; double2* hist = nullptr; const int* diag = nullptr; int n_diag = 0; };

// Enqueue the kernels of one evaluation of a plan at fixed times: the step kernel (whose last CTA
// also performs the deterministic reduction and, if `fin` asks for it, the P update) — or, for block
// models, step kernel + reduction kernel.  Nothing is synchronised here.  Returns through
// `finish_done` whether the P update was fused.
static int enqueue_step(qiw_context* ctx, Plan& pl, double t_i, double t_w, double t_f, const FinishArgs* fin = nullptr,
                        bool* finish_done = nullptr, bool collective = false, bool* collective_done = nullptr,
                        int n_times = 0, const uint32_t* dev_seq_sobol = nullptr, int sobol_z_stride = 0) {
    if (finish_done) *finish_done = false;
    if (collective_done) *collective_done = false;
    const HostModel& m = ctx->model;
    StepParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.entries = ctx->dEntries.p; sp.dyn = dev_seq_sobol ? ctx->dSeqDyn.p : pl.d_dyn.p;
    sp.P = ctx->dP.p; sp.E = ctx->dE.p; sp.deltas = ctx->dDeltas.p;
    for (int t = 0; t < kInlineTables && t < (int)ctx->hDeltas.size(); ++t) sp.deltas_inline[t] = ctx->hDeltas[t];
    sp.tables_on_grid = (ctx->tables.size() <= (size_t)kInlineTables) ? 1 : 0;
    for (auto& tb : ctx->tables)
        if (tb.n > 0 && (tb.kind != 0 || tb.n != ctx->n_tau || tb.beta != ctx->beta)) sp.tables_on_grid = 0;
    sp.S = m.S; sp.bsize = m.bsize; sp.n_tau = ctx->n_tau; sp.h = ctx->beta / (ctx->n_tau - 1); sp.inv_h = 1.0 / sp.h;
    sp.t_i = t_i; sp.t_w = t_w; sp.t_f = t_f;
    sp.partials = pl.d_partials.p;
    sp.finish_k_f = -1;
    sp.sobol_z_stride = sobol_z_stride;
    if ((m.scalar || pl.block_real || pl.block_mma) && !pl.explicit_mode) {
        const size_t need = (size_t)std::max(n_times, 1);
        if (ctx->dCounter.cap < need) {
            CK(cudaStreamSynchronize(ctx->stream));
            CK(ctx->dCounter.reserve(std::max<size_t>(need, 1024)));
            CK(cudaMemsetAsync(ctx->dCounter.p, 0, ctx->dCounter.cap * sizeof(unsigned int), ctx->stream));
        }
        sp.done_counter = ctx->dCounter.p;
        sp.n_call_entries = (int)pl.ids.size();
        sp.out = pl.d_out.p;
        if (n_times > 0) {   // batched: per-triple partial rows and results, times read on the device
            sp.times_dev = ctx->dTimes3.p;
            sp.partials = ctx->dBatchPartials.p;
            sp.out = ctx->dBatchOut.p;
        }
        bool exchange_fused = false;
        if (collective && ctx->peer_ready && ctx->n_ranks > 1 && pl.ids.size() * (size_t)m.bsize * sizeof(double2) * 2 <= kPeerSlotBytes) {   // every double travels as two 8-byte words
            sp.peer_ranks = ctx->n_ranks; sp.peer_rank = ctx->rank; sp.peer_seq = ++ctx->peer_seq;
            sp.peer_mail = ctx->dPeerPtrs.p; sp.peer_status = ctx->dPeerStatus.p; sp.peer_timeout_ns = peer_timeout_ns();
            exchange_fused = true;
            if (collective_done) *collective_done = true;
        }
        // the P update can be fused only if the sums are already global when the tail reaches it
        if (fin && fin->k_f >= 0 && (exchange_fused || ctx->n_ranks == 1 || !collective)) {
            sp.finish_k_f = fin->k_f; sp.finish_normalize = fin->normalize; sp.finish_P = ctx->dP.p;
            sp.finish_diag = fin->diag; sp.finish_n_diag = fin->n_diag; sp.finish_hist = fin->hist;
            if (finish_done) *finish_done = true;
        }
    }
    if (pl.explicit_mode) { sp.explicit_times = ctx->dTimes.p; sp.per_sample_out = ctx->dPerSample.p; }
    const char* trace_path = getenv("QIW_TRACE");   // diagnostics: per-CTA timeline of the last step kernel
    size_t trace_words = 0;
    if (trace_path) {
        trace_words = (size_t)pl.pitch * pl.items.size() * 12;
        CK(ctx->dTrace.reserve(trace_words));
        CK(cudaMemsetAsync(ctx->dTrace.p, 0, trace_words * sizeof(unsigned long long), ctx->stream));
        sp.trace = ctx->dTrace.p;
    }
    if (!m.scalar) {
        BlockParams bp;
        bp.m.dim = ctx->dDim.p; bp.m.boff = ctx->dBoff.p; bp.m.eoff = ctx->dEoff.p; bp.m.op_target = ctx->dOpTarget.p;
        bp.m.op_off = ctx->dOpOff.p; bp.m.pool = ctx->dPool.p; bp.m.S = m.S; bp.m.bsize = m.bsize; bp.m.maxdim = m.maxdim;
        bp.m.n_ops = m.n_ops;
        bp.words = ctx->dWordsPtr.p; bp.tree_off = ctx->dTreeOffPtr.p; bp.n_trees = ctx->dNTrees.p;
        bp.scratch = ctx->dScratch.p; bp.scratch_per_thread = pl.scratch_per_thread;
        StepParams gp = sp;
        gp.items = pl.d_items.p;
        dim3 grid((unsigned)pl.pitch, (unsigned)pl.items.size(), (unsigned)((pl.block_real || pl.block_mma) ? std::max(n_times, 1) : 1));
        if (pl.block_mma) {
            BlockWalkParams wp;
            memset(&wp, 0, sizeof(wp));
            wp.pool_re = ctx->dPoolRe.p; wp.warps = pl.bw_warps; wp.nI_max = pl.bw_nI; wp.nD_max = pl.bw_nD;
            ctx->last_real_mode = 1;
            ProfScope ps(ctx, 3);
            CK(launch_block_mma(gp, bp, wp, grid, pl.bw_warps * 32, pl.groups[0].smem[0], ctx->stream));
        } else if (pl.block_real) {
            BlockWalkParams wp;
            wp.pool_n = pl.bw_pool_n;
            wp.pool_re = ctx->dPoolRe.p; wp.xwords = ctx->dXWordsPtr.p; wp.unit_off = ctx->dXTreeOffPtr.p; wp.chunk_bounds = pl.d_bounds.p; wp.warps = pl.bw_warps; wp.max_sp = pl.bw_max_sp;
            wp.nI_max = pl.bw_nI; wp.nD_max = pl.bw_nD;
            ctx->last_real_mode = 1;
            ProfScope ps(ctx, 7);
            CK(launch_block_walk(gp, bp, wp, grid, pl.bw_warps * 32, pl.groups[0].smem[0], ctx->stream));
        } else {
            ctx->last_real_mode = 0;
            ProfScope ps(ctx, 7);
            CK(launch_block_step(gp, bp, grid, pl.block_threads, pl.groups[0].smem[0], ctx->stream));
        }
        ctx->launches++;
    } else
    for (auto& g : pl.groups) {
        StepParams gp = sp;
        gp.items = pl.d_items.p + g.item0;
        const int real = real_mode_possible(ctx, (int)pl.ids.size(), pl.ids.data()) ? 1 : 0;
        ctx->last_real_mode = real;
        gp.max_slots = g.max_slots;
        gp.max_nodes1 = g.max_nodes1;
        gp.max_dslots = g.max_dslots;
        gp.chunk_off = pl.d_chunk_off.p; gp.runs = pl.d_runs.p;
        gp.spb = g.spb[real];
        gp.aux_off = g.aux_off[real]; gp.red_off = g.red_off[real];
        {
            // measured on B200 (C1): overlapping consecutive steps gains nothing (7.25 vs 6.96 ms per run), because
            // the dependent grid can only start once the last wave of the running grid has started; off by default
            const char* env = getenv("QIW_PDL");
            gp.allow_overlap = (env && env[0] == '1') ? 1 : 0;
        }
        gp.spb_log2 = 0;
        while ((1 << gp.spb_log2) < gp.spb) ++gp.spb_log2;
        dim3 grid((unsigned)pl.pitch, (unsigned)g.n_items, (unsigned)std::max(n_times, 1));
        {
            ProfScope ps(ctx, real ? 1 : 0);
            CK(launch_scalar_step(real != 0, pl.lane_dual, gp, grid, g.warps[real] * 32, g.smem[real], ctx->stream));
        }
        ctx->launches++;
    }
    if (trace_path) {
        std::vector<unsigned long long> tr(trace_words);
        CK(cudaMemcpyAsync(tr.data(), ctx->dTrace.p, trace_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (FILE* f = fopen(trace_path, "w")) {
            fprintf(f, "cta,start_clk,tables_clk,walk_clk,end_clk,smid,entry,groups_warp0,start_ns,roots_clk,times_clk,fill_clk,tail_clk\n");
            for (size_t c = 0; c < trace_words / 12; ++c)
                fprintf(f, "%zu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu\n", c, tr[c * 12], tr[c * 12 + 1], tr[c * 12 + 2], tr[c * 12 + 3],
                        tr[c * 12 + 4], tr[c * 12 + 5], tr[c * 12 + 6], tr[c * 12 + 7], tr[c * 12 + 8], tr[c * 12 + 9], tr[c * 12 + 10], tr[c * 12 + 11]);
            fclose(f);
        }
    }
    if (!pl.explicit_mode && !m.scalar && !pl.block_real && !pl.block_mma) {
        {
            ProfScope ps(ctx, 4);
            CK(launch_reduce(pl.d_dyn.p, ctx->dEntries.p, pl.d_partials.p, pl.pitch, m.bsize, t_i, t_w, t_f, pl.d_out.p,
                             (int)pl.ids.size(), ctx->stream));
        }
        ctx->launches++;
    }
    return QIW_OK;
}

// ---- persistent run kernel: planning and launch ----------------------------------------------------
// The bold steps of qiw_inchworm_run as ONE cooperative launch (scalar_run_kernel) when a step is too small to fill
// the machine: every (entry, group of sample blocks, part of the lane program) becomes a job, one job per CTA for
// the whole run.  Light entries take m = 2, 4, 8, 16 sample blocks per job (their tables are small), heavy entries
// are split over several jobs, until the jobs are about equally long and fit the co-resident CTAs.
// `*used` = false when the plan does not fit (too many samples, more than kInlineTables tables, shared memory): the
// caller then issues one step kernel per step as before.
static int enqueue_run(qiw_context* ctx, Plan& pl, int k_first, int n_steps, double2* hist, size_t hist_stride, size_t hist_off,
                       const int* diag, int n_diag, bool collective, bool* used) {
    *used = false;
    const bool verbose = getenv("QIW_RUN_VERBOSE") != nullptr;
    auto skip = [&](const char* why) { if (verbose) fprintf(stderr, "qiw run kernel not used: %s\n", why); return QIW_OK; };
    const HostModel& m = ctx->model;
    if (!m.scalar || pl.explicit_mode || n_steps < 2) return skip("block model / explicit times / fewer than two steps");
    if (const char* env = getenv("QIW_NO_RUN_KERNEL")) if (env[0] == '1') return skip("QIW_NO_RUN_KERNEL");
    if (ctx->tables.size() > (size_t)kInlineTables) return skip("too many pair-interaction tables");
    bool on_grid = true;
    size_t table_elems = 0;       // staged pair-interaction tables: grid values (+ second derivatives for splines)
    for (auto& tb : ctx->tables) {
        if (tb.n > 0 && (tb.kind != 0 || tb.n != ctx->n_tau || tb.beta != ctx->beta)) on_grid = false;
        table_elems += (size_t)tb.n * (tb.kind == 1 ? 2 : 1);
    }
    if (on_grid) table_elems = ctx->tables.size() * (size_t)ctx->n_tau;
    if (collective && ctx->n_ranks > 1 && !ctx->peer_ready) return skip("multi-GPU without peer mailboxes");
    const int n_ent = (int)pl.ids.size(), S = m.S, n_tau = ctx->n_tau;
    if (collective && ctx->n_ranks > 1 && (size_t)n_ent * m.bsize * sizeof(double2) * 2 > kPeerSlotBytes) return skip("block sums exceed a mailbox slot");
    const int real = real_mode_possible(ctx, n_ent, pl.ids.data()) ? 1 : 0;
    Plan::RunPlan& rn = pl.run;
    if (rn.state != 0 && rn.real != real) rn.state = 0;
    if (rn.state < 0) return skip("plan does not fit (cached decision)");
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    if (rn.state == 0) {
        rn.state = -1; rn.real = real;
        int W = 12, ctas_per_sm = 2;     // tuning overrides: QIW_RUN_WARPS (<= 12), QIW_RUN_CTAS_PER_SM
        if (const char* env = getenv("QIW_RUN_WARPS")) { const int v = atoi(env); if (v >= 1 && v <= 12) W = v; }
        if (const char* env = getenv("QIW_RUN_CTAS_PER_SM")) { const int v = atoi(env); if (v >= 1 && v <= 4) ctas_per_sm = v; }
        const int threads = W * 32;
        const int G = ctas_per_sm * n_sm;
        const size_t cap = (size_t)228 * 1024 / ctas_per_sm - 1024;     // 228 KB per SM, 1 KB per resident CTA reserved by the system
        const size_t opsz = real ? sizeof(double) : sizeof(double2);
        const int n_tables = (int)ctx->tables.size();
        struct Ent { const EntryProgram* p; int64_t n_sb; int mm, r; int slots, pd; int64_t n_rec; };
        std::vector<Ent> ev(n_ent);
        int max_dslots = 1;
        for (int i = 0; i < n_ent; ++i) {
            const EntryProgram& p = ctx->entries[pl.ids[i]]->prog;
            Ent& x = ev[i];
            x.p = &p; x.mm = 1; x.r = 1;
            x.n_sb = p.order == 0 ? 1 : (int64_t)((pl.count + 31) / 32);
            x.pd = p.nP + (int)p.dslots.size(); x.slots = x.pd + p.nSegL;
            x.n_rec = 0;
            for (const auto& sec : p.lane_sections) x.n_rec += sec.n_rec;
            max_dslots = std::max(max_dslots, (int)p.dslots.size());
        }
        auto jobs_of = [&](const Ent& x) { return (int64_t)((x.n_sb + x.mm - 1) / x.mm) * x.r; };
        auto n_chunks_job = [&](const Ent& x) { return (int)std::max<int64_t>(1, std::min<int64_t>(std::max(1, W / x.mm), x.n_rec / x.r)); };
        // Cost of one job in shared-memory operations per 32 samples (what binds a step: calibrated on the per-phase
        // timeline of the README configuration, profiles/trace_run.py): configuration sums, table fill, segment products
        auto time_of = [&](const Ent& x) {
            const double walk = (double)x.p->lane_cost / x.r;
            const double build = 1.4 * (12.0 * x.p->dslots.size() + (7.0 + 4.0 * S) * (x.p->n_nodes - 1)) + 9.8 * x.p->nSegL + 30.0 * x.p->n_nodes;
            return 300.0 + x.mm * (walk + build);
        };
        auto aux_bytes = [&](const Ent& x) { return (size_t)(x.p->n_nodes + 1) * 32 * x.mm * (2 * sizeof(double) + sizeof(int)); };
        auto aux_off = [&](const Ent& x) {
            const size_t ns = (size_t)32 * x.mm, tb = (size_t)x.slots * ns * opsz, lo = (size_t)x.pd * ns * opsz, ab = aux_bytes(x);
            return (std::max(lo, tb > ab ? tb - ab : (size_t)0) + 15) & ~(size_t)15;
        };
        struct Layout { size_t ok, pw, red, ds, P, D, out, total; };
        auto layout = [&]() {
            size_t table = 0, okb = 0, pwb = 0;
            for (const Ent& x : ev) {
                const size_t ns = (size_t)32 * x.mm;
                table = std::max(table, std::max((size_t)x.slots * ns * opsz, aux_off(x) + aux_bytes(x)));
                okb = std::max(okb, ns * sizeof(int));
                pwb = std::max(pwb, (size_t)std::max(x.p->D, 1) * ns * sizeof(double));
            }
            Layout L;
            L.ok = (table + 15) & ~(size_t)15;
            L.pw = L.ok + okb;
            L.red = L.pw + pwb;
            L.ds = L.red + (size_t)S * W * sizeof(double2);
            L.P = (L.ds + (size_t)max_dslots * sizeof(uint32_t) + 15) & ~(size_t)15;
            L.D = L.P + (((size_t)n_tau * S * opsz + 15) & ~(size_t)15);
            L.out = L.D + ((std::max<size_t>(table_elems, 1) * opsz + 15) & ~(size_t)15);
            L.total = L.out + (size_t)n_ent * S * sizeof(double2) + (size_t)n_ent * sizeof(double) + (size_t)(n_ent + 2) * sizeof(int) + 16;
            return L;
        };
        if (layout().total > cap) return skip("tables exceed shared memory");
        auto total_jobs = [&]() { int64_t t = 0; for (const Ent& x : ev) t += jobs_of(x); return t; };
        // 1. too many jobs: let the cheapest jobs take twice the samples, as far as shared memory allows
        while (total_jobs() > G) {
            int best = -1;
            for (int i = 0; i < n_ent; ++i) {
                Ent& x = ev[i];
                if (x.mm >= 16 || x.mm >= x.n_sb) continue;
                x.mm *= 2;
                const size_t need = layout().total;
                const bool fits = need <= cap;
                x.mm /= 2;
                if (verbose && !fits) fprintf(stderr, "  merge: entry %d (order %d) m %d -> %d would need %zu B\n", pl.ids[i], x.p->order, x.mm, 2 * x.mm, need);
                if (!fits) continue;
                if (best < 0 || time_of(x) < time_of(ev[best])) best = i;
            }
            if (best < 0) break;             // the rest is packed: some CTAs take more than one job per step
            ev[best].mm *= 2;
            if (verbose) fprintf(stderr, "  merge: entry %d (order %d) -> m %d, jobs %lld, shared memory %zu of %zu\n", pl.ids[best], ev[best].p->order, ev[best].mm, (long long)total_jobs(), layout().total, cap);
        }
        if (total_jobs() > 8 * (int64_t)G) return skip("steps large enough to fill the machine: one step kernel per step");
        // 2. CTAs to spare: split the longest jobs
        for (;;) {
            int worst = 0;
            for (int i = 1; i < n_ent; ++i) if (time_of(ev[i]) > time_of(ev[worst])) worst = i;
            Ent& x = ev[worst];
            const int64_t more = (x.n_sb + x.mm - 1) / x.mm;
            if (total_jobs() + more > G || x.n_rec / (x.r + 1) < 2 * std::max(1, W / x.mm)) break;
            x.r += 1;
        }
        // 3. work items, chunks, jobs
        const Layout L = layout();
        std::vector<WorkItem> items;
        std::vector<uint32_t> chunk_off;
        std::vector<LaneRun> runs;
        std::vector<RunJob> jobs;
        std::vector<double> jtime;
        std::vector<int> jent;
        std::vector<int> entry_job0(n_ent + 1, 0);
        for (int i = 0; i < n_ent; ++i) {
            const Ent& x = ev[i];
            entry_job0[i] = (int)jobs.size();
            const int ncj = n_chunks_job(x);
            const int item0 = (int)items.size();
            for (int r = 0; r < x.r; ++r) {
                WorkItem it;
                it.entry = pl.ids[i]; it.slot = i; it.chunk0 = (int)chunk_off.size() + r * ncj; it.n_chunks = ncj;
                it.n_chunks_total = ncj * x.r; it.partial0 = 0;
                items.push_back(it);
            }
            append_entry_chunks(*x.p, ncj * x.r, chunk_off, runs);
            for (int r = 0; r < x.r; ++r)
                for (int64_t sb0 = 0; sb0 < x.n_sb; sb0 += x.mm) {
                    RunJob j;
                    memset(&j, 0, sizeof(j));
                    j.item = item0 + r; j.sb0 = (int)sb0; j.n_sub = x.mm; j.aux_off = (int)aux_off(x); j.row = (int)jobs.size();
                    j.stash_off = -1;
                    jobs.push_back(j);
                    jtime.push_back(time_of(x));
                    jent.push_back(i);
                }
        }
        entry_job0[n_ent] = (int)jobs.size();
        chunk_off.push_back((uint32_t)runs.size());
        if (runs.empty()) runs.push_back(make_uint4(0u, 0u, 0u, 0u));
        // on-chip copy of a job's part of the lane program (records, runs, chunk table, segment definitions and
        // coefficients), placed behind its operand table when the launch's table space leaves room (layout mirrored by
        // scalar_run_kernel's prologue); used when the job is its CTA's only one
        const size_t rows_bytes = jobs.size() * (size_t)S * opsz;
        const bool rows_staged = rows_bytes <= L.ok;
        for (size_t k = 0; k < jobs.size(); ++k) {
            const Ent& x = ev[jent[k]];
            const WorkItem& it = items[jobs[k].item];
            const uint32_t r0 = chunk_off[it.chunk0], r1 = chunk_off[it.chunk0 + it.n_chunks];
            size_t n4 = 8;
            for (uint32_t q = r0; q < r1; ++q) {
                const uint32_t code = runs[q].w, ni = (code >> 4) + ((((code >> 2) & 3u) + 1u) << (code & 3u));
                n4 += (size_t)runs[q].y * ((ni + 7u) / 8u);
            }
            const int NW = x.p->seg_stride > 7 ? 2 : 1;
            const size_t bytes = n4 * 16 + (size_t)(r1 - r0) * 16 + (((size_t)(it.n_chunks + 1) * 4 + 15) & ~(size_t)15) +
                                 (size_t)x.p->nSegL * NW * 16 + (((size_t)x.p->nSegL * opsz + 15) & ~(size_t)15);
            const size_t ns = (size_t)32 * x.mm;
            const size_t region_end = std::max((size_t)x.slots * ns * opsz, aux_off(x) + aux_bytes(x));
            const size_t off = (std::max(region_end, rows_staged ? rows_bytes : (size_t)0) + 15) & ~(size_t)15;
            if (off + bytes <= L.ok) jobs[k].stash_off = (int)off;
            // may a part of the CTA prepare the next step's times / pair-interaction rows while the rest reduces the
            // partial rows staged at the start of the table?  Only if the rows stay inside the propagator rows.
            jobs[k].flags = (!rows_staged || (rows_bytes <= (size_t)x.p->nP * ns * opsz && rows_bytes <= aux_off(x))) ? 1 : 0;
        }
        // Jobs -> SMs -> CTAs at about equal estimated cost (longest job first onto the least loaded bin): a step is bound
        // by the shared-memory pipe of the busiest SM, so the unit of balance is the SM; the CTAs that the hardware places
        // on one SM claim the job lists of one bin at run time (scalar_run_kernel).  With fewer jobs than CTA slots the
        // lists are simply one per CTA.
        const int max_ctas = scalar_run_max_ctas(real != 0, pl.lane_dual, threads, L.total, n_sm);
        if (max_ctas < n_sm) return skip("occupancy below one CTA per SM");
        const bool by_sm = (int)jobs.size() > n_sm && max_ctas >= G && !getenv("QIW_RUN_NO_SM_BINS");
        const int n_ctas = by_sm ? G : (int)std::min<size_t>(jobs.size(), (size_t)std::min(G, max_ctas));
        std::vector<int> ord(jobs.size());
        for (size_t k = 0; k < ord.size(); ++k) ord[k] = (int)k;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return jtime[a] > jtime[b]; });
        std::vector<std::vector<int>> lists(n_ctas);
        std::vector<double> load(n_ctas, 0.0);
        if (by_sm) {
            std::vector<std::vector<int>> bins(n_sm);
            std::vector<double> bload(n_sm, 0.0);
            for (int j : ord) {
                const int b = (int)(std::min_element(bload.begin(), bload.end()) - bload.begin());
                bins[b].push_back(j); bload[b] += jtime[j];
            }
            for (int b = 0; b < n_sm; ++b)            // inside a bin: longest first onto the least loaded of its CTAs
                for (int j : bins[b]) {
                    int best = b * ctas_per_sm;
                    for (int q = 1; q < ctas_per_sm; ++q) if (load[b * ctas_per_sm + q] < load[best]) best = b * ctas_per_sm + q;
                    lists[best].push_back(j); load[best] += jtime[j];
                }
        } else {
            for (int j : ord) {
                const int b = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                lists[b].push_back(j); load[b] += jtime[j];
            }
        }
        std::vector<RunJob> placed;
        std::vector<int> cta_job0(n_ctas + 1, 0);
        for (int k = 0; k < n_ctas; ++k) {
            cta_job0[k] = (int)placed.size();
            for (int j : lists[k]) placed.push_back(jobs[j]);
        }
        cta_job0[n_ctas] = (int)placed.size();
        rn.by_sm = by_sm; rn.ctas_per_sm = ctas_per_sm;
        CK(rn.d_cta_job0.upload(cta_job0.data(), cta_job0.size(), ctx->stream));
        CK(rn.d_jobs.upload(placed.data(), placed.size(), ctx->stream));
        CK(rn.d_entry_job0.upload(entry_job0.data(), entry_job0.size(), ctx->stream));
        CK(rn.d_items.upload(items.data(), items.size(), ctx->stream));
        CK(rn.d_chunk_off.upload(chunk_off.data(), chunk_off.size(), ctx->stream));
        CK(rn.d_runs.upload(runs.data(), runs.size(), ctx->stream));
        CK(rn.d_partials.reserve(2 * placed.size() * (size_t)S));
        CK(rn.d_barrier.reserve(2052));     // [0, 2049): SM map of the run kernel, [2049]: grid barrier counter
        CK(cudaStreamSynchronize(ctx->stream));
        rn.n_jobs = (int)placed.size(); rn.n_ctas = n_ctas; rn.threads = threads; rn.smem = L.total;
        rn.ok_off = (int)L.ok; rn.pw_off = (int)L.pw; rn.red_off = (int)L.red; rn.ds_off = (int)L.ds; rn.P_off = (int)L.P;
        rn.D_off = (int)L.D; rn.out_off = (int)L.out; rn.rows_staged = rows_staged ? 1 : 0;
        rn.state = 1;
        if (getenv("QIW_RUN_VERBOSE")) {
            fprintf(stderr, "qiw run kernel: %d jobs on %d CTAs (max %d), %zu B shared memory, %s arithmetic, CTA load %.0f .. %.0f\n", rn.n_jobs, n_ctas, max_ctas, rn.smem, real ? "real" : "complex", *std::min_element(load.begin(), load.end()), *std::max_element(load.begin(), load.end()));
            for (int i = 0; i < n_ent; ++i)
                fprintf(stderr, "  entry %d order %d k %d: cost %lld slots %d -> m %d, split %d, %lld jobs of %d chunks, est %.0f\n", pl.ids[i], ev[i].p->order,
                        ev[i].p->n_pts_after, (long long)ev[i].p->lane_cost, ev[i].slots, ev[i].mm, ev[i].r, (long long)jobs_of(ev[i]), n_chunks_job(ev[i]), time_of(ev[i]));
        }
    }
    RunParams rp;
    memset(&rp, 0, sizeof(rp));
    StepParams& sp = rp.sp;
    sp.entries = ctx->dEntries.p; sp.dyn = pl.d_dyn.p; sp.items = rn.d_items.p; sp.chunk_off = rn.d_chunk_off.p; sp.runs = rn.d_runs.p;
    sp.P = ctx->dP.p; sp.E = ctx->dE.p; sp.deltas = ctx->dDeltas.p;
    for (int t = 0; t < kInlineTables && t < (int)ctx->hDeltas.size(); ++t) sp.deltas_inline[t] = ctx->hDeltas[t];
    sp.tables_on_grid = on_grid ? 1 : 0;
    {
        size_t off = 0;
        for (size_t t = 0; t < ctx->tables.size() && t < (size_t)kInlineTables; ++t) {
            rp.D_table_off[t] = (int)off;
            off += (size_t)ctx->tables[t].n * (ctx->tables[t].kind == 1 ? 2 : 1);
        }
    }
    sp.S = S; sp.bsize = m.bsize; sp.n_tau = n_tau; sp.h = ctx->beta / (n_tau - 1); sp.inv_h = 1.0 / sp.h;
    sp.n_call_entries = n_ent;
    sp.finish_P = ctx->dP.p; sp.finish_diag = diag; sp.finish_n_diag = n_diag;
    if (collective && ctx->peer_ready && ctx->n_ranks > 1) {
        sp.peer_ranks = ctx->n_ranks; sp.peer_rank = ctx->rank; sp.peer_seq = ctx->peer_seq + 1;
        ctx->peer_seq += (unsigned long long)n_steps;
        sp.peer_mail = ctx->dPeerPtrs.p; sp.peer_status = ctx->dPeerStatus.p; sp.peer_timeout_ns = peer_timeout_ns();
    }
    rp.jobs = rn.d_jobs.p; rp.cta_job0 = rn.d_cta_job0.p; rp.entry_job0 = rn.d_entry_job0.p; rp.n_jobs = rn.n_jobs;
    rp.k_first = k_first; rp.n_steps = n_steps;
    rp.partials = rn.d_partials.p; rp.barrier = rn.d_barrier.p + 2049;
    rp.sm_map = rn.by_sm ? rn.d_barrier.p : nullptr; rp.ctas_per_sm = rn.ctas_per_sm;
    rp.post_warps = std::max(1, rn.threads / 64);     // half of the CTA (measured on the README run: 2 / 3 / 4 / 6 / 8 / 12 of 12 warps: 3.99 / 3.78 / 3.66 / 3.57 / 3.74 / 3.91 ms)
    if (const char* env = getenv("QIW_RUN_POST_WARPS")) { const int v = atoi(env); if (v >= 1 && v <= 32) rp.post_warps = v; }
    rp.n_tables = (int)ctx->tables.size();
    rp.ok_off = rn.ok_off; rp.pw_off = rn.pw_off; rp.red_off = rn.red_off; rp.ds_off = rn.ds_off; rp.P_off = rn.P_off;
    rp.D_off = rn.D_off; rp.out_off = rn.out_off; rp.rows_staged = rn.rows_staged;
    rp.hist = hist; rp.hist_stride = hist_stride; rp.hist_off = hist_off;
    CK(cudaMemsetAsync(rn.d_barrier.p, 0, 2052 * sizeof(unsigned int), ctx->stream));
    const char* trace_path = getenv("QIW_TRACE");   // diagnostics (make trace): per-CTA timeline of the middle step
    if (trace_path) {
        CK(ctx->dTrace.reserve((size_t)rn.n_ctas * 16));
        CK(cudaMemsetAsync(ctx->dTrace.p, 0, (size_t)rn.n_ctas * 16 * sizeof(unsigned long long), ctx->stream));
        sp.trace = ctx->dTrace.p;
    }
    ctx->last_real_mode = real;
    {
        ProfScope ps(ctx, 2);
        CK(launch_scalar_run(real != 0, pl.lane_dual, rp, rn.n_ctas, rn.threads, rn.smem, ctx->stream));
    }
    ctx->launches++;
    if (trace_path) {
        std::vector<unsigned long long> tr((size_t)rn.n_ctas * 16);
        CK(cudaMemcpyAsync(tr.data(), ctx->dTrace.p, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (FILE* f = fopen(trace_path, "w")) {
            fprintf(f, "cta,t0,times,fill,seg,walk,jobs_done,barrier,reduce,update,smid,n_jobs,entry,n_sub,end_ns\n");
            for (int c2 = 0; c2 < rn.n_ctas; ++c2) {
                const unsigned long long* t = tr.data() + (size_t)c2 * 16;
                fprintf(f, "%d,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu\n", c2, t[0], t[1] - t[0], t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0],
                        t[6] - t[0], t[7] - t[0], t[8] - t[0], t[9], t[10], t[11], t[12], t[13]);
            }
            fclose(f);
        }
    }
    *used = true;
    return QIW_OK;
}

// After the first evaluation of a plan the simplex roots of its sequence are in the cache.
static int mark_ucache_valid(qiw_context* ctx, Plan& pl) {
    if (!pl.ucache_enabled || pl.ucache_valid) return QIW_OK;
    pl.ucache_valid = true;
    for (auto& dy : pl.h_dyn) dy.ucache_valid = dy.ucache ? 1 : 0;
    CK(cudaMemcpyAsync(pl.d_dyn.p, pl.h_dyn.data(), pl.h_dyn.size() * sizeof(DevEntryDyn), cudaMemcpyHostToDevice, ctx->stream));
    return QIW_OK;
}

static int nccl_allreduce(qiw_context* ctx, double2* buf, size_t n_complex) {
    if (!ctx->comm) {
        if (ctx->n_ranks > 1) return fail(ctx, QIW_ERR_NCCL, "this call needs the NCCL communicator (qiw_comm_init) in addition to the peer mailboxes");
        return QIW_OK;
    }
    ProfScope ps(ctx, 6);
    int nrc = g_nccl.AllReduce(buf, buf, n_complex * 2, kNcclDouble, kNcclSum, ctx->comm, ctx->stream);
    if (nrc) return fail(ctx, QIW_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error"));
    return QIW_OK;
}

// After a synchronised call that used the peer exchange: did every peer answer?
static int check_peer_status(qiw_context* ctx) {
    if (!ctx->peer_ready || ctx->n_ranks <= 1) return QIW_OK;
    int st = 0;
    CK(cudaMemcpy(&st, ctx->dPeerStatus.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (st) {
        cudaMemset(ctx->dPeerStatus.p, 0, sizeof(int));
        return fail(ctx, QIW_ERR_NCCL, "peer-memory all-reduce timed out: a rank of the job did not reach the collective");
    }
    return QIW_OK;
}

static int ensure_host_out_impl(qiw_context* ctx, size_t n) {
    if (ctx->hOutCap < n) {
        if (ctx->hOut) cudaFreeHost(ctx->hOut);
        ctx->hOut = nullptr; ctx->hOutCap = 0;
        CK(cudaMallocHost((void**)&ctx->hOut, n * sizeof(double2)));
        ctx->hOutCap = n;
    }
    return QIW_OK;
}

// Common body of qiw_eval / qiw_eval_range / qiw_eval_at_times (scalar models).
static int eval_scalar(qiw_context* ctx, double t_i, double t_w, double t_f, int n_entries, const int32_t* ids,
                       const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t start, uint64_t count,
                       uint64_t N_total, bool allreduce, double* out, const double* explicit_times, int n_explicit) {
    const HostModel& m = ctx->model;
    const bool explicit_mode = explicit_times != nullptr;
    int rc = sync_static_tables(ctx);
    if (rc) return rc;
    Plan* plp = nullptr;
    rc = get_plan(ctx, n_entries, ids, count, explicit_mode, &plp);
    if (rc) return rc;
    Plan& pl = *plp;
    rc = stage_call(ctx, pl, sobol_m, sobol_x0, start, count, N_total, allreduce);
    if (rc) return rc;
    if (explicit_mode) {
        const int D = ctx->entries[ids[0]]->prog.D;
        CK(ctx->dTimes.upload(explicit_times, (size_t)n_explicit * std::max(D, 1), ctx->stream));
        CK(ctx->dPerSample.reserve((size_t)n_explicit * m.bsize));
        CK(cudaMemsetAsync(ctx->dPerSample.p, 0, (size_t)n_explicit * m.bsize * sizeof(double2), ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    bool coll_done = false;
    rc = enqueue_step(ctx, pl, t_i, t_w, t_f, nullptr, nullptr, allreduce && !explicit_mode, &coll_done);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = mark_ucache_valid(ctx, pl);
    if (rc) return rc;
    const size_t n_out = explicit_mode ? (size_t)n_explicit * m.bsize : (size_t)n_entries * m.bsize;
    double2* src = explicit_mode ? ctx->dPerSample.p : pl.d_out.p;
    if (allreduce && !explicit_mode && !coll_done) { rc = nccl_allreduce(ctx, src, n_out); if (rc) return rc; }
    rc = ensure_host_out(ctx, n_out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->hOut, src, n_out * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    memcpy(out, ctx->hOut, n_out * sizeof(double2));
    if (coll_done) return check_peer_status(ctx);
    return QIW_OK;
}

static int check_entries(qiw_context* ctx, int n_entries, const int32_t* ids, const char* who) {
    if (!ctx) return QIW_ERR_BAD_ARG;
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, std::string(who) + ": planning-only context (QIW_DEVICE_NONE) cannot compute; there is no CPU path");
    if (!ctx->have_model || ctx->n_tau == 0) return fail(ctx, QIW_ERR_BAD_ARG, std::string(who) + ": model and grid must be set first");
    if (n_entries <= 0 || !ids) return fail(ctx, QIW_ERR_BAD_ARG, std::string(who) + ": no entries");
    for (int i = 0; i < n_entries; ++i) {
        if (ids[i] < 0 || ids[i] >= (int)ctx->entries.size() || !ctx->entries[ids[i]] || !ctx->entries[ids[i]]->valid)
            return fail(ctx, QIW_ERR_BAD_ARG, std::string(who) + ": unknown entry id");
        for (int j = 0; j < i; ++j) if (ids[j] == ids[i]) return fail(ctx, QIW_ERR_BAD_ARG, std::string(who) + ": duplicate entry id");
    }
    return QIW_OK;
}

extern "C" {

int qiw_eval_range(qiw_context* ctx, double t_i, double t_w, double t_f, int32_t n_entries, const int32_t* ids,
                   const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t start, uint64_t count,
                   uint64_t N_total, double* out) {
    int rc = check_entries(ctx, n_entries, ids, "qiw_eval_range");
    if (rc) return rc;
    if (!out || N_total == 0 || start + count > N_total || N_total > 0xFFFFFFFFull) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_eval_range: bad sample range");
    cudaSetDevice(ctx->device);
    return eval_scalar(ctx, t_i, t_w, t_f, n_entries, ids, sobol_m, sobol_x0, start, count, N_total, false, out, nullptr, 0);
}

int qiw_eval(qiw_context* ctx, double t_i, double t_w, double t_f, int32_t n_entries, const int32_t* ids,
             const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t N_total, double* out) {
    int rc = check_entries(ctx, n_entries, ids, "qiw_eval");
    if (rc) return rc;
    if (!out || N_total == 0 || N_total > 0xFFFFFFFFull) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_eval: bad N_total");
    cudaSetDevice(ctx->device);
    uint64_t start = 0, count = N_total;
    rank_sub_range(N_total, ctx->n_ranks, ctx->rank, &start, &count);
    return eval_scalar(ctx, t_i, t_w, t_f, n_entries, ids, sobol_m, sobol_x0, start, count, N_total, true, out, nullptr, 0);
}

int qiw_eval_batch(qiw_context* ctx, int32_t n_times, const double* times, int32_t n_entries, const int32_t* ids,
                   const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t N_total, double* out) {
    int rc = check_entries(ctx, n_entries, ids, "qiw_eval_batch");
    if (rc) return rc;
    if (!out || !times || n_times <= 0 || n_times > 65535 || N_total == 0 || N_total > 0xFFFFFFFFull)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_eval_batch: bad argument");
    cudaSetDevice(ctx->device);
    const HostModel& m = ctx->model;
    uint64_t start = 0, count = N_total;
    rank_sub_range(N_total, ctx->n_ranks, ctx->rank, &start, &count);
    rc = sync_static_tables(ctx);
    if (rc) return rc;
    Plan* plp = nullptr;
    rc = get_plan(ctx, n_entries, ids, count, false, &plp);
    if (rc) return rc;
    Plan& pl = *plp;
    if (!m.scalar && !pl.block_real && !pl.block_mma) {   // block models in complex arithmetic: one launch per triple (the general block kernel has no batched form)
        for (int z = 0; z < n_times; ++z) {
            rc = qiw_eval(ctx, times[3 * z], times[3 * z + 1], times[3 * z + 2], n_entries, ids, sobol_m, sobol_x0, N_total,
                          out + (size_t)z * n_entries * m.bsize * 2);
            if (rc) return rc;
        }
        return QIW_OK;
    }
    rc = stage_call(ctx, pl, sobol_m, sobol_x0, start, count, N_total, true);
    if (rc) return rc;
    const size_t n_out = (size_t)n_times * n_entries * m.bsize;
    CK(ctx->dTimes3.upload(times, (size_t)n_times * 3, ctx->stream));
    CK(ctx->dBatchPartials.reserve((size_t)n_times * pl.partial_rows * m.bsize));
    CK(ctx->dBatchOut.reserve(n_out));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = enqueue_step(ctx, pl, 0.0, 0.0, 0.0, nullptr, nullptr, false, nullptr, n_times);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = mark_ucache_valid(ctx, pl);
    if (rc) return rc;
    rc = nccl_allreduce(ctx, ctx->dBatchOut.p, n_out);   // one collective for the whole batch
    if (rc) return rc;
    rc = ensure_host_out(ctx, n_out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->hOut, ctx->dBatchOut.p, n_out * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    memcpy(out, ctx->hOut, n_out * sizeof(double2));
    return QIW_OK;
}

int qiw_eval_seqs(qiw_context* ctx, double t_i, double t_w, double t_f, int32_t n_seqs, int32_t n_entries, const int32_t* ids,
                  const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t N_total, double* out) {
    int rc = check_entries(ctx, n_entries, ids, "qiw_eval_seqs");
    if (rc) return rc;
    if (!out || !sobol_m || !sobol_x0 || n_seqs <= 0 || n_seqs > 65535 || N_total == 0 || N_total > 0xFFFFFFFFull)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_eval_seqs: bad argument");
    cudaSetDevice(ctx->device);
    const HostModel& m = ctx->model;
    size_t md = 0, xd = 0;     // words per sequence in the caller's arrays
    for (int i = 0; i < n_entries; ++i) { md += (size_t)ctx->entries[ids[i]]->prog.D * 32; xd += ctx->entries[ids[i]]->prog.D; }
    uint64_t start = 0, count = N_total;
    rank_sub_range(N_total, ctx->n_ranks, ctx->rank, &start, &count);
    rc = sync_static_tables(ctx);
    if (rc) return rc;
    Plan* plp = nullptr;
    rc = get_plan(ctx, n_entries, ids, count, false, &plp);
    if (rc) return rc;
    Plan& pl = *plp;
    if (!m.scalar && !pl.block_real && !pl.block_mma) {   // block models in complex arithmetic: one launch per sequence
        for (int z = 0; z < n_seqs; ++z) {
            rc = qiw_eval(ctx, t_i, t_w, t_f, n_entries, ids, sobol_m + (size_t)z * md, sobol_x0 + (size_t)z * xd, N_total,
                          out + (size_t)z * n_entries * m.bsize * 2);
            if (rc) return rc;
        }
        return QIW_OK;
    }
    rc = stage_call(ctx, pl, sobol_m, sobol_x0, start, count, N_total, true);   // weights, ranges (sequence 0's parameters)
    if (rc) return rc;
    // per-sequence parameter blocks in the plan's per-entry layout (m[D][32] then x0[D] per entry)
    const size_t tot = pl.h_sobol.size();
    std::vector<uint32_t> hs(tot * (size_t)n_seqs, 0u);
    for (int z = 0; z < n_seqs; ++z) {
        size_t moff = 0, xoff = 0;
        for (int i = 0; i < n_entries; ++i) {
            const int D = ctx->entries[ids[i]]->prog.D;
            uint32_t* sb = hs.data() + (size_t)z * tot + pl.sobol_off[i];
            memcpy(sb, sobol_m + (size_t)z * md + moff, (size_t)D * 32 * sizeof(uint32_t));
            memcpy(sb + (size_t)D * 32, sobol_x0 + (size_t)z * xd + xoff, (size_t)D * sizeof(uint32_t));
            moff += (size_t)D * 32; xoff += D;
        }
    }
    CK(ctx->dSeqSobol.upload(hs.data(), hs.size(), ctx->stream));
    std::vector<DevEntryDyn> dyn = pl.h_dyn;
    for (int i = 0; i < n_entries; ++i) { dyn[i].sobol = ctx->dSeqSobol.p + pl.sobol_off[i]; dyn[i].ucache = nullptr; dyn[i].ucache_valid = 0; }
    CK(ctx->dSeqDyn.upload(dyn.data(), dyn.size(), ctx->stream));
    pl.ucache_valid = false;   // stage_call recorded sequence 0 as the plan's sequence: its cached roots are not trustworthy
    pl.default_sobol_resident = false;
    std::vector<double> times((size_t)n_seqs * 3);
    for (int z = 0; z < n_seqs; ++z) { times[3 * z] = t_i; times[3 * z + 1] = t_w; times[3 * z + 2] = t_f; }
    const size_t n_out = (size_t)n_seqs * n_entries * m.bsize;
    CK(ctx->dTimes3.upload(times.data(), times.size(), ctx->stream));
    CK(ctx->dBatchPartials.reserve((size_t)n_seqs * pl.partial_rows * m.bsize));
    CK(ctx->dBatchOut.reserve(n_out));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = enqueue_step(ctx, pl, t_i, t_w, t_f, nullptr, nullptr, false, nullptr, n_seqs, ctx->dSeqSobol.p, (int)tot);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = nccl_allreduce(ctx, ctx->dBatchOut.p, n_out);
    if (rc) return rc;
    rc = ensure_host_out(ctx, n_out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->hOut, ctx->dBatchOut.p, n_out * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    memcpy(out, ctx->hOut, n_out * sizeof(double2));
    return QIW_OK;
}

int qiw_eval_at_times(qiw_context* ctx, int32_t entry_id, double t_i, double t_w, double t_f, int32_t n_samples,
                      const double* times, double* out) {
    int rc = check_entries(ctx, 1, &entry_id, "qiw_eval_at_times");
    if (rc) return rc;
    if (n_samples <= 0 || !times || !out) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_eval_at_times: bad argument");
    cudaSetDevice(ctx->device);
    return eval_scalar(ctx, t_i, t_w, t_f, 1, &entry_id, nullptr, nullptr, 0, (uint64_t)n_samples, (uint64_t)n_samples, false,
                       out, times, n_samples);
}

int qiw_last_device_ms(qiw_context* ctx, double* ms) { if (!ctx || !ms) return QIW_ERR_BAD_ARG; *ms = ctx->last_ms; return QIW_OK; }
int qiw_launch_count(qiw_context* ctx, int64_t* n) { if (!ctx || !n) return QIW_ERR_BAD_ARG; *n = ctx->launches; return QIW_OK; }

// inchworm!'s loop (src/inchworm.jl:400-493) with the per-step state update on the device.
int qiw_inchworm_run(qiw_context* ctx, int32_t n_bare, const int32_t* bare_ids, int32_t n_bold, const int32_t* bold_ids,
                     const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t N_total, double* order_contribs) {
    int rc = check_entries(ctx, n_bare, bare_ids, "qiw_inchworm_run");
    if (rc) return rc;
    if (n_bold > 0) { rc = check_entries(ctx, n_bold, bold_ids, "qiw_inchworm_run"); if (rc) return rc; }
    if (N_total == 0 || N_total > 0xFFFFFFFFull) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_inchworm_run: bad N_total");
    for (int i = 0; i < n_bare; ++i) if (ctx->entries[bare_ids[i]]->prog.mode != QIW_MODE_BARE) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_inchworm_run: bare_ids must hold QIW_MODE_BARE entries");
    for (int i = 0; i < n_bold; ++i) if (ctx->entries[bold_ids[i]]->prog.mode != QIW_MODE_BOLD) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_inchworm_run: bold_ids must hold QIW_MODE_BOLD entries");
    cudaSetDevice(ctx->device);
    const HostModel& m = ctx->model;
    const int n_tau = ctx->n_tau, bs = m.bsize;
    const double h = ctx->beta / (n_tau - 1);
    rc = sync_static_tables(ctx);
    if (rc) return rc;
    uint64_t start = 0, count = N_total;
    rank_sub_range(N_total, ctx->n_ranks, ctx->rank, &start, &count);
    Plan *pb = nullptr, *pd = nullptr;
    rc = get_plan(ctx, n_bare, bare_ids, count, false, &pb);
    if (rc) return rc;
    if (n_bold > 0) {
        rc = get_plan(ctx, n_bold, bold_ids, count, false, &pd);
        if (rc) return rc;
        rc = get_plan(ctx, n_bare, bare_ids, count, false, &pb);   // re-fetch: building pd may have evicted it
        if (rc) return rc;
    }
    size_t moff = 0, xoff = 0;
    for (int i = 0; i < n_bare; ++i) { moff += (size_t)ctx->entries[bare_ids[i]]->prog.D * 32; xoff += ctx->entries[bare_ids[i]]->prog.D; }
    rc = stage_call(ctx, *pb, sobol_m, sobol_x0, start, count, N_total, true);
    if (rc) return rc;
    if (n_bold > 0) {
        rc = stage_call(ctx, *pd, sobol_m ? sobol_m + moff : nullptr, sobol_x0 ? sobol_x0 + xoff : nullptr, start, count, N_total, true);
        if (rc) return rc;
    }
    const int n_hist = n_bare + n_bold;
    if (order_contribs) {
        CK(ctx->dHist.reserve((size_t)n_tau * n_hist * bs));
        CK(cudaMemsetAsync(ctx->dHist.p, 0, (size_t)n_tau * n_hist * bs * sizeof(double2), ctx->stream));
    }
    // diagonal element indices of the packed block vector
    std::vector<int> diag;
    for (int s = 0; s < m.S; ++s) for (int i = 0; i < m.dim[s]; ++i) diag.push_back(m.boff[s] + i + m.dim[s] * i);
    CK(ctx->dDiag.upload(diag.data(), diag.size(), ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    // bare step: grid[0] -> grid[1], no normalisation (src/inchworm.jl:400-416)
    FinishArgs fin;
    bool coll_done = false;
    fin.diag = ctx->dDiag.p; fin.n_diag = (int)diag.size();
    bool fused = false;
    fin.k_f = 1; fin.normalize = 0; fin.hist = order_contribs ? ctx->dHist.p + (size_t)1 * n_hist * bs : nullptr;
    rc = enqueue_step(ctx, *pb, 0.0, 0.0, h, &fin, &fused, true, &coll_done);
    if (rc) return rc;
    if (!coll_done) { rc = nccl_allreduce(ctx, pb->d_out.p, (size_t)n_bare * bs); if (rc) return rc; }
    if (!fused) {
        ProfScope ps(ctx, 5);
        CK(launch_finish_step(ctx->dP.p, n_tau, bs, ctx->dDiag.p, (int)diag.size(), h, 1, pb->d_out.p, n_bare, 0,
                              order_contribs ? ctx->dHist.p + (size_t)1 * n_hist * bs : nullptr, ctx->stream));
        ctx->launches++;
    }
    // bold steps n = 2 .. n_tau-1 (1-based): tau_w = grid[n], tau_f = grid[n+1] (:474-493): one persistent launch when
    // the steps are small, else one step kernel per step
    bool run_used = false;
    if (n_bold > 0 && n_tau > 2) {
        rc = enqueue_run(ctx, *pd, 1, n_tau - 2, order_contribs ? ctx->dHist.p : nullptr, (size_t)n_hist * bs, (size_t)n_bare * bs,
                         ctx->dDiag.p, (int)diag.size(), true, &run_used);
        if (rc) return rc;
        if (run_used && ctx->n_ranks > 1) coll_done = true;
        if (run_used) { rc = mark_ucache_valid(ctx, *pd); if (rc) return rc; }   // the run kernel's prologue filled the root cache
    }
    for (int n = 1; !run_used && n_bold > 0 && n < n_tau - 1; ++n) {
        fin.k_f = n + 1; fin.normalize = 1;
        fin.hist = order_contribs ? ctx->dHist.p + ((size_t)(n + 1) * n_hist + n_bare) * bs : nullptr;
        rc = enqueue_step(ctx, *pd, 0.0, n * h, (n + 1) * h, &fin, &fused, true, &coll_done);
        if (rc) return rc;
        rc = mark_ucache_valid(ctx, *pd);
        if (rc) return rc;
        if (!coll_done) { rc = nccl_allreduce(ctx, pd->d_out.p, (size_t)n_bold * bs); if (rc) return rc; }
        if (!fused) {
            ProfScope ps(ctx, 5);
            CK(launch_finish_step(ctx->dP.p, n_tau, bs, ctx->dDiag.p, (int)diag.size(), h, n + 1, pd->d_out.p, n_bold, 1,
                                  order_contribs ? ctx->dHist.p + ((size_t)(n + 1) * n_hist + n_bare) * bs : nullptr, ctx->stream));
            ctx->launches++;
        }
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    const size_t n_hist_el = (size_t)n_tau * n_hist * bs;
    if (order_contribs) {      // through the pinned staging buffer: a copy into pageable memory is staged by the driver in small pieces
        rc = ensure_host_out(ctx, n_hist_el);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->hOut, ctx->dHist.p, n_hist_el * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (order_contribs) memcpy(order_contribs, ctx->hOut, n_hist_el * sizeof(double2));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    // rows written in complex arithmetic may carry a real part: the next qiw_set_P re-examines them
    if (!ctx->last_real_mode) for (int k = 1; k < n_tau; ++k) ctx->p_row_complex[k] = 1;
    if (coll_done) return check_peer_status(ctx);
    return QIW_OK;
}

// ---- Sobol / topologies / partitioning -----------------------------------------------------------

int qiw_sobol_direction_numbers(int32_t D, uint32_t* m) { return (m && sobol_direction_numbers(D, m) == 0) ? QIW_OK : QIW_ERR_BAD_ARG; }

int qiw_sobol_scramble(int32_t D, uint32_t* m, uint32_t* x0, const uint8_t* shift_bits, const uint8_t* ltm_bits) {
    if (D < 0 || !m || !x0 || !shift_bits || !ltm_bits) return QIW_ERR_BAD_ARG;
    return sobol_scramble(D, m, x0, shift_bits, ltm_bits) == 0 ? QIW_OK : QIW_ERR_BAD_ARG;
}

int qiw_sobol_points(qiw_context* ctx, int32_t D, const uint32_t* m, const uint32_t* x0, uint64_t start, uint64_t count,
                     uint32_t* out) {
    if (!ctx || D <= 0 || !m || !out || start + count > 0x100000000ull) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_sobol_points: bad argument");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context cannot compute");
    cudaSetDevice(ctx->device);
    DevBuf<uint32_t> dm, dx, dout;
    std::vector<uint32_t> zeros(D, 0u);
    CK(dm.upload(m, (size_t)D * 32, ctx->stream));
    CK(dx.upload(x0 ? x0 : zeros.data(), D, ctx->stream));
    CK(dout.reserve(count * D));
    if (count) {
        CK(launch_sobol_points(D, dm.p, dx.p, start, count, dout.p, ctx->stream));
        ctx->launches++;
        CK(cudaMemcpyAsync(out, dout.p, count * D * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    dm.release(); dx.release(); dout.release();
    return QIW_OK;
}

int qiw_topologies(int32_t order, int32_t k, int32_t with_external_arc, int64_t* n_top, int32_t* pairs, int32_t* parity) {
    if (!n_top) return QIW_ERR_BAD_ARG;
    int64_t n = enumerate_topologies(order, k, with_external_arc != 0, pairs, parity);
    if (n < 0) return QIW_ERR_BAD_ARG;
    *n_top = n;
    return QIW_OK;
}

int qiw_rank_sub_range(uint64_t N, int32_t n_ranks, int32_t rank, uint64_t* start, uint64_t* count) {
    if (n_ranks <= 0 || rank < 0 || rank >= n_ranks || !start || !count) return QIW_ERR_BAD_ARG;
    rank_sub_range(N, n_ranks, rank, start, count);
    return QIW_OK;
}

// ---- NCCL ------------------------------------------------------------------------------------------

int qiw_comm_unique_id(uint8_t id[QIW_UNIQUE_ID_BYTES]) {
    if (!id || !g_nccl.load()) return QIW_ERR_NCCL;
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u)) return QIW_ERR_NCCL;
    memcpy(id, u.internal, QIW_UNIQUE_ID_BYTES);
    return QIW_OK;
}

int qiw_comm_init(qiw_context* ctx, int32_t n_ranks, int32_t rank, const uint8_t id[QIW_UNIQUE_ID_BYTES]) {
    if (!ctx || n_ranks <= 0 || rank < 0 || rank >= n_ranks || !id) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_comm_init: bad argument");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context cannot communicate");
    if (!g_nccl.load()) return fail(ctx, QIW_ERR_NCCL, g_nccl.err);
    cudaSetDevice(ctx->device);
    ncclUniqueId u;
    memcpy(u.internal, id, QIW_UNIQUE_ID_BYTES);
    int rc = g_nccl.CommInitRank(&ctx->comm, n_ranks, u, rank);
    if (rc) return fail(ctx, QIW_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    ctx->n_ranks = n_ranks; ctx->rank = rank;
    drop_plans(ctx);
    return QIW_OK;
}

int qiw_comm_destroy(qiw_context* ctx) {
    if (!ctx) return QIW_ERR_BAD_ARG;
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr; ctx->n_ranks = 1; ctx->rank = 0;
    drop_plans(ctx);
    return QIW_OK;
}

int qiw_peer_handle(qiw_context* ctx, uint8_t handle[QIW_PEER_HANDLE_BYTES]) {
    if (!ctx || !handle) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_peer_handle: bad argument");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context cannot communicate");
    static_assert(sizeof(cudaIpcMemHandle_t) == QIW_PEER_HANDLE_BYTES, "IPC handle size");
    cudaSetDevice(ctx->device);
    if (!ctx->peerLocal) {
        CK(cudaMalloc((void**)&ctx->peerLocal, kPeerMailBytes));
        CK(cudaMemset(ctx->peerLocal, 0, kPeerMailBytes));
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->peerLocal));
    memcpy(handle, &h, QIW_PEER_HANDLE_BYTES);
    return QIW_OK;
}

int qiw_peer_init(qiw_context* ctx, int32_t n_ranks, int32_t rank, const uint8_t* handles) {
    if (ctx && n_ranks == 0) {   // switch the peer path off again (e.g. another rank could not map the mailboxes)
        ctx->peer_ready = false;
        return QIW_OK;
    }
    if (!ctx || n_ranks <= 0 || n_ranks > kMaxPeers || rank < 0 || rank >= n_ranks || !handles)
        return fail(ctx, QIW_ERR_BAD_ARG, "qiw_peer_init: bad argument (at most 16 ranks)");
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context cannot communicate");
    if (!ctx->peerLocal) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_peer_init: call qiw_peer_handle first");
    if (ctx->comm && (ctx->n_ranks != n_ranks || ctx->rank != rank)) return fail(ctx, QIW_ERR_BAD_ARG, "qiw_peer_init: rank layout differs from qiw_comm_init");
    cudaSetDevice(ctx->device);
    ctx->peerPtrs.assign(n_ranks, nullptr);
    for (int q = 0; q < n_ranks; ++q) {
        if (q == rank) { ctx->peerPtrs[q] = ctx->peerLocal; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * QIW_PEER_HANDLE_BYTES, QIW_PEER_HANDLE_BYTES);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->peerPtrs.clear();
            return fail(ctx, QIW_ERR_CUDA, std::string("qiw_peer_init: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) +
                                           " (no peer access between the GPUs? the NCCL path of qiw_comm_init remains usable)");
        }
        ctx->peerPtrs[q] = (unsigned char*)ptr;
    }
    CK(ctx->dPeerPtrs.upload(ctx->peerPtrs.data(), ctx->peerPtrs.size(), ctx->stream));
    CK(ctx->dPeerStatus.reserve(1));
    CK(cudaMemsetAsync(ctx->dPeerStatus.p, 0, sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_ranks = n_ranks; ctx->rank = rank;
    ctx->peer_seq = 0;
    ctx->peer_ready = true;
    drop_plans(ctx);
    return QIW_OK;
}

int qiw_profile_enable(qiw_context* ctx, int32_t on) {
    if (!ctx) return QIW_ERR_BAD_ARG;
    ctx->profiling = (on != 0) && !ctx->no_device;
    return QIW_OK;
}

int qiw_profile_read(qiw_context* ctx, double* ms, int64_t* launches, int32_t reset) {
    if (!ctx) return QIW_ERR_BAD_ARG;
    if (!ctx->no_device) {
        cudaSetDevice(ctx->device);
        CK(cudaStreamSynchronize(ctx->stream));
        for (auto& pe : ctx->prof_events) {
            float t = 0;
            if (cudaEventElapsedTime(&t, pe.a, pe.b) == cudaSuccess) { ctx->prof_ms[pe.cls] += t; ctx->prof_n[pe.cls] += 1; }
            cudaEventDestroy(pe.a); cudaEventDestroy(pe.b);
        }
        ctx->prof_events.clear();
    }
    for (int k = 0; k < QIW_PROFILE_CLASSES; ++k) {
        if (ms) ms[k] = ctx->prof_ms[k];
        if (launches) launches[k] = ctx->prof_n[k];
        if (reset) { ctx->prof_ms[k] = 0; ctx->prof_n[k] = 0; }
    }
    return QIW_OK;
}

// ---- measurement --------------------------------------------------------------------------------------

int qiw_measure_fp64_peak(qiw_context* ctx, double* tflops) {
    if (!ctx || !tflops) return QIW_ERR_BAD_ARG;
    if (ctx->no_device) return fail(ctx, QIW_ERR_CUDA, "planning-only context cannot compute");
    cudaSetDevice(ctx->device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int blocks = sms * 8, iters = 1 << 14;
    DevBuf<double> buf;
    CK(buf.reserve((size_t)blocks * 256));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        CK(launch_dfma_peak(buf.p, blocks, iters, ctx->stream));
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->launches++;
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        const double fl = (double)blocks * 256.0 * iters * 8.0 * 2.0;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    buf.release();
    *tflops = best;
    return QIW_OK;
}

}  // extern "C"

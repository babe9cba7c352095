// libta_b200.so -- host runtime + C ABI (include/ta_b200.h) around the sm_100a
// kernels in kernels.cuh.  Plain CUDA runtime + NCCL (dlopen'ed, only touched
// when more than one device / rank takes part); no PyTorch, no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ta_b200.h"
#include "fft_plan.h"
#include "kernels.cuh"

using namespace ta;

namespace {

std::string g_last_error;  // for failures before a context exists

// ------------------------------------------------------------------ NCCL (lazy)
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool nccl_load(std::string* err) {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        *err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return false;
    }
#define TA_SYM(field, name)                                             \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
    if (!g_nccl.field) { *err = std::string("NCCL symbol missing: ") + name; return false; }
    TA_SYM(GetUniqueId, "ncclGetUniqueId");
    TA_SYM(CommInitRank, "ncclCommInitRank");
    TA_SYM(CommInitAll, "ncclCommInitAll");
    TA_SYM(CommDestroy, "ncclCommDestroy");
    TA_SYM(AllReduce, "ncclAllReduce");
    TA_SYM(GroupStart, "ncclGroupStart");
    TA_SYM(GroupEnd, "ncclGroupEnd");
    TA_SYM(GetErrorString, "ncclGetErrorString");
#undef TA_SYM
    g_nccl.handle = h;
    return true;
}

constexpr int kNumSlabs = 3;
constexpr int kNumDevStage = 2;

// a contiguous range of this shard's particles handled by one kernel launch
struct LaunchRange {
    int64_t a0 = 0, n = 0;
    cudaEvent_t ready = nullptr;   // recorded when the range's series are in place (may be null)
};

struct Shard {
    int dev = 0;
    int num_sms = 0;
    int max_smem = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_k0 = nullptr;
    cudaEvent_t ev_copy_done[kNumDevStage] = {nullptr, nullptr};
    cudaEvent_t ev_k0_done[kNumDevStage] = {nullptr, nullptr};
    cudaEvent_t ev_slab[kNumSlabs] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_ka = nullptr, ev_kb = nullptr;   // around the last K1/K2/K3 launch
    bool kernel_timed = false;
    ncclComm_t comm = nullptr;
    // problem-sized state
    int64_t atom0 = 0, natoms = 0;
    double natoms_d = 0.0;
    void* series = nullptr;             // [natoms][D][Tld] of double (FP64) or float (FP32 mode)
    double* by_particle = nullptr;
    double* masses = nullptr;
    double* ts_sum = nullptr;
    double* ts_mean = nullptr;          // [Tld] atom mean of the last compute call (first shard only); + scratch of ta_green_kubo
    double* partial = nullptr;
    size_t partial_rows = 0;
    void* dstage[kNumDevStage][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int stage_toggle = 0;
    // whole-trajectory staging (ta_stage_bulk): particle chunks, each with all T frames
    void* dbulk[kNumDevStage][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    size_t dbulk_bytes = 0;
    cudaEvent_t ev_bulk_copy[kNumDevStage] = {nullptr, nullptr};
    cudaEvent_t ev_bulk_k0[kNumDevStage] = {nullptr, nullptr};
    std::vector<LaunchRange> chunks;    // particle ranges in staging order, with their "series ready" events
    bool chunks_pending = false;        // the next compute call is pipelined behind the staging, chunk by chunk
    double* lagmajor_tmp = nullptr;
    size_t lagmajor_bytes = 0;
    void* l2_scratch = nullptr;
    void* win_scratch = nullptr;        // K2/K3 with T beyond shared memory: per-CTA series + lag sums
    size_t win_scratch_bytes = 0;
    // FFT tables
    void* tw_lo = nullptr;
    void* tw_hi = nullptr;
    uint32_t* ftab = nullptr;
    uint32_t* pair0 = nullptr;
    uint32_t* own0 = nullptr;
    // fast-path FFT tables (k1_fast.cuh), in the arithmetic type of the plan
    void* f_omega = nullptr;
    void* f_tw2 = nullptr;
    uint32_t* f_map = nullptr;
    void* f_wbase = nullptr;
    void* f_inv = nullptr;
    // FFT route of the Helfand MSD: counters + per-warp lists of the (particle, lag) pairs that need the exact evaluation
    uint32_t* hflags = nullptr;
    size_t hflags_bytes = 0;
};

}  // namespace

struct ta_ctx {
    std::vector<Shard> sh;
    int rank = 0, nranks = 1;
    bool have_comm = false;
    std::string err;
    // problem
    bool begun = false;
    int64_t T = 0, N = 0, Tld = 0;
    int D = 0, dims[3] = {0, 1, 2};
    int DS = 0;              // series rows per particle: D, + 1 row of sum_d g^2 for Helfand in FP64 (K5 reads it)
    int src_dtype = TA_DTYPE_F32, n_fields = 1, precision = TA_PRECISION_FP64;
    size_t elt = 4;
    int64_t frames_staged = 0;
    // pinned slab ring
    void* slab[kNumSlabs] = {nullptr, nullptr, nullptr};
    int64_t slab_frames = 0;
    int cur_slab = -1;       // slab handed out by the last ta_stage_slot
    int next_slab = 0;
    bool slab_used[kNumSlabs] = {false, false, false};
    // FFT plan cache
    FftPlanHost plan;
    int64_t plan_T = -1;
    int plan_prec = -1;
    int npairs0 = 0;
    int fast_r1 = 0;         // > 0: the plan is for the three-pass path (k1_fast.cuh) with this R1
    // debugging knobs, read from the environment ONCE when the context is created (never on a launch path):
    //   TA_B200_K1_PATH = general        the general mixed-radix FFT kernel even where the three-pass kernel applies
    //                   = notmem         the three-pass kernel with its output stage through L2 even where the tensor-memory
    //                                    build applies (the tests compare the kernels with each other)
    //   TA_B200_BULK_CHUNK = n           particles per staging chunk of ta_stage_bulk (tests: any chunking, same bits)
    //   TA_B200_HELFAND_FFT_THR = x      refinement threshold of ta_helfand_fft (scripts/helfand_fft_error_constant.py)
    bool opt_general_fft = false, opt_no_tmem = false;
    int64_t opt_bulk_chunk = 0;
    double opt_helfand_thr = -2.0;       // < -1.5: the built-in rule
    int k1_threads = 0, k1_smem = 0, k1_grid = 0;
    bool k1_tmem = false;    // the last K1 launch kept its output-stage streams in tensor memory
    bool have_mean = false;  // ts_mean of the first shard holds the result of a compute call
    int64_t launches = 0;
    long long helfand_fft_flagged = 0;   // (particle, lag) pairs the last ta_helfand_fft evaluated exactly; -1: all (K3 took over)
};

namespace {

int fail(ta_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    g_last_error = msg;
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(ctx, TA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define CKN(call)                                                                         \
    do {                                                                                  \
        ncclResult_t r_ = (call);                                                         \
        if (r_ != ncclSuccess)                                                            \
            return fail(ctx, TA_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

void free_problem(ta_ctx* c) {
    for (auto& s : c->sh) {
        cudaSetDevice(s.dev);
        cudaStreamSynchronize(s.s_compute);
        cudaStreamSynchronize(s.s_copy);
        cudaStreamSynchronize(s.s_k0);
        cudaFree(s.series); s.series = nullptr;
        cudaFree(s.by_particle); s.by_particle = nullptr;
        cudaFree(s.masses); s.masses = nullptr;
        cudaFree(s.ts_sum); s.ts_sum = nullptr;
        cudaFree(s.ts_mean); s.ts_mean = nullptr;
        cudaFree(s.partial); s.partial = nullptr; s.partial_rows = 0;
        cudaFree(s.lagmajor_tmp); s.lagmajor_tmp = nullptr; s.lagmajor_bytes = 0;
        cudaFree(s.win_scratch); s.win_scratch = nullptr; s.win_scratch_bytes = 0;
        for (int b = 0; b < kNumDevStage; ++b)
            for (int f = 0; f < 2; ++f) {
                cudaFree(s.dstage[b][f]); s.dstage[b][f] = nullptr;
                cudaFree(s.dbulk[b][f]); s.dbulk[b][f] = nullptr;
            }
        s.dbulk_bytes = 0;
        for (auto& c : s.chunks) if (c.ready) cudaEventDestroy(c.ready);
        s.chunks.clear();
        s.chunks_pending = false;
        cudaFree(s.tw_lo); s.tw_lo = nullptr;
        cudaFree(s.tw_hi); s.tw_hi = nullptr;
        cudaFree(s.ftab); s.ftab = nullptr;
        cudaFree(s.pair0); s.pair0 = nullptr;
        cudaFree(s.own0); s.own0 = nullptr;
        cudaFree(s.f_omega); s.f_omega = nullptr;
        cudaFree(s.f_tw2); s.f_tw2 = nullptr;
        cudaFree(s.f_map); s.f_map = nullptr;
        cudaFree(s.f_wbase); s.f_wbase = nullptr;
        cudaFree(s.f_inv); s.f_inv = nullptr;
        cudaFree(s.hflags); s.hflags = nullptr; s.hflags_bytes = 0;
    }
    for (int i = 0; i < kNumSlabs; ++i) {
        if (c->slab[i]) cudaFreeHost(c->slab[i]);
        c->slab[i] = nullptr;
        c->slab_used[i] = false;
    }
    c->begun = false;
    c->have_mean = false;
    c->plan_T = -1;
    c->frames_staged = 0;
    c->cur_slab = -1;
    c->next_slab = 0;
}

int init_shard(ta_ctx* ctx, Shard& s, int dev) {
    s.dev = dev;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(ctx, TA_ERR_UNSUPPORTED,
                    std::string("device ") + prop.name + " is not sm_100 class; this library is built for sm_100a only");
    s.num_sms = prop.multiProcessorCount;
    CK(cudaDeviceGetAttribute(&s.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CK(cudaStreamCreateWithFlags(&s.s_compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.s_copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.s_k0, cudaStreamNonBlocking));
    for (int b = 0; b < kNumDevStage; ++b) {
        CK(cudaEventCreateWithFlags(&s.ev_copy_done[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.ev_k0_done[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.ev_bulk_copy[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.ev_bulk_k0[b], cudaEventDisableTiming));
    }
    for (int i = 0; i < kNumSlabs; ++i) CK(cudaEventCreateWithFlags(&s.ev_slab[i], cudaEventDisableTiming));
    CK(cudaEventCreate(&s.ev_t0));
    CK(cudaEventCreate(&s.ev_t1));
    CK(cudaEventCreate(&s.ev_ka));
    CK(cudaEventCreate(&s.ev_kb));
    return TA_OK;
}

int sync_all(ta_ctx* ctx) {
    for (auto& s : ctx->sh) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_copy));
        CK(cudaStreamSynchronize(s.s_k0));
        CK(cudaStreamSynchronize(s.s_compute));
    }
    return TA_OK;
}

// K0 on `stream`: staged slab [nframes][n][3] (particles a0 .. a0+n-1 of the shard) -> series.
template <typename SRC, typename OUT>
int launch_k0_t(ta_ctx* ctx, Shard& s, const void* vsrc, const void* xsrc, int64_t a0, int64_t n, int64_t nframes,
                int64_t frame0, cudaStream_t stream) {
    dim3 block(32, 8);
    // frames on grid.x (up to 2^31 - 1 tiles), particle tiles on grid.y (the kernel loops when there are more than 65,535)
    dim3 grid((unsigned)((nframes + K0_FR - 1) / K0_FR), (unsigned)std::min<int64_t>((n + K0_AT - 1) / K0_AT, 65535));
    const int d0 = ctx->dims[0], d1 = ctx->dims[1], d2 = ctx->dims[2];
    OUT* series = (OUT*)s.series + (size_t)a0 * ctx->DS * ctx->Tld;
    const double* masses = s.masses ? s.masses + a0 : nullptr;
    const SRC* v = (const SRC*)vsrc;
    const SRC* x = (const SRC*)xsrc;
    if (ctx->n_fields == 2) k0_stage<SRC, true, OUT><<<grid, block, 0, stream>>>(v, x, masses, series, (int)n, (int)nframes, frame0, ctx->Tld, ctx->D, ctx->DS, d0, d1, d2);
    else k0_stage<SRC, false, OUT><<<grid, block, 0, stream>>>(v, x, masses, series, (int)n, (int)nframes, frame0, ctx->Tld, ctx->D, ctx->DS, d0, d1, d2);
    CK(cudaGetLastError());
    ctx->launches++;
    return TA_OK;
}

int launch_k0(ta_ctx* ctx, Shard& s, const void* vsrc, const void* xsrc, int64_t a0, int64_t n, int64_t nframes,
              int64_t frame0, cudaStream_t stream) {
    const bool f32src = ctx->src_dtype == TA_DTYPE_F32, fp64 = ctx->precision == TA_PRECISION_FP64;
    if (f32src) return fp64 ? launch_k0_t<float, double>(ctx, s, vsrc, xsrc, a0, n, nframes, frame0, stream)
                            : launch_k0_t<float, float>(ctx, s, vsrc, xsrc, a0, n, nframes, frame0, stream);
    return fp64 ? launch_k0_t<double, double>(ctx, s, vsrc, xsrc, a0, n, nframes, frame0, stream)
                : launch_k0_t<double, float>(ctx, s, vsrc, xsrc, a0, n, nframes, frame0, stream);
}

// The particle ranges the next K1/K2/K3 call launches over: the staging chunks (each launch waits
// for its chunk's series, so the correlation of chunk c overlaps the H2D copy of chunk c+1) right
// after ta_stage_bulk, otherwise the whole shard in one launch.
std::vector<LaunchRange> take_launch_ranges(Shard& s) {
    std::vector<LaunchRange> r;
    if (s.chunks_pending && !s.chunks.empty()) r = s.chunks;
    else { LaunchRange all; all.a0 = 0; all.n = s.natoms; r.push_back(all); }
    s.chunks_pending = false;
    return r;
}

// Enqueue H2D of `nframes` frames for every shard + the K0 transposition.
// `base[f]`: host pointer to (frame 0 of this chunk, source atom 0 of particle
// 0) of field f; consecutive analysed frames are `pitch` bytes apart.
int enqueue_chunk(ta_ctx* ctx, const char* const* base, size_t pitch, int64_t frame0, int64_t nframes,
                  int slab_index) {
    const size_t elt = ctx->elt;
    for (auto& s : ctx->sh) {
        if (s.natoms == 0) continue;
        CK(cudaSetDevice(s.dev));
        const int b = s.stage_toggle;
        s.stage_toggle ^= 1;
        CK(cudaStreamWaitEvent(s.s_copy, s.ev_k0_done[b], 0));
        const size_t width = (size_t)s.natoms * 3 * elt;
        for (int f = 0; f < ctx->n_fields; ++f) {
            const char* src = base[f] + (size_t)s.atom0 * 3 * elt;
            CK(cudaMemcpy2DAsync(s.dstage[b][f], width, src, pitch, width, (size_t)nframes,
                                 cudaMemcpyHostToDevice, s.s_copy));
        }
        CK(cudaEventRecord(s.ev_copy_done[b], s.s_copy));
        if (slab_index >= 0) CK(cudaEventRecord(s.ev_slab[slab_index], s.s_copy));
        CK(cudaStreamWaitEvent(s.s_compute, s.ev_copy_done[b], 0));
        int rc = launch_k0(ctx, s, s.dstage[b][0], s.dstage[b][1], 0, s.natoms, nframes, frame0, s.s_compute);
        if (rc) return rc;
        CK(cudaEventRecord(s.ev_k0_done[b], s.s_compute));
    }
    ctx->frames_staged += nframes;
    return TA_OK;
}

int ensure_partial(ta_ctx* ctx, Shard& s, size_t rows) {
    if (s.partial_rows < rows) {
        cudaFree(s.partial);
        s.partial = nullptr;
        s.partial_rows = 0;
        CK(cudaMalloc(&s.partial, rows * (size_t)ctx->Tld * sizeof(double)));
        s.partial_rows = rows;
    }
    CK(cudaMemsetAsync(s.partial, 0, rows * (size_t)ctx->Tld * sizeof(double), s.s_compute));
    return TA_OK;
}

// Sum the per-CTA partial rows, all-reduce over devices / ranks, divide by the
// total particle count: results.timeseries = by_particle.mean(axis=1)
// (velocityautocorr.py:214,237; viscosity.py:233).
int finish_timeseries(ta_ctx* ctx, const std::vector<int>& grids, double* ts_out) {
    const int T = (int)ctx->T;
    const size_t cnt = (size_t)ctx->Tld + 1;   // + particle count in the last slot
    for (size_t i = 0; i < ctx->sh.size(); ++i) {
        Shard& s = ctx->sh[i];
        CK(cudaSetDevice(s.dev));
        CK(cudaMemsetAsync(s.ts_sum, 0, cnt * sizeof(double), s.s_compute));
        if (s.natoms > 0) {
            k_sum_partials<<<(T + 255) / 256, 256, 0, s.s_compute>>>(s.partial, grids[i], ctx->Tld, T, s.ts_sum);
            CK(cudaGetLastError());
            ctx->launches++;
        }
        s.natoms_d = (double)s.natoms;
        CK(cudaMemcpyAsync(s.ts_sum + ctx->Tld, &s.natoms_d, sizeof(double), cudaMemcpyHostToDevice, s.s_compute));
    }
    if (ctx->have_comm) {
        CKN(g_nccl.GroupStart());
        for (auto& s : ctx->sh) {
            ncclResult_t r = g_nccl.AllReduce(s.ts_sum, s.ts_sum, cnt, ncclDouble, ncclSum, s.comm, s.s_compute);
            if (r != ncclSuccess) {
                g_nccl.GroupEnd();
                return fail(ctx, TA_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
            }
        }
        CKN(g_nccl.GroupEnd());
    }
    // the atom mean is formed on the device (it stays there for ta_green_kubo) and copied out
    Shard& s0 = ctx->sh[0];
    CK(cudaSetDevice(s0.dev));
    k_atom_mean<<<(T + 255) / 256, 256, 0, s0.s_compute>>>(s0.ts_sum, ctx->Tld, T, s0.ts_mean);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(ts_out, s0.ts_mean, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, s0.s_compute));
    int rc = sync_all(ctx);
    if (rc) return rc;
    ctx->have_mean = true;
    return TA_OK;
}

int check_ready(ta_ctx* ctx) {
    if (!ctx) return fail(nullptr, TA_ERR_INVALID, "null context");
    if (!ctx->begun) return fail(ctx, TA_ERR_INVALID, "ta_stage_begin has not been called");
    if (ctx->frames_staged < ctx->T)
        return fail(ctx, TA_ERR_INVALID, "only " + std::to_string(ctx->frames_staged) + " of " +
                                             std::to_string(ctx->T) + " frames were staged");
    return TA_OK;
}

// Table uploads go through the COMPUTE stream and are complete on return.  A plain cudaMemcpy from pageable memory
// returns once the data sits in the driver's staging buffer; its DMA is ordered on the legacy stream only, which the
// library's non-blocking streams do not wait for -- with the copy engine busy streaming trajectory chunks an 80 KB
// table reached the device AFTER the first K1 launches had read it (zeros for every chunk but the last: found in
// round 2, profiles/r02_bulk_pipeline_race.txt).
template <typename V>
int upload_vec(ta_ctx* ctx, Shard& s, void** dst, const std::vector<V>& v) {
    CK(cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(V)));
    if (!v.empty()) CK(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice, s.s_compute));
    return TA_OK;
}

template <typename R>
int upload_fft_tables(ta_ctx* ctx, const std::vector<uint32_t>& own0) {
    const FftPlanHost& p = ctx->plan;
    const int nlo = 1 << p.lo_bits, nhi = (int)p.tw_hi.size() / 2;
    std::vector<cplx<R>> lo(nlo), hi(nhi);
    for (int i = 0; i < nlo; ++i) lo[i] = cmake<R>((R)p.tw_lo[2 * i], (R)p.tw_lo[2 * i + 1]);
    for (int i = 0; i < nhi; ++i) hi[i] = cmake<R>((R)p.tw_hi[2 * i], (R)p.tw_hi[2 * i + 1]);
    for (auto& s : ctx->sh) {
        CK(cudaSetDevice(s.dev));
        cudaFree(s.tw_lo); cudaFree(s.tw_hi); cudaFree(s.ftab); cudaFree(s.pair0); cudaFree(s.own0);
        s.tw_lo = s.tw_hi = nullptr; s.ftab = s.pair0 = s.own0 = nullptr;
        int rc;
        if ((rc = upload_vec(ctx, s, &s.tw_lo, lo)) || (rc = upload_vec(ctx, s, &s.tw_hi, hi)) ||
            (rc = upload_vec(ctx, s, (void**)&s.ftab, p.ftab)) || (rc = upload_vec(ctx, s, (void**)&s.pair0, p.pair0)) ||
            (rc = upload_vec(ctx, s, (void**)&s.own0, own0)))
            return rc;
        CK(cudaStreamSynchronize(s.s_compute));          // the host vectors go out of scope; the tables are in place
    }
    return TA_OK;
}

template <typename RT>
int upload_fast_tables(ta_ctx* ctx, const K1FastPlan& p) {
    auto conv = [](const std::vector<double>& v) { return std::vector<RT>(v.begin(), v.end()); };
    const std::vector<RT> omega = conv(p.omega), tw2 = conv(p.tw2), wbase = conv(p.wbase), inv = conv(p.inv);
    for (auto& s : ctx->sh) {
        CK(cudaSetDevice(s.dev));
        cudaFree(s.f_omega); cudaFree(s.f_tw2); cudaFree(s.f_map); cudaFree(s.f_wbase); cudaFree(s.f_inv);
        s.f_omega = s.f_tw2 = s.f_wbase = s.f_inv = nullptr; s.f_map = nullptr;
        int rc;
        if ((rc = upload_vec(ctx, s, &s.f_omega, omega)) || (rc = upload_vec(ctx, s, &s.f_tw2, tw2)) ||
            (rc = upload_vec(ctx, s, (void**)&s.f_map, p.map)) || (rc = upload_vec(ctx, s, &s.f_wbase, wbase)) ||
            (rc = upload_vec(ctx, s, &s.f_inv, inv)))
            return rc;
        CK(cudaStreamSynchronize(s.s_compute));
    }
    return TA_OK;
}

// Which K1 kernel serves T: the three-pass radix-16 kernel where it has an instantiation (H = 256 R1 >= T / 2 with at
// most 34 % padding: T from ~1,540 to 12,288; FP64 and FP32), else the general mixed-radix kernel (any T).
int ensure_fft_plan(ta_ctx* ctx) {
    if (ctx->plan_T == ctx->T && ctx->plan_prec == ctx->precision) return TA_OK;
    ctx->fast_r1 = 0;
    const bool fp64 = ctx->precision == TA_PRECISION_FP64;
    const int r1 = ctx->opt_general_fft ? 0 : k1f_choose_r1(ctx->T);
    if (r1 > 0) {
        K1FastPlan fp;
        int rcf = k1f_build_plan(ctx->T, ctx->Tld, r1, &fp);
        if (rcf) return fail(ctx, rcf, "cannot plan the fast FFT path for T=" + std::to_string(ctx->T));
        if ((rcf = fp64 ? upload_fast_tables<double>(ctx, fp) : upload_fast_tables<float>(ctx, fp))) return rcf;
        ctx->fast_r1 = r1;
        ctx->plan.H = fp.H; ctx->plan.L = fp.L; ctx->plan.npasses = 3;
        ctx->plan.radix[0] = r1; ctx->plan.radix[1] = 16; ctx->plan.radix[2] = 16;
        ctx->plan_T = ctx->T;
        ctx->plan_prec = ctx->precision;
        return TA_OK;
    }
    int rc = ta_build_fft_plan(ctx->T, &ctx->plan);
    if (rc) return fail(ctx, rc, "cannot plan an FFT for T=" + std::to_string(ctx->T));
    std::vector<uint32_t> own0;
    for (int p = 0; p < ctx->plan.H; ++p)
        if ((uint32_t)p <= ctx->plan.pair0[p]) own0.push_back((uint32_t)p);
    ctx->npairs0 = (int)own0.size();
    rc = fp64 ? upload_fft_tables<double>(ctx, own0) : upload_fft_tables<float>(ctx, own0);
    if (rc) return rc;
    ctx->plan_T = ctx->T;
    ctx->plan_prec = ctx->precision;
    return TA_OK;
}

template <typename R>
int launch_fft(ta_ctx* ctx, std::vector<int>* grids) {
    const FftPlanHost& p = ctx->plan;
    const int nlo = 1 << p.lo_bits, nhi = (int)p.tw_hi.size() / 2;
    const size_t tab_bytes = (size_t)(nlo + nhi) * sizeof(cplx<R>);
    size_t work = (size_t)p.H * sizeof(cplx<R>) + (size_t)(p.H + 1) * sizeof(R);   // FFT buffer + pair accumulators
    work = (work + 15) & ~(size_t)15;
    size_t smem = (work + tab_bytes + 15) & ~(size_t)15;
    int nthr = ((p.H / 8 + 31) / 32) * 32;
    nthr = std::max(32, std::min(K1_MAX_THREADS, nthr));
    grids->assign(ctx->sh.size(), 0);
    for (size_t i = 0; i < ctx->sh.size(); ++i) {
        Shard& s = ctx->sh[i];
        if (s.natoms == 0) continue;
        CK(cudaSetDevice(s.dev));
        // buffer + accumulators of one particle: shared memory when they fit, else a per-CTA global work area (any T)
        const bool in_smem = smem <= (size_t)s.max_smem;
        const size_t dyn = in_smem ? smem : ((tab_bytes + 15) & ~(size_t)15);
        if (dyn > (size_t)s.max_smem)
            return fail(ctx, TA_ERR_UNSUPPORTED, "FFT route: the twiddle tables for T=" + std::to_string(ctx->T) +
                                                     " do not fit in shared memory; use fft=False");
        auto kern = in_smem ? k1_fft_acf<R, false> : k1_fft_acf<R, true>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthr, dyn));
        if (occ < 1) return fail(ctx, TA_ERR_UNSUPPORTED, "FFT kernel does not fit on an SM");
        if (!in_smem) occ = std::min(occ, 2);        // the work areas should stay L2-resident
        int grid = (int)std::min<int64_t>(s.natoms, (int64_t)s.num_sms * occ);
        (*grids)[i] = grid;
        int rc = ensure_partial(ctx, s, (size_t)grid);
        if (rc) return rc;
        if (!in_smem) {
            const size_t need = work * (size_t)grid;
            if (s.win_scratch_bytes < need) {
                cudaFree(s.win_scratch);
                s.win_scratch = nullptr; s.win_scratch_bytes = 0;
                CK(cudaMalloc(&s.win_scratch, need));
                s.win_scratch_bytes = need;
            }
        }
        K1Args<R> a;
        a.t.T = (int)ctx->T; a.t.H = p.H; a.t.L = p.L; a.t.npasses = p.npasses;
        for (int q = 0; q < TA_MAX_PASSES; ++q) a.t.radix[q] = p.radix[q];
        a.t.lo_bits = p.lo_bits;
        a.t.tw_lo = (const cplx<R>*)s.tw_lo;
        a.t.tw_hi = (const cplx<R>*)s.tw_hi;
        a.t.ftab = s.ftab; a.t.pair0 = s.pair0; a.t.own0 = s.own0; a.t.npairs0 = ctx->npairs0;
        a.nlo = nlo; a.nhi = nhi;
        a.partial = s.partial;
        a.D = ctx->D; a.DS = ctx->DS; a.Tld = ctx->Tld;
        a.scratch = in_smem ? nullptr : (unsigned char*)s.win_scratch;
        a.scratch_stride = (long long)work;
        CK(cudaEventRecord(s.ev_ka, s.s_compute));
        for (const LaunchRange& rg : take_launch_ranges(s)) {
            if (rg.ready) CK(cudaStreamWaitEvent(s.s_compute, rg.ready, 0));
            a.series = (const R*)s.series + (size_t)rg.a0 * ctx->DS * ctx->Tld;
            a.by_particle = s.by_particle + (size_t)rg.a0 * ctx->Tld;
            a.natoms = (int)rg.n;
            kern<<<(int)std::min<int64_t>(grid, rg.n), nthr, dyn, s.s_compute>>>(a);
            CK(cudaGetLastError());
            ctx->launches++;
        }
        CK(cudaEventRecord(s.ev_kb, s.s_compute));
        s.kernel_timed = true;
        ctx->k1_threads = nthr; ctx->k1_smem = (int)dyn; ctx->k1_grid = grid; ctx->k1_tmem = false;
    }
    return TA_OK;
}

// PART = false (FP64 only): without the per-CTA particle sums (ta_helfand_fft does not use them)
template <int R1, typename RT, bool PART = true>
int launch_fft_fast_r1(ta_ctx* ctx, std::vector<int>* grids) {
    constexpr int NT = k1f_threads(R1);
    const int smem = k1f_smem_bytes(R1, k1f_prefetch(R1, (int)sizeof(RT)), (int)sizeof(RT));
    grids->assign(ctx->sh.size(), 0);
    for (size_t i = 0; i < ctx->sh.size(); ++i) {
        Shard& s = ctx->sh[i];
        if (s.natoms == 0) continue;
        CK(cudaSetDevice(s.dev));
        void (*kern)(const K1FArgs<RT>) = k1f_fft_acf<R1, RT, PART>;
        if (smem > s.max_smem) return fail(ctx, TA_ERR_UNSUPPORTED, "three-pass FFT kernel needs more shared memory than the device has");
        int occ = 0;
        bool tmem = false;
        if constexpr (sizeof(RT) == 8 && (R1 == 16 || R1 == 20)) {
            // one CTA per SM, one P1 column per thread: the output stage keeps its per-thread streams in tensor memory
            void (*kt)(const K1FArgs<RT>) = NT == 320 ? k1f_fft_acf_mr<R1, RT, PART, true> : k1f_fft_acf<R1, RT, PART, true>;
            if (!ctx->opt_no_tmem &&
                cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kt, NT, (size_t)smem) == cudaSuccess && occ == 1) {
                kern = kt;
                tmem = true;
            } else cudaGetLastError();
        }
        if (!tmem) if constexpr (NT == 320 && sizeof(RT) == 8) {
            // ten FP64 warps per SM: the build with the register cap stated outright (kernels.cuh); should a compiler
            // settle above what fits, the launch-bounds build of the same kernel takes over
            cudaFuncSetAttribute(k1f_fft_acf_mr<R1, RT, PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k1f_fft_acf_mr<R1, RT, PART>, NT, (size_t)smem) == cudaSuccess && occ >= 1)
                kern = k1f_fft_acf_mr<R1, RT, PART>;
            else cudaGetLastError();
        }
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, (size_t)smem));
        if (occ < 1) return fail(ctx, TA_ERR_UNSUPPORTED, "three-pass FFT kernel does not fit on an SM");
        int grid = (int)std::min<int64_t>(s.natoms, (int64_t)s.num_sms * occ);
        (*grids)[i] = grid;
        int rc = ensure_partial(ctx, s, (size_t)grid);
        if (rc) return rc;
        K1FArgs<RT> a;
        a.partial = s.partial;
        a.omega = (const cplx<RT>*)s.f_omega; a.tw2 = (const cplx<RT>*)s.f_tw2; a.map = s.f_map;
        a.wbase = (const cplx<RT>*)s.f_wbase; a.inv = (const RT*)s.f_inv;
        a.D = ctx->D; a.DS = ctx->DS; a.T = (int)ctx->T; a.nh = (int)((ctx->T + 1) / 2); a.Tld = ctx->Tld;
        CK(cudaEventRecord(s.ev_ka, s.s_compute));
        for (const LaunchRange& rg : take_launch_ranges(s)) {
            if (rg.ready) CK(cudaStreamWaitEvent(s.s_compute, rg.ready, 0));
            a.series = (const RT*)s.series + (size_t)rg.a0 * ctx->DS * ctx->Tld;
            a.by_particle = s.by_particle + (size_t)rg.a0 * ctx->Tld;
            a.natoms = (int)rg.n;
            kern<<<(int)std::min<int64_t>(grid, rg.n), NT, smem, s.s_compute>>>(a);
            CK(cudaGetLastError());
            ctx->launches++;
        }
        CK(cudaEventRecord(s.ev_kb, s.s_compute));
        s.kernel_timed = true;
        ctx->k1_threads = NT; ctx->k1_smem = smem; ctx->k1_grid = grid; ctx->k1_tmem = tmem;
    }
    return TA_OK;
}

template <typename RT, bool PART = true>
int launch_fft_fast(ta_ctx* ctx, std::vector<int>* grids) {
    switch (ctx->fast_r1) {
        case 4: return launch_fft_fast_r1<4, RT, PART>(ctx, grids);
        case 6: return launch_fft_fast_r1<6, RT, PART>(ctx, grids);
        case 8: return launch_fft_fast_r1<8, RT, PART>(ctx, grids);
        case 10: return launch_fft_fast_r1<10, RT, PART>(ctx, grids);
        case 12: return launch_fft_fast_r1<12, RT, PART>(ctx, grids);
        case 16: return launch_fft_fast_r1<16, RT, PART>(ctx, grids);
        case 20: return launch_fft_fast_r1<20, RT, PART>(ctx, grids);
        case 24: return launch_fft_fast_r1<24, RT, PART>(ctx, grids);
    }
    return fail(ctx, TA_ERR_UNSUPPORTED, "no fast FFT instantiation for R1=" + std::to_string(ctx->fast_r1));
}

// the FFT autocorrelation pass of ta_vacf_fft / ta_helfand_fft: by_particle = sum_d acf_d, per-CTA partial rows
int launch_k1(ta_ctx* ctx, std::vector<int>* grids, bool with_partial = true) {
    const bool fp64 = ctx->precision == TA_PRECISION_FP64;
    if (ctx->fast_r1 > 0 && fp64 && !with_partial) return launch_fft_fast<double, false>(ctx, grids);
    if (ctx->fast_r1 > 0) return fp64 ? launch_fft_fast<double>(ctx, grids) : launch_fft_fast<float>(ctx, grids);
    return fp64 ? launch_fft<double>(ctx, grids) : launch_fft<float>(ctx, grids);
}

template <typename R, int MODE>
int launch_windowed(ta_ctx* ctx, double denom, std::vector<int>* grids) {
    const int T = (int)ctx->T;
    const int ne = win_smem_elems(T);
    size_t smem = (size_t)((ne + 1) & ~1) * sizeof(R) + (size_t)T * sizeof(double);
    smem = (smem + 15) & ~(size_t)15;
    const int npairs = win_num_pairs(win_num_lag_blocks(T));
    const int nthr = 32 * std::max(1, std::min(KW_MAX_THREADS / 32, npairs));
    grids->assign(ctx->sh.size(), 0);
    for (size_t i = 0; i < ctx->sh.size(); ++i) {
        Shard& s = ctx->sh[i];
        if (s.natoms == 0) continue;
        CK(cudaSetDevice(s.dev));
        // series + lag sums of one particle: shared memory when they fit, else a per-CTA global scratch area
        const bool in_smem = smem <= (size_t)s.max_smem;
        const size_t dyn_smem = in_smem ? smem : 0;
        if (in_smem) CK(cudaFuncSetAttribute(k_windowed<R, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        if (in_smem) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_windowed<R, MODE, false>, nthr, dyn_smem));
        else occ = 1;
        if (occ < 1) return fail(ctx, TA_ERR_UNSUPPORTED, "windowed kernel does not fit on an SM");
        // fewer than four particles per resident CTA: deal each particle's lag-block pairs to several CTAs, so that the
        // last wave of the grid is full (BASELINE configs[1]: 1,000 particles on 592 resident CTAs)
        const int64_t resident = (int64_t)s.num_sms * occ;
        const int nwarps = std::max(1, nthr / 32);
        const int max_split = std::max(1, (npairs + nwarps - 1) / nwarps);      // every part keeps at least one round of pairs
        int nsplit = 1;
        if (in_smem && s.natoms < 4 * resident) {
            // the split whose last wave is fullest (each part re-stages the series: ~2 % per extra part)
            double best = 1e30;
            for (int ns = 1; ns <= max_split; ++ns) {
                const double units = (double)s.natoms * ns, waves = std::ceil(units / (double)resident);
                const double cost = waves * (double)resident / units * (1.0 + 0.02 * (ns - 1));
                if (cost < best - 1e-9) { best = cost; nsplit = ns; }
            }
        }
        int grid = (int)std::min<int64_t>(s.natoms * nsplit, resident);
        if (!in_smem) {
            const size_t need = smem * (size_t)grid;
            if (s.win_scratch_bytes < need) {
                cudaFree(s.win_scratch);
                s.win_scratch = nullptr; s.win_scratch_bytes = 0;
                CK(cudaMalloc(&s.win_scratch, need));
                s.win_scratch_bytes = need;
            }
        }
        (*grids)[i] = grid;
        int rc = ensure_partial(ctx, s, (size_t)grid);
        if (rc) return rc;
        WinArgs a;
        a.partial = s.partial;
        a.D = ctx->D; a.DS = ctx->DS; a.T = T; a.Tld = ctx->Tld; a.denom = denom;
        a.scratch = in_smem ? nullptr : (unsigned char*)s.win_scratch;
        a.scratch_stride = (long long)smem;
        a.nsplit = nsplit;
        CK(cudaEventRecord(s.ev_ka, s.s_compute));
        for (const LaunchRange& rg : take_launch_ranges(s)) {
            if (rg.ready) CK(cudaStreamWaitEvent(s.s_compute, rg.ready, 0));
            a.series = (const R*)s.series + (size_t)rg.a0 * ctx->DS * ctx->Tld;
            a.by_particle = s.by_particle + (size_t)rg.a0 * ctx->Tld;
            a.natoms = (int)rg.n;
            if (in_smem) k_windowed<R, MODE, false><<<(int)std::min<int64_t>(grid, rg.n * nsplit), nthr, dyn_smem, s.s_compute>>>(a);
            else k_windowed<R, MODE, true><<<(int)std::min<int64_t>(grid, rg.n), nthr, 0, s.s_compute>>>(a);
            CK(cudaGetLastError());
            ctx->launches++;
        }
        CK(cudaEventRecord(s.ev_kb, s.s_compute));
        s.kernel_timed = true;
    }
    return TA_OK;
}

int create_common(ta_ctx* ctx, int ndev, const int* devices) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ctx, TA_ERR_CUDA,
                    std::string("no CUDA device available (") + cudaGetErrorString(e) +
                        "); libta_b200 has no CPU fallback");
    if (ndev < 1) return fail(ctx, TA_ERR_INVALID, "ndev must be >= 1");
    // debugging knobs (see ta_ctx): read here, once, never on a launch path
    if (const char* v = getenv("TA_B200_K1_PATH")) {
        ctx->opt_general_fft = std::string(v) == "general";
        ctx->opt_no_tmem = std::string(v) == "notmem";
    }
    if (const char* v = getenv("TA_B200_BULK_CHUNK")) ctx->opt_bulk_chunk = atoll(v);
    if (const char* v = getenv("TA_B200_HELFAND_FFT_THR")) ctx->opt_helfand_thr = atof(v);
    ctx->sh.resize(ndev);
    for (int i = 0; i < ndev; ++i) {
        int dev = devices ? devices[i] : i;
        if (dev < 0 || dev >= count)
            return fail(ctx, TA_ERR_INVALID, "device index " + std::to_string(dev) + " out of range");
        int rc = init_shard(ctx, ctx->sh[i], dev);
        if (rc) return rc;
    }
    return TA_OK;
}

void destroy_ctx(ta_ctx* c) {
    if (!c) return;
    free_problem(c);
    for (auto& s : c->sh) {
        cudaSetDevice(s.dev);
        if (s.comm && g_nccl.handle) g_nccl.CommDestroy(s.comm);
        cudaFree(s.l2_scratch);
        for (int b = 0; b < kNumDevStage; ++b) {
            if (s.ev_copy_done[b]) cudaEventDestroy(s.ev_copy_done[b]);
            if (s.ev_k0_done[b]) cudaEventDestroy(s.ev_k0_done[b]);
            if (s.ev_bulk_copy[b]) cudaEventDestroy(s.ev_bulk_copy[b]);
            if (s.ev_bulk_k0[b]) cudaEventDestroy(s.ev_bulk_k0[b]);
        }
        for (int i = 0; i < kNumSlabs; ++i) if (s.ev_slab[i]) cudaEventDestroy(s.ev_slab[i]);
        if (s.ev_t0) cudaEventDestroy(s.ev_t0);
        if (s.ev_t1) cudaEventDestroy(s.ev_t1);
        if (s.ev_ka) cudaEventDestroy(s.ev_ka);
        if (s.ev_kb) cudaEventDestroy(s.ev_kb);
        if (s.s_compute) cudaStreamDestroy(s.s_compute);
        if (s.s_copy) cudaStreamDestroy(s.s_copy);
        if (s.s_k0) cudaStreamDestroy(s.s_k0);
    }
    delete c;
}

}  // namespace

// =============================================================== C ABI
extern "C" {

int ta_version(void) { return 100; }

int ta_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (count) *count = n;
    return TA_OK;
}

const char* ta_last_error(const ta_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int ta_ctx_create(int ndev, const int* devices, ta_ctx** out) {
    if (!out) return fail(nullptr, TA_ERR_INVALID, "out is null");
    *out = nullptr;
    ta_ctx* ctx = new ta_ctx();
    int rc = create_common(ctx, ndev, devices);
    if (rc == TA_OK && ndev > 1) {
        std::string err;
        if (!nccl_load(&err)) rc = fail(ctx, TA_ERR_NCCL, err);
        if (rc == TA_OK) {
            std::vector<ncclComm_t> comms(ndev);
            std::vector<int> devs(ndev);
            for (int i = 0; i < ndev; ++i) devs[i] = ctx->sh[i].dev;
            ncclResult_t r = g_nccl.CommInitAll(comms.data(), ndev, devs.data());
            if (r != ncclSuccess) rc = fail(ctx, TA_ERR_NCCL, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
            else {
                for (int i = 0; i < ndev; ++i) ctx->sh[i].comm = comms[i];
                ctx->have_comm = true;
            }
        }
    }
    if (rc != TA_OK) {
        g_last_error = ctx->err;
        destroy_ctx(ctx);
        return rc;
    }
    *out = ctx;
    return TA_OK;
}

int ta_nccl_unique_id(void* id128) {
    std::string err;
    if (!id128) return fail(nullptr, TA_ERR_INVALID, "id128 is null");
    if (!nccl_load(&err)) return fail(nullptr, TA_ERR_NCCL, err);
    static_assert(sizeof(ncclUniqueId) == TA_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, TA_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    memcpy(id128, &id, sizeof(id));
    return TA_OK;
}

int ta_ctx_create_rank(int device, int rank, int nranks, const void* id128, ta_ctx** out) {
    if (!out) return fail(nullptr, TA_ERR_INVALID, "out is null");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(nullptr, TA_ERR_INVALID, "bad rank / nranks");
    ta_ctx* ctx = new ta_ctx();
    ctx->rank = rank;
    ctx->nranks = nranks;
    int rc = create_common(ctx, 1, &device);
    if (rc == TA_OK && nranks > 1) {
        std::string err;
        if (!id128) rc = fail(ctx, TA_ERR_INVALID, "id128 is null");
        else if (!nccl_load(&err)) rc = fail(ctx, TA_ERR_NCCL, err);
        else {
            ncclUniqueId id;
            memcpy(&id, id128, sizeof(id));
            cudaSetDevice(ctx->sh[0].dev);
            ncclResult_t r = g_nccl.CommInitRank(&ctx->sh[0].comm, nranks, id, rank);
            if (r != ncclSuccess) rc = fail(ctx, TA_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
            else ctx->have_comm = true;
        }
    }
    if (rc != TA_OK) {
        g_last_error = ctx->err;
        destroy_ctx(ctx);
        return rc;
    }
    *out = ctx;
    return TA_OK;
}

void ta_ctx_destroy(ta_ctx* ctx) { destroy_ctx(ctx); }

int ta_host_register(void* ptr, uint64_t bytes) {
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();               // not sticky: the caller may go on with pageable copies
        return fail(nullptr, TA_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    }
    return TA_OK;
}
int ta_host_unregister(void* ptr) {
    ta_ctx* ctx = nullptr;
    CK(cudaHostUnregister(ptr));
    return TA_OK;
}

int ta_stage_begin(ta_ctx* ctx, int64_t T, int64_t N, int D, const int* dims, int src_dtype,
                   int n_fields, const double* masses, int precision) {
    if (!ctx) return fail(nullptr, TA_ERR_INVALID, "null context");
    if (T < 1 || N < 0) return fail(ctx, TA_ERR_INVALID, "T must be >= 1 and N >= 0");
    if (T > (int64_t)1 << 24) return fail(ctx, TA_ERR_UNSUPPORTED, "T too large");
    if (D < 1 || D > 3 || !dims) return fail(ctx, TA_ERR_INVALID, "D must be 1..3 with dims given");
    for (int i = 0; i < D; ++i)
        if (dims[i] < 0 || dims[i] > 2) return fail(ctx, TA_ERR_INVALID, "dims entries must be 0, 1 or 2");
    if (src_dtype != TA_DTYPE_F32 && src_dtype != TA_DTYPE_F64) return fail(ctx, TA_ERR_INVALID, "bad src_dtype");
    if (n_fields != 1 && n_fields != 2) return fail(ctx, TA_ERR_INVALID, "n_fields must be 1 or 2");
    if (n_fields == 2 && !masses && N > 0) return fail(ctx, TA_ERR_INVALID, "masses are required with n_fields == 2");
    if (precision != TA_PRECISION_FP64 && precision != TA_PRECISION_FP32) return fail(ctx, TA_ERR_INVALID, "bad precision");
    if (ctx->begun && ctx->T == T && ctx->N == N && ctx->D == D && ctx->src_dtype == src_dtype &&
        ctx->n_fields == n_fields && ctx->precision == precision) {
        // same shape as the previous run on this context: keep every buffer
        // (the pad region of the series is never written, so it is still zero)
        int rc0 = sync_all(ctx);
        if (rc0) return rc0;
        for (int i = 0; i < 3; ++i) ctx->dims[i] = (i < D) ? dims[i] : 0;
        ctx->precision = precision;
        ctx->frames_staged = 0;
        ctx->cur_slab = -1;
        for (auto& s : ctx->sh) s.chunks_pending = false;
        for (auto& s : ctx->sh) {
            if (s.natoms == 0 || n_fields != 2) continue;
            CK(cudaSetDevice(s.dev));
            // on the stream K0 runs on, and complete on return (a plain cudaMemcpy is not ordered with it: see upload_vec)
            CK(cudaMemcpyAsync(s.masses, masses + s.atom0, (size_t)s.natoms * sizeof(double), cudaMemcpyHostToDevice, s.s_compute));
            CK(cudaStreamSynchronize(s.s_compute));
        }
        return TA_OK;
    }
    free_problem(ctx);
    ctx->T = T; ctx->N = N; ctx->D = D;
    ctx->DS = D + ((n_fields == 2 && precision == TA_PRECISION_FP64) ? 1 : 0);
    for (int i = 0; i < 3; ++i) ctx->dims[i] = (i < D) ? dims[i] : 0;
    ctx->src_dtype = src_dtype;
    ctx->elt = (src_dtype == TA_DTYPE_F32) ? 4 : 8;
    ctx->n_fields = n_fields;
    ctx->precision = precision;
    ctx->Tld = ((T + 15) / 16) * 16;

    // pinned slab ring: ~32 MB per slab, at most 512 frames
    const size_t frame_bytes = (size_t)n_fields * (size_t)std::max<int64_t>(N, 1) * 3 * ctx->elt;
    int64_t F = (int64_t)((32u << 20) / frame_bytes);
    F = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(F, 512), T));
    ctx->slab_frames = F;

    const int ndev = (int)ctx->sh.size();
    const int64_t base = N / ndev, rem = N % ndev;
    int64_t a0 = 0;
    for (int i = 0; i < ndev; ++i) {
        Shard& s = ctx->sh[i];
        s.atom0 = a0;
        s.natoms = base + (i < rem ? 1 : 0);
        a0 += s.natoms;
        s.stage_toggle = 0;
        CK(cudaSetDevice(s.dev));
        CK(cudaMalloc(&s.ts_sum, ((size_t)ctx->Tld + 16) * sizeof(double)));
        CK(cudaMalloc(&s.ts_mean, (3 * (size_t)ctx->Tld + 16) * sizeof(double)));   // mean | times | running integral, out[2]
        if (s.natoms == 0) continue;
        // the series are stored in the arithmetic type: double, or float in the FP32 mode
        const size_t ser_bytes = (size_t)s.natoms * ctx->DS * ctx->Tld * (precision == TA_PRECISION_FP64 ? sizeof(double) : sizeof(float));
        const size_t out_bytes = (size_t)s.natoms * ctx->Tld * sizeof(double);
        cudaError_t e = cudaMalloc(&s.series, ser_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&s.by_particle, out_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            free_problem(ctx);
            return fail(ctx, TA_ERR_NOMEM, "device " + std::to_string(s.dev) + ": cannot allocate " +
                                               std::to_string((ser_bytes + out_bytes) >> 20) + " MiB for series + results");
        }
        CK(cudaMemsetAsync(s.series, 0, ser_bytes, s.s_compute));
        CK(cudaMemsetAsync(s.by_particle, 0, out_bytes, s.s_compute));
        if (n_fields == 2) {
            CK(cudaMalloc(&s.masses, (size_t)s.natoms * sizeof(double)));
            CK(cudaMemcpyAsync(s.masses, masses + s.atom0, (size_t)s.natoms * sizeof(double), cudaMemcpyHostToDevice, s.s_compute));
        }
        for (int b = 0; b < kNumDevStage; ++b)
            for (int f = 0; f < n_fields; ++f)
                CK(cudaMalloc(&s.dstage[b][f], (size_t)F * s.natoms * 3 * ctx->elt));
        // events start "signalled": record them once on an idle stream
        for (int b = 0; b < kNumDevStage; ++b) {
            CK(cudaEventRecord(s.ev_k0_done[b], s.s_compute));
            CK(cudaEventRecord(s.ev_copy_done[b], s.s_copy));
        }
    }
    ctx->begun = true;
    int rc = sync_all(ctx);
    if (rc) return rc;
    // The FFT plan and its tables now, while the copy engine is idle: a compute call that follows ta_stage_bulk is
    // queued behind the trajectory copies chunk by chunk, and a table upload issued then would wait for all of them.
    // (Cheap: ~100 KB of tables.  A T the FFT route cannot plan is reported by the FFT compute calls, not here.)
    if (ensure_fft_plan(ctx) != TA_OK) {
        ctx->plan_T = -1;
        ctx->err.clear();
    }
    return TA_OK;
}

int ta_stage_slot(ta_ctx* ctx, void** host_ptr, int64_t* frames_capacity) {
    if (!ctx || !ctx->begun) return fail(ctx, TA_ERR_INVALID, "ta_stage_begin has not been called");
    if (!host_ptr) return fail(ctx, TA_ERR_INVALID, "host_ptr is null");
    const int i = ctx->next_slab;
    if (!ctx->slab[i]) {
        const size_t bytes = (size_t)ctx->slab_frames * ctx->n_fields * (size_t)std::max<int64_t>(ctx->N, 1) * 3 * ctx->elt;
        CK(cudaHostAlloc(&ctx->slab[i], bytes, cudaHostAllocPortable));
    } else if (ctx->slab_used[i]) {
        for (auto& s : ctx->sh) {   // H2D copies that read this slab must have finished
            if (s.natoms == 0) continue;
            CK(cudaSetDevice(s.dev));
            CK(cudaEventSynchronize(s.ev_slab[i]));
        }
    }
    ctx->cur_slab = i;
    ctx->next_slab = (i + 1) % kNumSlabs;
    *host_ptr = ctx->slab[i];
    if (frames_capacity) *frames_capacity = ctx->slab_frames;
    return TA_OK;
}

int ta_stage_commit(ta_ctx* ctx, int64_t frame0, int64_t nframes) {
    if (!ctx || !ctx->begun) return fail(ctx, TA_ERR_INVALID, "ta_stage_begin has not been called");
    if (ctx->cur_slab < 0) return fail(ctx, TA_ERR_INVALID, "ta_stage_commit without ta_stage_slot");
    if (nframes < 1 || nframes > ctx->slab_frames || frame0 < 0 || frame0 + nframes > ctx->T)
        return fail(ctx, TA_ERR_INVALID, "frame range outside [0, T) or larger than the slab");
    const size_t row = (size_t)ctx->N * 3 * ctx->elt;
    const char* base[2];
    base[0] = (const char*)ctx->slab[ctx->cur_slab];
    base[1] = base[0] + row;
    const int slab = ctx->cur_slab;
    ctx->slab_used[slab] = true;
    ctx->cur_slab = -1;
    return enqueue_chunk(ctx, base, (size_t)ctx->n_fields * row, frame0, nframes, slab);
}

int ta_stage_bulk(ta_ctx* ctx, const void* const* fields, int64_t src_atoms, int64_t atom_first,
                  int64_t frame_first, int64_t frame_step, int64_t nframes) {
    if (!ctx || !ctx->begun) return fail(ctx, TA_ERR_INVALID, "ta_stage_begin has not been called");
    if (!fields || !fields[0] || (ctx->n_fields == 2 && !fields[1])) return fail(ctx, TA_ERR_INVALID, "field pointer is null");
    if (nframes != ctx->T) return fail(ctx, TA_ERR_INVALID, "ta_stage_bulk must deliver all T frames");
    if (frame_step < 1 || frame_first < 0 || atom_first < 0 || atom_first + ctx->N > src_atoms)
        return fail(ctx, TA_ERR_INVALID, "bad frame / atom window");
    const size_t src_row = (size_t)src_atoms * 3 * ctx->elt;
    const size_t pitch = (size_t)frame_step * src_row;
    const int64_t T = ctx->T;
    // Particle chunks, each carrying all T frames (a 2-D copy: T rows of chunk*3 values), so that the
    // correlation kernel of a chunk can start as soon as the chunk has landed: ~256 MB per chunk and
    // field, a multiple of 296 particles (two resident CTAs on each of the 148 SMs).
    const int64_t want = std::max<int64_t>(1, ((int64_t)256 << 20) / (T * 3 * (int64_t)ctx->elt));
    const int64_t env_chunk = ctx->opt_bulk_chunk;
    for (auto& s : ctx->sh) {
        if (s.natoms == 0) continue;
        CK(cudaSetDevice(s.dev));
        int64_t CA = env_chunk > 0 ? env_chunk : std::max<int64_t>(296, want / 296 * 296);
        CA = std::min<int64_t>(CA, s.natoms);
        const size_t need = (size_t)T * CA * 3 * ctx->elt;
        if (s.dbulk_bytes < need) {
            for (int b = 0; b < kNumDevStage; ++b)
                for (int f = 0; f < ctx->n_fields; ++f) {
                    cudaFree(s.dbulk[b][f]); s.dbulk[b][f] = nullptr;
                }
            s.dbulk_bytes = 0;
            const int nbuf = (s.natoms > CA) ? kNumDevStage : 1;
            for (int b = 0; b < nbuf; ++b)
                for (int f = 0; f < ctx->n_fields; ++f) CK(cudaMalloc(&s.dbulk[b][f], need));
            s.dbulk_bytes = need;
            for (int b = 0; b < kNumDevStage; ++b) {
                CK(cudaEventRecord(s.ev_bulk_k0[b], s.s_k0));
                CK(cudaEventRecord(s.ev_bulk_copy[b], s.s_copy));
            }
        }
        const size_t nchunks = (size_t)((s.natoms + CA - 1) / CA);
        while (s.chunks.size() > nchunks) { cudaEventDestroy(s.chunks.back().ready); s.chunks.pop_back(); }
        while (s.chunks.size() < nchunks) {
            LaunchRange r;
            CK(cudaEventCreateWithFlags(&r.ready, cudaEventDisableTiming));
            s.chunks.push_back(r);
        }
        for (size_t c = 0; c < nchunks; ++c) {
            const int b = (s.dbulk[1][0] != nullptr) ? (int)(c & 1) : 0;
            const int64_t a0 = (int64_t)c * CA, n = std::min<int64_t>(CA, s.natoms - a0);
            const size_t width = (size_t)n * 3 * ctx->elt;
            CK(cudaStreamWaitEvent(s.s_copy, s.ev_bulk_k0[b], 0));     // K0 has consumed this buffer
            for (int f = 0; f < ctx->n_fields; ++f) {
                const char* src = (const char*)fields[f] + (size_t)frame_first * src_row +
                                  (size_t)(atom_first + s.atom0 + a0) * 3 * ctx->elt;
                CK(cudaMemcpy2DAsync(s.dbulk[b][f], width, src, pitch, width, (size_t)T, cudaMemcpyHostToDevice, s.s_copy));
            }
            CK(cudaEventRecord(s.ev_bulk_copy[b], s.s_copy));
            CK(cudaStreamWaitEvent(s.s_k0, s.ev_bulk_copy[b], 0));
            int rc = launch_k0(ctx, s, s.dbulk[b][0], s.dbulk[b][1], a0, n, T, 0, s.s_k0);
            if (rc) return rc;
            CK(cudaEventRecord(s.ev_bulk_k0[b], s.s_k0));
            s.chunks[c].a0 = a0;
            s.chunks[c].n = n;
            CK(cudaEventRecord(s.chunks[c].ready, s.s_k0));
        }
        s.chunks_pending = true;
    }
    ctx->frames_staged = T;
    return TA_OK;
}

int ta_stage_end(ta_ctx* ctx) {
    if (!ctx || !ctx->begun) return fail(ctx, TA_ERR_INVALID, "ta_stage_begin has not been called");
    return sync_all(ctx);
}

int ta_vacf_fft(ta_ctx* ctx, double* ts_out) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!ts_out) return fail(ctx, TA_ERR_INVALID, "ts_out is null");
    if ((rc = ensure_fft_plan(ctx))) return rc;
    std::vector<int> grids;
    if ((rc = launch_k1(ctx, &grids))) return rc;
    return finish_timeseries(ctx, grids, ts_out);
}

int ta_vacf_windowed(ta_ctx* ctx, double* ts_out) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!ts_out) return fail(ctx, TA_ERR_INVALID, "ts_out is null");
    std::vector<int> grids;
    rc = (ctx->precision == TA_PRECISION_FP64) ? launch_windowed<double, TA_WIN_PRODUCT>(ctx, 1.0, &grids)
                                               : launch_windowed<float, TA_WIN_PRODUCT>(ctx, 1.0, &grids);
    if (rc) return rc;
    return finish_timeseries(ctx, grids, ts_out);
}

int ta_helfand(ta_ctx* ctx, const double* volumes, double boltzmann, double temp_avg, double* ts_out) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!ts_out || !volumes) return fail(ctx, TA_ERR_INVALID, "volumes / ts_out is null");
    if (ctx->n_fields != 2) return fail(ctx, TA_ERR_INVALID, "ta_helfand needs velocities and positions (n_fields == 2)");
    double vsum = 0.0;
    for (int64_t i = 0; i < ctx->T; ++i) vsum += volumes[i];
    const double vol_avg = vsum / (double)ctx->T;                 // viscosity.py:205
    const double denom = 2 * boltzmann * vol_avg * temp_avg;      // viscosity.py:229-231
    std::vector<int> grids;
    rc = (ctx->precision == TA_PRECISION_FP64) ? launch_windowed<double, TA_WIN_SQDIFF>(ctx, denom, &grids)
                                               : launch_windowed<float, TA_WIN_SQDIFF>(ctx, denom, &grids);
    if (rc) return rc;
    return finish_timeseries(ctx, grids, ts_out);
}

int ta_helfand_fft(ta_ctx* ctx, const double* volumes, double boltzmann, double temp_avg, double* ts_out) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!ts_out || !volumes) return fail(ctx, TA_ERR_INVALID, "volumes / ts_out is null");
    if (ctx->n_fields != 2) return fail(ctx, TA_ERR_INVALID, "ta_helfand_fft needs velocities and positions (n_fields == 2)");
    if (ctx->precision != TA_PRECISION_FP64) return fail(ctx, TA_ERR_UNSUPPORTED, "the FFT route of the Helfand MSD is FP64 only");
    double vsum = 0.0;
    for (int64_t i = 0; i < ctx->T; ++i) vsum += volumes[i];
    const double denom = 2 * boltzmann * (vsum / (double)ctx->T) * temp_avg;   // viscosity.py:205, :229-231
    // K5 / K6 keep T + 1 prefix sums in shared memory: refuse BEFORE the FFT pass runs, so that a caller that falls
    // back to ta_helfand has not paid for it
    for (auto& s : ctx->sh)
        if (s.natoms > 0 && ((size_t)ctx->Tld * sizeof(double) + 320 > (size_t)s.max_smem || ctx->T > 32768 || ctx->DS != ctx->D + 1))
            return fail(ctx, TA_ERR_UNSUPPORTED, "FFT Helfand route: T=" + std::to_string(ctx->T) + " does not fit shared memory");
    if ((rc = ensure_fft_plan(ctx))) return rc;
    std::vector<int> grids;
    if ((rc = launch_k1(ctx, &grids, /*with_partial=*/false))) return rc;   // by_particle = sum_d acf_d; K5 forms the particle sums
    // K5 forms S1 - 2 S2 and lists the lags whose result is not good to 1e-10 (cancellation); K6 evaluates those exactly.
    // thr = C eps / tol: C = 100 + T / 100 bounds the error constant of S1 - 2 S2 (FFT autocorrelation + prefix sums;
    // measured 9 - 20 on random, random-walk, ramp, spike, offset and piecewise-constant moments, 18 - 33 on smooth ones:
    // scripts/helfand_fft_error_constant.py -> profiles/r02_helfand_fft_error_constant.txt), tol = 2e-11 leaves a factor 5
    // to the 1e-10 bar on top of that.  If more than 2 % of a shard's (particle, lag) pairs are listed -- or a warp's list
    // overflows -- the direct kernel K3 does the whole call instead.
    const double thr = ctx->opt_helfand_thr >= -1.5 ? ctx->opt_helfand_thr
                                                    : (100.0 + (double)ctx->T / 100.0) * 1.1102230246251565e-16 / 2e-11;
    const size_t row_bytes = (size_t)ctx->Tld * sizeof(double);
    std::vector<unsigned long long> totals(2 * ctx->sh.size(), 0);
    std::vector<HelfandFftArgs> hargs(ctx->sh.size());
    for (size_t i = 0; i < ctx->sh.size(); ++i) {
        Shard& s = ctx->sh[i];
        if (s.natoms == 0) continue;
        CK(cudaSetDevice(s.dev));
        // one q buffer per CTA; two CTAs per SM where two rows fit (T <= ~14,000): they cover each other's load latencies
        // better than one CTA with the next row prefetched into a second buffer (7.3 against 11.0 ms at 125,000 x 10,000)
        const bool two = 2 * (row_bytes + 320) <= (size_t)s.max_smem;
        const int nbuf = 1;
        const size_t smem = (size_t)nbuf * row_bytes;
        auto k5 = two ? k5_helfand_fft_finish<2> : k5_helfand_fft_finish<1>;
        CK(cudaFuncSetAttribute(k5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k5, K5_THREADS, smem));
        if (occ < 1) return fail(ctx, TA_ERR_UNSUPPORTED, "K5 does not fit on an SM");
        // one grid for K5 and K6: K6 walks the lists K5's warps wrote and corrects the per-CTA sum rows K5 formed
        const int grid = (int)std::min<int64_t>(s.natoms, (int64_t)s.num_sms * occ);
        grids[i] = grid;
        if ((rc = ensure_partial(ctx, s, (size_t)grid))) return rc;    // zeroes the rows (K1's sums of the ACF are not wanted)
        // lists: room for twice the 2 % of pairs at which K3 takes over, shared evenly by the grid's warps
        const size_t nlists = (size_t)grid * K5_WARPS;
        // ... and for up to four particles of a CTA that are marked at every lag (a spike, a particle at rest among moving ones)
        const size_t per_cta = (size_t)((s.natoms + grid - 1) / grid);
        const size_t cap = std::max((size_t)(0.04 * (double)s.natoms * (double)ctx->T / (double)nlists) + 64,
                                    ((size_t)ctx->T / K5_WARPS + 1) * std::min<size_t>(per_cta, 4));
        const size_t lbytes = nlists * cap * sizeof(unsigned long long) + nlists * sizeof(unsigned) + 64;
        if (s.hflags_bytes < lbytes) {
            cudaFree(s.hflags);
            s.hflags = nullptr; s.hflags_bytes = 0;
            CK(cudaMalloc((void**)&s.hflags, lbytes));
            s.hflags_bytes = lbytes;
        }
        HelfandFftArgs& a = hargs[i];
        a.series = (const double*)s.series; a.by_particle = s.by_particle; a.partial = s.partial;
        a.natoms = (int)s.natoms; a.D = ctx->D; a.DS = ctx->DS; a.T = (int)ctx->T; a.Tld = ctx->Tld; a.denom = denom;
        unsigned char* base = reinterpret_cast<unsigned char*>(s.hflags);
        a.total = reinterpret_cast<unsigned long long*>(base);                       // [2], then the lists, then the counts
        a.list = reinterpret_cast<unsigned long long*>(base + 64);
        a.count = reinterpret_cast<unsigned*>(base + 64 + nlists * cap * sizeof(unsigned long long));
        a.cap = (unsigned)cap; a.thr = thr; a.nbuf = nbuf;
        CK(cudaMemsetAsync(a.total, 0, 2 * sizeof(unsigned long long), s.s_compute));
        k5<<<grid, K5_THREADS, smem, s.s_compute>>>(a);
        CK(cudaGetLastError());
        ctx->launches++;
        CK(cudaEventRecord(s.ev_kb, s.s_compute));                   // ta_last_kernel_ms: K1 + K5 (+ K6 below)
        CK(cudaMemcpyAsync(&totals[2 * i], a.total, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.s_compute));
    }
    bool need_k3 = false;
    ctx->helfand_fft_flagged = 0;
    for (size_t i = 0; i < ctx->sh.size(); ++i) {
        Shard& s = ctx->sh[i];
        if (s.natoms == 0) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_compute));
        ctx->helfand_fft_flagged += (long long)totals[2 * i];
        if ((double)totals[2 * i] > 0.02 * (double)s.natoms * (double)ctx->T || totals[2 * i + 1] > 0) need_k3 = true;
    }
    if (need_k3) {
        // too much of the result needs the exact evaluation (series that barely move): the direct kernel does it all
        ctx->helfand_fft_flagged = -1;
        rc = launch_windowed<double, TA_WIN_SQDIFF>(ctx, denom, &grids);
        if (rc) return rc;
        return finish_timeseries(ctx, grids, ts_out);
    }
    for (size_t i = 0; i < ctx->sh.size(); ++i) {
        Shard& s = ctx->sh[i];
        if (s.natoms == 0 || totals[2 * i] == 0) continue;           // nothing to correct on this shard
        CK(cudaSetDevice(s.dev));
        k6_helfand_refine<<<grids[i], K5_THREADS, 0, s.s_compute>>>(hargs[i]);
        CK(cudaGetLastError());
        ctx->launches++;
        CK(cudaEventRecord(s.ev_kb, s.s_compute));                   // ta_last_kernel_ms: K1 + K5 + K6
    }
    return finish_timeseries(ctx, grids, ts_out);
}

int ta_fetch_by_particle(ta_ctx* ctx, int64_t atom0, int64_t natoms, int layout, double* out) {
    if (!ctx || !ctx->begun) return fail(ctx, TA_ERR_INVALID, "ta_stage_begin has not been called");
    if (!out || atom0 < 0 || natoms < 0 || atom0 + natoms > ctx->N) return fail(ctx, TA_ERR_INVALID, "bad particle range");
    if (layout != TA_LAYOUT_ATOM_MAJOR && layout != TA_LAYOUT_LAG_MAJOR) return fail(ctx, TA_ERR_INVALID, "bad layout");
    const int64_t T = ctx->T, Tld = ctx->Tld;
    for (auto& s : ctx->sh) {
        const int64_t lo = std::max(atom0, s.atom0), hi = std::min(atom0 + natoms, s.atom0 + s.natoms);
        if (lo >= hi) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_compute));
        if (layout == TA_LAYOUT_ATOM_MAJOR) {
            CK(cudaMemcpy2D(out + (size_t)(lo - atom0) * T, (size_t)T * 8, s.by_particle + (size_t)(lo - s.atom0) * Tld,
                            (size_t)Tld * 8, (size_t)T * 8, (size_t)(hi - lo), cudaMemcpyDeviceToHost));
        } else {
            const int64_t chunk = std::max<int64_t>(32, ((int64_t)(64u << 20) / (T * 8)) / 32 * 32);
            const size_t need = (size_t)std::min<int64_t>(chunk, hi - lo) * T * 8;
            if (s.lagmajor_bytes < need) {
                cudaFree(s.lagmajor_tmp);
                s.lagmajor_tmp = nullptr; s.lagmajor_bytes = 0;
                CK(cudaMalloc(&s.lagmajor_tmp, need));
                s.lagmajor_bytes = need;
            }
            for (int64_t c0 = lo; c0 < hi; c0 += chunk) {
                const int64_t cn = std::min<int64_t>(chunk, hi - c0);
                dim3 block(32, 8), grid((unsigned)((T + 31) / 32), (unsigned)((cn + 31) / 32));
                k_to_lag_major<<<grid, block, 0, s.s_compute>>>(s.by_particle, Tld, (int)T, c0 - s.atom0, (int)cn, s.lagmajor_tmp);
                CK(cudaGetLastError());
                ctx->launches++;
                CK(cudaMemcpy2DAsync(out + (size_t)(c0 - atom0), (size_t)natoms * 8, s.lagmajor_tmp, (size_t)cn * 8,
                                     (size_t)cn * 8, (size_t)T, cudaMemcpyDeviceToHost, s.s_compute));
                CK(cudaStreamSynchronize(s.s_compute));
            }
        }
    }
    return TA_OK;
}

int ta_green_kubo(ta_ctx* ctx, const double* times, int64_t start, int64_t stop, int64_t step, double initial,
                  double* integral, double* running, double* slope) {
    if (!ctx || !ctx->begun || !ctx->have_mean) return fail(ctx, TA_ERR_INVALID, "ta_green_kubo needs the timeseries of a compute call");
    if (!times) return fail(ctx, TA_ERR_INVALID, "times is null");
    const int64_t T = ctx->T;
    if (step < 1 || start < 0 || stop > T || start > stop) return fail(ctx, TA_ERR_INVALID, "bad window");
    const int64_t n = (stop - start + step - 1) / step;
    Shard& s = ctx->sh[0];
    CK(cudaSetDevice(s.dev));
    double* d_times = s.ts_mean + ctx->Tld;
    double* d_run = d_times + ctx->Tld;
    double* d_out = d_run + ctx->Tld;
    CK(cudaMemcpyAsync(d_times, times, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, s.s_compute));
    k7_green_kubo<<<1, K7_THREADS, 0, s.s_compute>>>(s.ts_mean, d_times, start, step, (int)n, initial, running ? d_run : nullptr, d_out);
    CK(cudaGetLastError());
    ctx->launches++;
    double out[2] = {0.0, 0.0};
    CK(cudaMemcpyAsync(out, d_out, sizeof(out), cudaMemcpyDeviceToHost, s.s_compute));
    if (running && n > 0) CK(cudaMemcpyAsync(running, d_run, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s.s_compute));
    CK(cudaStreamSynchronize(s.s_compute));
    if (integral) *integral = out[0];
    if (slope) *slope = out[1];
    return TA_OK;
}

int ta_timer_begin(ta_ctx* ctx) {
    if (!ctx) return fail(nullptr, TA_ERR_INVALID, "null context");
    for (auto& s : ctx->sh) {
        CK(cudaSetDevice(s.dev));
        CK(cudaEventRecord(s.ev_t0, s.s_compute));
    }
    return TA_OK;
}

int ta_timer_end(ta_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ctx, TA_ERR_INVALID, "null argument");
    float worst = 0.f;
    for (auto& s : ctx->sh) {
        CK(cudaSetDevice(s.dev));
        CK(cudaStreamSynchronize(s.s_copy));
        CK(cudaEventRecord(s.ev_t1, s.s_compute));
        CK(cudaEventSynchronize(s.ev_t1));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, s.ev_t0, s.ev_t1));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return TA_OK;
}

int ta_probe_fp64(ta_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return fail(ctx, TA_ERR_INVALID, "null argument");
    Shard& s = ctx->sh[0];
    CK(cudaSetDevice(s.dev));
    const int grid = s.num_sms * 4, iters = 20000;           // 8 warps per sub-partition
    double* out = nullptr;
    CK(cudaMalloc(&out, (size_t)grid * KP_THREADS * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {                        // the first launch warms up
        CK(cudaEventRecord(e0, s.s_compute));
        k_probe_fp64<<<grid, KP_THREADS, 0, s.s_compute>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1, s.s_compute));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CK(cudaFree(out));
    const double flop = 2.0 * (double)grid * KP_THREADS * (double)iters * KP_UNROLL * KP_CHAINS;
    *tflops = flop / ((double)best * 1e-3) / 1e12;
    return TA_OK;
}

int ta_probe_h2d(ta_ctx* ctx, uint64_t bytes, double* gbps) {
    if (!ctx || !gbps || bytes == 0) return fail(ctx, TA_ERR_INVALID, "null argument");
    Shard& s = ctx->sh[0];
    CK(cudaSetDevice(s.dev));
    void *h = nullptr, *d = nullptr;
    CK(cudaHostAlloc(&h, (size_t)bytes, cudaHostAllocPortable));
    memset(h, 1, (size_t)bytes);
    cudaError_t e = cudaMalloc(&d, (size_t)bytes);
    if (e != cudaSuccess) { cudaFreeHost(h); return fail(ctx, TA_ERR_NOMEM, "ta_probe_h2d: cannot allocate the device buffer"); }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaMemcpyAsync(d, h, (size_t)bytes, cudaMemcpyHostToDevice, s.s_copy));    // warm-up
    CK(cudaStreamSynchronize(s.s_copy));
    // twelve back-to-back copies: long enough (0.2 - 0.6 s at 1 GiB) that ranks probing at the same time overlap for
    // nearly the whole timed region, whatever the skew of their allocations -- a SUSTAINED concurrent rate, which is
    // what the staging path of a multi-rank run gets (three copies per rank mostly ran one rank after the other and
    // reported 40 - 55 GB/s per rank on a host that sustains 23 GB/s per rank with eight ranks copying)
    const int reps = 12;
    CK(cudaEventRecord(e0, s.s_copy));
    for (int rep = 0; rep < reps; ++rep) CK(cudaMemcpyAsync(d, h, (size_t)bytes, cudaMemcpyHostToDevice, s.s_copy));
    CK(cudaEventRecord(e1, s.s_copy));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d); cudaFreeHost(h);
    *gbps = (double)reps * (double)bytes / ((double)ms * 1e-3) / 1e9;
    return TA_OK;
}

int ta_flush_l2(ta_ctx* ctx) {
    if (!ctx) return fail(nullptr, TA_ERR_INVALID, "null context");
    const size_t bytes = (size_t)512 << 20;   // 4x the 126 MB L2
    for (auto& s : ctx->sh) {
        CK(cudaSetDevice(s.dev));
        if (!s.l2_scratch) CK(cudaMalloc(&s.l2_scratch, bytes));
        CK(cudaMemsetAsync(s.l2_scratch, 0x5a, bytes, s.s_compute));
        CK(cudaStreamSynchronize(s.s_compute));
    }
    return TA_OK;
}

int ta_last_kernel_ms(ta_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ctx, TA_ERR_INVALID, "null argument");
    float worst = 0.f;
    for (auto& s : ctx->sh) {
        if (!s.kernel_timed) continue;
        CK(cudaSetDevice(s.dev));
        CK(cudaEventSynchronize(s.ev_kb));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, s.ev_ka, s.ev_kb));
        worst = std::max(worst, t);
    }
    *ms = worst;
    return TA_OK;
}

int64_t ta_launch_count(const ta_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t ta_helfand_fft_refined(const ta_ctx* ctx) { return ctx ? ctx->helfand_fft_flagged : 0; }

int ta_fft_plan_info(const ta_ctx* ctx, int* H, int* npasses, int* radices, int* threads, int* smem_bytes, int* grid) {
    if (!ctx) return TA_ERR_INVALID;
    const bool have = ctx->plan_T >= 0;
    if (H) *H = have ? ctx->plan.H : 0;
    if (npasses) *npasses = have ? ctx->plan.npasses : 0;
    if (radices) for (int i = 0; i < TA_MAX_PASSES; ++i) radices[i] = have ? ctx->plan.radix[i] : 0;
    if (threads) *threads = ctx->k1_threads;
    if (smem_bytes) *smem_bytes = ctx->k1_smem;
    if (grid) *grid = ctx->k1_grid;
    return TA_OK;
}

int ta_k1_uses_tmem(const ta_ctx* ctx) { return ctx && ctx->k1_tmem ? 1 : 0; }

}  // extern "C"

// sm_100a kernels of the time-correlation hot path.  The per-CTA arithmetic
// lives in fft_core.cuh / windowed_core.cuh; this file wraps it in persistent
// CTAs (one particle at a time, static round-robin so results are
// bit-reproducible run to run) and adds the staging transposition and the
// small reduction / layout kernels.
#pragma once
#include <cuda_runtime.h>
#include "ta_common.cuh"
#include "fft_core.cuh"
#include "windowed_core.cuh"
#include "k1_fast.cuh"

namespace ta {

// ---------------------------------------------------------------------------
// K0: staging transposition.  Frame-major slab [nframes][natoms][3] (f32 or
// f64, as copied from the host) -> atom-major, time-contiguous series
// [natoms][D][Tld] f64 at frame offset frame0, selecting the dim_type columns.
// Replaces the per-frame copies of VelocityAutocorr._single_frame
// (velocityautocorr.py:192-194) and ViscosityHelfand._single_frame
// (viscosity.py:189-199); with HELFAND it also forms g = (m*v)*x, the product
// viscosity.py:213-218 evaluates per lag (same association, so bit-identical).
// ---------------------------------------------------------------------------
constexpr int K0_FR = 32;   // frames per tile
constexpr int K0_AT = 32;   // atoms per tile

template <typename SRC, bool HELFAND, typename OUT>
__global__ void __launch_bounds__(256)
k0_stage(const SRC* __restrict__ v, const SRC* __restrict__ x, const double* __restrict__ masses,
         OUT* __restrict__ series, int natoms, int nframes, long long frame0, long long Tld,
         int D, int DS, int d0, int d1, int d2) {
    // frames on grid.x (2^31 - 1 tiles: any T the library accepts), particles on grid.y, tiled further by the loop below.
    // A particle owns DS rows of Tld values: its D selected columns and, when DS > D (Helfand in FP64), a row of
    // q[i] = sum_d g_d[i]^2 -- what the O(T log T) route's finishing kernel K5 prefix-sums (it then reads 8 bytes per
    // atom-frame instead of the 8 D of the series).
    __shared__ OUT tile[K0_FR][K0_AT * 3 + 1];
    const int f0 = blockIdx.x * K0_FR;
    const int nf = min(K0_FR, nframes - f0);
    const int dims[3] = {d0, d1, d2};
    for (int a0 = blockIdx.y * K0_AT; a0 < natoms; a0 += gridDim.y * K0_AT) {
        const int na = min(K0_AT, natoms - a0);
        const int ncol = na * 3;
        for (int f = threadIdx.y; f < nf; f += blockDim.y) {
            const size_t rowoff = ((size_t)(f0 + f) * natoms + a0) * 3;
            for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
                double val = (double)v[rowoff + c];
                if (HELFAND) {
                    double m = masses[a0 + c / 3];
                    val = (m * val) * (double)x[rowoff + c];
                }
                tile[f][c] = (OUT)val;
            }
        }
        __syncthreads();
        const int nrows = na * DS;
        for (int r = threadIdx.y; r < nrows; r += blockDim.y) {
            const int a = r / DS, d = r - a * DS;
            OUT* dst = series + ((size_t)(a0 + a) * DS + d) * Tld + frame0 + f0;
            if (d < D) {
                const int col = a * 3 + dims[d];
                for (int f = threadIdx.x; f < nf; f += blockDim.x) dst[f] = tile[f][col];
            } else {
                for (int f = threadIdx.x; f < nf; f += blockDim.x) {
                    double q = 0.0;
                    for (int e = 0; e < D; ++e) { const double g = (double)tile[f][a * 3 + dims[e]]; q = fma(g, g, q); }
                    dst[f] = (OUT)q;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// K1: FFT autocorrelation, one particle per CTA iteration.
// ---------------------------------------------------------------------------
template <typename R>
struct K1Args {
    FftTables<R> t;              // tw_lo / tw_hi point to GLOBAL copies here
    int nlo, nhi;
    const R* series;             // [natoms][DS][Tld]  (stored in the arithmetic type; the first D rows are transformed)
    double* by_particle;         // [natoms][Tld]
    double* partial;             // [gridDim.x][Tld]
    int natoms, D, DS;
    long long Tld;
    unsigned char* scratch;      // SCRATCH: per-CTA work area in global memory (FFT buffer + pair accumulators)
    long long scratch_stride;    // bytes per CTA
};

constexpr int K1_MAX_THREADS = 640;

// SCRATCH = false: the H-point buffer and the pair accumulators live in shared memory (H up to ~9,600 in FP64).
// SCRATCH = true : they live in a per-CTA global work area (served by L2), the twiddle tables stay in shared memory:
//                  the same passes, slower, for any T -- the FFT route never has to refuse a trajectory for its length.
template <typename R, bool SCRATCH = false>
__global__ void __launch_bounds__(K1_MAX_THREADS)
k1_fft_acf(const K1Args<R> args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = args.t.H;
    unsigned char* work = SCRATCH ? args.scratch + (size_t)blockIdx.x * (size_t)args.scratch_stride : smem_raw;
    cplx<R>* buf = reinterpret_cast<cplx<R>*>(work);
    cplx<R>* tw_lo = SCRATCH ? reinterpret_cast<cplx<R>*>(smem_raw) : buf + H;
    cplx<R>* tw_hi = tw_lo + args.nlo;
    R* sd = SCRATCH ? reinterpret_cast<R*>(buf + H) : reinterpret_cast<R*>(tw_hi + args.nhi);

    const int tid = threadIdx.x, nthr = blockDim.x;
    FftTables<R> t = args.t;
    for (int i = tid; i < args.nlo; i += nthr) tw_lo[i] = args.t.tw_lo[i];
    for (int i = tid; i < args.nhi; i += nthr) tw_hi[i] = args.t.tw_hi[i];
    t.tw_lo = tw_lo;
    t.tw_hi = tw_hi;
    __syncthreads();

    double* partial = args.partial + (size_t)blockIdx.x * args.Tld;
    for (int a = blockIdx.x; a < args.natoms; a += gridDim.x) {
        const R* ser = args.series + (size_t)a * args.DS * args.Tld;
        double* row = args.by_particle + (size_t)a * args.Tld;
        for (int r = 0; r < 2; ++r) {
            fft_zero_acc<R>(tid, nthr, sd, t);
            for (int d = 0; d < args.D; ++d) {
                fft_load<R>(tid, nthr, buf, ser + (size_t)d * args.Tld, t, r);
                __syncthreads();
                int size = H;
                for (int ps = 0; ps < t.npasses; ++ps) {
                    const int s = size / t.radix[ps];
                    dif_pass_any<R>(t.radix[ps], tid, nthr, buf, t, s);
                    __syncthreads();
                    size = s;
                }
                fft_accumulate<R>(tid, nthr, buf, sd, t, r);
                __syncthreads();
            }
            fft_build<R>(tid, nthr, buf, sd, t, r);
            __syncthreads();
            int s = 1;
            for (int ps = t.npasses - 1; ps >= 0; --ps) {
                dit_pass_any<R>(t.radix[ps], tid, nthr, buf, t, s);
                __syncthreads();
                s *= t.radix[ps];
            }
            fft_store<R>(tid, nthr, buf, row, partial, t, r);
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// K1 fast path (k1_fast.cuh): H = 256 R1 in three register-DFT passes.
// DevCtx: the device side of the `Ctx` policy the kernel bodies are written against (tests/emu has the CPU side).
// ---------------------------------------------------------------------------
struct DevCtx {
    static TA_HD void sync() {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
    static TA_HD void sync_warp() {
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
    template <class V> static TA_HD V shfl_xor(V v, int mask) {
#if defined(__CUDA_ARCH__)
        return __shfl_xor_sync(0xffffffffu, v, mask);
#else
        return v;
#endif
    }
    template <class V> static TA_HD V shfl_xor16(V v) {
#if defined(__CUDA_ARCH__)
        return __shfl_xor_sync(0xffffffffu, v, 16);
#else
        return v;
#endif
    }
    // streaming load: through L2 only, so that L1 keeps the tables every particle re-reads
    static TA_HD cd ld_stream(const cd* p) {
#if defined(__CUDA_ARCH__)
        const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
        return cmake<double>(v.x, v.y);
#else
        return *p;
#endif
    }
    static TA_HD cf ld_stream(const cf* p) {
#if defined(__CUDA_ARCH__)
        const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
        return cmake<float>(v.x, v.y);
#else
        return *p;
#endif
    }
    static TA_HD void compiler_fence() {
#if defined(__CUDA_ARCH__)
        asm volatile("" ::: "memory");
#endif
    }
    // mbarrier with one arrival per phase (the thread that issues the bulk load)
    static TA_HD void mbar_init(unsigned long long* bar) {
#if defined(__CUDA_ARCH__)
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    // dst[0..bytes) (shared) <- src (global) by the bulk-copy engine (TMA); completion flips the phase of `bar`
    static TA_HD void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
#if defined(__CUDA_ARCH__)
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(src), "r"(bytes), "r"(a) : "memory");
#endif
    }
    // ---- tensor memory as a per-thread scratchpad (k1f_body<..., TMEM = true>)
    // 512 columns for the CTA: warp 0 allocates, everybody learns the base address.  Called by ALL threads.
    static TA_HD uint32_t tmem_alloc(uint32_t* slot, int tid) {
#if defined(__CUDA_ARCH__)
        if ((tid >> 5) == 0) {
            const unsigned a = (unsigned)__cvta_generic_to_shared(slot);
            const unsigned ncols = 512u;
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        return *reinterpret_cast<volatile uint32_t*>(slot);
#else
        return 0;
#endif
    }
    static TA_HD void tmem_free(uint32_t base, int tid) {
#if defined(__CUDA_ARCH__)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if ((tid >> 5) == 0) {
            const unsigned ncols = 512u;
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
        }
#endif
    }
    // four complex doubles = 16 columns of the calling thread's own lane (whole warps only: .sync.aligned)
    static TA_HD void tmem_st4(uint32_t taddr, const cd (&v)[4]) {
#if defined(__CUDA_ARCH__)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     ::"r"(taddr),
                       "r"(__double2loint(v[0].x)), "r"(__double2hiint(v[0].x)), "r"(__double2loint(v[0].y)), "r"(__double2hiint(v[0].y)),
                       "r"(__double2loint(v[1].x)), "r"(__double2hiint(v[1].x)), "r"(__double2loint(v[1].y)), "r"(__double2hiint(v[1].y)),
                       "r"(__double2loint(v[2].x)), "r"(__double2hiint(v[2].x)), "r"(__double2loint(v[2].y)), "r"(__double2hiint(v[2].y)),
                       "r"(__double2loint(v[3].x)), "r"(__double2hiint(v[3].x)), "r"(__double2loint(v[3].y)), "r"(__double2hiint(v[3].y))
                     : "memory");
#endif
    }
    // load + wait in one statement: the registers are valid when it returns
    static TA_HD void tmem_ld4(uint32_t taddr, cd (&v)[4]) {
#if defined(__CUDA_ARCH__)
        int w[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                       "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                     : "r"(taddr) : "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = cmake<double>(__hiloint2double(w[4 * i + 1], w[4 * i]), __hiloint2double(w[4 * i + 3], w[4 * i + 2]));
#endif
    }
    static TA_HD void tmem_wait_st() {
#if defined(__CUDA_ARCH__)
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
    }
    static TA_HD void mbar_wait(unsigned long long* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
        unsigned done = 0, tries = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(a), "r"(parity) : "memory");
            if (!done && ++tries > (1u << 22)) __trap();          // a bulk copy that never lands is a bug: fail, do not hang
        }
#endif
    }
};

// One kernel per (R1, arithmetic type).  NT = 16 R1 threads; the bulk series prefetch where k1f_prefetch says so.
template <int R1, typename RT, bool PART = true, bool TMEM = false>
__global__ void __launch_bounds__(k1f_threads(R1), k1f_min_blocks(k1f_threads(R1), (int)sizeof(RT)))
k1f_fft_acf(const K1FArgs<RT> args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    k1f_body<R1, k1f_threads(R1), DevCtx, RT, k1f_prefetch(R1, (int)sizeof(RT)), PART, TMEM>(args, smem_raw, (int)threadIdx.x,
                                                                                             (int)blockIdx.x, (int)gridDim.x);
}

// The FP64 kernel of ten warps (R1 = 20) with the register cap stated outright instead of derived from launch bounds.
// With three warps on two of the sub-partitions (16,384 registers each) 170 registers per thread is the ceiling either
// way, but ptxas schedules differently under the two attributes: __launch_bounds__(320, 1) gives 168 registers and
// 16-32 B of spills, __maxnreg__(184) 166 registers and none -- 24.2 -> 23.5 ms at 100k x 10k (caps 152 / 160 / 168 /
// 176 / 184: 24.21 / 24.08 / 23.66 / 23.52 / 23.46 ms; 200 does not fit).
constexpr int K1F_MAXREG = 184;
template <int R1, typename RT, bool PART = true, bool TMEM = false>
__global__ void __maxnreg__(TMEM ? 168 : K1F_MAXREG)       // 168: the most three warps of a sub-partition can have
k1f_fft_acf_mr(const K1FArgs<RT> args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    k1f_body<R1, k1f_threads(R1), DevCtx, RT, k1f_prefetch(R1, (int)sizeof(RT)), PART, TMEM>(args, smem_raw, (int)threadIdx.x,
                                                                                             (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------
// K2 / K3: windowed lag sums, one particle per CTA iteration.
//   MODE = TA_WIN_PRODUCT: vacf[k]  = sum_d sum_i g_d[i] g_d[i+k] / (T-k)
//   MODE = TA_WIN_SQDIFF : visc[k]  = sum_d sum_i (g_d[i]-g_d[i+k])^2 / (D (T-k)) / denom
// ---------------------------------------------------------------------------
constexpr int KW_MAX_THREADS = 512;

template <typename R, int MODE, bool SCRATCH = false>
__global__ void __launch_bounds__(KW_MAX_THREADS)
k_windowed(const WinArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    win_body<R, MODE, DevCtx, SCRATCH>(args, smem_raw, (int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------
// K5: finishing pass of the opt-in FFT route for the Helfand MSD
// (docs/tutorials/helfand_dev_toy_system.ipynb:134 "consider whether fft is possible later";
// same quantity as ViscosityHelfand._conclude, viscosity.py:201-233):
//   sum_i (g[i] - g[i+k])^2 = S1[k] - 2 S2[k],
//   S2[k] = sum_i g[i] g[i+k]                      (K1 left  sum_d S2_d[k] / (T-k)  in by_particle)
//   S1[k] = sum_{i<T-k} q[i] + sum_{i>=k} q[i],    q[i] = sum_d g_d[i]^2   (the extra series row K0 wrote)
// The difference cancels: both terms carry rounding errors of C eps sum g^2 (C = 9 - 33 measured,
// profiles/r02_helfand_fft_error_constant.txt), so a lag whose un-normalised MSD is below thr * sum g^2 is NOT good to
// 1e-10 and goes on a list that K6 evaluates with the reference's own sum.
//
// One CTA of 32 warps per SM, persistent over particles b, b + grid, ...  Per particle: the q row arrives by the
// bulk-copy engine (TMA) into one of two shared buffers -- the row of the NEXT particle is in flight while this one is
// worked on --, each warp scans its contiguous 1/32 of the row with shuffle scans (no bank conflicts), one barrier,
// then thread t finishes lags t, t + 1024, ...: two prefix-sum reads, the row K1 left, the value, the test.  The
// lag k is always handled by thread k % 1024 (warp (k / 32) % 32), which alone updates entry k of the CTA's
// particle-sum row, and each warp appends the lags it marks to ITS OWN list (a counter in a register: no atomics, and the
// order -- hence every later floating-point sum -- is the same in every run).  HBM traffic: 8 (q) + 8 (row in) +
// 8 (row out) bytes per atom-frame.
// ---------------------------------------------------------------------------
struct HelfandFftArgs {
    const double* series;   // [natoms][DS][Tld]; row D of a particle is q
    double* by_particle;    // [natoms][Tld]  in: sum_d acf_d ; out: viscosity function
    double* partial;        // [gridDim.x][Tld]   particle sums of this CTA's rows (K5 writes, K6 corrects)
    int natoms, D, DS, T;
    long long Tld;
    double denom;           // 2 kB <V> temp_avg
    unsigned long long* list;   // [gridDim.x * 32][cap]  (particle << 32 | lag) pairs that need the exact evaluation, per warp
    unsigned* count;            // [gridDim.x * 32]       entries of each list (may exceed cap: the excess was dropped -> overflow)
    unsigned long long* total;  // [2]  sum of all counts; number of lists that overflowed
    unsigned cap;
    double thr;             // a lag is marked when its un-normalised MSD is below thr * sum_i sum_d g^2
    int nbuf;               // K5: shared q / prefix-sum buffers (2: next row prefetched; 1: series too long for two)
};

constexpr int K5_THREADS = 512;
constexpr int K5_WARPS = K5_THREADS / 32;

// 1 / a for a > 0: hardware seed + two Newton steps (within 1 ulp; no division sequence)
__device__ __forceinline__ double rcp_pos(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    return fma(y, e, y);
}

// MINB = 2: compiled for two co-resident CTAs (64 registers) with one q buffer each, which cover each other's load
// latencies -- 7.3 ms at 125,000 x 10,000 against 11.0 ms for one CTA with two buffers, whether or not that one has
// every row value in registers before its scan starts (profiles/r02_k5_k6.txt); MINB = 1: rows too long for two CTAs.
template <int MINB>
__global__ void __launch_bounds__(K5_THREADS, MINB)
k5_helfand_fft_finish(const HelfandFftArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = a.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned row_bytes = (unsigned)a.Tld * (unsigned)sizeof(double);          // Tld is a multiple of 16 values
    double* bufs[2] = {reinterpret_cast<double*>(smem_raw), reinterpret_cast<double*>(smem_raw + (a.nbuf > 1 ? row_bytes : 0))};
    __shared__ double wsum[K5_WARPS];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int nj = (T + K5_THREADS - 1) / K5_THREADS;               // lags per thread
    const int R = nj | 1;                                           // values per lane in the scan: odd -> conflict-free 64-bit accesses
    const int C = 32 * R;                                           // values per warp
    const unsigned magicC = (unsigned)((0x100000000ull + (unsigned)C - 1) / (unsigned)C);   // i / C = umulhi(i, magicC) for i < 2^32 / C
    double* partial = a.partial + (size_t)blockIdx.x * a.Tld;       // zeroed by the host; stays L2-resident (one row per CTA)
    unsigned my_count = 0;                                          // warp-uniform: entries this warp has appended
    unsigned long long* my_list = a.list + ((size_t)blockIdx.x * K5_WARPS + warp) * a.cap;
    const double cscale = 1.0 / ((double)a.D * a.denom);

    if (tid == 0) { DevCtx::mbar_init(&mbar[0]); DevCtx::mbar_init(&mbar[1]); }
    __syncthreads();
    auto qrow = [&](int n) { return a.series + ((size_t)n * a.DS + a.D) * a.Tld; };
    if (tid == 0 && (int)blockIdx.x < a.natoms) DevCtx::bulk_load(bufs[0], qrow(blockIdx.x), row_bytes, &mbar[0]);
    unsigned phase0 = 0u, phase1 = 0u;
    int it = 0;
    for (int n = blockIdx.x; n < a.natoms; n += gridDim.x, ++it) {
        const int b = a.nbuf > 1 ? (it & 1) : 0;
        double* P = bufs[b];                                       // q, then its inclusive prefix sums within each warp's stretch
        double* row = a.by_particle + (size_t)n * a.Tld;
        if (a.nbuf > 1 && tid == 0 && n + (int)gridDim.x < a.natoms)
            DevCtx::bulk_load(bufs[b ^ 1], qrow(n + gridDim.x), row_bytes, &mbar[b ^ 1]);      // its readers left it at the last barrier
        if (b == 0) { DevCtx::mbar_wait(&mbar[0], phase0); phase0 ^= 1u; }
        else { DevCtx::mbar_wait(&mbar[1], phase1); phase1 ^= 1u; }
        // ---- scan: warp w owns values [w C, (w + 1) C), lane l the run [w C + l R, + R)
        const int i0 = warp * C + lane * R;
        double run = 0.0;
        for (int t = 0; t < R; ++t) {
            const int i = i0 + t;
            if (i < T) { run += P[i]; P[i] = run; }
        }
        double incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const double lo = incl - run;                               // sum of the runs of the lanes below
        for (int t = 0; t < R; ++t) {
            const int i = i0 + t;
            if (i < T) P[i] += lo;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        // every warp: exclusive scan of the stretch totals (lane l < 16 holds the offset of stretch l)
        double woff = lane < K5_WARPS ? wsum[lane] : 0.0, tot;
        {
            double inc2 = woff;
#pragma unroll
            for (int o = 1; o < K5_WARPS; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += u;
            }
            tot = __shfl_sync(0xffffffffu, inc2, K5_WARPS - 1);
            woff = inc2 - woff;
        }
        // prefix sum Pj = sum_{i<j} q[i], j = 0 .. T (all lanes call it: the stretch offset travels by shuffle)
        auto prefix = [&](int j) {
            const int i = j > 0 ? j - 1 : 0;
            const double off = __shfl_sync(0xffffffffu, woff, (int)__umulhi((unsigned)i, magicC));
            return j > 0 ? P[i] + off : 0.0;
        };
        const double floor_msd = a.thr * tot;
        // one lag: value, test, list entry (all lanes of the warp call it together)
        auto finish_lag = [&](int k, double rk, double pak) {
            const int kc = k < T ? k : 0;
            const double pk = prefix(kc), ptk = prefix(T - kc);
            bool flag = false;
            if (k < T) {
                double val = 0.0;                                   // lag 0 stays exactly 0 (viscosity.py:207-210)
                if (k > 0) {
                    const double s1 = ptk + (tot - pk);
                    const double nk = (double)(T - k);
                    const double msd = s1 - 2.0 * rk * nk;          // un-normalised, what the threshold is about
                    flag = !(msd >= floor_msd);                     // also catches a NaN
                    val = msd * rcp_pos(nk) * cscale;
                }
                __stcs(row + k, val);
                partial[k] = pak + (flag ? 0.0 : val);              // a marked lag enters the sum with its exact value, in K6
            }
            const unsigned m = __ballot_sync(0xffffffffu, flag);
            if (flag) {
                const unsigned idx = my_count + __popc(m & ((1u << lane) - 1u));
                if (idx < a.cap) my_list[idx] = ((unsigned long long)(unsigned)n << 32) | (unsigned)k;
            }
            my_count += __popc(m);
        };
        for (int j0 = 0; j0 < nj; j0 += 4) {                        // the row K1 left and this CTA's sum row, four lags at a time
            double r4[4], p4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = tid + (j0 + u) * K5_THREADS;
                const bool in = j0 + u < nj && k < T;
                r4[u] = in ? __ldcs(row + k) : 0.0;
                p4[u] = in ? partial[k] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j0 + u < nj) finish_lag(tid + (j0 + u) * K5_THREADS, r4[u], p4[u]);       // uniform over the CTA
        }
        __syncthreads();                                            // P is free: the next-but-one row may land in it
        if (a.nbuf == 1 && tid == 0 && n + (int)gridDim.x < a.natoms)
            DevCtx::bulk_load(bufs[0], qrow(n + gridDim.x), row_bytes, &mbar[0]);
    }
    if (lane == 0) {
        a.count[(size_t)blockIdx.x * K5_WARPS + warp] = my_count;
        if (my_count) atomicAdd(a.total, (unsigned long long)my_count);
        if (my_count > a.cap) atomicAdd(a.total + 1, 1ull);
    }
}

// K6: the marked lags, evaluated with the reference's own sum  sum_d sum_i (g_d[i] - g_d[i+k])^2  (viscosity.py:212-226).
// Same grid as K5.  Phase A: the entries of all the CTA's lists are dealt to all its warps (the last lags of every particle
// sit on one or two lists) -- entries with few origins 32 at a time, one lane each, long ones one at a time with the lanes
// striding the origins -- and the exact value replaces the row entry.  Phase B: warp w walks the list warp w of K5 wrote,
// in order, and adds the exact values to the entries of the CTA's particle-sum row that only it touches (K5 left the
// marked lags out of that sum): fixed order, no atomics, the same bits in every run.
constexpr int K6_SHORT = 64;    // origins up to which an entry is a lane's job

__global__ void __launch_bounds__(K5_THREADS)
k6_helfand_refine(const HelfandFftArgs a) {
    const int T = a.T, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double cscale = 1.0 / ((double)a.D * a.denom);
    auto exact = [&](int n, int k, int i0, int istep) {            // sum over the origins i0, i0 + istep, ... of all dimensions
        const double* ser = a.series + (size_t)n * a.DS * a.Tld;
        double acc = 0.0;
        for (int d = 0; d < a.D; ++d) {
            const double* sd = ser + (size_t)d * a.Tld;
            for (int i = i0; i < T - k; i += istep) { const double df = sd[i] - sd[i + k]; acc = fma(df, df, acc); }
        }
        return acc;
    };
    // ---- phase A
    for (int lw = 0; lw < K5_WARPS; ++lw) {
        const unsigned cnt = min(a.count[(size_t)blockIdx.x * K5_WARPS + lw], a.cap);
        const unsigned long long* list = a.list + ((size_t)blockIdx.x * K5_WARPS + lw) * a.cap;
        for (unsigned e = threadIdx.x; e < cnt; e += K5_THREADS) {                       // short entries: a lane each
            const unsigned long long ent = list[e];
            const int n = (int)(ent >> 32), k = (int)(ent & 0xffffffffu);
            if (T - k > K6_SHORT) continue;
            a.by_particle[(size_t)n * a.Tld + k] = exact(n, k, 0, 1) / (double)(T - k) * cscale;
        }
        for (unsigned e = warp; e < cnt; e += K5_WARPS) {                                 // long entries: a warp each
            const unsigned long long ent = list[e];
            const int n = (int)(ent >> 32), k = (int)(ent & 0xffffffffu);
            if (T - k <= K6_SHORT) continue;                                              // uniform over the warp
            double acc = exact(n, k, lane, 32);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) a.by_particle[(size_t)n * a.Tld + k] = acc / (double)(T - k) * cscale;
        }
    }
    __syncthreads();            // the exact values written by the other warps of this CTA are visible
    // ---- phase B
    const unsigned n_ent = min(a.count[(size_t)blockIdx.x * K5_WARPS + warp], a.cap);
    const unsigned long long* my_list = a.list + ((size_t)blockIdx.x * K5_WARPS + warp) * a.cap;
    double* partial = a.partial + (size_t)blockIdx.x * a.Tld;
    for (unsigned e0 = 0; e0 < n_ent; e0 += 32) {
        const unsigned e = e0 + lane;
        const bool mine = e < n_ent;
        const unsigned long long ent = mine ? my_list[e] : 0ull;
        const int n = (int)(ent >> 32), k = (int)(ent & 0xffffffffu);
        const double v = mine ? a.by_particle[(size_t)n * a.Tld + k] : 0.0;
        // the values of one lag, summed in lane (= particle) order by the first lane that holds it, then one update of the sum row
        const unsigned same = __match_any_sync(0xffffffffu, mine ? k : -1 - lane);
        double sum = 0.0;
        for (int src = 0; src < 32; ++src) {                        // uniform trip count: every lane takes part in every shuffle
            const double dv = __shfl_sync(0xffffffffu, v, src);
            if ((same >> src) & 1u) sum += dv;
        }
        if (mine && lane == __ffs(same) - 1) partial[k] += sum;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Probe: the FP64 FMA rate of the whole device (bench.py's measured FP64 roofline denominator; MEASURED_PEAKS.json
// holds HBM and bf16 figures only).  16 independent DFMA chains per thread, 8 warps per SM sub-partition.
// ---------------------------------------------------------------------------
constexpr int KP_CHAINS = 16, KP_UNROLL = 8, KP_THREADS = 256;
__global__ void __launch_bounds__(KP_THREADS)
k_probe_fp64(double* out, int iters, double a, double b) {
    double x[KP_CHAINS];
#pragma unroll
    for (int i = 0; i < KP_CHAINS; ++i) x[i] = a + (double)(i + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < KP_UNROLL; ++rep) {
#pragma unroll
            for (int i = 0; i < KP_CHAINS; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < KP_CHAINS; ++i) sum += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// ---------------------------------------------------------------------------
// K4a: fixed-order sum of the per-CTA partial rows -> atom sum of this shard.
// ---------------------------------------------------------------------------
__global__ void k_sum_partials(const double* __restrict__ partial, int nrows, long long Tld, int T,
                               double* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T) return;
    double s = 0.0;
    for (int r = 0; r < nrows; ++r) s += partial[(size_t)r * Tld + k];
    out[k] = s;
}

// K4c: the atom mean itself, on the device: mean[k] = sum[k] / count (count = the all-reduced particle number in slot Tld)
__global__ void k_atom_mean(const double* __restrict__ ts_sum, long long Tld, int T, double* __restrict__ mean) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < T) mean[k] = ts_sum[k] / ts_sum[Tld];
}

// ---------------------------------------------------------------------------
// K7: Green-Kubo post-processing of the device-resident atom-mean timeseries: trapezoid integral and
// running integral of y[w] over x[w], w = start, start + step, ... (n points) -- what self_diffusivity_gk
// (velocityautocorr.py:316-322) and plot_running_integral (:407-414) compute with scipy.integrate -- and the
// least-squares slope of y over x (the linear fit of viscosity.py:235-245).  One CTA; O(T).
//   out[0] = integral, out[1] = slope; running[0] = initial, running[i] = sum of the first i trapezoids (may be null)
// ---------------------------------------------------------------------------
constexpr int K7_THREADS = 1024;
__global__ void __launch_bounds__(K7_THREADS)
k7_green_kubo(const double* __restrict__ y, const double* __restrict__ x, long long start, long long step, int n,
              double initial, double* __restrict__ running, double* __restrict__ out) {
    __shared__ double wtot[K7_THREADS / 32];
    __shared__ double carry_s;
    __shared__ double red[2][K7_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0.0;
    if (running && tid == 0 && n > 0) running[0] = initial;
    __syncthreads();
    // trapezoids i = 0 .. n-2 in tiles of K7_THREADS, block scan per tile
    for (int i0 = 0; i0 < n - 1; i0 += K7_THREADS) {
        const int i = i0 + tid;
        double t = 0.0;
        if (i < n - 1) {
            const long long a = start + (long long)i * step, b = a + step;
            t = 0.5 * (x[b] - x[a]) * (y[b] + y[a]);
        }
        double v = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) wtot[warp] = v;
        __syncthreads();
        double off = carry_s;
        for (int w = 0; w < warp; ++w) off += wtot[w];
        v += off;
        if (running && i < n - 1) running[i + 1] = v;
        __syncthreads();
        if (tid == K7_THREADS - 1) carry_s = v;
        __syncthreads();
    }
    // least squares of y over x: two passes (means, then centred sums), every sum in a fixed order
    auto block_sum2 = [&](double a, double b, double* ra, double* rb) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        __syncthreads();
        if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
        __syncthreads();
        double sa = 0.0, sb = 0.0;
        for (int w = 0; w < K7_THREADS / 32; ++w) { sa += red[0][w]; sb += red[1][w]; }
        *ra = sa; *rb = sb;
    };
    double sx = 0.0, sy = 0.0;
    for (int i = tid; i < n; i += K7_THREADS) { const long long a = start + (long long)i * step; sx += x[a]; sy += y[a]; }
    double mx, my;
    block_sum2(sx, sy, &mx, &my);
    mx /= (double)(n > 0 ? n : 1); my /= (double)(n > 0 ? n : 1);
    double sxx = 0.0, sxy = 0.0;
    for (int i = tid; i < n; i += K7_THREADS) {
        const long long a = start + (long long)i * step;
        const double dx = x[a] - mx;
        sxx += dx * dx; sxy += dx * (y[a] - my);
    }
    double txx, txy;
    block_sum2(sxx, sxy, &txx, &txy);
    if (tid == 0) {
        out[0] = carry_s;
        out[1] = (n >= 2 && txx != 0.0) ? txy / txx : 0.0;
    }
}

// K4b: lag-major view of the per-particle result for a range of atoms:
// out[k][j] = by_particle[atom0 + j][k]   (reference layout, velocityautocorr.py:145-147)
__global__ void k_to_lag_major(const double* __restrict__ by_particle, long long Tld, int T,
                               long long atom0, int natoms, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        int j = j0 + jj, k = k0 + threadIdx.x;
        if (j < natoms && k < T) tile[jj][threadIdx.x] = by_particle[(size_t)(atom0 + j) * Tld + k];
    }
    __syncthreads();
    for (int kk = threadIdx.y; kk < 32; kk += blockDim.y) {
        int k = k0 + kk, j = j0 + threadIdx.x;
        if (j < natoms && k < T) out[(size_t)k * natoms + j] = tile[threadIdx.x][kk];
    }
}

}  // namespace ta

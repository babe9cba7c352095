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
         int D, int d0, int d1, int d2) {
    // frames on grid.x (2^31 - 1 tiles: any T the library accepts), particles on grid.y, tiled further by the loop below
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
        const int nrows = na * D;
        for (int r = threadIdx.y; r < nrows; r += blockDim.y) {
            const int a = r / D, d = r - a * D;
            OUT* dst = series + ((size_t)(a0 + a) * D + d) * Tld + frame0 + f0;
            const int col = a * 3 + dims[d];
            for (int f = threadIdx.x; f < nf; f += blockDim.x) dst[f] = tile[f][col];
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
    const R* series;             // [natoms][D][Tld]  (stored in the arithmetic type)
    double* by_particle;         // [natoms][Tld]
    double* partial;             // [gridDim.x][Tld]
    int natoms, D;
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
        const R* ser = args.series + (size_t)a * args.D * args.Tld;
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
    static TA_HD void mbar_wait(unsigned long long* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
        const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(a), "r"(parity) : "memory");
        }
#endif
    }
};

// One kernel per (R1, arithmetic type).  NT = 16 R1 threads; the bulk series prefetch where k1f_prefetch says so.
template <int R1, typename RT>
__global__ void __launch_bounds__(k1f_threads(R1), k1f_min_blocks(k1f_threads(R1), (int)sizeof(RT)))
k1f_fft_acf(const K1FArgs<RT> args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    k1f_body<R1, k1f_threads(R1), DevCtx, RT, k1f_prefetch(R1, (int)sizeof(RT))>(args, smem_raw, (int)threadIdx.x, (int)blockIdx.x,
                                                                                 (int)gridDim.x);
}

// The FP64 kernel of ten warps (R1 = 20) with the register cap stated outright instead of derived from launch bounds.
// With three warps on two of the sub-partitions (16,384 registers each) 170 registers per thread is the ceiling either
// way, but ptxas schedules differently under the two attributes: __launch_bounds__(320, 1) gives 168 registers and
// 16-32 B of spills, __maxnreg__(184) 166 registers and none -- 24.2 -> 23.5 ms at 100k x 10k (caps 152 / 160 / 168 /
// 176 / 184: 24.21 / 24.08 / 23.66 / 23.52 / 23.46 ms; 200 does not fit).
constexpr int K1F_MAXREG = 184;
template <int R1, typename RT>
__global__ void __maxnreg__(K1F_MAXREG)
k1f_fft_acf_mr(const K1FArgs<RT> args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    k1f_body<R1, k1f_threads(R1), DevCtx, RT, k1f_prefetch(R1, (int)sizeof(RT))>(args, smem_raw, (int)threadIdx.x, (int)blockIdx.x,
                                                                                 (int)gridDim.x);
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
//   S1[k] = sum_{i<T-k} g[i]^2 + sum_{i>=k} g[i]^2 (from one prefix sum of q[i] = sum_d g_d[i]^2)
// One CTA per particle at a time; prefix sums in shared memory.  The difference cancels, so the
// relative error grows like eps * S1 / MSD (largest at small lags): this route is opt-in.
// ---------------------------------------------------------------------------
struct HelfandFftArgs {
    const double* series;   // [natoms][D][Tld]
    double* by_particle;    // [natoms][Tld]  in: sum_d acf_d ; out: viscosity function
    double* partial;        // [gridDim.x][Tld]   (K6)
    int natoms, D, T;
    long long Tld;
    double denom;           // 2 kB <V> temp_avg
    uint32_t* flags;        // [natoms][nwords]  bit k of a particle: lag k needs the exact evaluation
    int nwords;             // ceil(T / 32)
    unsigned long long* nflagged;   // total number of flagged (particle, lag) pairs
    double thr;             // a lag is flagged when its un-normalised MSD is below thr * sum_i sum_d g^2
};

constexpr int K5_THREADS = 1024;

// K5: S1[k] - 2 S2[k] per particle.  Both terms are of the size of sum g^2 and carry rounding errors of that size
// (C eps sum g^2: the FFT autocorrelation and the prefix sums), so the difference is only as accurate as
// C eps sum g^2 / MSD[k] relative.  Lags whose MSD is too small for the 1e-10 bar (short lags of smooth series, the last
// lags of any series) are marked in a per-particle bitmap and evaluated exactly by K6.
__global__ void __launch_bounds__(K5_THREADS)
k5_helfand_fft_finish(const HelfandFftArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* P = reinterpret_cast<double*>(smem_raw);   // P[j] = sum_{i<j} q[i], j = 0..T
    __shared__ double wsum[K5_THREADS / 32];
    __shared__ unsigned cta_flagged;
    const int T = a.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = (T + K5_THREADS - 1) / K5_THREADS;
    const int lo = min(T, tid * seg), hi = min(T, lo + seg);
    if (tid == 0) cta_flagged = 0;
    double* partial = a.partial + (size_t)blockIdx.x * a.Tld;   // particle sum of the rows as K5 leaves them; K6 corrects it
    for (int n = blockIdx.x; n < a.natoms; n += gridDim.x) {
        const double* ser = a.series + (size_t)n * a.D * a.Tld;
        __syncthreads();   // previous particle's P fully consumed
        // q[i] into P[i + 1] (coalesced), then a three-level inclusive scan
        for (int i0 = 0; i0 < T; i0 += 4 * K5_THREADS) {       // four samples per thread in flight
            double q[4] = {0.0, 0.0, 0.0, 0.0};
            for (int d = 0; d < a.D; ++d) {
                double gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { const int i = i0 + j * K5_THREADS + tid; gv[j] = i < T ? ser[(size_t)d * a.Tld + i] : 0.0; }
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = fma(gv[j], gv[j], q[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int i = i0 + j * K5_THREADS + tid; if (i < T) P[i + 1] = q[j]; }
        }
        if (tid == 0) P[0] = 0.0;
        __syncthreads();
        double run = 0.0;
        for (int i = lo; i < hi; ++i) { run += P[i + 1]; P[i + 1] = run; }
        double incl = run;   // inclusive scan of the segment totals over the block
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        double base = incl - run;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        for (int i = lo; i < hi; ++i) P[i + 1] += base;
        __syncthreads();
        double* row = a.by_particle + (size_t)n * a.Tld;
        uint32_t* fl = a.flags + (size_t)n * a.nwords;
        const double tot = P[T];
        const double floor_msd = a.thr * tot;
        const double cscale = 1.0 / ((double)a.D * a.denom);
        unsigned mine = 0;
        // lags k = k0 + tid, in batches of KB: the global loads of a batch (row, partial) are issued together;
        // uniform trip count (the ballot needs whole warps)
        constexpr int KB = 4;
        for (int k0 = 0; k0 < T; k0 += KB * K5_THREADS) {
            double r[KB], pa[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int k = k0 + j * K5_THREADS + tid;
                r[j] = k < T ? row[k] : 0.0;
                pa[j] = k < T ? partial[k] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int k = k0 + j * K5_THREADS + tid;
                bool flag = false;
                if (k < T) {
                    double val = 0.0;                           // lag 0 stays exactly 0 (viscosity.py:207-210)
                    if (k > 0) {
                        const double s1 = P[T - k] + (tot - P[k]);
                        const double nk = (double)(T - k);
                        const double msd = s1 - 2.0 * r[j] * nk;    // un-normalised, what the threshold is about
                        flag = !(msd >= floor_msd);                 // also catches a NaN
                        val = (s1 / nk - 2.0 * r[j]) * cscale;
                    }
                    row[k] = val;
                    partial[k] = pa[j] + val;
                }
                const unsigned m = __ballot_sync(0xffffffffu, flag);
                if (lane == 0 && k < T) { fl[k >> 5] = m; mine += __popc(m); }
            }
        }
        if (lane == 0 && mine) atomicAdd(&cta_flagged, mine);
    }
    __syncthreads();
    if (tid == 0 && cta_flagged) atomicAdd(a.nflagged, (unsigned long long)cta_flagged);
}

// K6: exact evaluation of the flagged lags, sum_d sum_i (g_d[i] - g_d[i+k])^2 (viscosity.py:212-226; one warp per lag, lanes
// stride the origins, fixed-order reduction); the row takes the exact value and the per-CTA partial row (the particle sum K5
// formed, same CTA -> particle map) the difference.  A particle whose flagged lags add up to little work (the usual case: the
// last one or two lags, a handful of origins) reads its few samples straight from global memory; otherwise its series are
// staged in shared memory one dimension at a time.
__global__ void __launch_bounds__(K5_THREADS)
k6_helfand_refine(const HelfandFftArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* g = reinterpret_cast<double*>(smem_raw);   // one series of the particle (staged particles only)
    __shared__ unsigned long long work_sum;
    const int T = a.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = K5_THREADS / 32;
    double* partial = a.partial + (size_t)blockIdx.x * a.Tld;
    for (int n = blockIdx.x; n < a.natoms; n += gridDim.x) {
        const double* ser = a.series + (size_t)n * a.D * a.Tld;
        double* row = a.by_particle + (size_t)n * a.Tld;
        const uint32_t* fl = a.flags + (size_t)n * a.nwords;
        __syncthreads();
        if (tid == 0) work_sum = 0ull;
        __syncthreads();
        unsigned long long w_mine = 0ull;               // origins to visit: sum over flagged lags of (T - k)
        for (int w = tid; w < a.nwords; w += K5_THREADS) {
            uint32_t m = fl[w];
            while (m) { const int k = (w << 5) + __ffs(m) - 1; m &= m - 1; w_mine += (unsigned long long)(T - k); }
        }
        if (w_mine) atomicAdd(&work_sum, w_mine);
        __syncthreads();
        const unsigned long long work = work_sum;
        if (work == 0ull) continue;
        const bool staged = work > 4ull * (unsigned long long)T;
        for (int d = 0; d < a.D; ++d) {
            const double* sd = ser + (size_t)d * a.Tld;
            if (staged) {
                __syncthreads();
                for (int i = tid; i < T; i += K5_THREADS) g[i] = sd[i];
                __syncthreads();
            }
            const double* src = staged ? g : sd;
            for (int w = warp; w < a.nwords; w += NW) {  // flagged lags of this particle, dealt to the warps word by word
                uint32_t m = fl[w];
                while (m) {
                    const int k = (w << 5) + __ffs(m) - 1;
                    m &= m - 1;
                    double acc = 0.0;
                    for (int i = lane; i < T - k; i += 32) { const double df = src[i] - src[i + k]; acc = fma(df, df, acc); }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0) {
                        // the running exact sum over the dimensions waits in the row itself; what K5 had put there leaves
                        // the particle sum first and the exact value enters it after the last dimension
                        if (d == 0) { partial[k] -= row[k]; row[k] = acc; }
                        else row[k] += acc;
                        if (d == a.D - 1) {
                            const double e = row[k] / (double)(T - k) / (double)a.D / a.denom;
                            row[k] = e;
                            partial[k] += e;
                        }
                    }
                }
            }
        }
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

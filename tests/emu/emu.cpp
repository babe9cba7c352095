// CPU thread-emulation harness for the per-CTA phase functions of the CUDA
// kernels (TEST INFRASTRUCTURE ONLY -- compiled and loaded by tests/, never by
// the product).  It includes the very same headers the kernels are built from
// and runs each phase for tid = 0..nthreads-1 where the kernel has a
// __syncthreads(), so index arithmetic, scramble tables and twiddles can be
// checked against the oracle on a machine without a GPU.
#include <vector>
#include <cstring>
#include "../../transport_analysis_b200/csrc/ta_common.cuh"
#include "../../transport_analysis_b200/csrc/fft_plan.h"
#include "../../transport_analysis_b200/csrc/fft_core.cuh"
#include "../../transport_analysis_b200/csrc/windowed_core.cuh"
#include "../../transport_analysis_b200/csrc/k1_fast.cuh"
#include "fiber_cta.h"

using namespace ta;

// the series as the kernels find them in HBM: stored in the arithmetic type R
template <typename R>
static std::vector<R> series_as(const double* series, size_t n) { return std::vector<R>(series, series + n); }

template <typename R>
static int run_fft(const double* series_f64, int T, int D, int Tld, int nthr, double* row, double* partial) {
    const std::vector<R> series_r = series_as<R>(series_f64, (size_t)D * Tld);
    const R* series = series_r.data();
    FftPlanHost hp;
    int rc = ta_build_fft_plan(T, &hp);
    if (rc) return rc;
    int nlo = 1 << hp.lo_bits, nhi = (int)hp.tw_hi.size() / 2;
    std::vector<cplx<R>> lo(nlo), hi(nhi);
    for (int i = 0; i < nlo; ++i) lo[i] = cmake<R>((R)hp.tw_lo[2 * i], (R)hp.tw_lo[2 * i + 1]);
    for (int i = 0; i < nhi; ++i) hi[i] = cmake<R>((R)hp.tw_hi[2 * i], (R)hp.tw_hi[2 * i + 1]);
    std::vector<uint32_t> own0;
    for (int p = 0; p < hp.H; ++p) if ((uint32_t)p <= hp.pair0[p]) own0.push_back(p);
    FftTables<R> t;
    t.T = T; t.H = hp.H; t.L = hp.L; t.npasses = hp.npasses;
    for (int i = 0; i < hp.npasses; ++i) t.radix[i] = hp.radix[i];
    t.lo_bits = hp.lo_bits; t.tw_lo = lo.data(); t.tw_hi = hi.data();
    t.ftab = hp.ftab.data(); t.pair0 = hp.pair0.data(); t.own0 = own0.data(); t.npairs0 = (int)own0.size();
    std::vector<cplx<R>> buf(hp.H);
    std::vector<R> sd(hp.H + 1);
#define PHASE(call) for (int tid = 0; tid < nthr; ++tid) { call; }
    for (int r = 0; r < 2; ++r) {
        PHASE(fft_zero_acc<R>(tid, nthr, sd.data(), t));
        for (int d = 0; d < D; ++d) {
            PHASE(fft_load<R>(tid, nthr, buf.data(), series + (size_t)d * Tld, t, r));
            int size = t.H;
            for (int ps = 0; ps < t.npasses; ++ps) {
                int s = size / t.radix[ps];
                PHASE(dif_pass_any<R>(t.radix[ps], tid, nthr, buf.data(), t, s));
                size = s;
            }
            PHASE(fft_accumulate<R>(tid, nthr, buf.data(), sd.data(), t, r));
        }
        PHASE(fft_build<R>(tid, nthr, buf.data(), sd.data(), t, r));
        int s = 1;
        for (int ps = t.npasses - 1; ps >= 0; --ps) {
            PHASE(dit_pass_any<R>(t.radix[ps], tid, nthr, buf.data(), t, s));
            s *= t.radix[ps];
        }
        PHASE(fft_store<R>(tid, nthr, buf.data(), row, partial, t, r));
    }
    return 0;
}

template <typename R>
static int run_win(const double* series_f64, int T, int D, int Tld, int mode, int nwarps, int nsplit, double* res) {
    const std::vector<R> series_r = series_as<R>(series_f64, (size_t)D * Tld);
    const R* series = series_r.data();
    // the kernel body itself (windowed_core.cuh win_body) for one particle, CTAs of nwarps warps, one per part of the
    // particle (nsplit); returns the un-normalised lag sums (row * (T - k) [* D * denom])
    const int ne = win_smem_elems(T);
    std::vector<unsigned char> smem((size_t)((ne + 1) & ~1) * sizeof(R) + (size_t)T * sizeof(double) + 64);
    unsigned char* sm = smem.data() + (16 - ((uintptr_t)smem.data() & 15)) % 16;
    const int nblk = nsplit;
    std::vector<double> row(Tld, 0.0), partial((size_t)nblk * Tld, 0.0);
    WinArgs a;
    a.series = series; a.by_particle = row.data(); a.partial = partial.data();
    a.natoms = 1; a.D = D; a.DS = D; a.T = T; a.Tld = Tld; a.denom = 1.0;
    a.scratch = nullptr; a.scratch_stride = 0; a.nsplit = nsplit;
    const int nthr = 32 * nwarps;
    for (int bid = 0; bid < nblk; ++bid) {
        if (mode == TA_WIN_PRODUCT) emu::run_cta(nthr, [&](int tid) { win_body<R, TA_WIN_PRODUCT, emu::EmuCtx>(a, sm, tid, nthr, bid, nblk); });
        else emu::run_cta(nthr, [&](int tid) { win_body<R, TA_WIN_SQDIFF, emu::EmuCtx>(a, sm, tid, nthr, bid, nblk); });
    }
    for (int k = 0; k < T; ++k) {
        double psum = 0.0;                              // every lag is finished by exactly one CTA
        int writers = 0;
        for (int b = 0; b < nblk; ++b) { psum += partial[(size_t)b * Tld + k]; writers += partial[(size_t)b * Tld + k] != 0.0; }
        if (row[k] != psum || writers > 1) return -2;
        res[k] = row[k] * (double)(T - k) * (mode == TA_WIN_PRODUCT ? 1.0 : (double)D);
    }
    return 0;
}

template <int R1, typename RT, bool TMEM = false>
static int run_k1fast(const double* series_f64, int T, int D, int Tld, int natoms, int nblk, double* by_particle,
                      double* partial) {
    K1FastPlan p;
    int rc = k1f_build_plan(T, Tld, R1, &p);
    if (rc) return rc;
    const std::vector<RT> series = series_as<RT>(series_f64, (size_t)natoms * D * Tld);
    const std::vector<RT> omega(p.omega.begin(), p.omega.end()), tw2(p.tw2.begin(), p.tw2.end()),
        wbase(p.wbase.begin(), p.wbase.end()), inv(p.inv.begin(), p.inv.end());
    K1FArgs<RT> a;
    a.series = series.data(); a.by_particle = by_particle; a.partial = partial;
    a.omega = reinterpret_cast<const cplx<RT>*>(omega.data());
    a.tw2 = reinterpret_cast<const cplx<RT>*>(tw2.data());
    a.map = p.map.data();
    a.wbase = reinterpret_cast<const cplx<RT>*>(wbase.data());
    a.inv = inv.data();
    a.natoms = natoms; a.D = D; a.DS = D; a.T = T; a.nh = p.nh; a.Tld = Tld;
    constexpr bool PREF = k1f_prefetch(R1, (int)sizeof(RT));        // the build the library ships for this (R1, RT)
    std::vector<unsigned char> smem(k1f_smem_bytes(R1, PREF, (int)sizeof(RT)) + 64);
    unsigned char* sm = smem.data() + (16 - ((uintptr_t)smem.data() & 15)) % 16;
    for (int bid = 0; bid < nblk; ++bid)
        emu::run_cta(k1f_threads(R1), [&](int tid) { k1f_body<R1, k1f_threads(R1), emu::EmuCtx, RT, PREF, true, TMEM>(a, sm, tid, bid, nblk); });
    return 0;
}

template <typename RT>
static int run_k1fast_any(const double* series, int T, int D, int Tld, int natoms, int nblk, int R1, double* by_particle, double* partial) {
    switch (R1) {
        case 4: return run_k1fast<4, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 6: return run_k1fast<6, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 8: return run_k1fast<8, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 10: return run_k1fast<10, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 12: return run_k1fast<12, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 16: return run_k1fast<16, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 20: return run_k1fast<20, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 24: return run_k1fast<24, RT>(series, T, D, Tld, natoms, nblk, by_particle, partial);
    }
    return -1;
}

extern "C" {
int emu_k1fast_r1(int T) { return k1f_choose_r1(T); }
// the three-pass kernel body as shipped for (R1, precision): use_f32 = 0 FP64, 1 FP32 (float series, float arithmetic),
// 2 FP64 with the tensor-memory output stage (R1 = 16, 20)
int emu_k1fast(const double* series, int T, int D, int Tld, int natoms, int nblk, int R1, int use_f32, double* by_particle,
               double* partial) {
    if (use_f32 == 2) {     // FP64 with the output stage's per-thread streams in tensor memory (the R1 the library ships it for)
        if (R1 == 16) return run_k1fast<16, double, true>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        if (R1 == 20) return run_k1fast<20, double, true>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        return -1;
    }
    return use_f32 ? run_k1fast_any<float>(series, T, D, Tld, natoms, nblk, R1, by_particle, partial)
                   : run_k1fast_any<double>(series, T, D, Tld, natoms, nblk, R1, by_particle, partial);
}
int emu_fft_acf(const double* series, int T, int D, int Tld, int nthr, int use_f32, double* row, double* partial) {
    return use_f32 ? run_fft<float>(series, T, D, Tld, nthr, row, partial)
                   : run_fft<double>(series, T, D, Tld, nthr, row, partial);
}
int emu_windowed(const double* series, int T, int D, int Tld, int mode, int nwarps, int nsplit, int use_f32, double* res) {
    return use_f32 ? run_win<float>(series, T, D, Tld, mode, nwarps, nsplit, res)
                   : run_win<double>(series, T, D, Tld, mode, nwarps, nsplit, res);
}
int emu_plan(int T, int* H, int* npasses, int* radix) {
    FftPlanHost hp;
    int rc = ta_build_fft_plan(T, &hp);
    if (rc) return rc;
    *H = hp.H; *npasses = hp.npasses;
    for (int i = 0; i < hp.npasses; ++i) radix[i] = hp.radix[i];
    return 0;
}
}

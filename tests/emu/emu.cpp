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
#include "../../transport_analysis_b200/csrc/k1_r8.cuh"
#include "fiber_cta.h"

using namespace ta;

template <typename R>
static int run_fft(const double* series, int T, int D, int Tld, int nthr, double* row, double* partial) {
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
static int run_win(const double* series, int T, int D, int Tld, int mode, int nwarps, double* res) {
    // the kernel body itself (windowed_core.cuh win_body) for one particle, one CTA of nwarps warps;
    // returns the un-normalised lag sums (row * (T - k) [* D * denom])
    const int ne = win_smem_elems(T);
    std::vector<unsigned char> smem((size_t)((ne + 1) & ~1) * sizeof(R) + (size_t)T * sizeof(double) + 64);
    unsigned char* sm = smem.data() + (16 - ((uintptr_t)smem.data() & 15)) % 16;
    std::vector<double> row(Tld, 0.0), partial(Tld, 0.0);
    WinArgs a;
    a.series = series; a.by_particle = row.data(); a.partial = partial.data();
    a.natoms = 1; a.D = D; a.T = T; a.Tld = Tld; a.denom = 1.0;
    a.scratch = nullptr; a.scratch_stride = 0;
    const int nthr = 32 * nwarps;
    if (mode == TA_WIN_PRODUCT) emu::run_cta(nthr, [&](int tid) { win_body<R, TA_WIN_PRODUCT, emu::EmuCtx>(a, sm, tid, nthr, 0, 1); });
    else emu::run_cta(nthr, [&](int tid) { win_body<R, TA_WIN_SQDIFF, emu::EmuCtx>(a, sm, tid, nthr, 0, 1); });
    for (int k = 0; k < T; ++k) {
        if (row[k] != partial[k]) return -2;
        res[k] = row[k] * (double)(T - k) * (mode == TA_WIN_PRODUCT ? 1.0 : (double)D);
    }
    return 0;
}

template <int R1, int VAR = 0>
static int run_k1fast(const double* series, int T, int D, int Tld, int natoms, int nblk, double* by_particle,
                      double* partial) {
    K1FastPlan p;
    int rc = k1f_build_plan(T, Tld, R1, &p);
    if (rc) return rc;
    K1FArgs a;
    a.series = series; a.by_particle = by_particle; a.partial = partial;
    a.omega = reinterpret_cast<const cd*>(p.omega.data());
    a.tw2 = reinterpret_cast<const cd*>(p.tw2.data());
    a.tw8 = reinterpret_cast<const cd*>(p.tw8.data());
    a.map = p.map.data();
    a.wbase = reinterpret_cast<const cd*>(p.wbase.data());
    a.inv = p.inv.data();
    a.natoms = natoms; a.D = D; a.T = T; a.nh = p.nh; a.Tld = Tld; a.prefetch = 0; a.stagger = 0; a.prof = nullptr;
    std::vector<unsigned char> smem(k1f_smem_bytes(R1, VAR) + 64);
    unsigned char* sm = smem.data() + (16 - ((uintptr_t)smem.data() & 15)) % 16;
    for (int bid = 0; bid < nblk; ++bid)
        emu::run_cta(k1f_threads(R1), [&](int tid) { k1f_body<R1, k1f_threads(R1), emu::EmuCtx, false, VAR>(a, sm, tid, bid, nblk); });
    return 0;
}

template <int R>
static int run_k1r8(const double* series, int T, int D, int Tld, int natoms, int nblk, double* by_particle, double* partial) {
    K1R8Plan p;
    int rc = k1e_build_plan(T, R, &p);
    if (rc) return rc;
    K1EArgs a;
    a.series = series; a.by_particle = by_particle; a.partial = partial;
    a.omega = reinterpret_cast<const cd*>(p.omega.data());
    a.tw2 = reinterpret_cast<const cd*>(p.tw2.data());
    a.tw3 = reinterpret_cast<const cd*>(p.tw3.data());
    a.map = p.map.data();
    a.wbase = reinterpret_cast<const cd*>(p.wbase.data());
    a.natoms = natoms; a.D = D; a.T = T; a.nh = p.nh; a.Tld = Tld; a.sm_slots = nullptr; a.stagger = 0;
    std::vector<unsigned char> smem(k1e_smem_bytes(R) + 64);
    unsigned char* sm = smem.data() + (16 - ((uintptr_t)smem.data() & 15)) % 16;
    for (int bid = 0; bid < nblk; ++bid)
        emu::run_cta(k1e_threads(R), [&](int tid) { k1e_body<R, emu::EmuCtx>(a, sm, tid, bid, nblk); });
    return 0;
}

extern "C" {
int emu_k1r8_r(int T) { return k1e_choose_r(T); }
int emu_k1r8(const double* series, int T, int D, int Tld, int natoms, int nblk, int R, double* by_particle, double* partial) {
    switch (R) {
        case 4: return run_k1r8<4>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 5: return run_k1r8<5>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 6: return run_k1r8<6>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 8: return run_k1r8<8>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 10: return run_k1r8<10>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 12: return run_k1r8<12>(series, T, D, Tld, natoms, nblk, by_particle, partial);
    }
    return -1;
}
// the plan's thread maps, for structural checks: returns NT, fills map[2 * NT]
int emu_k1r8_map(int T, int R, unsigned* map) {
    K1R8Plan p;
    int rc = k1e_build_plan(T, R, &p);
    if (rc) return rc;
    for (size_t i = 0; i < p.map.size(); ++i) map[i] = p.map[i];
    return p.NT;
}
int emu_k1fast_r1(int T) { return k1f_choose_r1(T); }
// the experiment variants of the three-pass kernel (k1_fast.cuh VAR bits) at R1 = 20 and R1 = 10
int emu_k1fast_var(const double* series, int T, int D, int Tld, int natoms, int nblk, int R1, int var, double* by_particle,
                   double* partial) {
    if (R1 == 20) switch (var) {
        case 1: return run_k1fast<20, 1>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 2: return run_k1fast<20, 2>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 4: return run_k1fast<20, 4>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 6: return run_k1fast<20, 6>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 8: return run_k1fast<20, 8>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 12: return run_k1fast<20, 12>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 52: return run_k1fast<20, 52>(series, T, D, Tld, natoms, nblk, by_particle, partial);
    }
    if (R1 == 10) switch (var) {
        case 12: return run_k1fast<10, 12>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 2: return run_k1fast<10, 2>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 4: return run_k1fast<10, 4>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 6: return run_k1fast<10, 6>(series, T, D, Tld, natoms, nblk, by_particle, partial);
    }
    return -1;
}
int emu_k1fast(const double* series, int T, int D, int Tld, int natoms, int nblk, int R1, double* by_particle,
               double* partial) {
    switch (R1) {
        case 4: return run_k1fast<4>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 6: return run_k1fast<6>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 8: return run_k1fast<8>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 10: return run_k1fast<10>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 12: return run_k1fast<12>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 16: return run_k1fast<16>(series, T, D, Tld, natoms, nblk, by_particle, partial);
        case 20: return run_k1fast<20>(series, T, D, Tld, natoms, nblk, by_particle, partial);
    }
    return -1;
}
int emu_fft_acf(const double* series, int T, int D, int Tld, int nthr, int use_f32, double* row, double* partial) {
    return use_f32 ? run_fft<float>(series, T, D, Tld, nthr, row, partial)
                   : run_fft<double>(series, T, D, Tld, nthr, row, partial);
}
int emu_windowed(const double* series, int T, int D, int Tld, int mode, int nwarps, int use_f32, double* res) {
    return use_f32 ? run_win<float>(series, T, D, Tld, mode, nwarps, res)
                   : run_win<double>(series, T, D, Tld, mode, nwarps, res);
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

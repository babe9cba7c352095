// K1 fast path: FFT autocorrelation with H = 256 * R1 in three passes.
//
// Same mathematics as the general kernel (fft_plan.h / fft_core.cuh; replaces
// tidynamics.acf as called from transport_analysis/velocityautocorr.py:210-214):
// the real series is packed z[n] = x[2n] + i x[2n+1], the two residue chains
// r = 0, 1 of the zero-padded length-2H spectrum are H-point complex FFTs, the
// power spectrum of the real series is formed from conjugate bin pairs and
// summed over the D series before ONE inverse per residue.
//
// What is different is the data movement (the general kernel makes ~12 shared-
// memory transfers of the H-point buffer per FFT and was bound by them):
//   * H = R1 * 16 * 16, decimation in frequency, three passes per FFT:
//       P1  radix R1, stride 256   fused with the global load and the residue twist
//       P2  radix 16, stride 16    shared -> registers -> shared
//       P3  radix 16, stride 1     shared -> registers, fused with the pair accumulation
//     and the mirror image P3' P2' P1' for the inverse, P3' fed from registers,
//     P1' fused with the final twist / normalisation / store: 2 shared-memory
//     round trips per FFT.
//   * In P3 the thread -> butterfly map puts conjugate-partner butterflies on
//     lanes l and l ^ 16, so partner bins are exchanged with shuffles and the
//     (Sigma, Delta) accumulators of a thread's 8 bin pairs stay in registers
//     over the D series; the same exchange feeds the inverse.
//   * P2 runs on the same warp as the P3 butterflies that consume its output (a warp owns the two
//     256-point blocks k1 and R1-k1 -- r = 1: k1 and R1-1-k1 -- in both passes), so P2 -> P3 and
//     P3' -> P2' need a warp-level sync only; the CTA barriers that remain are P1 -> P2, the
//     buffer hand-over to the next series, and P2' -> P1'.  Warps drift apart inside the
//     barrier-free stretch, which lets one warp's shared-memory traffic overlap another's math.
//   * All butterflies are register DFTs with compile-time constants
//     (dft_regs.cuh); P2 twiddles come from a 240-entry table, P1 twiddles are
//     powers of one table entry per thread.
//   * Shared buffer layout: element e lives at e + (e >> 4) (one pad element
//     per 16), which makes all three access patterns bank-conflict free.
#pragma once
#include <cstdint>
#include <vector>
#include "ta_common.cuh"
#include "dft_regs.cuh"
#include "fft_plan.h"

namespace ta {

struct K1FArgs {
    const double* series;        // [natoms][D][Tld]
    double* by_particle;         // [natoms][Tld]
    double* partial;             // [grid][Tld]
    const cd* omega;             // [256]      w_{2H}^j
    const cd* tw2;               // [15][16]   w_256^{j k}, k = 1..15
    const cd* tw8;               // [8][16]    TwDit table of om = w_256^j (K1F_VAR_DEFTW)
    const uint32_t* map;         // [2][16 R1] per residue: bits 0-15 P3 butterfly of a thread, 16-23 its P2 block, 30/31 flags
    const cd* wbase;             // [2][16 R1] w_L^{2 G0 + r} of that butterfly
    const double* inv;           // [Tld]      1 / (L (T - k)), 0 beyond T
    int natoms, D, T, nh;
    long long Tld;
    long long* prof;             // optional [grid][12] per-phase clock totals of thread 0 (debug; may be null)
    int prefetch;                // 1: ask L2 for the next particle's series at the start of each particle
    int stagger;                 // clocks the second wave of CTAs (bid >= nblk / 2) idles before its first particle
};

constexpr uint32_t K1F_SELF0 = 1u << 30;   // butterfly holding g = 0 (pairs k <-> 16-k in-thread)
constexpr uint32_t K1F_SELF8 = 1u << 31;   // butterfly holding g = H/2 (pairs k <-> 15-k in-thread)

constexpr int k1f_smem_bytes(int R1, int VAR = 0) {
    return (256 * R1 + 16 * R1 + 256 + 240 + ((VAR & 2) ? 256 * R1 : 0) + ((VAR & 4) ? 256 * R1 + 1 : 0)) * (int)sizeof(cd);
}
// Threads per CTA (NT) and resident CTAs per SM the kernel is compiled for.  A pass has NV = 16 R1
// radix-16 butterflies; a thread owns the butterflies vt = tid, tid + NT, ... (whole warps, so the
// lane exchange of P3 stays inside a warp).  Default: one butterfly per thread (NT = NV).
// Measured on B200 at R1 = 20, 100k x 10k (profiles/r01_k1f_cta_shapes.txt): one CTA of 320 threads
// per SM 25.5 ms; two CTAs of 160 threads 26.6 ms; two CTAs of 192 threads (which balances the
// butterfly rounds 5/5/5/5 over the four SM sub-partitions instead of 3/3/2/2) 26.7 ms, because
// the second set of pair accumulators spills (300 B / thread to local memory = L2 traffic).
// <= 390 threads per SM keeps 168 registers per thread.
constexpr int k1f_threads(int R1) { return 16 * R1; }
constexpr int k1f_min_blocks(int NT) { return 390 / NT < 1 ? 1 : 390 / NT; }

// ---------------------------------------------------------------------------
// P1 twiddles: tw[k] = om^(2k + r), k < R1, from om = w_{2H}^j, two-level.
// ---------------------------------------------------------------------------
template <int R1>
TA_HD void k1f_p1_twiddles(cd om, int r, cd* e, cd* g) {
    // e[b] = om^(2b + r), b < 4 ; g[a] = om^(8a), a < ceil(R1/4)
    cd w2 = cmul(om, om), w4 = cmul(w2, w2), w6 = cmul(w4, w2), w8 = cmul(w4, w4);
    if (r) {
        e[0] = om; e[1] = cmul(w2, om); e[2] = cmul(w4, om); e[3] = cmul(w6, om);
    } else {
        e[0] = cmake<double>(1.0, 0.0); e[1] = w2; e[2] = w4; e[3] = w6;
    }
    constexpr int NG = (R1 + 3) / 4;
    g[0] = cmake<double>(1.0, 0.0);
    if (NG > 1) g[1] = w8;
    if (NG > 2) g[2] = cmul(w8, w8);
    if (NG > 3) g[3] = cmul(g[2], w8);
    if (NG > 4) g[4] = cmul(g[2], g[2]);
    if (NG > 5) g[5] = cmul(g[4], w8);
}

// ---------------------------------------------------------------------------
// The kernel body for one CTA.  Ctx supplies sync() and shfl_xor16(); on the
// device these are __syncthreads / __shfl_xor_sync, in tests/emu they are
// cooperative-fiber versions so the very same code runs on the CPU.
// ---------------------------------------------------------------------------
// VAR bits (experiments; 0 = the measured default):
//   1  token-ordered shared-memory / global load phases after a CTA barrier: the warps of one SM sub-partition
//      (w, w + 4, w + 8) issue their loads one rank after the other instead of all at once, so rank 0 is already
//      on the FP64 pipe while the LSU serves rank 1
//   2  staged output: V_0 waits in a second shared buffer instead of the global row, the finished row is
//      normalised with a computed 1 / (L (T - k)) and leaves the SM as one bulk (TMA) store of the per-particle
//      row plus one bulk reduce-add into the per-CTA partial row, issued by a warp that is idle in P1':
//      no parked-row / table / partial-row loads and no L2 round trips in the output phase
//   4  series prefetch: while a chain runs P2 / P3, the bulk-copy engine (TMA) brings the next chain's series into a
//      second shared buffer (one mbarrier, phase per chain), so P1 is a shared -> registers -> shared pass like the
//      others and no warp waits for L2 / HBM
constexpr int K1F_VAR_TURNS = 1;
constexpr int K1F_VAR_STAGED = 2;
//   8  deferred twiddles: the P2 -> P3 twiddles w_256^(j k) are not multiplied onto the P2 outputs (15 table loads and
//      60 FP64 instructions per butterfly) but ride on the FMAs of the P3 butterfly (TwDit in dft_regs.cuh: 8 table loads,
//      +32 instructions); the inverse P2' uses the same butterfly with the conjugate table
constexpr int K1F_VAR_PREFETCH = 4;
//  16  output phase in chunks of 10 elements per thread instead of 5 (two L2 round trips instead of four)
//  32  1 / (L (T - k)) computed (hardware reciprocal seed + two Newton steps) instead of loaded from the table
constexpr int K1F_VAR_DEFTW = 8;
constexpr int K1F_VAR_OUT10 = 16;
constexpr int K1F_VAR_RCPINV = 32;

// Rank r (warps 4r .. 4r+3 of the first NW warps) may issue its loads once rank r-1 has issued its own.
// Named barriers id0 + r; consecutive uses of one site are separated by a CTA barrier.
template <class Ctx, int NW>
TA_HD void k1f_turn_wait(int warp, int id0) {
    const int rank = warp >> 2;
    if (rank > 0 && warp < NW) {
        const int prev = 4, cur = (NW - 4 * rank) < 4 ? (NW - 4 * rank) : 4;
        Ctx::bar_sync(id0 + rank, 32 * (prev + cur));
    }
}
template <class Ctx, int NW>
TA_HD void k1f_turn_pass(int warp, int id0) {
    const int rank = warp >> 2;
    if (warp < NW && 4 * (rank + 1) < NW) {
        const int nxt = (NW - 4 * (rank + 1)) < 4 ? (NW - 4 * (rank + 1)) : 4;
        Ctx::bar_arrive(id0 + rank + 1, 32 * (4 + nxt));
    }
}

template <int R1, int NT, class Ctx, bool PROF = false, int VAR = 0>
TA_HD void k1f_body(const K1FArgs& A, unsigned char* smem_raw, int tid, int bid, int nblk) {
    constexpr int H = 256 * R1;
    constexpr int NV = 16 * R1;          // radix-16 butterflies per pass ("virtual threads")
    constexpr int NB = (NV + NT - 1) / NT;   // butterfly rounds of P2 / P3; a thread owns vt = tid + it NT < NV
    constexpr int NG = (R1 + 3) / 4;
    static_assert(NT % 32 == 0 && NV % 32 == 0, "the lane exchange needs whole warps");
    constexpr bool TURNS = (VAR & K1F_VAR_TURNS) != 0 && NT == NV && NT >= 256;   // one butterfly per thread, >= 2 ranks
    constexpr int NW = NT / 32, NW1 = 256 / 32;                                     // warps of P2 / P3 and of P1 / P1'
    const int warp = tid >> 5;
    cd* buf = reinterpret_cast<cd*>(smem_raw);       // H + H/16 elements, padded layout
    cd* s_om = buf + (H + H / 16);                   // 256
    cd* s_tw2 = s_om + 256;                          // 240
    cd* stg = s_tw2 + 240;                           // STAGED: H elements, plain layout (V_0, then the finished row)
    constexpr bool STAGED = (VAR & K1F_VAR_STAGED) != 0;
    constexpr bool PREF = (VAR & K1F_VAR_PREFETCH) != 0;
    cd* pre = stg + (STAGED ? H : 0);                // PREF: H elements, the series of the coming chain as it lies in HBM
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(pre + H);   // PREF: "series has landed"
    unsigned pre_phase = 0;
    // the warp that issues the bulk store / reduce of a finished row: the last one, idle in P1' when NT > 256
    constexpr int NPART = NT < 256 ? NT : 256;       // threads that take part in P1'
    constexpr int WISS = NT / 32 - 1;
    constexpr bool ISS_IDLE = NT > 256;
    const double Ld = (double)(4 * H);

    for (int i = tid; i < 256; i += NT) s_om[i] = A.omega[i];
    constexpr bool DEFTW = (VAR & K1F_VAR_DEFTW) != 0;
    for (int i = tid; i < 240; i += NT) s_tw2[i] = DEFTW ? (i < 128 ? A.tw8[i] : cmake<double>(0.0, 0.0)) : A.tw2[i];
    if (PREF && tid == 0) Ctx::mbar_init(mbar);
    Ctx::sync();

    const int nh = A.nh;
    const unsigned ser_bytes = (unsigned)nh * (unsigned)sizeof(cd);
    if (PREF && tid == 0 && bid < A.natoms) Ctx::bulk_load(pre, A.series + (size_t)bid * A.D * A.Tld, ser_bytes, mbar);
    const int j2 = tid & 15;                         // NT is a multiple of 16: the same for every owned butterfly
    cd* part = reinterpret_cast<cd*>(A.partial + (size_t)bid * A.Tld);
    const cd* inv2 = reinterpret_cast<const cd*>(A.inv);

    long long tprev = 0;
    long long* prof = (PROF && A.prof != nullptr && tid == 0) ? A.prof + 32 * (size_t)bid : nullptr;
    if (PROF && prof) tprev = Ctx::clock();
    // debug instrumentation (PROF instantiation only): clocks of thread 0 between phase boundaries;
    // `dep` makes the clock read wait for a value (e.g. the last load of a batch)
#define K1F_TICK(ph, dep) do { if (PROF && prof) { const long long tn_ = Ctx::clock_after(dep); prof[r * 16 + (ph)] += tn_ - tprev; tprev = tn_; } } while (0)

    if (A.stagger > 0 && 2 * bid >= nblk) Ctx::spin(A.stagger);
    for (int atom = bid; atom < A.natoms; atom += nblk) {
        const double* ser = A.series + (size_t)atom * A.D * A.Tld;
        cd* row = reinterpret_cast<cd*>(A.by_particle + (size_t)atom * A.Tld);
        if (A.prefetch && atom + nblk < A.natoms) Ctx::prefetch_l2(ser + (size_t)nblk * A.D * A.Tld, A.D * A.Tld * sizeof(double), tid, NT);
        for (int r = 0; r < 2; ++r) {
            double acc_s[NB][8], acc_d[NB][8], acc8[NB];
#pragma unroll
            for (int it = 0; it < NB; ++it) {
                acc8[it] = 0.0;
#pragma unroll
                for (int m = 0; m < 8; ++m) { acc_s[it][m] = 0.0; acc_d[it][m] = 0.0; }
            }

            for (int d = 0; d < A.D; ++d) {
                // ---------------- P1: global -> registers -> shared
                const cd* src = reinterpret_cast<const cd*>(ser + (size_t)d * A.Tld);
                for (int j = tid; j < 256; j += NT) {
                    cd x[R1];
                    if (TURNS) k1f_turn_wait<Ctx, NW1>(warp, 1);
                    if (PREF) {
                        if (j == tid) Ctx::mbar_wait(mbar, pre_phase);
#pragma unroll
                        for (int q = 0; q < R1; ++q) {
                            const int n = j + 256 * q;
                            x[q] = (n < nh) ? pre[n] : cmake<double>(0.0, 0.0);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < R1; ++q) {
                            const int n = j + 256 * q;
                            x[q] = (n < nh) ? Ctx::ld_stream(src + n) : cmake<double>(0.0, 0.0);
                        }
                    }
                    if (TURNS) k1f_turn_pass<Ctx, NW1>(warp, 1);
                    K1F_TICK(0, x[R1 - 1].y + x[0].x);
                    if (r) {
                        static_for<1, R1>([&](auto iq) {
                            constexpr int q = decltype(iq)::value;
                            x[q] = mul_tw<q, 2 * R1, -1>(x[q]);
                        });
                    }
                    Dft<R1, -1>::run(x);
                    cd e[4], g[NG];
                    k1f_p1_twiddles<R1>(s_om[j], r, e, g);
                    cd* dst = buf + j + (j >> 4);
#pragma unroll
                    for (int k = 0; k < R1; ++k) {
                        cd y = x[k];
                        if (k >= 4) y = cmul(y, g[k >> 2]);
                        if (r || (k & 3)) y = cmul(y, e[k & 3]);
                        x[k] = y;
                    }
                    K1F_TICK(1, x[R1 - 1].y + x[1].x);
#pragma unroll
                    for (int k = 0; k < R1; ++k) dst[272 * k] = x[k];
                }
                Ctx::sync();
                if (PREF) {
                    // every P1 thread has read the prefetch buffer: hand it to the bulk-copy engine for the next chain
                    pre_phase ^= 1u;
                    if (tid == NT - 1) {
                        const double* nxt = nullptr;
                        if (d + 1 < A.D) nxt = ser + (size_t)(d + 1) * A.Tld;
                        else if (r == 0) nxt = ser;
                        else if (atom + nblk < A.natoms) nxt = ser + (size_t)nblk * A.D * A.Tld;
                        if (nxt) Ctx::bulk_load(pre, nxt, ser_bytes, mbar);
                    }
                }
                K1F_TICK(2, 0.0);
                // ---------------- P2: radix 16, stride 16
#pragma unroll
                for (int it = 0; it < NB; ++it) {
                    const int vt = tid + it * NT;
                    if (NV % NT != 0 && vt >= NV) break;
                    const int blk2 = (int)((A.map[r * NV + vt] >> 16) & 0xffu);
                    const int p2base = blk2 * 272 + j2;          // padded address of (blk*256 + j), + 17 q
                    cd x[16];
                    if (TURNS) k1f_turn_wait<Ctx, NW>(warp, 4);
#pragma unroll
                    for (int q = 0; q < 16; ++q) x[q] = buf[p2base + 17 * q];
                    if (TURNS) k1f_turn_pass<Ctx, NW>(warp, 4);
                    K1F_TICK(3, x[15].y + x[0].x);
                    Dft<16, -1>::run(x);
                    if (!DEFTW) {
#pragma unroll
                        for (int k = 1; k < 16; ++k) x[k] = cmul(x[k], s_tw2[(k - 1) * 16 + j2]);
                    }
                    K1F_TICK(4, x[15].y + x[1].x);
#pragma unroll
                    for (int k = 0; k < 16; ++k) buf[p2base + 17 * k] = x[k];
                }
                Ctx::sync_warp();        // P3 of this warp reads what this warp's P2 wrote
                K1F_TICK(5, 0.0);
                // ---------------- P3: radix 16, stride 1, + pair accumulation
                static_for<0, NB>([&](auto iit) {
                    constexpr int it = decltype(iit)::value;
                    const int vt = tid + it * NT;
                    if (NV % NT != 0 && vt >= NV) return;
                    const uint32_t mp = A.map[r * NV + vt];
                    const cd wb = A.wbase[r * NV + vt];
                    const int p3base = (int)(mp & 0xffffu) * 17;
                    const bool self0 = (mp & K1F_SELF0) != 0, self8 = (mp & K1F_SELF8) != 0;
                    cd v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = buf[p3base + q];
                    K1F_TICK(6, v[15].y + v[0].x);
                    if (DEFTW) TwDit<16, 0, -1, false, 16>::run(v, s_tw2 + (int)(mp & 15u));   // om = w_256^k2
                    else Dft<16, -1>::run(v);
                    static_for<0, 8>([&](auto im) {
                        constexpr int m = decltype(im)::value;
                        cd snd = v[15 - m], rec;
                        rec.x = Ctx::shfl_xor16(snd.x);
                        rec.y = Ctx::shfl_xor16(snd.y);
                        if (self8) rec = snd;
                        if (self0) rec = v[(16 - m) & 15];
                        const cd w = mul_tw<m, 32, -1>(wb);
                        const cd U = v[m];
                        const double nu = cnorm2(U), nv = cnorm2(rec);
                        const double B = U.x * rec.y + U.y * rec.x;
                        acc_s[it][m] += nu + nv;
                        acc_d[it][m] += 2.0 * w.x * B + w.y * (nu - nv);
                    });
                    if (self0) acc8[it] += 2.0 * cnorm2(v[8]);
                    K1F_TICK(7, acc_d[it][7] + acc_s[it][0]);
                });
                if (d + 1 < A.D) Ctx::sync();   // P1 of the next series overwrites the buffer
                K1F_TICK(8, 0.0);
            }

            // ---------------- inverse: build from the accumulators, P3'
            static_for<0, NB>([&](auto iit) {
                constexpr int it = decltype(iit)::value;
                const int vt = tid + it * NT;
                if (NV % NT != 0 && vt >= NV) return;
                const uint32_t mp = A.map[r * NV + vt];
                const cd wb = A.wbase[r * NV + vt];
                const int p3base = (int)(mp & 0xffffu) * 17;
                const bool self0 = (mp & K1F_SELF0) != 0, self8 = (mp & K1F_SELF8) != 0;
                cd v[16], ap[8], rc[8];
                static_for<0, 8>([&](auto im) {
                    constexpr int m = decltype(im)::value;
                    const cd w = mul_tw<m, 32, -1>(wb);
                    const double sig = acc_s[it][m], del = acc_d[it][m];
                    v[m] = cmake<double>(sig + w.y * del, w.x * del);
                    ap[m] = cmake<double>(sig - w.y * del, w.x * del);
                    rc[m].x = Ctx::shfl_xor16(ap[m].x);
                    rc[m].y = Ctx::shfl_xor16(ap[m].y);
                });
                static_for<8, 16>([&](auto ii) {
                    constexpr int idx = decltype(ii)::value;
                    cd val = rc[15 - idx];
                    if (self8) val = ap[15 - idx];
                    if (self0) val = (idx == 8) ? cmake<double>(acc8[it], 0.0) : ap[16 - idx];
                    v[idx] = val;
                });
                Dft<16, +1>::run(v);
                K1F_TICK(9, v[15].y + v[0].x);
#pragma unroll
                for (int q = 0; q < 16; ++q) buf[p3base + q] = v[q];
            });
            Ctx::sync_warp();
            K1F_TICK(10, 0.0);
            // ---------------- P2'
#pragma unroll
            for (int it = 0; it < NB; ++it) {
                const int vt = tid + it * NT;
                if (NV % NT != 0 && vt >= NV) break;
                const int blk2 = (int)((A.map[r * NV + vt] >> 16) & 0xffu);
                const int p2base = blk2 * 272 + j2;
                cd x[16];
                if (DEFTW) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) x[k] = buf[p2base + 17 * k];
                    TwDit<16, 0, +1, true, 16>::run(x, s_tw2 + j2);                          // om = conj(w_256^j2)
                } else {
                    x[0] = buf[p2base];
#pragma unroll
                    for (int k = 1; k < 16; ++k) x[k] = cmulc(buf[p2base + 17 * k], s_tw2[(k - 1) * 16 + j2]);
                    Dft<16, +1>::run(x);
                }
                K1F_TICK(11, x[15].y + x[0].x);
#pragma unroll
                for (int q = 0; q < 16; ++q) buf[p2base + 17 * q] = x[q];
            }
            // STAGED: the bulk operations of the previous particle have read (and reduced) the staging buffer
            if (STAGED && tid == 32 * WISS) Ctx::bulk_wait_all();
            Ctx::sync();
            K1F_TICK(12, 0.0);
            // ---------------- P1' + output
            for (int j = tid; j < 256; j += NT) {
                cd e[4], g[NG];
                k1f_p1_twiddles<R1>(s_om[j], r, e, g);
                const cd* srcb = buf + j + (j >> 4);
                cd x[R1];
                if (TURNS) k1f_turn_wait<Ctx, NW1>(warp, 7);
#pragma unroll
                for (int k = 0; k < R1; ++k) x[k] = srcb[272 * k];
                if (TURNS) k1f_turn_pass<Ctx, NW1>(warp, 7);
#pragma unroll
                for (int k = 0; k < R1; ++k) {
                    cd y = x[k];
                    if (k >= 4) y = cmulc(y, g[k >> 2]);
                    if (r || (k & 3)) y = cmulc(y, e[k & 3]);
                    x[k] = y;
                }
                Dft<R1, +1>::run(x);
                if (r) {
                    static_for<1, R1>([&](auto iq) {
                        constexpr int q = decltype(iq)::value;
                        x[q] = mul_tw<q, 2 * R1, +1>(x[q]);
                    });
                }
                // x[q] = V_r[n] (r = 1: already multiplied by conj(w_L^{2n})), n = j + 256 q
                K1F_TICK(13, x[R1 - 1].y + x[0].x);
                if (STAGED) {
                    if (r == 0) {
#pragma unroll
                        for (int q = 0; q < R1; ++q) stg[j + 256 * q] = x[q];     // V_0 waits here for residue 1
                    } else {
#pragma unroll
                        for (int q = 0; q < R1; ++q) {
                            const int n = j + 256 * q;
                            const int k = 2 * n;
                            const cd v0 = stg[n];
                            const double sx = k < A.T ? Ctx::rcp(Ld * (double)(A.T - k)) : 0.0;
                            const double sy = k + 1 < A.T ? Ctx::rcp(Ld * (double)(A.T - k - 1)) : 0.0;
                            stg[n] = cmake<double>((v0.x + x[q].x) * sx, (v0.y + x[q].y) * sy);
                        }
                    }
                } else if (r == 0) {
#pragma unroll
                    for (int q = 0; q < R1; ++q) {
                        const int n = j + 256 * q;
                        if (n < nh) row[n] = x[q];           // parked raw; finished by residue 1
                    }
                } else {
                    // Park the finished V_1 values in this thread's own buffer slots, then stream the
                    // output in chunks with all global loads of a chunk issued first: with x[] out of
                    // the registers there is room to keep a whole chunk of L2 round trips in flight
                    // (the row / part stores may alias the loads as far as the compiler knows).
                    cd* own = buf + j + (j >> 4);
#pragma unroll
                    for (int q = 0; q < R1; ++q) own[272 * q] = x[q];
                    Ctx::compiler_fence();
                    constexpr bool OUT10 = (VAR & K1F_VAR_OUT10) != 0, RCPINV = (VAR & K1F_VAR_RCPINV) != 0;
                    constexpr int QB = OUT10 ? ((R1 % 10 == 0) ? 10 : 8) : ((R1 % 5 == 0) ? 5 : 4);
                    static_for<0, (R1 + QB - 1) / QB>([&](auto ic) {
                        constexpr int q0 = decltype(ic)::value * QB;
                        constexpr int nq = (R1 - q0) < QB ? (R1 - q0) : QB;
                        cd a[nq], sc[nq], ps[nq];
#pragma unroll
                        for (int i = 0; i < nq; ++i) {
                            const int n = j + 256 * (q0 + i);
                            const int nc = n < nh ? n : nh - 1;      // clamped: the loads stay branch-free
                            a[i] = Ctx::ld_stream(row + nc); ps[i] = Ctx::ld_stream(part + nc);
                            if (!RCPINV) sc[i] = inv2[nc];
                        }
#pragma unroll
                        for (int i = 0; i < nq; ++i) {
                            const int n = j + 256 * (q0 + i);
                            if (n < nh) {
                                if (RCPINV) {
                                    const int k = 2 * n;
                                    sc[i].x = Ctx::rcp(Ld * (double)(A.T - k));
                                    sc[i].y = k + 1 < A.T ? Ctx::rcp(Ld * (double)(A.T - k - 1)) : 0.0;
                                }
                                const cd v1 = own[272 * (q0 + i)];
                                const cd o = cmake<double>((a[i].x + v1.x) * sc[i].x, (a[i].y + v1.y) * sc[i].y);
                                row[n] = o;
                                part[n] = cmake<double>(ps[i].x + o.x, ps[i].y + o.y);
                            }
                        }
                    });
                }
            }
            if (STAGED && r == 1) {
                // every P1' thread has written its part of the finished row -> one warp hands it to the
                // bulk-copy engine: row store + reduce-add into this CTA's partial row (fixed order:
                // the previous particle's group has completed, see bulk_wait_all above)
                const unsigned nbytes = (unsigned)nh * (unsigned)sizeof(cd);
                if (ISS_IDLE) {
                    if (tid < NPART) { Ctx::fence_async_smem(); Ctx::bar_arrive(10, NPART + 32); }
                    else if ((tid >> 5) == WISS) {
                        Ctx::bar_sync(10, NPART + 32);
                        if (tid == 32 * WISS) Ctx::bulk_store_and_add(row, part, stg, nbytes);
                    }
                } else {
                    Ctx::fence_async_smem();
                    Ctx::sync();
                    if (tid == 32 * WISS) Ctx::bulk_store_and_add(row, part, stg, nbytes);
                }
            }
            // no barrier here: P1 of the next chain writes exactly the elements this
            // thread has just read in P1'
            K1F_TICK(14, 0.0);
        }
    }
    if (STAGED && tid == 32 * WISS) Ctx::bulk_wait_all();
#undef K1F_TICK
}

// ---------------------------------------------------------------------------
// Host-side plan for the fast path.
// ---------------------------------------------------------------------------
struct K1FastPlan {
    int R1 = 0, H = 0, L = 0, NT = 0, nh = 0;
    std::vector<double> omega;    // 256 x (re, im)
    std::vector<double> tw2;      // 240 x (re, im)
    std::vector<double> tw8;      // 128 x (re, im)
    std::vector<uint32_t> map;    // 2 x NT
    std::vector<double> wbase;    // 2 x NT x (re, im)
    std::vector<double> inv;      // Tld
};

// radices R1 the library instantiates (even: the lane exchange needs whole warps)
inline const int* k1f_supported_r1(int* n) {
    static const int r1s[] = {4, 6, 8, 10, 12, 16, 20};
    *n = (int)(sizeof(r1s) / sizeof(r1s[0]));
    return r1s;
}

// R1 for a series of T frames, or 0 when the general kernel should be used
// (tiny problems, > 34 % padding, or longer than the largest instantiation).
inline int k1f_choose_r1(int64_t T) {
    const int64_t nh = (T + 1) / 2;
    int n;
    const int* r1s = k1f_supported_r1(&n);
    for (int i = 0; i < n; ++i) {
        const int64_t H = 256 * (int64_t)r1s[i];
        if (H >= nh) return (3 * H <= 4 * nh + 3) ? r1s[i] : 0;
    }
    return 0;
}

inline int k1f_build_plan(int64_t T, int64_t Tld, int R1, K1FastPlan* p) {
    if (R1 < 2 || (R1 & 1) || T < 1 || (T + 1) / 2 > 256 * (int64_t)R1) return TA_ERR_INVALID;
    p->R1 = R1; p->H = 256 * R1; p->L = 4 * p->H; p->NT = 16 * R1; p->nh = (int)((T + 1) / 2);
    const int64_t L = p->L;
    p->omega.resize(2 * 256);
    for (int j = 0; j < 256; ++j) ta_twiddle(2 * j, L, &p->omega[2 * j], &p->omega[2 * j + 1]);   // w_{2H}^j = w_L^{2j}
    p->tw2.resize(2 * 240);
    for (int k = 1; k < 16; ++k)
        for (int j = 0; j < 16; ++j)
            ta_twiddle((int64_t)j * k, 256, &p->tw2[2 * ((k - 1) * 16 + j)], &p->tw2[2 * ((k - 1) * 16 + j) + 1]);
    p->tw8.resize(2 * 128);
    for (int j = 0; j < 16; ++j) {
        const int64_t ex[8] = {8 * j, 4 * j, 2 * j, 2 * j + 32, j, j + 16, j + 32, j + 48};   // exponents of w_256
        for (int v = 0; v < 8; ++v) ta_twiddle(ex[v], 256, &p->tw8[2 * (v * 16 + j)], &p->tw8[2 * (v * 16 + j) + 1]);
    }
    const int NT = p->NT;
    p->map.assign(2 * (size_t)NT, 0);
    p->wbase.assign(4 * (size_t)NT, 0.0);
    // residue 0: owner / partner butterflies (see DESIGN.md "K1 fast path").  Pair index pidx = 16 w + i
    // (warp w, i = lane & 15): warp 0 holds the two self-paired blocks k1 = 0 (i < 8) and k1 = R1/2
    // (i >= 8), warp w >= 1 the blocks k1 = w (owners) and R1 - w (partners).
    std::vector<uint32_t> own(8 * R1), par(8 * R1);
    for (int k2 = 0; k2 < 8; ++k2) {
        own[k2] = (uint32_t)k2 | (k2 == 0 ? K1F_SELF0 : 0u);
        par[k2] = (k2 == 0) ? (8u | K1F_SELF8) : (uint32_t)(16 - k2);
        own[8 + k2] = (uint32_t)((R1 / 2) * 16 + k2);
        par[8 + k2] = (uint32_t)((R1 / 2) * 16 + 15 - k2);
    }
    for (int k1 = 1; k1 < R1 / 2; ++k1)
        for (int k2 = 0; k2 < 16; ++k2) {
            own[16 * k1 + k2] = (uint32_t)(k1 * 16 + k2);
            par[16 * k1 + k2] = (uint32_t)((R1 - k1) * 16 + 15 - k2);
        }
    for (int tid = 0; tid < NT; ++tid) {
        const int w = tid >> 5, l = tid & 31, pidx = w * 16 + (l & 15), side = l >> 4;
        p->map[tid] = side ? par[pidx] : own[pidx];
        p->map[NT + tid] = (uint32_t)(side ? (NT - 1 - pidx) : pidx);
        // P2 block of this thread (one block per half-warp), the same two blocks the warp's P3 butterflies cover
        const uint32_t blk2_r0 = (w == 0) ? (side ? (uint32_t)(R1 / 2) : 0u) : (side ? (uint32_t)(R1 - w) : (uint32_t)w);
        const uint32_t blk2_r1 = side ? (uint32_t)(R1 - 1 - w) : (uint32_t)w;
        p->map[tid] |= blk2_r0 << 16;
        p->map[NT + tid] |= blk2_r1 << 16;
        for (int r = 0; r < 2; ++r) {
            const int blk3 = (int)(p->map[r * NT + tid] & 0xffffu);
            const int k1 = blk3 >> 4, k2 = blk3 & 15;
            const int64_t G0 = k1 + (int64_t)R1 * k2;
            ta_twiddle(2 * G0 + r, L, &p->wbase[2 * (r * NT + tid)], &p->wbase[2 * (r * NT + tid) + 1]);
        }
    }
    p->inv.assign((size_t)Tld, 0.0);
    for (int64_t k = 0; k < T; ++k) p->inv[k] = 1.0 / ((double)L * (double)(T - k));
    return TA_OK;
}

}  // namespace ta

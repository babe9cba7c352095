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
//       P1  radix R1, stride 256   fused with the series load and the residue twist
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
//     buffer hand-over to the next series, and P2' -> P1'.
//   * All butterflies are register DFTs with compile-time constants
//     (dft_regs.cuh); P2 twiddles come from a 240-entry table, P1 twiddles are
//     powers of one table entry per thread.
//   * Shared buffer layout: element e lives at e + (e >> 4) (one pad element
//     per 16), which makes all three access patterns bank-conflict free.
//   * PREF: while a chain runs P2 / P3 the bulk-copy engine (TMA, cp.async.bulk completing on one mbarrier, one
//     phase per chain) brings the series of the next chain -- next dimension, next residue or next particle -- into
//     a second shared buffer, so P1 is a shared -> registers -> shared pass like the others and no warp waits for
//     L2 / HBM.
//
//   * TMEM (FP64, R1 = 16 / 20): the per-thread streams of the output stage -- the parked residue-0 result, the CTA's particle
//     sums, the 1/(L(T-k)) table -- live in tensor memory (tcgen05.st / ld on each thread's own lane) instead of going through
//     L2: 80 instead of 480 KB per particle between the SM and L2 (23.8 -> 21.5 ms at 100k x 10k).
//
// The body is a template on the arithmetic type RT: double (the reference's precision, rel. 1e-10) or float (the
// optional FP32 mode, stated tolerance 1e-5; the series are then stored as float in HBM, the per-particle rows and
// the particle sums stay double).  The round-1 experiment variants (token-ordered loads, staged bulk output,
// deferred twiddles, other CTA shapes, the four-pass radix-8 kernel) were measured slower and are gone from the
// product; they are in the history (commit ba2c124) and their measurements in profiles/r01_k1f_cta_shapes.txt.
#pragma once
#include <cstdint>
#include <vector>
#include "ta_common.cuh"
#include "dft_regs.cuh"
#include "fft_plan.h"

namespace ta {

template <typename RT>
struct K1FArgs {
    const RT* series;            // [natoms][DS][Tld]: the first D rows of a particle are transformed (DS > D: Helfand keeps a row of sum_d g^2 behind them)
    double* by_particle;         // [natoms][Tld]
    double* partial;             // [grid][Tld]
    const cplx<RT>* omega;       // [256]      w_{2H}^j
    const cplx<RT>* tw2;         // [15][16]   w_256^{j k}, k = 1..15
    const uint32_t* map;         // [2][16 R1] per residue: bits 0-15 P3 butterfly of a thread, 16-23 its P2 block, 30/31 flags
    const cplx<RT>* wbase;       // [2][16 R1] w_L^{2 G0 + r} of that butterfly
    const RT* inv;               // [Tld]      1 / (L (T - k)), 0 beyond T
    int natoms, D, DS, T, nh;
    long long Tld;
};

constexpr uint32_t K1F_SELF0 = 1u << 30;   // butterfly holding g = 0 (pairs k <-> 16-k in-thread)
constexpr uint32_t K1F_SELF8 = 1u << 31;   // butterfly holding g = H/2 (pairs k <-> 15-k in-thread)

// dynamic shared memory: FFT buffer (padded) + omega + tw2 (+ series buffer and its mbarrier), in complex elements of RT
constexpr int k1f_smem_bytes(int R1, bool pref, int real_bytes) {
    return (256 * R1 + 16 * R1 + 256 + 240 + (pref ? 256 * R1 + 2 : 0)) * 2 * real_bytes;
}
// Threads per CTA (NT): one radix-16 butterfly of P2 / P3 per thread (a pass has NV = 16 R1 of them; whole warps,
// so the lane exchange of P3 stays inside a warp).  Measured on B200 at R1 = 20, 100k x 10k
// (profiles/r01_k1f_cta_shapes.txt): one CTA of 320 threads per SM beats two of 160 or 192 and one of 128 or 256.
constexpr int k1f_threads(int R1) { return 16 * R1; }
// resident CTAs per SM the kernel is compiled for: FP64 <= 390 threads per SM keeps 168 registers per thread;
// FP32 (half the registers per value) twice that
constexpr int k1f_min_blocks(int NT, int real_bytes) {
    return (real_bytes == 8 ? 390 : 780) / NT < 1 ? 1 : (real_bytes == 8 ? 390 : 780) / NT;
}
// bulk series prefetch: on where its buffer does not cost a resident CTA (FP64: R1 >= 8; measured 25.4 -> 24.2 ms
// at 100k x 10k, 25.1 -> 23.6 ms at 200k x 5k; off at R1 = 4 and 6: 6.15 -> 6.21 ms at 150k x 2,000, 6.41 -> 6.87 ms
// at 100k x 3,000); FP32: on everywhere.
constexpr bool k1f_prefetch(int R1, int real_bytes) { return real_bytes == 4 || R1 >= 8; }

// ---------------------------------------------------------------------------
// P1 twiddles: tw[k] = om^(2k + r), k < R1, from om = w_{2H}^j, two-level.
// ---------------------------------------------------------------------------
template <int R1, class C>
TA_HD void k1f_p1_twiddles(C om, int r, C* e, C* g) {
    using RT = real_t<C>;
    // e[b] = om^(2b + r), b < 4 ; g[a] = om^(8a), a < ceil(R1/4)
    C w2 = cmul(om, om), w4 = cmul(w2, w2), w6 = cmul(w4, w2), w8 = cmul(w4, w4);
    if (r) {
        e[0] = om; e[1] = cmul(w2, om); e[2] = cmul(w4, om); e[3] = cmul(w6, om);
    } else {
        e[0] = cmake<RT>((RT)1, (RT)0); e[1] = w2; e[2] = w4; e[3] = w6;
    }
    constexpr int NG = (R1 + 3) / 4;
    g[0] = cmake<RT>((RT)1, (RT)0);
    if (NG > 1) g[1] = w8;
    if (NG > 2) g[2] = cmul(w8, w8);
    if (NG > 3) g[3] = cmul(g[2], w8);
    if (NG > 4) g[4] = cmul(g[2], g[2]);
    if (NG > 5) g[5] = cmul(g[4], w8);
}

// ---------------------------------------------------------------------------
// The kernel body for one CTA.  Ctx supplies sync(), shfl_xor16(), the bulk-copy and mbarrier calls; on the
// device these are the hardware's, in tests/emu they are cooperative-fiber versions so the very same code runs on
// the CPU.
// ---------------------------------------------------------------------------
// PART = false: the per-CTA particle-sum row is not wanted (ta_helfand_fft: K5 forms its own sums from the finished rows),
// so the output stage neither reads nor writes it -- a third of that stage's L2 traffic.
// TMEM = true (FP64, one P1 column per thread: NT >= 256, R1 a multiple of 4): the three per-thread streams of the output stage
// live in TENSOR MEMORY instead of going through L2 -- the parked V_0 of residue 0, the CTA's particle-sum row and the
// 1/(L(T-k)) table, 4 R1 32-bit columns each.  Every P1' thread only ever touches the values of its own column j, which is
// exactly what tcgen05.st / tcgen05.ld offer (a thread reads and writes its own TMEM lane; a warp owns the 32 lanes of
// its quarter, the two P1 warps of a quarter take 256 columns each).  The stage then moves 80 KB per particle between the SM
// and L2 (the finished row) instead of 480 KB, and has no L2 round trip to wait for.
template <int R1, int NT, class Ctx, typename RT, bool PREF, bool PART = true, bool TMEM = false>
TA_HD void k1f_body(const K1FArgs<RT>& A, unsigned char* smem_raw, int tid, int bid, int nblk) {
    using C = cplx<RT>;
    static_assert(!TMEM || (PREF && sizeof(RT) == 8 && NT >= 256 && R1 % 4 == 0 && 12 * R1 <= 256),
                  "tensor-memory build: FP64, one P1 column per thread, three 4 R1-column arrays in a 256-column half");
    constexpr int H = 256 * R1;
    constexpr int NV = 16 * R1;          // radix-16 butterflies per pass ("virtual threads")
    constexpr int NB = (NV + NT - 1) / NT;   // butterfly rounds of P2 / P3; a thread owns vt = tid + it NT < NV
    constexpr int NG = (R1 + 3) / 4;
    static_assert(NT % 32 == 0 && NV % 32 == 0, "the lane exchange needs whole warps");
    C* buf = reinterpret_cast<C*>(smem_raw);         // H + H/16 elements, padded layout
    C* s_om = buf + (H + H / 16);                    // 256
    C* s_tw2 = s_om + 256;                           // 240
    C* pre = s_tw2 + 240;                            // PREF: H elements, the series of the coming chain as it lies in HBM
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(pre + H);   // PREF: "series has landed" (16-byte slot)
    unsigned pre_phase = 0;
    const C czero = cmake<RT>((RT)0, (RT)0);

    for (int i = tid; i < 256; i += NT) s_om[i] = A.omega[i];
    for (int i = tid; i < 240; i += NT) s_tw2[i] = A.tw2[i];
    if (PREF && tid == 0) Ctx::mbar_init(mbar);
    Ctx::sync();

    const int nh = A.nh;
    // TMEM: 512 columns allocated by warp 0; this thread's slice = lane quarter of its warp, column half of its warp pair;
    // columns [0, 4 R1) parked V_0, [4 R1, 8 R1) particle sums, [8 R1, 12 R1) 1 / (L (T - k)), one complex double = 4 columns
    uint32_t tm_base = 0, tm = 0;
    constexpr uint32_t TM_PART = 4 * R1, TM_INV = 8 * R1;
    if (TMEM) {
        tm_base = Ctx::tmem_alloc(reinterpret_cast<uint32_t*>(mbar + 2), tid);
        tm = tm_base + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)((tid >> 7) & 1) * 256u;
        if (tid < 256) {
            const C* inv2i = reinterpret_cast<const C*>(A.inv);
            static_for<0, R1 / 4>([&](auto ic) {
                constexpr int c = decltype(ic)::value;
                cd z[4], sc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = tid + 256 * (4 * c + i);
                    z[i] = cmake<double>(0.0, 0.0);
                    sc[i] = z[i];
                    if (n < nh) { const C v = inv2i[n]; sc[i] = cmake<double>((double)v.x, (double)v.y); }   // 0 beyond the series: nothing is added there
                }
                if (PART) Ctx::tmem_st4(tm + TM_PART + 16 * c, z);
                Ctx::tmem_st4(tm + TM_INV + 16 * c, sc);
            });
            Ctx::tmem_wait_st();
        }
    }
    // bytes of one series the bulk copy moves: nh complex values, rounded up to the 16 bytes the engine works in
    // (FP32: the extra 8 bytes are the zero padding of the row, Tld is a multiple of 16 elements)
    const unsigned ser_bytes = ((unsigned)nh * (unsigned)sizeof(C) + 15u) & ~15u;
    if (PREF && tid == 0 && bid < A.natoms) Ctx::bulk_load(pre, A.series + (size_t)bid * A.DS * A.Tld, ser_bytes, mbar);
    const int j2 = tid & 15;                         // NT is a multiple of 16: the same for every owned butterfly
    cd* part = reinterpret_cast<cd*>(A.partial + (size_t)bid * A.Tld);
    const C* inv2 = reinterpret_cast<const C*>(A.inv);

    for (int atom = bid; atom < A.natoms; atom += nblk) {
        const RT* ser = A.series + (size_t)atom * A.DS * A.Tld;
        cd* row = reinterpret_cast<cd*>(A.by_particle + (size_t)atom * A.Tld);
        for (int r = 0; r < 2; ++r) {
            RT acc_s[NB][8], acc_d[NB][8], acc8[NB];
#pragma unroll
            for (int it = 0; it < NB; ++it) {
                acc8[it] = (RT)0;
#pragma unroll
                for (int m = 0; m < 8; ++m) { acc_s[it][m] = (RT)0; acc_d[it][m] = (RT)0; }
            }

            for (int d = 0; d < A.D; ++d) {
                // ---------------- P1: series -> registers -> shared
                const C* src = reinterpret_cast<const C*>(ser + (size_t)d * A.Tld);
                for (int j = tid; j < 256; j += NT) {
                    C x[R1];
                    if (PREF) {
                        if (j == tid) Ctx::mbar_wait(mbar, pre_phase);
#pragma unroll
                        for (int q = 0; q < R1; ++q) {
                            const int n = j + 256 * q;
                            x[q] = (n < nh) ? pre[n] : czero;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < R1; ++q) {
                            const int n = j + 256 * q;
                            x[q] = (n < nh) ? Ctx::ld_stream(src + n) : czero;
                        }
                    }
                    if (r) {
                        static_for<1, R1>([&](auto iq) {
                            constexpr int q = decltype(iq)::value;
                            x[q] = mul_tw<q, 2 * R1, -1>(x[q]);
                        });
                    }
                    Dft<R1, -1>::run(x);
                    C e[4], g[NG];
                    k1f_p1_twiddles<R1>(s_om[j], r, e, g);
                    C* dst = buf + j + (j >> 4);
#pragma unroll
                    for (int k = 0; k < R1; ++k) {
                        C y = x[k];
                        if (k >= 4) y = cmul(y, g[k >> 2]);
                        if (r || (k & 3)) y = cmul(y, e[k & 3]);
                        x[k] = y;
                    }
#pragma unroll
                    for (int k = 0; k < R1; ++k) dst[272 * k] = x[k];
                }
                Ctx::sync();
                if (PREF) {
                    // every P1 thread has read the prefetch buffer: hand it to the bulk-copy engine for the next chain
                    pre_phase ^= 1u;
                    if (tid == NT - 1) {
                        const RT* nxt = nullptr;
                        if (d + 1 < A.D) nxt = ser + (size_t)(d + 1) * A.Tld;
                        else if (r == 0) nxt = ser;
                        else if (atom + nblk < A.natoms) nxt = ser + (size_t)nblk * A.DS * A.Tld;
                        if (nxt) Ctx::bulk_load(pre, nxt, ser_bytes, mbar);
                    }
                }
                // ---------------- P2: radix 16, stride 16
#pragma unroll
                for (int it = 0; it < NB; ++it) {
                    const int vt = tid + it * NT;
                    if (NV % NT != 0 && vt >= NV) break;
                    const int blk2 = (int)((A.map[r * NV + vt] >> 16) & 0xffu);
                    const int p2base = blk2 * 272 + j2;          // padded address of (blk*256 + j), + 17 q
                    C x[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) x[q] = buf[p2base + 17 * q];
                    Dft<16, -1>::run(x);
#pragma unroll
                    for (int k = 1; k < 16; ++k) x[k] = cmul(x[k], s_tw2[(k - 1) * 16 + j2]);
#pragma unroll
                    for (int k = 0; k < 16; ++k) buf[p2base + 17 * k] = x[k];
                }
                Ctx::sync_warp();        // P3 of this warp reads what this warp's P2 wrote
                // ---------------- P3: radix 16, stride 1, + pair accumulation
                static_for<0, NB>([&](auto iit) {
                    constexpr int it = decltype(iit)::value;
                    const int vt = tid + it * NT;
                    if (NV % NT != 0 && vt >= NV) return;
                    const uint32_t mp = A.map[r * NV + vt];
                    const C wb = A.wbase[r * NV + vt];
                    const int p3base = (int)(mp & 0xffffu) * 17;
                    const bool self0 = (mp & K1F_SELF0) != 0, self8 = (mp & K1F_SELF8) != 0;
                    C v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = buf[p3base + q];
                    Dft<16, -1>::run(v);
                    static_for<0, 8>([&](auto im) {
                        constexpr int m = decltype(im)::value;
                        C snd = v[15 - m], rec;
                        rec.x = Ctx::shfl_xor16(snd.x);
                        rec.y = Ctx::shfl_xor16(snd.y);
                        if (self8) rec = snd;
                        if (self0) rec = v[(16 - m) & 15];
                        const C w = mul_tw<m, 32, -1>(wb);
                        const C U = v[m];
                        const RT nu = cnorm2(U), nv = cnorm2(rec);
                        const RT B = U.x * rec.y + U.y * rec.x;
                        acc_s[it][m] += nu + nv;
                        acc_d[it][m] += (RT)2 * w.x * B + w.y * (nu - nv);
                    });
                    if (self0) acc8[it] += (RT)2 * cnorm2(v[8]);
                });
                if (d + 1 < A.D) Ctx::sync();   // P1 of the next series overwrites the buffer
            }

            // ---------------- inverse: build from the accumulators, P3'
            static_for<0, NB>([&](auto iit) {
                constexpr int it = decltype(iit)::value;
                const int vt = tid + it * NT;
                if (NV % NT != 0 && vt >= NV) return;
                const uint32_t mp = A.map[r * NV + vt];
                const C wb = A.wbase[r * NV + vt];
                const int p3base = (int)(mp & 0xffffu) * 17;
                const bool self0 = (mp & K1F_SELF0) != 0, self8 = (mp & K1F_SELF8) != 0;
                C v[16], ap[8], rc[8];
                static_for<0, 8>([&](auto im) {
                    constexpr int m = decltype(im)::value;
                    const C w = mul_tw<m, 32, -1>(wb);
                    const RT sig = acc_s[it][m], del = acc_d[it][m];
                    v[m] = cmake<RT>(sig + w.y * del, w.x * del);
                    ap[m] = cmake<RT>(sig - w.y * del, w.x * del);
                    rc[m].x = Ctx::shfl_xor16(ap[m].x);
                    rc[m].y = Ctx::shfl_xor16(ap[m].y);
                });
                static_for<8, 16>([&](auto ii) {
                    constexpr int idx = decltype(ii)::value;
                    C val = rc[15 - idx];
                    if (self8) val = ap[15 - idx];
                    if (self0) val = (idx == 8) ? cmake<RT>(acc8[it], (RT)0) : ap[16 - idx];
                    v[idx] = val;
                });
                Dft<16, +1>::run(v);
#pragma unroll
                for (int q = 0; q < 16; ++q) buf[p3base + q] = v[q];
            });
            Ctx::sync_warp();
            // ---------------- P2'
#pragma unroll
            for (int it = 0; it < NB; ++it) {
                const int vt = tid + it * NT;
                if (NV % NT != 0 && vt >= NV) break;
                const int blk2 = (int)((A.map[r * NV + vt] >> 16) & 0xffu);
                const int p2base = blk2 * 272 + j2;
                C x[16];
                x[0] = buf[p2base];
#pragma unroll
                for (int k = 1; k < 16; ++k) x[k] = cmulc(buf[p2base + 17 * k], s_tw2[(k - 1) * 16 + j2]);
                Dft<16, +1>::run(x);
#pragma unroll
                for (int q = 0; q < 16; ++q) buf[p2base + 17 * q] = x[q];
            }
            Ctx::sync();
            // ---------------- P1' + output
            for (int j = tid; j < 256; j += NT) {
                C e[4], g[NG];
                k1f_p1_twiddles<R1>(s_om[j], r, e, g);
                const C* srcb = buf + j + (j >> 4);
                C x[R1];
#pragma unroll
                for (int k = 0; k < R1; ++k) x[k] = srcb[272 * k];
#pragma unroll
                for (int k = 0; k < R1; ++k) {
                    C y = x[k];
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
                if (TMEM) {
                    if (r == 0) {
                        // parked in this thread's tensor-memory columns until residue 1 is through
                        static_for<0, R1 / 4>([&](auto ic) {
                            constexpr int c = decltype(ic)::value;
                            cd v[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = cmake<double>((double)x[4 * c + i].x, (double)x[4 * c + i].y);
                            Ctx::tmem_st4(tm + 16 * c, v);
                        });
                        Ctx::tmem_wait_st();
                    } else {
                        static_for<0, R1 / 4>([&](auto ic) {
                            constexpr int c = decltype(ic)::value;
                            // one 16-column load at a time (each returns in ~12 clocks): the chunk's four values of x[] are
                            // finished in place, so no more than 16 extra registers are live beside x[]
                            cd t[4];
                            Ctx::tmem_ld4(tm + 16 * c, t);                       // parked V_0
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                x[4 * c + i] = cmake<RT>((RT)t[i].x + x[4 * c + i].x, (RT)t[i].y + x[4 * c + i].y);
                            Ctx::tmem_ld4(tm + TM_INV + 16 * c, t);              // 1 / (L (T - k)), 0 beyond the series
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int n = j + 256 * (4 * c + i);
                                x[4 * c + i] = cmake<RT>(x[4 * c + i].x * (RT)t[i].x, x[4 * c + i].y * (RT)t[i].y);
                                if (n < nh) row[n] = cmake<double>((double)x[4 * c + i].x, (double)x[4 * c + i].y);
                            }
                            if (PART) {
                                Ctx::tmem_ld4(tm + TM_PART + 16 * c, t);         // particle sums
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int n = j + 256 * (4 * c + i);
                                    if (n < nh) t[i] = cmake<double>(t[i].x + (double)x[4 * c + i].x, t[i].y + (double)x[4 * c + i].y);
                                }
                                Ctx::tmem_st4(tm + TM_PART + 16 * c, t);
                            }
                        });
                        Ctx::tmem_wait_st();
                    }
                } else if (r == 0) {
#pragma unroll
                    for (int q = 0; q < R1; ++q) {
                        const int n = j + 256 * q;
                        if (n < nh) row[n] = cmake<double>((double)x[q].x, (double)x[q].y);   // parked raw; finished by residue 1
                    }
                } else {
                    // Park the finished V_1 values in this thread's own buffer slots, then stream the
                    // output in chunks with all global loads of a chunk issued first: with x[] out of
                    // the registers there is room to keep a whole chunk of L2 round trips in flight
                    // (the row / part stores may alias the loads as far as the compiler knows).
                    C* own = buf + j + (j >> 4);
#pragma unroll
                    for (int q = 0; q < R1; ++q) own[272 * q] = x[q];
                    Ctx::compiler_fence();
                    constexpr int QB = (R1 % 5 == 0) ? 5 : 4;
                    static_for<0, (R1 + QB - 1) / QB>([&](auto ic) {
                        constexpr int q0 = decltype(ic)::value * QB;
                        constexpr int nq = (R1 - q0) < QB ? (R1 - q0) : QB;
                        cd a[nq], ps[nq];
                        C sc[nq];
#pragma unroll
                        for (int i = 0; i < nq; ++i) {
                            const int n = j + 256 * (q0 + i);
                            const int nc = n < nh ? n : nh - 1;      // clamped: the loads stay branch-free
                            a[i] = Ctx::ld_stream(row + nc);
                            if (PART) ps[i] = Ctx::ld_stream(part + nc);
                            sc[i] = inv2[nc];
                        }
#pragma unroll
                        for (int i = 0; i < nq; ++i) {
                            const int n = j + 256 * (q0 + i);
                            if (n < nh) {
                                const C v1 = own[272 * (q0 + i)];
                                const cd o = cmake<double>((double)(((RT)a[i].x + v1.x) * sc[i].x),
                                                           (double)(((RT)a[i].y + v1.y) * sc[i].y));
                                row[n] = o;
                                if (PART) part[n] = cmake<double>(ps[i].x + o.x, ps[i].y + o.y);
                            }
                        }
                    });
                }
            }
            // no barrier here: P1 of the next chain writes exactly the elements this
            // thread has just read in P1'
        }
    }
    if (TMEM) {
        // add the particle sums of this launch to the CTA's global partial row (the host zeroed it before the first launch
        // of the compute call; a call that follows ta_stage_bulk launches once per staging chunk), then give the tensor
        // memory back
        if (PART && tid < 256) {
            static_for<0, R1 / 4>([&](auto ic) {
                constexpr int c = decltype(ic)::value;
                cd ps[4];
                Ctx::tmem_ld4(tm + TM_PART + 16 * c, ps);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = tid + 256 * (4 * c + i);
                    if (n < nh) { const cd g0 = part[n]; part[n] = cmake<double>(g0.x + ps[i].x, g0.y + ps[i].y); }
                }
            });
        }
        Ctx::tmem_free(tm_base, tid);
    }
}

// ---------------------------------------------------------------------------
// Host-side plan for the fast path.
// ---------------------------------------------------------------------------
struct K1FastPlan {
    int R1 = 0, H = 0, L = 0, NT = 0, nh = 0;
    std::vector<double> omega;    // 256 x (re, im)
    std::vector<double> tw2;      // 240 x (re, im)
    std::vector<uint32_t> map;    // 2 x NT
    std::vector<double> wbase;    // 2 x NT x (re, im)
    std::vector<double> inv;      // Tld
};

// radices R1 the library instantiates (even: the lane exchange needs whole warps)
inline const int* k1f_supported_r1(int* n) {
    static const int r1s[] = {4, 6, 8, 10, 12, 16, 20, 24};
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

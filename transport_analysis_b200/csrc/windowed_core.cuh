// Per-lane core of the windowed-lag kernel K2/K3.
//
// K2 replaces the lag loop of VelocityAutocorr._conclude_simple
// (transport_analysis/velocityautocorr.py:223-235): sum_i g[i] * g[i+k].
// K3 replaces the lag loop of ViscosityHelfand._conclude
// (transport_analysis/viscosity.py:210-226): sum_i (g[i] - g[i+k])^2 with
// g = (m*v)*x formed at staging time.  Both are direct O(T^2) accumulations
// (no S1 - 2 S2 rewrite: it cancels catastrophically on ramp-like data).
//
// Work decomposition: lags are grouped in blocks of 16 (k0 = 16 kb); the
// origins i of a block are cut into chunks of 16; one lane owns one 16 x 16
// (origin x lag) tile at a time: 16 + 31 series values from shared memory feed
// 256 FMAs.  A warp works on a PAIR of lag blocks (kb, nlb-1-kb); its lanes are
// split between the two blocks in proportion to their chunk counts and the
// 16 partial sums of each block are reduced with halving shuffle exchanges.
//
// Shared-memory layout of one series: element x lives at x + 2 * (x >> 4)
// (two pad doubles per 16), so the 128-bit loads of the 32 lanes -- whose
// tiles start 16 elements apart -- fall 9 * 16 bytes apart: conflict free.
// The series is followed by >= 48 zeros.
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int TA_WIN_LAGS = 16;    // lags per block
constexpr int TA_WIN_CHUNK = 16;   // origins per tile

TA_HD int win_addr(int x) { return x + 2 * (x >> 4); }
// doubles needed for a series of length T in the padded layout (incl. zero tail)
TA_HD int win_smem_elems(int T) { return win_addr(((T + 15) / 16) * 16 + 48) ; }
TA_HD int win_num_lag_blocks(int T) { return (T + TA_WIN_LAGS - 1) / TA_WIN_LAGS; }

template <typename R, int MODE, bool MASK>
TA_HD void win_tile(const R* S, int i0, int k0, int T, R* acc) {
    const cplx<R>* A2 = reinterpret_cast<const cplx<R>*>(S + win_addr(i0));
    const cplx<R>* B2a = reinterpret_cast<const cplx<R>*>(S + win_addr(i0 + k0));
    const cplx<R>* B2b = reinterpret_cast<const cplx<R>*>(S + win_addr(i0 + k0 + 16));
    R a[16], b[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        cplx<R> va = A2[q], vb = B2a[q], vc = B2b[q];
        a[2 * q] = va.x; a[2 * q + 1] = va.y;
        b[2 * q] = vb.x; b[2 * q + 1] = vb.y;
        b[16 + 2 * q] = vc.x; b[16 + 2 * q + 1] = vc.y;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (MODE == TA_WIN_PRODUCT) {
                acc[m] += a[u] * b[u + m];
            } else {
                R d = a[u] - b[u + m];
                if (MASK && (i0 + u + k0 + m >= T)) d = (R)0;
                acc[m] += d * d;
            }
        }
    }
}

// Chunks of lag block kb whose 16 x 16 tile needs no validity mask.  PRODUCT: all of them (the zero
// tail makes out-of-range products vanish).  SQDIFF: those with i0 + 15 + k0 + 15 < T; the remaining
// one to three "tail" chunks per block are handled by win_tail_block.
template <int MODE>
TA_HD int win_num_chunks(int T, int kb) { return (T - kb * TA_WIN_LAGS + TA_WIN_CHUNK - 1) / TA_WIN_CHUNK; }
template <int MODE>
TA_HD int win_full_chunks(int T, int kb) {
    if (MODE == TA_WIN_PRODUCT) return win_num_chunks<MODE>(T, kb);
    const int n = T - kb * TA_WIN_LAGS - 30;
    return n > 0 ? (n + TA_WIN_CHUNK - 1) / TA_WIN_CHUNK : 0;
}

// Lag-block schedule of one warp: blocks are paired (kb, nlb-1-kb) so that every pair carries about
// T + 16 origins; warp w takes pairs w, w + nwarps, ...  The 32 lanes are split between the two
// blocks of a pair in proportion to their chunk counts (lanes [0, La) -> first block), so that a
// pair costs ceil-ish((nfa + nfb) / 32) tile rounds instead of ceil(nfa / 32) + ceil(nfb / 32).
TA_HD int win_num_pairs(int nlb) { return (nlb + 1) / 2; }
TA_HD void win_pair_blocks(int pair, int nlb, int* kb_a, int* kb_b) {
    *kb_a = pair;
    int other = nlb - 1 - pair;
    *kb_b = (other != pair) ? other : -1;
}
TA_HD int win_rounds(int nfa, int nfb, int La) {
    const int ra = nfa > 0 ? (nfa + La - 1) / La : 0;
    const int rb = nfb > 0 ? (nfb + (32 - La) - 1) / (32 - La) : 0;
    return ra > rb ? ra : rb;
}
TA_HD int win_split(int nfa, int nfb) {
    if (nfb <= 0) return 32;
    if (nfa <= 0) return 0;
    int La = (32 * nfa + (nfa + nfb) / 2) / (nfa + nfb);
    La = La < 1 ? 1 : (La > 31 ? 31 : La);
    int best = La, br = win_rounds(nfa, nfb, La);
    for (int c = La - 1; c <= La + 1; c += 2)
        if (c >= 1 && c <= 31 && win_rounds(nfa, nfb, c) < br) { best = c; br = win_rounds(nfa, nfb, c); }
    return best;
}

// Sum over the lanes with `take` set of their 16 partial sums, by halving exchanges (16 shuffles of
// a double instead of 16 x 5): on return lane l holds the total of lag  m = l >> 1.
template <class Ctx, typename R>
TA_HD double win_reduce16(const R* acc, bool take, int lane) {
    double v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = take ? (double)acc[m] : 0.0;
    {
        const bool hi = (lane & 16) != 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const double send = hi ? v[m] : v[m + 8], keep = hi ? v[m + 8] : v[m];
            v[m] = keep + Ctx::shfl_xor(send, 16);
        }
    }
    {
        const bool hi = (lane & 8) != 0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const double send = hi ? v[m] : v[m + 4], keep = hi ? v[m + 4] : v[m];
            v[m] = keep + Ctx::shfl_xor(send, 8);
        }
    }
    {
        const bool hi = (lane & 4) != 0;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const double send = hi ? v[m] : v[m + 2], keep = hi ? v[m + 2] : v[m];
            v[m] = keep + Ctx::shfl_xor(send, 4);
        }
    }
    {
        const bool hi = (lane & 2) != 0;
        const double send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
        v[0] = keep + Ctx::shfl_xor(send, 2);
    }
    return v[0] + Ctx::shfl_xor(v[0], 1);
}

struct WinArgs {
    const void* series;     // [natoms][DS][Tld] of the arithmetic type R (double, or float in the FP32 mode); rows 0 .. D-1 are used
    double* by_particle;    // [natoms][Tld]
    double* partial;        // [nblk][Tld]
    int natoms, D, DS, T;
    long long Tld;
    double denom;           // Helfand: 2 kB <V> temp_avg ; VACF: unused
    // Series too long for shared memory: per-CTA global scratch (same layout, read through L1/L2) -- slower,
    // but the direct lag sums then work for any T.  Null: the series and the lag sums live in shared memory.
    unsigned char* scratch;
    long long scratch_stride;   // bytes per CTA
    // Few particles (fewer than ~4 per resident CTA): a particle's lag-block pairs are dealt to `nsplit` CTAs, each of which
    // stages the whole series but computes, finishes and stores only the lags of its own pairs -- the work units
    // (particle, part) are then small enough to fill the last wave of the grid.  1: a CTA does a whole particle.
    int nsplit;
};

// ---------------------------------------------------------------------------
// The kernel body for one CTA (persistent over work units bid, bid + nblk, ...; unit u = part u % nsplit of particle
// u / nsplit).  Ctx supplies
// sync() and shfl_xor(); tests/emu runs the same code on the CPU under cooperative fibers.
//   MODE = TA_WIN_PRODUCT: vacf[k] = sum_d sum_i g_d[i] g_d[i+k] / (T-k)
//   MODE = TA_WIN_SQDIFF : visc[k] = sum_d sum_i (g_d[i]-g_d[i+k])^2 / (D (T-k)) / denom
// ---------------------------------------------------------------------------
template <typename R, int MODE, class Ctx, bool SCRATCH = false>
TA_HD void win_body(const WinArgs& A, unsigned char* smem_raw, int tid, int nthr, int bid, int nblk) {
    const int T = A.T;
    const int ne = win_smem_elems(T);
    if (SCRATCH) smem_raw = A.scratch + (size_t)bid * (size_t)A.scratch_stride;   // compile-time: the shared-memory build keeps LDS/STS
    R* S = reinterpret_cast<R*>(smem_raw);
    double* res = reinterpret_cast<double*>(S + ((ne + 1) & ~1));
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int nlb = win_num_lag_blocks(T), npairs = win_num_pairs(nlb);
    double* partial = A.partial + (size_t)bid * A.Tld;

    const int nsplit = A.nsplit > 1 ? A.nsplit : 1;
    // lag block blk belongs to pair min(blk, nlb - 1 - blk); pair p is computed by part (p / nwarps) % nsplit
    auto part_of = [&](int blk) { const int p = blk < nlb - 1 - blk ? blk : nlb - 1 - blk; return (p / nwarps) % nsplit; };
    for (long long u = bid; u < (long long)A.natoms * nsplit; u += nblk) {
        const int a = (int)(u / nsplit), part = (int)(u % nsplit);
        for (int k = tid; k < T; k += nthr) res[k] = 0.0;
        for (int d = 0; d < A.D; ++d) {
            const R* ser = reinterpret_cast<const R*>(A.series) + ((size_t)a * A.DS + d) * A.Tld;
            Ctx::sync();   // previous series fully consumed
            for (int x = tid; x < ne; x += nthr) S[x] = (R)0;
            Ctx::sync();
            for (int x = tid; x < T; x += nthr) S[win_addr(x)] = ser[x];
            Ctx::sync();
            // ---- unmasked tiles: a warp per block pair, lanes split between the two blocks
            for (int pair = part * nwarps + warp; pair < npairs; pair += nwarps * nsplit) {
                int ka, kb;
                win_pair_blocks(pair, nlb, &ka, &kb);
                const int nfa = win_full_chunks<MODE>(T, ka);
                const int nfb = kb >= 0 ? win_full_chunks<MODE>(T, kb) : 0;
                if (nfa + nfb == 0) continue;
                const int La = win_split(nfa, nfb);
                const bool mine_a = lane < La;
                const int k0 = (mine_a ? ka : kb) * TA_WIN_LAGS;
                const int cnt = mine_a ? nfa : nfb;
                const int stride = mine_a ? La : 32 - La;
                R acc[TA_WIN_LAGS];
#pragma unroll
                for (int m = 0; m < TA_WIN_LAGS; ++m) acc[m] = (R)0;
                for (int c = mine_a ? lane : lane - La; c < cnt; c += stride)
                    win_tile<R, MODE, false>(S, c * TA_WIN_CHUNK, k0, T, acc);
                const double ra = win_reduce16<Ctx, R>(acc, mine_a, lane);
                const double rb = (nfb > 0) ? win_reduce16<Ctx, R>(acc, !mine_a, lane) : 0.0;
                if ((lane & 1) == 0) {      // this warp is the only one that touches the pair's lags in this phase
                    const int m = lane >> 1;
                    if (nfa > 0 && ka * TA_WIN_LAGS + m < T) res[ka * TA_WIN_LAGS + m] += ra;
                    if (nfb > 0 && kb * TA_WIN_LAGS + m < T) res[kb * TA_WIN_LAGS + m] += rb;
                }
            }
            if (MODE == TA_WIN_SQDIFF) {
                // ---- masked tail tiles: one lane per lag block (no cross-lane reduction, fixed order)
                Ctx::sync();
                for (int blk = tid; blk < nlb; blk += nthr) {
                    if (nsplit > 1 && part_of(blk) != part) continue;
                    const int k0 = blk * TA_WIN_LAGS;
                    R acc[TA_WIN_LAGS];
#pragma unroll
                    for (int m = 0; m < TA_WIN_LAGS; ++m) acc[m] = (R)0;
                    const int nch = win_num_chunks<MODE>(T, blk);
                    for (int c = win_full_chunks<MODE>(T, blk); c < nch; ++c)
                        win_tile<R, MODE, true>(S, c * TA_WIN_CHUNK, k0, T, acc);
#pragma unroll
                    for (int m = 0; m < TA_WIN_LAGS; ++m)
                        if (k0 + m < T) res[k0 + m] += (double)acc[m];
                }
            }
        }
        Ctx::sync();
        double* row = A.by_particle + (size_t)a * A.Tld;
        for (int k = tid; k < T; k += nthr) {
            if (nsplit > 1 && part_of(k / TA_WIN_LAGS) != part) continue;      // another CTA finishes this lag
            double val;
            if (MODE == TA_WIN_PRODUCT) val = res[k] / (double)(T - k);
            else val = res[k] / ((double)A.D * (double)(T - k)) / A.denom;
            row[k] = val;
            partial[k] += val;
        }
        Ctx::sync();
    }
}

}  // namespace ta

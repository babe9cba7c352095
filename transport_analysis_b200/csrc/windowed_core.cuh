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
// 256 FMAs.  The 32 lanes of a warp take chunks c = lane, lane + 32, ... of
// the same lag block and their 16 partial sums are tree-reduced with shuffles.
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

// Partial sums of one lane for lag block kb: acc[m] += sum over the lane's
// tiles of f(g[i], g[i + 16 kb + m]).
template <typename R, int MODE>
TA_HD void win_lane_accumulate(int lane, int nlanes, const R* S, int T, int kb, R* acc) {
    const int k0 = kb * TA_WIN_LAGS;
    const int ni = T - k0;                       // valid origins for lag k0
    const int nch = (ni + TA_WIN_CHUNK - 1) / TA_WIN_CHUNK;
    for (int c = lane; c < nch; c += nlanes) {
        const int i0 = c * TA_WIN_CHUNK;
        if (MODE == TA_WIN_PRODUCT) {
            win_tile<R, MODE, false>(S, i0, k0, T, acc);   // zero tail makes masking unnecessary
        } else {
            if (i0 + 15 + k0 + 15 >= T) win_tile<R, MODE, true>(S, i0, k0, T, acc);
            else win_tile<R, MODE, false>(S, i0, k0, T, acc);
        }
    }
}

// Lag-block schedule of one warp: blocks are paired (kb, nlb-1-kb) so that
// every pair carries about T + 16 origins; warp w takes pairs w, w + nwarps...
TA_HD int win_num_pairs(int nlb) { return (nlb + 1) / 2; }
TA_HD void win_pair_blocks(int pair, int nlb, int* kb_a, int* kb_b) {
    *kb_a = pair;
    int other = nlb - 1 - pair;
    *kb_b = (other != pair) ? other : -1;
}

}  // namespace ta

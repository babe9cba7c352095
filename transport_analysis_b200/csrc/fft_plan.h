// Host-side plan for the FFT autocorrelation kernel (K1).
//
// What K1 computes (replaces tidynamics.acf as called from
// transport_analysis/velocityautocorr.py:210-214): for one particle and D
// real series x_d[0..T), acf[k] = sum_d sum_i x_d[i] x_d[i+k] / (T-k).
//
// How (see DESIGN.md "K1"): the length-L circular correlation, L = 4H >= 2T,
// is never formed as one transform.  A real series is packed into the complex
// series z[n] = x[2n] + i x[2n+1] (n < H, zero beyond T/2), whose length-2H
// spectrum splits into its even / odd bins Z[2g+r] = FFT_H(z[n] w_{2H}^{rn})[g]
// because the upper half of the padded input is zero.  Each residue r in {0,1}
// is an independent chain: twist -> in-place DIF FFT_H (digit-scrambled output)
// -> pairwise power-spectrum accumulation over the D series -> in-place DIT
// inverse FFT_H (scrambled in, natural out) -> twisted add into the output.
// All of it lives in one H-point complex shared-memory buffer.
//
// This header only chooses H, the radix schedule and builds the small tables.
#pragma once
#include <cstdint>
#include <vector>
#include <cmath>
#include "ta_common.cuh"

namespace ta {

struct FftPlanHost {
    int64_t T = 0;        // series length
    int H = 0;            // complex FFT length, H >= ceil(T/2)
    int L = 0;            // 4H: length of the equivalent zero-padded real transform
    int npasses = 0;
    int radix[TA_MAX_PASSES] = {0};
    int lo_bits = 0;      // two-level twiddle table: w_L^j = hi[j >> lo_bits] * lo[j & mask]
    std::vector<double> tw_lo;   // interleaved re,im ; 1<<lo_bits entries
    std::vector<double> tw_hi;   // interleaved re,im ; ceil(L / (1<<lo_bits)) entries
    std::vector<uint32_t> ftab;  // scrambled position p -> frequency index g
    std::vector<uint32_t> pair0; // residue 0: position of the partner bin (H - g) mod H
};

inline bool ta_smooth235(int64_t n, int* twos, int* threes, int* fives) {
    *twos = *threes = *fives = 0;
    while (n % 2 == 0) { n /= 2; ++*twos; }
    while (n % 3 == 0) { n /= 3; ++*threes; }
    while (n % 5 == 0) { n /= 5; ++*fives; }
    return n == 1;
}

// Smallest 2^a 3^b 5^c >= n (b <= 1, c <= 2 keeps the odd passes few).
inline int ta_choose_fft_len(int64_t n) {
    if (n < 1) n = 1;
    for (int64_t h = n;; ++h) {
        int a, b, c;
        if (ta_smooth235(h, &a, &b, &c) && b <= 1 && c <= 2) return (int)h;
    }
}

// exact-ish w_L^j = exp(-2 pi i j / L) using octant reduction in long double.
inline void ta_twiddle(int64_t j, int64_t L, double* re, double* im) {
    j %= L;
    if (j < 0) j += L;
    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)L;
    *re = (double)cosl(a);
    *im = (double)sinl(a);
    // exact values on the axes
    if ((4 * j) % L == 0) {
        int q = (int)((4 * j) / L);
        const double cr[4] = {1, 0, -1, 0}, ci[4] = {0, -1, 0, 1};
        *re = cr[q]; *im = ci[q];
    }
}

inline int ta_build_fft_plan(int64_t T, FftPlanHost* p) {
    if (T < 1) return TA_ERR_INVALID;
    p->T = T;
    int64_t half = (T + 1) / 2;
    int H = ta_choose_fft_len(half < 2 ? 2 : half);
    p->H = H;
    p->L = 4 * H;
    int a, b, c;
    ta_smooth235(H, &a, &b, &c);
    // Radix schedule, DIF order (first pass has the largest stride).  Powers of
    // two first (8s, then a 4 or 2), odd radices last: a stride-1 pass of odd
    // radix is free of shared-memory bank conflicts, a stride-1 radix-8 is not.
    int np = 0;
    while (a >= 3) { p->radix[np++] = 8; a -= 3; }
    if (a == 2) p->radix[np++] = 4;
    if (a == 1) p->radix[np++] = 2;
    for (int i = 0; i < b; ++i) p->radix[np++] = 3;
    for (int i = 0; i < c; ++i) p->radix[np++] = 5;
    if (np > TA_MAX_PASSES) return TA_ERR_UNSUPPORTED;
    p->npasses = np;

    // two-level twiddle table over w_L^j, j in [0, L)
    int lo_bits = 0;
    while ((1 << (2 * lo_bits)) < p->L) ++lo_bits;   // ~sqrt(L)
    if (lo_bits < 1) lo_bits = 1;
    p->lo_bits = lo_bits;
    int nlo = 1 << lo_bits;
    int nhi = (p->L + nlo - 1) / nlo;
    p->tw_lo.assign(2 * (size_t)nlo, 0.0);
    p->tw_hi.assign(2 * (size_t)nhi, 0.0);
    for (int j = 0; j < nlo; ++j) ta_twiddle(j, p->L, &p->tw_lo[2 * j], &p->tw_lo[2 * j + 1]);
    for (int j = 0; j < nhi; ++j) ta_twiddle((int64_t)j * nlo, p->L, &p->tw_hi[2 * j], &p->tw_hi[2 * j + 1]);

    // scramble tables: position p = sum_i k_i s_i  <->  g = k_1 + r_1 k_2 + r_1 r_2 k_3 + ...
    p->ftab.assign(H, 0);
    p->pair0.assign(H, 0);
    std::vector<uint32_t> postab(H, 0);
    for (int pos = 0; pos < H; ++pos) {
        int rem = pos, g = 0, R = 1, size = H;
        for (int i = 0; i < np; ++i) {
            int s = size / p->radix[i];
            int k = rem / s;
            rem -= k * s;
            g += k * R;
            R *= p->radix[i];
            size = s;
        }
        p->ftab[pos] = (uint32_t)g;
        postab[g] = (uint32_t)pos;
    }
    for (int pos = 0; pos < H; ++pos) {
        int g = (int)p->ftab[pos];
        p->pair0[pos] = postab[(H - g) % H];
    }
    return TA_OK;
}

}  // namespace ta

// Per-CTA phases of the FFT autocorrelation kernel K1 (see fft_plan.h for the
// algorithm).  Every phase is a function of (tid, nthreads) that touches the
// shared buffer only at indices it owns for that phase; the kernel separates
// phases with __syncthreads(), the CPU emulation harness (tests/emu) runs each
// phase for tid = 0..nthreads-1 in turn.  Replaces the arithmetic of
// tidynamics.acf at transport_analysis/velocityautocorr.py:211-213.
#pragma once
#include "ta_common.cuh"

namespace ta {

// Device-side view of the plan (pointers are shared-memory or global copies).
template <typename R>
struct FftTables {
    int T, H, L;
    int npasses;
    int radix[TA_MAX_PASSES];
    int lo_bits;
    const cplx<R>* tw_lo;      // 1 << lo_bits entries
    const cplx<R>* tw_hi;      // ceil(L >> lo_bits) entries
    const uint32_t* ftab;      // position -> frequency index g
    const uint32_t* pair0;     // residue 0 partner position
    const uint32_t* own0;      // residue 0: owner positions (p <= pair0[p]), npairs0 entries
    int npairs0;
};

template <typename R>
TA_HD cplx<R> tw_get(const FftTables<R>& t, int idx) {
    cplx<R> hi = t.tw_hi[idx >> t.lo_bits];
    cplx<R> lo = t.tw_lo[idx & ((1 << t.lo_bits) - 1)];
    return cmul(hi, lo);
}

// ---------------------------------------------------------------------------
// Radix butterflies in registers.  DIR = -1: forward kernel exp(-2 pi i qk/r),
// DIR = +1: inverse (conjugate) kernel.  Unnormalised.
// ---------------------------------------------------------------------------
template <typename R, int DIR> TA_HD cplx<R> rot90(cplx<R> a) {  // a * (DIR * i)
    return DIR < 0 ? cmul_mi(a) : cmul_pi(a);
}

template <typename R, int DIR> TA_HD void bfly2(cplx<R>& a, cplx<R>& b) {
    cplx<R> t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <typename R, int DIR> TA_HD void bfly4(cplx<R>& v0, cplx<R>& v1, cplx<R>& v2, cplx<R>& v3) {
    cplx<R> t0 = cadd(v0, v2), t1 = csub(v0, v2);
    cplx<R> t2 = cadd(v1, v3), t3 = rot90<R, DIR>(csub(v1, v3));
    v0 = cadd(t0, t2);
    v1 = cadd(t1, t3);
    v2 = csub(t0, t2);
    v3 = csub(t1, t3);
}

template <typename R, int DIR> TA_HD void bfly3(cplx<R>* v) {
    const R hs3 = (R)0.86602540378443864676372317075294;  // sqrt(3)/2
    cplx<R> t = cadd(v[1], v[2]);
    cplx<R> u = csub(v[1], v[2]);
    cplx<R> m = cmake<R>(v[0].x - (R)0.5 * t.x, v[0].y - (R)0.5 * t.y);
    cplx<R> ru = rot90<R, DIR>(cmake<R>(hs3 * u.x, hs3 * u.y));
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, ru);
    v[2] = csub(m, ru);
}

template <typename R, int DIR> TA_HD void bfly5(cplx<R>* v) {
    const R c1 = (R)0.30901699437494742410229341718282;   // cos(2pi/5)
    const R c2 = (R)-0.80901699437494742410229341718282;  // cos(4pi/5)
    const R s1 = (R)0.95105651629515357211643933337938;   // sin(2pi/5)
    const R s2 = (R)0.58778525229247312916870595463907;   // sin(4pi/5)
    cplx<R> t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    cplx<R> t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    cplx<R> a1 = cmake<R>(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    cplx<R> a2 = cmake<R>(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    cplx<R> b1 = rot90<R, DIR>(cmake<R>(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
    cplx<R> b2 = rot90<R, DIR>(cmake<R>(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
    v[0] = cmake<R>(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
    v[1] = cadd(a1, b1);
    v[4] = csub(a1, b1);
    v[2] = cadd(a2, b2);
    v[3] = csub(a2, b2);
}

template <typename R, int DIR> TA_HD void bfly8(cplx<R>* v) {
    const R h = (R)0.70710678118654752440084436210485;  // 1/sqrt(2)
    // even / odd radix-4 sub-transforms
    cplx<R> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    cplx<R> o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    bfly4<R, DIR>(e0, e1, e2, e3);
    bfly4<R, DIR>(o0, o1, o2, o3);
    // o_k *= w8^k, w8 = exp(DIR * i pi/4)
    cplx<R> r1 = rot90<R, DIR>(o1);                        // o1 * (DIR i)
    o1 = cmake<R>(h * (o1.x + r1.x), h * (o1.y + r1.y));   // o1 * (1 + DIR i)/sqrt2
    cplx<R> r3 = rot90<R, DIR>(o3);
    o3 = cmake<R>(h * (r3.x - o3.x), h * (r3.y - o3.y));   // o3 * (-1 + DIR i)/sqrt2
    o2 = rot90<R, DIR>(o2);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <typename R, int RADIX, int DIR> TA_HD void bfly(cplx<R>* v) {
    if (RADIX == 2) bfly2<R, DIR>(v[0], v[1]);
    else if (RADIX == 3) bfly3<R, DIR>(v);
    else if (RADIX == 4) bfly4<R, DIR>(v[0], v[1], v[2], v[3]);
    else if (RADIX == 5) bfly5<R, DIR>(v);
    else if (RADIX == 8) bfly8<R, DIR>(v);
}

// w^1 .. w^(RADIX-1) from w by repeated multiplication (a few ulp; the
// tolerance budget is 1e-10 relative, see tests/test_gpu_parity.py).
template <typename R, int RADIX> TA_HD void tw_powers(cplx<R> w, cplx<R>* p) {
    p[1] = w;
    if (RADIX > 2) p[2] = cmul(w, w);
    if (RADIX > 3) p[3] = cmul(p[2], w);
    if (RADIX > 4) p[4] = cmul(p[2], p[2]);
    if (RADIX > 5) p[5] = cmul(p[4], w);
    if (RADIX > 6) p[6] = cmul(p[3], p[3]);
    if (RADIX > 7) p[7] = cmul(p[4], p[3]);
}

// ---------------------------------------------------------------------------
// Phase: one in-place decimation-in-frequency pass, sub-transform size r*s.
//   y_k[j] = (sum_q x[j + q s] w_r^{qk}) * w_{rs}^{jk}   stored at j + k s
// ---------------------------------------------------------------------------
template <typename R, int RADIX>
TA_HD void dif_pass(int tid, int nthr, cplx<R>* buf, const FftTables<R>& t, int s) {
    const int nb = t.H / RADIX;
    const int tstep = t.L / (RADIX * s);   // w_{rs}^j = w_L^{tstep * j}
    for (int b = tid; b < nb; b += nthr) {
        int blk = b / s;
        int j = b - blk * s;
        cplx<R>* base = buf + blk * (RADIX * s) + j;
        cplx<R> v[RADIX];
#pragma unroll
        for (int q = 0; q < RADIX; ++q) v[q] = base[q * s];
        bfly<R, RADIX, -1>(v);
        if (s > 1) {
            cplx<R> p[RADIX];
            tw_powers<R, RADIX>(tw_get(t, tstep * j), p);
#pragma unroll
            for (int k = 1; k < RADIX; ++k) v[k] = cmul(v[k], p[k]);
        }
#pragma unroll
        for (int k = 0; k < RADIX; ++k) base[k * s] = v[k];
    }
}

// Phase: one in-place decimation-in-time inverse pass (exact mirror of dif_pass).
template <typename R, int RADIX>
TA_HD void dit_pass(int tid, int nthr, cplx<R>* buf, const FftTables<R>& t, int s) {
    const int nb = t.H / RADIX;
    const int tstep = t.L / (RADIX * s);
    for (int b = tid; b < nb; b += nthr) {
        int blk = b / s;
        int j = b - blk * s;
        cplx<R>* base = buf + blk * (RADIX * s) + j;
        cplx<R> v[RADIX];
#pragma unroll
        for (int k = 0; k < RADIX; ++k) v[k] = base[k * s];
        if (s > 1) {
            cplx<R> p[RADIX];
            tw_powers<R, RADIX>(tw_get(t, tstep * j), p);
#pragma unroll
            for (int k = 1; k < RADIX; ++k) v[k] = cmulc(v[k], p[k]);
        }
        bfly<R, RADIX, +1>(v);
#pragma unroll
        for (int q = 0; q < RADIX; ++q) base[q * s] = v[q];
    }
}

template <typename R>
TA_HD void dif_pass_any(int radix, int tid, int nthr, cplx<R>* buf, const FftTables<R>& t, int s) {
    switch (radix) {
        case 2: dif_pass<R, 2>(tid, nthr, buf, t, s); break;
        case 3: dif_pass<R, 3>(tid, nthr, buf, t, s); break;
        case 4: dif_pass<R, 4>(tid, nthr, buf, t, s); break;
        case 5: dif_pass<R, 5>(tid, nthr, buf, t, s); break;
        default: dif_pass<R, 8>(tid, nthr, buf, t, s); break;
    }
}
template <typename R>
TA_HD void dit_pass_any(int radix, int tid, int nthr, cplx<R>* buf, const FftTables<R>& t, int s) {
    switch (radix) {
        case 2: dit_pass<R, 2>(tid, nthr, buf, t, s); break;
        case 3: dit_pass<R, 3>(tid, nthr, buf, t, s); break;
        case 4: dit_pass<R, 4>(tid, nthr, buf, t, s); break;
        case 5: dit_pass<R, 5>(tid, nthr, buf, t, s); break;
        default: dit_pass<R, 8>(tid, nthr, buf, t, s); break;
    }
}

// ---------------------------------------------------------------------------
// Phase: load one real series (length T, stored in the arithmetic type R, zero beyond T up to
// the next even index) as z[n] = x[2n] + i x[2n+1], n < H, twisted by
// w_{2H}^{n} = w_L^{2n} for the odd residue.
// ---------------------------------------------------------------------------
template <typename R>
TA_HD void fft_load(int tid, int nthr, cplx<R>* buf, const R* series, const FftTables<R>& t, int r) {
    const int nh = (t.T + 1) / 2;
    const cplx<R>* src = reinterpret_cast<const cplx<R>*>(series);
    for (int n = tid; n < t.H; n += nthr) {
        cplx<R> z = cmake<R>((R)0, (R)0);
        if (n < nh) {
            z = src[n];
            if (r) z = cmul(z, tw_get(t, 2 * n));
        }
        buf[n] = z;
    }
}

// Phase: zero the pair accumulators (H + 1 reals).
template <typename R>
TA_HD void fft_zero_acc(int tid, int nthr, R* sd, const FftTables<R>& t) {
    for (int i = tid; i <= t.H; i += nthr) sd[i] = (R)0;
}

// ---------------------------------------------------------------------------
// Phase: after the forward transform of one series, add its contribution to
// the pair accumulators.  For the bin pair (f, 2H - f), f = 2g + r, held at
// scrambled positions (p, p'):  U = buf[p], U' = buf[p'],
//   Sigma += |U|^2 + |U'|^2            (= P[f] + P[2H-f])
//   Delta += 2 Re(w) Im(U U') + Im(w) (|U|^2 - |U'|^2)   (= P[f] - P[2H-f]),  w = w_L^f
// Sigma is kept at sd[p], Delta at sd[p'] (the self pair g = 0 of residue 0
// keeps Delta at sd[H]; other self pairs have Delta = 0).
// ---------------------------------------------------------------------------
template <typename R>
TA_HD void fft_pair_slots(const FftTables<R>& t, int r, int i, int* p, int* pp, int* slot_d) {
    if (r) {
        *p = i;
        *pp = t.H - 1 - i;
    } else {
        *p = (int)t.own0[i];
        *pp = (int)t.pair0[*p];
    }
    *slot_d = (*pp != *p) ? *pp : t.H;
}
template <typename R> TA_HD int fft_npairs(const FftTables<R>& t, int r) {
    return r ? (t.H + 1) / 2 : t.npairs0;
}

template <typename R>
TA_HD void fft_accumulate(int tid, int nthr, const cplx<R>* buf, R* sd, const FftTables<R>& t, int r) {
    const int np = fft_npairs(t, r);
    for (int i = tid; i < np; i += nthr) {
        int p, pp, sl;
        fft_pair_slots(t, r, i, &p, &pp, &sl);
        cplx<R> U = buf[p], V = buf[pp];
        cplx<R> w = tw_get(t, 2 * (int)t.ftab[p] + r);
        R nu = cnorm2(U), nv = cnorm2(V);
        R B = U.x * V.y + U.y * V.x;
        R del = (R)2 * w.x * B + w.y * (nu - nv);
        if (p != pp) {
            sd[p] += nu + nv;
            sd[pp] += del;
        } else {
            sd[p] += nu + nv;           // |S|^2/2 with U' = U
            if (sl == t.H && r == 0 && t.ftab[p] == 0) sd[t.H] += del;
        }
    }
}

// ---------------------------------------------------------------------------
// Phase: turn the accumulators into the (scrambled) input of the inverse
// transform:  A[f] = (Sigma + Im(w) Delta) + i Re(w) Delta,
//             A[2H-f] = (Sigma - Im(w) Delta) + i Re(w) Delta.
// ---------------------------------------------------------------------------
template <typename R>
TA_HD void fft_build(int tid, int nthr, cplx<R>* buf, const R* sd, const FftTables<R>& t, int r) {
    const int np = fft_npairs(t, r);
    for (int i = tid; i < np; i += nthr) {
        int p, pp, sl;
        fft_pair_slots(t, r, i, &p, &pp, &sl);
        cplx<R> w = tw_get(t, 2 * (int)t.ftab[p] + r);
        R sig = sd[p];
        R del = (p != pp) ? sd[pp] : ((r == 0 && t.ftab[p] == 0) ? sd[t.H] : (R)0);
        buf[p] = cmake<R>(sig + w.y * del, w.x * del);
        if (p != pp) buf[pp] = cmake<R>(sig - w.y * del, w.x * del);
    }
}

// ---------------------------------------------------------------------------
// Phase: output.  After the inverse transform buf[n] = V_r[n] (natural order).
// Residue 0 parks V_0 in the output row (raw); residue 1 adds its twisted part,
// normalises by L (T - k) and stores the finished lags, and adds them into this
// CTA's partial atom-sum row.  `row` must hold 2*ceil(T/2) doubles.
// ---------------------------------------------------------------------------
template <typename R>
TA_HD void fft_store(int tid, int nthr, const cplx<R>* buf, double* row, double* partial,
                     const FftTables<R>& t, int r) {
    const int nh = (t.T + 1) / 2;
    cplx<double>* out = reinterpret_cast<cplx<double>*>(row);
    cplx<double>* part = reinterpret_cast<cplx<double>*>(partial);
    for (int n = tid; n < nh; n += nthr) {
        cplx<R> v = buf[n];
        if (r == 0) {
            out[n] = cmake<double>((double)v.x, (double)v.y);
        } else {
            cplx<R> tv = cmulc(v, tw_get(t, 2 * n));      // V_1[n] * conj(w_L^{2n})
            cplx<double> a = out[n];
            double re = a.x + (double)tv.x, im = a.y + (double)tv.y;
            int k0 = 2 * n, k1 = 2 * n + 1;
            double o0 = re / ((double)t.L * (double)(t.T - k0));
            double o1 = (k1 < t.T) ? im / ((double)t.L * (double)(t.T - k1)) : 0.0;
            out[n] = cmake<double>(o0, o1);
            cplx<double> ps = part[n];
            part[n] = cmake<double>(ps.x + o0, ps.y + o1);
        }
    }
}

}  // namespace ta

// Small DFTs held entirely in registers, with compile-time twiddle constants.
// Building blocks of the fast FFT-autocorrelation kernel (k1_fast.cuh), which
// replaces the arithmetic of tidynamics.acf as called from
// transport_analysis/velocityautocorr.py:211-213.
//
//   Dft<N, DIR>::run(v):  v[k] <- sum_q v[q] exp(DIR * 2 pi i q k / N)
// in place, natural order in and out, unnormalised.  DIR = -1 is the forward
// kernel, DIR = +1 the inverse (conjugate) kernel.  Every index is a
// compile-time constant after unrolling, so v[] lives in registers.
#pragma once
#include <utility>
#include "ta_common.cuh"

namespace ta {

// ---------------------------------------------------------------------------
// compile-time trigonometry: cos / sin of 2 pi k / n, exact on the axes and
// diagonals, a couple of ulp elsewhere (Taylor series on an argument reduced
// to [0, pi/4] with integer arithmetic).
// ---------------------------------------------------------------------------
namespace ct {
constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kSqrtHalf = 0.70710678118654752440084436210485;

constexpr double sin_small(double a) {   // |a| <= pi/4
    const double a2 = a * a;
    double s = 1.0 / 51090942171709440000.0;           // 1/21!
    s = 1.0 / 121645100408832000.0 - a2 * s;             // 1/19!
    s = 1.0 / 355687428096000.0 - a2 * s;                // 1/17!
    s = 1.0 / 1307674368000.0 - a2 * s;                  // 1/15!
    s = 1.0 / 6227020800.0 - a2 * s;                     // 1/13!
    s = 1.0 / 39916800.0 - a2 * s;                       // 1/11!
    s = 1.0 / 362880.0 - a2 * s;                         // 1/9!
    s = 1.0 / 5040.0 - a2 * s;                           // 1/7!
    s = 1.0 / 120.0 - a2 * s;                            // 1/5!
    s = 1.0 / 6.0 - a2 * s;                              // 1/3!
    return a - a * a2 * s;
}
constexpr double cos_small(double a) {
    const double a2 = a * a;
    double c = 1.0 / 2432902008176640000.0;              // 1/20!
    c = 1.0 / 6402373705728000.0 - a2 * c;               // 1/18!
    c = 1.0 / 20922789888000.0 - a2 * c;                 // 1/16!
    c = 1.0 / 87178291200.0 - a2 * c;                    // 1/14!
    c = 1.0 / 479001600.0 - a2 * c;                      // 1/12!
    c = 1.0 / 3628800.0 - a2 * c;                        // 1/10!
    c = 1.0 / 40320.0 - a2 * c;                          // 1/8!
    c = 1.0 / 720.0 - a2 * c;                            // 1/6!
    c = 1.0 / 24.0 - a2 * c;                             // 1/4!
    c = 0.5 - a2 * c;                                    // 1/2!
    return 1.0 - a2 * c;
}
// angle = 2 pi m / n with 0 <= 8 m <= n
constexpr double cos_oct(long long m, long long n) {
    return m == 0 ? 1.0 : (8 * m == n ? kSqrtHalf : cos_small(kTwoPi * (double)m / (double)n));
}
constexpr double sin_oct(long long m, long long n) {
    return m == 0 ? 0.0 : (8 * m == n ? kSqrtHalf : sin_small(kTwoPi * (double)m / (double)n));
}
// first quadrant: angle = 2 pi m / n with 0 <= 4 m <= n
constexpr double cos_quad(long long m, long long n) {
    return (8 * m <= n) ? cos_oct(m, n) : sin_oct(n - 4 * m, 4 * n);
}
constexpr double sin_quad(long long m, long long n) {
    return (8 * m <= n) ? sin_oct(m, n) : cos_oct(n - 4 * m, 4 * n);
}
constexpr long long pmod(long long k, long long n) { return ((k % n) + n) % n; }
constexpr double cos2pi(long long k, long long n) {
    const long long m = pmod(k, n);
    const long long q = (4 * m) / n;                 // quadrant
    const long long m4 = 4 * m - q * n, n4 = 4 * n;  // angle - q*pi/2 = 2 pi m4 / n4
    return q == 0 ? cos_quad(m4, n4) : q == 1 ? -sin_quad(m4, n4) : q == 2 ? -cos_quad(m4, n4) : sin_quad(m4, n4);
}
constexpr double sin2pi(long long k, long long n) {
    const long long m = pmod(k, n);
    const long long q = (4 * m) / n;
    const long long m4 = 4 * m - q * n, n4 = 4 * n;
    return q == 0 ? sin_quad(m4, n4) : q == 1 ? cos_quad(m4, n4) : q == 2 ? -sin_quad(m4, n4) : -cos_quad(m4, n4);
}
}  // namespace ct

// static loop: f(std::integral_constant<int, I>) for I in [0, N)
template <int I, int N, class F>
TA_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

using cd = cplx<double>;
using cf = cplx<float>;

// real type of a complex value type
template <class C> struct real_of;
template <class R> struct real_of<cplx<R>> { using type = R; };
template <class C> using real_t = typename real_of<C>::type;

// a * exp(DIR * 2 pi i K / N), constant folded; axis and diagonal cases are cheaper
template <int K, int N, int DIR, class C>
TA_HD C mul_tw(C a) {
    using R = real_t<C>;
    constexpr long long k = ct::pmod(K, N);
    if constexpr (k == 0) {
        return a;
    } else if constexpr (2 * k == N) {
        return cmake<R>(-a.x, -a.y);
    } else if constexpr (4 * k == N) {               // exp(DIR i pi/2) = DIR i
        return DIR < 0 ? cmul_mi(a) : cmul_pi(a);
    } else if constexpr (4 * k == 3 * N) {           // exp(DIR i 3pi/2) = -DIR i
        return DIR < 0 ? cmul_pi(a) : cmul_mi(a);
    } else {
        constexpr R c = (R)ct::cos2pi(k, N);
        constexpr R s = (R)((DIR < 0 ? -1.0 : 1.0) * ct::sin2pi(k, N));
        return cmake<R>(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

template <int N, int DIR> struct Dft;

template <int DIR> struct Dft<1, DIR> {
    template <class C> static TA_HD void run(C*) {}
};

template <int DIR> struct Dft<2, DIR> {
    template <class C> static TA_HD void run(C* v) {
        C t = csub(v[0], v[1]);
        v[0] = cadd(v[0], v[1]);
        v[1] = t;
    }
};

template <int DIR> struct Dft<3, DIR> {
    template <class C> static TA_HD void run(C* v) {
        using R = real_t<C>;
        constexpr R hs3 = (R)0.86602540378443864676372317075294;   // sin(pi/3)
        constexpr R half = (R)0.5;
        C t = cadd(v[1], v[2]);
        C u = csub(v[1], v[2]);
        C m = cmake<R>(v[0].x - half * t.x, v[0].y - half * t.y);
        C su = cmake<R>(hs3 * u.x, hs3 * u.y);
        C ru = DIR < 0 ? cmul_mi(su) : cmul_pi(su);
        v[0] = cadd(v[0], t);
        v[1] = cadd(m, ru);
        v[2] = csub(m, ru);
    }
};

template <int DIR> struct Dft<4, DIR> {
    template <class C> static TA_HD void run(C* v) {
        C t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        C t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        C t3 = DIR < 0 ? cmul_mi(d) : cmul_pi(d);
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

template <int DIR> struct Dft<5, DIR> {
    template <class C> static TA_HD void run(C* v) {
        using R = real_t<C>;
        constexpr R c1 = (R)0.30901699437494742410229341718282;    // cos(2pi/5)
        constexpr R c2 = (R)-0.80901699437494742410229341718282;   // cos(4pi/5)
        constexpr R s1 = (R)0.95105651629515357211643933337938;    // sin(2pi/5)
        constexpr R s2 = (R)0.58778525229247312916870595463907;    // sin(4pi/5)
        C t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
        C t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        C a1 = cmake<R>(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        C a2 = cmake<R>(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        C q1 = cmake<R>(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
        C q2 = cmake<R>(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
        C b1 = DIR < 0 ? cmul_mi(q1) : cmul_pi(q1);
        C b2 = DIR < 0 ? cmul_mi(q2) : cmul_pi(q2);
        v[0] = cmake<R>(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
        v[1] = cadd(a1, b1);
        v[4] = csub(a1, b1);
        v[2] = cadd(a2, b2);
        v[3] = csub(a2, b2);
    }
};

// N = A * B by one Cooley-Tukey step in registers:
//   q = A m + a,  k = k1 + B k2:
//   y[k1 + B k2] = sum_a w_A^{a k2} ( w_N^{a k1} sum_m x[A m + a] w_B^{m k1} )
template <int A, int B, int DIR>
struct DftComposite {
    template <class C> static TA_HD void run(C* v) {
        constexpr int N = A * B;
        C u[N];   // u[a * B + k1]
        static_for<0, A>([&](auto ia) {
            constexpr int a = decltype(ia)::value;
            C t[B];
            static_for<0, B>([&](auto im) {
                constexpr int m = decltype(im)::value;
                t[m] = v[A * m + a];
            });
            Dft<B, DIR>::run(t);
            static_for<0, B>([&](auto ik) {
                constexpr int k1 = decltype(ik)::value;
                u[a * B + k1] = mul_tw<a * k1, N, DIR>(t[k1]);
            });
        });
        static_for<0, B>([&](auto ik) {
            constexpr int k1 = decltype(ik)::value;
            C s[A];
            static_for<0, A>([&](auto ia) {
                constexpr int a = decltype(ia)::value;
                s[a] = u[a * B + k1];
            });
            Dft<A, DIR>::run(s);
            static_for<0, A>([&](auto ik2) {
                constexpr int k2 = decltype(ik2)::value;
                v[k1 + B * k2] = s[k2];
            });
        });
    }
};

template <int DIR> struct Dft<6, DIR> : DftComposite<2, 3, DIR> {};
template <int DIR> struct Dft<8, DIR> : DftComposite<2, 4, DIR> {};
template <int DIR> struct Dft<10, DIR> : DftComposite<2, 5, DIR> {};
template <int DIR> struct Dft<12, DIR> : DftComposite<4, 3, DIR> {};
template <int DIR> struct Dft<16, DIR> : DftComposite<4, 4, DIR> {};
template <int DIR> struct Dft<20, DIR> : DftComposite<4, 5, DIR> {};
template <int DIR> struct Dft<24, DIR> : DftComposite<8, 3, DIR> {};

}  // namespace ta

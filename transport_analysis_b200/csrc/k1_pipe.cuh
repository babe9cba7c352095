// K1 pipelined: the three-pass FFT autocorrelation of k1_fast.cuh (same mathematics, same tables, same
// butterflies; replaces tidynamics.acf as called from transport_analysis/velocityautocorr.py:210-214) with the
// CTA-wide barriers taken out of the warps' way.
//
// What bounds the barrier version (profiles/r02_k1_fp64_ncu_summary.json): between two __syncthreads() every warp
// does exactly one load -> butterfly -> store round, all ten at the same time, so the shared-memory pipe and the FP64
// pipe take turns (51 % + 49 % of the elapsed time) although they run side by side when different warps feed them.
// Here a chain (one H-point transform: a forward transform F of one series, or the inverse I of one residue) is cut
// at its one all-to-all exchange into a FRONT stage (F: P1; I: P3' + P2') and a BACK stage (F: P2 + P3 with the pair
// accumulation; I: P1' + output), two FFT buffers alternate between consecutive chains, and a warp runs
//     front(chain k+1), back(chain k), front(chain k+2), back(chain k+1), ...
// with mbarriers instead of __syncthreads(): "full[b]" (every warp has finished the front stage that fills buffer b)
// is waited for one stage after it was signalled, "empty[b]" (every warp has finished the back stage that read b) only
// where the next writer of b is not the reader itself, and then only just before the stores of P1, after its loads
// and arithmetic.  A warp therefore never waits for a barrier the others have not passed long ago, warps drift apart
// (the two without P1 work first), and loads, arithmetic and stores of different warps overlap.
//
// Chain order of a particle's residue r: F(d = 0), I(previous residue), F(1), ..., F(D-1).  The inverse of a
// residue follows the first forward chain of the next one, so that its front stage (which consumes the pair
// accumulators) comes after the back stage of F(D-1) (which completes them) and before the back stage of the
// next residue's F(0) (which starts them again): one set of accumulators.
//
// Two buffers of H complex doubles do not fit beside an H-point double prefetch buffer; the kernel is for float
// series under FP64 arithmetic (float sources: K0 stores them as they are, P1 upcasts, bit-identical), whose
// prefetch buffer is half the size: 2 x 85 KB + 40 KB + tables = 218 KB at R1 = 20.
#pragma once
#include "k1_fast.cuh"

namespace ta {

constexpr int k1p_smem_bytes(int R1, int real_bytes, int store_bytes, bool pref) {
    return (2 * (256 * R1 + 16 * R1) + 256 + 240) * 2 * real_bytes + (pref ? 256 * R1 * 2 * store_bytes : 0) + 64;
}
// instantiated for the radices whose CTA has at least the eight warps of P1 (256 butterflies of radix R1)
constexpr bool k1p_supported(int R1) { return R1 == 16 || R1 == 20; }

enum : int { K1P_NONE = 0, K1P_F = 1, K1P_I = 2 };

#if defined(TA_EXPERIMENTS) && defined(__CUDA_ARCH__)
#define K1P_TRACE(ev) do { if (A.trace && bid == 0 && lane == 0 && pos <= 96) A.trace[(((tid >> 5) * 96 + (pos - 1)) * 8) + (ev)] = clock64(); } while (0)
#else
#define K1P_TRACE(ev) do { } while (0)
#endif

template <int R1, int NT, class Ctx, typename RT, typename ST, bool PREF>
TA_HD void k1p_body(const K1FArgs<RT, ST>& A, unsigned char* smem_raw, int tid, int bid, int nblk) {
    using C = cplx<RT>;
    using CS = cplx<ST>;
    constexpr int H = 256 * R1;
    constexpr int NV = 16 * R1;
    constexpr int NG = (R1 + 3) / 4;
    constexpr int NW = NT / 32;              // warps of the CTA
    constexpr int NPW = 256 / 32;            // warps with P1 / P1' work (thread j < 256 owns column j)
    static_assert(NT == NV && NT >= 256 && NT % 32 == 0, "one P2 / P3 butterfly per thread, one P1 column per thread");
    if (bid >= A.natoms) return;

    C* buf0 = reinterpret_cast<C*>(smem_raw);        // H + H/16 elements each, padded layout e + (e >> 4)
    C* buf1 = buf0 + (H + H / 16);
    C* s_om = buf1 + (H + H / 16);                   // 256
    C* s_tw2 = s_om + 256;                           // 240
    CS* pre = reinterpret_cast<CS*>(s_tw2 + 240);    // PREF: H elements, the series of the coming forward chain as it lies in HBM
    unsigned long long* mb = reinterpret_cast<unsigned long long*>(pre + (PREF ? H : 0));
    unsigned long long* mb_pre = mb;                 // the bulk copy of `pre` has landed
    unsigned long long* mb_full0 = mb + 1;           // front stage into buffer b finished by all warps
    unsigned long long* mb_full1 = mb + 2;
    unsigned long long* mb_empty0 = mb + 3;          // back stage on buffer b finished by all warps
    unsigned long long* mb_empty1 = mb + 4;
    unsigned* pre_cnt = reinterpret_cast<unsigned*>(mb + 5);   // P1 warps that have read `pre`
    const C czero = cmake<RT>((RT)0, (RT)0);
    const int lane = tid & 31;
    const bool p1_thread = tid < 256;

    for (int i = tid; i < 256; i += NT) s_om[i] = A.omega[i];
    for (int i = tid; i < 240; i += NT) s_tw2[i] = A.tw2[i];
    if (tid == 0) {
        Ctx::mbar_init(mb_pre, 1);
        Ctx::mbar_init(mb_full0, NW); Ctx::mbar_init(mb_full1, NW);
        Ctx::mbar_init(mb_empty0, NW); Ctx::mbar_init(mb_empty1, NW);
        *pre_cnt = 0u;
    }
    Ctx::sync();

    const int nh = A.nh;
    const unsigned ser_bytes = ((unsigned)nh * (unsigned)sizeof(CS) + 15u) & ~15u;
    if (PREF && tid == 0) Ctx::bulk_load(pre, A.series + (size_t)bid * A.DS * A.Tld, ser_bytes, mb_pre);
    const int j2 = tid & 15;
    cd* part = reinterpret_cast<cd*>(A.partial + (size_t)bid * A.Tld);
    const C* inv2 = reinterpret_cast<const C*>(A.inv);

    // per-thread synchronisation state, bit b = buffer b (no indexed arrays: they would live in local memory)
    unsigned ph_pre = 0u;        // parity of the `pre` phase to wait for next
    unsigned ph_full = 0u;       // parity of the full[b] phase to wait for next
    unsigned par_empty = 0u;     // parity of the empty[b] phase this thread's next arrival belongs to
    unsigned pend_empty = 0u;    // the previous empty[b] phase has not been waited for yet
    unsigned last_kind = 0u;     // 2 bits per buffer: kind of the last chain on it
    // wait for the empty[b] phase of this warp's last arrival.  Unambiguous with one parity bit: the barrier is in that
    // phase or in the next one, which cannot complete without this warp's next arrival.
    auto wait_empty = [&](int b) {
        Ctx::mbar_wait(b ? mb_empty1 : mb_empty0, ((par_empty >> b) & 1u) ^ 1u);
        pend_empty &= ~(1u << b);
    };
    // this warp has finished a back stage on buffer b.  The arriving lane must not arrive twice within one phase, so it
    // first waits for the phase of its previous arrival if no hazard wait has done that (the other lanes must NOT wait
    // here: by the time they looked, the arrival below may have completed the next phase as well).
    auto arrive_empty = [&](int b) {
        Ctx::sync_warp();
        if (lane == 0) {
            if (pend_empty & (1u << b)) Ctx::mbar_wait(b ? mb_empty1 : mb_empty0, ((par_empty >> b) & 1u) ^ 1u);
            Ctx::mbar_arrive(b ? mb_empty1 : mb_empty0);
        }
        par_empty ^= 1u << b;
        pend_empty |= 1u << b;
    };

    RT acc_s[8], acc_d[8], acc8 = (RT)0;
#pragma unroll
    for (int m = 0; m < 8; ++m) { acc_s[m] = (RT)0; acc_d[m] = (RT)0; }
    int pos = 0;                 // chains emitted so far
    // start the second warp of every sub-partition one stage late: its front stages then coincide with the first
    // warp's back stages, and the barriers' one stage of slack keeps them apart
    if (A.stagger > 0 && (tid >> 5) >= 4 && (tid >> 5) < 8) Ctx::delay((unsigned)A.stagger);

    // ------------------------------------------------------------------ front stage of a forward chain: P1
    auto front_F = [&](int atom, int r, int d, int b) {
        C* buf = b ? buf1 : buf0;
        const bool hazard = ((last_kind >> (2 * b)) & 3u) == (unsigned)K1P_F;   // the previous chain on b was read by other warps
        if (p1_thread) {
            const int j = tid;
            C x[R1];
            if (PREF) {
                Ctx::mbar_wait(mb_pre, ph_pre);
#pragma unroll
                for (int q = 0; q < R1; ++q) {
                    const int n = j + 256 * q;
                    if (n < nh) { const CS zs = pre[n]; x[q] = cmake<RT>((RT)zs.x, (RT)zs.y); }
                    else x[q] = czero;
                }
            } else {
                // straight from L2 / HBM: the other warps' stages cover the latency
                const CS* src = reinterpret_cast<const CS*>(A.series + ((size_t)atom * A.DS + d) * A.Tld);
                CS zs[R1];
#pragma unroll
                for (int q = 0; q < R1; ++q) {
                    const int n = j + 256 * q;
                    zs[q] = Ctx::ld_stream(src + (n < nh ? n : nh - 1));
                }
#pragma unroll
                for (int q = 0; q < R1; ++q) {
                    const int n = j + 256 * q;
                    x[q] = (n < nh) ? cmake<RT>((RT)zs[q].x, (RT)zs[q].y) : czero;
                }
            }
            // `pre` is in registers (shared-memory accesses of a warp are performed in order, so the counter update below
            // follows the loads): the last P1 warp to say so hands the buffer to the bulk-copy engine for the next forward chain
            if (PREF) Ctx::sync_warp();
            if (PREF && lane == 0) {
                if (Ctx::atomic_inc_acq_rel(pre_cnt) == (unsigned)(NPW - 1)) {
                    *pre_cnt = 0u;
                    const ST* ser = A.series + (size_t)atom * A.DS * A.Tld;
                    const ST* nxt = nullptr;
                    if (d + 1 < A.D) nxt = ser + (size_t)(d + 1) * A.Tld;
                    else if (r == 0) nxt = ser;
                    else if (atom + nblk < A.natoms) nxt = ser + (size_t)nblk * A.DS * A.Tld;
                    if (nxt) Ctx::bulk_load(pre, nxt, ser_bytes, mb_pre);
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
#pragma unroll
            for (int k = 0; k < R1; ++k) {
                C y = x[k];
                if (k >= 4) y = cmul(y, g[k >> 2]);
                if (r || (k & 3)) y = cmul(y, e[k & 3]);
                x[k] = y;
            }
            K1P_TRACE(1);
            if (hazard) wait_empty(b);
            K1P_TRACE(2);
            C* dst = buf + j + (j >> 4);
#pragma unroll
            for (int k = 0; k < R1; ++k) dst[272 * k] = x[k];
        }
        ph_pre ^= 1u;
        Ctx::sync_warp();
        if (lane == 0) Ctx::mbar_arrive(b ? mb_full1 : mb_full0);
    };

    // ------------------------------------------------------------------ back stage of a forward chain: P2, P3 + pair accumulation
    auto back_F = [&](int r, int b) {
        C* buf = b ? buf1 : buf0;
        Ctx::mbar_wait(b ? mb_full1 : mb_full0, (ph_full >> b) & 1u);
        ph_full ^= 1u << b;
        K1P_TRACE(4);
        const uint32_t mp = A.map[r * NV + tid];
        {
            const int blk2 = (int)((mp >> 16) & 0xffu);
            const int p2base = blk2 * 272 + j2;
            C x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = buf[p2base + 17 * q];
            Dft<16, -1>::run(x);
#pragma unroll
            for (int k = 1; k < 16; ++k) x[k] = cmul(x[k], s_tw2[(k - 1) * 16 + j2]);
#pragma unroll
            for (int k = 0; k < 16; ++k) buf[p2base + 17 * k] = x[k];
        }
        Ctx::sync_warp();
        {
            const C wb = A.wbase[r * NV + tid];
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
                acc_s[m] += nu + nv;
                acc_d[m] += (RT)2 * w.x * B + w.y * (nu - nv);
            });
            if (self0) acc8 += (RT)2 * cnorm2(v[8]);
        }
        arrive_empty(b);
    };

    // ------------------------------------------------------------------ front stage of an inverse chain: P3' from the accumulators, P2'
    auto front_I = [&](int r, int b) {
        C* buf = b ? buf1 : buf0;
        const bool hazard = ((last_kind >> (2 * b)) & 3u) == (unsigned)K1P_I;   // D = 1: the previous inverse was read column-wise
        const uint32_t mp = A.map[r * NV + tid];
        {
            const C wb = A.wbase[r * NV + tid];
            const int p3base = (int)(mp & 0xffffu) * 17;
            const bool self0 = (mp & K1F_SELF0) != 0, self8 = (mp & K1F_SELF8) != 0;
            C v[16], ap[8], rc[8];
            static_for<0, 8>([&](auto im) {
                constexpr int m = decltype(im)::value;
                const C w = mul_tw<m, 32, -1>(wb);
                const RT sig = acc_s[m], del = acc_d[m];
                v[m] = cmake<RT>(sig + w.y * del, w.x * del);
                ap[m] = cmake<RT>(sig - w.y * del, w.x * del);
                rc[m].x = Ctx::shfl_xor16(ap[m].x);
                rc[m].y = Ctx::shfl_xor16(ap[m].y);
            });
            static_for<8, 16>([&](auto ii) {
                constexpr int idx = decltype(ii)::value;
                C val = rc[15 - idx];
                if (self8) val = ap[15 - idx];
                if (self0) val = (idx == 8) ? cmake<RT>(acc8, (RT)0) : ap[16 - idx];
                v[idx] = val;
            });
            acc8 = (RT)0;
#pragma unroll
            for (int m = 0; m < 8; ++m) { acc_s[m] = (RT)0; acc_d[m] = (RT)0; }
            Dft<16, +1>::run(v);
            if (hazard) wait_empty(b);
#pragma unroll
            for (int q = 0; q < 16; ++q) buf[p3base + q] = v[q];
        }
        Ctx::sync_warp();
        {
            const int blk2 = (int)((mp >> 16) & 0xffu);
            const int p2base = blk2 * 272 + j2;
            C x[16];
            x[0] = buf[p2base];
#pragma unroll
            for (int k = 1; k < 16; ++k) x[k] = cmulc(buf[p2base + 17 * k], s_tw2[(k - 1) * 16 + j2]);
            Dft<16, +1>::run(x);
#pragma unroll
            for (int q = 0; q < 16; ++q) buf[p2base + 17 * q] = x[q];
        }
        Ctx::sync_warp();
        if (lane == 0) Ctx::mbar_arrive(b ? mb_full1 : mb_full0);
    };

    // ------------------------------------------------------------------ back stage of an inverse chain: P1' + output
    auto back_I = [&](int atom, int r, int b) {
        C* buf = b ? buf1 : buf0;
        Ctx::mbar_wait(b ? mb_full1 : mb_full0, (ph_full >> b) & 1u);
        ph_full ^= 1u << b;
        K1P_TRACE(4);
        if (p1_thread) {
            const int j = tid;
            cd* row = reinterpret_cast<cd*>(A.by_particle + (size_t)atom * A.Tld);
            C e[4], g[NG];
            k1f_p1_twiddles<R1>(s_om[j], r, e, g);
            C* own = buf + j + (j >> 4);
            C x[R1];
#pragma unroll
            for (int k = 0; k < R1; ++k) x[k] = own[272 * k];
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
            if (r == 0) {
#pragma unroll
                for (int q = 0; q < R1; ++q) {
                    const int n = j + 256 * q;
                    if (n < nh) row[n] = cmake<double>((double)x[q].x, (double)x[q].y);   // parked raw; finished by residue 1
                }
            } else {
                // V_1 parked in this thread's own column, the output streamed in chunks with all global loads of a chunk first
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
                        const int nc = n < nh ? n : nh - 1;
                        a[i] = Ctx::ld_stream(row + nc); ps[i] = Ctx::ld_stream(part + nc);
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
                            part[n] = cmake<double>(ps[i].x + o.x, ps[i].y + o.y);
                        }
                    }
                });
            }
        }
        arrive_empty(b);
    };

    // ------------------------------------------------------------------ the chain stream
    int cur_atom = bid, cur_r = 0, slot = 0;         // next forward chain to emit: residue (cur_atom, cur_r), slot within it
    int inv_atom = 0, inv_r = 0;                     // residue whose inverse is outstanding
    bool inv_pending = false;
    int pk = K1P_NONE, p_atom = 0, p_r = 0, p_b = 0; // the chain whose back stage comes next
    for (;;) {
        int kind = K1P_NONE, atom = 0, r = 0, d = 0;
        for (;;) {
            if (cur_atom >= A.natoms) {
                // the tail: the last inverse waits one (empty) step for the back stage that completes its accumulators
                if (inv_pending && pk != K1P_F) { kind = K1P_I; atom = inv_atom; r = inv_r; inv_pending = false; }
                break;
            }
            const int s = slot++;
            if (s == 1) {
                if (!inv_pending) continue;
                kind = K1P_I; atom = inv_atom; r = inv_r; inv_pending = false;
                break;
            }
            if (s <= A.D) { kind = K1P_F; atom = cur_atom; r = cur_r; d = s == 0 ? 0 : s - 1; break; }
            inv_atom = cur_atom; inv_r = cur_r; inv_pending = true;      // every forward chain of the residue is out
            slot = 0;
            cur_r ^= 1;
            if (cur_r == 0) cur_atom += nblk;
        }
        if (kind == K1P_NONE && pk == K1P_NONE) break;
        const int b = pos & 1;                       // consecutive chains alternate between the two buffers
        if (kind != K1P_NONE) ++pos;
#ifdef TA_EMU_TRACE
        if (tid == 225) fprintf(stderr, "step: front kind %d atom %d r %d d %d b %d | back kind %d r %d b %d | lk %x phf %x pare %x pend %x\n", kind, atom, r, d, b, pk, p_r, p_b, last_kind, ph_full, par_empty, pend_empty);
#endif
        K1P_TRACE(0);
        if (kind == K1P_F) front_F(atom, r, d, b);
        else if (kind == K1P_I) front_I(r, b);
        K1P_TRACE(3);
        if (kind != K1P_NONE) last_kind = (last_kind & ~(3u << (2 * b))) | ((unsigned)kind << (2 * b));
        if (pk == K1P_F) back_F(p_r, p_b);
        else if (pk == K1P_I) back_I(p_atom, p_r, p_b);
        K1P_TRACE(5);
        pk = kind; p_atom = atom; p_r = r; p_b = b;
    }
}

}  // namespace ta

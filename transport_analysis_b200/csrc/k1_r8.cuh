// K1 radix-8 path: FFT autocorrelation with H = 512 * R in four passes, 64 R threads per CTA.
//
// Same mathematics as k1_fast.cuh (which replaces tidynamics.acf as called from
// transport_analysis/velocityautocorr.py:210-214): residue chains r = 0, 1 of the zero-padded
// length-2H spectrum, power spectrum from conjugate bin pairs summed over the D series, one inverse
// per residue.  What changes is the shape of the work.  The three-pass kernel keeps a radix-16
// butterfly (64 registers of data + 34 of pair accumulators) in every thread, so an SM holds 10
// warps and every load, barrier and dependent FP64 instruction is exposed (measured: FP64 pipe 48 %
// busy, 7.1 clk between two instructions of a warp; profiles/r01_k1f_cta_shapes.txt).  Here
//   * every butterfly is radix 8 (32 registers of data, 9 accumulators): 64 R threads = 20 warps at
//     R = 10, five per SM sub-partition, <= 100 registers each,
//       P1  radix R, stride 512   fused with the global load and the residue twist      (512 threads)
//       P2  radix 8, stride 64    shared -> registers -> shared
//       P3  radix 8, stride 8     shared -> registers -> shared
//       P4  radix 8, stride 1     shared -> registers, fused with the pair accumulation
//     and the mirror image P4' P3' P2' P1' for the inverse (3 shared-memory round trips per FFT).
//   * the 512-point blocks k and R - k (r = 1: R - 1 - k) are owned by one group of four warps in P2,
//     P3 and P4, so P2 -> P3 needs a 128-thread named barrier and P3 -> P4 a warp-level sync; CTA
//     barriers remain at P1 -> P2 and at the buffer hand-over.
//   * in P4 conjugate-partner butterflies sit on lanes l and l ^ 16 (partner bins travel by shuffle,
//     the (Sigma, Delta) accumulators of a thread's 4 bin pairs stay in registers over the D series).
//   * buffer layout: element e lives at e + (e >> 3); all four access patterns are conflict free.
//   * output: V_0 waits in a second shared buffer, the finished row is normalised with a computed
//     1 / (L (T - k)) and leaves the SM as one bulk (TMA) store of the per-particle row plus one bulk
//     f64 reduce-add into the per-CTA partial row, issued by a warp that is idle in P1'.
#pragma once
#include <cstdint>
#include <algorithm>
#include <vector>
#include "ta_common.cuh"
#include "dft_regs.cuh"
#include "fft_plan.h"
#include "k1_fast.cuh"

namespace ta {

struct K1EArgs {
    const double* series;        // [natoms][D][Tld]
    double* by_particle;         // [natoms][Tld]
    double* partial;             // [grid][Tld]
    const cd* omega;             // [512]      w_{2H}^j
    const cd* tw2;               // [7][64]    w_512^{j k}, k = 1..7
    const cd* tw3;               // [7][8]     w_64^{j k},  k = 1..7
    const uint32_t* map;         // [2][64 R]  per residue and thread, see K1E_* below
    const cd* wbase;             // [2][64 R]  w_L^{2 G0 + r} of the thread's P4 butterfly
    int natoms, D, T, nh;
    long long Tld;
    unsigned* sm_slots;          // [num SMs] zeroed per launch: arrival order of the CTAs of one SM (stagger), may be null
    int stagger;                 // clocks the second CTA of an SM idles before its first particle (two CTAs per SM only)
};

// map word: bits 0-9 P4 butterfly (k * 64 + k2 * 8 + k3), bits 10-16 64-point block of the P3 butterfly
// (k * 8 + k2), bits 17-20 512-point block of the P2 butterfly, bits 21-24 group (named barrier id - 1),
// bits 25-28 warps in the group, bit 30 / 31 self-paired butterflies
constexpr uint32_t K1E_SELF0 = 1u << 30;   // bins 64 R k4: pairs k4 <-> 8 - k4 in-thread, k4 = 0 and 4 alone
constexpr uint32_t K1E_SELF4 = 1u << 31;   // bins 32 R + 64 R k4: pairs k4 <-> 7 - k4 in-thread

constexpr int k1e_threads(int R) { return 64 * R; }
constexpr int k1e_smem_bytes(int R) { return (512 * R + 64 * R + 512 + 448 + 64 + 512 * R) * (int)sizeof(cd); }

template <int R, class Ctx>
TA_HD void k1e_body(const K1EArgs& A, unsigned char* smem_raw, int tid, int bid, int nblk) {
    constexpr int H = 512 * R;
    constexpr int NT = 64 * R;
    constexpr int NG = (R + 3) / 4;
    constexpr int NPART = NT < 512 ? NT : 512;       // threads that take part in P1 / P1'
    constexpr int WISS = NT / 32 - 1;                // warp that issues the bulk store / reduce of a finished row
    constexpr bool ISS_IDLE = NT > 512;              // ... idle in P1' when the CTA has more than 16 warps
    cd* buf = reinterpret_cast<cd*>(smem_raw);       // H + H/8 elements, padded layout
    cd* s_om = buf + (H + H / 8);                    // 512
    cd* s_tw2 = s_om + 512;                          // 448
    cd* s_tw3 = s_tw2 + 448;                         // 56 (+ 8)
    cd* stg = s_tw3 + 64;                            // H, plain layout: V_0, then the finished row

    for (int i = tid; i < 512; i += NT) s_om[i] = A.omega[i];
    for (int i = tid; i < 448; i += NT) s_tw2[i] = A.tw2[i];
    for (int i = tid; i < 56; i += NT) s_tw3[i] = A.tw3[i];
    Ctx::sync();

    const int nh = A.nh;
    const int j0 = tid & 63, j1 = tid & 7;
    const double Ld = (double)(4 * H);
    cd* part = reinterpret_cast<cd*>(A.partial + (size_t)bid * A.Tld);

    // Two CTAs on one SM run the same phases at the same pace; offset by about half a chain, the FP64-bound
    // phases of one meet the shared-memory-bound phases of the other.
    if (A.stagger > 0 && A.sm_slots != nullptr) Ctx::stagger_second_cta(A.sm_slots, A.stagger, tid);

    for (int atom = bid; atom < A.natoms; atom += nblk) {
        const double* ser = A.series + (size_t)atom * A.D * A.Tld;
        cd* row = reinterpret_cast<cd*>(A.by_particle + (size_t)atom * A.Tld);
        for (int r = 0; r < 2; ++r) {
            const uint32_t mp = A.map[r * NT + tid];
            const cd wb = A.wbase[r * NT + tid];
            const int p4base = (int)(mp & 0x3ffu) * 9;
            const int p3base = (int)((mp >> 10) & 0x7fu) * 72 + j1;
            const int p2base = (int)((mp >> 17) & 0xfu) * 576 + j0 + (j0 >> 3);
            const int grp_bar = 1 + (int)((mp >> 21) & 0xfu), grp_thr = 32 * (int)((mp >> 25) & 0xfu);
            const bool self0 = (mp & K1E_SELF0) != 0, self4 = (mp & K1E_SELF4) != 0;
            double acc_s[4], acc_d[4], accx = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) { acc_s[m] = 0.0; acc_d[m] = 0.0; }

            for (int d = 0; d < A.D; ++d) {
                // ---------------- P1: global -> registers -> shared
                const cd* src = reinterpret_cast<const cd*>(ser + (size_t)d * A.Tld);
                for (int j = tid; j < 512; j += NT) {
                    cd x[R];
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int n = j + 512 * q;
                        x[q] = (n < nh) ? Ctx::ld_stream(src + n) : cmake<double>(0.0, 0.0);
                    }
                    if (r) {
                        static_for<1, R>([&](auto iq) {
                            constexpr int q = decltype(iq)::value;
                            x[q] = mul_tw<q, 2 * R, -1>(x[q]);
                        });
                    }
                    Dft<R, -1>::run(x);
                    cd e[4], g[NG];
                    k1f_p1_twiddles<R>(s_om[j], r, e, g);
                    cd* dst = buf + j + (j >> 3);
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        cd y = x[k];
                        if (k >= 4) y = cmul(y, g[k >> 2]);
                        if (r || (k & 3)) y = cmul(y, e[k & 3]);
                        dst[576 * k] = y;
                    }
                }
                Ctx::sync();
                // ---------------- P2: radix 8, stride 64
                {
                    cd x[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = buf[p2base + 72 * q];
                    Dft<8, -1>::run(x);
#pragma unroll
                    for (int k = 1; k < 8; ++k) x[k] = cmul(x[k], s_tw2[(k - 1) * 64 + j0]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) buf[p2base + 72 * k] = x[k];
                }
                Ctx::bar_sync(grp_bar, grp_thr);     // P3 of this group reads what this group's P2 wrote
                // ---------------- P3: radix 8, stride 8
                {
                    cd x[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = buf[p3base + 9 * q];
                    Dft<8, -1>::run(x);
#pragma unroll
                    for (int k = 1; k < 8; ++k) x[k] = cmul(x[k], s_tw3[(k - 1) * 8 + j1]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) buf[p3base + 9 * k] = x[k];
                }
                Ctx::sync_warp();                    // P4 of this warp reads what this warp's P3 wrote
                // ---------------- P4: radix 8, stride 1, + pair accumulation
                {
                    cd v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = buf[p4base + q];
                    Dft<8, -1>::run(v);
                    static_for<0, 4>([&](auto im) {
                        constexpr int m = decltype(im)::value;
                        cd snd = v[7 - m], rec;
                        rec.x = Ctx::shfl_xor16(snd.x);
                        rec.y = Ctx::shfl_xor16(snd.y);
                        if (self4) rec = snd;
                        if (self0) rec = v[(8 - m) & 7];
                        const cd w = mul_tw<m, 16, -1>(wb);
                        const cd U = v[m];
                        const double nu = cnorm2(U), nv = cnorm2(rec);
                        const double B = U.x * rec.y + U.y * rec.x;
                        acc_s[m] += nu + nv;
                        acc_d[m] += 2.0 * w.x * B + w.y * (nu - nv);
                    });
                    if (self0) accx += 2.0 * cnorm2(v[4]);
                }
                if (d + 1 < A.D) Ctx::sync();        // P1 of the next series overwrites the buffer
            }

            // ---------------- inverse: build from the accumulators, P4'
            {
                cd v[8], ap[4], rc[4];
                static_for<0, 4>([&](auto im) {
                    constexpr int m = decltype(im)::value;
                    const cd w = mul_tw<m, 16, -1>(wb);
                    const double sig = acc_s[m], del = acc_d[m];
                    v[m] = cmake<double>(sig + w.y * del, w.x * del);
                    ap[m] = cmake<double>(sig - w.y * del, w.x * del);
                    rc[m].x = Ctx::shfl_xor16(ap[m].x);
                    rc[m].y = Ctx::shfl_xor16(ap[m].y);
                });
                static_for<4, 8>([&](auto ii) {
                    constexpr int idx = decltype(ii)::value;
                    cd val = rc[7 - idx];
                    if (self4) val = ap[7 - idx];
                    if (self0) val = (idx == 4) ? cmake<double>(accx, 0.0) : ap[8 - idx];
                    v[idx] = val;
                });
                Dft<8, +1>::run(v);
#pragma unroll
                for (int q = 0; q < 8; ++q) buf[p4base + q] = v[q];
            }
            Ctx::sync_warp();
            // ---------------- P3'
            {
                cd x[8];
                x[0] = buf[p3base];
#pragma unroll
                for (int k = 1; k < 8; ++k) x[k] = cmulc(buf[p3base + 9 * k], s_tw3[(k - 1) * 8 + j1]);
                Dft<8, +1>::run(x);
#pragma unroll
                for (int q = 0; q < 8; ++q) buf[p3base + 9 * q] = x[q];
            }
            Ctx::bar_sync(grp_bar, grp_thr);
            // ---------------- P2'
            {
                cd x[8];
                x[0] = buf[p2base];
#pragma unroll
                for (int k = 1; k < 8; ++k) x[k] = cmulc(buf[p2base + 72 * k], s_tw2[(k - 1) * 64 + j0]);
                Dft<8, +1>::run(x);
#pragma unroll
                for (int q = 0; q < 8; ++q) buf[p2base + 72 * q] = x[q];
            }
            // the bulk operations of the previous particle have read (and reduced) the staging buffer
            if (tid == 32 * WISS) Ctx::bulk_wait_all();
            Ctx::sync();
            // ---------------- P1' + output
            for (int j = tid; j < 512; j += NT) {
                cd e[4], g[NG];
                k1f_p1_twiddles<R>(s_om[j], r, e, g);
                const cd* srcb = buf + j + (j >> 3);
                cd x[R];
#pragma unroll
                for (int k = 0; k < R; ++k) x[k] = srcb[576 * k];
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    cd y = x[k];
                    if (k >= 4) y = cmulc(y, g[k >> 2]);
                    if (r || (k & 3)) y = cmulc(y, e[k & 3]);
                    x[k] = y;
                }
                Dft<R, +1>::run(x);
                if (r) {
                    static_for<1, R>([&](auto iq) {
                        constexpr int q = decltype(iq)::value;
                        x[q] = mul_tw<q, 2 * R, +1>(x[q]);
                    });
                }
                // x[q] = V_r[n] (r = 1: already multiplied by conj(w_L^{2n})), n = j + 512 q
                if (r == 0) {
#pragma unroll
                    for (int q = 0; q < R; ++q) stg[j + 512 * q] = x[q];     // V_0 waits here for residue 1
                } else {
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int n = j + 512 * q;
                        const int k = 2 * n;
                        const cd v0 = stg[n];
                        const double sx = k < A.T ? Ctx::rcp(Ld * (double)(A.T - k)) : 0.0;
                        const double sy = k + 1 < A.T ? Ctx::rcp(Ld * (double)(A.T - k - 1)) : 0.0;
                        stg[n] = cmake<double>((v0.x + x[q].x) * sx, (v0.y + x[q].y) * sy);
                    }
                }
            }
            if (r == 1) {
                // every P1' thread has written its part of the finished row -> one warp hands it to the
                // bulk-copy engine: row store + reduce-add into this CTA's partial row (fixed order: the
                // previous particle's group has completed, see bulk_wait_all above)
                const unsigned nbytes = (unsigned)nh * (unsigned)sizeof(cd);
                if (ISS_IDLE) {
                    if (tid < NPART) { Ctx::fence_async_smem(); Ctx::bar_arrive(15, NPART + 32); }
                    else if ((tid >> 5) == WISS) {
                        Ctx::bar_sync(15, NPART + 32);
                        if (tid == 32 * WISS) Ctx::bulk_store_and_add(row, part, stg, nbytes);
                    }
                } else {
                    Ctx::fence_async_smem();
                    Ctx::sync();
                    if (tid == 32 * WISS) Ctx::bulk_store_and_add(row, part, stg, nbytes);
                }
            }
            // no barrier here: P1 of the next chain writes exactly the elements this thread has just read in P1'
        }
    }
    if (tid == 32 * WISS) Ctx::bulk_wait_all();
}

// ---------------------------------------------------------------------------
// Host-side plan.
// ---------------------------------------------------------------------------
struct K1R8Plan {
    int R = 0, H = 0, L = 0, NT = 0, nh = 0;
    std::vector<double> omega;    // 512 x (re, im)
    std::vector<double> tw2;      // 448 x (re, im)
    std::vector<double> tw3;      // 56 x (re, im)
    std::vector<uint32_t> map;    // 2 x NT
    std::vector<double> wbase;    // 2 x NT x (re, im)
};

inline const int* k1e_supported_r(int* n) {
    static const int rs[] = {4, 5, 6, 8, 10, 12};
    *n = (int)(sizeof(rs) / sizeof(rs[0]));
    return rs;
}

// R for a series of T frames, or 0 (same padding rule as k1f_choose_r1)
inline int k1e_choose_r(int64_t T) {
    const int64_t nh = (T + 1) / 2;
    int n;
    const int* rs = k1e_supported_r(&n);
    for (int i = 0; i < n; ++i) {
        const int64_t H = 512 * (int64_t)rs[i];
        if (H >= nh) return (3 * H <= 4 * nh + 3) ? rs[i] : 0;
    }
    return 0;
}

inline int k1e_build_plan(int64_t T, int R, K1R8Plan* p) {
    if (R < 2 || R > 15 || T < 1 || (T + 1) / 2 > 512 * (int64_t)R) return TA_ERR_INVALID;
    const int H = 512 * R, NT = 64 * R;
    p->R = R; p->H = H; p->L = 4 * H; p->NT = NT; p->nh = (int)((T + 1) / 2);
    const int64_t L = p->L;
    p->omega.resize(2 * 512);
    for (int j = 0; j < 512; ++j) ta_twiddle(2 * j, L, &p->omega[2 * j], &p->omega[2 * j + 1]);   // w_{2H}^j = w_L^{2j}
    p->tw2.resize(2 * 448);
    for (int k = 1; k < 8; ++k)
        for (int j = 0; j < 64; ++j)
            ta_twiddle((int64_t)j * k, 512, &p->tw2[2 * ((k - 1) * 64 + j)], &p->tw2[2 * ((k - 1) * 64 + j) + 1]);
    p->tw3.resize(2 * 56);
    for (int k = 1; k < 8; ++k)
        for (int j = 0; j < 8; ++j)
            ta_twiddle((int64_t)j * k, 64, &p->tw3[2 * ((k - 1) * 8 + j)], &p->tw3[2 * ((k - 1) * 8 + j) + 1]);
    p->map.assign(2 * (size_t)NT, 0);
    p->wbase.assign(4 * (size_t)NT, 0.0);

    // bin of output k4 of P4 butterfly b = k * 64 + k2 * 8 + k3:  g = k + R k2 + 8R k3 + 64R k4
    auto bin_of = [&](int b, int k4) { return (b >> 6) + R * ((b >> 3) & 7) + 8 * R * (b & 7) + 64 * R * k4; };
    auto bfly_of = [&](int g) { return (g % R) * 64 + ((g / R) & 7) * 8 + ((g / (8 * R)) & 7); };
    for (int r = 0; r < 2; ++r) {
        auto partner = [&](int b) { const int g = bin_of(b, 1); return bfly_of(r == 0 ? (H - g) % H : H - 1 - g); };
        const int nblk = 8 * R;                       // 64-point blocks, blk = b >> 3
        // slots of 8 + 8 butterflies: lanes c (lower) and 16 + c (upper) hold conjugate partners
        std::vector<std::vector<int>> lower, upper;
        std::vector<int> selfblk;
        for (int blk = 0; blk < nblk; ++blk) {
            const int pblk = partner(blk * 8) >> 3;
            if (pblk == blk) { selfblk.push_back(blk); continue; }
            if (pblk < blk) continue;
            std::vector<int> lo(8), up(8);
            for (int c = 0; c < 8; ++c) { lo[c] = blk * 8 + c; up[c] = partner(lo[c]); }
            lower.push_back(lo); upper.push_back(up);
        }
        if (selfblk.size() == 2) {
            // r = 0: blocks (0, 0) and (0, 4).  Owners are chosen so that the k3 digits of a quarter warp differ
            // (bank-conflict-free P4 loads): k3 = 0..3 of the block whose pairs are k3 <-> 7 - k3, and k3 = 4, 7, 6, 5
            // of the block with pairs k3 <-> 8 - k3 (k3 = 0 and 4 are the self-paired butterflies, lane partners)
            int bs = selfblk[0], bt = selfblk[1];     // bs: contains a butterfly that is its own partner
            if (partner(bs * 8) != bs * 8) std::swap(bs, bt);
            if (partner(bs * 8) != bs * 8 || partner(bs * 8 + 4) != bs * 8 + 4) return TA_ERR_UNSUPPORTED;
            std::vector<int> lo(8), up(8);
            for (int c = 0; c < 4; ++c) { lo[c] = bt * 8 + c; up[c] = partner(lo[c]); }
            lo[4] = bs * 8 + 4; up[4] = bs * 8;
            for (int c = 5; c < 8; ++c) { lo[c] = bs * 8 + (12 - c); up[c] = partner(lo[c]); }
            // P3 duty of this slot: lower quarter warp = block bt, upper = block bs (any split inside the warp works)
            lower.push_back(lo); upper.push_back(up);
        } else if (!selfblk.empty()) {
            return TA_ERR_UNSUPPORTED;
        }
        if ((int)lower.size() != 4 * R) return TA_ERR_UNSUPPORTED;
        // Groups: the slots are ordered so that a group of consecutive warps owns whole 512-point blocks.
        // Sort slots by the set {k, partner k} they touch.
        std::vector<int> order(lower.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        auto keyk = [&](int s) { const int a = lower[s][0] >> 6, b = upper[s][0] >> 6; return a < b ? a : b; };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keyk(a) < keyk(b); });
        // walk the slots; a group closes when the 512-blocks it touches are complete (8 blocks each)
        int slot_idx = 0, group = 0;
        while (slot_idx < (int)order.size()) {
            // collect slots with the same key
            const int key = keyk(order[slot_idx]);
            int end = slot_idx;
            while (end < (int)order.size() && keyk(order[end]) == key) ++end;
            int nslots = end - slot_idx;              // 8 (two 512-blocks) or 4 (one self-paired 512-block)
            // merge two 4-slot families into one group when the next family is also a 4-slot one
            int end2 = end;
            if (nslots == 4 && end < (int)order.size()) {
                const int key2 = keyk(order[end]);
                int e2 = end;
                while (e2 < (int)order.size() && keyk(order[e2]) == key2) ++e2;
                if (e2 - end == 4) end2 = e2;
            }
            nslots = end2 - slot_idx;
            if (nslots % 2) return TA_ERR_UNSUPPORTED;
            const int nwarps_g = nslots / 2;
            if (nwarps_g > 15 || group > 14) return TA_ERR_UNSUPPORTED;
            // 512-blocks of this group, for the P2 duty: 64 butterflies per block
            std::vector<int> kblocks;
            for (int s = slot_idx; s < end2; ++s)
                for (int side = 0; side < 2; ++side) {
                    const std::vector<int>& v = side ? upper[order[s]] : lower[order[s]];
                    for (int c = 0; c < 8; ++c) {
                        const int k = v[c] >> 6;
                        bool seen = false;
                        for (int kk : kblocks) seen = seen || kk == k;
                        if (!seen) kblocks.push_back(k);
                    }
                }
            if ((int)kblocks.size() * 64 != nwarps_g * 32) return TA_ERR_UNSUPPORTED;
            const int warp0 = slot_idx / 2;
            for (int s = slot_idx; s < end2; ++s) {
                const int w = s / 2, i = s & 1;
                for (int side = 0; side < 2; ++side) {
                    const std::vector<int>& v = side ? upper[order[s]] : lower[order[s]];
                    // P3 duty: the 64-block of this quarter warp (for the self slot: lower -> bt, upper -> bs)
                    int blk3 = v[0] >> 3;
                    if (order[s] == (int)lower.size() - 1 && selfblk.size() == 2) blk3 = side ? (upper[order[s]][4] >> 3) : (lower[order[s]][0] >> 3);
                    for (int c = 0; c < 8; ++c) {
                        const int lane = 16 * side + 8 * i + c, tid = 32 * w + lane;
                        const int b = v[c];
                        uint32_t word = (uint32_t)b | ((uint32_t)blk3 << 10);
                        // P2 duty: thread t of the group takes butterfly (kblocks[t / 64], t % 64)
                        const int tg = tid - 32 * warp0;
                        word |= (uint32_t)kblocks[tg >> 6] << 17;
                        word |= (uint32_t)group << 21;
                        word |= (uint32_t)nwarps_g << 25;
                        if (r == 0 && partner(b) == b) word |= (b & 7) == 0 ? K1E_SELF0 : K1E_SELF4;
                        p->map[(size_t)r * NT + tid] = word;
                        const int64_t G0 = bin_of(b, 0);
                        ta_twiddle(2 * G0 + r, L, &p->wbase[2 * ((size_t)r * NT + tid)], &p->wbase[2 * ((size_t)r * NT + tid) + 1]);
                    }
                }
            }
            slot_idx = end2;
            ++group;
        }
    }
    return TA_OK;
}

}  // namespace ta

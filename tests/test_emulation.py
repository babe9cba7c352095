"""The kernels' per-CTA phase functions (fft_core.cuh, windowed_core.cuh),
compiled for the CPU by tests/emu and run phase by phase, against the oracle.
This is what checks index arithmetic, scramble tables and twiddles in the
container that has no GPU; the GPU parity tests are in test_gpu_parity.py."""
import ctypes

import numpy as np
import pytest

import oracle
from _util import assert_close_normwise

DP = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(DP)


def emu_fft(lib, x, nthr=96, f32=0):
    T, D = x.shape
    Tld = (T + 15) // 16 * 16
    ser = np.zeros((D, Tld))
    ser[:, :T] = x.T
    row, part = np.zeros(Tld), np.zeros(Tld)
    assert lib.emu_fft_acf(_p(ser), T, D, Tld, nthr, f32, _p(row), _p(part)) == 0
    assert np.array_equal(row, part)
    assert np.all(row[T:] == 0)
    return row[:T]


def emu_win(lib, x, mode, nwarps=4, f32=0):
    T, D = x.shape
    Tld = (T + 15) // 16 * 16
    ser = np.zeros((D, Tld))
    ser[:, :T] = x.T
    res = np.zeros(T)
    assert lib.emu_windowed(_p(ser), T, D, Tld, mode, nwarps, f32, _p(res)) == 0
    return res


@pytest.mark.parametrize("T", [1, 2, 3, 5, 10, 16, 17, 31, 33, 99, 100, 129, 250, 1000, 2047, 5001, 20001])
def test_fft_phases_match_tidynamics_restatement(emu, T):
    rng = np.random.default_rng(T)
    for D in (1, 2, 3):
        x = rng.standard_normal((T, D))
        assert_close_normwise(emu_fft(emu, x), oracle.tidynamics_acf(x), 1e-11, f"T={T} D={D}")


def test_fft_phases_thread_count_independent(emu):
    x = np.random.default_rng(0).standard_normal((777, 3))
    a = emu_fft(emu, x, nthr=32)
    b = emu_fft(emu, x, nthr=640)
    assert np.array_equal(a, b)


def test_fft_plan_covers_input_and_uses_supported_radices(emu):
    for T in list(range(1, 300)) + [4097, 5000, 5001, 8192, 10000, 16384]:
        H, npz = ctypes.c_int(), ctypes.c_int()
        rad = (ctypes.c_int * 12)()
        assert emu.emu_plan(T, ctypes.byref(H), ctypes.byref(npz), rad) == 0
        r = list(rad)[: npz.value]
        assert H.value >= (T + 1) // 2 and int(np.prod(r)) == H.value
        assert set(r) <= {2, 3, 4, 5, 8}


@pytest.mark.parametrize("T", [1, 2, 15, 16, 17, 47, 48, 49, 100, 333, 1000])
def test_windowed_phases(emu, T):
    rng = np.random.default_rng(100 + T)
    x = rng.standard_normal((T, 3)) * 3 + 1
    prod = np.array([np.sum(x[: T - k] * x[k:]) for k in range(T)])
    sq = np.array([np.sum((x[: T - k] - x[k:]) ** 2) for k in range(T)])
    for nwarps in (1, 5):
        assert_close_normwise(emu_win(emu, x, 0, nwarps), prod, 1e-13, "product")
        got = emu_win(emu, x, 1, nwarps)
        assert got[0] == 0.0
        np.testing.assert_allclose(got, sq, rtol=1e-13, atol=0)


def test_fp32_mode_tolerance(emu):
    x = np.random.default_rng(5).standard_normal((2000, 3))
    assert_close_normwise(emu_fft(emu, x, f32=1), oracle.tidynamics_acf(x), 1e-5, "fp32 fft")
    prod = np.array([np.sum(x[: 2000 - k] * x[k:]) for k in range(2000)])
    assert_close_normwise(emu_win(emu, x, 0, f32=1), prod, 1e-5, "fp32 windowed")


# ------------------------------------------------------------------ K1 fast path under the fiber emulator
def emu_fast(lib, x, nblk=1, R1=None):
    """x: [T, N, D].  Runs k1f_body (k1_fast.cuh) for nblk CTAs of 16*R1 fibers."""
    T, N, D = x.shape
    R1 = R1 or lib.emu_k1fast_r1(T)
    assert R1 > 0
    Tld = (T + 15) // 16 * 16
    ser = np.zeros((N, D, Tld))
    ser[:, :, :T] = x.transpose(1, 2, 0)
    bp, part = np.zeros((N, Tld)), np.zeros((nblk, Tld))
    assert lib.emu_k1fast(_p(ser), T, D, Tld, N, nblk, R1, _p(bp), _p(part)) == 0
    assert np.all(bp[:, T:] == 0)
    return bp[:, :T], part[:, :T]


@pytest.mark.parametrize("T,D,N,nblk", [(1600, 2, 1, 1), (2000, 3, 2, 1), (2047, 1, 1, 1), (3000, 3, 2, 2), (4000, 1, 1, 1),
                                        (5000, 3, 3, 2), (5001, 2, 1, 1), (6000, 3, 1, 1), (8192, 3, 1, 1), (10000, 3, 2, 1)])
def test_fast_fft_kernel_body_matches_tidynamics_restatement(emu, T, D, N, nblk):
    x = np.random.default_rng(T + D).standard_normal((T, N, D))
    bp, part = emu_fast(emu, x, nblk)
    for a in range(N):
        assert_close_normwise(bp[a], oracle.tidynamics_acf(x[:, a, :]), 1e-12, f"T={T} atom {a}")
    np.testing.assert_allclose(part.sum(axis=0), bp.sum(axis=0), rtol=1e-13, atol=1e-13)


def test_fast_fft_path_selection(emu):
    """R1 = 0 (general kernel) for tiny or badly padded lengths, else the smallest even R1 with 256 R1 >= ceil(T/2)."""
    for T, want in [(1, 0), (700, 0), (1533, 0), (1535, 4), (2048, 4), (2049, 0), (2400, 6), (4096, 8), (5000, 10),
                    (5001, 10), (6144, 12), (8192, 16), (10000, 20), (10240, 20), (10241, 0), (20000, 0)]:
        assert emu.emu_k1fast_r1(T) == want, T


def test_fast_fft_ramp_is_exact_enough(emu):
    """Ramp data (the reference's step trajectory) has a 1e7 dynamic range; the closed form is the reference's own."""
    T = 5001
    x = np.repeat(np.arange(T, dtype=np.float64)[:, None, None], 3, axis=2)
    bp, _ = emu_fast(emu, x)
    assert_close_normwise(bp[0], oracle.characteristic_poly(T, 3), 1e-11)


# ------------------------------------------------------------------ K1 radix-8 path (k1_r8.cuh) under the fiber emulator
def emu_r8(lib, x, nblk=1, R=None):
    """x: [T, N, D].  Runs k1e_body for nblk CTAs of 64*R fibers (named barriers and the bulk hand-over emulated)."""
    T, N, D = x.shape
    R = R or lib.emu_k1r8_r(T)
    assert R > 0
    Tld = (T + 15) // 16 * 16
    ser = np.zeros((N, D, Tld))
    ser[:, :, :T] = x.transpose(1, 2, 0)
    bp, part = np.zeros((N, Tld)), np.zeros((nblk, Tld))
    assert lib.emu_k1r8(_p(ser), T, D, Tld, N, nblk, R, _p(bp), _p(part)) == 0
    assert np.all(bp[:, T:] == 0)
    return bp[:, :T], part[:, :T]


@pytest.mark.parametrize("T,D,N,nblk", [(3100, 2, 1, 1), (4000, 3, 2, 1), (4096, 1, 1, 1), (5000, 3, 3, 2), (5001, 2, 1, 1),
                                        (6000, 3, 1, 1), (8192, 3, 1, 1), (9999, 1, 1, 1), (10000, 3, 2, 1), (12000, 2, 1, 1)])
def test_radix8_fft_kernel_body_matches_tidynamics_restatement(emu, T, D, N, nblk):
    x = np.random.default_rng(T + D).standard_normal((T, N, D))
    bp, part = emu_r8(emu, x, nblk)
    for a in range(N):
        assert_close_normwise(bp[a], oracle.tidynamics_acf(x[:, a, :]), 1e-12, f"T={T} atom {a}")
    np.testing.assert_allclose(part.sum(axis=0), bp.sum(axis=0), rtol=1e-13, atol=1e-13)


def test_radix8_path_selection(emu):
    for T, want in [(1, 0), (3000, 0), (3071, 4), (4096, 4), (4097, 5), (5000, 5), (5121, 6), (6145, 8), (8192, 8),
                    (8193, 10), (10000, 10), (10240, 10), (10241, 12), (12288, 12), (12289, 0)]:
        assert emu.emu_k1r8_r(T) == want, T


@pytest.mark.parametrize("R", [4, 5, 6, 8, 10, 12])
def test_radix8_plan_thread_maps(emu, R):
    """Every P4 butterfly and every P3 / P2 duty is taken exactly once per residue; partners sit on lanes l, l ^ 16;
    a quarter warp's P4 butterflies have distinct k3 digits (conflict-free loads); groups own whole 512-point blocks."""
    NT = 64 * R
    mp = (ctypes.c_uint * (2 * NT))()
    assert emu.emu_k1r8_map(512 * R * 2, R, mp) == NT
    H = 512 * R
    for r in range(2):
        m = np.array(mp[r * NT:(r + 1) * NT], dtype=np.uint64)
        b4 = (m & 0x3FF).astype(int)
        blk3 = ((m >> 10) & 0x7F).astype(int)
        k2b = ((m >> 17) & 0xF).astype(int)
        grp = ((m >> 21) & 0xF).astype(int)
        gw = ((m >> 25) & 0xF).astype(int)
        assert sorted(b4) == list(range(NT))
        tid = np.arange(NT)
        assert sorted(blk3 * 8 + (tid & 7)) == list(range(NT))          # P3 butterflies
        assert sorted(k2b * 64 + (tid & 63)) == list(range(NT))        # P2 butterflies
        g0 = (b4 >> 6) + R * ((b4 >> 3) & 7) + 8 * R * (b4 & 7)
        for t in range(NT):
            p = t ^ 16
            ga, gb = g0[t] + 64 * R * 1, g0[p] + 64 * R * 6
            want = (H - ga) % H if r == 0 else H - 1 - ga
            selfp = bool(m[t] >> np.uint64(30))
            if selfp:
                assert r == 0 and b4[t] in (0, 4) and b4[p] in (0, 4)
            else:
                assert gb == want, (r, t)
        for q in range(NT // 8):
            assert len(set(b4[8 * q:8 * q + 8] & 7)) == 8
        for g in set(grp):
            th = tid[grp == g]
            assert len(th) == 32 * gw[th[0]] and th.max() - th.min() + 1 == len(th) and th.min() % 64 == 0
            ks = set(k2b[th])
            # P3 and P4 work of the group stays inside the 512-point blocks its P2 butterflies write
            assert set(blk3[th] >> 3) == ks and set(b4[th] >> 6) == ks
        # P3 -> P4 is warp-local
        for w in range(NT // 32):
            sl = slice(32 * w, 32 * w + 32)
            assert set(blk3[sl]) == set(b4[sl] >> 3)


@pytest.mark.parametrize("T,R1,var", [(10000, 20, 1), (10000, 20, 2), (9999, 20, 4), (10000, 20, 6), (5001, 10, 2),
                                      (5000, 10, 4), (4999, 10, 6), (10000, 20, 8), (9999, 20, 12), (5001, 10, 12),
                                      (9999, 20, 52)])
def test_fast_fft_kernel_variants(emu, T, R1, var):
    """k1_fast.cuh VAR bits: token-ordered loads (1), staged bulk output (2), bulk series prefetch (4), deferred
    P2 -> P3 twiddles (8), output in chunks of 10 (16), computed 1 / (L (T - k)) (32)."""
    N, D, nblk = 3, 3, 2
    x = np.random.default_rng(T + var).standard_normal((T, N, D))
    Tld = (T + 15) // 16 * 16
    ser = np.zeros((N, D, Tld))
    ser[:, :, :T] = x.transpose(1, 2, 0)
    bp, part = np.zeros((N, Tld)), np.zeros((nblk, Tld))
    assert emu.emu_k1fast_var(_p(ser), T, D, Tld, N, nblk, R1, var, _p(bp), _p(part)) == 0
    assert np.all(bp[:, T:] == 0)
    for a in range(N):
        assert_close_normwise(bp[a, :T], oracle.tidynamics_acf(x[:, a, :]), 1e-12, f"T={T} atom {a}")
    np.testing.assert_allclose(part.sum(axis=0)[:T], bp.sum(axis=0)[:T], rtol=1e-13, atol=1e-13)


# ------------------------------------------------------------------ boundary lengths of every fast-path instantiation
def _boundary_lengths(step, radices, lo_ok):
    """For H = step * R: the longest T (= 2H), the odd one below it, and the shortest T the selection rule sends to R."""
    out = []
    for R in radices:
        H = step * R
        out += [(2 * H, R), (2 * H - 1, R)]
        t_min = max(lo_ok, 2 * ((3 * H - 3 + 3) // 4) - 1)      # 3 H <= 4 nh + 3  ->  nh >= (3 H - 3) / 4
        out.append((t_min + 2, R))
    return out


@pytest.mark.parametrize("T,R1", _boundary_lengths(256, [4, 6, 8, 10, 12, 16, 20], 1535))
def test_three_pass_kernel_at_the_ends_of_every_length_range(emu, T, R1):
    if emu.emu_k1fast_r1(T) != R1:
        pytest.skip(f"T={T} is served by R1={emu.emu_k1fast_r1(T)}")
    x = np.random.default_rng(T).standard_normal((T, 1, 2))
    bp, part = emu_fast(emu, x, 1, R1)
    assert_close_normwise(bp[0], oracle.tidynamics_acf(x[:, 0, :]), 1e-12, f"T={T} R1={R1}")
    np.testing.assert_array_equal(part[0], bp[0])


@pytest.mark.parametrize("T,R", _boundary_lengths(512, [4, 5, 6, 8, 10, 12], 3071))
def test_radix8_kernel_at_the_ends_of_every_length_range(emu, T, R):
    if emu.emu_k1r8_r(T) != R:
        pytest.skip(f"T={T} is served by R={emu.emu_k1r8_r(T)}")
    x = np.random.default_rng(T).standard_normal((T, 1, 2))
    bp, part = emu_r8(emu, x, 1, R)
    assert_close_normwise(bp[0], oracle.tidynamics_acf(x[:, 0, :]), 1e-12, f"T={T} R={R}")
    np.testing.assert_allclose(part[0], bp[0], rtol=0, atol=0)

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


def emu_win(lib, x, mode, nwarps=4, f32=0, nsplit=1):
    T, D = x.shape
    Tld = (T + 15) // 16 * 16
    ser = np.zeros((D, Tld))
    ser[:, :T] = x.T
    res = np.zeros(T)
    assert lib.emu_windowed(_p(ser), T, D, Tld, mode, nwarps, nsplit, f32, _p(res)) == 0
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


@pytest.mark.parametrize("T,nwarps,nsplit", [(100, 1, 2), (333, 2, 3), (1000, 4, 2), (1000, 2, 7), (2000, 16, 4), (17, 1, 4)])
def test_windowed_particle_split_over_ctas(emu, T, nwarps, nsplit):
    """A particle's lag-block pairs dealt to nsplit CTAs (few-particle workloads): every lag is finished by exactly one
    of them, and the values are those of the one-CTA run bit for bit."""
    x = np.random.default_rng(T + nsplit).standard_normal((T, 2)) + 0.5
    for mode in (0, 1):
        one = emu_win(emu, x, mode, nwarps)
        split = emu_win(emu, x, mode, nwarps, nsplit=nsplit)
        assert np.array_equal(one, split)


def test_fp32_mode_tolerance(emu):
    x = np.random.default_rng(5).standard_normal((2000, 3))
    assert_close_normwise(emu_fft(emu, x, f32=1), oracle.tidynamics_acf(x), 1e-5, "fp32 fft")
    prod = np.array([np.sum(x[: 2000 - k] * x[k:]) for k in range(2000)])
    assert_close_normwise(emu_win(emu, x, 0, f32=1), prod, 1e-5, "fp32 windowed")


# ------------------------------------------------------------------ K1 fast path under the fiber emulator
def emu_fast(lib, x, nblk=1, R1=None, f32=0):
    """x: [T, N, D].  Runs k1f_body (k1_fast.cuh) for nblk CTAs of 16*R1 fibers."""
    T, N, D = x.shape
    R1 = R1 or lib.emu_k1fast_r1(T)
    assert R1 > 0
    Tld = (T + 15) // 16 * 16
    ser = np.zeros((N, D, Tld))
    ser[:, :, :T] = x.transpose(1, 2, 0)
    bp, part = np.zeros((N, Tld)), np.zeros((nblk, Tld))
    assert lib.emu_k1fast(_p(ser), T, D, Tld, N, nblk, R1, f32, _p(bp), _p(part)) == 0
    assert np.all(bp[:, T:] == 0)
    return bp[:, :T], part[:, :T]


@pytest.mark.parametrize("T,D,N,nblk", [(1600, 2, 1, 1), (2000, 3, 2, 1), (2047, 1, 1, 1), (3000, 3, 2, 2), (4000, 1, 1, 1),
                                        (5000, 3, 3, 2), (5001, 2, 1, 1), (6000, 3, 1, 1), (8192, 3, 1, 1), (10000, 3, 2, 1), (11000, 3, 2, 1), (12288, 1, 1, 1)])
def test_fast_fft_kernel_body_matches_tidynamics_restatement(emu, T, D, N, nblk):
    x = np.random.default_rng(T + D).standard_normal((T, N, D))
    bp, part = emu_fast(emu, x, nblk)
    for a in range(N):
        assert_close_normwise(bp[a], oracle.tidynamics_acf(x[:, a, :]), 1e-12, f"T={T} atom {a}")
    np.testing.assert_allclose(part.sum(axis=0), bp.sum(axis=0), rtol=1e-13, atol=1e-13)


def test_fast_fft_path_selection(emu):
    """R1 = 0 (general kernel) for tiny or badly padded lengths, else the smallest even R1 with 256 R1 >= ceil(T/2)."""
    for T, want in [(1, 0), (700, 0), (1533, 0), (1535, 4), (2048, 4), (2049, 0), (2400, 6), (4096, 8), (5000, 10),
                    (5001, 10), (6144, 12), (8192, 16), (10000, 20), (10240, 20), (10241, 24), (12288, 24), (12289, 0), (20000, 0)]:
        assert emu.emu_k1fast_r1(T) == want, T


def test_fast_fft_ramp_is_exact_enough(emu):
    """Ramp data (the reference's step trajectory) has a 1e7 dynamic range; the closed form is the reference's own."""
    T = 5001
    x = np.repeat(np.arange(T, dtype=np.float64)[:, None, None], 3, axis=2)
    bp, _ = emu_fast(emu, x)
    assert_close_normwise(bp[0], oracle.characteristic_poly(T, 3), 1e-11)


def test_fast_fft_kernel_body_fp32_mode(emu):
    """The same body instantiated for float (FP32 mode: float series in HBM, float arithmetic, double rows and particle
    sums), every R1, against the FP64 restatement at the stated 1e-5 (normwise); the bulk series copy of an odd number
    of complex floats is rounded up to 16 bytes (T = 4999: nh = 2500, T = 9998: nh = 4999)."""
    for T, D, N, nblk in [(1600, 3, 1, 1), (3000, 2, 2, 2), (4000, 3, 1, 1), (4999, 3, 2, 1), (6000, 1, 1, 1), (8192, 3, 1, 1),
                          (9998, 3, 3, 2), (12000, 2, 1, 1)]:
        x = np.random.default_rng(T).standard_normal((T, N, D)).astype(np.float32).astype(np.float64)
        bp, part = emu_fast(emu, x, nblk, f32=1)
        for a in range(N):
            ref = oracle.tidynamics_acf(x[:, a, :])
            # the float rounding floor of the un-normalised correlation (~0.04 eps32 sum x^2, the same at every lag) is
            # divided by the number of origins T - k: the stated 1e-5 holds for every lag with at least 8 origins, the
            # last 8 lags are within 1e-4 (DESIGN.md "FP32 mode")
            assert_close_normwise(bp[a][:-8], ref[:-8], 1e-5, f"fp32 T={T} atom {a}")
            assert_close_normwise(bp[a], ref, 1e-4, f"fp32 T={T} atom {a}, last lags")
        np.testing.assert_allclose(part.sum(axis=0), bp.sum(axis=0), rtol=1e-12, atol=1e-12)


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


@pytest.mark.parametrize("T,R1", _boundary_lengths(256, [4, 6, 8, 10, 12, 16, 20, 24], 1535))
def test_three_pass_kernel_at_the_ends_of_every_length_range(emu, T, R1):
    if emu.emu_k1fast_r1(T) != R1:
        pytest.skip(f"T={T} is served by R1={emu.emu_k1fast_r1(T)}")
    x = np.random.default_rng(T).standard_normal((T, 1, 2))
    bp, part = emu_fast(emu, x, 1, R1)
    assert_close_normwise(bp[0], oracle.tidynamics_acf(x[:, 0, :]), 1e-12, f"T={T} R1={R1}")
    np.testing.assert_array_equal(part[0], bp[0])


@pytest.mark.parametrize("T,D,N,nblk", [(10000, 3, 1, 1), (10000, 3, 5, 2), (9999, 2, 3, 1), (8192, 1, 4, 2), (7000, 3, 2, 3),
                                        (8000, 2, 3, 1), (10240, 1, 1, 1), (6200, 3, 2, 1)])
def test_tensor_memory_output_stage_gives_the_bits_of_the_global_one(emu, T, D, N, nblk):
    """k1f_body<..., TMEM = true>: parked V_0, particle sums and the 1/(L(T-k)) table in each thread's own tensor-memory
    columns instead of global memory.  Same arithmetic in the same order: rows and per-CTA particle sums bit for bit, for
    every D, several particles per CTA, CTAs without work, odd T (a last complex value with an empty imaginary lag) and
    series that fill H exactly."""
    x = np.random.default_rng(T + D + N).standard_normal((T, N, D))
    R1 = emu.emu_k1fast_r1(T)
    assert R1 in (16, 20)
    ref_bp, ref_part = emu_fast(emu, x, nblk, R1)
    bp, part = emu_fast(emu, x, nblk, R1, f32=2)
    assert np.array_equal(bp, ref_bp)
    assert np.array_equal(part, ref_part)
    assert_close_normwise(bp[0], oracle.tidynamics_acf(x[:, 0, :]), 1e-12, f"T={T}")

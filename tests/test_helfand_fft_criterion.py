"""The numerical design of the FFT route of the Helfand MSD, checked on the CPU (numpy stands in for K1 / K5):
S1 - 2 S2 is good to C eps sum(g^2) / MSD[k]; every lag whose un-normalised MSD is below thr * sum(g^2), thr =
(100 + T/100) eps / 2e-11 (csrc/ta_b200.cu ta_helfand_fft), is handed to the exact sum.  The test asserts that on every
trajectory family (a) the lags the rule keeps are within the 1e-10 bar with room to spare, (b) the error constant stays
far below the assumed 100, (c) the rule hands over only a small part of the lags, except for series that barely move."""
import numpy as np
import pytest

EPS = 2.0 ** -53


def family(kind, T, N, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(T, dtype=np.float64)[:, None]
    if kind == "white":
        return rng.standard_normal((T, N)) * 10.0 * rng.standard_normal((T, N))
    if kind == "walk":
        v = rng.standard_normal((T, N))
        for i in range(1, T):
            v[i] = 0.99 * v[i - 1] + np.sqrt(1 - 0.99 ** 2) * v[i]
        return 12.0 * v * (np.cumsum(v, axis=0) + 20.0 * rng.random((1, N)))
    if kind == "smooth":
        ph = rng.uniform(0, 6.28, (1, N))
        return (np.sin(2e-3 * t + ph) + 2.0) * (np.cos(1.3e-3 * t + 2 * ph) + 3.0)
    if kind == "ramp":                                   # the reference's step trajectory: m v x = 16 t t^2/2
        return np.repeat(16.0 * t * (t * t / 2), N, axis=1)
    if kind == "still":                                  # barely moving moments
        return 5.0 + 1e-9 * rng.standard_normal((T, N))
    raise ValueError(kind)


def exact_sums(g):
    """sum_i (g[i] - g[i+k])^2 accumulated in extended precision (the yardstick must be better than what it measures)."""
    T = g.shape[0]
    out = np.zeros_like(g)
    for k in range(1, T):
        d = (g[:-k] - g[k:]).astype(np.longdouble)
        out[k] = (d * d).sum(axis=0).astype(np.float64)
    return out


def k5_prefix_sums(q):
    """P[j] = sum_{i<j} q[i] in the association K5 uses: sequential inside a thread's segment (1,024 threads), a
    Hillis-Steele scan over the segment totals."""
    T, N = q.shape
    seg = -(-T // 1024)
    nseg = -(-T // seg)
    pad = np.zeros((nseg * seg, N))
    pad[:T] = q
    inner = np.cumsum(pad.reshape(nseg, seg, N), axis=1)
    tot = inner[:, -1, :].copy()
    o = 1
    while o < nseg:
        tot[o:] = tot[o:] + tot[:-o]
        o *= 2
    base = np.concatenate([np.zeros((1, N)), tot[:-1]])
    incl = (inner + base[:, None, :]).reshape(nseg * seg, N)[:T]
    return np.concatenate([np.zeros((1, N)), incl])


def fft_route(g):
    """un-normalised S1[k] - 2 S2[k] the way K1 + K5 form it (FFT autocorrelation, prefix sums)."""
    T = g.shape[0]
    L = 1 << int(np.ceil(np.log2(2 * T)))
    F = np.fft.rfft(g, n=L, axis=0)
    s2 = np.fft.irfft(F * np.conj(F), n=L, axis=0)[:T]
    P = k5_prefix_sums(g * g)
    k = np.arange(T)
    s1 = P[T - k] + (P[T] - P[k])
    return s1 - 2.0 * s2, P[T]


@pytest.mark.parametrize("kind,T", [("white", 600), ("white", 3000), ("walk", 3000), ("smooth", 3000), ("smooth", 700),
                                    ("ramp", 2000), ("still", 500)])
def test_threshold_rule_keeps_only_lags_that_meet_the_bar(kind, T):
    N = 6
    g = family(kind, T, N, seed=T)
    ex = exact_sums(g)
    approx, tot = fft_route(g)
    thr = (100.0 + T / 100.0) * EPS / 2e-11
    keep = approx >= thr * tot[None, :]
    keep[0] = False                                      # lag 0 is defined as 0
    with np.errstate(over="ignore", divide="ignore"):
        rel = np.abs(approx - ex) / np.maximum(ex, 1e-300)
    assert np.all(rel[keep] < 2e-11), (kind, rel[keep].max())
    C = np.abs(approx - ex)[1:] / (EPS * tot[None, :])
    assert C.max() < 50, (kind, C.max())
    handed_over = 1.0 - keep[1:].mean()
    if kind == "still":
        assert handed_over > 0.9                         # -> more than 2 % marked: the direct kernel takes the shard
    elif kind in ("white", "walk"):
        assert handed_over < 0.01
    else:
        assert handed_over < 0.2

"""Host logic of the parallel COI-entry march: the closed form of the reference's running sums
(`t += dt`, `ma += dr * n * dt`; orbit_intersect.py:108-121,170) must give, for any step k, the
very bits of the serial loop -- through binade changes, zero crossings, ties and stalls.
Pure host code of libvsb200.so (no GPU)."""
import ctypes as C

import numpy as np
import pytest

from volatilespace_b200._lib import VsbError, check, load, ptr


def closed(x0, c, kmax, ks, bounded=False, bound=0.0):
    L = load()
    ks = np.ascontiguousarray(ks, dtype=np.int64)
    xs = np.empty(len(ks))
    K, ns = C.c_int64(), C.c_int64()
    check(L.vsb_host_running_sum(float(x0), float(c), int(kmax), int(bounded), float(bound),
                                 len(ks), ptr(ks), ptr(xs), C.addressof(K), C.addressof(ns)))
    return xs, K.value, ns.value


def serial(x0, c, kmax):
    arr = np.full(kmax + 1, c, dtype=np.float64)
    arr[0] = x0
    return np.cumsum(arr)            # numpy's cumsum is the plain left-to-right loop


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def test_random_sums_match_the_loop_bit_for_bit():
    rng = np.random.default_rng(1)
    for trial in range(400):
        x0 = rng.choice([0.0, rng.uniform(-10, 10), rng.uniform(0, 7),
                         10 ** rng.uniform(-8, 4) * rng.choice([-1, 1])])
        c = 10 ** rng.uniform(-6, 1) * rng.choice([-1, 1])
        if trial % 5 == 0:           # few-bit increments: exact ties on every step of a binade
            c = float(np.ldexp(float(rng.integers(1, 64)), -int(rng.integers(0, 12)))) * rng.choice([-1, 1])
        kmax = int(rng.integers(1, 60000))
        xs, K, ns = closed(x0, c, kmax, np.arange(kmax + 1))
        assert K == kmax and ns <= 288
        assert np.array_equal(bits(xs), bits(serial(x0, c, kmax))), (x0, c, kmax)


@pytest.mark.parametrize("x0,c,K", [(3.7, -2.13e-6, 12_000_000), (0.0, 17.3331, 20_000_000),
                                    (5.1, 1.9e-7, 20_000_000), (-2.0, 3.3e-7, 16_000_000)])
def test_long_marches(x0, c, K):
    want = serial(x0, c, K)
    ks = np.concatenate([np.arange(0, K + 1, 997), [K]])
    xs, _, ns = closed(x0, c, K, ks)
    assert np.array_equal(bits(xs), bits(want[ks]))
    assert ns < 120                  # O(log K) segments


def test_trip_count_of_the_time_loop():
    """`while t < t_end: t += dt` (orbit_intersect.py:140,170)."""
    rng = np.random.default_rng(2)
    for _ in range(300):
        period = 10 ** rng.uniform(1, 6)
        steps = int(rng.integers(300, 40000))
        dt = period / steps
        t = serial(0.0, dt, steps + 3)
        k = int(np.argmax(~(t < period)))
        xs, K, _ = closed(0.0, dt, 1 << 60, [0, 1, k - 1, k], bounded=True, bound=period)
        assert K == k
        assert np.array_equal(bits(xs), bits(t[[0, 1, k - 1, k]]))
    assert closed(0.0, 1.0, 1 << 60, [0], bounded=True, bound=float("nan"))[1] == 0
    assert closed(0.0, 1.0, 1 << 60, [0], bounded=True, bound=-5.0)[1] == 0


def test_stalls():
    # |c| below half an ulp: the sum stays put (fine for a mean anomaly ...)
    xs, K, _ = closed(1.0, 1e-17, 1000, [1, 500, 1000])
    assert np.all(xs == 1.0)
    # ... but a time that stops growing below t_end never ends the reference's loop
    with pytest.raises(VsbError):
        closed(1.0, 1e-17, 1 << 60, [0], bounded=True, bound=2.0)
    with pytest.raises(VsbError):
        closed(0.0, 0.5, 10, [11])

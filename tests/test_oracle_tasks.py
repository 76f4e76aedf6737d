"""CPU tests of the oracle's restatement of the offline TimeWindow tasks (SPEC Y, oracle/tasks.c) against
independent float64 numpy statements of the same reference loops and against known answers."""
import numpy as np
import pytest

from sigdigger_b200 import synth


def _noise(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return (scale * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)


def test_delayed_conj_matches_float64_statement(oracle):
    x = _noise(5000, 1)
    for delay in (1, 7, 4096, 6000):
        y = oracle.delayed_conj(x, delay)
        xd = x.astype(np.complex128)
        ref = np.zeros_like(xd)
        if delay < len(x):
            prev = xd[:-delay]
            ref[delay:] = xd[delay:] * np.conj(prev) / (np.abs(prev) + 1e-3)
        assert np.all(y[:min(delay, len(x))] == 0)
        assert np.allclose(y, ref, rtol=2e-6, atol=1e-6)


def test_histogram_feed_spaces(oracle):
    x = _noise(3000, 2)
    xd = x.astype(np.complex128)
    a = oracle.histogram_feed(x, "amplitude")
    p = oracle.histogram_feed(x, "phase")
    f = oracle.histogram_feed(x, "frequency")
    assert len(a) == len(p) == len(x) and len(f) == len(x) - 1
    assert np.allclose(a, np.abs(xd), rtol=1e-6)
    assert np.allclose(p, np.angle(xd), atol=2e-6)
    assert np.allclose(f, np.angle(xd[1:] * np.conj(xd[:-1])), atol=5e-6)


def _manual_ref(x, space, symbol_count, symbol_sync):
    """literal float64 walk of WaveSampler::sampleManual"""
    n = len(x)
    delta = n / symbol_count
    so = symbol_sync / delta
    out = []
    prev = 0j
    xd = x.astype(np.complex128)
    for p in range(int(symbol_count)):
        start = (p - so) * delta + symbol_sync
        end = start + delta
        i0, i1 = int(np.floor(start)), int(np.ceil(end))
        t0, t1 = 1 - (start - i0), 1 - (i1 - end)
        avg = 0j
        for i in range(i0, i1 + 1):
            v = 0j
            if 0 <= i < n:
                v = t0 * xd[i] if i == i0 else (t1 * xd[i] if i == i1 else xd[i])
            avg += v * np.conj(v) if space == "amplitude" else v * np.conj(prev)
            prev = v
        out.append(np.sqrt(avg.real / delta) if space == "amplitude" else avg / delta)
    return np.asarray(out)


@pytest.mark.parametrize("space", ["amplitude", "phase", "frequency"])
def test_manual_sampler_matches_float64_walk(oracle, space):
    x = _noise(2000, 3)
    for count, sync in ((100.0, 0), (173.3, 5), (2000.0, 0), (61.5, 13)):
        got = oracle.sample_manual(x, space, count, sync)
        ref = _manual_ref(x, space, count, sync)
        assert len(got) == int(count)
        assert np.allclose(got, ref, rtol=2e-5, atol=2e-5)


def test_manual_sampler_recovers_ask_levels(oracle):
    sps, nsym = 16, 200
    rng = np.random.default_rng(4)
    bits = rng.integers(0, 2, nsym)
    x = np.repeat(0.2 + 0.8 * bits, sps).astype(np.complex64)
    got = oracle.sample_manual(x, "amplitude", float(nsym), 0)
    # the box-car runs one sample into the next symbol with weight ~1 (edge weights 1 - frac), so allow for it
    dec = (got.real > 0.6 * np.sqrt((sps + 1) / sps)).astype(int)
    assert np.mean(dec == bits) > 0.97


def test_zero_crossing_run_lengths(oracle):
    sps = 20
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2, 700)
    lvl = np.repeat(2.0 * bits - 1.0, sps)
    # PHASE space with zeroCrossingAngle = e^{-0.2i}: var = arg(x) - 0.2 = +0.3 for a "1" and -0.7 for a "0"
    x = np.exp(0.5j * lvl).astype(np.complex64)
    sym, total = oracle.sample_zero_crossing(x, "phase", 1.0 / sps, zc_angle=np.exp(-0.2j))
    assert total == len(sym)
    # A run is reported when it ENDS, with the polarity of the run that FOLLOWS it (block[i] = var > 0 at the
    # crossing), so a two-level signal comes out inverted; and inside the last 4096-block every sample counts as
    # a crossing (`last` is per work() call), so that block yields nothing at 1/20 symbols per sample.
    n_last = len(x) - (len(x) - 1) // 4096 * 4096
    expect = (len(x) - n_last) // sps
    assert abs(total - expect) <= 4
    runs = 1 - bits
    m = min(total, len(runs)) - 3
    best = max(np.mean(sym[k:m] == runs[:m - k]) for k in range(0, 3))
    assert best > 0.97


def test_zero_crossing_blocks_cap_and_amplitude(oracle):
    n = 3 * 4096 + 100
    t = np.arange(n)
    x = (0.5 + 0.45 * np.sign(np.sin(2 * np.pi * t / 64.0))).astype(np.complex64)
    sym, total = oracle.sample_zero_crossing(x, "amplitude", 1.0 / 32, amplitude=True, threshold=0.5 + 0j)
    assert total > 300 and set(np.unique(sym)) <= {0, 1}
    # bnor = 1, constant positive variable, 9000 samples = blocks of 4096, 4096, 808.  Every block re-seeds
    # prevVar = -1, so its first sample is a crossing: block 0 reports 0 samples, block 1 the 4096 before it;
    # in the last block every sample is a crossing: 4096 for the first, then 1 each until the per-block cap of
    # 4096 symbols is hit -> 4096 + 4096.
    y = np.ones(9000, np.complex64)
    sym, total = oracle.sample_zero_crossing(y, "amplitude", 1.0, amplitude=True, threshold=0.5 + 0j)
    assert total == 8192 and np.all(sym == 1)
    s2, t2 = oracle.sample_zero_crossing(y, "amplitude", 1.0, amplitude=True, threshold=0.5 + 0j, cap=10)
    assert t2 == total and len(s2) == 10 and np.array_equal(s2, sym[:10])


def _snr_literal(history, bps, sigma, alpha):
    """one SNREstimator::feed (Misc/SNREstimator.cpp:133-158 -> iterate :80-120 -> recalculateModel :30-78),
    float32 statement by statement, numpy's exp"""
    f = np.float32
    L = len(history)
    n_int = 1 << bps
    hx = f(1) / f(L)
    mx = max(int(history.max()), 1)
    ht = (history.astype(np.float32) / f(mx)).astype(np.float32)
    sigma = f(sigma)
    sigma2 = f(sigma * sigma)
    x = (np.arange(L).astype(np.float32) * hx).astype(np.float32)
    x = np.where(x >= f(.5), x - f(1), x).astype(np.float32)
    g = np.exp((-x * x / sigma2).astype(np.float32)).astype(np.float32)
    intlen = f(1) / f(n_int)
    start = f(.5) * intlen
    hi = np.zeros(L, np.float32)
    idx = np.arange(L)
    for j in range(n_int):
        skip = f(start + f(j) * intlen)
        t = f(f(1) - f(skip - np.floor(skip)))
        skipint = int(np.floor(f(f(L) * skip)))
        i1 = (L + idx - skipint) % L
        i2 = (L + i1 - 1) % L
        hi = (hi + t * g[i1]).astype(np.float32)
        hi = (hi + f(f(1) - t) * g[i2]).astype(np.float32)
    if hi.max() > 0:
        hi = (hi / hi.max()).astype(np.float32)
    sinv = f(1) / sigma
    s3 = f(f(sinv * sinv) * sinv)
    delta = f(0)
    for i in range(L):
        term = f(0)
        for j in range(n_int):
            skip = f(start + f(j) * intlen)
            d = f(x[i] - skip)
            term = f(term + f(d * d))
        term = f(term * f(f(hi[i] - ht[i]) / s3))
        delta = f(delta + term)
    delta = f(delta / f(L))
    return f(sigma + f(-f(alpha) * delta)), hi


def _histogram(bps, length, sigma, n=200000, seed=0):
    rng = np.random.default_rng(seed)
    k = 1 << bps
    centres = (rng.integers(0, k, n) + 0.5) / k
    v = (centres + sigma * rng.standard_normal(n)) % 1.0
    return np.bincount((v * length).astype(int) % length, minlength=length).astype(np.uint32)


def test_snr_estimator_matches_literal_and_converges(oracle):
    for bps, length in ((1, 256), (2, 256), (3, 400)):
        h = _histogram(bps, length, 0.03, seed=bps)
        e = oracle.SnrEstimator(bps, length, alpha=1.0)
        sigma = 1.0 / 8
        for it in range(4):
            want, hi = _snr_literal(h, bps, sigma, 1.0)
            e.feed(h)
            assert abs(e.sigma - want) <= 2e-5 * abs(want), (bps, it)
            assert np.abs(e.model() - hi).max() < 2e-5
            sigma = e.sigma                                   # keep the literal on the oracle's trajectory
        assert e.snr == pytest.approx(1.0 / ((1 << bps) * e.sigma), rel=1e-6)
        e.close()
    # the fit moves sigma towards the histogram's true width (gradient descent on the model error)
    h = _histogram(2, 256, 0.02, seed=9)
    e = oracle.SnrEstimator(2, 256, alpha=0.05)
    s0 = e.sigma
    for _ in range(300):
        e.feed(h)
    assert abs(e.sigma - 0.02 * np.sqrt(2)) < abs(s0 - 0.02 * np.sqrt(2))
    assert 0.02 < e.sigma < s0          # slowly: the step carries a factor sigma^3 (`/ sigma3inv`, :110)
    e.close()


def test_carrier_detect_finds_tone(oracle):
    rng = np.random.default_rng(6)
    for n, f in ((1000, 0.0371), (4096, -0.21), (50000, 0.3003)):
        x = synth.awgn(n, 0.05, rng) + 0.7 * np.exp(2j * np.pi * f * np.arange(n))
        w = oracle.carrier_detect(x.astype(np.complex64), 0.002, 0.0)
        assert abs(w / (2 * np.pi) - f) < 1.5 / n + 1e-4
    # DC notch: a strong DC term is ignored, the weaker tone wins
    n = 8192
    x = (1.0 + 0.3 * np.exp(2j * np.pi * 0.125 * np.arange(n))).astype(np.complex64)
    assert abs(oracle.carrier_detect(x, 0.002, 0.0)) < 1e-3
    assert abs(oracle.carrier_detect(x, 0.002, 0.05) / (2 * np.pi) - 0.125) < 1e-3

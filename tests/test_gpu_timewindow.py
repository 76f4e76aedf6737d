"""-m gpu: the offline TimeWindow tasks that are fully specified in-repo (SPEC Y) through the C-ABI against the
oracle's loop-by-loop restatement (oracle/tasks.c): bit-exact, batched, including the reference's block quirks."""
import numpy as np
import pytest

from sigdigger_b200 import synth

pytestmark = pytest.mark.gpu


def _noise(shape, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return (scale * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))).astype(np.complex64)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_delayed_conj_bit_exact(sdb, oracle):
    x = _noise((5, 20000), 1)
    x[2, 100:140] = 0                                  # |prev| = 0 -> 1 / 1e-3
    for delay in (1, 13, 4096, 19999, 30000):
        got = sdb.delayed_conj(x, delay)
        for b in range(x.shape[0]):
            assert np.array_equal(_bits(got[b]), _bits(oracle.delayed_conj(x[b], delay)))
    with pytest.raises(sdb.SdbError):
        sdb.delayed_conj(x, 0)


@pytest.mark.parametrize("space", ["amplitude", "phase", "frequency"])
def test_histogram_feed_bit_exact(sdb, oracle, space):
    x = _noise((4, 33333), 2)
    x[1, 5] = 0
    got = sdb.histogram_feed(x, space)
    assert got.shape == (4, 33333 - (space == "frequency"))
    for b in range(4):
        assert np.array_equal(_bits(got[b]), _bits(oracle.histogram_feed(x[b], space)))
    one = sdb.histogram_feed(x[0], space)
    assert np.array_equal(_bits(one), _bits(got[0]))


@pytest.mark.parametrize("space", ["amplitude", "phase", "frequency"])
def test_manual_sampler_bit_exact(sdb, oracle, space):
    x = _noise((3, 50000), 3)
    for count, sync in ((500.0, 0), (1733.3, 7), (50000.0, 0), (61.5, 130), (12345.678, 3)):
        got = sdb.sample_manual(x, space, count, sync)
        assert got.shape == (3, int(count))
        for b in range(3):
            ref = oracle.sample_manual(x[b], space, count, sync)
            assert np.array_equal(_bits(got[b]), _bits(ref)), (space, count, sync, b)
    with pytest.raises(sdb.SdbError):
        sdb.sample_manual(x, space, 0.5, 0)


def test_manual_sampler_then_decider(sdb, oracle):
    """sampler -> decider, as WaveSampler::work() does for MANUAL / GARDNER (Tasks/WaveSampler.cpp:316-317)"""
    sps, nsym = 8, 4000
    sig, syms = synth.psk_signal(sps * nsym, float(sps), order=4, seed=9)
    x = (sig * np.exp(0.1j)).astype(np.complex64)
    soft = sdb.sample_manual(x, "phase", float(nsym), 0)
    hard = sdb.decide(soft, "argument", 2, -np.pi, np.pi)
    ref_soft = oracle.sample_manual(x, "phase", float(nsym), 0)
    assert np.array_equal(_bits(soft), _bits(ref_soft))
    assert np.array_equal(hard, oracle.decide(ref_soft, "argument", 2, -np.pi, np.pi))
    m = sdb.decide(soft, "modulus", 3, 0.0, 2.0)
    assert np.array_equal(m, oracle.decide(ref_soft, "modulus", 3, 0.0, 2.0))
    assert hard.max() <= 3 and len(np.unique(hard)) >= 2


def _fsk(n, sps, seed, h=0.5):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, n // sps + 1)
    f = np.repeat(2.0 * bits - 1.0, sps)[:n] * (h / (2 * sps))
    return np.exp(2j * np.pi * np.cumsum(f)).astype(np.complex64)


@pytest.mark.parametrize("space,amplitude", [("amplitude", True), ("amplitude", False), ("phase", False),
                                             ("frequency", False)])
def test_zero_crossing_bit_exact(sdb, oracle, space, amplitude):
    n = 5 * 4096 + 777                                  # several work() blocks + a short last one
    xs = []
    for b in range(4):
        if space == "frequency":
            x = _fsk(n, 10 + b, 20 + b) + _noise(n, 30 + b, 0.02)
        elif space == "phase":
            lvl = np.repeat(2.0 * np.random.default_rng(40 + b).integers(0, 2, n // 16 + 1) - 1.0, 16)[:n]
            x = np.exp(0.6j * lvl) + _noise(n, 50 + b, 0.05)
        else:
            lvl = np.repeat(np.random.default_rng(60 + b).integers(0, 2, n // 12 + 1), 12)[:n]
            x = (0.2 + 0.8 * lvl) * np.exp(0.3j) + _noise(n, 70 + b, 0.03)
        xs.append(x.astype(np.complex64))
    x = np.stack(xs)
    x[3, 4096 * 2 + 5: 4096 * 2 + 9] = 0                # exact zeros: var == 0 is neither sign
    kw = dict(amplitude=amplitude, threshold=0.6 + 0.1j, zc_angle=np.exp(-0.3j))
    for bnor in (1.0 / 12, 0.37, 1.0, 2.5):
        got, cnt = sdb.sample_zero_crossing(x, space, bnor, **kw)
        for b in range(4):
            ref, total = oracle.sample_zero_crossing(x[b], space, bnor, **kw)
            assert int(cnt[b]) == total, (space, bnor, b)
            assert np.array_equal(got[b], ref), (space, bnor, b)
    # cap smaller than the output: the count still says how many there were
    got, cnt = sdb.sample_zero_crossing(x, space, 1.0, cap=100, **kw)
    ref, total = oracle.sample_zero_crossing(x[0], space, 1.0, cap=100, **kw)
    assert int(cnt[0]) == total > 100 and np.array_equal(got[0], ref) and len(got[0]) == 100


def test_zero_crossing_block_cap_known_answer(sdb):
    y = np.ones((2, 9000), np.complex64)
    got, cnt = sdb.sample_zero_crossing(y, "amplitude", 1.0, amplitude=True, threshold=0.5 + 0j)
    assert list(cnt) == [8192, 8192] and np.all(got[0] == 1)


@pytest.mark.parametrize("n", [1000, 4096, 65536, 200000])
def test_carrier_detect_bit_exact(sdb, oracle, n):
    rng = np.random.default_rng(n)
    fs = [0.0371, -0.21, 0.3003, 0.0004]
    x = np.stack([synth.awgn(n, 0.05, rng) + 0.7 * np.exp(2j * np.pi * f * np.arange(n)) for f in fs])
    x = x.astype(np.complex64)
    for rel, notch in ((0.002, 0.0), (0.01, 0.001), (0.05, 0.3)):
        got = sdb.carrier_detect(x, rel, notch)
        for b in range(len(fs)):
            ref = oracle.carrier_detect(x[b], rel, notch)
            assert np.float32(got[b]).view(np.uint32) == np.float32(ref).view(np.uint32), (n, rel, notch, b)
    got = sdb.carrier_detect(x, 0.002, 0.0)
    for b, f in enumerate(fs):
        assert abs(got[b] / (2 * np.pi) - f) < 1.5 / n + 1e-4
    with pytest.raises(sdb.SdbError):
        sdb.carrier_detect(np.zeros((1 << 20) + 1, np.complex64))

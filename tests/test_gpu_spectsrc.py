"""Inspector spectrum sources and baud estimators (SPEC U): the CUDA path through the C-ABI
against the oracle on the channel samples the engine itself produced (bit-exact), over several streams and feeds."""
import numpy as np
import pytest

from sigdigger_b200 import synth

pytestmark = pytest.mark.gpu

KINDS = ["psd", "cyclo", "fmspect", "timediff", "abstimediff", "exp_2", "exp_4", "exp_8", "fac"]


def _capture(S, n, sps_in, seed):
    xs = []
    for s in range(S):
        sig, _ = synth.psk_signal(n, sps_in, order=4, seed=seed + s)
        rng = np.random.default_rng(seed + s)
        xs.append(0.3 * synth.mix(sig, 0.1 + 1e-4 * s, 0.2 * s) + synth.awgn(n, 1e-3, rng))
    return np.stack(xs).astype(np.complex64)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ns", [256, 4096])
def test_spectrum_source_bit_exact(sdb, oracle, kind, ns):
    W, hops, S = 8192, 48, 3
    n = W // 2 * hops
    x = _capture(S, n, 64.0, seed=11)                     # 1/64 baud at the input rate
    e = sdb.Engine(n_streams=S, psd_size=0, st_window_size=W, max_feed=n)
    h = e.open_channel(2 * np.pi * 0.1, 2 * np.pi * 4 / 64.0, 1.0)   # decimation 16 -> 4 samples per symbol
    e.set_spectrum_source(h, kind, ns)
    e.commit()
    for lo, hi in ((0, n // 2), (n // 2, n)):              # two feeds: the frame is the tail of each feed's output
        e.feed(x[:, lo:hi])
        got, sizes = e.read_spectrum(h)
        for s in range(S):
            chan = e.read_channel(s, h)
            ref = oracle.spectsrc_frame(kind, ns, chan)
            if ref is None:
                assert sizes[s] == 0
                continue
            assert sizes[s] == len(ref) == (ns // 2 if kind == "fac" else ns)
            assert np.array_equal(got[s, :len(ref)].view(np.uint32), ref.view(np.uint32)), (kind, ns, s)


def test_short_feed_emits_nothing(sdb, oracle):
    W, S = 4096, 2
    n = W // 2 * 4
    x = _capture(S, n, 32.0, seed=5)
    e = sdb.Engine(n_streams=S, psd_size=0, st_window_size=W, max_feed=n)
    h = e.open_channel(2 * np.pi * 0.1, 2 * np.pi * 4 / 32.0, 1.0)
    e.set_spectrum_source(h, "psd", 4096)
    e.set_estimator(h, "baud-nonlinear")
    e.commit()
    e.feed(x)
    _, sizes = e.read_spectrum(h)
    vals, valid = e.read_estimate(h, "baud-nonlinear")
    assert not sizes.any() and not valid.any()


@pytest.mark.parametrize("ns", [1024, 4096])
def test_baud_estimators_bit_exact_and_right(sdb, oracle, ns):
    W, hops, S = 8192, 24, 4
    n = W // 2 * hops
    x = _capture(S, n, 64.0, seed=23)
    e = sdb.Engine(n_streams=S, psd_size=0, st_window_size=W, max_feed=n)
    h = e.open_channel(2 * np.pi * 0.1, 2 * np.pi * 8 / 64.0, 1.0)   # decimation 8 -> 8 samples per symbol
    e.set_spectrum_source(h, "abstimediff", ns)
    e.set_estimator(h, "baud-fac")
    e.set_estimator(h, "baud-nonlinear")
    e.commit()
    e.feed(x)
    fs_ch = e.channel_rate(h)
    true_baud = e.samp_rate / 64.0
    for est in ("baud-fac", "baud-nonlinear"):
        vals, valid = e.read_estimate(h, est)
        for s in range(S):
            ref = oracle.estimate_baud(est, ns, fs_ch, e.read_channel(s, h))
            assert (ref is not None) == bool(valid[s])
            assert ref is not None
            assert np.float32(ref).view(np.uint32) == vals[s].view(np.uint32), (est, s, ref, vals[s])
            # baud-fac resolves whole lags only (8 +- 1 samples per symbol here); the spectral line is sharp
            tol = 0.15 if est == "baud-fac" else 0.02
            assert abs(vals[s] - true_baud) / true_baud < tol, (est, vals[s], true_baud)


def test_registry_names_and_errors(sdb):
    L = sdb.load_library()
    assert [L.sdb_spectsrc_name(i).decode() for i in range(10)] == [
        "none", "psd", "cyclo", "fmspect", "timediff", "abstimediff", "exp_2", "exp_4", "exp_8", "fac"]
    assert [L.sdb_estimator_name(i).decode() for i in range(2)] == ["baud-fac", "baud-nonlinear"]
    assert L.sdb_spectsrc_name(10) is None
    e = sdb.Engine(n_streams=1, psd_size=0, st_window_size=4096, max_feed=8192)
    h = e.open_channel(1.0, 0.2, 1.0)
    assert L.sdb_engine_set_spectrum_source(e._h, h, 1, 1000) == -1      # not a power of two
    assert L.sdb_engine_set_spectrum_source(e._h, h, 42, 1024) == -1
    assert L.sdb_engine_set_spectrum_source(e._h, h + 1, 1, 1024) == -1   # wrong handle
    assert L.sdb_engine_set_estimator(e._h, h, 7, 1) == -1
    e.commit()
    assert L.sdb_engine_set_spectrum_source(e._h, h, 1, 1024) == -1      # after commit

"""-m gpu parity tests: the CUDA path through the C-ABI vs the CPU oracle on the same seeded input."""
import numpy as np
import pytest

import parity
from sigdigger_b200 import synth

pytestmark = pytest.mark.gpu


def _noise(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


# ------------------------------------------------------------------------------------------------
# main PSD
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,window", [(512, "hann"), (4096, "blackmann_harris"), (8192, "hann"),
                                      (16384, "none"), (32768, "hamming"), (65536, "blackmann_harris"),
                                      (65536, "flat_top"), (1 << 17, "hann")])
def test_psd_matches_oracle(sdb, oracle, N, window):
    frames = 3
    x = _noise(N * frames, seed=N % 97, scale=0.3)
    x += (0.5 * np.exp(2j * np.pi * 0.1234 * np.arange(N * frames))).astype(np.complex64)
    e = sdb.Engine(n_streams=1, psd_size=N, psd_window=window, max_feed=N * frames)
    e.commit()
    e.feed(x[None, :])
    got = e.read_psd()[0]
    ref = oracle.psd_frames(x, N, window)                 # SPEC transform: bit-identical
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), parity.psd_err_ratio(got, ref)
    # ... and within tolerance of the oracle's independent radix-2 transform
    parity.assert_psd_close(got, oracle.psd_frames(x, N, window, spec=False))
    # both sides must sit within float32 rounding of the float64 truth
    w = oracle.window(N, window).astype(np.float64)
    truth = np.abs(np.fft.fft(x.reshape(frames, N).astype(np.complex128) * w, axis=1)) ** 2 / N
    parity.assert_psd_close(got, truth)


def test_psd_multistream_and_chunks(sdb, oracle):
    N, S, frames = 8192, 5, 4
    x = np.stack([_noise(N * frames, seed=10 + s, scale=0.1 * (s + 1)) for s in range(S)])
    e = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=N * frames)
    e.commit()
    e.feed(x)
    got = e.read_psd()
    for s in range(S):
        assert np.array_equal(got[s].view(np.uint32), oracle.psd_frames(x[s], N, "hann").view(np.uint32))
    # feeding in two halves gives the same frames
    e2 = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=N * frames)
    e2.commit()
    e2.feed(x[:, :N * 2])
    a = e2.read_psd().copy()
    e2.feed(x[:, N * 2:])
    b = e2.read_psd().copy()
    assert np.array_equal(np.concatenate([a, b], axis=1), got)


@pytest.mark.parametrize("N", [16384, 32768, 65536])          # generic four-step, SPEC F.5 and F.4 row kernels
def test_psd_shift_db_epilogue(sdb, oracle, N):
    x = _noise(N * 2, seed=5, scale=0.2)
    e = sdb.Engine(n_streams=1, psd_size=N, psd_window="hann", max_feed=N * 2, flags=sdb.FLAG_PSD_SHIFT_DB)
    e.commit()
    e.feed(x[None, :])
    got = e.read_psd()[0]
    ref = oracle.psd_frames(x, N, "hann")
    for f in range(2):
        r = ref[f].copy()
        oracle.lib().sdo_psd_shift_db(oracle.ptr(r), N)
        assert np.max(np.abs(got[f] - r)) < 2e-3   # dB; bins at the -80 dB floor amplify float noise


def test_psd_known_answers(sdb):
    N = 65536
    # impulse -> flat 1/N ; tone on a bin -> single bin N*|a|^2
    x = np.zeros((1, N * 2), np.complex64)
    x[0, 0] = 1.0
    k = 1234
    x[0, N:] = 0.5 * np.exp(2j * np.pi * k * np.arange(N) / N)
    e = sdb.Engine(n_streams=1, psd_size=N, psd_window="none", max_feed=N * 2)
    e.commit()
    e.feed(x)
    p = e.read_psd()[0]
    assert np.allclose(p[0], 1.0 / N, rtol=1e-5)
    assert abs(p[1][k] - N * 0.25) / (N * 0.25) < 1e-5
    assert np.delete(p[1], k).max() < 1e-8 * N
    # Parseval: sum psd = sum |x|^2
    assert abs(p[1].sum() - 0.25 * N) / (0.25 * N) < 1e-5


# ------------------------------------------------------------------------------------------------
# channeliser
# ------------------------------------------------------------------------------------------------
def _channels_case(W):
    return [dict(f0=2 * np.pi * 0.125, bw=2 * np.pi * 0.03, guard=1.0),
            dict(f0=2 * np.pi * 0.7003, bw=2 * np.pi * 0.011, guard=1.5),
            dict(f0=2 * np.pi * 0.0007, bw=2 * np.pi * 0.02, guard=1.0),           # wraps around DC
            dict(f0=2 * np.pi * 0.3301, bw=2 * np.pi * 0.004, guard=2.0, precise=True)]


@pytest.mark.parametrize("W", [4096, 8192, 65536])
def test_channeliser_matches_oracle(sdb, oracle, W):
    hops = 12
    n = W // 2 * hops
    x = _noise(n, seed=W % 89, scale=0.05)
    t = np.arange(n)
    for f, a in [(0.125, 0.4), (0.7003, 0.3), (0.0007, 0.2), (0.3301, 0.25), (0.45, 0.5)]:
        x += (a * np.exp(2j * np.pi * (f + 1e-4) * t)).astype(np.complex64)
    chans = _channels_case(W)
    ref = oracle.specttuner_run(x, W, chans)
    e = sdb.Engine(n_streams=1, psd_size=0, st_window_size=W, max_feed=n)
    hs = [e.open_channel(c["f0"], c["bw"], c["guard"], c.get("precise", False)) for c in chans]
    e.commit()
    e.feed(x[None, :])
    x_rms = float(np.sqrt(np.mean(np.abs(x) ** 2)))
    for h, r in zip(hs, ref):
        got = e.read_channel(0, h)
        info = e.channel_info(h)
        assert len(got) == len(r) == (hops - 1) * info.size // 2
        parity.assert_channel_close(got, r, x_rms, info.decimation)
        assert np.array_equal(got.view(np.uint32), r.view(np.uint32)), "channel %d not bit-identical" % h


def test_channeliser_block_size_independence(sdb, oracle):
    W, hops = 8192, 16
    n = W // 2 * hops
    x = _noise(n, seed=77, scale=0.2)
    chans = _channels_case(W)
    outs = []
    for split in ([n], [W, W // 2 * 3, n - W - W // 2 * 3], [W // 2] * hops):
        e = sdb.Engine(n_streams=1, psd_size=0, st_window_size=W, max_feed=n)
        hs = [e.open_channel(c["f0"], c["bw"], c["guard"], c.get("precise", False)) for c in chans]
        e.commit()
        acc = [[] for _ in hs]
        off = 0
        for m in split:
            e.feed(x[None, off:off + m])
            off += m
            for i, h in enumerate(hs):
                acc[i].append(e.read_channel(0, h))
        outs.append([np.concatenate(a) for a in acc])
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ------------------------------------------------------------------------------------------------
# inspector chains on identical channel-rate input: bit-exact
# ------------------------------------------------------------------------------------------------
def _psk_capture(n, sps, order, seed, noise_db=-40.0, cfo=2e-4):
    s, _ = synth.psk_signal(n, sps, order=order, seed=seed)
    rng = np.random.default_rng(seed)
    x = synth.mix(s, cfo, 0.3) * 0.2 + synth.awgn(n, 10 ** (noise_db / 20), rng)
    return x.astype(np.complex64)


@pytest.mark.parametrize("order,costas,bps", [(2, 1, 1), (4, 2, 2), (8, 3, 3)])
def test_psk_chain_bit_exact(sdb, oracle, order, costas, bps):
    fs, sps, n = 1.0, 3.125, 40000
    x = _psk_capture(n, sps, order, seed=order)
    kw = dict(baud=fs / sps, costas_order=costas, bits_per_symbol=bps, loop_bw=fs * 2e-3, mf_type=1,
              mf_rolloff=0.35, clock_type=1, clock_gain=0.1)
    (soft, hard), = sdb.inspector_run("psk", fs, x, **kw)
    rs, rh = oracle.inspector_run(oracle.insp_config("psk", fs, **kw), x)
    assert len(rs) > n / sps * 0.9
    parity.assert_symbols_match(soft, hard, rs, rh, exact_soft=True)


@pytest.mark.parametrize("kw", [
    dict(agc_enabled=0, agc_gain_db=6.0, costas_order=0, offset=2e-4, mf_type=0, clock_type=0, clock_phase=0.3),
    dict(agc_enabled=1, costas_order=2, mf_type=0, clock_type=1, clock_gain=0.3),
    dict(agc_enabled=1, costas_order=2, mf_type=1, mf_rolloff=0.5, clock_type=0, clock_phase=0.0),
    dict(agc_enabled=1, costas_order=2, mf_type=1, clock_type=1, clock_running=0),
])
def test_psk_chain_variants_bit_exact(sdb, oracle, kw):
    fs, sps, n = 1.0, 4.0, 20000
    x = _psk_capture(n, sps, 4, seed=9)
    kw = dict(baud=fs / sps, bits_per_symbol=2, loop_bw=fs * 1e-3, **kw)
    (soft, hard), = sdb.inspector_run("psk", fs, x, **kw)
    rs, rh = oracle.inspector_run(oracle.insp_config("psk", fs, **kw), x)
    parity.assert_symbols_match(soft, hard, rs, rh, exact_soft=True)


@pytest.mark.parametrize("locked,rate", [(0, 1e-3), (0, 2e-2), (1, 1e-3)])
def test_psk_chain_cma_bit_exact(sdb, oracle, locked, rate):
    """SPEC E: CMA equaliser after the clock recovery, multipath capture, across two feeds (state kept)."""
    fs, sps, n = 1.0, 4.0, 24000
    x = _psk_capture(n, sps, 4, seed=21)
    x = (x + 0.35 * np.exp(0.7j) * np.roll(x, 3) - 0.2j * np.roll(x, 9)).astype(np.complex64)
    kw = dict(baud=fs / sps, costas_order=2, bits_per_symbol=2, loop_bw=fs * 1e-3, mf_type=1, clock_type=1,
              clock_gain=0.1, eq_type=1, eq_rate=rate, eq_locked=locked)
    (soft, hard), = sdb.inspector_run("psk", fs, x, **kw)
    rs, rh = oracle.inspector_run(oracle.insp_config("psk", fs, **kw), x)
    assert len(rs) > n / sps * 0.9
    parity.assert_symbols_match(soft, hard, rs, rh, exact_soft=True)
    if not locked:   # the equaliser does something: the output differs from the unequalised chain
        kw0 = dict(kw, eq_type=0)
        (soft0, _), = sdb.inspector_run("psk", fs, x, **kw0)
        assert not np.array_equal(soft0.view(np.uint32), soft.view(np.uint32))


@pytest.mark.parametrize("quad", [0, 1])
def test_fsk_chain_bit_exact(sdb, oracle, quad):
    fs, sps, n = 1.0, 5.0, 30000
    s, _ = synth.fsk_signal(n, sps, h=1.0, seed=4)
    rng = np.random.default_rng(4)
    x = (0.3 * s + synth.awgn(n, 10 ** (-35 / 20), rng)).astype(np.complex64)
    kw = dict(baud=fs / sps, bits_per_symbol=1, fsk_phase=0.2, fsk_quad_demod=quad, mf_type=1, clock_type=1,
              clock_gain=0.2)
    (soft, hard), = sdb.inspector_run("fsk", fs, x, **kw)
    rs, rh = oracle.inspector_run(oracle.insp_config("fsk", fs, **kw), x)
    assert len(rs) > 0.9 * n / sps
    parity.assert_symbols_match(soft, hard, rs, rh, exact_soft=True)


@pytest.mark.parametrize("use_pll,chan", [(0, 0), (1, 1), (1, 2), (0, 1)])
def test_ask_chain_bit_exact(sdb, oracle, use_pll, chan):
    fs, sps, n = 1.0, 6.25, 30000
    s, _ = synth.ask_signal(n, sps, levels=4, seed=6)
    rng = np.random.default_rng(6)
    x = (0.4 * synth.mix(s, 1e-4, 0.5) + synth.awgn(n, 10 ** (-40 / 20), rng)).astype(np.complex64)
    kw = dict(baud=fs / sps, bits_per_symbol=2, ask_use_pll=use_pll, ask_channel=chan, loop_bw=fs * 5e-3,
              offset=1e-4, mf_type=1, clock_type=1, clock_gain=0.2)
    (soft, hard), = sdb.inspector_run("ask", fs, x, **kw)
    rs, rh = oracle.inspector_run(oracle.insp_config("ask", fs, **kw), x)
    parity.assert_symbols_match(soft, hard, rs, rh, exact_soft=True)


@pytest.mark.parametrize("demod", ["am", "fm", "usb", "lsb"])
def test_audio_chain_bit_exact(sdb, oracle, demod):
    fs, n = 200000.0, 60000
    t = np.arange(n) / fs
    tone = np.cos(2 * np.pi * 1000 * t)
    if demod == "am":
        s = (1 + 0.5 * tone) * np.exp(1j * 0.4)
    elif demod == "fm":
        s = np.exp(1j * 2 * np.pi * 5000 * np.cumsum(tone) / fs)
    else:
        s = np.exp(2j * np.pi * (700 if demod == "usb" else -700) * t) + 0.5 * np.exp(
            2j * np.pi * (1900 if demod == "usb" else -1900) * t)
    rng = np.random.default_rng(3)
    x = (0.2 * s + synth.awgn(n, 1e-3, rng)).astype(np.complex64)
    kw = dict(audio_demod=sdb.AUDIO[demod], audio_cutoff=5000.0, audio_sample_rate=44100, agc_enabled=1,
              agc_ts=0.01, offset=1500.0, audio_squelch=1, audio_squelch_level=1e-6)
    (soft, hard), = sdb.inspector_run("audio", fs, x, **kw)
    okw = dict(kw)
    rs, _ = oracle.inspector_run(oracle.insp_config("audio", fs, **okw), x)
    assert abs(len(rs) - n * 44100 / fs) <= 2
    assert len(soft) == len(rs)
    assert np.array_equal(soft.view(np.uint32), rs.view(np.uint32))


def test_inspector_batch_is_independent(sdb, oracle):
    fs, sps, n, B = 1.0, 3.125, 8000, 40
    xs = np.stack([_psk_capture(n, sps, 4, seed=100 + b) for b in range(B)])
    kw = dict(baud=fs / sps, costas_order=2, bits_per_symbol=2, loop_bw=fs * 2e-3, mf_type=1, clock_type=1,
              clock_gain=0.1)
    res = sdb.inspector_run("psk", fs, xs, **kw)
    cfg = oracle.insp_config("psk", fs, **kw)
    for b in (0, 7, 31, 39):
        rs, rh = oracle.inspector_run(cfg, xs[b])
        parity.assert_symbols_match(res[b][0], res[b][1], rs, rh, exact_soft=True)


# ------------------------------------------------------------------------------------------------
# Tasks/ primitives
# ------------------------------------------------------------------------------------------------
def test_task_carrier_xlate_bit_exact(sdb, oracle):
    import ctypes as C
    n = 20000
    x = _noise(n, 1)
    got = sdb.carrier_xlate(x, 0.0123, 0.7)
    o = oracle.Ncqo()
    L = oracle.lib()
    L.sdo_ncqo_init(C.byref(o), C.c_float(-0.0123))
    L.sdo_ncqo_set_phase(C.byref(o), C.c_float(-0.7))
    ref = np.empty_like(x)
    L.sdo_carrier_xlate(oracle.ptr(x), oracle.ptr(ref), n, C.byref(o))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_task_quad_demod_bit_exact(sdb, oracle):
    import ctypes as C
    n = 30000
    x = _noise(n, 2)
    got = sdb.quad_demod(x)
    ref = np.empty_like(x)
    prev = oracle.Cpx(0, 0)
    primed = C.c_int(0)
    oracle.lib().sdo_quad_demod(oracle.ptr(x), oracle.ptr(ref), n, C.byref(prev), C.byref(primed))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert got[0] == 0 and np.all(got.real == 0)


def _oracle_loop(oracle, init, feed, x):
    ref = np.empty_like(x)
    v = x.view(np.float32).reshape(-1, 2)
    for i in range(len(x)):
        r = feed(oracle.Cpx(float(v[i, 0]), float(v[i, 1])))
        ref[i] = np.float32(r.re) + 1j * np.float32(r.im)
    return ref


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_task_costas_bit_exact(sdb, oracle, kind):
    import ctypes as C
    n = 6000
    x = _psk_capture(n, 4.0, {1: 2, 2: 4, 3: 8}[kind], seed=kind, cfo=3e-4)
    tau, loop = 4.0, 2e-3
    got = sdb.costas(x, kind, tau, loop)
    L = oracle.lib()
    c = oracle.Costas()
    assert L.sdo_costas_init(C.byref(c), kind, 0.0, 1.0 / tau, 3, loop) == 0
    ref = _oracle_loop(oracle, None, lambda v: L.sdo_costas_feed(C.byref(c), v), x)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_task_pll_and_agc_bit_exact(sdb, oracle):
    import ctypes as C
    n = 6000
    x = (0.3 * np.exp(1j * (2 * np.pi * 2e-3 * np.arange(n) + 0.5))).astype(np.complex64) + _noise(n, 3, 0.01)
    L = oracle.lib()
    p = oracle.Pll()
    L.sdo_pll_init(C.byref(p), 0.0, 0.01)
    ref = _oracle_loop(oracle, None, lambda v: L.sdo_pll_track(C.byref(p), v), x)
    assert np.array_equal(sdb.pll(x, 0.01).view(np.uint32), ref.view(np.uint32))
    # AGC: AGCTask parameterisation (fractions doubled, history sizes at their defaults)
    tau = 10.0
    env = (0.05 + 0.5 * (np.arange(n) > n // 2)).astype(np.float32)
    y = (x * env).astype(np.complex64)
    ap = oracle.AgcParams()
    L.sdo_agc_params_from_tau(C.byref(ap), tau, 2.0)
    ap.delay_line_size, ap.mag_history_size = 20, 20
    a = oracle.Agc()
    assert L.sdo_agc_init(C.byref(a), C.byref(ap)) == 0
    ref = _oracle_loop(oracle, None, lambda v: L.sdo_agc_feed(C.byref(a), v), y)
    assert np.array_equal(sdb.agc(y, tau).view(np.uint32), ref.view(np.uint32))


def test_task_lpf_matches_oracle_and_length(sdb, oracle):
    n = 30000
    x = _noise(n, 8, 0.3)
    bw = 0.2
    got = sdb.lpf(x, bw)
    assert got.shape == x.shape
    bw_ang = np.float32(np.pi) * np.float32(bw)
    guard = np.float32(2 * np.pi) / bw_ang
    pad = np.concatenate([x, np.zeros(4096, np.complex64)])
    ref = oracle.specttuner_run(pad, 4096, [dict(f0=0.0, bw=float(bw_ang), guard=float(guard))])[0][:n]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # it is a low-pass: out-of-band power is gone, in-band kept
    G = np.abs(np.fft.fft(got[4096:4096 + 16384])) ** 2
    X = np.abs(np.fft.fft(x[4096:4096 + 16384])) ** 2
    f = np.fft.fftfreq(16384)
    assert G[np.abs(f) > 0.07].sum() < 1e-3 * X[np.abs(f) > 0.07].sum()
    assert abs(G[np.abs(f) < 0.04].sum() / X[np.abs(f) < 0.04].sum() - 1) < 0.05


# ------------------------------------------------------------------------------------------------
# whole pipeline
# ------------------------------------------------------------------------------------------------
def _pipeline_case(N, hops, seed, carriers, S=1):
    n = N // 2 * hops
    xs, metas = [], []
    for s in range(S):
        x, meta = synth.multi_carrier(n, 1.0, carriers, noise_db=-50.0, seed=seed + s)
        xs.append(x)
        metas.append(meta)
    return np.stack(xs), metas


def test_pipeline_qpsk_matches_oracle(sdb, oracle):
    N, hops = 65536, 24
    baud = 1.0 / 100
    x, _ = _pipeline_case(N, hops, 21, [("qpsk", 0.125 + 3e-6, baud, -12.0, {})])
    f0, bw = 2 * np.pi * 0.125, 2 * np.pi * 3 * baud
    e = sdb.Engine(n_streams=1, psd_size=N, psd_window="blackmann_harris", max_feed=x.shape[1])
    h = e.open_channel(f0, bw, 1.0)
    kw = dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=e.channel_rate(h) * 2e-3, mf_type=1,
              mf_rolloff=0.35, clock_type=1, clock_gain=0.1)
    e.set_inspector(h, "psk", **kw)
    e.commit()
    e.feed(x)
    ic = oracle.insp_config("psk", e.channel_rate(h), **kw)
    ref = oracle.analyzer_run(oracle.make_an_params(N, "blackmann_harris", [(f0, bw, 1.0, 0, ic)]), x[0])
    assert np.array_equal(e.read_psd()[0].view(np.uint32), ref["psd"].view(np.uint32))
    soft, hard = e.read_symbols(0, h)
    assert len(hard) > 1000
    # end to end: PSD bins, soft symbols and hard symbols all bit-identical to the oracle
    parity.assert_symbols_match(soft, hard, ref["soft"][0], ref["hard"][0], exact_soft=True)


def test_pipeline_mixed_inspectors_streams_and_chunks(sdb, oracle):
    N, hops, S = 16384, 20, 3
    specs = [("qpsk", 0.10, 1 / 128., -14.0, {}), ("fsk", 0.30, 1 / 160., -14.0, {}),
             ("ask", 0.55, 1 / 256., -12.0, {"levels": 2}), ("bpsk", 0.80, 1 / 128., -14.0, {})]
    x, _ = _pipeline_case(N, hops, 40, specs, S=S)
    e = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=N * 4)
    chans, hs = [], []
    for kind, f, baud, _, _ in specs:
        f0, bw = 2 * np.pi * f, 2 * np.pi * 3 * baud
        h = e.open_channel(f0, bw, 1.0)
        fs_ch = e.channel_rate(h)
        if kind in ("qpsk", "bpsk"):
            cls = "psk"
            kw = dict(baud=baud, costas_order=2 if kind == "qpsk" else 1, bits_per_symbol=2 if kind == "qpsk" else 1,
                      loop_bw=fs_ch * 2e-3, mf_type=1, clock_type=1, clock_gain=0.1)
        elif kind == "fsk":
            cls = "fsk"
            kw = dict(baud=baud, bits_per_symbol=1, mf_type=1, clock_type=1, clock_gain=0.2)
        else:
            cls = "ask"
            kw = dict(baud=baud, bits_per_symbol=1, ask_use_pll=1, ask_channel=0, loop_bw=fs_ch * 5e-3, mf_type=1,
                      clock_type=1, clock_gain=0.2)
        e.set_inspector(h, cls, **kw)
        chans.append((f0, bw, 1.0, 0, oracle.insp_config(cls, fs_ch, **kw)))
        hs.append(h)
    e.commit()
    got_soft = [[[] for _ in hs] for _ in range(S)]
    got_hard = [[[] for _ in hs] for _ in range(S)]
    psd = []
    n = x.shape[1]
    for off in range(0, n, N * 4):
        seg = x[:, off:off + N * 4]
        if seg.shape[1] % N:
            seg = seg[:, :seg.shape[1] // N * N]
        if seg.shape[1] == 0:
            break
        e.feed(seg)
        psd.append(e.read_psd().copy())
        for s in range(S):
            for i, h in enumerate(hs):
                a, b = e.read_symbols(s, h)
                got_soft[s][i].append(a)
                got_hard[s][i].append(b)
    used = sum(p.shape[1] for p in psd) * N
    failures = []
    for s in range(S):
        ref = oracle.analyzer_run(oracle.make_an_params(N, "hann", chans), x[s, :used])
        assert np.array_equal(np.concatenate([p[s] for p in psd]).view(np.uint32), ref["psd"].view(np.uint32))
        for i in range(len(hs)):
            try:
                parity.assert_symbols_match(np.concatenate(got_soft[s][i]), np.concatenate(got_hard[s][i]),
                                            ref["soft"][i], ref["hard"][i], exact_soft=True)
            except AssertionError as exc:
                failures.append("stream %d channel %d (%s): %s" % (s, i, specs[i][0], exc))
    assert not failures, "\n".join(failures)


@pytest.mark.parametrize("fmt", ["u8", "s8", "s16"])
def test_native_sample_formats(sdb, oracle, fmt):
    """8 / 16-bit SDR sample formats (Default/SourceConfig/FileSourcePage.cpp:80-104) converted inside the first
    load (SPEC Q): identical to feeding the converted float32 samples, for PSD (small and four-step paths),
    channeliser history across feeds, and symbols."""
    N, hops, S = 65536 if fmt == "s16" else 8192, 10, 2
    n = N // 2 * hops
    baud = 1 / 100.
    xs = []
    for s in range(S):
        xf, _ = synth.multi_carrier(n, 1.0, [("qpsk", 0.125 + 3e-6, baud, -8.0, {})], noise_db=-45.0, seed=60 + s)
        xs.append(xf)
    xf = np.stack(xs)
    iq = np.stack([xf.real, xf.imag], axis=-1)
    if fmt == "u8":
        q = np.clip(np.rint(iq * 128.0 + 128.0), 0, 255).astype(np.uint8)
        back = (q.astype(np.float32) - 128.0) / 128.0
    elif fmt == "s8":
        q = np.clip(np.rint(iq * 128.0), -128, 127).astype(np.int8)
        back = q.astype(np.float32) / 128.0
    else:
        q = np.clip(np.rint(iq * 32768.0), -32768, 32767).astype(np.int16)
        back = q.astype(np.float32) / 32768.0
    xc = (back[..., 0] + 1j * back[..., 1]).astype(np.complex64)      # exactly what the kernels must see
    f0, bw = 2 * np.pi * 0.125, 2 * np.pi * 3 * baud
    e = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=n, input_format=fmt)
    h = e.open_channel(f0, bw, 1.0)
    kw = dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=e.channel_rate(h) * 2e-3, mf_type=1,
              clock_type=1, clock_gain=0.1)
    e.set_inspector(h, "psk", **kw)
    e.commit()
    ic = oracle.insp_config("psk", e.channel_rate(h), **kw)
    psd, soft, hard = [], [[] for _ in range(S)], [[] for _ in range(S)]
    half = n // 2 // N * N
    for seg in (slice(0, half), slice(half, n)):
        e.feed(q[:, seg])
        psd.append(e.read_psd().copy())
        for s in range(S):
            a, b = e.read_symbols(s, h)
            soft[s].append(a)
            hard[s].append(b)
    for s in range(S):
        ref = oracle.analyzer_run(oracle.make_an_params(N, "hann", [(f0, bw, 1.0, 0, ic)]), xc[s], want_chan=False)
        assert np.array_equal(np.concatenate([p[s] for p in psd]).view(np.uint32), ref["psd"].view(np.uint32))
        parity.assert_symbols_match(np.concatenate(soft[s]), np.concatenate(hard[s]), ref["soft"][0], ref["hard"][0],
                                    exact_soft=True)


@pytest.mark.parametrize("fmt,N", [("f32", 65536), ("f32", 4096), ("s16", 65536), ("u8", 8192), ("f32", 32768),
                                   ("s16", 32768), ("u8", 32768)])
def test_iq_reverse_flag(sdb, fmt, N):
    """SDB_FLAG_IQ_REVERSE (suscan_analyzer_set_iq_reverse): the swap happens inside the first load, for every
    transform size path and sample format, history across feeds included: bit-identical to feeding (Q, I)."""
    S, n = 2, N * 4
    rng = np.random.default_rng(N)
    iq = (0.4 * rng.standard_normal((S, n, 2))).astype(np.float32)
    if fmt == "s16":
        q = np.clip(np.rint(iq * 32768.0), -32768, 32767).astype(np.int16)
    elif fmt == "u8":
        q = np.clip(np.rint(iq * 128.0 + 128.0), 0, 255).astype(np.uint8)
    else:
        q = iq.view(np.complex64)[..., 0]
    swapped = np.ascontiguousarray(q[..., ::-1]) if fmt != "f32" else (q.imag + 1j * q.real).astype(np.complex64)

    def run(data, flags):
        e = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=n, input_format=fmt, flags=flags)
        h = e.open_channel(2 * np.pi * 0.2, 2 * np.pi / 16, 1.0)
        e.commit()
        out = []
        for seg in (slice(0, n // 2), slice(n // 2, n)):
            e.feed(data[:, seg])
            out.append((e.read_psd().copy(), [e.read_channel(s, h) for s in range(S)]))
        return out

    a = run(q, sdb.FLAG_IQ_REVERSE)
    b = run(swapped, 0)
    for (pa, ca), (pb, cb) in zip(a, b):
        assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32))
        for x, y in zip(ca, cb):
            assert len(x) == len(y) and np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_async_pipeline_matches_sync(sdb):
    """feed_host + read_*_async queued back to back (H2D / kernels / D2H of neighbouring feeds overlap,
    results double-buffered) must give exactly what feed + blocking reads give."""
    N, S, feeds = 16384, 4, 5
    n = N * 4
    specs = [("qpsk", 0.10, 1 / 128., -14.0, {}), ("fsk", 0.30, 1 / 160., -14.0, {})]
    x, _ = _pipeline_case(N, 8 * feeds, 70, specs, S=S)

    def make():
        e = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=n)
        hs = []
        for kind, f, baud, _, _ in specs:
            h = e.open_channel(2 * np.pi * f, 2 * np.pi * 3 * baud, 1.0)
            if kind == "qpsk":
                e.set_inspector(h, "psk", baud=baud, costas_order=2, bits_per_symbol=2,
                                loop_bw=e.channel_rate(h) * 2e-3, mf_type=1, clock_type=1, clock_gain=0.1)
            else:
                e.set_inspector(h, "fsk", baud=baud, bits_per_symbol=1, mf_type=1, clock_type=1, clock_gain=0.2)
            hs.append(h)
        e.commit()
        return e, hs

    e1, hs = make()
    cap = e1.symbol_capacity
    K = len(hs)
    ref = []
    for i in range(feeds):
        e1.feed(x[:, i * n:(i + 1) * n])
        cnt = np.zeros(S * K, np.uint32)
        soft = np.zeros((S * K, cap), np.complex64)
        hard = np.zeros((S * K, cap), np.uint8)
        e1.read_all_symbols(cnt, soft, hard, cap)
        ref.append((e1.read_psd().copy(), cnt, soft, hard))
    e2, _ = make()
    outs = []
    for i in range(feeds):
        seg = np.ascontiguousarray(x[:, i * n:(i + 1) * n])
        outs.append((seg, np.zeros((S, n // N, N), np.float32), np.zeros(S * K, np.uint32),
                     np.zeros((S * K, cap), np.complex64), np.zeros((S * K, cap), np.uint8)))
        e2.feed_host_ptr(seg.ctypes.data, n, n)
        e2.read_psd_async(outs[-1][1])
        e2.read_all_symbols_async(outs[-1][2], outs[-1][3], outs[-1][4], cap)
    e2.sync()
    for i in range(feeds):
        assert np.array_equal(outs[i][1].view(np.uint32), ref[i][0].view(np.uint32)), "psd of feed %d" % i
        assert np.array_equal(outs[i][2], ref[i][1])
        for c in range(S * K):
            m = ref[i][1][c]
            assert np.array_equal(outs[i][3][c, :m].view(np.uint32), ref[i][2][c, :m].view(np.uint32))
            assert np.array_equal(outs[i][4][c, :m], ref[i][3][c, :m])


def test_packed_symbol_read_matches_unpacked(sdb):
    """sdb_engine_read_symbols_packed_async: the GPU lays the chains' symbols back to back (16-symbol aligned starts,
    zero-filled gaps) straight into pinned host memory; same symbols as the [chains][cap] read, queued asynchronously
    behind two feeds; pageable destinations and too small buffers are refused / reported"""
    import torch
    N, S, feeds = 16384, 5, 3
    n = N * 4
    specs = [("qpsk", 0.10, 1 / 128., -14.0, {}), ("fsk", 0.30, 1 / 160., -14.0, {}), ("qpsk", 0.55, 1 / 256., -14.0, {})]
    x, _ = _pipeline_case(N, 8 * feeds, 71, specs, S=S)
    e = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=n)
    for kind, f, baud, _, _ in specs:
        h = e.open_channel(2 * np.pi * f, 2 * np.pi * 3 * baud, 1.0)
        if kind == "qpsk":
            e.set_inspector(h, "psk", baud=baud, costas_order=2, bits_per_symbol=2,
                            loop_bw=e.channel_rate(h) * 2e-3, mf_type=1, clock_type=1, clock_gain=0.1)
        else:
            e.set_inspector(h, "fsk", baud=baud, bits_per_symbol=1, mf_type=1, clock_type=1, clock_gain=0.2)
    e.commit()
    cap, chains = e.symbol_capacity, S * len(specs)
    total_cap = chains * ((cap + 15) // 16 * 16)
    pin = lambda shape, dt: torch.zeros(shape, dtype=dt, pin_memory=True).numpy()
    outs = []
    for i in range(feeds):
        seg = np.ascontiguousarray(x[:, i * n:(i + 1) * n])
        o = dict(seg=seg, cnt=pin((chains,), torch.int32).view(np.uint32), off=pin((chains + 1,), torch.int64).view(np.uint64),
                 soft=pin((total_cap,), torch.complex64), hard=pin((total_cap,), torch.uint8))
        o["soft"][:] = 7 + 7j
        o["hard"][:] = 77
        e.feed_host_ptr(seg.ctypes.data, n, n)
        e.read_symbols_packed_async(o["cnt"], o["off"], o["soft"], o["hard"], total_cap)
        # the unpacked read of the same feed, for comparison
        o["rc"], o["rs"], o["rh"] = (np.zeros(chains, np.uint32), np.zeros((chains, cap), np.complex64),
                                     np.zeros((chains, cap), np.uint8))
        e.read_all_symbols_async(o["rc"], o["rs"], o["rh"], cap)
        outs.append(o)
    e.sync()
    for i, o in enumerate(outs):
        assert np.array_equal(o["cnt"], o["rc"]) and o["cnt"].sum() > 0
        assert o["off"][0] == 0 and np.all(o["off"] % 16 == 0)
        assert np.array_equal(np.diff(o["off"].astype(np.int64)), (o["cnt"].astype(np.int64) + 15) // 16 * 16)
        for c in range(chains):
            m, a = int(o["cnt"][c]), int(o["off"][c])
            assert np.array_equal(o["soft"][a:a + m].view(np.uint32), o["rs"][c, :m].view(np.uint32)), (i, c)
            assert np.array_equal(o["hard"][a:a + m], o["rh"][c, :m]), (i, c)
            pad = int(o["off"][c + 1]) - a - m
            assert not o["soft"][a + m:a + m + pad].any() and not o["hard"][a + m:a + m + pad].any()
        assert np.all(o["hard"][int(o["off"][-1]):] == 77)       # nothing written past the extent
    # a destination that is too small: the rows that fit are written, the extent tells
    o = outs[-1]
    small = int(o["off"][chains // 2])
    cnt2, off2 = np.zeros(chains, np.uint32), np.zeros(chains + 1, np.uint64)
    soft2 = pin((total_cap,), torch.complex64)
    soft2[:] = 9
    e.read_symbols_packed(cnt2, off2, soft2, None, small)
    assert int(off2[-1]) > small and np.array_equal(soft2[:small].view(np.uint32), o["soft"][:small].view(np.uint32))
    assert np.all(soft2[small:] == 9)
    # device destination
    dsoft = torch.zeros(total_cap, dtype=torch.complex64, device="cuda")
    e.read_symbols_packed(cnt2, off2, dsoft.data_ptr(), None, total_cap)
    assert np.array_equal(dsoft.cpu().numpy()[:int(off2[-1])].view(np.uint32), o["soft"][:int(off2[-1])].view(np.uint32))
    with pytest.raises(sdb.SdbError):                          # pageable memory cannot be written by the GPU
        e.read_symbols_packed(cnt2, off2, np.zeros(total_cap + 2, np.complex64)[2:], None, total_cap)


def test_error_paths(sdb):
    e = sdb.Engine(n_streams=1, psd_size=8192, max_feed=8192)
    with pytest.raises(sdb.SdbError):
        e.open_channel(7.0, 0.1)          # f0 outside [0, 2 pi)
    with pytest.raises(sdb.SdbError):
        e.open_channel(1.0, 0.1, guard=0.5)
    h = e.open_channel(1.0, 0.1)
    with pytest.raises(sdb.SdbError):
        cfg = sdb.InspectorConfig()
        sdb._check(sdb.load_library().sdb_engine_set_inspector(e._h, h + 5, __import__("ctypes").byref(cfg)))  # wrong handle
    with pytest.raises(sdb.SdbError):
        e.feed(np.zeros((1, 8192), np.complex64))   # not committed
    e.commit()
    with pytest.raises(sdb.SdbError):
        e.feed(np.zeros((1, 100), np.complex64))    # not a multiple of the frame
    with pytest.raises(sdb.SdbError):
        sdb.Engine(n_streams=1, psd_size=1000, max_feed=1000)

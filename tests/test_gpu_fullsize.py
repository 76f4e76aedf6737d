"""-m gpu: BASELINE.json configs at their full sizes (65536-pt PSD, 100 / 200 MS/s geometry).

Oracle comparison on a few streams (the CPU oracle does ~10 MS/s per core, so a stream of 2^20 samples is
checked directly), plus size-independent properties on every stream of the batch: Parseval on the PSD,
transmit PRBS -> recover the same PRBS from the hard symbols, identical results whatever the batch."""
import numpy as np
import pytest

import parity
from sigdigger_b200 import synth

pytestmark = pytest.mark.gpu

N = 65536
FS = 100e6


def _cfg2_stream(n, seed, baud=1e6, fc=12.5e6 + 300.0, amp_db=-12.0, noise_db=-60.0):
    x, meta = synth.multi_carrier(n, FS, [("qpsk", fc, baud, amp_db, {})], noise_db=noise_db, seed=seed)
    return x, meta[0]


def _qpsk_kw(fs_ch, baud=1e6):
    return dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=fs_ch * 2e-3, mf_type=1, mf_rolloff=0.35,
                clock_type=1, clock_gain=0.1)


def test_cfg2_full_size_batch(sdb, oracle):
    """configs[1]: 100 MS/s, 65536-pt Blackman-Harris spectrum + 1 QPSK inspector (1 MBd, RRC 0.35)."""
    S, hops = 6, 32
    n = hops * N // 2                                  # 2^20 samples per stream
    xs, refs = [], []
    for s in range(S):
        x, sym = _cfg2_stream(n, seed=100 + s)
        xs.append(x)
        refs.append(sym)
    x = np.stack(xs)
    f0 = float(np.float32(2 * np.pi * 12.5e6 / FS))
    bw = float(np.float32(2 * np.pi * 3e6 / FS))
    e = sdb.Engine(n_streams=S, psd_size=N, psd_window="blackmann_harris", max_feed=n, samp_rate=FS)
    h = e.open_channel(f0, bw, 1.0)
    info = e.channel_info(h)
    assert (info.size, info.decimation) == (2048, 32.0)    # SURVEY section 8 table, cfg 2
    kw = _qpsk_kw(e.channel_rate(h))
    e.set_inspector(h, "psk", **kw)
    e.commit()
    e.feed(x)
    psd = e.read_psd()
    # --- oracle on two of the streams: everything bit-identical
    for s in (0, S - 1):
        ic = oracle.insp_config("psk", e.channel_rate(h), **kw)
        ref = oracle.analyzer_run(oracle.make_an_params(N, "blackmann_harris", [(f0, bw, 1.0, 0, ic)]), x[s],
                                  want_chan=False)
        assert np.array_equal(psd[s].view(np.uint32), ref["psd"].view(np.uint32))
        soft, hard = e.read_symbols(s, h)
        parity.assert_symbols_match(soft, hard, ref["soft"][0], ref["hard"][0], exact_soft=True)
    # --- properties on every stream
    w = oracle.window(N, "blackmann_harris").astype(np.float64)
    for s in range(S):
        frames = x[s].reshape(-1, N).astype(np.complex128)
        energy = np.sum(np.abs(frames * w) ** 2, axis=1)             # Parseval: sum_k psd[k] = sum |w x|^2
        assert np.allclose(psd[s].astype(np.float64).sum(axis=1), energy, rtol=2e-6)
        soft, hard = e.read_symbols(s, h)
        assert abs(len(hard) - (hops - 1) * 1024 / 3.125) < 8        # symbol count: (n - N/2) / 100
        # recover the transmitted PRBS symbols (up to the loop's 4-fold phase ambiguity and a lag)
        tx = refs[s]
        hh = hard[4000:9000].astype(int)
        best = min((np.count_nonzero(((hh + rot) % 4) != tx[lag:lag + len(hh)]), lag, rot)
                   for lag in range(3980, 4060) for rot in range(4))
        assert best[0] == 0, "stream %d: PRBS not recovered: %r" % (s, best)


def test_cfg4_audio_path(sdb, oracle):
    """configs[3]: 50 MS/s, 32768-pt spectrum + 8 audio channels (3 AM, 3 FM, 2 USB), 200 kHz wide
    (Default/Audio/AudioProcessor.cpp:118-121) -> 256-pt IFFT, decimation 128, resampled to 44.1 kS/s."""
    fs, Np, hops = 50e6, 32768, 40
    n = hops * Np // 2
    t = np.arange(n) / fs
    tone = np.cos(2 * np.pi * 1000 * t)
    x = synth.awgn(n, 10 ** (-60 / 20), np.random.default_rng(4))
    demods, f_off = ["am"] * 3 + ["fm"] * 3 + ["usb"] * 2, []
    for k, d in enumerate(demods):
        f = (k - 3.5) * 2.0e6
        f_off.append(f)
        if d == "am":
            s = (1 + 0.5 * tone) * 0.2
        elif d == "fm":
            s = 0.2 * np.exp(1j * 2 * np.pi * 5000 * np.cumsum(tone) / fs)
        else:
            s = 0.1 * (np.exp(2j * np.pi * 700 * t) + 0.5 * np.exp(2j * np.pi * 1900 * t))
        x = x + s * np.exp(2j * np.pi * f * t)
    x = x.astype(np.complex64)
    e = sdb.Engine(n_streams=1, psd_size=Np, psd_window="blackmann_harris", max_feed=n, samp_rate=fs)
    hs, ochans = [], []
    for d, f in zip(demods, f_off):
        f0 = float(np.float32(2 * np.pi * ((f / fs) % 1.0)))
        bw = float(np.float32(2 * np.pi * 200e3 / fs))
        h = e.open_channel(f0, bw, 1.0)
        assert e.channel_info(h).size == 256 and e.channel_info(h).decimation == 128.0
        kw = dict(audio_demod=sdb.AUDIO[d], audio_cutoff=5000.0, audio_sample_rate=44100, agc_enabled=1, agc_ts=0.0005,
                  offset=1300.0, audio_squelch=0, audio_volume=1.0)
        e.set_inspector(h, "audio", **kw)
        hs.append(h)
        ochans.append((f0, bw, 1.0, 0, oracle.insp_config("audio", e.channel_rate(h), **kw)))
    e.commit()
    e.feed(x[None, :])
    ref = oracle.analyzer_run(oracle.make_an_params(Np, "blackmann_harris", ochans), x, want_chan=False)
    assert np.array_equal(e.read_psd()[0].view(np.uint32), ref["psd"].view(np.uint32))
    for i, h in enumerate(hs):
        soft, _ = e.read_symbols(0, h)
        assert abs(len(soft) - (n - Np // 2) / 128 * 44100 / 390625.0) <= 2
        assert np.array_equal(soft.view(np.uint32), ref["soft"][i].view(np.uint32)), demods[i]
        # the 1 kHz tone comes out of the AM and FM channels
        if demods[i] in ("am", "fm"):
            a = soft.real[200:]
            spec = np.abs(np.fft.rfft((a - a.mean()) * np.hanning(len(a))))
            fpk = np.argmax(spec[2:]) + 2
            assert abs(fpk * 44100.0 / len(a) - 1000.0) < 150.0, (demods[i], fpk * 44100.0 / len(a))


def test_cfg3_full_size_64_inspectors(sdb, oracle):
    """configs[2]: 200 MS/s, 65536-pt spectrum + 64 inspectors (2-FSK / QPSK / ASK mix) on a 3 MHz raster."""
    fs = 200e6
    hops = 12
    n = hops * N // 2
    carriers, chans = [], []
    for k in range(64):
        f = (k - 31.5) * 3e6
        kind = ("fsk", "qpsk", "ask")[k % 3]
        baud = 0.5e6 if kind == "ask" else 1e6
        carriers.append((kind, f + 300.0, baud, -24.0, {"levels": 2} if kind == "ask" else {}))
    x, _ = synth.multi_carrier(n, fs, carriers, noise_db=-70.0, seed=7)
    e = sdb.Engine(n_streams=2, psd_size=N, psd_window="blackmann_harris", max_feed=n, samp_rate=fs)
    hs, ochans = [], []
    for kind, f, baud, _, _ in carriers:
        f0 = float(np.float32(2 * np.pi * (((f - 300.0) / fs) % 1.0)))
        bw = float(np.float32(2 * np.pi * 2.5e6 / fs))
        h = e.open_channel(f0, bw, 1.0)
        fs_ch = e.channel_rate(h)
        assert e.channel_info(h).size == 1024                        # SURVEY section 8 table, cfg 3
        if kind == "qpsk":
            cls, kw = "psk", _qpsk_kw(fs_ch)
        elif kind == "fsk":
            cls, kw = "fsk", dict(baud=baud, bits_per_symbol=1, mf_type=1, mf_rolloff=0.35, clock_type=1,
                                  clock_gain=0.2)
        else:
            cls, kw = "ask", dict(baud=baud, bits_per_symbol=1, ask_use_pll=1, ask_channel=0, loop_bw=fs_ch * 5e-3,
                                  mf_type=1, mf_rolloff=0.35, clock_type=1, clock_gain=0.2)
        e.set_inspector(h, cls, **kw)
        hs.append(h)
        ochans.append((f0, bw, 1.0, 0, oracle.insp_config(cls, fs_ch, **kw)))
    e.commit()
    xb = np.stack([x, x[::-1].copy()])                               # second stream: a different signal
    e.feed(xb)
    ref = oracle.analyzer_run(oracle.make_an_params(N, "blackmann_harris", ochans), x, want_chan=False)
    assert np.array_equal(e.read_psd()[0].view(np.uint32), ref["psd"].view(np.uint32))
    for i, h in enumerate(hs):
        soft, hard = e.read_symbols(0, h)
        assert len(hard) > 100
        parity.assert_symbols_match(soft, hard, ref["soft"][i], ref["hard"][i], exact_soft=True)

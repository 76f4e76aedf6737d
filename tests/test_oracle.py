"""CPU tests (-m "not gpu"): pin the oracle against closed-form answers, float64 numpy/scipy and the
in-repo reference arithmetic; check the host logic and that the C-ABI library exports every symbol
include/sigdigger_b200.h declares.  The reference ships no golden vectors, so these
known-answer tests are the pinning this project can do; parity stays "unpinned" w.r.t. upstream."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy import signal

from sigdigger_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- SPEC M: elementary functions --
def test_math_accuracy(oracle):
    L = oracle.lib()
    x = np.linspace(-30, 30, 30001).astype(np.float32)
    s, c = oracle.sincos(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 2e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 2e-7
    xs = np.logspace(-37, 37, 4001).astype(np.float32)
    l = oracle.vec1(L.sdo_log10f, xs)
    ref = np.log10(xs.astype(np.float64))
    assert np.all(np.abs(l - ref) <= 3e-7 * np.maximum(1.0, np.abs(ref)))
    e = np.linspace(-36, 37, 4001).astype(np.float32)
    ee = oracle.vec1(L.sdo_exp10f, e)
    assert np.abs(ee / 10.0 ** e.astype(np.float64) - 1).max() < 3e-7
    rng = np.random.default_rng(0)
    yy, xx = rng.standard_normal(4000).astype(np.float32), rng.standard_normal(4000).astype(np.float32)
    at = np.array([L.sdo_atan2f(float(a), float(b)) for a, b in zip(yy, xx)])
    assert np.abs(at - np.arctan2(yy.astype(np.float64), xx.astype(np.float64))).max() < 5e-7
    assert L.sdo_atan2f(0.0, 0.0) == 0.0
    assert abs(L.sdo_atan2f(0.0, -1.0) - np.pi) < 1e-6
    assert abs(L.sdo_atan2f(-1.0, 0.0) + np.pi / 2) < 1e-6


# ---------------------------------------------------------------- SPEC F/W/P: FFT, windows, PSD --
@pytest.mark.parametrize("n", [2, 8, 64, 4096, 65536])
def test_fft_vs_numpy(oracle, n):
    rng = np.random.default_rng(n)
    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    F = oracle.fft(z)
    R = np.fft.fft(z.astype(np.complex128))
    assert np.abs(F - R).max() <= 2e-6 * np.sqrt(np.mean(np.abs(R) ** 2)) * max(1, np.log2(n)) ** 0.5
    back = oracle.fft(F, +1) / n
    assert np.abs(back - z).max() < 5e-6


@pytest.mark.parametrize("n,four", [(16, False), (64, True), (256, False), (2048, False), (4096, True),
                                    (8192, False), (32768, False), (65536, False), (1 << 17, False)])
def test_spec_fft_vs_radix2_and_float64(oracle, n, four):
    """The SPEC dataflows (F.2 Stockham, F.3 four-step, F.4 256x256 fft16) against the independent
    radix-2 transform and numpy float64: same transform to float32 accuracy."""
    rng = np.random.default_rng(n + four)
    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    S = oracle.spec_fft(z, four=four)
    R2 = oracle.fft(z)
    T = np.fft.fft(z.astype(np.complex128))
    rms = np.sqrt(np.mean(np.abs(T) ** 2))
    assert np.abs(S - T).max() <= 3e-6 * rms
    assert np.abs(S - R2).max() <= 4e-6 * rms
    assert np.sqrt(np.mean(np.abs(S - T) ** 2)) <= 3e-7 * rms


def test_windows_vs_scipy(oracle):
    for k, nm in [("hamming", "hamming"), ("hann", "hann"), ("blackmann_harris", "blackmanharris")]:
        assert np.abs(oracle.window(1024, k) - signal.get_window(nm, 1024, fftbins=False)).max() < 1e-7
    ft = oracle.window(1024, "flat_top")
    x = 2 * np.pi * np.arange(1024) / 1023
    assert np.abs(ft - (1 - 1.93 * np.cos(x) + 1.29 * np.cos(2 * x) - 0.388 * np.cos(3 * x)
                        + 0.028 * np.cos(4 * x))).max() < 1e-6
    assert np.all(oracle.window(64, "none") == 1)


def test_psd_known_answers(oracle):
    N = 8192
    x = np.zeros(N * 2, np.complex64)
    x[0] = 1
    k = 321
    x[N:] = 0.5 * np.exp(2j * np.pi * k * np.arange(N) / N)
    p = oracle.psd_frames(x, N, "none")
    assert np.allclose(p[0], 1 / N, rtol=1e-6)
    assert abs(p[1][k] - 0.25 * N) < 1e-5 * 0.25 * N
    assert abs(p[1].sum() - 0.25 * N) < 1e-5 * 0.25 * N    # Parseval
    # windowed tone: coherent gain of the window
    w = oracle.window(N, "hann").astype(np.float64)
    ph = oracle.psd_frames(x, N, "hann")
    assert abs(ph[1][k] - 0.25 * w.sum() ** 2 / N) / (0.25 * w.sum() ** 2 / N) < 1e-4


def test_psd_shift_db_and_averager(oracle):
    N = 1024
    rng = np.random.default_rng(1)
    p = rng.random(N).astype(np.float32) + 1e-3
    q = p.copy()
    oracle.lib().sdo_psd_shift_db(oracle.ptr(q), N)
    ref = 10 * np.log10(np.fft.fftshift(p).astype(np.float64) + 1e-8)
    assert np.abs(q - ref).max() < 1e-4
    last = q.copy()
    frame = (q + 1).astype(np.float32)
    oracle.lib().sdo_averager_feed(oracle.ptr(last), oracle.ptr(frame), N, C.c_float(0.25))
    assert np.allclose(last, q + 0.25, atol=1e-5)


# ---------------------------------------------------------------- SPEC I: filters ----------------
def test_butterworth_vs_scipy(oracle):
    L = oracle.lib()
    for order, fc in [(1, 0.2), (2, 0.3), (2, 0.64), (4, 0.05), (5, 0.1)]:
        b = (C.c_float * (order + 1))()
        a = (C.c_float * (order + 1))()
        assert L.sdo_butter_lp(order, fc, b, a) == 0
        sb, sa = signal.butter(order, fc)
        assert np.abs(np.array(b[:]) - sb).max() < 1e-6
        assert np.abs(np.array(a[:]) - sa).max() < 1e-6


def test_filter_feed_vs_lfilter(oracle):
    L = oracle.lib()
    sb, sa = signal.butter(2, 0.3)
    b = (C.c_float * 3)(*sb)
    a = (C.c_float * 3)(*sa)
    f = oracle.Filt()
    assert L.sdo_filt_init(C.byref(f), 3, a, 3, b) == 0
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(500) + 1j * rng.standard_normal(500)).astype(np.complex64)
    y = np.array([complex(r.re, r.im) for r in (L.sdo_filt_feed(C.byref(f), oracle.Cpx(float(v.real), float(v.imag)))
                                                for v in x)])
    ref = signal.lfilter(np.float32(sb).astype(np.float64), np.float32(sa).astype(np.float64), x.astype(np.complex128))
    assert np.abs(y - ref).max() < 1e-5


def test_rrc_taps_properties(oracle):
    T, beta = 3.125, 0.35
    n = oracle.lib().sdo_mf_span(T)
    assert n == 19
    h = np.empty(n, np.float32)
    oracle.lib().sdo_taps_rrc(h.ctypes.data_as(oracle.c_float_p), n, T, beta)
    assert abs(h.sum() - 1.0) < 0.06            # unity DC gain up to truncation
    assert np.argmax(h) in (n // 2, n // 2 + 1)
    # symmetric about n/2 (sample n/2 is the centre for even spans; here centre between taps)
    t = (np.arange(n) - n / 2) / T
    assert np.all(h[np.abs(t) < 0.5] > 0.1)


# ---------------------------------------------------------------- SPEC N/C/G: loops ---------------
def test_ncqo_phase_continuity_and_xlate(oracle):
    L = oracle.lib()
    n = 5000
    x = np.ones(n, np.complex64)
    o = oracle.Ncqo()
    L.sdo_ncqo_init(C.byref(o), C.c_float(0.01))
    y = np.empty_like(x)
    # feed in ragged blocks: identical to one straight loop
    o2 = oracle.Ncqo()
    L.sdo_ncqo_init(C.byref(o2), C.c_float(0.01))
    y2 = np.empty_like(x)
    L.sdo_carrier_xlate(oracle.ptr(x), oracle.ptr(y), n, C.byref(o))
    off = 0
    for m in (1, 7, 4096, 896):
        L.sdo_carrier_xlate(oracle.ptr(x[off:off + m]), oracle.ptr(y2[off:off + m]), m, C.byref(o2))
        off += m
    assert np.array_equal(y.view(np.uint32), y2.view(np.uint32))
    ref = np.exp(1j * np.pi * 0.01 * np.arange(n))
    assert np.abs(y - ref).max() < 2e-3        # float32 phase accumulation drift only


@pytest.mark.parametrize("kind,order", [(1, 2), (2, 4), (3, 8)])
def test_costas_pulls_in_offset(oracle, kind, order):
    L = oracle.lib()
    n = 20000
    s, _ = synth.psk_signal(n, 4.0, order=order, seed=kind)
    x = synth.mix(s, 2e-4, 0.7).astype(np.complex64)     # positive frequency offset must be pulled in
    c = oracle.Costas()
    assert L.sdo_costas_init(C.byref(c), kind, 0.0, 0.5, 3, 4e-3) == 0
    for v in x:
        L.sdo_costas_feed(C.byref(c), oracle.Cpx(float(v.real), float(v.imag)))
    # NCO angular frequency converges to 2 pi * offset
    assert abs(c.ncqo.omega - 2 * np.pi * 2e-4) < 2e-4
    assert c.lock > 0.5


def test_pll_tracks_carrier(oracle):
    L = oracle.lib()
    n = 8000
    x = np.exp(1j * (2 * np.pi * 1e-3 * np.arange(n) + 1.0)).astype(np.complex64)
    p = oracle.Pll()
    L.sdo_pll_init(C.byref(p), 0.0, 0.01)
    out = [L.sdo_pll_track(C.byref(p), oracle.Cpx(float(v.real), float(v.imag))) for v in x]
    tail = np.array([complex(r.re, r.im) for r in out[-500:]])
    assert np.abs(np.angle(tail)).max() < 0.05 and abs(p.ncqo.omega - 2 * np.pi * 1e-3) < 1e-5


def test_agc_levels_and_delay(oracle):
    L = oracle.lib()
    ap = oracle.AgcParams()
    L.sdo_agc_params_from_tau(C.byref(ap), 10.0, 1.0)
    assert ap.delay_line_size == 39 and ap.mag_history_size == 39 and ap.hang_max == 19
    a = oracle.Agc()
    assert L.sdo_agc_init(C.byref(a), C.byref(ap)) == 0
    out = []
    for i in range(3000):
        amp = 0.01 if i < 1500 else 0.3
        out.append(L.sdo_agc_feed(C.byref(a), oracle.Cpx(amp, 0.0)))
    y = np.array([r.re for r in out])
    # 30 dB input step -> compressed to < 2 dB at the output (slope 0.06), after the loop settles
    lo, hi = y[1400], y[2900]
    assert 0 < lo < hi and 20 * np.log10(hi / lo) < 2.5
    assert y[:39].max() == 0          # look-ahead delay line


def test_gardner_recovers_prbs_exactly(oracle):
    """transmit PRBS -> RRC QPSK at non-integer sps -> oracle psk inspector -> exact bits."""
    fs, sps, n = 1.0, 3.125, 60000
    s, idx = synth.psk_signal(n, sps, order=4, seed=5)
    rng = np.random.default_rng(5)
    x = (0.25 * synth.mix(s, 1e-4, 0.2) + synth.awgn(n, 10 ** (-45 / 20), rng)).astype(np.complex64)
    cfg = oracle.insp_config("psk", fs, baud=fs / sps, costas_order=2, bits_per_symbol=2, loop_bw=fs * 2e-3,
                             mf_type=1, mf_rolloff=0.35, clock_type=1, clock_gain=0.1)
    soft, hard = oracle.inspector_run(cfg, x)
    h = hard[2000:12000].astype(int)
    best = min((np.count_nonzero(((h + rot) % 4) != idx[lag:lag + len(h)]), lag, rot)
               for lag in range(1990, 2030) for rot in range(4))
    assert best[0] == 0, best
    # block-size independence of the whole chain
    soft2, hard2 = oracle.inspector_run(cfg, x, chunk=777)
    assert np.array_equal(soft.view(np.uint32), soft2.view(np.uint32)) and np.array_equal(hard, hard2)


def test_cma_equalizer_restores_constant_modulus(oracle):
    """SPEC E: two-ray multipath QPSK; CMA shrinks the modulus spread and settles on |y| = 1 (0.75 after
    the -2.5 dB output gain); eq_locked freezes the identity weights, i.e. passes the symbols through."""
    fs, sps, n = 1.0, 4.0, 160000
    s, _ = synth.psk_signal(n, sps, order=4, seed=8)
    rng = np.random.default_rng(8)
    x = 0.25 * (s + 0.3 * np.exp(0.9j) * np.roll(s, 4)) + synth.awgn(n, 10 ** (-50 / 20), rng)
    x = x.astype(np.complex64)
    kw = dict(baud=fs / sps, costas_order=2, bits_per_symbol=2, loop_bw=fs * 1e-3, mf_type=1, clock_type=1,
              clock_gain=0.05)
    off, _ = oracle.inspector_run(oracle.insp_config("psk", fs, **kw), x)
    on, _ = oracle.inspector_run(oracle.insp_config("psk", fs, eq_type=1, eq_rate=5e-3, **kw), x)
    lock, _ = oracle.inspector_run(oracle.insp_config("psk", fs, eq_type=1, eq_rate=5e-3, eq_locked=1, **kw), x)
    assert len(on) == len(off) == len(lock)
    assert np.array_equal(lock.view(np.uint32), off.view(np.uint32))
    t = len(on) * 2 // 3
    spread_off, spread_on = np.std(np.abs(off[t:])), np.std(np.abs(on[t:]))
    assert spread_on < 0.6 * spread_off, (spread_on, spread_off)
    # CMA's fixed point has E|y|^4 = E|y|^2, i.e. |y| ~ 1 before the 0.75 gain
    m2, m4 = np.mean(np.abs(on[t:] / 0.75) ** 2), np.mean(np.abs(on[t:] / 0.75) ** 4)
    assert abs(m4 / m2 - 1.0) < 0.05


def test_decider_intervals(oracle):
    L = oracle.lib()
    d = oracle.Decider()
    L.sdo_decider_init(C.byref(d), 0, 2, C.c_float(-np.pi), C.c_float(np.pi))
    ang = np.array([-3.0, -1.0, 0.5, 2.0, np.pi])
    x = np.exp(1j * ang).astype(np.complex64)
    sym = np.empty(len(x), np.uint8)
    L.sdo_decider_decide(C.byref(d), oracle.ptr(x), oracle.ptr(sym), len(x))
    assert list(sym) == [0, 1, 2, 3, 3]
    L.sdo_decider_init(C.byref(d), 1, 1, C.c_float(0), C.c_float(1))
    x = np.array([0.1, 0.6, 1.7], np.complex64)
    sym = np.empty(3, np.uint8)
    L.sdo_decider_decide(C.byref(d), oracle.ptr(x), oracle.ptr(sym), 3)
    assert list(sym) == [0, 1, 1]


def test_quad_demod_matches_reference_formula(oracle):
    n = 1000
    f = 0.05
    x = np.exp(2j * np.pi * f * np.arange(n)).astype(np.complex64)
    y = np.empty_like(x)
    prev, primed = oracle.Cpx(0, 0), C.c_int(0)
    oracle.lib().sdo_quad_demod(oracle.ptr(x), oracle.ptr(y), n, C.byref(prev), C.byref(primed))
    assert y[0] == 0
    assert np.allclose(y[1:].imag, 2 * f, atol=1e-5) and np.all(y.real == 0)   # +-1 at +-fs/2


# ---------------------------------------------------------------- SPEC S: spectral tuner ----------
def test_specttuner_geometry(oracle):
    # cfg2 of SURVEY.md section 8: 3 MHz at 100 MS/s in a 65536 window -> 2048-pt IFFT, D = 32
    c, s, w = oracle.channel_geometry(65536, 2 * np.pi * 0.125, 2 * np.pi * 0.03, 1.0)
    assert (c, s, w) == (8192, 2048, 1967)
    assert c % 2 == 0
    # LPFTask trick: guard = 2 pi / bw -> no decimation (Tasks/LPFTask.cpp:65)
    bw = np.float32(np.pi * 0.2)
    c, s, w = oracle.channel_geometry(4096, 0.0, float(bw), float(np.float32(2 * np.pi) / bw))
    assert (c, s) == (0, 4096) and abs(w - 4096 * 0.1) <= 1


def test_specttuner_tone_gain_latency_and_blocks(oracle):
    W, hops = 4096, 20
    n = W // 2 * hops
    f = 0.125
    x = (0.5 * np.exp(2j * np.pi * f * np.arange(n))).astype(np.complex64)
    ch = dict(f0=2 * np.pi * f, bw=2 * np.pi * 0.03, guard=1.0)
    y = oracle.specttuner_run(x, W, [ch])[0]
    c, s, w = oracle.channel_geometry(W, ch["f0"], ch["bw"], 1.0)
    assert len(y) == (hops - 1) * s // 2          # latency of half a window
    assert np.abs(np.abs(y[s:]) - 0.5).max() < 1e-3   # unity pass-band gain, tone lands on DC
    y2 = oracle.specttuner_run(x, W, [ch], chunk=1000)[0]
    assert np.array_equal(y.view(np.uint32), y2.view(np.uint32))


def test_specttuner_lpf_equals_fft_brickwall(oracle):
    """LPFTask contract: with guard = 2 pi / bw the tuner is a (smoothed) FFT-domain low-pass."""
    W = 4096
    n = W * 8
    rng = np.random.default_rng(4)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    bw = np.float32(np.pi * 0.25)
    y = oracle.specttuner_run(np.concatenate([x, np.zeros(W, np.complex64)]), W,
                              [dict(f0=0.0, bw=float(bw), guard=float(np.float32(2 * np.pi) / bw))])[0][:n]
    X = np.fft.fft(x[W:W * 5].astype(np.complex128))
    Y = np.fft.fft(y[W:W * 5].astype(np.complex128))
    f = np.fft.fftfreq(W * 4)
    inb, outb = np.abs(f) < 0.05, np.abs(f) > 0.075
    assert abs(np.sum(np.abs(Y[inb]) ** 2) / np.sum(np.abs(X[inb]) ** 2) - 1) < 0.02
    assert np.sum(np.abs(Y[outb]) ** 2) < 1e-4 * np.sum(np.abs(X[outb]) ** 2)


# ---------------------------------------------------------------- SPEC V: SpectrumView ------------
def _sview_literal(fmin, fmax, feeds, rel_bw=0.5):
    """Literal Python transcription of Panoramic/Scanner.cpp:56-256 (float32/float64 as in the C++)."""
    f32 = np.float32
    rng_ = fmax - fmin
    size = 1
    while size < int(rng_ / 1000.0):
        size <<= 1
    size = min(size, 65536)
    psd = np.zeros(65536, f32)
    acc = np.zeros(65536, f32)
    cnt = np.zeros(65536, f32)
    for feed in feeds:
        if isinstance(feed[0], str):                     # ("view", psdAccum, psdCount, freqMin, freqMax) of another
            _, data, count, lo, hi = feed                # view: SpectrumView::feed(SpectrumView const &), :276-286
            adjust = False
        else:
            data, center, fftbw = feed
            count, adjust = None, True
            lo, hi = center - fftbw / 2, center + fftbw / 2
        psize = len(data)
        inp_bw = hi - lo
        assert (hi - lo) / rng_ * size >= 2              # linear mode only
        skip = int(f32(0.5) * (f32(1) - f32(rel_bw)) * f32(psize)) if adjust else 0
        fskip = float(skip) / psize * inp_bw
        bw = inp_bw - 2 * fskip
        fft_count = rng_ / bw
        bins = size / fft_count
        sw, dw = inp_bw / psize, rng_ / size
        delta = dw / sw
        pos = (fskip + lo - fmin) / rng_ * size
        j = int(pos) if pos > 0 else 0
        k = int(pos + bins) if pos + bins < size else size
        while j < k:
            fj = fmin + dw * j
            sb = (fj - lo) / sw
            a, b = int(sb), int(sb + delta)
            a = min(max(a, 0), psize - 1)
            b = min(max(b, a + 1), psize)
            s = f32(0)
            c = f32(0)
            for i in range(a, b):
                s = f32(s + data[i])
                c = f32(c + (f32(1) if count is None else count[i]))
            if c > 0:
                acc[j] = f32(acc[j] + f32(s / c))
                cnt[j] = f32(cnt[j] + f32(1))
            j += 1
        # interpolate
        in_gap, first, left, count, zp = False, True, f32(-200.0), 1, 0
        for i in range(size):
            empty = cnt[i] <= f32(0.5)
            if not in_gap:
                if empty:
                    in_gap, zp, count = True, i, 1
                    first = i == 0
                    if not first:
                        left = psd[i - 1]
                else:
                    psd[i] = f32(acc[i] / cnt[i])
                    if cnt[i] > f32(5):
                        cnt[i] = f32(1)
                        acc[i] = f32(psd[i] * f32(1))
            elif empty:
                count += 1
            else:
                in_gap = False
                right = psd[i] = f32(acc[i] / cnt[i])
                for jj in range(count):
                    if first:
                        psd[jj + zp] = right
                    else:
                        t = f32(f32(jj + f32(0.5)) / f32(count))
                        psd[jj + zp] = f32(f32(f32(1) - t) * left + f32(t * right))
        if in_gap:
            for jj in range(count):
                psd[jj + zp] = left
    return psd[:size], acc[:size], cnt[:size], size


def test_spectrum_view_matches_literal_transcription(oracle):
    L = oracle.lib()
    fmin, fmax, fftbw, psize = 100e6, 140e6, 2e6, 2048
    rng = np.random.default_rng(7)
    feeds = []
    for hop in range(60):
        center = fmin + 0.5e6 + hop * 0.66e6
        feeds.append(((rng.random(psize).astype(np.float32) * 10 - 90), center, fftbw))
    feeds += feeds[10:30]          # revisits exercise the count > 5 forgetting rule
    feeds += feeds[10:30] * 4
    v = oracle.SpectrumView()
    assert L.sdo_sview_init(C.byref(v)) == 0
    L.sdo_sview_set_range(C.byref(v), fmin, fmax)
    v.fft_bandwidth = fftbw
    for data, center, _ in feeds:
        L.sdo_sview_feed(C.byref(v), oracle.ptr(data), None, psize, center, 1)
    psd, acc, cnt, size = _sview_literal(fmin, fmax, feeds)
    assert v.spectrum_size == size
    got = np.ctypeslib.as_array(v.psd, shape=(65536,))[:size]
    gcnt = np.ctypeslib.as_array(v.psd_count, shape=(65536,))[:size]
    assert np.array_equal(gcnt, cnt)
    assert np.array_equal(got.view(np.uint32), psd.view(np.uint32))
    L.sdo_sview_free(C.byref(v))


def _zoom_case():
    """Scanner::setViewRange (Panoramic/Scanner.cpp:444-487): sweep a wide range, zoom into a part of it -- the new
    view is seeded with feed(previous view) -- keep sweeping there, then zoom out again."""
    wide, narrow, fftbw, psize = (400e6, 460e6), (417.3e6, 431.9e6), 2e6, 1024
    rng = np.random.default_rng(23)

    def hops(lo, hi, n):
        cs = lo + (hi - lo) * (np.arange(n) + 0.5) / n
        return [((rng.random(psize).astype(np.float32) * 20 - 100), float(c), fftbw) for c in cs]
    return wide, narrow, fftbw, psize, hops(*wide, 45), hops(*narrow, 25), hops(*wide, 10)


def test_spectrum_view_zoom_feeds_previous_view(oracle):
    """SpectrumView::feed(SpectrumView const &) (Scanner.cpp:276-286) against the literal transcription, both ways"""
    L = oracle.lib()
    L.sdo_sview_feed_view.argtypes = [C.c_void_p, C.c_void_p]
    wide, narrow, fftbw, psize, h1, h2, h3 = _zoom_case()
    views = [oracle.SpectrumView(), oracle.SpectrumView(), oracle.SpectrumView()]
    for v, r in zip(views, (wide, narrow, wide)):
        assert L.sdo_sview_init(C.byref(v)) == 0
        L.sdo_sview_set_range(C.byref(v), *r)
        v.fft_bandwidth = fftbw

    def arr(p, n):
        return np.ctypeslib.as_array(p, shape=(65536,))[:n].copy()
    for d, c, _ in h1:
        L.sdo_sview_feed(C.byref(views[0]), oracle.ptr(d), None, psize, c, 1)
    _, a0, c0, s0 = _sview_literal(*wide, h1)
    assert views[0].spectrum_size == s0
    L.sdo_sview_feed_view(C.byref(views[1]), C.byref(views[0]))                      # zoom in
    for d, c, _ in h2:
        L.sdo_sview_feed(C.byref(views[1]), oracle.ptr(d), None, psize, c, 1)
    p1, a1, c1, s1 = _sview_literal(*narrow, [("view", a0, c0, wide[0], wide[1])] + h2)
    assert views[1].spectrum_size == s1 and s1 < s0
    assert np.array_equal(arr(views[1].psd_count, s1), c1)
    assert np.array_equal(arr(views[1].psd, s1).view(np.uint32), p1.view(np.uint32))
    assert np.array_equal(arr(views[1].psd_accum, s1).view(np.uint32), a1.view(np.uint32))
    L.sdo_sview_feed_view(C.byref(views[2]), C.byref(views[1]))                      # zoom out: detail into wide
    for d, c, _ in h3:
        L.sdo_sview_feed(C.byref(views[2]), oracle.ptr(d), None, psize, c, 1)
    p2, a2, c2, s2 = _sview_literal(*wide, [("view", a1, c1, narrow[0], narrow[1])] + h3)
    assert np.array_equal(arr(views[2].psd_count, s2), c2)
    assert np.array_equal(arr(views[2].psd, s2).view(np.uint32), p2.view(np.uint32))
    inside = (np.arange(s2) * (wide[1] - wide[0]) / s2 + wide[0] > narrow[0] + 1e6) & \
             (np.arange(s2) * (wide[1] - wide[0]) / s2 + wide[0] < narrow[1] - 1e6)
    assert np.all(c2[inside] >= 1)                                                   # the detail landed there
    for v in views:
        L.sdo_sview_free(C.byref(v))


# ---------------------------------------------------------------- C-ABI surface --------------------
def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: the header must compile as C99 (-pedantic) and as C++11, nothing but
    <stddef.h> / <stdint.h> / <sys/time.h> types in the signatures"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "sigdigger_b200.h"\nint main(void) { sdb_engine_params p; (void) p; return 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror"], ["g++", "-std=c++11", "-Wall", "-x", "c++"]):
        r = subprocess.run(cmd + ["-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "sigdigger_b200.h")).read(), flags=re.S)   # code only
    assert "torch::" not in text and "cudaStream_t" not in text and "at::Tensor" not in text


def test_library_exports_every_declared_symbol():
    import sigdigger_b200 as sdb
    hdr = open(os.path.join(ROOT, "include", "sigdigger_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(sdb.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(sdb.EXPORTED_SYMBOLS), declared ^ set(sdb.EXPORTED_SYMBOLS)
    sdb.load_library()


def test_no_device_fails_loudly():
    import sigdigger_b200 as sdb
    if sdb.device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(sdb.SdbError, match="no CUDA device"):
        sdb.Engine(n_streams=1, psd_size=8192, max_feed=8192)
    with pytest.raises(sdb.SdbError, match="no CUDA device"):
        sdb.quad_demod(np.zeros(16, np.complex64))


def test_spectrum_sources_and_baud_estimators(oracle):
    """SPEC U on a QPSK capture at 8 samples/symbol with a 0.01 cycles/sample carrier offset."""
    n, sps = 6000, 8.0
    s, _ = synth.psk_signal(n, sps, order=4, seed=3)
    rng = np.random.default_rng(3)
    x = (0.3 * synth.mix(s, 0.01, 0.3) + synth.awgn(n, 1e-3, rng)).astype(np.complex64)
    ns = 1024
    w = signal.get_window("blackmanharris", ns, fftbins=False)
    fr = x[-ns:].astype(np.complex128)
    prev = x[-ns - 1:-1].astype(np.complex128)

    def spec(p):
        return np.abs(np.fft.fft(p * w)) ** 2 / ns

    refs = {"psd": spec(fr), "cyclo": spec(fr * np.conj(prev)), "timediff": spec(fr - prev),
            "abstimediff": spec(np.abs(fr - prev)), "exp_2": spec(fr ** 2), "exp_4": spec(fr ** 4),
            "exp_8": spec(fr ** 8), "fmspect": spec(np.angle(fr * np.conj(prev)) / np.pi)}
    for kind, ref in refs.items():
        got = oracle.spectsrc_frame(kind, ns, x)
        assert got is not None and len(got) == ns
        assert np.max(np.abs(got - ref)) <= 2e-6 * ref.max(), kind
    fac = oracle.spectsrc_frame("fac", ns, x)
    r = np.abs(np.fft.ifft(np.abs(np.fft.fft(fr)) ** 2))[:ns // 2]      # FACTab.cpp:209-221 up to the 1/ns scale
    assert len(fac) == ns // 2 and np.max(np.abs(fac - r)) <= 2e-6 * r.max()
    # the 4th power of QPSK has a line at 4x the carrier offset; the estimators find the symbol rate
    e4 = oracle.spectsrc_frame("exp_4", 4096, x)
    assert int(np.argmax(e4)) == round(4 * 0.01 * 4096)
    for est in ("baud-fac", "baud-nonlinear"):
        for m in (1024, 4096):
            assert oracle.estimate_baud(est, m, 1.0, x) == pytest.approx(1 / sps, rel=1e-6)
    assert oracle.spectsrc_frame("psd", ns, x[:ns]) is None             # needs ns + 1 samples
    assert oracle.estimate_baud("baud-fac", ns, 1.0, x[:ns]) is None


def test_channel_detector_known_answers(oracle):
    """SPEC K on synthetic spectra: averaging, exact quartile noise floor, runs, width filter, edges."""
    N = 1024
    psd = np.ones((1, N), np.float32)
    psd[0, 100:110] = 50.0            # unshifted bins 100..109 -> ascending-frequency bins 612..621
    psd[0, 700:702] = 50.0            # 2 bins wide: below min_bins = 3
    psd[0, N // 2 - 5:N // 2 + 4] = 20.0   # wraps the ascending-frequency edges: [1019, 1024) and [0, 4)
    d = oracle.ChannelDetector(N, alpha=0.5, gamma=0.5, snr=4.0, min_bins=3)
    ch, total = d.feed(psd)
    assert total == 3 and [(c[0], c[1]) for c in ch] == [(0, 4), (612, 622), (1019, 1024)]
    assert [c[2] for c in ch] == [20.0, 50.0, 20.0] and all(c[3] == 1.0 for c in ch)
    assert ch[1][4] == 50.0
    # second update with the carrier gone: the average halves its excess, the floor stays
    psd2 = np.ones((1, N), np.float32)
    ch2, _ = d.feed(psd2)
    assert (ch2[1][0], ch2[1][1], ch2[1][2]) == (612, 622, 25.5)
    # noise floor = (N/4)-th smallest averaged bin, smoothed by gamma
    rng = np.random.default_rng(0)
    p3 = rng.exponential(1.0, (1, N)).astype(np.float32)
    d3 = oracle.ChannelDetector(N, 1.0, 0.25, 1e9, 1)
    d3.feed(p3)
    assert d3.d.n0 == np.sort(p3[0])[N // 4]
    p4 = (p3 * 3).astype(np.float32)
    d3.feed(p4)
    n0 = np.float32(np.sort(p3[0])[N // 4])
    assert d3.d.n0 == np.float32(n0 + np.float32(0.25) * (np.float32(np.sort(p4[0])[N // 4]) - n0))
    for x in (d, d3):
        x.close()


def test_fast_transform_of_the_cpu_baseline_matches_the_spec_transform(oracle):
    """oracle/fft_fast.c (speed leg of bench.py's CPU arm) against numpy float64 and the SPEC transform: equal to
    float32 rounding, forward and inverse, with and without a window.  It is never used for parity."""
    import ctypes as C
    L = oracle.lib()
    L.sdo_fast_fft.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_int]
    rng = np.random.default_rng(5)
    for n in (8, 32, 1024, 2048, 65536):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        w = oracle.window(n, "blackmann_harris")
        y = np.empty_like(x)
        L.sdo_fast_fft(x.ctypes.data, w.ctypes.data, y.ctypes.data, n, -1)
        ref = np.fft.fft(x.astype(np.complex128) * w)
        assert np.max(np.abs(y - ref)) <= 2e-6 * np.max(np.abs(ref))
        z = np.empty_like(x)
        L.sdo_fast_fft(y.ctypes.data, None, z.ctypes.data, n, +1)
        assert np.max(np.abs(z / n - x * w)) <= 5e-6 * np.max(np.abs(x))
    # PSD of the SPEC transform and of the fast one agree to the parity tolerance of tests/parity.py (1e-5 of the peak)
    n = 65536
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    spec = oracle.psd_frames(x, n, "blackmann_harris")[0]
    y = np.empty_like(x)
    w = oracle.window(n, "blackmann_harris")
    L.sdo_fast_fft(x.ctypes.data, w.ctypes.data, y.ctypes.data, n, -1)
    fast = (y.real.astype(np.float64) ** 2 + y.imag.astype(np.float64) ** 2) / n
    assert np.max(np.abs(fast - spec)) <= 1e-5 * np.max(spec)

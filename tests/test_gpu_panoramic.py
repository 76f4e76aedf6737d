"""-m gpu: panoramic scanner path (BASELINE.json configs[4]): per-hop 65536-pt PSD in the PSDMessage layout
(fft-shift + dB) -> SpectrumView stitch.  The CUDA project + accumulate must reproduce the reference's
feed() / interpolate() sequence (Panoramic/Scanner.cpp:239-256) value by value."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hops(n_hops, N, fs, seed):
    rng = np.random.default_rng(seed)
    x = (0.02 * (rng.standard_normal((n_hops, N)) + 1j * rng.standard_normal((n_hops, N)))).astype(np.complex64)
    t = np.arange(N)
    for h in range(n_hops):
        for k in range(3):
            f = rng.uniform(-0.2, 0.2)
            x[h] += (rng.uniform(0.05, 0.4) * np.exp(2j * np.pi * f * t)).astype(np.complex64)
    return x


def _oracle_view(oracle, psd_db, centers, fmin, fmax, fftbw, rel_bw):
    L = oracle.lib()
    v = oracle.SpectrumView()
    assert L.sdo_sview_init(C.byref(v)) == 0
    L.sdo_sview_set_range(C.byref(v), fmin, fmax)
    v.fft_bandwidth = fftbw
    v.fft_rel_bw = rel_bw
    for h, c in enumerate(centers):
        L.sdo_sview_feed(C.byref(v), oracle.ptr(psd_db[h]), None, psd_db.shape[1], float(c), 1)
    n = v.spectrum_size
    out = tuple(np.ctypeslib.as_array(p, shape=(65536,))[:n].copy() for p in (v.psd, v.psd_accum, v.psd_count))
    L.sdo_sview_free(C.byref(v))
    return out


@pytest.mark.parametrize("N,n_hops,revisit", [(65536, 24, True), (8192, 40, True), (4096, 16, False)])
def test_panoramic_sweep_matches_reference_sequence(sdb, oracle, N, n_hops, revisit):
    import torch
    fs = 100e6
    fftbw = fs
    rel_bw = 0.5
    fmin, fmax = 1.0e9, 1.0e9 + n_hops * fs * rel_bw * 0.9
    centers = fmin + fs * rel_bw * (0.4 + 0.9 * np.arange(n_hops))
    order = list(range(n_hops))
    if revisit:                       # exercise the count > 5 forgetting rule and out-of-order hops
        order += order[3:9] * 6 + order[::-1][:5]
    x = _hops(n_hops, N, fs, seed=N % 13)
    xs = np.stack([x[h] for h in order])
    cs = centers[order]
    # oracle: SPEC PSD + GUI shift/dB, then the literal feed sequence
    psd_db = np.stack([oracle.psd_frames(xs[h], N, "blackmann_harris")[0] for h in range(len(order))])
    for h in range(len(order)):
        oracle.lib().sdo_psd_shift_db(oracle.ptr(psd_db[h]), N)
    ref = _oracle_view(oracle, psd_db, cs, fmin, fmax, fftbw, rel_bw)
    # CUDA: hop PSDs with the fused shift/dB epilogue, then project + accumulate
    e = sdb.Engine(n_streams=len(order), psd_size=N, psd_window="blackmann_harris", max_feed=N,
                   flags=sdb.FLAG_PSD_SHIFT_DB)
    e.commit()
    e.feed(xs)
    got_db = e.read_psd()[:, 0, :]
    assert np.array_equal(got_db.view(np.uint32), psd_db.view(np.uint32)), "hop PSDs (dB) not bit-identical"
    v = sdb.SpectrumView(fmin, fmax, fftbw, rel_bw)
    v.project(e.psd_device_ptr, N, cs)
    v.accumulate()
    psd, acc, cnt = v.read()
    assert len(psd) == len(ref[0])
    assert np.array_equal(cnt, ref[2])
    assert np.array_equal(acc.view(np.uint32), ref[1].view(np.uint32))
    assert np.array_equal(psd.view(np.uint32), ref[0].view(np.uint32))
    # split in two accumulate() calls == one (what the multi-GPU gather does)
    v2 = sdb.SpectrumView(fmin, fmax, fftbw, rel_bw)
    v2.project(e.psd_device_ptr, N, cs)
    j0, nb, va, vc = v2.contrib_ptrs()
    half = len(order) // 2
    mb = v2.max_bins
    v2.accumulate(j0, nb, va, vc, half)
    v2.accumulate(j0 + 4 * half, nb + 4 * half, va + 4 * half * mb, vc + 4 * half * mb, len(order) - half)
    psd2, acc2, cnt2 = v2.read()
    assert np.array_equal(acc2.view(np.uint32), acc.view(np.uint32)) and np.array_equal(cnt2, cnt)
    assert np.array_equal(psd2.view(np.uint32), psd.view(np.uint32))
    del torch


def test_panoramic_sweep_driver_single_rank(sdb, oracle):
    """sigdigger_b200.panoramic.sweep() with world = 1 (the multi-GPU form is tests/test_gpu_multi.py)."""
    import torch
    from sigdigger_b200 import panoramic
    N, n_hops, fs, rel_bw = 16384, 21, 100e6, 0.5
    fmin, fmax = 3.0e9, 3.0e9 + n_hops * fs * rel_bw
    centers = fmin + fs * rel_bw * (0.5 + np.arange(n_hops))
    x = _hops(n_hops, N, fs, seed=2)
    psd, acc, cnt = panoramic.sweep(sdb, torch, None, torch.from_numpy(x).cuda(), centers, N, "hann", (fmin, fmax),
                                    fs, rel_bw)
    psd_db = np.stack([oracle.psd_frames(x[h], N, "hann")[0] for h in range(n_hops)])
    for h in range(n_hops):
        oracle.lib().sdo_psd_shift_db(oracle.ptr(psd_db[h]), N)
    ref = _oracle_view(oracle, psd_db, centers, fmin, fmax, fs, rel_bw)
    assert np.array_equal(cnt, ref[2]) and np.array_equal(acc.view(np.uint32), ref[1].view(np.uint32))
    assert np.array_equal(psd.view(np.uint32), ref[0].view(np.uint32))


def test_psd_shift_db_pass_equals_fused_epilogue(sdb):
    """sdb_psd_shift_db_device on a linear-PSD engine == the engine's SDB_FLAG_PSD_SHIFT_DB output, bit for bit"""
    import torch
    for N, S, frames in ((65536, 3, 2), (1024, 5, 4)):
        x = _hops(S, N * frames, 1.0, seed=N % 13)
        outs = []
        for flags in (0, sdb.FLAG_PSD_SHIFT_DB):
            e = sdb.Engine(n_streams=S, psd_size=N, psd_window="hann", max_feed=N * frames, flags=flags)
            e.commit()
            e.feed(x)
            outs.append(e.read_psd())
            if flags == 0:
                db = torch.empty((S, frames, N), dtype=torch.float32, device="cuda")
                sdb.psd_shift_db(e.psd_device_ptr, db.data_ptr(), S * frames, N)
                torch.cuda.synchronize()
                got = db.cpu().numpy()
            e.close()
        assert np.array_equal(got.view(np.uint32), outs[1].view(np.uint32))
        assert not np.array_equal(outs[0], outs[1])
    with pytest.raises(sdb.SdbError):
        sdb.psd_shift_db(db.data_ptr(), db.data_ptr(), 1, 1024)
    with pytest.raises(sdb.SdbError):
        sdb.psd_shift_db(db.data_ptr(), db.data_ptr() + 4096, 1, 1000)


def test_spectrum_averager_bit_exact(sdb, oracle):
    """Misc/Averager.cpp on the device: engine PSD (PSDMessage layout) -> per-stream EMA, frame after frame"""
    N, S, frames, feeds = 4096, 3, 5, 3
    x = _hops(S, N * frames * feeds, 1.0, seed=21)
    e = sdb.Engine(n_streams=S, psd_size=N, psd_window="blackmann_harris", max_feed=N * frames,
                   flags=sdb.FLAG_PSD_SHIFT_DB)
    e.commit()
    avg = sdb.Averager(N, S, alpha=0.25)
    ref = np.zeros((S, N), np.float32)
    L = oracle.lib()
    primed = False
    for f in range(feeds):
        if f == 2:
            avg.set_alpha(0.05)
        alpha = 0.25 if f < 2 else 0.05
        e.feed(x[:, f * N * frames:(f + 1) * N * frames])
        avg.feed_ptr(e.psd_device_ptr, frames)
        psd = e.read_psd()
        for s in range(S):
            for k in range(frames):
                # the first frame is copied (Averager.cpp:30-38), like alpha >= 1
                L.sdo_averager_feed(oracle.ptr(ref[s]), oracle.ptr(np.ascontiguousarray(psd[s, k])), N,
                                    C.c_float(alpha if (primed or k > 0) else 1.0))
        primed = True
        assert np.array_equal(avg.read().view(np.uint32), ref.view(np.uint32))
    # alpha >= 1 and reset() both copy the frame
    avg.set_alpha(1.0)
    e.feed(x[:, :N * frames])
    avg.feed_ptr(e.psd_device_ptr, frames)
    assert np.array_equal(avg.read().view(np.uint32), e.read_psd()[:, -1].view(np.uint32))
    avg.set_alpha(0.5)
    avg.reset()
    avg.feed_ptr(e.psd_device_ptr, 1, frames * N)      # first frame of every stream's row
    assert np.array_equal(avg.read().view(np.uint32), e.read_psd()[:, 0].view(np.uint32))


def test_panoramic_sweep_with_channel_detector(sdb, oracle):
    """configs[4] in full: per-hop PSD -> channel detector on the rank that owns the hop + SpectrumView stitch.
    The stitched spectrum must not change when the detector rides along (linear PSD + separate dB pass), and the
    channel lists must equal the oracle detector's run on each hop's own PSD, at absolute frequencies."""
    import torch
    from sigdigger_b200 import panoramic
    N, n_hops, frames, fs, rel_bw = 16384, 9, 3, 50e6, 0.5
    fmin, fmax = 1.0e9, 1.0e9 + n_hops * fs * rel_bw
    centers = fmin + fs * rel_bw * (0.5 + np.arange(n_hops))
    x = _hops(n_hops, N * frames, fs, seed=11)
    xt = torch.from_numpy(x).cuda()
    det = dict(alpha=0.5, gamma=0.5, snr=10.0, min_bins=2)
    psd, acc, cnt, chans = panoramic.sweep(sdb, torch, None, xt, centers, N, "hann", (fmin, fmax), fs, rel_bw,
                                           detect=det)
    # same stitch as a detector-less sweep over the last frame of every hop
    last = np.ascontiguousarray(x[:, (frames - 1) * N:])
    ref = panoramic.sweep(sdb, torch, None, torch.from_numpy(last).cuda(), centers, N, "hann", (fmin, fmax), fs,
                          rel_bw)
    for a, b in zip((psd, acc, cnt), ref):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert len(chans) == n_hops
    found = 0
    for h in range(n_hops):
        lin = oracle.psd_frames(x[h], N, "hann")
        d = oracle.ChannelDetector(N, det["alpha"], det["gamma"], det["snr"], det["min_bins"])
        want, _ = d.feed(lin)
        d.close()
        assert len(chans[h]) == len(want)
        for c, w in zip(chans[h], want):
            for a, b in ((c["S0"], w[2]), (c["N0"], w[3]), (c["snr"], w[4])):
                assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
            assert centers[h] - fs / 2 <= c["f_lo"] <= c["fc"] <= c["f_hi"] <= centers[h] + fs / 2
            assert c["fc"] == pytest.approx(0.5 * (c["f_lo"] + c["f_hi"]))
        found += len(want)
    assert found >= n_hops          # every hop carries three tones well above the noise


def test_panoramic_histogram_mode(sdb, oracle):
    """hop narrower than two destination bins -> feedHistogramMode (Scanner.cpp:187-237)"""
    N, n_hops = 4096, 30
    fftbw = 1e6
    fmin, fmax = 100e6, 100e6 + 65536 * 1000.0 * 4          # 4 kHz bins, hops of 1 MHz * ... wide range
    fmax = fmin + 65536 * 1.2e6                               # destination bin 1.2 MHz > hop width
    rng = np.random.default_rng(3)
    cs = fmin + rng.uniform(0.01, 0.99, n_hops) * (fmax - fmin)
    x = _hops(n_hops, N, fftbw, seed=5)
    psd_db = np.stack([oracle.psd_frames(x[h], N, "hann")[0] for h in range(n_hops)])
    for h in range(n_hops):
        oracle.lib().sdo_psd_shift_db(oracle.ptr(psd_db[h]), N)
    ref = _oracle_view(oracle, psd_db, cs, fmin, fmax, fftbw, 0.5)
    e = sdb.Engine(n_streams=n_hops, psd_size=N, psd_window="hann", max_feed=N, flags=sdb.FLAG_PSD_SHIFT_DB)
    e.commit()
    e.feed(x)
    v = sdb.SpectrumView(fmin, fmax, fftbw, 0.5)
    v.project(e.psd_device_ptr, N, cs)
    v.accumulate()
    psd, acc, cnt = v.read()
    assert np.array_equal(cnt.view(np.uint32), ref[2].view(np.uint32))
    assert np.array_equal(acc.view(np.uint32), ref[1].view(np.uint32))
    assert np.array_equal(psd.view(np.uint32), ref[0].view(np.uint32))


@pytest.mark.parametrize("detect", [False, True])
def test_panoramic_sweep_histogram_mode_geometry(sdb, oracle, detect):
    """The sweep's stream-ordered path with hops narrower than two view bins (feedHistogramMode): the tiled projection
    does not apply, the sweep converts to dB (detector on: linear PSD + separate pass) and runs the general kernel."""
    import torch
    from sigdigger_b200 import panoramic
    N, n_hops, fftbw = 4096, 30, 1e6
    fmin = 100e6
    fmax = fmin + 65536 * 1.2e6
    rng = np.random.default_rng(3)
    cs = fmin + rng.uniform(0.01, 0.99, n_hops) * (fmax - fmin)
    x = _hops(n_hops, N, fftbw, seed=5)
    psd_db = np.stack([oracle.psd_frames(x[h], N, "hann")[0] for h in range(n_hops)])
    for h in range(n_hops):
        oracle.lib().sdo_psd_shift_db(oracle.ptr(psd_db[h]), N)
    ref = _oracle_view(oracle, psd_db, cs, fmin, fmax, fftbw, 0.5)
    det = dict(alpha=0.5, gamma=0.5, snr=10.0, min_bins=2) if detect else None
    out = panoramic.sweep(sdb, torch, None, torch.from_numpy(x).cuda(), cs, N, "hann", (fmin, fmax), fftbw, 0.5,
                          detect=det)
    for a, b in zip(out[:3], ref):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_panoramic_channel_lists_beyond_the_packed_read_back(sdb, oracle, monkeypatch):
    """The sweep reads the channel lists back packed (a few per hop); a sweep that finds more than the packed buffer
    holds is served from the full [hops][cap] array on demand: same lists either way."""
    import torch
    from sigdigger_b200 import panoramic
    N, n_hops, fs, rel_bw = 16384, 9, 50e6, 0.5
    fmin, fmax = 1.0e9, 1.0e9 + n_hops * fs * rel_bw
    centers = fmin + fs * rel_bw * (0.5 + np.arange(n_hops))
    xt = torch.from_numpy(_hops(n_hops, N, fs, seed=11)).cuda()
    det = dict(alpha=1.0, gamma=0.5, snr=10.0, min_bins=2)
    want = panoramic.sweep(sdb, torch, None, xt, centers, N, "hann", (fmin, fmax), fs, rel_bw, detect=det)[3]
    assert sum(len(c) for c in want) > 4
    monkeypatch.setenv("SDB_PANORAMIC_DENSE_CAP", "4")
    got = panoramic.sweep(sdb, torch, None, xt, centers, N, "hann", (fmin, fmax), fs, rel_bw, detect=det)[3]
    assert got == want

"""-m gpu: two device paths against the oracle (which the CPU suite pins to the compiled reference / literal
transcriptions of it):
  * SpectrumView::feed(SpectrumView const &) (Panoramic/Scanner.cpp:276-286, the zoom path of Scanner::setViewRange):
    k_sview_project_view + the accumulate / fill passes;
  * the inspector tab's SNR estimator (Misc/SNREstimator.cpp:30-169): k_snr_feed.
First hardware run: round 2 (both bit-exact as written)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


def test_snr_estimator_batch_bit_exact(sdb, oracle):
    """Misc/SNREstimator.cpp on the device (k_snr_feed, one CTA per estimator) against the oracle restatement, which
    tests/test_oracle_tasks.py checks against a literal transcription: sigma, SNR and model after every feed"""
    from test_oracle_tasks import _histogram
    length, cases = 256, [(1, 0.04, 1.0), (2, 0.03, 1.0), (3, 0.015, 0.5), (2, 0.02, 0.05)]
    hist = np.stack([_histogram(b, length, s, seed=10 + i) for i, (b, s, _) in enumerate(cases)])
    g = sdb.SnrEstimator(len(cases), length)
    refs = []
    for i, (b, _, a) in enumerate(cases):
        g.set_bps(i, b)
        g.set_alpha(i, a)
        refs.append(oracle.SnrEstimator(b, length, alpha=a))
    for it in range(6):
        if it == 3:                                           # a change of bps restarts sigma (setBps)
            g.set_bps(0, 2)
            oracle.lib().sdo_snr_set_bps(C.byref(refs[0].e), 2)
        g.feed(hist)
        sigma, snr, model = g.read(model=True)
        for i, r in enumerate(refs):
            r.feed(hist[i])
            assert np.float32(sigma[i]).view(np.uint32) == np.float32(r.sigma).view(np.uint32), (it, i)
            assert np.float32(snr[i]).view(np.uint32) == np.float32(r.snr).view(np.uint32), (it, i)
            assert np.array_equal(model[i].view(np.uint32), r.model().view(np.uint32)), (it, i)
    for r in refs:
        r.close()


def test_view_feed_zoom_in_and_out_bit_exact(sdb, oracle):
    import torch
    from test_oracle import _zoom_case
    L = oracle.lib()
    L.sdo_sview_feed_view.argtypes = [C.c_void_p, C.c_void_p]
    wide, narrow, fftbw, psize, h1, h2, h3 = _zoom_case()
    ov = [oracle.SpectrumView(), oracle.SpectrumView(), oracle.SpectrumView()]
    gv = []
    for v, r in zip(ov, (wide, narrow, wide)):
        assert L.sdo_sview_init(C.byref(v)) == 0
        L.sdo_sview_set_range(C.byref(v), *r)
        v.fft_bandwidth = fftbw
        gv.append(sdb.SpectrumView(r[0], r[1], fftbw, 0.5))

    def feed_hops(i, hops):
        for d, c, _ in hops:
            L.sdo_sview_feed(C.byref(ov[i]), oracle.ptr(d), None, psize, c, 1)
        t = torch.from_numpy(np.stack([d for d, _, _ in hops])).cuda()
        gv[i].project(t.data_ptr(), psize, [c for _, c, _ in hops])
        gv[i].accumulate()

    def same(i):
        n = ov[i].spectrum_size
        psd, acc, cnt = gv[i].read()
        assert len(psd) == n
        for got, ref in ((cnt, ov[i].psd_count), (acc, ov[i].psd_accum), (psd, ov[i].psd)):
            want = np.ctypeslib.as_array(ref, shape=(65536,))[:n]
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))

    feed_hops(0, h1)
    same(0)
    L.sdo_sview_feed_view(C.byref(ov[1]), C.byref(ov[0]))
    gv[1].feed_view(gv[0])                                  # zoom in: the new view starts from the wide one
    same(1)
    feed_hops(1, h2)
    same(1)
    L.sdo_sview_feed_view(C.byref(ov[2]), C.byref(ov[1]))
    gv[2].feed_view(gv[1])                                  # zoom out: the detail lands in a fresh wide view
    same(2)
    feed_hops(2, h3)
    same(2)
    with pytest.raises(sdb.SdbError):
        gv[0].feed_view(gv[0])
    for v in ov:
        L.sdo_sview_free(C.byref(v))

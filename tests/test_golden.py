"""Committed fixtures (tests/golden/, generator make_golden.py): the oracle against the reference-specified arithmetic
(`ref_*.npz`, literal transcriptions of Panoramic/Scanner.cpp:56-256, Tasks/QuadDemodTask.cpp:44-60,
Default/GenericInspector/FACTab.cpp:209-221, and of the TimeWindow tasks Tasks/DelayedConjTask.cpp:58-100,
Tasks/WaveSampler.cpp:96-175,222-292, Tasks/HistogramFeeder.cpp:35-87, Tasks/CarrierDetector.cpp:49-147) and against
its own frozen outputs (`pin_oracle.json`: self-pins, not
upstream vectors).  The GPU tests compare the CUDA path with the same files through the C-ABI."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)


def test_oracle_spectrum_view_vs_reference_fixture(oracle):
    import make_golden as G
    ref = np.load(os.path.join(GOLD, "ref_spectrumview.npz"))
    fmin, fmax, fftbw, psize, feeds = G.sview_case()
    L = oracle.lib()
    v = oracle.SpectrumView()
    assert L.sdo_sview_init(C.byref(v)) == 0
    L.sdo_sview_set_range(C.byref(v), fmin, fmax)
    v.fft_bandwidth = fftbw
    for data, center, _ in feeds:
        L.sdo_sview_feed(C.byref(v), oracle.ptr(data), None, psize, center, 1)
    size = int(ref["size"])
    assert v.spectrum_size == size
    got = np.ctypeslib.as_array(v.psd, shape=(65536,))[:size]
    cnt = np.ctypeslib.as_array(v.psd_count, shape=(65536,))[:size]
    assert np.array_equal(cnt, ref["count"])
    assert np.array_equal(got.view(np.uint32), ref["psd"].view(np.uint32))
    L.sdo_sview_free(C.byref(v))


def test_oracle_quad_demod_and_fac_vs_reference_fixtures(oracle):
    q = np.load(os.path.join(GOLD, "ref_quad_demod.npz"))
    x = q["x"]
    y = np.empty_like(x)
    prev, primed = oracle.Cpx(0, 0), C.c_int(0)
    oracle.lib().sdo_quad_demod(oracle.ptr(x), oracle.ptr(y), len(x), C.byref(prev), C.byref(primed))
    assert y[0] == 0 and np.abs(y - q["y"]).max() < 3e-7       # float32 atan2 (SPEC M) against float64
    f = np.load(os.path.join(GOLD, "ref_fac.npz"))
    fac = oracle.spectsrc_frame("fac", 1024, f["x"])
    assert fac is not None and len(fac) == 512
    assert np.abs(fac - f["fac"]).max() <= 2e-6 * f["fac"].max()


def _check_timewindow(impl):
    """impl: the oracle wrappers (CPU) or the package (CUDA path through the C-ABI); same calls, same fixture"""
    r = np.load(os.path.join(GOLD, "ref_timewindow.npz"))
    bits = lambda a: np.ascontiguousarray(a).view(np.uint32)            # noqa: E731
    for d in (7, 500):
        got, ref = impl.delayed_conj(r["tone"][:3000], d), r["dconj_%d" % d]
        assert np.all(got[:d] == 0)
        # cabsf of the reference's libm against sqrt(re^2 + im^2) of SPEC M: one unit in the last place of k
        assert np.abs(got - ref).max() <= 2.5e-7 * np.abs(ref).max()
    for space, sig in (("amplitude", "ask"), ("phase", "psk"), ("frequency", "psk")):
        got = impl.sample_manual(r[sig][:6000], space, 487.3, 5)
        assert np.array_equal(bits(got), bits(r["manual_" + space])), space        # no libm involved: bit for bit
    for name, sig, space, amp, thr, z, bnor in (
            ("zc_amp_power", "ask", "amplitude", True, 0.6 + 0.1j, 1 + 0j, 1.0 / 12),
            ("zc_amp_proj", "ask", "amplitude", False, 0.55 + 0.2j, np.exp(-0.3j), 1.0 / 12),
            ("zc_phase", "psk", "phase", False, 0j, np.exp(0.1j), 1.0 / 12),
            ("zc_amp_fast", "ask", "amplitude", True, 0.6 + 0.1j, 1 + 0j, 1.0)):
        got, total = impl.sample_zero_crossing(r[sig], space, bnor, amplitude=amp, threshold=thr, zc_angle=z)
        assert total == len(r[name]) and np.array_equal(got, r[name]), name
    for space in ("amplitude", "phase", "frequency"):
        assert np.abs(impl.histogram_feed(r["psk"], space) - r["hist_" + space]).max() < 4e-7
    for (m, notch), ref in zip(((4096, 0.0), (6000, 0.0), (9692, 0.02)), r["carrier"]):
        assert abs(impl.carrier_detect(r["tone"][:m], 0.004, notch) - ref) < 2e-6
        assert abs(ref / (2 * np.pi) - 0.0731) < 2e-4


def test_oracle_timewindow_tasks_vs_reference_fixture(oracle):
    _check_timewindow(oracle)


@pytest.mark.gpu
def test_gpu_timewindow_tasks_vs_reference_fixture(sdb):
    _check_timewindow(sdb)


def test_oracle_frozen_outputs(oracle):
    import make_golden as G
    want = json.load(open(os.path.join(GOLD, "pin_oracle.json")))
    got = G.pins()
    assert got == want, {k: (got[k], want[k]) for k in want if got.get(k) != want[k]}


@pytest.mark.gpu
def test_gpu_spectrum_view_vs_reference_fixture(sdb):
    import make_golden as G
    ref = np.load(os.path.join(GOLD, "ref_spectrumview.npz"))
    fmin, fmax, fftbw, psize, feeds = G.sview_case()
    import torch
    hops = torch.from_numpy(np.stack([d for d, _, _ in feeds])).cuda()
    v = sdb.SpectrumView(fmin, fmax, fftbw, 0.5)
    v.project(hops.data_ptr(), psize, [c for _, c, _ in feeds])
    v.accumulate()
    psd, _, cnt = v.read()
    size = int(ref["size"])
    assert len(psd) == size
    assert np.array_equal(cnt, ref["count"])
    assert np.array_equal(psd.view(np.uint32), ref["psd"].view(np.uint32))


@pytest.mark.gpu
def test_gpu_quad_demod_vs_reference_fixture(sdb):
    q = np.load(os.path.join(GOLD, "ref_quad_demod.npz"))
    y = sdb.quad_demod(q["x"])
    assert y[0] == 0 and np.abs(y - q["y"]).max() < 3e-7

#!/usr/bin/env python
"""Generates the fixtures under tests/golden/ (run from the repo root: `python tests/golden/make_golden.py`).

The reference repo ships no test vectors and its DSP lives in libraries that are absent, so there
are two kinds of fixture, kept apart on purpose:

* `ref_*.npz` -- outputs of LITERAL Python transcriptions of the arithmetic that IS in /root/reference:
    - Panoramic/Scanner.cpp:56-256 (SpectrumView::feed, linear + histogram modes, interpolate, forgetting rule):
      transcription `_sview_literal` in tests/test_oracle.py;
    - Tasks/QuadDemodTask.cpp:44-60 (dst[0] = 0, dst[p] = i/pi * arg(x[p] conj(x[p-1]))) evaluated in float64;
    - Default/GenericInspector/FACTab.cpp:209-221 (FFT -> x conj(x) -> inverse FFT -> |.| of the first half) in float64.
  These pin the oracle to the reference where the reference is specific.  Nothing here needs the reference at test time.
* `pin_*.json` -- SHA-256 of the ORACLE's own outputs on seeded inputs for the parts whose inner arithmetic is upstream
  (FFT dataflows, AGC / Costas / Gardner chains, channeliser, spectrum sources, channel detector).  They are NOT
  upstream golden vectors ("parity unpinned"); they freeze SPEC.md so that a change of the oracle -- against which the
  CUDA path is compared bit for bit -- cannot go unnoticed.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O                      # noqa: E402
from sigdigger_b200 import synth            # noqa: E402
from test_oracle import _sview_literal      # noqa: E402


def sview_case():
    fmin, fmax, fftbw, psize = 100e6, 140e6, 2e6, 2048
    rng = np.random.default_rng(7)
    feeds = []
    for hop in range(60):
        feeds.append(((rng.random(psize).astype(np.float32) * 10 - 90), fmin + 0.5e6 + hop * 0.66e6, fftbw))
    feeds += feeds[10:30]
    feeds += feeds[10:30] * 4
    return fmin, fmax, fftbw, psize, feeds


def quad_case():
    rng = np.random.default_rng(11)
    n = 4096
    x = (np.exp(2j * np.pi * np.cumsum(0.05 + 0.02 * rng.standard_normal(n))) * (0.5 + rng.random(n))).astype(np.complex64)
    ref = np.zeros(n, np.complex128)
    xd = x.astype(np.complex128)
    ref[1:] = 1j / np.pi * np.angle(xd[1:] * np.conj(xd[:-1]))
    return x, ref


def fac_case():
    n = 4096
    s, _ = synth.psk_signal(n + 1, 8.0, order=4, seed=3)
    rng = np.random.default_rng(3)
    x = (0.3 * s + synth.awgn(n + 1, 1e-3, rng)).astype(np.complex64)
    fr = x[-1024:].astype(np.complex128)
    ref = np.abs(np.fft.ifft(np.abs(np.fft.fft(fr)) ** 2))[:512]      # FACTab: unnormalised inverse / N here
    return x, ref


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pins():
    """seeded oracle outputs (see module docstring: self-pins, not upstream vectors)"""
    out = {}
    rng = np.random.default_rng(2026)
    z = (rng.standard_normal(65536) + 1j * rng.standard_normal(65536)).astype(np.complex64)
    out["spec_fft_65536"] = sha(O.spec_fft(z))
    out["spec_fft_4096_stockham"] = sha(O.spec_fft(z[:4096]))
    out["spec_fft_8192_fourstep"] = sha(O.spec_fft(z[:8192]))
    out["psd_bh_8192"] = sha(O.psd_frames(z[:8192 * 4], 8192, "blackmann_harris"))
    n, sps = 40000, 3.125
    s, _ = synth.psk_signal(n, sps, order=4, seed=5)
    x = (0.25 * synth.mix(s, 1e-4, 0.2) + synth.awgn(n, 10 ** (-45 / 20), np.random.default_rng(5))).astype(np.complex64)
    kw = dict(baud=1.0 / sps, costas_order=2, bits_per_symbol=2, loop_bw=2e-3, mf_type=1, mf_rolloff=0.35, clock_type=1,
              clock_gain=0.1)
    soft, hard = O.inspector_run(O.insp_config("psk", 1.0, **kw), x)
    out["psk_soft"], out["psk_hard"] = sha(soft), sha(hard)
    soft, hard = O.inspector_run(O.insp_config("psk", 1.0, eq_type=1, eq_rate=5e-3, **kw), x)
    out["psk_cma_soft"] = sha(soft)
    fs_, _ = synth.fsk_signal(30000, 5.0, h=1.0, seed=4)
    xf = (0.3 * fs_ + synth.awgn(30000, 10 ** (-35 / 20), np.random.default_rng(4))).astype(np.complex64)
    soft, hard = O.inspector_run(O.insp_config("fsk", 1.0, baud=0.2, bits_per_symbol=1, mf_type=1, clock_type=1,
                                               clock_gain=0.2), xf)
    out["fsk_soft"] = sha(soft)
    W = 8192
    xc = (z[:W * 8] * 0.05 + 0.4 * np.exp(2j * np.pi * 0.1251 * np.arange(W * 8))).astype(np.complex64)
    ch = O.specttuner_run(xc, W, [dict(f0=2 * np.pi * 0.125, bw=2 * np.pi / 32, guard=1.0)])[0]
    out["channeliser"] = sha(ch)
    out["spectsrc_cyclo"] = sha(O.spectsrc_frame("cyclo", 1024, ch))
    out["spectsrc_fac"] = sha(O.spectsrc_frame("fac", 1024, ch))
    d = O.ChannelDetector(8192, 0.25, 0.5, 6.0, 3)
    chans, total = d.feed(O.psd_frames(xc[:8192 * 4], 8192, "hann"))
    d.close()
    out["chdet"] = sha(np.array([(a, b) for a, b, *_ in chans], np.uint32))
    out["chdet_total"] = int(total)
    return out


def main():
    fmin, fmax, fftbw, psize, feeds = sview_case()
    psd, acc, cnt, size = _sview_literal(fmin, fmax, feeds)
    np.savez_compressed(os.path.join(HERE, "ref_spectrumview.npz"), psd=psd, count=cnt, size=size)
    x, ref = quad_case()
    np.savez_compressed(os.path.join(HERE, "ref_quad_demod.npz"), x=x, y=ref)
    x, ref = fac_case()
    np.savez_compressed(os.path.join(HERE, "ref_fac.npz"), x=x, fac=ref)
    json.dump(pins(), open(os.path.join(HERE, "pin_oracle.json"), "w"), indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()

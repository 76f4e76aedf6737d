"""Recording side of the path: the baseband capture writer, the audio saver and the
inspector recording formats, host-only.  Known answers from the reference's conventions (file names:
Default/Source/SourceWidget.cpp:1092-1100, Audio/AudioFileSaver.cpp:86-96; data variables:
Default/GenericInspector/InspectorUI.cpp:860-930) and round trips through the capture reader."""
import json
import os
import wave

import numpy as np
import pytest


@pytest.fixture
def sdbh():
    import sigdigger_b200
    sigdigger_b200.load_library()
    return sigdigger_b200


def _iq(n, seed=1, amp=0.4):
    rng = np.random.default_rng(seed)
    return (amp * (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n))).astype(np.complex64)


def test_capture_name_is_the_one_the_reader_guesses_from(sdbh, tmp_path):
    assert sdbh.capture_file_name(1706745599, 2e6, 433.92e6) == \
        "sigdigger_20240131_235959Z_2000000_433920000_float32_iq.raw"
    x = _iq(5000)
    r = sdbh.Recorder(str(tmp_path), samp_rate=250000.0, frequency=145.8e6, start_time=1700000000, auto_name=True)
    assert os.path.basename(r.path) == "sigdigger_20231114_221320Z_250000_145800000_float32_iq.raw"
    for k in range(0, 5000, 777):                      # the hook delivers blocks of any size
        assert r.write(x[k:k + 777]) == len(x[k:k + 777])
    assert r.samples == 5000
    r.close()
    assert oct(os.stat(r.path).st_mode & 0o777) == "0o600"        # creat(path, 0600)
    c = sdbh.Capture(r.path)
    assert (c.info.samp_rate, c.info.frequency, c.info.start_time, c.info.n_samples) == (250000.0, 145.8e6, 1700000000, 5000)
    assert np.array_equal(c.samples, x)                # float32: the file is the sample stream itself
    c.close()


@pytest.mark.parametrize("fmt,dtype,scale,off", [("u8", np.uint8, 128.0, 128.0), ("s8", np.int8, 128.0, 0.0),
                                                 ("s16", np.int16, 32768.0, 0.0)])
def test_native_formats_invert_the_loader_scaling(sdbh, tmp_path, fmt, dtype, scale, off):
    info = np.iinfo(dtype)
    # every code of the format, as the loader would present it (SPEC Q), comes back as the same code
    codes = np.arange(info.min, info.max + 1, 1 if fmt != "s16" else 7).astype(np.int64)
    if len(codes) % 2:
        codes = codes[:-1]
    f = ((codes - off) / scale).astype(np.float32)
    x = (f[0::2] + 1j * f[1::2]).astype(np.complex64)
    r = sdbh.Recorder(str(tmp_path), samp_rate=1e6, frequency=1e8, sample_format=fmt, start_time=1700000000,
                      auto_name=True)
    r.write(x)
    r.write(np.array([5 + 5j, -5 - 5j, np.nan + 0j], np.complex64))          # saturation; NaN -> mid-scale
    r.close()
    c = sdbh.Capture(r.path)                           # format token in the name -> guessed by the reader
    assert c.info.sample_format == sdbh.FORMAT[fmt] and c.info.guessed & 8
    got = c.samples.astype(np.int64)
    assert np.array_equal(got[:len(codes)], codes)
    mid = int(off)
    assert list(got[len(codes):]) == [info.max, info.max, info.min, info.min, mid, mid]
    c.close()


def test_wav_and_sigmf_containers_round_trip(sdbh, tmp_path):
    x = _iq(3000, seed=3)
    for fmt, tol in (("f32", 0.0), ("s16", 1.0 / 32768), ("u8", 1.0 / 128)):
        p = str(tmp_path / ("cap_%s.wav" % fmt))
        r = sdbh.Recorder(p, samp_rate=48000.0, container="wav", sample_format=fmt)
        r.write(x[:1000]); r.write(x[1000:])
        r.close()
        with open(p, "rb") as f:
            h = f.read(44)
        assert h[:4] == b"RIFF" and int.from_bytes(h[4:8], "little") == os.path.getsize(p) - 8
        assert int.from_bytes(h[40:44], "little") == os.path.getsize(p) - 44
        c = sdbh.Capture(p)
        assert (c.info.container, c.info.sample_format, c.info.n_samples, c.info.samp_rate) == \
            (sdbh.CONTAINER["wav"], sdbh.FORMAT[fmt], 3000, 48000.0)
        s = c.samples
        if fmt == "f32":
            assert np.array_equal(s, x)
        else:
            v = s.astype(np.float32).reshape(-1, 2)
            v = (v - 128.0) / 128.0 if fmt == "u8" else v / 32768.0
            assert np.max(np.abs(v[:, 0] - x.real)) <= tol / 2 + 1e-7 and np.max(np.abs(v[:, 1] - x.imag)) <= tol / 2 + 1e-7
        c.close()
    with pytest.raises(sdbh.SdbError):
        sdbh.Recorder(str(tmp_path / "x.wav"), samp_rate=1.0, container="wav", sample_format="s8")
    for fmt in ("f32", "s16", "s8", "u8"):
        p = str(tmp_path / ("rec_%s" % fmt))
        r = sdbh.Recorder(p, samp_rate=2.4e6, frequency=1090e6, container="sigmf", sample_format=fmt,
                          start_time=1700000000)
        assert r.path.endswith(".sigmf-data")
        r.write(x)
        r.close()
        meta = json.load(open(p + ".sigmf-meta"))
        assert meta["global"]["core:sample_rate"] == 2.4e6 and meta["captures"][0]["core:frequency"] == 1090e6
        assert meta["captures"][0]["core:datetime"] == "2023-11-14T22:13:20Z"
        c = sdbh.Capture(p + ".sigmf-meta")
        assert (c.info.container, c.info.sample_format, c.info.n_samples) == (sdbh.CONTAINER["sigmf"], sdbh.FORMAT[fmt], 3000)
        assert (c.info.samp_rate, c.info.frequency) == (2.4e6, 1090e6)
        c.close()


def test_audio_saver_names_and_pcm(sdbh, tmp_path):
    t = np.arange(4410) / 44100.0
    a = (0.5 * np.sin(2 * np.pi * 1000 * t) + 0.25j).astype(np.complex64)     # the consumer uses Re{x} only
    paths = []
    for k in range(3):
        r = sdbh.Recorder(str(tmp_path), samp_rate=44100, frequency=145.8e6, audio="fm")
        r.write(a)
        r.write(np.array([2.0, -2.0], np.complex64))                             # clipped, not wrapped
        r.close()
        paths.append(os.path.basename(r.path))
    assert paths == ["audio-FM-145800000-44100-%04d.wav" % i for i in (1, 2, 3)]   # first free index
    w = wave.open(str(tmp_path / paths[0]))
    assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 44100, 4412)
    pcm = np.frombuffer(w.readframes(4412), np.int16)
    assert np.array_equal(pcm[:4410], np.rint(a.real * np.float32(32767.0)).astype(np.int16))
    assert list(pcm[4410:]) == [32767, -32768]
    assert os.path.basename(sdbh.Recorder(str(tmp_path), samp_rate=8000, frequency=7.1e6, audio="usb").path) == \
        "audio-USB-7100000-8000-0001.wav"
    with pytest.raises(sdbh.SdbError):
        sdbh.Recorder(str(tmp_path / "no" / "such" / "dir"), samp_rate=8000, audio="am")


def test_audio_plan_follows_the_audio_processor(sdbh):
    """Default/Audio/AudioProcessor.cpp:118-151,200-228,250-269 by hand"""
    p = sdbh.audio_plan(2e6, 44100, "fm", lo=150e3, bw=12500.0)
    assert (p.max_audio_bw, p.sample_rate, p.true_bw, p.true_lo) == (2e5, 44100, 12500.0, 150e3)
    assert (p.ch_fc, p.ch_ft, p.ch_bw, p.ch_f_lo, p.ch_f_hi) == (150e3, 0.0, 2e5, -1e5, 1e5)
    # side bands: half the bandwidth, LO moved by half of THAT (a quarter of the selected width)
    u = sdbh.audio_plan(2e6, 44100, "usb", lo=7000.0, bw=2800.0)
    l = sdbh.audio_plan(2e6, 44100, "lsb", lo=7000.0, bw=2800.0)
    assert (u.true_bw, u.true_lo, l.true_bw, l.true_lo) == (1400.0, 7700.0, 1400.0, 6300.0)
    # narrow analyzer: everything is limited by fs / 2, the playback rate is floored to it
    n = sdbh.audio_plan(60000.0, 44100, "am", lo=0.0, bw=1e6)
    assert (n.max_audio_bw, n.sample_rate, n.true_bw) == (30000.0, 30000, 30000.0)
    assert sdbh.audio_plan(2e6, 8000, "am", lo=0.0, bw=0.2).true_bw == 1.0           # never below 1 Hz
    assert sdbh.audio_plan(2e6, 8000, "am", lo=3e5, bw=1e4).ch_fc == 0.0             # outside +-maxAudioBw: opened at 0
    assert sdbh.audio_plan(2e6, 8000, "am", lo=-2e5, bw=1e4).ch_fc == -2e5           # the limit itself is kept
    with pytest.raises(sdbh.SdbError):
        sdbh.audio_plan(1.5, 44100, "am", lo=0.0, bw=1.0)                            # rate floors to 0
    cfg = sdbh.audio_plan_config(u, "usb", cutoff=3000.0, squelch=True, squelch_level=0.05, agc=False, agc_ts=0.5)
    assert (cfg.audio_demod, cfg.audio_sample_rate, cfg.audio_volume) == (3, 44100, 1.0)   # USB = 2, +1 on the wire
    assert (cfg.audio_cutoff, cfg.audio_squelch, cfg.agc_enabled, cfg.agc_ts) == (3000.0, 1, 0, 0.5)
    assert abs(cfg.audio_squelch_level - 0.05) < 1e-9


def test_inspector_recording_formats(sdbh):
    rng = np.random.default_rng(5)
    soft = (rng.standard_normal(1000) + 1j * rng.standard_normal(1000)).astype(np.complex64)
    hard = rng.integers(0, 4, 1000).astype(np.uint8)
    assert np.array_equal(sdbh.inspector_forward("soft_bits", soft), soft)
    assert np.array_equal(sdbh.inspector_forward("soft_bits_i", soft), soft.real)
    assert np.array_equal(sdbh.inspector_forward("soft_bits_q", soft), soft.imag)
    assert np.array_equal(sdbh.inspector_forward("symbols", soft, hard), hard)
    d = sdbh.inspector_forward("decision_space", soft, decision_mode="modulus")
    assert d.dtype == np.float32 and np.allclose(d, np.abs(soft.astype(np.complex128)), rtol=1e-6)
    d = sdbh.inspector_forward("decision_space", soft, decision_mode="argument")
    assert np.allclose(d, np.angle(1j * soft.astype(np.complex128)) / np.pi, atol=2e-6) and np.all(np.abs(d) <= 1.0)
    with pytest.raises(sdbh.SdbError):
        sdbh.inspector_forward("symbols", soft)          # no decision available

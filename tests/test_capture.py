"""Capture-file front end (Default/SourceConfig/FileSourcePage.cpp:68-140): container parsing on
the CPU (no GPU needed) and, on the GPU, a WAV / raw capture pushed through the analyzer in its native sample format."""
import json
import os
import struct
import wave

import numpy as np
import pytest


@pytest.fixture
def sdbh():
    """the binding without the device check of the `sdb` fixture: container parsing is host-only"""
    import sigdigger_b200
    sigdigger_b200.load_library()
    return sigdigger_b200


def _iq_s16(n, seed=1):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = 0.3 * np.exp(2j * np.pi * 0.11 * t) + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty(2 * n, np.int16)
    iq[0::2] = np.round(x.real * 32767).astype(np.int16)
    iq[1::2] = np.round(x.imag * 32767).astype(np.int16)
    return iq


def test_raw_capture_and_name_guess(sdbh, tmp_path):
    x = (np.arange(2000) + 1j * np.arange(2000)).astype(np.complex64)
    p = tmp_path / "sigdigger_20240131_235959Z_2000000_433920000_float32_iq.raw"
    x.tofile(p)
    c = sdbh.Capture(str(p))
    i = c.info
    assert (i.container, i.sample_format, i.n_samples, i.data_offset) == (sdbh.CONTAINER["raw"], sdbh.FORMAT["f32"], 2000, 0)
    assert (i.samp_rate, i.frequency) == (2e6, 433.92e6) and i.guessed & 7 == 7
    assert i.start_time == 1706745599                       # 2024-01-31 23:59:59 UTC
    assert np.array_equal(c.samples, x)
    c.close()
    q = tmp_path / "capture.u8"
    np.arange(256, dtype=np.uint8).tofile(q)
    c = sdbh.Capture(str(q), sample_format="u8")             # explicit format, nothing to guess
    assert (c.info.sample_format, c.info.n_samples, c.info.guessed, c.info.samp_rate) == (sdbh.FORMAT["u8"], 128, 0, 0.0)
    assert np.array_equal(c.samples, np.arange(256, dtype=np.uint8))
    c.close()
    with pytest.raises(sdbh.SdbError):
        sdbh.Capture(str(tmp_path / "missing.raw"))


def test_wav_capture(sdbh, tmp_path):
    iq = _iq_s16(3000)
    p = tmp_path / "iq.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(250000)
        w.writeframes(iq.tobytes())
    c = sdbh.Capture(str(p))
    assert (c.info.container, c.info.sample_format, c.info.n_samples, c.info.samp_rate) == (
        sdbh.CONTAINER["wav"], sdbh.FORMAT["s16"], 3000, 250000.0)
    assert c.info.data_offset == 44 and np.array_equal(c.samples, iq)
    c.close()
    # IEEE float WAV with an extra chunk before "data" and an odd-sized chunk to skip
    x = (np.arange(100) * (1 + 2j) / 100).astype(np.complex64)
    body = x.tobytes()
    fmt = struct.pack("<HHIIHH", 3, 2, 48000, 48000 * 8, 8, 32)
    junk = b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\0"
    riff = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + junk + b"data" + struct.pack("<I", len(body)) + body
    q = tmp_path / "f32.wav"
    q.write_bytes(b"RIFF" + struct.pack("<I", len(riff)) + riff)
    c = sdbh.Capture(str(q))
    assert (c.info.sample_format, c.info.n_samples, c.info.samp_rate) == (sdbh.FORMAT["f32"], 100, 48000.0)
    assert np.array_equal(c.samples, x)
    c.close()
    # mono audio is not an IQ capture
    m = tmp_path / "mono.wav"
    with wave.open(str(m), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(8000)
        w.writeframes(iq[:100].tobytes())
    with pytest.raises(sdbh.SdbError, match="two channels"):
        sdbh.Capture(str(m))


@pytest.mark.parametrize("dtype,name,fmt", [("cf32_le", "f32", np.complex64), ("ci16_le", "s16", np.int16),
                                            ("ci8", "s8", np.int8), ("cu8", "u8", np.uint8)])
def test_sigmf_capture(sdbh, tmp_path, dtype, name, fmt):
    n = 500
    if fmt is np.complex64:
        data = (np.arange(n) * (1 - 1j)).astype(np.complex64)
    else:
        data = (np.arange(2 * n) % 100).astype(fmt)
    stem = tmp_path / "rec"
    data.tofile(str(stem) + ".sigmf-data")
    meta = {"global": {"core:datatype": dtype, "core:sample_rate": 2.4e6, "core:version": "1.0.0"},
            "captures": [{"core:sample_start": 0, "core:frequency": 1.4204e9}], "annotations": []}
    open(str(stem) + ".sigmf-meta", "w").write(json.dumps(meta, indent=2))
    for path in (str(stem) + ".sigmf-meta", str(stem) + ".sigmf-data"):
        c = sdbh.Capture(path)
        assert (c.info.container, c.info.sample_format, c.info.n_samples) == (sdbh.CONTAINER["sigmf"], sdbh.FORMAT[name], n)
        assert (c.info.samp_rate, c.info.frequency) == (2.4e6, 1.4204e9)
        assert np.array_equal(c.samples, data)
        c.close()


@pytest.mark.gpu
def test_wav_capture_through_the_analyzer(sdb, oracle, tmp_path):
    """s16 WAV -> sdb_capture -> analyzer in the native format (conversion in the first load): PSD frames
    bit-identical to the oracle on v / 32768."""
    from sigdigger_b200.analyzer import Analyzer
    N = 4096
    iq = _iq_s16(N * 8, seed=3)
    p = tmp_path / "iq.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(1000000)
        w.writeframes(iq.tobytes())
    c = sdb.Capture(str(p))
    a = Analyzer(c.info.samp_rate, window_size=N, window="hann", psd_update_int=0.0, data=c.samples, read_size=N * 4)
    psd = []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "PSD":
            psd.append(m["psd"])
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    a.close()
    c.close()
    assert name == "EOS" and len(psd) == 8
    x = (iq[0::2].astype(np.float32) / np.float32(32768) + 1j * (iq[1::2].astype(np.float32) / np.float32(32768)))
    ref = oracle.psd_frames(x.astype(np.complex64), N, "hann")
    assert np.array_equal(np.stack(psd).view(np.uint32), ref.view(np.uint32))

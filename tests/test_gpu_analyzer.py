"""-m gpu: the suscan-style asynchronous analyzer (message queue + requests), SURVEY.md 8(a) a18 / 8(b).
Message flow follows the reference's handshake (Suscan/AnalyzerRequestTracker.cpp:138-157): OPEN reply ->
SET_ID -> SET_CONFIG -> SAMPLES keyed by the caller's inspector id; PSD messages; EOS terminates."""
import ctypes as C
import threading

import numpy as np
import pytest

import parity
from sigdigger_b200 import synth

pytestmark = pytest.mark.gpu


class Gate:
    """Gate of the in-memory source: the analyzer's worker handles its request queue, then reads a block.  The read
    callback parks in wait(); `parked` tells the test that the worker is past its request loop, so that requests sent
    now take effect after the block about to be released (without it the tests raced the worker's start-up)."""

    def __init__(self):
        self.go, self.parked = threading.Event(), threading.Event()

    def wait(self, timeout):
        self.parked.set()
        return self.go.wait(timeout)

    def set(self):
        self.go.set()



def test_analyzer_message_flow_and_parity(sdb, oracle):
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 8192, 1.0e6
    baud = fs / 100.0
    blocks, per_block = 6, N * 8
    n = blocks * per_block
    x, _ = synth.multi_carrier(n, fs, [("qpsk", 0.125 * fs + 3.0, baud, -10.0, {})], noise_db=-50.0, seed=31)
    go = Gate()
    pos = [0]

    def read(priv, dst, maxn):                    # plays the role of a file / SDR source back-end
        go.wait(30)
        take = min(maxn, n - pos[0])
        if take > 0:
            C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
            pos[0] += take
        return take

    a = Analyzer(fs, window_size=N, window="blackmann_harris", psd_update_int=N / fs, read=read, read_size=per_block)
    name, info = a.read(5000)
    assert name == "SOURCE_INFO"
    assert go.parked.wait(10)
    # the worker is now blocked in the first read; queue the handshake, then let the source run
    a.open("psk", 0.125 * fs, 3 * baud, req_id=11)
    a.open("nope", 0.0, 1000.0, req_id=12)                       # unknown class -> WRONG_KIND
    a.set_inspector_id(0, 0xBEEF, req_id=13)
    cfg = sdb.InspectorConfig()
    fs_ch = fs * 256 / N
    sdb._check(sdb.load_library().sdb_inspector_config_default(C.byref(cfg), sdb.INSP["psk"], fs_ch))
    kw = dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=fs_ch * 2e-3, mf_type=1, mf_rolloff=0.35,
              clock_type=1, clock_gain=0.1, clock_running=1)
    for k, v in kw.items():
        setattr(cfg, k, v)
    a.set_inspector_config(0, cfg, req_id=14)
    a.set_inspector_id(7, 1, req_id=15)                          # bad handle -> WRONG_HANDLE
    go.set()
    psd, soft, hard, replies = [], [], [], []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "PSD":
            psd.append(m["psd"])
            assert m["samp_rate"] == fs
        elif name == "SAMPLES":
            assert m["inspector_id"] == 0xBEEF
            soft.append(m["samples"])
            hard.append(m["symbols"])
        elif name == "INSPECTOR":
            replies.append((m["kind"], m["req_id"], m["handle"]))
            if m["kind"] == "OPEN":
                assert m["class_name"] == "psk" and abs(m["equiv_fs"] - fs_ch) < 1e-3 and m["config"].clock_running == 0
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    assert name == "EOS"
    a.close()
    assert replies == [("OPEN", 11, 0), ("WRONG_KIND", 12, -1), ("SET_ID", 13, 0), ("SET_CONFIG", 14, 0),
                       ("WRONG_HANDLE", 15, 7)]
    # every PSD frame of every block, bit-identical to the oracle
    ref_psd = oracle.psd_frames(x, N, "blackmann_harris")
    assert len(psd) == n // N
    assert np.array_equal(np.stack(psd).view(np.uint32), ref_psd.view(np.uint32))
    # the inspector became active at the second block (requests are applied at block boundaries)
    ic = oracle.insp_config("psk", fs_ch, **kw)
    f0, bw = float(np.float32(2.0 * np.pi * 0.125)), float(np.float32(2.0 * np.pi * (3 * baud) / fs))
    ref = oracle.analyzer_run(oracle.make_an_params(N, "blackmann_harris", [(f0, bw, 1.0, 0, ic)]), x[per_block:])
    parity.assert_symbols_match(np.concatenate(soft), np.concatenate(hard), ref["soft"][0], ref["hard"][0],
                                exact_soft=True)


def test_analyzer_memory_source_halt(sdb):
    from sigdigger_b200.analyzer import Analyzer
    N = 4096
    x = (0.1 * np.exp(2j * np.pi * 0.1 * np.arange(N * 16))).astype(np.complex64)
    a = Analyzer(1e6, window_size=N, window="hann", psd_update_int=0.0, data=x, loop=True, read_size=N * 4)
    seen = 0
    for _ in range(50):
        name, m = a.read(10000)
        if name == "PSD":
            seen += 1
            k = int(np.argmax(m["psd"]))
            assert abs(k - 0.1 * N) <= 1
        if seen >= 12:
            break
    assert seen >= 12
    a.halt()
    for _ in range(1000):
        name, m = a.read(10000)
        if name == "HALT":
            break
    assert name == "HALT"
    a.close()


def test_analyzer_wide_spectrum_sweep(sdb, oracle):
    """SUSCAN_ANALYZER_MODE_WIDE_SPECTRUM (Panoramic/Scanner.cpp:296-372, 492-523): retune, drop the buffering
    samples, one PSD per hop with the hop centre in `fc`; hop PSDs bit-identical to the oracle on the samples the
    source delivered; progressive plan, then a narrower range."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 4096, 1.0e6
    fmin, fmax = 100.0e6, 104.0e6
    carriers = [(100.6e6, 0.30), (102.1e6, 0.20), (103.3e6, 0.25)]
    state = {"fc": 0.0, "t": 0, "hops": [], "retunes": []}
    total_hops = 40

    def set_frequency(f):
        state["fc"] = f
        state["retunes"].append(f)
        return 0

    def read(priv, dst, n):
        if len(state["hops"]) >= total_hops:
            return 0
        t = state["t"] + np.arange(n)
        state["t"] += n
        x = np.zeros(n, np.complex128)
        for f, amp in carriers:
            if abs(f - state["fc"]) < fs / 2:
                x += amp * np.exp(2j * np.pi * ((f - state["fc"]) / fs) * t)
        x += 1e-3 * np.exp(2j * np.pi * 0.37 * t * t / n)             # a deterministic low-level chirp as "noise"
        x = x.astype(np.complex64)
        C.memmove(dst, x.ctypes.data, 8 * n)
        if n == N:
            state["hops"].append((state["fc"], x))
        return n

    a = Analyzer(fs, window_size=N, window="blackmann_harris", read=read, wide=True, min_freq=fmin, max_freq=fmax,
                 set_frequency=set_frequency)
    a.set_buffering_size(1000)
    a.set_rel_bandwidth(0.5)
    msgs = []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "PSD":
            msgs.append(m)
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    a.close()
    assert name == "EOS" and len(msgs) == total_hops == len(state["hops"])
    # progressive plan: step = rel_bw fs = 0.5 MHz, centres min + step/2 + i step, wrapping after 8 hops
    centres = [fmin + 0.25e6 + 0.5e6 * (i % 8) for i in range(total_hops)]
    assert [m["fc"] for m in msgs] == [int(round(c)) for c in centres]
    assert state["retunes"][:total_hops] == centres
    for m, (fc, x) in zip(msgs, state["hops"]):
        assert m["fc"] == int(round(fc))
        ref = oracle.psd_frames(x, N, "blackmann_harris")[0]
        assert np.array_equal(m["psd"].view(np.uint32), ref.view(np.uint32))
        inband = [f for f, _ in carriers if abs(f - fc) < 0.2 * fs]
        if inband:                                                       # the carrier shows where it should
            k = int(np.argmax(m["psd"]))
            assert abs(((k + N // 2) % N - N // 2) - round((inband[0] - fc) / fs * N)) <= 1
    # the buffering samples were requested and dropped: 1000 + 4096 per hop (requests apply from the next batch of 16
    # hops on, so the first batch may have run without them)
    assert state["t"] in (total_hops * (1000 + N), total_hops * N + (total_hops - 16) * 1000)


def test_analyzer_history_and_replay(sdb):
    """suscan_analyzer_set_history_size / _replay (Suscan/Analyzer.cpp:157-167): the last samples stay in a ring;
    replay pauses the source and loops over the ring."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 4096, 1.0e6
    blk = N * 2
    go = Gate()
    step = threading.Semaphore(0)
    calls = [0]

    def read(priv, dst, n):                               # block k carries a tone in bin 100 + 50 k
        go.wait(30)
        if not step.acquire(timeout=30):
            return 0
        k = calls[0]
        calls[0] += 1
        x = (0.3 * np.exp(2j * np.pi * (100 + 50 * k) / N * np.arange(n))).astype(np.complex64)
        C.memmove(dst, x.ctypes.data, 8 * n)
        return n

    a = Analyzer(fs, window_size=N, window="hann", psd_update_int=0.0, read=read, read_size=blk)
    assert a.read(5000)[0] == "SOURCE_INFO"
    assert go.parked.wait(10)          # requests sent from here on are handled after the first block
    assert a.set_history_size(3 * blk)
    go.set()

    def frames(k, release=True):
        out = []
        if release:
            for _ in range(k):
                step.release()
        while len(out) < 2 * k:
            name, m = a.read(20000)
            assert name not in ("TIMEOUT", "EOS", "READ_ERROR", "HALT")
            if name == "PSD":
                out.append((int(np.argmax(m["psd"])), m["looped"], m["history_size"]))
        return out

    live = frames(6)                                      # blocks 0..5; the ring is armed from block 1 on
    assert [p for p, _, _ in live] == [100 + 50 * (i // 2) for i in range(12)]
    assert live[-1][2] == 3 * blk and all(lp == 0 for _, lp, _ in live)
    a.replay(True)
    frames(1)                                             # the block in flight when the request lands
    n_calls = calls[0]
    rep = frames(5, release=False)                        # the source is paused: the ring (blocks 4, 5, 6) loops
    assert calls[0] == n_calls
    assert [p for p, _, _ in rep] == [100 + 50 * b for b in (4, 4, 5, 5, 6, 6, 4, 4, 5, 5)]
    assert [lp for _, lp, _ in rep] == [0] * 6 + [1] * 4  # looped is raised when the ring starts over
    a.halt()
    for _ in range(4):
        step.release()
    a.close()


def test_analyzer_seek(sdb):
    """suscan_analyzer_seek on a seekable (in-memory) source; callback sources refuse."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 4096, 1.0e6
    blk = N * 4
    na, nb = 400 * blk, 4 * blk
    t = np.arange(na + nb)
    x = np.where(t < na, np.exp(2j * np.pi * 0.1 * t), np.exp(2j * np.pi * 0.3 * t)).astype(np.complex64) * np.float32(0.2)
    a = Analyzer(fs, window_size=N, window="hann", psd_update_int=0.0, data=x, read_size=blk)
    assert a.seek(na / fs)                                   # jump over (most of) the first tone
    peaks = []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "PSD":
            peaks.append(int(np.argmax(m["psd"])))
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    a.close()
    assert name == "EOS"
    assert len(peaks) < (na + nb) // N                       # blocks were skipped
    assert peaks[-16:] == [round(0.3 * N)] * 16              # and the capture ended on its last four blocks
    b = Analyzer(fs, window_size=N, read=lambda priv, dst, n: 0, read_size=blk)
    assert not b.seek(0.5)
    b.close()


def test_analyzer_inspector_watermark(sdb):
    """suscan_analyzer_set_inspector_watermark_async (Default/Audio/AudioProcessor.cpp:745-747): batches are held
    back until they carry `watermark` samples; the stream itself is unchanged."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 8192, 1.0e6
    baud = fs / 100.0
    n = 8 * N * 8
    x, _ = synth.multi_carrier(n, fs, [("qpsk", 0.125 * fs, baud, -10.0, {})], noise_db=-50.0, seed=2)

    def run(watermark):
        go = Gate()
        pos = [0]

        def read(priv, dst, maxn):
            go.wait(30)
            take = min(maxn, n - pos[0])
            if take > 0:
                C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
                pos[0] += take
            return take

        a = Analyzer(fs, window_size=N, window="hann", psd_update_int=1.0, read=read, read_size=N * 8)
        assert a.read(5000)[0] == "SOURCE_INFO"
        assert go.parked.wait(10)          # requests sent from here on are handled after the first block
        a.open("psk", 0.125 * fs, 3 * baud, req_id=1)
        a.set_inspector_id(0, 5, req_id=2)
        cfg = sdb.InspectorConfig()
        sdb._check(sdb.load_library().sdb_inspector_config_default(C.byref(cfg), sdb.INSP["psk"], fs * 256 / N))
        cfg.baud, cfg.costas_order, cfg.bits_per_symbol, cfg.clock_type, cfg.clock_running = baud, 2, 2, 1, 1
        a.set_inspector_config(0, cfg, req_id=3)
        if watermark:
            a.set_inspector_watermark(0, watermark, req_id=4)
        go.set()
        sizes, soft = [], []
        while True:
            name, m = a.read(20000)
            assert name != "TIMEOUT"
            if name == "SAMPLES":
                sizes.append(len(m["samples"]))
                soft.append(m["samples"])
            elif name in ("EOS", "READ_ERROR", "HALT"):
                break
        a.close()
        return sizes, np.concatenate(soft)

    s0, a0 = run(0)
    s1, a1 = run(1500)
    assert len(s0) == 7 and all(600 < v < 700 for v in s0)             # one batch per block (~655 symbols)
    assert all(v >= 1500 for v in s1[:-1]) and len(s1) < len(s0)        # held back; the tail is flushed at EOS
    assert np.array_equal(a0.view(np.uint32), a1.view(np.uint32))


def test_analyzer_overridable_freq_and_bandwidth(sdb):
    """suscan_analyzer_set_inspector_freq_overridable / _bandwidth_overridable (Suscan/Analyzer.cpp:509-526): a raw
    inspector opened beside a tone is retuned onto it, then widened."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 4096, 1.0e6
    n = 10 * N * 4
    x = (0.3 * np.exp(2j * np.pi * 0.2 * np.arange(n))).astype(np.complex64)
    go = Gate()
    step = threading.Semaphore(0)
    pos = [0]

    def read(priv, dst, maxn):                        # one block per permit: the test decides when time advances
        go.wait(30)
        step.acquire(timeout=30)
        take = min(maxn, n - pos[0])
        if take > 0:
            C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
            pos[0] += take
        return take

    a = Analyzer(fs, window_size=N, window="hann", psd_update_int=0.0, read=read, read_size=N * 4)
    assert a.read(5000)[0] == "SOURCE_INFO"
    assert go.parked.wait(10)          # requests sent from here on are handled after the first block
    a.open("raw", 0.1 * fs, fs / 16.0, req_id=1)
    a.set_inspector_id(0, 9, req_id=2)
    go.set()

    def advance(k, batches=None):
        """let k blocks through (4 PSD frames each, then one SAMPLES batch per block once the inspector runs) and
        return the batches"""
        batches = k if batches is None else batches
        for _ in range(k):
            step.release()
        psds, out = 0, []
        while psds < 4 * k or len(out) < batches:
            name, m = a.read(20000)
            assert name not in ("TIMEOUT", "EOS", "READ_ERROR", "HALT")
            if name == "PSD":
                psds += 1
            elif name == "SAMPLES":
                out.append(m["samples"])
        return out

    assert advance(1, 0) == []                        # block 0: the requests are still queued
    off = advance(2)[-1]                              # tuned 0.1 fs away from the tone: almost nothing in the channel
    a.set_inspector_freq(0, 0.2 * fs)
    advance(1)                                        # the block in flight when the request lands: either tuning
    on = advance(2)[-1]                               # retuned: the tone sits at the channel centre
    p_off, p_on = np.mean(np.abs(off) ** 2), np.mean(np.abs(on) ** 2)
    assert p_on > 0.05 and p_off < 1e-3 * p_on
    a.set_inspector_bandwidth(0, fs / 4.0)
    advance(1)
    wide = advance(2)[-1]
    assert len(wide) == 4 * len(on)                   # four times the bandwidth: four times the channel rate
    a.halt()
    for _ in range(4):
        step.release()
    a.close()


def test_analyzer_source_options(sdb, oracle):
    """iq_reverse, baseband filter hook and throttle (Suscan/Analyzer.cpp:117-135, 238-244;
    Default/Source/SourceWidget.cpp:1156-1184)."""
    import time
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 4096, 1.0e6
    blocks = 6
    x = (0.2 * np.exp(2j * np.pi * 0.1 * np.arange(N * 4 * blocks))).astype(np.complex64)
    seen = []

    def recorder(samples, offset):               # the GUI's baseband recorder: sees every block, in order
        seen.append((offset, samples.copy()))
        return True

    def scaler(samples, offset):                 # a filter may rewrite the block the analyzer will see
        samples *= np.float32(0.5)
        return True

    go = Gate()
    pos = [0]

    def read(priv, dst, maxn):
        go.wait(30)
        take = min(maxn, len(x) - pos[0])
        if take > 0:
            C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
            pos[0] += take
        return take

    a = Analyzer(fs, window_size=N, window="hann", psd_update_int=0.0, read=read, read_size=N * 4)
    assert a.read(5000)[0] == "SOURCE_INFO"
    assert go.parked.wait(10)          # requests sent from here on are handled after the first block
    a.register_baseband_filter(recorder)
    a.register_baseband_filter(scaler)
    a.set_iq_reverse(True)
    a.set_throttle(2e6)                          # 5 more blocks of 16384 samples at 2 MS/s: >= 41 ms
    t0 = time.time()
    go.set()
    psd = []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "PSD":
            psd.append(m["psd"])
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    dt = time.time() - t0
    a.close()
    assert name == "EOS" and len(psd) == 4 * blocks
    # the options are requests: they apply from the block after the one in flight when they arrive, so compare
    # frame by frame with the two possible inputs and require the swapped one from some block on
    half = (x * np.float32(0.5)).astype(np.complex64)
    ref_n = oracle.psd_frames(half, N, "hann")
    ref_s = oracle.psd_frames((half.imag + 1j * half.real).astype(np.complex64), N, "hann")
    kinds = []
    for i, p in enumerate(psd):
        if np.array_equal(p.view(np.uint32), ref_s[i].view(np.uint32)):
            kinds.append("s")
        else:
            assert np.array_equal(p.view(np.uint32), ref_n[i].view(np.uint32)), i
            kinds.append("n")
    assert kinds[-4:] == ["s"] * 4 and "".join(kinds) == "n" * kinds.count("n") + "s" * kinds.count("s")
    # the mirrored tone of (Q, I) sits at -0.1 fs
    assert abs(int(np.argmax(psd[-1])) - round(0.9 * N)) <= 1
    assert [o for o, _ in seen] == [i * N * 4 for i in range(blocks)]
    assert np.array_equal(np.concatenate([s for _, s in seen]), x)       # recorder ran before the scaler
    assert dt >= 0.035


def test_analyzer_subcarrier_inspector(sdb, oracle):
    """open_ex with a parent handle (Default/GenericInspector/GenericInspector.cpp:502-525): a PSK inspector on a
    sub-carrier of a raw inspector's channel = two channelisers in series; symbols bit-identical to the oracle."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 8192, 1.0e6
    baud = fs / 512.0
    blocks, per_block = 6, N * 8
    n = blocks * per_block
    x, _ = synth.multi_carrier(n, fs, [("qpsk", 0.125 * fs + fs / 64.0, baud, -10.0, {})], noise_db=-55.0, seed=17)
    go = Gate()
    pos = [0]

    def read(priv, dst, maxn):
        go.wait(30)
        take = min(maxn, n - pos[0])
        if take > 0:
            C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
            pos[0] += take
        return take

    a = Analyzer(fs, window_size=N, window="hann", psd_update_int=1.0, read=read, read_size=per_block)
    assert a.read(5000)[0] == "SOURCE_INFO"
    assert go.parked.wait(10)          # requests sent from here on are handled after the first block
    a.open("raw", 0.125 * fs, fs / 8.0, req_id=1)                       # parent: 1024-point channel, fs / 8
    a.open("psk", fs / 64.0, 3 * baud, req_id=2, parent=0)             # child, relative to the parent's centre
    a.open("psk", 0.0, 1e9, req_id=3, parent=1)                        # wider than the parent's channel -> INVALID_CHANNEL
    a.open("psk", 0.0, 1000.0, req_id=4, parent=9)                     # unknown parent   -> WRONG_HANDLE
    a.set_inspector_id(1, 0x51, req_id=5)
    cfg = sdb.InspectorConfig()
    fs_par = fs / 8.0
    fs_ch = fs_par * 64 / 1024
    sdb._check(sdb.load_library().sdb_inspector_config_default(C.byref(cfg), sdb.INSP["psk"], fs_ch))
    kw = dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=fs_ch * 2e-3, mf_type=1, mf_rolloff=0.35,
              clock_type=1, clock_gain=0.1, clock_running=1)
    for k, v in kw.items():
        setattr(cfg, k, v)
    a.set_inspector_config(1, cfg, req_id=6)
    go.set()
    soft, hard, replies = [], [], []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "SAMPLES":
            assert m["inspector_id"] == 0x51
            soft.append(m["samples"]); hard.append(m["symbols"])
        elif name == "INSPECTOR":
            replies.append((m["kind"], m["req_id"], m["handle"]))
            if m["req_id"] == 2:
                assert abs(m["equiv_fs"] - fs_ch) < 1e-3
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    a.close()
    assert name == "EOS"
    assert [r[:2] for r in replies] == [("OPEN", 1), ("OPEN", 2), ("INVALID_CHANNEL", 3), ("WRONG_HANDLE", 4),
                                        ("SET_ID", 5), ("SET_CONFIG", 6)]
    assert [r[2] for r in replies[:2]] == [0, 1] and replies[3][2] == 9 and replies[4][2] == replies[5][2] == 1
    # oracle: baseband -> parent channel (window N) -> child channel (window 1024) -> psk chain
    f0p, bwp = float(np.float32(2.0 * np.pi * 0.125)), float(np.float32(2.0 * np.pi / 8.0))
    par = oracle.specttuner_run(x[per_block:], N, [dict(f0=f0p, bw=bwp, guard=1.0)])[0]
    f0c, bwc = float(np.float32(2.0 * np.pi * 0.125)), float(np.float32(2.0 * np.pi * (3 * baud) / fs_par))
    chi = oracle.specttuner_run(par, 1024, [dict(f0=f0c, bw=bwc, guard=1.0)])[0]
    rs, rh = oracle.inspector_run(oracle.insp_config("psk", fs_ch, **kw), chi)
    assert len(rs) > 300
    parity.assert_symbols_match(np.concatenate(soft), np.concatenate(hard), rs, rh, exact_soft=True)


def test_analyzer_nested_subcarrier_inspectors(sdb, oracle):
    """Sub-carrier inspection two levels deep (a sub-carrier tab opens sub-carrier tabs of its own,
    Default/GenericInspector/GenericInspector.cpp:502-525), with a spectrum source on the middle inspector: three
    channelisers in series; the grandchild's symbols are bit-identical to the oracle, the middle inspector emits
    SPECTRUM messages of its own channel."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 8192, 1.0e6
    baud = fs / 1024.0
    blocks, per_block = 6, N * 16
    n = blocks * per_block
    # carrier at fs/8 (parent centre) + fs/64 (child centre, in the parent) + fs/512 (grandchild centre, in the child)
    x, _ = synth.multi_carrier(n, fs, [("qpsk", 0.125 * fs + fs / 64.0 + fs / 512.0, baud, -10.0, {})], noise_db=-55.0,
                               seed=23)
    go = Gate()
    pos = [0]

    def read(priv, dst, maxn):
        go.wait(30)
        take = min(maxn, n - pos[0])
        if take > 0:
            C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
            pos[0] += take
        return take

    a = Analyzer(fs, window_size=N, window="hann", psd_update_int=1.0, read=read, read_size=per_block)
    assert a.read(5000)[0] == "SOURCE_INFO"
    assert go.parked.wait(10)          # requests sent from here on are handled after the first block
    fs_par = fs / 8.0                                                   # 1024-point channel
    fs_mid = fs_par / 4.0                                               # 256-point channel of the parent's 1024
    a.open("raw", 0.125 * fs, fs / 8.0, req_id=1)                       # handle 0
    a.open("raw", fs / 64.0, fs_par / 4.0, req_id=2, parent=0)          # handle 1, in the parent's channel
    a.open("psk", fs / 512.0, 3 * baud, req_id=3, parent=1)             # handle 2, in the child's channel
    a.set_inspector_id(2, 0x77, req_id=4)
    a.set_inspector_id(1, 0x66, req_id=5)
    a.set_spectrum_source(1, sdb.SPECTSRC["psd"], req_id=6)
    fs_ch = fs_mid * 32 / 256                                           # 3 baud -> 24 of 256 bins -> 32-point channel
    cfg = sdb.InspectorConfig()
    sdb._check(sdb.load_library().sdb_inspector_config_default(C.byref(cfg), sdb.INSP["psk"], fs_ch))
    kw = dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=fs_ch * 2e-3, mf_type=1, mf_rolloff=0.35,
              clock_type=1, clock_gain=0.1, clock_running=1)
    for k, v in kw.items():
        setattr(cfg, k, v)
    a.set_inspector_config(2, cfg, req_id=7)
    go.set()
    soft, hard, opens, spectra = [], [], {}, []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "SAMPLES" and m["inspector_id"] == 0x77:
            soft.append(m["samples"]); hard.append(m["symbols"])
        elif name == "INSPECTOR":
            if m["kind"] == "OPEN":
                opens[m["req_id"]] = (m["handle"], m["equiv_fs"])
            if m["kind"] == "SPECTRUM" and m["spectrum"] is not None:
                assert m["inspector_id"] == 0x66
                spectra.append(m)
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    a.close()
    assert name == "EOS"
    assert opens[1][0] == 0 and opens[2][0] == 1 and opens[3][0] == 2
    assert abs(opens[2][1] - fs_mid) < 1e-3 and abs(opens[3][1] - fs_ch) < 1e-3
    # oracle: baseband -> parent (window N) -> child (window 1024) -> grandchild (window 256) -> psk chain
    f0p, bwp = float(np.float32(2.0 * np.pi * 0.125)), float(np.float32(2.0 * np.pi / 8.0))
    par = oracle.specttuner_run(x[per_block:], N, [dict(f0=f0p, bw=bwp, guard=1.0)])[0]
    f0c, bwc = float(np.float32(2.0 * np.pi * 0.125)), float(np.float32(2.0 * np.pi / 4.0))
    mid = oracle.specttuner_run(par, 1024, [dict(f0=f0c, bw=bwc, guard=1.0)])[0]
    f0g, bwg = float(np.float32(2.0 * np.pi * 0.0625)), float(np.float32(2.0 * np.pi * (3 * baud) / fs_mid))
    gch = oracle.specttuner_run(mid, 256, [dict(f0=f0g, bw=bwg, guard=1.0)])[0]
    rs, rh = oracle.inspector_run(oracle.insp_config("psk", fs_ch, **kw), gch)
    assert len(rs) > 100
    parity.assert_symbols_match(np.concatenate(soft), np.concatenate(hard), rs, rh, exact_soft=True)
    # the middle inspector's spectrum source: the PSD of the tail of its channel samples of every block (SPEC U).
    # A fresh engine's first feed yields one window less, at every level: the first block gives H - 2 child hops.
    H = per_block // (N // 2)
    ends = np.cumsum([(H - 2) * 128] + [H * 128] * (blocks - 2))
    starts = [0] + list(ends[:-1])
    ns = len(spectra[0]["spectrum"])
    assert ns == 2048 and abs(spectra[0]["equiv_fs"] - fs_mid) < 1e-3 and len(spectra) == blocks - 1
    for sp, s0, e0 in zip(spectra, starts, ends):
        ref = oracle.spectsrc_frame("psd", ns, mid[s0:e0])
        assert np.array_equal(sp["spectrum"].view(np.uint32), ref.view(np.uint32))


def test_analyzer_spectrum_estimator_and_channel_messages(sdb, oracle):
    """kind=SPECTRUM / kind=ESTIMATOR inspector messages and MESSAGE_TYPE_CHANNEL lists
    (Suscan/Analyzer.cpp:539-565, GenericInspector.cpp:232-264, ChannelMessage.cpp:25-70)."""
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 8192, 1.0e6
    baud = fs / 64.0
    blocks, per_block = 5, N * 16
    n = blocks * per_block
    x, _ = synth.multi_carrier(n, fs, [("qpsk", 0.125 * fs, baud, -10.0, {})], noise_db=-50.0, seed=5)
    go = Gate()
    pos = [0]

    def read(priv, dst, maxn):
        go.wait(30)
        take = min(maxn, n - pos[0])
        if take > 0:
            C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
            pos[0] += take
        return take

    a = Analyzer(fs, window_size=N, window="blackmann_harris", psd_update_int=1.0, read=read, read_size=per_block,
                 channel_update_int=per_block / fs, alpha=0.5, gamma=0.5, snr=10.0, freq=433e6)
    assert a.read(5000)[0] == "SOURCE_INFO"
    assert go.parked.wait(10)          # requests sent from here on are handled after the first block
    a.open("psk", 0.125 * fs, 8 * baud, req_id=1)          # 8 samples per symbol at the channel rate
    a.set_inspector_id(0, 77, req_id=2)
    a.set_spectrum_source(0, sdb.SPECTSRC["exp_4"], req_id=3)
    a.estimator_cmd(0, sdb.ESTIMATOR["baud-nonlinear"], True, req_id=4)
    a.set_spectrum_source(0, 99, req_id=5)                  # unknown source -> WRONG_OBJECT
    go.set()
    spectra, estimates, channels, acks = [], [], [], []
    while True:
        name, m = a.read(20000)
        assert name != "TIMEOUT"
        if name == "INSPECTOR":
            if m["kind"] == "SPECTRUM" and m["spectrum"] is not None:
                assert m["spectsrc_id"] == sdb.SPECTSRC["exp_4"] and m["inspector_id"] == 77
                spectra.append(m)
            elif m["kind"] == "ESTIMATOR" and m["req_id"] == 0:
                estimates.append(m)
            else:
                acks.append((m["kind"], m["req_id"]))
                if m["kind"] == "OPEN":
                    assert (m["spectsrc_count"], m["estimator_count"]) == (9, 2)
        elif name == "CHANNEL":
            channels.append(m["channels"])
        elif name in ("EOS", "READ_ERROR", "HALT"):
            break
    a.close()
    assert name == "EOS"
    assert acks == [("OPEN", 1), ("SET_ID", 2), ("SPECTRUM", 3), ("ESTIMATOR", 4), ("WRONG_OBJECT", 5)]
    # requests are applied at block boundaries: blocks 2.. carry a spectrum and an estimate each
    assert len(spectra) >= 3 and len(estimates) >= 3
    fs_ch = spectra[0]["equiv_fs"]
    assert spectra[0]["samp_rate"] == int(fs_ch) and len(spectra[0]["spectrum"]) in (1024, 2048, 4096)
    for e in estimates:
        assert e["estimator_id"] == 1 and abs(e["value"] - baud) / baud < 0.02
    # the 4th power of QPSK: a line at 4x the residual carrier offset (here ~0) dominates the spectrum
    s = spectra[-1]["spectrum"]
    k = int(np.argmax(s))
    assert min(k, len(s) - k) <= 2
    # one CHANNEL list per block; the carrier is found at its place, reported in absolute frequency
    assert len(channels) == blocks
    hit = [c for c in channels[-1] if c["f_lo"] <= 433e6 + 0.125 * fs <= c["f_hi"]]
    assert len(hit) == 1 and 0.7 * baud < hit[0]["bw"] < 2.5 * baud and hit[0]["snr"] > 10.0

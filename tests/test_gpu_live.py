"""-m gpu: source-side options and live re-planning of the worker loop.
* DC removal (SPEC R) bit-identical to the oracle, float32 and 8-bit sources, block by block;
* sdb_engine_migrate: a re-planned engine continues the loops of the inspectors that stay open -- opening, closing,
  retuning or reconfiguring ONE inspector leaves the symbol streams of the others bit-identical to an undisturbed run
  (suscan keeps them running too: Suscan/Analyzer.cpp:459-537; the GUI fires set_inspector_config on every slider
  move, Default/GenericInspector/InspectorCtl/InspectorCtl.cpp:68);
* one process, two GPUs (per-device kernel attributes)."""
import ctypes as C

import numpy as np
import pytest

import parity
from sigdigger_b200 import synth

pytestmark = pytest.mark.gpu


def _dc_oracle(oracle, blocks, fs):
    L = oracle.lib()
    c = (C.c_float * 2)(0.0, 0.0)
    out = []
    for b in blocks:
        b = np.ascontiguousarray(b, np.complex64)
        y = np.empty_like(b)
        alpha = np.float32(1.0 - np.exp(-len(b) / (fs * 0.1)))
        L.sdo_dc_remove(c, oracle.ptr(b), oracle.ptr(y), C.c_size_t(len(b)), C.c_float(float(alpha)))
        out.append(y)
    return out


@pytest.mark.parametrize("fmt", ["f32", "u8"])
def test_dc_removal_matches_oracle(sdb, oracle, fmt):
    N, fs = 8192, 1.0e6
    n_blk, blocks = N * 4, 5
    rng = np.random.default_rng(3)
    x = (0.05 * (rng.standard_normal(n_blk * blocks) + 1j * rng.standard_normal(n_blk * blocks)) + (0.21 - 0.13j)).astype(np.complex64)
    x += (0.1 * np.exp(2j * np.pi * 0.0731 * np.arange(len(x)))).astype(np.complex64)
    if fmt == "u8":
        q = np.clip(np.round(x.view(np.float32) * 128.0 + 128.0), 0, 255).astype(np.uint8).reshape(-1, 2)
        xf = np.ascontiguousarray((q.astype(np.float32) - 128.0) / 128.0).view(np.complex64).reshape(-1)
        feed = q
    else:
        xf, feed = x, x
    e = sdb.Engine(n_streams=1, psd_size=N, psd_window="hann", max_feed=n_blk, samp_rate=fs, flags=sdb.FLAG_DC_REMOVE,
                   input_format=fmt)
    e.commit()
    ref_blocks = _dc_oracle(oracle, [xf[i * n_blk:(i + 1) * n_blk] for i in range(blocks)], fs)
    for i in range(blocks):
        blk = feed[i * n_blk:(i + 1) * n_blk]
        e.feed(blk.reshape(1, n_blk, 2) if fmt == "u8" else blk.reshape(1, -1))
        psd = e.read_psd()[0]
        ref = oracle.psd_frames(ref_blocks[i], N, "hann")
        assert np.array_equal(psd.view(np.uint32), ref.view(np.uint32)), (fmt, i)
    # the estimate converges: the DC bin of the last block is far below that of the first
    assert psd[-1][0] < 0.15 * oracle.psd_frames(xf[:N], N, "hann")[0][0]


def _two_channel_engine(sdb, N, n_blk, fs, cfgs, freqs, baud):
    e = sdb.Engine(n_streams=1, psd_size=N, psd_window="blackmann_harris", max_feed=n_blk, samp_rate=fs)
    hs = []
    for f, kw in zip(freqs, cfgs):
        h = e.open_channel(float(np.float32(2 * np.pi * (f / fs % 1.0))), float(np.float32(2 * np.pi * 3 * baud / fs)), 1.0)
        e.set_inspector(h, "psk", **kw)
        hs.append(h)
    e.commit()
    return e, hs


def test_migrate_keeps_running_inspectors_bit_identical(sdb, oracle):
    """Inspector B runs undisturbed while inspector A is reconfigured (loop bandwidth: loops continue), retuned
    (channel moves: loops continue, cross-fade restarts), closed and another one opened."""
    N, fs = 8192, 1.0e6
    baud = fs / 100.0
    n_blk, blocks = N * 4, 8
    fA, fB, fC = 0.125 * fs, -0.2 * fs, 0.3 * fs
    x, _ = synth.multi_carrier(n_blk * blocks, fs, [("qpsk", fA + 3.0, baud, -12.0, {}), ("qpsk", fB - 2.0, baud, -12.0, {}),
                                                   ("qpsk", fC, baud, -12.0, {})], noise_db=-55.0, seed=41)
    fs_ch = fs * 256 / N
    kw = dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=fs_ch * 2e-3, mf_type=1, mf_rolloff=0.35,
              clock_type=1, clock_gain=0.1)
    kw2 = dict(kw, loop_bw=fs_ch * 5e-3)
    blk = lambda i: x[i * n_blk:(i + 1) * n_blk].reshape(1, -1)
    # undisturbed run: A and B for all blocks
    e0, (a0, b0) = _two_channel_engine(sdb, N, n_blk, fs, [kw, kw], [fA, fB], baud)
    und = []
    for i in range(blocks):
        e0.feed(blk(i))
        und.append(e0.read_symbols(0, b0))
    # disturbed run
    L = sdb.load_library()
    got = []
    e1, (a1, b1) = _two_channel_engine(sdb, N, n_blk, fs, [kw, kw], [fA, fB], baud)
    plan = {2: ([kw2, kw], [fA, fB]),                 # afc.loop-bw of A changes
            4: ([kw2, kw], [fA + 7000.0, fB]),        # A dragged by 7 kHz
            6: ([kw, kw], [fC, fB])}                  # A closed, C opened in its slot
    cur, bh = e1, b1
    keep = []
    for i in range(blocks):
        if i in plan:
            cfgs, freqs = plan[i]
            e2, (a2, b2) = _two_channel_engine(sdb, N, n_blk, fs, cfgs, freqs, baud)
            omap = (C.c_int32 * 2)(a1 if i != 6 else -1, bh)
            sdb._check(L.sdb_engine_migrate_map(e2._h, cur._h, omap, 2))
            keep.append(cur)
            cur, a1, bh = e2, a2, b2
        cur.feed(blk(i))
        got.append(cur.read_symbols(0, bh))
    for i in range(blocks):
        assert np.array_equal(got[i][0].view(np.uint32), und[i][0].view(np.uint32)), i
        assert np.array_equal(got[i][1], und[i][1]), i
    # and A itself kept its loops across the loop-bandwidth change: equal to an oracle that switches loop_bw at block 2
    # is out of the oracle's reach (no live reconfiguration there); check continuity instead: symbol count per block
    assert sum(len(g[1]) for g in got) == sum(len(u[1]) for u in und)


def test_analyzer_reconfigure_one_inspector_leaves_the_other_untouched(sdb, oracle):
    """Through the asynchronous analyzer: set_inspector_config on A (loop bandwidth) and
    set_inspector_freq_overridable on A mid-stream; B's SAMPLES equal an undisturbed run bit for bit; acks in order."""
    import threading
    from sigdigger_b200.analyzer import Analyzer
    N, fs = 8192, 1.0e6
    baud = fs / 100.0
    blocks, per_block = 8, N * 4
    n = blocks * per_block
    fA, fB = 0.125 * fs, -0.2 * fs
    x, _ = synth.multi_carrier(n, fs, [("qpsk", fA + 3.0, baud, -12.0, {}), ("qpsk", fB - 2.0, baud, -12.0, {})],
                               noise_db=-55.0, seed=43)
    fs_ch = fs * 256 / N
    kw = dict(baud=baud, costas_order=2, bits_per_symbol=2, loop_bw=fs_ch * 2e-3, mf_type=1, mf_rolloff=0.35,
              clock_type=1, clock_gain=0.1, clock_running=1)

    def run(disturb):
        step = threading.Semaphore(0)
        pos = [0]

        def read(priv, dst, maxn):
            step.acquire()
            take = min(maxn, n - pos[0])
            if take > 0:
                C.memmove(dst, x.ctypes.data + 8 * pos[0], 8 * take)
                pos[0] += take
            return take

        a = Analyzer(fs, window_size=N, window="blackmann_harris", psd_update_int=1.0, read=read, read_size=per_block)
        name, _ = a.read(5000)
        assert name == "SOURCE_INFO"
        # the worker now sits in its first read: both inspectors exist from block 2 on, in either run
        a.open("psk", fA, 3 * baud, req_id=1)
        a.open("psk", fB, 3 * baud, req_id=2)
        a.set_inspector_id(0, 100, req_id=3)
        a.set_inspector_id(1, 200, req_id=4)
        cfg = sdb.InspectorConfig()
        sdb._check(sdb.load_library().sdb_inspector_config_default(C.byref(cfg), sdb.INSP["psk"], fs_ch))
        for k, v in kw.items():
            setattr(cfg, k, v)
        a.set_inspector_config(0, cfg, req_id=5)
        a.set_inspector_config(1, cfg, req_id=6)
        acks, outB = [], []

        def consumed(blocks_read, timeout=20.0):
            import time
            t0 = time.time()
            while pos[0] < blocks_read * per_block and time.time() - t0 < timeout:
                time.sleep(0.002)
            assert pos[0] >= blocks_read * per_block

        for _ in range(3):
            step.release()
        consumed(3)
        if disturb:                                   # lands before block 4 or 5: mid-stream either way
            cfg.loop_bw = fs_ch * 6e-3
            a.set_inspector_config(0, cfg, req_id=7)
        for _ in range(2):
            step.release()
        consumed(5)
        if disturb:
            a.set_inspector_freq(0, fA + 5000.0)
        for _ in range(blocks - 5 + 1):
            step.release()
        done = False
        while not done:
            name, m = a.read(30000)
            assert name != "TIMEOUT"
            if name == "SAMPLES" and m["inspector_id"] == 200:
                outB.append(m["samples"])
            elif name == "INSPECTOR":
                acks.append((m["kind"], m["req_id"]))
            elif name in ("EOS", "READ_ERROR", "HALT"):
                done = True
        a.close()
        return np.concatenate(outB), acks

    ref, acks0 = run(False)
    got, acks1 = run(True)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert [r for _, r in acks0] == [1, 2, 3, 4, 5, 6]
    assert [r for _, r in acks1] == [1, 2, 3, 4, 5, 6, 7]


def test_two_engines_on_two_gpus_in_one_process(sdb, oracle):
    """One process, two devices (what a single GUI process with two analyzers does): the > 48 KB dynamic
    shared-memory attributes of the kernels are per device, not per process."""
    if sdb.device_count() < 2:
        pytest.skip("needs two GPUs")
    N = 65536
    n = N * 2
    x, _ = synth.multi_carrier(n, 1.0, [("qpsk", 0.125, 0.01, -10.0, {})], noise_db=-50.0, seed=5)
    outs = []
    for dev in (0, 1):
        e = sdb.Engine(n_streams=1, psd_size=N, psd_window="blackmann_harris", max_feed=n, samp_rate=1.0, device=dev)
        h = e.open_channel(float(np.float32(2 * np.pi * 0.125)), float(np.float32(2 * np.pi * 0.03)), 1.0)
        e.set_inspector(h, "psk", baud=0.01, costas_order=2, bits_per_symbol=2, loop_bw=3e-5, mf_type=1, clock_type=1,
                        clock_gain=0.1)
        e.commit()
        e.feed(x.reshape(1, -1))
        outs.append((e.read_psd()[0].copy(), e.read_symbols(0, h)))
    assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))
    assert np.array_equal(outs[0][1][0].view(np.uint32), outs[1][1][0].view(np.uint32))

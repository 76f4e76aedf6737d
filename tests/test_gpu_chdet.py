"""Channel detector (SPEC K): CUDA path through the C-ABI against the oracle, bit-exact
(bin ranges, levels, noise floor), engine-integrated on the engine's own PSD and stand-alone on arbitrary spectra."""
import numpy as np
import pytest

from sigdigger_b200 import synth

pytestmark = pytest.mark.gpu


def _scene(S, N, frames, seed):
    """S streams: noise + a few carriers of different widths."""
    n = N * frames
    rng = np.random.default_rng(seed)
    xs = []
    truth = []
    for s in range(S):
        x = synth.awgn(n, 0.01, rng)
        tr = []
        for f, sps, amp in ((0.11 + 0.01 * s, 16.0, 0.3), (-0.27, 64.0, 0.2), (0.40, 8.0, 0.25)):
            sig, _ = synth.psk_signal(n, sps, order=4, seed=seed + 7 * s + int(sps))
            x = x + amp * synth.mix(sig, f, 0.1)
            tr.append((f, 1.0 / sps))
        xs.append(x)
        truth.append(tr)
    return np.stack(xs).astype(np.complex64), truth


def _same(got, ref):
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert (g.bin_lo, g.bin_hi) == (r[0], r[1])
        for a, b in ((g.S0, r[2]), (g.N0, r[3]), (g.snr, r[4])):
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)


@pytest.mark.parametrize("N", [4096, 65536])
def test_engine_channel_detector_bit_exact(sdb, oracle, N):
    S, frames, feeds = 3, 4, 3
    x, truth = _scene(S, N, frames * feeds, seed=N % 97)
    e = sdb.Engine(n_streams=S, psd_size=N, psd_window="blackmann_harris", max_feed=N * frames)
    kw = dict(alpha=0.25, gamma=0.5, snr=8.0, min_bins=max(2, N // 2048))
    e.set_channel_detector(beta=0.1, **kw)
    e.commit()
    dets = [oracle.ChannelDetector(N, kw["alpha"], kw["gamma"], kw["snr"], kw["min_bins"]) for _ in range(S)]
    for f in range(feeds):
        e.feed(x[:, f * N * frames:(f + 1) * N * frames])
        psd = e.read_psd()
        for s in range(S):
            ref, rtot = dets[s].feed(psd[s])
            got, tot = e.read_channels(s)
            assert tot == rtot
            _same(got, ref)
    # after averaging the three carriers are found where they were put, roughly as wide as their symbol rate
    got, _ = e.read_channels(0, center_freq=100e6)
    for f, baud in truth[0]:
        hit = [c for c in got if c.f_lo - 100e6 <= f <= c.f_hi - 100e6]
        assert len(hit) == 1, (f, [(c.f_lo, c.f_hi) for c in got])
        assert 0.7 * baud < hit[0].bw < 2.5 * baud
        assert hit[0].fc == pytest.approx(0.5 * (hit[0].f_lo + hit[0].f_hi))
        assert hit[0].snr > 8.0
    for d in dets:
        d.close()


def test_standalone_detector_edges_caps_and_noise(sdb, oracle):
    import torch
    N, S = 8192, 4
    rng = np.random.default_rng(3)
    psd = rng.exponential(1.0, (S, 2, N)).astype(np.float32)
    psd[0, :, :] += 0.0
    psd[1, :, N // 2 - 40:N // 2 + 7] += 50.0          # touches both array edges in shifted order (around index N/2)
    psd[2, :, 0:30] += 40.0
    psd[2, :, N - 25:] += 40.0                          # straddles DC in unshifted order = one run in shifted order
    psd[3, :, ::3] += 30.0                              # thousands of 1-bin runs: all rejected by min_bins
    d = sdb.ChannelDetector(N, S, alpha=0.5, gamma=1.0, snr=6.0, min_bins=3)
    dets = [oracle.ChannelDetector(N, 0.5, 1.0, 6.0, 3) for _ in range(S)]
    t = torch.from_numpy(psd).cuda()
    for rep in range(2):
        d.feed(t)
        for s in range(S):
            ref, rtot = dets[s].feed(psd[s])
            got, tot = d.read(s, samp_rate=1.0)
            assert tot == rtot
            _same(got, ref)
    got, _ = d.read(2)
    assert any(c.bin_lo <= N // 2 - 25 and c.bin_hi >= N // 2 + 30 for c in got)
    got1, _ = d.read(1)
    assert got1 and got1[0].bin_lo == 0 and got1[-1].bin_hi == N      # runs clipped at the band edges, not merged
    got3, _ = d.read(3)
    assert all(c.bin_hi - c.bin_lo >= 3 for c in got3)               # min_bins removed the 1- and 2-bin runs
    d.close()
    for o in dets:
        o.close()


def test_many_channels_cap(sdb, oracle):
    import torch
    N = 16384
    psd = np.full((1, 1, N), 1.0, np.float32)
    for i in range(400):                                 # 400 carriers of 8 bins: more than the 256-entry report cap
        psd[0, 0, 20 + 40 * i:28 + 40 * i] = 100.0
    d = sdb.ChannelDetector(N, 1, alpha=1.0, gamma=1.0, snr=4.0, min_bins=2)
    d.feed(torch.from_numpy(psd).cuda())
    got, tot = d.read(0)
    o = oracle.ChannelDetector(N, 1.0, 1.0, 4.0, 2)
    ref, rtot = o.feed(psd[0])
    assert tot == rtot == 400 and len(got) == len(ref) == 256
    _same(got, ref)
    d.close()
    o.close()

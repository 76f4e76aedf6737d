"""Host logic of the multi-GPU panoramic sweep (BASELINE.json configs[4]).

Replaces, for the wide-spectrum mode, what `SigDigger::Scanner` does on the GUI thread per PSD message
(`Panoramic/Scanner.cpp:503-523`: `view.feed(psd, nullptr, fftSize, fc)`), spread over ranks:

    rank r : hops [lo_r, hi_r)  --PSD-->  project  --(j0, nb, va, vc)--+
                                 `-> channel detector (optional)         | one all_gather (NCCL over NVLink)
    rank 0 : accumulate(all contributions, in global hop order) + fill <-+

Per-bin state of the SpectrumView depends only on that bin's own contributions in hop order, so gathering the
contribution lists in rank order (= hop order, shards are contiguous) and applying them on rank 0 reproduces
the reference's sequential feed() exactly.  Only this exchange uses a collective; the PSDs never move.

With `detect=...` every rank also runs the channel detector (SPEC K; the analyzer's CHANNEL message,
Suscan/Analyzer.cpp:570-577) over its own hops, each hop an independent detector state as in a fresh
wide-spectrum analyzer, and the per-hop channel lists (absolute frequencies) ride the same gather.
"""
import numpy as np

CHANNEL_FIELDS = ("fc", "f_lo", "f_hi", "bw", "snr", "S0", "N0")


def shard(n_hops, world, rank):
    """Contiguous, balanced hop range of `rank` (progressive sweep => contiguous in frequency)."""
    base, rem = divmod(n_hops, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def padded_len(n_hops, world):
    """Every rank contributes the same number of rows to all_gather: the largest shard."""
    return (n_hops + world - 1) // world


def gather_order(n_hops, world):
    """Row indices into the [world * padded_len] gathered array that hold real hops, in global hop order."""
    pl = padded_len(n_hops, world)
    idx = []
    for r in range(world):
        lo, hi = shard(n_hops, world, r)
        idx.extend(r * pl + i for i in range(hi - lo))
    return np.asarray(idx, dtype=np.int64)


def pack_channels(per_hop, pl, cap):
    """Host side of the channel gather: per-hop lists of DetectedChannel -> ([pl, cap, 7] f64, [pl] i32)."""
    rows = np.zeros((pl, cap, len(CHANNEL_FIELDS)), np.float64)
    cnt = np.zeros(pl, np.int32)
    for i, chans in enumerate(per_hop):
        cnt[i] = min(len(chans), cap)
        for j, c in enumerate(chans[:cap]):
            rows[i, j] = [getattr(c, f) for f in CHANNEL_FIELDS]
    return rows, cnt


def unpack_channels(rows, cnt):
    """Gathered (rows, cnt) in global hop order -> list (per hop) of dicts."""
    return [[dict(zip(CHANNEL_FIELDS, map(float, rows[h, j]))) for j in range(int(cnt[h]))]
            for h in range(len(cnt))]


def sweep(sdb, torch, dist, x_local, centers_all, psd_size, window, view_range, fft_bandwidth, rel_bw=0.5,
          device=0, detect=None, channel_cap=64):
    """One sweep. x_local: [local_hops, frames * psd_size] complex64 cuda tensor of this rank's hops.
    Returns (psd, accum, count) numpy arrays on rank 0, None elsewhere.  With `detect` (a dict of
    set_channel_detector arguments, e.g. {"alpha": 1.0, "snr": 6.0}) rank 0 gets
    (psd, accum, count, channels), channels[h] = list of dicts (CHANNEL_FIELDS, absolute Hz) for hop h."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_hops = len(centers_all)
    lo, hi = shard(n_hops, world, rank)
    assert x_local.shape[0] == hi - lo
    n_local = max(1, hi - lo)
    eng = sdb.Engine(n_streams=n_local, psd_size=psd_size, psd_window=window, max_feed=x_local.shape[1],
                     samp_rate=fft_bandwidth, device=device, flags=0 if detect is not None else sdb.FLAG_PSD_SHIFT_DB)
    if detect is not None:
        eng.set_channel_detector(**detect)
    eng.commit()
    view = sdb.SpectrumView(view_range[0], view_range[1], fft_bandwidth, rel_bw, device=device)
    mb = view.max_bins
    pl = padded_len(n_hops, world)
    j0 = torch.zeros(pl, dtype=torch.int32, device="cuda")
    nb = torch.zeros(pl, dtype=torch.int32, device="cuda")
    va = torch.zeros((pl, mb), dtype=torch.float32, device="cuda")
    vc = torch.zeros((pl, mb), dtype=torch.float32, device="cuda")
    per_hop = []
    if hi > lo:
        eng.feed(x_local)
        frames = x_local.shape[1] // psd_size
        psd_ptr = eng.psd_device_ptr
        if detect is not None:
            # the detector consumed the linear PSD; the view wants the PSDMessage layout of the last frame
            db = torch.empty((hi - lo, psd_size), dtype=torch.float32, device="cuda")
            lin_last = psd_ptr + (frames - 1) * psd_size * 4
            if frames == 1:
                sdb.psd_shift_db(lin_last, db.data_ptr(), hi - lo, psd_size)
            else:
                for s in range(hi - lo):
                    sdb.psd_shift_db(lin_last + s * frames * psd_size * 4, db.data_ptr() + s * psd_size * 4, 1,
                                     psd_size)
            psd_ptr = db.data_ptr()
            per_hop = [eng.read_channels(s, center_freq=float(centers_all[lo + s]), cap=channel_cap)[0]
                       for s in range(hi - lo)]
        else:
            assert frames == 1, "without a detector the sweep takes one PSD frame per hop"
        view.project(psd_ptr, psd_size, centers_all[lo:hi])
        # device-to-device copy of the contribution lists into the (padded) send buffers
        view.contrib_copy(j0.data_ptr(), nb.data_ptr(), va.data_ptr(), vc.data_ptr(), hi - lo)
    torch.cuda.synchronize()
    bufs = [j0, nb, va, vc]
    if detect is not None:
        rows, cnt = pack_channels(per_hop, pl, channel_cap)
        bufs += [torch.from_numpy(rows).cuda(), torch.from_numpy(cnt).cuda()]
    if world > 1:
        outs = []
        for t in bufs:
            g = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device="cuda")
            dist.all_gather_into_tensor(g, t)
            outs.append(g.reshape((world * pl,) + tuple(t.shape[1:])))
        order = torch.from_numpy(gather_order(n_hops, world)).cuda()
        bufs = [o.index_select(0, order).contiguous() for o in outs]
    else:
        bufs = [t[:n_hops].contiguous() for t in bufs]
    if rank != 0:
        return None
    j0, nb, va, vc = bufs[:4]
    view.accumulate(j0.data_ptr(), nb.data_ptr(), va.data_ptr(), vc.data_ptr(), n_hops)
    out = view.read()
    if detect is not None:
        out = tuple(out) + (unpack_channels(bufs[4].cpu().numpy(), bufs[5].cpu().numpy()),)
    return out

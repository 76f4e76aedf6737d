"""Host logic of the multi-GPU panoramic sweep (BASELINE.json configs[4]).

Replaces, for the wide-spectrum mode, what `SigDigger::Scanner` does on the GUI thread per PSD message
(`Panoramic/Scanner.cpp:503-523`: `view.feed(psd, nullptr, fftSize, fc)`), spread over ranks:

    rank r : hops [lo_r, hi_r)  --PSD-->  project  --(j0, nb, va, vc)--+
                                                                         | one all_gather (NCCL over NVLink)
    rank 0 : accumulate(all contributions, in global hop order) + fill <-+

Per-bin state of the SpectrumView depends only on that bin's own contributions in hop order, so gathering the
contribution lists in rank order (= hop order, shards are contiguous) and applying them on rank 0 reproduces
the reference's sequential feed() exactly.  Only this exchange uses a collective; the PSDs never move.
"""
import numpy as np


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


def sweep(sdb, torch, dist, x_local, centers_all, psd_size, window, view_range, fft_bandwidth, rel_bw=0.5,
          device=0):
    """One sweep. x_local: [local_hops, psd_size] complex64 cuda tensor of this rank's hops.
    Returns (psd, accum, count) numpy arrays on rank 0, None elsewhere."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_hops = len(centers_all)
    lo, hi = shard(n_hops, world, rank)
    assert x_local.shape[0] == hi - lo
    eng = sdb.Engine(n_streams=max(1, hi - lo), psd_size=psd_size, psd_window=window, max_feed=psd_size,
                     device=device, flags=sdb.FLAG_PSD_SHIFT_DB)
    eng.commit()
    view = sdb.SpectrumView(view_range[0], view_range[1], fft_bandwidth, rel_bw, device=device)
    mb = view.max_bins
    pl = padded_len(n_hops, world)
    j0 = torch.zeros(pl, dtype=torch.int32, device="cuda")
    nb = torch.zeros(pl, dtype=torch.int32, device="cuda")
    va = torch.zeros((pl, mb), dtype=torch.float32, device="cuda")
    vc = torch.zeros((pl, mb), dtype=torch.float32, device="cuda")
    if hi > lo:
        eng.feed(x_local)
        view.project(eng.psd_device_ptr, psd_size, centers_all[lo:hi])
        # device-to-device copy of the contribution lists into the (padded) send buffers
        view.contrib_copy(j0.data_ptr(), nb.data_ptr(), va.data_ptr(), vc.data_ptr(), hi - lo)
    torch.cuda.synchronize()
    if world > 1:
        outs = []
        for t in (j0, nb, va, vc):
            g = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device="cuda")
            dist.all_gather_into_tensor(g, t)
            outs.append(g.reshape((world * pl,) + tuple(t.shape[1:])))
        order = torch.from_numpy(gather_order(n_hops, world)).cuda()
        j0, nb, va, vc = [o.index_select(0, order).contiguous() for o in outs]
    else:
        j0, nb, va, vc = j0[:n_hops], nb[:n_hops], va[:n_hops].contiguous(), vc[:n_hops].contiguous()
    if rank != 0:
        return None
    view.accumulate(j0.data_ptr(), nb.data_ptr(), va.data_ptr(), vc.data_ptr(), n_hops)
    return view.read()

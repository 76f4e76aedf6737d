"""Host logic of the multi-GPU panoramic sweep (BASELINE.json configs[4]).

Replaces, for the wide-spectrum mode, what `SigDigger::Scanner` does on the GUI thread per PSD message
(`Panoramic/Scanner.cpp:503-523`: `view.feed(psd, nullptr, fftSize, fc)`), spread over ranks:

    rank r : hops [lo_r, hi_r)  --PSD-->  project  --(j0, nb, va, vc)--+
                                 `-> channel detector (optional)         | one gather to rank 0 (NCCL over NVLink)
    rank 0 : accumulate(all contributions, in global hop order) + fill <-+

The sweep itself is C++ behind the C-ABI (csrc/panoramic.cu, sdb_panoramic_*); this module only shards the hop list
for the tests / the bench and carries the 128-byte NCCL id between the ranks.

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


def exchange_unique_id(sdb, torch, dist):
    """rank 0 creates the NCCL id of the sweep's own communicator; torch.distributed only carries its 128 bytes."""
    if dist is None or dist.get_world_size() == 1:
        return None
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(sdb.panoramic_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def sweep(sdb, torch, dist, x_local, centers_all, psd_size, window, view_range, fft_bandwidth, rel_bw=0.5,
          device=0, detect=None, channel_cap=64, handle=None):
    """One sweep through the C-ABI (sdb_panoramic_*: engine PSDs, projection, NCCL gather to rank 0, stitch).
    x_local: [local_hops, psd_size] complex64 cuda tensor of this rank's hops (sdb.Panoramic.shard).
    Returns (psd, accum, count) numpy arrays on rank 0, None elsewhere; with `detect` (dict: alpha / gamma / snr)
    rank 0 gets (psd, accum, count, channels), channels[h] = list of dicts (CHANNEL_FIELDS, absolute Hz) for hop h.
    `handle`: an sdb.Panoramic to reuse (the bench keeps one across steps)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_hops = len(centers_all)
    lo, hi = shard(n_hops, world, rank)
    assert x_local.shape[0] == hi - lo and (hi == lo or x_local.shape[1] % psd_size == 0)
    frames = max(1, x_local.shape[1] // psd_size) if hi > lo else 1
    p = handle
    if p is None:
        p = sdb.Panoramic(psd_size, window, fft_bandwidth, view_range, rel_bw, device=device, rank=rank, world=world,
                          unique_id=exchange_unique_id(sdb, torch, dist), detect=detect, channel_cap=channel_cap,
                          frames_per_hop=frames)
    p.sweep(x_local.contiguous(), centers_all)
    if rank != 0:
        return None
    out = p.read()
    if detect is not None:
        chans = [[{f: float(getattr(c, f)) for f in CHANNEL_FIELDS} for c in p.read_channels(h)] for h in range(n_hops)]
        out = tuple(out) + (chans,)
    return out

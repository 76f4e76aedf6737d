"""CPU tests of the N>1 host logic with the gloo backend (world_size 2, 3): hop sharding, the all_gather
of the panoramic contribution lists and their re-ordering into global hop order, and the weak-scaling
aggregation rule of bench.py (max of rank times)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sigdigger_b200 import panoramic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_hops, mb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = panoramic.shard(n_hops, world, rank)
    pl = panoramic.padded_len(n_hops, world)
    # fake contribution lists: row h carries its global hop id
    j0 = torch.full((pl,), -1, dtype=torch.int32)
    va = torch.zeros((pl, mb), dtype=torch.float32)
    for i, h in enumerate(range(lo, hi)):
        j0[i] = h
        va[i] = h + torch.arange(mb) * 1e-3
    gj = [torch.empty_like(j0) for _ in range(world)]
    gv = [torch.empty_like(va) for _ in range(world)]
    dist.all_gather(gj, j0)
    dist.all_gather(gv, va)
    order = torch.from_numpy(panoramic.gather_order(n_hops, world))
    j_all = torch.cat(gj).index_select(0, order)
    v_all = torch.cat(gv).index_select(0, order)
    # per-hop channel lists (cfg5 "per-GPU channel detector, gather to rank 0"): hop h reports h % 3 channels
    class Ch:
        def __init__(self, h, j):
            self.fc, self.f_lo, self.f_hi, self.bw = 1e9 + h, 1e9 + h - j - 1, 1e9 + h + j + 1, 2.0 * (j + 1)
            self.snr, self.S0, self.N0 = 10.0 + j, -20.0 - h, -60.0
    rows, cnt = panoramic.pack_channels([[Ch(h, j) for j in range(h % 3)] for h in range(lo, hi)], pl, 4)
    rows, cnt = torch.from_numpy(rows), torch.from_numpy(cnt)
    gr = [torch.empty_like(rows) for _ in range(world)]
    gc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(gr, rows)
    dist.all_gather(gc, cnt)
    chans = panoramic.unpack_channels(torch.cat(gr).index_select(0, order).numpy(),
                                      torch.cat(gc).index_select(0, order).numpy())
    # weak-scaling timing rule: whole-job time = max over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((j_all.numpy(), v_all.numpy(), float(t.item()), chans))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_hops", [(2, 16), (2, 7), (3, 10)])
def test_gather_restores_global_hop_order(world, n_hops):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_hops, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    j_all, v_all, tmax, chans = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert list(j_all) == list(range(n_hops))
    assert np.allclose(v_all[:, 0], np.arange(n_hops))
    assert tmax == 10.0 + world - 1
    assert [len(c) for c in chans] == [h % 3 for h in range(n_hops)]
    for h, cs in enumerate(chans):
        for j, c in enumerate(cs):
            assert c["fc"] == 1e9 + h and c["bw"] == 2.0 * (j + 1) and c["snr"] == 10.0 + j and c["S0"] == -20.0 - h


def test_shards_are_contiguous_and_cover():
    for world in (1, 2, 3, 4, 8):
        for n in (1, 7, 8, 1024, 1025):
            edges = [panoramic.shard(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= panoramic.padded_len(n, world)
            assert len(panoramic.gather_order(n, world)) == n

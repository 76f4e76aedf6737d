"""torchrun worker: multi-GPU panoramic sweep (NCCL all_gather of the contribution lists) checked on rank 0
against the oracle's literal feed sequence.  Launched by tests/test_gpu_multi.py."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import sigdigger_b200 as sdb
    from sigdigger_b200 import panoramic

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()

    N, n_hops, fs, rel_bw = 65536, 37, 100e6, 0.5
    fmin, fmax = 2.0e9, 2.0e9 + n_hops * fs * rel_bw
    centers = fmin + fs * rel_bw * (0.5 + np.arange(n_hops))
    rng = np.random.default_rng(11)                       # same data on every rank; each uses its shard
    x = (0.02 * (rng.standard_normal((n_hops, N)) + 1j * rng.standard_normal((n_hops, N)))).astype(np.complex64)
    t = np.arange(N)
    for h in range(n_hops):
        x[h] += (0.3 * np.exp(2j * np.pi * rng.uniform(-0.2, 0.2) * t)).astype(np.complex64)
    lo, hi = panoramic.shard(n_hops, world, rank)
    xl = torch.from_numpy(x[lo:hi]).cuda()
    res = panoramic.sweep(sdb, torch, dist, xl, centers, N, "blackmann_harris", (fmin, fmax), fs, rel_bw,
                          device=local)
    # the same sweep with the per-hop channel detector: the lists cross NVLink section by section and are read back
    # packed on rank 0; they must equal those of a single-rank sweep over all hops
    det = dict(alpha=1.0, gamma=0.5, snr=8.0, min_bins=2)
    res_det = panoramic.sweep(sdb, torch, dist, xl, centers, N, "blackmann_harris", (fmin, fmax), fs, rel_bw,
                              device=local, detect=det)
    ok = True
    if rank == 0:
        import oracle_lib as O
        L = O.lib()
        v = O.SpectrumView()
        assert L.sdo_sview_init(C.byref(v)) == 0
        L.sdo_sview_set_range(C.byref(v), fmin, fmax)
        v.fft_bandwidth = fs
        v.fft_rel_bw = rel_bw
        for h in range(n_hops):
            p = O.psd_frames(x[h], N, "blackmann_harris")[0]
            L.sdo_psd_shift_db(O.ptr(p), N)
            L.sdo_sview_feed(C.byref(v), O.ptr(p), None, N, float(centers[h]), 1)
        n = v.spectrum_size
        ref = [np.ctypeslib.as_array(q, shape=(65536,))[:n].copy() for q in (v.psd, v.psd_accum, v.psd_count)]
        psd, acc, cnt = res
        ok = (np.array_equal(cnt, ref[2]) and np.array_equal(acc.view(np.uint32), ref[1].view(np.uint32))
              and np.array_equal(psd.view(np.uint32), ref[0].view(np.uint32)))
        one = panoramic.sweep(sdb, torch, None, torch.from_numpy(x).cuda(), centers, N, "blackmann_harris", (fmin, fmax),
                              fs, rel_bw, device=local, detect=det)
        same_view = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(res_det[:3], res))
        same_lists = res_det[3] == one[3] and sum(len(c) for c in one[3]) >= n_hops
        ok = ok and same_view and same_lists
        print("panoramic multi-GPU sweep: world=%d hops=%d bins=%d exact=%s (detector lists equal: %s)"
              % (world, n_hops, n, ok, same_lists))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

"""PCIe ceilings of the box: pinned H2D / D2H alone and together, ordinary vs write-combined host memory.
python tools/pcie_probe.py   (needs a GPU)"""
import ctypes as C
import time

import torch

rt = C.CDLL("libcudart.so")
N = 320 << 20


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
    return p


def run(h_up, h_down, up=True, down=True, reps=10):
    d_up = torch.empty(N, dtype=torch.uint8, device="cuda")
    d_down = torch.ones(N, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            rt.cudaMemcpyAsync(C.c_void_p(d_up.data_ptr()), h_up, C.c_size_t(N), 1, C.c_void_p(s1.cuda_stream))
        if down:
            rt.cudaMemcpyAsync(h_down, C.c_void_p(d_down.data_ptr()), C.c_size_t(N), 2, C.c_void_p(s2.cuda_stream))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return N * reps / dt / 1e9


torch.cuda.init()
plain_up, plain_down = host_alloc(N, 0), host_alloc(N, 0)
wc_up = host_alloc(N, 4)       # cudaHostAllocWriteCombined
C.memset(plain_up, 1, N)
C.memset(wc_up, 1, N)
for name, hu in (("pinned", plain_up), ("write-combined", wc_up)):
    print("%-15s H2D alone %.1f GB/s | D2H alone %.1f | both: %.1f each way" %
          (name, run(hu, plain_down, True, False), run(hu, plain_down, False, True), run(hu, plain_down, True, True)))

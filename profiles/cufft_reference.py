import torch, time
torch.cuda.init()
for batch in (148, 1024, 4096):
    x = torch.randn(batch, 65536, dtype=torch.complex64, device='cuda')
    for _ in range(3): y = torch.fft.fft(x)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    n = 10
    for _ in range(n): y = torch.fft.fft(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print("cufft batch", batch, "ms", ms, "us/window", ms * 1e3 / batch, "GB/s r+w", batch * 65536 * 16 / ms / 1e6)
# plain copy for reference
x = torch.randn(4096, 65536, dtype=torch.complex64, device='cuda'); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): y.copy_(x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("copy 2GB ms", ms, "GB/s r+w", 2 * x.numel() * 8 / ms / 1e6)
x = torch.randn(148, 65536, dtype=torch.complex64, device='cuda'); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
a.record()
for _ in range(100): y.copy_(x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 100
print("copy 77MB ms", ms, "GB/s r+w", 2 * x.numel() * 8 / ms / 1e6)
s = x.abs().sum()
torch.cuda.synchronize()
a.record()
for _ in range(100): s = x.real.sum()
b.record(); torch.cuda.synchronize()

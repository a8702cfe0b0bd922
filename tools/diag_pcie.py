"""Scratch diagnostic: pinned D2H bandwidth (contiguous and the pitched 2-D copy sample() uses)."""
import sys, time
import torch
sys.path.insert(0, ".")
from littlemcmc_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
for mb in (64, 256, 1024):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("contiguous D2H %4d MiB: %.1f GB/s" % (mb, n / dt / 1e9))
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("contiguous H2D %4d MiB: %.1f GB/s" % (mb, n / dt / 1e9))
# pitched: 1024 rows of 32*1000*8 bytes, host pitch 400*1000*8
C_, n, D, T = 1024, 32, 1000, 400
src = torch.empty(C_, n, D, dtype=torch.float64, device=dev)
dst = torch.empty(C_, T, D, dtype=torch.float64, pin_memory=True)
s = torch.cuda.current_stream(dev).cuda_stream
def go():
    v = dst[:, :n]
    L.check(lib.lmc_memcpy2d_d2h(v.data_ptr(), v.stride(0) * 8, src.data_ptr(), src.stride(0) * 8, n * D * 8, C_, s), "cp")
go(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    go()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("pitched 2-D D2H %d MiB: %.1f GB/s" % (src.numel() * 8 >> 20, src.numel() * 8 / dt / 1e9))

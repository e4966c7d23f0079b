"""Host->device copy rate of 2 MB from different kinds of pinned memory."""
import ctypes as C
import time

import torch

rt = C.CDLL("libcudart.so.12")
n = 1920 * 1080
d = torch.empty(n, dtype=torch.uint8, device="cuda")


def bench(ptr, label):
    for _ in range(5):
        rt.cudaMemcpy(C.c_void_p(d.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 1)
    t0 = time.perf_counter()
    for _ in range(50):
        rt.cudaMemcpy(C.c_void_p(d.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 1)
    dt = (time.perf_counter() - t0) / 50
    print(f"{label:28s} {dt * 1e6:7.1f} us  {n / dt * 1e-9:6.1f} GB/s")


for flags, label in ((0, "cudaHostAlloc default"), (4, "cudaHostAlloc writeCombined"), (1, "cudaHostAlloc portable")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 7, n)
    bench(p.value, label)
t = torch.empty(n, dtype=torch.uint8).pin_memory()
bench(t.data_ptr(), "torch pin_memory()")
t2 = torch.empty(n, dtype=torch.uint8)
bench(t2.data_ptr(), "pageable")
# D2H for comparison
h = torch.empty(n, dtype=torch.uint8).pin_memory()
t0 = time.perf_counter()
for _ in range(50):
    rt.cudaMemcpy(C.c_void_p(h.data_ptr()), C.c_void_p(d.data_ptr()), C.c_size_t(n), 2)
dt = (time.perf_counter() - t0) / 50
print(f"{'D2H to pinned':28s} {dt * 1e6:7.1f} us  {n / dt * 1e-9:6.1f} GB/s")

import time, numpy as np, torch
rt = torch.cuda.cudart()
n = 1 << 30
a = np.ones(n, dtype=np.uint8)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
t = torch.from_numpy(a)
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(t); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pageable H2D 1 GiB ms", round((t1 - t0) * 1e3, 1))
t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, n, 0); t1 = time.perf_counter()
print("cudaHostRegister 1 GiB ms", round((t1 - t0) * 1e3, 1), r)
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(t, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("registered H2D 1 GiB ms", round((t1 - t0) * 1e3, 1))
h = torch.empty(n, dtype=torch.uint8)
for _ in range(2):
    t0 = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pageable D2H 1 GiB ms", round((t1 - t0) * 1e3, 1))
t0 = time.perf_counter(); r = rt.cudaHostUnregister(a.ctypes.data); t1 = time.perf_counter()
print("cudaHostUnregister ms", round((t1 - t0) * 1e3, 1), r)
b = np.empty(n, dtype=np.uint8)   # untouched pages
t0 = time.perf_counter(); r = rt.cudaHostRegister(b.ctypes.data, n, 0); t1 = time.perf_counter()
print("cudaHostRegister untouched 1 GiB ms", round((t1 - t0) * 1e3, 1), r)

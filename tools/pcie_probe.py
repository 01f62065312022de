import torch, time
n=1024*1440000
h=torch.empty(n,dtype=torch.int16,pin_memory=True); d=torch.empty(n,dtype=torch.int16,device='cuda')
for _ in range(2): d.copy_(h,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); d.copy_(h,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print('H2D pinned 2.95GB: %.1f ms = %.1f GB/s'%(dt*1e3, n*2/dt/1e9))
h2=torch.empty(842639360,dtype=torch.uint8,pin_memory=True); d2=torch.empty(842639360,dtype=torch.uint8,device='cuda')
h2.copy_(d2,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); h2.copy_(d2,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print('D2H pinned 0.84GB: %.1f ms = %.1f GB/s'%(dt*1e3, 842639360/dt/1e9))
# concurrent both directions
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
t=time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('both concurrently: %.1f ms'%(dt*1e3))

import torch, time
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device='cuda'); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("D2H pinned GB/s", n/dt/1e9)
t0=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("H2D pinned GB/s", n/dt/1e9)
# chunks of 1.17GB on 4 streams
ss=[torch.cuda.Stream() for _ in range(4)]
c = n//8
t0=time.perf_counter()
for i in range(8):
    with torch.cuda.stream(ss[i%4]): h[i*c:(i+1)*c].copy_(d[i*c:(i+1)*c], non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("D2H 8 chunks on 4 streams GB/s", n/dt/1e9)

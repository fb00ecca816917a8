import torch, time
x = torch.empty(34603008//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda:0")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(20): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt=(time.perf_counter()-t0)/20
print("H2D 34.6MB pinned: %.3f ms = %.1f GB/s" % (dt*1e3, 34.603/dt/1e3))
y = torch.empty(3932160//4, dtype=torch.float32).pin_memory(); e = torch.empty_like(y, device="cuda:0")
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): y.copy_(e, non_blocking=True)
torch.cuda.synchronize(); print("D2H 3.9MB: %.3f ms" % ((time.perf_counter()-t0)/20*1e3))

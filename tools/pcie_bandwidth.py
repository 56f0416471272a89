import torch, time
n=112_500_000
d=torch.empty(n,dtype=torch.uint8,device='cuda'); h=torch.empty(n,dtype=torch.uint8).pin_memory()
for _ in range(3): h.copy_(d,non_blocking=True); torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): h.copy_(d,non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("D2H GB/s", 10*n/ (e0.elapsed_time(e1)*1e-3)/1e9)
e0.record()
for _ in range(10): d.copy_(h,non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D GB/s", 10*n/ (e0.elapsed_time(e1)*1e-3)/1e9)

import torch
dev=torch.device("cuda:0")
n=805306368//2
a=torch.empty(n,dtype=torch.bfloat16,device=dev); b=torch.empty_like(a)
def t(f,reps=10):
    f(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
ms=t(lambda: b.copy_(a)); print(f"copy  : {ms*1e3:.0f} us  {2*a.numel()*2/ms/1e9:.2f} TB/s (read+write)")
ms=t(lambda: b.zero_()); print(f"write : {ms*1e3:.0f} us  {a.numel()*2/ms/1e9:.2f} TB/s")
ms=t(lambda: a.float().sum() if False else torch.sum(a, dtype=torch.float32)); print(f"read  : {ms*1e3:.0f} us  {a.numel()*2/ms/1e9:.2f} TB/s")
c=torch.empty(n//4,dtype=torch.bfloat16,device=dev)
ms=t(lambda: torch.add(a[:n//4], c, out=b[:n//4])); 

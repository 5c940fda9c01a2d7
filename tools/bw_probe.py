import torch
dev="cuda"
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e-3
for mb in (315, 1024, 4096):
    x=torch.empty(mb*1024*1024//4,device=dev); y=torch.empty_like(x)
    s=t(lambda: x.fill_(1.0)); print(f"fill  {mb} MB: {mb/1024/s/1000*1.0737:.2f} TB/s written")
    s=t(lambda: y.copy_(x)); print(f"copy  {mb} MB: {2*mb/1024/s/1000*1.0737:.2f} TB/s (r+w)")
    s=t(lambda: x.sum()); print(f"read  {mb} MB: {mb/1024/s/1000*1.0737:.2f} TB/s read")

import torch,time
x=torch.empty(32<<20,dtype=torch.uint8).pin_memory(); d=torch.empty_like(x,device='cuda')
y=torch.empty(16<<20,dtype=torch.uint8).pin_memory(); e=torch.empty(16<<20,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for name,fn in (("h2d32",lambda: d.copy_(x,non_blocking=True)),("d2h16",lambda: y.copy_(e,non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
    print(name, dt*1e3,'ms', (32 if 'h2d' in name else 16)/1024/dt,'GB/s')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d.copy_(x,non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print("both", dt*1e3,'ms')

import torch, time
dev=torch.device('cuda:0')
n=256<<20
h1=torch.empty(n,dtype=torch.uint8).pin_memory(); h2=torch.empty(n,dtype=torch.uint8).pin_memory()
d1=torch.empty(n,dtype=torch.uint8,device=dev); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,reps=5):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
def both(): h2d(); d2h()
print("H2D GB/s", n/t(h2d)/1e9); print("D2H GB/s", n/t(d2h)/1e9); tb=t(both); print("both: each GB/s", n/tb/1e9)

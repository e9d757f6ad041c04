import torch, time
dev=torch.device('cuda',0)
n=1<<30
h=[torch.empty(n,dtype=torch.uint8,pin_memory=True) for _ in range(3)]
d=[torch.empty(n,dtype=torch.uint8,device=dev) for _ in range(3)]
s=[torch.cuda.Stream() for _ in range(3)]
def run(label, fn, nbytes, reps=4):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/reps
    print(f"{label:50s} {nbytes/dt/1e9:6.1f} GB/s")
def d2h1():
    with torch.cuda.stream(s[0]): h[0].copy_(d[0],non_blocking=True)
def d2h2():
    with torch.cuda.stream(s[0]): h[0][:n//2].copy_(d[0][:n//2],non_blocking=True)
    with torch.cuda.stream(s[1]): h[0][n//2:].copy_(d[0][n//2:],non_blocking=True)
def h2d1():
    with torch.cuda.stream(s[2]): d[2].copy_(h[2],non_blocking=True)
def both():
    d2h1(); h2d1()
def both2():
    d2h2(); h2d1()
run("D2H 1 GiB, one stream", d2h1, n)
run("D2H 1 GiB, two streams (halves)", d2h2, n)
run("H2D 1 GiB, one stream", h2d1, n)
run("D2H + H2D concurrently (bytes of both)", both, 2*n)
run("D2H (two streams) + H2D concurrently", both2, 2*n)
def small():
    with torch.cuda.stream(s[0]):
        for k in range(8): h[0][k*(n//8):(k+1)*(n//8)].copy_(d[0][k*(n//8):(k+1)*(n//8)],non_blocking=True)
run("D2H 8 x 128 MiB, one stream", small, n)

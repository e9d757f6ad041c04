"""End-to-end timeline of one config-4 pass from pinned host memory: constructor, compute() return, and how long the
device was busy (CUDA events around the compute stream's work) -- tells host-side gaps from link-bound tails."""
import sys, time, torch, numpy as np
sys.path.insert(0, __file__.rsplit('/', 2)[0])
import bench, spectral_connectivity_b200 as sc
from spectral_connectivity_b200 import _lib
wl = bench.WORKLOADS['cfg4']
dev = torch.device('cuda', 0)
x_dev = bench.make_recording(wl, 0, dev)
x_host = torch.empty(x_dev.shape, dtype=torch.float32, pin_memory=True); x_host.copy_(x_dev); torch.cuda.synchronize()
x_np = x_host.numpy()
kw = dict(sampling_frequency=wl['fs'], time_halfbandwidth_product=wl['NW'], time_window_duration=wl['duration'])
bufs = None
if len(sys.argv) > 1:
    sc.Connectivity._GRANGER_SUB_WINDOWS = int(sys.argv[1])
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    m = sc.Multitaper(x_np, **kw); t1 = time.perf_counter()
    c = sc.Connectivity.from_multitaper(m)
    _lib.TIMER = _lib.StageTimer()
    res = c.compute(bench.MEASURES, out=bufs); t2 = time.perf_counter()
    e1.record(); torch.cuda.synchronize(); t3 = time.perf_counter()
    tot = _lib.TIMER.totals()
    evs = _lib.TIMER.events; _lib.TIMER = None
    # timeline of the compute stream relative to e0: start / end of every timed kernel
    tl = [(name, e0.elapsed_time(a), e0.elapsed_time(b_)) for name, a, b_ in evs]
    first, last = tl[0][1], tl[-1][2]
    gaps = sum(max(0.0, tl[k + 1][1] - tl[k][2]) for k in range(len(tl) - 1))
    big = sorted(((tl[k + 1][1] - tl[k][2], tl[k][0], tl[k + 1][0], k) for k in range(len(tl) - 1)), reverse=True)[:6]
    if i >= 6:
        print(f"   first kernel starts at {first:.1f} ms, last ends at {last:.1f} ms, gaps between timed kernels {gaps:.1f} ms; largest: "
              + ", ".join(f"{g:.1f} ms after {a_}->{b2}#{k}" for g, a_, b2, k in big))
    if bufs is None:
        bufs = {k: sc.pinned_empty(v.shape, v.dtype) for k, v in res.items()}
    kern = sum(v[0] for v in tot.values())
    print(f"call {i}: ctor {1e3*(t1-t0):6.1f} ms, compute() returned after {1e3*(t2-t0):6.1f} ms, compute stream done {e0.elapsed_time(e1):6.1f} ms, "
          f"sum of stage kernels {kern:6.1f} ms ({ {k: round(v[0],1) for k,v in tot.items()} })")
    del res, c, m

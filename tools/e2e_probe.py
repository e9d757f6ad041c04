import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench, spectral_connectivity_b200 as sc
wl = bench.WORKLOADS['cfg4']
dev = torch.device('cuda', 0)
x_dev = bench.make_recording(wl, 0, dev)
x_host = torch.empty(x_dev.shape, dtype=torch.float32, pin_memory=True); x_host.copy_(x_dev); torch.cuda.synchronize()
del x_dev
x_np = x_host.numpy()
kw = dict(sampling_frequency=wl['fs'], time_halfbandwidth_product=wl['NW'], time_window_duration=wl['duration'])
for i in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = sc.Multitaper(x_np, **kw); t1 = time.perf_counter()
    c = sc.Connectivity.from_multitaper(m)
    res = c.compute(bench.MEASURES); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"call {i}: ctor {1e3*(t1-t0):.1f} ms, compute {1e3*(t2-t1):.1f} ms, total {1e3*(t2-t0):.1f} ms")
    del res, c, m

"""One Granger launch on 8 windows of config 4 (for ncu): python tools/granger_profile.py [max_iterations]"""
import sys, torch
sys.path.insert(0, __file__.rsplit('/', 2)[0])
import bench, spectral_connectivity_b200 as sc
wl = dict(bench.WORKLOADS['cfg4']); wl['N'] = 8000
x = bench.make_recording(wl, 0, torch.device('cuda', 0))
kw = dict(sampling_frequency=wl['fs'], time_halfbandwidth_product=wl['NW'], time_window_duration=wl['duration'])
opts = dict(max_iterations=int(sys.argv[1])) if len(sys.argv) > 1 else {}
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(2):
        c = sc.Connectivity.from_multitaper(sc.Multitaper(x, **kw), output="torch")
        c.compute(["pairwise_spectral_granger_prediction"], **opts)
torch.cuda.synchronize()

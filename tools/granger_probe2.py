import sys, torch
sys.path.insert(0, '/root/repo')
import bench, spectral_connectivity_b200 as sc
from spectral_connectivity_b200 import _lib
wl = dict(bench.WORKLOADS['cfg4']); wl['N'] = 8000
dev = torch.device('cuda', 0)
x = bench.make_recording(wl, 0, dev)
kw = dict(sampling_frequency=wl['fs'], time_halfbandwidth_product=wl['NW'], time_window_duration=wl['duration'])
def run(**opts):
    c = sc.Connectivity.from_multitaper(sc.Multitaper(x, **kw), output="torch")
    _lib.TIMER = _lib.StageTimer()
    out = c.compute(["pairwise_spectral_granger_prediction"], **opts)
    t = _lib.TIMER.totals(); _lib.TIMER = None
    return t['granger'][0], out["pairwise_spectral_granger_prediction"]
ref = None
for name, opts in [("default", {}), ("max_iter=0", dict(max_iterations=0)), ("max_iter=4", dict(max_iterations=4)), ("max_iter=8", dict(max_iterations=8))]:
    run(**opts)
    ms, gc = run(**opts)
    print(f"{_lib.LIB_PATH.split('/')[-1]} {name:12s} granger {ms:8.2f} ms for 8 windows  (x7.5 = {ms*7.5:7.1f} ms per step) checksum {float(torch.nan_to_num(gc).double().sum()):.6f}")

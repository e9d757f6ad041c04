"""Time the Granger stage of an 8-window slice of config 4 with alternative builds of the library.

usage: python tools/variant_probe.py [lib.so ...]   (each run in a fresh process via SC_B200_LIB)
"""
import os
import subprocess
import sys

CHILD = r'''
import sys, torch
sys.path.insert(0, %(root)r)
import bench, spectral_connectivity_b200 as sc
from spectral_connectivity_b200 import _lib
wl = dict(bench.WORKLOADS['cfg4']); wl['N'] = 8000
x = bench.make_recording(wl, 0, torch.device('cuda', 0))
kw = dict(sampling_frequency=wl['fs'], time_halfbandwidth_product=wl['NW'], time_window_duration=wl['duration'])
def run(**opts):
    c = sc.Connectivity.from_multitaper(sc.Multitaper(x, **kw), output="torch")
    _lib.TIMER = _lib.StageTimer()
    c.compute(["pairwise_spectral_granger_prediction"], **opts)
    t = _lib.TIMER.totals(); _lib.TIMER = None
    return t['granger'][0]
for name, opts in [("default", {}), ("it=0", dict(max_iterations=0)), ("it=4", dict(max_iterations=4)),
                   ("no mixed", dict(mixed_precision=False))]:
    run(**opts)
    ms = min(run(**opts) for _ in range(3))
    print("  %%-10s %%8.2f ms / 8 windows" %% (name, ms))
'''

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in (sys.argv[1:] or [""]):
    env = dict(os.environ)
    if lib:
        env["SC_B200_LIB"] = os.path.abspath(lib)
    print(lib or "default build", flush=True)
    subprocess.run([sys.executable, "-c", CHILD % dict(root=root)], env=env, check=False)

import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
import spectral_connectivity_b200 as sc
g = np.load('/root/repo/tests/golden/round2.npz')
x = O.synthetic_series(300, 5, 6, 100.0, seed=9)
m = sc.Multitaper(x, sampling_frequency=100.0, time_halfbandwidth_product=3, time_window_duration=1.0)
c = sc.Connectivity.from_multitaper(m)
for rank in (1, 2, 3):
    val, vec = c.global_coherence(max_rank=rank)
    ref = g[f"global_rank{rank}_values"]
    print(rank, val.shape, ref.shape, "got", val[0, :2], "ref", ref[0, :2])
    d = np.abs(val - ref)
    w = np.unravel_index(np.argmax(d), d.shape)
    print(" worst", w, val[w[0], w[1]], ref[w[0], w[1]])

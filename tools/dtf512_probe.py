#!/usr/bin/env python
"""One window of BASELINE config 5's geometry (512 channels x 128 trials x 120 samples, 9 tapers) through
directed_transfer_function: prints iterations, flags and the wall time (used under ncu for the launch list of
the blocked Wilson path, profiles/r01_launches_dtf512_summary.txt)."""
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_connectivity_b200 as sc  # noqa: E402

rng = np.random.default_rng(55)
x = rng.standard_normal((120 * int(os.environ.get("WINDOWS", "1")), 128, 512)).astype(np.float32)
x[1:, :, 1::2] += 0.5 * x[:-1, :, 0::2]
warnings.simplefilter("ignore")
for rep in range(int(os.environ.get("REPS", "1"))):
    m = sc.Multitaper(x, sampling_frequency=2000.0, time_halfbandwidth_product=5, time_window_duration=0.060)
    c = sc.Connectivity.from_multitaper(m, output=os.environ.get("OUTPUT", "torch"))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dtf = c.directed_transfer_function()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("iterations", c.last_wilson_iterations.tolist(), "flags", c.last_wilson_flags.tolist(), f"{dt * 1e3:.1f} ms",
          "row-sum err", float(np.abs(np.asarray(dtf[:2].cpu() if hasattr(dtf, "cpu") else dtf[:2]).sum(-1) - 1).max()),
          "cached" if getattr(c, "_mvar_cache", None) is not None else "streamed", tuple(dtf.shape))

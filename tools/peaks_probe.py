"""Measure FP32 / FP64 SIMT FMA peaks with sc_simt_peak (CUDA events), print JSON."""
import ctypes, json, sys
import torch
sys.path.insert(0, '/root/repo')
from spectral_connectivity_b200 import _lib


def measure(dtype, n=1 << 14, reps=5):
    lib = _lib.load()
    scratch = torch.empty(lib.sc_simt_peak_scratch_bytes(), dtype=torch.uint8, device="cuda")
    flops = ctypes.c_double(0.0)
    best = 0.0
    for _ in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.sc_simt_peak(dtype, n, _lib.ptr(scratch), ctypes.byref(flops), _lib.stream_ptr()), "peak")
        b.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    return best


if __name__ == "__main__":
    print(json.dumps({"fp32_fma_tflops": measure(0), "fp64_fma_tflops": measure(1, n=1 << 13)}))

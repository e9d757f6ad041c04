"""One tensor-core CSM launch on a config-4 chunk (8 windows); with a -DSC_CSM_PROFILE build of csm_tc.cu the kernel
prints, for CTA 0, how many cycles every warp role spent waiting on its mbarriers (which role is the bottleneck)."""
import sys
import torch
sys.path.insert(0, __file__.rsplit('/', 2)[0])
from spectral_connectivity_b200 import _lib
lib = _lib.load()
B, F, R, S = 8, 501, 448, 256
xp = torch.randn((B, F, 2, R, S), device='cuda', dtype=torch.float32)
out = torch.empty((B, F, S, S), dtype=torch.complex64, device='cuda')
_lib.check(lib.sc_csm(_lib.ptr(xp), B, F, R, S, 1.0 / R, 0, _lib.ptr(out), _lib.stream_ptr()), 'tc')
torch.cuda.synchronize()

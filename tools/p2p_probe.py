"""2-GPU microbenchmark of the peer-memory reduce kernel (torchrun): barrier alone, kernel alone, NCCL reduce_scatter."""
import ctypes, os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectral_connectivity_b200 import _lib

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
W, F, S = 8, 501, 256
n = W * F * S * S * 8
buf = symm_mem.empty(n, dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
part = torch.view_as_complex(buf.view(torch.float32).view(W, F, S, S, 2))
part.copy_(torch.randn(W, F, S, S, dtype=torch.complex64, device=dev))
own = W // world
out = torch.empty((own, F, S, S), dtype=torch.complex64, device=dev)
pw = torch.empty((own, F, S), dtype=torch.float32, device=dev)
coh = torch.empty((own, F, S, S), dtype=torch.float32, device=dev)
ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

t_bar = timeit(lambda: hdl.barrier(channel=0))
def kern(measure=1):
    _lib.check(lib.sc_peer_reduce_csm(ptrs, world, rank * own * F, own * F, S, _lib.ptr(out), _lib.ptr(pw), measure,
                                      _lib.ptr(coh), _lib.stream_ptr()), "peer")
t_k = timeit(kern)
t_k0 = timeit(lambda: kern(-1))
def both():
    hdl.barrier(channel=0); kern()
t_both = timeit(both)
outn = torch.empty((own, F, S, S), dtype=torch.complex64, device=dev)
t_nccl = timeit(lambda: dist.reduce_scatter_tensor(torch.view_as_real(outn), torch.view_as_real(part)))
ref = outn.clone(); kern(); torch.cuda.synchronize()
err = float((out - ref).abs().max())
gb = own * F * S * S * 8 / 1e9
print(f"[rank {rank}] chunk {W} windows, own {own} ({gb:.2f} GB per peer): barrier {t_bar:.3f} ms, kernel+coh {t_k:.2f} ms "
      f"({gb * world / t_k * 1e3:.0f} GB/s read), kernel plain {t_k0:.2f} ms, barrier+kernel {t_both:.2f} ms, "
      f"NCCL reduce_scatter {t_nccl:.2f} ms, max |p2p - nccl| {err:.2e}", flush=True)
dist.destroy_process_group()

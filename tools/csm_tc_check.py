import torch, numpy as np, sys, time
sys.path.insert(0,'/root/repo')
from spectral_connectivity_b200 import _lib
lib=_lib.load()
for n_obs in (64,128,256,448,896):
    n_bf,n_sig=2,256
    g=torch.Generator(device='cuda').manual_seed(1)
    xp=torch.randn((n_bf,1,2,n_obs,n_sig),generator=g,device='cuda',dtype=torch.float32)
    out=torch.empty((n_bf,1,n_sig,n_sig),dtype=torch.complex64,device='cuda'); out2=torch.empty_like(out)
    _lib.check(lib.sc_csm(_lib.ptr(xp),n_bf,1,n_obs,n_sig,1.0,0,_lib.ptr(out),_lib.stream_ptr()),'tc')
    _lib.check(lib.sc_csm_simt(_lib.ptr(xp),n_bf,1,n_obs,n_sig,1.0,0,_lib.ptr(out2),_lib.stream_ptr()),'simt')
    torch.cuda.synchronize()
    z=torch.complex(xp[:,0,0].double(),xp[:,0,1].double()); ref=torch.einsum('bri,brj->bij',z,z.conj()).cpu().numpy()
    got=out[:,0].cpu().numpy(); simt=out2[:,0].cpu().numpy()
    d=np.einsum('bii->bi',got).real; dr=np.einsum('bii->bi',ref).real; ds=np.einsum('bii->bi',simt).real
    off=~np.eye(n_sig,dtype=bool)
    print(n_obs,'diag signed rel bias tc %.3e simt %.3e'%(((d-dr)/dr).mean(), ((ds-dr)/dr).mean()),
          'offdiag rms err/sqrt(PiPj) tc %.2e simt %.2e'%(np.sqrt((np.abs(got-ref)[:,off]**2).mean())/dr.mean(), np.sqrt((np.abs(simt-ref)[:,off]**2).mean())/dr.mean()))
# speed at cfg4 chunk size: 8 windows
B,F,R,S=8,501,448,256
xp=torch.randn((B,F,2,R,S),device='cuda',dtype=torch.float32)
out=torch.empty((B,F,S,S),dtype=torch.complex64,device='cuda')
for name,fn in (('tc',lib.sc_csm),('simt',lib.sc_csm_simt)):
    for _ in range(2): _lib.check(fn(_lib.ptr(xp),B,F,R,S,1.0/R,0,_lib.ptr(out),_lib.stream_ptr()),name)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): _lib.check(fn(_lib.ptr(xp),B,F,R,S,1.0/R,0,_lib.ptr(out),_lib.stream_ptr()),name)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/3
    print(name,'ms per 8-window chunk',ms,'-> per 60 windows',ms*7.5, 'TFLOP/s (8*S^2*R real flop)',8*S*S*R*B*F/ms/1e9)

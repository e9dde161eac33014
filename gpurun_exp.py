import sys, time, torch, numpy as np, os
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps*1e3
class KF:
    def __init__(s,c): s.c=torch.tensor(c)
    def get_coeffs(s): return s.c
    def get_deriv_coeffs(s): return s.c
for name,(N,d,L) in {'A':(1_000_000,8,16),'B':(16600,18,11),'A1':(1_000_000,8,1)}.items():
    torch.manual_seed(0)
    x=torch.randn(N,d,device='cuda'); v=torch.randn(N,L,device='cuda'); go=torch.randn(N,L,device='cuda')
    kf=KF([0.34608543,1,0.34608543])
    xr=x.clone().requires_grad_(True)
    def fwd():
        with torch.no_grad(): return sg.LatticeFilterGeneral.apply(v,xr,kf)
    def fwdbwd():
        xr.grad=None
        out=sg.LatticeFilterGeneral.apply(v,xr,kf); out.backward(go)
    for ch in (None,1,2,4,8):
        sg.LatticeFilterGeneral.grad_chunk=ch
        print(name,'chunk',ch,'fwd ms',round(wall(fwd),3),'fwd+bwd ms',round(wall(fwdbwd),3), 'cache builds',sg.lattice_cache.builds,'hits',sg.lattice_cache.hits)

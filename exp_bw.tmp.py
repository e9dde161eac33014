import sys, torch, ctypes as C
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
from simplex_gp_b200 import _capi
from simplex_gp_b200.lattice import _ptr, _stream_ptr, _fp
from simplex_gp_b200.function import lattice_filter_grad
torch.manual_seed(0)
N,d,L=1_000_000,8,16
x=torch.randn(N,d,device='cuda'); g=torch.randn(N,L,device='cuda'); v=torch.randn(N,L,device='cuda')
c=[0.34608543,1,0.34608543]
lat=sg.Lattice(x,c)
def timeit(fn, reps=10, warm=2):
    for i in range(warm): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
print('backward total ms', timeit(lambda: lattice_filter_grad(lat,g,v,x,c,True)))
lib=_capi.lib(); st=_stream_ptr(lat.device)
for nl in (8,4,2):
    W=18*nl
    packed=torch.zeros(N,W,device='cuda'); filt=torch.empty(N,W,device='cuda'); gx=torch.empty(N,d,device='cuda')
    t_pack=timeit(lambda: _capi.check(lib.sgp_grad_pack(_ptr(g),L,_ptr(v),L,_ptr(x),d,N,d,0,nl,_ptr(packed),W,st)))
    t_splat=timeit(lambda: lat.splat(packed,mode=4))
    sp=lat.splat(packed,mode=4)
    t_blur=timeit(lambda: lat.blur(sp,groups=True,exact=False))
    bl=lat.blur(sp,groups=True,exact=False)
    t_mvm=timeit(lambda: lat.mvm(packed,out=filt))
    t_con=timeit(lambda: _capi.check(lib.sgp_grad_contract(_ptr(filt),W,_ptr(g),L,_ptr(v),L,_ptr(x),d,N,d,0,nl,1,1,_ptr(gx),d,None,0,st)))
    print(f'nl={nl} ({W} ch): pack {t_pack:.2f} splat {t_splat:.2f} blur {t_blur:.2f} mvm {t_mvm:.2f} (slice ~{t_mvm-t_splat-t_blur:.2f}) contract {t_con:.2f} ms; per column {(t_pack+t_mvm+t_con)/nl:.3f}')
    del packed,filt,sp,bl

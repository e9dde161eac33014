import sys, time, torch, numpy as np, ctypes as C, os
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
from simplex_gp_b200 import _capi
from simplex_gp_b200.lattice import _fp, _ptr, _stream_ptr
torch.manual_seed(0)
N,d,L=1_000_000,8,16
x=torch.randn(N,d,device='cuda'); v=torch.randn(N,L,device='cuda')
c=[0.34608543,1,0.34608543]
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1000
lat=sg.Lattice(x,c); torch.cuda.synchronize()
t0=time.time(); lat=sg.Lattice(x,c); torch.cuda.synchronize(); bt=(time.time()-t0)*1e3
lib=_capi.lib(); buf0,buf1=lat._scratch(L); st=_stream_ptr(lat.device); where=C.c_int(0); arr=lat.groups['array']; cnp=lat.coeffs
out=torch.empty(N,L,device='cuda')
r=lat.rows
t_sr=timeit(lambda: _capi.check(lib.sgp_splat_rows(_ptr(r['ent']),_ptr(r['ent_row']),N,d,lat.M,_ptr(v),v.stride(0),L,_ptr(buf0),st)))
vi=lat._view(lat._table(False,False),None,False,True)
t_sa=timeit(lambda: _capi.check(lib.sgp_splat(C.byref(vi),_ptr(v),v.stride(0),L,_ptr(buf0),1,st)))
t_bg=timeit(lambda: _capi.check(lib.sgp_blur_groups(arr,len(arr),lat.M,1,_fp(cnp),3,L,_ptr(buf0),_ptr(buf1),C.byref(where),1,st)))
vo=lat._view(lat._table(False,True),None,False,True)
t_sl=timeit(lambda: _capi.check(lib.sgp_slice(C.byref(vo),_ptr(buf1),L,_ptr(out),out.stride(0),st)))
vo2=lat._view(exact=False)
t_sl2=timeit(lambda: _capi.check(lib.sgp_slice(C.byref(vo2),_ptr(buf1),L,_ptr(out),out.stride(0),st)))
t_m=timeit(lambda: lat.mvm(v,out=out))
print(f'build {bt:.1f} ms: splat rows {t_sr:.1f} atomic {t_sa:.1f} | blur groups {t_bg:.1f} | slice transposed {t_sl:.1f} plain {t_sl2:.1f} | mvm {t_m:.1f} us -> {1e6/t_m:.0f} MVM/s')

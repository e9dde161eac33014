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
for ga, gr in ((3,512),(3,384),(2,512),(4,1500),(5,1600),(1,512)):
    torch.cuda.synchronize(); t0=time.time()
    lat=sg.Lattice(x,c,group_axes=ga,group_rows=gr); torch.cuda.synchronize(); bt=time.time()-t0
    if lat.groups is None: print('no groups', ga, gr); continue
    gl=lat.groups['list']
    lib=_capi.lib(); buf0,buf1=lat._scratch(L); st=_stream_ptr(lat.device); where=C.c_int(0); arr=lat.groups['array']; cnp=lat.coeffs
    view=lat._view()
    t_g=timeit(lambda: _capi.check(lib.sgp_blur_groups(arr,len(arr),lat.M,1,_fp(cnp),3,L,_ptr(buf0),_ptr(buf1),C.byref(where),st)))
    t_a=timeit(lambda: _capi.check(lib.sgp_blur(C.byref(view),_fp(cnp),3,L,_ptr(buf0),_ptr(buf1),C.byref(where),st)))
    out=torch.empty(N,L,device='cuda')
    t_mg=timeit(lambda: lat.mvm(v,out=out,mode=1,blur='groups'))
    t_ma=timeit(lambda: lat.mvm(v,out=out,mode=1,blur='axis'))
    print(f'axes {ga} rows {gr}: build {bt*1e3:.1f} ms groups', [(g['j0'],g['j1'],g['max_class'],g['rows_cap'],g['n_batches']) for g in gl],
          f'blur groups {t_g:.1f} us, axis {t_a:.1f} us; mvm groups {t_mg:.1f} us, axis {t_ma:.1f} us')

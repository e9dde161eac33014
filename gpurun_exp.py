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
cfgs=[(None if a.split(',')[0]=='a' else int(a.split(',')[0]), int(a.split(',')[1])) for a in sys.argv[1:]] or [(None,512)]
for ga, gr in cfgs:
    lat=sg.Lattice(x,c,group_axes=ga,group_rows=gr); torch.cuda.synchronize()
    lib=_capi.lib(); buf0,buf1=lat._scratch(L); st=_stream_ptr(lat.device); where=C.c_int(0); arr=lat.groups['array']; cnp=lat.coeffs
    out=torch.empty(N,L,device='cuda')
    t_bg=timeit(lambda: _capi.check(lib.sgp_blur_groups(arr,len(arr),lat.M,1,_fp(cnp),3,L,_ptr(buf0),_ptr(buf1),C.byref(where),1,st)))
    t_m=timeit(lambda: lat.mvm(v,out=out))
    print(f'T={os.environ.get("SGP_GROUP_THREADS")} axes {ga} rows {gr} groups {[(g["j0"],g["j1"],g["n_batches"],g["rows_cap"]) for g in lat.groups["list"]]}: blur {t_bg:.1f} | mvm {t_m:.1f} us -> {1e6/t_m:.0f} MVM/s')

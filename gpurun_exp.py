import sys, time, torch, numpy as np, ctypes as C, os, shutil
which=sys.argv[1]
shutil.copy(f'/root/repo/lib_v{which}.tmp.so','/root/repo/simplex-gp_b200/libsgp_lattice.so')
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
from simplex_gp_b200 import _capi
from simplex_gp_b200.lattice import _fp, _ptr, _stream_ptr
torch.manual_seed(0)
N,d,L=1_000_000,8,16
x=torch.randn(N,d,device='cuda'); vs=[torch.randn(N,L,device='cuda') for _ in range(4)]
c=[0.34608543,1,0.34608543]
def timeit(fn, reps=100, warm=10):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1000
outs=[torch.empty(N,L,device='cuda') for _ in range(4)]
lat=sg.Lattice(x,c); torch.cuda.synchronize()
lib=_capi.lib(); buf0,buf1=lat._scratch(L); st=_stream_ptr(lat.device)
r=lat.rows
vo=lat._view(lat._table(False,True),None,False)
t_sr=timeit(lambda i: _capi.check(lib.sgp_splat_rows(_ptr(r['ent']),_ptr(r['ent_row']),N,d,lat.M,_ptr(vs[i%4]),L,L,_ptr(buf0),st)))
t_sl=timeit(lambda i: _capi.check(lib.sgp_slice(C.byref(vo),_ptr(buf1),L,_ptr(outs[i%4]),L,st)))
graphs=[lat.capture(vs[k],outs[k]) for k in range(4)]
t_g=timeit(lambda i: graphs[i%4].replay())
print(f'variant {which}: splat {t_sr:.1f} slice {t_sl:.1f} graph {t_g:.1f} us/mvm -> {1e6/t_g:.0f} MVM/s')

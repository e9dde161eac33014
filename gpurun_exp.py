import sys, time, torch, numpy as np, ctypes as C, os
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
from simplex_gp_b200 import _capi
from simplex_gp_b200.lattice import _fp, _ptr, _stream_ptr
def timeit(fn, reps=10, warm=2):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1000
CFG={'B':(16600,18,11,[0.34608543,1,0.34608543]),'C':(2_050_000,11,16,[0.15233751,0.50067621,1.0,0.50067621,0.15233751]),'D10':(1_000_000,24,4,[0.08435782,0.24239115,0.60311586,1.0,0.60311586,0.24239115,0.08435782])}
for name in sys.argv[1].split(','):
    N,d,L,c=CFG[name]
    torch.manual_seed(0)
    x=torch.randn(N,d,device='cuda'); v=torch.randn(N,L,device='cuda'); out=torch.empty(N,L,device='cuda')
    for ga,gr in [(None,512),(None,768),(None,1024),(2,512)]:
        lat=sg.Lattice(x,c,group_axes=ga,group_rows=gr); torch.cuda.synchronize()
        if lat.groups is None: print(name,ga,gr,'no groups'); continue
        lib=_capi.lib(); buf0,buf1=lat._scratch(L); st=_stream_ptr(lat.device); where=C.c_int(0); arr=lat.groups['array']; cnp=lat.coeffs
        t_bg=timeit(lambda i: _capi.check(lib.sgp_blur_groups(arr,len(arr),lat.M,lat.order,_fp(cnp),len(c),L,_ptr(buf0),_ptr(buf1),C.byref(where),1,st)))
        t_m=timeit(lambda i: lat.mvm(v,out=out))
        print(name,'axes',ga,'rows',gr,'groups',[(g['j0'],g['j1'],g['max_class'],g['n_batches']) for g in lat.groups['list']],'blur us',round(t_bg,1),'mvm us',round(t_m,1))
        del lat

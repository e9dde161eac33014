import sys, time, torch, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
from simplex_gp_b200 import _capi
from simplex_gp_b200.lattice import _fp, _ptr, _stream_ptr
torch.manual_seed(0)
N,d,L=1_000_000,8,16
x=torch.randn(N,d,device='cuda'); v=torch.randn(N,L,device='cuda')
c=[0.34608543,1,0.34608543]
def stages(lat, v, mode, reps=20):
    lib=_capi.lib(); view=lat._view(); buf0,buf1=lat._scratch(L); cnp=lat.coeffs; st=_stream_ptr(lat.device)
    out=torch.empty(N,L,device='cuda'); where=C.c_int(0)
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(4)]; acc=[0,0,0]
    for i in range(reps+3):
        ev[0].record(); _capi.check(lib.sgp_splat(C.byref(view), _ptr(v), v.stride(0), L, _ptr(buf0), mode, st))
        ev[1].record(); _capi.check(lib.sgp_blur(C.byref(view), _fp(cnp), 3, L, _ptr(buf0), _ptr(buf1), C.byref(where), st))
        ev[2].record(); _capi.check(lib.sgp_slice(C.byref(view), _ptr(buf1 if where.value else buf0), L, _ptr(out), out.stride(0), st))
        ev[3].record(); torch.cuda.synchronize()
        if i>=3:
            for k in range(3): acc[k]+=ev[k].elapsed_time(ev[k+1])
    return [a/reps*1000 for a in acc]
def stages_tiles(lat, v, reps=20):
    lib=_capi.lib(); view=lat._view(); tv=lat._tiles_view(); buf0,buf1=lat._scratch(L); cnp=lat.coeffs; st=_stream_ptr(lat.device)
    out=torch.empty(N,L,device='cuda'); where=C.c_int(0)
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(4)]; acc=[0,0,0]
    for i in range(reps+3):
        ev[0].record(); _capi.check(lib.sgp_splat_tiles(C.byref(tv), _ptr(v), v.stride(0), L, _ptr(buf0), st))
        ev[1].record(); _capi.check(lib.sgp_blur(C.byref(view), _fp(cnp), 3, L, _ptr(buf0), _ptr(buf1), C.byref(where), st))
        ev[2].record(); _capi.check(lib.sgp_slice_tiles(C.byref(tv), _ptr(buf1 if where.value else buf0), L, _ptr(out), out.stride(0), st))
        ev[3].record(); torch.cuda.synchronize()
        if i>=3:
            for k in range(3): acc[k]+=ev[k].elapsed_time(ev[k+1])
    return [a/reps*1000 for a in acc]
import time
for T in (128,256,512):
    torch.cuda.synchronize(); t0=time.time(); lat=sg.Lattice(x,c,tile_points=T); torch.cuda.synchronize(); bt=time.time()-t0
    print('T',T,'build s',bt,'S',lat.tiles['S'],'S/pv',lat.tiles['S']/(N*(d+1)),'max_dict',lat.tiles['max_dict'],'tiles: splat/blur/slice us', stages_tiles(lat,v))
print('atomic', stages(lat,v,1))
sys.exit(0)
# sort points lexicographically by greedy
g=lat.greedy.cpu().numpy().astype(np.int32)
order=torch.from_numpy(np.lexsort(g[:, ::-1].T)).cuda()
xs=x[order].contiguous(); vs=v[order].contiguous()
lat2=sg.Lattice(xs,c)
print('sorted pts: splat/blur/slice us atomic', stages(lat2,vs,1), 'gather', stages(lat2,vs,2))
cnt=torch.bincount(lat.offsets.reshape(-1).long(), minlength=lat.M)
print('max row', cnt.max().item())

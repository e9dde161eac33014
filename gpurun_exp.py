import sys, time, torch, numpy as np, ctypes as C, os
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
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
lat=sg.Lattice(x,c); torch.cuda.synchronize()
outs=[torch.empty(N,L,device='cuda') for _ in range(4)]
ref=lat.mvm(vs[0], mode=1, blur='axis', exact=True).clone()
t_e=timeit(lambda i: lat.mvm(vs[i%4],out=outs[i%4]))
graphs=[lat.capture(vs[k],outs[k]) for k in range(4)]
t_g=timeit(lambda i: graphs[i%4].replay())
torch.cuda.synchronize()
print(f'PDL={os.environ.get("SGP_PDL")}: eager {t_e:.1f} graph {t_g:.1f} us/mvm -> {1e6/t_g:.0f} MVM/s; rel err', float((outs[0]-ref).norm()/ref.norm()))

import sys, time, torch, numpy as np, ctypes as C, os
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
torch.manual_seed(0)
N,d,L=1_000_000,8,16
x=torch.randn(N,d,device='cuda'); vs=[torch.randn(N,L,device='cuda') for _ in range(4)]
c=[0.34608543,1,0.34608543]
def timeit(fn, reps=50, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1000
lat=sg.Lattice(x,c); torch.cuda.synchronize()
outs=[torch.empty(N,L,device='cuda') for _ in range(4)]
print('eager us/mvm', timeit(lambda i: lat.mvm(vs[i%4],out=outs[i%4])))
g=torch.cuda.CUDAGraph()
s=torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(4): lat.mvm(vs[i],out=outs[i])
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    for i in range(4): lat.mvm(vs[i],out=outs[i])
print('graph (4 mvm) us/mvm', timeit(lambda i: g.replay(), reps=20)/4)
ref=outs[0].clone(); lat.mvm(vs[0],out=outs[0]); torch.cuda.synchronize(); print('rel', float((ref-outs[0]).norm()/ref.norm()))

import sys, time, torch, numpy as np, os
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
torch.manual_seed(0)
N,d,L=1_000_000,8,16
x=torch.randn(N,d).pin_memory(); v=torch.randn(N,L).pin_memory()
c=torch.tensor([0.34608543,1,0.34608543])
ts=[]
r=None
for i in range(12):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=sg.filter(v,x,c); ts.append((time.perf_counter()-t0)*1e3)
print('filter per-call ms', [round(t,2) for t in ts])
ts=[]
for i in range(8):
    t0=time.perf_counter(); o=torch.empty((N,L),pin_memory=True); ts.append((time.perf_counter()-t0)*1e3)
print('alloc pinned (rebinding o) ms', [round(t,2) for t in ts])
ts=[]
for i in range(8):
    t0=time.perf_counter(); o=None; o=torch.empty((N,L),pin_memory=True); ts.append((time.perf_counter()-t0)*1e3)
print('alloc pinned (free first) ms', [round(t,2) for t in ts])

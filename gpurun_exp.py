import sys, time, torch
sys.path.insert(0, '/root/repo')
import simplex_gp_b200 as sg
torch.manual_seed(0)
N,d=1_000_000,8
x=torch.randn(N,d,device='cuda')
c=[0.34608543,1,0.34608543]
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps*1e3
print('build plain ms', wall(lambda: sg.Lattice(x,c,build_groups=False,build_rows=False)))
print('build +groups ms', wall(lambda: sg.Lattice(x,c,build_rows=False)))
print('build +groups(3 fixed) ms', wall(lambda: sg.Lattice(x,c,build_rows=False,group_axes=3)))
print('build full ms', wall(lambda: sg.Lattice(x,c)))

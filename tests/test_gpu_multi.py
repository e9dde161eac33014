"""Multi-GPU paths on real devices (NCCL): lattice broadcast, RHS-column sharding and point sharding against the
single-GPU result.  Needs >= 2 visible GPUs (`gpurun --gpus 2`); with fewer the ranks-as-one-process variants still run
(world size 1 exercises the same code without the collectives)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import MAT15_2, RBF1, ROOT, make_inputs

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import simplex_gp_b200 as sg
        from simplex_gp_b200.distributed import (ColumnShardedOperator, PointShardedLattice, broadcast_lattice,
                                                 shard_points)
        N, d, L = 40_000, 6, 12
        x, v = make_inputs(N, d, L, seed=123)
        xd, vd = x.to(dev), v.to(dev)
        ref_lat = sg.Lattice(xd, MAT15_2)
        want = ref_lat.mvm(vd, mode=1, blur="axis", exact=True).clone()
        # 1. lattice built on rank 0 and broadcast; every rank filters its own columns
        lat = broadcast_lattice(ref_lat if rank == 0 else None, src=0, device=dev)
        assert lat.M == ref_lat.M and torch.equal(lat.keys, ref_lat.keys) and torch.equal(lat.nbr, ref_lat.nbr)
        op = ColumnShardedOperator(lat)
        full = op.matmul_full(vd)
        e_col = float((full - want).norm() / want.norm())
        dots = op.dots(vd[:, slice(*op.columns(L))], full[:, slice(*op.columns(L))], L)
        e_dot = float((dots - (vd * want).sum(0)).abs().max() / (vd * want).sum(0).abs().max())
        # 2. point sharding: local splat, all-reduce of the lattice values, blur, local slice
        ps = PointShardedLattice(xd, MAT15_2)
        lo, hi = shard_points(N, world, rank)
        mine = ps.mvm(vd[lo:hi].contiguous(), column_blur=False)      # all-reduce, replicated blur
        e_pt = float((mine - want[lo:hi]).norm() / want[lo:hi].norm())
        if world > 1:                                                   # reduce-scatter, column-sharded blur, all-gather
            mine2 = ps.mvm(vd[lo:hi].contiguous(), column_blur=True)
            e_pt = max(e_pt, float((mine2 - want[lo:hi]).norm() / want[lo:hi].norm()))
            assert ps.M == ref_lat.M and torch.equal(ps.local.keys, ref_lat.keys)
        # 3. a CG training step with the probe / RHS columns sharded: value and lengthscale gradient equal the
        #    single-GPU ones (gradients summed over ranks)
        from simplex_gp_b200 import gp
        n2 = 4000
        x2, v2 = make_inputs(n2, 4, 1, seed=77)
        y2 = (torch.sin(x2.sum(1)) + 0.1 * v2[:, 0]).to(dev)
        probes = torch.randn(n2, 5, generator=torch.Generator().manual_seed(5)).sign().to(dev)
        res = []
        for sharded in (False, True):
            torch.manual_seed(0)
            k = sg.RBFLattice(ard_num_dims=4, order=1).to(dev)
            s_, nz, mu = torch.tensor(0.9, device=dev), torch.tensor(0.2, device=dev), torch.tensor(0.0, device=dev)
            opr = k(x2.to(dev))
            if sharded:
                val, sur = gp.mll_cg_sharded(opr.matmul, y2, mu, s_, nz, probes, tol=1e-5)
            else:
                val, sur = gp.mll_cg(opr.matmul, y2, mu, s_, nz, probes=probes, tol=1e-5)
            sur.backward()
            g = k.raw_lengthscale.grad.clone()
            if sharded and world > 1:
                dist.all_reduce(g)
            res.append((val, g))
        e_val = abs(res[0][0] - res[1][0]) / abs(res[0][0])
        e_grad = float((res[0][1] - res[1][1]).norm() / res[0][1].norm())
        assert e_val < 1e-5 and e_grad < 1e-3, (e_val, e_grad)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", (e_col, e_dot, e_pt)))
    except Exception as exc:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc() + repr(exc)))


@pytest.mark.parametrize("world", [1, 2])
def test_column_and_point_sharding_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r
        e_col, e_dot, e_pt = r[2]
        assert e_col < 1e-5 and e_dot < 1e-4 and e_pt < 1e-5, r

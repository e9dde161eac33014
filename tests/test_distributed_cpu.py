"""Host-side logic of the multi-GPU path on CPU: gloo backend, world_size 2 (one process per rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r
    return dict((r[0], r[2]) for r in res)


def _entry(fn, rank, world, port, q, *args):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", out))
    except Exception as exc:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc() + repr(exc)))


def test_shard_ranges():
    from simplex_gp_b200.distributed import shard_columns, shard_points
    for L in (0, 1, 7, 16, 17):
        for world in (1, 2, 3, 8):
            ranges = [shard_columns(L, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == L
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert shard_points(10, 2, 1) == (5, 10)
    with pytest.raises(ValueError):
        shard_columns(4, 2, 2)


def _bcast_worker(rank, world):
    from simplex_gp_b200.distributed import broadcast_lattice_arrays
    N, d, M, r = 50, 3, 17, 2
    if rank == 0:
        g = torch.Generator().manual_seed(3)
        arrays = {"replay": torch.randint(0, 100, (N, d + 1, 2), generator=g, dtype=torch.int32),
                  "keys": torch.randint(-50, 50, (M, d), generator=g, dtype=torch.int16),
                  "nbr": torch.randint(-1, M, (d + 1, M, 2 * r), generator=g, dtype=torch.int32)}
        meta = {"N": N, "M": M, "d": d, "order": r, "coeffs": [0.1, 0.5, 1.0, 0.5, 0.1]}
    else:
        arrays, meta = None, None
    arrays, meta = broadcast_lattice_arrays(arrays, meta, src=0, device="cpu")
    return {k: v.numpy().copy() for k, v in arrays.items()}, meta


def test_broadcast_lattice_arrays_gloo():
    res = _run(_bcast_worker, 2)
    (a0, m0), (a1, m1) = res[0], res[1]
    assert m0 == m1 and m1["M"] == 17
    for k in a0:
        assert np.array_equal(a0[k], a1[k]) and a0[k].size > 0


def _colshard_worker(rank, world):
    """Column sharding: every rank filters its own RHS columns of the same lattice; gathered == unsharded.
    The MVM itself is the oracle here (CPU): the test is about the sharding arithmetic and the collectives."""
    from oracle import oracle
    from simplex_gp_b200.distributed import all_gather_columns, shard_columns
    from conftest import RBF1, make_inputs
    x, v = make_inputs(400, 3, 5, seed=9)
    O = oracle.OracleLattice(x.numpy(), RBF1)
    lo, hi = shard_columns(v.shape[1], world, rank)
    mine = torch.from_numpy(O.mvm(v[:, lo:hi].contiguous().numpy())) if hi > lo else torch.empty(400, 0)
    full = all_gather_columns(mine, v.shape[1])
    want = torch.from_numpy(O.mvm(v.numpy()))
    return bool(torch.equal(full, want))


def test_column_sharding_gloo():
    res = _run(_colshard_worker, 2)
    assert res[0] and res[1]


def _pointshard_worker(rank, world):
    """Point sharding: each rank splats its own points into the full lattice, values are all-reduced before the
    blur, each rank slices its own points.  Structure comes from the oracle (global numbering)."""
    from oracle import oracle
    from simplex_gp_b200.distributed import allreduce_lattice_values, shard_points
    from conftest import RBF1, make_inputs
    x, v = make_inputs(600, 4, 3, seed=10)
    O = oracle.OracleLattice(x.numpy(), RBF1)
    lo, hi = shard_points(600, world, rank)
    off, w = O.offsets, O.weights
    vals = np.zeros((O.M, 3), dtype=np.float64)
    for n in range(lo, hi):
        for r_ in range(O.d + 1):
            vals[off[n, r_]] += np.float64(w[n, r_]) * v[n].numpy().astype(np.float64)
    vals_t = torch.from_numpy(vals)
    allreduce_lattice_values(vals_t)
    _, sp, _ = O.mvm(v.numpy(), return_intermediates=True)
    err = float(np.abs(vals_t.numpy() - sp.astype(np.float64)).max() / np.abs(sp).max())
    return err


def test_point_sharded_splat_allreduce_gloo():
    res = _run(_pointshard_worker, 2)
    assert res[0] < 1e-6 and res[1] < 1e-6


def _keylists_worker(rank, world):
    """The ragged all-gather of the ranks' key lists (point-sharded build): same lists, in rank order, on every rank."""
    from simplex_gp_b200.distributed import gather_key_lists
    g = torch.Generator().manual_seed(100 + rank)
    mine = torch.randint(-300, 300, (5 + 7 * rank, 4), generator=g, dtype=torch.int16)
    lists = gather_key_lists(mine)
    return [t.numpy().copy() for t in lists]


def test_gather_key_lists_gloo():
    res = _run(_keylists_worker, 2)
    assert [a.shape for a in res[0]] == [(5, 4), (12, 4)]
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    g = torch.Generator().manual_seed(101)
    assert np.array_equal(res[0][1], torch.randint(-300, 300, (12, 4), generator=g, dtype=torch.int16).numpy())


def _bcast_lattice_meta_worker(rank, world):
    from simplex_gp_b200.distributed import broadcast_lattice_arrays
    if rank == 0:
        arrays = {"replay": torch.zeros((0, 3, 2), dtype=torch.int32), "keys": torch.zeros((0, 2), dtype=torch.int16),
                  "nbr": torch.zeros((3, 0, 2), dtype=torch.int32)}
        meta = {"N": 0, "M": 0, "d": 2, "order": 1, "coeffs": [0.34608543, 1.0, 0.34608543], "exact": True}
    else:
        arrays, meta = None, None
    arrays, meta = broadcast_lattice_arrays(arrays, meta, src=0, device="cpu")
    return meta, {k: tuple(v.shape) for k, v in arrays.items()}


def test_broadcast_header_carries_exact_and_stencil_bits():
    res = _run(_bcast_lattice_meta_worker, 2)
    for meta, shapes in (res[0], res[1]):
        assert meta["exact"] is True and meta["order"] == 1
        assert np.array_equal(np.asarray(meta["coeffs"], np.float32).view(np.int32),
                              np.asarray([0.34608543, 1.0, 0.34608543], np.float32).view(np.int32))
        assert shapes == {"replay": (0, 3, 2), "keys": (0, 2), "nbr": (3, 0, 2)}

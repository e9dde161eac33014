"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars: lattice structure (greedy, rank, keys, vertex offsets, neighbour indices) bit-exact; barycentric weights
bit-exact; values bit-exact on the deterministic path (gather splat, blur, slice) and within 1e-5 relative on the
atomic splat path (fp32 atomics reorder the sums).
"""
import numpy as np
import pytest
import torch

from conftest import MAT15_2, MAT15_3, RBF1, RBF2, bits, make_inputs

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star tolerance for fp32 values

CASES = [
    # N, d, L, coeffs, dist
    (4, 2, 1, [0.5, 1.0, 0.5], "randn"),
    (200, 1, 3, RBF1, "randn"),
    (1000, 5, 2, [1.0], "randn"),
    (5000, 3, 4, RBF1, "randn"),
    (20000, 8, 16, RBF1, "randn"),
    (20000, 8, 2, RBF1, "rand"),
    (3000, 18, 11, RBF1, "randn"),
    (20000, 11, 3, MAT15_2, "randn"),
    (2000, 24, 2, MAT15_3, "randn"),
    (777, 7, 5, RBF2, "randn"),
    (300, 30, 1, RBF1, "randn"),   # generic-d kernels (d > 24)
]


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


@pytest.mark.parametrize("N,d,L,coeffs,dist", CASES)
def test_structure_and_values_match_oracle(sg, oracle, N, d, L, coeffs, dist):
    x, v = make_inputs(N, d, L, seed=N + d, dist=dist)
    O = oracle.OracleLattice(x.numpy(), coeffs)
    lat = sg.Lattice(x.cuda(), coeffs, build_csr=True, build_tiles=True, sort_points=True)
    assert lat.M == O.M
    assert np.array_equal(lat.scale.view(np.int32), O.scale.view(np.int32))
    assert np.array_equal(lat.greedy.cpu().numpy(), O.greedy)
    assert np.array_equal(lat.rank.cpu().numpy(), O.rank)
    assert np.array_equal(bits(lat.weights.cpu().numpy()), bits(O.weights))
    assert np.array_equal(lat.offsets.cpu().numpy(), O.offsets)
    assert np.array_equal(lat.keys.cpu().numpy(), O.keys)
    if lat.order > 0:
        assert np.array_equal(lat.nbr.cpu().numpy(), O.nbr)

    out_o, sp_o, bl_o = O.mvm(v.numpy(), return_intermediates=True)
    vd = v.cuda()
    # deterministic path: bit-exact at every stage
    sp = lat.splat(vd, mode=2)
    assert np.array_equal(bits(sp.cpu().numpy()), bits(sp_o))
    bl = lat.blur(sp)
    assert np.array_equal(bits(bl.cpu().numpy()), bits(bl_o))
    out = lat.slice(bl, mode=1)
    assert np.array_equal(bits(out.cpu().numpy()), bits(out_o))
    assert np.array_equal(bits(lat.mvm(vd, mode=2, blur="axis", exact=True).cpu().numpy()), bits(out_o))
    # locality order of the points: the sorted slice computes every output row with the same arithmetic
    assert lat.sorted is not None
    assert np.array_equal(np.sort(lat.sorted["perm"].cpu().numpy()), np.arange(N))
    assert np.array_equal(bits(lat.slice(bl, mode=1, sorted=True).cpu().numpy()), bits(out_o))
    assert _rel(lat.splat(vd, mode=1, sorted=True).cpu().numpy(), sp_o) < REL_TOL
    # production defaults (sorted atomic splat, fused multiply-adds, blur groups): 1e-5 relative
    assert _rel(lat.mvm(vd).cpu().numpy(), out_o) < REL_TOL
    assert _rel(lat.mvm(vd, blur="axis").cpu().numpy(), out_o) < REL_TOL
    assert _rel(lat.mvm(vd, sorted=False).cpu().numpy(), out_o) < REL_TOL
    assert _rel(lat.mvm(vd, sorted=False, blur="axis", exact=True).cpu().numpy(), out_o) < REL_TOL
    assert _rel(lat.slice(bl, mode=1, exact=False).cpu().numpy(), out_o) < 1e-6
    assert _rel(lat.blur(sp, exact=False).cpu().numpy(), bl_o) < 1e-6
    # row-sorted segmented-gather splat (the production splat)
    assert _rel(lat.splat(vd, mode=4).cpu().numpy(), sp_o) < REL_TOL
    assert _rel(lat.mvm(vd, mode=4).cpu().numpy(), out_o) < REL_TOL
    assert _rel(lat.mvm(vd, mode=1).cpu().numpy(), out_o) < REL_TOL
    # blur groups (several axes per launch through shared memory): same arithmetic per pass, bit-exact
    if lat.order > 0:
        assert lat.groups is not None
        bl_g = lat.blur(sp, groups=True)
        assert np.array_equal(bits(bl_g.cpu().numpy()), bits(bl_o))
        assert np.array_equal(bits(lat.mvm(vd, mode=2, blur="groups", exact=True).cpu().numpy()), bits(out_o))
        assert _rel(lat.blur(sp, groups=True, exact=False).cpu().numpy(), bl_o) < 1e-6
    # locality tiles: slice is the same arithmetic staged through shared memory (bit-exact on the same lattice
    # values); splat sums per-tile partials, then one reduction per segment (1e-5 relative)
    out_t = lat.slice(bl, mode=3)
    assert np.array_equal(bits(out_t.cpu().numpy()), bits(out_o))
    assert _rel(lat.slice(bl, mode=3, exact=False).cpu().numpy(), out_o) < 1e-6
    sp_t = lat.splat(vd, mode=3)
    assert _rel(sp_t.cpu().numpy(), sp_o) < REL_TOL
    assert _rel(lat.mvm(vd, mode=3).cpu().numpy(), out_o) < REL_TOL
    assert _rel(lat.mvm(vd, mode=3, blur="axis", exact=True).cpu().numpy(), out_o) < REL_TOL
    # atomic scatter path: 1e-5 relative
    sp_a = lat.splat(vd, mode=1)
    assert _rel(sp_a.cpu().numpy(), sp_o) < REL_TOL
    out_a = lat.mvm(vd, mode=1)
    assert _rel(out_a.cpu().numpy(), out_o) < REL_TOL


@pytest.mark.parametrize("group_axes,group_rows", [(1, 512), (2, 64), (4, 512), (9, 1500), (3, 16)])
def test_blur_group_partitions(sg, oracle, group_axes, group_rows):
    """Any partition of the axes into groups gives the per-axis result bit for bit; ranges whose classes do not fit
    are shortened, and a lattice whose single-axis lines do not fit falls back to the per-axis blur."""
    x, v = make_inputs(6000, 8, 8, seed=41)
    O = oracle.OracleLattice(x.numpy(), RBF2)
    out_o, sp_o, bl_o = O.mvm(v.numpy(), return_intermediates=True)
    lat = sg.Lattice(x.cuda(), RBF2, build_csr=True, group_axes=group_axes, group_rows=group_rows)
    sp = lat.splat(v.cuda(), mode=2)
    if lat.groups is None:
        assert group_rows == 16   # lines longer than 16 lattice points exist here
        assert np.array_equal(bits(lat.mvm(v.cuda(), mode=2, exact=True).cpu().numpy()), bits(out_o))
        return
    covered = [(g["j0"], g["j1"]) for g in lat.groups["list"]]
    assert covered[0][0] == 0 and covered[-1][1] == 9 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    assert all(g["rows_cap"] <= group_rows for g in lat.groups["list"])
    assert np.array_equal(bits(lat.blur(sp, groups=True).cpu().numpy()), bits(bl_o))
    assert np.array_equal(bits(lat.mvm(v.cuda(), mode=2, blur="groups", exact=True).cpu().numpy()), bits(out_o))


def test_groups_from_hash_table_without_neighbour_table(sg, oracle):
    """build_nbr=False: the blur groups are built straight from the key hash table, the (d+1) x M x 2r neighbour table
    never exists; the MVM is bit-identical to the one of the lattice that has it."""
    x, v = make_inputs(8000, 11, 8, seed=43)
    a = sg.Lattice(x.cuda(), MAT15_2, build_csr=True)
    b = sg.Lattice(x.cuda(), MAT15_2, build_csr=True, build_nbr=False)
    assert b.nbr is None and b.groups is not None
    for ga, gb in zip(a.groups["list"], b.groups["list"]):
        assert (ga["j0"], ga["j1"], ga["n_batches"]) == (gb["j0"], gb["j1"], gb["n_batches"])
        assert torch.equal(ga["lnb"], gb["lnb"]) and torch.equal(ga["src"], gb["src"])
    out_a = a.mvm(v.cuda(), mode=2, exact=True)
    out_b = b.mvm(v.cuda(), mode=2, exact=True)
    assert torch.equal(out_a, out_b)
    assert np.array_equal(bits(out_b.cpu().numpy()), bits(oracle.OracleLattice(x.numpy(), MAT15_2).mvm(v.numpy())))
    with pytest.raises(RuntimeError):
        b.mvm(v.cuda(), blur="axis")
    # a lattice whose lines do not fit a CTA falls back to the per-axis blur and builds the table after all
    x1, v1 = make_inputs(20000, 1, 2, seed=42, scale=2000.0)
    c = sg.Lattice(x1.cuda(), RBF1, build_nbr=False)
    assert c.groups is None and c.nbr is not None
    assert _rel(c.mvm(v1.cuda()).cpu().numpy(), oracle.OracleLattice(x1.numpy(), RBF1).mvm(v1.numpy())) < REL_TOL


def test_long_line_falls_back_to_axis_blur(sg, oracle):
    x, v = make_inputs(20000, 1, 2, seed=42, scale=2000.0)   # d = 1: one lattice line holds every point
    lat = sg.Lattice(x.cuda(), RBF1, build_csr=True)
    O = oracle.OracleLattice(x.numpy(), RBF1)
    assert lat.M == O.M and lat.M > 1000 and lat.groups is None
    assert np.array_equal(bits(lat.mvm(v.cuda(), mode=2, exact=True).cpu().numpy()), bits(O.mvm(v.numpy())))


def test_cuda_graph_replay(sg, oracle):
    x, v = make_inputs(5000, 6, 8, seed=51)
    O = oracle.OracleLattice(x.numpy(), RBF1)
    lat = sg.Lattice(x.cuda(), RBF1)
    src = torch.zeros(5000, 8, device="cuda")
    out = torch.empty(5000, 8, device="cuda")
    graph = lat.capture(src, out)
    for seed in (1, 2):
        w = torch.randn(5000, 8, generator=torch.Generator().manual_seed(seed))
        src.copy_(w)
        graph.replay()
        assert _rel(out.cpu().numpy(), O.mvm(w.numpy())) < REL_TOL


def test_filter_dropin_cpu_and_cuda_inputs(sg, oracle):
    x, v = make_inputs(3000, 4, 3, seed=5)
    c = torch.tensor(RBF1)
    want = oracle.filter(v.numpy(), x.numpy(), c.numpy())
    got_cpu = sg.filter(v, x, c)
    assert got_cpu.device.type == "cpu" and got_cpu.shape == v.shape
    assert _rel(got_cpu.numpy(), want) < REL_TOL
    got_gpu = sg.filter(v.cuda(), x.cuda(), c.cuda())
    assert got_gpu.is_cuda
    assert _rel(got_gpu.cpu().numpy(), want) < REL_TOL


def test_edge_cases(sg, oracle):
    # no points
    lat = sg.Lattice(torch.empty(0, 3, device="cuda"), RBF1)
    assert lat.M == 0 and lat.mvm(torch.empty(0, 2, device="cuda")).shape == (0, 2)
    # one point; many copies of one point (a lattice of d+1 points, every row touched N times)
    for x in (torch.tensor([[0.3, -1.2, 0.7]]), torch.tensor([[0.3, -1.2, 0.7]]).repeat(3000, 1)):
        v = torch.randn(x.shape[0], 3, generator=torch.Generator().manual_seed(1))
        O = oracle.OracleLattice(x.numpy(), RBF2)
        lat = sg.Lattice(x.cuda(), RBF2)
        assert lat.M == O.M == 4
        assert _rel(lat.mvm(v.cuda()).cpu().numpy(), O.mvm(v.numpy())) < REL_TOL
    # RHS and output that are column slices of wider tensors (row stride > L), odd channel counts
    x, _ = make_inputs(4000, 5, 1, seed=61)
    O = oracle.OracleLattice(x.numpy(), RBF1)
    lat = sg.Lattice(x.cuda(), RBF1)
    wide = torch.randn(4000, 12, generator=torch.Generator().manual_seed(2)).cuda()
    for lo, hi in ((0, 12), (2, 7), (1, 2), (4, 12)):
        src = wide[:, lo:hi]
        outw = torch.full((4000, 16), float("nan"), device="cuda")
        res = lat.mvm(src, out=outw[:, 3:3 + hi - lo])
        want = O.mvm(wide[:, lo:hi].contiguous().cpu().numpy())
        assert _rel(res.cpu().numpy(), want) < REL_TOL
        assert torch.isnan(outw[:, :3]).all() and torch.isnan(outw[:, 3 + hi - lo:]).all()
    # a column-major RHS is made contiguous by the host
    assert _rel(lat.mvm(wide.t().contiguous().t()).cpu().numpy(), O.mvm(wide.cpu().numpy())) < REL_TOL
    # wrong shapes / devices / dtypes fail loudly
    with pytest.raises(ValueError):
        lat.mvm(torch.zeros(3999, 2, device="cuda"))
    with pytest.raises(TypeError):
        lat.mvm(torch.zeros(4000, 2))
    with pytest.raises(TypeError):
        sg.Lattice(x.double().cuda(), RBF1)
    with pytest.raises(ValueError):
        lat.mvm(wide, coeffs=RBF2)


def test_against_reference_cuda_extension(sg):
    """The reference's own CUDA path (compiled unmodified into oracle/_ref/cuda_build by
    profiles/run_reference_cuda.py --build; shipped prebuilt) on the same GPU: allclose-level agreement, M well past the
    size where the reference's CPU table mis-files keys."""
    import importlib.util
    import os
    from conftest import ROOT
    so = os.path.join(ROOT, "oracle", "_ref", "cuda_build", "sgp_ref_gpu_lattice.so")
    if not os.path.exists(so):
        pytest.skip("reference CUDA extension not built")
    spec = importlib.util.spec_from_file_location("sgp_ref_gpu_lattice", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    x, v = make_inputs(60_000, 8, 4, seed=71)
    c = torch.tensor(RBF1).cuda()
    theirs = ref.filter(v.cuda(), x.cuda(), c)
    lat = sg.Lattice(x.cuda(), RBF1)
    assert lat.M > 3 * 16383
    assert _rel(lat.mvm(v.cuda()).cpu().numpy(), theirs.cpu().numpy()) < REL_TOL
    assert _rel(sg.filter(v.cuda(), x.cuda(), c).cpu().numpy(), theirs.cpu().numpy()) < REL_TOL


def test_hash_table_overflow_is_an_error(sg):
    x, _ = make_inputs(5000, 6, 1, seed=62)
    with pytest.raises(RuntimeError):
        sg.Lattice(x.cuda(), RBF1, hash_capacity=1024)


def test_key_range_error(sg):
    x = torch.full((10, 3), 1e6)
    with pytest.raises(RuntimeError):
        sg.Lattice(x.cuda(), RBF1)


@pytest.mark.parametrize("d", [1, 2, 3, 8, 11, 18, 23, 24, 30])
def test_exact_division(sg, d):
    """slice divides by the constant 1 + 2^-d with a 3-instruction Markstein sequence; it must equal the IEEE
    division (what the reference's `/` compiles to) for every finite fp32 input with |a| >= 2^-100 or a == 0,
    and be within 1e-37 absolute below that.  Exhaustive over all 2^32 bit patterns."""
    import ctypes as C
    from simplex_gp_b200 import _capi
    lib = _capi.lib()
    bad = torch.zeros(2, dtype=torch.int64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    chunk = 1 << 30
    for k in range(4):
        _capi.check(lib.sgp_debug_division_mismatches(d, k * chunk, chunk, C.c_void_p(bad.data_ptr()), st))
    assert bad.tolist() == [0, 0]


@pytest.mark.parametrize("N,Nn,d,coeffs", [(4000, 700, 5, RBF1), (900, 2500, 3, MAT15_2), (20000, 1, 8, RBF1),
                                            (1, 300, 2, RBF1), (500, 0, 4, RBF1)])
def test_extend_equals_union_build(sg, oracle, N, Nn, d, coeffs):
    """Lattice.extend (sgp_hash_seed / sgp_hash_extend / sgp_count_extension / sgp_number_extension): the lattice of
    cat([x, x_new]) made without revisiting x is bit for bit the one built from the concatenation -- and therefore
    the oracle's (first-touch numbering is sequential over the points, permutohedral.h:467-485)."""
    x, v = make_inputs(N + Nn, d, 4, seed=300 + N)
    xd = x.cuda()
    base = sg.Lattice(xd[:N], coeffs)
    ext = base.extend(xd[N:])
    full = sg.Lattice(xd, coeffs)
    assert (ext.N, ext.M) == (full.N, full.M)
    assert torch.equal(ext.keys, full.keys)
    assert torch.equal(ext.replay, full.replay)
    assert torch.equal(ext.nbr, full.nbr)
    assert torch.equal(ext.greedy, full.greedy) and torch.equal(ext.rank, full.rank)
    O = oracle.OracleLattice(x.numpy(), np.asarray(coeffs, dtype=np.float32))
    assert O.M == ext.M and np.array_equal(O.keys, ext.keys.cpu().numpy())
    assert np.array_equal(O.offsets, ext.offsets.cpu().numpy())
    vd = v.cuda()
    a, b = ext.mvm(vd, exact=True), full.mvm(vd, exact=True)
    assert float((a - b).norm()) <= 2e-6 * float(b.norm())      # same tables; splat atomics may reorder sums
    want = O.mvm(v.numpy())
    assert np.linalg.norm(a.cpu().numpy().astype(np.float64) - want) <= 1e-5 * np.linalg.norm(want)
    # the lattice that was extended is untouched
    again = sg.Lattice(xd[:N], coeffs)
    assert torch.equal(base.keys, again.keys) and torch.equal(base.replay, again.replay)
    assert float((base.mvm(vd[:N], exact=True) - again.mvm(vd[:N], exact=True)).norm()) <= 2e-6 * float(vd.norm())


def test_extend_twice_and_from_arrays(sg):
    x, v = make_inputs(6000, 6, 2, seed=909)
    xd = x.cuda()
    a = sg.Lattice(xd[:2000], RBF1, keep_structure=False).extend(xd[2000:3500]).extend(xd[3500:])
    full = sg.Lattice(xd, RBF1)
    assert torch.equal(a.keys, full.keys) and torch.equal(a.replay, full.replay) and torch.equal(a.nbr, full.nbr)
    assert a.greedy is None
    base = sg.Lattice(xd[:2000], RBF1)
    wrapped = sg.Lattice.from_arrays(RBF1, base.replay, base.keys, base.nbr)
    b = wrapped.extend(xd[2000:])
    assert torch.equal(b.keys, full.keys) and torch.equal(b.replay, full.replay)
    assert float((b.mvm(v.cuda(), exact=True) - full.mvm(v.cuda(), exact=True)).norm()) <= 2e-6 * float(v.norm())
    with pytest.raises(ValueError):
        base.extend(xd[:10, :3])


def test_extend_lazy_tables(sg, oracle):
    """lazy_tables=k: the first k products run on the neighbour table alone, the k+1-th builds the blur groups and the
    row-sorted entries; every product gives the oracle's numbers."""
    x, v = make_inputs(5000, 5, 8, seed=911)
    xd, vd = x.cuda(), v.cuda()
    lat = sg.Lattice(xd[:4000], RBF1).extend(xd[4000:], lazy_tables=2)
    assert lat.groups is None and lat.rows is None and lat.nbr is not None
    want = oracle.OracleLattice(x.numpy(), np.asarray(RBF1, dtype=np.float32)).mvm(v.numpy())
    for k in range(4):
        got = lat.mvm(vd).cpu().numpy()
        assert np.linalg.norm(got.astype(np.float64) - want) <= 1e-5 * np.linalg.norm(want)
        assert (lat.groups is not None and lat.rows is not None) == (k >= 2)
    lat2 = sg.Lattice(xd[:4000], RBF1).extend(xd[4000:], lazy_tables=5)
    out = torch.empty_like(vd)
    g = lat2.capture(vd, out)        # a graph wants the production chain: tables are built before the capture
    assert lat2.groups is not None and lat2.rows is not None
    g.replay()
    torch.cuda.synchronize()
    assert np.linalg.norm(out.cpu().numpy().astype(np.float64) - want) <= 1e-5 * np.linalg.norm(want)


def test_row_sorted_splat_on_a_point_subset_with_the_full_key_set(sg, oracle):
    """A rank's share of the points under point sharding: lattice arrays of the full point set, replay of a subset, so
    many lattice rows have no entry.  The row-sorted entries then carry one weightless filler per row
    (sgp_build_rowsorted, fill_rows = M); splat, blur and slice must agree with the oracle's partial splat."""
    x, v = make_inputs(6000, 6, 8, seed=77)
    full = sg.Lattice(x.cuda(), RBF1)
    lo, hi = 1500, 2600
    part = sg.Lattice.from_arrays(RBF1, full.replay[lo:hi].contiguous(), full.keys, full.nbr, build_csr=True)
    assert part.rows["entries"] == (hi - lo) * 7 + full.M            # fillers were needed
    whole = sg.Lattice.from_arrays(RBF1, full.replay, full.keys, full.nbr)
    assert whole.rows["entries"] == 6000 * 7                           # ... and are not when every row is touched
    vd = v.cuda()[lo:hi].contiguous()
    sp_rows = part.splat(vd, mode=4)
    sp_atomic = part.splat(vd, mode=1)
    sp_gather = part.splat(vd, mode=2)
    O = oracle.OracleLattice(x.numpy(), np.asarray(RBF1, dtype=np.float32))
    vz = np.zeros((6000, 8), dtype=np.float32)
    vz[lo:hi] = v.numpy()[lo:hi]
    _, sp_o, _ = O.mvm(vz, return_intermediates=True)
    for got in (sp_rows, sp_atomic, sp_gather):
        assert _rel(got.cpu().numpy(), sp_o) < REL_TOL
    want = O.mvm(vz)[lo:hi]
    assert _rel(part.mvm(vd).cpu().numpy(), want) < REL_TOL


@pytest.mark.parametrize("N,d,coeffs,shards", [(6000, 5, RBF1, 3), (2500, 11, MAT15_2, 2), (900, 2, RBF1, 4), (40, 3, RBF1, 5)])
def test_sharded_build_equals_full_build(sg, N, d, coeffs, shards):
    """Point-sharded lattice build, its ranks emulated in one process: each share of the points builds its own lattice,
    the key lists are merged in rank order (sgp_hash_append_keys / count / number) -- keys, vertex indices and the
    tables derived from the merged hash table are those of the lattice built from all points at once
    (the reference's sequential first-touch numbering, permutohedral.h:73-79,467-485)."""
    from simplex_gp_b200.distributed import merge_key_lists, shard_points
    x, v = make_inputs(N, d, 6, seed=N + shards)
    xd, vd = x.cuda(), v.cuda()
    full = sg.Lattice(xd, coeffs)
    parts = []
    for g in range(shards):
        lo, hi = shard_points(N, shards, g)
        parts.append(sg.Lattice(xd[lo:hi].contiguous(), coeffs, build_groups=False, build_rows=False, build_nbr=False))
        assert parts[-1].nbr is None and parts[-1].groups is None
    lists = [p.keys for p in parts]
    want_out = full.mvm(vd, mode=1, blur="axis", exact=True)
    total = torch.zeros(full.M, 6, device="cuda")
    outs = []
    for g in range(shards):
        lo, hi = shard_points(N, shards, g)
        keys, table, lmap = merge_key_lists(lists, want_map_of=g)
        assert torch.equal(keys, full.keys)
        replay = parts[g].replay.clone()
        replay[..., 0] = lmap[replay[..., 0].long()]
        assert torch.equal(replay, full.replay[lo:hi])
        loc = sg.Lattice.from_arrays(coeffs, replay, keys.contiguous(), None, table=table)
        assert torch.equal(loc.nbr, full.nbr)
        # this share's splat into the full lattice; the sum over the shares is the exchange step of point sharding
        total += loc.splat(vd[lo:hi].contiguous(), mode=4)
        outs.append(loc)
    assert float((total - full.splat(vd, mode=1)).norm() / total.norm()) < 1e-6
    for g, loc in enumerate(outs):
        lo, hi = shard_points(N, shards, g)
        state = {}
        got = loc.mvm(vd[lo:hi].contiguous(), after_splat=lambda vals: vals[:, :6].copy_(total))
        assert float((got - want_out[lo:hi]).norm() / want_out[lo:hi].norm()) < 1e-5
        # without the whole-lattice neighbour table: blur groups straight from the merged hash table
        keys, table, lmap = merge_key_lists(lists, want_map_of=g)
        lean = sg.Lattice.from_arrays(coeffs, loc.replay, keys.contiguous(), None, table=table, build_nbr=False)
        if lean.groups is not None:
            assert lean.nbr is None
            got = lean.mvm(vd[lo:hi].contiguous(), after_splat=lambda vals: vals[:, :6].copy_(total))
            assert float((got - want_out[lo:hi]).norm() / want_out[lo:hi].norm()) < 1e-5


def test_hash_table_sized_from_the_previous_build(sg, oracle):
    """The build sizes its hash table from the lattice of the last build of the same (N, d) (4x its M instead of
    2 N(d+1)); a lattice that outgrew the hint fills the table, which the insertion reports quickly (bounded probing) and
    the build repeats at the safe size.  Numbering never depends on the table."""
    from simplex_gp_b200 import lattice as LT
    LT._M_BUILD_HINT.clear()
    x, v = make_inputs(20000, 4, 3, seed=91)
    xd = x.cuda()
    first = sg.Lattice(xd, RBF1)                      # no hint: safe size
    again = sg.Lattice(xd.clone(), RBF1)              # hinted: small table
    assert again.hash_capacity < first.hash_capacity
    assert again.M == first.M and torch.equal(again.keys, first.keys) and torch.equal(again.replay, first.replay)
    assert torch.equal(again.nbr, first.nbr)
    big = sg.Lattice((xd * 6).contiguous(), RBF1)     # many more lattice points than the hint allows for
    O = oracle.OracleLattice((x * 6).numpy(), RBF1)
    assert big.M == O.M > 4 * first.M
    assert np.array_equal(big.keys.cpu().numpy(), O.keys) and np.array_equal(big.offsets.cpu().numpy(), O.offsets)
    assert np.array_equal(big.nbr.cpu().numpy(), O.nbr)
    small = sg.Lattice((xd * 0.2).contiguous(), RBF1)  # far fewer: the hinted table is merely roomy
    O = oracle.OracleLattice((x * 0.2).numpy(), RBF1)
    assert small.M == O.M and np.array_equal(small.keys.cpu().numpy(), O.keys)
    LT._M_BUILD_HINT.clear()


@pytest.mark.parametrize("N,d,L,coeffs", [(5000, 8, 16, RBF1), (3000, 5, 11, MAT15_2), (2000, 3, 4, RBF1), (900, 2, 1, RBF2),
                                           (70, 8, 16, RBF1), (4097, 11, 32, RBF1)])
@pytest.mark.parametrize("scan", [0, 1])
def test_ring_kernels_forced_at_small_sizes(sg, oracle, N, d, L, coeffs, scan):
    """The TMA-ring splat / slice (csrc/sgp_ring.cu) are selected for large dense lattices and wide rows only; here they
    are forced (SGP_RING_FORCE) on small and ragged shapes -- partial tiles, 11 -> 12 padded channels, one channel chunk,
    fewer tiles than warps -- in both reduction forms (per-run reductions + warp-uniform aggregation; tile scan + stores)."""
    import os
    x, v = make_inputs(N, d, L, seed=N + L + scan)
    want = oracle.OracleLattice(x.numpy(), coeffs).mvm(v.numpy())
    lat = sg.Lattice(x.cuda(), coeffs)
    old = {k: os.environ.get(k) for k in ("SGP_RING_FORCE", "SGP_SPLAT_SCAN")}
    os.environ.update(SGP_RING_FORCE="1", SGP_SPLAT_SCAN=str(scan))
    try:
        got = lat.mvm(v.cuda()).cpu().numpy()
        graph_out = torch.empty(N, L, device="cuda")
        graph = lat.capture(v.cuda(), graph_out)
        for _ in range(3):       # the splat buffer is re-zeroed at the end of every replay
            graph.replay()
        torch.cuda.synchronize()
    finally:
        for k, val in old.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
    assert _rel(got, want) < REL_TOL
    assert _rel(graph_out.cpu().numpy(), want) < REL_TOL


@pytest.mark.parametrize("L", [5, 11])
def test_ragged_right_hand_sides_through_the_padded_copy(sg, oracle, L):
    """L = 11 (the reference's training block [y | 10 probes]): at large N the MVM copies the block into a zero-padded
    one so that the splat gathers 16-byte vectors (SGP_MVM_SRC_PADDED); forced here at a size the oracle finishes, eager
    and captured, and compared with the channel-by-channel path."""
    import os
    N, d = 3000, 6
    x, v = make_inputs(N, d, L, seed=L)
    want = oracle.OracleLattice(x.numpy(), RBF1).mvm(v.numpy())
    lat = sg.Lattice(x.cuda(), RBF1)
    old = os.environ.get("SGP_PAD_SRC")
    try:
        os.environ["SGP_PAD_SRC"] = "0"
        plain = lat.mvm(v.cuda()).cpu().numpy()
        os.environ["SGP_PAD_SRC"] = "1"
        src = v.cuda()
        got = lat.mvm(src).cpu().numpy()
        out = torch.empty(N, L, device="cuda")
        graph = lat.capture(src, out)
        src.mul_(2.0)                      # the graph reads the caller's buffer at replay time, not a stale copy
        graph.replay()
        graph.replay()
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("SGP_PAD_SRC", None)
        else:
            os.environ["SGP_PAD_SRC"] = old
    assert _rel(got, want) < REL_TOL and _rel(plain, want) < REL_TOL
    assert _rel(out.cpu().numpy(), 2.0 * want) < REL_TOL


@pytest.mark.parametrize("N,d,L", [(1000, 4, 16), (333, 7, 12), (77, 3, 8), (2049, 8, 20)])
def test_ring_kernels_stay_inside_their_buffers(sg, N, d, L):
    """Guard bands around every buffer the TMA-ring kernels write (compute-sanitizer is not available on the GPU pool):
    lattice values and outputs are views into the middle of larger NaN-filled allocations; after the forced ring splat
    and slice (partial tiles, entry counts that are not a multiple of the tile) the bands are untouched, and so is the
    zero-padded copy of a ragged block."""
    import ctypes as C
    import os
    from simplex_gp_b200 import _capi
    from simplex_gp_b200.lattice import _ptr, _stream_ptr
    x, v = make_inputs(N, d, L, seed=3 * N + L)
    lat = sg.Lattice(x.cuda(), RBF1)
    lat.mvm(v.cuda())                      # builds the lazy tables
    lib, st, M, rows = _capi.lib(), _stream_ptr(lat.device), lat.M, lat.rows
    G = 4096                                # guard floats either side
    nan = float("nan")
    vals_all = torch.full((G + M * L + G,), nan, device="cuda")
    out_all = torch.full((G + N * L + G,), nan, device="cuda")
    vals = vals_all[G:G + M * L].view(M, L)
    out = out_all[G:G + N * L].view(N, L)
    src = v.cuda()
    old = {k: os.environ.get(k) for k in ("SGP_RING_FORCE", "SGP_SPLAT_SCAN")}
    try:
        os.environ["SGP_RING_FORCE"] = "1"
        want = None
        for scan in ("0", "1"):
            os.environ["SGP_SPLAT_SCAN"] = scan
            vals.fill_(nan)
            _capi.check(lib.sgp_splat_rows(_ptr(rows["ent"]), _ptr(rows["seg_row"]), rows["n"], N, M, _ptr(src), src.stride(0),
                                           L, _ptr(vals), L, st))
            torch.cuda.synchronize()
            assert torch.isnan(vals_all[:G]).all() and torch.isnan(vals_all[G + M * L:]).all()
            assert not torch.isnan(vals).any()
            if want is None:
                want = vals.clone()
            else:
                assert float((vals - want).abs().max()) <= 1e-5 * float(want.abs().max())
        v_out = lat._view(lat._table(False, True), None, lat.exact)     # slice of values held in lattice-index order
        _capi.check(lib.sgp_slice(C.byref(v_out), _ptr(vals), L, _ptr(out), out.stride(0), L, st))
        torch.cuda.synchronize()
        assert torch.isnan(out_all[:G]).all() and torch.isnan(out_all[G + N * L:]).all()
        assert not torch.isnan(out).any()
    finally:
        for k, val in old.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val


@pytest.mark.parametrize("L", [9, 11, 12])
@pytest.mark.parametrize("ring", ["0", "1"])
def test_twelve_columns_on_sixteen_channel_lattice_rows(sg, oracle, L, ring):
    """9-12 column blocks run on 16-channel lattice rows on large dense lattices (Lattice.lattice_width): the splat's and the
    slice's spare lanes idle.  Forced here at a size the oracle finishes (SGP_LV16=2), with the one-shot and the ring
    kernels, eager, captured, and through the CG form (sweep folded into the slice)."""
    import os
    from simplex_gp_b200 import _capi
    N, d = 4000, 5
    x, v = make_inputs(N, d, L, seed=40 + L)
    want = oracle.OracleLattice(x.numpy(), RBF1).mvm(v.numpy())
    lat = sg.Lattice(x.cuda(), RBF1)
    src = v.cuda()
    lat.mvm(src)
    old = {k: os.environ.get(k) for k in ("SGP_LV16", "SGP_RING_FORCE", "SGP_PAD_SRC")}
    os.environ.update(SGP_LV16="2", SGP_RING_FORCE=ring, SGP_PAD_SRC="1")
    try:
        assert lat.lattice_width(L) == 16
        got = lat.mvm(src).cpu().numpy()
        out = torch.empty(N, L, device="cuda")
        graph = lat.capture(src, out)
        graph.replay()
        graph.replay()
        torch.cuda.synchronize()
        cg_out = cg_dot = None
        if L % 4 == 0:
            lib = _capi.lib()
            s, noise = torch.tensor([0.5], device="cuda"), torch.tensor([0.25], device="cuda")
            cg_out = torch.empty(N, L, device="cuda")
            cg_dot = torch.empty(L, device="cuda")
            scratch = torch.empty(int(lib.sgp_cg_scratch_floats(L)), device="cuda")
            lat.mvm(src, out=cg_out, cg=(s, noise, cg_dot, scratch))
            torch.cuda.synchronize()
    finally:
        for k, val in old.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
    assert _rel(got, want) < REL_TOL
    assert _rel(out.cpu().numpy(), want) < REL_TOL
    if cg_out is not None:
        ap = 0.5 * want + 0.25 * v.numpy()
        assert _rel(cg_out.cpu().numpy(), ap) < REL_TOL
        dots = (v.numpy().astype(np.float64) * ap.astype(np.float64)).sum(0)
        assert np.abs(cg_dot.cpu().numpy() - dots).max() <= 1e-4 * np.abs(dots).max()

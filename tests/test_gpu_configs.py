"""Full-size checks at the BASELINE.json configurations (B200), through size-independent properties: the oracle cannot
run these sizes in seconds (config C takes the reference ~400 s on a CPU), so the CUDA path is checked against
  * structural invariants of the lattice (weights sum to 1, indices in range, neighbour tables mutually inverse, keys
    unique, every lattice point touched),
  * linearity of the operator,
  * agreement of the production path (row-sorted splat, blur groups, fused multiply-adds) with the independent plain
    path (atomic scatter splat, one blur launch per axis, reference arithmetic),
  * bit-exact agreement with the oracle on a prefix of the points (lattice of the first n points).
"""
import numpy as np
import pytest
import torch

from conftest import MAT15_2, MAT15_3, RBF1, bits

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: N, d, L, coeffs   (BASELINE.json configs[1..3]; D is run at a tenth of its N)
    "A_metric": (1_000_000, 8, 16, RBF1),
    "B_elevators": (16_600, 18, 11, RBF1),
    "C_houseelectric": (2_050_000, 11, 16, MAT15_2),
    "D_stress_tenth": (1_000_000, 24, 4, MAT15_3),
}


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_size_properties(sg, oracle, name):
    N, d, L, coeffs = CONFIGS[name]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, d, generator=g)
    xd = x.cuda()
    lat = sg.Lattice(xd, coeffs)
    M, r = lat.M, lat.order
    assert 0 < M <= N * (d + 1)
    # ---- structure ------------------------------------------------------------------------------------------
    w = lat.weights
    assert float((w.sum(1) - 1).abs().max()) < 1e-5 and float(w.min()) > -1e-5
    off = lat.offsets
    assert int(off.min()) >= 0 and int(off.max()) == M - 1
    assert int(torch.bincount(off.reshape(-1).long(), minlength=M).min()) >= 1      # every lattice point is touched
    nbr = lat.nbr.long()
    idx = torch.arange(M, device="cuda")
    for j in (0, d // 2, d):
        for t in range(r):          # offset -(r-t) and its mirror +(r-t)
            lo, hi = nbr[j, :, t], nbr[j, :, 2 * r - 1 - t]
            has = lo >= 0
            assert bool((nbr[j, lo[has], 2 * r - 1 - t] == idx[has]).all())
            has = hi >= 0
            assert bool((nbr[j, hi[has], t] == idx[has]).all())
    # keys are distinct: compare a 64-bit polynomial hash of the rows (collisions are astronomically unlikely)
    k = lat.keys.long()
    mult = torch.tensor([(1_000_003 ** (i + 1)) % (2 ** 61 - 1) for i in range(d)], device="cuda")
    h = (k * mult).sum(1)
    assert int(torch.unique(h).numel()) == M
    # ---- prefix parity with the oracle ---------------------------------------------------------------------------
    n = 3000
    O = oracle.OracleLattice(x[:n].numpy(), coeffs)
    sub = sg.Lattice(xd[:n].contiguous(), coeffs)
    assert sub.M == O.M and np.array_equal(sub.keys.cpu().numpy(), O.keys)
    assert np.array_equal(sub.offsets.cpu().numpy(), O.offsets)
    assert np.array_equal(lat.greedy[:n].cpu().numpy(), O.greedy) and np.array_equal(lat.rank[:n].cpu().numpy(), O.rank)
    # ---- operator ---------------------------------------------------------------------------------------------------
    u = torch.randn(N, L, generator=g).cuda()
    v = torch.randn(N, L, generator=g).cuda()
    Ku, Kv = lat.mvm(u).clone(), lat.mvm(v).clone()
    assert torch.isfinite(Ku).all()
    comb = lat.mvm(0.5 * u - 2.0 * v)
    assert _rel(comb, 0.5 * Ku - 2.0 * Kv) < 2e-5
    plain = lat.mvm(u, mode=1, blur="axis", exact=True, sorted=False)
    assert _rel(Ku, plain) < 1e-5
    ones = lat.mvm(torch.ones(N, 1, device="cuda"))
    assert float(ones.min()) > 0


def test_full_stress_configuration_on_one_gpu(sg, oracle):
    """BASELINE.json configs[4] at FULL size on one B200: N = 10M, d = 24, Matern-1.5 order 3 (M = 250M lattice points,
    ~122 GB).  The blur groups are built straight from the hash table (build_nbr=False): the whole-lattice neighbour
    table would be 150 GB.  The reference cannot run this size at all (its `int keyIdx = filled*kd` overflows,
    permutohedral.h:75,166)."""
    free, total = torch.cuda.mem_get_info()
    if total < 160e9:
        pytest.skip("needs a 180 GB GPU")
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    N, d, L = 10_000_000, 24, 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, d, generator=g)
    try:
        lat = sg.Lattice(x.cuda(), MAT15_3, build_nbr=False)
    except torch.OutOfMemoryError:
        pytest.skip("not enough free GPU memory in this process")
    M = lat.M
    assert 0.99 * N * (d + 1) < M <= N * (d + 1) and lat.nbr is None and lat.groups is not None
    w = lat.weights
    assert float((w.sum(1) - 1).abs().max()) < 1e-5
    assert int(lat.offsets.min()) >= 0 and int(lat.offsets.max()) == M - 1
    n = 2000
    O = oracle.OracleLattice(x[:n].numpy(), MAT15_3)
    assert np.array_equal(lat.greedy[:n].cpu().numpy(), O.greedy) and np.array_equal(lat.rank[:n].cpu().numpy(), O.rank)
    assert np.array_equal(bits(lat.weights[:n].cpu().numpy()), bits(O.weights))
    u = torch.randn(N, L, generator=g).cuda()
    v = torch.randn(N, L, generator=g).cuda()
    Ku, Kv = lat.mvm(u).clone(), lat.mvm(v).clone()
    assert torch.isfinite(Ku).all()
    assert _rel(lat.mvm(0.5 * u - 2.0 * v), 0.5 * Ku - 2.0 * Kv) < 2e-5
    assert float(lat.mvm(torch.ones(N, 1, device="cuda")).min()) > 0
    del lat
    gc.collect()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("L", [1, 16])
def test_metric_configuration_against_oracle_at_full_size(sg, oracle, L):
    """BASELINE.json configs[1] / the metric shape, ALL points: N = 1M, d = 8, RBF order 1 against the oracle's
    restatement of the reference (permutohedral.h:259-340) -- lattice structure bit-exact (M, keys in first-touch
    order, vertex indices, neighbour table), the deterministic path (ordered-gather splat, reference arithmetic)
    bit-exact on every output, the production path within the north star's 1e-5.  Also measures how far the UNMODIFIED
    reference (stale-bucket defect of its hash-table growth, permutohedral.h:104-106 vs :61-63; oracle
    ``reference_table=True``) is from the correct table the product implements; the figure goes to DESIGN.md section 2."""
    import json
    import os
    N, d = 1_000_000, 8
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, d, generator=g)
    v = torch.randn(N, L, generator=g)
    O = oracle.OracleLattice(x.numpy(), RBF1)
    lat = sg.Lattice(x.cuda(), RBF1, build_csr=True)
    assert lat.M == O.M
    assert np.array_equal(lat.keys.cpu().numpy(), O.keys)
    assert np.array_equal(lat.offsets.cpu().numpy(), O.offsets)
    assert np.array_equal(bits(lat.weights.cpu().numpy()), bits(O.weights))
    assert np.array_equal(lat.greedy.cpu().numpy(), O.greedy) and np.array_equal(lat.rank.cpu().numpy(), O.rank)
    assert np.array_equal(lat.nbr.cpu().numpy(), O.nbr)
    want = O.mvm(v.numpy())
    vd = v.cuda()
    exact = lat.mvm(vd, mode=2, exact=True).cpu().numpy()
    assert np.array_equal(bits(exact), bits(want))                      # bit-exact on all N x L outputs
    prod = lat.mvm(vd).cpu().numpy().astype(np.float64)
    rel = float(np.linalg.norm(prod - want) / np.linalg.norm(want))
    assert rel < 1e-5
    one_call = sg.filter(vd, x.cuda(), torch.tensor(RBF1)).cpu().numpy().astype(np.float64)
    rel_filter = float(np.linalg.norm(one_call - want) / np.linalg.norm(want))
    assert rel_filter < 1e-5
    # the unmodified reference: same geometry, a few orphaned / duplicated lattice points per table doubling
    R = oracle.OracleLattice(x.numpy(), RBF1, reference_table=True)
    ref = R.mvm(v.numpy()).astype(np.float64)
    dev = np.abs(ref - want)
    rel_ref = float(np.linalg.norm(ref - want) / np.linalg.norm(want))
    rows_off = int((dev.max(axis=1) > 1e-6 * np.abs(want).max()).sum())
    worst = float(dev.max() / np.abs(want).max())
    assert 0 < rel_ref < 2e-2 and R.M != O.M
    rec = {"N": N, "d": d, "L": L, "M": int(O.M), "M_unmodified_reference": int(R.M),
           "production_rel_l2_vs_oracle": rel, "one_call_filter_rel_l2_vs_oracle": rel_filter,
           "unmodified_reference_rel_l2_vs_correct_table": rel_ref,
           "unmodified_reference_rows_differing": rows_off, "unmodified_reference_worst_abs_over_max": worst}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"parity_full_size_L{L}.json"), "w") as f:
            json.dump(rec, f, indent=1)
    print(json.dumps(rec))


@pytest.mark.parametrize("L", [11, 12])
def test_training_block_at_full_size(sg, oracle, L):
    """The reference's training block ([y | 10 probes]: 11 columns; 12 as this package's solver pads it) at the metric
    shape, ALL points: here the production chain runs on 16-channel lattice rows (Lattice.lattice_width), the 11-column
    block through the zero-padded copy; eager, captured and, for 12 columns, the CG form with the sweep folded into the
    slice -- all within the north star's 1e-5 of the oracle."""
    from simplex_gp_b200 import _capi
    N, d = 1_000_000, 8
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, d, generator=g)
    v = torch.randn(N, L, generator=g)
    want = oracle.OracleLattice(x.numpy(), RBF1).mvm(v.numpy()).astype(np.float64)
    lat = sg.Lattice(x.cuda(), RBF1)
    vd = v.cuda()
    lat.mvm(vd)                                   # builds the postponed tables
    assert lat.lattice_width(L) == 16
    prod = lat.mvm(vd).cpu().numpy().astype(np.float64)
    assert float(np.linalg.norm(prod - want) / np.linalg.norm(want)) < 1e-5
    out = torch.empty(N, L, device="cuda")
    graph = lat.capture(vd, out)
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert float(np.linalg.norm(out.cpu().numpy() - want) / np.linalg.norm(want)) < 1e-5
    if L % 4 == 0:
        lib = _capi.lib()
        s, noise = torch.tensor([0.5], device="cuda"), torch.tensor([0.25], device="cuda")
        ap = torch.empty(N, L, device="cuda")
        dots = torch.empty(L, device="cuda")
        scratch = torch.empty(int(lib.sgp_cg_scratch_floats(L)), device="cuda")
        lat.mvm(vd, out=ap, cg=(s, noise, dots, scratch))
        ref_ap = 0.5 * want + 0.25 * v.numpy().astype(np.float64)
        assert float(np.linalg.norm(ap.cpu().numpy() - ref_ap) / np.linalg.norm(ref_ap)) < 1e-5
        ref_dots = (v.numpy().astype(np.float64) * ref_ap).sum(0)
        assert float(np.abs(dots.cpu().numpy() - ref_dots).max() / np.abs(ref_dots).max()) < 1e-4

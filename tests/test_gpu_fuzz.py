"""Seeded random sweep of shapes and input distributions against the oracle: structure bit-exact, the deterministic
chain bit-exact, the production chain within 1e-5.  The distributions aim at what the reference's arithmetic is
sensitive to: rank ties (points on lattice hyperplanes: integer grids, all-equal coordinates, exact zeros), repeated
points, clusters much tighter than a lattice cell, widely spread points (large keys), one-point and one-dimensional
inputs, column counts that are not multiples of the vector width."""
import numpy as np
import pytest
import torch

from conftest import MAT15_2, MAT15_3, RBF1, RBF2, bits

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
STENCILS = [RBF1, RBF1, RBF2, MAT15_2, MAT15_3]
DISTS = ["randn", "rand", "grid", "equal", "clusters", "repeats", "wide", "zeros_mixed"]


def _draw(seed):
    g = torch.Generator().manual_seed(7000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    d = [1, 2, 3, 5, 8, 11, 18, 24, 27][ri(0, 8)]
    N = [1, 2, 17, 300, 2500, 9000][ri(0, 5)]
    if d >= 18:
        N = min(N, 2500)
    L = [1, 2, 3, 4, 5, 8, 11, 16, 23, 37][ri(0, 9)]
    coeffs = STENCILS[ri(0, len(STENCILS) - 1)]
    dist = DISTS[seed % len(DISTS)]
    if dist == "randn":
        x = torch.randn(N, d, generator=g)
    elif dist == "rand":
        x = torch.rand(N, d, generator=g)
    elif dist == "grid":        # integer multiples of a step: many points exactly on lattice hyperplanes
        x = torch.randint(-3, 4, (N, d), generator=g).float() * [0.25, 0.5, 1.0][ri(0, 2)]
    elif dist == "equal":       # all coordinates of a point equal: every elevated difference ties
        x = torch.randn(N, 1, generator=g).round().repeat(1, d)
    elif dist == "clusters":
        c = torch.randn(max(1, N // 50), d, generator=g) * 2
        x = c[torch.randint(0, c.shape[0], (N,), generator=g)] + 1e-3 * torch.randn(N, d, generator=g)
    elif dist == "repeats":
        base = torch.randn(max(1, N // 7), d, generator=g)
        x = base[torch.randint(0, base.shape[0], (N,), generator=g)]
    elif dist == "wide":
        x = torch.randn(N, d, generator=g) * 40.0
    else:                       # exact zeros and signed zeros mixed into ordinary points
        x = torch.randn(N, d, generator=g)
        x[torch.rand(N, d, generator=g) < 0.4] = 0.0
        x[torch.rand(N, d, generator=g) < 0.1] = -0.0
    v = torch.randn(N, L, generator=g)
    return x.contiguous(), v.contiguous(), coeffs, dist


def _rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("seed", range(48))
def test_random_shape_against_oracle(sg, oracle, seed):
    x, v, coeffs, dist = _draw(seed)
    O = oracle.OracleLattice(x.numpy(), np.asarray(coeffs, dtype=np.float32))
    lat = sg.Lattice(x.cuda(), coeffs, build_csr=True)
    what = f"seed={seed} dist={dist} N={x.shape[0]} d={x.shape[1]} L={v.shape[1]} r={len(coeffs) // 2} M={O.M}"
    assert lat.M == O.M, what
    assert np.array_equal(lat.greedy.cpu().numpy(), O.greedy), what
    assert np.array_equal(lat.rank.cpu().numpy(), O.rank), what
    assert np.array_equal(bits(lat.weights.cpu().numpy()), bits(O.weights)), what
    assert np.array_equal(lat.keys.cpu().numpy(), O.keys), what
    assert np.array_equal(lat.offsets.cpu().numpy(), O.offsets), what
    assert np.array_equal(lat.nbr.cpu().numpy(), O.nbr), what
    want = O.mvm(v.numpy())
    vd = v.cuda()
    assert np.array_equal(bits(lat.mvm(vd, mode=2, blur="axis", exact=True).cpu().numpy()), bits(want)), what
    scale = max(np.linalg.norm(want.astype(np.float64)), 1e-30)
    for kw in ({}, {"exact": True}, {"mode": 1, "blur": "axis"}):
        got = lat.mvm(vd, **kw).cpu().numpy()
        assert np.isfinite(got).all(), what
        assert np.linalg.norm(got.astype(np.float64) - want) <= REL_TOL * scale, (what, kw)
    # the same lattice grown in two steps
    if x.shape[0] >= 2:
        k = x.shape[0] // 3 + 1
        ext = sg.Lattice(x[:k].cuda(), coeffs).extend(x[k:].cuda())
        assert torch.equal(ext.keys, lat.keys) and torch.equal(ext.replay, lat.replay), what
        assert torch.equal(ext.nbr, lat.nbr), what

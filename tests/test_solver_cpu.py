"""Host-side solver logic on CPU (no lattice, a dense SPD kernel matrix stands in for the operator): the blocked
pivoted-Cholesky preconditioner the reference's solver settings ask for (experiments/train_simplexgp.py:34-37,63-67:
``max_preconditioner_size = 100``), its symmetric use inside batched CG, and the log-determinant correction."""
import math

import torch

from simplex_gp_b200 import gp


def _problem(n=500, noise=0.01, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 2, generator=g)
    K = torch.exp(-0.5 * ((x[:, None] - x[None]) ** 2).sum(-1) / 0.5 ** 2)
    y = torch.sin(x[:, 0]) + 0.1 * torch.randn(n, generator=g)
    return K, y, torch.tensor(1.7), torch.tensor(noise)


def test_pivoted_cholesky_is_a_low_rank_factor():
    K, _, s, _ = _problem()
    errs = []
    for rank in (10, 40, 160):
        Lm = gp.pivoted_cholesky(lambda V: s * (K @ V), K.shape[0], float(s), rank=rank, block=16)
        assert Lm.shape == (K.shape[0], rank) and torch.isfinite(Lm).all()
        errs.append(float((s * K - Lm @ Lm.T).norm() / (s * K).norm()))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 0.05
    # block = 1 is the textbook sequence: exact on a matrix of that rank
    B = torch.randn(60, 7, generator=torch.Generator().manual_seed(1))
    G = B @ B.T
    G = G / G.diagonal().max()
    Lm = gp.pivoted_cholesky(lambda V: G @ V, 60, 1.0, rank=7, block=1, rel_tol=1e-6)
    # (the residual starts from the nominal diagonal 1.0, as GPyTorch's does from LazyTensor.diag(); G's is <= 1)
    assert torch.isfinite(Lm).all()


def test_preconditioner_algebra():
    K, _, s, noise = _problem()
    n = K.shape[0]
    Lm = gp.pivoted_cholesky(lambda V: s * (K @ V), n, float(s), rank=50)
    pre = gp.LowRankPreconditioner(Lm, float(noise))
    P = (noise * torch.eye(n) + Lm @ Lm.T).double()
    ev, Q = torch.linalg.eigh(P)
    want = ((Q * ev.rsqrt()) @ Q.T).float()
    V = torch.randn(n, 4, generator=torch.Generator().manual_seed(2))
    assert float((pre.inv_sqrt(V) - want @ V).norm() / (want @ V).norm()) < 1e-3
    assert abs(pre.logdet() - float(torch.logdet(P))) < 1e-2


def test_preconditioned_cg_same_answer_fewer_iterations():
    K, y, s, noise = _problem()
    n = K.shape[0]
    probes = torch.randn(n, 20, generator=torch.Generator().manual_seed(3)).sign()
    st0, st1 = {}, {}
    v0, s0 = gp.mll_cg(lambda V: K @ V, y, torch.tensor(0.0), s, noise, probes=probes, tol=1e-6, max_iter=3000, stats=st0)
    v1, s1 = gp.mll_cg(lambda V: K @ V, y, torch.tensor(0.0), s, noise, probes=probes, tol=1e-6, max_iter=3000,
                       preconditioner_size=60, stats=st1)
    assert st1["cg_iterations"] < 0.6 * st0["cg_iterations"] and st1["preconditioner_rank"] > 40
    A = (s * K + noise * torch.eye(n)).double()
    Lc = torch.linalg.cholesky(A)
    a = torch.cholesky_solve(y.double()[:, None], Lc)
    exact = float((-0.5 * ((y.double()[:, None] * a).sum() + 2 * torch.log(torch.diagonal(Lc)).sum()
                           + n * math.log(2 * math.pi))) / n)
    # the quadratic term is exact on both sides; the log-determinant is a 20-probe stochastic estimate
    assert abs(v0 - exact) < 0.03 and abs(v1 - exact) < 0.03


def test_cg_stopping_rules_on_cpu():
    """stop="mean" (GPyTorch's rule) never runs longer than stop="all"; min_iter is honoured; a zero right-hand side stays
    zero and does not hold the mean rule up."""
    from simplex_gp_b200 import gp
    g = torch.Generator().manual_seed(3)
    Q = torch.randn(300, 40, generator=g)
    K = Q @ Q.T / 40 + 0.05 * torch.eye(300)
    B = torch.randn(300, 5, generator=g)
    B[:, 2] = 0.0
    A = lambda V: K @ V
    _, a_all, _ = gp.batched_cg(A, B, tol=0.05, max_iter=200, stop="all")
    X, a_mean, _ = gp.batched_cg(A, B, tol=0.05, max_iter=200, stop="mean")
    assert 1 <= a_mean.shape[0] <= a_all.shape[0]
    assert float(X[:, 2].abs().max()) == 0.0
    _, a_min, _ = gp.batched_cg(A, B, tol=0.5, max_iter=200, stop="mean", min_iter=20)
    assert a_min.shape[0] == 20
    _, a_cap, _ = gp.batched_cg(A, B, tol=0.5, max_iter=5, stop="mean", min_iter=20)
    assert a_cap.shape[0] <= 5

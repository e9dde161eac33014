"""Marginal-likelihood parity and the reference's one acceptance test (tests/train_snelson.py), restated without
GPyTorch: the SAME solver (simplex_gp_b200.gp) runs over this package's CUDA filter and over the CPU oracle filter,
so the only difference is the MVM backend."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, RBF1, make_inputs

pytestmark = pytest.mark.gpu


def _snelson():
    g = np.load(os.path.join(GOLDEN_DIR, "kat.npz"))
    return torch.tensor(g["snelson/x"]), torch.tensor(g["snelson/v"][:, 0])


def _oracle_matmul(oracle, x_scaled, coeffs):
    O = oracle.OracleLattice(x_scaled.numpy(), coeffs)
    return lambda V: torch.from_numpy(O.mvm(V.contiguous().numpy()))


def test_dense_mll_matches_oracle_backend(sg, oracle):
    from simplex_gp_b200 import gp
    x, y = _snelson()
    ls, s, noise, mu = 0.8, torch.tensor(1.3), torch.tensor(0.05), torch.tensor(0.1)
    want = gp.mll_dense(_oracle_matmul(oracle, x / ls, RBF1), y, mu, s, noise)
    lat = sg.Lattice((x / ls).cuda(), RBF1)
    got = gp.mll_dense(lambda V: lat.mvm(V), y.cuda(), mu.cuda(), s.cuda(), noise.cuda())
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))


def test_cg_mll_matches_oracle_backend(sg, oracle):
    from simplex_gp_b200 import gp
    N, d = 3000, 4
    x, v = make_inputs(N, d, 1, seed=12)
    y = torch.sin(x.sum(1)) + 0.1 * v[:, 0]
    ls, s, noise, mu = 1.2, torch.tensor(0.9), torch.tensor(0.2), torch.tensor(0.0)
    probes = torch.randn(N, 10, generator=torch.Generator().manual_seed(5)).sign()
    want, _ = gp.mll_cg(_oracle_matmul(oracle, x / ls, RBF1), y, mu, s, noise, probes=probes, tol=1e-6)
    lat = sg.Lattice((x / ls).cuda(), RBF1)
    got, _ = gp.mll_cg(lambda V: lat.mvm(V), y.cuda(), mu.cuda(), s.cuda(), noise.cuda(), probes=probes, tol=1e-6)
    assert abs(got - want) <= 1e-5 * abs(want), (got, want)


def _exact_rbf_mll(x, y, raw):
    import math
    sp = torch.nn.functional.softplus
    ls, s, noise, mu = sp(raw[0]), sp(raw[1]), sp(raw[2]) + 1e-4, raw[3]
    d2 = (x[:, None, :] - x[None, :, :]).pow(2).sum(-1) / ls ** 2
    K = s * torch.exp(-0.5 * d2) + noise * torch.eye(x.shape[0], device=x.device)
    Lc = torch.linalg.cholesky(K)
    r = (y - mu).unsqueeze(-1)
    a = torch.cholesky_solve(r, Lc)
    return (-0.5 * ((r * a).sum() + 2 * torch.log(torch.diagonal(Lc)).sum() + x.shape[0] * math.log(2 * math.pi))) / x.shape[0]


def test_snelson_training_matches_exact_gp(sg):
    """tests/train_snelson.py:79-96 of the reference: 100 Adam(lr=0.1) steps of Simplex-GP (RBFLattice order 1) and of
    an exact RBF GP on Snelson-1D; the final per-datum MLLs must agree to 0.1."""
    from simplex_gp_b200 import gp
    x, y = _snelson()
    x, y = x.cuda(), y.cuda()
    model = gp.ExactGPModel(x, y, sg.RBFLattice(order=1).cuda()).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=0.1)
    for _ in range(100):
        opt.zero_grad()
        loss = -model.mll()
        loss.backward()
        opt.step()
    sgp_mll = -float(loss)
    raw = torch.zeros(4, device="cuda", requires_grad=True)
    opt = torch.optim.Adam([raw], lr=0.1)
    for _ in range(100):
        opt.zero_grad()
        loss = -_exact_rbf_mll(x, y, raw)
        loss.backward()
        opt.step()
    exact_mll = -float(loss)
    assert np.isfinite(sgp_mll) and abs(sgp_mll - exact_mll) < 0.1, (sgp_mll, exact_mll)


class _OracleFilter(torch.autograd.Function):
    """The reference's autograd operator (bilateral_kernel.py:76-124) with the CPU oracle as its native ``filter``:
    forward = filter(src, ref, coeffs); backward = one filter of [g | g(x)x | v | v(x)x] with the derivative stencil
    and the contraction of :122."""

    @staticmethod
    def forward(ctx, source, reference, oracle, coeffs, deriv_coeffs):
        ctx.save_for_backward(source, reference)
        ctx.oracle, ctx.coeffs, ctx.deriv = oracle, coeffs, deriv_coeffs
        return torch.from_numpy(oracle.filter(source.detach().numpy(), reference.detach().numpy(), coeffs))

    @staticmethod
    def backward(ctx, g):
        src, ref = ctx.saved_tensors
        f = lambda a, c: torch.from_numpy(ctx.oracle.filter(a.contiguous().numpy(), ref.detach().numpy(), c))
        n, L = src.shape
        d = ref.shape[1]
        grad_source = grad_reference = None
        if ctx.needs_input_grad[0] and not ctx.needs_input_grad[1]:
            grad_source = f(g, ctx.coeffs)
        if ctx.needs_input_grad[1]:
            gf = g[..., None] * ref[..., None, :]
            sf = src[..., None] * ref[..., None, :]
            all_ = torch.cat([g, gf.reshape(n, L * d), src, sf.reshape(n, L * d)], dim=-1)
            wg, wgf, ws, wsf = torch.split(f(all_, ctx.deriv), [L, L * d, L, L * d], dim=-1)
            grad_reference = -2 * (sf * wg[..., None] - src[..., None] * wgf.view(-1, L, d) + gf * ws[..., None]
                                   - g[..., None] * wsf.view(-1, L, d)).sum(-2)
            if ctx.needs_input_grad[0]:
                grad_source = wg
        return grad_source, grad_reference, None, None, None


def test_cg_training_step_gradients(sg, oracle):
    """A CG training step (config 3's structure: y + 10 probes = 11 RHS, ARD lengthscales) at a size the CPU oracle
    finishes in seconds: value and every hyper-parameter gradient of the CG / stochastic-trace surrogate over the CUDA
    operator equal those of the SAME solver over the reference's autograd formulas with the CPU oracle filter
    (tolerance: 1e-4 of each gradient's norm -- the solves are converged to 1e-6 on both sides, the filters agree to
    1e-7).  d = 5 rather than 18: at d = 18 and this N no two points share a lattice point and the lengthscale gradient
    is 1e-9, i.e. rounding noise."""
    from simplex_gp_b200 import gp
    N, d = 2000, 5
    x, v = make_inputs(N, d, 1, seed=21)
    y = torch.tanh(x[:, 0]) + 0.1 * v[:, 0]
    probes = torch.randn(N, 10, generator=torch.Generator().manual_seed(9)).sign()
    k = sg.RBFLattice(ard_num_dims=d, order=1).cuda()
    model = gp.ExactGPModel(x.cuda(), y.cuda(), k, max_cholesky_size=0).cuda()
    value, surrogate = model.mll(probes=probes, tol=1e-6, max_iter=500)
    (-surrogate).backward()
    got = {"ls": k.raw_lengthscale.grad.flatten().cpu().double(), "noise": model.raw_noise.grad.cpu().double(),
           "scale": model.raw_outputscale.grad.cpu().double(), "mean": model.raw_mean.grad.cpu().double()}
    # the same step on the CPU: same parameterisation (softplus of raw values initialised to 0), oracle-backed operator
    sp = torch.nn.functional.softplus
    raw_ls = torch.zeros(1, d, requires_grad=True)
    raw_s, raw_n, raw_mu = (torch.zeros((), requires_grad=True) for _ in range(3))
    coeffs = k.dkernel_fn.get_coeffs().numpy()
    deriv = k.dkernel_fn.get_deriv_coeffs().numpy()
    xs = x / sp(raw_ls)
    matmul = lambda V: _OracleFilter.apply(V, xs, oracle, coeffs, deriv)
    want_value, want_surr = gp.mll_cg(matmul, y, raw_mu, sp(raw_s), sp(raw_n) + model.min_noise, probes=probes, tol=1e-6,
                                      max_iter=500)
    (-want_surr).backward()
    want = {"ls": raw_ls.grad.flatten().double(), "noise": raw_n.grad.double(), "scale": raw_s.grad.double(),
            "mean": raw_mu.grad.double()}
    assert abs(value - want_value) <= 1e-5 * abs(want_value), (value, want_value)
    assert float(want["ls"].abs().min()) > 1e-4      # a real signal in every ARD dimension
    for name in got:
        err = float((got[name] - want[name]).norm())
        ref = float(want[name].norm())
        assert err <= 1e-4 * ref + 1e-7, (name, got[name], want[name])


def test_predict_mean_and_variance_against_dense_algebra(sg):
    """ExactGPModel.predict (CG path) against the same posterior written with dense matrices made from the operator:
    mean = mu + s K(test, train) A^-1 (y - mu), var = s - s^2 k^T A^-1 k, A = s K + noise I."""
    from simplex_gp_b200 import gp
    torch.manual_seed(5)
    n, nt, d = 1200, 40, 2
    x = torch.rand(n, d, device="cuda") * 4 - 2
    y = torch.sin(2 * x[:, 0]) * torch.cos(x[:, 1]) + 0.05 * torch.randn(n, device="cuda")
    xt = torch.rand(nt, d, device="cuda") * 4 - 2
    kernel = sg.RBFLattice(ard_num_dims=d, order=2).cuda()
    with torch.no_grad():
        kernel.raw_lengthscale.fill_(-0.5)
    model = gp.ExactGPModel(x, y, kernel, max_cholesky_size=0).cuda()
    with torch.no_grad():
        model.raw_noise.fill_(-1.0)    # noise 0.31: keeps s K + noise I well conditioned (the lattice operator is not
                                       # exactly positive semi-definite), so CG and the dense solve agree
    mean, var = model.predict(xt, tol=1e-6, max_iter=2000, variance=True, var_block=16)
    with torch.no_grad():
        s, noise, mu = model.outputscale.double(), model.noise.double(), model.raw_mean.double()
        ls = kernel.lengthscale.detach()
        union = torch.cat([x, xt]) / ls
        W = sg.Lattice(union, kernel.dkernel_fn.get_coeffs()).mvm(torch.eye(n + nt, device="cuda"), exact=True).double()
        Kst, Kts = W[n:, :n], W[:n, n:]
        # the training block of the union filter differs from the filter on the training lattice alone (the test points
        # add lattice points that the blur passes through), so alpha comes from the model's own operator
        Ktr = kernel(x).matmul(torch.eye(n, device="cuda")).double()
        asym = float((Ktr - Ktr.T).norm() / Ktr.norm())
        A = s * Ktr + noise * torch.eye(n, device="cuda", dtype=torch.float64)
        alpha = torch.linalg.solve(A, (y.double() - mu))
        want_mean = mu + s * (Kst @ alpha)
        want_var = s - (s * Kts * torch.linalg.solve(A, s * Kts)).sum(0)
    assert float((mean.double() - want_mean).norm() / want_mean.norm()) < 1e-3, asym
    assert float((var.double() - want_var.clamp_min(0)).abs().max()) < 2e-3
    # and it predicts: far better than the constant predictor
    yt = torch.sin(2 * xt[:, 0]) * torch.cos(xt[:, 1])
    assert float((mean - yt).pow(2).mean().sqrt()) < 0.5 * float(yt.std())
    # the small-N (Cholesky) path gives the same mean
    model.max_cholesky_size = 5000
    mean_dense = model.predict(xt)
    assert float((mean_dense - mean).norm() / mean.norm()) < 1e-3


def test_experiment_harness_learns_on_toy_shape(sg):
    """experiments/train_simplexgp.py end to end (train steps by CG + SLQ, prediction through the extended lattice,
    early-stopping bookkeeping) on the small synthetic problem."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("train_simplexgp", os.path.join(root, "experiments", "train_simplexgp.py"))
    h = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(h)
    summary = h.main(dataset="toy", epochs=8, seed=0, nu=1.5, order=1, min_noise=0.1, quiet=True, variance_points=32)
    assert summary["val/best_step"] >= 1
    assert summary["test/best_rmse"] < 0.75          # targets are standardised: the constant predictor scores 1.0
    assert np.isfinite(summary["test/best_nll"])
    assert len(summary["lengthscale"]) == 3


@pytest.mark.parametrize("n,L", [(5000, 11), (777, 1), (3000, 16), (1200, 37), (40, 256), (2001, 12), (13, 8), (9000, 4)])
def test_cuda_cg_sweeps_match_tensor_expressions(sg, n, L):
    """batched_cg on the sweeps of csrc/sgp_solver.cu against the same iteration written with tensor expressions, on a
    dense SPD operator (so that only the vector updates differ)."""
    from simplex_gp_b200 import gp
    g = torch.Generator().manual_seed(n + L)
    Q = torch.randn(n, 64, generator=g).cuda()
    Kmat = Q @ Q.T / 64
    B = torch.randn(n, L, generator=g).cuda()
    s, noise = torch.tensor(0.7, device="cuda"), torch.tensor(0.3, device="cuda")
    matmul = lambda V: Kmat @ V
    A = lambda V: s * matmul(V) + noise * V
    X0, a0, b0 = gp.batched_cg(A, B, tol=1e-5, max_iter=200)
    X1, a1, b1 = gp.batched_cg(A, B, tol=1e-5, max_iter=200, matmul=matmul, scale=s, shift=noise)
    assert 0 <= a1.shape[0] - a0.shape[0] <= 2 or abs(a0.shape[0] - a1.shape[0]) <= 2   # the flag is read one iteration late
    k = min(a0.shape[0], a1.shape[0]) - 1
    assert float((X0 - X1).norm() / X0.norm()) < 1e-4
    early = min(max(k, 1), 4)      # the coefficient histories of fp32 CG drift apart in the late iterations
    torch.testing.assert_close(a1[:early], a0[:early], rtol=2e-3, atol=1e-6)
    torch.testing.assert_close(b1[:early], b0[:early], rtol=2e-3, atol=1e-6)
    torch.testing.assert_close(a1[: max(k, 1)], a0[: max(k, 1)], rtol=0.1, atol=1e-3)
    resid = (A(X1) - B).norm(dim=0) / B.norm(dim=0)
    assert float(resid.max()) < 5e-5
    # a single iteration (max_iter = 1) and an immediately converged system
    X2, a2, b2 = gp.batched_cg(A, B, tol=1e-5, max_iter=1, matmul=matmul, scale=s, shift=noise)
    assert a2.shape == (1, L)
    X3, a3, _ = gp.batched_cg(A, B, tol=1e9, max_iter=50, matmul=matmul, scale=s, shift=noise)
    assert a3.shape[0] <= 2


def test_cg_on_the_lattice_operator_uses_the_lattice_directly(sg):
    """mll_cg through the kernel module: the CUDA sweeps drive the cached Lattice directly and agree with the tensor-expression iteration over the autograd operator."""
    from simplex_gp_b200 import gp
    x, _ = make_inputs(6000, 3, 1, seed=12)
    xd = x.cuda()
    y = torch.sin(xd[:, 0]) + 0.1 * torch.randn(6000, device="cuda")
    kernel = sg.RBFLattice(ard_num_dims=3, order=1).cuda()
    op = kernel(xd)
    assert op.lattice() is not None
    mu, s, noise = torch.tensor(0.1, device="cuda"), torch.tensor(0.9, device="cuda"), torch.tensor(0.2, device="cuda")
    probes = torch.randn(6000, 10, generator=torch.Generator().manual_seed(3)).sign().cuda()
    with torch.no_grad():
        B = torch.cat([(y - mu).unsqueeze(-1), probes], 1)
        A = lambda V: s * op.matmul(V) + noise * V
        X0, a0, _ = gp.batched_cg(A, B, tol=1e-4, max_iter=200)
        X1, a1, _ = gp.batched_cg(A, B, tol=1e-4, max_iter=200, matmul=op.matmul, scale=s, shift=noise)
    # ~120 iterations on this system: the count moves by a few with the rounding of the dot products
    assert a1.shape[0] > 3 and abs(a0.shape[0] - a1.shape[0]) <= max(2, a0.shape[0] // 10)
    assert float((X0 - X1).norm() / X0.norm()) < 1e-3
    v0, _ = gp.mll_cg(lambda V: op.matmul(V), y, mu, s, noise, probes=probes, tol=1e-4)   # a plain function: no fast path
    v1, _ = gp.mll_cg(op.matmul, y, mu, s, noise, probes=probes, tol=1e-4)
    assert abs(v0 - v1) < 2e-3 * max(abs(v0), 1.0)


def test_pivoted_cholesky_preconditioner_over_the_lattice_operator(sg):
    """The rank-100 pivoted-Cholesky preconditioner of the reference's solver settings
    (experiments/train_simplexgp.py:34-37,63-67) over the CUDA lattice operator: same solution as plain CG at the
    reference's evaluation tolerance, in fewer iterations; its rows come from 16-column MVMs with one-hot blocks."""
    from simplex_gp_b200 import gp
    torch.manual_seed(11)
    n, d = 30000, 2
    x = (torch.rand(n, d, device="cuda") * 6 - 3)
    y = torch.sin(2 * x[:, 0]) * torch.cos(x[:, 1]) + 0.05 * torch.randn(n, device="cuda")
    lat = sg.Lattice((x / 0.7).contiguous(), RBF1)
    s, noise = torch.tensor(1.5, device="cuda"), torch.tensor(0.1, device="cuda")
    mm = lambda V: lat.mvm(V.contiguous())
    Lm = gp.pivoted_cholesky(lambda V: s * mm(V), n, float(s), rank=100, device=x.device)
    assert Lm.shape == (n, 100) and torch.isfinite(Lm).all() and int((Lm.abs().sum(0) > 0).sum()) > 50
    B = torch.cat([y[:, None], torch.randn(n, 3, device="cuda")], 1)
    tol = 1e-2      # the reference's eval_cg_tolerance
    X0, a0, _ = gp.batched_cg(lambda V: s * mm(V) + noise * V, B, tol=tol, max_iter=1500, matmul=mm, scale=s, shift=noise)
    pre = gp.LowRankPreconditioner(Lm, float(noise))
    At = lambda V: pre.inv_sqrt(s * mm(pre.inv_sqrt(V)) + noise * pre.inv_sqrt(V))
    one, zero = torch.ones((), device="cuda"), torch.zeros((), device="cuda")
    Xt, a1, _ = gp.batched_cg(At, pre.inv_sqrt(B), tol=tol, max_iter=3000, matmul=At, scale=one, shift=zero)
    X1 = pre.inv_sqrt(Xt)
    # measured on B200: 140 preconditioned iterations; plain CG has not reached 1e-2 after 3000 on this problem
    assert a1.shape[0] < 0.5 * a0.shape[0] and a1.shape[0] < 1000, (a1.shape[0], a0.shape[0])
    res = lambda X: float(((s * mm(X) + noise * X) - B).norm() / B.norm())
    assert res(X1) < 5 * tol and res(X1) <= res(X0) + tol, (res(X0), res(X1))
    # ... and through the model API (training value + prediction)
    k = sg.RBFLattice(ard_num_dims=d, order=1).cuda()
    model = gp.ExactGPModel(x, y, k, max_cholesky_size=0).cuda()
    probes = torch.randn(n, 10, device="cuda").sign()
    st0, st1 = {}, {}
    v0, _ = model.mll(probes=probes, tol=1e-2, max_iter=1000, stats=st0)
    v1, sur = model.mll(probes=probes, tol=1e-2, max_iter=1000, preconditioner_size=100, stats=st1)
    assert abs(v0 - v1) < 0.1 * abs(v0) + 0.1 and st1["preconditioner_rank"] > 50, (v0, v1, st0, st1)
    assert st1["cg_iterations"] <= st0["cg_iterations"]
    (-sur).backward()
    assert torch.isfinite(k.raw_lengthscale.grad).all()
    xt = torch.rand(50, d, device="cuda") * 4 - 2
    m0 = model.predict(xt, preconditioner_size=0)
    it0 = list(model.last_solve_iterations)
    m1 = model.predict(xt, preconditioner_size=100)
    assert float((m0 - m1).abs().max()) < 0.1 and model.last_solve_iterations[0] <= it0[0], (it0, model.last_solve_iterations)


@pytest.mark.parametrize("stop,min_iter", [("mean", 0), ("mean", 20), ("all", 7)])
def test_cuda_cg_stopping_rules_match_tensor_expressions(sg, stop, min_iter):
    """The stopping rule of GPyTorch's linear_cg (mean relative residual over the non-zero columns, a minimum number of
    iterations) on the device sweeps against the tensor-expression iteration: same iteration count (the device flag is
    read one iteration late) and the same solution.  One right-hand side is zero and one converges much later than the
    others, so 'mean' and 'all' stop at different iterations."""
    from simplex_gp_b200 import gp
    n, L = 1500, 8
    g = torch.Generator().manual_seed(5)
    Q = torch.randn(n, 48, generator=g).cuda()
    Kmat = Q @ Q.T / 48
    B = torch.randn(n, L, generator=g).cuda()
    B[:, 3] = 0.0
    B[:, 5] = Q[:, 0] * 30.0        # lies in the span of the large eigenvalues: slow to converge relative to its norm
    s, noise = torch.tensor(1.0, device="cuda"), torch.tensor(0.05, device="cuda")
    matmul = lambda V: Kmat @ V
    A = lambda V: s * matmul(V) + noise * V
    tol = 0.05
    X0, a0, _ = gp.batched_cg(A, B, tol=tol, max_iter=300, stop=stop, min_iter=min_iter)
    X1, a1, _ = gp.batched_cg(A, B, tol=tol, max_iter=300, matmul=matmul, scale=s, shift=noise, stop=stop, min_iter=min_iter)
    assert a0.shape[0] >= min_iter and a1.shape[0] >= min_iter
    assert 0 <= a1.shape[0] - a0.shape[0] <= 1
    assert float((X0 - X1).norm() / X0.norm()) < 2e-2          # one iteration apart at a loose tolerance
    assert float(X1[:, 3].abs().max()) == 0.0
    if stop == "mean" and min_iter == 0:
        Xa, aa, _ = gp.batched_cg(A, B, tol=tol, max_iter=300, matmul=matmul, scale=s, shift=noise, stop="all")
        assert aa.shape[0] >= a1.shape[0]
    with pytest.raises(ValueError):
        gp.batched_cg(A, B, stop="median")


@pytest.mark.parametrize("N,d,L", [(6000, 6, 16), (5000, 8, 12), (3000, 5, 8), (2500, 4, 11), (1200, 3, 1), (4100, 7, 32)])
def test_cg_sweep_folded_into_the_slice(sg, N, d, L):
    """Lattice.mvm(cg=...) -- AP = s K P + noise P and pAp = sum_n P * AP, the sweep in the ring slice's epilogue where
    that kernel applies (L >= 12), as its own launch elsewhere -- against the product followed by the tensor expressions."""
    import os
    from simplex_gp_b200 import _capi
    g = torch.Generator().manual_seed(N + L)
    x = torch.randn(N, d, generator=g).cuda()
    P = torch.randn(N, L, generator=g).cuda()
    lat = sg.Lattice(x, [0.34608543, 1.0, 0.34608543])
    s, noise = torch.tensor([0.7], device="cuda"), torch.tensor([0.3], device="cuda")
    KP = lat.mvm(P)
    want = 0.7 * KP + 0.3 * P
    want_dot = (P.double() * want.double()).sum(0)
    lib = _capi.lib()
    scratch = torch.empty(int(lib.sgp_cg_scratch_floats(L)), device="cuda")
    for force in ("0", "1"):
        os.environ["SGP_RING_FORCE"] = force           # 1: the epilogue form also on narrow rows
        try:
            AP = torch.full((N, L), float("nan"), device="cuda")
            pAp = torch.full((L,), float("nan"), device="cuda")
            lat.mvm(P, out=AP, cg=(s, noise, pAp, scratch))
            torch.cuda.synchronize()
        finally:
            os.environ.pop("SGP_RING_FORCE")
        assert float((AP - want).norm() / want.norm()) < 1e-6
        assert float(((pAp.double() - want_dot) / want_dot.abs().clamp_min(1e-12)).abs().max()) < 1e-4
    # plain chain (atomic splat, per-axis blur): the sweep as its own launch
    AP = torch.empty((N, L), device="cuda")
    pAp = torch.empty((L,), device="cuda")
    lat.mvm(P, out=AP, mode=1, blur="axis", cg=(s, noise, pAp, scratch))
    assert float((AP - want).norm() / want.norm()) < 1e-5
    assert float(((pAp.double() - want_dot) / want_dot.abs().clamp_min(1e-12)).abs().max()) < 1e-4

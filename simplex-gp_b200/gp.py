"""Marginal log-likelihood of an exact GP whose kernel matrix is the lattice operator -- the callers either side
of the MVM, restated in plain torch because GPyTorch is not available here.

The reference delegates this to GPyTorch (``ExactMarginalLogLikelihood`` -> ``inv_quad_logdet``: dense Cholesky for
N <= 800, otherwise preconditioned CG + stochastic Lanczos quadrature; tests/train_snelson.py:48-61,
experiments/train_simplexgp.py:29-57).  Two routines with the same split:

* ``mll_dense``  -- N <= ``max_cholesky_size``: ``K = s * Op @ I + noise * I`` through ONE lattice filter with L = N
  (the Snelson path), Cholesky, exact MLL.  Gradients flow through ``LatticeFilterGeneral``.
* ``mll_cg``     -- batched conjugate gradients on ``[y - mu | probes]`` (one lattice filter with L = 1 + n_probes per
  iteration), the log-determinant from the CG/Lanczos tridiagonals, and the standard stochastic-trace surrogate for
  the gradient (one more filter forward + one backward).

``mll`` values are per datum, like GPyTorch's ``ExactMarginalLogLikelihood``.  Both routines take the operator as a
callable ``matmul(V) -> K_base @ V`` so the very same solver runs over this package's CUDA filter and over the
reference's CPU filter in the parity tests; the only difference is then the MVM backend.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch

__all__ = ["mll_dense", "mll_cg", "mll_cg_sharded", "batched_cg", "ExactGPModel", "pivoted_cholesky",
           "LowRankPreconditioner"]


def mll_dense(matmul: Callable, y: torch.Tensor, mean: torch.Tensor, outputscale: torch.Tensor,
              noise: torch.Tensor) -> torch.Tensor:
    n = y.shape[0]
    eye = torch.eye(n, dtype=y.dtype, device=y.device)
    K = outputscale * matmul(eye) + noise * eye
    K = 0.5 * (K + K.transpose(-1, -2))   # the lattice operator is symmetric only to rounding
    Lc = torch.linalg.cholesky(K)
    r = (y - mean).unsqueeze(-1)
    alpha = torch.cholesky_solve(r, Lc)
    quad = (r * alpha).sum()
    logdet = 2.0 * torch.log(torch.diagonal(Lc)).sum()
    return (-0.5 * (quad + logdet + n * math.log(2 * math.pi))) / n


@torch.no_grad()
def batched_cg(A: Callable, B: torch.Tensor, tol: float = 1e-4, max_iter: int = 500, *, matmul: Optional[Callable] = None,
               scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None, stop: str = "all",
               min_iter: int = 0):
    """Solve ``A X = B`` column-wise (A symmetric positive definite, given as ``A(V)``).  Returns ``X`` and the CG
    coefficients ``(alphas[k, L], betas[k, L])`` from which the Lanczos tridiagonal of every column follows.

    When the operator is also given in parts, ``A(V) = scale * matmul(V) + shift * V``, and ``B`` is a float32 CUDA
    block of at most 256 columns, the three vector sweeps of an iteration run as one launch each with their dot products
    folded in (``csrc/sgp_solver.cu``) instead of ~17 tensor expressions -- at N = 1M, L = 11 those cost more device
    time than the lattice MVM between them.

    ``stop``: ``"all"`` -- every column's relative residual is below ``tol``; ``"mean"`` -- their mean over the non-zero
    columns is, the rule of GPyTorch's ``linear_cg`` (``residual_norm.mean() < tolerance``), which is what the
    reference's ``cg_tolerance`` settings mean (experiments/train_simplexgp.py:34-37,63-67).  ``min_iter``: iterations
    run whatever the residual (GPyTorch: 10, and ``max_lanczos_quadrature_iterations`` = 20 when the coefficients feed
    a log-determinant)."""
    if stop not in ("all", "mean"):
        raise ValueError(f"stop must be 'all' or 'mean', got {stop!r}")
    min_iter = min(int(min_iter), max(int(max_iter) - 1, 0))
    if (matmul is not None and scale is not None and shift is not None and B.is_cuda and B.dtype == torch.float32
            and B.dim() == 2 and 1 <= B.shape[1] <= 256 and B.shape[0] >= 1):
        return _batched_cg_cuda(matmul, scale, shift, B, tol, max_iter, stop, min_iter)
    X = torch.zeros_like(B)
    R = B.clone()
    P = R.clone()
    rs = (R * R).sum(0)
    b_norm = rs.sqrt().clamp_min(1e-30)
    alphas, betas = [], []
    nonzero = rs > 0
    for it in range(max_iter):
        AP = A(P)
        pAp = (P * AP).sum(0)
        alpha = rs / pAp.clamp_min(1e-30)
        X += P * alpha
        R -= AP * alpha
        rs_new = (R * R).sum(0)
        beta = rs_new / rs.clamp_min(1e-30)
        alphas.append(alpha)
        betas.append(beta)
        rel = rs_new.sqrt() / b_norm
        converged = bool((rel < tol).all()) if stop == "all" else (not bool(nonzero.any()) or bool(rel[nonzero].mean() < tol))
        if converged and it + 1 >= min_iter:
            break
        P = R + P * beta
        rs = rs_new
    return X, torch.stack(alphas), torch.stack(betas)


def _batched_cg_cuda(matmul: Callable, scale: torch.Tensor, shift: torch.Tensor, B: torch.Tensor, tol: float,
                     max_iter: int, stop: str = "all", min_iter: int = 0):
    """The same iteration as above on the sweeps of ``csrc/sgp_solver.cu`` (``sgp_cg_apply / update / direction``).

    * A lattice operator is driven through its ``Lattice`` directly (no autograd node per product, output written in
      place), on blocks padded to a multiple of four columns: the 11-column block of a training step ([y - mu | 10
      probes]) then moves 16-byte vectors end to end (246 -> 175 us per MVM at N = 1M).  Padding columns are all-zero
      right-hand sides; they stay zero.
    * The convergence flag of iteration i is read while iteration i+1 runs, so the device never waits for the host to
      enqueue the next iteration; the price is one iteration more than strictly needed."""
    from . import _capi
    from .lattice import _ptr, _stream_ptr
    lib = _capi.lib()
    dev = B.device
    N, L0 = int(B.shape[0]), int(B.shape[1])
    owner = getattr(matmul, "__self__", None)
    lat = owner.lattice() if hasattr(owner, "lattice") else None
    if lat is not None and (lat.N != N or lat.device != dev):
        lat = None
    # the operator's own stencil, passed to every product (the lattice cache keys on its length and variance only)
    op_coeffs = owner.dkernel.get_coeffs() if (lat is not None and getattr(owner, "dkernel", None) is not None) else None
    L = (L0 + 3) // 4 * 4 if (lat is not None and L0 > 4) else L0
    if L > 256:
        L = L0
    X = torch.zeros((N, L), dtype=torch.float32, device=dev)
    R = torch.zeros((N, L), dtype=torch.float32, device=dev)
    R[:, :L0] = B
    P = R.clone()
    rs = (R * R).sum(0).contiguous()
    bnorm = rs.sqrt().clamp_min(1e-30).contiguous()
    pAp = torch.empty(L, dtype=torch.float32, device=dev)
    max_iter = max(1, int(max_iter))
    alphas = torch.zeros((max_iter, L), dtype=torch.float32, device=dev)
    betas = torch.zeros((max_iter, L), dtype=torch.float32, device=dev)
    done = torch.zeros(max_iter, dtype=torch.int32, device=dev)
    done_host = torch.zeros(max_iter, dtype=torch.int32).pin_memory()
    events = [torch.cuda.Event(), torch.cuda.Event()]
    scratch = torch.empty(int(lib.sgp_cg_scratch_floats(L)), dtype=torch.float32, device=dev)
    s = scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
    nz = shift.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
    AP = torch.empty((N, L), dtype=torch.float32, device=dev) if lat is not None else None
    k = 0
    criterion = 1 if stop == "mean" else 0   # SGP_CG_MEAN / SGP_CG_ALL_COLUMNS
    import os
    fuse = os.environ.get("SGP_CG_FUSE", "1") != "0"   # 0: the sweep after the product as its own launch (experiments)
    one_call = None
    if lat is not None and fuse and (L % 4 == 0 or L <= 4):
        one_call = _cg_one_call_setup(lat, op_coeffs, L)
    if one_call is not None:
        # the whole iteration (product, its epilogue sweep, update, direction) enqueued by ONE C call: issued piece by
        # piece from Python an iteration costs ~500 us of host time at N = 1M, more than the ~290 us the device needs
        import ctypes as C
        v_out, ent, seg, n_ent, arr, coeffs_np, bufs = one_call
        args_head = (C.byref(v_out), _ptr(ent), _ptr(seg), n_ent, arr, len(arr), _capi_fp(coeffs_np), coeffs_np.shape[0],
                     _ptr(bufs[0]), _ptr(bufs[1]), 3, _ptr(X), _ptr(R), _ptr(P), _ptr(AP), _ptr(rs), _ptr(pAp), _ptr(bnorm),
                     _ptr(s), _ptr(nz), float(tol), criterion, L, int(bufs[0].shape[1]), _ptr(alphas), _ptr(betas), _ptr(done),
                     C.c_void_p(done_host.data_ptr()))
        flags_np = done_host.numpy()
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            for it in range(max_iter):
                _capi.check(lib.sgp_cg_iteration(*args_head, it, _ptr(scratch), st))
                events[it & 1].record()
                k = it + 1
                if it > 0:
                    events[(it - 1) & 1].synchronize()      # iteration it-1 finished long ago; iteration it is running
                    if int(flags_np[it - 1]) and it >= min_iter:
                        break
        return X[:, :L0].contiguous(), alphas[:k, :L0].contiguous(), betas[:k, :L0].contiguous()
    with torch.cuda.device(dev):
        st = _stream_ptr(dev)
        for it in range(max_iter):
            if lat is not None:
                # the product and the sweep after it (AP = s K P + noise P, pAp = sum P * AP) in one call: the sweep runs
                # in the slice's epilogue where the TMA-ring slice applies
                if fuse:
                    lat.mvm(P, out=AP, coeffs=op_coeffs, cg=(s, nz, pAp, scratch))
                else:
                    lat.mvm(P, out=AP, coeffs=op_coeffs)
                    _capi.check(lib.sgp_cg_apply(_ptr(AP), _ptr(P), _ptr(s), _ptr(nz), N, L, _ptr(pAp), _ptr(scratch), st))
            else:
                AP = matmul(P)
                if AP.dtype != torch.float32 or not AP.is_contiguous() or AP.data_ptr() == P.data_ptr():
                    AP = AP.to(torch.float32).contiguous().clone()
                _capi.check(lib.sgp_cg_apply(_ptr(AP), _ptr(P), _ptr(s), _ptr(nz), N, L, _ptr(pAp), _ptr(scratch), st))
            # X += alpha P is deferred into the direction sweep (which reads P anyway): 8 passes over [N, L] per iteration
            # instead of 9; when the loop stops after an update, the last alpha P is added here
            _capi.check(lib.sgp_cg_update_r(_ptr(R), _ptr(AP), _ptr(rs), _ptr(pAp), _ptr(bnorm), float(tol), criterion,
                                            N, L, _ptr(alphas[it]), _ptr(betas[it]), _ptr(done[it:]), _ptr(scratch), st))
            done_host[it:it + 1].copy_(done[it:it + 1], non_blocking=True)
            events[it & 1].record()
            k = it + 1
            stop_now = it + 1 >= max_iter
            if it > 0 and not stop_now:
                events[(it - 1) & 1].synchronize()      # iteration it-1 finished long ago; iteration it is running
                stop_now = bool(int(done_host[it - 1])) and it >= min_iter   # (iteration it - 1 is the it-th one)
            if stop_now:
                X.addcmul_(P, alphas[it].unsqueeze(0))
                break
            _capi.check(lib.sgp_cg_direction_x(_ptr(P), _ptr(R), _ptr(X), _ptr(alphas[it]), _ptr(betas[it]), N, L, st))
    return X[:, :L0].contiguous(), alphas[:k, :L0].contiguous(), betas[:k, :L0].contiguous()


def _capi_fp(a):
    from .lattice import _fp
    return _fp(a)


def _cg_one_call_setup(lat, op_coeffs, L: int):
    """Tables and private work buffers for ``sgp_cg_iteration`` on ``lat`` (production chain), or ``None`` when the
    lattice has no row-sorted entries / blur groups."""
    from .lattice import _coeffs_np
    lazy = getattr(lat, "_lazy", None)
    if lazy is not None:            # many products follow: build the postponed tables now
        lazy["calls"] = lazy["after"]
        lat._count_product()
    if lat.rows is None or lat.groups is None or lat.N == 0 or lat.M == 0:
        return None
    c = lat.coeffs if op_coeffs is None else _coeffs_np(op_coeffs)
    if c.shape[0] != 2 * lat.order + 1:
        return None
    Lv = lat.lattice_width(L)
    v_out = lat._slice_view(Lv, True, lat.exact)
    bufs = (torch.zeros((lat.M, Lv), dtype=torch.float32, device=lat.device),
            torch.empty((lat.M, Lv), dtype=torch.float32, device=lat.device))
    return v_out, lat.rows["ent"], lat.rows["seg_row"], lat.rows["n"], lat.groups["array"], c, bufs


@torch.no_grad()
def pivoted_cholesky(matmul: Callable, n: int, diag_value, rank: int = 100, block: int = 16, rel_tol: float = 1e-3,
                     device=None, dtype=torch.float32) -> torch.Tensor:
    """Rank-``rank`` pivoted Cholesky factor ``Lm [n, rank]`` of ``K = diag_value-scaled operator``: ``K ~ Lm Lm^T``.

    The reference's solver settings ask GPyTorch for this preconditioner (``max_preconditioner_size = 100``,
    experiments/train_simplexgp.py:34-37,63-67).  GPyTorch's loop fetches one row of ``K`` per pivot -- for the lattice
    operator a row is an MVM with a one-hot vector, and a 1-column MVM costs half of a 16-column one -- so the pivots are
    taken ``block`` at a time: the ``block`` largest residual diagonals are chosen, their rows fetched by ONE
    ``block``-column MVM, and the Cholesky updates run over them in order; a pivot whose residual has meanwhile dropped
    under ``rel_tol * diag_value`` (it was close to an earlier pivot of the block) contributes a zero column.  Like
    GPyTorch the residual starts from the operator's nominal diagonal (``LazyTensor.diag()``, ones for the lattice kernel:
    bilateral_kernel.py:139-140, times the output scale).  No host synchronisation inside."""
    dv = float(diag_value)
    d = torch.full((n,), dv, device=device, dtype=dtype)
    Lm = torch.zeros((n, rank), device=device, dtype=dtype)
    k = 0
    while k < rank:
        b = min(block, rank - k)
        piv = torch.topk(d, b).indices
        E = torch.zeros((n, b), device=device, dtype=dtype)
        E[piv, torch.arange(b, device=device)] = 1.0
        R = matmul(E)                                   # rows K[piv, :] as columns (the operator is symmetric)
        for j in range(b):
            p = piv[j]
            dp = d[p]
            ok = (dp > rel_tol * dv).to(dtype)
            col = R[:, j] - Lm[:, :k] @ Lm[p, :k] if k > 0 else R[:, j].clone()
            col = col * (ok / torch.sqrt(dp.clamp_min(1e-30)))
            Lm[:, k] = col
            d = (d - col * col).clamp_min(0.0)
            d[p] = 0.0
            k += 1
    return Lm


class LowRankPreconditioner:
    """``P = noise * I + Lm Lm^T`` through the thin eigen-decomposition ``Lm Lm^T = U diag(lam) U^T`` (``U [n, k]``
    orthonormal): ``P^{-1/2} v = v / sqrt(noise) + U ((lam + noise)^{-1/2} - noise^{-1/2}) U^T v`` -- two skinny GEMMs.
    CG then runs on ``P^{-1/2} A P^{-1/2}`` (symmetric preconditioning), which keeps the plain CG sweeps of
    ``csrc/sgp_solver.cu`` and makes the Lanczos tridiagonals those of the preconditioned operator, so that
    ``log|A| = log|P| + log|P^{-1/2} A P^{-1/2}|`` (GPyTorch corrects its log-determinant estimate the same way)."""

    def __init__(self, Lm: torch.Tensor, noise):
        n, k = Lm.shape
        self.n, self.noise = n, float(noise)
        G = (Lm.T @ Lm).double()
        lam, Q = torch.linalg.eigh(G)
        keep = lam > 1e-10 * lam.max().clamp_min(1e-30)
        lam, Q = lam[keep], Q[:, keep]
        self.lam = lam.to(Lm.dtype)
        self.U = (Lm @ (Q / lam.sqrt()).to(Lm.dtype)).contiguous()
        self.coef = ((self.lam + self.noise).rsqrt() - self.noise ** -0.5)

    def inv_sqrt(self, V: torch.Tensor) -> torch.Tensor:
        return V * (self.noise ** -0.5) + self.U @ (self.coef.unsqueeze(1) * (self.U.T @ V))

    def logdet(self) -> float:
        import math
        return float((self.n - self.lam.numel()) * math.log(self.noise) + torch.log(self.lam.double() + self.noise).sum())


def _lanczos_logdet(alphas: torch.Tensor, betas: torch.Tensor, n: int) -> torch.Tensor:
    """Stochastic Lanczos quadrature: mean over probe columns of ``n * e1^T log(T) e1``."""
    k, p = alphas.shape
    a, b = alphas.double().cpu(), betas.double().cpu()
    # T = tridiag of the Lanczos process that CG runs implicitly: diag_i = 1/a_i + b_{i-1}/a_{i-1}, off_i = sqrt(b_i)/a_i;
    # all probe columns at once
    diag = 1.0 / a
    diag[1:] += b[:-1] / a[:-1]
    off = torch.sqrt(b[:-1]) / a[:-1]
    T = torch.zeros(p, k, k, dtype=torch.float64)
    idx = torch.arange(k)
    T[:, idx, idx] = diag.T
    if k > 1:
        T[:, idx[:-1], idx[1:]] = off.T
        T[:, idx[1:], idx[:-1]] = off.T
    ev, V = torch.linalg.eigh(T)
    total = float((V[:, 0, :] ** 2 * torch.log(ev.clamp_min(1e-30))).sum())
    return torch.tensor(n * total / p)


def mll_cg(matmul: Callable, y: torch.Tensor, mean: torch.Tensor, outputscale: torch.Tensor, noise: torch.Tensor,
           n_probes: int = 10, tol: float = 1e-4, max_iter: int = 500, generator: Optional[torch.Generator] = None,
           probes: Optional[torch.Tensor] = None, preconditioner_size: int = 0, stats: Optional[dict] = None,
           stop: str = "all", min_iter: int = 0):
    """Returns ``(mll_value, surrogate)``: ``mll_value`` is the (detached) per-datum MLL estimate, ``surrogate`` a scalar
    whose gradient with respect to the hyper-parameters is the usual CG / stochastic-trace MLL gradient estimate.

    ``preconditioner_size = k > 0``: rank-``k`` pivoted-Cholesky preconditioner of the reference's solver settings
    (``pivoted_cholesky`` / ``LowRankPreconditioner``).  ``stats`` (a dict) receives the CG iteration count.
    ``stop`` / ``min_iter``: the stopping rule of ``batched_cg``; GPyTorch's own (what the reference trains with) is
    ``stop="mean", min_iter=20``."""
    n = y.shape[0]
    r = (y - mean)
    if probes is None:
        probes = torch.randn(n, n_probes, dtype=y.dtype, device=y.device, generator=generator).sign()   # Rademacher
    Z = probes.to(device=y.device, dtype=y.dtype)
    n_probes = Z.shape[1]
    B = torch.cat([r.detach().unsqueeze(-1), Z], dim=1)

    def A(V):
        with torch.no_grad():
            return outputscale.detach() * matmul(V) + noise.detach() * V

    pre = None
    if preconditioner_size > 0:
        with torch.no_grad():
            s_ = outputscale.detach()
            Lm = pivoted_cholesky(lambda V: s_ * matmul(V), n, float(s_), rank=int(preconditioner_size), device=y.device,
                                  dtype=y.dtype)
            pre = LowRankPreconditioner(Lm, float(noise.detach()))
            del Lm
    if pre is None:
        X, al, be = batched_cg(A, B, tol=tol, max_iter=max_iter, matmul=matmul, scale=outputscale, shift=noise,
                               stop=stop, min_iter=min_iter)
        logdet_p = 0.0
    else:
        # symmetric preconditioning: CG on P^-1/2 A P^-1/2 with right-hand sides [P^-1/2 r | probes]; the probe columns
        # are Rademacher in the preconditioned space (what the Lanczos quadrature needs), their images P^-1/2 z are the
        # vectors the trace estimator pairs the solutions with (E[z~^T A~^-1 P^-1/2 dA P^-1/2 z~] = tr(A^-1 dA))
        Bt = torch.cat([pre.inv_sqrt(B[:, :1]), B[:, 1:]], dim=1)
        one, zero = torch.ones((), device=y.device, dtype=y.dtype), torch.zeros((), device=y.device, dtype=y.dtype)
        At = lambda V: pre.inv_sqrt(A(pre.inv_sqrt(V)))
        Xt, al, be = batched_cg(At, Bt, tol=tol, max_iter=max_iter, matmul=At, scale=one, shift=zero, stop=stop,
                                min_iter=min_iter)
        X = pre.inv_sqrt(Xt)
        Z = pre.inv_sqrt(Z)
        logdet_p = pre.logdet()
    if stats is not None:
        stats["cg_iterations"] = int(al.shape[0])
        stats["preconditioner_rank"] = 0 if pre is None else int(pre.lam.numel())
    alpha, U = X[:, :1], X[:, 1:]
    quad = float((r.detach().unsqueeze(-1) * alpha).sum())
    logdet = logdet_p + float(_lanczos_logdet(al[:, 1:], be[:, 1:], n))
    value = (-0.5 * (quad + logdet + n * math.log(2 * math.pi))) / n
    # surrogate: d/dtheta [ -1/2 r^T K^-1 r - 1/2 log|K| ] = 1/2 a^T dK a - a^T dmu ... - 1/2 E[u^T dK z]
    V = torch.cat([alpha, Z], dim=1)
    KV = outputscale * matmul(V) + noise * V
    s = 0.5 * (alpha[:, 0] * KV[:, 0]).sum() - 0.5 * (U * KV[:, 1:]).sum() / n_probes
    s = s + (alpha[:, 0] * (mean - mean.detach())).sum() if mean.requires_grad else s
    return value, s / n


def mll_cg_sharded(matmul: Callable, y: torch.Tensor, mean: torch.Tensor, outputscale: torch.Tensor,
                   noise: torch.Tensor, probes: torch.Tensor, tol: float = 1e-4, max_iter: int = 500, group=None):
    """``mll_cg`` with the RHS columns ``[y - mu | probes]`` sharded over the ranks of ``group`` (north star: "the
    CG/Lanczos probe and RHS columns are sharded across GPUs").  Every rank holds the same lattice and runs CG on its own
    columns -- columns are independent, so the solve needs no communication at all --, the per-column Lanczos
    coefficients and solutions are all-gathered once at the end, and each rank back-propagates the surrogate of its own
    columns; the caller all-reduces (sums) the hyper-parameter gradients, as data-parallel training does.  ``probes``
    must be identical on every rank.  Returns ``(mll_value, local_surrogate)``."""
    import torch.distributed as dist

    from .distributed import shard_columns

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = y.shape[0]
    r = y - mean
    Z = probes.to(device=y.device, dtype=y.dtype)
    n_probes = Z.shape[1]
    B = torch.cat([r.detach().unsqueeze(-1), Z], dim=1)
    Lt = B.shape[1]
    lo, hi = shard_columns(Lt, world, rank)

    def A(V):
        with torch.no_grad():
            return outputscale.detach() * matmul(V) + noise.detach() * V

    if hi > lo:
        X_loc, al, be = batched_cg(A, B[:, lo:hi].contiguous(), tol=tol, max_iter=max_iter, matmul=matmul,
                                   scale=outputscale, shift=noise)
    else:
        X_loc, al, be = B.new_zeros(n, 0), B.new_zeros(0, 0), B.new_zeros(0, 0)
    # gather: pad the coefficient histories to the longest one (alpha = inf, beta = 0 leave the tridiagonal's leading
    # block untouched -- handled below by truncating each column at its own length)
    iters = torch.tensor([al.shape[0]], device=y.device)
    if world > 1:
        dist.all_reduce(iters, op=dist.ReduceOp.MAX, group=group)
    kmax = int(iters.item())
    wmax = max(h - l for l, h in (shard_columns(Lt, world, q) for q in range(world)))
    pack = B.new_zeros(n + 2 * kmax + 1, wmax)
    w = hi - lo
    if w > 0:
        pack[:n, :w] = X_loc
        pack[n:n + al.shape[0], :w] = al
        pack[n + kmax:n + kmax + be.shape[0], :w] = be
        pack[n + 2 * kmax, :w] = float(al.shape[0])
    parts = [pack]
    if world > 1:
        parts = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(parts, pack, group=group)
    X = torch.cat([p[:n, : (shard_columns(Lt, world, q)[1] - shard_columns(Lt, world, q)[0])] for q, p in enumerate(parts)], 1)
    alpha = X[:, :1]
    quad = float((r.detach().unsqueeze(-1) * alpha).sum())
    # log-determinant from the probe columns' tridiagonals (every rank computes the same number)
    total, cnt = 0.0, 0
    for q, p in enumerate(parts):
        l_q, h_q = shard_columns(Lt, world, q)
        for j in range(h_q - l_q):
            if l_q + j == 0:
                continue   # column 0 is y - mu, not a probe
            k_j = int(p[n + 2 * kmax, j].item())
            total += float(_lanczos_logdet(p[n:n + k_j, j:j + 1], p[n + kmax:n + kmax + k_j, j:j + 1], n))
            cnt += 1
    logdet = total / max(cnt, 1)
    value = (-0.5 * (quad + logdet + n * math.log(2 * math.pi))) / n
    # surrogate over this rank's columns
    s = y.new_zeros(())
    if hi > lo:
        V_loc = torch.cat([alpha, Z], dim=1)[:, lo:hi].contiguous()
        KV = outputscale * matmul(V_loc) + noise * V_loc
        U_loc = X[:, lo:hi]
        for j in range(hi - lo):
            col = lo + j
            if col == 0:
                s = s + 0.5 * (alpha[:, 0] * KV[:, j]).sum()
                if mean.requires_grad:
                    s = s + (alpha[:, 0] * (mean - mean.detach())).sum()
            else:
                s = s - 0.5 * (U_loc[:, j] * KV[:, j]).sum() / n_probes
    return value, s / n


class ExactGPModel(torch.nn.Module):
    """Constant mean + ``outputscale * kernel`` + Gaussian noise: the model of tests/train_snelson.py:11-23 and
    experiments/train_simplexgp.py (``ScaleKernel(RBFLattice(...))``, ``GaussianLikelihood(noise > min_noise)``), with
    GPyTorch's parameterisation (softplus of raw parameters, raw values initialised to 0)."""

    def __init__(self, train_x, train_y, kernel, min_noise: float = 1e-4, max_cholesky_size: int = 800):
        super().__init__()
        self.train_x, self.train_y = train_x, train_y
        self.kernel = kernel
        self.min_noise = min_noise
        self.max_cholesky_size = max_cholesky_size
        self.raw_mean = torch.nn.Parameter(torch.zeros(()))
        self.raw_outputscale = torch.nn.Parameter(torch.zeros(()))
        self.raw_noise = torch.nn.Parameter(torch.zeros(()))

    @property
    def outputscale(self):
        return torch.nn.functional.softplus(self.raw_outputscale)

    @property
    def noise(self):
        return torch.nn.functional.softplus(self.raw_noise) + self.min_noise

    def operator(self):
        return self.kernel(self.train_x)

    def mll(self, **cg_kwargs):
        """Per-datum marginal log-likelihood as a differentiable scalar (dense) or ``(value, surrogate)`` (CG)."""
        op = self.operator()
        matmul = op.matmul if hasattr(op, "matmul") else (lambda V: op @ V)
        if self.train_x.shape[0] <= self.max_cholesky_size:
            return mll_dense(matmul, self.train_y, self.raw_mean, self.outputscale, self.noise)
        return mll_cg(matmul, self.train_y, self.raw_mean, self.outputscale, self.noise, **cg_kwargs)

    @torch.no_grad()
    def predict(self, test_x, tol: float = 1e-2, max_iter: int = 1000, variance: bool = False, var_block: int = 16,
                preconditioner_size: int = 100):
        """Posterior mean (and, on request, the latent variance) at ``test_x`` -- what ``model(test_x)`` of
        experiments/train_simplexgp.py:60-84 returns in evaluation mode.

        ``alpha = (s K + noise I)^-1 (y - mu)`` by a dense factorisation (N <= ``max_cholesky_size``) or CG at relative residual
        ``tol`` (the reference's ``eval_cg_tolerance``); ``mean = mu + s K(test, train) alpha`` is ONE rectangular
        product on the training lattice extended by the test points.  The variance
        ``s - s^2 k_i^T (s K + noise I)^-1 k_i`` is exact up to ``tol`` and costs a solve with ``var_block`` columns per
        ``var_block`` test points (the reference uses GPyTorch's LOVE approximation, ``fast_pred_var``, instead).
        ``preconditioner_size``: rank of the pivoted-Cholesky preconditioner of the CG solves (0 = none)."""
        op = self.operator()
        n = self.train_x.shape[0]
        s, noise, mu = self.outputscale, self.noise, self.raw_mean
        r = (self.train_y - mu).unsqueeze(-1)
        dense = None
        if n <= self.max_cholesky_size:
            # LU of the operator as it is: on a finite lattice the blur passes do not commute exactly, so the filter is
            # symmetric only approximately, and CG (below) solves the unsymmetrised system too
            eye = torch.eye(n, dtype=r.dtype, device=r.device)
            dense = torch.linalg.lu_factor(s * op.matmul(eye) + noise * eye)

        pre = None
        if dense is None and preconditioner_size > 0:
            # the reference's evaluation settings: eval_cg_tolerance 1e-2 with a rank-100 pivoted-Cholesky preconditioner
            # (experiments/train_simplexgp.py:63-67); built once per prediction, shared by every solve below
            pre = LowRankPreconditioner(pivoted_cholesky(lambda V: s * op.matmul(V), n, float(s), rank=preconditioner_size,
                                                         device=r.device, dtype=r.dtype), float(noise))
        self.last_solve_iterations = []

        def solve(B):
            if dense is not None:
                return torch.linalg.lu_solve(*dense, B)
            if pre is None:
                X, al, _ = batched_cg(lambda V: s * op.matmul(V) + noise * V, B, tol=tol, max_iter=max_iter,
                                      matmul=op.matmul, scale=s, shift=noise)
            else:
                At = lambda V: pre.inv_sqrt(s * op.matmul(pre.inv_sqrt(V)) + noise * pre.inv_sqrt(V))
                one, zero = torch.ones((), device=B.device, dtype=B.dtype), torch.zeros((), device=B.device, dtype=B.dtype)
                Xt, al, _ = batched_cg(At, pre.inv_sqrt(B), tol=tol, max_iter=max_iter, matmul=At, scale=one, shift=zero)
                X = pre.inv_sqrt(Xt)
            self.last_solve_iterations.append(int(al.shape[0]))
            return X

        cross = self.kernel(test_x, self.train_x)              # K(test, train)
        mean = mu + s * cross.matmul(solve(r))[:, 0]
        if not variance:
            return mean
        n_test = test_x.shape[0]
        back = cross.transpose(-1, -2)                         # K(train, test)
        var = torch.empty(n_test, dtype=mean.dtype, device=mean.device)
        for lo in range(0, n_test, var_block):
            hi = min(lo + var_block, n_test)
            E = torch.zeros(n_test, hi - lo, dtype=mean.dtype, device=mean.device)
            E[torch.arange(lo, hi), torch.arange(hi - lo)] = 1.0
            Kb = s * back.matmul(E)                            # s k_i as columns [n, block]
            var[lo:hi] = s - (Kb * solve(Kb)).sum(0)
        return mean, var.clamp_min(0.0)

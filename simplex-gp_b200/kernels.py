"""GPyTorch surface of the lattice kernel: ``RBFLattice`` / ``MaternLattice`` / ``BilateralKernel``.

Mirrors gpytorch_lattice_kernel/bilateral_kernel.py:127-160, 183-200, 247-254 of the reference: the same class and
factory names, constructor arguments (``order``, ``nu``, and every ``gpytorch.kernels.Kernel`` keyword such as
``ard_num_dims`` passed through), ``forward`` contract (``diag=True`` -> ones; ``x1 == x2`` -> square operator,
otherwise the rectangular one built on the union lattice) and the fact that the lengthscale enters only as
``x / lengthscale``.

GPyTorch is a third-party dependency of the reference that is not vendored with it.  When it is importable the classes
below derive from ``gpytorch.kernels.Kernel`` and ``gpytorch.lazy.LazyTensor`` (GPyTorch < 1.9) or
``linear_operator.LinearOperator`` (>= 1.9), so ``gpytorch.kernels.ScaleKernel(RBFLattice(ard_num_dims=d))`` works
unchanged.  When it is not (this image has no GPyTorch and no network) a minimal stand-in with the same
``lengthscale`` / ``raw_lengthscale`` semantics (softplus-positive constraint, initial raw value 0) keeps the module
usable and testable; see ``_compat``.
"""
from __future__ import annotations

import torch

from ._compat import HAVE_GPYTORCH, Kernel, LazyTensor
from .coeffs import DiscretizedKernelFN, Matern, rbf
from .function import LatticeFilterGeneral

__all__ = ["SquareLazyLattice", "RectangularLazyLattice", "LatticeAccelerated", "RBFLattice", "MaternLattice",
           "BilateralKernel", "HAVE_GPYTORCH"]


class SquareLazyLattice(LazyTensor):
    """``K(X, X)`` as an operator: ``_matmul(V)`` is one lattice filter (bilateral_kernel.py:127-140)."""

    def __init__(self, x, dkernel=None):
        super().__init__(x, dkernel=dkernel)
        self.x = x
        self.dkernel = dkernel

    def _matmul(self, V):
        return LatticeFilterGeneral.apply(V, self.x, self.dkernel)

    def lattice(self):
        """The cached ``Lattice`` behind this operator (forward stencil), or ``None`` when ``x`` is not a float32 CUDA
        matrix.  Solvers that do not need autograd drive it directly (``lattice().mvm(P, out=AP)``, ``capture``)
        instead of paying for an autograd node per product."""
        x = self.x
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
            return None
        return LatticeFilterGeneral.cache.get(x, self.dkernel.get_coeffs())

    def _size(self):
        return torch.Size((self.x.shape[-2], self.x.shape[-2]))

    def _transpose_nonbatch(self):
        return self   # the operator is symmetric

    def diag(self):
        return torch.ones_like(self.x[..., 0])

    # linear_operator (GPyTorch >= 1.9) spells it this way
    def _diagonal(self):
        return self.diag()


class RectangularLazyLattice(LazyTensor):
    """``K(Xin, Xout)``: filter on the union of both point sets with the ``Xin`` rows of ``V`` zero-padded, keep the
    ``Xin`` rows of the result (bilateral_kernel.py:142-160)."""

    def __init__(self, xin, xout, dkernel=None):
        super().__init__(xin, xout, dkernel=dkernel)
        self.xin = xin
        self.xout = xout
        self.dkernel = dkernel
        self._union = None
        self._partner = None    # set on a transposed view: the operator whose union lattice it shares

    def _union_reference(self):
        """``cat([xout, xin])`` with its lattice made by extending the cached lattice of ``xout`` (the training inputs
        in a prediction) by the rows of ``xin`` -- the reference builds the union lattice from nothing in every product
        (bilateral_kernel.py:150-156).  The tensor is kept so that later products of this operator hit the lattice
        cache by identity; with gradients attached it is rebuilt per call and found by value."""
        if self._union is not None:
            return self._union
        union = torch.cat([self.xout, self.xin], dim=-2)
        if union.is_cuda and union.dtype == torch.float32 and union.dim() == 2 and self.dkernel is not None:
            LatticeFilterGeneral.cache.get_union(union, self.xout.shape[-2], self.dkernel.get_coeffs(),
                                                 base=self.xout)
        if not union.requires_grad:
            self._union = union
        return union

    def _matmul(self, V):
        n_out = self.xout.shape[-2]
        assert V.shape[-2] == n_out, f"mismatched shapes? {V.shape, self.xout.shape}"
        if self._partner is not None:
            # transposed view: the union filter is symmetric, so K(xin, xout) is the other off-diagonal block of the
            # partner's union operator -- same lattice, V on the last rows, the product read from the first
            union = self._partner._union_reference()
            n_in = self.xin.shape[-2]
            padding = V.new_zeros(*V.shape[:-2], n_in, V.shape[-1])
            filtered = LatticeFilterGeneral.apply(torch.cat([padding, V], dim=-2), union, self.dkernel)
            return filtered[..., :n_in, :]
        # one square filter on the union of both point sets: the rows of V sit on xout, the xin rows carry zeros and
        # receive the product
        union = self._union_reference()
        padding = V.new_zeros(*V.shape[:-2], self.xin.shape[-2], V.shape[-1])
        filtered = LatticeFilterGeneral.apply(torch.cat([V, padding], dim=-2), union, self.dkernel)
        return filtered[..., n_out:, :]

    def _size(self):
        return torch.Size((*self.xin.shape[:-1], self.xout.shape[-2]))

    def _transpose_nonbatch(self):
        if self._partner is not None:
            return self._partner
        t = RectangularLazyLattice(self.xout, self.xin, self.dkernel)
        t._partner = self       # shares this operator's union lattice instead of building cat([xin, xout])'s
        return t


class LatticeAccelerated(Kernel):
    """A stationary kernel ``k(|x1 - x2|^2)`` evaluated through the permutohedral lattice (bilateral_kernel.py:183-200)."""

    has_lengthscale = True

    def __init__(self, kernel_fn, *args, order=2, **kwargs):
        super().__init__(*args, **kwargs)
        self.dkernel_fn = DiscretizedKernelFN(kernel_fn, order)

    def forward(self, x1, x2, diag=False, **params):
        if diag is True:
            return torch.ones_like(x1[..., 0])
        if x1.shape == x2.shape and (x1 is x2 or bool(x1.eq(x2).all())):
            return SquareLazyLattice(x1.div(self.lengthscale), self.dkernel_fn)
        return RectangularLazyLattice(x1.div(self.lengthscale), x2.div(self.lengthscale), self.dkernel_fn)


def RBFLattice(*args, order=2, **kwargs):
    return LatticeAccelerated(rbf, *args, order=order, **kwargs)


def BilateralKernel(*args, **kwargs):
    return RBFLattice(*args, **kwargs)


def MaternLattice(*args, nu=1.5, order=3, **kwargs):
    if nu not in (1.5, 2.5):
        raise NotImplementedError(f"nu={nu}: the reference supports 1.5 and 2.5 (bilateral_kernel.py:212-217)")
    return LatticeAccelerated(lambda d2: Matern.apply(d2, nu), *args, order=order, **kwargs)

"""Base classes for ``kernels.py``: GPyTorch's when it is installed, otherwise a minimal stand-in.

The stand-in reproduces only what ``LatticeAccelerated`` and its operators rely on
(``gpytorch.kernels.Kernel``: ``ard_num_dims``, ``batch_shape``, ``active_dims``, ``lengthscale_constraint``,
``raw_lengthscale`` + ``lengthscale`` with the softplus-positive constraint and initial raw value 0, ``__call__``;
``gpytorch.lazy.LazyTensor``: ``matmul``, ``size``/``shape``, ``diag``, ``evaluate``, ``transpose``).  It is not a
re-implementation of GPyTorch.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

HAVE_GPYTORCH = False
Kernel = None
LazyTensor = None

try:  # pragma: no cover - GPyTorch is absent from the build image
    import gpytorch as _gp
    from gpytorch.kernels import Kernel as _GPKernel

    try:
        from gpytorch.lazy import LazyTensor as _GPLazy          # GPyTorch < 1.9 (what the reference imports)
    except Exception:
        from linear_operator.operators import LinearOperator as _GPLazy   # GPyTorch >= 1.9
    Kernel, LazyTensor, HAVE_GPYTORCH = _GPKernel, _GPLazy, True
except Exception:
    pass


class _Positive(nn.Module):
    """softplus constraint (GPyTorch's default ``Positive()`` for lengthscales)."""

    def transform(self, raw):
        return F.softplus(raw)

    def inverse_transform(self, value):
        value = torch.as_tensor(value)
        return value + torch.log(-torch.expm1(-value))


class _StandInKernel(nn.Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, lengthscale_prior=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self.batch_shape = torch.Size(batch_shape)
        self.active_dims = None if active_dims is None else torch.as_tensor(active_dims, dtype=torch.long)
        self.eps = eps
        if lengthscale_prior is not None:
            raise NotImplementedError("priors need GPyTorch")
        if self.has_lengthscale:
            n = 1 if ard_num_dims is None else ard_num_dims
            self.raw_lengthscale = nn.Parameter(torch.zeros(*self.batch_shape, 1, n))
            self.raw_lengthscale_constraint = lengthscale_constraint if lengthscale_constraint is not None else _Positive()

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale) if self.has_lengthscale else None

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_lengthscale.dtype, device=self.raw_lengthscale.device)
        with torch.no_grad():
            self.raw_lengthscale.copy_(self.raw_lengthscale_constraint.inverse_transform(value).expand_as(self.raw_lengthscale))

    def forward(self, x1, x2, diag=False, **params):  # pragma: no cover
        raise NotImplementedError

    def __call__(self, x1, x2=None, diag=False, **params):
        if x1.dim() == 1:
            x1 = x1.unsqueeze(-1)
        if x2 is None:
            x2 = x1
        elif x2.dim() == 1:
            x2 = x2.unsqueeze(-1)
        if self.active_dims is not None:
            same = x2 is x1
            x1 = x1.index_select(-1, self.active_dims.to(x1.device))
            x2 = x1 if same else x2.index_select(-1, self.active_dims.to(x2.device))
        return self.forward(x1, x2, diag=diag, **params)


class _StandInLazyTensor:
    def __init__(self, *args, **kwargs):
        self._args = args
        self._kwargs = kwargs

    def _matmul(self, rhs):  # pragma: no cover
        raise NotImplementedError

    def _size(self):  # pragma: no cover
        raise NotImplementedError

    def _transpose_nonbatch(self):  # pragma: no cover
        raise NotImplementedError

    def size(self, dim=None):
        s = self._size()
        return s if dim is None else s[dim]

    @property
    def shape(self):
        return self._size()

    def matmul(self, rhs):
        if rhs.dim() == 1:
            return self._matmul(rhs.unsqueeze(-1)).squeeze(-1)
        return self._matmul(rhs)

    __matmul__ = matmul

    def transpose(self, a=-1, b=-2):
        return self._transpose_nonbatch()

    def t(self):
        return self._transpose_nonbatch()

    def evaluate(self):
        n = self._size()[-1]
        ref = self._args[0]
        return self._matmul(torch.eye(n, dtype=ref.dtype, device=ref.device))

    to_dense = evaluate


if not HAVE_GPYTORCH:
    Kernel, LazyTensor = _StandInKernel, _StandInLazyTensor

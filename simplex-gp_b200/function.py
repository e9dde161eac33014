"""The autograd operator of the lattice filter and the lattice cache behind it.

Replaces ``LatticeFilterGeneral`` of the reference (gpytorch_lattice_kernel/bilateral_kernel.py:59-124): same
``apply(source, reference, kernel_fn)`` signature, same forward value and the same gradient formulas.

Differences in *how*, not *what*:

* The reference rebuilds the lattice inside every ``filter`` call.  Within one optimiser step GPyTorch calls the
  operator ~100 + n_cg times with the SAME ``x / lengthscale`` tensor, so lattices are cached
  on ``(tensor identity, tensor version, stencil variance)``; a Matern kernel needs two per step because its
  derivative stencil has a different variance than its forward stencil (permutohedral.h:388-389).
* The backward never materialises the reference's ``N x 2L(1+d)`` block (bilateral_kernel.py:113-118): RHS columns are
  packed, filtered and contracted ``chunk`` columns at a time by the kernels in ``csrc/sgp_grad.cu``.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
from torch.autograd import Function

from . import _capi
from ._capi import check
from .lattice import Lattice, _coeffs_np, _ptr, _stream_ptr, stencil_variance

__all__ = ["LatticeFilterGeneral", "LatticeCache", "lattice_cache", "lattice_filter_grad"]


class LatticeCache:
    """Small LRU of built lattices keyed on the position tensor.

    Fast path: the identity of the tensor object.  An entry matches only while the tensor it was built from is alive,
    is the very same object, and has not been modified in place (``Tensor._version``) -- a new tensor that happens to
    reuse the address never matches.  Slow path: a tensor that is a different object with the same shape is compared
    value for value with the cached ones (one ``torch.equal`` each, ~50 us at N = 1M): GPyTorch divides the training
    inputs by the lengthscale afresh in every call, so in evaluation mode each prediction arrives with a new tensor
    holding the same numbers as the last one.

    ``get_union`` serves the rectangular operator: the lattice of ``cat([x_base, x_new])`` is made by extending the
    cached lattice of ``x_base`` (``Lattice.extend``) instead of being built from nothing."""

    # products a union lattice serves from its neighbour table before building the blur groups and row-sorted entries
    # (2.0 ms at the metric shape, 0.1 ms saved per product afterwards: profiles/predict_step.py)
    lazy_union_products = 16

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self._entries = OrderedDict()
        self.hits = 0
        self.content_hits = 0
        self.builds = 0
        self.extensions = 0

    def clear(self) -> None:
        self._entries.clear()

    @staticmethod
    def _key(x: torch.Tensor, c: np.ndarray):
        var_bits = int(np.float32(stencil_variance(c)).view(np.int32))
        return (id(x), x.data_ptr(), tuple(x.shape), tuple(x.stride()), str(x.device), c.shape[0], var_bits)

    def _lookup(self, x: torch.Tensor, c: np.ndarray) -> Optional[Lattice]:
        key = self._key(x, c)
        ent = self._entries.get(key)
        if ent is not None:
            ref, version, lat, alias = ent
            if ref() is x and version == x._version:
                self._entries.move_to_end(key)
                self.hits += 1
                return lat
            del self._entries[key]
        for k, (ref, version, lat, alias) in list(self._entries.items()):
            if (k[2], k[4:]) != (key[2], key[4:]):      # shape, device, stencil length and variance
                continue
            if alias._version != version:       # the cached tensor was modified in place: the lattice is stale
                del self._entries[k]
                continue
            if alias.shape == x.shape and alias.device == x.device and torch.equal(alias, x.detach()):
                self.content_hits += 1
                self._store(x, c, lat)      # the next lookup of this tensor object is an identity hit
                return lat
        return None

    def _store(self, x: torch.Tensor, c: np.ndarray, lat: Lattice) -> None:
        try:
            ref = weakref.ref(x)
        except TypeError:  # pragma: no cover
            return
        self._entries[self._key(x, c)] = (ref, x._version, lat, x.detach())
        while len(self._entries) > self.capacity:
            self._entries.popitem(last=False)

    def get(self, x: torch.Tensor, coeffs, **build_kwargs) -> Lattice:
        c = _coeffs_np(coeffs)
        lat = self._lookup(x, c)
        if lat is None:
            lat = Lattice(x, c, **build_kwargs)
            self.builds += 1
            self._store(x, c, lat)
        return lat

    def get_union(self, union: torch.Tensor, n_base: int, coeffs, base: Optional[torch.Tensor] = None) -> Lattice:
        """Lattice of ``union = cat([x_base, x_new])`` (``n_base`` rows of ``x_base`` first), registered under
        ``union`` so that ``get(union, coeffs)`` finds it.  ``base``: the tensor ``x_base`` itself when the caller has
        it (its cached lattice is then found by identity)."""
        c = _coeffs_np(coeffs)
        lat = self._lookup(union, c)
        if lat is None:
            x_base = base if base is not None else union[:n_base]
            lat = self.get(x_base, c).extend(union[n_base:].detach(), lazy_tables=self.lazy_union_products)
            self.extensions += 1
            self._store(union, c, lat)
        return lat


lattice_cache = LatticeCache()


def _to_device(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t if t.device == dev else t.to(dev, non_blocking=True)


def _compute_device(source: torch.Tensor) -> torch.device:
    if source.is_cuda:
        return source.device
    if not torch.cuda.is_available():
        raise RuntimeError("LatticeFilterGeneral: no CUDA device; this package has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def lattice_filter_grad(lat: Lattice, g: torch.Tensor, v: torch.Tensor, x: torch.Tensor, deriv_coeffs,
                        want_grad_src: bool, chunk: Optional[int] = None):
    """``grad_reference`` (and the filtered ``g``) of bilateral_kernel.py:113-123 on the lattice ``lat`` (built from
    ``x`` with the derivative stencil).  All tensors fp32 on ``lat.device``; returns ``(grad_x [N,d], wg [N,L] or None)``."""
    lib = _capi.lib()
    dev = lat.device
    N, L = int(g.shape[0]), int(g.shape[1])
    d = lat.d
    g, v, x = g.contiguous(), v.contiguous(), x.contiguous()
    c = _coeffs_np(deriv_coeffs)
    grad_x = torch.empty((N, d), dtype=torch.float32, device=dev)
    wg = torch.empty((N, L), dtype=torch.float32, device=dev) if want_grad_src else None
    if N == 0 or L == 0:
        return grad_x.zero_(), wg
    per = 2 * (d + 1)
    if chunk is None:
        # columns per pass: as many as keep the packed block near 144 channels (measured best on B200: 5.8 ms vs 7.4 ms
        # at 54 channels for N=1M, d=8, L=16), rounded so that the channel count is a multiple of 4 whenever possible
        chunk = max(1, 144 // per)
        if (per * chunk) % 4 and (per * chunk * 2) % 4 == 0 and chunk * 2 <= max(L, 2):
            chunk *= 2
    chunk = max(1, min(int(chunk), L))
    width = per * chunk
    ldp = (width + 3) // 4 * 4
    packed = torch.empty((N, ldp), dtype=torch.float32, device=dev)
    filtered = torch.empty((N, ldp), dtype=torch.float32, device=dev)
    st = _stream_ptr(dev)
    with torch.cuda.device(dev):
        l0 = 0
        while l0 < L:
            nl = min(chunk, L - l0)
            # a short last pass filters only its own channels (the first ceil4(per * nl) columns of the buffers, same
            # row stride); sgp_grad_pack zero-fills up to the multiple of four
            w = (per * nl + 3) // 4 * 4
            pk, fl = (packed, filtered) if w == ldp else (packed[:, :w], filtered[:, :w])
            check(lib.sgp_grad_pack(_ptr(g), g.stride(0), _ptr(v), v.stride(0), _ptr(x), x.stride(0), N, d, l0, nl,
                                    _ptr(packed), ldp, st))
            lat.mvm(pk, out=fl, coeffs=c)
            check(lib.sgp_grad_contract(_ptr(filtered), ldp, _ptr(g), g.stride(0), _ptr(v), v.stride(0), _ptr(x),
                                        x.stride(0), N, d, l0, nl, int(l0 == 0), int(l0 + nl >= L), _ptr(grad_x),
                                        grad_x.stride(0), _ptr(wg), wg.stride(0) if wg is not None else 0, st))
            l0 += nl
    return grad_x, wg


class LatticeFilterGeneral(Function):
    """``LatticeFilterGeneral.apply(source[N,L], reference[N,d], kernel_fn) -> filtered[N,L]``.

    ``kernel_fn`` is a ``DiscretizedKernelFN`` (``get_coeffs()`` / ``get_deriv_coeffs()``).  ``reference`` is the
    input already divided by the lengthscale; its gradient is what carries the lengthscale gradient."""

    cache = lattice_cache
    grad_chunk = None   # RHS columns per backward pass (None = automatic)

    @staticmethod
    def forward(ctx, source, reference, kernel_fn):
        assert source.shape[0] == reference.shape[0], \
            "Incompatible shapes {}, and {}".format(source.shape, reference.shape)
        if source.dim() != 2 or reference.dim() != 2:
            raise NotImplementedError("batch dimensions are not supported (reference: bilateral_kernel.py:88)")
        dev = _compute_device(source)
        coeffs = kernel_fn.get_coeffs()
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(source, reference)
            ctx.kernel_fn = kernel_fn
            ctx.coeffs = coeffs
            ctx.deriv_coeffs = kernel_fn.get_deriv_coeffs()
        if reference.device == dev and reference.dtype == torch.float32:
            lat = LatticeFilterGeneral.cache.get(reference, coeffs)   # keyed on the caller's tensor object
            ctx.lat = lat
        else:
            lat = Lattice(_to_device(reference.detach().float().contiguous(), dev), coeffs)
            ctx.lat = None
        # the stencil is passed explicitly: the cache keys a lattice on the stencil's length and variance only, and two
        # stencils can share both (the variance is scale-invariant)
        out = lat.mvm(_to_device(source.detach().float(), dev), coeffs=coeffs)
        return out.to(device=source.device, dtype=source.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        with torch.no_grad():
            src, ref = ctx.saved_tensors
            dev = _compute_device(src)
            grad_source = grad_reference = None
            g = _to_device(grad_output.detach().float(), dev)
            on_dev = ref.device == dev and ref.dtype == torch.float32
            if ctx.needs_input_grad[0] and not ctx.needs_input_grad[1]:
                # the operator is symmetric: grad_source = filter(g) with the forward stencil (:110-111)
                lat = ctx.lat if ctx.lat is not None else Lattice(_to_device(ref.detach().float().contiguous(), dev),
                                                                  ctx.coeffs)
                grad_source = lat.mvm(g, coeffs=ctx.coeffs).to(device=src.device, dtype=src.dtype)
            if ctx.needs_input_grad[1]:
                x_dev = ref.detach() if on_dev else _to_device(ref.detach().float().contiguous(), dev)
                if on_dev:
                    lat = LatticeFilterGeneral.cache.get(ref, ctx.deriv_coeffs)
                else:
                    lat = Lattice(x_dev, ctx.deriv_coeffs)
                v = _to_device(src.detach().float(), dev)
                gx, wg = lattice_filter_grad(lat, g, v, x_dev.float(), ctx.deriv_coeffs, ctx.needs_input_grad[0],
                                             chunk=LatticeFilterGeneral.grad_chunk)
                grad_reference = gx.to(device=ref.device, dtype=ref.dtype)
                if ctx.needs_input_grad[0]:
                    grad_source = wg.to(device=src.device, dtype=src.dtype)   # filtered with the derivative stencil (:123)
        return grad_source, grad_reference, None

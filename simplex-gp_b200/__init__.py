"""simplex-gp_b200: B200-native Simplex-GP lattice kernel MVM.

Drop-in for the hot path of activatedgeek/simplex-gp (``gpytorch_lattice_kernel``): the native operator
``filter(src, ref, coeffs)``, the autograd op ``LatticeFilterGeneral`` and the GPyTorch kernels
``RBFLattice`` / ``MaternLattice``.  All compute runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/sgp_lattice.h``; there is no CPU fallback.

The directory name contains a hyphen, so the importable name is ``simplex_gp_b200`` (see ``simplex_gp_b200.py``
at the repository root).
"""
from .coeffs import DiscretizedKernelFN, Matern, get_coeffs, matern, rbf
from .function import LatticeCache, LatticeFilterGeneral, lattice_cache, lattice_filter_grad
from .kernels import (BilateralKernel, LatticeAccelerated, MaternLattice, RBFLattice, RectangularLazyLattice,
                      SquareLazyLattice)
from .lattice import Lattice, lattice_filter, scale_factors, slice_divisor, stencil_variance

filter = lattice_filter  # the reference's operator name (cpp/lattice.cpp:14-16)

__all__ = [
    "Lattice", "lattice_filter", "filter", "stencil_variance", "scale_factors", "slice_divisor",
    "get_coeffs", "DiscretizedKernelFN", "rbf", "matern", "Matern",
    "LatticeFilterGeneral", "LatticeCache", "lattice_cache", "lattice_filter_grad",
    "RBFLattice", "MaternLattice", "BilateralKernel", "LatticeAccelerated", "SquareLazyLattice",
    "RectangularLazyLattice",
]

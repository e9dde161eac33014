"""GPU parity of the autograd operator and the kernel modules against the reference's LatticeFilterGeneral
(golden vectors in tests/golden/autograd.npz, produced by the reference's Python + C++ code) and the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, RBF1, make_inputs

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


class FixedCoeffs:
    """kernel_fn stand-in that returns stored stencils (what DiscretizedKernelFN would compute)."""

    def __init__(self, fwd, deriv):
        self.f, self.d = torch.as_tensor(fwd), torch.as_tensor(deriv)

    def get_coeffs(self):
        return self.f

    def get_deriv_coeffs(self):
        return self.d


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


GOLD = np.load(os.path.join(GOLDEN_DIR, "autograd.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files})


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_forward_backward_match_reference(sg, case, device):
    g = {k.split("/")[1]: GOLD[k] for k in GOLD.files if k.startswith(case + "/")}
    kf = FixedCoeffs(g["coeffs"], g["deriv_coeffs"])
    x = torch.tensor(g["x"], device=device, requires_grad=True)
    v = torch.tensor(g["v"], device=device, requires_grad=True)
    go = torch.tensor(g["grad_out"], device=device)
    out = sg.LatticeFilterGeneral.apply(v, x, kf)
    assert out.device.type == device and out.shape == v.shape
    out.backward(go)
    assert _rel(out.detach().cpu().numpy(), g["out"]) < REL_TOL
    assert _rel(v.grad.cpu().numpy(), g["grad_src"]) < REL_TOL
    assert _rel(x.grad.cpu().numpy(), g["grad_ref"]) < REL_TOL
    # case A of the reference backward: only the source needs a gradient
    v2 = torch.tensor(g["v"], device=device, requires_grad=True)
    x2 = torch.tensor(g["x"], device=device)
    sg.LatticeFilterGeneral.apply(v2, x2, kf).backward(go)
    assert _rel(v2.grad.cpu().numpy(), g["grad_src_only"]) < REL_TOL


def test_gradient_chunking_is_invariant(sg):
    x, v = make_inputs(500, 4, 7, seed=3)
    kf = FixedCoeffs(RBF1, RBF1)
    go = torch.randn(500, 7, generator=torch.Generator().manual_seed(1)).cuda()
    grads = []
    for chunk in (None, 1, 3, 7):
        sg.LatticeFilterGeneral.grad_chunk = chunk
        xr = x.cuda().requires_grad_(True)
        vr = v.cuda().requires_grad_(True)
        sg.LatticeFilterGeneral.apply(vr, xr, kf).backward(go)
        grads.append((xr.grad.cpu().numpy(), vr.grad.cpu().numpy()))
    sg.LatticeFilterGeneral.grad_chunk = None
    for gx, gv in grads[1:]:
        assert _rel(gx, grads[0][0]) < 2e-6 and _rel(gv, grads[0][1]) < 2e-6


def test_lattice_cache(sg):
    cache = sg.lattice_cache
    cache.clear()
    x, v = make_inputs(300, 3, 2, seed=4)
    xd, vd = x.cuda(), v.cuda()
    kf = FixedCoeffs(RBF1, RBF1)
    b0, h0 = cache.builds, cache.hits
    a = sg.LatticeFilterGeneral.apply(vd, xd, kf)
    b = sg.LatticeFilterGeneral.apply(vd, xd, kf)
    assert cache.builds == b0 + 1 and cache.hits == h0 + 1
    xd.mul_(2.0)   # in-place change: the cached lattice must not be reused
    c = sg.LatticeFilterGeneral.apply(vd, xd, kf)
    assert cache.builds == b0 + 2
    assert not torch.allclose(a, c)
    y = xd.clone()  # a different tensor object with equal content: found by value, no rebuild
    ch0 = cache.content_hits
    e = sg.LatticeFilterGeneral.apply(vd, y, kf)
    assert cache.builds == b0 + 2 and cache.content_hits == ch0 + 1
    assert _rel(e.cpu().numpy(), c.cpu().numpy()) < 2e-6
    y2 = xd.clone()
    y2[5, 1] += 1e-3   # one value differs: a different lattice
    sg.LatticeFilterGeneral.apply(vd, y2, kf)
    assert cache.builds == b0 + 3
    xd.add_(1.0)       # the cached tensor changed in place after the build: its entry must not answer for y's content
    z = y.clone()
    sg.LatticeFilterGeneral.apply(vd, z, kf)
    assert cache.builds == b0 + 3 and cache.content_hits == ch0 + 2   # served by y's entry
    del y, y2, z
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_kernel_modules_and_lengthscale_gradient(sg, oracle):
    torch.manual_seed(0)
    N, d = 400, 3
    x, v = make_inputs(N, d, 2, seed=6)
    k = sg.RBFLattice(ard_num_dims=d, order=1).cuda()
    with torch.no_grad():
        k.raw_lengthscale.copy_(torch.tensor([[0.3, -0.2, 0.5]]))
    xd, vd = x.cuda(), v.cuda()
    op = k(xd)
    assert type(op).__name__ == "SquareLazyLattice" and tuple(op.shape) == (N, N)
    assert torch.equal(op.diag(), torch.ones(N, device="cuda"))
    assert torch.equal(k(xd, xd, diag=True), torch.ones(N, device="cuda"))
    out = op.matmul(vd)
    ls = k.lengthscale.detach().cpu()
    want = oracle.filter(v.numpy(), (x / ls).numpy(), k.dkernel_fn.get_coeffs().numpy())
    assert _rel(out.detach().cpu().numpy(), want) < REL_TOL
    # lengthscale gradient = chain rule through x / lengthscale of the operator's grad_reference
    go = torch.randn(N, 2, generator=torch.Generator().manual_seed(2)).cuda()
    (out * go).sum().backward()
    xs = (xd / k.lengthscale.detach()).requires_grad_(True)
    sg.LatticeFilterGeneral.apply(vd, xs, k.dkernel_fn).backward(go)
    manual = (xs.grad * (-xd / k.lengthscale.detach() ** 2)).sum(0, keepdim=True) * torch.sigmoid(k.raw_lengthscale.detach())
    assert _rel(k.raw_lengthscale.grad.cpu().numpy(), manual.cpu().numpy()) < 1e-4
    assert k.raw_lengthscale.grad.abs().sum() > 0


def test_rectangular_operator(sg, oracle):
    xin, _ = make_inputs(150, 3, 1, seed=7)
    xout, v = make_inputs(220, 3, 2, seed=8)
    k = sg.MaternLattice(nu=1.5, order=2).cuda()
    op = k(xin.cuda(), xout.cuda())
    assert type(op).__name__ == "RectangularLazyLattice" and tuple(op.shape) == (150, 220)
    got = op.matmul(v.cuda())
    assert tuple(got.shape) == (150, 2)
    ls = float(k.lengthscale.detach())
    big_x = torch.cat([xout, xin]) / ls
    big_v = torch.cat([v, torch.zeros(150, 2)])
    want = oracle.filter(big_v.numpy(), big_x.numpy(), k.dkernel_fn.get_coeffs().numpy())[220:]
    assert _rel(got.detach().cpu().numpy(), want) < REL_TOL
    t = op.transpose(-1, -2)
    assert tuple(t.shape) == (220, 150)
    # the transpose shares the union lattice: it is the other off-diagonal block of the same union filter, which is
    # what the reference computes on cat([xin, xout]) (the same key set numbered differently)
    u = torch.randn(150, 2, generator=torch.Generator().manual_seed(3))
    b0 = sg.lattice_cache.builds + sg.lattice_cache.extensions
    tu = t.matmul(u.cuda())
    assert sg.lattice_cache.builds + sg.lattice_cache.extensions == b0
    assert tuple(tu.shape) == (220, 2)
    want_t = oracle.filter(torch.cat([torch.zeros(220, 2), u]).numpy(), big_x.numpy(),
                           k.dkernel_fn.get_coeffs().numpy())[:220]
    assert _rel(tu.detach().cpu().numpy(), want_t) < REL_TOL
    other = oracle.filter(torch.cat([u, torch.zeros(220, 2)]).numpy(), (torch.cat([xin, xout]) / ls).numpy(),
                          k.dkernel_fn.get_coeffs().numpy())[150:]
    assert _rel(tu.detach().cpu().numpy(), other) < REL_TOL
    assert t.transpose(-1, -2) is op


def test_rectangular_operator_extends_the_cached_training_lattice(sg, oracle):
    """Prediction-shaped use: the square operator on the training inputs first, then K(test, train) products.  The
    union lattice must come from extending the cached training lattice (no second build of the training points), be
    reused across products and across fresh `x / lengthscale` tensors, and give the reference's numbers."""
    cache = sg.lattice_cache
    cache.clear()
    xtr, v = make_inputs(3000, 4, 3, seed=21)
    xte, _ = make_inputs(400, 4, 1, seed=22)
    k = sg.RBFLattice(ard_num_dims=4, order=1).cuda()
    xtr_d, xte_d, vd = xtr.cuda(), xte.cuda(), v.cuda()
    with torch.no_grad():
        k(xtr_d).matmul(vd)                              # builds the training lattice
        b0, e0 = cache.builds, cache.extensions
        op = k(xte_d, xtr_d)                             # fresh xtr / lengthscale tensor: found by value
        got = op.matmul(vd)
        got2 = op.matmul(vd * 2)
        assert cache.builds == b0 and cache.extensions == e0 + 1
        got3 = k(xte_d, xtr_d).matmul(vd)               # a new operator object: union found by value
        assert cache.builds == b0 and cache.extensions == e0 + 1
    ls = k.lengthscale.detach().cpu()
    big_x = torch.cat([xtr, xte]) / ls
    big_v = torch.cat([v, torch.zeros(400, 3)])
    want = oracle.filter(big_v.numpy(), big_x.numpy(), k.dkernel_fn.get_coeffs().numpy())[3000:]
    assert _rel(got.cpu().numpy(), want) < REL_TOL
    assert _rel(got3.cpu().numpy(), got.cpu().numpy()) < 2e-6
    assert _rel(got2.cpu().numpy(), 2 * want) < REL_TOL


def test_matern_nu_validation(sg):
    with pytest.raises(NotImplementedError):
        sg.MaternLattice(nu=0.7)


def test_cached_lattice_serves_stencils_of_equal_variance(sg):
    """The lattice cache keys on the stencil's length and variance, and the variance is scale-invariant: a second
    kernel function whose stencil is a multiple of the first one's reuses the cached lattice and must still be applied
    with ITS coefficients (every product passes the stencil explicitly)."""
    class Scaled:
        def __init__(self, c, dc):
            self.c, self.dc = torch.tensor(c), torch.tensor(dc)

        def get_coeffs(self):
            return self.c

        def get_deriv_coeffs(self):
            return self.dc

    d = 3
    x, v = make_inputs(400, d, 2, seed=8)
    xd, vd = x.cuda(), v.cuda()
    base = Scaled(RBF1, RBF1)
    twice = Scaled([2 * c for c in RBF1], [2 * c for c in RBF1])
    sg.lattice_cache.clear()
    builds = sg.lattice_cache.builds
    a = sg.LatticeFilterGeneral.apply(vd, xd, base)
    b = sg.LatticeFilterGeneral.apply(vd, xd, twice)
    assert sg.lattice_cache.builds == builds + 1                 # one lattice, two stencils
    assert float((b - 2.0 ** (d + 1) * a).norm() / b.norm()) < 1e-5   # d+1 blur passes, each doubled

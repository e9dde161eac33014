"""Runs only where GPyTorch is installed (it is not in the build image: no network).  The north star's promise is that
``RBFLattice`` / ``MaternLattice`` drop into ``gpytorch.kernels.ScaleKernel`` unchanged
(gpytorch_lattice_kernel/bilateral_kernel.py:127-140,183-200; tests/train_snelson.py:11-23): these tests execute exactly
that through real GPyTorch -- the kernel derives from ``gpytorch.kernels.Kernel``, the operator from GPyTorch's lazy
tensor / ``LinearOperator`` -- and compare GPyTorch's exact marginal log-likelihood with this package's own solver."""
import pytest
import torch

gpytorch = pytest.importorskip("gpytorch")
pytestmark = pytest.mark.gpu


class _Model(gpytorch.models.ExactGP):
    def __init__(self, x, y, likelihood, kernel):
        super().__init__(x, y, likelihood)
        self.mean_module = gpytorch.means.ConstantMean()
        self.covar_module = gpytorch.kernels.ScaleKernel(kernel)

    def forward(self, x):
        return gpytorch.distributions.MultivariateNormal(self.mean_module(x), self.covar_module(x))


def _snelson_like(n=200):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 1, generator=g) * 6
    y = torch.sin(x[:, 0] * 2) + 0.1 * torch.randn(n, generator=g)
    return x.cuda(), y.cuda()


def test_kernel_classes_derive_from_gpytorch(sg):
    from simplex_gp_b200 import kernels
    assert kernels.HAVE_GPYTORCH
    k = sg.RBFLattice(ard_num_dims=3, order=1)
    assert isinstance(k, gpytorch.kernels.Kernel)
    assert k.lengthscale.shape[-1] == 3
    sk = gpytorch.kernels.ScaleKernel(sg.MaternLattice(nu=1.5, order=2, ard_num_dims=3)).cuda()
    x = torch.randn(50, 3, device="cuda")
    v = torch.randn(50, 2, device="cuda")
    op = sk(x)
    out = op.matmul(v) if hasattr(op, "matmul") else op @ v
    assert out.shape == (50, 2) and torch.isfinite(out).all()


def test_scale_kernel_mll_and_gradients_through_gpytorch(sg):
    """tests/train_snelson.py of the reference, through real GPyTorch: one MLL evaluation + backward, and a few Adam steps."""
    from simplex_gp_b200 import gp
    x, y = _snelson_like()
    lik = gpytorch.likelihoods.GaussianLikelihood().cuda()
    model = _Model(x, y, lik, sg.RBFLattice(order=1)).cuda()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    model.train(); lik.train()
    with gpytorch.settings.max_cholesky_size(800):
        loss = -mll(model(x), y)
    loss.backward()
    ls_grad = model.covar_module.base_kernel.raw_lengthscale.grad
    assert torch.isfinite(loss) and ls_grad is not None and torch.isfinite(ls_grad).all() and float(ls_grad.abs().sum()) > 0
    # the same model in this package's own solver (same parameterisation: softplus of raw values initialised to 0)
    own = gp.ExactGPModel(x, y, sg.RBFLattice(order=1).cuda()).cuda()
    assert abs(float(-own.mll()) - float(loss)) < 5e-3
    opt = torch.optim.Adam(model.parameters(), lr=0.1)
    first = float(loss)
    for _ in range(20):
        opt.zero_grad()
        with gpytorch.settings.max_cholesky_size(800):
            loss = -mll(model(x), y)
        loss.backward()
        opt.step()
    assert float(loss) < first


def test_cg_path_and_prediction_through_gpytorch(sg):
    """N > max_cholesky_size: GPyTorch's own CG / Lanczos / pivoted-Cholesky machinery over the lattice operator with
    the reference's settings (experiments/train_simplexgp.py:29-84)."""
    g = torch.Generator().manual_seed(1)
    n, d = 3000, 3
    x = torch.randn(n, d, generator=g).cuda()
    y = (torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g).cuda())
    lik = gpytorch.likelihoods.GaussianLikelihood().cuda()
    model = _Model(x, y, lik, sg.RBFLattice(ard_num_dims=d, order=1)).cuda()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    model.train(); lik.train()
    with gpytorch.settings.cg_tolerance(1.0), gpytorch.settings.max_cg_iterations(1000), \
            gpytorch.settings.max_preconditioner_size(100), gpytorch.settings.max_root_decomposition_size(100), \
            gpytorch.settings.max_cholesky_size(0):
        loss = -mll(model(x), y)
        loss.backward()
    assert torch.isfinite(loss)
    assert torch.isfinite(model.covar_module.base_kernel.raw_lengthscale.grad).all()
    model.eval(); lik.eval()
    xt = torch.randn(40, d, generator=g).cuda()
    with torch.no_grad(), gpytorch.settings.eval_cg_tolerance(1e-2), gpytorch.settings.max_cholesky_size(0), \
            gpytorch.settings.fast_pred_var():
        pred = model(xt)
        assert torch.isfinite(pred.mean).all() and pred.mean.shape == (40,)

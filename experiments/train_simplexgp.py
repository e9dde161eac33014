"""Train and evaluate an exact GP whose kernel matrix is the lattice operator -- the caller either side of the hot path.

Restates experiments/train_simplexgp.py and experiments/utils.py:21-198 of the reference (the sweep of
configs/simplexgp.yml: Adam lr 0.1, 100 epochs, CG tolerance 1.0 for training and 1e-2 for evaluation, Matern-1.5
order 1, noise >= 0.1, early stopping on the validation RMSE) on this package's operator and solver
(``simplex_gp_b200.gp``; GPyTorch, wandb and fire are not available here).  The UCI ``.mat`` files cannot be downloaded,
so by default a synthetic regression problem with the named dataset's shape is generated; ``--data-dir`` loads the
real file (``<dir>/<name>/<name>.mat``, key ``data``, last column the target) exactly as the reference does.

    python experiments/train_simplexgp.py --dataset elevators --epochs 20
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from typing import Iterator, Optional, Tuple

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# rows x input dimensions of the datasets the reference sweeps (configs/simplexgp.yml:5-11), from
# notebooks/viz_compute.ipynb:102-103 and BASELINE.json (elevators-shaped: d = 18)
SHAPES = {
    "elevators": (16_599, 18), "houseelectric": (2_049_280, 11), "3droad": (434_874, 3),
    "keggdirected": (48_827, 20), "protein": (45_730, 9), "precipitation3d_all": (628_474, 3),
    "toy": (4_000, 3),
}


def set_seeds(seed: Optional[int]) -> None:
    if seed is not None and seed >= 0:
        torch.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)


def synthetic_table(name: str, seed: int = 0, max_n: Optional[int] = None) -> torch.Tensor:
    """``[N, d+1]`` float32 table shaped like ``name`` (inputs, then the target in the last column): correlated
    inputs, a smooth low-dimensional target plus noise -- something a stationary kernel can learn."""
    N, d = SHAPES[name]
    if max_n is not None:
        N = min(N, int(max_n))
    g = torch.Generator().manual_seed(1000 + seed)
    mix = torch.randn(d, d, generator=g) / math.sqrt(d) + torch.eye(d)
    x = torch.randn(N, d, generator=g) @ mix
    w = torch.randn(d, 3, generator=g) / math.sqrt(d)
    z = x @ w
    y = torch.sin(z[:, 0]) + 0.5 * torch.cos(1.5 * z[:, 1]) * z[:, 2] + 0.1 * torch.randn(N, generator=g)
    return torch.cat([x, y[:, None]], dim=1).float()


def load_table(name: str, data_dir: Optional[str], seed: int = 0, max_n: Optional[int] = None) -> torch.Tensor:
    if data_dir is None:
        return synthetic_table(name, seed, max_n)
    from scipy.io import loadmat   # the reference's loader (utils.py:60)
    data = torch.as_tensor(loadmat(os.path.join(data_dir, name, name + ".mat"))["data"], dtype=torch.float32)
    return data if max_n is None else data[: int(max_n)]


def prepare_dataset(name: str, data_dir: Optional[str] = None, device="cpu", train_val_split: float = 0.8,
                    seed: int = 0, max_n: Optional[int] = None) -> Iterator[Tuple[str, torch.Tensor, torch.Tensor]]:
    """Yields ``("train" | "val" | "test", x, y)``: the first 64 % / next 16 % / last 20 % of the rows, standardised
    with the training split's statistics (utils.py:21-45, 66-72)."""
    data = load_table(name, data_dir, seed, max_n).to(device)
    N = data.shape[0]
    n_train_val = int(train_val_split * N)
    n_train = int(train_val_split * n_train_val)
    parts = {"train": data[:n_train], "val": data[n_train:n_train_val], "test": data[n_train_val:]}
    tx, ty = parts["train"][:, :-1], parts["train"][:, -1]
    x_mean, x_scale = tx.mean(0, keepdim=True), tx.std(0, keepdim=True) + 2e-6   # the reference adds 1e-6 twice
    y_mean, y_scale = ty.mean(0, keepdim=True), ty.std(0, keepdim=True) + 2e-6
    for mode, part in parts.items():
        yield mode, ((part[:, :-1] - x_mean) / x_scale).contiguous(), ((part[:, -1] - y_mean) / y_scale).contiguous()


class EarlyStopper:
    """Keeps the best ``info`` by score (higher is better); done after ``patience`` calls without an improvement of at
    least ``delta`` (utils.py:170-198)."""

    def __init__(self, patience: int = 10, delta: float = 1e-4):
        self.patience, self.delta = patience, delta
        self.stale = 0
        self.best_score = None
        self.best_info = None

    def is_done(self) -> bool:
        return self.patience >= 0 and self.stale >= self.patience

    def __call__(self, score: float, info) -> None:
        assert not self.is_done()
        if self.best_score is None or score >= self.best_score + self.delta:
            self.best_score, self.best_info = score, info
        else:
            self.stale += 1


def build_model(train_x, train_y, nu: Optional[float], order: int, min_noise: float):
    import simplex_gp_b200 as sg
    from simplex_gp_b200 import gp
    d = train_x.shape[-1]
    kernel = sg.MaternLattice(ard_num_dims=d, nu=nu, order=order) if nu is not None else \
        sg.RBFLattice(ard_num_dims=d, order=order)
    return gp.ExactGPModel(train_x, train_y, kernel.to(train_x.device), min_noise=min_noise).to(train_x.device)


def train(model, optim, probes, cg_iter: int = 500, cg_tol: float = 1.0, pre_size: int = 100) -> dict:
    """One optimiser step on the negative marginal log-likelihood (train_simplexgp.py:29-57)."""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    optim.zero_grad()
    res = model.mll(probes=probes, tol=cg_tol, max_iter=cg_iter, preconditioner_size=pre_size) \
        if model.train_x.shape[0] > model.max_cholesky_size \
        else model.mll()
    value, surrogate = res if isinstance(res, tuple) else (float(res.detach()), res)
    torch.cuda.synchronize()
    loss_ts = time.perf_counter() - t0
    (-surrogate).backward()
    optim.step()
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    return {"train/mll": float(value), "train/loss_ts": loss_ts, "train/bw_ts": total - loss_ts, "train/total_ts": total}


def test(x, y, model, cg_iter: int = 500, cg_tol: float = 1e-2, label: str = "test", variance_points: int = 0,
         pre_size: int = 100) -> dict:
    """RMSE / MAE of the posterior mean (train_simplexgp.py:60-84); the NLL on the first ``variance_points`` points
    when asked for (exact variance, one CG solve per 16 points)."""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mean = model.predict(x, tol=cg_tol, max_iter=cg_iter, preconditioner_size=pre_size)
    torch.cuda.synchronize()
    out = {f"{label}/rmse": float((mean - y).pow(2).mean().sqrt()), f"{label}/mae": float((mean - y).abs().mean()),
           f"{label}/pred_ts": time.perf_counter() - t0}
    if variance_points > 0:
        k = min(int(variance_points), x.shape[0])
        m, var = model.predict(x[:k].contiguous(), tol=cg_tol, max_iter=cg_iter, variance=True, preconditioner_size=pre_size)
        sd = (var + model.noise.detach()).sqrt()
        out[f"{label}/nll"] = float(-torch.distributions.Normal(m, sd).log_prob(y[:k]).mean())
    return out


def main(dataset: str = "elevators", data_dir: Optional[str] = None, log_int: int = 1, seed: Optional[int] = None,
         device: int = 0, epochs: int = 100, lr: float = 0.1, p_epochs: int = 200, n_probes: int = 10,
         cg_iter: int = 500, cg_tol: float = 1.0, cg_eval_tol: float = 1e-2, nu: Optional[float] = 1.5, order: int = 1,
         min_noise: float = 0.1, max_n: Optional[int] = None, variance_points: int = 0, quiet: bool = False,
         pre_size: int = 100) -> dict:
    if not torch.cuda.is_available():
        raise RuntimeError("experiments/train_simplexgp.py needs a CUDA device: the lattice operator has no CPU path")
    set_seeds(seed)
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    splits = {m: (x, y) for m, x, y in prepare_dataset(dataset, data_dir, dev, seed=seed or 0, max_n=max_n)}
    (train_x, train_y), (val_x, val_y), (test_x, test_y) = splits["train"], splits["val"], splits["test"]
    log = (lambda *a: None) if quiet else (lambda rec: print(json.dumps(rec), flush=True))
    log({"dataset": dataset, "D": train_x.shape[-1], "N_train": train_x.shape[0], "N_val": val_x.shape[0],
         "N_test": test_x.shape[0], "synthetic": data_dir is None})
    model = build_model(train_x, train_y, nu, order, min_noise)
    optim = torch.optim.Adam(model.parameters(), lr=lr)
    probes = torch.randn(train_x.shape[0], n_probes, device=dev).sign()
    stopper = EarlyStopper(patience=p_epochs)
    for i in range(epochs):
        rec = {"step": i + 1, **train(model, optim, probes, cg_iter=cg_iter, cg_tol=cg_tol, pre_size=pre_size)}
        if i % log_int == 0:
            rec.update(test(val_x, val_y, model, cg_iter, cg_eval_tol, "val", pre_size=pre_size))
            rec.update(test(test_x, test_y, model, cg_iter, cg_eval_tol, "test", variance_points, pre_size=pre_size))
            rec.update({"param/noise": float(model.noise.detach()), "param/outputscale": float(model.outputscale.detach())})
            stopper(-rec["val/rmse"], {"state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
                                       "summary": {"test/best_rmse": rec["test/rmse"], "val/best_step": i + 1,
                                                   **({"test/best_nll": rec["test/nll"]} if "test/nll" in rec else {})}})
        log(rec)
        if stopper.is_done():
            break
    summary = dict(stopper.best_info["summary"]) if stopper.best_info else {}
    summary["lengthscale"] = [float(v) for v in model.kernel.lengthscale.detach().flatten()]
    log({"summary": summary})
    return summary


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--dataset", default="elevators", choices=sorted(SHAPES))
    ap.add_argument("--data-dir", default=None)
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--p-epochs", type=int, default=200)
    ap.add_argument("--log-int", type=int, default=1)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--n-probes", type=int, default=10)
    ap.add_argument("--cg-iter", type=int, default=500)
    ap.add_argument("--cg-tol", type=float, default=1.0)
    ap.add_argument("--cg-eval-tol", type=float, default=1e-2)
    ap.add_argument("--nu", type=float, default=1.5, help="Matern smoothness (1.5 or 2.5); negative selects the RBF kernel")
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--min-noise", type=float, default=0.1)
    ap.add_argument("--max-n", type=int, default=None, help="use only the first MAX_N rows")
    ap.add_argument("--variance-points", type=int, default=0)
    a = ap.parse_args()
    kw = vars(a)
    kw["nu"] = None if a.nu is not None and a.nu < 0 else a.nu
    main(**kw)

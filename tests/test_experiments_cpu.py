"""Host logic of the experiment harness (experiments/train_simplexgp.py): dataset split, standardisation, early
stopping -- the rules of the reference's experiments/utils.py:21-45,66-72,170-198."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _harness():
    spec = importlib.util.spec_from_file_location("train_simplexgp", os.path.join(ROOT, "experiments", "train_simplexgp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_dataset_shapes_split_and_standardisation():
    h = _harness()
    assert h.SHAPES["elevators"] == (16_599, 18) and h.SHAPES["houseelectric"] == (2_049_280, 11)
    parts = {m: (x, y) for m, x, y in h.prepare_dataset("elevators", seed=3)}
    n = 16_599
    n_tv = int(0.8 * n)
    n_tr = int(0.8 * n_tv)
    assert parts["train"][0].shape == (n_tr, 18) and parts["val"][0].shape == (n_tv - n_tr, 18)
    assert parts["test"][0].shape == (n - n_tv, 18) and parts["test"][1].shape == (n - n_tv,)
    tx, ty = parts["train"]
    assert tx.dtype == torch.float32 and tx.is_contiguous()
    assert float(tx.mean(0).abs().max()) < 1e-4 and float((tx.std(0) - 1).abs().max()) < 1e-4
    assert abs(float(ty.mean())) < 1e-4 and abs(float(ty.std()) - 1) < 1e-4
    # the other splits are scaled with the training statistics: the three splits are one table, cut in order
    raw = h.synthetic_table("elevators", seed=3)
    mu, sd = raw[:n_tr, :-1].mean(0, keepdim=True), raw[:n_tr, :-1].std(0, keepdim=True) + 2e-6
    torch.testing.assert_close(parts["test"][0], (raw[n_tv:, :-1] - mu) / sd)
    # deterministic in the seed, different across seeds, max_n truncates
    again = h.synthetic_table("elevators", seed=3)
    assert torch.equal(raw, again) and not torch.equal(raw, h.synthetic_table("elevators", seed=4))
    assert h.synthetic_table("houseelectric", max_n=1000).shape == (1000, 12)


def test_early_stopper():
    h = _harness()
    s = h.EarlyStopper(patience=2, delta=0.1)
    s(-1.0, "a")
    s(-0.95, "b")          # improvement below delta: stale
    assert s.best_info == "a" and not s.is_done()
    s(-0.5, "c")
    assert s.best_info == "c" and s.stale == 1
    s(-0.6, "d")
    assert s.is_done() and s.best_info == "c"
    never = h.EarlyStopper(patience=-1)
    for k in range(5):
        never(0.0, k)
    assert not never.is_done()

"""Host logic of the training loop (no GPU): the reference's stop rule
``while abs(old - new) > thresh`` (HiC-GNN_main.py:118-131, HiC_GAT_generalize_directly.py:202,228) as
``train.fit`` implements it around the fused step, checked with a scripted step."""
import pytest
import torch

from hic_gnn_b200 import train


def reference_loop(losses, thresh, max_steps=None):
    """The reference's loop shape: oldloss = 1, lossdiff = 1; while lossdiff > thresh: step; lossdiff = |old - new|."""
    hist, old, diff, it = [], 1.0, 1.0, iter(losses)
    while diff > thresh and (max_steps is None or len(hist) < max_steps):
        new = next(it)
        hist.append(new)
        diff, old = abs(old - new), new
    return hist


class ScriptedStep:
    def __init__(self, losses):
        self.losses, self.calls = list(losses), 0

    def __call__(self):
        v = self.losses[self.calls]
        self.calls += 1
        return torch.tensor(v, dtype=torch.float64), None


class _Model:
    def train(self):
        return self


LOSSES = [0.9, 0.5, 0.4, 0.4 + 5e-9, 0.3, 0.2, 0.2, 0.1, 0.05, 0.05]


@pytest.mark.parametrize("thresh", [1e-8, 1e-2, 0.2, 0.0])
@pytest.mark.parametrize("max_steps", [None, 3, 7])
def test_fit_stops_like_the_reference_loop(monkeypatch, thresh, max_steps):
    if thresh == 0.0 and max_steps is None:
        max_steps = len(LOSSES)  # |old - new| <= 0 only fires on an exact repeat (index 6)
    step = ScriptedStep(LOSSES)
    monkeypatch.setattr(train, "TrainStep", lambda *a, **k: step)
    got = train.fit(_Model(), None, None, None, mode="mse", thresh=thresh, max_steps=max_steps)
    want = reference_loop(LOSSES, thresh, max_steps)
    assert got == want and step.calls == len(want)


@pytest.mark.parametrize("check_every", [2, 3, 4])
def test_fit_batched_readback_overruns_by_less_than_one_batch(monkeypatch, check_every):
    step = ScriptedStep(LOSSES)
    monkeypatch.setattr(train, "TrainStep", lambda *a, **k: step)
    got = train.fit(_Model(), None, None, None, mode="mse", thresh=1e-8, check_every=check_every)
    want = reference_loop(LOSSES, 1e-8)
    assert got[: len(want)] == want                       # same trajectory ...
    assert len(want) <= len(got) < len(want) + check_every  # ... read back in batches
    assert got == LOSSES[: len(got)]


def test_unknown_mode_is_rejected():
    with pytest.raises(ValueError):
        train.TrainStep(_Model(), None, None, None, mode="spearman")


def test_nan_loss_ends_training_like_the_reference():
    """`while lossdiff > thresh` is False for a NaN difference: the reference stops after a divergence, and so must
    fit() (an `abs(old - new) <= thresh` test would spin forever when max_steps is None)."""
    import pytest as _pt

    losses = [0.9, 0.5, float("nan"), 0.4, 0.3]
    step = ScriptedStep(losses)
    mp = _pt.MonkeyPatch()
    try:
        mp.setattr(train, "TrainStep", lambda *a, **k: step)
        got = train.fit(_Model(), None, None, None, mode="mse", thresh=1e-8)
    finally:
        mp.undo()
    assert len(got) == 3 and got[2] != got[2] and step.calls == 3
    assert len(reference_loop(losses, 1e-8)) == 3

"""CPU checks of the loss restatement (utils/train_utils.py:146-185): the oracle
against the vectors produced by the reference's own cls_loss / reg_loss source
(tests/golden/make_golden.py), known answers, and the analytic gradients against float64
finite differences of the same formulas."""
import numpy as np
import pytest

from oracle import rpn_oracle as O

F32 = np.float32
# TF sums the per-entry terms in float32 in an unspecified order; the oracle sums in float64.
LOSS_RTOL = 2e-6


def test_losses_match_reference_vectors(golden):
    reg = O.reg_loss(golden["loss_reg_true"], golden["loss_reg_pred"])
    cls = O.cls_loss(golden["loss_cls_true"], golden["loss_cls_pred"])
    assert abs(float(reg) - float(golden["loss_reg"])) <= LOSS_RTOL * abs(float(golden["loss_reg"]))
    assert abs(float(cls) - float(golden["loss_cls"])) <= LOSS_RTOL * abs(float(golden["loss_cls"]))


def test_loss_known_answers():
    # one positive row: |e| = [0.5, 2, 0, 1] -> 0.125 + 1.5 + 0 + 0.5 = 2.125; one row -> / 1
    t = np.zeros((1, 3, 4), F32)
    t[0, 1] = [1, 1, 1, 1]
    p = np.zeros((1, 3, 4), F32)
    p[0, 1] = [1.5, 3, 1, 0]
    p[0, 2] = [9, 9, 9, 9]                      # true row is zero: masked out
    assert float(O.reg_loss(t, p)) == pytest.approx(2.125, rel=1e-7)
    # no positive row: 0 / max(1, 0)
    assert float(O.reg_loss(np.zeros((2, 5, 4), F32), np.ones((2, 5, 4), F32))) == 0.0
    # BCE of p = 0.5 is log 2 whatever the target; ignored entries do not count
    y = np.array([1, 0, -1, -1], F32)
    q = np.array([0.5, 0.5, 0.01, 0.99], F32)
    assert float(O.cls_loss(y, q)) == pytest.approx(np.log(2.0), rel=1e-6)
    # clipping: p = 0 with target 1 -> -log(2e-7)
    assert float(O.cls_loss(np.array([1], F32), np.array([0], F32))) == pytest.approx(-np.log(2e-7), rel=1e-5)
    # nothing to average -> NaN (mean of an empty tensor), as TF
    assert np.isnan(O.cls_loss(np.full((4,), -1, F32), np.full((4,), 0.5, F32)))


def _f64_losses(td, pd, tl, pl):
    e = np.abs(pd - td)
    q = np.minimum(e, 1.0)
    h = (0.5 * q * q + (e - q)).sum(-1)
    pos = np.any(td != 0, axis=-1)
    reg = (h * pos).sum() / max(1, pos.sum())
    m = tl != -1
    eps = np.float64(np.float32(1e-7))
    p = np.clip(pl, eps, 1 - eps)
    bce = -(tl * np.log(p + eps) + (1 - tl) * np.log(1 - p + eps))
    return reg, bce[m].mean()


def test_loss_gradients_match_finite_differences():
    rng = np.random.default_rng(11)
    td = np.zeros((2, 40, 4), F32)
    td[:, :6] = rng.normal(0, 1, size=(2, 6, 4))
    pd = rng.normal(0, 1.2, size=(2, 40, 4)).astype(F32)
    tl = rng.integers(-1, 2, size=(2, 40)).astype(F32)
    pl = rng.uniform(0.02, 0.98, size=(2, 40)).astype(F32)
    gd, gl = O.loss_grads(td, pd, tl, pl)
    td64, pd64, tl64, pl64 = (a.astype(np.float64) for a in (td, pd, tl, pl))
    h = 1e-6
    for idx in [(0, 0, 0), (1, 5, 3), (0, 3, 2), (1, 20, 1)]:
        a, b = pd64.copy(), pd64.copy()
        a[idx] += h
        b[idx] -= h
        fd = (_f64_losses(td64, a, tl64, pl64)[0] - _f64_losses(td64, b, tl64, pl64)[0]) / (2 * h)
        assert gd[idx] == pytest.approx(fd, rel=1e-4, abs=1e-7)
    for idx in [(0, 0), (1, 7), (0, 39), (1, 21)]:
        a, b = pl64.copy(), pl64.copy()
        a[idx] += h
        b[idx] -= h
        fd = (_f64_losses(td64, pd64, tl64, a)[1] - _f64_losses(td64, pd64, tl64, b)[1]) / (2 * h)
        assert gl[idx] == pytest.approx(fd, rel=1e-4, abs=1e-7)
    assert np.all(gd[:, 6:] == 0) and np.all(gl[tl == -1] == 0)

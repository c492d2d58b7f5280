"""CPU check of the hand-written adjoint orchestration (dsvi_step.py).

The C-ABI wrappers in ``_ops`` are monkeypatched with their CPU specifications
(oracle/kernel_specs.py) -- test infrastructure only -- so the sequence of kernels and
every adjoint formula is compared with the autograd oracle / the golden vectors of the
reference without a GPU.  The GPU tests then only have to show kernel == spec.
"""
import inspect

import numpy as np
import pytest
import torch

from oracle import kernel_specs as specs
from oracle import nmgp_oracle as orc
from tests import golden_util as gu

pkg = pytest.importorskip("collaborative_nonstationary_multivariate_gaussian_process_b200")
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops, dsvi_step  # noqa: E402


@pytest.fixture
def spec_ops(monkeypatch):
    names = [n for n, f in inspect.getmembers(specs, inspect.isfunction) if not n.startswith("_")]
    for n in names:
        assert hasattr(_ops, n), "spec %s has no C-ABI wrapper" % n
        monkeypatch.setattr(_ops, n, getattr(specs, n))
    return names


def run_step(g, sample_chunk=None):
    p = gu.case_params(g)
    I = torch.from_numpy(g["I"]).to(torch.int32)
    loss, grads = dsvi_step.dsvi_step(
        p, torch.from_numpy(g["Z"]), torch.from_numpy(g["x"]), torch.from_numpy(gu.case_targets(g)), I, int(g["N"]),
        torch.from_numpy(g["z_v"]), torch.from_numpy(g["z_ell"]), torch.from_numpy(g["z_L"]),
        sample_chunk=sample_chunk)
    return loss, grads


@pytest.mark.parametrize("name", gu.DSVI_CASES)
def test_step_matches_reference_golden(spec_ops, name):
    g = gu.load(name)
    loss, grads = run_step(g, sample_chunk=1 if name == "dsvi_ragged" else None)
    assert abs(float(loss) - float(g["loss"])) <= 1e-10 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    train_len = bool(int(g["train_len"]))
    for k in orc.PARAM_NAMES:
        if k.startswith("length_scales") and not train_len:
            continue
        gu.check_grad(k, grads[k].numpy(), g, 1e-9)


def test_every_wrapper_has_a_spec():
    names = [n for n, f in inspect.getmembers(_ops, inspect.isfunction)
             if not n.startswith("_") and f.__module__ == _ops.__name__]
    missing = [n for n in names if not hasattr(specs, n)]
    assert not missing, missing


def _subject_case(S=3):
    """dsvi_ragged rows with S different target vectors (subjects) and the case's first noise draw repeated."""
    g = gu.load("dsvi_ragged")
    rng = np.random.default_rng(5)
    ys = np.stack([g["y"]] + [g["y"] + 0.3 * rng.standard_normal(g["y"].shape) for _ in range(S - 1)])
    reps = lambda a: np.concatenate([a[:1]] * S)
    return g, ys, reps(g["z_v"]), reps(g["z_ell"]), reps(g["z_L"])


def test_per_subject_targets_match_the_mean_of_reference_forwards(spec_ops):
    """y [S, B] (HCP-style step): loss and gradients equal the mean over subjects of one-draw oracle forwards."""
    g, ys, zv, zell, zL = _subject_case()
    S = ys.shape[0]
    p = gu.case_params(g)
    I = torch.from_numpy(g["I"]).to(torch.int32)
    loss, grads = dsvi_step.dsvi_step(p, torch.from_numpy(g["Z"]), torch.from_numpy(g["x"]), torch.from_numpy(ys), I,
                                      int(g["N"]), torch.from_numpy(zv), torch.from_numpy(zell), torch.from_numpy(zL))
    Xl, _ = gu.case_lists(g)
    D = int(g["D"])
    ref_loss, ref_grads = 0.0, None
    for s in range(S):
        Yl = [torch.from_numpy(ys[s][g["I"] == d]).view(-1, 1) for d in range(D)]
        l_s, g_s = orc.step_loss_and_grads(p, torch.from_numpy(g["Z"]).view(-1, 1), int(g["N"]), Xl, Yl,
                                           draws=gu.replay_draws(g)[:1])
        ref_loss += float(l_s) / S
        ref_grads = {k: (v / S if v is not None else None) for k, v in g_s.items()} if ref_grads is None else \
            {k: (ref_grads[k] + v / S if v is not None else None) for k, v in g_s.items()}
    assert abs(float(loss) - ref_loss) <= 1e-10 * abs(ref_loss)
    for k, gr in ref_grads.items():
        if gr is None:
            continue
        den = max(float(torch.linalg.norm(gr)), 1e-300)
        assert float(torch.linalg.norm(grads[k].reshape(-1) - gr.reshape(-1))) / den <= 1e-9, k


def test_sample_sharding_sums_to_the_unsharded_step(spec_ops):
    """Two ranks each holding all rows and half of the subjects (S_total / sample_offset, KL pairs split by kl_shard):
    the sum of (loss, gradients) over the ranks equals the single-rank step."""
    g, ys, zv, zell, zL = _subject_case(S=4)
    p = gu.case_params(g)
    I = torch.from_numpy(g["I"]).to(torch.int32)
    args = lambda sl: (p, torch.from_numpy(g["Z"]), torch.from_numpy(g["x"]), torch.from_numpy(ys[sl]), I, int(g["N"]),
                       torch.from_numpy(zv[sl]), torch.from_numpy(zell[sl]), torch.from_numpy(zL[sl]))
    full_loss, full_grads = dsvi_step.dsvi_step(*args(slice(0, 4)))
    tot, acc = 0.0, None
    for rank in range(2):
        l_r, g_r = dsvi_step.dsvi_step(*args(slice(2 * rank, 2 * rank + 2)), kl_weight=0.5, kl_shard=(rank, 2),
                                       S_total=4, sample_offset=2 * rank)
        tot += float(l_r)
        acc = g_r if acc is None else {k: acc[k] + g_r[k] for k in g_r}
    assert abs(tot - float(full_loss)) <= 1e-12 * abs(float(full_loss))
    for k in full_grads:
        den = max(float(torch.linalg.norm(full_grads[k])), 1e-300)
        assert float(torch.linalg.norm(acc[k] - full_grads[k])) / den <= 1e-10, k


def test_exact_kl_flag_matches_the_oracle_with_the_true_kl(spec_ops):
    """exact_kl=True (not the reference's behaviour: quirk q10) against the oracle's exact-KL variant."""
    g = gu.load("dsvi_ragged")
    p = gu.case_params(g)
    I = torch.from_numpy(g["I"]).to(torch.int32)
    loss, grads = dsvi_step.dsvi_step(p, torch.from_numpy(g["Z"]), torch.from_numpy(g["x"]), torch.from_numpy(g["y"]), I,
                                      int(g["N"]), torch.from_numpy(g["z_v"]), torch.from_numpy(g["z_ell"]),
                                      torch.from_numpy(g["z_L"]), exact_kl=True)
    Xl, Yl = gu.case_lists(g)
    ref_loss, ref_grads = orc.step_loss_and_grads(p, torch.from_numpy(g["Z"]).view(-1, 1), int(g["N"]), Xl, Yl,
                                                  draws=gu.replay_draws(g), exact_kl=True)
    assert abs(float(loss) - float(g["loss"])) > 1e-6 * abs(float(g["loss"]))          # it IS a different objective
    assert abs(float(loss) - float(ref_loss)) <= 1e-10 * abs(float(ref_loss))
    for k, gr in ref_grads.items():
        if gr is None:
            continue
        den = max(float(torch.linalg.norm(gr)), 1e-300)
        assert float(torch.linalg.norm(grads[k].reshape(-1) - gr.reshape(-1))) / den <= 1e-9, k

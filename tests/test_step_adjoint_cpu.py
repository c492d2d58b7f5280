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
        p, torch.from_numpy(g["Z"]), torch.from_numpy(g["x"]), torch.from_numpy(g["y"]), I, int(g["N"]),
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

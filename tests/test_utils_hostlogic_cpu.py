"""Host logic of the drop-in utils.* wrappers on the CPU (C-ABI wrappers replaced by their specifications)."""
import inspect

import pytest

from oracle import kernel_specs as specs
from tests import utils_cases
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops


@pytest.fixture(autouse=True)
def spec_ops(monkeypatch):
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            monkeypatch.setattr(_ops, n, f)


def test_utils_primitives_forward_and_autograd():
    utils_cases.run_all("cpu")

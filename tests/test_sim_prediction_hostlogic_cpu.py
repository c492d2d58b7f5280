"""Host logic of prediction.point_predmap / pointwise_predmap / test_predmap on the CPU against the reference's golden
outputs (C-ABI wrappers replaced by their specifications)."""
import inspect

import pytest

from oracle import kernel_specs as specs
from tests import sim_prediction_cases
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops


@pytest.fixture(autouse=True)
def spec_ops(monkeypatch):
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            monkeypatch.setattr(_ops, n, f)


def test_map_prediction_matches_reference():
    print(sim_prediction_cases.run_all("cpu"))

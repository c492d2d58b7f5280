"""Host logic of the drop-in API on the CPU: the C-ABI wrappers are replaced by their CPU specifications
(test infrastructure only), so what is exercised here is row regrouping, noise draw order, the
autograd.Function plumbing, Adam/DataLoader sequencing and checkpoint-key compatibility."""
import inspect

import pytest

from oracle import kernel_specs as specs
from tests import api_cases
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops


@pytest.fixture(autouse=True)
def spec_ops(monkeypatch):
    for n, f in inspect.getmembers(specs, inspect.isfunction):
        if not n.startswith("_"):
            monkeypatch.setattr(_ops, n, f)


def test_predict_modelpt():
    api_cases.predict_modelpt("cpu")


def test_forward_backward_reference_noise():
    api_cases.forward_backward_reference_noise("cpu")


def test_unsorted_index():
    api_cases.unsorted_index_matches_sorted("cpu")


def test_inference_trace():
    api_cases.inference_trace("cpu")


def test_compute_elbo_replay():
    api_cases.compute_elbo_replay("cpu")


def test_posterior_sampling_replay():
    api_cases.posterior_sampling_replay("cpu")

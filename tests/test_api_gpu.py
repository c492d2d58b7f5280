"""Host API (drop-in nmgp_dsvi) on the GPU, through the C ABI."""
import pytest

pytestmark = pytest.mark.gpu
from tests import api_cases  # noqa: E402

DEV = "cuda:0"


def test_predict_modelpt():
    api_cases.predict_modelpt(DEV)


def test_forward_backward_reference_noise():
    api_cases.forward_backward_reference_noise(DEV)


def test_unsorted_index():
    api_cases.unsorted_index_matches_sorted(DEV)


def test_inference_trace():
    api_cases.inference_trace(DEV)


def test_compute_elbo_replay():
    api_cases.compute_elbo_replay(DEV)


def test_posterior_sampling_replay():
    api_cases.posterior_sampling_replay(DEV)

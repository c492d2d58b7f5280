"""prediction.point_predmap / pointwise_predmap / test_predmap on the GPU (through the C ABI) against the reference."""
import pytest

pytestmark = pytest.mark.gpu
from tests import sim_prediction_cases  # noqa: E402


def test_map_prediction_matches_reference():
    print(sim_prediction_cases.run_all("cuda:0"))

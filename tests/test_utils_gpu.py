"""Drop-in utils.* primitives on the GPU (through the C ABI)."""
import pytest

pytestmark = pytest.mark.gpu
from tests import utils_cases  # noqa: E402


def test_utils_primitives_forward_and_autograd():
    utils_cases.run_all("cuda:0")

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


def test_graphed_step_replays_the_eager_iteration():
    """A whole iteration replayed from a CUDA graph (graph_step.GraphedStep) follows the eager loop: same losses and
    same parameters after several Adam steps with the device-keyed noise (fresh noise on every replay)."""
    import numpy as np
    import torch
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import nmgp_dsvi
    from collaborative_nonstationary_multivariate_gaussian_process_b200.graph_step import GraphedStep
    dev = torch.device("cuda:0")
    D, Q, T = 3, 20, 40
    rng = np.random.default_rng(0)
    x = torch.from_numpy(np.tile(np.linspace(0, 1, T), D)).to(dev)
    y = torch.from_numpy(rng.standard_normal(T * D)).to(dev)
    I = torch.from_numpy(np.repeat(np.arange(D, dtype=np.int32), T)).to(dev)

    def make():
        m = nmgp_dsvi.NMGP(T * D, D, torch.linspace(0, 1, Q, dtype=torch.float64).view(-1, 1), mu_v=-1.0 * np.ones(Q),
                           seed=7, device=dev, noise="device")
        for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
            getattr(m, k).data.fill_(-1.0)
        return m, torch.optim.Adam(m.parameters(), lr=0.01, capturable=True)
    m1, o1 = make()
    eager = []
    for _ in range(3 + 4):                                   # GraphedStep warms up with 3 real steps before capturing
        o1.zero_grad(set_to_none=True)
        loss = m1.forward_rows(x, y, I, n_mc=2)
        loss.backward()
        o1.step()
        eager.append(float(loss))
    m2, o2 = make()
    gs = GraphedStep(m2, o2, x, y, I, n_mc=2, warmup=3)
    # capture itself does not execute; replays continue from the state after the warm-up steps
    got = [float(gs.step()) for _ in range(4)]
    gs.check()
    assert gs.kernels_per_replay > 50
    for a, b in zip(got, eager[3:]):
        assert abs(a - b) <= 1e-9 * abs(b), (got, eager)
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert float(torch.linalg.norm(p1 - p2)) <= 1e-9 * max(float(torch.linalg.norm(p1)), 1e-300), k

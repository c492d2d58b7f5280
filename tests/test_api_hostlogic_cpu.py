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


def test_device_minibatches_follow_the_reference_loader_order():
    """DeviceMinibatches(order="reference") must see the batches DataLoader(shuffle=True) would produce and leave the
    global CPU generator in the same state (quirk q9: the reference's loss trace depends on it)."""
    import numpy as np
    import torch
    from torch.utils.data import DataLoader, TensorDataset
    from collaborative_nonstationary_multivariate_gaussian_process_b200.nmgp_dsvi import DeviceMinibatches
    n, bs, D = 23, 5, 3
    rng = np.random.default_rng(0)
    X = rng.standard_normal(n); Y = rng.standard_normal(n); I = np.sort(rng.integers(0, D, n))
    torch.manual_seed(5)
    dl = DataLoader(TensorDataset(torch.arange(n)), batch_size=bs, shuffle=True)
    want = []
    for ep in range(2):
        for (b,) in dl:
            idx = b.numpy()
            grp = np.argsort(I[idx], kind="stable")           # vec2list regrouping (code/nmgp_dsvi.py:745-755)
            want.append((X[idx[grp]], I[idx[grp]]))
        want.append(float(torch.randn(1)))
    torch.manual_seed(5)
    mb = DeviceMinibatches(X, Y, I, D, bs, "cpu", order="reference")
    got = []
    for ep in range(2):
        for xb, yb, Ib, Ih in mb.epoch():
            got.append((xb.numpy(), Ib.numpy().astype(np.int64)))
            assert np.array_equal(Ih, Ib.numpy())
        got.append(float(torch.randn(1)))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        if isinstance(a, float):
            assert a == b
        else:
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the device order covers every row exactly once per epoch, grouped by output
    mb2 = DeviceMinibatches(X, Y, I, D, bs, "cpu", order="device", seed=3)
    seen = np.concatenate([xb.numpy() for xb, _, Ib, _ in mb2.epoch()])
    assert sorted(seen.tolist()) == sorted(X.tolist())
    assert all(bool((Ib[1:] >= Ib[:-1]).all()) for _, _, Ib, _ in mb2.epoch())


def test_whole_module_pickle_round_trip():
    """The reference's drivers pickle the whole nn.Module next to the loss trace (NMGP_PM25.py:101-106,
    NMGP_ECoG_full.py:169-174): the drop-in model must survive pickle.dump / pickle.load with its parameters, its
    inducing grid Z (a plain attribute, not in the state dict) and its shapes, and keep the reference's state-dict keys."""
    import io
    import pickle
    import numpy as np
    import torch
    from collaborative_nonstationary_multivariate_gaussian_process_b200.nmgp_dsvi import NMGP
    Z = np.linspace(0.0, 1.0, 6)
    m = NMGP(number_observations=40, dim_outputs=3, Z=Z, minibatch_size=10, seed=3, device="cpu")
    buf = io.BytesIO()
    pickle.dump([m, [1.0, 2.0], [0.1, 0.2]], buf)               # [model, loss_list, time_list] as the drivers write it
    m2, losses, times = pickle.loads(buf.getvalue())
    assert losses == [1.0, 2.0] and times == [0.1, 0.2]
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd) == list(sd2)
    assert set(sd) >= {"mu_W", "sqrt_W", "mu_v", "sqrt_v", "mu_U", "sqrt_U", "sigma2_err_log"}
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k
    assert torch.equal(torch.as_tensor(m.Z), torch.as_tensor(m2.Z))
    assert (m2.N, m2.D, m2.M) == (m.N, m.D, m.M)

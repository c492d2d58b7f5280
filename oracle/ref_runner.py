"""Runs the UNMODIFIED reference (vendored byte for byte into oracle/_ref/ by oracle/build_ref.py) for timing.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by ``bench.py --impl reference`` and its ``cpu_baseline`` leg.
The torch>=2 compatibility shim of SURVEY.md 8c (stub matplotlib, ``torch.solve``, ``torch.symeig``) is installed at
import time; the reference files themselves are never edited.

One *reference step* is exactly the loop body of code/nmgp_dsvi.py:832-854::

    optimizer.zero_grad(); loss = model(X_list, Y_list); loss.backward(retain_graph=True); optimizer.step()

S > 1 has no code path in the reference: S Monte-Carlo samples are S such forward/backward calls on unchanged
parameters, so the time of one S-sample iteration is S times the time of one call.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CODE = os.path.join(HERE, "_ref", "code")
_mod = {}


def available():
    return os.path.exists(os.path.join(REF_CODE, "nmgp_dsvi.py"))


def install_shim():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(torch, "solve") or getattr(torch.solve, "__name__", "") != "_nmgp_shim_solve":
        def _nmgp_shim_solve(input, A):
            return torch.linalg.solve(A, input), None
        torch.solve = _nmgp_shim_solve
    torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO="U")


def reference_module():
    """The reference's ``nmgp_dsvi`` module, imported from oracle/_ref/code."""
    if "nmgp_dsvi" not in _mod:
        if not available():
            raise RuntimeError("oracle/_ref is not built (run `python oracle/build_ref.py` where /root/reference exists)")
        install_shim()
        import importlib.util
        import warnings
        warnings.filterwarnings("ignore")
        sys.path.insert(0, REF_CODE)
        try:
            spec = importlib.util.spec_from_file_location("_nmgp_ref_nmgp_dsvi", os.path.join(REF_CODE, "nmgp_dsvi.py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
        finally:
            sys.path.remove(REF_CODE)
            # the reference's `utils` must not shadow anything else that imports a module of that name later
            sys.modules.pop("utils", None)
        _mod["nmgp_dsvi"] = m
    return _mod["nmgp_dsvi"]


def sim_modules():
    """(kernels, kronecker_operation, distributions) of the reference's SIM_code line."""
    if "sim" not in _mod:
        install_shim()
        p = os.path.join(REF_CODE, "SIM_code")
        sys.path.insert(0, p)
        try:
            import importlib
            for k in [k for k in sys.modules if k == "Utility" or k.startswith("Utility.")]:
                del sys.modules[k]
            ker = importlib.import_module("Utility.kernels")
            kro = importlib.import_module("Utility.kronecker_operation")
            dis = importlib.import_module("Utility.distributions")
        finally:
            sys.path.remove(p)
        _mod["sim"] = (ker, kro, dis)
    return _mod["sim"]


def make_step(T, D, Q, rows, hyper, seed=0, lr=0.005, mu_v_one=True):
    """Reference model on the T x D shared grid (X_d = arange(T), Z = linspace(0, T-1, Q), NMGP(seed=22, mu_v=1), driver
    hyper-parameters, frozen length-scales) and a closure running one reference step on ``rows`` random grid rows."""
    ref = reference_module()
    rows = min(int(rows), T * D)
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(T * D, size=rows, replace=False))
    Xl = [torch.from_numpy((pick[(pick // T) == d] % T).astype(np.float64)).view(-1, 1) for d in range(D)]
    Yl = [torch.from_numpy(rng.standard_normal(x.shape[0])).view(-1, 1) for x in Xl]
    Z = torch.linspace(0, T - 1, Q, dtype=torch.float64).view(-1, 1)
    model = ref.NMGP(T * D, D, Z, mu_v=np.ones(Q) if mu_v_one else None, seed=22)
    for k, v in hyper.items():
        getattr(model, k).data.fill_(float(v))
    for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
        getattr(model, k).requires_grad = False                      # fix_hyperpars=True (nmgp_dsvi.py:795-814)
    opt = torch.optim.Adam(model.parameters(), lr=lr)

    def step():
        opt.zero_grad()
        loss = model(Xl, Yl)
        loss.backward(retain_graph=True)
        opt.step()
        return float(loss.detach())
    return step, rows

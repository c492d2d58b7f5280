"""Shared cases of the SIM_code log-posterior tests: run on the CPU with the C-ABI wrappers replaced by their
specifications (test_sim_logpos_hostlogic_cpu) and on the GPU through the C ABI (test_sim_logpos_gpu).
Golden values AND gradients come from the unmodified reference (oracle/gen_golden_logpos.py: its own autograd)."""
import numpy as np
import torch

from tests import golden_util as gu

VTOL = 1e-9       # north_star tolerance on log-densities
GTOL = 1e-9       # norm-wise on gradient vectors


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1); b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check_objective_gradients(dev):
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import logpos
    g = gu.load("sim_logpos")
    d = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.float64)).to(dev)
    sc = lambda v: torch.tensor(float(v), dtype=torch.float64, device=dev)
    hyp = [float(v) for v in g["hyp"]]
    a, b, c = (float(v) for v in g["abc"])
    ts2 = sc(g["ts2"])
    out = {}
    # Kronecker NMGP posterior
    pars = torch.cat([d("tilde_l"), d("tilde_sigma"), d("uL_vec"), ts2.view(1)]).requires_grad_(True)
    val = logpos.nlogpos_obj(pars, d("Y"), d("x"), *hyp, a, b, c)
    val.backward()
    out["nlogpos_obj"] = (abs(float(val) - float(g["nlogpos_obj"])) / abs(float(g["nlogpos_obj"])),
                          _rel(pars.grad.cpu().numpy(), g["grad_nlogpos_obj"]))
    # deviance
    pd_ = torch.cat([d("tilde_l"), d("tilde_sigma"), d("L_vec"), ts2.view(1)]).requires_grad_(True)
    val = logpos.deviance_obj(pd_, d("Y"), d("x"))
    val.backward()
    out["deviance_obj"] = (abs(float(val) - float(g["deviance"])) / abs(float(g["deviance"])),
                           _rel(pd_.grad.cpu().numpy(), g["grad_deviance_obj"]))
    # stationary variant
    pS = torch.cat([sc(g["tlS"]).view(1), sc(g["tsS"]).view(1), d("uL_vec"), ts2.view(1)]).requires_grad_(True)
    val = logpos.nlogpos_obj_S(pS, d("Y"), d("x"), sc(-1.0), sc(0.7), a, b, c)
    val.backward()
    out["nlogpos_obj_S"] = (abs(float(val) - float(g["nlogpos_obj_S"])) / abs(float(g["nlogpos_obj_S"])),
                            _rel(pS.grad.cpu().numpy(), g["grad_nlogpos_obj_S"]))
    for k, (ev, eg) in out.items():
        assert ev <= VTOL, (k, "value", ev)
        assert eg <= GTOL, (k, "gradient", eg)
    return out

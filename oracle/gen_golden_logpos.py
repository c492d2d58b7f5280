"""Golden values of the SIM_code log-posteriors / deviance (code/SIM_code/Utility/logpos.py) from the UNMODIFIED reference
under the shim of oracle/gen_golden.py.  TEST INFRASTRUCTURE ONLY:  python oracle/gen_golden_logpos.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, install_shim  # noqa: E402


def main():
    install_shim()
    from Utility import logpos, utils as sim_utils
    g = torch.Generator().manual_seed(23)
    N, M = 40, 3
    P = M * (M + 1) // 2
    x = torch.sort(torch.rand(N, generator=g).double())[0]
    tilde_l = (3 * (x - 1) ** 3 - 1.5) + 0.05 * torch.randn(N, generator=g).double()
    tilde_sigma = 0.1 * torch.randn(N, generator=g).double()
    uL_vec = 0.4 * torch.randn(P, generator=g).double()
    L_vec = sim_utils.uLvec2Lvec(uL_vec, M)
    ts2 = torch.tensor(np.log(2e-2)).double()
    Y = torch.randn(N, M, generator=g).double()
    hyp = [torch.tensor(v).double() for v in (-1.0, 1.5, 0.3, 0.0, 0.8, 0.4)]
    a, b, c = 1.0, 1.0, 10.0
    out = {}
    v = logpos.logpos(tilde_l, tilde_sigma, uL_vec, ts2, Y, x, *hyp, a, b, c, verbose=True)
    out["logpos_verbose"] = np.array([float(t) for t in v])
    out["logpos_noprior"] = float(logpos.logpos(tilde_l, tilde_sigma, uL_vec, ts2, Y, x, *hyp, a, b, c, Prior=False))
    pars = torch.cat([tilde_l, tilde_sigma, uL_vec, ts2.view(1)])
    out["nlogpos_obj"] = float(logpos.nlogpos_obj(pars, Y, x, *[float(h) for h in hyp], a, b, c))
    out["deviance"] = float(logpos.deviance(tilde_l, tilde_sigma, L_vec, ts2, Y, x))
    tlS, tsS = torch.tensor(-1.7).double(), torch.tensor(0.15).double()
    vS = logpos.logpos_S(tlS, tsS, uL_vec, ts2, Y, x, torch.tensor(-1.0).double(), torch.tensor(0.7).double(), a, b, c, verbose=True)
    out["logpos_S_verbose"] = np.array([float(t) for t in vS])
    # Hadamard layout
    keep = torch.rand(N, M, generator=g) < 0.6
    xh = torch.cat([x[keep[:, m]] for m in range(M)])
    ih = torch.cat([torch.full((int(keep[:, m].sum()),), m, dtype=torch.long) for m in range(M)])
    yh = torch.cat([Y[keep[:, m], m] for m in range(M)])
    tlh = (3 * (xh - 1) ** 3 - 1.5) + 0.05 * torch.randn(xh.numel(), generator=g).double()
    tsh = 0.1 * torch.randn(xh.numel(), generator=g).double()
    vH = logpos.logpos_hadamard(tlh, tsh, L_vec, ts2, xh, ih, yh, *hyp, a, b, c, verbose=True)
    out["logpos_hadamard_verbose"] = np.array([float(t) for t in vH])
    vHS = logpos.logpos_hadamard_S(tlS, tsS, L_vec, ts2, xh, ih, yh, torch.tensor(-1.0).double(), torch.tensor(0.7).double(), a, b, c, verbose=True)
    out["logpos_hadamard_S_verbose"] = np.array([float(t) for t in vHS])
    parsH = torch.cat([tlh, tsh, L_vec, ts2.view(1)])
    out["nlogpos_obj_hadamard"] = float(logpos.nlogpos_obj_hadamard(parsH, xh, ih, yh, *[float(h) for h in hyp], a, b, c))
    # spatially varying coregionalisation
    Ni, Mi = 20, 2
    Pi = Mi * (Mi + 1) // 2
    xi = torch.sort(torch.rand(Ni, generator=g).double())[0]
    tli = (3 * (xi - 1) ** 3 - 1.5) + 0.05 * torch.randn(Ni, generator=g).double()
    uLi = 0.3 * torch.randn(Ni * Pi, generator=g).double()
    Yi = torch.randn(Ni, Mi, generator=g).double()
    hyp_i = [torch.tensor(v_).double() for v_ in (-1.0, 1.5, 0.3, 0.1, 0.7, 0.35)]
    vSVC = logpos.logpos_SVC(tli, uLi, ts2, Yi, xi, *hyp_i, a, b, verbose=True)
    out["logpos_SVC_verbose"] = np.array([float(t) for t in vSVC])
    out["nlogpos_obj_SVC"] = float(logpos.nlogpos_obj_SVC(torch.cat([tli, uLi, ts2.view(1)]), Yi, xi, *[float(h) for h in hyp_i], a, b))
    Nh = xh.numel()
    Lv_h = 0.4 * torch.randn(Nh * P, generator=g).double() + 0.3
    vHSVC = logpos.logpos_hadamard_SVC(tlh, Lv_h, ts2, xh, ih, yh, *hyp_i, a, b, verbose=True)
    out["logpos_hadamard_SVC_verbose"] = np.array([float(t) for t in vHSVC])
    # gradients of the objectives w.r.t. the parameter vector: the reference's own autograd (through torch.symeig ->
    # linalg.eigh under the shim, torch.distributions, ...).  These pin the hand-written SIM_code adjoints.
    pg = pars.clone().requires_grad_(True)
    logpos.nlogpos_obj(pg, Y, x, *[float(h) for h in hyp], a, b, c).backward()
    out["grad_nlogpos_obj"] = pg.grad.numpy().copy()
    pd_ = torch.cat([tilde_l, tilde_sigma, L_vec, ts2.view(1)]).clone().requires_grad_(True)
    logpos.deviance_obj(pd_, Y, x).backward()
    out["grad_deviance_obj"] = pd_.grad.numpy().copy()
    pS = torch.cat([tlS.view(1), tsS.view(1), uL_vec, ts2.view(1)]).clone().requires_grad_(True)
    vSo = logpos.nlogpos_obj_S(pS, Y, x, torch.tensor(-1.0).double(), torch.tensor(0.7).double(), a, b, c)
    vSo.backward()
    out["nlogpos_obj_S"] = float(vSo)
    out["grad_nlogpos_obj_S"] = pS.grad.numpy().copy()
    # dense (Hadamard / spatially-varying-coregionalisation) objectives: value + reference autograd gradient
    fh = [float(h) for h in hyp]
    fhi = [float(h) for h in hyp_i]
    pH = parsH.clone().requires_grad_(True)
    logpos.nlogpos_obj_hadamard(pH, xh, ih, yh, *fh, a, b, c).backward()
    out["grad_nlogpos_obj_hadamard"] = pH.grad.numpy().copy()
    pHS = torch.cat([tlS.view(1), tsS.view(1), L_vec, ts2.view(1)]).clone().requires_grad_(True)
    vHSo = logpos.nlogpos_obj_hadamard_S(pHS, xh, ih, yh, torch.tensor(-1.0).double(), torch.tensor(0.7).double(), a, b, c)
    vHSo.backward()
    out["nlogpos_obj_hadamard_S"] = float(vHSo)
    out["grad_nlogpos_obj_hadamard_S"] = pHS.grad.numpy().copy()
    pSVC = torch.cat([tli, uLi, ts2.view(1)]).clone().requires_grad_(True)
    logpos.nlogpos_obj_SVC(pSVC, Yi, xi, *fhi, a, b).backward()
    out["grad_nlogpos_obj_SVC"] = pSVC.grad.numpy().copy()
    pHSVC = torch.cat([tlh, Lv_h, ts2.view(1)]).clone().requires_grad_(True)
    vHSVCo = logpos.nlogpos_obj_hadamard_SVC(pHSVC, xh, ih, yh, *fhi, a, b)
    vHSVCo.backward()
    out["nlogpos_obj_hadamard_SVC"] = float(vHSVCo)
    out["grad_nlogpos_obj_hadamard_SVC"] = pHSVC.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "sim_logpos.npz"), xi=xi.numpy(), tli=tli.numpy(), uLi=uLi.numpy(), Yi=Yi.numpy(),
                        hyp_i=np.array([float(h) for h in hyp_i]), Lv_h=Lv_h.numpy(), x=x.numpy(), tilde_l=tilde_l.numpy(), tilde_sigma=tilde_sigma.numpy(),
                        uL_vec=uL_vec.numpy(), L_vec=L_vec.numpy(), ts2=float(ts2), Y=Y.numpy(), hyp=np.array([float(h) for h in hyp]),
                        abc=np.array([a, b, c]), tlS=float(tlS), tsS=float(tsS), xh=xh.numpy(), ih=ih.numpy(), yh=yh.numpy(),
                        tlh=tlh.numpy(), tsh=tsh.numpy(), **out)
    print({k: (v if np.ndim(v) == 0 else v[:2]) for k, v in out.items()})


if __name__ == "__main__":
    main()

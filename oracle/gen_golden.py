"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only at
/root/reference) under the 3-patch torch>=2 shim of SURVEY.md 8c.

TEST INFRASTRUCTURE ONLY.  Run once in the build container:

    python oracle/gen_golden.py

The GPU box has no /root/reference; tests there read only the committed .npz files.
Large gradient tensors are stored as (norm, seeded random projection, strided
sample) triplets instead of in full; parameters of the larger cases are
re-created from their seed by ``oracle.nmgp_oracle.init_params`` and pinned by
checksums stored next to them.
"""
import os
import pickle
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
FULL_LIMIT = 4096      # store tensors up to this many elements in full


def install_shim():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    torch.solve = lambda input, A: (torch.linalg.solve(A, input), None)
    torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO="U")
    sys.path.insert(0, os.path.join(REF, "code"))
    sys.path.insert(0, os.path.join(REF, "code", "SIM_code"))


def proj_vector(n, tag):
    rng = np.random.default_rng(abs(hash(tag)) % (2 ** 32))
    return rng.standard_normal(n)


def stable_tag_seed(tag: str) -> int:
    import zlib
    return zlib.crc32(tag.encode())


def summarize(name, arr, out):
    a = np.asarray(arr, dtype=np.float64)
    if a.size <= FULL_LIMIT:
        out[name] = a
    else:
        flat = a.reshape(-1)
        rng = np.random.default_rng(stable_tag_seed(name))
        out[name + "__norm"] = np.linalg.norm(flat)
        out[name + "__proj"] = np.array([flat @ rng.standard_normal(flat.size) for _ in range(4)])
        idx = rng.choice(flat.size, size=512, replace=False)
        out[name + "__idx"] = idx
        out[name + "__val"] = flat[idx]


class Recorder:
    """Wraps torch.randn to log every draw of one reference call."""

    def __init__(self):
        self.log = []
        self._orig = torch.randn

    def __enter__(self):
        def rec(*a, **k):
            t = self._orig(*a, **k)
            self.log.append(t.clone())
            return t
        torch.randn = rec
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig


HYPER_KEYS = ("sigma2_tildeell_log", "length_scales_tildeell_log", "sigma2_L0_log",
              "length_scales_L0_log", "sigma2_L1_log", "length_scales_L1_log", "sigma2_err_log")


def dsvi_case(nmgp_dsvi, name, X_list, Y_list, z, N, hyper, seed, init=None, train_len=False,
              store_params=True, n_forward=1, Y_lists=None):
    """Y_lists (one Y_list per forward): the forwards see different targets on the same inputs -- the subjects of an
    HCP-shaped step; the recorded loss / gradients are those of the mean over the forwards."""
    D = len(X_list)
    Q = z.shape[0]
    T = torch.DoubleTensor
    Z = torch.from_numpy(z).type(T).unsqueeze(1)
    init = init or {}
    model = nmgp_dsvi.NMGP(N, D, Z, minibatch_size=None, seed=seed, **init)
    for k, v in hyper.items():
        getattr(model, k).data.fill_(v)
    if not train_len:
        for k in ("length_scales_tildeell_log", "length_scales_L0_log", "length_scales_L1_log"):
            getattr(model, k).requires_grad = False
    Xl = [torch.from_numpy(np.asarray(x)).type(T).view(-1, 1) for x in X_list]
    Yl = [torch.from_numpy(np.asarray(y)).type(T).view(-1, 1) for y in Y_list]
    I = np.hstack([np.repeat(j, x.shape[0]) for j, x in enumerate(Xl)]).astype(np.int64)
    B = I.shape[0]
    torch.manual_seed(1000 + seed)
    losses = []
    zs = []
    total = 0
    Yls = None if Y_lists is None else [[torch.from_numpy(np.asarray(y)).type(T).view(-1, 1) for y in Yl_s]
                                        for Yl_s in Y_lists]
    for s in range(n_forward):
        with Recorder() as rec:
            loss = model(Xl, Yl if Yls is None else Yls[s])
        total = total + loss
        losses.append(float(loss.detach()))
        log = rec.log
        assert len(log) == 2 + D * (D + 1) // 2
        zL = np.zeros((B, D))
        k = 2
        for i in range(D):
            for j in range(i + 1):
                zz = log[k].numpy().astype(np.float64)
                zL[I == i, j] = zz[I == i]
                k += 1
        zs.append((log[0].numpy().astype(np.float64), log[1].numpy().astype(np.float64), zL))
    (total / n_forward).backward()
    out = dict(name=name, D=D, Q=Q, B=B, N=N, seed=seed, train_len=int(train_len), n_forward=n_forward,
               x=torch.cat(Xl).view(-1).numpy(), y=torch.cat(Yl).view(-1).numpy(), I=I, Z=z,
               loss=float((total / n_forward).detach()), losses=np.array(losses),
               z_v=np.stack([a for a, _, _ in zs]), z_ell=np.stack([b for _, b, _ in zs]),
               z_L=np.stack([c for _, _, c in zs]), store_params=int(store_params))
    if Yls is not None:
        out["ys"] = np.stack([torch.cat(Yl_s).view(-1).numpy() for Yl_s in Yls])
    for k, v in init.items():
        out["init_" + k] = np.asarray(v)
    sd = model.state_dict()
    for k in sd:
        a = sd[k].detach().numpy()
        if store_params or a.size <= FULL_LIMIT:
            out["param_" + k] = a
        out["paramsum_" + k] = np.array([a.sum(), (a.astype(np.float64) ** 2).sum()])
    for k, prm in model.named_parameters():
        if prm.grad is not None:
            summarize("grad_" + k, prm.grad.numpy(), out)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", out["loss"], "B", B)
    return model


def main():
    os.makedirs(OUT, exist_ok=True)
    install_shim()
    import nmgp_dsvi
    from Utility import kernels, kronecker_operation, distributions
    torch.set_num_threads(8)

    sim_hyper = {"sigma2_L0_log": 0., "length_scales_L0_log": 2., "sigma2_L1_log": 0.,
                 "length_scales_L1_log": 2., "sigma2_tildeell_log": 0., "length_scales_tildeell_log": 0.,
                 "sigma2_err_log": -2.}
    # ---- config 1: the three shipped pickles, one full-batch step each ----------------
    for tag in ("low", "high", "varying"):
        with open(os.path.join(REF, "data", "simulation", "sim_illustration_%s_freq.pickle" % tag), "rb") as f:
            X_list, Y_list, Xt_list, Yt_list = pickle.load(f)
        z = np.linspace(0, 1, 20)
        dsvi_case(nmgp_dsvi, "dsvi_sim_%s" % tag, X_list, Y_list, z, 200, sim_hyper, seed=0)

    # ---- inference(): first losses of the seed-0 low_freq run (NMGP_SIM.ipynb settings) ---
    with open(os.path.join(REF, "data", "simulation", "sim_illustration_low_freq.pickle"), "rb") as f:
        X_list, Y_list, Xt_list, Yt_list = pickle.load(f)
    z = np.linspace(0, 1, 20)
    # NB quirk q3: hyperpars['sigma2_L1_log'] lands in sigma2_L0_log
    model, loss_list, time_list = nmgp_dsvi.inference(
        X_list, Y_list, z, 200, 2, hyperpars=dict(sim_hyper), lr=0.005, itnum=5, seed=0, show_ELBO=False)
    pred = nmgp_dsvi.predict_Y(model, Xt_list)
    np.savez_compressed(os.path.join(OUT, "inference_sim_low.npz"),
                        losses=np.array([float(l) for l in loss_list]), pred_after=pred,
                        X=np.concatenate(X_list).reshape(-1), Y=np.concatenate(Y_list).reshape(-1),
                        n_per_output=np.array([len(x) for x in X_list]),
                        Xt=np.concatenate(Xt_list).reshape(-1), nt_per_output=np.array([len(x) for x in Xt_list]))
    print("inference losses", [float(l) for l in loss_list])

    # ---- model.pt + predict_Y (deterministic known answer) -----------------------------
    ck = torch.load(os.path.join(REF, "code", "notebook", "model.pt"), weights_only=True)
    sd = ck["model_state_dict"]
    Z = torch.from_numpy(np.linspace(0, 1, 20)).type(torch.DoubleTensor).unsqueeze(1)
    m = nmgp_dsvi.NMGP(200, 2, Z)
    m.load_state_dict(sd)
    pred = nmgp_dsvi.predict_Y(m, Xt_list)
    Yt = np.concatenate(Yt_list)
    rmse = float(np.sqrt(np.mean((pred[:, None] - Yt) ** 2)))
    out = {"param_" + k: v.numpy() for k, v in sd.items()}
    out.update(pred=pred, rmse=rmse, Xt=np.concatenate(Xt_list).reshape(-1), Yt=Yt.reshape(-1),
               nt_per_output=np.array([len(x) for x in Xt_list]), Z=np.linspace(0, 1, 20))
    np.savez_compressed(os.path.join(OUT, "predict_modelpt.npz"), **out)
    print("model.pt predict", pred[:3], rmse)

    # ---- compute_ELBO with 3 draws on low_freq (quirk q5 included) ---------------------
    T = torch.DoubleTensor
    Xl = [torch.from_numpy(x).type(T) for x in X_list]
    Yl = [torch.from_numpy(y).type(T) for y in Y_list]
    torch.manual_seed(77)
    with Recorder() as rec:
        elbo = m.compute_ELBO(Xl, Yl, n_sample=3)
    np.savez_compressed(os.path.join(OUT, "elbo_modelpt.npz"), elbo=float(elbo),
                        noise=np.concatenate([t.numpy().astype(np.float64).reshape(-1) for t in rec.log]),
                        X=np.concatenate(X_list).reshape(-1), Y=np.concatenate(Y_list).reshape(-1),
                        n_per_output=np.array([len(x) for x in X_list]))
    print("compute_ELBO", float(elbo))

    # ---- posterior sampling: sample_Y / sample_FY with 2 draws -----------------------------
    Xs_small = [torch.from_numpy(Xt_list[0][:7]).type(T), torch.from_numpy(Xt_list[1][:5]).type(T)]
    torch.manual_seed(31)
    with Recorder() as rec:
        sYs, sLs, sGs, sEll = m.sample_Y(Xs_small, n_sample=2)
    np.savez_compressed(os.path.join(OUT, "sample_Y_modelpt.npz"), X=np.concatenate([a.numpy().reshape(-1) for a in Xs_small]),
                        n_per_output=np.array([7, 5]), Ys=sYs.numpy(), Ls=sLs.numpy(), Gs=sGs.numpy(), ells=sEll.numpy())
    m3 = nmgp_dsvi.NMGP(50, 3, torch.from_numpy(np.linspace(0, 1, 8)).type(T).unsqueeze(1), seed=4,
                        mu_v=-1.0 * np.ones(8))
    for k, v in {"length_scales_tildeell_log": -1.0, "length_scales_L0_log": -0.5, "length_scales_L1_log": -0.8,
                 "sigma2_err_log": -3.0}.items():
        getattr(m3, k).data.fill_(v)
    grid = torch.from_numpy(np.linspace(0.05, 0.95, 9)).type(T)
    torch.manual_seed(32)
    tE, tY, tC = m3.sample_FY(grid, n_sample=2)
    out3 = {"param_" + k: v.detach().numpy() for k, v in m3.state_dict().items()}
    out3.update(grid=grid.numpy(), ells=tE.numpy(), Ys=tY.numpy(), corrs=tC.numpy(), Z=np.linspace(0, 1, 8))
    np.savez_compressed(os.path.join(OUT, "sample_FY_d3.npz"), **out3)
    print("sample_Y/FY", sYs.shape, sLs.shape, sGs.shape, sEll.shape, tE.shape, tY.shape, tC.shape)

    # ---- small ragged case with an empty output, N != B, trainable length-scales, S=2 ---
    rng = np.random.default_rng(5)
    counts = [15, 0, 25, 9]
    Xs = [np.sort(rng.uniform(0, 1, c)) for c in counts]
    Ys = [5 * np.cos(2 * np.pi * 5 * x ** 2) * (0.5 + d) + rng.uniform(0, 1, x.shape) for d, x in enumerate(Xs)]
    hyper = {"sigma2_tildeell_log": -0.3, "length_scales_tildeell_log": -1.2, "sigma2_L0_log": 0.2,
             "length_scales_L0_log": -0.7, "sigma2_L1_log": -0.4, "length_scales_L1_log": -0.9,
             "sigma2_err_log": -1.5}
    dsvi_case(nmgp_dsvi, "dsvi_ragged", Xs, Ys, np.linspace(0, 1, 8), 137, hyper, seed=3,
              init={"mu_v": -1.5 * np.ones(8)}, train_len=True, n_forward=2)

    # ---- ECoG-like: shared grid T=800, D=8, Q=50, B=512 minibatch, driver hyper-parameters ---
    Tn, D, Q, B = 800, 8, 50, 512
    rng = np.random.default_rng(11)
    grid = np.arange(Tn, dtype=np.float64)
    Yfull = np.stack([np.sin(grid / 40.0 * (1 + 0.1 * d)) + 0.3 * rng.standard_normal(Tn) for d in range(D)])
    pick = np.sort(rng.choice(Tn * D, size=B, replace=False))
    Xs = [grid[pick[(pick // Tn) == d] % Tn] for d in range(D)]
    Ys = [Yfull[d][(pick[(pick // Tn) == d] % Tn)] for d in range(D)]
    hyper = {"length_scales_L0_log": 10., "length_scales_L1_log": 10., "length_scales_tildeell_log": 5.,
             "sigma2_err_log": -5.}
    dsvi_case(nmgp_dsvi, "dsvi_ecog_like", Xs, Ys, np.linspace(0, Tn - 1, Q), Tn * D, hyper, seed=22,
              init={"mu_v": np.ones(Q)}, store_params=False)

    # ---- PM2.5-like: D=6, Q=100, B=1000 (driver shape), len = e^10 ----------------------
    Tn, D, Q, B = 2048, 6, 100, 1000
    rng = np.random.default_rng(12)
    grid = np.arange(Tn, dtype=np.float64)
    Yfull = np.stack([np.cos(grid / 100.0 + d) + 0.2 * rng.standard_normal(Tn) for d in range(D)])
    pick = np.sort(rng.choice(Tn * D, size=B, replace=False))
    Xs = [grid[pick[(pick // Tn) == d] % Tn] for d in range(D)]
    Ys = [Yfull[d][(pick[(pick // Tn) == d] % Tn)] for d in range(D)]
    hyper = {"length_scales_L0_log": 10., "length_scales_L1_log": 10., "length_scales_tildeell_log": 10.}
    dsvi_case(nmgp_dsvi, "dsvi_pm25_like", Xs, Ys, np.linspace(0, Tn - 1, Q), Tn * D, hyper, seed=22,
              init={"mu_v": np.ones(Q)}, store_params=False)

    # ---- SIM_code line --------------------------------------------------------------------
    torch.manual_seed(4)
    T1, T2, Dm = 48, 31, 3
    x1 = torch.sort(torch.rand(T1).double())[0].view(-1, 1)
    x2 = torch.rand(T2).double().view(-1, 1)
    ell1 = torch.exp(3 * (x1.view(-1) - 1) ** 3 - 1.0); ell2 = torch.exp(0.3 * torch.randn(T2).double() - 1.5)
    sg1 = torch.exp(0.2 * torch.randn(T1).double()); sg2 = torch.exp(0.2 * torch.randn(T2).double())
    K_self = kernels.Nonstationary_RBF_cov(x1, sigma1=sg1, ell1=ell1)
    K_cross = kernels.Nonstationary_RBF_cov(x1, sigma1=sg1, ell1=ell1, X2=x2, sigma2=sg2, ell2=ell2)
    K_def = kernels.Nonstationary_RBF_cov(x1)
    R_self = kernels.RBF_cov(x1, alpha=1.3, beta=0.2)
    R_cross = kernels.RBF_cov(x1, x2, alpha=0.7, beta=0.35)
    Lb = torch.tril(torch.randn(Dm, Dm).double()); Bf = Lb @ Lb.t()
    yv = torch.randn(Dm * T1).double(); mu = 0.1 * torch.randn(Dm * T1).double()
    s2 = torch.tensor(1e-2).double()
    mv = kronecker_operation.kron_mv(Bf, K_self, yv)
    ld = kronecker_operation.kron_logdet(s2, Bf, K_self)
    inv = kronecker_operation.kron_inv(s2, Bf, K_self)
    kp = kronecker_operation.kronecker_product(Bf, K_self[:5, :4])
    kd = kronecker_operation.kronecker_product_diag(torch.diagonal(Bf), torch.diagonal(K_self))
    lp0 = distributions.multivariate_normal_logpdf0(yv, mu, Bf, K_self, s2)
    lp2 = distributions.multivariate_normal_logpdf2(yv, mu, Bf, K_self, s2)
    torch.manual_seed(9)
    rB = torch.rand(Dm); rK = torch.rand(T1)
    torch.manual_seed(9)
    lp1 = distributions.multivariate_normal_logpdf1(yv, mu, Bf, K_self, s2)
    np.savez_compressed(os.path.join(OUT, "sim_code.npz"), x1=x1.numpy().reshape(-1), x2=x2.numpy().reshape(-1),
                        ell1=ell1.numpy(), ell2=ell2.numpy(), sg1=sg1.numpy(), sg2=sg2.numpy(),
                        K_self=K_self.numpy(), K_cross=K_cross.numpy(), K_def=K_def.numpy(),
                        R_self=R_self.numpy(), R_cross=R_cross.numpy(), Bf=Bf.numpy(), y=yv.numpy(), mu=mu.numpy(),
                        s2=float(s2), kron_mv=mv.numpy(), kron_logdet=float(ld), kron_inv_diag=torch.diagonal(inv).numpy(),
                        kron_inv_row7=inv[7].numpy(), kron_prod=kp.numpy(), kron_diag=kd.numpy(),
                        logpdf0=float(lp0), logpdf2=float(lp2), logpdf1=float(lp1), rand_B=rB.numpy(), rand_K=rK.numpy())
    print("sim_code logpdf0/2/1", float(lp0), float(lp2), float(lp1))

    # a second, larger Kronecker case (T=200, D=2) matching SURVEY 8c's identity probe
    torch.manual_seed(6)
    T1, Dm = 200, 2
    x1 = torch.sort(torch.rand(T1).double())[0].view(-1, 1)
    ell1 = torch.exp(3 * (x1.view(-1) - 1) ** 3 - 3.0)
    K = kernels.Nonstationary_RBF_cov(x1, ell1=ell1)
    Lb = torch.tril(torch.randn(Dm, Dm).double()); Bf = Lb @ Lb.t()
    yv = torch.randn(Dm * T1).double()
    lp0 = distributions.multivariate_normal_logpdf0(yv, torch.zeros_like(yv), Bf, K, s2)
    lp2 = distributions.multivariate_normal_logpdf2(yv, torch.zeros_like(yv), Bf, K, s2)
    np.savez_compressed(os.path.join(OUT, "sim_code_t200.npz"), x1=x1.numpy().reshape(-1), ell1=ell1.numpy(),
                        Bf=Bf.numpy(), y=yv.numpy(), s2=float(s2), logpdf0=float(lp0), logpdf2=float(lp2),
                        K_diag=torch.diagonal(K).numpy(), K_row17=K[17].numpy())
    print("t200 logpdf0/2", float(lp0), float(lp2))


if __name__ == "__main__":
    main()

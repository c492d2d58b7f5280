"""Drop-in for the MAP predictive functions of code/SIM_code/Utility/prediction.py (point_predmap 337-408,
pointwise_predmap 410-430, test_predmap 432-458) and the vector <-> lower-triangle helpers of
code/SIM_code/Utility/utils.py:10-88 they use.

The reference redoes, for EVERY test point, two T x T LU solves and torch.symeig(K_x).  Here everything that does not
depend on the test point is factorised once (`KroneckerPosterior`): the two GP-conditional systems, the eigen-blocks
sigma2 I + lambda_k K_x of sigma2 I + B_f (x) K_x (one blocked Cholesky per block, no symeig of K_x), and
alpha_k = A_k^-1 (V^T Y^T)_k.  Because k_f = B_f (x) k_x for a single test input, in the eigenbasis of B_f

    mean_m     = sum_k lambda_k V[m,k] * (k_x . alpha_k)
    k_f,m^T S^-1 k_f,m = sum_k (lambda_k V[m,k])^2 * (k_x . A_k^-1 k_x)

so one test point costs D vector solves and 2D dot products.  Results equal the reference's to rounding (tests).
"""
import numpy as np
import torch

from . import _ops as ops
from . import kernels, kronecker_operation, settings


def vec2pars(pars_hist, N, M):
    """prediction.py:14-24: split a history of flat parameter vectors (slicing only)."""
    P = M * (M + 1) // 2
    return pars_hist[:, :N], pars_hist[:, N:2 * N], pars_hist[:, 2 * N:2 * N + P], pars_hist[:, -1]


def vec2list(y_vec, indx):
    """prediction.py:26-30: observations of each output as a list (indexing only)."""
    M = int(np.unique(np.asarray(indx.cpu() if torch.is_tensor(indx) else indx)).shape[0])
    return [y_vec[indx == i] for i in range(M)]


# ---- code/SIM_code/Utility/utils.py:10-88 (index plumbing on tiny vectors) -------------------------------------------
def _diag_positions(M):
    return list(np.cumsum(np.arange(1, M + 1)) - 1)


def uLvec2Lvec(uL_vec, M):
    """utils.py:10-22: exponentiate the diagonal entries of the packed lower triangle."""
    on = _diag_positions(M)
    L_vec = uL_vec.clone()
    L_vec[on] = torch.exp(uL_vec[on])
    return L_vec


def Lvec2uLvec(L_vec, M):
    """utils.py:24-36."""
    on = _diag_positions(M)
    uL_vec = L_vec.clone()
    uL_vec[on] = torch.log(L_vec[on])
    return uL_vec


def uLvecs2Lvecs(uL_vecs, N, M):
    """utils.py:38-46: uLvec2Lvec applied to each of the N consecutive packed triangles."""
    P = M * (M + 1) // 2
    on = _diag_positions(M)
    U = uL_vecs.reshape(N, P).clone()
    U[:, on] = torch.exp(U[:, on])
    return U.reshape(-1)


def vec2lowtriangle(x, N=None):
    """utils.py:56-74."""
    if N * (N + 1) / 2 != x.shape[0]:
        raise ValueError("check the dimension size!")
    mat = torch.zeros(N, N, dtype=torch.float64, device=x.device)      # settings.torchType = DoubleTensor
    idx = torch.tril_indices(N, N, device=x.device)
    mat[idx[0], idx[1]] = x
    return mat


def lowtriangle2vec(L, N=None):
    """utils.py:77-88."""
    idx = torch.tril_indices(N, N, device=L.device)
    return L[idx[0], idx[1]]


# ---------------------------------------------------------------------------------------------------------------------
class _ConditionalGP:
    """GP conditional of a log-hyper-function (log-ell or log-sigma) at a new input (prediction.py:52-57, 353-358):
    Sigma = RBF_cov(x) (+1e-6 I) is factorised once; it depends only on (x, alpha, beta)."""

    def __init__(self, x, mu, alpha, beta):
        self.x, self.mu, self.alpha, self.beta = x, float(mu), float(alpha), float(beta)
        self.Sigma = kernels.RBF_cov(x, alpha=self.alpha, beta=self.beta)
        self.Lc, _ = ops.potrf_big(self.Sigma.clone())

    def solve(self, b):
        """Sigma^-1 b with one step of iterative refinement in FP64: Sigma = RBF + 1e-6 I has a condition number of
        ~1e8, so a plain Cholesky solve carries ~cond * eps = 1e-8; the refinement step brings the solution (and the
        conditional variance that is a cancellation of it) to rounding level, which is what lets these predictors be
        compared with the reference at 1e-9."""
        b = b.contiguous()
        x = ops.potrs_vec(self.Lc, b)
        r = ops.axpby(b, ops.gemm_nt(x.view(1, -1), self.Sigma).view(-1), 1.0, -1.0)     # b - Sigma x (Sigma symmetric)
        return ops.axpby(x, ops.potrs_vec(self.Lc, r), 1.0, 1.0)

    def weights(self, tilde):
        """Sigma^-1 (tilde - mu): enough for the conditional mean."""
        return self.solve(tilde - self.mu)

    def projection(self, xs):
        """(k, Sigma^-1 k, k** - k . Sigma^-1 k) for the single input xs [1,1]."""
        k = kernels.RBF_cov(self.x, xs, alpha=self.alpha, beta=self.beta).view(-1).contiguous()
        proj = self.solve(k)
        kss = kernels.RBF_cov(xs, alpha=self.alpha, beta=self.beta).view(())
        return k, proj, kss - ops.dot(proj, k).reshape(())


class _KronState:
    """sigma2 I + B_f (x) K_x factorised through the eigen-blocks of B_f for ONE parameter sample
    (prediction.py:60-77 / 360-381), plus alpha_k = A_k^-1 (V^T Y^T)_k."""

    def __init__(self, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, stationary=False):
        N, M = Y.shape
        self.x = x
        self.sigma2_err = torch.exp(tilde_sigma2_err)
        self.l = torch.exp(tilde_l).contiguous()
        self.sigma = torch.exp(tilde_sigma).contiguous()
        L = vec2lowtriangle(uLvec2Lvec(uL_vec, M), M)
        self.B_f = ops.gemm_nt(L.contiguous(), L.contiguous())
        if stationary:                                     # the *_S variants: scalar sigma, ell (prediction.py:1547)
            K_x = kernels.RBF_cov(x, alpha=float(self.sigma), beta=float(self.l))
        else:
            K_x = kernels.Nonstationary_RBF_cov(x, sigma1=self.sigma, ell1=self.l)
        self._factor(K_x, Y)

    def _factor(self, K_x, Y):
        N, M = Y.shape
        y = Y.t().contiguous().view(-1)
        self.blocks = []          # (lambda_k, chol(sigma2 I + lambda_k K_x), alpha_k)
        Rt = None
        for k, lam_k, Lk, hld, V in kronecker_operation._factor_blocks(self.sigma2_err, self.B_f, K_x):
            if Rt is None:
                self.V = V
                Rt = ops.gemm_nt(V.t().contiguous(), y.view(M, N).t().contiguous())        # rows: (V^T (x) I) y
            self.blocks.append((lam_k, Lk, ops.potrs_vec(Lk, Rt[k].contiguous())))
        self.lam = torch.tensor([b[0] for b in self.blocks], dtype=torch.float64, device=Y.device)

    def predict_with(self, k_x, k_ss):
        """Mean and variance of y for a cross-covariance vector k_x [N] and prior variance factor k_ss (scalar)."""
        dots = torch.stack([torch.stack((ops.dot(k_x, a_k).reshape(()),
                                         ops.dot(k_x, ops.potrs_vec(Lk, k_x)).reshape(())))
                            for _, Lk, a_k in self.blocks])                                        # [D, 2]
        c = self.V * self.lam.view(1, -1)                       # c[m, k] = lambda_k V[m, k]
        mu_f = c @ dots[:, 0]
        sigma2_f = torch.diagonal(self.B_f) * k_ss - (c * c) @ dots[:, 1]
        return mu_f, sigma2_f + self.sigma2_err

    def predict(self, xs, l_star, sigma_star, self_jitter=True):
        """Predictive mean and variance of y at xs given (ell, sigma) there (prediction.py:78-93 / 382-402).
        self_jitter=False: prior variance sigma_star^2 B_f[m,m] without the 1e-6 of the self-covariance, as
        point_predmap_sampling writes it (prediction.py:262)."""
        k_x = kernels.Nonstationary_RBF_cov(X1=self.x, sigma1=self.sigma, ell1=self.l, X2=xs, sigma2=sigma_star,
                                            ell2=l_star).view(-1).contiguous()
        if self_jitter:
            k_ss = kernels.Nonstationary_RBF_cov(X1=xs, sigma1=sigma_star, ell1=l_star).view(())  # incl. the 1e-6 jitter
        else:
            k_ss = (sigma_star * sigma_star).view(())
        mu_f, sigma2_y = self.predict_with(k_x, k_ss)
        sigma2_y = torch.where(sigma2_y <= 0, torch.full_like(sigma2_y, settings.precision), sigma2_y)
        return mu_f, sigma2_y

    def predict_stationary(self, xs):
        """The *_S variants (prediction.py:1550-1556): k_x = RBF_cov(x, xs; sigma, ell), prior variance sigma^2 B_f[m,m],
        variances clipped where < 0 (strictly, as the reference writes it)."""
        k_x = kernels.RBF_cov(self.x, xs, alpha=float(self.sigma), beta=float(self.l)).view(-1).contiguous()
        mu_f, sigma2_y = self.predict_with(k_x, (self.sigma * self.sigma).view(()))
        sigma2_y = torch.where(sigma2_y < 0, torch.full_like(sigma2_y, settings.precision), sigma2_y)
        return mu_f, sigma2_y


class KroneckerPosterior:
    """Everything of point_predmap that does not depend on x_star (prediction.py:350-381)."""

    def __init__(self, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                 mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma):
        self.x = x.contiguous().view(-1, 1)
        self.gp_l = _ConditionalGP(self.x, mu_tilde_l, alpha_tilde_l, beta_tilde_l)
        self.gp_s = _ConditionalGP(self.x, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
        self.w_l, self.w_s = self.gp_l.weights(tilde_l), self.gp_s.weights(tilde_sigma)
        self.state = _KronState(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, self.x)

    def point(self, x_star):
        """prediction.py:353-408 for one test input: [3, M] = (mean - 1.96 sd, mean, mean + 1.96 sd)."""
        xs = x_star.reshape(1, 1).to(self.x.dtype)
        est = []
        for gp, wts in ((self.gp_l, self.w_l), (self.gp_s, self.w_s)):
            k = kernels.RBF_cov(self.x, xs, alpha=gp.alpha, beta=gp.beta).view(-1).contiguous()
            est.append(gp.mu + ops.dot(k, wts).reshape(()))
        mu_f, sigma2_y = self.state.predict(xs, torch.exp(est[0]).view(1), torch.exp(est[1]).view(1))
        sd = torch.sqrt(sigma2_y)
        return torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd])


def point_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_star, mu_tilde_l, alpha_tilde_l,
                  beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:337-408."""
    post = KroneckerPosterior(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l,
                              beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
    return post.point(x_star)


def pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l,
                      beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:410-430: [N_grid, 3, M]; the factorisations are shared by all grid points."""
    post = KroneckerPosterior(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l,
                              beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
    return torch.stack([post.point(g) for g in grids])


def test_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                 beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:432-458: same loop over held-out inputs."""
    return pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                             beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)





# ---- sampling predictors (prediction.py:34-184) -----------------------------------------------------------------------
def _draw(loc, scale):
    """torch.distributions.Normal(loc, scale).sample() of the reference: the same torch.normal call on the CPU generator,
    so a seeded run consumes the random stream exactly as the reference does."""
    return torch.normal(loc.detach().cpu(), scale.detach().cpu()).to(loc.device)


def _predsample(hist, Y, x, points, hyp_l, hyp_s, N_sample):
    tl_h, ts_h, uL_h, s2_h = (h[-N_sample:] for h in hist)
    xcol = x.contiguous().view(-1, 1)
    gp_l, gp_s = _ConditionalGP(xcol, *hyp_l), _ConditionalGP(xcol, *hyp_s)
    # one factorisation per parameter sample, shared by all test inputs (the reference redoes it per input)
    states = [_KronState(tl, ts, uL, s2, Y, xcol) for tl, ts, uL, s2 in zip(tl_h, ts_h, uL_h, s2_h)]
    floor = torch.tensor(settings.precision, dtype=torch.float64, device=Y.device)
    out = []
    for x_star in points:                                   # draw order of the reference: inputs outer, samples inner
        xs = x_star.reshape(1, 1).to(xcol.dtype)
        _, proj_l, var_l = gp_l.projection(xs)
        _, proj_s, var_s = gp_s.projection(xs)
        var_l = torch.where(var_l < 0, floor, var_l)
        var_s = torch.where(var_s < 0, floor, var_s)
        rows = []
        for st, tl, ts in zip(states, tl_h, ts_h):
            mu_l = gp_l.mu + ops.dot(proj_l, (tl - gp_l.mu).contiguous()).reshape(())
            l_star = torch.exp(_draw(mu_l, torch.sqrt(var_l))).view(1)
            mu_s = gp_s.mu + ops.dot(proj_s, (ts - gp_s.mu).contiguous()).reshape(())
            sigma_star = torch.exp(_draw(mu_s, torch.sqrt(var_s))).view(1)
            mu_f, sigma2_y = st.predict(xs, l_star, sigma_star)
            rows.append(_draw(mu_f, torch.sqrt(sigma2_y)))
        out.append(torch.stack(rows))
    return out


def point_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, x_star, mu_tilde_l,
                     alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample,
                     *args, **kwargs):
    """prediction.py:34-131: one posterior draw of y(x_star) per retained parameter sample, [N_sample, M]."""
    return _predsample((tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist), Y, x, [x_star],
                       (mu_tilde_l, alpha_tilde_l, beta_tilde_l), (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma),
                       N_sample)[0]


def pointwise_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, grids, mu_tilde_l,
                         alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample,
                         *args, **kwargs):
    """prediction.py:133-156: numpy array [N_grid, N_sample, M] (the reference returns numpy here)."""
    res = _predsample((tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist), Y, x, list(grids),
                      (mu_tilde_l, alpha_tilde_l, beta_tilde_l), (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma),
                      N_sample)
    return torch.stack(res).cpu().numpy()


def test_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, x_test, mu_tilde_l,
                    alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample,
                    *args, **kwargs):
    """prediction.py:158-184: the same loop over held-out inputs."""
    return pointwise_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, x_test,
                                mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
                                beta_tilde_sigma, N_sample)


def _predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, points, hyp_l, hyp_s):
    xcol = x.contiguous().view(-1, 1)
    gp_l, gp_s = _ConditionalGP(xcol, *hyp_l), _ConditionalGP(xcol, *hyp_s)
    state = _KronState(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, xcol)
    floor = torch.tensor(settings.precision, dtype=torch.float64, device=Y.device)
    dl, ds = (tilde_l - gp_l.mu).contiguous(), (tilde_sigma - gp_s.mu).contiguous()
    out = []
    for x_star in points:
        xs = x_star.reshape(1, 1).to(xcol.dtype)
        _, proj_l, var_l = gp_l.projection(xs)
        _, proj_s, var_s = gp_s.projection(xs)
        var_l = torch.where(var_l < 0, floor, var_l)
        var_s = torch.where(var_s < 0, floor, var_s)
        mu_l = gp_l.mu + ops.dot(proj_l, dl).reshape(())
        mu_s = gp_s.mu + ops.dot(proj_s, ds).reshape(())
        draws = []
        for _ in range(n_sample):                           # draw order of the reference: ell*, sigma*, y per sample
            l_star = torch.exp(_draw(mu_l, torch.sqrt(var_l))).view(1)
            sigma_star = torch.exp(_draw(mu_s, torch.sqrt(var_s))).view(1)
            mu_f, sigma2_y = state.predict(xs, l_star, sigma_star, self_jitter=False)
            draws.append(_draw(mu_f, torch.sqrt(sigma2_y)))
        ys = torch.stack(draws).cpu().numpy()
        out.append((np.percentile(ys, q=[2.5, 97.5], axis=0), np.mean(ys, axis=0), np.std(ys, axis=0)))
    return out


def point_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_star, mu_tilde_l,
                           alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma,
                           *args, **kwargs):
    """prediction.py:189-277: (2.5 % / 97.5 % quantiles [2, M], mean [M], std [M]) of n_sample predictive draws."""
    return _predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, [x_star],
                             (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                             (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))[0]


def pointwise_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, mu_tilde_l,
                               alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma,
                               *args, **kwargs):
    """prediction.py:279-306: stacked over the grid: ([G, 2, M], [G, M], [G, M])."""
    res = _predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, list(grids),
                            (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                            (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return (np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.stack([r[2] for r in res]))


def test_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l,
                          alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma,
                          *args, **kwargs):
    """prediction.py:308-335."""
    return pointwise_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test,
                                      mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
                                      beta_tilde_sigma)


# ---- stationary variants (prediction.py:1532-1658): dense inverse of B_f (x) K_x + sigma2 I -> eigen-block factors ---------
def pointwise_predmap_S(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, *args, **kwargs):
    """prediction.py:1532-1565: [N_grid, 3, M]."""
    st = _KronState(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x.contiguous().view(-1, 1), stationary=True)
    res = []
    for grid in grids:
        mu_f, s2 = st.predict_stationary(grid.reshape(1, 1).to(torch.float64))
        sd = torch.sqrt(s2)
        res.append(torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd]))
    return torch.stack(res)


def test_predmap_S(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, test_x, *args, **kwargs):
    """prediction.py:1567-1604: (mean [N_test, M], std [N_test, M])."""
    st = _KronState(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x.contiguous().view(-1, 1), stationary=True)
    out = [st.predict_stationary(xs.reshape(1, 1).to(torch.float64)) for xs in test_x]
    return torch.stack([o[0] for o in out]), torch.stack([torch.sqrt(o[1]) for o in out])


def pointwise_predsample_S(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, grids, *args, **kwargs):
    """prediction.py:1606-1631: numpy [N_sample, N_grid, M]; ONE np.random.randn() scalar per (sample, grid point), shared
    by the M outputs, drawn from numpy's global generator in the reference's order."""
    xcol = x.contiguous().view(-1, 1)
    samples = []
    for tl, ts, uL, s2 in zip(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs):
        st = _KronState(tl, ts, uL, s2, Y, xcol, stationary=True)
        res = []
        for grid in grids:
            mu_f, var = st.predict_stationary(grid.reshape(1, 1).to(torch.float64))
            res.append(mu_f + np.random.randn() * torch.sqrt(var))
        samples.append(torch.stack(res))
    return torch.stack(samples).cpu().numpy()


def test_predsample_S(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, test_x, *args, **kwargs):
    """prediction.py:1633-1658."""
    return pointwise_predsample_S(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, test_x)


# ---- Hadamard (irregular observations) MAP predictors (prediction.py:710-910) ------------------------------------------
class _HadamardState:
    """S = K_x * K_i + sigma2 I for observations (x_n, indx_n, y_n) (prediction.py:742-750): one dense blocked Cholesky
    instead of symeig(K) and an explicit inverse; alpha = S^-1 y."""

    def __init__(self, tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, M, stationary=False):
        self.stationary = stationary
        self.x = x.contiguous().view(-1, 1)
        self.indx = indx.to(torch.int32).contiguous()
        self.sigma2_err = torch.exp(tilde_sigma2_err)
        self.l = torch.exp(tilde_l).contiguous()
        self.sigma = torch.exp(tilde_sigma).contiguous()
        Lm = vec2lowtriangle(L_vec, M)
        self.B_f = ops.gemm_nt(Lm.contiguous(), Lm.contiguous())
        if stationary:                                     # scalar sigma, ell (prediction.py:1669)
            K_x = kernels.RBF_cov(self.x, alpha=float(self.sigma), beta=float(self.l))
        else:
            K_x = kernels.Nonstationary_RBF_cov(self.x, sigma1=self.sigma, ell1=self.l)
        S = ops.hadamard_index_cov(K_x, self.B_f, self.indx, self.indx, float(self.sigma2_err))
        self.Lc, _ = ops.potrf_big(S)
        self.alpha = ops.potrs_vec(self.Lc, y.contiguous())
        self.M = M

    def cross(self, xs, l_star, sigma_star):
        return kernels.Nonstationary_RBF_cov(X1=self.x, sigma1=self.sigma, ell1=self.l, X2=xs, sigma2=sigma_star,
                                             ell2=l_star).contiguous()                            # [N, 1]

    def output(self, k_x, m, k_ss, prior_index=None):
        """mean and variance of output m at the test input with cross-covariance column k_x (prior variance taken from
        output `prior_index`, default m)."""
        mi = torch.full((1,), int(m), dtype=torch.int32, device=k_x.device)
        kf = ops.hadamard_index_cov(k_x, self.B_f, self.indx, mi, 0.0).view(-1).contiguous()      # B_f[indx_n, m] k_x[n]
        mu = ops.dot(kf, self.alpha).reshape(())
        pi = int(m) if prior_index is None else int(prior_index)
        var = self.B_f[pi, pi] * k_ss - ops.dot(kf, ops.potrs_vec(self.Lc, kf)).reshape(()) + self.sigma2_err
        return mu, var

    def stationary_point(self, x_star, outputs, prior_index=None):
        xs = x_star.reshape(1, 1).to(torch.float64)
        a, b = float(self.sigma), float(self.l)
        k_x = kernels.RBF_cov(self.x, xs, alpha=a, beta=b).contiguous()
        k_ss = kernels.RBF_cov(xs, alpha=a, beta=b).view(())                                       # sigma^2 + 1e-6
        mus, vs = zip(*[self.output(k_x, m, k_ss, prior_index) for m in outputs])
        mu_f, s2 = torch.stack(mus), torch.stack(vs)
        return mu_f, torch.where(s2 <= 0, torch.full_like(s2, settings.precision), s2)


def _hadamard_setup(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, hyp_l, hyp_s):
    M = int(torch.unique(indx).numel())
    xcol = x.contiguous().view(-1, 1)
    gp_l, gp_s = _ConditionalGP(xcol, *hyp_l), _ConditionalGP(xcol, *hyp_s)
    return (M, gp_l, gp_s, gp_l.weights(tilde_l), gp_s.weights(tilde_sigma),
            _HadamardState(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, M))


def _hadamard_point(setup, x_star, outputs):
    M, gp_l, gp_s, w_l, w_s, st = setup
    xs = x_star.reshape(1, 1).to(torch.float64)
    est = []
    for gp, wts in ((gp_l, w_l), (gp_s, w_s)):
        k = kernels.RBF_cov(st.x, xs, alpha=gp.alpha, beta=gp.beta).view(-1).contiguous()
        est.append(gp.mu + ops.dot(k, wts).reshape(()))
    l_star, sigma_star = torch.exp(est[0]).view(1), torch.exp(est[1]).view(1)
    k_x = st.cross(xs, l_star, sigma_star)
    k_ss = kernels.Nonstationary_RBF_cov(X1=xs, sigma1=sigma_star, ell1=l_star).view(())          # incl. the 1e-6 jitter
    mus, vs = zip(*[st.output(k_x, m, k_ss) for m in outputs])
    mu_f, s2 = torch.stack(mus), torch.stack(vs)
    s2 = torch.where(s2 <= 0, torch.full_like(s2, settings.precision), s2)
    sd = torch.sqrt(s2)
    return torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd])


def point_predmap_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, x_star, mu_tilde_l, alpha_tilde_l,
                           beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:710-785: [3, M] band of all outputs at x_star."""
    setup = _hadamard_setup(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y,
                            (mu_tilde_l, alpha_tilde_l, beta_tilde_l), (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return _hadamard_point(setup, x_star, range(setup[0]))


def pointwise_predmap_hadmard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, grids, mu_tilde_l,
                              alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma,
                              *args, **kwargs):
    """prediction.py:787-808 (the reference's spelling): [N_grid, 3, M], one factorisation for all grid points."""
    setup = _hadamard_setup(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y,
                            (mu_tilde_l, alpha_tilde_l, beta_tilde_l), (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return torch.stack([_hadamard_point(setup, g, range(setup[0])) for g in grids])


def indexedpoint_predmap_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, x_star, indx_star,
                                  mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
                                  beta_tilde_sigma, *args, **kwargs):
    """prediction.py:810-885: [3] band of output indx_star at x_star."""
    setup = _hadamard_setup(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y,
                            (mu_tilde_l, alpha_tilde_l, beta_tilde_l), (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return _hadamard_point(setup, x_star, [int(indx_star)]).view(3)


def test_predmap_harmard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, x_test, indx_test, mu_tilde_l,
                         alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma,
                         *args, **kwargs):
    """prediction.py:887-909 (the reference's spelling): [N_test, 3]."""
    setup = _hadamard_setup(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y,
                            (mu_tilde_l, alpha_tilde_l, beta_tilde_l), (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return torch.stack([_hadamard_point(setup, xs, [int(ii)]).view(3) for xs, ii in zip(x_test, indx_test)])


# ---- Hadamard sampling predictors (prediction.py:461-708) ---------------------------------------------------------------
def _hadamard_predsample(hist, x, indx, y, points, outputs_of, hyp_l, hyp_s):
    tl_h, ts_h, L_h, s2_h = hist
    M = int(torch.unique(indx).numel())
    xcol = x.contiguous().view(-1, 1)
    gp_l, gp_s = _ConditionalGP(xcol, *hyp_l), _ConditionalGP(xcol, *hyp_s)
    states = [_HadamardState(tl, ts, Lv, s2, x, indx, y, M) for tl, ts, Lv, s2 in zip(tl_h, ts_h, L_h, s2_h)]
    floor = torch.tensor(settings.precision, dtype=torch.float64, device=y.device)
    out = []
    for ip, x_star in enumerate(points):                    # draw order of the reference: inputs outer, samples inner
        xs = x_star.reshape(1, 1).to(torch.float64)
        _, proj_l, var_l = gp_l.projection(xs)
        _, proj_s, var_s = gp_s.projection(xs)
        var_l = torch.where(var_l < 0, floor, var_l)
        var_s = torch.where(var_s < 0, floor, var_s)
        outputs = outputs_of(ip, M)
        rows = []
        for st, tl, ts in zip(states, tl_h, ts_h):
            mu_l = gp_l.mu + ops.dot(proj_l, (tl - gp_l.mu).contiguous()).reshape(())
            l_star = torch.exp(_draw(mu_l, torch.sqrt(var_l))).view(1)
            mu_s = gp_s.mu + ops.dot(proj_s, (ts - gp_s.mu).contiguous()).reshape(())
            sigma_star = torch.exp(_draw(mu_s, torch.sqrt(var_s))).view(1)
            k_x = st.cross(xs, l_star, sigma_star)
            k_ss = kernels.Nonstationary_RBF_cov(X1=xs, sigma1=sigma_star, ell1=l_star).view(())
            mus, vs = zip(*[st.output(k_x, m, k_ss) for m in outputs])
            mu_f, s2 = torch.stack(mus), torch.stack(vs)
            s2 = torch.where(s2 <= 0, torch.full_like(s2, settings.precision), s2)
            rows.append(_draw(mu_f, torch.sqrt(s2)))
        out.append(rows)
    return out


def point_predsample_hadamard(tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist, x, indx, y, x_star,
                              mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
                              beta_tilde_sigma, *args, **kwargs):
    """prediction.py:461-553: [N_hist, M]."""
    res = _hadamard_predsample((tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist), x, indx, y, [x_star],
                               lambda ip, M: range(M), (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                               (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return torch.stack(res[0])


def pointwise_predsample_hadamard(tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist, x, indx, y, grids,
                                  mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
                                  beta_tilde_sigma, *args, **kwargs):
    """prediction.py:555-583: [N_grid, N_hist, M]."""
    res = _hadamard_predsample((tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist), x, indx, y,
                               list(grids), lambda ip, M: range(M), (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                               (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return torch.stack([torch.stack(r) for r in res])


def indexedpoint_predsample_hadamard(tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist, x, indx, y,
                                     x_star, indx_star, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma,
                                     alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """prediction.py:585-676: [N_hist] draws of output indx_star at x_star."""
    res = _hadamard_predsample((tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist), x, indx, y, [x_star],
                               lambda ip, M: [int(indx_star)], (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                               (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return torch.cat(res[0])


def test_predsample_hadamard(tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist, x, indx, y, x_test,
                             indx_test, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
                             beta_tilde_sigma, *args, **kwargs):
    """prediction.py:678-708: [N_test, N_hist]."""
    idx = [int(i) for i in indx_test]
    res = _hadamard_predsample((tilde_l_hist, tilde_sigma_hist, L_vec_hist, tilde_sigma2_err_hist), x, indx, y,
                               list(x_test), lambda ip, M: [idx[ip]], (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                               (mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma))
    return torch.stack([torch.cat(r) for r in res])


# ---- stationary Hadamard MAP predictors (prediction.py:1661-1762) -------------------------------------------------------
def _S_hadamard_state(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y):
    return _HadamardState(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, int(torch.unique(indx).numel()),
                          stationary=True)


def point_predmap_S_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, x_star, *args, **kwargs):
    """prediction.py:1661-1694: [3, M]."""
    st = _S_hadamard_state(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y)
    mu_f, s2 = st.stationary_point(x_star, range(st.M))
    sd = torch.sqrt(s2)
    return torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd])


def pointwise_predmap_S_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, grids, *args, **kwargs):
    """prediction.py:1696-1706: [N_grid, 3, M]."""
    st = _S_hadamard_state(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y)
    res = []
    for grid in grids:
        mu_f, s2 = st.stationary_point(grid, range(st.M))
        sd = torch.sqrt(s2)
        res.append(torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd]))
    return torch.stack(res)


def _S_indexed(st, x_star, indx_star):
    # the reference takes (A - B)[0, 0] with A = B_f (x) k**: the prior variance of output 0 whatever indx_star is
    # (prediction.py:1738-1740); reproduced
    mu_f, s2 = st.stationary_point(x_star, [int(indx_star)], prior_index=0)
    return torch.stack([mu_f.view(()), torch.sqrt(s2).view(())])


def indexedpoint_predmap_S_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, x_star, indx_star,
                                    *args, **kwargs):
    """prediction.py:1708-1744: tensor [mean, std] of output indx_star at x_star."""
    return _S_indexed(_S_hadamard_state(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y), x_star, indx_star)


def test_predmap_S_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, x_test, indx_test,
                            *args, **kwargs):
    """prediction.py:1746-1762: (means [N_test], stds [N_test])."""
    st = _S_hadamard_state(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y)
    res = torch.stack([_S_indexed(st, xs, ii) for xs, ii in zip(x_test, indx_test)])
    return res[:, 0], res[:, 1]


# ---- spatially varying coregionalisation ("inhomogeneous") MAP predictors (prediction.py:912-1036) -------------------
class _InhomogeneousState:
    """K[(m,n),(m',n')] = K_x[n,n'] (L_n L_n'^T)[m,m'] + sigma2 I in the output-major order of y = Y^T.view(-1)
    (prediction.py:944-953): built as (L_row L_row^T) * K_x[n,n'] and factorised once by the blocked Cholesky; the GP
    conditionals of log-ell and of the P = M(M+1)/2 entries of the packed triangle are hoisted as well."""

    def __init__(self, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, hyp_l, hyp_L, constrained_conditional=False):
        N, M = Y.shape
        P = M * (M + 1) // 2
        dev = Y.device
        self.N, self.M, self.P = N, M, P
        self.x = x.contiguous().view(-1, 1)
        self.gp_l, self.gp_L = _ConditionalGP(self.x, *hyp_l), _ConditionalGP(self.x, *hyp_L)
        self.w_l = self.gp_l.weights(tilde_l)
        # the MAP predictors condition the GP on the unconstrained entries (prediction.py:931-934); the history-based
        # sampler conditions it on the constrained ones, diagonals already exponentiated (prediction.py:1259-1266)
        U = (uLvecs2Lvecs(uL_vecs, N, M) if constrained_conditional else uL_vecs).reshape(N, P)
        self.W_L = torch.stack([self.gp_L.weights(U[:, p_].contiguous()) for p_ in range(P)])       # [P, N]
        self.sigma2_err = torch.exp(tilde_sigma2_err)
        self.l = torch.exp(tilde_l).contiguous()
        Lmat = torch.zeros(N, M, M, dtype=torch.float64, device=dev)
        idx = torch.tril_indices(M, M, device=dev)
        Lmat[:, idx[0], idx[1]] = uLvecs2Lvecs(uL_vecs, N, M).reshape(N, P)
        self.Lrow = Lmat.permute(1, 0, 2).reshape(M * N, M).contiguous()          # row (m, n) = L_n[m, :]
        self.nidx = torch.arange(N, dtype=torch.int32, device=dev).repeat(M).contiguous()
        self.zero_idx = torch.zeros(M, dtype=torch.int32, device=dev)
        K_i = ops.gemm_nt(self.Lrow, self.Lrow)
        K_x = kernels.Nonstationary_RBF_cov(self.x, ell1=self.l)
        S = ops.hadamard_index_cov(K_i, K_x, self.nidx, self.nidx, float(self.sigma2_err))
        self.Lc, _ = ops.potrf_big(S)
        self.alpha = ops.potrs_vec(self.Lc, Y.t().contiguous().view(-1))

    def cond_l(self, xs, with_var=False):
        """conditional mean (and variance) of log-ell at xs"""
        k = kernels.RBF_cov(self.x, xs, alpha=self.gp_l.alpha, beta=self.gp_l.beta).view(-1).contiguous()
        mu = self.gp_l.mu + ops.dot(k, self.w_l).reshape(())
        return (mu, self.gp_l.projection(xs)[2]) if with_var else mu

    def cond_uL(self, xs, with_var=False):
        """conditional means of the P packed-triangle entries at xs (and their common variance)"""
        kL = kernels.RBF_cov(self.x, xs, alpha=self.gp_L.alpha, beta=self.gp_L.beta).view(1, -1).contiguous()
        mu = self.gp_L.mu + ops.gemm_nt(self.W_L.contiguous(), kL).view(-1)                        # [P]
        return (mu, self.gp_L.projection(xs)[2]) if with_var else mu

    def predict(self, xs, l_star, L_vec_star):
        """mean and variance of y(xs) given ell and the (constrained) packed triangle there"""
        N, M = self.N, self.M
        dev = self.x.device
        L_star = vec2lowtriangle(L_vec_star, M).contiguous()
        one = torch.ones(1, dtype=torch.float64, device=dev)
        k_x = kernels.Nonstationary_RBF_cov(X1=self.x, sigma1=torch.ones(N, dtype=torch.float64, device=dev), ell1=self.l,
                                            X2=xs, sigma2=one, ell2=l_star).contiguous()           # [N, 1]
        A_f = ops.hadamard_index_cov(self.Lrow, k_x, self.nidx, self.zero_idx, 0.0)                # row (m,n): k_x[n] L_n[m,:]
        k_fT = ops.gemm_nt(L_star, A_f)                                                            # [M, M N]
        k_ss = kernels.Nonstationary_RBF_cov(X1=xs, sigma1=one, ell1=l_star).view(())             # 1 + 1e-6
        prior = ops.gemm_nt(L_star, L_star)
        mus, vs = [], []
        for m in range(M):
            kf = k_fT[m].contiguous()
            mus.append(ops.dot(kf, self.alpha).reshape(()))
            vs.append(prior[m, m] * k_ss - ops.dot(kf, ops.potrs_vec(self.Lc, kf)).reshape(()) + self.sigma2_err)
        mu_f, s2 = torch.stack(mus), torch.stack(vs)
        return mu_f, torch.where(s2 <= 0, torch.full_like(s2, settings.precision), s2)

    def point(self, x_star):
        xs = x_star.reshape(1, 1).to(torch.float64)
        l_star = torch.exp(self.cond_l(xs)).view(1)
        L_vec_star = uLvec2Lvec(self.cond_uL(xs), self.M)
        mu_f, s2 = self.predict(xs, l_star, L_vec_star)
        sd = torch.sqrt(s2)
        return torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd]), L_vec_star

    def sample_point(self, x_star, n_sample, pred_smoothness=False, pred_cov=False):
        """prediction.py:1056-1158: n_sample draws at one input; the three modes of the reference and its draw order."""
        xs = x_star.reshape(1, 1).to(torch.float64)
        floor = torch.tensor(settings.precision, dtype=torch.float64, device=self.x.device)
        ls, Ls, ys = [], [], []
        if not pred_cov:
            mu_l, var_l = self.cond_l(xs, with_var=True)
            var_l = torch.where(var_l < 0, floor, var_l)
        if pred_cov or not pred_smoothness:
            mu_u, var_u = self.cond_uL(xs, with_var=True)
            sd_u = torch.sqrt(torch.where(var_u < 0, floor, var_u)).expand(self.P).contiguous()
        for _ in range(n_sample):
            if pred_smoothness:
                ls.append(_draw(mu_l, torch.sqrt(var_l)))
                continue
            if pred_cov:
                Ls.append(vec2lowtriangle(uLvec2Lvec(_draw(mu_u, sd_u), self.M), self.M))
                continue
            tl = _draw(mu_l, torch.sqrt(var_l))
            L_vec_star = uLvec2Lvec(_draw(mu_u, sd_u), self.M)
            mu_f, s2 = self.predict(xs, torch.exp(tl).view(1), L_vec_star)
            ys.append(_draw(mu_f, torch.sqrt(s2)))
        if pred_smoothness:
            return torch.stack(ls).cpu().numpy()
        if pred_cov:
            return torch.stack(Ls).cpu().numpy()
        ys = torch.stack(ys).cpu().numpy()
        return np.percentile(ys, q=[2.5, 97.5], axis=0), np.mean(ys, axis=0), np.std(ys, axis=0)


def point_predmap_inhomogeneous(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_star, mu_tilde_l, alpha_tilde_l,
                                beta_tilde_l, mu_L, alpha_L, beta_L, *args, **kwargs):
    """prediction.py:912-988: ([3, M] band, estimated packed triangle L_vec at x_star)."""
    st = _InhomogeneousState(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                             (mu_L, alpha_L, beta_L))
    return st.point(x_star)


def pointwise_predmap_inhomogeneous(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l,
                                    beta_tilde_l, mu_L, alpha_L, beta_L, *args, **kwargs):
    """prediction.py:990-1012: ([N_grid, 3, M], [N_grid, P]); one factorisation for all grid points."""
    st = _InhomogeneousState(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                             (mu_L, alpha_L, beta_L))
    res = [st.point(g) for g in grids]
    return torch.stack([r[0] for r in res]), torch.stack([r[1] for r in res])


def test_predmap_inhomogeneous(tilde_l, L_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                               mu_L, alpha_L, beta_L, *args, **kwargs):
    """prediction.py:1014-1036 (its `L_vecs` argument is passed on as the unconstrained uL_vecs, as in the reference)."""
    return pointwise_predmap_inhomogeneous(tilde_l, L_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                                           beta_tilde_l, mu_L, alpha_L, beta_L)


# ---- SVC Hadamard MAP predictors (prediction.py:1367-1530): irregular observations, a triangle per observation -------
class _SVCHadamardState:
    """K[n,n'] = K_x[n,n'] (L_n[indx_n,:] . L_n'[indx_n',:]) + sigma2 I (prediction.py:1391-1397), factorised once."""

    def __init__(self, tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, hyp_l, hyp_L):
        N = y.shape[0]
        M = int(torch.unique(indx).numel())
        P = M * (M + 1) // 2
        dev = y.device
        self.N, self.M, self.P = N, M, P
        self.x = x.contiguous().view(-1, 1)
        self.gp_l, self.gp_L = _ConditionalGP(self.x, *hyp_l), _ConditionalGP(self.x, *hyp_L)
        self.w_l = self.gp_l.weights(tilde_l)
        Lv = L_vecs.reshape(N, P)                            # used as given (no exp of the diagonal, prediction.py:1385)
        self.W_L = torch.stack([self.gp_L.weights(Lv[:, p_].contiguous()) for p_ in range(P)])    # [P, N]
        self.sigma2_err = torch.exp(tilde_sigma2_err)
        self.l = torch.exp(tilde_l).contiguous()
        Lmat = torch.zeros(N, M, M, dtype=torch.float64, device=dev)
        idx = torch.tril_indices(M, M, device=dev)
        Lmat[:, idx[0], idx[1]] = Lv
        self.Lsel = Lmat[torch.arange(N, device=dev), indx.long().to(dev)].contiguous()            # row n = L_n[indx_n, :]
        self.ident = torch.arange(N, dtype=torch.int32, device=dev)
        self.zero_idx = torch.zeros(M, dtype=torch.int32, device=dev)
        K_i = ops.gemm_nt(self.Lsel, self.Lsel)
        K_x = kernels.Nonstationary_RBF_cov(self.x, ell1=self.l)
        S = ops.hadamard_index_cov(K_i, K_x, self.ident, self.ident, float(self.sigma2_err))
        self.Lc, _ = ops.potrf_big(S)
        self.alpha = ops.potrs_vec(self.Lc, y.contiguous())

    def _common(self, x_star):
        N, M = self.N, self.M
        dev = self.x.device
        xs = x_star.reshape(1, 1).to(torch.float64)
        k = kernels.RBF_cov(self.x, xs, alpha=self.gp_l.alpha, beta=self.gp_l.beta).view(-1).contiguous()
        l_star = torch.exp(self.gp_l.mu + ops.dot(k, self.w_l).reshape(())).view(1)
        kL = kernels.RBF_cov(self.x, xs, alpha=self.gp_L.alpha, beta=self.gp_L.beta).view(1, -1).contiguous()
        L_star = vec2lowtriangle(self.gp_L.mu + ops.gemm_nt(self.W_L.contiguous(), kL).view(-1), M).contiguous()
        one = torch.ones(1, dtype=torch.float64, device=dev)
        k_x = kernels.Nonstationary_RBF_cov(X1=self.x, sigma1=torch.ones(N, dtype=torch.float64, device=dev), ell1=self.l,
                                            X2=xs, sigma2=one, ell2=l_star).contiguous()           # [N, 1]
        k_ss = kernels.Nonstationary_RBF_cov(X1=xs, ell1=l_star).view(())                          # 1 + 1e-6
        prior = torch.diagonal(ops.gemm_nt(L_star, L_star)) * k_ss                                 # diag of A
        k_i = ops.gemm_nt(self.Lsel, L_star)                                                       # [N, M]
        k_f = ops.hadamard_index_cov(k_i, k_x, self.ident, self.zero_idx, 0.0)                     # rows scaled by k_x[n]
        return prior, k_f.t().contiguous()                                                         # [M], [M, N]

    def _quad(self, kf):
        return ops.dot(kf, self.alpha).reshape(()), ops.dot(kf, ops.potrs_vec(self.Lc, kf)).reshape(())

    def point(self, x_star):
        prior, k_fT = self._common(x_star)
        mq = [self._quad(k_fT[m].contiguous()) for m in range(self.M)]
        mu_f = torch.stack([a for a, _ in mq])
        s2 = prior - torch.stack([b for _, b in mq]) + self.sigma2_err
        s2 = torch.where(s2 <= 0, torch.full_like(s2, settings.precision), s2)
        sd = torch.sqrt(s2)
        return torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd])

    def indexed(self, x_star, indx_star):
        """[mean, A[0,0]-b+s2, ..., A[M-1,M-1]-b+s2]: the reference subtracts the 1x1 explained variance of the requested
        output from the whole prior diagonal and returns VARIANCES (prediction.py:1505-1512); reproduced."""
        prior, k_fT = self._common(x_star)
        mu, b = self._quad(k_fT[int(indx_star)].contiguous())
        s2 = prior - b + self.sigma2_err
        s2 = torch.where(s2 <= 0, torch.full_like(s2, settings.precision), s2)
        return torch.cat([mu.view(1), s2.view(-1)])


def _svc_state(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, hyper):
    mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L = hyper[:6]
    return _SVCHadamardState(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, (mu_tilde_l, alpha_tilde_l, beta_tilde_l),
                             (mu_L, alpha_L, beta_L))


def point_predmap_SVC_hadamard(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, x_star, mu_tilde_l, alpha_tilde_l,
                               beta_tilde_l, mu_L, alpha_L, beta_L, *args, **kwargs):
    """prediction.py:1367-1431: [3, M]."""
    return _svc_state(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y,
                      (mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L)).point(x_star)


def pointwise_predmap_SVC_hadamard(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, grids, *args, **kwargs):
    """prediction.py:1433-1444: [N_grid, 3, M]; the six hyper-parameters travel in *args as in the reference."""
    st = _svc_state(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, args)
    return torch.stack([st.point(g) for g in grids])


def indexedpoint_predmap_SVC_hadamard(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, x_star, indx_star, mu_tilde_l,
                                      alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, *args, **kwargs):
    """prediction.py:1446-1513: tensor of length 1 + M (see _SVCHadamardState.indexed)."""
    return _svc_state(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y,
                      (mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L)).indexed(x_star, indx_star)


def test_predmap_SVC_hadamard(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, x_test, indx_test, *args, **kwargs):
    """prediction.py:1515-1530: (res[:, 0], res[:, 1]) of the stacked indexed results."""
    st = _svc_state(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, args)
    res = torch.stack([st.indexed(xs, ii) for xs, ii in zip(x_test, indx_test)])
    return res[:, 0], res[:, 1]


def _inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, points, hyper, pred_smoothness, pred_cov):
    st = _InhomogeneousState(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, hyper[:3], hyper[3:6])
    res = [st.sample_point(p_, n_sample, pred_smoothness, pred_cov) for p_ in points]
    if pred_smoothness or pred_cov:
        return res
    return [r[0] for r in res], [r[1] for r in res], [r[2] for r in res]


def point_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_star, mu_tilde_l,
                                         alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, pred_smoothness=False,
                                         pred_cov=False, *args, **kwargs):
    """prediction.py:1038-1158: (quantiles [2, M], mean [M], std [M]); or the n_sample draws of log-ell (pred_smoothness)
    / of the coregionalisation triangle [n_sample, M, M] (pred_cov) at x_star."""
    res = _inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, [x_star],
                                  (mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L), pred_smoothness, pred_cov)
    return res[0] if (pred_smoothness or pred_cov) else (res[0][0], res[1][0], res[2][0])


def pointwise_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l,
                                             alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, pred_smoothness=False,
                                             pred_cov=False, *args, **kwargs):
    """prediction.py:1160-1201: the same stacked over the grid (numpy)."""
    res = _inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, list(grids),
                                  (mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L), pred_smoothness, pred_cov)
    if pred_smoothness or pred_cov:
        return np.stack(res)
    return np.stack(res[0]), np.stack(res[1]), np.stack(res[2])


def test_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l,
                                        alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, *args, **kwargs):
    """prediction.py:1203-1229."""
    return pointwise_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l,
                                                    alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L)


def _predsample_inhomogeneous(hist, Y, x, points, hyper, N_sample):
    tl_h, uL_h, s2_h = (h[-N_sample:] for h in hist)
    states = [_InhomogeneousState(tl, uL, s2, Y, x, hyper[:3], hyper[3:6], constrained_conditional=True)
              for tl, uL, s2 in zip(tl_h, uL_h, s2_h)]
    floor = torch.tensor(settings.precision, dtype=torch.float64, device=Y.device)
    out = []
    for x_star in points:                                   # draw order of the reference: inputs outer, samples inner
        xs = x_star.reshape(1, 1).to(torch.float64)
        rows = []
        for st in states:
            mu_l, var_l = st.cond_l(xs, with_var=True)
            tl = _draw(mu_l, torch.sqrt(torch.where(var_l < 0, floor, var_l)))
            mu_L, var_L = st.cond_uL(xs, with_var=True)
            sd_L = torch.sqrt(torch.where(var_L < 0, floor, var_L)).expand(st.P).contiguous()
            L_vec_star = _draw(mu_L, sd_L)                  # used as the triangle itself (no exp), prediction.py:1267-1268
            mu_f, s2 = st.predict(xs, torch.exp(tl).view(1), L_vec_star)
            rows.append(_draw(mu_f, torch.sqrt(s2)))
        out.append(torch.stack(rows))
    return out


def point_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, x_star, mu_tilde_l,
                                   alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, *args, **kwargs):
    """prediction.py:1231-1323: [N_sample, M]."""
    return _predsample_inhomogeneous((tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist), Y, x, [x_star],
                                     (mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L), N_sample)[0]


def pointwise_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, grids, mu_tilde_l,
                                       alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, *args, **kwargs):
    """prediction.py:1325-1344: numpy [N_grid, N_sample, M]."""
    res = _predsample_inhomogeneous((tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist), Y, x, list(grids),
                                    (mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L), N_sample)
    return torch.stack(res).cpu().numpy()


def test_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, x_test, mu_tilde_l,
                                  alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, *args, **kwargs):
    """prediction.py:1346-1365."""
    return pointwise_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, x_test, mu_tilde_l,
                                              alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, N_sample)


test_predmap.__test__ = False      # not pytest tests
test_predsample_inhomogeneous.__test__ = False
test_predmap_inhomogeneous_sampling.__test__ = False
test_predmap_SVC_hadamard.__test__ = False
test_predmap_inhomogeneous.__test__ = False
test_predmap_S_hadamard.__test__ = False
test_predsample_hadamard.__test__ = False
test_predmap_harmard.__test__ = False
test_predmap_S.__test__ = False
test_predsample_S.__test__ = False
test_predmap_sampling.__test__ = False
test_predsample.__test__ = False

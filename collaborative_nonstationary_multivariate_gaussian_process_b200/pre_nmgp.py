"""Optional initialiser of code/pre_nmgp.py: local maximum-likelihood estimates of the log length-scale and
the noise variance at every inducing point, from its nearest observations, with the coregionalisation factor fixed to
the Cholesky factor of the empirical output covariance.  The reference imports it (code/nmgp_dsvi.py:16) and only uses it
from commented-out driver code; it is host-side NumPy/SciPy work on P = 10 points per inducing location and is NOT part
of the GPU hot path (nothing here touches the CUDA extension).

The local model is y ~ N(0, K (x) B + s2 I) with K = RBF(x_local; ell), B = L L^T.  The reference evaluates its density
through a dense (P D) x (P D) ``np.kron`` and ``scipy.stats.multivariate_normal`` (pre_nmgp.py:48-56); here the same
value comes from the two small eigen-decompositions of K and B (the Kronecker identity of SURVEY 3.4), which is what
``kronecker_operation`` does on the GPU for the big problem."""
import numpy as np

tridiagonal_jitter = 1e-6


def search_nearest_neighhood(x, Y, z_m, P=10):
    """pre_nmgp.py:9-12 (the reference always keeps 10 neighbours, whatever P says -- reproduced)."""
    indices = np.argsort(np.abs(x - z_m))[:10]
    return x[indices], Y[indices]


def create_RBF(X, X2=None, scale2=1., length_scales=1.):
    """pre_nmgp.py:28-33."""
    length_scales = max(float(length_scales), 1e-8)
    A = np.asarray(X, dtype=np.float64) / length_scales
    Bm = A if X2 is None else np.asarray(X2, dtype=np.float64) / length_scales
    d = A[:, None, :] - Bm[None, :, :]
    return scale2 * np.exp(-0.5 * np.sum(d * d, -1))


def _kron_gauss_logpdf(Y, K, Bm, s2):
    """log N(vec(Y) | 0, K (x) B + s2 I) for Y [N, D] (row-major vec: index n D + d), via eigh(K), eigh(B)."""
    wk, Vk = np.linalg.eigh(K)
    wb, Vb = np.linalg.eigh(Bm)
    lam = np.outer(wk, wb) + s2                               # eigenvalues of the Kronecker sum, [N, D]
    if np.any(lam <= 0):
        return -np.inf
    R = Vk.T @ Y @ Vb                                         # rotated data
    n = Y.size
    return float(-0.5 * np.sum(R * R / lam) - 0.5 * np.sum(np.log(lam)) - 0.5 * n * np.log(2.0 * np.pi))


def compute_loglik_part(pars, x, Y, L):
    """pre_nmgp.py:48-56: pars = (log sigma2_err, log ell), L fixed."""
    K = create_RBF(np.asarray(x, dtype=np.float64).reshape(-1, 1), length_scales=np.exp(pars[1]))
    return _kron_gauss_logpdf(np.asarray(Y, dtype=np.float64), K, L @ L.T, float(np.exp(pars[0])))


def compute_loglik(pars, x, Y):
    """pre_nmgp.py:35-46: pars = (log sigma2_err, log ell, vec(tril L))."""
    D = Y.shape[1]
    L = np.zeros((D, D))
    L[np.tril_indices(D)] = pars[2:]
    return compute_loglik_part(pars[:2], x, Y, L)


def objective(pars, x, Y):
    return -compute_loglik(pars, x, Y)


def objective_part(pars, x, Y, L):
    return -compute_loglik_part(pars, x, Y, L)


def pre_estimation_partial(x, Y, z, P=10):
    """pre_nmgp.py:102-125.  Returns (v_array [M] = log ell at the inducing points, L_tensor [D, D, M], sigma2_err_log [M])."""
    from scipy.optimize import minimize
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    Y = np.asarray(Y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64).reshape(-1)
    N = Y.shape[0]
    est_L = np.linalg.cholesky(Y.T @ Y / (N - 1))
    L_tensor = np.stack([est_L for _ in range(z.shape[0])], axis=-1)
    log_s2, ell = [], []
    for z_local in z:
        x_local, Y_local = search_nearest_neighhood(x, Y, z_local, P=P)
        res = minimize(objective_part, np.array([-6.0, -6.0]), args=(x_local, Y_local, est_L))
        log_s2.append(res.x[0])
        ell.append(np.exp(res.x[1]))
    return np.log(np.array(ell)), L_tensor, np.array(log_s2)

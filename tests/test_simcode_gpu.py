"""SIM_code (exact / Kronecker) line on the GPU against the reference's golden vectors and the CPU oracle.
Tolerances: 1e-12 relative for the element-wise kernel builds, 1e-9 for anything that goes through a
factorisation (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import kernel_specs as specs
from oracle import nmgp_oracle as orc
from tests import golden_util as gu

pytestmark = pytest.mark.gpu

from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops  # noqa: E402
from collaborative_nonstationary_multivariate_gaussian_process_b200 import distributions, kernels, kronecker_operation  # noqa: E402

DEV = "cuda:0"


def d(a):
    return torch.as_tensor(a, dtype=torch.float64).to(DEV).contiguous()


def rel(a, b):
    a = torch.as_tensor(a).detach().cpu().double().reshape(-1); b = torch.as_tensor(b).detach().cpu().double().reshape(-1)
    return float(torch.linalg.norm(a - b) / max(float(torch.linalg.norm(b)), 1e-300))


def test_kernel_builds_match_reference_golden():
    g = gu.load("sim_code")
    x1 = d(g["x1"]).view(-1, 1); x2 = d(g["x2"]).view(-1, 1)
    assert rel(kernels.Nonstationary_RBF_cov(x1, sigma1=d(g["sg1"]), ell1=d(g["ell1"])), g["K_self"]) < 1e-13
    assert rel(kernels.Nonstationary_RBF_cov(x1, d(g["sg1"]), d(g["ell1"]), x2, d(g["sg2"]), d(g["ell2"])), g["K_cross"]) < 1e-13
    assert rel(kernels.Nonstationary_RBF_cov(x1), g["K_def"]) < 1e-13
    assert rel(kernels.RBF_cov(x1, alpha=1.3, beta=0.2), g["R_self"]) < 1e-13
    assert rel(kernels.RBF_cov(x1, x2, alpha=0.7, beta=0.35), g["R_cross"]) < 1e-13
    xc = torch.from_numpy(g["x1"]).view(-1, 1)
    assert rel(kernels.pairwise_distances(x1), orc.sq_dists_gemm(xc)) < 1e-13


@pytest.mark.parametrize("T1,T2", [(130, 130), (64, 64), (257, 131), (70, 200), (1, 3), (1100, 1100)])
def test_nonstationary_cov_tiles_and_adjoint(T1, T2):
    """Tiled builds (ragged tiles, odd row lengths = scalar-store path, mirrored self-covariance) against the CPU
    specification, and the hand-written adjoint w.r.t. sigma / ell against autograd of the specification."""
    gen = torch.Generator().manual_seed(T1 * 7 + T2)
    x1 = torch.sort(torch.rand(T1, generator=gen, dtype=torch.float64))[0].view(-1, 1)
    x2 = torch.sort(torch.rand(T2, generator=gen, dtype=torch.float64))[0].view(-1, 1)
    s1 = 0.5 + torch.rand(T1, generator=gen, dtype=torch.float64); l1 = torch.exp(torch.randn(T1, generator=gen, dtype=torch.float64) - 1.5)
    s2 = 0.5 + torch.rand(T2, generator=gen, dtype=torch.float64); l2 = torch.exp(torch.randn(T2, generator=gen, dtype=torch.float64) - 1.5)
    Kc = kernels.Nonstationary_RBF_cov(d(x1), d(s1), d(l1), d(x2), d(s2), d(l2))
    assert rel(Kc, specs.nonstationary_cov(x1, s1, l1, x2, s2, l2, 0.0)) < 1e-13
    Ks = kernels.Nonstationary_RBF_cov(d(x1), d(s1), d(l1))
    assert rel(Ks, specs.nonstationary_cov(x1, s1, l1, x1, s1, l1, 1e-6)) < 1e-13
    assert torch.equal(Ks, Ks.t()), "a self-covariance must be exactly symmetric"
    assert rel(kernels.RBF_cov(d(x1), alpha=1.7, beta=0.3), specs.sim_rbf_cov(x1, x1, 1.7, 0.3, 1e-6)) < 1e-13
    assert rel(kernels.RBF_cov(d(x1), d(x2), alpha=0.9, beta=0.4), specs.sim_rbf_cov(x1, x2, 0.9, 0.4, 0.0)) < 1e-13
    # adjoint, cross-covariance
    Kb = torch.randn(T1, T2, generator=gen, dtype=torch.float64)
    got = ops.nonstationary_cov_bwd(d(x1), d(s1), d(l1), d(x2), d(s2), d(l2), d(Kb))
    ref = specs.nonstationary_cov_bwd(x1, s1, l1, x2, s2, l2, Kb)
    for a, b, n in zip(got, ref, ("sigma1", "ell1", "sigma2", "ell2")):
        assert rel(a, b) < 1e-11, n
    # autograd through the drop-in, self-covariance (row- and column-side contributions added)
    sv = d(s1).requires_grad_(True); lv = d(l1).requires_grad_(True)
    Kb2 = torch.randn(T1, T1, generator=gen, dtype=torch.float64)
    (kernels.Nonstationary_RBF_cov(d(x1), sv, lv) * d(Kb2)).sum().backward()
    scpu = s1.clone().requires_grad_(True); lcpu = l1.clone().requires_grad_(True)
    (specs.nonstationary_cov(x1, scpu, lcpu, x1, scpu, lcpu, 1e-6) * Kb2).sum().backward()
    assert rel(sv.grad, scpu.grad) < 1e-11
    # (T = 1: the exact gradient w.r.t. ell is 0 -- compare on the scale of the sigma gradient)
    assert float(torch.linalg.norm(lv.grad.cpu() - lcpu.grad)) < 1e-11 * max(float(torch.linalg.norm(lcpu.grad)), float(torch.linalg.norm(scpu.grad)))


def test_kronecker_and_logpdf_match_reference_golden():
    g = gu.load("sim_code")
    K = d(g["K_self"]); Bf = d(g["Bf"]); y = d(g["y"]); mu = d(g["mu"]); s2 = torch.tensor(float(g["s2"]), dtype=torch.float64)
    assert rel(kronecker_operation.kron_mv(Bf, K, y), g["kron_mv"]) < 1e-12
    assert rel(kronecker_operation.kronecker_product(Bf, K[:5, :4].contiguous()), g["kron_prod"]) < 1e-14
    assert rel(kronecker_operation.kronecker_product_diag(torch.diagonal(Bf).contiguous(), torch.diagonal(K).contiguous()), g["kron_diag"]) < 1e-14
    ld = float(kronecker_operation.kron_logdet(s2, Bf, K).cpu())
    assert abs(ld - float(g["kron_logdet"])) <= 1e-9 * abs(float(g["kron_logdet"]))
    inv = kronecker_operation.kron_inv(s2, Bf, K).cpu()
    assert np.allclose(torch.diagonal(inv).numpy(), g["kron_inv_diag"], rtol=1e-8)
    assert np.allclose(inv[7].numpy(), g["kron_inv_row7"], rtol=1e-7, atol=1e-8)
    lp0 = float(distributions.multivariate_normal_logpdf0(y, mu, Bf, K, s2).cpu())
    lp2 = float(distributions.multivariate_normal_logpdf2(y, mu, Bf, K, s2).cpu())
    assert abs(lp0 - float(g["logpdf0"])) <= 1e-9 * abs(lp0), (lp0, float(g["logpdf0"]))
    assert abs(lp2 - float(g["logpdf2"])) <= 1e-9 * abs(lp2)
    torch.manual_seed(9)
    lp1 = float(distributions.multivariate_normal_logpdf1(y, mu, Bf, K, s2).cpu())
    assert abs(lp1 - float(g["logpdf1"])) <= 1e-9 * abs(lp1)
    # dense-path helper (distributions.py:10-23)
    lp = float(distributions.multivariate_normal_logpdf(y, mu, torch.tensor(ld, device=DEV), d(inv)).cpu())
    assert abs(lp - float(g["logpdf2"])) <= 1e-8 * abs(lp)


def test_t200_logpdf():
    g = gu.load("sim_code_t200")
    x1 = d(g["x1"]).view(-1, 1)
    K = kernels.Nonstationary_RBF_cov(x1, ell1=d(g["ell1"]))
    assert rel(K[17], g["K_row17"]) < 1e-13
    y = d(g["y"]); Bf = d(g["Bf"]); s2 = torch.tensor(float(g["s2"]), dtype=torch.float64)
    lp0 = float(distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), Bf, K, s2).cpu())
    assert abs(lp0 - float(g["logpdf0"])) <= 1e-9 * abs(lp0), (lp0, float(g["logpdf0"]))


# even K (16-byte aligned rows): the TMA-fed kernel; odd K: the cp.async kernel (no tensor map possible)
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 77, 45), (1, 5, 3), (513, 260, 129), (512, 256, 128), (300, 78, 46),
                                   (1000, 130, 34), (129, 65, 2), (70, 300, 1000), (2048, 2048, 256)])
def test_gemm_nt(M, N, K):
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen, dtype=torch.float64); B = torch.randn(N, K, generator=gen, dtype=torch.float64)
    assert rel(ops.gemm_nt(d(A), d(B)), A @ B.t()) < 1e-14
    C = torch.randn(M, N, generator=gen, dtype=torch.float64); Cd = d(C.clone())
    ops.gemm_nt(d(A), d(B), alpha=-0.5, beta=2.0, C=Cd)
    assert rel(Cd, -0.5 * (A @ B.t()) + 2.0 * C) < 1e-14


def test_gemm_nt_on_views_of_a_larger_matrix():
    """Row-strided blocks (what the blocked triangular inverse and the Cholesky panels pass): tensor maps / leading
    dimensions of the parent matrix, output written into a block of another matrix."""
    gen = torch.Generator().manual_seed(3)
    P = torch.randn(700, 900, generator=gen, dtype=torch.float64); Q_ = torch.randn(500, 900, generator=gen, dtype=torch.float64)
    O = torch.randn(800, 600, generator=gen, dtype=torch.float64)
    Pd, Qd, Od = d(P), d(Q_), d(O.clone())
    ops.gemm_nt(Pd[100:420, 64:576], Qd[6:206, 64:576], alpha=1.5, beta=-1.0, C=Od[32:352, 100:300])
    ref = O.clone()
    ref[32:352, 100:300] = 1.5 * (P[100:420, 64:576] @ Q_[6:206, 64:576].t()) - O[32:352, 100:300]
    assert rel(Od, ref) < 1e-14


@pytest.mark.parametrize("T,pb", [(7, 0), (128, 0), (200, 0), (1000, 0), (1537, 0), (2500, 0), (1537, 256), (2100, 512)])
def test_potrf_big_and_solve(T, pb, monkeypatch):
    if pb:
        monkeypatch.setenv("NMGP_POTRF_PB", str(pb))   # panel-width knob: exercises the multi-block left-looking panels
    gen = torch.Generator().manual_seed(T)
    x = torch.sort(torch.rand(T, generator=gen, dtype=torch.float64))[0].view(-1, 1)
    K = orc.sim_nonstationary_cov(x, ell1=torch.exp(3 * (x.view(-1) - 1) ** 3 - 2.0)) + 1e-2 * torch.eye(T, dtype=torch.float64)
    L, hld = ops.potrf_big(d(K.clone()))
    Lr = torch.linalg.cholesky(K)
    assert rel(L, Lr) < 1e-9
    assert abs(float(hld.cpu()) - float(Lr.diagonal().log().sum())) <= 1e-11 * max(1.0, abs(float(Lr.diagonal().log().sum())))
    assert float(torch.triu(L, 1).abs().max()) == 0.0
    b = torch.randn(T, generator=gen, dtype=torch.float64)
    xs = ops.potrs_vec(d(Lr), d(b))
    assert rel(xs, torch.cholesky_solve(b.view(-1, 1), Lr).view(-1)) < 1e-9
    # size-independent property: residual of the factorisation itself
    Ld = L.cpu()
    assert rel(Ld @ Ld.t(), K) < 1e-13


def test_potrf_big_raises_on_non_pd():
    A = torch.eye(300, dtype=torch.float64); A[250, 250] = -1.0
    with pytest.raises(RuntimeError):
        ops.potrf_big(d(A))
    A = torch.eye(1500, dtype=torch.float64); A[1333, 1333] = -1.0
    with pytest.raises(RuntimeError, match="1334"):
        ops.potrf_big(d(A))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 64, 127, 128])
def test_eigh_small(n):
    gen = torch.Generator().manual_seed(n)
    L = torch.tril(torch.randn(n, n, generator=gen, dtype=torch.float64)); A = L @ L.t()
    w, V = ops.eigh_small(d(A))
    wr, _ = torch.linalg.eigh(A)
    assert rel(w, wr) < 1e-12
    Vc = V.cpu()
    assert rel(Vc @ torch.diag(w.cpu()) @ Vc.t(), A) < 1e-12
    assert rel(Vc.t() @ Vc, torch.eye(n, dtype=torch.float64)) < 1e-12


def test_hadamard_index_cov():
    """out[i,j] = Kx[i,j] Bf[indx1[i], indx2[j]] (+ diag): logpos.generate_K_index fused with the Hadamard product."""
    from oracle import kernel_specs as specs
    gen = torch.Generator().manual_seed(8)
    for (n1, n2, M) in ((37, 37, 3), (300, 1, 5), (5, 70, 2)):
        Kx = torch.randn(n1, n2, generator=gen, dtype=torch.float64)
        L = torch.randn(M, M, generator=gen, dtype=torch.float64); Bf = L @ L.t()
        i1 = torch.randint(0, M, (n1,), generator=gen).to(torch.int32); i2 = torch.randint(0, M, (n2,), generator=gen).to(torch.int32)
        got = ops.hadamard_index_cov(d(Kx), d(Bf), i1.cuda(), i2.cuda(), 0.25)
        assert rel(got, specs.hadamard_index_cov(Kx, Bf, i1, i2, 0.25)) < 1e-15
    # non-square weight table (rows scaled by a per-row factor: Bf is [R, 1], all column indices 0)
    Kx = torch.randn(40, 3, generator=gen, dtype=torch.float64); Bf = torch.randn(8, 1, generator=gen, dtype=torch.float64)
    i1 = torch.randint(0, 8, (40,), generator=gen).to(torch.int32); i2 = torch.zeros(3, dtype=torch.int32)
    assert rel(ops.hadamard_index_cov(d(Kx), d(Bf), i1.cuda(), i2.cuda(), 0.0), specs.hadamard_index_cov(Kx, Bf, i1, i2, 0.0)) < 1e-15


@pytest.mark.parametrize("T", [8192, 12288])
def test_potrf_big_native_panels_at_sweep_sizes(T):
    """The 256 / 512-wide panels the factorisation picks by itself above T = 6144 / 12288 (BASELINE config 5 sizes),
    checked on the device: residual of L L^T against A (size-independent property) and log-determinant against
    cuSOLVER's factor (a yardstick, not on the product path)."""
    gen = torch.Generator().manual_seed(T)
    x = torch.sort(torch.rand(T, generator=gen, dtype=torch.float64))[0].view(-1, 1)
    K = kernels.Nonstationary_RBF_cov(d(x), ell1=d(torch.exp(3 * (x.view(-1) - 1) ** 3 - 3.0)))
    A = ops.scale_add_diag(K, 1.0, 1e-2)
    L, hld = ops.potrf_big(A.clone())
    R = ops.gemm_nt(L, L)                                     # L L^T on the tensor cores
    assert float(torch.linalg.norm(R - A) / torch.linalg.norm(A)) < 1e-13
    Lr = torch.linalg.cholesky(A)
    ref = float(Lr.diagonal().log().sum())
    assert abs(float(hld) - ref) <= 1e-11 * abs(ref)
    assert float(torch.linalg.norm(L - Lr) / torch.linalg.norm(Lr)) < 1e-9


@pytest.mark.parametrize("T,D", [(2048, 3), (4096, 2), (8192, 2)])
def test_kron_logpdf0_matches_the_eigen_route_at_sweep_sizes(T, D):
    """multivariate_normal_logpdf0 through the eigen-block Cholesky pipeline (augmented systems, blocks in flight on
    several streams) against the reference's own route -- symeig of both factors (distributions.py:37-51), evaluated
    here with torch.linalg.eigh on the device -- at the sizes of the scale sweep.  1e-9 relative."""
    gen = torch.Generator().manual_seed(T + D)
    x = torch.sort(torch.rand(T, generator=gen, dtype=torch.float64))[0].view(-1, 1)
    K = kernels.Nonstationary_RBF_cov(d(x), ell1=d(torch.exp(3 * (x.view(-1) - 1) ** 3 - 3.0)))
    Lb = torch.tril(torch.randn(D, D, generator=gen, dtype=torch.float64)); Bf = d(Lb @ Lb.t() / D)
    y = d(torch.randn(D * T, generator=gen, dtype=torch.float64))
    s2 = torch.tensor(1e-2, dtype=torch.float64)
    got = float(distributions.multivariate_normal_logpdf0(y, torch.zeros_like(y), Bf, K, s2))
    wB, VB = torch.linalg.eigh(Bf, UPLO="U")
    wK, VK = torch.linalg.eigh(K, UPLO="U")
    a = (VK.t() @ y.view(D, T).t() @ VB).t().reshape(-1)        # kron_mv(V_B^T, V_K^T, y)
    tt = torch.kron(wB, wK) + float(s2)
    ref = float(-0.5 * torch.log(tt).sum() - 0.5 * (a * a / tt).sum())
    assert abs(got - ref) <= 1e-9 * abs(ref), (got, ref)
    # the adjoint-capable path (explicit solves) gives the same value
    yv = y.clone().requires_grad_(True)
    got2 = float(distributions.multivariate_normal_logpdf0(yv, torch.zeros_like(y), Bf, K, s2))
    assert abs(got2 - ref) <= 1e-9 * abs(ref), (got2, ref)


def test_nonstationary_cov_adjoint_multidimensional_inputs():
    """dx > 1 (no reference call site uses it, but kernels.Nonstationary_RBF_cov accepts it): forward and the (sigma, ell)
    adjoint against autograd of the specification."""
    gen = torch.Generator().manual_seed(11)
    T1, T2, dx = 70, 45, 3
    x1 = torch.rand(T1, dx, generator=gen, dtype=torch.float64); x2 = torch.rand(T2, dx, generator=gen, dtype=torch.float64)
    s1 = 0.5 + torch.rand(T1, generator=gen, dtype=torch.float64); l1 = torch.exp(torch.randn(T1, generator=gen, dtype=torch.float64) - 1.0)
    s2 = 0.5 + torch.rand(T2, generator=gen, dtype=torch.float64); l2 = torch.exp(torch.randn(T2, generator=gen, dtype=torch.float64) - 1.0)
    assert rel(kernels.Nonstationary_RBF_cov(d(x1), d(s1), d(l1), d(x2), d(s2), d(l2)),
               specs.nonstationary_cov(x1, s1, l1, x2, s2, l2, 0.0)) < 1e-13
    Kb = torch.randn(T1, T2, generator=gen, dtype=torch.float64)
    got = ops.nonstationary_cov_bwd(d(x1), d(s1), d(l1), d(x2), d(s2), d(l2), d(Kb))
    ref = specs.nonstationary_cov_bwd(x1, s1, l1, x2, s2, l2, Kb)
    for a, b, n in zip(got, ref, ("sigma1", "ell1", "sigma2", "ell2")):
        assert rel(a, b) < 1e-11, n
    sv = d(s1).requires_grad_(True); lv = d(l1).requires_grad_(True)
    Kb2 = torch.randn(T1, T1, generator=gen, dtype=torch.float64)
    (kernels.Nonstationary_RBF_cov(d(x1), sv, lv) * d(Kb2)).sum().backward()
    scpu = s1.clone().requires_grad_(True); lcpu = l1.clone().requires_grad_(True)
    (specs.nonstationary_cov(x1, scpu, lcpu, x1, scpu, lcpu, 1e-6) * Kb2).sum().backward()
    assert rel(sv.grad, scpu.grad) < 1e-11 and rel(lv.grad, lcpu.grad) < 1e-11


@pytest.mark.parametrize("N,M,table", [(37, 3, "small"), (300, 5, "small"), (90, 90, "identity")])
def test_dense_loglik_bwd(N, M, table):
    """Adjoint kernel of the dense indexed log-likelihood (shared-memory privatised table / global atomics) vs its spec."""
    gen = torch.Generator().manual_seed(N + M)
    A = torch.randn(N, N, generator=gen, dtype=torch.float64); A = A @ A.t() / N
    Bt = torch.randn(M, M, generator=gen, dtype=torch.float64); Bt = Bt @ Bt.t() / M
    if table == "identity":
        i1 = torch.arange(N, dtype=torch.int32)
    else:
        i1 = torch.randint(0, M, (N,), generator=gen, dtype=torch.int32)
    Sinv = torch.randn(N, N, generator=gen, dtype=torch.float64); Sinv = Sinv + Sinv.t()
    alpha = torch.randn(N, generator=gen, dtype=torch.float64)
    g = torch.tensor([0.7], dtype=torch.float64)
    got = ops.dense_loglik_bwd(d(Sinv), d(alpha), d(A), d(Bt), i1.cuda(), i1.cuda(), d(g))
    ref = specs.dense_loglik_bwd(Sinv, alpha, A, Bt, i1, i1, g)
    for a, b, n in zip(got, ref, ("Abar", "Btbar", "s2bar")):
        assert rel(a, b) < 1e-12, n


def test_concurrent_eigen_block_pipeline_is_reproducible_and_exact():
    """Several eigen-block factorisations in flight (streams / scratch slots) must give bit-identical results run to run
    and agree with a plain torch Cholesky per block.  (With TMA-fed GEMMs in the concurrent pipeline a T = 12288 run was
    not reproducible -- relative 2e-7 -- which is why that pipeline stages its GEMM operands with cp.async.)"""
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import kronecker_operation as ko
    T, D = 6400, 6
    gen = torch.Generator().manual_seed(5)
    x = torch.sort(torch.rand(T, generator=gen, dtype=torch.float64))[0].view(-1, 1)
    ell = torch.exp(3 * (x.view(-1) - 1) ** 3 - 3.0)
    Lb = torch.tril(torch.randn(D, D, generator=gen, dtype=torch.float64)); Bf = Lb @ Lb.t() / D
    y = torch.randn(D * T, generator=gen, dtype=torch.float64)
    s2 = torch.tensor(1e-2, dtype=torch.float64)
    K = kernels.Nonstationary_RBF_cov(d(x), ell1=d(ell))
    lam, V = ops.eigh_small(d(Bf))
    Rt = (V.t() @ d(y).view(D, T)).contiguous()
    runs = []
    for _ in range(4):
        res = ko.block_pipeline(s2, d(Bf), K, Rt=Rt, want_alpha=False)
        torch.cuda.synchronize()
        assert int(res["info"].abs().sum()) == 0
        runs.append((res["hld"].clone(), res["quad"].clone()))
    for h, q in runs[1:]:
        assert torch.equal(h, runs[0][0]) and torch.equal(q, runs[0][1])
    for m in range(D):
        A = K * lam[m] + torch.eye(T, dtype=torch.float64, device=DEV) * 1e-2
        L = torch.linalg.cholesky(A)
        z = torch.linalg.solve_triangular(L, Rt[m].view(-1, 1), upper=False)
        href, qref = float(torch.log(torch.diagonal(L)).sum()), float((z * z).sum())
        assert abs(float(runs[0][0][m]) - href) <= 1e-12 * abs(href)
        assert abs(float(runs[0][1][m]) - qref) <= 1e-11 * abs(qref)

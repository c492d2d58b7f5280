/*
 * nmgp_b200.h -- C ABI of libnmgp_b200.so: the B200 (sm_100a, FP64) kernels behind the NMGP DSVI hot path.
 *
 * The reference (Corleno/Collaborative_Nonstationary_Multivariate_Gaussian_Process) is pure Python/torch and has no
 * FFI; its boundary is the Python call surface of code/utils.py, code/nmgp_dsvi.py and code/SIM_code/Utility/*.py.
 * Each entry point below names the reference site(s) it replaces.  INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major data; `double` unless noted; `int` is 32-bit;
 *   - sizes: B rows in the minibatch, Q inducing points (<= 112), D outputs, ns Monte-Carlo samples of a chunk,
 *     nb / np batch counts of Q x Q matrices;
 *   - `hyp` is the vector of exponentiated hyper-parameters in the order
 *     {s2_tildeell, len_tildeell, s2_L0, len_L0, s2_L1, len_L1, s2_err} (code/nmgp_dsvi.py:180-188);
 *     `ghyp` accumulates gradients w.r.t. the corresponding *_log parameters;
 *   - rows are sorted by output id I[n]; seg[d] = first row with I >= d, seg[D] = B;
 *   - "+=" marks accumulating outputs (caller zero-initialises), "=" overwriting ones;
 *   - all work is enqueued on `stream` (a cudaStream_t); no per-call allocation.  Two entry points keep a small
 *     library-owned scratch per process, created on first use and reused: nmgp_latent_fused (padded copies of
 *     Sigma_W / mu_W, D * ~25 KB) and nmgp_potrf_big (two 128 x 128 inverse blocks, a helper stream and two events for
 *     the look-ahead); they are therefore not re-entrant from several host threads;
 *   - return value: 0 ok, < 0 argument/launch error, > 0 numerical failure; nmgp_last_error() gives the text
 *     (thread-local).
 */
#ifndef NMGP_B200_H
#define NMGP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* nmgp_stream_t; /* cudaStream_t */

const char* nmgp_last_error(void);
int nmgp_version(void);
/* number of kernel launches this library has issued since it was loaded (host-side counter; measurement aid) */
unsigned long long nmgp_launch_count(void);

/* hyp[i] = exp(logs[i])                                                    code/nmgp_dsvi.py:180-188 */
int nmgp_hyper_exp(const double* logs, double* hyp, int n, nmgp_stream_t stream);
/* seg[0..D] from the sorted output ids                                     code/nmgp_dsvi.py:163-167 */
int nmgp_segment_offsets(const int* I, int* seg, long long B, int D, nmgp_stream_t stream);

/* Sigma = tril(S) tril(S)^T, batched                                       utils.py:68-72 + nmgp_dsvi.py:172-177 */
int nmgp_tril_syrk_fwd(const double* S, double* Sigma, int nb, int Q, nmgp_stream_t stream);
int nmgp_tril_syrk_bwd(const double* S, const double* SigBar, double* Sbar, int nb, int Q, nmgp_stream_t stream);

/* C = chol(A + jitter I) (lower), hld = sum log diag C; *info = 1 + index of a non-PD matrix (0 if none)
 *                                                                          utils.py:46,276,347-348 (torch.cholesky) */
int nmgp_potrf_batched(const double* A, double jitter, double* C, double* hld, int* info, int nb, int Q,
                       nmgp_stream_t stream);
/* Abar = sym(C^-T Phi(C^T Cbar') C^-1), Cbar' = tril(Cbar) + diag(hldbar / diag C)   (autograd of the above) */
int nmgp_potrf_bwd_batched(const double* C, const double* Cbar, const double* hldbar, double* Abar, int nb, int Q,
                           nmgp_stream_t stream);
/* same adjoint for 64 < Q <= 128 on the FP64 tensor cores: one A^T B reduction and two right solves "rows x C^-1" with
 * the DMMA row-solve kernel; work1, work2: [nb,Q,Q] scratch */
int nmgp_potrf_bwd_batched_lq(const double* C, const double* Cbar, const double* hldbar, double* Abar, double* work1,
                              double* work2, int nb, int Q, nmgp_stream_t stream);

/* kl[p,b] = KL(N(mu_b, CS_b CS_b^T) || N(0, R_p R_p^T)) in the reference's form (quirk q10), every (p,b) pair in
 * parallel: rs[b,a] = sum_{c<=a} CS_b[a,c]^2; t[p,b,:] = (R_p R_p^T)^-1 mu_b by the DMMA row solve (saved for the
 * adjoint); work [np,nb,Q] is scratch.                                      utils.py:332-351 KL_Gaussian */
int nmgp_kl_fwd(const double* CS, const double* hldS, const double* mu, const double* R, const double* hldR,
                double* kl, double* t, double* rs, double* work, int np, int nb, int Q, nmgp_stream_t stream);
/* adjoint of the above; work [np,nb,Q], rsb [nb,Q], G [np,Q,Q] are scratch */
int nmgp_kl_bwd(const double* klbar, const double* CS, const double* R, const double* t, const double* rs,
                double* CSbar, double* hldSbar, double* mubar, double* Rbar /* += */, double* hldRbar, double* work,
                double* rsb, double* G, int np, int nb, int Q, nmgp_stream_t stream);

/* building blocks of the mathematically exact KL variant (explicit flag; the reference's utils.py:349 only uses
 * diag(chol(K)), quirk q10): C[s] += sign A[s]^T B[s] (A, B: [ns,B,Q]) and Rbar_p += -tril(G_p R_p) */
int nmgp_atb(const double* A, const double* Bm, double* C /* += */, double sign, int ns, long long B, int Q,
             nmgp_stream_t stream);
int nmgp_kl_rbar(const double* R, const double* G, double* Rbar /* += */, int np, int Q, nmgp_stream_t stream);

/* K[n,q] = hyp[is2] exp(-(x_n/len - z_q/len)^2 / 2) (+ jitter on n == q)    utils.py:75-94 create_RBF */
int nmgp_rbf_build_fwd(const double* x, const double* z, const double* hyp, int is2, int ilen, double jitter,
                       double* K, long long B, int Q, nmgp_stream_t stream);
int nmgp_rbf_build_bwd(const double* x, const double* z, const double* hyp, int is2, int ilen, const double* Kbar,
                       double* ghyp /* += */, long long B, int Q, nmgp_stream_t stream);

/* K[s,n,q] = sqrt(2ab/(a^2+b^2)) exp(-(x_n-z_q)^2/(a^2+b^2)), a = ellx[s,n], b = ellz[s,q]   utils.py:97-103 create_Gibbs */
int nmgp_gibbs_build_fwd(const double* x, const double* z, const double* ellx, const double* ellz, double jitter,
                         double* K, int ns, long long B, int Q, nmgp_stream_t stream);
/* Kfwd: the forward values K (built with jitter 0) if the caller still holds them, else NULL (they are recomputed) */
int nmgp_gibbs_build_bwd(const double* x, const double* z, const double* ellx, const double* ellz, const double* Kbar,
                         const double* Kfwd, double* ellxbar /* = */, double* ellzbar /* += */, int ns, long long B,
                         int Q, nmgp_stream_t stream);

/* P = K (R R^T)^-1, c = rowsum(P o K)                                      utils.py:117-122 (torch.solve of K22 + eps I) */
int nmgp_solve_rows_fwd(const double* K, const double* R, double* P, double* c, int ns, long long B, int Q,
                        nmgp_stream_t stream);
int nmgp_solve_rows_bwd(const double* Pbar, const double* cbar, const double* K, const double* P, const double* R,
                        double* Kbar /* = */, double* Abar /* += */, double* work /* [ns,B,Q] scratch, Q <= 64 */,
                        int ns, long long B, int Q, nmgp_stream_t stream);

/* q[s,n,j] = p^T Sig[idx] p, m[s,n,j] = p . Mu[idx] for j <= I[n]          utils.py:120-122,143-144 (MGP_d, MGP_mu_sigma2)
 * mode 0 (latent functions): idx = j; mode 1 (coefficients): idx = packed pair (I[n], j), p = Pb row if j == I[n] */
int nmgp_quadform_fwd(const double* Pa, const double* Pb, const int* I, const int* seg, const double* Sig,
                      const double* Mu, double* q, double* m, int ns, long long B, int Q, int D, int mode,
                      nmgp_stream_t stream);
int nmgp_quadform_bwd(const double* Pa, const double* Pb, const int* I, const int* seg, const double* Sig,
                      const double* Mu, const double* qbar, const double* mbar, double* Pabar, double* Pbbar, int ns,
                      long long B, int Q, int D, int mode, nmgp_stream_t stream);
int nmgp_weighted_gram(const double* Pa, const double* Pb, const int* I, const int* seg, const double* qbar,
                       const double* mbar, double* SigBar /* += */, double* MuBar /* += */, int ns, long long B, int Q,
                       int D, int mode, nmgp_stream_t stream);

/* quadform_fwd (mode 0) + lik_rows + quadform_bwd in one pass over 128-row tiles on the FP64 tensor cores
 * (DMMA) for Q <= 64; larger Q runs the three kernels (work_q, work_m: [ns,B,D] scratch, may be NULL if Q <= 64)
 *                                                                          utils.py:143-144 + nmgp_dsvi.py:255-258 */
int nmgp_latent_fused(const double* PG, const double* cG, const double* l, const double* y, const int* I,
                      const int* seg, const double* SigW, const double* muW, const double* hyp, double scale,
                      double* Rsum /* += */, double* ghyp /* += */, double* lbar, double* mgbar, double* qgbar,
                      double* cGbar, double* PGbar, double* work_q, double* work_m, int ns, long long B, int Q, int D,
                      long long y_stride /* 0: y[B] shared by the samples; >= B: y[ns, y_stride], one target vector
                                            per sample (the subjects of an HCP-style step) */,
                      nmgp_stream_t stream);

/* v = mu_v + C_v z, ellz = exp(v)                                          utils.py:225-227, nmgp_dsvi.py:215 */
int nmgp_sample_v_fwd(const double* mu_v, const double* Cv, const double* zv, double* v, double* ellz, int S, int Q,
                      nmgp_stream_t stream);
int nmgp_sample_v_bwd(const double* ellzbar, const double* vbar, const double* ellz, const double* zv,
                      double* mu_v_bar /* += */, double* Cvbar /* += */, int S, int Q, nmgp_stream_t stream);

/* sd = sqrt(s2_tildeell - c + eps)                                         utils.py:233 + :32 */
int nmgp_ell_sd_fwd(const double* c, const double* hyp, double* sd, long long B, nmgp_stream_t stream);
int nmgp_ell_sd_bwd(const double* sdbar, const double* sd, const double* hyp, double* ghyp, double* cbar, long long B,
                    nmgp_stream_t stream);
/* ellx[s,n] = exp(P_ell[n,:] . v[s,:] + z[s,n] sd[n])                      utils.py:231-235, nmgp_dsvi.py:216 */
int nmgp_ell_rows_fwd(const double* Pell, const double* v, const double* zell, const double* sd, double* ellx, int ns,
                      long long B, int Q, nmgp_stream_t stream);
int nmgp_ell_rows_bwd(const double* ellxbar, const double* ellx, const double* Pell, const double* v,
                      const double* zell, double* vbar /* += */, double* Pellbar /* += */, double* sdbar /* += */,
                      int ns, long long B, int Q, nmgp_stream_t stream);

/* sd[n,j] = sqrt(s2_k - c_k[n] + q[n,j] + eps)                             utils.py:122 + :32 (inside MGP_d) */
int nmgp_coef_sd_fwd(const double* q, const double* cL0, const double* cL1, const int* I, const double* hyp,
                     double* sd, long long B, int D, nmgp_stream_t stream);
int nmgp_coef_sd_bwd(const double* sdbar, const double* sd, const int* I, const double* hyp, double* ghyp,
                     double* qbar, double* cL0bar, double* cL1bar, long long B, int D, nmgp_stream_t stream);
/* l[s,n,j] = m + z sd (exp on the diagonal coefficient)                    nmgp_dsvi.py:228-238
 * zL == NULL: the N(0,1) draws are generated inside the kernel by the counter-based generator of nmgp_noise_fill
 * from (seed, stream_id, s0 + s, gid[n] (or n), j) and never stored; the backward kernel regenerates them. */
int nmgp_coef_sample_fwd(const double* m, const double* sd, const double* zL, const int* I, double* l, int ns,
                         long long B, int D, unsigned long long seed, unsigned long long stream_id, int s0,
                         const long long* gid, const unsigned long long* step_dev, nmgp_stream_t stream);
int nmgp_coef_sample_bwd(const double* lbar, const double* l, const double* zL, const int* I, double* mbar /* += */,
                         double* sdbar /* += */, int ns, long long B, int D, unsigned long long seed,
                         unsigned long long stream_id, int s0, const long long* gid, const unsigned long long* step_dev,
                         nmgp_stream_t stream);
/* out[s,n,c] = N(0,1), float32 draws widened to double (quirk q2: utils.py:123,226,234), Philox4x32-10 keyed by
 * (seed, stream_id) with counter (gid[n] or n, s0 + s, c/4): independent of how rows are sharded over ranks */
int nmgp_noise_fill(double* out, int ns, long long B, int C, unsigned long long seed, unsigned long long stream_id,
                    int s0, const long long* gid, const unsigned long long* step_dev, nmgp_stream_t stream);
/* step_dev (may be NULL): DEVICE step counter; the effective stream id is stream_id | (*step_dev << 8).  A step that
 * is replayed from a CUDA graph passes stream_id = kind and increments the counter inside the graph. */

/* ---- 64 < Q <= 128 (the PM2.5 / HCP drivers use Q = 100): ring-pipelined DMMA kernels, csrc/nmgp_quadform_lq.cu ----
 * The Q x Q covariances are first copied into padded, half-split records (the order the kernels stream them through
 * shared memory with cp.async.bulk); nmgp_lq_record_doubles(Q) doubles per record, 0 if Q is outside 65..128. */
long long nmgp_lq_record_doubles(int Q);
int nmgp_lq_pad_records(const double* Sig /* [n,Q,Q] */, double* rec /* [n, record_doubles] */, int n, int Q,
                        nmgp_stream_t stream);
/* = nmgp_latent_fused with recW = padded Sigma_W                           utils.py:143-144 + nmgp_dsvi.py:255-258 */
int nmgp_lq_latent_fused(const double* PG, const double* cG, const double* l, const double* y, const int* I,
                         const double* recW, const double* muW, const double* hyp, double scale, double* Rsum /* += */,
                         double* ghyp /* += */, double* lbar, double* mgbar, double* qgbar, double* cGbar, double* PGbar,
                         int ns, long long B, int Q, int D, long long y_stride, nmgp_stream_t stream);
/* = nmgp_quadform_fwd (bwd == 0: q, m; entries of pairs a row does not consume stay as the caller initialised them)
 * / nmgp_quadform_bwd (bwd != 0: Pabar, Pbbar) in coefficient mode, recU = padded Sigma_U[packed pairs]
 *                                                                          utils.py:120-122 in the loop nmgp_dsvi.py:228-237 */
int nmgp_lq_coef_quadform(int bwd, const double* Pa, const double* Pb, const int* I, const double* recU,
                          const double* Mu, double* q, double* m, const double* qbar, const double* mbar,
                          double* Pabar, double* Pbbar, int ns, long long B, int Q, int D, nmgp_stream_t stream);

/* expected log-likelihood of a sample chunk and its cotangents             nmgp_dsvi.py:255-258, utils.py:268-272 */
int nmgp_lik_rows(const double* l, const double* mg, const double* qg, const double* cG, const double* y, const int* I,
                  const double* hyp, double scale, double* Rsum /* += */, double* ghyp /* += */, double* lbar,
                  double* mgbar, double* qgbar, double* cGbar, int ns, long long B, int D, long long y_stride,
                  nmgp_stream_t stream);

/* means only: m[s,n,j] = p . Mu[idx], j <= I[n]                            utils.py:149-157 MGP_mu (predict_Y) */
int nmgp_pair_means(const double* Pa, const double* Pb, const int* I, const double* Mu, double* m, int ns, long long B,
                    int Q, int D, int mode, nmgp_stream_t stream);
/* F[s,n] = sum_{j <= I[n]} l[s,n,j] g[s,n,j]                               nmgp_dsvi.py:255, :721-722 */
int nmgp_rowdot_live(const double* l, const double* g, const int* I, double* F, int ns, long long B, int D,
                     nmgp_stream_t stream);

/* small utilities: reparameterize (diagonal), Normal_logprob, batch_trace_XXT      utils.py:15-33, 268-287 */
int nmgp_reparam_diag(const double* mean, const double* var, const double* z, double* out, long long n,
                      nmgp_stream_t stream);
int nmgp_normal_logprob_sum(const double* loc, const double* scale /* device scalar */, const double* y,
                            double* out /* += */, long long n, nmgp_stream_t stream);
int nmgp_sumsq_rows(const double* x, double* out, long long rows, long long cols, nmgp_stream_t stream);

/* per-point output correlation matrices from the sampled mixing factors      nmgp_dsvi.py:567-569 (sample_FY) */
int nmgp_lcorr(const double* L, double* corr, long long nmat, int D, nmgp_stream_t stream);

/* SIM_code line: code/SIM_code/Utility/kernels.py:46-73 Nonstationary_RBF_cov and :24-43 RBF_cov.
 * sigma/ell pointers may be NULL (= ones).  self != 0: the reference's X2=None call (X2, sigma2, ell2 are X1, sigma1,
 * ell1): only the 64 x 64 tiles on and below the diagonal are evaluated, each is written twice (itself and mirrored,
 * 16-byte coalesced stores through shared memory) and `jitter` (1e-6) is added on the diagonal.  dx == 1 (every
 * reference call site, quirk q11) takes the tiled path: one rsqrt + one exp per entry. */
int nmgp_nonstationary_cov(const double* X1, const double* sigma1, const double* ell1, const double* X2,
                           const double* sigma2, const double* ell2, double jitter, double* K, long long T1,
                           long long T2, int dx, int self, nmgp_stream_t stream);
int nmgp_sim_rbf_cov(const double* X1, const double* X2, double alpha, double beta, double jitter, double* K,
                     long long T1, long long T2, int dx, int self, nmgp_stream_t stream);
/* adjoint of nmgp_nonstationary_cov w.r.t. the per-point sigma / ell (g_* +=, any may be NULL; any dx): what autograd
 * gives the reference when logpos.nlogpos_obj* are differentiated (logpos.py:216-296).  For a
 * self-covariance the caller adds the row-side (g_*1) and column-side (g_*2) results. */
int nmgp_nonstationary_cov_bwd(const double* X1, const double* sigma1, const double* ell1, const double* X2,
                               const double* sigma2, const double* ell2, const double* Kbar, double* g_sigma1,
                               double* g_ell1, double* g_sigma2, double* g_ell2, long long T1, long long T2, int dx,
                               nmgp_stream_t stream);

/* out[i,j] = Kx[i,j] * Bf[indx1[i], indx2[j]] (+ diag on i == j): logpos.py:87-98 generate_K_index fused with the
 * Hadamard product K_x * K_i and the sigma2_err I of prediction.py:746-750 (irregular observations); Bf is row-major
 * with M columns (square M x M on the reference's call sites), indices int32 */
int nmgp_hadamard_index_cov(const double* Kx, const double* Bf, const int* indx1, const int* indx2, double diag,
                            double* out, long long N1, long long N2, int M, nmgp_stream_t stream);

/* adjoint of the dense indexed log-likelihood -1/2 logdet S - 1/2 y^T S^-1 y, S = A o Bt[i1, i2] + sigma2 I, i.e. what
 * autograd gives the reference for the torch.inverse / torch.logdet likelihoods of the Hadamard and spatially-varying
 * coregionalisation posteriors (logpos.py:350-352, 521-526, 611-616, 690-694).  Sinv = S^-1, alpha = S^-1 y, g = upstream
 * cotangent (device scalar); Abar [N,N] =, Btbar [Mrows,M] += , s2bar[1] += */
int nmgp_dense_loglik_bwd(const double* Sinv, const double* alpha, const double* A, const double* Bt, const int* i1,
                          const int* i2, const double* g, double* Abar, double* Btbar, double* s2bar, long long N,
                          int Mrows, int M, nmgp_stream_t stream);

int nmgp_pairwise_dist(const double* X1, const double* X2, double* out, long long T1, long long T2, int dx,
                       nmgp_stream_t stream);                               /* kernels.py:5-21 */

/* C = alpha A B^T + beta C on the FP64 tensor cores (A [M,K], B [N,K], row-major)
 *                                                    kronecker_operation.py:72-85 kron_mv (torch.mm x2), trailing updates */
int nmgp_gemm_nt(const double* A, const double* B, double* C, long long M, long long N, long long K, long long lda,
                 long long ldb, long long ldc, double alpha, double beta, nmgp_stream_t stream);
/* in-place blocked lower Cholesky, hld = sum log diag L      replaces torch.symeig / torch.logdet / torch.inverse at
 *                                                    kronecker_operation.py:45-47,66-67; distributions.py:37-40,109-110
 * *info = 1 + index of the FIRST non-positive pivot (0 if none); strict upper triangle zeroed on return */
int nmgp_potrf_big(double* A, long long T, long long lda, double* hld, int* info, nmgp_stream_t stream);
/* same with an explicit scratch slot (0..3): independent factorisations running concurrently on different streams
 * (the eigen-blocks of a Kronecker log-density) must use different slots; scratch is kept per (device, slot) */
int nmgp_potrf_big_slot(double* A, long long T, long long lda, double* hld, int* info, int slot,
                        int panel /* panel width, multiple of 128; 0 = chosen from T */, nmgp_stream_t stream);
/* on != 0 while the caller keeps several factorisations / GEMMs in flight on different streams: operand tiles are then
 * staged with cp.async instead of TMA tensor copies (reproducibility finding in profiles/README.md) */
void nmgp_gemm_concurrent_mode(int on);
/* out = scale * inv(L) for a lower-triangular nb x nb block (nb <= 128): building block of the blocked triangular
 * inverse behind the Cholesky-based log-density adjoint (autograd of distributions.py:26-52) */
int nmgp_tri_inv_block(const double* L, long long lda, int nb, double* out, long long ldo, double scale,
                       nmgp_stream_t stream);
/* augmented eigen-block system (T+1) x (T+1), leading dimension lda: rows 0..T-1 = alpha K + sigma2 I, row T =
 * (r^T, 1 + |r|^2 / sigma2).  Its Cholesky factor holds (L^-1 r)^T in the last row, so r^T A^-1 r needs no triangular
 * solve launches: nmgp_augmented_results reads quad = |L^-1 r|^2 and hld = 1/2 logdet A from the factor. */
int nmgp_build_augmented(const double* K, const double* r, double* A, long long T, long long lda,
                         const double* alpha_dev, const double* sigma2_dev, const double* rnorm2_dev,
                         nmgp_stream_t stream);
int nmgp_augmented_results(const double* A, long long T, long long lda, const double* hld_aug, double* hld,
                           double* quad, nmgp_stream_t stream);
/* A = alpha_dev[0] K + sigma2_dev[0] I, scalars read on the device */
int nmgp_scale_add_diag_dev(const double* K, double* A, long long T, const double* alpha_dev, const double* sigma2_dev,
                            nmgp_stream_t stream);
/* x <- (L L^T)^-1 x */
int nmgp_potrs_vec(const double* L, long long T, long long lda, double* x, nmgp_stream_t stream);
/* A = alpha K + sigma2 I (the eigen-block sigma2 I + lambda_m K of sigma2 I + B (x) K) */
int nmgp_scale_add_diag(const double* K, double* A, long long T, double alpha, double sigma2, nmgp_stream_t stream);
/* dense Kronecker product                                                  kronecker_operation.py:5-33 */
int nmgp_kron_product(const double* t1, const double* t2, double* out, int h1, int w1, long long h2, long long w2,
                      nmgp_stream_t stream);
/* Jacobi eigen-decomposition of a small symmetric matrix (upper triangle read), ascending eigenvalues
 *                                                                          torch.symeig(B) at kronecker_operation.py:45 */
int nmgp_eigh_small(const double* A, double* w, double* V, double* work, int n, nmgp_stream_t stream);
long long nmgp_eigh_small_work(int n);   /* doubles of scratch nmgp_eigh_small needs (rotation log of the Jacobi sweeps) */
int nmgp_axpby(const double* x, const double* y, double* out, long long n, double a, double b, nmgp_stream_t stream);
/* out = (a_scale * a_dev[0]) x + b y, the scalar a_dev read on the device */
int nmgp_axpby_dev(const double* x, const double* y, double* out, long long n, const double* a_dev, double a_scale,
                   double b, nmgp_stream_t stream);
int nmgp_dot(const double* x, const double* y, double* out /* += */, long long n, nmgp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NMGP_B200_H */

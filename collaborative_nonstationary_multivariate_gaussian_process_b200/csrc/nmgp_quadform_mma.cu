// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) formulation of the FLOP-dominant kernels.
//
//  k_latent_fused : per (sample, 128-row tile), for every latent j <= I[n]
//        V = P Sigma_W[j]           (DMMA, A = P tile held in registers for the whole j loop,
//                                    B = Sigma_W[j] streamed through a cp.async double buffer)
//        q = rowdot(V, P)           -> s2_g, expected log-likelihood, all row cotangents
//        Pbar += 2 qbar V + mbar mu -> the adjoint w.r.t. P reuses V (no second GEMM)
//     = quadform_fwd + lik_rows + quadform_bwd of nmgp_quadform.cu / nmgp_rows.cu in one pass
//     (reference: code/utils.py:143-144 + code/nmgp_dsvi.py:255-258 and their autograd).
//  k_gram_mma     : SigBar[idx] += P^T diag(qbar[:,j]) P, MuBar[idx] += P^T mbar[:,j] over the rows of one output,
//                   NG latents per CTA, lower-triangular 8x8 blocks only (symmetric result).
//
// Register-resident design: valid for Q <= 64 (NB = ceil(Q/8) <= 8); larger Q uses the shared-memory FMA kernels.
#include "common.cuh"
#include <stdlib.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
// copy a Q x Q row-major matrix into a padded shared tile (row stride ld): one warp per row, no integer division,
// 16-byte chunks when the rows are 16-byte aligned (Q even)
__device__ __forceinline__ void stage_matrix(double* __restrict__ dst, const double* __restrict__ src, int Q, int ld,
                                             int warp, int nwarps, int lane) {
    if ((Q & 1) == 0) {
        for (int a = warp; a < Q; a += nwarps)
            for (int c = 2 * lane; c < Q; c += 64) cp_async16(&dst[a * ld + c], &src[(size_t)a * Q + c]);
    } else {
        for (int a = warp; a < Q; a += nwarps)
            for (int c = lane; c < Q; c += 32) cp_async8(&dst[a * ld + c], &src[(size_t)a * Q + c]);
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__host__ __device__ constexpr int pad4mod8(int n) { return ((n + 3) / 8) * 8 + 4; }   // smallest m >= n, m % 8 == 4
// smallest m >= n with m % 16 == 8: rows of the P tile are then read conflict-free as 16-byte (2-column) accesses by
// the C-fragment layout of the epilogue (8 rows x 4 column pairs per warp)
__host__ __device__ constexpr int pad8mod16(int n) { return ((n + 7) / 16) * 16 + 8; }


// mbarrier + 1-D bulk copy (TMA engine): one elected thread moves a whole padded (Sigma_W[j], mu_W[j]) record
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done = 0;
    const unsigned a = smem_u32(b);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
// rec[j] = [ Sigma_W[j] padded to KP x LDS | mu_W[j] padded to NP ], zeros in the padding; block D appends the padded
// mean block  mu[DP][LDM]  (DP = D rounded up to 8) that the fused kernel brings in by one bulk copy for its two mean-path
// products (phase 1 and the mbar mu_W part of Pbar)
__global__ void k_pad_records(const double* __restrict__ SigW, const double* __restrict__ muW, double* __restrict__ rec,
                              int Q, int KP, int LDS, int NP, int D, int DP, int LDM) {
    const int j = blockIdx.x, REC = KP * LDS + NP;
    if (j == D) {
        double* out = rec + (size_t)D * REC;
        for (int e = threadIdx.x; e < DP * LDM; e += blockDim.x) {
            const int jj = e / LDM, c = e - jj * LDM;
            out[e] = (jj < D && c < Q) ? muW[(size_t)jj * Q + c] : 0.0;
        }
        return;
    }
    double* out = rec + (size_t)j * REC;
    for (int e = threadIdx.x; e < KP * LDS; e += blockDim.x) {
        const int a = e / LDS, b = e - a * LDS;
        out[e] = (a < Q && b < Q) ? SigW[((size_t)j * Q + a) * Q + b] : 0.0;
    }
    for (int c = threadIdx.x; c < NP; c += blockDim.x) out[KP * LDS + c] = c < Q ? muW[(size_t)j * Q + c] : 0.0;
}

// ------------------------------------------------------------------------------------------------------------
// k_latent_fused: 64-row tiles, 4 warps, TWO CTAs per SM.  The two CTAs of an SM share no barrier, so the epilogue /
// prologue / output phases of one overlap the DMMA phase of the other (68.1 -> 64.9 ms per step against one 128-row,
// 8-warp CTA per SM).  This only pays because a latent's record arrives by one bulk copy: with per-thread cp.async
// staging every CTA paid the staging instructions itself and the same split measured 18% slower (profiles/README.md).
// k_coef_quadform_mma keeps 128-row tiles (LF_ROWS / LF_THREADS).
#define LF_ROWS 128
#define LF_THREADS 256
#define LFK_ROWS 64       // k_latent_fused: 64-row tiles, 4 warps, two CTAs per SM
#define LFK_THREADS 128
#define LFK_CTAS 2

template <int NB, int KS>
struct LFShape {
    static constexpr int NP = 8 * NB;                      // padded N (columns of Sigma / P in C layout)
    static constexpr int KP = 4 * KS;                      // padded K
    static constexpr int LDP = pad8mod16(NP > KP ? NP : KP);
    static constexpr int LDS = pad4mod8(NP);
    static constexpr int REC = KP * LDS + NP;              // one latent's padded record: Sigma_W[j] then mu_W[j]
    // padded mean block mu[DP][LDM] (row = latent): staged in a record buffer when it fits (D <= MU_DMAX)
    static constexpr int LDM = pad4mod8(NP > KP ? NP : KP);
    static constexpr int MU_DMAX = 64;
    static constexpr int MUB = MU_DMAX * LDM;
    static constexpr int BUF = REC > MUB ? REC : MUB;      // doubles per staging buffer
    static constexpr size_t smem_doubles = (size_t)LFK_ROWS * LDP + 2 * (size_t)BUF + 2 * LFK_ROWS + 6;
    static constexpr size_t smem_bytes = smem_doubles * 8 + LFK_ROWS * 4;
};

template <int NB, int KS>
__global__ void __launch_bounds__(LFK_THREADS, LFK_CTAS)
k_latent_fused(const double* __restrict__ PG, const double* __restrict__ cG, const double* __restrict__ l,
               const double* __restrict__ y, const int* __restrict__ I, const double* __restrict__ rec,
               const double* __restrict__ muW, const double* __restrict__ hyp, double scale,
               double* __restrict__ Rsum, double* __restrict__ ghyp, double* __restrict__ lbar,
               double* __restrict__ mgbar, double* __restrict__ qgbar, double* __restrict__ cGbar,
               double* __restrict__ PGbar, long long B, int Q, int D, long long ystride) {
    using SH = LFShape<NB, KS>;
    constexpr int LDP = SH::LDP, LDS = SH::LDS, REC = SH::REC, LDM = SH::LDM, BUF = SH::BUF;
    extern __shared__ __align__(16) double sm[];
    double* Ps = sm;                                   // [LFK_ROWS][LDP]
    double* Ss = Ps + (size_t)LFK_ROWS * LDP;           // [2][BUF]: double-buffered records of the latent j / mean block
    double* rrs = Ss + 2 * (size_t)BUF;                // [LFK_ROWS]
    double* omcs = rrs + LFK_ROWS;                      // [LFK_ROWS]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(omcs + LFK_ROWS);   // full[2], prologue, pad, empty[2]
    unsigned long long* mempty = mbar + 4;
    int* Is = reinterpret_cast<int*>(mbar + 6);        // [LFK_ROWS]

    const int s = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * LFK_ROWS;   // (longest-tiles-first order measured: no gain)
    const int nrows = (int)min((long long)LFK_ROWS, B - row0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const size_t rbase = (size_t)s * B + row0;          // first row of this tile in [ns*B]
    const double s2e = hyp[H_S2_ERR];
    const int DP = (D + 7) & ~7;
    const bool mu_smem = DP <= SH::MU_DMAX && DP * LDM <= BUF;   // the padded mean block fits a staging buffer
    const double* mublk = rec + (size_t)D * REC;
    const bool bulk_rows = (Q & 1) == 0;                // 16-byte aligned rows of P: one bulk copy per row

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar[2], 1);
        mbar_init(&mempty[0], LFK_THREADS / 32);
        mbar_init(&mempty[1], LFK_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // everything the prologue needs is put in flight at once on the bulk-copy engine: the P rows of the tile (mbar[2]),
    // the mean block into buffer 1 (mbar[2]) and the record of latent 0 into buffer 0 (mbar[0])
    if (tid == 0) {
        const unsigned bytes = (bulk_rows ? (unsigned)nrows * (unsigned)Q * 8u : 0u) + (mu_smem ? (unsigned)(DP * LDM) * 8u : 0u);
        mbar_expect_tx(&mbar[2], bytes);
        if (mu_smem) bulk_g2s(Ss + BUF, mublk, (unsigned)(DP * LDM) * 8u, &mbar[2]);
        mbar_expect_tx(&mbar[0], (unsigned)(REC * sizeof(double)));
        bulk_g2s(Ss, rec, (unsigned)(REC * sizeof(double)), &mbar[0]);
    }
    if (bulk_rows) {
        if (tid < nrows) bulk_g2s(Ps + (size_t)tid * LDP, PG + (rbase + tid) * Q, (unsigned)Q * 8u, &mbar[2]);
        for (int e = tid; e < LFK_ROWS * (LDP - Q); e += LFK_THREADS) {       // padding columns
            const int r = e / (LDP - Q), a = Q + (e - r * (LDP - Q));
            Ps[r * LDP + a] = 0.0;
        }
        for (int e = tid; e < (LFK_ROWS - nrows) * Q; e += LFK_THREADS) {     // rows past the end of the batch
            const int r = nrows + e / Q, a = e - (e / Q) * Q;
            Ps[r * LDP + a] = 0.0;
        }
    } else {
        for (int e = tid; e < LFK_ROWS * LDP; e += LFK_THREADS) {
            int r = e / LDP, a = e - r * LDP;
            Ps[e] = (r < nrows && a < Q) ? PG[(rbase + r) * Q + a] : 0.0;
        }
    }
    for (int r = tid; r < LFK_ROWS; r += LFK_THREADS) {
        Is[r] = r < nrows ? I[row0 + r] : -1;
        omcs[r] = r < nrows ? 1.0 - cG[rbase + r] : 0.0;
    }
    mbar_wait(&mbar[2], 0u);
    __syncthreads();

    // A fragments of this warp's 16 rows: rows 16w + 8mb + g, k = 4ks + t
    double afr[2][KS];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) afr[mb][ks] = Ps[(16 * w + 8 * mb + g) * LDP + 4 * ks + t];
    const int rloc[2] = {16 * w + g, 16 * w + 8 + g};
    const int myI[2] = {Is[rloc[0]], Is[rloc[1]]};

    // ---- phase 1: m[n,j] = p_n . mu_W[j], F_n, residual -------------------------------------------------
    double Fp[2] = {0.0, 0.0};
    for (int jb = 0; jb * 8 < D; ++jb) {
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        const int jn = 8 * jb + g;                       // B-fragment column (latent) of this lane
        double lv[2][2];                                 // coefficients of this lane's (row, latent) results, loaded ahead
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * jb + 2 * t + e;
                lv[mb][e] = (rloc[mb] < nrows && j <= myI[mb]) ? __ldg(&l[(rbase + rloc[mb]) * D + j]) : 0.0;
            }
        if (mu_smem) {
            const double* Mu = Ss + BUF;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double b = Mu[jn * LDM + 4 * ks + t];
                dmma884(acc[0][0], acc[0][1], afr[0][ks], b);
                dmma884(acc[1][0], acc[1][1], afr[1][ks], b);
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int k = 4 * ks + t;
                double b = (jn < D && k < Q) ? __ldg(&muW[(size_t)jn * Q + k]) : 0.0;
                dmma884(acc[0][0], acc[0][1], afr[0][ks], b);
                dmma884(acc[1][0], acc[1][1], afr[1][ks], b);
            }
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            if (rloc[mb] < nrows) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * jb + 2 * t + e;
                    if (j <= myI[mb]) {
                        Fp[mb] = fma(lv[mb][e], acc[mb][e], Fp[mb]);
                        mgbar[(rbase + rloc[mb]) * D + j] = acc[mb][e];   // m[n,j] parked in the slot its cotangent overwrites at latent j
                    }
                }
            }
        }
    }
    double racc = 0.0, gacc = 0.0;                       // per-thread partial sums of the sample statistics
    double rr[2];
    const double cst = -0.5 * log(s2e) - log(sqrt(2.0 * 3.14159265358979323846));
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        double F = Fp[mb];
        F += __shfl_xor_sync(0xffffffffu, F, 1);
        F += __shfl_xor_sync(0xffffffffu, F, 2);
        double r = 0.0;
        if (rloc[mb] < nrows) r = y[(size_t)s * ystride + row0 + rloc[mb]] - F;
        rr[mb] = r / s2e;
        if (t == 0 && rloc[mb] < nrows) {
            rrs[rloc[mb]] = rr[mb];
            racc += -(r * r) / (2.0 * s2e) + cst;
            gacc += (r * r) / (2.0 * s2e) - 0.5;
        }
    }
    // entries of latents a row does not use (j > I[n]) are exact zeros in every [ns,B,D] output
    {
        const int r = tid >> 1;                            // two threads per row
        if (r < nrows)
            for (int j = Is[r] + 1 + (tid & 1); j < D; j += 2) {
                const size_t o = (rbase + r) * D + j;
                lbar[o] = 0.0;
                mgbar[o] = 0.0;
                qgbar[o] = 0.0;
            }
    }
    // ---- phase 2: quadratic forms, j loop with a cp.async double buffer -----------------------------------
    const int jmax = Is[nrows - 1];
    const int warpmaxI = (16 * w < nrows) ? Is[min(16 * w + 15, nrows - 1)] : -1;
    auto stage = [&](int j, int buf) {                    // one thread: whole record by the bulk-copy engine
        mbar_expect_tx(&mbar[buf], (unsigned)(REC * sizeof(double)));
        bulk_g2s(Ss + (size_t)buf * BUF, rec + (size_t)j * REC, (unsigned)(REC * sizeof(double)), &mbar[buf]);
    };
    auto stage_mu = [&](int buf) {                        // the mean block as "record jmax + 1" (for the product after the loop)
        mbar_expect_tx(&mbar[buf], (unsigned)(DP * LDM) * 8u);
        bulk_g2s(Ss + (size_t)buf * BUF, mublk, (unsigned)(DP * LDM) * 8u, &mbar[buf]);
    };
    double pacc[2][NB][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) pacc[mb][nb][0] = pacc[mb][nb][1] = 0.0;
    double pen[2] = {0.0, 0.0}, gsum[2] = {0.0, 0.0};

    __syncthreads();                                       // phase 1 is done with the mean block in buffer 1 (refilled at j = 0)
    double lnext[2], mnext[2];                             // (record 0 is already in flight since the prologue)
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        lnext[mb] = (rloc[mb] < nrows) ? __ldg(&l[(rbase + rloc[mb]) * D]) : 0.0;
        mnext[mb] = (rloc[mb] < nrows && 0 <= myI[mb]) ? __ldcg(&mgbar[(rbase + rloc[mb]) * D]) : 0.0;
    }
    for (int j = 0; j <= jmax; ++j) {
        const int buf = j & 1;
        // No CTA barrier in the loop: a warp releases a record buffer (empty mbarrier, one arrival per warp) as soon as its
        // DMMA run over it is done, and warp 0 refills the other buffer once all four warps have released it -- the
        // epilogue of a slow warp no longer holds the others back, and the refill starts one epilogue earlier.
        if (w == 0) {
            if (j >= 1) mbar_wait(&mempty[buf ^ 1], (unsigned)(((j - 1) >> 1) & 1));    // latent j - 1 is done everywhere
            if (lane == 0) {
                if (j + 1 <= jmax) stage(j + 1, buf ^ 1);
                else if (mu_smem) stage_mu(buf ^ 1);
            }
            __syncwarp();
        }
        if (warpmaxI < j) {                                // warp-uniform: none of this warp's rows uses latent j
            // a skipping warp reads no record, so nothing paces it: it must not arrive for latent j while the buffer's
            // previous phase (latent j - 2) is still open, or that phase would complete with two of its arrivals
            if (j >= 2) mbar_wait(&mempty[buf], (unsigned)(((j - 2) >> 1) & 1));
            if (lane == 0) mbar_arrive(&mempty[buf]);
            continue;
        }
        const double lcur[2] = {lnext[0], lnext[1]}, mcur[2] = {mnext[0], mnext[1]};
        if (j + 1 < D) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                lnext[mb] = (rloc[mb] < nrows) ? __ldg(&l[(rbase + rloc[mb]) * D + j + 1]) : 0.0;
                mnext[mb] = (rloc[mb] < nrows && j + 1 <= myI[mb]) ? __ldcg(&mgbar[(rbase + rloc[mb]) * D + j + 1]) : 0.0;
            }
        }
        mbar_wait(&mbar[buf], (unsigned)((j >> 1) & 1));  // record j landed
        const double* Sd = Ss + (size_t)buf * BUF;
        double V[2][NB][2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) V[mb][nb][0] = V[mb][nb][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const double b = Sd[(4 * ks + t) * LDS + 8 * nb + g];
                dmma884(V[0][nb][0], V[0][nb][1], afr[0][ks], b);
                dmma884(V[1][nb][0], V[1][nb][1], afr[1][ks], b);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&mempty[buf]);          // this warp is done reading record j
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int rl = rloc[mb];
            double qp = 0.0, qp1 = 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const double2 pv = *reinterpret_cast<const double2*>(&Ps[rl * LDP + 8 * nb + 2 * t]);
                qp = fma(V[mb][nb][0], pv.x, qp);
                qp1 = fma(V[mb][nb][1], pv.y, qp1);
            }
            qp += qp1;
            qp += __shfl_xor_sync(0xffffffffu, qp, 1);
            qp += __shfl_xor_sync(0xffffffffu, qp, 2);
            const double mp = mcur[mb];                                // p_n . mu_W[j], computed by phase 1
            const bool live = (rl < nrows) && (j <= myI[mb]);
            const double lj = live ? lcur[mb] : 0.0;
            const double gq = scale * (0.5 / s2e) * lj * lj;          // cotangent of s2_g[n,j]
            const double gm = -scale * rr[mb] * lj;                    // cotangent of mu_g[n,j]
            const double s2g = omcs[rl] + qp;
            if (live && t == 0) {
                const size_t o = (rbase + rl) * D + j;
                pen[mb] = fma(lj * lj, s2g, pen[mb]);
                gsum[mb] += gq;
                lbar[o] = -scale * (rr[mb] * mp - (1.0 / s2e) * lj * s2g);
                qgbar[o] = gq;
                mgbar[o] = gm;
            }
            const double g2 = 2.0 * gq;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                pacc[mb][nb][0] = fma(g2, V[mb][nb][0], pacc[mb][nb][0]);
                pacc[mb][nb][1] = fma(g2, V[mb][nb][1], pacc[mb][nb][1]);
            }
        }
    }
    // Pbar += mbar mu_W (the mean-path part of the adjoint) as one DMMA product over the latents: A = mbar rows of this
    // warp (just written to mgbar; zeros beyond I[n]), B = mu_W.  Replaces 28 FMAs per thread and latent in the loop.
    __syncthreads();
    if (mu_smem) {
        // B = the mean block staged as record jmax + 1; A = this warp's mbar rows, all loads issued before the first DMMA
        const int bm = (jmax + 1) & 1;
        const double* Mu = Ss + (size_t)bm * BUF;
        constexpr int KSM = SH::MU_DMAX / 4;
        double av[2][KSM];
#pragma unroll
        for (int ks = 0; ks < KSM; ++ks)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const int jj = 4 * ks + t;
                av[mb][ks] = (rloc[mb] < nrows && jj < D) ? __ldcg(&mgbar[(rbase + rloc[mb]) * D + jj]) : 0.0;
            }
        mbar_wait(&mbar[bm], (unsigned)(((jmax + 1) >> 1) & 1));
#pragma unroll
        for (int ks = 0; ks < KSM; ++ks) {
            if (4 * ks < D) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const double b = Mu[(4 * ks + t) * LDM + 8 * nb + g];
                    dmma884(pacc[0][nb][0], pacc[0][nb][1], av[0][ks], b);
                    dmma884(pacc[1][nb][0], pacc[1][nb][1], av[1][ks], b);
                }
            }
        }
    } else {
        for (int ks = 0; 4 * ks < D; ++ks) {
            const int jj = 4 * ks + t;
            double av[2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
                av[mb] = (rloc[mb] < nrows && jj < D) ? __ldcg(&mgbar[(rbase + rloc[mb]) * D + jj]) : 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const int c = 8 * nb + g;
                const double b = (jj < D && c < Q) ? __ldg(&muW[(size_t)jj * Q + c]) : 0.0;
                dmma884(pacc[0][nb][0], pacc[0][nb][1], av[0], b);
                dmma884(pacc[1][nb][0], pacc[1][nb][1], av[1], b);
            }
        }
    }

    // ---- outputs ---------------------------------------------------------------------------------------------
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const int rl = rloc[mb];
        if (rl < nrows) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = 8 * nb + 2 * t + e;
                    if (c < Q) PGbar[(rbase + rl) * Q + c] = pacc[mb][nb][e];
                }
            }
            if (t == 0) {
                cGbar[rbase + rl] = -gsum[mb];
                racc -= (0.5 / s2e) * pen[mb];
                gacc += (0.5 / s2e) * pen[mb];
            }
        }
    }
    racc = block_sum(racc);
    gacc = block_sum(gacc);
    if (tid == 0) {
        atomicAdd(&Rsum[s], racc);
        atomicAdd(&ghyp[H_S2_ERR], -scale * gacc);
    }
}

template <int NB, int KS>
static int launch_latent_fused(const double* PG, const double* cG, const double* l, const double* y, const int* I,
                               const double* SigW, const double* muW, const double* hyp, double scale, double* Rsum,
                               double* ghyp, double* lbar, double* mgbar, double* qgbar, double* cGbar, double* PGbar,
                               int ns, long long B, int Q, int D, long long ystride, cudaStream_t st) {
    using SH = LFShape<NB, KS>;
    size_t smem = SH::smem_bytes;
    if (int r = nmgp_opt_in_smem(k_latent_fused<NB, KS>, smem, "nmgp_latent_fused")) return r;
    // padded records of (Sigma_W[j], mu_W[j]) in a library-owned scratch buffer (grown on demand, one per device and
    // kernel shape; a process drives one GPU in the intended one-rank-per-GPU use, but a second device must not be handed
    // the first one's pointer)
    constexpr int MAXDEV = 16;
    static double* recs[MAXDEV] = {nullptr};
    static size_t rec_caps[MAXDEV] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) {
        nmgp_set_error("nmgp_latent_fused: unsupported device ordinal %d", dev);
        return -4;
    }
    double*& rec = recs[dev];
    size_t& rec_cap = rec_caps[dev];
    const int DP = (D + 7) & ~7;
    const size_t need = (size_t)D * SH::REC + (size_t)DP * SH::LDM;
    if (need > rec_cap) {
        if (rec) {
            cudaDeviceSynchronize();
            cudaFree(rec);
        }
        if (cudaMalloc(&rec, need * sizeof(double)) != cudaSuccess) {
            rec = nullptr;
            rec_cap = 0;
            nmgp_set_error("nmgp_latent_fused: cannot allocate %zu bytes of scratch", need * sizeof(double));
            return -4;
        }
        rec_cap = need;
    }
    k_pad_records<<<NMGP_L(D + 1), 256, 0, st>>>(SigW, muW, rec, Q, SH::KP, SH::LDS, SH::NP, D, DP, SH::LDM);
    dim3 grid((unsigned)((B + LFK_ROWS - 1) / LFK_ROWS), ns);
    k_latent_fused<NB, KS><<<NMGP_L(grid), LFK_THREADS, smem, st>>>(PG, cG, l, y, I, rec, muW, hyp, scale, Rsum, ghyp, lbar,
                                                           mgbar, qgbar, cGbar, PGbar, B, Q, D, ystride);
    return nmgp_launch_status("nmgp_latent_fused");
}

// generic-Q fallbacks (nmgp_quadform.cu / nmgp_rows.cu)
extern "C" int nmgp_quadform_fwd(const double*, const double*, const int*, const int*, const double*, const double*,
                                 double*, double*, int, long long, int, int, int, cudaStream_t);
extern "C" int nmgp_quadform_bwd(const double*, const double*, const int*, const int*, const double*, const double*,
                                 const double*, const double*, double*, double*, int, long long, int, int, int,
                                 cudaStream_t);
extern "C" int nmgp_lik_rows(const double*, const double*, const double*, const double*, const double*, const int*,
                             const double*, double, double*, double*, double*, double*, double*, double*, int,
                             long long, int, long long, cudaStream_t);

#define LF_CASE(nb, ks)                                                                                              \
    if (NBr == nb && KSr == ks)                                                                                      \
        return launch_latent_fused<nb, ks>(PG, cG, l, y, I, SigW, muW, hyp, scale, Rsum, ghyp, lbar, mgbar, qgbar,   \
                                           cGbar, PGbar, ns, B, Q, D, ystride, st);

// Fused latent-function statistics + expected log-likelihood + cotangents.  `work_q`, `work_m` ([ns,B,D] each) are
// only used on the generic-Q path (they receive q and m).
NMGP_API int nmgp_latent_fused(const double* PG, const double* cG, const double* l, const double* y, const int* I,
                               const int* seg, const double* SigW, const double* muW, const double* hyp, double scale,
                               double* Rsum, double* ghyp, double* lbar, double* mgbar, double* qgbar, double* cGbar,
                               double* PGbar, double* work_q, double* work_m, int ns, long long B, int Q, int D,
                               long long ystride, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128 && D > 0, "nmgp_latent_fused");
    if (ns == 0 || B == 0) return 0;
    const int NBr = (Q + 7) / 8, KSr = (Q + 3) / 4;
    if (NBr <= 8) {
        LF_CASE(1, 1) LF_CASE(1, 2) LF_CASE(2, 3) LF_CASE(2, 4) LF_CASE(3, 5) LF_CASE(3, 6) LF_CASE(4, 7) LF_CASE(4, 8)
        LF_CASE(5, 9) LF_CASE(5, 10) LF_CASE(6, 11) LF_CASE(6, 12) LF_CASE(7, 13) LF_CASE(7, 14) LF_CASE(8, 15)
        LF_CASE(8, 16)
    }
    // Q > 64: three-kernel path
    if (int r = nmgp_quadform_fwd(PG, PG, I, seg, SigW, muW, work_q, work_m, ns, B, Q, D, MODE_W, st)) return r;
    if (int r = nmgp_lik_rows(l, work_m, work_q, cG, y, I, hyp, scale, Rsum, ghyp, lbar, mgbar, qgbar, cGbar, ns, B, D,
                              ystride, st))
        return r;
    return nmgp_quadform_bwd(PG, PG, I, seg, SigW, muW, qgbar, mgbar, PGbar, PGbar, ns, B, Q, D, MODE_W, st);
}

// ------------------------------------------------------------------------------------------------------------
// Coefficient-side quadratic forms on DMMA (MODE_U): for the rows of output i and every j <= i
//     q[n,j] = p^T Sigma_U[pair(i,j)] p,  m[n,j] = p . mu_U[pair(i,j)],   p = P_L1[n] if j == i else P_L0[n]
// (forward, code/utils.py:120-122 inside the D(D+1)/2 loop of code/nmgp_dsvi.py:228-237, only for the pairs the row
// consumes) and the adjoint  Pbar += 2 qbar Sigma p + mbar mu  (backward).  Same tiling as k_latent_fused: 128-row
// tiles, A = P tile in registers (reloaded when the source system changes), B = Sigma_U[pair] through a cp.async
// double buffer; the tile walks the (i, j) pairs of every output present in it.
template <int NB, int KS, bool BWD>
__global__ void __launch_bounds__(LF_THREADS, 1)
k_coef_quadform_mma(const double* __restrict__ Pa, const double* __restrict__ Pb, const int* __restrict__ I,
                    const double* __restrict__ Sig, const double* __restrict__ Mu, double* __restrict__ qout,
                    double* __restrict__ mout, const double* __restrict__ qbar, const double* __restrict__ mbar,
                    double* __restrict__ Pabar, double* __restrict__ Pbbar, long long B, int Q, int D, FastDiv qdiv) {
    using SH = LFShape<NB, KS>;
    constexpr int LDP = SH::LDP, LDS = SH::LDS, NP = SH::NP, KP = SH::KP;
    extern __shared__ __align__(16) double sm[];
    double* PaS = sm;                                   // [LF_ROWS][LDP]
    double* PbS = PaS + (size_t)LF_ROWS * LDP;          // [LF_ROWS][LDP]
    double* Ss = PbS + (size_t)LF_ROWS * LDP;           // [2][KP][LDS]
    double* mus = Ss + 2 * (size_t)KP * LDS;            // [2][NP]
    double* qcol = mus + 2 * NP;                        // [2][LF_ROWS]  (BWD: qbar column of the task)
    double* mcol = qcol + 2 * LF_ROWS;                  // [2][LF_ROWS]
    int* Is = reinterpret_cast<int*>(mcol + 2 * LF_ROWS);

    const int s = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * LF_ROWS;
    const int nrows = (int)min((long long)LF_ROWS, B - row0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const size_t rbase = (size_t)s * B + row0;

    for (int e = tid; e < LF_ROWS * LDP; e += LF_THREADS) {
        int r = e / LDP, a = e - r * LDP;
        bool ok = r < nrows && a < Q;
        PaS[e] = ok ? Pa[(rbase + r) * Q + a] : 0.0;
        PbS[e] = ok ? Pb[(rbase + r) * Q + a] : 0.0;
    }
    for (int e = tid; e < 2 * KP * LDS + 2 * NP; e += LF_THREADS) Ss[e] = 0.0;
    for (int r = tid; r < LF_ROWS; r += LF_THREADS) Is[r] = r < nrows ? I[row0 + r] : -1;
    __syncthreads();

    const int rloc[2] = {16 * w + g, 16 * w + 8 + g};
    const int myI[2] = {Is[rloc[0]], Is[rloc[1]]};
    const int wlo = (16 * w < nrows) ? Is[16 * w] : 1 << 30;                         // outputs spanned by this warp
    const int whi = (16 * w < nrows) ? Is[min(16 * w + 15, nrows - 1)] : -1;
    const int i_lo = Is[0], i_hi = Is[nrows - 1];

    auto stage = [&](int i, int j, int buf) {
        const int idx = pair_slot(i, j, D);
        const double* Sg = Sig + (size_t)idx * Q * Q;
        double* Sd = Ss + (size_t)buf * KP * LDS;
        for (int e = tid; e < Q * Q; e += LF_THREADS) {
            int a = (int)qdiv.div((unsigned)e), b = e - a * Q;
            cp_async8(&Sd[a * LDS + b], &Sg[e]);
        }
        for (int c = tid; c < Q; c += LF_THREADS) cp_async8(&mus[buf * NP + c], &Mu[(size_t)idx * Q + c]);
        if (BWD) {
            for (int r = tid; r < nrows; r += LF_THREADS) {
                cp_async8(&qcol[buf * LF_ROWS + r], &qbar[(rbase + r) * D + j]);
                cp_async8(&mcol[buf * LF_ROWS + r], &mbar[(rbase + r) * D + j]);
            }
        }
    };

    double afr[2][KS];
    double pacc[2][NB][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) pacc[mb][nb][0] = pacc[mb][nb][1] = 0.0;
    int cursrc = -1;                                   // 0: Pa fragments loaded, 1: Pb fragments loaded

    int i = i_lo, j = 0, it = 0;
    stage(i, j, 0);
    cp_async_commit();
    while (i <= i_hi) {
        const int buf = it & 1;
        int ni = i, nj = j + 1;
        if (nj > i) { ni = i + 1; nj = 0; }
        cp_async_wait<0>();
        __syncthreads();
        if (ni <= i_hi) stage(ni, nj, buf ^ 1);
        cp_async_commit();
        if (i >= wlo && i <= whi) {                    // warp-uniform: some row of this warp belongs to output i
            const int src = (j == i) ? 1 : 0;
            const double* Psrc = src ? PbS : PaS;
            if (src != cursrc) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) afr[mb][ks] = Psrc[(16 * w + 8 * mb + g) * LDP + 4 * ks + t];
                cursrc = src;
            }
            const double* Sd = Ss + (size_t)buf * KP * LDS;
            double V[2][NB][2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) V[mb][nb][0] = V[mb][nb][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const double b = Sd[(4 * ks + t) * LDS + 8 * nb + g];
                    dmma884(V[0][nb][0], V[0][nb][1], afr[0][ks], b);
                    dmma884(V[1][nb][0], V[1][nb][1], afr[1][ks], b);
                }
            }
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const int rl = rloc[mb];
                const bool live = (rl < nrows) && (myI[mb] == i);
                if (!BWD) {
                    double qp = 0.0, mp = 0.0;
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        const double p0 = Psrc[rl * LDP + 8 * nb + 2 * t], p1 = Psrc[rl * LDP + 8 * nb + 2 * t + 1];
                        qp = fma(V[mb][nb][0], p0, qp);
                        qp = fma(V[mb][nb][1], p1, qp);
                        mp = fma(mus[buf * NP + 8 * nb + 2 * t], p0, mp);
                        mp = fma(mus[buf * NP + 8 * nb + 2 * t + 1], p1, mp);
                    }
                    qp += __shfl_xor_sync(0xffffffffu, qp, 1);
                    qp += __shfl_xor_sync(0xffffffffu, qp, 2);
                    mp += __shfl_xor_sync(0xffffffffu, mp, 1);
                    mp += __shfl_xor_sync(0xffffffffu, mp, 2);
                    if (live && t == 0) {
                        qout[(rbase + rl) * D + j] = qp;
                        mout[(rbase + rl) * D + j] = mp;
                    }
                } else {
                    const double g2 = live ? 2.0 * qcol[buf * LF_ROWS + rl] : 0.0;
                    const double gm = live ? mcol[buf * LF_ROWS + rl] : 0.0;
                    if (src == 0) {
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            pacc[mb][nb][0] = fma(g2, V[mb][nb][0], fma(gm, mus[buf * NP + 8 * nb + 2 * t], pacc[mb][nb][0]));
                            pacc[mb][nb][1] = fma(g2, V[mb][nb][1], fma(gm, mus[buf * NP + 8 * nb + 2 * t + 1], pacc[mb][nb][1]));
                        }
                    } else if (live) {                 // the single diagonal pair of this row: direct store
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = 8 * nb + 2 * t + e;
                                if (c < Q)
                                    Pbbar[(rbase + rl) * Q + c] = fma(g2, V[mb][nb][e], gm * mus[buf * NP + c]);
                            }
                    }
                }
            }
        }
        i = ni; j = nj; ++it;
    }
    cp_async_wait<0>();
    if (BWD) {
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int rl = rloc[mb];
            if (rl < nrows) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * nb + 2 * t + e;
                        if (c < Q) Pabar[(rbase + rl) * Q + c] = pacc[mb][nb][e];
                    }
            }
        }
    }
}

template <int NB, int KS, bool BWD>
static int launch_coef_quadform(const double* Pa, const double* Pb, const int* I, const double* Sig, const double* Mu,
                                double* q, double* m, const double* qbar, const double* mbar, double* Pabar,
                                double* Pbbar, int ns, long long B, int Q, int D, cudaStream_t st) {
    using SH = LFShape<NB, KS>;
    size_t smem = 8 * (2 * (size_t)LF_ROWS * SH::LDP + 2 * (size_t)SH::KP * SH::LDS + 2 * SH::NP + 4 * LF_ROWS) + LF_ROWS * 4;
    if (int r = nmgp_opt_in_smem(k_coef_quadform_mma<NB, KS, BWD>, smem, "nmgp_quadform(mma)")) return r;
    dim3 grid((unsigned)((B + LF_ROWS - 1) / LF_ROWS), ns);
    k_coef_quadform_mma<NB, KS, BWD><<<NMGP_L(grid), LF_THREADS, smem, st>>>(Pa, Pb, I, Sig, Mu, q, m, qbar, mbar, Pabar, Pbbar,
                                                                     B, Q, D, FastDiv((unsigned)Q));
    return nmgp_launch_status("nmgp_quadform(mma)");
}

#define CQ_CASE(nb, ks)                                                                                               \
    if (NBr == nb && KSr == ks)                                                                                       \
        return bwd ? launch_coef_quadform<nb, ks, true>(Pa, Pb, I, Sig, Mu, q, m, qbar, mbar, Pabar, Pbbar, ns, B, Q, \
                                                        D, st)                                                        \
                   : launch_coef_quadform<nb, ks, false>(Pa, Pb, I, Sig, Mu, q, m, qbar, mbar, Pabar, Pbbar, ns, B,   \
                                                         Q, D, st);

// MODE_U forward (bwd = false: q, m) / backward (bwd = true: Pabar, Pbbar).  Returns 1 when Q > 64 (caller falls back).
int nmgp_coef_quadform_mma(bool bwd, const double* Pa, const double* Pb, const int* I, const double* Sig,
                           const double* Mu, double* q, double* m, const double* qbar, const double* mbar,
                           double* Pabar, double* Pbbar, int ns, long long B, int Q, int D, cudaStream_t st) {
    const int NBr = (Q + 7) / 8, KSr = (Q + 3) / 4;
    if (NBr > 8) return 1;
    CQ_CASE(1, 1) CQ_CASE(1, 2) CQ_CASE(2, 3) CQ_CASE(2, 4) CQ_CASE(3, 5) CQ_CASE(3, 6) CQ_CASE(4, 7) CQ_CASE(4, 8)
    CQ_CASE(5, 9) CQ_CASE(5, 10) CQ_CASE(6, 11) CQ_CASE(6, 12) CQ_CASE(7, 13) CQ_CASE(7, 14) CQ_CASE(8, 15)
    CQ_CASE(8, 16)
    return 1;
}

// ------------------------------------------------------------------------------------------------------------
// Weighted Gram matrices on DMMA.  grid (D outputs, jgroups [+1 for the MODE_U diagonal pair], ns), 128 threads.
#define GM_TROWS 32     // rows per staged tile (8 k-steps; 64-row tiles measured the same: 4.30 vs 4.32 ms)
#define GM_THREADS 256  // 8 warps.  NB <= 8: warp w & 3 = block-row role, w >> 2 = which half of the 4 latents of the CTA
                        // it accumulates.  8 < NB <= 16 (WIDE): 8 block-row roles, 2 latents per CTA, both in every warp
#define GM_NGW 2        // latents accumulated per warp

template <int NB>
struct GMShape {
    static constexpr bool WIDE = NB > 8;
    static constexpr int NG = WIDE ? 2 : 4;                // latents per CTA
    static constexpr int NP = 8 * NB;
    static constexpr int LDP = pad4mod8(NP);
    // Block-row roles.  Even NB: role w owns block rows (w, NB-1-w): NB+1 blocks each.  Odd NB: rows (w, NB-2-w) for
    // w < (NB-1)/2 and the last row alone: NB blocks each -- e.g. NB = 7 (Q = 50): 7/7/7/7 instead of 8/8/8/4 with the
    // even rule, which left the fourth warp scheduler's tensor pipe idle half of the time (a warp's scheduler is its
    // index mod 4, so a light role is a light scheduler).
    static constexpr int NSLOT = (NB & 1) ? NB : NB + 1;
    static constexpr size_t smem_doubles = 2 * (size_t)GM_TROWS * LDP + 2 * 2 * NG * GM_TROWS;
};

// block-row roles (see GMShape): role w of a CTA with NB block rows
__host__ __device__ constexpr bool gm_active(int NB, int w) { return (NB & 1) ? w <= (NB - 1) / 2 : w <= NB - 1 - w; }
__host__ __device__ constexpr int gm_a1(int NB, int w) { return (NB & 1) ? (w < (NB - 1) / 2 ? w : NB - 1) : w; }
__host__ __device__ constexpr int gm_a2(int NB, int w) { return (NB & 1) ? (w < (NB - 1) / 2 ? NB - 2 - w : NB - 1) : NB - 1 - w; }

// One staged tile (GM_TROWS rows) for a warp whose role (block rows A1 <= A2) is known at compile time.  The A fragment
// of block row a and the B fragment of block column a are the same shared-memory element (Gram matrix: both operands
// are P), so the warp loads the A2 + 1 distinct fragments of a k-step once and uses them on both sides: 4-7 shared
// loads per k-step at NB = 7 instead of 9-11 when the slot -> block mapping was resolved at run time.
template <int NB, int A1, int A2, bool WIDE, int NG, int NSLOT, int LDP>
__device__ __forceinline__ void gram_tile(double (&acc)[GM_NGW][NSLOT][2], double (&accm)[2][2], const double* __restrict__ Pd,
                                          const double* __restrict__ wqb, const double* __restrict__ wmb, int g, int t, int ug) {
    const int u0 = ug * GM_NGW;
#pragma unroll
    for (int kk = 0; kk < GM_TROWS / 4; ++kk) {
        const int n = 4 * kk + t;                            // row of the tile this lane feeds as k index
        double f[A2 + 1];
#pragma unroll
        for (int b = 0; b <= A2; ++b) f[b] = Pd[n * LDP + 8 * b + g];
        double sa1[GM_NGW], sa2[GM_NGW];
#pragma unroll
        for (int u = 0; u < GM_NGW; ++u) {
            const double wv = wqb[(u0 + u) * GM_TROWS + n];
            sa1[u] = f[A1] * wv;
            sa2[u] = f[A2] * wv;
        }
#pragma unroll
        for (int b = 0; b <= A2; ++b) {
#pragma unroll
            for (int u = 0; u < GM_NGW; ++u) {
                if (b <= A1) dmma884(acc[u][b][0], acc[u][b][1], sa1[u], f[b]);
                if (A2 != A1) dmma884(acc[u][A1 + 1 + b][0], acc[u][A1 + 1 + b][1], sa2[u], f[b]);
            }
        }
        // MuBar: (Q x rows)(rows x NG): B fragment column g carries mbar of latent g (< NG), k = n.  One product gives all
        // NG latents of the CTA, so the two latent halves share the work: half 0 takes block row A1, half 1 block row A2
        if (WIDE || ug == 0 || A2 != A1) {
            const double bm = (g < NG) ? wmb[g * GM_TROWS + n] : 0.0;
            if (WIDE || ug == 0) dmma884(accm[0][0], accm[0][1], f[A1], bm);
            if ((WIDE || ug == 1) && A2 != A1) dmma884(accm[1][0], accm[1][1], f[A2], bm);
        }
    }
}
#define GM_ROLE(W)                                                                                                     \
    case W:                                                                                                            \
        if constexpr (gm_active(NB, W))                                                                                \
            gram_tile<NB, gm_a1(NB, W), gm_a2(NB, W), WIDE, GM_NG, NSLOT, LDP>(acc, accm, Pd, wqb, wmb, g, t, ug);     \
        break;

template <int NB>
__global__ void __launch_bounds__(GM_THREADS, NB > 8 ? 1 : 2)
k_gram_mma(const double* __restrict__ Pa, const double* __restrict__ Pb, const int* __restrict__ seg,
           const double* __restrict__ qbar, const double* __restrict__ mbar, double* __restrict__ SigBar,
           double* __restrict__ MuBar, long long B, int Q, int D, int mode) {
    using SH = GMShape<NB>;
    constexpr int LDP = SH::LDP, GM_NG = SH::NG, NSLOT = SH::NSLOT;
    constexpr bool WIDE = SH::WIDE;
    extern __shared__ __align__(16) double sm[];
    double* Pt = sm;                                        // [2][GM_TROWS][LDP]
    double* wq = Pt + 2 * (size_t)GM_TROWS * LDP;           // [2][GM_NG][GM_TROWS]
    double* wm = wq + 2 * GM_NG * GM_TROWS;                 // [2][GM_NG][GM_TROWS]
    const int i = blockIdx.x, s = blockIdx.z;
    const int ngroups = (D + GM_NG - 1) / GM_NG;
    int j0, nj;
    const double* P = Pa;
    if (mode == MODE_U && (int)blockIdx.y == ngroups) {     // the diagonal coefficient pair (i,i): L1 system rows
        j0 = i; nj = 1; P = Pb;
    } else {
        j0 = GM_NG * blockIdx.y;
        const int jlast = (mode == MODE_U) ? i - 1 : i;      // MODE_U groups cover the strictly-lower pairs only
        if (j0 > jlast) return;
        nj = min(GM_NG, jlast - j0 + 1);
    }
    const long long rbeg = seg[i], rend = seg[i + 1];
    if (rbeg >= rend) return;
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int w = WIDE ? (tid >> 5) : ((tid >> 5) & 3), ug = WIDE ? 0 : (tid >> 7);
    const int u0 = ug * GM_NGW;                            // this warp's latents: u0 .. u0 + GM_NGW - 1
    const int a1 = gm_a1(NB, w), a2 = gm_a2(NB, w);
    const bool active = gm_active(NB, w);
    const int nslots = !active ? 0 : (a1 == a2 ? a1 + 1 : a1 + a2 + 2);

    for (int e = tid; e < (int)SH::smem_doubles; e += GM_THREADS) sm[e] = 0.0;
    __syncthreads();

    double acc[GM_NGW][NSLOT][2];
    double accm[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int u = 0; u < GM_NGW; ++u)
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) acc[u][sl][0] = acc[u][sl][1] = 0.0;

    const long long ntiles = (rend - rbeg + GM_TROWS - 1) / GM_TROWS;
    auto stage = [&](long long tile, int buf) {
        const long long r0 = rbeg + tile * GM_TROWS;
        const int nr = (int)min((long long)GM_TROWS, rend - r0);
        double* Pd = Pt + (size_t)buf * GM_TROWS * LDP;
        for (int r = (tid >> 5); r < GM_TROWS; r += GM_THREADS / 32) {
            if (r < nr) {
                const double* src = &P[((size_t)s * B + r0 + r) * Q];
                if ((Q & 1) == 0) {
                    for (int c = 2 * lane; c < Q; c += 64) cp_async16(&Pd[r * LDP + c], &src[c]);
                } else {
                    for (int c = lane; c < Q; c += 32) cp_async8(&Pd[r * LDP + c], &src[c]);
                }
            } else {
                for (int c = lane; c < Q; c += 32) Pd[r * LDP + c] = 0.0;
            }
        }
        for (int e = tid; e < GM_NG * GM_TROWS; e += GM_THREADS) {
            int u = e / GM_TROWS, r = e - u * GM_TROWS;
            double* dq = &wq[(buf * GM_NG + u) * GM_TROWS + r];
            double* dm = &wm[(buf * GM_NG + u) * GM_TROWS + r];
            if (r < nr && u < nj) {
                const size_t o = ((size_t)s * B + r0 + r) * D + j0 + u;
                cp_async8(dq, &qbar[o]);
                cp_async8(dm, &mbar[o]);
            } else {
                *dq = 0.0;
                *dm = 0.0;
            }
        }
    };
    stage(0, 0);
    cp_async_commit();
    for (long long tile = 0; tile < ntiles; ++tile) {
        const int buf = (int)(tile & 1);
        cp_async_wait<0>();
        __syncthreads();
        if (tile + 1 < ntiles) stage(tile + 1, buf ^ 1);
        cp_async_commit();
        if (!active) continue;
        const double* Pd = Pt + (size_t)buf * GM_TROWS * LDP;
        const double* wqb = wq + (size_t)buf * GM_NG * GM_TROWS;
        const double* wmb = wm + (size_t)buf * GM_NG * GM_TROWS;
        switch (w) {                                         // the role is warp-uniform; each case is fully static
            GM_ROLE(0) GM_ROLE(1) GM_ROLE(2) GM_ROLE(3) GM_ROLE(4) GM_ROLE(5) GM_ROLE(6) GM_ROLE(7)
            default: break;
        }
    }
    cp_async_wait<0>();
    if (!active) return;
    // ---- write-out: block (a, bb) holds rows 8a+g, cols 8bb+2t+e; mirror the strictly-lower blocks ----------
#pragma unroll
    for (int uu = 0; uu < GM_NGW; ++uu) {
        const int u = u0 + uu;
        if (u < nj) {
            const int idx = (mode == MODE_U) ? pair_slot(i, j0 + u, D) : j0 + u;
            double* Sb = SigBar + (size_t)idx * Q * Q;
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl) {
                if (sl < nslots) {
                    const bool first = sl <= a1;
                    const int a = first ? a1 : a2, bb = first ? sl : sl - a1 - 1;
                    const int r = 8 * a + g;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * bb + 2 * t + e;
                        if (r < Q && c < Q) {
                            atomicAdd(&Sb[(size_t)r * Q + c], acc[uu][sl][e]);
                            if (a != bb) atomicAdd(&Sb[(size_t)c * Q + r], acc[uu][sl][e]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int st2 = 0; st2 < 2; ++st2) {
        if (st2 == 1 && a2 == a1) break;
        if (!WIDE && ug != st2) continue;                   // half 0 accumulated block row a1, half 1 block row a2
        const int a = st2 == 0 ? a1 : a2;
        const int r = 8 * a + g;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int u = 2 * t + e;
            if (u < nj && r < Q) {
                const int idx = (mode == MODE_U) ? pair_slot(i, j0 + u, D) : j0 + u;
                atomicAdd(&MuBar[(size_t)idx * Q + r], accm[st2][e]);
            }
        }
    }
}

template <int NB>
static int launch_gram(const double* Pa, const double* Pb, const int* seg, const double* qbar, const double* mbar,
                       double* SigBar, double* MuBar, int ns, long long B, int Q, int D, int mode, cudaStream_t st) {
    size_t smem = GMShape<NB>::smem_doubles * sizeof(double);
    if (int r = nmgp_opt_in_smem(k_gram_mma<NB>, smem, "nmgp_weighted_gram")) return r;
    constexpr int GM_NG = GMShape<NB>::NG;
    const int ngroups = (D + GM_NG - 1) / GM_NG;
    dim3 grid(D, ngroups + (mode == MODE_U ? 1 : 0), ns);
    k_gram_mma<NB><<<NMGP_L(grid), GM_THREADS, smem, st>>>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, B, Q, D, mode);
    return nmgp_launch_status("nmgp_weighted_gram(mma)");
}

// returns 1 if Q is outside the register-resident range (caller falls back to the FMA kernel)
int nmgp_weighted_gram_mma(const double* Pa, const double* Pb, const int* seg, const double* qbar, const double* mbar,
                           double* SigBar, double* MuBar, int ns, long long B, int Q, int D, int mode,
                           cudaStream_t st) {
    switch ((Q + 7) / 8) {
        case 1: return launch_gram<1>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 2: return launch_gram<2>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 3: return launch_gram<3>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 4: return launch_gram<4>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 5: return launch_gram<5>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 6: return launch_gram<6>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 7: return launch_gram<7>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 8: return launch_gram<8>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 9: return launch_gram<9>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 10: return launch_gram<10>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 11: return launch_gram<11>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 12: return launch_gram<12>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 13: return launch_gram<13>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 14: return launch_gram<14>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 15: return launch_gram<15>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        case 16: return launch_gram<16>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
        default: return 1;
    }
}

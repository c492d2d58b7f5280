// Diagonal quadratic forms through the variational covariances -- the FLOP-dominant part of the step:
//   q[n,j] = p_n^T Sigma_idx p_n,  m[n,j] = p_n . mu_idx      for j <= I[n]
// (reference: torch.mul(P.matmul(Sigma), P).sum(-1) and P mu at code/utils.py:120-122,143-144, evaluated there
// for ALL (pair, row) combinations; here only for the pairs a row consumes).
//   MODE_W: idx = j (latent functions, Sigma_W[j]);  MODE_U: idx = packed pair (I[n], j), p from the L1 system
//   when j == I[n] else from the L0 system.
// Rows are sorted by output id, so every (matrix, row-range) task is a contiguous range.
//
// This file holds the shared-memory FMA formulation (generic in Q).  The DMMA (mma.sync m8n8k4.f64)
// formulation for the large-batch case lives in nmgp_quadform_mma.cu.
#include "common.cuh"

#define QF_SUB 4  // threads per row

// DMMA formulations (nmgp_quadform_mma.cu), Q <= 64
int nmgp_coef_quadform_mma(bool bwd, const double* Pa, const double* Pb, const int* I, const double* Sig,
                           const double* Mu, double* q, double* m, const double* qbar, const double* mbar,
                           double* Pabar, double* Pbbar, int ns, long long B, int Q, int D, cudaStream_t st);

__device__ __forceinline__ void load_matrix(double* __restrict__ dst, const double* __restrict__ src, int n) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = src[e];
}

// grid (ceil(B/TR), ns); block TR*QF_SUB threads; thread (r,u) handles row r, columns b = u, u+4, ...
template <bool BWD>
__global__ void k_quadform(const double* __restrict__ Pa, const double* __restrict__ Pb, const int* __restrict__ I,
                           const double* __restrict__ Sig, const double* __restrict__ Mu,
                           // fwd outputs
                           double* __restrict__ qout, double* __restrict__ mout,
                           // bwd inputs/outputs
                           const double* __restrict__ qbar, const double* __restrict__ mbar,
                           double* __restrict__ Pabar, double* __restrict__ Pbbar,
                           long long B, int Q, int D, int mode) {
    extern __shared__ double sm[];
    const int TR = blockDim.x / QF_SUB, ldp = Q + 1;
    double* Sg = sm;                          // [Q*Q]
    double* Mus = Sg + (size_t)Q * Q;         // [Q]
    double* PaS = Mus + Q;                    // [TR][ldp]
    double* PbS = PaS + (size_t)TR * ldp;     // [TR][ldp]  (MODE_U only, else alias of PaS)
    double* GaS = PbS + (mode == MODE_U ? (size_t)TR * ldp : 0);   // BWD accumulators [TR][ldp]
    double* GbS = GaS + (size_t)TR * ldp;                           // BWD, MODE_U only
    if (mode != MODE_U) PbS = PaS;
    const int s = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * TR;
    const int nrows = (int)min((long long)TR, B - row0);
    const size_t base = ((size_t)s * B + row0) * Q;
    for (int e = threadIdx.x; e < TR * Q; e += blockDim.x) {
        int r = e / Q, a = e - r * Q;
        bool ok = r < nrows;
        PaS[r * ldp + a] = ok ? Pa[base + e] : 0.0;
        if (mode == MODE_U) PbS[r * ldp + a] = ok ? Pb[base + e] : 0.0;
        if (BWD) {
            GaS[r * ldp + a] = 0.0;
            if (mode == MODE_U) GbS[r * ldp + a] = 0.0;
        }
    }
    const int r = threadIdx.x / QF_SUB, u = threadIdx.x % QF_SUB;
    const bool rowok = r < nrows;
    const int myI = rowok ? I[row0 + r] : -1;
    const int i_lo = I[row0], i_hi = I[row0 + nrows - 1];
    const size_t obase = ((size_t)s * B + row0 + r) * D;
    __syncthreads();

    const int outer_lo = (mode == MODE_U) ? i_lo : 0;
    const int outer_hi = (mode == MODE_U) ? i_hi : 0;
    for (int io = outer_lo; io <= outer_hi; ++io) {
        const int jmax = (mode == MODE_U) ? io : i_hi;
        for (int j = 0; j <= jmax; ++j) {
            const int idx = (mode == MODE_U) ? pair_slot(io, j, D) : j;
            const bool active = rowok && ((mode == MODE_U) ? (myI == io) : (myI >= j));
            const bool useB = (mode == MODE_U) && (j == io);
            __syncthreads();   // previous task done with Sg
            load_matrix(Sg, Sig + (size_t)idx * Q * Q, Q * Q);
            load_matrix(Mus, Mu + (size_t)idx * Q, Q);
            __syncthreads();
            if (active) {   // NB: warp-divergent only at segment boundaries; shuffles below use the active mask of 4 lanes
                const double* p = (useB ? PbS : PaS) + r * ldp;
                double qacc = 0.0, macc = 0.0;
                double qb = 0.0, mb = 0.0;
                if (BWD) {
                    qb = qbar[obase + j];
                    mb = mbar[obase + j];
                }
                double* g = BWD ? ((useB ? GbS : GaS) + r * ldp) : nullptr;
                for (int b = u; b < Q; b += QF_SUB) {
                    double v0 = 0.0, v1 = 0.0;
                    int a = 0;
                    for (; a + 1 < Q; a += 2) {
                        v0 = fma(p[a], Sg[a * Q + b], v0);
                        v1 = fma(p[a + 1], Sg[(a + 1) * Q + b], v1);
                    }
                    if (a < Q) v0 = fma(p[a], Sg[a * Q + b], v0);
                    double v = v0 + v1;
                    if (BWD) {
                        // pbar += qbar * (Sigma + Sigma^T) p + mbar * mu ; Sigma is exactly symmetric by construction
                        g[b] += 2.0 * qb * v + mb * Mus[b];
                    } else {
                        qacc = fma(v, p[b], qacc);
                        macc = fma(p[b], Mus[b], macc);
                    }
                }
                if (!BWD) {
                    // reduce over the QF_SUB lanes of this row (adjacent lanes, same activity)
                    unsigned mask = __activemask();
                    qacc += __shfl_xor_sync(mask, qacc, 1);
                    qacc += __shfl_xor_sync(mask, qacc, 2);
                    macc += __shfl_xor_sync(mask, macc, 1);
                    macc += __shfl_xor_sync(mask, macc, 2);
                    if (u == 0) {
                        qout[obase + j] = qacc;
                        mout[obase + j] = macc;
                    }
                }
            }
        }
    }
    if (BWD) {
        __syncthreads();
        for (int e = threadIdx.x; e < nrows * Q; e += blockDim.x) {
            int rr = e / Q, a = e - rr * Q;
            Pabar[base + e] = GaS[rr * ldp + a];
            if (mode == MODE_U) Pbbar[base + e] = GbS[rr * ldp + a];
        }
    }
}

static int quadform_rows(int Q, int mode, bool bwd) {
    int ntile = 1 + (mode == MODE_U ? 1 : 0) + (bwd ? (mode == MODE_U ? 2 : 1) : 0);
    int TR = 64;
    while (TR > 8 && 8.0 * ((double)Q * Q + Q + (double)ntile * TR * (Q + 1)) > 200.0 * 1024) TR >>= 1;
    return TR;
}
static size_t quadform_smem(int Q, int mode, bool bwd, int TR) {
    int ntile = 1 + (mode == MODE_U ? 1 : 0) + (bwd ? (mode == MODE_U ? 2 : 1) : 0);
    return sizeof(double) * ((size_t)Q * Q + Q + (size_t)ntile * TR * (Q + 1));
}

NMGP_API int nmgp_quadform_fwd(const double* Pa, const double* Pb, const int* I, const int* seg, const double* Sig,
                               const double* Mu, double* q /* pre-zeroed */, double* m /* pre-zeroed */, int ns,
                               long long B, int Q, int D, int mode, cudaStream_t st) {
    (void)seg;
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128 && D > 0 && (mode == MODE_W || mode == MODE_U),
                 "nmgp_quadform_fwd");
    if (ns == 0 || B == 0) return 0;
    if (mode == MODE_U && Q <= 64)
        return nmgp_coef_quadform_mma(false, Pa, Pb, I, Sig, Mu, q, m, nullptr, nullptr, nullptr, nullptr, ns, B, Q, D, st);
    const int TR = quadform_rows(Q, mode, false);
    size_t smem = quadform_smem(Q, mode, false, TR);
    if (int r = nmgp_opt_in_smem(k_quadform<false>, smem, "nmgp_quadform_fwd")) return r;
    dim3 grid((unsigned)((B + TR - 1) / TR), ns);
    k_quadform<false><<<NMGP_L(grid), TR * QF_SUB, smem, st>>>(Pa, Pb, I, Sig, Mu, q, m, nullptr, nullptr, nullptr, nullptr, B,
                                                       Q, D, mode);
    return nmgp_launch_status("nmgp_quadform_fwd");
}

NMGP_API int nmgp_quadform_bwd(const double* Pa, const double* Pb, const int* I, const int* seg, const double* Sig,
                               const double* Mu, const double* qbar, const double* mbar, double* Pabar, double* Pbbar,
                               int ns, long long B, int Q, int D, int mode, cudaStream_t st) {
    (void)seg;
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128 && D > 0 && (mode == MODE_W || mode == MODE_U),
                 "nmgp_quadform_bwd");
    if (ns == 0 || B == 0) return 0;
    if (mode == MODE_U && Q <= 64)
        return nmgp_coef_quadform_mma(true, Pa, Pb, I, Sig, Mu, nullptr, nullptr, qbar, mbar, Pabar, Pbbar, ns, B, Q, D,
                                      st);
    const int TR = quadform_rows(Q, mode, true);
    size_t smem = quadform_smem(Q, mode, true, TR);
    if (int r = nmgp_opt_in_smem(k_quadform<true>, smem, "nmgp_quadform_bwd")) return r;
    dim3 grid((unsigned)((B + TR - 1) / TR), ns);
    k_quadform<true><<<NMGP_L(grid), TR * QF_SUB, smem, st>>>(Pa, Pb, I, Sig, Mu, nullptr, nullptr, qbar, mbar, Pabar, Pbbar, B,
                                                      Q, D, mode);
    return nmgp_launch_status("nmgp_quadform_bwd");
}

// ------------------------------------------------------------------------------------------------
// SigBar[idx] += sum_{s, n in task rows} qbar[s,n,j] p p^T ;  MuBar[idx] += sum mbar[s,n,j] p
// grid (chunks, ntasks, ns).  MODE_W: task = j, rows [seg[j], B).  MODE_U: task = packed pair slot, rows of output i.
#define WG_SUB 32     // rows staged per pass
#define WG_CHUNK 1024 // rows per CTA
__global__ void k_weighted_gram(const double* __restrict__ Pa, const double* __restrict__ Pb,
                                const int* __restrict__ seg, const double* __restrict__ qbar,
                                const double* __restrict__ mbar, double* __restrict__ SigBar,
                                double* __restrict__ MuBar, long long B, int Q, int D, int mode) {
    extern __shared__ double sm[];
    const int ldp = Q + 1;
    double* Acc = sm;                         // [Q*Q]
    double* MAcc = Acc + (size_t)Q * Q;       // [Q]
    double* Pt = MAcc + Q;                    // [WG_SUB][ldp]
    double* wq = Pt + (size_t)WG_SUB * ldp;   // [WG_SUB]
    double* wm = wq + WG_SUB;                 // [WG_SUB]
    const int task = blockIdx.y, s = blockIdx.z;
    int i, j;
    if (mode == MODE_W) {
        j = task; i = -1;
    } else if (task < D) {
        i = task; j = task;
    } else {
        int t = task - D;                    // t = i(i-1)/2 + j, j < i
        i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
        while (i * (i - 1) / 2 > t) --i;
        while ((i + 1) * i / 2 <= t) ++i;
        j = t - i * (i - 1) / 2;
    }
    const long long rbeg = (mode == MODE_W) ? seg[j] : seg[i];
    const long long rend = (mode == MODE_W) ? B : seg[i + 1];
    const long long c0 = rbeg + (long long)blockIdx.x * WG_CHUNK;
    if (c0 >= rend) return;
    const long long c1 = min(rend, c0 + WG_CHUNK);
    const double* P = (mode == MODE_U && i == j) ? Pb : Pa;
    for (int e = threadIdx.x; e < Q * Q + Q; e += blockDim.x) Acc[e] = 0.0;   // Acc and MAcc are contiguous
    for (long long t0 = c0; t0 < c1; t0 += WG_SUB) {
        const int nr = (int)min((long long)WG_SUB, c1 - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * Q; e += blockDim.x) {
            int r = e / Q, a = e - r * Q;
            Pt[r * ldp + a] = P[((size_t)s * B + t0) * Q + e];
        }
        for (int r = threadIdx.x; r < nr; r += blockDim.x) {
            size_t o = ((size_t)s * B + t0 + r) * D + j;
            wq[r] = qbar[o];
            wm[r] = mbar[o];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
            int a = e / Q, b = e - a * Q;
            double acc = 0.0;
            for (int r = 0; r < nr; ++r) acc = fma(wq[r] * Pt[r * ldp + a], Pt[r * ldp + b], acc);
            Acc[e] += acc;
        }
        for (int a = threadIdx.x; a < Q; a += blockDim.x) {
            double acc = 0.0;
            for (int r = 0; r < nr; ++r) acc = fma(wm[r], Pt[r * ldp + a], acc);
            MAcc[a] += acc;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) atomicAdd(&SigBar[(size_t)task * Q * Q + e], Acc[e]);
    for (int a = threadIdx.x; a < Q; a += blockDim.x) atomicAdd(&MuBar[(size_t)task * Q + a], MAcc[a]);
}
int nmgp_weighted_gram_mma(const double* Pa, const double* Pb, const int* seg, const double* qbar, const double* mbar,
                           double* SigBar, double* MuBar, int ns, long long B, int Q, int D, int mode, cudaStream_t st);

NMGP_API int nmgp_weighted_gram(const double* Pa, const double* Pb, const int* I, const int* seg, const double* qbar,
                                const double* mbar, double* SigBar, double* MuBar, int ns, long long B, int Q, int D,
                                int mode, cudaStream_t st) {
    (void)I;
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 0 && Q <= 128 && D > 0 && (mode == MODE_W || mode == MODE_U),
                 "nmgp_weighted_gram");
    if (ns == 0 || B == 0) return 0;
    if (Q <= 128) return nmgp_weighted_gram_mma(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, ns, B, Q, D, mode, st);
    const int ntasks = (mode == MODE_W) ? D : D * (D + 1) / 2;
    NMGP_REQUIRE(ntasks <= 65535, "nmgp_weighted_gram");
    size_t smem = sizeof(double) * ((size_t)Q * Q + Q + (size_t)WG_SUB * (Q + 1) + 2 * WG_SUB);
    if (int r = nmgp_opt_in_smem(k_weighted_gram, smem, "nmgp_weighted_gram")) return r;
    dim3 grid((unsigned)((B + WG_CHUNK - 1) / WG_CHUNK), ntasks, ns);
    k_weighted_gram<<<NMGP_L(grid), 256, smem, st>>>(Pa, Pb, seg, qbar, mbar, SigBar, MuBar, B, Q, D, mode);
    return nmgp_launch_status("nmgp_weighted_gram");
}

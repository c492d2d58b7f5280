// FP64 tensor-core (DMMA) quadratic-form kernels for 64 < Q <= 128 inducing points (the PM2.5 / HCP drivers use
// Q = 100: code/NMGP_PM25.py:219, code/NMGP_HCP.py:210).  Same mathematics as nmgp_quadform_mma.cu
// (code/utils.py:120-122,143-144 + code/nmgp_dsvi.py:255-258 and their autograd) but the register-resident tiling of
// that file stops at Q = 64, so here
//   * a CTA owns 128 rows (8 warps x 16 rows), one CTA per SM; the P tile lives in shared memory and A fragments are
//     read from it (they no longer fit in registers);
//   * V = P Sigma[slot] is produced in two column halves (<= 8 blocks of 8 columns each), so that V (one half) and the
//     adjoint accumulator Pbar (all columns) fit in registers; the adjoint weight 2 qbar does not depend on the
//     quadratic form itself, so Pbar += 2 qbar V is applied half by half;
//   * Sigma[slot] does not fit in shared memory twice (Q = 100: 87 KB), so every record is pre-padded in global memory
//     in consumption order [half][k][columns of the half] and streamed through a 4-stage ring of ~16 KB chunks by the
//     bulk-copy engine (cp.async.bulk + mbarrier complete_tx; one elected producer thread, per-stage full/empty
//     mbarriers, no CTA-wide barrier inside the task loop).
// One kernel template serves three modes:
//   LQ_W  : latent side, fused with the expected log-likelihood and every row cotangent (= k_latent_fused);
//   LQ_UF : coefficient side forward   q[n,j], m[n,j] for the pairs (I[n], j) the row consumes (= k_coef_quadform_mma<fwd>);
//   LQ_UB : coefficient side adjoint   Pbar += 2 qbar Sigma p + mbar mu                        (= k_coef_quadform_mma<bwd>).
// The coefficient modes run twice: strictly-lower pairs (j < i) on the L0 system rows, diagonal pairs (i, i) on the L1
// system rows (`diag`).
#include "common.cuh"

namespace {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
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

__host__ __device__ constexpr int pad4mod8(int n) { return ((n + 3) / 8) * 8 + 4; }      // smallest m >= n, m % 8 == 4
__host__ __device__ constexpr int pad8mod16(int n) { return ((n + 7) / 16) * 16 + 8; }   // smallest m >= n, m % 16 == 8

enum { LQ_W = 0, LQ_UF = 1, LQ_UB = 2 };
#define LQ_ROWS 128
#define LQ_THREADS 256
#define LQ_WARPS 8
#define LQ_NST 4

template <int NB>
struct LQShape {
    static constexpr int NB0 = (NB + 1) / 2;               // 8-column blocks of half 0 (half 1: NB - NB0)
    static constexpr int NB1 = NB - NB0;
    static constexpr int NP = 8 * NB;                      // padded columns
    static constexpr int KS = 2 * NB;                      // k-steps of 4 (K padded like N)
    static constexpr int KP = 4 * KS;
    static constexpr int LDP = pad8mod16(NP);              // row stride of the P tile (conflict-free 16-byte C-layout reads)
    static constexpr int LDSH = pad4mod8(8 * NB0);         // row stride of a half record (conflict-free B-fragment reads)
    static constexpr int CR = ((16384 / (LDSH * 8)) / 4) * 4;     // k-rows per ring chunk (multiple of 4, ~16 KB)
    static constexpr int NCH = (KP + CR - 1) / CR;         // chunks per half
    static constexpr int HALF = KP * LDSH;                 // doubles of one half record
    static constexpr int REC = 2 * HALF;                   // doubles of one padded record
    static constexpr size_t smem_bytes = sizeof(double) * ((size_t)LQ_ROWS * LDP + (size_t)LQ_NST * CR * LDSH + 2 * LQ_ROWS)
                                         + sizeof(int) * LQ_ROWS + sizeof(unsigned long long) * 2 * LQ_NST;
};

// rec[s] = [half][k < KP][c < LDSH] of Sigma[s] (Q x Q), zero padding
template <int NB>
__global__ void k_lq_pad_records(const double* __restrict__ Sig, double* __restrict__ rec, int Q) {
    using SH = LQShape<NB>;
    const int s = blockIdx.x;
    double* out = rec + (size_t)s * SH::REC;
    const double* S = Sig + (size_t)s * Q * Q;
    for (int e = threadIdx.x; e < SH::REC; e += blockDim.x) {
        const int h = e / SH::HALF, r = e - h * SH::HALF;
        const int k = r / SH::LDSH, c = r - k * SH::LDSH;
        const int col = 8 * SH::NB0 * h + c;
        const bool in = (k < Q) && (col < Q) && (c < 8 * (h == 0 ? SH::NB0 : SH::NB1));
        out[e] = in ? S[(size_t)k * Q + col] : 0.0;
    }
}

struct LQArgs {
    const double* P;        // [ns, B, Q] rows (latent side: P_G; coefficient side: L0 / L1 system rows)
    const double* rec;      // [nslot][REC] padded covariance records
    const double* Mu;       // [nslot][Q]
    const int* I;
    long long B;
    int Q, D, diag;
    // LQ_W
    const double* cG; const double* l; const double* y; long long ystride; const double* hyp; double scale;
    double* Rsum; double* ghyp; double* lbar; double* mgbar; double* qgbar; double* cGbar;
    // LQ_UF
    double* qout; double* mout;
    // LQ_UB
    const double* qbar; const double* mbar;
    double* Pbar;           // LQ_W: PGbar; LQ_UB: Pabar (lower pairs) / Pbbar (diagonal pairs)
};

// task cursor: LQ_W walks latents j = 0..jmax; the coefficient modes walk (i, j < i) or (i, i) over the outputs of the tile
template <int MODE>
struct Cursor {
    int i, j;
    __device__ __forceinline__ void first(int i_lo, int diag) {
        if (MODE == LQ_W) { i = 0; j = 0; }
        else if (diag) { i = i_lo; j = i_lo; }
        else { i = i_lo > 1 ? i_lo : 1; j = 0; }
    }
    __device__ __forceinline__ void next(int diag) {
        if (MODE == LQ_W) { ++j; }
        else if (diag) { ++i; j = i; }
        else if (++j >= i) { ++i; j = 0; }
    }
    __device__ __forceinline__ bool valid(int i_hi, int jmax) const { return MODE == LQ_W ? j <= jmax : i <= i_hi; }
    __device__ __forceinline__ int slot(int D) const { return MODE == LQ_W ? j : pair_slot(i, j, D); }
};

template <int NB, int MODE>
__global__ void __launch_bounds__(LQ_THREADS, 1) k_lq(const LQArgs a) {
    using SH = LQShape<NB>;
    constexpr int NB0 = SH::NB0, NB1 = SH::NB1, KS = SH::KS, LDP = SH::LDP, LDSH = SH::LDSH, CR = SH::CR, NCH = SH::NCH;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* ring = reinterpret_cast<double*>(smraw);                  // [LQ_NST][CR][LDSH]   (128-byte aligned chunks)
    double* Ps = ring + (size_t)LQ_NST * CR * LDSH;                    // [LQ_ROWS][LDP]
    double* rrs = Ps + (size_t)LQ_ROWS * LDP;                          // [LQ_ROWS]
    double* omcs = rrs + LQ_ROWS;                                      // [LQ_ROWS]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(omcs + LQ_ROWS);   // [LQ_NST]
    unsigned long long* empty = full + LQ_NST;                         // [LQ_NST]
    int* Is = reinterpret_cast<int*>(empty + LQ_NST);                  // [LQ_ROWS]

    const int Q = a.Q, D = a.D, diag = a.diag;
    const long long B = a.B;
    const int s = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * LQ_ROWS;
    const int nrows = (int)min((long long)LQ_ROWS, B - row0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const size_t rbase = (size_t)s * B + row0;

    for (int e = tid; e < LQ_ROWS * LDP; e += LQ_THREADS) {
        const int r = e / LDP, c = e - r * LDP;
        Ps[e] = (r < nrows && c < Q) ? a.P[(rbase + r) * Q + c] : 0.0;
    }
    if (tid == 0) {
        for (int k = 0; k < LQ_NST; ++k) {
            mbar_init(&full[k], 1);
            mbar_init(&empty[k], LQ_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int r = tid; r < LQ_ROWS; r += LQ_THREADS) {
        Is[r] = r < nrows ? a.I[row0 + r] : -1;
        if (MODE == LQ_W) omcs[r] = r < nrows ? 1.0 - a.cG[rbase + r] : 0.0;
    }
    __syncthreads();

    const int rloc[2] = {16 * w + g, 16 * w + 8 + g};
    const int myI[2] = {Is[rloc[0]], Is[rloc[1]]};
    const int i_lo = Is[0], i_hi = Is[nrows - 1];
    const int jmax = i_hi;
    const int wlo = (16 * w < nrows) ? Is[16 * w] : (1 << 30);
    const int whi = (16 * w < nrows) ? Is[min(16 * w + 15, nrows - 1)] : -1;
    const double* Pw = Ps + (size_t)(16 * w) * LDP;                    // this warp's 16 rows
    double s2e = 1.0;
    if (MODE == LQ_W) s2e = a.hyp[H_S2_ERR];

    // ---- phase 1: means m[n, j] = p_n . Mu[slot] on the tensor pipe (latents / pairs as the N dimension) -----------
    double Fp[2] = {0.0, 0.0};
    if (MODE != LQ_UB) {                                                // the adjoint needs no means
        const int o_lo = (MODE == LQ_W) ? 0 : i_lo, o_hi = (MODE == LQ_W) ? 0 : i_hi;
        for (int io = o_lo; io <= o_hi; ++io) {
            const int jcount = (MODE == LQ_W) ? D : (diag ? 1 : io);    // columns of this output's product
            if (MODE != LQ_W && (io < wlo || io > whi)) continue;
            for (int jb = 0; jb * 8 < jcount; ++jb) {
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                const int jn = 8 * jb + g;                              // B-fragment column of this lane
                int slot_n = -1;
                if (jn < jcount) slot_n = (MODE == LQ_W) ? jn : pair_slot(io, diag ? io : jn, D);
#pragma unroll 2
                for (int ks = 0; ks < KS; ++ks) {
                    const int k = 4 * ks + t;
                    const double b = (slot_n >= 0 && k < Q) ? __ldg(&a.Mu[(size_t)slot_n * Q + k]) : 0.0;
                    dmma884(acc[0][0], acc[0][1], Pw[g * LDP + k], b);
                    dmma884(acc[1][0], acc[1][1], Pw[(8 + g) * LDP + k], b);
                }
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    if (rloc[mb] >= nrows) continue;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int jc = 8 * jb + 2 * t + e;
                        if (jc >= jcount) continue;
                        if (MODE == LQ_W) {
                            if (jc <= myI[mb]) {
                                const size_t o = (rbase + rloc[mb]) * D + jc;
                                Fp[mb] = fma(a.l[o], acc[mb][e], Fp[mb]);
                                a.mgbar[o] = acc[mb][e];            // parked in the slot its cotangent overwrites at latent jc
                            }
                        } else if (myI[mb] == io) {
                            a.mout[(rbase + rloc[mb]) * D + (diag ? io : jc)] = acc[mb][e];
                        }
                    }
                }
            }
        }
    }
    double racc = 0.0, gacc = 0.0, rr[2] = {0.0, 0.0};
    if (MODE == LQ_W) {
        const double cst = -0.5 * log(s2e) - log(sqrt(2.0 * 3.14159265358979323846));
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            double F = Fp[mb];
            F += __shfl_xor_sync(0xffffffffu, F, 1);
            F += __shfl_xor_sync(0xffffffffu, F, 2);
            double r = 0.0;
            if (rloc[mb] < nrows) r = a.y[(size_t)s * a.ystride + row0 + rloc[mb]] - F;
            rr[mb] = r / s2e;
            if (t == 0 && rloc[mb] < nrows) {
                rrs[rloc[mb]] = rr[mb];
                racc += -(r * r) / (2.0 * s2e) + cst;
                gacc += (r * r) / (2.0 * s2e) - 0.5;
            }
        }
        // entries of latents a row does not use (j > I[n]) are exact zeros in every [ns,B,D] output
        const int r = tid >> 1;
        if (r < nrows)
            for (int j = Is[r] + 1 + (tid & 1); j < D; j += 2) {
                const size_t o = (rbase + r) * D + j;
                a.lbar[o] = 0.0;
                a.mgbar[o] = 0.0;
                a.qgbar[o] = 0.0;
            }
    }
    __syncthreads();                       // the parked means (global) are read back by other lanes of the warp

    // ---- phase 2: task loop over the ring ---------------------------------------------------------------------------
    double pacc[2][NB][2];
    if (MODE != LQ_UF) {
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) pacc[mb][nb][0] = pacc[mb][nb][1] = 0.0;
    }
    double pen[2] = {0.0, 0.0}, gsum[2] = {0.0, 0.0};

    // producer state (thread 0 only): cursor and chunk position of the NEXT chunk to issue
    Cursor<MODE> pc;
    pc.first(i_lo, diag);
    int p_h = 0, p_kc = 0, p_n = 0;                                     // half, chunk-in-half, flat chunk index
    auto issue = [&]() {                                                // thread 0: issue chunk p_n if any
        if (!pc.valid(i_hi, jmax)) return;
        const int st = p_n % LQ_NST;
        if (p_n >= LQ_NST) mbar_wait(&empty[st], (unsigned)(((p_n / LQ_NST) - 1) & 1));
        const int k0 = p_kc * CR, rows = min(CR, SH::KP - k0);
        const unsigned bytes = (unsigned)(rows * LDSH * sizeof(double));
        const double* src = a.rec + (size_t)pc.slot(D) * SH::REC + (size_t)p_h * SH::HALF + (size_t)k0 * LDSH;
        mbar_expect_tx(&full[st], bytes);
        bulk_g2s(ring + (size_t)st * CR * LDSH, src, bytes, &full[st]);
        ++p_n;
        if (++p_kc == NCH) {
            p_kc = 0;
            if (++p_h == 2) { p_h = 0; pc.next(diag); }
        }
    };
    if (tid == 0)
        for (int k = 0; k < LQ_NST - 1; ++k) issue();

    Cursor<MODE> cc;
    cc.first(i_lo, diag);
    int cn = 0;                                                         // flat index of the chunk being consumed
    while (cc.valid(i_hi, jmax)) {
        const int ti = cc.i, tj = cc.j;
        const bool skip = (MODE == LQ_W) ? (whi < tj) : (ti < wlo || ti > whi);
        bool live[2];
        double g2[2] = {0.0, 0.0}, lcur[2] = {0.0, 0.0}, mcur[2] = {0.0, 0.0};
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            live[mb] = (rloc[mb] < nrows) && ((MODE == LQ_W) ? (tj <= myI[mb]) : (myI[mb] == ti));
            if (live[mb] && !skip) {
                const size_t o = (rbase + rloc[mb]) * D + tj;
                if (MODE == LQ_W) {
                    lcur[mb] = __ldg(&a.l[o]);
                    mcur[mb] = __ldcg(&a.mgbar[o]);
                    g2[mb] = 2.0 * a.scale * (0.5 / s2e) * lcur[mb] * lcur[mb];
                } else if (MODE == LQ_UB) {
                    g2[mb] = 2.0 * __ldg(&a.qbar[o]);
                }
            }
        }
        double qp[2] = {0.0, 0.0};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            constexpr int NBHMAX = NB0;
            const int nbh = h == 0 ? NB0 : NB1;
            double V[2][NBHMAX][2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < NBHMAX; ++nb) V[mb][nb][0] = V[mb][nb][1] = 0.0;
            for (int kc = 0; kc < NCH; ++kc) {
                if (tid == 0) issue();                                  // keeps LQ_NST - 1 chunks in flight
                const int st = cn % LQ_NST;
                mbar_wait(&full[st], (unsigned)((cn / LQ_NST) & 1));
                if (!skip) {
                    const double* Sd = ring + (size_t)st * CR * LDSH;
                    const int k0 = kc * CR, nks = min(CR, SH::KP - k0) / 4;
#pragma unroll 2
                    for (int ks = 0; ks < nks; ++ks) {
                        const double a0 = Pw[g * LDP + k0 + 4 * ks + t];
                        const double a1 = Pw[(8 + g) * LDP + k0 + 4 * ks + t];
#pragma unroll
                        for (int nb = 0; nb < NBHMAX; ++nb) {
                            if (nb < nbh) {
                                const double b = Sd[(4 * ks + t) * LDSH + 8 * nb + g];
                                dmma884(V[0][nb][0], V[0][nb][1], a0, b);
                                dmma884(V[1][nb][0], V[1][nb][1], a1, b);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
                ++cn;
            }
            if (!skip) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    double q0 = 0.0, q1 = 0.0;
#pragma unroll
                    for (int nb = 0; nb < NBHMAX; ++nb) {
                        if (nb < nbh) {
                            const int nbg = h * NB0 + nb;
                            const double2 pv = *reinterpret_cast<const double2*>(&Ps[rloc[mb] * LDP + 8 * nbg + 2 * t]);
                            q0 = fma(V[mb][nb][0], pv.x, q0);
                            q1 = fma(V[mb][nb][1], pv.y, q1);
                            if (MODE != LQ_UF) {
                                pacc[mb][nbg][0] = fma(g2[mb], V[mb][nb][0], pacc[mb][nbg][0]);
                                pacc[mb][nbg][1] = fma(g2[mb], V[mb][nb][1], pacc[mb][nbg][1]);
                            }
                        }
                    }
                    qp[mb] += q0 + q1;
                }
            }
        }
        if (!skip && MODE != LQ_UB) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                double q = qp[mb];
                q += __shfl_xor_sync(0xffffffffu, q, 1);
                q += __shfl_xor_sync(0xffffffffu, q, 2);
                if (live[mb] && t == 0) {
                    const size_t o = (rbase + rloc[mb]) * D + tj;
                    if (MODE == LQ_W) {
                        const double lj = lcur[mb], gq = 0.5 * g2[mb];
                        const double s2g = omcs[rloc[mb]] + q;
                        pen[mb] = fma(lj * lj, s2g, pen[mb]);
                        gsum[mb] += gq;
                        a.lbar[o] = -a.scale * (rr[mb] * mcur[mb] - (1.0 / s2e) * lj * s2g);
                        a.qgbar[o] = gq;
                        a.mgbar[o] = -a.scale * rr[mb] * lj;
                    } else {
                        a.qout[o] = q;
                    }
                }
            }
        }
        cc.next(diag);
    }

    // ---- Pbar += mbar Mu (mean path of the adjoint): one DMMA product over the latents / pairs ----------------------
    if (MODE != LQ_UF) {
        __syncthreads();                   // LQ_W: the mbar entries just written by the t == 0 lanes
        const double* mb_src = (MODE == LQ_W) ? a.mgbar : a.mbar;
        const int o_lo = (MODE == LQ_W) ? 0 : i_lo, o_hi = (MODE == LQ_W) ? 0 : i_hi;
        for (int io = o_lo; io <= o_hi; ++io) {
            if (MODE != LQ_W && (io < wlo || io > whi)) continue;
            const int jcount = (MODE == LQ_W) ? D : (diag ? 1 : io);
            for (int ks = 0; 4 * ks < jcount; ++ks) {
                const int jj = 4 * ks + t;
                const int jcol = (MODE == LQ_W) ? jj : (diag ? io : jj);
                double av[2];
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    const bool ok = rloc[mb] < nrows && jj < jcount && (MODE == LQ_W || myI[mb] == io);
                    av[mb] = ok ? __ldcg(&mb_src[(rbase + rloc[mb]) * D + jcol]) : 0.0;
                }
                const int slot = (jj < jcount) ? ((MODE == LQ_W) ? jj : pair_slot(io, jcol, D)) : -1;
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const int c = 8 * nb + g;
                    const double b = (slot >= 0 && c < Q) ? __ldg(&a.Mu[(size_t)slot * Q + c]) : 0.0;
                    dmma884(pacc[0][nb][0], pacc[0][nb][1], av[0], b);
                    dmma884(pacc[1][nb][0], pacc[1][nb][1], av[1], b);
                }
            }
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int rl = rloc[mb];
            // coefficient adjoint: rows of outputs without a task of this launch (output 0 has no strictly-lower pair)
            // still own their Pbar row: it is written as zeros
            if (rl < nrows) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * nb + 2 * t + e;
                        if (c < Q) a.Pbar[(rbase + rl) * Q + c] = pacc[mb][nb][e];
                    }
            }
        }
    }
    if (MODE == LQ_W) {
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            if (rloc[mb] < nrows && t == 0) {
                a.cGbar[rbase + rloc[mb]] = -gsum[mb];
                racc -= (0.5 / s2e) * pen[mb];
                gacc += (0.5 / s2e) * pen[mb];
            }
        }
        racc = block_sum(racc);
        gacc = block_sum(gacc);
        if (tid == 0) {
            atomicAdd(&a.Rsum[s], racc);
            atomicAdd(&a.ghyp[H_S2_ERR], -a.scale * gacc);
        }
    }
}

template <int NB, int MODE>
int launch_lq(const LQArgs& a, int ns, cudaStream_t st, const char* what) {
    using SH = LQShape<NB>;
    if (int r = nmgp_opt_in_smem(k_lq<NB, MODE>, SH::smem_bytes, what)) return r;
    dim3 grid((unsigned)((a.B + LQ_ROWS - 1) / LQ_ROWS), ns);
    k_lq<NB, MODE><<<NMGP_L(grid), LQ_THREADS, SH::smem_bytes, st>>>(a);
    return nmgp_launch_status(what);
}

#define LQ_SWITCH(MODE, args, ns, st, what)                       \
    switch ((args.Q + 7) / 8) {                                   \
        case 9: return launch_lq<9, MODE>(args, ns, st, what);    \
        case 10: return launch_lq<10, MODE>(args, ns, st, what);  \
        case 11: return launch_lq<11, MODE>(args, ns, st, what);  \
        case 12: return launch_lq<12, MODE>(args, ns, st, what);  \
        case 13: return launch_lq<13, MODE>(args, ns, st, what);  \
        case 14: return launch_lq<14, MODE>(args, ns, st, what);  \
        case 15: return launch_lq<15, MODE>(args, ns, st, what);  \
        case 16: return launch_lq<16, MODE>(args, ns, st, what);  \
        default: nmgp_set_error("%s: Q = %d outside 65..128", what, args.Q); return -1; \
    }

int rec_doubles(int Q) {
    switch ((Q + 7) / 8) {
        case 9: return LQShape<9>::REC;
        case 10: return LQShape<10>::REC;
        case 11: return LQShape<11>::REC;
        case 12: return LQShape<12>::REC;
        case 13: return LQShape<13>::REC;
        case 14: return LQShape<14>::REC;
        case 15: return LQShape<15>::REC;
        case 16: return LQShape<16>::REC;
        default: return 0;
    }
}

}  // namespace

// doubles per padded record for this Q (0 when Q is outside 65..128)
NMGP_API long long nmgp_lq_record_doubles(int Q) { return Q > 64 ? rec_doubles(Q) : 0; }

// rec[s] (s < n) = padded, half-split copy of Sig[s] (Q x Q) in the order the ring consumes it
NMGP_API int nmgp_lq_pad_records(const double* Sig, double* rec, int n, int Q, cudaStream_t st) {
    NMGP_REQUIRE(n >= 0 && Q > 64 && Q <= 128, "nmgp_lq_pad_records");
    if (n == 0) return 0;
    switch ((Q + 7) / 8) {
        case 9: k_lq_pad_records<9><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
        case 10: k_lq_pad_records<10><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
        case 11: k_lq_pad_records<11><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
        case 12: k_lq_pad_records<12><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
        case 13: k_lq_pad_records<13><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
        case 14: k_lq_pad_records<14><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
        case 15: k_lq_pad_records<15><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
        default: k_lq_pad_records<16><<<NMGP_L(n), 256, 0, st>>>(Sig, rec, Q); break;
    }
    return nmgp_launch_status("nmgp_lq_pad_records");
}

// = nmgp_latent_fused for 64 < Q <= 128; recW = nmgp_lq_pad_records(Sigma_W)
NMGP_API int nmgp_lq_latent_fused(const double* PG, const double* cG, const double* l, const double* y, const int* I,
                                  const double* recW, const double* muW, const double* hyp, double scale, double* Rsum,
                                  double* ghyp, double* lbar, double* mgbar, double* qgbar, double* cGbar, double* PGbar,
                                  int ns, long long B, int Q, int D, long long ystride, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 64 && Q <= 128 && D > 0, "nmgp_lq_latent_fused");
    if (ns == 0 || B == 0) return 0;
    LQArgs a = {};
    a.P = PG; a.rec = recW; a.Mu = muW; a.I = I; a.B = B; a.Q = Q; a.D = D; a.diag = 0;
    a.cG = cG; a.l = l; a.y = y; a.ystride = ystride; a.hyp = hyp; a.scale = scale; a.Rsum = Rsum; a.ghyp = ghyp;
    a.lbar = lbar; a.mgbar = mgbar; a.qgbar = qgbar; a.cGbar = cGbar; a.Pbar = PGbar;
    LQ_SWITCH(LQ_W, a, ns, st, "nmgp_lq_latent_fused")
}

static int lq_coef(int mode, int diag, const double* P, const int* I, const double* recU, const double* Mu, double* q,
                   double* m, const double* qbar, const double* mbar, double* Pbar, int ns, long long B, int Q, int D,
                   cudaStream_t st) {
    LQArgs a = {};
    a.P = P; a.rec = recU; a.Mu = Mu; a.I = I; a.B = B; a.Q = Q; a.D = D; a.diag = diag;
    a.qout = q; a.mout = m; a.qbar = qbar; a.mbar = mbar; a.Pbar = Pbar;
    if (mode == LQ_UF) { LQ_SWITCH(LQ_UF, a, ns, st, "nmgp_lq_coef_quadform(fwd)") }
    LQ_SWITCH(LQ_UB, a, ns, st, "nmgp_lq_coef_quadform(bwd)")
}

// = nmgp_quadform_fwd / nmgp_quadform_bwd in MODE_U for 64 < Q <= 128; recU = nmgp_lq_pad_records(Sigma_U[packed pairs]).
// bwd == 0: q, m [ns,B,D] (=, entries of pairs a row does not consume must be pre-zeroed by the caller);
// bwd != 0: Pabar, Pbbar [ns,B,Q] (=).
NMGP_API int nmgp_lq_coef_quadform(int bwd, const double* Pa, const double* Pb, const int* I, const double* recU,
                                   const double* Mu, double* q, double* m, const double* qbar, const double* mbar,
                                   double* Pabar, double* Pbbar, int ns, long long B, int Q, int D, cudaStream_t st) {
    NMGP_REQUIRE(ns >= 0 && ns <= 65535 && B >= 0 && Q > 64 && Q <= 128 && D > 0, "nmgp_lq_coef_quadform");
    if (ns == 0 || B == 0) return 0;
    const int mode = bwd ? LQ_UB : LQ_UF;
    if (int r = lq_coef(mode, 0, Pa, I, recU, Mu, q, m, qbar, mbar, Pabar, ns, B, Q, D, st)) return r;
    return lq_coef(mode, 1, Pb, I, recU, Mu, q, m, qbar, mbar, Pbbar, ns, B, Q, D, st);
}

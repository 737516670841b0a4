// gpvar.cu -- variance of the GP-surrogate expected log joint (fp64 only).
//
// Replaces the K^2/2 pairs of triangular solves of pyvbmc/vbmc/variational_optimization.py:1472-1518
// (reference; 2.5 s per call on the CPU at D=20, N=400, K=50, S=8).  Per hyper-sample s:
//   Z      [K][N]   z_kn = exp(lnnf_k - 1/2 |delta_kn|^2)                       (written by gplj_kernel)
//   L_chol:  V = L^-T Z^T  (blocked forward substitution, L upper triangular)   -> z_k'(L'L)^-1 z_j = v_k . v_j
//            J_jk = prior_jk - v_j . v_k / sn2_eff                                (:1486-1501)
//   else  :  W = L Z^T     (dense product, L = -(K + sn2 I)^-1)
//            J_jk = prior_jk + z_k . w_j                                          (:1502-1503)
//   prior_jk = exp(lnnf_jk - 1/2 sum_d ((mu_dj - mu_dk) / tau_jk,d)^2),  tau_jk^2 = (sigma_j^2 + sigma_k^2) lambda^2 + ell^2
//   varG_s  = sum_k w_k^2 max(eps, J_kk) + 2 sum_{j<k} w_j w_k J_jk              (:1505-1514)
// This is a catastrophic cancellation (prior ~ explained part), so everything is fp64 and the
// solve keeps the Cholesky factor (error ~ eps * cond(L), not eps * cond(L)^2).
//
// The N x N factor (S * N^2 * 8 bytes: 10 MB at C3) is the one operand of this library whose HBM/L2
// traffic matters; each CTA streams its L_s panel once per 32-row block with coalesced 256-byte rows.
#include <stdlib.h>

#include "common.cuh"

namespace vbmc {
namespace {

constexpr int kNB = 32;  // rows per block of the substitution
constexpr int kCT = 8;   // right-hand sides (mixture components) per CTA

// grid (ceil(K / kCT), S), block (32, kCT): thread (r, c) owns row r of the current block, column c.
// V is stored [S][N][K] (k contiguous) for the Gram kernel.
__global__ void __launch_bounds__(kNB *kCT)
trsm_kernel(const double *__restrict__ Lall, const double *__restrict__ hyp, int hs, int DP, int N, int K,
            const double *__restrict__ Z, double *__restrict__ V) {
    extern __shared__ double sv[];  // [N][kCT] solved rows of this column tile
    __shared__ double sdiag[kNB][kNB + 1];
    const int s = blockIdx.y, c0 = blockIdx.x * kCT;
    const int r = threadIdx.x, cc = threadIdx.y, col = c0 + cc;
    const bool chol = hyp[(size_t)s * hs + 3 * DP + 4] != 0.0;
    const double *L = Lall + (size_t)s * N * N;
    const double *z = Z + ((size_t)s * K + (col < K ? col : 0)) * N;
    const int nblk = (N + kNB - 1) / kNB;
    for (int b = 0; b < nblk; ++b) {
        const int row = b * kNB + r;
        const bool live = row < N && col < K;
        double acc = 0.0;
        if (chol) {
            // acc = z_row - sum_{c < b*NB} L[c][row] v_c        (L^T lower triangular; L[c][row] contiguous in row)
            if (live) {
                acc = z[row];
                const int cend = b * kNB;
#pragma unroll 8
                for (int c = 0; c < cend; ++c) acc = fma(-L[(size_t)c * N + row], sv[c * kCT + cc], acc);
            }
            // diagonal block: stage L[b*NB + i][b*NB + j] (i <= j) and substitute row by row
            if (cc == 0 || kCT == 1) {
                for (int i = 0; i < kNB; ++i) {
                    const int gi = b * kNB + i;
                    sdiag[i][r] = (gi < N && row < N) ? L[(size_t)gi * N + row] : (i == r ? 1.0 : 0.0);
                }
            }
            __syncthreads();
            for (int t = 0; t < kNB; ++t) {
                // v_t = acc_t / L_tt ; rows > t of this block subtract L[t][row] v_t
                if (r == t) sv[(b * kNB + t) * kCT + cc] = live ? acc / sdiag[t][t] : 0.0;
                __syncthreads();
                if (r > t) acc = fma(-sdiag[t][r], sv[(b * kNB + t) * kCT + cc], acc);
            }
        } else {
            // W = L z  (dense; L is symmetric, read it column-wise so that rows are contiguous)
            if (live) {
#pragma unroll 8
                for (int c = 0; c < N; ++c) acc = fma(L[(size_t)c * N + row], z[c], acc);
            }
            sv[(b * kNB + r) * kCT + cc] = acc;
            __syncthreads();
        }
        if (live) V[((size_t)s * N + row) * K + col] = sv[row * kCT + cc];
    }
}

// grid (K, S): CTA (k, s) fills row k of J_s.  threads over j.
__global__ void __launch_bounds__(128)
gram_kernel(const double *__restrict__ prm, ParamLayout lay, const double *__restrict__ hyp, int hs, int N,
            const double *__restrict__ Z, const double *__restrict__ V, double *__restrict__ J) {
    const int D = lay.D, DP = lay.DP, K = lay.K, k = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const double *h = hyp + (size_t)s * hs;
    const bool chol = h[3 * DP + 4] != 0.0;
    const double sn2 = h[3 * DP + 3], ln_sf2 = h[3 * DP + 0], sum_lnell = h[3 * DP + 1];
    const double *mu = prm + lay.mu(), *sigma = prm + lay.sigma(), *lambd = prm + lay.lambd();
    const double *Vs = V + (size_t)s * N * K, *Zs = Z + (size_t)s * K * N;
    for (int j = tid; j < K; j += nt) {
        // prior term (:1479-1488)
        const double s2 = sigma[j] * sigma[j] + sigma[k] * sigma[k];
        double lnt = 0.0, d2 = 0.0;
        for (int d = 0; d < D; ++d) {
            const double t2 = s2 * lambd[d] * lambd[d] + h[d] * h[d];
            const double dm = mu[j * D + d] - mu[k * D + d];
            lnt += log(t2);
            d2 += dm * dm / t2;
        }
        double v = exp(ln_sf2 + sum_lnell - 0.5 * lnt - 0.5 * d2);
        double q = 0.0;
        if (chol) {
#pragma unroll 4
            for (int n = 0; n < N; ++n) q = fma(Vs[(size_t)n * K + k], Vs[(size_t)n * K + j], q);
            v -= q / sn2;
        } else {
#pragma unroll 4
            for (int n = 0; n < N; ++n) q = fma(Zs[(size_t)k * N + n], Vs[(size_t)n * K + j], q);
            v += q;
        }
        J[((size_t)s * K + k) * K + j] = v;
    }
}

// ---- round 2: V = M^T Z^T with the explicit factor inverse ---------------------------------------------------------
// The blocked substitution above is inherently sequential (13 dependent row blocks at N = 400: 449 us at C3).  With
// M = L^-1 (upper triangular; built once per packed GP by gppred.cu, shared with the GP predictions) the solve is a
// product, V[n][k] = sum_{c <= n} M[c][n] z_k[c], in which every (n, k) entry is independent: grid (ceil(N / 16), S),
// block (16 points n) x (16 component lanes, 4 components each); the z chunk of all K components is staged in shared
// memory (double-buffered), the 32 factor entries of a chunk are loaded up front (coalesced along n).  Low-noise
// samples use M = L (dense) and the full range of c.
__global__ void __launch_bounds__(256)
vmul_kernel(const double *__restrict__ Mall, const double *__restrict__ hyp, int hs, int DP, int N, int NP, int K,
            const double *__restrict__ Z, double *__restrict__ V) {
    extern __shared__ double sZ[];  // [2][64][33]
    constexpr int NB = 16;          // points per CTA; block = 16 points x 16 component lanes, 4 components per lane
    const int s = blockIdx.y, n0 = blockIdx.x * NB, r = threadIdx.x & (NB - 1), q = threadIdx.x / NB, n = n0 + r;
    const bool chol = hyp[(size_t)s * hs + 3 * DP + 4] != 0.0;
    const double *M = Mall + (size_t)s * N * NP;
    const double *Zs = Z + (size_t)s * K * N;
    const int c_end = chol ? min(N, n0 + NB) : N;
    const int nn = n < N ? n : N - 1;  // (clamped: out-of-range lanes load valid memory and are not stored)
    for (int k0 = 0; k0 < K; k0 += 64) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        auto stage = [&](int c0, int buf) {
            double *dst = sZ + buf * 64 * 33;
            for (int e = threadIdx.x; e < 64 * 32; e += 256) {
                const int kk = e >> 5, cc = e & 31;
                dst[kk * 33 + cc] = (k0 + kk < K && c0 + cc < N) ? Zs[(size_t)(k0 + kk) * N + c0 + cc] : 0.0;
            }
        };
        __syncthreads();
        stage(0, 0);
        int buf = 0;
        for (int c0 = 0; c0 < c_end; c0 += 32, buf ^= 1) {
            // the 32 factor entries of this chunk: independent loads, all in flight before the first use
            double m[32];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) m[cc] = (c0 + cc < N) ? M[(size_t)(c0 + cc) * NP + nn] : 0.0;
            __syncthreads();  // chunk c0 staged (and every thread is done reading the other buffer)
            if (c0 + 32 < c_end) stage(c0 + 32, buf ^ 1);
            const double *z = sZ + buf * 64 * 33;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fma(m[cc], z[(q + 16 * i) * 33 + cc], acc[i]);
            }
        }
        if (n < N) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + q + 16 * i;
                if (k < K) V[((size_t)s * N + n) * K + k] = acc[i];
            }
        }
    }
}

// grid (K, S), block (64 components j) x (4 quarters of the training points): CTA (k, s) fills row k of J_s; the four
// quarters (and the four slices of the dimension sum of the prior term) meet in shared memory.
__global__ void __launch_bounds__(256)
gram4_kernel(const double *__restrict__ prm, ParamLayout lay, const double *__restrict__ hyp, int hs, int N,
             const double *__restrict__ Z, const double *__restrict__ V, double *__restrict__ J) {
    const int D = lay.D, DP = lay.DP, K = lay.K, k = blockIdx.x, s = blockIdx.y, tx = threadIdx.x & 63, part = threadIdx.x >> 6;
    __shared__ double sq[4][64], sl[4][64], sd[4][64];
    const double *h = hyp + (size_t)s * hs;
    const bool chol = h[3 * DP + 4] != 0.0;
    const double sn2 = h[3 * DP + 3], ln_sf2 = h[3 * DP + 0], sum_lnell = h[3 * DP + 1];
    const double *mu = prm + lay.mu(), *sigma = prm + lay.sigma(), *lambd = prm + lay.lambd();
    const double *Vs = V + (size_t)s * N * K, *Zs = Z + (size_t)s * K * N;
    const int nq = (N + 3) / 4, nb = part * nq, ne = min(N, nb + nq);
    for (int j0 = 0; j0 < K; j0 += 64) {
        const int j = j0 + tx;
        double q0 = 0.0, q1 = 0.0, lnt = 0.0, d2 = 0.0;
        if (j < K) {
            // prior term (:1479-1488), dimensions d = part, part + 4, ...
            const double s2 = sigma[j] * sigma[j] + sigma[k] * sigma[k];
            for (int d = part; d < D; d += 4) {
                const double t2 = s2 * lambd[d] * lambd[d] + h[d] * h[d];
                const double dm = mu[j * D + d] - mu[k * D + d];
                lnt += log(t2);
                d2 += dm * dm / t2;
            }
            if (chol) {
                int n = nb;
                for (; n + 1 < ne; n += 2) {
                    q0 = fma(Vs[(size_t)n * K + k], Vs[(size_t)n * K + j], q0);
                    q1 = fma(Vs[(size_t)(n + 1) * K + k], Vs[(size_t)(n + 1) * K + j], q1);
                }
                if (n < ne) q0 = fma(Vs[(size_t)n * K + k], Vs[(size_t)n * K + j], q0);
            } else {
                for (int n = nb; n < ne; ++n) q0 = fma(Zs[(size_t)k * N + n], Vs[(size_t)n * K + j], q0);
            }
        }
        __syncthreads();
        sq[part][tx] = q0 + q1, sl[part][tx] = lnt, sd[part][tx] = d2;
        __syncthreads();
        if (part == 0 && j < K) {
            const double q = (sq[0][tx] + sq[1][tx]) + (sq[2][tx] + sq[3][tx]);
            const double L = (sl[0][tx] + sl[1][tx]) + (sl[2][tx] + sl[3][tx]);
            const double Dd = (sd[0][tx] + sd[1][tx]) + (sd[2][tx] + sd[3][tx]);
            double v = exp(ln_sf2 + sum_lnell - 0.5 * L - 0.5 * Dd);
            v = chol ? v - q / sn2 : v + q;
            J[((size_t)s * K + k) * K + j] = v;
        }
    }
}

// grid S: symmetrise from the (j <= k) evaluations like the reference, varG_s, and (CTA 0, after a
// ticket) the across-sample statistics (:1578-1587).  out_var = [varG, var_ss, varG_s (S)...]
__global__ void __launch_bounds__(256)
var_reduce_kernel(const double *__restrict__ prm, ParamLayout lay, int S, double *__restrict__ J,
                  const double *__restrict__ gps, int gps_stride, int avg, double *__restrict__ out_var,
                  unsigned int *__restrict__ ticket) {
    const int K = lay.K, s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    __shared__ double scratch[40];
    __shared__ bool last;
    const double *w = prm + lay.w();
    double *Js = J + (size_t)s * K * K;
    const double eps = 2.220446049250313e-16;
    double acc = 0.0;
    for (int e = tid; e < K * K; e += nt) {
        const int k = e / K, j = e - k * K;
        if (j < k)
            acc += 2.0 * w[j] * w[k] * Js[e];  // reference keeps J[k][j], j < k (:1510-1514)
        else if (j == k)
            acc += w[k] * w[k] * fmax(eps, Js[e]);  // :1506-1507
    }
    acc = block_sum(acc, scratch);
    __syncthreads();
    for (int e = tid; e < K * K; e += nt) {
        const int k = e / K, j = e - k * K;
        if (j > k) Js[e] = Js[(size_t)j * K + k];  // mirror the lower triangle
    }
    if (tid == 0) {
        out_var[2 + s] = fmax(acc, eps);  // :1517-1518
        __threadfence();
        last = atomicAdd(ticket, 1u) == (unsigned)(S - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (tid == 0) {
        *ticket = 0;
        double varG = out_var[2], var_ss = 0.0;
        if (S > 1 && avg) {
            double Gb = 0.0, vb = 0.0;
            for (int i = 0; i < S; ++i) Gb += gps[(size_t)i * gps_stride], vb += out_var[2 + i];
            Gb /= S;
            const double vmean = vb / S;
            double ss = 0.0, sv = 0.0;
            for (int i = 0; i < S; ++i) {
                const double dg = gps[(size_t)i * gps_stride] - Gb, dv = out_var[2 + i] - vmean;
                ss += dg * dg;
                sv += dv * dv;
            }
            const double varG_ss = ss / (S - 1);
            var_ss = varG_ss + sqrt(sv / (S - 1));  // sic: a std added to a variance (:1586)
            varG = vb / S + varG_ss;                // :1587
        }
        out_var[0] = varG;
        out_var[1] = var_ss;
    }
}

}  // namespace

// Workspace layout in c->d_var: Z [S K N] | V [S N K] | J [S K K] | out_var [2 + S] | ticket
size_t gpvar_workspace(int S, int K, int N) { return (size_t)2 * S * K * N + (size_t)S * K * K + 2 + S + 2; }
double *gpvar_Z(Ctx *c) { return c->d_var; }
double *gpvar_J(Ctx *c, int K) { return c->d_var + (size_t)2 * c->S * K * c->N; }
double *gpvar_out(Ctx *c, int K) { return gpvar_J(c, K) + (size_t)c->S * K * K; }

int gpvar_launch(Ctx *c, const double *d_params, int K, int avg) {
    VBMC_REQUIRE(c->has_gp && c->has_L, VBMC_ERR_STATE, "log-joint variance needs the GP factor L (pack the GP with L)");
    const int D = c->gD, DP = c->gDP, N = c->N, S = c->S, hs = hyp_stride(DP);
    ParamLayout lay{D, DP, K};
    double *Z = gpvar_Z(c), *V = Z + (size_t)S * K * N, *J = gpvar_J(c, K), *ov = gpvar_out(c, K);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(ov + 2 + S);
    const size_t smem = (size_t)(((N + kNB - 1) / kNB) * kNB) * kCT * sizeof(double);
    VBMC_REQUIRE(smem <= 200 * 1024, VBMC_ERR_UNSUPPORTED, "log-joint variance: N too large for the substitution kernel");
    if (smem > 48 * 1024)
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VBMC_CUDA_CHECK(cudaMemsetAsync(ticket, 0, 2 * sizeof(unsigned int), c->stream));
    static int env_old = -1;
    if (env_old < 0) {
        const char *e = getenv("VBMC_GPVAR_OLD");
        env_old = e ? atoi(e) : 0;
    }
    const int NP = gppred_np(N);
    if (!env_old && gppred_prepare(c) == VBMC_OK) {
        // product with the explicit inverse (default)
        vmul_kernel<<<dim3((N + 15) / 16, S), 256, 2 * 64 * 33 * sizeof(double), c->stream>>>(c->d_Linv, c->d_hyp, hs, DP, N, NP, K, Z, V);
        VBMC_CUDA_CHECK(cudaGetLastError());
        gram4_kernel<<<dim3(K, S), 256, 0, c->stream>>>(d_params, lay, c->d_hyp, hs, N, Z, V, J);
        VBMC_CUDA_CHECK(cudaGetLastError());
    } else {
        cudaGetLastError();
        // fallback (N beyond the prediction kernel's limit): blocked substitution
        trsm_kernel<<<dim3((K + kCT - 1) / kCT, S), dim3(kNB, kCT), smem, c->stream>>>(c->d_L, c->d_hyp, hs, DP, N, K, Z, V);
        VBMC_CUDA_CHECK(cudaGetLastError());
        gram_kernel<<<dim3(K, S), K > 64 ? 128 : 64, 0, c->stream>>>(d_params, lay, c->d_hyp, hs, N, Z, V, J);
        VBMC_CUDA_CHECK(cudaGetLastError());
    }
    var_reduce_kernel<<<S, 256, 0, c->stream>>>(d_params, lay, S, J, c->d_gps, 1 + RawLayout{D, K}.block(), avg, ov, ticket);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches += 3;
    return VBMC_OK;
}

}  // namespace vbmc

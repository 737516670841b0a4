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
    trsm_kernel<<<dim3((K + kCT - 1) / kCT, S), dim3(kNB, kCT), smem, c->stream>>>(c->d_L, c->d_hyp, hs, DP, N, K, Z, V);
    VBMC_CUDA_CHECK(cudaGetLastError());
    gram_kernel<<<dim3(K, S), K > 64 ? 128 : 64, 0, c->stream>>>(d_params, lay, c->d_hyp, hs, N, Z, V, J);
    VBMC_CUDA_CHECK(cudaGetLastError());
    var_reduce_kernel<<<S, 256, 0, c->stream>>>(d_params, lay, S, J, c->d_gps, 1 + RawLayout{D, K}.block(), avg, ov, ticket);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches += 3;
    return VBMC_OK;
}

}  // namespace vbmc

// gppred.cu -- GP posterior predictive mean / variance at a batch of search points, and the variational-posterior
// density at the same points: the two O(Nx) ingredients of every acquisition function
// (pyvbmc/acquisition_functions/abstract_acq_fcn.py:76-97: `f_mu, f_s2 = gp.predict(Xs, separate_samples=True)`;
//  acq_fcn_log.py:38-47: `vp.pdf(Xs, orig_flag=False, log_flag=True)`).  fp64 only.
//
// gp.predict is gpyreg's (third-party, not vendored); per hyper-sample s it evaluates (gplite / gpyreg
// parameterisation, the same posterior record `_gp_log_joint` consumes, variational_optimization.py:1375-1398)
//     k*_n    = sf2 exp(-1/2 sum_d ((x*_d - X_nd) / ell_d)^2)                       cross-covariance, SE-ARD
//     f_mu    = m(x*) + k* . alpha
//     f_s2    = sf2 - |L^-T (sW k*)|^2        (L_chol: L upper Cholesky factor of K / sn2 + I,  sW = 1 / sqrt(sn2))
//             = sf2 + k* . (L k*)             (low-noise branch: L = -(K + sn2 I)^-1)
//     f_s2    = max(f_s2, 0)
// The variance is a cancellation (prior minus explained part), so the triangular structure is kept: with
// M = L^-1 (upper triangular, built once per packed GP by tri_inv_kernel), |L^-T k*|^2 = |M^T k*|^2 is a sum of
// squares of the rows of  W = K* M  -- a dense [Nx x N] x [N x N] product, 2 N^2 Nx flops per hyper-sample
// (1.3 GFLOP at N = 400, Nx = 8192), the one GEMM-shaped fp64 contraction of this library.
//
// gppred_kernel, one CTA per (32 search points, hyper-sample):
//   * the product runs on the fp64 tensor-core path: mma.sync.m8n8k4.f64 (DMMA); the 32 x N accumulator block W
//     stays in registers for the whole reduction (8 warps x 4 row tiles x 7 column tiles = 56 doubles per thread),
//     column tiles are dealt round-robin to the warps so the triangular skip (M[n][m] = 0 for m < n) stays balanced;
//   * the reduction index n is consumed in chunks of 8 training points: the K* chunk is computed on the fly with
//     direct differences (one element per thread, never stored to HBM) and the matching 8 rows of M are staged with
//     cp.async, both double-buffered in shared memory; one __syncthreads per chunk;
//   * f_mu falls out of the K* pass (per-thread partial dot products with alpha, 8-lane shuffle reduction);
//   * per-point variance: squares of the accumulators, shuffle reduction over the 4 lanes that share a row, then a
//     fixed-order sum over the 8 warps.
// vp_pdf_kernel: one thread per point, the direct form of variational_posterior.py:447-468 (transformed space).
#include "common.cuh"

namespace vbmc {
namespace {

constexpr int kTP = 32;       // search points per CTA
constexpr int kKC = 8;        // training points per reduction chunk (two k4 steps)
constexpr int kPThreads = 256;
constexpr int kColTiles = 7;  // column tiles (8 columns each) per warp: 8 warps x 7 x 8 = 448 columns
constexpr int kMaxNP = 8 * kColTiles * 8;
constexpr int kAStride = kKC + 4;  // sK row stride (doubles): conflict-free A-fragment loads

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// M[s] = L[s]^-1 (upper triangular) for Cholesky samples, M[s] = L[s] (dense, symmetric) otherwise; [S][N][NP],
// columns >= N are zero.  grid (ceil(N / 32), S), 32 threads: thread j solves L x = e_j by back substitution; the 32
// solutions of the CTA live in shared memory [N][33].  Every lane reads the same L[i][k] (broadcast).
__global__ void __launch_bounds__(32)
tri_inv_kernel(const double *__restrict__ Lall, const double *__restrict__ hyp, int hs, int DP, int N, int NP,
               double *__restrict__ Mall) {
    extern __shared__ double sx[];  // [N][33]
    const int s = blockIdx.y, j0 = blockIdx.x * 32, lane = threadIdx.x, j = j0 + lane;
    const double *L = Lall + (size_t)s * N * N;
    double *M = Mall + (size_t)s * N * NP;
    const bool chol = hyp[(size_t)s * hs + 3 * DP + 4] != 0.0;
    if (!chol) {
        for (int i = 0; i < N; ++i)
            if (j < NP) M[(size_t)i * NP + j] = j < N ? L[(size_t)i * N + j] : 0.0;
        if (blockIdx.x == gridDim.x - 1)
            for (int i = 0; i < N; ++i)
                for (int c = j0 + 32 + lane; c < NP; c += 32) M[(size_t)i * NP + c] = 0.0;
        return;
    }
    const int jmax = min(j0 + 31, N - 1);
    for (int i = N - 1; i > jmax; --i) sx[i * 33 + lane] = 0.0;
    for (int i = jmax; i >= 0; --i) {
        const double *Li = L + (size_t)i * N;
        double a0 = (i == j) ? 1.0 : 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = i + 1;
        for (; k + 3 <= jmax; k += 4) {
            a0 = fma(-Li[k], sx[k * 33 + lane], a0);
            a1 = fma(-Li[k + 1], sx[(k + 1) * 33 + lane], a1);
            a2 = fma(-Li[k + 2], sx[(k + 2) * 33 + lane], a2);
            a3 = fma(-Li[k + 3], sx[(k + 3) * 33 + lane], a3);
        }
        for (; k <= jmax; ++k) a0 = fma(-Li[k], sx[k * 33 + lane], a0);
        const double x = (i <= j && j < N) ? ((a0 + a1) + (a2 + a3)) / Li[i] : 0.0;
        sx[i * 33 + lane] = x;  // (own column only: no cross-lane dependency, no barrier needed)
    }
    for (int i = 0; i < N; ++i)
        if (j < NP) M[(size_t)i * NP + j] = (j < N) ? sx[i * 33 + lane] : 0.0;
    if (blockIdx.x == gridDim.x - 1)
        for (int i = 0; i < N; ++i)
            for (int c = j0 + 32 + lane; c < NP; c += 32) M[(size_t)i * NP + c] = 0.0;
}

template <int DP>
__global__ void __launch_bounds__(kPThreads, 1)
gppred_kernel(const double *__restrict__ Xs, int Nx, int D, const double *__restrict__ Xts_all, int N, int NP,
              const double *__restrict__ hyp, int hs, const double *__restrict__ alpha, const double *__restrict__ Mall,
              int S, int mean_kind, double *__restrict__ f_mu, double *__restrict__ f_s2) {
    extern __shared__ __align__(16) double psm[];
    const int s = blockIdx.y, p0 = blockIdx.x * kTP, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int MS = NP + 8;                     // sM row stride (doubles): NP % 32 == 0  =>  stride = 8 (mod 32)
    double *sXs = psm;                         // [kTP][DP]  search points / ell
    double *sK = sXs + kTP * DP;               // [2][kTP][kAStride]
    double *sM = sK + 2 * kTP * kAStride;      // [2][kKC][MS]
    double *sR = sM + 2 * kKC * MS;            // [8 warps][kTP]
    const double *h = hyp + (size_t)s * hs;
    const double *M = Mall + (size_t)s * N * NP;
    const double *al = alpha + (size_t)s * N;
    const double *Xts = Xts_all + (size_t)s * DP * N;  // [DP][N] training inputs / ell_s (padded rows are zero)
    const bool chol = h[3 * DP + 4] != 0.0;
    const double ln_sf2 = h[3 * DP + 0], sn2 = h[3 * DP + 3];

    for (int e = tid; e < kTP * DP; e += kPThreads) {
        const int i = e / DP, d = e - i * DP, p = p0 + i;
        sXs[e] = (p < Nx && d < D) ? Xs[(size_t)p * D + d] / h[d] : 0.0;
    }
    __syncthreads();

    // K* element of this thread in every chunk: point ki, training point n0 + kn
    const int ki = tid >> 3, kn = tid & 7;
    const double *xi = sXs + ki * DP;
    double mu_acc = 0.0;

    auto stage = [&](int c) {  // chunk c -> buffers (c & 1)
        const int n0 = c * kKC, n = n0 + kn, b = c & 1;
        double v = 0.0;
        if (n < N) {
            double d2 = 0.0;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                const double t = xi[d] - Xts[(size_t)d * N + n];  // (padded dimensions: 0 - 0)
                d2 = fma(t, t, d2);
            }
            v = exp(ln_sf2 - 0.5 * d2);
            mu_acc = fma(v, al[n], mu_acc);
        }
        sK[(b * kTP + ki) * kAStride + kn] = v;
        // rows n0 .. n0+7 of M; Cholesky samples: columns below n0 are zero and never read
        const int c_lo = chol ? (n0 & ~7) : 0;
        const int per_row = (NP - c_lo) >> 1;  // 16-byte pieces per row
        for (int e = tid; e < kKC * per_row; e += kPThreads) {
            const int r = e / per_row, q = e - r * per_row;
            double *dst = sM + (size_t)(b * kKC + r) * MS + c_lo + 2 * q;
            if (n0 + r < N)
                cp_async16(dst, M + (size_t)(n0 + r) * NP + c_lo + 2 * q);
            else
                dst[0] = 0.0, dst[1] = 0.0;
        }
        cp_async_commit();
    };

    double acc[4][kColTiles][2];
#pragma unroll
    for (int rt = 0; rt < 4; ++rt)
#pragma unroll
        for (int ct = 0; ct < kColTiles; ++ct) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;

    const int nchunk = (N + kKC - 1) / kKC;
    stage(0);
    cp_async_wait_all();
    __syncthreads();
    const int ar = lane >> 2, ak = lane & 3;  // A fragment: row (lane / 4), k (lane % 4);  B: k (lane % 4), col (lane / 4)
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) stage(c + 1);
        const int b = c & 1, n0 = c * kKC;
        const double *aK = sK + (size_t)b * kTP * kAStride, *bM = sM + (size_t)b * kKC * MS;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            double a[4];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) a[rt] = aK[(8 * rt + ar) * kAStride + 4 * ks + ak];
#pragma unroll
            for (int ct = 0; ct < kColTiles; ++ct) {
                const int col0 = 8 * (wid + 8 * ct);
                if (col0 >= NP || (chol && col0 + 8 <= n0)) continue;  // warp-uniform
                const double bb = bM[(4 * ks + ak) * MS + col0 + ar];
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) dmma(acc[rt][ct][0], acc[rt][ct][1], a[rt], bb);
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }

    // f_mu: 8 lanes (kn) share a point
    mu_acc += __shfl_xor_sync(0xffffffffu, mu_acc, 1);
    mu_acc += __shfl_xor_sync(0xffffffffu, mu_acc, 2);
    mu_acc += __shfl_xor_sync(0xffffffffu, mu_acc, 4);
    if (kn == 0 && p0 + ki < Nx) {
        double m = 0.0;
        if (mean_kind != VBMC_MEAN_ZERO) m = h[3 * DP + 2];
        if (mean_kind == VBMC_MEAN_NEGQUAD) {
            double q = 0.0;
            for (int d = 0; d < D; ++d) {
                const double t = xi[d] * h[d] - h[DP + d];  // x - x_m
                q = fma(t * t, h[2 * DP + d], q);
            }
            m -= 0.5 * q;
        }
        f_mu[(size_t)(p0 + ki) * S + s] = m + mu_acc;
    }

    // per-point variance.  accumulator (rt, ct, e) is W[row = 8 rt + lane / 4][col = 8 (wid + 8 ct) + 2 (lane % 4) + e]
    double r[4];
#pragma unroll
    for (int rt = 0; rt < 4; ++rt) {
        double v = 0.0;
#pragma unroll
        for (int ct = 0; ct < kColTiles; ++ct) {
            const int col0 = 8 * (wid + 8 * ct) + 2 * ak;
            if (chol) {
                v = fma(acc[rt][ct][0], acc[rt][ct][0], v);
                v = fma(acc[rt][ct][1], acc[rt][ct][1], v);
            } else {
                // low-noise branch: k* . (L k*) needs K* again at the accumulator's column (rare path: recomputed)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m_ = col0 + e;
                    if (m_ < N) {
                        double d2 = 0.0;
                        for (int d = 0; d < D; ++d) {
                            const double t = sXs[(8 * rt + ar) * DP + d] - Xts[(size_t)d * N + m_];
                            d2 = fma(t, t, d2);
                        }
                        v = fma(acc[rt][ct][e], exp(ln_sf2 - 0.5 * d2), v);
                    }
                }
            }
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        r[rt] = v;
    }
    if (ak == 0) {
#pragma unroll
        for (int rt = 0; rt < 4; ++rt) sR[wid * kTP + 8 * rt + ar] = r[rt];
    }
    __syncthreads();
    if (tid < kTP && p0 + tid < Nx) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += sR[w * kTP + tid];  // fixed order
        const double sf2 = exp(ln_sf2);
        const double fs2 = chol ? sf2 - v / sn2 : sf2 + v;
        f_s2[(size_t)(p0 + tid) * S + s] = fmax(fs2, 0.0);
    }
}

// y_i = sum_k w_k N(x_i; mu_k, sigma_k^2 Lambda)  (variational_posterior.py:447-468), optionally its logarithm and
// the gradient w.r.t. x (of the pdf, or of the log pdf when log_flag).  One thread per point.
__global__ void __launch_bounds__(128)
vp_pdf_kernel(const double *__restrict__ prm, ParamLayout lay, const double *__restrict__ Xs, int Nx, int log_flag,
              int grad_flag, double *__restrict__ y, double *__restrict__ dy) {
    extern __shared__ double sp[];  // mu [K][D] | sigma [K] | lambda [D] | w [K]
    const int D = lay.D, K = lay.K, tid = threadIdx.x;
    for (int e = tid; e < K * D + 2 * K + D; e += blockDim.x) sp[e] = prm[e];  // (the four leading blocks of ParamLayout)
    __syncthreads();
    const double *mu = sp + lay.mu(), *sigma = sp + lay.sigma(), *lambd = sp + lay.lambd(), *w = sp + lay.w();
    const int i = blockIdx.x * blockDim.x + tid;
    if (i >= Nx) return;
    double x[kMaxD], g[kMaxD];
    double prod_l = 1.0;
    for (int d = 0; d < D; ++d) x[d] = Xs[(size_t)i * D + d], g[d] = 0.0, prod_l *= lambd[d];
    const double nf = 1.0 / pow(2.0 * 3.14159265358979323846, 0.5 * D) / prod_l;
    double yy = 0.0;
    for (int k = 0; k < K; ++k) {
        const double sk = sigma[k];
        double d2 = 0.0;
        for (int d = 0; d < D; ++d) {
            const double t = (x[d] - mu[k * D + d]) / (sk * lambd[d]);
            d2 = fma(t, t, d2);
        }
        const double nn = nf * w[k] / pow(sk, (double)D) * exp(-0.5 * d2);
        yy += nn;
        if (grad_flag)
            for (int d = 0; d < D; ++d) g[d] -= nn * (x[d] - mu[k * D + d]) / (lambd[d] * lambd[d] * sk * sk);
    }
    if (grad_flag)
        for (int d = 0; d < D; ++d) dy[(size_t)i * D + d] = log_flag ? g[d] / yy : g[d];
    y[i] = log_flag ? (yy == 0.0 ? -INFINITY : log(yy)) : yy;
}

template <int DP>
int gppred_launch_dp(Ctx *c, const double *d_Xs, int Nx, int NP, double *d_mu, double *d_s2) {
    const int N = c->N, S = c->S, hs = hyp_stride(DP);
    const size_t smem = sizeof(double) * ((size_t)kTP * DP + 2 * kTP * kAStride + 2 * (size_t)kKC * (NP + 8) + 8 * kTP);
    static size_t smem_set = 0;
    if (smem > smem_set && smem > 48 * 1024) {
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(gppred_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    gppred_kernel<DP><<<dim3((Nx + kTP - 1) / kTP, S), kPThreads, smem, c->stream>>>(
        d_Xs, Nx, c->gD, c->d_Linv + (size_t)S * N * NP, N, NP, c->d_hyp, hs, c->d_alpha, c->d_Linv, S, c->mean_kind, d_mu, d_s2);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace

int gppred_np(int N) { return (N + 31) / 32 * 32; }

namespace {
// Xts[s][d][n] = X[n][d] / ell_{s,d}
__global__ void scale_x_kernel(const double *__restrict__ Xt, const double *__restrict__ hyp, int hs, int DP, int N, int S,
                               double *__restrict__ Xts) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * DP * N) return;
    const int s = e / (DP * N), r = e - s * DP * N, d = r / N;
    Xts[e] = Xt[r] / hyp[(size_t)s * hs + d];
}
}  // namespace

// L^-1 (or L itself on the low-noise branch) of every hyper-sample and the per-sample scaled training inputs, once per
// packed GP
int gppred_prepare(Ctx *c) {
    VBMC_REQUIRE(c->has_gp && c->has_L, VBMC_ERR_STATE, "GP prediction needs the factor L (pack the GP with L)");
    if (c->d_Linv) return VBMC_OK;
    const int N = c->N, S = c->S, DP = c->gDP, NP = gppred_np(N);
    VBMC_REQUIRE(NP <= kMaxNP, VBMC_ERR_UNSUPPORTED, "GP prediction: more than 448 training points are not supported yet");
    VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_Linv, ((size_t)S * N * NP + (size_t)S * DP * N) * sizeof(double)));
    scale_x_kernel<<<(S * DP * N + 255) / 256, 256, 0, c->stream>>>(c->d_Xt, c->d_hyp, hyp_stride(DP), DP, N, S,
                                                                     c->d_Linv + (size_t)S * N * NP);
    VBMC_CUDA_CHECK(cudaGetLastError());
    const size_t smem = (size_t)N * 33 * sizeof(double);
    if (smem > 48 * 1024)
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(tri_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tri_inv_kernel<<<dim3((NP + 31) / 32, S), 32, smem, c->stream>>>(c->d_L, c->d_hyp, hyp_stride(DP), DP, N, NP, c->d_Linv);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches += 2;
    return VBMC_OK;
}

int gppred_launch(Ctx *c, const double *d_Xs, int Nx, double *d_mu, double *d_s2) {
    VBMC_TRY(gppred_prepare(c));
    const int NP = gppred_np(c->N);
    switch (c->gDP) {
#define VBMC_CASE(DPV) \
    case DPV:          \
        return gppred_launch_dp<DPV>(c, d_Xs, Nx, NP, d_mu, d_s2)
        VBMC_CASE(4);
        VBMC_CASE(8);
        VBMC_CASE(12);
        VBMC_CASE(16);
        VBMC_CASE(20);
        VBMC_CASE(24);
        VBMC_CASE(28);
        VBMC_CASE(32);
#undef VBMC_CASE
    }
    set_error("gp_predict: unsupported padded dimension");
    return VBMC_ERR_UNSUPPORTED;
}

int vp_pdf_launch(Ctx *c, const double *d_params, int D, int K, const double *d_Xs, int Nx, int log_flag, int grad_flag,
                  double *d_y, double *d_dy) {
    ParamLayout lay{D, pad_dim(D), K};
    const size_t smem = sizeof(double) * ((size_t)K * D + 2 * K + D);
    VBMC_REQUIRE(smem <= 48 * 1024, VBMC_ERR_UNSUPPORTED, "vp_pdf: K * D too large");
    vp_pdf_kernel<<<(Nx + 127) / 128, 128, smem, c->stream>>>(d_params, lay, d_Xs, Nx, log_flag, grad_flag, d_y, d_dy);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace vbmc

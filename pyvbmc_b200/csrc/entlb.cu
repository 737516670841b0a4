// entlb.cu -- Jensen lower bound on the mixture entropy and its gradient (fp64).
//
// Replaces pyvbmc/entropy/entlb_vbmc.py:60-159 (reference).  O(K^2 D) work: three tiny kernels.
//   gamma_ij = N(mu_i; mu_j, (sigma_i^2 + sigma_j^2) Lambda)       (evaluated in the log domain)
//   H        = -sum_i w_i log sum_j w_j gamma_ij                                        (:84-97)
// Gradients are written RAW (before the log / softmax Jacobians, which finalize_kernel applies)
// into the entropy block of the raw vector.
#include "common.cuh"

namespace vbmc {
namespace {

constexpr double kLog2Pi = 1.8378770664093454836;

// A: CTA i -> lg[i][j], r2[i][j], lgs[i] = log sum_j w_j gamma_ij
__global__ void entlb_a_kernel(const double *__restrict__ prm, ParamLayout lay, double *__restrict__ lg,
                               double *__restrict__ r2o, double *__restrict__ lgs) {
    const int D = lay.D, K = lay.K, i = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    __shared__ double scratch[40];
    const double *mu = prm + lay.mu(), *sigma = prm + lay.sigma(), *lambd = prm + lay.lambd(), *w = prm + lay.w();
    double sumlnl = 0.0;
    for (int d = 0; d < D; ++d) sumlnl += log(lambd[d]);
    const double c0 = -0.5 * D * kLog2Pi - sumlnl;
    const double si2 = sigma[i] * sigma[i];
    double mx = -INFINITY;
    for (int j = tid; j < K; j += nt) {
        const double s2 = si2 + sigma[j] * sigma[j];
        double r2 = 0.0;
        for (int d = 0; d < D; ++d) {
            const double t = (mu[i * D + d] - mu[j * D + d]) / lambd[d];
            r2 = fma(t, t, r2);
        }
        const double v = c0 - 0.5 * D * log(s2) - 0.5 * r2 / s2;
        lg[(size_t)i * K + j] = v;
        r2o[(size_t)i * K + j] = r2;
        mx = fmax(mx, v + log(w[j]));
    }
    mx = block_max(mx, scratch);
    double se = 0.0;
    for (int j = tid; j < K; j += nt) se += exp(lg[(size_t)i * K + j] + log(w[j]) - mx);
    se = block_sum(se, scratch);
    if (tid == 0) lgs[i] = mx + log(se);
}

// B: CTA j -> raw gmu[:, j], gsig[j], gw[j], lambda contributions lamp[j][:]
__global__ void entlb_b_kernel(const double *__restrict__ prm, ParamLayout lay, const double *__restrict__ lg,
                               const double *__restrict__ r2i, const double *__restrict__ lgs,
                               double *__restrict__ raw_ent, RawLayout rl, double *__restrict__ lamp) {
    const int D = lay.D, K = lay.K, j = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    extern __shared__ double sm[];
    double *pw = sm;          // pair_ij / s2_ij
    double *gj = sm + K;      // w_i gamma_ij / gammasum_j
    double *is2 = sm + 2 * K; // 1 / s2_ij
    __shared__ double scratch[40];
    const double *mu = prm + lay.mu(), *sigma = prm + lay.sigma(), *lambd = prm + lay.lambd(), *w = prm + lay.w();
    const double sj2 = sigma[j] * sigma[j], lgsj = lgs[j], wj = w[j];
    double acc_sig = 0.0, acc_w = 0.0;
    for (int i = tid; i < K; i += nt) {
        const double l = lg[(size_t)i * K + j];
        const double gam_i = exp(l - lgs[i]), gam_j = exp(l - lgsj);
        const double s2 = sigma[i] * sigma[i] + sj2;
        const double pair = w[i] * wj * (gam_i + gam_j);
        pw[i] = pair / s2;
        gj[i] = w[i] * gam_j;
        is2[i] = 1.0 / s2;
        const double r2 = r2i[(size_t)i * K + j];
        acc_sig += pair * (-D / s2 + r2 / (s2 * s2));  // :112-115,133-139
        acc_w += w[i] * gam_i;                          // :158-159 (gamma symmetric)
    }
    acc_sig = block_sum(acc_sig, scratch);
    acc_w = block_sum(acc_w, scratch);
    if (tid == 0) {
        raw_ent[rl.o_sig() + j] = -sigma[j] * acc_sig;
        raw_ent[rl.o_w() + j] = -lgsj - acc_w;
    }
    __syncthreads();
    for (int d = 0; d < D; ++d) {
        double am = 0.0, al = 0.0;
        const double il = 1.0 / lambd[d];
        for (int i = tid; i < K; i += nt) {
            const double t = (mu[i * D + d] - mu[j * D + d]) * il;
            am = fma(pw[i], t, am);
            al = fma(gj[i], t * t * is2[i] - 1.0, al);
        }
        am = block_sum(am, scratch);
        al = block_sum(al, scratch);
        if (tid == 0) {
            raw_ent[rl.o_mu() + j * D + d] = -am * il;  // :107-110,121-131
            lamp[(size_t)j * D + d] = wj * al;          // :141-156
        }
    }
}

// C: single CTA -> H and the lambda gradient (sums over components); K == 1 closed form (:60-78)
__global__ void entlb_c_kernel(const double *__restrict__ prm, ParamLayout lay, const double *__restrict__ lgs,
                               const double *__restrict__ lamp, double *__restrict__ raw_ent, RawLayout rl,
                               double *__restrict__ Hout) {
    const int D = lay.D, K = lay.K, tid = threadIdx.x, nt = blockDim.x;
    __shared__ double scratch[40];
    const double *sigma = prm + lay.sigma(), *lambd = prm + lay.lambd(), *w = prm + lay.w();
    if (K == 1) {
        if (tid == 0) {
            double sl = 0.0;
            for (int d = 0; d < D; ++d) sl += log(lambd[d]);
            *Hout = 0.5 * D * (1.0 + kLog2Pi) + D * log(sigma[0]) + sl;
            raw_ent[rl.o_sig()] = D / sigma[0];
            raw_ent[rl.o_w()] = 0.0;
            for (int d = 0; d < D; ++d) {
                raw_ent[rl.o_mu() + d] = 0.0;
                raw_ent[rl.o_lam() + d] = 1.0 / lambd[d];
            }
        }
        return;
    }
    double h = 0.0;
    for (int i = tid; i < K; i += nt) h -= w[i] * lgs[i];
    h = block_sum(h, scratch);
    if (tid == 0) *Hout = h;
    for (int d = tid; d < D; d += nt) {
        double a = 0.0;
        for (int j = 0; j < K; ++j) a += lamp[(size_t)j * D + d];
        raw_ent[rl.o_lam() + d] = -a / lambd[d];
    }
}

}  // namespace

// d_raw_ent points at the entropy block of the raw vector; *d_H receives H.
int entlb_launch(Ctx *c, const double *d_params, int D, int K, const int grad[4], double *d_raw_ent, double *d_H) {
    (void)grad;
    ParamLayout lay{D, pad_dim(D) > 0 ? pad_dim(D) : D, K};
    RawLayout rl{D, K};
    const size_t need = (size_t)2 * K * K + K + (size_t)K * D;
    VBMC_TRY(ensure(&c->d_lbws, &c->lbws_cap, need));
    double *lg = c->d_lbws, *r2 = lg + (size_t)K * K, *lgs = r2 + (size_t)K * K, *lamp = lgs + K;
    const int nt = K >= 128 ? 128 : (K > 32 ? 64 : 32);
    if (K > 1) {
        entlb_a_kernel<<<K, nt, 0, c->stream>>>(d_params, lay, lg, r2, lgs);
        VBMC_CUDA_CHECK(cudaGetLastError());
        const size_t smem = (size_t)3 * K * sizeof(double);
        VBMC_REQUIRE(smem <= 200 * 1024, VBMC_ERR_UNSUPPORTED, "entlb: K too large");
        if (smem > 48 * 1024)
            VBMC_CUDA_CHECK(cudaFuncSetAttribute(entlb_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        entlb_b_kernel<<<K, nt, smem, c->stream>>>(d_params, lay, lg, r2, lgs, d_raw_ent, rl, lamp);
        VBMC_CUDA_CHECK(cudaGetLastError());
        c->launches += 2;
    }
    entlb_c_kernel<<<1, 128, 0, c->stream>>>(d_params, lay, lgs, lamp, d_raw_ent, rl, d_H);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace vbmc

// gplj.cu -- GP-surrogate expected log joint (Bayesian quadrature) per hyper-sample and component.
//
// Replaces the S x K Python loop of pyvbmc/vbmc/variational_optimization.py:1374-1465 (reference).
// All fp64: z . alpha is ill-conditioned on real GPs (sigma_f^2 ~ 1e4, sigma_n^2 ~ 1e-5).
//
// One CTA per (component k, hyper-sample s); threads stride over the N training points
// (X is stored TRANSPOSED [D][N] by vbmc_gp_pack so that lanes read consecutive doubles).
//   tau_kd   = sqrt(sigma_k^2 lambda_d^2 + ell_d^2)                                   (:1401)
//   delta_dn = (mu_dk - X_nd) / tau_kd ,  z_n = exp(lnnf_k - 1/2 sum_d delta_dn^2)    (:1402-1406)
//   U   = sum_n alpha_n z_n                      -> I_k = U + m0 + nu_k               (:1407-1424)
//   M_d = sum_n alpha_n z_n delta_dn             -> d/dmu      = -w_k M_d / tau_kd    (:1430-1436)
//   Q_d = sum_n alpha_n z_n delta_dn^2           -> d/dsigma, d/dlambda via (Q_d - U) (:1438-1462)
// The CTA's tail turns the three moments into the finished per-(s,k) terms: I_sk, and the w_k-weighted raw
// gradients w.r.t. mu_k, sigma_k (written into the per-sample block gps[s]) and the lambda contribution
// lamc[s][k][:] (summed over k by finalize_kernel).
#include "common.cuh"

namespace vbmc {
namespace {

template <int DP, bool GRAD>
__global__ void __launch_bounds__(128)
gplj_kernel(const double *__restrict__ prm, ParamLayout lay, const double *__restrict__ Xt,
            const double *__restrict__ alpha, const double *__restrict__ hyp, int hs, int N, int s_begin,
            int s_step, int mean_kind, double *__restrict__ gps, int gps_stride, double *__restrict__ lamc,
            double *__restrict__ Zout) {
    const int D = lay.D, K = lay.K;
    const int k = blockIdx.x, s = s_begin + blockIdx.y * s_step;
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ double s_itau[DP], s_mu[DP], s_red[4][1 + 2 * DP], s_lnnf;

    const double *h = hyp + (size_t)s * hs;
    if (tid < DP) {
        double it = 0.0, m = 0.0;
        if (tid < D) {
            const double sg = prm[lay.sigma() + k], lm = prm[lay.lambd() + tid], el = h[tid];
            it = 1.0 / sqrt(sg * sg * lm * lm + el * el);
            m = prm[lay.mu() + k * D + tid];
        }
        s_itau[tid] = it;
        s_mu[tid] = m;
    }
    __syncthreads();
    if (tid < 32) {
        double acc = 0.0;
        for (int d = tid; d < D; d += 32) acc += log(s_itau[d]);  // -sum ln tau
        acc = warp_sum(acc);
        if (tid == 0) s_lnnf = h[3 * DP + 0] + h[3 * DP + 1] + acc;  // ln sf^2 + sum ln ell - sum ln tau
    }
    __syncthreads();
    const double lnnf = s_lnnf;
    const double *al = alpha + (size_t)s * N;

    double U = 0.0, Mv[GRAD ? DP : 1], Qv[GRAD ? DP : 1];
    if constexpr (GRAD) {
#pragma unroll
        for (int d = 0; d < DP; ++d) Mv[d] = Qv[d] = 0.0;
    }
    for (int n = tid; n < N; n += nt) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int d = 0; d < DP; d += 2) {
            // padded rows of Xt are zero and s_mu / s_itau are zero there: no guards needed
            const double t0 = (s_mu[d] - Xt[(size_t)d * N + n]) * s_itau[d];
            const double t1 = (s_mu[d + 1] - Xt[(size_t)(d + 1) * N + n]) * s_itau[d + 1];
            a0 = fma(t0, t0, a0);
            a1 = fma(t1, t1, a1);
        }
        const double z = exp(lnnf - 0.5 * (a0 + a1));
        if (Zout) Zout[((size_t)s * K + k) * N + n] = z;
        const double za = z * al[n];
        U += za;
        if constexpr (GRAD) {
            // delta is recomputed (2 flops) instead of kept: 40 fewer live registers -> 2x the CTAs per SM
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                const double dl = (s_mu[d] - Xt[(size_t)d * N + n]) * s_itau[d];
                const double t = za * dl;
                Mv[d] += t;
                Qv[d] = fma(t, dl, Qv[d]);
            }
        }
    }

    // block reduction in fixed order (warp shuffle, then <= 4 warps serially)
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    U = warp_sum(U);
    if (lane == 0) s_red[wid][0] = U;
    if constexpr (GRAD) {
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            const double m = warp_sum(Mv[d]), q = warp_sum(Qv[d]);
            if (lane == 0) s_red[wid][1 + d] = m, s_red[wid][1 + DP + d] = q;
        }
    }
    __syncthreads();
    // ---- tail (first warp): moments -> finished per-(s,k) terms ------------------------------------
    if (tid < 32) {
        const RawLayout rl{D, K};
        const bool quad = mean_kind == VBMC_MEAN_NEGQUAD, zero = mean_kind == VBMC_MEAN_ZERO;
        double U = 0.0;
        for (int w = 0; w < nw; ++w) U += s_red[w][0];
        const double sg = prm[lay.sigma() + k], wk = prm[lay.w() + k], s2 = sg * sg;
        double nu = 0.0, asig = 0.0;
        const int d = tid;
        double *gs = gps + (size_t)s * gps_stride;
        if (d < D) {
            const double lm = prm[lay.lambd() + d], m = s_mu[d], it = s_itau[d];
            const double xm = h[DP + d], iom2 = h[2 * DP + d];
            if (quad) nu = iom2 * (m * m + s2 * lm * lm - 2.0 * m * xm + xm * xm);  // :1409-1424
            if constexpr (GRAD) {
                double M = 0.0, Q = 0.0;
                for (int w = 0; w < nw; ++w) M += s_red[w][1 + d], Q += s_red[w][1 + DP + d];
                double gm = -M * it;  // :1430-1436
                double gl = s2 * it * it * lm * (Q - U);  // :1452-1462
                asig = lm * lm * it * it * (Q - U);       // :1438-1450
                if (quad) {
                    gm -= iom2 * (m - xm);
                    gl -= s2 * lm * iom2;
                    asig -= lm * lm * iom2;
                }
                gs[1 + rl.o_mu() + k * D + d] = wk * gm;
                lamc[((size_t)s * K + k) * D + d] = wk * gl;
            }
        }
        nu = warp_sum(nu);
        asig = warp_sum(asig);
        if (tid == 0) {
            const double m0 = zero ? 0.0 : h[3 * DP + 2];
            gs[1 + rl.o_w() + k] = U + m0 - 0.5 * nu;          // I_sk  (:1407-1428, :1464-1465)
            if constexpr (GRAD) gs[1 + rl.o_sig() + k] = wk * sg * asig;
        }
    }
}

}  // namespace

int gplj_launch(Ctx *c, const double *d_params, int K, int s_begin, int s_step, bool anygrad, cudaStream_t stream,
                double *Zout) {
    VBMC_REQUIRE(c->has_gp, VBMC_ERR_STATE, "gp_log_joint: no GP packed (call vbmc_gp_pack first)");
    const int D = c->gD, DP = c->gDP;
    ParamLayout lay{D, DP, K};
    const int S_local = (c->S - s_begin + s_step - 1) / s_step;
    if (S_local <= 0) return VBMC_OK;
    dim3 grid(K, S_local);
    const int nt = c->N >= 96 ? 128 : (c->N >= 48 ? 64 : 32);
    const int hs = hyp_stride(DP);
    const int gst = 1 + RawLayout{D, K}.block();
#define VBMC_CASE(NDP)                                                                                          \
    case NDP:                                                                                                   \
        if (anygrad)                                                                                            \
            gplj_kernel<NDP, true><<<grid, nt, 0, stream>>>(d_params, lay, c->d_Xt, c->d_alpha, c->d_hyp, hs, c->N,   \
                                                            s_begin, s_step, c->mean_kind, c->d_gps, gst,       \
                                                            c->d_lamc, Zout);                                   \
        else                                                                                                    \
            gplj_kernel<NDP, false><<<grid, nt, 0, stream>>>(d_params, lay, c->d_Xt, c->d_alpha, c->d_hyp, hs, c->N,  \
                                                             s_begin, s_step, c->mean_kind, c->d_gps, gst,      \
                                                             c->d_lamc, Zout);                                  \
        break
    switch (DP) {
        VBMC_CASE(4);
        VBMC_CASE(8);
        VBMC_CASE(12);
        VBMC_CASE(16);
        VBMC_CASE(20);
        VBMC_CASE(24);
        VBMC_CASE(28);
        VBMC_CASE(32);
        default:
            set_error("gp_log_joint: unsupported dimension");
            return VBMC_ERR_UNSUPPORTED;
    }
#undef VBMC_CASE
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace vbmc

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
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------------
// Latency-oriented variant (default): ONE WARP per (component k, hyper-sample s), four pairs per CTA.
//   * the training inputs Xt [DP][N] (64 KB at N = 400) are brought into shared memory ONCE per CTA by a single
//     cp.async.bulk (TMA, completion on an mbarrier) while the warps compute their tau_kd / normalisation constants;
//     every later read of a coordinate is a conflict-free shared-memory load (lanes <-> consecutive points);
//   * delta_dn stays in registers between the distance pass and the moment pass (no recomputation, no second read);
//   * the 1 + 2 D moments are reduced inside the warp (one butterfly for U, two transposed 32-value reductions for
//     M_d and Q_d: 62 shuffles each instead of 5 per value) and the same warp finishes the per-(s, k) terms:
//     no __syncthreads after the copy, no cross-warp traffic.
// The kernel above (one CTA per pair, operands from L2) needed 224 registers x 128 threads, ran its 400 CTAs in 1.35
// waves and took 17.5 us at C3 against a DFMA floor of ~2 us; it remains the fallback when Xt does not fit.
constexpr int kGW = 4;

__device__ __forceinline__ uint32_t g_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// sums over the 32 lanes of 32 per-lane doubles at once: afterwards lane l holds sum_lanes v[l]
__device__ __forceinline__ double warp_transpose_sum32d(double (&v)[32], int lane) {
#pragma unroll
    for (int st = 16, n = 16; st >= 1; st >>= 1, n >>= 1) {
        const bool upper = (lane & st) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const double send = upper ? v[i] : v[i + n];
            const double keep = upper ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, st);
        }
    }
    return v[0];
}

template <int DP, bool GRAD>
__global__ void __launch_bounds__(32 * kGW)
gplj_kernel_w(const double *__restrict__ prm, ParamLayout lay, const double *__restrict__ Xt,
              const double *__restrict__ alpha, const double *__restrict__ hyp, int hs, int N, int s_begin, int s_step,
              int S_local, int mean_kind, double *__restrict__ gps, int gps_stride, double *__restrict__ lamc,
              double *__restrict__ Zout) {
    static_assert(DP <= 32, "one lane per dimension in the finishing step");
    extern __shared__ __align__(128) unsigned char gsm[];
    double *sX = reinterpret_cast<double *>(gsm);                    // [DP][N]
    double2 *sC = reinterpret_cast<double2 *>(sX + (size_t)DP * N);  // [kGW][DP] (mu_kd, 1 / tau_kd)
    uint64_t *bar = reinterpret_cast<uint64_t *>(sC + kGW * DP);
    const int D = lay.D, K = lay.K, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t bar_a = g_smem_u32(bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)((size_t)DP * N * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         g_smem_u32(sX)),
                     "l"(Xt), "r"(bytes), "r"(bar_a)
                     : "memory");
    }
    const int p = blockIdx.x * kGW + wid;
    if (p >= K * S_local) return;
    const int k = p % K, s = s_begin + (p / K) * s_step;
    const double *h = hyp + (size_t)s * hs;
    const double sg = prm[lay.sigma() + k];
    // per-pair constants while the copy is in flight: lane d owns dimension d
    double it = 0.0, mu_d = 0.0, lnit = 0.0;
    if (lane < D) {
        const double lm = prm[lay.lambd() + lane], el = h[lane];
        it = 1.0 / sqrt(sg * sg * lm * lm + el * el);
        mu_d = prm[lay.mu() + k * D + lane];
        lnit = log(it);
    }
    if (lane < DP) sC[wid * DP + lane] = make_double2(mu_d, it);
    const double lnnf = h[3 * DP + 0] + h[3 * DP + 1] + warp_sum(lnit);  // ln sf^2 + sum ln ell - sum ln tau
    __syncwarp();
    const double2 *cw = sC + wid * DP;
    const double *al = alpha + (size_t)s * N;
    {  // the tile has landed?
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.b32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a), "r"(0u)
                : "memory");
        }
    }
    double U = 0.0, Mv[GRAD ? DP : 1], Qv[GRAD ? DP : 1];
    if constexpr (GRAD) {
#pragma unroll
        for (int d = 0; d < DP; ++d) Mv[d] = Qv[d] = 0.0;
    }
    for (int n = lane; n < N; n += 32) {
        double dl[DP];
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int d = 0; d < DP; d += 2) {
            // padded rows of Xt are zero and mu / (1 / tau) are zero there: no guards needed
            const double2 c0 = cw[d], c1 = cw[d + 1];
            dl[d] = (c0.x - sX[(size_t)d * N + n]) * c0.y;
            dl[d + 1] = (c1.x - sX[(size_t)(d + 1) * N + n]) * c1.y;
            a0 = fma(dl[d], dl[d], a0);
            a1 = fma(dl[d + 1], dl[d + 1], a1);
        }
        const double z = exp(lnnf - 0.5 * (a0 + a1));
        if (Zout) Zout[((size_t)s * K + k) * N + n] = z;
        const double za = z * al[n];
        U += za;
        if constexpr (GRAD) {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                const double t = za * dl[d];
                Mv[d] += t;
                Qv[d] = fma(t, dl[d], Qv[d]);
            }
        }
    }
    U = warp_sum(U);
    double M = 0.0, Q = 0.0;
    if constexpr (GRAD) {
        double v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = i < DP ? Mv[i < DP ? i : 0] : 0.0;
        M = warp_transpose_sum32d(v, lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = i < DP ? Qv[i < DP ? i : 0] : 0.0;
        Q = warp_transpose_sum32d(v, lane);
    }
    // ---- moments -> finished per-(s, k) terms (lane d <-> dimension d) ---------------------------------------------
    const RawLayout rl{D, K};
    const bool quad = mean_kind == VBMC_MEAN_NEGQUAD, zero = mean_kind == VBMC_MEAN_ZERO;
    const double wk = prm[lay.w() + k], s2 = sg * sg;
    double nu = 0.0, asig = 0.0;
    double *gs = gps + (size_t)s * gps_stride;
    if (lane < D) {
        const int d = lane;
        const double lm = prm[lay.lambd() + d], m = mu_d;
        const double xm = h[DP + d], iom2 = h[2 * DP + d];
        if (quad) nu = iom2 * (m * m + s2 * lm * lm - 2.0 * m * xm + xm * xm);  // :1409-1424
        if constexpr (GRAD) {
            double gm = -M * it;                      // :1430-1436
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
    if (lane == 0) {
        const double m0 = zero ? 0.0 : h[3 * DP + 2];
        gs[1 + rl.o_w() + k] = U + m0 - 0.5 * nu;  // I_sk  (:1407-1428, :1464-1465)
        if constexpr (GRAD) gs[1 + rl.o_sig() + k] = wk * sg * asig;
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
    // default: one warp per (k, s), training inputs staged in shared memory by one bulk copy per CTA
    const size_t wsmem = (size_t)DP * c->N * sizeof(double) + (size_t)kGW * DP * sizeof(double2) + 16;
    static int env_old = -1;
    if (env_old < 0) {
        const char *e = getenv("VBMC_GPLJ_OLD");
        env_old = e ? atoi(e) : 0;
    }
    if (wsmem <= 200 * 1024 && !env_old) {
        const unsigned wgrid = (unsigned)((K * S_local + kGW - 1) / kGW);
#define VBMC_CASEW(NDP)                                                                                               \
    case NDP: {                                                                                                       \
        static size_t set_g = 0, set_v = 0;                                                                           \
        if (anygrad) {                                                                                                \
            if (wsmem > set_g && wsmem > 48 * 1024) {                                                                 \
                VBMC_CUDA_CHECK(cudaFuncSetAttribute(gplj_kernel_w<NDP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); \
                set_g = wsmem;                                                                                        \
            }                                                                                                         \
            gplj_kernel_w<NDP, true><<<wgrid, 32 * kGW, wsmem, stream>>>(d_params, lay, c->d_Xt, c->d_alpha, c->d_hyp, hs, c->N, \
                                                                          s_begin, s_step, S_local, c->mean_kind, c->d_gps, gst, c->d_lamc, Zout); \
        } else {                                                                                                      \
            if (wsmem > set_v && wsmem > 48 * 1024) {                                                                 \
                VBMC_CUDA_CHECK(cudaFuncSetAttribute(gplj_kernel_w<NDP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); \
                set_v = wsmem;                                                                                        \
            }                                                                                                         \
            gplj_kernel_w<NDP, false><<<wgrid, 32 * kGW, wsmem, stream>>>(d_params, lay, c->d_Xt, c->d_alpha, c->d_hyp, hs, c->N, \
                                                                           s_begin, s_step, S_local, c->mean_kind, c->d_gps, gst, c->d_lamc, Zout); \
        }                                                                                                             \
    } break
        switch (DP) {
            VBMC_CASEW(4);
            VBMC_CASEW(8);
            VBMC_CASEW(12);
            VBMC_CASEW(16);
            VBMC_CASEW(20);
            VBMC_CASEW(24);
            VBMC_CASEW(28);
            VBMC_CASEW(32);
            default:
                set_error("gp_log_joint: unsupported dimension");
                return VBMC_ERR_UNSUPPORTED;
        }
#undef VBMC_CASEW
        VBMC_CUDA_CHECK(cudaGetLastError());
        c->launches++;
        return VBMC_OK;
    }
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

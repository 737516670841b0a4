// sieve.cu -- batched, value-only negative ELCBO for the "sieve" of variational_optimization.py:775-787:
// the reference evaluates init_N (up to 50 K) candidate variational posteriors one after the other with
//     _neg_elcbo(theta_b, gp, vp_b, 0, ns_ent_K_fast = 0, compute_grad = 0, compute_var = 0, theta_bnd)
// i.e. F_b = -G_b - H_b + L_bound_b + L_pen_b with the deterministic entropy bound (entlb_vbmc.py:60-97) and
// the expected log joint (variational_optimization.py:1374-1428), no gradients.  Here ALL candidates are one
// launch: a persistent CTA per candidate slot keeps the training inputs in shared memory and loops
//     for s: tau, lnnf, nu per component  ->  one warp per component, lanes over the training points
// fp64 throughout (same arithmetic as gplj_kernel / entlb_a_kernel / tail_kernel, value parts only).
#include "common.cuh"

namespace vbmc {
namespace {

constexpr double kLog2Pi = 1.8378770664093454836;
constexpr int kSieveThreads = 256;

struct SieveArgs {
    ParamLayout lay;
    int B, N, S, hs, mean_kind, x_in_smem;
    int optimize[4], use_bounds, n_bnd;
    double tol_con, w_thr, w_pen;
    const double *prm;  // [B][lay.total()]
    const double *Xt, *alpha, *hyp, *lb, *ub;
    double *out;  // [B][4]: F, G, H, L_bound + L_pen
};

__global__ void __launch_bounds__(kSieveThreads) sieve_kernel(SieveArgs a) {
    const int D = a.lay.D, DP = a.lay.DP, K = a.lay.K, N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = kSieveThreads / 32;
    extern __shared__ double sm[];
    double *s_mu = sm;                 // [K][D]
    double *s_itau = s_mu + K * D;     // [K][D]
    double *s_sig = s_itau + K * D;    // [K]
    double *s_w = s_sig + K;           // [K]
    double *s_lam = s_w + K;           // [D]
    double *s_lnnf = s_lam + D;        // [K]
    double *s_I0 = s_lnnf + K;         // [K]  m0 - nu_k / 2
    double *s_lgs = s_I0 + K;          // [K]  entropy bound: log sum_j w_j gamma_ij
    double *s_red = s_lgs + K;         // [40]
    double *s_X = s_red + 40;          // [DP][N] when it fits
    const double *Xsrc = a.Xt;
    if (a.x_in_smem) {
        for (int i = tid; i < DP * N; i += kSieveThreads) s_X[i] = a.Xt[i];
        Xsrc = s_X;
    }
    __syncthreads();

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double *prm = a.prm + (size_t)b * a.lay.total();
        __syncthreads();
        for (int i = tid; i < K * D; i += kSieveThreads) s_mu[i] = prm[a.lay.mu() + i];
        for (int i = tid; i < K; i += kSieveThreads) s_sig[i] = prm[a.lay.sigma() + i], s_w[i] = prm[a.lay.w() + i];
        for (int i = tid; i < D; i += kSieveThreads) s_lam[i] = prm[a.lay.lambd() + i];
        __syncthreads();

        // ---- expected log joint: G = mean_s sum_k w_k I_sk (:1401-1428, :1581) ---------------------------------
        double Gacc = 0.0;  // lane 0 of every warp
        for (int s = 0; s < a.S; ++s) {
            const double *h = a.hyp + (size_t)s * a.hs;
            const double *al = a.alpha + (size_t)s * N;
            for (int i = tid; i < K * D; i += kSieveThreads) {
                const int k = i / D, d = i - k * D;
                const double sg = s_sig[k], lm = s_lam[d], el = h[d];
                s_itau[i] = 1.0 / sqrt(sg * sg * lm * lm + el * el);
            }
            __syncthreads();
            const bool quad = a.mean_kind == VBMC_MEAN_NEGQUAD, zero = a.mean_kind == VBMC_MEAN_ZERO;
            const double m0 = zero ? 0.0 : h[3 * DP + 2];
            for (int k = wid; k < K; k += nw) {
                double lt = 0.0, nu = 0.0;
                const double s2 = s_sig[k] * s_sig[k];
                for (int d = lane; d < D; d += 32) {
                    lt += log(s_itau[k * D + d]);  // -sum ln tau
                    if (quad) {
                        const double m = s_mu[k * D + d], lm = s_lam[d], xm = h[DP + d];
                        nu += h[2 * DP + d] * (m * m + s2 * lm * lm - 2.0 * m * xm + xm * xm);  // :1409-1424
                    }
                }
                lt = warp_sum(lt), nu = warp_sum(nu);
                if (lane == 0) s_lnnf[k] = h[3 * DP + 0] + h[3 * DP + 1] + lt, s_I0[k] = m0 - 0.5 * nu;
            }
            __syncthreads();
            for (int k = wid; k < K; k += nw) {
                const double lnnf = s_lnnf[k];
                const double *mu = s_mu + k * D, *it = s_itau + k * D;
                double U = 0.0;
                for (int n = lane; n < N; n += 32) {
                    double a0 = 0.0, a1 = 0.0;
                    int d = 0;
                    for (; d + 1 < D; d += 2) {
                        const double t0 = (mu[d] - Xsrc[(size_t)d * N + n]) * it[d];
                        const double t1 = (mu[d + 1] - Xsrc[(size_t)(d + 1) * N + n]) * it[d + 1];
                        a0 = fma(t0, t0, a0), a1 = fma(t1, t1, a1);
                    }
                    if (d < D) {
                        const double t0 = (mu[d] - Xsrc[(size_t)d * N + n]) * it[d];
                        a0 = fma(t0, t0, a0);
                    }
                    U += al[n] * exp(lnnf - 0.5 * (a0 + a1));
                }
                U = warp_sum(U);
                if (lane == 0) Gacc += s_w[k] * (U + s_I0[k]);  // I_sk = U + m0 - nu / 2
            }
            __syncthreads();
        }
        double G = block_sum(lane == 0 ? Gacc : 0.0, s_red) / (double)a.S;

        // ---- entropy lower bound (entlb_vbmc.py:60-97): H = -sum_i w_i log sum_j w_j gamma_ij ------------------
        double sl = 0.0;
        for (int d = lane; d < D; d += 32) sl += log(s_lam[d]);
        sl = warp_sum(sl);
        double H;
        if (K == 1) {
            H = 0.5 * D * (1.0 + kLog2Pi) + D * log(s_sig[0]) + sl;  // closed form (:60-78)
        } else {
            const double c0 = -0.5 * D * kLog2Pi - sl;
            for (int i = wid; i < K; i += nw) {
                const double si2 = s_sig[i] * s_sig[i];
                double mx = -INFINITY, se = 0.0;  // running log-sum-exp over this lane's components j
                for (int j = lane; j < K; j += 32) {
                    const double s2 = si2 + s_sig[j] * s_sig[j];
                    double r2 = 0.0;
                    for (int d = 0; d < D; ++d) {
                        const double t = (s_mu[i * D + d] - s_mu[j * D + d]) / s_lam[d];
                        r2 = fma(t, t, r2);
                    }
                    const double v = c0 - 0.5 * D * log(s2) - 0.5 * r2 / s2 + log(s_w[j]);
                    if (v == -INFINITY) continue;  // zero weight
                    if (v > mx) se = se * exp(mx - v) + 1.0, mx = v;
                    else se += exp(v - mx);
                }
                const double gm = warp_max(mx);
                se = warp_sum(mx == -INFINITY ? 0.0 : se * exp(mx - gm));
                if (lane == 0) s_lgs[i] = gm + log(se);
            }
            __syncthreads();
            double hp = 0.0;
            for (int i = tid; i < K; i += kSieveThreads) hp -= s_w[i] * s_lgs[i];
            H = block_sum(hp, s_red);
        }

        // ---- soft bounds and weight penalty (:503-657, :1212-1229), value parts -------------------------------
        double Lb = 0.0, Lp = 0.0;
        if (a.use_bounds && a.n_bnd > 0) {
            const int n_mu = a.optimize[0] ? K * D : 0, n_sc = K * D, n_eta = a.optimize[3] ? K : 0;
            for (int e = tid; e < n_mu + n_sc + n_eta; e += kSieveThreads) {
                double x;
                if (e < n_mu)
                    x = s_mu[e];
                else if (e < n_mu + n_sc) {
                    const int i = e - n_mu, k = i / D, d = i - k * D;  // column-major (D,K) ravel (:557-562)
                    x = prm[a.lay.lnlam_b() + d] + prm[a.lay.lnsig_b() + k];
                } else
                    x = prm[a.lay.eta_b() + e - n_mu - n_sc];
                const double lo = a.lb[e], hi = a.ub[e];
                const double viol = x < lo ? x - lo : (x > hi ? x - hi : 0.0);
                if (viol != 0.0) {
                    const double r = viol / ((hi - lo) * a.tol_con);
                    Lb += 0.5 * r * r;
                }
            }
            if (a.optimize[3])
                for (int k = tid; k < K; k += kSieveThreads) Lp += (s_w[k] < a.w_thr ? s_w[k] : a.w_thr) * a.w_pen;
            Lb = block_sum(Lb, s_red);
            Lp = block_sum(Lp, s_red);
        }
        if (tid == 0) {
            double *o = a.out + (size_t)b * 4;
            o[0] = -G - H + Lb + Lp;
            o[1] = G;
            o[2] = H;
            o[3] = Lb + Lp;
        }
    }
}

}  // namespace

// d_prm: [B][ParamLayout::total()] device, d_out: [B][4] device
int sieve_launch(Ctx *c, int B, int D, int K, const int optimize[4], bool use_bounds, const double *d_prm, double *d_out) {
    VBMC_REQUIRE(c->has_gp, VBMC_ERR_STATE, "negelcbo_batch: no GP packed (call vbmc_gp_pack first)");
    VBMC_REQUIRE(c->gD == D, VBMC_ERR_ARG, "negelcbo_batch: D does not match the packed GP");
    if (B <= 0) return VBMC_OK;
    SieveArgs a{};
    const int DP = c->gDP;
    a.lay = ParamLayout{D, DP, K};
    a.B = B, a.N = c->N, a.S = c->S, a.hs = hyp_stride(DP), a.mean_kind = c->mean_kind;
    for (int i = 0; i < 4; ++i) a.optimize[i] = optimize[i];
    a.use_bounds = use_bounds ? 1 : 0;
    a.n_bnd = use_bounds ? c->n_bnd : 0;
    if (use_bounds) {
        const int n_expect = (optimize[0] ? K * D : 0) + K * D + (optimize[3] ? K : 0);
        VBMC_REQUIRE(c->n_bnd == n_expect, VBMC_ERR_ARG, "soft bounds: lb/ub length does not match [mu|ln-scale|eta]");
    }
    a.tol_con = c->tol_con, a.w_thr = c->w_thr, a.w_pen = c->w_pen;
    a.prm = d_prm, a.Xt = c->d_Xt, a.alpha = c->d_alpha, a.hyp = c->d_hyp, a.lb = c->d_lb, a.ub = c->d_ub;
    a.out = d_out;
    const size_t base = ((size_t)2 * K * D + 5 * K + D + 40) * sizeof(double);
    const size_t xs = (size_t)DP * c->N * sizeof(double);
    a.x_in_smem = base + xs <= 200 * 1024;
    const size_t smem = base + (a.x_in_smem ? xs : 0);
    VBMC_REQUIRE(base <= 200 * 1024, VBMC_ERR_UNSUPPORTED, "negelcbo_batch: D*K too large for shared memory");
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(sieve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const int per_sm = smem > 110 * 1024 ? 1 : 2;
    const int grid = std::min(B, c->sm_count * per_sm);
    sieve_kernel<<<grid, kSieveThreads, smem, c->stream>>>(a);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace vbmc

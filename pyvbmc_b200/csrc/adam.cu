// adam.cu -- device-resident minimize_adam (pyvbmc/vbmc/minimize_adam.py:61-145) around the negative-ELCBO
// evaluation: theta, the moment estimates, the iterate table and the iteration counter live in HBM, one iteration =
//     adam_prepare_kernel   theta -> parameter block (VariationalPosterior.set_parameters, variational_posterior.py:
//                           680-759 incl. the lambda normalisation, softmax weights; the eta shift :1082-1085; the
//                           theta slices the bound loss reads :536-555; Philox offset = offset0 + iteration)
//     gplj / entmc / tail   the evaluation itself, output (F, dF) left in device memory
//     adam_update_kernel    moment update, bias correction, step-size schedule, clamp to [lb, ub], x_tab / y_tab
// captured as ONE CUDA graph and replayed back to back: no H2D / D2H / host synchronisation per iteration.  The
// early-stopping test (a linear fit every 20 iterations, :106-138) stays on the host and reads y_tab / x_tab in
// batches.
#include "common.cuh"

namespace vbmc {
namespace {

// theta -> parameter block: shared by the Adam loop (theta in HBM) and by vbmc_negelcbo_theta (theta, the template of
// the groups theta does not carry, and the Philox key in PINNED HOST memory, read through their device aliases; the
// normalised sigma / lambda / w go back to the host the same way in `vp_out` = [sigma (K) | lambda (D) | w (K)]).
__device__ __forceinline__ void theta_to_params(const AdamDev &a, double *__restrict__ prm, double *__restrict__ vp_out,
                                                const uint64_t *__restrict__ key_src, bool skip_mu = false) {
    const ParamLayout lay = a.lay;
    const int D = lay.D, K = lay.K, tid = threadIdx.x, nt = blockDim.x;
    __shared__ double scratch[40];
    __shared__ uint64_t skey[2];
    extern __shared__ double stg[];  // theta behind the means [<= 2K + D] | template sigma, lambda, w, eta [3K + D]
    double *sth = stg, *stm = stg + 2 * K + D;
    const double *th = a.theta, *tm = a.tmpl;
    const int n_mu = a.opt[0] ? D * K : 0, n_tail = a.P - n_mu;
    const int pos_s = 0, pos_l = a.opt[1] ? K : 0;  // offsets inside sth
    const bool need_tm = !(a.opt[0] && a.opt[1] && a.opt[2] && a.opt[3]);
    // ONE pass over the inputs, every load independent: on the drop-in path theta, the template and the key live in
    // pinned HOST memory, and each dependent round of reads would cost a PCIe round trip (the first version read theta
    // phase by phase: 16 us for this 1-CTA kernel)
    for (int e = tid; e < n_tail; e += nt) sth[e] = th[n_mu + e];
    if (need_tm)
        for (int e = tid; e < 3 * K + D; e += nt) stm[e] = tm[lay.sigma() + e];  // sigma | lambda | w | eta are contiguous
    if (tid == 0 && key_src) skey[0] = key_src[0], skey[1] = key_src[1];
    // mu (theta order is component-major already)
    if (!skip_mu)
        for (int e = tid; e < D * K; e += nt) prm[lay.mu() + e] = a.opt[0] ? th[e] : tm[lay.mu() + e];
    __syncthreads();
    // The three reductions (sum of lambda^2; max and sum of exp of eta) run in TWO WARPS side by side with shuffles only,
    // then one barrier publishes them: this 1-CTA kernel heads the critical path of every evaluation, and each
    // block-wide reduction of the first version cost two more barriers (13 us for the fused Adam kernel under ncu).
    const int lane = tid & 31, wid = tid >> 5;
    if (wid == 0) {  // lambda: exp, then the unit-RMS normalisation shared with sigma (:749-756)
        double l2 = 0.0;
        for (int d = lane; d < D; d += 32) {
            const double lm = a.opt[2] ? exp(sth[pos_l + d]) : stm[K + d];
            l2 += lm * lm;
        }
        l2 = warp_sum(l2);
        if (lane == 0) scratch[0] = sqrt(l2 / D);
    } else if (wid == 1 && a.opt[3]) {  // weights: softmax of eta with the max shift (:735-741)
        const double *eta = sth + n_tail - K;
        double mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmax(mx, eta[k]);
        mx = warp_max(mx);
        double se = 0.0;
        for (int k = lane; k < K; k += 32) se += exp(eta[k] - mx);
        se = warp_sum(se);
        if (lane == 0) scratch[1] = mx, scratch[2] = se;
    }
    __syncthreads();
    const double scale = scratch[0];
    for (int d = tid; d < D; d += nt) {
        const double lm = (a.opt[2] ? exp(sth[pos_l + d]) : stm[K + d]) / scale;
        prm[lay.lambd() + d] = lm;
        if (vp_out) vp_out[K + d] = lm;
        prm[lay.lnlam_b() + d] = a.opt[2] ? sth[pos_l + d] : log(lm);  // (:536-555: theta's slice, else log of the field)
    }
    for (int k = tid; k < K; k += nt) {
        const double sg = (a.opt[1] ? exp(sth[pos_s + k]) : stm[k]) * scale;
        prm[lay.sigma() + k] = sg;
        if (vp_out) vp_out[k] = sg;
        prm[lay.lnsig_b() + k] = a.opt[1] ? sth[pos_s + k] : log(sg);
    }
    // eta itself is stored shifted (:1082-1085).  The reference shifts the eta block of the caller's theta IN PLACE
    // (`vp.eta = theta[-K:]; vp.eta -= amax`: the slice is a view): the soft-bound loss reads the shifted eta
    // (:1195-1209) and minimize_adam's `x -= step` (minimize_adam.py:98) updates the renormalised iterate.  Same here:
    // theta's eta block is rewritten with max == 0 before the evaluation and the update.
    if (a.opt[3]) {
        const double *eta = sth + n_tail - K;
        const double mx = scratch[1], se = scratch[2];
        for (int k = tid; k < K; k += nt) {
            const double e = eta[k] - mx;
            const double wk = exp(e) / se;
            prm[lay.w() + k] = wk;
            if (vp_out) vp_out[K + D + k] = wk;
            prm[lay.eta() + k] = e;
            prm[lay.eta_b() + k] = e;
            a.theta[a.P - K + k] = e;
        }
    } else {
        for (int k = tid; k < K; k += nt) {
            prm[lay.w() + k] = stm[K + D + k];
            if (vp_out) vp_out[K + D + k] = stm[K + D + k];
            prm[lay.eta() + k] = stm[2 * K + D + k];
            prm[lay.eta_b() + k] = stm[2 * K + D + k];
        }
    }
    if (tid == 0) {  // Philox key rides behind the parameter block
        uint64_t *key = reinterpret_cast<uint64_t *>(prm + lay.total());
        if (key_src) {
            key[0] = skey[0], key[1] = skey[1];
        } else {
            key[0] = a.seed;
            key[1] = a.offset0 + (uint64_t)(*a.iter);
        }
    }
}

__global__ void __launch_bounds__(256) adam_prepare_kernel(AdamDev a, double *__restrict__ prm) {
    theta_to_params(a, prm, nullptr, nullptr);
}

__global__ void __launch_bounds__(256)
theta_prepare_kernel(AdamDev a, double *__restrict__ prm, double *__restrict__ vp_out, const uint64_t *__restrict__ key_src) {
    // theta lives in pinned HOST memory here, and reads over PCIe are limited by the requests in flight, not by their
    // size: the D*K means (9 of the 10 KB) are copied by CTAs 1.. while CTA 0 works on the 2K + D entries behind them
    if (blockIdx.x > 0) {
        const int n = a.lay.D * a.lay.K, stride = ((int)gridDim.x - 1) * (int)blockDim.x;
        const double *src = a.opt[0] ? a.theta : a.tmpl + a.lay.mu();
        for (int e = ((int)blockIdx.x - 1) * (int)blockDim.x + (int)threadIdx.x; e < n; e += stride) prm[a.lay.mu() + e] = src[e];
        return;
    }
    theta_to_params(a, prm, vp_out, key_src, gridDim.x > 1);
}

__device__ __forceinline__ void adam_update(const AdamDev &a, const double *__restrict__ out) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long i = *a.iter;  // 0-based iteration
    const double beta_1 = 0.9, beta_2 = 0.999, fudge = 1.4901161193847656e-08;  // sqrt(np.spacing(1))
    __shared__ double s_sc[3];
    if (tid < 3) {  // three lanes, one transcendental each (instead of two pow() and an exp() in every thread)
        s_sc[tid] = tid == 0 ? 1.0 - pow(beta_1, (double)(i + 1))
                             : (tid == 1 ? 1.0 - pow(beta_2, (double)(i + 1))
                                         : a.master_min + (a.master_max - a.master_min) * exp(-(double)(i + 1) / a.master_decay));
    }
    __syncthreads();
    const double c1 = s_sc[0], c2 = s_sc[1], step = s_sc[2];
    double *xrow = a.xtab + (size_t)i * a.P;
    for (int e = tid; e < a.P; e += nt) {
        const double g = out[kOutHead + e];
        const double m = beta_1 * a.m[e] + (1.0 - beta_1) * g;
        const double v = beta_2 * a.v[e] + (1.0 - beta_2) * g * g;
        a.m[e] = m, a.v[e] = v;
        double x = a.theta[e] - step * (m / c1) / (sqrt(v / c2) + fudge);
        if (a.lb) x = fmax(a.lb[e], x);
        if (a.ub) x = fmin(a.ub[e], x);
        a.theta[e] = x;
        xrow[e] = x;
    }
    if (tid == 0) a.ytab[i] = out[0];
    __syncthreads();
    if (tid == 0) *a.iter = i + 1;
}

__global__ void __launch_bounds__(1024) adam_update_kernel(AdamDev a, const double *__restrict__ out) { adam_update(a, out); }

// update of iteration i FUSED with the parameter block of iteration i + 1 (one launch less on the critical path of every
// iteration; the block prepared after the last iteration is simply never used)
__global__ void __launch_bounds__(1024)
adam_update_prepare_kernel(AdamDev a, const double *__restrict__ out, double *__restrict__ prm) {
    adam_update(a, out);
    __threadfence_block();
    __syncthreads();  // the new theta and the incremented iteration counter are visible to the whole CTA
    theta_to_params(a, prm, nullptr, nullptr);
}

}  // namespace

int adam_prepare_launch(Ctx *c, const AdamDev &a, double *d_prm) {
    adam_prepare_kernel<<<1, 256, (size_t)(5 * a.lay.K + 2 * a.lay.D) * sizeof(double), c->stream>>>(a, d_prm);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

int theta_prepare_launch(Ctx *c, const AdamDev &a, double *d_prm, double *vp_out, const uint64_t *key_src) {
    static const int mu_ctas = getenv("VBMC_PREP_MU_CTAS") ? atoi(getenv("VBMC_PREP_MU_CTAS")) : 4;
    theta_prepare_kernel<<<1 + mu_ctas, 256, (size_t)(5 * a.lay.K + 2 * a.lay.D) * sizeof(double), c->stream>>>(a, d_prm, vp_out, key_src);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

int adam_update_prepare_launch(Ctx *c, const AdamDev &a, const double *d_out, double *d_prm) {
    adam_update_prepare_kernel<<<1, 1024, (size_t)(5 * a.lay.K + 2 * a.lay.D) * sizeof(double), c->stream>>>(a, d_out, d_prm);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

int adam_update_launch(Ctx *c, const AdamDev &a, const double *d_out) {
    adam_update_kernel<<<1, 1024, 0, c->stream>>>(a, d_out);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace vbmc

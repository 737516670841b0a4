// finalize.cu -- fixed-order reduction of the kernel partials into the RAW vector, and the
// final assembly (Jacobians, soft-bound loss, weight penalty) of F and dF.
//
//   reduce_kernel   : entmc CTA records + gplj (s,k) records  ->  raw = [H, G, .. | ent | gp]
//                     (raw = pre-Jacobian sums, already scaled by the GLOBAL 1/Ns and 1/S, so a
//                      SUM all-reduce over ranks yields the single-GPU value)
//   finalize_kernel : raw -> out = [F, G, H, .., | dF | dH | dG]
//                     log/softmax Jacobians  entmc_vbmc.py:114-130, variational_optimization.py:1522-1548
//                     soft bounds            variational_optimization.py:503-657 (incl. the row-major
//                                            reshape of the column-major ln-scale block at :584-586)
//                     weight penalty         variational_optimization.py:1212-1229
// Both are single-CTA kernels: O(S K D) work, fp64, deterministic summation order.
#include "common.cuh"

namespace vbmc {
namespace {

constexpr double kLog2Pi = 1.8378770664093454836;

struct ReduceArgs {
    ParamLayout lay;
    RawLayout rl;
    EvalFlags f;
    // entmc
    const double *entpart;
    int slabs, ent_stride;
    double Ns_glob;     // draws per component over all ranks
    double draws_local; // draws per component on this rank
    double *crec;       // [K][ent_stride] per-component records
    // gp
    const double *gppart;
    const double *hyp;
    int hs, s_begin, s_step, S, S_glob;
    int mean_kind;
    double *gps;  // [S][1 + block]
    double *raw;
};

// Stage 1 (many CTAs): CTA j < K sums the slab records of mixture component j in slab order;
// CTA K + i turns the (s_i, k) log-joint records into G_s and the raw per-sample gradients.
__global__ void __launch_bounds__(128) reduce_kernel(const double *__restrict__ prm, ReduceArgs a) {
    const int D = a.lay.D, DP = a.lay.DP, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    __shared__ double scratch[40];
    const double *mu = prm + a.lay.mu(), *sigma = prm + a.lay.sigma(), *lambd = prm + a.lay.lambd(),
                 *w = prm + a.lay.w();
    const RawLayout rl = a.rl;
    const bool ent_on = a.f.have_ent && a.f.use_ent_mc;
    const int n_ent = ent_on ? K : 0;

    if ((int)blockIdx.x < n_ent) {
        // ---------------------------------------------------------- Monte-Carlo entropy, component j
        const int j = blockIdx.x, st = a.ent_stride;
        // last warp: sum_d ln lambda_d (lanes over d) while the others stream the slab records
        if (tid >= nt - 32) {
            double sl = 0.0;
            for (int d = tid - (nt - 32); d < D; d += 32) sl += log(lambd[d]);
            sl = warp_sum(sl);
            if (tid == nt - 32) scratch[0] = sl;
        }
        double v0 = 0.0;
        for (int t = tid; t < st; t += nt) {
            double v = 0.0;
#pragma unroll 4
            for (int s = 0; s < a.slabs; ++s) v += a.entpart[((size_t)j * a.slabs + s) * st + t];
            if (t == 0)
                v0 = v;
            else
                a.crec[(size_t)j * st + t] = v;
        }
        __syncthreads();
        // + (draws of this rank) * log( nconst / sigma_j^D )   (entmc_vbmc.py:53-56,77)
        if (tid == 0) a.crec[(size_t)j * st] = v0 + a.draws_local * (-0.5 * D * kLog2Pi - scratch[0] - D * log(sigma[j]));
        return;
    }
    if (!a.f.have_gp) return;
    // -------------------------------------------------------------- GP expected log joint, sample s
    const int s = a.s_begin + ((int)blockIdx.x - n_ent) * a.s_step;
    if (s >= a.S) return;
    const int gst = 1 + 2 * DP;
    const int blk = rl.block();
    const bool quad = a.mean_kind == VBMC_MEAN_NEGQUAD, zero = a.mean_kind == VBMC_MEAN_ZERO;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const double *h = a.hyp + (size_t)s * a.hs;
    const double *ell = h, *xm = h + DP, *iom2 = h + 2 * DP;
    const double m0 = zero ? 0.0 : h[3 * DP + 2];
    double *gs = a.gps + (size_t)s * (1 + blk);
    const double *rec_s = a.gppart + (size_t)s * K * gst;
    // I_sk -> the w block (variational_optimization.py:1407-1428,1464-1465)
    double gpart = 0.0;
    for (int k = tid; k < K; k += nt) {
        double I = rec_s[(size_t)k * gst] + m0;
        if (quad) {
            const double s2 = sigma[k] * sigma[k];
            double nu = 0.0;
            for (int d = 0; d < D; ++d) {
                const double m = mu[k * D + d];
                nu += iom2[d] * (m * m + s2 * lambd[d] * lambd[d] - 2.0 * m * xm[d] + xm[d] * xm[d]);
            }
            I -= 0.5 * nu;
        }
        gs[1 + rl.o_w() + k] = I;
        gpart += w[k] * I;
    }
    gpart = block_sum(gpart, scratch);
    if (tid == 0) gs[0] = gpart;  // G_s (:1425)
    if (!anyg) return;
    // d/dmu (:1430-1436)
    for (int e = tid; e < K * D; e += nt) {
        const int k = e / D, d = e - k * D;
        const double sl = sigma[k] * lambd[d];
        const double tau = sqrt(sl * sl + ell[d] * ell[d]);
        double g = -rec_s[(size_t)k * gst + 1 + d] / tau;
        if (quad) g -= iom2[d] * (mu[e] - xm[d]);
        gs[1 + rl.o_mu() + e] = w[k] * g;
    }
    // d/dsigma (:1438-1450)
    for (int k = tid; k < K; k += nt) {
        const double U = rec_s[(size_t)k * gst];
        double acc = 0.0, accq = 0.0;
        for (int d = 0; d < D; ++d) {
            const double sl = sigma[k] * lambd[d];
            const double t2 = sl * sl + ell[d] * ell[d];
            acc += lambd[d] * lambd[d] / t2 * (rec_s[(size_t)k * gst + 1 + DP + d] - U);
            accq += lambd[d] * lambd[d] * iom2[d];
        }
        double g = sigma[k] * acc;
        if (quad) g -= sigma[k] * accq;
        gs[1 + rl.o_sig() + k] = w[k] * g;
    }
    // d/dlambda (:1452-1462): warp per dimension, lanes over components, fixed-order shuffle tree
    {
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        for (int d = wid; d < D; d += nw) {
            double acc = 0.0;
            for (int k = lane; k < K; k += 32) {
                const double s2 = sigma[k] * sigma[k];
                const double t2 = s2 * lambd[d] * lambd[d] + ell[d] * ell[d];
                double g = s2 / t2 * lambd[d] * (rec_s[(size_t)k * gst + 1 + DP + d] - rec_s[(size_t)k * gst]);
                if (quad) g -= s2 * lambd[d] * iom2[d];
                acc += w[k] * g;
            }
            acc = warp_sum(acc);
            if (lane == 0) gs[1 + rl.o_lam() + d] = acc;
        }
    }
}

// Stage 2, part A (device function, whole CTA): per-component / per-sample records -> raw vector.
__device__ void assemble_raw(const double *__restrict__ prm, const ReduceArgs &a, double *scratch) {
    const int D = a.lay.D, DP = a.lay.DP, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    const double *sigma = prm + a.lay.sigma(), *lambd = prm + a.lay.lambd(), *w = prm + a.lay.w();
    const RawLayout rl = a.rl;
    double *raw = a.raw;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const bool ent_mc = a.f.have_ent && a.f.use_ent_mc;
    const bool ent_lb = a.f.have_ent && !a.f.use_ent_mc;  // entlb kernels already wrote raw[0] and the block

    if (tid >= 1 && tid < 4) raw[tid] = 0.0;
    if (tid == 0 && !ent_lb) raw[0] = 0.0;
    if (!a.f.have_ent)
        for (int e = tid; e < rl.block(); e += nt) raw[rl.ent() + e] = 0.0;
    if (!a.f.have_gp)
        for (int e = tid; e < rl.block(); e += nt) raw[rl.gp() + e] = 0.0;

    if (ent_mc) {
        const double inv_ns = 1.0 / a.Ns_glob;
        const int st = a.ent_stride;
        double *ent = raw + rl.ent();
        double hpart = 0.0;
        for (int j = tid; j < K; j += nt) hpart -= w[j] * a.crec[(size_t)j * st] * inv_ns;  // entmc_vbmc.py:80
        hpart = block_sum(hpart, scratch);
        if (tid == 0) raw[0] = hpart;
        if (anyg) {
            // d/dmu_j (:98): element-wise
            for (int e = tid; e < K * D; e += nt) {
                const int j = e / D, d = e - j * D;
                ent[rl.o_mu() + e] = w[j] * a.crec[(size_t)j * st + 1 + d] * inv_ns / lambd[d];
            }
            // cross sums: one WARP per output entry, lanes over the summed index, butterfly tree
            const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
            for (int idx = wid; idx < 2 * K + D; idx += nw) {
                double u = 0.0;
                if (idx < K) {  // d/dsigma_j (:102-103): sum over dimensions
                    for (int d = lane; d < D; d += 32) u += a.crec[(size_t)idx * st + 1 + DP + d];
                    u = warp_sum(u);
                    if (lane == 0) ent[rl.o_sig() + idx] = w[idx] * u * inv_ns / sigma[idx];
                } else if (idx < K + D) {  // d/dlambda_d (:106-108): sum over components
                    const int d = idx - K;
                    for (int j = lane; j < K; j += 32) u += w[j] * a.crec[(size_t)j * st + 1 + DP + d];
                    u = warp_sum(u);
                    if (lane == 0) ent[rl.o_lam() + d] = u * inv_ns / lambd[d];
                } else {  // d/dw_k (:111-112): direct term + sum over components
                    const int k = idx - K - D;
                    if (a.f.grad[3])
                        for (int j = lane; j < K; j += 32) u += w[j] * a.crec[(size_t)j * st + 1 + 2 * DP + k];
                    u = warp_sum(u);
                    if (lane == 0) ent[rl.o_w() + k] = -(a.crec[(size_t)k * st] + u) * inv_ns;
                }
            }
        }
    }
    if (a.f.have_gp) {
        // average over hyper-samples (:1578-1596); this rank contributes its own s / S_glob
        const int blk = rl.block();
        const double inv_S = 1.0 / (double)a.S_glob;
        for (int e = tid; e < blk + 1; e += nt) {
            if (e > 0 && !anyg) break;
            double v = 0.0;
#pragma unroll 4
            for (int s = a.s_begin; s < a.S; s += a.s_step) v += a.gps[(size_t)s * (1 + blk) + e];
            v *= inv_S;
            if (e == 0)
                raw[1] = v;
            else
                raw[rl.gp() + e - 1] = v;
        }
    }
    __syncthreads();
}

// Apply the reparameterisation Jacobians to one raw block and scatter it into theta order.
// es = sum exp(eta), dot = sum exp(eta_k) gw_k.  Returns the packed length via *P (thread 0 view).
__device__ __forceinline__ void pack_block(const double *__restrict__ blk, const RawLayout rl, const double *prm,
                                           const ParamLayout lay, const int grad[4], int jac, double es, double dot,
                                           double sign, double *__restrict__ dst, bool accumulate) {
    const int D = lay.D, K = lay.K, tid = threadIdx.x, nt = blockDim.x;
    int off = 0;
    if (grad[0]) {
        for (int e = tid; e < K * D; e += nt) {
            const double v = sign * blk[rl.o_mu() + e];
            dst[off + e] = accumulate ? dst[off + e] + v : v;
        }
        off += K * D;
    }
    if (grad[1]) {
        for (int k = tid; k < K; k += nt) {
            double v = sign * blk[rl.o_sig() + k];
            if (jac) v *= prm[lay.sigma() + k];
            dst[off + k] = accumulate ? dst[off + k] + v : v;
        }
        off += K;
    }
    if (grad[2]) {
        for (int d = tid; d < D; d += nt) {
            double v = sign * blk[rl.o_lam() + d];
            if (jac) v *= prm[lay.lambd() + d];
            dst[off + d] = accumulate ? dst[off + d] + v : v;
        }
        off += D;
    }
    if (grad[3]) {
        for (int k = tid; k < K; k += nt) {
            double v = blk[rl.o_w() + k];
            if (jac) {
                const double ek = exp(prm[lay.eta() + k]);
                v = ek / es * v - ek / (es * es) * dot;  // row k of J_w @ g   (entmc_vbmc.py:122-130)
            }
            v *= sign;
            dst[off + k] = accumulate ? dst[off + k] + v : v;
        }
    }
}

struct FinalArgs {
    ParamLayout lay;
    RawLayout rl;
    EvalFlags f;
    int do_assemble, do_finalize;
    ReduceArgs red;  // used when do_assemble
    const double *raw;
    const double *lb, *ub;
    int n_bnd;
    double tol_con, w_thr, w_pen;
    double *out;
    int Pfull;
};

// Stage 2 (single CTA, 1024 threads): [assemble raw] and/or [finalize].  On one GPU both run in one
// launch; with several ranks the all-reduce of raw sits between an assemble-only and a finalize-only launch.
// Every output element is written exactly once; cross-element sums (softmax, the reference's row-major
// reshape of the ln-scale block) go through shared memory with one warp per output entry.
__global__ void __launch_bounds__(1024) finalize_kernel(const double *__restrict__ prm, FinalArgs a) {
    const int D = a.lay.D, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    __shared__ double scratch[40];
    extern __shared__ double fsm[];  // tmp [K*D] | add_sig [K] | add_lam [D] | add_eta [K]
    double *tmp = fsm, *add_sig = tmp + K * D, *add_lam = add_sig + K, *add_eta = add_lam + D;
    if (a.do_assemble) assemble_raw(prm, a.red, scratch);
    if (!a.do_finalize) return;

    const RawLayout rl = a.rl;
    const double *eta = prm + a.lay.eta(), *w = prm + a.lay.w();
    double *out = a.out;
    double *dF = out + kOutHead, *dH = dF + a.Pfull, *dG = dH + a.Pfull;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const bool bounds = a.f.use_bounds && a.n_bnd > 0;

    for (int i = tid; i < 2 * K + D; i += nt) add_sig[i] = 0.0;  // add_sig | add_lam | add_eta are contiguous

    // ---- softmax pieces: es = sum exp(eta), <exp(eta), gw> for both blocks, and for the weight penalty
    double es = 0.0, dote = 0.0, dotg = 0.0, dotp = 0.0, Lp = 0.0;
    const bool need_sm = (anyg && a.f.grad[3] && a.f.jacobian) || (bounds && a.f.optimize[3]);
    if (need_sm) {
        for (int k = tid; k < K; k += nt) {
            const double ek = exp(eta[k]);
            es += ek;
            dote += ek * a.raw[rl.ent() + rl.o_w() + k];
            dotg += ek * a.raw[rl.gp() + rl.o_w() + k];
            if (bounds && a.f.optimize[3]) {
                const double wk = w[k];
                Lp += (wk < a.w_thr ? wk : a.w_thr) * a.w_pen;       // :1213-1219
                dotp += ek * (wk < a.w_thr ? a.w_pen : 0.0);
            }
        }
        es = block_sum(es, scratch);
        dote = block_sum(dote, scratch);
        dotg = block_sum(dotg, scratch);
        if (bounds && a.f.optimize[3]) {
            Lp = block_sum(Lp, scratch);
            dotp = block_sum(dotp, scratch);
        }
    }

    // ---- soft bounds on [mu | ln sigma_k + ln lambda_d | eta]  (:503-657) -----------------
    double Lb = 0.0;
    const int n_mu = a.f.optimize[0] ? K * D : 0, n_sc = K * D, n_eta = a.f.optimize[3] ? K : 0;
    if (bounds) {
        const double *lnsig = prm + a.lay.lnsig_b(), *lnlam = prm + a.lay.lnlam_b(), *etab = prm + a.lay.eta_b();
        const double *mu = prm + a.lay.mu();
        for (int e = tid; e < n_mu + n_sc + n_eta; e += nt) {
            double x;
            if (e < n_mu)
                x = mu[e];
            else if (e < n_mu + n_sc) {
                const int i = e - n_mu, k = i / D, d = i - k * D;  // column-major (D,K) ravel (:557-562)
                x = lnlam[d] + lnsig[k];
            } else
                x = etab[e - n_mu - n_sc];
            const double lo = a.lb[e], hi = a.ub[e];
            const double ell = (hi - lo) * a.tol_con;
            const double viol = x < lo ? x - lo : (x > hi ? x - hi : 0.0);
            double dy = 0.0;
            if (viol != 0.0) {
                const double r = viol / ell;
                Lb += 0.5 * r * r;
                dy = viol / (ell * ell);
            }
            // mu part is re-derived element-wise in the final pass; keep the other two in shared memory
            if (e >= n_mu && e < n_mu + n_sc)
                tmp[e - n_mu] = dy;
            else if (e >= n_mu + n_sc)
                add_eta[e - n_mu - n_sc] = dy;
        }
        Lb = block_sum(Lb, scratch);  // (contains the barrier that publishes tmp / add_eta)
        if (anyg) {
            // the reference reshapes the ln-scale gradient ROW-major to (D, K): dls[r][b] = tmp[r*K + b]
            // (:584-586); sigma gets the column sums, lambda the row sums.  One warp per output entry.
            for (int idx = wid; idx < K + D; idx += nw) {
                double v = 0.0;
                if (idx < K) {
                    for (int r = lane; r < D; r += 32) v += tmp[r * K + idx];
                    v = warp_sum(v);
                    if (lane == 0) add_sig[idx] = v;
                } else {
                    const int r = idx - K;
                    for (int b = lane; b < K; b += 32) v += tmp[r * K + b];
                    v = warp_sum(v);
                    if (lane == 0) add_lam[r] = v;
                }
            }
        }
    }
    __syncthreads();

    // ---- gradients: Jacobians (entmc_vbmc.py:114-130, variational_optimization.py:1522-1548) and dF -----
    if (anyg) {
        const double *be = a.raw + rl.ent(), *bg = a.raw + rl.gp();
        const double *sigma = prm + a.lay.sigma(), *lambd = prm + a.lay.lambd(), *mu = prm + a.lay.mu();
        const int jac = a.f.jacobian;
        const int n0 = a.f.grad[0] ? K * D : 0, n1 = a.f.grad[1] ? K : 0, n2 = a.f.grad[2] ? D : 0,
                  n3 = a.f.grad[3] ? K : 0;
        for (int e = tid; e < n0 + n1 + n2 + n3; e += nt) {
            double gh, gg, add = 0.0;
            if (e < n0) {
                gh = be[rl.o_mu() + e], gg = bg[rl.o_mu() + e];
                if (bounds && n_mu) {  // d(bound loss)/d mu, recomputed element-wise
                    const double x = mu[e], lo = a.lb[e], hi = a.ub[e], ell = (hi - lo) * a.tol_con;
                    const double viol = x < lo ? x - lo : (x > hi ? x - hi : 0.0);
                    if (viol != 0.0) add = viol / (ell * ell);
                }
            } else if (e < n0 + n1) {
                const int k = e - n0;
                const double sc = jac ? sigma[k] : 1.0;
                gh = be[rl.o_sig() + k] * sc, gg = bg[rl.o_sig() + k] * sc;
                add = add_sig[k];
            } else if (e < n0 + n1 + n2) {
                const int d = e - n0 - n1;
                const double sc = jac ? lambd[d] : 1.0;
                gh = be[rl.o_lam() + d] * sc, gg = bg[rl.o_lam() + d] * sc;
                add = add_lam[d];
            } else {
                const int k = e - n0 - n1 - n2;
                gh = be[rl.o_w() + k], gg = bg[rl.o_w() + k];
                const double ek = jac || bounds ? exp(eta[k]) : 0.0;
                if (jac) {  // row k of J_w @ g
                    gh = ek / es * gh - ek / (es * es) * dote;
                    gg = ek / es * gg - ek / (es * es) * dotg;
                }
                add = add_eta[k];
                if (bounds && a.f.optimize[3]) {  // weight penalty through the softmax Jacobian (:1221-1229)
                    const double g = w[k] < a.w_thr ? a.w_pen : 0.0;
                    add += ek / es * g - ek / (es * es) * dotp;
                }
            }
            dH[e] = gh;
            dG[e] = gg;
            dF[e] = -gg - gh + add;  // :1171-1173, :1200, :1227-1229
        }
    }
    if (tid == 0) {
        const double H = a.raw[0], G = a.raw[1];
        const double F = -G - H + Lb + Lp;
        out[0] = F;
        out[1] = G;
        out[2] = H;
        out[3] = 0.0;
        out[4] = 0.0;
        out[5] = Lb;
        out[6] = Lp;
        out[7] = isfinite(F) ? 0.0 : 1.0;
    }
}

// per-hyper-sample Jacobians for avg_flag == 0: CTA s -> out_s[s] = [G_s | dG_s (P)]
__global__ void __launch_bounds__(256)
gps_finalize_kernel(const double *__restrict__ prm, ParamLayout lay, RawLayout rl, EvalFlags f,
                    const double *__restrict__ gps, double *__restrict__ out_s, int Pfull) {
    const int K = lay.K, s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    __shared__ double scratch[40];
    const double *blk = gps + (size_t)s * (1 + rl.block()) + 1;
    double *dst = out_s + (size_t)s * (1 + Pfull);
    double es = 0.0, dot = 0.0;
    if (f.grad[3] && f.jacobian) {
        for (int k = tid; k < K; k += nt) {
            const double ek = exp(prm[lay.eta() + k]);
            es += ek;
            dot += ek * blk[rl.o_w() + k];
        }
        es = block_sum(es, scratch);
        dot = block_sum(dot, scratch);
    }
    if (tid == 0) dst[0] = blk[-1];
    pack_block(blk, rl, prm, lay, f.grad, f.jacobian, es, dot, 1.0, dst + 1, false);
}

}  // namespace

static size_t finalize_smem(int D, int K) { return sizeof(double) * ((size_t)K * D + 2 * K + D); }

static ReduceArgs make_reduce_args(Ctx *c, int D, int K, const EvalFlags &f, const EntmcPlan *plan, int64_t Ns_glob,
                                   int s_begin, int s_step, int S_glob, double *d_raw) {
    ReduceArgs a{};
    const int DP = pad_dim(D);
    a.lay = ParamLayout{D, DP, K};
    a.rl = RawLayout{D, K};
    a.f = f;
    a.entpart = c->d_entpart;
    a.crec = c->d_crec;
    a.ent_stride = entpart_stride(DP, K);
    if (plan) {
        a.slabs = plan->slabs;
        a.Ns_glob = (double)Ns_glob;
        a.draws_local = 2.0 * (double)plan->half;
    }
    a.gppart = c->d_gppart;
    a.hyp = c->d_hyp;
    a.hs = hyp_stride(DP);
    a.s_begin = s_begin;
    a.s_step = s_step;
    a.S = c->S;
    a.S_glob = S_glob;
    a.mean_kind = c->mean_kind;
    a.gps = c->d_gps;
    a.raw = d_raw;
    return a;
}

// records -> (optionally) raw vector.  With assemble == false the caller fuses the assembly into
// finalize_launch (single-GPU fast path: one launch less).
int reduce_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const EntmcPlan *plan,
                  int64_t Ns_glob, int s_begin, int s_step, int S_glob, double *d_raw, bool assemble) {
    const int DP = pad_dim(D);
    const bool ent_on = f.have_ent && f.use_ent_mc;
    if (ent_on) VBMC_TRY(ensure(&c->d_crec, &c->crec_cap, (size_t)K * entpart_stride(DP, K)));
    ReduceArgs a = make_reduce_args(c, D, K, f, plan, Ns_glob, s_begin, s_step, S_glob, d_raw);
    c->red_args_valid = true;
    c->red_plan_slabs = a.slabs, c->red_Ns_glob = a.Ns_glob, c->red_draws_local = a.draws_local;
    c->red_s_begin = s_begin, c->red_s_step = s_step, c->red_S_glob = S_glob;
    const int S_local = f.have_gp ? (c->S - s_begin + s_step - 1) / s_step : 0;
    const int grid = (ent_on ? K : 0) + (S_local > 0 ? S_local : 0);
    if (grid > 0) {
        reduce_kernel<<<grid, 128, 0, c->stream>>>(d_params, a);
        VBMC_CUDA_CHECK(cudaGetLastError());
        c->launches++;
    }
    if (assemble) {
        FinalArgs fa{};
        fa.lay = a.lay, fa.rl = a.rl, fa.f = f;
        fa.do_assemble = 1, fa.do_finalize = 0;
        fa.red = a;
        finalize_kernel<<<1, 1024, finalize_smem(D, K), c->stream>>>(d_params, fa);
        VBMC_CUDA_CHECK(cudaGetLastError());
        c->launches++;
    }
    return VBMC_OK;
}

int finalize_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const double *d_raw,
                    double *d_out, bool assemble_first) {
    FinalArgs a{};
    a.lay = ParamLayout{D, pad_dim(D), K};
    a.rl = RawLayout{D, K};
    a.f = f;
    a.do_assemble = assemble_first ? 1 : 0;
    a.do_finalize = 1;
    if (assemble_first) {
        VBMC_REQUIRE(c->red_args_valid, VBMC_ERR_STATE, "finalize: no reduce stage to assemble from");
        a.red = make_reduce_args(c, D, K, f, nullptr, 0, c->red_s_begin, c->red_s_step, c->red_S_glob,
                                 const_cast<double *>(d_raw));
        a.red.slabs = c->red_plan_slabs, a.red.Ns_glob = c->red_Ns_glob, a.red.draws_local = c->red_draws_local;
    }
    a.raw = d_raw;
    a.lb = c->d_lb;
    a.ub = c->d_ub;
    a.n_bnd = c->n_bnd;
    a.tol_con = c->tol_con;
    a.w_thr = c->w_thr;
    a.w_pen = c->w_pen;
    a.out = d_out;
    a.Pfull = a.rl.block();
    const size_t fsmem = finalize_smem(D, K);
    if (fsmem > 48 * 1024)
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    finalize_kernel<<<1, 1024, fsmem, c->stream>>>(d_params, a);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

int gps_finalize_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, double *d_out_s) {
    ParamLayout lay{D, pad_dim(D), K};
    RawLayout rl{D, K};
    gps_finalize_kernel<<<c->S, 256, 0, c->stream>>>(d_params, lay, rl, f, c->d_gps, d_out_s, rl.block());
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace vbmc

// finalize.cu -- fixed-order reduction of the kernel outputs into the RAW vector, and the final
// assembly (Jacobians, soft-bound loss, weight penalty) of F and dF.
//
//   raw_kernel   : entmc CTA records + per-(s,k) log-joint terms  ->  raw = [H, G, .., .. | ent | gp]
//                  (raw = pre-Jacobian sums, already scaled by the GLOBAL 1/Ns and 1/S, so a SUM
//                   all-reduce over ranks yields the single-GPU value)
//   final_kernel : raw -> out = [F, G, H, .., | dF | dH | dG]
//                  log/softmax Jacobians  entmc_vbmc.py:114-130, variational_optimization.py:1522-1548
//                  soft bounds            variational_optimization.py:503-657 (incl. the row-major
//                                         reshape of the column-major ln-scale block at :584-586)
//                  weight penalty         variational_optimization.py:1212-1229
// Both are latency-bound O(S K D) passes, so both are spread over many small CTAs with ONE output entry
// per thread (element-wise entries) or per warp (entries that are sums over components / samples / slabs:
// lanes stride over the summed index, butterfly tree => fixed summation order, bitwise reproducible).
// Small global scalars (softmax sums, penalty) are recomputed by every CTA instead of being exchanged.
#include "common.cuh"

namespace vbmc {
namespace {

constexpr double kLog2Pi = 1.8378770664093454836;

struct RawArgs {
    ParamLayout lay;
    RawLayout rl;
    EvalFlags f;
    // entmc records of ent_stride doubles.  chunk == 0: component j owns records [j*slabs, (j+1)*slabs);
    // chunk > 0 (warp-autonomous kernel): CTA c covers pairs [c*chunk, (c+1)*chunk) of the flattened
    // (component, pair) space and writes record c*maxseg + (j - first component of c) for component j.
    const double *entpart;
    int slabs, ent_stride, maxseg;
    long long chunk, half;
    double Ns_glob;      // draws per component over all ranks
    double draws_local;  // draws per component on this rank
    // log joint: gps[s] = [G_s | mu | sigma | lambda | w] per-sample raw block, lamc[s][k][d]
    double *gps;
    const double *lamc;
    int gps_stride, s_begin, s_step, S, S_glob, S_local;
    double *raw;  // global raw vector (the entlb kernels may already have written H and the ent block)
    int n_warp_entries, n_thread_entries;
};

// sum over the slab records of component j of field f (fixed order)
__device__ __forceinline__ double slab_sum(const RawArgs &a, int j, int f) {
    double v = 0.0;
    if (a.chunk > 0) {
        const long long lo = (long long)j * a.half, hi = lo + a.half - 1;
        const int c0 = (int)(lo / a.chunk), c1 = (int)(hi / a.chunk);
#pragma unroll 4
        for (int c = c0; c <= c1; ++c) {
            const int seg = j - (int)(((long long)c * a.chunk) / a.half);
            v += a.entpart[((size_t)c * a.maxseg + seg) * a.ent_stride + f];
        }
        return v;
    }
    const double *rec = a.entpart + (size_t)j * a.slabs * a.ent_stride + f;
#pragma unroll 8
    for (int s = 0; s < a.slabs; ++s) v += rec[(size_t)s * a.ent_stride];
    return v;
}

__global__ void __launch_bounds__(256) raw_kernel(const double *__restrict__ prm, RawArgs a) {
    const int D = a.lay.D, DP = a.lay.DP, K = a.lay.K;
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const double *sigma = prm + a.lay.sigma(), *lambd = prm + a.lay.lambd(), *w = prm + a.lay.w();
    const RawLayout rl = a.rl;
    double *raw = a.raw;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const bool ent_mc = a.f.have_ent && a.f.use_ent_mc;
    const bool ent_lb = a.f.have_ent && !a.f.use_ent_mc;
    const double inv_ns = ent_mc ? 1.0 / a.Ns_glob : 0.0, inv_S = 1.0 / (double)a.S_glob;
    const int st = a.ent_stride;
    double *ent = raw + rl.ent(), *gpb = raw + rl.gp();

    if (gw < a.n_warp_entries) {
        // ------------------------------------------------------------ one WARP per summed entry
        // order: [H | ent sigma (K) | ent lambda (D) | ent w (K) | G | gp lambda (D) | per-sample (S_local x (D+1))]
        int idx = gw;
        if (idx == 0) {  // H = -sum_j w_j mean log q  (entmc_vbmc.py:80)
            if (ent_mc) {
                double sl = 0.0;
                for (int d = lane; d < D; d += 32) sl += log(lambd[d]);
                sl = warp_sum(sl);
                double h = 0.0;
                for (int j = lane; j < K; j += 32) {
                    const double hs = slab_sum(a, j, 0) + a.draws_local * (-0.5 * D * kLog2Pi - sl - D * log(sigma[j]));
                    h -= w[j] * hs;
                }
                h = warp_sum(h);
                if (lane == 0) raw[0] = h * inv_ns;
            } else if (!ent_lb && lane == 0) {
                raw[0] = 0.0;
            }
            if (lane == 0) raw[2] = raw[3] = 0.0;
            return;
        }
        idx -= 1;
        if (idx < K) {  // d/dsigma_j (:102-103): sum over dimensions of Be
            if (ent_mc && anyg) {
                double v = 0.0;
                for (int d = lane; d < D; d += 32) v += slab_sum(a, idx, 1 + DP + d);
                v = warp_sum(v);
                if (lane == 0) ent[rl.o_sig() + idx] = w[idx] * v * inv_ns / sigma[idx];
            } else if (!ent_lb && lane == 0) {
                ent[rl.o_sig() + idx] = 0.0;
            }
            return;
        }
        idx -= K;
        if (idx < D) {  // d/dlambda_d (:106-108): sum over components
            if (ent_mc && anyg) {
                double v = 0.0;
                for (int j = lane; j < K; j += 32) v += w[j] * slab_sum(a, j, 1 + DP + idx);
                v = warp_sum(v);
                if (lane == 0) ent[rl.o_lam() + idx] = v * inv_ns / lambd[idx];
            } else if (!ent_lb && lane == 0) {
                ent[rl.o_lam() + idx] = 0.0;
            }
            return;
        }
        idx -= D;
        if (idx < K) {  // d/dw_k (:111-112): -E_k[log q] - sum_j w_j E_j[N_k / q]
            if (ent_mc && anyg) {
                double v = 0.0;
                if (a.f.grad[3])
                    for (int j = lane; j < K; j += 32) v += w[j] * slab_sum(a, j, 1 + 2 * DP + idx);
                double sl = 0.0;
                for (int d = lane; d < D; d += 32) sl += log(lambd[d]);
                sl = warp_sum(sl);
                v = warp_sum(v);
                if (lane == 0) {
                    const double hs = slab_sum(a, idx, 0) + a.draws_local * (-0.5 * D * kLog2Pi - sl - D * log(sigma[idx]));
                    ent[rl.o_w() + idx] = -(hs + v) * inv_ns;
                }
            } else if (!ent_lb && lane == 0) {
                ent[rl.o_w() + idx] = 0.0;
            }
            return;
        }
        idx -= K;
        if (idx == 0) {  // G = mean_s sum_k w_k I_sk  (:1425, :1581)
            double v = 0.0;
            if (a.f.have_gp)
                for (int i = lane; i < a.S_local * K; i += 32) {
                    const int s = a.s_begin + (i / K) * a.s_step, k = i % K;
                    v += w[k] * a.gps[(size_t)s * a.gps_stride + 1 + rl.o_w() + k];
                }
            v = warp_sum(v);
            if (lane == 0) raw[1] = v * inv_S;
            return;
        }
        idx -= 1;
        if (idx < D) {  // gp d/dlambda_d: mean_s sum_k lamc  (:1452-1462, :1596)
            double v = 0.0;
            if (a.f.have_gp && anyg)
                for (int i = lane; i < a.S_local * K; i += 32) {
                    const int s = a.s_begin + (i / K) * a.s_step, k = i % K;
                    v += a.lamc[((size_t)s * K + k) * D + idx];
                }
            v = warp_sum(v);
            if (lane == 0) gpb[rl.o_lam() + idx] = v * inv_S;
            return;
        }
        idx -= D;
        if (a.f.have_gp) {  // per-sample G_s and lambda block (consumers: variance path, avg_flag = 0, I_sk)
            const int si = idx / (D + 1), col = idx - si * (D + 1), s = a.s_begin + si * a.s_step;
            double *gs = a.gps + (size_t)s * a.gps_stride;
            double v = 0.0;
            if (col == 0)
                for (int k = lane; k < K; k += 32) v += w[k] * gs[1 + rl.o_w() + k];
            else if (anyg)
                for (int k = lane; k < K; k += 32) v += a.lamc[((size_t)s * K + k) * D + col - 1];
            v = warp_sum(v);
            if (lane == 0) gs[col == 0 ? 0 : 1 + rl.o_lam() + col - 1] = v;
        }
        return;
    }
    // ---------------------------------------------------------------- one THREAD per element-wise entry
    // order: [ent mu (K*D) | gp mu (K*D) | gp sigma (K) | gp w (K)]
    int e = (gw - a.n_warp_entries) * 32 + lane;
    if (e >= a.n_thread_entries) return;
    if (e < K * D) {  // d/dmu_j (:98)
        if (ent_mc && anyg) {
            const int j = e / D, d = e - j * D;
            ent[rl.o_mu() + e] = w[j] * slab_sum(a, j, 1 + d) * inv_ns / lambd[d];
        } else if (!ent_lb) {
            ent[rl.o_mu() + e] = 0.0;
        }
        return;
    }
    e -= K * D;
    // gp blocks: average over this rank's hyper-samples (:1578-1596)
    const int off = e < K * D ? rl.o_mu() + e : (e < K * D + K ? rl.o_sig() + (e - K * D) : rl.o_w() + (e - K * D - K));
    double v = 0.0;
    if (a.f.have_gp && anyg) {
#pragma unroll 4
        for (int s = a.s_begin; s < a.S; s += a.s_step) v += a.gps[(size_t)s * a.gps_stride + 1 + off];
    }
    gpb[off] = v * inv_S;
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
    const double *raw;
    const double *lb, *ub;
    int n_bnd;
    double tol_con, w_thr, w_pen;
    double *out;
    int Pfull;
};

// violation derivative of extended-theta entry e (and its loss), variational_optimization.py:639-653
__device__ __forceinline__ double bound_dy(const FinalArgs &a, const double *prm, int e, int n_mu, int n_sc,
                                           double *loss) {
    const int D = a.lay.D;
    double x;
    if (e < n_mu)
        x = prm[a.lay.mu() + e];
    else if (e < n_mu + n_sc) {
        const int i = e - n_mu, k = i / D, d = i - k * D;  // column-major (D,K) ravel (:557-562)
        x = prm[a.lay.lnlam_b() + d] + prm[a.lay.lnsig_b() + k];
    } else
        x = prm[a.lay.eta_b() + e - n_mu - n_sc];
    const double lo = a.lb[e], hi = a.ub[e];
    const double viol = x < lo ? x - lo : (x > hi ? x - hi : 0.0);
    if (viol == 0.0) return 0.0;
    const double ell = (hi - lo) * a.tol_con, r = viol / ell;
    if (loss) *loss += 0.5 * r * r;
    return r / ell;
}

__global__ void __launch_bounds__(256) final_kernel(const double *__restrict__ prm, FinalArgs a) {
    const int D = a.lay.D, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    __shared__ double scratch[40];
    __shared__ double ssm[5];  // es, dote, dotg, dotp, Lp
    extern __shared__ double tmp[];  // [K*D] d(bound loss)/d(ln-scale entry), recomputed by every CTA
    const RawLayout rl = a.rl;
    const double *raw = a.raw;
    const double *eta = prm + a.lay.eta(), *w = prm + a.lay.w();
    double *out = a.out;
    double *dF = out + kOutHead, *dH = dF + a.Pfull, *dG = dH + a.Pfull;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const bool bounds = a.f.use_bounds && a.n_bnd > 0;
    const bool pen = bounds && a.f.optimize[3];
    const int n_mu = a.f.optimize[0] ? K * D : 0, n_sc = K * D, n_eta = a.f.optimize[3] ? K : 0;

    // ---- softmax pieces (every CTA, first warp): es = sum exp(eta), <exp(eta), gw>, penalty terms
    const bool need_sm = (anyg && a.f.grad[3] && a.f.jacobian) || pen;
    if (tid < 32 && need_sm) {
        double es = 0.0, dote = 0.0, dotg = 0.0, dotp = 0.0, Lp = 0.0;
        for (int k = tid; k < K; k += 32) {
            const double ek = exp(eta[k]);
            es += ek;
            dote += ek * raw[rl.ent() + rl.o_w() + k];
            dotg += ek * raw[rl.gp() + rl.o_w() + k];
            if (pen) {
                const double wk = w[k];
                Lp += (wk < a.w_thr ? wk : a.w_thr) * a.w_pen;  // :1213-1219
                dotp += ek * (wk < a.w_thr ? a.w_pen : 0.0);
            }
        }
        es = warp_sum(es), dote = warp_sum(dote), dotg = warp_sum(dotg), dotp = warp_sum(dotp), Lp = warp_sum(Lp);
        if (tid == 0) ssm[0] = es, ssm[1] = dote, ssm[2] = dotg, ssm[3] = dotp, ssm[4] = Lp;
    }
    // ln-scale block of the bound loss: all K*D entries in one parallel pass (independent loads), the
    // row / column sums below then run out of shared memory.  CTA 0 also collects the loss itself.
    double Lb = 0.0;
    if (bounds) {
        for (int i = tid; i < n_sc; i += nt) tmp[i] = bound_dy(a, prm, n_mu + i, n_mu, n_sc, &Lb);
        if (blockIdx.x == 0) {
            for (int e = tid; e < n_mu; e += nt) bound_dy(a, prm, e, n_mu, n_sc, &Lb);
            for (int e = n_mu + n_sc + tid; e < n_mu + n_sc + n_eta; e += nt) bound_dy(a, prm, e, n_mu, n_sc, &Lb);
        }
    }
    __syncthreads();
    const double es = ssm[0], dote = ssm[1], dotg = ssm[2], dotp = ssm[3];

    // ---- one output entry per thread: Jacobians, bound-loss and penalty gradients, dF ---------------
    if (anyg) {
        const double *be = raw + rl.ent(), *bg = raw + rl.gp();
        const double *sigma = prm + a.lay.sigma(), *lambd = prm + a.lay.lambd();
        const int jac = a.f.jacobian;
        const int n0 = a.f.grad[0] ? K * D : 0, n1 = a.f.grad[1] ? K : 0, n2 = a.f.grad[2] ? D : 0,
                  n3 = a.f.grad[3] ? K : 0;
        const int e = blockIdx.x * nt + tid;
        if (e < n0 + n1 + n2 + n3) {
            double gh, gg, add = 0.0;
            if (e < n0) {
                gh = be[rl.o_mu() + e], gg = bg[rl.o_mu() + e];
                if (bounds && n_mu) add = bound_dy(a, prm, e, n_mu, n_sc, nullptr);
            } else if (e < n0 + n1) {
                const int k = e - n0;
                const double sc = jac ? sigma[k] : 1.0;
                gh = be[rl.o_sig() + k] * sc, gg = bg[rl.o_sig() + k] * sc;
                // the reference reshapes the ln-scale gradient ROW-major to (D, K): dls[r][b] = dy[r*K + b]
                // (:584-586); sigma_b gets the column sum over r
                if (bounds)
                    for (int r = 0; r < D; ++r) add += tmp[r * K + k];
            } else if (e < n0 + n1 + n2) {
                const int d = e - n0 - n1;
                const double sc = jac ? lambd[d] : 1.0;
                gh = be[rl.o_lam() + d] * sc, gg = bg[rl.o_lam() + d] * sc;
                if (bounds)  // ... and lambda_r the row sum over b
                    for (int b = 0; b < K; ++b) add += tmp[d * K + b];
            } else {
                const int k = e - n0 - n1 - n2;
                gh = be[rl.o_w() + k], gg = bg[rl.o_w() + k];
                const double ek = (jac || pen) ? exp(eta[k]) : 0.0;
                if (jac) {  // row k of J_w @ g
                    gh = ek / es * gh - ek / (es * es) * dote;
                    gg = ek / es * gg - ek / (es * es) * dotg;
                }
                if (bounds && n_eta) add = bound_dy(a, prm, n_mu + n_sc + k, n_mu, n_sc, nullptr);
                if (pen) {  // weight penalty through the softmax Jacobian (:1221-1229)
                    const double g = w[k] < a.w_thr ? a.w_pen : 0.0;
                    add += ek / es * g - ek / (es * es) * dotp;
                }
            }
            dH[e] = gh;
            dG[e] = gg;
            dF[e] = -gg - gh + add;  // :1171-1173, :1200, :1227-1229
        }
    }

    // ---- CTA 0: the scalars ----------------------------------------------------------------------------
    if (blockIdx.x == 0) {
        Lb = block_sum(Lb, scratch);
        if (tid == 0) {
            const double H = raw[0], G = raw[1], Lp = pen ? ssm[4] : 0.0;
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

// records / per-sample terms -> raw vector (device pointer d_raw)
int reduce_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const EntmcPlan *plan,
                  int64_t Ns_glob, int s_begin, int s_step, int S_glob, double *d_raw) {
    RawArgs a{};
    const int DP = pad_dim(D);
    a.lay = ParamLayout{D, DP, K};
    a.rl = RawLayout{D, K};
    a.f = f;
    a.entpart = c->d_entpart;
    a.ent_stride = entpart_stride(DP, K);
    if (plan) {
        a.slabs = plan->slabs;
        a.chunk = (plan->variant == ENTMC_WARP || plan->variant == ENTMC_TC) ? plan->chunk : 0;
        a.maxseg = plan->maxseg;
        a.half = plan->half;
        a.Ns_glob = (double)Ns_glob;
        a.draws_local = 2.0 * (double)plan->half;
    }
    a.gps = c->d_gps;
    a.lamc = c->d_lamc;
    a.gps_stride = 1 + a.rl.block();
    a.s_begin = s_begin, a.s_step = s_step, a.S = c->S, a.S_glob = S_glob;
    a.S_local = f.have_gp ? (c->S - s_begin + s_step - 1) / s_step : 0;
    if (a.S_local < 0) a.S_local = 0;
    a.raw = d_raw;
    a.n_warp_entries = 1 + K + D + K + 1 + D + a.S_local * (D + 1);
    a.n_thread_entries = 2 * K * D + 2 * K;
    const int warps = a.n_warp_entries + (a.n_thread_entries + 31) / 32;
    raw_kernel<<<(warps + 7) / 8, 256, 0, c->stream>>>(d_params, a);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

int finalize_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const double *d_raw,
                    double *d_out) {
    FinalArgs a{};
    a.lay = ParamLayout{D, pad_dim(D), K};
    a.rl = RawLayout{D, K};
    a.f = f;
    a.raw = d_raw;
    a.lb = c->d_lb;
    a.ub = c->d_ub;
    a.n_bnd = f.use_bounds ? c->n_bnd : 0;
    a.tol_con = c->tol_con;
    a.w_thr = c->w_thr;
    a.w_pen = c->w_pen;
    a.out = d_out;
    a.Pfull = a.rl.block();
    const int P = (f.grad[0] ? K * D : 0) + (f.grad[1] ? K : 0) + (f.grad[2] ? D : 0) + (f.grad[3] ? K : 0);
    const int grid = P > 0 ? (P + 255) / 256 : 1;
    const size_t smem = (size_t)K * D * sizeof(double);
    VBMC_REQUIRE(smem <= 200 * 1024, VBMC_ERR_UNSUPPORTED, "finalize: D*K too large");
    if (smem > c->finalize_smem_set && smem > 48 * 1024) {
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->finalize_smem_set = smem;
    }
    final_kernel<<<grid, 256, smem, c->stream>>>(d_params, a);
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

// finalize.cu -- fixed-order reduction of the kernel outputs into the RAW vector, and the final
// assembly (Jacobians, soft-bound loss, weight penalty) of F and dF.
//
//   finalize_kernel (ONE CTA, 1024 threads), two phases that can run in one launch or in two:
//     assemble : entmc CTA records + per-(s,k) log-joint terms  ->  raw = [H, G, .., .. | ent | gp]
//                (raw = pre-Jacobian sums, already scaled by the GLOBAL 1/Ns and 1/S, so a SUM
//                 all-reduce over ranks yields the single-GPU value)
//     finalize : raw -> out = [F, G, H, .., | dF | dH | dG]
//                log/softmax Jacobians  entmc_vbmc.py:114-130, variational_optimization.py:1522-1548
//                soft bounds            variational_optimization.py:503-657 (incl. the row-major
//                                       reshape of the column-major ln-scale block at :584-586)
//                weight penalty         variational_optimization.py:1212-1229
//   On one GPU both phases share a launch; with W ranks the all-reduce sits between them.
//   Latency is what matters here (O(S K D) work): the parameter block, the bounds and the raw vector
//   are staged into shared memory once and every later phase runs out of shared memory; cross-element
//   sums use one warp per output entry with a butterfly tree => deterministic order.
#include "common.cuh"

namespace vbmc {
namespace {

constexpr double kLog2Pi = 1.8378770664093454836;

struct AsmArgs {
    // entmc CTA records: component j owns records [j*slabs, (j+1)*slabs)
    const double *entpart;
    int slabs, ent_stride;
    double Ns_glob;      // draws per component over all ranks
    double draws_local;  // draws per component on this rank
    // log joint: gps[s] = [G_s | mu | sigma | lambda | w] per-sample raw block, lamc[s][k][d]
    double *gps;
    const double *lamc;
    int gps_stride, s_begin, s_step, S, S_glob;
    double *raw_out;  // global raw vector
};

struct FinalArgs {
    ParamLayout lay;
    RawLayout rl;
    EvalFlags f;
    int do_assemble, do_finalize, stage_bounds;
    AsmArgs as;
    const double *raw_in;  // global raw vector (all-reduced) when !do_assemble; entlb output when ent_lb
    const double *lb, *ub;
    int n_bnd;
    double tol_con, w_thr, w_pen;
    double *out;
    int Pfull;
};

struct Smem {
    double *prm, *raw, *tmp, *add_sig, *add_lam, *add_eta, *part, *gsig, *sgs, *lb, *ub;
};

__device__ __forceinline__ Smem carve(double *base, const FinalArgs &a) {
    const int D = a.lay.D, K = a.lay.K;
    Smem m;
    m.prm = base;
    m.raw = m.prm + a.lay.total();
    m.tmp = m.raw + a.rl.total();
    m.add_sig = m.tmp + K * D;
    m.add_lam = m.add_sig + K;
    m.add_eta = m.add_lam + D;
    m.part = m.add_eta + K;                       // [32][ent_stride + 1]
    m.gsig = m.part + 32 * (a.as.ent_stride + 1);  // [K]
    m.sgs = m.gsig + K;                            // [S][D + 1]
    m.lb = m.sgs + a.as.S * (D + 1);
    m.ub = m.lb + (a.stage_bounds ? a.n_bnd : 0);
    return m;
}

// ---- assemble: records -> raw (in shared memory, then copied to global) -------------------------------
__device__ void assemble_raw(const FinalArgs &a, const Smem &m, double *scratch) {
    const int D = a.lay.D, DP = a.lay.DP, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const AsmArgs &as = a.as;
    const RawLayout rl = a.rl;
    const double *sigma = m.prm + a.lay.sigma(), *lambd = m.prm + a.lay.lambd(), *w = m.prm + a.lay.w();
    double *raw = m.raw;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const bool ent_mc = a.f.have_ent && a.f.use_ent_mc;
    const bool ent_lb = a.f.have_ent && !a.f.use_ent_mc;  // the entlb kernels wrote H and the block to global raw

    for (int e = tid; e < rl.total(); e += nt) {
        const bool in_ent = e >= rl.ent() && e < rl.gp();
        raw[e] = (ent_lb && (e == 0 || in_ent)) ? a.raw_in[e] : 0.0;
    }
    for (int k = tid; k < K; k += nt) m.gsig[k] = 0.0;
    __syncthreads();

    if (ent_mc) {
        const double inv_ns = 1.0 / as.Ns_glob;
        const int st = as.ent_stride;
        // fields the producer actually wrote: [sum log q] always, [A | Be] with any gradient, [racc] with d/dw
        const int st_eff = !anyg ? 1 : (a.f.grad[3] ? st : 1 + 2 * DP);
        double *ent = raw + rl.ent();
        double sumlnl = 0.0;
        for (int d = lane; d < D; d += 32) sumlnl += log(lambd[d]);
        sumlnl = warp_sum(sumlnl);  // every warp computes the same value
        double hpart = 0.0;         // lane 0 of each warp
        for (int f = tid; f < 32 * (st + 1); f += nt) m.part[f] = 0.0;
        __syncthreads();
        for (int f0 = 0; f0 < st_eff; f0 += 32) {
            const int f = f0 + lane;
            const bool fin = f < st_eff;
            const bool is_be = fin && f >= 1 + DP && f < 1 + DP + D;
            double colacc = 0.0;  // sum_j w_j v_j[f] over this warp's components
            for (int j = wid; j < K; j += nw) {
                double v = 0.0;
                if (fin) {
                    const double *rec = as.entpart + (size_t)j * as.slabs * st + f;
#pragma unroll 4
                    for (int s = 0; s < as.slabs; ++s) v += rec[(size_t)s * st];
                }
                const double wj = w[j];
                if (f == 0) {
                    // + (draws of this rank) * log(nconst / sigma_j^D)   (entmc_vbmc.py:53-56,77)
                    const double hs = v + as.draws_local * (-0.5 * D * kLog2Pi - sumlnl - D * log(sigma[j]));
                    hpart -= wj * hs * inv_ns;        // :80
                    ent[rl.o_w() + j] = -hs * inv_ns;  // :111 (cross term added below)
                } else if (fin && f < 1 + D) {
                    ent[rl.o_mu() + j * D + (f - 1)] = wj * v * inv_ns / lambd[f - 1];  // :98
                }
                if (anyg && f0 < 1 + DP + D && f0 + 32 > 1 + DP) {  // chunk holds Be fields: d/dsigma_j (:102-103)
                    const double sb = warp_sum(is_be ? v : 0.0);
                    if (lane == 0) m.gsig[j] += sb;
                }
                if (fin && f >= 1 + DP) colacc += wj * v;
            }
            if (fin) m.part[wid * (st + 1) + f] = colacc;
        }
        if (lane == 0) m.part[wid * (st + 1) + st] = hpart;
        __syncthreads();
        // cross-warp sums in fixed order
        for (int f = tid; f <= st; f += nt) {
            double v = 0.0;
            for (int ww = 0; ww < nw; ++ww) v += m.part[ww * (st + 1) + f];
            if (f == st)
                raw[0] = v;  // H
            else if (f >= 1 + DP && f < 1 + DP + D)
                ent[rl.o_lam() + (f - 1 - DP)] = v * inv_ns / lambd[f - 1 - DP];  // :106-108
            else if (f >= 1 + 2 * DP && a.f.grad[3])
                ent[rl.o_w() + (f - 1 - 2 * DP)] -= v * inv_ns;  // :112
        }
        for (int j = tid; j < K; j += nt) ent[rl.o_sig() + j] = w[j] * m.gsig[j] * inv_ns / sigma[j];
        __syncthreads();
    }

    if (a.f.have_gp) {
        // per-sample cross-component sums: G_s = sum_k w_k I_sk (:1425), lambda block = sum_k lamc (:1452-1462)
        const int blk = rl.block();
        int si = 0;
        for (int s = as.s_begin; s < as.S; s += as.s_step, ++si) {
            double *gs = as.gps + (size_t)s * as.gps_stride;
            for (int idx = wid; idx < D + 1; idx += nw) {
                double v = 0.0;
                if (idx == 0) {
                    for (int k = lane; k < K; k += 32) v += w[k] * gs[1 + rl.o_w() + k];
                } else if (anyg) {
                    for (int k = lane; k < K; k += 32) v += as.lamc[((size_t)s * K + k) * D + idx - 1];
                }
                v = warp_sum(v);
                if (lane == 0) {
                    m.sgs[si * (D + 1) + idx] = v;
                    if (idx == 0)
                        gs[0] = v;
                    else
                        gs[1 + rl.o_lam() + idx - 1] = v;  // complete the per-sample block for its consumers
                }
            }
        }
        const int S_local = si;
        __syncthreads();
        // average over hyper-samples (:1578-1596); this rank contributes its own s / S_glob
        const double inv_S = 1.0 / (double)as.S_glob;
        for (int e = tid; e < blk + 1; e += nt) {
            if (e > 0 && !anyg) break;
            double v = 0.0;
            const bool lam = e >= 1 + rl.o_lam() && e < 1 + rl.o_w();
            if (e == 0 || lam) {
                const int col = e == 0 ? 0 : e - rl.o_lam();
                for (int i = 0; i < S_local; ++i) v += m.sgs[i * (D + 1) + col];
            } else {
#pragma unroll 4
                for (int s = as.s_begin; s < as.S; s += as.s_step) v += as.gps[(size_t)s * as.gps_stride + e];
            }
            v *= inv_S;
            if (e == 0)
                raw[1] = v;
            else
                raw[rl.gp() + e - 1] = v;
        }
    }
    __syncthreads();
    for (int e = tid; e < rl.total(); e += nt) as.raw_out[e] = raw[e];
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


__global__ void __launch_bounds__(1024) finalize_kernel(const double *__restrict__ prm, FinalArgs a) {
    const int D = a.lay.D, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    __shared__ double scratch[40];
    extern __shared__ double fsm[];
    const Smem m = carve(fsm, a);

    // ---- stage: parameter block, (all-reduced) raw vector, bounds -> shared memory -------------------
    for (int i = tid; i < a.lay.total(); i += nt) m.prm[i] = prm[i];
    if (!a.do_assemble)
        for (int i = tid; i < a.rl.total(); i += nt) m.raw[i] = a.raw_in[i];
    const bool bounds = a.do_finalize && a.f.use_bounds && a.n_bnd > 0;
    if (bounds && a.stage_bounds)
        for (int i = tid; i < a.n_bnd; i += nt) m.lb[i] = a.lb[i], m.ub[i] = a.ub[i];
    for (int i = tid; i < 2 * K + D; i += nt) m.add_sig[i] = 0.0;  // add_sig | add_lam | add_eta are contiguous
    __syncthreads();
    if (a.do_assemble) assemble_raw(a, m, scratch);
    if (!a.do_finalize) return;

    const RawLayout rl = a.rl;
    const double *raw = m.raw;
    const double *lb = a.stage_bounds ? m.lb : a.lb, *ub = a.stage_bounds ? m.ub : a.ub;
    const double *eta = m.prm + a.lay.eta(), *w = m.prm + a.lay.w();
    double *out = a.out;
    double *dF = out + kOutHead, *dH = dF + a.Pfull, *dG = dH + a.Pfull;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];

    // ---- softmax pieces: es = sum exp(eta), <exp(eta), gw> for both blocks, and for the weight penalty
    double es = 0.0, dote = 0.0, dotg = 0.0, dotp = 0.0, Lp = 0.0;
    const bool need_sm = (anyg && a.f.grad[3] && a.f.jacobian) || (bounds && a.f.optimize[3]);
    if (need_sm) {
        for (int k = tid; k < K; k += nt) {
            const double ek = exp(eta[k]);
            es += ek;
            dote += ek * raw[rl.ent() + rl.o_w() + k];
            dotg += ek * raw[rl.gp() + rl.o_w() + k];
            if (bounds && a.f.optimize[3]) {
                const double wk = w[k];
                Lp += (wk < a.w_thr ? wk : a.w_thr) * a.w_pen;  // :1213-1219
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
        const double *lnsig = m.prm + a.lay.lnsig_b(), *lnlam = m.prm + a.lay.lnlam_b(), *etab = m.prm + a.lay.eta_b();
        const double *mu = m.prm + a.lay.mu();
        for (int e = tid; e < n_mu + n_sc + n_eta; e += nt) {
            double x;
            if (e < n_mu)
                x = mu[e];
            else if (e < n_mu + n_sc) {
                const int i = e - n_mu, k = i / D, d = i - k * D;  // column-major (D,K) ravel (:557-562)
                x = lnlam[d] + lnsig[k];
            } else
                x = etab[e - n_mu - n_sc];
            const double lo = lb[e], hi = ub[e];
            const double ell = (hi - lo) * a.tol_con;
            const double viol = x < lo ? x - lo : (x > hi ? x - hi : 0.0);
            double dy = 0.0;
            if (viol != 0.0) {
                const double r = viol / ell;
                Lb += 0.5 * r * r;
                dy = viol / (ell * ell);
            }
            // the mu part is re-derived element-wise in the final pass; keep the other two in shared memory
            if (e >= n_mu && e < n_mu + n_sc)
                m.tmp[e - n_mu] = dy;
            else if (e >= n_mu + n_sc)
                m.add_eta[e - n_mu - n_sc] = dy;
        }
        Lb = block_sum(Lb, scratch);  // (contains the barrier that publishes tmp / add_eta)
        if (anyg) {
            // the reference reshapes the ln-scale gradient ROW-major to (D, K): dls[r][b] = tmp[r*K + b]
            // (:584-586); sigma gets the column sums, lambda the row sums.  One warp per output entry.
            for (int idx = wid; idx < K + D; idx += nw) {
                double v = 0.0;
                if (idx < K) {
                    for (int r = lane; r < D; r += 32) v += m.tmp[r * K + idx];
                    v = warp_sum(v);
                    if (lane == 0) m.add_sig[idx] = v;
                } else {
                    const int r = idx - K;
                    for (int b = lane; b < K; b += 32) v += m.tmp[r * K + b];
                    v = warp_sum(v);
                    if (lane == 0) m.add_lam[r] = v;
                }
            }
        }
    }
    __syncthreads();

    // ---- gradients: Jacobians (entmc_vbmc.py:114-130, variational_optimization.py:1522-1548) and dF -----
    if (anyg) {
        const double *be = raw + rl.ent(), *bg = raw + rl.gp();
        const double *sigma = m.prm + a.lay.sigma(), *lambd = m.prm + a.lay.lambd(), *mu = m.prm + a.lay.mu();
        const int jac = a.f.jacobian;
        const int n0 = a.f.grad[0] ? K * D : 0, n1 = a.f.grad[1] ? K : 0, n2 = a.f.grad[2] ? D : 0,
                  n3 = a.f.grad[3] ? K : 0;
        for (int e = tid; e < n0 + n1 + n2 + n3; e += nt) {
            double gh, gg, add = 0.0;
            if (e < n0) {
                gh = be[rl.o_mu() + e], gg = bg[rl.o_mu() + e];
                if (bounds && n_mu) {  // d(bound loss)/d mu, recomputed element-wise
                    const double x = mu[e], lo = lb[e], hi = ub[e], ell = (hi - lo) * a.tol_con;
                    const double viol = x < lo ? x - lo : (x > hi ? x - hi : 0.0);
                    if (viol != 0.0) add = viol / (ell * ell);
                }
            } else if (e < n0 + n1) {
                const int k = e - n0;
                const double sc = jac ? sigma[k] : 1.0;
                gh = be[rl.o_sig() + k] * sc, gg = bg[rl.o_sig() + k] * sc;
                add = m.add_sig[k];
            } else if (e < n0 + n1 + n2) {
                const int d = e - n0 - n1;
                const double sc = jac ? lambd[d] : 1.0;
                gh = be[rl.o_lam() + d] * sc, gg = bg[rl.o_lam() + d] * sc;
                add = m.add_lam[d];
            } else {
                const int k = e - n0 - n1 - n2;
                gh = be[rl.o_w() + k], gg = bg[rl.o_w() + k];
                const double ek = jac || bounds ? exp(eta[k]) : 0.0;
                if (jac) {  // row k of J_w @ g
                    gh = ek / es * gh - ek / (es * es) * dote;
                    gg = ek / es * gg - ek / (es * es) * dotg;
                }
                add = m.add_eta[k];
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
        const double H = raw[0], G = raw[1];
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

static size_t finalize_smem(const FinalArgs &a, bool stage_bounds) {
    const int D = a.lay.D, K = a.lay.K;
    size_t n = (size_t)a.lay.total() + a.rl.total() + (size_t)K * D + 2 * K + D + 32 * (size_t)(a.as.ent_stride + 1) + K +
               (size_t)a.as.S * (D + 1);
    if (stage_bounds) n += 2 * (size_t)a.n_bnd;
    return n * sizeof(double);
}

static int launch_finalize(Ctx *c, const double *d_params, FinalArgs &a) {
    a.stage_bounds = 1;
    size_t smem = finalize_smem(a, true);
    if (smem > 200 * 1024) {
        a.stage_bounds = 0;
        smem = finalize_smem(a, false);
    }
    VBMC_REQUIRE(smem <= 220 * 1024, VBMC_ERR_UNSUPPORTED, "finalize: D*K too large for the shared-memory staging");
    if (smem > c->finalize_smem_set) {
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->finalize_smem_set = smem;
    }
    finalize_kernel<<<1, 1024, smem, c->stream>>>(d_params, a);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

static void fill_asm(Ctx *c, int D, int K, FinalArgs &a, double *d_raw) {
    const int DP = pad_dim(D);
    a.as.entpart = c->d_entpart;
    a.as.ent_stride = entpart_stride(DP, K);
    a.as.slabs = c->red_plan_slabs;
    a.as.Ns_glob = c->red_Ns_glob;
    a.as.draws_local = c->red_draws_local;
    a.as.gps = c->d_gps;
    a.as.lamc = c->d_lamc;
    a.as.gps_stride = 1 + RawLayout{D, K}.block();
    a.as.s_begin = c->red_s_begin;
    a.as.s_step = c->red_s_step;
    a.as.S = c->S;
    a.as.S_glob = c->red_S_glob;
    a.as.raw_out = d_raw;
}

// Record the shard description of this evaluation; with assemble == true also launch the
// assemble-only pass (multi-GPU: the all-reduce of d_raw follows).
int reduce_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const EntmcPlan *plan,
                  int64_t Ns_glob, int s_begin, int s_step, int S_glob, double *d_raw, bool assemble) {
    c->red_args_valid = true;
    c->red_plan_slabs = plan ? plan->slabs : 0;
    c->red_Ns_glob = plan ? (double)Ns_glob : 0.0;
    c->red_draws_local = plan ? 2.0 * (double)plan->half : 0.0;
    c->red_s_begin = s_begin, c->red_s_step = s_step, c->red_S_glob = S_glob;
    if (!assemble) return VBMC_OK;
    FinalArgs a{};
    a.lay = ParamLayout{D, pad_dim(D), K};
    a.rl = RawLayout{D, K};
    a.f = f;
    a.do_assemble = 1, a.do_finalize = 0;
    fill_asm(c, D, K, a, d_raw);
    a.raw_in = d_raw;
    return launch_finalize(c, d_params, a);
}

int finalize_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const double *d_raw,
                    double *d_out, bool assemble_first) {
    FinalArgs a{};
    a.lay = ParamLayout{D, pad_dim(D), K};
    a.rl = RawLayout{D, K};
    a.f = f;
    a.do_assemble = assemble_first ? 1 : 0;
    a.do_finalize = 1;
    if (assemble_first) VBMC_REQUIRE(c->red_args_valid, VBMC_ERR_STATE, "finalize: no partials to assemble from");
    fill_asm(c, D, K, a, const_cast<double *>(d_raw));
    a.raw_in = d_raw;
    a.lb = c->d_lb;
    a.ub = c->d_ub;
    a.n_bnd = f.use_bounds ? c->n_bnd : 0;
    a.tol_con = c->tol_con;
    a.w_thr = c->w_thr;
    a.w_pen = c->w_pen;
    a.out = d_out;
    a.Pfull = a.rl.block();
    return launch_finalize(c, d_params, a);
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

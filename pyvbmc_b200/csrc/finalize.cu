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
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

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
    ChunkMap cm;  // CTA c covers pairs [cm.start(c), cm.start(c + 1)) of the flattened (component, pair) space
    double Ns_glob;      // draws per component over all ranks
    double draws_local;  // draws per component on this rank
    // log joint: gps[s] = [G_s | mu | sigma | lambda | w] per-sample raw block, lamc[s][k][d]
    double *gps;
    const double *lamc;
    int gps_stride, s_begin, s_step, S, S_glob, S_local;
    double *raw;   // global raw vector (the entlb kernels may already have written H and the ent block)
    double *csum;  // [K][ent_stride] per-component sums of the entmc records (scratch between the two raw phases)
};

// sum over the slab records of component j of field f (fixed order)
__device__ __forceinline__ double slab_sum(const RawArgs &a, int j, int f) {
    double v = 0.0;
    if (a.chunk > 0) {
        // component j spans CTAs c0..c1 of the flattened space.  Every CTA after c0 STARTS inside component j, so j is
        // its segment 0; only c0 needs the segment index.
        const long long lo = (long long)j * a.half, hi = lo + a.half - 1;
        const int c0 = a.cm.cta_of(lo), c1 = a.cm.cta_of(hi);
        const int seg0 = j - (int)(a.cm.start(c0) / a.half);
        const double *rec = a.entpart + f;
        v = rec[((size_t)c0 * a.maxseg + seg0) * a.ent_stride];
        const size_t step = (size_t)a.maxseg * a.ent_stride;
        rec += (size_t)(c0 + 1) * step;
        // batches of 8 independent (predicated) loads: one L2 round trip per batch, fixed summation order
        for (int c = c0 + 1; c <= c1; c += 8, rec += 8 * step) {
            double t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = (c + i <= c1) ? rec[(size_t)i * step] : 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) v += t[i];
        }
        return v;
    }
    const double *rec = a.entpart + (size_t)j * a.slabs * a.ent_stride + f;
#pragma unroll 8
    for (int s = 0; s < a.slabs; ++s) v += rec[(size_t)s * a.ent_stride];
    return v;
}

// ---- phase 1: everything that reads producer outputs (entmc records, per-(s,k) log-joint terms) ---------
// gtid / GT: index and number of the threads of the whole cluster
__device__ __forceinline__ void raw_phase1(const double *__restrict__ prm, const RawArgs &a, int gtid, int GT) {
    const int D = a.lay.D, DP = a.lay.DP, K = a.lay.K;
    const int lane = gtid & 31, gw = gtid >> 5, GW = GT >> 5;
    const double *lambd = prm + a.lay.lambd(), *w = prm + a.lay.w();
    const RawLayout rl = a.rl;
    double *raw = a.raw;
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const bool ent_mc = a.f.have_ent && a.f.use_ent_mc;
    const bool ent_lb = a.f.have_ent && !a.f.use_ent_mc;
    const double inv_ns = ent_mc ? 1.0 / a.Ns_glob : 0.0, inv_S = 1.0 / (double)a.S_glob;
    const int st = a.ent_stride;
    double *ent = raw + rl.ent(), *gpb = raw + rl.gp();

    // (a) per-component sums of the entmc records: csum[j][f]; the mu block of the entropy gradient (:98)
    //     needs nothing else and is finished here.  Threads run over f fastest => coalesced record reads.
    if (ent_mc) {
        const int nf = anyg ? st : 1;
        for (int e = gtid; e < K * nf; e += GT) {
            const int j = e / nf, f = e - j * nf;
            double v = slab_sum(a, j, f);
            // f = 0: complete sum_i log q(x_i) of component j up to the shared sum_d ln lambda_d term (phase 2)
            if (f == 0) v += a.draws_local * (-0.5 * D * kLog2Pi - D * log(prm[a.lay.sigma() + j]));
            a.csum[j * st + f] = v;
            if (f >= 1 && f < 1 + D) ent[rl.o_mu() + j * D + (f - 1)] = w[j] * v * inv_ns / lambd[f - 1];
        }
    }
    if (!ent_lb && !(ent_mc && anyg)) {  // no entropy gradient from this evaluation: keep the block defined
        for (int e = gtid; e < rl.block(); e += GT) ent[e] = 0.0;
        if (!ent_mc && gtid == 0) raw[0] = 0.0;
    }
    if (gtid == 0) raw[2] = raw[3] = 0.0;

    // (b) log joint, element-wise entries [gp mu (K*D) | gp sigma (K) | gp w (K)]: average over this
    //     rank's hyper-samples (:1578-1596)
    const int n_cs = (a.f.have_ent && a.f.use_ent_mc) ? K * (anyg ? st : 1) : 0;
    for (int e = (gtid + GT - n_cs % GT) % GT; e < K * D + 2 * K; e += GT) {
        const int off = e < K * D ? rl.o_mu() + e : (e < K * D + K ? rl.o_sig() + (e - K * D) : rl.o_w() + (e - K * D - K));
        double v = 0.0;
        if (a.f.have_gp && anyg) {
#pragma unroll 8
            for (int s = a.s_begin; s < a.S; s += a.s_step) v += a.gps[(size_t)s * a.gps_stride + 1 + off];
        }
        gpb[off] = v * inv_S;
    }

    // (c) log joint, one WARP per summed entry: per-sample G_s and lambda block, S_local x (D + 1) sums over the
    //     components (consumers: phase 2 for G and the lambda gradient; variance path, avg_flag = 0, I_sk)
    const int n_we = a.f.have_gp ? a.S_local * (D + 1) : 0;
    for (int idx = GW - 1 - gw; idx < n_we; idx += GW) {
        {
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
    }
    (void)DP;
}

// ---- phase 2: entropy entries that are sums over components / dimensions of csum (one warp each) ---------
__device__ __forceinline__ void raw_phase2(const double *__restrict__ prm, const RawArgs &a, int gtid, int GT) {
    const bool anyg = a.f.grad[0] || a.f.grad[1] || a.f.grad[2] || a.f.grad[3];
    const int D = a.lay.D, DP = a.lay.DP, K = a.lay.K;
    const int lane = gtid & 31, gw = gtid >> 5, GW = GT >> 5;
    {  // G = mean_s G_s (:1425, :1581) and gp d/dlambda_d = mean_s sum_k lamc (:1452-1462, :1596): last threads
        const int e = GT - 1 - gtid;
        if (e < 1 + D) {
            double v = 0.0;
            if (a.f.have_gp && (e == 0 || anyg)) {
                const int off = e == 0 ? 0 : 1 + a.rl.o_lam() + e - 1;
#pragma unroll 8
                for (int s = a.s_begin; s < a.S; s += a.s_step) v += a.gps[(size_t)s * a.gps_stride + off];
            }
            v *= 1.0 / (double)a.S_glob;
            if (e == 0)
                a.raw[1] = v;
            else
                a.raw[a.rl.gp() + a.rl.o_lam() + e - 1] = v;
        }
    }
    if (!(a.f.have_ent && a.f.use_ent_mc)) return;
    const double *sigma = prm + a.lay.sigma(), *lambd = prm + a.lay.lambd(), *w = prm + a.lay.w();
    const RawLayout rl = a.rl;
    const double inv_ns = 1.0 / a.Ns_glob;
    const int st = a.ent_stride;
    const double *cs = a.csum;
    double *ent = a.raw + rl.ent();
    const int n_we = anyg ? 1 + K + D + K : 1;
    for (int idx0 = gw; idx0 < n_we; idx0 += GW) {
        int idx = idx0;
        double sl = 0.0;  // sum_d ln lambda_d (normalisation of the own-component density)
        if (idx == 0 || idx >= 1 + K + D) {
            for (int d = lane; d < D; d += 32) sl += log(lambd[d]);
            sl = warp_sum(sl);
        }
        if (idx == 0) {  // H = -sum_j w_j mean log q  (entmc_vbmc.py:80)
            double h = 0.0;
            for (int j = lane; j < K; j += 32) {
                h -= w[j] * (cs[j * st] - a.draws_local * sl);
            }
            h = warp_sum(h);
            if (lane == 0) a.raw[0] = h * inv_ns;
            continue;
        }
        idx -= 1;
        if (idx < K) {  // d/dsigma_j (:102-103): sum over dimensions of Be
            double v = 0.0;
            for (int d = lane; d < D; d += 32) v += cs[idx * st + 1 + DP + d];
            v = warp_sum(v);
            if (lane == 0) ent[rl.o_sig() + idx] = w[idx] * v * inv_ns / sigma[idx];
            continue;
        }
        idx -= K;
        if (idx < D) {  // d/dlambda_d (:106-108): sum over components
            double v = 0.0;
            for (int j = lane; j < K; j += 32) v += w[j] * cs[j * st + 1 + DP + idx];
            v = warp_sum(v);
            if (lane == 0) ent[rl.o_lam() + idx] = v * inv_ns / lambd[idx];
            continue;
        }
        idx -= D;
        {  // d/dw_k (:111-112): -E_k[log q] - sum_j w_j E_j[N_k / q]
            double v = 0.0;
            if (a.f.grad[3])
                for (int j = lane; j < K; j += 32) v += w[j] * cs[j * st + 1 + 2 * DP + idx];
            v = warp_sum(v);
            if (lane == 0) {
                const double hs = cs[idx * st] - a.draws_local * sl;
                ent[rl.o_w() + idx] = -(hs + v) * inv_ns;
            }
        }
    }
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
    double *lpart;  // [cluster size] per-CTA shares of the bound loss (scratch behind csum)
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

// ---- phase 0 (needs only the parameter block, so it runs BEFORE the first cluster barrier and hides behind the
// record loads of phase 1): bound-loss derivative of the ln-scale entries (tmp, every CTA that finishes entries),
// exp(eta) (sek), and this CTA's share of the bound loss itself -> a.lpart[cluster rank].
__device__ __forceinline__ void final_prep(const double *__restrict__ prm, const FinalArgs &a, bool has_entries,
                                           int crank, int gtid, int GT, double *scratch, double *sek, double *tmp) {
    const int D = a.lay.D, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    const bool bounds = a.f.use_bounds && a.n_bnd > 0;
    const int n_mu = a.f.optimize[0] ? K * D : 0, n_sc = K * D, n_eta = a.f.optimize[3] ? K : 0;
    if (has_entries)
        for (int k = tid; k < K; k += nt) sek[k] = exp(prm[a.lay.eta() + k]);
    double Lb = 0.0;
    if (bounds) {
        if (has_entries) {
#pragma unroll 2
            for (int i = tid; i < n_sc; i += nt) tmp[i] = bound_dy(a, prm, n_mu + i, n_mu, n_sc, nullptr);
        }
#pragma unroll 2
        for (int e = gtid; e < n_mu + n_sc + n_eta; e += GT) bound_dy(a, prm, e, n_mu, n_sc, &Lb);
    }
    Lb = block_sum(Lb, scratch);  // (also the barrier that publishes tmp / sek inside the CTA)
    if (tid == 0) a.lpart[crank] = Lb;
}

// ---- phase 3: raw -> out.  `vb` is the (virtual) block index: block vb finishes entries [vb*nt, (vb+1)*nt) ----
__device__ __forceinline__ void final_phase(const double *__restrict__ prm, const FinalArgs &a, int vb, int nblk,
                                            double *ssm /*[5]: es, dote, dotg, dotp, Lp*/, const double *sek,
                                            const double *tmp /*[K*D] d(bound loss)/d(ln-scale entry), per CTA*/) {
    const int D = a.lay.D, K = a.lay.K, tid = threadIdx.x, nt = blockDim.x;
    const RawLayout rl = a.rl;
    const double *raw = a.raw;
    const double *w = prm + a.lay.w();
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
            const double ek = sek[k];
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
    // ---- one output entry per thread: Jacobians, bound-loss and penalty gradients, dF ---------------
    // (loads first, the softmax sums are only needed by the weight entries after the barrier)
    const int n0 = a.f.grad[0] ? K * D : 0, n1 = a.f.grad[1] ? K : 0, n2 = a.f.grad[2] ? D : 0, n3 = a.f.grad[3] ? K : 0;
    const int e = vb * nt + tid;
    const bool live = anyg && e < n0 + n1 + n2 + n3;
    double gh = 0.0, gg = 0.0, add = 0.0;
    int kw = -1;
    if (live) {
        const double *be = raw + rl.ent(), *bg = raw + rl.gp();
        const double *sigma = prm + a.lay.sigma(), *lambd = prm + a.lay.lambd();
        const int jac = a.f.jacobian;
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
            kw = e - n0 - n1 - n2;
            gh = be[rl.o_w() + kw], gg = bg[rl.o_w() + kw];
            if (bounds && n_eta) add = bound_dy(a, prm, n_mu + n_sc + kw, n_mu, n_sc, nullptr);
        }
    }
    __syncthreads();
    if (live) {
        if (kw >= 0) {
            const double es = ssm[0], dote = ssm[1], dotg = ssm[2], dotp = ssm[3];
            const double ek = sek[kw];
            if (a.f.jacobian) {  // row k of J_w @ g
                gh = ek / es * gh - ek / (es * es) * dote;
                gg = ek / es * gg - ek / (es * es) * dotg;
            }
            if (pen) {  // weight penalty through the softmax Jacobian (:1221-1229)
                const double g = w[kw] < a.w_thr ? a.w_pen : 0.0;
                add += ek / es * g - ek / (es * es) * dotp;
            }
        }
        if (a.f.parts) dH[e] = gh, dG[e] = gg;
        dF[e] = -gg - gh + add;  // :1171-1173, :1200, :1227-1229
    }

    // ---- block 0: the scalars ----------------------------------------------------------------------------
    if (vb == 0 && tid == 0) {
        double Lb = 0.0;
        if (bounds)
            for (int r = 0; r < nblk; ++r) Lb += a.lpart[r];  // fixed order
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

constexpr int kTailThreads = 1024, kTailCluster = 8, kTailGridCtas = 8;
enum { PH_RAW = 1, PH_FINAL = 2, PH_XCHG = 4 };

// All-reduce of the raw vector over NVLink peer memory, inside the tail kernel (no NCCL launch, no second tail
// launch).  Every rank owns an exchange buffer (cudaMalloc + CUDA IPC, mapped by all peers):
//     slot(par, src)  [2][W][stride] doubles : rank `src`'s raw vector of the step with parity `par`
//     flag(par, src)  [2][W] uint64          : epoch of the step whose slot(par, src) is complete
// Step `epoch`: a rank stores its raw vector into slot(epoch & 1, rank) of EVERY rank (its own included), fences at
// system scope, publishes flag = epoch on every rank, waits until all W flags of its own buffer show `epoch`, and
// sums the W slots in rank order -- the same data in the same order on every rank, so the results stay bitwise
// replicated.  Two parities suffice: a rank can run at most one step ahead of the slowest one (it needs that
// rank's flag to finish a step), and a slot is rewritten two steps later.
struct XchgArgs {
    int world, rank, n, stride;  // n = raw length, stride = slot pitch (doubles)
    unsigned long long epoch;
    double *peer[VBMC_P2P_MAX_WORLD];  // exchange buffers of all ranks (peer[rank] = the local one)
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// returns false (in every thread of the CTA that hosts a timed-out waiter ... the caller poisons the result) if a
// peer's flag did not arrive within ~2 s
// Barrier over all CTAs of the tail kernel.  cluster mode: hardware cluster barrier (the kernel is launched as ONE
// thread-block cluster).  grid mode: sense-reversing counter in global memory (the <= 32 CTAs are always co-resident:
// one per SM on an otherwise idle machine); the counter returns to zero and the sense flips on every use, so the two
// words survive from launch to launch and CUDA-graph replays need no reset.
struct TailBarrier {
    unsigned *count;
    volatile unsigned *sense;
    unsigned nblk;
    int use_cluster;
    __device__ __forceinline__ void sync() const {
        if (use_cluster) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
            return;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned s = *sense;
            __threadfence();
            if (atomicAdd(count, 1u) == nblk - 1u) {
                *count = 0u;
                __threadfence();
                *sense = s ^ 1u;
            } else {
                while (*sense == s) {
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
};

template <class Cluster>
__device__ __forceinline__ void raw_exchange(const XchgArgs &x, double *__restrict__ raw, int gtid, int GT, Cluster &cluster,
                                             int *s_bad) {
    const int W = x.world, par = (int)(x.epoch & 1ull);
    const size_t flags_at = (size_t)2 * W * x.stride;
    // 1. publish the local raw vector in every rank's slot(par, rank)
    for (int q = 0; q < W; ++q) {
        double *dst = x.peer[q] + ((size_t)par * W + x.rank) * x.stride;
        for (int i = gtid; i < x.n; i += GT) dst[i] = raw[i];
    }
    __threadfence_system();
    cluster.sync();
    if (gtid < W) {
        unsigned long long *f = reinterpret_cast<unsigned long long *>(x.peer[gtid] + flags_at) + (size_t)par * W + x.rank;
        st_release_sys_u64(f, x.epoch);
    }
    // 2. wait for the W flags of the local buffer
    if (gtid < W) {
        const unsigned long long *f =
            reinterpret_cast<const unsigned long long *>(x.peer[x.rank] + flags_at) + (size_t)par * W + gtid;
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(f) != x.epoch) {
            if (clock64() - t0 > 4000000000LL) {  // ~2 s: a lost peer must not hang the GPU
                *s_bad = 1;
                break;
            }
        }
    }
    cluster.sync();
    // 3. fixed-order sum of the W slots -> raw (identical on every rank)
    const double *mine = x.peer[x.rank] + (size_t)par * W * x.stride;
    for (int i = gtid; i < x.n; i += GT) {
        double v = 0.0;
        for (int src = 0; src < W; ++src) v += __ldcg(mine + (size_t)src * x.stride + i);  // (L2: peers wrote it)
        raw[i] = v;
    }
}

// One launch for everything behind the producers: a cluster of 8 CTAs; the phases are separated by cluster
// barriers (release/acquire at cluster scope, so plain global stores of one phase are visible to the next).
__global__ void __launch_bounds__(kTailThreads) tail_kernel(const double *__restrict__ prm, RawArgs ra, FinalArgs fa,
                                                           int phases, XchgArgs xa, TailBarrier cluster) {
    __shared__ double scratch[40];
    __shared__ double ssm[5];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    extern __shared__ double tmp[];  // [K*D] bound-loss derivatives of the ln-scale entries | [K] exp(eta)
    const int nblk = (int)gridDim.x, gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x), GT = nblk * (int)blockDim.x;
    int nvb = 0;
    double *sek = nullptr;
#ifdef VBMC_TAIL_DEBUG
    unsigned long long ts[7];
#define TS(i) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[i])::"memory")
#else
#define TS(i)
#endif
    TS(0);
    if (phases & PH_FINAL) {
        const bool anyg = fa.f.grad[0] || fa.f.grad[1] || fa.f.grad[2] || fa.f.grad[3];
        const int P = anyg ? (fa.f.grad[0] ? fa.lay.K * fa.lay.D : 0) + (fa.f.grad[1] ? fa.lay.K : 0) +
                                 (fa.f.grad[2] ? fa.lay.D : 0) + (fa.f.grad[3] ? fa.lay.K : 0)
                           : 0;
        nvb = P > 0 ? (P + (int)blockDim.x - 1) / (int)blockDim.x : 1;
        sek = tmp + fa.lay.K * fa.lay.D;
        final_prep(prm, fa, (int)blockIdx.x < nvb, (int)blockIdx.x, gtid, GT, scratch, sek, tmp);
    }
    TS(1);
    if (phases & PH_RAW) {
        raw_phase1(prm, ra, gtid, GT);
        TS(2);
        cluster.sync();
        TS(3);
        raw_phase2(prm, ra, gtid, GT);
        if (phases & PH_XCHG) {
            cluster.sync();  // the local raw vector is complete
            raw_exchange(xa, ra.raw, gtid, GT, cluster, &s_bad);
        }
    }
    TS(4);
    if (phases & PH_FINAL) {
        cluster.sync();  // raw vector (fused launch) and the bound-loss shares are complete
        TS(5);
        for (int vb = blockIdx.x; vb < nvb; vb += nblk) {
            final_phase(prm, fa, vb, nblk, ssm, sek, tmp);
            __syncthreads();
        }
    }
    if ((phases & PH_XCHG) && (phases & PH_FINAL)) {
        // a peer that never answered: make the failure visible instead of returning a partial sum
        __syncthreads();
        if (s_bad && threadIdx.x == 0) fa.out[0] = __longlong_as_double(0x7ff8000000000000LL), fa.out[7] = 2.0;
    }
    TS(6);
#ifdef VBMC_TAIL_DEBUG
    if (threadIdx.x == 0 && phases == (PH_RAW | PH_FINAL))
        printf("tail cta %d: prep %llu p1 %llu sync %llu p2 %llu sync %llu final %llu ns\n", (int)blockIdx.x, ts[1] - ts[0],
               ts[2] - ts[1], ts[3] - ts[2], ts[4] - ts[3], ts[5] - ts[4], ts[6] - ts[5]);
#endif
#undef TS
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

int fill_raw_args(Ctx *c, int D, int K, const EvalFlags &f, const EntmcPlan *plan, int64_t Ns_glob, int s_begin,
                  int s_step, int S_glob, double *d_raw, RawArgs *out) {
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
        a.cm = plan->variant == ENTMC_TC ? ChunkMap{(long long)plan->chunk, (long long)plan->chunk_small, plan->n_big}
                                         : ChunkMap{(long long)plan->chunk, (long long)plan->chunk, plan->grid};
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
    VBMC_TRY(ensure(&c->d_csum, &c->csum_cap, (size_t)K * a.ent_stride + 64));
    a.csum = c->d_csum + 64;
    *out = a;
    return VBMC_OK;
}

int fill_final_args(Ctx *c, int D, int K, const EvalFlags &f, const double *d_raw, double *d_out, FinalArgs *out) {
    FinalArgs a{};
    VBMC_TRY(ensure(&c->d_csum, &c->csum_cap, (size_t)K * entpart_stride(pad_dim(D), K) + 64));
    a.lpart = c->d_csum;
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
    *out = a;
    return VBMC_OK;
}

// raw phase deferred by reduce_launch(defer = true), consumed by the next finalize_launch of the same context
static_assert(sizeof(RawArgs) <= sizeof(Ctx::raw_pending_blob), "Ctx::raw_pending_blob is too small");

int tail_launch(Ctx *c, const double *d_params, const RawArgs &ra, const FinalArgs &fa, int phases, int D, int K,
                const XchgArgs &xa = XchgArgs{}) {
    const size_t smem = (phases & PH_FINAL) ? (size_t)(K * D + K) * sizeof(double) : 0;
    VBMC_REQUIRE(smem <= 200 * 1024, VBMC_ERR_UNSUPPORTED, "finalize: D*K too large");
    if (smem > c->finalize_smem_set && smem > 48 * 1024) {
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->finalize_smem_set = smem;
    }
    static int tail_ctas = 0, tail_threads = 0, tail_grid_mode = -1;
    if (tail_ctas == 0) {
        const char *e1 = getenv("VBMC_TAIL_CLUSTER"), *e2 = getenv("VBMC_TAIL_THREADS"), *e3 = getenv("VBMC_TAIL_MODE");
        // measured (C3, 1 GPU): hardware cluster barrier 75.4 us per step, global-memory barrier 78.2-78.7 us (8-32 CTAs)
        tail_grid_mode = (e3 && !strcmp(e3, "grid")) ? 1 : 0;
        tail_ctas = e1 ? atoi(e1) : (tail_grid_mode ? kTailGridCtas : kTailCluster);
        tail_threads = e2 ? atoi(e2) : 0;  // 0: by problem size, below
        if (tail_ctas > 32) tail_ctas = 32;
        if (!tail_grid_mode && tail_ctas > 8) tail_ctas = 8;
    }
    if (!c->d_tailsync) {  // barrier words of the grid mode: zero once, self-resetting afterwards
        VBMC_CUDA_CHECK(cudaMalloc((void **)&c->d_tailsync, 64));
        VBMC_CUDA_CHECK(cudaMemsetAsync(c->d_tailsync, 0, 64, c->stream));
    }
    TailBarrier tb{c->d_tailsync, c->d_tailsync + 8, (unsigned)tail_ctas, tail_grid_mode ? 0 : 1};
    cudaLaunchConfig_t cfg{};
    // 8 x 1024 threads at C3 size (K D = 1000; fewer or smaller CTAs are all slower there); small problems (K D <= 512:
    // C1, C2, C4) finish 2 us sooner with 512-thread CTAs, at every draw count (measured: 22.5 -> 20.5 us per device-resident
    // evaluation of C2 / C4 at the reference's default draw counts, 31.1 -> 28.7 / 40.1 -> 38.0 us at 100k / 200k draws)
    const int threads = tail_threads > 0 ? tail_threads : (K * D <= 512 ? 512 : kTailThreads);
    cfg.gridDim = dim3(tail_ctas), cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = tail_ctas, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = tail_grid_mode ? 0 : 1;
    VBMC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tail_kernel, d_params, ra, fa, phases, xa, tb));
    c->launches++;
    return VBMC_OK;
}

}  // namespace

namespace {
__global__ void __launch_bounds__(1024) stage_copy_kernel(double *__restrict__ dst, const double *__restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
}  // namespace

// pinned host block (device-addressable under the same pointer) -> device buffer, as a kernel on stream `st`
int stage_copy_launch(Ctx *c, double *d_dst, const double *h_pinned_src, int n, cudaStream_t st) {
    stage_copy_kernel<<<1, n >= 1024 ? 1024 : 32, 0, st>>>(d_dst, h_pinned_src, n);
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

// records / per-sample terms -> raw vector (device pointer d_raw).  defer = true (single GPU): nothing is
// launched; the raw phases run fused with the next finalize_launch on the same raw vector (one launch).
int reduce_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const EntmcPlan *plan,
                  int64_t Ns_glob, int s_begin, int s_step, int S_glob, double *d_raw, bool defer) {
    RawArgs ra;
    VBMC_TRY(fill_raw_args(c, D, K, f, plan, Ns_glob, s_begin, s_step, S_glob, d_raw, &ra));
    if (defer) {
        memcpy(c->raw_pending_blob, &ra, sizeof(RawArgs));
        c->raw_pending = true;
        c->raw_pending_p2p = s_step > 1;  // deferred with W > 1 ranks (s_step = W): the peer-memory all-reduce path
        return VBMC_OK;
    }
    c->raw_pending = false;
    return tail_launch(c, d_params, ra, FinalArgs{}, PH_RAW, D, K);
}

int finalize_launch(Ctx *c, const double *d_params, int D, int K, const EvalFlags &f, const double *d_raw,
                    double *d_out) {
    FinalArgs fa;
    VBMC_TRY(fill_final_args(c, D, K, f, d_raw, d_out, &fa));
    if (c->raw_pending) {
        c->raw_pending = false;
        RawArgs ra;
        memcpy(&ra, c->raw_pending_blob, sizeof(RawArgs));
        VBMC_REQUIRE(ra.raw == d_raw, VBMC_ERR_STATE, "finalize: raw vector differs from the one given to partials");
        if (c->raw_pending_p2p) {  // W ranks: raw phases -> peer-memory all-reduce -> final phase, ONE launch
            c->raw_pending_p2p = false;
            XchgArgs xa{};
            xa.world = c->p2p_world, xa.rank = c->p2p_rank, xa.n = fa.rl.total(), xa.stride = c->p2p_stride;
            xa.epoch = ++c->p2p_epoch;
            for (int q = 0; q < c->p2p_world; ++q) xa.peer[q] = c->p2p_peer[q];
            return tail_launch(c, d_params, ra, fa, PH_RAW | PH_XCHG | PH_FINAL, D, K, xa);
        }
        return tail_launch(c, d_params, ra, fa, PH_RAW | PH_FINAL, D, K);
    }
    return tail_launch(c, d_params, RawArgs{}, fa, PH_FINAL, D, K);
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

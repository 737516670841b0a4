// entmc.cu -- Monte-Carlo mixture entropy and its reparameterisation gradient.
//
// Replaces the NumPy loop of pyvbmc/entropy/entmc_vbmc.py:64-112 (reference).
//
// Work decomposition
//   grid = (slabs, K); CTA (slab, j) owns a contiguous range of ANTITHETIC PAIRS of
//   component j.  One thread evaluates one pair (+eps, -eps) at a time: the pair shares the
//   noise registers e_d = sigma_j*eps_d and every shared-memory load of the component table.
//
// Arithmetic (scaled coordinates y = x / lambda, so that lambda drops out of the distances)
//   Delta_kd = (mu_dj - mu_dk) / lambda_d                      (fp64 -> T, exact for k == j)
//   t(+-)_kd = Delta_kd +- e_d                                  = (x_d - mu_dk) / lambda_d
//   s_k      = D log2(sigma_j/sigma_k) - h_k |t_k|^2 + h_j |e|^2,   h_k = log2(e) / (2 sigma_k^2)
//            = log2( N_k(x) / N_j(x) )      -> the own component is the log-sum-exp reference
//              point: u_j = 1, every other u_k = 2^s_k is a density RATIO under x ~ N_j, whose
//              mean is 1 (overflow needs a 2^127-sigma event; checked downstream via isfinite).
//   q_rel    = sum_k w_k u_k ,  log q(x) = ln N_j(x) + ln q_rel
//   l_d      = sum_k (w_k u_k / sigma_k^2) t_kd      ( = lambda_d * lsum_d / N_j(x),  :93-95 )
//   per pair:  hacc += log q(x+) + log q(x-)
//              A_d  += l+_d/q+ + l-_d/q-                  (-> d/dmu_j,            :98)
//              Be_d += e_d (l+_d/q+ - l-_d/q-)            (-> d/dsigma_j, d/dlambda :102-108)
//              racc_k += u+_k/q+ + u-_k/q-                (-> d/dw_k,             :112)
//   Per-thread running sums live in shared memory columns (no bank conflicts, no atomics);
//   each CTA writes ONE record [hacc | A | Be | racc] of fp64 partial sums, reduced in fixed
//   order by reduce_kernel (finalize.cu) => bitwise run-to-run determinism.
//
// T = float : fp32 compute, fp64 accumulation across threads (default, ~6D+10 FP32-pipe
//             instructions per (pair, k)).   T = double: everything fp64 (validation mode).
#include <algorithm>

#include "common.cuh"
#include "philox.cuh"

namespace vbmc {

namespace {

template <typename T>
struct M;
template <>
struct M<float> {
    static __device__ __forceinline__ float ex2(float x) {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ double lg2(float x) { return (double)log2f(x); }
};
template <>
struct M<double> {
    static __device__ __forceinline__ double ex2(double x) { return exp2(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double lg2(double x) { return log2(x); }
};

template <typename T>
struct alignas(16) KConst {
    T ck, h, w, wis2;
};

template <typename T, int DP, bool WGRAD, bool ANYGRAD, bool PHILOX>
__global__ void __launch_bounds__(128)
entmc_kernel(const double *__restrict__ prm, ParamLayout lay, int64_t half, int64_t pair0, int64_t half_glob,
             int R, const double *__restrict__ eps, uint64_t seed, uint64_t offset, double *__restrict__ part,
             int part_stride) {
    constexpr bool kKeepT = sizeof(T) == 4;  // fp32: keep t(+-) in registers; fp64: recompute
    const int D = lay.D, K = lay.K;
    {  // the Philox key lives behind the parameter block so that a captured CUDA graph stays valid
        const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
        seed = rngp[0], offset = rngp[1];
    }
    const int j = blockIdx.y, slab = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sDl = reinterpret_cast<T *>(smem_raw);                        // [K][DP]
    KConst<T> *sKc = reinterpret_cast<KConst<T> *>(sDl + K * DP);    // [K]
    T *sAcc = reinterpret_cast<T *>(sKc + K);                        // accA [DP][nt], accB [DP][nt]
    T *sU = sAcc + (ANYGRAD ? 2 * DP * nt : 0);                      // Up [K][nt], Um [K][nt], racc [K][nt]
    double *scratch = reinterpret_cast<double *>(sU + (WGRAD ? 3 * K * nt : 0));  // [32] (16B aligned by sizes)

    const double *mu = prm + lay.mu();
    const double *sigma = prm + lay.sigma();
    const double *lambd = prm + lay.lambd();
    const double *w = prm + lay.w();
    const double sig_j = sigma[j];
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2

    // ---- component tables ---------------------------------------------------------------
    for (int i = tid; i < K * DP; i += nt) {
        const int k = i / DP, d = i - k * DP;
        sDl[i] = (d < D) ? (T)((mu[j * D + d] - mu[k * D + d]) / lambd[d]) : (T)0;
    }
    for (int k = tid; k < K; k += nt) {
        const double sk = sigma[k];
        KConst<T> c;
        c.ck = (T)(D * (log2(sig_j) - log2(sk)));
        c.h = (T)(kHalfLog2e / (sk * sk));
        c.w = (T)w[k];
        c.wis2 = (T)(w[k] / (sk * sk));
        sKc[k] = c;
    }
    if (ANYGRAD)
        for (int i = tid; i < 2 * DP * nt; i += nt) sAcc[i] = (T)0;
    if (WGRAD)
        for (int i = tid; i < K * nt; i += nt) sU[2 * K * nt + i] = (T)0;
    __syncthreads();

    T *accA = sAcc + tid, *accB = sAcc + DP * nt + tid;
    T *Up = sU + tid, *Um = sU + K * nt + tid, *racc = sU + 2 * K * nt + tid;
    const T hj = (T)(kHalfLog2e / (sig_j * sig_j));
    const double is2j = 1.0 / (sig_j * sig_j);
    const T sj = (T)sig_j;
    double hacc = 0.0;

    const int64_t slab_base = (int64_t)slab * nt * R;
    for (int r = 0; r < R; ++r) {
        const int64_t p = slab_base + (int64_t)r * nt + tid;  // local pair index
        if (p >= half) break;                                  // no barriers inside the loop
        const int64_t gpair = pair0 + p;

        T e[DP];
        if (PHILOX) {
            float z[DP];
            philox_normals<DP>(seed, offset, (uint32_t)j, (uint64_t)gpair, D, z);
#pragma unroll
            for (int d = 0; d < DP; ++d) e[d] = sj * (T)z[d];
        } else {
            const double *ep = eps + ((size_t)j * (size_t)half_glob + (size_t)gpair) * (size_t)D;
#pragma unroll
            for (int d = 0; d < DP; ++d) e[d] = (d < D) ? sj * (T)__ldg(ep + d) : (T)0;
        }
        T e2a = 0, e2b = 0;
#pragma unroll
        for (int d = 0; d < DP; d += 2) {
            e2a = fma(e[d], e[d], e2a);
            e2b = fma(e[d + 1], e[d + 1], e2b);
        }
        const T e2 = e2a + e2b;
        const T base = hj * e2;

        T lp[ANYGRAD ? DP : 1], lm[ANYGRAD ? DP : 1];
        if (ANYGRAD) {
#pragma unroll
            for (int d = 0; d < DP; ++d) lp[d] = lm[d] = (T)0;
        }
        T qp = 0, qm = 0;

#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const KConst<T> c = sKc[k];
            const T *dl = sDl + k * DP;
            T tp[kKeepT ? DP : 1], tm[kKeepT ? DP : 1];
            T a0 = 0, a1 = 0, b0 = 0, b1 = 0;
#pragma unroll
            for (int d = 0; d < DP; d += 2) {
                const T x0 = dl[d] + e[d], y0 = dl[d] - e[d];
                const T x1 = dl[d + 1] + e[d + 1], y1 = dl[d + 1] - e[d + 1];
                if (kKeepT) {
                    tp[d] = x0, tm[d] = y0, tp[d + 1] = x1, tm[d + 1] = y1;
                }
                a0 = fma(x0, x0, a0), b0 = fma(y0, y0, b0);
                a1 = fma(x1, x1, a1), b1 = fma(y1, y1, b1);
            }
            const T cb = c.ck + base;
            const T up = M<T>::ex2(fma(-c.h, a0 + a1, cb));
            const T um = M<T>::ex2(fma(-c.h, b0 + b1, cb));
            if (WGRAD) {
                Up[k * nt] = up;
                Um[k * nt] = um;
            }
            qp = fma(c.w, up, qp);
            qm = fma(c.w, um, qm);
            if (ANYGRAD) {
                const T gp = c.wis2 * up, gm = c.wis2 * um;
#pragma unroll
                for (int d = 0; d < DP; ++d) {
                    const T x = kKeepT ? tp[d] : dl[d] + e[d];
                    const T y = kKeepT ? tm[d] : dl[d] - e[d];
                    lp[d] = fma(gp, x, lp[d]);
                    lm[d] = fma(gm, y, lm[d]);
                }
            }
        }

        // log q(x+) + log q(x-)  without the per-component constant (added by reduce_kernel)
        hacc += 0.69314718055994530942 * (M<T>::lg2(qp) + M<T>::lg2(qm)) - (double)e2 * is2j;
        if (ANYGRAD) {
            const T iqp = M<T>::rcp(qp), iqm = M<T>::rcp(qm);
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                const T a = lp[d] * iqp, b = lm[d] * iqm;
                accA[d * nt] += a + b;
                accB[d * nt] += e[d] * (a - b);
            }
            if (WGRAD) {
                for (int k = 0; k < K; ++k) racc[k * nt] += fma(Up[k * nt], iqp, Um[k * nt] * iqm);
            }
        }
    }

    // ---- CTA record: fixed-order fp64 reduction over the thread columns ------------------
    double *rec = part + ((size_t)j * gridDim.x + slab) * (size_t)part_stride;
    const double hs = block_sum(hacc, scratch);
    if (tid == 0) rec[0] = hs;
    if (ANYGRAD) {
        __syncthreads();
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        const int rows = 2 * DP + (WGRAD ? K : 0);
        for (int row = wid; row < rows; row += nw) {
            const T *src = (row < 2 * DP) ? (sAcc + row * nt) : (sU + 2 * K * nt + (row - 2 * DP) * nt);
            double v = 0.0;
            for (int c = lane; c < nt; c += 32) v += (double)src[c];
            v = warp_sum(v);
            if (lane == 0) rec[1 + row] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 production kernel: same mathematics as entmc_kernel<float,...>, but
//   * all per-dimension arithmetic is issued as packed f32x2 instructions (sm_100 FADD2 / FFMA2,
//     two adjacent dimensions per instruction): 6*DP/2 packed ops per (pair, component) instead of
//     6*DP scalar ones -- the kernel is bound by instruction issue / fixed-latency stalls at
//     2 CTAs per SM, so halving the instruction count is what moves it towards the FMA-pipe roof;
//   * the per-thread gradient sums A_d, Be_d stay in registers across the thread's pairs and are
//     spilled to shared-memory columns once, right before the CTA record is reduced.
template <int DP, bool WGRAD, bool ANYGRAD, bool PHILOX>
__global__ void __launch_bounds__(128, 2)
entmc_kernel_f32x2(const double *__restrict__ prm, ParamLayout lay, int64_t half, int64_t pair0, int64_t half_glob,
                   int R, const double *__restrict__ eps, uint64_t seed, uint64_t offset,
                   double *__restrict__ part, int part_stride) {
    constexpr int H = DP / 2;  // packed pairs of dimensions
    const int D = lay.D, K = lay.K;
    {  // the Philox key lives behind the parameter block so that a captured CUDA graph stays valid
        const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
        seed = rngp[0], offset = rngp[1];
    }
    const int j = blockIdx.y, slab = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sDl = reinterpret_cast<float *>(smem_raw);                        // [K][DP]
    KConst<float> *sKc = reinterpret_cast<KConst<float> *>(sDl + K * DP);    // [K]
    float *sU = reinterpret_cast<float *>(sKc + K);  // Up [K][nt], Um [K][nt], racc [K][nt]  (WGRAD)
    // without WGRAD the region only serves as the [2*DP][nt] spill area of the record reduction
    const int u_floats = WGRAD ? 3 * K * nt : (ANYGRAD ? 2 * DP * nt : 0);
    double *scratch = reinterpret_cast<double *>(sU + ((u_floats + 3) & ~3));

    const double *mu = prm + lay.mu();
    const double *sigma = prm + lay.sigma();
    const double *lambd = prm + lay.lambd();
    const double *w = prm + lay.w();
    const double sig_j = sigma[j];
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2

    for (int i = tid; i < K * DP; i += nt) {
        const int k = i / DP, d = i - k * DP;
        sDl[i] = (d < D) ? (float)((mu[j * D + d] - mu[k * D + d]) / lambd[d]) : 0.0f;
    }
    for (int k = tid; k < K; k += nt) {
        const double sk = sigma[k];
        KConst<float> c;
        c.ck = (float)(D * (log2(sig_j) - log2(sk)));
        c.h = (float)(kHalfLog2e / (sk * sk));
        c.w = (float)w[k];
        c.wis2 = (float)(w[k] / (sk * sk));
        sKc[k] = c;
    }
    if (WGRAD)
        for (int i = tid; i < K * nt; i += nt) sU[2 * K * nt + i] = 0.0f;
    __syncthreads();

    float *Up = sU + tid, *Um = sU + K * nt + tid, *racc = sU + 2 * K * nt + tid;
    const float hj = (float)(kHalfLog2e / (sig_j * sig_j));
    const double is2j = 1.0 / (sig_j * sig_j);
    const float sj = (float)sig_j;
    double hacc = 0.0;
    float2 accA[ANYGRAD ? H : 1], accB[ANYGRAD ? H : 1];
    if constexpr (ANYGRAD) {
#pragma unroll
        for (int i = 0; i < H; ++i) accA[i] = accB[i] = make_float2(0.f, 0.f);
    }

    const int64_t slab_base = (int64_t)slab * nt * R;
    for (int r = 0; r < R; ++r) {
        const int64_t p = slab_base + (int64_t)r * nt + tid;  // local pair index
        if (p >= half) break;                                  // no barriers inside the loop
        const int64_t gpair = pair0 + p;

        float2 e2[H], ne2[H];
        {
            float z[DP];
            if (PHILOX) {
                philox_normals<DP>(seed, offset, (uint32_t)j, (uint64_t)gpair, D, z);
            } else {
                const double *ep = eps + ((size_t)j * (size_t)half_glob + (size_t)gpair) * (size_t)D;
#pragma unroll
                for (int d = 0; d < DP; ++d) z[d] = (d < D) ? (float)__ldg(ep + d) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < H; ++i) {
                e2[i] = make_float2(sj * z[2 * i], sj * z[2 * i + 1]);
                ne2[i] = make_float2(-e2[i].x, -e2[i].y);
            }
        }
        float2 ee = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < H; ++i) ee = __ffma2_rn(e2[i], e2[i], ee);
        const float e2sum = ee.x + ee.y;
        const float base = hj * e2sum;

        float2 lp[ANYGRAD ? H : 1], lm[ANYGRAD ? H : 1];
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int i = 0; i < H; ++i) lp[i] = lm[i] = make_float2(0.f, 0.f);
        }
        float qp = 0.f, qm = 0.f;

#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const KConst<float> c = sKc[k];
            const float2 *dl2 = reinterpret_cast<const float2 *>(sDl + k * DP);
            float2 tp[H], tm[H];
            float2 ap0 = make_float2(0.f, 0.f), ap1 = ap0, am0 = ap0, am1 = ap0;
#pragma unroll
            for (int i = 0; i < H; i += 2) {
                const float2 d0 = dl2[i], d1 = dl2[i + 1];  // one LDS.128
                tp[i] = __fadd2_rn(d0, e2[i]);
                tm[i] = __fadd2_rn(d0, ne2[i]);
                tp[i + 1] = __fadd2_rn(d1, e2[i + 1]);
                tm[i + 1] = __fadd2_rn(d1, ne2[i + 1]);
                ap0 = __ffma2_rn(tp[i], tp[i], ap0);
                am0 = __ffma2_rn(tm[i], tm[i], am0);
                ap1 = __ffma2_rn(tp[i + 1], tp[i + 1], ap1);
                am1 = __ffma2_rn(tm[i + 1], tm[i + 1], am1);
            }
            const float2 ap = __fadd2_rn(ap0, ap1), am = __fadd2_rn(am0, am1);
            const float cb = c.ck + base;
            const float up = M<float>::ex2(fmaf(-c.h, ap.x + ap.y, cb));
            const float um = M<float>::ex2(fmaf(-c.h, am.x + am.y, cb));
            if (WGRAD) {
                Up[k * nt] = up;
                Um[k * nt] = um;
            }
            qp = fmaf(c.w, up, qp);
            qm = fmaf(c.w, um, qm);
            if constexpr (ANYGRAD) {
                const float gpv = c.wis2 * up, gmv = c.wis2 * um;
                const float2 gp2 = make_float2(gpv, gpv), gm2 = make_float2(gmv, gmv);
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    lp[i] = __ffma2_rn(gp2, tp[i], lp[i]);
                    lm[i] = __ffma2_rn(gm2, tm[i], lm[i]);
                }
            }
        }

        hacc += 0.69314718055994530942 * ((double)log2f(qp) + (double)log2f(qm)) - (double)e2sum * is2j;
        if constexpr (ANYGRAD) {
            const float iqp = __frcp_rn(qp), iqm = __frcp_rn(qm);
            const float2 ip2 = make_float2(iqp, iqp), im2 = make_float2(iqm, iqm), nim2 = make_float2(-iqm, -iqm);
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const float2 a = __fmul2_rn(lp[i], ip2);
                accA[i] = __fadd2_rn(accA[i], __ffma2_rn(lm[i], im2, a));      // l+/q+ + l-/q-
                accB[i] = __ffma2_rn(e2[i], __ffma2_rn(lm[i], nim2, a), accB[i]);  // e (l+/q+ - l-/q-)
            }
            if (WGRAD) {
                for (int k = 0; k < K; ++k) racc[k * nt] += fmaf(Up[k * nt], iqp, Um[k * nt] * iqm);
            }
        }
    }

    // ---- CTA record: fixed-order fp64 reduction over the thread columns ------------------
    double *rec = part + ((size_t)j * gridDim.x + slab) * (size_t)part_stride;
    const double hs = block_sum(hacc, scratch);
    if (tid == 0) rec[0] = hs;
    if constexpr (ANYGRAD) {
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        __syncthreads();
        if (WGRAD) {
            for (int row = wid; row < K; row += nw) {
                const float *src = sU + 2 * K * nt + row * nt;
                double v = 0.0;
                for (int c = lane; c < nt; c += 32) v += (double)src[c];
                v = warp_sum(v);
                if (lane == 0) rec[1 + 2 * DP + row] = v;
            }
            __syncthreads();
        }
        // spill the register sums into columns [2*DP][nt] (the Up/Um scratch is free now)
#pragma unroll
        for (int i = 0; i < H; ++i) {
            sU[(2 * i) * nt + tid] = accA[i].x;
            sU[(2 * i + 1) * nt + tid] = accA[i].y;
            sU[(DP + 2 * i) * nt + tid] = accB[i].x;
            sU[(DP + 2 * i + 1) * nt + tid] = accB[i].y;
        }
        __syncthreads();
        for (int row = wid; row < 2 * DP; row += nw) {
            const float *src = sU + row * nt;
            double v = 0.0;
            for (int c = lane; c < nt; c += 32) v += (double)src[c];
            v = warp_sum(v);
            if (lane == 0) rec[1 + row] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 production kernel, "dimension-split": TWO ADJACENT LANES evaluate one antithetic pair, lane
// h in {0,1} holding dims [h*DP/2, (h+1)*DP/2).  Same instruction count per pair as the one-thread
// version (the per-dimension work splits evenly; only the scalar tail is duplicated), but half the
// registers (~110) and half the per-thread shared-memory columns, so 4 CTAs x 128 threads fit per
// SM instead of 2: the kernel is bound by fixed-latency stalls of the dependent FADD2->FFMA2 chains,
// and twice the resident warps is what hides them (ncu: issue-active 43 % -> see profiles/).
//   per (pair, k) and lane:  DP/4 FADD2 x2 (t+, t-), DP/4 FFMA2 x2 (|t|^2), 2 SHFL (combine the two
//   half-sums), 2 MUFU.EX2, DP/4 FFMA2 x2 (l+, l-).
template <int DP, bool WGRAD, bool ANYGRAD, bool PHILOX>
__global__ void __launch_bounds__(128, 4)
entmc_kernel_f32x2_ds(const double *__restrict__ prm, ParamLayout lay, int64_t half, int64_t pair0,
                      int64_t half_glob, int R, const double *__restrict__ eps, uint64_t seed, uint64_t offset,
                      double *__restrict__ part, int part_stride) {
    constexpr int DH = DP / 2;            // dims per lane
    constexpr int H2 = DH / 2;            // packed float2 per lane
    constexpr int DHP = (DH + 3) & ~3;    // table row length per half (float4 addressable)
    constexpr int NQ = DHP / 4;
    const int D = lay.D, K = lay.K;
    {  // the Philox key lives behind the parameter block so that a captured CUDA graph stays valid
        const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
        seed = rngp[0], offset = rngp[1];
    }
    const int j = blockIdx.y, slab = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int h = tid & 1, pt = tid >> 1, npt = nt >> 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sDl = reinterpret_cast<float *>(smem_raw);                             // [K][2][DHP]
    KConst<float> *sKc = reinterpret_cast<KConst<float> *>(sDl + K * 2 * DHP);    // [K]
    float *sU = reinterpret_cast<float *>(sKc + K);  // U [K][nt], racc [K][nt] (WGRAD); later the [2*DP][npt] spill
    const int u_floats = WGRAD ? max(2 * K * nt, DP * nt) : (ANYGRAD ? DP * nt : 0);
    double *scratch = reinterpret_cast<double *>(sU + ((u_floats + 3) & ~3));

    const double *mu = prm + lay.mu();
    const double *sigma = prm + lay.sigma();
    const double *lambd = prm + lay.lambd();
    const double *w = prm + lay.w();
    const double sig_j = sigma[j];
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2

    for (int i = tid; i < K * 2 * DHP; i += nt) {
        const int k = i / (2 * DHP), r = i - k * 2 * DHP, hh = r / DHP, c = r - hh * DHP;
        const int d = hh * DH + c;
        sDl[i] = (c < DH && d < D) ? (float)((mu[j * D + d] - mu[k * D + d]) / lambd[d]) : 0.0f;
    }
    for (int k = tid; k < K; k += nt) {
        const double sk = sigma[k];
        KConst<float> c;
        c.ck = (float)(D * (log2(sig_j) - log2(sk)));
        c.h = (float)(kHalfLog2e / (sk * sk));
        c.w = (float)w[k];
        c.wis2 = (float)(w[k] / (sk * sk));
        sKc[k] = c;
    }
    if (WGRAD)
        for (int i = tid; i < K * nt; i += nt) sU[K * nt + i] = 0.0f;
    __syncthreads();

    float *Ucol = sU + tid, *racc = sU + K * nt + tid;
    const float hj = (float)(kHalfLog2e / (sig_j * sig_j));
    const double is2j = 1.0 / (sig_j * sig_j);
    const float sj = (float)sig_j;
    double hacc = 0.0;
    float2 accA[ANYGRAD ? H2 : 1], accB[ANYGRAD ? H2 : 1];
    if constexpr (ANYGRAD) {
#pragma unroll
        for (int i = 0; i < H2; ++i) accA[i] = accB[i] = make_float2(0.f, 0.f);
    }

    const int64_t slab_base = (int64_t)slab * npt * R;
    for (int r = 0; r < R; ++r) {
        const int64_t p = slab_base + (int64_t)r * npt + pt;  // local pair index (same for both lanes)
        const bool live = p < half;                            // whole warps keep running: shuffles below
        if (!__any_sync(0xffffffffu, live)) break;
        const int64_t gpair = pair0 + (live ? p : 0);

        float2 e2[H2], ne2[H2];
        {
            float z[DH];
            if (PHILOX) {
                philox_normals_half<DH>(seed, offset, (uint32_t)j, (uint64_t)gpair, h, D, z);
            } else {
                const double *ep = eps + ((size_t)j * (size_t)half_glob + (size_t)gpair) * (size_t)D + h * DH;
#pragma unroll
                for (int i = 0; i < DH; ++i) z[i] = (h * DH + i < D) ? (float)__ldg(ep + i) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < H2; ++i) {
                e2[i] = make_float2(sj * z[2 * i], sj * z[2 * i + 1]);
                ne2[i] = make_float2(-e2[i].x, -e2[i].y);
            }
        }
        float2 ee = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < H2; ++i) ee = __ffma2_rn(e2[i], e2[i], ee);
        float e2sum = ee.x + ee.y;
        e2sum += __shfl_xor_sync(0xffffffffu, e2sum, 1);
        const float base = hj * e2sum;

        float2 lp[ANYGRAD ? H2 : 1], lm[ANYGRAD ? H2 : 1];
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int i = 0; i < H2; ++i) lp[i] = lm[i] = make_float2(0.f, 0.f);
        }
        float qp = 0.f, qm = 0.f;

#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const KConst<float> c = sKc[k];
            const float4 *row = reinterpret_cast<const float4 *>(sDl + (k * 2 + h) * DHP);
            float2 tp[H2], tm[H2];
            float2 ap0 = make_float2(0.f, 0.f), ap1 = ap0, am0 = ap0, am1 = ap0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 v = row[q];
                if (2 * q < H2) {
                    const float2 d0 = make_float2(v.x, v.y);
                    tp[2 * q] = __fadd2_rn(d0, e2[2 * q]);
                    tm[2 * q] = __fadd2_rn(d0, ne2[2 * q]);
                    ap0 = __ffma2_rn(tp[2 * q], tp[2 * q], ap0);
                    am0 = __ffma2_rn(tm[2 * q], tm[2 * q], am0);
                }
                if (2 * q + 1 < H2) {
                    const float2 d1 = make_float2(v.z, v.w);
                    tp[2 * q + 1] = __fadd2_rn(d1, e2[2 * q + 1]);
                    tm[2 * q + 1] = __fadd2_rn(d1, ne2[2 * q + 1]);
                    ap1 = __ffma2_rn(tp[2 * q + 1], tp[2 * q + 1], ap1);
                    am1 = __ffma2_rn(tm[2 * q + 1], tm[2 * q + 1], am1);
                }
            }
            const float2 ap = __fadd2_rn(ap0, ap1), am = __fadd2_rn(am0, am1);
            float pa = ap.x + ap.y, pm = am.x + am.y;
            pa += __shfl_xor_sync(0xffffffffu, pa, 1);  // |t+|^2 over all dims (commutative: both lanes agree bitwise)
            pm += __shfl_xor_sync(0xffffffffu, pm, 1);
            const float cb = c.ck + base;
            const float up = M<float>::ex2(fmaf(-c.h, pa, cb));
            const float um = M<float>::ex2(fmaf(-c.h, pm, cb));
            if (WGRAD) Ucol[k * nt] = h ? um : up;  // lane 0 keeps u+, lane 1 keeps u-
            qp = fmaf(c.w, up, qp);
            qm = fmaf(c.w, um, qm);
            if constexpr (ANYGRAD) {
                const float gpv = c.wis2 * up, gmv = c.wis2 * um;
                const float2 gp2 = make_float2(gpv, gpv), gm2 = make_float2(gmv, gmv);
#pragma unroll
                for (int i = 0; i < H2; ++i) {
                    lp[i] = __ffma2_rn(gp2, tp[i], lp[i]);
                    lm[i] = __ffma2_rn(gm2, tm[i], lm[i]);
                }
            }
        }

        if (live && h == 0)
            hacc += 0.69314718055994530942 * ((double)log2f(qp) + (double)log2f(qm)) - (double)e2sum * is2j;
        if constexpr (ANYGRAD) {
            const float iqp = live ? __frcp_rn(qp) : 0.f, iqm = live ? __frcp_rn(qm) : 0.f;
            const float2 ip2 = make_float2(iqp, iqp), im2 = make_float2(iqm, iqm), nim2 = make_float2(-iqm, -iqm);
#pragma unroll
            for (int i = 0; i < H2; ++i) {
                const float2 a = __fmul2_rn(lp[i], ip2);
                accA[i] = __fadd2_rn(accA[i], __ffma2_rn(lm[i], im2, a));          // l+/q+ + l-/q-
                accB[i] = __ffma2_rn(e2[i], __ffma2_rn(lm[i], nim2, a), accB[i]);  // e (l+/q+ - l-/q-)
            }
            if (WGRAD) {
                const float iq = h ? iqm : iqp;
                for (int k = 0; k < K; ++k) racc[k * nt] = fmaf(Ucol[k * nt], iq, racc[k * nt]);
            }
        }
    }

    // ---- CTA record: fixed-order fp64 reduction over the thread columns ------------------
    double *rec = part + ((size_t)j * gridDim.x + slab) * (size_t)part_stride;
    const double hs = block_sum(hacc, scratch);
    if (tid == 0) rec[0] = hs;
    if constexpr (ANYGRAD) {
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        __syncthreads();
        if (WGRAD) {
            for (int row = wid; row < K; row += nw) {
                const float *src = sU + K * nt + row * nt;
                double v = 0.0;
                for (int c = lane; c < nt; c += 32) v += (double)src[c];
                v = warp_sum(v);
                if (lane == 0) rec[1 + 2 * DP + row] = v;
            }
            __syncthreads();
        }
        // spill the register sums into [2*DP][npt] columns (the U scratch is free now)
#pragma unroll
        for (int i = 0; i < H2; ++i) {
            const int d = h * DH + 2 * i;
            sU[(d)*npt + pt] = accA[i].x;
            sU[(d + 1) * npt + pt] = accA[i].y;
            sU[(DP + d) * npt + pt] = accB[i].x;
            sU[(DP + d + 1) * npt + pt] = accB[i].y;
        }
        __syncthreads();
        for (int row = wid; row < 2 * DP; row += nw) {
            const float *src = sU + row * npt;
            double v = 0.0;
            for (int c = lane; c < npt; c += 32) v += (double)src[c];
            v = warp_sum(v);
            if (lane == 0) rec[1 + row] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 production kernel, "expanded" form.  One thread per antithetic pair.  Per (pair, component):
//   B      = sum_d Delta_kd e_d                                  (DP/2 FFMA2)
//   |t+-|^2 = A_k + E +- 2B,   A_k = |Delta_k|^2 (table), E = |e|^2 (per pair)
//   s+-    = (ck - h_k A_k) + (h_j - h_k) E -+ 2 h_k B           (3 FFMA, constants folded per k)
//   l+-    = sum_k g+-_k Delta_k  +- e sum_k g+-_k                (DP/2 FFMA2 each; the e-term once per pair)
// i.e. 1.5 DP packed FMAs per (pair, k) instead of the 3 DP of the direct form, and no t+- registers
// live across the exponential.  The expansion cancels when a draw lands close to ANOTHER component's
// centre relative to |Delta|: the absolute error of s is ~ h_k (A_k + E) 2^-24.  The set-up therefore
// flags every k with h_k (A_k + E_max) > kGuard and those components take the DIRECT path
// (t = Delta +- e, |t|^2 summed term by term, l += g t) inside the same loop; the flag is uniform over
// the CTA (one mixture component j per CTA), so the branch never diverges.  k == j has Delta = 0, A = 0:
// its distance is E exactly in both forms.
struct alignas(8) KFast {
    float ck2, h2, hd, w, wis2, flag;  // ck - h A | 2 h | h_j - h | w_k | w_k / sigma_k^2 | 1 => direct path
    float ck, h;                       // direct-path constants
};

template <int DP, bool WGRAD, bool ANYGRAD, bool PHILOX>
__global__ void __launch_bounds__(128, 2)
entmc_kernel_fast(const double *__restrict__ prm, ParamLayout lay, int64_t half, int64_t pair0, int64_t half_glob,
                  int R, const double *__restrict__ eps, uint64_t seed, uint64_t offset, double *__restrict__ part,
                  int part_stride, float guard) {
    constexpr int H = DP / 2;  // packed pairs of dimensions
    const int D = lay.D, K = lay.K;
    {  // the Philox key lives behind the parameter block so that a captured CUDA graph stays valid
        const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
        seed = rngp[0], offset = rngp[1];
    }
    const int j = blockIdx.y, slab = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sDl = reinterpret_cast<float *>(smem_raw);              // [K][DP]
    KFast *sKc = reinterpret_cast<KFast *>(sDl + K * DP);          // [K]
    float *sU = reinterpret_cast<float *>(sKc + K);  // Up [K][nt], Um [K][nt], racc [K][nt] (WGRAD); spill [2 DP][nt]
    const int u_floats = WGRAD ? max(3 * K * nt, 2 * DP * nt) : (ANYGRAD ? 2 * DP * nt : 0);
    double *scratch = reinterpret_cast<double *>(sU + ((u_floats + 3) & ~3));

    const double *mu = prm + lay.mu();
    const double *sigma = prm + lay.sigma();
    const double *lambd = prm + lay.lambd();
    const double *w = prm + lay.w();
    const double sig_j = sigma[j];
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2

    for (int i = tid; i < K * DP; i += nt) {
        const int k = i / DP, d = i - k * DP;
        sDl[i] = (d < D) ? (float)((mu[j * D + d] - mu[k * D + d]) / lambd[d]) : 0.0f;
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        const double sk = sigma[k];
        const double hk = kHalfLog2e / (sk * sk), hjd = kHalfLog2e / (sig_j * sig_j);
        const double ck = D * (log2(sig_j) - log2(sk));
        double A = 0.0;  // |Delta_k|^2 of the ROUNDED table entries (what the FFMA2s will see)
        for (int d = 0; d < D; ++d) A += (double)sDl[k * DP + d] * (double)sDl[k * DP + d];
        // bound on E = sigma_j^2 |eps|^2 : |eps|^2 <= D + 8 sqrt(2 D) + 32 except with probability < 1e-12
        const double Emax = sig_j * sig_j * (D + 8.0 * sqrt(2.0 * D) + 32.0);
        KFast c;
        c.ck2 = (float)(ck - hk * A);
        c.h2 = (float)(2.0 * hk);
        c.hd = (float)(hjd - hk);
        c.w = (float)w[k];
        c.wis2 = (float)(w[k] / (sk * sk));
        c.flag = (hk * (A + Emax) > (double)guard && k != j) ? 1.0f : 0.0f;
        c.ck = (float)ck;
        c.h = (float)hk;
        sKc[k] = c;
    }
    if (WGRAD)
        for (int i = tid; i < K * nt; i += nt) sU[2 * K * nt + i] = 0.0f;
    __syncthreads();

    float *Up = sU + tid, *Um = sU + K * nt + tid, *racc = sU + 2 * K * nt + tid;
    const float hj = (float)(kHalfLog2e / (sig_j * sig_j));
    const double is2j = 1.0 / (sig_j * sig_j);
    const float sj = (float)sig_j;
    double hacc = 0.0;
    float2 accA[ANYGRAD ? H : 1], accB[ANYGRAD ? H : 1];
    if constexpr (ANYGRAD) {
#pragma unroll
        for (int i = 0; i < H; ++i) accA[i] = accB[i] = make_float2(0.f, 0.f);
    }

    const int64_t slab_base = (int64_t)slab * nt * R;
    for (int r = 0; r < R; ++r) {
        const int64_t p = slab_base + (int64_t)r * nt + tid;  // local pair index
        if (p >= half) break;                                  // no barriers / shuffles inside the loop
        const int64_t gpair = pair0 + p;

        float2 e2[H];
        {
            float z[DP];
            if (PHILOX) {
                philox_normals<DP>(seed, offset, (uint32_t)j, (uint64_t)gpair, D, z);
            } else {
                const double *ep = eps + ((size_t)j * (size_t)half_glob + (size_t)gpair) * (size_t)D;
#pragma unroll
                for (int d = 0; d < DP; ++d) z[d] = (d < D) ? (float)__ldg(ep + d) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < H; ++i) e2[i] = make_float2(sj * z[2 * i], sj * z[2 * i + 1]);
        }
        float2 ee = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < H; ++i) ee = __ffma2_rn(e2[i], e2[i], ee);
        const float E = ee.x + ee.y;
        const float base = hj * E;

        float2 lp[ANYGRAD ? H : 1], lm[ANYGRAD ? H : 1];
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int i = 0; i < H; ++i) lp[i] = lm[i] = make_float2(0.f, 0.f);
        }
        float qp = 0.f, qm = 0.f, Gp = 0.f, Gm = 0.f;

#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const KFast c = sKc[k];
            const float2 *dl2 = reinterpret_cast<const float2 *>(sDl + k * DP);
            float2 dl[H];
#pragma unroll
            for (int i = 0; i < H; ++i) dl[i] = dl2[i];
            float up, um;
            if (c.flag == 0.0f) {
                float2 b0 = make_float2(0.f, 0.f), b1 = b0;
#pragma unroll
                for (int i = 0; i + 1 < H; i += 2) {
                    b0 = __ffma2_rn(dl[i], e2[i], b0);
                    b1 = __ffma2_rn(dl[i + 1], e2[i + 1], b1);
                }
                if (H & 1) b0 = __ffma2_rn(dl[H - 1], e2[H - 1], b0);
                const float2 bb = __fadd2_rn(b0, b1);
                const float B = bb.x + bb.y;
                const float s0 = fmaf(c.hd, E, c.ck2);
                up = M<float>::ex2(fmaf(-c.h2, B, s0));
                um = M<float>::ex2(fmaf(c.h2, B, s0));
                qp = fmaf(c.w, up, qp);
                qm = fmaf(c.w, um, qm);
                if constexpr (ANYGRAD) {
                    const float gpv = c.wis2 * up, gmv = c.wis2 * um;
                    Gp += gpv;
                    Gm += gmv;
                    const float2 gp2 = make_float2(gpv, gpv), gm2 = make_float2(gmv, gmv);
#pragma unroll
                    for (int i = 0; i < H; ++i) {
                        lp[i] = __ffma2_rn(gp2, dl[i], lp[i]);
                        lm[i] = __ffma2_rn(gm2, dl[i], lm[i]);
                    }
                }
            } else {
                // direct path: differences first, squared term by term (no cancellation)
                float2 a0 = make_float2(0.f, 0.f), a1 = a0;
                float2 tp[H], tm[H];
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    tp[i] = __fadd2_rn(dl[i], e2[i]);
                    tm[i] = __fadd2_rn(dl[i], make_float2(-e2[i].x, -e2[i].y));
                    a0 = __ffma2_rn(tp[i], tp[i], a0);
                    a1 = __ffma2_rn(tm[i], tm[i], a1);
                }
                const float cb = c.ck + base;
                up = M<float>::ex2(fmaf(-c.h, a0.x + a0.y, cb));
                um = M<float>::ex2(fmaf(-c.h, a1.x + a1.y, cb));
                qp = fmaf(c.w, up, qp);
                qm = fmaf(c.w, um, qm);
                if constexpr (ANYGRAD) {
                    const float gpv = c.wis2 * up, gmv = c.wis2 * um;
                    const float2 gp2 = make_float2(gpv, gpv), gm2 = make_float2(gmv, gmv);
#pragma unroll
                    for (int i = 0; i < H; ++i) {
                        lp[i] = __ffma2_rn(gp2, tp[i], lp[i]);
                        lm[i] = __ffma2_rn(gm2, tm[i], lm[i]);
                    }
                }
            }
            if (WGRAD) {
                Up[k * nt] = up;
                Um[k * nt] = um;
            }
        }

        hacc += 0.69314718055994530942 * ((double)log2f(qp) + (double)log2f(qm)) - (double)E * is2j;
        if constexpr (ANYGRAD) {
            const float iqp = __frcp_rn(qp), iqm = __frcp_rn(qm);
            const float2 ip2 = make_float2(iqp, iqp), im2 = make_float2(iqm, iqm);
            const float2 Gp2 = make_float2(Gp, Gp), nGm2 = make_float2(-Gm, -Gm);
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const float2 a = __fmul2_rn(__ffma2_rn(e2[i], Gp2, lp[i]), ip2);   // l+ / q+
                const float2 b = __fmul2_rn(__ffma2_rn(e2[i], nGm2, lm[i]), im2);  // l- / q-
                accA[i] = __fadd2_rn(accA[i], __fadd2_rn(a, b));
                accB[i] = __ffma2_rn(e2[i], __fadd2_rn(a, make_float2(-b.x, -b.y)), accB[i]);
            }
            if (WGRAD) {
                for (int k = 0; k < K; ++k) racc[k * nt] += fmaf(Up[k * nt], iqp, Um[k * nt] * iqm);
            }
        }
    }

    // ---- CTA record: fixed-order fp64 reduction over the thread columns ------------------
    double *rec = part + ((size_t)j * gridDim.x + slab) * (size_t)part_stride;
    const double hs = block_sum(hacc, scratch);
    if (tid == 0) rec[0] = hs;
    if constexpr (ANYGRAD) {
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        __syncthreads();
        if (WGRAD) {
            for (int row = wid; row < K; row += nw) {
                const float *src = sU + 2 * K * nt + row * nt;
                double v = 0.0;
                for (int c = lane; c < nt; c += 32) v += (double)src[c];
                v = warp_sum(v);
                if (lane == 0) rec[1 + 2 * DP + row] = v;
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < H; ++i) {
            sU[(2 * i) * nt + tid] = accA[i].x;
            sU[(2 * i + 1) * nt + tid] = accA[i].y;
            sU[(DP + 2 * i) * nt + tid] = accB[i].x;
            sU[(DP + 2 * i + 1) * nt + tid] = accB[i].y;
        }
        __syncthreads();
        for (int row = wid; row < 2 * DP; row += nw) {
            const float *src = sU + row * nt;
            double v = 0.0;
            for (int c = lane; c < nt; c += 32) v += (double)src[c];
            v = warp_sum(v);
            if (lane == 0) rec[1 + row] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 kernel for SMALL draw counts (the reference's defaults: ~28 draws per component).  Same arithmetic,
// tables, guard and record layout as entmc_kernel_fast; what changes is the mapping.  With 14 antithetic pairs per
// component, one thread per pair leaves 14 lanes of one warp walking serially over all K components (22.6 us at C3,
// 9.6 % issue-active, every other warp parked at the final reduction: profiles/r4d).  Here EIGHT lanes share a pair and
// split the components (k = lane, lane + 8, ...): the component loop is 8x shorter and every warp of the CTA works.
// Only q+ and q- (the mixture densities) are needed in full by every lane -- three butterfly shuffles each; everything
// else a pair contributes (the l+- / G+- partial sums) enters the results LINEARLY once 1/q+- is known, so each lane
// folds its own partial sums into its private accumulators and the CTA-wide column reduction at the end adds them up.
template <int DP, bool WGRAD, bool ANYGRAD, bool PHILOX>
__global__ void __launch_bounds__(128, 2)
entmc_kernel_small(const double *__restrict__ prm, ParamLayout lay, int64_t half, int64_t pair0, int64_t half_glob,
                   int R, const double *__restrict__ eps, uint64_t seed, uint64_t offset, double *__restrict__ part,
                   int part_stride, float guard) {
    constexpr int H = DP / 2;  // packed pairs of dimensions
    constexpr int KL = 8;      // lanes per antithetic pair
    const int D = lay.D, K = lay.K;
    {  // the Philox key lives behind the parameter block so that a captured CUDA graph stays valid
        const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
        seed = rngp[0], offset = rngp[1];
    }
    const int j = blockIdx.y, slab = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int G = nt / KL, g = tid / KL, l = tid % KL;  // pair slots per sweep; this thread's slot and component lane

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sDl = reinterpret_cast<float *>(smem_raw);      // [K][DP]
    KFast *sKc = reinterpret_cast<KFast *>(sDl + K * DP);  // [K]
    float *sU = reinterpret_cast<float *>(sKc + K);        // Up [K][G], Um [K][G], racc [K][G] (WGRAD); spill [2 DP][nt]
    const int u_floats = WGRAD ? max(3 * K * G, 2 * DP * nt) : (ANYGRAD ? 2 * DP * nt : 0);
    double *scratch = reinterpret_cast<double *>(sU + ((u_floats + 3) & ~3));

    const double *mu = prm + lay.mu();
    const double *sigma = prm + lay.sigma();
    const double *lambd = prm + lay.lambd();
    const double *w = prm + lay.w();
    const double sig_j = sigma[j];
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2

    // The set-up is a large share of this kernel (one wave of K CTAs, a few microseconds in all): one reciprocal per
    // dimension instead of a division per table entry, the loads of four entries in flight at a time, one logarithm per
    // component (of the ratio), four partial sums for |Delta_k|^2.
    double *sInvL = scratch + 40;  // [DP] 1 / lambda_d, then mu_j / lambda_d
    if (tid < DP) sInvL[tid] = tid < D ? 1.0 / lambd[tid] : 0.0;
    __syncthreads();
#pragma unroll 4
    for (int i = tid; i < K * DP; i += nt) {
        const int k = i / DP, d = i - k * DP;
        sDl[i] = (d < D) ? (float)((mu[j * D + d] - mu[k * D + d]) * sInvL[d]) : 0.0f;
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        const double sk = sigma[k];
        const double hk = kHalfLog2e / (sk * sk), hjd = kHalfLog2e / (sig_j * sig_j);
        const double ck = D * log2(sig_j / sk);
        double A0 = 0.0, A1 = 0.0, A2 = 0.0, A3 = 0.0;  // |Delta_k|^2 of the ROUNDED table entries (what the FFMA2s will see)
        const float4 *row = reinterpret_cast<const float4 *>(sDl + k * DP);  // (DP % 4 == 0, padded entries are zero)
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
            const float4 v = row[q];
            A0 += (double)v.x * (double)v.x, A1 += (double)v.y * (double)v.y;
            A2 += (double)v.z * (double)v.z, A3 += (double)v.w * (double)v.w;
        }
        const double A = (A0 + A1) + (A2 + A3);
        const double Emax = sig_j * sig_j * (D + 8.0 * sqrt(2.0 * D) + 32.0);  // (see entmc_kernel_fast)
        KFast c;
        c.ck2 = (float)(ck - hk * A);
        c.h2 = (float)(2.0 * hk);
        c.hd = (float)(hjd - hk);
        c.w = (float)w[k];
        c.wis2 = (float)(w[k] / (sk * sk));
        c.flag = (hk * (A + Emax) > (double)guard && k != j) ? 1.0f : 0.0f;
        c.ck = (float)ck;
        c.h = (float)hk;
        sKc[k] = c;
    }
    if (WGRAD)
        for (int i = tid; i < K * G; i += nt) sU[2 * K * G + i] = 0.0f;
    __syncthreads();

    float *Up = sU + g, *Um = sU + K * G + g, *racc = sU + 2 * K * G + g;  // element k at [k * G]
    const float hj = (float)(kHalfLog2e / (sig_j * sig_j));
    const double is2j = 1.0 / (sig_j * sig_j);
    const float sj = (float)sig_j;
    double hacc = 0.0;
    float2 accA[ANYGRAD ? H : 1], accB[ANYGRAD ? H : 1];
    if constexpr (ANYGRAD) {
#pragma unroll
        for (int i = 0; i < H; ++i) accA[i] = accB[i] = make_float2(0.f, 0.f);
    }

    const int64_t slab_base = (int64_t)slab * G * R;
    for (int r = 0; r < R; ++r) {  // (uniform trip count: the shuffles below need every lane of the warp)
        const int64_t p = slab_base + (int64_t)r * G + g;  // local pair index of this thread's slot
        const bool live = p < half;
        const int64_t gpair = pair0 + (live ? p : 0);

        float2 e2[H];
        {
            float z[DP];
            if (PHILOX) {
                philox_normals<DP>(seed, offset, (uint32_t)j, (uint64_t)gpair, D, z);  // (the 8 lanes of a slot draw the same numbers)
            } else {
                const double *ep = eps + ((size_t)j * (size_t)half_glob + (size_t)gpair) * (size_t)D;
#pragma unroll
                for (int d = 0; d < DP; ++d) z[d] = (d < D) ? (float)__ldg(ep + d) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < H; ++i) e2[i] = make_float2(sj * z[2 * i], sj * z[2 * i + 1]);
        }
        float2 ee = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < H; ++i) ee = __ffma2_rn(e2[i], e2[i], ee);
        const float E = ee.x + ee.y;
        const float base = hj * E;

        // partial sums over this lane's components
        float2 lp[ANYGRAD ? H : 1], lm[ANYGRAD ? H : 1];
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int i = 0; i < H; ++i) lp[i] = lm[i] = make_float2(0.f, 0.f);
        }
        float qp = 0.f, qm = 0.f, Gp = 0.f, Gm = 0.f;

        if (live) {
            for (int k = l; k < K; k += KL) {
                const KFast c = sKc[k];
                const float2 *dl2 = reinterpret_cast<const float2 *>(sDl + k * DP);
                float2 dl[H];
#pragma unroll
                for (int i = 0; i < H; ++i) dl[i] = dl2[i];
                float up, um;
                if (c.flag == 0.0f) {
                    float2 b0 = make_float2(0.f, 0.f), b1 = b0;
#pragma unroll
                    for (int i = 0; i + 1 < H; i += 2) {
                        b0 = __ffma2_rn(dl[i], e2[i], b0);
                        b1 = __ffma2_rn(dl[i + 1], e2[i + 1], b1);
                    }
                    if (H & 1) b0 = __ffma2_rn(dl[H - 1], e2[H - 1], b0);
                    const float2 bb = __fadd2_rn(b0, b1);
                    const float B = bb.x + bb.y;
                    const float s0 = fmaf(c.hd, E, c.ck2);
                    up = M<float>::ex2(fmaf(-c.h2, B, s0));
                    um = M<float>::ex2(fmaf(c.h2, B, s0));
                    qp = fmaf(c.w, up, qp);
                    qm = fmaf(c.w, um, qm);
                    if constexpr (ANYGRAD) {
                        const float gpv = c.wis2 * up, gmv = c.wis2 * um;
                        Gp += gpv;
                        Gm += gmv;
                        const float2 gp2 = make_float2(gpv, gpv), gm2 = make_float2(gmv, gmv);
#pragma unroll
                        for (int i = 0; i < H; ++i) {
                            lp[i] = __ffma2_rn(gp2, dl[i], lp[i]);
                            lm[i] = __ffma2_rn(gm2, dl[i], lm[i]);
                        }
                    }
                } else {
                    // direct path: differences first, squared term by term (no cancellation)
                    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
                    float2 tp[H], tm[H];
#pragma unroll
                    for (int i = 0; i < H; ++i) {
                        tp[i] = __fadd2_rn(dl[i], e2[i]);
                        tm[i] = __fadd2_rn(dl[i], make_float2(-e2[i].x, -e2[i].y));
                        a0 = __ffma2_rn(tp[i], tp[i], a0);
                        a1 = __ffma2_rn(tm[i], tm[i], a1);
                    }
                    const float cb = c.ck + base;
                    up = M<float>::ex2(fmaf(-c.h, a0.x + a0.y, cb));
                    um = M<float>::ex2(fmaf(-c.h, a1.x + a1.y, cb));
                    qp = fmaf(c.w, up, qp);
                    qm = fmaf(c.w, um, qm);
                    if constexpr (ANYGRAD) {
                        const float gpv = c.wis2 * up, gmv = c.wis2 * um;
                        const float2 gp2 = make_float2(gpv, gpv), gm2 = make_float2(gmv, gmv);
#pragma unroll
                        for (int i = 0; i < H; ++i) {
                            lp[i] = __ffma2_rn(gp2, tp[i], lp[i]);
                            lm[i] = __ffma2_rn(gm2, tm[i], lm[i]);
                        }
                    }
                }
                if (WGRAD) {
                    Up[k * G] = up;
                    Um[k * G] = um;
                }
            }
        }
        // the mixture densities of the pair: butterfly over the 8 lanes of the slot (fixed order => reproducible)
#pragma unroll
        for (int o = 1; o < KL; o <<= 1) {
            qp += __shfl_xor_sync(0xffffffffu, qp, o);
            qm += __shfl_xor_sync(0xffffffffu, qm, o);
        }
        if (live) {
            if (l == 0) hacc += 0.69314718055994530942 * ((double)log2f(qp) + (double)log2f(qm)) - (double)E * is2j;
            if constexpr (ANYGRAD) {
                const float iqp = __frcp_rn(qp), iqm = __frcp_rn(qm);
                const float2 ip2 = make_float2(iqp, iqp), im2 = make_float2(iqm, iqm);
                const float2 Gp2 = make_float2(Gp, Gp), nGm2 = make_float2(-Gm, -Gm);
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float2 a = __fmul2_rn(__ffma2_rn(e2[i], Gp2, lp[i]), ip2);   // this lane's share of l+ / q+
                    const float2 b = __fmul2_rn(__ffma2_rn(e2[i], nGm2, lm[i]), im2);  // ... of l- / q-
                    accA[i] = __fadd2_rn(accA[i], __fadd2_rn(a, b));
                    accB[i] = __ffma2_rn(e2[i], __fadd2_rn(a, make_float2(-b.x, -b.y)), accB[i]);
                }
                if (WGRAD) {
                    for (int k = l; k < K; k += KL) racc[k * G] += fmaf(Up[k * G], iqp, Um[k * G] * iqm);
                }
            }
        }
    }

    // ---- CTA record: fixed-order fp64 reduction over the thread columns ------------------
    double *rec = part + ((size_t)j * gridDim.x + slab) * (size_t)part_stride;
    const double hs = block_sum(hacc, scratch);
    if (tid == 0) rec[0] = hs;
    if constexpr (ANYGRAD) {
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        __syncthreads();
        if (WGRAD) {
            // G (= 16) columns per row: two rows per warp pass, one per half-warp
            const int hw = lane >> 4, hl = lane & 15;
            for (int r0 = 2 * wid; r0 < K; r0 += 2 * nw) {
                const int row = r0 + hw;
                double v = 0.0;
                if (row < K) {
                    const float *src = sU + 2 * K * G + row * G;
                    for (int c = hl; c < G; c += 16) v += (double)src[c];
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (row < K && hl == 0) rec[1 + 2 * DP + row] = v;
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < H; ++i) {
            sU[(2 * i) * nt + tid] = accA[i].x;
            sU[(2 * i + 1) * nt + tid] = accA[i].y;
            sU[(DP + 2 * i) * nt + tid] = accB[i].x;
            sU[(DP + 2 * i + 1) * nt + tid] = accB[i].y;
        }
        __syncthreads();
        for (int row = wid; row < 2 * DP; row += nw) {
            const float *src = sU + row * nt;
            double v = 0.0;
            for (int c = lane; c < nt; c += 32) v += (double)src[c];
            v = warp_sum(v);
            if (lane == 0) rec[1 + row] = v;
        }
    }
}

static size_t entmc_smem_small(int DP, int K, int nt, bool wgrad, bool anygrad) {
    size_t b = (size_t)K * DP * sizeof(float) + (size_t)K * sizeof(KFast);
    size_t u = wgrad ? (size_t)3 * K * (nt / 8) : 0;
    if (anygrad && u < (size_t)2 * DP * nt) u = (size_t)2 * DP * nt;
    u = (u + 3) & ~(size_t)3;
    b += u * sizeof(float);
    b = (b + 15) & ~(size_t)15;
    return b + (40 + DP) * sizeof(double);
}

static size_t entmc_smem_fast(int DP, int K, int nt, bool wgrad, bool anygrad) {
    size_t b = (size_t)K * DP * sizeof(float) + (size_t)K * sizeof(KFast);
    size_t u = wgrad ? (size_t)3 * K * nt : 0;
    if (anygrad && u < (size_t)2 * DP * nt) u = (size_t)2 * DP * nt;
    u = (u + 3) & ~(size_t)3;
    b += u * sizeof(float);
    b = (b + 15) & ~(size_t)15;
    return b + 40 * sizeof(double);
}

// ---------------------------------------------------------------------------------------------
// fp32 production kernel, "warp-autonomous" (default).  Arithmetic = the expanded form of
// entmc_kernel_fast (same guard, same direct-path fallback); what changes is the work distribution:
//   * the K * half antithetic pairs form ONE index space cut into equal chunks, one per CTA (grid <= 3
//     CTAs per SM: a single, balanced wave); a chunk that straddles components is processed segment by
//     segment, the component tables being rebuilt at each boundary (at most twice for realistic sizes);
//   * inside a segment every WARP is an independent worker on batches of 32 pairs: it owns a 32-column
//     tile of the u(+-) scratch, keeps its d/dw sums racc_k in registers (lane l owns rows l, l+32, ...:
//     after each batch it folds the tile row-wise with conflict-free LDS.128), and reduces its own
//     gradient sums through its tile -- no block-wide barrier inside a segment, no per-thread racc
//     columns (-200 B/thread of shared memory => 3 CTAs/SM instead of 2);
//   * per segment the four warp records are combined through shared memory and ONE fp64 record is
//     written: part[(cta * maxseg + seg)].  raw_kernel recomputes the same chunk arithmetic.
constexpr int kUS = 132;  // row stride (floats) of the u tile: 128 columns + 4 => LDS.128 row reads are conflict-free

template <int DP, bool WGRAD, bool ANYGRAD, bool PHILOX>
__global__ void __launch_bounds__(128, (DP <= 20 ? 3 : 2))
entmc_kernel_w(const double *__restrict__ prm, ParamLayout lay, int64_t half, int64_t pair0, int64_t half_glob,
               int64_t chunk, int maxseg, const double *__restrict__ eps, uint64_t seed, uint64_t offset,
               double *__restrict__ part, int part_stride, float guard) {
    constexpr int H = DP / 2;
    constexpr int NW = 4;
    constexpr int RR = 4;  // rows of the tile per lane: K <= 128 and 2*DP <= 64 rows both fit
    const int D = lay.D, K = lay.K;
    {
        const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
        seed = rngp[0], offset = rngp[1];
    }
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int urows = max(WGRAD ? 2 * K : 0, ANYGRAD ? 2 * DP : 0);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sDl = reinterpret_cast<float *>(smem_raw);             // [K][DP]   (rows in PERMUTED order, see sPerm)
    KFast *sKc = reinterpret_cast<KFast *>(sDl + K * DP);         // [K]       (permuted order)
    float *sU = reinterpret_cast<float *>(sKc + K);               // [urows][kUS]: u+ rows [0,K), u- rows [K,2K)
    float *sIq = sU + urows * kUS;                                // [2][128]
    double *sRec = reinterpret_cast<double *>(sIq + 2 * 128);     // [NW][part_stride]
    double *sInvL = sRec + NW * part_stride;                      // [DP]
    int *sPerm = reinterpret_cast<int *>(sInvL + DP);             // [K] position -> component; [K] = #fast
    float *sFlag = reinterpret_cast<float *>(sPerm + K + 1);      // [K] scratch of the set-up

    const double *mu = prm + lay.mu();
    const double *sigma = prm + lay.sigma();
    const double *lambd = prm + lay.lambd();
    const double *w = prm + lay.w();
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2

    const int64_t T = (int64_t)K * half;
    const int64_t g0 = (int64_t)blockIdx.x * chunk, g1 = min(g0 + chunk, T);
    if (g0 >= T) return;
    if (tid < DP) sInvL[tid] = tid < D ? 1.0 / lambd[tid] : 0.0;

    const int j_first = (int)(g0 / half);
    for (int seg = 0;; ++seg) {
        const int j = j_first + seg;
        const int64_t lo = max(g0, (int64_t)j * half), hi = min(g1, (int64_t)(j + 1) * half);
        if (j >= K || lo >= hi) break;
        const int64_t p_lo = lo - (int64_t)j * half;
        const int n = (int)(hi - lo);

        // ---- component tables ----------------------------------------------------------------------
        // (1) conditioning flag of every component: h_k (A_k + E_max) > guard => direct path (see _fast)
        __syncthreads();
        const double sig_j = sigma[j];
        const double hjd = kHalfLog2e / (sig_j * sig_j);
        const double Emax = sig_j * sig_j * (D + 8.0 * sqrt(2.0 * D) + 32.0);
        for (int k = tid; k < K; k += 128) {
            double A = 0.0;
            for (int d = 0; d < D; ++d) {
                const double t = (mu[j * D + d] - mu[k * D + d]) * sInvL[d];
                A = fma(t, t, A);
            }
            const double hk = kHalfLog2e / (sigma[k] * sigma[k]);
            sFlag[k] = (hk * (A + Emax) > (double)guard && k != j) ? 1.0f : 0.0f;
        }
        __syncthreads();
        // (2) stable partition: un-flagged components first (branch-free inner loop), flagged ones after
        if (tid == 0) {
            int nf = 0;
            for (int k = 0; k < K; ++k)
                if (sFlag[k] == 0.0f) sPerm[nf++] = k;
            sPerm[K] = nf;
            for (int k = 0; k < K; ++k)
                if (sFlag[k] != 0.0f) sPerm[nf++] = k;
        }
        __syncthreads();
        const int nfast = sPerm[K];
        // (3) tables in permuted order
        for (int i = tid; i < K * DP; i += 128) {
            const int pos = i / DP, d = i - pos * DP, k = sPerm[pos];
            sDl[i] = (d < D) ? (float)((mu[j * D + d] - mu[k * D + d]) * sInvL[d]) : 0.0f;
        }
        __syncthreads();
        for (int pos = tid; pos < K; pos += 128) {
            const int k = sPerm[pos];
            const double sk = sigma[k];
            const double hk = kHalfLog2e / (sk * sk);
            const double ck = D * (log2(sig_j) - log2(sk));
            double A = 0.0;  // |Delta_k|^2 of the ROUNDED table entries (what the FFMA2s will see)
            for (int d = 0; d < D; ++d) A += (double)sDl[pos * DP + d] * (double)sDl[pos * DP + d];
            KFast c;
            c.ck2 = (float)(ck - hk * A);
            c.h2 = (float)(2.0 * hk);
            c.hd = (float)(hjd - hk);
            c.w = (float)w[k];
            c.wis2 = (float)(w[k] / (sk * sk));
            c.flag = sFlag[k];
            c.ck = (float)ck;
            c.h = (float)hk;
            sKc[pos] = c;
        }
        __syncthreads();

        const float hj = (float)hjd;
        const double is2j = 1.0 / (sig_j * sig_j);
        const float sj = (float)sig_j;
        double hacc = 0.0;
        double rowacc[RR] = {0.0, 0.0, 0.0, 0.0};  // fp64 sums of the gradient rows lane + 32 r  (A | Be)
        float racc[RR] = {0.f, 0.f, 0.f, 0.f};     // d/dw sums of the (permuted) component rows lane + 32 r
        float *Ucol = sU + tid;                     // column of this thread inside its warp's tile
        const float4 *tile4 = reinterpret_cast<const float4 *>(sU + wid * 32);

        for (int b = wid; b * 32 < n; b += NW) {
            const int off = b * 32 + lane;
            const bool live = off < n;
            const int64_t gpair = pair0 + p_lo + (live ? off : 0);

            float2 e2[H];
            {
                float z[DP];
                if (PHILOX) {
                    philox_normals<DP>(seed, offset, (uint32_t)j, (uint64_t)gpair, D, z);
                } else {
                    const double *ep = eps + ((size_t)j * (size_t)half_glob + (size_t)gpair) * (size_t)D;
#pragma unroll
                    for (int d = 0; d < DP; ++d) z[d] = (d < D) ? (float)__ldg(ep + d) : 0.0f;
                }
#pragma unroll
                for (int i = 0; i < H; ++i) e2[i] = make_float2(sj * z[2 * i], sj * z[2 * i + 1]);
            }
            float2 ee = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < H; ++i) ee = __ffma2_rn(e2[i], e2[i], ee);
            const float E = ee.x + ee.y;
            const float base = hj * E;

            float2 lp[ANYGRAD ? H : 1], lm[ANYGRAD ? H : 1];
            if constexpr (ANYGRAD) {
#pragma unroll
                for (int i = 0; i < H; ++i) lp[i] = lm[i] = make_float2(0.f, 0.f);
            }
            float qp = 0.f, qm = 0.f, Gp = 0.f, Gm = 0.f;

            // ---- un-flagged components: expanded form, branch-free, two components in flight ---------
            float *ust = Ucol;
#pragma unroll 2
            for (int k = 0; k < nfast; ++k) {
                const KFast c = sKc[k];
                const float2 *dl2 = reinterpret_cast<const float2 *>(sDl + k * DP);
                float2 dl[H];
#pragma unroll
                for (int i = 0; i < H; ++i) dl[i] = dl2[i];
                float2 b0 = make_float2(0.f, 0.f), b1 = b0;
#pragma unroll
                for (int i = 0; i + 1 < H; i += 2) {
                    b0 = __ffma2_rn(dl[i], e2[i], b0);
                    b1 = __ffma2_rn(dl[i + 1], e2[i + 1], b1);
                }
                if (H & 1) b0 = __ffma2_rn(dl[H - 1], e2[H - 1], b0);
                const float2 bb = __fadd2_rn(b0, b1);
                const float B = bb.x + bb.y;
                const float s0 = fmaf(c.hd, E, c.ck2);
                const float up = M<float>::ex2(fmaf(-c.h2, B, s0));
                const float um = M<float>::ex2(fmaf(c.h2, B, s0));
                if (WGRAD) {
                    ust[0] = up;
                    ust[K * kUS] = um;
                    ust += kUS;
                }
                qp = fmaf(c.w, up, qp);
                qm = fmaf(c.w, um, qm);
                if constexpr (ANYGRAD) {
                    const float gpv = c.wis2 * up, gmv = c.wis2 * um;
                    Gp += gpv;
                    Gm += gmv;
#pragma unroll
                    for (int i = 0; i < H; ++i) {
                        lp[i] = __ffma2_rn(make_float2(gpv, gpv), dl[i], lp[i]);
                        lm[i] = __ffma2_rn(make_float2(gmv, gmv), dl[i], lm[i]);
                    }
                }
            }
            // ---- flagged components: direct differences, squared term by term (no cancellation) -------
#pragma unroll 1
            for (int k = nfast; k < K; ++k) {
                const KFast c = sKc[k];
                const float2 *dl2 = reinterpret_cast<const float2 *>(sDl + k * DP);
                float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float2 tp = __fadd2_rn(dl2[i], e2[i]);
                    const float2 tm = __fadd2_rn(dl2[i], make_float2(-e2[i].x, -e2[i].y));
                    a0 = __ffma2_rn(tp, tp, a0);
                    a1 = __ffma2_rn(tm, tm, a1);
                }
                const float cb = c.ck + base;
                const float up = M<float>::ex2(fmaf(-c.h, a0.x + a0.y, cb));
                const float um = M<float>::ex2(fmaf(-c.h, a1.x + a1.y, cb));
                if (WGRAD) {
                    ust[0] = up;
                    ust[K * kUS] = um;
                    ust += kUS;
                }
                qp = fmaf(c.w, up, qp);
                qm = fmaf(c.w, um, qm);
                if constexpr (ANYGRAD) {
                    const float gpv = c.wis2 * up, gmv = c.wis2 * um;
#pragma unroll
                    for (int i = 0; i < H; ++i) {  // t recomputed: this loop is rare, registers are not
                        const float2 tp = __fadd2_rn(dl2[i], e2[i]);
                        const float2 tm = __fadd2_rn(dl2[i], make_float2(-e2[i].x, -e2[i].y));
                        lp[i] = __ffma2_rn(make_float2(gpv, gpv), tp, lp[i]);
                        lm[i] = __ffma2_rn(make_float2(gmv, gmv), tm, lm[i]);
                    }
                }
            }

            if (live) hacc += 0.69314718055994530942 * ((double)log2f(qp) + (double)log2f(qm)) - (double)E * is2j;
            if constexpr (ANYGRAD) {
                const float iqp = live ? __frcp_rn(qp) : 0.f, iqm = live ? __frcp_rn(qm) : 0.f;
                if (WGRAD) {
                    // fold this batch's tile: racc_k += sum_c u+[k][c] / q+[c] + u-[k][c] / q-[c]
                    sIq[tid] = iqp;
                    sIq[128 + tid] = iqm;
                    __syncwarp();
                    const float4 *ip4 = reinterpret_cast<const float4 *>(sIq + wid * 32);
                    const float4 *im4 = reinterpret_cast<const float4 *>(sIq + 128 + wid * 32);
#pragma unroll
                    for (int rr = 0; rr < RR; ++rr) {
                        const int row = lane + 32 * rr;
                        if (row < K) {
                            const float4 *up4 = tile4 + row * (kUS / 4);
                            const float4 *um4 = tile4 + (K + row) * (kUS / 4);
                            float a0 = 0.f, a1 = 0.f;
#pragma unroll
                            for (int c4 = 0; c4 < 8; ++c4) {
                                const float4 u = up4[c4], v = um4[c4], p4 = ip4[c4], m4 = im4[c4];
                                a0 = fmaf(u.x, p4.x, a0), a1 = fmaf(v.x, m4.x, a1);
                                a0 = fmaf(u.y, p4.y, a0), a1 = fmaf(v.y, m4.y, a1);
                                a0 = fmaf(u.z, p4.z, a0), a1 = fmaf(v.z, m4.z, a1);
                                a0 = fmaf(u.w, p4.w, a0), a1 = fmaf(v.w, m4.w, a1);
                            }
                            racc[rr] += a0 + a1;
                        }
                    }
                    __syncwarp();  // the tile is about to be reused
                }
                // gradient rows of this batch through the tile: row d = l+_d/q+ + l-_d/q-, row DP+d = e_d (..-..);
                // lane r sums row r over the 32 pairs in fp64 (cross-thread accumulation never sees fp32)
                const float2 ip2 = make_float2(iqp, iqp), im2 = make_float2(iqm, iqm);
                const float2 Gp2 = make_float2(Gp, Gp), nGm2 = make_float2(-Gm, -Gm);
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float2 a = __fmul2_rn(__ffma2_rn(e2[i], Gp2, lp[i]), ip2);    // l+ / q+
                    const float2 bq = __fmul2_rn(__ffma2_rn(e2[i], nGm2, lm[i]), im2);  // l- / q-
                    const float2 sa = __fadd2_rn(a, bq);
                    const float2 sb = __fmul2_rn(e2[i], __fadd2_rn(a, make_float2(-bq.x, -bq.y)));
                    Ucol[(2 * i) * kUS] = sa.x;
                    Ucol[(2 * i + 1) * kUS] = sa.y;
                    Ucol[(DP + 2 * i) * kUS] = sb.x;
                    Ucol[(DP + 2 * i + 1) * kUS] = sb.y;
                }
                __syncwarp();
#pragma unroll
                for (int rr = 0; rr < RR; ++rr) {
                    const int row = lane + 32 * rr;
                    if (row < 2 * DP) {
                        const float4 *r4 = tile4 + row * (kUS / 4);
                        float v0 = 0.f, v1 = 0.f;
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const float4 x = r4[c4];
                            v0 += x.x + x.y;
                            v1 += x.z + x.w;
                        }
                        rowacc[rr] += (double)v0 + (double)v1;
                    }
                }
                __syncwarp();
            }
        }

        // ---- segment record: each warp contributes its sums, one combine across the 4 warps -----------
        double *myrec = sRec + wid * part_stride;
        const double hs = warp_sum(hacc);
        if (lane == 0) myrec[0] = hs;
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int rr = 0; rr < RR; ++rr) {
                const int row = lane + 32 * rr;
                if (row < 2 * DP) myrec[1 + row] = rowacc[rr];
                if (WGRAD && row < K) myrec[1 + 2 * DP + sPerm[row]] = (double)racc[rr];  // un-permute
            }
        }
        __syncthreads();
        double *rec = part + ((size_t)blockIdx.x * maxseg + seg) * (size_t)part_stride;
        const int nf = ANYGRAD ? (WGRAD ? part_stride : 1 + 2 * DP) : 1;
        for (int f = tid; f < nf; f += 128)
            rec[f] = (sRec[f] + sRec[part_stride + f]) + (sRec[2 * part_stride + f] + sRec[3 * part_stride + f]);
    }
}

static size_t entmc_smem_w(int DP, int K, int nt, bool wgrad, bool anygrad) {
    (void)nt;
    const int urows = std::max(wgrad ? 2 * K : 0, anygrad ? 2 * DP : 0);
    size_t b = (size_t)K * DP * sizeof(float) + (size_t)K * sizeof(KFast) + (size_t)urows * kUS * sizeof(float) +
               2 * 128 * sizeof(float);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)4 * (1 + 2 * DP + K) * sizeof(double) + (size_t)DP * sizeof(double);
    b += (size_t)(K + 1) * sizeof(int) + (size_t)K * sizeof(float) + 16;
    return b;
}

static size_t entmc_smem_ds(int DP, int K, int nt, bool wgrad, bool anygrad) {
    const int DHP = ((DP / 2) + 3) & ~3;
    size_t b = (size_t)K * 2 * DHP * sizeof(float) + (size_t)K * sizeof(KConst<float>);
    size_t u = wgrad ? (size_t)2 * K * nt : 0;
    if (anygrad && u < (size_t)DP * nt) u = (size_t)DP * nt;
    u = (u + 3) & ~(size_t)3;
    b += u * sizeof(float);
    b = (b + 15) & ~(size_t)15;
    return b + 40 * sizeof(double);
}

static size_t entmc_smem_f32x2(int DP, int K, int nt, bool wgrad, bool anygrad) {
    size_t fl = (size_t)K * DP;
    size_t b = fl * sizeof(float) + (size_t)K * sizeof(KConst<float>);
    size_t u = wgrad ? (size_t)3 * K * nt : (anygrad ? (size_t)2 * DP * nt : 0);
    if (wgrad && u < (size_t)2 * DP * nt) u = (size_t)2 * DP * nt;
    u = (u + 3) & ~(size_t)3;
    b += u * sizeof(float);
    b = (b + 15) & ~(size_t)15;
    return b + 40 * sizeof(double);
}

template <typename T>
size_t entmc_smem(int DP, int K, int nt, bool wgrad, bool anygrad) {
    size_t b = (size_t)K * DP * sizeof(T) + (size_t)K * sizeof(KConst<T>);
    if (anygrad) b += (size_t)2 * DP * nt * sizeof(T);
    if (wgrad) b += (size_t)3 * K * nt * sizeof(T);
    b = (b + 15) & ~(size_t)15;
    return b + 40 * sizeof(double);
}

template <typename T, int DP, bool WGRAD, bool ANYGRAD, bool PHILOX>
int launch_inst(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, const double *d_eps,
                uint64_t seed, uint64_t offset, double *d_part) {
    dim3 grid(plan.slabs, lay.K);
    if (c->time_entmc) VBMC_CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
#define VBMC_LAUNCH(KERN, ...)                                                                                   \
    do {                                                                                                         \
        auto kern = KERN;                                                                                        \
        static size_t smem_set = 0; /* per instantiation: the attribute call costs microseconds */              \
        if (plan.smem > smem_set) {                                                                              \
            VBMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem)); \
            smem_set = plan.smem;                                                                                \
        }                                                                                                        \
        kern<<<grid, plan.threads, plan.smem, c->stream>>>(d_params, lay, plan.half, plan.pair0, plan.half_glob,  \
                                                           plan.pairs_per_thread, d_eps, seed, offset, d_part,   \
                                                           entpart_stride(DP, lay.K) __VA_ARGS__);               \
    } while (0)
    if constexpr (sizeof(T) == 4) {
        switch (plan.variant) {
            case ENTMC_WARP: {
                auto kern = entmc_kernel_w<DP, WGRAD, ANYGRAD, PHILOX>;
                static size_t smem_set = 0;
                if (plan.smem > smem_set) {
                    VBMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
                    smem_set = plan.smem;
                }
                kern<<<plan.grid, 128, plan.smem, c->stream>>>(d_params, lay, plan.half, plan.pair0, plan.half_glob,
                                                              plan.chunk, plan.maxseg, d_eps, seed, offset, d_part,
                                                              entpart_stride(DP, lay.K), c->entmc_guard);
                break;
            }
            case ENTMC_FAST:
                VBMC_LAUNCH((entmc_kernel_fast<DP, WGRAD, ANYGRAD, PHILOX>), , c->entmc_guard);
                break;
            case ENTMC_SMALL:
                VBMC_LAUNCH((entmc_kernel_small<DP, WGRAD, ANYGRAD, PHILOX>), , c->entmc_guard);
                break;
            case ENTMC_DSPLIT:
                VBMC_LAUNCH((entmc_kernel_f32x2_ds<DP, WGRAD, ANYGRAD, PHILOX>));
                break;
            case ENTMC_PACKED:
                VBMC_LAUNCH((entmc_kernel_f32x2<DP, WGRAD, ANYGRAD, PHILOX>));
                break;
            default:
                VBMC_LAUNCH((entmc_kernel<float, DP, WGRAD, ANYGRAD, PHILOX>));
        }
    } else {
        VBMC_LAUNCH((entmc_kernel<T, DP, WGRAD, ANYGRAD, PHILOX>));
    }
#undef VBMC_LAUNCH
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    if (c->time_entmc) {
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
        VBMC_CUDA_CHECK(cudaEventSynchronize(c->ev1));
        float ms = 0;
        VBMC_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->entmc_ms_sum += ms;
        c->entmc_ms_n++;
    }
    return VBMC_OK;
}

template <typename T, int DP>
int launch_dp(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, bool anygrad, bool wgrad,
              bool philox, const double *d_eps, uint64_t seed, uint64_t offset, double *d_part) {
#define VBMC_GO(W, A, P) return launch_inst<T, DP, W, A, P>(c, d_params, lay, plan, d_eps, seed, offset, d_part)
    if (wgrad) {
        if (philox) VBMC_GO(true, true, true);
        VBMC_GO(true, true, false);
    }
    if (anygrad) {
        if (philox) VBMC_GO(false, true, true);
        VBMC_GO(false, true, false);
    }
    if (philox) VBMC_GO(false, false, true);
    VBMC_GO(false, false, false);
#undef VBMC_GO
}

template <typename T>
int launch_t(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, bool anygrad, bool wgrad,
             bool philox, const double *d_eps, uint64_t seed, uint64_t offset, double *d_part) {
    switch (lay.DP) {
#define VBMC_CASE(N) \
    case N:          \
        return launch_dp<T, N>(c, d_params, lay, plan, anygrad, wgrad, philox, d_eps, seed, offset, d_part)
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
    set_error("entmc: unsupported padded dimension");
    return VBMC_ERR_UNSUPPORTED;
}

template <int DP>
__global__ void philox_dump_kernel(int D, int K, int64_t half, uint64_t seed, uint64_t offset, double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * half) return;
    const int j = (int)(i / half);
    const int64_t p = i - (int64_t)j * half;
    float z[DP];
    philox_normals<DP>(seed, offset, (uint32_t)j, (uint64_t)p, D, z);
#pragma unroll
    for (int d = 0; d < DP; ++d)
        if (d < D) out[(size_t)i * D + d] = (double)z[d];
}

}  // namespace

// Choose threads / pairs-per-thread so that the grid is a whole number of waves of the
// resident-CTA capacity whenever possible (148 SMs x CTAs that fit by shared memory).
static size_t entmc_smem_variant(int variant, int precision, int DP, int K, int nt, bool wgrad, bool anygrad) {
    if (precision == VBMC_PREC_F64) return entmc_smem<double>(DP, K, nt, wgrad, anygrad);
    switch (variant) {
        case ENTMC_WARP:
            return entmc_smem_w(DP, K, nt, wgrad, anygrad);
        case ENTMC_FAST:
            return entmc_smem_fast(DP, K, nt, wgrad, anygrad);
        case ENTMC_SMALL:
            return entmc_smem_small(DP, K, nt, wgrad, anygrad);
        case ENTMC_DSPLIT:
            return entmc_smem_ds(DP, K, nt, wgrad, anygrad);
        case ENTMC_PACKED:
            return entmc_smem_f32x2(DP, K, nt, wgrad, anygrad);
        default:
            return entmc_smem<float>(DP, K, nt, wgrad, anygrad);
    }
}

int entmc_plan(const Ctx *c, int D, int K, int64_t half_local, bool wgrad, int precision, EntmcPlan *plan) {
    const int DP = pad_dim(D);
    VBMC_REQUIRE(DP > 0, VBMC_ERR_UNSUPPORTED, "entmc: D > 32 is not supported");
    VBMC_REQUIRE(half_local >= 0, VBMC_ERR_ARG, "entmc: negative draw count");
    int variant = precision == VBMC_PREC_F64 ? ENTMC_SCALAR : c->entmc_variant;
    if (variant < 0) {
        // auto: the tensor-core kernel wins once every SM gets several 128-pair tiles; the warp-autonomous kernel pays
        // off once there is a wave of >= 4-batch CTAs
        const int64_t T = (int64_t)K * half_local;
        // Measured in round 2 (scripts/variant_sweep.py, profiles/r3d_variant_sweep.txt; ~400k draws): the tensor-core
        // MAIN kernel beats both CUDA-core kernels at every (D, K) tried; what it adds is the table kernel (~7 us) in
        // front of it -- its noise generator runs beside the previous tail (device-resident loops) or beside the host's
        // preparation of the call (vbmc_noise_prefetch).  Table + main wins from K >= 32 (D = 6, 10, 20, 32) and from
        // K >= 24 at D >= 16; below that the warp-autonomous / expanded kernels stay ahead (C2: D=10, K=20; C4: D=6, K=30).
        const bool tc_shape = K >= 32 || (D >= 16 && K >= 24);
        if (T >= 150000 && tc_shape && entmc_tc_supported(DP, K)) variant = ENTMC_TC;
        else variant = T >= 65536 ? ENTMC_WARP : ENTMC_FAST;
        // the reference's default draw counts (tens of draws per component): eight lanes per pair instead of one thread
        static const int64_t small_half = getenv("VBMC_SMALL_HALF") ? atoll(getenv("VBMC_SMALL_HALF")) : 64;
        if (variant == ENTMC_FAST && half_local <= small_half) variant = ENTMC_SMALL;
    }
    const size_t smem_cap = 227 * 1024;
    if (variant == ENTMC_TC) {
        if (entmc_tc_supported(DP, K)) return entmc_tc_plan(c, D, K, half_local, plan);
        variant = ENTMC_WARP;
    }
    if (variant == ENTMC_WARP) {
        // the warp-autonomous kernel keeps racc rows lane + 32 r (r < 4) in registers and needs its tile in smem
        const size_t smem = entmc_smem_w(DP, K, 128, wgrad, true);
        if (K > 128 || smem > smem_cap) {
            variant = ENTMC_FAST;
        } else {
            int per_sm = (int)(smem_cap / (smem + 1024));
            const int reg_cap = DP <= 20 ? 3 : 2;  // __launch_bounds__ of entmc_kernel_w
            if (per_sm > reg_cap) per_sm = reg_cap;
            if (per_sm < 1) per_sm = 1;
            const int64_t T = (int64_t)K * half_local;
            const int64_t slots = (int64_t)c->sm_count * per_sm;
            // one wave over all resident slots; whole batches for all four warps (chunk multiple of 128).
            // Small problems still get one batch per warp and as many CTAs as that allows: resident warps
            // matter more than amortising the table set-up.
            int64_t chunk = (T + slots - 1) / slots;
            chunk = ((chunk + 127) / 128) * 128;
            if (chunk < 128) chunk = 128;
            const int64_t G = std::max<int64_t>(1, (T + chunk - 1) / chunk);
            plan->variant = variant;
            plan->threads = 128;
            plan->pairs_per_thread = (int)(chunk / 128);
            plan->chunk = chunk;
            plan->grid = (int)G;
            plan->maxseg = half_local > 0 ? (int)((chunk - 1) / half_local) + 2 : 1;
            plan->slabs = plan->grid * plan->maxseg;  // records allocated (not all used)
            plan->half = half_local;
            plan->pair0 = 0;
            plan->half_glob = half_local;
            plan->smem = smem;
            return VBMC_OK;
        }
    }
    int nt = 128;
    auto smem_of = [&](int t) { return entmc_smem_variant(variant, precision, DP, K, t, wgrad, true); };
    while (nt > 32 && smem_of(nt) > smem_cap) nt >>= 1;
    VBMC_REQUIRE(smem_of(nt) <= smem_cap, VBMC_ERR_UNSUPPORTED, "entmc: K too large for shared memory");
    const size_t smem = smem_of(nt);
    int per_sm = (int)(smem_cap / (smem + 1024));
    // register-limited CTAs per SM (see __launch_bounds__ of the kernels)
    const int reg_limit = ((precision != VBMC_PREC_F64 && variant == ENTMC_DSPLIT) ? 4 : 2) * (128 / nt);
    if (per_sm > reg_limit) per_sm = reg_limit;
    if (per_sm < 1) per_sm = 1;
    const int64_t slots = (int64_t)c->sm_count * per_sm;
    // pairs evaluated per CTA sweep: one per thread, or one per two lanes (dimension-split kernel)
    const int ppi = precision == VBMC_PREC_F64 ? nt : (variant == ENTMC_DSPLIT ? nt / 2 : (variant == ENTMC_SMALL ? nt / 8 : nt));

    // candidate R: cost ~ waves * (R + overhead); overhead ~ table set-up + record reduction
    int bestR = 1;
    double best = 1e300;
    const double overhead = 0.35;
    for (int R = 1; R <= 64; ++R) {
        const int64_t slabs = (half_local + (int64_t)ppi * R - 1) / ((int64_t)ppi * R);
        const int64_t ctas = slabs * K;
        const int64_t waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * (R + overhead);
        if (cost < best - 1e-9) best = cost, bestR = R;
        if (slabs <= 1) break;
    }
    plan->variant = variant;
    plan->chunk = 0, plan->grid = 0, plan->maxseg = 0;
    plan->threads = nt;
    plan->pairs_per_thread = bestR;
    plan->slabs = (int)((half_local + (int64_t)ppi * bestR - 1) / ((int64_t)ppi * bestR));
    if (plan->slabs < 1) plan->slabs = 1;
    plan->half = half_local;
    plan->pair0 = 0;
    plan->half_glob = half_local;
    plan->smem = smem;
    return VBMC_OK;
}

int entmc_launch(Ctx *c, const double *d_params, int D, int K, const EntmcPlan &plan, bool anygrad, bool wgrad,
                 int precision, int rng_mode, const double *d_eps, uint64_t seed, uint64_t offset, double *d_part) {
    ParamLayout lay{D, pad_dim(D), K};
    const bool philox = rng_mode == VBMC_RNG_PHILOX;
    VBMC_REQUIRE(philox || d_eps != nullptr, VBMC_ERR_ARG, "entmc: eps required in VBMC_RNG_EPS mode");
    VBMC_REQUIRE(plan.half_glob < ((int64_t)1 << 32), VBMC_ERR_UNSUPPORTED, "entmc: more than 2^32 pairs per component");
    // smem of the plan was sized for anygrad; recompute for the actual instantiation
    EntmcPlan p = plan;
    p.smem = entmc_smem_variant(plan.variant, precision, lay.DP, K, plan.threads, wgrad, anygrad);
    if (precision == VBMC_PREC_F64)
        return launch_t<double>(c, d_params, lay, p, anygrad, wgrad, philox, d_eps, seed, offset, d_part);
    if (plan.variant == ENTMC_TC) {
        if (c->time_entmc) VBMC_CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
        VBMC_TRY(entmc_tc_launch(c, d_params, lay, plan, anygrad, philox, d_eps, d_part));
        VBMC_CUDA_CHECK(cudaGetLastError());
        c->launches++;
        if (c->time_entmc) {
            VBMC_CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
            VBMC_CUDA_CHECK(cudaEventSynchronize(c->ev1));
            float ms = 0;
            VBMC_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            c->entmc_ms_sum += ms;
            c->entmc_ms_n++;
            if (c->ev2_recorded) {
                VBMC_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev2, c->ev1));
                c->entmc_main_ms_sum += ms;
                c->ev2_recorded = false;
            }
        }
        return VBMC_OK;
    }
    return launch_t<float>(c, d_params, lay, p, anygrad, wgrad, philox, d_eps, seed, offset, d_part);
}

int philox_normals_launch(Ctx *c, int D, int K, int64_t half, uint64_t seed, uint64_t offset, double *d_eps) {
    const int DP = pad_dim(D);
    VBMC_REQUIRE(DP > 0, VBMC_ERR_UNSUPPORTED, "philox: D > 32 is not supported");
    const int64_t n = (int64_t)K * half;
    if (n == 0) return VBMC_OK;
    const int nt = 128;
    const unsigned grid = (unsigned)((n + nt - 1) / nt);
    switch (DP) {
#define VBMC_CASE(N)                                                                               \
    case N:                                                                                        \
        philox_dump_kernel<N><<<grid, nt, 0, c->stream>>>(D, K, half, seed, offset, d_eps);        \
        break
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
    VBMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    return VBMC_OK;
}

}  // namespace vbmc

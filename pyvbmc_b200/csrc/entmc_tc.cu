// entmc_tc.cu -- Monte-Carlo mixture entropy (pyvbmc/entropy/entmc_vbmc.py:39-134) with the two
// GEMM-shaped contractions of the inner loop on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// For the draws x(+-) = mu_j + sigma_j lambda (.) (+-e) of component j, scaled by 1/lambda:
//     |t(+-)_k|^2 = |Delta_k|^2 + |e|^2 (+-) 2 B_k,       B_k = Delta_k . e,   Delta_k = (mu_j - mu_k) / lambda
//     u(+-)_k     = N_k(x+-) / N_j(x+-) = 2^(s0_k -+ h2_k B_k)                 (density ratios, log2 domain)
//     q(+-)       = sum_k w_k u(+-)_k
// and the reparameterisation gradient (entmc_vbmc.py:84-112) needs, per antithetic pair,
//     racc_k  += u+_k / q+ + u-_k / q-                                          (d/dw, and d/dmu through Delta)
//     v_d      = sum_k Delta_kd c_k ,   c_k = (w_k / sigma_k^2) (u+_k / q+ - u-_k / q-)
// B = E Delta^T  ([pairs x D] x [D x K])  and  V = C Delta  ([pairs x K] x [K x D])  are GEMMs.  A CTA owns tiles of
// 128 antithetic pairs (thread t <-> pair t <-> TMEM lane t):
//   GEMM1  B[128 x KP]  = E[128 x D8] Delta^T      tcgen05.mma kind::tf32, A and B from shared memory (K-major,
//                                                  no-swizzle canonical layout), 3xTF32 split (hi*hi + hi*lo + lo*hi)
//   pass 1 tcgen05.ld B rows -> u(+-) (2 MUFU.EX2 per (pair, k)), q(+-), G(+-); u(+-) parked in TMEM (tcgen05.st)
//   pass 2 u(+-) -> racc_k (accumulated in TMEM across the tiles of a segment), c_k split hi/lo -> TMEM
//   GEMM2  V[128 x N2]  = C[128 x KP] Delta        A operand straight from TMEM, B from shared memory, 3xTF32
//   epilogue: per-thread sums of e_d (v_d + e_d (G+/q+ + G-/q-)) and e_d (G+/q+ - G-/q-)
// which leaves ~20 CUDA-core instructions per (pair, component) instead of ~50 (30 of them packed FFMA2) in the
// CUDA-core kernels of entmc.cu.  Components whose expanded distance is badly conditioned (same guard as
// entmc_kernel_w) have their u(+-) recomputed with direct differences on the CUDA cores.
// Work distribution, records and determinism are those of entmc_kernel_w: the K * half pairs are one index space
// cut into equal chunks (one per CTA, a multiple of 128 pairs), a chunk that straddles components is processed
// segment by segment, ONE fp64 record [hacc | A_d | Be_d | racc_k] per (CTA, segment), no atomics.
#include "common.cuh"
#include "philox.cuh"

namespace vbmc {
namespace {

struct alignas(16) KTc {
    float ck2, h2, hd, w;  // s0_k = hd E + ck2 ;  log2 u(+-) = s0_k -+ h2 B_k
};
struct alignas(8) KDir {
    float ck, h;  // direct path: log2 u = ck + hj E - h |t|^2
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 2000000000LL) __trap();  // ~1 s: a lost MMA completion must not hang the GPU
    }
}

// ---- tcgen05 ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// shared-memory operand descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor; canonical layout in 16-byte
// units ((8, n), 2) : ((1, SBO), LBO)): 8 rows x 16 B core matrices, rows 16 B apart, the next 8-row group SBO
// bytes further, the second 16-byte chunk along K LBO bytes further.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_tf32(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

#define TM_R16(v) \
    "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define TM_W16(v) \
    "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), \
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])

// 16 consecutive fp32 columns of this thread's TMEM lane (load + wait in ONE asm statement: the registers
// are valid when it returns, whatever the compiler schedules around it)
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : TM_R16(v)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), TM_W16(v)
        : "memory");
}
#define TM_R8(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
#define TM_W8(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
__device__ __forceinline__ void tm_ld8x3(uint32_t ta, uint32_t (&a)[8], uint32_t tb, uint32_t (&b)[8], uint32_t tc,
                                         uint32_t (&c)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : TM_R8(a), TM_R8(b), TM_R8(c)
        : "r"(ta), "r"(tb), "r"(tc)
        : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), TM_W8(v) : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kTile = 128;     // antithetic pairs per tile = TMEM lanes
constexpr int kThreads = 256;  // two threads per pair: each owns half of the dimensions and half of the components

struct TcSmem {  // byte offsets inside the dynamic shared memory of entmc_kernel_tc
    uint32_t Ah, Al, B1h, B1l, B2h, B2l, Dl, Mu, Kc, Wis, Dir, Mask, E, Q, Rec, Tot, InvL, Bar, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int DP, int D, int K, int part_stride) {
    const int D8 = (DP + 7) / 8 * 8, NC1 = D8 / 4, N2 = D8 <= 16 ? 16 : 32;
    const int KP = (K + 15) / 16 * 16;
    TcSmem s;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        const uint32_t at = o;
        o += (bytes + 127u) & ~127u;
        return at;
    };
    s.Ah = take(NC1 * kTile * 16);
    s.Al = take(NC1 * kTile * 16);
    s.B1h = take(NC1 * KP * 16);
    s.B1l = take(NC1 * KP * 16);
    s.B2h = take((KP / 4) * N2 * 16);
    s.B2l = take((KP / 4) * N2 * 16);
    s.Dl = take(K * DP * 4);
    s.Mu = take(K * D * 8);
    s.Kc = take(KP * sizeof(KTc));
    s.Wis = take(KP * 4);
    s.Dir = take(KP * sizeof(KDir));
    s.Mask = take(16 * 4);
    s.E = take(2 * kTile * 4);
    s.Q = take(2 * 4 * kTile * 4);
    s.Rec = take(8 * part_stride * 8);
    s.Tot = take(part_stride * 8);
    s.InvL = take(DP * 8);
    s.Bar = take(64);
    s.total = o;
    return s;
}

template <int N>
struct TmLd;  // N consecutive fp32 columns of this thread's TMEM lane
template <>
struct TmLd<16> {
    static __device__ __forceinline__ void go(uint32_t taddr, uint32_t (&v)[16]) { tm_ld16(taddr, v); }
};
template <>
struct TmLd<8> {
    static __device__ __forceinline__ void go(uint32_t taddr, uint32_t (&v)[8]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
            : "r"(taddr)
            : "memory");
    }
};

template <int DP, bool ANYGRAD, bool PHILOX>
__global__ void __launch_bounds__(kThreads, 2)
entmc_kernel_tc(const double *__restrict__ prm, ParamLayout lay, int64_t half, int64_t pair0, int64_t half_glob,
                int64_t chunk, int maxseg, const double *__restrict__ eps, double *__restrict__ part, int part_stride,
                float guard, uint32_t tmem_cols, int desc_swap) {
    constexpr int DH = DP / 2;              // dimensions per thread
    constexpr int D8 = (DP + 7) / 8 * 8;    // GEMM1 reduction length (tf32 MMAs consume 8 per instruction)
    constexpr int NC1 = D8 / 4;             // 16-byte chunks along D
    constexpr int N2 = D8 <= 16 ? 16 : 32;  // GEMM2 output columns (M = 128 needs N % 16 == 0)
    constexpr int LW = DH <= 8 ? 8 : 16;    // columns of V loaded per thread in the epilogue (DH + LW <= N2)
    const int D = lay.D, K = lay.K;
    const int nch = (K + 15) >> 4, KP = nch * 16;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = tid & (kTile - 1), hsel = tid >> 7, quad = wid & 3;
    // components (16-column chunks) of this thread: the first half of the chunks for hsel = 0, the rest for 1
    const int c_begin = hsel ? (nch + 1) / 2 : 0, c_end = hsel ? nch : (nch + 1) / 2;
    uint64_t seed, offset;
    {
        const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
        seed = rngp[0], offset = rngp[1];
    }

    extern __shared__ __align__(128) unsigned char smem[];
    const TcSmem L = tc_smem_layout(DP, D, K, part_stride);
    float *sAh = reinterpret_cast<float *>(smem + L.Ah), *sAl = reinterpret_cast<float *>(smem + L.Al);
    float *sB1h = reinterpret_cast<float *>(smem + L.B1h), *sB1l = reinterpret_cast<float *>(smem + L.B1l);
    float *sB2h = reinterpret_cast<float *>(smem + L.B2h), *sB2l = reinterpret_cast<float *>(smem + L.B2l);
    float *sDl = reinterpret_cast<float *>(smem + L.Dl);
    double *sMu = reinterpret_cast<double *>(smem + L.Mu);
    KTc *sKc = reinterpret_cast<KTc *>(smem + L.Kc);
    float *sWis = reinterpret_cast<float *>(smem + L.Wis);
    KDir *sDir = reinterpret_cast<KDir *>(smem + L.Dir);
    uint32_t *sMask = reinterpret_cast<uint32_t *>(smem + L.Mask);
    float *sE = reinterpret_cast<float *>(smem + L.E);  // [2][128] partial |e|^2
    float *sQ = reinterpret_cast<float *>(smem + L.Q);  // [2][4][128] partial q+, q-, G+, G-
    double *sRec = reinterpret_cast<double *>(smem + L.Rec);
    double *sTot = reinterpret_cast<double *>(smem + L.Tot);
    double *sInvL = reinterpret_cast<double *>(smem + L.InvL);
    uint64_t *sBar = reinterpret_cast<uint64_t *>(smem + L.Bar);
    uint32_t *sTmem = reinterpret_cast<uint32_t *>(sBar + 2);

    const double *sigma = prm + lay.sigma();
    const double *lambd = prm + lay.lambd();
    const double *w = prm + lay.w();
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2

    const int64_t T = (int64_t)K * half;
    const int64_t g0 = (int64_t)blockIdx.x * chunk, g1 = min(g0 + chunk, T);
    if (g0 >= T) return;

    // ---- one-time set-up: barriers, tensor memory, component means, zeroed operand tile --------------------
    const uint32_t bar0 = smem_u32(sBar), bar1 = smem_u32(sBar + 1);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < DP) sInvL[tid] = tid < D ? 1.0 / lambd[tid] : 0.0;
    {
        const double *mu = prm + lay.mu();
        for (int i = tid; i < K * D; i += kThreads) sMu[i] = mu[i];
        for (int i = tid; i < NC1 * kTile * 4; i += kThreads) sAh[i] = 0.f, sAl[i] = 0.f;  // padding dims stay 0
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);  // this warp's lane quadrant
    // TMEM columns: R0 = B -> u+ -> c_hi, R1 = u- -> c_lo, R2 = racc, D2 = V
    const uint32_t cR0 = 0, cR1 = KP, cR2 = 2 * KP, cD2 = 3 * KP;
    uint32_t ph0 = 0, ph1 = 0;

    const uint32_t idesc1 = umma_idesc_tf32(KP), idesc2 = umma_idesc_tf32(N2);
    // operand strides: rows 16 B apart, 8-row groups 128 B apart, K chunks one whole row-block apart
    const uint32_t lboA = kTile * 16, lboB1 = KP * 16, lboB2 = N2 * 16, sbo = 128;
    (void)desc_swap;

    const int j_first = (int)(g0 / half);
    for (int seg = 0;; ++seg) {
        const int j = j_first + seg;
        const int64_t lo = max(g0, (int64_t)j * half), hi = min(g1, (int64_t)(j + 1) * half);
        if (j >= K || lo >= hi) break;
        const int64_t p_lo = lo - (int64_t)j * half;
        const int n = (int)(hi - lo);

        // ---- component tables ------------------------------------------------------------------------
        __syncthreads();
        const double sig_j = sigma[j];
        const double hjd = kHalfLog2e / (sig_j * sig_j);
        const double Emax = sig_j * sig_j * (D + 8.0 * sqrt(2.0 * D) + 32.0);
        if (tid < 16) sMask[tid] = 0u;
        for (int i = tid; i < K * DP; i += kThreads) {
            const int k = i / DP, d = i - k * DP;
            sDl[i] = (d < D) ? (float)((sMu[j * D + d] - sMu[k * D + d]) * sInvL[d]) : 0.0f;
        }
        __syncthreads();
        for (int k = tid; k < KP; k += kThreads) {
            KTc c;
            KDir dr;
            float wis = 0.f;
            c.ck2 = -200.0f, c.h2 = 0.f, c.hd = 0.f, c.w = 0.f;  // padding / flagged: expanded-form u = 2^-200 = 0
            dr.ck = -200.0f, dr.h = 0.f;
            if (k < K) {
                const double sk = sigma[k];
                const double hk = kHalfLog2e / (sk * sk);
                const double ck = D * (log2(sig_j) - log2(sk));
                double A = 0.0;  // |Delta_k|^2 of the rounded table entries
                for (int d = 0; d < D; ++d) A += (double)sDl[k * DP + d] * (double)sDl[k * DP + d];
                c.w = (float)w[k];
                wis = (float)(w[k] / (sk * sk));
                dr.ck = (float)ck;
                dr.h = (float)hk;
                // conditioning of the expanded form (see entmc_kernel_fast): direct differences beyond the guard
                if (k != j && hk * (A + Emax) > (double)guard) {
                    atomicOr(&sMask[k >> 4], 1u << (k & 15));
                } else {
                    c.ck2 = (float)(ck - hk * A);
                    c.h2 = (float)(2.0 * hk);
                    c.hd = (float)(hjd - hk);
                }
            }
            sKc[k] = c;
            sWis[k] = wis;
            sDir[k] = dr;
        }
        // operand tables of the two GEMMs (hi = upper 19 bits = exact tf32, lo = remainder)
        for (int i = tid; i < KP * D8; i += kThreads) {
            const int k = i / D8, d = i - k * D8;
            const float v = (k < K && d < DP) ? sDl[k * DP + d] : 0.0f;
            const float vh = __uint_as_float(__float_as_uint(v) & 0xffffe000u), vl = v - vh;
            const int o1 = (d >> 2) * (KP * 4) + k * 4 + (d & 3);  // GEMM1 B operand: rows = components, K dim = d
            sB1h[o1] = vh, sB1l[o1] = vl;
            if (ANYGRAD) {
                const int o2 = (k >> 2) * (N2 * 4) + d * 4 + (k & 3);  // GEMM2 B operand: rows = d, K dim = components
                sB2h[o2] = vh, sB2l[o2] = vl;
            }
        }
        if constexpr (ANYGRAD && N2 > D8) {
            constexpr int NZ = N2 - D8;
            for (int i = tid; i < KP * NZ; i += kThreads) {
                const int k = i / NZ, d = D8 + (i - k * NZ);
                const int o2 = (k >> 2) * (N2 * 4) + d * 4 + (k & 3);
                sB2h[o2] = 0.f, sB2l[o2] = 0.f;
            }
        }
        if (ANYGRAD) {  // racc = 0 in TMEM (own chunks)
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            for (int ci = c_begin; ci < c_end; ++ci) tm_st16(trow + cR2 + 16 * ci, z);
            tm_wait_st();
        }
        __syncthreads();

        const float hj = (float)hjd;
        const double is2j = 1.0 / (sig_j * sig_j);
        const float sj = (float)sig_j;
        double hacc = 0.0;
        float ae[ANYGRAD ? DH : 1], be[ANYGRAD ? DH : 1];
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int i = 0; i < DH; ++i) ae[i] = be[i] = 0.f;
        }

        for (int t0 = 0; t0 < n; t0 += kTile) {
            const int off = t0 + row;
            const bool live = off < n;
            const int64_t gpair = pair0 + p_lo + (live ? off : 0);

            // ---- 1. this thread's half of the noise of pair `row`, operand tile of GEMM1 ---------------------
            float e[DH];
            {
                float z[DH];
                if (PHILOX) {
                    philox_normals_half<DH>(seed, offset, (uint32_t)j, (uint64_t)gpair, hsel, D, z);
                } else {
                    const double *ep = eps + ((size_t)j * (size_t)half_glob + (size_t)gpair) * (size_t)D;
#pragma unroll
                    for (int i = 0; i < DH; ++i) z[i] = (hsel * DH + i < D) ? (float)__ldg(ep + hsel * DH + i) : 0.0f;
                }
#pragma unroll
                for (int i = 0; i < DH; ++i) e[i] = live ? sj * z[i] : 0.0f;
            }
            {
                float Eh = 0.f;
#pragma unroll
                for (int i = 0; i < DH; ++i) {
                    Eh = fmaf(e[i], e[i], Eh);
                    const int d = hsel * DH + i;
                    const float vh = __uint_as_float(__float_as_uint(e[i]) & 0xffffe000u);
                    const int o = (d >> 2) * (kTile * 4) + row * 4 + (d & 3);
                    sAh[o] = vh, sAl[o] = e[i] - vh;
                }
                sE[hsel * kTile + row] = Eh;
            }
            fence_async_smem();  // generic-proxy writes (tile, and the tables at a segment start) -> async proxy
            tc_fence_before();
            __syncthreads();

            // ---- 2. GEMM1: B = E Delta^T (3xTF32) -----------------------------------------------------------
            if (tid == 0) {
                tc_fence_after();
                const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl), bH = smem_u32(sB1h), bL = smem_u32(sB1l);
#pragma unroll
                for (int s = 0; s < D8 / 8; ++s) {
                    const uint64_t dAh = umma_desc(aH + s * 2 * lboA, lboA, sbo), dAl = umma_desc(aL + s * 2 * lboA, lboA, sbo);
                    const uint64_t dBh = umma_desc(bH + s * 2 * lboB1, lboB1, sbo), dBl = umma_desc(bL + s * 2 * lboB1, lboB1, sbo);
                    mma_ss(tmem + cR0, dAh, dBh, idesc1, s > 0 ? 1u : 0u);
                    mma_ss(tmem + cR0, dAh, dBl, idesc1, 1u);
                    mma_ss(tmem + cR0, dAl, dBh, idesc1, 1u);
                }
                tc_commit(bar0);
            }
            const float E = sE[row] + sE[kTile + row];
            mbar_wait(bar0, ph0);
            ph0 ^= 1u;
            tc_fence_after();

            // ---- 3. pass 1: density ratios u(+-) of this thread's components, partial mixture sums --------
            float qp = 0.f, qm = 0.f, Gp = 0.f, Gm = 0.f;
            for (int ci = c_begin; ci < c_end; ++ci) {
                uint32_t b[16], um[16];
                tm_ld16(trow + cR0 + 16 * ci, b);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = 16 * ci + i;
                    const KTc c = sKc[k];
                    const float wis = sWis[k];
                    const float s0 = fmaf(c.hd, E, c.ck2);
                    const float x = c.h2 * __uint_as_float(b[i]);
                    const float vp = ex2f(s0 - x), vm = ex2f(s0 + x);
                    qp = fmaf(c.w, vp, qp), qm = fmaf(c.w, vm, qm);
                    Gp = fmaf(wis, vp, Gp), Gm = fmaf(wis, vm, Gm);
                    b[i] = __float_as_uint(vp), um[i] = __float_as_uint(vm);
                }
                const uint32_t fm = sMask[ci];
                if (fm != 0u) {  // CTA-uniform, rare: badly conditioned components, direct differences
                    const float base = hj * E;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if ((fm >> i) & 1u) {
                            const int k = 16 * ci + i;
                            const KDir dr = sDir[k];
                            const float *dl = sDl + k * DP;
                            float a0 = 0.f, a1 = 0.f;
#pragma unroll 1
                            for (int d = 0; d < DP; ++d) {
                                const int o = (d >> 2) * (kTile * 4) + row * 4 + (d & 3);
                                const float ed = sAh[o] + sAl[o];
                                const float tp = dl[d] + ed, tm = dl[d] - ed;
                                a0 = fmaf(tp, tp, a0), a1 = fmaf(tm, tm, a1);
                            }
                            const float vp = ex2f(fmaf(-dr.h, a0, dr.ck + base));
                            const float vm = ex2f(fmaf(-dr.h, a1, dr.ck + base));
                            const float wk = sKc[k].w, wis = sWis[k];
                            qp = fmaf(wk, vp, qp), qm = fmaf(wk, vm, qm);
                            Gp = fmaf(wis, vp, Gp), Gm = fmaf(wis, vm, Gm);
                            b[i] = __float_as_uint(vp), um[i] = __float_as_uint(vm);
                        }
                    }
                }
                if constexpr (ANYGRAD) {
                    tm_st16(trow + cR0 + 16 * ci, b);
                    tm_st16(trow + cR1 + 16 * ci, um);
                }
            }
            // exchange the partial sums of the two threads of a pair
            sQ[(hsel * 4 + 0) * kTile + row] = qp, sQ[(hsel * 4 + 1) * kTile + row] = qm;
            if constexpr (ANYGRAD) {
                sQ[(hsel * 4 + 2) * kTile + row] = Gp, sQ[(hsel * 4 + 3) * kTile + row] = Gm;
                tm_wait_st();
            }
            __syncthreads();
            qp = sQ[0 * kTile + row] + sQ[4 * kTile + row];
            qm = sQ[1 * kTile + row] + sQ[5 * kTile + row];
            if (live && hsel == 0)
                hacc += 0.69314718055994530942 * ((double)log2f(qp) + (double)log2f(qm)) - (double)E * is2j;

            if constexpr (ANYGRAD) {
                Gp = sQ[2 * kTile + row] + sQ[6 * kTile + row];
                Gm = sQ[3 * kTile + row] + sQ[7 * kTile + row];
                const float iqp = live ? __frcp_rn(qp) : 0.f, iqm = live ? __frcp_rn(qm) : 0.f;
                // ---- 4. pass 2: racc_k += u+/q+ + u-/q- ;  c_k = wis2_k (u+/q+ - u-/q-) split hi/lo --------------
                for (int c8 = 2 * c_begin; c8 < 2 * c_end; ++c8) {  // 8 components at a time (register pressure)
                    uint32_t up[8], um[8], ra[8];
                    tm_ld8x3(trow + cR0 + 8 * c8, up, trow + cR1 + 8 * c8, um, trow + cR2 + 8 * c8, ra);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float a = __uint_as_float(up[i]) * iqp, bq = __uint_as_float(um[i]) * iqm;
                        ra[i] = __float_as_uint(__uint_as_float(ra[i]) + (a + bq));
                        const float c = sWis[8 * c8 + i] * (a - bq);
                        const uint32_t ch = __float_as_uint(c) & 0xffffe000u;
                        up[i] = ch;
                        um[i] = __float_as_uint(c - __uint_as_float(ch));
                    }
                    tm_st8(trow + cR0 + 8 * c8, up);
                    tm_st8(trow + cR1 + 8 * c8, um);
                    tm_st8(trow + cR2 + 8 * c8, ra);
                }
                tm_wait_st();
                tc_fence_before();
                __syncthreads();

                // ---- 5. GEMM2: V = C Delta (A from TMEM, 3xTF32) ---------------------------------------------
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t bH = smem_u32(sB2h), bL = smem_u32(sB2l);
                    for (int s = 0; s < KP / 8; ++s) {
                        const uint64_t dBh = umma_desc(bH + s * 2 * lboB2, lboB2, sbo), dBl = umma_desc(bL + s * 2 * lboB2, lboB2, sbo);
                        mma_ts(tmem + cD2, tmem + cR0 + 8 * s, dBh, idesc2, s > 0 ? 1u : 0u);
                        mma_ts(tmem + cD2, tmem + cR0 + 8 * s, dBl, idesc2, 1u);
                        mma_ts(tmem + cD2, tmem + cR1 + 8 * s, dBh, idesc2, 1u);
                    }
                    tc_commit(bar1);
                }
                const float sgp = Gp * iqp, sgm = Gm * iqm;
                const float gs = sgp + sgm, gd = sgp - sgm;
                mbar_wait(bar1, ph1);
                ph1 ^= 1u;
                tc_fence_after();

                // ---- 6. epilogue: per-thread gradient sums over this thread's dimensions -------------------------
                uint32_t v[LW];
                TmLd<LW>::go(trow + cD2 + hsel * DH, v);
#pragma unroll
                for (int i = 0; i < DH; ++i) {
                    be[i] = fmaf(e[i], fmaf(e[i], gs, __uint_as_float(v[i])), be[i]);
                    ae[i] = fmaf(e[i], gd, ae[i]);
                }
            }
        }

        // ---- segment record -----------------------------------------------------------------------------------
        // warp (quad, hsel) contributes: hacc (hsel = 0), A/Be of its dimensions, racc of its components
        double *myrec = sRec + wid * part_stride;
        const double hs = warp_sum(hacc);
        if (lane == 0) myrec[0] = hs;
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int i = 0; i < DH; ++i) {
                const float a = warp_sum_f(ae[i]), b = warp_sum_f(be[i]);
                if (lane == 0) myrec[1 + hsel * DH + i] = (double)a, myrec[1 + DP + hsel * DH + i] = (double)b;
            }
            for (int ci = c_begin; ci < c_end; ++ci) {
                uint32_t ra[16];
                tm_ld16(trow + cR2 + 16 * ci, ra);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float r = warp_sum_f(__uint_as_float(ra[i]));
                    if (lane == 0 && 16 * ci + i < K) myrec[1 + 2 * DP + 16 * ci + i] = (double)r;
                }
            }
            tc_fence_before();
        }
        __syncthreads();
        const int nf = ANYGRAD ? part_stride : 1;
        for (int f = tid; f < nf; f += kThreads) {
            // which half of the CTA owns field f
            int h = 0;
            if (f >= 1 && f < 1 + 2 * DP) h = ((f - 1) % DP) >= DH ? 1 : 0;
            if (f >= 1 + 2 * DP) h = ((f - 1 - 2 * DP) >> 4) >= (nch + 1) / 2 ? 1 : 0;
            const double *r4 = sRec + (size_t)(4 * h) * part_stride + f;
            sTot[f] = (r4[0] + r4[part_stride]) + (r4[2 * part_stride] + r4[3 * part_stride]);
        }
        __syncthreads();
        double *rec = part + ((size_t)blockIdx.x * maxseg + seg) * (size_t)part_stride;
        for (int f = tid; f < nf; f += kThreads) {
            double v = sTot[f];
            if (ANYGRAD && f >= 1 && f < 1 + DP) {  // A_d += sum_k Delta_kd (w_k / sigma_k^2) racc_k
                const int d = f - 1;
                double s = 0.0;
                for (int k = 0; k < K; ++k) s += (double)sDl[k * DP + d] * (double)sWis[k] * sTot[1 + 2 * DP + k];
                v += s;
            }
            rec[f] = v;
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (wid == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

uint32_t tc_tmem_cols(int DP, int K) {
    const int D8 = (DP + 7) / 8 * 8, N2 = D8 <= 16 ? 16 : 32, KP = (K + 15) / 16 * 16;
    const int need = 3 * KP + N2;
    uint32_t c = 32;
    while ((int)c < need) c <<= 1;
    return c;
}

template <int DP, bool ANYGRAD, bool PHILOX>
int tc_launch_inst(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, const double *d_eps,
                   double *d_part) {
    auto kern = entmc_kernel_tc<DP, ANYGRAD, PHILOX>;
    static size_t smem_set = 0;
    if (plan.smem > smem_set) {
        VBMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        smem_set = plan.smem;
    }
    static int desc_swap = -1;
    if (desc_swap < 0) {
        const char *e = getenv("VBMC_TC_DESC_SWAP");
        desc_swap = e ? atoi(e) : 0;
    }
    kern<<<plan.grid, kThreads, plan.smem, c->stream>>>(d_params, lay, plan.half, plan.pair0, plan.half_glob, plan.chunk,
                                                     plan.maxseg, d_eps, d_part, entpart_stride(DP, lay.K),
                                                     c->entmc_guard, tc_tmem_cols(DP, lay.K), desc_swap);
    return VBMC_OK;
}

template <int DP>
int tc_launch_dp(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, bool anygrad, bool philox,
                 const double *d_eps, double *d_part) {
    if (anygrad) {
        if (philox) return tc_launch_inst<DP, true, true>(c, d_params, lay, plan, d_eps, d_part);
        return tc_launch_inst<DP, true, false>(c, d_params, lay, plan, d_eps, d_part);
    }
    if (philox) return tc_launch_inst<DP, false, true>(c, d_params, lay, plan, d_eps, d_part);
    return tc_launch_inst<DP, false, false>(c, d_params, lay, plan, d_eps, d_part);
}

}  // namespace

bool entmc_tc_supported(int DP, int K) { return DP > 0 && K >= 1 && K <= 160 && tc_tmem_cols(DP, K) <= 512; }

// chunk (multiple of 128 pairs), grid and shared memory of the tensor-core kernel
int entmc_tc_plan(const Ctx *c, int D, int K, int64_t half_local, EntmcPlan *plan) {
    const int DP = pad_dim(D);
    const uint32_t cols = tc_tmem_cols(DP, K);
    const int per_sm = (int)(512 / cols);  // tensor memory: 512 columns per SM
    size_t smem = tc_smem_layout(DP, D, K, entpart_stride(DP, K)).total;
    VBMC_REQUIRE(smem <= 227 * 1024, VBMC_ERR_UNSUPPORTED, "entmc (tensor-core): tables do not fit in shared memory");
    // never let more CTAs become resident than tensor memory can serve (tcgen05.alloc would spin): pad the
    // shared-memory request so that exactly per_sm CTAs fit
    const size_t floor_smem = (size_t)(227 * 1024) / (per_sm + 1) + 1024;
    if (smem < floor_smem) smem = floor_smem;
    const int64_t T = (int64_t)K * half_local;
    const int64_t slots = (int64_t)c->sm_count * per_sm;
    int64_t chunk = (T + slots - 1) / slots;
    chunk = ((chunk + kTile - 1) / kTile) * kTile;
    if (chunk < kTile) chunk = kTile;
    plan->variant = ENTMC_TC;
    plan->threads = kThreads;
    plan->pairs_per_thread = (int)(chunk / kTile);
    plan->chunk = chunk;
    plan->grid = (int)std::max<int64_t>(1, (T + chunk - 1) / chunk);
    plan->maxseg = half_local > 0 ? (int)((chunk - 1) / half_local) + 2 : 1;
    plan->slabs = plan->grid * plan->maxseg;
    plan->half = half_local;
    plan->pair0 = 0;
    plan->half_glob = half_local;
    plan->smem = smem;
    return VBMC_OK;
}

int entmc_tc_launch(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, bool anygrad, bool philox,
                    const double *d_eps, double *d_part) {
    switch (lay.DP) {
#define VBMC_CASE(N) \
    case N:          \
        return tc_launch_dp<N>(c, d_params, lay, plan, anygrad, philox, d_eps, d_part)
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

}  // namespace vbmc

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
// X = E (h2 Delta)^T  ([pairs x D] x [D x K])  and  V = C' (wis Delta)  ([pairs x K] x [K x D]) are GEMMs.
//
// Two kernels per evaluation:
//   entmc_tc_gen_kernel   (a) CTAs [0, K): every j-dependent table of component j in its final shared-memory image
//                         (GEMM operand tiles split hi/lo for 3xTF32, per-component constants, guard mask);
//                         (b) one CTA per tile of 128 antithetic pairs: the noise tile (Philox + Box-Muller, or the
//                         eps input of parity mode) scaled by sigma_j, split hi/lo, in the K-major operand layout
//                         GEMM1 reads, plus |e|^2 per pair.  Both go to global memory and stay L2 resident.
//   entmc_kernel_tc       a CTA owns a chunk of the flattened (component, pair) space, tile by tile (2 threads per
//                         pair; thread <-> TMEM lane):
//     tile    cp.async.bulk (TMA, 1-D) of the prepared 25 KB image into a 3-deep shared-memory ring
//     GEMM1   X[128 x KP]  = E[128 x D8] (h2 Delta)^T   tcgen05.mma kind::tf32, operands from shared memory
//                                                       (K-major, no swizzle), 3xTF32 (hi*hi + hi*lo + lo*hi)
//     pass 1  tcgen05.ld X rows -> u(+-) (2 MUFU.EX2 per (pair, k)), q(+-), G(+-); u(+-) parked in TMEM
//     pass 2  u(+-) -> racc_k (registers), c'_k = u+/q+ - u-/q- split hi/lo -> TMEM
//     GEMM2   V[128 x N2]  = C'[128 x KP] (wis Delta)   A operand straight from TMEM, 3xTF32 (2xTF32 was tried: its
//                                                       2^-12 per-term error breaks the 1e-4 gradient bound at tiny N_s)
//     epilogue per-thread sums of e_d (v_d + e_d (G+/q+ + G-/q-)) and e_d (G+/q+ - G-/q-)
//   X is double-buffered in TMEM; there is NO CTA-wide barrier in the tile loop: the two threads of a pair meet on
//   a 64-thread named barrier, everything else is mbarriers (tile landed / X ready / V ready / warps done), so
//   warps drift apart and cover each other's TMEM and MUFU latencies; two CTAs per SM.
//   ~19 CUDA-core instructions per (pair, component) instead of ~50 in the CUDA-core kernels of entmc.cu.
// Components whose expanded distance is badly conditioned (same guard as entmc_kernel_w) have their u(+-)
// recomputed with direct differences on the CUDA cores.
// Work distribution, records and determinism are those of entmc_kernel_w: the K * half pairs are one index space
// cut into equal chunks (one per CTA, a multiple of 128 pairs), a chunk that straddles components is processed
// segment by segment, ONE fp64 record [hacc | A_d | Be_d | racc_k] per (CTA, segment), no atomics.
#include "common.cuh"
#include "philox.cuh"

namespace vbmc {
namespace {

struct alignas(16) KTc {
    float ck2, hd, w, wis;  // s0_k = hd E + ck2 ;  log2 u(+-) = s0_k -+ X_k ;  wis = w / sigma^2
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
// round-to-nearest tf32 (unbiased, unlike masking the low mantissa bits)
__device__ __forceinline__ uint32_t f32_to_tf32_rna(float x) {
    uint32_t y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
    return y;
}

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void pair_barrier(int quad) {  // the two warps that share a TMEM lane quadrant
    switch (quad) {  // immediate barrier ids: a register operand would reserve all 16 hardware barriers
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 2000000000LL) __trap();  // ~1 s: a lost completion must not hang the GPU
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
// two 16-column loads in flight, one wait
__device__ __forceinline__ void tm_ld16x2(uint32_t ta, uint32_t (&a)[16], uint32_t tb, uint32_t (&b)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : TM_R16(a), TM_R16(b)
        : "r"(ta), "r"(tb)
        : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), TM_W16(v)
        : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// sums over the 32 lanes of 32 per-lane values at once: afterwards lane l holds sum_lanes v[l] in v[0]
// (31 shuffles instead of 32 x 5: at every step a lane keeps one half of its values and trades the other)
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16, n = 16; s >= 1; s >>= 1, n >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = upper ? v[i] : v[i + n];
            const float keep = upper ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

constexpr int kTile = 128;     // antithetic pairs per tile = TMEM lanes
constexpr int kThreads = 256;  // two threads per pair: each owns half of the dimensions and half of the components
constexpr int kMaxChunks = 2;  // 16-column component chunks per thread (K <= 64): racc lives in registers
constexpr int kRing = 3;       // noise-tile buffers in shared memory

__host__ __device__ inline int tc_d8(int DP) { return (DP + 7) / 8 * 8; }
__host__ __device__ inline int tc_n2(int DP) { return tc_d8(DP) <= 16 ? 16 : 32; }
__host__ __device__ inline int tc_kp(int K) { return (K + 15) / 16 * 16; }

// Per-component table block (byte offsets): the image entmc_tc_gen_kernel writes to global memory and the main
// kernel copies verbatim into shared memory at a segment start.  Dl (plain Delta, fp32) is only read by the rare
// direct-difference path and stays in global memory (it sits behind `smem_bytes`).
struct TcTab {
    uint32_t B1h, B1l, B2h, B2l, Kc, Dir, Mask, Scal, smem_bytes, Dl, total;
};
__host__ __device__ inline TcTab tc_tab_layout(int DP, int K) {
    const int D8 = tc_d8(DP), NC1 = D8 / 4, N2 = tc_n2(DP), KP = tc_kp(K);
    TcTab s;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        const uint32_t at = o;
        o += (bytes + 127u) & ~127u;
        return at;
    };
    s.B1h = take(NC1 * KP * 16);
    s.B1l = take(NC1 * KP * 16);
    s.B2h = take((KP / 4) * N2 * 16);
    s.B2l = take((KP / 4) * N2 * 16);
    s.Kc = take(KP * sizeof(KTc));
    s.Dir = take(KP * sizeof(KDir));
    s.Mask = take(16 * 4);
    s.Scal = take(64);  // float sj, hj * sj^2
    s.smem_bytes = o;
    s.Dl = take(KP * DP * 4);
    s.total = o;
    return s;
}

// noise-tile image: [hi (NC1 x 128 x 16 B) | lo (same) | partial |e|^2 (2 halves x 128 floats)], K-major
// no-swizzle operand layout
__host__ __device__ inline uint32_t tc_a_bytes(int DP) { return (uint32_t)(tc_d8(DP) / 4) * kTile * 16; }  // one of hi / lo
__host__ __device__ inline uint32_t tc_tile_bytes(int DP) { return 2 * tc_a_bytes(DP) + 2 * kTile * 4; }

struct TcSmem {  // byte offsets inside the dynamic shared memory of entmc_kernel_tc
    uint32_t Tab, A, Q, Rec, Tot, Bar, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int DP, int K, int part_stride) {
    TcSmem s;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        const uint32_t at = o;
        o += (bytes + 127u) & ~127u;
        return at;
    };
    s.Tab = take(tc_tab_layout(DP, K).smem_bytes);
    s.A = take(kRing * tc_tile_bytes(DP));
    s.Q = take(2 * 2 * 4 * kTile * 4);  // [tile parity][half][q+, q-, G+, G-][row]
    s.Bar = take(128);
    s.total = o;
    // the record scratch is only used between segments, when the ring is idle: it aliases the first slots
    s.Rec = s.A;
    s.Tot = s.A + (((uint32_t)(8 * part_stride * 8) + 127u) & ~127u);
    return s;
}

// tiles of one CTA's chunk, in processing order: segment after segment, each segment's pairs in tiles of 128.
// Both kernels enumerate them with this walk; `tpc` images are reserved per CTA.
struct TcWork {
    int64_t half, pair0, half_glob, chunk;
    int K, tpc, n_img;
    ChunkMap cm;
};
#ifndef VBMC_GEN_TILES
#define VBMC_GEN_TILES 4
#endif
#ifndef VBMC_GEN_ROT
#define VBMC_GEN_ROT 1  // images dealt to the generator CTAs round-robin with a rotation: every CTA gets live and dead images alike (15.1 -> 13.6 us)
#endif
constexpr int kGenTiles = VBMC_GEN_TILES;  // noise-tile images per generator CTA

// ------------------------------------------------------------------------------------------------------------
// CTAs [0, K): tables of component j = blockIdx.x.  CTAs K + g: noise-tile image g.
// dims [D0, D0 + DH) of one row -> operand layout in global memory, in 16-byte pieces where the alignment allows
template <int DH, int D0>
__device__ __forceinline__ void store_half(const float (&eh)[DH], const float (&el)[DH], float *aH, float *aL, int row) {
    static_assert(DH % 2 == 0 && D0 % 2 == 0, "dimension halves are stored in 8- or 16-byte pieces");
#pragma unroll
    for (int i = 0; i < DH; i += 2) {
        const int d = D0 + i;
        const int o = (d >> 2) * (kTile * 4) + row * 4 + (d & 3);
        if ((d & 3) == 0 && i + 4 <= DH) {
            *reinterpret_cast<float4 *>(aH + o) = make_float4(eh[i], eh[i + 1], eh[i + 2 < DH ? i + 2 : i], eh[i + 3 < DH ? i + 3 : i]);
            *reinterpret_cast<float4 *>(aL + o) = make_float4(el[i], el[i + 1], el[i + 2 < DH ? i + 2 : i], el[i + 3 < DH ? i + 3 : i]);
        } else if ((d & 3) == 2 && i >= 2 && i + 2 <= DH) {
            // second half of the 16-byte piece stored in the previous step
        } else {
            *reinterpret_cast<float2 *>(aH + o) = make_float2(eh[i], eh[i + 1]);
            *reinterpret_cast<float2 *>(aL + o) = make_float2(el[i], el[i + 1]);
        }
    }
}

template <int DP, bool PHILOX>
__global__ void __launch_bounds__(kThreads)
entmc_tc_gen_kernel(const double *__restrict__ prm, ParamLayout lay, float guard, unsigned char *__restrict__ tab,
                    TcWork wk, const double *__restrict__ eps, unsigned char *__restrict__ tiles, int n_tab,
                    int64_t key_delta, int key_by_value, uint64_t seed_v, uint64_t offset_v) {
    const int D = lay.D, K = lay.K, tid = threadIdx.x, nt = blockDim.x;
    constexpr int DH = DP / 2, D8 = (DP + 7) / 8 * 8, N2 = D8 <= 16 ? 16 : 32;
    extern __shared__ __align__(16) unsigned char psm[];
    // programmatic dependent launch: the main kernel may become resident (barrier init, TMEM allocation) while this
    // grid is still running; it reads nothing this grid writes before its own griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;");
    // CTAs [0, n_tab): tables of component blockIdx.x (n_tab = K, or 0 when only noise is wanted: the look-ahead launch
    // that fills the OTHER tile buffer for the next evaluation while this one's tail runs).  The noise tiles are
    // UNSCALED standard normals: they depend on the Philox key only, never on theta.
    if ((int)blockIdx.x >= n_tab) {
        // ---- (b) noise tile -------------------------------------------------------------------------------
        __shared__ int s_info[kGenTiles][4];  // j, n, t0, valid
        __shared__ int64_t s_plo[kGenTiles];
        __shared__ uint64_t s_key[2];
        __shared__ int s_g[kGenTiles];
        const int n_gen = (int)gridDim.x - n_tab, gb = (int)blockIdx.x - n_tab;
        if (PHILOX && tid == kGenTiles) {
            // the key of THIS evaluation rides behind the parameter block (key_delta = 1: the next evaluation's draws); a
            // prefetch launched before the parameters exist on the device passes the key by value.  One load per CTA.
            uint64_t ks = seed_v, ko = offset_v;
            if (!key_by_value) {
                const uint64_t *rngp = reinterpret_cast<const uint64_t *>(prm + lay.total());
                ks = rngp[0], ko = rngp[1] + (uint64_t)key_delta;
            }
            s_key[0] = ks, s_key[1] = ko;
        }
        if (tid < kGenTiles) {  // which (component, pair range) is image g?  (64-bit divisions: once per image)
            const int g = VBMC_GEN_ROT ? tid * n_gen + (gb + 3 * tid) % n_gen : gb * kGenTiles + tid;
            s_g[tid] = g;
            const int cta = g / wk.tpc, l = g - cta * wk.tpc;
            const int64_t Tn = (int64_t)K * wk.half;
            const int64_t g0 = wk.cm.start(cta), g1 = min((int64_t)wk.cm.start(cta + 1), Tn);
            int valid = 0, j = 0, n = 0, t0 = 0;
            int64_t p_lo = 0;
            if (g0 < Tn) {
                int cnt = 0;
                for (j = (int)(g0 / wk.half);; ++j) {
                    const int64_t lo = max(g0, (int64_t)j * wk.half), hi = min(g1, (int64_t)(j + 1) * wk.half);
                    if (j >= K || lo >= hi) break;  // image not used by this chunk
                    n = (int)(hi - lo);
                    const int ntile = (n + kTile - 1) / kTile;
                    if (l < cnt + ntile) {
                        t0 = (l - cnt) * kTile;
                        p_lo = lo - (int64_t)j * wk.half;
                        valid = 1;
                        break;
                    }
                    cnt += ntile;
                }
            }
            s_info[tid][0] = j, s_info[tid][1] = n, s_info[tid][2] = t0, s_info[tid][3] = valid && g < wk.n_img;
            s_plo[tid] = p_lo;
        }
        __syncthreads();
        for (int gi = 0; gi < kGenTiles; ++gi) {
        if (!s_info[gi][3]) continue;
        const int g = s_g[gi];
        const int j = s_info[gi][0], n = s_info[gi][1], t0 = s_info[gi][2];
        const int64_t p_lo = s_plo[gi];
        constexpr uint32_t ABYTES = (uint32_t)(D8 / 4) * kTile * 16, TB = 2 * ABYTES + 2 * kTile * 4;
        unsigned char *img = tiles + (size_t)g * TB;
        float *aH = reinterpret_cast<float *>(img), *aL = reinterpret_cast<float *>(img + ABYTES);
        float *gE = reinterpret_cast<float *>(img + 2 * ABYTES);
        const int row = tid & (kTile - 1), hsel = tid >> 7;
        const int off = t0 + row;
        const bool live = off < n;
        const int64_t gpair = wk.pair0 + p_lo + (live ? off : 0);
        float z[DH];
        if (PHILOX) {
            philox_normals_half<DH>(s_key[0], s_key[1], (uint32_t)j, (uint64_t)gpair, hsel, D, z);
        } else {
            const double *ep = eps + ((size_t)j * (size_t)wk.half_glob + (size_t)gpair) * (size_t)D;
#pragma unroll
            for (int i = 0; i < DH; ++i) z[i] = (hsel * DH + i < D) ? (float)__ldg(ep + hsel * DH + i) : 0.0f;
        }
        // straight to global memory in the operand layout: element (row, d) sits at (d / 4) * 2048 + row * 16 + (d % 4) * 4
        // bytes, so the lanes of a warp write 32 consecutive 16-byte (or 8-byte: DH % 4 == 2) pieces per store
        float Eh = 0.f, eh[DH], el[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i) {
            const float e = live ? z[i] : 0.0f;  // UNSCALED: sigma_j is folded into the tables, the tile depends on the key only
            Eh = fmaf(e, e, Eh);
            eh[i] = __uint_as_float(__float_as_uint(e) & 0xffffe000u);
            el[i] = e - eh[i];
        }
        gE[hsel * kTile + row] = Eh;
        if (hsel == 0)
            store_half<DH, 0>(eh, el, aH, aL, row);
        else
            store_half<DH, DH>(eh, el, aH, aL, row);
        if constexpr (D8 > DP) {  // padded dimensions of the operand tile are zero (DP % 4 == 0: whole 16-byte pieces)
            if (hsel == 1) {
#pragma unroll
                for (int d = DP; d < D8; d += 4) {
                    const int o = (d >> 2) * (kTile * 4) + row * 4;
                    *reinterpret_cast<float4 *>(aH + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4 *>(aL + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        }  // images of this CTA
        return;
    }
    // ---- (a) tables of component j ------------------------------------------------------------------------
    const int j = blockIdx.x, KP = tc_kp(K);
    const TcTab L = tc_tab_layout(DP, K);
    unsigned char *base = tab + (size_t)j * L.total;
    float *gB1h = reinterpret_cast<float *>(base + L.B1h), *gB1l = reinterpret_cast<float *>(base + L.B1l);
    float *gB2h = reinterpret_cast<float *>(base + L.B2h), *gB2l = reinterpret_cast<float *>(base + L.B2l);
    KTc *gKc = reinterpret_cast<KTc *>(base + L.Kc);
    KDir *gDir = reinterpret_cast<KDir *>(base + L.Dir);
    uint32_t *gMask = reinterpret_cast<uint32_t *>(base + L.Mask);
    float *gDl = reinterpret_cast<float *>(base + L.Dl);
    float *sDl = reinterpret_cast<float *>(psm);              // [KP][DP] Delta (fp32-rounded)
    double *sH2 = reinterpret_cast<double *>(sDl + KP * DP);  // [KP] 2 h_k (0: padded / guarded component)
    double *sWis = sH2 + KP;                                  // [KP] w_k / sigma_k^2
    __shared__ uint32_t sMask[16];

    const double *mu = prm + lay.mu(), *sigma = prm + lay.sigma(), *lambd = prm + lay.lambd(), *w = prm + lay.w();
    const double kHalfLog2e = 0.72134752044448170368;  // log2(e) / 2
    const double sig_j = sigma[j];
    const double hjd = kHalfLog2e / (sig_j * sig_j);
    const double Emax = sig_j * sig_j * (D + 8.0 * sqrt(2.0 * D) + 32.0);
    // (latency matters here: these K small CTAs sit in front of the main kernel.  One reciprocal per dimension instead
    // of one per table entry, and the per-component constants -- two fp64 logarithms and three divisions each -- are
    // computed by KP threads side by side instead of by lane 0 of eight warps, component after component.)
    __shared__ double sInvL[kMaxD];
    if (tid < 16) sMask[tid] = 0u;
    if (tid < DP) sInvL[tid] = tid < D ? 1.0 / lambd[tid] : 0.0;
    __syncthreads();
    for (int i = tid; i < KP * DP; i += nt) {
        const int k = i / DP, d = i - k * DP;
        const float v = (k < K && d < D) ? (float)((mu[j * D + d] - mu[k * D + d]) * sInvL[d]) : 0.0f;
        sDl[i] = v;
        gDl[i] = v;
    }
    __syncthreads();
    if (tid < KP) {
        const int k = tid;
        KTc c;
        KDir dr;
        c.ck2 = -200.0f, c.hd = 0.f, c.w = 0.f, c.wis = 0.f;  // padding / guarded: expanded-form u = 2^-200 = 0
        dr.ck = -200.0f, dr.h = 0.f;
        double h2 = 0.0, wis = 0.0;
        if (k < K) {
            double A0 = 0.0, A1 = 0.0;  // |Delta_k|^2 of the rounded table entries
#pragma unroll
            for (int d = 0; d < DP; d += 2) {
                const double a = (double)sDl[k * DP + d], b2 = (double)sDl[k * DP + d + 1];
                A0 = fma(a, a, A0), A1 = fma(b2, b2, A1);
            }
            const double A = A0 + A1;
            const double sk = sigma[k];
            const double hk = kHalfLog2e / (sk * sk);
            const double ck = D * (log2(sig_j) - log2(sk));
            c.w = (float)w[k];
            wis = w[k] / (sk * sk);
            c.wis = (float)wis;
            dr.ck = (float)ck;
            dr.h = (float)hk;
            // conditioning of the expanded form (see entmc_kernel_fast): direct differences beyond the guard
            if (k != j && hk * (A + Emax) > (double)guard) {
                atomicOr(&sMask[k >> 4], 1u << (k & 15));
            } else {
                c.ck2 = (float)(ck - hk * A);
                c.hd = (float)((hjd - hk) * sig_j * sig_j);  // multiplies |z|^2 = |e|^2 / sigma_j^2
                h2 = 2.0 * hk;
            }
        }
        gKc[k] = c;
        gDir[k] = dr;
        sH2[k] = h2;
        sWis[k] = wis;
    }
    __syncthreads();
    if (tid < 16) gMask[tid] = sMask[tid];
    if (tid == 0) {
        float *sc = reinterpret_cast<float *>(base + L.Scal);
        sc[0] = (float)sig_j;
        sc[1] = (float)(hjd * sig_j * sig_j);  // direct path: log2 u = ck + (hj sigma_j^2) |z|^2 - h |t|^2
    }
    // operand tables of the two GEMMs (hi = upper 19 bits = exact tf32, lo = remainder); the per-component scales
    // are folded in: GEMM1 yields X_k = 2 h_k B_k directly, GEMM2 contracts (u+/q+ - u-/q-) with wis_k Delta_k
    for (int i = tid; i < KP * D8; i += nt) {
        const int k = i / D8, d = i - k * D8;
        const double dl = (k < K && d < DP) ? (double)sDl[k * DP + d] : 0.0;
        {
            const float v = (float)(sH2[k] * sig_j * dl);  // GEMM1 contracts the UNSCALED noise z: X_k = 2 h_k sigma_j Delta_k . z
            const float vh = __uint_as_float(__float_as_uint(v) & 0xffffe000u), vl = v - vh;
            const int o1 = (d >> 2) * (KP * 4) + k * 4 + (d & 3);  // GEMM1 B operand: rows = components, K dim = d
            gB1h[o1] = vh, gB1l[o1] = vl;
        }
        {
            const float v = (float)(sWis[k] * dl);
            const float vh = __uint_as_float(__float_as_uint(v) & 0xffffe000u), vl = v - vh;
            const int o2 = (k >> 2) * (N2 * 4) + d * 4 + (k & 3);  // GEMM2 B operand: rows = d, K dim = components
            gB2h[o2] = vh, gB2l[o2] = vl;
        }
    }
    if constexpr (N2 > D8) {
        constexpr int NZ = N2 - D8;
        for (int i = tid; i < KP * NZ; i += nt) {
            const int k = i / NZ, d = D8 + (i - k * NZ);
            const int o2 = (k >> 2) * (N2 * 4) + d * 4 + (k & 3);
            gB2h[o2] = 0.f, gB2l[o2] = 0.f;
        }
    }
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

template <int DP, bool ANYGRAD>
__global__ void __launch_bounds__(kThreads, 2)
entmc_kernel_tc(const double *__restrict__ prm, ParamLayout lay, TcWork wk, int maxseg, double *__restrict__ part,
                int part_stride, const unsigned char *__restrict__ tab, const unsigned char *__restrict__ tiles,
                uint32_t tmem_cols) {
    constexpr int DH = DP / 2;            // dimensions per thread
    constexpr int D8 = (DP + 7) / 8 * 8;  // GEMM1 reduction length (tf32 MMAs consume 8 per instruction)
    constexpr int NC1 = D8 / 4;           // 16-byte chunks along D
    constexpr int N2 = D8 <= 16 ? 16 : 32;  // GEMM2 output columns (M = 128 needs N % 16 == 0)
    constexpr int LW = DH <= 8 ? 8 : 16;    // columns of V loaded per thread in the epilogue (DH + LW <= N2)
    constexpr uint32_t ABYTES = (uint32_t)NC1 * kTile * 16, TB = 2 * ABYTES + 2 * kTile * 4;
    const int D = lay.D, K = lay.K;
    const int64_t half = wk.half;
    const int nch = (K + 15) >> 4, KP = nch * 16;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = tid & (kTile - 1), hsel = tid >> 7, quad = wid & 3;
    // components (16-column chunks) of this thread: the first half of the chunks for hsel = 0, the rest for 1
    const int c_begin = hsel ? (nch + 1) / 2 : 0, c_end = hsel ? nch : (nch + 1) / 2;

    extern __shared__ __align__(128) unsigned char smem[];
    const TcTab T = tc_tab_layout(DP, K);
    const TcSmem L = tc_smem_layout(DP, K, part_stride);
    unsigned char *sTab = smem + L.Tab;
    const float *sB2h = reinterpret_cast<const float *>(sTab + T.B2h), *sB2l = reinterpret_cast<const float *>(sTab + T.B2l);
    const KTc *sKc = reinterpret_cast<const KTc *>(sTab + T.Kc);
    const KDir *sDir = reinterpret_cast<const KDir *>(sTab + T.Dir);
    const uint32_t *sMask = reinterpret_cast<const uint32_t *>(sTab + T.Mask);
    const float *sScal = reinterpret_cast<const float *>(sTab + T.Scal);
    unsigned char *sA = smem + L.A;                     // ring of tile images [hi | lo | E]
    float *sQ0 = reinterpret_cast<float *>(smem + L.Q);  // [tile parity][half][4][128] partial q+, q-, G+, G-
    double *sRec = reinterpret_cast<double *>(smem + L.Rec);
    double *sTot = reinterpret_cast<double *>(smem + L.Tot);
    uint64_t *sBar = reinterpret_cast<uint64_t *>(smem + L.Bar);
    uint32_t *sTmem = reinterpret_cast<uint32_t *>(sBar + 8);  // (7 barriers above)

    const int64_t Tn = (int64_t)K * half;
    const int64_t g0 = wk.cm.start((int)blockIdx.x), g1 = min((int64_t)wk.cm.start((int)blockIdx.x + 1), Tn);
    if (g0 >= Tn) return;

    // ---- one-time set-up: barriers, tensor memory ---------------------------------------------------------------
    // sBar: [0..2] tile landed (tx) | [3] V ready (GEMM2 commit) | [4] warps done (8) | [5, 6] X0 / X1 ready (GEMM1 commit)
    const uint32_t barF = smem_u32(sBar), bar1 = smem_u32(sBar + 3), barC = smem_u32(sBar + 4), barX = smem_u32(sBar + 5);
    if (tid == 0) {
        for (int i = 0; i < kRing; ++i) mbar_init(barF + 8 * i, 1);
        mbar_init(barX, 1);
        mbar_init(barX + 8, 1);
        mbar_init(bar1, 1);
        mbar_init(barC, kThreads / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);  // this warp's lane quadrant
    // everything above overlapped the table (+ inline noise) kernel in front of this one; its outputs are read below
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // TMEM columns: X0 / X1 = GEMM1 output of even / odd tiles -> u+ -> c_hi (in place), U = u- -> c_lo, V = GEMM2 output
    const uint32_t cU = 2 * KP, cV = 3 * KP;
    uint32_t ph1 = 0, phC = 0;

    const uint32_t idesc1 = umma_idesc_tf32(KP), idesc2 = umma_idesc_tf32(N2);
    // operand strides: rows 16 B apart, 8-row groups 128 B apart, K chunks one whole row-block apart
    const uint32_t lboA = kTile * 16, lboB1 = KP * 16, lboB2 = N2 * 16, sbo = 128;
    const unsigned char *my_tiles = tiles + (size_t)blockIdx.x * wk.tpc * TB;

    // (elected thread) tile image r of this CTA -> ring slot r % kRing
    auto issue_load = [&](int r) {
        const uint32_t bar = barF + 8 * (r % kRing);
        mbar_expect_tx(bar, TB);
        bulk_g2s(smem_u32(sA) + (uint32_t)(r % kRing) * TB, my_tiles + (size_t)r * TB, TB, bar);
    };
    // (elected thread) X[r & 1] = E[r] (h2 Delta)^T, 3xTF32
    auto issue_gemm1 = [&](int r) {
        const uint32_t aH = smem_u32(sA) + (uint32_t)(r % kRing) * TB, aL = aH + ABYTES;
        const uint32_t bH = smem_u32(sTab + T.B1h), bL = smem_u32(sTab + T.B1l);
        const uint32_t dcol = tmem + (uint32_t)(r & 1) * KP;
#pragma unroll
        for (int s = 0; s < D8 / 8; ++s) {
            const uint64_t dAh = umma_desc(aH + s * 2 * lboA, lboA, sbo), dAl = umma_desc(aL + s * 2 * lboA, lboA, sbo);
            const uint64_t dBh = umma_desc(bH + s * 2 * lboB1, lboB1, sbo), dBl = umma_desc(bL + s * 2 * lboB1, lboB1, sbo);
            mma_ss(dcol, dAh, dBh, idesc1, s > 0 ? 1u : 0u);
            mma_ss(dcol, dAh, dBl, idesc1, 1u);
            mma_ss(dcol, dAl, dBh, idesc1, 1u);
        }
        tc_commit(barX + 8 * (uint32_t)(r & 1));
    };

    int r0 = 0;  // tiles of this CTA processed so far (ring slots and barrier parities run on across segments)
    const int j_first = (int)(g0 / half);
    for (int seg = 0;; ++seg) {
        const int j = j_first + seg;
        const int64_t lo = max(g0, (int64_t)j * half), hi = min(g1, (int64_t)(j + 1) * half);
        if (j >= K || lo >= hi) break;
        const int n = (int)(hi - lo);
        const int ntile = (n + kTile - 1) / kTile;

        // ---- component tables: verbatim copy of the prepared image; first two noise tiles in flight ----------------
        __syncthreads();
        if (tid == 0) {
            issue_load(r0);
            if (ntile > 1) issue_load(r0 + 1);
        }
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(tab + (size_t)j * T.total);
            uint4 *dst = reinterpret_cast<uint4 *>(sTab);
            for (int i = tid; i < (int)(T.smem_bytes / 16); i += kThreads) dst[i] = __ldg(src + i);
        }
        fence_async_smem();  // generic-proxy writes (tables) -> async proxy (MMA operand reads)
        __syncthreads();
        const float sj = sScal[0], hj = sScal[1];
        const float *gDl = reinterpret_cast<const float *>(tab + (size_t)j * T.total + T.Dl);
        if (tid == kTile) {
            mbar_wait(barF + 8 * (r0 % kRing), (uint32_t)(r0 / kRing) & 1u);
            tc_fence_after();
            issue_gemm1(r0);
        }

        double hacc = 0.0;
        float racc[ANYGRAD ? kMaxChunks * 16 : 1];
        float ae[ANYGRAD ? DH : 1], be[ANYGRAD ? DH : 1];
        if constexpr (ANYGRAD) {
#pragma unroll
            for (int i = 0; i < kMaxChunks * 16; ++i) racc[i] = 0.f;
#pragma unroll
            for (int i = 0; i < DH; ++i) ae[i] = be[i] = 0.f;
        }
        float gs = 0.f, gd = 0.f;  // sigma_j^2 (G+/q+ + G-/q-) and sigma_j (G+/q+ - G-/q-) of the tile whose epilogue is pending

        // epilogue of tile r (its GEMM2 has completed): per-thread gradient sums over this thread's dimensions
        auto epilogue = [&](int r) {
            const unsigned char *img = sA + (size_t)(r % kRing) * TB;
            const float *aH = reinterpret_cast<const float *>(img), *aL = reinterpret_cast<const float *>(img + ABYTES);
            uint32_t v[LW];
            TmLd<LW>::go(trow + cV + hsel * DH, v);
#pragma unroll
            for (int i = 0; i < DH; ++i) {
                const int d = hsel * DH + i;
                const int o = (d >> 2) * (kTile * 4) + row * 4 + (d & 3);
                const float z = aH[o] + aL[o];  // exact: hi + lo is the fp32 noise value; e = sigma_j z
                be[i] = fmaf(z, fmaf(z, gs, sj * __uint_as_float(v[i])), be[i]);
                ae[i] = fmaf(z, gd, ae[i]);
            }
        };

        // Tile loop.  MMA latencies never sit on the compute warps' critical path: GEMM1(t + 1) is issued in the
        // middle of tile t (its X buffer and noise tile are ready then), GEMM2(t) runs under the first half of pass 1
        // of tile t + 1, and the epilogue of tile t is deferred into tile t + 1.
        for (int t = 0; t < ntile; ++t) {
            const int r = r0 + t;
            const bool live = t * kTile + row < n;
            const uint32_t cX = (uint32_t)(r & 1) * KP;
            const unsigned char *img = sA + (size_t)(r % kRing) * TB;
            const float *aH = reinterpret_cast<const float *>(img), *aL = reinterpret_cast<const float *>(img + ABYTES);
            mbar_wait(barF + 8 * (r % kRing), (uint32_t)(r / kRing) & 1u);  // (long since complete: makes the image visible)
            const float E = reinterpret_cast<const float *>(img + 2 * ABYTES)[row] + reinterpret_cast<const float *>(img + 2 * ABYTES)[kTile + row];
            mbar_wait(barX + 8 * (uint32_t)(r & 1), (uint32_t)(r >> 1) & 1u);  // X(t): issued one tile ago
            tc_fence_after();

            // ---- pass 1: density ratios u(+-) of this thread's components, partial mixture sums ----------------
            float qp = 0.f, qm = 0.f, Gp = 0.f, Gm = 0.f;
            uint32_t um_keep[2][16];
#pragma unroll
            for (int cc = 0; cc < kMaxChunks; ++cc) {
                const int ci = c_begin + cc;
                uint32_t (&um)[16] = um_keep[cc & 1];
                if (ci < c_end) {
                    uint32_t x[16];
                    tm_ld16(trow + cX + 16 * ci, x);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const KTc c = sKc[16 * ci + i];
                        const float s0 = fmaf(c.hd, E, c.ck2);
                        const float xx = __uint_as_float(x[i]);
                        const float vp = ex2f(s0 - xx), vm = ex2f(s0 + xx);
                        qp = fmaf(c.w, vp, qp), qm = fmaf(c.w, vm, qm);
                        if constexpr (ANYGRAD) Gp = fmaf(c.wis, vp, Gp), Gm = fmaf(c.wis, vm, Gm);
                        x[i] = __float_as_uint(vp), um[i] = __float_as_uint(vm);
                    }
                    const uint32_t fm = sMask[ci];
                    if (fm != 0u) {  // CTA-uniform, rare: badly conditioned components, direct differences
                        const float base = hj * E;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if ((fm >> i) & 1u) {
                                const int k = 16 * ci + i;
                                const KDir dr = sDir[k];
                                const float *dl = gDl + k * DP;
                                float a0 = 0.f, a1 = 0.f;
#pragma unroll 1
                                for (int d = 0; d < DP; ++d) {
                                    const int o = (d >> 2) * (kTile * 4) + row * 4 + (d & 3);
                                    const float ed = sj * (aH[o] + aL[o]), dd = __ldg(dl + d);
                                    const float tp = dd + ed, tm = dd - ed;
                                    a0 = fmaf(tp, tp, a0), a1 = fmaf(tm, tm, a1);
                                }
                                const float vp = ex2f(fmaf(-dr.h, a0, dr.ck + base));
                                const float vm = ex2f(fmaf(-dr.h, a1, dr.ck + base));
                                // (the expanded-form value of a guarded component above is exactly 0: ck2 = -200)
                                const float wk_ = sKc[k].w, wis = sKc[k].wis;
                                qp = fmaf(wk_, vp, qp), qm = fmaf(wk_, vm, qm);
                                if constexpr (ANYGRAD) Gp = fmaf(wis, vp, Gp), Gm = fmaf(wis, vm, Gm);
                                x[i] = __float_as_uint(vp), um[i] = __float_as_uint(vm);
                            }
                        }
                    }
                    if constexpr (ANYGRAD) tm_st16(trow + cX + 16 * ci, x);
                }
                if constexpr (ANYGRAD) {
                    if (cc == 0 && t > 0) {
                        // U still feeds GEMM2 of the previous tile (c_lo): it ran under the arithmetic above
                        mbar_wait(bar1, ph1);
                        ph1 ^= 1u;
                        tc_fence_after();
                    }
                    if (ci < c_end) tm_st16(trow + cU + 16 * ci, um_keep[cc & 1]);
                }
            }
            if constexpr (ANYGRAD) {
                if (t > 0) epilogue(r - 1);  // deferred epilogue of the previous tile (V complete: bar1 above)
            }
            if (tid == kTile && t + 1 < ntile) {
                // second elected thread: GEMM1 of the NEXT tile.  Its X buffer was c' of tile t - 1 (GEMM2(t - 1) done:
                // bar1 above; all warps past their pass 2 of t - 1: barC below, waited at the end of the previous
                // iteration), its noise tile landed long ago
                if (t > 0) {
                    mbar_wait(barC, phC ^ 1u);
                    tc_fence_after();
                }
                mbar_wait(barF + 8 * ((r + 1) % kRing), (uint32_t)((r + 1) / kRing) & 1u);
                tc_fence_after();
                issue_gemm1(r + 1);
            }
            __syncwarp();
            // exchange the partial sums of the two threads of a pair (they sit in the two warps of one quadrant)
            float *sQ = sQ0 + (r & 1) * (8 * kTile);  // (by tile parity: a warp without components may run a tile ahead)
            sQ[(hsel * 4 + 0) * kTile + row] = qp, sQ[(hsel * 4 + 1) * kTile + row] = qm;
            if constexpr (ANYGRAD) sQ[(hsel * 4 + 2) * kTile + row] = Gp, sQ[(hsel * 4 + 3) * kTile + row] = Gm;
            pair_barrier(quad);
            qp = sQ[0 * kTile + row] + sQ[4 * kTile + row];
            qm = sQ[1 * kTile + row] + sQ[5 * kTile + row];
            if (live && hsel == 0)
                hacc += 0.69314718055994530942 * ((double)log2f(qp) + (double)log2f(qm)) - (double)E;  // E = |z|^2 = |e|^2 / sigma_j^2

            if constexpr (ANYGRAD) {
                Gp = sQ[2 * kTile + row] + sQ[6 * kTile + row];
                Gm = sQ[3 * kTile + row] + sQ[7 * kTile + row];
                const float iqp = live ? __frcp_rn(qp) : 0.f, iqm = live ? __frcp_rn(qm) : 0.f;
                gs = sj * sj * fmaf(Gp, iqp, Gm * iqm), gd = sj * fmaf(Gp, iqp, -(Gm * iqm));
                tm_wait_st();  // own u(+-) stores of pass 1
                // ---- pass 2: racc_k += u+/q+ + u-/q- ;  c'_k = u+/q+ - u-/q- split hi/lo -------------------------
#pragma unroll
                for (int cc = 0; cc < kMaxChunks; ++cc) {
                    const int ci = c_begin + cc;
                    if (ci < c_end) {
                        uint32_t up[16], um[16];
                        tm_ld16x2(trow + cX + 16 * ci, up, trow + cU + 16 * ci, um);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float a = __uint_as_float(up[i]) * iqp, bq = __uint_as_float(um[i]) * iqm;
                            racc[cc * 16 + i] += a + bq;
                            const float c = a - bq;
                            const uint32_t ch = __float_as_uint(c) & 0xffffe000u;
                            up[i] = ch;
                            um[i] = __float_as_uint(c - __uint_as_float(ch));
                        }
                        tm_st16(trow + cX + 16 * ci, up);
                        tm_st16(trow + cU + 16 * ci, um);
                    }
                }
                tm_wait_st();
            }
            // ---- this warp is done with X[r & 1] / U / V (and with the ring slot of tile r - 1) --------------------
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(barC);
            if (tid == 0) {
                // elected thread: once ALL warps are there, GEMM2 of this tile and the load of tile r + 2 (its ring
                // slot was tile r - 1's, whose deferred epilogue every warp finished before arriving)
                mbar_wait(barC, phC);
                tc_fence_after();
                if constexpr (ANYGRAD) {
                    const uint32_t bH = smem_u32(sB2h), bL = smem_u32(sB2l);
                    for (int s = 0; s < KP / 8; ++s) {
                        const uint64_t dBh = umma_desc(bH + s * 2 * lboB2, lboB2, sbo), dBl = umma_desc(bL + s * 2 * lboB2, lboB2, sbo);
                        mma_ts(tmem + cV, tmem + cX + 8 * s, dBh, idesc2, s > 0 ? 1u : 0u);
                        mma_ts(tmem + cV, tmem + cX + 8 * s, dBl, idesc2, 1u);
                        mma_ts(tmem + cV, tmem + cU + 8 * s, dBh, idesc2, 1u);
                    }
                    tc_commit(bar1);
                }
                if (t + 2 < ntile) issue_load(r + 2);
            }
            phC ^= 1u;
            __syncwarp();
        }
        if constexpr (ANYGRAD) {  // drain: epilogue of the last tile
            mbar_wait(bar1, ph1);
            ph1 ^= 1u;
            tc_fence_after();
            epilogue(r0 + ntile - 1);
        }
        r0 += ntile;

        // ---- segment record -----------------------------------------------------------------------------------
        // warp (quad, hsel) contributes: hacc (hsel = 0), A/Be of its dimensions, racc of its components
        __syncthreads();  // the record scratch aliases the tile ring: every warp must be past its last epilogue
        double *myrec = sRec + wid * part_stride;
        const double hs = warp_sum(hacc);
        if (lane == 0) myrec[0] = hs;
        if constexpr (ANYGRAD) {
            {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = i < DH ? ae[i < DH ? i : 0] : (i < 2 * DH ? be[i < 2 * DH ? i - DH : 0] : 0.f);
                const float sres = warp_transpose_sum32(v, lane);
                if (lane < DH) myrec[1 + hsel * DH + lane] = (double)sres;
                else if (lane < 2 * DH) myrec[1 + DP + hsel * DH + lane - DH] = (double)sres;
            }
            static_assert(2 * (DP / 2) <= 32, "A/Be sums of one thread must fit one transposed warp reduction");
            {
                const float sres = warp_transpose_sum32(racc, lane);  // lane l: component 16 * (c_begin + l / 16) + l % 16
                const int ci = c_begin + (lane >> 4), k = 16 * ci + (lane & 15);
                if (ci < c_end && k < K) myrec[1 + 2 * DP + k] = (double)sres;
            }
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
        for (int f = tid; f < nf; f += kThreads)
            if (!(ANYGRAD && f >= 1 && f < 1 + D)) rec[f] = sTot[f];
        if constexpr (ANYGRAD) {
            // A_d += sum_k Delta_kd (w_k / sigma_k^2) racc_k : one warp per dimension, lanes over the components
            for (int d = wid; d < D; d += kThreads / 32) {
                double s = 0.0;
                for (int k = lane; k < K; k += 32) {
                    const int o2 = (k >> 2) * (N2 * 4) + d * 4 + (k & 3);
                    s += ((double)sB2h[o2] + (double)sB2l[o2]) * sTot[1 + 2 * DP + k];  // (wis Delta) is folded in the table
                }
                s = warp_sum(s);
                if (lane == 0) rec[1 + d] = sTot[1 + d] + s;
            }
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
    const int need = 3 * tc_kp(K) + tc_n2(DP);
    uint32_t c = 32;
    while ((int)c < need) c <<= 1;
    return c;
}

TcWork tc_work(const EntmcPlan &plan, int K) {
    TcWork w;
    w.half = plan.half, w.pair0 = plan.pair0, w.half_glob = plan.half_glob, w.chunk = plan.chunk;
    w.cm = ChunkMap{(long long)plan.chunk, (long long)plan.chunk_small, plan.n_big};
    w.K = K;
    w.tpc = (int)(plan.chunk / kTile) + plan.maxseg;  // sum_seg ceil(n_seg / 128) <= chunk / 128 + segments
    w.n_img = plan.grid * w.tpc;
    return w;
}

template <int DP>
int tc_launch_dp(Ctx *c, const double *d_params, ParamLayout lay, const EntmcPlan &plan, bool anygrad, bool philox,
                 const double *d_eps, double *d_part) {
    const int K = lay.K;
    const TcTab T = tc_tab_layout(DP, K);
    const TcWork wk = tc_work(plan, K);
    const size_t TB = tc_tile_bytes(DP);
    const size_t n_img = (size_t)plan.grid * wk.tpc;
    VBMC_TRY(ensure(&c->d_tctab, &c->tctab_cap, ((size_t)K * T.total + 7) / 8));
    unsigned char *d_tab = reinterpret_cast<unsigned char *>(c->d_tctab);
    const int KP = tc_kp(K);
    const size_t psm = (size_t)KP * DP * 4 + (size_t)KP * 16;
    const unsigned tile_ctas = (unsigned)((n_img + kGenTiles - 1) / kGenTiles);
    auto gen = [&](cudaStream_t st, unsigned char *tiles, int n_tab, bool with_tiles, int64_t delta) -> int {
        const unsigned grid = (unsigned)n_tab + (with_tiles ? tile_ctas : 0u);
        if (philox)
            entmc_tc_gen_kernel<DP, true><<<grid, kThreads, psm, st>>>(d_params, lay, c->entmc_guard, d_tab, wk, d_eps, tiles, n_tab, delta, 0, 0, 0);
        else
            entmc_tc_gen_kernel<DP, false><<<grid, kThreads, psm, st>>>(d_params, lay, c->entmc_guard, d_tab, wk, d_eps, tiles, n_tab, delta, 0, 0, 0);
        VBMC_CUDA_CHECK(cudaGetLastError());
        c->launches++;
        return VBMC_OK;
    };
    // shape / work split the tile images depend on (besides the Philox key)
    const uint64_t sig[6] = {(uint64_t)lay.D << 32 | (uint64_t)K, (uint64_t)plan.grid << 32 | (uint64_t)plan.maxseg,
                             (uint64_t)plan.chunk << 20 ^ (uint64_t)plan.chunk_small << 8 ^ (uint64_t)plan.n_big, (uint64_t)plan.half,
                             (uint64_t)plan.pair0, (uint64_t)plan.half_glob};
    const uint64_t want_seed = c->cur_seed, want_offset = c->cur_offset + (uint64_t)c->key_delta;
    bool have = philox && c->noise_ready && c->noise_seed == want_seed && c->noise_offset == want_offset;
    for (int i = 0; i < 6 && have; ++i) have = c->noise_sig[i] == sig[i];
    const int b = c->noise_buf;
    VBMC_TRY(ensure(&c->d_tctiles[b], &c->tctiles_cap[b], (n_img * TB + 7) / 8));
    unsigned char *d_tiles = reinterpret_cast<unsigned char *>(c->d_tctiles[b]);
    if (have) {
        // the draws of this evaluation were generated under the previous evaluation's tail (finalize() of that
        // evaluation re-joined the side stream, so the tiles are complete in stream order), or by vbmc_noise_prefetch
        // while the host was still packing the parameters: tables only
        // (a prefetch launch was joined into the main stream by the entry point: capi.cu settle_prefetch)
        VBMC_TRY(gen(c->stream, d_tiles, K, false, 0));
    } else {
        VBMC_TRY(gen(c->stream, d_tiles, K, true, c->key_delta));
    }
    c->noise_ready = false;
    if (c->root_forked) {  // tiles generated on the stream forked in front of the parameter kernel (capi.cu stage()): join
        VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_root_join, 0));
        c->root_forked = false;
    }
    static size_t smem_set[2] = {0, 0};
    if (plan.smem > smem_set[anygrad ? 1 : 0]) {
        if (anygrad)
            VBMC_CUDA_CHECK(cudaFuncSetAttribute(entmc_kernel_tc<DP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        else
            VBMC_CUDA_CHECK(cudaFuncSetAttribute(entmc_kernel_tc<DP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        smem_set[anygrad ? 1 : 0] = plan.smem;
    }
    const int ps = entpart_stride(DP, K);
    const uint32_t cols = tc_tmem_cols(DP, K);
    if (c->time_entmc && c->ev2) {
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev2, c->stream));
        c->ev2_recorded = true;
    }
    {
        static int pdl = -1;
        if (pdl < 0) {
            const char *e = getenv("VBMC_PDL");
            pdl = e ? atoi(e) : 1;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(plan.grid), cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = plan.smem;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
        const unsigned char *tab_c = d_tab, *tiles_c = d_tiles;
        if (anygrad)
            VBMC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, entmc_kernel_tc<DP, true>, d_params, lay, wk, plan.maxseg, d_part, ps, tab_c, tiles_c, cols));
        else
            VBMC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, entmc_kernel_tc<DP, false>, d_params, lay, wk, plan.maxseg, d_part, ps, tab_c, tiles_c, cols));
    }
    VBMC_CUDA_CHECK(cudaGetLastError());
    if (philox && c->lookahead) {
        // draws of the NEXT evaluation (key offset + 1) into the other buffer, on the side stream, behind this main
        // kernel: it runs while the tail (8 CTAs) leaves the machine idle.  finalize() re-joins the side stream.
        // (Measured alternatives, profiles/r4_e2e_timeline.md: released by the tail's launch-completion event, enqueued
        // behind the tail, or as the tail's programmatic dependent on the main stream -- none faster, the last one slower.)
        const int nb = 1 - b;
        VBMC_TRY(ensure(&c->d_tctiles[nb], &c->tctiles_cap[nb], (n_img * TB + 7) / 8));
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
        VBMC_CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_main, 0));
        VBMC_TRY(gen(c->stream2, reinterpret_cast<unsigned char *>(c->d_tctiles[nb]), 0, true, c->key_delta + 1));
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev_noise, c->stream2));
        c->noise_pending_join = true;
        c->noise_ready = true;
        c->noise_seed = want_seed, c->noise_offset = want_offset + 1;
        for (int i = 0; i < 6; ++i) c->noise_sig[i] = sig[i];
        c->noise_buf = nb;
    }
    return VBMC_OK;
}

// vbmc_noise_prefetch: the noise tiles of the evaluation with Philox key (seed, offset), on the side stream, before
// its parameters have reached the device (the tiles do not depend on them)
// root = true: launched by stage() of an evaluation on the stream forked in front of its parameter kernel; the key is
// read from the device copy `d_key` of the pinned host key (a one-warp copy kernel in front of the generator), so that a
// captured graph picks up the key of every replay
template <int DP>
int tc_prefetch_dp(Ctx *c, ParamLayout lay, const EntmcPlan &plan, uint64_t seed, uint64_t offset, bool root) {
    const int K = lay.K;
    const TcWork wk = tc_work(plan, K);
    const size_t TB = tc_tile_bytes(DP);
    const size_t n_img = (size_t)plan.grid * wk.tpc;
    const int b = c->noise_buf;
    VBMC_TRY(ensure(&c->d_tctiles[b], &c->tctiles_cap[b], (n_img * TB + 7) / 8));
    const unsigned tile_ctas = (unsigned)((n_img + kGenTiles - 1) / kGenTiles);
    if (root) {
        // (every generator CTA reading the host copy itself serialises ~1 us PCIe reads of one line: 1.1 ms measured)
        VBMC_TRY(stage_copy_launch(c, c->d_key, c->key_host, 2, c->stream3));
        // padded shared-memory request: 3 generator CTAs per SM instead of 4, so that the parameter kernel and the table
        // kernel (main stream) find room beside this grid instead of waiting for its single wave to drain
        static const size_t pad = (size_t)(getenv("VBMC_ROOT_SMEM") ? atoi(getenv("VBMC_ROOT_SMEM")) : 57) * 1024;
        static bool pad_set = false;
        if (!pad_set && pad > 48 * 1024) {
            VBMC_CUDA_CHECK(cudaFuncSetAttribute(entmc_tc_gen_kernel<DP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
            pad_set = true;
        }
        entmc_tc_gen_kernel<DP, true><<<tile_ctas, kThreads, pad, c->stream3>>>(c->d_key - lay.total(), lay, 0.f, nullptr, wk, nullptr,
                                                                             reinterpret_cast<unsigned char *>(c->d_tctiles[b]), 0,
                                                                             0, 0, 0, 0);
        VBMC_CUDA_CHECK(cudaGetLastError());
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev_root_join, c->stream3));
    } else {
        entmc_tc_gen_kernel<DP, true><<<tile_ctas, kThreads, 0, c->stream2>>>(nullptr, lay, 0.f, nullptr, wk, nullptr,
                                                                             reinterpret_cast<unsigned char *>(c->d_tctiles[b]), 0,
                                                                             0, 1, seed, offset);
        VBMC_CUDA_CHECK(cudaGetLastError());
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev_noise, c->stream2));
    }
    c->launches++;
    const uint64_t sig[6] = {(uint64_t)lay.D << 32 | (uint64_t)K, (uint64_t)plan.grid << 32 | (uint64_t)plan.maxseg,
                             (uint64_t)plan.chunk << 20 ^ (uint64_t)plan.chunk_small << 8 ^ (uint64_t)plan.n_big, (uint64_t)plan.half,
                             (uint64_t)plan.pair0, (uint64_t)plan.half_glob};
    for (int i = 0; i < 6; ++i) c->noise_sig[i] = sig[i];
    c->noise_ready = true;
    if (!root) c->noise_needs_wait = true;
    c->noise_seed = seed, c->noise_offset = offset;
    return VBMC_OK;
}

}  // namespace

bool entmc_tc_supported(int DP, int K) {
    return DP > 0 && K >= 1 && tc_kp(K) <= 16 * 2 * kMaxChunks && tc_tmem_cols(DP, K) <= 512;
}

// chunk (multiple of 128 pairs), grid and shared memory of the tensor-core kernel
int entmc_tc_plan(const Ctx *c, int D, int K, int64_t half_local, EntmcPlan *plan) {
    const int DP = pad_dim(D);
    const uint32_t cols = tc_tmem_cols(DP, K);
    int per_sm = (int)(512 / cols);  // tensor memory: 512 columns per SM
    if (per_sm > 2) per_sm = 2;      // __launch_bounds__(256, 2)
    static int env_per_sm = -1;
    if (env_per_sm < 0) {
        const char *e = getenv("VBMC_TC_PER_SM");
        env_per_sm = e ? atoi(e) : 0;
    }
    if (env_per_sm >= 1 && env_per_sm < per_sm) per_sm = env_per_sm;
    size_t smem = tc_smem_layout(DP, K, entpart_stride(DP, K)).total;
    VBMC_REQUIRE(smem <= 227 * 1024, VBMC_ERR_UNSUPPORTED, "entmc (tensor-core): tables do not fit in shared memory");
    // never let more CTAs become resident than tensor memory can serve (tcgen05.alloc would spin): pad the
    // shared-memory request so that exactly per_sm CTAs fit
    const size_t floor_smem = (size_t)(227 * 1024) / (per_sm + 1) + 1024;
    if (smem < floor_smem) smem = floor_smem;
    const int64_t T = (int64_t)K * half_local;
    const int64_t slots = (int64_t)c->sm_count * per_sm;
    // tiles of 128 pairs over `slots` resident CTAs: the first n_big CTAs take one tile more than the others.  (A
    // uniform chunk of ceil(tiles / slots) tiles left SMs with 12 tiles next to SMs with 6 at C3: 0.88 waves.)
    const int64_t n_tiles = (T + kTile - 1) / kTile;
    int64_t base = n_tiles / slots, rem = n_tiles - base * slots;
    static int env_even = -1;
    if (env_even < 0) {
        const char *e = getenv("VBMC_TC_EVEN_CHUNKS");
        env_even = e ? atoi(e) : 0;
    }
    int64_t chunk, chunk_small;
    int grid, n_big;
    if (base == 0) {  // fewer tiles than CTA slots: one tile per CTA
        chunk = chunk_small = kTile;
        grid = (int)std::max<int64_t>(1, n_tiles), n_big = grid;
    } else if (rem == 0 || env_even) {
        chunk = chunk_small = (base + (rem ? 1 : 0)) * kTile;
        grid = (int)((T + chunk - 1) / chunk), n_big = grid;
    } else {
        chunk = (base + 1) * kTile, chunk_small = base * kTile;
        grid = (int)slots, n_big = (int)rem;
    }
    plan->variant = ENTMC_TC;
    plan->threads = kThreads;
    plan->pairs_per_thread = (int)(chunk / kTile);
    plan->chunk = chunk;
    plan->chunk_small = chunk_small;
    plan->n_big = n_big;
    plan->grid = grid;
    plan->maxseg = half_local > 0 ? (int)((chunk - 1) / half_local) + 2 : 1;
    plan->slabs = plan->grid * plan->maxseg;
    plan->half = half_local;
    plan->pair0 = 0;
    plan->half_glob = half_local;
    plan->smem = smem;
    return VBMC_OK;
}

int entmc_tc_prefetch(Ctx *c, ParamLayout lay, const EntmcPlan &plan, uint64_t seed, uint64_t offset, bool root) {
    switch (lay.DP) {
#define VBMC_CASE(N) \
    case N:          \
        return tc_prefetch_dp<N>(c, lay, plan, seed, offset, root)
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

// philox.cuh -- counter-based standard-normal draws for the Monte-Carlo entropy.
//
// The reference draws eps from NumPy's global MT19937 (pyvbmc/entropy/entmc_vbmc.py:64-68);
// a sequential generator cannot be sharded, so production mode keys Philox4x32-10 with
// (seed, offset | component j, pair index, 4-dim block): the draws are a pure function of
// the key and therefore identical for any grid shape or GPU count.
#pragma once
#include <stdint.h>

namespace vbmc {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// two uniforms -> two standard normals (Box-Muller, fp32 fast intrinsics)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &z0, float &z1) {
    const float u1 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-08f;  // (0, 1)
    const float u2 = (float)(b >> 8) * 5.9604644775390625e-08f;           // [0, 1)
    const float r = sqrtf(-1.3862943611198906f * __log2f(u1));            // sqrt(-2 ln u1)
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// The D (padded: DP) dimensions of a draw are generated in two halves of DH = DP/2 dimensions:
// half h in {0,1} owns dims [h*DH, h*DH + DH) and takes its normals from the Philox blocks
// (pair, j | b << 20 | h << 28), b = 0 .. ceil(DH/4)-1.  The split mirrors the fp32 entropy kernel,
// where two adjacent lanes evaluate one antithetic pair, each lane holding one half of the dims.
template <int DH>
__device__ __forceinline__ void philox_normals_half(uint64_t seed, uint64_t offset, uint32_t j, uint64_t pair,
                                                    int h, int D, float (&z)[DH]) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    constexpr int NB = (DH + 3) / 4;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint4 ctr = make_uint4((uint32_t)pair, j | ((uint32_t)b << 20) | ((uint32_t)h << 28), (uint32_t)offset,
                                     (uint32_t)(offset >> 32) ^ (uint32_t)(pair >> 32));
        const uint4 x = philox4x32_10(ctr, key);
        float n[4];
        box_muller(x.x, x.y, n[0], n[1]);
        box_muller(x.z, x.w, n[2], n[3]);
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (4 * b + t < DH) z[4 * b + t] = (h * DH + 4 * b + t < D) ? n[t] : 0.0f;
    }
}

// z[0..D) ~ N(0,1) for (component j, antithetic pair `pair`); z[d >= D] = 0
template <int DP>
__device__ __forceinline__ void philox_normals(uint64_t seed, uint64_t offset, uint32_t j, uint64_t pair, int D,
                                               float (&z)[DP]) {
    constexpr int DH = DP / 2;
    float a[DH], b[DH];
    philox_normals_half<DH>(seed, offset, j, pair, 0, D, a);
    philox_normals_half<DH>(seed, offset, j, pair, 1, D, b);
#pragma unroll
    for (int i = 0; i < DH; ++i) z[i] = a[i], z[DH + i] = b[i];
}

}  // namespace vbmc

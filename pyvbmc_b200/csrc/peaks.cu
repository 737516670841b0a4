// peaks.cu -- measurement-only kernels: FP32 / FP64 FMA issue peak of this GPU.
//
// MEASURED_PEAKS.json holds the HBM copy bandwidth and the cuBLAS bf16 peak; the Monte-Carlo
// entropy kernel is bound by the FP32 FMA pipe, so bench.py measures that denominator itself.
#include "common.cuh"

namespace vbmc {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
            x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// packed f32x2 variant (FFMA2): same flop count per thread, half the instructions
__global__ void __launch_bounds__(256) fma2_peak_kernel(float *out, int iters, float a, float b) {
    float2 x0 = make_float2((float)threadIdx.x, 1.f), x1 = x0, x2 = x0, x3 = x0;
    x1.x += 1.f, x2.x += 2.f, x3.x += 3.f;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __ffma2_rn(x0, a2, b2), x1 = __ffma2_rn(x1, a2, b2);
            x2 = __ffma2_rn(x2, a2, b2), x3 = __ffma2_rn(x3, a2, b2);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y;
}

template <typename T>
int run_peak(Ctx *c, double *tflops, bool packed = false) {
    const int ctas = c->sm_count * 8, nt = 256, iters = 2048;
    T *buf = nullptr;
    VBMC_CUDA_CHECK(cudaMalloc((void **)&buf, (size_t)ctas * nt * sizeof(T)));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
        if (packed)
            fma2_peak_kernel<<<ctas, nt, 0, c->stream>>>((float *)buf, iters, 0.999f, 0.001f);
        else
            fma_peak_kernel<T><<<ctas, nt, 0, c->stream>>>(buf, iters, (T)0.999, (T)0.001);
        VBMC_CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
        VBMC_CUDA_CHECK(cudaEventSynchronize(c->ev1));
        float ms = 0;
        VBMC_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const double flops = 2.0 * 64.0 * iters * (double)ctas * nt;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        c->launches++;
    }
    VBMC_CUDA_CHECK(cudaFree(buf));
    *tflops = best;
    return VBMC_OK;
}

}  // namespace
}  // namespace vbmc

extern "C" int vbmc_fma_peak(vbmc_ctx *p, int fp64, double *tflops) {
    using namespace vbmc;
    VBMC_REQUIRE(p && tflops, VBMC_ERR_ARG, "fma_peak: null argument");
    Ctx *c = reinterpret_cast<Ctx *>(p);  // CtxEx starts with its Ctx
    cudaSetDevice(c->device);
    if (fp64 == 2) return run_peak<float>(c, tflops, true);  // packed FFMA2
    return fp64 ? run_peak<double>(c, tflops) : run_peak<float>(c, tflops);
}

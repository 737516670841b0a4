// Micro-benchmark (GPU box): what do a cluster launch, zero-copy host writes and a D2H copy cost per step?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o launch_lat launch_lat.cu && ./launch_lat
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;
__global__ void k_plain(double *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)i;
}
__global__ void k_cluster(double *out, int n) {
    cg::this_cluster().sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)i;
}
__global__ void k_spin(long long cycles) {
    long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}
template <class F> float timeit(cudaStream_t s, F f, int reps = 200) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    for (int i = 0; i < 20; ++i) f();
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / reps;
}
int main() {
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    double *d, *h;
    const int n = 1128;
    cudaMalloc(&d, 8 * 4096);
    cudaMallocHost(&h, 8 * 4096);
    auto cl = [&](double *o, int nn) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(8), cfg.blockDim = dim3(512), cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_cluster, o, nn);
    };
    // back-to-back dependent launches: spin kernel (20 us) followed by X; report X's marginal cost
    const long long spin = 20 * 1900;
    float base = timeit(s, [&] { k_spin<<<148, 128, 0, s>>>(spin); });
    printf("spin alone                         %7.2f us\n", base);
    printf("spin + plain kernel -> device      %7.2f us\n", timeit(s, [&] { k_spin<<<148, 128, 0, s>>>(spin); k_plain<<<8, 512, 0, s>>>(d, n); }) - base);
    printf("spin + plain kernel -> pinned host %7.2f us\n", timeit(s, [&] { k_spin<<<148, 128, 0, s>>>(spin); k_plain<<<8, 512, 0, s>>>(h, n); }) - base);
    printf("spin + plain -> host 3x entries    %7.2f us\n", timeit(s, [&] { k_spin<<<148, 128, 0, s>>>(spin); k_plain<<<8, 512, 0, s>>>(h, 3 * n); }) - base);
    printf("spin + cluster kernel -> device    %7.2f us\n", timeit(s, [&] { k_spin<<<148, 128, 0, s>>>(spin); cl(d, n); }) - base);
    printf("spin + cluster kernel -> host      %7.2f us\n", timeit(s, [&] { k_spin<<<148, 128, 0, s>>>(spin); cl(h, n); }) - base);
    printf("spin + plain -> device + D2H copy  %7.2f us\n", timeit(s, [&] { k_spin<<<148, 128, 0, s>>>(spin); k_plain<<<8, 512, 0, s>>>(d, n); cudaMemcpyAsync(h, d, 8 * n, cudaMemcpyDeviceToHost, s); }) - base);
    printf("spin + H2D copy 10KB before        %7.2f us\n", timeit(s, [&] { cudaMemcpyAsync(d, h, 10320, cudaMemcpyHostToDevice, s); k_spin<<<148, 128, 0, s>>>(spin); }) - base);
    // sync-per-step variants (what the e2e path does): launch, then host waits
    auto sync_step = [&](auto f) {
        for (int i = 0; i < 20; ++i) { f(); cudaStreamSynchronize(s); }
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, s);
        for (int i = 0; i < 200; ++i) { f(); cudaStreamSynchronize(s); }
        cudaEventRecord(b, s); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        return ms * 1e3f / 200;
    };
    printf("sync/step: spin only               %7.2f us\n", sync_step([&] { k_spin<<<148, 128, 0, s>>>(spin); }));
    printf("sync/step: spin + plain->host      %7.2f us\n", sync_step([&] { k_spin<<<148, 128, 0, s>>>(spin); k_plain<<<8, 512, 0, s>>>(h, n); }));
    printf("sync/step: spin + plain->dev + D2H %7.2f us\n", sync_step([&] { k_spin<<<148, 128, 0, s>>>(spin); k_plain<<<8, 512, 0, s>>>(d, n); cudaMemcpyAsync(h, d, 8 * n, cudaMemcpyDeviceToHost, s); }));
    printf("sync/step: H2D + spin + plain->host%7.2f us\n", sync_step([&] { cudaMemcpyAsync(d, h, 10320, cudaMemcpyHostToDevice, s); k_spin<<<148, 128, 0, s>>>(spin); k_plain<<<8, 512, 0, s>>>(h, n); }));
    // kernel reads params straight from pinned host memory instead of an H2D copy
    return 0;
}

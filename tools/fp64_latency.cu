// DFMA issue/latency probe for sm_100a: throughput of dependent DFMA chains as a
// function of warps per SM and independent chains per thread (ILP).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chains(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

template <int ILP>
void run(int warps_per_sm, double* d) {
    const int iters = 4000;
    chains<ILP><<<148, 32 * warps_per_sm>>>(d, iters, 0.999, 1e-3);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chains<ILP><<<148, 32 * warps_per_sm>>>(d, iters, 0.999, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d + 148 * 32 * warps_per_sm, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * ILP;                       // DFMA per thread
    printf("{\"warps_per_sm\": %d, \"ilp\": %d, \"cycles_per_dfma_per_warp\": %.2f, \"tflops\": %.2f}\n",
           warps_per_sm, ILP, cyc / n, 2.0 * n * 148 * 32 * warps_per_sm / (ms * 1e-3) / 1e12);
}

int main() {
    double* d; cudaMalloc(&d, 8 * (148 * 1024 + 8));
    const int ws[] = { 1, 4, 6, 8, 12, 16, 32 };
    for (int w : ws) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); }
    return 0;
}

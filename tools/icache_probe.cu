// Instruction-cache probe for sm_100a: one warp per SM scheduler loops over a straight-line body of
// N FFMA instructions (16 bytes each); cycles per instruction rise once the body no longer fits.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/icache_probe tools/icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void body(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N / 4; i++) { x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (float)(t1 - t0) / ((float)iters * N);
}
template <int N>
void run(float* d, int warps) {
    body<N><<<148, 32 * warps>>>(d, 200, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    body<N><<<148, 32 * warps>>>(d, 200, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    float c; cudaMemcpy(&c, d + 148 * 32 * warps, 4, cudaMemcpyDeviceToHost);
    printf("{\"body_instructions\": %d, \"body_kib\": %.0f, \"warps_per_sm\": %d, \"cycles_per_instr_per_warp\": %.2f}\n", N, N * 16 / 1024.0, warps, c);
}
int main() {
    float* d; cudaMalloc(&d, 4 * (148 * 1024 + 8));
    for (int w = 1; w <= 4; w *= 4) {
        run<512>(d, w); run<1024>(d, w); run<2048>(d, w); run<3072>(d, w); run<4096>(d, w); run<6144>(d, w); run<8192>(d, w); run<12288>(d, w); run<16384>(d, w);
    }
    return 0;
}

// FP64 pipe probe for sm_100a: issue cost of DFMA / DMUL / DADD with three DISTINCT register
// operands (the common case in real code), against the accumulate-only DFMA of fp64_peak.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mix tools/fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(double* out, int iters, double s) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0 + 1e-9 * (i + threadIdx.x); c[i] = 1e-7 * i + s; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = fma(a[i], s, s);                    // 1 varying operand
            if (MODE == 1) a[i] = fma(b[i], c[(i + 1) & 7], a[i]);    // 3 distinct register operands
            if (MODE == 2) a[i] = __dmul_rn(a[i], b[i]);              // DMUL, 2 distinct
            if (MODE == 3) a[i] = __dadd_rn(a[i], c[i]);              // DADD, 2 distinct
            if (MODE == 4) { a[i] = fma(b[i], c[(i + 1) & 7], a[i]); b[i] = __dadd_rn(b[i], c[i]); }   // DFMA + DADD mix
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

template <int MODE>
void run(const char* name, int warps, double* d) {
    const int iters = 4000;
    probe<MODE><<<148, 32 * warps>>>(d, iters, 0.999);
    cudaDeviceSynchronize();
    probe<MODE><<<148, 32 * warps>>>(d, iters, 0.999);
    cudaDeviceSynchronize();
    double cyc; cudaMemcpy(&cyc, d + 148 * 32 * warps, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * (MODE == 4 ? 2 : 1);
    printf("{\"op\": \"%s\", \"warps_per_sm\": %d, \"cycles_per_instr_per_warp\": %.2f, \"warp_instr_per_cycle_per_sm\": %.2f}\n",
           name, warps, cyc / n, warps * n / cyc);
}

int main() {
    double* d; cudaMalloc(&d, 8 * (148 * 1024 + 8));
    const int ws[] = { 1, 4, 8, 16 };
    for (int w : ws) {
        run<0>("dfma_1_operand", w, d); run<1>("dfma_3_operands", w, d); run<2>("dmul", w, d); run<3>("dadd", w, d); run<4>("dfma+dadd", w, d);
    }
    return 0;
}

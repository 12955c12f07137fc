// DMMA pilot (north_star: "tensor cores only if ncu shows the contraction dominates").
// 1. peak of mma.sync.aligned.m8n8k4.row.col.f64 on this GPU (independent accumulator chains per warp);
// 2. the congruence K_ij = S^T C'_ij S of one Shell_1 Gauss point (S: 5 x 9, C'_ij: 5 x 5, nine (i,j) pairs) as
//    m8n8k4 tiles -- 12 MMAs per (i,j): (5x5)(5x9) padded to (8x8)(8x16) = 4, (9x5)(5x9) padded to (16x8)(8x16) = 8 --
//    against the same product with scalar DFMA on the 3x3-block structure (what phase B of shell::eval_kernel does).
// Prints TFLOP/s of issued tensor flops, and useful congruences per second for both forms.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CHAINS>
__global__ void dmma_peak(double* out, int iters) {
    double acc[CHAINS][2];
    for (int c = 0; c < CHAINS; c++) { acc[c][0] = threadIdx.x; acc[c][1] = c; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) dmma(acc[c][0], acc[c][1], a, b);
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += acc[c][0] + acc[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// one warp = one Gauss point per iteration: 9 x 12 MMAs (operands from registers; fragment shuffling of a real kernel not charged)
__global__ void congruence_dmma(double* out, int iters) {
    double acc[12][2];
    for (int c = 0; c < 12; c++) { acc[c][0] = threadIdx.x; acc[c][1] = c; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; i++)
        for (int ij = 0; ij < 9; ij++)
#pragma unroll
            for (int c = 0; c < 12; c++) dmma(acc[c][0], acc[c][1], a + ij, b);
    double s = 0;
    for (int c = 0; c < 12; c++) s += acc[c][0] + acc[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// scalar form: one LANE = one (element, K column); per Gauss point the m-step (C' column x S) and the row loop,
// 165 DFMA per point for a u column (the mix of shell::uu_item / rot_item averaged), operands in registers
__global__ void congruence_dfma(double* out, int iters) {
    double k[21], m[6];
    for (int i = 0; i < 21; i++) k[i] = threadIdx.x + i;
    const double s1 = 1.0 + 1e-9 * threadIdx.x, s2 = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 6; i++) m[i] = fma(s1, k[i], s2 * k[i + 6]);
#pragma unroll
        for (int a = 0; a < 7; a++)
#pragma unroll
            for (int i = 0; i < 3; i++) k[3 * a + i] = fma(s2, m[3 + i], fma(s1, m[i], k[3 * a + i]));
    }
    double s = 0;
    for (int i = 0; i < 21; i++) s += k[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
    const int iters = 20000, grid = 148 * 8, block = 256;
    const double warps = (double)grid * block / 32;
    float ms;
    ms = timeit([&] { dmma_peak<1><<<grid, block>>>(out, iters); });
    printf("{\"dmma_m8n8k4_chains1_tflops\": %.2f,\n", warps * iters * 1 * 512.0 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { dmma_peak<4><<<grid, block>>>(out, iters); });
    printf(" \"dmma_m8n8k4_chains4_tflops\": %.2f,\n", warps * iters * 4 * 512.0 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { dmma_peak<8><<<grid, block>>>(out, iters); });
    const double peak = warps * iters * 8 * 512.0 / (ms * 1e-3) / 1e12;
    printf(" \"dmma_m8n8k4_chains8_tflops\": %.2f,\n", peak);
    const int it2 = 2000;
    ms = timeit([&] { congruence_dmma<<<grid, block>>>(out, it2); });
    const double gp_dmma = warps * it2 / (ms * 1e-3);
    printf(" \"congruence_dmma_gauss_points_per_s\": %.3e, \"issued_tflops\": %.2f, \"useful_fraction\": %.3f,\n", gp_dmma, gp_dmma * 108 * 512 / 1e12, 9 * 1260.0 / (108 * 512));
    ms = timeit([&] { congruence_dfma<<<grid, block>>>(out, it2 * 10); });
    // a Shell_1 element has 27 columns and 3 Gauss points; one loop pass = one (column, Gauss point) at ~48 DFMA
    const double colgp = (double)grid * block * it2 * 10 / (ms * 1e-3);
    printf(" \"congruence_dfma_column_points_per_s\": %.3e, \"dfma_tflops\": %.2f,\n", colgp, colgp * 54 * 2 / 1e12);
    printf(" \"shell_elements_per_s_phaseB_dmma\": %.3e, \"shell_elements_per_s_phaseB_dfma\": %.3e}\n", gp_dmma / 3.0, colgp / 81.0);
    return 0;
}

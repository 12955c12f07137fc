// Can a thin kernel become resident BESIDE a fat persistent kernel on every SM?  (ring pipeline)
// A: one CTA per SM, TA threads, NREG live doubles per thread, SMEM bytes of dynamic shared memory; spins until a
// flag is set (or gives up after 200 ms).  B: one CTA per SM, 256 threads, sets the flag.  Prints whether A saw it.
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long gns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int NREG>
__global__ void __launch_bounds__(224, 1) A(volatile int* flag, double* out, int* seen) {
    extern __shared__ double sm[];
    double v[NREG];
#pragma unroll
    for (int i = 0; i < NREG; i++) v[i] = out[(size_t)i * 148 * 256 + blockIdx.x * blockDim.x + threadIdx.x];
    const unsigned long long t0 = gns();
    int ok = 0;
    while (gns() - t0 < 200000000ULL) { if (*flag) { ok = 1; break; } __nanosleep(200); }
    if (threadIdx.x == 0) { sm[0] = 1.0; seen[blockIdx.x] = ok; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NREG; i++) out[(size_t)i * 148 * 256 + blockIdx.x * blockDim.x + threadIdx.x] = v[i] + sm[0];
}
__global__ void __launch_bounds__(256, 8) B(int* flag, int* nb) { if (threadIdx.x == 0) { atomicAdd(nb, 1); *flag = 1; __threadfence(); } }
template <int NREG> void run(int ta, int smem, bool carve, bool b_first) {
    int *flag, *seen, *nb; double* out;
    cudaMalloc(&flag, 4); cudaMalloc(&nb, 4); cudaMalloc(&seen, 148 * 4); cudaMalloc(&out, (size_t)128 * 148 * 256 * 8); cudaMemset(out, 0, (size_t)128 * 148 * 256 * 8);
    cudaMemset(flag, 0, 4); cudaMemset(nb, 0, 4); cudaMemset(seen, 0, 148 * 4);
    cudaFuncSetAttribute(A<NREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (carve) { cudaFuncSetAttribute(A<NREG>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); cudaFuncSetAttribute(B, cudaFuncAttributePreferredSharedMemoryCarveout, 100); }
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, A<NREG>);
    if (b_first) { /* not meaningful here: B exits at once */ }
    A<NREG><<<148, ta, smem, s1>>>(flag, out, seen);
    B<<<148, 256, 0, s2>>>(flag, nb);
    cudaError_t e = cudaDeviceSynchronize();
    int h[148]; cudaMemcpy(h, seen, sizeof(h), cudaMemcpyDeviceToHost);
    int n = 0; for (int i = 0; i < 148; i++) n += h[i];
    printf("A: %d threads, %d regs, %d B smem, carve-out hint %d -> %d of 148 CTAs saw B's flag while running (%s)\n", ta, fa.numRegs, smem, (int)carve, n, cudaGetErrorString(e));
    cudaFree(flag); cudaFree(seen); cudaFree(out); cudaFree(nb);
}
int main() {
    for (int carve = 0; carve < 2; carve++) {
        run<8>(224, 32 * 1024, carve, false);
        run<8>(224, 225792, carve, false);
        run<120>(128, 129024, carve, false);
        run<120>(160, 161280, carve, false);
        run<120>(192, 193536, carve, false);
        run<120>(224, 225792, carve, false);
        run<120>(224, 32768, carve, false);
        run<100>(224, 225792, carve, false);
        run<80>(224, 225792, carve, false);
    }
    return 0;
}

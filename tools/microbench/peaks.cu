// Secondary ceilings of the frame pipeline on this GPU ("microbenchmark it on the box"):
//   * FP64 FMA issue rate (the raster / shade kernels are float64 by contract),
//   * shared-memory atomics: 64-bit atomicMin (z-buffer keys) and 32-bit atomicAdd (stencil counts),
//   * global (L2) 32-bit atomicAdd on distinct addresses (tile-list binning).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/peaks tools/microbench/peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void k_smem_min64(unsigned long long* out, int iters) {
    __shared__ unsigned long long s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = ~0ull;
    __syncthreads();
    unsigned long long v = 0x7000000000000000ull - threadIdx.x;
    for (int i = 0; i < iters; ++i) { atomicMin(&s[(threadIdx.x * 7 + i * 33) & 1023], v); v -= 1024; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}

__global__ void k_smem_add32(int* out, int iters) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = 0;
    __syncthreads();
    for (int i = 0; i < iters; ++i) atomicAdd(&s[(threadIdx.x * 7 + i * 33) & 1023], 1);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}

__global__ void k_gmem_add32(int* buf, int n, int iters) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < iters; ++i) atomicAdd(&buf[(id * 9 + i * 4099u) % (unsigned)n], 1);
}

template <class F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    unsigned long long* u; cudaMalloc(&u, 8 * blocks);
    int* o; cudaMalloc(&o, 4 * blocks);
    const int n = 1 << 24; int* g; cudaMalloc(&g, 4 * n); cudaMemset(g, 0, 4 * n);
    const int it = 1 << 14;
    float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(d, it); });
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, sms);
    printf(" \"fp64_fma_tflops\": %.2f,\n", 2.0 * 8 * it * (double)blocks * threads / ms / 1e9);
    const int ia = 1 << 12;
    ms = time_ms([&] { k_smem_min64<<<blocks, threads>>>(u, ia); });
    printf(" \"smem_atomic_min_u64_Gops\": %.1f,\n", (double)ia * blocks * threads / ms / 1e6);
    ms = time_ms([&] { k_smem_add32<<<blocks, threads>>>(o, ia); });
    printf(" \"smem_atomic_add_i32_Gops\": %.1f,\n", (double)ia * blocks * threads / ms / 1e6);
    const int ig = 1 << 8;
    ms = time_ms([&] { k_gmem_add32<<<blocks, threads>>>(g, n, ig); });
    printf(" \"l2_atomic_add_i32_Gops\": %.1f}\n", (double)ig * blocks * threads / ms / 1e6);
    return 0;
}

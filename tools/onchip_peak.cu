// onchip_peak.cu — measured on-chip roofs for kernels whose messages live in shared memory (bench.py's roofline block).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/onchip_peak tools/onchip_peak.cu
//   ./tools/onchip_peak > profiles/onchip_peaks.json          (on the B200; a few seconds)
//
// Three streaming loops, each at the path kernel's launch shape (128-thread CTAs, 8 per SM, 27 KB of dynamic shared
// memory each) and at one 1024-thread CTA per SM:
//   smem_ld64_st64   conflict-free LDS.64 + STS.64 over the CTA's buffer (bytes read + written per second)
//   smem_ld64        conflict-free LDS.64 only
//   issue            independent 32-bit integer chains (warp instructions per second: the issue roof)
//   fp64             independent DFMA chains (the fp64 pipe's roof, warp instructions per second)
// Timed with CUDA events, best of 5 after a warm-up launch.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

extern __shared__ __align__(16) unsigned char smem[];

__global__ void smem_rw_kernel(int words, int iters, double *sink) {
    double *buf = (double *)smem;
    const int T = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < words; i += T) buf[i] = (double)i;
    __syncthreads();
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll 4
        for (int i = tid; i < words; i += T) { const double v = buf[i]; acc += v; buf[i] = acc; }
    }
    if (acc == 123.456) sink[0] = acc;
}

__global__ void smem_r_kernel(int words, int iters, double *sink) {
    double *buf = (double *)smem;
    const int T = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < words; i += T) buf[i] = (double)i;
    __syncthreads();
    unsigned long long acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll 8
        for (int i = tid; i < words; i += T) acc ^= (unsigned long long)__double_as_longlong(buf[i]) + it;
    }
    if (acc == 0x123456789ull) sink[0] = (double)acc;
}

__global__ void issue_kernel(int iters, unsigned *sink) {
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {      // 8 independent chains x 16 = 128 ALU instructions per trip
            a0 = a0 * 3u + 1u; a1 = (a1 ^ 0x9e3779b9u) + a0 * 0u + 7u; a2 = a2 * 5u + 3u; a3 = (a3 ^ 0x7f4a7c15u) + 11u;
            a4 = a4 * 7u + 5u; a5 = (a5 ^ 0x85ebca6bu) + 13u; a6 = a6 * 9u + 7u; a7 = (a7 ^ 0xc2b2ae35u) + 17u;
        }
    }
    const unsigned r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0xdeadbeefu) sink[0] = r;
}

__global__ void fp64_kernel(int iters, double *sink) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) sink[0] = r;
}

template <typename F>
static float best_ms(F launch) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *sink; CK(cudaMalloc(&sink, 64));
    CK(cudaFuncSetAttribute(smem_rw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(smem_r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(fp64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d, \"how\": \"tools/onchip_peak.cu, CUDA events, best of 5\", \"shapes\": [", p.name, sms, clk);
    const int shapes[2][3] = {{128, 8, 27 * 1024}, {1024, 1, 200 * 1024}};      // threads, CTAs per SM, dynamic smem bytes
    for (int s = 0; s < 2; s++) {
        const int T = shapes[s][0], per = shapes[s][1], bytes = shapes[s][2];
        const int words = bytes / 8, grid = sms * per;
        const int iters = (s == 0) ? 2000 : 300;
        int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, smem_rw_kernel, T, bytes));
        const float ms_rw = best_ms([&] { smem_rw_kernel<<<grid, T, bytes>>>(words, iters, sink); });
        const float ms_r = best_ms([&] { smem_r_kernel<<<grid, T, bytes>>>(words, iters, sink); });
        const int it_issue = 4000;
        const float ms_i = best_ms([&] { issue_kernel<<<grid, T, bytes>>>(it_issue, (unsigned *)sink); });
        const float ms_f = best_ms([&] { fp64_kernel<<<grid, T, bytes>>>(it_issue, sink); });
        const double rw_gbs = (double)grid * words * 16.0 * iters / (ms_rw * 1e6);
        const double r_gbs = (double)grid * words * 8.0 * iters / (ms_r * 1e6);
        const double issue = (double)grid * (T / 32) * 128.0 * it_issue / (ms_i * 1e-3);      // ALU warp instructions / s (loop overhead not counted)
        const double f64 = (double)grid * (T / 32) * 64.0 * it_issue / (ms_f * 1e-3);
        printf("%s{\"threads\": %d, \"ctas_per_sm\": %d, \"resident_ctas_per_sm\": %d, \"smem_bytes_per_cta\": %d, "
               "\"smem_ld64_st64_gbs\": %.1f, \"smem_ld64_gbs\": %.1f, \"issue_gwarp_inst_per_s\": %.2f, \"fp64_gwarp_inst_per_s\": %.2f, "
               "\"fp64_tflops\": %.2f}",
               s ? ", " : "", T, per, occ, bytes, rw_gbs, r_gbs, issue / 1e9, f64 / 1e9, f64 * 32 * 2 / 1e12);
    }
    printf("]}\n");
    return 0;
}

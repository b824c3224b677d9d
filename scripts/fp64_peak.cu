// Measures the sustained fp64 FMA rate of the device (MEASURED_PEAKS.json has HBM and bf16 only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_peak scripts/fp64_peak.cu && scripts/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double fmas = (double)blocks * threads * iters * 8;
    std::printf("%s: %d SMs, fp64 FMA %.2f TFMA/s = %.2f TFLOP/s (%.1f FMA/clk/SM at %.0f MHz nominal)\n", p.name, p.multiProcessorCount,
                fmas / best * 1e-9, 2 * fmas / best * 1e-9, fmas / (best * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate * 1e-3);
    return 0;
}

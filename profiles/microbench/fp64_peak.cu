// FP64 FMA latency / throughput microbenchmark for the backward-pass roofline (SURVEY.md 8(d):
// "FP64/FP32 vector peaks are not in MEASURED_PEAKS.json - measure with an FMA microbenchmark").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void fma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int blocks_per_sm, int threads, int sms, double clock_ghz) {
    double* out;
    int grid = sms * blocks_per_sm;
    cudaMalloc(&out, sizeof(double) * grid * threads);
    int iters = 20000;
    fma_kernel<ILP><<<grid, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    cudaEventRecord(s);
    fma_kernel<ILP><<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms;
    cudaEventElapsedTime(&ms, s, e);
    double fmas = (double)grid * threads * iters * ILP;
    double tflops = 2 * fmas / (ms * 1e-3) / 1e12;
    double warps_per_smsp = blocks_per_sm * threads / 32.0 / 4.0;
    double cyc_per_warp_fma = (ms * 1e-3 * clock_ghz * 1e9) / ((double)iters * ILP) / warps_per_smsp;   // SMSP cycles per warp-instruction
    printf("ILP=%d warps/SMSP=%.2f : %.2f TFLOP/s, %.2f SMSP-cycles per warp-DFMA (at %.3f GHz)\n", ILP, warps_per_smsp, tflops,
           cyc_per_warp_fma, clock_ghz);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double ghz = p.clockRate * 1e-6;
    printf("%s: %d SMs, %.3f GHz max\n", p.name, sms, ghz);
    run<1>(1, 128, sms, ghz);     // 1 warp/SMSP, dependent chain -> latency
    run<2>(1, 128, sms, ghz);
    run<4>(1, 128, sms, ghz);
    run<8>(1, 128, sms, ghz);
    run<1>(2, 128, sms, ghz);
    run<4>(2, 128, sms, ghz);
    run<8>(2, 128, sms, ghz);
    run<8>(4, 128, sms, ghz);
    run<4>(8, 128, sms, ghz);
    run<8>(8, 256, sms, ghz);
    return 0;
}

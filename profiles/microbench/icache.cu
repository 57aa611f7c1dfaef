// How large may a hot loop body be before instruction fetch limits issue?  Straight-line FFMA
// bodies of N instructions (8 independent accumulators), 8 warps per SM as in the solve kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void __launch_bounds__(128) body(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i & 7] = fmaf(x[i & 7], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N>
void run(int sms, int ctas_per_sm, double ghz) {
    float* out;
    int grid = sms * ctas_per_sm;
    cudaMalloc(&out, sizeof(float) * grid * 128);
    int iters = (1 << 22) / N;
    body<N><<<grid, 128>>>(out, 4, 1.0000001f, 1e-9f);
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    cudaEventRecord(s);
    body<N><<<grid, 128>>>(out, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms;
    cudaEventElapsedTime(&ms, s, e);
    double warp_instr = (double)grid * 4 * iters * N;
    double ipc_sm = warp_instr / (ms * 1e-3 * ghz * 1e9) / sms;
    printf("body %6d instr (%4d KB), %d warps/SM: IPC per SM = %.2f\n", N, N * 16 / 1024, ctas_per_sm * 4, ipc_sm);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double ghz = p.clockRate * 1e-6;
    for (int c = 1; c <= 2; ++c) {
        run<256>(p.multiProcessorCount, c, ghz);
        run<512>(p.multiProcessorCount, c, ghz);
        run<1024>(p.multiProcessorCount, c, ghz);
        run<2048>(p.multiProcessorCount, c, ghz);
        run<4096>(p.multiProcessorCount, c, ghz);
        run<8192>(p.multiProcessorCount, c, ghz);
        run<16384>(p.multiProcessorCount, c, ghz);
    }
    return 0;
}

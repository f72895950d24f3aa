// Microbenchmark: fp64 issue-rate ceilings on B200 (sm_100a).
// DMMA.8x8x4 (the only f64 tensor shape ptxas emits for sm_100a) vs scalar DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, const double* in, int iters) {
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, const double* in, int iters) {
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double *in, *out; cudaMalloc(&in, 1024 * 8); cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8);
    cudaMemset(in, 0, 1024 * 8);
    const int iters = 4096;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int grid = sms * bps;
            float ms = time_it([&] { dmma_kernel<16><<<grid, threads>>>(out, in, iters); });
            double flop = 2.0 * 256 * 16 * (double)iters * (threads / 32) * grid;
            printf("DMMA.8x8x4 nacc=16 threads=%4d blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", threads, bps, ms, flop / ms * 1e-9);
            ms = time_it([&] { dmma_kernel<4><<<grid, threads>>>(out, in, iters); });
            flop = 2.0 * 256 * 4 * (double)iters * (threads / 32) * grid;
            printf("DMMA.8x8x4 nacc=4  threads=%4d blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", threads, bps, ms, flop / ms * 1e-9);
            ms = time_it([&] { dfma_kernel<16><<<grid, threads>>>(out, in, iters); });
            flop = 2.0 * 16 * (double)iters * threads * grid;
            printf("DFMA       nacc=16 threads=%4d blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", threads, bps, ms, flop / ms * 1e-9);
        }
    }
    // sustained: ~3 s of DMMA to see the power-capped rate
    {
        int grid = sms * 2, threads = 512;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int reps = 0; float ms = 0;
        while (ms < 3000.f) {
            for (int k = 0; k < 20; k++) dmma_kernel<16><<<grid, threads>>>(out, in, iters);
            reps += 20; cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double flop = 2.0 * 256 * 16 * (double)iters * (threads / 32) * grid * reps;
        printf("DMMA sustained %.0f ms: %.2f TFLOP/s\n", ms, flop / ms * 1e-9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// FP64 throughput of the FMA pipe vs the DMMA (mma.sync.m8n8k4.f64) tensor path on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_pipes.cu -o fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
    double a = threadIdx.x * 1e-9 + 1.0, b = 0.999999;
    double acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], b, a);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters) {
    double a = threadIdx.x * 1e-9 + 1.0, b = 0.999999;
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 8, threads = 256;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * iters * (double)blocks * threads;
        printf("DFMA : %.3f ms  %.2f TFLOP/s\n", ms, fl / (ms * 1e-3) * 1e-12);
        cudaEventRecord(e0);
        k_dmma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * (threads / 32);
        printf("DMMA : %.3f ms  %.2f TFLOP/s\n", ms, fl / (ms * 1e-3) * 1e-12);
    }
    return 0;
}

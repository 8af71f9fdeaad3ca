// DMMA.8x8x4 throughput as a function of resident warps per SM and of independent accumulators per warp (sm_100a).
// Question it answers: can 16 warps / SM (the multifrontal GEMM's occupancy) saturate the FP64 tensor path, and how
// many non-DMMA instructions per DMMA can be interleaved before the rate drops?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a dmma_occupancy.cu -o dmma_occupancy
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, int EXTRA>
__global__ void k_dmma(double *out, int iters, int *sink) {
    double a = threadIdx.x * 1e-9 + 1.0, b = 0.999999;
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; }
    int x = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int e = 0; e < EXTRA; ++e) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(it), "r"(e));
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (x == 0x7fffffff) *sink = x;
}

template <int NACC, int EXTRA>
void run(double *out, int *sink, int warps_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
    const int ctas_per_sm = warps_per_sm * 32 / threads;
    const int blocks = 148 * ctas_per_sm;
    const int iters = 4000;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_dmma<NACC, EXTRA><<<blocks, threads>>>(out, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double fl = 2.0 * 8 * 8 * 4 * NACC * iters * (double)blocks * (threads / 32);
    printf("warps/SM %2d  acc/warp %2d  extra int instr per DMMA %d : %7.3f ms  %6.2f TFLOP/s\n", warps_per_sm, NACC, EXTRA, best,
           fl / (best * 1e-3) * 1e-12);
}

int main() {
    double *out; int *sink;
    cudaMalloc(&out, 148 * 64 * 32 * sizeof(double));
    cudaMalloc(&sink, 4);
    for (int w : {4, 8, 16, 32, 64}) run<16, 0>(out, sink, w);
    for (int w : {4, 8, 16, 32}) run<8, 0>(out, sink, w);
    for (int w : {8, 16, 32}) run<4, 0>(out, sink, w);
    for (int w : {8, 16, 32}) run<16, 2>(out, sink, w);
    for (int w : {8, 16, 32}) run<16, 4>(out, sink, w);
    for (int w : {8, 16, 32}) run<16, 6>(out, sink, w);
    return 0;
}

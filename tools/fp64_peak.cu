// FP64 peak microbenchmarks for the roofline denominators (MEASURED_PEAKS.json has no FP64 figure):
// DFMA chains on the FP64 pipe and DMMA (mma.sync m8n8k4 / m16n8k8 f64) chains.  Prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double ra = a + threadIdx.x * 1e-9, rb = b;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(ra), "d"(rb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(ra), "d"(rb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(ra), "d"(rb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(ra), "d"(rb));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
    double ra[4] = {a, a + 1e-9 * threadIdx.x, a, a}, rb[2] = {b, b};
    for (int i = 0; i < iters; ++i) {
#define MMA(c) asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(ra[0]), "d"(ra[1]), "d"(ra[2]), "d"(ra[3]), "d"(rb[0]), "d"(rb[1]))
        MMA(c0); MMA(c1); MMA(c2); MMA(c3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[2] + c3[3] + c0[3] + c1[2] + c2[1] + c3[0];
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); launch();
    cudaDeviceSynchronize();
    double best = 1e30;
    for (int r = 0; r < 10; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 256, blocks = sms * 8, iters = 1 << 14;
    double* out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    double ms1 = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double dfma = 2.0 * 8 * iters * (double)threads * blocks / (ms1 * 1e-3) / 1e12;
    double ms2 = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double dmma884 = 512.0 * 4 * iters * (double)(threads / 32) * blocks / (ms2 * 1e-3) / 1e12;
    double ms3 = time_ms([&] { dmma1688_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double dmma1688 = 2048.0 * 4 * iters * (double)(threads / 32) * blocks / (ms3 * 1e-3) / 1e12;
    cudaError_t e = cudaGetLastError();
    printf("{\"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, \"dmma_m16n8k8_tflops\": %.3f, \"cuda_error\": \"%s\"}\n",
           sms, dfma, dmma884, dmma1688, cudaGetErrorString(e));
    return 0;
}

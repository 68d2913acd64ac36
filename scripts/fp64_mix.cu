// fp64_mix.cu - throughput of the individual fp64-pipe instructions the force kernels use (developer probe).
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9 + 1.0, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    int cnt = 0;
    for (int it = 0; it < iters; ++it) {
        if (OP == 0) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b); }
        if (OP == 1) { x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b); x4 = __dadd_rn(x4, b); x5 = __dadd_rn(x5, b); x6 = __dadd_rn(x6, b); x7 = __dadd_rn(x7, b); }
        if (OP == 2) { x0 = __dmul_rn(x0, a); x1 = __dmul_rn(x1, a); x2 = __dmul_rn(x2, a); x3 = __dmul_rn(x3, a); x4 = __dmul_rn(x4, a); x5 = __dmul_rn(x5, a); x6 = __dmul_rn(x6, a); x7 = __dmul_rn(x7, a); }
        if (OP == 3) { cnt += (x0 < b) + (x1 < a) + (x2 < b) + (x3 < a) + (x4 < b) + (x5 < a) + (x6 < b) + (x7 < a); x0 = __longlong_as_double(__double_as_longlong(x0) + cnt); }
        if (OP == 4) {
            double y0, y1, y2, y3, y4, y5, y6, y7;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x0)); asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y1) : "d"(x1));
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y2) : "d"(x2)); asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y3) : "d"(x3));
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y4) : "d"(x4)); asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y5) : "d"(x5));
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y6) : "d"(x6)); asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y7) : "d"(x7));
            x0 = y0; x1 = y1; x2 = y2; x3 = y3; x4 = y4; x5 = y5; x6 = y6; x7 = y7;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + cnt;
}

template <int OP>
void run(const char* name, int sms) {
    const int iters = 2048, blocks = 8 * sms, threads = 256;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double ops = 8.0 * iters * (double)blocks * threads;
    printf("%-12s %8.3f ms  %7.2f Gop/s/SM  = %5.1f lanes/clk/SM @1.9GHz\n", name, best, ops / best / 1e6 / sms, ops / best / 1e6 / sms / 1.9);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    run<0>("DFMA", p.multiProcessorCount); run<1>("DADD", p.multiProcessorCount); run<2>("DMUL", p.multiProcessorCount);
    run<3>("DSETP", p.multiProcessorCount); run<4>("MUFU.RCP64H", p.multiProcessorCount);
    return 0;
}

// fp64_peak.cu - measured DFMA throughput / latency of one B200 (roofline denominator for the force kernels).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-9 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int blocks_per_sm, int threads, int sms) {
    const int iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks_per_sm * sms * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<ILP><<<blocks_per_sm * sms, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        dfma_kernel<ILP><<<blocks_per_sm * sms, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * ILP * (double)iters * threads * blocks_per_sm * sms;
    printf("ILP=%d warps/SM=%3d  %8.3f ms  %7.2f TFLOP/s\n", ILP, blocks_per_sm * threads / 32, best, flops / best / 1e9);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s: %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int sms = p.multiProcessorCount;
    // sweep: resident warps per SM x independent chains per thread (how much parallelism the FP64 pipe needs)
    for (int w : {4, 8, 16, 32, 64}) {
        const int threads = w >= 8 ? 256 : 128, blocks = w * 32 / threads;
        run<1>(blocks, threads, sms); run<2>(blocks, threads, sms); run<4>(blocks, threads, sms); run<8>(blocks, threads, sms);
    }
    return 0;
}

// Microbenchmark: issue rate of packed FP32 (FFMA2 / FADD2 / FMUL2, sm_100) against scalar FFMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float *sink, int iters) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2((threadIdx.x + i) * 1e-3f, (threadIdx.x + i) * 2e-3f);
    const float2 a = make_float2(0.999f + blockIdx.x * 1e-9f, 0.998f), b = make_float2(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }  // 2 scalar FFMA
                else if (MODE == 1) acc[i] = __ffma2_rn(acc[i], a, b);                                          // 1 FFMA2
                else if (MODE == 2) acc[i] = __fadd2_rn(acc[i], b);
                else acc[i] = __fmul2_rn(acc[i], a);
            }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    if (s == 12345.678f) sink[0] = s;
}
template <int MODE> double run(float *sink, int sms) {
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(sink, 64);
    cudaEventRecord(e0); k<MODE><<<sms * 8, 256>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return (double)sms * 8 * 256 * iters * 64 * 2 /* float lanes per op pair */ / (ms * 1e-3) / 1e12; // T fp32-ops/s (FMA counted once)
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *sink; cudaMalloc(&sink, 4);
    printf("SMs %d\n", sms);
    printf("scalar FFMA : %.2f T lane-ops/s (x2 = TFLOP/s)\n", run<0>(sink, sms));
    printf("FFMA2       : %.2f T lane-ops/s\n", run<1>(sink, sms));
    printf("FADD2       : %.2f T lane-ops/s\n", run<2>(sink, sms));
    printf("FMUL2       : %.2f T lane-ops/s\n", run<3>(sink, sms));
    return 0;
}

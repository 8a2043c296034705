// Development micro-benchmarks: issue / pipe cost of FFMA, FFMA2, FMUL2 and MUFU.EX2 mixes (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu && tools/microbench2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// NP packed FFMA2 + NS scalar FFMA + NM MUFU per iteration, all independent chains
template <int NP, int NS, int NM, bool MUL>
__global__ void __launch_bounds__(256) mix(float* out, int iters, float a, float b) {
    float2 y[12]; float x[12], z[8];
#pragma unroll
    for (int j = 0; j < 12; ++j) { y[j] = make_float2(0.1f * threadIdx.x + j, 0.2f * j); x[j] = 0.3f * j + threadIdx.x; }
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 0.01f * j;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < NP; ++j) y[j % 12] = (MUL && (j & 1)) ? __fmul2_rn(y[j % 12], a2) : __ffma2_rn(y[j % 12], a2, b2);
#pragma unroll
            for (int j = 0; j < NS; ++j) x[j % 12] = fmaf(x[j % 12], a, b);
#pragma unroll
            for (int j = 0; j < NM; ++j) z[j % 8] = ex2a(-z[j % 8]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) s += y[j].x + y[j].y + x[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += z[j];
    if (s == 123.456f) out[0] = s;
}

template <typename F>
double time_ms(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, ms;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* d; cudaMalloc(&d, 64);
    const int iters = 2048;
    const int wps = 32, blocks = sms * (wps / 8);
    const double warp_iters_per_smsp = (double)(wps / 4) * iters * 4;   // warp-iterations (of the r loop) per scheduler
    const double clk = khz * 1e3;
#define RUN(NP, NS, NM, MUL)                                                                         \
    {                                                                                                \
        double ms = time_ms([&] { mix<NP, NS, NM, MUL><<<blocks, 256>>>(d, iters, .999f, .001f); }); \
        printf("packed %2d scalar %2d mufu %d mul %d : %.2f cycles per group (at %d MHz)\n", NP, NS, NM, (int)MUL, \
               ms * 1e-3 * clk / warp_iters_per_smsp, khz / 1000);                                   \
    }
    RUN(0, 12, 0, false)
    RUN(12, 0, 0, false)
    RUN(12, 0, 0, true)
    RUN(6, 6, 0, false)
    RUN(0, 0, 2, false)
    RUN(0, 8, 2, false)
    RUN(0, 12, 2, false)
    RUN(8, 0, 2, false)
    RUN(10, 0, 2, false)
    RUN(12, 0, 2, false)
    RUN(10, 0, 2, true)
    RUN(10, 2, 2, true)
    RUN(10, 4, 2, true)
    RUN(12, 0, 1, false)
    RUN(12, 0, 4, false)
    return 0;
}

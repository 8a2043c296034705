// Development micro-benchmark: FFMA2/FMUL2 throughput with realistic operand patterns (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE 0: y = y*a+b (shared operands)   1: y[j] = c[j]*e[j&3]+y[j] (distinct regs)   2: y[j] = c[j]*bcast(s[j&3]) + y[j]
// 3: like the pixel loop: u = C*bcast(e); m = R*u + m  (dependent pairs)
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(float* out, int iters, float a, float b) {
    float2 y[8], c[8], e[4]; float sc[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { y[j] = make_float2(0.1f * threadIdx.x + j, 0.2f * j); c[j] = make_float2(a + 1e-4f * j, a - 1e-4f * j); }
#pragma unroll
    for (int j = 0; j < 4; ++j) { e[j] = make_float2(b + 1e-5f * j, b - 1e-5f * j); sc[j] = a + 1e-6f * j; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = __ffma2_rn(y[j], a2, b2);
            } else if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = __ffma2_rn(c[j], e[j & 3], y[j]);
            } else if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = __ffma2_rn(c[j], make_float2(sc[j & 3], sc[j & 3]), y[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 u = __fmul2_rn(c[j], make_float2(sc[(j + r) & 3], sc[(j + r) & 3]));
                    y[j] = __ffma2_rn(e[j], u, y[j]);
                    const float2 v = __fmul2_rn(c[4 + j], make_float2(sc[(j + r + 1) & 3], sc[(j + r + 1) & 3]));
                    y[4 + j] = __ffma2_rn(e[j], v, y[4 + j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[j] = sc[j] * 0.999f + 1e-7f;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += y[j].x + y[j].y;
    if (s == 123.456f) out[0] = s;
}

template <int MODE, int WARPS>
void run(float* d, int sms, int khz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096; float best = 1e30f, ms;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE, WARPS><<<sms, WARPS * 32>>>(d, iters, .999f, .001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // packed instructions per scheduler: WARPS/4 warps x iters x 32
    printf("mode %d warps/SM %2d: %.2f cycles per packed instruction per scheduler\n", MODE, WARPS,
           best * 1e-3 * khz * 1e3 / ((WARPS / 4.0) * iters * 32.0));
}

int main() {
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* d; cudaMalloc(&d, 64);
    run<0, 16>(d, sms, khz); run<1, 16>(d, sms, khz); run<2, 16>(d, sms, khz); run<3, 16>(d, sms, khz);
    run<0, 8>(d, sms, khz); run<1, 8>(d, sms, khz); run<3, 8>(d, sms, khz);
    run<1, 4>(d, sms, khz); run<3, 4>(d, sms, khz);
    return 0;
}

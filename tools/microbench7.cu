// Development micro-benchmark (not part of the product): what does a packed FFMA2 cost on sm_100a
// as a function of its operand forms, and which other instructions issue in its shadow?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench7 tools/microbench7.cu
// Every variant runs 4 warps per scheduler (16 per SM), one CTA per SM, and reports cycles per
// loop trip per scheduler divided by the packed (or scalar) FP32 instructions of a trip.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int NACC = 12;

// MODE  0: acc = a[j] * b[j] + acc            three distinct pairs per instruction
//       1: acc = a[j] * bcast(s[j & 3]) + acc  pair, broadcast scalar (4 different), pair
//       2: acc = a[j] * bcast(s0) + acc        pair, ONE broadcast scalar, pair
//       3: acc = acc * a0 + b0                 one changing pair, two shared pairs
//       4: scalar FFMA, three distinct registers (2 x NACC accumulators)
//       5: scalar FFMA, acc = acc * a0 + b0
//       6: acc = a[j] * acc + bcast(s0)        broadcast in the addend
//       7: u = a[j] * bcast(s[j&3]);  acc = b[j] * u + acc      the pixel loop's pattern (2 packed)
// EXTRA 0: nothing   1: + 4 MUFU.EX2   2: + 8 IADD3 (ALU pipe)   3: + 4 LDS.128 (+ LOP3 to keep them)
//       4: + 8 MOV-like (PRMT)  5: + 4 MUFU + 4 LDS.128 + 4 IADD
template <int MODE, int EXTRA>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float a0f, float b0f, const float* gsrc) {
    __shared__ float4 sm[512];
    sm[threadIdx.x] = make_float4(a0f * threadIdx.x, b0f, a0f, 1.f);
    __syncthreads();
    float2 acc[NACC], a[NACC], b[NACC];
    float s[4];
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
        acc[j] = make_float2(0.1f * threadIdx.x + j, 0.2f * j);
        a[j] = make_float2(a0f + 1e-4f * j, a0f - 1e-4f * j);
        b[j] = make_float2(b0f + 1e-5f * j, b0f - 1e-5f * j);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = a0f + 1e-6f * j;
    const float2 a2 = make_float2(a0f, a0f * 1.0001f), b2 = make_float2(b0f, b0f * 1.0001f);
    float x[4] = {0.1f, 0.2f, 0.3f, 0.4f};
    int n[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    unsigned lx = 0;
    const float4* sp = sm + ((threadIdx.x >> 4) & 1);   // two addresses per warp, like the block table
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            if (MODE == 0) acc[j] = __ffma2_rn(a[j], b[j], acc[j]);
            if (MODE == 1) acc[j] = __ffma2_rn(a[j], make_float2(s[j & 3], s[j & 3]), acc[j]);
            if (MODE == 2) acc[j] = __ffma2_rn(a[j], make_float2(s[0], s[0]), acc[j]);
            if (MODE == 3) acc[j] = __ffma2_rn(acc[j], a2, b2);
            if (MODE == 4) { acc[j].x = fmaf(a[j].x, b[j].x, acc[j].x); acc[j].y = fmaf(a[j].y, b[j].y, acc[j].y); }
            if (MODE == 5) { acc[j].x = fmaf(acc[j].x, a2.x, b2.x); acc[j].y = fmaf(acc[j].y, a2.y, b2.y); }
            if (MODE == 6) acc[j] = __ffma2_rn(a[j], acc[j], make_float2(s[0], s[0]));
            if (MODE == 7) {
                const float2 u = __fmul2_rn(a[j], make_float2(s[j & 3], s[j & 3]));
                acc[j] = __ffma2_rn(b[j], u, acc[j]);
            }
            if (MODE == 8 && (j & 3) == 0) {   // grouped: 4 FMUL2 sharing one scalar, then the 4 dependent FFMA2
                float2 u[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) u[q] = __fmul2_rn(a[j + q], make_float2(s[j >> 2], s[j >> 2]));
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[j + q] = __ffma2_rn(b[j + q], u[q], acc[j + q]);
            }
            if (MODE == 9 && (j & 3) == 0) {   // grouped, and the addend is a shared scalar (class 0 of the pixel loop)
                float2 u[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) u[q] = __fmul2_rn(a[j + q], make_float2(s[j >> 2], s[j >> 2]));
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[j + q] = __ffma2_rn(b[j + q], u[q], make_float2(s[3], s[3]));
            }
            if (EXTRA == 1 || EXTRA == 5) { if (j % 3 == 0) x[j / 3] = ex2a(x[j / 3]); }
            if (EXTRA == 2) { if (j < 8) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[j]) : "r"(i)); }
            if (EXTRA == 5) { if (j % 3 == 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[j / 3]) : "r"(i)); }
            if (EXTRA == 3 || EXTRA == 5) {
                if (j % 3 == 2) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                 : "r"((unsigned)__cvta_generic_to_shared(sp + 32 * (j / 3))));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lx) : "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.w)));
                }
            }
            if (EXTRA == 4) { if (j < 8) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(n[j]) : "r"(i)); }
        }
        if (MODE == 1 || MODE >= 7) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] += 1e-7f;      // the scalars change per trip, like E in the pixel loop
        }
    }
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j) t += acc[j].x + acc[j].y;
    t += x[0] + x[1] + x[2] + x[3];
    int m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) m += n[j];
    if (t == 123.456f || m == 0x7fffffff || lx == 0x12345678u) out[0] = t + m;
}

template <int MODE, int EXTRA>
void run(float* d, int sms, int khz, const char* what) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15; float best = 1e30f, ms;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE, EXTRA><<<sms, 512>>>(d, iters, .999f, .001f, d); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * khz * 1e3 / (4.0 * iters);     // per trip per scheduler (4 warps each)
    const int fp = (MODE >= 7) ? 2 * NACC : ((MODE == 4 || MODE == 5) ? 2 * NACC : NACC);
    printf("mode %d extra %d: %7.2f cycles per trip per scheduler, %5.2f per FP32 instruction  (%s)\n", MODE, EXTRA, cyc, cyc / fp, what);
}

int main() {
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* d; cudaMalloc(&d, 4096);
    run<0, 0>(d, sms, khz, "FFMA2 pair*pair+pair, all distinct");
    run<1, 0>(d, sms, khz, "FFMA2 pair*bcast(4 scalars)+pair");
    run<2, 0>(d, sms, khz, "FFMA2 pair*bcast(one scalar)+pair");
    run<3, 0>(d, sms, khz, "FFMA2 acc*a0+b0");
    run<4, 0>(d, sms, khz, "FFMA distinct regs");
    run<5, 0>(d, sms, khz, "FFMA acc*a0+b0");
    run<6, 0>(d, sms, khz, "FFMA2 pair*acc+bcast");
    run<7, 0>(d, sms, khz, "FMUL2 pair*bcast then FFMA2 pair*u+acc");
    run<8, 0>(d, sms, khz, "grouped: 4 FMUL2 pair*bcast(shared), 4 FFMA2 pair*u+acc");
    run<9, 0>(d, sms, khz, "grouped: 4 FMUL2 pair*bcast(shared), 4 FFMA2 pair*u+bcast");
    run<2, 1>(d, sms, khz, "mode 2 + 4 MUFU");
    run<2, 2>(d, sms, khz, "mode 2 + 8 IADD");
    run<3, 1>(d, sms, khz, "mode 3 + 4 MUFU");
    run<3, 2>(d, sms, khz, "mode 3 + 8 IADD");
    run<9, 1>(d, sms, khz, "mode 9 + 4 MUFU");
    run<0, 1>(d, sms, khz, "mode 0 + 4 MUFU");
    run<0, 2>(d, sms, khz, "mode 0 + 8 IADD");
    run<0, 3>(d, sms, khz, "mode 0 + 4 LDS.128 + 4 LOP3");
    run<0, 4>(d, sms, khz, "mode 0 + 8 PRMT");
    run<0, 5>(d, sms, khz, "mode 0 + 4 MUFU + 4 LDS.128 + 4 LOP3 + 4 IADD");
    run<7, 1>(d, sms, khz, "mode 7 + 4 MUFU");
    run<7, 3>(d, sms, khz, "mode 7 + 4 LDS.128 + 4 LOP3");
    run<7, 5>(d, sms, khz, "mode 7 + 4 MUFU + 4 LDS.128 + 4 LOP3 + 4 IADD");
    run<3, 5>(d, sms, khz, "mode 3 + 4 MUFU + 4 LDS.128 + 4 LOP3 + 4 IADD");
    run<4, 5>(d, sms, khz, "mode 4 + 4 MUFU + 4 LDS.128 + 4 LOP3 + 4 IADD");
    return 0;
}

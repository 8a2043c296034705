// Development micro-benchmark: the factorised row step in registers, with broadcast scalar operands
// (current form) against genuine register pairs (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// FORM 0: pairs are adjacent pixels of one group, E and the row terms are broadcast scalars
// FORM 1: pairs are (group A, group B) at the same offset: E pairs come straight from the two MUFU,
//         row factors and row terms are duplicated pairs held in registers (as if loaded from a table)
template <int FORM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) step_kernel(float* out, int iters, float a, float b) {
    float2 dxa[4], C[4][4], R[2][4], bc[4][2], sa[2], d[4], w[4];
    float Rs[2][4], bcs[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = (float)((threadIdx.x & 7) * 4) + 1.5f - 3.3f * k * b - 30.f;
        dxa[k] = make_float2(v, v + 32.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) C[k][j] = make_float2(1.f + 0.01f * j * b + 0.001f * k, 1.f - 0.01f * j * b);
        bc[k][0] = make_float2(0.01f * b * k, 0.01f * b * k); bc[k][1] = make_float2(-0.1f * a * k, -0.1f * a * k);
        bcs[k][0] = 0.01f * b * k; bcs[k][1] = -0.1f * a * k;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        sa[c] = make_float2(-0.02f * a * (c + 1), -0.02f * a * (c + 1));
#pragma unroll
        for (int j = 0; j < 4; ++j) { R[c][j] = make_float2(1.f + 0.001f * j * a, 1.f + 0.001f * j * a); Rs[c][j] = 1.f + 0.001f * j * a; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { d[j] = make_float2(7.f + 0.1f * j, 7.1f); w[j] = make_float2(-0.5f, -0.4f - 0.01f * j); }
    float2 s0 = make_float2(0.f, 0.f), s1 = s0;
    for (int i = 0; i < iters; ++i) {
        float2 m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = make_float2(6.4f, 6.4f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float2 u[4];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int k = 2 * o + c;
                float2 t, q;
                if (FORM == 0) {
                    t = __ffma2_rn(make_float2(sa[c].x, sa[c].x), dxa[k], make_float2(bcs[k][0], bcs[k][0]));
                    q = __ffma2_rn(dxa[k], t, make_float2(bcs[k][1], bcs[k][1]));
                } else {
                    t = __ffma2_rn(sa[c], dxa[k], bc[k][0]);
                    q = __ffma2_rn(dxa[k], t, bc[k][1]);
                }
                const float eA = ex2a(q.x), eB = ex2a(q.y);
                if (FORM == 0) {
                    const float2 ea = make_float2(eA, eA), eb = make_float2(eB, eB);
                    if (o == 0) { u[0] = __fmul2_rn(C[k][0], ea); u[1] = __fmul2_rn(C[k][1], ea); u[2] = __fmul2_rn(C[k][0], eb); u[3] = __fmul2_rn(C[k][1], eb); }
                    else { u[0] = __ffma2_rn(C[k][0], ea, u[0]); u[1] = __ffma2_rn(C[k][1], ea, u[1]); u[2] = __ffma2_rn(C[k][0], eb, u[2]); u[3] = __ffma2_rn(C[k][1], eb, u[3]); }
                } else {
                    const float2 e = make_float2(eA, eB);
                    if (o == 0) { u[0] = __fmul2_rn(C[k][0], e); u[1] = __fmul2_rn(C[k][1], e); u[2] = __fmul2_rn(C[k][2], e); u[3] = __fmul2_rn(C[k][3], e); }
                    else { u[0] = __ffma2_rn(C[k][0], e, u[0]); u[1] = __ffma2_rn(C[k][1], e, u[1]); u[2] = __ffma2_rn(C[k][2], e, u[2]); u[3] = __ffma2_rn(C[k][3], e, u[3]); }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = __ffma2_rn(R[c][j], u[j], m[j]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float2 ra = __ffma2_rn(w[j], m[j], d[j]), rb = __ffma2_rn(w[2 + j], m[2 + j], d[2 + j]);
            s0 = __ffma2_rn(ra, ra, s0); s1 = __ffma2_rn(rb, rb, s1);
        }
        // keep the row terms changing so nothing is hoisted
#pragma unroll
        for (int k = 0; k < 4; ++k) { bc[k][0].x += 1e-6f; bc[k][0].y += 1e-6f; bcs[k][0] += 1e-6f; }
    }
    if (s0.x + s0.y + s1.x + s1.y == 123.456f) out[0] = s0.x;
}

template <int FORM, int WARPS>
void run(float* d, int sms, int khz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16; float best = 1e30f, ms;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); step_kernel<FORM, WARPS><<<sms, WARPS * 32>>>(d, iters, .999f, .001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("form %d warps/SM %2d: %.1f cycles per row step per scheduler\n", FORM, WARPS, best * 1e-3 * khz * 1e3 / ((WARPS / 4.0) * iters));
}

// FORM 2: one anchor per 2 x 4 pixel block: 4 scalar anchor exponents (2 FFMA + 1 MUFU each), pairs are
// (row i, pixel pair p); C[k][4] lane constants per component, T[c][4] block factors (as if from a table)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) block_kernel(float* out, int iters, float a, float b) {
    float dxa[4], sa[2], bys[4], cys[4];
    float2 C[4][4], T[2][4], d[4], w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dxa[k] = (float)((threadIdx.x & 15) * 4) + 1.5f - 3.3f * k * b - 30.f;
        bys[k] = 0.01f * b * k; cys[k] = -0.1f * a * k;
#pragma unroll
        for (int j = 0; j < 4; ++j) C[k][j] = make_float2(1.f + 0.01f * j * b + 0.001f * k, 1.f - 0.01f * j * b);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        sa[c] = -0.02f * a * (c + 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) T[c][j] = make_float2(1.f + 0.001f * j * a, 1.f - 0.001f * j * a);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { d[j] = make_float2(7.f + 0.1f * j, 7.1f); w[j] = make_float2(-0.5f, -0.4f - 0.01f * j); }
    float2 s0 = make_float2(0.f, 0.f), s1 = s0;
    for (int i = 0; i < iters; ++i) {
        float2 m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = make_float2(6.4f, 6.4f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float2 u[4];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int k = 2 * o + c;
                const float q = fmaf(dxa[k], fmaf(sa[c], dxa[k], bys[k]), cys[k]);
                const float e = ex2a(q);
                const float2 e2 = make_float2(e, e);
#pragma unroll
                for (int p = 0; p < 4; ++p) u[p] = (o == 0) ? __fmul2_rn(C[k][p], e2) : __ffma2_rn(C[k][p], e2, u[p]);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) m[p] = __ffma2_rn(T[c][p], u[p], m[p]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float2 ra = __ffma2_rn(w[j], m[j], d[j]), rb = __ffma2_rn(w[2 + j], m[2 + j], d[2 + j]);
            s0 = __ffma2_rn(ra, ra, s0); s1 = __ffma2_rn(rb, rb, s1);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) bys[k] += 1e-6f;
    }
    if (s0.x + s0.y + s1.x + s1.y == 123.456f) out[0] = s0.x;
}

template <int WARPS>
void run_block(float* d, int sms, int khz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16; float best = 1e30f, ms;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); block_kernel<WARPS><<<sms, WARPS * 32>>>(d, iters, .999f, .001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("form 2 (2x4 blocks) warps/SM %2d: %.1f cycles per row step per scheduler\n", WARPS, best * 1e-3 * khz * 1e3 / ((WARPS / 4.0) * iters));
}

int main() {
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* d; cudaMalloc(&d, 64);
    run<0, 16>(d, sms, khz); run<1, 16>(d, sms, khz); run_block<16>(d, sms, khz); run<0, 8>(d, sms, khz); run_block<8>(d, sms, khz);
    return 0;
}

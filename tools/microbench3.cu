// Development micro-benchmark: per-lane pixel data from TMEM (tcgen05.ld) vs shared memory (LDS.128)
// under the factorised pixel loop's arithmetic (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu && tools/microbench3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                    "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                    "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}

// value of plane p (0 data, 1 weight) at row r, column c of the 64 x 64 stamp
__device__ __forceinline__ float pix(int p, int r, int c) { return p ? -(0.5f + 0.001f * (float)((r * 7 + c * 3) & 63)) : 3.f + 0.01f * (float)((r * 5 + c) & 127); }

// MODE 0: TMEM, 1: shared memory, 2: no loads (constants)
template <int MODE>
__global__ void __launch_bounds__(512, 1) loop_kernel(float* out, int passes, float a, float b) {
    extern __shared__ __align__(128) float smem[];   // [2][64][64] planes
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = lane & 7, g = lane >> 3;
    for (int i = threadIdx.x; i < 2 * 64 * 64; i += blockDim.x) smem[i] = pix(i >> 12, (i >> 6) & 63, i & 63);
    uint32_t tbase = 0;
    if (MODE == 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tbase = tmem_base_s;
        if (warp < 4) {
            // lane's pixels of row step i: 8 data then 8 weights -> columns 16 i .. 16 i + 15 of TMEM lane 32 warp + lane
            for (int i = 0; i < 16; ++i) {
                float v[16];
                const int r = 4 * i + g;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = smem[r * 64 + 4 * c + j]; v[4 + j] = smem[r * 64 + 32 + 4 * c + j];
                    v[8 + j] = smem[4096 + r * 64 + 4 * c + j]; v[12 + j] = smem[4096 + r * 64 + 32 + 4 * c + j];
                }
                tmem_st16(tbase + ((uint32_t)(32 * warp) << 16) + 16 * i, v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (MODE == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    float2 dxa[4], C[4][4];
    float2 sa[2] = {make_float2(-0.15f * a, -0.15f * a), make_float2(-0.02f * a, -0.02f * a)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = (float)(c * 4) + 1.5f - 3.3f * k * b - 30.f;
        dxa[k] = make_float2(v, v + 32.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) C[k][j] = make_float2(1.f + 0.01f * j * b + 0.001f * k, 1.f - 0.01f * j * b);
    }
    const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    float2 s = make_float2(0.f, 0.f);
    for (int p = 0; p < passes; ++p) {
        const float* dp = smem + g * 64 + 4 * c;
#pragma unroll 1
        for (int i = 0; i < 16; ++i) {
            float v[16];
            if (MODE == 0) {
                tmem_ld16(taddr + 16 * i, v);
            } else if (MODE == 1) {
                const float4 dA = *reinterpret_cast<const float4*>(dp), dB = *reinterpret_cast<const float4*>(dp + 32);
                const float4 wA = *reinterpret_cast<const float4*>(dp + 4096), wB = *reinterpret_cast<const float4*>(dp + 4096 + 32);
                v[0] = dA.x; v[1] = dA.y; v[2] = dA.z; v[3] = dA.w; v[4] = dB.x; v[5] = dB.y; v[6] = dB.z; v[7] = dB.w;
                v[8] = wA.x; v[9] = wA.y; v[10] = wA.z; v[11] = wA.w; v[12] = wB.x; v[13] = wB.y; v[14] = wB.z; v[15] = wB.w;
                dp += 4 * 64;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = j < 8 ? 3.f + 0.01f * j : -0.5f;
            }
            const float fr = (float)(4 * i + g);
            float2 m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = make_float2(6.4f, 6.4f);
#pragma unroll
            for (int cl = 0; cl < 2; ++cl) {
                const float h = 0.001f * b * (fr - 30.f) * (cl + 1);
                const float2 R01 = make_float2(1.f - 3.f * h, 1.f - h), R23 = make_float2(1.f + h, 1.f + 3.f * h);
                float2 u[4];
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int k = 2 * o + cl;
                    const float yd = fr - 30.f - 2.2f * k;
                    const float bb = 0.01f * b * yd, cc = (sa[cl].x * yd) * yd;
                    const float2 t = __ffma2_rn(sa[cl], dxa[k], make_float2(bb, bb));
                    const float2 q = __ffma2_rn(dxa[k], t, make_float2(cc, cc));
                    const float eA = ex2a(q.x), eB = ex2a(q.y);
                    const float2 ea = make_float2(eA, eA), eb = make_float2(eB, eB);
                    if (o == 0) {
                        u[0] = __fmul2_rn(C[k][0], ea); u[1] = __fmul2_rn(C[k][1], ea);
                        u[2] = __fmul2_rn(C[k][2], eb); u[3] = __fmul2_rn(C[k][3], eb);
                    } else {
                        u[0] = __ffma2_rn(C[k][0], ea, u[0]); u[1] = __ffma2_rn(C[k][1], ea, u[1]);
                        u[2] = __ffma2_rn(C[k][2], eb, u[2]); u[3] = __ffma2_rn(C[k][3], eb, u[3]);
                    }
                }
                m[0] = __ffma2_rn(R01, u[0], m[0]); m[1] = __ffma2_rn(R23, u[1], m[1]);
                m[2] = __ffma2_rn(R01, u[2], m[2]); m[3] = __ffma2_rn(R23, u[3], m[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 r = __ffma2_rn(make_float2(v[8 + 2 * j], v[9 + 2 * j]), m[j], make_float2(v[2 * j], v[2 * j + 1]));
                s = __ffma2_rn(r, r, s);
            }
        }
    }
    // every lane's sum (TMEM and shared-memory modes must agree bit for bit)
    float tot = s.x + s.y;
    for (int off = 16; off; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    if (lane == 0 && blockIdx.x == 0) out[warp] = tot;
    __syncthreads();
    if (MODE == 0 && warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(256));
    }
}

template <int MODE>
void run(const char* name, float* d, int sms, int khz) {
    cudaFuncSetAttribute(loop_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 64 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int passes = 2000;
    float best = 1e30f, ms;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        loop_kernel<MODE><<<sms, 512, 2 * 64 * 64 * 4>>>(d, passes, .999f, .001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    float h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // 4 warps per scheduler, passes x 16 row steps each
    printf("%-8s %.3f ms  %.1f cycles per row step per scheduler  (sums %.6e %.6e %.6e)\n", name, best,
           best * 1e-3 * khz * 1e3 / (4.0 * passes * 16), h[0], h[5], h[15]);
}

int main() {
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* d; cudaMalloc(&d, 64 * 4);
    run<2>("none", d, sms, khz);
    run<1>("smem", d, sms, khz);
    run<0>("tmem", d, sms, khz);
    return 0;
}

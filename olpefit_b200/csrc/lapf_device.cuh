// Device-side building blocks of the LAPF step-2 hot path for sm_100a.
//
//   Layout<NB>      parameter-vector index map (apf_step2.py:108; 3body/apf_step2_3body.py:266-288)
//   Coef<NB>        per-proposal coefficients of the K = 2*NB elliptical Gaussians
//   warp_chi2<>     fused model / residual / square / reduce over one stamp by ONE warp
//                   (replaces build_analytical_model + chi_squared, apf_step2.py:78-137)
//   philox / draws  counter-based random stream (replaces numpy's global MT, apf_step2.py:64,68,143,302)
//
// Numerics: pixel terms in FP32 with stamp-local coordinates, ex2.approx
// with coefficients pre-scaled by -log2(e); per-row FP32 partial sums are folded into an FP64
// accumulator; everything that crosses lanes, the stored chi-square and the Metropolis
// difference are FP64.  The pixel->lane map and the reduction tree are fixed, so chi-square is
// a pure function of the parameter vector.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lapf {

constexpr float kLog2e = 1.4426950408889634f;
constexpr unsigned kFull = 0xffffffffu;

template <int NB>
struct Layout {
    static constexpr int NOBJ = NB;
    static constexpr int P = 3 * NB + 10;   // 16 / 19
    static constexpr int K = 2 * NB;        // Gaussian components: (narrow, wide) per object
    static constexpr int I_DX = 2 * NB, I_DY = 2 * NB + 1;
    static constexpr int I_AMP = 2 * NB + 2;
    static constexpr int I_RATIO = 3 * NB + 2, I_BKGD = 3 * NB + 3;
    static constexpr int I_SX = 3 * NB + 4, I_SY = 3 * NB + 5, I_SX2 = 3 * NB + 6, I_SY2 = 3 * NB + 7;
    static constexpr int I_TH = 3 * NB + 8, I_TH2 = 3 * NB + 9;
};

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU.EX2
    return y;
}

__device__ __forceinline__ double shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(kFull, lo, src);
    hi = __shfl_sync(kFull, hi, src);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(kFull, lo, mask);
    hi = __shfl_xor_sync(kFull, hi, mask);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += shfl_xor_f64(v, off);
    return v;   // identical in every lane: xor-butterfly adds the same pairs everywhere
}

// mbarrier + 1-D TMA bulk copy (global -> shared), used to stage a stamp once per frame.
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Coefficients
// ---------------------------------------------------------------------------------------------
template <int NB>
struct Coef {
    static constexpr int K = 2 * NB;
    float x0[K], y0[K], amp[K];   // component 2o = narrow core of object o, 2o+1 = its wide wing
    float sa[2], sb[2], sc[2];    // shape 0 = narrow, 1 = wide; a, b, c of A.1 times -log2(e)
    float floor;
};

// a, b, c of astropy Gaussian2D.evaluate, pre-scaled so that the
// exponential is a bare ex2:  G = A * 2^(sa*dx^2 + sb*dx*dy + sc*dy^2).
template <int NB>
__device__ __forceinline__ void set_shape(Coef<NB>& cf, int which, float sx, float sy, float th) {
    float s, c;
    sincosf(th, &s, &c);
    const float ivx = 1.0f / (sx * sx), ivy = 1.0f / (sy * sy);
    cf.sa[which] = -0.5f * kLog2e * (c * c * ivx + s * s * ivy);
    cf.sb[which] = -kLog2e * (s * c) * (ivx - ivy);     // sin(2t)/2 = s*c
    cf.sc[which] = -0.5f * kLog2e * (s * s * ivx + c * c * ivy);
}

// Centres (stamp-local) and amplitudes from a parameter vector in frame coordinates
// (build_2d_gaussian, apf_step2.py:95-101).  `pv` may point to shared or global memory.
template <int NB>
__device__ __forceinline__ void set_centres_amps(Coef<NB>& cf, const double* pv, int ox, int oy,
                                                 int floor_index) {
    using L = Layout<NB>;
    const double dx = pv[L::I_DX], dy = pv[L::I_DY];
    const float ratio = (float)pv[L::I_RATIO], bkgd = (float)pv[L::I_BKGD];
#pragma unroll
    for (int o = 0; o < NB; ++o) {
        const double xc = pv[2 * o] - (double)ox, yc = pv[2 * o + 1] - (double)oy;
        cf.x0[2 * o] = (float)xc;
        cf.y0[2 * o] = (float)yc;
        cf.x0[2 * o + 1] = (float)(xc + dx);
        cf.y0[2 * o + 1] = (float)(yc + dy);
        const float a = (float)pv[L::I_AMP + o] - bkgd;   // :95
        const float aw = a * ratio;                       // :96
        cf.amp[2 * o] = a - aw;                           // :97
        cf.amp[2 * o + 1] = aw;
    }
    cf.floor = (float)pv[floor_index];                    // apf_step2.py:119-120
}

template <int NB>
__device__ __forceinline__ void set_all(Coef<NB>& cf, const double* pv, int ox, int oy, int floor_index) {
    using L = Layout<NB>;
    set_centres_amps<NB>(cf, pv, ox, oy, floor_index);
    set_shape<NB>(cf, 0, (float)pv[L::I_SX], (float)pv[L::I_SY], (float)pv[L::I_TH]);
    set_shape<NB>(cf, 1, (float)pv[L::I_SX2], (float)pv[L::I_SY2], (float)pv[L::I_TH2]);
}

// ---------------------------------------------------------------------------------------------
// The pixel loop: one warp evaluates chi-square of one parameter vector over an NY x NX stamp.
//
// Lane geometry.  Columns are processed in panels of PW = min(NX, 64).  Inside a panel a lane
// owns 8 fixed columns (two groups of 4 for 128-bit loads) and every RG-th row, so
//   * the column offsets dx[k][j] = x_j - x0_k live in registers for the whole panel
//     (no per-pixel subtract),
//   * the row-dependent terms  sb*dy, sc*dy^2  are computed once per row and shared by 8 pixels,
//   * per pixel and component the work is 2 FFMA + 1 MUFU.EX2 + 1 FFMA.
// Each quarter-warp reads 128 contiguous bytes per LDS.128: conflict-free for NX = 64/128; for
// NX = 32 odd row groups take their two column groups in swapped order, which keeps the two rows
// a quarter-warp touches on disjoint banks.
// ---------------------------------------------------------------------------------------------
template <int NX>
struct Geo {
    static constexpr int PW = NX >= 64 ? 64 : 32;
    static constexpr int PANELS = NX / PW;
    static constexpr int LPR = PW / 8;        // lanes per row
    static constexpr int RG = 32 / LPR;       // rows handled by a warp per step
    static_assert(NX % PW == 0 && (NX == 32 || NX % 64 == 0), "unsupported stamp width");
};

template <int NB, int NX, int NY, bool STORE>
__device__ __forceinline__ double warp_chi2(const Coef<NB>& cf, const float* __restrict__ d,
                                            const float* __restrict__ w, float* __restrict__ model_out,
                                            int lane) {
    using G = Geo<NX>;
    constexpr int K = 2 * NB;
    static_assert(NY % G::RG == 0, "unsupported stamp height");
    const int c = lane % G::LPR, g = lane / G::LPR;
    const int swap = (G::PW == 32) ? (g & 1) : 0;
    double acc = 0.0;
#pragma unroll 1
    for (int pan = 0; pan < G::PANELS; ++pan) {
        const int colA = pan * G::PW + 4 * c + (G::PW / 2) * swap;
        const int colB = pan * G::PW + 4 * c + (G::PW / 2) * (1 - swap);
        float xd[K][8];
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                xd[k][j] = (float)(colA + j) - cf.x0[k];
                xd[k][4 + j] = (float)(colB + j) - cf.x0[k];
            }
        }
#pragma unroll 1
        for (int i = 0; i < NY / G::RG; ++i) {
            const int r = i * G::RG + g;
            const float fr = (float)r;
            float by[K], cy[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float yd = fr - cf.y0[k];
                by[k] = cf.sb[k & 1] * yd;
                cy[k] = (cf.sc[k & 1] * yd) * yd;
            }
            const float4 dA = *reinterpret_cast<const float4*>(d + r * NX + colA);
            const float4 dB = *reinterpret_cast<const float4*>(d + r * NX + colB);
            const float4 wA = *reinterpret_cast<const float4*>(w + r * NX + colA);
            const float4 wB = *reinterpret_cast<const float4*>(w + r * NX + colB);
            float m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = cf.floor;
#pragma unroll
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t = fmaf(cf.sa[k & 1], xd[k][j], by[k]);
                    const float q = fmaf(xd[k][j], t, cy[k]);
                    m[j] = fmaf(cf.amp[k], ex2_approx(q), m[j]);
                }
            }
            if (STORE) {
                *reinterpret_cast<float4*>(model_out + r * NX + colA) = make_float4(m[0], m[1], m[2], m[3]);
                *reinterpret_cast<float4*>(model_out + r * NX + colB) = make_float4(m[4], m[5], m[6], m[7]);
            }
            const float dv[8] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w};
            const float wv[8] = {wA.x, wA.y, wA.z, wA.w, wB.x, wB.y, wB.z, wB.w};
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float r0 = dv[j] - m[j], r1 = dv[4 + j] - m[4 + j];
                s0 = fmaf(wv[j] * r0, r0, s0);
                s1 = fmaf(wv[4 + j] * r1, r1, s1);
            }
            acc += (double)(s0 + s1);
        }
    }
    return warp_sum_f64(acc);
}

// ---------------------------------------------------------------------------------------------
// Random stream
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kPhiloxTag = 0x4C415046u;   // 'LAPF'

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += kPhiloxW0;
        k.y += kPhiloxW1;
    }
    return c;
}

// The draws of update t of one walker, in the reference's order: parameter index
// (apf_step2.py:302), one standard normal (:64/:68), one uniform (:143).
struct Draw {
    int k;        // parameter index
    double z;     // standard normal
    double lnu;   // log of the uniform (-inf when the uniform is 0)
};

__device__ __forceinline__ Draw make_draw(uint64_t seed, uint64_t walker_id, uint64_t t, int nparam) {
    const uint4 r = philox4x32_10(
        make_uint4((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)(seed >> 32), kPhiloxTag),
        make_uint2((uint32_t)seed, (uint32_t)walker_id));
    Draw d;
    d.k = (int)__umulhi(r.x, (uint32_t)nparam);
    const double u1 = (double)((r.y >> 8) + 1u) * 0x1p-24;   // (0, 1]
    const double u2 = (double)(r.z >> 8) * 0x1p-24;          // [0, 1)
    d.z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    const double u = (double)(r.w >> 8) * 0x1p-24;           // [0, 1)
    d.lnu = log(u);
    return d;
}

}  // namespace lapf

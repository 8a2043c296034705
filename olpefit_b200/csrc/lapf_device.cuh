// Device-side building blocks of the LAPF step-2 hot path for sm_100a.
//
//   Layout<NB>      parameter-vector index map (apf_step2.py:108; 3body/apf_step2_3body.py:266-288)
//   Coef<NB>        per-proposal coefficients of the K = 2*NB elliptical Gaussians, culling
//                   segments, safe-range flag; CoefImg its shared-memory image
//   warp_chi2<>     fused model / residual / square / reduce over one stamp by ONE warp
//                   (replaces build_analytical_model + chi_squared, apf_step2.py:78-137):
//                   row table + column table -> factorised loop (row_steps_fast: one exponential
//                   per 4-pixel group and component) or plain loop (row_steps: one per pixel)
//   tmem_*          tensor memory as a per-lane pixel store (tcgen05.alloc / st / ld)
//   philox / draws  counter-based random stream (replaces numpy's global MT, apf_step2.py:64,68,143,302)
//
// Numerics: pixel terms in FP32 with stamp-local coordinates, ex2.approx
// with coefficients pre-scaled by -log2(e); FP32 partial sums of a few rows are folded into an FP64
// accumulator; everything that crosses lanes, the stored chi-square and the Metropolis
// difference are FP64.  The pixel->lane map and the reduction tree are fixed, so chi-square is
// a pure function of the parameter vector.
#pragma once
#ifndef LAPF_LOOP_UNROLL
#define LAPF_LOOP_UNROLL 2   /* row steps per trip of the factorised loop (measured: 2 is 2 % faster than 1 or 4) */
#endif
#include <cstdint>
#include <cuda_runtime.h>

namespace lapf {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kLoopUnroll = LAPF_LOOP_UNROLL;
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

// Tensor memory (TMEM) as a per-lane, read-only pixel store.  The sampler never uses the tensor
// cores, so their 256 KB of TMEM per SM are free: every lane keeps ITS pixels of the staged stamp
// (prepared data and weight of the 8 columns x NY/RG rows it always works on) in the TMEM lane it
// is allowed to read -- lane 32 (warp % 4) + lane id, shape .32x32b -- and fetches the 16 values
// of a row step with one tcgen05.ld instead of four LDS.128.  That takes 2 KB per warp and row
// step off the shared-memory pipe, which the factorised loop would otherwise saturate.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {   // one warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {          // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t addr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
                   "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                   "r"(__float_as_uint(v[15]))
                 : "memory");
}
// issue the load of 16 consecutive columns; the registers are valid only after tmem_ld16_wait
__device__ __forceinline__ void tmem_ld16_issue(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_ld8_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
// the operands tie every later use of the registers to the wait
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
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
    // far-field culling (set_cull -> set_segments): per column panel the row steps
    //   [0, seg0) floor only, [seg0, seg1) wide wings, [seg1, seg2) all, [seg2, seg3) wings, rest floor only
    // and the component evaluations (pixels x components) they add up to
    int seg[2][4];
    unsigned nexp;
    // the factorised pixel loop (row_steps_fast) is safe for this vector (set_fast)
    bool fast;
};

// a, b, c of astropy Gaussian2D.evaluate, pre-scaled so that the
// exponential is a bare ex2:  G = A * 2^(sa*dx^2 + sb*dx*dy + sc*dy^2).
template <int NB>
__device__ __forceinline__ void set_shape_sc(Coef<NB>& cf, int which, float sx, float sy, float s, float c) {
    const float ivx = 1.0f / (sx * sx), ivy = 1.0f / (sy * sy);
    cf.sa[which] = -0.5f * kLog2e * (c * c * ivx + s * s * ivy);
    cf.sb[which] = -kLog2e * (s * c) * (ivx - ivy);     // sin(2t)/2 = s*c
    cf.sc[which] = -0.5f * kLog2e * (s * s * ivx + c * c * ivy);
}

template <int NB>
__device__ __forceinline__ void set_shape(Coef<NB>& cf, int which, float sx, float sy, float th) {
    float s, c;
    sincosf(th, &s, &c);
    set_shape_sc<NB>(cf, which, sx, sy, s, c);
}

// Per-warp staging of one (trial) parameter vector as FP32 in shared memory.  Lane j < P holds
// parameter j in FP64 frame coordinates; positions are converted to stamp-local coordinates in
// FP64 before rounding to FP32.  Layout of tf[]:
//   tf[j], j < P            parameter j (positions: local)
//   tf[P + j], j < 2*NB     centre coordinate j of the WIDE component: position j + dx or dy
// Every conversion is done once, by the owning lane; afterwards all lanes read tf[] as broadcasts.
template <int NB>
__device__ __forceinline__ void stage_trial(float* tf, int lane, double mine, double oxd, double oyd) {
    using L = Layout<NB>;
    const double dxv = shfl_f64(mine, L::I_DX), dyv = shfl_f64(mine, L::I_DY);
    __syncwarp();   // readers of the previous vector are done
    if (lane < 2 * NB) {
        const double base = mine - ((lane & 1) ? oyd : oxd);
        tf[lane] = (float)base;
        tf[L::P + lane] = (float)(base + ((lane & 1) ? dyv : dxv));
    } else if (lane < L::P) {
        tf[lane] = (float)mine;
    }
    __syncwarp();
}

// Centres and amplitudes from the staged vector (build_2d_gaussian, apf_step2.py:95-101).
template <int NB>
__device__ __forceinline__ void load_centres_amps(Coef<NB>& cf, const float* tf) {
    using L = Layout<NB>;
    const float ratio = tf[L::I_RATIO], bkgd = tf[L::I_BKGD];
#pragma unroll
    for (int o = 0; o < NB; ++o) {
        cf.x0[2 * o] = tf[2 * o];
        cf.y0[2 * o] = tf[2 * o + 1];
        cf.x0[2 * o + 1] = tf[L::P + 2 * o];
        cf.y0[2 * o + 1] = tf[L::P + 2 * o + 1];
        const float a = tf[L::I_AMP + o] - bkgd;          // :95
        const float aw = a * ratio;                       // :96
        cf.amp[2 * o] = a - aw;                           // :97
        cf.amp[2 * o + 1] = aw;
    }
}
// the floor of the model is parameter floor_index of the staged vector (apf_step2.py:119-120)

template <int NB>
__device__ __forceinline__ void load_shape(Coef<NB>& cf, int which, const float* tf) {
    using L = Layout<NB>;
    if (which)
        set_shape<NB>(cf, 1, tf[L::I_SX2], tf[L::I_SY2], tf[L::I_TH2]);
    else
        set_shape<NB>(cf, 0, tf[L::I_SX], tf[L::I_SY], tf[L::I_TH]);
}

// Shared-memory image of a Coef: 16-byte slots, an ODD number of them per walker, so that both the
// one-walker-per-lane accesses of the batched sampler and its broadcast reads are conflict-free
// 128-bit transactions.
template <int NB, int PANELS>
struct CoefImg {
    static constexpr int K = 2 * NB;
    static constexpr int WORDS = 3 * K + 7 + 4 * PANELS + 2;   // x0 y0 amp | sa sb sc floor | segments | nexp fast
    static constexpr int V4 = ((WORDS + 3) / 4) | 1;
    static constexpr int STRIDE = 4 * V4;             // floats per walker
};

template <int NB, int PANELS>
__device__ __forceinline__ void store_coef(float* __restrict__ dst, const Coef<NB>& cf) {
    constexpr int K = 2 * NB;
    using I = CoefImg<NB, PANELS>;
    float v[I::STRIDE];
#pragma unroll
    for (int i = 0; i < I::STRIDE; ++i) v[i] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { v[k] = cf.x0[k]; v[K + k] = cf.y0[k]; v[2 * K + k] = cf.amp[k]; }
#pragma unroll
    for (int c = 0; c < 2; ++c) { v[3 * K + c] = cf.sa[c]; v[3 * K + 2 + c] = cf.sb[c]; v[3 * K + 4 + c] = cf.sc[c]; }
    v[3 * K + 6] = cf.floor;
#pragma unroll
    for (int p = 0; p < PANELS; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[3 * K + 7 + 4 * p + q] = __int_as_float(cf.seg[p][q]);
    v[3 * K + 7 + 4 * PANELS] = __uint_as_float(cf.nexp);
    v[3 * K + 8 + 4 * PANELS] = __int_as_float(cf.fast ? 1 : 0);
#pragma unroll
    for (int q = 0; q < I::V4; ++q)
        reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

template <int NB, int PANELS>
__device__ __forceinline__ void load_coef(Coef<NB>& cf, const float* __restrict__ src) {
    constexpr int K = 2 * NB;
    using I = CoefImg<NB, PANELS>;
    float v[I::STRIDE];
#pragma unroll
    for (int q = 0; q < I::V4; ++q) {
        const float4 t = reinterpret_cast<const float4*>(src)[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) { cf.x0[k] = v[k]; cf.y0[k] = v[K + k]; cf.amp[k] = v[2 * K + k]; }
#pragma unroll
    for (int c = 0; c < 2; ++c) { cf.sa[c] = v[3 * K + c]; cf.sb[c] = v[3 * K + 2 + c]; cf.sc[c] = v[3 * K + 4 + c]; }
    cf.floor = v[3 * K + 6];
#pragma unroll
    for (int p = 0; p < PANELS; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) cf.seg[p][q] = __float_as_int(v[3 * K + 7 + 4 * p + q]);
    cf.nexp = __float_as_uint(v[3 * K + 7 + 4 * PANELS]);
    cf.fast = __float_as_int(v[3 * K + 8 + 4 * PANELS]) != 0;
}

// Coefficients of a parameter vector given in FP64 frame coordinates, by ONE thread: the same
// conversions as stage_trial + load_centres_amps + load_shape, so the numbers are the ones the
// cooperative path produces.  v[] must be indexed with constants only (registers).
template <int NB>
__device__ __forceinline__ void coef_from_vector(Coef<NB>& cf, const double (&v)[Layout<NB>::P], double floor_v,
                                                 double oxd, double oyd) {
    using L = Layout<NB>;
    float tf[L::P + 2 * NB];
#pragma unroll
    for (int j = 0; j < L::P; ++j) {
        if (j < 2 * NB) {
            const double base = v[j] - ((j & 1) ? oyd : oxd);
            tf[j] = (float)base;
            tf[L::P + j] = (float)(base + ((j & 1) ? v[L::I_DY] : v[L::I_DX]));
        } else {
            tf[j] = (float)v[j];
        }
    }
    load_centres_amps<NB>(cf, tf);
    cf.floor = (float)floor_v;
    set_shape<NB>(cf, 0, tf[L::I_SX], tf[L::I_SY], tf[L::I_TH]);
    set_shape<NB>(cf, 1, tf[L::I_SX2], tf[L::I_SY2], tf[L::I_TH2]);
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

// Row table: everything that depends on the row only, computed ONCE per proposal by the warp
// (one or two rows per lane) instead of once per row step by every lane.  One entry of Tab::RS
// floats per row r:
//   [2k], [2k+1]          sb_k * dy_k,  sc_k * dy_k^2     (dy_k = r - y0_k), for component k < K
//   [2K + 4c .. +3]       2^(j h)  for j = -3, -1, 1, 3,  h  = sb_c (r - y0_c) / 2,   shape class c
//   [2K + 8 + 4c .. +3]   2^(j h') for j = -3, -1, 1, 3,  h' = h + sa_c * D
// The first pair gives the exponent of component k at any column of the row:
//   q = fma(dx, fma(sa, dx, sb*dy), sc*dy^2);
// the second the factors that move it half a pixel / one and a half pixels left or right of the
// anchor column of a lane's column group A, the third the same for its group B, whose anchor is
// D = +-PW/2 columns away (see row_steps_fast; the sign follows the row for 32-pixel panels, where
// odd rows take their groups in swapped order).  y0_c is the centre of the class's FIRST component;
// the other components of the class fold their offset into the lane constants.
// The entry is padded to 28 floats so the 4 (or 8) row groups of a warp read distinct banks.
template <int NB>
struct Tab {
    static constexpr int K = 2 * NB;
    static constexpr int RS = 28;
    static constexpr int OFF_RA = 2 * K;
    static constexpr int OFF_RB = 2 * K + 8;
    static_assert(OFF_RB + 8 <= RS && OFF_RA % 4 == 0, "row table entry layout");
};

// A table covers TR rows of the stamp; taller stamps are walked table by table ("half").  A team
// member (TEAM warps share one walker, warp tw takes every TEAM-th row step) only builds the
// TR / TEAM rows it evaluates.
template <int NY, int TEAM = 1>
struct Rows {
    static constexpr int TR = (NY >= 128 && TEAM == 1) ? 32 : (NY < 64 ? NY : 64);   // rows per table
    static constexpr int HALVES = NY / TR;
    static constexpr int NR = TR / TEAM;                                             // rows a warp builds
    static_assert(NY % TR == 0 && TR % TEAM == 0, "unsupported stamp height");
};

// per-warp shared-memory scratch of the pixel loop: the row table, then the column table
// (coop_consts: [component][group-A anchor 0..7][4 pixel offsets])
template <int NB, int NY, int TEAM = 1>
struct Scratch {
    static constexpr int TAB = Rows<NY, TEAM>::NR * Tab<NB>::RS;
    static constexpr int CT = 2 * NB * 8 * 4;
    static constexpr int FLOATS = TAB + CT;
};

template <int NB, int NX, int TR, int TEAM>
__device__ __forceinline__ void build_row_table(float* __restrict__ rt, const Coef<NB>& cf, int lane, int row0, int tw) {
    constexpr int K = 2 * NB;
    constexpr int NR = TR / TEAM;
    constexpr int RG = Geo<NX>::RG;
    constexpr float HALF = 0.5f * Geo<NX>::PW;
    using T = Tab<NB>;
    __syncwarp();   // readers of the previous table are done
    if (NR == 64) {
        // two rows (r, r+32) per pass, as the two halves of packed FP32 operations (TEAM = 1 here)
        const float2 fr = make_float2((float)(row0 + lane), (float)(row0 + 32 + lane));
        float4* o0 = reinterpret_cast<float4*>(rt + lane * T::RS);
        float4* o1 = reinterpret_cast<float4*>(rt + (32 + lane) * T::RS);
        float2 v[2 * K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float2 yd = __fadd2_rn(fr, make_float2(-cf.y0[k], -cf.y0[k]));
            const float2 sc2 = make_float2(cf.sc[k & 1], cf.sc[k & 1]);
            v[2 * k] = __fmul2_rn(make_float2(cf.sb[k & 1], cf.sb[k & 1]), yd);
            v[2 * k + 1] = __fmul2_rn(__fmul2_rn(sc2, yd), yd);
        }
#pragma unroll
        for (int q = 0; q < 2 * K / 4; ++q) {
            o0[q] = make_float4(v[4 * q].x, v[4 * q + 1].x, v[4 * q + 2].x, v[4 * q + 3].x);
            o1[q] = make_float4(v[4 * q].y, v[4 * q + 1].y, v[4 * q + 2].y, v[4 * q + 3].y);
        }
        // 64-row tables belong to 64-pixel panels: group B is always +PW/2 away (no swapped rows)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float2 h = __fmul2_rn(v[2 * c], make_float2(0.5f, 0.5f));     // sb_c * dy_c / 2
            const float2 h3 = __fmul2_rn(v[2 * c], make_float2(1.5f, 1.5f));
            o0[T::OFF_RA / 4 + c] = make_float4(ex2_approx(-h3.x), ex2_approx(-h.x), ex2_approx(h.x), ex2_approx(h3.x));
            o1[T::OFF_RA / 4 + c] = make_float4(ex2_approx(-h3.y), ex2_approx(-h.y), ex2_approx(h.y), ex2_approx(h3.y));
            const float sd = cf.sa[c] * HALF;
            const float2 g = __fadd2_rn(h, make_float2(sd, sd));
            const float2 g3 = __fmul2_rn(g, make_float2(3.f, 3.f));
            o0[T::OFF_RB / 4 + c] = make_float4(ex2_approx(-g3.x), ex2_approx(-g.x), ex2_approx(g.x), ex2_approx(g3.x));
            o1[T::OFF_RB / 4 + c] = make_float4(ex2_approx(-g3.y), ex2_approx(-g.y), ex2_approx(g.y), ex2_approx(g3.y));
        }
    } else {
#pragma unroll
        for (int j0 = 0; j0 < NR; j0 += 32) {
            const int j = j0 + lane;                    // local row: row step j / RG of this warp, row group j % RG
            if (NR % 32 == 0 || j < NR) {
                const int r = row0 + ((j / RG) * TEAM + tw) * RG + (j % RG);
                const float fr = (float)r;
                float4* o = reinterpret_cast<float4*>(rt + j * T::RS);
                float v[2 * K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float yd = fr - cf.y0[k];
                    v[2 * k] = cf.sb[k & 1] * yd;
                    v[2 * k + 1] = (cf.sc[k & 1] * yd) * yd;
                }
#pragma unroll
                for (int q = 0; q < 2 * K / 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                // odd rows of a 32-pixel panel take their column groups in swapped order
                const float dlt = (Geo<NX>::PW == 32 && (r & 1)) ? -HALF : HALF;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float h = v[2 * c] * 0.5f, h3 = v[2 * c] * 1.5f;
                    o[T::OFF_RA / 4 + c] = make_float4(ex2_approx(-h3), ex2_approx(-h), ex2_approx(h), ex2_approx(h3));
                    const float g = h + cf.sa[c] * dlt, g3 = g * 3.f;
                    o[T::OFF_RB / 4 + c] = make_float4(ex2_approx(-g3), ex2_approx(-g), ex2_approx(g), ex2_approx(g3));
                }
            }
        }
    }
    __syncwarp();
}

// Far-field culling.  |A_k| 2^(q) <= |A_k| 2^(kappa dy^2) for every pixel of a row at distance dy
// from the component's centre (kappa = sc - sb^2/(4 sa), the exponent maximised over dx), and the
// same with the roles of x and y swapped for a column panel.  A component whose bound over a whole
// row step (or panel) is below tau = 2^-27 |floor| is skipped there: the components of a class are
// added to the pixel as ONE term (see row_steps_fast), and up to four terms below tau sum to less
// than half an ulp of a model value that is at least the floor, so the FMA that would add them
// rounds them away: no MUFU, no FFMA, and (for non-negative amplitudes) bit-identical results.  Rows are
// culled per CLASS (all narrow cores / all wide wings): the active rows of a class are one interval
// of row steps, so a panel is walked as at most five segments (none, wings, all, wings, none), each
// a tight loop without per-row tests.  On a 128-pixel stamp the cores matter in ~1/4 of the rows and
// the wings in ~2/3.  The decision is a pure function of the coefficients (chi-square stays a
// function of the parameter vector); a nan/inf in them disables culling so it reaches chi-square.
template <int NX, int NY>
__device__ __forceinline__ void cull_one(float a, float x0, float y0, float sa, float sb, float sc, float floor_v,
                                         int& lo, int& hi, uint32_t& pm) {
    using G = Geo<NX>;
    constexpr int STEPS = NY / G::RG;
    constexpr uint32_t kAllPans = (1u << G::PANELS) - 1u;
    // approximate log2 / divide / sqrt are fine here: the radius gets a whole pixel of slack
    const float tau = 0x1p-27f * fabsf(floor_v);
    const float L = __log2f(__fdividef(fabsf(a), tau));    // bits of headroom above tau
    lo = 0; hi = STEPS - 1;                                // default: everything (also for nan / inf)
    pm = kAllPans;
    if (L <= 0.f) {                                        // below tau everywhere
        lo = STEPS; hi = -1; pm = 0u;
    } else {
        const float ky = sc - __fdividef(sb * sb, 4.f * sa), kx = sa - __fdividef(sb * sb, 4.f * sc);   // both < 0
        const float Y = __fsqrt_rn(__fdividef(L, -ky)) + 1.f, X = __fsqrt_rn(__fdividef(L, -kx)) + 1.f;
        if (Y < 1e6f && fabsf(y0) < 1e6f) {
            lo = max(0, (int)ceilf((y0 - Y - (float)(G::RG - 1)) / (float)G::RG));
            hi = min(STEPS - 1, (int)floorf((y0 + Y) / (float)G::RG));
            if (lo > hi) { lo = STEPS; hi = -1; }
        }
        if (G::PANELS > 1 && X < 1e6f && fabsf(x0) < 1e6f) {
            pm = 0u;
#pragma unroll
            for (int p = 0; p < G::PANELS; ++p)
                if (x0 + X >= (float)(p * G::PW) && x0 - X <= (float)(p * G::PW + G::PW - 1)) pm |= 1u << p;
        }
    }
}

// From the class intervals to the segments of every panel (the wing interval is widened to contain
// the core interval: evaluating a component where it is not needed is harmless).  The plain loop
// and team members (whose update time is set by the warp with the busiest rows) evaluate densely.
// Needs cf.fast.
template <int NB, int NX, int NY, int TEAM>
__device__ __forceinline__ void set_segments(Coef<NB>& cf, const int (&lo)[2], const int (&hi)[2], const uint32_t (&pm)[2]) {
    using G = Geo<NX>;
    constexpr int STEPS = NY / G::RG;
    unsigned n = 0;
#pragma unroll
    for (int pan = 0; pan < G::PANELS; ++pan) {
        const bool n_on = ((pm[0] >> pan) & 1u) && lo[0] <= hi[0];
        const bool w_on = ((pm[1] >> pan) & 1u) && lo[1] <= hi[1];
        int nlo = n_on ? lo[0] : STEPS, nhi1 = n_on ? hi[0] + 1 : STEPS;
        int wlo = w_on ? lo[1] : nlo, whi1 = w_on ? hi[1] + 1 : nhi1;
        wlo = min(wlo, nlo);
        whi1 = max(whi1, nhi1);
        if (!n_on) { nlo = whi1; nhi1 = whi1; }
        if (!cf.fast || TEAM > 1) { wlo = 0; nlo = 0; nhi1 = STEPS; whi1 = STEPS; }
        cf.seg[pan][0] = wlo; cf.seg[pan][1] = nlo; cf.seg[pan][2] = nhi1; cf.seg[pan][3] = whi1;
        n += (unsigned)((nhi1 - nlo) + (whi1 - wlo));
    }
    cf.nexp = (unsigned)(G::PW * G::RG * NB) * n;
}

// cooperative form: lane k < K works out component k, the hulls are formed with shuffles
template <int NB, int NX, int NY, int TEAM = 1>
__device__ __forceinline__ void set_cull(Coef<NB>& cf, int lane) {
    using G = Geo<NX>;
    constexpr int K = 2 * NB;
    constexpr int STEPS = NY / G::RG;
    float a = cf.amp[0], x0 = cf.x0[0], y0 = cf.y0[0];
#pragma unroll
    for (int k = 1; k < K; ++k)
        if (lane == k) { a = cf.amp[k]; x0 = cf.x0[k]; y0 = cf.y0[k]; }
    const int sh = lane & 1;
    const float sa = sh ? cf.sa[1] : cf.sa[0], sb = sh ? cf.sb[1] : cf.sb[0], sc = sh ? cf.sc[1] : cf.sc[0];
    int lo, hi;
    uint32_t pm;
    cull_one<NX, NY>(a, x0, y0, sa, sb, sc, cf.floor, lo, hi, pm);
    if (lane >= K) { lo = STEPS; hi = -1; pm = 0u; }       // neutral element
    // hull over the components of each class: lanes of equal parity (xor 2, 4 keeps the parity)
#pragma unroll
    for (int off = 2; off <= 4; off <<= 1) {
        lo = min(lo, __shfl_xor_sync(kFull, lo, off));
        hi = max(hi, __shfl_xor_sync(kFull, hi, off));
        pm |= __shfl_xor_sync(kFull, pm, off);
    }
    int clo[2], chi[2];
    uint32_t cpm[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        clo[c] = __shfl_sync(kFull, lo, c);
        chi[c] = __shfl_sync(kFull, hi, c);
        cpm[c] = __shfl_sync(kFull, pm, c);
    }
    set_segments<NB, NX, NY, TEAM>(cf, clo, chi, cpm);
}

// one-thread form (the batched sampler prepares one walker per lane): same decisions
template <int NB, int NX, int NY>
__device__ __forceinline__ void set_cull_serial(Coef<NB>& cf) {
    constexpr int K = 2 * NB;
    constexpr int STEPS = NY / Geo<NX>::RG;
    int clo[2] = {STEPS, STEPS}, chi[2] = {-1, -1};
    uint32_t cpm[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < K; ++k) {
        int lo, hi;
        uint32_t pm;
        cull_one<NX, NY>(cf.amp[k], cf.x0[k], cf.y0[k], cf.sa[k & 1], cf.sb[k & 1], cf.sc[k & 1], cf.floor, lo, hi, pm);
        clo[k & 1] = min(clo[k & 1], lo);
        chi[k & 1] = max(chi[k & 1], hi);
        cpm[k & 1] |= pm;
    }
    set_segments<NB, NX, NY, 1>(cf, clo, chi, cpm);
}

template <int NB, int NX, int NY, int TEAM = 1>
__device__ __forceinline__ void no_cull(Coef<NB>& cf) {
    const int clo[2] = {0, 0}, chi[2] = {NY / Geo<NX>::RG - 1, NY / Geo<NX>::RG - 1};
    const uint32_t cpm[2] = {0xffffffffu, 0xffffffffu};
    set_segments<NB, NX, NY, TEAM>(cf, clo, chi, cpm);
}

// The factorised pixel loop multiplies factors whose exponents can be large although their sum is
// not.  It is used only when, for every pixel of the stamp, the column factor's exponent
// |j (sa (2 dxa + j) + w_k)| and the row factor's |1.5 (sb dy + sa PW)| stay below 40 -- no factor overflows
// or underflows on its own -- and when an anchor exponential that underflows (q0 < -126) implies
// that its four pixels are below 2^-46 |A| <= 2^-25 |floor|.  Anything else, nan and inf included,
// takes the plain loop (one exponential per pixel and component).  The decision is a function of
// the coefficients only and is the same in every lane.
template <int NX, int NY>
__device__ __forceinline__ bool fast_one(float a, float x0, float y0, float yref, float sa, float sb, float sc,
                                         float floor_v) {
    const float dxm = fabsf(x0 - 0.5f * NX) + 0.5f * NX;      // >= |anchor - x0| for every anchor column
    const float argc = 1.5f * (fabsf(sa) * (2.f * dxm + 1.5f) + fabsf(sb * (yref - y0)));
    const float argr = 1.5f * (fabsf(sb) * (fabsf(yref - 0.5f * NY) + 0.5f * NY) + fabsf(sa) * Geo<NX>::PW);
    return argc <= 40.f && argr <= 40.f && fabsf(sc) <= 1e30f &&
           fabsf(a) <= 0x1p21f * fabsf(floor_v) && fabsf(a) <= 1e12f;    // all false on nan
}

template <int NB, int NX, int NY>
__device__ __forceinline__ void set_fast(Coef<NB>& cf, int lane) {
    constexpr int K = 2 * NB;
    float a = cf.amp[0], x0 = cf.x0[0], y0 = cf.y0[0];
#pragma unroll
    for (int k = 1; k < K; ++k)
        if (lane == k) { a = cf.amp[k]; x0 = cf.x0[k]; y0 = cf.y0[k]; }
    const int sh = lane & 1;
    const float sa = sh ? cf.sa[1] : cf.sa[0], sb = sh ? cf.sb[1] : cf.sb[0], sc = sh ? cf.sc[1] : cf.sc[0];
    const bool ok = fast_one<NX, NY>(a, x0, y0, sh ? cf.y0[1] : cf.y0[0], sa, sb, sc, cf.floor);
    cf.fast = __all_sync(kFull, ok || lane >= K);
}

template <int NB, int NX, int NY>
__device__ __forceinline__ void set_fast_serial(Coef<NB>& cf) {
    constexpr int K = 2 * NB;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < K; ++k)
        ok = ok && fast_one<NX, NY>(cf.amp[k], cf.x0[k], cf.y0[k], cf.y0[k & 1], cf.sa[k & 1], cf.sb[k & 1],
                                    cf.sc[k & 1], cf.floor);
    cf.fast = ok;
}

// Row steps [i0, i1) of one panel for the component classes KIND says (0: none, the model is the
// floor; 1: wide wings only; 2: all components).  chi-square terms are added to the four FP32
// accumulators s0, s1 in row order, whatever the segmentation (so culling does not regroup sums).
// PREP = true: the planes hold d*sqrt(w) and -sqrt(w) (the sampler converts a stamp once after
// staging it); PREP = false: raw data / weight planes, converted per pixel.  Both give
// bit-identical chi-square: the residual is always  r = fma(-sqrt(w), m, d*sqrt(w)),  chi2 += r*r.
struct StepPtrs {          // where the next row step of this lane lives
    const float* rp;       // row table
    const float* dp;       // data plane
    const float* wp;       // weight plane
    float* mp;             // model output (STORE only)
    uint32_t tm;           // TMEM address of the lane's 16 values (TM only)
};

// model of 8 pixels -> (optional store) -> residuals -> chi-square accumulators
template <int NX, bool STORE, bool PREP>
__device__ __forceinline__ void finish_step(const float2 (&m)[4], const float4& dA, const float4& dB,
                                            const float4& wA, const float4& wB, float* mp, int colA, int colB,
                                            float2& s0, float2& s1) {
    if (STORE) {
        *reinterpret_cast<float4*>(mp + colA) = make_float4(m[0].x, m[0].y, m[1].x, m[1].y);
        *reinterpret_cast<float4*>(mp + colB) = make_float4(m[2].x, m[2].y, m[3].x, m[3].y);
    }
    float2 dv[4] = {make_float2(dA.x, dA.y), make_float2(dA.z, dA.w), make_float2(dB.x, dB.y),
                    make_float2(dB.z, dB.w)};
    float2 wv[4] = {make_float2(wA.x, wA.y), make_float2(wA.z, wA.w), make_float2(wB.x, wB.y),
                    make_float2(wB.z, wB.w)};
    if (!PREP) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float rx = sqrtf(wv[j].x), ry = sqrtf(wv[j].y);
            dv[j] = make_float2(dv[j].x * rx, dv[j].y * ry);
            wv[j] = make_float2(-rx, -ry);
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float2 ra = __ffma2_rn(wv[j], m[j], dv[j]);
        const float2 rb = __ffma2_rn(wv[2 + j], m[2 + j], dv[2 + j]);
        s0 = __ffma2_rn(ra, ra, s0);
        s1 = __ffma2_rn(rb, rb, s1);
    }
}

// ---- plain loop: one exponential per pixel and component -------------------------------------
// Arithmetic is issued as packed FFMA2 (fma.rn.f32x2, new on sm_100): two adjacent pixels per
// instruction, scalar coefficients as broadcast operands.  Per pixel PAIR and component that is
// 3 FFMA2 + 2 MUFU.EX2: SFU-bound.  Kept for parameter vectors set_fast turns away.
template <int NB, int NX, int NY, bool STORE, bool PREP, int KIND>
__device__ __forceinline__ void row_steps(const Coef<NB>& cf, const float2 (&xd)[2 * NB][4], float2& s0, float2& s1,
                                          int& i, int i1, StepPtrs& sp, int colA, int colB) {
    using G = Geo<NX>;
    using T = Tab<NB>;
    constexpr int K = 2 * NB;
    const float* rp = sp.rp;
    const float* dp = sp.dp;
    const float* wp = sp.wp;
    float* mp = sp.mp;
#pragma unroll 1
    for (; i < i1; ++i) {
        const float4 dA = *reinterpret_cast<const float4*>(dp + colA);
        const float4 dB = *reinterpret_cast<const float4*>(dp + colB);
        const float4 wA = *reinterpret_cast<const float4*>(wp + colA);
        const float4 wB = *reinterpret_cast<const float4*>(wp + colB);
        float2 m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = make_float2(cf.floor, cf.floor);
        if (KIND > 0) {
            float rc[2 * K];
#pragma unroll
            for (int q = 0; q < 2 * K / 4; ++q) {
                const float4 t4 = reinterpret_cast<const float4*>(rp)[q];
                rc[4 * q] = t4.x; rc[4 * q + 1] = t4.y; rc[4 * q + 2] = t4.z; rc[4 * q + 3] = t4.w;
            }
#pragma unroll
            for (int k = (KIND == 2 ? 0 : 1); k < K; k += (KIND == 2 ? 1 : 2)) {
                const float2 sa2 = make_float2(cf.sa[k & 1], cf.sa[k & 1]);
                const float2 by2 = make_float2(rc[2 * k], rc[2 * k]);
                const float2 cy2 = make_float2(rc[2 * k + 1], rc[2 * k + 1]);
                const float2 am2 = make_float2(cf.amp[k], cf.amp[k]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 t = __ffma2_rn(sa2, xd[k][j], by2);
                    const float2 q = __ffma2_rn(xd[k][j], t, cy2);
                    const float2 e = make_float2(ex2_approx(q.x), ex2_approx(q.y));
                    m[j] = __ffma2_rn(am2, e, m[j]);
                }
            }
        }
        finish_step<NX, STORE, PREP>(m, dA, dB, wA, wB, mp, colA, colB, s0, s1);
        if (STORE) mp += G::RG * NX;
        rp += G::RG * T::RS;
        dp += G::RG * NX;
        wp += G::RG * NX;
    }
    sp.rp = rp; sp.dp = dp; sp.wp = wp; sp.mp = mp;
}

// ---- factorised loop -------------------------------------------------------------------------
// A lane's 8 columns are two groups of 4 adjacent pixels.  With the anchor xa at the centre of a
// group (between its 2nd and 3rd pixel) and j in {-1.5, -0.5, 0.5, 1.5},
//     q_k(xa + j, r) = q_k(xa, r) + j (sa (2 dxa_k + j)) + j sb dy_k,        dxa_k = xa - x0_k,
// and dy_k = (r - y0_c) + (y0_c - y0_k) with c the first component of k's shape class, so
//     A_k 2^q_k(xa + j, r) = E_k(r) * C_kj * R_cj(r)
//     E_k(r)  = 2^q_k(xa, r)                         one MUFU.EX2 per GROUP and component
//     C_kj    = A_k 2^(j (sa (2 dxa_k + j) + sb (y0_c - y0_k)))    lane constant (registers)
//     R_cj(r) = 2^(j sb (r - y0_c))                  row table, shared by the class
// and a class adds  R_cj * sum_k C_kj E_k  to the pixel: per pixel pair NB + 1 packed operations
// for NB components, plus 2 FFMA2 + 2 MUFU per component for the two anchors.  The anchor of
// group B is D = +-PW/2 columns from group A's, so C_kj(B) = C_kj(A) 2^(2 j sa D): the lane keeps
// the constants of group A only and group B takes its factor from a second row-table entry
// R'_cj = R_cj 2^(2 j sa D).  A row step of 8 pixels x 4 components is 40 packed FP32
// instructions and 8 MUFU (plain loop: 56 and 32): the loop is bound by the FP32 pipe, not the SFU.
template <int NB>
struct LaneK {
    static constexpr int K = 2 * NB;
    float2 dxa[K];       // (group A, group B) anchor offset to the centre of component k
    float2 C[K][2];      // group A, pixel pairs (-1.5,-0.5) and (0.5,1.5)
};

// The C_kj of the 8 group-A anchors of a panel (columns 4a + 1.5, a = 0..7), worked out once per
// proposal by the warp -- lane (a, k) does component k at anchor a -- and handed round through the
// column table ct[k][a][4]; every lane then reads the 4K values of its own anchor.
template <int NB, int NX>
__device__ __forceinline__ void coop_consts(LaneK<NB>& lk, float* __restrict__ ct, const Coef<NB>& cf, int lane,
                                            int pan, int colA, int colB) {
    using G = Geo<NX>;
    constexpr int K = 2 * NB;
    __syncwarp();   // readers of the previous column table are done
    const int a = lane & 7;
    const float xa = (float)(pan * G::PW + 4 * a) + 1.5f;
#pragma unroll
    for (int kk0 = 0; kk0 < K; kk0 += 4) {
        const int kk = kk0 + (lane >> 3);
        if (K % 4 == 0 || kk < K) {
            float amp = cf.amp[0], x0 = cf.x0[0], y0 = cf.y0[0];
#pragma unroll
            for (int k = 1; k < K; ++k)
                if (kk == k) { amp = cf.amp[k]; x0 = cf.x0[k]; y0 = cf.y0[k]; }
            const int c = kk & 1;
            const float sa = c ? cf.sa[1] : cf.sa[0], sb = c ? cf.sb[1] : cf.sb[0];
            const float wk = sb * ((c ? cf.y0[1] : cf.y0[0]) - y0);      // exactly 0 for the class's first component
            const float d2 = 2.f * (xa - x0);
            const float2 w2 = make_float2(wk, wk), sa2 = make_float2(sa, sa), am = make_float2(amp, amp);
            const float2 jlo = make_float2(-1.5f, -0.5f), jhi = make_float2(0.5f, 1.5f);
            const float2 alo = __fmul2_rn(jlo, __ffma2_rn(sa2, __fadd2_rn(make_float2(d2, d2), jlo), w2));
            const float2 ahi = __fmul2_rn(jhi, __ffma2_rn(sa2, __fadd2_rn(make_float2(d2, d2), jhi), w2));
            const float2 clo = __fmul2_rn(am, make_float2(ex2_approx(alo.x), ex2_approx(alo.y)));
            const float2 chi = __fmul2_rn(am, make_float2(ex2_approx(ahi.x), ex2_approx(ahi.y)));
            reinterpret_cast<float4*>(ct)[kk * 8 + a] = make_float4(clo.x, clo.y, chi.x, chi.y);
        }
    }
    __syncwarp();
    const int mine = ((colA - pan * G::PW) >> 2) & 7;   // this lane's group-A anchor
    const float xaA = (float)colA + 1.5f, xaB = (float)colB + 1.5f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float4 v = reinterpret_cast<const float4*>(ct)[k * 8 + mine];
        lk.C[k][0] = make_float2(v.x, v.y);
        lk.C[k][1] = make_float2(v.z, v.w);
        lk.dxa[k] = make_float2(xaA - cf.x0[k], xaB - cf.x0[k]);
    }
}

template <int NB, int NX, int NY, bool STORE, bool PREP, int KIND, int TM = 0>
__device__ __forceinline__ void row_steps_fast(const Coef<NB>& cf, const LaneK<NB>& lk, float2& s0, float2& s1,
                                               int& i, int i1, StepPtrs& sp, int colA, int colB) {
    using G = Geo<NX>;
    using T = Tab<NB>;
    constexpr int K = 2 * NB;
    const float* rp = sp.rp;
    const float* dp = sp.dp;
    const float* wp = sp.wp;
    float* mp = sp.mp;
    uint32_t tm = sp.tm;
#pragma unroll kLoopUnroll
    for (; i < i1; ++i) {
        float4 dA, dB, wA, wB;
        uint32_t tv[16];
        if (TM == 1) {
            tmem_ld16_issue(tm, tv);
        } else if (TM == 2) {
            tmem_ld8_issue(tm, tv);
            dA = *reinterpret_cast<const float4*>(dp + colA);
            dB = *reinterpret_cast<const float4*>(dp + colB);
        } else {
            dA = *reinterpret_cast<const float4*>(dp + colA);
            dB = *reinterpret_cast<const float4*>(dp + colB);
            wA = *reinterpret_cast<const float4*>(wp + colA);
            wB = *reinterpret_cast<const float4*>(wp + colB);
        }
        float2 m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = make_float2(cf.floor, cf.floor);
        if (KIND > 0) {
            float rc[2 * K];
#pragma unroll
            for (int q = 0; q < 2 * K / 4; ++q) {
                const float4 t4 = reinterpret_cast<const float4*>(rp)[q];
                rc[4 * q] = t4.x; rc[4 * q + 1] = t4.y; rc[4 * q + 2] = t4.z; rc[4 * q + 3] = t4.w;
            }
#pragma unroll
            for (int c = (KIND == 2 ? 0 : 1); c < 2; ++c) {
                const float4 RA = reinterpret_cast<const float4*>(rp + T::OFF_RA)[c];
                const float4 RB = reinterpret_cast<const float4*>(rp + T::OFF_RB)[c];
                const float2 sa2 = make_float2(cf.sa[c], cf.sa[c]);
                float2 u[4];
#pragma unroll
                for (int o = 0; o < NB; ++o) {
                    const int k = 2 * o + c;
                    const float2 t = __ffma2_rn(sa2, lk.dxa[k], make_float2(rc[2 * k], rc[2 * k]));
                    const float2 q = __ffma2_rn(lk.dxa[k], t, make_float2(rc[2 * k + 1], rc[2 * k + 1]));
                    const float eA = ex2_approx(q.x), eB = ex2_approx(q.y);
                    const float2 ea = make_float2(eA, eA), eb = make_float2(eB, eB);
                    if (o == 0) {
                        u[0] = __fmul2_rn(lk.C[k][0], ea); u[1] = __fmul2_rn(lk.C[k][1], ea);
                        u[2] = __fmul2_rn(lk.C[k][0], eb); u[3] = __fmul2_rn(lk.C[k][1], eb);
                    } else {
                        u[0] = __ffma2_rn(lk.C[k][0], ea, u[0]); u[1] = __ffma2_rn(lk.C[k][1], ea, u[1]);
                        u[2] = __ffma2_rn(lk.C[k][0], eb, u[2]); u[3] = __ffma2_rn(lk.C[k][1], eb, u[3]);
                    }
                }
                m[0] = __ffma2_rn(make_float2(RA.x, RA.y), u[0], m[0]); m[1] = __ffma2_rn(make_float2(RA.z, RA.w), u[1], m[1]);
                m[2] = __ffma2_rn(make_float2(RB.x, RB.y), u[2], m[2]); m[3] = __ffma2_rn(make_float2(RB.z, RB.w), u[3], m[3]);
            }
        }
        if (TM == 1) {
            tmem_ld16_wait(tv);
            dA = make_float4(__uint_as_float(tv[0]), __uint_as_float(tv[1]), __uint_as_float(tv[2]), __uint_as_float(tv[3]));
            dB = make_float4(__uint_as_float(tv[4]), __uint_as_float(tv[5]), __uint_as_float(tv[6]), __uint_as_float(tv[7]));
            tm += 16;
        } else if (TM == 2) {
            tmem_ld8_wait(tv);
            tm += 8;
        }
        if (TM) {
            wA = make_float4(__uint_as_float(tv[8]), __uint_as_float(tv[9]), __uint_as_float(tv[10]), __uint_as_float(tv[11]));
            wB = make_float4(__uint_as_float(tv[12]), __uint_as_float(tv[13]), __uint_as_float(tv[14]), __uint_as_float(tv[15]));
        }
        finish_step<NX, STORE, PREP>(m, dA, dB, wA, wB, mp, colA, colB, s0, s1);
        if (STORE) mp += G::RG * NX;
        rp += G::RG * T::RS;
        dp += G::RG * NX;
        wp += G::RG * NX;
    }
    sp.rp = rp; sp.dp = dp; sp.wp = wp; sp.mp = mp; sp.tm = tm;
}

// chi-square of one parameter vector over the stamp by ONE warp (TEAM = 1), or this warp's share
// of it (TEAM > 1: warp `tw` takes every TEAM-th row step; the caller adds the partials in a fixed
// order).  Builds the row and column tables in `scratch` (Scratch<NB, NY, TEAM>::FLOATS floats of
// this warp's own shared memory) itself.  `exps` counts the component evaluations (pixels x
// components) the far-field culling left to do.
template <int NB, int NX, int NY, bool STORE, bool PREP, int TEAM = 1, int TM = 0>
__device__ __forceinline__ double warp_chi2(const Coef<NB>& cf, float* __restrict__ scratch,
                                            const float* __restrict__ d, const float* __restrict__ w,
                                            float* __restrict__ model_out, int lane, int tw = 0,
                                            unsigned* exps = nullptr, uint32_t tmem = 0) {
    using G = Geo<NX>;
    using T = Tab<NB>;
    using R = Rows<NY, TEAM>;
    constexpr int K = 2 * NB;
    constexpr int TR = R::TR;
    constexpr int STEPS = TR / G::RG;            // row steps per table (per panel)
    static_assert(TM == 0 || (TEAM == 1 && PREP), "the TMEM pixel store holds prepared stamps for whole-warp passes");
    static_assert(TM != 1 || (R::HALVES == 1 && G::PANELS == 1), "both planes fit the 512 TMEM columns up to 64 x 64 pixels");
    constexpr int TM_STEP = TM == 1 ? 16 : 8;     // TMEM columns per row step
    static_assert(TR % G::RG == 0, "unsupported stamp height");
    static_assert(STEPS % TEAM == 0, "team size must divide the row steps");
    float* rt = scratch;
    float* ct = scratch + Scratch<NB, NY, TEAM>::TAB;
    const int c = lane % G::LPR, g = lane / G::LPR;
    const int swap = (G::PW == 32) ? (g & 1) : 0;
    double acc = 0.0;
    if (exps) *exps += cf.nexp;
#pragma unroll 1
    for (int half = 0; half < R::HALVES; ++half) {
        build_row_table<NB, NX, TR, TEAM>(rt, cf, lane, half * TR, tw);
        const float* dh = d + half * TR * NX;
        const float* wh = w + half * TR * NX;
        float* mh = STORE ? model_out + half * TR * NX : nullptr;
#pragma unroll 1
        for (int pan = 0; pan < G::PANELS; ++pan) {
            const int colA = pan * G::PW + 4 * c + (G::PW / 2) * swap;
            const int colB = pan * G::PW + 4 * c + (G::PW / 2) * (1 - swap);
            // segments of this panel, clipped to the row steps of this table:
            //   [0,wlo) none, [wlo,nlo) wings, [nlo,nhi1) all, [nhi1,whi1) wings, [whi1,STEPS) none
            const bool p1 = G::PANELS > 1 && pan > 0;   // (constant indices keep the coefficients in registers)
            int wlo = p1 ? cf.seg[1][0] : cf.seg[0][0], nlo = p1 ? cf.seg[1][1] : cf.seg[0][1];
            int nhi1 = p1 ? cf.seg[1][2] : cf.seg[0][2], whi1 = p1 ? cf.seg[1][3] : cf.seg[0][3];
            if (R::HALVES > 1) {
                const int base = half * STEPS;
                wlo = min(max(wlo - base, 0), STEPS); nlo = min(max(nlo - base, 0), STEPS);
                nhi1 = min(max(nhi1 - base, 0), STEPS); whi1 = min(max(whi1 - base, 0), STEPS);
            }
            float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
            if (cf.fast) {
                LaneK<NB> lk;
                coop_consts<NB, NX>(lk, ct, cf, lane, pan, colA, colB);
                if (TEAM == 1) {
                    // contiguous steps: the pointers run through the segments
                    StepPtrs sp{rt + g * T::RS, dh + g * NX, wh + g * NX, STORE ? mh + g * NX : nullptr,
                                tmem + (uint32_t)((half * G::PANELS + pan) * STEPS * TM_STEP)};
                    int i = 0;
                    if (NX < 64) {
                        // 32-pixel stamps have no far field (set_cull is never called for them)
                        row_steps_fast<NB, NX, NY, STORE, PREP, 2, TM>(cf, lk, s0, s1, i, STEPS, sp, colA, colB);
                    } else {
                        row_steps_fast<NB, NX, NY, STORE, PREP, 0, TM>(cf, lk, s0, s1, i, wlo, sp, colA, colB);
                        row_steps_fast<NB, NX, NY, STORE, PREP, 1, TM>(cf, lk, s0, s1, i, nlo, sp, colA, colB);
                        row_steps_fast<NB, NX, NY, STORE, PREP, 2, TM>(cf, lk, s0, s1, i, nhi1, sp, colA, colB);
                        row_steps_fast<NB, NX, NY, STORE, PREP, 1, TM>(cf, lk, s0, s1, i, whi1, sp, colA, colB);
                        row_steps_fast<NB, NX, NY, STORE, PREP, 0, TM>(cf, lk, s0, s1, i, STEPS, sp, colA, colB);
                    }
                } else {
                    // a team member owns only STEPS/TEAM steps (no culling in teams): one dense step at a time;
                    // its table holds just those rows
#pragma unroll 1
                    for (int mth = 0; mth < STEPS / TEAM; ++mth) {
                        const int r0 = (mth * TEAM + tw) * G::RG + g;
                        StepPtrs sp{rt + (mth * G::RG + g) * T::RS, dh + r0 * NX, wh + r0 * NX, STORE ? mh + r0 * NX : nullptr, 0u};
                        int i = 0;
                        row_steps_fast<NB, NX, NY, STORE, PREP, 2>(cf, lk, s0, s1, i, 1, sp, colA, colB);
                    }
                }
            } else {
                float2 xd[K][4];   // pixel pairs: (0,1) (2,3) of group A, (0,1) (2,3) of group B
                const float fa = (float)colA, fb = (float)colB;
                const float2 cols[4] = {make_float2(fa, fa + 1.f), make_float2(fa + 2.f, fa + 3.f),
                                        make_float2(fb, fb + 1.f), make_float2(fb + 2.f, fb + 3.f)};
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float2 nx0 = make_float2(-cf.x0[k], -cf.x0[k]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) xd[k][j] = __fadd2_rn(cols[j], nx0);
                }
#pragma unroll 1
                for (int mth = 0; mth < STEPS / TEAM; ++mth) {
                    const int r0 = (mth * TEAM + tw) * G::RG + g;
                    StepPtrs sp{rt + (mth * G::RG + g) * T::RS, dh + r0 * NX, wh + r0 * NX, STORE ? mh + r0 * NX : nullptr, 0u};
                    int i = 0;
                    row_steps<NB, NX, NY, STORE, PREP, 2>(cf, xd, s0, s1, i, 1, sp, colA, colB);
                }
            }
            // FP32 partial sums of one panel (at most 16 steps x 8 pixels over 4 accumulators) -> FP64
            acc += (double)((s0.x + s0.y) + (s1.x + s1.y));
        }
    }
    return warp_sum_f64(acc);
}

// Copy of the prepared stamp into the TMEM pixel store: called by warps 0..3 (one per TMEM lane
// quadrant) after prep_stamp; row step i of the lane goes to columns 16 i .. 16 i + 15 in the order
// the loop consumes them (data A, data B, weight A, weight B).
template <int NX, int NY>
__device__ __forceinline__ void tmem_fill_stamp(uint32_t tmem_base, const float* __restrict__ sd,
                                                const float* __restrict__ sw, int warp, int lane) {
    using G = Geo<NX>;
    static_assert(G::PANELS == 1, "one panel");
    const int c = lane % G::LPR, g = lane / G::LPR;
    const int swap = (G::PW == 32) ? (g & 1) : 0;
    const int colA = 4 * c + (G::PW / 2) * swap, colB = 4 * c + (G::PW / 2) * (1 - swap);
    const uint32_t addr = tmem_base + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
    for (int i = 0; i < NY / G::RG; ++i) {
        const int r = i * G::RG + g;
        const float4 dA = *reinterpret_cast<const float4*>(sd + r * NX + colA);
        const float4 dB = *reinterpret_cast<const float4*>(sd + r * NX + colB);
        const float4 wA = *reinterpret_cast<const float4*>(sw + r * NX + colA);
        const float4 wB = *reinterpret_cast<const float4*>(sw + r * NX + colB);
        const float v[16] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w, wA.x, wA.y, wA.z, wA.w, wB.x, wB.y, wB.z, wB.w};
        tmem_st16(addr + 16 * i, v);
    }
    tmem_wait_st();
}

// 128-pixel stamps: only the weight plane fits (8 columns per row step, 4 tables x 2 panels x 8
// steps = 512 columns, the whole TMEM); the data plane stays in shared memory.  Same order as the
// loop nest of warp_chi2: table, panel, row step.
template <int NX, int NY>
__device__ __forceinline__ void tmem_fill_weights(uint32_t tmem_base, const float* __restrict__ sw, int warp, int lane) {
    using G = Geo<NX>;
    using R = Rows<NY, 1>;
    constexpr int STEPS = R::TR / G::RG;
    static_assert(R::HALVES * G::PANELS * STEPS * 8 <= 512, "weight plane exceeds the TMEM columns");
    const int c = lane % G::LPR, g = lane / G::LPR;
    const int swap = (G::PW == 32) ? (g & 1) : 0;
    const uint32_t addr = tmem_base + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
    for (int hp = 0; hp < R::HALVES * G::PANELS; ++hp) {
        const int half = hp / G::PANELS, pan = hp % G::PANELS;
        const int colA = pan * G::PW + 4 * c + (G::PW / 2) * swap, colB = pan * G::PW + 4 * c + (G::PW / 2) * (1 - swap);
#pragma unroll 1
        for (int i = 0; i < STEPS; ++i) {
            const int r = half * R::TR + i * G::RG + g;
            const float4 wA = *reinterpret_cast<const float4*>(sw + r * NX + colA);
            const float4 wB = *reinterpret_cast<const float4*>(sw + r * NX + colB);
            const float v[8] = {wA.x, wA.y, wA.z, wA.w, wB.x, wB.y, wB.z, wB.w};
            tmem_st8(addr + (uint32_t)((hp * STEPS + i) * 8), v);
        }
    }
    tmem_wait_st();
}

// In-place conversion of a staged stamp to (d*sqrt(w), -sqrt(w)); called by the whole CTA.
__device__ __forceinline__ void prep_stamp(float* __restrict__ sd, float* __restrict__ sw, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float r = sqrtf(sw[i]);
        sw[i] = -r;
        sd[i] *= r;
    }
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

// Per-warp sampler scratch in shared memory: the staged trial vector, the random numbers of the
// next 32 updates (one update per lane, read back as broadcasts) and the shapes of the current
// state.  Keeping these out of registers leaves the register file to the pixel loop.
struct WarpScratch {
    float tf[32];        // staged trial vector (stage_trial)
    double step[32];     // proposal step of update slot i: w*z, or 10^(w*z) for log10 parameters
    double lnu[32];      // log of the accept/reject uniform of update slot i
    int k[32];           // parameter index of update slot i
    float shape[8];      // sa0 sb0 sc0 - sa1 sb1 sc1 -  of the CURRENT state
    float trig[4];       // sin, cos of theta (narrow), sin, cos of theta2 (wide) of the CURRENT state
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

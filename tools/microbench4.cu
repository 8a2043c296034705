// Development micro-benchmark: the product's own warp_chi2 (factorised loop, 64 x 64, 2-body) with
// pixel data from shared memory vs from the TMEM pixel store, 16 warps per SM (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench4 tools/microbench4.cu && tools/microbench4
#include <cstdio>
#include "../olpefit_b200/csrc/lapf_device.cuh"
using namespace lapf;

template <bool TM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) pass_kernel(double* out, int passes) {
    constexpr int NB = 2, NX = 64, NY = 64;
    extern __shared__ __align__(128) float smem[];
    __shared__ uint32_t tslot;
    float* sd = smem;
    float* sw = sd + NX * NY;
    float* rt = sw + NX * NY;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NX * NY; i += blockDim.x) {
        const int r = i / NX, c = i % NX;
        const float m = 6.4f + 9000.f * __expf(-0.1f * ((r - 31.7f) * (r - 31.7f) + (c - 32.3f) * (c - 32.3f)));
        sd[i] = m + (float)((i * 37) % 11) - 5.f;
        sw[i] = 1.f / (1444.f + fabsf(sd[i]));
    }
    __syncthreads();
    prep_stamp(sd, sw, NX * NY);
    __syncthreads();
    uint32_t tbase = 0;
    if (TM) {
        if (warp == 0) tmem_alloc(&tslot, 256);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        tbase = tslot;
        if (warp < 4) tmem_fill_stamp<NX, NY>(tbase, sd, sw, warp, lane);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
    }
    Coef<NB> cf;
    const float x[4] = {32.3f + 0.01f * warp, 32.6f + 0.01f * warp, 41.7f, 42.0f}, y[4] = {31.7f, 31.5f, 39.2f, 39.0f};
    const float amp[4] = {12000.f, 3000.f, 240.f, 60.f};
    for (int k = 0; k < 4; ++k) { cf.x0[k] = x[k]; cf.y0[k] = y[k]; cf.amp[k] = amp[k]; }
    set_shape<NB>(cf, 0, 2.138f, 2.35f, 0.3f);
    set_shape<NB>(cf, 1, 6.41f, 6.84f, 0.5f);
    cf.floor = 6.4f;
    set_fast<NB, NX, NY>(cf, lane);
    set_cull<NB, NX, NY>(cf, lane);
    double acc = 0.0;
    unsigned e = 0;
    for (int p = 0; p < passes; ++p) {
        cf.x0[0] += 1e-4f;
        acc += warp_chi2<NB, NX, NY, false, true, 1, (TM ? 1 : 0)>(cf, rt + warp * (Scratch<NB, NX, NY>::FLOATS), sd, sw, nullptr, lane, 0, &e,
                                                            tbase + ((uint32_t)(32 * (warp & 3)) << 16));
    }
    if (lane == 0 && blockIdx.x == 0) out[warp] = acc;
    if (lane == 0 && blockIdx.x == 0 && warp == 0) out[16] = (double)e / passes;
    __syncthreads();
    if (TM && warp == 0) tmem_dealloc(tbase, 256);
}

template <bool TM, int WARPS>
void run(const char* name, double* d, int sms, int khz) {
    const int smem = (2 * 64 * 64 + 16 * 1408) * 4;
    cudaFuncSetAttribute(pass_kernel<TM, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int passes = 4000;
    float best = 1e30f, ms;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        pass_kernel<TM, WARPS><<<sms, WARPS * 32, smem>>>(d, passes);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double h[17]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-6s %.3f ms  %.0f cycles per pass per scheduler  -> %.3e pixel-evals/s  (chi2 sums %.10e %.10e; comp-evals/pass %.0f)\n",
           name, best, best * 1e-3 * khz * 1e3 / ((WARPS / 4.0) * passes), (double)sms * WARPS * passes * 4096 / (best * 1e-3), h[0], h[7], h[16]);
}

int main() {
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* d; cudaMalloc(&d, 17 * 8);
    run<true, 16>("tmem16", d, sms, khz);
    run<true, 12>("tmem12", d, sms, khz);
    run<true, 8>("tmem8", d, sms, khz);
    return 0;
}

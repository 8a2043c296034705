// Development probe (not part of the product): cp.async.bulk.tensor from a 2-D / 3-D tensor map on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, int VAR>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmap_p, const CUtensorMap* tmap_g, int x0, int y0, int f, int rows, int nx, float* out) {
    const CUtensorMap* tm = (VAR & 2) ? tmap_g : &tmap_p;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(rows * nx * 4) : "memory");
        const uint64_t desc = reinterpret_cast<uint64_t>(tm);
        if (RANK == 3) {
            if (VAR & 1)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(smem_u32(tile)), "l"(desc), "r"(x0), "r"(y0), "r"(f), "r"(smem_u32(&bar)) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(smem_u32(tile)), "l"(desc), "r"(x0), "r"(y0), "r"(f), "r"(smem_u32(&bar)) : "memory");
        } else {
            if (VAR & 1)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(tile)), "l"(desc), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(tile)), "l"(desc), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
        }
    }
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n }" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < rows * nx; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK, int VAR>
void launch(const CUtensorMap& tm, const CUtensorMap* tg, int x0, int y0, int f, int rows, int nx, float* out) {
    probe<RANK, VAR><<<1, 128, rows * nx * 4>>>(tm, tg, x0, y0, f, rows, nx, out);
}

int main(int argc, char** argv) {
    const int var = argc > 1 ? atoi(argv[1]) : 0, only_rank = argc > 2 ? atoi(argv[2]) : 2;
    const int F = 3, fy = 200, fx = 256, rows = 32, nx = 64;
    std::vector<float> h((size_t)F * fy * fx);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, rows * nx * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s, query %d, ptr %p\n", cudaGetErrorString(e), (int)q, p);
    Enc enc = (Enc)p;
    for (int rank = only_rank; rank <= only_rank; ++rank) {
        alignas(64) CUtensorMap tm;
        const cuuint64_t dims[3] = {(cuuint64_t)fx, (cuuint64_t)fy, (cuuint64_t)F};
        const cuuint64_t strides[2] = {(cuuint64_t)fx * 4, (cuuint64_t)fx * fy * 4};
        const cuuint32_t box[3] = {(cuuint32_t)nx, (cuuint32_t)rows, 1u};
        const cuuint32_t es[3] = {1u, 1u, 1u};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d + (rank == 2 ? (size_t)fy * fx : 0), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("rank %d: encode -> %d\n", rank, (int)r);
        const int x0 = argc > 3 ? atoi(argv[3]) : 10, y0 = 20, f = 1;
        CUtensorMap* tg; cudaMalloc(&tg, sizeof(CUtensorMap)); cudaMemcpy(tg, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        printf("variant %d (bit0: no .tile qualifier, bit1: descriptor in global memory), rank %d\n", var, rank);
        if (rank == 3) { if (var == 0) launch<3, 0>(tm, tg, x0, y0, f, rows, nx, out); if (var == 1) launch<3, 1>(tm, tg, x0, y0, f, rows, nx, out);
                         if (var == 2) launch<3, 2>(tm, tg, x0, y0, f, rows, nx, out); if (var == 3) launch<3, 3>(tm, tg, x0, y0, f, rows, nx, out); }
        else { if (var == 0) launch<2, 0>(tm, tg, x0, y0, f, rows, nx, out); if (var == 1) launch<2, 1>(tm, tg, x0, y0, f, rows, nx, out);
               if (var == 2) launch<2, 2>(tm, tg, x0, y0, f, rows, nx, out); if (var == 3) launch<2, 3>(tm, tg, x0, y0, f, rows, nx, out); }
        e = cudaDeviceSynchronize();
        printf("rank %d: kernel -> %s\n", rank, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> o(rows * nx);
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < rows; ++r2)
            for (int c = 0; c < nx; ++c)
                if (o[r2 * nx + c] != h[((size_t)f * fy + y0 + r2) * fx + x0 + c]) ++bad;
        printf("rank %d: %d mismatches\n", rank, bad);
    }
    return 0;
}

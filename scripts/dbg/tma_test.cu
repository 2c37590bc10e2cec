// standalone TMA + mbarrier probe (developer aid)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, float *out, int x0, int row, int pl, int variant) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *buf = (float *)smem;
    uint32_t bar = smem_u32(smem + 12 * 128 + 128 * (threadIdx.x >> 5) * 0 + 2048);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    buf += warp * 1024;
    bar += warp * 8;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (variant & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(12 * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(buf)), "l"(&map), "r"(x0), "r"(row), "r"(pl), "r"(bar) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(0) : "memory");
    } while (!done);
    for (int f = 0; f < 12; ++f) out[(warp * 12 + f) * 32 + lane] = buf[f * 32 + lane];
}
int main(int argc, char **argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int nx = 96, ny = 96, planes = 12;
    std::vector<float> h((size_t)nx * ny * planes);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, 4 * 12 * 32 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    cuuint32_t box[3] = {32, 1, 12}, es[3] = {1, 1, 1};
    CUresult r = ((PFN_encodeTiled)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d variant=%d\n", (int)r, variant);
    int threads = (variant & 2) ? 128 : 32;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    k<<<1, threads, (variant & 4) ? 68096 : 8192>>>(m, o, (variant & 8) ? 4 : 0, 5, 0, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> ho(12 * 32);
        cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost);
        printf("out[0]=%g (expect %g) out[33]=%g (expect %g)\n", ho[0], (float)(5 * nx + ((variant & 8) ? 4 : 0)), ho[33],
               (float)(nx * ny + 5 * nx + 1 + ((variant & 8) ? 4 : 0)));
    }
    return 0;
}

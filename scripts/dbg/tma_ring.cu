// standalone probe of the per-warp TMA ring used by kernels_fused.cu (developer aid)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int RING = 8, PF = 2, SLOT_F = 13 * 32;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
struct Args { int la, lb, nx, variant; float *out; };
__global__ void __launch_bounds__(128, 3) k(const __grid_constant__ Args A, const __grid_constant__ CUtensorMap map) {
    extern __shared__ __align__(128) unsigned char smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WB = RING * SLOT_F * 4 + 3072 + 576 + 64;
    unsigned char *wbase = smem + warp * ((WB + 127) & ~127);
    float *ring = (float *)wbase;
    uint32_t bar0 = smem_u32(wbase + RING * SLOT_F * 4 + ((A.variant & 1) ? 0 : 3072 + 576));
    int x0 = ((A.variant & 32) ? 8 : 10) + 24 * warp;
    if (x0 + 32 > A.nx) x0 = A.nx - 32;
    if (lane == 0) {
        for (int s = 0; s < RING; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + s * 8), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!(A.variant & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    float acc = 0.f;
    if ((A.variant & 8) && warp != 0) return;
    for (int r = (A.la - PF) & ~3; r < A.lb + 4; ++r) {
        __syncwarp();
        int rp = r + PF;
        if (lane == 0 && rp >= A.la && rp < A.lb) {
            int slot = (rp - A.la) & (RING - 1);
            uint32_t bar = bar0 + slot * 8, dst = smem_u32(ring + slot * SLOT_F);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(12 * 128) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(&map), "r"(x0), "r"(rp), "r"(0), "r"(bar) : "memory");
        }
        if (r >= A.la && r < A.lb) {
            int rel = r - A.la;
            mbar_wait(bar0 + (rel & (RING - 1)) * 8, (rel >> 3) & 1);
            acc += ring[(rel & (RING - 1)) * SLOT_F + lane];
        }
    }
    A.out[warp * 32 + lane] = acc;
}
__global__ void __launch_bounds__(128, 3) k2(const __grid_constant__ CUtensorMap map, const __grid_constant__ Args A) {
    extern __shared__ __align__(128) unsigned char smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WB = RING * SLOT_F * 4 + 3072 + 576 + 64;
    unsigned char *wbase = smem + warp * ((WB + 127) & ~127);
    float *ring = (float *)wbase;
    uint32_t bar0 = smem_u32(wbase + RING * SLOT_F * 4 + ((A.variant & 1) ? 0 : 3072 + 576));
    int x0 = ((A.variant & 32) ? 8 : 10) + 24 * warp;
    if (x0 + 32 > A.nx) x0 = A.nx - 32;
    if (lane == 0) {
        for (int s = 0; s < RING; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + s * 8), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!(A.variant & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    float acc = 0.f;
    if ((A.variant & 8) && warp != 0) return;
    for (int r = (A.la - PF) & ~3; r < A.lb + 4; ++r) {
        __syncwarp();
        int rp = r + PF;
        if (lane == 0 && rp >= A.la && rp < A.lb) {
            int slot = (rp - A.la) & (RING - 1);
            uint32_t bar = bar0 + slot * 8, dst = smem_u32(ring + slot * SLOT_F);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(12 * 128) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(&map), "r"(x0), "r"(rp), "r"(0), "r"(bar) : "memory");
        }
        if (r >= A.la && r < A.lb) {
            int rel = r - A.la;
            mbar_wait(bar0 + (rel & (RING - 1)) * 8, (rel >> 3) & 1);
            acc += ring[(rel & (RING - 1)) * SLOT_F + lane];
        }
    }
    A.out[warp * 32 + lane] = acc;
}
int main(int argc, char **argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int rows = argc > 2 ? atoi(argv[2]) : 40;
    int nx = 96, ny = 96, planes = 12;
    std::vector<float> h((size_t)nx * ny * planes, 1.0f);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, 4 * 32 * 4);
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
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 68096);
    Args A{10, 10 + rows, nx, variant, o};
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 68096);
    if (variant & 16) k2<<<1, 128, 68096>>>(m, A); else k<<<1, 128, 68096>>>(A, m);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant=%d rows=%d encode=%d sync: %s\n", variant, rows, (int)r, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        float ho[128];
        cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
        printf("acc[0]=%g acc[127]=%g (expect %d)\n", ho[0], ho[127], rows);
    }
    return 0;
}

// Micro-benchmark (developer aid, not product): issue throughput of the instruction classes the fused RK4
// kernel is made of, on sm_100a: FFMA vs FFMA2 (packed f32x2), FADD vs FADD2, SHFL, LDS.32 vs LDS.64 and mixes.
// Reports warp-instructions per clock per SM sub-partition.  nvcc -arch=sm_100a -O3 -o ubench_fp ubench_fp.cu
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 4096;
constexpr int NACC = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, long long *cyc, float a, float b) {
    __shared__ float sm[2048];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = (float)i * a;
    __syncthreads();
    float x[NACC];
    float2 y[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) {
        x[q] = (float)(threadIdx.x + q);
        y[q] = make_float2(x[q], x[q] + 1.0f);
    }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            if (MODE == 0) x[q] = fmaf(x[q], a, b);                       // FFMA
            if (MODE == 1) y[q] = __ffma2_rn(y[q], a2, b2);               // FFMA2
            if (MODE == 2) x[q] = x[q] + a;                               // FADD
            if (MODE == 3) y[q] = __fadd2_rn(y[q], a2);                   // FADD2
            if (MODE == 4) x[q] = __shfl_down_sync(0xffffffffu, x[q], 1);  // SHFL
            if (MODE == 5) x[q] += sm[(lane + q * 32 + it) & 2047];         // LDS.32 + FADD
            if (MODE == 6) {                                              // LDS.64 + FADD2
                const float2 v = *reinterpret_cast<const float2 *>(&sm[(2 * lane + q * 64 + 2 * it) & 2047]);
                y[q] = __fadd2_rn(y[q], v);
            }
            if (MODE == 7) {  // FFMA + SHFL 3:1 mix
                x[q] = fmaf(x[q], a, b);
                if ((q & 3) == 3) x[q] = __shfl_down_sync(0xffffffffu, x[q], 1);
            }
            if (MODE == 8) {  // FFMA2 + FFMA mix (1:1)
                y[q] = __ffma2_rn(y[q], a2, b2);
                x[q] = fmaf(x[q], a, b);
            }
            if (MODE == 9) {  // FFMA2 + IADD-ish (alu pipe) mix
                y[q] = __ffma2_rn(y[q], a2, b2);
                x[q] = __int_as_float(__float_as_int(x[q]) + q + it);
            }
            if (MODE == 10) {  // FFMA + alu mix
                x[q] = fmaf(x[q], a, b);
                y[q].x = __int_as_float(__float_as_int(y[q].x) + q + it);
            }
            if (MODE == 11) y[q] = __fmul2_rn(y[q], a2);  // FMUL2
            if (MODE == 12) {                              // FFMA with 3 distinct register operands
                x[q] = fmaf(x[q], x[(q + 1) % NACC], x[(q + 2) % NACC]);
            }
            if (MODE == 13) {  // FFMA2 3 distinct register operands
                y[q] = __ffma2_rn(y[q], y[(q + 1) % NACC], y[(q + 2) % NACC]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int q = 0; q < NACC; ++q) s += x[q] + y[q].x + y[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_iter, int warps_per_smsp) {
    float *out;
    long long *cyc;
    const int blocks = 148, threads = 128 * warps_per_smsp;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c[148];
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)c[i] / blocks;
    const double insts = (double)ITERS * NACC * per_iter * warps_per_smsp;  // per SMSP
    printf("%-28s warps/SMSP=%d  %7.3f inst/clk/SMSP  (%.0f cyc, %.3f ms) err=%s\n", name, warps_per_smsp, insts / avg, avg, ms,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w = 2; w <= 8; w *= 2) {
        run<0>("FFMA (imm/const operands)", 1, w);
        run<12>("FFMA 3 reg operands", 1, w);
        run<1>("FFMA2", 1, w);
        run<13>("FFMA2 3 reg operands", 1, w);
        run<2>("FADD", 1, w);
        run<3>("FADD2", 1, w);
        run<11>("FMUL2", 1, w);
        run<4>("SHFL", 1, w);
        run<5>("LDS.32+FADD (count 2)", 2, w);
        run<6>("LDS.64+FADD2 (count 2)", 2, w);
        run<7>("FFMA+SHFL 4:1 (count 1.25)", 1, w);
        run<8>("FFMA2+FFMA (count 2)", 2, w);
        run<9>("FFMA2+IADD (count 2)", 2, w);
        run<10>("FFMA+IADD (count 2)", 2, w);
    }
    return 0;
}

// Host emulation of the CUDA execution model for waves.jl_b200/csrc/latent_core.cuh -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's kernel source unchanged for the CPU: one pthread per CUDA thread, __syncthreads() = a pthread
// barrier, dynamic shared memory = one heap block per CTA, CTAs run one after the other.  It lets the CPU test suite
// (no GPU in the build container) check the kernels' indexing, shared-memory carve-up and barrier placement against
// oracle/latent_oracle.py.  Nothing under waves.jl_b200/ links or loads this file.
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#define LAT_EMU
struct EmuDim {
    int x;
};
static thread_local EmuDim threadIdx, blockIdx;
static EmuDim blockDim;
static pthread_barrier_t g_bar;
static unsigned char *g_smem;
static double g_part[1024 * 3];

#define LAT_DEV static inline
#define LAT_GLOBAL static
#define LAT_SMEM unsigned char *lat_smem = g_smem;
#define LAT_SYNC() pthread_barrier_wait(&g_bar)

// warp-shuffle reduction of the CUDA build, restated with an (emulation-only) barrier
static inline void lat_emu_reduce3(double v[3], double *red, int tid) {
    for (int q = 0; q < 3; ++q) g_part[tid * 3 + q] = v[q];
    pthread_barrier_wait(&g_bar);
    if ((tid & 31) == 0) {
        for (int q = 0; q < 3; ++q) {
            double s = 0.0;
            for (int l = 0; l < 32 && tid + l < blockDim.x; ++l) s += g_part[(tid + l) * 3 + q];
            red[(tid >> 5) * 3 + q] = s;
        }
    }
    pthread_barrier_wait(&g_bar);  // g_part is rewritten by the next call
}

static inline void lat_emu_reduce3f(float v[3], float *red, int tid) {
    for (int q = 0; q < 3; ++q) g_part[tid * 3 + q] = v[q];
    pthread_barrier_wait(&g_bar);
    if ((tid & 31) == 0) {
        for (int q = 0; q < 3; ++q) {
            float s = 0.0f;
            for (int l = 0; l < 32 && tid + l < blockDim.x; ++l) s += (float)g_part[(tid + l) * 3 + q];
            red[(tid >> 5) * 3 + q] = s;
        }
    }
    pthread_barrier_wait(&g_bar);
}

#include "../../waves.jl_b200/csrc/latent_core.cuh"

template <class K>
static void launch(K kernel, const LatentP &p, int nt, size_t smem) {
    blockDim.x = nt;
    g_smem = (unsigned char *)aligned_alloc(16, (smem + 15) & ~(size_t)15);
    for (int b = 0; b < p.batch; ++b) {
        memset(g_smem, 0xff, smem);  // NaN pattern: reads of never-written shared memory show up in the results
        pthread_barrier_init(&g_bar, nullptr, nt);
        std::vector<std::thread> ts;
        for (int t = 0; t < nt; ++t)
            ts.emplace_back([=]() {
                threadIdx.x = t;
                blockIdx.x = b;
                kernel(p);
            });
        for (auto &t : ts) t.join();
        pthread_barrier_destroy(&g_bar);
    }
    free(g_smem);
}

extern "C" int emu_sizeof_latentp() { return (int)sizeof(LatentP); }
extern "C" void emu_latent_integrate(const LatentP *p, int nt) { launch(k_latent_integrate, *p, nt, LAT_FWD_SMEM(p->n)); }
extern "C" void emu_latent_integrate_r1(const LatentP *p, int nt) { launch(k_latent_integrate_r1, *p, nt, LAT_R1_SMEM(p->n)); }
extern "C" void emu_latent_integrate_r2(const LatentP *p, int nt) { launch(k_latent_integrate_r2, *p, nt, LAT_R2_SMEM(p->n)); }
extern "C" void emu_latent_adjoint_r1(const LatentP *p, int nt) { launch(k_latent_adjoint_r1, *p, nt, LAT_ADJ1_SMEM(p->n)); }
extern "C" void emu_latent_adjoint_r2(const LatentP *p, int nt) { launch(k_latent_adjoint_r2, *p, nt, LAT_ADJ1_SMEM(p->n)); }
extern "C" void emu_latent_adjoint(const LatentP *p, int nt) { launch(k_latent_adjoint, *p, nt, LAT_ADJ_SMEM(p->n)); }

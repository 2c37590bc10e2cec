// C ABI of the 1-D latent dynamics (include/waves_b200.h, "latent" section): AcousticDynamics{OneDim} + Integrator +
// compute_latent_energy + adjoint_sensitivity of the reference's AcousticEnergyModel (src/dynamics.jl:97-128, :190-222,
// src/model/acoustic_energy_model.jl:6-15, :86-106).  No CPU compute path: everything runs in the two kernels of
// latent_core.cuh.  Compiled with -fmad=false (the reference never contracts a*b+c).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime.h>

#include "../../include/waves_b200.h"
#include "latent_core.cuh"

int waves_set_error(const char *msg);  // waves_abi.cu: thread-local message, returns 1

#define LFAIL(...)                                  \
    do {                                            \
        char buf_[400];                             \
        snprintf(buf_, sizeof(buf_), __VA_ARGS__);  \
        return waves_set_error(buf_);               \
    } while (0)

#define LCU(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) LFAIL("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct waves_latent {
    int device, n, max_smem;
    float c0, dt, pml_scale, dx;
    float grad8[8];
    cudaStream_t stream;
    // device staging of host arguments (grown on demand): 0 z0, 1 tspan, 2 X, 3 Y, 4 shape, 5 pml, 6 z, 7 energy,
    // 8 w_energy, 9 dL_dz, 10 g_z0, 11 g_Y, 12 g_shape, 13 g_pml
    float *stage[14];
    size_t cap[14];
    int64_t launches;
    cudaEvent_t ev0, ev1;  // around the last kernel launch
    float last_ms;
    int force_generic;  // 1: never use the register fast path (tests compare the two)
    int pair_forward;   // 1: forward pass with the pair kernel where it applies (WAVES_LATENT_PAIR)
    int fast_adjoint;   // 1: reverse pass with the register kernel where it applies (WAVES_LATENT_ADJ_R1)
};

static bool on_device(const void *p) {
    cudaPointerAttributes at;
    const bool dev = p && cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    return dev;
}

static int grow(waves_latent *h, int slot, size_t elems) {
    if (h->cap[slot] >= elems) return 0;
    LCU(cudaStreamSynchronize(h->stream));
    if (h->stage[slot]) cudaFree(h->stage[slot]);
    h->stage[slot] = nullptr;
    h->cap[slot] = 0;
    cudaError_t e = cudaMalloc((void **)&h->stage[slot], sizeof(float) * elems);
    if (e != cudaSuccess) LFAIL("waves_latent: cannot allocate %.1f MB on the device (%s)", elems * 4e-6, cudaGetErrorString(e));
    h->cap[slot] = elems;
    return 0;
}

// device view of an input: the pointer itself when it already is device memory, else a staged copy
static int dev_in(waves_latent *h, int slot, const float *src, size_t elems, const float **out) {
    *out = nullptr;
    if (!src) return 0;
    if (on_device(src)) {
        *out = src;
        return 0;
    }
    if (grow(h, slot, elems)) return 1;
    LCU(cudaMemcpyAsync(h->stage[slot], src, sizeof(float) * elems, cudaMemcpyHostToDevice, h->stream));
    *out = h->stage[slot];
    return 0;
}

// device buffer behind an output: the pointer itself or a staging block to copy back from
static int dev_out(waves_latent *h, int slot, float *dst, size_t elems, float **out) {
    *out = nullptr;
    if (!dst) return 0;
    if (on_device(dst)) {
        *out = dst;
        return 0;
    }
    if (grow(h, slot, elems)) return 1;
    *out = h->stage[slot];
    return 0;
}

static int copy_back(waves_latent *h, int slot, float *dst, size_t elems) {
    if (!dst || on_device(dst)) return 0;
    LCU(cudaMemcpyAsync(dst, h->stage[slot], sizeof(float) * elems, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

extern "C" int waves_latent_create(const waves_latent_config *cfg, waves_latent **out) {
    if (!cfg || !out) LFAIL("waves_latent_create: null argument");
    *out = nullptr;
    if (cfg->n < 3) LFAIL("waves_latent_create: need n >= 3 grid points (three-point boundary rows, src/operators.jl:3-4)");
    if (!cfg->x && !cfg->grad8) LFAIL("waves_latent_create: need the grid x or the gradient rows grad8");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        LFAIL("waves_latent_create: no CUDA device (%s); this library has no CPU path",
              ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) LFAIL("waves_latent_create: device %d out of range (%d)", cfg->device, ndev);
    LCU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    LCU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) LFAIL("waves_latent_create: needs an sm_100a device (Blackwell B200); found sm_%d%d", prop.major, prop.minor);
    if (LAT_FWD_SMEM(cfg->n) > (size_t)prop.sharedMemPerBlockOptin)
        LFAIL("waves_latent_create: n = %d needs %zu bytes of shared memory per sample (limit %zu)", cfg->n,
              (size_t)LAT_FWD_SMEM(cfg->n), (size_t)prop.sharedMemPerBlockOptin);

    waves_latent *h = new waves_latent();
    memset(h, 0, sizeof(*h));
    h->pair_forward = h->fast_adjoint = 1;  // WAVES_LATENT_AUTO
    h->device = cfg->device;
    h->n = cfg->n;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->c0 = cfg->c0;
    h->dt = cfg->dt;
    if (cfg->grad8) {
        memcpy(h->grad8, cfg->grad8, sizeof(h->grad8));
    } else if (waves_build_gradient8(cfg->x, cfg->n, h->grad8)) {
        delete h;
        return 1;
    }
    // pml_scale = dyn.pml[[1]] (src/dynamics.jl:192): the first value of build_pml(::OneDim) (src/pml.jl:6-15)
    if (cfg->pml0 >= 0.0f) {
        h->pml_scale = cfg->pml0;
    } else {
        if (!cfg->x) {
            delete h;
            LFAIL("waves_latent_create: pml0 < 0 asks for build_pml(dim, pml_width, pml_scale)[1], which needs x");
        }
        float p0 = 0.0f;
        if (waves_latent_build_pml(cfg->x, cfg->n, cfg->pml_width, cfg->pml_scale, nullptr, &p0)) {
            delete h;
            return 1;
        }
        h->pml_scale = p0;
    }
    h->dx = cfg->dx > 0.0f ? cfg->dx : (cfg->x ? waves_mean_diff(cfg->x, cfg->n) : 0.0f);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        LFAIL("waves_latent_create: cannot create a stream");
    }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaStreamDestroy(h->stream);
        delete h;
        LFAIL("waves_latent_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return 0;
}

extern "C" int waves_latent_destroy(waves_latent *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int k = 0; k < 14; ++k)
        if (h->stage[k]) cudaFree(h->stage[k]);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int waves_latent_build_pml(const float *x, int n, float width, float scale, float *out, float *first) {
    // build_pml(::OneDim, width, scale), src/pml.jl:6-15
    if (!x || n < 1) return waves_set_error("waves_latent_build_pml: bad arguments");
    const float a0 = fabsf(x[0]), a1 = fabsf(x[n - 1]);
    const float start = (a0 < a1 ? a0 : a1) - width;
    for (int i = 0; i < n; ++i) {
        float v = fabsf(x[i]) - start;
        v = (v > 0.0f ? v : 0.0f) / width;
        v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        const float r = ((v * v) * v) * scale;
        if (out) out[i] = r;
        if (i == 0 && first) *first = r;
    }
    return 0;
}

static int threads_for(int n) {
    int nt = (n + 31) & ~31;
    return nt > 1024 ? 1024 : nt;
}

static void fill_common(const waves_latent *h, LatentP &p, int batch, int steps, int nseq, float freq) {
    memset(&p, 0, sizeof(p));
    p.n = h->n;
    p.batch = batch;
    p.steps = steps;
    p.nseq = nseq;
    p.c0 = h->c0;
    p.dt = h->dt;
    p.hdt = 0.5f * h->dt;
    p.pml_scale = h->pml_scale;
    p.freq = freq;
    p.dx = h->dx;
    for (int k = 0; k < 3; ++k) p.gf[k] = h->grad8[k];
    for (int k = 0; k < 2; ++k) p.gc[k] = h->grad8[3 + k];
    for (int k = 0; k < 3; ++k) p.gl[k] = h->grad8[5 + k];
}

extern "C" int waves_latent_integrate(waves_latent *h, int batch, int steps, int nseq, const float *z0, const float *tspan,
                                      const float *X, const float *Y, const float *shape, float freq, const float *pml,
                                      float *z, float *energy, float *z_last) {
    if (!h) LFAIL("waves_latent_integrate: null handle");
    LCU(cudaSetDevice(h->device));
    if (batch < 1 || steps < 0 || nseq < 2) LFAIL("waves_latent_integrate: need batch >= 1, steps >= 0 and nseq >= 2");
    if (!z0 || !tspan || !X || !Y || !pml) LFAIL("waves_latent_integrate: z0, tspan, X, Y and pml are required");
    const size_t n = h->n, st = 4 * n * (size_t)batch, T = (size_t)steps + 1;
    LatentP p;
    fill_common(h, p, batch, steps, nseq, freq);
    if (dev_in(h, 0, z0, st, &p.z0) || dev_in(h, 1, tspan, T * batch, &p.tspan) || dev_in(h, 2, X, (size_t)nseq * batch, &p.X) ||
        dev_in(h, 3, Y, n * nseq * batch, &p.Y) || dev_in(h, 4, shape, n * batch, &p.shape) ||
        dev_in(h, 5, pml, n * batch, &p.pml))
        return 1;
    if (dev_out(h, 6, z, st * T, &p.z) || dev_out(h, 7, energy, 3 * T * batch, &p.energy) ||
        dev_out(h, 10, z_last, st, &p.z_last))
        return 1;
    // the attribute belongs to the function, not the handle: handles with different n may alternate
    const bool fast = !h->force_generic && h->n <= 1024 && nseq <= LAT_FAST_NSEQ;  // one element per thread, knots in shared memory
    const bool pair = fast && h->pair_forward && (h->n % 2) == 0 && h->n >= 4;  // two elements per thread, packed f32x2
    if (pair)
        LCU(cudaFuncSetAttribute(k_latent_integrate_r2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAT_R2_SMEM(n)));
    else if (fast)
        LCU(cudaFuncSetAttribute(k_latent_integrate_r1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAT_R1_SMEM(n)));
    else
        LCU(cudaFuncSetAttribute(k_latent_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAT_FWD_SMEM(n)));
    LCU(cudaEventRecord(h->ev0, h->stream));
    if (pair)
        k_latent_integrate_r2<<<batch, threads_for(h->n / 2), LAT_R2_SMEM(n), h->stream>>>(p);
    else if (fast)
        k_latent_integrate_r1<<<batch, threads_for(h->n), LAT_R1_SMEM(n), h->stream>>>(p);
    else
        k_latent_integrate<<<batch, threads_for(h->n), LAT_FWD_SMEM(n), h->stream>>>(p);
    h->launches++;
    LCU(cudaGetLastError());
    LCU(cudaEventRecord(h->ev1, h->stream));
    if (copy_back(h, 6, z, st * T) || copy_back(h, 7, energy, 3 * T * batch) || copy_back(h, 10, z_last, st)) return 1;
    LCU(cudaStreamSynchronize(h->stream));
    LCU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

extern "C" int waves_latent_adjoint(waves_latent *h, int batch, int steps, int nseq, const float *z, const float *tspan,
                                    const float *X, const float *Y, const float *shape, float freq, const float *pml,
                                    int adj_mode, const float *w_energy, const float *dL_dz, float *dL_dz0, float *dL_dY,
                                    float *dL_dshape, float *dL_dpml) {
    if (!h) LFAIL("waves_latent_adjoint: null handle");
    LCU(cudaSetDevice(h->device));
    if (batch < 1 || steps < 1 || nseq < 2) LFAIL("waves_latent_adjoint: need batch >= 1, steps >= 1 and nseq >= 2");
    if (!z || !tspan || !X || !Y || !pml || !dL_dz0) LFAIL("waves_latent_adjoint: z, tspan, X, Y, pml and dL_dz0 are required");
    if (!w_energy && !dL_dz) LFAIL("waves_latent_adjoint: need a cotangent (w_energy and/or dL_dz)");
    if (adj_mode != WAVES_ADJ_EXACT && adj_mode != WAVES_ADJ_COMPAT) LFAIL("waves_latent_adjoint: unknown adjoint mode %d", adj_mode);
    const bool fast = h->fast_adjoint && !h->force_generic && h->n <= 1024 && nseq <= LAT_FAST_NSEQ;  // register kernels
    if (!fast && LAT_ADJ_SMEM(h->n) > (size_t)h->max_smem)
        LFAIL("waves_latent_adjoint: n = %d needs %zu bytes of shared memory per sample (limit %d)", h->n,
              (size_t)LAT_ADJ_SMEM(h->n), h->max_smem);
    const size_t n = h->n, st = 4 * n * (size_t)batch, T = (size_t)steps + 1;
    LatentP p;
    fill_common(h, p, batch, steps, nseq, freq);
    p.compat = adj_mode == WAVES_ADJ_COMPAT;
    if (dev_in(h, 6, z, st * T, &p.zt) || dev_in(h, 1, tspan, T * batch, &p.tspan) || dev_in(h, 2, X, (size_t)nseq * batch, &p.X) ||
        dev_in(h, 3, Y, n * nseq * batch, &p.Y) || dev_in(h, 4, shape, n * batch, &p.shape) ||
        dev_in(h, 5, pml, n * batch, &p.pml) || dev_in(h, 8, w_energy, 3 * T * batch, &p.w_energy) ||
        dev_in(h, 9, dL_dz, st * T, &p.dL_dz))
        return 1;
    if (dev_out(h, 10, dL_dz0, st, &p.g_z0) || dev_out(h, 11, dL_dY, n * nseq * batch, &p.g_Y) ||
        dev_out(h, 12, dL_dshape, n * batch, &p.g_shape) || dev_out(h, 13, dL_dpml, n * batch, &p.g_pml))
        return 1;
    if (p.g_Y) LCU(cudaMemsetAsync(p.g_Y, 0, sizeof(float) * n * nseq * batch, h->stream));
    const bool pair = fast && h->pair_forward && (h->n % 2) == 0 && h->n >= 4;  // PAIR | ADJ_R1: two elements per thread
    if (pair)
        LCU(cudaFuncSetAttribute(k_latent_adjoint_r2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAT_ADJ1_SMEM(n)));
    else if (fast)
        LCU(cudaFuncSetAttribute(k_latent_adjoint_r1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAT_ADJ1_SMEM(n)));
    else
        LCU(cudaFuncSetAttribute(k_latent_adjoint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAT_ADJ_SMEM(n)));
    LCU(cudaEventRecord(h->ev0, h->stream));
    if (pair)
        k_latent_adjoint_r2<<<batch, threads_for(h->n / 2), LAT_ADJ1_SMEM(n), h->stream>>>(p);
    else if (fast)
        k_latent_adjoint_r1<<<batch, threads_for(h->n), LAT_ADJ1_SMEM(n), h->stream>>>(p);
    else
        k_latent_adjoint<<<batch, threads_for(h->n), LAT_ADJ_SMEM(n), h->stream>>>(p);
    h->launches++;
    LCU(cudaGetLastError());
    LCU(cudaEventRecord(h->ev1, h->stream));
    if (copy_back(h, 10, dL_dz0, st) || copy_back(h, 11, dL_dY, n * nseq * batch) || copy_back(h, 12, dL_dshape, n * batch) ||
        copy_back(h, 13, dL_dpml, n * batch))
        return 1;
    LCU(cudaStreamSynchronize(h->stream));
    LCU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

extern "C" int waves_latent_set_generic(waves_latent *h, int on) {
    if (!h) LFAIL("waves_latent_set_generic: null handle");
    h->force_generic = on != 0;
    h->pair_forward = h->fast_adjoint = on ? 0 : 1;  // off: back to WAVES_LATENT_AUTO
    return 0;
}

extern "C" int waves_latent_set_variant(waves_latent *h, int variant) {
    if (!h) LFAIL("waves_latent_set_variant: null handle");
    const int known = WAVES_LATENT_GENERIC | WAVES_LATENT_PAIR | WAVES_LATENT_ADJ_R1 | WAVES_LATENT_SINGLE;
    if (variant < 0 || (variant & ~known) || ((variant & WAVES_LATENT_GENERIC) && variant != WAVES_LATENT_GENERIC) ||
        ((variant & WAVES_LATENT_SINGLE) && (variant & WAVES_LATENT_PAIR)))
        LFAIL("waves_latent_set_variant: bad variant flags %d", variant);
    h->force_generic = (variant & WAVES_LATENT_GENERIC) != 0;
    // AUTO = the fastest kernels measured on a B200 (profiles/r2_latent_bench.jsonl): the pair forms of both passes
    h->pair_forward = variant == WAVES_LATENT_AUTO || (variant & WAVES_LATENT_PAIR) != 0;
    h->fast_adjoint = variant == WAVES_LATENT_AUTO || (variant & WAVES_LATENT_ADJ_R1) != 0;
    return 0;
}

extern "C" float waves_latent_last_kernel_ms(waves_latent *h) { return h ? h->last_ms : 0.0f; }

extern "C" int64_t waves_latent_launch_count(waves_latent *h) { return h ? h->launches : 0; }

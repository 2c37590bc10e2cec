// C ABI of the B200-native Waves.jl hot path (include/waves_b200.h): handle management, host-side
// restatements of the reference constructors, and the orchestration of (iter::Integrator)(ui,tspan,θ)
// + the WaveEnv energy metric.  No CPU compute path exists: everything runs on the CUDA device.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <utility>
#include <vector>

#include "waves_internal.h"

static thread_local char g_err[512] = "";

static int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define CU_TRY(call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CHECK_H(h)                                   \
    do {                                             \
        if (!(h)) return fail("null waves_handle");  \
        CU_TRY(cudaSetDevice((h)->device));          \
    } while (0)

int waves_set_error(const char *msg) { return fail("%s", msg); }

extern "C" int waves_version(void) { return WAVES_B200_VERSION; }
extern "C" int waves_build_flags(void) {
#ifdef WAVES_DEV
    return WAVES_BUILD_DEV;
#else
    return 0;
#endif
}
extern "C" const char *waves_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// Host builders (reference constructors restated in float32, same evaluation order)
// ---------------------------------------------------------------------------------------------
extern "C" int waves_range_f32(float start, float stop, int n, float *out) {
    if (n < 1 || !out) return fail("waves_range_f32: bad arguments");
    if (n == 1) {
        out[0] = start;
        return 0;
    }
    // Julia evaluates Float32 ranges in Float64 and rounds each element once (src/dims.jl:56-60)
    double a = start, b = stop;
    for (int i = 0; i < n; ++i) out[i] = (float)((((double)(n - 1 - i)) * a + ((double)i) * b) / (double)(n - 1));
    out[0] = start;
    out[n - 1] = stop;
    return 0;
}

extern "C" int waves_build_pml_profile(const float *x, int n, float width, float scale, float *out) {
    // src/pml.jl:21-29
    if (n < 1 || !x || !out) return fail("waves_build_pml_profile: bad arguments");
    float start = fabsf(x[0]) - width;
    float mn = INFINITY;
    for (int i = 0; i < n; ++i) {
        float ax = fabsf(x[i]);
        if (ax > start && ax < mn) mn = ax;
    }
    for (int i = 0; i < n; ++i) {
        float ax = fabsf(x[i]);
        float v = 0.0f;
        if (ax > start) v = (ax - mn) / width;
        float cube = (v * v) * v;
        out[i] = cube * scale;
    }
    return 0;
}

extern "C" int waves_build_gradient8(const float *x, int n, float *o) {
    // src/operators.jl:10-22: each coefficient is one float32 division by (2Δ)
    if (n < 3 || !x || !o) return fail("waves_build_gradient8: need n >= 3");
    float delta = (x[n - 1] - x[0]) / (float)(n - 1);
    float two = 2.0f * delta;
    const float f[3] = {-3.0f, 4.0f, -1.0f}, c[2] = {-1.0f, 1.0f}, l[3] = {1.0f, -4.0f, 3.0f};
    for (int k = 0; k < 3; ++k) o[k] = f[k] / two;
    for (int k = 0; k < 2; ++k) o[3 + k] = c[k] / two;
    for (int k = 0; k < 3; ++k) o[5 + k] = l[k] / two;
    return 0;
}

extern "C" int waves_build_normal(const float *x, int nx, const float *y, int ny, int n, const float *mu,
                                  const float *sigma, const float *a, float *out) {
    // src/utils.jl:12-18
    if (!x || !y || !mu || !sigma || !a || !out) return fail("waves_build_normal: null argument");
    const float two_pi = 2.0f * 3.14159274101257324f;
    for (size_t q = 0; q < (size_t)nx * ny; ++q) out[q] = 0.0f;
    for (int k = 0; k < n; ++k) {
        float s2 = sigma[k] * sigma[k];
        float coef = 1.0f / (two_pi * s2);
        float den = 2.0f * s2;
        float ca = coef * a[k];
        for (int j = 0; j < ny; ++j) {
            float dy = y[j] - mu[2 * k + 1];
            float dy2 = dy * dy;
            for (int i = 0; i < nx; ++i) {
                float dx = x[i] - mu[2 * k];
                float r2 = (dx * dx) + dy2;
                float e = (float)exp((double)((-r2) / den));
                float fk = ca * e;
                size_t q = (size_t)j * nx + i;
                out[q] = k == 0 ? fk : out[q] + fk;
            }
        }
    }
    return 0;
}

extern "C" float waves_mean_diff(const float *x, int n) {
    // Flux.mean(diff(x)) (src/dims.jl:126): float32 differences, pairwise-summed in the reference;
    // accumulated in double here (agrees to <= 1 ulp)
    double s = 0;
    for (int i = 0; i + 1 < n; ++i) s += (double)(float)(x[i + 1] - x[i]);
    return (float)((float)s / (float)(n - 1));
}

// ---------------------------------------------------------------------------------------------
// Handle
// ---------------------------------------------------------------------------------------------
static void graph_drop(waves_handle *h) {
    if (h->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->graph_exec);
    h->graph_exec = nullptr;
}

static void free_handle(waves_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int k = 0; k < 9; ++k)
        if (h->adj[k]) cudaFree(h->adj[k]);
    if (h->adj_g) cudaFree(h->adj_g);
    graph_drop(h);
    if (h->traj) cudaFree(h->traj);
    if (h->ckpt) cudaFree(h->ckpt);
    if (h->pconst) cudaFree(h->pconst);
    for (int sd = 0; sd < 2; ++sd)
        for (int k = 0; k < 3; ++k)
            if (h->ipc_ptr[sd][k]) cudaIpcCloseMemHandle(h->ipc_ptr[sd][k]);
    if (h->flags) cudaFree(h->flags);
    if (h->obs_in) cudaFree(h->obs_in);
    if (h->obs_out) cudaFree(h->obs_out);
    float **bufs[] = {&h->u[0], &h->u[1], &h->k,      &h->ys,     &h->acc,    &h->b2,     &h->shape,  &h->cplane,
                      &h->d_x,  &h->d_y,  &h->d_sigma, &h->d_cyl0, &h->d_cyl1, &h->d_tspan, &h->d_stage, &h->d_energy};
    for (auto b : bufs)
        if (*b) cudaFree(*b);
    if (h->d_env) cudaFree(h->d_env);
    if (h->d_epart) cudaFree(h->d_epart);
    fused_release(h);
    if (h->h_env) cudaFreeHost(h->h_env);
    if (h->h_cyl0) cudaFreeHost(h->h_cyl0);
    if (h->h_cyl1) cudaFreeHost(h->h_cyl1);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_params) cudaEventDestroy(h->ev_params);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int waves_create(const waves_config *cfg, waves_handle **out) {
    if (!cfg || !out) return fail("waves_create: null argument");
    *out = nullptr;
    int ny_global = cfg->ny_global > 0 ? cfg->ny_global : cfg->ny;
    if (cfg->nx < 32 || cfg->ny < 8 || ny_global < 16)
        return fail("waves_create: need nx >= 32, ny >= 8 and ny_global >= 16 (got %d x %d of %d)", cfg->nx, cfg->ny, ny_global);
    if (cfg->nx != ny_global)
        return fail("waves_create: nx (%d) must equal ny_global (%d): the reference takes sigma_y = sigma_x' "
                    "(src/dynamics.jl:162)", cfg->nx, ny_global);
    if (cfg->n_env < 1) return fail("waves_create: n_env must be >= 1");
    if (!cfg->x || !cfg->y) return fail("waves_create: x and y grids are required");
    if (cfg->row0 < 0 || cfg->row0 + cfg->ny > ny_global) return fail("waves_create: slab rows out of range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail("waves_create: no CUDA device (%s); this library has no CPU path",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return fail("waves_create: device %d out of range (%d)", cfg->device, ndev);
    CU_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail("waves_create: needs an sm_100a device (Blackwell B200); found sm_%d%d", prop.major, prop.minor);

    waves_handle *h = new waves_handle();
    memset(h, 0, sizeof(*h));
    h->device = cfg->device;
    h->sm_count = prop.multiProcessorCount;
    GridP &gp = h->gp;
    gp.nx = cfg->nx;
    gp.nxp = (cfg->nx + 3) & ~3;
    int gtop = cfg->row0 > 0 ? WAVES_HALO : 0;
    int gbot = cfg->row0 + cfg->ny < ny_global ? WAVES_HALO : 0;
    gp.ny_own0 = gtop;
    gp.ny_own = cfg->ny;
    gp.ny_alloc = cfg->ny + gtop + gbot;
    gp.ny_global = ny_global;
    gp.grow0 = cfg->row0 - gtop;
    gp.n_env = cfg->n_env;
    gp.plane = (long long)gp.ny_alloc * gp.nxp;
    gp.env_stride = 12 * gp.plane;
    gp.c0 = cfg->c0;
    gp.b0 = cfg->c0 * cfg->c0;
    gp.dt = cfg->dt;
    gp.hdt = 0.5f * cfg->dt;

    std::vector<float> sig(cfg->nx), g8(8);
    if (cfg->sigma)
        memcpy(sig.data(), cfg->sigma, sizeof(float) * cfg->nx);
    else if (waves_build_pml_profile(cfg->x, cfg->nx, cfg->pml_width, cfg->pml_scale, sig.data())) {
        delete h;  // nothing but the struct exists yet
        return 1;
    }
    if (cfg->grad8)
        memcpy(g8.data(), cfg->grad8, sizeof(float) * 8);
    else if (waves_build_gradient8(cfg->x, cfg->nx, g8.data())) {
        delete h;
        return 1;
    }
    memcpy(gp.g_first, g8.data(), 12);
    memcpy(gp.g_central, g8.data() + 3, 8);
    memcpy(gp.g_last, g8.data() + 5, 12);
    h->d_omega = cfg->d_omega > 0 ? cfg->d_omega
                                  : (float)(waves_mean_diff(cfg->x, cfg->nx) * waves_mean_diff(cfg->y, ny_global));

#define ALLOC(ptr, count)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), sizeof(*(ptr)) * (size_t)(count));       \
        if (e_ != cudaSuccess) {                                                              \
            free_handle(h);                                                                   \
            return fail("cudaMalloc(%s, %zu B): %s", #ptr, sizeof(*(ptr)) * (size_t)(count), cudaGetErrorString(e_)); \
        }                                                                                     \
        cudaMemset((ptr), 0, sizeof(*(ptr)) * (size_t)(count));                               \
    } while (0)

    cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    size_t state = (size_t)gp.env_stride * gp.n_env;
    ALLOC(h->u[0], state);
    ALLOC(h->u[1], state);
    ALLOC(h->shape, (size_t)gp.plane * gp.n_env);
    ALLOC(h->pconst, (size_t)gp.plane * gp.n_env * 2);
    ALLOC(h->flags, 2);
    ALLOC(h->d_x, gp.nx);
    ALLOC(h->d_y, ny_global);
    ALLOC(h->d_sigma, gp.nx + 4);  // (+4: the vectorised reverse kernel reads the profile four columns at a time)
    ALLOC(h->d_env, gp.n_env);
    h->cyl_cap = 32;
    ALLOC(h->d_cyl0, (size_t)gp.n_env * h->cyl_cap * 4);
    ALLOC(h->d_cyl1, (size_t)gp.n_env * h->cyl_cap * 4);
    // pinned host staging of the per-environment parameters: their upload is a true asynchronous copy
    {
        const size_t cb = sizeof(float) * 4 * (size_t)gp.n_env * h->cyl_cap;
        if (cudaMallocHost((void **)&h->h_cyl0, cb) != cudaSuccess || cudaMallocHost((void **)&h->h_cyl1, cb) != cudaSuccess ||
            cudaMallocHost((void **)&h->h_env, sizeof(EnvParams) * gp.n_env) != cudaSuccess) {
            free_handle(h);
            return fail("waves_create: cudaMallocHost failed");
        }
        memset(h->h_cyl0, 0, cb);
        memset(h->h_cyl1, 0, cb);
        memset(h->h_env, 0, sizeof(EnvParams) * gp.n_env);
    }
    h->epart_blocks = 64;
    ALLOC(h->d_epart, (size_t)gp.n_env * h->epart_blocks * 3);
    cudaMemcpy(h->d_x, cfg->x, sizeof(float) * gp.nx, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_y, cfg->y, sizeof(float) * ny_global, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_sigma, sig.data(), sizeof(float) * gp.nx, cudaMemcpyHostToDevice);
    gp.x = h->d_x;
    gp.y = h->d_y;
    gp.sigma = h->d_sigma;
    h->env_dirty = true;
    h->cyl_dirty = true;
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaEventCreateWithFlags(&h->ev_params, cudaEventDisableTiming);
    if (fused_prepare(h)) {
        free_handle(h);
        return 1;  // message set by fused_prepare
    }
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) {
        free_handle(h);
        return fail("waves_create: %s", cudaGetErrorString(ce));
    }
    *out = h;
    return 0;
}

extern "C" int waves_destroy(waves_handle *h) {
    free_handle(h);
    return 0;
}

extern "C" int waves_sync(waves_handle *h) {
    CHECK_H(h);
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" void *waves_stream(waves_handle *h) { return h ? (void *)h->stream : nullptr; }

// host writes to the pinned staging buffers must wait for the last upload that reads them
static void staging_writable(waves_handle *h) {
    if (h->ev_params) cudaEventSynchronize(h->ev_params);
}

static int flush_params(waves_handle *h) {
    const bool any = h->env_dirty || h->cyl_dirty;
    if (h->env_dirty) {
        CU_TRY(cudaMemcpyAsync(h->d_env, h->h_env, sizeof(EnvParams) * h->gp.n_env, cudaMemcpyHostToDevice, h->stream));
        h->env_dirty = false;
    }
    if (h->cyl_dirty) {
        size_t n = sizeof(float) * 4 * (size_t)h->cyl_cap * h->gp.n_env;
        CU_TRY(cudaMemcpyAsync(h->d_cyl0, h->h_cyl0, n, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaMemcpyAsync(h->d_cyl1, h->h_cyl1, n, cudaMemcpyHostToDevice, h->stream));
        h->cyl_dirty = false;
    }
    if (any) CU_TRY(cudaEventRecord(h->ev_params, h->stream));
    return 0;
}

// copy a dense (rows, nx) block <-> the pitched owned rows of one field plane
static int copy_plane(waves_handle *h, float *dev_plane, const float *src, float *dst, int planes) {
    const GridP &gp = h->gp;
    for (int p = 0; p < planes; ++p) {
        float *d = dev_plane + (size_t)p * gp.plane + (size_t)gp.ny_own0 * gp.nxp;
        if (src)
            CU_TRY(cudaMemcpy2DAsync(d, sizeof(float) * gp.nxp, src + (size_t)p * gp.ny_own * gp.nx, sizeof(float) * gp.nx,
                                     sizeof(float) * gp.nx, gp.ny_own, cudaMemcpyDefault, h->stream));
        else
            CU_TRY(cudaMemcpy2DAsync(dst + (size_t)p * gp.ny_own * gp.nx, sizeof(float) * gp.nx, d, sizeof(float) * gp.nxp,
                                     sizeof(float) * gp.nx, gp.ny_own, cudaMemcpyDefault, h->stream));
    }
    return 0;
}

// copy a dense (ny_alloc, nx) block (owned + ghost rows) into one pitched field plane
static int copy_plane_with_ghosts(waves_handle *h, float *dev_plane, const float *src) {
    const GridP &gp = h->gp;
    CU_TRY(cudaMemcpy2DAsync(dev_plane, sizeof(float) * gp.nxp, src, sizeof(float) * gp.nx, sizeof(float) * gp.nx, gp.ny_alloc,
                             cudaMemcpyDefault, h->stream));
    return 0;
}

static int copy_planes_fast(waves_handle *h, float *dev_plane, const float *src, float *dst, int planes) {
    // without ghost rows consecutive planes are consecutive rows: one 2-D copy
    const GridP &gp = h->gp;
    if (gp.ny_alloc != gp.ny_own) return copy_plane(h, dev_plane, src, dst, planes);
    size_t rows = (size_t)planes * gp.ny_own;
    if (src)
        CU_TRY(cudaMemcpy2DAsync(dev_plane, sizeof(float) * gp.nxp, src, sizeof(float) * gp.nx, sizeof(float) * gp.nx, rows,
                                 cudaMemcpyDefault, h->stream));
    else
        CU_TRY(cudaMemcpy2DAsync(dst, sizeof(float) * gp.nx, dev_plane, sizeof(float) * gp.nxp, sizeof(float) * gp.nx, rows,
                                 cudaMemcpyDefault, h->stream));
    return 0;
}

extern "C" int waves_set_state(waves_handle *h, int env, const float *u12) {
    CHECK_H(h);
    if (!u12) return fail("waves_set_state: null buffer");
    const GridP &gp = h->gp;
    if (env >= gp.n_env) return fail("waves_set_state: env %d out of range", env);
    int e0 = env < 0 ? 0 : env, e1 = env < 0 ? gp.n_env : env + 1;
    size_t per_env = (size_t)12 * gp.ny_own * gp.nx;
    for (int e = e0; e < e1; ++e)
        if (copy_planes_fast(h, h->u[h->cur] + (size_t)e * gp.env_stride, u12 + (size_t)(e - e0) * per_env, nullptr, 12))
            return 1;
    h->aux_synced = 0;
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int waves_get_state(waves_handle *h, int env, float *u12) {
    CHECK_H(h);
    if (!u12) return fail("waves_get_state: null buffer");
    const GridP &gp = h->gp;
    if (env >= gp.n_env) return fail("waves_get_state: env %d out of range", env);
    int e0 = env < 0 ? 0 : env, e1 = env < 0 ? gp.n_env : env + 1;
    size_t per_env = (size_t)12 * gp.ny_own * gp.nx;
    for (int e = e0; e < e1; ++e)
        if (copy_planes_fast(h, h->u[h->cur] + (size_t)e * gp.env_stride, nullptr, u12 + (size_t)(e - e0) * per_env, 12))
            return 1;
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int waves_set_source(waves_handle *h, int env, const float *shape, float freq) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (env >= gp.n_env) return fail("waves_set_source: env %d out of range", env);
    int e0 = env < 0 ? 0 : env, e1 = env < 0 ? gp.n_env : env + 1;
    staging_writable(h);
    for (int e = e0; e < e1; ++e) {
        EnvParams &ep = h->h_env[e];
        float *d = h->shape + (size_t)e * gp.plane;
        if (!shape) {
            CU_TRY(cudaMemsetAsync(d, 0, sizeof(float) * gp.plane, h->stream));
            ep.has_source = 0;
            ep.freq = 0.f;
            ep.src_j0 = ep.src_j1 = ep.src_i0 = ep.src_i1 = 0;
        } else {
            // the same plane is broadcast to every env when env < 0
            if (copy_plane_with_ghosts(h, d, shape)) return 1;
            ep.has_source = 1;
            ep.freq = freq;
            int bb[4];
            if (source_bbox(h, e, bb)) return 1;
            ep.src_i0 = bb[0];
            ep.src_i1 = bb[1];
            ep.src_j0 = bb[2];
            ep.src_j1 = bb[3];
        }
    }
    h->env_dirty = true;
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int waves_set_design(waves_handle *h, int env, int ncyl, const float *cyl0, const float *cyl1, float ti,
                                float tf) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (env >= gp.n_env) return fail("waves_set_design: env %d out of range", env);
    if (ncyl < 0 || (ncyl > 0 && (!cyl0 || !cyl1))) return fail("waves_set_design: bad cylinder arguments");
    if (ncyl > h->cyl_cap) {
        int cap = h->cyl_cap;
        while (cap < ncyl) cap *= 2;
        float *n0 = nullptr, *n1 = nullptr;
        CU_TRY(cudaStreamSynchronize(h->stream));  // a pending upload may still read the old staging buffers
        CU_TRY(cudaMallocHost((void **)&n0, sizeof(float) * 4 * (size_t)gp.n_env * cap));
        CU_TRY(cudaMallocHost((void **)&n1, sizeof(float) * 4 * (size_t)gp.n_env * cap));
        memset(n0, 0, sizeof(float) * 4 * (size_t)gp.n_env * cap);
        memset(n1, 0, sizeof(float) * 4 * (size_t)gp.n_env * cap);
        for (int e = 0; e < gp.n_env; ++e) {
            memcpy(n0 + (size_t)e * cap * 4, h->h_cyl0 + (size_t)e * h->cyl_cap * 4, sizeof(float) * 4 * h->cyl_cap);
            memcpy(n1 + (size_t)e * cap * 4, h->h_cyl1 + (size_t)e * h->cyl_cap * 4, sizeof(float) * 4 * h->cyl_cap);
        }
        cudaFreeHost(h->h_cyl0);
        cudaFreeHost(h->h_cyl1);
        h->h_cyl0 = n0;
        h->h_cyl1 = n1;
        CU_TRY(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_cyl0);
        cudaFree(h->d_cyl1);
        CU_TRY(cudaMalloc((void **)&h->d_cyl0, sizeof(float) * 4 * (size_t)cap * gp.n_env));
        CU_TRY(cudaMalloc((void **)&h->d_cyl1, sizeof(float) * 4 * (size_t)cap * gp.n_env));
        h->cyl_cap = cap;
    }
    int e0 = env < 0 ? 0 : env, e1 = env < 0 ? gp.n_env : env + 1;
    staging_writable(h);
    for (int e = e0; e < e1; ++e) {
        EnvParams &ep = h->h_env[e];
        ep.ncyl = ncyl;
        ep.ti = ti;
        ep.tf = tf;
        if (ncyl > 0) {
            memcpy(h->h_cyl0 + (size_t)e * h->cyl_cap * 4, cyl0, sizeof(float) * 4 * ncyl);
            memcpy(h->h_cyl1 + (size_t)e * h->cyl_cap * 4, cyl1, sizeof(float) * 4 * ncyl);
        }
    }
    h->env_dirty = true;
    h->cyl_dirty = true;
    return 0;
}

extern "C" int waves_set_design_batch(waves_handle *h, int ncyl, const float *cyl0, const float *cyl1, float ti, float tf) {
    CHECK_H(h);
    if (ncyl < 0 || (ncyl > 0 && (!cyl0 || !cyl1))) return fail("waves_set_design_batch: bad cylinder arguments");
    for (int e = 0; e < h->gp.n_env; ++e)
        if (waves_set_design(h, e, ncyl, cyl0 ? cyl0 + (size_t)e * ncyl * 4 : nullptr, cyl1 ? cyl1 + (size_t)e * ncyl * 4 : nullptr, ti, tf))
            return 1;
    return 0;
}

extern "C" int waves_set_speed_field(waves_handle *h, int env, const float *c) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (env >= gp.n_env) return fail("waves_set_speed_field: env %d out of range", env);
    int e0 = env < 0 ? 0 : env, e1 = env < 0 ? gp.n_env : env + 1;
    if (c && !h->cplane) {
        CU_TRY(cudaMalloc((void **)&h->cplane, sizeof(float) * (size_t)gp.plane * gp.n_env));
        CU_TRY(cudaMemsetAsync(h->cplane, 0, sizeof(float) * (size_t)gp.plane * gp.n_env, h->stream));
    }
    staging_writable(h);
    for (int e = e0; e < e1; ++e) {
        h->h_env[e].has_cplane = c ? 1 : 0;
        if (c && copy_plane_with_ghosts(h, h->cplane + (size_t)e * gp.plane, c)) return 1;
    }
    h->env_dirty = true;
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

static int ensure_exact_scratch(waves_handle *h) {
    size_t state = (size_t)h->gp.env_stride * h->gp.n_env;
    if (!h->k) CU_TRY(cudaMalloc((void **)&h->k, sizeof(float) * state));
    if (!h->ys) CU_TRY(cudaMalloc((void **)&h->ys, sizeof(float) * state));
    if (!h->acc) CU_TRY(cudaMalloc((void **)&h->acc, sizeof(float) * state));
    if (!h->b2) CU_TRY(cudaMalloc((void **)&h->b2, sizeof(float) * (size_t)h->gp.plane * h->gp.n_env));
    return 0;
}

static int ensure_stage(waves_handle *h, int steps) {
    if (steps + 1 > h->stage_cap) {
        CU_TRY(cudaStreamSynchronize(h->stream));
        if (h->d_tspan) cudaFree(h->d_tspan);
        if (h->d_stage) cudaFree(h->d_stage);
        h->stage_cap = steps + 1 < 128 ? 128 : steps + 1;
        CU_TRY(cudaMalloc((void **)&h->d_tspan, sizeof(float) * h->stage_cap));
        CU_TRY(cudaMalloc((void **)&h->d_stage, sizeof(float) * STAGE_ROW * (size_t)h->stage_cap * h->gp.n_env));
    }
    return 0;
}

static bool any_speed_plane(const waves_handle *h) {
    for (int e = 0; e < h->gp.n_env; ++e)
        if (h->h_env[e].ncyl > 0 || h->h_env[e].has_cplane) return true;
    return false;
}

// one RK4 step in the reference's exact order: 4 x (speed -> rhs -> accumulate), then the update
static int step_exact(waves_handle *h, int steps, int step) {
    if (ensure_exact_scratch(h)) return 1;
    const bool sp = any_speed_plane(h);
    float *u = h->u[h->cur];
    for (int s = 0; s < 4; ++s) {
        const float *in = s == 0 ? u : h->ys;
        if (sp) launch_speed2(h, 0, h->gp.n_env, h->d_stage, steps, step, s, h->b2);
        launch_rhs_exact(h, 0, h->gp.n_env, in, h->k, h->d_stage, steps, step, s);
        launch_rk_update(h, s, u, h->k, h->acc, h->ys);
    }
    launch_rk_final(h, u, h->acc, u);
    return 0;
}

static int step_any(waves_handle *h, int steps, int step, int mode, float *d_e3, int defer_slot = -1) {
    if (mode == WAVES_MODE_EXACT) {
        if (step_exact(h, steps, step)) return 1;
        if (d_e3) launch_energy(h, h->u[h->cur], d_e3, 3 * (steps + 1));
        return 0;
    }
    if (mode != WAVES_MODE_FUSED) return fail("unknown mode %d", mode);
    return launch_fused_step(h, h->d_stage, steps, step, d_e3, defer_slot);
}

extern "C" int waves_rhs(waves_handle *h, int env, float t, float *du12) {
    CHECK_H(h);
    if (env < 0 || env >= h->gp.n_env || !du12) return fail("waves_rhs: bad arguments");
    if (ensure_exact_scratch(h) || ensure_stage(h, 1) || flush_params(h)) return 1;
    CU_TRY(cudaMemcpyAsync(h->d_tspan, &t, sizeof(float), cudaMemcpyHostToDevice, h->stream));
    launch_stage_table(h, h->d_tspan, 1, h->d_stage);
    const EnvParams &ep = h->h_env[env];
    if (ep.ncyl > 0 || ep.has_cplane) launch_speed2(h, env, 1, h->d_stage, 1, 0, 0, h->b2);
    launch_rhs_exact(h, env, 1, h->u[h->cur], h->k, h->d_stage, 1, 0, 0);
    if (copy_planes_fast(h, h->k + (size_t)env * h->gp.env_stride, nullptr, du12, 12)) return 1;
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int waves_step(waves_handle *h, float t, int mode) {
    CHECK_H(h);
    if (ensure_stage(h, 1) || flush_params(h)) return 1;
    CU_TRY(cudaMemcpyAsync(h->d_tspan, &t, sizeof(float), cudaMemcpyHostToDevice, h->stream));
    launch_stage_table(h, h->d_tspan, 1, h->d_stage);
    if (step_any(h, 1, 0, mode & ~WAVES_STEP_ASYNC, nullptr)) return 1;
    if (!(mode & WAVES_STEP_ASYNC)) CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int waves_observe(waves_handle *h, const float *frames, int nsave, int res_x, int res_y, float *out) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (!frames || !out || nsave < 0 || res_x < 1 || res_y < 1) return fail("waves_observe: bad arguments");
    if (res_x > gp.nx || res_y > gp.ny_own) return fail("waves_observe: the resolution must not exceed the grid (src/env.jl:52)");
    if (gp.ny_own != gp.ny_global) return fail("waves_observe: not available on slab handles");
    cudaPointerAttributes at;
    const bool in_dev = cudaPointerGetAttributes(&at, frames) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    const bool out_dev = cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    const size_t plane_e = (size_t)gp.ny_own * gp.nx, frame_e = 12 * plane_e;
    const float *in = frames;
    long long env_stride = (long long)nsave * frame_e, chan_stride = (long long)frame_e;
    if (!in_dev && nsave > 0) {  // only the U_tot planes cross to the device
        const size_t need = (size_t)gp.n_env * nsave * plane_e;
        if (h->obs_in_cap < need) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            if (h->obs_in) cudaFree(h->obs_in);
            CU_TRY(cudaMalloc((void **)&h->obs_in, sizeof(float) * need));
            h->obs_in_cap = need;
        }
        CU_TRY(cudaMemcpy2DAsync(h->obs_in, sizeof(float) * plane_e, frames, sizeof(float) * frame_e, sizeof(float) * plane_e,
                                 (size_t)gp.n_env * nsave, cudaMemcpyHostToDevice, h->stream));
        in = h->obs_in;
        env_stride = (long long)nsave * plane_e;
        chan_stride = (long long)plane_e;
    }
    const size_t out_e = (size_t)gp.n_env * (nsave + 1) * res_x * res_y;
    float *dout = out;
    if (!out_dev) {
        if (h->obs_out_cap < out_e) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            if (h->obs_out) cudaFree(h->obs_out);
            CU_TRY(cudaMalloc((void **)&h->obs_out, sizeof(float) * out_e));
            h->obs_out_cap = out_e;
        }
        dout = h->obs_out;
    }
    launch_imresize(h, in, env_stride, chan_stride, nsave, gp.nx, h->shape, gp.plane, res_x, res_y, dout);
    if (!out_dev) CU_TRY(cudaMemcpyAsync(out, dout, sizeof(float) * out_e, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int waves_energy(waves_handle *h, float *e3) {
    CHECK_H(h);
    if (!e3) return fail("waves_energy: null buffer");
    if (h->energy_cap < 1) {
        CU_TRY(cudaMalloc((void **)&h->d_energy, sizeof(float) * 3 * 128 * (size_t)h->gp.n_env));
        h->energy_cap = 128;
    }
    cudaPointerAttributes at;
    const bool out_dev = cudaPointerGetAttributes(&at, e3) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (out_dev) {  // device output: written by the kernel itself, asynchronous on the handle's stream (no host round trip)
        launch_energy(h, h->u[h->cur], e3, 3);
    } else {
        launch_energy(h, h->u[h->cur], h->d_energy, 3);
        CU_TRY(cudaMemcpyAsync(e3, h->d_energy, sizeof(float) * 3 * h->gp.n_env, cudaMemcpyDefault, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
    }
    CU_TRY(cudaGetLastError());
    return 0;
}

// Is `p` memory a captured CUDA graph may copy to / from asynchronously (device, managed or pinned host memory)?
static bool graph_safe_ptr(const void *p) {
    if (!p) return true;
    cudaPointerAttributes at;
    const bool ok = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered;
    cudaGetLastError();
    return ok;
}

static bool device_ptr_aligned16(const void *p, int device) {
    if (!p || ((uintptr_t)p & 15u)) return false;
    cudaPointerAttributes at;
    const bool ok = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device == device;
    cudaGetLastError();
    return ok;
}

extern "C" int waves_set_traj_stride(waves_handle *h, int stride) {
    CHECK_H(h);
    if (stride < 1) return fail("waves_set_traj_stride: stride must be >= 1");
    h->traj_stride = stride;
    return 0;
}

extern "C" int waves_set_graph(waves_handle *h, int on) {
    CHECK_H(h);
    if (on < 0 || on > 2) return fail("waves_set_graph: mode must be 0, 1 or 2");
    h->graph_off = on ? 0 : 1;
    h->coop_on = on == 2;
    if (!on) graph_drop(h);
    return 0;
}

extern "C" int waves_integrate(waves_handle *h, const float *tspan, int steps, int mode, float *energy,
                               const int32_t *save_steps, int nsave, float *frames, float *u_tot_traj,
                               float *u_inc_traj) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (!tspan || steps < 1) return fail("waves_integrate: need tspan with steps >= 1");
    if (nsave > 0 && (!save_steps || !frames)) return fail("waves_integrate: save_steps/frames missing");
    for (int i = 0; i < nsave; ++i)
        if (save_steps[i] < 0 || save_steps[i] > steps || (i > 0 && save_steps[i] <= save_steps[i - 1]))
            return fail("waves_integrate: save_steps must be ascending within [0, steps]");
    if (ensure_stage(h, steps) || flush_params(h)) return 1;
    if (energy && h->energy_cap < steps + 1) {
        CU_TRY(cudaStreamSynchronize(h->stream));
        graph_drop(h);
        if (h->d_energy) cudaFree(h->d_energy);
        h->energy_cap = steps + 1 < 128 ? 128 : steps + 1;
        CU_TRY(cudaMalloc((void **)&h->d_energy, sizeof(float) * 3 * (size_t)h->energy_cap * gp.n_env));
    }
    const bool fused = mode == WAVES_MODE_FUSED;
    if (!fused && ensure_exact_scratch(h)) return 1;  // (allocations must not happen inside a stream capture)
    CU_TRY(cudaMemcpyAsync(h->d_tspan, tspan, sizeof(float) * (steps + 1), cudaMemcpyDefault, h->stream));
    launch_stage_table(h, h->d_tspan, steps, h->d_stage);

    const size_t frame_elems = (size_t)12 * gp.ny_own * gp.nx, plane_elems = (size_t)gp.ny_own * gp.nx;
    const int stride = h->traj_stride > 0 ? h->traj_stride : 1;
    const size_t traj_frames = (size_t)steps / stride + 1;
    const bool dense = gp.nxp == gp.nx && gp.ny_alloc == gp.ny_own;
    // destinations in device memory whose layout keeps every block 16-byte aligned are written by a copy kernel
    const bool al4 = dense && plane_elems % 4 == 0 && gp.env_stride % 4 == 0 && gp.plane % 4 == 0;
    const bool frames_dev = al4 && device_ptr_aligned16(frames, h->device);
    const bool traj_dev[2] = {al4 && device_ptr_aligned16(u_tot_traj, h->device), al4 && device_ptr_aligned16(u_inc_traj, h->device)};
    int isave = 0;
    auto emit = [&](int frame) -> int {
        const float *u = h->u[h->cur];
        if (isave < nsave && save_steps[isave] == frame) {
            if (dense && frames_dev) {
                // device destination: an SM copy kernel moves this frame of every environment at the HBM copy rate
                launch_copy_blocks(h, u, frames + (size_t)isave * frame_elems, (long long)frame_elems, (long long)gp.env_stride,
                                   (long long)(frame_elems * nsave), gp.n_env);
            } else if (dense) {
                // dense planes: one strided copy moves this frame of every environment
                CU_TRY(cudaMemcpy2DAsync(frames + (size_t)isave * frame_elems, sizeof(float) * frame_elems * nsave, u,
                                         sizeof(float) * gp.env_stride, sizeof(float) * frame_elems, gp.n_env, cudaMemcpyDefault, h->stream));
            } else {
                for (int e = 0; e < gp.n_env; ++e)
                    if (copy_planes_fast(h, const_cast<float *>(u) + (size_t)e * gp.env_stride, nullptr,
                                         frames + ((size_t)e * nsave + isave) * frame_elems, 12))
                        return 1;
            }
            ++isave;
        }
        // U trajectories for renderers (src/plot.jl:25): every `stride`-th frame, one strided copy per field for all environments
        if ((u_tot_traj || u_inc_traj) && frame % stride == 0) {
            const size_t slot = (size_t)(frame / stride);
            for (int f = 0; f < 2; ++f) {
                float *dst = f == 0 ? u_tot_traj : u_inc_traj;
                if (!dst) continue;
                const float *src = u + (size_t)f * 6 * gp.plane;
                if (dense && (f == 0 ? traj_dev[0] : traj_dev[1])) {
                    launch_copy_blocks(h, src, dst + slot * plane_elems, (long long)plane_elems, (long long)gp.env_stride,
                                       (long long)(plane_elems * traj_frames), gp.n_env);
                } else if (dense) {
                    CU_TRY(cudaMemcpy2DAsync(dst + slot * plane_elems, sizeof(float) * plane_elems * traj_frames, src,
                                             sizeof(float) * gp.env_stride, sizeof(float) * plane_elems, gp.n_env, cudaMemcpyDefault, h->stream));
                } else {
                    for (int e = 0; e < gp.n_env; ++e)
                        if (copy_plane(h, const_cast<float *>(src) + (size_t)e * gp.env_stride, nullptr,
                                       dst + ((size_t)e * traj_frames + slot) * plane_elems, 1))
                            return 1;
                }
            }
        }
        return 0;
    };

    // The fused kernel reports the energies of the frame it READS (they ride on the row arrival), so frame n
    // comes out of step n and the last frame needs one stand-alone reduction; the exact path reports the frame
    // it wrote.
    auto run_steps = [&]() -> int {
        isave = 0;
        if (energy && !fused) launch_energy(h, h->u[h->cur], h->d_energy, 3 * (steps + 1));
        if (emit(0)) return 1;
        // short integrations of small batches keep the per-warp energy partials of every step and reduce them in ONE launch at
        // the end: a single environment is latency-bound, and a reduction kernel between consecutive steps sits on its critical path
        const bool defer = energy && fused && steps <= fused_epart_slots(h);
        auto needs_emit = [&](int frame) {
            return (isave < nsave && save_steps[isave] == frame) || ((u_tot_traj || u_inc_traj) && frame % stride == 0);
        };
        for (int n = 0; n < steps;) {
            // small batches: every step up to the next frame that has to be copied out in ONE cooperative launch
            if (fused && h->coop_on && (defer || !energy)) {
                int stop = n + 1;
                while (stop < steps && !needs_emit(stop)) ++stop;
                if (stop - n >= 2) {
                    const int rc = launch_fused_multi(h, h->d_stage, steps, n, stop - n, energy != nullptr);
                    if (rc == 1) return 1;
                    if (rc == 0) {
                        n = stop;
                        if (emit(n)) return 1;
                        continue;
                    }
                }
            }
            float *d_e3 = energy ? h->d_energy + 3 * (size_t)(fused ? n : n + 1) : nullptr;
            if (step_any(h, steps, n, mode, d_e3, defer ? n : -1)) return 1;
            if (emit(n + 1)) return 1;
            ++n;
        }
        if (defer) fused_reduce_deferred(h, steps, h->d_energy, 3 * (steps + 1));
        if (energy && fused) launch_energy(h, h->u[h->cur], h->d_energy + 3 * (size_t)steps, 3 * (steps + 1));
        return 0;
    };

    // The whole integration (every kernel of every step, the forks to the side streams, the frame copies) is captured ONCE
    // into a CUDA graph and replayed by later calls with the same shape: a single environment is launch-bound (a 700^2 step
    // is ~10 us of device work behind 5 launches and 6 event operations), and replaying removes the per-launch host cost.
    // The graph bakes in buffer addresses and which ping-pong buffer is current, so the key below holds all of them.
    // (a batch small enough for the multi-step cooperative launches needs no graph: a handful of launches per call)
    const bool want_graph = fused && !h->graph_off && !h->profile && !h->peer_on && !fused_is_small_batch(h) && graph_safe_ptr(frames) && graph_safe_ptr(u_tot_traj) &&
                            graph_safe_ptr(u_inc_traj) && steps >= 4;
    if (want_graph) {
        uint64_t key[12] = {(uint64_t)steps, (uint64_t)mode, (uint64_t)h->cur, (uint64_t)h->aux_synced, (uint64_t)(energy != nullptr),
                            (uint64_t)nsave, (uint64_t)(uintptr_t)frames, (uint64_t)(uintptr_t)u_tot_traj, (uint64_t)(uintptr_t)u_inc_traj,
                            (uint64_t)stride, (uint64_t)(uintptr_t)h->d_stage, (uint64_t)(uintptr_t)h->d_energy};
        uint64_t hs = 1469598103934665603ull;
        for (int i = 0; i < nsave; ++i) hs = (hs ^ (uint64_t)save_steps[i]) * 1099511628211ull;
        const uint64_t more[5] = {(uint64_t)(uintptr_t)h->d_cyl0, (uint64_t)(uintptr_t)h->d_cyl1, (uint64_t)(uintptr_t)h->d_env,
                                  (uint64_t)h->cyl_cap, (uint64_t)(uintptr_t)h->d_tspan};
        for (uint64_t v : more) hs = (hs ^ v) * 1099511628211ull;
        key[5] ^= hs << 8;
        if (!h->graph_exec || memcmp(key, h->graph_key, sizeof(key)) != 0) {
            graph_drop(h);
            const int cur0 = h->cur, aux0 = h->aux_synced;
            const int64_t l0 = h->launches;
            cudaGraph_t g = nullptr;
            CU_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = run_steps();
            cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
            if (rc || ce != cudaSuccess || !g) {
                if (g) cudaGraphDestroy(g);
                cudaGetLastError();
                h->cur = cur0;
                h->aux_synced = aux0;
                h->launches = l0;
                if (rc) return 1;
                return fail("waves_integrate: stream capture failed: %s", cudaGetErrorString(ce));
            }
            cudaGraphExec_t ge = nullptr;
            ce = cudaGraphInstantiate(&ge, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) {
                h->cur = cur0;
                h->aux_synced = aux0;
                h->launches = l0;
                return fail("waves_integrate: cudaGraphInstantiate: %s", cudaGetErrorString(ce));
            }
            h->graph_exec = ge;
            memcpy(h->graph_key, key, sizeof(key));
            h->graph_launches = h->launches - l0;
            h->graph_cur_after = h->cur;
        } else {  // the replay leaves the handle as the captured run left it
            h->cur = h->graph_cur_after;
            h->aux_synced = 1;
            h->launches += h->graph_launches;
        }
        CU_TRY(cudaGraphLaunch((cudaGraphExec_t)h->graph_exec, h->stream));
        h->graph_replays++;
    } else if (run_steps())
        return 1;
    if (energy)
        CU_TRY(cudaMemcpyAsync(energy, h->d_energy, sizeof(float) * 3 * (size_t)(steps + 1) * gp.n_env, cudaMemcpyDefault,
                               h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}


// ---- reverse pass -----------------------------------------------------------------------------
extern "C" int waves_set_adjoint_checkpoint(waves_handle *h, int every) {
    CHECK_H(h);
    if (every < 0) return fail("waves_set_adjoint_checkpoint: need every >= 0 (0: chosen from the free device memory)");
    h->adj_ckpt = every;
    return 0;
}

static int grow_floats(waves_handle *h, float **buf, size_t *cap, size_t need, const char *what) {
    if (*cap >= need) return 0;
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    cudaError_t ce = cudaMalloc((void **)buf, sizeof(float) * need);
    if (ce != cudaSuccess) return fail("waves_adjoint: cannot allocate %.1f MB for %s (%s)", need * 4e-6, what, cudaGetErrorString(ce));
    *cap = need;
    return 0;
}

extern "C" int waves_adjoint(waves_handle *h, const float *tspan, int steps, int fwd_mode, int adj_mode,
                             const float *w_energy, const float *dL_dzN, float *dL_dz0, float *dL_dc, float *loss) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (!tspan || steps < 1 || !dL_dz0) return fail("waves_adjoint: need tspan, steps >= 1 and dL_dz0");
    if (gp.ny_own != gp.ny_global) return fail("waves_adjoint: not available on slab handles");
    const int adj_flags = adj_mode;
    const bool stagewise = (adj_mode & WAVES_ADJ_STAGEWISE) != 0, tiles_only = (adj_mode & WAVES_ADJ_TILES) != 0;
    adj_mode &= ~(WAVES_ADJ_STAGEWISE | WAVES_ADJ_TILES | WAVES_ADJ_NO_RING | WAVES_ADJ_RING);
    if (adj_mode != WAVES_ADJ_EXACT && adj_mode != WAVES_ADJ_COMPAT) return fail("waves_adjoint: unknown adjoint mode %d", adj_mode);
    if (fwd_mode != WAVES_MODE_FUSED && fwd_mode != WAVES_MODE_EXACT) return fail("waves_adjoint: unknown forward mode %d", fwd_mode);
    const size_t state = (size_t)gp.env_stride * gp.n_env, planes = (size_t)gp.plane * gp.n_env;
    // Without the speed-plane gradient the reverse step is one fused launch set (kernels_adjoint_fused.cu) and needs only the
    // two U planes of every stored state (for the energy cotangent); dL/dc needs the forward stage states and runs the
    // per-stage kernels of kernels_adjoint.cu on full stored states.
    const bool fused_rev = !stagewise && dL_dc == nullptr && gp.nx >= 8 && gp.ny_global >= 8;   // (the fused kernel's border form needs 8 cells)
    const size_t slot = fused_rev ? 2 * planes : state;

    // scratch: 0 w, 1 wsum (both paths); stagewise: 2 lk, 3 ly, 4 y1, 5 y2, 6 y3 (state-sized); 7: three b^2 planes; 8: dL/dc
    const int nstate = fused_rev ? (dL_dzN ? 3 : 2) : 7;
    for (int k = 0; k < nstate; ++k)
        if (!h->adj[k]) CU_TRY(cudaMalloc((void **)&h->adj[k], sizeof(float) * state));
    if (!h->adj[7]) CU_TRY(cudaMalloc((void **)&h->adj[7], sizeof(float) * 3 * planes));
    if (fused_rev && !tiles_only && !h->adj_g) CU_TRY(cudaMalloc((void **)&h->adj_g, sizeof(float) * 2 * planes));
    if (dL_dc && !h->adj[8]) CU_TRY(cudaMalloc((void **)&h->adj[8], sizeof(float) * planes));
    if (fwd_mode == WAVES_MODE_EXACT || !fused_rev)
        if (ensure_exact_scratch(h)) return 1;

    // ---- storage plan: every state of the LAST segment is kept while the forward pass runs; earlier segments keep only the
    // state at their start (checkpoint) and are re-run during the reverse sweep.  One segment (no
    // recomputation) when everything fits.
    int K = steps;
    {
        size_t free_b = 0, total_b = 0;
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
        const double have = 0.8 * ((double)free_b + 4.0 * ((double)h->traj_cap + (double)h->ckpt_cap));
        auto bytes = [&](int k) { return 4.0 * ((double)(k + 1) * slot + (k < steps ? (double)((steps + k - 1) / k) * state : 0.0)); };
        if (h->adj_ckpt > 0)
            K = h->adj_ckpt < steps ? h->adj_ckpt : steps;
        else if (bytes(steps) > have) {
            K = (int)ceil(sqrt((double)steps * (double)state / (double)slot));
            K = K < 1 ? 1 : (K > steps ? steps : K);
            if (bytes(K) > have)
                return fail("waves_adjoint: %d steps of %d environments need %.1f GB even with checkpoints every %d steps (%.1f GB free)",
                            steps, gp.n_env, bytes(K) * 1e-9, K, free_b * 1e-9);
        }
    }
    const int nseg = (steps + K - 1) / K;
    if (grow_floats(h, &h->traj, &h->traj_cap, (size_t)(K + 1) * slot, "the stored states of one segment")) return 1;
    if (nseg > 1 && grow_floats(h, &h->ckpt, &h->ckpt_cap, (size_t)nseg * state, "checkpoints")) return 1;
    h->adj_last_K = K;

    float *W = h->adj[0], *WS = h->adj[1];  // W / WS swap roles after every step
    float *LK = h->adj[2], *LY = h->adj[3];
    float *Y[3] = {h->adj[4], h->adj[5], h->adj[6]};
    float *B2[3] = {h->adj[7], h->adj[7] + planes, h->adj[7] + 2 * planes};
    float *GC = h->adj[8];

    // ---- forward: (iter::Integrator)(z0, t, θ) (src/dynamics.jl:121) ----
    // the stage table gets steps+1 rows: the reference's loop also pulls back through a step taken at t_N
    const int rows = steps + 1;
    const int erows = rows + 1;  // the step kernels lay the energy trace out for a `rows`-step integration: rows + 1 frames per env
    if (ensure_stage(h, rows + 1) || flush_params(h)) return 1;
    if (h->energy_cap < erows) {
        CU_TRY(cudaStreamSynchronize(h->stream));
        graph_drop(h);
        if (h->d_energy) cudaFree(h->d_energy);
        h->energy_cap = erows < 128 ? 128 : erows;
        CU_TRY(cudaMalloc((void **)&h->d_energy, sizeof(float) * 3 * (size_t)h->energy_cap * gp.n_env));
    }
    CU_TRY(cudaMemcpyAsync(h->d_tspan, tspan, sizeof(float) * rows, cudaMemcpyDefault, h->stream));
    launch_stage_table(h, h->d_tspan, rows, h->d_stage);
    auto record = [&](int slot_idx) -> int {  // the state a reverse step needs: its U planes, or all of it
        float *dst = h->traj + (size_t)slot_idx * slot;
        if (fused_rev)
            launch_gather_u(h, h->u[h->cur], dst);
        else
            CU_TRY(cudaMemcpyAsync(dst, h->u[h->cur], sizeof(float) * state, cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    };
    const bool want_loss = loss != nullptr && w_energy != nullptr;
    const bool fused_fwd = fwd_mode == WAVES_MODE_FUSED;
    if (want_loss && !fused_fwd) launch_energy(h, h->u[h->cur], h->d_energy, 3 * erows);
    for (int n = 0; n < steps; ++n) {
        const int seg = n / K, j = n - seg * K;
        if (j == 0 && seg < nseg - 1)
            CU_TRY(cudaMemcpyAsync(h->ckpt + (size_t)seg * state, h->u[h->cur], sizeof(float) * state, cudaMemcpyDeviceToDevice, h->stream));
        if (seg == nseg - 1 && record(j)) return 1;
        float *d_e3 = want_loss ? h->d_energy + 3 * (size_t)(fused_fwd ? n : n + 1) : nullptr;
        if (step_any(h, rows, n, fwd_mode, d_e3)) return 1;
    }
    if (record(steps - (nseg - 1) * K)) return 1;
    if (want_loss && fused_fwd) launch_energy(h, h->u[h->cur], h->d_energy + 3 * (size_t)steps, 3 * erows);
    if (nseg > 1)  // z_N: the handle is left at the last state, like waves_integrate
        CU_TRY(cudaMemcpyAsync(h->ckpt + (size_t)(nseg - 1) * state, h->u[h->cur], sizeof(float) * state, cudaMemcpyDeviceToDevice, h->stream));
    if (loss) {
        for (int e = 0; e < gp.n_env; ++e) loss[e] = 0.0f;
        if (want_loss) {
            std::vector<float> en((size_t)3 * erows * gp.n_env);
            CU_TRY(cudaMemcpyAsync(en.data(), h->d_energy, sizeof(float) * en.size(), cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
            for (int e = 0; e < gp.n_env; ++e) {
                double acc = 0.0;
                for (int i = 0; i <= steps; ++i)
                    for (int k = 0; k < 3; ++k) acc += (double)w_energy[3 * i + k] * (double)en[((size_t)e * erows + i) * 3 + k];
                loss[e] = (float)acc;
            }
        }
    }

    // ---- reverse sweep ----
    const bool sp = any_speed_plane(h);
    // a design that does not move (initial == final cylinders, or a frozen speed plane) has one speed field for every stage
    bool design_static = true, b2_ready = false;
    for (int e = 0; e < gp.n_env && design_static; ++e) {
        const EnvParams &ep = h->h_env[e];
        if (ep.has_cplane || ep.ncyl == 0) continue;
        design_static = memcmp(h->h_cyl0 + (size_t)e * h->cyl_cap * 4, h->h_cyl1 + (size_t)e * h->cyl_cap * 4, sizeof(float) * 4 * ep.ncyl) == 0;
    }
    const float dt = gp.dt, hdt = gp.hdt, s6 = dt / 6.0f, s3 = dt / 3.0f;
    static const float zero3[3] = {0.0f, 0.0f, 0.0f};
    auto w3_of = [&](int i) -> const float * {
        if (!w_energy) return nullptr;
        const float *w3 = w_energy + 3 * i;
        return (w3[0] != 0.0f || w3[1] != 0.0f || w3[2] != 0.0f) ? w3 : nullptr;
    };
    if (dL_dc) CU_TRY(cudaMemsetAsync(GC, 0, sizeof(float) * planes, h->stream));
    CU_TRY(cudaMemsetAsync(W, 0, sizeof(float) * state, h->stream));
    // Fused reverse step: its interior runs on the march kernel (kernels_fused.cu, stage_T) and keeps the auxiliary cotangents of
    // those cells as ONE accumulated plane G, added to them after the sweep; the frame around it takes the shared-memory tiles.
    // march: -1 not decided yet (the first step tells whether the handle's work plan allows it), 0 tiles everywhere, 1 march.
    int march = (fused_rev && !tiles_only) ? -1 : 0;
    // same convention, for the PML ring around the interior: three more launches per step, which pay once the batch is
    // throughput-bound (measured: 32 x 700^2 32.0 -> 36.2 Gcell-updates/s, one 700^2 environment 0.036 -> 0.040 s)
    const bool ring_pays = (long long)gp.n_env * gp.nx * gp.ny_global >= 2000000LL || (adj_flags & WAVES_ADJ_RING);
    int ring = (march && ring_pays && !(adj_flags & WAVES_ADJ_NO_RING)) ? -1 : 0;
    int rect[4] = {0, 0, 0, 0};
    if (march) {
        // the march never writes the auxiliary cotangents of its cells: both ping-pong buffers must hold the same (initial) values there
        CU_TRY(cudaMemsetAsync(h->adj_g, 0, sizeof(float) * 2 * planes, h->stream));
        CU_TRY(cudaMemsetAsync(WS, 0, sizeof(float) * state, h->stream));
    }
    auto speeds = [&](int i) {
        if (sp && !(design_static && b2_ready))
            for (int tau = 0; tau < 3; ++tau) launch_speed2(h, 0, gp.n_env, h->d_stage, rows, i, tau == 0 ? 0 : (tau == 1 ? 1 : 3), B2[tau]);
        b2_ready = true;
    };
    // W += a_i, the energy cotangent of the stored state z_i at `zslot`
    auto add_energy = [&](int i, const float *zslot) {
        if (const float *w3 = w3_of(i)) {
            if (fused_rev)
                launch_energy_cotangent(h, zslot, 2 * gp.plane, gp.plane, W, w3);
            else
                launch_energy_cotangent(h, zslot, gp.env_stride, 6 * gp.plane, W, w3);
        }
    };
    // W += the caller's cotangent of the last state
    auto add_final = [&]() -> int {
        float *T = h->adj[2];  // the caller's cotangent in the pitched layout
        CU_TRY(cudaMemsetAsync(T, 0, sizeof(float) * state, h->stream));
        for (int e = 0; e < gp.n_env; ++e)
            if (copy_planes_fast(h, T + (size_t)e * gp.env_stride, dL_dzN + (size_t)e * 12 * gp.ny_own * gp.nx, nullptr, 12))
                return 1;
        launch_lin3(h, W, 1.0f, W, 1.0f, T, 0.0f, nullptr);
        if (march) CU_TRY(cudaMemcpyAsync(WS, W, sizeof(float) * state, cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    };
    // W <- W + J_step(z_i, t_i)^T W with the per-stage kernels, dL/dc accumulated: the pullback of one runge_kutta call
    // (src/dynamics.jl:105-107)
    auto step_vjp_stagewise = [&](int i, const float *z) {
        speeds(i);
        float *gc = dL_dc ? GC : nullptr;
        // forward stage states y1, y2, y3 (k4 is not needed): y_s = z + a k(y_{s-1}) in one launch each.  The dynamics are
        // linear in the state, so J^T does not depend on them: only the dL/dc term reads them.
        if (gc) {
            float *saved_b2 = h->b2;
            for (int s = 0; s < 3; ++s) {
                h->b2 = B2[s == 0 ? 0 : 1];
                launch_rhs_exact(h, 0, gp.n_env, s == 0 ? z : Y[s - 1], Y[s], h->d_stage, rows, i, s, z, s == 2 ? dt : hdt);
            }
            h->b2 = saved_b2;
        }
        const float *b0 = sp ? B2[0] : nullptr, *b1 = sp ? B2[1] : nullptr, *b2p = sp ? B2[2] : nullptr;
        // each launch forms its cotangent on the fly (a w + b λ_y of the previous stage), returns λ_y = J^T(.) and keeps the
        // running sum WS = w + λ_y3 + λ_y2 + λ_y1 + λ_z; LK / LY alternate as the λ_y buffers
        launch_rhs_transposed(h, s6, W, 0.0f, nullptr, Y[2], b2p, LY, WS, 1, gc);  // λ_k4 = dt/6 w           -> λ_y3 = J4^T λ_k4
        launch_rhs_transposed(h, s3, W, dt, LY, Y[1], b1, LK, WS, 0, gc);          // λ_k3 = dt/3 w + dt λ_y3   -> λ_y2
        launch_rhs_transposed(h, s3, W, hdt, LK, Y[0], b1, LY, WS, 0, gc);         // λ_k2 = dt/3 w + dt/2 λ_y2 -> λ_y1
        launch_rhs_transposed(h, s6, W, hdt, LY, z, b0, LK, WS, 0, gc);            // λ_k1 = dt/6 w + dt/2 λ_y1 -> λ_z
        std::swap(W, WS);                                                          // w <- w + J_step^T w
    };
    // one reverse step i with the stored state at `zslot`: exact  W <- a_i + (I + J_i^T) W ; compat  W <- (I + J_i^T)(W + a_i)
    auto reverse_step = [&](int i, const float *zslot) -> int {
        const float *w3 = w3_of(i);
        if (fused_rev) {
            speeds(i);
            // (a design that does not move has ONE speed plane: the kernel then stages a single plane in shared memory)
            const float *b2v[3] = {sp ? B2[0] : nullptr, sp ? B2[design_static ? 0 : 1] : nullptr, sp ? B2[design_static ? 0 : 2] : nullptr};
            const bool pre = adj_mode == WAVES_ADJ_COMPAT && w3, post = adj_mode == WAVES_ADJ_EXACT && w3;
            if (march) {
                const int rc = launch_adjoint_interior_march(h, W, WS, h->adj_g, h->traj, h->traj_cap, zslot, w3, pre ? 1 : 2, h->d_stage, rows, i, rect);
                if (rc == 1) return 1;
                if (rc < 0 && march > 0) return fail("waves_adjoint: the march kernel became unavailable in the middle of a sweep");
                march = rc == 0 ? 1 : 0;
            }
            if (march) {
                // the PML ring on the march too where the handle has one: the tiles then keep the outer 8 cells of the domain
                int frame[4] = {rect[0], rect[1], rect[2], rect[3]};
                int rc2 = -1;
                if (ring) {
                    rc2 = launch_adjoint_ring_march(h, W, WS, h->traj, h->traj_cap, zslot, w3, pre ? 1 : 2, h->d_stage, rows, i, frame);
                    if (rc2 == 1) return 1;
                    if (rc2 < 0 && ring > 0) return fail("waves_adjoint: the ring march became unavailable in the middle of a sweep");
                    ring = rc2 == 0 ? 1 : 0;
                }
                if (launch_adjoint_step_frame(h, W, WS, b2v, pre ? zslot : nullptr, pre ? w3 : zero3, post ? zslot : nullptr, post ? w3 : zero3, frame))
                    return 1;
                adjoint_march_join(h, ring);
            } else if (launch_adjoint_step_fused(h, W, WS, b2v, pre ? zslot : nullptr, pre ? w3 : zero3, post ? zslot : nullptr, post ? w3 : zero3))
                return 1;
            std::swap(W, WS);
            return 0;
        }
        if (adj_mode == WAVES_ADJ_COMPAT) add_energy(i, zslot);
        step_vjp_stagewise(i, zslot);
        if (adj_mode == WAVES_ADJ_EXACT) add_energy(i, zslot);
        return 0;
    };
    for (int seg = nseg - 1; seg >= 0; --seg) {
        const int s0 = seg * K, len = (steps - s0) < K ? (steps - s0) : K;
        if (seg < nseg - 1) {  // re-run the segment from its checkpoint, storing what its reverse steps need
            CU_TRY(cudaMemcpyAsync(h->u[h->cur], h->ckpt + (size_t)seg * state, sizeof(float) * state, cudaMemcpyDeviceToDevice, h->stream));
            h->aux_synced = 0;
            for (int j = 0; j < len; ++j) {
                if (record(j)) return 1;
                if (step_any(h, rows, s0 + j, fwd_mode, nullptr)) return 1;
            }
        } else {  // the last state: lambda_N = a_N (exact), or the extra step-vjp of the loop as written (compat)
            const float *zN = h->traj + (size_t)len * slot;
            if (dL_dzN && add_final()) return 1;
            if (adj_mode == WAVES_ADJ_EXACT)
                add_energy(steps, zN);
            else if (reverse_step(steps, zN))
                return 1;
        }
        for (int j = len - 1; j >= 0; --j)
            if (reverse_step(s0 + j, h->traj + (size_t)j * slot)) return 1;
    }
    if (nseg > 1) {
        CU_TRY(cudaMemcpyAsync(h->u[h->cur], h->ckpt + (size_t)(nseg - 1) * state, sizeof(float) * state, cudaMemcpyDeviceToDevice, h->stream));
        h->aux_synced = 0;
    }
    if (march > 0) launch_apply_aux_cotangent(h, W, h->adj_g, rect);
    for (int e = 0; e < gp.n_env; ++e) {
        if (copy_planes_fast(h, W + (size_t)e * gp.env_stride, nullptr, dL_dz0 + (size_t)e * 12 * gp.ny_own * gp.nx, 12)) return 1;
        if (dL_dc && copy_planes_fast(h, GC + (size_t)e * gp.plane, nullptr, dL_dc + (size_t)e * gp.ny_own * gp.nx, 1)) return 1;
    }
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaGetLastError());
    return 0;
}

// ---- slab halo plumbing ----------------------------------------------------------------------
extern "C" int waves_halo_describe(waves_handle *h, waves_halo_desc *o) {
    CHECK_H(h);
    if (!o) return fail("waves_halo_describe: null");
    const GridP &gp = h->gp;
    float *u = h->u[h->cur];
    bool lo = gp.ny_own0 > 0, hi = gp.ny_alloc > gp.ny_own0 + gp.ny_own;
    o->send_lo = lo ? u + (size_t)gp.ny_own0 * gp.nxp : nullptr;
    o->recv_lo = lo ? u + (size_t)(gp.ny_own0 - WAVES_HALO) * gp.nxp : nullptr;
    o->send_hi = hi ? u + (size_t)(gp.ny_own0 + gp.ny_own - WAVES_HALO) * gp.nxp : nullptr;
    o->recv_hi = hi ? u + (size_t)(gp.ny_own0 + gp.ny_own) * gp.nxp : nullptr;
    o->plane_stride = gp.plane;
    o->n_planes = gp.n_env * 12;
    o->block_floats = WAVES_HALO * gp.nxp;
    o->pitch = gp.nxp;
    return 0;
}

extern "C" int waves_halo_pack(waves_handle *h, float *lo_buf, float *hi_buf) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (gp.ny_own0 == 0) lo_buf = nullptr;
    if (gp.ny_alloc == gp.ny_own0 + gp.ny_own) hi_buf = nullptr;
    if (lo_buf || hi_buf) launch_pack_halo(h, h->u[h->cur], lo_buf, hi_buf);
    return 0;
}

extern "C" int waves_halo_unpack(waves_handle *h, const float *lo_buf, const float *hi_buf) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (gp.ny_own0 == 0) lo_buf = nullptr;
    if (gp.ny_alloc == gp.ny_own0 + gp.ny_own) hi_buf = nullptr;
    if (lo_buf || hi_buf) launch_unpack_halo(h, h->u[h->cur], lo_buf, hi_buf);
    // aux_synced is NOT cleared: the auxiliary fields a window treats as constant are constant in the neighbour's rows too
    // (same sigma profile), so every exchange delivers the same ghost values for them; new initial conditions arrive through
    // waves_set_state on every rank, which clears the flag.
    return 0;
}

// ---- peer halo (NVLink peer stores instead of an exchange step) -------------------------------------
extern "C" int waves_peer_export(waves_handle *h, waves_peer_info *out) {
    CHECK_H(h);
    if (!out) return fail("waves_peer_export: null");
    memset(out, 0, sizeof(*out));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->ipc_u0, h->u[0]));
    CU_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->ipc_u1, h->u[1]));
    CU_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->ipc_flags, h->flags));
    out->ny_alloc = h->gp.ny_alloc;
    out->ny_own0 = h->gp.ny_own0;
    out->ny_own = h->gp.ny_own;
    out->device = h->device;
    return 0;
}

extern "C" int waves_peer_attach(waves_handle *h, const waves_peer_info *lo, const waves_peer_info *hi) {
    CHECK_H(h);
    const GridP &gp = h->gp;
    if (h->peer_on) return fail("waves_peer_attach: already attached");
    if ((lo != nullptr) != (gp.ny_own0 > 0) || (hi != nullptr) != (gp.ny_alloc > gp.ny_own0 + gp.ny_own))
        return fail("waves_peer_attach: a neighbour record is needed exactly on the sides where this slab has ghost rows");
    const waves_peer_info *nb[2] = {lo, hi};
    for (int s = 0; s < 2; ++s) {
        if (!nb[s]) continue;
        if (s == 0 ? (nb[s]->ny_alloc != nb[s]->ny_own0 + nb[s]->ny_own + WAVES_HALO) : (nb[s]->ny_own0 != WAVES_HALO))
            return fail("waves_peer_attach: neighbour %d has no ghost rows facing this slab", s);
        void *p[3] = {nullptr, nullptr, nullptr};
        const unsigned char *hd[3] = {nb[s]->ipc_u0, nb[s]->ipc_u1, nb[s]->ipc_flags};
        for (int k = 0; k < 3; ++k) {
            cudaIpcMemHandle_t mh;
            memcpy(&mh, hd[k], sizeof(mh));
            CU_TRY(cudaIpcOpenMemHandle(&p[k], mh, cudaIpcMemLazyEnablePeerAccess));
            h->ipc_ptr[s][k] = p[k];
        }
        h->peer_u[s][0] = (float *)p[0];
        h->peer_u[s][1] = (float *)p[1];
        h->peer_plane[s] = (unsigned)((size_t)nb[s]->ny_alloc * gp.nxp);
        if (s == 0) {  // my first owned rows -> the lower neighbour's bottom ghost rows; I am its UPPER neighbour
            h->peer_j0[0] = gp.ny_own0;
            h->peer_dj[0] = nb[s]->ny_own0 + nb[s]->ny_own - gp.ny_own0;
            h->peer_flag[0] = (int *)p[2] + 1;
        } else {       // my last owned rows -> the upper neighbour's top ghost rows; I am its LOWER neighbour
            h->peer_j0[1] = gp.ny_own0 + gp.ny_own - WAVES_HALO;
            h->peer_dj[1] = (nb[s]->ny_own0 - WAVES_HALO) - h->peer_j0[1];
            h->peer_flag[1] = (int *)p[2] + 0;
        }
    }
    h->peer_steps = 0;
    CU_TRY(cudaMemsetAsync(h->flags, 0, sizeof(int) * 2, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->peer_on = 1;
    return 0;
}

// ---- introspection ---------------------------------------------------------------------------
extern "C" int64_t waves_launch_count(waves_handle *h) { return h ? h->launches : 0; }
extern "C" int64_t waves_graph_replays(waves_handle *h) { return h ? h->graph_replays : 0; }

extern "C" int waves_profile(waves_handle *h, int on) {
    CHECK_H(h);
    fused_profile_drain(h);
    h->profile = on;
    h->fused_ms = 0;
    h->fused_launches = 0;
    return 0;
}

extern "C" int waves_profile_read(waves_handle *h, double *ms, int64_t *n) {
    CHECK_H(h);
    CU_TRY(cudaStreamSynchronize(h->stream));
    fused_profile_drain(h);
    if (ms) *ms = h->fused_ms;
    if (n) *n = h->fused_launches;
    return 0;
}

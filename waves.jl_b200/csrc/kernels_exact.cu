// Per-stage CUDA kernels that evaluate the reference's dynamics in its EXACT float32 evaluation
// order (compiled with -fmad=false: Julia never contracts a*b+c).  They serve WAVES_MODE_EXACT,
// waves_rhs (B3 parity, src/dynamics.jl:179-188) and the energy metric (src/env.jl:104-114), and are
// the on-device cross-check of the fused kernel.  sm_100a only.
#include "waves_internal.h"

namespace {

// ---- stage table: RK4 stage times (src/dynamics.jl:10-13) and source factors (src/sources.jl:68) ---
__global__ void k_stage_table(const float *__restrict__ tspan, int steps, const EnvParams *__restrict__ env,
                              int n_env, float dt, float hdt, float *__restrict__ table) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= steps * n_env) return;
    int e = idx / steps, n = idx - e * steps;
    float t = tspan[n];
    float ts[3] = {t, t + hdt, t + dt};
    float *row = table + (size_t)idx * STAGE_ROW;
    const float two_pi = 2.0f * 3.14159274101257324f;  // 2.0f0 * Float32(π)
    for (int s = 0; s < 3; ++s) {
        row[s] = ts[s];
        float arg = (two_pi * ts[s]) * env[e].freq;
        // Julia's sin(::Float32) is the rounding of an accurate double evaluation
        row[3 + s] = env[e].has_source ? (float)sin((double)arg) : 0.0f;
    }
    row[6] = 0.f;
    row[7] = 0.f;
}

// DesignInterpolator(t) for one scalar parameter (src/designs.jl:287-292, :45-49)
__device__ __forceinline__ float interp_param(float p0, float p1, float ti, float tf, float t) {
    float dt = tf - ti;
    dt = dt > 0.0f ? dt : 1.0f;
    float inv = 1.0f / dt;
    float tc = fminf(fmaxf(t, ti), tf);
    float s = tc - ti;
    float dy = p1 + (p0 * -1.0f);
    float slope = dy * inv;
    return p0 + (slope * s);
}

// speed(design(t), grid, c0)^2 (src/designs.jl:99-116) -> b2 plane of each env
__global__ void k_speed2(GridP gp, const EnvParams *__restrict__ env, const float *__restrict__ cyl0,
                         const float *__restrict__ cyl1, int cyl_cap, const float *__restrict__ cplane, int env0,
                         const float *__restrict__ table, int steps, int step, int stage, float *__restrict__ b2) {
    extern __shared__ float cur[];  // [ncyl][4]
    int e = env0 + blockIdx.z;
    const EnvParams ep = env[e];
    float t = table[((size_t)e * steps + step) * STAGE_ROW + (stage == 0 ? 0 : (stage == 3 ? 2 : 1))];
    for (int k = threadIdx.y * blockDim.x + threadIdx.x; k < 4 * ep.ncyl; k += blockDim.x * blockDim.y) {
        size_t o = (size_t)e * cyl_cap * 4 + k;
        cur[k] = interp_param(cyl0[o], cyl1[o], ep.ti, ep.tf, t);
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int jl = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= gp.nx || jl >= gp.ny_alloc) return;
    size_t q = (size_t)e * gp.plane + (size_t)jl * gp.nxp + i;
    if (ep.has_cplane) {
        float c = cplane[q];
        b2[q] = c * c;
        return;
    }
    int g = gp.grow0 + jl;
    float xs = gp.x[i], ys = gp.y[min(max(g, 0), gp.ny_global - 1)];
    int cnt = 0;
    float cd = 0.0f;
    for (int k = 0; k < ep.ncyl; ++k) {
        float ddx = xs - cur[4 * k], ddy = ys - cur[4 * k + 1];
        float r2 = cur[4 * k + 2] * cur[4 * k + 2];
        float d2 = (ddx * ddx) + (ddy * ddy);
        bool m = d2 < r2;
        cnt += m;
        cd = cd + (m ? 1.0f : 0.0f) * cur[4 * k + 3];
    }
    float c = ((cnt == 0 ? 1.0f : 0.0f) * gp.c0) + cd;
    b2[q] = c * c;
}

struct Plane {
    const float *p;
    const float *sh;  // source shape plane or nullptr
    float s;          // source factor
    __device__ __forceinline__ float at(long long q) const { return sh ? (p[q] + (sh[q] * s)) : p[q]; }
};

// ∇*u along x (src/operators.jl:45), SparseMatrixCSC accumulation order
__device__ __forceinline__ float d_x(const GridP &gp, const Plane &pl, long long row, int i) {
    int n = gp.nx;
    if (i == 0)
        return ((gp.g_first[0] * pl.at(row)) + (gp.g_first[1] * pl.at(row + 1))) + (gp.g_first[2] * pl.at(row + 2));
    if (i == n - 1)
        return ((gp.g_last[0] * pl.at(row + n - 3)) + (gp.g_last[1] * pl.at(row + n - 2))) +
               (gp.g_last[2] * pl.at(row + n - 1));
    return (gp.g_central[0] * pl.at(row + i - 1)) + (gp.g_central[1] * pl.at(row + i + 1));
}

// (∇*u')' along y (src/operators.jl:46); g = global row, jl = local row
__device__ __forceinline__ float d_y(const GridP &gp, const Plane &pl, long long base, int jl, int g, int i) {
    long long P = gp.nxp;
    if (g == 0)
        return ((gp.g_first[0] * pl.at(base + jl * P + i)) + (gp.g_first[1] * pl.at(base + (jl + 1) * P + i))) +
               (gp.g_first[2] * pl.at(base + (jl + 2) * P + i));
    if (g == gp.ny_global - 1)
        return ((gp.g_last[0] * pl.at(base + (jl - 2) * P + i)) + (gp.g_last[1] * pl.at(base + (jl - 1) * P + i))) +
               (gp.g_last[2] * pl.at(base + jl * P + i));
    int jm = max(jl - 1, 0), jp = min(jl + 1, gp.ny_alloc - 1);  // clamped only in never-consumed ghost rows
    return (gp.g_central[0] * pl.at(base + jm * P + i)) + (gp.g_central[1] * pl.at(base + jp * P + i));
}

// acoustic_dynamics for both wavefields (src/dynamics.jl:151-188); blockIdx.z = env*2 + wavefield
__global__ void __launch_bounds__(256)
k_rhs_exact(GridP gp, const EnvParams *__restrict__ env, int env0, const float *__restrict__ u, float *__restrict__ kout,
            const float *__restrict__ b2, const float *__restrict__ shape, const float *__restrict__ table, int steps,
            int step, int stage, const float *__restrict__ zadd, float za) {
    // zadd != nullptr: write the RK stage state zadd + za * k instead of k (used by the reverse pass to re-derive y1..y3)
    int e = env0 + (blockIdx.z >> 1), w = blockIdx.z & 1;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int jl = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= gp.nx || jl >= gp.ny_alloc) return;
    int g = gp.grow0 + jl;
    if (g < 0 || g >= gp.ny_global) return;
    const EnvParams ep = env[e];
    const long long P = gp.plane;
    const float *x = u + (long long)e * gp.env_stride + (long long)w * 6 * P;
    float *k = kout + (long long)e * gp.env_stride + (long long)w * 6 * P;
    float sf = table[((size_t)e * steps + step) * STAGE_ROW + 3 + (stage == 0 ? 0 : (stage == 3 ? 2 : 1))];
    const float *sh = shape + (long long)e * P;
    Plane pU{x, nullptr, 0.f}, pUf{x, sh, sf}, pVx{x + P, nullptr, 0.f}, pVy{x + 2 * P, nullptr, 0.f};
    if (!ep.has_source) pUf.sh = nullptr;  // U .+ 0.0f0
    long long row = (long long)jl * gp.nxp, q = row + i;
    float b = (w == 0 && (ep.ncyl > 0 || ep.has_cplane)) ? b2[(long long)e * P + q] : gp.b0;
    float sx = gp.sigma[i], sy = gp.sigma[g];
    bool border = (i == 0) || (i == gp.nx - 1) || (g == 0) || (g == gp.ny_global - 1);
    float bc = border ? 0.0f : 1.0f;  // build_dirichlet (src/dims.jl:117-124)
    float U = pU.at(q), Vx = pVx.at(q), Vy = pVy.at(q);
    float Px = x[3 * P + q], Py = x[4 * P + q], Om = x[5 * P + q];
    float Vxx = d_x(gp, pVx, row, i);
    float Vyy = d_y(gp, pVy, 0, jl, g, i);
    float Ux = d_x(gp, pUf, row, i);
    float Uy = d_y(gp, pUf, 0, jl, g, i);
    float dU = ((((b * (Vxx + Vyy)) + Px) + Py) - ((sx + sy) * U)) - Om;
    float r6[6] = {bc * dU, Ux - (sx * Vx), Uy - (sy * Vy), (b * sx) * Vyy, (b * sy) * Vxx, (sx * sy) * U};
    const float *zb = zadd ? zadd + (long long)e * gp.env_stride + (long long)w * 6 * P : nullptr;
#pragma unroll
    for (int f = 0; f < 6; ++f) k[f * P + q] = zb ? zb[f * P + q] + za * r6[f] : r6[f];
}

// runge_kutta accumulation (src/dynamics.jl:10-15): k1 .+ 2*k2 .+ 2*k3 .+ k4 left to right; y = u .+ (a*k)
__global__ void k_rk_update(long long n, int stage, float a, const float *__restrict__ u, const float *__restrict__ k,
                            float *__restrict__ acc, float *__restrict__ ys) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    float kk = k[q];
    if (stage == 0)
        acc[q] = kk;
    else if (stage == 3)
        acc[q] = acc[q] + kk;
    else
        acc[q] = acc[q] + 2.0f * kk;
    if (stage < 3) ys[q] = u[q] + a * kk;
}

// du = 1/6f0 * (...); return du * dt; u' = u .+ du  (src/dynamics.jl:14-15,41)
__global__ void k_rk_final(long long n, float dt, const float *__restrict__ u, const float *__restrict__ acc,
                           float *__restrict__ out) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float sixth = 1.0f / 6.0f;
    float du = sixth * acc[q];
    out[q] = u[q] + du * dt;
}

// Energy metric (src/env.jl:104-111): Σ u_tot², Σ u_inc², Σ (u_tot-u_inc)² over owned rows, f64 accumulation.
__global__ void __launch_bounds__(256) k_energy_partial(GridP gp, const float *__restrict__ u, double *__restrict__ part) {
    int e = blockIdx.y;
    const float *ut = u + (long long)e * gp.env_stride, *ui = ut + 6 * gp.plane;
    double st = 0, si = 0, ss = 0;
    long long cells = (long long)gp.ny_own * gp.nx;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        int j = (int)(c / gp.nx), i = (int)(c - (long long)j * gp.nx);
        long long q = (long long)(j + gp.ny_own0) * gp.nxp + i;
        float a = ut[q], b = ui[q], d = a - b;
        st += (double)a * (double)a;
        si += (double)b * (double)b;
        ss += (double)d * (double)d;
    }
    __shared__ double sm[3][8];
    for (int o = 16; o > 0; o >>= 1) {
        st += __shfl_down_sync(0xffffffffu, st, o);
        si += __shfl_down_sync(0xffffffffu, si, o);
        ss += __shfl_down_sync(0xffffffffu, ss, o);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sm[0][w] = st;
        sm[1][w] = si;
        sm[2][w] = ss;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0;
        for (int k = 0; k < 8; ++k) s += sm[threadIdx.x][k];
        part[((size_t)e * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
    }
}

__global__ void k_energy_final(const double *__restrict__ part, int nblk, float d_omega, float *__restrict__ e3,
                               int env_stride3) {
    int e = blockIdx.x;
    if (threadIdx.x < 3) {
        double s = 0;
        for (int k = 0; k < nblk; ++k) s += part[((size_t)e * nblk + k) * 3 + threadIdx.x];
        e3[(size_t)e * env_stride3 + threadIdx.x] = (float)s * d_omega;
    }
}

// Halo pack/unpack for slab decomposition: WAVES_HALO rows x nxp floats per field plane.
__global__ void k_halo_copy(GridP gp, float *__restrict__ u, float *__restrict__ lo, float *__restrict__ hi, int unpack) {
    int blk = WAVES_HALO * gp.nxp;
    long long total = (long long)gp.n_env * 12 * blk;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        long long pl = q / blk;
        int r = (int)(q - pl * blk);
        float *base = u + pl * gp.plane;
        if (!unpack) {
            if (lo) lo[q] = base[(long long)gp.ny_own0 * gp.nxp + r];
            if (hi) hi[q] = base[(long long)(gp.ny_own0 + gp.ny_own - WAVES_HALO) * gp.nxp + r];
        } else {
            if (lo) base[(long long)(gp.ny_own0 - WAVES_HALO) * gp.nxp + r] = lo[q];
            if (hi) base[(long long)(gp.ny_own0 + gp.ny_own) * gp.nxp + r] = hi[q];
        }
    }
}

// imresize(w, resolution) of Images.jl / ImageTransformations.jl (bilinear, no prefilter) for the observation of
// RLBase.state(env) (src/env.jl:132-137): output pixel i (1-based) samples the input at s (i - 0.5) + 0.5 with s = n_in / n_out
// along each axis.  Coordinates and weights in double like the reference (Int / Int scale factors), one rounding to float.
// in: channel c of environment e is at in + e * env_stride + c * chan_stride (pitch nxp); out: (n_env, n_chan, ry, rx).
__global__ void k_imresize(int nx, int ny, int in_pitch, const float *__restrict__ in, long long env_stride, long long chan_stride, int n_chan,
                           const float *__restrict__ last, long long last_env_stride, int last_pitch, int rx, int ry,
                           float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    const int e = blockIdx.z / (n_chan + 1), c = blockIdx.z - e * (n_chan + 1);
    if (i >= rx || j >= ry) return;
    const float *src = c < n_chan ? in + (long long)e * env_stride + (long long)c * chan_stride : last + (long long)e * last_env_stride;
    const int nxp = c < n_chan ? in_pitch : last_pitch;
    const double sx = (double)nx / (double)rx, sy = (double)ny / (double)ry;
    const double x = sx * ((double)i + 0.5) - 0.5, y = sy * ((double)j + 0.5) - 0.5;  // 0-based source coordinates
    int x0 = (int)floor(x), y0 = (int)floor(y);
    const double fx = x - (double)x0, fy = y - (double)y0;
    const int x1 = min(max(x0 + 1, 0), nx - 1), y1 = min(max(y0 + 1, 0), ny - 1);
    x0 = min(max(x0, 0), nx - 1);
    y0 = min(max(y0, 0), ny - 1);
    const double a = src[(long long)y0 * nxp + x0], b = src[(long long)y0 * nxp + x1];
    const double cc = src[(long long)y1 * nxp + x0], d = src[(long long)y1 * nxp + x1];
    const double top = (1.0 - fx) * a + fx * b, bot = (1.0 - fx) * cc + fx * d;
    out[(((long long)e * (n_chan + 1) + c) * ry + j) * rx + i] = (float)((1.0 - fy) * top + fy * bot);
}

}  // namespace

void launch_imresize(waves_handle *h, const float *in, long long env_stride, long long chan_stride, int n_chan, int in_pitch,
                     const float *last, long long last_env_stride, int rx, int ry, float *out) {
    dim3 blk(32, 8), grd((rx + 31) / 32, (ry + 7) / 8, h->gp.n_env * (n_chan + 1));
    k_imresize<<<grd, blk, 0, h->stream>>>(h->gp.nx, h->gp.ny_own, in_pitch, in, env_stride, chan_stride, n_chan, last, last_env_stride,
                                          h->gp.nxp, rx, ry, out);
    h->launches++;
}

void launch_stage_table(waves_handle *h, const float *d_tspan, int steps, float *d_table) {
    int n = steps * h->gp.n_env;
    k_stage_table<<<(n + 127) / 128, 128, 0, h->stream>>>(d_tspan, steps, h->d_env, h->gp.n_env, h->gp.dt, h->gp.hdt,
                                                           d_table);
    h->launches++;
}

void launch_speed2(waves_handle *h, int env0, int nenv, const float *d_table, int steps, int step, int stage,
                   float *b2) {
    dim3 blk(32, 8), grd((h->gp.nx + 31) / 32, (h->gp.ny_alloc + 7) / 8, nenv);
    size_t sm = sizeof(float) * 4 * (size_t)(h->cyl_cap > 0 ? h->cyl_cap : 1);
    k_speed2<<<grd, blk, sm, h->stream>>>(h->gp, h->d_env, h->d_cyl0, h->d_cyl1, h->cyl_cap, h->cplane, env0, d_table,
                                          steps, step, stage, b2);
    h->launches++;
}

void launch_rhs_exact(waves_handle *h, int env0, int nenv, const float *u_in, float *k_out, const float *d_table,
                      int steps, int step, int stage, const float *zadd, float za) {
    dim3 blk(32, 8), grd((h->gp.nx + 31) / 32, (h->gp.ny_alloc + 7) / 8, nenv * 2);
    k_rhs_exact<<<grd, blk, 0, h->stream>>>(h->gp, h->d_env, env0, u_in, k_out, h->b2, h->shape, d_table, steps, step,
                                            stage, zadd, za);
    h->launches++;
}

void launch_rk_update(waves_handle *h, int stage, const float *u, const float *k, float *acc, float *ys) {
    long long n = h->gp.env_stride * h->gp.n_env;
    float a = stage == 2 ? h->gp.dt : h->gp.hdt;
    k_rk_update<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, stage, a, u, k, acc, ys);
    h->launches++;
}

void launch_rk_final(waves_handle *h, const float *u_in, const float *acc, float *u_out) {
    long long n = h->gp.env_stride * h->gp.n_env;
    k_rk_final<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, h->gp.dt, u_in, acc, u_out);
    h->launches++;
}

void launch_energy(waves_handle *h, const float *u, float *d_e3, int env_stride3) {
    dim3 grd(h->epart_blocks, h->gp.n_env);
    k_energy_partial<<<grd, 256, 0, h->stream>>>(h->gp, u, h->d_epart);
    k_energy_final<<<h->gp.n_env, 32, 0, h->stream>>>(h->d_epart, h->epart_blocks, h->d_omega, d_e3, env_stride3);
    h->launches += 2;
}

// Strided block copy on the SMs: block b of `n_blocks` is `elems` contiguous floats at src + b * src_stride, stored at
// dst + b * dst_stride.  Saved frames (env.wave, src/env.jl:116) and U trajectories of a large batch are tens of GB per
// frame; a pitched cudaMemcpy2DAsync between device buffers ran at a fraction of the HBM copy rate (measured).
// All strides, `elems` and both base addresses are multiples of 4 floats (checked by the caller): 16-byte accesses only.
namespace {
__global__ void __launch_bounds__(256) k_copy_blocks(const float4 *__restrict__ src, float4 *__restrict__ dst, long long elems4,
                                                     long long src_stride4, long long dst_stride4, int chunks) {
    const int b = blockIdx.x / chunks, c = blockIdx.x - b * chunks;
    const float4 *s = src + (long long)b * src_stride4;
    float4 *d = dst + (long long)b * dst_stride4;
    const long long per = (elems4 + chunks - 1) / chunks, i0 = (long long)c * per, i1 = min(i0 + per, elems4);
    long long i = i0 + threadIdx.x;
    // four independent 16-byte loads in flight per thread
    for (; i + 3 * 256 < i1; i += 4 * 256) {
        const float4 a = __ldcs(s + i), b4 = __ldcs(s + i + 256), c4 = __ldcs(s + i + 512), d4 = __ldcs(s + i + 768);
        __stcs(d + i, a);
        __stcs(d + i + 256, b4);
        __stcs(d + i + 512, c4);
        __stcs(d + i + 768, d4);
    }
    for (; i < i1; i += 256) __stcs(d + i, __ldcs(s + i));
}
}  // namespace

void launch_copy_blocks(waves_handle *h, const float *src, float *dst, long long elems, long long src_stride, long long dst_stride,
                        int n_blocks) {
    // ~16 resident CTAs per SM's worth of work, at least 64 KB per CTA
    long long want = (long long)h->sm_count * 16 / (n_blocks > 0 ? n_blocks : 1);
    const long long cap = elems / 16384 + 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    const int chunks = (int)want;
    k_copy_blocks<<<(unsigned)((long long)n_blocks * chunks), 256, 0, h->stream>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst),
                                                                                 elems / 4, src_stride / 4, dst_stride / 4, chunks);
    h->launches++;
}

void launch_pack_halo(waves_handle *h, const float *u, float *lo, float *hi) {
    k_halo_copy<<<h->sm_count * 4, 256, 0, h->stream>>>(h->gp, const_cast<float *>(u), lo, hi, 0);
    h->launches++;
}

void launch_unpack_halo(waves_handle *h, float *u, const float *lo, const float *hi) {
    k_halo_copy<<<h->sm_count * 4, 256, 0, h->stream>>>(h->gp, u, const_cast<float *>(lo), const_cast<float *>(hi), 1);
    h->launches++;
}

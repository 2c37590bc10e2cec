// 1-D latent dynamics of Waves.jl's AcousticEnergyModel: kernels.
//
//   (dyn::AcousticDynamics{OneDim})(x, t, θ)   src/dynamics.jl:190-222      θ = [C, F, PML]
//   runge_kutta, (iter::Integrator)(ui, tspan::Matrix, θ)   src/dynamics.jl:9-16, :37-49
//   LinearInterpolation (θ[1])                 src/utils.jl:70-98
//   Source(shape, freq)(t::Vector) (θ[2])      src/sources.jl:21-23
//   compute_latent_energy                      src/model/acoustic_energy_model.jl:6-15
//   adjoint_sensitivity (batchwise OneDim)     src/dynamics.jl:97-118
//
// B200 mapping: the reference integrates a batch of ~1 k-element samples with ~100 CUDA.jl kernel launches per RK4
// step.  Here ONE CTA owns one sample for the whole integration: its state, the four RK4 stage states and the running
// stage sum live in shared memory (84 KB at n = 1024), neighbours are read from shared memory, each RK4 step costs
// four block barriers and no launch; HBM sees one 16-byte-per-element frame write per step (when the trajectory is
// wanted) and nothing else.  The arithmetic is the reference's float32 evaluation order (this file is compiled with
// -fmad=false), so the forward kernel is bit-exact against oracle/latent_oracle.py.
//
// The file contains only thread-index / barrier / shared-memory code so that tests/emu/ can compile the SAME source
// for the host (threads = pthreads, barrier = pthread barrier) and check indexing and hazards without a GPU.  That
// host build is a test of this source, never a product path.
#pragma once
#include <math.h>
#include <stddef.h>

#ifndef LAT_EMU
#define LAT_DEV __device__ __forceinline__
#define LAT_GLOBAL __global__ __launch_bounds__(1024, 1)  // one 1024-thread CTA per SM: <= 64 registers per thread
#define LAT_SMEM extern __shared__ __align__(16) unsigned char lat_smem[];
#define LAT_SYNC() __syncthreads()
#endif

#define LAT_MAX_WARPS 32

// layouts: C order of the reference's column-major arrays (same memory image)
struct LatentP {
    int n, batch, steps, nseq;
    float c0, dt, hdt, pml_scale, freq, dx;
    float gf[3], gc[2], gl[3];  // rows of dyn.grad: first(3), central(2), last(3)  (src/operators.jl:10-22)
    const float *z0;            // (n, 4, batch)          [batch][4][n]
    const float *tspan;         // (steps+1, batch)       [batch][steps+1]
    const float *X;             // (nseq, batch)          [batch][nseq]
    const float *Y;             // (n, nseq, batch)       [batch][nseq][n]
    const float *shape;         // (n, batch)             [batch][n]      nullptr: NoSource
    const float *pml;           // (n, batch)             [batch][n]
    float *z;                   // (n, 4, batch, steps+1) [steps+1][batch][4][n]   nullable
    float *energy;              // (steps+1, 3, batch)    [batch][3][steps+1]      nullable
    float *z_last;              // (n, 4, batch)                                   nullable
    // ---- reverse pass only ----
    int compat;                 // 0: exact discrete adjoint, 1: the reference loop as written
    const float *zt;            // stored trajectory [steps+1][batch][4][n]
    const float *w_energy;      // nullable [batch][3][steps+1]: dL/d(energy)
    const float *dL_dz;         // nullable [steps+1][batch][4][n]
    float *g_z0;                // [batch][4][n]
    float *g_Y;                 // nullable [batch][nseq][n]   (zero-initialised by the caller)
    float *g_shape;             // nullable [batch][n]
    float *g_pml;               // nullable [batch][n]
};

// (∇ * v)[i]: SparseMatrixCSC * dense accumulates nzval*v in increasing column order from 0, no FMA (src/operators.jl:45)
template <class V>
LAT_DEV float lat_grad(const LatentP &p, int i, V v) {
    const int n = p.n;
    if (i == 0) return ((p.gf[0] * v(0)) + (p.gf[1] * v(1))) + (p.gf[2] * v(2));
    if (i == n - 1) return ((p.gl[0] * v(n - 3)) + (p.gl[1] * v(n - 2))) + (p.gl[2] * v(n - 1));
    return (p.gc[0] * v(i - 1)) + (p.gc[1] * v(i + 1));
}

// (∇' * w)[j] for the reverse pass
template <class V>
LAT_DEV float lat_grad_t(const LatentP &p, int j, V w) {
    const int n = p.n;
    float r = 0.0f;
    if (j + 1 >= 1 && j + 1 <= n - 2) r += p.gc[0] * w(j + 1);  // row j+1 holds gc[0] in column j
    if (j - 1 >= 1 && j - 1 <= n - 2) r += p.gc[1] * w(j - 1);  // row j-1 holds gc[1] in column j
    if (j <= 2) r += p.gf[j] * w(0);
    if (j >= n - 3) r += p.gl[j - (n - 3)] * w(n - 1);
    return r;
}

// sin(2.0f0 * pi * t * freq) (src/sources.jl:22): ((2f0*π)*t)*freq in Float32; Julia's sin(::Float32) is the rounding
// of an accurate double evaluation
LAT_DEV float lat_sin_factor(float t, float freq) {
    const float two_pi = 2.0f * 3.14159274101257324f;
    const float arg = (two_pi * t) * freq;
    return (float)sin((double)arg);
}

// the reference's segment mask (src/utils.jl:79-80)
LAT_DEV bool lat_mask(float l, float r, float rend, float t) { return ((l <= t) && (t < r)) || ((r == rend) && (rend == t)); }

// c0 * C(t) at element i (src/utils.jl:70-86, src/dynamics.jl:210): masks and sums as written.  Xb: the knots of this
// sample (global or shared), Yb: its values [nseq][n]
LAT_DEV float lat_speed_elem(const LatentP &p, const float *Xb, const float *Yb, float t, int i) {
    const float rend = Xb[p.nseq - 1];
    float x0 = 0.0f, y0 = 0.0f, dydx = 0.0f;
    for (int k = 0; k + 1 < p.nseq; ++k) {
        const float l = Xb[k], r = Xb[k + 1];
        if (lat_mask(l, r, rend, t)) {
            const float dd = (r - t) - (l - t);  // diff(X .- x): the query is subtracted before the difference
            const float yk = Yb[(size_t)k * p.n + i], yk1 = Yb[(size_t)(k + 1) * p.n + i];
            x0 = x0 + l;
            y0 = y0 + yk;
            dydx = dydx + (yk1 - yk) / dd;
        }
    }
    return p.c0 * (y0 + (t - x0) * dydx);
}

// the same for every element this thread owns (generic kernels)
LAT_DEV void lat_speed(const LatentP &p, int b, float t, float *cout, int tid, int nt) {
    const float *Xb = p.X + (size_t)b * p.nseq;
    const float *Yb = p.Y + (size_t)b * p.nseq * p.n;
    for (int i = tid; i < p.n; i += nt) cout[i] = lat_speed_elem(p, Xb, Yb, t, i);
}

// one right-hand side at element i: S = stage state [4][n] in shared memory, c0c = c0*C(t), fs = source factor
LAT_DEV void lat_rhs_s(const LatentP &p, const float *S, float a, const float *shp, float s, float fs, int i, float k[4]) {
    const int n = p.n;
    const float *Ut = S, *Vt = S + n, *Ui = S + 2 * n, *Vi = S + 3 * n;
    const float bc = (i == 0 || i == n - 1) ? 0.0f : 1.0f;  // build_dirichlet(::OneDim), src/dims.jl:111-115
    const float gVt = lat_grad(p, i, [&](int j) { return Vt[j]; });
    const float gUt = lat_grad(p, i, [&](int j) { return Ut[j] + (shp[j] * fs); });  // ∇ * (U_tot .+ f)
    const float gVi = lat_grad(p, i, [&](int j) { return Vi[j]; });
    const float gUi = lat_grad(p, i, [&](int j) { return Ui[j] + (shp[j] * fs); });
    k[0] = ((a * gVt) - (s * Ut[i])) * bc;     // src/dynamics.jl:210, :217
    k[1] = (a * gUt) - (s * Vt[i]);            // :211
    k[2] = ((p.c0 * gVi) - (s * Ui[i])) * bc;  // :213, :219
    k[3] = (gUi * p.c0) - (s * Vi[i]);         // :214
}

LAT_DEV void lat_rhs(const LatentP &p, const float *S, const float *c0c, const float *shp, const float *sig, float fs, int i,
                     float k[4]) {
    lat_rhs_s(p, S, c0c[i], shp, sig[i], fs, i, k);
}

// Σ over the block of three doubles per thread; result valid in out[0..2] (shared) after the NEXT barrier
LAT_DEV void lat_reduce3_warp(double v[3], double *red, int tid) {
#ifndef LAT_EMU
    for (int o = 16; o > 0; o >>= 1)
        for (int q = 0; q < 3; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
    if ((tid & 31) == 0)
        for (int q = 0; q < 3; ++q) red[(tid >> 5) * 3 + q] = v[q];
#else
    lat_emu_reduce3(v, red, tid);
#endif
}

// shared-memory carve-up of the forward kernel: doubles first (alignment), then floats
#define LAT_FWD_FLOATS(n) (21 * (size_t)(n) + 8)
#define LAT_FWD_SMEM(n) (sizeof(double) * 3 * LAT_MAX_WARPS + sizeof(float) * LAT_FWD_FLOATS(n))

LAT_GLOBAL void k_latent_integrate(LatentP p) {
    LAT_SMEM
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x, nw = (nt + 31) >> 5;
    const int T = p.steps + 1;
    double *red = (double *)lat_smem;
    float *u = (float *)(red + 3 * LAT_MAX_WARPS);
    float *acc = u + 4 * n, *sa = acc + 4 * n, *sb = sa + 4 * n;
    float *cA = sb + 4 * n, *cB = cA + n, *cC = cB + n, *shp = cC + n, *sig = shp + n;
    float *fac = sig + n;  // [2][4]: source factors of the three stage times, double-buffered by step parity

    for (int i = tid; i < n; i += nt) {
        for (int f = 0; f < 4; ++f) {
            const float v = p.z0[((size_t)b * 4 + f) * n + i];
            u[f * n + i] = v;
            if (p.z) p.z[((size_t)b * 4 + f) * n + i] = v;  // frame 0 = ui (src/dynamics.jl:46)
        }
        shp[i] = p.shape ? p.shape[(size_t)b * n + i] : 0.0f;
        sig[i] = p.pml_scale * p.pml[(size_t)b * n + i];  // σ = dyn.pml[[1]] .* PML (src/dynamics.jl:192-193)
    }

    const float sixth = 1.0f / 6.0f;
    for (int s = 0; s <= p.steps; ++s) {
        // ---- phase 0: energy of frame s, parameters of step s (every array written here is owner-only or double-buffered)
        if (p.energy) {
            double e[3] = {0.0, 0.0, 0.0};
            for (int i = tid; i < n; i += nt) {
                const float a = u[i], c = u[2 * n + i], d = a - c;  // tot .- inc in Float32 first
                e[0] += (double)a * a;
                e[1] += (double)c * c;
                e[2] += (double)d * d;
            }
            lat_reduce3_warp(e, red, tid);
        }
        float *fc = fac + (s & 1) * 4;
        if (s < p.steps) {
            const float t0 = p.tspan[(size_t)b * T + s];
            const float t1 = t0 + p.hdt, t2 = t0 + p.dt;
            lat_speed(p, b, t0, cA, tid, nt);
            lat_speed(p, b, t1, cB, tid, nt);
            lat_speed(p, b, t2, cC, tid, nt);
            if (tid < 3) fc[tid] = p.shape ? lat_sin_factor(tid == 0 ? t0 : (tid == 1 ? t1 : t2), p.freq) : 0.0f;
        }
        LAT_SYNC();
        if (p.energy && tid < 3) {
            double sum = 0.0;
            for (int w = 0; w < nw; ++w) sum += red[w * 3 + tid];
            p.energy[((size_t)b * 3 + tid) * T + s] = (float)sum * p.dx;  // sum(tot .^ 2, dims = 1) * dx
        }
        if (s == p.steps) break;

        // ---- stage 1: k1 = f(u, t)
        for (int i = tid; i < n; i += nt) {
            float k[4];
            lat_rhs(p, u, cA, shp, sig, fc[0], i, k);
            for (int f = 0; f < 4; ++f) {
                acc[f * n + i] = k[f];
                sa[f * n + i] = u[f * n + i] + (p.hdt * k[f]);
            }
        }
        LAT_SYNC();
        // ---- stage 2: k2 = f(u + dt/2 k1, t + dt/2)
        for (int i = tid; i < n; i += nt) {
            float k[4];
            lat_rhs(p, sa, cB, shp, sig, fc[1], i, k);
            for (int f = 0; f < 4; ++f) {
                acc[f * n + i] = acc[f * n + i] + (2.0f * k[f]);
                sb[f * n + i] = u[f * n + i] + (p.hdt * k[f]);
            }
        }
        LAT_SYNC();
        // ---- stage 3: k3 = f(u + dt/2 k2, t + dt/2)
        for (int i = tid; i < n; i += nt) {
            float k[4];
            lat_rhs(p, sb, cB, shp, sig, fc[1], i, k);
            for (int f = 0; f < 4; ++f) {
                acc[f * n + i] = acc[f * n + i] + (2.0f * k[f]);
                sa[f * n + i] = u[f * n + i] + (p.dt * k[f]);
            }
        }
        LAT_SYNC();
        // ---- stage 4: k4 = f(u + dt k3, t + dt); u <- u + (1/6 (k1 + 2k2 + 2k3 + k4)) dt   (src/dynamics.jl:14-15, :41)
        for (int i = tid; i < n; i += nt) {
            float k[4];
            lat_rhs(p, sa, cC, shp, sig, fc[2], i, k);
            for (int f = 0; f < 4; ++f) {
                const float du = (sixth * (acc[f * n + i] + k[f])) * p.dt;
                const float v = u[f * n + i] + du;
                u[f * n + i] = v;  // owner-only: stage 4 reads neighbours of sa, never of u
                if (p.z) p.z[(((size_t)(s + 1) * p.batch + b) * 4 + f) * n + i] = v;
            }
        }
        // no barrier here: phase 0 of the next step touches only this thread's own elements (and the other fac buffer)
    }
    if (p.z_last)
        for (int i = tid; i < n; i += nt)
            for (int f = 0; f < 4; ++f) p.z_last[((size_t)b * 4 + f) * n + i] = u[f * n + i];
}

// ------------------------------------------------------------------------------------------------------------------
// Register fast path of the forward kernel: n <= 1024 (the reference's latent grid), one element per thread.
// ncu on the generic kernel above (profiles/r1_latent_ncu_summary.json): issue-bound, ~1300 warp instructions per warp per
// RK4 step.  Here the element's state, stage sum, damping and the interpolation operands stay in registers; only the four
// stage values are exchanged through two ping-pong shared buffers (4 STS + 8 LDS per stage); the knots live in shared
// memory and, when they increase strictly (then at most one segment mask is true), a cursor replaces the scan; the sine
// factors of LAT_CH steps are tabulated by all threads at once instead of by three threads on the critical path.
// Same float32 evaluation order: bit-exact with the generic kernel and the oracle.  Interior threads read i-1 / i+1;
// the two boundary threads evaluate the three-point rows from shared memory.
#define LAT_CH 256
#define LAT_FAST_NSEQ 64
#define LAT_R1_FLOATS(n) (9 * (size_t)(n) + LAT_FAST_NSEQ + 2 * 3 * LAT_CH + 3 * LAT_MAX_WARPS + 8)
#define LAT_R1_SMEM(n) (sizeof(float) * LAT_R1_FLOATS(n))

LAT_DEV void lat_reduce3f_warp(float v[3], float *red, int tid) {
#ifndef LAT_EMU
    for (int o = 16; o > 0; o >>= 1)
        for (int q = 0; q < 3; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
    if ((tid & 31) == 0)
        for (int q = 0; q < 3; ++q) red[(tid >> 5) * 3 + q] = v[q];
#else
    lat_emu_reduce3f(v, red, tid);
#endif
}

// Σ of one double per lane, valid in lane 0 (the emulation sums the shared slots directly)
LAT_DEV double lat_warp_sum(double v, const float *red, int nw, int q, int lane) {
#ifndef LAT_EMU
    (void)red; (void)nw; (void)q; (void)lane;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
#else
    double sum = 0.0;
    if (lane == 0)
        for (int w = 0; w < nw; ++w) sum += (double)red[w * 3 + q];
    (void)v;
    return sum;
#endif
}

LAT_GLOBAL void k_latent_integrate_r1(LatentP p) {
    LAT_SMEM
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x, nw = (nt + 31) >> 5;
    const int T = p.steps + 1, i = tid;
    float *B0 = (float *)lat_smem, *B1 = B0 + 4 * n;  // ping-pong stage states [4][n]
    float *shs = B1 + 4 * n;                          // source shape (the boundary rows read three of them)
    float *Xs = shs + n;                              // knots of this sample
    float *fcs = Xs + LAT_FAST_NSEQ;                  // [2][LAT_CH][3] source factors, double-buffered by chunk
    float *red = fcs + 2 * 3 * LAT_CH;                // [warps][3]
    const bool act = i < n, inner = act && i > 0 && i < n - 1;
    const float *Yb = p.Y + (size_t)b * p.nseq * n;

    float u[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc[4], ys[4], sg = 0.0f;
    if (act) {
        for (int f = 0; f < 4; ++f) {
            u[f] = p.z0[((size_t)b * 4 + f) * n + i];
            B0[f * n + i] = u[f];
            if (p.z) p.z[((size_t)b * 4 + f) * n + i] = u[f];
        }
        shs[i] = p.shape ? p.shape[(size_t)b * n + i] : 0.0f;
        sg = p.pml_scale * p.pml[(size_t)b * n + i];
    }
    for (int k = tid; k < p.nseq; k += nt) Xs[k] = p.X[(size_t)b * p.nseq + k];
    LAT_SYNC();
    bool mono = true;  // strictly increasing knots: the segment masks exclude one another
    for (int k = 0; k + 1 < p.nseq; ++k) mono = mono && (Xs[k] < Xs[k + 1]);
    const float rend = Xs[p.nseq - 1];
    int kcur = 0;
    float yk = 0.0f, yk1 = 0.0f;
    float lc = 1.0f, rc = 0.0f;  // knots of the segment whose values yk, yk1 are loaded: empty until the first query

    // c0 * C(t) at this thread's element.  Common case (ncu: the three queries were 240 of 800 instructions per step):
    // t lies in the segment of the previous query, whose knots and values are in registers -- then its mask is the only
    // true one (strictly increasing knots) and the result is the reference's expression for that segment.
    auto speed_slow = [&](float t) -> float {
        if (!mono) return lat_speed_elem(p, Xs, Yb, t, i);
        while (kcur + 2 < p.nseq && t >= Xs[kcur + 1]) ++kcur;
        while (kcur > 0 && t < Xs[kcur]) --kcur;
        const float l = Xs[kcur], r = Xs[kcur + 1];
        if (!lat_mask(l, r, rend, t)) return p.c0 * (0.0f + (t - 0.0f) * 0.0f);  // no segment: sums of zeros
        if (l != lc || r != rc) {
            yk = Yb[(size_t)kcur * n + i];
            yk1 = Yb[(size_t)(kcur + 1) * n + i];
            lc = l;
            rc = r;
        }
        const float dd = (r - t) - (l - t);
        const float x0 = 0.0f + l, y0 = 0.0f + yk, dydx = 0.0f + (yk1 - yk) / dd;
        return p.c0 * (y0 + (t - x0) * dydx);
    };
    auto speed = [&](float t) -> float {
        if (!act) return 0.0f;
        if (mono && lc <= t && t < rc) {
            const float dd = (rc - t) - (lc - t);
            const float x0 = 0.0f + lc, y0 = 0.0f + yk, dydx = 0.0f + (yk1 - yk) / dd;
            return p.c0 * (y0 + (t - x0) * dydx);
        }
        return speed_slow(t);
    };
    // The two warps that hold elements 0 and n-1 run a branch-free three-term form of the derivative rows (a per-thread
    // select keeps the two-term value for their interior lanes), so no warp diverges: ncu/timing showed the barriers
    // waiting for those two warps when the boundary rows were a divergent out-of-line path.
    const int wlo = i & ~31;
    const bool edge_warp = wlo == 0 || (wlo <= n - 1 && n - 1 < wlo + 32);
    const bool first = act && i == 0, last = act && i == n - 1;
    const int o0 = first ? 0 : (last ? -2 : -1), o1 = first ? 1 : (last ? -1 : 1), o2 = first ? 2 : 0;
    const float q0 = first ? p.gf[0] : (last ? p.gl[0] : p.gc[0]), q1 = first ? p.gf[1] : (last ? p.gl[1] : p.gc[1]);
    const float q2 = first ? p.gf[2] : (last ? p.gl[2] : 0.0f);
    const float bcv = (first || last) ? 0.0f : 1.0f;
    const float s0 = act ? shs[i + o0] : 0.0f, s1 = act ? shs[i + o1] : 0.0f, s2 = act ? shs[i + o2] : 0.0f;
    auto rhs = [&](const float *S, float a, float fs, const float own[4], float k[4]) {
        const float *Ut = S, *Vt = S + n, *Ui = S + 2 * n, *Vi = S + 3 * n;
        if (!edge_warp && inner) {  // every lane of the warp is interior
            const float fm = s0 * fs, fq = s1 * fs;  // interior lanes: s0 = shape[i-1], s1 = shape[i+1]
            const float gVt = (p.gc[0] * Vt[i - 1]) + (p.gc[1] * Vt[i + 1]);
            const float gUt = (p.gc[0] * (Ut[i - 1] + fm)) + (p.gc[1] * (Ut[i + 1] + fq));
            const float gVi = (p.gc[0] * Vi[i - 1]) + (p.gc[1] * Vi[i + 1]);
            const float gUi = (p.gc[0] * (Ui[i - 1] + fm)) + (p.gc[1] * (Ui[i + 1] + fq));
            k[0] = (a * gVt) - (sg * own[0]);  // bc = 1 inside
            k[1] = (a * gUt) - (sg * own[1]);
            k[2] = (p.c0 * gVi) - (sg * own[2]);
            k[3] = (gUi * p.c0) - (sg * own[3]);
        } else if (act) {
            const bool three = first || last;
            const float f0 = s0 * fs, f1 = s1 * fs, f2 = s2 * fs;
            auto row = [&](float v0, float v1, float v2) {
                const float two = (q0 * v0) + (q1 * v1);
                const float thr = two + (q2 * v2);
                return three ? thr : two;
            };
            const float gVt = row(Vt[i + o0], Vt[i + o1], Vt[i + o2]);
            const float gUt = row(Ut[i + o0] + f0, Ut[i + o1] + f1, Ut[i + o2] + f2);
            const float gVi = row(Vi[i + o0], Vi[i + o1], Vi[i + o2]);
            const float gUi = row(Ui[i + o0] + f0, Ui[i + o1] + f1, Ui[i + o2] + f2);
            k[0] = ((a * gVt) - (sg * own[0])) * bcv;  // x * 1 is exact: one instruction stream for all lanes
            k[1] = (a * gUt) - (sg * own[1]);
            k[2] = ((p.c0 * gVi) - (sg * own[2])) * bcv;
            k[3] = (gUi * p.c0) - (sg * own[3]);
        }
    };

    const float sixth = 1.0f / 6.0f;
    float t0 = p.steps > 0 ? p.tspan[(size_t)b * T] : 0.0f;
    float *zp = (p.z && act) ? p.z + ((size_t)p.batch + b) * 4 * n + i : nullptr;  // this thread's U_tot slot of frame 1
    const size_t zstep = (size_t)p.batch * 4 * n;
    for (int s = 0; s <= p.steps; ++s) {
        // ---- phase 0 (owner-only data, the red slots and the OTHER factor buffer): energy of frame s, speeds of step s
        if (p.energy) {
            const float a = u[0], c = u[2], d = a - c;  // tot .- inc in Float32 first
            float e[3] = {a * a, c * c, d * d};
            lat_reduce3f_warp(e, red, tid);
        }
        float cA = 0.0f, cB = 0.0f, cC = 0.0f;
        if (s < p.steps) {
            if (s % LAT_CH == 0) {  // tabulate the source factors of the next LAT_CH steps
                float *fw = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH;
                for (int q = tid; q < 3 * LAT_CH; q += nt) {
                    const int st = s + q / 3, w = q - 3 * (q / 3);
                    if (st < p.steps) {
                        const float ts = p.tspan[(size_t)b * T + st];
                        fw[q] = p.shape ? lat_sin_factor(w == 0 ? ts : (w == 1 ? ts + p.hdt : ts + p.dt), p.freq) : 0.0f;
                    }
                }
            }
            cA = speed(t0);
            cB = speed(t0 + p.hdt);
            cC = speed(t0 + p.dt);
        }
        LAT_SYNC();
        if (p.energy) {  // Σ over the warps, in double: one warp per quantity (warps 1..3, none of them an edge warp)
            const int w = tid >> 5, lane = tid & 31;
            if (nw >= 5) {
                if (w >= 1 && w <= 3) {
                    const double sum = lat_warp_sum(lane < nw ? (double)red[lane * 3 + (w - 1)] : 0.0, red, nw, w - 1, lane);
                    if (lane == 0) p.energy[((size_t)b * 3 + (w - 1)) * T + s] = (float)sum * p.dx;
                }
            } else if (tid < 3) {
                double sum = 0.0;
                for (int ww = 0; ww < nw; ++ww) sum += (double)red[ww * 3 + tid];
                p.energy[((size_t)b * 3 + tid) * T + s] = (float)sum * p.dx;
            }
        }
        if (s == p.steps) break;
        const float *fc = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH + 3 * (s % LAT_CH);
        const float f0 = fc[0], f1 = fc[1], f2 = fc[2];
        if (s + 1 < p.steps) t0 = p.tspan[(size_t)b * T + s + 1];  // next step's time, off the critical path

        float k[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        // ---- stage 1: reads B0 (= u), writes y2 to B1
        rhs(B0, cA, f0, u, k);
        for (int f = 0; f < 4; ++f) {
            acc[f] = k[f];
            ys[f] = u[f] + (p.hdt * k[f]);
            if (act) B1[f * n + i] = ys[f];
        }
        LAT_SYNC();
        // ---- stage 2: reads B1 (= y2), writes y3 to B0 (u itself lives in registers)
        rhs(B1, cB, f1, ys, k);
        for (int f = 0; f < 4; ++f) {
            acc[f] = acc[f] + (2.0f * k[f]);
            ys[f] = u[f] + (p.hdt * k[f]);
            if (act) B0[f * n + i] = ys[f];
        }
        LAT_SYNC();
        // ---- stage 3: reads B0 (= y3), writes y4 to B1
        rhs(B0, cB, f1, ys, k);
        for (int f = 0; f < 4; ++f) {
            acc[f] = acc[f] + (2.0f * k[f]);
            ys[f] = u[f] + (p.dt * k[f]);
            if (act) B1[f * n + i] = ys[f];
        }
        LAT_SYNC();
        // ---- stage 4: reads B1 (= y4), writes the new state to B0 (its last readers passed two barriers ago)
        rhs(B1, cC, f2, ys, k);
        for (int f = 0; f < 4; ++f) {
            const float du = (sixth * (acc[f] + k[f])) * p.dt;
            u[f] = u[f] + du;
            if (act) B0[f * n + i] = u[f];
        }
        if (zp) {
            for (int f = 0; f < 4; ++f) zp[(size_t)f * n] = u[f];
            zp += zstep;
        }
    }
    if (p.z_last && act)
        for (int f = 0; f < 4; ++f) p.z_last[((size_t)b * 4 + f) * n + i] = u[f];
}

// ------------------------------------------------------------------------------------------------------------------
// Pair variant of the forward fast path (what WAVES_LATENT_AUTO selects for even n <= 1024): every thread owns TWO
// adjacent elements held as float2 and the arithmetic is packed f32x2 (FADD2 / FFMA2: one issue slot per two elements --
// the forward kernels are issue-bound, profiles/r1_latent_ncu_fastpath_summary.json), neighbours cost 8 LDS per two
// elements instead of 16, stores are STS.64, and a barrier joins 16 warps instead of 32.  Multiplies and adds stay
// separately rounded (see lf2_mul below), so it is bit-identical to the other kernels: green on the B200
// (tests/test_gpu_zlatent.py::test_pair_variant_forward_bit_exact), 2.89 against 3.01 us per RK4 step at 148 x 1024
// (profiles/r2_latent_bench.jsonl), static SASS check in profiles/r2_latent_pair_sass_digest.json (scripts/sass_check_latent.py).
// 127 registers at 512 threads: one sample per SM.  n even, 4 <= n <= 1024.
#ifndef LAT_EMU
typedef float2 lf2;
#define LAT_GLOBAL2 __global__ __launch_bounds__(512, 1)
LAT_DEV lf2 lf2_mk(float a, float b) { return make_float2(a, b); }
LAT_DEV lf2 lf2_add(lf2 a, lf2 b) { return __fadd2_rn(a, b); }
LAT_DEV lf2 lf2_sub(lf2 a, lf2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }  // a + (-b) == a - b in IEEE
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad=false (inline PTX included; seen on the
// B200 as a 1-ulp difference).  The product is therefore formed as fma(a, b, -0.0) with the -0.0 read from constant memory:
// exact (x + -0 == x for every x), opaque to ptxas, and an FFMA2 followed by an FADD2 cannot be contracted.
__constant__ float2 lat_negzero = {-0.0f, -0.0f};
LAT_DEV lf2 lf2_mul(lf2 a, lf2 b) { return __ffma2_rn(a, b, lat_negzero); }
#else
struct lf2 {
    float x, y;
};
#define LAT_GLOBAL2 static
LAT_DEV lf2 lf2_mk(float a, float b) { return lf2{a, b}; }
LAT_DEV lf2 lf2_add(lf2 a, lf2 b) { return lf2{a.x + b.x, a.y + b.y}; }
LAT_DEV lf2 lf2_sub(lf2 a, lf2 b) { return lf2{a.x - b.x, a.y - b.y}; }
LAT_DEV lf2 lf2_mul(lf2 a, lf2 b) { return lf2{a.x * b.x, a.y * b.y}; }
#endif
LAT_DEV lf2 lf2_bc(float s) { return lf2_mk(s, s); }

#define LAT_R2_FLOATS(n) LAT_R1_FLOATS(n)
#define LAT_R2_SMEM(n) LAT_R1_SMEM(n)

LAT_GLOBAL2 void k_latent_integrate_r2(LatentP p) {
    LAT_SMEM
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x, nw = (nt + 31) >> 5;
    const int T = p.steps + 1, e0 = 2 * tid, half = n >> 1;
    float *B0 = (float *)lat_smem, *B1 = B0 + 4 * n;
    float *shs = B1 + 4 * n, *Xs = shs + n, *fcs = Xs + LAT_FAST_NSEQ, *red = fcs + 2 * 3 * LAT_CH;
    const bool act = tid < half;
    const float *Yb = p.Y + (size_t)b * p.nseq * n;

    lf2 u[4], acc[4], ys[4], sg = lf2_bc(0.0f), sown = lf2_bc(0.0f);
    for (int f = 0; f < 4; ++f) u[f] = acc[f] = ys[f] = lf2_bc(0.0f);
    if (act) {
        for (int f = 0; f < 4; ++f) {
            const float *src = p.z0 + ((size_t)b * 4 + f) * n + e0;
            u[f] = lf2_mk(src[0], src[1]);
            *(lf2 *)&B0[f * n + e0] = u[f];
            if (p.z) {
                float *dst = p.z + ((size_t)b * 4 + f) * n + e0;
                dst[0] = u[f].x;
                dst[1] = u[f].y;
            }
        }
        if (p.shape) sown = lf2_mk(p.shape[(size_t)b * n + e0], p.shape[(size_t)b * n + e0 + 1]);
        *(lf2 *)&shs[e0] = sown;
        sg = lf2_mul(lf2_bc(p.pml_scale), lf2_mk(p.pml[(size_t)b * n + e0], p.pml[(size_t)b * n + e0 + 1]));
    }
    for (int k = tid; k < p.nseq; k += nt) Xs[k] = p.X[(size_t)b * p.nseq + k];
    LAT_SYNC();
    bool mono = true;
    for (int k = 0; k + 1 < p.nseq; ++k) mono = mono && (Xs[k] < Xs[k + 1]);
    const float rend = Xs[p.nseq - 1];

    // neighbours of the pair: element e0-1 and element e0+2 (clamped for the two boundary threads, whose rows do not use them
    // in that role)
    const bool first = act && tid == 0, last = act && tid == half - 1;
    const int iL = e0 > 0 ? e0 - 1 : 0, iR = e0 + 2 < n ? e0 + 2 : n - 1;
    const lf2 sN = act ? lf2_mk(shs[iL], shs[iR]) : lf2_bc(0.0f);
    const int wlo = tid & ~31;
    const bool edge_warp = wlo == 0 || (wlo <= half - 1 && half - 1 < wlo + 32);
    // rows of dyn.grad for the two elements: coefficients of the operands (X0, X1, X2) chosen in grad() below
    const lf2 A0 = lf2_mk(first ? p.gf[0] : p.gc[0], last ? p.gl[0] : p.gc[0]);
    const lf2 A1 = lf2_mk(first ? p.gf[1] : p.gc[1], last ? p.gl[1] : p.gc[1]);
    const lf2 A2 = lf2_mk(first ? p.gf[2] : 0.0f, last ? p.gl[2] : 0.0f);
    const lf2 bcv = lf2_mk(first ? 0.0f : 1.0f, last ? 0.0f : 1.0f);
    const lf2 g0 = lf2_bc(p.gc[0]), g1 = lf2_bc(p.gc[1]), c02 = lf2_bc(p.c0);
    // (∇ * v) at the two elements: N = (v[e0-1], v[e0+2]), O = (v[e0], v[e0+1])
    auto grad = [&](lf2 N, lf2 O) -> lf2 {
        if (!edge_warp) return lf2_add(lf2_mul(g0, lf2_mk(N.x, O.x)), lf2_mul(g1, lf2_mk(O.y, N.y)));
        const lf2 X0 = lf2_mk(first ? O.x : N.x, last ? N.x : O.x);
        const lf2 X1 = lf2_mk(O.y, last ? O.x : N.y);
        const lf2 X2 = lf2_mk(N.y, O.y);
        const lf2 two = lf2_add(lf2_mul(A0, X0), lf2_mul(A1, X1));
        const lf2 thr = lf2_add(two, lf2_mul(A2, X2));
        return lf2_mk(first ? thr.x : two.x, last ? thr.y : two.y);
    };
    auto rhs = [&](const float *S, lf2 a, float fs, const lf2 own[4], lf2 k[4]) {
        if (!act) return;
        const float *Ut = S, *Vt = S + n, *Ui = S + 2 * n, *Vi = S + 3 * n;
        const lf2 fs2 = lf2_bc(fs);
        const lf2 fN = lf2_mul(sN, fs2), fO = lf2_mul(sown, fs2);
        const lf2 gVt = grad(lf2_mk(Vt[iL], Vt[iR]), own[1]);
        const lf2 gUt = grad(lf2_add(lf2_mk(Ut[iL], Ut[iR]), fN), lf2_add(own[0], fO));  // ∇ * (U_tot .+ f)
        const lf2 gVi = grad(lf2_mk(Vi[iL], Vi[iR]), own[3]);
        const lf2 gUi = grad(lf2_add(lf2_mk(Ui[iL], Ui[iR]), fN), lf2_add(own[2], fO));
        k[0] = lf2_sub(lf2_mul(a, gVt), lf2_mul(sg, own[0]));
        k[1] = lf2_sub(lf2_mul(a, gUt), lf2_mul(sg, own[1]));
        k[2] = lf2_sub(lf2_mul(c02, gVi), lf2_mul(sg, own[2]));
        k[3] = lf2_sub(lf2_mul(gUi, c02), lf2_mul(sg, own[3]));
        if (edge_warp) {  // bc: x * 1 is exact for the interior lanes
            k[0] = lf2_mul(k[0], bcv);
            k[2] = lf2_mul(k[2], bcv);
        }
    };

    // c0 * C(t) at the two elements, as in k_latent_integrate_r1
    int kcur = 0;
    lf2 yk = lf2_bc(0.0f), yk1 = lf2_bc(0.0f);
    float lc = 1.0f, rc = 0.0f;
    auto speed = [&](float t) -> lf2 {
        if (!act) return lf2_bc(0.0f);
        if (!mono) return lf2_mk(lat_speed_elem(p, Xs, Yb, t, e0), lat_speed_elem(p, Xs, Yb, t, e0 + 1));
        if (!(lc <= t && t < rc)) {
            while (kcur + 2 < p.nseq && t >= Xs[kcur + 1]) ++kcur;
            while (kcur > 0 && t < Xs[kcur]) --kcur;
            const float l = Xs[kcur], r = Xs[kcur + 1];
            if (!lat_mask(l, r, rend, t)) return lf2_bc(p.c0 * (0.0f + (t - 0.0f) * 0.0f));
            if (l != lc || r != rc) {
                yk = lf2_mk(Yb[(size_t)kcur * n + e0], Yb[(size_t)kcur * n + e0 + 1]);
                yk1 = lf2_mk(Yb[(size_t)(kcur + 1) * n + e0], Yb[(size_t)(kcur + 1) * n + e0 + 1]);
                lc = l;
                rc = r;
            }
        }
        const float dd = (rc - t) - (lc - t), x0 = 0.0f + lc;
        const lf2 dy = lf2_sub(yk1, yk);
        const lf2 dydx = lf2_add(lf2_bc(0.0f), lf2_mk(dy.x / dd, dy.y / dd));
        const lf2 y0 = lf2_add(lf2_bc(0.0f), yk);
        return lf2_mul(c02, lf2_add(y0, lf2_mul(lf2_bc(t - x0), dydx)));
    };

    const lf2 hdt2 = lf2_bc(p.hdt), dt2 = lf2_bc(p.dt), two2 = lf2_bc(2.0f), sixth2 = lf2_bc(1.0f / 6.0f);
    float t0 = p.steps > 0 ? p.tspan[(size_t)b * T] : 0.0f;
    float *zp = (p.z && act) ? p.z + ((size_t)p.batch + b) * 4 * n + e0 : nullptr;
    const size_t zstep = (size_t)p.batch * 4 * n;
    for (int s = 0; s <= p.steps; ++s) {
        if (p.energy) {
            const lf2 d = lf2_sub(u[0], u[2]);  // tot .- inc in Float32 first
            const lf2 ea = lf2_mul(u[0], u[0]), ec = lf2_mul(u[2], u[2]), ed = lf2_mul(d, d);
            float e[3] = {ea.x + ea.y, ec.x + ec.y, ed.x + ed.y};
            lat_reduce3f_warp(e, red, tid);
        }
        lf2 cA = lf2_bc(0.0f), cB = cA, cC = cA;
        if (s < p.steps) {
            if (s % LAT_CH == 0) {
                float *fw = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH;
                for (int q = tid; q < 3 * LAT_CH; q += nt) {
                    const int st = s + q / 3, w = q - 3 * (q / 3);
                    if (st < p.steps) {
                        const float ts = p.tspan[(size_t)b * T + st];
                        fw[q] = p.shape ? lat_sin_factor(w == 0 ? ts : (w == 1 ? ts + p.hdt : ts + p.dt), p.freq) : 0.0f;
                    }
                }
            }
            cA = speed(t0);
            cB = speed(t0 + p.hdt);
            cC = speed(t0 + p.dt);
        }
        LAT_SYNC();
        if (p.energy) {
            const int w = tid >> 5, lane = tid & 31;
            if (nw >= 5) {
                if (w >= 1 && w <= 3) {
                    const double sum = lat_warp_sum(lane < nw ? (double)red[lane * 3 + (w - 1)] : 0.0, red, nw, w - 1, lane);
                    if (lane == 0) p.energy[((size_t)b * 3 + (w - 1)) * T + s] = (float)sum * p.dx;
                }
            } else if (tid < 3) {
                double sum = 0.0;
                for (int ww = 0; ww < nw; ++ww) sum += (double)red[ww * 3 + tid];
                p.energy[((size_t)b * 3 + tid) * T + s] = (float)sum * p.dx;
            }
        }
        if (s == p.steps) break;
        const float *fc = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH + 3 * (s % LAT_CH);
        const float f0 = fc[0], f1 = fc[1], f2 = fc[2];
        if (s + 1 < p.steps) t0 = p.tspan[(size_t)b * T + s + 1];

        lf2 k[4];
        for (int f = 0; f < 4; ++f) k[f] = lf2_bc(0.0f);
        rhs(B0, cA, f0, u, k);  // stage 1: reads B0 (= u), writes y2 to B1
        for (int f = 0; f < 4; ++f) {
            acc[f] = k[f];
            ys[f] = lf2_add(u[f], lf2_mul(hdt2, k[f]));
            if (act) *(lf2 *)&B1[f * n + e0] = ys[f];
        }
        LAT_SYNC();
        rhs(B1, cB, f1, ys, k);  // stage 2: reads B1 (= y2), writes y3 to B0
        for (int f = 0; f < 4; ++f) {
            acc[f] = lf2_add(acc[f], lf2_mul(two2, k[f]));
            ys[f] = lf2_add(u[f], lf2_mul(hdt2, k[f]));
            if (act) *(lf2 *)&B0[f * n + e0] = ys[f];
        }
        LAT_SYNC();
        rhs(B0, cB, f1, ys, k);  // stage 3: reads B0 (= y3), writes y4 to B1
        for (int f = 0; f < 4; ++f) {
            acc[f] = lf2_add(acc[f], lf2_mul(two2, k[f]));
            ys[f] = lf2_add(u[f], lf2_mul(dt2, k[f]));
            if (act) *(lf2 *)&B1[f * n + e0] = ys[f];
        }
        LAT_SYNC();
        rhs(B1, cC, f2, ys, k);  // stage 4: reads B1 (= y4), writes the new state to B0
        for (int f = 0; f < 4; ++f) {
            const lf2 du = lf2_mul(lf2_mul(sixth2, lf2_add(acc[f], k[f])), dt2);
            u[f] = lf2_add(u[f], du);
            if (act) *(lf2 *)&B0[f * n + e0] = u[f];
        }
        if (zp) {
            for (int f = 0; f < 4; ++f) {
                zp[(size_t)f * n] = u[f].x;
                zp[(size_t)f * n + 1] = u[f].y;
            }
            zp += zstep;
        }
    }
    if (p.z_last && act)
        for (int f = 0; f < 4; ++f) {
            p.z_last[((size_t)b * 4 + f) * n + e0] = u[f].x;
            p.z_last[((size_t)b * 4 + f) * n + e0 + 1] = u[f].y;
        }
}

// ------------------------------------------------------------------------------------------------------------------
// reverse pass: adjoint_sensitivity (src/dynamics.jl:97-118) for the batchwise OneDim simulation.
//   exact  : λ_N = a_N;  λ_i = a_i + (I + J_iᵀ) λ_{i+1}
//   compat : acc = 0; for i = N..0: acc = (I + J_iᵀ)(acc + a_i)          (the loop as written)
// a_i = dL_dz[i] + d(Σ w_energy · energy)/dz_i.  One CTA per sample; per stored state z_i the three forward stage
// states are recomputed in shared memory and four transposed right-hand sides are applied; parameter gradients
// (C.Y through the interpolation weights, F.shape, PML) are accumulated by the owner thread of every element.
#define LAT_ADJ_FLOATS(n) (39 * (size_t)(n) + 8)
#define LAT_ADJ_SMEM(n) (sizeof(float) * LAT_ADJ_FLOATS(n))

// vjp of one right-hand side: kb = cotangent of its output (shared, [4][n]), S = the stage state it was evaluated at.
// Returns the cotangent of the stage state at element j in yb[4] and accumulates the parameter gradients of element j.
LAT_DEV void lat_rhs_vjp(const LatentP &p, int b, const float *kb, const float *S, const float *c0c, const float *shp,
                         const float *sig, float fs, float t, int j, float yb[4], float *gshp, float *gsig) {
    const int n = p.n;
    const float *Ut = S, *Vt = S + n, *Ui = S + 2 * n, *Vi = S + 3 * n;
    auto bcf = [&](int i) { return (i == 0 || i == n - 1) ? 0.0f : 1.0f; };
    const float bj = bcf(j), sj = sig[j];
    const float tp = lat_grad_t(p, j, [&](int i) { return c0c[i] * kb[n + i]; });           // ∇ᵀ(a ⊙ w_Vtot)
    const float tq = lat_grad_t(p, j, [&](int i) { return c0c[i] * bcf(i) * kb[i]; });      // ∇ᵀ(a ⊙ bc ⊙ w_Utot)
    const float tr = lat_grad_t(p, j, [&](int i) { return kb[3 * n + i]; });                // ∇ᵀ(w_Vinc)
    const float ts = lat_grad_t(p, j, [&](int i) { return bcf(i) * kb[2 * n + i]; });       // ∇ᵀ(bc ⊙ w_Uinc)
    yb[0] = tp - sj * bj * kb[j];
    yb[1] = tq - sj * kb[n + j];
    yb[2] = p.c0 * tr - sj * bj * kb[2 * n + j];
    yb[3] = p.c0 * ts - sj * kb[3 * n + j];
    // parameters
    if (gshp) gshp[j] += (tp + p.c0 * tr) * fs;  // f enters as ∇(U + f) of both wavefields
    if (gsig) gsig[j] -= bj * kb[j] * Ut[j] + kb[n + j] * Vt[j] + bj * kb[2 * n + j] * Ui[j] + kb[3 * n + j] * Vi[j];
    if (p.g_Y) {
        const float gVt = lat_grad(p, j, [&](int i) { return Vt[i]; });
        const float gUt = lat_grad(p, j, [&](int i) { return Ut[i] + shp[i] * fs; });
        const float cbar = p.c0 * (bj * kb[j] * gVt + kb[n + j] * gUt);  // dL/dC(t)[j]
        const float *Xb = p.X + (size_t)b * p.nseq;
        float *gY = p.g_Y + (size_t)b * p.nseq * n;
        const float rend = Xb[p.nseq - 1];
        float x0 = 0.0f;
        for (int k = 0; k + 1 < p.nseq; ++k)
            if (lat_mask(Xb[k], Xb[k + 1], rend, t)) x0 = x0 + Xb[k];
        for (int k = 0; k + 1 < p.nseq; ++k) {
            const float l = Xb[k], r = Xb[k + 1];
            if (lat_mask(l, r, rend, t)) {
                const float wk = (t - x0) / ((r - t) - (l - t));
                gY[(size_t)k * n + j] += cbar * (1.0f - wk);   // owner-only read-modify-write: no atomics needed
                gY[(size_t)(k + 1) * n + j] += cbar * wk;
            }
        }
    }
}

LAT_GLOBAL void k_latent_adjoint(LatentP p) {
    LAT_SMEM
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x;
    const int T = p.steps + 1, N = p.steps;
    float *zs = (float *)lat_smem;                       // z_i                        (neighbour reads)
    float *y2 = zs + 4 * n, *y3 = y2 + 4 * n, *y4 = y3 + 4 * n;
    float *kbA = y4 + 4 * n, *kbB = kbA + 4 * n;         // stage cotangents, double-buffered (neighbour reads)
    float *lam = kbB + 4 * n, *zsum = lam + 4 * n;       // owner-only
    float *cA = zsum + 4 * n, *cB = cA + n, *cC = cB + n, *shp = cC + n, *sig = shp + n;
    float *gshp = sig + n, *gsig = gshp + n;             // owner-only accumulators
    float *fac = gsig + n;                               // [2][4]

    for (int i = tid; i < n; i += nt) {
        shp[i] = p.shape ? p.shape[(size_t)b * n + i] : 0.0f;
        sig[i] = p.pml_scale * p.pml[(size_t)b * n + i];
        gshp[i] = 0.0f;
        gsig[i] = 0.0f;
        for (int f = 0; f < 4; ++f) lam[f * n + i] = 0.0f;
    }
    // a_i at the elements this thread owns, added to lam; zi = state i (global)
    auto add_cotangent = [&](int i) {
        const float *zi = p.zt + ((size_t)i * p.batch + b) * 4 * n;
        float w3[3] = {0.0f, 0.0f, 0.0f};
        if (p.w_energy)
            for (int q = 0; q < 3; ++q) w3[q] = p.w_energy[((size_t)b * 3 + q) * T + i];
        for (int e = tid; e < n; e += nt) {
            if (p.w_energy) {
                const float ut = zi[e], ui = zi[2 * n + e], d = ut - ui;
                lam[e] += 2.0f * p.dx * (w3[0] * ut + w3[2] * d);
                lam[2 * n + e] += 2.0f * p.dx * (w3[1] * ui - w3[2] * d);
            }
            if (p.dL_dz)
                for (int f = 0; f < 4; ++f) lam[f * n + e] += p.dL_dz[(((size_t)i * p.batch + b) * 4 + f) * n + e];
        }
    };
    if (!p.compat) add_cotangent(N);  // λ_N = a_N

    const float dt = p.dt, hdt = p.hdt, sixth = 1.0f / 6.0f;
    int it = 0;
    for (int i = p.compat ? N : N - 1; i >= 0; --i, ++it) {
        if (p.compat) add_cotangent(i);  // acc += a_i before the step vjp (src/dynamics.jl:110)
        float *fc = fac + (it & 1) * 4;
        const float t0 = p.tspan[(size_t)b * T + i];
        const float t1 = t0 + hdt, t2 = t0 + dt;
        lat_speed(p, b, t0, cA, tid, nt);
        lat_speed(p, b, t1, cB, tid, nt);
        lat_speed(p, b, t2, cC, tid, nt);
        if (tid < 3) fc[tid] = p.shape ? lat_sin_factor(tid == 0 ? t0 : (tid == 1 ? t1 : t2), p.freq) : 0.0f;
        const float *zi = p.zt + ((size_t)i * p.batch + b) * 4 * n;
        for (int e = tid; e < n; e += nt)
            for (int f = 0; f < 4; ++f) zs[f * n + e] = zi[f * n + e];
        LAT_SYNC();
        // ---- forward stage states y2 = z + dt/2 k1(z), y3 = z + dt/2 k2(y2), y4 = z + dt k3(y3)
        for (int e = tid; e < n; e += nt) {
            float k[4];
            lat_rhs(p, zs, cA, shp, sig, fc[0], e, k);
            for (int f = 0; f < 4; ++f) y2[f * n + e] = zs[f * n + e] + (hdt * k[f]);
        }
        LAT_SYNC();
        for (int e = tid; e < n; e += nt) {
            float k[4];
            lat_rhs(p, y2, cB, shp, sig, fc[1], e, k);
            for (int f = 0; f < 4; ++f) y3[f * n + e] = zs[f * n + e] + (hdt * k[f]);
        }
        LAT_SYNC();
        for (int e = tid; e < n; e += nt) {
            float k[4];
            lat_rhs(p, y3, cB, shp, sig, fc[1], e, k);
            for (int f = 0; f < 4; ++f) {
                y4[f * n + e] = zs[f * n + e] + (dt * k[f]);
                kbA[f * n + e] = (sixth * dt) * lam[f * n + e];  // cotangent of k4
            }
        }
        LAT_SYNC();
        // ---- reverse stages: ȳ4 = J4ᵀ k̄4; k̄3 = dt/3 λ + dt ȳ4; ȳ3 = J3ᵀ k̄3; k̄2 = dt/3 λ + dt/2 ȳ3; ȳ2 = J2ᵀ k̄2;
        //      k̄1 = dt/6 λ + dt/2 ȳ2; ȳ1 = J1ᵀ k̄1;  δz = ȳ1 + ȳ2 + ȳ3 + ȳ4
        for (int e = tid; e < n; e += nt) {
            float yb[4];
            lat_rhs_vjp(p, b, kbA, y4, cC, shp, sig, fc[2], t2, e, yb, p.g_shape ? gshp : nullptr, p.g_pml ? gsig : nullptr);
            for (int f = 0; f < 4; ++f) {
                zsum[f * n + e] = yb[f];
                kbB[f * n + e] = (2.0f * sixth * dt) * lam[f * n + e] + dt * yb[f];
            }
        }
        LAT_SYNC();
        for (int e = tid; e < n; e += nt) {
            float yb[4];
            lat_rhs_vjp(p, b, kbB, y3, cB, shp, sig, fc[1], t1, e, yb, p.g_shape ? gshp : nullptr, p.g_pml ? gsig : nullptr);
            for (int f = 0; f < 4; ++f) {
                zsum[f * n + e] += yb[f];
                kbA[f * n + e] = (2.0f * sixth * dt) * lam[f * n + e] + hdt * yb[f];
            }
        }
        LAT_SYNC();
        for (int e = tid; e < n; e += nt) {
            float yb[4];
            lat_rhs_vjp(p, b, kbA, y2, cB, shp, sig, fc[1], t1, e, yb, p.g_shape ? gshp : nullptr, p.g_pml ? gsig : nullptr);
            for (int f = 0; f < 4; ++f) {
                zsum[f * n + e] += yb[f];
                kbB[f * n + e] = (sixth * dt) * lam[f * n + e] + hdt * yb[f];
            }
        }
        LAT_SYNC();
        for (int e = tid; e < n; e += nt) {
            float yb[4];
            lat_rhs_vjp(p, b, kbB, zs, cA, shp, sig, fc[0], t0, e, yb, p.g_shape ? gshp : nullptr, p.g_pml ? gsig : nullptr);
            for (int f = 0; f < 4; ++f) lam[f * n + e] += zsum[f * n + e] + yb[f];  // λ <- λ + δz
        }
        if (!p.compat) add_cotangent(i);  // λ_i = a_i + (I + J_iᵀ) λ_{i+1}
        LAT_SYNC();  // zs, kbB and the stage states are rewritten by the next iteration
    }
    for (int e = tid; e < n; e += nt) {
        for (int f = 0; f < 4; ++f) p.g_z0[((size_t)b * 4 + f) * n + e] = lam[f * n + e];
        if (p.g_shape) p.g_shape[(size_t)b * n + e] = gshp[e];
        if (p.g_pml) p.g_pml[(size_t)b * n + e] = p.pml_scale * gsig[e];  // σ = pml_scale .* PML
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Register fast path of the reverse kernel: n <= 1024, one element per thread (the layout of k_latent_integrate_r1).
// ncu on the generic reverse kernel (profiles/r1_latent_ncu_summary.json): issue-bound, ~3060 warp instructions per warp
// per step.  Here λ, the step sum, the stage cotangent and the parameter-gradient accumulators are registers; shared memory
// holds what neighbours read: the stored state (double-buffered by step so no barrier is needed at the end of a step), the
// three recomputed stage states, and two ping-pong buffers with the PRODUCTS the transposed rows consume
// (c0·C ⊙ bc ⊙ w_U, c0·C ⊙ w_V, bc ⊙ w_Uinc, w_Vinc) so no speed array has to be shared; ∂L/∂C.Y of the current segment
// accumulates in two registers and is flushed to HBM only when the stage times move to another segment.
// Eight barriers per step.  Same mathematics as k_latent_adjoint (float32 accumulation order differs).
#define LAT_ADJ1_FLOATS(n) (29 * (size_t)(n) + LAT_FAST_NSEQ + 2 * 3 * LAT_CH + 8)
#define LAT_ADJ1_SMEM(n) (sizeof(float) * LAT_ADJ1_FLOATS(n))

LAT_GLOBAL void k_latent_adjoint_r1(LatentP p) {
    LAT_SMEM
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x, i = tid;
    const int T = p.steps + 1, N = p.steps;
    float *ZA = (float *)lat_smem, *ZB = ZA + 4 * n;     // stored state z_i, double-buffered by iteration parity
    float *Y2 = ZB + 4 * n, *Y3 = Y2 + 4 * n, *Y4 = Y3 + 4 * n;
    float *PA = Y4 + 4 * n, *PB = PA + 4 * n;            // products read by the transposed rows
    float *shs = PB + 4 * n, *Xs = shs + n, *fcs = Xs + LAT_FAST_NSEQ;
    const bool act = i < n, inner = act && i > 0 && i < n - 1;
    const float *Yb = p.Y + (size_t)b * p.nseq * n;
    float *gY = p.g_Y ? p.g_Y + (size_t)b * p.nseq * n : nullptr;

    float sg = 0.0f;
    if (act) {
        shs[i] = p.shape ? p.shape[(size_t)b * n + i] : 0.0f;
        sg = p.pml_scale * p.pml[(size_t)b * n + i];
    }
    for (int k = tid; k < p.nseq; k += nt) Xs[k] = p.X[(size_t)b * p.nseq + k];
    LAT_SYNC();
    bool mono = true;
    for (int k = 0; k + 1 < p.nseq; ++k) mono = mono && (Xs[k] < Xs[k + 1]);
    const float rend = Xs[p.nseq - 1];

    // per-lane form of the derivative rows, as in k_latent_integrate_r1
    const int wlo = i & ~31;
    const bool edge_warp = wlo == 0 || (wlo <= n - 1 && n - 1 < wlo + 32);
    const bool first = act && i == 0, last = act && i == n - 1, three = first || last;
    const int o0 = first ? 0 : (last ? -2 : -1), o1 = first ? 1 : (last ? -1 : 1), o2 = first ? 2 : 0;
    const float q0 = first ? p.gf[0] : (last ? p.gl[0] : p.gc[0]), q1 = first ? p.gf[1] : (last ? p.gl[1] : p.gc[1]);
    const float q2 = first ? p.gf[2] : (last ? p.gl[2] : 0.0f);
    const float bcv = three ? 0.0f : 1.0f;
    // the TRANSPOSED rows differ from the central form at elements 0, 1, 2 and n-3, n-2, n-1 (the one-sided rows reach them)
    const bool edge_t = wlo == 0 || wlo + 31 >= n - 3;
    const float s0 = act ? shs[i + o0] : 0.0f, s1 = act ? shs[i + o1] : 0.0f, s2 = act ? shs[i + o2] : 0.0f;
    auto row = [&](float v0, float v1, float v2) {
        const float two = (q0 * v0) + (q1 * v1);
        return three ? two + (q2 * v2) : two;
    };
    // forward right-hand side at this thread's element (reference order, as in the forward kernel)
    auto rhs = [&](const float *S, float a, float fs, const float own[4], float k[4]) {
        if (!act) return;
        const float *Ut = S, *Vt = S + n, *Ui = S + 2 * n, *Vi = S + 3 * n;
        const float f0 = s0 * fs, f1 = s1 * fs, f2 = s2 * fs;
        float gVt, gUt, gVi, gUi;
        if (!edge_warp) {
            gVt = (p.gc[0] * Vt[i - 1]) + (p.gc[1] * Vt[i + 1]);
            gUt = (p.gc[0] * (Ut[i - 1] + f0)) + (p.gc[1] * (Ut[i + 1] + f1));
            gVi = (p.gc[0] * Vi[i - 1]) + (p.gc[1] * Vi[i + 1]);
            gUi = (p.gc[0] * (Ui[i - 1] + f0)) + (p.gc[1] * (Ui[i + 1] + f1));
        } else {
            gVt = row(Vt[i + o0], Vt[i + o1], Vt[i + o2]);
            gUt = row(Ut[i + o0] + f0, Ut[i + o1] + f1, Ut[i + o2] + f2);
            gVi = row(Vi[i + o0], Vi[i + o1], Vi[i + o2]);
            gUi = row(Ui[i + o0] + f0, Ui[i + o1] + f1, Ui[i + o2] + f2);
        }
        k[0] = ((a * gVt) - (sg * own[0])) * bcv;
        k[1] = (a * gUt) - (sg * own[1]);
        k[2] = ((p.c0 * gVi) - (sg * own[2])) * bcv;
        k[3] = (gUi * p.c0) - (sg * own[3]);
    };

    // C(t): value c0*C(t)[i], segment index (-1: no segment) and interpolation weight of its right knot
    int kcur = 0, kreg = -1;  // search cursor; segment whose knots (lc, rc) and values (yk, yk1) are in registers
    float yk = 0.0f, yk1 = 0.0f, lc = 1.0f, rc = 0.0f;
    auto query = [&](float t, float &c, int &kseg, float &w) {
        kseg = -1;
        w = 0.0f;
        c = 0.0f;
        if (!act) return;
        if (!mono) {
            c = lat_speed_elem(p, Xs, Yb, t, i);
            return;
        }
        if (!(lc <= t && t < rc)) {
            while (kcur + 2 < p.nseq && t >= Xs[kcur + 1]) ++kcur;
            while (kcur > 0 && t < Xs[kcur]) --kcur;
            const float l = Xs[kcur], r = Xs[kcur + 1];
            if (!lat_mask(l, r, rend, t)) {
                c = p.c0 * (0.0f + (t - 0.0f) * 0.0f);
                return;
            }
            if (kreg != kcur) {
                yk = Yb[(size_t)kcur * n + i];
                yk1 = Yb[(size_t)(kcur + 1) * n + i];
                lc = l;
                rc = r;
                kreg = kcur;
            }
        }
        const float dd = (rc - t) - (lc - t);
        const float x0 = 0.0f + lc, y0 = 0.0f + yk, dydx = 0.0f + (yk1 - yk) / dd;
        c = p.c0 * (y0 + (t - x0) * dydx);
        kseg = kreg;
        w = (t - x0) / dd;
    };

    // ∂L/∂C.Y of segment kacc: coefficients of Y[kacc] and Y[kacc+1] at this element
    int kacc = -1;
    float gYa = 0.0f, gYb = 0.0f;
    auto flush_gy = [&]() {
        if (gY && act && kacc >= 0) {
            gY[(size_t)kacc * n + i] += gYa;  // owner-only read-modify-write
            gY[(size_t)(kacc + 1) * n + i] += gYb;
        }
        gYa = 0.0f;
        gYb = 0.0f;
    };
    auto add_gy = [&](float cbar, float t, int kseg, float w) {
        if (!gY || !act) return;
        if (!mono) {  // knots that are not increasing: every true mask contributes (as in k_latent_adjoint)
            float x0 = 0.0f;
            for (int k = 0; k + 1 < p.nseq; ++k)
                if (lat_mask(Xs[k], Xs[k + 1], rend, t)) x0 = x0 + Xs[k];
            for (int k = 0; k + 1 < p.nseq; ++k) {
                const float l = Xs[k], r = Xs[k + 1];
                if (lat_mask(l, r, rend, t)) {
                    const float wk = (t - x0) / ((r - t) - (l - t));
                    gY[(size_t)k * n + i] += cbar * (1.0f - wk);
                    gY[(size_t)(k + 1) * n + i] += cbar * wk;
                }
            }
            return;
        }
        if (kseg < 0) return;
        if (kseg != kacc) {
            flush_gy();
            kacc = kseg;
        }
        gYa += cbar * (1.0f - w);
        gYb += cbar * w;
    };

    float lam[4] = {0.0f, 0.0f, 0.0f, 0.0f}, zr[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float gshp = 0.0f, gsig = 0.0f;
    // λ += a_idx at this element; zv = z_idx at this element
    auto add_cotangent = [&](int idx, const float zv[4]) {
        if (!act) return;
        if (p.w_energy) {
            const float w0 = p.w_energy[((size_t)b * 3 + 0) * T + idx], w1 = p.w_energy[((size_t)b * 3 + 1) * T + idx];
            const float w2 = p.w_energy[((size_t)b * 3 + 2) * T + idx];
            const float ut = zv[0], ui = zv[2], d = ut - ui;
            lam[0] += 2.0f * p.dx * (w0 * ut + w2 * d);
            lam[2] += 2.0f * p.dx * (w1 * ui - w2 * d);
        }
        if (p.dL_dz)
            for (int f = 0; f < 4; ++f) lam[f] += p.dL_dz[(((size_t)idx * p.batch + b) * 4 + f) * n + i];
    };
    auto load_state = [&](int idx) {
        if (!act) return;
        const float *zi = p.zt + ((size_t)idx * p.batch + b) * 4 * n;
        for (int f = 0; f < 4; ++f) zr[f] = zi[(size_t)f * n + i];
    };
    // vjp of one right-hand side.  Pin: the products written by the previous stage; S: the stage state; kb: own cotangent
    auto vjp = [&](const float *S, const float *Pin, float a, float fs, float t, int kseg, float w, const float kb[4],
                   float yb[4]) {
        if (!act) return;
        const float *Q0 = Pin, *Q1 = Pin + n, *Q2 = Pin + 2 * n, *Q3 = Pin + 3 * n;
        float tp, tq, tr, ts;
        if (!edge_t) {  // 3 <= i <= n-4 for every lane of this warp: rows i-1 and i+1 are central, no one-sided row reaches i
            tp = (p.gc[0] * Q1[i + 1]) + (p.gc[1] * Q1[i - 1]);
            tq = (p.gc[0] * Q0[i + 1]) + (p.gc[1] * Q0[i - 1]);
            tr = (p.gc[0] * Q3[i + 1]) + (p.gc[1] * Q3[i - 1]);
            ts = (p.gc[0] * Q2[i + 1]) + (p.gc[1] * Q2[i - 1]);
        } else {
            tp = lat_grad_t(p, i, [&](int j) { return Q1[j]; });
            tq = lat_grad_t(p, i, [&](int j) { return Q0[j]; });
            tr = lat_grad_t(p, i, [&](int j) { return Q3[j]; });
            ts = lat_grad_t(p, i, [&](int j) { return Q2[j]; });
        }
        yb[0] = tp - sg * bcv * kb[0];
        yb[1] = tq - sg * kb[1];
        yb[2] = p.c0 * tr - sg * bcv * kb[2];
        yb[3] = p.c0 * ts - sg * kb[3];
        const float *Ut = S, *Vt = S + n, *Ui = S + 2 * n, *Vi = S + 3 * n;
        if (p.g_shape) gshp += (tp + p.c0 * tr) * fs;
        if (p.g_pml) gsig -= bcv * kb[0] * Ut[i] + kb[1] * Vt[i] + bcv * kb[2] * Ui[i] + kb[3] * Vi[i];
        if (gY) {
            const float f0 = s0 * fs, f1 = s1 * fs, f2 = s2 * fs;
            float gVt, gUt;
            if (!edge_warp) {
                gVt = (p.gc[0] * Vt[i - 1]) + (p.gc[1] * Vt[i + 1]);
                gUt = (p.gc[0] * (Ut[i - 1] + f0)) + (p.gc[1] * (Ut[i + 1] + f1));
            } else {
                gVt = row(Vt[i + o0], Vt[i + o1], Vt[i + o2]);
                gUt = row(Ut[i + o0] + f0, Ut[i + o1] + f1, Ut[i + o2] + f2);
            }
            add_gy(p.c0 * (bcv * kb[0] * gVt + kb[1] * gUt), t, kseg, w);
        }
        (void)a;
    };
    // the products the NEXT transposed rows read, for the stage whose speed is a
    auto put_products = [&](float *P, float a, const float kb[4]) {
        if (!act) return;
        P[i] = a * bcv * kb[0];
        P[n + i] = a * kb[1];
        P[2 * n + i] = bcv * kb[2];
        P[3 * n + i] = kb[3];
    };

    if (!p.compat) {  // λ_N = a_N
        load_state(N);
        add_cotangent(N, zr);
    }
    const float dt = p.dt, hdt = p.hdt, sixth = 1.0f / 6.0f;
    const int istart = p.compat ? N : N - 1;
    int it = 0;
    for (int s = istart; s >= 0; --s, ++it) {
        float *Z = (it & 1) ? ZB : ZA;
        // ---- phase 0: state, cotangent (compat), the three speeds, the table of source factors
        load_state(s);
        if (p.compat) add_cotangent(s, zr);  // acc += a_i before the step vjp (src/dynamics.jl:110)
        if (s == istart || (s % LAT_CH) == LAT_CH - 1) {
            const int c0s = (s / LAT_CH) * LAT_CH;
            float *fw = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH;
            for (int q = tid; q < 3 * LAT_CH; q += nt) {
                const int st = c0s + q / 3, w = q - 3 * (q / 3);
                if (st <= s) {
                    const float ts = p.tspan[(size_t)b * T + st];
                    fw[q] = p.shape ? lat_sin_factor(w == 0 ? ts : (w == 1 ? ts + hdt : ts + dt), p.freq) : 0.0f;
                }
            }
        }
        const float t0 = p.tspan[(size_t)b * T + s], t1 = t0 + hdt, t2 = t0 + dt;
        float cA, cB, cC, wA, wB, wC;
        int kA, kB, kC;
        query(t2, cC, kC, wC);  // time-decreasing order, like the sweep
        query(t1, cB, kB, wB);
        query(t0, cA, kA, wA);
        if (act)
            for (int f = 0; f < 4; ++f) Z[f * n + i] = zr[f];
        LAT_SYNC();
        const float *fc = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH + 3 * (s % LAT_CH);
        const float f0 = fc[0], f1 = fc[1], f2 = fc[2];

        float k[4] = {0.0f, 0.0f, 0.0f, 0.0f}, ys[4] = {0.0f, 0.0f, 0.0f, 0.0f}, kb[4], yb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float zsum[4];
        // ---- forward stage states y2 = z + dt/2 k1(z), y3 = z + dt/2 k2(y2), y4 = z + dt k3(y3)
        rhs(Z, cA, f0, zr, k);
        for (int f = 0; f < 4; ++f) {
            ys[f] = zr[f] + (hdt * k[f]);
            if (act) Y2[f * n + i] = ys[f];
        }
        LAT_SYNC();
        rhs(Y2, cB, f1, ys, k);
        for (int f = 0; f < 4; ++f) {
            ys[f] = zr[f] + (hdt * k[f]);
            if (act) Y3[f * n + i] = ys[f];
        }
        LAT_SYNC();
        rhs(Y3, cB, f1, ys, k);
        for (int f = 0; f < 4; ++f) {
            ys[f] = zr[f] + (dt * k[f]);
            if (act) Y4[f * n + i] = ys[f];
            kb[f] = (sixth * dt) * lam[f];  // cotangent of k4
        }
        put_products(PA, cC, kb);
        LAT_SYNC();
        // ---- reverse stages: ȳ4 = J4ᵀ k̄4; k̄3 = dt/3 λ + dt ȳ4; ȳ3 = J3ᵀ k̄3; k̄2 = dt/3 λ + dt/2 ȳ3; ȳ2 = J2ᵀ k̄2;
        //      k̄1 = dt/6 λ + dt/2 ȳ2; ȳ1 = J1ᵀ k̄1;  δz = ȳ1 + ȳ2 + ȳ3 + ȳ4
        vjp(Y4, PA, cC, f2, t2, kC, wC, kb, yb);
        for (int f = 0; f < 4; ++f) {
            zsum[f] = yb[f];
            kb[f] = (2.0f * sixth * dt) * lam[f] + dt * yb[f];
        }
        put_products(PB, cB, kb);
        LAT_SYNC();
        vjp(Y3, PB, cB, f1, t1, kB, wB, kb, yb);
        for (int f = 0; f < 4; ++f) {
            zsum[f] += yb[f];
            kb[f] = (2.0f * sixth * dt) * lam[f] + hdt * yb[f];
        }
        put_products(PA, cB, kb);
        LAT_SYNC();
        vjp(Y2, PA, cB, f1, t1, kB, wB, kb, yb);
        for (int f = 0; f < 4; ++f) {
            zsum[f] += yb[f];
            kb[f] = (sixth * dt) * lam[f] + hdt * yb[f];
        }
        put_products(PB, cA, kb);
        LAT_SYNC();
        vjp(Z, PB, cA, f0, t0, kA, wA, kb, yb);
        for (int f = 0; f < 4; ++f) lam[f] += zsum[f] + yb[f];  // λ <- λ + δz
        if (!p.compat) add_cotangent(s, zr);                     // λ_i = a_i + (I + J_iᵀ) λ_{i+1}
        // no barrier: the next iteration writes the OTHER Z buffer, and every other buffer is rewritten only after barriers
    }
    flush_gy();
    if (act) {
        for (int f = 0; f < 4; ++f) p.g_z0[((size_t)b * 4 + f) * n + i] = lam[f];
        if (p.g_shape) p.g_shape[(size_t)b * n + i] = gshp;
        if (p.g_pml) p.g_pml[(size_t)b * n + i] = p.pml_scale * gsig;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Pair variant of the register reverse kernel (opt-in: WAVES_LATENT_PAIR | WAVES_LATENT_ADJ_R1): k_latent_adjoint_r1 with
// two adjacent elements per thread as float2 and packed f32x2 arithmetic, as k_latent_integrate_r2 does for the forward
// pass.  n even, 4 <= n <= 1024.  NOT YET RUN ON A B200; equal to the generic reverse kernel (to float32 rounding) under the
// host emulation, race-free under ThreadSanitizer.
LAT_GLOBAL2 void k_latent_adjoint_r2(LatentP p) {
    LAT_SMEM
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x, e0 = 2 * tid, half = n >> 1;
    const int T = p.steps + 1, N = p.steps;
    float *ZA = (float *)lat_smem, *ZB = ZA + 4 * n;
    float *Y2 = ZB + 4 * n, *Y3 = Y2 + 4 * n, *Y4 = Y3 + 4 * n;
    float *PA = Y4 + 4 * n, *PB = PA + 4 * n;
    float *shs = PB + 4 * n, *Xs = shs + n, *fcs = Xs + LAT_FAST_NSEQ;
    const bool act = tid < half;
    const float *Yb = p.Y + (size_t)b * p.nseq * n;
    float *gY = p.g_Y ? p.g_Y + (size_t)b * p.nseq * n : nullptr;
    const lf2 zero2 = lf2_bc(0.0f);

    lf2 sg = zero2, sown = zero2;
    if (act) {
        if (p.shape) sown = lf2_mk(p.shape[(size_t)b * n + e0], p.shape[(size_t)b * n + e0 + 1]);
        *(lf2 *)&shs[e0] = sown;
        sg = lf2_mul(lf2_bc(p.pml_scale), lf2_mk(p.pml[(size_t)b * n + e0], p.pml[(size_t)b * n + e0 + 1]));
    }
    for (int k = tid; k < p.nseq; k += nt) Xs[k] = p.X[(size_t)b * p.nseq + k];
    LAT_SYNC();
    bool mono = true;
    for (int k = 0; k + 1 < p.nseq; ++k) mono = mono && (Xs[k] < Xs[k + 1]);
    const float rend = Xs[p.nseq - 1];

    const bool first = act && tid == 0, last = act && tid == half - 1;
    const int iL = e0 > 0 ? e0 - 1 : 0, iR = e0 + 2 < n ? e0 + 2 : n - 1;
    const lf2 sN = act ? lf2_mk(shs[iL], shs[iR]) : zero2;
    const int wlo = tid & ~31;
    const bool edge_warp = wlo == 0 || (wlo <= half - 1 && half - 1 < wlo + 32);
    const lf2 A0 = lf2_mk(first ? p.gf[0] : p.gc[0], last ? p.gl[0] : p.gc[0]);
    const lf2 A1 = lf2_mk(first ? p.gf[1] : p.gc[1], last ? p.gl[1] : p.gc[1]);
    const lf2 A2 = lf2_mk(first ? p.gf[2] : 0.0f, last ? p.gl[2] : 0.0f);
    const lf2 bcv = lf2_mk(first ? 0.0f : 1.0f, last ? 0.0f : 1.0f);
    const lf2 g0 = lf2_bc(p.gc[0]), g1 = lf2_bc(p.gc[1]), c02 = lf2_bc(p.c0);
    auto grad = [&](lf2 Nn, lf2 O) -> lf2 {  // (∇ * v) at the two elements, as in k_latent_integrate_r2
        if (!edge_warp) return lf2_add(lf2_mul(g0, lf2_mk(Nn.x, O.x)), lf2_mul(g1, lf2_mk(O.y, Nn.y)));
        const lf2 X0 = lf2_mk(first ? O.x : Nn.x, last ? Nn.x : O.x);
        const lf2 X1 = lf2_mk(O.y, last ? O.x : Nn.y);
        const lf2 X2 = lf2_mk(Nn.y, O.y);
        const lf2 two = lf2_add(lf2_mul(A0, X0), lf2_mul(A1, X1));
        const lf2 thr = lf2_add(two, lf2_mul(A2, X2));
        return lf2_mk(first ? thr.x : two.x, last ? thr.y : two.y);
    };
    auto rhs = [&](const float *S, lf2 a, float fs, const lf2 own[4], lf2 k[4]) {
        if (!act) return;
        const float *Ut = S, *Vt = S + n, *Ui = S + 2 * n, *Vi = S + 3 * n;
        const lf2 fs2 = lf2_bc(fs);
        const lf2 fN = lf2_mul(sN, fs2), fO = lf2_mul(sown, fs2);
        const lf2 gVt = grad(lf2_mk(Vt[iL], Vt[iR]), own[1]);
        const lf2 gUt = grad(lf2_add(lf2_mk(Ut[iL], Ut[iR]), fN), lf2_add(own[0], fO));
        const lf2 gVi = grad(lf2_mk(Vi[iL], Vi[iR]), own[3]);
        const lf2 gUi = grad(lf2_add(lf2_mk(Ui[iL], Ui[iR]), fN), lf2_add(own[2], fO));
        k[0] = lf2_mul(lf2_sub(lf2_mul(a, gVt), lf2_mul(sg, own[0])), bcv);
        k[1] = lf2_sub(lf2_mul(a, gUt), lf2_mul(sg, own[1]));
        k[2] = lf2_mul(lf2_sub(lf2_mul(c02, gVi), lf2_mul(sg, own[2])), bcv);
        k[3] = lf2_sub(lf2_mul(gUi, c02), lf2_mul(sg, own[3]));
    };
    // transposed rows: (∇ᵀ w)[j] = gc0 w[j+1] (row j+1 central) + gc1 w[j-1] (row j-1 central) + gf[j] w[0] (j <= 2)
    //                              + gl[j-(n-3)] w[n-1] (j >= n-3); per-thread coefficient pairs make the edge form branch-free
    auto tcoef = [&](int j, float &ta, float &tb, float &tf, float &tl) {
        ta = (j + 1 >= 1 && j + 1 <= n - 2) ? p.gc[0] : 0.0f;
        tb = (j - 1 >= 1 && j - 1 <= n - 2) ? p.gc[1] : 0.0f;
        tf = (j >= 0 && j <= 2) ? p.gf[j] : 0.0f;
        tl = (j >= n - 3 && j <= n - 1) ? p.gl[j - (n - 3)] : 0.0f;
    };
    float ta0, tb0, tf0, tl0, ta1, tb1, tf1, tl1;
    tcoef(act ? e0 : 3, ta0, tb0, tf0, tl0);
    tcoef(act ? e0 + 1 : 3, ta1, tb1, tf1, tl1);
    const lf2 TA = lf2_mk(ta0, ta1), TB = lf2_mk(tb0, tb1), TF = lf2_mk(tf0, tf1), TL = lf2_mk(tl0, tl1);
    const bool edge_t = wlo == 0 || wlo + 31 >= half - 2;  // warps holding elements 0,1,2 or n-3,n-2,n-1
    auto gradT = [&](const float *Q, lf2 O) -> lf2 {       // O = this thread's own pair of Q
        const lf2 Nn = lf2_mk(Q[iL], Q[iR]);
        if (!edge_t) return lf2_add(lf2_mul(g0, lf2_mk(O.y, Nn.y)), lf2_mul(g1, lf2_mk(Nn.x, O.x)));
        const lf2 inner2 = lf2_add(lf2_mul(TA, lf2_mk(O.y, Nn.y)), lf2_mul(TB, lf2_mk(Nn.x, O.x)));
        return lf2_add(inner2, lf2_add(lf2_mul(TF, lf2_bc(Q[0])), lf2_mul(TL, lf2_bc(Q[n - 1]))));
    };

    int kcur = 0, kreg = -1;
    lf2 yk = zero2, yk1 = zero2;
    float lc = 1.0f, rc = 0.0f;
    auto query = [&](float t, lf2 &c, int &kseg, float &w) {
        kseg = -1;
        w = 0.0f;
        c = zero2;
        if (!act) return;
        if (!mono) {
            c = lf2_mk(lat_speed_elem(p, Xs, Yb, t, e0), lat_speed_elem(p, Xs, Yb, t, e0 + 1));
            return;
        }
        if (!(lc <= t && t < rc)) {
            while (kcur + 2 < p.nseq && t >= Xs[kcur + 1]) ++kcur;
            while (kcur > 0 && t < Xs[kcur]) --kcur;
            const float l = Xs[kcur], r = Xs[kcur + 1];
            if (!lat_mask(l, r, rend, t)) {
                c = lf2_bc(p.c0 * (0.0f + (t - 0.0f) * 0.0f));
                return;
            }
            if (kreg != kcur) {
                yk = lf2_mk(Yb[(size_t)kcur * n + e0], Yb[(size_t)kcur * n + e0 + 1]);
                yk1 = lf2_mk(Yb[(size_t)(kcur + 1) * n + e0], Yb[(size_t)(kcur + 1) * n + e0 + 1]);
                lc = l;
                rc = r;
                kreg = kcur;
            }
        }
        const float dd = (rc - t) - (lc - t), x0 = 0.0f + lc;
        const lf2 dy = lf2_sub(yk1, yk);
        const lf2 dydx = lf2_add(zero2, lf2_mk(dy.x / dd, dy.y / dd));
        c = lf2_mul(c02, lf2_add(lf2_add(zero2, yk), lf2_mul(lf2_bc(t - x0), dydx)));
        kseg = kreg;
        w = (t - x0) / dd;
    };

    int kacc = -1;
    lf2 gYa = zero2, gYb = zero2;
    auto flush_gy = [&]() {
        if (gY && act && kacc >= 0) {
            gY[(size_t)kacc * n + e0] += gYa.x;
            gY[(size_t)kacc * n + e0 + 1] += gYa.y;
            gY[(size_t)(kacc + 1) * n + e0] += gYb.x;
            gY[(size_t)(kacc + 1) * n + e0 + 1] += gYb.y;
        }
        gYa = zero2;
        gYb = zero2;
    };
    auto add_gy = [&](lf2 cbar, float t, int kseg, float w) {
        if (!gY || !act) return;
        if (!mono) {
            float x0 = 0.0f;
            for (int k = 0; k + 1 < p.nseq; ++k)
                if (lat_mask(Xs[k], Xs[k + 1], rend, t)) x0 = x0 + Xs[k];
            for (int k = 0; k + 1 < p.nseq; ++k) {
                const float l = Xs[k], r = Xs[k + 1];
                if (lat_mask(l, r, rend, t)) {
                    const float wk = (t - x0) / ((r - t) - (l - t));
                    gY[(size_t)k * n + e0] += cbar.x * (1.0f - wk);
                    gY[(size_t)k * n + e0 + 1] += cbar.y * (1.0f - wk);
                    gY[(size_t)(k + 1) * n + e0] += cbar.x * wk;
                    gY[(size_t)(k + 1) * n + e0 + 1] += cbar.y * wk;
                }
            }
            return;
        }
        if (kseg < 0) return;
        if (kseg != kacc) {
            flush_gy();
            kacc = kseg;
        }
        gYa = lf2_add(gYa, lf2_mul(cbar, lf2_bc(1.0f - w)));
        gYb = lf2_add(gYb, lf2_mul(cbar, lf2_bc(w)));
    };

    lf2 lam[4], zr[4];
    for (int f = 0; f < 4; ++f) lam[f] = zr[f] = zero2;
    lf2 gshp = zero2, gsig = zero2;
    auto add_cotangent = [&](int idx, const lf2 zv[4]) {
        if (!act) return;
        if (p.w_energy) {
            const float w0 = p.w_energy[((size_t)b * 3 + 0) * T + idx], w1 = p.w_energy[((size_t)b * 3 + 1) * T + idx];
            const float w2 = p.w_energy[((size_t)b * 3 + 2) * T + idx];
            const lf2 d = lf2_sub(zv[0], zv[2]), k2 = lf2_bc(2.0f * p.dx);
            lam[0] = lf2_add(lam[0], lf2_mul(k2, lf2_add(lf2_mul(lf2_bc(w0), zv[0]), lf2_mul(lf2_bc(w2), d))));
            lam[2] = lf2_add(lam[2], lf2_mul(k2, lf2_sub(lf2_mul(lf2_bc(w1), zv[2]), lf2_mul(lf2_bc(w2), d))));
        }
        if (p.dL_dz)
            for (int f = 0; f < 4; ++f) {
                const float *src = p.dL_dz + (((size_t)idx * p.batch + b) * 4 + f) * n + e0;
                lam[f] = lf2_add(lam[f], lf2_mk(src[0], src[1]));
            }
    };
    auto load_state = [&](int idx) {
        if (!act) return;
        const float *zi = p.zt + ((size_t)idx * p.batch + b) * 4 * n + e0;
        for (int f = 0; f < 4; ++f) zr[f] = lf2_mk(zi[(size_t)f * n], zi[(size_t)f * n + 1]);
    };
    lf2 po[4];  // this thread's own pairs of the products last written by put_products
    for (int f = 0; f < 4; ++f) po[f] = zero2;
    auto put_products = [&](float *P, lf2 a, const lf2 kb[4]) {
        if (!act) return;
        po[0] = lf2_mul(lf2_mul(a, bcv), kb[0]);
        po[1] = lf2_mul(a, kb[1]);
        po[2] = lf2_mul(bcv, kb[2]);
        po[3] = kb[3];
        for (int f = 0; f < 4; ++f) *(lf2 *)&P[f * n + e0] = po[f];
    };
    auto vjp = [&](const float *S, const float *Pin, float fs, float t, int kseg, float w, const lf2 kb[4], lf2 yb[4]) {
        if (!act) return;
        const lf2 tq = gradT(Pin, po[0]);          // ∇ᵀ(a ⊙ bc ⊙ w_Utot)
        const lf2 tp = gradT(Pin + n, po[1]);      // ∇ᵀ(a ⊙ w_Vtot)
        const lf2 ts = gradT(Pin + 2 * n, po[2]);  // ∇ᵀ(bc ⊙ w_Uinc)
        const lf2 tr = gradT(Pin + 3 * n, po[3]);  // ∇ᵀ(w_Vinc)
        yb[0] = lf2_sub(tp, lf2_mul(lf2_mul(sg, bcv), kb[0]));
        yb[1] = lf2_sub(tq, lf2_mul(sg, kb[1]));
        yb[2] = lf2_sub(lf2_mul(c02, tr), lf2_mul(lf2_mul(sg, bcv), kb[2]));
        yb[3] = lf2_sub(lf2_mul(c02, ts), lf2_mul(sg, kb[3]));
        const lf2 Ut = *(const lf2 *)&S[e0], Vt = *(const lf2 *)&S[n + e0], Ui = *(const lf2 *)&S[2 * n + e0],
                  Vi = *(const lf2 *)&S[3 * n + e0];
        const lf2 fs2 = lf2_bc(fs);
        if (p.g_shape) gshp = lf2_add(gshp, lf2_mul(lf2_add(tp, lf2_mul(c02, tr)), fs2));
        if (p.g_pml) {
            const lf2 a0 = lf2_mul(lf2_mul(bcv, kb[0]), Ut), a1 = lf2_mul(kb[1], Vt);
            const lf2 a2 = lf2_mul(lf2_mul(bcv, kb[2]), Ui), a3 = lf2_mul(kb[3], Vi);
            gsig = lf2_sub(gsig, lf2_add(lf2_add(a0, a1), lf2_add(a2, a3)));
        }
        if (gY) {
            const lf2 fN = lf2_mul(sN, fs2), fO = lf2_mul(sown, fs2);
            const lf2 gVt = grad(lf2_mk(S[n + iL], S[n + iR]), Vt);
            const lf2 gUt = grad(lf2_add(lf2_mk(S[iL], S[iR]), fN), lf2_add(Ut, fO));
            const lf2 cbar = lf2_mul(c02, lf2_add(lf2_mul(lf2_mul(bcv, kb[0]), gVt), lf2_mul(kb[1], gUt)));
            add_gy(cbar, t, kseg, w);
        }
    };

    if (!p.compat) {
        load_state(N);
        add_cotangent(N, zr);
    }
    const float dt = p.dt, hdt = p.hdt, sixth = 1.0f / 6.0f;
    const lf2 hdt2 = lf2_bc(hdt), dt2 = lf2_bc(dt);
    const lf2 w4 = lf2_bc(sixth * dt), w23 = lf2_bc(2.0f * sixth * dt);
    const int istart = p.compat ? N : N - 1;
    int it = 0;
    for (int s = istart; s >= 0; --s, ++it) {
        float *Z = (it & 1) ? ZB : ZA;
        load_state(s);
        if (p.compat) add_cotangent(s, zr);
        if (s == istart || (s % LAT_CH) == LAT_CH - 1) {
            const int c0s = (s / LAT_CH) * LAT_CH;
            float *fw = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH;
            for (int q = tid; q < 3 * LAT_CH; q += nt) {
                const int st = c0s + q / 3, w = q - 3 * (q / 3);
                if (st <= s) {
                    const float ts = p.tspan[(size_t)b * T + st];
                    fw[q] = p.shape ? lat_sin_factor(w == 0 ? ts : (w == 1 ? ts + hdt : ts + dt), p.freq) : 0.0f;
                }
            }
        }
        const float t0 = p.tspan[(size_t)b * T + s], t1 = t0 + hdt, t2 = t0 + dt;
        lf2 cA, cB, cC;
        float wA, wB, wC;
        int kA, kB, kC;
        query(t2, cC, kC, wC);
        query(t1, cB, kB, wB);
        query(t0, cA, kA, wA);
        if (act)
            for (int f = 0; f < 4; ++f) *(lf2 *)&Z[f * n + e0] = zr[f];
        LAT_SYNC();
        const float *fc = fcs + ((s / LAT_CH) & 1) * 3 * LAT_CH + 3 * (s % LAT_CH);
        const float f0 = fc[0], f1 = fc[1], f2 = fc[2];

        lf2 k[4], ys[4], kb[4], yb[4], zsum[4];
        for (int f = 0; f < 4; ++f) k[f] = ys[f] = kb[f] = yb[f] = zsum[f] = zero2;
        rhs(Z, cA, f0, zr, k);
        for (int f = 0; f < 4; ++f) {
            ys[f] = lf2_add(zr[f], lf2_mul(hdt2, k[f]));
            if (act) *(lf2 *)&Y2[f * n + e0] = ys[f];
        }
        LAT_SYNC();
        rhs(Y2, cB, f1, ys, k);
        for (int f = 0; f < 4; ++f) {
            ys[f] = lf2_add(zr[f], lf2_mul(hdt2, k[f]));
            if (act) *(lf2 *)&Y3[f * n + e0] = ys[f];
        }
        LAT_SYNC();
        rhs(Y3, cB, f1, ys, k);
        for (int f = 0; f < 4; ++f) {
            ys[f] = lf2_add(zr[f], lf2_mul(dt2, k[f]));
            if (act) *(lf2 *)&Y4[f * n + e0] = ys[f];
            kb[f] = lf2_mul(w4, lam[f]);
        }
        put_products(PA, cC, kb);
        LAT_SYNC();
        vjp(Y4, PA, f2, t2, kC, wC, kb, yb);
        for (int f = 0; f < 4; ++f) {
            zsum[f] = yb[f];
            kb[f] = lf2_add(lf2_mul(w23, lam[f]), lf2_mul(dt2, yb[f]));
        }
        put_products(PB, cB, kb);
        LAT_SYNC();
        vjp(Y3, PB, f1, t1, kB, wB, kb, yb);
        for (int f = 0; f < 4; ++f) {
            zsum[f] = lf2_add(zsum[f], yb[f]);
            kb[f] = lf2_add(lf2_mul(w23, lam[f]), lf2_mul(hdt2, yb[f]));
        }
        put_products(PA, cB, kb);
        LAT_SYNC();
        vjp(Y2, PA, f1, t1, kB, wB, kb, yb);
        for (int f = 0; f < 4; ++f) {
            zsum[f] = lf2_add(zsum[f], yb[f]);
            kb[f] = lf2_add(lf2_mul(w4, lam[f]), lf2_mul(hdt2, yb[f]));
        }
        put_products(PB, cA, kb);
        LAT_SYNC();
        vjp(Z, PB, f0, t0, kA, wA, kb, yb);
        for (int f = 0; f < 4; ++f) lam[f] = lf2_add(lam[f], lf2_add(zsum[f], yb[f]));
        if (!p.compat) add_cotangent(s, zr);
    }
    flush_gy();
    if (act) {
        for (int f = 0; f < 4; ++f) {
            p.g_z0[((size_t)b * 4 + f) * n + e0] = lam[f].x;
            p.g_z0[((size_t)b * 4 + f) * n + e0 + 1] = lam[f].y;
        }
        if (p.g_shape) {
            p.g_shape[(size_t)b * n + e0] = gshp.x;
            p.g_shape[(size_t)b * n + e0 + 1] = gshp.y;
        }
        if (p.g_pml) {
            p.g_pml[(size_t)b * n + e0] = p.pml_scale * gsig.x;
            p.g_pml[(size_t)b * n + e0 + 1] = p.pml_scale * gsig.y;
        }
    }
}

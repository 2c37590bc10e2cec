/*
 * CPU oracle (C restatement) of the Waves.jl 2-D acoustic RK4 hot path.
 * TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Never part of the product.
 *
 * PARITY UNPINNED: the reference is Julia (absent from this image, no golden
 * vectors for the 2-D path).  This file is pinned to oracle/waves_oracle.py
 * bit-for-bit (tests/test_oracle.py) which restates the reference line by line.
 *
 * Same IEEE-float32 evaluation order as the reference's CPU path; build with
 * -ffp-contract=off so the compiler never fuses a*b+c (Julia does not).
 *
 * Layout: state (12, ny, nx), x fastest == the reference's column-major (nx, ny, 12).
 * Field order tot{U,Vx,Vy,Psix,Psiy,Omega}, inc{...}   (src/dynamics.jl:152-157,185-187)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int nx, ny;
    const float *first, *central, *last; /* gradient rows, src/operators.jl:10-22 */
    const float *sigma;                  /* 1-D pml profile, src/pml.jl:21-29 */
    float c0;
} wo_dyn;

/* ∇*u along x at (j,i): SparseMatrixCSC*dense accumulation order (src/operators.jl:45) */
static inline float d_x(const wo_dyn *d, const float *p, int j, int i) {
    const float *r = p + (size_t)j * d->nx;
    int n = d->nx;
    if (i == 0) return ((d->first[0] * r[0]) + (d->first[1] * r[1])) + (d->first[2] * r[2]);
    if (i == n - 1) return ((d->last[0] * r[n - 3]) + (d->last[1] * r[n - 2])) + (d->last[2] * r[n - 1]);
    return (d->central[0] * r[i - 1]) + (d->central[1] * r[i + 1]);
}

/* (∇*u')' along y (src/operators.jl:46) */
static inline float d_y(const wo_dyn *d, const float *p, int j, int i) {
    size_t nx = (size_t)d->nx;
    int n = d->ny;
    const float *c = p + i;
    if (j == 0) return ((d->first[0] * c[0]) + (d->first[1] * c[nx])) + (d->first[2] * c[2 * nx]);
    if (j == n - 1)
        return ((d->last[0] * c[(size_t)(n - 3) * nx]) + (d->last[1] * c[(size_t)(n - 2) * nx])) +
               (d->last[2] * c[(size_t)(n - 1) * nx]);
    return (d->central[0] * c[(size_t)(j - 1) * nx]) + (d->central[1] * c[(size_t)(j + 1) * nx]);
}

/* acoustic_dynamics, src/dynamics.jl:151-177.  x: 6 planes in, k: 6 planes out.
 * b2: per-cell c^2 plane or NULL (then scalar bs); uf: scratch plane for U .+ f. */
static void rhs6(const wo_dyn *d, const float *x, const float *b2, float bs, const float *f, float *uf, float *k) {
    const int nx = d->nx, ny = d->ny;
    const size_t P = (size_t)nx * ny;
    const float *U = x, *Vx = x + P, *Vy = x + 2 * P, *Px = x + 3 * P, *Py = x + 4 * P, *Om = x + 5 * P;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            size_t q = (size_t)j * nx + i;
            uf[q] = f ? (U[q] + f[q]) : (U[q] + 0.0f);
        }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j) {
        const float sy = d->sigma[j];
        const int jb = (j == 0 || j == ny - 1);
        for (int i = 0; i < nx; ++i) {
            size_t q = (size_t)j * nx + i;
            const float sx = d->sigma[i];
            const float b = b2 ? b2[q] : bs;
            const float bc = (jb || i == 0 || i == nx - 1) ? 0.0f : 1.0f;
            float Vxx = d_x(d, Vx, j, i);
            float Vyy = d_y(d, Vy, j, i);
            float Ux = d_x(d, uf, j, i);
            float Uy = d_y(d, uf, j, i);
            float dU = ((((b * (Vxx + Vyy)) + Px[q]) + Py[q]) - ((sx + sy) * U[q])) - Om[q];
            k[q] = bc * dU;
            k[P + q] = Ux - (sx * Vx[q]);
            k[2 * P + q] = Uy - (sy * Vy[q]);
            k[3 * P + q] = (b * sx) * Vyy;
            k[4 * P + q] = (b * sy) * Vxx;
            k[5 * P + q] = (sx * sy) * U[q];
        }
    }
}

typedef struct {
    int ncyl;
    const float *c0v, *c1v; /* (ncyl,4): x, y, r, c at ti and tf */
    float ti, tf;
} wo_design;

/* speed(interp(t), grid, c0)^2: src/designs.jl:99-116, 274-292 */
static void speed2(const wo_dyn *d, const wo_design *g, const float *xs, const float *ys, float t, float *b2) {
    const int nx = d->nx, ny = d->ny, n = g->ncyl;
    float *cur = (float *)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
    float dt = g->tf - g->ti;
    dt = dt > 0.0f ? dt : 1.0f;
    float inv = 1.0f / dt;
    float tc = t < g->ti ? g->ti : (t > g->tf ? g->tf : t);
    float s = tc - g->ti;
    for (int k = 0; k < 4 * n; ++k) {
        float dy = g->c1v[k] + (g->c0v[k] * -1.0f);
        float slope = dy * inv;
        cur[k] = g->c0v[k] + (slope * s);
    }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            int cnt = 0;
            float cd = 0.0f;
            for (int k = 0; k < n; ++k) {
                float ddx = xs[i] - cur[4 * k], ddy = ys[j] - cur[4 * k + 1];
                float r2 = cur[4 * k + 2] * cur[4 * k + 2];
                float d2 = (ddx * ddx) + (ddy * ddy);
                int m = d2 < r2;
                cnt += m;
                cd = cd + (m ? 1.0f : 0.0f) * cur[4 * k + 3];
            }
            float c = ((cnt == 0 ? 1.0f : 0.0f) * d->c0) + cd;
            b2[(size_t)j * nx + i] = c * c;
        }
    free(cur);
}

static float src_factor(float t, float freq) { /* src/sources.jl:68 */
    float two_pi = 2.0f * (float)M_PI;
    float arg = (two_pi * t) * freq;
    return (float)sin((double)arg);
}

/* Energies of one frame, src/env.jl:104-114; f64 accumulation, f32 result * dOmega */
static void frame_energy(const float *u, int nx, int ny, float dOmega, float *e3) {
    const size_t P = (size_t)nx * ny;
    const float *ut = u, *ui = u + 6 * P;
    double st = 0, si = 0, ss = 0;
#pragma omp parallel for reduction(+ : st, si, ss) schedule(static)
    for (long q = 0; q < (long)P; ++q) {
        float a = ut[q], b = ui[q], c = a - b;
        st += (double)a * a;
        si += (double)b * b;
        ss += (double)c * c;
    }
    e3[0] = (float)st * dOmega;
    e3[1] = (float)si * dOmega;
    e3[2] = (float)ss * dOmega;
}

/*
 * dyn(x,t,θ) for the whole 12-field state (src/dynamics.jl:179-188).
 * shape NULL -> NoSource (f = 0.0f0); ncyl < 0 -> NoDesign (scalar c0);
 * cplane != NULL overrides the design with a fixed c plane.
 */
int wo_rhs(int nx, int ny, const float *grad8, const float *sigma, float c0, const float *xs, const float *ys,
           int ncyl, const float *cyl0, const float *cyl1, float ti, float tf, const float *cplane,
           const float *shape, float freq, float t, const float *state, float *k) {
    const size_t P = (size_t)nx * ny;
    wo_dyn d = {nx, ny, grad8, grad8 + 3, grad8 + 5, sigma, c0};
    wo_design g = {ncyl, cyl0, cyl1, ti, tf};
    float *b2 = NULL, *f = NULL, *uf = (float *)malloc(sizeof(float) * P);
    if (cplane) {
        b2 = (float *)malloc(sizeof(float) * P);
        for (size_t q = 0; q < P; ++q) b2[q] = cplane[q] * cplane[q];
    } else if (ncyl >= 0) {
        b2 = (float *)malloc(sizeof(float) * P);
        speed2(&d, &g, xs, ys, t, b2);
    }
    if (shape) {
        f = (float *)malloc(sizeof(float) * P);
        float s = src_factor(t, freq);
        for (size_t q = 0; q < P; ++q) f[q] = shape[q] * s;
    }
    rhs6(&d, state, b2, c0 * c0, f, uf, k);
    rhs6(&d, state + 6 * P, NULL, c0 * c0, f, uf, k + 6 * P);
    free(uf);
    free(b2);
    free(f);
    return 0;
}

/*
 * Integrator (src/dynamics.jl:37-53) with runge_kutta (:9-16) for `steps` steps
 * starting at tspan[0]; state is updated in place.  energy (steps+1,3) nullable;
 * frames (nsave,12,ny,nx) receives the states after save_steps[] (0 = initial).
 */
int wo_integrate(int nx, int ny, const float *grad8, const float *sigma, float c0, float dt, const float *xs,
                 const float *ys, int ncyl, const float *cyl0, const float *cyl1, float ti, float tf,
                 const float *cplane, const float *shape, float freq, const float *tspan, int steps, float *state,
                 float dOmega, float *energy, const int *save_steps, int nsave, float *frames) {
    const size_t P = (size_t)nx * ny, N = 12 * P;
    wo_dyn d = {nx, ny, grad8, grad8 + 3, grad8 + 5, sigma, c0};
    wo_design g = {ncyl, cyl0, cyl1, ti, tf};
    float *k = (float *)malloc(sizeof(float) * N), *y = (float *)malloc(sizeof(float) * N);
    float *acc = (float *)malloc(sizeof(float) * N), *uf = (float *)malloc(sizeof(float) * P);
    float *b2 = (cplane || ncyl >= 0) ? (float *)malloc(sizeof(float) * P) : NULL;
    float *f = shape ? (float *)malloc(sizeof(float) * P) : NULL;
    if (cplane)
        for (size_t q = 0; q < P; ++q) b2[q] = cplane[q] * cplane[q];
    const float hdt = 0.5f * dt, sixth = 1.0f / 6.0f, bs = c0 * c0;
    int isave = 0;
    if (energy) frame_energy(state, nx, ny, dOmega, energy);
    while (isave < nsave && save_steps[isave] == 0) memcpy(frames + (size_t)(isave++) * N, state, sizeof(float) * N);

    for (int n = 0; n < steps; ++n) {
        const float t = tspan[n];
        const float ts[4] = {t, t + hdt, t + hdt, t + dt};
        const float as[3] = {hdt, hdt, dt};
        for (int s = 0; s < 4; ++s) {
            const float *in = s == 0 ? state : y;
            if (!cplane && ncyl >= 0) speed2(&d, &g, xs, ys, ts[s], b2);
            if (shape) {
                float sf = src_factor(ts[s], freq);
#pragma omp parallel for schedule(static)
                for (long q = 0; q < (long)P; ++q) f[q] = shape[q] * sf;
            }
            rhs6(&d, in, b2, bs, f, uf, k);
            rhs6(&d, in + 6 * P, NULL, bs, f, uf, k + 6 * P);
            /* runge_kutta: k1 .+ 2*k2 .+ 2*k3 .+ k4 summed left to right; stage state u .+ (a*k) */
#pragma omp parallel for schedule(static)
            for (long q = 0; q < (long)N; ++q) {
                float kk = k[q];
                if (s == 0) acc[q] = kk;
                else if (s == 3) acc[q] = acc[q] + kk;
                else acc[q] = acc[q] + 2.0f * kk;
                if (s < 3) y[q] = state[q] + as[s] * kk;
            }
        }
#pragma omp parallel for schedule(static)
        for (long q = 0; q < (long)N; ++q) {
            float du = sixth * acc[q];
            state[q] = state[q] + du * dt;
        }
        if (energy) frame_energy(state, nx, ny, dOmega, energy + 3 * (size_t)(n + 1));
        while (isave < nsave && save_steps[isave] == n + 1) memcpy(frames + (size_t)(isave++) * N, state, sizeof(float) * N);
    }
    free(k);
    free(y);
    free(acc);
    free(uf);
    free(b2);
    free(f);
    return 0;
}

int wo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py --impl reference: use every host core even when the launcher exported OMP_NUM_THREADS=1 (torchrun does) */
void wo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

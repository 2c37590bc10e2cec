// Fused reverse RK4 step for sm_100a: ONE launch set pulls a cotangent of the whole batch back through one runge_kutta call
// (src/dynamics.jl:9-16 under adjoint_sensitivity, src/dynamics.jl:97-118), including the energy-metric cotangent
// (src/env.jl:104-111) of the frame it passes.
//
// The right-hand side of the 2-D dynamics (src/dynamics.jl:151-188) is affine in the state, so the Jacobian of a step does not
// depend on the state: without a parameter gradient the reverse of one step is itself an RK4 step of the TRANSPOSED operator,
//     l4 = J(t+dt)^T (dt/6 w)            l3 = J(t+dt/2)^T (dt/3 w + dt l4)
//     l2 = J(t+dt/2)^T (dt/3 w + dt/2 l3) l1 = J(t)^T (dt/6 w + dt/2 l2)          w' = w + l4 + l3 + l2 + l1
// with (kernels_adjoint.cu; checked against float64 autograd in tests/test_gpu_adjoint_fused.py)
//     (J^T l)_U  = -(sx+sy) m lU + Dx^T lVx + Dy^T lVy + sx sy lOm        m = Dirichlet mask (src/dims.jl:117-124)
//     (J^T l)_Vx = Dx^T[b (m lU + sy lPsiy)] - sx lVx                      b = c^2 at the stage time
//     (J^T l)_Vy = Dy^T[b (m lU + sx lPsix)] - sy lVy
//     (J^T l)_Psix = (J^T l)_Psiy = m lU ;  (J^T l)_Om = -m lU
// and D^T the transpose of the 3-band matrix of src/operators.jl:10-22 (its one-sided first / last rows couple the three
// outermost cells to the edge cell).
//
// Temporal blocking in shared memory: a CTA owns a TX x TY tile of ONE wavefield of one environment, loads it with a 4-cell
// halo (one cell per RK stage), runs the four transposed stages on the tile in shared memory (the stage cotangent's
// differentiated quantities are exchanged through four shared planes, everything else lives in registers of the thread that
// owns the cell) and writes the tile once: 12 planes read + 12 written per reverse step instead of 30 per STAGE for the
// per-stage kernels of kernels_adjoint.cu.  The halo is recomputed redundantly; halo cells go stale one ring per stage and
// never feed an owned cell.  Two variants: INTERIOR tiles (sigma == 0 in the whole tile region, away from the domain border:
// mask 1, central stencils only, the three auxiliary cotangents never feed back and are only accumulated) and general tiles.
#include <algorithm>
#include <vector>

#include "waves_internal.h"

namespace {

constexpr int ATX = 64;        // tile width (owned columns)
constexpr int ASWP = ATX + 12;  // shared row pitch; a tile region is at most ATX + 10 columns wide
#ifndef WV_ADJ_NT_INT
#define WV_ADJ_NT_INT 1024
#endif
#ifndef WV_ADJ_NT_GEN
#define WV_ADJ_NT_GEN 512
#endif
#ifndef WV_ADJ_MINB_GEN
#define WV_ADJ_MINB_GEN 2
#endif
constexpr int ATY_INT = 28, ANT_INT = WV_ADJ_NT_INT;  // interior tiles: owned rows, threads (1024: 3 cells per thread, <= 64 registers)
constexpr int ATY_GEN = 12, ANT_GEN = WV_ADJ_NT_GEN;  // general tiles
constexpr int AGUARD = 2 * ASWP + 4;        // floats in front of / behind the planes: stencil reads of edge cells stay in bounds

struct AdjFArgs {
    GridP gp;
    const float *w_in;   // [n_env][12][plane] cotangent before the step
    float *w_out;        // same layout, after it (must not alias w_in: neighbouring tiles read the halo)
    const float *b2[3];  // c^2 planes of the total field at t, t + dt/2, t + dt: [n_env][plane]; nullptr: ambient c0^2
    // energy cotangent a = dL/dz of sum_k w3[k] E_k(z) (src/env.jl:104-111), z given by its two U planes [n_env][2][plane]:
    const float *pre_u;  // added to w_in BEFORE the step (the reference loop as written), nullptr: none
    const float *post_u; // added to w_out AFTER the step (exact discrete adjoint), nullptr: none
    float pre_w[3], post_w[3];
    float two_dO;        // 2 * dOmega
    int tiles_x;         // tiles per row of this launch
    int x_org, y_org;    // first owned column / row of tile (0, 0)
    int y_end;           // owned rows stop here (exclusive)
    int skip_x0, skip_x1, skip_y0, skip_y1;  // general launch: tiles inside this rectangle belong to the interior launch
};

extern __shared__ __align__(16) float adj_sm[];

__device__ __forceinline__ float energy_cot(const float *__restrict__ up, long long plane, int wf, const float w3[3], float two_dO) {
    const float ut = up[0], ui = up[plane], d = ut - ui;
    return wf == 0 ? two_dO * (w3[0] * ut + w3[2] * d) : two_dO * (w3[1] * ui - w3[2] * d);
}

// (D^T v)[i] along one axis for a cell at coordinate i of n, v read from shared memory at stride `st` around offset o
__device__ __forceinline__ float dT_gen(const float *__restrict__ v, int o, int st, int i, int n, const GridP &gp) {
    float acc = 0.0f;
    if (i + 1 <= n - 2) acc += gp.g_central[0] * v[o + st];   // interior row r = i+1 holds g_central[0] on column r-1
    if (i - 1 >= 1) acc += gp.g_central[1] * v[o - st];       // interior row r = i-1 holds g_central[1] on column r+1
    if (i <= 2) acc += (i == 0 ? gp.g_first[0] : (i == 1 ? gp.g_first[1] : gp.g_first[2])) * v[o - i * st];
    if (i >= n - 3) acc += (i == n - 3 ? gp.g_last[0] : (i == n - 2 ? gp.g_last[1] : gp.g_last[2])) * v[o + (n - 1 - i) * st];
    return acc;
}

template <bool INTERIOR, int TY, int NT>
__global__ void __launch_bounds__(NT, INTERIOR ? 1 : WV_ADJ_MINB_GEN) k_adjoint_step(const __grid_constant__ AdjFArgs A) {
    constexpr int SH = TY + 10, PL = SH * ASWP;            // rows / floats of one shared plane
    constexpr int NW = INTERIOR ? 3 : 6;                   // planes of w kept in shared memory
    constexpr int NL = INTERIOR ? 3 : 4;                   // exchanged planes of the stage cotangent
    constexpr int CPT = ((ATX + 10) * SH + NT - 1) / NT;   // cells per thread
    const GridP &gp = A.gp;
    const int nx = gp.nx, ny = gp.ny_global, nxp = gp.nxp;
    const long long P = gp.plane;
    const int tx = blockIdx.x % A.tiles_x, ty = blockIdx.x / A.tiles_x;
    const int e = blockIdx.y >> 1, wf = blockIdx.y & 1;
    const int x0 = A.x_org + tx * ATX, y0 = A.y_org + ty * TY;
    if (x0 >= nx || y0 >= A.y_end) return;
    if (!INTERIOR && x0 >= A.skip_x0 && x0 + ATX <= A.skip_x1 && y0 >= A.skip_y0 && y0 + TY <= A.skip_y1) return;
    const int x1 = min(x0 + ATX, nx), y1 = min(y0 + TY, A.y_end);
    // tile region [L, R) x [T0, B): 4 halo cells per side, clipped to the domain; a region that would end 1 or 2 cells short of
    // the domain edge is extended to it (the transposed one-sided rows couple the edge cell to its two inner neighbours)
    const int L = max(x0 - 4, 0), T0 = max(y0 - 4, 0);
    int R = min(x1 + 4, nx), B = min(y1 + 4, ny);
    if (nx - R <= 2) R = nx;
    if (ny - B <= 2) B = ny;
    const int Wr = R - L, ncell = Wr * (B - T0);

    float *Wt = adj_sm + AGUARD;        // NW planes: w of this tile (U, Vx, Vy[, Psix, Psiy, Omega])
    float *Lx = Wt + NW * PL;           // NL planes: INTERIOR {b lU, lVx, lVy}; general {qx, qy, lVx, lVy}
    const float *win = A.w_in + ((long long)e * 12 + wf * 6) * P;
    float *wout = A.w_out + ((long long)e * 12 + wf * 6) * P;

    // ---- ownership: cell c = tid + k NT of the region, row-major ----
    int off[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = threadIdx.x + k * NT;
        if (c < ncell) {
            const int lr = c / Wr, lc = c - lr * Wr;
            off[k] = lr * ASWP + lc;
        } else
            off[k] = -1;
    }
    const int gq0 = T0 * nxp + L;
    auto gq_of = [&](int o) { const int lr = o / ASWP; return gq0 + lr * nxp + (o - lr * ASWP); };  // global offset in a plane
    // ---- load w (+ the cotangent injected before the step) ----
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        if (off[k] < 0) continue;
        const int q = gq_of(off[k]);
#pragma unroll
        for (int f = 0; f < NW; ++f) Wt[f * PL + off[k]] = win[f * P + q];
        if (A.pre_u) Wt[off[k]] += energy_cot(A.pre_u + (long long)e * 2 * P + q, P, wf, A.pre_w, A.two_dO);
    }
    float oU[CPT], oVx[CPT], oVy[CPT], og[CPT], sU[CPT], sVx[CPT], sVy[CPT], sg[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) oU[k] = oVx[k] = oVy[k] = og[k] = sU[k] = sVx[k] = sVy[k] = sg[k] = 0.0f;
    __syncthreads();

    const float dt = gp.dt;
    const float cm = gp.g_central[0], cp = gp.g_central[1];
#pragma unroll
    for (int s = 0; s < 4; ++s) {   // stages 4, 3, 2, 1 of the forward step
        const float a = (s == 0 || s == 3) ? dt * (1.0f / 6.0f) : dt * (1.0f / 3.0f);
        const float be = s == 0 ? 0.0f : (s == 1 ? dt : 0.5f * dt);
        const int tau = s == 0 ? 2 : (s == 3 ? 0 : 1);
        const float *b2p = (wf == 0 && A.b2[tau]) ? A.b2[tau] + (long long)e * P : nullptr;
        // phase A: the differentiated quantities of the stage cotangent l = a w + be l_prev
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            if (off[k] < 0) continue;
            const int o = off[k];
            const float b = b2p ? __ldg(b2p + gq_of(o)) : gp.b0;
            const float lU = a * Wt[o] + be * oU[k];
            if (INTERIOR) {
                Lx[o] = b * lU;
                Lx[PL + o] = a * Wt[PL + o] + be * oVx[k];
                Lx[2 * PL + o] = a * Wt[2 * PL + o] + be * oVy[k];
            } else {
                const int lr = o / ASWP, lc = o - lr * ASWP, x = L + lc, y = T0 + lr;
                const float sx = __ldg(gp.sigma + x), sy = __ldg(gp.sigma + y);
                const float m = (x == 0 || x == nx - 1 || y == 0 || y == ny - 1) ? 0.0f : 1.0f;
                const float lPx = a * Wt[3 * PL + o] + be * og[k], lPy = a * Wt[4 * PL + o] + be * og[k];
                const float mU = m * lU;
                Lx[o] = b * (mU + sy * lPy);
                Lx[PL + o] = b * (mU + sx * lPx);
                Lx[2 * PL + o] = a * Wt[PL + o] + be * oVx[k];
                Lx[3 * PL + o] = a * Wt[2 * PL + o] + be * oVy[k];
            }
        }
        __syncthreads();
        // phase B: l_new = J^T l
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            if (off[k] < 0) continue;
            const int o = off[k];
            const float lU = a * Wt[o] + be * oU[k];
            if (INTERIOR) {
                const float *q = Lx, *lvx = Lx + PL, *lvy = Lx + 2 * PL;
                oU[k] = (cm * lvx[o + 1] + cp * lvx[o - 1]) + (cm * lvy[o + ASWP] + cp * lvy[o - ASWP]);
                oVx[k] = cm * q[o + 1] + cp * q[o - 1];
                oVy[k] = cm * q[o + ASWP] + cp * q[o - ASWP];
                sg[k] += lU;
            } else {
                const int lr = o / ASWP, lc = o - lr * ASWP, x = L + lc, y = T0 + lr;
                const float sx = __ldg(gp.sigma + x), sy = __ldg(gp.sigma + y);
                const float m = (x == 0 || x == nx - 1 || y == 0 || y == ny - 1) ? 0.0f : 1.0f;
                const float lOm = a * Wt[5 * PL + o] - be * og[k];
                const float mU = m * lU;
                const float *qx = Lx, *qy = Lx + PL, *lvx = Lx + 2 * PL, *lvy = Lx + 3 * PL;
                oU[k] = -(sx + sy) * mU + dT_gen(lvx, o, 1, x, nx, gp) + dT_gen(lvy, o, ASWP, y, ny, gp) + (sx * sy) * lOm;
                oVx[k] = dT_gen(qx, o, 1, x, nx, gp) - sx * lvx[o];
                oVy[k] = dT_gen(qy, o, ASWP, y, ny, gp) - sy * lvy[o];
                og[k] = mU;
                sg[k] += mU;
            }
            sU[k] += oU[k];
            sVx[k] += oVx[k];
            sVy[k] += oVy[k];
        }
        __syncthreads();
    }

    // ---- w' = w + sum of the four stage results (+ the cotangent injected after the step), owned cells only ----
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        if (off[k] < 0) continue;
        const int o = off[k], lr = o / ASWP, lc = o - lr * ASWP, x = L + lc, y = T0 + lr;
        if (x < x0 || x >= x1 || y < y0 || y >= y1) continue;
        const int q = gq0 + lr * nxp + lc;
        float vU = Wt[o] + sU[k];
        if (A.post_u) vU += energy_cot(A.post_u + (long long)e * 2 * P + q, P, wf, A.post_w, A.two_dO);
        wout[q] = vU;
        wout[P + q] = Wt[PL + o] + sVx[k];
        wout[2 * P + q] = Wt[2 * PL + o] + sVy[k];
        if (INTERIOR) {
            wout[3 * P + q] = win[3 * P + q] + sg[k];
            wout[4 * P + q] = win[4 * P + q] + sg[k];
            wout[5 * P + q] = win[5 * P + q] - sg[k];
        } else {
            wout[3 * P + q] = Wt[3 * PL + o] + sg[k];
            wout[4 * P + q] = Wt[4 * PL + o] + sg[k];
            wout[5 * P + q] = Wt[5 * PL + o] - sg[k];
        }
    }
}

template <bool INTERIOR, int TY>
constexpr size_t adj_smem() {
    return sizeof(float) * ((size_t)(INTERIOR ? 6 : 10) * (TY + 10) * ASWP + 2 * AGUARD);
}

// gather the two U planes of every environment of a state: [n_env][12][plane] -> [n_env][2][plane]
__global__ void k_gather_u(GridP gp, const float *__restrict__ u, float *__restrict__ out) {
    const long long n = gp.plane, q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y >> 1, f = blockIdx.y & 1;
    if (q < n) out[((long long)e * 2 + f) * n + q] = u[((long long)e * 12 + f * 6) * n + q];
}

}  // namespace

// One fused reverse step for every environment: w_out = [pre-injection] (I + J_step^T) w_in [+ post-injection].
// b2[tau]: c^2 planes of the total field at the three stage times (nullptr: no design anywhere).
int launch_adjoint_step_fused(waves_handle *h, const float *w_in, float *w_out, const float *const b2[3], const float *pre_u,
                              const float *pre_w3, const float *post_u, const float *post_w3) {
    const GridP &gp = h->gp;
    if (gp.ny_own != gp.ny_global) return waves_set_error("fused reverse step: not available on slab handles");
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t ce = cudaFuncSetAttribute(k_adjoint_step<true, ATY_INT, ANT_INT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)adj_smem<true, ATY_INT>());
        if (ce == cudaSuccess)
            ce = cudaFuncSetAttribute(k_adjoint_step<false, ATY_GEN, ANT_GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)adj_smem<false, ATY_GEN>());
        if (ce != cudaSuccess) return waves_set_error("fused reverse step: cudaFuncSetAttribute failed (was the library built for sm_100a?)");
        attr_done = true;
    }
    // zero-sigma zone [z0, z1) of the profile, cached per handle
    if (h->adj_z1 == 0 && h->adj_z0 == 0) {
        std::vector<float> sig(gp.nx);
        cudaMemcpy(sig.data(), h->d_sigma, sizeof(float) * gp.nx, cudaMemcpyDeviceToHost);
        int z0 = 0, z1 = gp.nx;
        while (z0 < gp.nx && sig[z0] != 0.0f) ++z0;
        while (z1 > z0 && sig[z1 - 1] != 0.0f) --z1;
        bool clean = true;
        for (int i = z0; i < z1; ++i) clean = clean && sig[i] == 0.0f;
        if (!clean) z0 = z1 = -1;
        h->adj_z0 = z0;
        h->adj_z1 = z1 == 0 ? -1 : z1;
    }
    const int z0 = std::max(h->adj_z0, 3), z1 = std::min(h->adj_z1, gp.nx - 3);
    // interior rectangle: multiples of the tile sizes of BOTH launches, every tile region (4-cell halo) inside the zero zone
    const int lcm_y = ATY_GEN;  // interior rows start / stop on general-tile boundaries
    int xi0 = ((z0 + 4 + ATX - 1) / ATX) * ATX, xi1 = ((z1 - 4) / ATX) * ATX;
    int yi0 = ((z0 + 4 + lcm_y - 1) / lcm_y) * lcm_y, yi1 = ((z1 - 4) / lcm_y) * lcm_y;
    const bool has_int = h->adj_z0 >= 0 && xi1 > xi0 && yi1 > yi0;

    AdjFArgs A;
    A.gp = gp;
    A.w_in = w_in;
    A.w_out = w_out;
    for (int t = 0; t < 3; ++t) A.b2[t] = b2 ? b2[t] : nullptr;
    A.pre_u = pre_u;
    A.post_u = post_u;
    for (int k = 0; k < 3; ++k) {
        A.pre_w[k] = pre_w3 ? pre_w3[k] : 0.0f;
        A.post_w[k] = post_w3 ? post_w3[k] : 0.0f;
    }
    A.two_dO = 2.0f * h->d_omega;
    if (has_int) {
        A.tiles_x = (xi1 - xi0) / ATX;
        A.x_org = xi0;
        A.y_org = yi0;
        A.y_end = yi1;
        A.skip_x0 = A.skip_x1 = A.skip_y0 = A.skip_y1 = 0;
        const int tiles_y = (yi1 - yi0 + ATY_INT - 1) / ATY_INT;
        dim3 grd(A.tiles_x * tiles_y, gp.n_env * 2);
        k_adjoint_step<true, ATY_INT, ANT_INT><<<grd, ANT_INT, adj_smem<true, ATY_INT>(), h->stream>>>(A);
        h->launches++;
    }
    A.tiles_x = (gp.nx + ATX - 1) / ATX;
    A.x_org = 0;
    A.y_org = 0;
    A.y_end = gp.ny_global;
    A.skip_x0 = has_int ? xi0 : 0;
    A.skip_x1 = has_int ? xi1 : 0;
    A.skip_y0 = has_int ? yi0 : 0;
    A.skip_y1 = has_int ? yi1 : 0;
    {
        const int tiles_y = (gp.ny_global + ATY_GEN - 1) / ATY_GEN;
        dim3 grd(A.tiles_x * tiles_y, gp.n_env * 2);
        k_adjoint_step<false, ATY_GEN, ANT_GEN><<<grd, ANT_GEN, adj_smem<false, ATY_GEN>(), h->stream>>>(A);
        h->launches++;
    }
    return 0;
}

void launch_gather_u(waves_handle *h, const float *u, float *out) {
    dim3 grd((unsigned)((h->gp.plane + 255) / 256), h->gp.n_env * 2);
    k_gather_u<<<grd, 256, 0, h->stream>>>(h->gp, u, out);
    h->launches++;
}

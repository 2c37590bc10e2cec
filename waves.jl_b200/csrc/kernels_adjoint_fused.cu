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
// Since round 2 the INTERIOR of the step runs on the forward kernel's march (kernels_fused.cu: stage_T, k_fused_step<6>); the kernels
// of this file advance the frame around it (k_adjoint_step_gen4 on four bands, launch_adjoint_step_frame) or, with
// WAVES_ADJ_TILES or where the march does not apply, everything (launch_adjoint_step_fused).
//
// Temporal blocking in shared memory: a CTA owns a 64 x TY tile of ONE wavefield of one environment, loads it with a 4-cell
// halo (one cell per RK stage), runs the four transposed stages on the tile in shared memory and writes the tile once: 12
// planes read + 12 written per reverse step instead of 30 per STAGE for the per-stage kernels of kernels_adjoint.cu.  The halo
// is recomputed redundantly; halo cells go stale one ring per stage and never feed an owned cell.  A thread owns four
// consecutive cells of a row (float4 global and shared accesses, its own exchanged values in registers between the two phases
// of a stage).  Shared memory per CTA: 6 planes of w, 3 or 4 exchanged planes of the stage cotangent, 0 / 1 / 3 planes of c^2,
// each 4 * 512 floats, pitched to the region width so that consecutive threads touch consecutive banks.  Two variants:
// INTERIOR tiles (sigma == 0 in the whole tile region, away from the domain border: mask 1, central stencils only, the three
// auxiliary cotangents never feed back and are only accumulated) and general tiles.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "waves_internal.h"

namespace {

constexpr int ATX = 64;             // tile width (owned columns)
constexpr int AWMAX = ATX + 10;     // a tile region is at most ATX + 10 columns wide (4 halo columns per side, + 2 at a domain edge)
#ifndef WV_ADJ_TY_INT
#define WV_ADJ_TY_INT 17
#endif
#ifndef WV_ADJ_TY_GEN
#define WV_ADJ_TY_GEN 16
#endif
#ifndef WV_ADJ_TX_INT
#define WV_ADJ_TX_INT 64   // owned columns of an interior tile (64 or 32)
#endif
#ifndef WV_ADJ_NT_INT
#define WV_ADJ_NT_INT 512  // threads of an interior CTA (512: two CTAs per SM, 256: four)
#endif
#ifndef WV_ADJ_SHFL
#define WV_ADJ_SHFL 1      // x neighbours of a 4-cell group by warp shuffle (the scalar shared-memory reads are 4-way bank conflicts)
#endif
constexpr int ATY_INT = WV_ADJ_TY_INT, ATY_GEN = WV_ADJ_TY_GEN;  // owned rows of interior / general tiles
constexpr int ATX_INT = WV_ADJ_TX_INT, ANT_INT = WV_ADJ_NT_INT;
constexpr int ANT = 512;            // threads per CTA; two CTAs per SM (<= 64 registers, <= 113 KB of shared memory each)
constexpr int AGUARD = 2 * AWMAX + 4;  // floats in front of / behind the planes: stencil reads of edge cells stay in bounds

struct AdjFArgs {
    GridP gp;
    const float *w_in;   // [n_env][12][plane] cotangent before the step
    float *w_out;        // same layout, after it (must not alias w_in: neighbouring tiles read the halo)
    const float *b2[3];  // c^2 planes of the total field at t, t + dt/2, t + dt: [n_env][plane]; nullptr: ambient c0^2
    int nb;              // distinct planes among b2[0..2]: 0 (none), 1 (a design that does not move) or 3
    // energy cotangent a = dL/dz of sum_k w3[k] E_k(z) (src/env.jl:104-111), z given by its two U planes [n_env][2][plane]:
    const float *pre_u;  // added to w_in BEFORE the step (the reference loop as written), nullptr: none
    const float *post_u; // added to w_out AFTER the step (exact discrete adjoint), nullptr: none
    float pre_w[3], post_w[3];
    float two_dO;        // 2 * dOmega
    int tiles_x;         // tiles per row of this launch
    int x_org, y_org;    // first owned column / row of tile (0, 0)
    int y_end;           // owned rows stop here (exclusive)
    int skip_x0, skip_x1, skip_y0, skip_y1;  // general launch: tiles inside this rectangle belong to the interior launch
    // general launch over up to four rectangular bands (the frame around the rectangle the march kernel owns): band k holds the
    // blocks [bstart[k], bstart[k+1]) as tiles of ATX x TY from (bx0, by0), clipped to (bx1, by1); nband == 0: one band = the domain
    int nband;
    int bx0[4], bx1[4], by0[4], by1[4], btx[4], bty[4], bstart[5];   // bty: owned rows per tile of the band (as many as its width lets 512 threads hold)
};

extern __shared__ __align__(16) float adj_sm[];

__device__ __forceinline__ float energy_cot(const float *__restrict__ up, long long plane, int wf, const float w3[3], float two_dO) {
    const float ut = up[0], ui = up[plane], d = ut - ui;
    return wf == 0 ? two_dO * (w3[0] * ut + w3[2] * d) : two_dO * (w3[1] * ui - w3[2] * d);
}

// Interior tiles, vectorised along x: the region of an interior tile is always 72 columns wide (full halos) and starts on a
// multiple of 4 columns, so a thread owns FOUR consecutive cells of one row: global memory moves as float4, its own cells'
// exchanged values stay in registers between the two phases of a stage, the y neighbours arrive as two LDS.128 per plane and
// the x neighbours as two scalar loads -- a quarter of the shared-memory and address instructions of the scalar form.
// 512 threads x 4 cells cover up to 27 rows of 72 (TY <= 17).
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 operator*(float s, float4 v) { return make_float4(s * v.x, s * v.y, s * v.z, s * v.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 b) { return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w)); }

template <int TY, int TX, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) k_adjoint_step_int4(const __grid_constant__ AdjFArgs A) {
    constexpr int WR = TX + 8, GPR = WR / 4;      // region width, 4-cell groups per row
    constexpr int PLF = 4 * NT;                    // floats per plane
    static_assert((TY + 8) * GPR <= NT, "an interior region must fit one group per thread");
    const GridP &gp = A.gp;
    const int nxp = gp.nxp;
    const long long P = gp.plane;
    const int tx = blockIdx.x % A.tiles_x, ty = blockIdx.x / A.tiles_x;
    const int e = blockIdx.y >> 1, wf = blockIdx.y & 1;
    const int x0 = A.x_org + tx * TX, y0 = A.y_org + ty * TY;
    if (y0 >= A.y_end) return;
    const int y1 = min(y0 + TY, A.y_end);
    const int L = x0 - 4, T0 = y0 - 4, Hr = y1 + 4 - T0;   // interior tiles keep full halos inside the domain
    const int ngrp = Hr * GPR;
    // (the three auxiliary cotangents of an interior tile feed nothing -- sigma == 0 -- and only receive the sum of the stage
    // cotangents of their own cell: they are not staged in shared memory, the owned cells read and write them at the end)
    float *Wt = adj_sm + AGUARD, *Lx = Wt + 3 * PLF, *Bt = Lx + 3 * PLF;
    const float *win = A.w_in + ((long long)e * 12 + wf * 6) * P;
    float *wout = A.w_out + ((long long)e * 12 + wf * 6) * P;
    const int nb = wf == 0 ? A.nb : 0;

    const int t = threadIdx.x, c0 = 4 * t;            // first of this thread's four cells: float c0 of every plane
    const bool act = t < ngrp;
    const int lr = act ? t / GPR : 0, lc = act ? 4 * (t - lr * GPR) : 0;
    const int q = (T0 + lr) * nxp + L + lc;           // first cell in a global plane (16-byte aligned)
#pragma unroll
    for (int f = 0; f < 3; ++f) st4(Wt + f * PLF + c0, ld4(win + f * P + q));
    if (A.pre_u) {
        const float *up = A.pre_u + (long long)e * 2 * P + q;
        const float4 ut = ld4(up), ui = ld4(up + P), d = ut - ui;
        const float4 a4 = wf == 0 ? A.two_dO * (A.pre_w[0] * ut + A.pre_w[2] * d) : A.two_dO * (A.pre_w[1] * ui - A.pre_w[2] * d);
        st4(Wt + c0, ld4(Wt + c0) + a4);
    }
    for (int tb = 0; tb < nb; ++tb) st4(Bt + tb * PLF + c0, ld4(A.b2[tb] + (long long)e * P + q));
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 oU = z4, oVx = z4, oVy = z4, sU = z4, sVx = z4, sVy = z4, sg = z4;
    __syncthreads();

    const float dt = gp.dt, cp = gp.g_central[1];
#pragma unroll
    for (int s = 0; s < 4; ++s) {   // stages 4, 3, 2, 1 of the forward step
        const float a = (s == 0 || s == 3) ? dt * (1.0f / 6.0f) : dt * (1.0f / 3.0f);
        const float be = s == 0 ? 0.0f : (s == 1 ? dt : 0.5f * dt);
        const int tau = s == 0 ? 2 : (s == 3 ? 0 : 1);
        // phase A: l = a w + be l_prev; exchange b lU, lVx, lVy
        const float4 lU = fma4(be, oU, a * ld4(Wt + c0));
        const float4 b4 = nb == 0 ? make_float4(gp.b0, gp.b0, gp.b0, gp.b0) : ld4(Bt + (nb == 1 ? 0 : tau) * PLF + c0);
        const float4 qc = b4 * lU;
        const float4 vx = fma4(be, oVx, a * ld4(Wt + PLF + c0)), vy = fma4(be, oVy, a * ld4(Wt + 2 * PLF + c0));
        st4(Lx + c0, qc);
        st4(Lx + PLF + c0, vx);
        st4(Lx + 2 * PLF + c0, vy);
#if WV_ADJ_SHFL
        float ql = __shfl_up_sync(0xffffffffu, qc.w, 1), qr = __shfl_down_sync(0xffffffffu, qc.x, 1);
        float vl = __shfl_up_sync(0xffffffffu, vx.w, 1), vr = __shfl_down_sync(0xffffffffu, vx.x, 1);
#endif
        __syncthreads();
        // phase B: l_new = J^T l with central rows only (D^T = -D)
#if WV_ADJ_SHFL
        // the x neighbours sit in the neighbouring lanes' registers (a group that starts / ends a row gets a value of another row:
        // those are the outermost halo columns, stale by construction); lanes 0 and 31 read theirs from shared memory
        if ((t & 31) == 0) {
            ql = Lx[c0 - 1];
            vl = Lx[PLF + c0 - 1];
        }
        if ((t & 31) == 31) {
            qr = Lx[c0 + 4];
            vr = Lx[PLF + c0 + 4];
        }
#else
        const float ql = Lx[c0 - 1], qr = Lx[c0 + 4], vl = Lx[PLF + c0 - 1], vr = Lx[PLF + c0 + 4];
#endif
        const float4 qu = ld4(Lx + c0 - WR), qd = ld4(Lx + c0 + WR), yu = ld4(Lx + 2 * PLF + c0 - WR), yd = ld4(Lx + 2 * PLF + c0 + WR);
        const float4 dxv = make_float4(vl - vx.y, vx.x - vx.z, vx.y - vx.w, vx.z - vr);
        const float4 dxq = make_float4(ql - qc.y, qc.x - qc.z, qc.y - qc.w, qc.z - qr);
        oU = cp * (dxv + (yu - yd));
        oVx = cp * dxq;
        oVy = cp * (qu - qd);
        sg = sg + lU;
        sU = sU + oU;
        sVx = sVx + oVx;
        sVy = sVy + oVy;
        __syncthreads();
    }
    const int y = T0 + lr;
    if (!act || lc < 4 || lc >= WR - 4 || y < y0 || y >= y1) return;   // whole groups are owned or not (x0 = L + 4)
    float4 vU = ld4(Wt + c0) + sU;
    if (A.post_u) {
        const float *up = A.post_u + (long long)e * 2 * P + q;
        const float4 ut = ld4(up), ui = ld4(up + P), d = ut - ui;
        vU = vU + (wf == 0 ? A.two_dO * (A.post_w[0] * ut + A.post_w[2] * d) : A.two_dO * (A.post_w[1] * ui - A.post_w[2] * d));
    }
    st4(wout + q, vU);
    st4(wout + P + q, ld4(Wt + PLF + c0) + sVx);
    st4(wout + 2 * P + q, ld4(Wt + 2 * PLF + c0) + sVy);
    st4(wout + 3 * P + q, ld4(win + 3 * P + q) + sg);
    st4(wout + 4 * P + q, ld4(win + 4 * P + q) + sg);
    st4(wout + 5 * P + q, ld4(win + 5 * P + q) - sg);
}

// General tiles (PML strips, corners, domain border), vectorised along x like the interior form: four consecutive cells of a row
// per thread, planes pitched to a multiple of 4 columns.  The transposed derivative with its one-sided first / last rows
// (src/operators.jl:3-4) is the CENTRAL transposed stencil with zeros behind the domain plus a rank-one correction on the three
// outermost cells of each side,
//     (D^T v)[i] = cp (v[i-1] - v[i+1])  +  (gf0, gf1 - cp, gf2)[i] v[0]        for i = 0, 1, 2
//                                         +  (gl0, gl1 - cm, gl2)[i-(n-3)] v[n-1]  for i = n-3, n-2, n-1
// (needs n >= 8 so that the two sides do not meet; waves_adjoint takes the per-stage kernels on smaller grids).  v[0] of a row is
// the first component of the thread that holds columns 0..3, v[n-1] is read from the exchanged plane; along y the corrections
// are one extra row read with a per-thread scalar coefficient.  Every thread runs the same few instructions: the border costs
// two selects and short predicated blocks instead of a divergent out-of-line routine.  Columns behind the domain's last one
// (the pad of a pitched row) are kept at zero in the exchanged planes.
template <int TY>
__global__ void __launch_bounds__(ANT, 2) k_adjoint_step_gen4(const __grid_constant__ AdjFArgs A) {
    constexpr int PLF = 4 * ANT;
    static_assert((TY + 10) * ((AWMAX + 3) / 4) <= ANT, "a general region must fit one group per thread");
    const GridP &gp = A.gp;
    const int nx = gp.nx, ny = gp.ny_global, nxp = gp.nxp;
    const long long P = gp.plane;
    const int e = blockIdx.y >> 1, wf = blockIdx.y & 1;
    int x0, y0, xe, ye, tyr = TY;
    if (A.nband > 0) {
        int k = 0;
        while (k + 1 < A.nband && (int)blockIdx.x >= A.bstart[k + 1]) ++k;
        const int bi = (int)blockIdx.x - A.bstart[k];
        const int tx = bi % A.btx[k], ty = bi / A.btx[k];
        tyr = A.bty[k];
        x0 = A.bx0[k] + tx * ATX, y0 = A.by0[k] + ty * tyr;
        xe = A.bx1[k], ye = A.by1[k];
    } else {
        const int tx = blockIdx.x % A.tiles_x, ty = blockIdx.x / A.tiles_x;
        x0 = A.x_org + tx * ATX, y0 = A.y_org + ty * TY;
        xe = nx, ye = A.y_end;
        if (x0 >= A.skip_x0 && x0 + ATX <= A.skip_x1 && y0 >= A.skip_y0 && y0 + TY <= A.skip_y1) return;
    }
    if (x0 >= xe || y0 >= ye) return;
    const int x1 = min(x0 + ATX, xe), y1 = min(y0 + tyr, ye);
    const int L = max(x0 - 4, 0), T0 = max(y0 - 4, 0);
    int R = min(x1 + 4, nx), B = min(y1 + 4, ny);
    if (nx - R <= 2) R = nx;
    if (ny - B <= 2) B = ny;
    const int Wp = (R - L + 3) & ~3, gpr = Wp >> 2, ngrp = gpr * (B - T0);   // pitch (floats), groups per row, groups
    float *Wt = adj_sm + AGUARD, *Lx = Wt + 6 * PLF, *Bt = Lx + 4 * PLF;
    const float *win = A.w_in + ((long long)e * 12 + wf * 6) * P;
    float *wout = A.w_out + ((long long)e * 12 + wf * 6) * P;
    const int nb = wf == 0 ? A.nb : 0;

    const int t = threadIdx.x, c0 = 4 * t;
    const bool act = t < ngrp;
    const int lr = act ? t / gpr : 0, lc = act ? 4 * (t - lr * gpr) : 0;
    const int xg = L + lc, y = T0 + lr;               // first column of the group, its row
    const int q = y * nxp + xg;
#pragma unroll
    for (int f = 0; f < 6; ++f) st4(Wt + f * PLF + c0, ld4(win + f * P + q));
    if (A.pre_u) {
        const float *up = A.pre_u + (long long)e * 2 * P + q;
        const float4 ut = ld4(up), ui = ld4(up + P), d = ut - ui;
        const float4 a4 = wf == 0 ? A.two_dO * (A.pre_w[0] * ut + A.pre_w[2] * d) : A.two_dO * (A.pre_w[1] * ui - A.pre_w[2] * d);
        st4(Wt + c0, ld4(Wt + c0) + a4);
    }
    for (int tb = 0; tb < nb; ++tb) st4(Bt + tb * PLF + c0, ld4(A.b2[tb] + (long long)e * P + q));
    const float sy = __ldg(gp.sigma + y);
    const bool ytop = y == 0, ybot = y == ny - 1;
    // Dirichlet mask of the four cells as bits (1: the cell is on the domain border, its lU is masked to zero)
    const unsigned mb = ((ytop || ybot) ? 15u : 0u) | ((xg == 0 || xg == nx - 1) ? 1u : 0u) | (xg + 1 == nx - 1 ? 2u : 0u) |
                        (xg + 2 == nx - 1 ? 4u : 0u) | (xg + 3 == nx - 1 ? 8u : 0u);
    // pad columns of this group (bits), kept at zero in the exchanged planes: they are the "v[n] = 0" of the last column
    const unsigned pad = (xg + 1 >= nx ? 2u : 0u) | (xg + 2 >= nx ? 4u : 0u) | (xg + 3 >= nx ? 8u : 0u) | (xg >= nx ? 1u : 0u);
    const bool xl0 = xg == 0;                                        // columns 0..3: nothing to the left, left one-sided corrections
    const bool xr0 = xg + 4 >= nx;                                   // nothing inside the domain to the right of this group
    const bool xrs = R == nx && xg + 3 >= nx - 3 && xg < nx;         // holds one of the three last columns
    const int oN = (nx - 1 - L) - lc;                                // from this group's first float to column nx-1 of its row
    const float cm = gp.g_central[0], cp = gp.g_central[1];
    // coefficient of v[row 0] / v[row ny-1] in (Dy^T v) of this thread's row, and where those rows are (region rows 0 / B-1-T0)
    const float kyT = y == 0 ? gp.g_first[0] : (y == 1 ? gp.g_first[1] - cp : (y == 2 ? gp.g_first[2] : 0.0f));
    const float kyB = y == ny - 3 ? gp.g_last[0] : (y == ny - 2 ? gp.g_last[1] - cm : (y == ny - 1 ? gp.g_last[2] : 0.0f));
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // (D^T v) along x of this thread's four cells; own: its values of the plane (registers or re-read)
    auto dxT = [&](const float *__restrict__ pl, const float4 own) -> float4 {
        const float l = xl0 ? 0.0f : pl[c0 - 1], r = xr0 ? 0.0f : pl[c0 + 4];
        float4 d = cp * make_float4(l - own.y, own.x - own.z, own.y - own.w, own.z - r);   // central rows: cm = -cp
        if (xl0) {
            d.x = fmaf(gp.g_first[0], own.x, d.x);
            d.y = fmaf(gp.g_first[1] - cp, own.x, d.y);
            d.z = fmaf(gp.g_first[2], own.x, d.z);
        }
        if (xrs) {
            const float vN = pl[c0 + oN];
            const int k0 = xg - (nx - 3);   // cell j of the group is column nx-3 + (k0 + j)
            const float g0 = gp.g_last[0], g1 = gp.g_last[1] - cm, g2 = gp.g_last[2];
            auto kr = [&](int k) { return k == 0 ? g0 : (k == 1 ? g1 : (k == 2 ? g2 : 0.0f)); };
            d.x = fmaf(kr(k0), vN, d.x);
            d.y = fmaf(kr(k0 + 1), vN, d.y);
            d.z = fmaf(kr(k0 + 2), vN, d.z);
            d.w = fmaf(kr(k0 + 3), vN, d.w);
        }
        return d;
    };
    auto dyT = [&](const float *__restrict__ pl) -> float4 {
        const float4 u = ytop ? z4 : ld4(pl + c0 - Wp), dn = ybot ? z4 : ld4(pl + c0 + Wp);
        float4 d = cp * (u - dn);
        if (kyT != 0.0f) d = fma4(kyT, ld4(pl + c0 - lr * Wp), d);                  // (y <= 2: the region starts at row 0)
        if (kyB != 0.0f) d = fma4(kyB, ld4(pl + c0 + (ny - 1 - y) * Wp), d);        // (y >= ny-3: the region ends at row ny-1)
        return d;
    };
    auto zero_pad = [&](float4 v) -> float4 {
        if (pad) {
            v.x = (pad & 1u) ? 0.f : v.x;
            v.y = (pad & 2u) ? 0.f : v.y;
            v.z = (pad & 4u) ? 0.f : v.z;
            v.w = (pad & 8u) ? 0.f : v.w;
        }
        return v;
    };
    float4 oU = z4, oVx = z4, oVy = z4, og = z4, sU = z4, sVx = z4, sVy = z4, sg = z4;
    __syncthreads();

    const float dt = gp.dt;
#pragma unroll
    for (int s = 0; s < 4; ++s) {   // stages 4, 3, 2, 1 of the forward step
        const float a = (s == 0 || s == 3) ? dt * (1.0f / 6.0f) : dt * (1.0f / 3.0f);
        const float be = s == 0 ? 0.0f : (s == 1 ? dt : 0.5f * dt);
        const int tau = s == 0 ? 2 : (s == 3 ? 0 : 1);
        // phase A: l = a w + be l_prev; exchange qx = b (m lU + sy lPsiy), qy = b (m lU + sx lPsix), lVx, lVy
        float4 lU = fma4(be, oU, a * ld4(Wt + c0));
        if (mb) {   // (rare: border rows / columns only)
            lU.x = (mb & 1u) ? 0.f : lU.x;
            lU.y = (mb & 2u) ? 0.f : lU.y;
            lU.z = (mb & 4u) ? 0.f : lU.z;
            lU.w = (mb & 8u) ? 0.f : lU.w;
        }
        {
            const float4 sx = ld4(gp.sigma + xg);   // (L1-resident; the profile is allocated with 4 floats of padding)
            const float4 b4 = nb == 0 ? make_float4(gp.b0, gp.b0, gp.b0, gp.b0) : ld4(Bt + (nb == 1 ? 0 : tau) * PLF + c0);
            const float4 lPx = fma4(be, og, a * ld4(Wt + 3 * PLF + c0)), lPy = fma4(be, og, a * ld4(Wt + 4 * PLF + c0));
            st4(Lx + c0, zero_pad(b4 * (lU + sy * lPy)));
            st4(Lx + PLF + c0, zero_pad(b4 * (lU + sx * lPx)));
            st4(Lx + 2 * PLF + c0, zero_pad(fma4(be, oVx, a * ld4(Wt + PLF + c0))));
            st4(Lx + 3 * PLF + c0, zero_pad(fma4(be, oVy, a * ld4(Wt + 2 * PLF + c0))));
        }
        __syncthreads();
        // phase B: l_new = J^T l
        {
            const float *qx = Lx, *qy = Lx + PLF, *lvx = Lx + 2 * PLF, *lvy = Lx + 3 * PLF;
            const float4 sx = ld4(gp.sigma + xg);
            // (one derivative at a time, each result folded into its output at once: the live set stays within 64 registers)
            {   // new lVx = Dx^T qx - sx lVx ; Dx^T lVx is the first part of the new lU
                const float4 vx = ld4(lvx + c0);
                oVx = dxT(qx, ld4(qx + c0)) - sx * vx;
                sVx = sVx + oVx;
                oU = dxT(lvx, vx);
            }
            {   // new lVy = Dy^T qy - sy lVy
                oVy = dyT(qy) - sy * ld4(lvy + c0);
                sVy = sVy + oVy;
            }
            oU = oU + dyT(lvy);
            {
                const float4 lOm = a * ld4(Wt + 5 * PLF + c0) - be * og;
                oU = oU + (sy * sx) * lOm - (sx + make_float4(sy, sy, sy, sy)) * lU;
            }
            sU = sU + oU;
            og = lU;
            sg = sg + lU;
        }
        __syncthreads();
    }
    if (!act || y < y0 || y >= y1) return;
    // owned columns: [x0, x1); a group may straddle x1 (pad columns / a domain whose width is not a multiple of 4)
    float4 vU = ld4(Wt + c0) + sU;
    if (A.post_u) {
        const float *up = A.post_u + (long long)e * 2 * P + q;
        const float4 ut = ld4(up), ui = ld4(up + P), d = ut - ui;
        vU = vU + (wf == 0 ? A.two_dO * (A.post_w[0] * ut + A.post_w[2] * d) : A.two_dO * (A.post_w[1] * ui - A.post_w[2] * d));
    }
    const float4 r1 = ld4(Wt + PLF + c0) + sVx, r2 = ld4(Wt + 2 * PLF + c0) + sVy, r3 = ld4(Wt + 3 * PLF + c0) + sg,
                 r4 = ld4(Wt + 4 * PLF + c0) + sg, r5 = ld4(Wt + 5 * PLF + c0) - sg;
    if (xg >= x0 && xg + 3 < x1) {
        st4(wout + q, vU);
        st4(wout + P + q, r1);
        st4(wout + 2 * P + q, r2);
        st4(wout + 3 * P + q, r3);
        st4(wout + 4 * P + q, r4);
        st4(wout + 5 * P + q, r5);
    } else {
        const float v[6][4] = {{vU.x, vU.y, vU.z, vU.w}, {r1.x, r1.y, r1.z, r1.w}, {r2.x, r2.y, r2.z, r2.w},
                               {r3.x, r3.y, r3.z, r3.w}, {r4.x, r4.y, r4.z, r4.w}, {r5.x, r5.y, r5.z, r5.w}};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (xg + j >= x0 && xg + j < x1)
#pragma unroll
                for (int f = 0; f < 6; ++f) wout[f * P + q + j] = v[f][j];
    }
}

size_t adj_smem_int4(int nb) { return sizeof(float) * ((size_t)(6 + nb) * 4 * ANT_INT + 2 * AGUARD); }
size_t adj_smem_gen4(int nb) { return sizeof(float) * ((size_t)(10 + nb) * 4 * ANT + 2 * AGUARD); }

// gather the two U planes of every environment of a state: [n_env][12][plane] -> [n_env][2][plane]
__global__ void k_gather_u(GridP gp, const float *__restrict__ u, float *__restrict__ out) {
    const long long n = gp.plane, q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y >> 1, f = blockIdx.y & 1;
    if (q < n) out[((long long)e * 2 + f) * n + q] = u[((long long)e * 12 + f * 6) * n + q];
}

// After a sweep whose interior ran on the march (kernels_fused.cu, stage_T): the three auxiliary cotangents of the cells of the
// rectangle [x0, x1) x [y0, y1) receive the accumulated G = sum over the steps of dt/6 (y0 + 2 y1 + 2 y2 + y3)_U.
__global__ void k_apply_aux_cotangent(GridP gp, float *__restrict__ w, const float *__restrict__ G, int x0, int x1, int y0, int y1) {
    const int x = x0 + 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = y0 + blockIdx.y;
    const int e = blockIdx.z >> 1, wf = blockIdx.z & 1;
    if (x >= x1 || y >= y1) return;
    const long long q = (long long)y * gp.nxp + x;
    const float4 g = ld4(G + ((long long)e * 2 + wf) * gp.plane + q);
    float *wp = w + ((long long)e * 12 + wf * 6) * gp.plane + q;
    st4(wp + 3 * gp.plane, ld4(wp + 3 * gp.plane) + g);
    st4(wp + 4 * gp.plane, ld4(wp + 4 * gp.plane) + g);
    st4(wp + 5 * gp.plane, ld4(wp + 5 * gp.plane) - g);
}

}  // namespace

void launch_apply_aux_cotangent(waves_handle *h, float *w, const float *G, const int rect[4]) {
    const int groups = (rect[1] - rect[0] + 3) / 4;
    dim3 grd((groups + 63) / 64, rect[3] - rect[2], h->gp.n_env * 2);
    k_apply_aux_cotangent<<<grd, 64, 0, h->stream>>>(h->gp, w, G, rect[0], rect[1], rect[2], rect[3]);
    h->launches++;
}

static int adj_attr_once() {
    static bool attr_done = false;
    if (attr_done) return 0;
    cudaError_t ce = cudaFuncSetAttribute(k_adjoint_step_int4<ATY_INT, ATX_INT, ANT_INT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)adj_smem_int4(3));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(k_adjoint_step_gen4<ATY_GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem_gen4(3));
    if (ce != cudaSuccess) return waves_set_error("fused reverse step: cudaFuncSetAttribute failed (was the library built for sm_100a?)");
    attr_done = true;
    return 0;
}

// The general tiles of one fused reverse step on the frame around `rect` (whose cells another kernel advances): four bands.
int launch_adjoint_step_frame(waves_handle *h, const float *w_in, float *w_out, const float *const b2[3], const float *pre_u,
                              const float *pre_w3, const float *post_u, const float *post_w3, const int rect[4]) {
    const GridP &gp = h->gp;
    if (adj_attr_once()) return 1;
    AdjFArgs A;
    A.gp = gp;
    A.w_in = w_in;
    A.w_out = w_out;
    for (int t = 0; t < 3; ++t) A.b2[t] = b2 ? b2[t] : nullptr;
    A.nb = (!b2 || !b2[0]) ? 0 : ((b2[0] == b2[1] && b2[1] == b2[2]) ? 1 : 3);
    A.pre_u = pre_u;
    A.post_u = post_u;
    for (int k = 0; k < 3; ++k) {
        A.pre_w[k] = pre_w3 ? pre_w3[k] : 0.0f;
        A.post_w[k] = post_w3 ? post_w3[k] : 0.0f;
    }
    A.two_dO = 2.0f * h->d_omega;
    A.tiles_x = 1;
    A.x_org = A.y_org = 0;
    A.y_end = gp.ny_global;
    A.skip_x0 = A.skip_x1 = A.skip_y0 = A.skip_y1 = 0;
    // left and right bands over the full height, top and bottom bands between them
    const int bx0[4] = {0, rect[1], rect[0], rect[0]}, bx1[4] = {rect[0], gp.nx, rect[1], rect[1]};
    const int by0[4] = {0, 0, 0, rect[3]}, by1[4] = {gp.ny_global, gp.ny_global, rect[2], gp.ny_global};
    int nb = 0, blocks = 0;
    for (int k = 0; k < 4; ++k) {
        if (bx1[k] <= bx0[k] || by1[k] <= by0[k]) continue;
        A.bx0[nb] = bx0[k], A.bx1[nb] = bx1[k], A.by0[nb] = by0[k], A.by1[nb] = by1[k];
        A.btx[nb] = (bx1[k] - bx0[k] + ATX - 1) / ATX;
        // rows per tile: a region of W columns (tile + halos, + 2 where it is extended to a domain edge) is W/4 groups wide, and
        // (rows + 8 halo rows + 2 of an extension) x groups must fit the CTA's 512 threads; the band's height is split evenly
        int gprmax = 1;
        for (int tx = 0; tx < A.btx[nb]; ++tx) {   // (the kernel's own region arithmetic)
            const int x0 = bx0[k] + tx * ATX, x1 = std::min(x0 + ATX, bx1[k]), L = std::max(x0 - 4, 0);
            int R = std::min(x1 + 4, gp.nx);
            if (gp.nx - R <= 2) R = gp.nx;
            gprmax = std::max(gprmax, ((R - L + 3) & ~3) >> 2);
        }
        const int tymax = std::max(4, ANT / gprmax - 10), H = by1[k] - by0[k];
        const int nt = (H + tymax - 1) / tymax;
        A.bty[nb] = (H + nt - 1) / nt;
        A.bstart[nb] = blocks;
        blocks += A.btx[nb] * nt;
        ++nb;
    }
    for (int k = nb; k < 4; ++k) A.bx0[k] = A.bx1[k] = A.by0[k] = A.by1[k] = 0, A.btx[k] = 1, A.bty[k] = ATY_GEN, A.bstart[k] = blocks;
    A.bstart[4] = blocks;
    if (nb < 4) A.bstart[nb] = blocks;
    A.nband = nb;
    if (nb == 0) return 0;
    dim3 grd(blocks, gp.n_env * 2);
    k_adjoint_step_gen4<ATY_GEN><<<grd, ANT, adj_smem_gen4(A.nb), h->stream>>>(A);
    h->launches++;
    return 0;
}

// One fused reverse step for every environment: w_out = [pre-injection] (I + J_step^T) w_in [+ post-injection].
// b2[tau]: c^2 planes of the total field at the three stage times (nullptr: no design anywhere).
int launch_adjoint_step_fused(waves_handle *h, const float *w_in, float *w_out, const float *const b2[3], const float *pre_u,
                              const float *pre_w3, const float *post_u, const float *post_w3) {
    const GridP &gp = h->gp;
    if (gp.ny_own != gp.ny_global) return waves_set_error("fused reverse step: not available on slab handles");
    if (adj_attr_once()) return 1;
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
    A.nb = (!b2 || !b2[0]) ? 0 : ((b2[0] == b2[1] && b2[1] == b2[2]) ? 1 : 3);
    A.pre_u = pre_u;
    A.post_u = post_u;
    for (int k = 0; k < 3; ++k) {
        A.pre_w[k] = pre_w3 ? pre_w3[k] : 0.0f;
        A.post_w[k] = post_w3 ? post_w3[k] : 0.0f;
    }
    A.two_dO = 2.0f * h->d_omega;
    A.nband = 0;
    for (int k = 0; k < 4; ++k) A.bx0[k] = A.bx1[k] = A.by0[k] = A.by1[k] = 0, A.btx[k] = 1, A.bty[k] = ATY_GEN, A.bstart[k] = 0;
    A.bstart[4] = 0;
    if (has_int) {
        A.tiles_x = (xi1 - xi0) / ATX_INT;
        A.x_org = xi0;
        A.y_org = yi0;
        A.y_end = yi1;
        A.skip_x0 = A.skip_x1 = A.skip_y0 = A.skip_y1 = 0;
        const int tiles_y = (yi1 - yi0 + ATY_INT - 1) / ATY_INT;
        dim3 grd(A.tiles_x * tiles_y, gp.n_env * 2);
        k_adjoint_step_int4<ATY_INT, ATX_INT, ANT_INT><<<grd, ANT_INT, adj_smem_int4(A.nb), h->stream>>>(A);
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
        k_adjoint_step_gen4<ATY_GEN><<<grd, ANT, adj_smem_gen4(A.nb), h->stream>>>(A);
        h->launches++;
    }
    return 0;
}

void launch_gather_u(waves_handle *h, const float *u, float *out) {
    const GridP &gp = h->gp;
    if ((gp.plane & 3) == 0 && (((uintptr_t)u | (uintptr_t)out) & 15u) == 0) {
        // two strided block copies (U of the total field, U of the incident field) at the HBM copy rate (float4, streaming)
        launch_copy_blocks(h, u, out, (long long)gp.plane, (long long)gp.env_stride, 2LL * gp.plane, gp.n_env);
        launch_copy_blocks(h, u + 6 * gp.plane, out + gp.plane, (long long)gp.plane, (long long)gp.env_stride, 2LL * gp.plane, gp.n_env);
        return;
    }
    dim3 grd((unsigned)((gp.plane + 255) / 256), gp.n_env * 2);
    k_gather_u<<<grd, 256, 0, h->stream>>>(gp, u, out);
    h->launches++;
}

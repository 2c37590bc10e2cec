// Fused RK4 step kernel for sm_100a: ONE launch advances every cell of every environment by a full
// runge_kutta step (src/dynamics.jl:9-16) -- 4 evaluations of acoustic_dynamics for the total and the
// incident wavefield (src/dynamics.jl:151-188), the design speed field (src/designs.jl:99-116) at the
// three distinct stage times, the Gaussian*sine source (src/sources.jl:67-69), PML and Dirichlet mask,
// the RK4 combination, and the energy metric of the new frame (src/env.jl:104-111).
//
// Design:
//  * work item = one WARP = a strip of 32 columns marched along y through a slab of rows;
//    lane <-> column, so d/dy neighbours live in the lane's own registers and d/dx neighbours come
//    from warp shuffles.  No block-level synchronisation exists in the kernel.
//  * temporal blocking: the 4 RK stages are software-pipelined along y (stage s works one row behind
//    stage s-1), so a row of state is read from HBM once and written once per step (96 B/cell-update).
//    The 4-column / 4-row halo per side is recomputed redundantly (24 valid of 32 lanes).
//  * rows are fetched by TMA (cp.async.bulk.tensor.3d, box 32 x 1 x 12 fields) into a per-warp
//    shared-memory ring, two rows ahead, completion tracked by mbarriers; out-of-range columns/rows
//    are zero-filled by the TMA unit.
//  * rotating register windows are indexed by (row & 3) with the row loop unrolled by 4, so all
//    window indices are compile-time constants (no register moves).
//  * domain-border rows (one-sided 3-point stencils, src/operators.jl:3-4) are processed together
//    with their inward neighbour by a small out-of-line routine that keeps its state in shared memory.
//  * GENERAL=false is the lean interior variant (sigma == 0 everywhere in the warp's window): Psi/Omega
//    pass through unchanged; GENERAL=true handles PML strips and domain borders.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "waves_internal.h"

namespace {

constexpr int RING = 8;       // ring slots (rows) per warp
constexpr int PF = 2;         // TMA prefetch distance in rows
constexpr int SLOT_F = 13 * 32;  // floats per ring slot: 12 fields + source shape
constexpr int CYL_CAP = 12;   // culled cylinders kept per warp
constexpr int BSET_F = 24 * 32;  // border-row state per warp (floats)
constexpr int WARPS = 4;      // warps per CTA (independent work items)

struct Item {
    int x0;        // first column of the 32-lane window
    int vlo, vhi;  // lanes [vlo, vhi) own output columns
    int j0, j1;    // output local rows [j0, j1)
    int la, lb;    // loaded local rows [la, lb)
    int top, bot;  // window touches the domain's first / last row
};

struct FusedArgs {
    GridP gp;
    const EnvParams *env;
    const float *cyl0, *cyl1;
    int cyl_cap;
    const float *table;
    int steps, step;
    float *out;
    const Item *items;
    int n_items;
    float *epart;  // [n_env][n_items_total][3] or nullptr
    int epart_stride;  // items per env in epart (all kernels of a step share one buffer)
    int epart_off;     // offset of this kernel's items
    int dbg;           // developer bisecting flags (WAVES_DEBUG_FLAGS)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// DesignInterpolator(t) for one parameter, exact float32 order (src/designs.jl:287-292)
__device__ __forceinline__ float interp_param(float p0, float p1, float ti, float tf, float t) {
    float dt = __fsub_rn(tf, ti);
    dt = dt > 0.0f ? dt : 1.0f;
    float inv = __fdiv_rn(1.0f, dt);
    float tc = fminf(fmaxf(t, ti), tf);
    float s = __fsub_rn(tc, ti);
    float dy = __fadd_rn(p1, __fmul_rn(p0, -1.0f));
    float slope = __fmul_rn(dy, inv);
    return __fadd_rn(p0, __fmul_rn(slope, s));
}

struct WarpCtx {
    // geometry
    int lane, col, x0, la, lb, j0, j1;
    bool valid_lane, xb, top, bot, is_first_col, is_last_col;
    float xs, sx, bcm;
    // step constants
    float kd, dt, hdt, b0;
    float sf[3];  // source factor at t, t+dt/2, t+dt
    // smem
    float *ring;   // [RING][13][32]
    float *bset;   // [24][32]
    float *cyl;    // [CYL_CAP][3][4]
    uint32_t bar0; // smem address of mbarrier 0
    int nact;      // active cylinders (-1: overflow, use the slow loop)
    // source
    bool src_cols;
    int src_j0, src_j1;
    // global
    const float *ys_g, *sig_g;
    int grow0, ny_global;
};

__device__ __forceinline__ int ring_slot(const WarpCtx &c, int j) { return (j - c.la) & (RING - 1); }
__device__ __forceinline__ const float *ring_row(const WarpCtx &c, int j) { return c.ring + ring_slot(c, j) * SLOT_F + c.lane; }
__device__ __forceinline__ bool src_row(const WarpCtx &c, int j) { return c.src_cols && j >= c.src_j0 && j < c.src_j1; }
__device__ __forceinline__ float sh_at(const WarpCtx &c, int j) { return src_row(c, j) ? ring_row(c, j)[12 * 32] : 0.0f; }

// c(x,y,t)^2 for this lane at local row j and stage-time index tau (src/designs.jl:99-116), exact order
__device__ __forceinline__ float speed2(const WarpCtx &c, const FusedArgs &A, const EnvParams &ep, int e, int j, int tau) {
    if (c.nact == 0) return c.b0;
    float yv = c.ys_g[min(max(c.grow0 + j, 0), c.ny_global - 1)];
    int cnt = 0;
    float cd = 0.0f;
    if (c.nact > 0) {
        for (int a = 0; a < c.nact; ++a) {
            const float4 p = *reinterpret_cast<const float4 *>(c.cyl + (a * 3 + tau) * 4);  // px, py, r2, c
            float dy = __fsub_rn(yv, p.y);
            float dy2 = __fmul_rn(dy, dy);
            if (dy2 >= p.z) continue;  // warp-uniform: the row misses this cylinder
            float dx = __fsub_rn(c.xs, p.x);
            float d2 = __fadd_rn(__fmul_rn(dx, dx), dy2);
            bool m = d2 < p.z;
            cnt += m;
            cd = __fadd_rn(cd, m ? p.w : 0.0f);
        }
    } else {  // overflow of the per-warp list: evaluate every cylinder
        float t = A.table[((size_t)e * A.steps + A.step) * STAGE_ROW + tau];
        for (int k = 0; k < ep.ncyl; ++k) {
            size_t o = ((size_t)e * A.cyl_cap + k) * 4;
            float px = interp_param(A.cyl0[o], A.cyl1[o], ep.ti, ep.tf, t);
            float py = interp_param(A.cyl0[o + 1], A.cyl1[o + 1], ep.ti, ep.tf, t);
            float r = interp_param(A.cyl0[o + 2], A.cyl1[o + 2], ep.ti, ep.tf, t);
            float cc = interp_param(A.cyl0[o + 3], A.cyl1[o + 3], ep.ti, ep.tf, t);
            float dx = __fsub_rn(c.xs, px), dy = __fsub_rn(yv, py);
            float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            bool m = d2 < __fmul_rn(r, r);
            cnt += m;
            cd = __fadd_rn(cd, m ? cc : 0.0f);
        }
    }
    float cv = __fadd_rn(cnt == 0 ? A.gp.c0 : 0.0f, cd);
    return __fmul_rn(cv, cv);
}

// d/dx across lanes: central (src/operators.jl:5) or the one-sided border rows (:3-4) on the domain's edge columns
template <bool GENERAL>
__device__ __forceinline__ float ddx(const WarpCtx &c, const GridP &gp, float v) {
    float e1 = __shfl_down_sync(0xffffffffu, v, 1), w1 = __shfl_up_sync(0xffffffffu, v, 1);
    float d = c.kd * (e1 - w1);
    if (GENERAL) {
        if (c.xb) {
            float e2 = __shfl_down_sync(0xffffffffu, v, 2), w2 = __shfl_up_sync(0xffffffffu, v, 2);
            if (c.is_first_col) d = ((gp.g_first[0] * v) + (gp.g_first[1] * e1)) + (gp.g_first[2] * e2);
            if (c.is_last_col) d = ((gp.g_last[0] * w2) + (gp.g_last[1] * w1)) + (gp.g_last[2] * v);
        }
    }
    return d;
}

// Register state of one warp: rotating windows indexed [stage][wavefield][row & 3]
template <bool GENERAL>
struct Regs {
    float Uf[4][2][4];  // U + f of stage state y_s (s = 0: the loaded row)
    float Vy[4][2][4];
    float Vx[4][2][4];  // s = 1..3
    float aU[2][4], aVx[2][4], aVy[2][4];  // k1 + 2k2 + 2k3 accumulators
    // GENERAL only
    float Uc[4][2][4], Px[4][2][4], Py[4][2][4], Om[4][2][4];
    float aPx[2][4], aPy[2][4], aOm[2][4];
    float e_tot, e_inc, e_sc;
};

// Border-row state kept in shared memory (per warp): index = field * 32 + lane
// fields: for wf w (0,1): base = w*12: 0 Vx, 1 Uc, 2 Px, 3 Py, 4 Om, 5 aU, 6 aVx, 7 aVy, 8 aPx, 9 aPy, 10 aOm, 11 spare
enum { B_VX = 0, B_UC, B_PX, B_PY, B_OM, B_AU, B_AVX, B_AVY, B_APX, B_APY, B_AOM };

// One RK stage on a domain-border row (global row 0 or ny-1).  v3[w][0..2] / f3[w][0..2] hold
// Vy and U+f of the previous stage state on the three rows the one-sided stencil spans, border row
// first for TOP (rows 0,1,2) and last for BOT (rows L-2,L-1,L).  Returns the new Uf / Vy of the border row.
__device__ __noinline__ void border_row_stage(const WarpCtx c, const FusedArgs &A, const EnvParams ep, int e, int S,
                                              bool top, int jb, const float (*f3)[3], const float (*v3)[3],
                                              float *uf_out, float *vy_out, float *e3) {
    const GridP &gp = A.gp;
    const float *g = top ? gp.g_first : gp.g_last;
    const float *ur = ring_row(c, jb);
    const float sy = c.sig_g[c.grow0 + jb];
    const float sx = c.sx;
    const float a = (S == 3) ? c.dt : c.hdt;
    const int tau = (S == 1) ? 0 : (S == 4 ? 2 : 1);
    const float shv = sh_at(c, jb);
    const float sf_next = (S == 1 || S == 2) ? c.sf[1] : c.sf[2];
    const float btot = speed2(c, A, ep, e, jb, tau);
    float outU[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float *bs = c.bset + w * 12 * 32 + c.lane;
        const float *uw = ur + w * 6 * 32;
        const float uU = uw[0], uVx = uw[32], uVy = uw[64], uPx = uw[96], uPy = uw[128], uOm = uw[160];
        const float b = (w == 0) ? btot : c.b0;
        const int ib = top ? 0 : 2;  // position of the border row in f3/v3
        float ufC = f3[w][ib], vyC = v3[w][ib];
        float vxC = (S == 1) ? uVx : bs[B_VX * 32];
        float uC = (S == 1) ? uU : bs[B_UC * 32];
        float px = (S == 1) ? uPx : bs[B_PX * 32];
        float py = (S == 1) ? uPy : bs[B_PY * 32];
        float om = (S == 1) ? uOm : bs[B_OM * 32];
        float Ux = ddx<true>(c, gp, ufC);
        float Vxx = ddx<true>(c, gp, vxC);
        float Uy = ((g[0] * f3[w][0]) + (g[1] * f3[w][1])) + (g[2] * f3[w][2]);
        float Vyy = ((g[0] * v3[w][0]) + (g[1] * v3[w][1])) + (g[2] * v3[w][2]);
        // bc == 0 on a border row (src/dims.jl:117-124): dU = 0
        float kU = 0.0f;
        float kVx = Ux - sx * vxC;
        float kVy = Uy - sy * vyC;
        float kPx = (b * sx) * Vyy;
        float kPy = (b * sy) * Vxx;
        float kOm = (sx * sy) * uC;
        if (S < 4) {
            float Us = uU + a * kU;
            uf_out[w] = Us + shv * sf_next;
            vy_out[w] = uVy + a * kVy;
            bs[B_VX * 32] = uVx + a * kVx;
            bs[B_UC * 32] = Us;
            bs[B_PX * 32] = uPx + a * kPx;
            bs[B_PY * 32] = uPy + a * kPy;
            bs[B_OM * 32] = uOm + a * kOm;
            if (S == 1) {
                bs[B_AU * 32] = kU;
                bs[B_AVX * 32] = kVx;
                bs[B_AVY * 32] = kVy;
                bs[B_APX * 32] = kPx;
                bs[B_APY * 32] = kPy;
                bs[B_AOM * 32] = kOm;
            } else {
                bs[B_AU * 32] += 2.0f * kU;
                bs[B_AVX * 32] += 2.0f * kVx;
                bs[B_AVY * 32] += 2.0f * kVy;
                bs[B_APX * 32] += 2.0f * kPx;
                bs[B_APY * 32] += 2.0f * kPy;
                bs[B_AOM * 32] += 2.0f * kOm;
            }
        } else {
            const float sixth = 1.0f / 6.0f;
            float oU = uU + (sixth * (bs[B_AU * 32] + kU)) * c.dt;
            float oVx = uVx + (sixth * (bs[B_AVX * 32] + kVx)) * c.dt;
            float oVy = uVy + (sixth * (bs[B_AVY * 32] + kVy)) * c.dt;
            float oPx = uPx + (sixth * (bs[B_APX * 32] + kPx)) * c.dt;
            float oPy = uPy + (sixth * (bs[B_APY * 32] + kPy)) * c.dt;
            float oOm = uOm + (sixth * (bs[B_AOM * 32] + kOm)) * c.dt;
            outU[w] = oU;
            if (c.valid_lane && jb >= c.j0 && jb < c.j1) {
                float *o = A.out + (long long)e * gp.env_stride + (long long)w * 6 * gp.plane + (long long)jb * gp.nxp + c.col;
                o[0] = oU;
                o[gp.plane] = oVx;
                o[2 * gp.plane] = oVy;
                o[3 * gp.plane] = oPx;
                o[4 * gp.plane] = oPy;
                o[5 * gp.plane] = oOm;
            }
        }
    }
    if (S == 4 && c.valid_lane && jb >= c.j0 && jb < c.j1) {
        float d = outU[0] - outU[1];
        e3[0] += outU[0] * outU[0];
        e3[1] += outU[1] * outU[1];
        e3[2] += d * d;
    }
    __syncwarp();
}

// One RK stage S (1..4) on interior row j = r - S, for both wavefields.  PH = r & 3.
template <bool GENERAL, int S, int PH>
__device__ __forceinline__ void stage(const WarpCtx &c, const FusedArgs &A, const EnvParams &ep, int e, Regs<GENERAL> &R, int j) {
    const GridP &gp = A.gp;
    constexpr int sc = (PH - S + 8) & 3, sm = (sc + 3) & 3, sp = (sc + 1) & 3;  // slots of rows j, j-1, j+1
    const float *ur = ring_row(c, j);
    const float a = (S == 3) ? c.dt : c.hdt;
    constexpr int tau = (S == 1) ? 0 : (S == 4 ? 2 : 1);
    const float sy = GENERAL ? c.sig_g[c.grow0 + j] : 0.0f;
    const float sx = c.sx;
    const float shv = (S < 4) ? sh_at(c, j) : 0.0f;
    const float sf_next = (S == 1 || S == 2) ? c.sf[1] : c.sf[2];
    const float btot = speed2(c, A, ep, e, j, tau);
    float outU[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const float *uw = ur + w * 6 * 32;
        const float uU = uw[0], uVx = uw[32], uVy = uw[64];
        const float b = (w == 0) ? btot : c.b0;
        const float ufC = R.Uf[S - 1][w][sc];
        const float vxC = (S == 1) ? uVx : R.Vx[S - 1][w][sc];
        const float Ux = ddx<GENERAL>(c, gp, ufC);
        const float Vxx = ddx<GENERAL>(c, gp, vxC);
        const float Uy = c.kd * (R.Uf[S - 1][w][sp] - R.Uf[S - 1][w][sm]);
        const float Vyy = c.kd * (R.Vy[S - 1][w][sp] - R.Vy[S - 1][w][sm]);
        float kU, kVx, kVy, kPx = 0.f, kPy = 0.f, kOm = 0.f;
        float uPx = 0.f, uPy = 0.f, uOm = 0.f;
        if (GENERAL) {
            uPx = uw[96];
            uPy = uw[128];
            uOm = uw[160];
            const float vyC = R.Vy[S - 1][w][sc];
            const float uC = (S == 1) ? uU : R.Uc[S - 1][w][sc];
            const float px = (S == 1) ? uPx : R.Px[S - 1][w][sc];
            const float py = (S == 1) ? uPy : R.Py[S - 1][w][sc];
            const float om = (S == 1) ? uOm : R.Om[S - 1][w][sc];
            kU = c.bcm * (((((b * (Vxx + Vyy)) + px) + py) - ((sx + sy) * uC)) - om);
            kVx = Ux - sx * vxC;
            kVy = Uy - sy * vyC;
            kPx = (b * sx) * Vyy;
            kPy = (b * sy) * Vxx;
            kOm = (sx * sy) * uC;
        } else {
            // sigma == 0 in the whole window: Psi, Omega never change within the step
            const float P = (uw[96] + uw[128]) - uw[160];
            kU = (b * (Vxx + Vyy)) + P;
            kVx = Ux;
            kVy = Uy;
        }
        if (S < 4) {
            const float Us = uU + a * kU;
            R.Uf[S][w][sc] = Us + shv * sf_next;
            R.Vx[S][w][sc] = uVx + a * kVx;
            R.Vy[S][w][sc] = uVy + a * kVy;
            if (GENERAL) {
                R.Uc[S][w][sc] = Us;
                R.Px[S][w][sc] = uPx + a * kPx;
                R.Py[S][w][sc] = uPy + a * kPy;
                R.Om[S][w][sc] = uOm + a * kOm;
            }
            if (S == 1) {
                R.aU[w][sc] = kU;
                R.aVx[w][sc] = kVx;
                R.aVy[w][sc] = kVy;
                if (GENERAL) {
                    R.aPx[w][sc] = kPx;
                    R.aPy[w][sc] = kPy;
                    R.aOm[w][sc] = kOm;
                }
            } else {
                R.aU[w][sc] += 2.0f * kU;
                R.aVx[w][sc] += 2.0f * kVx;
                R.aVy[w][sc] += 2.0f * kVy;
                if (GENERAL) {
                    R.aPx[w][sc] += 2.0f * kPx;
                    R.aPy[w][sc] += 2.0f * kPy;
                    R.aOm[w][sc] += 2.0f * kOm;
                }
            }
        } else {
            const float sixth = 1.0f / 6.0f;
            const float oU = uU + (sixth * (R.aU[w][sc] + kU)) * c.dt;
            const float oVx = uVx + (sixth * (R.aVx[w][sc] + kVx)) * c.dt;
            const float oVy = uVy + (sixth * (R.aVy[w][sc] + kVy)) * c.dt;
            float oPx, oPy, oOm;
            if (GENERAL) {
                oPx = uPx + (sixth * (R.aPx[w][sc] + kPx)) * c.dt;
                oPy = uPy + (sixth * (R.aPy[w][sc] + kPy)) * c.dt;
                oOm = uOm + (sixth * (R.aOm[w][sc] + kOm)) * c.dt;
            } else {
                oPx = uw[96];
                oPy = uw[128];
                oOm = uw[160];
            }
            outU[w] = oU;
            if (c.valid_lane && j >= c.j0 && j < c.j1) {
                float *o = A.out + (long long)e * gp.env_stride + (long long)w * 6 * gp.plane + (long long)j * gp.nxp + c.col;
                o[0] = oU;
                o[gp.plane] = oVx;
                o[2 * gp.plane] = oVy;
                o[3 * gp.plane] = oPx;
                o[4 * gp.plane] = oPy;
                o[5 * gp.plane] = oOm;
            }
        }
    }
    if (S == 4 && c.valid_lane && j >= c.j0 && j < c.j1) {
        const float d = outU[0] - outU[1];
        R.e_tot += outU[0] * outU[0];
        R.e_inc += outU[1] * outU[1];
        R.e_sc += d * d;
    }
    if (GENERAL) {
        // domain-border rows ride along with their inward neighbour (see file header)
        const bool do_top = c.top && j == 1, do_bot = c.bot && j == c.lb - 2;
        if (do_top || do_bot) {
            float f3[2][3], v3[2][3], ufo[2], vyo[2], e3[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                f3[w][0] = R.Uf[S - 1][w][sm];
                f3[w][1] = R.Uf[S - 1][w][sc];
                f3[w][2] = R.Uf[S - 1][w][sp];
                v3[w][0] = R.Vy[S - 1][w][sm];
                v3[w][1] = R.Vy[S - 1][w][sc];
                v3[w][2] = R.Vy[S - 1][w][sp];
            }
            if (do_top) {
                border_row_stage(c, A, ep, e, S, true, c.la, f3, v3, ufo, vyo, e3);
                if (S < 4) {
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        R.Uf[S][w][sm] = ufo[w];
                        R.Vy[S][w][sm] = vyo[w];
                    }
                }
            }
            if (do_bot) {
                border_row_stage(c, A, ep, e, S, false, c.lb - 1, f3, v3, ufo, vyo, e3);
                if (S < 4) {
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        R.Uf[S][w][sp] = ufo[w];
                        R.Vy[S][w][sp] = vyo[w];
                    }
                }
            }
            R.e_tot += e3[0];
            R.e_inc += e3[1];
            R.e_sc += e3[2];
        }
    }
}

template <bool GENERAL, int PH>
__device__ __forceinline__ void row_step(const WarpCtx &c, const FusedArgs &A, const EnvParams &ep, int e, Regs<GENERAL> &R, int r,
                                         const CUtensorMap *map_u, const CUtensorMap *map_sh) {
    // 1. prefetch row r + PF into the slot that row r + PF - RING vacated (all its readers finished >= 1 step ago)
    __syncwarp();
    const int rp = r + PF;
    if (c.lane == 0 && rp >= c.la && rp < c.lb && !(A.dbg & 2)) {
        const int slot = (rp - c.la) & (RING - 1);
        const uint32_t bar = c.bar0 + slot * 8, dst = smem_u32(c.ring + slot * SLOT_F);
        const bool s = src_row(c, rp);
        mbar_expect_tx(bar, 12 * 128 + (s ? 128 : 0));
        tma_load_3d(dst, map_u, c.x0, rp, e * 12, bar);
        if (s) tma_load_3d(dst + 12 * 128, map_sh, c.x0, rp, e, bar);
    }
    // 2. arrival of row r: stage-0 windows
    if (r >= c.la && r < c.lb) {
        const int rel = r - c.la;
        if (!(A.dbg & 2)) mbar_wait(c.bar0 + (rel & (RING - 1)) * 8, (rel >> 3) & 1);
        const float *ur = ring_row(c, r);
        const float shv = sh_at(c, r);
        constexpr int s0 = PH & 3;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            R.Uf[0][w][s0] = ur[w * 6 * 32] + shv * c.sf[0];
            R.Vy[0][w][s0] = ur[w * 6 * 32 + 64];
        }
    }
    // 3. the four stages, each one row behind the previous
    const int jlo = c.la + 1, jhi = c.lb - 2;  // rows with both y-neighbours loaded
    if (A.dbg & 1) return;
    if (r - 1 >= jlo && r - 1 <= jhi) stage<GENERAL, 1, PH>(c, A, ep, e, R, r - 1);
    if (r - 2 >= jlo && r - 2 <= jhi) stage<GENERAL, 2, PH>(c, A, ep, e, R, r - 2);
    if (r - 3 >= jlo && r - 3 <= jhi) stage<GENERAL, 3, PH>(c, A, ep, e, R, r - 3);
    if (r - 4 >= jlo && r - 4 <= jhi) stage<GENERAL, 4, PH>(c, A, ep, e, R, r - 4);
}

template <bool GENERAL>
__global__ void __launch_bounds__(WARPS * 32, GENERAL ? 2 : 3)
k_fused_step(const __grid_constant__ FusedArgs A, const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_sh) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * WARPS + warp;
    const int e = (int)(gw / A.n_items), it = (int)(gw - (long long)e * A.n_items);
    if (e >= A.gp.n_env) return;
    const GridP &gp = A.gp;
    const Item item = A.items[it];
    const EnvParams ep = A.env[e];

    // per-warp shared memory carve-up
    constexpr int WARP_BYTES = RING * SLOT_F * 4 + BSET_F * 4 + CYL_CAP * 12 * 4 + RING * 8;
    unsigned char *wbase = smem_raw + (size_t)warp * ((WARP_BYTES + 127) & ~127);
    WarpCtx c;
    c.ring = reinterpret_cast<float *>(wbase);
    c.bset = c.ring + RING * SLOT_F;
    c.cyl = c.bset + BSET_F;
    c.bar0 = smem_u32(c.cyl + CYL_CAP * 12);
    c.lane = lane;
    c.x0 = item.x0;
    c.col = item.x0 + lane;
    c.la = item.la;
    c.lb = item.lb;
    c.j0 = item.j0;
    c.j1 = item.j1;
    c.top = item.top;
    c.bot = item.bot;
    c.valid_lane = lane >= item.vlo && lane < item.vhi;
    c.is_first_col = c.col == 0;
    c.is_last_col = c.col == gp.nx - 1;
    c.xb = (item.x0 == 0) || (item.x0 + 32 >= gp.nx);
    c.bcm = (c.is_first_col || c.is_last_col) ? 0.0f : 1.0f;
    c.xs = gp.x[min(c.col, gp.nx - 1)];  // lanes past the last column (nx % 4 != 0) are never valid
    c.sx = GENERAL ? gp.sigma[min(c.col, gp.nx - 1)] : 0.0f;
    c.kd = gp.g_central[1];
    c.dt = gp.dt;
    c.hdt = gp.hdt;
    c.b0 = gp.b0;
    c.ys_g = gp.y;
    c.sig_g = gp.sigma;
    c.grow0 = gp.grow0;
    c.ny_global = gp.ny_global;
    const float *trow = A.table + ((size_t)e * A.steps + A.step) * STAGE_ROW;
    c.sf[0] = trow[3];
    c.sf[1] = trow[4];
    c.sf[2] = trow[5];
    c.src_cols = ep.has_source && item.x0 < ep.src_i1 && item.x0 + 32 > ep.src_i0;
    c.src_j0 = ep.src_j0;
    c.src_j1 = ep.src_j1;

    if (lane == 0 && !(A.dbg & 8)) {
        for (int s = 0; s < RING; ++s) mbar_init(c.bar0 + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // cull the design's cylinders against this warp's window, at the three stage times (src/designs.jl:287-292)
    c.nact = 0;
    if (ep.ncyl > 0 && !(A.dbg & 4)) {
        const float xlo = gp.x[item.x0], xhi = gp.x[min(item.x0 + 31, gp.nx - 1)];
        const float ylo = gp.y[min(max(gp.grow0 + item.la, 0), gp.ny_global - 1)];
        const float yhi = gp.y[min(max(gp.grow0 + item.lb - 1, 0), gp.ny_global - 1)];
        int n = 0;
        for (int k0 = 0; k0 < ep.ncyl; k0 += 32) {
            const int k = k0 + lane;
            float P[3][4];
            bool hit = false;
            if (k < ep.ncyl) {
                const size_t o = ((size_t)e * A.cyl_cap + k) * 4;
                float rmax = 0.f, pxmin = 1e30f, pxmax = -1e30f, pymin = 1e30f, pymax = -1e30f;
#pragma unroll
                for (int tau = 0; tau < 3; ++tau) {
                    const float t = trow[tau];
                    P[tau][0] = interp_param(A.cyl0[o], A.cyl1[o], ep.ti, ep.tf, t);
                    P[tau][1] = interp_param(A.cyl0[o + 1], A.cyl1[o + 1], ep.ti, ep.tf, t);
                    const float r = interp_param(A.cyl0[o + 2], A.cyl1[o + 2], ep.ti, ep.tf, t);
                    P[tau][2] = __fmul_rn(r, r);
                    P[tau][3] = interp_param(A.cyl0[o + 3], A.cyl1[o + 3], ep.ti, ep.tf, t);
                    rmax = fmaxf(rmax, fabsf(r));
                    pxmin = fminf(pxmin, P[tau][0]);
                    pxmax = fmaxf(pxmax, P[tau][0]);
                    pymin = fminf(pymin, P[tau][1]);
                    pymax = fmaxf(pymax, P[tau][1]);
                }
                const float m = rmax * 1.0001f + 1e-6f;
                hit = (pxmax + m >= xlo) && (pxmin - m <= xhi) && (pymax + m >= ylo) && (pymin - m <= yhi);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            const int pos = n + __popc(bal & ((1u << lane) - 1));
            if (hit && pos < CYL_CAP) {
#pragma unroll
                for (int tau = 0; tau < 3; ++tau)
                    *reinterpret_cast<float4 *>(c.cyl + (pos * 3 + tau) * 4) = make_float4(P[tau][0], P[tau][1], P[tau][2], P[tau][3]);
            }
            n += __popc(bal);
        }
        c.nact = n <= CYL_CAP ? n : -1;
        __syncwarp();
    }

    Regs<GENERAL> R;
    R.e_tot = R.e_inc = R.e_sc = 0.0f;
    // zero-init windows so never-consumed warm-up lanes hold finite values
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                R.Uf[s][w][q] = 0.f;
                R.Vy[s][w][q] = 0.f;
                R.Vx[s][w][q] = 0.f;
                R.Uc[s][w][q] = 0.f;
                R.Px[s][w][q] = 0.f;
                R.Py[s][w][q] = 0.f;
                R.Om[s][w][q] = 0.f;
            }
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            R.aU[w][q] = R.aVx[w][q] = R.aVy[w][q] = 0.f;
            R.aPx[w][q] = R.aPy[w][q] = R.aOm[w][q] = 0.f;
        }

    const int r_begin = (c.la - PF) & ~3, r_end = c.lb + 4;  // the first PF steps only prefetch
    for (int r = r_begin; r < r_end; r += 4) {
        row_step<GENERAL, 0>(c, A, ep, e, R, r, &map_u, &map_sh);
        row_step<GENERAL, 1>(c, A, ep, e, R, r + 1, &map_u, &map_sh);
        row_step<GENERAL, 2>(c, A, ep, e, R, r + 2, &map_u, &map_sh);
        row_step<GENERAL, 3>(c, A, ep, e, R, r + 3, &map_u, &map_sh);
    }

    if (A.epart) {
        float et = R.e_tot, ei = R.e_inc, es = R.e_sc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            et += __shfl_down_sync(0xffffffffu, et, o);
            ei += __shfl_down_sync(0xffffffffu, ei, o);
            es += __shfl_down_sync(0xffffffffu, es, o);
        }
        if (lane == 0) {
            float *p = A.epart + ((size_t)e * A.epart_stride + A.epart_off + it) * 3;
            p[0] = et;
            p[1] = ei;
            p[2] = es;
        }
    }
}

// Sum the per-warp energy partials of one step in a fixed order (deterministic), f64, -> (tot, inc, sc) * dΩ
__global__ void k_energy_reduce(const float *__restrict__ part, int n_items, float d_omega, float *__restrict__ e3, int env_stride3) {
    const int e = blockIdx.x;
    __shared__ double sm[3][32];
    double s[3] = {0, 0, 0};
    for (int k = threadIdx.x; k < n_items; k += blockDim.x) {
        const float *p = part + ((size_t)e * n_items + k) * 3;
        s[0] += p[0];
        s[1] += p[1];
        s[2] += p[2];
    }
    for (int q = 0; q < 3; ++q) {
        for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_down_sync(0xffffffffu, s[q], o);
        if ((threadIdx.x & 31) == 0) sm[q][threadIdx.x >> 5] = s[q];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sm[threadIdx.x][k];
        e3[(size_t)e * env_stride3 + threadIdx.x] = (float)t * d_omega;
    }
}

__global__ void k_bbox(GridP gp, const float *__restrict__ shape, int *__restrict__ bb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= gp.nx || j >= gp.ny_alloc) return;
    if (shape[(size_t)j * gp.nxp + i] != 0.0f) {
        atomicMin(bb + 0, i);
        atomicMax(bb + 1, i + 1);
        atomicMin(bb + 2, j);
        atomicMax(bb + 3, j + 1);
    }
}

// ---- host side -------------------------------------------------------------------------------
struct FusedPlan {
    Item *d_items_int = nullptr, *d_items_gen = nullptr;
    int n_int = 0, n_gen = 0;
    float *d_epart = nullptr;
    int *d_bb = nullptr;
    int smem_bytes = 0;
};

FusedPlan *plan_of(waves_handle *h, bool create) {
    if (!h->plan && create) h->plan = new FusedPlan();
    return static_cast<FusedPlan *>(h->plan);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(PFN_encodeTiled enc, CUtensorMap *m, float *base, const GridP &gp, int planes, int box_planes) {
    cuuint64_t dims[3] = {(cuuint64_t)gp.nx, (cuuint64_t)gp.ny_alloc, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)gp.nxp * 4, (cuuint64_t)gp.plane * 4};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)box_planes};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace

int source_bbox(waves_handle *h, int env, int *bb) {
    FusedPlan *p = plan_of(h, false);
    if (!p) return 1;
    int init[4] = {1 << 30, 0, 1 << 30, 0};
    cudaMemcpyAsync(p->d_bb, init, sizeof(init), cudaMemcpyHostToDevice, h->stream);
    dim3 blk(32, 8), grd((h->gp.nx + 31) / 32, (h->gp.ny_alloc + 7) / 8);
    k_bbox<<<grd, blk, 0, h->stream>>>(h->gp, h->shape + (size_t)env * h->gp.plane, p->d_bb);
    h->launches++;
    cudaMemcpyAsync(bb, p->d_bb, sizeof(init), cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return 1;
    if (bb[1] == 0) bb[0] = bb[1] = bb[2] = bb[3] = 0;
    return 0;
}

// Build the work-item lists: 24-column strips x row slabs, split into interior items (sigma == 0 in the
// whole 32 x rows window, no domain border) and general items.
int fused_prepare(waves_handle *h) {
    const GridP &gp = h->gp;
    FusedPlan *p = plan_of(h, true);
    if (!p) return waves_set_error("fused_prepare: too many handles");
    std::vector<float> sig(gp.nx);
    cudaMemcpy(sig.data(), h->d_sigma, sizeof(float) * gp.nx, cudaMemcpyDeviceToHost);

    // zero-sigma zone [z0, z1) of the profile (used for columns and, by global row, for rows)
    int z0 = 0, z1 = gp.nx;
    while (z0 < gp.nx && sig[z0] != 0.0f) ++z0;
    while (z1 > z0 && sig[z1 - 1] != 0.0f) --z1;
    bool clean = true;
    for (int i = z0; i < z1; ++i) clean = clean && sig[i] == 0.0f;
    if (!clean) z0 = z1 = 0;  // unusual profile: everything general

    // --- columns: output ranges [olo, ohi) ---
    // MEASURED ON B200: cp.async.bulk.tensor (tiled) raises "illegal instruction" unless the innermost
    // coordinate is 16-byte aligned, so every window starts at a multiple of 4 columns (scripts/dbg/tma_ring.cu).
    struct Col { int olo, ohi, x0; bool interior; };
    std::vector<Col> cols;
    auto add_cols = [&](int a, int b, bool interior) {  // a is a multiple of 4
        int o = a;
        while (o < b) {
            const int x0 = o == 0 ? 0 : o - 4;
            int hi = x0 + 28;  // 24 owned columns + 4 halo (28 at the left domain edge)
            if (hi > b) hi = b;
            cols.push_back({o, hi, x0, interior});
            o = hi;
        }
    };
    // interior output columns keep their 4-column halo inside the zero-sigma zone
    int ci0 = (z0 + 4 + 3) & ~3, ci1 = (z1 - 4) & ~3;
    if (ci1 - ci0 < 24) {
        add_cols(0, gp.nx, false);
    } else {
        add_cols(0, ci0, false);
        add_cols(ci0, ci1, true);
        add_cols(ci1, gp.nx, false);
    }
    for (auto &cc : cols) {
        // interior strips must have the whole 32-lane window inside the zero zone
        if (cc.interior && !(cc.x0 >= z0 && cc.x0 + 32 <= z1)) cc.interior = false;
    }

    // --- rows: output ranges in local rows ---
    struct Row { int j0, j1; bool interior; };
    std::vector<Row> rows;
    const int own0 = gp.ny_own0, own1 = gp.ny_own0 + gp.ny_own;
    // global rows [z0+4, z1-4) are interior candidates
    int ri0 = z0 + 4 - gp.grow0, ri1 = z1 - 4 - gp.grow0;
    if (ri0 < own0) ri0 = own0;
    if (ri1 > own1) ri1 = own1;
    const int SEG = 64;  // rows per slab
    auto add_rows = [&](int a, int b, bool interior) {
        if (b <= a) return;
        int n = (b - a + SEG - 1) / SEG;
        for (int k = 0; k < n; ++k) {
            int lo = a + (int)((long long)(b - a) * k / n), hi = a + (int)((long long)(b - a) * (k + 1) / n);
            rows.push_back({lo, hi, interior});
        }
    };
    if (ri1 - ri0 < 16) {
        add_rows(own0, own1, false);
    } else {
        add_rows(own0, ri0, false);
        add_rows(ri0, ri1, true);
        add_rows(ri1, own1, false);
    }

    std::vector<Item> ii, gg;
    for (auto &rr : rows)
        for (auto &cc : cols) {
            Item it;
            it.x0 = cc.x0;
            it.vlo = cc.olo - cc.x0;
            it.vhi = cc.ohi - cc.x0;
            it.j0 = rr.j0;
            it.j1 = rr.j1;
            it.la = rr.j0 - 4 < 0 ? 0 : rr.j0 - 4;
            it.lb = rr.j1 + 4 > gp.ny_alloc ? gp.ny_alloc : rr.j1 + 4;
            it.top = (gp.grow0 + it.la == 0) ? 1 : 0;
            it.bot = (gp.grow0 + it.lb == gp.ny_global) ? 1 : 0;
            bool interior = cc.interior && rr.interior && !it.top && !it.bot;
            // rows la..lb-1 must also be inside the zero zone
            if (interior && !(gp.grow0 + it.la >= z0 && gp.grow0 + it.lb <= z1)) interior = false;
            (interior ? ii : gg).push_back(it);
        }
    p->n_int = (int)ii.size();
    p->n_gen = (int)gg.size();
    if (p->d_items_int) cudaFree(p->d_items_int);
    if (p->d_items_gen) cudaFree(p->d_items_gen);
    cudaMalloc((void **)&p->d_items_int, sizeof(Item) * (ii.size() + 1));
    cudaMalloc((void **)&p->d_items_gen, sizeof(Item) * (gg.size() + 1));
    cudaMemcpy(p->d_items_int, ii.data(), sizeof(Item) * ii.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_items_gen, gg.data(), sizeof(Item) * gg.size(), cudaMemcpyHostToDevice);
    cudaMalloc((void **)&p->d_epart, sizeof(float) * 3 * (size_t)(ii.size() + gg.size()) * gp.n_env);
    cudaMalloc((void **)&p->d_bb, sizeof(int) * 4);

    constexpr int WARP_BYTES = RING * SLOT_F * 4 + BSET_F * 4 + CYL_CAP * 12 * 4 + RING * 8;
    p->smem_bytes = WARPS * ((WARP_BYTES + 127) & ~127);
    cudaError_t ce = cudaFuncSetAttribute(k_fused_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_fused_step<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_bytes);
    if (ce != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof(buf), "fused_prepare: cudaFuncSetAttribute: %s (was the library built for sm_100a?)", cudaGetErrorString(ce));
        return waves_set_error(buf);
    }
    h->fused_smem = p->smem_bytes;

    // TMA descriptors
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (ce != cudaSuccess || !fn) return waves_set_error("fused_prepare: cuTensorMapEncodeTiled entry point not found");
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    int r0 = make_map(enc, &h->map_u[0], h->u[0], gp, 12 * gp.n_env, 12);
    int r1 = make_map(enc, &h->map_u[1], h->u[1], gp, 12 * gp.n_env, 12);
    int r2 = make_map(enc, &h->map_shape, h->shape, gp, gp.n_env, 1);
    if (r0 || r1 || r2) {
        char buf[256];
        snprintf(buf, sizeof(buf), "fused_prepare: cuTensorMapEncodeTiled failed (%d %d %d)", r0, r1, r2);
        return waves_set_error(buf);
    }
    h->maps_ready = true;
    return 0;
}

void fused_release(waves_handle *h) {
    FusedPlan *p = static_cast<FusedPlan *>(h->plan);
    if (!p) return;
    if (p->d_items_int) cudaFree(p->d_items_int);
    if (p->d_items_gen) cudaFree(p->d_items_gen);
    if (p->d_epart) cudaFree(p->d_epart);
    if (p->d_bb) cudaFree(p->d_bb);
    delete p;
    h->plan = nullptr;
}

int fused_item_counts(waves_handle *h, int *n_int, int *n_gen) {
    FusedPlan *p = plan_of(h, false);
    if (!p) return 1;
    *n_int = p->n_int;
    *n_gen = p->n_gen;
    return 0;
}

int launch_fused_step(waves_handle *h, const float *d_table, int steps, int step, float *d_e3) {
    FusedPlan *p = plan_of(h, false);
    if (!p || !h->maps_ready) return waves_set_error("fused step: handle not prepared");
    for (int e = 0; e < h->gp.n_env; ++e)
        if (h->h_env[e].has_cplane)
            return waves_set_error("fused step: a fixed speed plane (waves_set_speed_field) needs WAVES_MODE_EXACT");
    FusedArgs A;
    A.gp = h->gp;
    A.env = h->d_env;
    A.cyl0 = h->d_cyl0;
    A.cyl1 = h->d_cyl1;
    A.cyl_cap = h->cyl_cap;
    A.table = d_table;
    A.steps = steps;
    A.step = step;
    A.out = h->u[h->cur ^ 1];
    A.epart = d_e3 ? p->d_epart : nullptr;
    A.epart_stride = p->n_int + p->n_gen;
    static const int dbg_flags = getenv("WAVES_DEBUG_FLAGS") ? atoi(getenv("WAVES_DEBUG_FLAGS")) : 0;
    A.dbg = dbg_flags;
    static const int dbg_skip = getenv("WAVES_DEBUG_SKIP") ? atoi(getenv("WAVES_DEBUG_SKIP")) : 0;  // developer bisecting aid
    if (h->profile) cudaEventRecord(h->ev0, h->stream);
    if (p->n_int > 0 && !(dbg_skip & 1)) {
        A.items = p->d_items_int;
        A.n_items = p->n_int;
        A.epart_off = 0;
        long long warps = (long long)p->n_int * h->gp.n_env;
        k_fused_step<false><<<(unsigned)((warps + WARPS - 1) / WARPS), WARPS * 32, p->smem_bytes, h->stream>>>(A, h->map_u[h->cur],
                                                                                                          h->map_shape);
        h->launches++;
    }
    if (p->n_gen > 0 && !(dbg_skip & 2)) {
        A.items = p->d_items_gen;
        A.n_items = p->n_gen;
        A.epart_off = p->n_int;
        long long warps = (long long)p->n_gen * h->gp.n_env;
        k_fused_step<true><<<(unsigned)((warps + WARPS - 1) / WARPS), WARPS * 32, p->smem_bytes, h->stream>>>(A, h->map_u[h->cur],
                                                                                                         h->map_shape);
        h->launches++;
    }
    if (h->profile) {
        cudaEventRecord(h->ev1, h->stream);
        cudaEventSynchronize(h->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        h->fused_ms += ms;
        h->fused_launches++;
    }
    if (d_e3) {
        k_energy_reduce<<<h->gp.n_env, 256, 0, h->stream>>>(p->d_epart, p->n_int + p->n_gen, h->d_omega, d_e3, 3 * (steps + 1));
        h->launches++;
    }
    h->cur ^= 1;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof(buf), "fused step launch: %s", cudaGetErrorString(ce));
        return waves_set_error(buf);
    }
    return 0;
}

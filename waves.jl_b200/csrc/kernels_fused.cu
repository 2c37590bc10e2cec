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

#include <algorithm>
#include <vector>

#include "waves_internal.h"

namespace {

constexpr int RING = 8;      // ring slots (rows) per warp
constexpr int PF = 2;        // TMA prefetch distance in rows
constexpr int CYL_CAP = 12;  // culled cylinders kept per warp
constexpr int WARPS = 1;     // warps per CTA: one, so every item-derived value is provably CTA-uniform

// Per-variant layout.  The interior variant marches BOTH wavefields in one warp (it needs both for
// the scattered energy and has registers to spare); the general variant (PML strips, domain borders)
// carries twice the state per wavefield, so each warp takes ONE wavefield.
// Ring slot rows (32 floats each): [0, NF) state fields written by TMA, row NF: source shape written by TMA
// (only for rows inside the source's bounding box).  On arrival the warp rewrites some rows IN PLACE
// (generic-proxy writes followed by fence.proxy.async before the slot's next TMA):
//   F_SHV (= the shape row): shape or 0;
//   interior: Psix rows -> P = Psix+Psiy-Omega (Psi/Omega are passed through to the output on arrival),
//             Psiy_tot, Omega_tot, Psiy_inc rows -> kd*c^2 at the three stage times;
//   general:  three extra rows hold kd*c^2 (all six fields stay live through the four stages).
template <bool GENERAL>
struct Cfg {
    static constexpr int NW = GENERAL ? 1 : 2;
    static constexpr int NF = 6 * NW;
    static constexpr int F_SH = NF;
    static constexpr int F_SHV = NF;
    static constexpr int SLOT_ROWS = GENERAL ? NF + 4 : NF + 1;
    static constexpr int SLOT_F = SLOT_ROWS * 32;
    static constexpr int BSET_F = GENERAL ? 20 * 32 : 0;  // border-row state + call scratch (general only)
    static constexpr int RING_F = RING * SLOT_F;
    static constexpr int CYL_OFF = RING_F + BSET_F;
    static constexpr int BAR_OFF = CYL_OFF + CYL_CAP * 12;
    static constexpr int WARP_F = ((BAR_OFF + RING * 2) + 31) & ~31;  // floats per warp (128-byte multiple)
    // row holding kd*c^2 at stage-time index tau
    __host__ __device__ static constexpr int f_bk(int tau) { return GENERAL ? NF + 1 + tau : (tau == 0 ? 4 : (tau == 1 ? 5 : 10)); }
    // interior: row holding P of wavefield w
    __host__ __device__ static constexpr int f_p(int w) { return w * 6 + 3; }
};

struct Item {
    int x0;        // first column of the 32-lane window (multiple of 4: TMA needs 16-byte aligned boxes)
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
    float *epart;      // [n_env][epart_stride][3] or nullptr
    int epart_stride;  // items per env in epart (all kernels of a step share one buffer)
    int epart_off;     // offset of this kernel's items
    int dbg;           // developer bisecting flags (WAVES_DEBUG_FLAGS)
    // host-computed step constants: read straight from the constant bank as FFMA operands
    float kd, b0kd;             // 1/(2Δ) and c0^2/(2Δ)
    float akd_h, akd_f, dt6kd;  // (dt/2)kd, dt*kd, (dt/6)kd
    float hdt, dt, dt6;
    unsigned plane, nxp;
};

extern __shared__ __align__(128) float smf[];  // all shared memory, indexed with 32-bit arithmetic

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

// Per-warp context.  Everything here is either warp-uniform or a per-lane constant of the whole march.
struct WarpCtx {
    int lane, wb;      // lane, first float of this warp's shared memory
    int x0, la, lb;
    int js0;           // first row the regular stages may store
    unsigned jsn;      // rows [js0, js0+jsn) are stored by this lane (0 for halo lanes)
    float *out_e;      // output state of this env at this lane's column (+ wavefield offset in the general variant)
    bool xb, top, bot, is_first_col, is_last_col, is_tot;
    int w0;
    float xs, sx, bcm;
    float sf[3];       // source factor at t, t+dt/2, t+dt
    uint32_t bar0, ring_sa;  // shared-window addresses of mbarrier 0 and of the ring
    int nact;          // culled cylinders (0: none touch this window, -1: list overflow -> slow loop)
    int src_j0;
    unsigned src_n;    // rows [src_j0, src_j0+src_n) of this window carry a non-zero source shape
};

__device__ __forceinline__ bool src_row(const WarpCtx &c, int j) { return (unsigned)(j - c.src_j0) < c.src_n; }

// c(x,y,t)^2 with every cylinder (list overflow), exact order of src/designs.jl:99-116
__device__ __noinline__ float speed2_slow(const FusedArgs &A, int e, int tau, float xs, float yv) {
    const EnvParams ep = A.env[e];
    const float *cyl0 = A.cyl0 + (size_t)e * A.cyl_cap * 4, *cyl1 = A.cyl1 + (size_t)e * A.cyl_cap * 4;
    const int ncyl = ep.ncyl;
    const float ti = ep.ti, tf = ep.tf, c0 = A.gp.c0;
    const float t = A.table[((size_t)e * A.steps + A.step) * STAGE_ROW + tau];
    int cnt = 0;
    float cd = 0.0f;
    for (int k = 0; k < ncyl; ++k) {
        float px = interp_param(cyl0[4 * k], cyl1[4 * k], ti, tf, t);
        float py = interp_param(cyl0[4 * k + 1], cyl1[4 * k + 1], ti, tf, t);
        float r = interp_param(cyl0[4 * k + 2], cyl1[4 * k + 2], ti, tf, t);
        float cc = interp_param(cyl0[4 * k + 3], cyl1[4 * k + 3], ti, tf, t);
        float dx = __fsub_rn(xs, px), dy = __fsub_rn(yv, py);
        float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        bool m = d2 < __fmul_rn(r, r);
        cnt += m;
        cd = __fadd_rn(cd, m ? cc : 0.0f);
    }
    float cv = __fadd_rn(cnt == 0 ? c0 : 0.0f, cd);
    return __fmul_rn(cv, cv);
}

// kd * speed(design(t), grid, c0)^2 of local row j at the three stage times -> smf[dst + tau*32]
// (src/designs.jl:99-116: strict '<', speeds of overlapping cylinders add, ambient where none).
template <bool GENERAL>
__device__ __forceinline__ void speed_row(const WarpCtx &c, const FusedArgs &A, int e, int j, int uri) {
    using C = Cfg<GENERAL>;
    const float yv = A.gp.y[min(max(A.gp.grow0 + j, 0), A.gp.ny_global - 1)];
#pragma unroll 1
    for (int tau = 0; tau < 3; ++tau) {
        float b;
        if (c.nact > 0) {
            int cnt = 0;
            float cd = 0.0f;
#pragma unroll 1
            for (int a = 0; a < c.nact; ++a) {
                const float4 p = *reinterpret_cast<const float4 *>(&smf[c.wb + C::CYL_OFF + (a * 3 + tau) * 4]);  // px, py, r^2, c
                float dy = __fsub_rn(yv, p.y);
                float dy2 = __fmul_rn(dy, dy);
                if (dy2 >= p.z) continue;  // warp-uniform: the row misses this cylinder
                float dx = __fsub_rn(c.xs, p.x);
                float d2 = __fadd_rn(__fmul_rn(dx, dx), dy2);
                bool m = d2 < p.z;
                cnt += m;
                cd = __fadd_rn(cd, m ? p.w : 0.0f);
            }
            float cv = __fadd_rn(cnt == 0 ? A.gp.c0 : 0.0f, cd);
            b = __fmul_rn(cv, cv);
        } else {
            b = speed2_slow(A, e, tau, c.xs, yv);
        }
        smf[uri + (tau == 0 ? C::f_bk(0) : (tau == 1 ? C::f_bk(1) : C::f_bk(2))) * 32] = b * A.kd;
    }
}

// Out-of-line copy for the general variant (keeps its instruction footprint inside the I-cache)
__device__ __noinline__ void speed_row_general(const FusedArgs &A, int e, int it, int j, int uri, int nact) {
    using C = Cfg<true>;
    const int lane = threadIdx.x & 31, wb = (threadIdx.x >> 5) * C::WARP_F;
    const int col = A.items[it].x0 + lane;
    const float xs = A.gp.x[min(col, A.gp.nx - 1)];
    const float yv = A.gp.y[min(max(A.gp.grow0 + j, 0), A.gp.ny_global - 1)];
    for (int tau = 0; tau < 3; ++tau) {
        float b;
        if (nact > 0) {
            int cnt = 0;
            float cd = 0.0f;
            for (int a = 0; a < nact; ++a) {
                const float4 p = *reinterpret_cast<const float4 *>(&smf[wb + C::CYL_OFF + (a * 3 + tau) * 4]);
                float dy = __fsub_rn(yv, p.y);
                float dy2 = __fmul_rn(dy, dy);
                if (dy2 >= p.z) continue;
                float dx = __fsub_rn(xs, p.x);
                float d2 = __fadd_rn(__fmul_rn(dx, dx), dy2);
                bool m = d2 < p.z;
                cnt += m;
                cd = __fadd_rn(cd, m ? p.w : 0.0f);
            }
            float cv = __fadd_rn(cnt == 0 ? A.gp.c0 : 0.0f, cd);
            b = __fmul_rn(cv, cv);
        } else {
            b = speed2_slow(A, e, tau, xs, yv);
        }
        smf[uri + C::f_bk(tau) * 32] = b * A.kd;
    }
}

// Register state of one warp: rotating windows indexed [stage][wavefield][row & 3]
template <bool GENERAL>
struct Regs {
    static constexpr int NW = Cfg<GENERAL>::NW;
    float Uf[4][NW][4];  // U + f of stage state y_s (s = 0: the loaded row)
    float Vy[4][NW][4];
    float Vx[4][NW][4];  // s = 1..3
    float aU[NW][4], aVx[NW][4], aVy[NW][4];  // k1 + 2k2 + 2k3 accumulators (aVx, aVy: un-scaled differences)
    // GENERAL only
    float Uc[4][NW][4], Px[4][NW][4], Py[4][NW][4], Om[4][NW][4];
    float aPx[NW][4], aPy[NW][4], aOm[NW][4];
    float e_tot, e_inc, e_sc;
};

// Border-row state kept in shared memory (general variant, per warp): row index * 32 + lane
enum { B_VX = 0, B_UC, B_PX, B_PY, B_OM, B_AU, B_AVX, B_AVY, B_APX, B_APY, B_AOM, B_F0 = 12, B_F1, B_F2, B_V0, B_V1, B_V2 };

// One RK stage on a domain-border row (global row 0 or ny-1) of this warp's wavefield.  The caller left
// U+f and Vy of the previous stage state on the three rows the one-sided stencil spans (ascending rows)
// in the border scratch B_F0..B_V2; everything else is re-derived here (rare path, kept out of line and
// out of the hot path's registers).  Returns the new (Uf, Vy) of the border row.
__device__ __noinline__ float2 border_row_stage(const FusedArgs &A, int e, int it, int w0, int S, bool top, int nact) {
    using C = Cfg<true>;
    const GridP &gp = A.gp;
    const int lane = threadIdx.x & 31, wb = (threadIdx.x >> 5) * C::WARP_F;
    const Item item = A.items[it];
    const int jb = top ? item.la : item.lb - 1;
    const int col = item.x0 + lane;
    const int uri = wb + ((jb - item.la) & (RING - 1)) * C::SLOT_F + lane;
    const int bsi = wb + C::RING_F + lane;
    const float *g = top ? gp.g_first : gp.g_last;
    const float kd = gp.g_central[1];
    const float sx = gp.sigma[min(col, gp.nx - 1)], sy = gp.sigma[gp.grow0 + jb];
    const float a = (S == 3) ? gp.dt : gp.hdt;
    const int tau = (S == 1) ? 0 : (S == 4 ? 2 : 1);
    const float *trow = A.table + ((size_t)e * A.steps + A.step) * STAGE_ROW;
    const float sf_next = (S == 1 || S == 2) ? trow[4] : trow[5];
    const float shv = smf[uri + C::F_SHV * 32];
    const float b = (w0 == 0 && nact != 0) ? smf[uri + C::f_bk(tau) * 32] / kd : gp.b0;
    const float uU = smf[uri], uVx = smf[uri + 32], uVy = smf[uri + 64], uPx = smf[uri + 96], uPy = smf[uri + 128], uOm = smf[uri + 160];
    const float f0 = smf[bsi + B_F0 * 32], f1 = smf[bsi + B_F1 * 32], f2 = smf[bsi + B_F2 * 32];
    const float v0 = smf[bsi + B_V0 * 32], v1 = smf[bsi + B_V1 * 32], v2 = smf[bsi + B_V2 * 32];
    const float ufC = top ? f0 : f2, vyC = top ? v0 : v2;
    const float vxC = (S == 1) ? uVx : smf[bsi + B_VX * 32];
    const float uC = (S == 1) ? uU : smf[bsi + B_UC * 32];
    auto ddx = [&](float v) {
        float e1 = __shfl_down_sync(0xffffffffu, v, 1), w1 = __shfl_up_sync(0xffffffffu, v, 1);
        float e2 = __shfl_down_sync(0xffffffffu, v, 2), w2 = __shfl_up_sync(0xffffffffu, v, 2);
        float d = kd * (e1 - w1);
        if (col == 0) d = ((gp.g_first[0] * v) + (gp.g_first[1] * e1)) + (gp.g_first[2] * e2);
        if (col == gp.nx - 1) d = ((gp.g_last[0] * w2) + (gp.g_last[1] * w1)) + (gp.g_last[2] * v);
        return d;
    };
    const float Ux = ddx(ufC), Vxx = ddx(vxC);
    const float Uy = ((g[0] * f0) + (g[1] * f1)) + (g[2] * f2);
    const float Vyy = ((g[0] * v0) + (g[1] * v1)) + (g[2] * v2);
    // bc == 0 on a border row (src/dims.jl:117-124): dU = 0
    const float kU = 0.0f;
    const float kVx = Ux - sx * vxC, kVy = Uy - sy * vyC;
    const float kPx = (b * sx) * Vyy, kPy = (b * sy) * Vxx, kOm = (sx * sy) * uC;
    float2 ret = make_float2(0.f, 0.f);
    if (S < 4) {
        const float Us = uU + a * kU;
        ret.x = Us + shv * sf_next;
        ret.y = uVy + a * kVy;
        smf[bsi + B_VX * 32] = uVx + a * kVx;
        smf[bsi + B_UC * 32] = Us;
        smf[bsi + B_PX * 32] = uPx + a * kPx;
        smf[bsi + B_PY * 32] = uPy + a * kPy;
        smf[bsi + B_OM * 32] = uOm + a * kOm;
        const float m = (S == 1) ? 0.0f : 1.0f, w = (S == 1) ? 1.0f : 2.0f;
        smf[bsi + B_AU * 32] = m * smf[bsi + B_AU * 32] + w * kU;
        smf[bsi + B_AVX * 32] = m * smf[bsi + B_AVX * 32] + w * kVx;
        smf[bsi + B_AVY * 32] = m * smf[bsi + B_AVY * 32] + w * kVy;
        smf[bsi + B_APX * 32] = m * smf[bsi + B_APX * 32] + w * kPx;
        smf[bsi + B_APY * 32] = m * smf[bsi + B_APY * 32] + w * kPy;
        smf[bsi + B_AOM * 32] = m * smf[bsi + B_AOM * 32] + w * kOm;
    } else if (lane >= item.vlo && lane < item.vhi && jb >= item.j0 && jb < item.j1) {
        const float sixth = 1.0f / 6.0f;
        float *o = A.out + (long long)e * gp.env_stride + (long long)w0 * 6 * gp.plane + (long long)jb * gp.nxp + col;
        o[0] = uU + (sixth * (smf[bsi + B_AU * 32] + kU)) * gp.dt;
        o[gp.plane] = uVx + (sixth * (smf[bsi + B_AVX * 32] + kVx)) * gp.dt;
        o[2 * gp.plane] = uVy + (sixth * (smf[bsi + B_AVY * 32] + kVy)) * gp.dt;
        o[3 * gp.plane] = uPx + (sixth * (smf[bsi + B_APX * 32] + kPx)) * gp.dt;
        o[4 * gp.plane] = uPy + (sixth * (smf[bsi + B_APY * 32] + kPy)) * gp.dt;
        o[5 * gp.plane] = uOm + (sixth * (smf[bsi + B_AOM * 32] + kOm)) * gp.dt;
    }
    __syncwarp();
    return ret;
}

// One RK stage S (1..4) on row j = r - S.  PH = r & 3.  Rows whose inputs are not loaded yet (warm-up)
// produce values that no stored cell depends on, and stores are predicated.
// Derivatives are kept un-scaled (differences); the 1/(2Δ) factor is folded into the coefficients.
template <bool GENERAL, int S, int PH>
__device__ __forceinline__ void stage(const WarpCtx &c, const FusedArgs &A, int e, int it, Regs<GENERAL> &R, int j) {
    using C = Cfg<GENERAL>;
    constexpr int sc = (PH - S + 8) & 3, sm = (sc + 3) & 3, sp = (sc + 1) & 3;  // slots of rows j, j-1, j+1
    const int uri = c.wb + ((j - c.la) & (RING - 1)) * C::SLOT_F + c.lane;
    const float a = (S == 3) ? A.dt : A.hdt;
    const float akd = (S == 3) ? A.akd_f : A.akd_h;
    constexpr int tau = (S == 1) ? 0 : (S == 4 ? 2 : 1);
    const float shv = (S < 4) ? smf[uri + C::F_SHV * 32] : 0.0f;
    const float sf_next = (S == 1 || S == 2) ? c.sf[1] : c.sf[2];
    const float bk_tot = smf[uri + C::f_bk(tau) * 32];  // kd * c^2 (written on arrival)
    const bool st = (unsigned)(j - c.js0) < c.jsn;         // this lane stores row j
    float outU[C::NW];
    if (!GENERAL) {
#pragma unroll
        for (int w = 0; w < C::NW; ++w) {
            const int uw = uri + w * 6 * 32;
            const float uU = smf[uw], uVx = smf[uw + 32], uVy = smf[uw + 64];
            const float bk = (w == 0) ? bk_tot : A.b0kd;
            const float ufC = R.Uf[S - 1][w][sc];
            const float vxC = (S == 1) ? uVx : R.Vx[S - 1][w][sc];
            const float dUx = __shfl_down_sync(0xffffffffu, ufC, 1) - __shfl_up_sync(0xffffffffu, ufC, 1);
            const float dVx = __shfl_down_sync(0xffffffffu, vxC, 1) - __shfl_up_sync(0xffffffffu, vxC, 1);
            const float dUy = R.Uf[S - 1][w][sp] - R.Uf[S - 1][w][sm];
            const float dVy = R.Vy[S - 1][w][sp] - R.Vy[S - 1][w][sm];
            // sigma == 0 in the whole window: Psi, Omega never change within the step; P was formed on arrival
            const float kU = bk * (dVx + dVy) + smf[uri + C::f_p(w) * 32];
            if (S < 4) {
                const float Us = uU + a * kU;
                R.Uf[S][w][sc] = Us + shv * sf_next;
                R.Vx[S][w][sc] = uVx + akd * dUx;
                R.Vy[S][w][sc] = uVy + akd * dUy;
                if (S == 1) {
                    R.aU[w][sc] = kU;
                    R.aVx[w][sc] = dUx;
                    R.aVy[w][sc] = dUy;
                } else {
                    R.aU[w][sc] += 2.0f * kU;
                    R.aVx[w][sc] += 2.0f * dUx;
                    R.aVy[w][sc] += 2.0f * dUy;
                }
            } else {
                const float oU = uU + A.dt6 * (R.aU[w][sc] + kU);
                outU[w] = oU;
                if (st) {
                    float *o = c.out_e + (unsigned)(w * 6) * A.plane + (unsigned)j * A.nxp;
                    o[0] = oU;
                    o[A.plane] = uVx + A.dt6kd * (R.aVx[w][sc] + dUx);
                    o[2u * A.plane] = uVy + A.dt6kd * (R.aVy[w][sc] + dUy);
                }
            }
        }
        if (S == 4 && st) {
            const float d = outU[0] - outU[C::NW - 1];
            R.e_tot += outU[0] * outU[0];
            R.e_inc += outU[C::NW - 1] * outU[C::NW - 1];
            R.e_sc += d * d;
        }
    } else {
        const GridP &gp = A.gp;
        const float sy = gp.sigma[min(max(A.gp.grow0 + j, 0), A.gp.ny_global - 1)];
        const float sx = c.sx;
        const float uU = smf[uri], uVx = smf[uri + 32], uVy = smf[uri + 64], uPx = smf[uri + 96], uPy = smf[uri + 128], uOm = smf[uri + 160];
        const float bk = c.is_tot ? bk_tot : A.b0kd;
        const float ufC = R.Uf[S - 1][0][sc];
        const float vxC = (S == 1) ? uVx : R.Vx[S - 1][0][sc];
        const float vyC = R.Vy[S - 1][0][sc];
        const float uC = (S == 1) ? uU : R.Uc[S - 1][0][sc];
        const float px = (S == 1) ? uPx : R.Px[S - 1][0][sc];
        const float py = (S == 1) ? uPy : R.Py[S - 1][0][sc];
        const float om = (S == 1) ? uOm : R.Om[S - 1][0][sc];
        // un-scaled x-differences (one-sided 3-point rows on the domain's edge columns, pre-divided by kd)
        float e1 = __shfl_down_sync(0xffffffffu, ufC, 1), w1 = __shfl_up_sync(0xffffffffu, ufC, 1);
        float dUx = e1 - w1;
        float e1v = __shfl_down_sync(0xffffffffu, vxC, 1), w1v = __shfl_up_sync(0xffffffffu, vxC, 1);
        float dVx = e1v - w1v;
        if (c.xb) {
            // one-sided 3-point rows on the domain's edge columns (src/operators.jl:3-4); relative to 1/(2Δ) the
            // coefficients are (-3, 4, -1) and (1, -4, 3)
            const float e2 = __shfl_down_sync(0xffffffffu, ufC, 2), w2 = __shfl_up_sync(0xffffffffu, ufC, 2);
            const float e2v = __shfl_down_sync(0xffffffffu, vxC, 2), w2v = __shfl_up_sync(0xffffffffu, vxC, 2);
            if (c.is_first_col) {
                dUx = (4.0f * e1 - 3.0f * ufC) - e2;
                dVx = (4.0f * e1v - 3.0f * vxC) - e2v;
            }
            if (c.is_last_col) {
                dUx = (3.0f * ufC - 4.0f * w1) + w2;
                dVx = (3.0f * vxC - 4.0f * w1v) + w2v;
            }
        }
        const float dUy = R.Uf[S - 1][0][sp] - R.Uf[S - 1][0][sm];
        const float dVy = R.Vy[S - 1][0][sp] - R.Vy[S - 1][0][sm];
        const float kU = c.bcm * ((((bk * (dVx + dVy) + px) + py) - (sx + sy) * uC) - om);
        const float kVx = A.kd * dUx - sx * vxC;
        const float kVy = A.kd * dUy - sy * vyC;
        const float kPx = (bk * sx) * dVy;
        const float kPy = (bk * sy) * dVx;
        const float kOm = (sx * sy) * uC;
        if (S < 4) {
            const float Us = uU + a * kU;
            R.Uf[S][0][sc] = Us + shv * sf_next;
            R.Vx[S][0][sc] = uVx + a * kVx;
            R.Vy[S][0][sc] = uVy + a * kVy;
            R.Uc[S][0][sc] = Us;
            R.Px[S][0][sc] = uPx + a * kPx;
            R.Py[S][0][sc] = uPy + a * kPy;
            R.Om[S][0][sc] = uOm + a * kOm;
            if (S == 1) {
                R.aU[0][sc] = kU;
                R.aVx[0][sc] = kVx;
                R.aVy[0][sc] = kVy;
                R.aPx[0][sc] = kPx;
                R.aPy[0][sc] = kPy;
                R.aOm[0][sc] = kOm;
            } else {
                R.aU[0][sc] += 2.0f * kU;
                R.aVx[0][sc] += 2.0f * kVx;
                R.aVy[0][sc] += 2.0f * kVy;
                R.aPx[0][sc] += 2.0f * kPx;
                R.aPy[0][sc] += 2.0f * kPy;
                R.aOm[0][sc] += 2.0f * kOm;
            }
        } else if (st) {
            float *o = c.out_e + (unsigned)j * A.nxp;
            o[0] = uU + A.dt6 * (R.aU[0][sc] + kU);
            o[A.plane] = uVx + A.dt6 * (R.aVx[0][sc] + kVx);
            o[2u * A.plane] = uVy + A.dt6 * (R.aVy[0][sc] + kVy);
            o[3u * A.plane] = uPx + A.dt6 * (R.aPx[0][sc] + kPx);
            o[4u * A.plane] = uPy + A.dt6 * (R.aPy[0][sc] + kPy);
            o[5u * A.plane] = uOm + A.dt6 * (R.aOm[0][sc] + kOm);
        }
        // domain-border rows ride along with their inward neighbour (see file header)
        const bool do_top = c.top && j == c.la + 1, do_bot = c.bot && j == c.lb - 2;
        if (do_top || do_bot) {
            const int bsi = c.wb + C::RING_F + c.lane;
            smf[bsi + B_F0 * 32] = R.Uf[S - 1][0][sm];
            smf[bsi + B_F1 * 32] = R.Uf[S - 1][0][sc];
            smf[bsi + B_F2 * 32] = R.Uf[S - 1][0][sp];
            smf[bsi + B_V0 * 32] = R.Vy[S - 1][0][sm];
            smf[bsi + B_V1 * 32] = R.Vy[S - 1][0][sc];
            smf[bsi + B_V2 * 32] = R.Vy[S - 1][0][sp];
            if (do_top) {
                const float2 o = border_row_stage(A, e, it, c.w0, S, true, c.nact);
                if (S < 4) {
                    R.Uf[S][0][sm] = o.x;
                    R.Vy[S][0][sm] = o.y;
                }
            }
            if (do_bot) {
                const float2 o = border_row_stage(A, e, it, c.w0, S, false, c.nact);
                if (S < 4) {
                    R.Uf[S][0][sp] = o.x;
                    R.Vy[S][0][sp] = o.y;
                }
            }
        }
    }
}

template <bool GENERAL, int PH>
__device__ __forceinline__ void row_step(const WarpCtx &c, const FusedArgs &A, int e, int it, Regs<GENERAL> &R, int r,
                                         const CUtensorMap *map_u, const CUtensorMap *map_sh) {
    using C = Cfg<GENERAL>;
    // 1. prefetch row r + PF into the slot that row r + PF - RING vacated (all its readers finished >= 1 step ago)
    __syncwarp();
    const int rp = r + PF;
    if (c.lane == 0 && rp >= c.la && rp < c.lb && !(A.dbg & 2)) {
        const int slot = (rp - c.la) & (RING - 1);
        const uint32_t bar = c.bar0 + slot * 8, dst = c.ring_sa + slot * (C::SLOT_F * 4);
        const bool s = src_row(c, rp);
        mbar_expect_tx(bar, C::NF * 128 + (s ? 128 : 0));
        tma_load_3d(dst, map_u, c.x0, rp, e * 12 + c.w0 * 6, bar);
        if (s) tma_load_3d(dst + C::F_SH * 128, map_sh, c.x0, rp, e, bar);
    }
    // 2. arrival of row r: stage-0 windows, speed field of the row, interior pass-through of Psi/Omega
    if (r >= c.la && r < c.lb) {
        const int rel = r - c.la;
        if (!(A.dbg & 2)) mbar_wait(c.bar0 + (rel & (RING - 1)) * 8, (rel >> 3) & 1);
        const int uri = c.wb + (rel & (RING - 1)) * C::SLOT_F + c.lane;
        const float shv = src_row(c, r) ? smf[uri + C::F_SH * 32] : 0.0f;
        smf[uri + C::F_SHV * 32] = shv;
        constexpr int s0 = PH & 3;
#pragma unroll
        for (int w = 0; w < C::NW; ++w) {
            R.Uf[0][w][s0] = smf[uri + w * 6 * 32] + shv * c.sf[0];
            R.Vy[0][w][s0] = smf[uri + w * 6 * 32 + 64];
        }
        float pP[C::NW];
        if (!GENERAL) {
            const bool st = (unsigned)(r - c.js0) < c.jsn;
#pragma unroll
            for (int w = 0; w < C::NW; ++w) {
                const float px = smf[uri + (w * 6 + 3) * 32], py = smf[uri + (w * 6 + 4) * 32], om = smf[uri + (w * 6 + 5) * 32];
                pP[w] = (px + py) - om;
                if (st) {
                    float *o = c.out_e + (unsigned)(w * 6 + 3) * A.plane + (unsigned)r * A.nxp;
                    o[0] = px;
                    o[A.plane] = py;
                    o[2u * A.plane] = om;
                }
            }
#pragma unroll
            for (int w = 0; w < C::NW; ++w) smf[uri + C::f_p(w) * 32] = pP[w];
        }
        if (c.nact != 0) {
            if (GENERAL)
                speed_row_general(A, e, it, r, uri, c.nact);
            else
                speed_row<GENERAL>(c, A, e, r, uri);
        } else {
            smf[uri + C::f_bk(0) * 32] = A.b0kd;
            smf[uri + C::f_bk(1) * 32] = A.b0kd;
            smf[uri + C::f_bk(2) * 32] = A.b0kd;
        }
        // the rows rewritten above are TMA targets again RING rows later: order the generic writes before it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (A.dbg & 1) return;
    // 3. the four stages, each one row behind the previous.  Rows below la + 1 run unguarded (whatever they
    //    compute is overwritten before a stored cell reads it).  In the general variant rows above lb - 2
    //    must not run: they would clobber the border row's window slots.
    if (GENERAL) {
        const int jhi = c.lb - 2;
        if (r - 1 <= jhi) stage<GENERAL, 1, PH>(c, A, e, it, R, r - 1);
        if (r - 2 <= jhi) stage<GENERAL, 2, PH>(c, A, e, it, R, r - 2);
        if (r - 3 <= jhi) stage<GENERAL, 3, PH>(c, A, e, it, R, r - 3);
        if (r - 4 <= jhi) stage<GENERAL, 4, PH>(c, A, e, it, R, r - 4);
    } else {
        stage<GENERAL, 1, PH>(c, A, e, it, R, r - 1);
        stage<GENERAL, 2, PH>(c, A, e, it, R, r - 2);
        stage<GENERAL, 3, PH>(c, A, e, it, R, r - 3);
        stage<GENERAL, 4, PH>(c, A, e, it, R, r - 4);
    }
}

template <bool GENERAL>
__global__ void __launch_bounds__(WARPS * 32, 12 / WARPS)
k_fused_step(const __grid_constant__ FusedArgs A, const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_sh) {
    using C = Cfg<GENERAL>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long gw = (long long)blockIdx.x * WARPS + warp;
    const int w0 = GENERAL ? (int)(gw & 1) : 0;  // general: two warps (tot, inc) per item
    if (GENERAL) gw >>= 1;
    const int e = (int)(gw / A.n_items), it = (int)(gw - (long long)e * A.n_items);
    if (e >= A.gp.n_env) return;
    const GridP &gp = A.gp;
    const Item item = A.items[it];
    const EnvParams ep = A.env[e];

    WarpCtx c;
    c.lane = lane;
    c.wb = warp * C::WARP_F;
    c.ring_sa = smem_u32(&smf[c.wb]);
    c.bar0 = smem_u32(&smf[c.wb + C::BAR_OFF]);
    c.w0 = w0;
    c.is_tot = w0 == 0;
    c.x0 = item.x0;
    c.la = item.la;
    c.lb = item.lb;
    c.top = item.top;
    c.bot = item.bot;
    const int col = item.x0 + lane;
    const bool valid_lane = lane >= item.vlo && lane < item.vhi;
    {
        // regular stages store rows that have both y-neighbours loaded; border rows belong to border_row_stage
        const int js0 = max(item.j0, item.la + 1), js1 = min(item.j1, item.lb - 1);
        c.js0 = js0;
        c.jsn = (valid_lane && js1 > js0) ? (unsigned)(js1 - js0) : 0u;
    }
    c.out_e = A.out + (long long)e * gp.env_stride + (long long)w0 * 6 * gp.plane + min(col, gp.nx - 1);
    c.is_first_col = col == 0;
    c.is_last_col = col == gp.nx - 1;
    c.xb = (item.x0 == 0) || (item.x0 + 32 >= gp.nx);
    c.bcm = (c.is_first_col || c.is_last_col) ? 0.0f : 1.0f;
    c.xs = gp.x[min(col, gp.nx - 1)];  // lanes past the last column (nx % 4 != 0) are never valid
    c.sx = GENERAL ? gp.sigma[min(col, gp.nx - 1)] : 0.0f;
    const float *trow = A.table + ((size_t)e * A.steps + A.step) * STAGE_ROW;
    c.sf[0] = trow[3];
    c.sf[1] = trow[4];
    c.sf[2] = trow[5];
    {
        const bool src_cols = ep.has_source && item.x0 < ep.src_i1 && item.x0 + 32 > ep.src_i0;
        c.src_j0 = ep.src_j0;
        c.src_n = (src_cols && ep.src_j1 > ep.src_j0) ? (unsigned)(ep.src_j1 - ep.src_j0) : 0u;
    }

    if (lane == 0 && !(A.dbg & 8)) {
        for (int s = 0; s < RING; ++s) mbar_init(c.bar0 + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // cull the design's cylinders against this warp's window, at the three stage times (src/designs.jl:287-292)
    c.nact = 0;
    if (ep.ncyl > 0 && (!GENERAL || c.is_tot) && !(A.dbg & 4)) {
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
                    *reinterpret_cast<float4 *>(&smf[c.wb + C::CYL_OFF + (pos * 3 + tau) * 4]) =
                        make_float4(P[tau][0], P[tau][1], P[tau][2], P[tau][3]);
            }
            n += __popc(bal);
        }
        c.nact = n <= CYL_CAP ? n : -1;
        __syncwarp();
    }

    Regs<GENERAL> R;
    R.e_tot = R.e_inc = R.e_sc = 0.0f;
    // zero-init windows so never-consumed warm-up values are finite
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int w = 0; w < C::NW; ++w)
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
    for (int w = 0; w < C::NW; ++w)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            R.aU[w][q] = R.aVx[w][q] = R.aVy[w][q] = 0.f;
            R.aPx[w][q] = R.aPy[w][q] = R.aOm[w][q] = 0.f;
        }

    const int r_begin = (c.la - PF) & ~3, r_end = c.lb + 4;  // the first PF steps only prefetch
#pragma unroll 1
    for (int r = r_begin; r < r_end; r += 4) {
        row_step<GENERAL, 0>(c, A, e, it, R, r, &map_u, &map_sh);
        row_step<GENERAL, 1>(c, A, e, it, R, r + 1, &map_u, &map_sh);
        row_step<GENERAL, 2>(c, A, e, it, R, r + 2, &map_u, &map_sh);
        row_step<GENERAL, 3>(c, A, e, it, R, r + 3, &map_u, &map_sh);
    }

    if (!GENERAL && A.epart) {
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

// Energy partials of the output cells owned by the general items (PML strips / borders): one warp per item,
// reading the freshly written U_tot / U_inc rows (src/env.jl:104-111).
__global__ void __launch_bounds__(128) k_energy_items(GridP gp, const float *__restrict__ u, const Item *__restrict__ items, int n_items,
                                                      float *__restrict__ epart, int epart_stride, int epart_off) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int e = (int)(gw / n_items), it = (int)(gw - (long long)e * n_items);
    if (e >= gp.n_env) return;
    const Item item = items[it];
    const float *ut = u + (long long)e * gp.env_stride, *ui = ut + 6 * gp.plane;
    float et = 0.f, ei = 0.f, es = 0.f;
    if (lane >= item.vlo && lane < item.vhi) {
        for (int j = item.j0; j < item.j1; ++j) {
            const long long q = (long long)j * gp.nxp + item.x0 + lane;
            const float a = ut[q], b = ui[q], d = a - b;
            et += a * a;
            ei += b * b;
            es += d * d;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        et += __shfl_down_sync(0xffffffffu, et, o);
        ei += __shfl_down_sync(0xffffffffu, ei, o);
        es += __shfl_down_sync(0xffffffffu, es, o);
    }
    if (lane == 0) {
        float *p = epart + ((size_t)e * epart_stride + epart_off + it) * 3;
        p[0] = et;
        p[1] = ei;
        p[2] = es;
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
    int smem_int = 0, smem_gen = 0;
};

FusedPlan *plan_of(waves_handle *h, bool create) {
    if (!h->plan && create) h->plan = new FusedPlan();
    return static_cast<FusedPlan *>(h->plan);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(PFN_encodeTiled enc, CUtensorMap *m, float *base, const GridP &gp, int planes, int box_planes) {
    // box = 32 columns x 1 row x box_planes field planes
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
            if (x0 + 32 >= gp.nx) hi = gp.nx;  // the window reaches the right domain edge: no halo needed there
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
    // rows per slab: tall slabs amortise the 8 warm-up rows, but keep >= ~8 waves of warps in flight
    const int SEG = std::max(48, std::min(192, (int)((long long)gp.ny_own * (long long)cols.size() * gp.n_env / 14000)));
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

    p->smem_gen = WARPS * Cfg<true>::WARP_F * 4;
    p->smem_int = WARPS * Cfg<false>::WARP_F * 4;
    cudaError_t ce = cudaFuncSetAttribute(k_fused_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_gen);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_fused_step<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem_int);
    if (ce != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof(buf), "fused_prepare: cudaFuncSetAttribute: %s (was the library built for sm_100a?)", cudaGetErrorString(ce));
        return waves_set_error(buf);
    }
    h->fused_smem = p->smem_int;

    // TMA descriptors
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (ce != cudaSuccess || !fn) return waves_set_error("fused_prepare: cuTensorMapEncodeTiled entry point not found");
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    int r0 = make_map(enc, &h->map_u[0], h->u[0], gp, 12 * gp.n_env, 12);
    int r1 = make_map(enc, &h->map_u[1], h->u[1], gp, 12 * gp.n_env, 12);
    int r2 = make_map(enc, &h->map_shape, h->shape, gp, gp.n_env, 1);
    r0 |= make_map(enc, &h->map_u6[0], h->u[0], gp, 12 * gp.n_env, 6);
    r1 |= make_map(enc, &h->map_u6[1], h->u[1], gp, 12 * gp.n_env, 6);
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
    A.kd = h->gp.g_central[1];
    A.b0kd = h->gp.b0 * A.kd;
    A.dt = h->gp.dt;
    A.hdt = h->gp.hdt;
    A.dt6 = h->gp.dt * (1.0f / 6.0f);
    A.akd_h = A.hdt * A.kd;
    A.akd_f = A.dt * A.kd;
    A.dt6kd = A.dt6 * A.kd;
    A.plane = (unsigned)h->gp.plane;
    A.nxp = (unsigned)h->gp.nxp;
    static const int dbg_flags = getenv("WAVES_DEBUG_FLAGS") ? atoi(getenv("WAVES_DEBUG_FLAGS")) : 0;
    A.dbg = dbg_flags;
    static const int dbg_skip = getenv("WAVES_DEBUG_SKIP") ? atoi(getenv("WAVES_DEBUG_SKIP")) : 0;  // developer bisecting aid
    if (h->profile) cudaEventRecord(h->ev0, h->stream);
    if (p->n_int > 0 && !(dbg_skip & 1)) {
        A.items = p->d_items_int;
        A.n_items = p->n_int;
        A.epart_off = 0;
        long long warps = (long long)p->n_int * h->gp.n_env;
        k_fused_step<false><<<(unsigned)((warps + WARPS - 1) / WARPS), WARPS * 32, p->smem_int, h->stream>>>(A, h->map_u[h->cur],
                                                                                                        h->map_shape);
        h->launches++;
    }
    if (p->n_gen > 0 && !(dbg_skip & 2)) {
        A.items = p->d_items_gen;
        A.n_items = p->n_gen;
        A.epart_off = p->n_int;
        long long warps = 2LL * p->n_gen * h->gp.n_env;  // one warp per wavefield
        k_fused_step<true><<<(unsigned)((warps + WARPS - 1) / WARPS), WARPS * 32, p->smem_gen, h->stream>>>(A, h->map_u6[h->cur],
                                                                                                       h->map_shape);
        h->launches++;
        if (d_e3) {
            long long ew = (long long)p->n_gen * h->gp.n_env;
            k_energy_items<<<(unsigned)((ew + 3) / 4), 128, 0, h->stream>>>(h->gp, A.out, p->d_items_gen, p->n_gen, p->d_epart,
                                                                            A.epart_stride, p->n_int);
            h->launches++;
        }
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

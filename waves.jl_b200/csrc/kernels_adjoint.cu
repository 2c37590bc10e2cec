// Reverse pass of the integrator: adjoint_sensitivity / rrule(::Integrator) (src/dynamics.jl:97-128) for the 2-D
// AcousticDynamics (src/dynamics.jl:151-188).  One kernel applies the TRANSPOSED right-hand side
//   (J^T λ)_U  = -(σx+σy) m λU + Dx^T λVx + Dy^T λVy + σxσy λΩ          m = Dirichlet mask (src/dims.jl:117-124)
//   (J^T λ)_Vx = Dx^T[b (m λU + σy λΨy)] - σx λVx
//   (J^T λ)_Vy = Dy^T[b (m λU + σx λΨx)] - σy λVy
//   (J^T λ)_Ψx = (J^T λ)_Ψy = m λU ;  (J^T λ)_Ω = -m λU
// and accumulates the sensitivity to the speed plane of the total field,
//   dL/dc += 2 c [ m λU (DxVx + DyVy) + σx λΨx DyVy + σy λΨy DxVx ]      (derivatives of the forward stage state),
// where D is the 3-band matrix of src/operators.jl:10-22 with its one-sided first / last rows and D^T its transpose.
// The step-level reverse of runge_kutta (src/dynamics.jl:9-16) is orchestrated in waves_abi.cu (waves_adjoint).
// Checked in tests/test_gpu_adjoint.py against an independent autodiff of the forward program.  sm_100a only.
#include <stdlib.h>

#include "waves_internal.h"

namespace {

// coefficient of v[r] in (D^T v)[i]: D[r][i]
struct Band {
    float gf[3], gc[2], gl[3];
    int n;
    // sum over rows r of D[r][i] * v(r), v given by a callable on the row index
    template <class F>
    __device__ __forceinline__ float transposed(int i, F v) const {
        float acc = 0.0f;
        if (i + 1 <= n - 2) acc += gc[0] * v(i + 1);          // interior row r = i+1 holds gc[0] on column r-1
        if (i - 1 >= 1) acc += gc[1] * v(i - 1);              // interior row r = i-1 holds gc[1] on column r+1
        // (selects instead of gf[i]: dynamic indexing would put the coefficient arrays in local memory)
        if (i <= 2) acc += (i == 0 ? gf[0] : (i == 1 ? gf[1] : gf[2])) * v(0);                             // first row: columns 0,1,2
        if (i >= n - 3) acc += (i == n - 3 ? gl[0] : (i == n - 2 ? gl[1] : gl[2])) * v(n - 1);           // last row: columns n-3,n-2,n-1
        return acc;
    }
    template <class F>
    __device__ __forceinline__ float forward(int i, F v) const {
        if (i == 0) return gf[0] * v(0) + gf[1] * v(1) + gf[2] * v(2);
        if (i == n - 1) return gl[0] * v(n - 3) + gl[1] * v(n - 2) + gl[2] * v(n - 1);
        return gc[0] * v(i - 1) + gc[1] * v(i + 1);
    }
};

// Interior cells of k_rhs_transposed (>= 3 cells from every domain edge): central stencils, mask == 1, neighbours addressed by
// constant offsets from one index.  HAS_LP: the cotangent is a w + b lyp, else a w.
template <bool HAS_LP>
__device__ __forceinline__ void rhs_transposed_interior(const GridP &gp, int e, int wf, int i, int j, float a, const float *__restrict__ w, float b,
                                                        const float *__restrict__ lyp, const float *__restrict__ y, const float *__restrict__ b2,
                                                        float *__restrict__ out, float *__restrict__ ws, int first, float *__restrict__ gcacc) {
    const long long P = gp.plane;
    const int nxp = gp.nxp;
    const long long q = (long long)e * gp.env_stride + (long long)wf * 6 * P + (long long)j * nxp + i;
    const float *W = w + q, *LP = HAS_LP ? lyp + q : nullptr;
    auto ld = [&](int f, int off) { return HAS_LP ? a * W[f * P + off] + b * LP[f * P + off] : a * W[f * P + off]; };
    const float *bp = (wf == 0 && b2) ? b2 + (long long)e * P + (long long)j * nxp + i : nullptr;
    auto bv = [&](int off) { return bp ? bp[off] : gp.b0; };
    const float cm = gp.g_central[0], cp = gp.g_central[1];  // D[r][r-1], D[r][r+1]: (D^T v)[i] = cm v[i+1] + cp v[i-1]
    const float sx = gp.sigma[i], sy = gp.sigma[j];
    const float lU = ld(0, 0), lVx = ld(1, 0), lVy = ld(2, 0), lPx = ld(3, 0), lPy = ld(4, 0), lOm = ld(5, 0);
    const float o0 = -(sx + sy) * lU + (cm * ld(1, 1) + cp * ld(1, -1)) + (cm * ld(2, nxp) + cp * ld(2, -nxp)) + (sx * sy) * lOm;
    const float o1 = (cm * (bv(1) * (ld(0, 1) + sy * ld(4, 1))) + cp * (bv(-1) * (ld(0, -1) + sy * ld(4, -1)))) - sx * lVx;
    const float o2 = (cm * (bv(nxp) * (ld(0, nxp) + sx * ld(3, nxp))) + cp * (bv(-nxp) * (ld(0, -nxp) + sx * ld(3, -nxp)))) - sy * lVy;
    const float o[6] = {o0, o1, o2, lU, lU, -lU};
    float *O = out + q, *WS = ws + q;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        O[f * P] = o[f];
        WS[f * P] = (first ? W[f * P] : WS[f * P]) + o[f];
    }
    if (wf == 0 && gcacc) {
        const float *Y = y + (long long)e * gp.env_stride + (long long)j * nxp + i;
        const float Vxx = cm * Y[P - 1] + cp * Y[P + 1], Vyy = cm * Y[2 * P - nxp] + cp * Y[2 * P + nxp];
        const float gb = lU * (Vxx + Vyy) + (sx * lPx) * Vyy + (sy * lPy) * Vxx;
        gcacc[(long long)e * P + (long long)j * nxp + i] += 2.0f * sqrtf(bv(0)) * gb;
    }
}

// Body of k_rhs_transposed for one cell.  INTERIOR: the cell is at least 3 cells away from every domain edge, so every
// stencil is the central one and the Dirichlet mask is 1 (decided per thread block: no divergence, a fraction of the
// instructions of the general form).
template <bool INTERIOR>
__device__ __forceinline__ void rhs_transposed_cell(const GridP &gp, int e, int wf, int i, int j, float a, const float *__restrict__ w, float b,
                                                    const float *__restrict__ lyp, const float *__restrict__ y, const float *__restrict__ b2,
                                                    float *__restrict__ out, float *__restrict__ ws, int first, float *__restrict__ gcacc) {
    const long long P = gp.plane;
    const int nxp = gp.nxp, nx = gp.nx, ny = gp.ny_global;
    const long long base = (long long)e * gp.env_stride + (long long)wf * 6 * P;
    const float *W = w + base, *LP = lyp ? lyp + base : nullptr;
    float *O = out + base, *WS = ws + base;
    const float *bpl = (wf == 0 && b2) ? b2 + (long long)e * P : nullptr;
    Band bx{{gp.g_first[0], gp.g_first[1], gp.g_first[2]}, {gp.g_central[0], gp.g_central[1]}, {gp.g_last[0], gp.g_last[1], gp.g_last[2]}, nx};
    Band by = bx;
    by.n = ny;
    auto mask = [&](int ii, int jj) { return (!INTERIOR && (ii == 0 || ii == nx - 1 || jj == 0 || jj == ny - 1)) ? 0.0f : 1.0f; };
    auto bval = [&](int ii, int jj) { return bpl ? bpl[(long long)jj * nxp + ii] : gp.b0; };
    auto at = [&](int f, int ii, int jj) {
        const long long q = (long long)f * P + (long long)jj * nxp + ii;
        return LP ? a * W[q] + b * LP[q] : a * W[q];
    };
    // D^T of a callable along x (row j fixed) / along y (column i fixed)
    auto dxT = [&](auto v) { return INTERIOR ? bx.gc[0] * v(i + 1) + bx.gc[1] * v(i - 1) : bx.transposed(i, v); };
    auto dyT = [&](auto v) { return INTERIOR ? by.gc[0] * v(j + 1) + by.gc[1] * v(j - 1) : by.transposed(j, v); };
    const float sx = gp.sigma[i], sy = gp.sigma[j];
    const long long q = (long long)j * nxp + i;
    const float lU = at(0, i, j), lVx = at(1, i, j), lVy = at(2, i, j), lPx = at(3, i, j), lPy = at(4, i, j), lOm = at(5, i, j);
    const float mU = mask(i, j) * lU;
    const float dxT_lVx = dxT([&](int r) { return at(1, r, j); });
    const float dyT_lVy = dyT([&](int r) { return at(2, i, r); });
    const float dxT_qx = dxT([&](int r) { return bval(r, j) * (mask(r, j) * at(0, r, j) + sy * at(4, r, j)); });
    const float dyT_qy = dyT([&](int r) { return bval(i, r) * (mask(i, r) * at(0, i, r) + sx * at(3, i, r)); });
    float o[6];
    o[0] = -(sx + sy) * mU + dxT_lVx + dyT_lVy + (sx * sy) * lOm;
    o[1] = dxT_qx - sx * lVx;
    o[2] = dyT_qy - sy * lVy;
    o[3] = mU;
    o[4] = mU;
    o[5] = -mU;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        O[f * P + q] = o[f];
        WS[f * P + q] = (first ? W[f * P + q] : WS[f * P + q]) + o[f];
    }
    if (wf == 0 && gcacc) {
        const float *Y = y + (long long)e * gp.env_stride;
        auto vx = [&](int r) { return Y[P + (long long)j * nxp + r]; };
        auto vy = [&](int r) { return Y[2 * P + (long long)r * nxp + i]; };
        const float Vxx = INTERIOR ? bx.gc[0] * vx(i - 1) + bx.gc[1] * vx(i + 1) : bx.forward(i, vx);
        const float Vyy = INTERIOR ? by.gc[0] * vy(j - 1) + by.gc[1] * vy(j + 1) : by.forward(j, vy);
        const float gb = mU * (Vxx + Vyy) + (sx * lPx) * Vyy + (sy * lPy) * Vxx;
        gcacc[(long long)e * P + q] += 2.0f * sqrtf(bval(i, j)) * gb;
    }
}

// blockIdx.z = env * 2 + wavefield.  The cotangent is formed on the fly as lam = a * w + b * lyp (lyp nullable): the step-level
// reverse (waves_adjoint) never materialises it.  w, lyp, y, out, ws: [n_env][12][plane]; b2: [n_env][plane] or nullptr
// (scalar c0^2); gcacc: [n_env][plane] or nullptr.  Besides out = J^T lam the kernel keeps the running sum of the step:
// ws = (first ? w : ws) + out.  out must not alias lyp (neighbouring cells read it).
__global__ void __launch_bounds__(256)
k_rhs_transposed(GridP gp, float a, const float *__restrict__ w, float b, const float *__restrict__ lyp, const float *__restrict__ y,
                 const float *__restrict__ b2, float *__restrict__ out, float *__restrict__ ws, int first, float *__restrict__ gcacc) {
    const int e = blockIdx.z >> 1, wf = blockIdx.z & 1;
    const int i0 = blockIdx.x * blockDim.x, j0 = blockIdx.y * blockDim.y;
    const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
    const bool interior_block = i0 >= 3 && i0 + (int)blockDim.x + 3 <= gp.nx && j0 >= 3 && j0 + (int)blockDim.y + 3 <= gp.ny_own;
    if (i >= gp.nx || j >= gp.ny_own) return;
    if (interior_block) {
        if (lyp)
            rhs_transposed_interior<true>(gp, e, wf, i, j, a, w, b, lyp, y, b2, out, ws, first, gcacc);
        else
            rhs_transposed_interior<false>(gp, e, wf, i, j, a, w, b, lyp, y, b2, out, ws, first, gcacc);
    } else
        rhs_transposed_cell<false>(gp, e, wf, i, j, a, w, b, lyp, y, b2, out, ws, first, gcacc);
}

// out = a*x + b*y + c*z (y, z nullable); elementwise over the whole state
__global__ void k_lin3(long long n, float *__restrict__ out, float a, const float *__restrict__ x, float b, const float *__restrict__ y,
                       float c, const float *__restrict__ z) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    float v = a * x[q];
    if (y) v += b * y[q];
    if (z) v += c * z[q];
    out[q] = v;
}

// w += dL/dz_i of the energy-weighted loss sum_k w3[k] E_k (src/env.jl:104-111): only the two U planes are touched.
// z: U_tot of environment e at z + e * z_env + q, U_inc at + z_inc (a full state, or the two gathered U planes).
__global__ void k_energy_cotangent(GridP gp, const float *__restrict__ z, long long z_env, long long z_inc, float *__restrict__ w,
                                   float w_tot, float w_inc, float w_sc, float d_omega) {
    const int e = blockIdx.z;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= gp.nx || j >= gp.ny_own) return;
    const long long qz = (long long)e * z_env + (long long)j * gp.nxp + i;
    const long long q = (long long)e * gp.env_stride + (long long)j * gp.nxp + i;
    const float ut = z[qz], ui = z[qz + z_inc], d = ut - ui;
    const float two = 2.0f * d_omega;
    w[q] += two * (w_tot * ut + w_sc * d);
    w[q + 6 * gp.plane] += two * (w_inc * ui - w_sc * d);
}

}  // namespace

void launch_rhs_transposed(waves_handle *h, float a, const float *w, float b, const float *lyp, const float *y, const float *b2,
                           float *out, float *ws, int first, float *gcacc) {
    static const int bx = waves_dev_env("WAVES_DEBUG_ADJ_BX", 64);  // developer tuning aid (measured: 64 x 4 tiles)
    dim3 blk(bx, 256 / bx), grd((h->gp.nx + bx - 1) / bx, (h->gp.ny_own + blk.y - 1) / blk.y, h->gp.n_env * 2);
    k_rhs_transposed<<<grd, blk, 0, h->stream>>>(h->gp, a, w, b, lyp, y, b2, out, ws, first, gcacc);
    h->launches++;
}

void launch_lin3(waves_handle *h, float *out, float a, const float *x, float b, const float *y, float c, const float *z) {
    const long long n = h->gp.env_stride * h->gp.n_env;
    k_lin3<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, out, a, x, b, y, c, z);
    h->launches++;
}

void launch_energy_cotangent(waves_handle *h, const float *z, long long z_env, long long z_inc, float *w, const float *w3) {
    dim3 blk(32, 8), grd((h->gp.nx + 31) / 32, (h->gp.ny_own + 7) / 8, h->gp.n_env);
    k_energy_cotangent<<<grd, blk, 0, h->stream>>>(h->gp, z, z_env, z_inc, w, w3[0], w3[1], w3[2], h->d_omega);
    h->launches++;
}

// Race check of the latent kernels' shared-memory protocol -- TEST INFRASTRUCTURE ONLY.
// Built with -fsanitize=thread: the kernels of csrc/latent_core.cuh run under the host emulation (one pthread per CUDA
// thread, __syncthreads() = pthread barrier); ThreadSanitizer reports every pair of conflicting shared/global accesses
// that is not ordered by a barrier, i.e. a missing __syncthreads().  Exit code 0 and no "WARNING: ThreadSanitizer" = clean.
#include <math.h>
#include <stdio.h>

#include <vector>

#include "latent_emu.cpp"

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 70, nt = argc > 2 ? atoi(argv[2]) : 96, steps = argc > 3 ? atoi(argv[3]) : 6;
    const int batch = 1, nseq = 3, T = steps + 1;
    const bool with_energy = argc > 4 ? atoi(argv[4]) != 0 : true;  // 0: no emulation-only barriers in the forward kernels
    std::vector<float> z0(4 * n), tspan(T), X(nseq), Y((size_t)nseq * n), shape(n), pml(n), z((size_t)T * 4 * n), e(3 * T),
        last(4 * n), wE(3 * T, 1.0f), dz((size_t)T * 4 * n, 0.01f), g0(4 * n), gY((size_t)nseq * n, 0.0f), gs(n), gp(n);
    for (int i = 0; i < n; ++i) {
        const float x = -100.0f + 200.0f * i / (n - 1);
        z0[i] = expf(-x * x / 200.0f);
        z0[n + i] = 0.001f * sinf(x);
        z0[2 * n + i] = expf(-(x - 5) * (x - 5) / 300.0f);
        z0[3 * n + i] = 0.0f;
        shape[i] = expf(-(x + 20) * (x + 20) / 100.0f);
        const float r = fmaxf(fabsf(x) - 90.0f, 0.0f) / 10.0f;
        pml[i] = r * r * r;
        for (int k = 0; k < nseq; ++k) Y[(size_t)k * n + i] = 1.0f + 0.1f * k + 0.2f * sinf(0.1f * i);
    }
    for (int s = 0; s < T; ++s) tspan[s] = 1e-5f * s;
    X[0] = tspan[0];
    X[1] = tspan[steps / 2];
    X[2] = tspan[steps];
    LatentP p = {};
    p.n = n; p.batch = batch; p.steps = steps; p.nseq = nseq;
    p.c0 = 1531.0f; p.dt = 1e-5f; p.hdt = 0.5e-5f; p.pml_scale = 10000.0f; p.freq = 1000.0f; p.dx = 200.0f / (n - 1);
    const float k2 = 1.0f / (2.0f * p.dx);
    p.gf[0] = -3 * k2; p.gf[1] = 4 * k2; p.gf[2] = -k2; p.gc[0] = -k2; p.gc[1] = k2; p.gl[0] = k2; p.gl[1] = -4 * k2; p.gl[2] = 3 * k2;
    p.z0 = z0.data(); p.tspan = tspan.data(); p.X = X.data(); p.Y = Y.data(); p.shape = shape.data(); p.pml = pml.data();
    p.z = z.data(); p.energy = with_energy ? e.data() : nullptr; p.z_last = last.data();
    emu_latent_integrate(&p, nt);
    if (nt >= n) emu_latent_integrate_r1(&p, nt);
    if (n % 2 == 0 && nt >= n / 2) emu_latent_integrate_r2(&p, ((n / 2 + 31) / 32) * 32);
    p.zt = z.data(); p.w_energy = wE.data(); p.dL_dz = dz.data(); p.g_z0 = g0.data(); p.g_Y = gY.data(); p.g_shape = gs.data();
    p.g_pml = gp.data();
    for (int compat = 0; compat < 2; ++compat) {
        p.compat = compat;
        emu_latent_adjoint(&p, nt);
        if (nt >= n) emu_latent_adjoint_r1(&p, nt);
        if (n % 2 == 0 && nt >= n / 2) emu_latent_adjoint_r2(&p, ((n / 2 + 31) / 32) * 32);
    }
    double chk = 0;
    for (float v : g0) chk += v;
    printf("done n=%d nt=%d steps=%d checksum %.6g energy %.6g\n", n, nt, steps, chk, (double)e[3 * T - 1]);
    return isfinite(chk) ? 0 : 3;
}

"""Seeded inputs of the 1-D latent dynamics tests (shared by the CPU emulation tests and the GPU parity tests)."""
import numpy as np

from oracle import latent_oracle as lo
from oracle import waves_oracle as wo

F32 = np.float32


def make_case(n=96, batch=3, steps=24, nseq=4, seed=0, gs=100.0, dt=1e-5, c0=1531.0, pml_width=10.0, pml_scale=10000.0,
              freq=1000.0, t0=0.0, knots="actions"):
    """A batch like the one AcousticEnergyModel integrates (scripts/main.jl:120-141: latent_gs = 100, pml_width = 10,
    pml_scale = 10000, dt = 1e-5), at a size the oracle finishes in seconds.  Every sample has its own time column."""
    rng = np.random.default_rng(seed)
    dim = wo.OneDim.make(gs, n)
    dyn = lo.LatentDynamics.make(dim, c0, pml_width, pml_scale)
    x = dim.x.astype(np.float64)
    tspan = np.stack([wo.build_tspan(F32(t0 + 1e-3 * b), dt, steps) for b in range(batch)])          # [batch][steps+1]
    if knots == "actions":     # DesignEncoder: t[1:integration_steps:end, :] (src/model/design_encoder.jl:27)
        stride = steps // (nseq - 1)
        assert stride * (nseq - 1) == steps
        X = np.ascontiguousarray(tspan[:, ::stride])
    elif knots == "repeated":  # knots that are not increasing: two segment masks are true at once (their terms add up)
        X = np.ascontiguousarray(tspan[:, [0, steps, steps // 2, steps]])
        assert nseq == 4
    else:                      # knots that do not cover the whole span: queries outside every segment give C = 0
        X = np.stack([np.linspace(tspan[b, 2], tspan[b, -3], nseq).astype(F32) for b in range(batch)])
    Y = (1.0 + 0.5 * rng.random((batch, nseq, n))).astype(F32)                                        # 2σ(·) ∈ (0, 2)
    shape = np.stack([np.exp(-((x - rng.uniform(-30, 30)) / 8.0) ** 2) for _ in range(batch)]).astype(F32)
    pml = (dyn.pml / dyn.pml.max())[None, :].repeat(batch, 0).astype(F32)                             # θ[3] ∈ [0, 1]
    pml = (pml * rng.uniform(0.8, 1.0, (batch, 1))).astype(F32)
    z0 = np.stack([np.stack([np.exp(-((x - rng.uniform(-40, 40)) / 10.0) ** 2) * rng.uniform(0.5, 1.0),
                             1e-3 * rng.standard_normal(n),
                             np.exp(-((x - rng.uniform(-40, 40)) / 12.0) ** 2) * rng.uniform(0.5, 1.0),
                             1e-3 * rng.standard_normal(n)]) for _ in range(batch)]).astype(F32)
    theta = lo.LatentTheta(X=X, Y=Y, shape=shape, freq=F32(freq), pml=pml)
    return dict(dim=dim, dyn=dyn, z0=z0, tspan=tspan, theta=theta, dt=F32(dt), steps=steps, n=n, batch=batch, nseq=nseq)

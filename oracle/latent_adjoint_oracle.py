"""Oracle for the reverse pass of the 1-D latent dynamics -- TEST INFRASTRUCTURE ONLY (only tests/ may import it).

Restates `adjoint_sensitivity` for the batchwise OneDim simulation (src/dynamics.jl:97-118) and the
`rrule(::Integrator, z0, t, θ)` that calls it (:120-128) in torch float64:

  * exact mode  -- reverse-mode autodiff through the whole unrolled trajectory (what Zygote would compute without the
                   custom rrule): the independent truth of the discrete adjoint.
  * compat mode -- the reference loop AS WRITTEN: for i = N+1 … 1 { acc += aᵢ; (δz, δθ) = pullback of ONE runge_kutta
                   call at the stored (zᵢ, tᵢ) applied to acc; acc += δz; gθ += δθ }.

θ = [C, F, PML]: gradients are returned for C.Y, F.shape and PML (the trainable leaves reached in
src/model/acoustic_energy_model.jl:86-94) and for z0.  The loss is  Σ w_energy ⊙ compute_latent_energy(z, dx) + Σ dL_dz ⊙ z.

PARITY UNPINNED: the reference has no test or fixture for this path (scripts/adjoint_sensitivity.jl only plots).
"""
from __future__ import annotations

import numpy as np

from . import latent_oracle as lo
from . import waves_oracle as wo

F32 = np.float32


def _torch():
    import torch
    return torch


def _interp_weights(X: np.ndarray, t: np.ndarray):
    """The masks of linear_interp (src/utils.jl:79-80) evaluated in Float32 like the forward pass, and the Float64
    interpolation weights they select: c = Σ_k W0[b,k] Y[b,k] + W1[b,k] Y[b,k+1]."""
    xr = t[:, None]
    left, r = X[:, :-1], X[:, 1:]
    mask = ((left <= xr) & (xr < r)) | ((r == r[:, -1:]) & (r[:, -1:] == xr))
    x0 = np.where(mask, left, 0).astype(np.float64).sum(1, keepdims=True)
    dd = (r.astype(np.float64) - left.astype(np.float64))
    w = (xr.astype(np.float64) - x0) / dd
    W1 = np.where(mask, w, 0.0)
    W0 = np.where(mask, 1.0 - w, 0.0)
    return W0, W1


def rk4_du_torch(z, t32, cs, Y, shape, pml):
    """runge_kutta (src/dynamics.jl:9-16) of the OneDim dynamics (:190-222); z [batch][4][n] float64 tensor."""
    torch = _torch()
    dyn, th = cs["dyn"], cs["theta"]
    D = torch.tensor(dyn.grad.to_scipy().toarray().astype(np.float64))
    bc = torch.tensor(dyn.bc.astype(np.float64))
    c0, scale, freq = float(dyn.c0), float(dyn.pml[0]), float(th.freq)
    dt, hdt = float(cs["dt"]), float(F32(0.5) * cs["dt"])
    sig = scale * pml

    def rhs(w, tq):
        W0, W1 = _interp_weights(th.X, tq)
        c = (torch.tensor(W0)[:, :, None] * Y[:, :-1]).sum(1) + (torch.tensor(W1)[:, :, None] * Y[:, 1:]).sum(1)
        fac = torch.tensor([float(wo.source_factor(tb, freq, F32)) for tb in tq], dtype=torch.float64)
        f = shape * fac[:, None]
        Ut, Vt, Ui, Vi = w[:, 0], w[:, 1], w[:, 2], w[:, 3]
        a = c0 * c
        return torch.stack([(a * (Vt @ D.T) - sig * Ut) * bc, a * ((Ut + f) @ D.T) - sig * Vt,
                            (c0 * (Vi @ D.T) - sig * Ui) * bc, c0 * ((Ui + f) @ D.T) - sig * Vi], 1)

    t1 = (t32 + F32(hdt)).astype(F32)
    t2 = (t32 + F32(dt)).astype(F32)
    k1 = rhs(z, t32)
    k2 = rhs(z + hdt * k1, t1)
    k3 = rhs(z + hdt * k2, t1)
    k4 = rhs(z + dt * k3, t2)
    return (1.0 / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4) * dt


def _energy_loss(zi, w3, dx):
    """Σ_k w3[b,k] E_k(zi): compute_latent_energy of one frame (src/model/acoustic_energy_model.jl:6-15)."""
    torch = _torch()
    tot, inc = zi[:, 0], zi[:, 2]
    E = torch.stack([(tot ** 2).sum(1), (inc ** 2).sum(1), ((tot - inc) ** 2).sum(1)], 1) * dx
    return (E * w3).sum()


def adjoint_truth(cs, w_energy=None, dL_dz=None, compat=False, z_stored=None):
    """dict(z0, Y, shape, pml) of float64 gradients."""
    torch = _torch()
    th = cs["theta"]
    N = cs["steps"]
    dx = float(wo.get_dx(cs["dim"]))
    leaf = lambda a: torch.tensor(a.astype(np.float64), requires_grad=True)  # noqa: E731
    Y, shape, pml = leaf(th.Y), leaf(th.shape), leaf(th.pml)
    wE = None if w_energy is None else torch.tensor(w_energy.astype(np.float64))
    aZ = None if dL_dz is None else torch.tensor(dL_dz.astype(np.float64))
    tcol = lambda i: np.ascontiguousarray(cs["tspan"][:, i])  # noqa: E731

    def frame_loss(zi, i):
        L = 0.0
        if wE is not None:
            L = L + _energy_loss(zi, wE[:, :, i], dx)
        if aZ is not None:
            L = L + (aZ[i] * zi).sum()
        return L

    if not compat:
        z0 = leaf(cs["z0"])
        z, L = z0, frame_loss(z0, 0)
        for i in range(N):
            z = z + rk4_du_torch(z, tcol(i), cs, Y, shape, pml)
            L = L + frame_loss(z, i + 1)
        L.backward()
        return {"z0": z0.grad.numpy(), "Y": Y.grad.numpy(), "shape": shape.grad.numpy(), "pml": pml.grad.numpy()}

    if z_stored is None:
        z_stored = lo.integrate(cs["dyn"], cs["z0"], cs["tspan"], th, cs["dt"])
    acc = torch.zeros(cs["z0"].shape, dtype=torch.float64)
    gY, gs, gp = torch.zeros_like(Y), torch.zeros_like(shape), torch.zeros_like(pml)
    for i in range(N, -1, -1):
        zi = leaf(z_stored[i])
        ai = torch.autograd.grad(frame_loss(zi, i), zi)[0]            # aᵢ = ∂L_∂z[:, :, :, i]
        acc = acc + ai                                                 # :110
        du = rk4_du_torch(zi, tcol(i), cs, Y, shape, pml)              # :105-107
        dz, dY, dS, dP = torch.autograd.grad(du, [zi, Y, shape, pml], grad_outputs=acc)   # :112
        acc = acc + dz                                                 # :113
        gY, gs, gp = gY + dY, gs + dS, gp + dP                         # :114
    return {"z0": acc.numpy(), "Y": gY.detach().numpy(), "shape": gs.detach().numpy(), "pml": gp.detach().numpy()}

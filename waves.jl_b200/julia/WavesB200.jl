# WavesB200.jl -- thin `ccall` shim that puts libwaves_b200.so (include/waves_b200.h) behind the
# reference's own call surface:
#     (env::WaveEnv)(action)                  src/env.jl:91-121      (seam B1)
#     (iter::Integrator)(ui, tspan, θ)        src/dynamics.jl:37-53  (seam B2)
#
# STATUS: UNVERIFIED.  Julia is not installed in the build image or on the GPU boxes, so this file has
# never been executed; it is written against the C ABI that the Python ctypes mirror exercises
# symbol for symbol (waves.jl_b200/_lib.py, tests/test_gpu_parity.py).  See INTEGRATION.md.
#
# Usage inside the reference repo:
#     include("path/to/WavesB200.jl"); using .WavesB200
#     benv = WavesB200.B200Env(env; lib = "/path/to/libwaves_b200.so", device = 0)
#     while !is_terminated(env); benv(policy(env)); end       # instead of env(policy(env))
module WavesB200

using Waves
using Waves: OneDim, TwoDim, WaveEnv, Integrator, AcousticDynamics, DesignInterpolator, AbstractDesign, NoDesign,
             Cylinders, Cloak, AbstractScatterers, NoSource, build_tspan, get_dx, get_dy, stack
using SparseArrays
using Libdl

# The library is opened once with Libdl and every entry point is resolved with dlsym: `ccall((:name, lib), ...)` needs `lib` to
# be a constant or global expression, so a path held in a local variable or a struct field cannot be used there.
const LIBH = Ref{Ptr{Cvoid}}(C_NULL)
const LIBPATH = Ref{String}("")
"""    load!(path)  -- open libwaves_b200.so (idempotent); every handle of this session uses it"""
function load!(path::AbstractString)
    if LIBH[] == C_NULL || LIBPATH[] != path
        LIBH[] = Libdl.dlopen(path)
        LIBPATH[] = String(path)
    end
    return LIBH[]
end
sym(name::Symbol) = (LIBH[] == C_NULL && error("waves_b200: call WavesB200.load!(\"/path/to/libwaves_b200.so\") first"); Libdl.dlsym(LIBH[], name))

const MODE_FUSED = Cint(0)
const MODE_EXACT = Cint(1)

# struct waves_config (include/waves_b200.h) -- field order and types must match exactly
struct WavesConfig
    nx::Int32; ny::Int32; n_env::Int32; device::Int32
    c0::Float32; dt::Float32; pml_width::Float32; pml_scale::Float32
    x::Ptr{Float32}; y::Ptr{Float32}; sigma::Ptr{Float32}; grad8::Ptr{Float32}
    d_omega::Float32; ny_global::Int32; row0::Int32; flags::UInt32
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    lib::String
end

function check(h_lib::String, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall(sym(:waves_last_error), Cstring, ()))
    error("waves_b200: " * msg)
end

# rows of dyn.grad (src/operators.jl:10-22): first(3) central(2) last(3)
function grad8(grad::AbstractMatrix{Float32})
    g = Matrix(grad)   # tiny: only the three distinct rows are read
    n = size(g, 1)
    return Float32[g[1, 1], g[1, 2], g[1, 3], g[2, 1], g[2, 3], g[n, n-2], g[n, n-1], g[n, n]]
end

"""cylinder table (ncyl, 4) rows {x, y, r, c}, row-major, of the stacked design (src/designs.jl:133-138, :228)"""
cyl_table(c::Cylinders) = Matrix{Float32}(permutedims(hcat(c.pos, c.r, c.c)))      # 4 x ncyl column-major == (ncyl,4) row-major
cyl_table(d::AbstractScatterers) = cyl_table(d.cylinders)
cyl_table(d::Cloak) = cyl_table(stack(d.config.cylinders, d.core))

function create(dyn::AcousticDynamics{TwoDim}, dt::Float32; lib::String, device::Integer = 0, n_env::Integer = 1)
    dim = dyn.dim
    x, y = Vector{Float32}(dim.x), Vector{Float32}(dim.y)
    sigma = Vector{Float32}(dyn.pml[:, 1])
    g8 = grad8(dyn.grad)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve x y sigma g8 begin
        cfg = Ref(WavesConfig(length(x), length(y), n_env, device, dyn.c0, dt, 0f0, 0f0,
                              pointer(x), pointer(y), pointer(sigma), pointer(g8),
                              get_dx(dim) * get_dy(dim), length(y), 0, 0))
        check(lib, ccall(sym(:waves_create), Cint, (Ref{WavesConfig}, Ref{Ptr{Cvoid}}), cfg, out))
    end
    h = Handle(out[], lib)
    finalizer(h -> (h.ptr != C_NULL && ccall(sym(:waves_destroy), Cint, (Ptr{Cvoid},), h.ptr); h.ptr = C_NULL), h)
    return h
end

set_state!(h::Handle, u12::Array{Float32, 3}, env::Integer = -1) =
    check(h.lib, ccall(sym(:waves_set_state), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}), h.ptr, env, u12))

function get_state(h::Handle, nx, ny, env::Integer = 0)
    u = Array{Float32}(undef, nx, ny, 12)
    check(h.lib, ccall(sym(:waves_get_state), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}), h.ptr, env, u))
    return u
end

set_source!(h::Handle, ::NoSource, env::Integer = -1) =
    check(h.lib, ccall(sym(:waves_set_source), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Cfloat), h.ptr, env, C_NULL, 0f0))
set_source!(h::Handle, src, env::Integer = -1) =   # Source / RandomPosGaussianSource: fields shape, freq
    check(h.lib, ccall(sym(:waves_set_source), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Cfloat), h.ptr, env, Array{Float32}(src.shape), src.freq))

set_design!(h::Handle, ::Nothing, env::Integer = -1) =
    check(h.lib, ccall(sym(:waves_set_design), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float32}, Ptr{Float32}, Cfloat, Cfloat), h.ptr, env, 0, C_NULL, C_NULL, 0f0, 0f0))
function set_design!(h::Handle, interp::DesignInterpolator, env::Integer = -1)
    interp.initial isa NoDesign && return set_design!(h, nothing, env)
    a, b = cyl_table(interp.initial), cyl_table(interp.final)
    check(h.lib, ccall(sym(:waves_set_design), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float32}, Ptr{Float32}, Cfloat, Cfloat),
                       h.ptr, env, size(a, 2), a, b, interp.ti, interp.tf))
end

"""
    observe(h, frames, resolution)

The image of `RLBase.state(env)` (src/env.jl:132-137) computed on the device: `imresize(cat(frames[:, :, 1, :], shape; dims = 3),
resolution)` for the `(nx, ny, 12, nsave)` block `integrate!` returned.  Returns `(res_x, res_y, nsave + 1)`.
"""
function observe(h::Handle, frames::Array{Float32, 4}, resolution::Tuple{Int, Int})
    nsave = size(frames, 4)
    out = Array{Float32}(undef, resolution[1], resolution[2], nsave + 1)
    check(h.lib, ccall(sym(:waves_observe), Cint, (Ptr{Cvoid}, Ptr{Float32}, Cint, Cint, Cint, Ptr{Float32}),
                       h.ptr, frames, nsave, resolution[1], resolution[2], out))
    return out
end

"""One DesignInterpolator per environment of a batch handle over a common [ti, tf]: `a`, `b` are (4, ncyl, n_env) arrays."""
set_design_batch!(h::Handle, a::Array{Float32, 3}, b::Array{Float32, 3}, ti::Float32, tf::Float32) =
    check(h.lib, ccall(sym(:waves_set_design_batch), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Ptr{Float32}, Cfloat, Cfloat),
                       h.ptr, size(a, 2), a, b, ti, tf))

"""
    integrate!(h, tspan; save_steps, frames, energy, mode)

(iter::Integrator)(ui, tspan, θ) + the energy metric (src/dynamics.jl:37-53, src/env.jl:104-114) for the
handle's environments.  Returns `(energy::Matrix{Float32} (3, steps+1), frames::Array{Float32,4} (nx,ny,12,nsave))`
for a single-environment handle.
"""
function integrate!(h::Handle, tspan::Vector{Float32}, nx::Int, ny::Int; save_steps::Vector{Int32} = Int32[],
                    mode::Cint = MODE_FUSED, u_tot = C_NULL, u_inc = C_NULL)
    steps = length(tspan) - 1
    energy = Array{Float32}(undef, 3, steps + 1)                  # column-major (3, steps+1) == C (steps+1, 3)
    frames = Array{Float32}(undef, nx, ny, 12, length(save_steps))
    check(h.lib, ccall(sym(:waves_integrate), Cint,
                       (Ptr{Cvoid}, Ptr{Float32}, Cint, Cint, Ptr{Float32}, Ptr{Int32}, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                       h.ptr, tspan, steps, mode, energy, save_steps, length(save_steps), frames, u_tot, u_inc))
    return energy, frames
end

"""
    adjoint!(h, tspan, nx, ny; w_energy, dL_dzN, fwd_mode, adj_mode)

rrule(::Integrator, z0, t, θ) + adjoint_sensitivity (src/dynamics.jl:97-128) from the handle's current state for the
loss  L = Σᵢ Σₖ w_energy[k, i] Eₖ(zᵢ) + ⟨dL_dzN, z_N⟩.  `adj_mode = 0` is the exact discrete adjoint, `1` the reference
loop as written.  Returns `(loss, ∂L/∂z0 (nx,ny,12), ∂L/∂c (nx,ny))` for a single-environment handle.
"""
function adjoint!(h::Handle, tspan::Vector{Float32}, nx::Int, ny::Int; w_energy = C_NULL, dL_dzN = C_NULL,
                  fwd_mode::Cint = MODE_FUSED, adj_mode::Cint = Cint(0))
    steps = length(tspan) - 1
    gz = Array{Float32}(undef, nx, ny, 12)
    gc = Array{Float32}(undef, nx, ny)
    loss = zeros(Float32, 1)
    check(h.lib, ccall(sym(:waves_adjoint), Cint,
                       (Ptr{Cvoid}, Ptr{Float32}, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                       h.ptr, tspan, steps, fwd_mode, adj_mode, w_energy, dL_dzN, gz, gc, loss))   # w_energy: (3, steps+1) column-major
    return loss[1], gz, gc
end

# ---- controls of the two calls above ----------------------------------------------------------------------
"""u_tot / u_inc trajectories of `integrate!` keep every `stride`-th frame only (what a renderer needs, src/plot.jl:25)"""
set_traj_stride!(h::Handle, stride::Integer) =
    check(h.lib, ccall(sym(:waves_set_traj_stride), Cint, (Ptr{Cvoid}, Cint), h.ptr, stride))
"""how `integrate!` issues kernels: 0 every kernel directly, 1 (default) CUDA-graph replay / one launch per step, 2 cooperative multi-step launches"""
set_graph!(h::Handle, mode::Integer) = check(h.lib, ccall(sym(:waves_set_graph), Cint, (Ptr{Cvoid}, Cint), h.ptr, mode))
"""`adjoint!` stores a checkpoint every `every` steps and re-runs one segment at a time (0: chosen from the free device memory)"""
set_adjoint_checkpoint!(h::Handle, every::Integer) =
    check(h.lib, ccall(sym(:waves_set_adjoint_checkpoint), Cint, (Ptr{Cvoid}, Cint), h.ptr, every))
"""wait for everything queued on the handle's stream"""
sync!(h::Handle) = check(h.lib, ccall(sym(:waves_sync), Cint, (Ptr{Cvoid},), h.ptr))

# ---- seam B2: (iter::Integrator)(ui, tspan, θ) for the parameterised θ the environment builds -------------
"""Drop-in for `iter(ui, tspan, [C, F])` when C is the design interpolation of `(env::WaveEnv)(action)`
(src/env.jl:96-99): pass the DesignInterpolator (or `nothing` for a constant c0) instead of the closure."""
function integrate(iter::Integrator, h::Handle, ui::Array{Float32, 3}, tspan::Vector{Float32}, interp, source)
    nx, ny = size(ui, 1), size(ui, 2)
    set_state!(h, ui)
    set_design!(h, interp)
    set_source!(h, source)
    steps = length(tspan) - 1
    _, frames = integrate!(h, tspan, nx, ny; save_steps = Int32.(0:steps))
    return frames                                                  # (nx, ny, 12, steps+1) like cat(ui, ...; dims = 4)
end

# ---- seam B1: (env::WaveEnv)(action) ------------------------------------------------------------------------
mutable struct B200Env
    env::WaveEnv
    h::Handle
    return_frames::Bool          # also return the full u_tot/u_inc trajectories (only render! needs them)
end

function B200Env(env::WaveEnv; lib::String, device::Integer = 0, return_frames::Bool = false)
    h = create(env.iter.dynamics, env.dt; lib = lib, device = device)
    set_state!(h, Array{Float32}(env.wave[:, :, :, end]))
    set_source!(h, env.source)
    return B200Env(env, h, return_frames)
end

const FRAMESKIP = 10   # src/env.jl:90

function (b::B200Env)(action::AbstractDesign)
    env = b.env
    tspan = build_tspan(env)                                       # src/env.jl:92
    ti = time(env)
    next_design = env.design_space(env.design, action)             # src/env.jl:96
    interp = DesignInterpolator(env.design, next_design, ti, tspan[end])
    set_design!(b.h, interp)
    nx, ny = size(env.dim)
    n = env.integration_steps
    u_tot = b.return_frames ? Array{Float32}(undef, nx, ny, n + 1) : C_NULL
    u_inc = b.return_frames ? Array{Float32}(undef, nx, ny, n + 1) : C_NULL
    energy, frames = integrate!(b.h, tspan, nx, ny; save_steps = Int32[n - 2FRAMESKIP, n - FRAMESKIP, n], u_tot = u_tot, u_inc = u_inc)
    env.signal = permutedims(energy)                               # (steps+1, 3)  src/env.jl:114
    env.design = next_design
    env.wave = frames                                              # frames 81, 91, 101  src/env.jl:116
    env.time_step += n
    return tspan, interp, (b.return_frames ? u_tot : nothing), (b.return_frames ? u_inc : nothing)
end

"""Call after `reset!(env)` (src/env.jl:81-88): pushes the zeroed wave and the re-drawn source to the device."""
function sync_reset!(b::B200Env)
    set_state!(b.h, Array{Float32}(b.env.wave[:, :, :, end]))
    set_source!(b.h, b.env.source)
end

# ---- 1-D latent dynamics: model.iter(z0, t, θ) of AcousticEnergyModel (src/model/acoustic_energy_model.jl:86-106) -------
# struct waves_latent_config (include/waves_b200.h)
struct LatentConfig
    n::Int32; device::Int32
    c0::Float32; dt::Float32; pml_width::Float32; pml_scale::Float32; pml0::Float32; dx::Float32
    x::Ptr{Float32}; grad8::Ptr{Float32}
end

mutable struct LatentHandle
    ptr::Ptr{Cvoid}
    lib::String
    n::Int
end

"""Handle for `Integrator(runge_kutta, dyn::AcousticDynamics{OneDim}, dt)` (src/dynamics.jl:18-22, :190-222)."""
function create_latent(iter::Integrator; lib::String, device::Integer = 0)
    dyn = iter.dynamics
    x = Vector{Float32}(dyn.dim.x)
    g8 = grad8(dyn.grad)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve x g8 begin
        cfg = LatentConfig(length(x), device, dyn.c0, iter.dt, 0f0, 0f0, Array(dyn.pml)[1], get_dx(dyn.dim),
                           pointer(x), pointer(g8))                       # pml0 = dyn.pml[[1]] (src/dynamics.jl:192)
        check(lib, ccall(sym(:waves_latent_create), Cint, (Ref{LatentConfig}, Ref{Ptr{Cvoid}}), cfg, out))
    end
    return LatentHandle(out[], lib, length(x))
end

# Raw address of a host `Array` or of a `CuArray` (the library asks the CUDA runtime which kind it got); `nothing` -> NULL.
# For a CuArray `pointer(a)` is a `CuPtr{Float32}`; its bits are the device address.
rawptr(::Nothing) = Ptr{Float32}(C_NULL)
rawptr(a::Array{Float32}) = pointer(a)
rawptr(a::AbstractArray{Float32}) = reinterpret(Ptr{Float32}, pointer(a))

"""
    latent_integrate(h, z0, t, C, F, PML; want_z = true)

`iter(z0, t, [C, F, PML])` (src/dynamics.jl:37-49) and `compute_latent_energy(z, dx)` (src/model/acoustic_energy_model.jl:6-15)
in one launch.  z0 (n, 4, batch), t (steps+1, batch), C::LinearInterpolation (X (nseq, batch), Y (n, nseq, batch)),
F::Source (shape (n, batch)), PML (n, batch): all host `Array`s or all `CuArray`s (outputs are `similar` to `z0`).
Returns `(z (n, 4, batch, steps+1) or nothing, energy (steps+1, 3, batch))`.
"""
function latent_integrate(h::LatentHandle, z0, t, C, F, PML; want_z::Bool = true)
    n, batch, steps, nseq = h.n, size(z0, 3), size(t, 1) - 1, size(C.X, 1)
    z = want_z ? similar(z0, n, 4, batch, steps + 1) : nothing
    energy = similar(z0, steps + 1, 3, batch)
    X, Y, shape = C.X, C.Y, F.shape
    GC.@preserve z0 t X Y shape PML z energy begin
        check(h.lib, ccall(sym(:waves_latent_integrate), Cint,
                           (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
                            Cfloat, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                           h.ptr, batch, steps, nseq, rawptr(z0), rawptr(t), rawptr(X), rawptr(Y), rawptr(shape), F.freq,
                           rawptr(PML), rawptr(z), rawptr(energy), C_NULL))
    end
    return z, energy
end

"""
    latent_adjoint(h, z, t, C, F, PML; w_energy = nothing, dL_dz = nothing, adj_mode = 0)

`adjoint_sensitivity(iter, z, t, θ, ∂L_∂z)` (src/dynamics.jl:97-118): returns `(∂L/∂z0, ∂L/∂C.Y, ∂L/∂F.shape, ∂L/∂PML)`.
`w_energy` (steps+1, 3, batch) is the cotangent of `compute_latent_energy(z, dx)`; `adj_mode = 1` is the reference loop as written.
"""
function latent_adjoint(h::LatentHandle, z, t, C, F, PML; w_energy = nothing, dL_dz = nothing, adj_mode::Integer = 0)
    n, batch, steps, nseq = h.n, size(z, 3), size(t, 1) - 1, size(C.X, 1)
    X, Y, shape = C.X, C.Y, F.shape
    gz0, gY, gS, gP = similar(z, n, 4, batch), similar(Y), similar(shape), similar(PML)
    GC.@preserve z t X Y shape PML w_energy dL_dz gz0 gY gS gP begin
        check(h.lib, ccall(sym(:waves_latent_adjoint), Cint,
                           (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
                            Cfloat, Ptr{Float32}, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32},
                            Ptr{Float32}),
                           h.ptr, batch, steps, nseq, rawptr(z), rawptr(t), rawptr(X), rawptr(Y), rawptr(shape), F.freq,
                           rawptr(PML), adj_mode, rawptr(w_energy), rawptr(dL_dz), rawptr(gz0), rawptr(gY), rawptr(gS),
                           rawptr(gP)))
    end
    return gz0, gY, gS, gP
end

# What replaces `rrule(iter::Integrator, z0, t, θ)` (src/dynamics.jl:120-128) for the OneDim dynamics:
#
#   function Flux.ChainRulesCore.rrule(iter::Integrator, z0::AbstractArray{Float32, 3}, t::AbstractMatrix{Float32}, θ)
#       C, F, PML = θ
#       z, _ = WavesB200.latent_integrate(LATENT[], z0, t, C, F, PML)
#       function Integrator_back(adj)
#           gz0, gY, gS, gP = WavesB200.latent_adjoint(LATENT[], z, t, C, F, PML; dL_dz = adj, adj_mode = 1)
#           return nothing, gz0, nothing, [(X = nothing, Y = gY), (shape = gS, freq = nothing), gP]
#       end
#       return z, Integrator_back
#   end


# ---- episodes written by the Python mirror (waves.jl_b200/data.py: Episode.save -> .npz) as reference Episodes ----------------
# The reference stores episodes as BSON of Julia structs (src/data.jl:60-71: s::Vector{WaveEnvState}, a::Vector{<:AbstractDesign},
# t::Vector{Vector{Float32}}, y::Vector{Matrix{Float32}}), which only Julia can write faithfully.  The batched generator writes
# plain arrays; this turns one file into the reference's Episode (NPZ.jl is the only extra dependency) and
#     FileIO.save(WavesB200.episode_from_npz("episode1.npz", env), "episode1.bson")
# produces exactly what scripts/main.jl:144-150 loads with Episode(path = ...).
"""
    episode_from_npz(path, env::WaveEnv; npzread) -> Waves.Episode

`images` (A, C, res_y, res_x), `t` (A, steps+1), `y` (A, steps+1, 3), `designs` / `actions` (A, ncyl, 4) rows {x, y, r, c} of the
stacked cylinders; `env` supplies `dim` and a design of the right type (the cylinder tables are poured into copies of it).
Pass `npzread = NPZ.npzread` (NPZ.jl keeps NumPy's logical shapes).
"""
function episode_from_npz(path::AbstractString, env::WaveEnv; npzread)
    f = npzread(path)
    A = size(f["t"], 1)
    rows(tab, k0, n) = (Matrix{Float32}(tab[k0:k0+n-1, 1:2]), Vector{Float32}(tab[k0:k0+n-1, 3]), Vector{Float32}(tab[k0:k0+n-1, 4]))
    fill_design(proto::Cylinders, tab, k0 = 1) = Cylinders(rows(tab, k0, length(proto))...)
    fill_design(proto::AbstractScatterers, tab, k0 = 1) = typeof(proto)(fill_design(proto.cylinders, tab, k0))
    fill_design(proto::Cloak, tab, k0 = 1) = Cloak(fill_design(proto.config, tab, 1), fill_design(proto.core, tab, length(proto.config.cylinders) + 1))
    s = Waves.WaveEnvState[]
    a = AbstractDesign[]
    t = Vector{Float32}[]
    y = Matrix{Float32}[]
    for i in 1:A
        img = permutedims(f["images"][i, :, :, :], (3, 2, 1))          # (C, res_y, res_x) -> (res_x, res_y, C), src/env.jl:134-135
        push!(s, Waves.WaveEnvState(env.dim, Vector{Float32}(f["t"][i, :]), Array{Float32}(img), fill_design(env.design, f["designs"][i, :, :])))
        push!(a, fill_design(env.design, f["actions"][i, :, :]))
        push!(t, Vector{Float32}(f["t"][i, :]))
        push!(y, Matrix{Float32}(f["y"][i, :, :]))                      # (steps+1, 3) like env.signal, src/env.jl:114
    end
    return Waves.Episode(s, a, t, y)
end

end # module

# Generates vectors FROM THE REAL REFERENCE that pin oracle/waves_oracle.py and oracle/latent_oracle.py.
#
# STATUS: never run -- the build image and the GPU boxes have no Julia.  Until someone runs it the oracles stay
# "parity unpinned" (DESIGN.md section 2).  To pin them, in a checkout of gladisor/Waves.jl:
#
#     julia --project=. /path/to/tests/golden/make_reference_vectors.jl /path/to/tests/golden
#
# and commit the four ref_*.f32 files it writes (1.2 MB).  tests/test_oracle.py::test_oracle_matches_reference_vectors and
# tests/test_latent_cpu.py::test_latent_oracle_matches_reference_vectors pick them up (they skip while the files are absent).
# Everything is deterministic (no RNG) and runs on the CPU path of the reference.
using Waves, Flux

out = length(ARGS) > 0 ? ARGS[1] : "."
wr(name, a) = open(io -> write(io, Array{Float32}(a)), joinpath(out, name), "w")

# ---- (1) BASELINE config 1: TwoDim(15, 700), Gaussian source at (-10, 0), no design, 100 RK4 steps ---------------------
dim = TwoDim(15.0f0, 700)
dyn = AcousticDynamics(dim, WATER, 2.0f0, 20000.0f0)
iter = Integrator(runge_kutta, dyn, 1f-5)
shape = build_normal(build_grid(dim), [-10.0f0 0.0f0], [0.3f0], [1.0f0])
F = Source(shape, 1000.0f0)
C = t -> WATER                                                      # NoDesign: scalar ambient speed (src/designs.jl:63)
tspan = build_tspan(0.0f0, 1f-5, 100)
sol = iter(build_wave(dim, 12), tspan, [C, F])                      # (700, 700, 12, 101)
dΩ = get_dx(dim) * get_dy(dim)
u_tot, u_inc = sol[:, :, 1, :], sol[:, :, 7, :]
energy = vcat(sum(u_tot .^ 2, dims = (1, 2)) * dΩ, sum(u_inc .^ 2, dims = (1, 2)) * dΩ,
              sum((u_tot .- u_inc) .^ 2, dims = (1, 2)) * dΩ)       # src/env.jl:104-111 -> (3, 1, 101)
wr("ref_config1_energy.f32", reshape(energy, 3, :))                 # memory image [101][3]
wr("ref_config1_final_rows.f32", sol[:, [1, 24, 350, 699], :, end]) # (700, 4, 12): rows j = 1, 24, 350, 699 of all 12 fields

# ---- (2) a 96^2 case with three cylinders whose radii move during the integration (src/env.jl:96-99) -------------------
dim2 = TwoDim(3.0f0, 96)
dyn2 = AcousticDynamics(dim2, WATER, 0.6f0, 20000.0f0)
iter2 = Integrator(runge_kutta, dyn2, 1f-5)
grid2 = build_grid(dim2)
pos = Float32[0.5 0.0; 0.9 0.3; -0.2 -1.1]
d0 = Cylinders(pos, Float32[0.4, 0.3, 0.5], Float32[1032, 1032, 2120])
d1 = Cylinders(pos, Float32[0.6, 0.25, 0.35], Float32[1032, 1032, 2120])
tspan2 = build_tspan(3f-4, 1f-5, 40)
interp = DesignInterpolator(d0, d1, tspan2[1], tspan2[end])
C2 = t -> speed(interp(t), grid2, dyn2.c0)
F2 = Source(build_normal(grid2, [-1.0f0 0.2f0], [0.15f0], [1.0f0]), 1000.0f0)
x = dim2.x
u0 = zeros(Float32, 96, 96, 12)
u0[:, :, 1] .= 1f-3 .* sin.(x) .* cos.(x')                          # deterministic non-zero initial fields
u0[:, :, 7] .= u0[:, :, 1]
sol2 = iter2(u0, tspan2, [C2, F2])
wr("ref_design96_final.f32", sol2[:, :, :, end])                    # (96, 96, 12)

# ---- (3) 1-D latent dynamics (src/dynamics.jl:190-222): 256 elements, batch of 2, 40 steps --------------------------
ldim = OneDim(100.0f0, 256)
ldyn = AcousticDynamics(ldim, WATER, 10.0f0, 10000.0f0)
liter = Integrator(runge_kutta, ldyn, 1f-5)
lx = ldim.x
t = hcat(build_tspan(0.0f0, 1f-5, 40), build_tspan(1f-3, 1f-5, 40))           # (41, 2)
X = t[[1, 21, 41], :]                                                          # knots (3, 2)
Y = zeros(Float32, 256, 3, 2)
for b in 1:2, k in 1:3
    Y[:, k, b] .= 1.0f0 .+ 0.1f0 * k .+ 0.2f0 .* sin.(0.05f0 * b .* lx)
end
Cl = LinearInterpolation(X, Y)
Fl = Source(hcat(exp.(-(lx .+ 20.0f0) .^ 2 ./ 100.0f0), exp.(-(lx .- 10.0f0) .^ 2 ./ 150.0f0)), 1000.0f0)
PML = repeat(ldyn.pml ./ maximum(ldyn.pml), 1, 2)
z0 = zeros(Float32, 256, 4, 2)
z0[:, 1, 1] .= exp.(-lx .^ 2 ./ 200.0f0); z0[:, 3, 1] .= exp.(-(lx .- 5.0f0) .^ 2 ./ 300.0f0)
z0[:, 1, 2] .= exp.(-(lx .+ 30.0f0) .^ 2 ./ 250.0f0); z0[:, 3, 2] .= z0[:, 1, 2]
z = liter(z0, t, [Cl, Fl, PML])                                                # (256, 4, 2, 41)
wr("ref_latent_final_and_energy.f32", vcat(vec(z[:, :, :, end]), vec(compute_latent_energy(z, get_dx(ldim)))))
println("wrote ref_config1_energy.f32, ref_config1_final_rows.f32, ref_design96_final.f32, ref_latent_final_and_energy.f32 to ", out)

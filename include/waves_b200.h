/*
 * waves_b200.h -- C ABI of the B200-native Waves.jl acoustic RK4 hot path.
 *
 * Drop-in boundary.  The reference (gladisor/Waves.jl) has no FFI:
 * the seam is Julia dispatch on
 *     (env::WaveEnv)(action)                      src/env.jl:91-121      (B1)
 *     (iter::Integrator)(ui, tspan, θ)            src/dynamics.jl:37-53  (B2)
 *     runge_kutta(f,u,t,θ,dt) / dyn(x,t,θ)        src/dynamics.jl:9-16, :179-188 (B3)
 * Each entry point below names the reference interface it replaces.  A Julia `ccall`
 * shim (waves.jl_b200/julia/WavesB200.jl) and the Python ctypes mirror
 * (waves.jl_b200/) both bind exactly these symbols.
 *
 * Conventions
 *  - Plain C types only; every pointer may be a HOST or a DEVICE pointer (the library
 *    asks the CUDA runtime which) unless stated otherwise.  The caller owns all buffers
 *    it passes; the handle owns all device memory and streams it creates.
 *  - Layout is the reference's column-major (nx, ny, field[, frame]) == C-order
 *    [frame][field][ny][nx], x fastest, fields tot{U,Vx,Vy,Psix,Psiy,Omega}, inc{...}
 *    (src/dynamics.jl:152-157,185-187).
 *  - Every function returns 0 on success, non-zero on failure; waves_last_error()
 *    returns a thread-local message.  Nothing throws across the ABI.  There is NO CPU
 *    fallback: without a CUDA device waves_create fails.
 *  - A handle is not thread-safe.  Calls that write caller buffers synchronise before
 *    returning; everything else is asynchronous on the handle's stream.
 */
#ifndef WAVES_B200_H
#define WAVES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAVES_B200_VERSION 100 /* 0.1.0 */
#define WAVES_FIELDS 12
#define WAVES_HALO 4 /* ghost rows per side of a slab: one per RK4 stage */

typedef struct waves_handle waves_handle;

/* integrate/step modes */
#define WAVES_MODE_FUSED 0 /* one fused kernel per RK4 step (the product path) */
#define WAVES_MODE_EXACT 1 /* per-stage kernels in the reference's exact float32 evaluation order */
#define WAVES_STEP_ASYNC 0x100 /* OR into the mode of waves_step: return without synchronising the handle's stream */

/* adjoint modes (src/dynamics.jl:97-118) */
#define WAVES_ADJ_EXACT 0  /* exact discrete adjoint  lambda_i = a_i + (I+J_i^T) lambda_{i+1} */
#define WAVES_ADJ_COMPAT 1 /* the reference loop as written (one extra step-vjp) */
#define WAVES_ADJ_STAGEWISE 0x100 /* OR into adj_mode: reverse sweep with the per-stage kernels even without dL_dc (cross-check) */
#define WAVES_ADJ_NO_RING 0x400   /* OR into adj_mode: the marching kernel takes the interior only, the whole PML frame stays on tiles */
#define WAVES_ADJ_RING 0x800      /* OR into adj_mode: the PML ring on the marching kernel even for a small batch (default: batches of
                                     >= 2 M cells; a single environment is latency-bound and better off with fewer launches) */
#define WAVES_ADJ_TILES 0x200     /* OR into adj_mode: the fused reverse step takes the shared-memory tile kernels everywhere instead of
                                     the marching kernel in the interior (cross-check of the two) */

/*
 * Mirrors AcousticDynamics(dim, c0, pml_width, pml_scale) (src/dynamics.jl:141-149)
 * plus the WaveEnv keyword arguments that reach the integrator (src/env.jl:37-50).
 */
typedef struct waves_config {
    int32_t nx, ny;     /* local grid points; the reference needs nx == ny_global (σy = σx', src/dynamics.jl:162) */
    int32_t n_env;      /* independent environments held by this handle (>= 1) */
    int32_t device;     /* CUDA device ordinal */
    float c0;           /* ambient wave speed, dyn.c0 */
    float dt;           /* iter.dt */
    float pml_width;    /* used only when sigma == NULL */
    float pml_scale;    /* used only when sigma == NULL */
    const float *x;     /* HOST, nx values: dim.x (src/dims.jl:56-60) */
    const float *y;     /* HOST, ny_global values: dim.y */
    const float *sigma; /* HOST, optional nx values: dyn.pml[:,1]; NULL -> build_pml (src/pml.jl:21-29) from x */
    const float *grad8; /* HOST, optional 8 values: rows of dyn.grad = first(3) central(2) last(3);
                           NULL -> gradient(x) (src/operators.jl:10-22) */
    float d_omega;      /* get_dx(dim)*get_dy(dim) (src/env.jl:108); <= 0 -> computed from x, y */
    /* slab decomposition of ONE large grid along y (dim 2): this handle owns global rows
       [row0, row0+ny) of ny_global.  Single-device use: ny_global = ny, row0 = 0. */
    int32_t ny_global;
    int32_t row0;
    uint32_t flags;     /* reserved, 0 */
} waves_config;

int waves_version(void);
/* 0 for the release library.  WAVES_BUILD_DEV: a developer build (-DWAVES_DEV) that honours the WAVES_DEBUG_* environment
 * switches (tile sizes, launch order, skipped kernel variants); benchmarks must refuse it.  The release build reads no
 * environment variable at all. */
#define WAVES_BUILD_DEV 1
int waves_build_flags(void);
const char *waves_last_error(void);

/* AcousticDynamics + Integrator + WaveEnv state allocation (src/dynamics.jl:141-149, src/env.jl:52-66). */
int waves_create(const waves_config *cfg, waves_handle **out);
int waves_destroy(waves_handle *h);
int waves_sync(waves_handle *h);

/* env.wave[:,:,:,end] <- u12 / -> u12 : (12, ny, nx) floats (src/env.jl:102,116). env = -1: all envs, contiguous. */
int waves_set_state(waves_handle *h, int env, const float *u12);
int waves_get_state(waves_handle *h, int env, float *u12);

/* θ[2] = F: Source / RandomPosGaussianSource (src/sources.jl:10-23,25-69): f(t) = shape .* sin(2π t freq).
 * shape (ny, nx) plane; shape == NULL -> NoSource (src/sources.jl:7-8).
 * Slab handles (ny < ny_global) pass their rows INCLUDING the WAVES_HALO ghost rows that exist on each
 * interior side: global rows [row0 - ghosts_above, row0 + ny + ghosts_below). */
int waves_set_source(waves_handle *h, int env, const float *shape, float freq);

/* θ[1] = C: t -> speed(DesignInterpolator(initial, final, ti, tf)(t), grid, c0)
 * (src/env.jl:96-99, src/designs.jl:99-116, 274-292).  cyl0/cyl1: HOST (ncyl,4) rows {x, y, r, c}
 * of the stacked cylinders (Cloak: config then core, src/designs.jl:228) at ti and tf.
 * ncyl == 0 -> NoDesign: scalar c0 (src/designs.jl:63). */
int waves_set_design(waves_handle *h, int env, int ncyl, const float *cyl0, const float *cyl1, float ti, float tf);

/* Same for every environment of the handle in one call: cyl0 / cyl1 are HOST (n_env, ncyl, 4) blocks (one DesignInterpolator
 * per environment over a common [ti, tf]: a batch of WaveEnv stepping in lockstep). */
int waves_set_design_batch(waves_handle *h, int ncyl, const float *cyl0, const float *cyl1, float ti, float tf);

/* θ[1] = C: t -> a fixed speed plane c (ny, nx) (generic C closure frozen in time). NULL clears it.
 * Slab handles pass ghost rows too, like waves_set_source. */
int waves_set_speed_field(waves_handle *h, int env, const float *c);

/* dyn(x, t, θ) (src/dynamics.jl:179-188) on the current state of `env`: du12 (12, ny, nx). EXACT order. */
int waves_rhs(waves_handle *h, int env, float t, float *du12);

/* u <- u + runge_kutta(dyn, u, t, θ, dt) for every env (src/dynamics.jl:9-16, :40-42). */
int waves_step(waves_handle *h, float t, int mode);

/*
 * (iter::Integrator)(ui, tspan, θ) (src/dynamics.jl:37-53) fused with the energy metric of
 * (env::WaveEnv)(action) (src/env.jl:104-114), for every env of the handle.
 *   tspan        HOST, steps+1 times (build_tspan, src/dynamics.jl:5-7); step i uses tspan[i]
 *   energy       nullable, (n_env, steps+1, 3) floats {tot, inc, sc}; frame 0 is the initial state
 *   save_steps   HOST, nsave ascending step indices in [0, steps]; frames (n_env, nsave, 12, ny, nx)
 *                (env.wave = frames 80,90,100: src/env.jl:116)
 *   u_tot_traj / u_inc_traj  nullable (n_env, steps+1, ny, nx): the full U trajectories that
 *                (env::WaveEnv)(action) returns for rendering (src/env.jl:120, src/plot.jl:25)
 *                With waves_set_traj_stride(h, k) they hold every k-th frame only: (n_env, steps / k + 1, ny, nx), frame 0 first
 *                (what a renderer that shows a few frames per action needs, instead of 2 x 198 MB per action at 700^2).
 * The state of the handle is left at the last step.
 * Fused mode amortises launches (waves_set_graph): a batch that fits into one wave of warps (e.g. a single 700^2 environment,
 * which is latency-bound) takes one launch per step and one energy reduction per call; larger batches replay the whole
 * integration from a CUDA graph captured by the first call of the same shape (same steps, save steps, output buffers; frames /
 * trajectories in pageable host memory disable the graph: use device or pinned buffers).
 */
int waves_integrate(waves_handle *h, const float *tspan, int steps, int mode, float *energy, const int32_t *save_steps,
                    int nsave, float *frames, float *u_tot_traj, float *u_inc_traj);
/* Decimation of the U trajectories of waves_integrate (stride >= 1, default 1: every frame; src/plot.jl:25 consumers). */
int waves_set_traj_stride(waves_handle *h, int stride);
/* How waves_integrate issues the kernels of a fused integration (results are bitwise the same in every mode):
 *   0  every kernel of every step launched directly;
 *   1  (default) large batches replay the whole integration from a CUDA graph captured by the first call of a shape; a batch
 *      that fits into one wave of warps (a single 700^2 environment) takes ONE launch per step (all variants in one kernel);
 *   2  like 1, but such a small batch advances all steps between two saved frames in one COOPERATIVE launch with a grid
 *      barrier between steps (fewest launches; measured slower per step than mode 1 on a B200). */
int waves_set_graph(waves_handle *h, int on);

/*
 * rrule(::Integrator, z0, t, θ) + adjoint_sensitivity (src/dynamics.jl:97-128) for the 2-D dynamics: runs the forward
 * integration from the handle's current state, then the reverse sweep.  The loss is
 *     L = sum_i sum_k w_energy[i][k] * E_k(z_i)  +  <dL_dzN, z_N>          E = {tot, inc, sc} of src/env.jl:104-111
 *   w_energy   HOST, nullable, (steps+1, 3)           dL_dzN   nullable, (n_env, 12, ny, nx) extra cotangent of the last state
 *   dL_dz0     out, (n_env, 12, ny, nx)               dL_dc    out, nullable, (n_env, ny, nx): sensitivity to the speed
 *              plane of the total field (a time-constant perturbation, accumulated over every RK stage; the hard cylinder
 *              mask of src/designs.jl:99-104 has no derivative, so no design-parameter gradient exists in the reference)
 *   loss       out, nullable, HOST (n_env): the energy part of L
 *   fwd_mode   WAVES_MODE_FUSED / WAVES_MODE_EXACT for the forward pass;  adj_mode  WAVES_ADJ_EXACT / WAVES_ADJ_COMPAT
 * With dL_dc == NULL (the gradient the reference itself can produce) every reverse step is ONE fused launch set -- the
 * dynamics are affine in the state, so the reverse step is an RK4 step of the transposed operator and needs no forward stage
 * state -- and only the two U planes of each state are stored (for the energy cotangent).  With dL_dc the per-stage kernels
 * run on fully stored states.  Storage: the reference keeps all steps+1 states; here the forward pass keeps a checkpoint per
 * segment and the reverse sweep re-runs one segment at a time when steps+1 stored states do not fit in free device memory
 * (waves_set_adjoint_checkpoint forces a segment length).  The handle is left at the last state z_N.
 * Not available on slab handles.
 */
int waves_adjoint(waves_handle *h, const float *tspan, int steps, int fwd_mode, int adj_mode, const float *w_energy,
                  const float *dL_dzN, float *dL_dz0, float *dL_dc, float *loss);
/* Segment length of the reverse pass: a checkpoint every `every` steps, the steps of one segment re-run and stored while the
 * sweep crosses it.  0 (default): one segment when everything fits in free device memory, else about sqrt(steps * 6). */
int waves_set_adjoint_checkpoint(waves_handle *h, int every);

/*
 * RLBase.state(env) (src/env.jl:132-137): x = imresize(cat(env.wave[:, :, 1, :], env.source.shape; dims = 3), env.resolution).
 *   frames   the block waves_integrate filled, (n_env, nsave, 12, ny, nx), host or device: the U_tot plane of every saved frame
 *            is a channel; the handle's own source shape is the last channel
 *   out      (n_env, nsave + 1, res_y, res_x), host or device
 * imresize is Images.jl's (third party, version not pinned by the reference: no Manifest): bilinear sampling at
 * s (i - 1/2) + 1/2, no prefilter, restated from ImageTransformations.jl; needs res <= grid size like src/env.jl:52.
 */
int waves_observe(waves_handle *h, const float *frames, int nsave, int res_x, int res_y, float *out);

/* tot/inc/sc energy of the current state (src/env.jl:104-111): e3 (n_env, 3).  A DEVICE e3 is written asynchronously on the
 * handle's stream (no synchronisation: per-step energies of a slab run stay on the device until the end). */
int waves_energy(waves_handle *h, float *e3);

/* ---- slab decomposition (one process per GPU) ----------------------------------- */
/* Device pointers + sizes of the rows a neighbour needs / provides after every RK4 step:
 * send_lo/send_hi: the first/last WAVES_HALO owned rows; recv_lo/recv_hi: the ghost rows.
 * Each region is n_env*12 row blocks of WAVES_HALO*pitch floats, strided by plane_stride floats. */
typedef struct waves_halo_desc {
    float *send_lo, *send_hi, *recv_lo, *recv_hi; /* NULL where there is no neighbour */
    int64_t plane_stride;                          /* floats between consecutive field planes */
    int32_t n_planes;                              /* n_env*12 */
    int32_t block_floats;                          /* WAVES_HALO * pitch */
    int32_t pitch;                                 /* floats per row in device memory */
} waves_halo_desc;
int waves_halo_describe(waves_handle *h, waves_halo_desc *out);
/* Pack the send rows into two contiguous device buffers / unpack received ghost rows.
 * Buffers are n_planes*block_floats floats each (device). NULL skips that side.
 * The ghost rows must come from the neighbouring slab of the SAME global state (after waves_set_state on one rank every
 * rank must set its state too): the fused path caches fields that are constant where sigma is zero. */
int waves_halo_pack(waves_handle *h, float *lo_buf, float *hi_buf);
int waves_halo_unpack(waves_handle *h, const float *lo_buf, const float *hi_buf);
/* ---- halo without an exchange step: peer memory over NVLink --------------------------------------------------------
 * With one process per GPU, every slab handle exports CUDA IPC handles of its two state buffers and of a small flag array
 * (waves_peer_export); the caller ships that 256-byte record to the neighbouring ranks (any transport) and attaches the
 * neighbours' records (waves_peer_attach).  From then on the fused step of this rank stores the rows a neighbour needs
 * straight into that neighbour's ghost rows while it computes them, and consecutive steps of neighbouring ranks are ordered
 * by stream-ordered flags (a rank starts step n when both neighbours have finished step n-1): no pack, no NCCL call, no
 * unpack, no host synchronisation.  WAVES_MODE_FUSED only; every rank must take the same sequence of steps.
 */
typedef struct waves_peer_info {
    unsigned char ipc_u0[64], ipc_u1[64], ipc_flags[64]; /* cudaIpcMemHandle_t of u[0], u[1], flags */
    int32_t ny_alloc, ny_own0, ny_own, device;
    int32_t reserved[12];
} waves_peer_info;
int waves_peer_export(waves_handle *h, waves_peer_info *out);
/* lo / hi: the records of the ranks owning the rows below / above this slab (NULL at the ends of the grid) */
int waves_peer_attach(waves_handle *h, const waves_peer_info *lo, const waves_peer_info *hi);

/* CUDA stream (cudaStream_t) the handle launches on, so the caller can order NCCL calls with it. */
void *waves_stream(waves_handle *h);

/* ---- host-side builders restating reference constructors (used by the Python mirror and tests) --- */
/* collect(range(start, stop, n)) for Float32 endpoints, src/dims.jl:56-60 (Float64 interpolation, one rounding). */
int waves_range_f32(float start, float stop, int n, float *out);
/* build_pml(::TwoDim) 1-D profile, src/pml.jl:21-29 */
int waves_build_pml_profile(const float *x, int n, float width, float scale, float *out);
/* gradient(x) rows first(3) central(2) last(3), src/operators.jl:10-22 */
int waves_build_gradient8(const float *x, int n, float *out8);
/* build_normal(grid, mu, sigma, a), src/utils.jl:12-18: out (ny, nx) */
int waves_build_normal(const float *x, int nx, const float *y, int ny, int n, const float *mu_xy, const float *sigma,
                       const float *a, float *out);
/* get_dx(dim) = mean(diff(x)), src/dims.jl:126 */
float waves_mean_diff(const float *x, int n);

/* ---- 1-D latent dynamics: the training-time loop of AcousticEnergyModel -------------------------
 * Replaces, for AcousticDynamics{OneDim} (src/dynamics.jl:190-222),
 *     model.iter(z0, t, θ)                with θ = [C, F, PML]   src/model/acoustic_energy_model.jl:86-99
 *     compute_latent_energy(z, dx)                               src/model/acoustic_energy_model.jl:6-15
 *     rrule(::Integrator, z0, t, θ) / adjoint_sensitivity        src/dynamics.jl:97-128
 * One CTA integrates one batch element entirely in shared memory (no launch per step).  Arrays keep the reference's
 * column-major memory image (host or device pointers):
 *     z0 (n, 4, batch) fields U_tot, V_tot, U_inc, V_inc      tspan (steps+1, batch)   (one time column per sample)
 *     X (nseq, batch), Y (n, nseq, batch): C = LinearInterpolation(X, Y) (src/utils.jl:88-98), queried at every RK stage
 *     shape (n, batch), freq: F = Source(shape, freq) (src/sources.jl:10-23); shape == NULL -> no source
 *     pml (n, batch): θ[3]; σ = dyn.pml[[1]] .* pml
 *     z (n, 4, batch, steps+1)     energy (steps+1, 3, batch) {tot, inc, sc}
 * Same conventions as above: 0 / non-zero return codes with waves_last_error(), no CPU path, a handle is not thread-safe,
 * every call synchronises the handle's stream before it returns (all outputs are complete on return).
 */
typedef struct waves_latent waves_latent;
typedef struct waves_latent_config {
    int32_t n;        /* grid points of the OneDim latent grid (reference: 1024) */
    int32_t device;   /* CUDA device ordinal */
    float c0;         /* dyn.c0 */
    float dt;         /* iter.dt */
    float pml_width;  /* AcousticDynamics(dim, c0, pml_width, pml_scale): used when pml0 < 0 */
    float pml_scale;
    float pml0;       /* >= 0: dyn.pml[1] given directly; < 0: computed with build_pml(::OneDim) (src/pml.jl:6-15) from x */
    float dx;         /* get_dx(latent_dim) (src/dims.jl:126); <= 0: computed from x */
    const float *x;     /* HOST, n values: latent_dim.x (src/dims.jl:48-50); may be NULL when grad8, pml0 and dx are given */
    const float *grad8; /* HOST, optional: rows of dyn.grad first(3) central(2) last(3); NULL -> gradient(x) */
} waves_latent_config;
int waves_latent_create(const waves_latent_config *cfg, waves_latent **out);
int waves_latent_destroy(waves_latent *h);
/* z = iter(z0, tspan, θ) and/or compute_latent_energy(z, dx); z, energy, z_last (the last state, (n, 4, batch)) are each
 * nullable: with z == NULL no trajectory ever reaches HBM. */
int waves_latent_integrate(waves_latent *h, int batch, int steps, int nseq, const float *z0, const float *tspan,
                           const float *X, const float *Y, const float *shape, float freq, const float *pml, float *z,
                           float *energy, float *z_last);
/* adjoint_sensitivity(iter, z, t, θ, ∂L_∂z) (src/dynamics.jl:97-118).  z is the stored solution of waves_latent_integrate.
 * The cotangent of the solution is  dL_dz (n, 4, batch, steps+1, nullable)  plus, when w_energy (steps+1, 3, batch) is
 * given, the pullback of compute_latent_energy (so a loss on the energies needs no (n,4,batch,time) cotangent at all).
 * adj_mode: WAVES_ADJ_EXACT (discrete adjoint) or WAVES_ADJ_COMPAT (the reference loop as written).
 * Outputs: dL_dz0 (n, 4, batch); nullable dL_dY (n, nseq, batch), dL_dshape (n, batch), dL_dpml (n, batch). */
int waves_latent_adjoint(waves_latent *h, int batch, int steps, int nseq, const float *z, const float *tspan, const float *X,
                         const float *Y, const float *shape, float freq, const float *pml, int adj_mode,
                         const float *w_energy, const float *dL_dz, float *dL_dz0, float *dL_dY, float *dL_dshape,
                         float *dL_dpml);
/* build_pml(::OneDim, width, scale), src/pml.jl:6-15: out (n, nullable), *first = its first value (nullable). HOST. */
int waves_latent_build_pml(const float *x, int n, float width, float scale, float *out, float *first);
int64_t waves_latent_launch_count(waves_latent *h);
/* on != 0: always run the generic shared-memory kernels (any n that fits) instead of the register fast path taken for
 * n <= 1024 and nseq <= 64; both are bit-identical, the switch exists so tests can compare them */
int waves_latent_set_generic(waves_latent *h, int on);
/* Kernel variants of the latent path (flags, OR-ed).  Every variant is green on the B200 (tests/test_gpu_zlatent.py) and timed
 * (profiles/r2_latent_bench.jsonl).  AUTO, the default, takes the fastest measured kernel that applies:
 *   forward  n <= 1024, nseq <= 64: two elements per thread, packed f32x2 (n even; 2.9 us per RK4 step at 148 x 1024), else one
 *            element per thread (3.0 us); otherwise the generic shared-memory kernel.  All three are bit-identical.
 *   reverse  same conditions: the register kernels (pair form 7.6 us per step, single 9.4 us); otherwise the generic kernel (15.5).
 * GENERIC (exclusive): as waves_latent_set_generic(h, 1).  The other flags select kernels explicitly (tests, benchmarks):
 * SINGLE: one element per thread in the forward pass, generic reverse kernel unless ADJ_R1 is also given;
 * PAIR: forward pair kernel; ADJ_R1: register reverse kernel, in its pair form with PAIR | ADJ_R1. */
#define WAVES_LATENT_AUTO 0
#define WAVES_LATENT_GENERIC 1
#define WAVES_LATENT_PAIR 2
#define WAVES_LATENT_ADJ_R1 4
#define WAVES_LATENT_SINGLE 8
int waves_latent_set_variant(waves_latent *h, int variant);
/* device time (ms, CUDA events on the handle's stream) of the kernel of the last integrate / adjoint call */
float waves_latent_last_kernel_ms(waves_latent *h);

/* ---- introspection for benchmarks -------------------------------------------------------- */
/* number of kernel launches issued by this handle so far */
int64_t waves_launch_count(waves_handle *h);
/* number of waves_integrate calls that ran as ONE CUDA graph launch (capture + first launch included) */
int64_t waves_graph_replays(waves_handle *h);
/* average device time (ms) of the fused step kernel launches since the last reset, measured with
 * CUDA events on the handle's stream when profiling is on (waves_profile(h, 1)). */
int waves_profile(waves_handle *h, int on);
int waves_profile_read(waves_handle *h, double *fused_ms_total, int64_t *fused_launches);

#ifdef __cplusplus
}
#endif
#endif /* WAVES_B200_H */

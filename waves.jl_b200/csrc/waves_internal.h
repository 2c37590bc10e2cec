// Internal declarations shared by the C-ABI host code and the CUDA kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/waves_b200.h"

// Developer tuning / bisecting switches exist only in -DWAVES_DEV builds (scripts/tune_build.sh).  The release library reads no
// environment variable: a benchmarked .so cannot be made to skip work from outside (waves_build_flags() reports which build it is).
#ifdef WAVES_DEV
#include <stdlib.h>
static inline int waves_dev_env(const char *name, int dflt) {
    const char *s = getenv(name);
    return s ? atoi(s) : dflt;
}
#else
#define waves_dev_env(name, dflt) (dflt)
#endif

// Per-environment parameters of θ = [C, F] (src/env.jl:96-102), device-resident.
struct EnvParams {
    int ncyl;        // 0 -> NoDesign
    int has_source;  // 0 -> NoSource
    int has_cplane;  // 1 -> fixed speed plane overrides the design
    float ti, tf;    // DesignInterpolator window (src/designs.jl:274-279)
    float freq;      // source frequency (src/sources.jl:12,35)
    int src_j0, src_j1;  // local row range [j0, j1) where the source shape is non-zero
    int src_i0, src_i1;  // column range where the source shape is non-zero
};

// Geometry + constants passed by value to every kernel.
struct GridP {
    int nx;         // grid points along x (contiguous)
    int nxp;        // row pitch in floats (nx rounded up to a multiple of 4)
    int ny_alloc;   // rows held in device memory (owned + ghost)
    int ny_own0;    // first owned local row (== number of ghost rows on top)
    int ny_own;     // owned rows
    int ny_global;  // rows of the whole grid
    int grow0;      // global row index of local row 0
    int n_env;
    long long plane;       // floats per field plane  (ny_alloc * nxp)
    long long env_stride;  // floats per environment  (12 * plane)
    float g_first[3], g_central[2], g_last[3];  // rows of the gradient matrix (src/operators.jl:10-22)
    float c0, b0;   // ambient speed and c0*c0
    float dt, hdt;  // dt and 0.5f*dt
    const float *x, *y, *sigma;  // device: dim.x (nx), dim.y (ny_global), pml profile (max(nx, ny_global))
};

// One row of the per-(env, step) stage table: times and source factors of the RK4 stages.
// [0..2] = t, t+dt/2, t+dt ; [3..5] = sin(2π t f) at those times ; [6..7] pad
#define STAGE_ROW 8

struct waves_handle {
    GridP gp;
    int device;
    cudaStream_t stream;
    float *u[2];  // ping-pong state [n_env][12][ny_alloc][nxp]
    int cur;
    int aux_synced;       // 1: a fused step ran since the state was last written from outside: both buffers hold the constant fields
    float *k, *ys, *acc;  // exact-mode scratch, lazily allocated
    float *b2;            // exact-mode c^2 plane per env [n_env][plane], lazily allocated
    float *shape;         // [n_env][plane] source shape (zeros when NoSource)
    float *cplane;        // [n_env][plane] fixed speed plane, lazily allocated
    float *pconst;        // [n_env][2][plane]: Psix + Psiy - Omega of the interior cells of both wavefields (kernels_fused.cu)
    float *adj[9];        // adjoint scratch: w, wsum, lk, ly, y1, y2, y3 (state-sized), b2 planes x3 (one block), dL/dc
    float *obs_in, *obs_out;  // observation scratch (waves_observe)
    size_t obs_in_cap, obs_out_cap;
    float *traj;          // reverse pass: what the reverse steps of one segment need of the stored states (waves_adjoint)
    size_t traj_cap;      // floats
    float *ckpt;          // reverse pass: the state at the start of every segment + the last state
    size_t ckpt_cap;      // floats
    int adj_ckpt;         // steps per segment requested by the caller (0: automatic)
    int adj_last_K;       // steps per segment the last waves_adjoint call used
    int adj_z0, adj_z1;   // zero-sigma zone of the profile (fused reverse step), 0/0: not computed yet
    float *adj_g;         // [n_env][2][plane] accumulated auxiliary cotangent of the interior cells (reverse pass on the march)
    int adj_march;        // waves_set_adjoint_march: 1 (default) the interior of the fused reverse step runs on the march kernel
    float *d_x, *d_y, *d_sigma;
    EnvParams *h_env, *d_env;
    bool env_dirty;
    float *d_cyl0, *d_cyl1;  // [n_env][cyl_cap][4]
    float *h_cyl0, *h_cyl1;
    int cyl_cap;
    bool cyl_dirty;
    float *d_tspan;
    float *d_stage;
    int stage_cap;  // steps capacity of d_tspan / d_stage
    double *d_epart;
    int epart_blocks;
    float *d_energy;
    int energy_cap;  // frames capacity
    float d_omega;
    CUtensorMap map_u[2];   // TMA descriptors over the two state buffers, box = 64 x 1 x 7 planes (total field + U_inc)
    CUtensorMap map_u6[2];  // same tensors, box = 64 x 1 x 6 planes (one wavefield)
    CUtensorMap map_u3[2];  // box = 64 x 1 x 3 planes (U, Vx, Vy of one wavefield: the lean interior variant)
    CUtensorMap map_u1[2];  // box = 64 x 1 x 1 plane
    CUtensorMap map_p;      // the P planes, box = 64 x 1 x 1
    CUtensorMap map_shape;
    bool maps_ready;
    int64_t launches;
    int profile;
    double fused_ms;
    int64_t fused_launches;
    cudaEvent_t ev0, ev1;
    cudaEvent_t ev_params;  // recorded after every upload of h_env / h_cyl*: the (pinned) staging may be rewritten once it completed
    int fused_smem;
    int sm_count;
    void *plan;  // FusedPlan (kernels_fused.cu)
    // ---- peer halo (slabs without an exchange step): neighbours' state buffers mapped through CUDA IPC ----
    int peer_on;
    float *peer_u[2][2];      // [side][buffer]: neighbour's u[0], u[1] (nullptr: no neighbour)
    unsigned peer_plane[2];
    int peer_j0[2], peer_dj[2];
    int *flags;               // [2] device ints on THIS GPU: steps completed by the lower / upper neighbour
    int *peer_flag[2];        // the slot in neighbour s's flags that this rank writes
    void *ipc_ptr[2][3];      // opened IPC mappings (to close)
    int peer_steps;           // fused steps taken in peer mode
    // ---- CUDA graph of a whole waves_integrate call (waves_abi.cu) ----
    void *graph_exec;         // cudaGraphExec_t of the last captured integration, nullptr: none
    uint64_t graph_key[12];   // what the captured graph depends on (shape of the call, buffer addresses, ping-pong parity)
    int64_t graph_launches;   // kernel launches one replay stands for
    int64_t graph_replays;
    int graph_cur_after;      // h->cur after the captured run
    int graph_off;            // 1: launch every kernel directly (waves_set_graph)
    int coop_on;              // 1: small batches take several steps per cooperative launch (waves_set_graph(h, 2))
    int traj_stride;          // U trajectories keep every traj_stride-th frame (waves_set_traj_stride); 0 == 1
};

// ---- kernels_exact.cu ----
void launch_stage_table(waves_handle *h, const float *d_tspan, int steps, float *d_table);
void launch_speed2(waves_handle *h, int env0, int nenv, const float *d_table, int steps, int step, int stage,
                   float *b2);
void launch_rhs_exact(waves_handle *h, int env0, int nenv, const float *u_in, float *k_out, const float *d_table,
                      int steps, int step, int stage, const float *zadd = nullptr, float za = 0.0f);
void launch_rk_update(waves_handle *h, int stage, const float *u, const float *k, float *acc, float *ys);
void launch_rk_final(waves_handle *h, const float *u_in, const float *acc, float *u_out);
void launch_energy(waves_handle *h, const float *u, float *d_e3, int frame_stride3);
void launch_imresize(waves_handle *h, const float *in, long long env_stride, long long chan_stride, int n_chan, int in_pitch,
                     const float *last, long long last_env_stride, int rx, int ry, float *out);
void launch_pack_halo(waves_handle *h, const float *u, float *lo, float *hi);
void launch_copy_blocks(waves_handle *h, const float *src, float *dst, long long elems, long long src_stride, long long dst_stride,
                        int n_blocks);  // device -> device, everything a multiple of 4 floats
void launch_unpack_halo(waves_handle *h, float *u, const float *lo, const float *hi);

// ---- kernels_adjoint.cu ----
void launch_rhs_transposed(waves_handle *h, float a, const float *w, float b, const float *lyp, const float *y, const float *b2,
                           float *out, float *ws, int first, float *gcacc);
void launch_lin3(waves_handle *h, float *out, float a, const float *x, float b, const float *y, float c, const float *z);
void launch_energy_cotangent(waves_handle *h, const float *z, long long z_env, long long z_inc, float *w, const float *w3 /*host, 3 weights*/);

// ---- kernels_adjoint_fused.cu ----
int launch_adjoint_step_fused(waves_handle *h, const float *w_in, float *w_out, const float *const b2[3], const float *pre_u,
                              const float *pre_w3, const float *post_u, const float *post_w3);
void launch_gather_u(waves_handle *h, const float *u, float *out);
int launch_adjoint_step_frame(waves_handle *h, const float *w_in, float *w_out, const float *const b2[3], const float *pre_u,
                              const float *pre_w3, const float *post_u, const float *post_w3, const int rect[4]);
void launch_apply_aux_cotangent(waves_handle *h, float *w, const float *G, const int rect[4]);

// ---- kernels_fused.cu ----
int fused_prepare(waves_handle *h);  // work items, tensor maps, smem attribute; 0 on success
void fused_release(waves_handle *h);
int fused_item_counts(waves_handle *h, int *n_int, int *n_gen);
void fused_profile_drain(waves_handle *h);   // waves_profile: collect the recorded launch-set durations (synchronises)
int source_bbox(waves_handle *h, int env, int *bbox4);
int waves_set_error(const char *msg);  // sets the thread-local message, returns 1
int launch_fused_step(waves_handle *h, const float *d_table, int steps, int step, float *d_e3 /*nullable*/, int defer_slot = -1);
int launch_fused_multi(waves_handle *h, const float *d_table, int steps, int step0, int count, bool energy);  // 0 done, -1 n/a, 1 error
int launch_adjoint_interior_march(waves_handle *h, const float *w_in, float *w_out, float *G, const float *zbase, size_t zfloats,
                                  const float *z, const float *w3, int inj, const float *d_table, int steps, int step, int rect[4]);
int launch_adjoint_ring_march(waves_handle *h, const float *w_in, float *w_out, const float *zbase, size_t zfloats, const float *z,
                              const float *w3, int inj, const float *d_table, int steps, int step, int outer[4]);
void adjoint_march_join(waves_handle *h, int ring);   // the handle's stream waits for the march kernels of the current reverse step
bool fused_is_small_batch(waves_handle *h);
int fused_epart_slots(waves_handle *h);
void fused_reduce_deferred(waves_handle *h, int count, float *d_e3, int env_stride3);

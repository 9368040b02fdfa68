/* flipgpu.h -- C ABI of the B200-native grid-fluid step (libflipgpu.so).
 *
 * Drop-in boundary for ONE hot path of owenbharrison/simsAndGames: the per-sim
 * state + step surface that the reference UI shells touch,
 *   struct FlipFluid   flip_fluid/src/flip_fluid.h:13-582   (caller: flip_fluid/src/main.cpp:32-168)
 *   class  Fluid       eulerian_fluid/src/fluid.h:19-260     (caller: eulerian_fluid/src/fluid_ui.h:64-111)
 * The reference has no FFI of its own; these entry points are what a
 * binding for that surface needs: create / upload / step / readback / destroy.
 *
 * Conventions
 *  - plain C, opaque handles, host pointers + element counts, no torch/CUDA types.
 *  - every call returns 0 on success, a FLIPGPU_E* code otherwise;
 *    flipgpu_last_error() gives the message (thread-local).
 *  - a handle is not thread-safe (the reference is single-threaded).
 *  - arrays cross the boundary in the REFERENCE layout and index order:
 *    FLIP grid x-fastest  fIX(i,j)=i+f_num_x*j (flip_fluid.h:147-149), particles AoS xy / rgb
 *    in the caller's particle index order; Eulerian grid y-fastest ix(i,j)=i*num_y+j (fluid.h:94-96).
 *  - undefined reference state (flip_fluid.h:76-110 never initialises its arrays) is
 *    defined as all-zero here (particle colour blue = 1 as flip_fluid.h:94-96).
 *  - there is no CPU fallback: without a CUDA device every create call fails.
 */
#ifndef FLIPGPU_H
#define FLIPGPU_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FLIPGPU_OK 0
#define FLIPGPU_EINVAL 1   /* bad argument / bad handle / count mismatch */
#define FLIPGPU_ECUDA 2    /* a CUDA runtime call failed */
#define FLIPGPU_ENODEV 3   /* no usable CUDA device (no CPU fallback exists) */
#define FLIPGPU_ENCCL 4    /* an NCCL call failed */
#define FLIPGPU_ESTATE 5   /* call not valid in this state (e.g. sharded-only call on a 1-GPU handle) */

const char* flipgpu_last_error(void);
int flipgpu_version(void);
int flipgpu_device_count(int* count);

/* ------------------------------------------------------------------ FLIP (flip_fluid.h) */
typedef struct FlipSim FlipSim;

/* constructor arguments, FlipFluid(d,width,height,sp,r,m) flip_fluid.h:62-65 */
typedef struct {
	float density, width, height, spacing, particle_radius;
	int max_particles;
	int device; /* CUDA ordinal; -1 = current device */
} FlipParams;

/* derived sizes, computed on the host in fp32 exactly as flip_fluid.h:69-74,103-106 */
typedef struct {
	int f_num_x, f_num_y, f_num_cells;
	int p_num_x, p_num_y, p_num_cells;
	int max_particles, num_particles;
	float density, h, f_inv_spacing, particle_radius, p_inv_spacing;
} FlipDims;

/* state arrays of struct FlipFluid (flip_fluid.h:23-55), addressable for upload/readback */
typedef enum {
	FLIP_U = 0, FLIP_V, FLIP_DU, FLIP_DV, FLIP_PREV_U, FLIP_PREV_V, FLIP_P, FLIP_S, /* float[C] */
	FLIP_CELL_TYPE,          /* int[C]: 0 fluid, 1 air, 2 solid (flip_fluid.h:30-34) */
	FLIP_CELL_COLOR,         /* float[3C] */
	FLIP_PARTICLE_POS,       /* float[2*num_particles] */
	FLIP_PARTICLE_VEL,       /* float[2*num_particles] */
	FLIP_PARTICLE_COLOR,     /* float[3*num_particles] */
	FLIP_PARTICLE_DENSITY,   /* float[C] */
	FLIP_NUM_CELL_PARTICLES, /* int[p_num_cells] */
	FLIP_FIRST_CELL_PARTICLE,/* int[p_num_cells+1] */
	FLIP_CELL_PARTICLE_IDS,  /* int[num_particles] */
	FLIP_FIELD_COUNT
} FlipField;

typedef enum {
	FLIP_SOLVER_LEX_WAVEFRONT = 0, /* the reference's lexicographic loop order (flip_fluid.h:458-494) as a blocked anti-diagonal
	                                  wavefront: skewed square tiles with temporal blocking in shared memory; bit-identical
	                                  u,v,p.  Value 0: a zero-initialised FlipStepParams runs the reference's own order */
	FLIP_SOLVER_RED_BLACK = 1,     /* red-black order (north_star's named order); bounded against the reference order (T2).
	                                  At the reference's 50 sweeps it is less stable than the reference order on tall
	                                  liquid columns (profiles/r1_stability.md) */
	FLIP_SOLVER_LEX_DIAGONAL = 2,  /* same order, unblocked one-diagonal-per-grid-sync kernel (cross-check / non-binary s) */
	FLIP_SOLVER_LEX_SYSTOLIC = 3   /* same order, register-systolic pipeline (one warp per 32-column strip, the iterations of a
	                                  chunk as pipeline stages, no block-wide barrier; sor_systolic.cu): bit-identical, slower
	                                  than the tiled kernel on B200 so far (profiles/r2_systolic.md) */
} FlipSolverOrder;

/* the 13 arguments of FlipFluid::simulate (flip_fluid.h:560-565) + two switches */
typedef struct {
	float dt, gravity, flip_ratio;
	int num_pressure_iters, num_particle_iters;
	float over_relaxation;
	int compensate_drift, separate_particles;
	float obstacle_x, obstacle_y, obstacle_vel_x, obstacle_vel_y, obstacle_radius;
	int solver_order;   /* FlipSolverOrder */
	int update_colors;  /* run updateParticleColors/updateCellColors + colour diffusion (flip_fluid.h:239-245,498-557) */
} FlipStepParams;

typedef enum {
	FLIP_SCALAR_REST_DENSITY = 0, /* particle_rest_density (flip_fluid.h:44,321-332) */
	FLIP_SCALAR_RESIDUAL_MAX,     /* max |div| over cells the solver updates, incl. drift term */
	FLIP_SCALAR_RESIDUAL_RMS
} FlipScalar;

/* phases of simulate() for flip_get_timings, ms accumulated since flip_reset_timings */
enum { FLIP_PH_INTEGRATE_BIN = 0, FLIP_PH_SEPARATE, FLIP_PH_COLLIDE_MARK, FLIP_PH_P2G, FLIP_PH_SOLVE,
       FLIP_PH_G2P, FLIP_PH_COLORS, FLIP_PH_COUNT };

int flip_create(const FlipParams* params, FlipSim** out);
int flip_destroy(FlipSim* sim);
int flip_get_dims(FlipSim* sim, FlipDims* dims);
/* num_particles is a public member the caller sets (flip_fluid/src/main.cpp:59) */
int flip_set_num_particles(FlipSim* sim, int n);
/* host -> device, reference layout; count = number of elements (not bytes) and must match the field */
int flip_upload(FlipSim* sim, int field, const void* host, size_t count);
/* the per-frame form (flip_fluid/src/main.cpp:106-122 rewrites s, u, v every drag frame): the copy is only QUEUED on the handle's
 * stream.  `host` must stay unchanged until the next synchronising call (flip_synchronize, flip_readback, flip_readback_wait),
 * unless it is pageable memory, which the CUDA runtime stages before it returns.  Whether an uploaded s is 0/1-valued (the
 * blocked solvers need to know) is checked on the device behind the copy. */
int flip_upload_async(FlipSim* sim, int field, const void* host, size_t count);
/* device -> host, reference layout and particle index order */
int flip_readback(FlipSim* sim, int field, void* host, size_t count);
/* particle_pos (caller order) -> pinned host memory without blocking: the PCIe copy runs on its own stream while the next
 * flip_step computes; flip_readback_wait joins it.  This is the frame loop of main.cpp:160-182 with the draw of frame k
 * overlapping the simulation of frame k+1. */
int flip_readback_pos_async(FlipSim* sim, void* host, size_t count);
int flip_readback_wait(FlipSim* sim);
/* device equivalent of FluidUI::setObstacle (flip_fluid/src/main.cpp:95-126): rewrites s (and u,v inside
 * the disc) over 1<=i<nx-2, 1<=j<ny-2 (Q10) with the given obstacle velocity */
int flip_set_obstacle(FlipSim* sim, float x, float y, float vx, float vy, float radius);
/* n_steps x FlipFluid::simulate (flip_fluid.h:560-581); asynchronous on the handle's stream */
int flip_step(FlipSim* sim, const FlipStepParams* sp, int n_steps);
/* flip_step replays a captured whole-step CUDA graph once the simulation is in steady state (same parameters for two steps,
 * rest density latched, reference solver order, no phase timers, nothing pending from the host); results are identical to the
 * launch-by-launch path, the gain is launch overhead on small scenes.  on = 0 switches the replay off. */
int flip_set_graph_mode(FlipSim* sim, int on);
int flip_get_graph_replays(FlipSim* sim, long long* replays);
/* single phases, for same-state replay tests (each cites its reference lines in flipgpu.cu) */
int flip_phase_integrate(FlipSim* sim, float dt, float gravity);                    /* :156-162 */
int flip_phase_bin(FlipSim* sim);                                                    /* :169-198 */
int flip_phase_separate(FlipSim* sim, int num_iters, int update_colors);             /* :201-250, parallel schedule */
int flip_phase_collide(FlipSim* sim, float ox, float oy, float ovx, float ovy, float orad); /* :254-289 */
int flip_phase_p2g(FlipSim* sim);                                                    /* :339-403,432-446 + :292-333 */
int flip_phase_solve(FlipSim* sim, int iters, float dt, float omega, int drift, int order); /* :451-495 */
int flip_phase_g2p(FlipSim* sim, float flip_ratio);                                  /* :404-429 */
int flip_phase_colors(FlipSim* sim);                                                 /* :498-557 */
/* Reference-order ("exact") mode for parity runs on small scenes: the separation sweeps run in the reference's serial
 * particle order (level-scheduled, flip_fluid.h:203-250 incl. Q1), the P2G adds of every node run in caller order
 * (:398-403, :315-318) and the rest density is the serial running sum of :321-332.  Together with
 * FLIP_SOLVER_LEX_WAVEFRONT a whole flip_step is then bit-identical to FlipFluid::simulate.  Much slower: hundreds of dependent
 * levels per sweep, and the level schedule is copied to the host every step (a parity mode for small scenes, single GPU only). */
int flip_set_exact_order(FlipSim* sim, int on);
/* levels of the last serial-order separation schedule; drifted != 0 if a particle left the 1-cell neighbourhood of its
 * bin during the sweeps (the schedule's validity assumption; never seen on the reference scenes) */
int flip_get_exact_order_info(FlipSim* sim, int* levels, int* drifted);
int flip_get_scalar(FlipSim* sim, int which, float* out);
int flip_set_rest_density(FlipSim* sim, float rest);
int flip_synchronize(FlipSim* sim);
/* the CUDA stream all work of this handle is issued on (a cudaStream_t), for event timing by the caller */
int flip_get_stream(FlipSim* sim, void** stream);
int flip_get_timings(FlipSim* sim, float* ms, int count);
int flip_reset_timings(FlipSim* sim);
int flip_enable_timings(FlipSim* sim, int on);
/* kernels launched by this handle since creation (all hand-written sm_100a kernels of this library) */
int flip_get_launch_count(FlipSim* sim, long long* launches);
/* diagnostic (FLIPGPU_SOR_PROFILE=1): SM cycles per phase of the blocked solver: wait, load, compute, write-back, tiles, n_tiles, Kc, chunks */
int flip_debug_solver_profile(FlipSim* sim, unsigned long long* out8);

/* scene helper = the particle lattice + tank walls of flip_fluid/src/main.cpp:47-77 written into host
 * buffers (pos: 2*max floats, s: nx*ny floats); returns the particle count through *num_particles.
 * Pass pos==NULL to only query the count. */
int flip_scene_tank(float tank_width, float tank_height, float spacing, float radius, float rel_w, float rel_h,
                    int f_num_x, int f_num_y, float* pos, size_t pos_capacity, float* s, int* num_particles);

/* ------------------------------------------------------------------ Eulerian (fluid.h) */
typedef struct EulerSim EulerSim;
typedef enum {
	EULER_U = 0, EULER_V, EULER_NEW_U, EULER_NEW_V, EULER_PRESSURE, EULER_M, EULER_NEW_M, /* float[C] */
	EULER_SOLID,  /* unsigned char[C] (the reference's bool[]) */
	EULER_FIELD_COUNT
} EulerField;
typedef struct { int num_x, num_y, num_cells; float density, h, h1; } EulerDims;

int euler_create(int x, int y, float density, float h, int device, EulerSim** out);  /* Fluid(x,y,d,h) fluid.h:41 */
int euler_destroy(EulerSim* sim);
int euler_get_dims(EulerSim* sim, EulerDims* dims);
int euler_upload(EulerSim* sim, int field, const void* host, size_t count);
int euler_readback(EulerSim* sim, int field, void* host, size_t count);
int euler_project(EulerSim* sim, int iters, float dt, int order);  /* solveIncompressibility fluid.h:98-139 */
int euler_extrapolate(EulerSim* sim);                              /* fluid.h:142-151 */
int euler_advect_vel(EulerSim* sim, float dt);                     /* fluid.h:210-239 */
int euler_advect_smoke(EulerSim* sim, float dt);                   /* fluid.h:241-259 */
/* n_steps x (project, extrapolate, advectVel, advectSmoke) = eulerian_fluid/src/fluid_ui.h:107-111 */
int euler_step(EulerSim* sim, int iters, float dt, int order, int n_steps);
/* device FluidUI::setObstacle, eulerian_fluid/src/fluid_ui.h:24-62: clears `solid`, rebuilds the left/top/bottom walls and
 * a solid disc of radius r cells centred on cell (x, y) whose cells get m = 1, u = vx, v = vy.  The caller supplies
 * (vx, vy) = (x - previous x, y - previous y) / dt as :39-43 does when reset == false, else 0. */
int euler_set_obstacle(EulerSim* sim, int x, int y, int r, float vx, float vy);
/* sampleField fluid.h:153-188 at n query points (field: 0 U, 1 V, 2 S) */
int euler_sample_field(EulerSim* sim, int field, const float* xs, const float* ys, float* out, size_t n);
int euler_get_residual(EulerSim* sim, float* max_div, float* rms_div);
int euler_synchronize(EulerSim* sim);
int euler_get_stream(EulerSim* sim, void** stream);
int euler_get_launch_count(EulerSim* sim, long long* launches);

/* ------------------------------------------------------------------ 3D Eulerian smoke (fluid3d.h)
 * struct Fluid3D, eulerian_fluid/src/fluid3d.h:19-368 (the 3D transcription of fluid.h; no reference target compiles it).
 * Same state arrays and calls; x-fastest layout ix(i,j,k) = i + num_x*j + num_y*num_x*k (fluid3d.h:131-133) on the wire. */
typedef struct Fluid3DSim Fluid3DSim;
typedef enum {
	FLUID3D_U = 0, FLUID3D_V, FLUID3D_W, FLUID3D_NEW_U, FLUID3D_NEW_V, FLUID3D_NEW_W, FLUID3D_PRESSURE, FLUID3D_M, FLUID3D_NEW_M, /* float[C] */
	FLUID3D_SOLID,  /* unsigned char[C] (the reference's bool[]) */
	FLUID3D_FIELD_COUNT
} Fluid3DField;
typedef struct { int num_x, num_y, num_z, num_cells; float density, h; } Fluid3DDims;

int fluid3d_create(int x, int y, int z, float density, float h, int device, Fluid3DSim** out);  /* Fluid3D(x,y,z,d,h) fluid3d.h:40 */
int fluid3d_destroy(Fluid3DSim* sim);
int fluid3d_get_dims(Fluid3DSim* sim, Fluid3DDims* dims);
int fluid3d_upload(Fluid3DSim* sim, int field, const void* host, size_t count);
int fluid3d_readback(Fluid3DSim* sim, int field, void* host, size_t count);
/* solveIncompressibility fluid3d.h:135-182; order: FLIP_SOLVER_LEX_WAVEFRONT (the reference's loop order as a plane wavefront,
 * bit-identical) or FLIP_SOLVER_RED_BLACK (parity of i+j+k) */
int fluid3d_project(Fluid3DSim* sim, int iters, float dt, int order);
int fluid3d_extrapolate(Fluid3DSim* sim);                 /* fluid3d.h:185-211 */
int fluid3d_advect_vel(Fluid3DSim* sim, float dt);        /* fluid3d.h:305-346, incl. its k < num_y loop bound (:311) */
int fluid3d_advect_smoke(Fluid3DSim* sim, float dt);      /* fluid3d.h:348-367 */
int fluid3d_step(Fluid3DSim* sim, int iters, float dt, int order, int n_steps);   /* project, extrapolate, advectVel, advectSmoke */
/* sampleField fluid3d.h:213-264 at n query points (field: 0 U, 1 V, 2 W, 3 S) */
int fluid3d_sample_field(Fluid3DSim* sim, int field, const float* xs, const float* ys, const float* zs, float* out, size_t n);
int fluid3d_synchronize(Fluid3DSim* sim);
int fluid3d_get_launch_count(Fluid3DSim* sim, long long* launches);

/* ------------------------------------------------------------------ multi-GPU: slab-sharded FLIP step
 * One process per GPU.  The grid is cut into nranks slabs of rows (y); rank r owns rows [j0,j1) and the particles
 * whose cell lies in them (ownership by position; global_particles = the number of caller indices in use).
 * Every flip_step: one NCCL neighbour exchange of ghosts (the owned particles within `ghost_band_rows` of a
 * boundary; the band follows the fastest particle, so a particle that crosses a boundary is always already present
 * on its new owner), the single-GPU kernels on owned + ghost particles, halo-row
 * exchanges of the grid fields after P2G and after the solve, and the red-black projection with the
 * communication-avoiding deep halo (halo_rows rows every halo_rows-1 colour passes).  Positions stay in GLOBAL
 * coordinates and particles carry their global caller index, so with the rest density given the result is
 * bit-identical to the one-GPU FLIP_SOLVER_RED_BLACK run (tests/test_gpu_multi.py).  params->max_particles is the
 * per-rank capacity: the owned particles plus TWO steps' ghosts (a step's arrivals are appended behind what the rank
 * already holds; what it no longer holds drops out in that step's sort).  Running out of room is an error on every rank
 * (FLIPGPU_ESTATE), never a silent drop.  Only FLIP_SOLVER_RED_BLACK, update_colors = 0. */
int flip_create_sharded(const FlipParams* params, int global_particles, int rank, int nranks, const void* unique_id128, FlipSim** out);
/* same with caller-chosen slabs: row_starts[0..nranks] ascending from 0 to f_num_y, rank r owns rows [row_starts[r], row_starts[r+1])
 * (every slab at least halo_rows+1 rows); NULL = equal slabs.  Lets the caller balance a scene whose liquid does not fill the
 * partition axis evenly (BASELINE configs[3]). */
int flip_create_sharded_rows(const FlipParams* params, int global_particles, int rank, int nranks, const int* row_starts, const void* unique_id128,
                             FlipSim** out);
int flip_get_shard_info(FlipSim* sim, int* j0, int* j1, int* halo_rows, int* ghost_band_rows, long long* exchanges);
int flip_upload_particles(FlipSim* sim, const float* pos, const float* vel /* may be NULL = 0 */, const int* global_ids, int n);
int flip_readback_particles(FlipSim* sim, float* pos, float* vel, int* global_ids, int capacity, int* n_out);
int flip_upload_rows(FlipSim* sim, int field, const void* host, int j_start, int n_rows);   /* grid fields, nx elements per row */
int flip_readback_rows(FlipSim* sim, int field, void* host, int j_start, int n_rows);

/* ------------------------------------------------------------------ multi-GPU: slab-sharded projection
 * One process per GPU.  The grid of nf x ns cells (nf = fast/contiguous axis) is cut into nranks slabs of rows
 * along the slow axis; rank r owns rows [j0,j1) = sorshard_slab(ns, r, nranks) plus H halo rows on each side.
 * sorshard_solve runs red-black SOR sweeps (the update of flip_fluid.h:461-491 / fluid.h:110-135) with one
 * NCCL neighbour exchange of H velocity rows every H-1 colour passes; results are bit-identical to the single-GPU
 * FLIP_SOLVER_RED_BLACK solve.  Rank 0 makes the communicator id, the host program broadcasts it (MPI,
 * torch.distributed, a file ...) and every rank passes it to sorshard_create. */
typedef struct SorShard SorShard;
typedef enum {
	SORSHARD_VEL_FAST = 0, /* velocity along the fast axis at the low face: u for FlipFluid, v for Fluid */
	SORSHARD_VEL_SLOW,     /* velocity along the slow axis: v for FlipFluid, u for Fluid */
	SORSHARD_P, SORSHARD_S, SORSHARD_DENSITY, /* float */
	SORSHARD_CELL_TYPE     /* int, 0 = fluid */
} SorShardField;
int flipgpu_nccl_unique_id(void* out128);                     /* 128 bytes (ncclUniqueId) */
void sorshard_slab(int ns, int rank, int nranks, int* j0, int* j1);
int sorshard_create(int nf, int ns, int x_fast, int rank, int nranks, const void* unique_id128, int device, SorShard** out);
int sorshard_destroy(SorShard* h);
int sorshard_get_rows(SorShard* h, int* j0, int* j1, int* halo_rows);
/* rows [j_start, j_start+n_rows) in global numbering, inside [j0-halo, j1+halo); host holds n_rows*nf elements */
int sorshard_upload_rows(SorShard* h, int field, const void* host, int j_start, int n_rows);
int sorshard_readback_rows(SorShard* h, int field, void* host, int j_start, int n_rows);
int sorshard_solve(SorShard* h, int iters, float cp, float omega, int compensate_drift, float rest_density);
int sorshard_synchronize(SorShard* h);
int sorshard_get_stream(SorShard* h, void** stream);
int sorshard_get_counters(SorShard* h, long long* launches, long long* exchanges);

#ifdef __cplusplus
}
#endif
#endif

/*
 * femflow_mpm.h -- C ABI of the B200-native MLS/APIC MPM substep.
 *
 * This is the drop-in boundary for the one hot path of jparr721/FEMFlow that
 * this library replaces: P2G scatter -> grid update -> G2P gather.  The
 * reference has no FFI of its own (it is pure Python + numba); each entry point
 * below names the reference interface it stands in for (paths relative to the
 * reference tree):
 *
 *   ffmpm_substep      femflow/solvers/mpm/mls_mpm.py:40-79    solve_mls_mpm_3d
 *                      (+ the 2D phase sequence, SURVEY 3.4, which has no driver)
 *   ffmpm_p2g          femflow/solvers/mpm/three_d/p2g.py:14-80, two_d/p2g.py:11-76
 *   ffmpm_grid_op      femflow/solvers/mpm/three_d/grid_op.py:5-47, two_d/grid_op.py:5-24
 *   ffmpm_g2p          femflow/solvers/mpm/three_d/g2p.py:9-59, two_d/g2p.py:5-47
 *   ffmpm_clear_grid   the per-substep np.zeros of mls_mpm.py:55-56
 *   ffmpm_bin          (new) cell binning of base_coord, three_d/p2g.py:50
 *   ffmpm_poll_error   the RuntimeError of three_d/p2g.py:51-52,70-71, g2p.py:23-24,35-36
 *   ffmpm_snapshot     femflow/solvers/mpm/particle.py:30-33   map_particles_to_pos
 *   ffmpm_set_materials
 *                      femflow/solvers/mpm/particle.py:7-13  Particle.mass / mu_0 / lambda_0
 *   ffmpm_set_colliders / ffmpm_collide
 *                      femflow/solvers/mpm/three_d/grid_op.py:50-67  check_collision_points
 *   ffmpm_scatter / ffmpm_gather / ffmpm_grid_op_halo
 *                      (new) the halves of a substep around the grid update, for slab drivers
 *   ffmpm_migrate_pack / ffmpm_migrate_unpack / ffmpm_set_owned_range / ffmpm_set_owned_slack
 *                      (new) slab ownership by base cell (three_d/p2g.py:50) and particle hand-over between slabs
 *   ffmpm_export_state (new) the live state back in the caller's particle order: the reference updates the caller's
 *                      arrays in place, particle i stays particle i (three_d/g2p.py:43-59)
 *   ffmpm_gen_implicit_points / ffmpm_gen_cube_points
 *                      femflow/simulation/mpm/primitives.py:46-76, numerics/geometry.py:101-116 (scene generators)
 *   ffmpm_g2p with model = FFMPM_SNOW in 3D: three_d/g2p.py:48-58 forms U clip(sig) Vh^T -- numpy's Vh transposed once
 *                      more -- so its result depends on the SIGNS of LAPACK's singular vectors.  csrc/mpm_svd3.cuh walks
 *                      DGESDD's operation sequence for 3x3 input (Householder bidiagonalisation, DBDSQR sweeps, the sign
 *                      flip and the sort, DORMBR) in fp64 and is pinned against np.linalg.svd itself and the reference's
 *                      own F_out / Jp_out (tests/golden/snow3d.npz).  It runs as a second launch after the G2P, which
 *                      carries F unchanged; the reference's driver never reaches this branch (mls_mpm.py:58).
 *
 * Conventions
 *   - Plain C, no torch types.  All array arguments are DEVICE pointers owned by
 *     the caller (torch tensors' data_ptr()); the library never allocates or
 *     frees device memory, so every call is CUDA-graph capturable.
 *   - Every function returns 0 on success or a negative FFMPM_E_* code; nothing
 *     throws across the boundary and nothing synchronises the device except
 *     ffmpm_poll_error.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - Particle state is SoA with one contiguous plane per scalar component:
 *     component c of particle p of a field with K components lives at
 *     field[c * stride + p].  Matrices are row-major: F[r][c] is component r*d+c.
 *   - Grid: node-major, 4 scalars per node, C order over (nx, ny, nz):
 *     {mom_x, mom_y, mom_z, mass} after P2G, {v_x, v_y, v_z, mass} after the
 *     grid update (2D: {mom_x, mom_y, mass, 0}, nz = 1).
 */
#ifndef FEMFLOW_MPM_H_
#define FEMFLOW_MPM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFMPM_ABI_VERSION 5

enum {
  FFMPM_OK = 0,
  FFMPM_E_INVALID = -1,   /* bad argument / unsupported configuration      */
  FFMPM_E_CUDA = -2,      /* a CUDA runtime call failed (see ffmpm_last_error) */
  FFMPM_E_OOB = -3,       /* a particle stencil left the grid (RuntimeError in the reference) */
  FFMPM_E_STATE = -4,     /* state not bound / workspace too small          */
  FFMPM_E_NONFINITE = -5  /* reserved */
};

enum { FFMPM_F32 = 0, FFMPM_F64 = 1 };            /* storage type of state and grid   */
enum { FFMPM_NEO_HOOKEAN = 0, FFMPM_SNOW = 1 };   /* `model` of three_d/p2g.py:28     */
enum {
  FFMPM_P2G_AUTO = 0,      /* tiled when the state is binned, else scatter */
  FFMPM_P2G_SCATTER = 1,   /* one thread per particle, one vector red per node */
  FFMPM_P2G_TILED = 2,     /* binned, warp-autonomous cell runs (needs ffmpm_bin) */
  FFMPM_P2G_FUSED = 3      /* as AUTO, and ffmpm_gather runs G2P(k) + P2G(k+1) as ONE kernel */
};

/* Scalars of solve_mls_mpm_3d's argument list (mls_mpm.py:40-53) plus the slab
 * extension (non-cubic local grid at an integer node offset inside a global
 * grid; reduces to the reference when origin = 0 and n = res + 1). */
typedef struct FfMpmConfig {
  int32_t dim;            /* 2 or 3                                              */
  int32_t dtype;          /* FFMPM_F32 / FFMPM_F64                               */
  int32_t model;          /* FFMPM_NEO_HOOKEAN / FFMPM_SNOW                      */
  int32_t res[3];         /* GLOBAL grid_resolution per axis (walls use it)      */
  int32_t n[3];           /* LOCAL node counts per axis (res+1 for one GPU)      */
  int32_t origin[3];      /* global index of local node 0 per axis (walls act on GLOBAL
                             indices, so only the ranks holding a global face see them) */
  double inv_dx, dx, dt, volume, gravity, hardening;
  /* 2D only (two_d/p2g.py:14-16 takes global material scalars); also used in
   * 3D when the per-particle arrays of FfMpmState are NULL. */
  double mass, mu_0, lambda_0;
  int32_t p2g_mode;       /* FFMPM_P2G_*                                         */
  int32_t reserved[7];
} FfMpmConfig;

/* One particle-state buffer (device pointers).  Planes as described above. */
typedef struct FfMpmState {
  void* x;        /* dim   planes */
  void* v;        /* dim   planes */
  void* C;        /* dim^2 planes */
  void* F;        /* dim^2 planes */
  void* Jp;       /* 1 plane (may be NULL for 3D neo-hookean: never touched, quirk 7) */
  void* mass;     /* 1 plane, or NULL -> cfg.mass     */
  void* mu0;      /* 1 plane, or NULL -> cfg.mu_0     */
  void* lam0;     /* 1 plane, or NULL -> cfg.lambda_0 */
  int32_t* id;    /* original particle index (carried through reordering), or NULL */
  uint8_t* material; /* 1 byte per particle: row of the ffmpm_set_materials table; NULL -> row 0.
                      * Only read once a table is set, and then mass/mu0/lam0 must be NULL. */
  int64_t stride; /* elements between consecutive component planes (>= n)   */
} FfMpmState;

typedef struct FfMpmHandle FfMpmHandle;

int32_t ffmpm_abi_version(void);
const char* ffmpm_last_error(void);

/* Bytes of caller-provided device workspace needed for `capacity` particles
 * (grid + binning buffers + error flag).  Returns a negative code on error. */
int64_t ffmpm_workspace_bytes(const FfMpmConfig* cfg, int64_t capacity);

int ffmpm_create(const FfMpmConfig* cfg, int32_t device, FfMpmHandle** out);
void ffmpm_destroy(FfMpmHandle* h);
int ffmpm_set_workspace(FfMpmHandle* h, void* workspace, int64_t bytes);

/* Bind the particle buffers.  `cur` holds the live state; `alt` (same layout,
 * may be NULL) is the ping-pong target used when G2P writes particles back in
 * binned order.  n <= capacity given to ffmpm_workspace_bytes. */
int ffmpm_bind_state(FfMpmHandle* h, const FfMpmState* cur, const FfMpmState* alt, int64_t n);
/* Which of the two bound buffers holds the live state (0 = cur, 1 = alt). */
int ffmpm_live_buffer(const FfMpmHandle* h);
int64_t ffmpm_num_particles(const FfMpmHandle* h);
int ffmpm_set_num_particles(FfMpmHandle* h, int64_t n);

/* Phase entry points (each asynchronous on `stream`). */
int ffmpm_clear_grid(FfMpmHandle* h, void* stream);
int ffmpm_bin(FfMpmHandle* h, void* stream);
int ffmpm_p2g(FfMpmHandle* h, void* stream);
int ffmpm_grid_op(FfMpmHandle* h, void* stream);
int ffmpm_g2p(FfMpmHandle* h, void* stream);
/* Slab decomposition along axis 0 (new; the reference is single-process): the grid
 * update with the halo SUM fused into its load.  recv_lo / recv_hi hold the
 * neighbour ranks' partial {momentum, mass} for the first planes_lo / last planes_hi
 * node planes of this rank's local grid (same node-major layout, contiguous because
 * axis 0 is slowest).  With 0 planes it is ffmpm_grid_op. */
int ffmpm_grid_op_halo(FfMpmHandle* h, const void* recv_lo, int32_t planes_lo, const void* recv_hi,
                       int32_t planes_hi, void* stream);
/* Slabs: the GLOBAL base cells [own_lo, own_hi) along axis 0 owned by this rank.  The binned
 * G2P then counts, per substep, the particles whose new base cell lies outside that range in a
 * device counter (ffmpm_leaver_count_ptr; cleared by the next ffmpm_bin), so the migration logic
 * needs no pass of its own over the positions. */
int ffmpm_set_owned_range(FfMpmHandle* h, int32_t own_lo, int32_t own_hi);
/* count[0] = particles outside the owned range, count[1] = those of them MORE than `slack` cells outside it
 * (ffmpm_set_owned_slack, default 0): a slab driver whose halo margin has room lets particles stray and migrates
 * only when count[1] > 0, instead of every period. */
int ffmpm_set_owned_slack(FfMpmHandle* h, int32_t slack);
int ffmpm_leaver_count_ptr(FfMpmHandle* h, int32_t** count);
/* Slab migration on the device (new).  One round, all on `stream`, no host involvement:
 *   ffmpm_migrate_pack    every live particle whose GLOBAL base cell along axis 0 (three_d/p2g.py:50) left the owned
 *                         range of ffmpm_set_owned_range is copied into `out_lo` (cell < own_lo) or `out_hi`
 *                         (cell >= own_hi) and its slot is back-filled from the tail of the live buffer: only the
 *                         leavers and as many keepers move.  At most `cap` particles per side; the rest stays one
 *                         more round.  A NULL outbox = no neighbour on that side: those particles stay (a particle
 *                         outside the GLOBAL grid is then reported by the binning as in the reference).
 *   (caller: exchange out_lo / out_hi with the neighbour ranks -- fixed-size messages)
 *   ffmpm_migrate_unpack  appends the particles of `in_lo` / `in_hi` (NULL = none) behind the keepers and copies the
 *                         round's record {out_lo, out_hi, in_lo, in_hi, n_new, overflow} (6 x int32) to `record`
 *                         (device or pinned host memory).  The caller reads it ONCE, then ffmpm_set_num_particles(n_new).
 * Message = (1 + ffmpm_migrate_rows) x cap scalars of cfg.dtype: row 0 = header (element 0: particle count), then one
 * row per component: x3 v3 C9 F9, material (mass mu0 lam0 planes, or the table row and two unused rows), id bits, [Jp].
 * Needs the id plane; invalidates the binning of the live buffer (its key/rank/perm arrays are the scratch lists). */
int32_t ffmpm_migrate_rows(const FfMpmHandle* h);
int ffmpm_migrate_pack(FfMpmHandle* h, void* out_lo, void* out_hi, int32_t cap, void* stream);
int ffmpm_migrate_unpack(FfMpmHandle* h, const void* in_lo, const void* in_hi, int32_t cap, int32_t* record, void* stream);
/* The two halves of a substep either side of the grid update, as ffmpm_substep issues
 * them (slab drivers put the halo exchange in between):
 *   ffmpm_scatter  zeroed grid + cell binning + P2G   (mls_mpm.py:54-73).  The library keeps
 *                  two grids and an internal stream: the binning runs underneath the
 *                  compute-bound P2G, the idle grid is cleared underneath G2P.
 *   ffmpm_gather   G2P (mls_mpm.py:79), after joining the binning.
 * Both are ordered on `stream` like any other call. */
int ffmpm_scatter(FfMpmHandle* h, void* stream);
int ffmpm_gather(FfMpmHandle* h, void* stream);
/* n_substeps x (scatter, grid_op, gather).
 * The grid of a substep is one of TWO the library keeps (ffmpm_grid_ptr / ffmpm_grid_view return the current one: ask
 * again after every call that scatters).  3D: the idle grid is cleared on the internal stream underneath P2G.  2D:
 * ffmpm_grid_op zeroes the idle grid in the same pass, so a 2D substep is three launches (P2G, grid update, G2P:
 * two_d/{p2g,grid_op,g2p}.py) and an EVEN number of substeps ends on the grid it started from (what CUDA-graph
 * capture needs).  2D fp32 states whose x, v, C, F planes sit back to back at one stride run the warp-window kernels
 * (csrc/mpm_2d_window.cuh; FFMPM_2D_WINDOW=0 in the environment selects the thread-per-particle ones). */
int ffmpm_substep(FfMpmHandle* h, int32_t n_substeps, void* stream);

/* Material table (femflow/solvers/mpm/particle.py:7-13: every Particle carries mass, mu_0 and
 * lambda_0, but a scene only ever holds a handful of distinct triples -- one per mesh in
 * simulation.py:92-104).  With a table set, particles name their triple by a 1-byte row index
 * (FfMpmState.material) instead of three scalar planes: 1 instead of 12 bytes read by P2G and
 * 2 instead of 24 moved by the reordering G2P, bit-identical results.  `mass`, `mu0`, `lam0`:
 * HOST arrays of `count` doubles (rounded to the storage type like the planes would be);
 * 1 <= count <= FFMPM_MAX_MATERIALS, count == 0 removes the table.  Synchronous. */
#define FFMPM_MAX_MATERIALS 256
int ffmpm_set_materials(FfMpmHandle* h, const double* mass, const double* mu0, const double* lam0, int32_t count);

/* Plane colliders (femflow/solvers/mpm/three_d/grid_op.py:50-67 check_collision_points, unused
 * by the reference's driver): every node with dot(I*dx - point, normal + 1/|normal|) < 0 for
 * some (point, normal) gets zero velocity at the end of the grid update.  `points`, `normals`:
 * HOST arrays of count x 3 doubles (copied); count <= FFMPM_MAX_COLLIDERS, 0 clears them. */
#define FFMPM_MAX_COLLIDERS 8
int ffmpm_set_colliders(FfMpmHandle* h, const double* points, const double* normals, int32_t count);
/* check_collision_points on its own (phase-level parity): applies the colliders to the grid. */
int ffmpm_collide(FfMpmHandle* h, void* stream);

/* Phase-level access for parity tests: device pointer of the node-major grid
 * (n[0]*n[1]*n[2]*4 scalars of cfg.dtype). */
int ffmpm_grid_ptr(FfMpmHandle* h, void** grid);
/* The same pointer for READING only (slab drivers send halo planes straight out of grid memory).
 * ffmpm_grid_ptr must assume the caller writes, which turns off the block-list grid update and clear
 * for that substep; this accessor keeps them. */
int ffmpm_grid_view(FfMpmHandle* h, const void** grid);
/* Binning results (device pointers into the workspace, valid after ffmpm_bin):
 *   keys[p]          bin key of particle p of the live buffer: tile-major cell id
 *                    (tile = 4x4x4 base cells in 3D, 8x8 in 2D);
 *                    key == n_cells marks a particle whose stencil leaves the grid
 *   perm[s]          binned slot s -> particle index in the live buffer
 *   cell_offsets[k]  first slot of cell k (n_cells + 2 entries, exclusive scan)   */
int ffmpm_bin_ptrs(FfMpmHandle* h, int32_t** keys, int32_t** perm, int32_t** cell_offsets,
                   int64_t* n_cells);

/* Synchronises `stream`, reads and clears the sticky device error flag.
 * Returns FFMPM_E_OOB when n_oob > 0. */
int ffmpm_poll_error(FfMpmHandle* h, void* stream, int32_t* code, int64_t* n_oob);

/* Position snapshot, particle.py:30-33: out[3*id + c] = x[c] / coeff as f64,
 * `out` a device or pinned-host-mapped pointer of n*dim doubles. */
int ffmpm_snapshot(FfMpmHandle* h, double coeff, double* out, void* stream);

/* The live state written into `dst` (same SoA layout, any stride >= n; only x, v, C, F and Jp are touched) in the
 * caller's ORIGINAL particle order: dst.field[c * dst.stride + id[p]] = live.field[c * stride + p].  The reference
 * updates the caller's arrays in place, particle i staying particle i (three_d/g2p.py:43-59); the CUDA path stores
 * particles cell-sorted and carries their index in the id plane -- this is the way back for callers that hold
 * their state outside the library (HostSubstepPipeline).  Without an id plane it is a plain copy. */
int ffmpm_export_state(FfMpmHandle* h, const FfMpmState* dst, void* stream);

/* Scene point generators on the device (the step in front of the path).  No handle needed.
 *   ffmpm_gen_implicit_points  femflow/simulation/mpm/primitives.py:46-61 generate_implicit_points over the lattice of
 *                              numerics/geometry.py:101-116 grid((res, res, res)): the points p with f(p) - t > t for
 *                              f = gyroid (kind 0) / diamond (1) / primitive (2), primitives.py:8-43, in lattice order.
 *                              Asynchronous; the number of selected points is left as an int64 at scratch[0] (a call
 *                              with capacity 0 only counts); `out` receives min(count, capacity) rows of 3 doubles.
 *                              `scratch`: ffmpm_scene_scratch_bytes(res) bytes of device memory, 256-byte aligned.
 *   ffmpm_gen_cube_points      primitives.py:64-76 generate_cube_points: res^3 rows [z, y, x] (x fastest) of the
 *                              np.linspace lattices of the three intervals; bounds6 = {x0, x1, y0, y1, z0, z1} (host). */
int64_t ffmpm_scene_scratch_bytes(int32_t res);
int ffmpm_gen_implicit_points(int32_t kind, double k, double t, int32_t res, void* scratch, double* out, int64_t capacity,
                              void* stream);
int ffmpm_gen_cube_points(const double* bounds6_host, int32_t res, double* out, void* stream);

/* Number of kernel launches issued by this handle so far. */
int64_t ffmpm_launch_count(const FfMpmHandle* h);

/* Diagnostic (no reference counterpart): dst[4*i .. 4*i+3] += value[0..3] for i < count with the one
 * 16-byte vector reduction P2G issues per node (red.global.add.v4.f32).  `dst` may be PEER memory mapped
 * over NVLink: scripts/peer_red_probe.py uses it to find out whether P2G could scatter its halo planes
 * straight into a neighbour GPU's inbox.  fp32 only; asynchronous on `stream`. */
int ffmpm_debug_red_add4(float* dst, const float* value4_host, int64_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEMFLOW_MPM_H_ */

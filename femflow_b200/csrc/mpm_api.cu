// C ABI of the MPM substep library (see include/femflow_mpm.h).
#include <cuda_runtime.h>

#include <cmath>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/femflow_mpm.h"
#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"
#include "mpm_p2g_bulk.cuh"
#include "mpm_tiled.cuh"
#include "mpm_g2p2g.cuh"
#include "mpm_migrate.cuh"
#include "mpm_2d.cuh"
#include "mpm_2d_window.cuh"
#include "mpm_scene.cuh"
#include "mpm_svd3.cuh"

using namespace ffmpm;

static thread_local char g_last_error[512] = "";

static int set_err(int code, const char* fmt, const char* detail = "") {
  snprintf(g_last_error, sizeof(g_last_error), fmt, detail);
  return code;
}

#define CUDA_TRY(expr)                                                          \
  do {                                                                          \
    cudaError_t e__ = (expr);                                                   \
    if (e__ != cudaSuccess) return set_err(FFMPM_E_CUDA, #expr ": %s", cudaGetErrorString(e__)); \
  } while (0)

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct FfMpmHandle {
  FfMpmConfig cfg;
  DevCfg dev;
  Colliders colliders;
  int device;
  int sm_count;
  int64_t n_nodes;
  // workspace carve-up
  char* ws;
  int64_t ws_bytes;
  ErrRec* err;
  void* grid;          // grid of the current substep (one of grids[])
  void* grids[2];      // ping-pong: the idle one is cleared behind G2P on the auxiliary stream
  int grid_cur;
  bool grid_clean[2];  // known to be all-zero
  bool scatter_ahead;  // grids[grid_cur ^ 1] already holds P2G of the live state (fused G2P2G)
  int fuse;            // FFMPM_FUSE=0 disables the fused G2P2G kernel
  int win2;            // 2D fp32, unbinned: warp-window kernels of mpm_2d_window.cuh (FFMPM_2D_WINDOW=0: thread per particle)
  int gg_blocks_per_sm;
  bool bin_pending;    // the binning of this substep is in flight on `aux` (ev_join)
  cudaStream_t aux;    // internal stream for work that overlaps the compute-bound P2G
  cudaEvent_t ev_fork, ev_join;
  int overlap;         // FFMPM_OVERLAP=0 disables the auxiliary stream
  BinBuffers bin;
  // state
  FfMpmState st[2];
  bool have_alt;
  int live;
  int64_t n;
  int64_t capacity;   // capacity the workspace was sized for (derived from ws_bytes)
  bool binned;        // bin buffers describe the live buffer
  int64_t launches;
  bool prebinned;     // keys/rank/histogram of the live buffer were emitted by the last reordering G2P
  int p2g_variant;    // see p2g_t (FFMPM_P2G_VARIANT)
  int p2g_blocks_per_sm, g2p_blocks_per_sm;   // persistent-grid sizing (tunable: FFMPM_P2G_BPS / FFMPM_G2P_BPS)
  bool grid_in_blocks; // the current grid was zero before a P2G of exactly the binned particles: everything
                       // non-zero lies inside the node blocks listed by the binning (bin.node_tiles)
  int sparse_grid_op;  // FFMPM_SPARSE_GRID_OP=0: always update / clear the whole grid
  int bin_slot;        // which of bin.node_tiles2[] the last binning filled (= grid_cur at that time)
  bool grid_listed[2]; // everything non-zero in grids[g] lies inside the blocks of bin.node_tiles2[g]
  void* mat_table;    // device, [3][MAT_ROWS] of the storage type (ffmpm_set_materials)
  int n_materials;
};

static size_t elem_size(const FfMpmConfig& c) { return c.dtype == FFMPM_F64 ? 8 : 4; }

static int validate(const FfMpmConfig* c) {
  if (!c) return set_err(FFMPM_E_INVALID, "null config");
  if (c->dim != 2 && c->dim != 3) return set_err(FFMPM_E_INVALID, "dim must be 2 or 3");
  if (c->dtype != FFMPM_F32 && c->dtype != FFMPM_F64) return set_err(FFMPM_E_INVALID, "bad dtype");
  if (c->model != FFMPM_NEO_HOOKEAN && c->model != FFMPM_SNOW) return set_err(FFMPM_E_INVALID, "bad model");
  for (int d = 0; d < c->dim; ++d) {
    if (c->n[d] < 3 || c->res[d] < 2) return set_err(FFMPM_E_INVALID, "grid too small");
  }
  if (c->dim == 2 && c->n[2] != 1) return set_err(FFMPM_E_INVALID, "2D needs n[2] == 1");
  int64_t nodes = (int64_t)c->n[0] * c->n[1] * c->n[2];
  if (nodes >= (1LL << 31)) return set_err(FFMPM_E_INVALID, "more than 2^31 nodes per GPU");
  if (!(c->dt > 0) || !(c->dx > 0) || !(c->inv_dx > 0)) return set_err(FFMPM_E_INVALID, "dt, dx, inv_dx must be > 0");
  return FFMPM_OK;
}

struct WsLayout {
  int64_t err_off, grid_off, grid2_off, bin_off, total;
};

static WsLayout layout(const FfMpmConfig& c, int64_t capacity) {
  WsLayout L;
  int64_t nodes = (int64_t)c.n[0] * c.n[1] * c.n[2];
  L.err_off = 0;
  L.grid_off = 256;
  L.grid2_off = align_up(L.grid_off + nodes * 4 * (int64_t)elem_size(c), 256);
  L.bin_off = align_up(L.grid2_off + nodes * 4 * (int64_t)elem_size(c), 256);
  L.total = L.bin_off + bin_workspace_bytes(c.dim, c.n, capacity);
  return L;
}

// Definitions below pick up C linkage from their declarations in femflow_mpm.h.

int32_t ffmpm_abi_version(void) { return FFMPM_ABI_VERSION; }
const char* ffmpm_last_error(void) { return g_last_error; }

int64_t ffmpm_workspace_bytes(const FfMpmConfig* cfg, int64_t capacity) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (capacity < 0) return set_err(FFMPM_E_INVALID, "negative capacity");
  return layout(*cfg, capacity).total;
}

int ffmpm_create(const FfMpmConfig* cfg, int32_t device, FfMpmHandle** out) {
  if (!out) return set_err(FFMPM_E_INVALID, "null out");
  int rc = validate(cfg);
  if (rc) return rc;
  int count = 0;
  CUDA_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return set_err(FFMPM_E_INVALID, "no such CUDA device");
  FfMpmHandle* h = new (std::nothrow) FfMpmHandle();
  if (!h) return set_err(FFMPM_E_INVALID, "out of host memory");
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->device = device;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete h; return set_err(FFMPM_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
  h->sm_count = prop.multiProcessorCount;
  DevCfg& d = h->dev;
  d.dim = cfg->dim; d.model = cfg->model;
  for (int i = 0; i < 3; ++i) {
    d.n[i] = cfg->n[i]; d.origin[i] = cfg->origin[i]; d.res[i] = cfg->res[i];
  }
  if (cfg->dim == 2) { d.n[2] = 1; d.origin[2] = 0; d.res[2] = 2; }
  d.inv_dx = cfg->inv_dx; d.dx = cfg->dx; d.dt = cfg->dt; d.volume = cfg->volume;
  d.gravity = cfg->gravity; d.hardening = cfg->hardening;
  d.mass = cfg->mass; d.mu0 = cfg->mu_0; d.lam0 = cfg->lambda_0;
  {
    int e = 0;
    const double m = frexp(cfg->inv_dx, &e);
    d.index_fp32 = (m == 0.5 && e > -100 && e < 100) ? 1 : 0;   // inv_dx == 2^(e-1) exactly
    if (const char* ev = getenv("FFMPM_INDEX_FP32")) d.index_fp32 = d.index_fp32 && atoi(ev) != 0;
  }
  d.own_lo = INT32_MIN;   // single domain: nobody leaves
  d.own_hi = INT32_MAX;
  d.own_slack = 0;
  d.fp32_stress = 1;   // fp32 build: left-form perturbation series (mpm_math.cuh); FFMPM_FP32_STRESS=0: fp64 stress always
  if (const char* e = getenv("FFMPM_FP32_STRESS")) d.fp32_stress = atoi(e) != 0;
  h->n_nodes = (int64_t)d.n[0] * d.n[1] * d.n[2];
  h->overlap = 1;
  if (const char* e = getenv("FFMPM_OVERLAP")) h->overlap = atoi(e) != 0;
  // Fused G2P2G: opt-in (p2g_mode FUSED or FFMPM_FUSE=1).  Measured on B200 (profiles/r01i): the fused
  // kernel needs 128 registers (16 warps/SM) and runs 1.61 ms against 0.78 + 0.83 ms for the two
  // separate kernels whose binning/clear additionally hide under P2G, so it is not the default.
  h->fuse = cfg->p2g_mode == FFMPM_P2G_FUSED ? 1 : 0;
  if (const char* e = getenv("FFMPM_FUSE")) h->fuse = atoi(e) != 0;
  h->win2 = 1;
  if (const char* e = getenv("FFMPM_2D_WINDOW")) h->win2 = atoi(e) != 0;
  h->gg_blocks_per_sm = 4;
  if (const char* e = getenv("FFMPM_GG_BPS")) { int v = atoi(e); if (v > 0 && v <= 32) h->gg_blocks_per_sm = v; }
  {
    cudaError_t ce = cudaSetDevice(device);
    int prio_lo = 0, prio_hi = 0;
    if (ce == cudaSuccess) ce = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    // highest priority: its short memory-bound kernels slot in as CTAs of the compute-bound P2G retire
    if (ce == cudaSuccess) ce = cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, prio_hi);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    if (ce != cudaSuccess) { delete h; return set_err(FFMPM_E_CUDA, "stream/event creation: %s", cudaGetErrorString(ce)); }
  }
  h->sparse_grid_op = 1;
  if (const char* e = getenv("FFMPM_SPARSE_GRID_OP")) h->sparse_grid_op = atoi(e) != 0;
  h->p2g_blocks_per_sm = 5;
  h->g2p_blocks_per_sm = 8;
  h->p2g_variant = 5;   // physical-order P2G with cp.async-prefetched state when eligible (profiles/r01j, r02a)
  if (const char* e = getenv("FFMPM_P2G_VARIANT")) h->p2g_variant = atoi(e);
  if (const char* e = getenv("FFMPM_P2G_BPS")) { int v = atoi(e); if (v > 0 && v <= 32) h->p2g_blocks_per_sm = v; }
  if (const char* e = getenv("FFMPM_G2P_BPS")) { int v = atoi(e); if (v > 0 && v <= 32) h->g2p_blocks_per_sm = v; }
  *out = h;
  return FFMPM_OK;
}

void ffmpm_destroy(FfMpmHandle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->aux) { cudaStreamSynchronize(h->aux); cudaStreamDestroy(h->aux); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->mat_table) cudaFree(h->mat_table);
  delete h;
}

int ffmpm_set_workspace(FfMpmHandle* h, void* workspace, int64_t bytes) {
  if (!h || !workspace) return set_err(FFMPM_E_INVALID, "null handle/workspace");
  if (((uintptr_t)workspace & 255) != 0) return set_err(FFMPM_E_INVALID, "workspace must be 256-byte aligned");
  WsLayout L0 = layout(h->cfg, 0);
  if (bytes < L0.total) return set_err(FFMPM_E_STATE, "workspace too small for the grid");
  // largest capacity that fits
  int64_t lo = 0, hi = (int64_t)1 << 40;
  while (lo < hi) {
    int64_t mid = lo + (hi - lo + 1) / 2;
    if (layout(h->cfg, mid).total <= bytes) lo = mid; else hi = mid - 1;
  }
  h->capacity = lo;
  WsLayout L = layout(h->cfg, lo);
  h->ws = (char*)workspace;
  h->ws_bytes = bytes;
  h->err = (ErrRec*)(h->ws + L.err_off);
  h->grids[0] = h->ws + L.grid_off;
  h->grids[1] = h->ws + L.grid2_off;
  h->grid_cur = 0;
  h->grid = h->grids[0];
  h->grid_clean[0] = h->grid_clean[1] = false;
  h->grid_listed[0] = h->grid_listed[1] = false;
  h->grid_in_blocks = false;
  h->bin_pending = false;
  h->scatter_ahead = false;
  bin_carve(h->bin, h->ws + L.bin_off, h->cfg.dim, h->dev.n, lo);
  h->binned = false;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemset(h->err, 0, sizeof(ErrRec)));
  return FFMPM_OK;
}

static bool state_ok(const FfMpmHandle* h, const FfMpmState* s) {
  if (!s->x || !s->v || !s->C || !s->F) return false;
  // per-particle material: all three planes or none (then the config scalars apply)
  const int mats = (s->mass != nullptr) + (s->mu0 != nullptr) + (s->lam0 != nullptr);
  if (mats != 0 && mats != 3) return false;
  if (h->cfg.dim == 2 && !s->Jp) return false;
  if (h->cfg.model == FFMPM_SNOW && !s->Jp) return false;
  return true;
}

int ffmpm_bind_state(FfMpmHandle* h, const FfMpmState* cur, const FfMpmState* alt, int64_t n) {
  if (!h || !cur) return set_err(FFMPM_E_INVALID, "null handle/state");
  if (n < 0 || n > cur->stride) return set_err(FFMPM_E_INVALID, "n must be in [0, stride]");
  if (!state_ok(h, cur)) return set_err(FFMPM_E_INVALID, "state is missing a required plane");
  h->st[0] = *cur;
  h->have_alt = false;
  if (alt) {
    if (!state_ok(h, alt) || alt->stride != cur->stride) return set_err(FFMPM_E_INVALID, "alt buffer layout mismatch");
    if ((cur->mass != nullptr) != (alt->mass != nullptr) || (cur->mu0 != nullptr) != (alt->mu0 != nullptr) ||
        (cur->lam0 != nullptr) != (alt->lam0 != nullptr) || (cur->id != nullptr) != (alt->id != nullptr) ||
        (cur->material != nullptr) != (alt->material != nullptr) ||
        (cur->Jp != nullptr) != (alt->Jp != nullptr))
      return set_err(FFMPM_E_INVALID, "alt buffer must carry the same optional planes");
    h->st[1] = *alt;
    h->have_alt = true;
  }
  h->live = 0;
  h->n = n;
  h->binned = false;
  h->prebinned = false;
  h->scatter_ahead = false;
  h->grid_in_blocks = false;
  h->grid_listed[0] = h->grid_listed[1] = false;
  h->grid_clean[0] = h->grid_clean[1] = false;
  return FFMPM_OK;
}

int ffmpm_set_materials(FfMpmHandle* h, const double* mass, const double* mu0, const double* lam0, int32_t count) {
  if (!h) return set_err(FFMPM_E_INVALID, "null handle");
  if (count < 0 || count > FFMPM_MAX_MATERIALS) return set_err(FFMPM_E_INVALID, "material count must be in [0, 256]");
  if (count > 0 && (!mass || !mu0 || !lam0)) return set_err(FFMPM_E_INVALID, "null material arrays");
  static_assert(FFMPM_MAX_MATERIALS == MAT_ROWS, "table rows");
  CUDA_TRY(cudaSetDevice(h->device));
  if (count > 0) {
    const size_t es = elem_size(h->cfg);
    if (!h->mat_table) CUDA_TRY(cudaMalloc(&h->mat_table, 3 * MAT_ROWS * sizeof(double)));
    alignas(8) unsigned char host[3 * MAT_ROWS * sizeof(double)] = {};
    const double* rows[3] = {mass, mu0, lam0};
    for (int r = 0; r < 3; ++r)
      for (int i = 0; i < count; ++i) {
        if (es == 8) ((double*)host)[r * MAT_ROWS + i] = rows[r][i];
        else ((float*)host)[r * MAT_ROWS + i] = (float)rows[r][i];
      }
    CUDA_TRY(cudaDeviceSynchronize());   // nothing in flight may still read the old table
    CUDA_TRY(cudaMemcpy(h->mat_table, host, 3 * MAT_ROWS * es, cudaMemcpyHostToDevice));
  }
  h->n_materials = count;
  h->scatter_ahead = false;
  // one material: the kernels take it from the config scalars, rounded to the storage type like a plane would be
  const bool f64 = h->cfg.dtype == FFMPM_F64;
  h->dev.mass = count == 1 ? (f64 ? mass[0] : (double)(float)mass[0]) : h->cfg.mass;
  h->dev.mu0 = count == 1 ? (f64 ? mu0[0] : (double)(float)mu0[0]) : h->cfg.mu_0;
  h->dev.lam0 = count == 1 ? (f64 ? lam0[0] : (double)(float)lam0[0]) : h->cfg.lambda_0;
  return FFMPM_OK;
}

int ffmpm_live_buffer(const FfMpmHandle* h) { return h ? h->live : FFMPM_E_INVALID; }
int64_t ffmpm_num_particles(const FfMpmHandle* h) { return h ? h->n : FFMPM_E_INVALID; }
int ffmpm_set_num_particles(FfMpmHandle* h, int64_t n) {
  if (!h || n < 0 || n > h->st[0].stride) return set_err(FFMPM_E_INVALID, "bad particle count");
  h->n = n;
  h->binned = false;
  h->prebinned = false;
  h->scatter_ahead = false;
  h->grid_in_blocks = false;
  h->grid_listed[0] = h->grid_listed[1] = false;
  h->grid_clean[0] = h->grid_clean[1] = false;
  return FFMPM_OK;
}

template <typename T>
static StateView<T> view(const FfMpmHandle* h, const FfMpmState& s) {
  StateView<T> v;
  v.x = (T*)s.x; v.v = (T*)s.v; v.C = (T*)s.C; v.F = (T*)s.F; v.Jp = (T*)s.Jp;
  v.mass = (T*)s.mass; v.mu0 = (T*)s.mu0; v.lam0 = (T*)s.lam0; v.id = s.id; v.stride = s.stride;
  // The row plane is only meaningful with a table (ready() rejects the other combinations).  A one-row
  // table is served through the config scalars (ffmpm_set_materials overrides them in h->dev): no
  // per-particle material traffic or lookups at all.
  v.mat_table = h->n_materials > 1 ? (const T*)h->mat_table : nullptr;
  v.material = h->n_materials > 1 ? s.material : nullptr;
  return v;
}

static int ready(FfMpmHandle* h) {
  if (!h) return set_err(FFMPM_E_INVALID, "null handle");
  if (!h->ws) return set_err(FFMPM_E_STATE, "workspace not set");
  if (!h->st[0].x) return set_err(FFMPM_E_STATE, "state not bound");
  if (h->n_materials > 0 && h->st[0].mass) return set_err(FFMPM_E_STATE, "material table and mass/mu0/lam0 planes are exclusive");
  if (h->n_materials == 0 && h->st[0].material) return set_err(FFMPM_E_STATE, "material rows bound but no table set (ffmpm_set_materials)");
  cudaError_t e = cudaSetDevice(h->device);  // callable from any host thread (simulation.py:117)
  if (e != cudaSuccess) return set_err(FFMPM_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return FFMPM_OK;
}

static int check_launch(FfMpmHandle* h, int n_launches) {
  h->launches += n_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(FFMPM_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return FFMPM_OK;
}

int ffmpm_clear_grid(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(h->grid, 0, (size_t)h->n_nodes * 4 * elem_size(h->cfg), (cudaStream_t)stream));
  h->grid_clean[h->grid_cur] = true;
  return FFMPM_OK;
}

template <typename T>
static int bin_t(FfMpmHandle* h, cudaStream_t s) {
  if (h->n > h->capacity) return set_err(FFMPM_E_STATE, "workspace too small for this particle count");
  h->bin_slot = h->grid_cur;   // the node-block list belongs to the grid this substep's P2G writes
  h->bin.node_tiles = h->bin.node_tiles2[h->bin_slot];
  h->bin.node_count = h->bin.node_counts + h->bin_slot;
  int nl = bin_particles<T>(h->dev, view<T>(h, h->st[h->live]), h->n, h->bin, h->err, h->prebinned, s);
  h->binned = true;
  h->prebinned = false;
  return check_launch(h, nl);
}

int ffmpm_bin(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (h->n == 0) { h->binned = true; return FFMPM_OK; }
  return h->cfg.dtype == FFMPM_F64 ? bin_t<double>(h, (cudaStream_t)stream) : bin_t<float>(h, (cudaStream_t)stream);
}

template <typename T>
static int p2g_t(FfMpmHandle* h, cudaStream_t s) {
  StateView<T> sv = view<T>(h, h->st[h->live]);
  int mode = h->cfg.p2g_mode;
  if (mode == FFMPM_P2G_FUSED) mode = FFMPM_P2G_AUTO;
  if (mode == FFMPM_P2G_AUTO) mode = h->binned ? FFMPM_P2G_TILED : FFMPM_P2G_SCATTER;
  if (mode == FFMPM_P2G_TILED) {
    if (!h->binned) return set_err(FFMPM_E_STATE, "tiled P2G needs ffmpm_bin first");
    if (h->cfg.dim == 2) {
      // 2D: warp-autonomous cell runs in physical order (mpm_2d.cuh); correct for any order, fast for a sorted one
      const long long windows = (h->n + P2G_WINDOW - 1) / P2G_WINDOW;
      long long want = (windows + P2G_RUN_WARPS - 1) / P2G_RUN_WARPS, cap = (long long)h->sm_count * 8;
      const int blocks = (int)(want < cap ? want : cap);
      p2g_runs2_kernel<T><<<blocks < 1 ? 1 : blocks, P2G_RUN_WARPS * 32, 0, s>>>(h->dev, sv, h->n, (T*)h->grid, h->err);
      return check_launch(h, 1);
    }
    // P2G variant (FFMPM_P2G_VARIANT): 0 = through the permutation, 1 = physical order (kept sorted by G2P),
    // 5 (default) = physical order with the window state prefetched by per-lane cp.async (fp32, mpm_p2g_bulk.cuh)
    if constexpr (sizeof(T) == 4) {
      if (h->p2g_variant >= 3 && p2g_bulk_eligible(h->dev, sv)) {
        if (!p2g_bulk_launch<4, 1>(h->dev, sv, h->n, (T*)h->grid, h->err, h->sm_count, h->p2g_blocks_per_sm, s))
          return set_err(FFMPM_E_CUDA, "could not configure the bulk P2G kernel");
        return check_launch(h, 1);
      }
    }
    const bool use_perm = h->p2g_variant == 0;
    int nl = p2g_runs<T>(h->dev, sv, h->n, h->bin, (T*)h->grid, h->err, h->sm_count, h->p2g_blocks_per_sm, use_perm, s);
    return check_launch(h, nl);
  }
  if constexpr (sizeof(T) == 4) {
    if (h->win2 && w2_eligible(h->dev, sv)) {
      if (!p2g_window2_launch(h->dev, sv, h->n, (T*)h->grid, h->err, h->sm_count, s))
        return set_err(FFMPM_E_CUDA, "could not configure the 2D window P2G kernel");
      return check_launch(h, 1);
    }
  }
  unsigned blocks = (unsigned)((h->n + 127) / 128);
  if (h->cfg.dim == 3)
    p2g_scatter3_kernel<T><<<blocks, 128, 0, s>>>(h->dev, sv, h->n, (T*)h->grid, h->err);
  else
    p2g_scatter2_kernel<T><<<blocks, 128, 0, s>>>(h->dev, sv, h->n, (T*)h->grid, h->err);
  return check_launch(h, 1);
}

int ffmpm_p2g(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  // a zero grid + the particles the bin buffers describe: what P2G writes stays inside bin.node_tiles
  h->grid_in_blocks = h->grid_clean[h->grid_cur] && h->binned && h->bin_slot == h->grid_cur && h->n > 0;
  h->grid_listed[h->grid_cur] = h->grid_in_blocks;
  h->grid_clean[h->grid_cur] = false;
  if (h->n == 0) return FFMPM_OK;
  return h->cfg.dtype == FFMPM_F64 ? p2g_t<double>(h, (cudaStream_t)stream) : p2g_t<float>(h, (cudaStream_t)stream);
}

template <typename T>
static int grid_op_t(FfMpmHandle* h, cudaStream_t s, const void* halo_lo = nullptr, long long nodes_lo = 0,
                     const void* halo_hi = nullptr, long long nodes_hi = 0) {
  unsigned blocks = (unsigned)((h->n_nodes + 255) / 256);
  if (h->grid_in_blocks && h->sparse_grid_op) {
    // the node-block list comes from the binning, which may still be in flight on the internal stream
    if (h->bin_pending) {
      CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join, 0));
      h->bin_pending = false;
    }
    h->grid_in_blocks = false;   // velocities now: a second update would have to see every node again
    if (h->cfg.dim == 2)
      grid_op2_blocks_kernel<T><<<h->sm_count * 4, 256, 0, s>>>(h->dev, (T*)h->grid, h->bin.node_tiles2[h->grid_cur],
                                                               h->bin.node_counts + h->grid_cur, h->bin.ntile[1]);
    else
    grid_op3_blocks_kernel<T><<<h->sm_count * 8, 256, 0, s>>>(h->dev, (T*)h->grid, h->n_nodes, (const T*)halo_lo, nodes_lo,
                                                             (const T*)halo_hi, nodes_hi, h->colliders,
                                                             h->bin.node_tiles2[h->grid_cur], h->bin.node_counts + h->grid_cur,
                                                             h->bin.ntile[1], h->bin.ntile[2]);
  } else if (h->cfg.dim == 3)
    grid_op3_kernel<T><<<blocks, 256, 0, s>>>(h->dev, (T*)h->grid, h->n_nodes, (const T*)halo_lo, nodes_lo,
                                               (const T*)halo_hi, nodes_hi, h->colliders);
  else {
    // 2D, dense: the same pass zeroes the IDLE grid (last read by the previous substep's G2P), which the next
    // ffmpm_scatter then switches to -- no separate clear on the critical path of a 2D substep
    const int idle = h->grid_cur ^ 1;
    T* clear = nullptr;
    if (h->win2 && !h->grid_clean[idle]) {
      clear = (T*)h->grids[idle];
      h->grid_clean[idle] = true;
      h->grid_listed[idle] = false;
    }
    grid_op2_kernel<T><<<blocks, 256, 0, s>>>(h->dev, (T*)h->grid, h->n_nodes, clear);
  }
  return check_launch(h, 1);
}

int ffmpm_grid_op(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  return h->cfg.dtype == FFMPM_F64 ? grid_op_t<double>(h, (cudaStream_t)stream) : grid_op_t<float>(h, (cudaStream_t)stream);
}

int ffmpm_grid_op_halo(FfMpmHandle* h, const void* recv_lo, int32_t planes_lo, const void* recv_hi, int32_t planes_hi,
                       void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (h->cfg.dim != 3) return set_err(FFMPM_E_INVALID, "slab halos are 3D only");
  if (planes_lo < 0 || planes_hi < 0 || planes_lo + planes_hi > h->dev.n[0] || (planes_lo && !recv_lo) ||
      (planes_hi && !recv_hi))
    return set_err(FFMPM_E_INVALID, "bad halo planes");
  const long long plane = (long long)h->dev.n[1] * h->dev.n[2];
  return h->cfg.dtype == FFMPM_F64
             ? grid_op_t<double>(h, (cudaStream_t)stream, recv_lo, planes_lo * plane, recv_hi, planes_hi * plane)
             : grid_op_t<float>(h, (cudaStream_t)stream, recv_lo, planes_lo * plane, recv_hi, planes_hi * plane);
}

template <typename T>
static int g2p_t(FfMpmHandle* h, cudaStream_t s) {
  StateView<T> sv = view<T>(h, h->st[h->live]);
  if (h->binned && h->have_alt && h->cfg.dim == 2) {
    // 2D, binned: thread per binned slot, new state in cell order into the other buffer, next substep pre-binned
    StateView<T> dst = view<T>(h, h->st[h->live ^ 1]);
    g2p_reorder2_kernel<T><<<(unsigned)((h->n + 255) / 256), 256, 0, s>>>(h->dev, sv, dst, h->n, h->bin, (const T*)h->grid, h->err);
    h->live ^= 1;
    h->binned = false;
    h->prebinned = true;
    return check_launch(h, 1);
  }
  if (h->binned && h->have_alt && h->cfg.dim == 3) {
    // binned: write the particles back in cell order into the other buffer
    StateView<T> dst = view<T>(h, h->st[h->live ^ 1]);
    // (the kernel pre-bins the advected particles for the next substep into the histogram that
    // bin_particles left cleared)
    int nl = g2p_tiled<T>(h->dev, sv, dst, h->n, h->bin, (const T*)h->grid, h->err, h->sm_count, h->g2p_blocks_per_sm, s);
    h->live ^= 1;
    h->binned = false;  // positions moved: perm / cell offsets are stale ...
    h->prebinned = true;  // ... but keys, ranks and the histogram of the new live buffer are ready
    if (h->cfg.model == FFMPM_SNOW) {   // three_d/g2p.py:48-58 on the F the kernel carried over unchanged
      snow_project3_kernel<T><<<(unsigned)((h->n + 127) / 128), 128, 0, s>>>(h->dev, dst, h->n);
      ++nl;
    }
    return check_launch(h, nl);
  }
  h->binned = false;
  h->prebinned = false;
  if constexpr (sizeof(T) == 4) {
    if (h->win2 && w2_eligible(h->dev, sv)) {
      if (!g2p_window2_launch(h->dev, sv, h->n, (const T*)h->grid, h->err, h->sm_count, s))
        return set_err(FFMPM_E_CUDA, "could not configure the 2D window G2P kernel");
      return check_launch(h, 1);
    }
  }
  unsigned blocks = (unsigned)((h->n + 127) / 128);
  if (h->cfg.dim == 3 && h->cfg.model == FFMPM_SNOW) {
    g2p_gather3_kernel<T, true><<<blocks, 128, 0, s>>>(h->dev, sv, h->n, (const T*)h->grid, h->err);
    snow_project3_kernel<T><<<blocks, 128, 0, s>>>(h->dev, sv, h->n);
    return check_launch(h, 2);
  }
  if (h->cfg.dim == 3)
    g2p_gather3_kernel<T><<<blocks, 128, 0, s>>>(h->dev, sv, h->n, (const T*)h->grid, h->err);
  else
    g2p_gather2_kernel<T><<<blocks, 128, 0, s>>>(h->dev, sv, h->n, (const T*)h->grid, h->err);
  return check_launch(h, 1);
}

int ffmpm_g2p(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (h->n == 0) return FFMPM_OK;
  return h->cfg.dtype == FFMPM_F64 ? g2p_t<double>(h, (cudaStream_t)stream) : g2p_t<float>(h, (cudaStream_t)stream);
}

// True when P2G walks the particles in physical order, i.e. does not consume the binning:
// the binning can then run on the auxiliary stream underneath it.
static bool p2g_independent_of_bin(const FfMpmHandle* h) {
  return h->p2g_variant >= 1;
}

static bool binned_pipeline(const FfMpmHandle* h) {
  return h->have_alt && h->cfg.p2g_mode != FFMPM_P2G_SCATTER;
}

// Binning of the live buffer (+ a cleared idle grid for the fused G2P2G) on the auxiliary
// stream, forked from `s` here and joined in ffmpm_gather.
static int fork_bin(FfMpmHandle* h, cudaStream_t s, bool clear_idle) {
  int rc;
  cudaStream_t w = h->overlap ? h->aux : s;
  if (h->overlap) {
    CUDA_TRY(cudaEventRecord(h->ev_fork, s));
    CUDA_TRY(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
  }
  if ((rc = ffmpm_bin(h, (void*)w))) return rc;
  if (clear_idle) {
    const int idle = h->grid_cur ^ 1;
    if (h->grid_listed[idle] && h->sparse_grid_op) {
      // only the node blocks its last substep touched (listed by that substep's binning) are non-zero
      if (h->cfg.dim == 2 && h->cfg.dtype == FFMPM_F64)
        grid_clear_blocks2_kernel<double><<<h->sm_count * 4, 256, 0, w>>>(h->dev, (double*)h->grids[idle], h->bin.node_tiles2[idle],
                                                                         h->bin.node_counts + idle, h->bin.ntile[1]);
      else if (h->cfg.dim == 2)
        grid_clear_blocks2_kernel<float><<<h->sm_count * 4, 256, 0, w>>>(h->dev, (float*)h->grids[idle], h->bin.node_tiles2[idle],
                                                                        h->bin.node_counts + idle, h->bin.ntile[1]);
      else if (h->cfg.dtype == FFMPM_F64)
        grid_clear_blocks_kernel<double><<<h->sm_count * 8, 256, 0, w>>>(h->dev, (double*)h->grids[idle], h->bin.node_tiles2[idle],
                                                                        h->bin.node_counts + idle, h->bin.ntile[1], h->bin.ntile[2]);
      else
        grid_clear_blocks_kernel<float><<<h->sm_count * 8, 256, 0, w>>>(h->dev, (float*)h->grids[idle], h->bin.node_tiles2[idle],
                                                                       h->bin.node_counts + idle, h->bin.ntile[1], h->bin.ntile[2]);
      if ((rc = check_launch(h, 1))) return rc;
    } else {
      CUDA_TRY(cudaMemsetAsync(h->grids[idle], 0, (size_t)h->n_nodes * 4 * elem_size(h->cfg), w));
    }
    h->grid_clean[idle] = true;
    h->grid_listed[idle] = false;
  }
  if (h->overlap) {
    CUDA_TRY(cudaEventRecord(h->ev_join, h->aux));
    h->bin_pending = true;
  }
  return FFMPM_OK;
}

int ffmpm_scatter(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const bool binned = binned_pipeline(h) && h->n > 0;
  const bool fuse = binned && h->fuse && h->cfg.dim == 3;
  if (h->scatter_ahead) {
    // the previous fused gather already scattered the live state into the idle grid
    h->scatter_ahead = false;
    h->grid_cur ^= 1;
    h->grid = h->grids[h->grid_cur];
    return fork_bin(h, s, fuse);
  }
  if (!h->grid_clean[h->grid_cur] && h->grid_clean[h->grid_cur ^ 1]) {
    h->grid_cur ^= 1;   // the grid that was cleared underneath the previous P2G
    h->grid = h->grids[h->grid_cur];
  }
  if (!h->grid_clean[h->grid_cur] && (rc = ffmpm_clear_grid(h, stream))) return rc;
  if (binned) {
    if (p2g_independent_of_bin(h)) {
      // binning and the clear of the idle grid (next substep's, or the fused kernel's target): underneath P2G
      if ((rc = fork_bin(h, s, true))) return rc;
    } else {
      if ((rc = ffmpm_bin(h, stream))) return rc;
      if (fuse) {
        CUDA_TRY(cudaMemsetAsync(h->grids[h->grid_cur ^ 1], 0, (size_t)h->n_nodes * 4 * elem_size(h->cfg), s));
        h->grid_clean[h->grid_cur ^ 1] = true;
      }
    }
  }
  return ffmpm_p2g(h, stream);
}

template <typename T>
static int g2p2g_t(FfMpmHandle* h, cudaStream_t s) {
  StateView<T> sv = view<T>(h, h->st[h->live]), dst = view<T>(h, h->st[h->live ^ 1]);
  int nl = g2p2g_tiled<T>(h->dev, sv, dst, h->bin, (const T*)h->grid, (T*)h->grids[h->grid_cur ^ 1], h->err, h->sm_count,
                          h->gg_blocks_per_sm, s);
  h->live ^= 1;
  h->binned = false;
  h->prebinned = true;
  h->grid_clean[h->grid_cur ^ 1] = false;
  h->grid_listed[h->grid_cur ^ 1] = false;
  h->grid_in_blocks = false;
  h->scatter_ahead = true;
  return check_launch(h, nl);
}

int ffmpm_gather(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->bin_pending) {
    CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join, 0));
    h->bin_pending = false;
  }
  const bool can_fuse = h->fuse && h->cfg.dim == 3 && binned_pipeline(h) && h->n > 0 && h->binned &&
                        h->grid_clean[h->grid_cur ^ 1] && !(h->cfg.dim == 3 && h->cfg.model == FFMPM_SNOW);
  if (can_fuse)
    return h->cfg.dtype == FFMPM_F64 ? g2p2g_t<double>(h, s) : g2p2g_t<float>(h, s);
  return ffmpm_g2p(h, stream);
}

int ffmpm_substep(FfMpmHandle* h, int32_t n_substeps, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  for (int32_t it = 0; it < n_substeps; ++it) {
    if ((rc = ffmpm_scatter(h, stream))) return rc;
    if ((rc = ffmpm_grid_op(h, stream))) return rc;
    if ((rc = ffmpm_gather(h, stream))) return rc;
  }
  return FFMPM_OK;
}

int ffmpm_set_colliders(FfMpmHandle* h, const double* points, const double* normals, int32_t count) {
  if (!h) return set_err(FFMPM_E_INVALID, "null handle");
  if (count < 0 || count > FFMPM_MAX_COLLIDERS || (count > 0 && (!points || !normals)))
    return set_err(FFMPM_E_INVALID, "bad collider arguments");
  if (count > 0 && h->cfg.dim != 3) return set_err(FFMPM_E_INVALID, "colliders are 3D only (three_d/grid_op.py:50-67)");
  h->colliders.count = count;
  for (int c = 0; c < count; ++c) {
    const double* nrm = normals + 3 * c;
    const double denom = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    for (int d = 0; d < 3; ++d) {
      h->colliders.point[c][d] = points[3 * c + d];
      h->colliders.normal[c][d] = nrm[d] + (1.0 / denom);   // grid_op.py:59-60: scalar added to every component
    }
  }
  return FFMPM_OK;
}

int ffmpm_set_owned_range(FfMpmHandle* h, int32_t own_lo, int32_t own_hi) {
  if (!h || own_lo > own_hi) return set_err(FFMPM_E_INVALID, "bad owned range");
  h->dev.own_lo = own_lo;
  h->dev.own_hi = own_hi;
  return FFMPM_OK;
}

int ffmpm_set_owned_slack(FfMpmHandle* h, int32_t slack) {
  if (!h || slack < 0) return set_err(FFMPM_E_INVALID, "bad slack");
  h->dev.own_slack = slack;
  return FFMPM_OK;
}

int ffmpm_leaver_count_ptr(FfMpmHandle* h, int32_t** count) {
  if (!h || !h->ws || !count) return set_err(FFMPM_E_STATE, "workspace not set");
  *count = &h->bin.counters[3];
  return FFMPM_OK;
}

// ---- slab migration (mpm_migrate.cuh) ----
static int mig_rows(const FfMpmHandle* h) { return MIG_ROWS + (h->st[0].Jp ? 1 : 0); }

int32_t ffmpm_migrate_rows(const FfMpmHandle* h) { return h ? mig_rows(h) : FFMPM_E_INVALID; }

static MigRec* mig_rec(FfMpmHandle* h) { return reinterpret_cast<MigRec*>(h->bin.counters + 32); }

template <typename T>
static int migrate_pack_t(FfMpmHandle* h, void* out_lo, void* out_hi, int cap, cudaStream_t s) {
  StateView<T> sv = view<T>(h, h->st[h->live]);
  sv.material = h->st[h->live].material;   // rows travel with their particle whatever the table size
  MigRec* rec = mig_rec(h);
  CUDA_TRY(cudaMemsetAsync(rec, 0, sizeof(MigRec), s));
  const int rows = mig_rows(h);
  if (h->n > 0) {
    mig_pack_kernel<T><<<(unsigned)((h->n + 255) / 256), 256, 0, s>>>(h->dev, sv, h->n, (T*)out_lo, (T*)out_hi, cap, rows, rec, h->bin.keys);
    const unsigned blocks = (unsigned)((2 * (long long)cap + 255) / 256);
    mig_match_kernel<T><<<blocks, 256, 0, s>>>(sv, h->n, cap, rec, h->bin.keys, h->bin.rank, h->bin.perm);
    mig_fill_kernel<T><<<blocks, 256, 0, s>>>(sv, rows, rec, h->bin.rank, h->bin.perm);
  }
  mig_headers_kernel<T><<<1, 32, 0, s>>>((T*)out_lo, (T*)out_hi, cap, rec);
  return check_launch(h, h->n > 0 ? 4 : 1);
}

int ffmpm_migrate_pack(FfMpmHandle* h, void* out_lo, void* out_hi, int32_t cap, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (h->cfg.dim != 3) return set_err(FFMPM_E_INVALID, "slab migration is 3D only");
  if (!h->st[h->live].id) return set_err(FFMPM_E_STATE, "slab migration needs the id plane (it marks the vacated slots)");
  if (cap < 1 || cap > (1 << 22) || 2 * (int64_t)cap > h->capacity) return set_err(FFMPM_E_INVALID, "outbox capacity must be in [1, min(2^22, capacity / 2)]");
  // the binning of the live buffer dies with the round: its key / rank / perm arrays are the scratch lists
  if (h->bin_pending) {
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_join, 0));
    h->bin_pending = false;
  }
  h->binned = false;
  h->prebinned = false;
  h->scatter_ahead = false;
  return h->cfg.dtype == FFMPM_F64 ? migrate_pack_t<double>(h, out_lo, out_hi, cap, (cudaStream_t)stream)
                                   : migrate_pack_t<float>(h, out_lo, out_hi, cap, (cudaStream_t)stream);
}

template <typename T>
static int migrate_unpack_t(FfMpmHandle* h, const void* in_lo, const void* in_hi, int cap, cudaStream_t s) {
  StateView<T> sv = view<T>(h, h->st[h->live]);
  sv.material = h->st[h->live].material;
  mig_unpack_kernel<T><<<(unsigned)((2 * (long long)cap + 255) / 256), 256, 0, s>>>(sv, (const T*)in_lo, (const T*)in_hi, cap, mig_rows(h), mig_rec(h));
  return check_launch(h, 1);
}

int ffmpm_migrate_unpack(FfMpmHandle* h, const void* in_lo, const void* in_hi, int32_t cap, int32_t* record, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (cap < 1 || !record) return set_err(FFMPM_E_INVALID, "bad migration arguments");
  rc = h->cfg.dtype == FFMPM_F64 ? migrate_unpack_t<double>(h, in_lo, in_hi, cap, (cudaStream_t)stream)
                                 : migrate_unpack_t<float>(h, in_lo, in_hi, cap, (cudaStream_t)stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(record, mig_rec(h), 6 * sizeof(int32_t), cudaMemcpyDefault, (cudaStream_t)stream));
  return FFMPM_OK;
}

int ffmpm_collide(FfMpmHandle* h, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (h->cfg.dim != 3) return set_err(FFMPM_E_INVALID, "colliders are 3D only (three_d/grid_op.py:50-67)");
  if (h->colliders.count == 0) return FFMPM_OK;
  unsigned blocks = (unsigned)((h->n_nodes + 255) / 256);
  if (h->cfg.dtype == FFMPM_F64)
    collide3_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(h->dev, (double*)h->grid, h->n_nodes, h->colliders);
  else
    collide3_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(h->dev, (float*)h->grid, h->n_nodes, h->colliders);
  return check_launch(h, 1);
}

int ffmpm_grid_ptr(FfMpmHandle* h, void** grid) {
  if (!h || !grid || !h->ws) return set_err(FFMPM_E_STATE, "workspace not set");
  *grid = h->grid;
  h->grid_clean[h->grid_cur] = false;   // the caller may write through this pointer
  h->grid_in_blocks = false;
  h->grid_listed[h->grid_cur] = false;
  return FFMPM_OK;
}

int ffmpm_grid_view(FfMpmHandle* h, const void** grid) {
  if (!h || !grid || !h->ws) return set_err(FFMPM_E_STATE, "workspace not set");
  *grid = h->grid;
  return FFMPM_OK;
}

int ffmpm_bin_ptrs(FfMpmHandle* h, int32_t** keys, int32_t** perm, int32_t** cell_offsets, int64_t* n_cells) {
  if (!h || !h->ws) return set_err(FFMPM_E_STATE, "workspace not set");
  if (keys) *keys = h->bin.keys;
  if (perm) *perm = h->bin.perm;
  if (cell_offsets) *cell_offsets = h->bin.cell_off;
  if (n_cells) *n_cells = h->bin.n_cells;
  return FFMPM_OK;
}

int ffmpm_poll_error(FfMpmHandle* h, void* stream, int32_t* code, int64_t* n_oob) {
  int rc = ready(h);
  if (rc) return rc;
  ErrRec rec;
  CUDA_TRY(cudaMemcpyAsync(&rec, h->err, sizeof(rec), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  if (rec.code != 0 || rec.n_oob != 0) CUDA_TRY(cudaMemsetAsync(h->err, 0, sizeof(rec), (cudaStream_t)stream));
  if (code) *code = (int32_t)rec.code;
  if (n_oob) *n_oob = (int64_t)rec.n_oob;
  if (rec.n_oob) return set_err(FFMPM_E_OOB, "particle stencil left the grid");
  return FFMPM_OK;
}

template <typename T>
__global__ void snapshot_kernel(StateView<T> s, long long n, int dim, double coeff, double* __restrict__ out) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  long long dst = s.id ? (long long)s.id[p] : p;
  for (int c = 0; c < dim; ++c) out[dst * dim + c] = (double)s.x[c * s.stride + p] / coeff;
}

int ffmpm_snapshot(FfMpmHandle* h, double coeff, double* out, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (!out || coeff == 0.0) return set_err(FFMPM_E_INVALID, "bad snapshot arguments");
  if (h->n == 0) return FFMPM_OK;
  unsigned blocks = (unsigned)((h->n + 255) / 256);
  // particle.py:32 divides by coeff (x / coeff != x * (1/coeff) in general): divide on the device too
  if (h->cfg.dtype == FFMPM_F64)
    snapshot_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(view<double>(h, h->st[h->live]), h->n, h->cfg.dim, coeff, out);
  else
    snapshot_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(view<float>(h, h->st[h->live]), h->n, h->cfg.dim, coeff, out);
  return check_launch(h, 1);
}

// Live state -> caller buffers in ORIGINAL particle order (the id plane undoes the cell-sorted storage order).
template <typename T>
__global__ void __launch_bounds__(256) export_by_id_kernel(StateView<T> s, StateView<T> d, long long n, int dim) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long q = s.id ? (long long)s.id[p] : p;
  const long long ss = s.stride, ds = d.stride;
  const int dd = dim * dim;
  for (int c = 0; c < dim; ++c) { d.x[c * ds + q] = s.x[c * ss + p]; d.v[c * ds + q] = s.v[c * ss + p]; }
  for (int c = 0; c < dd; ++c) { d.C[c * ds + q] = s.C[c * ss + p]; d.F[c * ds + q] = s.F[c * ss + p]; }
  if (s.Jp && d.Jp) d.Jp[q] = s.Jp[p];
}

int ffmpm_export_state(FfMpmHandle* h, const FfMpmState* dst, void* stream) {
  int rc = ready(h);
  if (rc) return rc;
  if (!dst || !dst->x || !dst->v || !dst->C || !dst->F || dst->stride < h->n) return set_err(FFMPM_E_INVALID, "bad export target");
  if (h->n == 0) return FFMPM_OK;
  const unsigned blocks = (unsigned)((h->n + 255) / 256);
  if (h->cfg.dtype == FFMPM_F64)
    export_by_id_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(view<double>(h, h->st[h->live]), view<double>(h, *dst), h->n, h->cfg.dim);
  else
    export_by_id_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(view<float>(h, h->st[h->live]), view<float>(h, *dst), h->n, h->cfg.dim);
  return check_launch(h, 1);
}

// ---- scene point generators (mpm_scene.cuh) ----
int64_t ffmpm_scene_scratch_bytes(int32_t res) {
  if (res < 1 || res > 2048) return set_err(FFMPM_E_INVALID, "lattice resolution must be in [1, 2048]");
  const long long total = (long long)res * res * res;
  const long long blocks = (total + SCENE_THREADS - 1) / SCENE_THREADS;
  return 256 + ((blocks * 4 + 255) / 256) * 256 + blocks * 8;
}

int ffmpm_gen_implicit_points(int32_t kind, double k, double t, int32_t res, void* scratch, double* out, int64_t capacity,
                              void* stream) {
  if (kind < 0 || kind > 2) return set_err(FFMPM_E_INVALID, "Invalid implicit function specified");   // primitives.py:57
  if (res < 1 || res > 2048 || !scratch || capacity < 0 || (capacity > 0 && !out) || !(k != 0.0))
    return set_err(FFMPM_E_INVALID, "bad scene generator arguments");
  const long long total = (long long)res * res * res;
  const int blocks = (int)((total + SCENE_THREADS - 1) / SCENE_THREADS);
  long long* count = (long long*)scratch;
  int* counts = (int*)((char*)scratch + 256);
  long long* offsets = (long long*)((char*)scratch + 256 + (((long long)blocks * 4 + 255) / 256) * 256);
  cudaStream_t s = (cudaStream_t)stream;
  implicit_count_kernel<<<blocks, SCENE_THREADS, 0, s>>>(kind, k, t, res, total, counts);
  scene_scan_kernel<<<1, 1024, 0, s>>>(counts, offsets, blocks, count);
  if (capacity > 0) implicit_write_kernel<<<blocks, SCENE_THREADS, 0, s>>>(kind, k, t, res, total, offsets, out, capacity);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(FFMPM_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return FFMPM_OK;
}

int ffmpm_gen_cube_points(const double* bounds6_host, int32_t res, double* out, void* stream) {
  if (!bounds6_host || !out || res < 1 || res > 2048) return set_err(FFMPM_E_INVALID, "bad scene generator arguments");
  const long long total = (long long)res * res * res;
  const double* b = bounds6_host;
  cube_points_kernel<<<(unsigned)((total + SCENE_THREADS - 1) / SCENE_THREADS), SCENE_THREADS, 0, (cudaStream_t)stream>>>(
      b[0], b[1], b[2], b[3], b[4], b[5], res, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(FFMPM_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return FFMPM_OK;
}

int64_t ffmpm_launch_count(const FfMpmHandle* h) { return h ? h->launches : 0; }

__global__ void debug_red_add4_kernel(float* __restrict__ dst, float a, float b, float c, float d, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) red_add4(dst + 4 * i, a, b, c, d);
}

int ffmpm_debug_red_add4(float* dst, const float* v, int64_t count, void* stream) {
  if (!dst || !v || count < 0 || ((uintptr_t)dst & 15) != 0) return set_err(FFMPM_E_INVALID, "bad red probe arguments");
  if (count == 0) return FFMPM_OK;
  debug_red_add4_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, v[0], v[1], v[2], v[3], count);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(FFMPM_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return FFMPM_OK;
}


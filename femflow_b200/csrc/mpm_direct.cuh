// Direct (unbinned) kernels: one thread per particle / per node.  These are the
// order-independent baseline of the path -- used for state that has not been
// binned, for the fp64 build, and as the in-library cross-check of the tiled
// kernels.  Reference loops: three_d/p2g.py:49-80, three_d/grid_op.py:27-47,
// three_d/g2p.py:21-59 and the two_d/ equivalents.
#pragma once
#include "mpm_common.cuh"

namespace ffmpm {

// ----------------------------------------------------------------------------
// Per-particle P2G payload, shared by the scatter and the tiled kernels.
// ----------------------------------------------------------------------------
template <typename T>
struct P2GParticle3 {
  int bx, by, bz;       // LOCAL base node
  T fx, fy, fz;
  T mvx, mvy, mvz, m;   // mass * v, mass
  T a00, a01, a02, a10, a11, a12, a20, a21, a22;  // affine = stress + mass*C
  bool ok;
};

// Plane numbering of the 3D per-particle inputs (the order of the bulk-prefetch slab too).
enum { P2G_X = 0, P2G_V = 3, P2G_C = 6, P2G_F = 15, P2G_MASS = 24, P2G_MU = 25, P2G_LAM = 26, P2G_NPLANES = 27 };

// `get(k)` returns plane k of the particle (k is a literal at every call site, so a
// loader that switches on k folds away); has_mat: per-particle mass/mu0/lam0 planes exist.
template <typename T, typename Get>
FFMPM_HD P2GParticle3<T> p2g_prepare3_from(const DevCfg& cfg, Get get, bool has_mat, double jp) {
  P2GParticle3<T> q;
  T x0 = get(P2G_X), x1 = get(P2G_X + 1), x2 = get(P2G_X + 2);
  int gx, gy, gz;
  base_fx(x0, cfg, gx, q.fx);
  base_fx(x1, cfg, gy, q.fy);
  base_fx(x2, cfg, gz, q.fz);
  q.bx = gx - cfg.origin[0]; q.by = gy - cfg.origin[1]; q.bz = gz - cfg.origin[2];
  // utils.py:138-150 with res = G: base < 0 or base + 2 >= G  -> RuntimeError
  q.ok = !(isnan((double)x0) || isnan((double)x1) || isnan((double)x2)) &&
         q.bx >= 0 && q.by >= 0 && q.bz >= 0 && q.bx + 2 < cfg.n[0] && q.by + 2 < cfg.n[1] && q.bz + 2 < cfg.n[2];
  if (!q.ok) return q;
  double mass = has_mat ? (double)get(P2G_MASS) : cfg.mass;
  double mu = has_mat ? (double)get(P2G_MU) : cfg.mu0;
  double lam = has_mat ? (double)get(P2G_LAM) : cfg.lam0;
  double e = cfg.hardening;                       // constant_hardening: a plain multiplier (quirk 8)
  if (cfg.model == 1) e = exp(cfg.hardening * (1.0 - jp));  // snow_hardening, utils.py:48
  mu *= e; lam *= e;
  const double k = (cfg.dt * cfg.volume) * (4.0 * cfg.inv_dx * cfg.inv_dx);
  const T f00 = get(P2G_F + 0), f01 = get(P2G_F + 1), f02 = get(P2G_F + 2);
  const T f10 = get(P2G_F + 3), f11 = get(P2G_F + 4), f12 = get(P2G_F + 5);
  const T f20 = get(P2G_F + 6), f21 = get(P2G_F + 7), f22 = get(P2G_F + 8);
  const T c00 = get(P2G_C + 0), c01 = get(P2G_C + 1), c02 = get(P2G_C + 2);
  const T c10 = get(P2G_C + 3), c11 = get(P2G_C + 4), c12 = get(P2G_C + 5);
  const T c20 = get(P2G_C + 6), c21 = get(P2G_C + 7), c22 = get(P2G_C + 8);
  bool done = false;
  if constexpr (sizeof(T) == 4) {
    // fp32 build, moderate strain: perturbation-form stress entirely in fp32 (mpm_math.cuh)
    if (cfg.fp32_stress) {
      Mat3<float> Ff{f00, f01, f02, f10, f11, f12, f20, f21, f22}, Cf{c00, c01, c02, c10, c11, c12, c20, c21, c22}, Af;
      done = fixed_corotated_affine3_f32(Ff, Cf, (float)mu, (float)lam, (float)mass, (float)k, Af);
      if (done) {
        q.a00 = Af.a00; q.a01 = Af.a01; q.a02 = Af.a02;
        q.a10 = Af.a10; q.a11 = Af.a11; q.a12 = Af.a12;
        q.a20 = Af.a20; q.a21 = Af.a21; q.a22 = Af.a22;
      }
    }
  }
  if (!done) {
    Mat3<double> F{f00, f01, f02, f10, f11, f12, f20, f21, f22}, C{c00, c01, c02, c10, c11, c12, c20, c21, c22};
    Mat3<double> A = fixed_corotated_affine3(F, C, mu, lam, mass, k);
    q.a00 = (T)A.a00; q.a01 = (T)A.a01; q.a02 = (T)A.a02;
    q.a10 = (T)A.a10; q.a11 = (T)A.a11; q.a12 = (T)A.a12;
    q.a20 = (T)A.a20; q.a21 = (T)A.a21; q.a22 = (T)A.a22;
  }
  q.m = (T)mass;
  if (sizeof(T) == 4 && has_mat) {
    // the product of two floats is exact in fp64, so rounding it once is the fp32 product: same bits, no fp64 pipe
    q.mvx = q.m * get(P2G_V); q.mvy = q.m * get(P2G_V + 1); q.mvz = q.m * get(P2G_V + 2);
  } else {
    q.mvx = (T)(mass * (double)get(P2G_V)); q.mvy = (T)(mass * (double)get(P2G_V + 1)); q.mvz = (T)(mass * (double)get(P2G_V + 2));
  }
  return q;
}

template <typename T>
__device__ __forceinline__ P2GParticle3<T> p2g_prepare3(const DevCfg& cfg, const StateView<T>& s, long long p) {
  const long long st = s.stride;
  const int mode = mat_mode_of(s);
  const int row = (mode == MAT_TABLE && s.material) ? (int)s.material[p] : 0;
  auto get = [&](int k) -> T {
    if (k < P2G_V) return s.x[k * st + p];
    if (k < P2G_C) return s.v[(k - P2G_V) * st + p];
    if (k < P2G_F) return s.C[(k - P2G_C) * st + p];
    if (k < P2G_MASS) return s.F[(k - P2G_F) * st + p];
    if (mode == MAT_TABLE) return s.mat_table[(k - P2G_MASS) * MAT_ROWS + row];
    if (k == P2G_MASS) return s.mass[p];
    if (k == P2G_MU) return s.mu0[p];
    return s.lam0[p];
  };
  const double jp = cfg.model == 1 ? (double)s.Jp[p] : 1.0;
  return p2g_prepare3_from<T>(cfg, get, mode != MAT_CFG, jp);
}

template <typename T>
struct P2GParticle2 {
  int bx, by;
  T fx, fy;
  T mvx, mvy, m;
  T a00, a01, a10, a11;
  bool ok;
};

template <typename T>
__device__ __forceinline__ P2GParticle2<T> p2g_prepare2(const DevCfg& cfg, const StateView<T>& s, long long p) {
  P2GParticle2<T> q;
  const long long st = s.stride;
  T x0 = s.x[p], x1 = s.x[st + p];
  int gx, gy;
  base_fx(x0, cfg, gx, q.fx);
  base_fx(x1, cfg, gy, q.fy);
  q.bx = gx - cfg.origin[0]; q.by = gy - cfg.origin[1];
  // The 2D reference has no bounds check (UB there); we flag and skip instead.
  q.ok = !(isnan((double)x0) || isnan((double)x1)) && q.bx >= 0 && q.by >= 0 && q.bx + 2 < cfg.n[0] && q.by + 2 < cfg.n[1];
  if (!q.ok) return q;
  double mass = cfg.mass, mu = cfg.mu0, lam = cfg.lam0;
  const int mode = mat_mode_of(s);
  if (mode == MAT_TABLE) {
    const int row = s.material ? (int)s.material[p] : 0;
    mass = (double)s.mat_table[row]; mu = (double)s.mat_table[MAT_ROWS + row]; lam = (double)s.mat_table[2 * MAT_ROWS + row];
  } else if (mode == MAT_PLANES) {
    mass = (double)s.mass[p]; mu = (double)s.mu0[p]; lam = (double)s.lam0[p];
  }
  double e = cfg.hardening;
  if (cfg.model == 1) e = exp(cfg.hardening * (1.0 - (double)s.Jp[p]));
  mu *= e; lam *= e;
  const double k = (cfg.dt * cfg.volume) * (4.0 * cfg.inv_dx * cfg.inv_dx);
  const T f00 = s.F[p], f01 = s.F[st + p], f10 = s.F[2 * st + p], f11 = s.F[3 * st + p];
  const T c00 = s.C[p], c01 = s.C[st + p], c10 = s.C[2 * st + p], c11 = s.C[3 * st + p];
  bool done = false;
  if constexpr (sizeof(T) == 4) {
    // fp32 build: the cancellation-free closed form (mpm_math.cuh), the reference's 1e-10 restored analytically
    if (cfg.fp32_stress) {
      Mat2<float> Af;
      done = fixed_corotated_affine2_f32(Mat2<float>{f00, f01, f10, f11}, Mat2<float>{c00, c01, c10, c11}, (float)mu, (float)lam,
                                         (float)mass, (float)k, Af);
      if (done) { q.a00 = Af.a00; q.a01 = Af.a01; q.a10 = Af.a10; q.a11 = Af.a11; }
    }
  }
  if (!done) {
    Mat2<double> F, C;
    F.a00 = f00; F.a01 = f01; F.a10 = f10; F.a11 = f11;
    C.a00 = c00; C.a01 = c01; C.a10 = c10; C.a11 = c11;
    Mat2<double> A = fixed_corotated_affine2(F, C, mu, lam, mass, k);
    q.a00 = (T)A.a00; q.a01 = (T)A.a01; q.a10 = (T)A.a10; q.a11 = (T)A.a11;
  }
  q.m = (T)mass;
  q.mvx = (T)(mass * (double)s.v[p]); q.mvy = (T)(mass * (double)s.v[st + p]);
  return q;
}

// ----------------------------------------------------------------------------
// P2G, scatter form
// ----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) p2g_scatter3_kernel(DevCfg cfg, StateView<T> s, long long n, T* __restrict__ grid,
                                                           ErrRec* err) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  P2GParticle3<T> q = p2g_prepare3(cfg, s, p);
  if (!q.ok) { atomicAdd(&err->n_oob, 1ULL); return; }
  T wx[3], wy[3], wz[3];
  bspline(q.fx, wx[0], wx[1], wx[2]);
  bspline(q.fy, wy[0], wy[1], wy[2]);
  bspline(q.fz, wz[0], wz[1], wz[2]);
  const T dx = (T)cfg.dx;
  const long long ny = cfg.n[1], nz = cfg.n[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T dpx = ((T)i - q.fx) * dx;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T dpy = ((T)j - q.fy) * dx;
      T wij = wx[i] * wy[j];
      T* row = grid + (((long long)(q.bx + i) * ny + (q.by + j)) * nz + q.bz) * 4;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        T dpz = ((T)k - q.fz) * dx;
        T w = wij * wz[k];
        T mx = q.mvx + (q.a00 * dpx + q.a01 * dpy + q.a02 * dpz);
        T my = q.mvy + (q.a10 * dpx + q.a11 * dpy + q.a12 * dpz);
        T mz = q.mvz + (q.a20 * dpx + q.a21 * dpy + q.a22 * dpz);
        red_add4(row + 4 * k, w * mx, w * my, w * mz, w * q.m);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(128) p2g_scatter2_kernel(DevCfg cfg, StateView<T> s, long long n, T* __restrict__ grid,
                                                           ErrRec* err) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  P2GParticle2<T> q = p2g_prepare2(cfg, s, p);
  if (!q.ok) { atomicAdd(&err->n_oob, 1ULL); return; }
  T wx[3], wy[3];
  bspline(q.fx, wx[0], wx[1], wx[2]);
  bspline(q.fy, wy[0], wy[1], wy[2]);
  const T dx = (T)cfg.dx;
  const long long ny = cfg.n[1];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T dpx = ((T)i - q.fx) * dx;
    T* row = grid + ((long long)(q.bx + i) * ny + q.by) * 4;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T dpy = ((T)j - q.fy) * dx;
      T w = wx[i] * wy[j];
      T mx = q.mvx + (q.a00 * dpx + q.a01 * dpy);
      T my = q.mvy + (q.a10 * dpx + q.a11 * dpy);
      red_add4(row + 4 * j, w * mx, w * my, w * q.m, (T)0);
    }
  }
}

// ----------------------------------------------------------------------------
// Grid update fused with the boundary projection
// ----------------------------------------------------------------------------
// 3D (three_d/grid_op.py:25-47): v = mom/mass, v.y += dt*g, clamp to +-0.9 dx/dt,
// then zero component d on nodes with global index I[d] < 1 or I[d] >= R-1 (quirk 5).
template <typename T>
FFMPM_HD void grid_op3_node(const DevCfg& cfg, T* __restrict__ grid, long long n_nodes, long long node,
                                              int i, int j, int k, const T* __restrict__ halo_lo, long long nodes_lo,
                                              const T* __restrict__ halo_hi, long long nodes_hi, const Colliders& col) {
  using V4 = typename Vec4<T>::type;
  V4 g = reinterpret_cast<V4*>(grid)[node];
  // Halo SUM fused into the load: the neighbour slabs' partial {momentum, mass} of the
  // shared node planes (first nodes_lo / last nodes_hi nodes of the C-order grid).
  if (node < nodes_lo) {
    const V4 h = reinterpret_cast<const V4*>(halo_lo)[node];
    g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
  }
  if (node >= n_nodes - nodes_hi) {
    const V4 h = reinterpret_cast<const V4*>(halo_hi)[node - (n_nodes - nodes_hi)];
    g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
  }
  // empty node with zero momentum: velocity stays zero under the wall projection
  if (!(g.w > (T)0) && g.x == (T)0 && g.y == (T)0 && g.z == (T)0) return;
  if (g.w > (T)0) {
    const T va = (T)(cfg.dx * 0.9 / cfg.dt);
    const T dtg = (T)(cfg.dt * cfg.gravity);
    T vx = g.x / g.w, vy = g.y / g.w, vz = g.z / g.w;
    vy += dtg;
    g.x = fmin(fmax(vx, -va), va);
    g.y = fmin(fmax(vy, -va), va);
    g.z = fmin(fmax(vz, -va), va);
  }
  const int boundary = 1;
  int I0 = i + cfg.origin[0], I1 = j + cfg.origin[1], I2 = k + cfg.origin[2];
  if (I0 < boundary || I0 >= cfg.res[0] - boundary) g.x = (T)0;
  if (I1 < boundary || I1 >= cfg.res[1] - boundary) g.y = (T)0;
  if (I2 < boundary || I2 >= cfg.res[2] - boundary) g.z = (T)0;
  // plane colliders (three_d/grid_op.py:50-67), predicate in fp64 like the reference
  for (int c = 0; c < col.count; ++c) {
    const double ox = I0 * cfg.dx - col.point[c][0], oy = I1 * cfg.dx - col.point[c][1], oz = I2 * cfg.dx - col.point[c][2];
    if (ox * col.normal[c][0] + oy * col.normal[c][1] + oz * col.normal[c][2] < 0) { g.x = g.y = g.z = (T)0; }
  }
  reinterpret_cast<V4*>(grid)[node] = g;
}

template <typename T>
__global__ void __launch_bounds__(256) grid_op3_kernel(DevCfg cfg, T* __restrict__ grid, long long n_nodes,
                                                       const T* __restrict__ halo_lo, long long nodes_lo,
                                                       const T* __restrict__ halo_hi, long long nodes_hi, Colliders col) {
  long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  int k = (int)(node % cfg.n[2]);
  long long r = node / cfg.n[2];
  int j = (int)(r % cfg.n[1]);
  int i = (int)(r / cfg.n[1]);
  grid_op3_node<T>(cfg, grid, n_nodes, node, i, j, k, halo_lo, nodes_lo, halo_hi, nodes_hi, col);
}

// The same update over a list of 4x4x4 node blocks (mpm_bin.cuh: node_tiles_kernel) instead of the
// whole grid: every node outside the listed blocks is zero {momentum, mass} by construction (the grid
// was cleared, P2G of the binned particles wrote only inside them), and the dense kernel leaves such
// nodes untouched.  One node per thread, four blocks per CTA round, grid-stride over the list.
template <typename T>
__global__ void __launch_bounds__(256) grid_op3_blocks_kernel(DevCfg cfg, T* __restrict__ grid, long long n_nodes,
                                                              const T* __restrict__ halo_lo, long long nodes_lo,
                                                              const T* __restrict__ halo_hi, long long nodes_hi,
                                                              Colliders col, const int* __restrict__ node_tiles,
                                                              const int* __restrict__ n_listed, int nt1, int nt2) {
  const int count = *n_listed;
  const int local = threadIdx.x & 63;
  const int di = local >> 4, dj = (local >> 2) & 3, dk = local & 3;
  for (int idx = blockIdx.x * 4 + (threadIdx.x >> 6); idx < count; idx += gridDim.x * 4) {
    const int t = node_tiles[idx];
    const int c = t % nt2, b = (t / nt2) % nt1, a = t / (nt2 * nt1);
    const int i = a * 4 + di, j = b * 4 + dj, k = c * 4 + dk;
    if (i < cfg.n[0] && j < cfg.n[1] && k < cfg.n[2]) {
      const long long node = ((long long)i * cfg.n[1] + j) * cfg.n[2] + k;
      grid_op3_node<T>(cfg, grid, n_nodes, node, i, j, k, halo_lo, nodes_lo, halo_hi, nodes_hi, col);
    }
  }
}

// Zeroes the listed node blocks: after a substep nothing else of that grid is non-zero.
template <typename T>
__global__ void __launch_bounds__(256) grid_clear_blocks_kernel(DevCfg cfg, T* __restrict__ grid,
                                                                const int* __restrict__ node_tiles,
                                                                const int* __restrict__ n_listed, int nt1, int nt2) {
  using V4 = typename Vec4<T>::type;
  const int count = *n_listed;
  const int local = threadIdx.x & 63;
  const int di = local >> 4, dj = (local >> 2) & 3, dk = local & 3;
  V4 z;
  z.x = z.y = z.z = z.w = (T)0;
  for (int idx = blockIdx.x * 4 + (threadIdx.x >> 6); idx < count; idx += gridDim.x * 4) {
    const int t = node_tiles[idx];
    const int c = t % nt2, b = (t / nt2) % nt1, a = t / (nt2 * nt1);
    const int i = a * 4 + di, j = b * 4 + dj, k = c * 4 + dk;
    if (i < cfg.n[0] && j < cfg.n[1] && k < cfg.n[2])
      reinterpret_cast<V4*>(grid)[((long long)i * cfg.n[1] + j) * cfg.n[2] + k] = z;
  }
}

// Stand-alone plane colliders (three_d/grid_op.py:50-67) for the phase-level API; inside a
// substep the same predicate runs fused at the end of grid_op3_kernel.
template <typename T>
__global__ void __launch_bounds__(256) collide3_kernel(DevCfg cfg, T* __restrict__ grid, long long n_nodes, Colliders col) {
  long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  const int k = (int)(node % cfg.n[2]);
  const long long r = node / cfg.n[2];
  const int I0 = (int)(r / cfg.n[1]) + cfg.origin[0], I1 = (int)(r % cfg.n[1]) + cfg.origin[1], I2 = k + cfg.origin[2];
  bool hit = false;
  for (int c = 0; c < col.count; ++c) {
    const double ox = I0 * cfg.dx - col.point[c][0], oy = I1 * cfg.dx - col.point[c][1], oz = I2 * cfg.dx - col.point[c][2];
    hit = hit || (ox * col.normal[c][0] + oy * col.normal[c][1] + oz * col.normal[c][2] < 0);
  }
  if (hit) { grid[node * 4 + 0] = (T)0; grid[node * 4 + 1] = (T)0; grid[node * 4 + 2] = (T)0; }
}

// 2D (two_d/grid_op.py:13-24): only nodes with mass > 0; walls are f64 predicates on
// i/R against 0.05 and 1-0.05, evaluated here in f64 exactly as the reference (quirk 6).
// `clear` (optional): a second grid of the same shape that this pass zeroes on the way (the idle one of the ping-pong pair).
template <typename T>
__global__ void __launch_bounds__(256) grid_op2_kernel(DevCfg cfg, T* __restrict__ grid, long long n_nodes, T* __restrict__ clear) {
  long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  using V4 = typename Vec4<T>::type;
  if (clear) {
    V4 z;
    z.x = z.y = z.z = z.w = (T)0;
    reinterpret_cast<V4*>(clear)[node] = z;
  }
  V4 g = reinterpret_cast<V4*>(grid)[node];
  if (!(g.z > (T)0)) return;
  int j = (int)(node % cfg.n[1]);
  int i = (int)(node / cfg.n[1]);
  T vx = g.x / g.z, vy = g.y / g.z;
  vy += (T)(cfg.dt * cfg.gravity);
  const double boundary = 0.05;
  double x = (double)(i + cfg.origin[0]) / (double)cfg.res[0];
  double y = (double)(j + cfg.origin[1]) / (double)cfg.res[1];
  if (x < boundary || x > 1 - boundary || y > 1 - boundary) { vx = (T)0; vy = (T)0; }
  if (y < boundary) vy = fmax((T)0, vy);
  g.x = vx; g.y = vy;
  reinterpret_cast<V4*>(grid)[node] = g;
}

// ----------------------------------------------------------------------------
// G2P stencil sums (three_d/g2p.py:31-43): v = sum w gv,  C = sum (w gv) (x) dpos  (the
// caller scales C by 4*inv_dx).  The weights and dpos are separable per axis, so the 27
// node terms are folded axis by axis (z, then y, then x): ~40% fewer FMAs than the
// node-by-node outer product, same sums up to round-off.
// ----------------------------------------------------------------------------
template <typename T, typename Fetch>
FFMPM_HD void g2p_accumulate3(Fetch fetch, T fx, T fy, T fz, T& vx, T& vy, T& vz, T& c00, T& c01,
                                                T& c02, T& c10, T& c11, T& c12, T& c20, T& c21, T& c22) {
  T wx[3], wy[3], wz[3];
  bspline(fx, wx[0], wx[1], wx[2]);
  bspline(fy, wy[0], wy[1], wy[2]);
  bspline(fz, wz[0], wz[1], wz[2]);
  T dz[3] = {wz[0] * ((T)0 - fz), wz[1] * ((T)1 - fz), wz[2] * ((T)2 - fz)};
  T dy[3] = {wy[0] * ((T)0 - fy), wy[1] * ((T)1 - fy), wy[2] * ((T)2 - fy)};
  T dxw[3] = {wx[0] * ((T)0 - fx), wx[1] * ((T)1 - fx), wx[2] * ((T)2 - fx)};
  vx = vy = vz = (T)0;
  c00 = c01 = c02 = c10 = c11 = c12 = c20 = c21 = c22 = (T)0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T px = 0, py = 0, pz = 0;        // sum_j wy (sum_k wz gv)
    T qx = 0, qy = 0, qz = 0;        // sum_j wy (j - fy) (sum_k wz gv)
    T rx = 0, ry = 0, rz = 0;        // sum_j wy (sum_k wz (k - fz) gv)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T tx = 0, ty = 0, tz = 0, ux = 0, uy = 0, uz = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const auto g = fetch(i, j, k);
        tx += wz[k] * g.x; ty += wz[k] * g.y; tz += wz[k] * g.z;
        ux += dz[k] * g.x; uy += dz[k] * g.y; uz += dz[k] * g.z;
      }
      px += wy[j] * tx; py += wy[j] * ty; pz += wy[j] * tz;
      qx += dy[j] * tx; qy += dy[j] * ty; qz += dy[j] * tz;
      rx += wy[j] * ux; ry += wy[j] * uy; rz += wy[j] * uz;
    }
    vx += wx[i] * px; vy += wx[i] * py; vz += wx[i] * pz;
    c00 += dxw[i] * px; c10 += dxw[i] * py; c20 += dxw[i] * pz;
    c01 += wx[i] * qx; c11 += wx[i] * qy; c21 += wx[i] * qz;
    c02 += wx[i] * rx; c12 += wx[i] * ry; c22 += wx[i] * rz;
  }
}

// ----------------------------------------------------------------------------
// G2P, gather form (in place)
// ----------------------------------------------------------------------------
// KEEPF (3D snow): F is left as it was; snow_project3_kernel forms (I + dt C) F in fp64 and projects it.
template <typename T, bool KEEPF = false>
__global__ void __launch_bounds__(128) g2p_gather3_kernel(DevCfg cfg, StateView<T> s, long long n, const T* __restrict__ grid,
                                                          ErrRec* err) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long st = s.stride;
  T x0 = s.x[p], x1 = s.x[st + p], x2 = s.x[2 * st + p];
  int gx, gy, gz;
  T fx, fy, fz;
  base_fx(x0, cfg, gx, fx);
  base_fx(x1, cfg, gy, fy);
  base_fx(x2, cfg, gz, fz);
  int bx = gx - cfg.origin[0], by = gy - cfg.origin[1], bz = gz - cfg.origin[2];
  bool ok = !(isnan((double)x0) || isnan((double)x1) || isnan((double)x2)) && bx >= 0 && by >= 0 && bz >= 0 &&
            bx + 2 < cfg.n[0] && by + 2 < cfg.n[1] && bz + 2 < cfg.n[2];
  if (!ok) { atomicAdd(&err->n_oob, 1ULL); return; }
  const long long ny = cfg.n[1], nz = cfg.n[2];
  const T* gb = grid + (((long long)bx * ny + by) * nz + bz) * 4;
  T vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22;
  g2p_accumulate3<T>([&](int i, int j, int k) { return ld_node(gb + (((long long)i * ny + j) * nz + k) * 4); }, fx, fy, fz,
                     vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22);
  const T s4 = (T)(4.0 * cfg.inv_dx);
  c00 *= s4; c01 *= s4; c02 *= s4; c10 *= s4; c11 *= s4; c12 *= s4; c20 *= s4; c21 *= s4; c22 *= s4;
  const T dt = (T)cfg.dt;
  T f00 = s.F[0 * st + p], f01 = s.F[1 * st + p], f02 = s.F[2 * st + p];
  T f10 = s.F[3 * st + p], f11 = s.F[4 * st + p], f12 = s.F[5 * st + p];
  T f20 = s.F[6 * st + p], f21 = s.F[7 * st + p], f22 = s.F[8 * st + p];
  // F <- (I + dt C) F   (three_d/g2p.py:46)
  T m00 = (T)1 + dt * c00, m01 = dt * c01, m02 = dt * c02;
  T m10 = dt * c10, m11 = (T)1 + dt * c11, m12 = dt * c12;
  T m20 = dt * c20, m21 = dt * c21, m22 = (T)1 + dt * c22;
  if constexpr (!KEEPF) {
    s.F[0 * st + p] = m00 * f00 + m01 * f10 + m02 * f20;
    s.F[1 * st + p] = m00 * f01 + m01 * f11 + m02 * f21;
    s.F[2 * st + p] = m00 * f02 + m01 * f12 + m02 * f22;
    s.F[3 * st + p] = m10 * f00 + m11 * f10 + m12 * f20;
    s.F[4 * st + p] = m10 * f01 + m11 * f11 + m12 * f21;
    s.F[5 * st + p] = m10 * f02 + m11 * f12 + m12 * f22;
    s.F[6 * st + p] = m20 * f00 + m21 * f10 + m22 * f20;
    s.F[7 * st + p] = m20 * f01 + m21 * f11 + m22 * f21;
    s.F[8 * st + p] = m20 * f02 + m21 * f12 + m22 * f22;
  }
  s.C[0 * st + p] = c00; s.C[1 * st + p] = c01; s.C[2 * st + p] = c02;
  s.C[3 * st + p] = c10; s.C[4 * st + p] = c11; s.C[5 * st + p] = c12;
  s.C[6 * st + p] = c20; s.C[7 * st + p] = c21; s.C[8 * st + p] = c22;
  s.v[p] = vx; s.v[st + p] = vy; s.v[2 * st + p] = vz;
  s.x[p] = x0 + dt * vx; s.x[st + p] = x1 + dt * vy; s.x[2 * st + p] = x2 + dt * vz;
}

// ----------------------------------------------------------------------------
// G2P of one 2D particle (two_d/g2p.py:17-47): shared by the in-place gather kernel and the reordering one.
// ----------------------------------------------------------------------------
template <typename T>
struct G2POut2 {
  T x0, x1, v0, v1, c00, c01, c10, c11, f00, f01, f10, f11, jp;
};

template <typename T>
__device__ __forceinline__ void g2p_particle2(const DevCfg& cfg, const T* __restrict__ grid, int bx, int by, T fx, T fy, T x0, T x1,
                                              T f00, T f01, T f10, T f11, T jp_in, bool has_jp, G2POut2<T>& o) {
  T wx[3], wy[3];
  bspline(fx, wx[0], wx[1], wx[2]);
  bspline(fy, wy[0], wy[1], wy[2]);
  const long long ny = cfg.n[1];
  T vx = 0, vy = 0, c00 = 0, c01 = 0, c10 = 0, c11 = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const T dpx = (T)i - fx;
    const T* row = grid + ((long long)(bx + i) * ny + by) * 4;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const T dpy = (T)j - fy;
      const T w = wx[i] * wy[j];
      const auto g = ld_node(row + 4 * j);
      const T ux = w * g.x, uy = w * g.y;
      vx += ux; vy += uy;
      c00 += ux * dpx; c01 += ux * dpy; c10 += uy * dpx; c11 += uy * dpy;
    }
  }
  const T s4 = (T)(4.0 * cfg.inv_dx);
  c00 *= s4; c01 *= s4; c10 *= s4; c11 *= s4;
  const T dt = (T)cfg.dt;
  const T m00 = (T)1 + dt * c00, m01 = dt * c01, m10 = dt * c10, m11 = (T)1 + dt * c11;
  const T n00 = m00 * f00 + m01 * f10, n01 = m00 * f01 + m01 * f11, n10 = m10 * f00 + m11 * f10, n11 = m10 * f01 + m11 * f11;
  o.jp = jp_in;
  bool done = false;
  if constexpr (sizeof(T) == 4) {
    // two_d/g2p.py:37-47 for det F > 0 without plasticity: U diag(sig) Vh^T is F itself (LAPACK's 2x2 Vh is symmetric
    // there), and Jp <- clip(Jp * J / (J + 1e-10)): no SVD, no fp64 in the fp32 build.  det F <= 0
    // (the V.T quirk rotates F) and snow (singular values clamped) take the closed forms of svd_roundtrip2 in fp64.
    // "safely positive": the fp32 determinant cannot have the wrong sign when it exceeds 1e-3 of its own terms
    const float p0 = n00 * n11, p1 = n01 * n10;
    const float old_J = p0 - p1;
    if (cfg.model != 1 && old_J > 1e-3f * (fabsf(p0) + fabsf(p1))) {
      o.f00 = n00; o.f01 = n01; o.f10 = n10; o.f11 = n11;
      if (has_jp) o.jp = fminf(fmaxf(jp_in * (old_J / (old_J + 1e-10f)), 0.6f), 20.0f);
      done = true;
    }
  }
  if (!done) {
    Mat2<double> Fn;
    Fn.a00 = n00; Fn.a01 = n01; Fn.a10 = n10; Fn.a11 = n11;
    // two_d/g2p.py:37-47: SVD round trip and Jp update for every model (quirk 7)
    const double old_J = Fn.a00 * Fn.a11 - Fn.a01 * Fn.a10;
    double det_new;
    const Mat2<double> Fr = svd_roundtrip2(Fn, cfg.model == 1, det_new);
    if (has_jp) {
      const double jp = (double)jp_in * old_J / (det_new + 1e-10);
      o.jp = (T)fmin(fmax(jp, 0.6), 20.0);
    }
    o.f00 = (T)Fr.a00; o.f01 = (T)Fr.a01; o.f10 = (T)Fr.a10; o.f11 = (T)Fr.a11;
  }
  o.c00 = c00; o.c01 = c01; o.c10 = c10; o.c11 = c11;
  o.v0 = vx; o.v1 = vy;
  o.x0 = x0 + dt * vx; o.x1 = x1 + dt * vy;
}

template <typename T>
__global__ void __launch_bounds__(128) g2p_gather2_kernel(DevCfg cfg, StateView<T> s, long long n, const T* __restrict__ grid,
                                                          ErrRec* err) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long st = s.stride;
  T x0 = s.x[p], x1 = s.x[st + p];
  int gx, gy;
  T fx, fy;
  base_fx(x0, cfg, gx, fx);
  base_fx(x1, cfg, gy, fy);
  int bx = gx - cfg.origin[0], by = gy - cfg.origin[1];
  bool ok = !(isnan((double)x0) || isnan((double)x1)) && bx >= 0 && by >= 0 && bx + 2 < cfg.n[0] && by + 2 < cfg.n[1];
  if (!ok) { atomicAdd(&err->n_oob, 1ULL); return; }
  G2POut2<T> o;
  g2p_particle2<T>(cfg, grid, bx, by, fx, fy, x0, x1, s.F[p], s.F[st + p], s.F[2 * st + p], s.F[3 * st + p],
                   s.Jp ? s.Jp[p] : (T)1, s.Jp != nullptr, o);
  if (s.Jp) s.Jp[p] = o.jp;
  s.F[p] = o.f00; s.F[st + p] = o.f01; s.F[2 * st + p] = o.f10; s.F[3 * st + p] = o.f11;
  s.C[p] = o.c00; s.C[st + p] = o.c01; s.C[2 * st + p] = o.c10; s.C[3 * st + p] = o.c11;
  s.v[p] = o.v0; s.v[st + p] = o.v1;
  s.x[p] = o.x0; s.x[st + p] = o.x1;
}

}  // namespace ffmpm

// The binned 2D pipeline: the architecture of the 3D one for two_d/{p2g,grid_op,g2p}.py, selectable with
// MpmSolver(reorder=True).  At BASELINE configs[1] (1 M particles on 1024^2: 54 MB of state, resident in L2) it is SLOWER
// than the unbinned kernels (86.8 us against 43.0 us per substep for the warp-window kernels of mpm_2d_window.cuh,
// profiles/r02s_2d_series.json), so those stay the 2D default; this path is for scenes whose state does not fit L2.
//   p2g_runs2_kernel     warp-autonomous P2G in physical order (two_d/p2g.py:49-76): lane per particle -> runs of equal
//                        base cell -> lane per (run, x-slab) accumulating the slab's three nodes in registers -> ONE
//                        vector RED per node and run;
//   g2p_reorder2_kernel  thread per binned slot (two_d/g2p.py:17-47 incl. the SVD round trip and Jp): gathers its
//                        particle through `perm`, reads the 9 nodes straight from the grid, writes the new state in cell
//                        order into the other buffer and emits the next substep's key, rank and histogram;
//   node_tiles2 / grid_op2_blocks / grid_clear_blocks2   grid update and clear over the 8x8-node blocks the binned
//                        particles can have written.
// (Also measured and removed, profiles/r02g / r02h: the whole 2D substep loop as ONE persistent cooperative kernel with
// three grid-wide barriers per substep -- 79-81 us per substep, slower than the four separate kernels.)
#pragma once
#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"
#include "mpm_p2g_runs.cuh"

namespace ffmpm {

constexpr int NODE_TILE2 = 8;   // nodes per block edge, 2D (= TILE2: a tile of base cells 8t .. 8t+7 scatters into blocks t, t+1)

// ---------------------------------------------------------------------------------------------------------------
// P2G
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct P2GWarpSlab2 {
  P2GVec4<T> pay[2][P2G_PADDED];   // {mvx, mvy, m, fx}  {a00, a01, a10, a11} * dx
  T fy[P2G_PADDED];
  int node0[P2G_WINDOW];           // linear LOCAL node id of the base cell (-1: outside the grid)
  int run_start[P2G_WINDOW + 1];
};

template <typename T>
__global__ void __launch_bounds__(P2G_RUN_WARPS * 32, 4)
p2g_runs2_kernel(DevCfg cfg, StateView<T> s, long long n, T* __restrict__ grid, ErrRec* err) {
  __shared__ P2GWarpSlab2<T> slabs[P2G_RUN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  P2GWarpSlab2<T>& S = slabs[warp];
  const T dx = (T)cfg.dx;
  const int ny = cfg.n[1];
  const int n_windows = (int)((n + P2G_WINDOW - 1) / P2G_WINDOW);
  const int total_warps = gridDim.x * P2G_RUN_WARPS;
  for (int win = blockIdx.x * P2G_RUN_WARPS + warp; win < n_windows; win += total_warps) {
    const long long w0 = (long long)win * P2G_WINDOW;
    const int cnt = (int)min((long long)P2G_WINDOW, n - w0);
    int node[2];
    // ---- phase 1: lane per particle (slots lane and 32 + lane) ----
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = h * 32 + lane;
      node[h] = -1;
      if (idx < cnt) {
        P2GParticle2<T> q = p2g_prepare2(cfg, s, w0 + idx);
        const int ph = p2g_pad(idx);
        if (q.ok) {
          S.pay[0][ph] = P2GVec4<T>{q.mvx, q.mvy, q.m, q.fx};
          S.pay[1][ph] = P2GVec4<T>{q.a00 * dx, q.a01 * dx, q.a10 * dx, q.a11 * dx};
          S.fy[ph] = q.fy;
          node[h] = q.bx * ny + q.by;
        }
        S.node0[idx] = node[h];
      }
    }
    __syncwarp();
    // ---- runs ----
    unsigned heads[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = h * 32 + lane;
      const int prev = idx > 0 ? S.node0[idx - 1] : -2;
      heads[h] = __ballot_sync(0xffffffffu, idx < cnt && node[h] != prev);
    }
    const int n0 = __popc(heads[0]);
    const int n_runs = n0 + __popc(heads[1]);
    {
      const unsigned below = (1u << lane) - 1u;
      if (heads[0] & (1u << lane)) S.run_start[__popc(heads[0] & below)] = lane;
      if (heads[1] & (1u << lane)) S.run_start[n0 + __popc(heads[1] & below)] = 32 + lane;
      if (lane == 0) S.run_start[n_runs] = cnt;
    }
    __syncwarp();
    // ---- phase 2: lane per (run, x-slab) ----
    const int n_items = n_runs * 3;
    for (int item = lane; item < n_items; item += 32) {
      const int r = item / 3, li = item - r * 3;
      const int r0 = S.run_start[r], r1 = S.run_start[r + 1];
      const int nd = S.node0[r0];
      if (nd < 0) continue;
      const T ci = (T)li;
      const T sx = li == 1 ? (T)-1 : (T)0.5, cx_ = (T)1.5 - (T)0.5 * ci, ox_ = li == 1 ? (T)0.75 : (T)0;
      T mx[3] = {0, 0, 0}, my[3] = {0, 0, 0}, mm[3] = {0, 0, 0};
      for (int qi = r0; qi < r1; ++qi) {
        const int ph = p2g_pad(qi);
        const P2GVec4<T> p0 = S.pay[0][ph], p1 = S.pay[1][ph];
        const T fy = S.fy[ph], fx = p0.w;
        T wy[3];
        bspline(fy, wy[0], wy[1], wy[2]);
        const T tx_ = fx - cx_;
        const T wxi = sx * tx_ * tx_ + ox_;
        const T dpx = ci - fx;
        const T bx = p0.x + p1.x * dpx, by = p0.y + p1.z * dpx;   // m v + affine[:, 0] * dpos.x  (two_d/p2g.py:72-75)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const T dpy = (T)j - fy;
          const T w = wxi * wy[j];
          mx[j] += w * (bx + p1.y * dpy);
          my[j] += w * (by + p1.w * dpy);
          mm[j] += w * p0.z;
        }
      }
      T* g = grid + ((long long)nd + (long long)li * ny) * 4;
#pragma unroll
      for (int j = 0; j < 3; ++j) red_add4(g + 4 * j, mx[j], my[j], mm[j], (T)0);
    }
    __syncwarp();   // the slab is rewritten by the next window
  }
}

template <typename T>
__global__ void __launch_bounds__(256) g2p_reorder2_kernel(DevCfg cfg, StateView<T> src, StateView<T> dst, long long n, BinBuffers B,
                                                           const T* __restrict__ grid, ErrRec* err) {
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ss = src.stride, ds = dst.stride;
  int next_key = -1;
  if (slot < n) {
    const long long p = B.perm[slot];
    const T x0 = src.x[p], x1 = src.x[ss + p];
    int gx, gy;
    T fx, fy;
    base_fx(x0, cfg, gx, fx);
    base_fx(x1, cfg, gy, fy);
    const int bx = gx - cfg.origin[0], by = gy - cfg.origin[1];
    const bool ok = x0 == x0 && x1 == x1 && bx >= 0 && by >= 0 && bx + 2 < cfg.n[0] && by + 2 < cfg.n[1];
    const T jp_in = src.Jp ? src.Jp[p] : (T)1;
    if (ok) {
      G2POut2<T> o;
      g2p_particle2<T>(cfg, grid, bx, by, fx, fy, x0, x1, src.F[p], src.F[ss + p], src.F[2 * ss + p], src.F[3 * ss + p], jp_in,
                       src.Jp != nullptr, o);
      dst.x[slot] = o.x0; dst.x[ds + slot] = o.x1;
      dst.v[slot] = o.v0; dst.v[ds + slot] = o.v1;
      dst.C[slot] = o.c00; dst.C[ds + slot] = o.c01; dst.C[2 * ds + slot] = o.c10; dst.C[3 * ds + slot] = o.c11;
      dst.F[slot] = o.f00; dst.F[ds + slot] = o.f01; dst.F[2 * ds + slot] = o.f10; dst.F[3 * ds + slot] = o.f11;
      if (src.Jp) dst.Jp[slot] = o.jp;
      next_key = bin_key_of<T>(cfg, B, o.x0, o.x1, (T)0);
    } else {
      // outside the grid (already flagged by the binning): carried over unchanged, stays in the trailing bin
      dst.x[slot] = x0; dst.x[ds + slot] = x1;
      dst.v[slot] = src.v[p]; dst.v[ds + slot] = src.v[ss + p];
      for (int c = 0; c < 4; ++c) { dst.C[c * ds + slot] = src.C[c * ss + p]; dst.F[c * ds + slot] = src.F[c * ss + p]; }
      if (src.Jp) dst.Jp[slot] = jp_in;
      next_key = B.n_cells;
    }
    if (src.mass) dst.mass[slot] = src.mass[p];
    if (src.mu0) dst.mu0[slot] = src.mu0[p];
    if (src.lam0) dst.lam0[slot] = src.lam0[p];
    if (src.id) dst.id[slot] = src.id[p];
    if (src.material) dst.material[slot] = src.material[p];
    B.keys[slot] = next_key;
  }
  // next substep's histogram + within-cell rank (all 32 lanes take part; key < 0 = idle lane)
  bin_rank_warp(B, next_key, slot);
}

// ---------------------------------------------------------------------------------------------------------------
// Node blocks (8 x 8 nodes)
// ---------------------------------------------------------------------------------------------------------------
// Block (a, b) is live iff one of the base-cell tiles (a - {0,1}, b - {0,1}) holds particles: a tile of base cells
// 8t .. 8t+7 scatters to nodes 8t .. 8t+9, i.e. into blocks t and t+1.
__global__ void __launch_bounds__(256) node_tiles2_kernel(BinBuffers B) {
  const int T_ = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  if (T_ < B.n_node_tiles) {
    const int b = T_ % B.ntile[1], a = T_ / B.ntile[1];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int ta = a - (d >> 1), tb = b - (d & 1);
      if (ta >= 0 && tb >= 0 && ta < B.tiles[0] && tb < B.tiles[1]) live = live || B.tile_flag[ta * B.tiles[1] + tb] != 0;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, live);
  if (m) {
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(B.node_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) B.node_tiles[base + __popc(m & ((1u << lane) - 1u))] = T_;
  }
}

// two_d/grid_op.py:13-24 on one node; the walls are f64 predicates on i/R (quirk 6).
template <typename T>
__device__ __forceinline__ void grid_op2_node(const DevCfg& cfg, T* __restrict__ grid, long long node, int i, int j) {
  using V4 = typename Vec4<T>::type;
  V4 g = reinterpret_cast<V4*>(grid)[node];
  if (!(g.z > (T)0)) return;
  T vx = g.x / g.z, vy = g.y / g.z;
  vy += (T)(cfg.dt * cfg.gravity);
  const double boundary = 0.05;
  const double x = (double)(i + cfg.origin[0]) / (double)cfg.res[0];
  const double y = (double)(j + cfg.origin[1]) / (double)cfg.res[1];
  if (x < boundary || x > 1 - boundary || y > 1 - boundary) { vx = (T)0; vy = (T)0; }
  if (y < boundary) vy = fmax((T)0, vy);
  g.x = vx; g.y = vy;
  reinterpret_cast<V4*>(grid)[node] = g;
}

template <typename T>
__global__ void __launch_bounds__(256) grid_op2_blocks_kernel(DevCfg cfg, T* __restrict__ grid, const int* __restrict__ node_tiles,
                                                              const int* __restrict__ n_listed, int nt1) {
  const int count = *n_listed;
  const int local = threadIdx.x & 63;
  const int di = local >> 3, dj = local & 7;
  for (int idx = blockIdx.x * 4 + (threadIdx.x >> 6); idx < count; idx += gridDim.x * 4) {
    const int t = node_tiles[idx];
    const int i = (t / nt1) * NODE_TILE2 + di, j = (t % nt1) * NODE_TILE2 + dj;
    if (i < cfg.n[0] && j < cfg.n[1]) grid_op2_node<T>(cfg, grid, (long long)i * cfg.n[1] + j, i, j);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) grid_clear_blocks2_kernel(DevCfg cfg, T* __restrict__ grid, const int* __restrict__ node_tiles,
                                                                 const int* __restrict__ n_listed, int nt1) {
  using V4 = typename Vec4<T>::type;
  const int count = *n_listed;
  const int local = threadIdx.x & 63;
  const int di = local >> 3, dj = local & 7;
  V4 z;
  z.x = z.y = z.z = z.w = (T)0;
  for (int idx = blockIdx.x * 4 + (threadIdx.x >> 6); idx < count; idx += gridDim.x * 4) {
    const int t = node_tiles[idx];
    const int i = (t / nt1) * NODE_TILE2 + di, j = (t % nt1) * NODE_TILE2 + dj;
    if (i < cfg.n[0] && j < cfg.n[1]) reinterpret_cast<V4*>(grid)[(long long)i * cfg.n[1] + j] = z;
  }
}

}  // namespace ffmpm

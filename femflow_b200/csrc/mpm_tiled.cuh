// Tiled 3D kernels over binned particles (see mpm_bin.cuh for the key layout).
//
// P2G: persistent CTAs pull active tiles (4x4x4 base cells -> 6x6x6 nodes) from a
// work counter.  Each warp takes 32 consecutive binned particles: lanes first
// evaluate the per-particle payload (fp64 polar decomposition + fixed-corotated
// stress, three_d/p2g.py:57-65) and park it in shared memory; then the warp walks
// those 32 particles with ONE LANE PER STENCIL NODE (27 of 32 lanes), accumulating
// the contributions of all particles of the same cell in registers -- particles are
// cell-sorted, so same-node contributions are aggregated across the warp without a
// single conflict -- and flushes into the CTA's shared-memory tile only when the
// cell changes.  The tile is finally added to the global grid with one vector
// reduction (red.global.add.v4.f32) per touched node.
//
// G2P: the CTA stages the tile's 6x6x6 velocity block in shared memory, each thread
// gathers its particle's 3x3x3 stencil from there, and the updated state is written
// to the OTHER particle buffer at the binned slot -- so the state is physically in
// cell order for the next substep and every store is fully coalesced.
#pragma once
#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"

namespace ffmpm {

constexpr int TN3 = TILE3 + 2;            // nodes per tile edge
constexpr int TNODES3 = TN3 * TN3 * TN3;  // 216
constexpr int P2G_WARPS = 8;
constexpr int G2P_THREADS = 128;

template <typename T>
struct alignas(16) P2GPayload {
  T mvx, mvy, mvz, m;
  T a00, a01, a02, fx;
  T a10, a11, a12, fy;
  T a20, a21, a22, fz;
};

template <typename T>
__global__ void __launch_bounds__(P2G_WARPS * 32) p2g_tiled3_kernel(DevCfg cfg, StateView<T> s, BinBuffers B,
                                                                    T* __restrict__ grid, ErrRec* err) {
  __shared__ T tile[4][TNODES3];
  __shared__ P2GPayload<T> stage[P2G_WARPS][32];
  __shared__ int stage_cell[P2G_WARPS][32];
  __shared__ int s_work;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_active = B.counters[0];
  // lane -> stencil node (i, j, k)
  const int li = lane / 9, lj = (lane / 3) % 3, lk = lane % 3;
  const bool node_lane = lane < 27;
  const T ci = (T)li, cj = (T)lj, ck = (T)lk;
  // B-spline piece of this lane per axis: w = s * (fx - c)^2 + o
  const T sx = li == 1 ? (T)-1 : (T)0.5, cx_ = (T)1.5 - (T)0.5 * ci, ox_ = li == 1 ? (T)0.75 : (T)0;
  const T sy = lj == 1 ? (T)-1 : (T)0.5, cy_ = (T)1.5 - (T)0.5 * cj, oy_ = lj == 1 ? (T)0.75 : (T)0;
  const T sz = lk == 1 ? (T)-1 : (T)0.5, cz_ = (T)1.5 - (T)0.5 * ck, oz_ = lk == 1 ? (T)0.75 : (T)0;
  const int lane_node = (li * TN3 + lj) * TN3 + lk;
  const T dx = (T)cfg.dx;

  for (;;) {
    if (threadIdx.x == 0) s_work = atomicAdd(&B.counters[1], 1);
    for (int i = threadIdx.x; i < 4 * TNODES3; i += blockDim.x) (&tile[0][0])[i] = (T)0;
    __syncthreads();
    const int wi = s_work;
    if (wi >= n_active) break;
    const int t = B.active_tiles[wi];
    const int start = B.cell_off[t * TILE_CELLS], end = B.cell_off[(t + 1) * TILE_CELLS];
    const int tz = t % B.tiles[2], ty = (t / B.tiles[2]) % B.tiles[1], tx = t / (B.tiles[2] * B.tiles[1]);
    const int ox = tx * TILE3, oy = ty * TILE3, oz = tz * TILE3;

    for (int chunk = start + warp * 32; chunk < end; chunk += P2G_WARPS * 32) {
      const int slot = chunk + lane;
      // ---- phase 1: one lane per particle ----
      if (slot < end) {
        const long long p = B.perm[slot];
        P2GParticle3<T> q = p2g_prepare3(cfg, s, p);
        P2GPayload<T> pl;
        pl.mvx = q.mvx; pl.mvy = q.mvy; pl.mvz = q.mvz; pl.m = q.m;
        pl.a00 = q.a00 * dx; pl.a01 = q.a01 * dx; pl.a02 = q.a02 * dx; pl.fx = q.fx;
        pl.a10 = q.a10 * dx; pl.a11 = q.a11 * dx; pl.a12 = q.a12 * dx; pl.fy = q.fy;
        pl.a20 = q.a20 * dx; pl.a21 = q.a21 * dx; pl.a22 = q.a22 * dx; pl.fz = q.fz;
        stage[warp][lane] = pl;
        // node index of the cell's base inside the tile
        stage_cell[warp][lane] = ((q.bx - ox) * TN3 + (q.by - oy)) * TN3 + (q.bz - oz);
      }
      __syncwarp();
      // ---- phase 2: one lane per stencil node ----
      const int cnt = min(32, end - chunk);
      T ax = 0, ay = 0, az = 0, am = 0;
      int cur = stage_cell[warp][0];
      for (int qi = 0; qi < cnt; ++qi) {
        const int c = stage_cell[warp][qi];
        if (c != cur) {
          if (node_lane) {
            const int nd = cur + lane_node;
            atomicAdd(&tile[0][nd], ax); atomicAdd(&tile[1][nd], ay);
            atomicAdd(&tile[2][nd], az); atomicAdd(&tile[3][nd], am);
          }
          ax = ay = az = am = (T)0;
          cur = c;
        }
        const P2GPayload<T>& pl = stage[warp][qi];
        const T tx_ = pl.fx - cx_, ty_ = pl.fy - cy_, tz_ = pl.fz - cz_;
        const T wx = sx * tx_ * tx_ + ox_, wy = sy * ty_ * ty_ + oy_, wz = sz * tz_ * tz_ + oz_;
        const T w = wx * wy * wz;
        const T dpx = ci - pl.fx, dpy = cj - pl.fy, dpz = ck - pl.fz;
        const T mx = pl.mvx + (pl.a00 * dpx + pl.a01 * dpy + pl.a02 * dpz);
        const T my = pl.mvy + (pl.a10 * dpx + pl.a11 * dpy + pl.a12 * dpz);
        const T mz = pl.mvz + (pl.a20 * dpx + pl.a21 * dpy + pl.a22 * dpz);
        ax += w * mx; ay += w * my; az += w * mz; am += w * pl.m;
      }
      if (node_lane) {
        const int nd = cur + lane_node;
        atomicAdd(&tile[0][nd], ax); atomicAdd(&tile[1][nd], ay);
        atomicAdd(&tile[2][nd], az); atomicAdd(&tile[3][nd], am);
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- tile -> global grid: one vector reduction per touched node ----
    for (int nd = threadIdx.x; nd < TNODES3; nd += blockDim.x) {
      const T m = tile[3][nd];
      if (m != (T)0) {
        const int k = nd % TN3, j = (nd / TN3) % TN3, i = nd / (TN3 * TN3);
        const int gx = ox + i, gy = oy + j, gz = oz + k;
        if (gx < cfg.n[0] && gy < cfg.n[1] && gz < cfg.n[2]) {
          T* g = grid + (((long long)gx * cfg.n[1] + gy) * cfg.n[2] + gz) * 4;
          red_add4(g, tile[0][nd], tile[1][nd], tile[2][nd], m);
        }
      }
    }
    __syncthreads();
  }
}

// Moves the optional planes that G2P does not compute.
template <typename T>
__device__ __forceinline__ void carry_planes(const StateView<T>& src, const StateView<T>& dst, long long p, long long slot,
                                             bool with_jp) {
  if (src.mass) dst.mass[slot] = src.mass[p];
  if (src.mu0) dst.mu0[slot] = src.mu0[p];
  if (src.lam0) dst.lam0[slot] = src.lam0[p];
  if (src.id) dst.id[slot] = src.id[p];
  if (with_jp && src.Jp) dst.Jp[slot] = src.Jp[p];
}

template <typename T>
__global__ void __launch_bounds__(G2P_THREADS) g2p_tiled3_kernel(DevCfg cfg, StateView<T> src, StateView<T> dst,
                                                                 BinBuffers B, const T* __restrict__ grid, ErrRec* err) {
  using V4 = typename Vec4<T>::type;
  __shared__ V4 tile[TNODES3];
  __shared__ int s_work;
  const int n_active = B.counters[0];
  const long long ss = src.stride, ds = dst.stride;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_work = atomicAdd(&B.counters[2], 1);
    __syncthreads();
    const int wi = s_work;
    if (wi > n_active) break;
    if (wi == n_active) {
      // trailing bin: particles outside the grid are carried over unchanged (already flagged by ffmpm_bin)
      const int start = B.cell_off[B.n_cells], end = B.cell_off[B.n_cells + 1];
      for (int slot = start + threadIdx.x; slot < end; slot += blockDim.x) {
        const long long p = B.perm[slot];
        for (int c = 0; c < 3; ++c) { dst.x[c * ds + slot] = src.x[c * ss + p]; dst.v[c * ds + slot] = src.v[c * ss + p]; }
        for (int c = 0; c < 9; ++c) { dst.C[c * ds + slot] = src.C[c * ss + p]; dst.F[c * ds + slot] = src.F[c * ss + p]; }
        carry_planes(src, dst, p, slot, true);
      }
      continue;
    }
    const int t = B.active_tiles[wi];
    const int start = B.cell_off[t * TILE_CELLS], end = B.cell_off[(t + 1) * TILE_CELLS];
    const int tz = t % B.tiles[2], ty = (t / B.tiles[2]) % B.tiles[1], tx = t / (B.tiles[2] * B.tiles[1]);
    const int ox = tx * TILE3, oy = ty * TILE3, oz = tz * TILE3;
    // ---- stage the tile's velocity block ----
    for (int nd = threadIdx.x; nd < TNODES3; nd += blockDim.x) {
      const int k = nd % TN3, j = (nd / TN3) % TN3, i = nd / (TN3 * TN3);
      const int gx = ox + i, gy = oy + j, gz = oz + k;
      V4 g;
      g.x = g.y = g.z = g.w = (T)0;
      if (gx < cfg.n[0] && gy < cfg.n[1] && gz < cfg.n[2])
        g = ld_node(grid + (((long long)gx * cfg.n[1] + gy) * cfg.n[2] + gz) * 4);
      tile[nd] = g;
    }
    __syncthreads();
    for (int slot = start + threadIdx.x; slot < end; slot += blockDim.x) {
      const long long p = B.perm[slot];
      const T x0 = src.x[p], x1 = src.x[ss + p], x2 = src.x[2 * ss + p];
      int gx, gy, gz;
      T fx, fy, fz;
      base_fx(x0, cfg.inv_dx, gx, fx);
      base_fx(x1, cfg.inv_dx, gy, fy);
      base_fx(x2, cfg.inv_dx, gz, fz);
      const int cb = ((gx - cfg.origin[0] - ox) * TN3 + (gy - cfg.origin[1] - oy)) * TN3 + (gz - cfg.origin[2] - oz);
      T wx[3], wy[3], wz[3];
      bspline(fx, wx[0], wx[1], wx[2]);
      bspline(fy, wy[0], wy[1], wy[2]);
      bspline(fz, wz[0], wz[1], wz[2]);
      T vx = 0, vy = 0, vz = 0;
      T c00 = 0, c01 = 0, c02 = 0, c10 = 0, c11 = 0, c12 = 0, c20 = 0, c21 = 0, c22 = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const T dpx = (T)i - fx;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const T dpy = (T)j - fy;
          const T wij = wx[i] * wy[j];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const T dpz = (T)k - fz;
            const T w = wij * wz[k];
            const V4 g = tile[cb + (i * TN3 + j) * TN3 + k];
            const T ux = w * g.x, uy = w * g.y, uz = w * g.z;
            vx += ux; vy += uy; vz += uz;
            c00 += ux * dpx; c01 += ux * dpy; c02 += ux * dpz;
            c10 += uy * dpx; c11 += uy * dpy; c12 += uy * dpz;
            c20 += uz * dpx; c21 += uz * dpy; c22 += uz * dpz;
          }
        }
      }
      const T s4 = (T)(4.0 * cfg.inv_dx);
      c00 *= s4; c01 *= s4; c02 *= s4; c10 *= s4; c11 *= s4; c12 *= s4; c20 *= s4; c21 *= s4; c22 *= s4;
      const T dt = (T)cfg.dt;
      const T f00 = src.F[0 * ss + p], f01 = src.F[1 * ss + p], f02 = src.F[2 * ss + p];
      const T f10 = src.F[3 * ss + p], f11 = src.F[4 * ss + p], f12 = src.F[5 * ss + p];
      const T f20 = src.F[6 * ss + p], f21 = src.F[7 * ss + p], f22 = src.F[8 * ss + p];
      const T m00 = (T)1 + dt * c00, m01 = dt * c01, m02 = dt * c02;
      const T m10 = dt * c10, m11 = (T)1 + dt * c11, m12 = dt * c12;
      const T m20 = dt * c20, m21 = dt * c21, m22 = (T)1 + dt * c22;
      dst.F[0 * ds + slot] = m00 * f00 + m01 * f10 + m02 * f20;
      dst.F[1 * ds + slot] = m00 * f01 + m01 * f11 + m02 * f21;
      dst.F[2 * ds + slot] = m00 * f02 + m01 * f12 + m02 * f22;
      dst.F[3 * ds + slot] = m10 * f00 + m11 * f10 + m12 * f20;
      dst.F[4 * ds + slot] = m10 * f01 + m11 * f11 + m12 * f21;
      dst.F[5 * ds + slot] = m10 * f02 + m11 * f12 + m12 * f22;
      dst.F[6 * ds + slot] = m20 * f00 + m21 * f10 + m22 * f20;
      dst.F[7 * ds + slot] = m20 * f01 + m21 * f11 + m22 * f21;
      dst.F[8 * ds + slot] = m20 * f02 + m21 * f12 + m22 * f22;
      dst.C[0 * ds + slot] = c00; dst.C[1 * ds + slot] = c01; dst.C[2 * ds + slot] = c02;
      dst.C[3 * ds + slot] = c10; dst.C[4 * ds + slot] = c11; dst.C[5 * ds + slot] = c12;
      dst.C[6 * ds + slot] = c20; dst.C[7 * ds + slot] = c21; dst.C[8 * ds + slot] = c22;
      dst.v[slot] = vx; dst.v[ds + slot] = vy; dst.v[2 * ds + slot] = vz;
      dst.x[slot] = x0 + dt * vx; dst.x[ds + slot] = x1 + dt * vy; dst.x[2 * ds + slot] = x2 + dt * vz;
      carry_planes(src, dst, p, slot, true);
    }
  }
}

template <typename T>
int p2g_tiled(const DevCfg& cfg, const StateView<T>& s, long long n, BinBuffers& B, T* grid, ErrRec* err, int sm_count,
              cudaStream_t st) {
  (void)n;
  cudaMemsetAsync(&B.counters[1], 0, sizeof(int32_t), st);
  int blocks = min(B.n_tiles, sm_count * 4);
  p2g_tiled3_kernel<T><<<blocks, P2G_WARPS * 32, 0, st>>>(cfg, s, B, grid, err);
  return 1;
}

template <typename T>
int g2p_tiled(const DevCfg& cfg, const StateView<T>& src, const StateView<T>& dst, long long n, BinBuffers& B,
              const T* grid, ErrRec* err, int sm_count, cudaStream_t st) {
  (void)n;
  cudaMemsetAsync(&B.counters[2], 0, sizeof(int32_t), st);
  int blocks = min(B.n_tiles + 1, sm_count * 8);
  g2p_tiled3_kernel<T><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);
  return 1;
}

}  // namespace ffmpm

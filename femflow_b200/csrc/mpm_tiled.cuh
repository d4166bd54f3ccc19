// Tiled 3D kernels over binned particles (see mpm_bin.cuh for the key layout).
//
// P2G: see p2g_tiled3_kernel below.
//
// G2P: the CTA stages the tile's 6x6x6 velocity block in shared memory, each thread
// gathers its particle's 3x3x3 stencil from there, and the updated state is written
// to the OTHER particle buffer at the binned slot -- so the state is physically in
// cell order for the next substep and every store is fully coalesced.
#pragma once
#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"
#include "mpm_p2g_runs.cuh"

namespace ffmpm {

constexpr int TN3 = TILE3 + 2;            // nodes per tile edge
constexpr int TNODES3 = TN3 * TN3 * TN3;  // 216
constexpr int P2G_THREADS = 288;          // 9 warps: 64 cells x 9 stencil columns = 2 x 288 work items per full tile
constexpr int G2P_THREADS = 128;

// Per-particle P2G payload parked in shared memory between the two phases.
template <typename T>
struct alignas(16) P2GPayload {
  T mvx, mvy, mvz, m;
  T a00, a01, a02, fx;   // a = affine * dx
  T a10, a11, a12, fy;
  T a20, a21, a22, fz;
  T wz0, wz1, wz2, pad;
};

template <typename T> struct P2GChunk { static constexpr int value = 512; };
template <> struct P2GChunk<double> { static constexpr int value = 256; };

// P2G over binned particles.  Persistent CTAs pull active tiles (4x4x4 base cells)
// from a work counter and walk the tile's contiguous, cell-sorted particle run in
// chunks:
//   phase 1 (thread per particle)  gather the state through `perm`, evaluate the
//           polar decomposition / fixed-corotated stress in fp64
//           (three_d/p2g.py:57-65) and park {m v, m, affine*dx, fx, wz} in smem;
//   phase 2 (thread per (cell, stencil column (i,j)))  walk the cell's particles,
//           accumulating the column's three nodes x {momentum, mass} in registers
//           (three_d/p2g.py:67-80) -- same-node contributions of all particles of a
//           cell are summed before they leave the SM -- then ONE vector reduction
//           (red.global.add.v4.f32) per node.
template <typename T>
__global__ void __launch_bounds__(P2G_THREADS) p2g_tiled3_kernel(DevCfg cfg, StateView<T> s, BinBuffers B,
                                                                 T* __restrict__ grid, ErrRec* err) {
  constexpr int CHUNK = P2GChunk<T>::value;
  __shared__ P2GPayload<T> pay[CHUNK];
  __shared__ int soff[TILE_CELLS + 1];
  __shared__ int s_work;

  const int n_active = B.counters[0];
  const T dx = (T)cfg.dx;
  const long long ny = cfg.n[1], nz = cfg.n[2];

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_work = atomicAdd(&B.counters[1], 1);
    __syncthreads();
    const int wi = s_work;
    if (wi >= n_active) break;
    const int t = B.active_tiles[wi];
    if (threadIdx.x <= TILE_CELLS) soff[threadIdx.x] = B.cell_off[t * TILE_CELLS + threadIdx.x];
    const int tz = t % B.tiles[2], ty = (t / B.tiles[2]) % B.tiles[1], tx = t / (B.tiles[2] * B.tiles[1]);
    const int ox = tx * TILE3, oy = ty * TILE3, oz = tz * TILE3;
    __syncthreads();
    const int start = soff[0], end = soff[TILE_CELLS];

    for (int chunk = start; chunk < end; chunk += CHUNK) {
      const int cend = min(chunk + CHUNK, end);
      // ---- phase 1: one thread per particle ----
      for (int slot = chunk + threadIdx.x; slot < cend; slot += P2G_THREADS) {
        const long long p = B.perm[slot];
        P2GParticle3<T> q = p2g_prepare3(cfg, s, p);
        P2GPayload<T> pl;
        pl.mvx = q.mvx; pl.mvy = q.mvy; pl.mvz = q.mvz; pl.m = q.m;
        pl.a00 = q.a00 * dx; pl.a01 = q.a01 * dx; pl.a02 = q.a02 * dx; pl.fx = q.fx;
        pl.a10 = q.a10 * dx; pl.a11 = q.a11 * dx; pl.a12 = q.a12 * dx; pl.fy = q.fy;
        pl.a20 = q.a20 * dx; pl.a21 = q.a21 * dx; pl.a22 = q.a22 * dx; pl.fz = q.fz;
        bspline(q.fz, pl.wz0, pl.wz1, pl.wz2);
        pl.pad = (T)0;
        pay[slot - chunk] = pl;
      }
      __syncthreads();
      // ---- phase 2: one thread per (cell, column) ----
      // cells intersecting this chunk: binary search of the first/last cell in soff
      int c_lo = 0, c_hi = TILE_CELLS - 1;
      {
        int lo = 0, hi = TILE_CELLS;      // last c with soff[c] <= chunk
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (soff[mid] <= chunk) lo = mid; else hi = mid; }
        c_lo = lo;
        lo = 0; hi = TILE_CELLS;          // last c with soff[c] <= cend - 1
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (soff[mid] <= cend - 1) lo = mid; else hi = mid; }
        c_hi = lo;
      }
      const int n_items = (c_hi - c_lo + 1) * 9;
      for (int item = threadIdx.x; item < n_items; item += P2G_THREADS) {
        const int c = c_lo + item / 9, col = item % 9;
        const int r0 = max(soff[c], chunk) - chunk, r1 = min(soff[c + 1], cend) - chunk;
        if (r0 >= r1) continue;
        const int li = col / 3, lj = col % 3;
        const T ci = (T)li, cj = (T)lj;
        // B-spline piece of this column per axis: w = s * (f - c)^2 + o  (three_d/p2g.py:55)
        const T sx = li == 1 ? (T)-1 : (T)0.5, cx_ = (T)1.5 - (T)0.5 * ci, ox_ = li == 1 ? (T)0.75 : (T)0;
        const T sy = lj == 1 ? (T)-1 : (T)0.5, cy_ = (T)1.5 - (T)0.5 * cj, oy_ = lj == 1 ? (T)0.75 : (T)0;
        T x0 = 0, y0 = 0, z0 = 0, m0 = 0, x1 = 0, y1 = 0, z1 = 0, m1 = 0, x2 = 0, y2 = 0, z2 = 0, m2 = 0;
#pragma unroll 2
        for (int qi = r0; qi < r1; ++qi) {
          const P2GPayload<T> pl = pay[qi];
          const T tx_ = pl.fx - cx_, ty_ = pl.fy - cy_;
          const T wij = (sx * tx_ * tx_ + ox_) * (sy * ty_ * ty_ + oy_);
          const T dpx = ci - pl.fx, dpy = cj - pl.fy;
          const T bx = pl.mvx + (pl.a00 * dpx + pl.a01 * dpy);
          const T by = pl.mvy + (pl.a10 * dpx + pl.a11 * dpy);
          const T bz = pl.mvz + (pl.a20 * dpx + pl.a21 * dpy);
          const T d0 = -pl.fz, d1 = (T)1 - pl.fz, d2 = (T)2 - pl.fz;
          const T w0 = wij * pl.wz0, w1 = wij * pl.wz1, w2 = wij * pl.wz2;
          x0 += w0 * (bx + pl.a02 * d0); y0 += w0 * (by + pl.a12 * d0); z0 += w0 * (bz + pl.a22 * d0); m0 += w0 * pl.m;
          x1 += w1 * (bx + pl.a02 * d1); y1 += w1 * (by + pl.a12 * d1); z1 += w1 * (bz + pl.a22 * d1); m1 += w1 * pl.m;
          x2 += w2 * (bx + pl.a02 * d2); y2 += w2 * (by + pl.a12 * d2); z2 += w2 * (bz + pl.a22 * d2); m2 += w2 * pl.m;
        }
        const int gx = ox + (c >> 4) + li, gy = oy + ((c >> 2) & 3) + lj, gz = oz + (c & 3);
        T* g = grid + (((long long)gx * ny + gy) * nz + gz) * 4;
        red_add4(g, x0, y0, z0, m0);
        red_add4(g + 4, x1, y1, z1, m1);
        red_add4(g + 8, x2, y2, z2, m2);
      }
      __syncthreads();   // phase 2 readers are done before the next chunk overwrites `pay`
    }
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
      T vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22;
      g2p_accumulate3<T>([&](int i, int j, int k) { return tile[cb + (i * TN3 + j) * TN3 + k]; }, fx, fy, fz,
                         vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22);
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
              int blocks_per_sm, cudaStream_t st) {
  (void)n;
  cudaMemsetAsync(&B.counters[1], 0, sizeof(int32_t), st);
  int blocks = min(B.n_tiles, sm_count * blocks_per_sm);
  p2g_tiled3_kernel<T><<<blocks, P2G_THREADS, 0, st>>>(cfg, s, B, grid, err);
  return 1;
}

template <typename T>
int g2p_tiled(const DevCfg& cfg, const StateView<T>& src, const StateView<T>& dst, long long n, BinBuffers& B,
              const T* grid, ErrRec* err, int sm_count, int blocks_per_sm, cudaStream_t st) {
  (void)n;
  cudaMemsetAsync(&B.counters[2], 0, sizeof(int32_t), st);
  int blocks = min(B.n_tiles + 1, sm_count * blocks_per_sm);
  g2p_tiled3_kernel<T><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);
  return 1;
}

}  // namespace ffmpm

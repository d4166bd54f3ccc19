// Tiled 3D G2P over binned particles (see mpm_bin.cuh for the key layout); the binned
// P2G lives in mpm_p2g_runs.cuh.
//
// G2P: persistent CTAs pull active tiles (4x4x4 base cells) from a work counter, stage
// the tile's 6x6x6 velocity block in shared memory, each thread gathers its particle's
// 3x3x3 stencil from there, and the updated state is written to the OTHER particle
// buffer at the binned slot -- so the state is physically in cell order for the next
// substep and every store is fully coalesced.  The kernel also emits the NEXT substep's
// cell key and within-cell rank of every particle (it is the one place that knows the
// advected position), which removes a whole pass over the positions from ffmpm_bin.
#pragma once
#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"
#include "mpm_p2g_runs.cuh"

namespace ffmpm {

constexpr int TN3 = TILE3 + 2;            // nodes per tile edge
constexpr int TNODES3 = TN3 * TN3 * TN3;  // 216
constexpr int G2P_THREADS = 128;

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
        // still outside (the position did not move): next substep's key is the trailing bin again
        B.keys[slot] = B.n_cells;
        B.rank[slot] = atomicAdd(&B.cell_count[B.n_cells], 1);
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
    for (int sbase = start; sbase < end; sbase += blockDim.x) {
      const int slot = sbase + threadIdx.x;
      int next_key = -1;
      if (slot < end) {
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
      const T nx0 = x0 + dt * vx, nx1 = x1 + dt * vy, nx2 = x2 + dt * vz;
      dst.x[slot] = nx0; dst.x[ds + slot] = nx1; dst.x[2 * ds + slot] = nx2;
      carry_planes(src, dst, p, slot, true);
      next_key = bin_key_of<T>(cfg, B, nx0, nx1, nx2);
      B.keys[slot] = next_key;
      }
      // next substep's histogram + within-cell rank (same scheme as bin_count_kernel)
      bin_rank_warp(B, next_key, slot);
    }
  }
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

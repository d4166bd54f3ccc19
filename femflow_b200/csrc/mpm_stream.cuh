// G2P in physical particle order ("stream" pipeline).
//
// The reordering G2P keeps the particle buffers sorted by the cell each particle had
// one substep ago, so a warp's 32 consecutive particles sit in a handful of
// neighbouring cells.  That is all the locality the gather needs: state planes are read
// with fully coalesced loads (no permutation in front of them), the 27 stencil nodes
// come through L1 (lanes of the same cell hit the same 16-byte node -> broadcast), and
// nothing in the kernel synchronises wider than a warp.
//
// Each thread then writes its particle to the OTHER buffer at
//     slot = cell_off[key] + rank
// where (key, rank) is the counting-sort position of the particle's CURRENT cell --
// emitted by the previous substep's G2P, scanned by ffmpm_bin -- and emits
// (key', rank', histogram) of the ADVECTED position for the next substep.  So the sort
// is applied by the one kernel that rewrites the state anyway: no separate permutation
// pass and no gather through an index array.
#pragma once
#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"

namespace ffmpm {

constexpr int G2P_STREAM_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(G2P_STREAM_THREADS)
g2p_stream3_kernel(DevCfg cfg, StateView<T> src, StateView<T> dst, long long n, BinBuffers B,
                   const int32_t* __restrict__ keys_cur, const int32_t* __restrict__ rank_cur,
                   int32_t* __restrict__ keys_next, int32_t* __restrict__ rank_next, const T* __restrict__ grid,
                   ErrRec* err) {
  const long long ss = src.stride, ds = dst.stride;
  const long long ny = cfg.n[1], nz = cfg.n[2];
  // block-uniform trip count: bin_rank_warp needs all 32 lanes
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int next_key = -1;
  long long slot = 0;
  if (p < n) {
    const int key = keys_cur[p];
    slot = (long long)B.cell_off[key] + rank_cur[p];
    const T x0 = src.x[p], x1 = src.x[ss + p], x2 = src.x[2 * ss + p];
    if (key >= B.n_cells) {
      // outside the grid (RuntimeError in the reference, three_d/g2p.py:23-24): carried over unchanged
      atomicAdd(&err->n_oob, 1ULL);
      for (int c = 0; c < 3; ++c) { dst.x[c * ds + slot] = src.x[c * ss + p]; dst.v[c * ds + slot] = src.v[c * ss + p]; }
      for (int c = 0; c < 9; ++c) { dst.C[c * ds + slot] = src.C[c * ss + p]; dst.F[c * ds + slot] = src.F[c * ss + p]; }
      next_key = B.n_cells;
    } else {
      int gx, gy, gz;
      T fx, fy, fz;
      base_fx(x0, cfg.inv_dx, gx, fx);
      base_fx(x1, cfg.inv_dx, gy, fy);
      base_fx(x2, cfg.inv_dx, gz, fz);
      const T* gb = grid + (((long long)(gx - cfg.origin[0]) * ny + (gy - cfg.origin[1])) * nz + (gz - cfg.origin[2])) * 4;
      T vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22;
      g2p_accumulate3<T>([&](int i, int j, int k) { return ld_node(gb + (((long long)i * ny + j) * nz + k) * 4); }, fx,
                         fy, fz, vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22);
      const T s4 = (T)(4.0 * cfg.inv_dx);
      c00 *= s4; c01 *= s4; c02 *= s4; c10 *= s4; c11 *= s4; c12 *= s4; c20 *= s4; c21 *= s4; c22 *= s4;
      const T dt = (T)cfg.dt;
      const T f00 = src.F[0 * ss + p], f01 = src.F[1 * ss + p], f02 = src.F[2 * ss + p];
      const T f10 = src.F[3 * ss + p], f11 = src.F[4 * ss + p], f12 = src.F[5 * ss + p];
      const T f20 = src.F[6 * ss + p], f21 = src.F[7 * ss + p], f22 = src.F[8 * ss + p];
      // F <- (I + dt C) F   (three_d/g2p.py:46)
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
      const T nx0 = x0 + dt * vx, nx1 = x1 + dt * vy, nx2 = x2 + dt * vz;   // three_d/g2p.py:45
      dst.x[slot] = nx0; dst.x[ds + slot] = nx1; dst.x[2 * ds + slot] = nx2;
      next_key = bin_key_of<T>(cfg, B, nx0, nx1, nx2);
    }
    if (src.mass) dst.mass[slot] = src.mass[p];
    if (src.mu0) dst.mu0[slot] = src.mu0[p];
    if (src.lam0) dst.lam0[slot] = src.lam0[p];
    if (src.id) dst.id[slot] = src.id[p];
    if (src.Jp) dst.Jp[slot] = src.Jp[p];
    keys_next[slot] = next_key;
  }
  // next substep's histogram + within-cell rank
  {
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, next_key);
    if (next_key >= 0) {
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if ((int)lane == leader) base = atomicAdd(&B.cell_count[next_key], __popc(peers));
      base = __shfl_sync(peers, base, leader);
      rank_next[slot] = base + __popc(peers & ((1u << lane) - 1u));
    }
  }
}

template <typename T>
int g2p_stream(const DevCfg& cfg, const StateView<T>& src, const StateView<T>& dst, long long n, BinBuffers& B,
               const int32_t* keys_cur, const int32_t* rank_cur, int32_t* keys_next, int32_t* rank_next, const T* grid,
               ErrRec* err, cudaStream_t st) {
  unsigned blocks = (unsigned)((n + G2P_STREAM_THREADS - 1) / G2P_STREAM_THREADS);
  g2p_stream3_kernel<T><<<blocks, G2P_STREAM_THREADS, 0, st>>>(cfg, src, dst, n, B, keys_cur, rank_cur, keys_next,
                                                               rank_next, grid, err);
  return 1;
}

}  // namespace ffmpm

// Fused G2P(k) + P2G(k+1) over binned particles ("G2P2G").
//
// P2G of substep k+1 consumes exactly what G2P of substep k produces (x, v, C, F of
// every particle), so the two are done by ONE kernel while the particle is still in
// registers: the 27 state planes are read once per substep instead of twice (-108 B per
// particle), the latency-bound gather half and the issue-bound scatter half share the
// SM, and a whole kernel boundary disappears.
//
//   * persistent CTAs pull active tiles (4x4x4 base cells), stage the tile's 6x6x6
//     velocity block of grid A (velocities of substep k) in shared memory;
//   * every warp owns 64 consecutive binned slots per round (two per lane): gather,
//     advect, F <- (I + dt C) F (three_d/g2p.py:21-59), write the particle to the other
//     state buffer in cell order, emit its next key / rank / histogram entry;
//   * still in registers: fixed-corotated stress of the NEW state (three_d/p2g.py:57-65),
//     payload parked in the warp's smem slab;
//   * runs of equal (new) base cell -> lane per (run, x-slab) register accumulation ->
//     one red.global.add.v4.f32 per node into grid B (mass/momentum of substep k+1,
//     three_d/p2g.py:67-80).  Particles are sorted by their OLD cell; one substep moves
//     them by less than a cell, so the runs stay long.
#pragma once
#include "mpm_tiled.cuh"

namespace ffmpm {

constexpr int GG_WARPS = 4;
constexpr int GG_THREADS = GG_WARPS * 32;
constexpr int GG_ROUND = GG_WARPS * P2G_WINDOW;   // 256 slots per CTA round

template <typename T, int MIN_BLOCKS>
__global__ void __launch_bounds__(GG_THREADS, MIN_BLOCKS)
g2p2g_tiled3_kernel(DevCfg cfg, StateView<T> src, StateView<T> dst, BinBuffers B, const T* __restrict__ grid_in,
                    T* __restrict__ grid_out, ErrRec* err) {
  using V4 = typename Vec4<T>::type;
  __shared__ V4 tile[TNODES3];
  __shared__ int s_work;
  __shared__ P2GWarpSlab<T> slabs[GG_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  P2GWarpSlab<T>& S = slabs[warp];
  const int n_active = B.counters[0];
  const long long ss = src.stride, ds = dst.stride;
  const int ny = cfg.n[1], nz = cfg.n[2];
  const T dxs = (T)cfg.dx;
  const int mat_mode = mat_mode_of(src);
  const bool has_mat = mat_mode != MAT_CFG;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_work = atomicAdd(&B.counters[2], 1);
    __syncthreads();
    const int wi = s_work;
    if (wi > n_active) break;
    if (wi == n_active) {
      // trailing bin: particles outside the grid are carried over unchanged and scatter nothing
      const int start = B.cell_off[B.n_cells], end = B.cell_off[B.n_cells + 1];
      for (int slot = start + threadIdx.x; slot < end; slot += blockDim.x) {
        const long long p = B.perm[slot];
        for (int c = 0; c < 3; ++c) { dst.x[c * ds + slot] = src.x[c * ss + p]; dst.v[c * ds + slot] = src.v[c * ss + p]; }
        for (int c = 0; c < 9; ++c) { dst.C[c * ds + slot] = src.C[c * ss + p]; dst.F[c * ds + slot] = src.F[c * ss + p]; }
        carry_planes(src, dst, p, slot, true);
        B.keys[slot] = B.n_cells;
        B.rank[slot] = atomicAdd(&B.cell_count[B.n_cells], 1);
      }
      continue;
    }
    const int t = B.active_tiles[wi];
    const int start = B.cell_off[t * TILE_CELLS], end = B.cell_off[(t + 1) * TILE_CELLS];
    const int tz = t % B.tiles[2], ty = (t / B.tiles[2]) % B.tiles[1], tx = t / (B.tiles[2] * B.tiles[1]);
    const int ox = tx * TILE3, oy = ty * TILE3, oz = tz * TILE3;
    for (int nd = threadIdx.x; nd < TNODES3; nd += blockDim.x) {
      const int k = nd % TN3, j = (nd / TN3) % TN3, i = nd / (TN3 * TN3);
      const int gx = ox + i, gy = oy + j, gz = oz + k;
      V4 g;
      g.x = g.y = g.z = g.w = (T)0;
      if (gx < cfg.n[0] && gy < cfg.n[1] && gz < cfg.n[2])
        g = ld_node(grid_in + (((long long)gx * cfg.n[1] + gy) * cfg.n[2] + gz) * 4);
      tile[nd] = g;
    }
    __syncthreads();
    for (int rbase = (start / GG_ROUND) * GG_ROUND; rbase < end; rbase += GG_ROUND) {
      int node[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int idx = h * 32 + lane;
        const int slot = rbase + warp * P2G_WINDOW + idx;
        const bool mine = slot >= start && slot < end;
        int next_key = -1;
        node[h] = -1;
        if (mine) {
          const long long p = B.perm[slot];
          T cm = 0, cmu = 0, cl = 0, cjp = 1;
          int cid = 0;
          if (src.mass) cm = src.mass[p];
          if (src.mu0) cmu = src.mu0[p];
          if (src.lam0) cl = src.lam0[p];
          if (src.id) cid = src.id[p];
          if (src.Jp) cjp = src.Jp[p];
          unsigned char row = 0;
          if (src.material) row = src.material[p];
          if (mat_mode == MAT_TABLE) {
            cm = src.mat_table[row]; cmu = src.mat_table[MAT_ROWS + row]; cl = src.mat_table[2 * MAT_ROWS + row];
          }
          const T x0 = src.x[p], x1 = src.x[ss + p], x2 = src.x[2 * ss + p];
          const T f00 = src.F[0 * ss + p], f01 = src.F[1 * ss + p], f02 = src.F[2 * ss + p];
          const T f10 = src.F[3 * ss + p], f11 = src.F[4 * ss + p], f12 = src.F[5 * ss + p];
          const T f20 = src.F[6 * ss + p], f21 = src.F[7 * ss + p], f22 = src.F[8 * ss + p];
          int gx, gy, gz;
          T fx, fy, fz;
          base_fx(x0, cfg, gx, fx);
          base_fx(x1, cfg, gy, fy);
          base_fx(x2, cfg, gz, fz);
          const int cb = ((gx - cfg.origin[0] - ox) * TN3 + (gy - cfg.origin[1] - oy)) * TN3 + (gz - cfg.origin[2] - oz);
          T o[24];
          {
            T vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22;
            g2p_accumulate3<T>([&](int i, int j, int k) { return tile[cb + (i * TN3 + j) * TN3 + k]; }, fx, fy, fz,
                               vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22);
            const T s4 = (T)(4.0 * cfg.inv_dx);
            c00 *= s4; c01 *= s4; c02 *= s4; c10 *= s4; c11 *= s4; c12 *= s4; c20 *= s4; c21 *= s4; c22 *= s4;
            const T dt = (T)cfg.dt;
            const T m00 = (T)1 + dt * c00, m01 = dt * c01, m02 = dt * c02;
            const T m10 = dt * c10, m11 = (T)1 + dt * c11, m12 = dt * c12;
            const T m20 = dt * c20, m21 = dt * c21, m22 = (T)1 + dt * c22;
            o[0] = x0 + dt * vx; o[1] = x1 + dt * vy; o[2] = x2 + dt * vz;
            o[3] = vx; o[4] = vy; o[5] = vz;
            o[6] = c00; o[7] = c01; o[8] = c02; o[9] = c10; o[10] = c11; o[11] = c12; o[12] = c20; o[13] = c21; o[14] = c22;
            o[15] = m00 * f00 + m01 * f10 + m02 * f20;
            o[16] = m00 * f01 + m01 * f11 + m02 * f21;
            o[17] = m00 * f02 + m01 * f12 + m02 * f22;
            o[18] = m10 * f00 + m11 * f10 + m12 * f20;
            o[19] = m10 * f01 + m11 * f11 + m12 * f21;
            o[20] = m10 * f02 + m11 * f12 + m12 * f22;
            o[21] = m20 * f00 + m21 * f10 + m22 * f20;
            o[22] = m20 * f01 + m21 * f11 + m22 * f21;
            o[23] = m20 * f02 + m21 * f12 + m22 * f22;
          }
          // ---- G2P output: the particle in cell order in the other buffer ----
#pragma unroll
          for (int k = 0; k < 3; ++k) { dst.x[k * ds + slot] = o[k]; dst.v[k * ds + slot] = o[3 + k]; }
#pragma unroll
          for (int k = 0; k < 9; ++k) { dst.C[k * ds + slot] = o[6 + k]; dst.F[k * ds + slot] = o[15 + k]; }
          if (src.mass) dst.mass[slot] = cm;
          if (src.mu0) dst.mu0[slot] = cmu;
          if (src.lam0) dst.lam0[slot] = cl;
          if (src.id) dst.id[slot] = cid;
          if (src.Jp) dst.Jp[slot] = cjp;
          if (src.material) dst.material[slot] = row;
          next_key = bin_key_of<T>(cfg, B, o[0], o[1], o[2]);
          B.keys[slot] = next_key;
          // ---- P2G of the next substep, phase 1, from registers ----
          P2GParticle3<T> q = p2g_prepare3_from<T>(
              cfg,
              [&](int k) -> T {
                if (k < P2G_MASS) return o[k];     // x v C F are laid out in P2G plane order
                if (k == P2G_MASS) return cm;
                if (k == P2G_MU) return cmu;
                return cl;
              },
              has_mat, (double)cjp);
          node[h] = p2g_park(S, q, idx, dxs, ny, nz);
        } else {
          S.node0[idx] = -1;   // padding slot of a partial round: an empty run
        }
        bin_rank_warp(B, next_key, slot);
      }
      __syncwarp();
      p2g_runs_phase2<T>(S, node, P2G_WINDOW, lane, ny, nz, grid_out);
      __syncwarp();
    }
  }
}

template <typename T>
int g2p2g_tiled(const DevCfg& cfg, const StateView<T>& src, const StateView<T>& dst, BinBuffers& B, const T* grid_in,
                T* grid_out, ErrRec* err, int sm_count, int blocks_per_sm, cudaStream_t st) {
  cudaMemsetAsync(&B.counters[2], 0, sizeof(int32_t), st);
  int blocks = min(B.n_tiles + 1, sm_count * blocks_per_sm);
  static int minb = [] { const char* e = getenv("FFMPM_GG_MINB"); return e ? atoi(e) : 4; }();
  if (minb >= 5)
    g2p2g_tiled3_kernel<T, 5><<<blocks, GG_THREADS, 0, st>>>(cfg, src, dst, B, grid_in, grid_out, err);
  else if (minb == 3)
    g2p2g_tiled3_kernel<T, 3><<<blocks, GG_THREADS, 0, st>>>(cfg, src, dst, B, grid_in, grid_out, err);
  else
    g2p2g_tiled3_kernel<T, 4><<<blocks, GG_THREADS, 0, st>>>(cfg, src, dst, B, grid_in, grid_out, err);
  return 1;
}

}  // namespace ffmpm

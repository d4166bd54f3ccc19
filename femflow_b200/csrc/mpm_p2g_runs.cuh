// P2G over binned (cell-sorted) particles, warp-autonomous form.
//
// Every warp owns windows of P2G_WINDOW consecutive binned slots and never talks to
// another warp (no block barrier, no shared grid tile, no work counter):
//   phase 1 (lane per particle)   gather the state through `perm`, evaluate the polar
//           decomposition / fixed-corotated stress in fp64 (three_d/p2g.py:57-65) and
//           park {m v, m, affine*dx, fx, wz} in the warp's private smem slab;
//   runs    a ballot over "my cell differs from my predecessor's" splits the window
//           into runs of particles that share a base cell;
//   phase 2 (lane per (run, stencil column (i,j)))   walk the run, accumulating the
//           column's three nodes x {momentum, mass} in registers
//           (three_d/p2g.py:67-80), then ONE red.global.add.v4.f32 per node.
// Same-node contributions of all particles of a cell are therefore summed on chip and
// leave the SM as 27 vector reductions per run instead of 27 per particle.
#pragma once
#include <cstdlib>

#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"

namespace ffmpm {

constexpr int P2G_WINDOW = 64;       // slots per warp window (2 per lane)
constexpr int P2G_RUN_WARPS = 4;     // warps per CTA

template <typename T>
struct alignas(16) P2GRunPayload {
  T mvx, mvy, mvz, m;
  T a00, a01, a02, fx;   // a = affine * dx
  T a10, a11, a12, fy;
  T a20, a21, a22, fz;
  T wz0, wz1, wz2, pad;
};

template <typename T>
struct P2GWarpSlab {
  P2GRunPayload<T> pay[P2G_WINDOW];
  int node0[P2G_WINDOW];          // linear LOCAL node id of the particle's base cell
  int run_start[P2G_WINDOW + 1];  // window-relative first slot of each run (+ sentinel)
};

template <typename T, int MIN_BLOCKS>
__global__ void __launch_bounds__(P2G_RUN_WARPS * 32, MIN_BLOCKS)
p2g_runs3_kernel(DevCfg cfg, StateView<T> s, BinBuffers B, T* __restrict__ grid, ErrRec* err) {
  __shared__ P2GWarpSlab<T> slabs[P2G_RUN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  P2GWarpSlab<T>& S = slabs[warp];
  const T dx = (T)cfg.dx;
  const int ny = cfg.n[1], nz = cfg.n[2];
  // slots [0, n_in) hold the particles that are inside the grid (the trailing bin is skipped)
  const int n_in = B.cell_off[B.n_cells];
  const int n_windows = (n_in + P2G_WINDOW - 1) / P2G_WINDOW;
  const int total_warps = gridDim.x * P2G_RUN_WARPS;

  for (int win = blockIdx.x * P2G_RUN_WARPS + warp; win < n_windows; win += total_warps) {
    const int w0 = win * P2G_WINDOW;
    const int cnt = min(P2G_WINDOW, n_in - w0);
    // ---- phase 1 ----
    int node[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = h * 32 + lane;
      node[h] = -1;
      if (idx < cnt) {
        const long long p = B.perm[w0 + idx];
        P2GParticle3<T> q = p2g_prepare3(cfg, s, p);
        P2GRunPayload<T> pl;
        pl.mvx = q.mvx; pl.mvy = q.mvy; pl.mvz = q.mvz; pl.m = q.m;
        pl.a00 = q.a00 * dx; pl.a01 = q.a01 * dx; pl.a02 = q.a02 * dx; pl.fx = q.fx;
        pl.a10 = q.a10 * dx; pl.a11 = q.a11 * dx; pl.a12 = q.a12 * dx; pl.fy = q.fy;
        pl.a20 = q.a20 * dx; pl.a21 = q.a21 * dx; pl.a22 = q.a22 * dx; pl.fz = q.fz;
        bspline(q.fz, pl.wz0, pl.wz1, pl.wz2);
        pl.pad = (T)0;
        S.pay[idx] = pl;
        node[h] = (q.bx * ny + q.by) * nz + q.bz;
        S.node0[idx] = node[h];
      }
    }
    __syncwarp();
    // run heads: first slot of the window, or base cell differs from the predecessor's
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
    // ---- phase 2 ----
    const int n_items = n_runs * 9;
    for (int item = lane; item < n_items; item += 32) {
      const int r = item / 9, col = item - r * 9;
      const int r0 = S.run_start[r], r1 = S.run_start[r + 1];
      const int li = col / 3, lj = col - li * 3;
      const T ci = (T)li, cj = (T)lj;
      // B-spline piece of this column per axis: w = s * (f - c)^2 + o  (three_d/p2g.py:55)
      const T sx = li == 1 ? (T)-1 : (T)0.5, cx_ = (T)1.5 - (T)0.5 * ci, ox_ = li == 1 ? (T)0.75 : (T)0;
      const T sy = lj == 1 ? (T)-1 : (T)0.5, cy_ = (T)1.5 - (T)0.5 * cj, oy_ = lj == 1 ? (T)0.75 : (T)0;
      T x0 = 0, y0 = 0, z0 = 0, m0 = 0, x1 = 0, y1 = 0, z1 = 0, m1 = 0, x2 = 0, y2 = 0, z2 = 0, m2 = 0;
      for (int qi = r0; qi < r1; ++qi) {
        const P2GRunPayload<T> pl = S.pay[qi];
        const T tx_ = pl.fx - cx_, ty_ = pl.fy - cy_;
        const T wij = (sx * tx_ * tx_ + ox_) * (sy * ty_ * ty_ + oy_);
        const T dpx = ci - pl.fx, dpy = cj - pl.fy;
        const T bx = pl.mvx + (pl.a00 * dpx + pl.a01 * dpy);
        const T by = pl.mvy + (pl.a10 * dpx + pl.a11 * dpy);
        const T bz = pl.mvz + (pl.a20 * dpx + pl.a21 * dpy);
        const T d0 = -pl.fz, d1 = (T)1 - pl.fz, d2 = (T)2 - pl.fz;
        const T w0_ = wij * pl.wz0, w1_ = wij * pl.wz1, w2_ = wij * pl.wz2;
        x0 += w0_ * (bx + pl.a02 * d0); y0 += w0_ * (by + pl.a12 * d0); z0 += w0_ * (bz + pl.a22 * d0); m0 += w0_ * pl.m;
        x1 += w1_ * (bx + pl.a02 * d1); y1 += w1_ * (by + pl.a12 * d1); z1 += w1_ * (bz + pl.a22 * d1); m1 += w1_ * pl.m;
        x2 += w2_ * (bx + pl.a02 * d2); y2 += w2_ * (by + pl.a12 * d2); z2 += w2_ * (bz + pl.a22 * d2); m2 += w2_ * pl.m;
      }
      T* g = grid + ((long long)S.node0[r0] + (long long)(li * ny + lj) * nz) * 4;
      red_add4(g, x0, y0, z0, m0);
      red_add4(g + 4, x1, y1, z1, m1);
      red_add4(g + 8, x2, y2, z2, m2);
    }
    __syncwarp();   // the slab is rewritten by the next window
  }
}

template <typename T>
int p2g_runs(const DevCfg& cfg, const StateView<T>& s, long long n, BinBuffers& B, T* grid, ErrRec* err, int sm_count,
             int blocks_per_sm, cudaStream_t st) {
  long long windows = (n + P2G_WINDOW - 1) / P2G_WINDOW;
  long long want = (windows + P2G_RUN_WARPS - 1) / P2G_RUN_WARPS;
  long long cap = (long long)sm_count * blocks_per_sm;
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  // MIN_BLOCKS trades registers (fp64 polar iteration) for resident warps; tunable via FFMPM_P2G_MINB
  static int minb = [] { const char* e = getenv("FFMPM_P2G_MINB"); return e ? atoi(e) : 4; }();
  if (minb >= 8)
    p2g_runs3_kernel<T, 8><<<blocks, P2G_RUN_WARPS * 32, 0, st>>>(cfg, s, B, grid, err);
  else if (minb >= 6)
    p2g_runs3_kernel<T, 6><<<blocks, P2G_RUN_WARPS * 32, 0, st>>>(cfg, s, B, grid, err);
  else if (minb == 5)
    p2g_runs3_kernel<T, 5><<<blocks, P2G_RUN_WARPS * 32, 0, st>>>(cfg, s, B, grid, err);
  else
    p2g_runs3_kernel<T, 4><<<blocks, P2G_RUN_WARPS * 32, 0, st>>>(cfg, s, B, grid, err);
  return 1;
}

}  // namespace ffmpm

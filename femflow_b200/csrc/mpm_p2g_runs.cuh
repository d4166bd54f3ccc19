// P2G over cell-sorted particles, warp-autonomous form.
//
// Every warp owns windows of P2G_WINDOW consecutive slots and never talks to another
// warp (no block barrier, no shared grid tile, no work counter):
//   phase 1 (lane per particle)   fetch the state (through `perm`, or in physical
//           order), evaluate the fixed-corotated stress (three_d/p2g.py:57-65) and park
//           {m v, m, affine*dx, fx} in the warp's private smem slab;
//   runs    a ballot over "my base cell differs from my predecessor's" splits the
//           window into runs of particles that share a base cell;
//   phase 2 (lane per (run, x-slab i))   walk the run, accumulating the slab's nine
//           nodes x {momentum, mass} in registers (three_d/p2g.py:67-80), then ONE
//           red.global.add.v4.f32 per node.
// Same-node contributions of all particles of a cell are therefore summed on chip and
// leave the SM as 27 vector reductions per run instead of 27 per particle.  The scheme
// is correct for ANY particle order (an unsorted particle merely splits a run); sorted
// order is what makes the runs long.
//
// Shared-memory layout: four float4 planes indexed by a padded slot q + q/8.  With
// ~8 particles per cell the lanes of one LDS.128 read particles ~8 slots apart;
// without the padding those 16-byte chunks fall on the same banks (measured: 5.3
// wavefronts per LDS.128 against 3.0 ideal, profiles/r01c).
#pragma once
#include <cstdlib>

#include "mpm_bin.cuh"
#include "mpm_common.cuh"
#include "mpm_direct.cuh"

namespace ffmpm {

constexpr int P2G_WINDOW = 64;                           // slots per warp window (2 per lane)
constexpr int P2G_PADDED = P2G_WINDOW + P2G_WINDOW / 8;  // padded slab length
constexpr int P2G_RUN_WARPS = 4;                         // warps per CTA

__device__ __forceinline__ int p2g_pad(int q) { return q + (q >> 3); }

template <typename T>
struct alignas(16) P2GVec4 {
  T x, y, z, w;
};

template <typename T>
struct P2GWarpSlab {
  P2GVec4<T> pay[4][P2G_PADDED];  // {mvx,mvy,mvz,m} {a00,a01,a02,fx} {a10,a11,a12,fy} {a20,a21,a22,fz}, a = affine*dx
  int node0[P2G_WINDOW];          // linear LOCAL node id of the particle's base cell (-1: outside the grid)
  int run_start[P2G_WINDOW + 1];  // window-relative first slot of each run (+ sentinel)
  unsigned char order[P2G_WINDOW];   // INDIRECT phase 2: position in cell-sorted order -> slot of the window
};

// Phase 1 tail: park one particle's payload.
template <typename T>
__device__ __forceinline__ int p2g_park(P2GWarpSlab<T>& S, P2GParticle3<T>& q, int idx, T dx, int ny, int nz) {
  const int ph = p2g_pad(idx);
  if (!q.ok) {   // outside the grid: contributes nothing (flagged by the binning / G2P)
    q.mvx = q.mvy = q.mvz = q.m = (T)0;
    q.a00 = q.a01 = q.a02 = q.a10 = q.a11 = q.a12 = q.a20 = q.a21 = q.a22 = (T)0;
    q.fx = q.fy = q.fz = (T)0.5;
  }
  S.pay[0][ph] = P2GVec4<T>{q.mvx, q.mvy, q.mvz, q.m};
  S.pay[1][ph] = P2GVec4<T>{q.a00 * dx, q.a01 * dx, q.a02 * dx, q.fx};
  S.pay[2][ph] = P2GVec4<T>{q.a10 * dx, q.a11 * dx, q.a12 * dx, q.fy};
  S.pay[3][ph] = P2GVec4<T>{q.a20 * dx, q.a21 * dx, q.a22 * dx, q.fz};
  const int node = q.ok ? (q.bx * ny + q.by) * nz + q.bz : -1;
  S.node0[idx] = node;
  return node;
}

// Phase 2 proper: lane per (run, x-slab) over the run table S.run_start[0 .. n_runs] of a parked window.
// INDIRECT: the runs are runs of the window's particles in CELL-SORTED order (S.order maps a sorted position to its
// slot) instead of runs of consecutive slots -- see p2g_sort_window.
template <typename T, bool INDIRECT = false>
__device__ __forceinline__ void p2g_accumulate_runs(P2GWarpSlab<T>& S, int n_runs, int lane, int ny, int nz,
                                                    T* __restrict__ grid) {
  const int n_items = n_runs * 3;
  for (int item = lane; item < n_items; item += 32) {
    const int r = item / 3, li = item - r * 3;
    const int r0 = S.run_start[r], r1 = S.run_start[r + 1];
    const int node_r = S.node0[INDIRECT ? (int)S.order[r0] : r0];
    if (node_r < 0) continue;   // a run of out-of-grid particles
    const T ci = (T)li;
    // B-spline piece of this slab along x: w = s * (f - c)^2 + o  (three_d/p2g.py:55)
    const T sx = li == 1 ? (T)-1 : (T)0.5, cx_ = (T)1.5 - (T)0.5 * ci, ox_ = li == 1 ? (T)0.75 : (T)0;
    T ax[9], ay[9], az[9], am[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) ax[e] = ay[e] = az[e] = am[e] = (T)0;
    for (int qi = r0; qi < r1; ++qi) {
      const int ph = p2g_pad(INDIRECT ? (int)S.order[qi] : qi);
      const P2GVec4<T> p0 = S.pay[0][ph], p1 = S.pay[1][ph], p2 = S.pay[2][ph], p3 = S.pay[3][ph];
      const T fx = p1.w, fy = p2.w, fz = p3.w;
      T wy[3], wz[3];
      bspline(fy, wy[0], wy[1], wy[2]);
      bspline(fz, wz[0], wz[1], wz[2]);
      const T tx_ = fx - cx_;
      const T wxi = sx * tx_ * tx_ + ox_;
      const T dpx = ci - fx;
      const T bx = p0.x + p1.x * dpx, by = p0.y + p2.x * dpx, bz = p0.z + p3.x * dpx;
      const T dz[3] = {-fz, (T)1 - fz, (T)2 - fz};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const T dpy = (T)j - fy;
        const T wij = wxi * wy[j];
        const T cxj = bx + p1.y * dpy, cyj = by + p2.y * dpy, czj = bz + p3.y * dpy;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const T w = wij * wz[k];
          ax[j * 3 + k] += w * (cxj + p1.z * dz[k]);
          ay[j * 3 + k] += w * (cyj + p2.z * dz[k]);
          az[j * 3 + k] += w * (czj + p3.z * dz[k]);
          am[j * 3 + k] += w * p0.w;
        }
      }
    }
    T* g = grid + ((long long)node_r + (long long)li * ny * nz) * 4;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        red_add4(g + ((long long)j * nz + k) * 4, ax[j * 3 + k], ay[j * 3 + k], az[j * 3 + k], am[j * 3 + k]);
  }
}

// Cell-sort of one window inside the warp.  The reordering G2P stores a particle at the slot of the cell it was in
// BEFORE it was advected, so with particles moving c cells per substep a fraction ~c of a window sits in a
// neighbouring cell's run: every such particle is a run of its own and cuts another run in two, the (run, slab)
// items overflow the 32 lanes and phase 2 takes two or three passes of short, ragged runs.  Sorting the window's
// 64 (base node, slot) pairs -- a bitonic network over two keys per lane, shuffles only -- merges all particles of
// a cell within the window into ONE run again.  lane l holds the elements 2l and 2l+1; key = node << 6 | slot
// (nodes are compared relative to the window's smallest: 26 bits), out-of-grid particles sort last.
// Returns the number of runs and fills S.order / S.run_start, or -1 when the window's nodes span more than 2^26 ids
// (the caller then keeps the unsorted run table).  ~150 instructions: only called when the unsorted run table would
// need more than one pass.
__device__ __forceinline__ int p2g_sort_window(P2GWarpSlab<float>& S, const int (&node)[2], int cnt, int lane) {
  const unsigned big = 0x03ffffffu;                                    // out of grid / past the end: sorts last
  unsigned lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (2 * lane + h < cnt && node[h] >= 0) { lo = min(lo, (unsigned)node[h]); hi = max(hi, (unsigned)node[h]); }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if (lo != 0xffffffffu && hi - lo >= big) return -1;
  unsigned key[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const bool live = 2 * lane + h < cnt && node[h] >= 0;
    const unsigned rel = live ? (unsigned)node[h] - lo : big;
    key[h] = (rel << 6) | (unsigned)(2 * lane + h);
  }
#pragma unroll
  for (int k = 2; k <= P2G_WINDOW; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j == 1) {
        // partner = the lane's other element; ascending iff bit k of the element index is clear
        const bool up = ((2 * lane) & k) == 0;
        const unsigned a = min(key[0], key[1]), b = max(key[0], key[1]);
        key[0] = up ? a : b;
        key[1] = up ? b : a;
      } else {
        const int lj = j >> 1;                                         // partner lane distance
        const bool lower = (lane & lj) == 0;                           // this lane holds the lower element index
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const unsigned other = __shfl_xor_sync(0xffffffffu, key[h], lj);
          const bool up = ((2 * lane + h) & k) == 0;
          key[h] = (lower == up) ? min(key[h], other) : max(key[h], other);
        }
      }
    }
  }
  // sorted position 2l + h holds key[h]
  S.order[2 * lane] = (unsigned char)(key[0] & 63u);
  S.order[2 * lane + 1] = (unsigned char)(key[1] & 63u);
  const unsigned c0 = key[0] >> 6, c1 = key[1] >> 6;
  unsigned prev = __shfl_up_sync(0xffffffffu, c1, 1);
  if (lane == 0) prev = 0xffffffffu;
  // elements past `cnt` carry the `big` key and sort behind everything: position < cnt <=> a real particle
  const unsigned h0 = __ballot_sync(0xffffffffu, 2 * lane < cnt && c0 != prev);
  const unsigned h1 = __ballot_sync(0xffffffffu, 2 * lane + 1 < cnt && c1 != c0);
  const unsigned below = (1u << lane) - 1u;
  const int r_lo = __popc(h0 & below) + __popc(h1 & below);
  const unsigned mine0 = (h0 >> lane) & 1u, mine1 = (h1 >> lane) & 1u;
  if (mine0) S.run_start[r_lo] = 2 * lane;
  if (mine1) S.run_start[r_lo + mine0] = 2 * lane + 1;
  const int n_runs = __popc(h0) + __popc(h1);
  if (lane == 0) S.run_start[n_runs] = cnt;
  return n_runs;
}

// Runs + phase 2 over a parked window (call after a __syncwarp() that follows phase 1); lane l owns the
// slots l and 32 + l.
template <typename T>
__device__ __forceinline__ void p2g_runs_phase2(P2GWarpSlab<T>& S, const int (&node)[2], int cnt, int lane, int ny,
                                                int nz, T* __restrict__ grid) {
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
  p2g_accumulate_runs<T>(S, n_runs, lane, ny, nz, grid);
}

// USE_PERM: walk the particles through the counting-sort permutation (exactly cell-sorted);
// otherwise in physical order, which the reordering G2P keeps sorted up to one substep of
// motion.
template <typename T, int MIN_BLOCKS, bool USE_PERM>
__global__ void __launch_bounds__(P2G_RUN_WARPS * 32, MIN_BLOCKS)
p2g_runs3_kernel(DevCfg cfg, StateView<T> s, long long n, BinBuffers B, T* __restrict__ grid, ErrRec* err) {
  __shared__ P2GWarpSlab<T> slabs[P2G_RUN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  P2GWarpSlab<T>& S = slabs[warp];
  const T dx = (T)cfg.dx;
  const int ny = cfg.n[1], nz = cfg.n[2];
  // with the permutation, slots [0, n_in) hold the particles inside the grid (the trailing bin is skipped)
  const int n_in = USE_PERM ? B.cell_off[B.n_cells] : (int)n;
  const int n_windows = (n_in + P2G_WINDOW - 1) / P2G_WINDOW;
  const int total_warps = gridDim.x * P2G_RUN_WARPS;

  for (int win = blockIdx.x * P2G_RUN_WARPS + warp; win < n_windows; win += total_warps) {
    const int w0 = win * P2G_WINDOW;
    const int cnt = min(P2G_WINDOW, n_in - w0);
    int node[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = h * 32 + lane;
      node[h] = -1;
      if (idx < cnt) {
        const long long p = USE_PERM ? (long long)B.perm[w0 + idx] : (long long)(w0 + idx);
        P2GParticle3<T> q = p2g_prepare3(cfg, s, p);
        node[h] = p2g_park(S, q, idx, dx, ny, nz);
      }
    }
    __syncwarp();
    p2g_runs_phase2<T>(S, node, cnt, lane, ny, nz, grid);
    __syncwarp();   // the slab is rewritten by the next window
  }
}

template <typename T>
int p2g_runs(const DevCfg& cfg, const StateView<T>& s, long long n, BinBuffers& B, T* grid, ErrRec* err, int sm_count,
             int blocks_per_sm, bool use_perm, cudaStream_t st) {
  long long windows = (n + P2G_WINDOW - 1) / P2G_WINDOW;
  long long want = (windows + P2G_RUN_WARPS - 1) / P2G_RUN_WARPS;
  long long cap = (long long)sm_count * blocks_per_sm;
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  // MIN_BLOCKS trades registers for resident warps; tunable via FFMPM_P2G_MINB
  static int minb = [] { const char* e = getenv("FFMPM_P2G_MINB"); return e ? atoi(e) : 5; }();
#define FFMPM_LAUNCH_P2G(MB)                                                                              \
  do {                                                                                                    \
    if (use_perm)                                                                                         \
      p2g_runs3_kernel<T, MB, true><<<blocks, P2G_RUN_WARPS * 32, 0, st>>>(cfg, s, n, B, grid, err);      \
    else                                                                                                  \
      p2g_runs3_kernel<T, MB, false><<<blocks, P2G_RUN_WARPS * 32, 0, st>>>(cfg, s, n, B, grid, err);     \
  } while (0)
  if (minb >= 6) FFMPM_LAUNCH_P2G(6);
  else if (minb == 4) FFMPM_LAUNCH_P2G(4);
  else FFMPM_LAUNCH_P2G(5);
#undef FFMPM_LAUNCH_P2G
  return 1;
}

}  // namespace ffmpm

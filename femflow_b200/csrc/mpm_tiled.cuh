// Tiled 3D G2P over binned particles (see mpm_bin.cuh for the key layout); the binned
// P2G lives in mpm_p2g_runs.cuh.
//
// G2P: persistent CTAs pull active tiles (4x4x4 base cells) from a work counter, stage
// the tile's 6x6x6 velocity block in shared memory, each thread gathers its particle's
// 3x3x3 stencil from there, and the updated state is written to the OTHER particle
// buffer at the binned slot -- so the state is physically in cell order for the next
// substep.  The kernel also emits the NEXT substep's cell key and within-cell rank of
// every particle (it is the one place that knows the advected position), which removes
// a whole pass over the positions from ffmpm_bin.
//
// Stores are plain coalesced 32-bit stores: a variant that parked each round's results in
// shared memory and drained them with cp.async.bulk (TMA) stores was measured slower
// (1.09 ms against 0.84 ms, profiles/r01f) and removed.
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
  if (src.material) dst.material[slot] = src.material[p];
  if (with_jp && src.Jp) dst.Jp[slot] = src.Jp[p];
}

constexpr int G2P_PRE_PLANES = 17;   // x3 F9 mass mu0 lam0 id Jp (+ the word holding the material row: ROWS)

__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}

// PRE (fp32): every thread prefetches the 17 input scalars of ITS particle of the next round
// into shared memory with 4-byte cp.async (the gather through `perm` rules out bulk copies)
// while the current round computes; the permutation entry itself is fetched two rounds
// ahead.  Each thread only ever reads what it copied itself, so no barrier is involved.
// ROWS: the state carries 1-byte material rows (table mode); compiled out otherwise, the kernel sits
// exactly at its 64-register budget.
// IDX32: exact fp32 cell indexing known at compile time (cfg.index_fp32; see base_fx).
// KEEPF (3D snow): F moves to its new slot unchanged; snow_project3_kernel (mpm_svd3.cuh) forms (I + dt C) F in fp64
// from it and the new C and projects it.  The instantiations of the neo-hookean path are untouched by the flag.
template <typename T, int MIN_BLOCKS, bool PRE = false, bool ROWS = false, bool IDX32 = false, bool KEEPF = false>
__global__ void __launch_bounds__(G2P_THREADS, MIN_BLOCKS) g2p_tiled3_kernel(DevCfg cfg, StateView<T> src, StateView<T> dst,
                                                                 BinBuffers B, const T* __restrict__ grid, ErrRec* err) {
  using V4 = typename Vec4<T>::type;
  __shared__ V4 tile[TNODES3];
  __shared__ int s_work;
  static_assert(!PRE || sizeof(T) == 4, "the cp.async prefetch is for the fp32 build");
  __shared__ __align__(16) float pre[PRE ? 2 : 1][PRE ? G2P_PRE_PLANES + (ROWS ? 1 : 0) : 1][PRE ? G2P_THREADS : 1];
  int pre_buf = 0;
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
    const int r_first = (start / G2P_THREADS) * G2P_THREADS;
    // PRE: post this thread's cp.async for the particle at permutation entry q into pre[buf]
    auto prefetch = [&](int q, int buf) {
      if constexpr (PRE) {
        if (q >= 0) {
          const int tid = threadIdx.x;
#pragma unroll
          for (int c = 0; c < 3; ++c) cp_async4(&pre[buf][c][tid], src.x + c * ss + q);
#pragma unroll
          for (int c = 0; c < 9; ++c) cp_async4(&pre[buf][3 + c][tid], src.F + c * ss + q);
          if (src.mass) cp_async4(&pre[buf][12][tid], src.mass + q);
          if (src.mu0) cp_async4(&pre[buf][13][tid], src.mu0 + q);
          if (src.lam0) cp_async4(&pre[buf][14][tid], src.lam0 + q);
          if (src.id) cp_async4(&pre[buf][15][tid], src.id + q);
          if (src.Jp) cp_async4(&pre[buf][16][tid], src.Jp + q);
          // a row is 1 byte, cp.async moves >= 4: fetch the aligned word around it
          if constexpr (ROWS) cp_async4(&pre[buf][G2P_PRE_PLANES][tid], src.material + (q & ~3));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    };
    auto perm_at = [&](int rb) -> int {
      const int sl = rb + (int)threadIdx.x;
      return (sl >= start && sl < end) ? B.perm[sl] : -1;
    };
    int q_cur = -1, q_nxt = -1;
    if constexpr (PRE) {
      q_cur = perm_at(r_first);
      q_nxt = r_first + G2P_THREADS < end ? perm_at(r_first + G2P_THREADS) : -1;
      prefetch(q_cur, pre_buf);
    }
    // rounds are aligned to multiples of 128 slots so that a round's plane segment is 16-byte aligned
    for (int rbase = r_first; rbase < end; rbase += G2P_THREADS) {
      const int slot = rbase + threadIdx.x;
      const bool mine = slot >= start && slot < end;
      int next_key = -1;
      bool leaving = false, urgent = false;
      T o[24];
      T cm = 0, cmu = 0, cl = 0, cjp = 0;
      int cid = 0;
      unsigned row = 0;
      int q_nn = -1;
      if constexpr (PRE) {
        const bool more = rbase + G2P_THREADS < end;
        if (more) {
          prefetch(q_nxt, pre_buf ^ 1);                                   // next round's inputs
          if (rbase + 2 * G2P_THREADS < end) q_nn = perm_at(rbase + 2 * G2P_THREADS);   // and the entry after that
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
      }
      if (mine) {
        T x0, x1, x2, f00, f01, f02, f10, f11, f12, f20, f21, f22;
        if constexpr (PRE) {
          const int tid = threadIdx.x;
          const float(*pb)[G2P_THREADS] = pre[pre_buf];
          x0 = pb[0][tid]; x1 = pb[1][tid]; x2 = pb[2][tid];
          f00 = pb[3][tid]; f01 = pb[4][tid]; f02 = pb[5][tid];
          f10 = pb[6][tid]; f11 = pb[7][tid]; f12 = pb[8][tid];
          f20 = pb[9][tid]; f21 = pb[10][tid]; f22 = pb[11][tid];
          if (src.mass) cm = pb[12][tid];
          if (src.mu0) cmu = pb[13][tid];
          if (src.lam0) cl = pb[14][tid];
          if (src.id) cid = __float_as_int(pb[15][tid]);
          if (src.Jp) cjp = pb[16][tid];
          if constexpr (ROWS) row = (__float_as_uint(pb[G2P_PRE_PLANES][tid]) >> ((q_cur & 3) * 8)) & 0xffu;
        } else {
          const long long p = B.perm[slot];
          // carried planes first: their latency hides behind the gather
          if (src.mass) cm = src.mass[p];
          if (src.mu0) cmu = src.mu0[p];
          if (src.lam0) cl = src.lam0[p];
          if (src.id) cid = src.id[p];
          if (src.Jp) cjp = src.Jp[p];
          if constexpr (ROWS) row = src.material[p];
          x0 = src.x[p]; x1 = src.x[ss + p]; x2 = src.x[2 * ss + p];
          f00 = src.F[0 * ss + p]; f01 = src.F[1 * ss + p]; f02 = src.F[2 * ss + p];
          f10 = src.F[3 * ss + p]; f11 = src.F[4 * ss + p]; f12 = src.F[5 * ss + p];
          f20 = src.F[6 * ss + p]; f21 = src.F[7 * ss + p]; f22 = src.F[8 * ss + p];
        }
        int gx, gy, gz;
        T fx, fy, fz;
        if constexpr (IDX32 && sizeof(T) == 4) {
          base_fx_f32(x0, (float)cfg.inv_dx, gx, fx);
          base_fx_f32(x1, (float)cfg.inv_dx, gy, fy);
          base_fx_f32(x2, (float)cfg.inv_dx, gz, fz);
        } else {
          base_fx(x0, cfg, gx, fx);
          base_fx(x1, cfg, gy, fy);
          base_fx(x2, cfg, gz, fz);
        }
        const int cb = ((gx - cfg.origin[0] - ox) * TN3 + (gy - cfg.origin[1] - oy)) * TN3 + (gz - cfg.origin[2] - oz);
        T vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22;
        g2p_accumulate3<T>([&](int i, int j, int k) { return tile[cb + (i * TN3 + j) * TN3 + k]; }, fx, fy, fz,
                           vx, vy, vz, c00, c01, c02, c10, c11, c12, c20, c21, c22);
        const T s4 = (T)(4.0 * cfg.inv_dx);
        c00 *= s4; c01 *= s4; c02 *= s4; c10 *= s4; c11 *= s4; c12 *= s4; c20 *= s4; c21 *= s4; c22 *= s4;
        const T dt = (T)cfg.dt;
        // F <- (I + dt C) F   (three_d/g2p.py:46)
        const T m00 = (T)1 + dt * c00, m01 = dt * c01, m02 = dt * c02;
        const T m10 = dt * c10, m11 = (T)1 + dt * c11, m12 = dt * c12;
        const T m20 = dt * c20, m21 = dt * c21, m22 = (T)1 + dt * c22;
        o[0] = x0 + dt * vx; o[1] = x1 + dt * vy; o[2] = x2 + dt * vz;   // three_d/g2p.py:45
        o[3] = vx; o[4] = vy; o[5] = vz;
        o[6] = c00; o[7] = c01; o[8] = c02; o[9] = c10; o[10] = c11; o[11] = c12; o[12] = c20; o[13] = c21; o[14] = c22;
        if constexpr (KEEPF) {
          o[15] = f00; o[16] = f01; o[17] = f02; o[18] = f10; o[19] = f11; o[20] = f12; o[21] = f20; o[22] = f21; o[23] = f22;
        } else {
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
        int gbx = 0;
        next_key = bin_key_of<T, IDX32>(cfg, B, o[0], o[1], o[2], &gbx);
        B.keys[slot] = next_key;
        leaving = gbx < cfg.own_lo || gbx >= cfg.own_hi;
        urgent = (long long)gbx - cfg.own_hi >= cfg.own_slack || (long long)cfg.own_lo - gbx > cfg.own_slack;
      }
      // next substep's histogram + within-cell rank (same scheme as bin_count_kernel): the atomic is posted here and
      // its return value is used after the 24 stores below, which cover its round trip to L2
      const BinRankTicket ticket = bin_rank_issue(B, next_key);
      {
        // slabs: count the particles that now belong to a neighbour rank (read by the migration logic)
        const unsigned lm = __ballot_sync(0xffffffffu, leaving);
        if (lm) {          // rare: only near a slab cut
          const unsigned um = __ballot_sync(lm, leaving && urgent);
          if ((threadIdx.x & 31) == (unsigned)(__ffs(lm) - 1)) {
            atomicAdd(&B.counters[3], __popc(lm));
            if (um) atomicAdd(&B.counters[4], __popc(um));
          }
        }
      }
      {
        if (mine) {
#pragma unroll
          for (int k = 0; k < 3; ++k) { dst.x[k * ds + slot] = o[k]; dst.v[k * ds + slot] = o[3 + k]; }
#pragma unroll
          for (int k = 0; k < 9; ++k) { dst.C[k * ds + slot] = o[6 + k]; dst.F[k * ds + slot] = o[15 + k]; }
          if (src.mass) dst.mass[slot] = cm;
          if (src.mu0) dst.mu0[slot] = cmu;
          if (src.lam0) dst.lam0[slot] = cl;
          if (src.id) dst.id[slot] = cid;
          if (src.Jp) dst.Jp[slot] = cjp;
          if constexpr (ROWS) dst.material[slot] = (unsigned char)row;
        }
      }
      bin_rank_finish(B, ticket, next_key, slot);
      if constexpr (PRE) {
        q_cur = q_nxt; q_nxt = q_nn;
        pre_buf ^= 1;
      }
    }
  }
}

template <typename T>
int g2p_tiled(const DevCfg& cfg, const StateView<T>& src, const StateView<T>& dst, long long n, BinBuffers& B,
              const T* grid, ErrRec* err, int sm_count, int blocks_per_sm, cudaStream_t st) {
  (void)n;
  cudaMemsetAsync(&B.counters[2], 0, sizeof(int32_t), st);
  int blocks = min(B.n_tiles + 1, sm_count * blocks_per_sm);
  if (cfg.model == 1) {   // snow: F is carried unchanged, the caller projects it afterwards (mpm_svd3.cuh)
    constexpr int MB = sizeof(T) == 4 ? 8 : 4;
    if (src.material) g2p_tiled3_kernel<T, MB, false, true, false, true><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);
    else g2p_tiled3_kernel<T, MB, false, false, false, true><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);
    return 1;
  }
  if constexpr (sizeof(T) == 4) {
    // FFMPM_G2P_PRE=0 disables the cp.async input prefetch (64 registers, 8 CTAs per SM either way)
    static int prefetch = [] { const char* e = getenv("FFMPM_G2P_PRE"); return e ? atoi(e) : 1; }();
    const bool rows = src.material != nullptr;
    const bool i32 = cfg.index_fp32 != 0;
#define FFMPM_G2P(PRE_, ROWS_)                                                                                          \
  do {                                                                                                                  \
    if (i32) g2p_tiled3_kernel<T, 8, PRE_, ROWS_, true><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);   \
    else g2p_tiled3_kernel<T, 8, PRE_, ROWS_, false><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);      \
  } while (0)
    if (prefetch && rows) FFMPM_G2P(true, true);
    else if (prefetch) FFMPM_G2P(true, false);
    else if (rows) FFMPM_G2P(false, true);
    else FFMPM_G2P(false, false);
#undef FFMPM_G2P
  } else if (src.material) {
    g2p_tiled3_kernel<T, 4, false, true><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);
  } else {
    g2p_tiled3_kernel<T, 4, false, false><<<blocks, G2P_THREADS, 0, st>>>(cfg, src, dst, B, grid, err);
  }
  return 1;
}

}  // namespace ffmpm

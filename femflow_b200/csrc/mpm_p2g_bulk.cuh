// P2G in physical particle order with TMA-prefetched particle state (fp32 build).
//
// Same warp-autonomous algorithm as mpm_p2g_runs.cuh (lane per particle -> runs of
// equal base cell -> lane per (run, x-slab) with register accumulation -> one vector
// RED per node), but the 27 SoA state planes of a warp's 64-particle window are
// brought into shared memory by the copy engine: one elected lane posts 27 bulk async
// copies (cp.async.bulk.shared::cluster.global, 256 contiguous bytes each) against an
// mbarrier, for window w+1 while the warp is still computing window w.  The loads
// need no registers, no address arithmetic per lane, and their latency is off the
// critical path regardless of occupancy.  Physical order is what makes this possible:
// a window is a contiguous, 256-byte aligned segment of every plane (no permutation
// gather); the reordering G2P keeps that order cell-sorted up to one substep of motion.
#pragma once
#include <type_traits>

#include "mpm_p2g_pair.cuh"
#include "mpm_p2g_runs.cuh"

namespace ffmpm {

template <int NBUF, int RAWP = P2G_NPLANES, int PAIR = 0>
struct P2GBulkWarp {
  alignas(128) float raw[NBUF][RAWP][P2G_WINDOW];   // RAWP = 24: no room for material planes (table / config material)
  std::conditional_t<(PAIR > 0), P2GPairSlab, P2GWarpSlab<float>> slab;   // PAIR: pair-major payload (mpm_p2g_pair.cuh)
  alignas(16) unsigned char mat[NBUF][P2G_WINDOW];   // material rows of the window (table mode)
  alignas(8) unsigned long long bar[NBUF];
};

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load_s(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(sdst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}

// Base pointer of every state plane, resolved once on the host and passed as a kernel
// parameter (constant bank): the per-window issue loop is then one uniform add + one
// UBLKCP per plane instead of a pointer-selection sequence (17 % of the kernel's
// instructions in profiles/r01_final before this).
struct P2GPlanes {
  const float* p[P2G_NPLANES];
  int n;                           // planes to prefetch: 27 with material planes, else 24
  const unsigned char* material;   // table mode: row per particle (nullptr: row 0)
  const float* table;              // table mode: [3][MAT_ROWS]
};

inline P2GPlanes p2g_planes_of(const StateView<float>& s) {
  P2GPlanes P;
  const long long st = s.stride;
  for (int k = 0; k < 3; ++k) { P.p[P2G_X + k] = s.x + k * st; P.p[P2G_V + k] = s.v + k * st; }
  for (int k = 0; k < 9; ++k) { P.p[P2G_C + k] = s.C + k * st; P.p[P2G_F + k] = s.F + k * st; }
  P.p[P2G_MASS] = s.mass; P.p[P2G_MU] = s.mu0; P.p[P2G_LAM] = s.lam0;
  const int mode = mat_mode_of(s);
  P.n = mode == MAT_PLANES ? P2G_NPLANES : P2G_MASS;
  P.material = mode == MAT_TABLE ? s.material : nullptr;
  P.table = mode == MAT_TABLE ? s.mat_table : nullptr;
  return P;
}

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// LDGSTS: 0 = bulk copies by the TMA engine (cp.async.bulk + mbarrier, one elected lane);
//         1 = per-lane 16-byte cp.async (LDGSTS): two planes per warp instruction, 14
//             instructions per window -- fewer issue slots than 27 elected UBLKCP sequences.
// RAWP / SMW: planes held per window and resident warps per SM the launch bounds ask for.  <27, 16> is the
// measured default (128 registers); <24, 20> (no material planes: 11.1 KB of shared memory per warp, 5 CTAs
// of 4 warps, <= 96 registers) is the occupancy experiment behind FFMPM_P2G_VARIANT=6.
// PAIR >= 1: phase 2 walks the runs two particles per instruction with packed fp32 (FFMA2; mpm_p2g_pair.cuh);
// PAIR == 2: phase 1 too -- the stress of the two particles a lane owns in a window is evaluated in packed
// fp32; PAIR == 3: ... in the left form only, chosen at compile time.  FFMPM_P2G_VARIANT=7 / 8, written after this round's GPU budget was spent: not yet measured.
template <int WARPS, int NBUF, int LDGSTS, int RAWP = P2G_NPLANES, int SMW = 16, int PAIR = 0>
__global__ void __launch_bounds__(WARPS * 32, SMW / WARPS)
p2g_bulk3_kernel(DevCfg cfg, P2GPlanes planes, long long n, float* __restrict__ grid, ErrRec* err, int wpw) {
  using T = float;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using WarpMem = P2GBulkWarp<NBUF, RAWP, PAIR>;
  WarpMem* warps = reinterpret_cast<WarpMem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpMem& W = warps[warp];
  auto& S = W.slab;
  const T dx = (T)cfg.dx;
  const int ny = cfg.n[1], nz = cfg.n[2];
  const int n_planes = planes.n;
  const bool has_mat = n_planes == P2G_NPLANES || planes.table != nullptr;
  int run_cap = 0;
  if constexpr (PAIR > 0) { run_cap = wpw >> 16; wpw &= 0xffff; }   // experimental kernels: run cap rides in the high half
  const int n_windows = (int)((n + P2G_WINDOW - 1) / P2G_WINDOW);
  // wpw > 0: every warp owns `wpw` consecutive windows and the CTA retires after them (a finite
  // grid lets the block scheduler interleave CTAs of kernels running on other streams);
  // wpw == 0: persistent, windows strided over the whole grid.
  const int total_warps = wpw > 0 ? 1 : gridDim.x * WARPS;
  const int first = wpw > 0 ? (blockIdx.x * WARPS + warp) * wpw : blockIdx.x * WARPS + warp;
  const int last_excl = wpw > 0 ? min(n_windows, first + wpw) : n_windows;

  if (lane == 0) {
    for (int b = 0; b < NBUF; ++b) mbar_init(&W.bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  auto issue = [&](int win, int buf) {
    if (LDGSTS) {
      const long long w0 = (long long)win * P2G_WINDOW + (lane & 15) * 4;
      const int hi = lane >> 4;
#pragma unroll
      for (int j = 0; j < (P2G_NPLANES + 1) / 2; ++j) {
        const int k0 = 2 * j, k1 = 2 * j + 1;
        const float* src = (hi && k1 < P2G_NPLANES) ? planes.p[k1 < P2G_NPLANES ? k1 : k0] : planes.p[k0];
        const int k = hi ? k1 : k0;
        if (k < n_planes && (RAWP == P2G_NPLANES || k < RAWP)) cp_async16(&W.raw[buf][k][(lane & 15) * 4], src + w0);
      }
      if (planes.material && lane < P2G_WINDOW / 16)
        cp_async16(&W.mat[buf][lane * 16], planes.material + (long long)win * P2G_WINDOW + lane * 16);
      asm volatile("cp.async.commit_group;" ::: "memory");
      return;
    }
    // one lane posts the whole window: n_planes x 256 B
    if (lane == 0) {
      mbar_expect_tx(&W.bar[buf], (unsigned)n_planes * P2G_WINDOW * 4u + (planes.material ? (unsigned)P2G_WINDOW : 0u));
      const long long w0 = (long long)win * P2G_WINDOW;
      if (planes.material) bulk_load_s(&W.mat[buf][0], planes.material + w0, P2G_WINDOW, &W.bar[buf]);
#pragma unroll
      for (int k = 0; k < P2G_NPLANES; ++k)
        if (k < n_planes && (RAWP == P2G_NPLANES || k < RAWP)) bulk_load_s(&W.raw[buf][k][0], planes.p[k] + w0, P2G_WINDOW * 4u, &W.bar[buf]);
    }
  };

  if (first < last_excl) issue(first, 0);
  int it = 0;
  for (int win = first; win < last_excl; win += total_warps, ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    const unsigned parity = NBUF == 2 ? ((it >> 1) & 1) : (it & 1);
    const int w0 = win * P2G_WINDOW;
    const int cnt = (int)min((long long)P2G_WINDOW, n - w0);
    if (NBUF == 2 && win + total_warps < last_excl) issue(win + total_warps, buf ^ 1);
    if (LDGSTS) {
      if (NBUF == 2 && win + total_warps < last_excl) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    } else {
      mbar_wait(&W.bar[buf], parity);
    }
    // ---- phase 1: lane per particle, state from the prefetched slab ----
    int node[2];
    if constexpr (PAIR >= 2) {
      // both particles of the lane (slots lane and lane + 32) in one packed evaluation
      const bool live_a = lane < cnt, live_b = 32 + lane < cnt;
      auto getter = [&](int idx) {
        return [&, idx](int k) -> T {
          const int row = planes.material ? (int)W.mat[buf][idx] : 0;
          if (k >= P2G_MASS && planes.table) return __ldg(planes.table + (k - P2G_MASS) * MAT_ROWS + row);
          if constexpr (RAWP < P2G_NPLANES) { if (k >= RAWP) return (T)0; }
          return W.raw[buf][k][idx];
        };
      };
      P2GPairParker park{S.pay, S.node0, lane, ny, nz, dx, {-1, -1}};
      p2g_prepare3_pair_sink<(PAIR == 3 ? 3 : 0)>(cfg, getter(lane), getter(32 + lane), has_mat, live_a, live_b, park);
      node[0] = park.node[0];
      node[1] = park.node[1];
      if (!live_a) p2g_park_pair_zero(S.pay, lane);
      if (!live_b) p2g_park_pair_zero(S.pay, 32 + lane);
    } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = h * 32 + lane;
      node[h] = -1;
      if (idx < cnt) {
        const int row = planes.material ? (int)W.mat[buf][idx] : 0;
        P2GParticle3<T> q = p2g_prepare3_from<T>(
            cfg,
            [&](int k) -> T {
              if (k >= P2G_MASS && planes.table) return __ldg(planes.table + (k - P2G_MASS) * MAT_ROWS + row);
              if constexpr (RAWP < P2G_NPLANES) { if (k >= RAWP) return (T)0; }   // RAWP = 24 is launched only without material planes
              return W.raw[buf][k][idx];
            },
            has_mat, 1.0);
        if constexpr (PAIR > 0) node[h] = p2g_park_pair(S.pay, S.node0, q, idx, dx, ny, nz);
        else node[h] = p2g_park(S, q, idx, dx, ny, nz);
      } else if constexpr (PAIR > 0) {
        p2g_park_pair_zero(S.pay, idx);   // the last window's tail: a masked partner must read finite values
      }
    }
    }
    __syncwarp();
    // single buffer: the raw slab is free again -> prefetch the next window behind phase 2
    if (NBUF == 1 && win + total_warps < last_excl) issue(win + total_warps, 0);
    if constexpr (PAIR > 0) p2g_runs_phase2_pair(S, node, cnt, lane, ny, nz, grid, run_cap);
    else p2g_runs_phase2<T>(S, node, cnt, lane, ny, nz, grid);
    __syncwarp();   // the payload slab is rewritten by the next window
  }
}

template <int WARPS, int NBUF, int LDGSTS, int RAWP = P2G_NPLANES, int SMW = 16, int PAIR = 0>
static bool p2g_bulk_launch(const DevCfg& cfg, const StateView<float>& s, long long n, float* grid, ErrRec* err,
                            int sm_count, int blocks_per_sm, cudaStream_t st, int run_cap = 0) {
  const size_t smem = sizeof(P2GBulkWarp<NBUF, RAWP, PAIR>) * WARPS;
  // function attributes are per device: a process that drives several GPUs configures each once
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(p2g_bulk3_kernel<WARPS, NBUF, LDGSTS, RAWP, SMW, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return false;
    configured[dev] = true;
  }
  long long windows = (n + P2G_WINDOW - 1) / P2G_WINDOW;
  static int wpw = [] { const char* e = getenv("FFMPM_P2G_WPW"); return e ? atoi(e) : 16; }();
  int blocks;
  if (wpw > 0) {
    blocks = (int)((windows + (long long)WARPS * wpw - 1) / ((long long)WARPS * wpw));
  } else {
    long long want = (windows + WARPS - 1) / WARPS;
    long long cap = (long long)sm_count * blocks_per_sm;
    blocks = (int)(want < cap ? want : cap);
  }
  if (blocks < 1) blocks = 1;
  const int wpw_arg = PAIR > 0 ? ((wpw & 0xffff) | (run_cap << 16)) : wpw;
  p2g_bulk3_kernel<WARPS, NBUF, LDGSTS, RAWP, SMW, PAIR><<<blocks, WARPS * 32, smem, st>>>(cfg, p2g_planes_of(s), n, grid, err, wpw_arg);
  return true;
}

// True when the state layout allows bulk copies: 16-byte aligned planes, stride multiple of the window.
inline bool p2g_bulk_eligible(const DevCfg& cfg, const StateView<float>& s) {
  if (cfg.model != 0 || cfg.dim != 3) return false;
  if ((s.stride % P2G_WINDOW) != 0) return false;
  const void* planes[] = {s.x, s.v, s.C, s.F, s.mass, s.mu0, s.lam0, s.material};
  for (const void* q : planes)
    if (((uintptr_t)q & 15) != 0) return false;
  return true;
}

}  // namespace ffmpm

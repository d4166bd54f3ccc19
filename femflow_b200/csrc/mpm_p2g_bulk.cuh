// P2G in physical particle order with the particle state prefetched into shared memory (fp32 build): the
// production P2G kernel.
//
// Same warp-autonomous algorithm as mpm_p2g_runs.cuh (lanes per particle -> runs of equal base cell -> lane per
// (run, x-slab) with register accumulation -> one vector RED per node), with three differences:
//   * the 24 (27 with material planes) SoA state planes of a warp's 64-particle window are copied into shared
//     memory by per-lane 16-byte cp.async (SASS LDGSTS) for window w+1 while phase 2 of window w runs: the loads
//     need no registers and their latency is off the critical path regardless of occupancy.  Physical order is
//     what makes this possible: a window is a contiguous, 256-byte aligned segment of every plane (no permutation
//     gather); the reordering G2P keeps that order cell-sorted up to one substep of motion.  (A variant that moved
//     the window with 27 TMA bulk copies against an mbarrier measured 0.85 ms against 0.72 ms: an elected
//     cp.async.bulk costs ~10 issue slots of uniform-register traffic and this kernel is issue-bound; removed.)
//   * a lane owns the ADJACENT slots 2*lane and 2*lane+1, so each plane is read with one 8-byte LDS for both
//     particles and the run heads come from a shuffle instead of a shared-memory round trip;
//   * the stress is the left-form fp32 series of mpm_math.cuh with its degree picked PER WARP (a warp pays for
//     its most strained particle anyway): one redux.sync on the strain norms, then straight-line code with
//     immediate coefficients -- no per-lane branching, no divergence bookkeeping.  All uniform scalars (material
//     of a one-material scene, dt vol 4/dx^2, dx) are folded on the way in, so that the parked affine matrix is
//     18 FFMA per particle.  Strains beyond the series (||F F^T - I||_F >= 0.15) take the fp64 Newton polar path
//     per particle through a non-inlined call that keeps its registers out of the hot path.
#pragma once
#include "mpm_p2g_runs.cuh"

namespace ffmpm {

template <int NBUF>
struct P2GBulkWarp {
  alignas(128) float raw[NBUF][P2G_NPLANES][P2G_WINDOW];
  P2GWarpSlab<float> slab;
  alignas(16) unsigned char mat[NBUF][P2G_WINDOW];   // material rows of the window (table mode)
};

// Base pointer of every state plane, resolved once on the host and passed as a kernel
// parameter (constant bank): the per-window issue loop is then one add + one LDGSTS per plane pair.
struct P2GPlanes {
  const float* p[P2G_NPLANES];
  int n;                           // planes to prefetch: 27 with material planes, else 24
  const unsigned char* material;   // table mode: row per particle (nullptr: row 0)
  const float* table;              // table mode: [3][MAT_ROWS]
  long long stride;                // contiguous(): plane k of x, v, C, F starts at p[0] + k * stride
};

// The 24 planes x3 v3 C9 F9 sit back to back at one stride (MpmSolver carves them from one allocation): the
// prefetch then needs ONE per-lane source pointer and twelve constant steps instead of a pointer per plane.
inline bool p2g_planes_contiguous(const StateView<float>& s) {
  const long long st = s.stride;
  return s.v == s.x + 3 * st && s.C == s.v + 3 * st && s.F == s.C + 9 * st;
}

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
  P.stride = s.stride;
  return P;
}

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// Large-strain fallback of one particle: affine * dx by the fp64 Newton polar path (mpm_math.cuh), re-reading
// F and C from the window image.  Deliberately not inlined.
__device__ __noinline__ void p2g_affine_fp64(const float (*raw)[P2G_WINDOW], int idx, double mu, double lam, double mass,
                                             double k, double dx, float* out9) {
  Mat3<double> F{raw[P2G_F + 0][idx], raw[P2G_F + 1][idx], raw[P2G_F + 2][idx], raw[P2G_F + 3][idx], raw[P2G_F + 4][idx],
                 raw[P2G_F + 5][idx], raw[P2G_F + 6][idx], raw[P2G_F + 7][idx], raw[P2G_F + 8][idx]};
  Mat3<double> C{raw[P2G_C + 0][idx], raw[P2G_C + 1][idx], raw[P2G_C + 2][idx], raw[P2G_C + 3][idx], raw[P2G_C + 4][idx],
                 raw[P2G_C + 5][idx], raw[P2G_C + 6][idx], raw[P2G_C + 7][idx], raw[P2G_C + 8][idx]};
  const Mat3<double> A = fixed_corotated_affine3(F, C, mu, lam, mass, k);
  out9[0] = (float)(A.a00 * dx); out9[1] = (float)(A.a01 * dx); out9[2] = (float)(A.a02 * dx);
  out9[3] = (float)(A.a10 * dx); out9[4] = (float)(A.a11 * dx); out9[5] = (float)(A.a12 * dx);
  out9[6] = (float)(A.a20 * dx); out9[7] = (float)(A.a21 * dx); out9[8] = (float)(A.a22 * dx);
}

// Run heads + run table for a window whose lane l owns slots 2l and 2l+1 (node[h] = base node of slot 2l+h, -1 when
// outside the grid), then phase 2.  Twin of p2g_runs_phase2 for the adjacent-slot ownership.
__device__ __forceinline__ void p2g_runs_phase2_adjacent(P2GWarpSlab<float>& S, const int (&node)[2], int cnt, int lane,
                                                         int ny, int nz, float* __restrict__ grid) {
  int prev0 = __shfl_up_sync(0xffffffffu, node[1], 1);
  if (lane == 0) prev0 = -2;
  unsigned h0 = __ballot_sync(0xffffffffu, 2 * lane < cnt && node[0] != prev0);
  unsigned h1 = __ballot_sync(0xffffffffu, 2 * lane + 1 < cnt && node[1] != node[0]);
  int n_runs = __popc(h0) + __popc(h1);
  // dense cells (a settled column: tens of particles per cell): a window holds two or three long runs and most lanes
  // of phase 2 would idle.  Cut the runs every 8 / 16 slots as well: up to 10 (run, slab) triples, at the price of one
  // more set of REDs per cut.
  const int cut = n_runs <= 2 ? 8 : (n_runs <= 5 ? 16 : 0);
  if (cut) {
    h0 = __ballot_sync(0xffffffffu, 2 * lane < cnt && (node[0] != prev0 || ((2 * lane) & (cut - 1)) == 0));
    n_runs = __popc(h0) + __popc(h1);
  }
  const unsigned below = (1u << lane) - 1u;
  const int r_lo = __popc(h0 & below) + __popc(h1 & below);
  const unsigned mine0 = (h0 >> lane) & 1u, mine1 = (h1 >> lane) & 1u;
  if (n_runs * 3 > 32) {
    // fragmented window (particles that changed cell since the G2P that placed them): merge the fragments
    const int merged = p2g_sort_window(S, node, cnt, lane);
    if (merged >= 0) {
      __syncwarp();
      p2g_accumulate_runs<float, true>(S, merged, lane, ny, nz, grid);
      return;
    }
  }
  if (mine0) S.run_start[r_lo] = 2 * lane;
  if (mine1) S.run_start[r_lo + mine0] = 2 * lane + 1;
  if (lane == 0) S.run_start[n_runs] = cnt;
  __syncwarp();
  p2g_accumulate_runs<float>(S, n_runs, lane, ny, nz, grid);
}

// CONTIG: the 24 state planes are one strided block (p2g_planes_contiguous) and there are no material planes;
// IDX32: cfg.index_fp32 (power-of-two inv_dx: exact fp32 cell indexing), resolved at compile time so that the
// fp64 indexing code is not even fetched.
template <int WARPS, int NBUF, bool CONTIG, bool IDX32>
__global__ void __launch_bounds__(WARPS * 32, 16 / WARPS)
p2g_bulk3_kernel(DevCfg cfg, P2GPlanes planes, long long n, float* __restrict__ grid, ErrRec* err, int wpw) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using WarpMem = P2GBulkWarp<NBUF>;
  WarpMem* warps = reinterpret_cast<WarpMem*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpMem& W = warps[warp];
  auto& S = W.slab;
  const int ny = cfg.n[1], nz = cfg.n[2];
  const int n_planes = planes.n;
  const bool mat_planes = n_planes == P2G_NPLANES, mat_table = planes.table != nullptr;
  const int n_windows = (int)((n + P2G_WINDOW - 1) / P2G_WINDOW);
  // wpw > 0: every warp owns `wpw` consecutive windows and the CTA retires after them (a finite
  // grid lets the block scheduler interleave CTAs of kernels running on other streams);
  // wpw == 0: persistent, windows strided over the whole grid.
  const int total_warps = wpw > 0 ? 1 : gridDim.x * WARPS;
  const int first = wpw > 0 ? (blockIdx.x * WARPS + warp) * wpw : blockIdx.x * WARPS + warp;
  const int last_excl = wpw > 0 ? min(n_windows, first + wpw) : n_windows;

  // uniform scalars, folded once per thread (three_d/p2g.py:57-65, utils.py:120-135)
  const float dxf = (float)cfg.dx;
  const float hard = (float)cfg.hardening;                                             // constant hardening: a multiplier (quirk 8)
  const double kd = (cfg.dt * cfg.volume) * (4.0 * cfg.inv_dx * cfg.inv_dx);
  const float nkdx = -(float)kd * dxf;                                                 // -(dt vol 4/dx^2) dx
  const float m_u = (float)cfg.mass, mu_u = (float)(cfg.mu0 * cfg.hardening), lam_u = (float)(cfg.lam0 * cfg.hardening);

  // CONTIG: lane l copies the 16-byte chunk (l & 15) of planes 2j + (l >> 4), j = 0 .. 11
  const float* const lane_src = planes.p[0] + (long long)(lane >> 4) * planes.stride + (lane & 15) * 4;
  auto issue = [&](int win, int buf) {
    if constexpr (CONTIG) {
      const float* src = lane_src + (long long)win * P2G_WINDOW;
      float* dst = &W.raw[buf][lane >> 4][(lane & 15) * 4];
      const long long step = 2 * planes.stride;
#pragma unroll
      for (int j = 0; j < P2G_MASS / 2; ++j) cp_async16(dst + j * 2 * P2G_WINDOW, src + j * step);
      if (planes.material && lane < P2G_WINDOW / 16)
        cp_async16(&W.mat[buf][lane * 16], planes.material + (long long)win * P2G_WINDOW + lane * 16);
      asm volatile("cp.async.commit_group;" ::: "memory");
      return;
    }
    const long long w0 = (long long)win * P2G_WINDOW + (lane & 15) * 4;
    const int hi = lane >> 4;
#pragma unroll
    for (int j = 0; j < (P2G_NPLANES + 1) / 2; ++j) {
      const int k0 = 2 * j, k1 = 2 * j + 1;
      const float* src = (hi && k1 < P2G_NPLANES) ? planes.p[k1 < P2G_NPLANES ? k1 : k0] : planes.p[k0];
      const int k = hi ? k1 : k0;
      if (k < n_planes) cp_async16(&W.raw[buf][k][(lane & 15) * 4], src + w0);
    }
    if (planes.material && lane < P2G_WINDOW / 16)
      cp_async16(&W.mat[buf][lane * 16], planes.material + (long long)win * P2G_WINDOW + lane * 16);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (first < last_excl) issue(first, 0);
  int it = 0;
  for (int win = first; win < last_excl; win += total_warps, ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    const int w0 = win * P2G_WINDOW;
    const int cnt = (int)min((long long)P2G_WINDOW, n - w0);
    if (NBUF == 2 && win + total_warps < last_excl) {
      issue(win + total_warps, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const float (*R)[P2G_WINDOW] = W.raw[buf];
    const int i0 = 2 * lane;
    auto ld2 = [&](int k) -> float2 { return *reinterpret_cast<const float2*>(&R[k][i0]); };

    // ---- phase 1: lane per slot pair (2 lane, 2 lane + 1), state from the prefetched window ----
    int node[2];
    bool ok[2];
    float fx[2], fy[2], fz[2];
    {
      const float2 x0 = ld2(P2G_X), x1 = ld2(P2G_X + 1), x2 = ld2(P2G_X + 2);
      const float xs[2][3] = {{x0.x, x1.x, x2.x}, {x0.y, x1.y, x2.y}};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int gx, gy, gz;
        if constexpr (IDX32) {
          base_fx_f32(xs[h][0], (float)cfg.inv_dx, gx, fx[h]);
          base_fx_f32(xs[h][1], (float)cfg.inv_dx, gy, fy[h]);
          base_fx_f32(xs[h][2], (float)cfg.inv_dx, gz, fz[h]);
        } else {
          base_fx(xs[h][0], cfg, gx, fx[h]);
          base_fx(xs[h][1], cfg, gy, fy[h]);
          base_fx(xs[h][2], cfg, gz, fz[h]);
        }
        const int bx = gx - cfg.origin[0], by = gy - cfg.origin[1], bz = gz - cfg.origin[2];
        // utils.py:138-150 with res = G: base < 0 or base + 2 >= G -> RuntimeError (flagged by the binning / G2P)
        ok[h] = i0 + h < cnt && xs[h][0] == xs[h][0] && xs[h][1] == xs[h][1] && xs[h][2] == xs[h][2] &&
                bx >= 0 && by >= 0 && bz >= 0 && bx + 2 < cfg.n[0] && by + 2 < cfg.n[1] && bz + 2 < cfg.n[2];
        node[h] = ok[h] ? (bx * ny + by) * nz + bz : -1;
      }
    }
    Mat3<float> E[2];
    Sym3f G[2], H[2];
    float r2[2], jm1[2];
    {
      const float2 f0 = ld2(P2G_F + 0), f1 = ld2(P2G_F + 1), f2 = ld2(P2G_F + 2), f3 = ld2(P2G_F + 3), f4 = ld2(P2G_F + 4),
                   f5 = ld2(P2G_F + 5), f6 = ld2(P2G_F + 6), f7 = ld2(P2G_F + 7), f8 = ld2(P2G_F + 8);
      left_strain3(Mat3<float>{f0.x, f1.x, f2.x, f3.x, f4.x, f5.x, f6.x, f7.x, f8.x}, E[0], G[0], r2[0]);
      left_strain3(Mat3<float>{f0.y, f1.y, f2.y, f3.y, f4.y, f5.y, f6.y, f7.y, f8.y}, E[1], G[1], r2[1]);
      jm1[0] = jm1_of(E[0]);
      jm1[1] = jm1_of(E[1]);
    }
    // the warp's largest strain picks the series degree for all of its particles (NaN compares as beyond every tier)
    int tier;
    {
      const float a = ok[0] ? r2[0] : 0.0f, b = ok[1] ? r2[1] : 0.0f;
      const float worst = (a != a || b != b) ? 1.0f : fmaxf(a, b);
      tier = cfg.fp32_stress ? stress_tier_of(__uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(worst)))) : kStressTiers;
    }
    if (tier == 0) { H[0] = left_stress_h<0>(G[0]); H[1] = left_stress_h<0>(G[1]); }
    else if (tier == 1) { H[0] = left_stress_h<1>(G[0]); H[1] = left_stress_h<1>(G[1]); }
    else if (tier == 2) { H[0] = left_stress_h<2>(G[0]); H[1] = left_stress_h<2>(G[1]); }
    else { H[0] = left_stress_h<3>(G[0]); H[1] = left_stress_h<3>(G[1]); }   // also evaluated (and discarded) beyond the series
    {
      const float2 c0 = ld2(P2G_C + 0), c1 = ld2(P2G_C + 1), c2 = ld2(P2G_C + 2), c3 = ld2(P2G_C + 3), c4 = ld2(P2G_C + 4),
                   c5 = ld2(P2G_C + 5), c6 = ld2(P2G_C + 6), c7 = ld2(P2G_C + 7), c8 = ld2(P2G_C + 8);
      const float2 v0 = ld2(P2G_V), v1 = ld2(P2G_V + 1), v2 = ld2(P2G_V + 2);
      const Mat3<float> Cm[2] = {{c0.x, c1.x, c2.x, c3.x, c4.x, c5.x, c6.x, c7.x, c8.x},
                                 {c0.y, c1.y, c2.y, c3.y, c4.y, c5.y, c6.y, c7.y, c8.y}};
      const float vs[2][3] = {{v0.x, v1.x, v2.x}, {v0.y, v1.y, v2.y}};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!ok[h]) continue;    // outside the grid (or past the end of the last window): its run is skipped in phase 2
        const int idx = i0 + h;
        float m = m_u, mu = mu_u, lam = lam_u;
        if (mat_planes) {
          m = R[P2G_MASS][idx]; mu = R[P2G_MU][idx] * hard; lam = R[P2G_LAM][idx] * hard;
        } else if (mat_table) {
          const int row = planes.material ? (int)W.mat[buf][idx] : 0;
          m = __ldg(planes.table + row); mu = __ldg(planes.table + MAT_ROWS + row) * hard;
          lam = __ldg(planes.table + 2 * MAT_ROWS + row) * hard;
        }
        Mat3<float> A;
        const bool series = cfg.fp32_stress && (tier < kStressTiers || r2[h] < kPerturbationMaxR * kPerturbationMaxR);
        if (series) {
          // affine * dx = -(k dx) (2 mu h(G) + lam (J-1) J [ALL entries, quirk 2]) + (mass dx) C
          const float l = lam * jm1[h] * (1.0f + jm1[h]);
          affine3_assemble(H[h], Cm[h], nkdx * 2.0f * mu, nkdx * l, m * dxf, A);
        } else {
          float a9[9];
          p2g_affine_fp64(R, idx, (double)mu, (double)lam, (double)m, kd, cfg.dx, a9);
          A = Mat3<float>{a9[0], a9[1], a9[2], a9[3], a9[4], a9[5], a9[6], a9[7], a9[8]};
        }
        const int ph = p2g_pad(idx);
        S.pay[0][ph] = P2GVec4<float>{m * vs[h][0], m * vs[h][1], m * vs[h][2], m};
        S.pay[1][ph] = P2GVec4<float>{A.a00, A.a01, A.a02, fx[h]};
        S.pay[2][ph] = P2GVec4<float>{A.a10, A.a11, A.a12, fy[h]};
        S.pay[3][ph] = P2GVec4<float>{A.a20, A.a21, A.a22, fz[h]};
      }
      S.node0[i0] = node[0];
      S.node0[i0 + 1] = node[1];
    }
    __syncwarp();
    // single buffer: the window image is free again -> prefetch the next window behind phase 2
    if (NBUF == 1 && win + total_warps < last_excl) issue(win + total_warps, 0);
    p2g_runs_phase2_adjacent(S, node, cnt, lane, ny, nz, grid);
    __syncwarp();   // the payload slab is rewritten by the next window
  }
}

template <int WARPS, int NBUF, bool CONTIG, bool IDX32>
static bool p2g_bulk_launch_t(const DevCfg& cfg, const StateView<float>& s, long long n, float* grid, ErrRec* err,
                            int sm_count, int blocks_per_sm, cudaStream_t st) {
  const size_t smem = sizeof(P2GBulkWarp<NBUF>) * WARPS;
  // function attributes are per device: a process that drives several GPUs configures each once
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(p2g_bulk3_kernel<WARPS, NBUF, CONTIG, IDX32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return false;
    configured[dev] = true;
  }
  long long windows = (n + P2G_WINDOW - 1) / P2G_WINDOW;
  // work item = WARPS * wpw windows per CTA.  8 measured best on the 16.8 M-particle block: 4 / 6 / 8 / 10 / 16 windows per warp
  // -> 1.363 / 1.359 / 1.355 / 1.356 / 1.366 ms per substep (profiles/r02u_p2g_work_item_sweep.json): smaller items retire
  // CTAs more often, which lets the binning kernels of the internal stream in earlier
  static int wpw = [] { const char* e = getenv("FFMPM_P2G_WPW"); return e ? atoi(e) : 8; }();
  int blocks;
  if (wpw > 0) {
    blocks = (int)((windows + (long long)WARPS * wpw - 1) / ((long long)WARPS * wpw));
  } else {
    long long want = (windows + WARPS - 1) / WARPS;
    long long cap = (long long)sm_count * blocks_per_sm;
    blocks = (int)(want < cap ? want : cap);
  }
  if (blocks < 1) blocks = 1;
  p2g_bulk3_kernel<WARPS, NBUF, CONTIG, IDX32><<<blocks, WARPS * 32, smem, st>>>(cfg, p2g_planes_of(s), n, grid, err, wpw);
  return true;
}

template <int WARPS, int NBUF>
static bool p2g_bulk_launch(const DevCfg& cfg, const StateView<float>& s, long long n, float* grid, ErrRec* err,
                            int sm_count, int blocks_per_sm, cudaStream_t st) {
  const bool contig = p2g_planes_contiguous(s) && mat_mode_of(s) != MAT_PLANES;
  if (contig && cfg.index_fp32) return p2g_bulk_launch_t<WARPS, NBUF, true, true>(cfg, s, n, grid, err, sm_count, blocks_per_sm, st);
  if (contig) return p2g_bulk_launch_t<WARPS, NBUF, true, false>(cfg, s, n, grid, err, sm_count, blocks_per_sm, st);
  if (cfg.index_fp32) return p2g_bulk_launch_t<WARPS, NBUF, false, true>(cfg, s, n, grid, err, sm_count, blocks_per_sm, st);
  return p2g_bulk_launch_t<WARPS, NBUF, false, false>(cfg, s, n, grid, err, sm_count, blocks_per_sm, st);
}

// True when the state layout allows 16-byte window copies: 16-byte aligned planes, stride multiple of the window.
inline bool p2g_bulk_eligible(const DevCfg& cfg, const StateView<float>& s) {
  if (cfg.model != 0 || cfg.dim != 3) return false;
  if ((s.stride % P2G_WINDOW) != 0) return false;
  const void* planes[] = {s.x, s.v, s.C, s.F, s.mass, s.mu0, s.lam0, s.material};
  for (const void* q : planes)
    if (((uintptr_t)q & 15) != 0) return false;
  return true;
}

}  // namespace ffmpm

// The 2D default of the fp32 build: warp-window P2G and G2P in physical particle order (no binning, no sort).
//
// What bounded the thread-per-particle 2D kernels at BASELINE configs[1] (1 M particles, 1024^2, state resident in L2;
// profiles/r02p_2d_*): (1) every thread is one dependent chain "12 loads -> stress -> 9 reductions" with nothing in
// flight meanwhile (long_scoreboard 44 % of the stall samples, issue slots 31 % busy, 397 instructions per particle,
// much of it fp64 scalar set-up), and (2) the 9.4 M vector reductions themselves: the same kernel with the REDs
// compiled out ran the substep in 38 us instead of 60.  Both kernels here are organised like the 3D production P2G:
//   * a warp owns windows of 64 consecutive slots (lane l: slots 2l, 2l+1 -> one 8-byte LDS / STG per plane for both)
//     and PREFETCHES the next window's planes into its shared-memory image with per-lane 16-byte cp.async while it works
//     on the current one: the state loads are off the critical path whatever the occupancy;
//   * all uniform scalars are folded once per thread; the stress is the cancellation-free fp32 closed form of
//     mpm_math.cuh (fp64 closed form next to a reflection / on request), the G2P round trip is the identity for
//     det F > 0 (mpm_direct.cuh: g2p_particle2).
// P2G additionally sums the window's contributions ON CHIP before they leave the SM.  Neighbouring particles of a window
// scatter into overlapping 3x3 stencils, whatever their order inside the window: the warp accumulates them in a private
// shared-memory NODE TILE spanning the window's bounding box of base cells (+2) and flushes each touched node with ONE
// red.global.add.v4.f32 -- for a 64-particle window of a 4-per-cell block that is ~76 reductions instead of 576.  The
// tile is updated without atomics (fp32 shared-memory atomics are CAS loops on sm_100): in one step every lane adds ONE
// stencil offset (i, j) of its particle, so two lanes touch the same node only if their particles share a base cell;
// such lanes (found with match.any) take turns, one round per duplicate, and a __syncwarp() separates the steps.  A window
// whose bounding box does not fit the tile (particles in random order) scatters its particles directly, like
// p2g_scatter2_kernel: the kernel is correct for any order and fast for a spatially coherent one.
// Measured (B200, 1 M particles, graph replay): 43.0 us per substep against 59.6 us for the thread-per-particle kernels with
// the same arithmetic (profiles/r02s_2d_series.json); compute-sanitizer memcheck / racecheck clean (profiles/r02q, r02r).
// Reference loops: two_d/p2g.py:49-76, two_d/g2p.py:17-47.
#pragma once
#include <climits>

#include "mpm_common.cuh"
#include "mpm_direct.cuh"
#include "mpm_p2g_bulk.cuh"   // cp_async16

namespace ffmpm {

constexpr int W2_WINDOW = 64;        // slots per warp window (2 per lane)
constexpr int W2_WARPS = 4;          // warps per CTA
constexpr int W2_MIN_CTAS = 5;       // resident CTAs per SM the kernels are compiled for (<= 96 registers; 6 and 8 measured slower: spills)
constexpr int W2_TILE_NODES = 160;   // capacity of a warp's node tile (a 4-per-cell window in lattice order needs 4 x 19)
constexpr int W2_NPLANES = 12;       // x2 v2 C4 F4 at one stride
enum { W2_X = 0, W2_V = 2, W2_C = 4, W2_F = 8 };

struct W2Planes {
  float* base;        // plane 0 of x; plane k of (x, v, C, F) at base + k * stride
  float* jp;          // may be nullptr
  long long stride;
};

// x, v, C, F carved from one allocation (MpmSolver does), 16-byte aligned windows, config material scalars.
inline bool w2_eligible(const DevCfg& cfg, const StateView<float>& s) {
  if (cfg.dim != 2) return false;
  const long long st = s.stride;
  if (st % W2_WINDOW != 0) return false;
  if (!(s.v == s.x + 2 * st && s.C == s.v + 2 * st && s.F == s.C + 4 * st)) return false;
  if (mat_mode_of(s) != MAT_CFG) return false;
  if (((uintptr_t)s.x & 15) != 0 || ((uintptr_t)s.Jp & 15) != 0) return false;
  if (cfg.model == 1 && !s.Jp) return false;
  return true;
}

inline W2Planes w2_planes_of(const StateView<float>& s) { return W2Planes{s.x, s.Jp, s.stride}; }

// Large-rotation / on-request fallback of one particle: two_d stress in fp64 (mpm_math.cuh).  Not inlined: keeps its
// registers out of the hot path.
__device__ __noinline__ void w2_affine_fp64(float f00, float f01, float f10, float f11, float c00, float c01, float c10,
                                            float c11, double mu, double lam, double mass, double k, float* out4) {
  Mat2<double> F, C;
  F.a00 = f00; F.a01 = f01; F.a10 = f10; F.a11 = f11;
  C.a00 = c00; C.a01 = c01; C.a10 = c10; C.a11 = c11;
  const Mat2<double> A = fixed_corotated_affine2(F, C, mu, lam, mass, k);
  out4[0] = (float)A.a00; out4[1] = (float)A.a01; out4[2] = (float)A.a10; out4[3] = (float)A.a11;
}

struct alignas(16) P2GWin2Warp {
  float raw[W2_NPLANES + 1][W2_WINDOW];   // window image of x2 v2 C4 F4 (+ Jp, snow only)
  float4 tile[W2_TILE_NODES];             // {mom_x, mom_y, mass, -} of the window's bounding box of nodes, row pitch = cols
};

template <bool IDX32>
__global__ void __launch_bounds__(W2_WARPS * 32, W2_MIN_CTAS)
p2g_window2_kernel(DevCfg cfg, W2Planes P, long long n, float* __restrict__ grid, ErrRec* err) {
  __shared__ P2GWin2Warp warps[W2_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  P2GWin2Warp& W = warps[warp];
  const int ny = cfg.n[1];
  const int n_windows = (int)((n + W2_WINDOW - 1) / W2_WINDOW);
  const int total_warps = gridDim.x * W2_WARPS;
  const bool snow = cfg.model == 1;

  // uniform scalars, folded once per thread (two_d/p2g.py:57-65, utils.py:53-92)
  const float dxf = (float)cfg.dx;
  const double kd = (cfg.dt * cfg.volume) * (4.0 * cfg.inv_dx * cfg.inv_dx);
  const float kf = (float)kd;
  const float m_u = (float)cfg.mass;
  const double mu_ud = cfg.mu0 * cfg.hardening, lam_ud = cfg.lam0 * cfg.hardening;                   // constant hardening (quirk 8)
  const float mu_u = (float)mu_ud, lam_u = (float)lam_ud;

  // lane l copies the 16-byte chunk (l & 15) of planes 2j + (l >> 4), j = 0 .. 5
  const float* const lane_src = P.base + (long long)(lane >> 4) * P.stride + (lane & 15) * 4;
  auto issue = [&](int win) {
    const float* src = lane_src + (long long)win * W2_WINDOW;
    float* dst = &W.raw[lane >> 4][(lane & 15) * 4];
    const long long step = 2 * P.stride;
#pragma unroll
    for (int j = 0; j < W2_NPLANES / 2; ++j) cp_async16(dst + j * 2 * W2_WINDOW, src + j * step);
    if (snow && lane < W2_WINDOW / 4) cp_async16(&W.raw[W2_NPLANES][lane * 4], P.jp + (long long)win * W2_WINDOW + lane * 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int win = blockIdx.x * W2_WARPS + warp;
  if (win < n_windows) issue(win);
  for (; win < n_windows; win += total_warps) {
    const long long w0 = (long long)win * W2_WINDOW;
    const int cnt = (int)min((long long)W2_WINDOW, n - w0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const int i0 = 2 * lane;
    auto ld2 = [&](int k) -> float2 { return *reinterpret_cast<const float2*>(&W.raw[k][i0]); };
    const float2 X0 = ld2(W2_X), X1 = ld2(W2_X + 1), V0 = ld2(W2_V), V1 = ld2(W2_V + 1);
    const float2 C0 = ld2(W2_C), C1 = ld2(W2_C + 1), C2 = ld2(W2_C + 2), C3 = ld2(W2_C + 3);
    const float2 F0 = ld2(W2_F), F1 = ld2(W2_F + 1), F2 = ld2(W2_F + 2), F3 = ld2(W2_F + 3);
    float2 JP = make_float2(1.0f, 1.0f);
    if (snow) JP = ld2(W2_NPLANES);
    __syncwarp();   // every lane holds its two particles in registers: the image is free for the next window
    if (win + total_warps < n_windows) issue(win + total_warps);

    // ---- phase 1: lane per slot pair ----
    int bx[2], by[2];
    bool ok[2];
    float fx[2], fy[2], mvx[2], mvy[2], a00[2], a01[2], a10[2], a11[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float xs0 = h ? X0.y : X0.x, xs1 = h ? X1.y : X1.x;
      int gx, gy;
      if constexpr (IDX32) {
        base_fx_f32(xs0, (float)cfg.inv_dx, gx, fx[h]);
        base_fx_f32(xs1, (float)cfg.inv_dx, gy, fy[h]);
      } else {
        base_fx(xs0, cfg, gx, fx[h]);
        base_fx(xs1, cfg, gy, fy[h]);
      }
      bx[h] = gx - cfg.origin[0]; by[h] = gy - cfg.origin[1];
      const bool in = i0 + h < cnt;
      // the 2D reference has no bounds check (UB there); we flag and skip instead
      ok[h] = in && xs0 == xs0 && xs1 == xs1 && bx[h] >= 0 && by[h] >= 0 && bx[h] + 2 < cfg.n[0] && by[h] + 2 < cfg.n[1];
      if (in && !ok[h]) atomicAdd(&err->n_oob, 1ULL);
      mvx[h] = mvy[h] = a00[h] = a01[h] = a10[h] = a11[h] = 0.0f;
      if (ok[h]) {
        const float f00 = h ? F0.y : F0.x, f01 = h ? F1.y : F1.x, f10 = h ? F2.y : F2.x, f11 = h ? F3.y : F3.x;
        const float c00 = h ? C0.y : C0.x, c01 = h ? C1.y : C1.x, c10 = h ? C2.y : C2.x, c11 = h ? C3.y : C3.x;
        float mu = mu_u, lam = lam_u;
        double mu_d = mu_ud, lam_d = lam_ud;
        if (snow) {   // snow_hardening, utils.py:48
          const double e = exp(cfg.hardening * (1.0 - (double)(h ? JP.y : JP.x)));
          mu_d = cfg.mu0 * e; lam_d = cfg.lam0 * e;
          mu = (float)mu_d; lam = (float)lam_d;
        }
        Mat2<float> Af;
        bool done = false;
        if (cfg.fp32_stress)
          done = fixed_corotated_affine2_f32(Mat2<float>{f00, f01, f10, f11}, Mat2<float>{c00, c01, c10, c11}, mu, lam, m_u, kf, Af);
        if (!done) {
          float a4[4];
          w2_affine_fp64(f00, f01, f10, f11, c00, c01, c10, c11, mu_d, lam_d, cfg.mass, kd, a4);
          Af.a00 = a4[0]; Af.a01 = a4[1]; Af.a10 = a4[2]; Af.a11 = a4[3];
        }
        a00[h] = Af.a00 * dxf; a01[h] = Af.a01 * dxf; a10[h] = Af.a10 * dxf; a11[h] = Af.a11 * dxf;   // affine * dx
        mvx[h] = m_u * (h ? V0.y : V0.x); mvy[h] = m_u * (h ? V1.y : V1.x);
      }
    }

    // ---- bounding box of the window's base cells ----
    int r_lo = INT_MAX, r_hi = INT_MIN, c_lo = INT_MAX, c_hi = INT_MIN;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (ok[h]) { r_lo = min(r_lo, bx[h]); r_hi = max(r_hi, bx[h]); c_lo = min(c_lo, by[h]); c_hi = max(c_hi, by[h]); }
    r_lo = __reduce_min_sync(0xffffffffu, r_lo); r_hi = __reduce_max_sync(0xffffffffu, r_hi);
    c_lo = __reduce_min_sync(0xffffffffu, c_lo); c_hi = __reduce_max_sync(0xffffffffu, c_hi);
    if (r_hi < r_lo) continue;   // no particle of this window is inside the grid (warp-uniform)
    // row pitch of the tile: the box's width rounded up to 4 mod 8 entries.  In lattice order even and odd lanes work on
    // two adjacent ROWS of the box at the same columns; with a pitch of 16 words mod 32 the eight lanes of a 128-bit
    // shared-memory wavefront then hit disjoint banks (measured before the padding: 181 excess wavefronts per window)
    const int rows = r_hi - r_lo + 3, width = c_hi - c_lo + 3;
    const int cols = ((width + 3) & ~7) + 4;
    const bool use_tile = rows <= W2_TILE_NODES && width <= W2_TILE_NODES && rows * cols <= W2_TILE_NODES;

    if (use_tile) {
      const int nt = rows * cols;
#pragma unroll
      for (int k = 0; k < W2_TILE_NODES / 32; ++k)
        if (lane + 32 * k < nt) W.tile[lane + 32 * k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      __syncwarp();
      int t0[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) t0[h] = ok[h] ? (bx[h] - r_lo) * cols + (by[h] - c_lo) : 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // lanes whose particles share a base cell would add to the same node in the same step: they take turns
        const int key = ok[h] ? bx[h] * ny + by[h] : -1 - (i0 + h);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int rounds = (int)__reduce_max_sync(0xffffffffu, (unsigned)__popc(peers));
        float wx[3], wy[3];
        bspline(fx[h], wx[0], wx[1], wx[2]);
        bspline(fy[h], wy[0], wy[1], wy[2]);
        for (int r = 0; r < rounds; ++r) {
          const bool act = ok[h] && rank == r;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float dpx = (float)i - fx[h];
            const float bxv = mvx[h] + a00[h] * dpx, byv = mvy[h] + a10[h] * dpx;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (act) {
                const float dpy = (float)j - fy[h];
                const float w = wx[i] * wy[j];
                float4 t = W.tile[t0[h] + i * cols + j];
                t.x += w * (bxv + a01[h] * dpy);
                t.y += w * (byv + a11[h] * dpy);
                t.z += w * m_u;
                t.w += w;   // unused sum: makes the write-back ONE 16-byte store
                W.tile[t0[h] + i * cols + j] = t;
              }
              __syncwarp();
            }
          }
        }
      }
      // ---- flush: one vector RED per touched node ----
      const float inv_cols = __frcp_rn((float)cols);
      for (int q = lane; q < nt; q += 32) {
        const float4 t = W.tile[q];
        if (t.x != 0.0f || t.y != 0.0f || t.z != 0.0f) {
          // q / cols without the integer division: exact for q < 160, 4 <= cols <= 160 (checked exhaustively on the host)
          const int r = __float2int_rz(__fmul_rn(__fadd_rn((float)q, 0.5f), inv_cols)), c = q - r * cols;
          red_add4(grid + ((long long)(r_lo + r) * ny + (c_lo + c)) * 4, t.x, t.y, t.z, 0.0f);
        }
      }
      __syncwarp();   // the tile is zeroed again by the next window
    } else {
      // incoherent window: every particle scatters its nine nodes itself
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!ok[h]) continue;
        float wx[3], wy[3];
        bspline(fx[h], wx[0], wx[1], wx[2]);
        bspline(fy[h], wy[0], wy[1], wy[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float dpx = (float)i - fx[h];
          const float bxv = mvx[h] + a00[h] * dpx, byv = mvy[h] + a10[h] * dpx;
          float* row = grid + ((long long)(bx[h] + i) * ny + by[h]) * 4;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float dpy = (float)j - fy[h];
            const float w = wx[i] * wy[j];
            red_add4(row + 4 * j, w * (bxv + a01[h] * dpy), w * (byv + a11[h] * dpy), w * m_u, 0.0f);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// G2P, in place (two_d/g2p.py:17-47): x, F (+ Jp) of the next window prefetched; v and C are outputs only.
// ---------------------------------------------------------------------------------------------------------------
constexpr int W2_G2P_PLANES = 6;   // x0 x1 F00 F01 F10 F11

struct alignas(16) G2PWin2Warp {
  float raw[W2_G2P_PLANES + 1][W2_WINDOW];   // + Jp
};

template <bool IDX32>
__global__ void __launch_bounds__(W2_WARPS * 32, W2_MIN_CTAS)
g2p_window2_kernel(DevCfg cfg, W2Planes P, long long n, const float* __restrict__ grid, ErrRec* err) {
  __shared__ G2PWin2Warp warps[W2_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  G2PWin2Warp& W = warps[warp];
  const int n_windows = (int)((n + W2_WINDOW - 1) / W2_WINDOW);
  const int total_warps = gridDim.x * W2_WARPS;
  const long long st = P.stride;
  const bool has_jp = P.jp != nullptr;

  // lane l copies the 16-byte chunk (l & 15) of the image rows 2j + (l >> 4), j = 0 .. 2: rows 0, 1 = x, rows 2 .. 5 = F
  auto issue = [&](int win) {
    const long long off = (long long)win * W2_WINDOW + (lane & 15) * 4;
#pragma unroll
    for (int j = 0; j < W2_G2P_PLANES / 2; ++j) {
      const int row = 2 * j + (lane >> 4);
      const int plane = row < 2 ? row : row + (W2_F - 2);
      cp_async16(&W.raw[row][(lane & 15) * 4], P.base + (long long)plane * st + off);
    }
    if (has_jp && lane < W2_WINDOW / 4) cp_async16(&W.raw[W2_G2P_PLANES][lane * 4], P.jp + (long long)win * W2_WINDOW + lane * 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int win = blockIdx.x * W2_WARPS + warp;
  if (win < n_windows) issue(win);
  for (; win < n_windows; win += total_warps) {
    const long long w0 = (long long)win * W2_WINDOW;
    const int cnt = (int)min((long long)W2_WINDOW, n - w0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const int i0 = 2 * lane;
    auto ld2 = [&](int k) -> float2 { return *reinterpret_cast<const float2*>(&W.raw[k][i0]); };
    const float2 X0 = ld2(0), X1 = ld2(1), F0 = ld2(2), F1 = ld2(3), F2 = ld2(4), F3 = ld2(5);
    float2 JP = make_float2(1.0f, 1.0f);
    if (has_jp) JP = ld2(W2_G2P_PLANES);
    __syncwarp();
    if (win + total_warps < n_windows) issue(win + total_warps);

    G2POut2<float> o[2];
    bool ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float xs0 = h ? X0.y : X0.x, xs1 = h ? X1.y : X1.x;
      int gx, gy;
      float fx, fy;
      if constexpr (IDX32) {
        base_fx_f32(xs0, (float)cfg.inv_dx, gx, fx);
        base_fx_f32(xs1, (float)cfg.inv_dx, gy, fy);
      } else {
        base_fx(xs0, cfg, gx, fx);
        base_fx(xs1, cfg, gy, fy);
      }
      const int bx = gx - cfg.origin[0], by = gy - cfg.origin[1];
      const bool in = i0 + h < cnt;
      ok[h] = in && xs0 == xs0 && xs1 == xs1 && bx >= 0 && by >= 0 && bx + 2 < cfg.n[0] && by + 2 < cfg.n[1];
      if (in && !ok[h]) atomicAdd(&err->n_oob, 1ULL);   // left unchanged, like g2p_gather2_kernel
      if (ok[h])
        g2p_particle2<float>(cfg, grid, bx, by, fx, fy, xs0, xs1, h ? F0.y : F0.x, h ? F1.y : F1.x, h ? F2.y : F2.x,
                             h ? F3.y : F3.x, h ? JP.y : JP.x, has_jp, o[h]);
    }
    // plane k of (x2 v2 C4 F4) sits k strides above plane 0: one running pointer instead of a product per plane
    if (ok[0] && ok[1]) {
      float* p = P.base + w0 + i0;
      auto st2 = [&](float a, float c) { *reinterpret_cast<float2*>(p) = make_float2(a, c); p += st; };
      st2(o[0].x0, o[1].x0); st2(o[0].x1, o[1].x1);
      st2(o[0].v0, o[1].v0); st2(o[0].v1, o[1].v1);
      st2(o[0].c00, o[1].c00); st2(o[0].c01, o[1].c01); st2(o[0].c10, o[1].c10); st2(o[0].c11, o[1].c11);
      st2(o[0].f00, o[1].f00); st2(o[0].f01, o[1].f01); st2(o[0].f10, o[1].f10); st2(o[0].f11, o[1].f11);
      if (has_jp) *reinterpret_cast<float2*>(P.jp + w0 + i0) = make_float2(o[0].jp, o[1].jp);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!ok[h]) continue;
        float* p = P.base + w0 + i0 + h;
        auto st1 = [&](float a) { *p = a; p += st; };
        st1(o[h].x0); st1(o[h].x1); st1(o[h].v0); st1(o[h].v1);
        st1(o[h].c00); st1(o[h].c01); st1(o[h].c10); st1(o[h].c11);
        st1(o[h].f00); st1(o[h].f01); st1(o[h].f10); st1(o[h].f11);
        if (has_jp) P.jp[w0 + i0 + h] = o[h].jp;
      }
    }
  }
}

// Launch geometry: persistent CTAs, as many as are resident (queried once per device and kernel).
template <typename Kernel>
static int w2_grid_size(Kernel kernel, long long n, int sm_count, int* cache) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cache[dev] == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, W2_WARPS * 32, 0) != cudaSuccess || per_sm < 1) return -1;
    cache[dev] = per_sm;
  }
  const long long windows = (n + W2_WINDOW - 1) / W2_WINDOW;
  const long long want = (windows + W2_WARPS - 1) / W2_WARPS, cap = (long long)sm_count * cache[dev];
  const long long blocks = want < cap ? want : cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static bool p2g_window2_launch(const DevCfg& cfg, const StateView<float>& s, long long n, float* grid, ErrRec* err, int sm_count,
                               cudaStream_t st) {
  static int occ[2][64] = {};
  if (cfg.index_fp32) {
    const int blocks = w2_grid_size(p2g_window2_kernel<true>, n, sm_count, occ[1]);
    if (blocks < 0) return false;
    p2g_window2_kernel<true><<<blocks, W2_WARPS * 32, 0, st>>>(cfg, w2_planes_of(s), n, grid, err);
  } else {
    const int blocks = w2_grid_size(p2g_window2_kernel<false>, n, sm_count, occ[0]);
    if (blocks < 0) return false;
    p2g_window2_kernel<false><<<blocks, W2_WARPS * 32, 0, st>>>(cfg, w2_planes_of(s), n, grid, err);
  }
  return true;
}

static bool g2p_window2_launch(const DevCfg& cfg, const StateView<float>& s, long long n, const float* grid, ErrRec* err,
                               int sm_count, cudaStream_t st) {
  static int occ[2][64] = {};
  if (cfg.index_fp32) {
    const int blocks = w2_grid_size(g2p_window2_kernel<true>, n, sm_count, occ[1]);
    if (blocks < 0) return false;
    g2p_window2_kernel<true><<<blocks, W2_WARPS * 32, 0, st>>>(cfg, w2_planes_of(s), n, grid, err);
  } else {
    const int blocks = w2_grid_size(g2p_window2_kernel<false>, n, sm_count, occ[0]);
    if (blocks < 0) return false;
    g2p_window2_kernel<false><<<blocks, W2_WARPS * 32, 0, st>>>(cfg, w2_planes_of(s), n, grid, err);
  }
  return true;
}

}  // namespace ffmpm

// Phase 2 of the warp-autonomous P2G with TWO PARTICLES PER INSTRUCTION (Blackwell packed fp32).
//
// sm_100 issues fma / mul / add on a pair of fp32 values held in an aligned register pair (PTX
// `fma.rn.f32x2`, SASS FFMA2 / FMUL2 / FADD2; CUDA intrinsics __ffma2_rn, __fmul2_rn, __fadd2_rn).  The
// default P2G is bound by issue slots (76 % busy, FMA pipe 50 %, profiles/r01_final_ncu_summary.txt), and
// its register-accumulation loop is a chain of independent FFMAs per particle: walking a run two
// particles at a time -- slots (2s, 2s+1) in the two halves of every register pair -- issues half the
// floating-point instructions for the same FMA-pipe work.
//
//   * shared-memory payload: eight float4 planes indexed by SLOT PAIR; plane c holds components 2c and
//     2c+1 of both particles as {c.A, c.B, c'.A, c'.B}, so one LDS.128 fills two register pairs.  The slot
//     pair index is padded (s + s/4): runs of ~8 particles start 4 pairs apart, and 5 x 16 bytes per run
//     spreads eight runs over all 32 banks.
//   * a run [r0, r1) covers the slot pairs r0/2 .. (r1-1)/2; a partner that belongs to the neighbouring
//     run (odd r0 / odd r1) is switched off through its x-weight (payloads are always finite: slots past
//     the end of the last window are parked as zeros).
//   * the two halves of the 36 accumulator pairs are summed once per run, before the nine vector REDs.
//
// Same sums as p2g_runs_phase2 up to the order of the additions.  Arithmetic per node (three_d/p2g.py:67-80)
// is spelled with explicit fused multiply-adds because the packed intrinsics are not contracted by the
// compiler.  The host instantiation (two plain floats) is what tests/test_kernel_math_host.py checks.
#pragma once
#include "mpm_p2g_runs.cuh"

namespace ffmpm {

struct F2 {
  float2 v;
};
FFMPM_HD F2 f2(float a, float b) { F2 r; r.v.x = a; r.v.y = b; return r; }
FFMPM_HD F2 f2(float a) { return f2(a, a); }
FFMPM_HD F2 f2_fma(F2 a, F2 b, F2 c) {
#ifdef __CUDA_ARCH__
  F2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r;
#else
  return f2(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y));
#endif
}
FFMPM_HD F2 f2_mul(F2 a, F2 b) {
#ifdef __CUDA_ARCH__
  F2 r; r.v = __fmul2_rn(a.v, b.v); return r;
#else
  return f2(a.v.x * b.v.x, a.v.y * b.v.y);
#endif
}
FFMPM_HD F2 f2_add(F2 a, F2 b) {
#ifdef __CUDA_ARCH__
  F2 r; r.v = __fadd2_rn(a.v, b.v); return r;
#else
  return f2(a.v.x + b.v.x, a.v.y + b.v.y);
#endif
}

constexpr int P2G_NPAIR = P2G_WINDOW / 2;
constexpr int P2G_PAIR_PADDED = P2G_NPAIR + P2G_NPAIR / 4;
constexpr int P2G_PAIR_PLANES = 8;   // 16 payload components, two per plane

FFMPM_HD int p2g_pair_pad(int s) { return s + (s >> 2); }

// Component order of the payload (= the four float4 of P2GWarpSlab::pay, flattened).
enum { PP_MVX = 0, PP_MVY, PP_MVZ, PP_M, PP_A00, PP_A01, PP_A02, PP_FX, PP_A10, PP_A11, PP_A12, PP_FY, PP_A20, PP_A21, PP_A22, PP_FZ };

struct P2GPairSlab {
  float4 pay[P2G_PAIR_PLANES][P2G_PAIR_PADDED];
  int node0[P2G_WINDOW];
  int run_start[P2G_WINDOW + 1];
};

// Address of component c of slot q inside the pair-major payload.
FFMPM_HD float* p2g_pair_slot(float4 (*pay)[P2G_PAIR_PADDED], int q, int c) {
  return reinterpret_cast<float*>(&pay[c >> 1][p2g_pair_pad(q >> 1)]) + ((c & 1) * 2 + (q & 1));
}

// Phase 1 tail for the pair-major layout (the scalar p2g_park's twin).
FFMPM_HD int p2g_park_pair(float4 (*pay)[P2G_PAIR_PADDED], int* node0, P2GParticle3<float>& q, int idx, float dx, int ny, int nz) {
  if (!q.ok) {
    q.mvx = q.mvy = q.mvz = q.m = 0.0f;
    q.a00 = q.a01 = q.a02 = q.a10 = q.a11 = q.a12 = q.a20 = q.a21 = q.a22 = 0.0f;
    q.fx = q.fy = q.fz = 0.5f;
  }
  const float val[16] = {q.mvx, q.mvy, q.mvz, q.m, q.a00 * dx, q.a01 * dx, q.a02 * dx, q.fx,
                         q.a10 * dx, q.a11 * dx, q.a12 * dx, q.fy, q.a20 * dx, q.a21 * dx, q.a22 * dx, q.fz};
#pragma unroll
  for (int c = 0; c < 16; ++c) *p2g_pair_slot(pay, idx, c) = val[c];
  const int node = q.ok ? (q.bx * ny + q.by) * nz + q.bz : -1;
  node0[idx] = node;
  return node;
}

// A slot nobody owns (past the end of the last window): finite zeros, so that a masked partner stays 0.
FFMPM_HD void p2g_park_pair_zero(float4 (*pay)[P2G_PAIR_PADDED], int idx) {
#pragma unroll
  for (int c = 0; c < 16; ++c) *p2g_pair_slot(pay, idx, c) = c == PP_FX || c == PP_FY || c == PP_FZ ? 0.5f : 0.0f;
}

// Packed quadratic B-spline pieces (three_d/p2g.py:55) of both particles, and the node offsets k - f.
FFMPM_HD void f2_bspline(F2 f, F2 (&w)[3], F2 (&d)[3]) {
  const F2 neg1 = f2(-1.0f);
  const F2 a = f2_fma(f, neg1, f2(1.5f));        // 1.5 - f
  const F2 nb = f2_fma(f, neg1, f2(1.0f));       // 1 - f  (= -(f - 1))
  const F2 c = f2_add(f, f2(-0.5f));             // f - 0.5
  w[0] = f2_mul(f2_mul(a, f2(0.5f)), a);
  w[1] = f2_fma(nb, f2_mul(nb, neg1), f2(0.75f));   // 0.75 - (f - 1)^2
  w[2] = f2_mul(f2_mul(c, f2(0.5f)), c);
  d[0] = f2_mul(f, neg1);
  d[1] = nb;
  d[2] = f2_fma(f, neg1, f2(2.0f));
}

// The nine nodes of x-slab `li` of one run [r0, r1) of a parked window, two particles per step.
// out[j*3+k] = {mom_x, mom_y, mom_z, mass} to add at node (li, j, k) of the run's base cell.
FFMPM_HD void p2g_pair_accumulate(const float4 (*pay)[P2G_PAIR_PADDED], int r0, int r1, int li, float (&out)[9][4]) {
  const float ci = (float)li;
  // B-spline piece of this slab along x: w = s * (f - c)^2 + o
  const F2 sx = f2(li == 1 ? -1.0f : 0.5f), ncx = f2(-(1.5f - 0.5f * ci)), ox = f2(li == 1 ? 0.75f : 0.0f);
  const F2 neg1 = f2(-1.0f), ci2 = f2(ci);
  F2 ax[9], ay[9], az[9], am[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) ax[e] = ay[e] = az[e] = am[e] = f2(0.0f);
  for (int s = r0 >> 1; 2 * s < r1; ++s) {
    const int ph = p2g_pair_pad(s);
    F2 p[16];
#pragma unroll
    for (int c = 0; c < P2G_PAIR_PLANES; ++c) {
      const float4 t = pay[c][ph];
      p[2 * c] = f2(t.x, t.y);
      p[2 * c + 1] = f2(t.z, t.w);
    }
    const F2 mask = f2(2 * s >= r0 ? 1.0f : 0.0f, 2 * s + 1 < r1 ? 1.0f : 0.0f);
    F2 wy[3], wz[3], dy[3], dz[3];
    f2_bspline(p[PP_FY], wy, dy);
    f2_bspline(p[PP_FZ], wz, dz);
    const F2 tx = f2_add(p[PP_FX], ncx);
    const F2 wxi = f2_mul(f2_fma(f2_mul(sx, tx), tx, ox), mask);
    const F2 dpx = f2_fma(p[PP_FX], neg1, ci2);
    const F2 bx = f2_fma(p[PP_A00], dpx, p[PP_MVX]), by = f2_fma(p[PP_A10], dpx, p[PP_MVY]), bz = f2_fma(p[PP_A20], dpx, p[PP_MVZ]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const F2 wij = f2_mul(wxi, wy[j]);
      const F2 cxj = f2_fma(p[PP_A01], dy[j], bx), cyj = f2_fma(p[PP_A11], dy[j], by), czj = f2_fma(p[PP_A21], dy[j], bz);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const F2 w = f2_mul(wij, wz[k]);
        ax[j * 3 + k] = f2_fma(w, f2_fma(p[PP_A02], dz[k], cxj), ax[j * 3 + k]);
        ay[j * 3 + k] = f2_fma(w, f2_fma(p[PP_A12], dz[k], cyj), ay[j * 3 + k]);
        az[j * 3 + k] = f2_fma(w, f2_fma(p[PP_A22], dz[k], czj), az[j * 3 + k]);
        am[j * 3 + k] = f2_fma(w, p[PP_M], am[j * 3 + k]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    out[e][0] = ax[e].v.x + ax[e].v.y;
    out[e][1] = ay[e].v.x + ay[e].v.y;
    out[e][2] = az[e].v.x + az[e].v.y;
    out[e][3] = am[e].v.x + am[e].v.y;
  }
}

// Runs + phase 2 over a window parked in the pair-major layout (twin of p2g_runs_phase2).
__device__ __forceinline__ void p2g_runs_phase2_pair(P2GPairSlab& S, const int (&node)[2], int cnt, int lane, int ny, int nz,
                                                     float* __restrict__ grid) {
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
  const int n_items = n_runs * 3;
  for (int item = lane; item < n_items; item += 32) {
    const int r = item / 3, li = item - r * 3;
    const int r0 = S.run_start[r], r1 = S.run_start[r + 1];
    if (S.node0[r0] < 0) continue;   // a run of out-of-grid particles
    float out[9][4];
    p2g_pair_accumulate(S.pay, r0, r1, li, out);
    float* g = grid + ((long long)S.node0[r0] + (long long)li * ny * nz) * 4;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        red_add4(g + ((long long)j * nz + k) * 4, out[j * 3 + k][0], out[j * 3 + k][1], out[j * 3 + k][2], out[j * 3 + k][3]);
  }
}

}  // namespace ffmpm

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


// ----------------------------------------------------------------------------
// The fp32 perturbation-form stress of mpm_math.cuh (fixed_corotated_affine3_f32) for TWO particles at
// once: every quantity is a register pair, every operation one packed instruction.  Same formulae, same
// series tiers; the degree is picked from the more strained of the two, and the function declines (the
// caller then evaluates both particles one by one, fp64 path included) when either is beyond the series.
// ----------------------------------------------------------------------------
struct Sym3x2 {
  F2 xx, xy, xz, yy, yz, zz;
};

FFMPM_HD Sym3x2 sym3_mul2(const Sym3x2& a, const Sym3x2& b) {
  Sym3x2 r;
  r.xx = f2_fma(a.xz, b.xz, f2_fma(a.xy, b.xy, f2_mul(a.xx, b.xx)));
  r.xy = f2_fma(a.xz, b.yz, f2_fma(a.xy, b.yy, f2_mul(a.xx, b.xy)));
  r.xz = f2_fma(a.xz, b.zz, f2_fma(a.xy, b.yz, f2_mul(a.xx, b.xz)));
  r.yy = f2_fma(a.yz, b.yz, f2_fma(a.yy, b.yy, f2_mul(a.xy, b.xy)));
  r.yz = f2_fma(a.yz, b.zz, f2_fma(a.yy, b.yz, f2_mul(a.xy, b.xz)));
  r.zz = f2_fma(a.zz, b.zz, f2_fma(a.yz, b.yz, f2_mul(a.xz, b.xz)));
  return r;
}

struct Mat3x2 {
  F2 a00, a01, a02, a10, a11, a12, a20, a21, a22;
};

// `load_C()` fetches the two APIC matrices only when the series is done: they are needed for the last nine
// FFMA2s, and holding 18 more registers through the series is what makes the packed phase 1 spill.
//
// `economised` (cfg.fp32_stress == 2, FFMPM_FP32_STRESS=2): q = M / G is not the truncated Taylor series but its
// Chebyshev interpolant on the tier's interval [-r_t, r_t] (scripts/series_economized.py prints the table below):
// the same 1e-7 bar with q of degree 2 / 3 / 4 / 5 / 6 for ||G||_F < 0.007 / 0.025 / 0.069 / 0.109 / 0.15 instead of
// 2 / 4 / 7 for 0.005 / 0.04 / 0.15 -- at the headline workload's strain (warp maximum ~ 0.08) five matrix
// products instead of seven.  Evaluated in fp32 both forms sit at ~2e-7 (rounding, not truncation).
//
// `form` 3 (FFMPM_FP32_STRESS=3), the LEFT form: with F = V R (V = (F F^T)^(1/2)), (F - R) F^T = F F^T - V, i.e. the
// stress term is the matrix function  h(G_B) = (I + G_B) - (I + G_B)^(1/2) = G_B p(G_B)  of the left Cauchy-Green
// strain G_B = F F^T - I = E + E^T + E E^T alone: no product with F afterwards (45 multiply-adds saved), and
// p(x) = (1 + x - sqrt(1 + x)) / x converges faster than q: degree 2 / 3 / 4 / 5 below 0.0136 / 0.0436 / 0.112 / 0.15
// -- four symmetric matrix products in all at the headline strain, against seven plus the two products with F.
template <typename LoadC>
FFMPM_HD bool fixed_corotated_affine3_f32x2(const Mat3x2& F, LoadC load_C, F2 mu, F2 lam, F2 mass, float dt_vol_dinv,
                                            Mat3x2& A, int form = 1) {
  const bool economised = form == 2, left = form == 3;
  const F2 neg1 = f2(-1.0f), two = f2(2.0f);
  Mat3x2 E = F;
  E.a00 = f2_add(F.a00, neg1); E.a11 = f2_add(F.a11, neg1); E.a22 = f2_add(F.a22, neg1);
  Sym3x2 G;
  if (!left) {     // G = F^T F - I = E + E^T + E^T E
    G.xx = f2_fma(two, E.a00, f2_fma(E.a20, E.a20, f2_fma(E.a10, E.a10, f2_mul(E.a00, E.a00))));
    G.yy = f2_fma(two, E.a11, f2_fma(E.a21, E.a21, f2_fma(E.a11, E.a11, f2_mul(E.a01, E.a01))));
    G.zz = f2_fma(two, E.a22, f2_fma(E.a22, E.a22, f2_fma(E.a12, E.a12, f2_mul(E.a02, E.a02))));
    G.xy = f2_add(f2_add(E.a01, E.a10), f2_fma(E.a20, E.a21, f2_fma(E.a10, E.a11, f2_mul(E.a00, E.a01))));
    G.xz = f2_add(f2_add(E.a02, E.a20), f2_fma(E.a20, E.a22, f2_fma(E.a10, E.a12, f2_mul(E.a00, E.a02))));
    G.yz = f2_add(f2_add(E.a12, E.a21), f2_fma(E.a21, E.a22, f2_fma(E.a11, E.a12, f2_mul(E.a01, E.a02))));
  } else {         // G_B = F F^T - I = E + E^T + E E^T
    G.xx = f2_fma(two, E.a00, f2_fma(E.a02, E.a02, f2_fma(E.a01, E.a01, f2_mul(E.a00, E.a00))));
    G.yy = f2_fma(two, E.a11, f2_fma(E.a12, E.a12, f2_fma(E.a11, E.a11, f2_mul(E.a10, E.a10))));
    G.zz = f2_fma(two, E.a22, f2_fma(E.a22, E.a22, f2_fma(E.a21, E.a21, f2_mul(E.a20, E.a20))));
    G.xy = f2_add(f2_add(E.a01, E.a10), f2_fma(E.a02, E.a12, f2_fma(E.a01, E.a11, f2_mul(E.a00, E.a10))));
    G.xz = f2_add(f2_add(E.a02, E.a20), f2_fma(E.a02, E.a22, f2_fma(E.a01, E.a21, f2_mul(E.a00, E.a20))));
    G.yz = f2_add(f2_add(E.a12, E.a21), f2_fma(E.a12, E.a22, f2_fma(E.a11, E.a21, f2_mul(E.a10, E.a20))));
  }
  const F2 d2 = f2_fma(G.zz, G.zz, f2_fma(G.yy, G.yy, f2_mul(G.xx, G.xx)));
  const F2 o2 = f2_fma(G.yz, G.yz, f2_fma(G.xz, G.xz, f2_mul(G.xy, G.xy)));
  const F2 r2p = f2_fma(two, o2, d2);
  const float r2 = fmaxf(r2p.v.x, r2p.v.y);
  if (!(r2 < kPerturbationMaxR * kPerturbationMaxR)) return false;      // also catches NaN in either half
  // c[0 .. top]: the coefficients the Horner steps add, ca / cb: the two highest (q starts as ca G + cb)
  float c[6] = {0.5f, -0.375f, 0.3125f, -0.2734375f, 0.24609375f, -0.2255859375f};
  float ca = -0.196380615234375f, cb = 0.20947265625f;
  int top = 5;
  if (left) {                               // p(x) = (1 + x - sqrt(1 + x)) / x, economised per tier
    if (r2 < 0.0136f * 0.0136f) {
      c[0] = 0.5f; cb = 0.12500542402267456f; ca = -0.06250379234552383f; top = 0;
    } else if (r2 < 0.0436f * 0.0436f) {
      c[0] = 0.5f; c[1] = 0.1249999925494194f; cb = -0.06255202740430832f; ca = 0.03910152614116669f; top = 1;
    } else if (r2 < 0.112f * 0.112f) {
      c[0] = 0.5f; c[1] = 0.12499897927045822f; c[2] = -0.06249919906258583f; cb = 0.03938665986061096f; ca = -0.02759857103228569f;
      top = 2;
    } else {
      c[0] = 0.5f; c[1] = 0.125f; c[2] = -0.062495309859514236f; c[3] = 0.039058685302734375f; cb = -0.027897052466869354f;
      ca = 0.02095773071050644f; top = 3;
    }
  } else if (!economised) {
    if (r2 < 0.005f * 0.005f) { ca = c[2]; cb = c[1]; top = 0; }
    else if (r2 < 0.04f * 0.04f) { ca = c[4]; cb = c[3]; top = 2; }
  } else if (r2 < 0.007f * 0.007f) {        // kEcon tier 0: q of degree 2
    c[0] = 0.5f; cb = -0.37501004338264465f; ca = 0.31250903010368347f; top = 0;
  } else if (r2 < 0.025f * 0.025f) {        // tier 1: degree 3
    c[0] = 0.5f; c[1] = -0.375f; cb = 0.31265386939048767f; ca = -0.27357855439186096f; top = 1;
  } else if (r2 < 0.069f * 0.069f) {        // tier 2: degree 4
    c[0] = 0.5f; c[1] = -0.37499839067459106f; c[2] = 0.3124985098838806f; cb = -0.2747856080532074f; ca = 0.24734565615653992f;
    top = 2;
  } else if (r2 < 0.109f * 0.109f) {        // tier 3: degree 5
    c[0] = 0.5f; c[1] = -0.375f; c[2] = 0.3124831020832062f; c[3] = -0.27342167496681213f; cb = 0.24987153708934784f;
    ca = -0.2291281670331955f; top = 3;
  } else {                                  // tier 4: degree 6, up to kPerturbationMaxR
    c[0] = 0.5f; c[1] = -0.3750002682209015f; c[2] = 0.3125002384185791f; c[3] = -0.2733475863933563f;
    c[4] = 0.24600879848003387f; cb = -0.2335180640220642f; ca = 0.2169661521911621f; top = 4;
  }
  const F2 ca2 = f2(ca), cb2 = f2(cb);
  Sym3x2 q;
  q.xx = f2_fma(ca2, G.xx, cb2); q.yy = f2_fma(ca2, G.yy, cb2); q.zz = f2_fma(ca2, G.zz, cb2);
  q.xy = f2_mul(ca2, G.xy); q.xz = f2_mul(ca2, G.xz); q.yz = f2_mul(ca2, G.yz);
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    if (i <= top) {
      q = sym3_mul2(G, q);
      const F2 ci = f2(c[i]);
      q.xx = f2_add(q.xx, ci); q.yy = f2_add(q.yy, ci); q.zz = f2_add(q.zz, ci);
    }
  }
  const Sym3x2 M = sym3_mul2(G, q);
  F2 x00, x01, x02, x11, x12, x22;
  if (left) {      // M is h(G_B) = (F - R) F^T itself
    x00 = M.xx; x01 = M.xy; x02 = M.xz; x11 = M.yy; x12 = M.yz; x22 = M.zz;
  } else {         // W = F M,  X = W F^T (symmetric)
    const F2 w00 = f2_fma(F.a02, M.xz, f2_fma(F.a01, M.xy, f2_mul(F.a00, M.xx)));
    const F2 w01 = f2_fma(F.a02, M.yz, f2_fma(F.a01, M.yy, f2_mul(F.a00, M.xy)));
    const F2 w02 = f2_fma(F.a02, M.zz, f2_fma(F.a01, M.yz, f2_mul(F.a00, M.xz)));
    const F2 w10 = f2_fma(F.a12, M.xz, f2_fma(F.a11, M.xy, f2_mul(F.a10, M.xx)));
    const F2 w11 = f2_fma(F.a12, M.yz, f2_fma(F.a11, M.yy, f2_mul(F.a10, M.xy)));
    const F2 w12 = f2_fma(F.a12, M.zz, f2_fma(F.a11, M.yz, f2_mul(F.a10, M.xz)));
    const F2 w20 = f2_fma(F.a22, M.xz, f2_fma(F.a21, M.xy, f2_mul(F.a20, M.xx)));
    const F2 w21 = f2_fma(F.a22, M.yz, f2_fma(F.a21, M.yy, f2_mul(F.a20, M.xy)));
    const F2 w22 = f2_fma(F.a22, M.zz, f2_fma(F.a21, M.yz, f2_mul(F.a20, M.xz)));
    x00 = f2_fma(w02, F.a02, f2_fma(w01, F.a01, f2_mul(w00, F.a00)));
    x01 = f2_fma(w02, F.a12, f2_fma(w01, F.a11, f2_mul(w00, F.a10)));
    x02 = f2_fma(w02, F.a22, f2_fma(w01, F.a21, f2_mul(w00, F.a20)));
    x11 = f2_fma(w12, F.a12, f2_fma(w11, F.a11, f2_mul(w10, F.a10)));
    x12 = f2_fma(w12, F.a22, f2_fma(w11, F.a21, f2_mul(w10, F.a20)));
    x22 = f2_fma(w22, F.a22, f2_fma(w21, F.a21, f2_mul(w20, F.a20)));
  }
  // J - 1 = tr E + principal 2x2 minors + det E, without cancellation
  const F2 trE = f2_add(f2_add(E.a00, E.a11), E.a22);
  const F2 m01 = f2_sub(f2_mul(E.a00, E.a11), f2_mul(E.a01, E.a10));
  const F2 m02 = f2_sub(f2_mul(E.a00, E.a22), f2_mul(E.a02, E.a20));
  const F2 m12 = f2_sub(f2_mul(E.a11, E.a22), f2_mul(E.a12, E.a21));
  const F2 c2 = f2_add(f2_add(m01, m02), m12);
  const F2 k0 = f2_sub(f2_mul(E.a11, E.a22), f2_mul(E.a12, E.a21));
  const F2 k1 = f2_sub(f2_mul(E.a10, E.a22), f2_mul(E.a12, E.a20));
  const F2 k2m = f2_sub(f2_mul(E.a10, E.a21), f2_mul(E.a11, E.a20));
  const F2 dE = f2_fma(E.a02, k2m, f2_sub(f2_mul(E.a00, k0), f2_mul(E.a01, k1)));
  const F2 jm1 = f2_add(f2_add(trE, c2), dE);
  const F2 l = f2_mul(f2_mul(lam, jm1), f2_add(jm1, f2(1.0f)));       // lam (J-1) J, broadcast onto ALL entries (quirk 2)
  const F2 nk = f2(-dt_vol_dinv);
  const F2 k2 = f2_mul(f2_mul(nk, two), mu), kl = f2_mul(nk, l);
#ifdef __CUDA_ARCH__
  asm volatile("" ::: "memory");   // keep the loads of C below the series (scheduling barrier only)
#endif
  const Mat3x2 C = load_C();
  A.a00 = f2_fma(mass, C.a00, f2_fma(k2, x00, kl)); A.a01 = f2_fma(mass, C.a01, f2_fma(k2, x01, kl));
  A.a02 = f2_fma(mass, C.a02, f2_fma(k2, x02, kl)); A.a10 = f2_fma(mass, C.a10, f2_fma(k2, x01, kl));
  A.a11 = f2_fma(mass, C.a11, f2_fma(k2, x11, kl)); A.a12 = f2_fma(mass, C.a12, f2_fma(k2, x12, kl));
  A.a20 = f2_fma(mass, C.a20, f2_fma(k2, x02, kl)); A.a21 = f2_fma(mass, C.a21, f2_fma(k2, x12, kl));
  A.a22 = f2_fma(mass, C.a22, f2_fma(k2, x22, kl));
  return true;
}

// Phase 1 for the two particles a lane owns in a window (slots `lane` and `lane + 32`): cell index, weights
// offset and material one by one (integer / fp64 work), the stress of both in packed fp32.  Falls back to
// the one-particle routine (p2g_prepare3_from, fp64 stress included) unless both are inside the grid and
// inside the series' strain range.  `live_b` false: the window holds no second particle for this lane.
// `sink` receives the results as they become final, so that a kernel can park them in shared memory at once
// instead of carrying two whole particles in registers through the stress evaluation (what made the first
// packed phase 1 spill):
//   sink.head(h, q)   particle h (0 / 1): base cell, weights offset, mass, mass*v are final (q.ok is true);
//   sink.affine(A)    the nine affine entries of both particles (pairs: .x = particle 0, .y = particle 1);
//   sink.full(h, q)   particle h evaluated by the one-particle routine (everything final, q.ok may be false).
// FORM: 0 = the stress form is cfg.fp32_stress at run time (1 Taylor, 2 economised, 3 left); 3 = the left form at
// compile time (F is dead once F F^T - I is formed, and the code of the other forms is not generated at all).
template <int FORM = 0, typename GetA, typename GetB, typename Sink>
FFMPM_HD void p2g_prepare3_pair_sink(const DevCfg& cfg, GetA ga, GetB gb, bool has_mat, bool live_a, bool live_b, Sink& sink) {
  bool packed = live_a && live_b && cfg.fp32_stress && cfg.model == 0;
  if (packed) {
    float mass_a, mu_a, lam_a, mass_b, mu_b, lam_b;
    P2GParticle3<float> qa, qb;
    auto index_part = [&](auto get, P2GParticle3<float>& q, float& mass_f, float& mu_f, float& lam_f) {
      const float x0 = get(P2G_X), x1 = get(P2G_X + 1), x2 = get(P2G_X + 2);
      int gx, gy, gz;
      base_fx(x0, cfg, gx, q.fx);
      base_fx(x1, cfg, gy, q.fy);
      base_fx(x2, cfg, gz, q.fz);
      q.bx = gx - cfg.origin[0]; q.by = gy - cfg.origin[1]; q.bz = gz - cfg.origin[2];
      q.ok = !(isnan((double)x0) || isnan((double)x1) || isnan((double)x2)) &&
             q.bx >= 0 && q.by >= 0 && q.bz >= 0 && q.bx + 2 < cfg.n[0] && q.by + 2 < cfg.n[1] && q.bz + 2 < cfg.n[2];
      const double mass = has_mat ? (double)get(P2G_MASS) : cfg.mass;
      const double mu = (has_mat ? (double)get(P2G_MU) : cfg.mu0) * cfg.hardening;     // constant hardening: a multiplier (quirk 8)
      const double lam = (has_mat ? (double)get(P2G_LAM) : cfg.lam0) * cfg.hardening;
      mass_f = (float)mass; mu_f = (float)mu; lam_f = (float)lam;
      q.m = (float)mass;
      if (has_mat) {
        q.mvx = q.m * get(P2G_V); q.mvy = q.m * get(P2G_V + 1); q.mvz = q.m * get(P2G_V + 2);
      } else {
        q.mvx = (float)(mass * (double)get(P2G_V)); q.mvy = (float)(mass * (double)get(P2G_V + 1));
        q.mvz = (float)(mass * (double)get(P2G_V + 2));
      }
    };
    index_part(ga, qa, mass_a, mu_a, lam_a);
    index_part(gb, qb, mass_b, mu_b, lam_b);
    packed = qa.ok && qb.ok;
    if (packed) {
      Mat3x2 F, A;
      F.a00 = f2(ga(P2G_F + 0), gb(P2G_F + 0)); F.a01 = f2(ga(P2G_F + 1), gb(P2G_F + 1)); F.a02 = f2(ga(P2G_F + 2), gb(P2G_F + 2));
      F.a10 = f2(ga(P2G_F + 3), gb(P2G_F + 3)); F.a11 = f2(ga(P2G_F + 4), gb(P2G_F + 4)); F.a12 = f2(ga(P2G_F + 5), gb(P2G_F + 5));
      F.a20 = f2(ga(P2G_F + 6), gb(P2G_F + 6)); F.a21 = f2(ga(P2G_F + 7), gb(P2G_F + 7)); F.a22 = f2(ga(P2G_F + 8), gb(P2G_F + 8));
      auto load_C = [&]() {
        Mat3x2 C;
        C.a00 = f2(ga(P2G_C + 0), gb(P2G_C + 0)); C.a01 = f2(ga(P2G_C + 1), gb(P2G_C + 1)); C.a02 = f2(ga(P2G_C + 2), gb(P2G_C + 2));
        C.a10 = f2(ga(P2G_C + 3), gb(P2G_C + 3)); C.a11 = f2(ga(P2G_C + 4), gb(P2G_C + 4)); C.a12 = f2(ga(P2G_C + 5), gb(P2G_C + 5));
        C.a20 = f2(ga(P2G_C + 6), gb(P2G_C + 6)); C.a21 = f2(ga(P2G_C + 7), gb(P2G_C + 7)); C.a22 = f2(ga(P2G_C + 8), gb(P2G_C + 8));
        return C;
      };
      const double k = (cfg.dt * cfg.volume) * (4.0 * cfg.inv_dx * cfg.inv_dx);
      // whether the series accepts the pair is known from F alone, but only at the end of its first stage; the
      // heads are parked first (a declined pair parks everything again through the one-particle routine)
      sink.head(0, qa);
      sink.head(1, qb);
      packed = fixed_corotated_affine3_f32x2(F, load_C, f2(mu_a, mu_b), f2(lam_a, lam_b), f2(mass_a, mass_b), (float)k, A,
                                             FORM != 0 ? FORM : cfg.fp32_stress);
      if (packed) {
        sink.affine(A);
        return;
      }
    }
  }
  if (live_a) { P2GParticle3<float> q = p2g_prepare3_from<float>(cfg, ga, has_mat, 1.0); sink.full(0, q); }
  if (live_b) { P2GParticle3<float> q = p2g_prepare3_from<float>(cfg, gb, has_mat, 1.0); sink.full(1, q); }
}

// The same with the two particles returned whole (host tests, and callers that keep them in registers).
struct P2GPairRegisters {
  P2GParticle3<float>* q[2];
  FFMPM_HD void head(int h, const P2GParticle3<float>& v) { *q[h] = v; }
  FFMPM_HD void affine(const Mat3x2& A) {
    P2GParticle3<float>& a = *q[0];
    P2GParticle3<float>& b = *q[1];
    a.a00 = A.a00.v.x; a.a01 = A.a01.v.x; a.a02 = A.a02.v.x; a.a10 = A.a10.v.x; a.a11 = A.a11.v.x; a.a12 = A.a12.v.x;
    a.a20 = A.a20.v.x; a.a21 = A.a21.v.x; a.a22 = A.a22.v.x;
    b.a00 = A.a00.v.y; b.a01 = A.a01.v.y; b.a02 = A.a02.v.y; b.a10 = A.a10.v.y; b.a11 = A.a11.v.y; b.a12 = A.a12.v.y;
    b.a20 = A.a20.v.y; b.a21 = A.a21.v.y; b.a22 = A.a22.v.y;
  }
  FFMPM_HD void full(int h, const P2GParticle3<float>& v) { *q[h] = v; }
};

template <typename GetA, typename GetB>
FFMPM_HD void p2g_prepare3_pair(const DevCfg& cfg, GetA ga, GetB gb, bool has_mat, bool live_a, bool live_b,
                                P2GParticle3<float>& qa, P2GParticle3<float>& qb) {
  P2GPairRegisters sink{{&qa, &qb}};
  p2g_prepare3_pair_sink<0>(cfg, ga, gb, has_mat, live_a, live_b, sink);
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

// Sink of p2g_prepare3_pair_sink that parks straight into the pair-major payload: slots `slot0` (particle 0)
// and `slot0 + 32` (particle 1) of the warp's window.
struct P2GPairParker {
  float4 (*pay)[P2G_PAIR_PADDED];
  int* node0;
  int slot0, ny, nz;
  float dx;
  int node[2];
  FFMPM_HD void head(int h, const P2GParticle3<float>& q) {
    const int idx = slot0 + 32 * h;
    *p2g_pair_slot(pay, idx, PP_MVX) = q.mvx; *p2g_pair_slot(pay, idx, PP_MVY) = q.mvy; *p2g_pair_slot(pay, idx, PP_MVZ) = q.mvz;
    *p2g_pair_slot(pay, idx, PP_M) = q.m;
    *p2g_pair_slot(pay, idx, PP_FX) = q.fx; *p2g_pair_slot(pay, idx, PP_FY) = q.fy; *p2g_pair_slot(pay, idx, PP_FZ) = q.fz;
    node[h] = (q.bx * ny + q.by) * nz + q.bz;
    node0[idx] = node[h];
  }
  FFMPM_HD void affine(const Mat3x2& A) {
    const F2 d = f2(dx);
    const F2 v[9] = {f2_mul(A.a00, d), f2_mul(A.a01, d), f2_mul(A.a02, d), f2_mul(A.a10, d), f2_mul(A.a11, d),
                     f2_mul(A.a12, d), f2_mul(A.a20, d), f2_mul(A.a21, d), f2_mul(A.a22, d)};
    const int comp[9] = {PP_A00, PP_A01, PP_A02, PP_A10, PP_A11, PP_A12, PP_A20, PP_A21, PP_A22};
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      *p2g_pair_slot(pay, slot0, comp[e]) = v[e].v.x;
      *p2g_pair_slot(pay, slot0 + 32, comp[e]) = v[e].v.y;
    }
  }
  FFMPM_HD void full(int h, const P2GParticle3<float>& q) {
    P2GParticle3<float> c = q;
    node[h] = p2g_park_pair(pay, node0, c, slot0 + 32 * h, dx, ny, nz);
  }
};

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
// `run_cap` (0 or a power of two): additionally cut a run at every run_cap-th slot of the window.  Where many
// particles share a cell (a column settling on the floor) a window holds one or two runs and most lanes
// of phase 2 idle; capped runs keep >= 64 / run_cap * 3 (run, slab) items per window at the price of one
// more set of REDs per cut.
__device__ __forceinline__ void p2g_runs_phase2_pair(P2GPairSlab& S, const int (&node)[2], int cnt, int lane, int ny, int nz,
                                                     float* __restrict__ grid, int run_cap) {
  unsigned heads[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int idx = h * 32 + lane;
    const int prev = idx > 0 ? S.node0[idx - 1] : -2;
    heads[h] = __ballot_sync(0xffffffffu, idx < cnt && (node[h] != prev || (run_cap > 0 && (idx & (run_cap - 1)) == 0)));
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

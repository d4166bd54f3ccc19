// Small dense linear algebra and constitutive math for the MPM kernels.
// Everything is register-resident (no local arrays with dynamic indexing).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// The constitutive math is plain arithmetic: it also compiles for the host, where tests/native/
// runs these very functions against LAPACK on CPU (tests/test_kernel_math_host.py).
#define FFMPM_HD __host__ __device__ __forceinline__

namespace ffmpm {

template <typename T>
struct Mat3 {
  T a00, a01, a02, a10, a11, a12, a20, a21, a22;
};

template <typename T>
struct Mat2 {
  T a00, a01, a10, a11;
};

template <typename T>
FFMPM_HD T det3(const Mat3<T>& m) {
  return m.a00 * (m.a11 * m.a22 - m.a12 * m.a21) - m.a01 * (m.a10 * m.a22 - m.a12 * m.a20) +
         m.a02 * (m.a10 * m.a21 - m.a11 * m.a20);
}

// Cofactor matrix cof(X) (so that X^{-T} = cof(X) / det X).
FFMPM_HD Mat3<double> cofactor3(const Mat3<double>& m) {
  Mat3<double> c;
  c.a00 = m.a11 * m.a22 - m.a12 * m.a21;
  c.a01 = m.a12 * m.a20 - m.a10 * m.a22;
  c.a02 = m.a10 * m.a21 - m.a11 * m.a20;
  c.a10 = m.a02 * m.a21 - m.a01 * m.a22;
  c.a11 = m.a00 * m.a22 - m.a02 * m.a20;
  c.a12 = m.a01 * m.a20 - m.a00 * m.a21;
  c.a20 = m.a01 * m.a12 - m.a02 * m.a11;
  c.a21 = m.a02 * m.a10 - m.a00 * m.a12;
  c.a22 = m.a00 * m.a11 - m.a01 * m.a10;
  return c;
}

// Rotation factor of the polar decomposition F = R S, the unique U*Vh of the SVD
// the reference takes (numerics/linear_algebra.py:131-132).  Newton iteration
// X <- (g X + X^{-T} / g) / 2 with determinant scaling while far from orthogonal;
// converges quadratically, det F < 0 converges to the det = -1 factor exactly as
// U*Vh does.  F = I returns I exactly.
FFMPM_HD Mat3<double> polar_rotation3(const Mat3<double>& F, double& detF) {
  Mat3<double> X = F;
  Mat3<double> c = cofactor3(X);
  double det = X.a00 * c.a00 + X.a01 * c.a01 + X.a02 * c.a02;
  detF = det;
#pragma unroll 1
  for (int it = 0; it < 48; ++it) {
    if (!(fabs(det) > 1e-300)) break;  // singular: R is not unique; keep the last iterate
    double ad = fabs(det);
    double g = 1.0;
    if (ad < 0.7 || ad > 1.4) g = exp(-log(ad) * (1.0 / 3.0));  // |det|^(-1/3)
    double hg = 0.5 * g;
    double hi = 0.5 / (g * det);
    Mat3<double> Y;
    Y.a00 = hg * X.a00 + hi * c.a00; Y.a01 = hg * X.a01 + hi * c.a01; Y.a02 = hg * X.a02 + hi * c.a02;
    Y.a10 = hg * X.a10 + hi * c.a10; Y.a11 = hg * X.a11 + hi * c.a11; Y.a12 = hg * X.a12 + hi * c.a12;
    Y.a20 = hg * X.a20 + hi * c.a20; Y.a21 = hg * X.a21 + hi * c.a21; Y.a22 = hg * X.a22 + hi * c.a22;
    double d2 = (Y.a00 - X.a00) * (Y.a00 - X.a00) + (Y.a01 - X.a01) * (Y.a01 - X.a01) +
                (Y.a02 - X.a02) * (Y.a02 - X.a02) + (Y.a10 - X.a10) * (Y.a10 - X.a10) +
                (Y.a11 - X.a11) * (Y.a11 - X.a11) + (Y.a12 - X.a12) * (Y.a12 - X.a12) +
                (Y.a20 - X.a20) * (Y.a20 - X.a20) + (Y.a21 - X.a21) * (Y.a21 - X.a21) +
                (Y.a22 - X.a22) * (Y.a22 - X.a22);
    X = Y;
    // quadratic convergence: |Y - R| ~ |Y - X|^2 once g == 1
    if (g == 1.0 && d2 < 1e-17) break;
    c = cofactor3(X);
    det = X.a00 * c.a00 + X.a01 * c.a01 + X.a02 * c.a02;
  }
  return X;
}

// affine = -(dt*vol)*(4 inv_dx^2) * (2 mu (F-R) F^T + lam (J-1) J [on ALL entries]) + mass*C
// (solvers/mpm/utils.py:120-135; the broadcast of the lambda term is quirk 2).
FFMPM_HD Mat3<double> fixed_corotated_affine3(const Mat3<double>& F, const Mat3<double>& C,
                                                                double mu, double lam, double mass,
                                                                double dt_vol_dinv) {
  double J;
  Mat3<double> R = polar_rotation3(F, J);
  Mat3<double> D;
  D.a00 = F.a00 - R.a00; D.a01 = F.a01 - R.a01; D.a02 = F.a02 - R.a02;
  D.a10 = F.a10 - R.a10; D.a11 = F.a11 - R.a11; D.a12 = F.a12 - R.a12;
  D.a20 = F.a20 - R.a20; D.a21 = F.a21 - R.a21; D.a22 = F.a22 - R.a22;
  double l = lam * (J - 1.0) * J;
  double m2 = 2.0 * mu;
  Mat3<double> A;
  // (D F^T)[r][c] = sum_k D[r][k] F[c][k]
  A.a00 = -dt_vol_dinv * (m2 * (D.a00 * F.a00 + D.a01 * F.a01 + D.a02 * F.a02) + l) + mass * C.a00;
  A.a01 = -dt_vol_dinv * (m2 * (D.a00 * F.a10 + D.a01 * F.a11 + D.a02 * F.a12) + l) + mass * C.a01;
  A.a02 = -dt_vol_dinv * (m2 * (D.a00 * F.a20 + D.a01 * F.a21 + D.a02 * F.a22) + l) + mass * C.a02;
  A.a10 = -dt_vol_dinv * (m2 * (D.a10 * F.a00 + D.a11 * F.a01 + D.a12 * F.a02) + l) + mass * C.a10;
  A.a11 = -dt_vol_dinv * (m2 * (D.a10 * F.a10 + D.a11 * F.a11 + D.a12 * F.a12) + l) + mass * C.a11;
  A.a12 = -dt_vol_dinv * (m2 * (D.a10 * F.a20 + D.a11 * F.a21 + D.a12 * F.a22) + l) + mass * C.a12;
  A.a20 = -dt_vol_dinv * (m2 * (D.a20 * F.a00 + D.a21 * F.a01 + D.a22 * F.a02) + l) + mass * C.a20;
  A.a21 = -dt_vol_dinv * (m2 * (D.a20 * F.a10 + D.a21 * F.a11 + D.a22 * F.a12) + l) + mass * C.a21;
  A.a22 = -dt_vol_dinv * (m2 * (D.a20 * F.a20 + D.a21 * F.a21 + D.a22 * F.a22) + l) + mass * C.a22;
  return A;
}

// ----------------------------------------------------------------------------
// fp32 evaluation of the same affine matrix in PERTURBATION FORM (used by the fp32
// build when the strain is moderate).  The cancellation that rules out a naive fp32
// stress -- F - R and J - 1 for F close to I -- is removed
// algebraically instead of being bought with fp64.  LEFT form: with F = V R
// (V = (F F^T)^(1/2) the left stretch),
//   (F - R) F^T = F F^T - V = B - B^(1/2),           B = F F^T = I + G,
//   E = F - I                      (exact in fp32 for entries near 1 / 0)
//   G = F F^T - I = E + E^T + E E^T                  (left Cauchy-Green strain)
//   h(G) = (I + G) - (I + G)^(1/2) = G p(G),         p(x) = (1 + x - sqrt(1 + x)) / x = 1/2 + x/8 - x^2/16 + ...
//   J - 1 = det(I + E) - 1 = tr E + (principal 2x2 minors of E) + det E
// so the stress term is a matrix function of the symmetric G alone (no product with F, no SVD, no
// iteration) and every quantity is O(strain): fp32 keeps ~2e-7 RELATIVE accuracy on the stress.
// p is not the truncated Taylor series but its Chebyshev interpolant on the tier's interval
// (scripts/series_economized.py prints the table below and the CPU suite regenerates and compares it):
// degree 2 / 3 / 4 / 5 for ||G||_F < 0.0136 / 0.0436 / 0.112 / 0.15 keeps the error of p below 5e-8 with
// fp32 coefficients; larger strains take the fp64 path.  Measured on B200 against the right form
// F (I - (I + F^T F - I)^(-1/2)) F^T with its 8-term Taylor series that round 1 shipped: four symmetric
// products instead of seven plus two products with F at the headline strain (profiles/r02a).
// ----------------------------------------------------------------------------
struct Sym3f {
  float xx, xy, xz, yy, yz, zz;
};

// Product of two COMMUTING symmetric matrices (symmetric again).
FFMPM_HD Sym3f sym3_mul(const Sym3f& a, const Sym3f& b) {
  Sym3f r;
  r.xx = a.xx * b.xx + a.xy * b.xy + a.xz * b.xz;
  r.xy = a.xx * b.xy + a.xy * b.yy + a.xz * b.yz;
  r.xz = a.xx * b.xz + a.xy * b.yz + a.xz * b.zz;
  r.yy = a.xy * b.xy + a.yy * b.yy + a.yz * b.yz;
  r.yz = a.xy * b.xz + a.yy * b.yz + a.yz * b.zz;
  r.zz = a.xz * b.xz + a.yz * b.yz + a.zz * b.zz;
  return r;
}

constexpr float kPerturbationMaxR = 0.15f;
constexpr int kStressTiers = 4;
// upper bound of ||G||_F per tier (tier t uses p of degree t + 2) and the power-basis coefficients c_0 .. c_deg
// of p per tier (scripts/series_economized.py, TIERS_LEFT).  Functions of literals rather than tables: with the
// tier a template argument and the Horner loop unrolled they fold to immediates in host and device code alike.
FFMPM_HD constexpr float stress_tier_r(int tier) {
  return tier == 0 ? 0.0136f : tier == 1 ? 0.0436f : tier == 2 ? 0.112f : 0.15f;
}
FFMPM_HD constexpr float stress_coef(int tier, int i) {
  switch (tier * 8 + i) {
    // ||G||_F < 0.0136: degree 2
    case 0: return 0.5f; case 1: return 0.12500542402267456f; case 2: return -0.06250379234552383f;
    // ||G||_F < 0.0436: degree 3
    case 8: return 0.5f; case 9: return 0.1249999925494194f; case 10: return -0.06255202740430832f; case 11: return 0.03910152614116669f;
    // ||G||_F < 0.112: degree 4
    case 16: return 0.5f; case 17: return 0.12499897927045822f; case 18: return -0.06249919906258583f; case 19: return 0.03938665986061096f;
    case 20: return -0.02759857103228569f;
    // ||G||_F < 0.15: degree 5
    case 24: return 0.5f; case 25: return 0.125f; case 26: return -0.062495309859514236f; case 27: return 0.039058685302734375f;
    case 28: return -0.027897052466869354f; case 29: return 0.02095773071050644f;
    default: return 0.0f;
  }
}

// E = F - I, G = E + E^T + E E^T and the squared Frobenius norm of G (an upper bound of its spectral radius^2).
FFMPM_HD void left_strain3(const Mat3<float>& F, Mat3<float>& E, Sym3f& G, float& r2) {
  E = F;
  E.a00 -= 1.0f; E.a11 -= 1.0f; E.a22 -= 1.0f;
  G.xx = 2.0f * E.a00 + (E.a00 * E.a00 + E.a01 * E.a01 + E.a02 * E.a02);
  G.yy = 2.0f * E.a11 + (E.a10 * E.a10 + E.a11 * E.a11 + E.a12 * E.a12);
  G.zz = 2.0f * E.a22 + (E.a20 * E.a20 + E.a21 * E.a21 + E.a22 * E.a22);
  G.xy = (E.a01 + E.a10) + (E.a00 * E.a10 + E.a01 * E.a11 + E.a02 * E.a12);
  G.xz = (E.a02 + E.a20) + (E.a00 * E.a20 + E.a01 * E.a21 + E.a02 * E.a22);
  G.yz = (E.a12 + E.a21) + (E.a10 * E.a20 + E.a11 * E.a21 + E.a12 * E.a22);
  r2 = (G.xx * G.xx + G.yy * G.yy + G.zz * G.zz) + 2.0f * (G.xy * G.xy + G.xz * G.xz + G.yz * G.yz);
}

// Tier of the series for a squared norm r2: 0 .. kStressTiers-1, or kStressTiers when the strain is beyond
// the series (also for NaN).
FFMPM_HD int stress_tier_of(float r2) {
  int t = kStressTiers;
#pragma unroll
  for (int i = kStressTiers - 1; i >= 0; --i)
    if (r2 < stress_tier_r(i) * stress_tier_r(i)) t = i;
  return t;
}

// h(G) = G p(G) by Horner, p of degree TIER + 2: TIER + 2 symmetric products, straight-line code with the
// coefficients as immediates (the kernels pick TIER per warp, so no per-lane branching).
template <int TIER>
FFMPM_HD Sym3f left_stress_h(const Sym3f& G) {
  constexpr int deg = TIER + 2;
  const float ca = stress_coef(TIER, deg), cb = stress_coef(TIER, deg - 1);
  Sym3f q;
  q.xx = ca * G.xx + cb; q.yy = ca * G.yy + cb; q.zz = ca * G.zz + cb;
  q.xy = ca * G.xy; q.xz = ca * G.xz; q.yz = ca * G.yz;
#pragma unroll
  for (int i = deg - 2; i >= 0; --i) {
    q = sym3_mul(G, q);
    q.xx += stress_coef(TIER, i); q.yy += stress_coef(TIER, i); q.zz += stress_coef(TIER, i);
  }
  return sym3_mul(G, q);
}

// J - 1 = tr E + principal 2x2 minors + det E, without cancellation.
FFMPM_HD float jm1_of(const Mat3<float>& E) {
  const float trE = E.a00 + E.a11 + E.a22;
  const float c2 = (E.a00 * E.a11 - E.a01 * E.a10) + (E.a00 * E.a22 - E.a02 * E.a20) + (E.a11 * E.a22 - E.a12 * E.a21);
  return trE + c2 + det3(E);
}

// A = k2 * h(G) + kl [on ALL entries, quirk 2] + mc * C with
//   k2 = -(dt vol 4 inv_dx^2) 2 mu * s,  kl = -(dt vol 4 inv_dx^2) lam (J-1) J * s,  mc = mass * s
// (s: a common scale the caller folds in -- the P2G kernels park affine * dx).
FFMPM_HD void affine3_assemble(const Sym3f& H, const Mat3<float>& C, float k2, float kl, float mc, Mat3<float>& A) {
  A.a00 = k2 * H.xx + kl + mc * C.a00; A.a01 = k2 * H.xy + kl + mc * C.a01; A.a02 = k2 * H.xz + kl + mc * C.a02;
  A.a10 = k2 * H.xy + kl + mc * C.a10; A.a11 = k2 * H.yy + kl + mc * C.a11; A.a12 = k2 * H.yz + kl + mc * C.a12;
  A.a20 = k2 * H.xz + kl + mc * C.a20; A.a21 = k2 * H.yz + kl + mc * C.a21; A.a22 = k2 * H.zz + kl + mc * C.a22;
}

// One particle, tier picked from its own strain.  Returns false when the strain is too large for the series
// (caller falls back to fp64).  Used by the thread-per-particle kernels; the warp-autonomous P2G picks the tier
// per warp and calls the pieces above directly.
FFMPM_HD bool fixed_corotated_affine3_f32(const Mat3<float>& F, const Mat3<float>& C, float mu,
                                                            float lam, float mass, float dt_vol_dinv, Mat3<float>& A) {
  Mat3<float> E;
  Sym3f G, H;
  float r2;
  left_strain3(F, E, G, r2);
  const int tier = stress_tier_of(r2);
  if (tier >= kStressTiers) return false;
  if (tier == 0) H = left_stress_h<0>(G);
  else if (tier == 1) H = left_stress_h<1>(G);
  else if (tier == 2) H = left_stress_h<2>(G);
  else H = left_stress_h<3>(G);
  const float jm1 = jm1_of(E);
  const float l = lam * jm1 * (1.0f + jm1);   // lam (J-1) J, broadcast onto ALL entries (quirk 2)
  affine3_assemble(H, C, -dt_vol_dinv * 2.0f * mu, -dt_vol_dinv * l, mass, A);
  return true;
}

// 2D: closed-form rotation with the reference's +1e-10 in the norm
// (numerics/linear_algebra.py:108-113; quirks 3 and 12), stress as utils.py:75-92.
FFMPM_HD Mat2<double> fixed_corotated_affine2(const Mat2<double>& F, const Mat2<double>& C,
                                                                double mu, double lam, double mass,
                                                                double dt_vol_dinv) {
  double J = F.a00 * F.a11 - F.a01 * F.a10;
  double x = F.a00 + F.a11;
  double y = F.a10 - F.a01;
  double scale = 1.0 / (sqrt(x * x + y * y) + 1e-10);
  double c = x * scale, s = y * scale;
  double d00 = F.a00 - c, d01 = F.a01 + s, d10 = F.a10 - s, d11 = F.a11 - c;
  double l = lam * (J - 1.0) * J;
  double m2 = 2.0 * mu;
  Mat2<double> A;
  A.a00 = -dt_vol_dinv * (m2 * (d00 * F.a00 + d01 * F.a01) + l) + mass * C.a00;
  A.a01 = -dt_vol_dinv * (m2 * (d00 * F.a10 + d01 * F.a11) + l) + mass * C.a01;
  A.a10 = -dt_vol_dinv * (m2 * (d10 * F.a00 + d11 * F.a01) + l) + mass * C.a10;
  A.a11 = -dt_vol_dinv * (m2 * (d10 * F.a10 + d11 * F.a11) + l) + mass * C.a11;
  return A;
}

// The same in fp32 WITHOUT cancellation (fp32 build, 2D): every term is formed from E = F - I.
//   x = F00 + F11 = 2 + tr E,  w = F10 - F01 = E10 - E01,  r = sqrt(x^2 + w^2),  R0 = [[x, -w], [w, x]] / r
//   q = r - 2 = (r^2 - 4) / (r + 2) = (4 tr E + (tr E)^2 + w^2) / (r + 2)
//   r (F - R0) = q F + D,  D = [[E00 - E11, E01 + E10], [E01 + E10, E11 - E00]]     (2 F - [[x, -w], [w, x]] = D)
// so F - R0 is built from the strain-sized numbers q and D only: D is exact for a pure rotation (Sterbenz) and the
// error of q is second order in the rotation angle -- no difference of O(1) (or O(angle)) numbers anywhere.
//   The reference's  scale = 1 / (r + 1e-10)  (numerics/linear_algebra.py:108-113, quirks 3, 12) gives
//   R = R0 (1 - e),  e = 1e-10 / (r + 1e-10),  so  F - R = (q F + D + e [[x, -w], [w, x]]) / r:
// the spurious isotropic stress of the reference is ADDED analytically instead of being lost below fp32 resolution.
//   J - 1 = tr E + det E.
// Returns false when r is not safely positive (F close to a reflection, or NaN) -- the caller takes the fp64 form.
// This removes a double-precision square root and division from the latency chain of every 2D particle.
FFMPM_HD bool fixed_corotated_affine2_f32(const Mat2<float>& F, const Mat2<float>& C, float mu, float lam, float mass,
                                          float dt_vol_dinv, Mat2<float>& A) {
  const float e00 = F.a00 - 1.0f, e01 = F.a01, e10 = F.a10, e11 = F.a11 - 1.0f;
  const float tr = e00 + e11, w = e10 - e01, sy = e10 + e01, dd = e00 - e11;
  const float x = 2.0f + tr;
  const float s2 = x * x + w * w;
#ifdef __CUDA_ARCH__
  // device: the approximate unit (MUFU.RSQ / MUFU.RCP, 2 ulp) -- every quantity formed with them is strain-sized or
  // multiplies one, so 2e-7 relative is far inside the 1e-5 bar; IEEE sqrt and three IEEE divisions cost ~30 instructions
  const float r = s2 * rsqrtf(s2);               // s2 == 0: NaN, declined below
  if (!(r > 1e-3f)) return false;
  const float inv_r = __fdividef(1.0f, r);
  const float q = __fdividef(4.0f * tr + tr * tr + w * w, r + 2.0f);
  const float e = __fdividef(1e-10f, r + 1e-10f);
#else
  const float r = sqrtf(s2);
  if (!(r > 1e-3f)) return false;
  const float inv_r = 1.0f / r;
  const float q = (4.0f * tr + tr * tr + w * w) / (r + 2.0f);
  const float e = 1e-10f / (r + 1e-10f);
#endif
  const float ex = e * x, ew = e * w;
  const float d00 = (q * F.a00 + dd + ex) * inv_r, d01 = (q * e01 + sy - ew) * inv_r;
  const float d10 = (q * e10 + sy + ew) * inv_r, d11 = (q * F.a11 - dd + ex) * inv_r;
  const float jm1 = tr + (e00 * e11 - e01 * e10);
  const float l = lam * jm1 * (1.0f + jm1);        // lam (J-1) J, on ALL entries (quirk 2)
  const float m2 = 2.0f * mu;
  A.a00 = -dt_vol_dinv * (m2 * (d00 * F.a00 + d01 * F.a01) + l) + mass * C.a00;
  A.a01 = -dt_vol_dinv * (m2 * (d00 * F.a10 + d01 * F.a11) + l) + mass * C.a01;
  A.a10 = -dt_vol_dinv * (m2 * (d10 * F.a00 + d11 * F.a01) + l) + mass * C.a10;
  A.a11 = -dt_vol_dinv * (m2 * (d10 * F.a10 + d11 * F.a11) + l) + mass * C.a11;
  return true;
}

// 2x2 SVD round trip of two_d/g2p.py:37-43: F <- U diag(sig) Vh^T with (U, sig, Vh)
// LAPACK's SVD and "V.T" applied to what is already Vh.  Measured convention of
// dgesdd on 2x2 input (tests/golden/quirk2d.npz): U is always a reflection; Vh is a
// symmetric reflection when det F > 0 (so Vh^T = Vh and the round trip is the
// identity) and a proper rotation [[c, s], [-s, c]] with (c, s) the principal right
// singular vector when det F < 0, so that
//     U S Vh^T = (U S Vh) (Vh^T)^2 = F * Rot(2 theta),
// which is independent of the sign LAPACK picks for the singular vectors.
// For `snow` the singular values are first clamped to [1-2.5e-2, 1+7.5e-3]:
//     U S' Vh = F * (r0 v1 v1^T + r1 v2 v2^T),  r_i = clamp(s_i) / s_i.
// Returns det of the result through `det_out` (sign follows det F).
FFMPM_HD Mat2<double> svd_roundtrip2(const Mat2<double>& F, bool snow, double& det_out) {
  double detF = F.a00 * F.a11 - F.a01 * F.a10;
  double m00 = F.a00 * F.a00 + F.a10 * F.a10;
  double m01 = F.a00 * F.a01 + F.a10 * F.a11;
  double m11 = F.a01 * F.a01 + F.a11 * F.a11;
  double dm = m00 - m11, om = 2.0 * m01;
  double h = sqrt(dm * dm + om * om);
  double c2 = 1.0, s2 = 0.0;
  if (h > 0.0) { c2 = dm / h; s2 = om / h; }
  Mat2<double> G = F;
  det_out = detF;
  if (snow) {
    double tr = m00 + m11;
    double sg0 = sqrt(fmax(0.5 * (tr + h), 0.0)), sg1 = sqrt(fmax(0.5 * (tr - h), 0.0));
    double c0 = fmin(fmax(sg0, 1.0 - 2.5e-2), 1.0 + 7.5e-3);
    double c1 = fmin(fmax(sg1, 1.0 - 2.5e-2), 1.0 + 7.5e-3);
    double r0 = sg0 > 0.0 ? c0 / sg0 : 0.0, r1 = sg1 > 0.0 ? c1 / sg1 : 0.0;
    // P1 = v1 v1^T = [[(1+c2)/2, s2/2], [s2/2, (1-c2)/2]]
    double p00 = 0.5 * (1.0 + c2), p01 = 0.5 * s2, p11 = 0.5 * (1.0 - c2);
    double q00 = r0 * p00 + r1 * (1.0 - p00), q01 = (r0 - r1) * p01, q11 = r0 * p11 + r1 * (1.0 - p11);
    G.a00 = F.a00 * q00 + F.a01 * q01; G.a01 = F.a00 * q01 + F.a01 * q11;
    G.a10 = F.a10 * q00 + F.a11 * q01; G.a11 = F.a10 * q01 + F.a11 * q11;
    det_out = (detF < 0.0 ? -1.0 : 1.0) * c0 * c1;
  }
  if (detF < 0.0) {
    Mat2<double> Hm;
    Hm.a00 = G.a00 * c2 + G.a01 * s2; Hm.a01 = -G.a00 * s2 + G.a01 * c2;
    Hm.a10 = G.a10 * c2 + G.a11 * s2; Hm.a11 = -G.a10 * s2 + G.a11 * c2;
    G = Hm;
  }
  return G;
}

}  // namespace ffmpm

// 3x3 SVD with LAPACK's singular-vector signs, and the 3D snow return map built on it.
//
// three_d/g2p.py:48-58 (model == "snow") computes  U @ diag(clip(sig)) @ Vh.T  -- numpy's Vh transposed once
// more -- which is not invariant under the sign freedom (u_i, v_i) -> (-u_i, -v_i) of an SVD: the reference's result
// is a function of the signs DGESDD returns, and those are a function of its operation sequence.  svd3_lapack()
// therefore walks that sequence for a 3x3 input (JOBZ = 'A', path 5): DGEBD2 (Householder bidiagonalisation: H1 on
// column 1, G1 on row 1, H2 on column 2; H3 and G2 are identities), DBDSQR through DBDSDC -> DLASDQ (n <= 25: zero-shift
// and implicit-shift QR sweeps in the direction of the larger end, DLASV2 on 2x2 blocks, negative singular values made
// positive by negating rows of VT, sort), DORMBR back-transformation.  DLARTG is LAPACK >= 3.10's (c >= 0).  LAPACK is
// not part of /root/reference (numpy links OpenBLAS 0.3.30 there).  tests/test_lapack_svd3.py compiles this file for the
// host and pins it against np.linalg.svd itself on 10^5 matrices (signs and values) and against the reference's own
// F_out / Jp_out in tests/golden/snow3d.npz.  fp64 throughout; plain arithmetic (tests/native/kernel_math_host.cu).
#pragma once
#include "mpm_math.cuh"

namespace ffmpm {

#define FFMPM_HDN __host__ __device__ __noinline__

namespace svd3 {

constexpr double kEps = 1.1102230246251565e-16;    // DLAMCH('Epsilon') = 2^-53
constexpr double kUnfl = 2.2250738585072014e-308;  // DLAMCH('Safe minimum')

__host__ __device__ inline double sgn(double a, double b) {   // Fortran SIGN(a, b)
  return signbit(b) ? -fabs(a) : fabs(a);
}

__host__ __device__ inline void lartg(double f, double g, double& c, double& s, double& r) {
  if (g == 0.0) { c = 1.0; s = 0.0; r = f; return; }
  if (f == 0.0) { c = 0.0; s = sgn(1.0, g); r = fabs(g); return; }
  const double d = sqrt(f * f + g * g);
  c = fabs(f) / d;
  r = sgn(d, f);
  s = g / r;
}

// Householder reflector H = I - tau [1; x][1; x]^T with H [alpha; x] = [beta; 0]; nx = 1 or 2
__host__ __device__ inline void larfg(double& alpha, double* x, int nx, double& tau) {
  double xn2 = 0.0;
  for (int i = 0; i < nx; ++i) xn2 += x[i] * x[i];
  if (xn2 == 0.0) { tau = 0.0; return; }
  const double beta = -sgn(sqrt(alpha * alpha + xn2), alpha);
  tau = (beta - alpha) / beta;
  const double sc = 1.0 / (alpha - beta);
  for (int i = 0; i < nx; ++i) x[i] *= sc;
  alpha = beta;
}

__host__ __device__ inline double las2_min(double f, double g, double h) {   // DLAS2: smaller singular value of [[f, g], [0, h]]
  const double fa = fabs(f), ga = fabs(g), ha = fabs(h);
  const double fhmn = fmin(fa, ha), fhmx = fmax(fa, ha);
  if (fhmn == 0.0) return 0.0;
  const double as = 1.0 + fhmn / fhmx, at = (fhmx - fhmn) / fhmx;
  if (ga < fhmx) {
    const double au = (ga / fhmx) * (ga / fhmx);
    return fhmn * (2.0 / (sqrt(as * as + au) + sqrt(at * at + au)));
  }
  const double au = fhmx / ga;
  if (au == 0.0) return (fhmn * fhmx) / ga;
  const double c = 1.0 / (sqrt(1.0 + (as * au) * (as * au)) + sqrt(1.0 + (at * au) * (at * au)));
  return 2.0 * ((fhmn * c) * au);
}

// DLASV2: SVD of [[f, g], [0, h]]
__host__ __device__ inline void lasv2(double f, double g, double h, double& ssmin, double& ssmax, double& snr, double& csr,
                                      double& snl, double& csl) {
  double ft = f, fa = fabs(f), ht = h, ha = fabs(h);
  int pmax = 1;
  const bool swap = ha > fa;
  if (swap) {
    pmax = 3;
    double t = ft; ft = ht; ht = t;
    t = fa; fa = ha; ha = t;
  }
  const double gt = g, ga = fabs(g);
  double clt, crt, slt, srt;
  if (ga == 0.0) {
    ssmin = ha; ssmax = fa; clt = 1.0; crt = 1.0; slt = 0.0; srt = 0.0;
  } else {
    bool gasmal = true;
    if (ga > fa) {
      pmax = 2;
      if (fa / ga < kEps) {
        gasmal = false;
        ssmax = ga;
        ssmin = ha > 1.0 ? fa / (ga / ha) : (fa / ga) * ha;
        clt = 1.0; slt = ht / gt; srt = 1.0; crt = ft / gt;
      }
    }
    if (gasmal) {
      const double d = fa - ha;
      double l = d == fa ? 1.0 : d / fa;
      const double m = gt / ft;
      double t = 2.0 - l;
      const double mm = m * m, tt = t * t;
      const double s = sqrt(tt + mm);
      const double r = l == 0.0 ? fabs(m) : sqrt(l * l + mm);
      const double a = 0.5 * (s + r);
      ssmin = ha / a; ssmax = fa * a;
      if (mm == 0.0) {
        if (l == 0.0) t = sgn(2.0, ft) * sgn(1.0, gt);
        else t = gt / sgn(d, ft) + m / t;
      } else {
        t = (m / (s + t) + m / (r + l)) * (1.0 + a);
      }
      l = sqrt(t * t + 4.0);
      crt = 2.0 / l; srt = t / l;
      clt = (crt + srt * m) / a;
      slt = (ht / ft) * srt / a;
    }
  }
  if (swap) { csl = srt; snl = crt; csr = slt; snr = clt; }
  else { csl = clt; snl = slt; csr = crt; snr = srt; }
  double tsign;
  if (pmax == 1) tsign = sgn(1.0, csr) * sgn(1.0, csl) * sgn(1.0, f);
  else if (pmax == 2) tsign = sgn(1.0, snr) * sgn(1.0, csl) * sgn(1.0, g);
  else tsign = sgn(1.0, snr) * sgn(1.0, snl) * sgn(1.0, h);
  ssmax = sgn(ssmax, tsign);
  ssmin = sgn(ssmin, tsign * sgn(1.0, f) * sgn(1.0, h));
}

__host__ __device__ inline void rot_rows(double (*vt)[3], int i, int j, double c, double s) {
  for (int k = 0; k < 3; ++k) {
    const double a = vt[i][k], b = vt[j][k];
    vt[i][k] = c * a + s * b;
    vt[j][k] = c * b - s * a;
  }
}
__host__ __device__ inline void rot_cols(double (*u)[3], int i, int j, double c, double s) {
  for (int k = 0; k < 3; ++k) {
    const double a = u[k][i], b = u[k][j];
    u[k][i] = c * a + s * b;
    u[k][j] = c * b - s * a;
  }
}

// DBDSQR('U', 3, 3, 3, 0): d[3], e[2] in place; rows of vt and columns of u rotate.  Indices are 1-based in the
// comments and in ll / m, as in the Fortran text; D(i) = d[i-1].  Returns false when the sweeps did not converge.
__host__ __device__ inline bool bdsqr3(double* d, double* e, double (*vt)[3], double (*u)[3]) {
  const int n = 3;
  const double tol = fmax(10.0, fmin(100.0, pow(kEps, -0.125))) * kEps;
  double sminoa = fabs(d[0]);
  if (sminoa != 0.0) {
    double mu = sminoa;
    for (int i = 1; i < n; ++i) {
      mu = fabs(d[i]) * (mu / (mu + fabs(e[i - 1])));
      sminoa = fmin(sminoa, mu);
      if (sminoa == 0.0) break;
    }
  }
  sminoa /= sqrt((double)n);
  const double thresh = fmax(tol * sminoa, 6.0 * (n * (n * kUnfl)));
  const int maxit = 6 * n * n;
  int it = 0, oldll = -1, oldm = -1, idir = 0, m = n;
  double rc1[2], rs1[2], rc2[2], rs2[2];   // the sweep's rotations, applied to the vectors after the sweep (DLASR)
  while (m > 1) {
    if (it > maxit) return false;
    double smax = fabs(d[m - 1]);
    bool split = false;
    int ll = 0;
    for (int lll = 1; lll < m; ++lll) {
      ll = m - lll;
      const double abss = fabs(d[ll - 1]), abse = fabs(e[ll - 1]);
      if (abse <= thresh) { split = true; break; }
      smax = fmax(smax, fmax(abss, abse));
    }
    if (split) {
      e[ll - 1] = 0.0;
      if (ll == m - 1) { m -= 1; continue; }   // the bottom singular value has converged
    } else {
      ll = 0;
    }
    ll += 1;
    if (ll == m - 1) {   // 2x2 block
      double sigmn, sigmx, sinr, cosr, sinl, cosl;
      lasv2(d[m - 2], e[m - 2], d[m - 1], sigmn, sigmx, sinr, cosr, sinl, cosl);
      d[m - 2] = sigmx; e[m - 2] = 0.0; d[m - 1] = sigmn;
      rot_rows(vt, m - 2, m - 1, cosr, sinr);
      rot_cols(u, m - 2, m - 1, cosl, sinl);
      m -= 2;
      continue;
    }
    if (ll > oldm || m < oldll) idir = fabs(d[ll - 1]) >= fabs(d[m - 1]) ? 1 : 2;
    double sminl;
    bool conv = false;
    if (idir == 1) {
      if (fabs(e[m - 2]) <= tol * fabs(d[m - 1])) { e[m - 2] = 0.0; continue; }
      double mu = fabs(d[ll - 1]);
      sminl = mu;
      for (int lll = ll; lll < m; ++lll) {
        if (fabs(e[lll - 1]) <= tol * mu) { e[lll - 1] = 0.0; conv = true; break; }
        mu = fabs(d[lll]) * (mu / (mu + fabs(e[lll - 1])));
        sminl = fmin(sminl, mu);
      }
    } else {
      if (fabs(e[ll - 1]) <= tol * fabs(d[ll - 1])) { e[ll - 1] = 0.0; continue; }
      double mu = fabs(d[m - 1]);
      sminl = mu;
      for (int lll = m - 1; lll >= ll; --lll) {
        if (fabs(e[lll - 1]) <= tol * mu) { e[lll - 1] = 0.0; conv = true; break; }
        mu = fabs(d[lll - 1]) * (mu / (mu + fabs(e[lll - 1])));
        sminl = fmin(sminl, mu);
      }
    }
    if (conv) continue;
    oldll = ll; oldm = m;
    double shift;
    if (n * tol * (sminl / smax) <= fmax(kEps, 0.01 * tol)) {
      shift = 0.0;
    } else {
      double sll;
      if (idir == 1) { sll = fabs(d[ll - 1]); shift = las2_min(d[m - 2], e[m - 2], d[m - 1]); }
      else { sll = fabs(d[m - 1]); shift = las2_min(d[ll - 1], e[ll - 1], d[ll]); }
      if (sll > 0.0 && (shift / sll) * (shift / sll) < kEps) shift = 0.0;
    }
    it += m - ll;
    int nr = 0;
    if (shift == 0.0) {
      double cs = 1.0, oldcs = 1.0, oldsn = 0.0, sn, r;
      if (idir == 1) {   // chase the bulge from top to bottom
        for (int i = ll; i < m; ++i) {
          lartg(d[i - 1] * cs, e[i - 1], cs, sn, r);
          if (i > ll) e[i - 2] = oldsn * r;
          lartg(oldcs * r, d[i] * sn, oldcs, oldsn, d[i - 1]);
          rc1[nr] = cs; rs1[nr] = sn; rc2[nr] = oldcs; rs2[nr] = oldsn; ++nr;
        }
        const double h = d[m - 1] * cs;
        d[m - 1] = h * oldcs;
        e[m - 2] = h * oldsn;
        for (int k = 0; k < nr; ++k) rot_rows(vt, ll - 1 + k, ll + k, rc1[k], rs1[k]);
        for (int k = 0; k < nr; ++k) rot_cols(u, ll - 1 + k, ll + k, rc2[k], rs2[k]);
        if (fabs(e[m - 2]) <= thresh) e[m - 2] = 0.0;
      } else {           // from bottom to top
        for (int i = m; i > ll; --i) {
          lartg(d[i - 1] * cs, e[i - 2], cs, sn, r);
          if (i < m) e[i - 1] = oldsn * r;
          lartg(oldcs * r, d[i - 2] * sn, oldcs, oldsn, d[i - 1]);
          rc1[nr] = cs; rs1[nr] = -sn; rc2[nr] = oldcs; rs2[nr] = -oldsn; ++nr;
        }
        const double h = d[ll - 1] * cs;
        d[ll - 1] = h * oldcs;
        e[ll - 1] = h * oldsn;
        for (int k = 0; k < nr; ++k) rot_rows(vt, m - 2 - k, m - 1 - k, rc2[k], rs2[k]);
        for (int k = 0; k < nr; ++k) rot_cols(u, m - 2 - k, m - 1 - k, rc1[k], rs1[k]);
        if (fabs(e[ll - 1]) <= thresh) e[ll - 1] = 0.0;
      }
    } else {
      double cosr, sinr, cosl, sinl, r;
      if (idir == 1) {
        double f = (fabs(d[ll - 1]) - shift) * (sgn(1.0, d[ll - 1]) + shift / d[ll - 1]);
        double g = e[ll - 1];
        for (int i = ll; i < m; ++i) {
          lartg(f, g, cosr, sinr, r);
          if (i > ll) e[i - 2] = r;
          f = cosr * d[i - 1] + sinr * e[i - 1];
          e[i - 1] = cosr * e[i - 1] - sinr * d[i - 1];
          g = sinr * d[i];
          d[i] = cosr * d[i];
          lartg(f, g, cosl, sinl, r);
          d[i - 1] = r;
          f = cosl * e[i - 1] + sinl * d[i];
          d[i] = cosl * d[i] - sinl * e[i - 1];
          if (i < m - 1) { g = sinl * e[i]; e[i] = cosl * e[i]; }
          rc1[nr] = cosr; rs1[nr] = sinr; rc2[nr] = cosl; rs2[nr] = sinl; ++nr;
        }
        e[m - 2] = f;
        for (int k = 0; k < nr; ++k) rot_rows(vt, ll - 1 + k, ll + k, rc1[k], rs1[k]);
        for (int k = 0; k < nr; ++k) rot_cols(u, ll - 1 + k, ll + k, rc2[k], rs2[k]);
        if (fabs(e[m - 2]) <= thresh) e[m - 2] = 0.0;
      } else {
        double f = (fabs(d[m - 1]) - shift) * (sgn(1.0, d[m - 1]) + shift / d[m - 1]);
        double g = e[m - 2];
        for (int i = m; i > ll; --i) {
          lartg(f, g, cosr, sinr, r);
          if (i < m) e[i - 1] = r;
          f = cosr * d[i - 1] + sinr * e[i - 2];
          e[i - 2] = cosr * e[i - 2] - sinr * d[i - 1];
          g = sinr * d[i - 2];
          d[i - 2] = cosr * d[i - 2];
          lartg(f, g, cosl, sinl, r);
          d[i - 1] = r;
          f = cosl * e[i - 2] + sinl * d[i - 2];
          d[i - 2] = cosl * d[i - 2] - sinl * e[i - 2];
          if (i > ll + 1) { g = sinl * e[i - 3]; e[i - 3] = cosl * e[i - 3]; }
          rc1[nr] = cosr; rs1[nr] = -sinr; rc2[nr] = cosl; rs2[nr] = -sinl; ++nr;
        }
        e[ll - 1] = f;
        if (fabs(e[ll - 1]) <= thresh) e[ll - 1] = 0.0;
        for (int k = 0; k < nr; ++k) rot_rows(vt, m - 2 - k, m - 1 - k, rc2[k], rs2[k]);
        for (int k = 0; k < nr; ++k) rot_cols(u, m - 2 - k, m - 1 - k, rc1[k], rs1[k]);
      }
    }
  }
  // singular values positive (rows of VT change sign), then decreasing order
  for (int i = 0; i < n; ++i)
    if (d[i] < 0.0) {
      d[i] = -d[i];
      for (int k = 0; k < 3; ++k) vt[i][k] = -vt[i][k];
    }
  for (int i = 1; i < n; ++i) {
    int isub = 1;
    double smin = d[0];
    for (int j = 2; j <= n + 1 - i; ++j)
      if (d[j - 1] <= smin) { isub = j; smin = d[j - 1]; }
    const int last = n + 1 - i;
    if (isub != last) {
      double t = d[isub - 1]; d[isub - 1] = d[last - 1]; d[last - 1] = t;
      for (int k = 0; k < 3; ++k) {
        t = vt[isub - 1][k]; vt[isub - 1][k] = vt[last - 1][k]; vt[last - 1][k] = t;
        t = u[k][isub - 1]; u[k][isub - 1] = u[k][last - 1]; u[k][last - 1] = t;
      }
    }
  }
  return true;
}

}  // namespace svd3

// a (row-major 3x3) = u diag(sig) vt, singular vectors signed as DGESDD signs them.  `a` is destroyed.
FFMPM_HDN bool svd3_lapack(double (*a)[3], double (*u)[3], double* sig, double (*vt)[3]) {
  using namespace svd3;
  // DGEBD2
  double v1[2] = {a[1][0], a[2][0]}, tauq1, d1 = a[0][0];
  larfg(d1, v1, 2, tauq1);
  const double h1[3] = {1.0, v1[0], v1[1]};
  for (int j = 1; j < 3; ++j) {
    const double w = h1[0] * a[0][j] + h1[1] * a[1][j] + h1[2] * a[2][j];
    for (int i = 0; i < 3; ++i) a[i][j] -= tauq1 * h1[i] * w;
  }
  double g1v[1] = {a[0][2]}, taup1, e1 = a[0][1];
  larfg(e1, g1v, 1, taup1);
  const double g1[2] = {1.0, g1v[0]};
  for (int i = 1; i < 3; ++i) {
    const double w = a[i][1] * g1[0] + a[i][2] * g1[1];
    a[i][1] -= taup1 * w * g1[0];
    a[i][2] -= taup1 * w * g1[1];
  }
  double v2[1] = {a[2][1]}, tauq2, d2 = a[1][1];
  larfg(d2, v2, 1, tauq2);
  const double h2[2] = {1.0, v2[0]};
  {
    const double w = h2[0] * a[1][2] + h2[1] * a[2][2];
    a[1][2] -= tauq2 * h2[0] * w;
    a[2][2] -= tauq2 * h2[1] * w;
  }
  double e[2] = {e1, a[1][2]};
  sig[0] = d1; sig[1] = d2; sig[2] = a[2][2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) u[i][j] = vt[i][j] = i == j ? 1.0 : 0.0;
  const bool ok = bdsqr3(sig, e, vt, u);
  // DORMBR('Q','L','N'): U <- H1 H2 U;  DORMBR('P','R','T'): VT <- VT G1
  for (int j = 0; j < 3; ++j) {
    const double w = h2[0] * u[1][j] + h2[1] * u[2][j];
    u[1][j] -= tauq2 * h2[0] * w;
    u[2][j] -= tauq2 * h2[1] * w;
  }
  for (int j = 0; j < 3; ++j) {
    const double w = h1[0] * u[0][j] + h1[1] * u[1][j] + h1[2] * u[2][j];
    for (int i = 0; i < 3; ++i) u[i][j] -= tauq1 * h1[i] * w;
  }
  for (int i = 0; i < 3; ++i) {
    const double w = vt[i][1] * g1[0] + vt[i][2] * g1[1];
    vt[i][1] -= taup1 * w * g1[0];
    vt[i][2] -= taup1 * w * g1[1];
  }
  return ok;
}

// three_d/g2p.py:46-59 for model == "snow", on one particle: F_ = (I + dt C) F in fp64 from the stored F and the
// NEW C, singular values clamped to [1 - 2.5e-2, 1 + 7.5e-3], F <- U diag(sig) Vh^T (the `V.T` of g2p.py:55),
// Jp <- clip(Jp det(F_) / (det(F) + 1e-10), 0.6, 20).  Evaluating F_ in fp64 here -- instead of reading an fp32 F_ back --
// keeps the perturbation F_ - I, which alone decides the singular vectors of a near-isotropic F, at the accuracy of C.
FFMPM_HDN void snow_return_map3(const double* Fold, const double* C, double dt, double jp_in, double* Fnew, double& jp_out) {
  double a[3][3], u[3][3], vt[3][3], sig[3], f_[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = Fold[i * 3 + j];
      for (int k = 0; k < 3; ++k) acc += dt * C[i * 3 + k] * Fold[k * 3 + j];
      f_[i][j] = a[i][j] = acc;
    }
  const double old_j = f_[0][0] * (f_[1][1] * f_[2][2] - f_[1][2] * f_[2][1]) - f_[0][1] * (f_[1][0] * f_[2][2] - f_[1][2] * f_[2][0]) +
                       f_[0][2] * (f_[1][0] * f_[2][1] - f_[1][1] * f_[2][0]);
  svd3_lapack(a, u, sig, vt);
  for (int k = 0; k < 3; ++k) sig[k] = fmin(fmax(sig[k], 1.0 - 2.5e-2), 1.0 + 7.5e-3);
  double g[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      // (U S Vh^T)_ij = sum_k U_ik s_k Vh^T_kj = sum_k U_ik s_k Vh_jk
      g[i][j] = u[i][0] * sig[0] * vt[j][0] + u[i][1] * sig[1] * vt[j][1] + u[i][2] * sig[2] * vt[j][2];
      Fnew[i * 3 + j] = g[i][j];
    }
  const double det = g[0][0] * (g[1][1] * g[2][2] - g[1][2] * g[2][1]) - g[0][1] * (g[1][0] * g[2][2] - g[1][2] * g[2][0]) +
                     g[0][2] * (g[1][0] * g[2][1] - g[1][1] * g[2][0]) + 1e-10;
  jp_out = fmin(fmax(jp_in * old_j / det, 0.6), 20.0);
}

#ifdef __CUDACC__
// The return map over the live state after a G2P that left F alone (KEEPF): thread per particle, fp64 registers,
// the SVD's small matrices in local memory.  Not a tuned kernel: the 3D snow branch is unreachable from the reference's
// driver (mls_mpm.py:58 passes "neo_hookean"); it is here so that the phase-level API answers like three_d/g2p.py does.
template <typename T>
__global__ void __launch_bounds__(128) snow_project3_kernel(DevCfg cfg, StateView<T> s, long long n) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long st = s.stride;
  double F[9], C[9], Fn[9], jp;
#pragma unroll
  for (int k = 0; k < 9; ++k) { F[k] = (double)s.F[k * st + p]; C[k] = (double)s.C[k * st + p]; }
  snow_return_map3(F, C, cfg.dt, (double)s.Jp[p], Fn, jp);
#pragma unroll
  for (int k = 0; k < 9; ++k) s.F[k * st + p] = (T)Fn[k];
  s.Jp[p] = (T)jp;
}
#endif

}  // namespace ffmpm

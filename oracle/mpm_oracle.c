/*
 * mpm_oracle.c -- plain-C (fp64) restatement of the FEMFlow MLS-MPM substep.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ as a second, independent checker (its
 * 3x3 polar factor comes from a one-sided Jacobi SVD, the CUDA path uses a Newton
 * iteration, the NumPy oracle uses LAPACK) and by bench.py as the CPU baseline
 * ("port").  The product package never links or loads it.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it against the
 * outputs of the real reference stored in tests/golden/.
 *
 * Each function names the reference lines it follows (paths relative to
 * /root/reference/femflow).  State is SoA/AoS-mixed exactly like the NumPy oracle:
 * x (N,d), v (N,d), F (N,d,d), C (N,d,d) row-major doubles; grids in the reference
 * layout grid_velocity (G..,d), grid_mass (G..,1).
 *
 * Loops over particles are OpenMP-parallel (the reference is serial numba); the
 * scatter uses `omp atomic`, so summation order -- and only that -- differs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* three_d/p2g.py:50-55: base = trunc(x*inv_dx - 0.5), fx = x*inv_dx - base, weights */
static inline void base_fx_w(double x, double inv_dx, int64_t* base, double* fx, double w[3]) {
  double s = x * inv_dx;
  *base = (int64_t)(s - 0.5); /* C cast truncates toward zero, like astype(int64) */
  *fx = s - (double)(*base);
  w[0] = 0.5 * (1.5 - *fx) * (1.5 - *fx);
  w[1] = 0.75 - (*fx - 1.0) * (*fx - 1.0);
  w[2] = 0.5 * (*fx - 0.5) * (*fx - 0.5);
}

/* numerics/linear_algebra.py:119-135: R = U Vh of the SVD.  One-sided (Hestenes)
 * Jacobi: rotate column pairs of A = F V until orthogonal; then U = A Sigma^-1 and
 * R = U V^T.  Sign/ordering conventions cancel in the product. */
static void polar3(const double F[9], double R[9]) {
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(A, F, sizeof(A));
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, b = 0, c = 0;
        for (int r = 0; r < 3; ++r) {
          a += A[3 * r + p] * A[3 * r + p];
          b += A[3 * r + q] * A[3 * r + q];
          c += A[3 * r + p] * A[3 * r + q];
        }
        if (fabs(c) <= 1e-300) continue;
        double rel = fabs(c) / sqrt(a * b > 0 ? a * b : 1e-300);
        if (rel > off) off = rel;
        double zeta = (b - a) / (2.0 * c);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int r = 0; r < 3; ++r) {
          double ap = A[3 * r + p], aq = A[3 * r + q];
          A[3 * r + p] = cs * ap - sn * aq;
          A[3 * r + q] = sn * ap + cs * aq;
          double vp = V[3 * r + p], vq = V[3 * r + q];
          V[3 * r + p] = cs * vp - sn * vq;
          V[3 * r + q] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-16) break;
  }
  double U[9];
  for (int c = 0; c < 3; ++c) {
    double nrm = sqrt(A[c] * A[c] + A[3 + c] * A[3 + c] + A[6 + c] * A[6 + c]);
    double inv = nrm > 0 ? 1.0 / nrm : 0.0;
    for (int r = 0; r < 3; ++r) U[3 * r + c] = A[3 * r + c] * inv;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = U[3 * r] * V[3 * c] + U[3 * r + 1] * V[3 * c + 1] + U[3 * r + 2] * V[3 * c + 2];
}

static inline double det3(const double m[9]) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

/* solvers/mpm/utils.py:95-135 fixed_corotated_stress_3d (lambda term on ALL entries) */
static void affine3(const double F[9], const double C[9], double inv_dx, double mu, double lam, double dt,
                    double volume, double mass, double out[9]) {
  double R[9];
  double J = det3(F);
  polar3(F, R);
  double D_inv = 4 * inv_dx * inv_dx;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double acc = 0;
      for (int k = 0; k < 3; ++k) acc += (2 * mu * (F[3 * r + k] - R[3 * r + k])) * F[3 * c + k];
      double PF = acc + lam * (J - 1) * J;
      out[3 * r + c] = -(dt * volume) * (D_inv * PF) + mass * C[3 * r + c];
    }
}

/* three_d/p2g.py:14-80.  Returns the number of out-of-grid particles (the
 * reference raises RuntimeError on the first one). */
int64_t oracle_p2g_3d(int64_t n, int64_t G, double inv_dx, double hardening, double dx, double dt, double volume,
                      double* grid_velocity, double* grid_mass, const double* x, const double* mass,
                      const double* mu0, const double* lam0, const double* v, const double* F, const double* C,
                      const double* Jp, int snow) {
  int64_t n_oob = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_oob)
  for (int64_t p = 0; p < n; ++p) {
    int64_t b[3];
    double fx[3], w[3][3];
    int bad = 0;
    for (int d = 0; d < 3; ++d) {
      base_fx_w(x[3 * p + d], inv_dx, &b[d], &fx[d], w[d]);
      if (b[d] < 0 || b[d] + 2 >= G) bad = 1; /* utils.py:138-150 with res = G */
    }
    if (bad) { n_oob += 1; continue; }
    double e = snow ? exp(hardening * (1.0 - Jp[p])) : hardening; /* utils.py:7-49 */
    double A[9];
    affine3(F + 9 * p, C + 9 * p, inv_dx, mu0[p] * e, lam0[p] * e, dt, volume, mass[p], A);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
          double dpos[3] = {(i - fx[0]) * dx, (j - fx[1]) * dx, (k - fx[2]) * dx};
          double weight = w[0][i] * w[1][j] * w[2][k];
          int64_t node = ((b[0] + i) * G + (b[1] + j)) * G + (b[2] + k);
          for (int r = 0; r < 3; ++r) {
            double val = weight * (v[3 * p + r] * mass[p] + (A[3 * r] * dpos[0] + A[3 * r + 1] * dpos[1] + A[3 * r + 2] * dpos[2]));
#pragma omp atomic
            grid_velocity[3 * node + r] += val;
          }
          double wm = weight * mass[p];
#pragma omp atomic
          grid_mass[node] += wm;
        }
  }
  return n_oob;
}

/* three_d/grid_op.py:5-47 */
void oracle_grid_op_3d(int64_t res, double dx, double dt, double gravity, double* grid_velocity, const double* grid_mass) {
  const int64_t G = res + 1;
  const double v_allowed = dx * 0.9 / dt;
  const int64_t boundary = 1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < G; ++i)
    for (int64_t j = 0; j < G; ++j)
      for (int64_t k = 0; k < G; ++k) {
        int64_t node = (i * G + j) * G + k;
        double* gv = grid_velocity + 3 * node;
        if (grid_mass[node] > 0) {
          for (int r = 0; r < 3; ++r) gv[r] /= grid_mass[node];
          gv[1] += dt * gravity;
          for (int r = 0; r < 3; ++r) gv[r] = gv[r] < -v_allowed ? -v_allowed : (gv[r] > v_allowed ? v_allowed : gv[r]);
        }
        int64_t I[3] = {i, j, k};
        for (int d = 0; d < 3; ++d) {
          if (I[d] < boundary) gv[d] = 0;
          if (I[d] >= res - boundary) gv[d] = 0;
        }
      }
}

/* three_d/g2p.py:9-59 (neo_hookean branch) */
int64_t oracle_g2p_3d(int64_t n, int64_t G, double inv_dx, double dt, const double* grid_velocity, double* x, double* v,
                      double* F, double* C) {
  int64_t n_oob = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_oob)
  for (int64_t p = 0; p < n; ++p) {
    int64_t b[3];
    double fx[3], w[3][3];
    int bad = 0;
    for (int d = 0; d < 3; ++d) {
      base_fx_w(x[3 * p + d], inv_dx, &b[d], &fx[d], w[d]);
      if (b[d] < 0 || b[d] + 2 >= G) bad = 1;
    }
    if (bad) { n_oob += 1; continue; }
    double nv[3] = {0, 0, 0}, nC[9] = {0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
          double dpos[3] = {i - fx[0], j - fx[1], k - fx[2]};
          const double* gv = grid_velocity + 3 * (((b[0] + i) * G + (b[1] + j)) * G + (b[2] + k));
          double weight = w[0][i] * w[1][j] * w[2][k];
          for (int r = 0; r < 3; ++r) {
            double wg = weight * gv[r];
            nv[r] += wg;
            for (int c = 0; c < 3; ++c) nC[3 * r + c] += 4 * inv_dx * (wg * dpos[c]);
          }
        }
    for (int r = 0; r < 3; ++r) { v[3 * p + r] = nv[r]; x[3 * p + r] += dt * nv[r]; }
    double Fo[9], Fn[9];
    memcpy(Fo, F + 9 * p, sizeof(Fo));
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += ((r == k ? 1.0 : 0.0) + dt * nC[3 * r + k]) * Fo[3 * k + c];
        Fn[3 * r + c] = acc;
      }
    memcpy(F + 9 * p, Fn, sizeof(Fn));
    memcpy(C + 9 * p, nC, sizeof(nC));
  }
  return n_oob;
}

/* solvers/mpm/mls_mpm.py:40-79: zeroed grids, p2g, grid_op, g2p.  `grid_velocity`
 * (G^3*3) and `grid_mass` (G^3) are caller-provided scratch. */
int64_t oracle_substep_3d(int64_t n, int64_t res, double inv_dx, double hardening, double dx, double dt, double volume,
                          double gravity, double* x, const double* mass, const double* mu0, const double* lam0,
                          double* v, double* F, double* C, double* grid_velocity, double* grid_mass) {
  const int64_t G = res + 1;
  memset(grid_velocity, 0, sizeof(double) * 3 * G * G * G);
  memset(grid_mass, 0, sizeof(double) * G * G * G);
  int64_t bad = oracle_p2g_3d(n, G, inv_dx, hardening, dx, dt, volume, grid_velocity, grid_mass, x, mass, mu0, lam0, v,
                              F, C, NULL, 0);
  oracle_grid_op_3d(res, dx, dt, gravity, grid_velocity, grid_mass);
  bad += oracle_g2p_3d(n, G, inv_dx, dt, grid_velocity, x, v, F, C);
  return bad;
}

/* ------------------------------- 2D ---------------------------------------- */
/* two_d/p2g.py:11-76 + utils.py:52-92 + linear_algebra.py:96-116 */
void oracle_p2g_2d(int64_t n, int64_t G, double inv_dx, double hardening, double mu_0, double lambda_0, double mass,
                   double dx, double dt, double volume, double* grid_velocity, double* grid_mass, const double* x,
                   const double* v, const double* F, const double* C, const double* Jp, int snow) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; ++p) {
    int64_t b[2];
    double fx[2], w[2][3];
    for (int d = 0; d < 2; ++d) base_fx_w(x[2 * p + d], inv_dx, &b[d], &fx[d], w[d]);
    double e = snow ? exp(hardening * (1.0 - Jp[p])) : hardening;
    double mu = mu_0 * e, lam = lambda_0 * e;
    const double* f = F + 4 * p;
    double J = f[0] * f[3] - f[1] * f[2];
    double xx = f[0] + f[3], yy = f[2] - f[1];
    double scale = 1.0 / (sqrt(xx * xx + yy * yy) + 1e-10);
    double c = xx * scale, s = yy * scale;
    double R[4] = {c, -s, s, c};
    double D_inv = 4 * inv_dx * inv_dx, A[4];
    for (int r = 0; r < 2; ++r)
      for (int cc = 0; cc < 2; ++cc) {
        double acc = 0;
        for (int k = 0; k < 2; ++k) acc += (2 * mu * (f[2 * r + k] - R[2 * r + k])) * f[2 * cc + k];
        A[2 * r + cc] = -(dt * volume) * (D_inv * (acc + lam * (J - 1) * J)) + mass * C[4 * p + 2 * r + cc];
      }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double dpos[2] = {(i - fx[0]) * dx, (j - fx[1]) * dx};
        double weight = w[0][i] * w[1][j];
        int64_t node = (b[0] + i) * G + (b[1] + j);
        for (int r = 0; r < 2; ++r) {
          double val = weight * (v[2 * p + r] * mass + (A[2 * r] * dpos[0] + A[2 * r + 1] * dpos[1]));
#pragma omp atomic
          grid_velocity[2 * node + r] += val;
        }
        double wm = weight * mass;
#pragma omp atomic
        grid_mass[node] += wm;
      }
  }
}

/* two_d/grid_op.py:5-24 */
void oracle_grid_op_2d(int64_t res, double dt, double gravity, double* grid_velocity, const double* grid_mass) {
  const int64_t G = res + 1;
  const double boundary = 0.05;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < G; ++i)
    for (int64_t j = 0; j < G; ++j) {
      int64_t node = i * G + j;
      if (grid_mass[node] > 0) {
        double* gv = grid_velocity + 2 * node;
        gv[0] /= grid_mass[node];
        gv[1] /= grid_mass[node];
        gv[1] += dt * gravity;
        double x = (double)i / (double)res, y = (double)j / (double)res;
        if (x < boundary || x > 1 - boundary || y > 1 - boundary) { gv[0] = 0.0; gv[1] = 0.0; }
        if (y < boundary) gv[1] = gv[1] > 0.0 ? gv[1] : 0.0;
      }
    }
}

/* two_d/g2p.py:5-47 (neo_hookean).  The SVD round trip U diag(sig) Vh^T is the
 * identity for det F > 0 and a rotation by 2 theta for det F < 0 (see below). */
void oracle_g2p_2d(int64_t n, int64_t G, double inv_dx, double dt, const double* grid_velocity, double* x, double* v,
                   double* F, double* C, double* Jp) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; ++p) {
    int64_t b[2];
    double fx[2], w[2][3];
    for (int d = 0; d < 2; ++d) base_fx_w(x[2 * p + d], inv_dx, &b[d], &fx[d], w[d]);
    double nv[2] = {0, 0}, nC[4] = {0, 0, 0, 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double dpos[2] = {i - fx[0], j - fx[1]};
        const double* gv = grid_velocity + 2 * ((b[0] + i) * G + (b[1] + j));
        double weight = w[0][i] * w[1][j];
        for (int r = 0; r < 2; ++r) {
          double wg = weight * gv[r];
          nv[r] += wg;
          for (int c = 0; c < 2; ++c) nC[2 * r + c] += 4 * inv_dx * (wg * dpos[c]);
        }
      }
    for (int r = 0; r < 2; ++r) { v[2 * p + r] = nv[r]; x[2 * p + r] += dt * nv[r]; }
    double Fo[4] = {F[4 * p], F[4 * p + 1], F[4 * p + 2], F[4 * p + 3]}, Fn[4];
    Fn[0] = (1 + dt * nC[0]) * Fo[0] + dt * nC[1] * Fo[2];
    Fn[1] = (1 + dt * nC[0]) * Fo[1] + dt * nC[1] * Fo[3];
    Fn[2] = dt * nC[2] * Fo[0] + (1 + dt * nC[3]) * Fo[2];
    Fn[3] = dt * nC[2] * Fo[1] + (1 + dt * nC[3]) * Fo[3];
    double old_J = Fn[0] * Fn[3] - Fn[1] * Fn[2];
    if (old_J < 0) {
      /* g2p.py:43 multiplies by Vh^T instead of Vh; for det F < 0 LAPACK's 2x2 Vh is the
       * proper rotation whose first row is the principal right singular vector, so
       * U S Vh^T = F (Vh^T)^2 = F Rot(2 theta) (measured: tests/golden/quirk2d.npz). */
      double m00 = Fn[0] * Fn[0] + Fn[2] * Fn[2], m01 = Fn[0] * Fn[1] + Fn[2] * Fn[3], m11 = Fn[1] * Fn[1] + Fn[3] * Fn[3];
      double dm = m00 - m11, om = 2 * m01, h = sqrt(dm * dm + om * om);
      double c2 = h > 0 ? dm / h : 1.0, s2 = h > 0 ? om / h : 0.0;
      double G0 = Fn[0] * c2 + Fn[1] * s2, G1 = -Fn[0] * s2 + Fn[1] * c2;
      double G2 = Fn[2] * c2 + Fn[3] * s2, G3 = -Fn[2] * s2 + Fn[3] * c2;
      Fn[0] = G0; Fn[1] = G1; Fn[2] = G2; Fn[3] = G3;
    }
    double det = old_J + 1e-10;
    double jp = Jp[p] * old_J / det;
    Jp[p] = jp < 0.6 ? 0.6 : (jp > 20.0 ? 20.0 : jp);
    memcpy(F + 4 * p, Fn, sizeof(Fn));
    memcpy(C + 4 * p, nC, sizeof(nC));
  }
}

void oracle_substep_2d(int64_t n, int64_t res, double inv_dx, double hardening, double mu_0, double lambda_0,
                       double mass, double dx, double dt, double volume, double gravity, double* x, double* v,
                       double* F, double* C, double* Jp, double* grid_velocity, double* grid_mass) {
  const int64_t G = res + 1;
  memset(grid_velocity, 0, sizeof(double) * 2 * G * G);
  memset(grid_mass, 0, sizeof(double) * G * G);
  oracle_p2g_2d(n, G, inv_dx, hardening, mu_0, lambda_0, mass, dx, dt, volume, grid_velocity, grid_mass, x, v, F, C, Jp, 0);
  oracle_grid_op_2d(res, dt, gravity, grid_velocity, grid_mass);
  oracle_g2p_2d(n, G, inv_dx, dt, grid_velocity, x, v, F, C, Jp);
}

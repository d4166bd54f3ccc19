// Host harness around the kernels' own per-particle device functions (femflow_b200/csrc):
// cell indexing, B-spline weights, the constitutive update in both precisions, the P2G payload and
// the separable G2P stencil sums are `__host__ __device__`, so the CPU test suite runs THE SAME
// SOURCE the GPU runs against LAPACK / the NumPy oracle (tests/test_kernel_math_host.py).  Test
// infrastructure only: nothing in the product links this file.
#include "../../femflow_b200/csrc/mpm_direct.cuh"
#include "../../femflow_b200/csrc/mpm_bin.cuh"
#include "../../femflow_b200/csrc/mpm_svd3.cuh"

using namespace ffmpm;

namespace {

DevCfg make_cfg(int res, int n_nodes, double inv_dx, double dx, double dt, double volume, double hardening, int model,
                int fp32_stress, int index_fp32) {
  DevCfg c{};
  c.dim = 3;
  c.model = model;
  for (int d = 0; d < 3; ++d) { c.n[d] = n_nodes; c.origin[d] = 0; c.res[d] = res; }
  c.inv_dx = inv_dx; c.dx = dx; c.dt = dt; c.volume = volume; c.gravity = 0.0; c.hardening = hardening;
  c.mass = c.mu0 = c.lam0 = 0.0;
  c.fp32_stress = fp32_stress;
  c.index_fp32 = index_fp32;
  c.own_lo = INT32_MIN; c.own_hi = INT32_MAX;
  return c;
}

template <typename T>
void prepare3(const DevCfg& cfg, long long n, const T* x, const T* v, const T* C, const T* F, const T* mass, const T* mu,
              const T* lam, const double* jp, int* base, T* fx, T* aff, T* mv, T* m, int* ok) {
  for (long long p = 0; p < n; ++p) {
    auto get = [&](int k) -> T {
      if (k < P2G_V) return x[3 * p + k];
      if (k < P2G_C) return v[3 * p + (k - P2G_V)];
      if (k < P2G_F) return C[9 * p + (k - P2G_C)];
      if (k < P2G_MASS) return F[9 * p + (k - P2G_F)];
      return k == P2G_MASS ? mass[p] : (k == P2G_MU ? mu[p] : lam[p]);
    };
    P2GParticle3<T> q = p2g_prepare3_from<T>(cfg, get, true, jp ? jp[p] : 1.0);
    ok[p] = q.ok ? 1 : 0;
    base[3 * p] = q.bx; base[3 * p + 1] = q.by; base[3 * p + 2] = q.bz;
    fx[3 * p] = q.fx; fx[3 * p + 1] = q.fy; fx[3 * p + 2] = q.fz;
    if (!q.ok) continue;
    const T a[9] = {q.a00, q.a01, q.a02, q.a10, q.a11, q.a12, q.a20, q.a21, q.a22};
    for (int e = 0; e < 9; ++e) aff[9 * p + e] = a[e];
    mv[3 * p] = q.mvx; mv[3 * p + 1] = q.mvy; mv[3 * p + 2] = q.mvz;
    m[p] = q.m;
  }
}

template <typename T>
struct V3 { T x, y, z; };

template <typename T>
void accumulate3(long long n, const T* f, const T* gv, T* v, T* c) {
  for (long long p = 0; p < n; ++p) {
    const T* g = gv + 81 * p;   // [3][3][3][3]
    T o[12];
    g2p_accumulate3<T>([&](int i, int j, int k) { const T* q = g + ((i * 3 + j) * 3 + k) * 3; return V3<T>{q[0], q[1], q[2]}; },
                       f[3 * p], f[3 * p + 1], f[3 * p + 2], o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9],
                       o[10], o[11]);
    for (int e = 0; e < 3; ++e) v[3 * p + e] = o[e];
    for (int e = 0; e < 9; ++e) c[9 * p + e] = o[3 + e];
  }
}

}  // namespace

// Grid update of every node (three_d/grid_op.py:5-67 through the kernels' grid_op3_node): a slab of `n0`
// node planes at global x origin `origin0` of a (res0, res1, res2) grid, halo partial sums of the first
// `planes_lo` / last `planes_hi` planes added while loading, `n_col` plane colliders (normals as
// ffmpm_set_colliders stores them).  In place on grid[n0*n1*n2][4].
template <typename T>
static void grid_op3_host(const int* res, const int* n, int origin0, double dx, double dt, double gravity, T* grid, const T* halo_lo,
                          int planes_lo, const T* halo_hi, int planes_hi, int n_col, const double* col_point, const double* col_normal) {
  DevCfg c{};
  c.dim = 3;
  for (int d = 0; d < 3; ++d) { c.res[d] = res[d]; c.n[d] = n[d]; c.origin[d] = 0; }
  c.origin[0] = origin0;
  c.dx = dx; c.inv_dx = 1.0 / dx; c.dt = dt; c.gravity = gravity;
  Colliders col{};
  col.count = n_col;
  for (int k = 0; k < n_col; ++k)
    for (int d = 0; d < 3; ++d) { col.point[k][d] = col_point[3 * k + d]; col.normal[k][d] = col_normal[3 * k + d]; }
  const long long plane = (long long)n[1] * n[2], nodes = plane * n[0];
  for (long long node = 0; node < nodes; ++node) {
    const int k = (int)(node % n[2]);
    const long long r = node / n[2];
    grid_op3_node<T>(c, grid, nodes, node, (int)(r / n[1]), (int)(r % n[1]), k, halo_lo, planes_lo * plane, halo_hi,
                     planes_hi * plane, col);
  }
}


extern "C" {

void km_prepare3_f32(int res, int n_nodes, double inv_dx, double dx, double dt, double volume, double hardening, int model,
                     int fp32_stress, int index_fp32, long long n, const float* x, const float* v, const float* C,
                     const float* F, const float* mass, const float* mu, const float* lam, const double* jp, int* base,
                     float* fx, float* aff, float* mv, float* m, int* ok) {
  prepare3<float>(make_cfg(res, n_nodes, inv_dx, dx, dt, volume, hardening, model, fp32_stress, index_fp32), n, x, v, C, F,
                  mass, mu, lam, jp, base, fx, aff, mv, m, ok);
}

void km_prepare3_f64(int res, int n_nodes, double inv_dx, double dx, double dt, double volume, double hardening, int model,
                     int fp32_stress, int index_fp32, long long n, const double* x, const double* v, const double* C,
                     const double* F, const double* mass, const double* mu, const double* lam, const double* jp, int* base,
                     double* fx, double* aff, double* mv, double* m, int* ok) {
  prepare3<double>(make_cfg(res, n_nodes, inv_dx, dx, dt, volume, hardening, model, fp32_stress, index_fp32), n, x, v, C,
                   F, mass, mu, lam, jp, base, fx, aff, mv, m, ok);
}

void km_g2p_accumulate3_f32(long long n, const float* f, const float* gv, float* v, float* c) { accumulate3<float>(n, f, gv, v, c); }
void km_g2p_accumulate3_f64(long long n, const double* f, const double* gv, double* v, double* c) { accumulate3<double>(n, f, gv, v, c); }

// 1 when the fp32 perturbation series accepted the strain (else the caller's fp64 path would run)
int km_affine3_f32(const float* F, const float* C, float mu, float lam, float mass, float k, float* A) {
  Mat3<float> f{F[0], F[1], F[2], F[3], F[4], F[5], F[6], F[7], F[8]}, c{C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], C[8]}, a;
  const bool done = fixed_corotated_affine3_f32(f, c, mu, lam, mass, k, a);
  if (done) { const float o[9] = {a.a00, a.a01, a.a02, a.a10, a.a11, a.a12, a.a20, a.a21, a.a22}; for (int e = 0; e < 9; ++e) A[e] = o[e]; }
  return done ? 1 : 0;
}

void km_polar3(long long n, const double* F, double* R, double* det) {
  for (long long p = 0; p < n; ++p) {
    const double* f = F + 9 * p;
    Mat3<double> m{f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8]};
    Mat3<double> r = polar_rotation3(m, det[p]);
    const double o[9] = {r.a00, r.a01, r.a02, r.a10, r.a11, r.a12, r.a20, r.a21, r.a22};
    for (int e = 0; e < 9; ++e) R[9 * p + e] = o[e];
  }
}

void km_affine2(long long n, const double* F, const double* C, double mu, double lam, double mass, double k, double* A) {
  for (long long p = 0; p < n; ++p) {
    Mat2<double> f{F[4 * p], F[4 * p + 1], F[4 * p + 2], F[4 * p + 3]}, c{C[4 * p], C[4 * p + 1], C[4 * p + 2], C[4 * p + 3]};
    Mat2<double> a = fixed_corotated_affine2(f, c, mu, lam, mass, k);
    A[4 * p] = a.a00; A[4 * p + 1] = a.a01; A[4 * p + 2] = a.a10; A[4 * p + 3] = a.a11;
  }
}

// fp32 closed form of the 2D stress (fp32 build): 1 per particle where it accepted the strain
void km_affine2_f32(long long n, const float* F, const float* C, float mu, float lam, float mass, float k, float* A, int* took) {
  for (long long p = 0; p < n; ++p) {
    Mat2<float> f{F[4 * p], F[4 * p + 1], F[4 * p + 2], F[4 * p + 3]}, c{C[4 * p], C[4 * p + 1], C[4 * p + 2], C[4 * p + 3]}, a{};
    took[p] = fixed_corotated_affine2_f32(f, c, mu, lam, mass, k, a) ? 1 : 0;
    A[4 * p] = a.a00; A[4 * p + 1] = a.a01; A[4 * p + 2] = a.a10; A[4 * p + 3] = a.a11;
  }
}

void km_svd_roundtrip2(long long n, const double* F, int snow, double* G, double* det) {
  for (long long p = 0; p < n; ++p) {
    Mat2<double> f{F[4 * p], F[4 * p + 1], F[4 * p + 2], F[4 * p + 3]};
    Mat2<double> g = svd_roundtrip2(f, snow != 0, det[p]);
    G[4 * p] = g.a00; G[4 * p + 1] = g.a01; G[4 * p + 2] = g.a10; G[4 * p + 3] = g.a11;
  }
}

void km_base_fx_f32(double inv_dx, int index_fp32, long long n, const float* x, int* base, float* fx) {
  DevCfg c{}; c.inv_dx = inv_dx; c.index_fp32 = index_fp32;
  for (long long p = 0; p < n; ++p) base_fx<float>(x[p], c, base[p], fx[p]);
}

void km_base_fx_f64(double inv_dx, long long n, const double* x, int* base, double* fx) {
  DevCfg c{}; c.inv_dx = inv_dx; c.index_fp32 = 0;
  for (long long p = 0; p < n; ++p) base_fx<double>(x[p], c, base[p], fx[p]);
}

// The left-form series h(G) = G p(G) of mpm_math.cuh at a GIVEN tier (the warp-autonomous P2G evaluates every
// particle of a warp at the tier of its most strained one): G as (xx, xy, xz, yy, yz, zz).
void km_left_stress_h(int tier, long long n, const float* G, float* H) {
  for (long long p = 0; p < n; ++p) {
    const float* g = G + 6 * p;
    const Sym3f s{g[0], g[1], g[2], g[3], g[4], g[5]};
    const Sym3f h = tier == 0 ? left_stress_h<0>(s) : tier == 1 ? left_stress_h<1>(s) : tier == 2 ? left_stress_h<2>(s) : left_stress_h<3>(s);
    float* o = H + 6 * p;
    o[0] = h.xx; o[1] = h.xy; o[2] = h.xz; o[3] = h.yy; o[4] = h.yz; o[5] = h.zz;
  }
}

// E = F - I, G = F F^T - I and ||G||_F^2 as the kernels form them, and the tier the norm selects.
void km_left_strain(long long n, const float* F, float* G, float* r2, int* tier, float* jm1) {
  for (long long p = 0; p < n; ++p) {
    const float* f = F + 9 * p;
    Mat3<float> E;
    Sym3f g;
    left_strain3(Mat3<float>{f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8]}, E, g, r2[p]);
    float* o = G + 6 * p;
    o[0] = g.xx; o[1] = g.xy; o[2] = g.xz; o[3] = g.yy; o[4] = g.yz; o[5] = g.zz;
    tier[p] = stress_tier_of(r2[p]);
    jm1[p] = jm1_of(E);
  }
}

float km_stress_coef(int tier, int i) { return stress_coef(tier, i); }
float km_stress_tier_r(int tier) { return stress_tier_r(tier); }

void km_grid_op3_f32(const int* res, const int* n, int origin0, double dx, double dt, double gravity, float* grid, const float* halo_lo,
                     int planes_lo, const float* halo_hi, int planes_hi, int n_col, const double* cp, const double* cn) {
  grid_op3_host<float>(res, n, origin0, dx, dt, gravity, grid, halo_lo, planes_lo, halo_hi, planes_hi, n_col, cp, cn);
}
void km_grid_op3_f64(const int* res, const int* n, int origin0, double dx, double dt, double gravity, double* grid, const double* halo_lo,
                     int planes_lo, const double* halo_hi, int planes_hi, int n_col, const double* cp, const double* cn) {
  grid_op3_host<double>(res, n, origin0, dx, dt, gravity, grid, halo_lo, planes_lo, halo_hi, planes_hi, n_col, cp, cn);
}

// Tile-major bin keys (mpm_bin.cuh: bin_key_of) of fp32 positions on an n-node grid at x origin `origin0`;
// also the global base cell along x that the slab migration logic reads.
int km_bin_keys_f32(int dim, const int* n, int origin0, double inv_dx, int index_fp32, long long count, const float* x, int* keys,
                    int* base_x) {
  DevCfg c{};
  c.dim = dim;
  for (int d = 0; d < 3; ++d) { c.n[d] = n[d]; c.origin[d] = 0; }
  c.origin[0] = origin0;
  c.inv_dx = inv_dx; c.index_fp32 = index_fp32;
  BinBuffers B{};
  bin_geometry(dim, n, B.tiles, B.n_tiles, B.n_cells);
  for (long long p = 0; p < count; ++p)
    keys[p] = bin_key_of<float>(c, B, x[3 * p], x[3 * p + 1], x[3 * p + 2], base_x + p);
  return B.n_cells;
}

// 3x3 SVD with LAPACK's singular-vector signs (mpm_svd3.cuh): row-major A -> U, sig, Vh
void km_svd3_lapack(long long n, const double* A, double* U, double* S, double* Vh, int* ok) {
  for (long long p = 0; p < n; ++p) {
    double a[3][3], u[3][3], vt[3][3];
    for (int i = 0; i < 9; ++i) a[i / 3][i % 3] = A[9 * p + i];
    ok[p] = svd3_lapack(a, u, S + 3 * p, vt) ? 1 : 0;
    for (int i = 0; i < 9; ++i) { U[9 * p + i] = u[i / 3][i % 3]; Vh[9 * p + i] = vt[i / 3][i % 3]; }
  }
}

// three_d/g2p.py:46-59, snow: F (old), C (new), Jp -> F, Jp
void km_snow_return_map3(long long n, double dt, const double* F, const double* Cm, const double* jp, double* Fout, double* jp_out) {
  for (long long p = 0; p < n; ++p) snow_return_map3(F + 9 * p, Cm + 9 * p, dt, jp[p], Fout + 9 * p, jp_out[p]);
}

}  // extern "C"

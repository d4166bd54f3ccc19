// Shared device-side types for the MPM kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mpm_math.cuh"

namespace ffmpm {

// Device copy of FfMpmConfig (passed by value as a kernel argument).
struct DevCfg {
  int dim, model;
  int n[3], origin[3], res[3];
  double inv_dx, dx, dt, volume, gravity, hardening;
  double mass, mu0, lam0;
  int fp32_stress;   // fp32 build: evaluate the stress in perturbation form in fp32 when the strain allows
  int index_fp32;      // inv_dx is a power of two: cell indexing is exact in fp32 (see base_fx)
  int own_lo, own_hi;  // slabs: GLOBAL base cells [own_lo, own_hi) along x owned by this rank (G2P counts leavers)
  int own_slack;       // ... and, separately, the leavers that are more than own_slack cells outside it ("urgent")
};

// Plane colliders of three_d/grid_op.py:50-67 (normals already shifted by 1/|normal|, the
// reference's quirk of adding the scalar to every component).
struct Colliders {
  int count;
  double point[8][3];
  double normal[8][3];
};

template <typename T>
struct StateView {
  T* x; T* v; T* C; T* F; T* Jp; T* mass; T* mu0; T* lam0;
  int* id;
  unsigned char* material;   // row of mat_table per particle (nullptr: row 0)
  const T* mat_table;        // [3][MAT_ROWS]: mass, mu0, lam0 rows; nullptr: planes or config scalars
  long long stride;
};
constexpr int MAT_ROWS = 256;

// Where a kernel takes a particle's (mass, mu_0, lambda_0) from (reference Particle, particle.py:7-13).
enum MatMode { MAT_CFG = 0, MAT_PLANES = 1, MAT_TABLE = 2 };
template <typename T>
__host__ __device__ __forceinline__ int mat_mode_of(const StateView<T>& s) {
  if (s.mat_table) return MAT_TABLE;
  return (s.mass && s.mu0 && s.lam0) ? MAT_PLANES : MAT_CFG;
}

// Sticky error record in the workspace: [0] = code, [1] = number of out-of-grid particles.
struct ErrRec {
  unsigned long long code;
  unsigned long long n_oob;
};

// base = trunc(x*inv_dx - 0.5) (C cast == numpy astype(int64): toward zero, quirk 1)
// and fx = x*inv_dx - base.  Evaluated in fp64 from the stored position so the binning is
// bit-exact with the fp64 reference for any grid resolution (quirk 11) -- except when
// inv_dx is a power of two (cfg.index_fp32, e.g. all BASELINE grids): then x*inv_dx,
// the subtraction of 0.5, the truncation and fx are all EXACT in fp32 for every position
// inside the grid, so the fp32 path returns the same bits at a fraction of the issue slots
// (6 fp64 conversion chains per particle in G2P, 3 in P2G).
// The exact-fp32 form (inv_dx a power of two): x*inv_dx, the subtraction of 0.5, the truncation and fx are exact.
FFMPM_HD void base_fx_f32(float xs, float inv_dx, int& base, float& fx) {
  const float s = xs * inv_dx;
  float t = s - 0.5f;
  t = fminf(fmaxf(t, -1.0e9f), 1.0e9f);   // keeps the int conversion defined; such values are out of grid anyway
  base = (int)t;
  fx = s - (float)base;
}

template <typename T>
FFMPM_HD void base_fx(T xs, const DevCfg& cfg, int& base, T& fx) {
  if (sizeof(T) == 4 && cfg.index_fp32) {
    float f;
    base_fx_f32((float)xs, (float)cfg.inv_dx, base, f);
    fx = (T)f;
    return;
  }
  double s = (double)xs * cfg.inv_dx;
  long long b = (long long)(s - 0.5);
  // clamp only to keep the int conversion defined for wild values; OOB is flagged by the caller
  if (b > 1000000000LL) b = 1000000000LL;
  if (b < -1000000000LL) b = -1000000000LL;
  base = (int)b;
  fx = (T)(s - (double)b);
}

// Quadratic B-spline weights (three_d/p2g.py:55).
template <typename T>
FFMPM_HD void bspline(T fx, T& w0, T& w1, T& w2) {
  T a = (T)1.5 - fx, b = fx - (T)1.0, c = fx - (T)0.5;
  w0 = (T)0.5 * a * a;
  w1 = (T)0.75 - b * b;
  w2 = (T)0.5 * c * c;
}

// One vector reduction per node: {mom_x, mom_y, mom_z, mass} += val.
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add4(double* addr, double a, double b, double c, double d) {
  atomicAdd(addr + 0, a);
  atomicAdd(addr + 1, b);
  atomicAdd(addr + 2, c);
  atomicAdd(addr + 3, d);
}

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

__device__ __forceinline__ float4 ld_node(const float* g) { return __ldg(reinterpret_cast<const float4*>(g)); }
__device__ __forceinline__ double4 ld_node(const double* g) {
  double2 a = __ldg(reinterpret_cast<const double2*>(g));
  double2 b = __ldg(reinterpret_cast<const double2*>(g) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

}  // namespace ffmpm

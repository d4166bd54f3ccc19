// Scene point generators on the device (reference femflow/simulation/mpm/primitives.py:8-76 and
// numerics/geometry.py:101-116): the lattice of res^3 points in [0,1]^3 (axis 0 fastest, coordinate =
// index / (res - 1)), filtered by an implicit function, and the axis-aligned box lattice of generate_cube_points.
// The reference walks the lattice point by point in Python (grid() + a loop over res^3 rows): the step in front of
// the hot path that takes minutes once scenes reach 10^7 particles.
//
// Order-preserving stream compaction in three passes: per-CTA counts of selected points, one-block exclusive scan of
// the counts, then every CTA re-evaluates its points and writes the selected ones at its offset in lattice order --
// the output is the reference's `g[inside > t]` row for row.  Coordinates are exact (one IEEE division per
// component); the implicit functions use the device's fp64 sin / cos, which may differ from the host libm in the
// last bit, so a lattice point within an ulp of the threshold could be classified differently (none is at the
// reference scene's parameters: tests/test_simulation.py pins the paper scene point for point).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ffmpm {

enum { IMPLICIT_GYROID = 0, IMPLICIT_DIAMOND = 1, IMPLICIT_PRIMITIVE = 2 };

// numerics/geometry.py:101-116: point i of the lattice, axis 0 fastest.
__device__ __forceinline__ void lattice_point(long long i, int res, double& x, double& y, double& z) {
  const int ix = (int)(i % res), iy = (int)((i / res) % res), iz = (int)(i / ((long long)res * res));
  const double den = res == 1 ? 1.0 : (double)(res - 1);
  x = (double)ix / den; y = (double)iy / den; z = (double)iz / den;
}

// primitives.py:8-43: the functions already subtract t, and generate_implicit_points compares the result with t again.
__device__ __forceinline__ bool implicit_selected(int kind, double k, double t, double x, double y, double z) {
  const double two_pi = (2.0 * 3.141592653589793) / k;
  const double ax = two_pi * x, ay = two_pi * y, az = two_pi * z;
  double f;
  if (kind == IMPLICIT_PRIMITIVE) {
    f = __dadd_rn(__dadd_rn(cos(ax), cos(ay)), cos(az));
  } else if (kind == IMPLICIT_GYROID) {
    f = __dadd_rn(__dadd_rn(__dmul_rn(sin(ax), cos(ay)), __dmul_rn(sin(ay), cos(az))), __dmul_rn(sin(az), cos(ax)));
  } else {
    const double sx = sin(ax), sy = sin(ay), sz = sin(az), cx = cos(ax), cy = cos(ay), cz = cos(az);
    f = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(sx, sy), sz), __dmul_rn(__dmul_rn(sx, cy), cz)),
                            __dmul_rn(__dmul_rn(cx, sy), cz)),
                  __dmul_rn(__dmul_rn(cx, cy), sz));
  }
  return __dadd_rn(f, -t) > t;
}

constexpr int SCENE_THREADS = 256;

__global__ void __launch_bounds__(SCENE_THREADS) implicit_count_kernel(int kind, double k, double t, int res, long long total,
                                                                      int* __restrict__ block_counts) {
  const long long i = (long long)blockIdx.x * SCENE_THREADS + threadIdx.x;
  bool sel = false;
  if (i < total) {
    double x, y, z;
    lattice_point(i, res, x, y, z);
    sel = implicit_selected(kind, k, t, x, y, z);
  }
  const int c = __syncthreads_count(sel);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of n_blocks ints in place (64-bit running total kept in out_total)
__global__ void __launch_bounds__(1024) scene_scan_kernel(int* __restrict__ counts, long long* __restrict__ offsets, int n_blocks,
                                                          long long* __restrict__ out_total) {
  __shared__ long long warp_sums[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int start = 0; start < n_blocks; start += 1024) {
    const int idx = start + threadIdx.x;
    const long long v = idx < n_blocks ? counts[idx] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long tmp = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += tmp;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      long long w = warp_sums[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long tmp = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= (unsigned)o) winc += tmp;
      }
      warp_sums[lane] = winc - w;
    }
    __syncthreads();
    const long long carry = carry_s;
    if (idx < n_blocks) offsets[idx] = carry + warp_sums[wid] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sums[wid] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_total = carry_s;
}

__global__ void __launch_bounds__(SCENE_THREADS) implicit_write_kernel(int kind, double k, double t, int res, long long total,
                                                                      const long long* __restrict__ offsets, double* __restrict__ out,
                                                                      long long capacity) {
  __shared__ int warp_base[SCENE_THREADS / 32];
  const long long i = (long long)blockIdx.x * SCENE_THREADS + threadIdx.x;
  double x = 0, y = 0, z = 0;
  bool sel = false;
  if (i < total) {
    lattice_point(i, res, x, y, z);
    sel = implicit_selected(kind, k, t, x, y, z);
  }
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, sel);
  if (lane == 0) warp_base[wid] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < SCENE_THREADS / 32; ++w) { const int c = warp_base[w]; warp_base[w] = run; run += c; }
  }
  __syncthreads();
  if (sel) {
    const long long dst = offsets[blockIdx.x] + warp_base[wid] + __popc(m & ((1u << lane) - 1u));
    if (dst < capacity) { out[3 * dst] = x; out[3 * dst + 1] = y; out[3 * dst + 2] = z; }
  }
}

// primitives.py:64-76: rows [z, y, x] with x fastest; np.linspace(a, b, res) = a + i * ((b - a) / (res - 1)), last = b.
__global__ void __launch_bounds__(SCENE_THREADS) cube_points_kernel(double x0, double x1, double y0, double y1, double z0, double z1,
                                                                   int res, double* __restrict__ out) {
  const long long total = (long long)res * res * res;
  const long long i = (long long)blockIdx.x * SCENE_THREADS + threadIdx.x;
  if (i >= total) return;
  const int ix = (int)(i % res), iy = (int)((i / res) % res), iz = (int)(i / ((long long)res * res));
  auto lin = [res](double a, double b, int idx) {
    if (res == 1) return a;
    if (idx == res - 1) return b;
    const double step = __ddiv_rn(__dadd_rn(b, -a), (double)(res - 1));
    return __dadd_rn(__dmul_rn((double)idx, step), a);
  };
  out[3 * i] = lin(z0, z1, iz);
  out[3 * i + 1] = lin(y0, y1, iy);
  out[3 * i + 2] = lin(x0, x1, ix);
}

}  // namespace ffmpm

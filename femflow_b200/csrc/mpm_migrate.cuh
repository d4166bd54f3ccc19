// Particle migration between neighbouring slabs, on the device (the reference is a single
// process, the ownership rule is its base cell three_d/p2g.py:50).
//
// A slab owns the GLOBAL base cells [own_lo, own_hi) along x (ffmpm_set_owned_range).  A migration round
//   pack    every live particle whose base cell left that range is copied into the outbox of its side
//           (at most `cap` per side; the rest stays one more round, the halo margin has the room) and its slot
//           becomes a hole; the holes below the new particle count are back-filled with the keepers that sit
//           above it -- O(leavers) traffic, the keepers stay where they are;
//   (the caller exchanges the two fixed-size outboxes with the neighbour ranks)
//   unpack  the particles of the two inboxes are appended behind the keepers.
// Nothing here needs the host: the counts live in the message headers and in a small device record that the
// caller reads ONCE per round (after unpack) to learn the new particle count.
//
// Message layout, scalars of the storage type T: row 0 is the header (element 0 = number of particles), rows
// 1 .. MIG_ROWS(+1 with Jp) hold one particle component per row, `cap` elements per row:
//   x3 v3 C9 F9 | mass mu0 lam0 (planes) or material row, 0, 0 (table / config material) | id (bit pattern) | [Jp]
#pragma once
#include "mpm_common.cuh"

namespace ffmpm {

constexpr int MIG_ROWS = 28;          // payload rows without Jp
constexpr int MIG_ID_ROW = 27;

struct MigRec {                       // device record of one round (int32 each)
  int out_lo, out_hi;                 // particles packed per side (<= cap)
  int in_lo, in_hi;                   // particles received per side
  int n_new;                          // particle count after the round
  int overflow;                       // leavers that did not fit an outbox (they stay), + capacity overflow on append
  int n_holes, n_lo_holes, n_movers;  // work counters of the back-fill
  int n_keep;                         // particle count after pack (before append)
  int pad[6];
};

template <typename T> __device__ __forceinline__ T mig_id_to_scalar(int id);
template <> __device__ __forceinline__ float mig_id_to_scalar<float>(int id) { return __int_as_float(id); }
template <> __device__ __forceinline__ double mig_id_to_scalar<double>(int id) { return (double)id; }
template <typename T> __device__ __forceinline__ int mig_scalar_to_id(T v);
template <> __device__ __forceinline__ int mig_scalar_to_id<float>(float v) { return __float_as_int(v); }
template <> __device__ __forceinline__ int mig_scalar_to_id<double>(double v) { return (int)v; }

template <typename T>
__device__ __forceinline__ T mig_component(const StateView<T>& s, long long p, int row) {
  const long long st = s.stride;
  if (row < 3) return s.x[row * st + p];
  if (row < 6) return s.v[(row - 3) * st + p];
  if (row < 15) return s.C[(row - 6) * st + p];
  if (row < 24) return s.F[(row - 15) * st + p];
  if (row == 24) return s.mass ? s.mass[p] : (s.material ? (T)s.material[p] : (T)0);
  if (row == 25) return s.mu0 ? s.mu0[p] : (T)0;
  if (row == 26) return s.lam0 ? s.lam0[p] : (T)0;
  if (row == MIG_ID_ROW) return mig_id_to_scalar<T>(s.id[p]);
  return s.Jp[p];
}

template <typename T>
__device__ __forceinline__ void mig_store_component(const StateView<T>& s, long long p, int row, T val) {
  const long long st = s.stride;
  if (row < 3) s.x[row * st + p] = val;
  else if (row < 6) s.v[(row - 3) * st + p] = val;
  else if (row < 15) s.C[(row - 6) * st + p] = val;
  else if (row < 24) s.F[(row - 15) * st + p] = val;
  else if (row == 24) { if (s.mass) s.mass[p] = val; else if (s.material) s.material[p] = (unsigned char)val; }
  else if (row == 25) { if (s.mu0) s.mu0[p] = val; }
  else if (row == 26) { if (s.lam0) s.lam0[p] = val; }
  else if (row == MIG_ID_ROW) s.id[p] = mig_scalar_to_id<T>(val);
  else s.Jp[p] = val;
}

// Pass 1: find the leavers (one read of the x plane), claim outbox slots warp-aggregated, copy them out, mark
// their slots as holes (id = -1) and list the holes.
template <typename T>
__global__ void __launch_bounds__(256) mig_pack_kernel(DevCfg cfg, StateView<T> s, long long n, T* __restrict__ out_lo,
                                                       T* __restrict__ out_hi, int cap, int rows, MigRec* rec,
                                                       int* __restrict__ holes) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int side = -1;
  if (p < n) {
    int gbx;
    T fx;
    base_fx(s.x[p], cfg, gbx, fx);
    if (gbx < cfg.own_lo && out_lo) side = 0;
    else if (gbx >= cfg.own_hi && out_hi) side = 1;
  }
  const unsigned lane = threadIdx.x & 31;
  int slot = -1;
#pragma unroll
  for (int sd = 0; sd < 2; ++sd) {
    const unsigned m = __ballot_sync(0xffffffffu, side == sd);
    if (m) {
      const int leader = __ffs(m) - 1;
      int base = 0;
      if ((int)lane == leader) base = atomicAdd(sd == 0 ? &rec->out_lo : &rec->out_hi, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (side == sd) slot = base + __popc(m & ((1u << lane) - 1u));
    }
  }
  if (side < 0) return;
  if (slot >= cap) {              // outbox full: this particle stays one more round
    atomicAdd(&rec->overflow, 1);
    return;
  }
  T* out = side == 0 ? out_lo : out_hi;
  for (int r = 0; r < rows; ++r) out[(long long)(1 + r) * cap + slot] = mig_component<T>(s, p, r);
  s.id[p] = -1;
  holes[atomicAdd(&rec->n_holes, 1)] = (int)p;
}

// Pass 2: with n_keep = n - packed known, list the holes below n_keep and the keepers at or above it.
template <typename T>
__global__ void __launch_bounds__(256) mig_match_kernel(StateView<T> s, long long n, int cap, MigRec* rec,
                                                        const int* __restrict__ holes, int* __restrict__ lo_holes,
                                                        int* __restrict__ movers) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int packed = min(rec->out_lo, cap) + min(rec->out_hi, cap);
  const long long n_keep = n - packed;
  if (t == 0) {
    rec->n_keep = (int)n_keep;
  }
  if (t < packed) {
    const int h = holes[t];
    if (h < n_keep) lo_holes[atomicAdd(&rec->n_lo_holes, 1)] = h;
    const long long j = n_keep + t;            // the tail [n_keep, n) has exactly `packed` slots
    if (j < n && s.id[j] != -1) movers[atomicAdd(&rec->n_movers, 1)] = (int)j;
  }
}

// Pass 3: move keeper k of the tail into hole k (the two lists have the same length by counting).
template <typename T>
__global__ void __launch_bounds__(256) mig_fill_kernel(StateView<T> s, int rows, const MigRec* rec,
                                                       const int* __restrict__ lo_holes, const int* __restrict__ movers) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rec->n_movers) return;
  const long long from = movers[t], to = lo_holes[t];
  for (int r = 0; r < rows; ++r) mig_store_component<T>(s, to, r, mig_component<T>(s, from, r));
}

// Header of an outbox = its particle count (as a scalar of the storage type; cap <= 2^22 keeps it exact in fp32).
template <typename T>
__global__ void mig_headers_kernel(T* out_lo, T* out_hi, int cap, const MigRec* rec) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (out_lo) out_lo[0] = (T)min(rec->out_lo, cap);
    if (out_hi) out_hi[0] = (T)min(rec->out_hi, cap);
  }
}

// Append the particles of both inboxes behind the keepers.
template <typename T>
__global__ void __launch_bounds__(256) mig_unpack_kernel(StateView<T> s, const T* __restrict__ in_lo, const T* __restrict__ in_hi,
                                                         int cap, int rows, MigRec* rec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c_lo = in_lo ? min((int)in_lo[0], cap) : 0, c_hi = in_hi ? min((int)in_hi[0], cap) : 0;
  const long long base = rec->n_keep;
  const long long room = s.stride - base;
  if (t == 0) {
    const long long take = min((long long)(c_lo + c_hi), room);
    rec->in_lo = c_lo;
    rec->in_hi = c_hi;
    rec->n_new = (int)(base + take);
    if (take < c_lo + c_hi) atomicAdd(&rec->overflow, (int)(c_lo + c_hi - take) + (1 << 24));   // capacity exceeded: flagged
    rec->out_lo = min(rec->out_lo, cap);
    rec->out_hi = min(rec->out_hi, cap);
  }
  if (t >= c_lo + c_hi || t >= room) return;
  const T* in = t < c_lo ? in_lo : in_hi;
  const int k = t < c_lo ? t : t - c_lo;
  for (int r = 0; r < rows; ++r) mig_store_component<T>(s, base + t, r, in[(long long)(1 + r) * cap + k]);
}

}  // namespace ffmpm

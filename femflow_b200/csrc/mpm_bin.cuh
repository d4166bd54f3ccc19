// Cell binning: a counting sort of the particles by the cell of their base node
// (three_d/p2g.py:50 `base_coord`), once per substep.
//
// Key layout (tile-major): the base cells are grouped into tiles of 4x4x4 (3D) or
// 8x8 (2D) cells = 64 cells, key = tile_id * 64 + cell_in_tile, so that one CTA of
// the tiled P2G/G2P kernels owns one contiguous run of binned particles.  Particles
// whose stencil would leave the grid get key = n_cells (a trailing bin) and are
// reported through the error record.
//
// Passes: (1) key + warp-aggregated histogram, which also yields each particle's
// rank inside its cell; (2) exclusive scan of the histogram (reduce / scan of block
// sums / downsweep); (3) active-tile list; (4) perm[offset[key] + rank] = p.
#pragma once
#include "mpm_common.cuh"

namespace ffmpm {

constexpr int TILE3 = 4;   // cells per tile edge, 3D
constexpr int TILE2 = 8;   // cells per tile edge, 2D
constexpr int TILE_CELLS = 64;
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct BinBuffers {
  int32_t* counters;     // [16]: 0 = n_active_tiles, 1 = p2g work counter, 2 = g2p work counter,
                         //       3 = particles whose base cell left the rank's owned x range (slabs),
                         //       4 = those of them more than own_slack cells outside it
  int32_t* cell_count;   // [n_cells + 2] histogram, bin n_cells = out-of-grid
  int32_t* cell_off;     // [n_cells + 2] exclusive scan of cell_count
  int32_t* block_sums;   // [n_scan_blocks + 1]
  int32_t* active_tiles; // [n_tiles]
  uint8_t* tile_flag;    // [n_tiles] 1 = the tile holds particles (rewritten by every binning)
  int32_t* node_tiles;   // [n_node_tiles] 4x4x4 node blocks P2G may have written (see node_tiles_kernel):
  int32_t* node_count;   //   the list + its length that the next binning fills (one of the two below)
  int32_t* node_tiles2[2];   // one list per grid of the ping-pong pair: the list also tells which blocks of
  int32_t* node_counts;      // that grid have to be cleared once the substep is over ([2] lengths)
  int ntile[3];          // node-tile grid: ceil(n / 4) per axis
  int n_node_tiles;
  int32_t* keys;         // [capacity]
  int32_t* rank;         // [capacity]
  int32_t* perm;         // [capacity]
  int tiles[3];
  int n_tiles;
  int n_cells;
  int n_scan_blocks;
  int64_t capacity;
};

inline void bin_geometry(int dim, const int* n, int* tiles, int& n_tiles, int& n_cells) {
  const int te = dim == 3 ? TILE3 : TILE2;
  n_tiles = 1;
  for (int d = 0; d < 3; ++d) {
    if (d < dim) {
      int nb = n[d] - 2;  // valid base cells per axis: 0 .. n-3
      tiles[d] = (nb + te - 1) / te;
    } else {
      tiles[d] = 1;
    }
    n_tiles *= tiles[d];
  }
  n_cells = n_tiles * TILE_CELLS;
}

// Node blocks: 4x4x4 nodes in 3D, 8x8 in 2D -- the edge of the base-cell tiles, so that a tile scatters into the
// blocks t and t+1 per axis.
inline int bin_node_tiles(int dim, const int* n, int* ntile) {
  const int edge = dim == 3 ? TILE3 : TILE2;
  int total = 1;
  for (int d = 0; d < 3; ++d) {
    ntile[d] = d < dim ? (n[d] + edge - 1) / edge : 1;
    total *= ntile[d];
  }
  return total;
}

inline int64_t bin_a256(int64_t v) { return (v + 255) / 256 * 256; }

inline int64_t bin_workspace_bytes(int dim, const int* n, int64_t capacity) {
  int tiles[3], n_tiles, n_cells;
  bin_geometry(dim, n, tiles, n_tiles, n_cells);
  int n_scan_blocks = (n_cells + 2 + SCAN_TILE - 1) / SCAN_TILE;
  int64_t b = 0;
  b += bin_a256(16 * 4);
  b += bin_a256((int64_t)(n_cells + 2) * 4) * 2;
  b += bin_a256((int64_t)(n_scan_blocks + 1) * 4);
  b += bin_a256((int64_t)n_tiles * 4);
  int ntile[3];
  b += bin_a256((int64_t)n_tiles);
  b += bin_a256((int64_t)bin_node_tiles(dim, n, ntile) * 4) * 2 + 256;
  b += bin_a256(capacity * 4) * 3;
  return b;
}

inline void bin_carve(BinBuffers& B, char* base, int dim, const int* n, int64_t capacity) {
  bin_geometry(dim, n, B.tiles, B.n_tiles, B.n_cells);
  B.n_scan_blocks = (B.n_cells + 2 + SCAN_TILE - 1) / SCAN_TILE;
  B.capacity = capacity;
  char* p = base;
  B.counters = (int32_t*)p; p += bin_a256(16 * 4);
  // counters and cell_count are contiguous so that one memset clears both
  B.cell_count = (int32_t*)p; p += bin_a256((int64_t)(B.n_cells + 2) * 4);
  B.cell_off = (int32_t*)p; p += bin_a256((int64_t)(B.n_cells + 2) * 4);
  B.block_sums = (int32_t*)p; p += bin_a256((int64_t)(B.n_scan_blocks + 1) * 4);
  B.active_tiles = (int32_t*)p; p += bin_a256((int64_t)B.n_tiles * 4);
  B.tile_flag = (uint8_t*)p; p += bin_a256((int64_t)B.n_tiles);
  B.n_node_tiles = bin_node_tiles(dim, n, B.ntile);
  B.node_tiles2[0] = (int32_t*)p; p += bin_a256((int64_t)B.n_node_tiles * 4);
  B.node_tiles2[1] = (int32_t*)p; p += bin_a256((int64_t)B.n_node_tiles * 4);
  B.node_counts = (int32_t*)p; p += 256;
  B.node_tiles = B.node_tiles2[0];
  B.node_count = B.node_counts;
  B.keys = (int32_t*)p; p += bin_a256(capacity * 4);
  B.rank = (int32_t*)p; p += bin_a256(capacity * 4);
  B.perm = (int32_t*)p; p += bin_a256(capacity * 4);
}

// Tile-major key of a LOCAL base cell.
FFMPM_HD int bin_key3(int bx, int by, int bz, int ty, int tz) {
  int tile = ((bx >> 2) * ty + (by >> 2)) * tz + (bz >> 2);
  return tile * TILE_CELLS + ((bx & 3) << 4) + ((by & 3) << 2) + (bz & 3);
}
FFMPM_HD int bin_key2(int bx, int by, int ty) {
  int tile = (bx >> 3) * ty + (by >> 3);
  return tile * TILE_CELLS + ((bx & 7) << 3) + (by & 7);
}

// Bin key of a particle position: tile-major id of its LOCAL base cell, or n_cells
// when the stencil would leave the grid (utils.py:138-150) / the position is NaN.
// IDX32: cfg.index_fp32 is known to hold (fp32 build, power-of-two inv_dx): the exact-fp32 indexing without the
// run-time choice (the fp64 code is then not generated at all).
template <typename T, bool IDX32 = false>
FFMPM_HD int bin_key_of(const DevCfg& cfg, const BinBuffers& B, T x0, T x1, T x2, int* base_x = nullptr) {
  const T xs[3] = {x0, x1, x2};
  int b[3] = {0, 0, 0};
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d < cfg.dim) {
      T fx;
      int g;
      if constexpr (IDX32 && sizeof(T) == 4) {
        float f;
        base_fx_f32((float)xs[d], (float)cfg.inv_dx, g, f);
        fx = (T)f;
      } else {
        base_fx(xs[d], cfg, g, fx);
      }
      b[d] = g - cfg.origin[d];
      if (d == 0 && base_x) *base_x = g;
      ok = ok && xs[d] == xs[d] && b[d] >= 0 && b[d] + 2 < cfg.n[d];
    }
  }
  if (!ok) return B.n_cells;
  return cfg.dim == 3 ? bin_key3(b[0], b[1], b[2], B.tiles[1], B.tiles[2]) : bin_key2(b[0], b[1], B.tiles[1]);
}

// Warp-aggregated histogram: one atomic per distinct key per warp; the lanes of a group
// take consecutive ranks in lane order.  Must be called by all 32 lanes; key < 0 = idle.
// The same in two halves, so that the caller can put work between the atomic and the use of its return value
// (the round trip to L2 is several hundred cycles): bin_rank_issue posts the atomic, bin_rank_finish writes the rank.
struct BinRankTicket {
  unsigned peers;
  int base;      // valid in the group's leader lane
};
__device__ __forceinline__ BinRankTicket bin_rank_issue(const BinBuffers& B, int key) {
  BinRankTicket t;
  const unsigned lane = threadIdx.x & 31;
  t.peers = __match_any_sync(0xffffffffu, key);
  t.base = 0;
  if (key >= 0 && (int)lane == __ffs(t.peers) - 1) t.base = atomicAdd(&B.cell_count[key], __popc(t.peers));
  return t;
}
__device__ __forceinline__ void bin_rank_finish(const BinBuffers& B, const BinRankTicket& t, int key, long long slot) {
  const unsigned lane = threadIdx.x & 31;
  const int base = __shfl_sync(0xffffffffu, t.base, __ffs(t.peers) - 1);
  if (key >= 0) B.rank[slot] = base + __popc(t.peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ void bin_rank_warp(const BinBuffers& B, int key, long long slot) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0) {
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(&B.cell_count[key], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    B.rank[slot] = base + __popc(peers & ((1u << lane) - 1u));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bin_count_kernel(DevCfg cfg, StateView<T> s, long long n, BinBuffers B) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int key = -1;
  if (p < n) {
    const long long st = s.stride;
    key = bin_key_of<T>(cfg, B, s.x[p], cfg.dim > 1 ? s.x[st + p] : (T)0, cfg.dim > 2 ? s.x[2 * st + p] : (T)0);
    B.keys[p] = key;
  }
  bin_rank_warp(B, key, p);
}

// ---- exclusive scan of cell_count[0 .. m) into cell_off, m = n_cells + 2 ----
__device__ __forceinline__ int block_exclusive_scan(int v, int* total_out) {
  __shared__ int warp_sums[32];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int nw = (blockDim.x + 31) >> 5;
    int w = (int)lane < nw ? warp_sums[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (unsigned)o) winc += t;
    }
    warp_sums[lane] = winc - w;  // exclusive
    if ((int)lane == nw - 1 && total_out) *total_out = winc;
  }
  __syncthreads();
  int r = warp_sums[wid] + inc - v;
  __syncthreads();
  return r;
}

// A thread owns SCAN_ITEMS = 8 consecutive counts = two 16-byte vectors (the arrays are 256-byte
// aligned): scalar loads at a 32-byte thread stride made these kernels L1-bound (72 % l1tex, 1.8 TB/s).
__device__ __forceinline__ void scan_load8(const int32_t* __restrict__ in, int base, int m, int (&v)[SCAN_ITEMS]) {
  if (base + SCAN_ITEMS <= m) {
    const int4 a = *reinterpret_cast<const int4*>(in + base), b = *reinterpret_cast<const int4*>(in + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = base + i < m ? in[base + i] : 0;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int32_t* __restrict__ in, int m,
                                                                   int32_t* __restrict__ block_sums) {
  __shared__ int total;
  static_assert(SCAN_ITEMS == 8, "two int4 per thread");
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  scan_load8(in, base, m, v);
  int sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) sum += v[i];
  block_exclusive_scan(sum, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int32_t* __restrict__ block_sums, int nb) {
  __shared__ int total;
  int carry = 0;
  for (int start = 0; start < nb; start += 1024) {
    int i = start + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int ex = block_exclusive_scan(v, &total);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[nb] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_downsweep_kernel(const int32_t* __restrict__ in, int m,
                                                                      const int32_t* __restrict__ block_sums,
                                                                      int32_t* __restrict__ out) {
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  scan_load8(in, base, m, v);
  int sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) sum += v[i];
  int ex = block_exclusive_scan(sum, nullptr) + block_sums[blockIdx.x];
  if (base + SCAN_ITEMS <= m) {
    int4 a, b;
    a.x = ex; a.y = a.x + v[0]; a.z = a.y + v[1]; a.w = a.z + v[2];
    b.x = a.w + v[3]; b.y = b.x + v[4]; b.z = b.y + v[5]; b.w = b.z + v[6];
    *reinterpret_cast<int4*>(out + base) = a;
    *reinterpret_cast<int4*>(out + base + 4) = b;
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      if (base + i < m) out[base + i] = ex;
      ex += v[i];
    }
  }
}

// Compact list of tiles that hold at least one particle (nearly ordered: one atomic per warp).
__global__ void __launch_bounds__(256) active_tiles_kernel(BinBuffers B) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = false;
  if (t < B.n_tiles) {
    act = B.cell_off[(t + 1) * TILE_CELLS] > B.cell_off[t * TILE_CELLS];
    B.tile_flag[t] = act ? 1 : 0;
  }
  unsigned m = __ballot_sync(0xffffffffu, act);
  if (m) {
    unsigned lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(&B.counters[0], __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (act) B.active_tiles[base + __popc(m & ((1u << lane) - 1u))] = t;
  }
}

// 3D: list of the 4x4x4 NODE blocks that the P2G of the binned particles can have written.  A tile of
// base cells 4t .. 4t+3 scatters to nodes 4t .. 4t+5, i.e. into node blocks t and t+1 per axis, so node
// block T is live iff one of the base-cell tiles T - {0,1}^3 holds particles.  The grid update then
// visits these blocks only (the grid is 8x larger than the occupied part in the 16M-particle block).
__global__ void __launch_bounds__(256) node_tiles_kernel(BinBuffers B) {
  const int T = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  if (T < B.n_node_tiles) {
    const int c = T % B.ntile[2], b = (T / B.ntile[2]) % B.ntile[1], a = T / (B.ntile[2] * B.ntile[1]);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const int ta = a - (d >> 2), tb = b - ((d >> 1) & 1), tc = c - (d & 1);
      if (ta >= 0 && tb >= 0 && tc >= 0 && ta < B.tiles[0] && tb < B.tiles[1] && tc < B.tiles[2])
        live = live || B.tile_flag[(ta * B.tiles[1] + tb) * B.tiles[2] + tc] != 0;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, live);
  if (m) {
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(B.node_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) B.node_tiles[base + __popc(m & ((1u << lane) - 1u))] = T;
  }
}

__global__ void node_tiles2_kernel(BinBuffers B);   // mpm_2d.cuh

__global__ void __launch_bounds__(256) bin_scatter_kernel(long long n, BinBuffers B, ErrRec* err) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int key = B.keys[p];
  B.perm[B.cell_off[key] + B.rank[p]] = (int)p;
  if (key == B.n_cells) atomicAdd(&err->n_oob, 1ULL);
}

// Clears the histogram (and nothing else): issued before a kernel that pre-bins.
inline void bin_clear_histogram(BinBuffers& B, cudaStream_t st) {
  cudaMemsetAsync(B.cell_count, 0, (size_t)(B.n_cells + 2) * sizeof(int32_t), st);
}

// Returns the number of kernel launches issued.  `prebinned`: keys / rank / histogram
// of the live buffer were already produced by the previous G2P.
template <typename T>
int bin_particles(const DevCfg& cfg, const StateView<T>& s, long long n, BinBuffers& B, ErrRec* err, bool prebinned,
                  cudaStream_t st) {
  const int m = B.n_cells + 2;
  unsigned pb = (unsigned)((n + 255) / 256);
  int launches = 5;
  if (prebinned) {
    cudaMemsetAsync(B.counters, 0, 16 * sizeof(int32_t), st);
  } else {
    // counters (16 ints, 256 B slot) + histogram in one clear
    cudaMemsetAsync(B.counters, 0, (size_t)((char*)(B.cell_count + m) - (char*)B.counters), st);
    bin_count_kernel<T><<<pb, 256, 0, st>>>(cfg, s, n, B);
    launches = 6;
  }
  scan_reduce_kernel<<<B.n_scan_blocks, SCAN_THREADS, 0, st>>>(B.cell_count, m, B.block_sums);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(B.block_sums, B.n_scan_blocks);
  scan_downsweep_kernel<<<B.n_scan_blocks, SCAN_THREADS, 0, st>>>(B.cell_count, m, B.block_sums, B.cell_off);
  // the histogram is consumed: clear it here (on the binning stream, underneath P2G) for the G2P that
  // pre-bins the next substep, instead of in front of that G2P on the main stream
  bin_clear_histogram(B, st);
  active_tiles_kernel<<<(B.n_tiles + 255) / 256, 256, 0, st>>>(B);
  cudaMemsetAsync(B.node_count, 0, sizeof(int32_t), st);
  if (cfg.dim == 3) node_tiles_kernel<<<(B.n_node_tiles + 255) / 256, 256, 0, st>>>(B);
  else node_tiles2_kernel<<<(B.n_node_tiles + 255) / 256, 256, 0, st>>>(B);
  ++launches;
  bin_scatter_kernel<<<pb, 256, 0, st>>>(n, B, err);
  return launches;
}

}  // namespace ffmpm

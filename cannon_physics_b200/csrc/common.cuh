// common.cuh — context, device buffers, launch helpers, exclusive scan and the hand-written radix sort
// (CUB-free by design: north_star asks for hand-written sort/scan primitives).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/cannon_cuda.h"

struct cannon_ctx {
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  std::vector<struct cannon_world*> pending;  // worlds with a cannon_world_step_async call that cannon_ctx_sync has not collected
};

#define CU_TRY(ctx, expr)                                                                         \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      char b__[512];                                                                              \
      snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      (ctx)->err = b__;                                                                           \
      return CANNON_E_CUDA;                                                                       \
    }                                                                                             \
  } while (0)

// growable device array
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n, bool keep = false, cudaStream_t s = nullptr) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = n + n / 4 + 16;
    T* q = nullptr;
    cudaError_t e = cudaMalloc((void**)&q, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return e;
      cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = q;
    cap = ncap;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// kernels launched by this library (reported through cannon_profile.kernel_launches)
static long long g_kernel_launches = 0;

// ---------------------------------------------------------------------------------------------------
// exclusive scan of int32 (n read from device memory so the whole step stays asynchronous)
// ---------------------------------------------------------------------------------------------------
#define SCAN_THREADS 512
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int warp_incl_scan(int v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (unsigned)o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int s_w[32];
  __shared__ int s_tot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int x = lane < nw ? s_w[lane] : 0;
    int xi = warp_incl_scan(x);
    if (lane < nw) s_w[lane] = xi - x;
    if (lane == nw - 1) s_tot = xi;
  }
  __syncthreads();
  int r = inc - v + s_w[wid];
  *total = s_tot;
  __syncthreads();
  return r;
}

struct ScanTmp {
  DBuf<int> tiles;
};

// Single-pass exclusive scan (decoupled look-back): every tile publishes its aggregate, then the inclusive prefix once the
// tiles before it are known; a tile only ever waits for tiles with a smaller ticket, and tickets are handed out in start
// order, so the wait always ends. One read and one write of the data, one launch (+ one memset of the tile states) instead
// of the three launches of the tile-sums / scan-of-sums / apply form - the step runs ~9 scans, this was 28 of its 73 launches.
#define SCAN_RESIDENT_BLOCKS 148  // one 512-thread block per SM is always resident on a B200
#define SCAN_FLAG_A 1ull  // aggregate of the tile
#define SCAN_FLAG_P 2ull  // inclusive prefix up to and including the tile
__device__ __forceinline__ void scan_publish(unsigned long long* p, unsigned long long flag, int v) {
  const unsigned long long w = (flag << 32) | (unsigned)v;
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long scan_peek(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ n_ptr, int n_fixed,
                                                               unsigned long long* __restrict__ state, int* __restrict__ ticket,
                                                               int* __restrict__ total_out) {
  __shared__ int s_tile, s_prefix;
  // tiles in start order: the block index when the whole grid is resident anyway (every block runs from the start, no block
  // can wait for one that has not been scheduled), a ticket otherwise
  if (gridDim.x <= SCAN_RESIDENT_BLOCKS) {
    if (threadIdx.x == 0) s_tile = blockIdx.x;
  } else if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int n = n_ptr ? *n_ptr : n_fixed;
  const int nTiles = n > 0 ? (n + SCAN_TILE - 1) / SCAN_TILE : 1;  // an empty input still reports its total through tile 0
  if (tile >= nTiles) return;
  const int base = tile * SCAN_TILE;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int i = base + threadIdx.x * SCAN_ITEMS + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  int tot;
  int ex = block_excl_scan(s, &tot);
  if (threadIdx.x == 0) {
    s_prefix = 0;
    scan_publish(&state[tile], tile == 0 ? SCAN_FLAG_P : SCAN_FLAG_A, tot);
  }
  if (tile > 0 && threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int prefix = 0;
    for (int t = tile - 1;; t -= 32) {  // lane k looks at tile t - k; tiles below 0 count as a zero prefix
      const int idx = t - lane;
      unsigned long long w;
      do {
        w = idx >= 0 ? scan_peek(&state[idx]) : (SCAN_FLAG_P << 32);
      } while (__any_sync(0xffffffffu, (w >> 32) == 0ull));
      const unsigned pm = __ballot_sync(0xffffffffu, (w >> 32) == SCAN_FLAG_P);
      const int upto = pm ? __ffs(pm) - 1 : 31;  // nearest tile that already knows its inclusive prefix
      int c = lane <= upto ? (int)(unsigned)w : 0;
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      prefix += c;
      if (pm) break;
    }
    if (lane == 0) {
      s_prefix = prefix;
      scan_publish(&state[tile], SCAN_FLAG_P, prefix + tot);
    }
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int i = base + threadIdx.x * SCAN_ITEMS + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
  if (tile == nTiles - 1 && threadIdx.x == 0 && total_out) *total_out = s_prefix + tot;
}

// out[i] = sum(in[0..i)), *total_out = sum(in[0..n)). n = *n_ptr (device) if n_ptr else n_fixed; cap bounds the grid.
// in == out is allowed (a tile reads its items before it writes them).
static inline cudaError_t scan_exclusive(const int* in, int* out, const int* n_ptr, int n_fixed, int cap, int* total_out, ScanTmp& tmp,
                                         cudaStream_t s) {
  int n_tiles = div_up(cap > 0 ? cap : 1, SCAN_TILE);
  cudaError_t e = tmp.tiles.reserve(2 * (size_t)n_tiles + 4);  // 64-bit tile states + the ticket counter
  if (e != cudaSuccess) return e;
  unsigned long long* state = (unsigned long long*)tmp.tiles.p;
  int* ticket = (int*)(state + n_tiles);
  if ((e = cudaMemsetAsync(state, 0, (size_t)n_tiles * 8 + 8, s)) != cudaSuccess) return e;
  k_scan_onepass<<<n_tiles, SCAN_THREADS, 0, s>>>(in, out, n_ptr, n_fixed, state, ticket, total_out);
  g_kernel_launches += 1;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// LSD radix sort of (u32 key, u32 value), 8 bits per pass, stable. n is host-known (body count).
//   pass = histogram per tile -> scan of [digit][tile] -> stable scatter with warp match ranking.
// Algorithmic traffic per pass and element: read key (hist) + read key,val + write key,val = 20 B.
// ---------------------------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ keys, int n, int shift, int n_tiles,
                                                        int* __restrict__ hist /* [256][n_tiles] */) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    int i = base + k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
                                                           int n_tiles, const int* __restrict__ hist_scanned) {
  __shared__ int s_base[256];            // global base of this tile for each digit (+ running count of earlier rounds)
  __shared__ int s_wcount[RS_THREADS / 32][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  s_base[threadIdx.x] = hist_scanned[threadIdx.x * n_tiles + blockIdx.x];
  const int base = blockIdx.x * RS_TILE;
  for (int k = 0; k < RS_ITEMS; k++) {
    for (int d = threadIdx.x; d < (RS_THREADS / 32) * 256; d += RS_THREADS) (&s_wcount[0][0])[d] = 0;
    __syncthreads();
    const int i = base + k * RS_THREADS + threadIdx.x;
    const bool valid = i < n;
    uint32_t key = valid ? keys[i] : 0xffffffffu;
    uint32_t val = valid ? vals[i] : 0u;
    const unsigned digit = valid ? ((key >> shift) & 255u) : 256u;  // 256 = "invalid" class
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) s_wcount[wid][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      int off = s_base[digit] + rank_in_warp;
      for (int w = 0; w < wid; w++) off += s_wcount[w][digit];
      keys_out[off] = key;
      vals_out[off] = val;
    }
    __syncthreads();
    {  // advance the running base by this round's totals
      int t = 0;
#pragma unroll
      for (int w = 0; w < RS_THREADS / 32; w++) t += s_wcount[w][threadIdx.x];
      s_base[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

struct SortTmp {
  DBuf<uint32_t> k2, v2;
  DBuf<int> hist;
  ScanTmp scan;
};

// sorts (keys, vals) of length n by the low `bits` bits of key; result ends up back in keys/vals.
static inline cudaError_t radix_sort_pairs(uint32_t* keys, uint32_t* vals, int n, int bits, SortTmp& t, cudaStream_t s) {
  if (n <= 1) return cudaSuccess;
  const int n_tiles = div_up(n, RS_TILE);
  cudaError_t e;
  if ((e = t.k2.reserve(n)) != cudaSuccess) return e;
  if ((e = t.v2.reserve(n)) != cudaSuccess) return e;
  if ((e = t.hist.reserve((size_t)256 * n_tiles)) != cudaSuccess) return e;
  const int passes = (bits + 7) / 8;  // an odd number of passes ends in the scratch buffers: copied back below
  uint32_t *ki = keys, *vi = vals, *ko = t.k2.p, *vo = t.v2.p;
  for (int p = 0; p < passes; p++) {
    const int shift = 8 * p;
    k_rs_hist<<<n_tiles, RS_THREADS, 0, s>>>(ki, n, shift, n_tiles, t.hist.p);
    if ((e = scan_exclusive(t.hist.p, t.hist.p, nullptr, 256 * n_tiles, 256 * n_tiles, nullptr, t.scan, s)) != cudaSuccess) return e;
    k_rs_scatter<<<n_tiles, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, n_tiles, t.hist.p);
    g_kernel_launches += 2;
    uint32_t* tk = ki; ki = ko; ko = tk;
    uint32_t* tv = vi; vi = vo; vo = tv;
  }
  if (passes & 1) {  // two small device copies are cheaper than a fourth pass (five launches)
    if ((e = cudaMemcpyAsync(keys, ki, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(vals, vi, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

// float -> u32 whose unsigned order equals the float order (-inf < ... < -0 = +0 handled as -0 < +0)
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// scan.cuh -- device-wide exclusive prefix sum (int32 output) over uint8 / int32 flags or counts.
// Three-phase reduce / scan-of-block-sums / downsweep; the block-sum array is scanned recursively.
#pragma once
#include "common.cuh"

namespace ahf {

constexpr int SC_THREADS = 512;
constexpr int SC_ITEMS   = 8;      // sc_load_items is written for 8
constexpr int SC_TILE    = SC_THREADS * SC_ITEMS;

template <typename T> __device__ __forceinline__ int sc_load(const T *a, uint64_t i) { return (int)a[i]; }
// SC_ITEMS consecutive inputs of one thread with vector loads when the thread's slice is whole and the array is aligned
__device__ __forceinline__ void sc_load_items(const uint8_t *a, uint64_t base, uint64_t n, int (&v)[8])
{
  if (base + 8 <= n && (reinterpret_cast<uintptr_t>(a) & 7) == 0) {
    const uint2 q = *reinterpret_cast<const uint2 *>(a + base);
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i] = (int)((q.x >> (8 * i)) & 255u); v[4 + i] = (int)((q.y >> (8 * i)) & 255u); }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (base + i < n) ? (int)a[base + i] : 0;
  }
}
__device__ __forceinline__ void sc_load_items(const int *a, uint64_t base, uint64_t n, int (&v)[8])
{
  if (base + 8 <= n && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
    const int4 q0 = *reinterpret_cast<const int4 *>(a + base), q1 = *reinterpret_cast<const int4 *>(a + base + 4);
    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (base + i < n) ? a[base + i] : 0;
  }
}
template <typename T> __device__ __forceinline__ void sc_load_items(const T *a, uint64_t base, uint64_t n, int (&v)[8])
{
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (base + i < n) ? (int)a[base + i] : 0;
}

__device__ __forceinline__ int block_exclusive_scan(int v)
{
  __shared__ int wsum[SC_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
  __syncthreads();
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = (lane < SC_THREADS / 32) ? wsum[lane] : 0, y = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int q = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += q; }
    if (lane < SC_THREADS / 32) wsum[lane] = y - x;
  }
  __syncthreads();
  return wsum[w] + inc - v;
}

template <typename T>
__global__ void __launch_bounds__(SC_THREADS) k_sc_reduce(const T *__restrict__ in, uint64_t n, int *__restrict__ bsum)
{
  __shared__ int red[SC_THREADS / 32];
  uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
  int s = 0, vv[SC_ITEMS];
  sc_load_items(in, base, n, vv);
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) s += vv[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < SC_THREADS / 32 ? red[threadIdx.x] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(SC_THREADS) k_sc_down(const T *in, uint64_t n, const int *__restrict__ boff,
                                                        int *out)
{
  uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
  int v[SC_ITEMS], s = 0;
  sc_load_items(in, base, n, v);
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) s += v[i];
  int ex = block_exclusive_scan(s) + boff[blockIdx.x];
  if (base + SC_ITEMS <= n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    int o[SC_ITEMS];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) { o[i] = ex; ex += v[i]; }
    *reinterpret_cast<int4 *>(out + base) = make_int4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<int4 *>(out + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
  } else {
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
  }
}

// single CTA exclusive scan of a short int array in place; writes the total to *total
static __global__ void __launch_bounds__(1024) k_sc_small(int *__restrict__ a, uint64_t m, int *__restrict__ total)
{
  __shared__ int wsum[32];
  const int      t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint64_t per = (m + 1023) / 1024;
  uint64_t       b = (uint64_t)t * per, e = b + per;
  if (b > m) b = m;
  if (e > m) e = m;
  int s = 0;
  for (uint64_t i = b; i < e; i++) s += a[i];
  int v = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += x; }
  if (lane == 31) wsum[w] = v;
  __syncthreads();
  if (w == 0) {
    int x = wsum[lane], y = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int q = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += q; }
    wsum[lane] = y - x;
    if (lane == 31 && total) *total = y;
  }
  __syncthreads();
  int run = wsum[w] + (v - s);
  for (uint64_t i = b; i < e; i++) { int x = a[i]; a[i] = run; run += x; }
}

// Single-pass scan (chained scan with decoupled look-back): one launch instead of reduce + scan of block sums + downsweep.  The path
// runs ~50 device-wide scans per pass, most of them over short arrays where launch latency is the cost.  Tiles are handed out by
// an atomic ticket, so a tile only ever waits for tiles that already run; per-tile state = (flag << 32 | value) in one 64-bit word
// (1: aggregate of the tile, 2: inclusive prefix up to and including the tile).  The state array lives in the context, starts
// zeroed and is zeroed again by the last tile to finish, so no memset precedes the launch.  In-place use (in == out) is fine: a
// tile reads its inputs before it writes its outputs and touches no other tile's data.
// building blocks of the chained scan, also used by kernels that fuse a compaction with their own work (mesh.cu k_relink_fused):
// sc_ticket: tile id in launch order; sc_lookback_prefix: called by the first warp with the tile's aggregate, returns the exclusive
// prefix (valid in lane 0) and publishes the inclusive one; sc_finish: the last tile to finish re-zeroes the state.
__device__ __forceinline__ unsigned sc_ticket(unsigned *__restrict__ ctr)
{
  __shared__ unsigned s_bid;
  if (threadIdx.x == 0) s_bid = atomicAdd(&ctr[0], 1u);
  __syncthreads();
  return s_bid;
}
__device__ __forceinline__ int sc_lookback_prefix(unsigned long long *__restrict__ state, unsigned bid, int agg, unsigned nblk, int *__restrict__ total)
{
  int prefix = 0;
  if (bid == 0) {
    if (threadIdx.x == 0) atomicExch(&state[0], (2ull << 32) | (unsigned)agg);
  } else {
    if (threadIdx.x == 0) atomicExch(&state[bid], (1ull << 32) | (unsigned)agg);
    long long j0 = (long long)bid - 1;                      // window [j0-31, j0], lane q looks at j0 - q
    for (;;) {
      const long long j = j0 - (long long)threadIdx.x;
      unsigned long long st = 2ull << 32;                   // before tile 0: prefix 0
      if (j >= 0) { do { st = *reinterpret_cast<volatile unsigned long long *>(&state[j]); } while ((st >> 32) == 0); }
      const unsigned isp = __ballot_sync(0xffffffffu, (st >> 32) == 2);
      const int      first = isp ? __ffs(isp) - 1 : 32;      // nearest predecessor that already has its inclusive prefix
      int add = ((int)threadIdx.x <= first) ? (int)(unsigned)st : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
      prefix += add;
      if (isp) break;
      j0 -= 32;
    }
    if (threadIdx.x == 0) atomicExch(&state[bid], (2ull << 32) | (unsigned)(prefix + agg));
  }
  if (threadIdx.x == 0 && bid == nblk - 1 && total) *total = prefix + agg;
  return prefix;
}
__device__ __forceinline__ void sc_finish(unsigned long long *__restrict__ state, unsigned *__restrict__ ctr, unsigned nblk)
{
  __shared__ int s_last;
  if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(&ctr[1], 1u) == nblk - 1) ? 1 : 0; }
  __syncthreads();
  if (s_last) {
    for (unsigned i = threadIdx.x; i < nblk; i += blockDim.x) state[i] = 0ull;
    if (threadIdx.x == 0) { ctr[0] = 0u; ctr[1] = 0u; }
  }
}

template <typename T, bool NZ = false>      // NZ: scan the flags (in[i] != 0) instead of the values
__global__ void __launch_bounds__(SC_THREADS) k_sc_lookback(const T *in, uint64_t n, int *out, unsigned long long *__restrict__ state,
                                                            unsigned *__restrict__ ctr, unsigned nblk, int *__restrict__ total)
{
  __shared__ int s_prefix, s_agg;
  const unsigned bid = sc_ticket(ctr);
  const uint64_t base = (uint64_t)bid * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
  int v[SC_ITEMS], s = 0;
  sc_load_items(in, base, n, v);
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) { if (NZ) v[i] = v[i] != 0; s += v[i]; }
  int ex = block_exclusive_scan(s);
  if (threadIdx.x == SC_THREADS - 1) s_agg = ex + s;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int prefix = sc_lookback_prefix(state, bid, s_agg, nblk, total);
    if (threadIdx.x == 0) s_prefix = prefix;
  }
  __syncthreads();
  ex += s_prefix;
  if (base + SC_ITEMS <= n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    int o[SC_ITEMS];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) { o[i] = ex; ex += v[i]; }
    *reinterpret_cast<int4 *>(out + base) = make_int4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<int4 *>(out + base + 4) = make_int4(o[4], o[5], o[6], o[7]);
  } else {
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
  }
  sc_finish(state, ctr, nblk);
}

// Segment heads of a sorted sequence in ONE launch: flag (key differs from the predecessor's), chained scan, and the per-element /
// per-head outputs -- instead of a heads kernel, a scan and a fill kernel (three launches, three sweeps over a flag array).
// F: key(i), emit(i, segment index of i, is head, key), end(number of segments)
template <typename F>
__global__ void __launch_bounds__(SC_THREADS) k_seg_heads(uint64_t n, F f, unsigned long long *__restrict__ state, unsigned *__restrict__ ctr,
                                                          unsigned nblk, int *__restrict__ total, const int *__restrict__ n_dev)
{
  __shared__ int s_prefix, s_agg;
  if (n_dev) n = (uint64_t)*n_dev;            // element count still on the device (the grid covers an upper bound; empty tiles pass the prefix on)
  const unsigned bid = sc_ticket(ctr);
  const uint64_t base = (uint64_t)bid * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
  uint64_t k[SC_ITEMS + 1];
  k[0] = (base > 0 && base - 1 < n) ? f.key(base - 1) : 0ull;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) k[i + 1] = (base + i < n) ? f.key(base + i) : 0ull;
  int v[SC_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) { v[i] = (base + i < n && (base + i == 0 || k[i + 1] != k[i])) ? 1 : 0; s += v[i]; }
  int ex = block_exclusive_scan(s);
  if (threadIdx.x == SC_THREADS - 1) s_agg = ex + s;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int prefix = sc_lookback_prefix(state, bid, s_agg, nblk, total);
    if (threadIdx.x == 0) { s_prefix = prefix; if (bid == nblk - 1) f.end(prefix + s_agg); }
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++)
    if (base + i < n) { ex += v[i]; f.emit(base + i, ex - 1, v[i], k[i + 1]); }
  sc_finish(state, ctr, nblk);
}

// Scan with the producer and the consumer of the values fused in: F: value(i) >= 0, emit(i, exclusive prefix, value), end(total)
template <typename F>
__global__ void __launch_bounds__(SC_THREADS) k_scan_emit(uint64_t n, F f, unsigned long long *__restrict__ state, unsigned *__restrict__ ctr,
                                                          unsigned nblk, int *__restrict__ total)
{
  __shared__ int s_prefix, s_agg;
  const unsigned bid = sc_ticket(ctr);
  const uint64_t base = (uint64_t)bid * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
  int v[SC_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) { v[i] = (base + i < n) ? f.value(base + i) : 0; s += v[i]; }
  int ex = block_exclusive_scan(s);
  if (threadIdx.x == SC_THREADS - 1) s_agg = ex + s;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int prefix = sc_lookback_prefix(state, bid, s_agg, nblk, total);
    if (threadIdx.x == 0) { s_prefix = prefix; if (bid == nblk - 1) f.end(prefix + s_agg); }
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++)
    if (base + i < n) { f.emit(base + i, ex, v[i]); ex += v[i]; }
  sc_finish(state, ctr, nblk);
}

// the zeroed look-back state of the context (grown on demand); the two counters sit in its last word
static inline void scan_state_reserve(ahfgpu_ctx *c, unsigned nblk)
{
  if ((size_t)nblk + 2 <= c->scan_cap) return;
  if (c->scan_state) ahf::dfree(c->scan_state);
  size_t cap = 4096; while (cap < (size_t)nblk + 2) cap <<= 1;
  c->scan_state = static_cast<unsigned long long *>(ahf::cache_alloc(cap * sizeof(unsigned long long)));
  CUDA_CHECK(cudaMemsetAsync(c->scan_state, 0, cap * sizeof(unsigned long long), c->stream));
  c->scan_cap = cap;
}
static inline unsigned *scan_state_ctr(ahfgpu_ctx *c) { return reinterpret_cast<unsigned *>(c->scan_state + (c->scan_cap - 1)); }

// out[i] = sum_{j<i} in[j]; no host synchronisation; d_total (device, may be null) receives the sum
template <typename T, bool NZ = false> void exclusive_scan_async(ahfgpu_ctx *c, const T *in, int *out, uint64_t n, int *d_total, DevBuf<int> &bs)
{
  if (n == 0) return;
  const unsigned nblk = (unsigned)((n + SC_TILE - 1) / SC_TILE);
  static const bool three_phase = getenv("AHFGPU_SCAN_V1") != nullptr;       // previous form (A/B timing; plain values only)
  if (three_phase && !NZ) {
    bs.reserve(nblk);
    LAUNCH(c, (k_sc_reduce<T>), nblk, SC_THREADS, 0, in, n, bs.p);
    LAUNCH(c, k_sc_small, 1, 1024, 0, bs.p, (uint64_t)nblk, d_total);
    LAUNCH(c, (k_sc_down<T>), nblk, SC_THREADS, 0, in, n, bs.p, out);
    return;
  }
  scan_state_reserve(c, nblk);
  unsigned *ctr = scan_state_ctr(c);
  LAUNCH(c, (k_sc_lookback<T, NZ>), nblk, SC_THREADS, 0, in, n, out, c->scan_state, ctr, nblk, d_total);
}

// segment heads of n > 0 sorted elements (see k_seg_heads); d_total (device, may be null) receives the number of segments
// n_dev (device, may be null): the real element count, n then being an upper bound known to the host
template <typename F> void seg_heads_async(ahfgpu_ctx *c, uint64_t n, const F &f, int *d_total, const int *n_dev = nullptr)
{
  const unsigned nblk = (unsigned)((n + SC_TILE - 1) / SC_TILE);
  scan_state_reserve(c, nblk);
  unsigned *ctr = scan_state_ctr(c);
  LAUNCH(c, (k_seg_heads<F>), nblk, SC_THREADS, 0, n, f, c->scan_state, ctr, nblk, d_total, n_dev);
}

template <typename F> void scan_emit_async(ahfgpu_ctx *c, uint64_t n, const F &f, int *d_total)
{
  const unsigned nblk = (unsigned)((n + SC_TILE - 1) / SC_TILE);
  scan_state_reserve(c, nblk);
  unsigned *ctr = scan_state_ctr(c);
  LAUNCH(c, (k_scan_emit<F>), nblk, SC_THREADS, 0, n, f, c->scan_state, ctr, nblk, d_total);
}

// out[i] = sum_{j<i} in[j]; returns the total (synchronises the stream)
template <typename T> int exclusive_scan(ahfgpu_ctx *c, const T *in, int *out, uint64_t n)
{
  if (n == 0) return 0;
  DevBuf<int> bs, tot;
  tot.reserve(1);
  exclusive_scan_async<T>(c, in, out, n, tot.p, bs);
  int h = 0;
  read_back(c, &h, tot.p, sizeof(int));
  bs.release(); tot.release();
  return h;
}

}  // namespace ahf

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

// out[i] = sum_{j<i} in[j]; no host synchronisation; d_total (device, may be null) receives the sum
template <typename T> void exclusive_scan_async(ahfgpu_ctx *c, const T *in, int *out, uint64_t n, int *d_total, DevBuf<int> &bs)
{
  if (n == 0) return;
  const unsigned nblk = (unsigned)((n + SC_TILE - 1) / SC_TILE);
  bs.reserve(nblk);
  LAUNCH(c, (k_sc_reduce<T>), nblk, SC_THREADS, 0, in, n, bs.p);
  LAUNCH(c, k_sc_small, 1, 1024, 0, bs.p, (uint64_t)nblk, d_total);
  LAUNCH(c, (k_sc_down<T>), nblk, SC_THREADS, 0, in, n, bs.p, out);
}

// out[i] = sum_{j<i} in[j]; returns the total (synchronises the stream)
template <typename T> int exclusive_scan(ahfgpu_ctx *c, const T *in, int *out, uint64_t n)
{
  if (n == 0) return 0;
  DevBuf<int> bs, tot;
  tot.reserve(1);
  exclusive_scan_async<T>(c, in, out, n, tot.p, bs);
  int h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, tot.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  bs.release(); tot.release();
  return h;
}

}  // namespace ahf

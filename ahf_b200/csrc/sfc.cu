// sfc.cu -- K1 Hilbert keys + K2 LSD radix sort + payload gather (sm_100a).
//
// Replaces the serial key loop and libc qsort of the reference (src/main.c:343-356): keys are the 63-bit
// Hilbert index of (trunc(x*2^21), trunc(y*2^21), trunc(z*2^21)) (src/libsfc/hilbert_util.c:69-92,
// hilbert.c:197-243).  HBM-bound integer work: no tensor cores.
#include "common.cuh"
#include "hilbert.cuh"
#include "scan.cuh"

namespace ahf {

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
__global__ void k_keys_soa(const float *__restrict__ pos3, uint64_t n, uint32_t bits, uint64_t *__restrict__ keys,
                           uint32_t *__restrict__ idx)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = pos3[3 * i], y = pos3[3 * i + 1], z = pos3[3 * i + 2];
  keys[i] = hilbert_key_pos(x, y, z, bits);
  if (idx) idx[i] = (uint32_t)i;
}

// 21-bit keys through the three-levels-per-look-up table (hilbert.cuh): the table (12 KB) is staged in shared memory by every
// CTA, which then strides over the particles [i0, i1).  The generic k_keys_soa needs ~420 dependent integer instructions per key
// and was issue bound (0.43 ms for 256^3); this form needs about a third.
__device__ uint16_t g_hil_tab3[12 * 512];
static void upload_hil_tab3() {}      // done per device in sfc_device_init (ahfgpu_init)
constexpr int KT_THREADS = 256;
__global__ void __launch_bounds__(KT_THREADS) k_keys_soa_tab(const float *__restrict__ pos3, uint64_t i0, uint64_t i1, uint64_t *__restrict__ keys,
                                                             uint32_t *__restrict__ idx)
{
  __shared__ __align__(16) uint16_t tab[12 * 512];
  for (int i = threadIdx.x; i < 12 * 512 / 8; i += KT_THREADS) reinterpret_cast<uint4 *>(tab)[i] = reinterpret_cast<const uint4 *>(g_hil_tab3)[i];
  __syncthreads();
  for (uint64_t i = i0 + blockIdx.x * (uint64_t)KT_THREADS + threadIdx.x; i < i1; i += (uint64_t)gridDim.x * KT_THREADS) {
    keys[i] = hilbert_key_pos21_tab(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], tab);
    if (idx) idx[i] = (uint32_t)i;
  }
}
static unsigned keys_tab_grid(uint64_t n) { const uint64_t b = (n + KT_THREADS - 1) / KT_THREADS; return (unsigned)(b < 148 * 16 ? (b ? b : 1) : 148 * 16); }

// reference AoS record: positions at byte offset off_pos of a record of `stride` bytes
__global__ void k_keys_aos(const unsigned char *__restrict__ rec, uint64_t n, uint32_t stride, int off_pos,
                           uint64_t *__restrict__ keys, uint32_t *__restrict__ idx)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *p = reinterpret_cast<const float *>(rec + i * stride + off_pos);
  keys[i] = hilbert_key_pos(p[0], p[1], p[2], 21);
  idx[i]  = (uint32_t)i;
}

// ------------------------------------------------------------------------------------------------
// K2: LSD radix sort, RS_BITS-bit digits, (u64 key, u32 value) pairs, stable.  8 bits (8 passes over the 63 key bits): measured against
//   9 bits (7 passes) at 256^3: 0.153 vs 0.187 ms per scatter launch -- with 512 digits the digit runs of a 2048-pair tile are 4 pairs
//   long and the run-by-run output no longer fills its sectors; 1.63 vs 1.80 ms for the whole sort (profiles/r2i_bench_256_n1.json)
//   per pass: block histograms -> exclusive scan over [digit][block] -> ranked scatter
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS   = RS_THREADS / 32;
constexpr int RS_BITS    = 8;
constexpr int RS_NB      = 1 << RS_BITS;             // digits per pass
constexpr int RS_DPT     = RS_NB / RS_THREADS;       // digits per thread in the digit loops (consecutive digits)
// pairs per thread: 8 (tile 2048, 3 CTAs/SM) or 16 (tile 4096: digit runs twice as long, so the run-by-run output fills its
// sectors better; 128 registers, 2 CTAs/SM).  Chosen at run time in radix_sort_pairs (AHFGPU_RS_ITEMS).
#define RS_TILE   (RS_THREADS * RS_ITEMS)
#define RS_WCHUNK (RS_TILE / RS_WARPS)

// Digit histogram of every scatter tile.  One CTA counts RS_HT consecutive tiles (32 keys per thread, all loads issued before the first
// shared atomic): with one 2048-key tile per CTA the kernel was bound by CTA turnover (8192 short-lived CTAs, 43 us for 134 MB at 256^3),
// and a thread now writes RS_HT consecutive words of the digit-major table instead of one word per 32-byte sector.
#define RS_HT (32 / RS_ITEMS)
template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t *__restrict__ keys, uint64_t n, int shift,
                                                        uint32_t *__restrict__ bhist, uint32_t nblk)
{
  __shared__ uint32_t h[RS_HT][RS_NB];
#pragma unroll
  for (int q = 0; q < RS_HT; q++)
#pragma unroll
    for (int d = 0; d < RS_DPT; d++) h[q][threadIdx.x + d * RS_THREADS] = 0;
  __syncthreads();
  const uint32_t t0 = blockIdx.x * RS_HT;
  const uint64_t base = (uint64_t)t0 * RS_TILE + threadIdx.x;
  uint64_t kk[RS_HT * RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_HT * RS_ITEMS; i++) {                       // element i * RS_THREADS of the CTA's span: tile i / RS_ITEMS
    const uint64_t j = base + (uint64_t)i * RS_THREADS;
    kk[i] = j < n ? keys[j] : 0ull;
  }
#pragma unroll
  for (int i = 0; i < RS_HT * RS_ITEMS; i++) {
    const uint64_t j = base + (uint64_t)i * RS_THREADS;
    if (j < n) atomicAdd(&h[i / RS_ITEMS][(uint32_t)(kk[i] >> shift) & (uint32_t)(RS_NB - 1)], 1u);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < RS_HT; q++)
    if (t0 + q < nblk) {
#pragma unroll
      for (int d = 0; d < RS_DPT; d++) { const int dg = threadIdx.x + d * RS_THREADS; bhist[(uint64_t)dg * nblk + t0 + q] = h[q][dg]; }
    }
}


// ranked scatter.  The tile is first sorted by digit in shared memory (stable), then written out digit run by digit run so that
// consecutive threads store consecutive addresses (a direct scatter writes 12 useful bytes per pair of 32-byte sectors).
#define RS_SMEM (RS_TILE * 12 + RS_WARPS * RS_NB * 4 + 2 * RS_NB * 4 + 64)
template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS, RS_ITEMS == 8 ? 4 : 2) k_rs_scatter(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                           uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, uint64_t n,
                                                           int shift, const uint32_t *__restrict__ bscan, uint32_t nblk)
{
  extern __shared__ __align__(16) unsigned char rsm[];
  uint64_t *sk = reinterpret_cast<uint64_t *>(rsm);                                  // [RS_TILE]
  uint32_t *sv = reinterpret_cast<uint32_t *>(rsm + RS_TILE * 8);                    // [RS_TILE]
  uint32_t (*whist)[RS_NB] = reinterpret_cast<uint32_t (*)[RS_NB]>(rsm + RS_TILE * 12);  // [RS_WARPS][RS_NB]
  uint32_t *gbase = reinterpret_cast<uint32_t *>(rsm + RS_TILE * 12 + RS_WARPS * RS_NB * 4);   // [RS_NB] global address of tile slot 0 of the digit's run
  uint32_t *wtot  = gbase + RS_NB;                                                   // [RS_WARPS] scratch of the digit scan
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * RS_NB; i += RS_THREADS) (&whist[0][0])[i] = 0;
  __syncthreads();
  const uint64_t tile0 = (uint64_t)blockIdx.x * RS_TILE, base = tile0 + (uint64_t)w * RS_WCHUNK;
  uint64_t k[RS_ITEMS];
  uint32_t v[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
#pragma unroll
  for (int s = 0; s < RS_ITEMS; s++) {
    uint64_t j = base + (uint64_t)s * 32 + lane;
    bool     valid = j < n;
    k[s] = valid ? kin[j] : ~0ull;
    v[s] = valid ? vin[j] : 0u;
  }
#pragma unroll
  for (int s = 0; s < RS_ITEMS; s++) {
    uint64_t j = base + (uint64_t)s * 32 + lane;
    bool     valid = j < n;
    uint32_t d  = (uint32_t)(k[s] >> shift) & (uint32_t)(RS_NB - 1);
    uint32_t dd = valid ? d : ((uint32_t)RS_NB + lane);
    uint32_t peers  = __match_any_sync(0xffffffffu, dd);
    int      leader = __ffs(peers) - 1;
    uint32_t r = __popc(peers & ((1u << lane) - 1u));
    uint32_t bc = 0;
    if (lane == leader && valid) { bc = whist[w][d]; whist[w][d] = bc + __popc(peers); }
    bc = __shfl_sync(0xffffffffu, bc, leader);
    rank[s] = bc + r;
    __syncwarp();
  }
  __syncthreads();
  {
    // thread t owns the RS_DPT consecutive digits t*RS_DPT ...: counts of the warps -> exclusive offsets inside the digit, then an
    // exclusive scan over all digits: where the digit's run starts inside the tile
    const int t = threadIdx.x;
    uint32_t  run[RS_DPT], tot = 0;
#pragma unroll
    for (int d = 0; d < RS_DPT; d++) {
      const int dg = t * RS_DPT + d;
      uint32_t  r = 0;
#pragma unroll
      for (int q = 0; q < RS_WARPS; q++) { uint32_t c = whist[q][dg]; whist[q][dg] = r; r += c; }
      run[d] = r; tot += r;
    }
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) wtot[w] = inc;
    __syncthreads();
    uint32_t wb = 0;
#pragma unroll
    for (int q = 0; q < RS_WARPS; q++) if (q < w) wb += wtot[q];
    uint32_t dstart = wb + inc - tot;
#pragma unroll
    for (int d = 0; d < RS_DPT; d++) {
      const int dg = t * RS_DPT + d;
#pragma unroll
      for (int q = 0; q < RS_WARPS; q++) whist[q][dg] += dstart;
      gbase[dg] = bscan[(uint64_t)dg * nblk + blockIdx.x] - dstart;
      dstart += run[d];
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < RS_ITEMS; s++) {
    uint64_t j = base + (uint64_t)s * 32 + lane;
    if (j < n) {
      uint32_t d = (uint32_t)(k[s] >> shift) & (uint32_t)(RS_NB - 1);
      uint32_t lp = whist[w][d] + rank[s];
      sk[lp] = k[s]; sv[lp] = v[s];
    }
  }
  __syncthreads();
  const int nv = (int)((n - tile0) < (uint64_t)RS_TILE ? (n - tile0) : (uint64_t)RS_TILE);
#pragma unroll 4
  for (int i = threadIdx.x; i < nv; i += RS_THREADS) {
    const uint64_t kk = sk[i];
    const uint32_t p = gbase[(uint32_t)(kk >> shift) & (uint32_t)(RS_NB - 1)] + (uint32_t)i;
    kout[p] = kk; vout[p] = sv[i];
  }
}

template <int RS_ITEMS>
static void radix_sort_pairs_t(ahfgpu_ctx *c, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, uint64_t n,
                               int key_bits, uint64_t **keys_sorted, uint32_t **vals_sorted, int first_bit)
{
  const uint32_t nblk = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
  DevBuf<uint32_t> bh;
  DevBuf<int>      bs;
  bh.reserve((size_t)RS_NB * nblk);
  uint64_t *ki = keys, *ko = keys_tmp;
  uint32_t *vi = vals, *vo = vals_tmp;
  // AHFGPU_KERNEL_STAGES=1 (bench.py's instrumented passes): event pair around every scatter launch, the largest kernel share of a pass
  const bool ktimer = getenv("AHFGPU_KERNEL_STAGES") != nullptr;
  for (int shift = first_bit; shift < key_bits; shift += RS_BITS) {    // bits below first_bit are left to the caller (ties)
    LAUNCH(c, k_rs_hist<RS_ITEMS>, (nblk + RS_HT - 1) / RS_HT, RS_THREADS, 0, ki, n, shift, bh.p, nblk);
    exclusive_scan_async<int>(c, (const int *)bh.p, (int *)bh.p, (uint64_t)RS_NB * nblk, nullptr, bs);   // in place: each tile is read before it is written
    {
      Stage sk(c, "rs_scatter_kernel", (int64_t)n, ktimer);
      LAUNCH(c, k_rs_scatter<RS_ITEMS>, nblk, RS_THREADS, RS_SMEM, ki, vi, ko, vo, n, shift, bh.p, nblk);
    }
    uint64_t *tk = ki; ki = ko; ko = tk;
    uint32_t *tv = vi; vi = vo; vo = tv;
  }
  bh.release(); bs.release();                       // stream-ordered block cache: no host sync needed
  *keys_sorted = ki; *vals_sorted = vi;
}

// number of passes radix_sort_pairs will make (callers that want the result in a particular buffer pick the start buffer by parity)
int radix_sort_passes(int key_bits, int first_bit) { int p = 0; for (int s = first_bit; s < key_bits; s += RS_BITS) p++; return p; }

// per-DEVICE setup of the sort kernels and the Hilbert table (called once per device from ahfgpu_init)
void sfc_device_init()
{
  {
    constexpr int RS_ITEMS = 8;
    CUDA_CHECK(cudaFuncSetAttribute(k_rs_scatter<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
  }
  {
    constexpr int RS_ITEMS = 16;
    CUDA_CHECK(cudaFuncSetAttribute(k_rs_scatter<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
  }
  std::vector<uint16_t> h(12 * 512);
  hilbert_build_tab3(h.data());
  CUDA_CHECK(cudaMemcpyToSymbol(g_hil_tab3, h.data(), h.size() * sizeof(uint16_t)));
}

int radix_sort_passes(int key_bits, int first_bit);
void radix_sort_pairs(ahfgpu_ctx *c, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, uint64_t n,
                      int key_bits, uint64_t **keys_sorted, uint32_t **vals_sorted, int first_bit)
{
  *keys_sorted = keys; *vals_sorted = vals;
  if (n == 0) return;
  static const int items = getenv("AHFGPU_RS_ITEMS") ? atoi(getenv("AHFGPU_RS_ITEMS")) : 8;      // measured at 256^3: 1.81 ms (8) vs 2.68 ms (16): the ranking loop is latency bound, occupancy wins
  if (items == 8) radix_sort_pairs_t<8>(c, keys, vals, keys_tmp, vals_tmp, n, key_bits, keys_sorted, vals_sorted, first_bit);
  else radix_sort_pairs_t<16>(c, keys, vals, keys_tmp, vals_tmp, n, key_bits, keys_sorted, vals_sorted, first_bit);
}

// ------------------------------------------------------------------------------------------------
// K2, the 63-bit Hilbert keys of the particles: six passes instead of eight.  The LSD passes sort bits KS_SKIP..62 only; particles
// whose keys agree in all of those bits sit in one cell of 2^-16 of the box per dimension -- none on the lattice part of a box, pairs
// and triples in clump cores -- and are still in input order.  k_fix_key_ties puts every such run into the order of the full stable
// sort (insertion sort by the whole key; equal keys keep their input order): one thread per run head, two key loads per particle
// otherwise.  A run longer than KS_MAXRUN (thousands of particles inside one 2^-16 cell: degenerate inputs) raises a flag and the
// arrangement is sorted again with all eight passes -- the stable sort of a stably pre-sorted array is the full stable sort.
// ------------------------------------------------------------------------------------------------
constexpr int KS_SKIP   = 15;
constexpr int KS_MAXRUN = 32;
__global__ void k_fix_key_ties(uint64_t *keys, uint32_t *vals, uint64_t n, int *flag)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i + 1 >= n) return;
  const uint64_t top = keys[i] >> KS_SKIP;
  if ((keys[i + 1] >> KS_SKIP) != top) return;                       // alone, or the last of its run
  if (i > 0 && (keys[i - 1] >> KS_SKIP) == top) return;              // not the head (the top bits of a position never change below)
  uint64_t j = i + 2;
  while (j < n && j - i <= (uint64_t)KS_MAXRUN && (keys[j] >> KS_SKIP) == top) j++;
  if (j - i > (uint64_t)KS_MAXRUN) { *flag = 1; return; }
  for (uint64_t a = i + 1; a < j; a++) {
    const uint64_t ka = keys[a];
    const uint32_t va = vals[a];
    uint64_t b = a;
    while (b > i) {
      const uint64_t kb = keys[b - 1];
      if (kb <= ka) break;
      keys[b] = kb; vals[b] = vals[b - 1]; b--;
    }
    if (b != a) { keys[b] = ka; vals[b] = va; }
  }
}
int sort_keys63_passes() { const bool full = getenv("AHFGPU_SORT_FULL") != nullptr; return radix_sort_passes(63, full ? 0 : KS_SKIP); }
void sort_keys63(ahfgpu_ctx *c, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, uint64_t **keys_sorted, uint32_t **vals_sorted)
{
  const bool full = getenv("AHFGPU_SORT_FULL") != nullptr;              // A/B timing and the parity tests of the tie fix
  if (full || n < 2) { radix_sort_pairs(c, keys, vals, keys_tmp, vals_tmp, n, 63, keys_sorted, vals_sorted, 0); return; }
  radix_sort_pairs(c, keys, vals, keys_tmp, vals_tmp, n, 63, keys_sorted, vals_sorted, KS_SKIP);
  DevBuf<int> flag;
  flag.reserve(1);
  CUDA_CHECK(cudaMemsetAsync(flag.p, 0, sizeof(int), c->stream));
  LAUNCH(c, k_fix_key_ties, (unsigned)((n + 255) / 256), 256, 0, *keys_sorted, *vals_sorted, n, flag.p);
  int h = 0;
  read_back(c, &h, flag.p, sizeof(int));
  flag.release();
  c->stage_cnt_extra["sort_full_fallback"] = h ? 1 : 0;
  if (h) {
    uint64_t *k0 = *keys_sorted, *k1 = (k0 == keys) ? keys_tmp : keys;
    uint32_t *v0 = *vals_sorted, *v1 = (v0 == vals) ? vals_tmp : vals;
    radix_sort_pairs(c, k0, v0, k1, v1, n, 63, keys_sorted, vals_sorted, 0);       // eight passes: ends in the buffer it started in
  }
}

// ------------------------------------------------------------------------------------------------
// Stable merge of two key-sorted runs of (u64, u32) pairs, A before B on equal keys (ahfgpu_sfc_sort_soa_async sorts the position chunks
// as they arrive over the bus and merges them once the last one is there: only one chunk sort and two merge sweeps are left behind the
// upload instead of the whole sort).  A CTA produces MG_TILE consecutive outputs: its share of A and of B by two merge-path searches
// (the split of diagonal d takes one more element of A as long as A[a] <= B[d - a - 1]), both shares into shared memory, then every
// element finds its output slot by one binary search in the other share (A: elements of B strictly smaller; B: elements of A not larger).
// ------------------------------------------------------------------------------------------------
constexpr int MG_THREADS = 256, MG_TILE = 2048;
__device__ __forceinline__ int64_t merge_path_split(const uint64_t *__restrict__ ka, int64_t na, const uint64_t *__restrict__ kb, int64_t nb, int64_t d)
{
  int64_t lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (ka[mid] <= kb[d - mid - 1]) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void __launch_bounds__(MG_THREADS) k_merge_pairs(const uint64_t *__restrict__ ka, const uint32_t *__restrict__ va, int64_t na,
                                                           const uint64_t *__restrict__ kb, const uint32_t *__restrict__ vb, int64_t nb,
                                                           uint64_t *__restrict__ ko, uint32_t *__restrict__ vo)
{
  __shared__ uint64_t sk[MG_TILE];
  __shared__ uint32_t sv[MG_TILE];
  __shared__ int64_t  s_split[2];
  const int64_t d0 = (int64_t)blockIdx.x * MG_TILE, d1 = d0 + MG_TILE < na + nb ? d0 + MG_TILE : na + nb;
  if (threadIdx.x == 0) s_split[0] = merge_path_split(ka, na, kb, nb, d0);
  if (threadIdx.x == 32) s_split[1] = merge_path_split(ka, na, kb, nb, d1);
  __syncthreads();
  const int64_t a0 = s_split[0], a1 = s_split[1], b0 = d0 - a0, b1 = d1 - a1;
  const int     ca = (int)(a1 - a0), cb = (int)(b1 - b0);                 // ca + cb = d1 - d0 <= MG_TILE
  for (int i = threadIdx.x; i < ca; i += MG_THREADS) { sk[i] = ka[a0 + i]; sv[i] = va[a0 + i]; }
  for (int i = threadIdx.x; i < cb; i += MG_THREADS) { sk[ca + i] = kb[b0 + i]; sv[ca + i] = vb[b0 + i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < ca + cb; i += MG_THREADS) {
    const uint64_t k = sk[i];
    int pos;
    if (i < ca) {                                       // from A: behind the elements of B that are strictly smaller
      int lo = 0, hi = cb;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sk[ca + mid] < k) lo = mid + 1; else hi = mid; }
      pos = i + lo;
    } else {                                            // from B: behind the elements of A that are not larger
      int lo = 0, hi = ca;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sk[mid] <= k) lo = mid + 1; else hi = mid; }
      pos = (i - ca) + lo;
    }
    ko[d0 + pos] = k; vo[d0 + pos] = sv[i];
  }
}
static void merge_pairs(ahfgpu_ctx *c, const uint64_t *ka, const uint32_t *va, int64_t na, const uint64_t *kb, const uint32_t *vb, int64_t nb, uint64_t *ko, uint32_t *vo)
{
  if (na + nb == 0) return;
  LAUNCH(c, k_merge_pairs, (unsigned)((na + nb + MG_TILE - 1) / MG_TILE), MG_THREADS, 0, ka, va, na, kb, vb, nb, ko, vo);
}

// ------------------------------------------------------------------------------------------------
// payload gather
// ------------------------------------------------------------------------------------------------
__global__ void k_gather_soa(const float *__restrict__ pos3, const float *__restrict__ mom3, const float *__restrict__ w,
                             const float *__restrict__ u, const uint32_t *__restrict__ order, uint64_t n,
                             float4 *__restrict__ pos4, float4 *__restrict__ mom4)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o = order[i];
  pos4[i] = make_float4(pos3[3 * o], pos3[3 * o + 1], pos3[3 * o + 2], w ? w[o] : 1.0f);
  mom4[i] = make_float4(mom3[3 * o], mom3[3 * o + 1], mom3[3 * o + 2], u ? u[o] : -1.0f);
}

// sorted AoS: copy whole records in 8-byte words, then patch sfckey and clear the leading `ll` pointer
__global__ void k_gather_aos(const unsigned char *__restrict__ in, unsigned char *__restrict__ out,
                             const uint32_t *__restrict__ order, const uint64_t *__restrict__ keys, uint64_t n, uint32_t stride,
                             int off_pos, int off_mom, int off_key, int off_w, int off_u, float4 *__restrict__ pos4,
                             float4 *__restrict__ mom4)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t *src = reinterpret_cast<const uint64_t *>(in + (uint64_t)order[i] * stride);
  uint64_t       *dst = reinterpret_cast<uint64_t *>(out + i * stride);
  const int       nw  = stride / 8;
  for (int q = 0; q < nw; q++) dst[q] = src[q];
  dst[0] = 0;                                                      // part.ll: first member (tdef.h:38-42)
  *reinterpret_cast<uint64_t *>(out + i * stride + off_key) = keys[i];
  const float *p = reinterpret_cast<const float *>(out + i * stride + off_pos);
  const float *m = reinterpret_cast<const float *>(out + i * stride + off_mom);
  float ww = off_w >= 0 ? *reinterpret_cast<const float *>(out + i * stride + off_w) : 1.0f;
  float uu = off_u >= 0 ? *reinterpret_cast<const float *>(out + i * stride + off_u) : -1.0f;
  pos4[i] = make_float4(p[0], p[1], p[2], ww);
  mom4[i] = make_float4(m[0], m[1], m[2], uu);
}

static void alloc_particles(ahfgpu_ctx *c, uint64_t n)
{
  c->free_particles();
  c->free_levels();
  c->free_halos();
  c->n = n;
  c->pos4 = static_cast<decltype(c->pos4)>(ahf::cache_alloc((n ? n : 1) * sizeof(float4)));
  c->mom4 = static_cast<decltype(c->mom4)>(ahf::cache_alloc((n ? n : 1) * sizeof(float4)));
}

void sfc_keys_only(ahfgpu_ctx *c, const float *pos3, uint64_t n, uint32_t bits, uint64_t *keys_out)
{
  DevBuf<float>    dpos;
  DevBuf<uint64_t> dk;
  dpos.reserve(3 * n); dk.reserve(n);
  CUDA_CHECK(cudaMemcpyAsync(dpos.p, pos3, 3 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  if (n) LAUNCH(c, k_keys_soa, (unsigned)((n + 255) / 256), 256, 0, dpos.p, n, bits, dk.p, (uint32_t *)nullptr);
  CUDA_CHECK(cudaMemcpyAsync(keys_out, dk.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  dpos.release(); dk.release();
}

void sfc_upload_soa(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *w, const float *u, uint64_t n)
{
  if (n >= (1ull << 32)) AHF_FAIL("more than 2^32-1 particles per device are not supported");
  c->wait_mom(false);
  ahf::dfree(c->in_pos); ahf::dfree(c->in_mom); ahf::dfree(c->in_w); ahf::dfree(c->in_u);
  c->in_pos = c->in_mom = c->in_w = c->in_u = nullptr; c->in_n = n;
  c->in_pos = static_cast<decltype(c->in_pos)>(ahf::cache_alloc((n ? n : 1) * 3 * sizeof(float)));
  c->in_mom = static_cast<decltype(c->in_mom)>(ahf::cache_alloc((n ? n : 1) * 3 * sizeof(float)));
  if (w) c->in_w = static_cast<decltype(c->in_w)>(ahf::cache_alloc((n ? n : 1) * sizeof(float)));
  if (u) c->in_u = static_cast<decltype(c->in_u)>(ahf::cache_alloc((n ? n : 1) * sizeof(float)));
  Stage st(c, "h2d", (int64_t)(24 * n + (w ? 4 * n : 0) + (u ? 4 * n : 0)));
  CUDA_CHECK(cudaMemcpyAsync(c->in_pos, pos3, 3 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(c->in_mom, mom3, 3 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  if (w) CUDA_CHECK(cudaMemcpyAsync(c->in_w, w, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  if (u) CUDA_CHECK(cudaMemcpyAsync(c->in_u, u, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void sfc_sort_resident(ahfgpu_ctx *c, uint64_t *keys_out, uint32_t *order_out)
{
  if (!c->in_pos) AHF_FAIL("no uploaded particles: call ahfgpu_upload_soa first");
  const uint64_t n = c->in_n;
  alloc_particles(c, n);
  c->has_weight = (c->in_w != nullptr); c->has_u = (c->in_u != nullptr);
  // the resident key / order arrays double as the first pair of sort buffers: an even number of passes ends in them
  c->keys = static_cast<decltype(c->keys)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint64_t)));
  c->order = static_cast<decltype(c->order)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint32_t)));
  DevBuf<uint64_t> k1;
  DevBuf<uint32_t> v1;
  k1.reserve(n); v1.reserve(n);
  const unsigned nb = (unsigned)((n + 255) / 256);
  // an odd number of passes ends in the OTHER buffer: start in the scratch pair then, so that the result lands in keys / order
  const bool odd = (sort_keys63_passes() & 1) != 0;
  uint64_t *kA = odd ? k1.p : c->keys, *kB = odd ? c->keys : k1.p;
  uint32_t *vA = odd ? v1.p : c->order, *vB = odd ? c->order : v1.p;
  {
    Stage st(c, "keys", (int64_t)n);
    if (n) { upload_hil_tab3(); LAUNCH(c, k_keys_soa_tab, keys_tab_grid(n), KT_THREADS, 0, c->in_pos, (uint64_t)0, n, kA, vA); }
  }
  uint64_t *ks; uint32_t *vs;
  {
    Stage st(c, "sort", (int64_t)n);
    sort_keys63(c, kA, vA, kB, vB, n, &ks, &vs);
    if (ks != c->keys) {
      CUDA_CHECK(cudaMemcpyAsync(c->keys, ks, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
      CUDA_CHECK(cudaMemcpyAsync(c->order, vs, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  {
    Stage st(c, "gather", (int64_t)n);
    if (n) LAUNCH(c, k_gather_soa, nb, 256, 0, c->in_pos, c->in_mom, c->in_w, c->in_u, c->order, n, c->pos4, c->mom4);
  }
  if (keys_out || order_out) {
    Stage st(c, "d2h", (int64_t)((keys_out ? 8 * n : 0) + (order_out ? 4 * n : 0)));
    if (keys_out) CUDA_CHECK(cudaMemcpyAsync(keys_out, c->keys, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    if (order_out) CUDA_CHECK(cudaMemcpyAsync(order_out, c->order, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  k1.release(); v1.release();
}

void sfc_sort_soa(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *w, const float *u, uint64_t n,
                  uint64_t *keys_out, uint32_t *order_out)
{
  sfc_upload_soa(c, pos3, mom3, w, u, n);
  sfc_sort_resident(c, keys_out, order_out);
}

// ------------------------------------------------------------------------------------------------
// Overlapped host->device path (ahfgpu_sfc_sort_soa_async).  All copies go through ctx->copy_stream in the order
// pos (four chunks), weight, mom, u; the main stream computes the keys of a chunk as soon as it has landed, sorts, and
// gathers pos4 -- everything ahfgpu_build_amr needs -- while the momenta are still on the bus.  The momentum gather runs on
// the copy stream behind its copy and signals ev_mom, which the halo pass waits for on the device.
// ------------------------------------------------------------------------------------------------
__global__ void k_gather_pos(const float *__restrict__ pos3, const float *__restrict__ w, const uint32_t *__restrict__ order, uint64_t n,
                             float4 *__restrict__ pos4)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o = order[i];
  pos4[i] = make_float4(pos3[3 * o], pos3[3 * o + 1], pos3[3 * o + 2], w ? w[o] : 1.0f);
}
__global__ void k_gather_mom(const float *__restrict__ mom3, const float *__restrict__ u, const uint32_t *__restrict__ order, uint64_t n,
                             float4 *__restrict__ mom4)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o = order[i];
  mom4[i] = make_float4(mom3[3 * o], mom3[3 * o + 1], mom3[3 * o + 2], u ? u[o] : -1.0f);
}

void sfc_sort_soa_async(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *w, const float *u, uint64_t n)
{
  if (n >= (1ull << 32)) AHF_FAIL("more than 2^32-1 particles per device are not supported");
  c->wait_mom(false);
  if (!c->copy_stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto &e : c->ev_copy) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_mom, cudaEventDisableTiming));
  }
  ahf::dfree(c->in_pos); ahf::dfree(c->in_mom); ahf::dfree(c->in_w); ahf::dfree(c->in_u);
  c->in_pos = c->in_mom = c->in_w = c->in_u = nullptr; c->in_n = n;
  c->in_pos = static_cast<decltype(c->in_pos)>(ahf::cache_alloc((n ? n : 1) * 3 * sizeof(float)));
  c->in_mom = static_cast<decltype(c->in_mom)>(ahf::cache_alloc((n ? n : 1) * 3 * sizeof(float)));
  if (w) c->in_w = static_cast<decltype(c->in_w)>(ahf::cache_alloc((n ? n : 1) * sizeof(float)));
  if (u) c->in_u = static_cast<decltype(c->in_u)>(ahf::cache_alloc((n ? n : 1) * sizeof(float)));
  alloc_particles(c, n);
  c->has_weight = (w != nullptr); c->has_u = (u != nullptr);
  c->keys = static_cast<decltype(c->keys)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint64_t)));
  c->order = static_cast<decltype(c->order)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint32_t)));
  DevBuf<uint64_t> k1;
  DevBuf<uint32_t> v1;
  k1.reserve(n); v1.reserve(n);                     // c->keys / c->order are the first pair of sort buffers
  // the blocks above are ordered on the main stream (they may be recycled): the copy stream starts behind this point
  CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
  CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
  constexpr int NCH = 4;
  upload_hil_tab3();
  const bool odd = (sort_keys63_passes() & 1) != 0;      // see sfc_sort_resident
  // chunk-wise sort + merge (default for boxes that are worth it; AHFGPU_ASYNC_SORT_WHOLE=1: keys per chunk, ONE sort behind the upload)
  const bool chunked = n >= (1u << 20) && !odd && getenv("AHFGPU_ASYNC_SORT_WHOLE") == nullptr;
  uint64_t *kA = odd ? k1.p : c->keys, *kB = odd ? c->keys : k1.p;
  uint32_t *vA = odd ? v1.p : c->order, *vB = odd ? c->order : v1.p;
  const uint64_t per = ((n + NCH - 1) / NCH + 255) & ~255ull;
  uint64_t ci0[NCH], ci1[NCH];
  // all copies are queued first: the copy stream runs on its own while the host waits for the tie flag of a chunk's sort
  for (int q = 0; q < NCH; q++) {
    ci0[q] = std::min(n, per * q); ci1[q] = std::min(n, per * (q + 1));
    if (ci1[q] > ci0[q]) CUDA_CHECK(cudaMemcpyAsync(c->in_pos + 3 * ci0[q], pos3 + 3 * ci0[q], 3 * (ci1[q] - ci0[q]) * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_CHECK(cudaEventRecord(c->ev_copy[q], c->copy_stream));
  }
  if (w) CUDA_CHECK(cudaMemcpyAsync(c->in_w, w, n * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
  CUDA_CHECK(cudaEventRecord(c->ev_copy[NCH], c->copy_stream));
  CUDA_CHECK(cudaMemcpyAsync(c->in_mom, mom3, 3 * n * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
  if (u) CUDA_CHECK(cudaMemcpyAsync(c->in_u, u, n * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
  {
    Stage st(c, "keys", (int64_t)n);
    for (int q = 0; q < NCH; q++) {
      CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_copy[q], 0));
      const uint64_t i0 = ci0[q], i1 = ci1[q];
      if (i1 > i0) LAUNCH(c, k_keys_soa_tab, keys_tab_grid(i1 - i0), KT_THREADS, 0, c->in_pos, i0, i1, kA, vA);
      if (chunked && i1 > i0) {
        // the chunk is sorted in place (an even number of passes ends in the buffer it started in) while the next one is on the bus
        uint64_t *cks; uint32_t *cvs;
        sort_keys63(c, kA + i0, vA + i0, kB + i0, vB + i0, i1 - i0, &cks, &cvs);
        if (cks != kA + i0) AHF_FAIL("chunk sort did not end in its own buffer");
        // chunks 0+1 merge as soon as both are sorted (kA -> kB), 2+3 likewise, the two halves at the end (kB -> kA)
        if (q == 1) merge_pairs(c, kA + ci0[0], vA + ci0[0], (int64_t)(ci1[0] - ci0[0]), kA + ci0[1], vA + ci0[1], (int64_t)(ci1[1] - ci0[1]), kB + ci0[0], vB + ci0[0]);
      }
    }
  }
  uint64_t *ks; uint32_t *vs;
  {
    Stage st(c, "sort", (int64_t)n);
    if (chunked) {
      static_assert(NCH == 4, "merge tree of sfc_sort_soa_async is written for four chunks");
      merge_pairs(c, kA + ci0[2], vA + ci0[2], (int64_t)(ci1[2] - ci0[2]), kA + ci0[3], vA + ci0[3], (int64_t)(ci1[3] - ci0[3]), kB + ci0[2], vB + ci0[2]);
      merge_pairs(c, kB, vB, (int64_t)ci0[2], kB + ci0[2], vB + ci0[2], (int64_t)(n - ci0[2]), kA, vA);
      ks = kA; vs = vA;
    } else
    sort_keys63(c, kA, vA, kB, vB, n, &ks, &vs);
    if (ks != c->keys) {
      CUDA_CHECK(cudaMemcpyAsync(c->keys, ks, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
      CUDA_CHECK(cudaMemcpyAsync(c->order, vs, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  const unsigned nb = (unsigned)((n + 255) / 256);
  {
    Stage st(c, "gather", (int64_t)n);
    CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_copy[NCH], 0));                 // weights
    if (n) LAUNCH(c, k_gather_pos, nb, 256, 0, c->in_pos, c->in_w, c->order, n, c->pos4);
  }
  // momentum gather on the copy stream, behind its copy and behind the sorted order
  CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
  CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
  if (n) { k_gather_mom<<<nb, 256, 0, c->copy_stream>>>(c->in_mom, c->in_u, c->order, n, c->mom4); c->n_launches++; CUDA_CHECK(cudaGetLastError()); }
  CUDA_CHECK(cudaEventRecord(c->ev_mom, c->copy_stream));
  c->mom_pending = true;
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  k1.release(); v1.release();
}

__global__ void k_keys_pos4(const float4 *__restrict__ pos4, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pos4[i];
  keys[i] = hilbert_key_pos(p.x, p.y, p.z, 21);
  idx[i]  = (uint32_t)i;
}
__global__ void __launch_bounds__(KT_THREADS) k_keys_pos4_tab(const float4 *__restrict__ pos4, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx)
{
  __shared__ __align__(16) uint16_t tab[12 * 512];
  for (int i = threadIdx.x; i < 12 * 512 / 8; i += KT_THREADS) reinterpret_cast<uint4 *>(tab)[i] = reinterpret_cast<const uint4 *>(g_hil_tab3)[i];
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)KT_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * KT_THREADS) {
    const float4 p = pos4[i];
    keys[i] = hilbert_key_pos21_tab(p.x, p.y, p.z, tab);
    idx[i] = (uint32_t)i;
  }
}
__global__ void k_gather4(const float4 *__restrict__ pin, const float4 *__restrict__ min_, const uint32_t *__restrict__ order, uint64_t n,
                          float4 *__restrict__ pos4, float4 *__restrict__ mom4)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o = order[i];
  pos4[i] = pin[o]; mom4[i] = min_[o];
}

__global__ void k_map_u32(const uint32_t *__restrict__ idx, const uint32_t *__restrict__ table, uint64_t n, uint32_t *__restrict__ out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = table[idx[i]];
}

// 21-bit keys of unsorted float3 positions on the device (slab.cu: block histogram of what a rank has read)
void sfc_keys_f3(ahfgpu_ctx *c, const float *pos3, uint64_t n, uint64_t *keys)
{
  if (n) LAUNCH(c, k_keys_soa_tab, keys_tab_grid(n), KT_THREADS, 0, pos3, (uint64_t)0, n, keys, (uint32_t *)nullptr);
}

void sfc_sort_device4_gid(ahfgpu_ctx *c, const void *pos4_dev, const void *mom4_dev, const uint32_t *gid, uint64_t n, bool has_w, bool has_u);
void sfc_sort_device4(ahfgpu_ctx *c, const void *pos4_dev, const void *mom4_dev, uint64_t n, bool has_w, bool has_u)
{
  sfc_sort_device4_gid(c, pos4_dev, mom4_dev, nullptr, n, has_w, has_u);
}
// gid (may be null): caller's identifier of input particle i; the resident `order` array then holds gid of the sorted particles
// instead of their input position (slab.cu: global input index of particles that came from several ranks)
void sfc_sort_device4_gid(ahfgpu_ctx *c, const void *pos4_dev, const void *mom4_dev, const uint32_t *gid, uint64_t n, bool has_w, bool has_u)
{
  if (n >= (1ull << 32)) AHF_FAIL("more than 2^32-1 particles per device are not supported");
  alloc_particles(c, n);
  c->has_weight = has_w; c->has_u = has_u;
  DevBuf<uint64_t> k0, k1;
  DevBuf<uint32_t> v0, v1;
  k0.reserve(n); k1.reserve(n); v0.reserve(n); v1.reserve(n);
  const unsigned nb = (unsigned)((n + 255) / 256);
  {
    Stage st(c, "keys", (int64_t)n);
    if (n) LAUNCH(c, k_keys_pos4_tab, keys_tab_grid(n), KT_THREADS, 0, (const float4 *)pos4_dev, n, k0.p, v0.p);
  }
  uint64_t *ks; uint32_t *vs;
  {
    Stage st(c, "sort", (int64_t)n);
    sort_keys63(c, k0.p, v0.p, k1.p, v1.p, n, &ks, &vs);
  }
  {
    Stage st(c, "gather", (int64_t)n);
    if (n) LAUNCH(c, k_gather4, nb, 256, 0, (const float4 *)pos4_dev, (const float4 *)mom4_dev, vs, n, c->pos4, c->mom4);
  }
  c->keys = static_cast<decltype(c->keys)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint64_t)));
  c->order = static_cast<decltype(c->order)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint32_t)));
  CUDA_CHECK(cudaMemcpyAsync(c->keys, ks, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
  if (gid && n) LAUNCH(c, k_map_u32, nb, 256, 0, vs, gid, n, c->order);
  else CUDA_CHECK(cudaMemcpyAsync(c->order, vs, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  k0.release(); k1.release(); v0.release(); v1.release();
}

void sfc_sort_aos(ahfgpu_ctx *c, void *part, uint64_t n, uint32_t stride, int off_pos, int off_mom, int off_key, int off_id,
                  int off_w, int off_u)
{
  (void)off_id;
  if (n >= (1ull << 32)) AHF_FAIL("more than 2^32-1 particles per device are not supported");
  if (stride % 8 != 0) AHF_FAIL("particle record stride must be a multiple of 8 bytes");
  alloc_particles(c, n);
  c->has_weight = off_w >= 0; c->has_u = off_u >= 0;
  DevBuf<unsigned char> in, out;
  DevBuf<uint64_t>      k0, k1;
  DevBuf<uint32_t>      v0, v1;
  in.reserve(n * stride); out.reserve(n * stride); k0.reserve(n); k1.reserve(n); v0.reserve(n); v1.reserve(n);
  {
    Stage st(c, "h2d", (int64_t)(n * stride));
    CUDA_CHECK(cudaMemcpyAsync(in.p, part, n * stride, cudaMemcpyHostToDevice, c->stream));
  }
  const unsigned nb = (unsigned)((n + 255) / 256);
  {
    Stage st(c, "keys", (int64_t)n);
    if (n) LAUNCH(c, k_keys_aos, nb, 256, 0, in.p, n, stride, off_pos, k0.p, v0.p);
  }
  uint64_t *ks; uint32_t *vs;
  {
    Stage st(c, "sort", (int64_t)n);
    sort_keys63(c, k0.p, v0.p, k1.p, v1.p, n, &ks, &vs);
  }
  {
    Stage st(c, "gather", (int64_t)n);
    if (n) LAUNCH(c, k_gather_aos, nb, 256, 0, in.p, out.p, vs, ks, n, stride, off_pos, off_mom, off_key, off_w, off_u, c->pos4, c->mom4);
  }
  c->keys = static_cast<decltype(c->keys)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint64_t)));
  c->order = static_cast<decltype(c->order)>(ahf::cache_alloc((n ? n : 1) * sizeof(uint32_t)));
  CUDA_CHECK(cudaMemcpyAsync(c->keys, ks, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(c->order, vs, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  {
    Stage st(c, "d2h", (int64_t)(n * stride));
    CUDA_CHECK(cudaMemcpyAsync(part, out.p, n * stride, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  in.release(); out.release(); k0.release(); k1.release(); v0.release(); v1.release();
}

}  // namespace ahf

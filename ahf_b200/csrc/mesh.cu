// mesh.cu -- D/F/R/L: TSC number-density deposit, refinement flags, next-level construction, relink, level loop.
//
// Replaces gen_domgrids / ll / zero_dens / assign_npart / refine_grid / relink / gen_AMRhierarchy of the reference
// (src/libamr_serial/*.c, called from src/main.c:616-648).  A level is the (z,y,x)-sorted list of its cell keys plus
// one bit per cell (`xbreak`: the reference's x-run ends after this cell, refine_grid.c:231-250), an open-addressing
// hash for coordinate lookup and a per-cell table of the 27 neighbours the reference's search would see
// (get_nnodes.c:459-862).  The domain level is a dense periodic L^3 block addressed arithmetically.
// HBM-bound integer/float scatter work: no tensor cores.
#include "common.cuh"
#include "scan.cuh"
#include "patches.cuh"
#include "hilbert.cuh"
#include "comm.cuh"

namespace ahf {

constexpr int    MIN_NNODES = 125;     // src/param.h:54
constexpr double CRITMULTI  = 8.0;     // src/param.h:118
// Deposit accumulators are u64 fixed point with a per-level scale 2^S: S = min(31 + ceil(log2(masstopartdens)), 44), so that the
// quantum seen by `dens` (masstopartdens * 2^-S) stays ~2^-31 on every level while a cell can still hold 2^20 particles.
static int fx_shift_for(double m2d) { int e = 0; while ((double)(1ull << e) < m2d && e < 40) e++; int S = 31 + e; return S > 44 ? 44 : S; }

// ------------------------------------------------------------------------------------------------
// device view of a level
// ------------------------------------------------------------------------------------------------
struct LV {
  long long       L;
  int             ncell, dense, logL;
  const uint64_t *ckey;
  const uint8_t  *xbreak;
  const uint64_t *hkey;
  const uint2    *hval;          // per slot: index of the block's first existing cell, 8-bit occupancy mask
  uint64_t        hmask;
  const int32_t  *crow, *row_c0; // rows of the level (build_rows_planes); used to decode the neighbour table on the periodic faces
};

static LV view(const Level &l)
{
  LV v; v.L = l.L; v.ncell = (int)l.ncell; v.dense = l.dense ? 1 : 0;
  v.logL = 0; while ((1ll << v.logL) < l.L) v.logL++;
  v.ckey = l.ckey; v.xbreak = l.xbreak; v.hkey = l.hkey; v.hval = reinterpret_cast<const uint2 *>(l.hval); v.hmask = l.hmask; v.crow = l.crow; v.row_c0 = l.row_c0;
  return v;
}

__device__ __forceinline__ uint64_t mix64(uint64_t k)
{
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
__device__ __forceinline__ void lv_coords(const LV &v, int c, int &x, int &y, int &z)
{
  uint64_t k = v.dense ? (uint64_t)c : v.ckey[c];
  x = (int)(k & (uint64_t)(v.L - 1)); y = (int)((k >> v.logL) & (uint64_t)(v.L - 1)); z = (int)(k >> (2 * v.logL));
}
__device__ __forceinline__ uint64_t lv_key(const LV &v, int x, int y, int z)
{
  return (((uint64_t)z << v.logL) | (uint64_t)y) << v.logL | (uint64_t)x;
}
// geometric lookup without periodic wrap: -1 when (x,y,z) is not a cell of the level.
// The hash is keyed on blocks of 8 x-consecutive cells (key >> 3), so that the lookups of a warp walking along a row share a few
// sectors instead of touching 32 random ones.  The cells of a block sit in one row, i.e. they are consecutive in the sorted cell
// array: a slot holds the block key (8 B), the index of the block's first existing cell and an occupancy mask (8 B) -- 16 bytes
// per slot to clear and probe instead of 40 with eight explicit indices.
__device__ __forceinline__ int slot_cell(const uint2 bm, unsigned b)
{
  return ((bm.y >> b) & 1u) ? (int)bm.x + __popc(bm.y & ((1u << b) - 1u)) : -1;
}
__device__ __forceinline__ int lv_lookup(const LV &v, int x, int y, int z)
{
  if ((unsigned)x >= (unsigned)v.L || (unsigned)y >= (unsigned)v.L || (unsigned)z >= (unsigned)v.L) return -1;
  uint64_t k = lv_key(v, x, y, z);
  if (v.dense) return (int)k;
  const uint64_t kb = k >> 3;
  uint64_t s = mix64(kb) & v.hmask;
  for (;;) {
    uint64_t hk = v.hkey[s];
    if (hk == kb) return slot_cell(v.hval[s], (unsigned)(k & 7));
    if (hk == ~0ull) return -1;
    s = (s + 1) & v.hmask;
  }
}

// Neighbour table of a sparse level, 10 words per cell, term-major [10][ncell]: words 0..8 = the (x, y+j-1, z+k-1) cell of row
// q = 3k + j as the reference's search sees it (-1: not visible), word 9 = visibility bits of the x-1 (bit q) and x+1 (bit 9+q)
// neighbours of those rows.  A visible x-1 / x+1 neighbour IS the array neighbour of the row's centre cell (same row, consecutive x)
// -- except across the periodic faces, where it is the last / first cell of that row.  40 bytes per cell instead of 27 explicit
// indices (108): the table is the largest array of a level and the cost of building it is the bytes it writes.
__device__ __forceinline__ int nb_get(const LV &v, const int32_t *__restrict__ nbr, int c, int q, int a, int x)
{
  const int rm = nbr[(size_t)q * (size_t)v.ncell + (size_t)c];
  if (a == 1 || rm < 0) return rm;
  const uint32_t m = (uint32_t)nbr[(size_t)9 * (size_t)v.ncell + (size_t)c];
  if (a == 0) return ((m >> q) & 1u) ? (x > 0 ? rm - 1 : v.row_c0[v.crow[rm] + 1] - 1) : -1;
  return ((m >> (9 + q)) & 1u) ? (x < (int)v.L - 1 ? rm + 1 : v.row_c0[v.crow[rm]]) : -1;
}

// ------------------------------------------------------------------------------------------------
// D2: ll() -- domain cell of every particle (lltools.c:59-66)
// ------------------------------------------------------------------------------------------------
__global__ void k_domain_cells(const float4 *__restrict__ pos4, uint64_t n, int L, int logL, int32_t *__restrict__ pcell)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pos4[i];
  // (unsigned long)((double)L * pos): L is a power of two, the product is exact in float as well
  long long cx = (long long)((double)L * (double)p.x), cy = (long long)((double)L * (double)p.y), cz = (long long)((double)L * (double)p.z);
  if (cx > L - 1 || cx < 0) cx = 0;
  if (cy > L - 1 || cy < 0) cy = 0;
  if (cz > L - 1 || cz < 0) cz = 0;
  pcell[i] = (int32_t)((((cz << logL) | cy) << logL) | cx);
}

// ------------------------------------------------------------------------------------------------
// D4 generic deposit (small levels): one thread per particle, 27 u64 fixed-point global reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tsc_weights(double s, double w[3])
{
  w[0] = 0.5 * (0.5 - s) * (0.5 - s);     // density.c:357-359
  w[1] = 0.75 - s * s;
  w[2] = 0.5 * (0.5 + s) * (0.5 + s);
}

__device__ __forceinline__ double sep_cell(float pos, int i, double L)
{
  double s = (double)pos * L - ((double)i + 0.5);
  if (fabs(s) > 0.5 * L) s -= copysign(L, s);   // density.c:347-354 periodic image
  return s;
}

__global__ void __launch_bounds__(256) k_deposit_generic(const float4 *__restrict__ pos4, const uint32_t *__restrict__ plist,
                                                         const int32_t *__restrict__ pcell, uint64_t np, LV v,
                                                         const int32_t *__restrict__ nbr, unsigned long long *__restrict__ acc, double fxscale)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const bool valid = i < np;
  int    c = -1;
  double wx[3], wy[3], wz[3];
  int    cx = 0, cy = 0, cz = 0;
  if (valid) {
    uint64_t p = plist ? plist[i] : i;
    float4   q = pos4[p];
    c = pcell[i];
    lv_coords(v, c, cx, cy, cz);
    const double L = (double)v.L;
    tsc_weights(sep_cell(q.x, cx, L), wx);
    tsc_weights(sep_cell(q.y, cy, L), wy);
    tsc_weights(sep_cell(q.z, cz, L), wz);
  }
  if (!valid) return;
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const unsigned long long t = (unsigned long long)(wz[k] * wy[j] * wx[a] * fxscale + 0.5);
        int tgt;
        if (v.dense) {
          int x = (cx + a - 1) & (int)(v.L - 1), y = (cy + j - 1) & (int)(v.L - 1), z = (cz + k - 1) & (int)(v.L - 1);
          tgt = (int)lv_key(v, x, y, z);
        } else tgt = nb_get(v, nbr, c, k * 3 + j, a, cx);
        if (tgt >= 0) atomicAdd(&acc[tgt], t);
      }
}

// ------------------------------------------------------------------------------------------------
// D4 domain level: shared-memory tile deposit.
//   The particle array is Hilbert sorted, so the particles of every aligned T^3-cell cube ("tile") are one contiguous
//   range.  A CTA takes one chunk (<= DT_CHUNK particles) of one tile: the chunk's float4 positions are staged into
//   shared memory by the TMA (cp.async.bulk, mbarrier completion), the (T+2)^3 tile is accumulated with native u32
//   shared-memory atomics in 2^-21 fixed point (ATOMS.ADD; float shared atomics would be CAS loops), and flushed as
//   2^-40 fixed point u64 into the level accumulators: plain stores for the cells no other CTA can touch, REDG.ADD.64
//   for the rest.  Integer accumulation makes the result independent of any ordering.
// ------------------------------------------------------------------------------------------------
constexpr int DT_T       = 16;
constexpr int DT_H       = DT_T + 2;
constexpr int DT_HH      = DT_H * DT_H * DT_H;
constexpr int DT_CHUNK   = 8192;         // particles per CTA (carry words make any count safe)
constexpr int DT_SUB     = 512;          // particles per TMA stage = one per thread
constexpr int DT_THREADS = 512;
constexpr int DT_SMEM    = 2 * DT_SUB * 16 + 2 * DT_HH * 4 + 16;

__global__ void k_tile_starts(const uint64_t *__restrict__ keys, int64_t n, int tbits, int ntile, int32_t *__restrict__ tstart)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntile) return;
  if (t == ntile) { tstart[t] = (int32_t)n; return; }
  const uint64_t kmin = (uint64_t)t << (3 * (21 - tbits));
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < kmin) lo = mid + 1; else hi = mid; }
  tstart[t] = (int32_t)lo;
}
__global__ void k_tile_nchunk(const int32_t *__restrict__ tstart, int ntile, int *__restrict__ nchunk)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntile) return;
  nchunk[t] = (tstart[t + 1] - tstart[t] + DT_CHUNK - 1) / DT_CHUNK;
}
__global__ void k_tile_work(const int *__restrict__ nchunk, const int *__restrict__ woff, int ntile, int2 *__restrict__ work)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntile) return;
  const int nc = nchunk[t], o = woff[t];
  for (int q = 0; q < nc; q++) work[o + q] = make_int2(t, q | (nc == 1 ? 0x40000000 : 0));
}

// dense domain level: self-contained work items {first particle, count, tile origin (10 bits per axis), flags}
__global__ void k_tile_work4(const int32_t *__restrict__ tstart, const int *__restrict__ nchunk, const int *__restrict__ woff, int ntile, int tbits,
                             int4 *__restrict__ work)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntile) return;
  const int nc = nchunk[t], o = woff[t], a = tstart[t], b = tstart[t + 1];
  uint32_t tx, ty, tz;
  hilbert_coords((uint64_t)t, (unsigned)tbits, tx, ty, tz);
  for (int q = 0; q < nc; q++) work[o + q] = make_int4(a + q * DT_CHUNK, min(DT_CHUNK, b - a - q * DT_CHUNK), (int)(tx | (ty << 10) | (tz << 20)), nc == 1 ? 1 : 0);
}

// functors of k_seg_heads (scan.cuh): deposit tiles of a refinement level, x-rows of a level's cells, z-planes of its rows
struct TileSeg {
  const uint64_t *keys; const uint32_t *plist; int sh; uint64_t np; uint32_t *tlist; int32_t *tstart; const int *np_dev;
  __device__ uint64_t key(uint64_t i) const { return keys[plist[i]] >> sh; }
  __device__ void emit(uint64_t i, int seg, int head, uint64_t k) const { if (head) { tlist[seg] = (uint32_t)k; tstart[seg] = (int32_t)i; } }
  __device__ void end(int nseg) const { tstart[nseg] = np_dev ? (int32_t)*np_dev : (int32_t)np; }
};
struct RowSeg {
  const uint64_t *ckey; int logL; int ncell; int32_t *crow; uint64_t *rowkey; int32_t *row_c0;
  __device__ uint64_t key(uint64_t i) const { return ckey[i] >> logL; }
  __device__ void emit(uint64_t i, int seg, int head, uint64_t k) const { crow[i] = seg; if (head) { rowkey[seg] = k; row_c0[seg] = (int32_t)i; } }
  __device__ void end(int nseg) const { row_c0[nseg] = ncell; }
};
struct PlaneSeg {
  const uint64_t *rowkey; int logL; int nrow; int32_t *rowplane; int32_t *plane_r0; int32_t *pz;
  __device__ uint64_t key(uint64_t i) const { return rowkey[i] >> logL; }
  __device__ void emit(uint64_t i, int seg, int head, uint64_t k) const { rowplane[i] = seg; if (head) { plane_r0[seg] = (int32_t)i; pz[seg] = (int32_t)k; } }
  __device__ void end(int nseg) const { plane_r0[nseg] = nrow; }
};

// refinement levels: tile id (Hilbert prefix of the particle key) of every level particle, heads of equal-id runs
__global__ void k_lvl_tile_heads(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ plist, uint64_t np, int sh, uint8_t *__restrict__ head)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= np) return;
  head[i] = (i == 0 || (keys[plist[i]] >> sh) != (keys[plist[i - 1]] >> sh)) ? 1 : 0;
}
__global__ void k_lvl_tile_fill(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ plist, uint64_t np, int sh, const uint8_t *__restrict__ head,
                                const int *__restrict__ hs, int ntile, uint32_t *__restrict__ tlist, int32_t *__restrict__ tstart)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= np) return;
  if (head[i]) { tlist[hs[i]] = (uint32_t)(keys[plist[i]] >> sh); tstart[hs[i]] = (int32_t)i; }
  if (i == np - 1) tstart[ntile] = (int32_t)np;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tile word update: low word += v with a native shared atomic; the (rare) carry goes to the carry word through a
// PREDICATED red.shared -- no branch, no convergence barrier in the hot loop
__device__ __forceinline__ void tile_add32(uint32_t saddr, uint32_t v)
{
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
  const uint32_t sum = old + v;
  asm volatile("{\n .reg .pred p;\n setp.lt.u32 p, %0, %1;\n @p red.shared.add.u32 [%2], 1;\n}" ::"r"(sum), "r"(old), "r"(saddr + DT_HH * 4) : "memory");
}
__device__ __forceinline__ void tile_add64(uint32_t saddr, unsigned long long v)
{
  const uint32_t lo = (uint32_t)v;
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(lo) : "memory");
  const uint32_t hi = (uint32_t)(v >> 32) + ((old + lo < old) ? 1u : 0u);
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %0, 0;\n @p red.shared.add.u32 [%1], %0;\n}" ::"r"(hi), "r"(saddr + DT_HH * 4) : "memory");
}

// work list of the deposit kernels in one launch (k_scan_emit): chunks per tile, their prefix sum, the list entries
struct TileWork {
  const int32_t *tstart; int2 *work;
  __device__ int  value(uint64_t t) const { return (tstart[t + 1] - tstart[t] + DT_CHUNK - 1) / DT_CHUNK; }
  __device__ void emit(uint64_t t, int o, int nc) const { for (int q = 0; q < nc; q++) work[o + q] = make_int2((int)t, q | (nc == 1 ? 0x40000000 : 0)); }
  __device__ void end(int) const {}
};
struct TileWork4 {
  const int32_t *tstart; int tbits; int4 *work;
  __device__ int  value(uint64_t t) const { return (tstart[t + 1] - tstart[t] + DT_CHUNK - 1) / DT_CHUNK; }
  __device__ void emit(uint64_t t, int o, int nc) const
  {
    const int a = tstart[t], b = tstart[t + 1];
    uint32_t tx, ty, tz;
    hilbert_coords((uint64_t)t, (unsigned)tbits, tx, ty, tz);
    for (int q = 0; q < nc; q++) work[o + q] = make_int4(a + q * DT_CHUNK, min(DT_CHUNK, b - a - q * DT_CHUNK), (int)(tx | (ty << 10) | (tz << 20)), nc == 1 ? 1 : 0);
  }
  __device__ void end(int) const {}
};

template <bool SPARSE>
__global__ void __launch_bounds__(DT_THREADS, 3)
k_deposit_tiles(const float4 *__restrict__ pos4, const int32_t *__restrict__ tstart, const int2 *__restrict__ work, int L, int logL, int tbits,
                unsigned long long *__restrict__ acc, const uint32_t *__restrict__ tlist, const int32_t *__restrict__ pcell, LV lvw,
                const int32_t *__restrict__ nbr, float fxs, const int *__restrict__ Wp)
{
  if ((int)blockIdx.x >= *Wp) return;                  // the grid is an upper bound of the work list (no host read-back of its length)
  extern __shared__ __align__(16) unsigned char dsm[];
  float4   *sp   = reinterpret_cast<float4 *>(dsm);                            // two stages of DT_SUB positions
  uint32_t *tile = reinterpret_cast<uint32_t *>(dsm + 2 * DT_SUB * 16);        // low 32 bits of the fixed-point sums
  uint32_t *tcar = tile + DT_HH;                                               // number of wrap-arounds of the low word
  uint64_t *mbar = reinterpret_cast<uint64_t *>(dsm + 2 * DT_SUB * 16 + 2 * DT_HH * 4);   // two mbarriers
  const int2 wk = work[blockIdx.x];
  const int  t = wk.x, chunk = wk.y & 0x3fffffff;
  const bool sole = (wk.y & 0x40000000) != 0;        // the tile's only chunk: its inner cells are touched by nobody else
  const int  s0 = tstart[t] + chunk * DT_CHUNK;
  int        np = tstart[t + 1] - s0;
  if (np > DT_CHUNK) np = DT_CHUNK;
  uint32_t tx, ty, tz;
  hilbert_coords(SPARSE ? (uint64_t)tlist[t] : (uint64_t)t, (unsigned)tbits, tx, ty, tz);
  const int x0 = (int)tx * DT_T, y0 = (int)ty * DT_T, z0 = (int)tz * DT_T;
  const uint32_t bar = smem_u32(mbar), dst = smem_u32(sp), tile_s = smem_u32(tile);
  const int nsub = (np + DT_SUB - 1) / DT_SUB;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // TMA stage `sub` -> buffer sub&1 (cp.async.bulk global->shared, completion on the stage's mbarrier)
  auto issue = [&](int sub) {
    const int      cnt = min(DT_SUB, np - sub * DT_SUB);
    const uint32_t bytes = (uint32_t)cnt * 16u, bb = bar + 8u * (sub & 1), dd = dst + (uint32_t)(sub & 1) * DT_SUB * 16u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dd), "l"(pos4 + s0 + (size_t)sub * DT_SUB), "r"(bytes), "r"(bb) : "memory");
  };
  if (threadIdx.x == 0) issue(0);
  for (int i = threadIdx.x; i < 2 * DT_HH / 4; i += DT_THREADS) reinterpret_cast<uint4 *>(tile)[i] = make_uint4(0, 0, 0, 0);  // overlaps the first copy
  __syncthreads();
  const float fL = (float)L;
  const int   M = L - 1;
  const int   lane = threadIdx.x & 31;
  for (int sub = 0; sub < nsub; sub++) {
    if (threadIdx.x == 0 && sub + 1 < nsub) issue(sub + 1);                    // the other buffer was released by the barrier below
    {
      const uint32_t bb = bar + 8u * (sub & 1), parity = (uint32_t)(sub >> 1) & 1u;
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bb), "r"(parity) : "memory");
      }
    }
    const int  i = sub * DT_SUB + threadIdx.x;
    const bool valid = i < np;
    const float4 q = valid ? sp[(sub & 1) * DT_SUB + threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
    // ll(): cell = (unsigned long)(L * pos), out-of-range -> 0 (lltools.c:59-66); exact in float for power-of-two L
    const float fx = q.x * fL, fy = q.y * fL, fz = q.z * fL;
    int cx = (int)fx, cy = (int)fy, cz = (int)fz;
    if (SPARSE) {          // the node the particle is linked to (relink's inclusive faces: floor(L*x) - face bit, carried in .w)
      const int fb = __float_as_int(q.w);
      cx -= fb & 1; cy -= (fb >> 1) & 1; cz -= (fb >> 2) & 1;
    }
    float sx = fx - ((float)cx + 0.5f), sy = fy - ((float)cy + 0.5f), sz = fz - ((float)cz + 0.5f);
    if (!SPARSE) {
      if (cx > M) { cx = 0; sx = fx - 0.5f - fL; }
      if (cy > M) { cy = 0; sy = fy - 0.5f - fL; }
      if (cz > M) { cz = 0; sz = fz - 0.5f - fL; }
    }
    float wx[3], wy[3], wz[3];
    wx[0] = 0.5f * (0.5f - sx) * (0.5f - sx); wx[1] = 0.75f - sx * sx; wx[2] = 0.5f * (0.5f + sx) * (0.5f + sx);
    wy[0] = 0.5f * (0.5f - sy) * (0.5f - sy); wy[1] = 0.75f - sy * sy; wy[2] = 0.5f * (0.5f + sy) * (0.5f + sy);
    wz[0] = 0.5f * (0.5f - sz) * (0.5f - sz); wz[1] = 0.75f - sz * sz; wz[2] = 0.5f * (0.5f + sz) * (0.5f + sz);
    const int lx = cx - x0, ly = cy - y0, lz = cz - z0;         // 0..T-1 inside the tile
    const bool intile = (unsigned)lx < (unsigned)DT_T && (unsigned)ly < (unsigned)DT_T && (unsigned)lz < (unsigned)DT_T;
    const int  cid = valid ? (intile ? ((lz * DT_T + ly) * DT_T + lx) : -2) : -1;
    // clump cores: a whole warp in ONE cell -> sum the 27 terms across the warp (REDUX on three 16-bit limbs) and let one
    // lane issue the shared-memory atomics instead of 32 lanes serialising on the same address.  (Grouping partial
    // matches with match.any was measured: the MATCH + partial-mask REDUX cost more than the conflicts they remove.)
    const int  cid0 = __shfl_sync(0xffffffffu, cid, 0);
    const bool grouped = __all_sync(0xffffffffu, cid == cid0) && cid0 >= 0;
    const unsigned peers = 0xffffffffu;
    const int      leader = 0;
    const uint32_t cell0 = tile_s + 4u * (uint32_t)((lz * DT_H + ly) * DT_H + lx);     // shared address of the (k=0,j=0,a=0) word
    if (grouped) {
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float wyz = wz[k] * wy[j] * fxs;
#pragma unroll
          for (int a = 0; a < 3; a++) {
            const unsigned long long v = __float2ull_rn(wyz * wx[a]);          // < 2^43
            const uint32_t l0 = __reduce_add_sync(peers, (uint32_t)(v & 0xffffu)), l1 = __reduce_add_sync(peers, (uint32_t)((v >> 16) & 0xffffu)),
                           l2 = __reduce_add_sync(peers, (uint32_t)(v >> 32));
            if (lane == leader && cid >= 0) tile_add64(cell0 + 4u * (uint32_t)((k * DT_H + j) * DT_H + a),
                                      (unsigned long long)l0 + ((unsigned long long)l1 << 16) + ((unsigned long long)l2 << 32));
          }
        }
    } else if (valid && intile) {
      if (!SPARSE || fxs <= 4294967296.0f) {         // S <= 32 (domain and first level): every term fits the low word
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const float wyz = wz[k] * wy[j] * fxs;
#pragma unroll
            for (int a = 0; a < 3; a++) tile_add32(cell0 + 4u * (uint32_t)((k * DT_H + j) * DT_H + a), __float2uint_rn(wyz * wx[a]));
          }
      } else {
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const float wyz = wz[k] * wy[j] * fxs;
#pragma unroll
            for (int a = 0; a < 3; a++) tile_add64(cell0 + 4u * (uint32_t)((k * DT_H + j) * DT_H + a), __float2ull_rn(wyz * wx[a]));
          }
      }
    }
    if (valid && !intile) {
      // the particle's cell is not in this tile (coordinate clamp of ll()): straight to the global accumulators
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
          for (int a = 0; a < 3; a++) {
            long long tgt;
            if (SPARSE) tgt = nb_get(lvw, nbr, pcell[s0 + i], k * 3 + j, a, cx);
            else { const int x = (cx + a - 1) & M, y = (cy + j - 1) & M, z = (cz + k - 1) & M; tgt = (long long)((((size_t)z << logL | y) << logL) | x); }
            if (tgt >= 0) atomicAdd(&acc[tgt], __float2ull_rn(wz[k] * wy[j] * fxs * wx[a]));
          }
    }
    __syncthreads();                                   // everybody is done with this stage's buffer
  }
  if (SPARSE) {
    // flush of a refinement-level tile: the touched cells are first compacted into a list (the TMA stage buffers are free now),
    // then ALL threads resolve them through the cell hash -- a thread sees a few independent lookups instead of marching
    // through 18 dependent ones
    __shared__ int s_nnz;
    uint16_t *list = reinterpret_cast<uint16_t *>(dsm);                     // DT_HH entries fit the 16 KB of stage buffers
    if (threadIdx.x == 0) s_nnz = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < DT_HH; i += DT_THREADS) {
      const bool nz = (tile[i] | tcar[i]) != 0;
      const unsigned bal = __ballot_sync(__activemask(), nz);
      if (nz) {
        const int leader = __ffs(bal) - 1;
        int basep = 0;
        if (lane == leader) basep = atomicAdd(&s_nnz, __popc(bal));
        basep = __shfl_sync(bal, basep, leader);
        list[basep + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
      }
    }
    __syncthreads();
    const int nnz = s_nnz;
    // four independent hash probes in flight per thread (the walk is latency bound: key slot -> value -> reduction)
    for (int e0 = threadIdx.x; e0 < nnz; e0 += 4 * DT_THREADS) {
      uint64_t kq[4], sq[4], hq[4];
      int      iq[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int e = e0 + q * DT_THREADS;
        iq[q] = e < nnz ? (int)list[e] : -1;
        const int i = iq[q] < 0 ? 0 : iq[q];
        const int hz = i / (DT_H * DT_H), r = i - hz * (DT_H * DT_H), hy = r / DT_H, hx = r - hy * DT_H;
        kq[q] = lv_key(lvw, (x0 + hx - 1) & M, (y0 + hy - 1) & M, (z0 + hz - 1) & M);
        sq[q] = mix64(kq[q] >> 3) & lvw.hmask;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) hq[q] = iq[q] >= 0 ? lvw.hkey[sq[q]] : ~0ull;
      int tg[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint64_t kb = kq[q] >> 3;
        uint64_t s = sq[q], hk = hq[q];
        while (hk != kb && hk != ~0ull) { s = (s + 1) & lvw.hmask; hk = lvw.hkey[s]; }
        tg[q] = (iq[q] >= 0 && hk == kb) ? slot_cell(lvw.hval[s], (unsigned)(kq[q] & 7)) : -1;    // particles sit on interior nodes: every touched cell exists
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (tg[q] >= 0) atomicAdd(&acc[tg[q]], ((unsigned long long)tcar[iq[q]] << 32) | tile[iq[q]]);      // already in the level's 2^-S units
    }
    return;
  }
  // flush (dense): thread = one (hx,hy) column of the tile, marching in z (no div/mod, constant strides)
  if (threadIdx.x < DT_H * DT_H) {
    const int hx = threadIdx.x % DT_H, hy = threadIdx.x / DT_H;
    const int x = (x0 + hx - 1) & M, y = (y0 + hy - 1) & M;
    const bool inxy = hx >= 2 && hx <= DT_T - 1 && hy >= 2 && hy <= DT_T - 1;
#pragma unroll 2
    for (int hz = 0; hz < DT_H; hz++) {
      const int i = (hz * DT_H + hy) * DT_H + hx;
      const uint32_t v = tile[i], cr = tcar[i];
      if ((v | cr) == 0) continue;
      const int z = (z0 + hz - 1) & M;
      const unsigned long long val = ((unsigned long long)cr << 32) | v;          // already in the level's 2^-S units
      unsigned long long *dstp = &acc[(((size_t)z << logL | y) << logL) | x];
      if (sole && inxy && hz >= 2 && hz <= DT_T - 1) *dstp = val; else atomicAdd(dstp, val);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// D4 refinement levels, run aggregation (k_deposit_runs).
//   On a refinement level the particles are concentrated: several (in clump cores hundreds) per cell, and since the level's
//   particle list is Hilbert sorted the particles of one cell are CONSECUTIVE.  Lane-per-particle shared atomics then serialise
//   on the same address (ncu, level 4 of the bench box: 20 wavefronts per ATOMS).  Here every thread walks DR_K consecutive
//   particles and keeps the 27 fixed-point terms of the current cell in registers; they go to the shared-memory tile only when
//   the cell changes.  The lanes of a warp sit DR_K particles apart, so their flushes mostly hit different cells.  The per-term
//   integers are the ones k_deposit_tiles<true> produces (same float expression, same rounding), so the level sums are
//   bit-identical to it whatever the grouping.
// ------------------------------------------------------------------------------------------------
constexpr int DR_THREADS = 256;
constexpr int DR_K       = 8;                                   // consecutive particles per thread = one 128-byte line of lpos
constexpr int DR_SMEM    = 2 * DT_HH * 4 + DT_HH * 2 + 16;       // low words | carry words | compaction list of the flush

__global__ void __launch_bounds__(DR_THREADS, 3)
k_deposit_runs(const float4 *__restrict__ lpos, const int32_t *__restrict__ tstart, const int2 *__restrict__ work, int L, int logL, int tbits,
               unsigned long long *__restrict__ acc, const uint32_t *__restrict__ tlist, const int32_t *__restrict__ pcell, LV lvw,
               const int32_t *__restrict__ nbr, float fxs, const int *__restrict__ Wp)
{
  if ((int)blockIdx.x >= *Wp) return;                  // the grid is an upper bound of the work list
  extern __shared__ __align__(16) unsigned char dsm[];
  uint32_t *tile = reinterpret_cast<uint32_t *>(dsm);
  uint32_t *tcar = tile + DT_HH;
  uint16_t *list = reinterpret_cast<uint16_t *>(dsm + 2 * DT_HH * 4);
  __shared__ int s_nnz;
  const int2 wk = work[blockIdx.x];
  const int  t = wk.x, chunk = wk.y & 0x3fffffff;
  const int  s0 = tstart[t] + chunk * DT_CHUNK;
  int        np = tstart[t + 1] - s0;
  if (np > DT_CHUNK) np = DT_CHUNK;
  uint32_t tx, ty, tz;
  hilbert_coords((uint64_t)tlist[t], (unsigned)tbits, tx, ty, tz);
  const int x0 = (int)tx * DT_T, y0 = (int)ty * DT_T, z0 = (int)tz * DT_T;
  const uint32_t tile_s = smem_u32(tile);
  for (int i = threadIdx.x; i < 2 * DT_HH / 4; i += DR_THREADS) reinterpret_cast<uint4 *>(tile)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) s_nnz = 0;
  __syncthreads();
  const float fL = (float)L;
  const int   M = L - 1;
  const int   lane = threadIdx.x & 31;
  // segment length: DR_K when the chunk is full, shorter for the thin tiles of deep levels so that all threads have work
  const int kk = min(DR_K, max(1, (np + DR_THREADS - 1) / DR_THREADS));
  for (int base = threadIdx.x * kk; base < np + (31 * kk); base += DR_THREADS * kk) {           // whole warps stay in the loop (REDUX below)
    if ((base - lane * kk) >= np) break;                                                        // the warp's first segment is past the end
    unsigned long long a27[27];
#pragma unroll
    for (int q = 0; q < 27; q++) a27[q] = 0ull;
    int cur = -1;                                    // tile word index of the cell the registers belong to
    auto flush = [&](int widx) {
      const uint32_t cell0 = tile_s + 4u * (uint32_t)widx;
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
          for (int a = 0; a < 3; a++) tile_add64(cell0 + 4u * (uint32_t)((k * DT_H + j) * DT_H + a), a27[k * 9 + j * 3 + a]);
    };
    const int nmine = min(kk, np - base);            // <= 0: nothing (the thread only takes part in the warp votes)
    float4 qn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nmine > 0) qn = lpos[s0 + base];
    for (int i = 0; i < nmine; i++) {
      const float4 q = qn;
      if (i + 1 < nmine) qn = lpos[s0 + base + i + 1];
      const float fx = q.x * fL, fy = q.y * fL, fz = q.z * fL;
      const int   fb = __float_as_int(q.w);          // relink's face bits: cell coordinate = floor(L*x) - bit
      const int   cx = (int)fx - (fb & 1), cy = (int)fy - ((fb >> 1) & 1), cz = (int)fz - ((fb >> 2) & 1);
      const float sx = fx - ((float)cx + 0.5f), sy = fy - ((float)cy + 0.5f), sz = fz - ((float)cz + 0.5f);
      float wx[3], wy[3], wz[3];
      wx[0] = 0.5f * (0.5f - sx) * (0.5f - sx); wx[1] = 0.75f - sx * sx; wx[2] = 0.5f * (0.5f + sx) * (0.5f + sx);
      wy[0] = 0.5f * (0.5f - sy) * (0.5f - sy); wy[1] = 0.75f - sy * sy; wy[2] = 0.5f * (0.5f + sy) * (0.5f + sy);
      wz[0] = 0.5f * (0.5f - sz) * (0.5f - sz); wz[1] = 0.75f - sz * sz; wz[2] = 0.5f * (0.5f + sz) * (0.5f + sz);
      const int lx = cx - x0, ly = cy - y0, lz = cz - z0;
      const bool intile = (unsigned)lx < (unsigned)DT_T && (unsigned)ly < (unsigned)DT_T && (unsigned)lz < (unsigned)DT_T;
      if (!intile) {
        // the particle's node is not in this tile (face case of relink): straight to the global accumulators
        const size_t pc = (size_t)pcell[s0 + base + i];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int j = 0; j < 3; j++)
#pragma unroll
            for (int a = 0; a < 3; a++) {
              const int tgt = nb_get(lvw, nbr, (int)pc, k * 3 + j, a, cx);
              if (tgt >= 0) atomicAdd(&acc[tgt], __float2ull_rn(wz[k] * wy[j] * fxs * wx[a]));
            }
        continue;
      }
      const int widx = (lz * DT_H + ly) * DT_H + lx;
      if (widx != cur) {
        if (cur >= 0) {
          flush(cur);
#pragma unroll
          for (int q = 0; q < 27; q++) a27[q] = 0ull;
        }
        cur = widx;
      }
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float wyz = wz[k] * wy[j] * fxs;
#pragma unroll
          for (int a = 0; a < 3; a++) a27[k * 9 + j * 3 + a] += __float2ull_rn(wyz * wx[a]);
        }
    }
    // last run of the segment.  Clump cores: when the whole warp ends in ONE cell the 27 sums are added across the warp first
    // (REDUX on three limbs, values < 2^48) and one lane touches shared memory.
    const int  cur0 = __shfl_sync(0xffffffffu, cur, 0);
    const bool same = __all_sync(0xffffffffu, cur == cur0) && cur0 >= 0;
    if (same) {
#pragma unroll
      for (int q = 0; q < 27; q++) {
        const unsigned long long v = a27[q];
        const uint32_t l0 = __reduce_add_sync(0xffffffffu, (uint32_t)(v & 0xffffu)), l1 = __reduce_add_sync(0xffffffffu, (uint32_t)((v >> 16) & 0xffffu)),
                       l2 = __reduce_add_sync(0xffffffffu, (uint32_t)(v >> 32));
        a27[q] = (unsigned long long)l0 + ((unsigned long long)l1 << 16) + ((unsigned long long)l2 << 32);
      }
      if (lane == 0) flush(cur0);
    } else if (cur >= 0) flush(cur);
  }
  __syncthreads();
  // flush of the tile: touched cells compacted into a list, then resolved through the cell hash with four probes in flight
  for (int i = threadIdx.x; i < DT_HH; i += DR_THREADS) {
    const bool nz = (tile[i] | tcar[i]) != 0;
    const unsigned bal = __ballot_sync(__activemask(), nz);
    if (nz) {
      const int leader = __ffs(bal) - 1;
      int basep = 0;
      if (lane == leader) basep = atomicAdd(&s_nnz, __popc(bal));
      basep = __shfl_sync(bal, basep, leader);
      list[basep + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
    }
  }
  __syncthreads();
  const int nnz = s_nnz;
  for (int e0 = threadIdx.x; e0 < nnz; e0 += 4 * DR_THREADS) {
    uint64_t kq[4], sq[4], hq[4];
    int      iq[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int e = e0 + q * DR_THREADS;
      iq[q] = e < nnz ? (int)list[e] : -1;
      const int i = iq[q] < 0 ? 0 : iq[q];
      const int hz = i / (DT_H * DT_H), r = i - hz * (DT_H * DT_H), hy = r / DT_H, hx = r - hy * DT_H;
      kq[q] = lv_key(lvw, (x0 + hx - 1) & M, (y0 + hy - 1) & M, (z0 + hz - 1) & M);
      sq[q] = mix64(kq[q] >> 3) & lvw.hmask;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) hq[q] = iq[q] >= 0 ? lvw.hkey[sq[q]] : ~0ull;
    int tg[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint64_t kb = kq[q] >> 3;
      uint64_t s = sq[q], hk = hq[q];
      while (hk != kb && hk != ~0ull) { s = (s + 1) & lvw.hmask; hk = lvw.hkey[s]; }
      tg[q] = (iq[q] >= 0 && hk == kb) ? slot_cell(lvw.hval[s], (unsigned)(kq[q] & 7)) : -1;      // particles sit on interior nodes: every touched cell exists
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (tg[q] >= 0) atomicAdd(&acc[tg[q]], ((unsigned long long)tcar[iq[q]] << 32) | tile[iq[q]]);
  }
}

// ------------------------------------------------------------------------------------------------
// D4 domain level, integer formulation (k_deposit_dom).
//   Positions are float32 in [0,1): u = trunc(x * 2^32) is exact (x >= 2^-8) and cell = u >> (32 - logL) equals the
//   reference's (unsigned long)(L * x) (lltools.c:59).  The in-cell fraction f = (u << logL) / 2^32 gives the TSC weights
//   (density.c:357-359 with s = f - 1/2) in 2^-32 fixed point:  W0 = (1-f)^2/2, W2 = f^2/2, W1 = 1 - W0 - W2.  Products use
//   IMAD.HI and the middle weight of every triple is the complement, so each particle deposits EXACTLY 2^32 units: the
//   level total is N * 2^32 whatever the order (a checksum the tests use).  No float->int conversion in the 27-term loop.
//   Two copies of the tile low words (even / odd lanes, 16 banks apart) halve the same-address serialisation of the
//   shared-memory atomics that Hilbert-adjacent particles of one cell cause.
// ------------------------------------------------------------------------------------------------
constexpr int DD_CO   = DT_HH + 8;                                   // word offset of copy B: == 16 (mod 32 banks)
constexpr int DD_CAR  = 2 * DD_CO;                                   // word offset of the carry counters
constexpr int DD_NS   = 4;                                           // TMA stages of DT_SUB particles (ring)
// Fixed-point scale of the domain deposit: every particle deposits EXACTLY 2^32 units (complement weights).  Measured and rejected in
// round 2 (profiles/r2l_bench_256_n1.json): 2^26 units per particle with the carry detection only where an atomic returned a value
// with its top bit set -- about 80 of the 354 warp instructions per 32 particles less (no carry replay loop, no per-term carry pair)
// -- did NOT change the kernel time (0.344 vs 0.349 ms): the kernel is not bound by instruction issue.
// Round 2b: the experiment k_deposit_dom2 works in 2^28 units per particle (D2_S): the z weights are taken at 28 bits (dom_wz_scale), the
// products stay IMAD.HI, the middle term of every triple stays the complement, so every particle deposits EXACTLY 2^28 units and a 32-bit
// word holds 16 particles' worth.  k_deposit_dom runs the same arithmetic when asked to (AHFGPU_DOM_S=28, implied by AHFGPU_DOM_V2=1): then
// both kernels add the same integers.  It is NOT the default: truncation errors have a sign (edge terms down, middle terms up), so next to a
// cell with 10^4 particles a neighbour's sum is off by ~10^4 units -- 1.1e-6 of its own density at 2^-28 (measured on the 128^3 box against
// the float64 sum; the north star's 1e-5 still holds), 16 times less at 2^-32.
constexpr int DD_S    = 32;
constexpr int D2_S    = 28;
constexpr int DD_SMEM = DD_NS * DT_SUB * 16 + (DD_CAR + DT_HH) * 4 + 8 * DD_NS * (DT_THREADS / 32);   // DT_SUB*16 = 16 warps x 512 B per ring slot

__device__ __forceinline__ uint32_t pos_q32(float x) { return x >= 1.0f ? 0u : __float2uint_rz(x * 4294967296.0f); }   // x == 1 -> cell 0 (lltools.c:61-64)
__device__ __forceinline__ void tsc_q32(uint32_t t32, uint32_t (&w)[3])
{
  const uint32_t t31 = t32 >> 1, a31 = 0x80000000u - t31;
  w[0] = (uint32_t)(((unsigned long long)a31 * a31) >> 31);
  w[2] = (uint32_t)(((unsigned long long)t31 * t31) >> 31);
  w[1] = 0u - w[0] - w[2];
}
__device__ __forceinline__ void dom_wz_scale(uint32_t (&wz)[3], const int wsh)      // z weights in 2^-(32 - wsh): the three still add up to exactly one particle
{
  if (wsh) { wz[0] >>= wsh; wz[2] >>= wsh; wz[1] = (1u << (32 - wsh)) - wz[0] - wz[2]; }
}
// low word += v (native ATOMS.ADD, old value returned).  Whether the low word wrapped is the carry-out of old + v; dom_carry
// shifts it into a per-thread bit mask (IADD3 with carry-out + IADD3.X: two instructions, no predicate, no branch).  The
// atomics of a plane are all issued before the first result is used; the few set bits are replayed onto the carry counters
// after the 27 terms.
__device__ __forceinline__ uint32_t dom_atom(uint32_t lo_addr, uint32_t v)
{
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(lo_addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void dom_carry(uint32_t &mask, uint32_t old, uint32_t v)
{
  asm volatile("{\n .reg .u32 s;\n add.cc.u32 s, %1, %2;\n addc.u32 %0, %0, %0;\n}" : "+r"(mask) : "r"(old), "r"(v));
}
__constant__ uint16_t c_dom_off[27];     // byte offset of term (k,j,a) inside the (T+2)^3 tile, indexed k*9+j*3+a

template <int VAR>      // 0 = product; 1..4 = timing experiments (AHFGPU_DOM_VARIANT): 1 no return/carry, 2 no atomics, 3 no flush, 4 one copy
__global__ void __launch_bounds__(DT_THREADS, 2)
k_deposit_dom(const float4 *__restrict__ pos4, const int4 *__restrict__ work, const int *__restrict__ Wp, int L, int logL,
              unsigned long long *__restrict__ acc, const uint32_t one /* == 1, a run-time value on purpose: see dom_carry */, const int rmax, const int wsh)
{
  const int W = *Wp;                                   // the grid is an upper bound of the work list (no host read-back of its length)
  if ((int)blockIdx.x >= W) return;
  extern __shared__ __align__(16) unsigned char dsm[];
  float4   *sp   = reinterpret_cast<float4 *>(dsm);
  uint32_t *tile = reinterpret_cast<uint32_t *>(dsm + DD_NS * DT_SUB * 16);        // copy A | copy B | carry counters
  uint64_t *mbar = reinterpret_cast<uint64_t *>(dsm + DD_NS * DT_SUB * 16 + (DD_CAR + DT_HH) * 4);
  // Persistent CTA: work items blockIdx.x, +gridDim.x, ...  Every warp streams its own 32-particle slices (slice w, w+16, ...
  // of each item) through a private DD_NS-deep TMA ring with private mbarriers.  The ring runs ahead ACROSS items, so the
  // loads of the next tile are in flight while this one is flushed; block-wide barriers only bracket the flush.
  constexpr int NW = DT_THREADS / 32;
  const int      lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const uint32_t tile_s = smem_u32(tile);
  const uint32_t bar = smem_u32(mbar) + 8u * (uint32_t)(wrp * DD_NS), dst = smem_u32(sp) + 512u * (uint32_t)(wrp * DD_NS);
  const int      M = L - 1, sh = 32 - logL;
  const uint32_t copy_off = (VAR != 4 && (lane & 1)) ? 4u * DD_CO : 0u;
  // producer state (warp-uniform): next slice to request
  int  p_item = blockIdx.x, p_k = 0, p_s0 = 0, p_np = 0, p_n = 0;
  uint32_t gi = 0, g = 0;                                       // slices requested / consumed by this warp so far
  if (p_item < W) { const int4 w4 = work[p_item]; p_s0 = w4.x; p_np = w4.y; const int nsl = (p_np + 31) >> 5; p_n = nsl > wrp ? (nsl - wrp + NW - 1) / NW : 0; }
  auto produce = [&]() {
    while (p_item < W && p_k >= p_n) {
      p_item += gridDim.x; p_k = 0;
      if (p_item < W) { const int4 w4 = work[p_item]; p_s0 = w4.x; p_np = w4.y; const int nsl = (p_np + 31) >> 5; p_n = nsl > wrp ? (nsl - wrp + NW - 1) / NW : 0; }
    }
    if (p_item >= W) return;
    if (lane == 0) {
      const int      first = (wrp + p_k * NW) << 5, cnt = min(32, p_np - first);
      const uint32_t bytes = (uint32_t)cnt * 16u, bb = bar + 8u * (gi % DD_NS), dd = dst + 512u * (gi % DD_NS);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dd), "l"(pos4 + p_s0 + first), "r"(bytes), "r"(bb) : "memory");
    }
    gi++; p_k++;
  };
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < DD_NS; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8u * q));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < DD_NS - 1; q++) produce();
  for (int i = threadIdx.x; i < (DD_CAR + DT_HH) / 4; i += DT_THREADS) reinterpret_cast<uint4 *>(tile)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int item = blockIdx.x; item < W; item += gridDim.x) {
  const int4 wk = work[item];
  const int  np = wk.y;
  const bool sole = wk.w != 0;
  const int  x0 = (wk.z & 1023) * DT_T, y0 = ((wk.z >> 10) & 1023) * DT_T, z0 = ((wk.z >> 20) & 1023) * DT_T;
  const int  nsl = (np + 31) >> 5;
  const int  nmine = nsl > wrp ? (nsl - wrp + NW - 1) / NW : 0;
  for (int k = 0; k < nmine; k++) {
    produce();                                  // refills the slot read in the previous iteration (ordered by the __syncwarp below)
    {
      const uint32_t bb = bar + 8u * (g % DD_NS), parity = (g / DD_NS) & 1u;
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bb), "r"(parity) : "memory");
      }
    }
    const int  i = ((wrp + k * NW) << 5) + lane;
    const bool valid = i < np;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(dst + 512u * (g % DD_NS) + 16u * lane) : "memory");
    __syncwarp();
    g++;
    const uint32_t ux = pos_q32(q.x), uy = pos_q32(q.y), uz = pos_q32(q.z);
    const int cx = (int)(ux >> sh), cy = (int)(uy >> sh), cz = (int)(uz >> sh);
    const int lx = cx - x0, ly = cy - y0, lz = cz - z0;
    const bool intile = (unsigned)(lx | ly | lz) < (unsigned)DT_T;
    const int  widx = (lz * DT_H + ly) * DT_H + lx;
    const int  cid = valid ? (intile ? widx : -2) : -1;
    uint32_t wx[3], wy[3], wz[3], wyz[9];
    tsc_q32(ux << logL, wx); tsc_q32(uy << logL, wy); tsc_q32(uz << logL, wz);
    dom_wz_scale(wz, wsh);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      wyz[k * 3 + 0] = __umulhi(wy[0], wz[k]); wyz[k * 3 + 2] = __umulhi(wy[2], wz[k]);
      wyz[k * 3 + 1] = wz[k] - wyz[k * 3 + 0] - wyz[k * 3 + 2];
    }
    // dense slices (clump cores): the particles of one cell are adjacent lanes (Hilbert order).  When the 32 lanes hold at most
    // `rmax` distinct cells, every run of equal cells sums its 27 terms with REDUX over the run's lane mask (two 16-bit limbs)
    // and only the run's first lane touches shared memory: no same-address serialisation of the atomics.
    const int      cprev = __shfl_up_sync(0xffffffffu, cid, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || cid != cprev);
    const bool grouped = __popc(heads) <= rmax && __all_sync(0xffffffffu, cid >= 0);
    if (grouped) {
      const unsigned upto = (2u << lane) - 1u;                                    // lanes 0..lane (lane 31: all)
      const int      rlo = 31 - __clz(heads & upto);
      const unsigned above = heads & ~upto;
      const unsigned rmask = (above ? ((1u << (__ffs(above) - 1)) - 1u) : 0xffffffffu) & ~((1u << rlo) - 1u);
      const uint32_t lo0 = tile_s + 4u * (uint32_t)widx, ca0 = lo0 + 4u * DD_CAR;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        uint32_t tl[9], th[9], old[9];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const uint32_t w = wyz[k * 3 + j];
          const uint32_t v0 = __umulhi(wx[0], w), v2 = __umulhi(wx[2], w), v1 = w - v0 - v2;
          const uint32_t vv[3] = { v0, v1, v2 };
#pragma unroll
          for (int a = 0; a < 3; a++) {
            const uint32_t l0 = __reduce_add_sync(rmask, vv[a] & 0xffffu), l1 = __reduce_add_sync(rmask, vv[a] >> 16);
            const unsigned long long tot = (unsigned long long)l0 + ((unsigned long long)l1 << 16);
            tl[j * 3 + a] = (uint32_t)tot; th[j * 3 + a] = (uint32_t)(tot >> 32);
          }
        }
        if (lane == rlo) {
#pragma unroll
          for (int q = 0; q < 9; q++) old[q] = dom_atom(lo0 + 4u * (uint32_t)((k * DT_H + q / 3) * DT_H + q % 3), tl[q]);
#pragma unroll
          for (int q = 0; q < 9; q++) {
            const uint32_t hi = th[q] + ((old[q] + tl[q] < old[q]) ? 1u : 0u);
            if (hi) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(ca0 + 4u * (uint32_t)((k * DT_H + q / 3) * DT_H + q % 3)), "r"(hi) : "memory");
          }
        }
      }
    } else if (intile && valid) {
      const uint32_t lo0 = tile_s + 4u * (uint32_t)widx + copy_off, ca0 = tile_s + 4u * (uint32_t)(widx + DD_CAR);
      uint32_t cmask = 0;                                   // bit 26-i: term i wrapped its low word
#pragma unroll
      for (int k = 0; k < 3; k++) {
        uint32_t v[9], old[9];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const uint32_t w = wyz[k * 3 + j];
          v[j * 3 + 0] = __umulhi(wx[0], w); v[j * 3 + 2] = __umulhi(wx[2], w); v[j * 3 + 1] = w - v[j * 3 + 0] - v[j * 3 + 2];
        }
        if (VAR == 1) {
#pragma unroll
          for (int q = 0; q < 9; q++) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(lo0 + 4u * (uint32_t)((k * DT_H + q / 3) * DT_H + q % 3)), "r"(v[q]) : "memory");
          continue;
        }
        if (VAR == 2) {
#pragma unroll
          for (int q = 0; q < 9; q++) cmask ^= v[q];
          continue;
        }
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
          for (int a = 0; a < 3; a++) old[j * 3 + a] = dom_atom(lo0 + 4u * (uint32_t)((k * DT_H + j) * DT_H + a), v[j * 3 + a]);
#pragma unroll
        for (int q = 0; q < 9; q++) dom_carry(cmask, old[q], v[q]);
      }
      if (VAR == 2) { if (cmask == 0x12345u) tile[DD_CAR + widx] = 1; cmask = 0; }
      while (cmask) {
        const int b = 31 - __clz(cmask);
        cmask ^= 1u << b;
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(ca0 + (uint32_t)c_dom_off[26 - b]), "r"(one) : "memory");
      }
    } else if (valid) {
      // ll()'s coordinate clamp put the particle into a cell outside this tile: straight to the global accumulators
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const uint32_t v0 = __umulhi(wx[0], wyz[r]), v2 = __umulhi(wx[2], wyz[r]), v1 = wyz[r] - v0 - v2;
        const uint32_t vv[3] = { v0, v1, v2 };
        const int y = (cy + (r % 3) - 1) & M, z = (cz + (r / 3) - 1) & M;
#pragma unroll
        for (int a = 0; a < 3; a++) atomicAdd(&acc[(((size_t)z << logL | y) << logL) | (size_t)((cx + a - 1) & M)], (unsigned long long)vv[a]);
      }
    }
  }
  __syncthreads();
  // flush + re-zero: thread = one (hx,hy) column of the tile, marching in z (no div/mod, constant strides)
  if (threadIdx.x < DT_H * DT_H) {
    const int hx = threadIdx.x % DT_H, hy = threadIdx.x / DT_H;
    const int x = (x0 + hx - 1) & M, y = (y0 + hy - 1) & M;
    const bool inxy = hx >= 2 && hx <= DT_T - 1 && hy >= 2 && hy <= DT_T - 1;
#pragma unroll 2
    for (int hz = 0; hz < DT_H; hz++) {
      const int i = (hz * DT_H + hy) * DT_H + hx;
      const unsigned long long val = ((unsigned long long)tile[DD_CAR + i] << 32) + (unsigned long long)tile[i] + (unsigned long long)tile[DD_CO + i];
      if (val == 0) continue;
      tile[i] = 0; tile[DD_CO + i] = 0; tile[DD_CAR + i] = 0;
      if (VAR == 3) continue;
      const int z = (z0 + hz - 1) & M;
      unsigned long long *dstp = &acc[(((size_t)z << logL | y) << logL) | x];
      if (sole && inxy && hz >= 2 && hz <= DT_T - 1) *dstp = val; else atomicAdd(dstp, val);
    }
  }
  __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// D4 domain level, round 2b: k_deposit_dom2 -- EXPERIMENT, opt-in (AHFGPU_DOM_V2=1), bit-identical to k_deposit_dom, measured SLOWER.
//   k_deposit_dom is bound by the shared-memory atomic pipe: 27 ATOMS per particle at ~3 wavefronts each (the cells of 32 Hilbert-adjacent
//   particles fall on effectively random banks), two cycles per wavefront = 86 % l1tex busy at 0.35 ms (profiles/r2n_*).  No tile layout removes
//   that (scripts/bank_sim.py); what removes it is giving the LANES a structure:
//     pass 0  the particles of a tile are one contiguous range in which the particles of one CELL are adjacent (the key's leading bits are the
//             cell): heads of the cell runs are found by comparing neighbours, first[cell] = index of the run's first particle; the others
//             go to lists (ranks 1 .. 7 of a run: list A; the body of a longer run: list B);
//     march   thread = one (x, y) column of the tile, a warp = 16 x-consecutive columns of rows y and y + 8 (8 x 18 words = 16 banks apart),
//             walking z = 0 .. 15.  The column's 27 running sums live in REGISTERS as a window of three z planes: the first particle of cell
//             (x, y, z) adds its 27 terms, then plane z is complete for this column and leaves with 9 ATOMS whose 32 addresses are 2 x 16
//             consecutive words: one wavefront each.  9 conflict-free atomics per CELL instead of 27 conflicting ones per particle;
//     rest    list A scatters lane per particle.
//   Sums are single 32-bit words in 2^-28 units (16 particles' worth).  Every particle deposits exactly 2^28 units, so a wrapped word shows as
//   a tile total that is short by a multiple of 2^32: the CTA checks sum(tile) == np * 2^28 and, if not (clump tiles; normally caught in pass 0
//   by a run longer than 8), redoes the tile in the `heavy` form: every sum leaves as two 16-bit limbs into two words per cell (no wrap for
//   any count a chunk can hold), and list B is walked: a thread takes 8 consecutive entries, keeps the 27 sums of the current cell in
//   registers, flushes on a cell change, with REDUX when the flushing lanes hold one cell.  Chunks of tiles above 8192 particles go there
//   directly.  All terms are the integers of k_deposit_dom: the sums are identical bit for bit whatever path served a tile
//   (tests/test_gpu_large.py::test_domain_deposit_kernels_agree).
//   MEASURED (profiles/r2w_dom2_ncu_summary.txt, 256^3): the march does what it was built for -- 1.00 wavefronts per ATOMS, shared-memory
//   wavefronts of the whole kernel 56 M -> 22 M, l1tex busy 86 % -> 48 % -- but the kernel needs MORE instructions than k_deposit_dom
//   (226 M against 185 M warp instructions: pass 0 57 M, march + list A 68 M, heavy tiles 58 M, flush 32 M, against a lattice that fills only
//   ~55 % of the march's lanes) and is issue / latency bound at 24 warps per SM: 0.39 ms against 0.35 ms on the bench box, 0.28 against 0.26
//   on a pure lattice.  A TSC deposit with exact integer sums costs >= ~100 instructions per particle (27 products, 27 sums, 36 for the
//   weights) before any bookkeeping: 52 M warp instructions at 256^3, 0.08 ms at the 60 % issue rate these kernels reach -- the shared-atomic
//   pipe is not the only wall between this stage and half the HBM roofline.
// ------------------------------------------------------------------------------------------------
constexpr int D2_THREADS = 256;
constexpr int D2_NW      = D2_THREADS / 32;
constexpr int D2_FP      = DT_T + 2;                   // row pitch of first[] (u16): rows y and y + 8 fall on different banks
constexpr int D2_FIRST_N = DT_T * DT_T * D2_FP;
constexpr int D2_EMPTY   = 0xFFFF;
constexpr int D2_WALK    = 8;
constexpr int D2_RUN     = 8;                          // particles of a cell run beyond this rank go to the walk list
constexpr int D2_LEFT_N  = DT_CHUNK + 32 * D2_NW;      // every warp's segment starts at a multiple of 32
constexpr int D2_SMEM    = 2 * DT_HH * 4 + D2_FIRST_N * 2 + D2_LEFT_N * 2;              // LO | HI | first[] | left[]
static_assert(DT_CHUNK <= 65534, "u16 particle indices");
static_assert((2 * DT_HH * 4) % 16 == 0 && (D2_FIRST_N * 2) % 16 == 0, "left[] is read with 16-byte loads");

__device__ __forceinline__ void red_s32(uint32_t addr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
// in-tile particle: no x == 1 clamp to think of (pass 0 sends such a tile to the heavy form, whose walk uses pos_q32)
__device__ __forceinline__ uint32_t pos_q32_in(float x) { return __float2uint_rz(x * 4294967296.0f); }
template <bool INTILE>
__device__ __forceinline__ void dom2_weights(const float4 q, const int logL, const bool act, uint32_t (&wx)[3], uint32_t (&wyz)[9])
{
  const uint32_t ux = INTILE ? pos_q32_in(q.x) : pos_q32(q.x), uy = INTILE ? pos_q32_in(q.y) : pos_q32(q.y), uz = INTILE ? pos_q32_in(q.z) : pos_q32(q.z);
  uint32_t wy[3], wz[3];
  tsc_q32(ux << logL, wx); tsc_q32(uy << logL, wy); tsc_q32(uz << logL, wz);
  dom_wz_scale(wz, 32 - D2_S);
#pragma unroll
  for (int k = 0; k < 3; k++) wz[k] = act ? wz[k] : 0u;                    // an empty cell adds zeros: no divergence in the 27-term block
#pragma unroll
  for (int k = 0; k < 3; k++) {
    wyz[k * 3 + 0] = __umulhi(wy[0], wz[k]); wyz[k * 3 + 2] = __umulhi(wy[2], wz[k]);
    wyz[k * 3 + 1] = wz[k] - wyz[k * 3 + 0] - wyz[k * 3 + 2];
  }
}
// one sum into the tile: a single word (light form) or two 16-bit limbs into LO and HI = LO + DT_HH words (heavy form)
template <bool SPLIT> __device__ __forceinline__ void dom2_put(uint32_t lo_addr, uint32_t v)
{
  if (SPLIT) { red_s32(lo_addr, v & 0xffffu); red_s32(lo_addr + 4u * DT_HH, v >> 16); }
  else red_s32(lo_addr, v);
}
// local cell of a particle inside the tile at (x0, y0, z0): (lz * 16 + ly) * 16 + lx, or -2 (outside: only x == 1.0f, which ll() clamps to
// cell 0, can do that).  trunc(x * L) == trunc(x * 2^32) >> (32 - logL) for 0 <= x < 1 (L is a power of two: the product is exact)
__device__ __forceinline__ int dom2_cell(const float4 q, const float Lf, const int x0, const int y0, const int z0)
{
  const int lx = __float2int_rz(q.x * Lf) - x0, ly = __float2int_rz(q.y * Lf) - y0, lz = __float2int_rz(q.z * Lf) - z0;
  return ((unsigned)(lx | ly | lz) < (unsigned)DT_T) ? (lz * DT_T + ly) * DT_T + lx : -2;
}

// march: column (x, y), source cells z = 0 .. 15; a[(z + c) % 3][b * 3 + t] = running sum of target plane z + c (tile coordinates with the rim).
// Steps z = 16, 17 have no source cell and only let the last two planes leave.  Unrolled by three (the window's period), not by 18: the
// fully unrolled kernel was 300 KB of code and a third of its stalls were instruction fetches.
template <bool SPLIT>
__device__ __forceinline__ void dom2_march(const float4 *__restrict__ P, const uint32_t first_s, const uint32_t lo_s, const int lane, const int wrp, const int logL)
{
  const int x = lane & 15, y = wrp + 8 * (lane >> 4);
  uint32_t pb = lo_s + 4u * (uint32_t)(y * DT_H + x);                     // target (a, b, c) = (0, 0, 0) of the current source cell
  uint32_t fa = first_s + 2u * (uint32_t)(y * D2_FP + x);
  uint32_t a[3][9];
#pragma unroll
  for (int t = 0; t < 9; t++) { a[0][t] = 0u; a[1][t] = 0u; }
  uint32_t idn = lds_u16(fa);
  float4   qn = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idn != (uint32_t)D2_EMPTY) qn = __ldg(P + idn);
#pragma unroll 1
  for (int zz = 0; zz < DT_T + 2; zz += 3) {
#pragma unroll
    for (int u = 0; u < 3; u++) {
      const int    z = zz + u;
      const bool   act = idn != (uint32_t)D2_EMPTY;
      const float4 q = qn;
      idn = (uint32_t)D2_EMPTY;
      if (z + 1 < DT_T) {
        fa += 2u * (uint32_t)(DT_T * D2_FP);
        idn = lds_u16(fa);
        if (idn != (uint32_t)D2_EMPTY) qn = __ldg(P + idn);
      }
      if (z < DT_T) {                                                     // warp uniform
        uint32_t wx[3], wyz[9];
        dom2_weights<true>(q, logL, act, wx, wyz);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const int slot = (u + c) % 3;
#pragma unroll
          for (int b = 0; b < 3; b++) {
            const uint32_t w = wyz[c * 3 + b];
            const uint32_t v0 = __umulhi(wx[0], w), v2 = __umulhi(wx[2], w), v1 = w - v0 - v2;
            if (c == 2) { a[slot][b * 3 + 0] = v0; a[slot][b * 3 + 1] = v1; a[slot][b * 3 + 2] = v2; }      // a plane enters the window
            else { a[slot][b * 3 + 0] += v0; a[slot][b * 3 + 1] += v1; a[slot][b * 3 + 2] += v2; }
          }
        }
      }
#pragma unroll
      for (int b = 0; b < 3; b++)                                         // plane z has all it gets from this column
#pragma unroll
        for (int t = 0; t < 3; t++) dom2_put<SPLIT>(pb + 4u * (uint32_t)(b * DT_H + t), a[u][b * 3 + t]);
      pb += 4u * (uint32_t)(DT_H * DT_H);
    }
  }
}
// lane per particle: the entries [0, n) of a warp's list
template <bool SPLIT>
__device__ __forceinline__ void dom2_scatter(const float4 *__restrict__ P, const uint32_t list_s, const int n, const uint32_t lo_s, const int lane, const int logL,
                                             const int sh, const int x0, const int y0, const int z0)
{
  for (int j = lane; j < n; j += 32) {
    const float4 q = __ldg(P + lds_u16(list_s + 2u * (uint32_t)j));
    const int lx = (int)(pos_q32_in(q.x) >> sh) - x0, ly = (int)(pos_q32_in(q.y) >> sh) - y0, lz = (int)(pos_q32_in(q.z) >> sh) - z0;
    const uint32_t t0 = lo_s + 4u * (uint32_t)((lz * DT_H + ly) * DT_H + lx);
    uint32_t wx[3], wyz[9];
    dom2_weights<true>(q, logL, true, wx, wyz);
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int b = 0; b < 3; b++) {
        const uint32_t w = wyz[c * 3 + b];
        const uint32_t v0 = __umulhi(wx[0], w), v2 = __umulhi(wx[2], w), v1 = w - v0 - v2;
        const uint32_t tb = t0 + 4u * (uint32_t)((c * DT_H + b) * DT_H);
        dom2_put<SPLIT>(tb, v0); dom2_put<SPLIT>(tb + 4u, v1); dom2_put<SPLIT>(tb + 8u, v2);
      }
  }
}

__global__ void __launch_bounds__(D2_THREADS, 3)
k_deposit_dom2(const float4 *__restrict__ pos4, const int4 *__restrict__ work, const int *__restrict__ Wp, int L, int logL,
               unsigned long long *__restrict__ acc, const int force_heavy, unsigned int *__restrict__ stats)
{
  const int W = *Wp;                                   // the grid is an upper bound of the work list (no host read-back of its length)
  if ((int)blockIdx.x >= W) return;
  extern __shared__ __align__(16) unsigned char dsm[];
  uint32_t *LO = reinterpret_cast<uint32_t *>(dsm);
  __shared__ int s_heavy;
  __shared__ unsigned long long s_red[D2_NW];
  const int4 wk = work[blockIdx.x];
  const int  np = wk.y;
  const bool sole = wk.w != 0;
  const int  x0 = (wk.z & 1023) * DT_T, y0 = ((wk.z >> 10) & 1023) * DT_T, z0 = ((wk.z >> 20) & 1023) * DT_T;
  const int  M = L - 1, sh = 32 - logL;
  const float Lf = (float)L;
  const int  lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const uint32_t lo_s = smem_u32(LO), first_s = lo_s + 8u * DT_HH, left_s = first_s + 2u * D2_FIRST_N;
  const float4 *__restrict__ P = pos4 + wk.x;
  for (int i = threadIdx.x; i < DT_HH; i += D2_THREADS) LO[i] = 0u;
  for (int i = threadIdx.x; i < D2_FIRST_N / 2; i += D2_THREADS) LO[2 * DT_HH + i] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) s_heavy = (!sole || force_heavy) ? 1 : 0;
  __syncthreads();
  // ---- pass 0: warp w owns the slices [sl0, sl1) of 32 particles.  Heads of the cell runs -> first[]; every other particle goes to the
  //      warp's own segment of left[]: ranks 1 .. D2_RUN - 1 of a run from the segment's start upwards (list A, nA entries), the body of a
  //      long run (clump cores) from its end downwards (list B, nB entries; in particle order reversed: the runs of one cell stay together)
  const int ns = (np + 31) >> 5, sl0 = (wrp * ns) / D2_NW, sl1 = ((wrp + 1) * ns) / D2_NW;
  const uint32_t segA = left_s + 2u * (uint32_t)(sl0 * 32 + 32 * wrp);
  const int      cap = (sl1 - sl0) * 32;
  int nA = 0, nB = 0;
  {
    int carry = -3, carry_len = 0;                     // cell of the particle before the slice; length of the run it ends
    if (sl0 > 0 && sl0 < sl1) carry = dom2_cell(__ldg(P + sl0 * 32 - 1), Lf, x0, y0, z0);
    bool flag = false;
    const unsigned lt = (1u << lane) - 1u;
    for (int sl = sl0; sl < sl1; sl += 4) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int i = (sl + u) * 32 + lane; if (sl + u < sl1 && i < np) q[u] = __ldg(P + i); }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (sl + u >= sl1) break;                       // warp uniform
        const int  i = (sl + u) * 32 + lane;
        const bool valid = i < np;
        const int  cid = valid ? dom2_cell(q[u], Lf, x0, y0, z0) : -1;
        int prev = __shfl_up_sync(0xffffffffu, cid, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(0xffffffffu, cid, 31);
        const bool     brk = cid != prev;                                  // a new run starts here (a clamped particle starts one too)
        const bool     head = valid && brk && cid >= 0;
        const unsigned bb = __ballot_sync(0xffffffffu, brk), vb = __ballot_sync(0xffffffffu, valid);
        const unsigned hb = __ballot_sync(0xffffffffu, head);
        flag = flag || (__ballot_sync(0xffffffffu, valid && cid < 0) != 0u);                   // a clamped particle: heavy form (its walk knows ll()'s clamp)
        // rank of the particle inside its run
        const unsigned below = bb & (lt | (1u << lane));
        const int rank = below ? lane - (31 - __clz(below)) : carry_len + lane + 1;
        carry_len = (bb ? 31 - (31 - __clz(bb)) : carry_len + 32);       // run length up to and including lane 31
        const bool toB = valid && !head && (rank >= D2_RUN || cid < 0);
        const bool toA = valid && !head && !toB;
        const unsigned ab = __ballot_sync(0xffffffffu, toA), Bb = __ballot_sync(0xffffffffu, toB);
        flag = flag || Bb != 0u;                                           // a cell with more than D2_RUN particles: single words will not do
        if (head) sts_u16(first_s + 2u * (uint32_t)((cid >> 4) * D2_FP + (cid & 15)), (uint32_t)i);
        if (toA) sts_u16(segA + 2u * (uint32_t)(nA + __popc(ab & lt)), (uint32_t)i);
        if (toB) sts_u16(segA + 2u * (uint32_t)(cap - 1 - nB - __popc(Bb & lt)), (uint32_t)i);
        nA += __popc(ab); nB += __popc(Bb);
        (void)hb; (void)vb;
      }
    }
    if (flag && lane == 0) s_heavy = 1;
  }
  __syncthreads();
  bool heavy = s_heavy != 0;
  for (;;) {
    if (!heavy) {
      dom2_march<false>(P, first_s, lo_s, lane, wrp, logL);
      dom2_scatter<false>(P, segA, nA, lo_s, lane, logL, sh, x0, y0, z0);
      __syncthreads();
      // every particle deposited exactly 2^DD_S units: a wrapped word (or a particle pass 0 lost) shows in the total
      unsigned long long s = 0;
      for (int i = threadIdx.x; i < DT_HH; i += D2_THREADS) s += LO[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_red[wrp] = s;
      __syncthreads();
      unsigned long long tot = 0;
#pragma unroll
      for (int w = 0; w < D2_NW; w++) tot += s_red[w];
      if (tot == ((unsigned long long)np << D2_S)) break;
      heavy = true;
      if (stats && threadIdx.x == 0) atomicAdd(stats + 1, 1u);
      __syncthreads();                                  // everybody has read LO
    }
    // ---- heavy form: sums leave as two 16-bit limbs (LO and HI get < 2^16 each per flush, at most DT_CHUNK flushes reach a word: no wrap).
    //      First particles by the same march, list A lane per particle; list B (the bodies of long runs) by a register walk: a thread takes
    //      D2_WALK consecutive entries (in a clump core: one cell), keeps the 27 sums of its current cell in registers (D2_WALK * 2^28 < 2^32)
    //      and flushes on a cell change.  When all flushing lanes of the warp hold the same cell the limbs are summed with REDUX and one
    //      lane issues the atomics.
    if (stats && threadIdx.x == 0) atomicAdd(stats, 1u);
    for (int i = threadIdx.x; i < 2 * DT_HH; i += D2_THREADS) LO[i] = 0u;
    __syncthreads();
    dom2_march<true>(P, first_s, lo_s, lane, wrp, logL);
    dom2_scatter<true>(P, segA, nA, lo_s, lane, logL, sh, x0, y0, z0);
    const uint32_t segB = segA + 2u * (uint32_t)(cap - nB);              // cap and nA + nB are multiples of ... not of 8: single loads
    for (int e0 = 0; e0 < nB; e0 += 32 * D2_WALK) {
      const int eb = e0 + lane * D2_WALK;
      uint32_t a[27];
      int      cur = -1;
      float4   qn = make_float4(0.f, 0.f, 0.f, 0.f);
      if (eb < nB) qn = __ldg(P + lds_u16(segB + 2u * (uint32_t)eb));
#pragma unroll 1
      for (int k = 0; k <= D2_WALK; k++) {
        const bool   have = k < D2_WALK && eb + k < nB;
        const float4 q = qn;
        if (k + 1 < D2_WALK && eb + k + 1 < nB) qn = __ldg(P + lds_u16(segB + 2u * (uint32_t)(eb + k + 1)));
        int widx = -1, cx = 0, cy = 0, cz = 0;
        uint32_t wx[3], wyz[9];
        if (have) {
          cx = (int)(pos_q32(q.x) >> sh); cy = (int)(pos_q32(q.y) >> sh); cz = (int)(pos_q32(q.z) >> sh);
          const int lx = cx - x0, ly = cy - y0, lz = cz - z0;
          if ((unsigned)(lx | ly | lz) < (unsigned)DT_T) widx = (lz * DT_H + ly) * DT_H + lx;
          dom2_weights<false>(q, logL, true, wx, wyz);
        }
        const bool     need = widx != cur && cur >= 0;
        const unsigned fm = __ballot_sync(0xffffffffu, need);
        if (fm) {
          const int  c0 = __shfl_sync(0xffffffffu, cur, __ffs(fm) - 1);
          const bool same = __popc(fm) >= 4 && __ballot_sync(0xffffffffu, need && cur == c0) == fm;
          if (same) {
            if (need) {
              const uint32_t t0 = lo_s + 4u * (uint32_t)cur;
              const bool     leader = lane == __ffs(fm) - 1;
#pragma unroll
              for (int t = 0; t < 27; t++) {
                const uint32_t l0 = __reduce_add_sync(fm, a[t] & 0xffffu), l1 = __reduce_add_sync(fm, a[t] >> 16);
                if (leader) {
                  const uint32_t ad = t0 + 4u * (uint32_t)(((t / 9) * DT_H + (t / 3) % 3) * DT_H + t % 3);
                  red_s32(ad, l0); red_s32(ad + 4u * DT_HH, l1);
                }
              }
            }
          } else if (need) {
            const uint32_t t0 = lo_s + 4u * (uint32_t)cur;
#pragma unroll
            for (int t = 0; t < 27; t++) dom2_put<true>(t0 + 4u * (uint32_t)(((t / 9) * DT_H + (t / 3) % 3) * DT_H + t % 3), a[t]);
          }
        }
        if (widx != cur) {
          cur = widx;
#pragma unroll
          for (int t = 0; t < 27; t++) a[t] = 0u;
        }
        if (have) {
#pragma unroll
          for (int r = 0; r < 9; r++) {
            const uint32_t w = wyz[r];
            const uint32_t v0 = __umulhi(wx[0], w), v2 = __umulhi(wx[2], w), v1 = w - v0 - v2;
            if (widx >= 0) { a[r * 3 + 0] += v0; a[r * 3 + 1] += v1; a[r * 3 + 2] += v2; }
            else {
              // ll()'s coordinate clamp put the particle into a cell outside this tile: straight to the global accumulators
              const int y = (cy + (r % 3) - 1) & M, z = (cz + (r / 3) - 1) & M;
              const uint32_t vv[3] = { v0, v1, v2 };
#pragma unroll
              for (int t = 0; t < 3; t++) atomicAdd(&acc[(((size_t)z << logL | y) << logL) | (size_t)((cx + t - 1) & M)], (unsigned long long)vv[t]);
            }
          }
        }
      }
    }
    __syncthreads();
    break;
  }
  // ---- flush: plain 8-byte stores for the cells no other CTA can touch, REDG.ADD.64 for the tile rim and multi-chunk tiles.
  //      Thread = one (hx, hy) column marching in z (no div / mod in the loop); the 68 columns beyond the 256 threads are spread item-wise
  auto put = [&](const uint32_t ad, unsigned long long *col, const int hz, const bool inxy) {
    unsigned long long val = lds_u32(ad);
    if (heavy) val += (unsigned long long)lds_u32(ad + 4u * DT_HH) << 16;
    if (val == 0) return;
    unsigned long long *dstp = col + ((size_t)((z0 + hz - 1) & M) << (2 * logL));
    if (inxy && hz >= 2 && hz <= DT_T - 1) *dstp = val; else atomicAdd(dstp, val);
  };
  {
    const int hx = threadIdx.x % DT_H, hy = threadIdx.x / DT_H;         // columns 0 .. 255
    const bool inxy = sole && hx >= 2 && hx <= DT_T - 1 && hy >= 2 && hy <= DT_T - 1;
    unsigned long long *col = acc + (((size_t)((y0 + hy - 1) & M) << logL) | (size_t)((x0 + hx - 1) & M));
#pragma unroll 1
    for (int hz = 0; hz < DT_H; hz++) put(lo_s + 4u * (uint32_t)(hz * DT_H * DT_H + threadIdx.x), col, hz, inxy);
    constexpr int REST = DT_H * DT_H - D2_THREADS;
    for (int e = threadIdx.x; e < REST * DT_H; e += D2_THREADS) {
      const int hz = e / REST, r = D2_THREADS + (e - hz * REST), hx2 = r % DT_H, hy2 = r / DT_H;
      put(lo_s + 4u * (uint32_t)(hz * DT_H * DT_H + r), acc + (((size_t)((y0 + hy2 - 1) & M) << logL) | (size_t)((x0 + hx2 - 1) & M)), hz,
          sole && hx2 >= 2 && hx2 <= DT_T - 1 && hy2 >= 2 && hy2 <= DT_T - 1);
    }
  }
}

// EXPERIMENT (AHFGPU_DOM_CELLSORT=1): order the particles of every tile by cell (z, y, x) before the domain deposit, to measure what
// bank-conflict-free lanes are worth to the shared-memory atomics (the sums are integers: any order gives the same result)
__global__ void k_cellsort_keys(const float4 *__restrict__ pos4, const uint64_t *__restrict__ keys, uint64_t n, int logL, int tbits,
                                uint64_t *__restrict__ ck, uint32_t *__restrict__ idx)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = pos4[i];
  const int sh = 32 - logL;
  const uint32_t cx = pos_q32(q.x) >> sh, cy = pos_q32(q.y) >> sh, cz = pos_q32(q.z) >> sh;
  const uint64_t tile = keys[i] >> (3 * (21 - tbits));
  ck[i] = (tile << 12) | (uint64_t)(((cz & 15u) << 8) | ((cy & 15u) << 4) | (cx & 15u));
  idx[i] = (uint32_t)i;
}
__global__ void k_gather_f4(const float4 *__restrict__ in, const uint32_t *__restrict__ idx, uint64_t n, float4 *__restrict__ out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

__global__ void k_finish_dens(const unsigned long long *__restrict__ acc, float *__restrict__ dens, int ncell, double m2d_over_scale)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const unsigned long long a = acc[c];
  dens[c] = a ? (float)(m2d_over_scale * (double)a - 1.0) : -1.0f;      // zero_dens: -mean_dens, density.c:480 (an empty box has an infinite scale)
}

// ------------------------------------------------------------------------------------------------
// F1: test_node (refine_grid.c:113-138)
// ------------------------------------------------------------------------------------------------
__global__ void k_test_node(LV v, const float *__restrict__ dens, const uint8_t *__restrict__ interior,
                            const int32_t *__restrict__ nbr, double thr, uint8_t *__restrict__ tn)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  bool hit = false;
  if (v.dense) {
    int x, y, z; lv_coords(v, c, x, y, z);
    const int M = (int)(v.L - 1);
#pragma unroll
    for (int k = -1; k <= 1; k++)
#pragma unroll
      for (int j = -1; j <= 1; j++)
#pragma unroll
        for (int a = 0; a <= 1; a++) {
          int t = (int)lv_key(v, (x + a) & M, (y + j) & M, (z + k) & M);
          hit |= ((double)dens[t] >= thr);
        }
  } else if (interior[c]) {
    // term-major table: the loads of a warp are coalesced.  An interior cell sees all 27 neighbours, so x+1 of row q is the array
    // neighbour of the row's centre (the row's first cell across the periodic face)
    const bool face = (int)(v.ckey[c] & (uint64_t)(v.L - 1)) == (int)v.L - 1;
    int t18[18];
#pragma unroll
    for (int q = 0; q < 9; q++) {
      const int rm = nbr[(size_t)q * (size_t)v.ncell + (size_t)c];
      t18[2 * q] = rm;
      t18[2 * q + 1] = face ? v.row_c0[v.crow[rm]] : rm + 1;
    }
#pragma unroll
    for (int q = 0; q < 18; q++) hit |= ((double)dens[t18[q]] >= thr);
  }
  tn[c] = hit ? 1 : 0;
}

// F1 on the dense domain grid (L a multiple of 32): a CTA stages the comparison bits of a 33 x 10 x 10 block in shared memory and
// evaluates the 18-cell stencil of its 32 x 8 x 8 cells from there (the per-cell form re-reads every density 18 times through L1)
__global__ void __launch_bounds__(256) k_test_node_dense(LV v, const float *__restrict__ dens, double thr, uint8_t *__restrict__ tn)
{
  __shared__ uint8_t b[10][10][36];                       // [z][y][x], 33 x used
  const int M = (int)(v.L - 1), logL = v.logL;
  const int tx = blockIdx.x * 32, ty = blockIdx.y * 8, tz = blockIdx.z * 8;
  for (int i = threadIdx.x; i < 10 * 10 * 33; i += 256) {
    const int x = i % 33, r = i / 33, y = r % 10, z = r / 10;
    const int gx = (tx + x) & M, gy = (ty + y - 1) & M, gz = (tz + z - 1) & M;
    b[z][y][x] = ((double)dens[((((size_t)gz << logL) | (size_t)gy) << logL) | (size_t)gx] >= thr) ? 1 : 0;
  }
  __syncthreads();
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
#pragma unroll
  for (int z = 0; z < 8; z++) {
    unsigned hit = 0;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int j = 0; j < 3; j++) hit |= (unsigned)b[z + k][y + j][x] | (unsigned)b[z + k][y + j][x + 1];
    tn[((((size_t)(tz + z) << logL) | (size_t)(ty + y)) << logL) | (size_t)(tx + x)] = hit ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// F2: which cells spawn children (ref_pquad / ref_cquad / ref_nquad)
// ------------------------------------------------------------------------------------------------
__global__ void k_mark_dense(LV v, const uint8_t *__restrict__ tn, uint8_t *__restrict__ mark)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  int x = c & (int)(v.L - 1);
  uint8_t m = 0;
  if (tn[c]) m = 1;
  else if (x != v.L - 1) {                         // the run's last node never gets a ghost pair
    int prev = (x == 0) ? c + (int)(v.L - 1) : c - 1;
    if (tn[prev]) m = 2;
  }
  mark[c] = m;
}

__device__ __forceinline__ bool tested_idx(long long t, int low, int up, long long len, bool case2)
{
  long long last = (len - 1) + up;
  if (t >= low && t < last) return true;
  long long e = last > low ? last : low;
  return case2 && t == e;
}
__device__ __forceinline__ void run_offsets(bool wrapped, long long r0, long long r1, long long L, int &low, int &up)
{
  low = 1; up = -1;
  if (wrapped) {
    if (r0 == 0 && r1 == L) { low = 0; up = 0; }
    else if (r0 == 0) { low = 0; up = -1; }
    else if (r1 == L) { low = 1; up = 0; }
  }
}

// y-run (cquad) bounds of every row as row indices [q0,q1): one thread per row walks to both ends of its run inside its plane
// (runs are a few tens of rows; one thread per PLANE walking all its rows left most of the GPU idle)
// y-runs of the rows inside their plane: the keys of a plane's rows increase strictly, so rowkey[r] - r is constant exactly along a run of
// consecutive y: both ends by binary search inside the plane
__global__ void k_row_runs(const uint64_t *__restrict__ rowkey, const int32_t *__restrict__ plane_r0, const int32_t *__restrict__ rowplane, int nrow,
                           int32_t *__restrict__ rq0, int32_t *__restrict__ rq1)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  const int P = rowplane[r], r0 = plane_r0[P], r1 = plane_r0[P + 1];
  const long long d = (long long)rowkey[r] - r;
  int lo = r0, hi = r;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)rowkey[mid] - mid < d) lo = mid + 1; else hi = mid; }
  rq0[r] = lo;
  lo = r + 1; hi = r1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)rowkey[mid] - mid <= d) lo = mid + 1; else hi = mid; }
  rq1[r] = lo;
}

// z-run (pquad) bounds of every plane as plane indices [p0,p1): one thread per plane walks to the ends of its run
__global__ void k_plane_z(const uint64_t *__restrict__ rowkey, const int32_t *__restrict__ plane_r0, int nplane, int logL, int32_t *__restrict__ pz)
{
  int P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P < nplane) pz[P] = (int32_t)(rowkey[plane_r0[P]] >> logL);
}
// z-runs of the planes: pz is strictly increasing, so pz[P] - P is constant exactly along a run of consecutive z and never decreases:
// both ends of the run by binary search (the walk along the run took 20 us per level on runs of hundreds of planes)
__global__ void k_plane_runs(const int32_t *__restrict__ pz, int nplane, int32_t *__restrict__ pp0, int32_t *__restrict__ pp1)
{
  int P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= nplane) return;
  const int d = pz[P] - P;
  int lo = 0, hi = P;                                   // first index with pz[i] - i == d
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (pz[mid] - mid < d) lo = mid + 1; else hi = mid; }
  pp0[P] = lo;
  lo = P + 1; hi = nplane;                              // first index behind P with pz[i] - i > d
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (pz[mid] - mid <= d) lo = mid + 1; else hi = mid; }
  pp1[P] = lo;
}

__global__ void k_row_tested(const uint64_t *__restrict__ rowkey, const int32_t *__restrict__ plane_r0, const int32_t *__restrict__ rowplane,
                             const int32_t *__restrict__ rq0, const int32_t *__restrict__ rq1, const int32_t *__restrict__ pp0,
                             const int32_t *__restrict__ pp1, int nrow, int nplane, long long L, int logL, uint8_t *__restrict__ row_tested,
                             uint8_t *__restrict__ row_flags)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  const uint64_t M = (uint64_t)(L - 1);
  int P = rowplane[r];
  // z direction (refine_grid.c:645-703, :818)
  int  a = pp0[P], b = pp1[P];
  long long z0 = (long long)(rowkey[plane_r0[a]] >> logL), z1 = (long long)(rowkey[plane_r0[b - 1]] >> logL) + 1;
  bool zwrapped = ((rowkey[plane_r0[0]] >> logL) == 0) && ((long long)(rowkey[plane_r0[nplane - 1]] >> logL) == L - 1);
  int  low, up;
  run_offsets(zwrapped, z0, z1, L, low, up);
  bool tz = tested_idx(P - a, low, up, b - a, (b == nplane) && (z1 == L));
  // y direction (refine_grid.c:346-405, :496); the wrap test looks at the FIRST plane of the z-run (:350)
  int  fr0 = plane_r0[a], fr1 = plane_r0[a + 1];
  bool ywrapped = ((rowkey[fr0] & M) == 0) && ((long long)(rowkey[fr1 - 1] & M) == L - 1);
  int  q0 = rq0[r], q1 = rq1[r];
  long long y0 = (long long)(rowkey[q0] & M), y1 = (long long)(rowkey[q1 - 1] & M) + 1;
  run_offsets(ywrapped, y0, y1, L, low, up);
  bool ty = tested_idx(r - q0, low, up, q1 - q0, (q1 == plane_r0[P + 1]) && (y1 == L));
  row_tested[r] = (tz && ty) ? 1 : 0;
  // first / last row of its cquad, first / last plane of its pquad (the run flags the query API exports)
  row_flags[r] = (uint8_t)((r == q0 ? 4 : 0) | (r == q1 - 1 ? 8 : 0) | (P == a ? 16 : 0) | (P == b - 1 ? 32 : 0));
}

__global__ void k_mark_sparse(LV v, const uint8_t *__restrict__ tn, const uint8_t *__restrict__ interior, const int32_t *__restrict__ crow,
                              const int32_t *__restrict__ row_c0, const uint8_t *__restrict__ row_tested, uint8_t *__restrict__ mark)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  int r = crow[c];
  uint8_t m = 0;
  if (row_tested[r]) {
    const int      c0 = row_c0[r], c1 = row_c0[r + 1];
    const uint64_t M = (uint64_t)(v.L - 1);
    const uint64_t k = v.ckey[c];
    const long long x = (long long)(k & M);
    const bool rowwrap = ((long long)(v.ckey[c1 - 1] & M) == v.L - 1);
    const bool first = (c == c0) || (v.ckey[c - 1] + 1 != k) || v.xbreak[c - 1];
    const bool last  = (c == c1 - 1) || (v.ckey[c + 1] != k + 1) || v.xbreak[c];
    const bool inloop = !last && (!first || (x == 0 && rowwrap));
    const bool case2  = last && (x == v.L - 1);
    if ((inloop || case2) && tn[c]) m = 1;
    else if (inloop && !tn[c] && interior[c]) {
      bool state;
      if (first) state = (x == 0 && rowwrap && tn[c1 - 1]);
      else {
        // c-1 belongs to the same run; it was part of the loop unless it is the run's untested first node
        const uint64_t kp = v.ckey[c - 1];
        const bool pfirst = (c - 1 == c0) || (v.ckey[c - 2] + 1 != kp) || v.xbreak[c - 2];
        const bool pin    = !pfirst || (((long long)(kp & M) == 0) && rowwrap);
        state = pin && tn[c - 1];
      }
      if (state) m = 2;
    }
  }
  mark[c] = m;
}

// ------------------------------------------------------------------------------------------------
// next level: 2x2x2 children of every marked cell, written directly in (z,y,x) order
//   idx = 8*MB(P) + k*4*MP(P) + 4*MBR(r) + j*2*MR(r) + 2*RR(c) + i        
// ------------------------------------------------------------------------------------------------
__global__ void k_make_children(LV v, const uint8_t *__restrict__ mark, const int *__restrict__ S, int Mtot,
                                const int32_t *__restrict__ crow, const int32_t *__restrict__ row_c0, const int32_t *__restrict__ rowplane,
                                const int32_t *__restrict__ plane_r0, int nrow, uint64_t *__restrict__ fkey, uint8_t *__restrict__ fbreak,
                                int flogL, int32_t *__restrict__ fparent, int32_t *__restrict__ cidx, int4 *__restrict__ cbase, int32_t *__restrict__ cpar)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  if (!mark[c]) { cidx[c] = -1; return; }
  int r, rc0, rc1, pc0, pc1;
  if (v.dense) {
    r = c >> v.logL; rc0 = r << v.logL; rc1 = rc0 + (int)v.L;
    int P = r >> v.logL; pc0 = P << (2 * v.logL); pc1 = pc0 + (1 << (2 * v.logL));
  } else {
    r = crow[c]; rc0 = row_c0[r]; rc1 = row_c0[r + 1];
    int P = rowplane[r]; pc0 = row_c0[plane_r0[P]]; pc1 = row_c0[plane_r0[P + 1]];
  }
  (void)nrow;
  auto Sat = [&](int i) { return i >= v.ncell ? Mtot : S[i]; };
  const int Sp0 = Sat(pc0), Sp1 = Sat(pc1), Sr0 = Sat(rc0), Sr1 = Sat(rc1);
  const int MB = Sp0, MP = Sp1 - Sp0, MBR = Sr0 - Sp0, MR = Sr1 - Sr0, RR = S[c] - Sr0;
  int x, y, z; lv_coords(v, c, x, y, z);
  const bool ghost = (mark[c] == 2);
  int cb[4];
#pragma unroll
  for (int k = 0; k < 2; k++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      cb[k * 2 + j] = (int)(8ll * MB + (long long)k * 4 * MP + 4ll * MBR + (long long)j * 2 * MR + 2ll * RR);
#pragma unroll
      for (int i = 0; i < 2; i++) {
        long long idx = 8ll * MB + (long long)k * 4 * MP + 4ll * MBR + (long long)j * 2 * MR + 2ll * RR + i;
        fkey[idx]   = ((((uint64_t)(2 * z + k)) << flogL | (uint64_t)(2 * y + j)) << flogL) | (uint64_t)(2 * x + i);
        fbreak[idx] = (ghost && i == 1) ? 1 : 0;                 // the reference's run ends after a ghost pair
        fparent[idx] = c;
      }
    }
  cidx[c] = S[c] | (ghost ? 0x40000000 : 0);
  cbase[S[c]] = make_int4(cb[0], cb[1], cb[2], cb[3]);
  cpar[S[c]] = c;
}

// one insertion per block of 8 x-consecutive cells: the block's cells are consecutive in the sorted cell array, so the thread of
// the block's first existing cell collects the occupancy mask from its (at most 7) successors and claims the slot -- one CAS and
// one 8-byte store per block instead of three atomics per cell
__global__ void k_hash_insert(const uint64_t *__restrict__ ckey, int ncell, uint64_t *__restrict__ hkey, uint2 *__restrict__ hval, uint64_t hmask)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const uint64_t k = ckey[c], kb = k >> 3;
  if (c > 0 && (ckey[c - 1] >> 3) == kb) return;
  uint32_t mask = 1u << (unsigned)(k & 7);
#pragma unroll
  for (int j = 1; j < 8; j++) {
    if (c + j >= ncell) break;
    const uint64_t kj = ckey[c + j];
    if ((kj >> 3) != kb) break;
    mask |= 1u << (unsigned)(kj & 7);
  }
  uint64_t s = mix64(kb) & hmask;
  for (;;) {
    unsigned long long old = atomicCAS((unsigned long long *)&hkey[s], ~0ull, (unsigned long long)kb);
    if (old == ~0ull) { hval[s] = make_uint2((uint32_t)c, mask); return; }
    s = (s + 1) & hmask;
  }
}
__global__ void k_row_heads(const uint64_t *__restrict__ ckey, int ncell, int logL, uint8_t *__restrict__ head)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  head[c] = (c == 0 || (ckey[c] >> logL) != (ckey[c - 1] >> logL)) ? 1 : 0;
}
__global__ void k_row_fill(const uint64_t *__restrict__ ckey, int ncell, int logL, const uint8_t *__restrict__ head, const int *__restrict__ hs,
                           int32_t *__restrict__ crow, uint64_t *__restrict__ rowkey, int32_t *__restrict__ row_c0, int nrow)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int r = hs[c] + head[c] - 1;
  crow[c] = r;
  if (head[c]) { rowkey[r] = ckey[c] >> logL; row_c0[r] = c; }
  if (c == ncell - 1) row_c0[nrow] = ncell;
}
__global__ void k_plane_heads(const uint64_t *__restrict__ rowkey, int nrow, int logL, uint8_t *__restrict__ head)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  head[r] = (r == 0 || (rowkey[r] >> logL) != (rowkey[r - 1] >> logL)) ? 1 : 0;
}
__global__ void k_plane_fill(int nrow, const uint8_t *__restrict__ head, const int *__restrict__ hs, int32_t *__restrict__ rowplane,
                             int32_t *__restrict__ plane_r0, int nplane)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  int P = hs[r] + head[r] - 1;
  rowplane[r] = P;
  if (head[r]) plane_r0[P] = r;
  if (r == nrow - 1) plane_r0[nplane] = nrow;
}

// D5: the 27 neighbours the reference's search sees from each cell + test_tsc (get_nnodes.c:41-51, :459-862)
__global__ void k_neighbours(LV v, int32_t *__restrict__ nbr, uint8_t *__restrict__ interior)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  int x, y, z; lv_coords(v, c, x, y, z);
  const int L = (int)v.L;
  // the 8 rows around the cell's own: all hash probes are issued together (independent loads in flight), resolved afterwards
  int      rowc[9];
  uint64_t kq[9], sq[9], hq[9];
#pragma unroll
  for (int q = 0; q < 9; q++) {
    const int k = q / 3, j = q % 3;
    int zz = z + k - 1; if (zz < 0) zz = L - 1; else if (zz >= L) zz = 0;
    int yy = y + j - 1; if (yy < 0) yy = L - 1; else if (yy >= L) yy = 0;
    kq[q] = lv_key(v, x, yy, zz);
    sq[q] = mix64(kq[q] >> 3) & v.hmask;
  }
#pragma unroll
  for (int q = 0; q < 9; q++) hq[q] = (q == 4) ? 0 : v.hkey[sq[q]];
#pragma unroll
  for (int q = 0; q < 9; q++) {
    if (q == 4) { rowc[q] = c; continue; }
    const uint64_t kb = kq[q] >> 3;
    uint64_t s = sq[q], hk = hq[q];
    while (hk != kb && hk != ~0ull) { s = (s + 1) & v.hmask; hk = v.hkey[s]; }
    sq[q] = s; hq[q] = hk;
  }
#pragma unroll
  for (int q = 0; q < 9; q++) if (q != 4) rowc[q] = (hq[q] == (kq[q] >> 3)) ? slot_cell(v.hval[sq[q]], (unsigned)(kq[q] & 7)) : -1;
  bool all = true;
  uint32_t vis = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int zz = z + k - 1; if (zz < 0) zz = L - 1; else if (zz >= L) zz = 0;
    const int pm = rowc[k * 3 + 1];                               // the plane is found through its (x, y) node (get_nnodes.c:514-651)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      int yy = y + j - 1; if (yy < 0) yy = L - 1; else if (yy >= L) yy = 0;
      const int rm = (pm >= 0) ? rowc[k * 3 + j] : -1;
      nbr[(size_t)(k * 3 + j) * (size_t)v.ncell + (size_t)c] = rm;
      if (rm < 0) { all = false; continue; }
      const uint64_t kk = v.ckey[rm];
      // x-1: same run, else the periodic image when on the face (get_nnodes.c:490-508)
      int xm = -1;
      if (rm > 0 && v.ckey[rm - 1] + 1 == kk && x > 0 && !v.xbreak[rm - 1]) xm = rm - 1;
      else if (x == 0) xm = lv_lookup(v, L - 1, yy, zz);
      // x+1 (get_nnodes.c:465-485)
      int xp = -1;
      if (!v.xbreak[rm] && rm + 1 < v.ncell && v.ckey[rm + 1] == kk + 1 && x < L - 1) xp = rm + 1;
      else if (x == L - 1) xp = lv_lookup(v, 0, yy, zz);
      if (xm >= 0) vis |= 1u << (k * 3 + j);
      if (xp >= 0) vis |= 1u << (9 + k * 3 + j);
      if (xm < 0 || xp < 0) all = false;
    }
  }
  nbr[(size_t)9 * (size_t)v.ncell + (size_t)c] = (int32_t)vis;
  interior[c] = all ? 1 : 0;
}

// Same table as k_neighbours, built WITHOUT hash probes.  A cell (x,y,z) = child (i,j,k) of coarse cell p; its neighbour
// (x+a, y+b, z+c) is child ((i+a)&1, (j+b)&1, (k+c)&1) of the coarse neighbour at offset ((i+a)>>1, (j+b)>>1, (k+c)>>1).  Every
// parent is interior on its own level (it passed test_node, or is the interior cell behind a ghost pair, refine_grid.c:231),
// so all 27 coarse neighbours exist in the coarse table (arithmetic on the dense domain level, periodic wrap included), and the
// children of a coarse cell are found through cidx/cbase.  "x-break" (the run ends after a ghost pair) is a property of the
// parent: the i=1 child of a ghost parent.  The visibility rules are those of k_neighbours (get_nnodes.c:459-862).
__global__ void __launch_bounds__(128) k_neighbours_pc(LV v, const int32_t *__restrict__ parent, LV cv, const int32_t *__restrict__ cnbr,
                                                       const int32_t *__restrict__ cidx, const int4 *__restrict__ cbase,
                                                       int32_t *__restrict__ nbr, uint8_t *__restrict__ interior)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  int x, y, z; lv_coords(v, c, x, y, z);
  const int L = (int)v.L;
  const int p = parent[c];
  const int bi = x & 1, bj = y & 1, bk = z & 1;
  // the 2x2x2 coarse cells the 27 neighbours live in: per dimension the offsets (b-1, b)
  int4 cbq[8];
  bool ghq[8];
  int  px = 0, py = 0, pz = 0;
  const int CM = (int)(cv.L - 1);
  lv_coords(cv, p, px, py, pz);
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int da = (q & 1) + bi - 1, db = ((q >> 1) & 1) + bj - 1, dc = ((q >> 2) & 1) + bk - 1;
    int qc;
    if (cv.dense) qc = (int)lv_key(cv, (px + da) & CM, (py + db) & CM, (pz + dc) & CM);
    else qc = (da == 0 && db == 0 && dc == 0) ? p : nb_get(cv, cnbr, p, (dc + 1) * 3 + (db + 1), da + 1, px);
    const int ci = qc >= 0 ? cidx[qc] : -1;
    ghq[q] = ci >= 0 && (ci & 0x40000000);
    cbq[q] = ci >= 0 ? cbase[ci & 0x3fffffff] : make_int4(-1, -1, -1, -1);
  }
  // geometric neighbour (a,b,cc in -1..1) and whether the run ends after it
  auto geo = [&](int a, int b, int cc, bool &brk) -> int {
    const int tx = bi + a, ty = bj + b, tz = bk + cc;                       // -1..2
    const int q = ((tx >> 1) - (bi - 1)) | (((ty >> 1) - (bj - 1)) << 1) | (((tz >> 1) - (bk - 1)) << 2);
    const int4 cb = cbq[q];
    const int  jk = (tz & 1) * 2 + (ty & 1);
    const int  b0 = jk == 0 ? cb.x : jk == 1 ? cb.y : jk == 2 ? cb.z : cb.w;
    brk = ghq[q] && (tx & 1);
    return b0 < 0 ? -1 : b0 + (tx & 1);
  };
  bool all = true;
  uint32_t vis = 0;
  const size_t os = (size_t)v.ncell;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    bool dummy;
    const int pm = (k == 1) ? c : geo(0, 0, k - 1, dummy);               // the plane is found through its (x, y) node (get_nnodes.c:514-651)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      bool brk_rm = false;
      const int rm = (pm >= 0) ? ((k == 1 && j == 1) ? c : geo(0, j - 1, k - 1, brk_rm)) : -1;
      if (k == 1 && j == 1) brk_rm = ghq[(1 - bi) | ((1 - bj) << 1) | ((1 - bk) << 2)] && bi;      // own parent is slot (1-bi, 1-bj, 1-bk)
      nbr[(size_t)(k * 3 + j) * os + (size_t)c] = rm;
      if (rm < 0) { all = false; continue; }
      bool brk_m = false, brk_p = false;
      const int gm = geo(-1, j - 1, k - 1, brk_m), gp = geo(1, j - 1, k - 1, brk_p);
      // x-1: same run, else the periodic image when on the face (get_nnodes.c:490-508); x+1 (get_nnodes.c:465-485)
      const int xm = (x > 0) ? ((gm >= 0 && !brk_m) ? gm : -1) : gm;
      const int xp = (x < L - 1) ? ((gp >= 0 && !brk_rm) ? gp : -1) : gp;
      if (xm >= 0) vis |= 1u << (k * 3 + j);
      if (xp >= 0) vis |= 1u << (9 + k * 3 + j);
      if (xm < 0 || xp < 0) all = false;
    }
  }
  nbr[(size_t)9 * os + (size_t)c] = (int32_t)vis;
  interior[c] = all ? 1 : 0;
}

// The same table once more, one thread per MARKED COARSE CELL (= the 8 children it spawned).  Existence and x-break of a fine cell
// are properties of its parent alone (a marked cell has all 8 children; the run breaks after the i=1 child of a ghost pair), and the
// compressed table needs an index only for the nine row centres, which always lie in the parent's own x column.  So a thread reads
// the marks of the 27 coarse neighbours (two 27-bit masks E, G) and the child bases of the nine with da = 0, and everything else is
// compile-time indexing over (i,j,k) x (b,c): ~90 instructions per fine cell instead of ~1000 in the per-cell kernel, which had to
// select among its 8 parents with run-time indices (ncu: issue bound, 79 % issue active).
__global__ void __launch_bounds__(128) k_neighbours_oct(LV v, LV cv, const int32_t *__restrict__ cnbr, const int32_t *__restrict__ cidx,
                                                        const int4 *__restrict__ cbase, const int32_t *__restrict__ cpar, int M,
                                                        int32_t *__restrict__ nbr, uint8_t *__restrict__ interior)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  const int p = cpar[s];
  int px, py, pz; lv_coords(cv, p, px, py, pz);
  const int CM = (int)(cv.L - 1), L = (int)v.L;
  uint32_t E = 0, G = 0;                       // bit (dc+1)*9 + (db+1)*3 + (da+1): that coarse neighbour has children / they are a ghost pair
  int4 CB0[9];                                 // child bases of the neighbours with da = 0, index (dc+1)*3 + (db+1)
#pragma unroll
  for (int t = 0; t < 27; t++) {
    const int da = t % 3 - 1, db = (t / 3) % 3 - 1, dc = t / 9 - 1;
    int qc;
    if (t == 13) qc = p;
    else if (cv.dense) qc = (int)lv_key(cv, (px + da) & CM, (py + db) & CM, (pz + dc) & CM);
    else qc = nb_get(cv, cnbr, p, (dc + 1) * 3 + (db + 1), da + 1, px);
    const int ci = qc >= 0 ? cidx[qc] : -1;
    if (ci >= 0) { E |= 1u << t; if (ci & 0x40000000) G |= 1u << t; }
    if (da == 0) CB0[(dc + 1) * 3 + (db + 1)] = ci >= 0 ? cbase[ci & 0x3fffffff] : make_int4(-1, -1, -1, -1);
  }
  const int4 own = CB0[4];
  const size_t os = (size_t)v.ncell;
#pragma unroll
  for (int k = 0; k < 2; k++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int jk = k * 2 + j;
      const int c0 = jk == 0 ? own.x : jk == 1 ? own.y : jk == 2 ? own.z : own.w;      // children (i=0,1) of this (j,k) pair: c0, c0+1
      uint32_t vis[2] = { 0u, 0u };
      bool     all[2] = { true, true };
#pragma unroll
      for (int cc = -1; cc <= 1; cc++)
#pragma unroll
        for (int b = -1; b <= 1; b++) {
          const int q = (cc + 1) * 3 + (b + 1);
          const int ty = j + b, tz = k + cc;                                            // -1..2
          const int db = (ty >> 1), dc = (tz >> 1), comp = (tz & 1) * 2 + (ty & 1);
          const int rowbit = (dc + 1) * 9 + (db + 1) * 3;                                // + (da+1)
          const bool plane_ok = (E >> ((dc + 1) * 9 + 3 + 1)) & 1u;                      // the (x, y, z+cc) cell: parent (0, 0, dc)
          const bool row_ok = plane_ok && ((E >> (rowbit + 1)) & 1u);
          const int4 cb = CB0[(dc + 1) * 3 + (db + 1)];
          const int  r0 = comp == 0 ? cb.x : comp == 1 ? cb.y : comp == 2 ? cb.z : cb.w;
          int2 rm = make_int2(row_ok ? r0 : -1, row_ok ? r0 + 1 : -1);
          *reinterpret_cast<int2 *>(nbr + (size_t)q * os + (size_t)c0) = rm;
          if (!row_ok) { all[0] = all[1] = false; continue; }
          const bool gh0 = (G >> (rowbit + 1)) & 1u;                                    // ghost flag of the row centre's parent (da = 0)
          // child i = 0 (x = 2 px): x-1 is the i=1 child of the da=-1 parent, x+1 its own sibling (i=1, da=0)
          {
            const bool em = (E >> (rowbit + 0)) & 1u, brk_m = (G >> (rowbit + 0)) & 1u;  // that neighbour is an i=1 child: break iff ghost
            const bool vm = em && ((2 * px > 0) ? !brk_m : true);
            const bool vp = true;                                                        // sibling exists; the centre (i=0) never ends a run
            if (vm) vis[0] |= 1u << q; else all[0] = false;
            if (vp) vis[0] |= 1u << (9 + q);
          }
          // child i = 1 (x = 2 px + 1): x-1 is its sibling (i=0, never a break), x+1 the i=0 child of the da=+1 parent
          {
            const bool ep = (E >> (rowbit + 2)) & 1u;
            const bool vm = true;
            const bool vp = ep && ((2 * px + 1 < L - 1) ? !gh0 : true);                  // the centre is an i=1 child: the run ends after it iff ghost
            if (vm) vis[1] |= 1u << q;
            if (vp) vis[1] |= 1u << (9 + q); else all[1] = false;
          }
        }
      *reinterpret_cast<int2 *>(nbr + (size_t)9 * os + (size_t)c0) = make_int2((int)vis[0], (int)vis[1]);
      *reinterpret_cast<uchar2 *>(interior + c0) = make_uchar2(all[0] ? 1 : 0, all[1] ? 1 : 0);
    }
}

// ------------------------------------------------------------------------------------------------
// R1: relink (relink.c:31-288) -- per particle: first (z,y,x)-ordered interior child that contains it
// ------------------------------------------------------------------------------------------------
// R1 relink of one particle: the cell of the finer level it moves to (-1: it stays), relink's face bits, its position
__device__ __forceinline__ int relink_one(const float4 *__restrict__ pos4, const float4 *__restrict__ lpos, const uint32_t *__restrict__ plist, const int32_t *__restrict__ pcell,
                                          uint64_t i, const LV &coa, const uint8_t *__restrict__ cmark, const int32_t *__restrict__ cidx, const int4 *__restrict__ cbase,
                                          const LV &fin, const uint8_t *__restrict__ finterior, uint8_t &dl, float4 &q)
{
  const int c = pcell[i];
  int res = -1;
  dl = 0;                                    // bit d: the child's coordinate d is floor(L*x_d) - 1 (particle exactly on the upper face)
  if (!cmark[c]) return -1;
  q = lpos ? lpos[i] : pos4[plist ? plist[i] : i];      // refinement levels: the level's contiguous copy (coalesced)
  int cx, cy, cz; lv_coords(coa, c, cx, cy, cz);
  const double Lf = (double)fin.L;
  const double t[3] = { (double)q.x * Lf, (double)q.y * Lf, (double)q.z * Lf };
  const int    b[3] = { 2 * cx, 2 * cy, 2 * cz };
  // child b+e (e = 0,1) contains the particle iff b+e <= t <= b+e+1 (inclusive on both faces, relink.c:153)
  const int4 cb = cbase[cidx[c] & 0x3fffffff];        // the cell's children on the fine level (k_make_children): no hash probe
  // common case: no coordinate sits exactly on a face of the fine grid, so exactly one child contains the particle
  const int e0 = (int)t[0] - b[0], e1 = (int)t[1] - b[1], e2 = (int)t[2] - b[2];
  if (t[0] != (double)(int)t[0] && t[1] != (double)(int)t[1] && t[2] != (double)(int)t[2] && (unsigned)(e0 | e1 | e2) <= 1u) {
    const int jk = e2 * 2 + e1;
    const int f = (jk == 0 ? cb.x : jk == 1 ? cb.y : jk == 2 ? cb.z : cb.w) + e0;
    return finterior[f] ? f : -1;
  }
  bool ok[3][2];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    ok[d][0] = (t[d] >= (double)b[d]) && (t[d] <= (double)(b[d] + 1));
    ok[d][1] = (t[d] >= (double)(b[d] + 1)) && (t[d] <= (double)(b[d] + 2));
  }
  for (int k = 0; k < 2 && res < 0; k++) {
    if (!ok[2][k]) continue;
    for (int j = 0; j < 2 && res < 0; j++) {
      if (!ok[1][j]) continue;
      for (int e = 0; e < 2 && res < 0; e++) {
        if (!ok[0][e]) continue;
        const int jk = k * 2 + j;
        const int f = (jk == 0 ? cb.x : jk == 1 ? cb.y : jk == 2 ? cb.z : cb.w) + e;
        if (finterior[f]) {
          res = f;
          dl = (uint8_t)(((int)t[0] != b[0] + e ? 1 : 0) | ((int)t[1] != b[1] + j ? 2 : 0) | ((int)t[2] != b[2] + k ? 4 : 0));
        }
      }
    }
  }
  return res;
}
__global__ void k_relink(const float4 *__restrict__ pos4, const float4 *__restrict__ lpos, const uint32_t *__restrict__ plist, const int32_t *__restrict__ pcell, uint64_t np,
                         LV coa, const uint8_t *__restrict__ cmark, const int32_t *__restrict__ cidx, const int4 *__restrict__ cbase,
                         LV fin, const uint8_t *__restrict__ finterior,
                         int32_t *__restrict__ newcell, uint8_t *__restrict__ moved, uint8_t *__restrict__ dlt)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= np) return;
  uint8_t dl; float4 q;
  const int res = relink_one(pos4, lpos, plist, pcell, i, coa, cmark, cidx, cbase, fin, finterior, dl, q);
  newcell[i] = res;
  moved[i]   = res >= 0 ? 1 : 0;
  dlt[i]     = dl;
}
// The prefix sum over the moved particles and their compaction into the finer level's list in ONE launch (single GPU): a warp owns
// RF_ITEMS x 32 consecutive particles, ranks its moved ones with ballots, the tile's total goes through the chained scan of scan.cuh,
// and the particle list, cells and level-local positions of the finer level are written directly -- no separate scan, no prefix array,
// and the count stays on the device (the tile list of the finer level is derived behind it before the host reads both counts back).
// Stable: the list keeps the Hilbert order.  (Fusing k_relink itself in as well was measured slower, 1.06 against 0.86 ms per 256^3
// pass: its chain of dependent loads wants more resident warps than a tile that has to stay alive for the look-back allows.)
constexpr int RF_THREADS = 512, RF_ITEMS = 8, RF_TILE = RF_THREADS * RF_ITEMS;
__global__ void __launch_bounds__(RF_THREADS)
k_compact_fused(const uint32_t *__restrict__ plist, const int32_t *__restrict__ newcell, const uint8_t *__restrict__ moved, const uint8_t *__restrict__ dlt, uint64_t np,
                const float4 *__restrict__ pos4, const float4 *__restrict__ lpos_in, uint32_t *__restrict__ plist_out, int32_t *__restrict__ pcell_out,
                float4 *__restrict__ lpos_out, int8_t *__restrict__ owner, int8_t newlevel,
                unsigned long long *__restrict__ state, unsigned *__restrict__ ctr, unsigned nblk, int *__restrict__ total)
{
  __shared__ int wsum[RF_THREADS / 32];
  __shared__ int s_prefix;
  const unsigned bid = sc_ticket(ctr);
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t wbase = (uint64_t)bid * RF_TILE + (uint64_t)w * (32 * RF_ITEMS);
  unsigned bal[RF_ITEMS];
  int      wt = 0;
#pragma unroll
  for (int k = 0; k < RF_ITEMS; k++) {
    const uint64_t i = wbase + (uint64_t)(k * 32 + lane);
    bal[k] = __ballot_sync(0xffffffffu, i < np && moved[i] != 0);
    wt += __popc(bal[k]);
  }
  if (lane == 0) wsum[w] = wt;
  __syncthreads();
  if (w == 0) {
    const int x = lane < RF_THREADS / 32 ? wsum[lane] : 0;
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    const int agg = __shfl_sync(0xffffffffu, inc, 31);
    if (lane < RF_THREADS / 32) wsum[lane] = inc - x;
    const int prefix = sc_lookback_prefix(state, bid, agg, nblk, total);
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  int o = s_prefix + wsum[w];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < RF_ITEMS; k++) {
    if ((bal[k] >> lane) & 1u) {
      const uint64_t i = wbase + (uint64_t)(k * 32 + lane);
      const int      dst = o + __popc(bal[k] & lt);
      const uint32_t p = plist ? plist[i] : (uint32_t)i;
      plist_out[dst] = p; pcell_out[dst] = newcell[i];
      float4 q = lpos_in ? lpos_in[i] : pos4[p]; q.w = __int_as_float((int)dlt[i]);      // .w carries relink's face bits (see k_compact_moved)
      lpos_out[dst] = q;
      owner[p] = newlevel;
    }
    o += __popc(bal[k]);
  }
  sc_finish(state, ctr, nblk);
}

__global__ void k_dbg_compare(const int32_t *__restrict__ a, const int32_t *__restrict__ b, uint64_t n, unsigned long long *__restrict__ out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (a[i] >= 0) atomicAdd(&out[1], 1ull);
  if (a[i] != b[i]) { atomicAdd(&out[0], 1ull); atomicMin(&out[2], (unsigned long long)i); }
}
__global__ void k_compact_moved(const uint32_t *__restrict__ plist, const int32_t *__restrict__ newcell, const uint8_t *__restrict__ moved,
                                const int *__restrict__ S, uint64_t np, uint32_t *__restrict__ plist_out, int32_t *__restrict__ pcell_out,
                                int8_t *__restrict__ owner, int8_t newlevel, const float4 *__restrict__ pos4, const float4 *__restrict__ lpos_in,
                                float4 *__restrict__ lpos_out, const uint8_t *__restrict__ dlt)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= np || !moved[i]) return;
  uint32_t p = plist ? plist[i] : (uint32_t)i;
  plist_out[S[i]] = p; pcell_out[S[i]] = newcell[i];
  // level-local contiguous copy: TMA-stageable, no index indirection in the deposit.  The number-density deposit has no use for
  // the weight, so .w carries relink's face bits (cell coordinate = floor(L*x) - bit) and the deposit needs no cell-key gather.
  float4 q = lpos_in ? lpos_in[i] : pos4[p]; q.w = __int_as_float((int)dlt[i]);
  lpos_out[S[i]] = q;
  owner[p] = newlevel;
}

__global__ void k_count_owner(const int8_t *__restrict__ owner, uint64_t n, unsigned long long *__restrict__ cnt)
{
  __shared__ unsigned int h[64];
  if (threadIdx.x < 64) h[threadIdx.x] = 0;
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const int l = owner[i];
    const unsigned peers = __match_any_sync(__activemask(), l);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[l & 63], (unsigned)__popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < 64 && h[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// host side of the level loop
// ------------------------------------------------------------------------------------------------
template <typename T> static T *dalloc(size_t n)
{
  T *p = nullptr;
  p = static_cast<T *>(ahf::cache_alloc((n ? n : 1) * sizeof(T)));
  return p;
}
static inline unsigned nblk(uint64_t n, int b) { return (unsigned)((n + b - 1) / b); }

// rows / planes of a level that hold at least one marked cell (S = exclusive prefix sum over the marks, *Mtot its total)
__global__ void k_count_marked(LV v, const int *__restrict__ S, const int *__restrict__ Mtot, const int32_t *__restrict__ row_c0,
                               const int32_t *__restrict__ plane_r0, int nrow, int nplane, int *__restrict__ out)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = *Mtot;
  auto Sat = [&](long long i) { return i >= v.ncell ? M : S[i]; };
  bool rm = false, pm = false;
  if (r < nrow) {
    const long long c0 = v.dense ? ((long long)r << v.logL) : row_c0[r], c1 = v.dense ? c0 + v.L : row_c0[r + 1];
    rm = Sat(c1) > Sat(c0);
  }
  if (r < nplane) {
    const long long c0 = v.dense ? ((long long)r << (2 * v.logL)) : row_c0[plane_r0[r]], c1 = v.dense ? c0 + v.L * v.L : row_c0[plane_r0[r + 1]];
    pm = Sat(c1) > Sat(c0);
  }
  const unsigned br = __ballot_sync(0xffffffffu, rm), bp = __ballot_sync(0xffffffffu, pm);
  if ((threadIdx.x & 31) == 0) { if (br) atomicAdd(&out[0], __popc(br)); if (bp) atomicAdd(&out[1], __popc(bp)); }
}

// plane tables, run bounds, tested flag and run flags of a sorted list of row keys (z * L + y).  nplane < 0: counted here (one
// read-back).  The list is the level's own rows on one GPU, and the rows of ALL ranks when the box is split (the reference's run
// structure is not local: whether a row / plane is tested depends on rows and planes arbitrarily far away, refine_grid.c:346-405,
// :645-703).
static void rows_tested_flags(ahfgpu_ctx *c, const uint64_t *rowkey, int64_t nrow, int64_t nplane, long long L, int logL, uint8_t *tested, uint8_t *flags,
                              int32_t **rowplane_out, int32_t **plane_r0_out)
{
  DevBuf<uint8_t> head; DevBuf<int> hs, bs;
  int32_t *rowplane = nullptr, *plane_r0 = nullptr, *pz = nullptr;
  if (nplane < 0 || c->env.seg_v1) {
    head.reserve(nrow); hs.reserve(nrow);
    LAUNCH(c, k_plane_heads, nblk(nrow, 256), 256, 0, rowkey, (int)nrow, logL, head.p);
    if (nplane < 0) nplane = exclusive_scan<uint8_t>(c, head.p, hs.p, nrow);
    else exclusive_scan_async<uint8_t>(c, head.p, hs.p, nrow, nullptr, bs);
    rowplane = dalloc<int32_t>(nrow); plane_r0 = dalloc<int32_t>(nplane + 1);
    LAUNCH(c, k_plane_fill, nblk(nrow, 256), 256, 0, (int)nrow, head.p, hs.p, rowplane, plane_r0, (int)nplane);
  } else {
    rowplane = dalloc<int32_t>(nrow); plane_r0 = dalloc<int32_t>(nplane + 1); pz = dalloc<int32_t>(nplane);
    seg_heads_async(c, (uint64_t)nrow, PlaneSeg{ rowkey, logL, (int)nrow, rowplane, plane_r0, pz }, nullptr);    // one launch: heads + scan + fill + z of the planes
  }
  int32_t *rq0 = dalloc<int32_t>(nrow), *rq1 = dalloc<int32_t>(nrow), *pp0 = dalloc<int32_t>(nplane), *pp1 = dalloc<int32_t>(nplane);
  LAUNCH(c, k_row_runs, nblk(nrow, 256), 256, 0, rowkey, plane_r0, rowplane, (int)nrow, rq0, rq1);
  if (!pz) { pz = dalloc<int32_t>(nplane); LAUNCH(c, k_plane_z, nblk(nplane, 128), 128, 0, rowkey, plane_r0, (int)nplane, logL, pz); }
  LAUNCH(c, k_plane_runs, nblk(nplane, 128), 128, 0, pz, (int)nplane, pp0, pp1);
  LAUNCH(c, k_row_tested, nblk(nrow, 256), 256, 0, rowkey, plane_r0, rowplane, rq0, rq1, pp0, pp1, (int)nrow, (int)nplane, L, logL, tested, flags);
  ahf::dfree(rq0); ahf::dfree(rq1); ahf::dfree(pp0); ahf::dfree(pp1); ahf::dfree(pz);      // stream-ordered block cache: no host sync needed
  head.release(); hs.release(); bs.release();
  if (rowplane_out) *rowplane_out = rowplane; else ahf::dfree(rowplane);
  if (plane_r0_out) *plane_r0_out = plane_r0; else ahf::dfree(plane_r0);
}

// lv.nrow / lv.nplane are known before the level exists (4 rows per marked coarse row, 2 planes per marked coarse plane: counted
// next to the prefix sum over the marks and read back with it), so nothing here waits for the device.
// local_only: only the level's own index tables (crow, rowkey, row_c0); the tested flags follow from the rows of all ranks
static void build_rows_planes(ahfgpu_ctx *c, Level &lv, bool with_tested)
{
  LV v = view(lv);
  const int nc = (int)lv.ncell;
  DevBuf<uint8_t> head; DevBuf<int> hs, bs;
  lv.crow = dalloc<int32_t>(nc); lv.rowkey = dalloc<uint64_t>(lv.nrow); lv.row_c0 = dalloc<int32_t>(lv.nrow + 1);
  if (c->env.seg_v1) {
    head.reserve(nc); hs.reserve(nc);
    LAUNCH(c, k_row_heads, nblk(nc, 256), 256, 0, lv.ckey, nc, v.logL, head.p);
    exclusive_scan_async<uint8_t>(c, head.p, hs.p, nc, nullptr, bs);
    LAUNCH(c, k_row_fill, nblk(nc, 256), 256, 0, lv.ckey, nc, v.logL, head.p, hs.p, lv.crow, lv.rowkey, lv.row_c0, (int)lv.nrow);
  } else
    seg_heads_async(c, (uint64_t)nc, RowSeg{ lv.ckey, v.logL, nc, lv.crow, lv.rowkey, lv.row_c0 }, nullptr);      // one launch: heads + scan + fill
  head.release(); hs.release(); bs.release();
  lv.row_tested = dalloc<uint8_t>(lv.nrow); lv.row_flags = dalloc<uint8_t>(lv.nrow);
  if (with_tested) rows_tested_flags(c, lv.rowkey, lv.nrow, lv.nplane, (long long)lv.L, v.logL, lv.row_tested, lv.row_flags, &lv.rowplane, &lv.plane_r0);
  else {
    // the plane index tables of the level's own rows are still needed (child order of the next level); tested / flags are set later
    DevBuf<uint8_t> t2, f2;
    t2.reserve(lv.nrow); f2.reserve(lv.nrow);
    rows_tested_flags(c, lv.rowkey, lv.nrow, lv.nplane, (long long)lv.L, v.logL, t2.p, f2.p, &lv.rowplane, &lv.plane_r0);
    t2.release(); f2.release();
  }
}

// ---- ONE box on several ranks: which cells / rows / moved particles belong to the rank's own key range
__device__ __forceinline__ bool cell_owned(const LV &v, int c, const uint8_t *__restrict__ own3, int bd, int rank)
{
  int x, y, z; lv_coords(v, c, x, y, z);
  const int s = v.logL - bd;
  return own3[((((size_t)(z >> s)) << bd) | (size_t)(y >> s)) << bd | (size_t)(x >> s)] == (uint8_t)rank;
}
__global__ void k_owned_cells(LV v, const uint8_t *__restrict__ own3, int bd, int rank, uint8_t *__restrict__ owned)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < v.ncell) owned[c] = cell_owned(v, c, own3, bd, rank) ? 1 : 0;
}
__global__ void k_count_owned_marked(LV v, const uint8_t *__restrict__ mark, const uint8_t *__restrict__ own3, int bd, int rank, int *__restrict__ out)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool hit = c < v.ncell && mark[c] && cell_owned(v, c, own3, bd, rank);
  const unsigned b = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}
__global__ void k_rows_owned(LV v, const int32_t *__restrict__ crow, const uint8_t *__restrict__ own3, int bd, int rank, uint8_t *__restrict__ rowown)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < v.ncell && cell_owned(v, c, own3, bd, rank)) rowown[crow[c]] = 1;
}
__global__ void k_compact_rows(const uint64_t *__restrict__ rowkey, const uint8_t *__restrict__ rowown, const int *__restrict__ pos, int nrow, uint64_t *__restrict__ out)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nrow && rowown[r]) out[pos[r]] = rowkey[r];
}
__global__ void k_count_owned_moved(const uint32_t *__restrict__ plist, const uint8_t *__restrict__ moved, uint64_t np, uint32_t own_lo, uint32_t own_hi, int *__restrict__ out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  bool hit = false;
  if (i < np && moved[i]) { const uint32_t p = plist ? plist[i] : (uint32_t)i; hit = p >= own_lo && p < own_hi; }
  const unsigned b = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}
__global__ void k_unique_heads(const uint64_t *__restrict__ k, uint64_t n, uint8_t *__restrict__ head)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || k[i] != k[i - 1]) ? 1 : 0;
}
__global__ void k_unique_fill(const uint64_t *__restrict__ k, uint64_t n, const uint8_t *__restrict__ head, const int *__restrict__ pos, uint64_t *__restrict__ out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n && head[i]) out[pos[i]] = k[i];
}
// R sorted lists of row keys, concatenated (list p = [off[p], off[p+1])): every element finds its place in the merged order
struct RowLists { int R; int off[33]; };
__global__ void k_rows_rank(const uint64_t *__restrict__ all, RowLists RL, uint64_t *__restrict__ merged, uint8_t *__restrict__ first)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= RL.off[RL.R]) return;
  int p = 0;
  while (i >= RL.off[p + 1]) p++;
  const uint64_t k = all[i];
  int  place = i - RL.off[p];
  bool dup = false;
  for (int q = 0; q < RL.R; q++) {
    if (q == p) continue;
    int lo = RL.off[q], hi = RL.off[q + 1];
    const int b = lo;
    // q < p: elements <= k come first (and an equal one makes k a duplicate); q > p: elements < k
    if (q < p) { while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (all[mid] <= k) lo = mid + 1; else hi = mid; } dup |= (lo > b && all[lo - 1] == k); }
    else       { while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (all[mid] < k) lo = mid + 1; else hi = mid; } }
    place += lo - b;
  }
  merged[place] = k;
  first[place] = dup ? 0 : 1;
}

// tested flag / run flags of the level's own rows from the table of ALL rows (rows nobody owns exist only in the rank's ghost fringe)
__global__ void k_map_rows(const uint64_t *__restrict__ rowkey, int nrow, const uint64_t *__restrict__ grow, int ng, const uint8_t *__restrict__ gtested,
                           const uint8_t *__restrict__ gflags, uint8_t *__restrict__ tested, uint8_t *__restrict__ flags)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  const uint64_t k = rowkey[r];
  int lo = 0, hi = ng;
  while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (grow[mid] < k) lo = mid + 1; else hi = mid; }
  const bool found = lo < ng && grow[lo] == k;
  tested[r] = found ? gtested[lo] : 0;
  flags[r] = found ? gflags[lo] : 0;
}

// the rows of a level that hold at least one cell of the rank's own key range: compacted keys (device, caller frees) and their number
static int64_t owned_rows(ahfgpu_ctx *c, Level &lv, uint64_t **send_out)
{
  Stage sto(c, "rows_owned", lv.nrow, c->env.stages);
  Comm *cm = c->comm; Slab *S = c->slab;
  DevBuf<uint8_t> ro; DevBuf<int> pos, bs, tot;
  ro.reserve(lv.nrow); pos.reserve(lv.nrow); tot.reserve(1);
  CUDA_CHECK(cudaMemsetAsync(ro.p, 0, lv.nrow, c->stream));
  CUDA_CHECK(cudaMemsetAsync(tot.p, 0, sizeof(int), c->stream));
  LAUNCH(c, k_rows_owned, nblk(lv.ncell, 256), 256, 0, view(lv), lv.crow, S->own3, S->bd, cm->rank, ro.p);
  exclusive_scan_async<uint8_t>(c, ro.p, pos.p, lv.nrow, tot.p, bs);
  uint64_t *send = dalloc<uint64_t>(lv.nrow);
  LAUNCH(c, k_compact_rows, nblk(lv.nrow, 256), 256, 0, lv.rowkey, ro.p, pos.p, (int)lv.nrow, send);
  int h = 0;
  read_back(c, &h, tot.p, sizeof(int));
  ro.release(); pos.release(); bs.release(); tot.release();
  *send_out = send;
  return h;
}

// rows of all ranks -> tested / run flags of this rank's rows.  Every rank contributes the rows that hold at least one of ITS cells
// (`send`, nrows_rank[rank] keys); lv == nullptr: the rank has no cells on this level and only takes part in the exchange.
static void rows_from_all_ranks(ahfgpu_ctx *c, Level *lv, const uint64_t *send, const std::vector<int64_t> &nrows_rank, long long L, int logL)
{
  Comm *cm = c->comm;
  const int R = cm->nranks;
  std::vector<size_t> bytes(R), off(R + 1, 0);
  for (int p = 0; p < R; p++) { bytes[p] = (size_t)nrows_rank[p] * 8; off[p + 1] = off[p] + bytes[p]; }
  const int64_t nall = (int64_t)(off[R] / 8);
  uint64_t *all = dalloc<uint64_t>(nall);
  {
    Stage st(c, "rows_allgather", (int64_t)nall * 8, c->env.stages);
    cm->allgatherv(c, send, all, bytes.data(), off.data());
  }
  if (lv && lv->nrow > 0) {
    Stage stm(c, "rows_merge", nall, c->env.stages);
    if (nall == 0) AHF_FAIL("a level without rows on any rank");
    // merge of the R sorted lists without a sort: the place of an element among all elements is its index in its own list plus, for every
    // other list, the number of elements below it (binary searches); an element that also occurs in an earlier list is a duplicate
    RowLists RL;
    RL.R = R;
    for (int p = 0; p < R; p++) { RL.off[p] = (int)(off[p] / 8); }
    RL.off[R] = (int)nall;
    uint64_t *merged = dalloc<uint64_t>(nall);
    DevBuf<uint8_t> first; DevBuf<int> pos;
    first.reserve(nall); pos.reserve(nall);
    LAUNCH(c, k_rows_rank, nblk(nall, 256), 256, 0, all, RL, merged, first.p);
    const int ng = exclusive_scan<uint8_t>(c, first.p, pos.p, (uint64_t)nall);
    uint64_t *grow = dalloc<uint64_t>(ng);
    LAUNCH(c, k_unique_fill, nblk(nall, 256), 256, 0, merged, (uint64_t)nall, first.p, pos.p, grow);
    uint8_t *gt = dalloc<uint8_t>(ng), *gf = dalloc<uint8_t>(ng);
    rows_tested_flags(c, grow, ng, -1, L, logL, gt, gf, nullptr, nullptr);
    LAUNCH(c, k_map_rows, nblk(lv->nrow, 256), 256, 0, lv->rowkey, (int)lv->nrow, grow, ng, gt, gf, lv->row_tested, lv->row_flags);
    first.release(); pos.release();
    ahf::dfree(merged); ahf::dfree(grow); ahf::dfree(gt); ahf::dfree(gf);
  }
  ahf::dfree(all);
}

static std::string lvl_name(const char *base, int lev) { char b[48]; snprintf(b, sizeof(b), "%s_L%d", base, lev); return b; }
// per-level stage timers (scripts/stage_breakdown.py) are opt-in: AHFGPU_LEVEL_STAGES=1 (MeshEnv); the per-pass totals are always taken


// per-DEVICE setup of the mesh kernels (function attributes and __constant__ symbols belong to the current device): called once per
// device from ahfgpu_init
void mesh_device_init()
{
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, DR_SMEM));
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_dom<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM));
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_dom2, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM));
#ifdef AHFGPU_EXPERIMENTS      // timing-only variants of the domain kernel (wrong densities on purpose): never in the shipped library
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_dom<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM));
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_dom<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM));
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_dom<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM));
  CUDA_CHECK(cudaFuncSetAttribute(k_deposit_dom<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM));
#endif
  uint16_t off[27];
  for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) for (int a = 0; a < 3; a++) off[k * 9 + j * 3 + a] = (uint16_t)(4 * ((k * DT_H + j) * DT_H + a));
  CUDA_CHECK(cudaMemcpyToSymbol(c_dom_off, off, sizeof(off)));
}

static void deposit_level(ahfgpu_ctx *c, Level &lv)
{
  const int lev_id = (int)c->levels.size() - 1;
  Stage st(c, "deposit", lv.npart_dep, c->env.stages);
  Stage stl(c, lvl_name("deposit", lev_id).c_str(), lv.npart_dep, c->env.level_stages);
  LV v = view(lv);
  const int nc = (int)lv.ncell;
  DevBuf<unsigned long long> acc;
  acc.reserve(nc);
  CUDA_CHECK(cudaMemsetAsync(acc.p, 0, sizeof(unsigned long long) * nc, c->stream));
  const bool generic_only = c->env.generic_deposit;
  const bool dom_v1       = c->env.deposit_v1;           // previous float-weight domain kernel (A/B timing)
  // choice of kernel and fixed-point scale from the counts of the WHOLE box (g_*): every rank of a split box rounds like one GPU
  const bool tiles_dense  = lv.dense && lv.L >= 2 * DT_T && lv.g_npart_dep > 0 && !generic_only;
  const int    S = (tiles_dense && !dom_v1) ? (c->env.dom_v2 ? D2_S : c->env.dom_s) : fx_shift_for(lv.masstopartdens);    // k_deposit_dom works in 2^-S units
  const double fxscale = (double)(1ull << S);
  const bool tiles_sparse = !lv.dense && lv.lpos && lv.g_npart_dep >= 2048 && v.logL - 4 <= 20 && !generic_only;
  if ((tiles_dense || tiles_sparse) && lv.npart_dep > 0) {
    // tiles are Hilbert cells of (logL - 4) bits per dimension: contiguous ranges of the (level's) particle list
    const int tbits = v.logL - 4;
    int ntile = 0;
    DevBuf<int32_t> tstart; DevBuf<int> nchunk, woff, bs, tot, hs; DevBuf<int2> work; DevBuf<int4> work4; DevBuf<uint32_t> tlist; DevBuf<uint8_t> head;
    tot.reserve(1);
    if (tiles_dense) {
      ntile = 1 << (3 * tbits);
      tstart.reserve(ntile + 1);
      LAUNCH(c, k_tile_starts, nblk(ntile + 1, 256), 256, 0, c->keys, (int64_t)c->n, tbits, ntile, tstart.p);
    } else {
      const uint64_t np = (uint64_t)lv.npart_dep;
      const int sh = 3 * (21 - tbits);
      if (lv.ntile >= 0 && lv.tlist && lv.tstart) {        // found behind the relink that made the list (amr_build)
        ntile = lv.ntile;
        tlist.adopt(lv.tlist); tstart.adopt(lv.tstart);
        lv.tlist = nullptr; lv.tstart = nullptr; lv.ntile = -1;
      } else if (c->env.seg_v1) {
        head.reserve(np); hs.reserve(np);
        LAUNCH(c, k_lvl_tile_heads, nblk(np, 256), 256, 0, c->keys, lv.plist, np, sh, head.p);
        ntile = exclusive_scan<uint8_t>(c, head.p, hs.p, np);
        tlist.reserve(ntile); tstart.reserve(ntile + 1);
        LAUNCH(c, k_lvl_tile_fill, nblk(np, 256), 256, 0, c->keys, lv.plist, np, sh, head.p, hs.p, ntile, tlist.p, tstart.p);
      } else {
        // one launch: heads + scan + fill.  The lists are sized before the number of tiles is known: a tile is named by the particle's
        // KEY, the particle's cell lies in it or -- a particle exactly on a cell face, relink's face bits -- in one of the 7 tiles towards
        // lower coordinates, and a tile with a cell holds a whole oct: at most 8 * ncell / 8 tiles
        const size_t cap = (size_t)std::min<uint64_t>(np, (uint64_t)lv.ncell + 1);
        tlist.reserve(cap); tstart.reserve(cap + 1);
        seg_heads_async(c, np, TileSeg{ c->keys, lv.plist, sh, np, tlist.p, tstart.p, nullptr }, tot.p);
        read_back(c, &ntile, tot.p, sizeof(int));
      }
    }
    // the work list has at most one entry per tile plus one per full chunk: launch that many CTAs, the kernels read the real
    // length from the device (no host read-back)
    const int W = ntile + (int)(lv.npart_dep / DT_CHUNK) + 1;
    const bool dom_default = tiles_dense && !dom_v1 && !(c->env.dom_v2 && c->env.dom_variant == 0);
    const bool fused_work = !c->env.seg_v1 && (tiles_sparse || dom_default);
    if (!fused_work) {
      nchunk.reserve(ntile); woff.reserve(ntile);
      LAUNCH(c, k_tile_nchunk, nblk(ntile, 256), 256, 0, tstart.p, ntile, nchunk.p);
      exclusive_scan_async<int>(c, nchunk.p, woff.p, ntile, tot.p, bs);
      work.reserve(W);
      LAUNCH(c, k_tile_work, nblk(ntile, 256), 256, 0, nchunk.p, woff.p, ntile, work.p);
    } else if (tiles_sparse) {
      work.reserve(W);
      scan_emit_async(c, (uint64_t)ntile, TileWork{ tstart.p, work.p }, tot.p);
    }
    const float4 *dom_pos = c->pos4;
    DevBuf<float4> pos_c;
    if (tiles_dense && !dom_v1 && getenv("AHFGPU_DOM_CELLSORT")) {
      const uint64_t np = c->n;
      DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> v0, v1;
      k0.reserve(np); k1.reserve(np); v0.reserve(np); v1.reserve(np); pos_c.reserve(np);
      LAUNCH(c, k_cellsort_keys, nblk(np, 256), 256, 0, c->pos4, c->keys, np, v.logL, tbits, k0.p, v0.p);
      uint64_t *ks; uint32_t *vs;
      radix_sort_pairs(c, k0.p, v0.p, k1.p, v1.p, np, ((3 * tbits + 12 + 7) / 8) * 8, &ks, &vs);
      LAUNCH(c, k_gather_f4, nblk(np, 256), 256, 0, c->pos4, vs, np, pos_c.p);
      dom_pos = pos_c.p;
      k0.release(); k1.release(); v0.release(); v1.release();
    }
    {
      Stage sk(c, lv.dense ? "deposit_dom_kernel" : "deposit_ref_kernel", lv.npart_dep, lv.dense || c->env.stages);
      Stage skl(c, lvl_name("depk", lev_id).c_str(), W, c->env.level_stages);
      if (tiles_dense && !dom_v1 && c->env.dom_v2 && c->env.dom_variant == 0) {
        work4.reserve(W);
        LAUNCH(c, k_tile_work4, nblk(ntile, 256), 256, 0, tstart.p, nchunk.p, woff.p, ntile, tbits, work4.p);
        DevBuf<unsigned int> d2s;
        if (c->env.dom2_stats) { d2s.reserve(2); CUDA_CHECK(cudaMemsetAsync(d2s.p, 0, 2 * sizeof(unsigned int), c->stream)); }
        LAUNCH(c, k_deposit_dom2, (unsigned)W, D2_THREADS, D2_SMEM, c->pos4, work4.p, tot.p, (int)lv.L, v.logL, acc.p, c->env.dom2_heavy ? 1 : 0, d2s.p);
        if (c->env.dom2_stats) {
          unsigned int h[2] = { 0, 0 };
          read_back(c, h, d2s.p, sizeof(h));
          c->stage_cnt_extra["dom2_heavy_ctas"] = h[0]; c->stage_cnt_extra["dom2_failed_light"] = h[1];
          d2s.release();
        }
      }
      else if (tiles_dense && !dom_v1) {
        const int var = c->env.dom_variant;
        const int rmax = c->env.dom_rmax;           // slices with at most this many distinct cells take the run-reduction path (measured: 1 = whole warp in one cell is best; partial-mask REDUX costs more than the conflicts it removes)
        work4.reserve(W);
        if (fused_work) scan_emit_async(c, (uint64_t)ntile, TileWork4{ tstart.p, tbits, work4.p }, tot.p);
        else LAUNCH(c, k_tile_work4, nblk(ntile, 256), 256, 0, tstart.p, nchunk.p, woff.p, ntile, tbits, work4.p);
        int nsm = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->dev));
        // AHFGPU_DOM_PERSIST=1: two persistent CTAs per SM striding over the items (measured slower: static striding loses the
        // hardware's dynamic balance between light and heavy tiles); default: one CTA per item
        const unsigned grid = c->env.dom_persist ? (unsigned)std::min(W, 2 * nsm) : (unsigned)W;
#define DOM_LAUNCH(V) LAUNCH(c, k_deposit_dom<V>, grid, DT_THREADS, DD_SMEM, dom_pos, work4.p, tot.p, (int)lv.L, v.logL, acc.p, 1u, rmax, 32 - S)
#ifdef AHFGPU_EXPERIMENTS
        if (var == 1) DOM_LAUNCH(1); else if (var == 2) DOM_LAUNCH(2); else if (var == 3) DOM_LAUNCH(3); else if (var == 4) DOM_LAUNCH(4); else DOM_LAUNCH(0);
#else
        if (var != 0) AHF_FAIL("AHFGPU_DOM_VARIANT needs a library built with -DAHFGPU_EXPERIMENTS (timing-only kernels, wrong densities)");
        DOM_LAUNCH(0);
#endif
#undef DOM_LAUNCH
      }
      else if (tiles_dense)
        LAUNCH(c, k_deposit_tiles<false>, (unsigned)W, DT_THREADS, DT_SMEM, c->pos4, tstart.p, work.p, (int)lv.L, v.logL, tbits, acc.p,
               (const uint32_t *)nullptr, (const int32_t *)nullptr, v, (const int32_t *)nullptr, (float)fxscale, tot.p);
      else if (c->env.sparse_v1 || (!c->env.sparse_v2 && (double)lv.g_npart_dep < 0.75 * (double)lv.g_ncell))
        // lane per particle: the thin tiles of the deepest levels (well under one particle per cell, a few hundred particles per
        // tile) have no runs to aggregate and want the larger CTA for the tile flush (measured: 0.23 vs 0.31 ms on level 6)
        LAUNCH(c, k_deposit_tiles<true>, (unsigned)W, DT_THREADS, DT_SMEM, lv.lpos, tstart.p, work.p, (int)lv.L, v.logL, tbits, acc.p,
               tlist.p, lv.pcell, v, lv.nbr, (float)fxscale, tot.p);
      else
        LAUNCH(c, k_deposit_runs, (unsigned)W, DR_THREADS, DR_SMEM, lv.lpos, tstart.p, work.p, (int)lv.L, v.logL, tbits, acc.p,
               tlist.p, lv.pcell, v, lv.nbr, (float)fxscale, tot.p);
    }
    pos_c.release();
    if (lv.dense) c->stage_cnt_extra["deposit_dom_ctas"] = W;
    tstart.release(); nchunk.release(); woff.release(); bs.release(); tot.release(); work.release(); work4.release(); tlist.release(); head.release(); hs.release();
  } else if (lv.npart_dep > 0 && !(tiles_dense || tiles_sparse)) {
    Stage sk(c, lv.dense ? "deposit_dom_kernel" : "deposit_ref_kernel", lv.npart_dep, lv.dense || c->env.stages);
    LAUNCH(c, k_deposit_generic, nblk(lv.npart_dep, 256), 256, 0, c->pos4, lv.plist, lv.pcell, (uint64_t)lv.npart_dep, v, lv.nbr, acc.p, fxscale);
  }
  LAUNCH(c, k_finish_dens, nblk(nc, 256), 256, 0, acc.p, lv.dens, nc, lv.masstopartdens / fxscale);
  acc.release();                                    // stream-ordered block cache: no host sync needed
}

static void alloc_cell_arrays(Level &lv)
{
  const size_t nc = (size_t)lv.ncell;
  lv.dens = dalloc<float>(nc); lv.tn = dalloc<uint8_t>(nc); lv.mark = dalloc<uint8_t>(nc); lv.count = dalloc<int32_t>(nc);
  CUDA_CHECK(cudaMemsetAsync(lv.mark, 0, nc, ahf::g_pool_stream));      // stream ordered: the block may be recycled memory that queued kernels still read
}

void amr_build(ahfgpu_ctx *c)
{
  c->env.read();
  Stage sall(c, "amr_total", (int64_t)c->n);
  c->free_levels();
  const uint64_t n = c->n;
  const ahfgpu_params &par = c->par;
  long long lmax = par.lgrid_max;
  if (lmax > (1 << 21) || lmax <= 0) lmax = (1 << 21);
  if (par.lgrid_dom > 1024) AHF_FAIL("dense domain grids above 1024^3 need 64-bit cell indices (not in this round)");
  // ONE box split over several contexts (slab.cu): the resident set is the rank's key range plus a ghost shell wide enough that every
  // level is exact on the rank's own cells; what is NOT local -- does any rank still refine, is the next level large enough, how many
  // particles deposit on it, which rows / planes exist -- is exchanged once per level (one small all-gather, one all-gather of row keys)
  Comm *cm = c->comm; Slab *S = c->slab;
  const bool split = cm != nullptr && S != nullptr;
  const int  R = split ? cm->nranks : 1;
  if (cm && !S) AHF_FAIL("a communicator is attached but the particles were not distributed: call ahfgpu_slab_distribute");
  const uint64_t n_box = c->n_total ? c->n_total : n;
  c->owner_level = dalloc<int8_t>(n);
  CUDA_CHECK(cudaMemsetAsync(c->owner_level, 0, n, c->stream));

  // ---- domain level (gen_domgrids, generate_grids.c:24-116)
  {
    Level d;
    d.L = par.lgrid_dom; d.ncell = d.L * d.L * d.L; d.dense = true;
    d.masstopartdens = ((double)d.L * (double)d.L * (double)d.L) / (double)n_box;
    d.critdens = par.nth_dom * d.masstopartdens;
    alloc_cell_arrays(d);
    d.npart_dep = (int64_t)n;
    d.g_ncell = d.ncell; d.g_npart_dep = (int64_t)n_box;
    d.pcell = dalloc<int32_t>(n);
    {
      Stage st(c, "ll", (int64_t)n, c->env.stages);
      LV v = view(d);
      if (n) LAUNCH(c, k_domain_cells, nblk(n, 256), 256, 0, c->pos4, n, (int)d.L, v.logL, d.pcell);
    }
    c->levels.push_back(d);
  }
  bool      active = true;                 // this rank still has cells on the current level
  long long curL = par.lgrid_dom;
  int       glev = 0;                      // index of the current level in the box-wide count
  for (;;) {
    const int lev = glev;
    if (active) deposit_level(c, c->levels.back());
    if (curL == lmax) break;                                           // generate_grids.c:307-308
    if (n_box == 0) break;                                             // empty box: the domain grid alone, nothing to flag
    int M = 0, h3[4] = { 0, 0, 0, 0 };
    DevBuf<int> S_;
    if (active) {
      Level &cur = c->levels.back();
      LV cv = view(cur);
      const int nc = (int)cur.ncell;
      {
        Stage st(c, "flag", nc, c->env.stages);
        Stage stl(c, lvl_name("flag", lev).c_str(), nc, c->env.level_stages);
        if (cur.dense && cur.L >= 32 && !c->env.testnode_v1)
          LAUNCH(c, k_test_node_dense, dim3((unsigned)(cur.L / 32), (unsigned)(cur.L / 8), (unsigned)(cur.L / 8)), 256, 0, cv, cur.dens, cur.critdens - 1.0, cur.tn);
        else
          LAUNCH(c, k_test_node, nblk(nc, 256), 256, 0, cv, cur.dens, cur.interior, cur.nbr, cur.critdens - 1.0, cur.tn);
        if (cur.dense) LAUNCH(c, k_mark_dense, nblk(nc, 256), 256, 0, cv, cur.tn, cur.mark);
        else LAUNCH(c, k_mark_sparse, nblk(nc, 256), 256, 0, cv, cur.tn, cur.interior, cur.crow, cur.row_c0, cur.row_tested, cur.mark);
      }
      Stage st(c, "refine", nc, c->env.stages);
      S_.reserve(nc);
      DevBuf<int> t3, bs;
      t3.reserve(4);
      CUDA_CHECK(cudaMemsetAsync(t3.p, 0, 4 * sizeof(int), c->stream));
      exclusive_scan_async<uint8_t, true>(c, cur.mark, S_.p, nc, t3.p, bs);            // marks are 0 / 1 (refined) / 2 (ghost pair): count the non-zero ones
      const int cnrow = cur.dense ? (int)(cur.L * cur.L) : (int)cur.nrow, cnplane = cur.dense ? (int)cur.L : (int)cur.nplane;
      LAUNCH(c, k_count_marked, nblk(cnrow, 256), 256, 0, cv, S_.p, t3.p, cur.row_c0, cur.plane_r0, cnrow, cnplane, t3.p + 1);
      if (split) LAUNCH(c, k_count_owned_marked, nblk(nc, 256), 256, 0, cv, cur.mark, S->own3, S->bd, cm->rank, t3.p + 3);
      read_back(c, h3, t3.p, 4 * sizeof(int));
      t3.release(); bs.release();
      M = h3[0];
      if (!split) h3[3] = M;
    }
    if (!split && M == 0) { S_.release(); break; }                     // refine_grid returned FALSE
    if ((long long)M * 8 > 2000000000ll) AHF_FAIL("refinement level exceeds 2^31 cells");
    // ---- next level, built locally from the rank's marks (exact on its own cells and well into the ghost shell)
    const bool built = active && M > 0;
    DevBuf<int32_t> newcell; DevBuf<uint8_t> moved, dlt; DevBuf<int> MS;
    int nmoved = 0, nmoved_own = 0;
    uint64_t *rows_send = nullptr; int64_t nrows_own = 0;
    if (built) {
      Level &cur = c->levels.back();
      LV cv = view(cur);
      const int nc = (int)cur.ncell;
      Stage st(c, "refine", nc, c->env.stages);
      Stage stl(c, lvl_name("refine", lev).c_str(), nc, c->env.level_stages);
      Level f;
      f.L = cur.L * 2; f.ncell = (int64_t)M * 8; f.dense = false;
      f.masstopartdens = cur.masstopartdens * CRITMULTI;               // generate_grids.c:164-170
      f.critdens = par.nth_ref * f.masstopartdens;
      f.ckey = dalloc<uint64_t>(f.ncell); f.xbreak = dalloc<uint8_t>(f.ncell);
      int flogL = cv.logL + 1;
      f.parent = dalloc<int32_t>(f.ncell); cur.cidx = dalloc<int32_t>(nc); cur.cbase = dalloc<int4>(M); cur.cpar = dalloc<int32_t>(M);
      LAUNCH(c, k_make_children, nblk(nc, 256), 256, 0, cv, cur.mark, S_.p, M, cur.crow, cur.row_c0, cur.dense ? nullptr : cur.rowplane,
             cur.plane_r0, (int)cur.nrow, f.ckey, f.xbreak, flogL, f.parent, cur.cidx, cur.cbase, cur.cpar);
      // hash
      // slots hold 8 x-consecutive cells; children come in x-pairs, so there are at most ncell/2 occupied slots
      uint64_t cap = 16; while (cap < (uint64_t)f.ncell + 2) cap <<= 1;
      f.hmask = cap - 1; f.hkey = dalloc<uint64_t>(cap); f.hval = dalloc<int32_t>(cap * 2);
      CUDA_CHECK(cudaMemsetAsync(f.hkey, 0xff, cap * sizeof(uint64_t), c->stream));          // empty slots: key ~0 (values are only read behind a key match)
      LAUNCH(c, k_hash_insert, nblk(f.ncell, 256), 256, 0, f.ckey, (int)f.ncell, f.hkey, reinterpret_cast<uint2 *>(f.hval), f.hmask);
      alloc_cell_arrays(f);
      f.interior = dalloc<uint8_t>(f.ncell); f.nbr = dalloc<int32_t>((size_t)f.ncell * 10);
      LV fv = view(f);
      if (c->env.nbr_v1) LAUNCH(c, k_neighbours, nblk(f.ncell, 128), 128, 0, fv, f.nbr, f.interior);      // hash probes (A/B timing)
      else if (c->env.nbr_v2) LAUNCH(c, k_neighbours_pc, nblk(f.ncell, 128), 128, 0, fv, f.parent, cv, cur.nbr, cur.cidx, cur.cbase, f.nbr, f.interior);   // per fine cell (A/B timing)
      else LAUNCH(c, k_neighbours_oct, nblk(M, 128), 128, 0, fv, cv, cur.nbr, cur.cidx, cur.cbase, cur.cpar, M, f.nbr, f.interior);
      if (c->env.debug_nbr) {                 // both constructions must give the same table
        DevBuf<int32_t> nb2; DevBuf<uint8_t> in2; DevBuf<unsigned long long> out;
        nb2.reserve((size_t)f.ncell * 10); in2.reserve(f.ncell); out.reserve(3);
        unsigned long long h0[3] = { 0, 0, ~0ull }, h[3];
        CUDA_CHECK(cudaMemcpyAsync(out.p, h0, sizeof(h0), cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, k_neighbours, nblk(f.ncell, 128), 128, 0, fv, nb2.p, in2.p);
        LAUNCH(c, k_dbg_compare, nblk((uint64_t)f.ncell * 10, 256), 256, 0, f.nbr, nb2.p, (uint64_t)f.ncell * 10, out.p);
        CUDA_CHECK(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        fprintf(stderr, "[nbr dbg] level %d: %llu of %llu neighbour entries differ between the parent/child and the hash construction (first at %llu)\n",
                lev + 1, h[0], (unsigned long long)f.ncell * 10, h[2]);
        nb2.release(); in2.release(); out.release();
      }
      f.nrow = 4ll * h3[1]; f.nplane = 2ll * h3[2];               // every marked coarse cell spawns 2x2x2 children
      build_rows_planes(c, f, !split);
      c->levels.push_back(f);
    }
    S_.release();
    // ---- relink, first half: which particles would move to the new level
    const bool fused_relink = !split && !c->env.relink_v1;
    if (built) {
      Level &coa = c->levels[c->levels.size() - 2];
      Level &fin = c->levels.back();
      Stage st(c, "relink", coa.npart_dep, c->env.stages);
      Stage stl(c, lvl_name("relink", lev).c_str(), coa.npart_dep, c->env.level_stages);
      const uint64_t np = (uint64_t)coa.npart_dep;
      newcell.reserve(np); moved.reserve(np); dlt.reserve(np);
      if (!fused_relink) MS.reserve(np);
      if (np && fused_relink)
        LAUNCH(c, k_relink, nblk(np, 256), 256, 0, c->pos4, coa.lpos, coa.plist, coa.pcell, np, view(coa), coa.mark, coa.cidx, coa.cbase, view(fin), fin.interior, newcell.p, moved.p, dlt.p);
      if (np && !fused_relink) {
        LAUNCH(c, k_relink, nblk(np, 256), 256, 0, c->pos4, coa.lpos, coa.plist, coa.pcell, np, view(coa), coa.mark, coa.cidx, coa.cbase, view(fin), fin.interior, newcell.p, moved.p, dlt.p);
        if (split) {
          DevBuf<int> t2, bs;
          t2.reserve(2);
          CUDA_CHECK(cudaMemsetAsync(t2.p, 0, 2 * sizeof(int), c->stream));
          exclusive_scan_async<uint8_t>(c, moved.p, MS.p, np, t2.p, bs);
          LAUNCH(c, k_count_owned_moved, nblk(np, 256), 256, 0, coa.plist, moved.p, np, (uint32_t)S->own_lo, (uint32_t)S->own_hi, t2.p + 1);
          int h2[2];
          read_back(c, h2, t2.p, sizeof(h2));
          nmoved = h2[0]; nmoved_own = h2[1];
          t2.release(); bs.release();
        } else {
          nmoved = exclusive_scan<uint8_t>(c, moved.p, MS.p, np);
          nmoved_own = nmoved;
        }
      }
      if (split) nrows_own = owned_rows(c, fin, &rows_send);
    }
    // ---- what the whole box decides
    long long g_M = h3[3], g_moved = nmoved_own;
    std::vector<int64_t> nrows_rank(R, 0);
    if (split) {
      Stage st(c, "level_allgather", 24 * R, c->env.stages);
      long long mine[3] = { (long long)h3[3], (long long)nmoved_own, (long long)nrows_own };
      std::vector<long long> all((size_t)3 * R);
      cm->allgather_host(c, mine, all.data(), sizeof(mine));
      g_M = 0; g_moved = 0;
      for (int p = 0; p < R; p++) { g_M += all[3 * p]; g_moved += all[3 * p + 1]; nrows_rank[p] = all[3 * p + 2]; }
    }
    const bool refined = g_M > 0;                                      // refine_grid returned TRUE somewhere
    const bool accepted = refined && g_M * 8 >= MIN_NNODES;            // generate_grids.c:231 / density.c:420
    if (!accepted) {
      if (built) { c->levels.back().free_all(); c->levels.pop_back(); }
      ahf::dfree(rows_send);
      newcell.release(); moved.release(); dlt.release(); MS.release();
      break;
    }
    if (split) {
      Stage st(c, "refine", 0, c->env.stages);
      int flogL = 0; while ((1ll << flogL) < curL * 2) flogL++;
      rows_from_all_ranks(c, built ? &c->levels.back() : nullptr, rows_send, nrows_rank, curL * 2, flogL);
      ahf::dfree(rows_send);
    }
    // ---- relink, second half
    if (built) {
      Level &coa = c->levels[c->levels.size() - 2];
      Level &fin = c->levels.back();
      Stage st(c, "relink", 0, c->env.stages);
      const uint64_t np = (uint64_t)coa.npart_dep;
      if (fused_relink) {
        // scan + compaction in one launch into lists sized for all np particles of the coarser level, then the deposit tiles of the new
        // list (the count still on the device), then ONE read-back of both counts
        Stage stl(c, lvl_name("relink", lev).c_str(), coa.npart_dep, c->env.level_stages);
        fin.plist = dalloc<uint32_t>(np); fin.pcell = dalloc<int32_t>(np); fin.lpos = dalloc<float4>(np);
        DevBuf<int> cnt;
        cnt.reserve(2);
        CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(int), c->stream));
        int h2[2] = { 0, 0 };
        if (np) {
          const unsigned nb = (unsigned)((np + RF_TILE - 1) / RF_TILE);
          scan_state_reserve(c, nb);
          LAUNCH(c, k_compact_fused, nb, RF_THREADS, 0, coa.plist, newcell.p, moved.p, dlt.p, np, c->pos4, coa.lpos, fin.plist, fin.pcell, fin.lpos,
                 c->owner_level, (int8_t)(lev + 1), c->scan_state, scan_state_ctr(c), nb, cnt.p);
          const int tbits = view(fin).logL - 4;
          const bool want_tiles = tbits <= 20 && tbits >= 0 && !c->env.generic_deposit && !c->env.seg_v1;
          if (want_tiles) {
            const size_t cap = (size_t)std::min<uint64_t>(np, (uint64_t)fin.ncell + 1);        // see deposit_level
            fin.tlist = dalloc<uint32_t>(cap); fin.tstart = dalloc<int32_t>(cap + 1);
            seg_heads_async(c, np, TileSeg{ c->keys, fin.plist, 3 * (21 - tbits), np, fin.tlist, fin.tstart, cnt.p }, cnt.p + 1, cnt.p);
          }
          read_back(c, h2, cnt.p, sizeof(h2));
          if (want_tiles) fin.ntile = h2[1];
        }
        cnt.release();
        nmoved = h2[0];
        if ((uint64_t)nmoved * 8 < np * 5) {                // much smaller than the bound (first refinement): exact lists, the big blocks go back to the cache
          const size_t m = (size_t)std::max(nmoved, 1);
          uint32_t *pl = dalloc<uint32_t>(m); int32_t *pc = dalloc<int32_t>(m); float4 *lp = dalloc<float4>(m);
          CUDA_CHECK(cudaMemcpyAsync(pl, fin.plist, sizeof(uint32_t) * nmoved, cudaMemcpyDeviceToDevice, c->stream));
          CUDA_CHECK(cudaMemcpyAsync(pc, fin.pcell, sizeof(int32_t) * nmoved, cudaMemcpyDeviceToDevice, c->stream));
          CUDA_CHECK(cudaMemcpyAsync(lp, fin.lpos, sizeof(float4) * nmoved, cudaMemcpyDeviceToDevice, c->stream));
          ahf::dfree(fin.plist); ahf::dfree(fin.pcell); ahf::dfree(fin.lpos);
          fin.plist = pl; fin.pcell = pc; fin.lpos = lp;
        }
        g_moved = nmoved;
      }
      fin.npart_dep = nmoved;
      fin.g_ncell = g_M * 8; fin.g_npart_dep = g_moved;
      if (!fused_relink) {
        fin.plist = dalloc<uint32_t>(nmoved); fin.pcell = dalloc<int32_t>(nmoved); fin.lpos = dalloc<float4>(nmoved);
        if (np) LAUNCH(c, k_compact_moved, nblk(np, 256), 256, 0, coa.plist, newcell.p, moved.p, MS.p, np, fin.plist, fin.pcell, c->owner_level, (int8_t)(lev + 1), c->pos4, coa.lpos, fin.lpos, dlt.p);
      }
    }
    newcell.release(); moved.release(); dlt.release(); MS.release();      // stream-ordered block cache: no host sync needed
    active = built;
    curL *= 2; glev++;
    if (glev + 1 >= 60) break;
  }
  c->g_nlevels = glev + 1;
  // final ownership counts (of the rank's own particles when the box is split)
  {
    DevBuf<unsigned long long> cnt;
    cnt.reserve(64);
    CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, 64 * sizeof(unsigned long long), c->stream));
    const uint64_t o0 = split ? S->own_lo : 0, on = split ? S->own_hi - S->own_lo : n;
    if (on) LAUNCH(c, k_count_owner, std::min(nblk(on, 256), 2368u), 256, 0, c->owner_level + o0, on, cnt.p);
    unsigned long long h[64];
    read_back(c, h, cnt.p, sizeof(h));
    for (size_t l = 0; l < c->levels.size(); l++) c->levels[l].npart_final = (int64_t)h[l];
    cnt.release();
  }
  if (split) {
    c->stage_cnt_extra["comm_calls"] = cm->coll_calls;
    c->stage_cnt_extra["comm_bytes"] = cm->coll_bytes;
    c->stage_cnt_extra["comm_us"] = (int64_t)(cm->coll_ms * 1000.0);
  }
}

// ------------------------------------------------------------------------------------------------
// queries
// ------------------------------------------------------------------------------------------------
__global__ void k_level_export(LV v, const int32_t *__restrict__ crow, const int32_t *__restrict__ row_c0, const uint64_t *__restrict__ rowkey,
                               const int32_t *__restrict__ rowplane, const int32_t *__restrict__ plane_r0, int nrow, int nplane, const uint8_t *__restrict__ row_flags,
                               int32_t *__restrict__ x, int32_t *__restrict__ y, int32_t *__restrict__ z, uint8_t *__restrict__ rf)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  int X, Y, Z; lv_coords(v, c, X, Y, Z);
  x[c] = X; y[c] = Y; z[c] = Z;
  uint8_t f = 0;
  const int L = (int)v.L;
  if (v.dense) {
    if (X == 0) f |= 1; if (X == L - 1) f |= 2; if (Y == 0) f |= 4; if (Y == L - 1) f |= 8; if (Z == 0) f |= 16; if (Z == L - 1) f |= 32;
  } else {
    int r = crow[c], c0 = row_c0[r], c1 = row_c0[r + 1];
    uint64_t k = v.ckey[c];
    if ((c == c0) || (v.ckey[c - 1] + 1 != k) || v.xbreak[c - 1]) f |= 1;
    if ((c == c1 - 1) || (v.ckey[c + 1] != k + 1) || v.xbreak[c]) f |= 2;
    f |= row_flags[r];            // first / last row of its cquad, plane of its pquad: from the rows of the whole box (k_row_tested)
  }
  (void)nrow; (void)nplane; (void)rowkey; (void)rowplane; (void)plane_r0;
  rf[c] = f;
}

// particles per node when the level was deposited: computed on demand for the query API
__global__ void k_count_cells(const int32_t *__restrict__ pcell, uint64_t np, int32_t *__restrict__ count)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const int c = i < np ? pcell[i] : -1;
  unsigned peers = __match_any_sync(0xffffffffu, c);
  if (c >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&count[c], __popc(peers));
}
__global__ void k_fill_i32(int32_t *a, uint64_t n, int32_t v)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_scatter_cells(const uint32_t *__restrict__ plist, const int32_t *__restrict__ pcell, uint64_t np, int32_t *__restrict__ out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= np) return;
  out[plist ? plist[i] : i] = pcell[i];
}


// ------------------------------------------------------------------------------------------------
// NEXT-1: patch labelling of a level -- the colouring sweep of ahf_gridinfo (src/libahf/ahf_gridinfo.c:236-577,
// numbering :751-775, periodic flags :640-672 / testBound :1090-1118) as connected components over the face neighbours the
// reference's search sees (patches.cuh states why this equals the sequential sweep).  The cell arrays are in traversal order.
// ------------------------------------------------------------------------------------------------
// face neighbour d (0 x-1, 1 x+1, 2 y-1, 3 y+1, 4 z-1, 5 z+1) of cell c, -1 when the reference's search does not see it
__device__ __forceinline__ int face_nb(const LV &v, const int32_t *__restrict__ nbr, int c, int d, int x)
{
  if (v.dense) {                                   // periodic full grid: everything is visible
    const int M = (int)v.L - 1, lg = v.logL;
    int y = (c >> lg) & M, z = c >> (2 * lg), xx = x;
    if (d == 0) xx = (x - 1) & M; else if (d == 1) xx = (x + 1) & M;
    else if (d == 2) y = (y - 1) & M; else if (d == 3) y = (y + 1) & M;
    else if (d == 4) z = (z - 1) & M; else z = (z + 1) & M;
    return (((z << lg) | y) << lg) | xx;
  }
  if (d == 0) return nb_get(v, nbr, c, 4, 0, x);
  if (d == 1) return nb_get(v, nbr, c, 4, 2, x);
  return nb_get(v, nbr, c, d == 2 ? 3 : d == 3 ? 5 : d == 4 ? 1 : 7, 1, x);
}
__device__ __forceinline__ int cell_x(const LV &v, int c) { return (int)((v.dense ? (uint64_t)c : v.ckey[c]) & (uint64_t)(v.L - 1)); }

// Initial forest: every cell points at its x-1 neighbour when it sees it (and that one is earlier), so the runs along x -- the bulk of
// all edges -- are chains before the first union; path halving shortens them on the first find.
// owned != nullptr (ONE box on several ranks): only the rank's own cells are edge sources -- their neighbour entries are exact, those of
// the ghost shell need not be; a ghost cell appears as the target of an own cell's edge only
__global__ void k_patch_init(LV v, const int32_t *__restrict__ nbr, int32_t *__restrict__ parent, const uint8_t *__restrict__ owned)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  if (owned && !owned[c]) { parent[c] = c; return; }
  const int p = face_nb(v, nbr, c, 0, cell_x(v, c));
  parent[c] = (p >= 0 && p < c) ? p : c;
}
// edges (c, n): n visible from c and earlier in traversal order (a later neighbour is still uncoloured when the sweep visits c).
// The x-1 edges are in the initial forest.  An edge (c, n) to the y-1 / z-1 neighbour is implied, and skipped, when the three edges
// (c, p), (p, m), (n, m) exist with p = x-1 neighbour of c and m = the same-direction neighbour of p = x-1 neighbour of n: only the first cell
// of every overlap of two runs does a union (the 5.7 ms of unions per 256^3 hierarchy were finds over already joined sets).
__global__ void k_patch_link(LV v, const int32_t *__restrict__ nbr, int32_t *parent, const uint8_t *__restrict__ owned)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  if (owned && !owned[c]) return;
  const int x = cell_x(v, c);
  const int p = face_nb(v, nbr, c, 0, x);
  const bool has_p = p >= 0 && p < c;
  const int xp = has_p ? cell_x(v, p) : 0;
#pragma unroll
  for (int d = 2; d <= 4; d += 2) {
    const int n = face_nb(v, nbr, c, d, x);
    if (n < 0 || n >= c) continue;
    if (has_p) {
      const int m = face_nb(v, nbr, p, d, xp);
      if (m >= 0 && m < p && m < n && (!owned || (owned[p] && owned[n])) && face_nb(v, nbr, n, 0, cell_x(v, n)) == m) continue;
    }
    uf_unite(parent, c, n);
  }
#pragma unroll
  for (int d = 1; d <= 5; d += 2) {                       // forward neighbours are earlier cells only across a periodic face
    const int n = face_nb(v, nbr, c, d, x);
    if (n >= 0 && n < c) uf_unite(parent, c, n);
  }
}
__global__ void k_patch_roots(int32_t *parent, int n, int32_t *__restrict__ root, uint8_t *__restrict__ isroot)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int r = uf_find(parent, c);
  root[c] = r; isroot[c] = (r == c) ? 1 : 0;
}
// isolated-refinement index = rank of the component's first cell among the first cells; periodic flags as testBound sets them
__global__ void k_patch_iso(LV v, const int32_t *__restrict__ nbr, const int32_t *__restrict__ root, const int *__restrict__ rank,
                            int32_t *__restrict__ iso, uint8_t *__restrict__ periodic3)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  const int i = rank[root[c]];
  iso[c] = i;
  int x, y, z;
  lv_coords(v, c, x, y, z);
  if (x == 0 && face_nb(v, nbr, c, 0, x) >= 0) periodic3[3 * (size_t)i + 0] = 1;
  if (y == 0 && face_nb(v, nbr, c, 2, x) >= 0) periodic3[3 * (size_t)i + 1] = 1;
  if (z == 0 && face_nb(v, nbr, c, 4, x) >= 0) periodic3[3 * (size_t)i + 2] = 1;
}


// ------------------------------------------------------------------------------------------------
// NEXT-2, first half: RefCentre (src/libahf/ahf_halos.c:935-1620) as segmented reductions over the patch labels.
// Per isolated refinement: node and particle counts, sums of node positions (plain and density weighted), maximum density, sum of the
// positions of the particles the level finally owns, extents.  Sums are double atomics: the counts, the maximum and the extents are
// exact, the centres differ from the reference's sequential sums in the last bits only.
// ------------------------------------------------------------------------------------------------
constexpr int PS_ACC = 16;      // 0 nodes 1 parts | 2-4 geom 5 norm | 6-8 dens-weighted 9 norm | 10 maxDens (bits) | 11-13 particle sums 14 norm
__device__ __forceinline__ double node_coord(int x, double L, double shift) { return fmod((double)x / L + shift + 1.0, 1.0); }   // ahf_halos.c:1024-1029
__device__ __forceinline__ void atomic_max_pos(double *a, double v) { atomicMax(reinterpret_cast<unsigned long long *>(a), (unsigned long long)__double_as_longlong(v)); }
__device__ __forceinline__ void atomic_min_pos(double *a, double v) { atomicMin(reinterpret_cast<unsigned long long *>(a), (unsigned long long)__double_as_longlong(v)); }

// Integer sums (PS_IACC u64 words per refinement: 0 nodes, 1-3 sum of (2x+1) [+2L across a periodic face], 4 particles, 5-7 sum of
// trunc(x * 2^36)): node coordinates are the dyadic rationals (2x+1)/(2L) and particle coordinates are float32, so these sums are EXACT
// (particles: for coordinates >= 2^-13, below that truncated to 2^-36), independent of the order of the atomics -- the table is bit
// reproducible from run to run and equals the reference's sequential double sums wherever those are exact themselves.
constexpr int PS_IACC = 8;
constexpr double PS_PSCALE = 68719476736.0;       // 2^36: 2 * 2^36 * 2^26 particles of one refinement fit 63 bits
// Cells come in (z, y, x) order, so the cells of one refinement are runs along x: a warp first adds up its runs of equal refinement
// index (segmented scan by shuffles over the run number) and only the last lane of a run touches the accumulators.  One atomic per value and RUN instead of per cell: the first version spent 4.4 ms per level
// of the 256^3 box in same-address atomics (profiles/r2n_launches_summary.txt), 28 % of all kernel time of that capture.
template <typename T> __device__ __forceinline__ T seg_shfl_up(T v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ unsigned long long seg_shfl_up<unsigned long long>(unsigned long long v, int o)
{
  const unsigned lo = __shfl_up_sync(0xffffffffu, (unsigned)v, o), hi = __shfl_up_sync(0xffffffffu, (unsigned)(v >> 32), o);
  return ((unsigned long long)hi << 32) | lo;
}
__global__ void k_pstat_cells(LV v, const int32_t *__restrict__ iso, const uint8_t *__restrict__ per3, const float *__restrict__ dens, double *acc,
                              unsigned long long *iacc, const uint8_t *__restrict__ owned)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = c < v.ncell && (!owned || owned[c]);
  const int lane = threadIdx.x & 31;
  int i = -1 - lane;                                       // invalid lanes: keys that match nobody
  unsigned long long iv[4] = { 0, 0, 0, 0 };
  double dv[4] = { 0, 0, 0, 0 }, dmax = 0.0;
  if (valid) {
    i = iso[c];
    int x, y, z;
    lv_coords(v, c, x, y, z);
    const double L = (double)v.L, shift = 0.5 / L;
    double xx = node_coord(x, L, shift), yy = node_coord(y, L, shift), zz = node_coord(z, L, shift);
    unsigned long long ix = 2ull * (unsigned long long)x + 1ull, iy = 2ull * (unsigned long long)y + 1ull, iz = 2ull * (unsigned long long)z + 1ull;
    if (per3[3 * i + 0] && xx < 0.5) { xx += 1.0; ix += 2ull * (unsigned long long)v.L; }                        // :1032-1040
    if (per3[3 * i + 1] && yy < 0.5) { yy += 1.0; iy += 2ull * (unsigned long long)v.L; }
    if (per3[3 * i + 2] && zz < 0.5) { zz += 1.0; iz += 2ull * (unsigned long long)v.L; }
    double d = (double)dens[c] + 1.0;                                  // + simu.mean_dens, :1057
    if (d < 0.0) d = 0.0;
    iv[0] = 1ull; iv[1] = ix; iv[2] = iy; iv[3] = iz;
    dv[0] = xx * d; dv[1] = yy * d; dv[2] = zz * d; dv[3] = d; dmax = d;
  }
  // runs of equal refinement index (the same index may come back later in the warp: A B A; the RUN number is what is contiguous)
  const int iprev = __shfl_up_sync(0xffffffffu, i, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || iprev != i);
  const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ko = __shfl_up_sync(0xffffffffu, run, o);
    const bool take = lane >= o && ko == run;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const unsigned long long a = seg_shfl_up(iv[q], o);
      const double b = __shfl_up_sync(0xffffffffu, dv[q], o);
      if (take) { iv[q] += a; dv[q] += b; }
    }
    const double m = __shfl_up_sync(0xffffffffu, dmax, o);
    if (take && m > dmax) dmax = m;
  }
  const int inext = __shfl_down_sync(0xffffffffu, i, 1);
  if (valid && (lane == 31 || inext != i)) {               // last lane of its run
    double *a = acc + (size_t)PS_ACC * i;
    unsigned long long *ia = iacc + (size_t)PS_IACC * i;
    atomicAdd(ia + 0, iv[0]); atomicAdd(ia + 1, iv[1]); atomicAdd(ia + 2, iv[2]); atomicAdd(ia + 3, iv[3]);
    atomicAdd(a + 6, dv[0]); atomicAdd(a + 7, dv[1]); atomicAdd(a + 8, dv[2]); atomicAdd(a + 9, dv[3]);
    atomic_max_pos(a + 10, dmax);
  }
}
// particles the level finally owns (node.ll at ahf_halos time): centre of mass of the refinement's particles (:1120-1180)
__global__ void k_pstat_parts(const float4 *__restrict__ pos4, const uint32_t *__restrict__ plist, const int32_t *__restrict__ pcell, uint64_t np,
                              const int8_t *__restrict__ owner, int lev, const int32_t *__restrict__ iso, const uint8_t *__restrict__ per3,
                              unsigned long long *iacc, uint64_t own_lo, uint64_t own_hi)
{
  // the particle list is in Hilbert order: the particles of one refinement come in runs.  Runs of equal refinement index inside a warp are
  // added up by shuffles (k_pstat_cells) and their last lane issues the four atomics -- one per run instead of one per particle (2.5 ms per
  // 256^3 hierarchy were same-address atomics of the clump cores).  Integer sums: the table does not depend on the grouping.
  const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int i = -1 - lane;                                       // lanes without a particle of this level: keys that match nobody
  unsigned long long iv[4] = { 0, 0, 0, 0 };
  if (k < np) {
    const uint64_t p = plist ? plist[k] : k;
    const int cc = pcell[k];
    if (owner[p] == lev && cc >= 0 && p >= own_lo && p < own_hi) {          // [own_lo, own_hi): the rank's own particles of a split box
      i = iso[cc];
      const float4 q = pos4[p];
      double xp = (double)q.x, yp = (double)q.y, zp = (double)q.z;
      if (per3[3 * i + 0] && xp < 0.5) xp += 1.0;
      if (per3[3 * i + 1] && yp < 0.5) yp += 1.0;
      if (per3[3 * i + 2] && zp < 0.5) zp += 1.0;
      iv[0] = 1ull; iv[1] = (unsigned long long)(xp * PS_PSCALE); iv[2] = (unsigned long long)(yp * PS_PSCALE); iv[3] = (unsigned long long)(zp * PS_PSCALE);
    }
  }
  const int iprev = __shfl_up_sync(0xffffffffu, i, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || iprev != i);
  const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ko = __shfl_up_sync(0xffffffffu, run, o);
    const bool take = lane >= o && ko == run;
#pragma unroll
    for (int q = 0; q < 4; q++) { const unsigned long long a = seg_shfl_up(iv[q], o); if (take) iv[q] += a; }
  }
  const int inext = __shfl_down_sync(0xffffffffu, i, 1);
  if (i >= 0 && (lane == 31 || inext != i)) {
    unsigned long long *ia = iacc + (size_t)PS_IACC * i;
    atomicAdd(ia + 4, iv[0]); atomicAdd(ia + 5, iv[1]); atomicAdd(ia + 6, iv[2]); atomicAdd(ia + 7, iv[3]);
  }
}
__device__ __forceinline__ double f1mod1(double v) { return v >= 2.0 ? v - 2.0 : v >= 1.0 ? v - 1.0 : v; }     // specific.c:120-129
// normalisation and fall-backs (:1240-1370), boundRefDiv of the periodic refinements (:1400-1470).
// out[i][18]: 0 numNodes 1 numParts 2-4 centre (= particle centre, AHFcomcentre) 5 maxDens 6-8 centreGEOM 9-11 centreDens 12-17 extents
__global__ void k_pstat_finish(const double *__restrict__ acc, const unsigned long long *__restrict__ iacc, const uint8_t *__restrict__ per3, int niso, double L,
                               double *__restrict__ out, double *__restrict__ div3)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= niso) return;
  const double *a = acc + (size_t)PS_ACC * i;
  const unsigned long long *ia = iacc + (size_t)PS_IACC * i;
  double *o = out + (size_t)18 * i;
  const double nnode = (double)ia[0], npart = (double)ia[4];
  o[0] = nnode; o[1] = npart; o[5] = a[10];
  for (int q = 0; q < 3; q++) o[6 + q] = nnode > 0 ? f1mod1(((double)ia[1 + q] / (2.0 * L)) / nnode + 1.0) : 0.0;
  for (int q = 0; q < 3; q++) o[9 + q] = a[9] > 0 ? f1mod1(a[6 + q] / a[9] + 1.0) : o[6 + q];
  for (int q = 0; q < 3; q++) o[2 + q] = npart > 0 ? f1mod1(((double)ia[5 + q] / PS_PSCALE) / npart + 1.0) : o[6 + q];
  if (a[10] <= 5e-16) for (int q = 0; q < 3; q++) o[9 + q] = o[6 + q];
  const double bl = 1.0 / L, vol = nnode * (bl * bl * bl);
  const double rad = pow((3.0 * vol) / (4 * 3.14159265358979323846), 0.333333333) * 1.1;
  for (int q = 0; q < 3; q++) {
    double dv = -1.0;
    if (per3[3 * i + q]) { const double aa = rad + o[9 + q], bb = 1.0 - rad + o[9 + q]; dv = fmod((aa + bb) / 2.0, 1.0); }
    div3[3 * i + q] = dv;
    o[12 + 2 * q] = 100000.0; o[13 + 2 * q] = 0.0;       // min sentinel as in the reference; max: node coordinates are > 0, so 0 = untouched
  }
}
// extents (:1480-1595): MinMax, or MinMaxBound at boundRefDiv for the periodic refinements (specific.c:204-254)
__global__ void k_pstat_extents(LV v, const int32_t *__restrict__ iso, const double *__restrict__ div3, double *out, const uint8_t *__restrict__ owned)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = c < v.ncell && (!owned || owned[c]);
  const int lane = threadIdx.x & 31;
  int i = -1 - lane;
  // per axis: candidate for the minimum and for the maximum of the refinement (+inf / -1: none), as MinMax / MinMaxBound would see this node
  double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1.0, -1.0, -1.0 };
  if (valid) {
    i = iso[c];
    int xyz[3];
    lv_coords(v, c, xyz[0], xyz[1], xyz[2]);
    const double L = (double)v.L, shift = 0.5 / L;
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const double xx = node_coord(xyz[q], L, shift), dv = div3[3 * i + q];
      if (dv < 0.0) { lo[q] = xx; hi[q] = xx; }
      else if (xx < dv) hi[q] = xx;
      else lo[q] = xx;
    }
  }
  const int iprev = __shfl_up_sync(0xffffffffu, i, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || iprev != i);
  const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {                         // runs of equal refinement index: segmented min / max (see k_pstat_cells)
    const int ko = __shfl_up_sync(0xffffffffu, run, o);
    const bool take = lane >= o && ko == run;
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const double a = __shfl_up_sync(0xffffffffu, lo[q], o), b = __shfl_up_sync(0xffffffffu, hi[q], o);
      if (take) { if (a < lo[q]) lo[q] = a; if (b > hi[q]) hi[q] = b; }
    }
  }
  const int inext = __shfl_down_sync(0xffffffffu, i, 1);
  if (valid && (lane == 31 || inext != i)) {
    double *o = out + (size_t)18 * i;
#pragma unroll
    for (int q = 0; q < 3; q++) {
      if (lo[q] < 1e299) atomic_min_pos(o + 12 + 2 * q, lo[q]);
      if (hi[q] >= 0.0) atomic_max_pos(o + 13 + 2 * q, hi[q]);
    }
  }
}
__global__ void k_pstat_fix(int niso, double *out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= niso) return;
  double *o = out + (size_t)18 * i;
  for (int q = 0; q < 3; q++) { if (o[12 + 2 * q] == 100000.0) o[12 + 2 * q] = 0.0; if (o[13 + 2 * q] == 0.0) o[13 + 2 * q] = 1.0; }    // :1597-1612
}


// ------------------------------------------------------------------------------------------------
// The same table for ONE box split over several ranks (slab.cu).  A refinement may cross rank boundaries, so the labelling is a
// connected-component problem over all ranks:
//   1. every rank labels its resident cells with edges whose SOURCE is one of its own cells (exact neighbour entries); ghost cells are
//      edge targets only.  Every edge of the box-wide graph has exactly one own source, so the union of the ranks' local components is
//      the box-wide partition once components that share a cell are joined;
//   2. a local component is named by the global key of its first cell.  A ghost cell that an own cell is joined to is an own cell of
//      another rank: (its key, its local name) goes to everybody, the owner answers with the pair (sender's name, owner's name);
//   3. the pairs of all ranks are a small graph over names: every rank solves it (host union-find), the name of a refinement becomes
//      the smallest key of its class = the key of its first cell in the box-wide traversal order;
//   4. refinements are numbered in the order of their first cells (what the reference's sweep produces): every rank contributes the
//      first cells it owns, the sorted union is the numbering;
//   5. sums / maxima / extents over OWN cells and OWN particles, combined over the ranks (integer sums exactly, double sums in rank
//      order), so every rank ends with the same table.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t cell_gkey(const LV &v, int c) { return v.dense ? (uint64_t)c : v.ckey[c]; }

__global__ void k_ps_haschild(const int32_t *__restrict__ root, int n, uint8_t *__restrict__ haschild)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n && root[c] != c) haschild[root[c]] = 1;
}
__global__ void k_ps_touched(const int32_t *__restrict__ root, const uint8_t *__restrict__ haschild, const uint8_t *__restrict__ owned, int n, uint8_t *__restrict__ t)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) t[c] = (!owned[c] && (root[c] != c || haschild[c])) ? 1 : 0;
}
__global__ void k_ps_emit(LV v, const int32_t *__restrict__ root, const uint8_t *__restrict__ t, const int *__restrict__ pos, uint64_t *__restrict__ rkey, uint64_t *__restrict__ rlab)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell || !t[c]) return;
  rkey[pos[c]] = cell_gkey(v, c); rlab[pos[c]] = cell_gkey(v, root[c]);
}
// records of all ranks -> pairs (sender's name, my name) for the cells I own.  bad: a record for a cell I own but do not have
__global__ void k_ps_pairs(LV v, const int32_t *__restrict__ root, const uint8_t *__restrict__ own3, int bd, int rank, const uint64_t *__restrict__ rkey,
                           const uint64_t *__restrict__ rlab, int nrec, uint8_t *__restrict__ keep, uint64_t *__restrict__ pa, uint64_t *__restrict__ pb, int *__restrict__ bad)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrec) return;
  const uint64_t k = rkey[i];
  const int x = (int)(k & (uint64_t)(v.L - 1)), y = (int)((k >> v.logL) & (uint64_t)(v.L - 1)), z = (int)(k >> (2 * v.logL));
  const int sh = v.logL - bd;
  keep[i] = 0;
  if (own3[((((size_t)(z >> sh)) << bd) | (size_t)(y >> sh)) << bd | (size_t)(x >> sh)] != (uint8_t)rank) return;
  const int c = v.ncell > 0 ? lv_lookup(v, x, y, z) : -1;
  if (c < 0) { atomicAdd(bad, 1); return; }
  const uint64_t mine = cell_gkey(v, root[c]);
  if (mine != rlab[i]) { keep[i] = 1; pa[i] = rlab[i]; pb[i] = mine; }
}
__global__ void k_ps_compact2(const uint8_t *__restrict__ keep, const int *__restrict__ pos, int n, const uint64_t *__restrict__ a, const uint64_t *__restrict__ b,
                              uint64_t *__restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) { out[2 * (size_t)pos[i]] = a[i]; out[2 * (size_t)pos[i] + 1] = b[i]; }
}
__device__ __forceinline__ int lower_bound_u64(const uint64_t *__restrict__ a, int n, uint64_t k)
{
  int lo = 0, hi = n;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (a[m] < k) lo = m + 1; else hi = m; }
  return lo;
}
// box-wide name of every resident cell's component; first[c]: c is an own cell and the first cell of its refinement
__global__ void k_ps_canon(LV v, const int32_t *__restrict__ root, const uint8_t *__restrict__ owned, const uint64_t *__restrict__ mfrom, const uint64_t *__restrict__ mto,
                           int nm, uint64_t *__restrict__ canon, uint8_t *__restrict__ first)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  uint64_t lab = cell_gkey(v, root[c]);
  const int j = lower_bound_u64(mfrom, nm, lab);
  if (j < nm && mfrom[j] == lab) lab = mto[j];
  canon[c] = lab;
  first[c] = (owned[c] && lab == cell_gkey(v, c)) ? 1 : 0;
}
__global__ void k_ps_firstkeys(LV v, const uint8_t *__restrict__ first, const int *__restrict__ pos, uint64_t *__restrict__ out)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < v.ncell && first[c]) out[pos[c]] = cell_gkey(v, c);
}
__global__ void k_ps_iso(LV v, const uint8_t *__restrict__ owned, const uint64_t *__restrict__ canon, const uint64_t *__restrict__ gfirst, int niso, const int32_t *__restrict__ nbr,
                         int32_t *__restrict__ iso, uint32_t *__restrict__ per3u, int *__restrict__ bad)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= v.ncell) return;
  if (!owned[c]) { iso[c] = -1; return; }
  const int i = lower_bound_u64(gfirst, niso, canon[c]);
  if (i >= niso || gfirst[i] != canon[c]) { iso[c] = -1; atomicAdd(bad, 1); return; }
  iso[c] = i;
  int x, y, z;
  lv_coords(v, c, x, y, z);
  if (x == 0 && face_nb(v, nbr, c, 0, x) >= 0) per3u[3 * (size_t)i + 0] = 1u;
  if (y == 0 && face_nb(v, nbr, c, 2, x) >= 0) per3u[3 * (size_t)i + 1] = 1u;
  if (z == 0 && face_nb(v, nbr, c, 4, x) >= 0) per3u[3 * (size_t)i + 2] = 1u;
}
__global__ void k_ps_per3(const uint32_t *__restrict__ u, int n, uint8_t *__restrict__ per3)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) per3[i] = u[i] ? 1 : 0;
}
// accumulators of all ranks ([R][niso * PS_ACC] doubles, [R][niso * PS_IACC] u64) -> box-wide: integer sums exactly, double sums in rank
// order, the maximum density as a maximum
__global__ void k_ps_combine_acc(const double *__restrict__ acc_all, const unsigned long long *__restrict__ iacc_all, int R, int niso, double *__restrict__ acc,
                                 unsigned long long *__restrict__ iacc)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < niso * PS_ACC) {
    const int q = t % PS_ACC;
    double v = 0.0;
    for (int p = 0; p < R; p++) { const double a = acc_all[(size_t)p * niso * PS_ACC + t]; if (q == 10) { if (a > v) v = a; } else v += a; }
    acc[t] = v;
  }
  if (t < niso * PS_IACC) {
    unsigned long long v = 0;
    for (int p = 0; p < R; p++) v += iacc_all[(size_t)p * niso * PS_IACC + t];
    iacc[t] = v;
  }
}
// extents of all ranks ([R][niso][18], the sentinels of k_pstat_finish where a rank holds none of the refinement's cells) -> box-wide
__global__ void k_ps_combine_ext(const double *__restrict__ out_all, int R, int niso, double *__restrict__ out)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= niso * 6) return;
  const int i = t / 6, q = 12 + t % 6;
  double v = out_all[(size_t)i * 18 + q];
  for (int p = 1; p < R; p++) { const double a = out_all[((size_t)p * niso + i) * 18 + q]; if (q & 1) { if (a > v) v = a; } else { if (a < v) v = a; } }
  out[(size_t)i * 18 + q] = v;
}

// every rank's `n_mine` elements (device) -> all elements in rank order (device, returned; caller frees), counts[p] per rank
template <typename T> static T *gather_var(ahfgpu_ctx *c, const T *send, int64_t n_mine, std::vector<int64_t> &counts, int64_t &total)
{
  Comm *cm = c->comm;
  const int R = cm->nranks;
  counts.assign(R, 0);
  long long mine = n_mine;
  std::vector<long long> all(R);
  cm->allgather_host(c, &mine, all.data(), sizeof(long long));
  std::vector<size_t> bytes(R), off(R + 1, 0);
  for (int p = 0; p < R; p++) { counts[p] = all[p]; bytes[p] = (size_t)all[p] * sizeof(T); off[p + 1] = off[p] + bytes[p]; }
  total = (int64_t)(off[R] / sizeof(T));
  T *out = dalloc<T>(total > 0 ? total : 1);
  cm->allgatherv(c, send, out, bytes.data(), off.data());
  return out;
}

static void patch_stats_split(ahfgpu_ctx *c, int lev, std::vector<double> &table)
{
  Comm *cm = c->comm; Slab *S = c->slab;
  const int R = cm->nranks;
  Level empty;
  Level &l = lev < (int)c->levels.size() ? c->levels[lev] : empty;
  if (lev >= (int)c->levels.size()) {                      // the rank's cells end on a coarser level: it only takes part in the exchanges
    empty.L = (int64_t)c->par.lgrid_dom << lev; empty.ncell = 0; empty.dense = false;
  }
  if (l.ncell > 0 && !l.dense && !l.nbr) AHF_FAIL("level has no neighbour table");
  LV v = view(l);
  const int nc = (int)l.ncell;
  DevBuf<int32_t> parent, root, diso; DevBuf<uint8_t> owned, isroot, hasch, touched, first, keep, per; DevBuf<int> pos, bad; DevBuf<uint64_t> rkey, rlab, canon, mfrom, mto;
  DevBuf<uint32_t> per3u; DevBuf<double> acc, accl, div3; DevBuf<unsigned long long> iacc, iaccl;
  parent.reserve(nc); root.reserve(nc); diso.reserve(nc); owned.reserve(nc); isroot.reserve(nc); hasch.reserve(nc); touched.reserve(nc); first.reserve(nc); pos.reserve(nc);
  canon.reserve(nc); bad.reserve(2);
  CUDA_CHECK(cudaMemsetAsync(bad.p, 0, 2 * sizeof(int), c->stream));
  int nt = 0;
  if (nc > 0) {
    LAUNCH(c, k_owned_cells, nblk(nc, 256), 256, 0, v, S->own3, S->bd, cm->rank, owned.p);
    LAUNCH(c, k_patch_init, nblk(nc, 256), 256, 0, v, l.nbr, parent.p, (const uint8_t *)owned.p);
    LAUNCH(c, k_patch_link, nblk(nc, 256), 256, 0, v, l.nbr, parent.p, (const uint8_t *)owned.p);
    LAUNCH(c, k_patch_roots, nblk(nc, 256), 256, 0, parent.p, nc, root.p, isroot.p);
    CUDA_CHECK(cudaMemsetAsync(hasch.p, 0, nc, c->stream));
    LAUNCH(c, k_ps_haschild, nblk(nc, 256), 256, 0, root.p, nc, hasch.p);
    LAUNCH(c, k_ps_touched, nblk(nc, 256), 256, 0, root.p, hasch.p, owned.p, nc, touched.p);
    nt = exclusive_scan<uint8_t>(c, touched.p, pos.p, (uint64_t)nc);
    rkey.reserve(nt); rlab.reserve(nt);
    if (nt) LAUNCH(c, k_ps_emit, nblk(nc, 256), 256, 0, v, root.p, touched.p, pos.p, rkey.p, rlab.p);
  } else { rkey.reserve(1); rlab.reserve(1); }
  // ---- 2. records to everybody, pairs from the owners
  std::vector<int64_t> cnt; int64_t nrec = 0, nrec2 = 0;
  uint64_t *akey = gather_var<uint64_t>(c, rkey.p, nt, cnt, nrec);
  uint64_t *alab = gather_var<uint64_t>(c, rlab.p, nt, cnt, nrec2);
  std::vector<uint64_t> mypairs;
  if (nrec > 0) {
    DevBuf<uint64_t> pa, pb, pc; DevBuf<int> ppos;
    keep.reserve(nrec); pa.reserve(nrec); pb.reserve(nrec); ppos.reserve(nrec);
    LAUNCH(c, k_ps_pairs, nblk(nrec, 256), 256, 0, v, root.p, S->own3, S->bd, cm->rank, akey, alab, (int)nrec, keep.p, pa.p, pb.p, bad.p);
    const int np2 = exclusive_scan<uint8_t>(c, keep.p, ppos.p, (uint64_t)nrec);
    if (np2) {
      pc.reserve((size_t)2 * np2);
      LAUNCH(c, k_ps_compact2, nblk(nrec, 256), 256, 0, keep.p, ppos.p, (int)nrec, pa.p, pb.p, pc.p);
      mypairs.resize((size_t)2 * np2);
      CUDA_CHECK(cudaMemcpyAsync(mypairs.data(), pc.p, sizeof(uint64_t) * 2 * np2, cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    pa.release(); pb.release(); pc.release(); ppos.release();
  }
  ahf::dfree(akey); ahf::dfree(alab);
  {                                                      // a joined ghost cell its owner does not have would be a broken ghost shell
    int hb[2] = { 0, 0 };
    read_back(c, hb, bad.p, sizeof(hb));
    if (hb[0]) AHF_FAIL("patch labels of a split box: a rank joined a ghost cell its owner does not hold (ghost shell too thin?)");
  }
  {                                                      // many ghost cells join the same two components: one pair each
    std::vector<std::pair<uint64_t, uint64_t>> pr(mypairs.size() / 2);
    for (size_t i = 0; i < pr.size(); i++) { uint64_t a = mypairs[2 * i], b = mypairs[2 * i + 1]; if (a > b) std::swap(a, b); pr[i] = { a, b }; }
    std::sort(pr.begin(), pr.end());
    pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
    mypairs.resize(pr.size() * 2);
    for (size_t i = 0; i < pr.size(); i++) { mypairs[2 * i] = pr[i].first; mypairs[2 * i + 1] = pr[i].second; }
  }
  // ---- 3. the name graph, on every rank
  std::vector<uint64_t> from, to;
  {
    DevBuf<uint64_t> dp;
    dp.reserve(mypairs.size() ? mypairs.size() : 1);
    if (!mypairs.empty()) CUDA_CHECK(cudaMemcpyAsync(dp.p, mypairs.data(), sizeof(uint64_t) * mypairs.size(), cudaMemcpyHostToDevice, c->stream));
    int64_t tot = 0;
    uint64_t *allp = gather_var<uint64_t>(c, dp.p, (int64_t)mypairs.size(), cnt, tot);
    std::vector<uint64_t> ap((size_t)tot);
    if (tot) { CUDA_CHECK(cudaMemcpyAsync(ap.data(), allp, sizeof(uint64_t) * tot, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); }
    ahf::dfree(allp); dp.release();
    std::vector<uint64_t> keys(ap);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    std::vector<int32_t> uf(keys.size());
    for (size_t i = 0; i < uf.size(); i++) uf[i] = (int32_t)i;
    for (size_t i = 0; i + 1 < ap.size(); i += 2) {
      const int32_t a = (int32_t)(std::lower_bound(keys.begin(), keys.end(), ap[i]) - keys.begin());
      const int32_t b = (int32_t)(std::lower_bound(keys.begin(), keys.end(), ap[i + 1]) - keys.begin());
      uf_unite(uf.data(), a, b);                           // smaller index = smaller key becomes the root (patches.cuh)
    }
    from = keys; to.resize(keys.size());
    for (size_t i = 0; i < keys.size(); i++) to[i] = keys[(size_t)uf_find(uf.data(), (int32_t)i)];
  }
  const int nm = (int)from.size();
  mfrom.reserve(nm ? nm : 1); mto.reserve(nm ? nm : 1);
  if (nm) {
    CUDA_CHECK(cudaMemcpyAsync(mfrom.p, from.data(), sizeof(uint64_t) * nm, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(mto.p, to.data(), sizeof(uint64_t) * nm, cudaMemcpyHostToDevice, c->stream));
  }
  // ---- 4. numbering by first cells
  int nfirst = 0;
  DevBuf<uint64_t> fk;
  if (nc > 0) {
    LAUNCH(c, k_ps_canon, nblk(nc, 256), 256, 0, v, root.p, owned.p, mfrom.p, mto.p, nm, canon.p, first.p);
    nfirst = exclusive_scan<uint8_t>(c, first.p, pos.p, (uint64_t)nc);
    fk.reserve(nfirst ? nfirst : 1);
    if (nfirst) LAUNCH(c, k_ps_firstkeys, nblk(nc, 256), 256, 0, v, first.p, pos.p, fk.p);
  } else { fk.reserve(1); CUDA_CHECK(cudaStreamSynchronize(c->stream)); }
  int64_t niso64 = 0;
  uint64_t *gfirst = gather_var<uint64_t>(c, fk.p, nfirst, cnt, niso64);
  const int ni = (int)niso64;
  {                                                      // the ranks' lists are disjoint and sorted: one host sort of the union
    std::vector<uint64_t> g((size_t)ni);
    if (ni) { CUDA_CHECK(cudaMemcpyAsync(g.data(), gfirst, sizeof(uint64_t) * ni, cudaMemcpyDeviceToHost, c->stream)); CUDA_CHECK(cudaStreamSynchronize(c->stream)); }
    std::sort(g.begin(), g.end());
    if (std::adjacent_find(g.begin(), g.end()) != g.end()) AHF_FAIL("patch labels of a split box: two ranks claim the first cell of one refinement");
    if (ni) CUDA_CHECK(cudaMemcpyAsync(gfirst, g.data(), sizeof(uint64_t) * ni, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));        // g leaves scope
  }
  table.assign((size_t)18 * ni, 0.0);
  if (ni > 0) {
    // ---- 5. labels, periodic flags, sums; combined over the ranks
    per3u.reserve((size_t)3 * ni); per.reserve((size_t)3 * ni);
    CUDA_CHECK(cudaMemsetAsync(per3u.p, 0, sizeof(uint32_t) * 3 * ni, c->stream));
    if (nc > 0) LAUNCH(c, k_ps_iso, nblk(nc, 256), 256, 0, v, owned.p, canon.p, gfirst, ni, l.nbr, diso.p, per3u.p, bad.p + 1);
    cm->allreduce_sum_u32(c, per3u.p, (size_t)3 * ni);
    LAUNCH(c, k_ps_per3, nblk((size_t)3 * ni, 256), 256, 0, per3u.p, 3 * ni, per.p);
    accl.reserve((size_t)PS_ACC * ni); iaccl.reserve((size_t)PS_IACC * ni); acc.reserve((size_t)PS_ACC * ni); iacc.reserve((size_t)PS_IACC * ni); div3.reserve((size_t)3 * ni);
    CUDA_CHECK(cudaMemsetAsync(accl.p, 0, sizeof(double) * PS_ACC * ni, c->stream));
    CUDA_CHECK(cudaMemsetAsync(iaccl.p, 0, sizeof(unsigned long long) * PS_IACC * ni, c->stream));
    if (nc > 0) {
      LAUNCH(c, k_pstat_cells, nblk(nc, 256), 256, 0, v, diso.p, per.p, l.dens, accl.p, iaccl.p, (const uint8_t *)owned.p);
      if (l.npart_dep > 0)
        LAUNCH(c, k_pstat_parts, nblk(l.npart_dep, 256), 256, 0, c->pos4, l.plist, l.pcell, (uint64_t)l.npart_dep, c->owner_level, (int)lev, diso.p, per.p, iaccl.p,
               (uint64_t)S->own_lo, (uint64_t)S->own_hi);
    }
    {
      int hb[2] = { 0, 0 };
      read_back(c, hb, bad.p, sizeof(hb));
      if (hb[1]) AHF_FAIL("patch labels of a split box: an own cell belongs to no numbered refinement");
    }
    std::vector<size_t> bytes(R), off(R + 1, 0);
    for (int p = 0; p < R; p++) { bytes[p] = sizeof(double) * PS_ACC * (size_t)ni; off[p + 1] = off[p] + bytes[p]; }
    double *acc_all = dalloc<double>((size_t)R * PS_ACC * ni);
    cm->allgatherv(c, accl.p, acc_all, bytes.data(), off.data());
    for (int p = 0; p < R; p++) { bytes[p] = sizeof(unsigned long long) * PS_IACC * (size_t)ni; off[p + 1] = off[p] + bytes[p]; }
    unsigned long long *iacc_all = dalloc<unsigned long long>((size_t)R * PS_IACC * ni);
    cm->allgatherv(c, iaccl.p, iacc_all, bytes.data(), off.data());
    LAUNCH(c, k_ps_combine_acc, nblk((size_t)ni * PS_ACC, 256), 256, 0, acc_all, iacc_all, R, ni, acc.p, iacc.p);
    double *out = dalloc<double>((size_t)18 * ni);
    LAUNCH(c, k_pstat_finish, nblk(ni, 128), 128, 0, acc.p, iacc.p, per.p, ni, (double)l.L, out, div3.p);
    if (nc > 0) LAUNCH(c, k_pstat_extents, nblk(nc, 256), 256, 0, v, diso.p, div3.p, out, (const uint8_t *)owned.p);
    for (int p = 0; p < R; p++) { bytes[p] = sizeof(double) * 18 * (size_t)ni; off[p + 1] = off[p] + bytes[p]; }
    double *out_all = dalloc<double>((size_t)R * 18 * ni);
    cm->allgatherv(c, out, out_all, bytes.data(), off.data());
    LAUNCH(c, k_ps_combine_ext, nblk((size_t)ni * 6, 256), 256, 0, out_all, R, ni, out);
    LAUNCH(c, k_pstat_fix, nblk(ni, 128), 128, 0, ni, out);
    CUDA_CHECK(cudaMemcpyAsync(table.data(), out, sizeof(double) * 18 * (size_t)ni, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    ahf::dfree(acc_all); ahf::dfree(iacc_all); ahf::dfree(out_all); ahf::dfree(out);
  }
  ahf::dfree(gfirst);
  parent.release(); root.release(); diso.release(); owned.release(); isroot.release(); hasch.release(); touched.release(); first.release(); keep.release(); per.release();
  pos.release(); bad.release(); rkey.release(); rlab.release(); canon.release(); mfrom.release(); mto.release(); per3u.release(); acc.release(); accl.release(); div3.release();
  iacc.release(); iaccl.release(); fk.release();
}

}  // namespace ahf

using namespace ahf;

extern "C" int ahfgpu_amr_patch_stats(ahfgpu_ctx *c, int32_t lev, int64_t *niso, double *stats, int64_t stats_cap)
{
  try {
    if (!c || lev < 0 || lev >= std::max((int)c->levels.size(), c->g_nlevels)) AHF_FAIL("bad level");
    if (!c->owner_level) AHF_FAIL("no hierarchy");
    CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
    if (c->slab && c->comm) {                       // ONE box on several ranks: a collective call, every rank ends with the same table
      auto it = c->pstat_split.find(lev);
      if (it == c->pstat_split.end()) {
        std::vector<double> t;
        patch_stats_split(c, lev, t);
        it = c->pstat_split.emplace(lev, std::move(t)).first;
      }
      const int64_t ni = (int64_t)(it->second.size() / 18);
      if (stats && ni > stats_cap) AHF_FAIL("stats buffer too small");
      if (stats && ni > 0) memcpy(stats, it->second.data(), sizeof(double) * 18 * (size_t)ni);
      if (niso) *niso = ni;
      return 0;
    }
    Level &l = c->levels[lev];
    if (!l.dense && !l.nbr) AHF_FAIL("level has no neighbour table");
    const int nc = (int)l.ncell;
    int ni = 0;
    if (nc > 0 && l.pstat_n >= 0) {                  // the table of this hierarchy is still there (a count query came first)
      ni = (int)l.pstat_n;
      if (stats && (int64_t)ni > stats_cap) AHF_FAIL("stats buffer too small");
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      if (stats && ni > 0) CUDA_CHECK(cudaMemcpy(stats, l.pstat, sizeof(double) * 18 * (size_t)ni, cudaMemcpyDeviceToHost));
    } else if (nc > 0) {
      LV v = view(l);
      DevBuf<int32_t> parent, root, diso; DevBuf<uint8_t> isroot, per; DevBuf<int> rank; DevBuf<double> acc, div3; DevBuf<unsigned long long> iacc;
      parent.reserve(nc); root.reserve(nc); diso.reserve(nc); isroot.reserve(nc); rank.reserve(nc); per.reserve((size_t)3 * nc);
      LAUNCH(c, k_patch_init, nblk(nc, 256), 256, 0, v, l.nbr, parent.p, (const uint8_t *)nullptr);
      LAUNCH(c, k_patch_link, nblk(nc, 256), 256, 0, v, l.nbr, parent.p, (const uint8_t *)nullptr);
      LAUNCH(c, k_patch_roots, nblk(nc, 256), 256, 0, parent.p, nc, root.p, isroot.p);
      ni = exclusive_scan<uint8_t>(c, isroot.p, rank.p, (uint64_t)nc);
      CUDA_CHECK(cudaMemsetAsync(per.p, 0, (size_t)3 * nc, c->stream));
      LAUNCH(c, k_patch_iso, nblk(nc, 256), 256, 0, v, l.nbr, root.p, rank.p, diso.p, per.p);
      if (stats && (int64_t)ni > stats_cap) AHF_FAIL("stats buffer too small");
      acc.reserve((size_t)PS_ACC * ni); div3.reserve((size_t)3 * ni);
      double *out = dalloc<double>((size_t)18 * ni);
      iacc.reserve((size_t)PS_IACC * ni);
      CUDA_CHECK(cudaMemsetAsync(acc.p, 0, sizeof(double) * PS_ACC * ni, c->stream));
      CUDA_CHECK(cudaMemsetAsync(iacc.p, 0, sizeof(unsigned long long) * PS_IACC * ni, c->stream));
      LAUNCH(c, k_pstat_cells, nblk(nc, 256), 256, 0, v, diso.p, per.p, l.dens, acc.p, iacc.p, (const uint8_t *)nullptr);
      if (l.npart_dep > 0)
        LAUNCH(c, k_pstat_parts, nblk(l.npart_dep, 256), 256, 0, c->pos4, l.plist, l.pcell, (uint64_t)l.npart_dep, c->owner_level, (int)lev, diso.p, per.p, iacc.p, (uint64_t)0, ~(uint64_t)0);
      LAUNCH(c, k_pstat_finish, nblk(ni, 128), 128, 0, acc.p, iacc.p, per.p, ni, (double)l.L, out, div3.p);
      LAUNCH(c, k_pstat_extents, nblk(nc, 256), 256, 0, v, diso.p, div3.p, out, (const uint8_t *)nullptr);
      LAUNCH(c, k_pstat_fix, nblk(ni, 128), 128, 0, ni, out);
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      if (stats && ni > 0) CUDA_CHECK(cudaMemcpy(stats, out, sizeof(double) * 18 * (size_t)ni, cudaMemcpyDeviceToHost));
      ahf::dfree(l.pstat); l.pstat = out; l.pstat_n = ni;
      parent.release(); root.release(); diso.release(); isroot.release(); rank.release(); per.release(); acc.release(); div3.release(); iacc.release();
    }
    if (niso) *niso = ni;
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

namespace ahf {
}  // namespace ahf

using namespace ahf;

extern "C" int ahfgpu_amr_patches(ahfgpu_ctx *c, int32_t lev, int32_t *iso, int64_t *niso, uint8_t *periodic3)
{
  try {
    if (!c || lev < 0 || lev >= (int)c->levels.size()) AHF_FAIL("bad level");
    CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
    Level &l = c->levels[lev];
    if (!l.dense && !l.nbr) AHF_FAIL("level has no neighbour table");
    const int nc = (int)l.ncell;
    int ni = 0;
    if (nc > 0) {
      LV v = view(l);
      DevBuf<int32_t> parent, root, diso; DevBuf<uint8_t> isroot, per; DevBuf<int> rank;
      parent.reserve(nc); root.reserve(nc); diso.reserve(nc); isroot.reserve(nc); rank.reserve(nc); per.reserve((size_t)3 * nc);
      LAUNCH(c, k_patch_init, nblk(nc, 256), 256, 0, v, l.nbr, parent.p, (const uint8_t *)nullptr);
      LAUNCH(c, k_patch_link, nblk(nc, 256), 256, 0, v, l.nbr, parent.p, (const uint8_t *)nullptr);
      LAUNCH(c, k_patch_roots, nblk(nc, 256), 256, 0, parent.p, nc, root.p, isroot.p);
      ni = exclusive_scan<uint8_t>(c, isroot.p, rank.p, (uint64_t)nc);
      CUDA_CHECK(cudaMemsetAsync(per.p, 0, (size_t)3 * nc, c->stream));
      LAUNCH(c, k_patch_iso, nblk(nc, 256), 256, 0, v, l.nbr, root.p, rank.p, diso.p, per.p);
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      if (iso) CUDA_CHECK(cudaMemcpy(iso, diso.p, sizeof(int32_t) * (size_t)nc, cudaMemcpyDeviceToHost));
      if (periodic3 && ni > 0) CUDA_CHECK(cudaMemcpy(periodic3, per.p, (size_t)3 * ni, cudaMemcpyDeviceToHost));
      parent.release(); root.release(); diso.release(); isroot.release(); rank.release(); per.release();
    }
    if (niso) *niso = ni;
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

extern "C" int ahfgpu_amr_level_owned(ahfgpu_ctx *c, int32_t lev, uint8_t *owned)
{
  try {
    if (!c || lev < 0 || lev >= (int)c->levels.size()) AHF_FAIL("bad level");
    if (!owned) AHF_FAIL("null argument");
    CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
    Level &l = c->levels[lev];
    const size_t nc = (size_t)l.ncell;
    if (!c->slab || !c->comm) { memset(owned, 1, nc); return 0; }
    DevBuf<uint8_t> o;
    o.reserve(nc);
    LAUNCH(c, k_owned_cells, nblk(nc, 256), 256, 0, view(l), c->slab->own3, c->slab->bd, c->comm->rank, o.p);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(owned, o.p, nc, cudaMemcpyDeviceToHost));
    o.release();
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

extern "C" int ahfgpu_amr_level_get(ahfgpu_ctx *c, int32_t lev, int32_t *x, int32_t *y, int32_t *z, float *dens, uint8_t *runflags,
                                    uint8_t *interior, uint8_t *mark, int32_t *count)
{
  try {
    if (!c || lev < 0 || lev >= (int)c->levels.size()) AHF_FAIL("bad level");
    CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
    Level &l = c->levels[lev];
    const size_t nc = (size_t)l.ncell;
    if (x || y || z || runflags) {
      DevBuf<int32_t> dx, dy, dz; DevBuf<uint8_t> rf;
      dx.reserve(nc); dy.reserve(nc); dz.reserve(nc); rf.reserve(nc);
      LAUNCH(c, k_level_export, nblk(nc, 256), 256, 0, view(l), l.crow, l.row_c0, l.rowkey, l.dense ? nullptr : l.rowplane, l.plane_r0,
             (int)l.nrow, (int)l.nplane, l.row_flags, dx.p, dy.p, dz.p, rf.p);
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      if (x) CUDA_CHECK(cudaMemcpy(x, dx.p, nc * 4, cudaMemcpyDeviceToHost));
      if (y) CUDA_CHECK(cudaMemcpy(y, dy.p, nc * 4, cudaMemcpyDeviceToHost));
      if (z) CUDA_CHECK(cudaMemcpy(z, dz.p, nc * 4, cudaMemcpyDeviceToHost));
      if (runflags) CUDA_CHECK(cudaMemcpy(runflags, rf.p, nc, cudaMemcpyDeviceToHost));
      dx.release(); dy.release(); dz.release(); rf.release();
    }
    if (dens) CUDA_CHECK(cudaMemcpy(dens, l.dens, nc * 4, cudaMemcpyDeviceToHost));
    if (interior) {
      if (l.dense) memset(interior, 1, nc);
      else CUDA_CHECK(cudaMemcpy(interior, l.interior, nc, cudaMemcpyDeviceToHost));
    }
    if (mark) CUDA_CHECK(cudaMemcpy(mark, l.mark, nc, cudaMemcpyDeviceToHost));
    if (count) {
      CUDA_CHECK(cudaMemsetAsync(l.count, 0, nc * 4, c->stream));
      if (l.npart_dep) LAUNCH(c, k_count_cells, nblk(l.npart_dep, 256), 256, 0, l.pcell, (uint64_t)l.npart_dep, l.count);
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      CUDA_CHECK(cudaMemcpy(count, l.count, nc * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

extern "C" int ahfgpu_amr_particle_levels(ahfgpu_ctx *c, int8_t *owner_level, int32_t *cell_of, int32_t nlev_cap)
{
  try {
    if (!c || !c->owner_level) AHF_FAIL("no hierarchy");
    CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
    const uint64_t n = c->n;
    if (owner_level) CUDA_CHECK(cudaMemcpy(owner_level, c->owner_level, n, cudaMemcpyDeviceToHost));
    if (cell_of) {
      DevBuf<int32_t> tmp;
      tmp.reserve(n);
      for (int l = 0; l < (int)c->levels.size() && l < nlev_cap; l++) {
        Level &lv = c->levels[l];
        LAUNCH(c, k_fill_i32, nblk(n, 256), 256, 0, tmp.p, n, -1);
        if (lv.npart_dep) LAUNCH(c, k_scatter_cells, nblk(lv.npart_dep, 256), 256, 0, lv.plist, lv.pcell, (uint64_t)lv.npart_dep, tmp.p);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaMemcpy(cell_of + (size_t)l * n, tmp.p, n * 4, cudaMemcpyDeviceToHost));
      }
      tmp.release();
    }
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

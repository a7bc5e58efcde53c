// slab.cu -- ONE box on several GPUs: SFC slab decomposition with a ghost shell.
//
// The model is the reference's MPI mode, re-formulated for GPUs:
//   * loadbalance_update / local_equalpart (src/libutility/loadbalance.c:206,383,480): histogram of the particles over the Hilbert
//     cells of a decomposition level (LevelDomainDecomp; "blocks" here), summed over the ranks, cut into equal-particle key ranges;
//   * comm_dist_part (src/comm.c:104-316): every particle goes to the owner of its key range;
//   * comm_dist_part_ahf + sfc_boundary_2_get (src/comm.c:324ff, src/libsfc/sfc_boundary.c:100-144): the particles of the boundary
//     shell are DUPLICATED on the neighbouring ranks.
// Here all three happen in one pass over the particles a rank has read: a particle is sent to its owner AND to every rank that owns
// a block within T blocks (Chebyshev, periodic) of its own block -- ONE personalised exchange (NCCL send/recv over NVLink), followed
// by the ONE sort the path needs anyway.  With the ghost shell wider than the reach of the hierarchy construction (7 domain cells,
// DESIGN.md section 6) and than the largest gathering radius, a rank builds every level over its own cells and runs the halo pass for
// the haloes centred in its range WITHOUT any per-level ghost-cell traffic; the only per-level collectives left are the row lists
// (the reference's run structure is non-local along x) and three scalars (mesh.cu).
// The partition is STABLE (particles keep their input order per destination, pieces arrive in rank order), so that the stable sort
// breaks equal keys by global input index -- exactly the order of the single-GPU sort of the whole file.
#include "comm.cuh"
#include "hilbert.cuh"
#include "scan.cuh"

namespace ahf {

void sfc_keys_f3(ahfgpu_ctx *c, const float *pos3, uint64_t n, uint64_t *keys);
void sfc_sort_device4_gid(ahfgpu_ctx *c, const void *pos4_dev, const void *mom4_dev, const uint32_t *gid, uint64_t n, bool has_w, bool has_u);

namespace {

template <typename T> T *dalloc(size_t n) { return static_cast<T *>(cache_alloc((n ? n : 1) * sizeof(T))); }
inline unsigned nblk(uint64_t n, int b) { return (unsigned)((n + b - 1) / b); }

// particles per block (warp-aggregated: the particles of a clump core share one block)
__global__ void k_block_hist(const uint64_t *__restrict__ keys, uint64_t n, int sh, uint32_t *__restrict__ hist)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const bool     valid = i < n;
  const uint32_t b = valid ? (uint32_t)(keys[i] >> sh) : 0xffffffffu;
  const unsigned peers = __match_any_sync(0xffffffffu, b);
  if (valid && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[b], (uint32_t)__popc(peers));
}

// equal-particle cuts of the block sequence (local_equalpart): split[r] = first block whose exclusive prefix reaches r * N / R
__global__ void k_splitters(const int *__restrict__ prefix, uint32_t nb3, unsigned long long ntot, int R, unsigned long long *__restrict__ split)
{
  const int r = threadIdx.x;
  if (r > R) return;
  if (r == 0) { split[0] = 0; return; }
  if (r == R) { split[R] = nb3; return; }
  const unsigned long long target = (ntot * (unsigned long long)r + (unsigned long long)R - 1) / (unsigned long long)R;
  uint32_t lo = 0, hi = nb3;
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((unsigned long long)prefix[mid] < target) lo = mid + 1; else hi = mid; }
  split[r] = lo;
}
__device__ __forceinline__ int owner_of(const unsigned long long *__restrict__ split, int R, unsigned long long h)
{
  int r = 0;
  while (r + 1 < R && split[r + 1] <= h) r++;
  return r;
}
// owner rank of every block on the (z, y, x) grid of blocks
__global__ void k_own3(const unsigned long long *__restrict__ split, int R, int bd, uint32_t nb3, uint8_t *__restrict__ own3)
{
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nb3) return;
  uint32_t x, y, z;
  hilbert_coords((uint64_t)h, (unsigned)bd, x, y, z);
  own3[(((size_t)z << bd) | y) << bd | x] = (uint8_t)owner_of(split, R, h);
}
// ghost masks: bit r of gm[b] = rank r owns a block within T blocks of b (periodic); three separable passes
__global__ void k_gm_init(const uint8_t *__restrict__ own3, uint32_t nb3, uint32_t *__restrict__ gm)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb3) gm[b] = 1u << own3[b];
}
__global__ void k_gm_dilate(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int bd, int axis, int T)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t B = 1u << bd, nb3 = 1u << (3 * bd);
  if (b >= nb3) return;
  const int      sh = axis * bd;
  const uint32_t q = (b >> sh) & (B - 1u), rest = b & ~((B - 1u) << sh);
  uint32_t m = 0;
  for (int d = -T; d <= T; d++) m |= in[rest | (((q + (uint32_t)(d + (int)B)) & (B - 1u)) << sh)];
  out[b] = m;
}

// block (z, y, x) index of a position: the coordinates the Hilbert key is made of (hilbert_key_pos), top bd bits
__device__ __forceinline__ uint32_t block_of(float x, float y, float z, int bd)
{
  const float    mx = 2097152.0f;
  const uint32_t top = 1u << 21;
  uint32_t c0 = (uint32_t)(int32_t)(x * mx), c1 = (uint32_t)(int32_t)(y * mx), c2 = (uint32_t)(int32_t)(z * mx);
  if (c0 >= top) c0 = top - 1;
  if (c1 >= top) c1 = top - 1;
  if (c2 >= top) c2 = top - 1;
  const int s = 21 - bd;
  return (((c2 >> s) << bd) | (c1 >> s)) << bd | (c0 >> s);
}

// ---- stable multi-destination partition: tile = PT_THREADS x PT_ITEMS consecutive particles, thread t owns items t*PT_ITEMS ...
constexpr int PT_THREADS = 256, PT_ITEMS = 4, PT_TILE = PT_THREADS * PT_ITEMS, PT_MAXR = 32;

__global__ void __launch_bounds__(PT_THREADS) k_part_count(const float *__restrict__ pos3, uint64_t n, int bd, const uint32_t *__restrict__ gm, int R,
                                                           uint32_t ntile, int *__restrict__ cnt /* [R][ntile] */)
{
  __shared__ int s_cnt[PT_MAXR];
  if (threadIdx.x < PT_MAXR) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * PT_TILE + (uint64_t)threadIdx.x * PT_ITEMS;
  int loc[PT_MAXR];
#pragma unroll
  for (int d = 0; d < PT_MAXR; d++) loc[d] = 0;
  for (int i = 0; i < PT_ITEMS; i++) {
    const uint64_t j = base + i;
    if (j >= n) break;
    const uint32_t m = gm[block_of(pos3[3 * j], pos3[3 * j + 1], pos3[3 * j + 2], bd)];
#pragma unroll
    for (int d = 0; d < PT_MAXR; d++) loc[d] += (m >> d) & 1u;
  }
#pragma unroll
  for (int d = 0; d < PT_MAXR; d++) {
    if (d >= R) break;
    int v = loc[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[d], v);
  }
  __syncthreads();
  if ((int)threadIdx.x < R) cnt[(size_t)threadIdx.x * ntile + blockIdx.x] = s_cnt[threadIdx.x];
}

__global__ void k_part_offsets(const int *__restrict__ scan, const int *__restrict__ total, uint32_t ntile, int R, long long *__restrict__ off /* [R+1] */)
{
  const int d = threadIdx.x;
  if (d < R) off[d] = scan[(size_t)d * ntile];
  if (d == R) off[R] = *total;
}

__global__ void __launch_bounds__(PT_THREADS) k_part_fill(const float *__restrict__ pos3, const float *__restrict__ mom3, const float *__restrict__ w,
                                                          const float *__restrict__ u, uint64_t n, uint32_t id_base, int bd, const uint32_t *__restrict__ gm,
                                                          int R, uint32_t ntile, const int *__restrict__ scan /* [R][ntile] exclusive */,
                                                          float4 *__restrict__ spos, float4 *__restrict__ smom, uint32_t *__restrict__ sgid)
{
  __shared__ int s_w[PT_MAXR][PT_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * PT_TILE + (uint64_t)threadIdx.x * PT_ITEMS;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  uint32_t msk[PT_ITEMS];
  for (int i = 0; i < PT_ITEMS; i++) {
    const uint64_t j = base + i;
    msk[i] = j < n ? gm[block_of(pos3[3 * j], pos3[3 * j + 1], pos3[3 * j + 2], bd)] : 0u;
  }
  // per destination: exclusive prefix of this thread's count over the threads of the tile (thread order = particle order)
  int pre[PT_MAXR];
#pragma unroll
  for (int d = 0; d < PT_MAXR; d++) {
    pre[d] = 0;
    if (d >= R) continue;
    int v = 0;
    for (int i = 0; i < PT_ITEMS; i++) v += (msk[i] >> d) & 1u;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) s_w[d][wrp] = inc;
    pre[d] = inc - v;
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < PT_MAXR; d++) {
    if (d >= R) continue;
    int wb = 0;
    for (int q = 0; q < wrp; q++) wb += s_w[d][q];
    pre[d] += wb + scan[(size_t)d * ntile + blockIdx.x];
  }
  for (int i = 0; i < PT_ITEMS; i++) {
    const uint64_t j = base + i;
    if (j >= n || !msk[i]) continue;
    const float4 p = make_float4(pos3[3 * j], pos3[3 * j + 1], pos3[3 * j + 2], w ? w[j] : 1.0f);
    const float4 m = make_float4(mom3[3 * j], mom3[3 * j + 1], mom3[3 * j + 2], u ? u[j] : -1.0f);
    const uint32_t g = id_base + (uint32_t)j;
#pragma unroll
    for (int d = 0; d < PT_MAXR; d++) {
      if (d >= R) break;
      if ((msk[i] >> d) & 1u) { const int o = pre[d]++; spos[o] = p; smom[o] = m; sgid[o] = g; }
    }
  }
}

__global__ void k_lower_bounds(const uint64_t *__restrict__ keys, uint64_t n, unsigned long long klo, unsigned long long khi, unsigned long long *__restrict__ out)
{
  if (threadIdx.x > 1) return;
  const unsigned long long k = threadIdx.x == 0 ? klo : khi;
  uint64_t lo = 0, hi = n;
  while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < k) lo = mid + 1; else hi = mid; }
  out[threadIdx.x] = lo;
}

}  // namespace

void slab_free(ahfgpu_ctx *c)
{
  if (!c->slab) return;
  dfree(c->slab->own3);
  delete c->slab;
  c->slab = nullptr;
}

void slab_distribute(ahfgpu_ctx *c, uint64_t id_base, double ghost_width, int decomp_bits)
{
  Comm *cm = c->comm;
  if (!cm) AHF_FAIL("no communicator: call ahfgpu_comm_init_nccl / ahfgpu_comm_init_local first");
  if (!c->in_pos) AHF_FAIL("no uploaded particles: call ahfgpu_upload_soa first");
  const int R = cm->nranks;
  if (R > PT_MAXR) AHF_FAIL("more than 32 ranks are not supported");
  const uint64_t n = c->in_n;
  int logL = 0; while ((1 << logL) < c->par.lgrid_dom) logL++;
  // decomposition level: blocks of 4 domain cells unless told otherwise (LevelDomainDecomp), at most 2^8 per dimension
  int bd = decomp_bits > 0 ? decomp_bits : logL - 2;
  if (bd > 8) bd = 8;
  if (bd > logL) bd = logL;
  if (bd < 1) bd = 1;
  if (!(ghost_width > 0.0)) ghost_width = 8.0 / (double)c->par.lgrid_dom;
  if (ghost_width < 8.0 / (double)c->par.lgrid_dom) ghost_width = 8.0 / (double)c->par.lgrid_dom;    // reach of the hierarchy construction: 7 cells
  const int B = 1 << bd;
  int T = (int)std::ceil(ghost_width * (double)B - 1e-9);
  if (T < 1) T = 1;
  if (2 * T + 1 > B) T = (B - 1) / 2;                 // the shell wraps the whole box: every rank sees everything
  const uint32_t nb3 = 1u << (3 * bd);
  const int      sh = 3 * (21 - bd);
  slab_free(c);
  c->free_halos(); c->free_levels(); c->free_particles();
  cm->coll_ms = 0.0; cm->coll_calls = 0; cm->coll_bytes = 0;
  Slab *S = new Slab();
  c->slab = S;
  S->bd = bd; S->T = T; S->ghost_width = ghost_width; S->split.assign(R + 1, 0);
  // ---- keys and block histogram of what this rank read
  uint64_t *keys = dalloc<uint64_t>(n);
  uint32_t *hist = dalloc<uint32_t>(nb3);
  int      *pre = dalloc<int>(nb3), *tot = dalloc<int>(1);
  unsigned long long *d_split = dalloc<unsigned long long>(R + 3);
  {
    Stage st(c, "slab_keys", (int64_t)n);
    sfc_keys_f3(c, c->in_pos, n, keys);
    CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * nb3, c->stream));
    if (n) LAUNCH(c, k_block_hist, nblk(n, 256), 256, 0, keys, n, sh, hist);
  }
  {
    Stage st(c, "slab_histogram_allreduce", (int64_t)nb3 * 4);
    cm->allreduce_sum_u32(c, hist, nb3);
  }
  uint32_t *gm = dalloc<uint32_t>(nb3), *gm2 = dalloc<uint32_t>(nb3);
  S->own3 = dalloc<uint8_t>(nb3);
  {
    Stage st(c, "slab_decompose", (int64_t)nb3);
    DevBuf<int> bs;
    CUDA_CHECK(cudaMemsetAsync(tot, 0, sizeof(int), c->stream));
    exclusive_scan_async<int>(c, reinterpret_cast<const int *>(hist), pre, nb3, tot, bs);
    int h_tot = 0;
    read_back(c, &h_tot, tot, sizeof(int));
    if (h_tot < 0) AHF_FAIL("more than 2^31 particles in one box are not supported by the decomposition");
    S->n_total = (uint64_t)h_tot;
    LAUNCH(c, k_splitters, 1, 64, 0, pre, nb3, (unsigned long long)S->n_total, R, d_split);
    LAUNCH(c, k_own3, nblk(nb3, 256), 256, 0, d_split, R, bd, nb3, S->own3);
    LAUNCH(c, k_gm_init, nblk(nb3, 256), 256, 0, S->own3, nb3, gm);
    LAUNCH(c, k_gm_dilate, nblk(nb3, 256), 256, 0, gm, gm2, bd, 0, T);
    LAUNCH(c, k_gm_dilate, nblk(nb3, 256), 256, 0, gm2, gm, bd, 1, T);
    LAUNCH(c, k_gm_dilate, nblk(nb3, 256), 256, 0, gm, gm2, bd, 2, T);
    std::vector<unsigned long long> hs(R + 1);
    read_back(c, hs.data(), d_split, sizeof(unsigned long long) * (R + 1));
    for (int r = 0; r <= R; r++) S->split[r] = hs[r];
    bs.release();
  }
  // ---- stable partition by destination (owner + ghost holders)
  const uint32_t ntile = (uint32_t)((n + PT_TILE - 1) / PT_TILE);
  int       *cnt = dalloc<int>((size_t)R * (ntile ? ntile : 1));
  long long *d_off = dalloc<long long>(R + 1);
  std::vector<long long> off(R + 1, 0);
  float4 *spos = nullptr, *smom = nullptr; uint32_t *sgid = nullptr;
  {
    Stage st(c, "slab_partition", (int64_t)n);
    if (n) {
      DevBuf<int> bs;
      LAUNCH(c, k_part_count, ntile, PT_THREADS, 0, c->in_pos, n, bd, gm2, R, ntile, cnt);
      CUDA_CHECK(cudaMemsetAsync(tot, 0, sizeof(int), c->stream));
      exclusive_scan_async<int>(c, cnt, cnt, (uint64_t)R * ntile, tot, bs);
      LAUNCH(c, k_part_offsets, 1, 64, 0, cnt, tot, ntile, R, d_off);
      read_back(c, off.data(), d_off, sizeof(long long) * (R + 1));
      bs.release();
    }
    const size_t nsend = (size_t)off[R];
    if (nsend >= (1ull << 31)) AHF_FAIL("more than 2^31 particles to send from one rank");
    spos = dalloc<float4>(nsend); smom = dalloc<float4>(nsend); sgid = dalloc<uint32_t>(nsend);
    if (id_base + n > (1ull << 32)) AHF_FAIL("global particle index above 2^32");
    if (n) LAUNCH(c, k_part_fill, ntile, PT_THREADS, 0, c->in_pos, c->in_mom, c->in_w, c->in_u, n, (uint32_t)id_base, bd, gm2, R, ntile, cnt, spos, smom, sgid);
  }
  // ---- exchange
  std::vector<long long> mycnt(R), allcnt((size_t)R * R);
  for (int d = 0; d < R; d++) mycnt[d] = off[d + 1] - off[d];
  cm->allgather_host(c, mycnt.data(), allcnt.data(), sizeof(long long) * R);
  std::vector<size_t> sb(R), rb(R), roff(R + 1, 0);
  for (int p = 0; p < R; p++) { sb[p] = (size_t)mycnt[p]; rb[p] = (size_t)allcnt[(size_t)p * R + cm->rank]; roff[p + 1] = roff[p] + rb[p]; }
  const size_t n_new = roff[R];
  if (n_new >= (1ull << 32)) AHF_FAIL("more than 2^32-1 particles on one rank");
  float4 *rpos = dalloc<float4>(n_new), *rmom = dalloc<float4>(n_new); uint32_t *rgid = dalloc<uint32_t>(n_new);
  {
    Stage st(c, "slab_exchange", (int64_t)n_new * 36);
    std::vector<const void *> sp(R); std::vector<void *> rp(R); std::vector<size_t> sbytes(R), rbytes(R);
    for (int pass = 0; pass < 3; pass++) {
      const size_t esz = pass < 2 ? sizeof(float4) : sizeof(uint32_t);
      const char  *sbase = pass == 0 ? (const char *)spos : pass == 1 ? (const char *)smom : (const char *)sgid;
      char        *rbase = pass == 0 ? (char *)rpos : pass == 1 ? (char *)rmom : (char *)rgid;
      for (int p = 0; p < R; p++) { sp[p] = sbase + (size_t)off[p] * esz; sbytes[p] = sb[p] * esz; rp[p] = rbase + roff[p] * esz; rbytes[p] = rb[p] * esz; }
      cm->alltoallv(c, sp.data(), sbytes.data(), rp.data(), rbytes.data());
    }
  }
  dfree(spos); dfree(smom); dfree(sgid); dfree(cnt); dfree(d_off); dfree(keys); dfree(hist); dfree(pre); dfree(gm); dfree(gm2);
  // ---- the one sort: keys of the received particles, stable radix sort, payload gather; order[] = global input index
  sfc_sort_device4_gid(c, rpos, rmom, rgid, n_new, c->in_w != nullptr, c->in_u != nullptr);
  dfree(rpos); dfree(rmom); dfree(rgid);
  {
    const unsigned long long klo = (unsigned long long)S->split[cm->rank] << sh, khi = (unsigned long long)S->split[cm->rank + 1] << sh;
    LAUNCH(c, k_lower_bounds, 1, 32, 0, c->keys, (uint64_t)c->n, klo, khi, d_split);
    unsigned long long h2[2];
    read_back(c, h2, d_split, sizeof(h2));
    S->own_lo = h2[0]; S->own_hi = h2[1];
  }
  dfree(d_split); dfree(tot);
  c->n_total = S->n_total;
  c->stage_cnt_extra["slab_owned"] = (int64_t)(S->own_hi - S->own_lo);
  c->stage_cnt_extra["slab_resident"] = (int64_t)c->n;
}

}  // namespace ahf

// common.cuh -- context, error handling, stage timers shared by the libahfgpu translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <chrono>
#include "../../include/ahfgpu.h"

namespace ahf {

extern thread_local std::string g_last_error;

struct Error { std::string msg; };

inline void fail(const char *file, int line, const std::string &what)
{
  char buf[64];
  snprintf(buf, sizeof(buf), " (%s:%d)", file, line);
  throw Error{what + buf};
}
#define AHF_FAIL(msg) ::ahf::fail(__FILE__, __LINE__, (msg))
#define CUDA_CHECK(expr)                                                                                 \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess) ::ahf::fail(__FILE__, __LINE__, std::string(#expr ": ") + cudaGetErrorString(e__)); \
  } while (0)

// every kernel launch of the library goes through this macro so that launches can be counted (bench.py gpu_launches)
// AHFGPU_KTIME=1 (diagnostic, off by default): an event pair around EVERY launch; ahfgpu_ctx::ktime_dump prints the per-kernel sums of a
// call and the idle time between consecutive kernels to stderr (warm caches, unlike an ncu launch list)
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                                      \
  do {                                                                                                   \
    if ((ctx)->ktime) (ctx)->ktime_mark(#kernel, true);                                                  \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                                     \
    if ((ctx)->ktime) (ctx)->ktime_mark(nullptr, false);                                                 \
    (ctx)->n_launches++;                                                                                 \
    CUDA_CHECK(cudaGetLastError());                                                                      \
  } while (0)

// stream-ordered allocation from the device's default memory pool (release threshold raised in ahfgpu_init so that
// freed blocks are reused instead of being returned to the driver): every API entry sets g_pool_stream = ctx->stream
extern thread_local cudaStream_t g_pool_stream;
// Size-class cache in front of cudaMallocAsync (api.cu).  A pass of the path allocates the same few hundred blocks every time
// (level arrays, sort and halo scratch, several GB at 512^3); handing a freed block of the same size class straight back costs
// nothing, whereas the driver pool may have to re-map physical memory when its free space is fragmented (measured: hundreds of
// ms per pass at 512^3).  Blocks are cached per stream, so reuse is ordered like any other work on that stream.
void *cache_alloc(size_t bytes);
void  cache_free(void *p);
void  cache_release_all();      // give everything cached back to the driver pool
inline void dfree(void *p) { if (p) cache_free(p); }

template <typename T> struct DevBuf {
  T     *p   = nullptr;
  size_t cap = 0;
  void reserve(size_t n)
  {
    if (n <= cap) return;
    if (p) ahf::dfree(p);
    p = nullptr; cap = 0;
    p = static_cast<T *>(ahf::cache_alloc((n ? n : 1) * sizeof(T)));
    cap = n ? n : 1;
  }
  void release() { if (p) ahf::dfree(p); p = nullptr; cap = 0; }
  void adopt(T *q) { release(); p = q; cap = q ? 1 : 0; }      // take over a block of the cache allocator (released like the others)
};

// one refinement level on the device (see DESIGN.md "data layout")
struct Level {
  int64_t  L = 0;           // l1dim
  int64_t  ncell = 0;
  bool     dense = false;   // domain level: cell index == (z*L+y)*L+x, no key list / hash
  uint64_t *ckey = nullptr;     // [ncell] sorted linear keys (z*L+y)*L+x          (sparse levels)
  uint8_t  *xbreak = nullptr;   // [ncell] reference nquad run ends after this cell (sparse levels)
  float    *dens = nullptr;     // [ncell]
  uint8_t  *interior = nullptr; // [ncell] (sparse levels; dense => all interior)
  uint8_t  *tn = nullptr;       // [ncell] test_node()
  uint8_t  *mark = nullptr;     // [ncell] 0 / 1 refined / 2 ghost
  int32_t  *nbr = nullptr;      // [10][ncell] (term-major) neighbour table: 9 row centres + visibility bits of their x-1 / x+1 cells (mesh.cu nb_get)
  int32_t  *crow = nullptr;     // [ncell] row index                             (sparse levels)
  int32_t  *count = nullptr;    // [ncell] particles linked when deposited
  uint64_t *hkey = nullptr; int32_t *hval = nullptr; uint64_t hmask = 0;   // open addressing hash: block key | (first cell, occupancy mask)
  // parent/child links between consecutive levels: cells of the next level are found through them, not through the hash
  int32_t  *parent = nullptr;   // [ncell] cell of the coarser level this cell is a child of      (sparse levels)
  int32_t  *cidx = nullptr;     // [ncell] -1: no children; else slot in cbase | 0x40000000 when the children are a ghost pair
  int4     *cbase = nullptr;    // [marked cells] index on the next level of child (i=0, j, k): .x (0,0) .y (j=1,k=0) .z (0,1) .w (1,1)
  int32_t  *cpar = nullptr;     // [marked cells] the marked cell of this level behind slot s of cbase
  // rows / planes of sparse levels
  int64_t  nrow = 0, nplane = 0;
  uint64_t *rowkey = nullptr;   // [nrow] z*L+y
  int32_t  *row_c0 = nullptr;   // [nrow+1]
  uint8_t  *row_tested = nullptr;
  uint8_t  *row_flags = nullptr;  // [nrow] bit 2/3 first/last row of its y-run (cquad), bit 4/5 first/last plane of its z-run (pquad)
  int32_t  *plane_r0 = nullptr; // [nplane+1]
  int32_t  *rowplane = nullptr; // [nrow]
  double   critdens = 0, masstopartdens = 0;
  int64_t  npart_dep = 0, npart_final = 0;
  // the same counts over the WHOLE box when the box is split over several contexts (slab.cu; equal to the local ones otherwise): kernel
  // and fixed-point scale of the deposit are chosen from these, so that every rank rounds every term exactly as one GPU would
  int64_t  g_ncell = 0, g_npart_dep = 0;
  // particles that reached this level (ascending sorted offsets) and their cell on this level
  uint32_t *plist = nullptr;    // [npart_dep]   (nullptr on the domain level = all particles)
  int32_t  *pcell = nullptr;    // [npart_dep]
  float4   *lpos = nullptr;     // [npart_dep] positions of the level's particles, contiguous (refinement levels)
  // deposit tiles of the level's particle list, found right behind the relink that made the list (one host read-back for both counts)
  uint32_t *tlist = nullptr; int32_t *tstart = nullptr; int ntile = -1;
  double   *pstat = nullptr;    // [pstat_n][18] RefCentre table of the level (ahfgpu_amr_patch_stats), kept until the hierarchy is rebuilt
  int64_t   pstat_n = -1;
  void free_all();
};

struct Comm;     // comm.cuh
struct Slab;

struct StageRec { std::string name; cudaEvent_t a, b; int64_t count; };

// A/B and debug switches of the mesh pass, read from the environment ONCE per ahfgpu_build_amr call (the level loop used to call
// getenv ~15 times per level; tests flip the switches between calls, so they are not cached for the life of the process)
struct MeshEnv {
  bool generic_deposit = false, deposit_v1 = false, dom_persist = false, sparse_v1 = false, sparse_v2 = false, testnode_v1 = false, nbr_v1 = false,
       nbr_v2 = false, debug_nbr = false, debug_relink = false, level_stages = false, stages = true;
  bool relink_v1 = false;           // AHFGPU_RELINK_V1: scan + compaction kernel instead of the fused k_compact_fused, tiles found by the deposit (A/B timing)
  bool seg_v1 = false;              // AHFGPU_SEG_V1: heads kernel + scan + fill kernel instead of the fused k_seg_heads (A/B timing)
  bool dom_v2 = false, dom2_heavy = false, dom2_stats = false;     // AHFGPU_DOM_V2=1: k_deposit_dom2 (measured slower, see mesh.cu) / force its heavy form / count heavy tiles
  int  dom_variant = 0, dom_rmax = 1, dom_s = 32;
  void read()
  {
    auto on = [](const char *n) { return getenv(n) != nullptr; };
    generic_deposit = on("AHFGPU_GENERIC_DEPOSIT"); deposit_v1 = on("AHFGPU_DEPOSIT_V1"); dom_persist = on("AHFGPU_DOM_PERSIST");
    sparse_v1 = on("AHFGPU_SPARSE_V1"); sparse_v2 = on("AHFGPU_SPARSE_V2"); testnode_v1 = on("AHFGPU_TESTNODE_V1"); nbr_v1 = on("AHFGPU_NBR_V1");
    nbr_v2 = on("AHFGPU_NBR_V2"); debug_nbr = on("AHFGPU_DEBUG_NBR"); debug_relink = on("AHFGPU_DEBUG_RELINK"); level_stages = on("AHFGPU_LEVEL_STAGES");
    // AHFGPU_STAGES=0: only the timers a caller cannot do without (amr_total, deposit_dom_kernel); every event record is a marker
    // between kernels on the stream, and ~90 of them per pass cost ~0.3 ms at 256^3
    seg_v1 = on("AHFGPU_SEG_V1"); relink_v1 = on("AHFGPU_RELINK_V1");
    const char *es = getenv("AHFGPU_STAGES"); stages = !(es && es[0] == '0');
    const char *e2 = getenv("AHFGPU_DOM_V2"); dom_v2 = e2 && e2[0] == '1'; dom2_heavy = on("AHFGPU_DOM2_HEAVY"); dom2_stats = on("AHFGPU_DOM2_STATS");
    const char *e3 = getenv("AHFGPU_DOM_S"); dom_s = (e3 && atoi(e3) == 28) ? 28 : 32;
    const char *ev = getenv("AHFGPU_DOM_VARIANT"); dom_variant = ev ? atoi(ev) : 0;
    const char *er = getenv("AHFGPU_DOM_RMAX"); dom_rmax = er ? atoi(er) : 1;
  }
};

}  // namespace ahf

struct ahfgpu_ctx {
  ahfgpu_params par{};
  int           dev = 0;
  cudaStream_t  stream = nullptr;
  int64_t       n_launches = 0;
  // resident sorted particles
  uint64_t  n = 0;
  float4   *pos4 = nullptr;     // x,y,z,weight
  float4   *mom4 = nullptr;     // px,py,pz,u
  uint64_t *keys = nullptr;
  uint32_t *order = nullptr;    // input position of sorted particle i
  bool      has_weight = false, has_u = false;
  bool      adopted = false;          // pos4/mom4/keys belong to the caller (ahfgpu_adopt_sorted)
  uint64_t  n_total = 0;              // particles of the whole box when the box is split over several contexts (0: n)
  ahf::Comm *comm = nullptr;          // several contexts working on ONE box (comm.cuh); owned by the context
  ahf::Slab *slab = nullptr;          // decomposition of the resident set (slab.cu): owned range + ghost shell
  int        g_nlevels = 0;           // levels of the whole box (a rank whose cells end earlier holds fewer)
  std::map<int, std::vector<double>> pstat_split;   // split box: per-refinement tables of the levels (host, identical on every rank) until the hierarchy is rebuilt
  // unsorted device copy kept by ahfgpu_upload_soa
  float    *in_pos = nullptr, *in_mom = nullptr, *in_w = nullptr, *in_u = nullptr;
  uint64_t  in_n = 0;
  cudaEvent_t ev[16] = {};
  unsigned long long *scan_state = nullptr; size_t scan_cap = 0;      // look-back state of the single-pass scan (scan.cuh), kept zeroed
  // second stream of ahfgpu_sfc_sort_soa_async: host->device copies and the momentum gather run here, so the momenta travel
  // while the main stream sorts and builds the hierarchy.  mom_pending: mom4 is complete only after ev_mom.
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t  ev_copy[8] = {}, ev_main = nullptr, ev_mom = nullptr;
  bool         mom_pending = false;
  // ahfgpu_particle_ids_async: its own stream, so that the permutation goes device -> host WHILE the momenta come in (the two directions of
  // the bus) instead of behind them -- queued behind them it travelled during the deep levels of the hierarchy build, whose many small
  // read-backs share the device -> host direction with it (mesh 6.47 -> 6.16 ms in the end-to-end pass without that copy)
  cudaStream_t d2h_stream = nullptr; cudaEvent_t ev_ids = nullptr; bool ids_pending = false;
  void wait_mom(bool host);     // order the main stream (host: the calling thread as well) after the momentum gather
  // hierarchy
  std::vector<ahf::Level> levels;
  int8_t   *owner_level = nullptr;   // [n]
  // halo pass results
  int64_t   nhalo = 0;
  double   *h_scal = nullptr;        // device [nhalo*NSCAL]
  int64_t  *h_moff = nullptr;        // device [nhalo+1]
  int64_t  *h_members = nullptr;     // device
  int64_t  *h_poff = nullptr;        // device [nhalo+1] (bins)
  double   *h_prof = nullptr;        // device
  double   *h_species = nullptr;     // device [nhalo*64]      (GAS_PARTICLES build: gas_only at 0, stars_only at 32)
  double   *h_prof_species = nullptr;// device [total bins*3]  (M_gas, M_star, u_gas)
  int64_t   h_total_members = 0, h_total_bins = 0;
  // ahfgpu_halo_members_buffer: the caller's pinned buffer, and whether the lists of the last halo pass already travel / sit there
  int64_t  *early_members = nullptr; int64_t early_cap = 0; bool early_sent = false; cudaEvent_t ev_members = nullptr;
  // stage timing
  std::vector<ahf::StageRec> stages;
  std::map<std::string, double>  stage_ms;
  std::map<std::string, int64_t> stage_cnt;
  std::map<std::string, int64_t> stage_cnt_extra;   // counters set directly by the stages (not event based)
  std::map<std::string, double>  stage_wall;        // host wall clock spent inside the stage scopes (ms); query "<name>@wall"
  bool stages_resolved = true;
  ahf::MeshEnv env;
  void  *h_up = nullptr; size_t h_up_bytes = 0;         // pinned staging of small host->device uploads (halo tile lists)
  void  *h_pin = nullptr; size_t h_pin_bytes = 0;       // pinned scratch of the small device->host read-backs (ahf::read_back)
  std::vector<cudaEvent_t> event_pool;        // stage-timer events are recycled, not created and destroyed every call
  // per-launch timing (AHFGPU_KTIME=1, see LAUNCH)
  struct KRec { const char *name; cudaEvent_t a, b; };
  bool ktime = false;
  std::vector<KRec> krecs;
  void ktime_mark(const char *name, bool begin);
  void ktime_dump(const char *label);

  void stage_reset();
  void stage_resolve();
  void free_particles();
  void free_levels();
  void free_halos();
};

namespace ahf {

// RAII stage timer: CUDA events on the library's stream around a group of launches
struct Stage {
  ahfgpu_ctx *c; size_t idx;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  Stage(ahfgpu_ctx *ctx, const char *name, int64_t count = 0, bool enabled = true) : c(ctx), idx((size_t)-1)
  {
    if (!enabled) return;                     // opt-in timers (per-level stages): no events, no host time
    StageRec r; r.name = name; r.count = count;
    auto take = [&](cudaEvent_t &e) { if (c->event_pool.empty()) CUDA_CHECK(cudaEventCreate(&e)); else { e = c->event_pool.back(); c->event_pool.pop_back(); } };
    take(r.a); take(r.b);
    CUDA_CHECK(cudaEventRecord(r.a, c->stream));
    c->stages.push_back(r); idx = c->stages.size() - 1; c->stages_resolved = false;
  }
  ~Stage()
  {
    if (idx == (size_t)-1) return;
    cudaEventRecord(c->stages[idx].b, c->stream);
    c->stage_wall[c->stages[idx].name] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

// small device->host read-back through pinned memory + stream sync (a copy into pageable memory goes through a driver staging
// buffer and blocks inside the call).  Up to 4 KB the words are written by a one-warp kernel straight into the pinned (device-visible)
// buffer instead of a cudaMemcpyAsync: a memcpy queues on the device->host copy engine, behind whatever bulk transfer the copy stream
// has in flight there (the end-to-end pass sends the 67 MB permutation home while the hierarchy is built: its ~30 per-level read-backs
// each waited for it, +1.3 ms per pass)
#if defined(__CUDACC__)
static __global__ void k_read_back_words(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int nwords)
{
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
}
#endif
inline void read_back(ahfgpu_ctx *c, void *host_dst, const void *dev_src, size_t bytes)
{
  if (bytes > c->h_pin_bytes) {
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr; c->h_pin_bytes = 0;
    const size_t cap = bytes > 65536 ? bytes : 65536;
    CUDA_CHECK(cudaHostAlloc(&c->h_pin, cap, cudaHostAllocDefault));
    c->h_pin_bytes = cap;
  }
#if defined(__CUDACC__)
  if (bytes <= 4096 && (bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(dev_src) & 3) == 0) {
    k_read_back_words<<<1, 64, 0, c->stream>>>(static_cast<const uint32_t *>(dev_src), static_cast<uint32_t *>(c->h_pin), (int)(bytes / 4));
    c->n_launches++;
    CUDA_CHECK(cudaGetLastError());
  } else
#endif
  CUDA_CHECK(cudaMemcpyAsync(c->h_pin, dev_src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  memcpy(host_dst, c->h_pin, bytes);
}

// entry points implemented in the .cu files
void sfc_sort_soa(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *w, const float *u, uint64_t n,
                  uint64_t *keys_out, uint32_t *order_out);
void sfc_sort_aos(ahfgpu_ctx *c, void *part, uint64_t n, uint32_t stride, int off_pos, int off_mom, int off_key,
                  int off_id, int off_w, int off_u);
void sfc_upload_soa(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *w, const float *u, uint64_t n);
void sfc_sort_resident(ahfgpu_ctx *c, uint64_t *keys_out, uint32_t *order_out);
void sfc_sort_soa_async(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *w, const float *u, uint64_t n);
void sfc_sort_device4(ahfgpu_ctx *c, const void *pos4_dev, const void *mom4_dev, uint64_t n, bool has_w, bool has_u);
void sfc_keys_only(ahfgpu_ctx *c, const float *pos3, uint64_t n, uint32_t bits, uint64_t *keys_out);
void radix_sort_pairs(ahfgpu_ctx *c, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, uint64_t n,
                      int key_bits, uint64_t **keys_sorted, uint32_t **vals_sorted, int first_bit = 0);
void amr_build(ahfgpu_ctx *c);
void mesh_device_init();
void sfc_device_init();
void halos_construct(ahfgpu_ctx *c, int64_t nhalo, const double *centre3, const double *gather_rad, const int64_t *seed);

}  // namespace ahf

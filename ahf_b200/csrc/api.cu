// api.cu -- extern "C" entry points of libahfgpu.so (see include/ahfgpu.h) and context housekeeping.
#include "common.cuh"
#include <chrono>
#include <algorithm>
#include "comm.cuh"
#include "hilbert.cuh"
#include <mutex>
#include <unordered_map>

namespace ahf {
thread_local std::string g_last_error;
thread_local cudaStream_t g_pool_stream = nullptr;

// ---- size-class block cache (see common.cuh)
namespace {
struct CacheKey { cudaStream_t st; size_t cls; bool operator<(const CacheKey &o) const { return st != o.st ? st < o.st : cls < o.cls; } };
struct CacheMeta { cudaStream_t st; size_t cls; };
std::mutex g_cache_mu;
std::map<CacheKey, std::vector<void *>> g_cache_free;
std::unordered_map<void *, CacheMeta>   g_cache_live;
// below 1 MiB: powers of two from 512 B; above: eight classes per octave (at most 12.5 % slack)
size_t cache_class(size_t bytes)
{
  if (bytes <= 512) return 512;
  size_t p2 = 512;
  while (p2 < bytes) p2 <<= 1;
  if (p2 <= (1u << 20)) return p2;
  const size_t step = p2 >> 4;                       // p2/2 .. p2 in eight steps
  return ((bytes + step - 1) / step) * step;
}
}  // namespace
void *cache_alloc(size_t bytes)
{
  const size_t cls = cache_class(bytes);
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache_free.find(CacheKey{ g_pool_stream, cls });
    if (it != g_cache_free.end() && !it->second.empty()) {
      void *p = it->second.back(); it->second.pop_back();
      g_cache_live[p] = CacheMeta{ g_pool_stream, cls };
      return p;
    }
  }
  void *p = nullptr;
  cudaError_t e = cudaMallocAsync(&p, cls, g_pool_stream);
  if (e != cudaSuccess) {                            // out of memory: drop the cache and retry once
    cudaGetLastError();
    cache_release_all();
    cudaStreamSynchronize(g_pool_stream);
    CUDA_CHECK(cudaMallocAsync(&p, cls, g_pool_stream));
  }
  std::lock_guard<std::mutex> lk(g_cache_mu);
  g_cache_live[p] = CacheMeta{ g_pool_stream, cls };
  return p;
}
void cache_free(void *p)
{
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  auto it = g_cache_live.find(p);
  if (it == g_cache_live.end()) { cudaFreeAsync(p, g_pool_stream); return; }     // not ours (should not happen)
  g_cache_free[CacheKey{ it->second.st, it->second.cls }].push_back(p);
  g_cache_live.erase(it);
}
void cache_release_all()
{
  std::lock_guard<std::mutex> lk(g_cache_mu);
  for (auto &kv : g_cache_free) for (void *p : kv.second) cudaFreeAsync(p, kv.first.st);
  g_cache_free.clear();
}

void Level::free_all()
{
  ahf::dfree(ckey); ahf::dfree(xbreak); ahf::dfree(dens); ahf::dfree(interior); ahf::dfree(tn); ahf::dfree(mark); ahf::dfree(nbr);
  ahf::dfree(crow); ahf::dfree(count); ahf::dfree(hkey); ahf::dfree(hval); ahf::dfree(rowkey); ahf::dfree(row_c0); ahf::dfree(row_tested); ahf::dfree(row_flags);
  ahf::dfree(plane_r0); ahf::dfree(rowplane); ahf::dfree(plist); ahf::dfree(pcell); ahf::dfree(lpos);
  ahf::dfree(parent); ahf::dfree(cidx); ahf::dfree(cbase); ahf::dfree(cpar); ahf::dfree(pstat); ahf::dfree(tlist); ahf::dfree(tstart);
  *this = Level();
}
}  // namespace ahf

using namespace ahf;

void ahfgpu_ctx::stage_reset()
{
  for (auto &s : stages) { event_pool.push_back(s.a); event_pool.push_back(s.b); }
  stages.clear(); stage_ms.clear(); stage_cnt.clear(); stage_cnt_extra.clear(); stage_wall.clear(); stages_resolved = true;
}
void ahfgpu_ctx::stage_resolve()
{
  if (stages_resolved) return;
  stage_ms.clear(); stage_cnt.clear();
  for (auto &s : stages) {
    float ms = 0.f;
    cudaEventSynchronize(s.b);
    cudaEventElapsedTime(&ms, s.a, s.b);
    stage_ms[s.name] += ms; stage_cnt[s.name] += s.count;
  }
  stages_resolved = true;
}
void ahfgpu_ctx::ktime_mark(const char *name, bool begin)
{
  auto take = [&](cudaEvent_t &e) { if (event_pool.empty()) cudaEventCreate(&e); else { e = event_pool.back(); event_pool.pop_back(); } };
  if (begin) { KRec r; r.name = name; take(r.a); take(r.b); cudaEventRecord(r.a, stream); krecs.push_back(r); }
  else if (!krecs.empty()) cudaEventRecord(krecs.back().b, stream);
}
void ahfgpu_ctx::ktime_dump(const char *label)
{
  if (krecs.empty()) return;
  cudaStreamSynchronize(stream);
  std::map<std::string, std::pair<double, int>> sum;
  double tk = 0.0, tg = 0.0, span = 0.0;
  float ms = 0.f;
  for (size_t i = 0; i < krecs.size(); i++) {
    cudaEventElapsedTime(&ms, krecs[i].a, krecs[i].b);
    sum[krecs[i].name].first += ms; sum[krecs[i].name].second++; tk += ms;
    if (i + 1 < krecs.size()) { cudaEventElapsedTime(&ms, krecs[i].b, krecs[i + 1].a); tg += ms; }
  }
  cudaEventElapsedTime(&ms, krecs.front().a, krecs.back().b); span = ms;
  std::vector<std::pair<double, std::string>> v;
  for (auto &kv : sum) v.push_back({ kv.second.first, kv.first });
  std::sort(v.begin(), v.end(), [](const std::pair<double, std::string> &x, const std::pair<double, std::string> &y) { return x.first > y.first; });
  fprintf(stderr, "[ktime] %s: %zu launches, span %.3f ms, kernels %.3f ms, between kernels %.3f ms\n", label, krecs.size(), span, tk, tg);
  for (auto &e : v) fprintf(stderr, "[ktime]   %-44s %4d x  %8.4f ms  %5.1f%%\n", e.second.c_str(), sum[e.second].second, e.first, 100.0 * e.first / span);
  for (auto &r : krecs) { event_pool.push_back(r.a); event_pool.push_back(r.b); }
  krecs.clear();
}
void ahfgpu_ctx::wait_mom(bool host)
{
  if (ids_pending) {
    if (host) cudaEventSynchronize(ev_ids);
    cudaStreamWaitEvent(stream, ev_ids, 0);
    ids_pending = false;
  }
  if (!mom_pending) return;
  if (host) cudaEventSynchronize(ev_mom);
  cudaStreamWaitEvent(stream, ev_mom, 0);
  mom_pending = false;
}
void ahfgpu_ctx::free_particles()
{
  wait_mom(false);
  if (adopted) { pos4 = mom4 = nullptr; keys = nullptr; adopted = false; }
  ahf::dfree(pos4); ahf::dfree(mom4); ahf::dfree(keys); ahf::dfree(order);
  pos4 = mom4 = nullptr; keys = nullptr; order = nullptr; n = 0;
}
void ahfgpu_ctx::free_levels()
{
  for (auto &l : levels) l.free_all();
  levels.clear(); pstat_split.clear();
  ahf::dfree(owner_level); owner_level = nullptr;
}
void ahfgpu_ctx::free_halos()
{
  ahf::dfree(h_scal); ahf::dfree(h_moff); ahf::dfree(h_members); ahf::dfree(h_poff); ahf::dfree(h_prof);
  ahf::dfree(h_species); ahf::dfree(h_prof_species); h_species = nullptr; h_prof_species = nullptr;
  h_scal = nullptr; h_moff = nullptr; h_members = nullptr; h_poff = nullptr; h_prof = nullptr;
  nhalo = 0; h_total_members = h_total_bins = 0; early_sent = false;
}

#define API_BEGIN try {
#define API_END                                                                                         \
  return 0; }                                                                                           \
  catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }                                  \
  catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }                          \
  catch (...) { ahf::g_last_error = "unknown exception"; return -3; }

static void check_params(const ahfgpu_params *p)
{
  if (!p) AHF_FAIL("null params");
  if (p->lgrid_dom < 4 || (p->lgrid_dom & (p->lgrid_dom - 1)) != 0) AHF_FAIL("lgrid_dom must be a power of two >= 4");
  if (p->lgrid_dom > (1 << 21)) AHF_FAIL("lgrid_dom above 2^21");
}

// context creation + module load + per-device kernel attributes / __constant__ symbols: once for each ordinal
static void device_setup(int dev)
{
  const bool tm = getenv("AHFGPU_INIT_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  CUDA_CHECK(cudaSetDevice(dev));
  static std::mutex mu; static bool done[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 64 || !done[dev]) {
    CUDA_CHECK(cudaFree(nullptr));
    const double t1 = now();
    ahf::mesh_device_init();
    const double t2 = now();
    ahf::sfc_device_init();
    if (dev < 64) done[dev] = true;
    if (tm) fprintf(stderr, "AHFGPU_INIT_TIMING context=%.3f mesh_device_init=%.3f sfc_device_init=%.3f\n", t1 - t0, t2 - t1, now() - t2);
  }
}

extern "C" {

const char *ahfgpu_last_error(void) { return ahf::g_last_error.c_str(); }

int ahfgpu_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// creates the CUDA context of `device` and loads the kernels ahead of ahfgpu_init: a host program calls it from a helper thread while it
// parses its parameter file, so that the 0.5-1 s of driver start-up do not sit on its critical path (ahf_b200/host/ahf_glue.c)
int ahfgpu_warmup(int32_t device)
{
  API_BEGIN
  int ndev = 0;
  CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) AHF_FAIL("no CUDA device: libahfgpu has no CPU fallback");
  if (device < 0 || device >= ndev) AHF_FAIL("device ordinal out of range");
  device_setup(device);
  API_END
}

int ahfgpu_init(ahfgpu_ctx **out, const ahfgpu_params *par)
{
  API_BEGIN
  if (!out) AHF_FAIL("null ctx pointer");
  check_params(par);
  int ndev = 0;
  CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) AHF_FAIL("no CUDA device: libahfgpu has no CPU fallback");
  if (par->device < 0 || par->device >= ndev) AHF_FAIL("device ordinal out of range");
  device_setup(par->device);
  ahfgpu_ctx *c = new ahfgpu_ctx();
  c->par = *par; c->dev = par->device;
  try {
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, par->device));
    uint64_t keep = ~0ull;
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  } catch (...) {
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    throw;
  }
  *out = c;
  API_END
}

int ahfgpu_set_params(ahfgpu_ctx *c, const ahfgpu_params *par)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  check_params(par);
  int dev = c->dev;
  c->par = *par; c->par.device = dev;
  API_END
}

int ahfgpu_finalize(ahfgpu_ctx *c)
{
  API_BEGIN
  if (!c) return 0;
  cudaSetDevice(c->dev); ahf::g_pool_stream = c->stream;
  c->wait_mom(true);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->stream);
  c->stage_reset(); c->free_halos(); c->free_levels(); c->free_particles();
  slab_free(c);
  if (c->comm) { delete c->comm; c->comm = nullptr; }
  ahf::dfree(c->in_pos); ahf::dfree(c->in_mom); ahf::dfree(c->in_w); ahf::dfree(c->in_u);
  ahf::dfree(c->scan_state); c->scan_state = nullptr; c->scan_cap = 0;
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->h_up) cudaFreeHost(c->h_up);
  c->h_up = nullptr; c->h_up_bytes = 0;
  c->h_pin = nullptr; c->h_pin_bytes = 0;
  for (auto &e : c->event_pool) cudaEventDestroy(e);
  c->event_pool.clear();
  for (auto &e : c->ev) if (e) cudaEventDestroy(e);
  if (c->ev_members) { cudaEventSynchronize(c->ev_members); cudaEventDestroy(c->ev_members); }
  if (c->d2h_stream) { cudaStreamSynchronize(c->d2h_stream); cudaStreamDestroy(c->d2h_stream); cudaEventDestroy(c->ev_ids); }
  {                                                   // blocks cached for this context's stream go back to the driver pool
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto it = g_cache_free.begin(); it != g_cache_free.end();) {
      if (it->first.st == c->stream) { for (void *p : it->second) cudaFreeAsync(p, c->stream); it = g_cache_free.erase(it); }
      else ++it;
    }
  }
  cudaStreamSynchronize(c->stream);
  cudaStreamDestroy(c->stream);
  if (c->copy_stream) {
    cudaStreamDestroy(c->copy_stream);
    for (auto &e : c->ev_copy) if (e) cudaEventDestroy(e);
    cudaEventDestroy(c->ev_main); cudaEventDestroy(c->ev_mom);
  }
  delete c;
  API_END
}

int ahfgpu_sfc_sort_particles(ahfgpu_ctx *c, void *part, uint64_t n, uint32_t stride, int32_t off_pos, int32_t off_mom,
                              int32_t off_key, int32_t off_id, int32_t off_weight, int32_t off_u)
{
  API_BEGIN
  if (!c || (!part && n)) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  sfc_sort_aos(c, part, n, stride, off_pos, off_mom, off_key, off_id, off_weight, off_u);
  API_END
}

int ahfgpu_sfc_sort_soa(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *weight, const float *u, uint64_t n,
                        uint64_t *keys_out, uint32_t *order_out)
{
  API_BEGIN
  if (!c || ((!pos3 || !mom3) && n)) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  c->ktime = getenv("AHFGPU_KTIME") != nullptr;
  sfc_sort_soa(c, pos3, mom3, weight, u, n, keys_out, order_out);
  c->ktime_dump("ahfgpu_sfc_sort_soa");
  API_END
}

int ahfgpu_sfc_sort_soa_async(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *weight, const float *u, uint64_t n)
{
  API_BEGIN
  if (!c || ((!pos3 || !mom3) && n)) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  sfc_sort_soa_async(c, pos3, mom3, weight, u, n);
  API_END
}

int ahfgpu_upload_soa(ahfgpu_ctx *c, const float *pos3, const float *mom3, const float *weight, const float *u, uint64_t n)
{
  API_BEGIN
  if (!c || ((!pos3 || !mom3) && n)) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  sfc_upload_soa(c, pos3, mom3, weight, u, n);
  API_END
}

int ahfgpu_sfc_sort_resident(ahfgpu_ctx *c)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  c->ktime = getenv("AHFGPU_KTIME") != nullptr;
  sfc_sort_resident(c, nullptr, nullptr);
  c->ktime_dump("ahfgpu_sfc_sort_resident");
  API_END
}

int ahfgpu_sfc_sort_device4(ahfgpu_ctx *c, const void *pos4_dev, const void *mom4_dev, uint64_t n, int32_t has_weight, int32_t has_u)
{
  API_BEGIN
  if (!c || ((!pos4_dev || !mom4_dev) && n)) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  sfc_sort_device4(c, pos4_dev, mom4_dev, n, has_weight != 0, has_u != 0);
  API_END
}

int ahfgpu_set_global_count(ahfgpu_ctx *c, uint64_t n_total)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  c->n_total = n_total;
  API_END
}

// ---- ONE box on several GPUs (comm.cuh, slab.cu) ----------------------------------------------------------------------------------
int ahfgpu_comm_nccl_unique_id(void *id128)
{
  API_BEGIN
  if (!id128) AHF_FAIL("null argument");
  comm_nccl_unique_id(id128);
  API_END
}

int ahfgpu_comm_init_nccl(ahfgpu_ctx *c, int32_t rank, int32_t nranks, const void *id128)
{
  API_BEGIN
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) AHF_FAIL("bad argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  if (c->comm) { delete c->comm; c->comm = nullptr; }
  c->comm = comm_create_nccl(rank, nranks, id128, c->dev);
  API_END
}

void *ahfgpu_comm_local_group_create(int32_t nranks)
{
  try { if (nranks < 1) return nullptr; return comm_local_group_create(nranks); } catch (...) { return nullptr; }
}

int ahfgpu_comm_local_group_destroy(void *group)
{
  API_BEGIN
  comm_local_group_destroy(group);
  API_END
}

int ahfgpu_comm_local_group_abort(void *group)
{
  API_BEGIN
  comm_local_group_abort(group);
  API_END
}

int ahfgpu_comm_init_local(ahfgpu_ctx *c, int32_t rank, void *group)
{
  API_BEGIN
  if (!c || !group) AHF_FAIL("bad argument");
  if (c->comm) { delete c->comm; c->comm = nullptr; }
  c->comm = comm_create_local(rank, group);
  API_END
}

int ahfgpu_slab_distribute(ahfgpu_ctx *c, uint64_t id_base, double ghost_width, int32_t decomp_bits)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  slab_distribute(c, id_base, ghost_width, decomp_bits);
  API_END
}

int ahfgpu_slab_info(ahfgpu_ctx *c, int64_t *iout, double *dout)
{
  API_BEGIN
  if (!c || !c->slab || !c->comm) AHF_FAIL("the resident particles are not a slab of a distributed box");
  const Slab &S = *c->slab;
  if (iout) {
    iout[0] = c->comm->rank; iout[1] = c->comm->nranks; iout[2] = (int64_t)c->n; iout[3] = (int64_t)S.own_lo; iout[4] = (int64_t)S.own_hi;
    iout[5] = (int64_t)S.n_total; iout[6] = S.bd; iout[7] = S.T; iout[8] = (int64_t)S.split[c->comm->rank]; iout[9] = (int64_t)S.split[c->comm->rank + 1];
    iout[10] = c->g_nlevels; iout[11] = c->comm->coll_calls;
  }
  if (dout) { dout[0] = S.ghost_width; dout[1] = c->comm->coll_ms; dout[2] = (double)c->comm->coll_bytes; dout[3] = 0.0; }
  API_END
}

int ahfgpu_slab_owner_of(ahfgpu_ctx *c, int64_t n, const double *pos3, int32_t *owner)
{
  API_BEGIN
  if (!c || !c->slab || !c->comm) AHF_FAIL("the resident particles are not a slab of a distributed box");
  if (n && (!pos3 || !owner)) AHF_FAIL("null argument");
  const Slab &S = *c->slab;
  const int R = c->comm->nranks;
  for (int64_t i = 0; i < n; i++) {
    double q[3];
    for (int d = 0; d < 3; d++) { q[d] = pos3[3 * i + d]; q[d] -= std::floor(q[d]); if (q[d] >= 1.0) q[d] = 0.0; }
    const uint64_t h = hilbert_key_posd(q[0], q[1], q[2], (unsigned)S.bd);
    int r = 0;
    while (r + 1 < R && S.split[r + 1] <= h) r++;
    owner[i] = r;
  }
  API_END
}

int ahfgpu_particles_get(ahfgpu_ctx *c, float *pos4, float *mom4)
{
  API_BEGIN
  if (!c || !c->pos4) AHF_FAIL("no resident particles");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->wait_mom(true);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (pos4 && c->n) CUDA_CHECK(cudaMemcpy(pos4, c->pos4, sizeof(float4) * c->n, cudaMemcpyDeviceToHost));
  if (mom4 && c->n) CUDA_CHECK(cudaMemcpy(mom4, c->mom4, sizeof(float4) * c->n, cudaMemcpyDeviceToHost));
  API_END
}

int ahfgpu_particle_ids(ahfgpu_ctx *c, uint32_t *ids)
{
  API_BEGIN
  if (!c || !c->order) AHF_FAIL("no resident particle index (order) array");
  if (!ids && c->n) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (c->n) CUDA_CHECK(cudaMemcpy(ids, c->order, sizeof(uint32_t) * c->n, cudaMemcpyDeviceToHost));
  API_END
}

// the same copy enqueued behind the momentum upload of ahfgpu_sfc_sort_soa_async on the library's copy stream: it travels (device -> host,
// the idle direction of the bus) while the hierarchy is built and is complete when the next call that waits for the momenta returns
// (ahfgpu_construct_halos).  ids should be pinned memory; without a pending asynchronous sort this is ahfgpu_particle_ids.
int ahfgpu_particle_ids_async(ahfgpu_ctx *c, uint32_t *ids)
{
  API_BEGIN
  if (!c || !c->order) AHF_FAIL("no resident particle index (order) array");
  if (!ids && c->n) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  if (!c->mom_pending || !c->copy_stream) {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->n) CUDA_CHECK(cudaMemcpy(ids, c->order, sizeof(uint32_t) * c->n, cudaMemcpyDeviceToHost));
  } else {
    // ahfgpu_sfc_sort_soa_async has returned: the main stream is idle and `order` is final.  Own stream: the copy shares the bus with the
    // incoming momenta (other direction) instead of queueing behind them
    if (!c->d2h_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_ids, cudaEventDisableTiming));
    }
    if (c->n) CUDA_CHECK(cudaMemcpyAsync(ids, c->order, sizeof(uint32_t) * c->n, cudaMemcpyDeviceToHost, c->d2h_stream));
    CUDA_CHECK(cudaEventRecord(c->ev_ids, c->d2h_stream));        // whoever waits for the momenta waits for this copy as well (wait_mom)
    c->ids_pending = true;
  }
  API_END
}

int ahfgpu_adopt_sorted(ahfgpu_ctx *c, const void *pos4_dev, const void *mom4_dev, const void *keys_dev, uint64_t n, int32_t has_weight, int32_t has_u)
{
  API_BEGIN
  if (!c || ((!pos4_dev || !mom4_dev || !keys_dev) && n)) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->free_halos(); c->free_levels(); c->free_particles();
  c->pos4 = (float4 *)pos4_dev; c->mom4 = (float4 *)mom4_dev; c->keys = (uint64_t *)keys_dev; c->order = nullptr;
  c->n = n; c->adopted = true; c->has_weight = has_weight != 0; c->has_u = has_u != 0;
  API_END
}

void *ahfgpu_device_ptr(ahfgpu_ctx *c, const char *name)
{
  if (!c || !name) return nullptr;
  if (!strcmp(name, "pos4")) return c->pos4;
  if (!strcmp(name, "mom4")) { c->wait_mom(true); return c->mom4; }
  if (!strcmp(name, "keys")) return c->keys;
  return nullptr;
}

int ahfgpu_event_record(ahfgpu_ctx *c, int32_t slot)
{
  API_BEGIN
  if (!c || slot < 0 || slot >= 16) AHF_FAIL("bad event slot");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  if (!c->ev[slot]) CUDA_CHECK(cudaEventCreate(&c->ev[slot]));
  CUDA_CHECK(cudaEventRecord(c->ev[slot], c->stream));
  API_END
}

double ahfgpu_event_elapsed_ms(ahfgpu_ctx *c, int32_t a, int32_t b)
{
  if (!c || a < 0 || b < 0 || a >= 16 || b >= 16 || !c->ev[a] || !c->ev[b]) return -1.0;
  float ms = 0.f;
  cudaSetDevice(c->dev);
  if (cudaEventSynchronize(c->ev[b]) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) return -1.0;
  return (double)ms;
}

int ahfgpu_synchronize(ahfgpu_ctx *c)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->wait_mom(true);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  API_END
}

int ahfgpu_hilbert_keys(ahfgpu_ctx *c, const float *pos3, uint64_t n, uint32_t bits, uint64_t *keys_out)
{
  API_BEGIN
  if (!c || ((!pos3 || !keys_out) && n)) AHF_FAIL("null argument");
  if (bits < 1 || bits > 21) AHF_FAIL("bits must be in 1..21");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  sfc_keys_only(c, pos3, n, bits, keys_out);
  API_END
}

int ahfgpu_build_amr(ahfgpu_ctx *c)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  if (!c->pos4) AHF_FAIL("no resident particles: call ahfgpu_sfc_sort_* first");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  c->ktime = getenv("AHFGPU_KTIME") != nullptr;
  amr_build(c);
  c->ktime_dump("ahfgpu_build_amr");
  API_END
}

int ahfgpu_amr_nlevels(ahfgpu_ctx *c) { return c ? (int)c->levels.size() : -1; }

int ahfgpu_amr_level_header(ahfgpu_ctx *c, int32_t lev, int64_t *iout, double *dout)
{
  API_BEGIN
  if (!c || lev < 0 || lev >= (int)c->levels.size()) AHF_FAIL("bad level");
  const Level &l = c->levels[lev];
  if (iout) { iout[0] = l.L; iout[1] = l.ncell; iout[2] = l.npart_dep; iout[3] = l.npart_final; }
  if (dout) { dout[0] = l.critdens; dout[1] = l.masstopartdens; }
  API_END
}

int ahfgpu_halo_sizes(ahfgpu_ctx *c, int64_t *total_members, int64_t *total_bins)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  if (total_members) *total_members = c->h_total_members;
  if (total_bins) *total_bins = c->h_total_bins;
  API_END
}

int ahfgpu_construct_halos(ahfgpu_ctx *c, int64_t nhalo, const double *centre3, const double *gather_rad, const int64_t *seed)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  if (!c->pos4) AHF_FAIL("no resident particles: call ahfgpu_sfc_sort_* first");
  if (nhalo && (!centre3 || !gather_rad)) AHF_FAIL("null argument");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  c->stage_reset();
  c->wait_mom(false);                                  // momenta of ahfgpu_sfc_sort_soa_async: device-side wait, the host does not block
  c->ktime = getenv("AHFGPU_KTIME") != nullptr;
  halos_construct(c, nhalo, centre3, gather_rad, seed);
  c->ktime_dump("ahfgpu_construct_halos");
  API_END
}

int ahfgpu_halo_members_buffer(ahfgpu_ctx *c, int64_t *members_pinned, int64_t capacity)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  if (members_pinned && capacity < 0) AHF_FAIL("negative capacity");
  CUDA_CHECK(cudaSetDevice(c->dev));
  if (c->early_sent && c->ev_members) CUDA_CHECK(cudaEventSynchronize(c->ev_members));       // a copy into the old buffer may still be in flight
  c->early_members = members_pinned; c->early_cap = members_pinned ? capacity : 0; c->early_sent = false;
  API_END
}

int ahfgpu_halo_fetch(ahfgpu_ctx *c, double *scal, int64_t *member_offset, int64_t *members, int64_t *prof_offset, double *prof)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
  CUDA_CHECK(cudaStreamSynchronize(c->stream));       // the library's stream is non-blocking: the legacy-stream copies below do not wait for it
  if (scal && c->nhalo) CUDA_CHECK(cudaMemcpy(scal, c->h_scal, sizeof(double) * AHFGPU_NSCAL * c->nhalo, cudaMemcpyDeviceToHost));
  if (member_offset) CUDA_CHECK(cudaMemcpy(member_offset, c->h_moff, sizeof(int64_t) * (c->nhalo + 1), cudaMemcpyDeviceToHost));
  if (members && c->h_total_members) {
    if (c->early_sent && members == c->early_members) CUDA_CHECK(cudaEventSynchronize(c->ev_members));      // already on its way (ahfgpu_halo_members_buffer)
    else CUDA_CHECK(cudaMemcpy(members, c->h_members, sizeof(int64_t) * c->h_total_members, cudaMemcpyDeviceToHost));
  }
  if (prof_offset) CUDA_CHECK(cudaMemcpy(prof_offset, c->h_poff, sizeof(int64_t) * (c->nhalo + 1), cudaMemcpyDeviceToHost));
  if (prof && c->h_total_bins) CUDA_CHECK(cudaMemcpy(prof, c->h_prof, sizeof(double) * AHFGPU_NPROFCOL * c->h_total_bins, cudaMemcpyDeviceToHost));
  API_END
}

int ahfgpu_halo_fetch_species(ahfgpu_ctx *c, double *species, double *prof_species)
{
  API_BEGIN
  if (!c) AHF_FAIL("null ctx");
  if (!c->h_species) AHF_FAIL("no per-species results: the particles carry no thermal energy / type information (u)");
  CUDA_CHECK(cudaSetDevice(c->dev));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (species && c->nhalo) CUDA_CHECK(cudaMemcpy(species, c->h_species, sizeof(double) * 64 * c->nhalo, cudaMemcpyDeviceToHost));
  if (prof_species && c->h_total_bins) CUDA_CHECK(cudaMemcpy(prof_species, c->h_prof_species, sizeof(double) * 3 * c->h_total_bins, cudaMemcpyDeviceToHost));
  API_END
}

double ahfgpu_stage_ms(ahfgpu_ctx *c, const char *name)
{
  if (!c || !name) return -1.0;
  c->stage_resolve();
  const std::string nm(name);
  if (nm.size() > 5 && nm.compare(nm.size() - 5, 5, "@wall") == 0) {
    auto iw = c->stage_wall.find(nm.substr(0, nm.size() - 5));
    return iw == c->stage_wall.end() ? -1.0 : iw->second;
  }
  auto it = c->stage_ms.find(name);
  return it == c->stage_ms.end() ? -1.0 : it->second;
}

int64_t ahfgpu_stage_count(ahfgpu_ctx *c, const char *name)
{
  if (!c || !name) return -1;
  if (!strcmp(name, "launches")) return c->n_launches;
  c->stage_resolve();
  auto ie = c->stage_cnt_extra.find(name);
  if (ie != c->stage_cnt_extra.end()) return ie->second;
  auto it = c->stage_cnt.find(name);
  return it == c->stage_cnt.end() ? -1 : it->second;
}

}  // extern "C"

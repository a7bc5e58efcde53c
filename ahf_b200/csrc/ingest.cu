// ingest.cu -- NEXT-4 of SURVEY 8f: bulk GADGET ingest with unit scaling on the device.
//
// Replaces, for the snapshot kinds the path's configs use, the reference's reader + scaling: io_gadget_readpart_raw
// (src/libio/io_gadget.c:427-568; one fread per VALUE through a function pointer: local_get_block_pos :1276-1399,
// local_get_block_vel :1401-1480, local_get_block_id :1482-1562) and io_gadget_scale_particles (:857-995).  Here the three blocks
// are read with ONE pread each into pinned memory, travel to the device while the next block is being read, and extreme
// positions, shift, box check and the conversion to AHF's internal units run as kernels -- in the reference's own float32
// arithmetic: x = (x + (float)shift) * (float)(1/boxsize), p = v * (float)(sqrt(a) a / (boxsize posscale 100)) (:919-925, :947-955),
// so that positions, momenta and therefore keys are bit-identical to what the reference holds after startrun().
// Supported: GADGET-1 and GADGET-2 framing, either byte order, float32 blocks, any particle types whose masses are in the header
// (massarr > 0: no MASS block) and no gas (no U block) -- BASELINE.json configs 1-4.  Everything else is refused loudly.
#include "common.cuh"
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <memory>
#include <mutex>

namespace ahf {

namespace {

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

struct GadgetHeader {                  // io_gadget_header_def.h:41-66
  int32_t  np[6];
  double   massarr[6];
  double   expansion, redshift;
  int32_t  flagsfr, flagfeedback;
  uint32_t nall[6];
  int32_t  flagcooling, numfiles;
  double   boxsize, omega0, omegalambda, hubble;
};

void pread_all(int fd, void *dst, size_t bytes, off_t off)
{
  char *p = static_cast<char *>(dst);
  while (bytes) {
    const ssize_t r = pread(fd, p, bytes, off);
    if (r <= 0) AHF_FAIL("short read from the snapshot file");
    p += r; off += r; bytes -= (size_t)r;
  }
}

__global__ void k_bswap32(uint32_t *a, uint64_t n)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = __byte_perm(a[i], 0, 0x0123);
}
// extreme positions (io_gadget.c:1359-1364), per component; float min/max through ordered-int atomics is avoided: block reduce + atomics on
// the bit patterns of non-negative / negative floats handled by two-pass sign trick
__device__ __forceinline__ void atomic_min_f(float *a, float v) { if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v)); else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v)); }
__device__ __forceinline__ void atomic_max_f(float *a, float v) { if (v >= 0.f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v)); else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v)); }
__global__ void k_minmax3(const float *__restrict__ pos3, uint64_t n, float *__restrict__ mm /* min xyz, max xyz */)
{
  float lo[3] = { 1e38f, 1e38f, 1e38f }, hi[3] = { -1e38f, -1e38f, -1e38f };
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; d++) { const float v = pos3[3 * i + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
    if ((threadIdx.x & 31) == 0) { atomic_min_f(mm + d, lo[d]); atomic_max_f(mm + 3 + d, hi[d]); }
  }
}
// SCALE_CALL(float) of io_gadget.c:947-955: two separate float operations per position component (no fused multiply-add)
__global__ void k_scale(float *__restrict__ pos3, float *__restrict__ mom3, uint64_t n3, float sx, float sy, float sz, float scale_pos, float scale_mom)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const int d = (int)(i % 3);
  const float sh = d == 0 ? sx : d == 1 ? sy : sz;
  pos3[i] = __fmul_rn(__fadd_rn(pos3[i], sh), scale_pos);
  mom3[i] = __fmul_rn(mom3[i], scale_mom);
}

}  // namespace

// what the host-only half of the ingest leaves: header, framing, and (when prefetched) the three blocks in pageable memory
struct Snapshot {
  GadgetHeader H;
  bool     swapped = false;
  int      ver = 1;
  uint64_t n = 0;
  double   pmass = 0.0;
  off_t    p_pos = 0, p_vel = 0, p_id = 0;
  std::vector<float>    pos, vel;       // filled by the prefetch only
  std::vector<uint32_t> id;
  double   read_ms = 0.0;
};

// framing + header + block offsets (io_gadget.c:1107-1231, :427-470); refuses what the bulk path does not cover
static void snapshot_open(int fd, Snapshot &S)
{
  uint32_t first = 0;
  pread_all(fd, &first, 4, 0);
  bool swapped = false; int ver = 1;
  if (first == 256u) { ver = 1; }
  else if (bswap32(first) == 256u) { ver = 1; swapped = true; }
  else if (first == 8u) { ver = 2; }
  else if (bswap32(first) == 8u) { ver = 2; swapped = true; }
  else AHF_FAIL("not a GADGET-1/2 file (first block length is neither 256 nor 8)");
  auto rd32 = [&](off_t off) { uint32_t v; pread_all(fd, &v, 4, off); return swapped ? bswap32(v) : v; };
  off_t off = 0;
  auto block = [&](const char *label, uint64_t expect_bytes) -> off_t {      // returns the payload offset of the next block, checks its length
    if (ver == 2) {                                                           // label block: [8]["NAME"][u32 next][8]
      if (rd32(off) != 8u) AHF_FAIL("GADGET-2 label block expected");
      char nm[5] = { 0, 0, 0, 0, 0 };
      pread_all(fd, nm, 4, off + 4);
      if (strncmp(nm, label, 4) != 0) AHF_FAIL(std::string("wrong block: expected ") + label + ", found " + nm);
      off += 16;
    }
    const uint32_t len = rd32(off);
    if (expect_bytes && len != expect_bytes) AHF_FAIL(std::string("unexpected length of block ") + label + " (only float32 / uint32 blocks are supported)");
    const off_t payload = off + 4;
    if (rd32(payload + len) != len) AHF_FAIL(std::string("block boundaries of ") + label + " disagree: corrupt file?");
    off = payload + len + 4;
    return payload;
  };
  GadgetHeader &H = S.H;
  {
    const off_t p = block("HEAD", 256);
    static_assert(sizeof(GadgetHeader) <= 256, "header layout");
    pread_all(fd, &H, sizeof(H), p);
    if (swapped) {
      for (auto &v : H.np) v = (int32_t)bswap32((uint32_t)v);
      auto sw = [](double &d) { uint64_t u; memcpy(&u, &d, 8); u = bswap64(u); memcpy(&d, &u, 8); };
      for (auto &v : H.massarr) sw(v);
      sw(H.expansion); sw(H.redshift); sw(H.boxsize); sw(H.omega0); sw(H.omegalambda); sw(H.hubble);
      H.numfiles = (int32_t)bswap32((uint32_t)H.numfiles);
    }
  }
  uint64_t n = 0;
  double pmass = 0.0;
  for (int t = 0; t < 6; t++) {
    if (H.np[t] < 0) AHF_FAIL("negative particle count in the header");
    n += (uint64_t)H.np[t];
    if (H.np[t] > 0 && !(H.massarr[t] > 0.0)) AHF_FAIL("particle types with individual masses (MASS block) are not supported by the bulk ingest");
    if (H.np[t] > 0 && (pmass == 0.0 || H.massarr[t] < pmass)) pmass = H.massarr[t];
  }
  if (H.np[1] > 0) pmass = H.massarr[1];                       // weights are in units of the type-1 mass (io_gadget.c:1595-1599)
  if (H.np[0] > 0) AHF_FAIL("gas particles (U block) are not supported by the bulk ingest");
  if (H.numfiles > 1) AHF_FAIL("multi-file snapshots are read one file per call: not in this round");
  if (n == 0 || n >= (1ull << 32)) AHF_FAIL("particle count out of range");
  S.swapped = swapped; S.ver = ver; S.n = n; S.pmass = pmass;
  S.p_pos = block("POS ", 12 * n); S.p_vel = block("VEL ", 12 * n); S.p_id = block("ID  ", 4 * n);
}

// snapshots read ahead of the CUDA context (ahfgpu_ingest_prefetch), by path
static std::mutex g_pref_mu;
static std::map<std::string, std::unique_ptr<Snapshot>> g_pref;

void ingest_prefetch(const char *path)
{
  const auto t0 = std::chrono::steady_clock::now();
  const int fd = open(path, O_RDONLY);
  if (fd < 0) AHF_FAIL(std::string("cannot open ") + path);
  struct Closer { int fd; ~Closer() { close(fd); } } closer{ fd };
  std::unique_ptr<Snapshot> S(new Snapshot());
  snapshot_open(fd, *S);
  const uint64_t n = S->n;
  S->pos.resize(3 * n); S->vel.resize(3 * n); S->id.resize(n);
  pread_all(fd, S->pos.data(), 12 * n, S->p_pos);
  pread_all(fd, S->vel.data(), 12 * n, S->p_vel);
  pread_all(fd, S->id.data(), 4 * n, S->p_id);
  S->read_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::lock_guard<std::mutex> lk(g_pref_mu);
  g_pref[path] = std::move(S);
}

// info[24]: 0 particles, 1 boxsize (after the extent check, file units x posscale), 2 expansion, 3 omega0, 4 lambda0, 5 pmass (mass of a
// type-1 particle x weightscale, or of the lightest type present), 6-8 shift applied, 9 scale_pos, 10 scale_mom, 11 version (1/2),
// 12 byte swapped, 13 hubble parameter of the file, 14 milliseconds reading (host wall clock), 15 milliseconds on the device (events),
// 16-18 smallest / 19-21 largest raw position per axis (f->minpos / f->maxpos of the reference's file object), 22 boxsize after the extent
// check in file units (what the reference leaves in header->boxsize), 23 reserved
void ingest_gadget(ahfgpu_ctx *c, const char *path, double posscale, double weightscale, uint64_t *ids_out, double *info)
{
  const auto t0 = std::chrono::steady_clock::now();
  std::unique_ptr<Snapshot> pre;
  {
    std::lock_guard<std::mutex> lk(g_pref_mu);
    auto it = g_pref.find(path);
    if (it != g_pref.end()) { pre = std::move(it->second); g_pref.erase(it); }
  }
  int fd = -1;
  struct Closer { int &fd; ~Closer() { if (fd >= 0) close(fd); } } closer{ fd };
  Snapshot S_local;
  if (!pre) {
    fd = open(path, O_RDONLY);
    if (fd < 0) AHF_FAIL(std::string("cannot open ") + path);
    snapshot_open(fd, S_local);
  }
  const Snapshot &S = pre ? *pre : S_local;
  const GadgetHeader &H = S.H;
  const uint64_t n = S.n;
  const bool swapped = S.swapped;
  const int ver = S.ver;
  const double pmass = S.pmass;
  const off_t p_pos = S.p_pos, p_vel = S.p_vel, p_id = S.p_id;
  // ---- bulk read + upload, block by block; ids stay on the host (the path carries the input index, ahfgpu_particle_ids)
  c->wait_mom(false);
  dfree(c->in_pos); dfree(c->in_mom); dfree(c->in_w); dfree(c->in_u);
  c->in_pos = c->in_mom = c->in_w = c->in_u = nullptr; c->in_n = n;
  c->in_pos = static_cast<float *>(cache_alloc(12 * n)); c->in_mom = static_cast<float *>(cache_alloc(12 * n));
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  double read_ms = 0.0;
  if (pre) {
    // the blocks were read while the CUDA context was still being created: upload from pageable memory (the driver stages it; pinning
    // 24 n bytes for one use would cost more than the staging)
    CUDA_CHECK(cudaEventRecord(e0, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->in_pos, pre->pos.data(), 12 * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->in_mom, pre->vel.data(), 12 * n, cudaMemcpyHostToDevice, c->stream));
    if (ids_out) for (uint64_t i = 0; i < n; i++) ids_out[i] = swapped ? bswap32(pre->id[i]) : pre->id[i];
    read_ms = pre->read_ms;
  } else {
  if (24 * n > c->h_up_bytes) {                                 // pinned staging of the context (kept between calls): positions | velocities
    if (c->h_up) cudaFreeHost(c->h_up);
    c->h_up = nullptr; c->h_up_bytes = 0;
    CUDA_CHECK(cudaHostAlloc(&c->h_up, 24 * n, cudaHostAllocDefault));
    c->h_up_bytes = 24 * n;
  }
  void *stage2 = static_cast<char *>(c->h_up) + 12 * n;
  pread_all(fd, c->h_up, 12 * n, p_pos);
  CUDA_CHECK(cudaEventRecord(e0, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(c->in_pos, c->h_up, 12 * n, cudaMemcpyHostToDevice, c->stream));
  pread_all(fd, stage2, 12 * n, p_vel);                         // overlaps the position upload
  CUDA_CHECK(cudaMemcpyAsync(c->in_mom, stage2, 12 * n, cudaMemcpyHostToDevice, c->stream));
  if (ids_out) {
    std::vector<uint32_t> raw(n);
    pread_all(fd, raw.data(), 4 * n, p_id);
    for (uint64_t i = 0; i < n; i++) ids_out[i] = swapped ? bswap32(raw[i]) : raw[i];
  }
  read_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  if (swapped) {
    LAUNCH(c, k_bswap32, (unsigned)((3 * n + 255) / 256), 256, 0, reinterpret_cast<uint32_t *>(c->in_pos), 3 * n);
    LAUNCH(c, k_bswap32, (unsigned)((3 * n + 255) / 256), 256, 0, reinterpret_cast<uint32_t *>(c->in_mom), 3 * n);
  }
  // ---- extreme positions, shift, box check, scaling (io_gadget_scale_particles)
  DevBuf<float> mm;
  mm.reserve(6);
  const float init[6] = { 1e38f, 1e38f, 1e38f, -1e38f, -1e38f, -1e38f };
  CUDA_CHECK(cudaMemcpyAsync(mm.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_minmax3, 148 * 8, 256, 0, c->in_pos, n, mm.p);
  float hmm[6];
  read_back(c, hmm, mm.p, sizeof(hmm));
  double boxsize = H.boxsize, shift[3];
  for (int d = 0; d < 3; d++) {
    const double ext = std::fabs((double)hmm[3 + d] - (double)hmm[d]);
    if (ext > boxsize) boxsize = ext;                            // :879-896
    shift[d] = ((double)hmm[d] < 0.0) ? -(double)hmm[d] : 0.0;   // :906-908
  }
  const double scale_pos = 1.0 / boxsize;
  const double scale_mom = std::sqrt(H.expansion) * H.expansion / (boxsize * posscale * 100.);
  LAUNCH(c, k_scale, (unsigned)((3 * n + 255) / 256), 256, 0, c->in_pos, c->in_mom, 3 * n, (float)shift[0], (float)shift[1], (float)shift[2], (float)scale_pos,
         (float)scale_mom);
  CUDA_CHECK(cudaEventRecord(e1, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  float dev_ms = 0.f;
  cudaEventElapsedTime(&dev_ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  mm.release();
  if (info) {
    info[0] = (double)n; info[1] = boxsize * posscale; info[2] = H.expansion; info[3] = H.omega0; info[4] = H.omegalambda; info[5] = pmass * weightscale;
    info[6] = shift[0]; info[7] = shift[1]; info[8] = shift[2]; info[9] = scale_pos; info[10] = scale_mom; info[11] = ver; info[12] = swapped ? 1 : 0;
    info[13] = H.hubble; info[14] = read_ms; info[15] = dev_ms;
    for (int d = 0; d < 3; d++) { info[16 + d] = (double)hmm[d]; info[19 + d] = (double)hmm[3 + d]; }
    info[22] = boxsize; info[23] = 0.0;
  }
}

}  // namespace ahf

// one particle of the UNSORTED input set (after ahfgpu_upload_soa / ahfgpu_ingest_gadget, before the sort releases it): the reference's
// startrun prints the first and the last particle into its log file (startrun.c:378-431)
// host-only half of ahfgpu_ingest_gadget: header checks and the three block reads into pageable memory, kept until the next
// ahfgpu_ingest_gadget of the same path consumes them -- callable before any CUDA context exists (the drop-in program reads the
// snapshot while its helper thread is still creating the context)
extern "C" int ahfgpu_ingest_prefetch(const char *path)
{
  try {
    if (!path) AHF_FAIL("null argument");
    ahf::ingest_prefetch(path);
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

extern "C" int ahfgpu_input_peek(ahfgpu_ctx *c, uint64_t index, float *pos3, float *mom3)
{
  try {
    if (!c || !pos3 || !mom3) AHF_FAIL("null argument");
    if (!c->in_pos || !c->in_mom || index >= c->in_n) AHF_FAIL("no unsorted input particle with that index is resident");
    CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
    CUDA_CHECK(cudaMemcpyAsync(pos3, c->in_pos + 3 * index, 12, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(mom3, c->in_mom + 3 * index, 12, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

extern "C" int ahfgpu_ingest_gadget(ahfgpu_ctx *c, const char *path, double posscale, double weightscale, uint64_t *ids_out, double *info)
{
  try {
    if (!c || !path) AHF_FAIL("null argument");
    CUDA_CHECK(cudaSetDevice(c->dev)); ahf::g_pool_stream = c->stream;
    c->stage_reset();
    ahf::ingest_gadget(c, path, posscale > 0 ? posscale : 1.0, weightscale > 0 ? weightscale : 1.0, ids_out, info);
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

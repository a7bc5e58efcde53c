// comm.cu -- transports behind ahf::Comm (comm.cuh): NCCL (one process per GPU) and an in-process group of host threads.
#include "comm.cuh"
#include <condition_variable>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>          // types and prototypes only: the library itself is loaded with dlopen

namespace ahf {

namespace {

__global__ void k_add_u32(uint32_t *__restrict__ acc, const uint32_t *__restrict__ x, size_t n)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) acc[i] += x[i];
}

struct CollTimer {           // CUDA events on the context's stream around one device collective
  ahfgpu_ctx *c; Comm *m; cudaEvent_t a = nullptr, b = nullptr; size_t bytes;
  CollTimer(ahfgpu_ctx *ctx, Comm *cm, size_t nbytes) : c(ctx), m(cm), bytes(nbytes)
  {
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    CUDA_CHECK(cudaEventRecord(a, c->stream));
  }
  void done()
  {
    CUDA_CHECK(cudaEventRecord(b, c->stream));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
    m->coll_ms += ms; m->coll_calls++; m->coll_bytes += (int64_t)bytes;
    cudaEventDestroy(a); cudaEventDestroy(b); a = b = nullptr;
  }
  ~CollTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

// ------------------------------------------------------------------------------------------------
// in-process group: one host thread per context; collectives = barrier, copies from the peers' buffers, barrier
// ------------------------------------------------------------------------------------------------
struct LocalGroup {
  int n = 0;
  std::mutex mu; std::condition_variable cv; int arrived = 0; uint64_t gen = 0;
  bool aborted = false;            // a rank failed: the others must not wait for it for ever
  struct Slot { const void *a = nullptr; const void *b = nullptr; const void *c = nullptr; };
  std::vector<Slot> slot;
  void barrier()
  {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) AHF_FAIL("local group aborted (another rank failed)");
    const uint64_t g = gen;
    if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g || aborted; });
    if (aborted) AHF_FAIL("local group aborted (another rank failed)");
  }
  void abort()
  {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
};

struct LocalComm : Comm {
  LocalGroup *G;
  LocalComm(int r, LocalGroup *g) : G(g) { rank = r; nranks = g->n; }
  const char *kind() const override { return "local (host threads of one process)"; }
  void allgather_host(ahfgpu_ctx *c, const void *send, void *recv, size_t bytes) override
  {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    G->slot[rank].a = send;
    G->barrier();
    for (int p = 0; p < nranks; p++) memcpy((char *)recv + (size_t)p * bytes, G->slot[p].a, bytes);
    G->barrier();
  }
  void allreduce_sum_u32(ahfgpu_ctx *c, uint32_t *buf, size_t n) override
  {
    CollTimer t(c, this, n * 4);
    uint32_t *tmp = static_cast<uint32_t *>(cache_alloc((n ? n : 1) * 4)), *stg = static_cast<uint32_t *>(cache_alloc((n ? n : 1) * 4));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    G->slot[rank].a = buf;
    G->barrier();
    CUDA_CHECK(cudaMemsetAsync(tmp, 0, n * 4, c->stream));
    for (int p = 0; p < nranks && n; p++) {           // rank order: every rank computes the same sum
      CUDA_CHECK(cudaMemcpyAsync(stg, G->slot[p].a, n * 4, cudaMemcpyDefault, c->stream));
      k_add_u32<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(tmp, stg, n);
      CUDA_CHECK(cudaGetLastError());
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    G->barrier();                                      // nobody reads a peer's buffer any more
    CUDA_CHECK(cudaMemcpyAsync(buf, tmp, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    cache_free(tmp); cache_free(stg);
    t.done();
  }
  void alltoallv(ahfgpu_ctx *c, const void *const *sendptr, const size_t *sendbytes, void *const *recvptr, const size_t *recvbytes) override
  {
    size_t tot = 0; for (int p = 0; p < nranks; p++) tot += recvbytes[p];
    CollTimer t(c, this, tot);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    G->slot[rank].a = sendptr; G->slot[rank].b = sendbytes;
    G->barrier();
    for (int p = 0; p < nranks; p++) {
      const void *const *sp = static_cast<const void *const *>(G->slot[p].a);
      const size_t      *sb = static_cast<const size_t *>(G->slot[p].b);
      if (sb[rank] != recvbytes[p]) AHF_FAIL("alltoallv: send and receive sizes disagree");
      if (recvbytes[p]) CUDA_CHECK(cudaMemcpyAsync(recvptr[p], sp[rank], recvbytes[p], cudaMemcpyDefault, c->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    G->barrier();
    t.done();
  }
  void allgatherv(ahfgpu_ctx *c, const void *send, void *recv, const size_t *bytes, const size_t *off) override
  {
    size_t tot = 0; for (int p = 0; p < nranks; p++) tot += bytes[p];
    CollTimer t(c, this, tot);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    G->slot[rank].a = send;
    G->barrier();
    for (int p = 0; p < nranks; p++)
      if (bytes[p]) CUDA_CHECK(cudaMemcpyAsync((char *)recv + off[p], G->slot[p].a, bytes[p], cudaMemcpyDefault, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    G->barrier();
    t.done();
  }
};

// ------------------------------------------------------------------------------------------------
// NCCL, loaded at run time
// ------------------------------------------------------------------------------------------------
struct NcclApi {
  void *h = nullptr;
  decltype(&ncclGetUniqueId)   GetUniqueId = nullptr;
  decltype(&ncclCommInitRank)  CommInitRank = nullptr;
  decltype(&ncclCommDestroy)   CommDestroy = nullptr;
  decltype(&ncclAllReduce)     AllReduce = nullptr;
  decltype(&ncclAllGather)     AllGather = nullptr;
  decltype(&ncclBroadcast)     Broadcast = nullptr;
  decltype(&ncclSend)          Send = nullptr;
  decltype(&ncclRecv)          Recv = nullptr;
  decltype(&ncclGroupStart)    GroupStart = nullptr;
  decltype(&ncclGroupEnd)      GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi &nccl()
{
  static NcclApi A;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (A.h) return A;
  // a process that already carries an NCCL (torch bundles one under the same SONAME) gets that one back
  A.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!A.h) A.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!A.h) AHF_FAIL(std::string("libnccl.so.2 not found: ") + dlerror());
#define AHF_NCCL_SYM(field, name)                                                  \
  *(void **)(&A.field) = dlsym(A.h, name);                                          \
  if (!A.field) AHF_FAIL(std::string("libnccl: missing symbol ") + name)
  AHF_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); AHF_NCCL_SYM(CommInitRank, "ncclCommInitRank"); AHF_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  AHF_NCCL_SYM(AllReduce, "ncclAllReduce"); AHF_NCCL_SYM(AllGather, "ncclAllGather"); AHF_NCCL_SYM(Broadcast, "ncclBroadcast");
  AHF_NCCL_SYM(Send, "ncclSend"); AHF_NCCL_SYM(Recv, "ncclRecv"); AHF_NCCL_SYM(GroupStart, "ncclGroupStart"); AHF_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  AHF_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef AHF_NCCL_SYM
  return A;
}
#define NCCL_CHECK(expr)                                                                                          \
  do {                                                                                                            \
    ncclResult_t r__ = (expr);                                                                                    \
    if (r__ != ncclSuccess) ::ahf::fail(__FILE__, __LINE__, std::string(#expr ": ") + nccl().GetErrorString(r__)); \
  } while (0)

struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  void      *d_small = nullptr;            // device staging of the small host all-gathers
  size_t     small_cap = 0;
  NcclComm(int r, int n, const void *id128, int device)
  {
    rank = r; nranks = n;
    NcclApi &A = nccl();
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(&id, id128, sizeof(id));
    CUDA_CHECK(cudaSetDevice(device));
    NCCL_CHECK(A.CommInitRank(&comm, n, id, r));
  }
  ~NcclComm() override
  {
    if (comm) nccl().CommDestroy(comm);
    if (d_small) cudaFree(d_small);
  }
  const char *kind() const override { return "nccl"; }
  void allgather_host(ahfgpu_ctx *c, const void *send, void *recv, size_t bytes) override
  {
    const size_t need = bytes * (size_t)(nranks + 1);
    if (need > small_cap) {
      if (d_small) CUDA_CHECK(cudaFree(d_small));
      d_small = nullptr; small_cap = 0;
      const size_t cap = need < 4096 ? 4096 : need;
      CUDA_CHECK(cudaMalloc(&d_small, cap)); small_cap = cap;
    }
    char *ds = static_cast<char *>(d_small), *dr = ds + bytes;
    CUDA_CHECK(cudaMemcpyAsync(ds, send, bytes, cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(nccl().AllGather(ds, dr, bytes, ncclChar, comm, c->stream));
    read_back(c, recv, dr, bytes * (size_t)nranks);
  }
  void allreduce_sum_u32(ahfgpu_ctx *c, uint32_t *buf, size_t n) override
  {
    CollTimer t(c, this, n * 4);
    if (n) NCCL_CHECK(nccl().AllReduce(buf, buf, n, ncclUint32, ncclSum, comm, c->stream));
    t.done();
  }
  void alltoallv(ahfgpu_ctx *c, const void *const *sendptr, const size_t *sendbytes, void *const *recvptr, const size_t *recvbytes) override
  {
    size_t tot = 0; for (int p = 0; p < nranks; p++) tot += recvbytes[p];
    CollTimer t(c, this, tot);
    NcclApi &A = nccl();
    if (recvbytes[rank]) CUDA_CHECK(cudaMemcpyAsync(recvptr[rank], sendptr[rank], recvbytes[rank], cudaMemcpyDeviceToDevice, c->stream));
    NCCL_CHECK(A.GroupStart());
    for (int p = 0; p < nranks; p++) {
      if (p == rank) continue;
      if (sendbytes[p]) NCCL_CHECK(A.Send(sendptr[p], sendbytes[p], ncclChar, p, comm, c->stream));
      if (recvbytes[p]) NCCL_CHECK(A.Recv(recvptr[p], recvbytes[p], ncclChar, p, comm, c->stream));
    }
    NCCL_CHECK(A.GroupEnd());
    t.done();
  }
  void allgatherv(ahfgpu_ctx *c, const void *send, void *recv, const size_t *bytes, const size_t *off) override
  {
    size_t tot = 0; for (int p = 0; p < nranks; p++) tot += bytes[p];
    CollTimer t(c, this, tot);
    NcclApi &A = nccl();
    NCCL_CHECK(A.GroupStart());
    for (int p = 0; p < nranks; p++) {
      if (!bytes[p]) continue;
      char *dst = (char *)recv + off[p];
      NCCL_CHECK(A.Broadcast(p == rank ? send : (const void *)dst, dst, bytes[p], ncclChar, p, comm, c->stream));
    }
    NCCL_CHECK(A.GroupEnd());
    t.done();
  }
};

}  // namespace

void comm_nccl_unique_id(void *id128)
{
  ncclUniqueId id;
  NCCL_CHECK(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
}
Comm *comm_create_nccl(int rank, int nranks, const void *id128, int device) { return new NcclComm(rank, nranks, id128, device); }
void *comm_local_group_create(int nranks)
{
  LocalGroup *g = new LocalGroup();
  g->n = nranks; g->slot.resize(nranks);
  return g;
}
void comm_local_group_destroy(void *group) { delete static_cast<LocalGroup *>(group); }
void comm_local_group_abort(void *group) { if (group) static_cast<LocalGroup *>(group)->abort(); }
Comm *comm_create_local(int rank, void *group)
{
  LocalGroup *g = static_cast<LocalGroup *>(group);
  if (!g || rank < 0 || rank >= g->n) AHF_FAIL("bad local group / rank");
  return new LocalComm(rank, g);
}

}  // namespace ahf

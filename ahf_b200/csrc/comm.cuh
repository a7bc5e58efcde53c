// comm.cuh -- the few collectives the ONE-box-on-several-GPUs path needs (slab.cu, mesh.cu), behind one small interface with two
// transports:
//   * NcclComm  -- one process per GPU, NCCL over NVLink / NVSwitch (libnccl.so.2 is loaded at run time with dlopen, so a single-GPU
//                  user of libahfgpu.so needs no NCCL at all);
//   * LocalComm -- several contexts of ONE process, each driven by its own host thread (the contexts may share a device):
//                  collectives are host barriers + device-to-device copies.  It exists so that the slab decomposition can be tested
//                  against the single-GPU result with 2, 4 or 8 "ranks" on a box with one GPU.
// The model is the reference's MPI mode: comm_dist_part (src/comm.c:104-316: particles to the owner of their SFC key range),
// loadbalance_update / local_equalpart (src/libutility/loadbalance.c:206,383: histogram per SFC cell, all-reduced, equal-particle
// ranges) and the boundary duplication (src/comm.c:324ff, src/libsfc/sfc_boundary.c:100-144).
#pragma once
#include "common.cuh"

namespace ahf {

struct Comm {
  int rank = 0, nranks = 1;
  virtual ~Comm() {}
  virtual const char *kind() const = 0;
  // small host blobs: every rank contributes `bytes`, all receive nranks * bytes in rank order (synchronises c->stream)
  virtual void allgather_host(ahfgpu_ctx *c, const void *send, void *recv, size_t bytes) = 0;
  // device, in place: buf[i] = sum over ranks
  virtual void allreduce_sum_u32(ahfgpu_ctx *c, uint32_t *buf, size_t n) = 0;
  // device, personalised exchange: sendptr[p] (sendbytes[p]) goes to rank p, recvptr[p] (recvbytes[p]) comes from rank p
  virtual void alltoallv(ahfgpu_ctx *c, const void *const *sendptr, const size_t *sendbytes, void *const *recvptr, const size_t *recvbytes) = 0;
  // device: recv + off[p] receives rank p's `send` (bytes[p]); every rank knows all sizes
  virtual void allgatherv(ahfgpu_ctx *c, const void *send, void *recv, const size_t *bytes, const size_t *off) = 0;
  // milliseconds spent inside the device collectives since the last reset (CUDA events on c->stream), and their number
  double  coll_ms = 0.0;
  int64_t coll_calls = 0, coll_bytes = 0;
};

// decomposition of one box over the ranks (slab.cu)
struct Slab {
  int      bd = 0;                 // bits per dimension of the decomposition cells ("blocks"): LevelDomainDecomp of the reference
  int      T = 0;                  // ghost shell thickness in blocks
  uint64_t n_total = 0;            // particles of the whole box
  std::vector<uint64_t> split;     // [nranks + 1] first Hilbert block of every rank (block = key >> 3 (21 - bd))
  uint8_t *own3 = nullptr;         // device [2^bd]^3 (z, y, x): owner rank of every block
  uint64_t own_lo = 0, own_hi = 0; // the rank's own particles are [own_lo, own_hi) of the resident key-sorted set (ghosts around them)
  double   ghost_width = 0.0;      // box units
};

Comm *comm_create_nccl(int rank, int nranks, const void *id128, int device);
void  comm_nccl_unique_id(void *id128);
void *comm_local_group_create(int nranks);
void  comm_local_group_destroy(void *group);
void  comm_local_group_abort(void *group);
Comm *comm_create_local(int rank, void *group);

void slab_distribute(ahfgpu_ctx *c, uint64_t id_base, double ghost_width, int decomp_bits);
void slab_free(ahfgpu_ctx *c);

}  // namespace ahf

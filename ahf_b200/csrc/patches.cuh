// patches.cuh -- lock-free union-find primitives of the patch labelling (SURVEY 8f NEXT-1: the colouring sweep of
// src/libahf/ahf_gridinfo.c:236-577).  Host+device: tests/test_host_cpu.py drives the same functions on the CPU against the
// literal restatement of the sweep in oracle/ahf_oracle_mesh.c.
//
// What the reference's sequential sweep computes, stated as a graph problem: cells are visited in traversal order (z, y, x);
// a cell takes the colour of the already coloured face neighbours its search SEES (get_TSCnodes visibility, six directions)
// and, where it sees several colours, the larger ones are rewritten to the smallest.  So two cells end up with one colour
// exactly when they are connected through edges (c, n) with n visible from c and n EARLIER than c in traversal order
// (a neighbour later in the order is still uncoloured when c is visited; across a periodic face the "forward" neighbour is
// an earlier cell and does count).  The surviving colours, in creation order, are the level's isolated refinements
// 0, 1, ...: component k is the one whose FIRST cell comes k-th.  With every union hooking the larger root under the smaller
// one, the root of a component is its first cell, and the rank of that cell among the roots is the reference's index.
#pragma once
#include <cstdint>

#ifndef AHF_HD
#if defined(__CUDACC__)
#define AHF_HD __host__ __device__ __forceinline__
#else
#define AHF_HD inline
#endif
#endif

namespace ahf {

AHF_HD int32_t uf_load(const int32_t *parent, int32_t i)
{
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const volatile int32_t *>(parent + i);     // other threads hook and compress concurrently
#else
  return parent[i];
#endif
}

// root of i, with path halving.  parent[i] <= i always (hooks go from the larger to the smaller root), and a halving store
// replaces a parent by one of its ancestors, so concurrent finds and unions stay correct without atomics on this path.
AHF_HD int32_t uf_find(int32_t *parent, int32_t i)
{
  for (;;) {
    const int32_t p = uf_load(parent, i);
    if (p == i) return i;
    const int32_t gp = uf_load(parent, p);
#if defined(__CUDA_ARCH__)
    if (gp != p) *reinterpret_cast<volatile int32_t *>(parent + i) = gp;
#else
    if (gp != p) parent[i] = gp;
#endif
    i = p;
  }
}

// union by index: the larger root is hooked under the smaller one (compare-and-swap on the root's own slot; a lost race means
// somebody else hooked it first: look the roots up again)
AHF_HD void uf_unite(int32_t *parent, int32_t a, int32_t b)
{
  for (;;) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int32_t t = a; a = b; b = t; }
#if defined(__CUDA_ARCH__)
    if (atomicCAS(reinterpret_cast<int *>(parent + a), (int)a, (int)b) == (int)a) return;
#else
    if (parent[a] == a) { parent[a] = b; return; }
#endif
  }
}

}  // namespace ahf

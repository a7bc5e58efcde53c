"""Host-side work partitioning for several GPUs of one box.

* haloes are independent units: greedy longest-processing-time assignment by (expected) gathered particle count
  (the reference gives one OpenMP thread one halo, `schedule(dynamic)`, src/libahf/ahf_halos.c:504-510);
* particles are split into SFC-contiguous slabs with (nearly) equal counts -- the equal-particle key ranges of the
  reference's MPI load balancer (src/libutility/loadbalance.c:383).
No data-path collective is needed for either: every rank derives the same partition from the same inputs.
"""
from __future__ import annotations

import heapq

import numpy as np


def assign_halos_lpt(weights: np.ndarray, nranks: int) -> np.ndarray:
    """rank of every halo; deterministic (ties by halo index)."""
    weights = np.asarray(weights, dtype=np.float64)
    order = np.lexsort((np.arange(len(weights)), -weights))
    heap = [(0.0, r) for r in range(nranks)]
    heapq.heapify(heap)
    out = np.empty(len(weights), dtype=np.int32)
    for h in order:
        load, r = heapq.heappop(heap)
        out[h] = r
        heapq.heappush(heap, (load + weights[h], r))
    return out


def slab_bounds(n: int, nranks: int) -> np.ndarray:
    """offsets [nranks+1] into the key-sorted particle array: rank r owns [b[r], b[r+1])"""
    return (np.arange(nranks + 1, dtype=np.int64) * n) // nranks


def slab_key_ranges(keys_sorted: np.ndarray, nranks: int) -> np.ndarray:
    """first key of every slab (and one past the last): equal-particle Hilbert key ranges; equal keys never straddle"""
    b = slab_bounds(len(keys_sorted), nranks)
    lo = np.empty(nranks + 1, dtype=np.uint64)
    lo[0] = 0
    for r in range(1, nranks):
        lo[r] = keys_sorted[b[r]]
    lo[nranks] = np.uint64(1) << np.uint64(63)
    return lo

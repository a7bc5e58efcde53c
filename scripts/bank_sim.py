"""Offline model of the shared-memory atomic wavefronts of the domain TSC deposit (k_deposit_dom): for every 32-particle slice of
a Hilbert-sorted box, count per stencil term the serialisation a tile layout causes (max lanes per bank; lanes on the same
address serialise as well).  Used to choose the layout before spending GPU time.  python scripts/bank_sim.py [n1d]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ahf_b200 import synth
from oracle import oracle as O

n1d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
box = synth.make_box(n1d, seed=43)
keys = O.hilbert_keys(box.pos)
order = O.argsort_keys(keys)
keys = keys[order]
pos = box.pos[order]
L = n1d
logL = int(np.log2(L))
tbits = logL - 4
tile = (keys >> np.uint64(3 * (21 - tbits))).astype(np.int64)
cell = np.minimum((pos.astype(np.float64) * L).astype(np.int64), L - 1)
lx, ly, lz = (cell[:, 0] & 15) + 1, (cell[:, 1] & 15) + 1, (cell[:, 2] & 15) + 1   # +1: rim

# slices: 32 consecutive particles of a <=8192 chunk of a tile
starts = np.flatnonzero(np.r_[True, tile[1:] != tile[:-1]])
ends = np.r_[starts[1:], len(tile)]
rng = np.random.default_rng(0)
sel = rng.choice(len(starts), size=min(len(starts), 400), replace=False)
rows = []
for t in sel:
    s, e = starts[t], ends[t]
    for c0 in range(s, e, 8192):
        c1 = min(e, c0 + 8192)
        for q in range(c0, c1, 32):
            idx = np.arange(q, min(q + 32, c1))
            if len(idx) < 32:
                continue
            rows.append(idx)
rows = np.array(rows)
print("slices", rows.shape[0], "tiles", len(sel), "mean particles/tile", (ends - starts).mean())
X, Y, Z = lx[rows], ly[rows], lz[rows]
lane = np.arange(32)[None, :]


def wavefronts(addr_fn, name, terms=None):
    tot = 0
    n = 0
    same = (X.max(1) == X.min(1)) & (Y.max(1) == Y.min(1)) & (Z.max(1) == Z.min(1))   # grouped slices take the REDUX path
    keep = ~same
    for k in (-1, 0, 1):
        for j in (-1, 0, 1):
            for a in (-1, 0, 1):
                ad = addr_fn(X[keep] + a, Y[keep] + j, Z[keep] + k, lane)
                bank = ad & 31
                cnt = np.zeros((ad.shape[0], 32), dtype=np.int32)
                np.add.at(cnt, (np.arange(ad.shape[0])[:, None].repeat(32, 1), bank), 1)
                tot += cnt.max(1).sum()
                n += ad.shape[0]
    print(f"{name:40s} wavefronts/ATOMS = {tot / n:.3f}   (grouped slices {same.mean():.3%})")
    return tot / n


H = 18
wavefronts(lambda x, y, z, l: (z * H + y) * H + x, "linear 18x18, one copy")
wavefronts(lambda x, y, z, l: (z * H + y) * H + x + (l & 1) * (H * H * H + 8), "linear 18x18, two copies (current)")
for rs, ps in ((20, 368), (19, 19 * 18 + 6), (18, 18 * 18 + 4), (20, 20 * 18 + 8), (17 + 4, 21 * 18 + 6)):
    wavefronts(lambda x, y, z, l: z * ps + y * rs + x, f"linear rs={rs} ps={ps} one copy")
    wavefronts(lambda x, y, z, l: z * ps + y * rs + x + (l & 1) * 16, f"linear rs={rs} ps={ps} two copies +16")


def swz(x, y, z, l, copies=1):
    X4, Y4, Z4 = x >> 2, y >> 2, z >> 2
    x0, x1, y0, y1, z0, z1 = x & 1, (x >> 1) & 1, y & 1, (y >> 1) & 1, z & 1, (z >> 1) & 1
    low = x0 | (y0 << 1) | (z0 << 2) | ((x1 ^ y1) << 3) | ((y1 ^ z1) << 4) | (y1 << 5)
    a = ((Z4 * 5 + Y4) * 5 + X4) * 64 + low
    if copies == 2:
        a = a ^ ((l & 1) << 4)
    return a


wavefronts(lambda x, y, z, l: swz(x, y, z, l, 1), "4x4x4 block xor swizzle, one copy")
wavefronts(lambda x, y, z, l: swz(x, y, z, l, 2), "4x4x4 block xor swizzle, lane-parity xor 16")


# ---- model of a run-aggregating domain kernel: thread = K consecutive particles, registers flushed when the cell changes
def runs_model(K, threads=512):
    tot_iter = tot_flush_exec = tot_flush_lanes = tot_wf = 0
    npart = 0
    for t in sel:
        s, e = starts[t], ends[t]
        for c0 in range(s, e, 8192):
            c1 = min(e, c0 + 8192)
            n = c1 - c0
            wid = ((lz[c0:c1] * H + ly[c0:c1]) * H + lx[c0:c1]).astype(np.int64)
            npart += n
            # segments of K per thread; warp = 32 consecutive threads
            nseg = (n + K - 1) // K
            pad = nseg * K - n
            w = np.concatenate([wid, np.full(pad, -1)]).reshape(nseg, K)
            nwarp = (nseg + 31) // 32
            w = np.concatenate([w, np.full((nwarp * 32 - nseg, K), -1)]).reshape(nwarp, 32, K)
            prev = np.full((nwarp, 32), -1)
            for i in range(K + 1):
                cur = w[:, :, i] if i < K else np.full((nwarp, 32), -1)
                flush = (prev >= 0) & (cur != prev)                 # lanes that flush prev before taking particle i
                anyf = flush.any(1)
                tot_flush_exec += int(anyf.sum())
                tot_flush_lanes += int(flush.sum())
                # wavefronts of one of the 27 flush atomics (term 0,0,0): max lanes per bank among flushing lanes
                if anyf.any():
                    bank = np.where(flush, prev & 31, -1)
                    cnt = np.zeros((nwarp, 33), dtype=np.int32)
                    np.add.at(cnt, (np.arange(nwarp)[:, None].repeat(32, 1), bank + 1), 1)
                    tot_wf += int(cnt[:, 1:].max(1).sum())
                if i < K:
                    tot_iter += int((cur >= 0).any(1).sum())
                    prev = np.where(cur >= 0, cur, prev)
    per32 = npart / 32.0
    print(f"K={K}: per 32 particles: accumulate iterations {tot_iter / per32:.2f}, flush executions {tot_flush_exec / per32:.2f}, "
          f"flushing lanes per execution {tot_flush_lanes / max(tot_flush_exec, 1):.1f}, wavefronts per flush atomic {tot_wf / max(tot_flush_exec, 1):.2f}")


for K in (1, 4, 8, 16):
    runs_model(K)

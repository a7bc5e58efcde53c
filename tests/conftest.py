import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["frag16", "edge32", "plain32", "species32"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """One fixture written by tests/golden/make_golden.py from the unmodified reference."""

    def __init__(self, name):
        self.name = name
        self.d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.n1d = int(self.d["n1d"]); self.nlev = int(self.d["nlev"])
        self.nper_dom = float(self.d["nper_dom"]); self.nper_ref = float(self.d["nper_ref"])
        self.keys = self.d["keys"]; self.pos = self.d["pos"]; self.mom = self.d["mom"]; self.ids = self.d["ids"]
        self.glob = self.d["halo_glob"]; self.hs = self.d["halo_s"]
        # multi-species fixtures (the reference's -DMULTIMASS -DGAS_PARTICLES build) carry weights and thermal energies
        self.weight = self.d["weight"] if "weight" in self.d.files else None
        self.u = self.d["u"] if "u" in self.d.files else None

    def input_order(self):
        """positions / momenta in snapshot-file order (ids are 0..N-1 in file order)"""
        n = self.pos.shape[0]
        pos = np.empty_like(self.pos); mom = np.empty_like(self.mom)
        pos[self.ids] = self.pos; mom[self.ids] = self.mom
        assert len(np.unique(self.ids)) == n
        return pos, mom

    def level(self, l):
        p = "L%d_" % l
        return {k: self.d[p + k] for k in ("l1dim", "critdens", "masstopartdens", "x", "y", "z", "dens", "runflags", "cnt",
                                           "plist", "cnt_final", "plist_final")}

    def patches(self):
        """(min_ref, {level: (iso[ncell], periodic[niso, 3])}) from tests/golden/patches.npz: the outcome of the reference's patch
        colouring (ahf_gridinfo) for this case, written by tests/golden/make_golden.py patches"""
        P = np.load(os.path.join(GOLDEN_DIR, "patches.npz"))
        m = int(P[self.name + "_min_ref"]); n = int(P[self.name + "_nlev"])
        return m, {l: (P["%s_L%d_iso" % (self.name, l)].astype(np.int32), P["%s_L%d_per" % (self.name, l)]) for l in range(m, m + n)}

    def gridtree(self):
        """{level: dict(centre, close, nodes, parts, daughter, sub_off, sub)} from tests/golden/gridtree.npz (the reference's own
        .AHF_gridtree, -DAHFgridtreefile build; default-build cases only) or None"""
        P = np.load(os.path.join(GOLDEN_DIR, "gridtree.npz"))
        if self.name + "_min_ref" not in P.files:
            return None
        m = int(P[self.name + "_min_ref"]); n = int(P[self.name + "_nlev"])
        return {l: {k: P["%s_L%d_%s" % (self.name, l, k)] for k in ("centre", "close", "nodes", "parts", "daughter", "sub_off", "sub")}
                for l in range(m, m + n)}

    def members(self, i):
        return self.d["halo_members"][self.d["halo_moff"][i]:self.d["halo_moff"][i + 1]].astype(np.int64)

    def species(self, i):
        return self.d["halo_species"][i] if "halo_species" in self.d.files else None

    def prof_species(self, i):
        a, b = int(self.d["halo_poff"][i]), int(self.d["halo_poff"][i + 1])
        if b == a or "halo_prof_species" not in self.d.files:
            return None
        return self.d["halo_prof_species"][a * 3:b * 3].reshape(3, b - a)

    def prof(self, i):
        a, b = int(self.d["halo_poff"][i]), int(self.d["halo_poff"][i + 1])
        if b == a:
            return None
        return self.d["halo_prof"][a * 25:b * 25].reshape(25, b - a)


@pytest.fixture(scope="session", params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


def lin(x, y, z, L):
    L = np.int64(L)
    return (z.astype(np.int64) * L + y.astype(np.int64)) * L + x.astype(np.int64)

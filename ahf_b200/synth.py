"""Synthetic inputs for the AHF hot path (BASELINE.json `configs`).

A box holds (i) a jittered lattice (Zel'dovich-style Gaussian displacements, sigma ~ 0.3 cell, small
Gaussian velocities) and (ii) Plummer spheres (about 30 % of the particles, masses log-uniform over
about 3 decades, isotropic Gaussian velocities with the Plummer 1-D dispersion).  Everything is
equal-mass dark matter (GADGET type 1), z = 0.

Two views of the same particles are produced:

* internal units, exactly what the reference holds in `struct particle` after
  `io_gadget_scale_particles` (reference src/libio/io_gadget.c:919-925): `pos` = x / boxsize as
  float32 in [0,1), `mom` = float32(v) * float32(1/(boxsize*100)) at a = 1;
* a GADGET-1 single file (reference src/libio/io_gadget_header_def.h:41-66, io_gadget.c:426-568)
  plus an `AHF.input` so that the compiled reference (oracle/_ref/ahf_ref) can be run on the same data.

The box size is a power of two (Mpc/h) so that x / boxsize is exact in float32 and both views agree
bit for bit.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

GRAV = 4.3006485e-9      # reference src/param.h:44  [Mpc km^2 / (Msun s^2)]
RHOC0 = 2.7755397e11     # reference src/param.h:43  [h^2 Msun / Mpc^3]
H0 = 100.0               # reference src/param.h:42


@dataclass
class Box:
    """One synthetic snapshot in both unit systems."""
    n1d: int
    boxsize: float           # Mpc/h
    omega0: float
    lambda0: float
    pmass: float             # Msun/h per particle
    pos: np.ndarray          # (N,3) float32, box units [0,1)
    mom: np.ndarray          # (N,3) float32, internal velocity units
    ids: np.ndarray          # (N,) uint64
    vel_kms: np.ndarray      # (N,3) float32 as written to the GADGET file
    clump_centres: np.ndarray  # (nc,3) float64 box units
    clump_npart: np.ndarray    # (nc,) int64
    clump_scale: np.ndarray    # (nc,) float64 Plummer a in box units

    @property
    def npart(self) -> int:
        return int(self.pos.shape[0])


def default_boxsize(n1d: int) -> float:
    """0.5 Mpc/h per mean inter-particle spacing, rounded to a power of two."""
    return float(2 ** int(round(np.log2(0.5 * n1d))))


def make_box(n1d: int, seed: int = 42, clump_frac: float = 0.3, n_clumps: int | None = None,
             boxsize: float | None = None, omega0: float = 0.3, lambda0: float = 0.7,
             sigma_cell: float = 0.3, mass_decades: float = 3.0, centres_box: np.ndarray | None = None) -> Box:
    rng = np.random.default_rng(seed)
    box = default_boxsize(n1d) if boxsize is None else float(boxsize)
    ntot = n1d ** 3
    if n_clumps is None:
        n_clumps = max(1, int(round(20 * (n1d / 128.0) ** 3)))
    pmass = omega0 * RHOC0 * box ** 3 / ntot

    # ---- clump membership: log-uniform masses over `mass_decades`, total = clump_frac * ntot
    n_cl_target = int(clump_frac * ntot)
    w = 10.0 ** (rng.uniform(0.0, mass_decades, size=n_clumps))
    cn = np.maximum(30, np.floor(w / w.sum() * n_cl_target)).astype(np.int64)
    n_cl = int(cn.sum())
    n_lat = ntot - n_cl
    if n_lat <= 0:
        raise ValueError("clump fraction too large")

    # ---- lattice part: a random subset of lattice sites keeps the total at n1d^3
    sites = rng.choice(ntot, size=n_lat, replace=False) if n_lat < ntot else np.arange(ntot)
    sites.sort()
    iz = sites // (n1d * n1d)
    iy = (sites // n1d) % n1d
    ix = sites % n1d
    cell = box / n1d
    lat = (np.stack([ix, iy, iz], axis=1).astype(np.float64) + 0.5) * cell
    lat += rng.normal(0.0, sigma_cell * cell, size=lat.shape)
    vlat = rng.normal(0.0, 50.0, size=lat.shape)

    # ---- Plummer clumps
    centres = rng.uniform(0.0, box, size=(n_clumps, 3))
    if centres_box is not None:                       # explicit centres (box units) for the first clumps
        cb = np.asarray(centres_box, dtype=np.float64).reshape(-1, 3)
        centres[:cb.shape[0]] = cb[:n_clumps] * box
    mass = cn * pmass
    a_pl = np.clip(0.1 * (mass / 1e14) ** (1.0 / 3.0), 0.05, 0.15)          # Mpc/h
    cpos = np.empty((n_cl, 3))
    cvel = np.empty((n_cl, 3))
    o = 0
    for c in range(n_clumps):
        m = int(cn[c])
        u = rng.uniform(1e-9, 1.0 - 1e-6, size=m)
        r = a_pl[c] / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        r = np.minimum(r, 20.0 * a_pl[c])
        d = rng.normal(size=(m, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        cpos[o:o + m] = centres[c] + d * r[:, None]
        sig2 = GRAV * mass[c] / (6.0 * a_pl[c]) / np.sqrt(1.0 + (r / a_pl[c]) ** 2)
        bulk = rng.normal(0.0, 200.0, size=3)
        cvel[o:o + m] = bulk + rng.normal(size=(m, 3)) * np.sqrt(sig2)[:, None]
        o += m

    x = np.concatenate([lat, cpos], axis=0)
    v = np.concatenate([vlat, cvel], axis=0)
    x = np.mod(x, box)
    x32 = x.astype(np.float32)
    # float32(box) must stay below box after rounding: pull the rare x32 == box down
    x32 = np.where(x32 >= np.float32(box), np.nextafter(np.float32(box), np.float32(0.0)), x32).astype(np.float32)
    v32 = v.astype(np.float32)

    pos = (x32 * np.float32(1.0 / box)).astype(np.float32)               # exact: box is 2^k
    scale_mom = 1.0 / (box * 1.0 * 100.0)                                  # a = 1 (io_gadget.c:921-923)
    mom = (v32 * np.float32(scale_mom)).astype(np.float32)
    ids = np.arange(ntot, dtype=np.uint64)
    return Box(n1d=n1d, boxsize=box, omega0=omega0, lambda0=lambda0, pmass=pmass, pos=pos, mom=mom, ids=ids,
               vel_kms=v32, clump_centres=centres / box, clump_npart=cn, clump_scale=a_pl / box)


def write_gadget1(box: Box, path: str, big_endian: bool = False, version: int = 1, pos_offset: float = 0.0) -> None:
    """GADGET-1 (default little-endian; big_endian / version=2 framing for the reader tests), all particles of type 1 with
    massarr[1] > 0 (no MASS block).  pos_offset shifts all positions (file units; negative coordinates make the reader shift them back)."""
    if big_endian or version != 1 or pos_offset != 0.0:
        return _write_gadget_variant(box, path, big_endian, version, pos_offset)
    n = box.npart
    if n >= 2 ** 31 // 12:
        raise ValueError("single GADGET file limited by 32-bit block lengths; split the snapshot")
    hdr = bytearray(256)
    np_ = [0, n, 0, 0, 0, 0]
    massarr = [0.0, box.pmass / 1e10, 0.0, 0.0, 0.0, 0.0]                  # GADGET_MUNIT = 1e10 Msun/h
    struct.pack_into("<6i", hdr, 0, *np_)
    struct.pack_into("<6d", hdr, 24, *massarr)
    struct.pack_into("<2d", hdr, 72, 1.0, 0.0)                             # expansion, redshift
    struct.pack_into("<2i", hdr, 88, 0, 0)
    struct.pack_into("<6I", hdr, 96, *np_)
    struct.pack_into("<2i", hdr, 120, 0, 1)                                # flagcooling, numfiles
    struct.pack_into("<4d", hdr, 128, box.boxsize, box.omega0, box.lambda0, 0.7)

    def block(f, payload: bytes):
        f.write(struct.pack("<I", len(payload)))
        f.write(payload)
        f.write(struct.pack("<I", len(payload)))

    x = (box.pos.astype(np.float32) * np.float32(box.boxsize)).astype("<f4")  # exact inverse of the scaling
    with open(path, "wb") as f:
        block(f, bytes(hdr))
        block(f, x.tobytes())
        block(f, box.vel_kms.astype("<f4").tobytes())
        block(f, box.ids.astype("<u4").tobytes())


def _write_gadget_variant(box: Box, path: str, big_endian: bool, version: int, pos_offset: float) -> None:
    e = ">" if big_endian else "<"
    n = box.npart
    hdr = bytearray(256)
    np_ = [0, n, 0, 0, 0, 0]
    struct.pack_into(e + "6i", hdr, 0, *np_)
    struct.pack_into(e + "6d", hdr, 24, 0.0, box.pmass / 1e10, 0.0, 0.0, 0.0, 0.0)
    struct.pack_into(e + "2d", hdr, 72, 1.0, 0.0)
    struct.pack_into(e + "2i", hdr, 88, 0, 0)
    struct.pack_into(e + "6I", hdr, 96, *np_)
    struct.pack_into(e + "2i", hdr, 120, 0, 1)
    struct.pack_into(e + "4d", hdr, 128, box.boxsize, box.omega0, box.lambda0, 0.7)

    def block(f, name: bytes, payload: bytes):
        if version == 2:
            f.write(struct.pack(e + "I", 8)); f.write(name); f.write(struct.pack(e + "I", len(payload) + 8)); f.write(struct.pack(e + "I", 8))
        f.write(struct.pack(e + "I", len(payload))); f.write(payload); f.write(struct.pack(e + "I", len(payload)))

    x = (box.pos.astype(np.float32) * np.float32(box.boxsize) + np.float32(pos_offset)).astype(e + "f4")
    with open(path, "wb") as f:
        block(f, b"HEAD", bytes(hdr))
        block(f, b"POS ", x.tobytes())
        block(f, b"VEL ", box.vel_kms.astype(e + "f4").tobytes())
        block(f, b"ID  ", box.ids.astype(e + "u4").tobytes())


def write_ahf_input(path: str, ic_filename: str, prefix: str, lgrid_domain: int, *, lgrid_max: int = 16777216,
                    nper_dom: float = 2.0, nper_ref: float = 2.5, vesc_tune: float = 1.5, nmin: int = 20,
                    rho_vir: int = 0, dvir: float = 200.0, max_gather_rad: float = 3.0) -> None:
    """AHF.input with the AHF.input-example settings (reference AHF.input-example:16-41)."""
    with open(path, "w") as f:
        f.write("[AHF]\n")
        f.write(f"ic_filename       = {ic_filename}\n")
        f.write("ic_filetype       = 60\n")
        f.write(f"outfile_prefix    = {prefix}\n")
        f.write(f"LgridDomain       = {lgrid_domain}\n")
        f.write(f"LgridMax          = {lgrid_max}\n")
        f.write(f"NperDomCell       = {nper_dom}\n")
        f.write(f"NperRefCell       = {nper_ref}\n")
        f.write(f"VescTune          = {vesc_tune}\n")
        f.write(f"NminPerHalo       = {nmin}\n")
        f.write(f"RhoVir            = {rho_vir}\n")
        f.write(f"Dvir              = {dvir}\n")
        f.write(f"MaxGatherRad      = {max_gather_rad}\n")
        f.write("LevelDomainDecomp = 6\nNcpuReading       = 1\n\n")
        f.write("[GADGET]\nGADGET_LUNIT      = 1.\nGADGET_MUNIT      = 1e10\n")


def write_reference_case(box: Box, workdir: str, lgrid_domain: int | None = None, **kw) -> str:
    """Write snapshot + AHF.input into `workdir`; returns the AHF.input path."""
    os.makedirs(workdir, exist_ok=True)
    snap = os.path.join(workdir, "snap.gadget")
    write_gadget1(box, snap)
    inp = os.path.join(workdir, "AHF.input")
    write_ahf_input(inp, snap, os.path.join(workdir, "ref"), lgrid_domain or box.n1d, **kw)
    return inp


def halo_seeds(box: Box, max_gather_rad_mpc: float = 3.0):
    """Halo seeds (centre, gathering radius, seed particle count) for the per-halo pass, standing in for what the
    reference's tree stage hands to ahf_halos_sfc_constructHalo: centres = the generator's clump centres,
    gatherRad = half the periodic distance to the nearest clump with more particles, at most
    min(MaxGatherRad/boxsize, 1/4); the richest clump gets the maximum (reference src/libahf/ahf_halos.c:2987-3051)."""
    c = np.mod(box.clump_centres, 1.0)
    npart = box.clump_npart.astype(np.int64)
    nh = len(npart)
    rmax = min(max_gather_rad_mpc / box.boxsize, 0.25)
    order = np.argsort(-npart, kind="stable")
    rad = np.full(nh, rmax)
    cs = c[order]
    for rank in range(1, nh):                      # clumps with more particles are the ones before `rank`
        d = np.abs(cs[:rank] - cs[rank])
        d = np.where(d > 0.5, 1.0 - d, d)
        dist = np.sqrt((d * d).sum(axis=1)).min()
        rad[order[rank]] = min(max(0.5 * dist, 4.0 * box.clump_scale[order[rank]]), rmax)
    return c.astype(np.float64), rad.astype(np.float64), npart


def make_box_slice(n1d: int, rank: int, world: int, seed: int = 42, clump_frac: float = 0.3, n_clumps: int | None = None,
                   boxsize: float | None = None, omega0: float = 0.3, sigma_cell: float = 0.3, mass_decades: float = 3.0):
    """The share of rank `rank` (of `world`) of a box of about n1d^3 particles, generated independently per rank so that very
    large boxes (512^3, 1024^3) never exist in one process: a contiguous z-slab of the jittered lattice (thinned by
    1 - clump_frac) plus every `world`-th Plummer clump.  Same statistics as make_box, not the same realisation.
    Returns (pos float32 (n,3) box units, mom float32 (n,3), clump centres/scale/npart of ALL clumps, boxsize, pmass)."""
    box = default_boxsize(n1d) if boxsize is None else float(boxsize)
    ntot = n1d ** 3
    if n_clumps is None:
        n_clumps = max(1, int(round(20 * (n1d / 128.0) ** 3)))
    pmass = omega0 * RHOC0 * box ** 3 / ntot
    rng0 = np.random.default_rng(seed)                      # global quantities: identical on every rank
    w = 10.0 ** (rng0.uniform(0.0, mass_decades, size=n_clumps))
    cn = np.maximum(30, np.floor(w / w.sum() * int(clump_frac * ntot))).astype(np.int64)
    centres = rng0.uniform(0.0, box, size=(n_clumps, 3))
    mass = cn * pmass
    a_pl = np.clip(0.1 * (mass / 1e14) ** (1.0 / 3.0), 0.05, 0.15)
    keep = 1.0 - cn.sum() / ntot
    rng = np.random.default_rng([seed, 1000 + rank])
    z0, z1 = (rank * n1d) // world, ((rank + 1) * n1d) // world
    cell = box / n1d
    parts_x, parts_v = [], []
    for iz in range(z0, z1):                                # plane by plane keeps the temporaries small
        m = rng.random(n1d * n1d) < keep
        idx = np.nonzero(m)[0]
        lat = np.empty((len(idx), 3))
        lat[:, 0] = (idx % n1d + 0.5) * cell; lat[:, 1] = (idx // n1d + 0.5) * cell; lat[:, 2] = (iz + 0.5) * cell
        lat += rng.normal(0.0, sigma_cell * cell, size=lat.shape)
        parts_x.append(lat.astype(np.float32)); parts_v.append(rng.normal(0.0, 50.0, size=lat.shape).astype(np.float32))
    for c in range(rank, n_clumps, world):
        mcl = int(cn[c])
        u = rng.uniform(1e-9, 1.0 - 1e-6, size=mcl)
        r = np.minimum(a_pl[c] / np.sqrt(u ** (-2.0 / 3.0) - 1.0), 20.0 * a_pl[c])
        d = rng.normal(size=(mcl, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        sig2 = GRAV * mass[c] / (6.0 * a_pl[c]) / np.sqrt(1.0 + (r / a_pl[c]) ** 2)
        parts_x.append((centres[c] + d * r[:, None]).astype(np.float32))
        parts_v.append((rng.normal(0.0, 200.0, size=3) + rng.normal(size=(mcl, 3)) * np.sqrt(sig2)[:, None]).astype(np.float32))
    x = np.mod(np.concatenate(parts_x), np.float32(box)).astype(np.float32)
    x = np.where(x >= np.float32(box), np.nextafter(np.float32(box), np.float32(0.0)), x).astype(np.float32)
    v = np.concatenate(parts_v)
    pos = (x * np.float32(1.0 / box)).astype(np.float32)
    mom = (v * np.float32(1.0 / (box * 100.0))).astype(np.float32)
    return pos, mom, dict(centres=centres / box, npart=cn, scale=a_pl / box), box, pmass


def halo_seeds_from(centres_box, npart, scale_box, boxsize, max_gather_rad_mpc: float = 3.0):
    b = Box(n1d=0, boxsize=boxsize, omega0=0.3, lambda0=0.7, pmass=0.0, pos=np.zeros((0, 3), np.float32), mom=np.zeros((0, 3), np.float32),
            ids=np.zeros(0, np.uint64), vel_kms=np.zeros((0, 3), np.float32), clump_centres=np.asarray(centres_box), clump_npart=np.asarray(npart),
            clump_scale=np.asarray(scale_box))
    return halo_seeds(b, max_gather_rad_mpc)


def make_host_box(n_host: int, n_sub: int = 40, n1d_bg: int = 64, seed: int = 47, boxsize: float = 64.0, omega0: float = 0.3,
                  lambda0: float = 0.7) -> Box:
    """BASELINE.json configs[4] stand-in for the cooperative unbinding path: ONE Plummer host of `n_host` particles with `n_sub`
    smaller Plummer subclumps inside it, on a sparse jittered-lattice background (equal-mass dark matter).  The host is far
    larger than anything one thread block should walk alone."""
    rng = np.random.default_rng(seed)
    box = float(boxsize)
    sub_n = np.maximum(200, (n_host * 10.0 ** rng.uniform(-4.0, -2.0, size=n_sub)).astype(np.int64))
    n_bg = n1d_bg ** 3
    ntot = int(n_host + sub_n.sum() + n_bg)
    pmass = omega0 * RHOC0 * box ** 3 / ntot
    cn = np.concatenate([[n_host], sub_n]).astype(np.int64)
    mass = cn * pmass
    a_pl = np.clip(0.1 * (mass / 1e14) ** (1.0 / 3.0), 0.02, 0.15)
    centres = np.empty((n_sub + 1, 3))
    centres[0] = 0.5 * box + rng.uniform(-1.0, 1.0, size=3)
    d = rng.normal(size=(n_sub, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    centres[1:] = centres[0] + d * (a_pl[0] * rng.uniform(1.5, 12.0, size=n_sub))[:, None]
    xs, vs = [], []
    cell = box / n1d_bg
    g = (np.arange(n1d_bg) + 0.5) * cell
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3) + rng.normal(0.0, 0.3 * cell, size=(n_bg, 3))
    xs.append(lat.astype(np.float32)); vs.append(rng.normal(0.0, 50.0, size=(n_bg, 3)).astype(np.float32))
    for c in range(n_sub + 1):
        m = int(cn[c])
        u = rng.uniform(1e-9, 1.0 - 1e-6, size=m)
        r = np.minimum(a_pl[c] / np.sqrt(u ** (-2.0 / 3.0) - 1.0), 20.0 * a_pl[c])
        dd = rng.normal(size=(m, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        sig2 = GRAV * mass[c] / (6.0 * a_pl[c]) / np.sqrt(1.0 + (r / a_pl[c]) ** 2)
        xs.append((centres[c] + dd * r[:, None]).astype(np.float32))
        vs.append((rng.normal(0.0, 200.0, size=3) + rng.normal(size=(m, 3)) * np.sqrt(sig2)[:, None]).astype(np.float32))
    x = np.mod(np.concatenate(xs), np.float32(box)).astype(np.float32)
    x = np.where(x >= np.float32(box), np.nextafter(np.float32(box), np.float32(0.0)), x).astype(np.float32)
    v = np.concatenate(vs)
    pos = (x * np.float32(1.0 / box)).astype(np.float32)
    mom = (v * np.float32(1.0 / (box * 100.0))).astype(np.float32)
    return Box(n1d=n1d_bg, boxsize=box, omega0=omega0, lambda0=lambda0, pmass=pmass, pos=pos, mom=mom, ids=np.arange(ntot, dtype=np.uint64),
               vel_kms=v, clump_centres=centres / box, clump_npart=cn, clump_scale=a_pl / box)


# ---------------------------------------------------------------------------------------------------------------------
# multi-species boxes (dark matter + gas + stars) for the -DMULTIMASS -DGAS_PARTICLES build of the reference
# ---------------------------------------------------------------------------------------------------------------------
PGAS, PDM, PSTAR = 0.0, -1.0, -4.0       # reference src/param.h:26-28: what `u` holds for non-gas particles


@dataclass
class SpeciesBox:
    box: Box                 # particles regrouped by GADGET type: gas (0), dark matter (1), stars (4)
    ngas: int
    ndm: int
    nstar: int
    mass: np.ndarray         # (N,) float32 Msun/h
    weight: np.ndarray       # (N,) float32 mass / dark-matter particle mass (io_gadget.c:1595-1599)
    u: np.ndarray            # (N,) float32: thermal energy (km/s)^2 for gas, -type for the others (io_gadget.c:1943-1958)


def make_species_box(n1d: int, seed: int = 42, gas_frac: float = 0.15, star_frac: float = 0.05, gas_mass: float = 0.2, star_mass: float = 0.1,
                     **kw) -> SpeciesBox:
    """make_box() with a random subset of particles turned into lighter gas / star particles."""
    b = make_box(n1d, seed=seed, **kw)
    rng = np.random.default_rng([seed, 77])
    n = b.npart
    t = rng.random(n)
    typ = np.where(t < gas_frac, 0, np.where(t < gas_frac + star_frac, 4, 1))
    order = np.argsort(typ, kind="stable")
    typ = typ[order]
    ngas, ndm, nstar = int((typ == 0).sum()), int((typ == 1).sum()), int((typ == 4).sum())
    w = np.where(typ == 0, gas_mass, np.where(typ == 4, star_mass, 1.0)).astype(np.float32)
    u = np.where(typ == 0, rng.uniform(1e3, 2e5, size=n), np.where(typ == 4, PSTAR, PDM)).astype(np.float32)
    nb = Box(n1d=b.n1d, boxsize=b.boxsize, omega0=b.omega0, lambda0=b.lambda0, pmass=b.pmass, pos=b.pos[order], mom=b.mom[order],
             ids=np.arange(n, dtype=np.uint64), vel_kms=b.vel_kms[order], clump_centres=b.clump_centres, clump_npart=b.clump_npart,
             clump_scale=b.clump_scale)
    return SpeciesBox(box=nb, ngas=ngas, ndm=ndm, nstar=nstar, mass=(w * np.float32(b.pmass)).astype(np.float32), weight=w, u=u)


def write_gadget1_species(sb: SpeciesBox, path: str) -> None:
    """GADGET-1 with types 0/1/4: massarr[1] > 0, a MASS block for gas and stars (the reference's GADGET-1 reader skips exactly four
    blocks POS, VEL, ID, MASS before U, io_gadget.c:1850-1853) and a U block for the gas."""
    b = sb.box
    n = b.npart
    hdr = bytearray(256)
    np_ = [sb.ngas, sb.ndm, 0, 0, sb.nstar, 0]
    massarr = [0.0, b.pmass / 1e10, 0.0, 0.0, 0.0, 0.0]
    struct.pack_into("<6i", hdr, 0, *np_)
    struct.pack_into("<6d", hdr, 24, *massarr)
    struct.pack_into("<2d", hdr, 72, 1.0, 0.0)
    struct.pack_into("<2i", hdr, 88, 0, 0)
    struct.pack_into("<6I", hdr, 96, *np_)
    struct.pack_into("<2i", hdr, 120, 0, 1)
    struct.pack_into("<4d", hdr, 128, b.boxsize, b.omega0, b.lambda0, 0.7)

    def block(f, payload: bytes):
        f.write(struct.pack("<I", len(payload))); f.write(payload); f.write(struct.pack("<I", len(payload)))

    x = (b.pos.astype(np.float32) * np.float32(b.boxsize)).astype("<f4")
    m = (sb.mass.astype(np.float64) / 1e10).astype("<f4")
    with open(path, "wb") as f:
        block(f, bytes(hdr))
        block(f, x.tobytes())
        block(f, b.vel_kms.astype("<f4").tobytes())
        block(f, b.ids.astype("<u4").tobytes())
        block(f, np.concatenate([m[:sb.ngas], m[sb.ngas + sb.ndm:]]).tobytes())          # types with massarr == 0, in type order
        block(f, sb.u[:sb.ngas].astype("<f4").tobytes())


def write_reference_case_species(sb: SpeciesBox, workdir: str, lgrid_domain: int | None = None, **kw) -> str:
    os.makedirs(workdir, exist_ok=True)
    snap = os.path.join(workdir, "snap.gadget")
    write_gadget1_species(sb, snap)
    inp = os.path.join(workdir, "AHF.input")
    write_ahf_input(inp, snap, os.path.join(workdir, "ref"), lgrid_domain or sb.box.n1d, **kw)
    return inp

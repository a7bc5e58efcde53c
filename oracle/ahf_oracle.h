/*
 * ahf_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C, single thread) of the AHF hot path
 * used ONLY as the checker for the CUDA implementation: tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may call it; the product (ahf_b200/csrc, libahfgpu.so) never does.
 *
 * Every function cites the reference file:line (relative to NegriAndrea/AHF) whose behaviour it restates.
 * Parity of this restatement itself is pinned against dumps of the compiled, unmodified reference
 * (oracle/_ref/ahf_ref, see oracle/build_ref.sh + oracle/ref_hooks.c; fixtures under tests/golden/).
 */
#ifndef AHF_ORACLE_H
#define AHF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- K1: Hilbert keys (src/libsfc/hilbert_util.c:69-92, hilbert.c:197-243) --------------------- */
uint64_t orc_hilbert_key(double x, double y, double z, unsigned bits);
void     orc_hilbert_keys(const float *pos3, int64_t n, unsigned bits, uint64_t *keys);
void     orc_hilbert_coords(uint64_t key, unsigned bits, uint32_t out[3]);      /* hilbert.c:139-181 */
/* K2: stable argsort by key (main.c:352-355 sorts unstably; ties live in one 2^-21 cell) */
void     orc_argsort_keys(const uint64_t *keys, int64_t n, int64_t *order);

/* ---- D/F/R/L: level hierarchy ------------------------------------------------------------------ */
typedef struct orc_hier orc_hier;
/* pos3: N x 3 float, box units, already in key-sorted order (array offset = particle handle). */
orc_hier *orc_hier_build(const float *pos3, int64_t n, int64_t lgrid_dom, int64_t lgrid_max,
                         double nth_dom, double nth_ref);
int       orc_hier_nlevels(const orc_hier *h);
/* iout[0]=l1dim iout[1]=ncell iout[2]=npart linked when the level was deposited/flagged
 * iout[3]=npart finally owned; dout[0]=critdens dout[1]=masstopartdens                       */
void      orc_hier_level_header(const orc_hier *h, int lev, int64_t *iout, double *dout);
/* any output pointer may be NULL.  Cells are in the reference's traversal order (z, y, x).
 * runflags: bit0/1 first/last node of its x-run, bit2/3 first/last row of its y-run, bit4/5 first/last
 * plane of its z-run.  mark: 0 untouched, 1 refined, 2 ghost pair (refine_grid.c:231-250).            */
void      orc_hier_level_get(const orc_hier *h, int lev, int32_t *x, int32_t *y, int32_t *z, float *dens,
                             uint8_t *runflags, uint8_t *interior, uint8_t *mark,
                             int32_t *cnt_flag, int64_t *plist_flag, int32_t *cnt_final, int64_t *plist_final);
/* NEXT-1: patch colouring of ahf_gridinfo (src/libahf/ahf_gridinfo.c:236-577, :719-775).  iso[ncell]: index of the
 * isolated refinement each cell of the level belongs to, numbered as the reference numbers them; periodic3[3*niso] (may be NULL):
 * its periodic flags (testBound, :1090-1118).  Returns the number of isolated refinements of the level. */
int64_t   orc_hier_patches(const orc_hier *h, int lev, int32_t *iso, uint8_t *periodic3);
/* six face neighbours per cell as the reference's neighbour search sees them: nb6[6*c + (x-1, x+1, y-1, y+1, z-1, z+1)], -1 = not visible */
void      orc_hier_face_neighbours(const orc_hier *h, int lev, int64_t *nb6);
/* NEXT-2, first half: RefCentre (src/libahf/ahf_halos.c:935-1390).  out[niso][ORC_NPATCH]: 0 numNodes, 1 numParts, 2-4 centre
 * (= centreCMpart: the shipped define.h:101 sets AHFcomcentre), 5 maxDens, 6-8 centreGEOM, 9-11 centreDens; iso / periodic3 / niso from
 * orc_hier_patches of the same level. */
#define ORC_NPATCH 12
void      orc_hier_patch_centres(const orc_hier *h, int lev, const int32_t *iso, const uint8_t *periodic3, int64_t niso, double *out);
void      orc_hier_free(orc_hier *h);

/* ---- G/U/P: halo pass -------------------------------------------------------------------------- */
typedef struct {
  /* unit factors and cosmology scalars (src/libahf/ahf_halos.c:199-221) */
  double r_fac, x_fac, v_fac, m_fac, rho_fac, phi_fac, Hubble;
  double ovlim, rho_vir;
  double vesc_tune;       /* simu.AHF_VTUNE   */
  int    min_part;        /* simu.AHF_MINPART */
} orc_halo_params;

#define ORC_NSCAL 64      /* same slot layout as oracle/ref_hooks.c dump_halos() */
#define ORC_NPROFCOL 25

typedef struct {
  int64_t  npart;             /* final number of members                                           */
  int64_t *ipart;             /* malloc'd, radius-sorted offsets into the key-sorted particle array  */
  int64_t  n_gather, n_rvir0, n_unbound, n_rvir1;
  double   s[ORC_NSCAL];      /* scalars, slots 10.. as in ref_hooks.c                              */
  int      nbins;
  double  *prof;              /* malloc'd, ORC_NPROFCOL x nbins, column-major (col*nbins + bin)      */
  /* GAS_PARTICLES build only (u != NULL): gas_only at 0, stars_only at 32 (SPECIESPROP, src/tdef.h:560-587):
     npart, Mass, pos_com(3), pos_mbp(3), vel(3), lambda, lambdaE, AngMom(3), axis(3), E1(3), E2(3), E3(3), Ekin, Epot */
  double   species[64];
  double  *prof_species;      /* malloc'd, 3 x nbins: M_gas, M_star (cumulative), u_gas (per shell) */
} orc_halo_result;

/* particle arrays in key-sorted order; weight/u may be NULL (equal-mass DM, default reference build) */
void orc_halo_construct(const uint64_t *keys, const float *pos3, const float *mom3, const float *weight,
                        const float *u, int64_t n, const orc_halo_params *par,
                        const double centre[3], double gather_rad, orc_halo_result *out);
void orc_halo_result_free(orc_halo_result *r);

#ifdef __cplusplus
}
#endif
#endif

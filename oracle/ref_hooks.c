/*
 * ref_hooks.c -- TEST INFRASTRUCTURE (oracle side).  Not part of the product.
 *
 * Pass-through wrappers around a handful of functions of the UNMODIFIED reference
 * (NegriAndrea/AHF, compiled from /root/reference/src by oracle/build_ref.sh).  The reference
 * translation units that CALL these functions are compiled with `-Dname=refhook_name`, so the
 * calls land here; every wrapper calls the real function (compiled without the rename) and
 *   (a) accumulates wall-clock time per phase (always), and
 *   (b) dumps intermediate state as flat binary files when $AHF_DUMP_DIR is set.
 * No arithmetic of the reference is touched.
 *
 * Hooked call sites (reference file:line of the call):
 *   main.c:616            gen_domgrids   -> dump of the key-sorted particle array (after main.c:343-356)
 *   main.c:623            ll
 *   main.c:631-632, generate_grids.c:221-222,:314-315   zero_dens / assign_npart
 *   generate_grids.c:206  refine_grid    -> dump of the coarse level as seen by the flagging stencil
 *   generate_grids.c:218  relink
 *   main.c:657            ahf_gridinfo   -> dump of the final hierarchy (cells, dens, owned particles)
 *   main.c:663            ahf_halos
 *   main.c:345,352        sfc_curve_calcKey / qsort (timing only)
 *   ahf_halos.c:508       ahf_halos_sfc_constructHalo -> halo seeds in
 *   ahf_halos.c:824       ahf_io_WriteHalos -> halo results out (all HALO scalars, members, profiles)
 *   ahf_halos_sfc.c:138-151  sort_halo_particles / rem_outsideRvir / rem_unbound / HaloProfiles (stage npart)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <omp.h>

#include "common.h"
#include "param.h"
#include "tdef.h"
#include "libamr_serial/amr_serial.h"
#include "libahf/ahf.h"
#include "libahf/ahf_halos.h"
#include "libahf/ahf_halos_sfc.h"
#include "libutility/utility.h"

extern double r_fac, x_fac, v_fac, m_fac, rho_fac, phi_fac, Hubble;

/* ------------------------------------------------------------------------------------------ */
static const char *dump_dir(void) { return getenv("AHF_DUMP_DIR"); }

static FILE *dump_open(const char *name)
{
  char  path[4096];
  FILE *f;
  snprintf(path, sizeof(path), "%s/%s", dump_dir(), name);
  f = fopen(path, "wb");
  if (!f) { fprintf(stderr, "ref_hooks: cannot open %s\n", path); exit(1); }
  return f;
}

/* phase timers [s] */
static double t_ll, t_deposit, t_refine, t_relink, t_gridinfo, t_halos_total;
static double t_halo_first = -1.0, t_halo_last = 0.0;
static long   n_deposit_part, n_deposit_nodes, n_halo_gathered, n_halo_calls;
static double t_main0;

/* ------------------------------------------------------------------------------------------ */
/* walk one level and write: header, per cell {x,y,z,dens,flags,count}, particle index CSR     */
/* flags: bit0 first node of nquad run, bit1 last node of nquad run, bit2 first row of cquad   */
/* run, bit3 last row, bit4 first plane of pquad run, bit5 last plane                          */
static void dump_level(gridls *g, const char *name)
{
  pqptr  pq; cqptr cq, icq; nqptr nq, inq; nptr nd; partptr p;
  long   x, y, z;
  int64_t ncell = 0, npart = 0, hdr[4];
  double  dh[2];
  FILE   *f;
  int     pass;
  int32_t *cx, *cy, *cz, *cnt; float *dens; uint8_t *flg; int64_t *plist;

  cx = cy = cz = cnt = NULL; dens = NULL; flg = NULL; plist = NULL;
  for (pass = 0; pass < 2; pass++) {
    int64_t ic = 0, ip = 0;
    for (pq = g->pquad; pq != NULL; pq = pq->next) {
      z = pq->z;
      for (cq = pq->loc; cq < pq->loc + pq->length; cq++, z++)
        for (icq = cq; icq != NULL; icq = icq->next) {
          y = icq->y;
          for (nq = icq->loc; nq < icq->loc + icq->length; nq++, y++)
            for (inq = nq; inq != NULL; inq = inq->next) {
              x = inq->x;
              for (nd = inq->loc; nd < inq->loc + inq->length; nd++, x++) {
                if (pass == 1) {
                  uint8_t fl = 0;
                  int32_t c  = 0;
                  if (nd == inq->loc)                        fl |= 1;
                  if (nd == inq->loc + inq->length - 1)      fl |= 2;
                  if (nq == icq->loc)                        fl |= 4;
                  if (nq == icq->loc + icq->length - 1)      fl |= 8;
                  if (cq == pq->loc)                         fl |= 16;
                  if (cq == pq->loc + pq->length - 1)        fl |= 32;
                  cx[ic] = (int32_t)x; cy[ic] = (int32_t)y; cz[ic] = (int32_t)z;
                  dens[ic] = nd->dens; flg[ic] = fl;
                  for (p = nd->ll; p != NULL; p = p->ll) { plist[ip++] = (int64_t)(p - global.fst_part); c++; }
                  cnt[ic] = c;
                } else {
                  for (p = nd->ll; p != NULL; p = p->ll) ip++;
                }
                ic++;
              }
            }
        }
    }
    if (pass == 0) {
      ncell = ic; npart = ip;
      cx   = malloc(sizeof(int32_t) * (ncell + 1)); cy  = malloc(sizeof(int32_t) * (ncell + 1));
      cz   = malloc(sizeof(int32_t) * (ncell + 1)); cnt = malloc(sizeof(int32_t) * (ncell + 1));
      dens = malloc(sizeof(float) * (ncell + 1));   flg = malloc(ncell + 1);
      plist = malloc(sizeof(int64_t) * (npart + 1));
    }
  }
  f = dump_open(name);
  hdr[0] = (int64_t)g->l1dim; hdr[1] = ncell; hdr[2] = npart; hdr[3] = 0;
  dh[0]  = g->critdens; dh[1] = g->masstopartdens;
  fwrite(hdr, sizeof(int64_t), 4, f);
  fwrite(dh, sizeof(double), 2, f);
  fwrite(cx, sizeof(int32_t), ncell, f); fwrite(cy, sizeof(int32_t), ncell, f); fwrite(cz, sizeof(int32_t), ncell, f);
  fwrite(dens, sizeof(float), ncell, f); fwrite(flg, 1, ncell, f); fwrite(cnt, sizeof(int32_t), ncell, f);
  fwrite(plist, sizeof(int64_t), npart, f);
  fclose(f);
  free(cx); free(cy); free(cz); free(cnt); free(dens); free(flg); free(plist);
}

/* ------------------------------------------------------------------------------------------ */
/* main.c:616 -- particles are key-sorted at this point                                         */
gridls *refhook_gen_domgrids(int *no_grids)
{
  t_main0 = omp_get_wtime();
  if (dump_dir()) {
    FILE    *f = dump_open("particles.bin");
    uint64_t n = global_info.no_part, i;
    partptr  p = global_info.fst_part;
    double   sc[8];
    fwrite(&n, sizeof(n), 1, f);
    sc[0] = simu.boxsize; sc[1] = simu.pmass; sc[2] = simu.t_unit; sc[3] = global.a;
    sc[4] = simu.omega0;  sc[5] = simu.lambda0; sc[6] = (double)simu.no_vpart; sc[7] = (double)sizeof(part);
    fwrite(sc, sizeof(double), 8, f);
    for (i = 0; i < n; i++) { uint64_t id = (uint64_t)p[i].id; fwrite(&id, 8, 1, f); }
    for (i = 0; i < n; i++) { uint64_t k = (uint64_t)p[i].sfckey; fwrite(&k, 8, 1, f); }
    for (i = 0; i < n; i++) { float v[3] = { p[i].pos[0], p[i].pos[1], p[i].pos[2] }; fwrite(v, 4, 3, f); }
    for (i = 0; i < n; i++) { float v[3] = { p[i].mom[0], p[i].mom[1], p[i].mom[2] }; fwrite(v, 4, 3, f); }
#ifdef MULTIMASS
    for (i = 0; i < n; i++) { float w = p[i].weight; fwrite(&w, 4, 1, f); }
#endif
#ifdef GAS_PARTICLES
    for (i = 0; i < n; i++) { float u = p[i].u; fwrite(&u, 4, 1, f); }
#endif
    fclose(f);
  }
  return gen_domgrids(no_grids);
}

void refhook_ll(long unsigned npart, partptr fst_part, gridls *cur_grid)
{
  double t = omp_get_wtime();
  ll(npart, fst_part, cur_grid);
  t_ll += omp_get_wtime() - t;
}

void refhook_zero_dens(gridls *g)
{
  double t = omp_get_wtime();
  zero_dens(g);
  t_deposit += omp_get_wtime() - t;
}

boolean refhook_assign_npart(gridls *g)
{
  double  t = omp_get_wtime();
  boolean r = assign_npart(g);
  t_deposit += omp_get_wtime() - t;
  n_deposit_part  += (long)g->size.no_part;
  n_deposit_nodes += (long)g->size.no_nodes;
  return r;
}

static int n_refine_calls = 0;
boolean refhook_refine_grid(gridls *fin, gridls *coa)
{
  double  t;
  boolean r;
  if (dump_dir()) {
    char name[64];
    snprintf(name, sizeof(name), "flag_level_%02d.bin", n_refine_calls);
    dump_level(coa, name);
  }
  n_refine_calls++;
  t = omp_get_wtime();
  r = refine_grid(fin, coa);
  t_refine += omp_get_wtime() - t;
  return r;
}

boolean refhook_relink(gridls *coa, gridls *fin)
{
  double  t = omp_get_wtime();
  boolean r = relink(coa, fin);
  t_relink += omp_get_wtime() - t;
  return r;
}

/* NEXT-1: the outcome of the patch colouring of ahf_gridinfo (ahf_gridinfo.c:236-577, :719-775), read from the
 * reference's own globals after the call: per coloured level the isolated-refinement index of every node in traversal order
 * (spatialRefIndex[node colour].isoRefIndex), the number of isolated refinements and their periodic flags. */
extern SRINDEX *spatialRefIndex;
extern int     *numIsoRef;
static void dump_patches(void)
{
  FILE   *f = dump_open("patches.bin");
  int32_t hdr[2];
  int     i;
  hdr[0] = ahf.min_ref; hdr[1] = ahf.no_grids;
  fwrite(hdr, sizeof(int32_t), 2, f);
  for (i = 0; i < ahf.no_grids; i++) {
    gridls *g = global.dom_grid + ahf.min_ref + i;
    pqptr   pq; cqptr cq, icq; nqptr nq, inq; nptr nd;
    int64_t ncell = 0, ic = 0, lh[2];
    int32_t *iso;
    int8_t  *per = calloc((size_t)(3 * numIsoRef[i] + 3), 1);
    int      pass;
    iso = NULL;
    for (pass = 0; pass < 2; pass++) {
      ic = 0;
      for (pq = g->pquad; pq != NULL; pq = pq->next)
        for (cq = pq->loc; cq < pq->loc + pq->length; cq++)
          for (icq = cq; icq != NULL; icq = icq->next)
            for (nq = icq->loc; nq < icq->loc + icq->length; nq++)
              for (inq = nq; inq != NULL; inq = inq->next)
                for (nd = inq->loc; nd < inq->loc + inq->length; nd++) {
                  if (pass == 1) {
                    const SRINDEX *sr = spatialRefIndex + nd->force.colour;
                    iso[ic] = (sr->refLevel == i) ? sr->isoRefIndex : -1000 - sr->refLevel;
                    if (sr->refLevel == i && sr->isoRefIndex >= 0 && sr->isoRefIndex < numIsoRef[i]) {
                      per[3 * sr->isoRefIndex + 0] = (int8_t)sr->periodic.x;
                      per[3 * sr->isoRefIndex + 1] = (int8_t)sr->periodic.y;
                      per[3 * sr->isoRefIndex + 2] = (int8_t)sr->periodic.z;
                    }
                  }
                  ic++;
                }
      if (pass == 0) { ncell = ic; iso = malloc(sizeof(int32_t) * (size_t)(ncell + 1)); }
    }
    lh[0] = ncell; lh[1] = numIsoRef[i];
    fwrite(lh, sizeof(int64_t), 2, f);
    fwrite(iso, sizeof(int32_t), (size_t)ncell, f);
    fwrite(per, 1, (size_t)(3 * numIsoRef[i]), f);
    free(iso); free(per);
  }
  fclose(f);
}

void refhook_ahf_gridinfo(gridls *grid_list, int curgrid_no)
{
  double t;
  if (dump_dir()) {
    int   i;
    FILE *f;
    for (i = global.domgrid_no; i <= curgrid_no; i++) {
      char name[64];
      snprintf(name, sizeof(name), "final_level_%02d.bin", i - global.domgrid_no);
      dump_level(grid_list + i, name);
    }
    f = dump_open("hierarchy.txt");
    fprintf(f, "nlevels %d\n", curgrid_no - global.domgrid_no + 1);
    for (i = global.domgrid_no; i <= curgrid_no; i++)
      fprintf(f, "level %d l1dim %lu nodes %lu parts %lu\n", i - global.domgrid_no, (grid_list + i)->l1dim,
              (grid_list + i)->size.no_nodes, (grid_list + i)->size.no_part);
    fclose(f);
  }
  t = omp_get_wtime();
  ahf_gridinfo(grid_list, curgrid_no);
  t_gridinfo += omp_get_wtime() - t;
  if (dump_dir()) dump_patches();
}

/* ------------------------------------------------------------------------------------------ */
/* halo pass                                                                                    */
typedef struct {
  HALO    *halo;                 /* address = ordering key                                      */
  double   in_pos[3], in_gatherRad;
  uint64_t in_npart;
  uint64_t n_gather, n_rvir0, n_unbound, n_rvir1;
  int      nbins;
  int      host_pre, hostlevel_pre, nsub_pre;     /* hostHalo / hostHaloLevel / subStruct[] as spatialRef2halos left them (before the re-hash) */
  int     *sub_pre;
} halorec;

static halorec *recs  = NULL;
static long     nrecs = 0, caprecs = 0;
static __thread halorec *cur_rec = NULL;

void refhook_sort(HALO *h)
{
  if (cur_rec) cur_rec->n_gather = h->npart;
  sort_halo_particles(h);
}
void refhook_rvir(HALO *h, int icall)
{
  rem_outsideRvir(h, icall);
  if (cur_rec) { if (icall == 0) cur_rec->n_rvir0 = h->npart; else cur_rec->n_rvir1 = h->npart; }
}
void refhook_unbound(HALO *h)
{
  rem_unbound(h);
  if (cur_rec) cur_rec->n_unbound = h->npart;
}
int refhook_profiles(HALO *h)
{
  return HaloProfiles(h);
}

void refhook_constructHalo(HALO *h)
{
  halorec r;
  double  t0 = omp_get_wtime(), t1;
  memset(&r, 0, sizeof(r));
  r.halo = h;
  r.in_pos[0] = h->pos.x; r.in_pos[1] = h->pos.y; r.in_pos[2] = h->pos.z;
  r.in_gatherRad = h->gatherRad; r.in_npart = h->npart;
  r.host_pre = h->hostHalo; r.hostlevel_pre = h->hostHaloLevel; r.nsub_pre = h->numSubStruct; r.sub_pre = NULL;
  if (dump_dir() && h->numSubStruct > 0 && h->subStruct) {
    r.sub_pre = malloc(h->numSubStruct * sizeof(int));
    memcpy(r.sub_pre, h->subStruct, h->numSubStruct * sizeof(int));
  }
  cur_rec = &r;
  ahf_halos_sfc_constructHalo(h);
  cur_rec = NULL;
  t1 = omp_get_wtime();
#pragma omp critical(refhook_rec)
  {
    if (t_halo_first < 0 || t0 < t_halo_first) t_halo_first = t0;
    if (t1 > t_halo_last) t_halo_last = t1;
    n_halo_gathered += (long)r.n_gather;
    n_halo_calls++;
    if (dump_dir()) {
      if (nrecs == caprecs) { caprecs = caprecs ? 2 * caprecs : 1024; recs = realloc(recs, caprecs * sizeof(halorec)); }
      recs[nrecs++] = r;
    }
  }
}


#define NSCAL 64
static HALO *g_halos = NULL; static int g_numHalos = 0;
static double u_fac_of_run(void) { double q = simu.boxsize / simu.t_unit; return q * q; }     /* ahf_halos.c:203 */
/* halo_tree.bin: per halo (halos[] order) hostHalo, hostHaloLevel, numSubStruct, subStruct[] BEFORE the sub-halo re-hash (int32 each),
 * then hostHalo, numSubStruct, subStruct[] AFTER it -- input and expected output of the re-hash restatement */
static void dump_halo_tree(halorec *byidx)
{
  FILE *f = dump_open("halo_tree.bin");
  long  i;
  int32_t n = g_numHalos;
  fwrite(&n, sizeof(int32_t), 1, f);
  for (i = 0; i < g_numHalos; i++) {
    HALO   *h = g_halos + i;
    int32_t v[3] = { byidx[i].host_pre, byidx[i].hostlevel_pre, byidx[i].nsub_pre }, w[2] = { h->hostHalo, h->numSubStruct };
    int     k;
    fwrite(v, sizeof(int32_t), 3, f);
    for (k = 0; k < byidx[i].nsub_pre; k++) { int32_t q = byidx[i].sub_pre ? byidx[i].sub_pre[k] : -1; fwrite(&q, sizeof(int32_t), 1, f); }
    fwrite(w, sizeof(int32_t), 2, f);
    for (k = 0; k < h->numSubStruct; k++) { int32_t q = h->subStruct[k]; fwrite(&q, sizeof(int32_t), 1, f); }
  }
  fclose(f);
}
static void dump_halos(void)
{
  FILE *f, *fi, *fp;
  long  i;
  int   k;
  halorec *byidx = calloc(g_numHalos > 0 ? g_numHalos : 1, sizeof(halorec));
  for (i = 0; i < nrecs; i++) { long j = recs[i].halo - g_halos; if (j >= 0 && j < g_numHalos) byidx[j] = recs[i]; }
  dump_halo_tree(byidx);
  f  = dump_open("halos.bin");
  fi = dump_open("halo_ipart.bin");
  fp = dump_open("halo_prof.bin");
#ifdef GAS_PARTICLES
  FILE *fs = dump_open("halo_species.bin"), *fps = dump_open("halo_prof_species.bin");
#endif
  {
    int64_t hdr[2] = { g_numHalos, NSCAL };
    double  g[16];
    memset(g, 0, sizeof(g));
    g[0] = r_fac; g[1] = x_fac; g[2] = v_fac; g[3] = m_fac; g[4] = rho_fac; g[5] = phi_fac; g[6] = Hubble;
    g[7] = global.ovlim; g[8] = global.rho_vir; g[9] = (double)simu.AHF_MINPART; g[10] = simu.AHF_VTUNE;
    g[11] = simu.MaxGatherRad; g[12] = global.a; g[13] = u_fac_of_run(); g[14] = simu.pmass; g[15] = global.z;
    fwrite(hdr, sizeof(int64_t), 2, f);
    fwrite(g, sizeof(double), 16, f);
  }
  for (i = 0; i < g_numHalos; i++) {
    HALO  *h = g_halos + i;
    double s[NSCAL];
    halorec *rc = byidx + i;
    int64_t np = (rc->halo == NULL || rc->in_npart == 0) ? 0 : (int64_t)h->npart;
    int64_t nb = 0;
    memset(s, 0, sizeof(s));
    k = 0;
    s[k++] = rc->in_pos[0]; s[k++] = rc->in_pos[1]; s[k++] = rc->in_pos[2];
    s[k++] = rc->in_gatherRad; s[k++] = (double)rc->in_npart;
    s[k++] = (double)rc->n_gather; s[k++] = (double)rc->n_rvir0;
    s[k++] = (double)rc->n_unbound; s[k++] = (double)rc->n_rvir1;
    s[k++] = (double)np;                                   /* 9 */
    s[k++] = h->M_vir; s[k++] = h->R_vir; s[k++] = h->ovdens; s[k++] = h->Phi0;       /* 10-13 */
    if (np >= simu.AHF_MINPART) {
      nb = h->prof.nbins;
      s[k++] = h->vel.x; s[k++] = h->vel.y; s[k++] = h->vel.z;                        /* 14-16 */
      s[k++] = h->sigV; s[k++] = h->v_esc2; s[k++] = h->V2_max; s[k++] = h->R_max; s[k++] = h->r2; /* 17-21 */
      s[k++] = h->lambda; s[k++] = h->lambdaE; s[k++] = h->Ekin; s[k++] = h->Epot; s[k++] = h->SurfP; /* 22-26 */
      s[k++] = h->pos_com.x; s[k++] = h->pos_com.y; s[k++] = h->pos_com.z; s[k++] = h->com_offset;  /* 27-30 */
      s[k++] = h->pos_mbp.x; s[k++] = h->pos_mbp.y; s[k++] = h->pos_mbp.z;            /* 31-33 */
      s[k++] = h->vel_mbp.x; s[k++] = h->vel_mbp.y; s[k++] = h->vel_mbp.z; s[k++] = h->mbp_offset;  /* 34-37 */
      s[k++] = h->AngMom.x; s[k++] = h->AngMom.y; s[k++] = h->AngMom.z;               /* 38-40 */
      s[k++] = h->axis.x; s[k++] = h->axis.y; s[k++] = h->axis.z;                     /* 41-43 */
      s[k++] = h->E1.x; s[k++] = h->E1.y; s[k++] = h->E1.z;                           /* 44-46 */
      s[k++] = h->E2.x; s[k++] = h->E2.y; s[k++] = h->E2.z;                           /* 47-49 */
      s[k++] = h->E3.x; s[k++] = h->E3.y; s[k++] = h->E3.z;                           /* 50-52 */
      s[k++] = h->fMhires; s[k++] = h->cNFW; s[k++] = h->cR1; s[k++] = h->R1;         /* 53-56 */
      s[k++] = (double)nb;                                                            /* 57 */
    }
    s[58] = (double)h->hostHalo; s[59] = (double)h->numSubStruct; s[60] = h->spaRes; s[61] = (double)h->refLev;
    s[62] = (double)h->numNodes;
    fwrite(s, sizeof(double), NSCAL, f);
    /* members */
    fwrite(&np, sizeof(int64_t), 1, fi);
    for (k = 0; k < np; k++) { int64_t ip = (int64_t)h->ipart[k]; fwrite(&ip, sizeof(int64_t), 1, fi); }
    /* profiles: nbins, then 25 columns of nbins doubles (npart cast to double) */
    fwrite(&nb, sizeof(int64_t), 1, fp);
    if (nb > 0) {
      int     b;
      double *cols[24] = { h->prof.r, h->prof.nvpart, h->prof.ovdens, h->prof.dens, h->prof.v2_circ, h->prof.v_esc2,
                           h->prof.sig_v, h->prof.Ekin, h->prof.Epot, h->prof.Lx, h->prof.Ly, h->prof.Lz,
                           h->prof.axis1, h->prof.E1x, h->prof.E1y, h->prof.E1z, h->prof.axis2, h->prof.E2x,
                           h->prof.E2y, h->prof.E2z, h->prof.axis3, h->prof.E3x, h->prof.E3y, h->prof.E3z };
      for (b = 0; b < nb; b++) { double v = (double)h->prof.npart[b]; fwrite(&v, sizeof(double), 1, fp); }
      for (k = 0; k < 24; k++) fwrite(cols[k], sizeof(double), nb, fp);
    }
#ifdef GAS_PARTICLES
    {
      /* gas_only at 0, stars_only at 32: npart, Mass, pos_com(3), pos_mbp(3), vel(3), lambda, lambdaE, AngMom(3), axis(3), E1, E2, E3, Ekin, Epot */
      double sp[64];
      int    q;
      memset(sp, 0, sizeof(sp));
      if (np >= simu.AHF_MINPART)
        for (q = 0; q < 2; q++) {
          SPECIESPROP *t = q ? &h->stars_only : &h->gas_only;
          double *o = sp + 32 * q;
          o[0] = (double)t->npart; o[1] = t->Mass; o[2] = t->pos_com.x; o[3] = t->pos_com.y; o[4] = t->pos_com.z;
          o[5] = t->pos_mbp.x; o[6] = t->pos_mbp.y; o[7] = t->pos_mbp.z; o[8] = t->vel.x; o[9] = t->vel.y; o[10] = t->vel.z;
          o[11] = t->lambda; o[12] = t->lambdaE; o[13] = t->AngMom.x; o[14] = t->AngMom.y; o[15] = t->AngMom.z;
          o[16] = t->axis.x; o[17] = t->axis.y; o[18] = t->axis.z;
          o[19] = t->E1.x; o[20] = t->E1.y; o[21] = t->E1.z; o[22] = t->E2.x; o[23] = t->E2.y; o[24] = t->E2.z;
          o[25] = t->E3.x; o[26] = t->E3.y; o[27] = t->E3.z; o[28] = t->Ekin; o[29] = t->Epot;
        }
      fwrite(sp, sizeof(double), 64, fs);
      fwrite(&nb, sizeof(int64_t), 1, fps);
      if (nb > 0) { fwrite(h->prof.M_gas, sizeof(double), nb, fps); fwrite(h->prof.M_star, sizeof(double), nb, fps); fwrite(h->prof.u_gas, sizeof(double), nb, fps); }
    }
#endif
  }
  fclose(f); fclose(fi); fclose(fp);
#ifdef GAS_PARTICLES
  fclose(fs); fclose(fps);
#endif
}

/* ahf_halos.c:824 -- first writer call; halos[] is complete (incl. the subhalo re-hash) and still alive */
void refhook_WriteHalos(const char *fprefix, HALO *halos, unsigned long *idx, int numHalos)
{
  if (dump_dir()) { g_halos = halos; g_numHalos = numHalos; dump_halos(); }
  ahf_io_WriteHalos(fprefix, halos, idx, numHalos);
}

/* main.c:343-356 -- key generation + qsort timing */
static double t_key_first = -1.0, t_keys = 0.0, t_sort = 0.0;
sfc_key_t refhook_calcKey(sfc_curve_t ctype, double x, double y, double z, uint32_t bits)
{
  if (t_key_first < 0) t_key_first = omp_get_wtime();
  return sfc_curve_calcKey(ctype, x, y, z, bits);
}
void refhook_qsort(void *base, size_t n, size_t sz, int (*cmp)(const void *, const void *))
{
  double t = omp_get_wtime();
  if (t_key_first >= 0 && t_keys == 0.0) t_keys = t - t_key_first;
  qsort(base, n, sz, cmp);
  t_sort += omp_get_wtime() - t;
}

void refhook_ahf_halos(gridls *grid_list)
{
  double t = omp_get_wtime();
  FILE  *f;
  ahf_halos(grid_list);
  t_halos_total = omp_get_wtime() - t;
  /* timing summary (always) */
  f = stderr;
  fprintf(f, "REFHOOK_TIMING threads=%d keys=%.6f sort=%.6f ll=%.6f deposit=%.6f deposit_parts=%ld deposit_nodes=%ld refine=%.6f relink=%.6f "
             "gridinfo=%.6f ahf_halos=%.6f halo_loop=%.6f halo_calls=%ld halo_gathered=%ld\n",
          omp_get_max_threads(), t_keys, t_sort, t_ll, t_deposit, n_deposit_part, n_deposit_nodes, t_refine, t_relink, t_gridinfo,
          t_halos_total, (t_halo_first < 0) ? 0.0 : (t_halo_last - t_halo_first), n_halo_calls, n_halo_gathered);
  if (dump_dir()) {
    FILE *g = dump_open("timing.txt");
    fprintf(g, "threads %d\nkeys %.6f\nsort %.6f\nll %.6f\ndeposit %.6f\ndeposit_parts %ld\ndeposit_nodes %ld\nrefine %.6f\nrelink %.6f\n"
               "gridinfo %.6f\nahf_halos %.6f\nhalo_loop %.6f\nhalo_calls %ld\nhalo_gathered %ld\n",
            omp_get_max_threads(), t_keys, t_sort, t_ll, t_deposit, n_deposit_part, n_deposit_nodes, t_refine, t_relink, t_gridinfo,
            t_halos_total, (t_halo_first < 0) ? 0.0 : (t_halo_last - t_halo_first), n_halo_calls, n_halo_gathered);
    fclose(g);
  }
}

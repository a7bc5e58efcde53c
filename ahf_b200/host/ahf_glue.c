/*
 * ahf_glue.c -- host side of the drop-in, in the reference's own language (C99).
 *
 * Compiled TOGETHER WITH the unmodified reference sources (NegriAndrea/AHF) by build_dropin.sh; nothing of the
 * reference is copied or edited: the reference translation units that contain the path's call sites are compiled with
 * `-Dcallee=ahfb200_callee`, which lands the calls in this file, and this file calls libahfgpu.so (include/ahfgpu.h).
 *
 *   src/main.c:343-356        key loop + qsort            -> ahfb200_calcKey (stub) + ahfb200_qsort  -> ahfgpu_sfc_sort_particles
 *   src/libahf/ahf_halos.c:504-510   OpenMP halo loop     -> ahfb200_constructHalo (collects the HALO pointers); the batched device pass
 *                                                            runs when the loop's parallel region ends (GOMP_parallel below: deterministic,
 *                                                            independent of VERBOSE) -> ahfgpu_construct_halos / ahfgpu_halo_fetch
 *                                                            -> HALO, c_profile arrays.  A build without OpenMP handles every halo at once.
 *   src/main.c:616-648        gen_domgrids / ll / zero_dens / assign_npart / gen_AMRhierarchy
 *                                                         -> ahfb200_gen_domgrids -> ahfgpu_build_amr + rebuild of the reference's quads
 *                                                            (only in the AHF-b200 build; AHF-b200-kh keeps the CPU mesh)
 * Everything else -- readers, ahf_gridinfo, tree, subhalo re-hash, writers -- is the reference, so the catalogues come out in
 * AHF's own formats.
 *
 * -DAHFB200_FULL (binaries AHF-b200-full, AHF-b200-mm-full): main.c's calls of ahf_gridinfo (main.c:657) and ahf_halos (:662) land here
 * as well.  ahfb200_halos does what ahf_halos() does (src/libahf/ahf_halos.c:166-930) with the library alone: patch labels and
 * per-refinement tables on the device (NEXT-1/2: ahfgpu_amr_patch_stats), tree and halo seeds (ahfgpu_tree_halos_ex), the halo pass
 * (ahfgpu_construct_halos), sub-halo re-hash, ordering and the four catalogue files (NEXT-3: ahfgpu_catalogue_write).  No quads are
 * rebuilt and none of the reference's mesh walkers run; of the reference there remain the reader, startrun and the log file.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE                /* RTLD_NEXT */
#endif
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stddef.h>
#include <string.h>
#include <omp.h>

#include "common.h"
#include "param.h"
#include "tdef.h"
#include "libutility/utility.h"
#include "libamr_serial/amr_serial.h"
#include "ahfgpu.h"

extern double r_fac, x_fac, v_fac, m_fac, rho_fac, phi_fac, Hubble;      /* src/libahf/ahf_halos.c:163 */

static ahfgpu_ctx *G = NULL;
static uint64_t   *G_ids = NULL;        /* ID block of a snapshot read by the bulk ingest (input order) */
static int         G_ingested = 0;

/* AHFB200_TIMING=1: wall clock of the program's phases on stderr at exit (one line, key=seconds), scripts/dropin_timing.py reads it */
#include <time.h>
static double T0 = 0.0, T_LAST = 0.0;
static char   TLINE[2048];
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
static void   tmark(const char *what)
{
  double t = now_s();
  size_t l = strlen(TLINE);
  if (l + 64 < sizeof(TLINE)) snprintf(TLINE + l, sizeof(TLINE) - l, " %s=%.4f", what, t - T_LAST);
  T_LAST = t;
}
static void tprint(void)
{
  if (!getenv("AHFB200_TIMING")) return;
  tmark("tail");
  fprintf(stderr, "AHFB200_TIMING total=%.4f%s\n", now_s() - T0, TLINE);
}

static void die(const char *what)
{
  fprintf(stderr, "ahf_glue: %s: %s\n", what, ahfgpu_last_error());
  common_terminate(EXIT_FAILURE);
}

static void fill_params(ahfgpu_params *p)
{
  memset(p, 0, sizeof(*p));
  p->device = 0;
  p->lgrid_dom = simu.NGRID_DOM; p->lgrid_max = simu.NGRID_MAX > (1 << 21) ? (1 << 21) : simu.NGRID_MAX;
  if (p->lgrid_dom < 4) p->lgrid_dom = 64;     /* the reader runs before AHF.input's LgridDomain reaches simu (startrun.c:93 / :499): replaced by ahfgpu_set_params */
  p->nth_dom = simu.Nth_dom; p->nth_ref = simu.Nth_ref; p->min_part = simu.AHF_MINPART; p->vesc_tune = simu.AHF_VTUNE;
  p->r_fac = r_fac; p->x_fac = x_fac; p->v_fac = v_fac; p->m_fac = m_fac; p->rho_fac = rho_fac; p->phi_fac = phi_fac;
  p->hubble = Hubble; p->ovlim = global.ovlim; p->rho_vir = global.rho_vir;
}

/* The CUDA context (driver start-up, module load: 0.5-1 s) is created by a helper thread from program start, while main() parses AHF.input
 * and opens the snapshot; ensure_ctx waits for it. */
#include <pthread.h>
static pthread_t WARM_T;
static int       WARM_ON = 0;
static void *warm_main(void *arg) { (void)arg; ahfgpu_warmup(0); return NULL; }
static void  warm_start(void) { if (!getenv("AHFB200_NO_WARMUP") && pthread_create(&WARM_T, NULL, warm_main, NULL) == 0) WARM_ON = 1; }
static void  warm_join(void) { if (WARM_ON) { pthread_join(WARM_T, NULL); WARM_ON = 0; } }

static void ensure_ctx(void)
{
  ahfgpu_params p;
  if (G) return;
  tmark("before_init");
  warm_join();
  fill_params(&p);
  if (ahfgpu_init(&G, &p)) die("ahfgpu_init");
  tmark("ahfgpu_init");
}

/* ---- K: src/main.c:343-356 ------------------------------------------------------------------------------- */
sfc_key_t ahfb200_calcKey(sfc_curve_t ctype, double x, double y, double z, uint32_t bits)
{
  (void)ctype; (void)x; (void)y; (void)z; (void)bits;
  return 0;                                   /* keys are produced by the device sort below */
}

void ahfb200_qsort(void *base, size_t n, size_t sz, int (*cmp)(const void *, const void *))
{
  if (base == (void *)global_info.fst_part && sz == sizeof(part)) {
    ensure_ctx();
    if (G_ingested) {                         /* the particles are on the device already (ahfb200_gadget_readpart): keys + sort there */
      if (ahfgpu_sfc_sort_resident(G)) die("ahfgpu_sfc_sort_resident");
      tmark("sfc_sort_resident");
      return;
    }
    if (ahfgpu_sfc_sort_particles(G, base, (uint64_t)n, (uint32_t)sizeof(part), (int32_t)offsetof(part, pos), (int32_t)offsetof(part, mom),
                                  (int32_t)offsetof(part, sfckey), (int32_t)offsetof(part, id),
#ifdef MULTIMASS
                                  (int32_t)offsetof(part, weight),
#else
                                  -1,
#endif
#ifdef GAS_PARTICLES
                                  (int32_t)offsetof(part, u)
#else
                                  -1
#endif
                                  )) die("ahfgpu_sfc_sort_particles");
    tmark("sfc_sort_particles");
    return;
  }
  qsort(base, n, sz, cmp);
}

/* ---- M: src/main.c:616-648 ------------------------------------------------------------------------------
 * gen_domgrids + ll + zero_dens + assign_npart + gen_AMRhierarchy  ->  ahfgpu_build_amr, then the device hierarchy is
 * rebuilt as the reference's own gridls / pquad / cquad / nquad / node structures (tdef.h:110-235) for the host code that
 * still consumes them (ahf_gridinfo, RefCentre ...).  The per-level cell list with run flags (ahfgpu_amr_level_get) carries
 * exactly the information of the quads; node particle lists are rebuilt in the reference's order (head insertion: a node of
 * an even level lists its particles by descending offset, of an odd level by ascending offset -- what ll()/relink() leave).
 * Everything is allocated with the reference's c_* allocators so that free_grid() (ahf_halos.c:483) works.             */
static void build_level_quads(gridls *g, int64_t nc, const int32_t *x, const int32_t *y, const int32_t *z, const float *dens,
                              const uint8_t *rf, nptr *nodeptr)
{
  int64_t c = 0;
  pqptr   pq = NULL, pq_tail = NULL;
  long    npq = 0;
  while (c < nc) {
    /* one z-run (pquad): planes until a "last plane" flag */
    int64_t c_run = c, nplanes = 0, cc = c;
    pqptr   newpq = c_pquad(1);
    /* count planes of the run */
    while (cc < nc) {
      int32_t zz = z[cc];
      int     lastplane = (rf[cc] & 32) != 0;
      while (cc < nc && z[cc] == zz) cc++;
      nplanes++;
      if (lastplane) break;
    }
    newpq->z = z[c_run]; newpq->length = (int)nplanes; newpq->loc = c_cquad(nplanes); newpq->next = NULL;
    if (pq_tail) pq_tail->next = newpq; else pq = newpq;
    pq_tail = newpq; npq++;
    for (int64_t ip = 0; ip < nplanes; ip++) {
      /* plane: cells [c, cend) */
      int64_t cend = c;
      cqptr   cq = newpq->loc + ip;
      int     first_yrun = 1;
      while (cend < nc && z[cend] == z[c]) cend++;
      while (c < cend) {
        /* one y-run (cquad): rows until a "last row" flag */
        int64_t r = c, nrows = 0;
        while (r < cend) {
          int32_t yy = y[r];
          int     lastrow = (rf[r] & 8) != 0;
          while (r < cend && y[r] == yy) r++;
          nrows++;
          if (lastrow) break;
        }
        if (!first_yrun) { cq->next = c_cquad(1); cq = cq->next; }
        first_yrun = 0;
        cq->y = y[c]; cq->length = (int)nrows; cq->loc = c_nquad(nrows); cq->next = NULL;
        for (int64_t ir = 0; ir < nrows; ir++) {
          int64_t rend = c;
          nqptr   nq = cq->loc + ir;
          int     first_xrun = 1;
          while (rend < cend && y[rend] == y[c]) rend++;
          while (c < rend) {
            /* one x-run (nquad): nodes until a "last node" flag */
            int64_t e = c;
            while (e < rend && !(rf[e] & 2)) e++;
            e++;                                            /* include the flagged last node */
            if (!first_xrun) { nq->next = c_nquad(1); nq = nq->next; }
            first_xrun = 0;
            nq->x = x[c]; nq->length = (int)(e - c); nq->loc = c_node(e - c); nq->next = NULL;
            for (int64_t k = c; k < e; k++) { nq->loc[k - c].dens = dens[k]; nq->loc[k - c].ll = NULL; nodeptr[k] = nq->loc + (k - c); }
            c = e;
          }
        }
      }
    }
  }
  g->pquad = pq;
  g->no_pquad = npq;
  g->pquad_array = (pqptr *)calloc(npq > 0 ? npq : 1, sizeof(pqptr));
  { long i = 0; for (pqptr q = pq; q != NULL; q = q->next) g->pquad_array[i++] = q; }
}

#ifdef AHFB200_FULL
/* nobody walks the reference's quads in this build: the grid list carries the level headers only (write_logfile, specific.c:1132) */
gridls *ahfb200_gen_domgrids(int *no_grids)
{
  ahfgpu_params p;
  gridls       *gl;
  int           nlev, l;
  if (simu.NGRID_MIN != simu.NGRID_DOM) { fprintf(stderr, "ahf_glue: NGRID_MIN != NGRID_DOM is not supported\n"); common_terminate(EXIT_FAILURE); }
  ensure_ctx();
  tmark("main_to_mesh");
  fill_params(&p);
  if (ahfgpu_set_params(G, &p)) die("ahfgpu_set_params");
  if (ahfgpu_build_amr(G)) die("ahfgpu_build_amr");
  tmark("build_amr");
  nlev = ahfgpu_amr_nlevels(G);
  gl = (gridls *)calloc(nlev, sizeof(gridls));
  for (l = 0; l < nlev; l++) {
    int64_t io[4]; double dd[2];
    gridls *g = gl + l;
    if (ahfgpu_amr_level_header(G, l, io, dd)) die("ahfgpu_amr_level_header");
    g->l1dim = (long unsigned)io[0]; g->spacing = 1.0 / (double)io[0]; g->spacing2 = g->spacing * g->spacing;
    g->critdens = dd[0]; g->masstopartdens = dd[1];
    g->masstodens = dd[1] * (double)simu.no_part / simu.no_vpart;
    g->size.no_part = (long unsigned)io[2]; g->size.no_nodes = (long unsigned)io[1];
    g->timecounter = global.super_t; g->next = (l < nlev - 1) ? TRUE : FALSE;
  }
  global.dom_grid = gl; global.domgrid_no = 0; global.fin_l1dim = gl[nlev - 1].l1dim;
  *no_grids = nlev;
  return gl;
}
#else
gridls *ahfb200_gen_domgrids(int *no_grids)
{
  ahfgpu_params p;
  gridls       *gl;
  int           nlev, l;
  uint64_t      n = global_info.no_part, i;
  int8_t       *owner;
  int32_t      *cell_of;
  nptr        **nodeptr;
  if (simu.NGRID_MIN != simu.NGRID_DOM) { fprintf(stderr, "ahf_glue: NGRID_MIN != NGRID_DOM is not supported\n"); common_terminate(EXIT_FAILURE); }
  ensure_ctx();
  fill_params(&p);
  if (ahfgpu_set_params(G, &p)) die("ahfgpu_set_params");
  if (ahfgpu_build_amr(G)) die("ahfgpu_build_amr");
  nlev = ahfgpu_amr_nlevels(G);
  gl = (gridls *)calloc(nlev, sizeof(gridls));
  nodeptr = (nptr **)calloc(nlev, sizeof(nptr *));
  for (l = 0; l < nlev; l++) {
    int64_t  io[4]; double dd[2];
    int32_t *x, *y, *z; float *dens; uint8_t *rf;
    gridls  *g = gl + l;
    if (ahfgpu_amr_level_header(G, l, io, dd)) die("ahfgpu_amr_level_header");
    x = malloc(io[1] * 4); y = malloc(io[1] * 4); z = malloc(io[1] * 4); dens = malloc(io[1] * 4); rf = malloc(io[1]);
    nodeptr[l] = malloc(io[1] * sizeof(nptr));
    if (!x || !y || !z || !dens || !rf || !nodeptr[l]) { fprintf(stderr, "ahf_glue: out of memory for level %d (%ld nodes)\n", l, (long)io[1]); common_terminate(EXIT_FAILURE); }
    if (ahfgpu_amr_level_get(G, l, x, y, z, dens, rf, NULL, NULL, NULL)) die("ahfgpu_amr_level_get");
    g->l1dim = (long unsigned)io[0]; g->spacing = 1.0 / (double)io[0]; g->spacing2 = g->spacing * g->spacing;
    g->critdens = dd[0]; g->masstopartdens = dd[1];
    g->masstodens = dd[1] * (double)simu.no_part / simu.no_vpart;          /* generate_grids.c:66-69 */
    g->size.no_part = (long unsigned)io[2]; g->size.no_nodes = (long unsigned)io[1];
    g->timecounter = global.super_t; g->next = (l < nlev - 1) ? TRUE : FALSE;
    build_level_quads(g, io[1], x, y, z, dens, rf, nodeptr[l]);
    free(x); free(y); free(z); free(dens); free(rf);
  }
  global.fin_l1dim = gl[nlev - 1].l1dim;
  /* node particle lists */
  owner = malloc(n); cell_of = malloc((size_t)nlev * n * sizeof(int32_t));
  if (!owner || !cell_of) { fprintf(stderr, "ahf_glue: out of memory for the particle -> node map (%d levels x %lu particles)\n", nlev, (unsigned long)n); common_terminate(EXIT_FAILURE); }
  if (ahfgpu_amr_particle_levels(G, owner, cell_of, nlev)) die("ahfgpu_amr_particle_levels");
  for (i = 0; i < n; i++) global_info.fst_part[i].ll = NULL;
  for (l = 0; l < nlev; l++) {
    if ((l & 1) == 0) {           /* descending list: head-insert in ascending order */
      for (i = 0; i < n; i++) if (owner[i] == l) { nptr nd = nodeptr[l][cell_of[(size_t)l * n + i]]; partptr q = global_info.fst_part + i; q->ll = nd->ll; nd->ll = q; }
    } else {                      /* ascending list: head-insert in descending order */
      for (i = n; i-- > 0;) if (owner[i] == l) { nptr nd = nodeptr[l][cell_of[(size_t)l * n + i]]; partptr q = global_info.fst_part + i; q->ll = nd->ll; nd->ll = q; }
    }
  }
  for (l = 0; l < nlev; l++) free(nodeptr[l]);
  free(nodeptr); free(owner); free(cell_of);
  global.dom_grid = gl; global.domgrid_no = 0;
  *no_grids = nlev;
  return gl;
}

#endif /* AHFB200_FULL */

void    ahfb200_ll(long unsigned npart, partptr fst_part, gridls *cur_grid) { (void)npart; (void)fst_part; (void)cur_grid; }
void    ahfb200_zero_dens(gridls *g) { (void)g; }
boolean ahfb200_assign_npart(gridls *g) { (void)g; return TRUE; }
boolean ahfb200_gen_AMRhierarchy(gridls **grid_list, int *no_grids) { (void)grid_list; (void)no_grids; return FALSE; }

/* ---- H: src/libahf/ahf_halos.c:504-510 ------------------------------------------------------------------- */
static HALO **pend = NULL;
static long   npend = 0, cappend = 0;

static void flush_halos(void);

void ahfb200_constructHalo(HALO *h)
{
#pragma omp critical(ahfb200_pend)
  {
    if (npend == cappend) { cappend = cappend ? 2 * cappend : 1024; pend = realloc(pend, cappend * sizeof(HALO *)); if (!pend) { fprintf(stderr, "ahf_glue: out of memory\n"); exit(EXIT_FAILURE); } }
    pend[npend++] = h;
  }
#ifndef _OPENMP
  flush_halos();                    /* serial build: no parallel region whose end could trigger the batch */
#endif
}

static int cmp_ptr(const void *a, const void *b)
{
  const HALO *x = *(HALO *const *)a, *y = *(HALO *const *)b;
  return (x < y) ? -1 : (x > y);
}

static void flush_halos(void)
{
  long          i, k, nh = npend;
  double       *ctr, *rad, *scal, *prof;
  int64_t      *seed, *moff, *mem, *poff, nmem = 0, nbin = 0;
  ahfgpu_params p;
  if (nh == 0) return;
  npend = 0;
  if (getenv("AHFB200_VERBOSE")) fprintf(stderr, "ahf_glue: constructing %ld haloes on the device\n", nh);
  qsort(pend, nh, sizeof(HALO *), cmp_ptr);           /* halos[] index order */
  ensure_ctx();
  fill_params(&p);
  if (ahfgpu_set_params(G, &p)) die("ahfgpu_set_params");
  ctr = malloc(3 * nh * sizeof(double)); rad = malloc(nh * sizeof(double)); seed = malloc(nh * sizeof(int64_t));
  for (i = 0; i < nh; i++) {
    ctr[3 * i] = pend[i]->pos.x; ctr[3 * i + 1] = pend[i]->pos.y; ctr[3 * i + 2] = pend[i]->pos.z;
    rad[i] = pend[i]->gatherRad; seed[i] = (int64_t)pend[i]->npart;
  }
  if (ahfgpu_construct_halos(G, nh, ctr, rad, seed)) die("ahfgpu_construct_halos");
  if (ahfgpu_halo_sizes(G, &nmem, &nbin)) die("ahfgpu_halo_sizes");
  scal = malloc(nh * AHFGPU_NSCAL * sizeof(double)); moff = malloc((nh + 1) * sizeof(int64_t)); poff = malloc((nh + 1) * sizeof(int64_t));
  mem = malloc((nmem > 0 ? nmem : 1) * sizeof(int64_t)); prof = malloc((nbin > 0 ? nbin : 1) * AHFGPU_NPROFCOL * sizeof(double));
  if (ahfgpu_halo_fetch(G, scal, moff, mem, poff, prof)) die("ahfgpu_halo_fetch");
#ifdef GAS_PARTICLES
  double *spc = malloc(nh * 64 * sizeof(double)), *psp = malloc((nbin > 0 ? nbin : 1) * 3 * sizeof(double));
  if (ahfgpu_halo_fetch_species(G, spc, psp)) die("ahfgpu_halo_fetch_species");
#endif
  for (i = 0; i < nh; i++) {
    HALO   *h = pend[i];
    double *s = scal + (size_t)AHFGPU_NSCAL * i;
    if (seed[i] == 0) continue;                        /* ahf_halos_sfc.c:122 */
    h->nll = 0; h->ll = NULL;
    h->npart = (unsigned long)s[9];
    h->ipart = (h->npart > 0) ? malloc(h->npart * sizeof(unsigned long)) : NULL;     /* freed with free(), ahf_halos.c:897 */
    for (k = 0; k < (long)h->npart; k++) h->ipart[k] = (unsigned long)mem[moff[i] + k];
    if ((long)s[5] >= simu.AHF_MINPART) {              /* rem_outsideRvir ran at least once (ahf_halos.c:3696) */
      h->M_vir = s[10]; h->R_vir = s[11]; h->ovdens = s[12]; h->Phi0 = s[13];
    }
    if ((long)h->npart >= simu.AHF_MINPART) {
      int     nb = (int)s[57], b;
      double *pr = prof + (size_t)AHFGPU_NPROFCOL * poff[i];
      c_profile(h, nb);                                /* alloc_struct.c:614; freed by dest_profile */
      h->vel.x = s[14]; h->vel.y = s[15]; h->vel.z = s[16]; h->sigV = s[17]; h->v_esc2 = s[18]; h->V2_max = s[19];
      h->R_max = s[20]; h->r2 = s[21]; h->lambda = s[22]; h->lambdaE = s[23]; h->Ekin = s[24]; h->Epot = s[25]; h->SurfP = s[26];
      h->pos_com.x = s[27]; h->pos_com.y = s[28]; h->pos_com.z = s[29]; h->com_offset = s[30];
      h->pos_mbp.x = s[31]; h->pos_mbp.y = s[32]; h->pos_mbp.z = s[33]; h->vel_mbp.x = s[34]; h->vel_mbp.y = s[35]; h->vel_mbp.z = s[36];
      h->mbp_offset = s[37]; h->AngMom.x = s[38]; h->AngMom.y = s[39]; h->AngMom.z = s[40];
      h->axis.x = s[41]; h->axis.y = s[42]; h->axis.z = s[43];
      h->E1.x = s[44]; h->E1.y = s[45]; h->E1.z = s[46]; h->E2.x = s[47]; h->E2.y = s[48]; h->E2.z = s[49];
      h->E3.x = s[50]; h->E3.y = s[51]; h->E3.z = s[52];
      h->fMhires = s[53]; h->cNFW = s[54]; h->cR1 = s[55]; h->R1 = s[56];
      for (b = 0; b < nb; b++) {
#define PR(col) pr[(col) * nb + b]
        h->prof.npart[b] = (unsigned long)PR(0); h->prof.r[b] = PR(1); h->prof.nvpart[b] = PR(2); h->prof.ovdens[b] = PR(3);
        h->prof.dens[b] = PR(4); h->prof.v2_circ[b] = PR(5); h->prof.v_esc2[b] = PR(6); h->prof.sig_v[b] = PR(7);
        h->prof.Ekin[b] = PR(8); h->prof.Epot[b] = PR(9); h->prof.Lx[b] = PR(10); h->prof.Ly[b] = PR(11); h->prof.Lz[b] = PR(12);
        h->prof.axis1[b] = PR(13); h->prof.E1x[b] = PR(14); h->prof.E1y[b] = PR(15); h->prof.E1z[b] = PR(16);
        h->prof.axis2[b] = PR(17); h->prof.E2x[b] = PR(18); h->prof.E2y[b] = PR(19); h->prof.E2z[b] = PR(20);
        h->prof.axis3[b] = PR(21); h->prof.E3x[b] = PR(22); h->prof.E3y[b] = PR(23); h->prof.E3z[b] = PR(24);
#undef PR
      }
#ifdef GAS_PARTICLES
      {                                               /* HaloProfiles' per-species blocks (ahf_halos.c:4712-4715, :5020-5181) */
        int q;
        for (b = 0; b < nb; b++) {
          h->prof.M_gas[b] = psp[3 * poff[i] + 0 * nb + b]; h->prof.M_star[b] = psp[3 * poff[i] + 1 * nb + b]; h->prof.u_gas[b] = psp[3 * poff[i] + 2 * nb + b];
        }
        for (q = 0; q < 2; q++) {
          SPECIESPROP *t = q ? &h->stars_only : &h->gas_only;
          const double *o = spc + 64 * (size_t)i + 32 * q;
          t->npart = (long unsigned)o[0]; t->Mass = o[1]; t->pos_com.x = o[2]; t->pos_com.y = o[3]; t->pos_com.z = o[4];
          t->pos_mbp.x = o[5]; t->pos_mbp.y = o[6]; t->pos_mbp.z = o[7]; t->vel.x = o[8]; t->vel.y = o[9]; t->vel.z = o[10];
          t->lambda = o[11]; t->lambdaE = o[12]; t->AngMom.x = o[13]; t->AngMom.y = o[14]; t->AngMom.z = o[15];
          t->axis.x = o[16]; t->axis.y = o[17]; t->axis.z = o[18];
          t->E1.x = o[19]; t->E1.y = o[20]; t->E1.z = o[21]; t->E2.x = o[22]; t->E2.y = o[23]; t->E2.z = o[24];
          t->E3.x = o[25]; t->E3.y = o[26]; t->E3.z = o[27]; t->Ekin = o[28]; t->Epot = o[29];
        }
      }
#endif
    }
  }
#ifdef GAS_PARTICLES
  free(spc); free(psp);
#endif
  if (getenv("AHFB200_VERBOSE")) { long ok = 0; for (i = 0; i < nh; i++) ok += ((long)pend[i]->npart >= simu.AHF_MINPART); fprintf(stderr, "ahf_glue: %ld haloes with npart >= %d\n", ok, simu.AHF_MINPART); }
  free(ctr); free(rad); free(seed); free(scal); free(moff); free(poff); free(mem); free(prof);
}

/* The halo loop is `#pragma omp parallel for` (ahf_halos.c:504-510): gcc lowers it to GOMP_parallel(outlined body).  This definition
 * takes precedence over libgomp's for the whole program; it runs the region through the real entry point and, when the region that
 * collected HALO pointers has ended, runs the batched device pass -- before the first serial statement that reads a HALO (the
 * sub-halo re-hash, ahf_halos.c:553), whatever that statement is and whether or not VERBOSE is defined. */
#ifdef _OPENMP
void GOMP_parallel(void (*fn)(void *), void *data, unsigned num_threads, unsigned flags)
{
  static void (*real)(void (*)(void *), void *, unsigned, unsigned) = NULL;
  if (!real) {
    *(void **)(&real) = dlsym(RTLD_NEXT, "GOMP_parallel");
    if (!real) { fprintf(stderr, "ahf_glue: libgomp's GOMP_parallel not found\n"); exit(EXIT_FAILURE); }
  }
  real(fn, data, num_threads, flags);
  if (npend > 0 && !omp_in_parallel()) flush_halos();
}
#endif

/* safety net: the catalogue writers must never see haloes the device pass has not filled */
static void check_flushed(void)
{
  if (npend > 0) { fprintf(stderr, "ahf_glue: %ld haloes were collected but never constructed on the device\n", npend); _Exit(EXIT_FAILURE); }
}
__attribute__((constructor)) static void ahfb200_register(void) { T0 = T_LAST = now_s(); atexit(check_flushed); atexit(tprint); warm_start(); }


#ifdef AHFB200_FULL
/* ---- main.c:657 / :662 -- ahf_gridinfo + ahf_halos ---------------------------------------------------------------------------------------
 * ahf_gridinfo's colouring runs on the device inside ahfb200_halos (ahfgpu_amr_patch_stats labels a level before it reduces it); what is
 * left of it here is the level count and the log lines. */
#include <math.h>
#include <time.h>
#include "libutility/cosmology.h"
extern double u_fac;                                       /* src/libahf/ahf_halos.c:163 */

/* ---- startrun.c:373 -> io_file_readpart -> io_gadget_readpart (libio/io_file.c:438, compiled with -Dio_gadget_readpart=ahfb200_gadget_readpart)
 * NEXT-4 of SURVEY 8f in the program: single-file GADGET snapshots without MASS / U blocks go through ahfgpu_ingest_gadget (three bulk reads,
 * unit scaling on the device); the file object is left as io_gadget_readpart_raw + io_gadget_scale_particles leave it (io_gadget.c:427-568,
 * :857-995), so that startrun's io_file_get calls (boxsize, pmass, no_vpart, weights, species) return the reference's values.  Everything
 * else -- and every file the ingest refuses -- is read by the reference's own reader.  The host AoS keeps only what the log file prints
 * (first / last particle); nobody else reads it in this build. */
#if !defined(MULTIMASS) && !defined(GAS_PARTICLES) && !defined(METALHACK)
#include "libio/io_gadget.h"
#include "libio/io_gadget_def.h"

uint64_t ahfb200_gadget_readpart(io_logging_t log, io_gadget_t f, uint64_t pskip, uint64_t pread, io_file_strg_struct_t strg)
{
  double    info[24], oldw = 0.0;
  uint64_t  n, i, first = 0;
  uint64_t *ids;
  int       t, d, k;
  if (getenv("AHFB200_NO_INGEST") || f == NULL || f->header == NULL || pskip != 0 || pread < f->no_part || f->multimass != 0
      || f->header->np[0] > 0 || strg.bytes_float != sizeof(float) || strg.weight.val != NULL || strg.u.val != NULL)
    return io_gadget_readpart(log, f, pskip, pread, strg);
  n = f->no_part;
  if (WARM_ON) ahfgpu_ingest_prefetch(f->fname);     /* read while the helper thread is still creating the CUDA context; a refusal shows below */
  tmark("ingest_prefetch");
  ensure_ctx();
  ids = malloc((n > 0 ? n : 1) * sizeof(uint64_t));
  if (!ids || ahfgpu_ingest_gadget(G, f->fname, f->posscale, f->weightscale, ids, info) || (uint64_t)info[0] != n) {
    io_logging_msg(log, INT32_C(1), "bulk ingest not used (%s): reading with io_gadget_readpart", ahfgpu_last_error());
    free(ids);
    return io_gadget_readpart(log, f, pskip, pread, strg);
  }
  tmark("ingest_gadget");
  /* extreme positions, box check, shift and scales (local_get_block_pos :1320-1347, io_gadget_scale_particles :873-935) */
  for (d = 0; d < 3; d++) { f->minpos[d] = info[16 + d]; f->maxpos[d] = info[19 + d]; simu.pos_shift[d] = info[6 + d]; }
  f->header->boxsize = info[22];
  simu.pos_scale = info[9];
  io_logging_msg(log, INT32_C(4), "Extreme positions: xmin = %g  xmax = %g", f->minpos[0], f->maxpos[0]);
  io_logging_msg(log, INT32_C(4), "                   ymin = %g  ymax = %g", f->minpos[1], f->maxpos[1]);
  io_logging_msg(log, INT32_C(4), "                   zmin = %g  zmax = %g", f->minpos[2], f->maxpos[2]);
  io_logging_msg(log, INT32_C(4), "Applying shift: (%g, %g, %g)", info[6], info[7], info[8]);
  io_logging_msg(log, INT32_C(3), "Scaling by:  positions:  %g", info[9]);
  io_logging_msg(log, INT32_C(3), "             velocities: %g", info[10]);
  /* weights from the header (local_get_block_mass :1636-1690 without a MASS block): the sum runs particle by particle, as there, so that
   * no_vpart = sumweight / mmass has the reference's bits */
  f->sumweight = 0.0; f->no_species = 0;
  for (t = 0; t < 6; t++) {
    const double w = f->header->massarr[t];
    const uint64_t cnt = (uint64_t)(f->header->np[t] > 0 ? f->header->np[t] : 0);
    if (cnt == 0) continue;
    if (w > oldw || w < oldw) {
      f->no_species++; oldw = w;
      if (w < f->minweight) f->minweight = w;
      if (w > f->maxweight) f->maxweight = w;
      if (t == 1 && w < f->mmass) f->mmass = w;      /* only halo particles (type 1) set mmass; first >= np[0] = 0 here */
    }
    for (i = 0; i < cnt; i++) f->sumweight += w;
    first += cnt;
  }
  /* what the log file shows of the particle array */
  for (k = 0; k < 2 && n > 0; k++) {
    const uint64_t j = k ? n - 1 : 0;
    float p3[3], m3[3];
    if (ahfgpu_input_peek(G, j, p3, m3)) die("ahfgpu_input_peek");
    *(float *)((char *)strg.posx.val + j * strg.posx.stride) = p3[0]; *(float *)((char *)strg.posy.val + j * strg.posy.stride) = p3[1];
    *(float *)((char *)strg.posz.val + j * strg.posz.stride) = p3[2];
    *(float *)((char *)strg.momx.val + j * strg.momx.stride) = m3[0]; *(float *)((char *)strg.momy.val + j * strg.momy.stride) = m3[1];
    *(float *)((char *)strg.momz.val + j * strg.momz.stride) = m3[2];
    if (strg.id.val != NULL) {
      if (strg.bytes_int == 8) *(uint64_t *)((char *)strg.id.val + j * strg.id.stride) = ids[j];
      else *(uint32_t *)((char *)strg.id.val + j * strg.id.stride) = (uint32_t)ids[j];
    }
  }
  G_ids = ids; G_ingested = 1;
  return pread < n ? pread : n;
}
#endif

void ahfb200_gridinfo(gridls *grid_list, int curgrid_no)
{
  (void)grid_list;
  ahf.no_grids = curgrid_no - global.domgrid_no;          /* ahf_gridinfo.c:123; min_ref is subtracted in ahfb200_halos */
  fprintf(io.logfile, "################## ahf_gridinfo ###################\n");
  fprintf(io.logfile, "Number of grids           = %d\n", curgrid_no + 1);
  fflush(io.logfile);
}

void ahfb200_halos(gridls *grid_list)
{
  ahfgpu_params       p;
  ahfgpu_catalogue_in cat;
  char     fprefix[MAXSTRING], file_no[MAXSTRING];
  double   a = global.a, a3 = a * a * a, omega, ovlim, rho_crit, rho_b, rho_vir, refine_ovdens = 0.0;
  int      nlev = ahfgpu_amr_nlevels(G), i, start = 1, nref, lev;
  int64_t *niso, rows = 0, nh = 0, nmem = 0, nbin = 0, r0, k;
  double  *stats, *ctr, *rad, *scal, *prof, *spc = NULL, *psp = NULL;
  int64_t *seed, *moff, *mem, *poff, *hsoff;
  int32_t *hhost, *hlev, *hsub;
  uint64_t *pid;
  float   *pu = NULL;
  uint64_t n = global_info.no_part;
  (void)grid_list;
  /* conversion factors and cosmology, as ahf_halos.c:196-222 */
  r_fac = simu.boxsize * global.a; x_fac = simu.boxsize; v_fac = simu.boxsize / simu.t_unit / global.a; m_fac = simu.pmass;
  u_fac = pow2(simu.boxsize / simu.t_unit); rho_fac = simu.pmass / pow3(simu.boxsize); phi_fac = Grav * simu.pmass / (simu.boxsize * global.a);
  omega = calc_omega(a); ovlim = calc_virial(a); Hubble = calc_Hubble(a);
  rho_crit = a3 * calc_rho_crit(a); rho_b = omega * rho_crit; rho_vir = a3 * calc_rho_vir(a);
  global.ovlim = ovlim; global.rho_b = rho_b; global.rho_vir = rho_vir;
  /* first coloured level, as ahf_gridinfo.c:147-190 */
  for (i = 0; i <= ahf.no_grids; i++) {
    double fl1dim = (double)global.dom_grid[i].l1dim, refine_len = (double)(simu.boxsize / fl1dim), refine_vol = pow3(refine_len);
    refine_ovdens = (simu.Nth_ref * (simu.pmass * simu.med_weight) / refine_vol) / rho_vir;
    fprintf(io.logfile, "l1dim = %16.0f refine_ovdens = %16.4g ovlim = %16.4g\n", fl1dim, refine_ovdens, ovlim);
    if (refine_ovdens < ovlim) start = i;
  }
  global.max_ovdens = refine_ovdens;
  ahf.min_ref = start + AHF_MIN_REF_OFFSET;
  ahf.no_grids = ahf.no_grids - ahf.min_ref + 1;
  fprintf(io.logfile, "min_ref = %d    (ahf_nogrids = %d)\n", ahf.min_ref, ahf.no_grids);
  fprintf(io.logfile, "#################### ahf_halos ####################\n");
  fflush(io.logfile);
  if (ahf.no_grids <= 0) return;                           /* ahf_halos.c:195 */
  snprintf(fprefix, MAXSTRING, "%s.", global_io.params->outfile_prefix);
  sprintf(file_no, "z%.3f", fabs(global.z));
  strcat(fprefix, file_no);
  fill_params(&p);
  if (ahfgpu_set_params(G, &p)) die("ahfgpu_set_params");
  /* per-refinement tables of the coloured levels (RefCentre), then tree + seeds (analyseRef, spatialRef2halos) */
  tmark("mesh_to_halos");
  timing.RefCentre -= time(NULL);
  nref = nlev - ahf.min_ref;
  niso = calloc(nref > 0 ? nref : 1, sizeof(int64_t));
  for (lev = ahf.min_ref; lev < nlev; lev++) {
    if (ahfgpu_amr_patch_stats(G, lev, niso + (lev - ahf.min_ref), NULL, 0)) die("ahfgpu_amr_patch_stats");
    rows += niso[lev - ahf.min_ref];
  }
  stats = malloc((rows > 0 ? rows : 1) * 18 * sizeof(double));
  for (lev = ahf.min_ref, r0 = 0; lev < nlev; lev++) {
    int64_t q = 0;
    if (ahfgpu_amr_patch_stats(G, lev, &q, stats + 18 * r0, rows - r0)) die("ahfgpu_amr_patch_stats");
    r0 += q;
  }
  timing.RefCentre += time(NULL);
  tmark("patch_stats");
  timing.analyseRef -= time(NULL);
  ctr = malloc((rows + 1) * 3 * sizeof(double)); rad = malloc((rows + 1) * sizeof(double)); seed = malloc((rows + 1) * sizeof(int64_t));
  hhost = malloc((rows + 1) * sizeof(int32_t)); hlev = malloc((rows + 1) * sizeof(int32_t)); hsoff = malloc((rows + 2) * sizeof(int64_t));
  hsub = malloc((rows + 1) * sizeof(int32_t));
  if (ahfgpu_tree_halos_ex(nref, niso, stats, simu.MaxGatherRad / simu.boxsize, NULL, NULL, NULL, NULL, 0, &nh, ctr, rad, seed, hhost, rows + 1,
                           hlev, hsoff, hsub, rows + 1)) die("ahfgpu_tree_halos_ex");
  timing.analyseRef += time(NULL);
  tmark("tree_halos");
  simu.no_halos = (int)nh;
  fprintf(io.logfile, "\nConstructing Halos (%ld)\n", (long)nh);
  fflush(io.logfile);
  /* halo pass */
  timing.ahf_halos_sfc_constructHalo -= time(NULL);
  if (ahfgpu_construct_halos(G, nh, ctr, rad, seed)) die("ahfgpu_construct_halos");
  if (ahfgpu_halo_sizes(G, &nmem, &nbin)) die("ahfgpu_halo_sizes");
  scal = malloc((nh > 0 ? nh : 1) * AHFGPU_NSCAL * sizeof(double)); moff = malloc((nh + 1) * sizeof(int64_t)); poff = malloc((nh + 1) * sizeof(int64_t));
  mem = malloc((nmem > 0 ? nmem : 1) * sizeof(int64_t)); prof = malloc((nbin > 0 ? nbin : 1) * AHFGPU_NPROFCOL * sizeof(double));
  if (ahfgpu_halo_fetch(G, scal, moff, mem, poff, prof)) die("ahfgpu_halo_fetch");
#ifdef GAS_PARTICLES
  spc = malloc((nh > 0 ? nh : 1) * 64 * sizeof(double)); psp = malloc((nbin > 0 ? nbin : 1) * 3 * sizeof(double));
  if (ahfgpu_halo_fetch_species(G, spc, psp)) die("ahfgpu_halo_fetch_species");
#endif
  timing.ahf_halos_sfc_constructHalo += time(NULL);
  tmark("construct_fetch");
  /* re-hash, ordering, catalogues */
  timing.ahf_io -= time(NULL);
  pid = malloc((n > 0 ? n : 1) * sizeof(uint64_t));
  if (G_ingested) {                                          /* IDs of the ingested file through the sorted-offset -> input-index permutation */
    uint32_t *perm = malloc((n > 0 ? n : 1) * sizeof(uint32_t));
    if (!perm || ahfgpu_particle_ids(G, perm)) die("ahfgpu_particle_ids");
#pragma omp parallel for schedule(static)
    for (k = 0; k < (int64_t)n; k++) pid[k] = G_ids[perm[k]];
    free(perm);
  } else
    for (k = 0; k < (int64_t)n; k++) pid[k] = (uint64_t)global_info.fst_part[k].id;
#ifdef GAS_PARTICLES
  pu = malloc((n > 0 ? n : 1) * sizeof(float));
  for (k = 0; k < (int64_t)n; k++) pu[k] = (float)global_info.fst_part[k].u;
#endif
  memset(&cat, 0, sizeof(cat));
  cat.nhalo = nh; cat.scal = scal; cat.pos3 = ctr; cat.member_off = moff; cat.members = mem; cat.prof_off = poff; cat.prof = prof;
  cat.species = spc; cat.prof_species = psp; cat.host = hhost; cat.host_level = hlev; cat.sub_off = hsoff; cat.sub = hsub;
  cat.part_id = pid; cat.part_u = pu; cat.part_weight = NULL;
  cat.x_fac = x_fac; cat.r_fac = r_fac; cat.v_fac = v_fac; cat.m_fac = m_fac; cat.rho_fac = rho_fac; cat.phi_fac = phi_fac; cat.u_fac = u_fac;
  cat.rho_vir = global.rho_vir; cat.pmass = simu.pmass; cat.min_part = simu.AHF_MINPART;
#ifdef GAS_PARTICLES
  cat.flags = 1;
#endif
  tmark("id_copy");
  if (ahfgpu_catalogue_write(fprefix, &cat, NULL, NULL, NULL)) die("ahfgpu_catalogue_write");
  timing.ahf_io += time(NULL);
  tmark("catalogue_write");
  free(niso); free(stats); free(ctr); free(rad); free(seed); free(hhost); free(hlev); free(hsoff); free(hsub);
  free(scal); free(moff); free(poff); free(mem); free(prof); free(spc); free(psp); free(pid); free(pu);
}
#endif /* AHFB200_FULL */

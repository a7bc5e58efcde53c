/*
 * ahf_oracle_mesh.c -- TEST INFRASTRUCTURE (see ahf_oracle.h).  Single-threaded CPU restatement of
 *   K1  Hilbert keys                       reference src/libsfc/hilbert.c:139-243, hilbert_util.c:69-92
 *   D2  particle -> domain cell            src/libamr_serial/lltools.c:32-92
 *   D3/D4 TSC number-density deposit       src/libamr_serial/density.c:238-438, :443-492
 *   D5  27-neighbour search semantics      src/libamr_serial/get_nnodes.c:41-51, :459-862
 *   F1/F2 refinement flags + child set     src/libamr_serial/refine_grid.c:113-138, :155-896
 *   R1  relink                             src/libamr_serial/relink.c:31-288
 *   L1  level loop                         src/libamr_serial/generate_grids.c:126-432, src/main.c:616-648
 *
 * The reference stores a level as run-length "quads" (pquad -> cquad -> nquad -> node).  This
 * restatement stores a level as the (z,y,x)-sorted list of its cells plus ONE extra bit per cell,
 * `xbreak`: "the reference's x-run (nquad) ends after this cell although cell x+1 may exist".  That
 * happens exactly after the second node of a ghost pair (refine_grid.c:231-250) and is what makes a
 * spatially adjacent node invisible to the reference's neighbour search (get_nnodes.c:465-508).
 * z-runs and y-runs of the reference are always maximal, so they follow from the cell set alone.
 * Summation order of the deposit (node traversal order x list order x 27 offsets) is reproduced, so
 * `dens` is expected to be BIT-identical to the reference's.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "ahf_oracle.h"

#define MIN_NNODES 125          /* src/param.h:54  */
#define CRITMULTI  8.0          /* src/param.h:118 */

/* ============================================================================================== */
/* K1: Hilbert keys, 3 dimensions                                                                  */
/* ============================================================================================== */
static inline unsigned rotr3(unsigned v, unsigned r) { return ((v >> r) | (v << (3 - r))) & 7u; }
static inline unsigned rotl3(unsigned v, unsigned r) { return ((v << r) | (v >> (3 - r))) & 7u; }
static inline unsigned next_rotation(unsigned rot, unsigned bits)
{
  unsigned low = bits & (0u - bits) & 3u;    /* lowest set bit if it is bit 0 or bit 1, else 0 */
  rot += (low == 1u) ? 1u : (low == 2u) ? 2u : 0u;
  rot += 1u;
  if (rot >= 3u) rot -= 3u;
  return rot;
}

/* index of integer cell (cx,cy,cz), `bits` bits per dimension; x is dimension 0 (hilbert.c:197-243) */
static uint64_t hilbert_index3(uint64_t cx, uint64_t cy, uint64_t cz, unsigned bits)
{
  uint64_t inter = 0, index = 0, mask;
  unsigned j, rot = 0, flip = 0;
  int      b;
  for (j = 0; j < bits; j++) {
    inter |= ((cx >> j) & 1u) << (3 * j);
    inter |= ((cy >> j) & 1u) << (3 * j + 1);
    inter |= ((cz >> j) & 1u) << (3 * j + 2);
  }
  if (bits > 1) {
    inter ^= inter >> 3;
    for (b = 3 * (int)bits - 3; b >= 0; b -= 3) {
      unsigned g = (unsigned)((inter >> b) & 7u);
      g     = rotr3(flip ^ g, rot);
      index = (index << 3) | g;
      flip  = 1u << rot;
      rot   = next_rotation(rot, g);
    }
    mask = 0;
    for (j = 1; j < bits; j++) mask |= (uint64_t)1 << (3 * j - 1);
    index ^= mask;
  } else {
    index = inter;
  }
  for (j = 1; j < 3 * bits; j *= 2) index ^= index >> j;
  return index;
}

void orc_hilbert_coords(uint64_t index, unsigned bits, uint32_t out[3])
{
  uint64_t coords = 0;
  unsigned j;
  if (bits > 1) {
    uint64_t nth = 0;
    unsigned rot = 0, flip = 0;
    int      b;
    for (j = 0; j < bits; j++) nth |= (uint64_t)1 << (3 * j);
    index ^= (index ^ nth) >> 1;
    for (b = 3 * (int)bits - 3; b >= 0; b -= 3) {
      unsigned g = (unsigned)((index >> b) & 7u);
      coords = (coords << 3) | (rotl3(g, rot) ^ flip);
      flip   = 1u << rot;
      rot    = next_rotation(rot, g);
    }
    for (j = 3; j < 3 * bits; j *= 2) coords ^= coords >> j;
  } else {
    coords = index ^ (index >> 1);
  }
  out[0] = out[1] = out[2] = 0;
  for (j = 0; j < bits; j++) {
    out[0] |= (uint32_t)((coords >> (3 * j)) & 1u) << j;
    out[1] |= (uint32_t)((coords >> (3 * j + 1)) & 1u) << j;
    out[2] |= (uint32_t)((coords >> (3 * j + 2)) & 1u) << j;
  }
}

/* hilbert_util.c:69-92: scale to 2^bits cells, truncate, clamp the single overflow value */
uint64_t orc_hilbert_key(double x, double y, double z, unsigned bits)
{
  uint64_t max = (uint64_t)1 << bits, c[3];
  c[0] = (uint64_t)trunc(x * (double)max); if (c[0] == max) c[0] = max - 1;
  c[1] = (uint64_t)trunc(y * (double)max); if (c[1] == max) c[1] = max - 1;
  c[2] = (uint64_t)trunc(z * (double)max); if (c[2] == max) c[2] = max - 1;
  return hilbert_index3(c[0], c[1], c[2], bits);
}

uint64_t orc_hilbert_key_grid(uint32_t x, uint32_t y, uint32_t z, unsigned bits)
{
  return hilbert_index3(x, y, z, bits);
}

void orc_hilbert_keys(const float *pos3, int64_t n, unsigned bits, uint64_t *keys)
{
  int64_t i;
  for (i = 0; i < n; i++)   /* main.c:343-350: float coordinates are widened to double first */
    keys[i] = orc_hilbert_key((double)pos3[3 * i], (double)pos3[3 * i + 1], (double)pos3[3 * i + 2], bits);
}

typedef struct { uint64_t k; int64_t i; } keyidx;
static int cmp_keyidx(const void *a, const void *b)
{
  const keyidx *x = a, *y = b;
  if (x->k != y->k) return (x->k < y->k) ? -1 : 1;
  return (x->i < y->i) ? -1 : (x->i > y->i);
}
void orc_argsort_keys(const uint64_t *keys, int64_t n, int64_t *order)
{
  keyidx *t = malloc(sizeof(keyidx) * (size_t)(n > 0 ? n : 1));
  int64_t i;
  for (i = 0; i < n; i++) { t[i].k = keys[i]; t[i].i = i; }
  qsort(t, (size_t)n, sizeof(keyidx), cmp_keyidx);
  for (i = 0; i < n; i++) order[i] = t[i].i;
  free(t);
}

/* ============================================================================================== */
/* level storage                                                                                   */
/* ============================================================================================== */
typedef struct {
  int64_t  L, ncell;
  int      dense;                 /* domain level: cell index = (z*L+y)*L+x                        */
  int32_t *x, *y, *z;
  uint8_t *xbreak;                /* the reference's nquad run ends after this cell                */
  float   *dens;
  int64_t *head;                  /* first particle of the cell's linked list, -1 = empty          */
  uint8_t *interior;              /* all 27 neighbours visible to the reference (test_tsc)         */
  uint8_t *mark;                  /* refinement outcome: 0 none, 1 refined, 2 ghost pair           */
  uint64_t *hkey; int64_t *hval; uint64_t hmask;      /* open-addressing hash (x,y,z) -> cell      */
  double   critdens, masstopartdens;
  int64_t  npart_flag, npart_final;
  int32_t *cnt_flag;  int64_t *plist_flag;            /* list snapshot at deposit/flag time        */
  int32_t *cnt_final; int64_t *plist_final;
} olevel;

struct orc_hier {
  int      nlev;
  olevel **lev;
  int64_t  n;
  int64_t *next;                  /* particle linked list (one membership at a time, like part.ll) */
  const float *pos;
};

static inline uint64_t pack3(int64_t L, int64_t x, int64_t y, int64_t z) { return ((uint64_t)z * (uint64_t)L + (uint64_t)y) * (uint64_t)L + (uint64_t)x; }

static uint64_t mix64(uint64_t k)
{
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

static void level_build_hash(olevel *lv)
{
  uint64_t cap = 16;
  int64_t  i;
  while (cap < (uint64_t)lv->ncell * 2 + 2) cap <<= 1;
  lv->hmask = cap - 1;
  lv->hkey  = malloc(sizeof(uint64_t) * cap);
  lv->hval  = malloc(sizeof(int64_t) * cap);
  memset(lv->hkey, 0xff, sizeof(uint64_t) * cap);
  for (i = 0; i < lv->ncell; i++) {
    uint64_t k = pack3(lv->L, lv->x[i], lv->y[i], lv->z[i]);
    uint64_t s = mix64(k) & lv->hmask;
    while (lv->hkey[s] != UINT64_MAX) s = (s + 1) & lv->hmask;
    lv->hkey[s] = k; lv->hval[s] = i;
  }
}

/* geometric lookup, no periodic wrap: -1 when the cell is not part of the level */
static inline int64_t cell_at(const olevel *lv, int64_t x, int64_t y, int64_t z)
{
  uint64_t k, s;
  if (x < 0 || y < 0 || z < 0 || x >= lv->L || y >= lv->L || z >= lv->L) return -1;
  if (lv->dense) return (int64_t)pack3(lv->L, x, y, z);
  k = pack3(lv->L, x, y, z);
  s = mix64(k) & lv->hmask;
  while (lv->hkey[s] != UINT64_MAX) {
    if (lv->hkey[s] == k) return lv->hval[s];
    s = (s + 1) & lv->hmask;
  }
  return -1;
}

static olevel *level_alloc(int64_t L, int64_t ncell, int dense)
{
  olevel *lv = calloc(1, sizeof(olevel));
  size_t  m  = (size_t)(ncell > 0 ? ncell : 1);
  int64_t i;
  lv->L = L; lv->ncell = ncell; lv->dense = dense;
  lv->x = malloc(sizeof(int32_t) * m); lv->y = malloc(sizeof(int32_t) * m); lv->z = malloc(sizeof(int32_t) * m);
  lv->xbreak   = calloc(m, 1);
  lv->dens     = malloc(sizeof(float) * m);
  lv->head     = malloc(sizeof(int64_t) * m);
  lv->interior = calloc(m, 1);
  lv->mark     = calloc(m, 1);
  for (i = 0; i < ncell; i++) lv->head[i] = -1;
  return lv;
}

static void level_free(olevel *lv)
{
  if (!lv) return;
  free(lv->x); free(lv->y); free(lv->z); free(lv->xbreak); free(lv->dens); free(lv->head); free(lv->interior);
  free(lv->mark); free(lv->hkey); free(lv->hval);
  free(lv->cnt_flag); free(lv->plist_flag); free(lv->cnt_final); free(lv->plist_final);
  free(lv);
}

/* ============================================================================================== */
/* D5: neighbour visibility (get_nnodes.c:459-862)                                                 */
/* ============================================================================================== */
/* x+1 as seen from cell c (search_TSCcolumn, get_nnodes.c:465-485): same run, else periodic image
 * when c sits at L-1 and the row's first run starts at 0, else missing. */
static int64_t nb_xplus(const olevel *lv, int64_t c)
{
  int64_t x = lv->x[c], y = lv->y[c], z = lv->z[c], r;
  if (!lv->xbreak[c]) { r = cell_at(lv, x + 1, y, z); if (r >= 0) return r; }
  if (x == lv->L - 1) return cell_at(lv, 0, y, z);
  return -1;
}
/* x-1 (get_nnodes.c:490-508): same run, else periodic image when c sits at 0 and the row's last run
 * ends at L. */
static int64_t nb_xminus(const olevel *lv, int64_t c)
{
  int64_t x = lv->x[c], y = lv->y[c], z = lv->z[c], r;
  r = cell_at(lv, x - 1, y, z);
  if (r >= 0 && !lv->xbreak[r]) return r;
  if (x == 0) return cell_at(lv, lv->L - 1, y, z);
  return -1;
}

/* nb[k][j][i], k=z offset, j=y offset, i=x offset (index 1 = the cell itself); -1 = not visible.
 * Order of tests follows get_TSCnodes: plane middle first, then row middle, then the x pair. */
static void neighbours27(const olevel *lv, int64_t c, int64_t nb[3][3][3])
{
  int64_t L = lv->L, x = lv->x[c], y = lv->y[c], z = lv->z[c];
  int     k, j;
  for (k = 0; k < 3; k++) {
    int64_t zz = z + k - 1, pm;
    if (zz < 0) zz = L - 1; else if (zz >= L) zz = 0;      /* wrap only ever applies at the faces */
    pm = (k == 1) ? c : cell_at(lv, x, y, zz);
    for (j = 0; j < 3; j++) {
      int64_t yy = y + j - 1, rm;
      if (yy < 0) yy = L - 1; else if (yy >= L) yy = 0;
      if (pm < 0) { nb[k][j][0] = nb[k][j][1] = nb[k][j][2] = -1; continue; }
      rm = (j == 1) ? pm : cell_at(lv, x, yy, zz);
      if (rm < 0) { nb[k][j][0] = nb[k][j][1] = nb[k][j][2] = -1; continue; }
      nb[k][j][1] = rm;
      nb[k][j][0] = nb_xminus(lv, rm);
      nb[k][j][2] = nb_xplus(lv, rm);
    }
  }
}

static int all27(int64_t nb[3][3][3])
{
  int k, j, i;
  for (k = 0; k < 3; k++) for (j = 0; j < 3; j++) for (i = 0; i < 3; i++) if (nb[k][j][i] < 0) return 0;
  return 1;
}

static void level_compute_interior(olevel *lv)
{
  int64_t c, nb[3][3][3];
  for (c = 0; c < lv->ncell; c++) {
    if (lv->dense) { lv->interior[c] = 1; continue; }       /* periodic full grid: everything is visible */
    neighbours27(lv, c, nb);
    lv->interior[c] = (uint8_t)all27(nb);
  }
}

/* ============================================================================================== */
/* D3 + D4: zero_dens + assign_npart                                                               */
/* ============================================================================================== */
static double f1mod1(double v)            /* specific.c:120-129, y == 1 */
{
  if (v >= 2.0) return v - 2.0;
  if (v >= 1.0) return v - 1.0;
  return v;
}

static void level_deposit(orc_hier *h, olevel *lv)
{
  const double L = (double)lv->L, shift = 0.5 / L, m2d = lv->masstopartdens;
  int64_t c, p, nb[3][3][3], npart = 0;
  int     d, k, j, i;
  for (c = 0; c < lv->ncell; c++) lv->dens[c] = (float)(-1.0);   /* density.c:480, simu.mean_dens = 1 */
  for (c = 0; c < lv->ncell; c++) {
    double ctr[3];
    if (lv->head[c] < 0) continue;                               /* neighbour search has no side effect */
    ctr[0] = f1mod1(((double)lv->x[c] / L + shift) + 1.0);
    ctr[1] = f1mod1(((double)lv->y[c] / L + shift) + 1.0);
    ctr[2] = f1mod1(((double)lv->z[c] / L + shift) + 1.0);
    neighbours27(lv, c, nb);
    for (p = lv->head[c]; p >= 0; p = h->next[p]) {
      double w[3][3];
      npart++;
      for (d = 0; d < 3; d++) {
        double xp = (double)h->pos[3 * p + d];
        double s  = (xp - ctr[d]) * L;
        if (fabs(s) > 0.5 * L) {                                /* density.c:347-354 periodic image */
          double a = f1mod1((xp + 0.5) + 1.0);
          double b = f1mod1((ctr[d] + 0.5) + 1.0);
          s = (a - b) * L;
        }
        w[1][d] = 0.75 - s * s;                                 /* density.c:357-359 */
        w[0][d] = ((0.5 - s) * (0.5 - s)) / 2;
        w[2][d] = ((0.5 + s) * (0.5 + s)) / 2;
      }
      for (k = 0; k < 3; k++) for (j = 0; j < 3; j++) for (i = 0; i < 3; i++)
        if (nb[k][j][i] >= 0) lv->dens[nb[k][j][i]] += m2d * w[k][2] * w[j][1] * w[i][0];   /* density.c:398 */
    }
  }
  lv->npart_flag = npart;
}

static void snapshot_lists(const orc_hier *h, const olevel *lv, int32_t **cnt, int64_t **plist, int64_t *ntot)
{
  int64_t c, p, n = 0, o = 0;
  for (c = 0; c < lv->ncell; c++) for (p = lv->head[c]; p >= 0; p = h->next[p]) n++;
  *cnt   = malloc(sizeof(int32_t) * (size_t)(lv->ncell > 0 ? lv->ncell : 1));
  *plist = malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  for (c = 0; c < lv->ncell; c++) {
    int32_t k = 0;
    for (p = lv->head[c]; p >= 0; p = h->next[p]) { (*plist)[o++] = p; k++; }
    (*cnt)[c] = k;
  }
  *ntot = n;
}

/* ============================================================================================== */
/* F1: test_node (refine_grid.c:113-138)                                                           */
/* ============================================================================================== */
static int test_node(const olevel *lv, int64_t c)
{
  int64_t nb[3][3][3];
  int     k, j, i;
  double  thr = lv->critdens - 1.0;                      /* critdens - simu.mean_dens */
  if (lv->dense) neighbours27(lv, c, nb);
  else { if (!lv->interior[c]) return 0; neighbours27(lv, c, nb); }
  if (!all27(nb)) return 0;
  for (k = 0; k < 3; k++) for (j = 0; j < 3; j++) for (i = 1; i < 3; i++)
    if (lv->dens[nb[k][j][i]] >= thr) return 1;
  return 0;
}

/* ============================================================================================== */
/* F2: which coarse cells spawn children (ref_pquad / ref_cquad / ref_nquad)                       */
/* ============================================================================================== */
typedef struct { int64_t z; int64_t row0, row1; } oplane;            /* rows [row0,row1) of the row table */
typedef struct { int64_t y, z; int64_t c0, c1; } orow;               /* cells [c0,c1)                     */

/* one reference x-run (nquad) [a,b) of cell indices inside a row; `first_c`/`last_c` delimit the whole
 * row so that the periodic start can look at the row's last run.  Restates ref_nquad for one run.   */
static int refine_xrun(olevel *lv, int64_t a, int64_t b, int64_t row_c0, int64_t row_c1, int has_next_run)
{
  int64_t L = lv->L, c;
  int     state = 0, refever = 0, xoffset = 1;
  if (lv->x[a] == 0) {
    /* refine_grid.c:169-199: row starts on the face; periodic partner = last node of the row's last run */
    int64_t last = row_c1 - 1;
    if (lv->x[last] == L - 1) {
      if (test_node(lv, last)) { state = 1; refever = 1; }
      xoffset = 0;
    }
  }
  (void)row_c0;
  for (c = a + xoffset; c < b - 1; c++) {                       /* :212-251 all but the last node */
    if (test_node(lv, c)) {
      state = 1; refever = 1;
      lv->mark[c] = 1;
    } else if (state) {
      state = 0;
      if (lv->interior[c] || lv->dense) lv->mark[c] = 2;        /* ghost pair only off the edge */
    }
  }
  if (has_next_run) {
    /* CASE 1 (:256-270): edge node of a run that is followed by another one is never refined */
  } else if (lv->x[b - 1] == L - 1 && b - 1 >= a + 0 && test_node(lv, b - 1)) {
    /* CASE 2 (:273-291): last node on the box face with a visible periodic partner */
    refever = 1;
    lv->mark[b - 1] = 1;
  }
  return refever;
}

/* restates ref_nquad over all linked runs of one row: returns refever||refined */
static int refine_row(olevel *lv, const orow *r)
{
  int64_t a = r->c0, c;
  int     any = 0;
  while (a < r->c1) {
    int64_t b = a + 1;
    while (b < r->c1 && !lv->xbreak[b - 1] && lv->x[b] == lv->x[b - 1] + 1) b++;
    (void)c;
    if (refine_xrun(lv, a, b, r->c0, r->c1, b < r->c1)) any = 1;
    a = b;
  }
  return any;
}

/* offsets of ref_cquad / ref_pquad (refine_grid.c:346-398, :645-696) */
static void run_offsets(int wrapped, int64_t r0, int64_t r1, int64_t L, int *low, int *up)
{
  if (wrapped) {
    if (r0 == 0 && r1 == L)      { *low = 0; *up = 0; }
    else if (r0 == 0)            { *low = 0; *up = -1; }
    else if (r1 == L)            { *low = 1; *up = 0; }
    else                         { *low = 1; *up = -1; }
  } else { *low = 1; *up = -1; }
}

static void level_mark_refinement(olevel *lv)
{
  int64_t  L = lv->L, c, nrow = 0, npl = 0, ir;
  orow    *rows;
  oplane  *pl;
  /* row / plane tables from the sorted cell list */
  for (c = 0; c < lv->ncell; c++)
    if (c == 0 || lv->y[c] != lv->y[c - 1] || lv->z[c] != lv->z[c - 1]) nrow++;
  rows = malloc(sizeof(orow) * (size_t)(nrow + 1));
  pl   = malloc(sizeof(oplane) * (size_t)(nrow + 1));
  nrow = 0;
  for (c = 0; c < lv->ncell; c++) {
    if (c == 0 || lv->y[c] != lv->y[c - 1] || lv->z[c] != lv->z[c - 1]) {
      if (nrow > 0) rows[nrow - 1].c1 = c;
      rows[nrow].y = lv->y[c]; rows[nrow].z = lv->z[c]; rows[nrow].c0 = c; nrow++;
    }
  }
  if (nrow > 0) rows[nrow - 1].c1 = lv->ncell;
  for (ir = 0; ir < nrow; ir++) {
    if (ir == 0 || rows[ir].z != rows[ir - 1].z) {
      if (npl > 0) pl[npl - 1].row1 = ir;
      pl[npl].z = rows[ir].z; pl[npl].row0 = ir; npl++;
    }
  }
  if (npl > 0) pl[npl - 1].row1 = nrow;
  memset(lv->mark, 0, (size_t)lv->ncell);
  if (npl == 0) { free(rows); free(pl); return; }

  {
    int     zwrapped = (pl[0].z == 0 && pl[npl - 1].z == L - 1);   /* refine_grid.c:656 */
    int64_t p0 = 0;
    while (p0 < npl) {                                            /* one z-run (pquad) [p0,p1) */
      int64_t p1 = p0 + 1, zlen, t, last_t;
      int     zlow, zup, is_last_run;
      while (p1 < npl && pl[p1].z == pl[p1 - 1].z + 1) p1++;
      zlen = p1 - p0;
      is_last_run = (p1 == npl);
      run_offsets(zwrapped, pl[p0].z, pl[p1 - 1].z + 1, L, &zlow, &zup);
      /* planes tested by the loop (:699-703) plus, for the last run touching the face, the exit plane (:818) */
      last_t = (zlen - 1) + zup;                                  /* loop runs t in [zlow, last_t) */
      for (t = zlow; t <= ((is_last_run && pl[p1 - 1].z + 1 == L) ? (last_t > zlow ? last_t : zlow) : last_t - 1); t++) {
        const oplane *P, *P0 = &pl[p0];                           /* P0: FIRST plane of the run (quirk, :350) */
        int64_t q0;
        int     ywrapped;
        if (t < 0 || t >= zlen) continue;
        P = &pl[p0 + t];
        ywrapped = (rows[P0->row0].y == 0 && rows[P0->row1 - 1].y == L - 1);
        q0 = P->row0;
        while (q0 < P->row1) {                                    /* one y-run (cquad) [q0,q1) */
          int64_t q1 = q0 + 1, ylen, u, last_u, u_hi;
          int     ylow, yup, y_is_last;
          while (q1 < P->row1 && rows[q1].y == rows[q1 - 1].y + 1) q1++;
          ylen = q1 - q0;
          y_is_last = (q1 == P->row1);
          run_offsets(ywrapped, rows[q0].y, rows[q1 - 1].y + 1, L, &ylow, &yup);
          last_u = (ylen - 1) + yup;
          u_hi = last_u - 1;
          if (y_is_last && rows[q1 - 1].y + 1 == L) u_hi = (last_u > ylow ? last_u : ylow);   /* CASE 2 (:496) */
          for (u = ylow; u <= u_hi; u++) {
            if (u < 0 || u >= ylen) continue;
            refine_row(lv, &rows[q0 + u]);
          }
          q0 = q1;
        }
      }
      p0 = p1;
    }
  }
  free(rows); free(pl);
}

/* ============================================================================================== */
/* build the fine level from the marks                                                             */
/* ============================================================================================== */
typedef struct { int32_t x, y, z; uint8_t brk; } fcell;
static int cmp_fcell(const void *a, const void *b)
{
  const fcell *p = a, *q = b;
  if (p->z != q->z) return p->z < q->z ? -1 : 1;
  if (p->y != q->y) return p->y < q->y ? -1 : 1;
  if (p->x != q->x) return p->x < q->x ? -1 : 1;
  return 0;
}

static olevel *level_refine(olevel *coa)
{
  int64_t c, nm = 0, o = 0;
  fcell  *fc;
  olevel *fin;
  int     i, j, k;
  level_mark_refinement(coa);
  for (c = 0; c < coa->ncell; c++) if (coa->mark[c]) nm++;
  if (nm == 0) return NULL;
  fc = malloc(sizeof(fcell) * (size_t)(8 * nm));
  for (c = 0; c < coa->ncell; c++) {
    if (!coa->mark[c]) continue;
    for (k = 0; k < 2; k++) for (j = 0; j < 2; j++) for (i = 0; i < 2; i++) {
      fc[o].x = 2 * coa->x[c] + i; fc[o].y = 2 * coa->y[c] + j; fc[o].z = 2 * coa->z[c] + k;
      fc[o].brk = (uint8_t)(coa->mark[c] == 2 && i == 1);       /* run ends after a ghost pair */
      o++;
    }
  }
  qsort(fc, (size_t)o, sizeof(fcell), cmp_fcell);
  fin = level_alloc(coa->L * 2, o, 0);
  for (c = 0; c < o; c++) { fin->x[c] = fc[c].x; fin->y[c] = fc[c].y; fin->z[c] = fc[c].z; fin->xbreak[c] = fc[c].brk; }
  free(fc);
  level_build_hash(fin);
  fin->masstopartdens = coa->masstopartdens * CRITMULTI;       /* generate_grids.c:164-170 */
  level_compute_interior(fin);
  return fin;
}

/* ============================================================================================== */
/* R1: relink (relink.c:31-288)                                                                    */
/* ============================================================================================== */
static int64_t level_relink(orc_hier *h, olevel *coa, olevel *fin)
{
  const double Lf = (double)fin->L, shift = 0.5 / Lf;
  int64_t f, moved = 0;
  for (f = 0; f < fin->ncell; f++) {                          /* fine nodes in traversal order */
    int64_t m = cell_at(coa, fin->x[f] / 2, fin->y[f] / 2, fin->z[f] / 2);
    int64_t p, prev;
    double  cx, cy, cz;
    if (m < 0 || coa->head[m] < 0) continue;
    if (!fin->interior[f]) continue;                           /* relink.c:127-131 */
    cx = (double)fin->x[f] / Lf + shift; cy = (double)fin->y[f] / Lf + shift; cz = (double)fin->z[f] / Lf + shift;
    prev = -1; p = coa->head[m];
    while (p >= 0) {
      int64_t nx = h->next[p];
      double  dx = fabs((double)h->pos[3 * p] - cx), dy = fabs((double)h->pos[3 * p + 1] - cy),
              dz = fabs((double)h->pos[3 * p + 2] - cz);
      if (dx <= shift && dy <= shift && dz <= shift) {         /* relink.c:153,:201 inclusive on both faces */
        if (prev < 0) coa->head[m] = nx; else h->next[prev] = nx;
        h->next[p] = fin->head[f]; fin->head[f] = p;           /* push front */
        moved++;
      } else prev = p;
      p = nx;
    }
  }
  return moved;
}

/* undo (relink.c:294-423): ownership goes back to the mother cell; list ORDER is not reproduced */
static void level_relink_back(orc_hier *h, olevel *coa, olevel *fin)
{
  int64_t f;
  for (f = 0; f < fin->ncell; f++) {
    int64_t m = cell_at(coa, fin->x[f] / 2, fin->y[f] / 2, fin->z[f] / 2), p = fin->head[f];
    while (p >= 0) { int64_t nx = h->next[p]; h->next[p] = coa->head[m]; coa->head[m] = p; p = nx; }
    fin->head[f] = -1;
  }
}

/* ============================================================================================== */
/* L1: gen_domgrids / ll / gen_AMRhierarchy                                                        */
/* ============================================================================================== */
orc_hier *orc_hier_build(const float *pos3, int64_t n, int64_t ldom, int64_t lmax, double nth_dom, double nth_ref)
{
  orc_hier *h = calloc(1, sizeof(orc_hier));
  olevel   *dom, *cur;
  int64_t   L = ldom, x, y, z, c, p;
  h->n = n; h->pos = pos3;
  h->next = malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  h->lev  = calloc(64, sizeof(olevel *));
  if (L > (1 << 21) || lmax > (1 << 21)) lmax = (1 << 21);     /* 63-bit cell packing limit */
  /* gen_domgrids (generate_grids.c:24-116): dense periodic L^3 block */
  dom = level_alloc(L, L * L * L, 1);
  c = 0;
  for (z = 0; z < L; z++) for (y = 0; y < L; y++) for (x = 0; x < L; x++) { dom->x[c] = (int32_t)x; dom->y[c] = (int32_t)y; dom->z[c] = (int32_t)z; c++; }
  dom->masstopartdens = ((double)L * (double)L * (double)L) / (double)n;      /* :66-69 pow3(dl1dim)/dno_part */
  dom->critdens       = nth_dom * dom->masstopartdens;                        /* :81-85 (equal masses)      */
  level_compute_interior(dom);
  /* ll (lltools.c:32-92): head insertion in array order */
  for (p = 0; p < n; p++) {
    uint64_t ci[3]; int d;
    for (d = 0; d < 3; d++) {
      ci[d] = (uint64_t)((double)L * (double)pos3[3 * p + d]);
      if (ci[d] > (uint64_t)(L - 1)) ci[d] = 0;
    }
    c = (int64_t)pack3(L, (int64_t)ci[0], (int64_t)ci[1], (int64_t)ci[2]);
    h->next[p] = dom->head[c]; dom->head[c] = p;
  }
  h->lev[0] = dom; h->nlev = 1;
  cur = dom;
  for (;;) {
    olevel *fin;
    int64_t moved;
    level_deposit(h, cur);                                     /* generate_grids.c:314-315 (== :221-222) */
    snapshot_lists(h, cur, &cur->cnt_flag, &cur->plist_flag, &cur->npart_flag);
    if (cur->L == lmax) break;                                 /* :307-308 */
    fin = level_refine(cur);                                   /* refine_grid */
    if (!fin) break;
    fin->critdens = nth_ref * fin->masstopartdens;             /* :166-170 */
    moved = level_relink(h, cur, fin);
    (void)moved;
    if (fin->ncell < MIN_NNODES) {                             /* :231, density.c:420 */
      level_relink_back(h, cur, fin);
      level_free(fin);
      break;
    }
    h->lev[h->nlev++] = fin;
    cur = fin;
  }
  {
    int l;
    for (l = 0; l < h->nlev; l++) snapshot_lists(h, h->lev[l], &h->lev[l]->cnt_final, &h->lev[l]->plist_final, &h->lev[l]->npart_final);
  }
  return h;
}

int orc_hier_nlevels(const orc_hier *h) { return h->nlev; }

void orc_hier_level_header(const orc_hier *h, int lev, int64_t *iout, double *dout)
{
  const olevel *lv = h->lev[lev];
  iout[0] = lv->L; iout[1] = lv->ncell; iout[2] = lv->npart_flag; iout[3] = lv->npart_final;
  dout[0] = lv->critdens; dout[1] = lv->masstopartdens;
}

void orc_hier_level_get(const orc_hier *h, int lev, int32_t *x, int32_t *y, int32_t *z, float *dens,
                        uint8_t *runflags, uint8_t *interior, uint8_t *mark,
                        int32_t *cnt_flag, int64_t *plist_flag, int32_t *cnt_final, int64_t *plist_final)
{
  const olevel *lv = h->lev[lev];
  size_t  nc = (size_t)lv->ncell;
  int64_t c;
  if (x) memcpy(x, lv->x, sizeof(int32_t) * nc);
  if (y) memcpy(y, lv->y, sizeof(int32_t) * nc);
  if (z) memcpy(z, lv->z, sizeof(int32_t) * nc);
  if (dens) memcpy(dens, lv->dens, sizeof(float) * nc);
  if (interior) memcpy(interior, lv->interior, nc);
  if (mark) memcpy(mark, lv->mark, nc);
  if (cnt_flag) memcpy(cnt_flag, lv->cnt_flag, sizeof(int32_t) * nc);
  if (plist_flag) memcpy(plist_flag, lv->plist_flag, sizeof(int64_t) * (size_t)lv->npart_flag);
  if (cnt_final) memcpy(cnt_final, lv->cnt_final, sizeof(int32_t) * nc);
  if (plist_final) memcpy(plist_final, lv->plist_final, sizeof(int64_t) * (size_t)lv->npart_final);
  if (runflags) {
    for (c = 0; c < lv->ncell; c++) {
      int64_t X = lv->x[c], Y = lv->y[c], Z = lv->z[c], r;
      uint8_t f = 0;
      r = cell_at(lv, X - 1, Y, Z); if (r < 0 || lv->xbreak[r]) f |= 1;
      r = cell_at(lv, X + 1, Y, Z); if (r < 0 || lv->xbreak[c]) f |= 2;
      if (lv->dense) {
        if (Y == 0) f |= 4; if (Y == lv->L - 1) f |= 8; if (Z == 0) f |= 16; if (Z == lv->L - 1) f |= 32;
      } else {
        /* y-runs / z-runs are maximal runs of non-empty rows / planes: detect emptiness by scanning the
         * sorted list neighbourhood (rows are contiguous in the list) */
        int64_t a;
        int     prow = 0, nrow = 0, ppl = 0, npl = 0;
        for (a = c; a >= 0 && lv->z[a] == Z && lv->y[a] >= Y - 1; a--) if (lv->y[a] == Y - 1) { prow = 1; break; }
        for (a = c; a < lv->ncell && lv->z[a] == Z && lv->y[a] <= Y + 1; a++) if (lv->y[a] == Y + 1) { nrow = 1; break; }
        for (a = c; a >= 0 && lv->z[a] >= Z - 1; a--) if (lv->z[a] == Z - 1) { ppl = 1; break; }
        for (a = c; a < lv->ncell && lv->z[a] <= Z + 1; a++) if (lv->z[a] == Z + 1) { npl = 1; break; }
        if (!prow) f |= 4; if (!nrow) f |= 8; if (!ppl) f |= 16; if (!npl) f |= 32;
      }
      runflags[c] = f;
    }
  }
}

/* ============================================================================================== */
/* NEXT-1: patch colouring of ahf_gridinfo (ahf_gridinfo.c:236-577)             */
/* ============================================================================================== */
/* colourInfo (ahf_gridinfo.c:1126-1312): the distinct non-zero colours among the six face neighbours the
 * reference's search sees, sorted ascending into holder[0..2] (zeros first), and the action code:
 * 0 new colour, 1 take holder[2], 2 holder[2] -> holder[1], 4 holder[2] and holder[1] -> holder[0].
 * (The reference's codes 2 with r==g / r==b and 3 need equal non-zero entries, which a list of DISTINCT
 * colours cannot hold.)  More than three distinct colours overflow col.holder[NDIM] in the reference:
 * reported as -1, the caller gives up on the level. */
static int colour_info(const olevel *lv, const int32_t *colour, int64_t c, int holder[3])
{
  int64_t nb[3][3][3];
  int     t[6], u[6], nu = 0, i, j;
  neighbours27(lv, c, nb);
  t[0] = nb[1][0][1] >= 0 ? colour[nb[1][0][1]] : 0;      /* y-1  (r) */
  t[1] = nb[1][1][0] >= 0 ? colour[nb[1][1][0]] : 0;      /* x-1  (g) */
  t[2] = nb[0][1][1] >= 0 ? colour[nb[0][1][1]] : 0;      /* z-1  (b) */
  t[3] = nb[1][2][1] >= 0 ? colour[nb[1][2][1]] : 0;      /* y+1  (t) */
  t[4] = nb[1][1][2] >= 0 ? colour[nb[1][1][2]] : 0;      /* x+1  (h) */
  t[5] = nb[2][1][1] >= 0 ? colour[nb[2][1][1]] : 0;      /* z+1  (n) */
  for (i = 0; i < 6; i++) {
    if (t[i] == 0) continue;
    for (j = 0; j < nu; j++) if (u[j] == t[i]) break;
    if (j == nu) u[nu++] = t[i];
  }
  if (nu > 3) return -1;
  holder[0] = holder[1] = holder[2] = 0;
  for (i = 0; i < nu; i++) holder[i] = u[i];
  for (i = 0; i < 3; i++) for (j = i + 1; j < 3; j++) if (holder[j] < holder[i]) { int w = holder[i]; holder[i] = holder[j]; holder[j] = w; }
  return nu == 0 ? 0 : nu == 1 ? 1 : nu == 2 ? 2 : 4;
}

static void colour_replace(int32_t *colour, int64_t c, int code, const int holder[3])
{
  const int cur = colour[c];
  if (code == 2) { if (cur == holder[2]) colour[c] = holder[1]; }
  else if (code == 4) { if (cur == holder[2] || cur == holder[1]) colour[c] = holder[0]; }
}

/* One level, the reference's sequential sweep restated literally: cells in traversal order; a cell that joins two (three)
 * colours takes the smallest, the SPATIALREF records of the others are deleted and the sweep JUMPS BACK to the first node of
 * the (smaller) deleted colour and walks forward replacing until it meets the first uncoloured node (ahf_gridinfo.c:300-330,
 * :395-450, :520-560).  Surviving records in creation order are the level's isolated refinements 0, 1, ... (:751-775).
 * iso[ncell]: isolated-refinement index per cell; periodic3[3*niso]: testBound flags (:1090-1118, :640-672).
 * returns the number of isolated refinements, -1 where the reference itself is undefined (see colour_info). */
int64_t orc_hier_patches(const orc_hier *h, int lev, int32_t *iso, uint8_t *periodic3)
{
  const olevel *lv = h->lev[lev];
  const int64_t nc = lv->ncell;
  int32_t *colour = (int32_t *)calloc((size_t)nc + 1, sizeof(int32_t));
  int64_t *first  = (int64_t *)malloc(sizeof(int64_t) * ((size_t)nc + 2));     /* first node of colour k (record position) */
  uint8_t *alive  = (uint8_t *)calloc((size_t)nc + 2, 1);
  int32_t *rank   = (int32_t *)malloc(sizeof(int32_t) * ((size_t)nc + 2));
  int      counter = 0, replace = 0, code = 0, holder[3] = { 0, 0, 0 };
  int64_t  c, niso = 0;
  for (c = 0; c < nc; c++) {
    if (replace && colour[c] == 0) replace = 0;
    if (replace) { colour_replace(colour, c, code, holder); continue; }
    code = colour_info(lv, colour, c, holder);
    if (code < 0) { niso = -1; goto done; }
    if (code == 0) { colour[c] = ++counter; first[counter] = c; alive[counter] = 1; }
    else if (code == 1) colour[c] = holder[2];
    else {
      int back;
      if (code == 2) { colour[c] = holder[1]; alive[holder[2]] = 0; back = holder[2]; }
      else           { colour[c] = holder[0]; alive[holder[2]] = 0; alive[holder[1]] = 0; back = holder[1]; }
      c = first[back];                          /* the for loop's c++ follows: the first node is treated here */
      colour_replace(colour, c, code, holder);
      replace = 1;
    }
  }
  for (c = 1; c <= counter; c++) rank[c] = alive[c] ? (int32_t)niso++ : -1;
  if (periodic3) memset(periodic3, 0, (size_t)(3 * niso));
  for (c = 0; c < nc; c++) {
    iso[c] = rank[colour[c]];
    if (periodic3 && (lv->x[c] == 0 || lv->y[c] == 0 || lv->z[c] == 0) && iso[c] >= 0) {
      int64_t nb[3][3][3];
      neighbours27(lv, c, nb);
      if (lv->x[c] == 0 && nb[1][1][0] >= 0) periodic3[3 * iso[c] + 0] = 1;
      if (lv->y[c] == 0 && nb[1][0][1] >= 0) periodic3[3 * iso[c] + 1] = 1;
      if (lv->z[c] == 0 && nb[0][1][1] >= 0) periodic3[3 * iso[c] + 2] = 1;
    }
  }
done:
  free(colour); free(first); free(alive); free(rank);
  return niso;
}

/* the six face neighbours of every cell as the reference's search sees them (get_TSCnodes): nb6[6*c + d], d = x-1, x+1, y-1,
 * y+1, z-1, z+1; -1 = not visible.  Lets tests state the colouring as a graph problem (ahf_b200/csrc/patches.cuh). */
void orc_hier_face_neighbours(const orc_hier *h, int lev, int64_t *nb6)
{
  const olevel *lv = h->lev[lev];
  int64_t c, nb[3][3][3];
  for (c = 0; c < lv->ncell; c++) {
    neighbours27(lv, c, nb);
    nb6[6 * c + 0] = nb[1][1][0]; nb6[6 * c + 1] = nb[1][1][2];
    nb6[6 * c + 2] = nb[1][0][1]; nb6[6 * c + 3] = nb[1][2][1];
    nb6[6 * c + 4] = nb[0][1][1]; nb6[6 * c + 5] = nb[2][1][1];
  }
}

/* ============================================================================================== */
/* NEXT-2, first half: RefCentre (ahf_halos.c:935-1390) -- per isolated refinement the node and     */
/* particle counts and the geometric / density-weighted / particle centres                         */
/* ============================================================================================== */
static double f1mod1x(double v) { return v >= 2.0 ? v - 2.0 : v >= 1.0 ? v - 1.0 : v; }     /* specific.c:120-129 */

/* iso / periodic3 / niso as returned by orc_hier_patches for the same level.  out[niso][ORC_NPATCH]:
 *   0 numNodes, 1 numParts, 2-4 centre = centreCMpart (the shipped define.h:101 sets AHFcomcentre: the halo centre is the centre of
 *   mass of the particles linked to the refinement; GEOM where it holds none, ahf_halos.c:1276-1298, :1366-1370), 5 maxDens,
 *   6-8 centreGEOM, 9-11 centreDens (density weighted, with the reference's fall-backs :1248-1274, :1329-1350).
 * Sums run in the reference's order: nodes in traversal order, the particles of a node in list order (ahf_halos.c:1012-1200);
 * node positions are cell centres in double, shifted by one box where the refinement is periodic and the coordinate < 0.5
 * (:1032-1040); tmpDens = dens + mean_dens (= 1), negative values count as zero (:1057-1075). */
void orc_hier_patch_centres(const orc_hier *h, int lev, const int32_t *iso, const uint8_t *periodic3, int64_t niso, double *out)
{
  const olevel *lv = h->lev[lev];
  const double  L = (double)lv->L, shift = 0.5 / (double)lv->L;
  double *acc = (double *)calloc((size_t)niso * 16 + 1, sizeof(double));   /* 0 nodes 1 parts | 2-5 geom+norm | 6-9 dens+norm | 10 maxDens | 11-14 cm+norm */
  int64_t c, i, ip = 0;
  for (i = 0; i < niso; i++) acc[16 * i + 10] = -1.0;                       /* ahf_halos.c:288 */
  for (c = 0; c < lv->ncell; c++) {
    double *a = acc + 16 * (int64_t)iso[c];
    const uint8_t *per = periodic3 + 3 * (int64_t)iso[c];
    double xx = fmod((double)lv->x[c] / L + shift + 1.0, 1.0), yy = fmod((double)lv->y[c] / L + shift + 1.0, 1.0),
           zz = fmod((double)lv->z[c] / L + shift + 1.0, 1.0), d;
    int32_t k;
    if (per[0] && xx < 0.5) xx += 1.0;
    if (per[1] && yy < 0.5) yy += 1.0;
    if (per[2] && zz < 0.5) zz += 1.0;
    a[0] += 1.0;
    a[2] += xx; a[3] += yy; a[4] += zz; a[5] += 1.0;
    d = (double)lv->dens[c] + 1.0;
    if (d < 0.0) d = 0.0;
    a[6] += xx * d; a[7] += yy * d; a[8] += zz * d; a[9] += d;
    if (d > a[10]) a[10] = d;
    for (k = 0; k < lv->cnt_final[c]; k++, ip++) {
      const int64_t p = lv->plist_final[ip];
      double xp = (double)h->pos[3 * p], yp = (double)h->pos[3 * p + 1], zp = (double)h->pos[3 * p + 2];
      if (per[0] && xp < 0.5) xp += 1.0;
      if (per[1] && yp < 0.5) yp += 1.0;
      if (per[2] && zp < 0.5) zp += 1.0;
      a[11] += xp; a[12] += yp; a[13] += zp; a[14] += 1.0; a[1] += 1.0;
    }
  }
  for (i = 0; i < niso; i++) {
    const double *a = acc + 16 * i;
    double *o = out + ORC_NPATCH * i;
    int q;
    o[0] = a[0]; o[1] = a[1]; o[5] = a[10];
    for (q = 0; q < 3; q++) o[6 + q] = a[5] > 0 ? f1mod1x(a[2 + q] / a[5] + 1.0) : a[2 + q];
    for (q = 0; q < 3; q++) o[9 + q] = a[9] > 0 ? f1mod1x(a[6 + q] / a[9] + 1.0) : o[6 + q];
    for (q = 0; q < 3; q++) o[2 + q] = a[14] > 0 ? f1mod1x(a[11 + q] / a[14] + 1.0) : o[6 + q];
    if (a[10] <= 5e-16) for (q = 0; q < 3; q++) o[9 + q] = o[6 + q];       /* MACHINE_ZERO, ahf_halos.c:1329-1350 */
  }
  free(acc);
}

void orc_hier_free(orc_hier *h)
{
  int l;
  if (!h) return;
  for (l = 0; l < h->nlev; l++) level_free(h->lev[l]);
  free(h->lev); free(h->next); free(h);
}

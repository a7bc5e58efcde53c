/*
 * ahf_oracle_halo.c -- TEST INFRASTRUCTURE (see ahf_oracle.h).  Single-threaded CPU restatement of the
 * per-halo pass of the reference (NegriAndrea/AHF), all arithmetic in double as there:
 *   G1  gather            src/libahf/ahf_halos_sfc.c:172-412 (+ hilbert_util.c:124-157 getShell)
 *   U1  radial sort       src/libahf/ahf_halos.c:5786-5894   (NR indexx is unstable; ties here are stable)
 *   U2  virial cut        src/libahf/ahf_halos.c:3687-3889
 *   U3  unbinding         src/libahf/ahf_halos.c:3292-3607
 *   P1  profiles          src/libahf/ahf_halos.c:3961-5018, src/libutility/specific.c:135-178,259-322,
 *                         1977-2072, src/libutility/general.c:548-650,1163-1240
 * Default (flag-less) reference build: equal-mass dark matter.  `weight` / `u` are honoured where the
 * MULTIMASS / GAS_PARTICLES build uses them in U2/U3 and the all-species sums of P1; the per-species
 * blocks of P1 (ahf_halos.c:5024-5256) are not restated (out of scope for round 1).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "ahf_oracle.h"

#define BITS_PER_DIM   21
#define GATHERRAD_FAC  1.001
#define MACHINE_ZERO   5e-16
#define ZERO_F         1e-6          /* param.h:121 (float build) */
#define PI_            3.14159265358979323846
#define GRAV_          4.3006485e-9
#define NIGNORE        5             /* AHF_Rmax_r2_NIGNORE */
#define MINPART_SHELL  10            /* AHF_MINPART_SHELL */

uint64_t orc_hilbert_key_grid(uint32_t x, uint32_t y, uint32_t z, unsigned bits);

/* ---------------------------------------------------------------------------------------------- */
static int64_t lower_bound_key(const uint64_t *keys, int64_t n, uint64_t k)
{
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = lo + (hi - lo) / 2; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
  return lo;
}

typedef struct { int64_t n, cap; int64_t *v; } ivec;
static void ivec_push(ivec *a, int64_t x)
{
  if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 256; a->v = realloc(a->v, sizeof(int64_t) * (size_t)a->cap); }
  a->v[a->n++] = x;
}

/* G1 */
static void gather(const uint64_t *keys, const float *pos, int64_t n, const double ctr[3], double R, ivec *out)
{
  unsigned bits = 1;
  uint64_t cells[27], ckey;
  uint32_t base[3];
  int      ncells, i, j, k, c;
  double   R2 = R * R;
  while ((GATHERRAD_FAC * R < 1. / (double)(1 << (bits + 1))) && ((bits + 1) <= BITS_PER_DIM)) bits++;   /* :244-256 */
  ckey = orc_hilbert_key(ctr[0], ctr[1], ctr[2], bits);
  if (bits == 1) {
    ncells = 8;
    for (i = 0; i < 8; i++) cells[i] = (uint64_t)i;                 /* :210-217 every octant, no distance test */
  } else {
    uint64_t L = (uint64_t)1 << bits;
    orc_hilbert_coords(ckey, bits, base);
    ncells = 0;
    for (i = -1; i <= 1; i++) for (j = -1; j <= 1; j++) for (k = -1; k <= 1; k++) {     /* hilbert_util.c:143-154 */
      uint32_t cx = (uint32_t)((base[0] + L + (uint64_t)(int64_t)i) % L), cy = (uint32_t)((base[1] + L + (uint64_t)(int64_t)j) % L),
               cz = (uint32_t)((base[2] + L + (uint64_t)(int64_t)k) % L);
      double   g = 1. / (double)((uint64_t)1 << bits), big = 0.5 * sqrt(3.) * g, cp[3], d[3], dist2;
      int      q;
      cp[0] = g * cx + 0.5 * g; cp[1] = g * cy + 0.5 * g; cp[2] = g * cz + 0.5 * g;               /* :279-283 */
      for (q = 0; q < 3; q++) { d[q] = fabs(cp[q] - ctr[q]); if (d[q] > 0.5) d[q] = 1.0 - d[q]; }
      dist2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      if (sqrt(dist2) > big + GATHERRAD_FAC * R) continue;                                        /* :294 */
      cells[ncells++] = orc_hilbert_key_grid(cx, cy, cz, bits);
    }
  }
  for (c = 0; c < ncells; c++) {
    unsigned sh = 3 * (BITS_PER_DIM - bits);
    uint64_t kmin = cells[c] << sh, kmax = kmin + (((uint64_t)1 << sh) - 1);
    int64_t  o = lower_bound_key(keys, n, kmin);
    for (; o < n && keys[o] <= kmax; o++) {
      double d[3], dist2;
      int    q;
      for (q = 0; q < 3; q++) { d[q] = fabs((double)pos[3 * o + q] - ctr[q]); if (d[q] > 0.5) d[q] = 1.0 - d[q]; }
      dist2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      if (dist2 <= R2) ivec_push(out, o);
    }
  }
}

/* signed minimum-image separation as in ahf_halos.c:5829-5841 */
static inline void sep3(const float *pos, int64_t p, const double c[3], double d[3])
{
  int q;
  for (q = 0; q < 3; q++) {
    d[q] = (double)pos[3 * p + q] - c[q];
    if (d[q] > 0.5) d[q] -= 1.0;
    if (d[q] < -0.5) d[q] += 1.0;
  }
}
/* |.|-then-wrap variant used by U2 / pass A of U3 / binning_parameter (ahf_halos.c:3838-3846) */
static inline double dist_abs(const float *pos, int64_t p, const double c[3])
{
  double d[3];
  int    q;
  for (q = 0; q < 3; q++) { d[q] = fabs((double)pos[3 * p + q] - c[q]); if (d[q] > 0.5) d[q] -= 1.0; }
  return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

typedef struct { double r2; int64_t ord; int64_t p; } rrec;
static int cmp_rrec(const void *a, const void *b)
{
  const rrec *x = a, *y = b;
  if (x->r2 != y->r2) return x->r2 < y->r2 ? -1 : 1;
  return x->ord < y->ord ? -1 : (x->ord > y->ord);
}

/* U2 */
static void rvir_cut(const float *pos, const float *wgt, const orc_halo_params *par, const double ctr[3],
                     int64_t *ip, int64_t *np, double *M_vir, double *R_vir, double *ovd)
{
  int64_t j = 0, ns = 0;
  double  M = 0.0, R = -1.0, od = 2 * par->ovlim;
  while (j < *np && od >= par->ovlim) {
    double w = wgt ? (double)wgt[ip[j]] : 1.0, V;
    M += w;
    R  = dist_abs(pos, ip[j], ctr);
    V  = 4. * PI_ / 3. * (R * R * R);
    od = M / V * par->rho_fac / par->rho_vir;
    ns++; j++;
  }
  *np = ns; *M_vir = M; *R_vir = R; *ovd = od;
}

/* NR-style index sort is only needed for the median |p|^2 seed of U3: reproduce "idx[n/2] of an ascending
 * index sort" -- ties between equal |p|^2 are resolved by position, which can differ from NR's unstable
 * order only for exactly equal values. */
static int64_t median_seed(const float *mom, const int64_t *ip, int64_t nv)
{
  rrec   *t = malloc(sizeof(rrec) * (size_t)(nv > 0 ? nv : 1));
  int64_t j, r;
  for (j = 0; j < nv; j++) {
    const float *m = mom + 3 * ip[j];
    /* pow2() of a float argument: the reference squares in float precision? pow2 is a macro (x)*(x) on flouble
       operands -> float products, float sums, then widened on assignment to the double array */
    float m2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    t[j].r2 = (double)m2; t[j].ord = j; t[j].p = j;
  }
  qsort(t, (size_t)nv, sizeof(rrec), cmp_rrec);
  r = t[nv / 2 - 1].p;      /* idx[] of NR indexx is 1-based: idx[n/2] is the (n/2)-th smallest */
  free(t);
  return r;
}

/* U3 */
static void unbind(const float *pos, const float *mom, const float *wgt, const float *u, const orc_halo_params *par,
                   const double ctr[3], int64_t *ip, int64_t *np, double *M_vir_o, double *R_vir_o, double *Phi0_o)
{
  double  v2_tune = par->vesc_tune * par->vesc_tune;
  int64_t no_vbulk = (int64_t)(par->min_part / 2), nremove = 4;
  int     niter = 0;
  while (nremove > 3) {
    double  I_now = 0.0, I_prev = 0.0, d_prev = 0.0, M_r = 0.0, Phi0 = 0.0, dist = 0.0, Phi = 0.0;
    double  M_vel, V[3], M_vir = 0.0, R_vir = 0.0;
    int64_t j, seed, nb = 0;
    niter++;
    /* pass A: Phi0 (:3359-3426) */
    for (j = 0; j < *np; j++) {
      double w = wgt ? (double)wgt[ip[j]] : 1.0;
      M_r += w;
      dist = dist_abs(pos, ip[j], ctr);
      if (dist > MACHINE_ZERO) {
        I_now = M_r / (dist * dist);
        Phi0 += ((I_now + I_prev) / 2.) * (dist - d_prev);
      }
      d_prev = dist; I_prev = I_now;
    }
    Phi0 += M_r / dist;
    *Phi0_o = Phi0;
    /* pass B (:3431-3583) */
    nremove = 0; I_now = I_prev = d_prev = M_r = 0.0;
    seed = (niter == 1) ? median_seed(mom, ip, no_vbulk) : 0;
    {
      double w = wgt ? (double)wgt[ip[seed]] : 1.0;
      M_vel = w;
      V[0] = w * mom[3 * ip[seed]]; V[1] = w * mom[3 * ip[seed] + 1]; V[2] = w * mom[3 * ip[seed] + 2];
    }
    for (j = 0; j < *np; j++) {
      int64_t p = ip[j];
      double  w = wgt ? (double)wgt[p] : 1.0, d[3], v_esc2, dv[3], vel2;
      int     q;
      M_r += w;
      sep3(pos, p, ctr, d);
      dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (dist > MACHINE_ZERO) {
        I_now = M_r / (dist * dist);
        Phi  += ((I_now + I_prev) / 2.) * (dist - d_prev);
        v_esc2 = (2 * fabs(Phi - Phi0) * par->phi_fac);
      } else v_esc2 = 1e30;
      for (q = 0; q < 3; q++) dv[q] = ((double)mom[3 * p + q] - V[q] / M_vel) * par->v_fac + par->Hubble * d[q] * par->r_fac;
      vel2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
      if (u) vel2 += (u[p] < 0.0f ? 0.0 : 2 * (double)u[p]);
      if (vel2 > v2_tune * v_esc2) nremove++;
      else {
        M_vel += w;
        /* weight * cur_part->mom[X]: double * float */
        V[0] += w * mom[3 * p]; V[1] += w * mom[3 * p + 1]; V[2] += w * mom[3 * p + 2];
        ip[nb++] = p;
        M_vir += w; R_vir = dist;
      }
      I_prev = I_now; d_prev = dist;
    }
    *np = nb; *M_vir_o = M_vir; *R_vir_o = R_vir;
    if ((double)nb < (double)par->min_part) break;
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* cyclic Jacobi eigen-solver for a symmetric 3x3 matrix (classical algorithm as used by general.c:1163-1240:
 * thresholded sweeps, first three sweeps with threshold 0.2*sum/9, rotation angle via t = sgn/(|theta|+sqrt(1+theta^2))) */
static void jacobi3(double a[3][3], double d[3], double v[3][3])
{
  double b[3], z[3];
  int    ip, iq, i, j;
  for (ip = 0; ip < 3; ip++) { for (iq = 0; iq < 3; iq++) v[ip][iq] = 0.0; v[ip][ip] = 1.0; }
  for (ip = 0; ip < 3; ip++) { b[ip] = d[ip] = a[ip][ip]; z[ip] = 0.0; }
  for (i = 1; i <= 50; i++) {
    double sm = 0.0, tresh;
    for (ip = 0; ip < 2; ip++) for (iq = ip + 1; iq < 3; iq++) sm += fabs(a[ip][iq]);
    if (sm == 0.0) return;
    tresh = (i < 4) ? 0.2 * sm / 9 : 0.0;
    for (ip = 0; ip < 2; ip++) for (iq = ip + 1; iq < 3; iq++) {
      double g = 100.0 * fabs(a[ip][iq]);
      if (i > 4 && (double)(fabs(d[ip]) + g) == (double)fabs(d[ip]) && (double)(fabs(d[iq]) + g) == (double)fabs(d[iq]))
        a[ip][iq] = 0.0;
      else if (fabs(a[ip][iq]) > tresh) {
        double h = d[iq] - d[ip], t, theta, c, s, tau, gg, hh;
        if ((double)(fabs(h) + g) == (double)fabs(h)) t = (a[ip][iq]) / h;
        else {
          theta = 0.5 * h / (a[ip][iq]);
          t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
          if (theta < 0.0) t = -t;
        }
        c = 1.0 / sqrt(1 + t * t); s = t * c; tau = s / (1.0 + c); h = t * a[ip][iq];
        z[ip] -= h; z[iq] += h; d[ip] -= h; d[iq] += h; a[ip][iq] = 0.0;
#define ROT(M, i1, j1, k1, l1) { gg = M[i1][j1]; hh = M[k1][l1]; M[i1][j1] = gg - s * (hh + gg * tau); M[k1][l1] = hh + s * (gg - hh * tau); }
        for (j = 0; j <= ip - 1; j++) ROT(a, j, ip, j, iq)
        for (j = ip + 1; j <= iq - 1; j++) ROT(a, ip, j, j, iq)
        for (j = iq + 1; j < 3; j++) ROT(a, ip, j, iq, j)
        for (j = 0; j < 3; j++) ROT(v, j, ip, j, iq)
#undef ROT
      }
    }
    for (ip = 0; ip < 3; ip++) { b[ip] += z[ip]; d[ip] = b[ip]; z[ip] = 0.0; }
  }
}

/* specific.c:135-178: eigenvalues descending; it[][c] <- eigenvector of the c-th largest */
static void get_axes(double it[3][3], double *ax1, double *ax2, double *ax3)
{
  double a[3][3], d[3], v[3][3];
  int    idx[3] = { 0, 1, 2 }, i, j;
  for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) a[i][j] = it[i][j];
  jacobi3(a, d, v);
  /* ascending index sort of 3 values (insertion sort, as NR indexx does for n < 7) */
  for (j = 1; j < 3; j++) {
    int t = idx[j]; double av = d[t];
    for (i = j - 1; i >= 0; i--) { if (d[idx[i]] <= av) break; idx[i + 1] = idx[i]; }
    idx[i + 1] = t;
  }
  *ax1 = d[idx[2]]; *ax2 = d[idx[1]]; *ax3 = d[idx[0]];
  for (i = 0; i < 3; i++) { it[i][0] = v[i][idx[2]]; it[i][1] = v[i][idx[1]]; it[i][2] = v[i][idx[0]]; }
}

static void smooth3(double *y, int n, int ns)
{
  double *t;
  int     i, j;
  if (ns == 0 || n < 3) return;
  t = calloc((size_t)n, sizeof(double));
  for (j = 0; j < ns; j++) {
    t[0] = (y[0] + y[1]) / 2.;
    for (i = 1; i < n - 1; i++) t[i] = (y[i - 1] + y[i] + y[i + 1]) / 3.;
    t[n - 1] = (y[n - 1] + y[n - 2]) / 2.;
    for (i = 0; i < n; i++) y[i] = t[i];
  }
  free(t);
}

/* general.c:608-650 */
static void find_max(const double *x, const double *yraw, int n, int ns, double *xmax)
{
  double *y = calloc((size_t)(n > 0 ? n : 1), sizeof(double)), ymax = -10.0;
  int     i, m0, m1;
  for (i = 0; i < n; i++) y[i] = yraw[i];
  smooth3(y, n, ns);
  m0 = n - 1;
  for (i = 0; i < n - 1; i++) if (y[i] > ymax) { ymax = y[i]; m0 = i; }
  m1 = m0;
  for (i = (n - 1 + m0) / 2; i > m0; i--) if (y[i] > ymax) { ymax = y[i]; m1 = i; }
  *xmax = (x[m0] + x[m1]) / 2;
  free(y);
}

static double cnfw_root(double c, double r) { return 0.216 * c / (log(1 + c) - c / (1 + c)) - r; }
static double calc_cNFW(double V2_max, double V2_vir)
{
  double r = V2_max / V2_vir, a = 2.2, b = 100, c;
  if (r <= 1 || r > 5.9) return -1;
  while (b - a > 1e-3) { c = (a + b) / 2; if (cnfw_root(a, r) * cnfw_root(c, r) > 0) a = c; else b = c; }
  return (a + b) / 2.0;
}

static double calc_lambdaE(const orc_halo_params *P, double absL, double Mh, double Mass, double Ekin, double Epot)
{
  double t1 = sqrt(P->m_fac * Mh), t2, t3;
  t1 = t1 * t1 * t1;
  t2 = Ekin * P->m_fac * (P->v_fac * P->v_fac);
  t3 = Epot * P->m_fac * P->phi_fac;
  t2 = sqrt(fabs(t2 + t3));
  t1 = t2 / t1;
  t2 = P->m_fac * P->r_fac * P->v_fac * absL;
  t2 = t2 / (P->m_fac * Mass);
  return t1 * t2 / GRAV_;
}

/* P1 */
static void profiles(const float *pos, const float *mom, const float *wgt, const float *u, const orc_halo_params *P,
                     const double ctr[3], const int64_t *ip, int64_t np, double R_vir_in, double Phi0, orc_halo_result *out)
{
  int     nbins = (int)(6.2 * (log10((double)np)) - 3.5), ibin, q;
  double  dist_min = -1.0, dist_max, ldmin, ldmax, ldr;
  int64_t jp, npart = 0, k, mb = -1;
  double  Phi = 0.0, pre_dist = 0.0, I_prev = 0.0, I_now = 0.0, M_prev = 0, V_prev = 0.0, rad_prev;
  double  M_sph_prev = 0.0, prev_dist = 0.0, M = 0.0, Vc[3] = { 0, 0, 0 }, a11 = 0, a22 = 0, a33 = 0, a12 = 0, a13 = 0, a23 = 0;
  double  sig_v = 0, Lv[3] = { 0, 0, 0 }, CoM[3] = { 0, 0, 0 }, Epot = 0, Ekin = 0, Emin = 1e30, M_hires = 0, M_lores = 0;
  double  cur_dist = -1.0, v_esc2 = 0.0, F43 = 4. * PI_ / 3.;
  double  M_gas = 0.0, M_star = 0.0;     /* GAS_PARTICLES build: cumulative gas / star mass (ahf_halos.c:4424-4484) */
  /* per species (0 gas, 1 stars): n, M, com(3), V(3), L(3), a11 a22 a33 a12 a13 a23, Epot, Ekin ; most bound ; u of the shell */
  double  sp[2][19], sp_emin[2] = { 1e30, 1e30 }, u_shell_gas = 0.0;
  int64_t sp_mb[2] = { -1, -1 };
  const double u_fac = (P->x_fac / 0.01) * (P->x_fac / 0.01);          /* (box / t_unit)^2, t_unit = 1/H0 (ahf_halos.c:205, startrun.c:547) */
  double *Vcirc2, *dens_r2, *ovd, *rad, *pr, x_max, r2, R_max, V_max, M_max, absL;
  double *s = out->s;
  if (nbins < 2) nbins = 2;
  /* binning_parameter (specific.c:259-322) */
  k = (int64_t)floor(((double)P->min_part / 10.) + 0.5);
  while (k < np - 1 && dist_min < MACHINE_ZERO) { dist_min = dist_abs(pos, ip[k], ctr); k++; }
  dist_max = dist_abs(pos, ip[np - 1], ctr);
  if (dist_min < MACHINE_ZERO) dist_min = dist_max / 2.;
  ldmin = log10(dist_min); ldmax = log10(dist_max); ldr = (ldmax - ldmin) / (double)nbins;
  out->nbins = nbins;
  out->prof  = calloc((size_t)(ORC_NPROFCOL * nbins), sizeof(double));
  pr = out->prof;
#define PR(col, b) pr[(col) * nbins + (b)]
  Vcirc2 = calloc((size_t)np, sizeof(double)); dens_r2 = calloc((size_t)np, sizeof(double));
  ovd = calloc((size_t)np, sizeof(double)); rad = calloc((size_t)np, sizeof(double));
  memset(sp, 0, sizeof(sp));
  if (u) out->prof_species = calloc((size_t)(3 * nbins), sizeof(double));
  rad_prev = dist_min;
  jp = 0;
  for (ibin = 0; ibin < nbins; ibin++) {
    double cur_rad = pow(10., ldmin + ((double)ibin + 1) * ldr), Volume, dM, dV, it[3][3], ax1, ax2, ax3;
    if (ibin == nbins - 1) cur_rad = dist_max + ZERO_F;
    u_shell_gas = 0.0;                                      /* :4274 */
    while (cur_dist < cur_rad && jp < np) {
      int64_t p = ip[jp];
      double  w = wgt ? (double)wgt[p] : 1.0, d[3], dv[3], Tpart, Upart, Epart, dVv, dMm;
      npart++;
      if (wgt) { if (fabs(w - 1.0) < ZERO_F) M_hires += w; else if (w > 1.0) M_lores += w; } else M_hires += w;
      M += w;
      sep3(pos, p, ctr, d);
      cur_dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (q = 0; q < 3; q++) CoM[q] += w * (ctr[q] + d[q]);
      a11 += w * d[0] * d[0]; a22 += w * d[1] * d[1]; a33 += w * d[2] * d[2];
      a12 += w * d[0] * d[1]; a13 += w * d[0] * d[2]; a23 += w * d[1] * d[2];
      for (q = 0; q < 3; q++) Vc[q] += w * mom[3 * p + q];
      for (q = 0; q < 3; q++) dv[q] = ((double)mom[3 * p + q] - Vc[q] / M);
      Lv[0] += w * (d[1] * dv[2] - d[2] * dv[1]);
      Lv[1] += w * (d[2] * dv[0] - d[0] * dv[2]);
      Lv[2] += w * (d[0] * dv[1] - d[1] * dv[0]);
      for (q = 0; q < 3; q++) dv[q] += P->Hubble * d[q] * P->r_fac / P->v_fac;
      Tpart = w * (dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
      if (cur_dist > MACHINE_ZERO) {
        I_now = M / (cur_dist * cur_dist);
        Phi  += ((I_now + I_prev) / 2.) * (cur_dist - pre_dist);
      }
      Upart  = (Phi - Phi0) * w;
      v_esc2 = 2 * fabs(Upart) / w;
      I_prev = I_now; pre_dist = cur_dist;
      Epot  += Upart;
      if (u && u[p] >= 0.0f) Tpart += w * (2 * (double)u[p] / (P->v_fac * P->v_fac));
      sig_v += Tpart; Ekin += Tpart;
      Epart  = (0.5 * Tpart + Upart);
      if (Epart < Emin) { Emin = Epart; mb = p; }
      if (u) {                                              /* species sums (:4420-4582) */
        const int isg = (u[p] >= 0.0f), iss = (fabs((double)u[p] - (-4.0)) < ZERO_F);
        int t;
        for (t = 0; t < 2; t++) {
          if (!(t == 0 ? isg : iss)) continue;
          double *a = sp[t];
          const double *dvr = dv;                           /* the reference's dVX at this point already carries the Hubble term (:4376-4378); d x d = 0 up to rounding */
          a[0] += 1.0; a[1] += w;
          for (q = 0; q < 3; q++) a[2 + q] += w * (ctr[q] + d[q]);
          for (q = 0; q < 3; q++) a[5 + q] += w * mom[3 * p + q];
          a[8]  += w * (d[1] * dvr[2] - d[2] * dvr[1]);
          a[9]  += w * (d[2] * dvr[0] - d[0] * dvr[2]);
          a[10] += w * (d[0] * dvr[1] - d[1] * dvr[0]);
          a[11] += w * d[0] * d[0]; a[12] += w * d[1] * d[1]; a[13] += w * d[2] * d[2];
          a[14] += w * d[0] * d[1]; a[15] += w * d[0] * d[2]; a[16] += w * d[1] * d[2];
          a[17] += Upart; a[18] += Tpart;
          if (Epart < sp_emin[t]) { sp_emin[t] = Epart; sp_mb[t] = p; }
        }
        if (isg) u_shell_gas += w * (double)u[p] / u_fac;
      }
      /* AHFparticle_Rmax_r2 per-member arrays (:4598-4616) */
      rad[jp] = cur_dist;
      ovd[jp] = M / (F43 * (cur_dist * cur_dist * cur_dist));
      dVv     = F43 * ((cur_dist * cur_dist * cur_dist) - (prev_dist * prev_dist * prev_dist));
      /* AHFdmonly_Rmax_r2 && GAS_PARTICLES (define.h:66, ahf_halos.c:4603-4611): R_max and r2 from the dark matter alone;
         gas: u >= PGAS (0), stars: |u - PSTAR| < ZERO with PSTAR = -4 (param.h:26-28) */
      if (u) { if (u[p] >= 0.0f) M_gas += w; if (fabs((double)u[p] - (-4.0)) < ZERO_F) M_star += w; }
      dMm     = (M - M_gas - M_star) - M_sph_prev;
      Vcirc2[jp] = (M - M_gas - M_star) / cur_dist;
      M_sph_prev = M - M_gas - M_star;
      dens_r2[jp] = dMm / dVv * (((cur_dist + prev_dist) / 2.) * ((cur_dist + prev_dist) / 2.));
      prev_dist = cur_dist;
      jp++;
    }
    Volume = F43 * (cur_rad * cur_rad * cur_rad);
    dM = M - M_prev; dV = Volume - V_prev;
    if (npart > MINPART_SHELL) {
      it[0][0] = a11; it[1][1] = a22; it[2][2] = a33; it[0][1] = it[1][0] = a12; it[0][2] = it[2][0] = a13; it[1][2] = it[2][1] = a23;
      get_axes(it, &ax1, &ax2, &ax3);
    } else { memset(it, 0, sizeof(it)); ax1 = 1; ax2 = 0; ax3 = 0; }
    PR(0, ibin) = (double)npart;  PR(1, ibin) = cur_rad; PR(2, ibin) = M; PR(3, ibin) = M / Volume;
    PR(4, ibin) = (dV > 0) ? dM / dV : 0.0;
    PR(5, ibin) = M / cur_rad; PR(6, ibin) = v_esc2; PR(7, ibin) = sqrt(sig_v / M);
    PR(8, ibin) = 0.5 * Ekin; PR(9, ibin) = 0.5 * Epot; PR(10, ibin) = Lv[0]; PR(11, ibin) = Lv[1]; PR(12, ibin) = Lv[2];
    PR(13, ibin) = 1.0; PR(14, ibin) = it[0][0]; PR(15, ibin) = it[1][0]; PR(16, ibin) = it[2][0];
    PR(17, ibin) = (ax1 > 0.) ? sqrt(ax2 / ax1) : 0.0; PR(18, ibin) = it[0][1]; PR(19, ibin) = it[1][1]; PR(20, ibin) = it[2][1];
    PR(21, ibin) = (ax1 > 0.) ? sqrt(ax3 / ax1) : 0.0; PR(22, ibin) = it[0][2]; PR(23, ibin) = it[1][2]; PR(24, ibin) = it[2][2];
    if (u) { out->prof_species[0 * nbins + ibin] = sp[0][1]; out->prof_species[1 * nbins + ibin] = sp[1][1]; out->prof_species[2 * nbins + ibin] = u_shell_gas; }
    M_prev = M; V_prev = Volume; rad_prev = cur_rad;
  }
  (void)rad_prev;
  for (q = 0; q < 3; q++) CoM[q] = fmod(CoM[q] / M + 1., 1.);
  find_max(rad + NIGNORE, dens_r2 + NIGNORE, (int)np - NIGNORE, 3, &x_max); r2 = x_max;
  find_max(rad + NIGNORE, Vcirc2 + NIGNORE, (int)np - NIGNORE, 1, &x_max); R_max = x_max;
  ibin = 0;
  while (rad[ibin] < x_max && ibin < np - 1) ibin++;
  M_max = ovd[ibin] * F43 * (rad[ibin] * rad[ibin] * rad[ibin]);
  V_max = M_max / R_max;
  free(ovd); free(Vcirc2); free(rad); free(dens_r2);
  absL = sqrt(PR(10, nbins - 1) * PR(10, nbins - 1) + PR(11, nbins - 1) * PR(11, nbins - 1) + PR(12, nbins - 1) * PR(12, nbins - 1));
  s[10] = M;                                             /* M_vir (R_vir keeps the value from U2) */
  s[14] = Vc[0] / M; s[15] = Vc[1] / M; s[16] = Vc[2] / M;
  s[17] = PR(7, nbins - 1); s[18] = PR(6, nbins - 1); s[19] = V_max; s[20] = R_max; s[21] = r2;
  s[24] = PR(8, nbins - 1); s[25] = PR(9, nbins - 1);
  if (absL > 0) {
    s[38] = PR(10, nbins - 1) / absL; s[39] = PR(11, nbins - 1) / absL; s[40] = PR(12, nbins - 1) / absL;
    s[22] = absL / M / sqrt(2. * M * R_vir_in);
    s[22] *= P->v_fac * sqrt(P->r_fac / (GRAV_ * P->m_fac));
    s[23] = calc_lambdaE(P, absL, M, M, s[24], s[25]);
  }
  s[41] = PR(13, nbins - 1); s[42] = PR(17, nbins - 1); s[43] = PR(21, nbins - 1);
  s[44] = PR(14, nbins - 1); s[45] = PR(15, nbins - 1); s[46] = PR(16, nbins - 1);
  s[47] = PR(18, nbins - 1); s[48] = PR(19, nbins - 1); s[49] = PR(20, nbins - 1);
  s[50] = PR(22, nbins - 1); s[51] = PR(23, nbins - 1); s[52] = PR(24, nbins - 1);
  s[54] = calc_cNFW(V_max, M / R_vir_in);
  {
    /* calc_R1 / calc_cR1 (specific.c:1977-2018) */
    double R1 = (PR(1, 0) / 2.0) * (PR(1, 0) / 2.0) * (PR(1, 0) / 2.0) * PR(4, 0) * PR(1, 0), a = 1.0, b = 500.0, c;
    for (ibin = 1; ibin < nbins; ibin++) {
      double rmid = (PR(1, ibin) + PR(1, ibin - 1)) / 2.0, dr = PR(1, ibin) - PR(1, ibin - 1);
      R1 += (rmid * rmid * rmid) * PR(4, ibin) * dr;
    }
    R1 = 4 * PI_ * R1 / M / R_vir_in;
    s[56] = R1;
    if (R1 <= 0.19 || 0.585 <= R1) s[55] = -1.0;
    else {
#define CR1ROOT(cc) (((cc) - 2.0 * log(1.0 + (cc)) + (cc) / (1.0 + (cc))) / ((cc) * (log(1.0 + (cc)) - (cc) / (1.0 + (cc)))) - R1)
      while (b - a > 1e-3) { c = (a + b) / 2; if (CR1ROOT(a) * CR1ROOT(c) > 0) a = c; else b = c; }
#undef CR1ROOT
      s[55] = (a + b) / 2.0;
    }
  }
  {
    double Ts = 2.0 * (PR(8, nbins - 1) - PR(8, nbins - 2)), fr = fabs(PR(1, nbins - 2) / PR(1, nbins - 1));
    s[26] = -0.125 * ((1. + fr) * (1. + fr) * (1. + fr)) / (1. - (fr * fr * fr)) * Ts;
  }
  s[53] = (M_hires > 0) ? M_hires / (M_hires + M_lores) : 0.0;
  if (mb >= 0) {
    double d[3];
    for (q = 0; q < 3; q++) { d[q] = fabs((double)pos[3 * mb + q] - ctr[q]); if (d[q] > 0.5) d[q] -= 1.0; }
    s[37] = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    s[31] = pos[3 * mb]; s[32] = pos[3 * mb + 1]; s[33] = pos[3 * mb + 2];
    s[34] = mom[3 * mb]; s[35] = mom[3 * mb + 1]; s[36] = mom[3 * mb + 2];
  } else s[37] = -1.0;
  {
    double d[3];
    for (q = 0; q < 3; q++) { d[q] = fabs(CoM[q] - ctr[q]); if (d[q] > 0.5) d[q] -= 1.0; }
    s[30] = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    s[27] = CoM[0]; s[28] = CoM[1]; s[29] = CoM[2];
  }
  s[57] = (double)nbins;
  if (u) {                                                  /* gas_only / stars_only (:5020-5181) */
    int t;
    for (t = 0; t < 2; t++) {
      double *a = sp[t], *o = out->species + 32 * t;
      if (a[0] <= 0.0) continue;                            /* reset_SPECIESPROP: all zero */
      o[0] = a[0]; o[1] = a[1];
      for (q = 0; q < 3; q++) { o[2 + q] = fmod(a[2 + q] / a[1] + 1.0, 1.0); o[8 + q] = a[5 + q] / a[1]; }
      o[28] = 0.5 * a[18]; o[29] = 0.5 * a[17];
      if (a[0] > 10.0) {                                    /* AHF_MINPART_GAS / AHF_MINPART_STARS (param.h:14-15) */
        double aL = sqrt(a[8] * a[8] + a[9] * a[9] + a[10] * a[10]), it[3][3], ax1, ax2, ax3;
        o[13] = a[8] / aL; o[14] = a[9] / aL; o[15] = a[10] / aL;
        o[11] = aL / a[1] / sqrt(2. * M * R_vir_in);
        o[11] *= P->v_fac * sqrt(P->r_fac / (GRAV_ * P->m_fac));
        o[12] = calc_lambdaE(P, aL, M, a[1], o[28], o[29]);
        it[0][0] = a[11]; it[1][1] = a[12]; it[2][2] = a[13]; it[0][1] = it[1][0] = a[14]; it[0][2] = it[2][0] = a[15]; it[1][2] = it[2][1] = a[16];
        get_axes(it, &ax1, &ax2, &ax3);
        o[16] = 1.0; o[17] = (ax1 > 0.) ? sqrt(ax2 / ax1) : 0.0; o[18] = (ax1 > 0.) ? sqrt(ax3 / ax1) : 0.0;
        o[19] = it[0][0]; o[20] = it[1][0]; o[21] = it[2][0]; o[22] = it[0][1]; o[23] = it[1][1]; o[24] = it[2][1];
        o[25] = it[0][2]; o[26] = it[1][2]; o[27] = it[2][2];
      }
      if (sp_mb[t] >= 0) { o[5] = pos[3 * sp_mb[t]]; o[6] = pos[3 * sp_mb[t] + 1]; o[7] = pos[3 * sp_mb[t] + 2]; }
    }
  }
#undef PR
}

/* ---------------------------------------------------------------------------------------------- */
void orc_halo_construct(const uint64_t *keys, const float *pos, const float *mom, const float *wgt, const float *u,
                        int64_t n, const orc_halo_params *par, const double ctr[3], double gather_rad, orc_halo_result *out)
{
  ivec    g = { 0, 0, NULL };
  int64_t np, j;
  double  M_vir = 0, R_vir = 0, ovd = 0, Phi0 = 0;
  memset(out, 0, sizeof(*out));
  out->s[0] = ctr[0]; out->s[1] = ctr[1]; out->s[2] = ctr[2]; out->s[3] = gather_rad;
  gather(keys, pos, n, ctr, gather_rad, &g);
  np = g.n;
  out->n_gather = out->n_rvir0 = out->n_unbound = out->n_rvir1 = np;
  if (np >= par->min_part) {                                /* U1 (:5795) */
    rrec *t = malloc(sizeof(rrec) * (size_t)np);
    for (j = 0; j < np; j++) {
      double d[3];
      sep3(pos, g.v[j], ctr, d);
      t[j].r2 = (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); t[j].ord = j; t[j].p = g.v[j];
    }
    qsort(t, (size_t)np, sizeof(rrec), cmp_rrec);
    for (j = 0; j < np; j++) g.v[j] = t[j].p;
    free(t);
  }
  if (np >= par->min_part) { rvir_cut(pos, wgt, par, ctr, g.v, &np, &M_vir, &R_vir, &ovd); Phi0 = 0.0; }
  out->n_rvir0 = out->n_unbound = out->n_rvir1 = np;
  if (np >= par->min_part) unbind(pos, mom, wgt, u, par, ctr, g.v, &np, &M_vir, &R_vir, &Phi0);
  out->n_unbound = out->n_rvir1 = np;
  if (np >= par->min_part) rvir_cut(pos, wgt, par, ctr, g.v, &np, &M_vir, &R_vir, &ovd);
  out->n_rvir1 = np;
  out->s[5] = (double)out->n_gather; out->s[6] = (double)out->n_rvir0; out->s[7] = (double)out->n_unbound;
  out->s[8] = (double)out->n_rvir1; out->s[9] = (double)np;
  out->s[10] = M_vir; out->s[11] = R_vir; out->s[12] = ovd; out->s[13] = Phi0;
  out->npart = np;
  out->ipart = g.v;
  if (np >= par->min_part) profiles(pos, mom, wgt, u, par, ctr, g.v, np, R_vir, Phi0, out);
}

void orc_halo_result_free(orc_halo_result *r)
{
  free(r->ipart); free(r->prof); free(r->prof_species);
  r->ipart = NULL; r->prof = NULL; r->prof_species = NULL;
}

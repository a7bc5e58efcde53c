// catalogue.cu -- NEXT-3 of SURVEY 8f: what ahf_halos() does after the halo loop, as HOST code behind the C-ABI (no device work):
//
//   * the sub-halo re-hash with the final radii (src/libahf/ahf_halos.c:550-640, check_subhalo :5900-5916): a halo the tree listed as
//     substructure stays with its host only if the two virial spheres overlap (|dx| < R_host + 0.5 R_sub, both above NminPerHalo);
//     otherwise it is offered to the host's host and so on, and becomes a field halo when nobody takes it.  The loop is sequential in
//     the reference on purpose (it appends to the lists of OTHER haloes, processed or not): restated with the same order of effects;
//   * the ordering by particle number (:720-741): Numerical Recipes' indexx (libutility/general.c:1093-1150) sorts ascending and the
//     result is reversed, so the order of haloes with EQUAL particle numbers is whatever that quicksort leaves -- restated step for
//     step (median of three, partition, insertion sort below 7, smaller part first), because halo IDs are positions in this order;
//   * the four catalogue files of the default build (libahf/ahf_io.c): <prefix>.AHF_halos (write_halos_line :2209-2345),
//     .AHF_profiles (WriteProfilesLegacy :637-826, with the convergence criterion of Power et al. that flips the sign of the inner
//     radii), .AHF_substructure (:429-503) and .AHF_particles (WriteParticlesLegacy :1094-1175).
//
// The reference writes one fprintf per value and, in the particle file, per ID, and looks halo IDs up with a linear search per
// reference (idx_inv, specific.c:183: O(N_h^2) for the host column alone).  Here: one pass per file into a large buffer, an inverse
// permutation for the IDs, integers formatted by hand; floating-point values go through snprintf with the reference's own
// conversion specifications and the reference's own operand order, so the files are byte-identical
// (tests/test_host_cpu.py::test_catalogue_writer_equals_reference_files on CPU, tests/test_gpu_dropin.py on the device path).
#include "common.cuh"
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

namespace ahf {
namespace {

// ---- ascending index sort, 0-based restatement of Numerical Recipes' indexx as the reference uses it (general.c:1074-1150, M = 7)
void nr_index_sort(const std::vector<double> &key, std::vector<int64_t> &ix)
{
  const int64_t n = (int64_t)key.size();
  ix.resize(n);
  for (int64_t j = 0; j < n; j++) ix[j] = j;
  if (n < 2) return;
  struct Range { int64_t lo, hi; };
  std::vector<Range> todo;                       // the reference's istack (it fails beyond 50 entries; a vector cannot)
  int64_t lo = 0, hi = n - 1;
  auto K = [&](int64_t p) { return key[ix[p]]; };
  for (;;) {
    if (hi - lo < 7) {
      for (int64_t j = lo + 1; j <= hi; j++) {   // straight insertion; NOTE the scan runs down to the array's first element, not to lo
        const int64_t t = ix[j];
        const double  a = key[t];
        int64_t i = j - 1;
        for (; i >= 0; i--) {
          if (K(i) <= a) break;
          ix[i + 1] = ix[i];
        }
        ix[i + 1] = t;
      }
      if (todo.empty()) break;
      lo = todo.back().lo; hi = todo.back().hi; todo.pop_back();
    } else {
      const int64_t mid = (lo + hi + 2) / 2 - 1;  // k = (l + ir) >> 1 in 1-based indices
      std::swap(ix[mid], ix[lo + 1]);
      if (K(lo + 1) > K(hi)) std::swap(ix[lo + 1], ix[hi]);
      if (K(lo) > K(hi)) std::swap(ix[lo], ix[hi]);
      if (K(lo + 1) > K(lo)) std::swap(ix[lo + 1], ix[lo]);
      int64_t i = lo + 1, j = hi;
      const int64_t t = ix[lo];
      const double  a = key[t];
      for (;;) {
        do i++; while (K(i) < a);
        do j--; while (K(j) > a);
        if (j < i) break;
        std::swap(ix[i], ix[j]);
      }
      ix[lo] = ix[j];
      ix[j] = t;
      if (hi - i + 1 >= j - lo) { todo.push_back({ i, hi }); hi = j - 1; }
      else { todo.push_back({ lo, j - 1 }); lo = i; }
    }
  }
}

struct Out {                                       // buffered file
  FILE *f = nullptr;
  std::vector<char> buf;
  size_t n = 0;
  explicit Out(const std::string &path) : buf(8u << 20)
  {
    f = fopen(path.c_str(), "w");
    if (!f) AHF_FAIL("could not open " + path);
  }
  ~Out() { if (f) fclose(f); }
  void flush() { if (n) { if (fwrite(buf.data(), 1, n, f) != n) AHF_FAIL("short write to a catalogue file"); n = 0; } }
  char *room(size_t want) { if (n + want > buf.size()) flush(); return buf.data() + n; }
  template <typename... A> void pf(const char *fmt, A... a)
  {
    char *p = room(512);
    const int k = snprintf(p, 512, fmt, a...);
    if (k < 0 || k >= 512) AHF_FAIL("catalogue field too long");
    n += (size_t)k;
  }
  void str(const char *s) { const size_t k = strlen(s); memcpy(room(k), s, k); n += k; }
  void u64(uint64_t v)                             // "%lu"
  {
    char t[24]; int k = 0;
    do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    char *p = room(24);
    for (int q = 0; q < k; q++) p[q] = t[k - 1 - q];
    n += (size_t)k;
  }
  void ch(char c) { *room(1) = c; n++; }
  void close() { flush(); if (fclose(f) != 0) { f = nullptr; AHF_FAIL("closing a catalogue file failed"); } f = nullptr; }
};

inline double sq(double x) { return x * x; }

}  // namespace

void catalogue_write(const char *fprefix, const ahfgpu_catalogue_in &in, int32_t *host_out, int32_t *nsub_out, int64_t *rank_out)
{
  const int64_t nh = in.nhalo;
  if (nh < 0 || nh >= (1ll << 31)) AHF_FAIL("halo count out of range");
  const int     minpart = in.min_part;
  const bool    gas = (in.flags & 1) != 0;
  if (gas && (!in.species || !in.prof_species || !in.part_u)) AHF_FAIL("the multi-species catalogue needs species, prof_species and part_u");
  auto S = [&](int64_t i, int k) { return in.scal[(size_t)i * AHFGPU_NSCAL + k]; };
  auto npart = [&](int64_t i) { return (long)S(i, 9); };
  // ---- re-hash (ahf_halos.c:550-640)
  std::vector<int32_t> host(nh);
  std::vector<std::vector<int32_t>> sub((size_t)nh);
  for (int64_t i = 0; i < nh; i++) {
    host[i] = in.host ? in.host[i] : -1;
    if (in.sub_off) sub[i].assign(in.sub + in.sub_off[i], in.sub + in.sub_off[i + 1]);
  }
  auto credible = [&](int64_t h, int64_t s) {     // check_subhalo (:5900-5916)
    double dx = std::fabs(in.pos3[3 * h] - in.pos3[3 * s]), dy = std::fabs(in.pos3[3 * h + 1] - in.pos3[3 * s + 1]), dz = std::fabs(in.pos3[3 * h + 2] - in.pos3[3 * s + 2]);
    if (dx > 0.5) dx = 1. - dx;
    if (dy > 0.5) dy = 1. - dy;
    if (dz > 0.5) dz = 1. - dz;
    return sq(dx) + sq(dy) + sq(dz) < sq(S(h, 11) + 0.5 * S(s, 11)) && npart(s) >= minpart && npart(h) >= minpart;
  };
  for (int64_t i = 0; i < nh; i++) {
    std::vector<int32_t> keep;
    for (size_t k = 0; k < sub[i].size(); k++) {
      const int32_t isub = sub[i][k];
      if (isub < 0 || isub >= nh) AHF_FAIL("substructure index out of range");
      int32_t ihost = host[isub];
      if ((in.host_level ? in.host_level[isub] : -1) >= 1) {           // AHF_HOSTHALOLEVEL (param.h:19)
        if (ihost < 0 || ihost >= nh) AHF_FAIL("a listed sub-halo without a host");
        if (credible(ihost, isub)) keep.push_back(isub);
        else {
          ihost = host[ihost];
          while (ihost != -1) {
            if (credible(ihost, isub)) { sub[ihost].push_back(isub); break; }
            ihost = host[ihost];
          }
          host[isub] = ihost;
        }
      } else host[isub] = -1;
    }
    sub[i].swap(keep);
  }
  // ---- order by particle number, descending (:720-741)
  std::vector<double> key((size_t)nh);
  for (int64_t i = 0; i < nh; i++) key[i] = (double)npart(i);
  std::vector<int64_t> asc, idx((size_t)nh), rank((size_t)nh);
  nr_index_sort(key, asc);
  for (int64_t i = 0; i < nh; i++) idx[nh - 1 - i] = asc[i];
  for (int64_t j = 0; j < nh; j++) rank[idx[j]] = j;                  // idx_inv without the search
  if (host_out) for (int64_t i = 0; i < nh; i++) host_out[i] = host[i];
  if (nsub_out) for (int64_t i = 0; i < nh; i++) nsub_out[i] = (int32_t)sub[i].size();
  if (rank_out) for (int64_t i = 0; i < nh; i++) rank_out[i] = rank[i];
  if (!fprefix) return;
  const std::string pre(fprefix);
  const double x_fac = in.x_fac, r_fac = in.r_fac, v_fac = in.v_fac, m_fac = in.m_fac, rho_fac = in.rho_fac, phi_fac = in.phi_fac, u_fac = in.u_fac;
  // ---- .AHF_halos
  {
    Out o(pre + ".AHF_halos");
    static const char *cols[] = { "ID", "hostHalo", "numSubStruct", "Mhalo", "npart", "Xc", "Yc", "Zc", "VXc", "VYc", "VZc", "Rhalo", "Rmax", "r2", "mbp_offset",
                                  "com_offset", "Vmax", "v_esc", "sigV", "lambda", "lambdaE", "Lx", "Ly", "Lz", "b", "c", "Eax", "Eay", "Eaz", "Ebx", "Eby", "Ebz",
                                  "Ecx", "Ecy", "Ecz", "ovdens", "nbins", "fMhires", "Ekin", "Epot", "SurfP", "Phi0", "cNFW" };
    static const char *spc[] = { "n", "M", "lambda", "lambdaE", "Lx", "Ly", "Lz", "b", "c", "Eax", "Eay", "Eaz", "Ebx", "Eby", "Ebz", "Ecx", "Ecy", "Ecz", "Ekin", "Epot" };
    int col = 1;
    o.ch('#');
    for (const char *c : cols) o.pf("%s(%i)\t", c, col++);
    if (gas)
      for (const char *t : { "gas", "star" })
        for (const char *c : spc) o.pf("%s_%s(%i)\t", c, t, col++);
    o.ch('\n');
    for (int64_t j = 0; j < nh; j++) {
      const int64_t i = idx[j];
      if (npart(i) < minpart) continue;
      o.pf("%10lu", (unsigned long)j);
      o.pf("\t%10d", host[i] < 0 ? -1 : (int)rank[host[i]]);
      o.pf("\t%10d", (int)sub[i].size());
      o.pf("\t%12.6g", S(i, 10) * m_fac);
      o.pf("\t%10ld", npart(i));
      o.pf("\t%16.8f", in.pos3[3 * i] * x_fac * 1000.); o.pf("\t%16.8f", in.pos3[3 * i + 1] * x_fac * 1000.); o.pf("\t%16.8f", in.pos3[3 * i + 2] * x_fac * 1000.);
      o.pf("\t%8.2f", S(i, 14) * v_fac); o.pf("\t%8.2f", S(i, 15) * v_fac); o.pf("\t%8.2f", S(i, 16) * v_fac);
      o.pf("\t%10.2f", S(i, 11) * x_fac * 1000.);
      o.pf("\t%10.2f", S(i, 20) * x_fac * 1000.);
      o.pf("\t%10.5f", S(i, 21) * x_fac * 1000.);
      o.pf("\t%10.5f", S(i, 37) * x_fac * 1000.);
      o.pf("\t%10.5f", S(i, 30) * x_fac * 1000.);
      o.pf("\t%8.2f", std::sqrt(S(i, 19) * phi_fac));
      o.pf("\t%12.6f", std::sqrt(S(i, 18) * phi_fac));
      o.pf("\t%8.2f", S(i, 17) * v_fac);
      o.pf("\t%10.6f", S(i, 22));
      o.pf("\t%12.6f", S(i, 23));
      o.pf("\t%12.4g", S(i, 38)); o.pf("\t%12.4g", S(i, 39)); o.pf("\t%12.4g", S(i, 40));
      o.pf("\t%10.6f", S(i, 42)); o.pf("\t%10.6f", S(i, 43));
      for (int k = 44; k <= 52; k++) o.pf("\t%10.6f", S(i, k));
      o.pf("\t%8.2f", S(i, 12));
      o.pf("\t%6d", (int)S(i, 57));
      o.pf("\t%8.6f", S(i, 53));
      o.pf("\t%12.6g", S(i, 24) * m_fac * sq(v_fac));
      o.pf("\t%12.6g", S(i, 25) * m_fac * phi_fac);
      o.pf("\t%12.6g", S(i, 26) * m_fac * sq(v_fac));
      o.pf("\t%12.6g", S(i, 13) * phi_fac);
      o.pf("\t%12.6g", S(i, 54));
      if (gas)
        for (int q = 0; q < 2; q++) {
          const double *t = in.species + 64 * (size_t)i + 32 * q;
          o.pf("\t%10ld", (long)t[0]);
          o.pf("\t%12.6g", t[1] * m_fac);
          o.pf("\t%10.6f", t[11]); o.pf("\t%10.6f", t[12]);
          o.pf("\t%10.6f", t[13]); o.pf("\t%10.6f", t[14]); o.pf("\t%10.6f", t[15]);
          o.pf("\t%10.6f", t[17]); o.pf("\t%10.6f", t[18]);
          for (int k = 19; k <= 27; k++) o.pf("\t%10.6f", t[k]);
          o.pf("\t%12.6g", t[28] * m_fac * sq(v_fac));
          o.pf("\t%12.6g", t[29] * m_fac * phi_fac);
        }
      o.ch('\n');
    }
    o.close();
  }
  // ---- .AHF_profiles
  {
    Out o(pre + ".AHF_profiles");
    static const char *cols[] = { "r", "npart", "M_in_r", "ovdens", "dens", "vcirc", "vesc", "sigv", "Lx", "Ly", "Lz", "b", "c", "Eax", "Eay", "Eaz", "Ebx", "Eby", "Ebz",
                                  "Ecx", "Ecy", "Ecz", "Ekin", "Epot" };
    int col = 1;
    o.ch('#');
    for (const char *c : cols) o.pf("%s(%i)\t", c, col++);
    if (gas) for (const char *c : { "M_gas", "M_star", "U_gas" }) o.pf("%s(%i)\t", c, col++);
    o.ch('\n');
    for (int64_t j = 0; j < nh; j++) {
      const int64_t i = idx[j];
      if (npart(i) < minpart) continue;
      const int     nb = (int)(in.prof_off[i + 1] - in.prof_off[i]);
      const double *pr = in.prof + (size_t)AHFGPU_NPROFCOL * in.prof_off[i];
      const double *ps = gas ? in.prof_species + (size_t)3 * in.prof_off[i] : nullptr;
#define PR(c, b) pr[(size_t)(c) * nb + (b)]
      int r_conv = 0;
      for (int b = 0; b < nb; b++) {             // converged radius (Power et al. 2003, their eq. 20), ahf_io.c:722-752
        const double rad = PR(1, b);
        double nDM;
        if (gas) { const double ngas = ps[0 * nb + b] / in.pmass, nstars = ps[1 * nb + b] / in.pmass; nDM = (double)(unsigned long)PR(0, b) - ngas - nstars; }
        else nDM = (double)(unsigned long)PR(0, b);
        double t_relax = nDM / std::log(nDM) / 8. * rad / std::sqrt(PR(5, b));
        const double t0 = 0.6 * PR(1, nb - 1) / std::sqrt(PR(5, nb - 1)) * r_fac / std::sqrt(phi_fac) * 1E3;
        t_relax *= r_fac / std::sqrt(phi_fac);
        t_relax *= 1E3;
        if (t_relax < t0) r_conv = b + 1;
      }
      for (int b = 0; b < nb; b++) {
        const double rad = (b < r_conv) ? -PR(1, b) : PR(1, b);
        o.pf("%8.4f", rad * x_fac * 1000.);
        o.pf("\t%10ld", (long)(unsigned long)PR(0, b));
        o.pf("\t%e", PR(2, b) * m_fac);
        o.pf("\t%12.2f", PR(3, b) * rho_fac / in.rho_vir);
        o.pf("\t%12.2f", PR(4, b) * rho_fac / in.rho_vir);
        o.pf("\t%8.2f", std::sqrt(PR(5, b) * phi_fac));
        o.pf("\t%10.6f", std::sqrt(PR(6, b) * phi_fac));
        o.pf("\t%8.2f", PR(7, b) * v_fac);
        o.pf("\t%12.4g", PR(10, b) * m_fac * r_fac * v_fac); o.pf("\t%12.4g", PR(11, b) * m_fac * r_fac * v_fac); o.pf("\t%12.4g", PR(12, b) * m_fac * r_fac * v_fac);
        o.pf("\t%10.6f", PR(17, b)); o.pf("\t%10.6f", PR(21, b));
        o.pf("\t%10.6f", PR(14, b)); o.pf("\t%10.6f", PR(15, b)); o.pf("\t%10.6f", PR(16, b));
        o.pf("\t%10.6f", PR(18, b)); o.pf("\t%10.6f", PR(19, b)); o.pf("\t%10.6f", PR(20, b));
        o.pf("\t%10.6f", PR(22, b)); o.pf("\t%10.6f", PR(23, b)); o.pf("\t%10.6f", PR(24, b));
        o.pf("\t%e", PR(8, b) * m_fac * sq(v_fac));
        o.pf("\t%e", PR(9, b) * m_fac * phi_fac);
        if (gas) { o.pf("\t%e", ps[0 * nb + b] * m_fac); o.pf("\t%e", ps[1 * nb + b] * m_fac); o.pf("\t%e", ps[2 * nb + b] * m_fac * u_fac); }
        o.ch('\n');
      }
#undef PR
    }
    o.close();
  }
  // ---- .AHF_substructure
  {
    Out o(pre + ".AHF_substructure");
    for (int64_t j = 0; j < nh; j++) {
      const int64_t i = idx[j];
      if (npart(i) < minpart || sub[i].empty()) continue;
      o.pf("%10d %12d\n", (int)j, (int)sub[i].size());
      for (int32_t s : sub[i]) o.pf("%10d ", (int)rank[s]);
      o.ch('\n');
    }
    o.close();
  }
  // ---- .AHF_particles
  {
    Out o(pre + ".AHF_particles");
    unsigned long ngood = 0;
    for (int64_t j = 0; j < nh; j++) if (npart(idx[j]) >= minpart) ngood++;
    o.pf("%15lu\n", ngood);                        // the reference writes 0 first and patches the count in afterwards: same bytes
    for (int64_t j = 0; j < nh; j++) {
      const int64_t i = idx[j];
      if (npart(i) < minpart) continue;
      o.pf("%lu %10ld\n", (unsigned long)npart(i), (long)j);
      const int64_t *m = in.members + in.member_off[i];
      const int64_t  np = in.member_off[i + 1] - in.member_off[i];
      if (np != npart(i)) AHF_FAIL("member list and particle number of a halo disagree");
      for (int64_t k = 0; k < np; k++) {
        const int64_t p = m[k];
        int type;
        if (in.part_u) type = (in.part_u[p] >= 0.0f) ? 0 : (int)(-in.part_u[p]);               // (u >= PGAS) ? PGAS : -u
        else if (in.part_weight) type = (int)in.part_weight[p];
        else type = 1;                                                                         // -PDM
        o.u64(in.part_id[p]); o.ch('\t');
        if (type < 0) { o.ch('-'); o.u64((uint64_t)(-(long long)type)); } else o.u64((uint64_t)type);
        o.ch('\n');
      }
    }
    o.close();
  }
}

}  // namespace ahf

extern "C" int ahfgpu_catalogue_write(const char *fprefix, const ahfgpu_catalogue_in *in, int32_t *host_out, int32_t *nsub_out, int64_t *rank_out)
{
  try {
    if (!in) AHF_FAIL("null argument");
    if (in->nhalo > 0 && (!in->scal || !in->pos3 || !in->member_off || !in->prof_off)) AHF_FAIL("scal, pos3, member_off and prof_off are required");
    if (fprefix && in->nhalo > 0 && (!in->members || !in->prof || !in->part_id)) AHF_FAIL("members, prof and part_id are required to write the files");
    ahf::catalogue_write(fprefix, *in, host_out, nsub_out, rank_out);
    return 0;
  } catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
    catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
    catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

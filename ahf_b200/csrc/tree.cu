// tree.cu -- HOST code of libahfgpu.so (no kernels): the refinement tree and the halo seeds from the per-refinement tables of
// ahfgpu_amr_patch_stats.  Replaces analyseRef (src/libahf/ahf_halos.c:1652-2300) and spatialRef2halos (:2405-3058) of the
// reference: both are loops over isolated refinements (hundreds to thousands per level), not over nodes or particles, so they stay
// on the host; only their inputs come from the device.  Default switches of the shipped define.h: PARDAU_PARTS (main branch = listed
// refinement with most particles), AHFcomcentre (halo centre = centre of mass of the refinement's particles).
#include "common.cuh"
#include <thread>
#include <algorithm>
#include <cmath>
#include <vector>
#include <array>
#include <functional>

namespace {

// fn(i0, i1) over [0, n) on the host threads (one call when n is small)
inline void run_blocks(int n, int min_parallel, const std::function<void(int, int)> &fn)
{
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr == 0) nthr = 1;
  if (nthr > 64) nthr = 64;
  if (n < min_parallel || nthr == 1) { fn(0, n); return; }
  std::vector<std::thread> pool;
  const int per = (int)((n + nthr - 1) / nthr);
  for (unsigned t = 0; t < nthr; t++) { const int a = (int)t * per, b = std::min(n, a + per); if (a < b) pool.emplace_back(fn, a, b); }
  for (auto &th : pool) th.join();
}

struct Ref {                      // one isolated refinement (SPATIALREF, src/tdef.h)
  double centre[3], cd[3];        // halo centre (stats 2-4), density-weighted centre (stats 9-11; analyseRef works on this one)
  double ext[3][2];               // min, max per dimension (stats 12-17); max < min across a periodic face
  long long nodes, parts;
  std::vector<int> sub, par;      // refinements of the next finer / next coarser level listed with this one
  int    daughter = -1;
  double close = -1.0;            // closeRefDist
  int    halo = -1;               // haloIndex
};

inline double pdist2(const double *a, const double *b)
{
  double s = 0.0, d[3];
  for (int q = 0; q < 3; q++) { d[q] = std::fabs(a[q] - b[q]); if (d[q] > 0.5) d[q] = 1.0 - d[q]; }
  s = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  return s;
}

// is v inside [lo, hi] of a refinement (ahf_halos.c:1710-1745; a periodic extent has hi < lo)
inline bool inside(double v, double lo, double hi)
{
  if (lo < hi) return v > lo && v < hi;
  return (v >= 0 && v < hi) || (v > lo && v <= 1.0);
}

// Uniform periodic cell list over points in [0, 1)^3: the two all-pairs loops of the reference (children x parents of consecutive levels,
// ahf_halos.c:1693-1800; every halo against every halo, :2985-3052) become searches over a few cells.  The answers are unchanged: the same
// candidates pass the same tests, minima are taken over the same double values.
struct CellList {
  int g = 1;
  std::vector<int> start, item;       // CSR: items of cell c are item[start[c] .. start[c + 1])
  int axis_cell(double v) const { int c = (int)std::floor(v * (double)g); return c < 0 ? 0 : (c >= g ? g - 1 : c); }
  size_t cell(int x, int y, int z) const { return ((size_t)z * g + y) * g + x; }
  template <typename PosOf> void build(int n, PosOf pos_of, int per_cell)
  {
    g = 1;
    while ((long long)g * g * g * per_cell < n && g < 256) g++;
    start.assign((size_t)g * g * g + 1, 0);
    std::vector<size_t> cof((size_t)n);
    for (int k = 0; k < n; k++) { const double *p = pos_of(k); cof[k] = cell(axis_cell(p[0]), axis_cell(p[1]), axis_cell(p[2])); start[cof[k] + 1]++; }
    for (size_t c = 0; c + 1 < start.size(); c++) start[c + 1] += start[c];
    item.resize((size_t)n);
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int k = 0; k < n; k++) item[(size_t)fill[cof[k]]++] = k;        // ascending index inside a cell
  }
  // offsets of one axis whose periodic index distance is <= r, each cell once
  void axis_range(int r, int &lo, int &hi) const { lo = -std::min(r, (g - 1) / 2); hi = std::min(r, g / 2); }
  int wrap(int c) const { c %= g; return c < 0 ? c + g : c; }
};

// nearest point (periodic distance, pdist2) among those accepted by `ok`, ties to the smaller index: rings of cells around the query until
// no unvisited cell can hold a closer point.  visit(item) returns false to stop scanning a cell (lists sorted so that the rest is rejected).
template <typename PosOf, typename Accept>
inline void nearest_in_cells(const CellList &G, const double *q, PosOf pos_of, Accept accept, double &best_d2, int &best)
{
  const int cx = G.axis_cell(q[0]), cy = G.axis_cell(q[1]), cz = G.axis_cell(q[2]);
  const double h = 1.0 / (double)G.g;
  const int rmax = G.g / 2;
  for (int r = 0; r <= rmax; r++) {
    if (r > 0 && best >= 0) { const double b = (double)(r - 1) * h * (1.0 - 1e-12); if (best_d2 < b * b) break; }   // cells of ring r are >= (r-1) h away
    int lo, hi; G.axis_range(r, lo, hi);
    auto scan = [&](int dx, int dy, int dz) {
      const size_t c = G.cell(G.wrap(cx + dx), G.wrap(cy + dy), G.wrap(cz + dz));
      for (int t = G.start[c]; t < G.start[c + 1]; t++) {
        const int k = G.item[(size_t)t];
        const int a = accept(k);
        if (a < 0) break;                 // the rest of this cell is rejected as well
        if (a == 0) continue;
        const double d = pdist2(q, pos_of(k));
        if (d < best_d2 || (d == best_d2 && k < best)) { best_d2 = d; best = k; }
      }
    };
    for (int dz = lo; dz <= hi; dz++)
      for (int dy = lo; dy <= hi; dy++) {
        if (std::abs(dz) == r || std::abs(dy) == r) { for (int dx = lo; dx <= hi; dx++) scan(dx, dy, dz); }
        else {                             // interior of the (dy, dz) square: only the two x faces of the ring
          if (-r >= lo) scan(-r, dy, dz);
          if (r > 0 && r <= hi) scan(r, dy, dz);
        }
      }
  }
}

}  // namespace

extern "C" int ahfgpu_tree_halos_ex(int32_t nlev, const int64_t *niso, const double *stats, double max_gather_rad,
                                   int32_t *daughter, double *close_ref_dist, int64_t *sub_offset, int32_t *sub, int64_t sub_cap,
                                   int64_t *nhalo, double *halo_pos3, double *halo_gather_rad, int64_t *halo_npart, int32_t *halo_host,
                                   int64_t halo_cap, int32_t *halo_host_level, int64_t *halo_sub_offset, int32_t *halo_sub, int64_t halo_sub_cap)
{
  try {
    if (nlev < 0 || (nlev && (!niso || !stats)) || !nhalo) AHF_FAIL("null argument");
    const int n = nlev;
    const bool tm = getenv("AHFGPU_TREE_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
      if (!tm) return;
      const auto t = std::chrono::steady_clock::now();
      fprintf(stderr, "[tree] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
      t_last = t;
    };
    std::vector<std::vector<Ref>> R(n);
    {
      const double *s = stats;
      for (int i = 0; i < n; i++) {
        R[i].resize((size_t)niso[i]);
        for (auto &r : R[i]) {
          r.nodes = (long long)s[0]; r.parts = (long long)s[1];
          for (int q = 0; q < 3; q++) { r.centre[q] = s[2 + q]; r.cd[q] = s[9 + q]; r.ext[q][0] = s[12 + 2 * q]; r.ext[q][1] = s[13 + 2 * q]; }
          s += 18;
        }
      }
    }
    // ---- analyseRef (1): a finer refinement is listed with every coarser one whose box holds its density centre (:1693-1800)
    //      The reference tests every pair of two consecutive levels (an O(n_i n_{i+1}) loop: 3.5 s for the 2e4 refinements per level of
    //      a 256^3 box with 2e4 haloes).  Same lists in the same order from a sorted x coordinate: the candidates of a parent are the
    //      children whose centre lies in its x extent (binary searches), tested in y and z and appended in ascending index order; the
    //      parent lists are filled afterwards, parents ascending -- exactly the order the nested loops produce.
    bool detail = false;
    for (int i = 0; i + 1 < n; i++) {
      std::vector<Ref> &Pn = R[i], &Cn = R[i + 1];
      const int nc = (int)Cn.size();
      CellList G;
      G.build(nc, [&](int k) { return Cn[k].cd; }, 2);
      // cells of one axis that can hold a centre inside (lo, hi) -- or inside the two pieces of a periodic extent (hi < lo)
      auto axis_cells = [&](double lo, double hi, std::vector<int> &out) {
        out.clear();
        if (lo < hi) { for (int c = G.axis_cell(lo); c <= G.axis_cell(hi); c++) out.push_back(c); }
        else {
          const int a = G.axis_cell(hi), b = G.axis_cell(lo);
          for (int c = 0; c <= a; c++) out.push_back(c);
          for (int c = std::max(b, a + 1); c < G.g; c++) out.push_back(c);
        }
      };
      auto work = [&](int j0, int j1) {
        std::vector<int> cand, ax[3];
        for (int j = j0; j < j1; j++) {
          const Ref &p = Pn[j];
          cand.clear();
          for (int q = 0; q < 3; q++) axis_cells(p.ext[q][0], p.ext[q][1], ax[q]);
          for (int z : ax[2]) for (int y : ax[1]) for (int x : ax[0]) {
            const size_t cc = G.cell(x, y, z);
            for (int t = G.start[cc]; t < G.start[cc + 1]; t++) {
              const int k = G.item[(size_t)t];
              const Ref &c = Cn[k];
              if (inside(c.cd[0], p.ext[0][0], p.ext[0][1]) && inside(c.cd[1], p.ext[1][0], p.ext[1][1]) && inside(c.cd[2], p.ext[2][0], p.ext[2][1])) cand.push_back(k);
            }
          }
          std::sort(cand.begin(), cand.end());                     // every child sits in one cell: no duplicates; ascending index = the nested loops' order
          Pn[j].sub.insert(Pn[j].sub.end(), cand.begin(), cand.end());
        }
      };
      lap("  (1) build");
      run_blocks((int)Pn.size(), 4096, work);
      lap("  (1) search");
      for (int j = 0; j < (int)Pn.size(); j++)
        for (int k : Pn[j].sub) { Cn[k].par.push_back(j); if (Cn[k].par.size() > 1) detail = true; }
    }
    lap("(1) lists");
    // ---- (2) several parents: keep the closest (first minimum), strike the refinement from the others.  The reference's loop runs over
    //      levels 1 .. n-2 only (:1817): a refinement of the finest level keeps all its parents
    if (detail)
      for (int i = 1; i + 1 < n; i++)
        for (int j = 0; j < (int)R[i].size(); j++) {
          Ref &r = R[i][j];
          if (r.par.size() <= 1) continue;
          int best = -1; double tmin = 10000000000000.0;
          for (int q : r.par) { const double d = pdist2(r.cd, R[i - 1][q].cd); if (d < tmin) { best = q; tmin = d; } }
          if (best < 0) AHF_FAIL("refinement with parents but none closest");
          for (int q : r.par)
            if (q != best) {
              std::vector<int> keep;
              for (int t : R[i - 1][q].sub) if (t != j) keep.push_back(t);
              R[i - 1][q].sub.swap(keep);
            }
          r.par.assign(1, best);
        }
    lap("(2) several parents");
    // ---- (3) no parent: adopt the closest refinement of the level above and inherit its centres (:2030-2165)
    for (int i = 1; i < n; i++) {
      // the centres of the level above as they are when level i is visited (adoption rewrites cd of level i only, after level i - 1 is done)
      CellList up; bool up_built = false;
      std::vector<std::array<double, 3>> cd_up(R[i - 1].size());
      for (size_t k = 0; k < R[i - 1].size(); k++) cd_up[k] = { R[i - 1][k].cd[0], R[i - 1][k].cd[1], R[i - 1][k].cd[2] };
      for (int j = 0; j < (int)R[i].size(); j++) {
        Ref &r = R[i][j];
        if (!r.par.empty()) continue;
        int best = -1; double tmin = 10000000000000.0;
        if (R[i - 1].size() < 512)
          for (int q = 0; q < (int)R[i - 1].size(); q++) { const double d = pdist2(r.cd, R[i - 1][q].cd); if (d < tmin) { best = q; tmin = d; } }
        else {                                                    // first minimum of the same distances, found through the cell list of the level above
          if (!up_built) { up.build((int)R[i - 1].size(), [&](int k) { return cd_up[(size_t)k].data(); }, 2); up_built = true; }
          nearest_in_cells(up, r.cd, [&](int k) { return cd_up[(size_t)k].data(); }, [](int) { return 1; }, tmin, best);
        }
        if (best < 0) continue;                                   // nothing above: the reference would index with -1 here
        r.par.assign(1, best); R[i - 1][best].sub.push_back(j);
        for (int q = 0; q < 3; q++) r.cd[q] = R[i - 1][best].cd[q];
      }
    }
    lap("(3) orphans");
    // ---- (4) main branch (PARDAU_PARTS: most particles, first maximum, :2190-2235) and closeRefDist of the other listed refinements
    //      (half the distance to the nearest sibling, :2245-2285)
    for (int i = 0; i + 1 < n; i++)
      for (auto &r : R[i]) {
        if (r.sub.size() > 1) {
          long long mp = -1; int best = -1;
          for (int k : r.sub) if (R[i + 1][k].parts > mp) { best = k; mp = R[i + 1][k].parts; }
          r.daughter = best;
          for (size_t a = 0; a < r.sub.size(); a++) {
            const int k = r.sub[a];
            if (k == best) continue;
            double tmin = 10000000000000.0;
            for (size_t b = 0; b < r.sub.size(); b++) if (a != b) { const double d = pdist2(R[i + 1][k].cd, R[i + 1][r.sub[b]].cd); if (d < tmin) tmin = d; }
            R[i + 1][k].close = 0.5 * std::sqrt(tmin);
          }
        } else if (r.sub.size() == 1) r.daughter = r.sub[0];
      }
    lap("(4) main branch, closeRefDist");
    // ---- tables out
    {
      int64_t row = 0, ns = 0;
      for (int i = 0; i < n; i++)
        for (auto &r : R[i]) {
          if (daughter) daughter[row] = r.daughter;
          if (close_ref_dist) close_ref_dist[row] = r.close;
          if (sub_offset) sub_offset[row] = ns;
          for (int k : r.sub) { if (sub) { if (ns >= sub_cap) AHF_FAIL("substructure buffer too small"); sub[ns] = k; } ns++; }
          row++;
        }
      if (sub_offset) sub_offset[row] = ns;
    }
    // ---- spatialRef2halos (:2405-2960): walk the tree level by level
    struct Halo { double pos[3] = { 0, 0, 0 }; long long npart = 0; double rvir = -1.0; int host = -1, host_level = -1; std::vector<int> subs; };
    std::vector<Halo> H;
    long long expect = 0;
    for (int i = 0; i < n; i++)
      for (auto &r : R[i]) {
        const long long ns = (long long)r.sub.size();
        expect += (i == 0) ? (ns == 0 ? 1 : ns) : (ns > 1 ? ns - 1 : 0);
      }
    for (int i = 0; i < n; i++)
      for (auto &r : R[i]) {
        const size_t ns = r.sub.size();
        int h;
        if (i == 0) {
          H.emplace_back(); h = (int)H.size() - 1;
          H[h].npart = r.parts;
          if (ns == 0) for (int q = 0; q < 3; q++) H[h].pos[q] = r.centre[q];
        } else {
          if (r.par.empty()) continue;                            // counted as a lost refinement by the reference (:2905-2925)
          h = r.halo;
          if (h < 0) AHF_FAIL("refinement without a halo (the reference exits here, ahf_halos.c:2676)");
          H[h].npart += r.parts;
          if (ns != 1) for (int q = 0; q < 3; q++) H[h].pos[q] = r.centre[q];
        }
        if (ns >= 1 && r.daughter != -1) R[i + 1][r.daughter].halo = h;
        if (ns > 1)
          for (int k : r.sub)
            if (k != r.daughter) {
              H.emplace_back(); const int c = (int)H.size() - 1;
              H[c].host = h; H[c].host_level = i; R[i + 1][k].halo = c; H[c].rvir = R[i + 1][k].close;
              H[h].subs.push_back(c);                              // halos[primHaloIndex].subStruct[kcount] = count (:2705, :2872)
            }
      }
    while ((long long)H.size() < expect) H.emplace_back();
    lap("tables + spatialRef2halos");
    // ---- gathering radius (:2985-3052): half the distance to the nearest halo with MORE particles, at least R_vir, at most
    //      min(MaxGatherRad / boxsize, 1/4)
    const int64_t nh = (int64_t)H.size();
    *nhalo = nh;
    if (nh > halo_cap && (halo_pos3 || halo_gather_rad || halo_npart || halo_host || halo_host_level || halo_sub_offset)) AHF_FAIL("halo buffers too small");
    const double maxg = max_gather_rad < 0.25 ? max_gather_rad : 0.25;
    // the reference's loop is O(N_h^2) (an OpenMP loop, ahf_halos.c:2989-2993: 0.5 s for 2e4 haloes on 8 threads, hours for the 1e6 of a
    // 1024^3 box); here every halo searches the cell list of all haloes, cells sorted by particle number (descending) so that a cell is
    // left at the first halo that is not larger.  The minimum is over the same distances: identical radii.
    std::vector<double> gr((size_t)nh);
    {
      CellList G;
      G.build((int)nh, [&](int k) { return H[(size_t)k].pos; }, 2);
      for (size_t c = 0; c + 1 < G.start.size(); c++)
        std::sort(G.item.begin() + G.start[c], G.item.begin() + G.start[c + 1], [&](int a, int b) { return H[(size_t)a].npart > H[(size_t)b].npart || (H[(size_t)a].npart == H[(size_t)b].npart && a < b); });
      auto work = [&](int i0, int i1) {
        for (int i = i0; i < i1; i++) {
          double g2 = 100000000000.0; int best = -1;
          const long long mine = H[(size_t)i].npart;
          nearest_in_cells(G, H[(size_t)i].pos, [&](int k) { return H[(size_t)k].pos; }, [&](int k) { return H[(size_t)k].npart > mine ? 1 : -1; }, g2, best);
          double g = best >= 0 ? std::sqrt(g2) * 0.5 : maxg;
          if (g < H[(size_t)i].rvir) g = H[(size_t)i].rvir;
          if (g > maxg) g = maxg;
          gr[(size_t)i] = g;
        }
      };
      run_blocks((int)nh, 2048, work);
    }
    lap("gathering radius");
    for (int64_t i = 0; i < nh; i++) {
      if (halo_pos3) for (int q = 0; q < 3; q++) halo_pos3[3 * i + q] = H[i].pos[q];
      if (halo_gather_rad) halo_gather_rad[i] = gr[(size_t)i];
      if (halo_npart) halo_npart[i] = H[i].npart;
      if (halo_host) halo_host[i] = H[i].host;
      if (halo_host_level) halo_host_level[i] = H[i].host_level;
    }
    if (halo_sub_offset) {
      int64_t ns = 0;
      for (int64_t i = 0; i < nh; i++) {
        halo_sub_offset[i] = ns;
        for (int k : H[i].subs) { if (halo_sub) { if (ns >= halo_sub_cap) AHF_FAIL("halo substructure buffer too small"); halo_sub[ns] = k; } ns++; }
      }
      halo_sub_offset[nh] = ns;
    }
    return 0;
  }
  catch (const ahf::Error &e) { ahf::g_last_error = e.msg; return -1; }
  catch (const std::exception &e) { ahf::g_last_error = e.what(); return -2; }
  catch (...) { ahf::g_last_error = "unknown exception"; return -3; }
}

extern "C" int ahfgpu_tree_halos(int32_t nlev, const int64_t *niso, const double *stats, double max_gather_rad,
                                 int32_t *daughter, double *close_ref_dist, int64_t *sub_offset, int32_t *sub, int64_t sub_cap,
                                 int64_t *nhalo, double *halo_pos3, double *halo_gather_rad, int64_t *halo_npart, int32_t *halo_host,
                                 int64_t halo_cap)
{
  return ahfgpu_tree_halos_ex(nlev, niso, stats, max_gather_rad, daughter, close_ref_dist, sub_offset, sub, sub_cap, nhalo, halo_pos3, halo_gather_rad,
                              halo_npart, halo_host, halo_cap, nullptr, nullptr, nullptr, 0);
}

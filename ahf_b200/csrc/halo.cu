// halo.cu -- G/U/P: per-halo gather, radial sort, virial cut, unbinding, virial cut, profiles (sm_100a).
//
// Replaces the OpenMP loop over ahf_halos_sfc_constructHalo (src/libahf/ahf_halos.c:504-510) and its callees
// gatherParts (ahf_halos_sfc.c:172-412), sort_halo_particles (ahf_halos.c:5786), rem_outsideRvir (:3687),
// rem_unbound (:3292), HaloProfiles (:3961).  All halo arithmetic is double precision as in the reference.
// One CTA per halo streams over the radius-sorted members in tiles; every "running" quantity of the
// reference's inside-out loops is a block scan with a carry, the causal running-mean-velocity test of
// rem_unbound is solved per tile as a fixed point (mask -> exclusive scan -> mask).  No tensor cores.
#include "common.cuh"
#include "hilbert.cuh"
#include "scan.cuh"

namespace ahf {

constexpr int    HB = 256;                 // threads per halo CTA
constexpr double MACHINE_ZERO = 5e-16;     // src/param.h:122
constexpr double ZERO_F = 1e-6;            // src/param.h:121
constexpr double PI_ = 3.14159265358979323846264338;
constexpr double GRAV_ = 4.3006485e-9;
constexpr double GATHERRAD_FAC = 1.001;
constexpr int    NIGNORE = 5;
constexpr int    MINPART_SHELL = 10;
constexpr int    MAXBINS = 64;

struct HP {   // parameters passed by value
  double r_fac, x_fac, v_fac, m_fac, rho_fac, phi_fac, hubble, ovlim, rho_vir, vesc_tune;
  int    min_part;
};

// ------------------------------------------------------------------------------------------------
// block-wide helpers (HB threads)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_incl_scan(double v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { double x = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += x; }
  return v;
}
// inclusive scan over the block; *total = block sum.  sm: HB/32 doubles of shared scratch
__device__ __forceinline__ double block_incl_scan(double v, double *sm, double *total)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double inc = warp_incl_scan(v);
  __syncthreads();
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  double base = 0.0, tot = 0.0;
#pragma unroll
  for (int q = 0; q < HB / 32; q++) { double s = sm[q]; if (q < w) base += s; tot += s; }
  *total = tot;
  return base + inc;
}
__device__ __forceinline__ double block_sum(double v, double *sm)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int q = 0; q < HB / 32; q++) t += sm[q];
  return t;
}
__device__ __forceinline__ int block_sum_i(int v, int *sm)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int q = 0; q < HB / 32; q++) t += sm[q];
  return t;
}
__device__ __forceinline__ int block_excl_scan_i(int v, int *sm, int *total)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
  __syncthreads();
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < HB / 32; q++) { int s = sm[q]; if (q < w) base += s; tot += s; }
  *total = tot;
  return base + inc - v;
}
__device__ __forceinline__ long long block_min_ll(long long v, long long *sm)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { long long x = __shfl_xor_sync(0xffffffffu, v, o); v = x < v ? x : v; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  long long t = sm[0];
#pragma unroll
  for (int q = 1; q < HB / 32; q++) t = sm[q] < t ? sm[q] : t;
  return t;
}

// signed minimum-image separation (ahf_halos.c:5829-5841)
__device__ __forceinline__ void sep3(const float4 &p, const double c[3], double d[3])
{
  d[0] = (double)p.x - c[0]; d[1] = (double)p.y - c[1]; d[2] = (double)p.z - c[2];
#pragma unroll
  for (int q = 0; q < 3; q++) { if (d[q] > 0.5) d[q] -= 1.0; if (d[q] < -0.5) d[q] += 1.0; }
}
__device__ __forceinline__ double dist3(const float4 &p, const double c[3])
{
  double d[3];
  sep3(p, c, d);
  return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

// ------------------------------------------------------------------------------------------------
// G1: key ranges of the (at most 27) search cells of every halo (ahf_halos_sfc.c:172-300, :369-412)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t lower_bound_key(const uint64_t *__restrict__ keys, int64_t n, uint64_t k)
{
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < k) lo = mid + 1; else hi = mid; }
  return lo;
}

// one warp per halo, one lane per search cell
__global__ void k_gather_ranges(const uint64_t *__restrict__ keys, int64_t n, const double *__restrict__ centre, const double *__restrict__ grad,
                                const int64_t *__restrict__ seed, int64_t nhalo, int64_t *__restrict__ rlo, int64_t *__restrict__ rhi,
                                int64_t *__restrict__ cand)
{
  const int64_t h = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int     q = threadIdx.x & 31;
  if (h >= nhalo) return;
  int64_t a = 0, b = 0;
  const bool skip = seed && seed[h] == 0;                                  // ahf_halos_sfc.c:122
  if (!skip && q < 27) {
    const double R = grad[h], cx = centre[3 * h], cy = centre[3 * h + 1], cz = centre[3 * h + 2];
    unsigned bits = 1;
    while ((GATHERRAD_FAC * R < 1. / (double)(1 << (bits + 1))) && ((bits + 1) <= 21)) bits++;     // :244-256
    const unsigned sh = 3 * (21 - bits);
    if (bits == 1) {
      if (q < 8) {                                                         // :210-217 all octants
        uint64_t kmin = (uint64_t)q << sh, kmax = kmin + ((1ull << sh) - 1);
        a = lower_bound_key(keys, n, kmin); b = lower_bound_key(keys, n, kmax + 1);
      }
    } else {
      const uint64_t ckey = hilbert_key_posd(cx, cy, cz, bits);
      uint32_t bx, by, bz;
      hilbert_coords(ckey, bits, bx, by, bz);
      const uint32_t L = 1u << bits;
      const double   g = 1. / (double)(1ull << bits), big = 0.5 * sqrt(3.) * g;
      const int i = q / 9 - 1, j = (q / 3) % 3 - 1, k = q % 3 - 1;          // hilbert_util.c:143-154: q = (i+1)*9 + (j+1)*3 + (k+1)
      uint32_t x = (bx + L + i) % L, y = (by + L + j) % L, z = (bz + L + k) % L;
      double dx = fabs((g * x + 0.5 * g) - cx), dy = fabs((g * y + 0.5 * g) - cy), dz = fabs((g * z + 0.5 * g) - cz);
      if (dx > 0.5) dx = 1.0 - dx; if (dy > 0.5) dy = 1.0 - dy; if (dz > 0.5) dz = 1.0 - dz;
      if (!(sqrt(dx * dx + dy * dy + dz * dz) > big + GATHERRAD_FAC * R)) {                 // :294
        uint64_t kc = hilbert_index(x, y, z, bits), kmin = kc << sh, kmax = kmin + ((1ull << sh) - 1);
        a = lower_bound_key(keys, n, kmin); b = lower_bound_key(keys, n, kmax + 1);
      }
    }
  }
  if (q < 27) { rlo[h * 27 + q] = a; rhi[h * 27 + q] = b; }
  long long tot = b - a;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (q == 0) cand[h] = tot;
}

// G2: one CTA per halo; members = candidates with periodic d^2 <= R^2, appended range by range (ahf_halos_sfc.c:327-356)
__global__ void __launch_bounds__(HB) k_gather_fill(const float4 *__restrict__ pos4, const double *__restrict__ centre, const double *__restrict__ grad,
                                                   const int64_t *__restrict__ rlo, const int64_t *__restrict__ rhi, const int64_t *__restrict__ candoff,
                                                   double *__restrict__ r2buf, uint32_t *__restrict__ idxbuf, int64_t *__restrict__ ngather)
{
  __shared__ int sm[HB / 32];
  const int64_t h = blockIdx.x;
  const double  c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const double  R2 = grad[h] * grad[h];
  int64_t out = candoff[h];
  for (int q = 0; q < 27; q++) {
    const int64_t a = rlo[h * 27 + q], b = rhi[h * 27 + q];
    for (int64_t base = a; base < b; base += HB) {
      int64_t o = base + threadIdx.x;
      bool    in = false;
      double  r2 = 0.0;
      if (o < b) {
        float4 p = pos4[o];
        double dx = fabs((double)p.x - c[0]), dy = fabs((double)p.y - c[1]), dz = fabs((double)p.z - c[2]);
        if (dx > 0.5) dx = 1.0 - dx; if (dy > 0.5) dy = 1.0 - dy; if (dz > 0.5) dz = 1.0 - dz;
        in = (dx * dx + dy * dy + dz * dz) <= R2;
        if (in) { double d[3]; sep3(p, c, d); r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]; }   // sort key (ahf_halos.c:5844)
      }
      int tot, pos = block_excl_scan_i(in ? 1 : 0, sm, &tot);
      if (in) { r2buf[out + pos] = r2; idxbuf[out + pos] = (uint32_t)o; }
      out += tot;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) ngather[h] = out - candoff[h];
}

// ------------------------------------------------------------------------------------------------
// U1: radial sort = LSD radix sort by r^2 bits, then stable sort by halo -> (halo, r^2, gather order)
// ------------------------------------------------------------------------------------------------
// Radial sort of ALL haloes in one LSD sort: key = halo index (hb bits) | 6 exponent bits | mb mantissa bits of r^2, a monotone but
// not injective image of (halo, r^2).  Exponents below 2^-64 collapse to zero and the mantissa is cut to mb bits; members whose keys
// agree are still in gather order and k_fix_ties orders every such run by the whole r^2 (then by gather position): the result is the
// full stable sort by (halo, r^2).  hb + 6 + mb is a multiple of 8: five passes for the 256^3 box (two sorts of 5 + 2 passes before).
__device__ __forceinline__ uint64_t halo_sort_key(uint64_t h, double r2, int mb)
{
  const uint64_t b = (uint64_t)__double_as_longlong(r2);               // r^2 >= 0: the bit pattern orders like the value
  const int      e = (int)(b >> 52) - 959;
  const uint64_t m = (b & ((1ull << 52) - 1ull)) >> (52 - mb);
  const uint64_t q = e < 0 ? 0ull : e > 63 ? ((64ull << mb) - 1ull) : (((uint64_t)e << mb) | m);
  return (h << (6 + mb)) | q;
}
// one thread per element of the eoff-packed layout (haloes below min_part have empty ranges there: ahf_halos.c:5795, left unsorted);
// its halo by binary search in eoff -- a CTA per halo was bound by the largest halo (0.26 ms at 256^3)
__global__ void k_sort_setup(const int64_t *__restrict__ candoff, const int64_t *__restrict__ eoff, int64_t nhalo, uint64_t ne,
                             const double *__restrict__ r2buf, uint64_t *__restrict__ key, uint32_t *__restrict__ val,
                             uint32_t *__restrict__ hid, int mb)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= ne) return;
  int64_t lo = 0, hi = nhalo;                                  // first h with eoff[h] > i, minus one
  while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (eoff[mid] <= (int64_t)i) lo = mid + 1; else hi = mid; }
  const int64_t h = lo - 1;
  key[i] = halo_sort_key((uint64_t)h, r2buf[candoff[h] + ((int64_t)i - eoff[h])], mb);
  val[i] = (uint32_t)i;
  hid[i] = (uint32_t)h;
}
// element e of the packed (>= min_part) layout lives at candoff[h] + (e - eoff[h]) of the candidate buffers
// One thread per run of equal sort keys (a few per 10^5 members: the mantissa bits kept resolve well below the spacing of neighbouring
// r^2) puts the run into the order the full stable sort gives: by the whole r^2, then by gather position.
__global__ void k_fix_ties(uint32_t *__restrict__ perm, const uint64_t *__restrict__ skey, const uint32_t *__restrict__ hid, const int64_t *__restrict__ candoff,
                           const int64_t *__restrict__ eoff, const double *__restrict__ r2buf, uint64_t ne)
{
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i + 1 >= ne) return;
  const uint64_t top = skey[i];
  if (skey[i + 1] != top) return;                                     // alone, or the last of its run
  if (i > 0 && skey[i - 1] == top) return;                            // not the head of its run
  auto key = [&](uint32_t e) { const uint32_t h = hid[e]; return (uint64_t)__double_as_longlong(r2buf[candoff[h] + ((int64_t)e - eoff[h])]); };
  uint64_t j = i + 2;
  while (j < ne && skey[j] == top) j++;
  if (j - i < 2) return;
  auto less = [&](uint32_t ea, uint32_t eb) { const uint64_t ka = key(ea), kb = key(eb); return ka < kb || (ka == kb && ea < eb); };
  if (j - i > 32) {                                            // long run (thin shells, huge haloes): heap sort, O(L log L)
    uint32_t *a = perm + i;
    const uint64_t L = j - i;
    auto sift = [&](uint64_t root, uint64_t end) {
      for (;;) {
        uint64_t ch = 2 * root + 1;
        if (ch >= end) return;
        if (ch + 1 < end && less(a[ch], a[ch + 1])) ch++;
        if (!less(a[root], a[ch])) return;
        const uint32_t t = a[root]; a[root] = a[ch]; a[ch] = t;
        root = ch;
      }
    };
    for (uint64_t st = L / 2; st-- > 0;) sift(st, L);
    for (uint64_t end = L - 1; end > 0; end--) { const uint32_t t = a[0]; a[0] = a[end]; a[end] = t; sift(0, end); }
    return;
  }
  for (uint64_t a = i + 1; a < j; a++) {                       // insertion sort by (key, gather position)
    const uint32_t ea = perm[a];
    const uint64_t ka = key(ea);
    uint64_t b = a;
    while (b > i) {
      const uint32_t eb = perm[b - 1];
      const uint64_t kb = key(eb);
      if (kb < ka || (kb == ka && eb < ea)) break;
      perm[b] = eb; b--;
    }
    perm[b] = ea;
  }
}
// sorted position i of the eoff-packed layout -> the halo's slot of the moff0-packed member list (the sort is by halo first, so
// position i belongs to the halo of the element that landed there)
__global__ void k_apply_perm(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ hid, const int64_t *__restrict__ candoff,
                             const int64_t *__restrict__ eoff, const int64_t *__restrict__ moff0, uint64_t ne, const uint32_t *__restrict__ idxbuf,
                             uint32_t *__restrict__ out)
{
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= ne) return;
  uint32_t e = perm[i], h = hid[e];
  out[moff0[h] + ((int64_t)i - eoff[h])] = idxbuf[candoff[h] + ((int64_t)e - eoff[h])];
}
__global__ void k_copy_unsorted(const int64_t *__restrict__ candoff, const int64_t *__restrict__ ngather, const int64_t *__restrict__ eoff2,
                                int min_part, const uint32_t *__restrict__ idxbuf, uint32_t *__restrict__ out)
{
  const int64_t h = blockIdx.x, ng = ngather[h];
  if (ng >= min_part) return;
  for (int64_t i = threadIdx.x; i < ng; i += blockDim.x) out[eoff2[h] + i] = idxbuf[candoff[h] + i];
}

struct RvirOut { long long np; double M, R, ovd; };

constexpr int HI = 4;            // members per thread and tile (blocked: thread t owns HI consecutive members)
constexpr int HT = HB * HI;      // members per tile

// exclusive block scan of NC doubles per thread at once; tot[] = block totals.  sm: (HB/32)*NC doubles
template <int NC> __device__ __forceinline__ void block_excl_scan_n(const double (&v)[NC], double (&ex)[NC], double (&tot)[NC], double *sm)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double inc[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    inc[c] = v[c];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double x = __shfl_up_sync(0xffffffffu, inc[c], o); if (lane >= o) inc[c] += x; }
  }
  __syncthreads();
  if (lane == 31) {
#pragma unroll
    for (int c = 0; c < NC; c++) sm[w * NC + c] = inc[c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NC; c++) {
    double base = 0.0, t = 0.0;
#pragma unroll
    for (int q = 0; q < HB / 32; q++) { double s = sm[q * NC + c]; if (q < w) base += s; t += s; }
    ex[c] = base + inc[c] - v[c]; tot[c] = t;
  }
}
template <int NC> __device__ __forceinline__ void block_sum_n(double (&v)[NC], double *sm)
{
#pragma unroll
  for (int c = 0; c < NC; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < NC; c++) sm[(threadIdx.x >> 5) * NC + c] = v[c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NC; c++) { double t = 0.0;
#pragma unroll
    for (int q = 0; q < HB / 32; q++) t += sm[q * NC + c];
    v[c] = t; }
}

// one tile of radius-sorted members in registers
struct TileMembers {
  uint32_t pid[HI];
  double   w[HI], r[HI], d[HI][3];
  bool     act[HI];
};
__device__ __forceinline__ void load_tile(TileMembers &T, const float4 *__restrict__ pos4, const uint32_t *__restrict__ ip, long long base, long long np,
                                          const double c[3])
{
#pragma unroll
  for (int i = 0; i < HI; i++) {
    long long j = base + (long long)threadIdx.x * HI + i;
    T.act[i] = j < np; T.pid[i] = 0; T.w[i] = 0.0; T.r[i] = 0.0; T.d[i][0] = T.d[i][1] = T.d[i][2] = 0.0;
    if (T.act[i]) {
      T.pid[i] = ip[j];
      float4 p = pos4[T.pid[i]];
      T.w[i] = (double)p.w; sep3(p, c, T.d[i]);
      T.r[i] = sqrt(T.d[i][0] * T.d[i][0] + T.d[i][1] * T.d[i][1] + T.d[i][2] * T.d[i][2]);
    }
  }
}
// M(<=j) for the tile's members: thread-local running sum + one block scan.  Returns the tile total.
__device__ __forceinline__ double tile_mass(const TileMembers &T, double carryM, double (&M)[HI], double *smd)
{
  double loc[1] = { 0.0 }, ex[1], tot[1];
#pragma unroll
  for (int i = 0; i < HI; i++) loc[0] += T.w[i];
  block_excl_scan_n<1>(loc, ex, tot, smd);
  double run = carryM + ex[0];
#pragma unroll
  for (int i = 0; i < HI; i++) { run += T.w[i]; M[i] = run; }
  return tot[0];
}
// Phi(r_j) = trapezoid sum of M(<r)/r^2 (ahf_halos.c:3398-3410, :3514-3526): needs (r, I) of member j-1
__device__ __forceinline__ double tile_phi(const TileMembers &T, const double (&M)[HI], double carryPhi, double &prev_r, double &prev_I,
                                           double (&Phi)[HI], double *nb_r, double *nb_I, double *smd, long long tile_n)
{
  double I[HI];
#pragma unroll
  for (int i = 0; i < HI; i++) I[i] = (T.act[i] && T.r[i] > MACHINE_ZERO) ? M[i] / (T.r[i] * T.r[i]) : 0.0;
  __syncthreads();
  nb_r[threadIdx.x] = T.r[HI - 1]; nb_I[threadIdx.x] = I[HI - 1];
  __syncthreads();
  double rp = threadIdx.x ? nb_r[threadIdx.x - 1] : prev_r, Ip = threadIdx.x ? nb_I[threadIdx.x - 1] : prev_I;
  double term[HI], loc[1] = { 0.0 }, ex[1], tot[1];
#pragma unroll
  for (int i = 0; i < HI; i++) {
    term[i] = (T.act[i] && T.r[i] > MACHINE_ZERO) ? ((I[i] + Ip) / 2.) * (T.r[i] - rp) : 0.0;
    loc[0] += term[i];
    rp = T.r[i]; Ip = I[i];
  }
  block_excl_scan_n<1>(loc, ex, tot, smd);
  double run = carryPhi + ex[0];
#pragma unroll
  for (int i = 0; i < HI; i++) { run += term[i]; Phi[i] = run; }
  // (r, I) of the tile's last member -> carry for the next tile
  const long long lt = (tile_n - 1) / HI; const int li = (int)((tile_n - 1) % HI);
  __syncthreads();
  if (threadIdx.x == lt) { nb_r[HB] = T.r[li]; nb_I[HB] = I[li]; }
  __syncthreads();
  prev_r = nb_r[HB]; prev_I = nb_I[HB];
  return tot[0];
}

// ------------------------------------------------------------------------------------------------
// U2: rem_outsideRvir (ahf_halos.c:3811-3869): first member whose mean enclosed overdensity drops below ovlim, inclusive
// ------------------------------------------------------------------------------------------------
__device__ RvirOut rvir_cut(const float4 *__restrict__ pos4, const uint32_t *__restrict__ ip, long long np, const double c[3], const HP &P,
                            double *smd, long long *sml)
{
  __shared__ RvirOut res;
  double carryM = 0.0;
  if (threadIdx.x == 0) { res.np = np; res.M = 0; res.R = -1.0; res.ovd = 2 * P.ovlim; }
  __syncthreads();
  for (long long base = 0; base < np; base += HT) {
    TileMembers T;
    load_tile(T, pos4, ip, base, np, c);
    double M[HI];
    double totM = tile_mass(T, carryM, M, smd);
    long long mine = 0x7fffffffffffffffll;
    double od[HI];
#pragma unroll
    for (int i = 0; i < HI; i++) {
      od[i] = 0.0;
      if (T.act[i]) {
        double V = 4. * PI_ / 3. * (T.r[i] * T.r[i] * T.r[i]);
        od[i] = M[i] / V * P.rho_fac / P.rho_vir;
        if (!(od[i] >= P.ovlim) && mine == 0x7fffffffffffffffll) mine = base + (long long)threadIdx.x * HI + i;
      }
    }
    long long first = block_min_ll(mine, sml);
    if (first != 0x7fffffffffffffffll) {
#pragma unroll
      for (int i = 0; i < HI; i++)
        if (base + (long long)threadIdx.x * HI + i == first) { res.np = first + 1; res.M = M[i]; res.R = T.r[i]; res.ovd = od[i]; }
      __syncthreads();
      return res;
    }
#pragma unroll
    for (int i = 0; i < HI; i++)
      if (base + (long long)threadIdx.x * HI + i == np - 1) { res.np = np; res.M = M[i]; res.R = T.r[i]; res.ovd = od[i]; }
    carryM += totM;
    __syncthreads();
  }
  __syncthreads();
  return res;
}

// ------------------------------------------------------------------------------------------------
// stage kernel 1: virial cut, unbinding, virial cut.  One CTA per halo; the member list is compacted in place.
// ------------------------------------------------------------------------------------------------
// MB: CTAs per SM the register budget is cut for -- 1 (255 registers) when a few hundred haloes wait for the longest of them, 2 (128
// registers, some spills) when thousands of small haloes queue for the SMs (2e4 haloes: 2.5 -> 1.8 ms)
template <int MB>
__global__ void __launch_bounds__(HB, MB) k_halo_unbind(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, int has_u,
                                                    const double *__restrict__ centre, const int64_t *__restrict__ moff0, const int64_t *__restrict__ ngather,
                                                    uint32_t *__restrict__ members, HP P, double *__restrict__ scal, int64_t *__restrict__ npart_out,
                                                    int64_t *__restrict__ iter_work, const int32_t *__restrict__ sel)
{
  __shared__ double    smd[(HB / 32) * 4];
  __shared__ long long sml[HB / 32];
  __shared__ int       smi[HB / 32];
  __shared__ double    nb_r[HB + 1], nb_I[HB + 1];
  __shared__ double    s_seed[4];
  __shared__ double    s_R;
  const int64_t h = sel ? (int64_t)sel[blockIdx.x] : (int64_t)blockIdx.x;      // sel: the haloes this launch serves (the small ones of the hybrid pass)
  uint32_t     *ip = members + moff0[h];
  long long     np = ngather[h];
  double       *S = scal + h * AHFGPU_NSCAL;
  const double  c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  if (threadIdx.x == 0) { S[5] = (double)np; S[6] = S[7] = S[8] = S[9] = (double)np; }
  double M_vir = 0, R_vir = 0, ovd = 0, Phi0 = 0;
  long long work = 0;
  if (np >= P.min_part) {
    RvirOut r = rvir_cut(pos4, ip, np, c, P, smd, sml);
    np = r.np; M_vir = r.M; R_vir = r.R; ovd = r.ovd; Phi0 = 0.0;
  }
  if (threadIdx.x == 0) S[6] = S[7] = S[8] = S[9] = (double)np;
  // ---- rem_unbound (ahf_halos.c:3292-3607)
  if (np >= P.min_part) {
    const double v2_tune = P.vesc_tune * P.vesc_tune;
    long long nremove = 4;
    int niter = 0;
    while (nremove > 3) {
      niter++;
      work += np;
      // pass A: Phi0 = sum of trapezoids of M(<r)/r^2 + M_tot/r_last (:3359-3426)
      double carryM = 0.0, carryPhi = 0.0, prev_r = 0.0, prev_I = 0.0;
      for (long long base = 0; base < np; base += HT) {
        TileMembers T;
        load_tile(T, pos4, ip, base, np, c);
        double M[HI], Phi[HI];
        const long long tile_n = (base + HT < np ? base + HT : np) - base;
        carryM += tile_mass(T, carryM, M, smd);
        carryPhi += tile_phi(T, M, carryPhi, prev_r, prev_I, Phi, nb_r, nb_I, smd, tile_n);
      }
      Phi0 = carryPhi + carryM / prev_r;
      // seed of the running bulk velocity (:3441-3470)
      if (threadIdx.x == 0) {
        long long seed = 0;
        if (niter == 1) {
          int nv = P.min_part / 2;
          float m2[64]; int id[64];
          if (nv > 64) nv = 64;
          for (int q = 0; q < nv; q++) { float4 m = mom4[ip[q]]; m2[q] = m.x * m.x + m.y * m.y + m.z * m.z; id[q] = q; }
          for (int a = 1; a < nv; a++) { int t = id[a]; float v = m2[t]; int b = a - 1; while (b >= 0 && m2[id[b]] > v) { id[b + 1] = id[b]; b--; } id[b + 1] = t; }
          seed = id[nv / 2 - 1];                                  // NR indexx is 1-based: idx[n/2] = (n/2)-th smallest
        }
        float4 p = pos4[ip[seed]], m = mom4[ip[seed]];
        double w = (double)p.w;
        s_seed[0] = w; s_seed[1] = w * m.x; s_seed[2] = w * m.y; s_seed[3] = w * m.z;
      }
      __syncthreads();
      double run0[4] = { s_seed[0], s_seed[1], s_seed[2], s_seed[3] };     // M_vel, V (bound so far, before this tile)
      // pass B (:3478-3583)
      carryM = 0.0; carryPhi = 0.0; prev_r = 0.0; prev_I = 0.0;
      long long nb = 0;
      double Mv_acc = 0.0, Rv_last = 0.0;
      nremove = 0;
      for (long long base = 0; base < np; base += HT) {
        TileMembers T;
        load_tile(T, pos4, ip, base, np, c);
        const long long tile_n = (base + HT < np ? base + HT : np) - base;
        double M[HI], Phi[HI], mom[HI][3], uu[HI], vesc2[HI];
#pragma unroll
        for (int i = 0; i < HI; i++) {
          mom[i][0] = mom[i][1] = mom[i][2] = 0.0; uu[i] = -1.0;
          if (T.act[i]) { float4 m = mom4[T.pid[i]]; mom[i][0] = (double)m.x; mom[i][1] = (double)m.y; mom[i][2] = (double)m.z; uu[i] = (double)m.w; }
        }
        carryM += tile_mass(T, carryM, M, smd);
        carryPhi += tile_phi(T, M, carryPhi, prev_r, prev_I, Phi, nb_r, nb_I, smd, tile_n);
#pragma unroll
        for (int i = 0; i < HI; i++) vesc2[i] = (T.act[i] && T.r[i] > MACHINE_ZERO) ? (2 * fabs(Phi[i] - Phi0) * P.phi_fac) : 1e30;
        // causal bound mask: inside a thread the members are tested sequentially; across threads the incoming prefix of
        // (M_vel, V) is iterated to its fixed point (a thread's answer only depends on the threads before it)
        bool bound[HI];
#pragma unroll
        for (int i = 0; i < HI; i++) bound[i] = T.act[i];
        double tot[4];
        for (int it = 0; it < HB + 2; it++) {
          double loc[4] = { 0, 0, 0, 0 }, ex[4];
#pragma unroll
          for (int i = 0; i < HI; i++)
            if (bound[i]) { loc[0] += T.w[i]; loc[1] += T.w[i] * mom[i][0]; loc[2] += T.w[i] * mom[i][1]; loc[3] += T.w[i] * mom[i][2]; }
          block_excl_scan_n<4>(loc, ex, tot, smd);
          double m = run0[0] + ex[0], vx = run0[1] + ex[1], vy = run0[2] + ex[2], vz = run0[3] + ex[3];
          bool changed = false;
#pragma unroll
          for (int i = 0; i < HI; i++) {
            bool nbnd = false;
            if (T.act[i]) {
              double dvx = (mom[i][0] - vx / m) * P.v_fac + P.hubble * T.d[i][0] * P.r_fac;
              double dvy = (mom[i][1] - vy / m) * P.v_fac + P.hubble * T.d[i][1] * P.r_fac;
              double dvz = (mom[i][2] - vz / m) * P.v_fac + P.hubble * T.d[i][2] * P.r_fac;
              double vel2 = dvx * dvx + dvy * dvy + dvz * dvz;
              if (has_u) vel2 += (uu[i] < 0.0 ? 0.0 : 2 * uu[i]);
              nbnd = !(vel2 > v2_tune * vesc2[i]);
              if (nbnd) { m += T.w[i]; vx += T.w[i] * mom[i][0]; vy += T.w[i] * mom[i][1]; vz += T.w[i] * mom[i][2]; }
            }
            changed |= (nbnd != bound[i]);
            bound[i] = nbnd;
          }
          if (!__syncthreads_or(changed ? 1 : 0)) break;
        }
        // `tot` was computed with the mask of the last (unchanged) evaluation
        run0[0] += tot[0]; run0[1] += tot[1]; run0[2] += tot[2]; run0[3] += tot[3];
        Mv_acc += tot[0];
        // append bound members in place (nb <= base)
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < HI; i++) cnt += bound[i] ? 1 : 0;
        int totb, posb = block_excl_scan_i(cnt, smi, &totb);
        long long lastj = -1;
#pragma unroll
        for (int i = 0; i < HI; i++)
          if (bound[i]) { ip[nb + posb] = T.pid[i]; posb++; lastj = base + (long long)threadIdx.x * HI + i; }
        long long lb = block_min_ll(-lastj, sml);                     // -(largest bound index), 1 if none
#pragma unroll
        for (int i = 0; i < HI; i++)
          if (bound[i] && base + (long long)threadIdx.x * HI + i == -lb) s_R = T.r[i];
        __syncthreads();
        if (totb > 0) Rv_last = s_R;
        nb += totb;
        nremove += tile_n - totb;
        __syncthreads();
      }
      np = nb; M_vir = Mv_acc; R_vir = Rv_last;
      if ((double)np < (double)P.min_part) break;
    }
  }
  if (threadIdx.x == 0) S[7] = S[8] = S[9] = (double)np;
  if (np >= P.min_part) {
    RvirOut r = rvir_cut(pos4, ip, np, c, P, smd, sml);
    np = r.np; M_vir = r.M; R_vir = r.R; ovd = r.ovd;
  }
  if (threadIdx.x == 0) {
    S[0] = c[0]; S[1] = c[1]; S[2] = c[2];
    S[8] = S[9] = (double)np;
    S[10] = M_vir; S[11] = R_vir; S[12] = ovd; S[13] = Phi0;
    int nbins = 0;
    if (np >= P.min_part) { nbins = (int)(6.2 * (log10((double)np)) - 3.5); if (nbins < 2) nbins = 2; if (nbins > MAXBINS) nbins = MAXBINS; }
    S[57] = (double)nbins;
    npart_out[h] = np;
    iter_work[h] = work;
  }
}

// ------------------------------------------------------------------------------------------------
// U2/U3, cooperative multi-block form (default).  Every halo is cut into tiles of HT radius-sorted members and ALL tiles
// of ALL active haloes run in one grid, so a 10^7-particle host spreads over the whole GPU and the step time no longer
// depends on the largest halo.  Each running quantity of the reference's inside-out loops is a segmented scan in three
// steps: tile totals (k_g_*_a) -> per-halo exclusive scan of the tile totals (k_g_scan, fixed order) -> tile-local scan plus
// carry (k_g_*_c).  The causal bound test of rem_unbound (mask_j depends on the running mean velocity of the BOUND members
// before j, ahf_halos.c:3538-3569) is solved as a fixed point over the whole halo: mask -> prefix sums -> mask, repeated
// until no member changes; because mask_j only depends on mask_<j the fixed point is unique and equals the sequential
// answer.  Deterministic: no floating-point atomics anywhere.
// ------------------------------------------------------------------------------------------------
constexpr int GNC = 5;      // scan components of the mask iteration: M_vel, P(3), bound count

struct GH {                 // per-halo device state of the cooperative pass (arrays indexed by halo)
  int64_t       *np;        // current member count
  double        *Mvir, *Rvir, *ovd, *Phi0, *seed;   // seed: 4 per halo
  unsigned long long *first;                         // rvir: first index below the overdensity limit
  int32_t       *tile0, *ntile;                      // tiles of the halo in the current tile list
  int64_t       *nb, *nremove;                       // result of an unbinding iteration
};

template <int NC>
__global__ void __launch_bounds__(HB) k_g_scan(const int32_t *__restrict__ act, const int32_t *__restrict__ tile0, const int32_t *__restrict__ ntile,
                                               const double *__restrict__ tt, double *__restrict__ tc, double *__restrict__ htot)
{
  __shared__ double smd[(HB / 32) * NC];
  const int h = act[blockIdx.x], t0 = tile0[h], nt = ntile[h];
  double carry[NC];
#pragma unroll
  for (int q = 0; q < NC; q++) carry[q] = 0.0;
  for (int b = 0; b < nt; b += HB) {
    const int t = b + threadIdx.x;
    double v[NC], ex[NC], tot[NC];
#pragma unroll
    for (int q = 0; q < NC; q++) v[q] = t < nt ? tt[(size_t)(t0 + t) * NC + q] : 0.0;
    block_excl_scan_n<NC>(v, ex, tot, smd);
    if (t < nt) {
#pragma unroll
      for (int q = 0; q < NC; q++) tc[(size_t)(t0 + t) * NC + q] = carry[q] + ex[q];
    }
#pragma unroll
    for (int q = 0; q < NC; q++) carry[q] += tot[q];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NC; q++) htot[(size_t)h * NC + q] = carry[q];
  }
}

// G2, tile-parallel: tile = up to HT consecutive candidates of one search cell of one halo.  COUNT pass -> k_g_scan<1> -> FILL pass;
// members are appended range by range in key order, as the reference does (ahf_halos_sfc.c:327-356)
template <bool FILL>
__global__ void __launch_bounds__(HB) k_gather_tiles(const float4 *__restrict__ pos4, const double *__restrict__ centre, const double *__restrict__ grad,
                                                     const int4 *__restrict__ tiles, const int64_t *__restrict__ candoff, const double *__restrict__ tc,
                                                     double *__restrict__ tt, double *__restrict__ r2buf, uint32_t *__restrict__ idxbuf)
{
  __shared__ int smi[HB / 32];
  const int4 tl = tiles[blockIdx.x];
  const int  h = tl.x, cnt = tl.w;
  const uint32_t start = (uint32_t)tl.z;
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const double R2 = grad[h] * grad[h];
  bool in[HI]; double r2[HI]; int mine = 0;
#pragma unroll
  for (int i = 0; i < HI; i++) {
    const int t = threadIdx.x * HI + i;
    in[i] = false; r2[i] = 0.0;
    if (t < cnt) {
      const float4 p = pos4[start + (uint32_t)t];
      double dx = fabs((double)p.x - c[0]), dy = fabs((double)p.y - c[1]), dz = fabs((double)p.z - c[2]);
      if (dx > 0.5) dx = 1.0 - dx; if (dy > 0.5) dy = 1.0 - dy; if (dz > 0.5) dz = 1.0 - dz;
      in[i] = (dx * dx + dy * dy + dz * dz) <= R2;
      if (FILL && in[i]) { double d[3]; sep3(p, c, d); r2[i] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]; }   // sort key (ahf_halos.c:5844)
      mine += in[i] ? 1 : 0;
    }
  }
  int tot, pos = block_excl_scan_i(mine, smi, &tot);
  if (!FILL) { if (threadIdx.x == 0) tt[blockIdx.x] = (double)tot; return; }
  int64_t out = candoff[h] + (int64_t)tc[blockIdx.x] + pos;
#pragma unroll
  for (int i = 0; i < HI; i++) if (in[i]) { r2buf[out] = r2[i]; idxbuf[out] = start + (uint32_t)(threadIdx.x * HI + i); out++; }
}
// tile list of the gather pass, built on the device (round 2: with 2e4 haloes the host version -- 27 ranges per halo read back, a host
// loop, the list uploaded again -- took 9 of the 11 ms of the stage): tiles per (halo, search cell) range -> exclusive scan -> fill
__global__ void k_gt_count(const int64_t *__restrict__ rlo, const int64_t *__restrict__ rhi, int64_t nrange, int *__restrict__ nt)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nrange) { const int64_t len = rhi[i] - rlo[i]; nt[i] = len > 0 ? (int)((len + HT - 1) / HT) : 0; }
}
__global__ void k_gt_fill(const int64_t *__restrict__ rlo, const int64_t *__restrict__ rhi, int64_t nrange, const int *__restrict__ toff, const int *__restrict__ ttot,
                          int4 *__restrict__ tiles, int32_t *__restrict__ act, int32_t *__restrict__ tile0, int32_t *__restrict__ ntile)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nrange) return;
  const int h = (int)(i / 27), q = (int)(i - (int64_t)h * 27);
  int o = toff[i];
  for (int64_t a = rlo[i]; a < rhi[i]; a += HT) { const int64_t len = rhi[i] - a; tiles[o++] = make_int4(h, q, (int)(uint32_t)a, (int)(len < HT ? len : HT)); }
  if (q == 0) {
    act[h] = h; tile0[h] = toff[i];
    const int64_t e = i + 27;
    ntile[h] = (e < nrange ? toff[e] : *ttot) - toff[i];
  }
}
__global__ void k_gather_counts(const double *__restrict__ htot, int64_t nhalo, int64_t *__restrict__ ngather)
{
  const int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (h < nhalo) ngather[h] = (int64_t)htot[h];
}

// cumulative mass of the tile's members: equal masses -> index + 1 exactly; otherwise the stored prefix
__device__ __forceinline__ void g_tile_M(const TileMembers &T, const double *__restrict__ Mpre, long long base, double (&M)[HI])
{
#pragma unroll
  for (int i = 0; i < HI; i++) {
    const long long j = base + (long long)threadIdx.x * HI + i;
    M[i] = Mpre ? (T.act[i] ? Mpre[j] : 0.0) : (double)(j + 1);
  }
}

// multimass only: tile totals of w, then M(<=j) per member
__global__ void __launch_bounds__(HB) k_g_mass_a(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                                                 const uint32_t *__restrict__ members, GH G, const int2 *__restrict__ tiles, double *__restrict__ tt)
{
  __shared__ double smd[HB / 32];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const uint32_t *ip = members + moff0[h];
  double loc = 0.0;
#pragma unroll
  for (int i = 0; i < HI; i++) { long long j = base + (long long)threadIdx.x * HI + i; if (j < np) loc += (double)pos4[ip[j]].w; }
  loc = block_sum(loc, smd);
  if (threadIdx.x == 0) tt[blockIdx.x] = loc;
}
__global__ void __launch_bounds__(HB) k_g_mass_c(const float4 *__restrict__ pos4, const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members, GH G,
                                                 const int2 *__restrict__ tiles, const double *__restrict__ tc, double *__restrict__ Mpre)
{
  __shared__ double smd[HB / 32];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const uint32_t *ip = members + moff0[h];
  double w[HI], loc[1] = { 0.0 }, ex[1], tot[1];
#pragma unroll
  for (int i = 0; i < HI; i++) { long long j = base + (long long)threadIdx.x * HI + i; w[i] = j < np ? (double)pos4[ip[j]].w : 0.0; loc[0] += w[i]; }
  block_excl_scan_n<1>(loc, ex, tot, smd);
  double run = tc[blockIdx.x] + ex[0];
#pragma unroll
  for (int i = 0; i < HI; i++) { long long j = base + (long long)threadIdx.x * HI + i; run += w[i]; if (j < np) Mpre[moff0[h] + j] = run; }
}

// rem_outsideRvir (ahf_halos.c:3811-3869): first member whose mean enclosed overdensity drops below ovlim
__global__ void __launch_bounds__(HB) k_g_rvir_find(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                                                    const uint32_t *__restrict__ members, GH G, const int2 *__restrict__ tiles, const double *__restrict__ Mpre, HP P)
{
  __shared__ long long sml[HB / 32];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  TileMembers T;
  load_tile(T, pos4, members + moff0[h], base, np, c);
  double M[HI];
  g_tile_M(T, Mpre ? Mpre + moff0[h] : nullptr, base, M);
  long long mine = 0x7fffffffffffffffll;
#pragma unroll
  for (int i = 0; i < HI; i++)
    if (T.act[i]) {
      const double V = 4. * PI_ / 3. * (T.r[i] * T.r[i] * T.r[i]), od = M[i] / V * P.rho_fac / P.rho_vir;
      if (!(od >= P.ovlim) && mine == 0x7fffffffffffffffll) mine = base + (long long)threadIdx.x * HI + i;
    }
  const long long first = block_min_ll(mine, sml);
  if (threadIdx.x == 0 && first != 0x7fffffffffffffffll) atomicMin(&G.first[h], (unsigned long long)first);
}
__global__ void k_g_rvir_apply(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                               const uint32_t *__restrict__ members, GH G, const int32_t *__restrict__ act, int nact, const double *__restrict__ Mpre, HP P)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nact) return;
  const int h = act[a];
  const long long np = G.np[h];
  const unsigned long long f = G.first[h];
  const long long js = (f == ~0ull) ? np - 1 : (long long)f;
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const double r = dist3(pos4[members[moff0[h] + js]], c), M = Mpre ? Mpre[moff0[h] + js] : (double)(js + 1);
  const double V = 4. * PI_ / 3. * (r * r * r);
  G.np[h] = js + 1; G.Mvir[h] = M; G.Rvir[h] = r; G.ovd[h] = M / V * P.rho_fac / P.rho_vir;
  G.first[h] = ~0ull;
}

// trapezoid terms of Phi for one tile (ahf_halos.c:3398-3410): (r, I) of the member before the tile come from that member itself
__device__ __forceinline__ double g_tile_phi(const TileMembers &T, const double (&M)[HI], double prev_r, double prev_I, double (&term)[HI],
                                             double *nb_r, double *nb_I)
{
  double I[HI];
#pragma unroll
  for (int i = 0; i < HI; i++) I[i] = (T.act[i] && T.r[i] > MACHINE_ZERO) ? M[i] / (T.r[i] * T.r[i]) : 0.0;
  nb_r[threadIdx.x] = T.r[HI - 1]; nb_I[threadIdx.x] = I[HI - 1];
  __syncthreads();
  double rp = threadIdx.x ? nb_r[threadIdx.x - 1] : prev_r, Ip = threadIdx.x ? nb_I[threadIdx.x - 1] : prev_I;
  double loc = 0.0;
#pragma unroll
  for (int i = 0; i < HI; i++) {
    term[i] = (T.act[i] && T.r[i] > MACHINE_ZERO) ? ((I[i] + Ip) / 2.) * (T.r[i] - rp) : 0.0;
    loc += term[i];
    rp = T.r[i]; Ip = I[i];
  }
  return loc;
}
__device__ __forceinline__ void g_prev_member(const float4 *__restrict__ pos4, const uint32_t *__restrict__ ip, const double *__restrict__ Mpre,
                                              long long base, const double c[3], double &prev_r, double &prev_I)
{
  prev_r = 0.0; prev_I = 0.0;
  if (base > 0) {
    prev_r = dist3(pos4[ip[base - 1]], c);
    const double Mp = Mpre ? Mpre[base - 1] : (double)base;
    prev_I = prev_r > MACHINE_ZERO ? Mp / (prev_r * prev_r) : 0.0;
  }
}
__global__ void __launch_bounds__(HB) k_g_phi_a(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                                                const uint32_t *__restrict__ members, GH G, const int2 *__restrict__ tiles, const double *__restrict__ Mpre,
                                                double *__restrict__ tt)
{
  __shared__ double smd[HB / 32];
  __shared__ double nb_r[HB], nb_I[HB];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const uint32_t *ip = members + moff0[h];
  const double   *Mp = Mpre ? Mpre + moff0[h] : nullptr;
  TileMembers T;
  load_tile(T, pos4, ip, base, np, c);
  double M[HI], term[HI], prev_r, prev_I;
  g_tile_M(T, Mp, base, M);
  g_prev_member(pos4, ip, Mp, base, c, prev_r, prev_I);
  double loc = g_tile_phi(T, M, prev_r, prev_I, term, nb_r, nb_I);
  loc = block_sum(loc, smd);
  if (threadIdx.x == 0) tt[blockIdx.x] = loc;
}
// Phi0 = sum of all trapezoids + M_tot / r_last (:3423-3426)
__global__ void k_g_phi0(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members,
                         GH G, const int32_t *__restrict__ act, int nact, const double *__restrict__ Mpre, const double *__restrict__ htot)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nact) return;
  const int h = act[a];
  const long long np = G.np[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const double r = dist3(pos4[members[moff0[h] + np - 1]], c), M = Mpre ? Mpre[moff0[h] + np - 1] : (double)np;
  G.Phi0[h] = htot[h] + M / r;
}
__global__ void __launch_bounds__(HB) k_g_phi_c(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                                                const uint32_t *__restrict__ members, GH G, const int2 *__restrict__ tiles, const double *__restrict__ Mpre,
                                                const double *__restrict__ tc, HP P, double *__restrict__ vesc2)
{
  __shared__ double smd[HB / 32];
  __shared__ double nb_r[HB], nb_I[HB];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const uint32_t *ip = members + moff0[h];
  const double   *Mp = Mpre ? Mpre + moff0[h] : nullptr;
  TileMembers T;
  load_tile(T, pos4, ip, base, np, c);
  double M[HI], term[HI], prev_r, prev_I;
  g_tile_M(T, Mp, base, M);
  g_prev_member(pos4, ip, Mp, base, c, prev_r, prev_I);
  double loc[1], ex[1], tot[1];
  loc[0] = g_tile_phi(T, M, prev_r, prev_I, term, nb_r, nb_I);
  block_excl_scan_n<1>(loc, ex, tot, smd);
  double run = tc[blockIdx.x] + ex[0];
  const double Phi0 = G.Phi0[h];
#pragma unroll
  for (int i = 0; i < HI; i++) {
    run += term[i];
    const long long j = base + (long long)threadIdx.x * HI + i;
    if (T.act[i]) vesc2[moff0[h] + j] = T.r[i] > MACHINE_ZERO ? (2 * fabs(run - Phi0) * P.phi_fac) : 1e30;      // :3514-3530
  }
}
// seed of the running bulk velocity (:3441-3470)
__global__ void k_g_seed(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members,
                         GH G, const int32_t *__restrict__ act, int nact, int first_iter, int min_part)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nact) return;
  const int h = act[a];
  const uint32_t *ip = members + moff0[h];
  long long seed = 0;
  if (first_iter) {
    int nv = min_part / 2;
    float m2[64]; int id[64];
    if (nv > 64) nv = 64;
    for (int q = 0; q < nv; q++) { float4 m = mom4[ip[q]]; m2[q] = m.x * m.x + m.y * m.y + m.z * m.z; id[q] = q; }
    for (int x = 1; x < nv; x++) { int t = id[x]; float v = m2[t]; int b = x - 1; while (b >= 0 && m2[id[b]] > v) { id[b + 1] = id[b]; b--; } id[b + 1] = t; }
    seed = id[nv / 2 - 1];                                  // NR indexx is 1-based: idx[n/2] = (n/2)-th smallest
  }
  const float4 p = pos4[ip[seed]], m = mom4[ip[seed]];
  const double w = (double)p.w;
  G.seed[4 * h] = w; G.seed[4 * h + 1] = w * m.x; G.seed[4 * h + 2] = w * m.y; G.seed[4 * h + 3] = w * m.z;
}
// one sweep of the mask fixed point.  FIRST: masks start as "all bound", only the tile totals are produced.
// Otherwise: prefix of the bound (M_vel, P) from the previous masks -> new masks (iterated inside the tile until stable for the
// given carry-in) -> their tile totals for the next sweep; *changed is raised when any mask differs from the one it replaces.
template <bool FIRST>
__global__ void __launch_bounds__(HB, 2) k_g_mask(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, int has_u, const double *__restrict__ centre,
                                               const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members, GH G, const int2 *__restrict__ tiles,
                                               const double *__restrict__ vesc2, const double *__restrict__ tc, HP P, uint8_t *__restrict__ mask,
                                               double *__restrict__ tt, int *__restrict__ changed)
{
  __shared__ double smd[(HB / 32) * GNC];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const uint32_t *ip = members + moff0[h];
  TileMembers T;
  load_tile(T, pos4, ip, base, np, c);
  double mom[HI][3], uu[HI];
#pragma unroll
  for (int i = 0; i < HI; i++) {
    mom[i][0] = mom[i][1] = mom[i][2] = 0.0; uu[i] = -1.0;
    if (T.act[i]) { float4 m = mom4[T.pid[i]]; mom[i][0] = (double)m.x; mom[i][1] = (double)m.y; mom[i][2] = (double)m.z; uu[i] = (double)m.w; }
  }
  bool bound[HI], was[HI];
#pragma unroll
  for (int i = 0; i < HI; i++) {
    const long long j = base + (long long)threadIdx.x * HI + i;
    bound[i] = T.act[i] && (FIRST ? true : mask[moff0[h] + j] != 0);
    was[i] = bound[i];
  }
  double tot[GNC];
  if (!FIRST) {
    const double v2_tune = P.vesc_tune * P.vesc_tune;
    double ve[HI];
#pragma unroll
    for (int i = 0; i < HI; i++) { const long long j = base + (long long)threadIdx.x * HI + i; ve[i] = T.act[i] ? vesc2[moff0[h] + j] : 1e30; }
    double run0[4];
#pragma unroll
    for (int q = 0; q < 4; q++) run0[q] = G.seed[4 * h + q] + tc[(size_t)blockIdx.x * GNC + q];
    for (int it = 0; it < HB + 2; it++) {
      double loc[4] = { 0, 0, 0, 0 }, ex[4], t4[4];
#pragma unroll
      for (int i = 0; i < HI; i++)
        if (bound[i]) { loc[0] += T.w[i]; loc[1] += T.w[i] * mom[i][0]; loc[2] += T.w[i] * mom[i][1]; loc[3] += T.w[i] * mom[i][2]; }
      block_excl_scan_n<4>(loc, ex, t4, smd);
      double m = run0[0] + ex[0], vx = run0[1] + ex[1], vy = run0[2] + ex[2], vz = run0[3] + ex[3];
      bool chg = false;
#pragma unroll
      for (int i = 0; i < HI; i++) {
        bool nbnd = false;
        if (T.act[i]) {
          double dvx = (mom[i][0] - vx / m) * P.v_fac + P.hubble * T.d[i][0] * P.r_fac;
          double dvy = (mom[i][1] - vy / m) * P.v_fac + P.hubble * T.d[i][1] * P.r_fac;
          double dvz = (mom[i][2] - vz / m) * P.v_fac + P.hubble * T.d[i][2] * P.r_fac;
          double vel2 = dvx * dvx + dvy * dvy + dvz * dvz;
          if (has_u) vel2 += (uu[i] < 0.0 ? 0.0 : 2 * uu[i]);
          nbnd = !(vel2 > v2_tune * ve[i]);
          if (nbnd) { m += T.w[i]; vx += T.w[i] * mom[i][0]; vy += T.w[i] * mom[i][1]; vz += T.w[i] * mom[i][2]; }
        }
        chg |= (nbnd != bound[i]);
        bound[i] = nbnd;
      }
      if (!__syncthreads_or(chg ? 1 : 0)) break;
    }
    bool diff = false;
#pragma unroll
    for (int i = 0; i < HI; i++) {
      diff |= (bound[i] != was[i]);
      const long long j = base + (long long)threadIdx.x * HI + i;
      if (T.act[i] && bound[i] != was[i]) mask[moff0[h] + j] = bound[i] ? 1 : 0;
    }
    if (__syncthreads_or(diff ? 1 : 0) && threadIdx.x == 0) atomicOr(changed, 1);
  }
  double loc[GNC] = { 0, 0, 0, 0, 0 };
#pragma unroll
  for (int i = 0; i < HI; i++)
    if (bound[i]) { loc[0] += T.w[i]; loc[1] += T.w[i] * mom[i][0]; loc[2] += T.w[i] * mom[i][1]; loc[3] += T.w[i] * mom[i][2]; loc[4] += 1.0; }
  block_sum_n<GNC>(loc, smd);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < GNC; q++) tt[(size_t)blockIdx.x * GNC + q] = loc[q];
  }
  (void)tot;
}
// compaction of the bound members (tc[.][4] = bound members of the halo before this tile) into a scratch list
__global__ void __launch_bounds__(HB) k_g_compact(const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members, GH G, const int2 *__restrict__ tiles,
                                                  const uint8_t *__restrict__ mask, const double *__restrict__ tc, uint32_t *__restrict__ out)
{
  __shared__ int smi[HB / 32];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  uint32_t pid[HI]; bool b[HI]; int cnt = 0;
#pragma unroll
  for (int i = 0; i < HI; i++) {
    const long long j = base + (long long)threadIdx.x * HI + i;
    b[i] = j < np && mask[moff0[h] + j] != 0; pid[i] = b[i] ? members[moff0[h] + j] : 0u; cnt += b[i] ? 1 : 0;
  }
  int tot, pos = block_excl_scan_i(cnt, smi, &tot);
  const long long o = moff0[h] + (long long)tc[(size_t)blockIdx.x * GNC + 4];
#pragma unroll
  for (int i = 0; i < HI; i++) if (b[i]) { out[o + pos] = pid[i]; pos++; }
}
__global__ void __launch_bounds__(HB) k_g_copyback(const int64_t *__restrict__ moff0, GH G, const int2 *__restrict__ tiles, const double *__restrict__ htot,
                                                   const uint32_t *__restrict__ in, uint32_t *__restrict__ members)
{
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, nb = (long long)htot[(size_t)h * GNC + 4];
#pragma unroll
  for (int i = 0; i < HI; i++) { const long long j = base + (long long)threadIdx.x + (long long)i * HB; if (j < nb) members[moff0[h] + j] = in[moff0[h] + j]; }
}
// end of an unbinding iteration (:3583-3599): npart, M_vir, R_vir of the bound set; Phi0 stays the one of this iteration
__global__ void k_g_iter_finish(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                                const uint32_t *__restrict__ members, GH G, const int32_t *__restrict__ act, int nact, const double *__restrict__ htot)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nact) return;
  const int h = act[a];
  const long long np = G.np[h], nb = (long long)htot[(size_t)h * GNC + 4];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  G.nb[h] = nb; G.nremove[h] = np - nb;
  G.Mvir[h] = htot[(size_t)h * GNC + 0];
  G.Rvir[h] = nb > 0 ? dist3(pos4[members[moff0[h] + nb - 1]], c) : 0.0;
  G.np[h] = nb;
}
__global__ void k_g_write_scal(GH G, const double *__restrict__ centre, const int64_t *__restrict__ ngather, const int64_t *__restrict__ n6, const int64_t *__restrict__ n7,
                               int64_t nhalo, int min_part, double *__restrict__ scal, int64_t *__restrict__ npart_out)
{
  const int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (h >= nhalo) return;
  double *S = scal + h * AHFGPU_NSCAL;
  const long long np = G.np[h];
  S[0] = centre[3 * h]; S[1] = centre[3 * h + 1]; S[2] = centre[3 * h + 2];
  S[5] = (double)ngather[h]; S[6] = (double)n6[h]; S[7] = (double)n7[h]; S[8] = S[9] = (double)np;
  S[10] = G.Mvir[h]; S[11] = G.Rvir[h]; S[12] = G.ovd[h]; S[13] = G.Phi0[h];
  int nbins = 0;
  if (np >= min_part) { nbins = (int)(6.2 * (log10((double)np)) - 3.5); if (nbins < 2) nbins = 2; if (nbins > MAXBINS) nbins = MAXBINS; }
  S[57] = (double)nbins;
  npart_out[h] = np;
}

// ------------------------------------------------------------------------------------------------
// P1 helpers: 3x3 Jacobi (general.c:1163-1240), get_axes (specific.c:135-178), calc_cNFW / R1 (specific.c:1972-2072)
// ------------------------------------------------------------------------------------------------
__device__ void jacobi3(double a[3][3], double d[3], double v[3][3])
{
  double b[3], z[3];
  for (int ip = 0; ip < 3; ip++) { for (int iq = 0; iq < 3; iq++) v[ip][iq] = 0.0; v[ip][ip] = 1.0; }
  for (int ip = 0; ip < 3; ip++) { b[ip] = d[ip] = a[ip][ip]; z[ip] = 0.0; }
  for (int i = 1; i <= 50; i++) {
    double sm = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (sm == 0.0) return;
    double tresh = (i < 4) ? 0.2 * sm / 9 : 0.0;
    for (int ip = 0; ip < 2; ip++) for (int iq = ip + 1; iq < 3; iq++) {
      double g = 100.0 * fabs(a[ip][iq]);
      if (i > 4 && (fabs(d[ip]) + g) == fabs(d[ip]) && (fabs(d[iq]) + g) == fabs(d[iq])) a[ip][iq] = 0.0;
      else if (fabs(a[ip][iq]) > tresh) {
        double hh = d[iq] - d[ip], t;
        if ((fabs(hh) + g) == fabs(hh)) t = (a[ip][iq]) / hh;
        else {
          double theta = 0.5 * hh / (a[ip][iq]);
          t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
          if (theta < 0.0) t = -t;
        }
        double cc = 1.0 / sqrt(1 + t * t), s = t * cc, tau = s / (1.0 + cc);
        hh = t * a[ip][iq];
        z[ip] -= hh; z[iq] += hh; d[ip] -= hh; d[iq] += hh; a[ip][iq] = 0.0;
#define ROT(Mx, i1, j1, k1, l1) { double g1 = Mx[i1][j1], h1 = Mx[k1][l1]; Mx[i1][j1] = g1 - s * (h1 + g1 * tau); Mx[k1][l1] = h1 + s * (g1 - h1 * tau); }
        for (int j = 0; j <= ip - 1; j++) ROT(a, j, ip, j, iq)
        for (int j = ip + 1; j <= iq - 1; j++) ROT(a, ip, j, j, iq)
        for (int j = iq + 1; j < 3; j++) ROT(a, ip, j, iq, j)
        for (int j = 0; j < 3; j++) ROT(v, j, ip, j, iq)
#undef ROT
      }
    }
    for (int ip = 0; ip < 3; ip++) { b[ip] += z[ip]; d[ip] = b[ip]; z[ip] = 0.0; }
  }
}
__device__ void get_axes(double it[3][3], double &ax1, double &ax2, double &ax3)
{
  double a[3][3], d[3], v[3][3];
  int    idx[3] = { 0, 1, 2 };
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = it[i][j];
  jacobi3(a, d, v);
  for (int j = 1; j < 3; j++) { int t = idx[j]; double av = d[t]; int i = j - 1; for (; i >= 0; i--) { if (d[idx[i]] <= av) break; idx[i + 1] = idx[i]; } idx[i + 1] = t; }
  ax1 = d[idx[2]]; ax2 = d[idx[1]]; ax3 = d[idx[0]];
  for (int i = 0; i < 3; i++) { it[i][0] = v[i][idx[2]]; it[i][1] = v[i][idx[1]]; it[i][2] = v[i][idx[0]]; }
}
__device__ double cnfw_root(double c, double r) { return 0.216 * c / (log(1 + c) - c / (1 + c)) - r; }
__device__ double calc_cNFW(double V2_max, double V2_vir)
{
  double r = V2_max / V2_vir, a = 2.2, b = 100, c;
  if (r <= 1 || r > 5.9) return -1;
  while (b - a > 1e-3) { c = (a + b) / 2; if (cnfw_root(a, r) * cnfw_root(c, r) > 0) a = c; else b = c; }
  return (a + b) / 2.0;
}
__device__ double cr1_root(double c, double R1) { return (c - 2.0 * log(1.0 + c) + c / (1.0 + c)) / (c * (log(1.0 + c) - c / (1.0 + c))) - R1; }

// ------------------------------------------------------------------------------------------------
// stage kernel 2: HaloProfiles (ahf_halos.c:3961-5018).  One CTA per halo.
//   per member prefix quantities (M, P, Phi) are block scans with carries; everything that is only sampled at bin
//   edges (CoM, inertia tensor, L, Ekin, Epot ...) is reduced per (tile, bin) in a fixed order and prefix-summed over bins.
// ------------------------------------------------------------------------------------------------
constexpr int NACC = 21;   // CoM3, a11 a22 a33 a12 a13 a23, L3, Ekin, Epot, Mhires, Mlores, M, npart, M_gas, M_star, u_gas (GAS_PARTICLES build)
constexpr int NSPC = 19;   // per species: n, M, com(3), P(3), L(3), a11 a22 a33 a12 a13 a23, Epot, Ekin

// per-bin cumulative values, Jacobi, profile columns and the integral properties (ahf_halos.c:4632-4710, :4870-5018); one thread
// one radial bin of HaloProfiles from the CUMULATIVE sums up to and including it (cum), the mass / volume of the previous bin, the
// escape velocity of the last non-empty bin so far and the bin's own u_gas sum
#define PR(col, bb) pr[(col) * nbins + (bb)]
__device__ __forceinline__ void prof_bin(const int b, const int nbins, const double *cum, const double shell_u, const double cur_rad, const double M_prev,
                                         const double V_prev, const double vesc_run, double *pr, double *prsp)
{
  const double F43 = 4. * PI_ / 3.;
  const double M = cum[16], Volume = F43 * (cur_rad * cur_rad * cur_rad), dM = M - M_prev, dV = Volume - V_prev;
  double it[3][3], ax1, ax2, ax3;
  if (cum[17] > (double)MINPART_SHELL) {
    it[0][0] = cum[3]; it[1][1] = cum[4]; it[2][2] = cum[5]; it[0][1] = it[1][0] = cum[6]; it[0][2] = it[2][0] = cum[7]; it[1][2] = it[2][1] = cum[8];
    get_axes(it, ax1, ax2, ax3);
  } else { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) it[i][j] = 0.0; ax1 = 1; ax2 = 0; ax3 = 0; }
  PR(0, b) = cum[17]; PR(1, b) = cur_rad; PR(2, b) = M; PR(3, b) = M / Volume; PR(4, b) = (dV > 0) ? dM / dV : 0.0;
  PR(5, b) = M / cur_rad; PR(6, b) = vesc_run; PR(7, b) = sqrt(cum[12] / M); PR(8, b) = 0.5 * cum[12]; PR(9, b) = 0.5 * cum[13];
  PR(10, b) = cum[9]; PR(11, b) = cum[10]; PR(12, b) = cum[11];
  PR(13, b) = 1.0; PR(14, b) = it[0][0]; PR(15, b) = it[1][0]; PR(16, b) = it[2][0];
  PR(17, b) = (ax1 > 0.) ? sqrt(ax2 / ax1) : 0.0; PR(18, b) = it[0][1]; PR(19, b) = it[1][1]; PR(20, b) = it[2][1];
  PR(21, b) = (ax1 > 0.) ? sqrt(ax3 / ax1) : 0.0; PR(22, b) = it[0][2]; PR(23, b) = it[1][2]; PR(24, b) = it[2][2];
  if (prsp) { prsp[0 * nbins + b] = cum[18]; prsp[1 * nbins + b] = cum[19]; prsp[2 * nbins + b] = shell_u; }   // M_gas, M_star cumulative; u_gas of the shell (:4713-4715)
}
__device__ void prof_tail(const int nbins, const double *cum, const double *Pl, double *pr, double *S, const double R_vir, const HP &P, const long long best_j,
                          const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, const uint32_t *__restrict__ ip, const double c[3]);

// serial form (one thread): the bins in order, then the integral properties
__device__ void prof_finalize(const int nbins, const double (*acc)[NACC], const double *edge, const double *vesc_bin, const double (*Vc_bin)[3],
                              double *pr, double *S, const double R_vir, const HP &P, const long long best_j, const float4 *__restrict__ pos4,
                              const float4 *__restrict__ mom4, const uint32_t *__restrict__ ip, const double c[3], double *prsp = nullptr)
{
  const double F43 = 4. * PI_ / 3.;
  double cum[NACC];
  for (int q = 0; q < NACC; q++) cum[q] = 0.0;
  double M_prev = 0.0, V_prev = 0.0, vesc_run = 0.0, Pl[3] = { 0, 0, 0 };
  for (int b = 0; b < nbins; b++) {
    for (int q = 0; q < NACC; q++) cum[q] += acc[b][q];
    if (acc[b][17] > 0.0) { vesc_run = vesc_bin[b]; Pl[0] = Vc_bin[b][0]; Pl[1] = Vc_bin[b][1]; Pl[2] = Vc_bin[b][2]; }
    const double cur_rad = edge[b];
    prof_bin(b, nbins, cum, acc[b][20], cur_rad, M_prev, V_prev, vesc_run, pr, prsp);
    M_prev = cum[16]; V_prev = F43 * (cur_rad * cur_rad * cur_rad);
  }
  prof_tail(nbins, cum, Pl, pr, S, R_vir, P, best_j, pos4, mom4, ip, c);
}

// integral properties of the halo from the finished profile (ahf_halos.c:4870-5018)
__device__ void prof_tail(const int nbins, const double *cum, const double *Pl, double *pr, double *S, const double R_vir, const HP &P, const long long best_j,
                          const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, const uint32_t *__restrict__ ip, const double c[3])
{
    const double M = cum[16];
    const int    lb = nbins - 1;
    double CoM[3];
    for (int q = 0; q < 3; q++) CoM[q] = fmod(cum[q] / M + 1., 1.);
    double absL = sqrt(PR(10, lb) * PR(10, lb) + PR(11, lb) * PR(11, lb) + PR(12, lb) * PR(12, lb));
    S[10] = M; S[14] = Pl[0] / M; S[15] = Pl[1] / M; S[16] = Pl[2] / M;
    S[17] = PR(7, lb); S[18] = PR(6, lb); S[24] = PR(8, lb); S[25] = PR(9, lb);
    if (absL > 0) {
      S[38] = PR(10, lb) / absL; S[39] = PR(11, lb) / absL; S[40] = PR(12, lb) / absL;
      double lam = absL / M / sqrt(2. * M * R_vir);
      lam *= P.v_fac * sqrt(P.r_fac / (GRAV_ * P.m_fac));
      S[22] = lam;
      double t1 = sqrt(P.m_fac * M); t1 = t1 * t1 * t1;
      double t2 = S[24] * P.m_fac * (P.v_fac * P.v_fac), t3 = S[25] * P.m_fac * P.phi_fac;
      t2 = sqrt(fabs(t2 + t3)); t1 = t2 / t1; t2 = P.m_fac * P.r_fac * P.v_fac * absL; t2 = t2 / (P.m_fac * M);
      S[23] = t1 * t2 / GRAV_;
    } else { S[38] = S[39] = S[40] = 0.0; S[22] = S[23] = 0.0; }
    S[41] = PR(13, lb); S[42] = PR(17, lb); S[43] = PR(21, lb);
    S[44] = PR(14, lb); S[45] = PR(15, lb); S[46] = PR(16, lb); S[47] = PR(18, lb); S[48] = PR(19, lb); S[49] = PR(20, lb);
    S[50] = PR(22, lb); S[51] = PR(23, lb); S[52] = PR(24, lb);
    {
      double R1 = (PR(1, 0) / 2.0) * (PR(1, 0) / 2.0) * (PR(1, 0) / 2.0) * PR(4, 0) * PR(1, 0);
      for (int b = 1; b < nbins; b++) { double rmid = (PR(1, b) + PR(1, b - 1)) / 2.0, dr = PR(1, b) - PR(1, b - 1); R1 += (rmid * rmid * rmid) * PR(4, b) * dr; }
      R1 = 4 * PI_ * R1 / M / R_vir;
      S[56] = R1;
      if (R1 <= 0.19 || 0.585 <= R1) S[55] = -1.0;
      else { double a = 1.0, b2 = 500.0, cc; while (b2 - a > 1e-3) { cc = (a + b2) / 2; if (cr1_root(a, R1) * cr1_root(cc, R1) > 0) a = cc; else b2 = cc; } S[55] = (a + b2) / 2.0; }
    }
    {
      double Ts = 2.0 * (PR(8, lb) - PR(8, lb - 1)), fr = fabs(PR(1, lb - 1) / PR(1, lb));
      S[26] = -0.125 * ((1. + fr) * (1. + fr) * (1. + fr)) / (1. - (fr * fr * fr)) * Ts;
    }
    S[53] = (cum[14] > 0) ? cum[14] / (cum[14] + cum[15]) : 0.0;
    if (best_j >= 0) {
      float4 p = pos4[ip[best_j]], m = mom4[ip[best_j]];
      double dx = fabs((double)p.x - c[0]), dy = fabs((double)p.y - c[1]), dz = fabs((double)p.z - c[2]);
      if (dx > 0.5) dx -= 1.0; if (dy > 0.5) dy -= 1.0; if (dz > 0.5) dz -= 1.0;
      S[37] = sqrt(dx * dx + dy * dy + dz * dz);
      S[31] = p.x; S[32] = p.y; S[33] = p.z; S[34] = m.x; S[35] = m.y; S[36] = m.z;
    } else S[37] = -1.0;
    {
      double dx = fabs(CoM[0] - c[0]), dy = fabs(CoM[1] - c[1]), dz = fabs(CoM[2] - c[2]);
      if (dx > 0.5) dx -= 1.0; if (dy > 0.5) dy -= 1.0; if (dz > 0.5) dz -= 1.0;
      S[30] = sqrt(dx * dx + dy * dy + dz * dz);
      S[27] = CoM[0]; S[28] = CoM[1]; S[29] = CoM[2];
    }
#undef PR
}


__global__ void __launch_bounds__(HB) k_halo_profiles(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, int has_w, int has_u,
                                                      const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                                                      const uint32_t *__restrict__ members, const int64_t *__restrict__ npart_in, HP P,
                                                      double *__restrict__ scal, const int64_t *__restrict__ poff, double *__restrict__ prof,
                                                      const int64_t *__restrict__ soff, double *__restrict__ scratch)
{
  __shared__ double smd[(HB / 32) * NACC];
  __shared__ double edge[MAXBINS];
  __shared__ double acc[MAXBINS][NACC];
  __shared__ double vesc_bin[MAXBINS], Vc_bin[MAXBINS][3];
  __shared__ double nb_r[HB + 1], nb_I[HB + 1];
  __shared__ double s_dmin, s_dmax;
  __shared__ double s_emin[HB / 32]; __shared__ long long s_eidx[HB / 32];
  const int64_t h = blockIdx.x;
  const long long np = npart_in[h];
  double *S = scal + h * AHFGPU_NSCAL;
  if (np < P.min_part) return;
  const uint32_t *ip = members + moff0[h];
  const double    c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const int       nbins = (int)S[57];
  const double    Phi0 = S[13], R_vir = S[11];
  double *pr = prof + poff[h] * AHFGPU_NPROFCOL;
  // per-member work arrays: r, y (value), t (smoothing ping-pong)
  double *w_r = scratch + soff[h] * 3, *w_y = w_r + np, *w_t = w_y + np;
  const double F43 = 4. * PI_ / 3.;
  // binning_parameter (specific.c:259-322)
  if (threadIdx.x == 0) {
    long long k = (long long)floor(((double)P.min_part / 10.) + 0.5);
    double dmin = -1.0;
    while (k < np - 1 && dmin < MACHINE_ZERO) { dmin = dist3(pos4[ip[k]], c); k++; }
    double dmax = dist3(pos4[ip[np - 1]], c);
    if (dmin < MACHINE_ZERO) dmin = dmax / 2.;
    s_dmin = dmin; s_dmax = dmax;
    double ldmin = log10(dmin), ldmax = log10(dmax), ldr = (ldmax - ldmin) / (double)nbins;
    for (int b = 0; b < nbins; b++) edge[b] = pow(10., ldmin + ((double)b + 1) * ldr);
    edge[nbins - 1] = dmax + ZERO_F;                                       // :4226-4241
  }
  for (int i = threadIdx.x; i < MAXBINS * NACC; i += HB) (&acc[0][0])[i] = 0.0;
  for (int i = threadIdx.x; i < MAXBINS; i += HB) { vesc_bin[i] = 0.0; Vc_bin[i][0] = Vc_bin[i][1] = Vc_bin[i][2] = 0.0; }
  __syncthreads();
  double carryM = 0.0, carryPhi = 0.0, cP[3] = { 0.0, 0.0, 0.0 }, prev_r = 0.0, prev_I = 0.0;
  double best_e = 1e30; long long best_j = -1;
  __shared__ int s_binlo, s_binhi;
  for (long long base = 0; base < np; base += HT) {
    TileMembers T;
    load_tile(T, pos4, ip, base, np, c);
    const long long tile_n = (base + HT < np ? base + HT : np) - base;
    double mom[HI][3], uu[HI], M[HI], Phi[HI], Pc[HI][3];
#pragma unroll
    for (int i = 0; i < HI; i++) {
      mom[i][0] = mom[i][1] = mom[i][2] = 0.0; uu[i] = -1.0;
      if (T.act[i]) { float4 m = mom4[T.pid[i]]; mom[i][0] = (double)m.x; mom[i][1] = (double)m.y; mom[i][2] = (double)m.z; uu[i] = (double)m.w; }
    }
    carryM += tile_mass(T, carryM, M, smd);
    {
      double loc[3] = { 0, 0, 0 }, ex[3], tot[3];
#pragma unroll
      for (int i = 0; i < HI; i++) { loc[0] += T.w[i] * mom[i][0]; loc[1] += T.w[i] * mom[i][1]; loc[2] += T.w[i] * mom[i][2]; }
      block_excl_scan_n<3>(loc, ex, tot, smd);
      double run[3] = { cP[0] + ex[0], cP[1] + ex[1], cP[2] + ex[2] };
#pragma unroll
      for (int i = 0; i < HI; i++) {
#pragma unroll
        for (int q = 0; q < 3; q++) { run[q] += T.w[i] * mom[i][q]; Pc[i][q] = run[q]; }
      }
      cP[0] += tot[0]; cP[1] += tot[1]; cP[2] += tot[2];
    }
    const double rr_in = (base == 0) ? -1.0 : prev_r;                 // r_{j-1} of the tile's first member (r_{-1} = -1)
    carryPhi += tile_phi(T, M, carryPhi, prev_r, prev_I, Phi, nb_r, nb_I, smd, tile_n);
    // bin of every member: the first bin whose edge exceeds r_{j-1} (ahf_halos.c:4283 `while (cur_dist < cur_rad)`)
    int bin[HI];
    {
      double rp = threadIdx.x ? nb_r[threadIdx.x - 1] : rr_in;
      int b = 0;
#pragma unroll
      for (int i = 0; i < HI; i++) {
        while (b < nbins - 1 && !(rp < edge[b])) b++;
        bin[i] = b;
        rp = T.r[i];
      }
    }
    if (threadIdx.x == 0) s_binlo = bin[0];
    if (threadIdx.x == (tile_n - 1) / HI) s_binhi = bin[(tile_n - 1) % HI];
    __syncthreads();
    const int blo = s_binlo, bhi = s_binhi;
    // per member energies (needed for the bin sums, v_esc and the most bound member)
    double Tp[HI], Up[HI], Lm[HI][3];
#pragma unroll
    for (int i = 0; i < HI; i++) {
      Tp[i] = Up[i] = 0.0; Lm[i][0] = Lm[i][1] = Lm[i][2] = 0.0;
      if (T.act[i]) {
        const double w = T.w[i];
        double dvx = mom[i][0] - Pc[i][0] / M[i], dvy = mom[i][1] - Pc[i][1] / M[i], dvz = mom[i][2] - Pc[i][2] / M[i];   // :4363-4370 mean INCLUDING j
        Lm[i][0] = w * (T.d[i][1] * dvz - T.d[i][2] * dvy);
        Lm[i][1] = w * (T.d[i][2] * dvx - T.d[i][0] * dvz);
        Lm[i][2] = w * (T.d[i][0] * dvy - T.d[i][1] * dvx);
        dvx += P.hubble * T.d[i][0] * P.r_fac / P.v_fac; dvy += P.hubble * T.d[i][1] * P.r_fac / P.v_fac; dvz += P.hubble * T.d[i][2] * P.r_fac / P.v_fac;
        Tp[i] = w * (dvx * dvx + dvy * dvy + dvz * dvz);
        Up[i] = (Phi[i] - Phi0) * w;
        const double vesc2 = 2 * fabs(Up[i]) / w;
        if (has_u && uu[i] >= 0.0) Tp[i] += w * (2 * uu[i] / (P.v_fac * P.v_fac));
        const long long j = base + (long long)threadIdx.x * HI + i;
        w_r[j] = T.r[i];
        // the last member of a bin owns the bin's v_esc2 and cumulative momentum
        bool lastofbin = (j == np - 1);
        if (!lastofbin) { int nbn = bin[i]; while (nbn < nbins - 1 && !(T.r[i] < edge[nbn])) nbn++; lastofbin = nbn != bin[i]; }
        if (lastofbin) { vesc_bin[bin[i]] = vesc2; Vc_bin[bin[i]][0] = Pc[i][0]; Vc_bin[bin[i]][1] = Pc[i][1]; Vc_bin[bin[i]][2] = Pc[i][2]; }
        const double Epart = 0.5 * Tp[i] + Up[i];
        if (Epart < best_e) { best_e = Epart; best_j = j; }          // provisional: thread-local, merged below
      }
    }
    // (tile, bin) reductions in a fixed order
    for (int b = blo; b <= bhi; b++) {
      double s[NACC];
#pragma unroll
      for (int q = 0; q < NACC; q++) s[q] = 0.0;
#pragma unroll
      for (int i = 0; i < HI; i++) {
        if (T.act[i] && bin[i] == b) {
          const double w = T.w[i];
          s[0] += w * (c[0] + T.d[i][0]); s[1] += w * (c[1] + T.d[i][1]); s[2] += w * (c[2] + T.d[i][2]);
          s[3] += w * T.d[i][0] * T.d[i][0]; s[4] += w * T.d[i][1] * T.d[i][1]; s[5] += w * T.d[i][2] * T.d[i][2];
          s[6] += w * T.d[i][0] * T.d[i][1]; s[7] += w * T.d[i][0] * T.d[i][2]; s[8] += w * T.d[i][1] * T.d[i][2];
          s[9] += Lm[i][0]; s[10] += Lm[i][1]; s[11] += Lm[i][2];
          s[12] += Tp[i]; s[13] += Up[i];
          if (has_w) { if (fabs(w - 1.0) < ZERO_F) s[14] += w; else if (w > 1.0) s[15] += w; } else s[14] += w;
          s[16] += w; s[17] += 1.0;
        }
      }
      block_sum_n<NACC>(s, smd);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NACC; q++) acc[b][q] += s[q];
      }
    }
    __syncthreads();
  }
  // most bound member: minimum of 0.5 T + U, first index on ties (:4590-4596); merge the thread-local candidates
  {
    double e = best_e; long long jj = best_j < 0 ? 0x7fffffffffffffffll : best_j;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double e2 = __shfl_xor_sync(0xffffffffu, e, o); long long j2 = __shfl_xor_sync(0xffffffffu, jj, o);
      if (e2 < e || (e2 == e && j2 < jj)) { e = e2; jj = j2; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_emin[threadIdx.x >> 5] = e; s_eidx[threadIdx.x >> 5] = jj; }
    __syncthreads();
    best_e = 1e30; best_j = -1;
    for (int q = 0; q < HB / 32; q++)
      if (s_eidx[q] != 0x7fffffffffffffffll && (s_emin[q] < best_e || (s_emin[q] == best_e && s_eidx[q] < best_j) || best_j < 0)) { best_e = s_emin[q]; best_j = s_eidx[q]; }
  }
  __syncthreads();
  // ---- per-bin cumulative values, Jacobi, profile columns
  if (threadIdx.x == 0) prof_finalize(nbins, acc, edge, vesc_bin, Vc_bin, pr, S, R_vir, P, best_j, pos4, mom4, ip, c);
  __syncthreads();
  // ---- R_max / r2 from per-member arrays (find_max, general.c:608-650; smooth3 :548-573)
  // M(<=j): equal masses -> j+1; general -> recomputed by a scan while filling y
  const long long nn = np - NIGNORE;          // arrays are offset by NIGNORE
  double x_r2 = 0.0, x_rmax = 0.0;
  for (int which = 0; which < 2; which++) {   // 0: dens_r2 (3 smoothing passes), 1: Vcirc2 (1 pass)
    if (!has_w) {
      // equal masses: M(<=j) = j + 1 exactly, no scan needed
      for (long long j = threadIdx.x; j < np; j += HB) {
        double r = w_r[j], rpv = j ? w_r[j - 1] : 0.0;
        if (which == 0) { double dV = F43 * ((r * r * r) - (rpv * rpv * rpv)); w_y[j] = 1.0 / dV * (((r + rpv) / 2.) * ((r + rpv) / 2.)); }
        else w_y[j] = (double)(j + 1) / r;
      }
    } else {
    double cM = 0.0;
    for (long long base = 0; base < np; base += HB) {
      long long j = base + threadIdx.x;
      double w = 0.0;
      if (j < np) w = (double)pos4[ip[j]].w;
      double tM, M = cM + block_incl_scan(w, smd, &tM);
      if (j < np) {
        double r = w_r[j], rpv = j ? w_r[j - 1] : 0.0;
        if (which == 0) { double dV = F43 * ((r * r * r) - (rpv * rpv * rpv)); w_y[j] = w / dV * (((r + rpv) / 2.) * ((r + rpv) / 2.)); }
        else w_y[j] = M / r;
      }
      cM += tM;
    }
    }
    __syncthreads();
    double *ya = w_y + NIGNORE, *yb = w_t + NIGNORE;
    const int ns = which == 0 ? 3 : 1;
    if (nn >= 3)
      for (int s = 0; s < ns; s++) {
        for (long long i = threadIdx.x; i < nn; i += HB) {
          double t;
          if (i == 0) t = (ya[0] + ya[1]) / 2.;
          else if (i == nn - 1) t = (ya[nn - 1] + ya[nn - 2]) / 2.;
          else t = (ya[i - 1] + ya[i] + ya[i + 1]) / 3.;
          yb[i] = t;
        }
        __syncthreads();
        double *tp = ya; ya = yb; yb = tp;
      }
    // first index of the maximum over [0, nn-2] with y > -10 (the right-to-left pass of find_max can never improve on it)
    double bv = -10.0; long long bi = nn - 1;
    for (long long i = threadIdx.x; i < nn - 1; i += HB) { double y = ya[i]; if (y > bv) { bv = y; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double v2 = __shfl_xor_sync(0xffffffffu, bv, o); long long i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      bool has2 = v2 > -10.0, has1 = bv > -10.0;
      if ((has2 && !has1) || (has2 && has1 && (v2 > bv || (v2 == bv && i2 < bi)))) { bv = v2; bi = i2; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_emin[threadIdx.x >> 5] = bv; s_eidx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    bv = -10.0; bi = nn - 1;
    for (int q = 0; q < HB / 32; q++) {
      double v2 = s_emin[q]; long long i2 = s_eidx[q];
      if (v2 > -10.0 && (!(bv > -10.0) || v2 > bv || (v2 == bv && i2 < bi))) { bv = v2; bi = i2; }
    }
    double xm = w_r[NIGNORE + bi];
    if (which == 0) x_r2 = xm; else x_rmax = xm;
    __syncthreads();
  }
  // V_max from the first member with r >= R_max (:4892-4901)
  if (!has_w) {
    if (threadIdx.x == 0) {
      long long lo = 0, hi = np - 1;                       // radii ascend: first j with !(r_j < x_rmax), capped at np-1
      while (lo < hi) { long long mid = lo + ((hi - lo) >> 1); if (w_r[mid] < x_rmax) lo = mid + 1; else hi = mid; }
      const double r = w_r[lo], od = (double)(lo + 1) / (F43 * (r * r * r)), M_max = od * F43 * (r * r * r), V_max = M_max / x_rmax;
      S[19] = V_max; S[20] = x_rmax; S[21] = x_r2;
      S[54] = calc_cNFW(V_max, S[10] / R_vir);
    }
    return;
  }
  {
    __shared__ long long sml[HB / 32];
    __shared__ double s_Mmax;
    double cM = 0.0;
    bool done = false;
    for (long long base = 0; base < np && !done; base += HB) {
      long long j = base + threadIdx.x;
      double w = 0.0;
      if (j < np) w = (double)pos4[ip[j]].w;
      double tM, M = cM + block_incl_scan(w, smd, &tM);
      bool hit = (j < np) && (!(w_r[j] < x_rmax) || j == np - 1);
      long long first = block_min_ll(hit ? j : 0x7fffffffffffffffll, sml);
      if (first != 0x7fffffffffffffffll) {
        if (j == first) { double r = w_r[j]; double od = M / (F43 * (r * r * r)); s_Mmax = od * F43 * (r * r * r); }
        done = true;
      }
      cM += tM;
      __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double V_max = s_Mmax / x_rmax;
      S[19] = V_max; S[20] = x_rmax; S[21] = x_r2;
      S[54] = calc_cNFW(V_max, S[10] / R_vir);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// P1, cooperative multi-block form (default): the tiles of all haloes in one grid, like the unbinding pass.
//   k_p_bins    per halo: binning_parameter
//   k_p_sum     per tile: totals of (w, w*mom) and the first / last radial bin the tile touches
//   k_g_scan<4> per halo: carries of M and P
//   k_p_phi     per tile: trapezoid total of Phi (needs M carry)            -> k_g_scan<1>
//   k_p_main    per tile: every per-member quantity; per (tile, bin) sums of the NACC accumulators in a fixed order
//   k_p_finish  per halo: sums the (tile, bin) partials in tile order, then prof_finalize
//   k_p_smooth / k_p_argmax / k_p_vmax: find_max of rho r^2 (3 smoothing passes) and v_circ^2 (1 pass), V_max
// ------------------------------------------------------------------------------------------------
struct PG {
  const int64_t *np;        // final member count per halo
  const int32_t *tile0, *ntile;
  double  *edge;            // [nhalo][MAXBINS]
  double  *vesc_bin;        // [nhalo][MAXBINS]
  double  *Vc_bin;          // [nhalo][MAXBINS][3]
  int32_t *tile_blo, *tile_ns;      // per tile: first bin, number of bins touched
  const int32_t *slot_off;          // per tile: offset into the partial sums
  double  *partial;         // [slots][NACC]
  double  *tbest_e; long long *tbest_j;   // per tile: most bound member
  double  *sp_part; double *sp_be; long long *sp_bj;   // GAS_PARTICLES build: per tile and species NSPC sums, most bound member
  double  *species, *prof_species;        // outputs: [nhalo][64], [total bins][3]
  double  *w_r, *y0a, *y0b, *y1a, *y1b, *Mpre;   // per member (moff0 layout); Mpre only with weights
};

constexpr int PNC = 5;     // scan components of the profile pass: M, P(3), baryon mass (gas + stars; GAS_PARTICLES build)
// species of a particle from `u` (param.h:26-28; ahf_halos.c:4424, :4470): gas u >= PGAS (0), stars u == PSTAR (-4)
__device__ __forceinline__ bool is_baryon(double u) { return u >= 0.0 || fabs(u + 4.0) < ZERO_F; }

__device__ __forceinline__ int bin_of(double rp, const double *__restrict__ edge, int nbins, int b)
{
  while (b < nbins - 1 && !(rp < edge[b])) b++;
  return b;
}

__global__ void k_p_bins(const float4 *__restrict__ pos4, const double *__restrict__ centre, const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members,
                         PG G, const int32_t *__restrict__ act, int nact, const double *__restrict__ scal, HP P)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nact) return;
  const int h = act[a];
  const long long np = G.np[h];
  const uint32_t *ip = members + moff0[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const int nbins = (int)scal[(size_t)h * AHFGPU_NSCAL + 57];
  double *edge = G.edge + (size_t)h * MAXBINS;
  long long k = (long long)floor(((double)P.min_part / 10.) + 0.5);          // binning_parameter (specific.c:259-322)
  double dmin = -1.0;
  while (k < np - 1 && dmin < MACHINE_ZERO) { dmin = dist3(pos4[ip[k]], c); k++; }
  const double dmax = dist3(pos4[ip[np - 1]], c);
  if (dmin < MACHINE_ZERO) dmin = dmax / 2.;
  const double ldmin = log10(dmin), ldmax = log10(dmax), ldr = (ldmax - ldmin) / (double)nbins;
  for (int b = 0; b < nbins; b++) edge[b] = pow(10., ldmin + ((double)b + 1) * ldr);
  edge[nbins - 1] = dmax + ZERO_F;                                           // ahf_halos.c:4226-4241
}

__global__ void __launch_bounds__(HB) k_p_sum(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, const double *__restrict__ centre,
                                              const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members, PG G, const int2 *__restrict__ tiles,
                                              const double *__restrict__ scal, int has_u, double *__restrict__ tt)
{
  __shared__ double smd[(HB / 32) * PNC];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const uint32_t *ip = members + moff0[h];
  double loc[PNC] = { 0, 0, 0, 0, 0 };
#pragma unroll
  for (int i = 0; i < HI; i++) {
    const long long j = base + (long long)threadIdx.x * HI + i;
    if (j < np) {
      const uint32_t pid = ip[j]; const double w = (double)pos4[pid].w; const float4 m = mom4[pid];
      loc[0] += w; loc[1] += w * m.x; loc[2] += w * m.y; loc[3] += w * m.z;
      if (has_u && is_baryon((double)m.w)) loc[4] += w;
    }
  }
  block_sum_n<PNC>(loc, smd);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < PNC; q++) tt[(size_t)blockIdx.x * PNC + q] = loc[q];
    // bins of the tile's first and last member: the first bin whose edge exceeds the radius of the member BEFORE (ahf_halos.c:4283)
    const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
    const int nbins = (int)scal[(size_t)h * AHFGPU_NSCAL + 57];
    const double *edge = G.edge + (size_t)h * MAXBINS;
    const long long last = (base + HT < np ? base + HT : np) - 1;
    const double rp0 = base ? dist3(pos4[ip[base - 1]], c) : -1.0, rp1 = last ? dist3(pos4[ip[last - 1]], c) : -1.0;
    const int blo = bin_of(rp0, edge, nbins, 0), bhi = bin_of(rp1, edge, nbins, blo);
    G.tile_blo[blockIdx.x] = blo; G.tile_ns[blockIdx.x] = bhi - blo + 1;
  }
}

// M(<=j), tile-local scan + carry (weights) or index + 1 (equal masses)
__device__ __forceinline__ void p_tile_M(const TileMembers &T, int has_w, double carryM, long long base, double (&M)[HI], double *smd)
{
  if (!has_w) {
#pragma unroll
    for (int i = 0; i < HI; i++) M[i] = (double)(base + (long long)threadIdx.x * HI + i + 1);
    return;
  }
  double loc[1] = { 0.0 }, ex[1], tot[1];
#pragma unroll
  for (int i = 0; i < HI; i++) loc[0] += T.w[i];
  block_excl_scan_n<1>(loc, ex, tot, smd);
  double run = carryM + ex[0];
#pragma unroll
  for (int i = 0; i < HI; i++) { run += T.w[i]; M[i] = run; }
  __syncthreads();
}
__device__ __forceinline__ void p_prev_member(const float4 *__restrict__ pos4, const uint32_t *__restrict__ ip, int has_w, double carryM, long long base,
                                              const double c[3], double &prev_r, double &prev_I)
{
  prev_r = 0.0; prev_I = 0.0;
  if (base > 0) {
    prev_r = dist3(pos4[ip[base - 1]], c);
    const double Mp = has_w ? carryM : (double)base;          // M(<= base-1) = carry of the tile
    prev_I = prev_r > MACHINE_ZERO ? Mp / (prev_r * prev_r) : 0.0;
  }
}
__global__ void __launch_bounds__(HB) k_p_phi(const float4 *__restrict__ pos4, int has_w, const double *__restrict__ centre, const int64_t *__restrict__ moff0,
                                              const uint32_t *__restrict__ members, PG G, const int2 *__restrict__ tiles, const double *__restrict__ tc4,
                                              double *__restrict__ tt)
{
  __shared__ double smd[HB / 32];
  __shared__ double nb_r[HB], nb_I[HB];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const uint32_t *ip = members + moff0[h];
  TileMembers T;
  load_tile(T, pos4, ip, base, np, c);
  const double carryM = tc4[(size_t)blockIdx.x * PNC];
  double M[HI], term[HI], prev_r, prev_I;
  p_tile_M(T, has_w, carryM, base, M, smd);
  p_prev_member(pos4, ip, has_w, carryM, base, c, prev_r, prev_I);
  double loc = g_tile_phi(T, M, prev_r, prev_I, term, nb_r, nb_I);
  loc = block_sum(loc, smd);
  if (threadIdx.x == 0) tt[blockIdx.x] = loc;
}

template <int MB>      // CTAs per SM the register budget is cut for (3 for thousands of one-tile haloes)
__global__ void __launch_bounds__(HB, MB) k_p_main(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, int has_w, int has_u, const double *__restrict__ centre,
                                               const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members, PG G, const int2 *__restrict__ tiles,
                                               const double *__restrict__ tc4, const double *__restrict__ tcphi, const double *__restrict__ scal, HP P)
{
  __shared__ double smd[(HB / 32) * NACC];
  __shared__ double edge[MAXBINS];
  __shared__ double nb_r[HB], nb_I[HB];
  __shared__ double s_emin[HB / 32]; __shared__ long long s_eidx[HB / 32];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h];
  const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
  const uint32_t *ip = members + moff0[h];
  const double *S = scal + (size_t)h * AHFGPU_NSCAL;
  const int    nbins = (int)S[57];
  const double Phi0 = S[13];
  const double F43 = 4. * PI_ / 3.;
  const double u_fac = (P.x_fac * 100.0) * (P.x_fac * 100.0);      // (box / t_unit)^2 with t_unit = 1/H0 (ahf_halos.c:205, startrun.c:547)
  for (int i = threadIdx.x; i < nbins; i += HB) edge[i] = G.edge[(size_t)h * MAXBINS + i];
  TileMembers T;
  load_tile(T, pos4, ip, base, np, c);
  double mom[HI][3], uu[HI], M[HI], Phi[HI], Pc[HI][3];
#pragma unroll
  for (int i = 0; i < HI; i++) {
    mom[i][0] = mom[i][1] = mom[i][2] = 0.0; uu[i] = -1.0;
    if (T.act[i]) { float4 m = mom4[T.pid[i]]; mom[i][0] = (double)m.x; mom[i][1] = (double)m.y; mom[i][2] = (double)m.z; uu[i] = (double)m.w; }
  }
  __syncthreads();
  const double carryM = tc4[(size_t)blockIdx.x * PNC];
  p_tile_M(T, has_w, carryM, base, M, smd);
  // GAS_PARTICLES build: R_max and r2 come from the dark matter alone (AHFdmonly_Rmax_r2, ahf_halos.c:4603-4611) -> cumulative baryon mass
  double Mb[HI];
#pragma unroll
  for (int i = 0; i < HI; i++) Mb[i] = 0.0;
  if (has_u) {
    double loc[1] = { 0.0 }, ex[1], tot[1];
#pragma unroll
    for (int i = 0; i < HI; i++) if (T.act[i] && is_baryon(uu[i])) loc[0] += T.w[i];
    block_excl_scan_n<1>(loc, ex, tot, smd);
    double run = tc4[(size_t)blockIdx.x * PNC + 4] + ex[0];
#pragma unroll
    for (int i = 0; i < HI; i++) { if (T.act[i] && is_baryon(uu[i])) run += T.w[i]; Mb[i] = run; }
    __syncthreads();
  }
  {
    double loc[3] = { 0, 0, 0 }, ex[3], tot[3];
#pragma unroll
    for (int i = 0; i < HI; i++) { loc[0] += T.w[i] * mom[i][0]; loc[1] += T.w[i] * mom[i][1]; loc[2] += T.w[i] * mom[i][2]; }
    block_excl_scan_n<3>(loc, ex, tot, smd);
    double run[3] = { tc4[(size_t)blockIdx.x * PNC + 1] + ex[0], tc4[(size_t)blockIdx.x * PNC + 2] + ex[1], tc4[(size_t)blockIdx.x * PNC + 3] + ex[2] };
#pragma unroll
    for (int i = 0; i < HI; i++) {
#pragma unroll
      for (int q = 0; q < 3; q++) { run[q] += T.w[i] * mom[i][q]; Pc[i][q] = run[q]; }
    }
    __syncthreads();
  }
  double prev_r, prev_I;
  p_prev_member(pos4, ip, has_w, carryM, base, c, prev_r, prev_I);
  {
    double term[HI], loc[1], ex[1], tot[1];
    loc[0] = g_tile_phi(T, M, prev_r, prev_I, term, nb_r, nb_I);
    block_excl_scan_n<1>(loc, ex, tot, smd);
    double run = tcphi[blockIdx.x] + ex[0];
#pragma unroll
    for (int i = 0; i < HI; i++) { run += term[i]; Phi[i] = run; }
  }
  // radius of the member before each of the thread's members (r_{-1}: -1 for binning, 0 for the per-member arrays)
  const double rr_in = (base == 0) ? -1.0 : prev_r;
  int bin[HI];
  double rprev[HI];
  {
    double rp = threadIdx.x ? nb_r[threadIdx.x - 1] : rr_in;
    int b = 0;
#pragma unroll
    for (int i = 0; i < HI; i++) { b = bin_of(rp, edge, nbins, b); bin[i] = b; rprev[i] = rp; rp = T.r[i]; }
  }
  const int blo = G.tile_blo[blockIdx.x], bhi = blo + G.tile_ns[blockIdx.x] - 1;
  double Tp[HI], Up[HI], Lm[HI][3];
  double best_e = 1e30; long long best_j = 0x7fffffffffffffffll;
#pragma unroll
  for (int i = 0; i < HI; i++) {
    Tp[i] = Up[i] = 0.0; Lm[i][0] = Lm[i][1] = Lm[i][2] = 0.0;
    if (T.act[i]) {
      const double w = T.w[i];
      double dvx = mom[i][0] - Pc[i][0] / M[i], dvy = mom[i][1] - Pc[i][1] / M[i], dvz = mom[i][2] - Pc[i][2] / M[i];   // :4363-4370 mean INCLUDING j
      Lm[i][0] = w * (T.d[i][1] * dvz - T.d[i][2] * dvy);
      Lm[i][1] = w * (T.d[i][2] * dvx - T.d[i][0] * dvz);
      Lm[i][2] = w * (T.d[i][0] * dvy - T.d[i][1] * dvx);
      dvx += P.hubble * T.d[i][0] * P.r_fac / P.v_fac; dvy += P.hubble * T.d[i][1] * P.r_fac / P.v_fac; dvz += P.hubble * T.d[i][2] * P.r_fac / P.v_fac;
      Tp[i] = w * (dvx * dvx + dvy * dvy + dvz * dvz);
      Up[i] = (Phi[i] - Phi0) * w;
      const double vesc2 = 2 * fabs(Up[i]) / w;
      if (has_u && uu[i] >= 0.0) Tp[i] += w * (2 * uu[i] / (P.v_fac * P.v_fac));
      const long long j = base + (long long)threadIdx.x * HI + i;
      const long long e = moff0[h] + j;
      // per-member arrays of find_max (ahf_halos.c:4598-4616): r, rho r^2, v_circ^2
      const double r = T.r[i], rpv = j ? rprev[i] : 0.0, dV = F43 * ((r * r * r) - (rpv * rpv * rpv));
      const double w_dm = (has_u && is_baryon(uu[i])) ? 0.0 : w;
      G.w_r[e] = r; G.y0a[e] = w_dm / dV * (((r + rpv) / 2.) * ((r + rpv) / 2.)); G.y1a[e] = (M[i] - Mb[i]) / r;
      if (has_w) G.Mpre[e] = M[i];
      // the last member of a bin owns the bin's v_esc2 and cumulative momentum
      bool lastofbin = (j == np - 1);
      if (!lastofbin) lastofbin = bin_of(r, edge, nbins, bin[i]) != bin[i];
      if (lastofbin) {
        G.vesc_bin[(size_t)h * MAXBINS + bin[i]] = vesc2;
        double *vc = G.Vc_bin + ((size_t)h * MAXBINS + bin[i]) * 3; vc[0] = Pc[i][0]; vc[1] = Pc[i][1]; vc[2] = Pc[i][2];
      }
      const double Epart = 0.5 * Tp[i] + Up[i];
      if (Epart < best_e) { best_e = Epart; best_j = j; }
    }
  }
  // Per-bin sums of the tile.  The members are radius sorted and a thread owns HI consecutive ones, so the bins of a WARP form one range
  // [wlo, whi] and only its first and last bin can also hold members of other warps: a warp reduces its own bins alone (no block-wide
  // barrier per bin -- a one-tile halo of 500 members has a dozen bins, and 2e4 such haloes spent 3 ms here), writes the bins strictly
  // inside its range straight to the partial table and leaves the two edge bins in shared memory, where they are added up over the warps
  // in warp order.  Per-thread terms, the butterfly inside a warp and the order over the warps are those of the block-wide reduction
  // this replaces (other warps contributed exact zeros), so the sums are the same bit for bit.
  {
    __shared__ double epart[HB / 32][2][NACC];
    __shared__ int    s_wlo[HB / 32], s_whi[HB / 32];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    int mylo = 0x7fffffff, myhi = -1;
#pragma unroll
    for (int i = 0; i < HI; i++) if (T.act[i]) { mylo = min(mylo, bin[i]); myhi = max(myhi, bin[i]); }
    const int wlo = __reduce_min_sync(0xffffffffu, mylo), whi = __reduce_max_sync(0xffffffffu, myhi);     // empty warp: wlo > whi
    if (lane == 0) { s_wlo[wrp] = wlo; s_whi[wrp] = whi; }
    double *pdst = G.partial + (size_t)G.slot_off[blockIdx.x] * NACC;
    for (int b = wlo; b <= whi; b++) {
      double s[NACC];
#pragma unroll
      for (int q = 0; q < NACC; q++) s[q] = 0.0;
#pragma unroll
      for (int i = 0; i < HI; i++) {
        if (T.act[i] && bin[i] == b) {
          const double w = T.w[i];
          s[0] += w * (c[0] + T.d[i][0]); s[1] += w * (c[1] + T.d[i][1]); s[2] += w * (c[2] + T.d[i][2]);
          s[3] += w * T.d[i][0] * T.d[i][0]; s[4] += w * T.d[i][1] * T.d[i][1]; s[5] += w * T.d[i][2] * T.d[i][2];
          s[6] += w * T.d[i][0] * T.d[i][1]; s[7] += w * T.d[i][0] * T.d[i][2]; s[8] += w * T.d[i][1] * T.d[i][2];
          s[9] += Lm[i][0]; s[10] += Lm[i][1]; s[11] += Lm[i][2];
          s[12] += Tp[i]; s[13] += Up[i];
          if (has_w) { if (fabs(w - 1.0) < ZERO_F) s[14] += w; else if (w > 1.0) s[15] += w; } else s[14] += w;
          s[16] += w; s[17] += 1.0;
          if (has_u) {
            if (uu[i] >= 0.0) { s[18] += w; s[20] += w * uu[i] / u_fac; }          // gas (:4424-4474)
            if (fabs(uu[i] + 4.0) < ZERO_F) s[19] += w;                            // stars (:4484)
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NACC; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
      }
      if (lane == 0) {
        double *dst = (b == wlo) ? epart[wrp][0] : (b == whi) ? epart[wrp][1] : pdst + (size_t)(b - blo) * NACC;
#pragma unroll
        for (int q = 0; q < NACC; q++) dst[q] = s[q];
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < (bhi - blo + 1) * NACC; idx += HB) {
      const int b = blo + idx / NACC, q = idx - (idx / NACC) * NACC;
      double t = 0.0;
      bool inside = false;
#pragma unroll
      for (int w = 0; w < HB / 32; w++) {
        const int lo = s_wlo[w], hi = s_whi[w];
        if (lo > hi) continue;
        if (b == lo) t += epart[w][0][q];
        else if (b == hi) t += epart[w][1][q];
        else if (b > lo && b < hi) inside = true;
      }
      if (!inside) pdst[idx] = t;                    // edge bins: the warps in order; bins without a member: zeros
    }
    __syncthreads();
  }
  // GAS_PARTICLES build: gas_only / stars_only sums of the tile (:4420-4530) and the most bound member of each species
  if (has_u) {
    for (int t = 0; t < 2; t++) {
      double s[NSPC];
#pragma unroll
      for (int q = 0; q < NSPC; q++) s[q] = 0.0;
      double se = 1e30; long long sj = 0x7fffffffffffffffll;
#pragma unroll
      for (int i = 0; i < HI; i++) {
        const bool in = T.act[i] && (t == 0 ? (uu[i] >= 0.0) : (fabs(uu[i] + 4.0) < ZERO_F));
        if (in) {
          const double w = T.w[i];
          s[0] += 1.0; s[1] += w;
          s[2] += w * (c[0] + T.d[i][0]); s[3] += w * (c[1] + T.d[i][1]); s[4] += w * (c[2] + T.d[i][2]);
          s[5] += w * mom[i][0]; s[6] += w * mom[i][1]; s[7] += w * mom[i][2];
          s[8] += Lm[i][0]; s[9] += Lm[i][1]; s[10] += Lm[i][2];                 // d x (H d) = 0: the Hubble term of the reference's dV drops out
          s[11] += w * T.d[i][0] * T.d[i][0]; s[12] += w * T.d[i][1] * T.d[i][1]; s[13] += w * T.d[i][2] * T.d[i][2];
          s[14] += w * T.d[i][0] * T.d[i][1]; s[15] += w * T.d[i][0] * T.d[i][2]; s[16] += w * T.d[i][1] * T.d[i][2];
          s[17] += Up[i]; s[18] += Tp[i];
          const double Epart = 0.5 * Tp[i] + Up[i];
          const long long j = base + (long long)threadIdx.x * HI + i;
          if (Epart < se) { se = Epart; sj = j; }
        }
      }
      block_sum_n<NSPC>(s, smd);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double e2 = __shfl_xor_sync(0xffffffffu, se, o); long long j2 = __shfl_xor_sync(0xffffffffu, sj, o);
        if (e2 < se || (e2 == se && j2 < sj)) { se = e2; sj = j2; }
      }
      __syncthreads();
      if ((threadIdx.x & 31) == 0) { s_emin[threadIdx.x >> 5] = se; s_eidx[threadIdx.x >> 5] = sj; }
      __syncthreads();
      if (threadIdx.x == 0) {
        se = s_emin[0]; sj = s_eidx[0];
        for (int q = 1; q < HB / 32; q++) if (s_emin[q] < se || (s_emin[q] == se && s_eidx[q] < sj)) { se = s_emin[q]; sj = s_eidx[q]; }
        double *dst = G.sp_part + ((size_t)blockIdx.x * 2 + t) * NSPC;
#pragma unroll
        for (int q = 0; q < NSPC; q++) dst[q] = s[q];
        G.sp_be[(size_t)blockIdx.x * 2 + t] = se; G.sp_bj[(size_t)blockIdx.x * 2 + t] = sj;
      }
      __syncthreads();
    }
  }
  // most bound member of the tile: minimum of 0.5 T + U, first index on ties (:4590-4596)
  {
    double e = best_e; long long jj = best_j;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double e2 = __shfl_xor_sync(0xffffffffu, e, o); long long j2 = __shfl_xor_sync(0xffffffffu, jj, o);
      if (e2 < e || (e2 == e && j2 < jj)) { e = e2; jj = j2; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_emin[threadIdx.x >> 5] = e; s_eidx[threadIdx.x >> 5] = jj; }
    __syncthreads();
    if (threadIdx.x == 0) {
      e = s_emin[0]; jj = s_eidx[0];
      for (int q = 1; q < HB / 32; q++) if (s_emin[q] < e || (s_emin[q] == e && s_eidx[q] < jj)) { e = s_emin[q]; jj = s_eidx[q]; }
      G.tbest_e[blockIdx.x] = e; G.tbest_j[blockIdx.x] = jj;
    }
  }
}

// NT threads: 256 when a few large haloes have thousands of tiles to sum, 64 when there are thousands of small haloes -- the tail of a halo
// (prof_tail: one thread) and its bins (one thread each) leave a large CTA idle, and 98 registers x 256 threads allow two CTAs per SM
template <int NT>
__global__ void __launch_bounds__(NT) k_p_finish(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, const double *__restrict__ centre,
                                                 const int64_t *__restrict__ moff0, const uint32_t *__restrict__ members, PG G, const int32_t *__restrict__ act,
                                                 double *__restrict__ scal, const int64_t *__restrict__ poff, double *__restrict__ prof, HP P)
{
  __shared__ double acc[MAXBINS][NACC];
  __shared__ double edge[MAXBINS], vesc_bin[MAXBINS], Vc_bin[MAXBINS][3];
  const int h = act[blockIdx.x];
  double *S = scal + (size_t)h * AHFGPU_NSCAL;
  const int nbins = (int)S[57], t0 = G.tile0[h], nt = G.ntile[h];
  // tiles touching bin b form one contiguous range (first and last bin of a tile both grow with the tile number): two binary
  // searches per bin, then the partials are summed in tile order -- a 10^7-particle host has 10^4 tiles
  __shared__ int s_tlo[MAXBINS], s_thi[MAXBINS];
  for (int b = threadIdx.x; b < nbins; b += NT) {
    int lo = t0, hi = t0 + nt;                                        // first tile whose last bin >= b
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (G.tile_blo[mid] + G.tile_ns[mid] - 1 < b) lo = mid + 1; else hi = mid; }
    s_tlo[b] = lo;
    hi = t0 + nt;                                                     // first tile whose first bin > b
    int l2 = lo;
    while (l2 < hi) { const int mid = (l2 + hi) >> 1; if (G.tile_blo[mid] <= b) l2 = mid + 1; else hi = mid; }
    s_thi[b] = l2;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nbins * NACC; idx += NT) {       // (bin, component)
    const int b = idx / NACC, q = idx - b * NACC;
    double a = 0.0;
    for (int t = s_tlo[b]; t < s_thi[b]; t++) a += G.partial[((size_t)G.slot_off[t] + (b - G.tile_blo[t])) * NACC + q];
    acc[b][q] = a;
  }
  for (int i = threadIdx.x; i < nbins; i += NT) {
    edge[i] = G.edge[(size_t)h * MAXBINS + i]; vesc_bin[i] = G.vesc_bin[(size_t)h * MAXBINS + i];
    for (int q = 0; q < 3; q++) Vc_bin[i][q] = G.Vc_bin[((size_t)h * MAXBINS + i) * 3 + q];
  }
  __syncthreads();
  // The bins in parallel (round 2): prof_finalize walked them with ONE thread per halo -- a dozen 3x3 Jacobi decompositions in a row,
  // 0.4 ms per halo; with 2e4 haloes that kernel alone took 6.6 ms.  Same arithmetic per bin (prof_bin), so the numbers are those of
  // the serial form: cumulative sums in place (one thread per component, bins in order), then one thread per bin, then the tail.
  __shared__ double raw17[MAXBINS], raw20[MAXBINS];
  for (int b = threadIdx.x; b < nbins; b += NT) { raw17[b] = acc[b][17]; raw20[b] = acc[b][20]; }
  __syncthreads();
  if (threadIdx.x < NACC) { double run = 0.0; for (int b = 0; b < nbins; b++) { run += acc[b][threadIdx.x]; acc[b][threadIdx.x] = run; } }
  __syncthreads();
  {
    const double F43 = 4. * PI_ / 3.;
    double *pr = prof + poff[h] * AHFGPU_NPROFCOL, *prsp = G.prof_species ? G.prof_species + poff[h] * 3 : nullptr;
    for (int b = threadIdx.x; b < nbins; b += NT) {
      double vesc_run = 0.0;                                            // escape velocity of the last non-empty bin up to b
      for (int q = b; q >= 0; q--) if (raw17[q] > 0.0) { vesc_run = vesc_bin[q]; break; }
      const double M_prev = b ? acc[b - 1][16] : 0.0, V_prev = b ? F43 * (edge[b - 1] * edge[b - 1] * edge[b - 1]) : 0.0;
      prof_bin(b, nbins, acc[b], raw20[b], edge[b], M_prev, V_prev, vesc_run, pr, prsp);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 1e30; long long bj = -1;
    for (int t = t0; t < t0 + nt; t++) {
      const long long jj = G.tbest_j[t];
      if (jj != 0x7fffffffffffffffll && (G.tbest_e[t] < e || bj < 0)) { e = G.tbest_e[t]; bj = jj; }
    }
    const double c[3] = { centre[3 * h], centre[3 * h + 1], centre[3 * h + 2] };
    double Pl[3] = { 0, 0, 0 };
    for (int q = nbins - 1; q >= 0; q--) if (raw17[q] > 0.0) { Pl[0] = Vc_bin[q][0]; Pl[1] = Vc_bin[q][1]; Pl[2] = Vc_bin[q][2]; break; }
    prof_tail(nbins, acc[nbins - 1], Pl, prof + poff[h] * AHFGPU_NPROFCOL, S, S[11], P, bj, pos4, mom4, members + moff0[h], c);
    if (G.species) {                       // gas_only / stars_only (ahf_halos.c:5020-5181): tile sums in tile order
      const double M = S[10], R_vir = S[11];
      for (int t = 0; t < 2; t++) {
        double a[NSPC], se = 1e30; long long sj = -1;
        for (int q = 0; q < NSPC; q++) a[q] = 0.0;
        for (int tl = t0; tl < t0 + nt; tl++) {
          const double *src = G.sp_part + ((size_t)tl * 2 + t) * NSPC;
          for (int q = 0; q < NSPC; q++) a[q] += src[q];
          const long long jj = G.sp_bj[(size_t)tl * 2 + t];
          if (jj != 0x7fffffffffffffffll && (G.sp_be[(size_t)tl * 2 + t] < se || sj < 0)) { se = G.sp_be[(size_t)tl * 2 + t]; sj = jj; }
        }
        double *o = G.species + (size_t)h * 64 + 32 * t;
        for (int q = 0; q < 32; q++) o[q] = 0.0;
        if (a[0] <= 0.0) continue;                                       // reset_SPECIESPROP
        o[0] = a[0]; o[1] = a[1];
        for (int q = 0; q < 3; q++) { o[2 + q] = fmod(a[2 + q] / a[1] + 1.0, 1.0); o[8 + q] = a[5 + q] / a[1]; }
        o[28] = 0.5 * a[18]; o[29] = 0.5 * a[17];
        if (a[0] > 10.0) {                                               // AHF_MINPART_GAS / AHF_MINPART_STARS (param.h:14-15)
          const double aL = sqrt(a[8] * a[8] + a[9] * a[9] + a[10] * a[10]);
          o[13] = a[8] / aL; o[14] = a[9] / aL; o[15] = a[10] / aL;
          double lam = aL / a[1] / sqrt(2. * M * R_vir);
          lam *= P.v_fac * sqrt(P.r_fac / (GRAV_ * P.m_fac));
          o[11] = lam;
          double t1 = sqrt(P.m_fac * M); t1 = t1 * t1 * t1;                  // calc_lambdaE (:3937)
          double t2 = o[28] * P.m_fac * (P.v_fac * P.v_fac), t3 = o[29] * P.m_fac * P.phi_fac;
          t2 = sqrt(fabs(t2 + t3)); t1 = t2 / t1; t2 = P.m_fac * P.r_fac * P.v_fac * aL; t2 = t2 / (P.m_fac * a[1]);
          o[12] = t1 * t2 / GRAV_;
          double it[3][3], ax1, ax2, ax3;
          it[0][0] = a[11]; it[1][1] = a[12]; it[2][2] = a[13]; it[0][1] = it[1][0] = a[14]; it[0][2] = it[2][0] = a[15]; it[1][2] = it[2][1] = a[16];
          get_axes(it, ax1, ax2, ax3);
          o[16] = 1.0; o[17] = (ax1 > 0.) ? sqrt(ax2 / ax1) : 0.0; o[18] = (ax1 > 0.) ? sqrt(ax3 / ax1) : 0.0;
          o[19] = it[0][0]; o[20] = it[1][0]; o[21] = it[2][0]; o[22] = it[0][1]; o[23] = it[1][1]; o[24] = it[2][1];
          o[25] = it[0][2]; o[26] = it[1][2]; o[27] = it[2][2];
        }
        if (sj >= 0) { const float4 pp = pos4[members[moff0[h] + sj]]; o[5] = pp.x; o[6] = pp.y; o[7] = pp.z; }
      }
    }
  }
}

// smooth3 (general.c:548-573) of both per-member arrays, offset by NIGNORE; `both` = 0 smooths only rho r^2
__global__ void __launch_bounds__(HB) k_p_smooth(const int64_t *__restrict__ moff0, PG G, const int2 *__restrict__ tiles, const double *__restrict__ a0,
                                                 double *__restrict__ b0, const double *__restrict__ a1, double *__restrict__ b1, int both)
{
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h], nn = np - NIGNORE, o = moff0[h] + NIGNORE;
#pragma unroll
  for (int k = 0; k < HI; k++) {
    const long long i = base + threadIdx.x + (long long)k * HB - NIGNORE;
    if (i < 0 || i >= nn) continue;
    for (int w = 0; w < (both ? 2 : 1); w++) {
      const double *ya = (w ? a1 : a0) + o; double *yb = (w ? b1 : b0) + o;
      double t;
      if (nn < 3) t = ya[i];
      else if (i == 0) t = (ya[0] + ya[1]) / 2.;
      else if (i == nn - 1) t = (ya[nn - 1] + ya[nn - 2]) / 2.;
      else t = (ya[i - 1] + ya[i] + ya[i + 1]) / 3.;
      yb[i] = t;
    }
  }
}
// first index of the maximum over [0, nn-2] with y > -10, per tile (find_max, general.c:608-650: the right-to-left pass can never improve on it)
__global__ void __launch_bounds__(HB) k_p_argmax(const int64_t *__restrict__ moff0, PG G, const int2 *__restrict__ tiles, const double *__restrict__ y0,
                                                 const double *__restrict__ y1, double *__restrict__ tval, long long *__restrict__ tidx)
{
  __shared__ double s_v[HB / 32]; __shared__ long long s_i[HB / 32];
  const int2 tl = tiles[blockIdx.x];
  const int  h = tl.x; const long long base = (long long)tl.y * HT, np = G.np[h], nn = np - NIGNORE, o = moff0[h] + NIGNORE;
  for (int w = 0; w < 2; w++) {
    const double *y = (w ? y1 : y0) + o;
    double bv = -10.0; long long bi = 0x7fffffffffffffffll;
#pragma unroll
    for (int k = 0; k < HI; k++) {
      const long long i = base + threadIdx.x + (long long)k * HB - NIGNORE;
      if (i >= 0 && i < nn - 1) { const double v = y[i]; if (v > bv) { bv = v; bi = i; } }
    }
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) {
      double v2 = __shfl_xor_sync(0xffffffffu, bv, q); long long i2 = __shfl_xor_sync(0xffffffffu, bi, q);
      if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      bv = s_v[0]; bi = s_i[0];
      for (int q = 1; q < HB / 32; q++) if (s_v[q] > bv || (s_v[q] == bv && s_i[q] < bi)) { bv = s_v[q]; bi = s_i[q]; }
      tval[(size_t)blockIdx.x * 2 + w] = bv; tidx[(size_t)blockIdx.x * 2 + w] = bi;
    }
  }
}
// R_max, r2, V_max (:4875-4901), cNFW
__global__ void k_p_vmax(const int64_t *__restrict__ moff0, PG G, const int32_t *__restrict__ act, int nact, int has_w, const double *__restrict__ tval,
                         const long long *__restrict__ tidx, double *__restrict__ scal)
{
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nact) return;
  const int h = act[a];
  const long long np = G.np[h], nn = np - NIGNORE;
  const int t0 = G.tile0[h], nt = G.ntile[h];
  const double *w_r = G.w_r + moff0[h];
  double *S = scal + (size_t)h * AHFGPU_NSCAL;
  const double F43 = 4. * PI_ / 3.;
  double xm[2];
  for (int w = 0; w < 2; w++) {
    double bv = -10.0; long long bi = nn - 1;
    for (int t = t0; t < t0 + nt; t++) {
      const double v = tval[(size_t)t * 2 + w]; const long long i = tidx[(size_t)t * 2 + w];
      if (i != 0x7fffffffffffffffll && v > bv) { bv = v; bi = i; }
    }
    long long k = NIGNORE + bi; if (k < 0) k = 0;
    xm[w] = w_r[k];
  }
  const double x_r2 = xm[0], x_rmax = xm[1];
  long long lo = 0, hi = np - 1;                       // radii ascend: first j with !(r_j < x_rmax), capped at np-1
  while (lo < hi) { long long mid = lo + ((hi - lo) >> 1); if (w_r[mid] < x_rmax) lo = mid + 1; else hi = mid; }
  const double r = w_r[lo], Mlo = has_w ? G.Mpre[moff0[h] + lo] : (double)(lo + 1);
  const double od = Mlo / (F43 * (r * r * r)), M_max = od * F43 * (r * r * r), V_max = M_max / x_rmax;
  S[19] = V_max; S[20] = x_rmax; S[21] = x_r2;
  S[54] = calc_cNFW(V_max, S[10] / S[11]);
}

// final member lists: one thread per member of the output (its halo by binary search in moff; a CTA per halo waited for the largest halo)
__global__ void k_members_out(const int64_t *__restrict__ moff0, const int64_t *__restrict__ moff, int64_t nhalo, int64_t total, const uint32_t *__restrict__ members,
                              const uint32_t *__restrict__ gid, int64_t *__restrict__ out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t lo = 0, hi = nhalo;                                  // first h with moff[h] > i, minus one
  while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (moff[mid] <= i) lo = mid + 1; else hi = mid; }
  const int64_t h = lo - 1;
  const uint32_t m = members[moff0[h] + (i - moff[h])];
  out[i] = (int64_t)(gid ? gid[m] : m);
}

// Halo-local particle copies.  After the radial sort the members of a halo are in radius order, i.e. scattered over the key-sorted
// particle arrays, and every sweep of the unbinding / profile kernels would gather pos4/mom4 through them (32 random bytes per
// member per sweep, ten to twenty sweeps).  The gathered members are copied ONCE into arrays in member order and the member lists
// are renumbered to positions in those arrays (gid keeps the particle offsets): the kernels are unchanged -- they still index
// pos[members[j]] -- but consecutive members are now consecutive addresses, also after unbound members are removed (the survivors
// stay in increasing order).  Only the unbinding uses the copies (1.1e7-member host: 155 -> 135 ms); the profile pass reads every
// member once and is faster on the shared particle arrays, which overlapping haloes keep in L2 (15.6 vs 20.5 ms), so the lists are
// translated back before it.
__global__ void k_localize_members(const float4 *__restrict__ pos4, const float4 *__restrict__ mom4, uint32_t *__restrict__ members, uint64_t tot,
                                   float4 *__restrict__ hpos, float4 *__restrict__ hmom, uint32_t *__restrict__ gid)
{
  const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j >= tot) return;
  const uint32_t p = members[j];
  hpos[j] = pos4[p]; hmom[j] = mom4[p]; gid[j] = p; members[j] = (uint32_t)j;
}
__global__ void k_globalize_members(uint32_t *__restrict__ members, uint64_t tot, const uint32_t *__restrict__ gid)
{
  const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j < tot) members[j] = gid[members[j]];
}

// exclusive scan of int64 on the host (nhalo-sized arrays; halos are few compared with particles)
static std::vector<int64_t> host_excl(const std::vector<int64_t> &v, int64_t *total)
{
  std::vector<int64_t> o(v.size() + 1, 0);
  for (size_t i = 0; i < v.size(); i++) o[i + 1] = o[i] + v[i];
  *total = o.back();
  return o;
}

template <typename T> static T *dalloc(size_t n)
{
  T *p = nullptr;
  p = static_cast<T *>(ahf::cache_alloc((n ? n : 1) * sizeof(T)));
  return p;
}
static inline unsigned nblk(uint64_t n, int b) { return (unsigned)((n + b - 1) / b); }

// host side of the cooperative U2/U3 pass: tile lists per phase, kernel sequences, the few per-halo read-backs
static void unbind_cooperative(ahfgpu_ctx *c, int64_t nhalo, const HP &P, const double *d_ctr, const int64_t *d_moff0, const std::vector<int64_t> &moff0,
                               const int64_t *d_ng, const std::vector<int64_t> &h_ng, uint32_t *d_members, int64_t tot_g, int64_t *d_np_out,
                               std::vector<int64_t> &h_np, int64_t *iter_members, const std::vector<char> &include)
{
  const bool has_w = c->has_weight;
  const int  has_u = c->has_u ? 1 : 0;
  GH G;
  G.np = dalloc<int64_t>(nhalo); G.nb = dalloc<int64_t>(nhalo); G.nremove = dalloc<int64_t>(nhalo);
  G.Mvir = dalloc<double>(nhalo); G.Rvir = dalloc<double>(nhalo); G.ovd = dalloc<double>(nhalo); G.Phi0 = dalloc<double>(nhalo); G.seed = dalloc<double>(4 * nhalo);
  G.first = dalloc<unsigned long long>(nhalo);
  int64_t *d_n6 = dalloc<int64_t>(nhalo), *d_n7 = dalloc<int64_t>(nhalo);
  CUDA_CHECK(cudaMemcpyAsync(G.np, d_ng, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToDevice, c->stream));
  for (double *q : { G.Mvir, G.Rvir, G.ovd, G.Phi0 }) CUDA_CHECK(cudaMemsetAsync(q, 0, sizeof(double) * nhalo, c->stream));
  CUDA_CHECK(cudaMemsetAsync(G.first, 0xff, sizeof(unsigned long long) * nhalo, c->stream));
  int64_t max_tiles = 0;
  for (int64_t h = 0; h < nhalo; h++) max_tiles += (h_ng[h] + HT - 1) / HT;
  if (max_tiles >= (1ll << 31)) AHF_FAIL("too many member tiles in one call");
  // active list | first tile | tile count | tile list in ONE device block, filled by ONE copy from pinned staging per phase
  // (four copies from pageable vectors before: each a blocking driver-staged transfer)
  const size_t nh2 = (size_t)((nhalo + 1) & ~1ll);                       // keeps the int2 list 8-byte aligned
  const size_t blk_bytes = 3 * nh2 * sizeof(int32_t) + (size_t)max_tiles * sizeof(int2);
  int32_t *d_blk = static_cast<int32_t *>(ahf::cache_alloc(blk_bytes ? blk_bytes : 8));
  int32_t *d_act = d_blk;
  G.tile0 = d_blk + nh2; G.ntile = d_blk + 2 * nh2;
  int2    *d_tiles = reinterpret_cast<int2 *>(d_blk + 3 * nh2);
  if (blk_bytes > c->h_up_bytes) {
    if (c->h_up) cudaFreeHost(c->h_up);
    c->h_up = nullptr; c->h_up_bytes = 0;
    CUDA_CHECK(cudaHostAlloc(&c->h_up, blk_bytes, cudaHostAllocDefault));
    c->h_up_bytes = blk_bytes;
  }
  int32_t *h_blk = static_cast<int32_t *>(c->h_up);
  double  *d_tt = dalloc<double>((size_t)max_tiles * GNC), *d_tc = dalloc<double>((size_t)max_tiles * GNC);
  double  *d_htot1 = dalloc<double>(nhalo), *d_htot5 = dalloc<double>((size_t)nhalo * GNC);
  double  *d_vesc2 = dalloc<double>(tot_g), *d_Mpre = has_w ? dalloc<double>(tot_g) : nullptr;
  uint8_t *d_mask = dalloc<uint8_t>(tot_g);
  uint32_t *d_tmp = dalloc<uint32_t>(tot_g);
  int     *d_changed = dalloc<int>(4);
  h_np = h_ng;
  std::vector<int32_t> act, tile0(nhalo), ntile(nhalo);
  std::vector<int2>    tiles;
  int nact = 0, nt = 0;
  // tile list of the haloes selected by `sel`
  auto build = [&](const std::vector<char> &sel) {
    act.clear(); tiles.clear();
    for (int64_t h = 0; h < nhalo; h++) {
      tile0[h] = 0; ntile[h] = 0;
      if (!sel[h]) continue;
      act.push_back((int32_t)h);
      tile0[h] = (int32_t)tiles.size(); ntile[h] = (int32_t)((h_np[h] + HT - 1) / HT);
      for (int t = 0; t < ntile[h]; t++) tiles.push_back(make_int2((int)h, t));
    }
    nact = (int)act.size(); nt = (int)tiles.size();
    if (!nact) return;
    // every phase ends with a read-back (stream synchronised), so the staging buffer is free again here
    memcpy(h_blk, act.data(), sizeof(int32_t) * nact);
    memcpy(h_blk + nh2, tile0.data(), sizeof(int32_t) * nhalo);
    memcpy(h_blk + 2 * nh2, ntile.data(), sizeof(int32_t) * nhalo);
    memcpy(h_blk + 3 * nh2, tiles.data(), sizeof(int2) * nt);
    CUDA_CHECK(cudaMemcpyAsync(d_blk, h_blk, 3 * nh2 * sizeof(int32_t) + (size_t)nt * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
  };
  auto mass_prefix = [&]() {
    if (!has_w) return;
    LAUNCH(c, k_g_mass_a, (unsigned)nt, HB, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_tiles, d_tt);
    LAUNCH(c, k_g_scan<1>, (unsigned)nact, HB, 0, d_act, G.tile0, G.ntile, d_tt, d_tc, d_htot1);
    LAUNCH(c, k_g_mass_c, (unsigned)nt, HB, 0, c->pos4, d_moff0, d_members, G, d_tiles, d_tc, d_Mpre);
  };
  auto fetch_np = [&]() {
    CUDA_CHECK(cudaMemcpyAsync(h_np.data(), G.np, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  };
  auto rvir = [&]() {
    std::vector<char> sel(nhalo);
    for (int64_t h = 0; h < nhalo; h++) sel[h] = include[h] && h_np[h] >= P.min_part;
    build(sel);
    if (!nact) return;
    mass_prefix();
    LAUNCH(c, k_g_rvir_find, (unsigned)nt, HB, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_tiles, d_Mpre, P);
    LAUNCH(c, k_g_rvir_apply, nblk(nact, 128), 128, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_act, nact, d_Mpre, P);
    fetch_np();
  };
  // ---- rem_outsideRvir, call 0
  rvir();
  CUDA_CHECK(cudaMemcpyAsync(d_n6, G.np, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToDevice, c->stream));
  // ---- rem_unbound (ahf_halos.c:3292-3607): all haloes iterate together, a halo leaves when nremove <= 3 or npart < min_part
  std::vector<char> active(nhalo);
  std::vector<int64_t> h_nrem(nhalo);
  int64_t work = 0, n_iter = 0, n_sweep = 0;
  for (int64_t h = 0; h < nhalo; h++) active[h] = include[h] && h_np[h] >= P.min_part;
  for (int iter = 1;; iter++) {
    build(active);
    if (!nact) break;
    for (int h : act) work += h_np[h];
    n_iter++;
    mass_prefix();
    LAUNCH(c, k_g_phi_a, (unsigned)nt, HB, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_tiles, d_Mpre, d_tt);
    LAUNCH(c, k_g_scan<1>, (unsigned)nact, HB, 0, d_act, G.tile0, G.ntile, d_tt, d_tc, d_htot1);
    LAUNCH(c, k_g_phi0, nblk(nact, 128), 128, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_act, nact, d_Mpre, d_htot1);
    LAUNCH(c, k_g_phi_c, (unsigned)nt, HB, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_tiles, d_Mpre, d_tc, P, d_vesc2);
    LAUNCH(c, k_g_seed, nblk(nact, 64), 64, 0, c->pos4, c->mom4, d_moff0, d_members, G, d_act, nact, iter == 1 ? 1 : 0, P.min_part);
    CUDA_CHECK(cudaMemsetAsync(d_mask, 1, (size_t)tot_g, c->stream));
    LAUNCH(c, k_g_mask<true>, (unsigned)nt, HB, 0, c->pos4, c->mom4, has_u, d_ctr, d_moff0, d_members, G, d_tiles, d_vesc2, d_tc, P, d_mask, d_tt, d_changed);
    for (;;) {              // fixed point: two sweeps per read-back; converged when the last sweep changed nothing
      CUDA_CHECK(cudaMemsetAsync(d_changed, 0, sizeof(int) * 2, c->stream));
      n_sweep += 2;
      for (int q = 0; q < 2; q++) {
        LAUNCH(c, k_g_scan<GNC>, (unsigned)nact, HB, 0, d_act, G.tile0, G.ntile, d_tt, d_tc, d_htot5);
        LAUNCH(c, k_g_mask<false>, (unsigned)nt, HB, 0, c->pos4, c->mom4, has_u, d_ctr, d_moff0, d_members, G, d_tiles, d_vesc2, d_tc, P, d_mask, d_tt, d_changed + q);
      }
      int chg[2];
      read_back(c, chg, d_changed, sizeof(int) * 2);
      if (!chg[1]) break;
    }
    LAUNCH(c, k_g_compact, (unsigned)nt, HB, 0, d_moff0, d_members, G, d_tiles, d_mask, d_tc, d_tmp);
    LAUNCH(c, k_g_copyback, (unsigned)nt, HB, 0, d_moff0, G, d_tiles, d_htot5, d_tmp, d_members);
    LAUNCH(c, k_g_iter_finish, nblk(nact, 128), 128, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_act, nact, d_htot5);
    CUDA_CHECK(cudaMemcpyAsync(h_nrem.data(), G.nremove, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToHost, c->stream));
    fetch_np();
    for (int h : act) active[h] = h_nrem[h] > 3 && h_np[h] >= P.min_part;
  }
  CUDA_CHECK(cudaMemcpyAsync(d_n7, G.np, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToDevice, c->stream));
  // ---- rem_outsideRvir, call 1
  rvir();
  LAUNCH(c, k_g_write_scal, nblk(nhalo, 128), 128, 0, G, d_ctr, d_ng, d_n6, d_n7, nhalo, P.min_part, c->h_scal, d_np_out);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  *iter_members = work;
  c->stage_cnt_extra["halo_unbind_iterations"] = n_iter;
  c->stage_cnt_extra["halo_unbind_mask_sweeps"] = n_sweep;
  for (void *q : { (void *)G.np, (void *)G.nb, (void *)G.nremove, (void *)G.Mvir, (void *)G.Rvir, (void *)G.ovd, (void *)G.Phi0, (void *)G.seed, (void *)G.first,
                   (void *)d_blk, (void *)d_n6, (void *)d_n7, (void *)d_tt, (void *)d_tc, (void *)d_htot1,
                   (void *)d_htot5, (void *)d_vesc2, (void *)d_Mpre, (void *)d_mask, (void *)d_tmp, (void *)d_changed })
    ahf::dfree(q);
}

// host side of the cooperative P1 pass
static void profiles_cooperative(ahfgpu_ctx *c, int64_t nhalo, const HP &P, const double *d_ctr, const int64_t *d_moff0, const uint32_t *d_members,
                                 int64_t tot_g, const int64_t *d_np, const std::vector<int64_t> &h_np)
{
  const int has_w = c->has_weight ? 1 : 0, has_u = c->has_u ? 1 : 0;
  std::vector<int32_t> act, tile0(nhalo, 0), ntile(nhalo, 0);
  std::vector<int2>    tiles;
  for (int64_t h = 0; h < nhalo; h++) {
    if (h_np[h] < P.min_part) continue;
    act.push_back((int32_t)h);
    tile0[h] = (int32_t)tiles.size(); ntile[h] = (int32_t)((h_np[h] + HT - 1) / HT);
    for (int t = 0; t < ntile[h]; t++) tiles.push_back(make_int2((int)h, t));
  }
  const int nact = (int)act.size(), nt = (int)tiles.size();
  if (!nact) return;
  int32_t *d_act = dalloc<int32_t>(nact), *d_tile0 = dalloc<int32_t>(nhalo), *d_ntile = dalloc<int32_t>(nhalo);
  int2    *d_tiles = dalloc<int2>(nt);
  CUDA_CHECK(cudaMemcpyAsync(d_act, act.data(), sizeof(int32_t) * nact, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(d_tile0, tile0.data(), sizeof(int32_t) * nhalo, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(d_ntile, ntile.data(), sizeof(int32_t) * nhalo, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(int2) * nt, cudaMemcpyHostToDevice, c->stream));
  PG G;
  G.np = d_np; G.tile0 = d_tile0; G.ntile = d_ntile;
  G.edge = dalloc<double>((size_t)nhalo * MAXBINS); G.vesc_bin = dalloc<double>((size_t)nhalo * MAXBINS); G.Vc_bin = dalloc<double>((size_t)nhalo * MAXBINS * 3);
  CUDA_CHECK(cudaMemsetAsync(G.vesc_bin, 0, sizeof(double) * nhalo * MAXBINS, c->stream));
  CUDA_CHECK(cudaMemsetAsync(G.Vc_bin, 0, sizeof(double) * nhalo * MAXBINS * 3, c->stream));
  G.tile_blo = dalloc<int32_t>(nt); G.tile_ns = dalloc<int32_t>(nt);
  int32_t *d_slot = dalloc<int32_t>(nt);
  G.slot_off = d_slot;
  G.tbest_e = dalloc<double>(nt); G.tbest_j = dalloc<long long>(nt);
  G.w_r = dalloc<double>(tot_g); G.y0a = dalloc<double>(tot_g); G.y0b = dalloc<double>(tot_g); G.y1a = dalloc<double>(tot_g); G.y1b = dalloc<double>(tot_g);
  G.Mpre = has_w ? dalloc<double>(tot_g) : nullptr;
  G.sp_part = nullptr; G.sp_be = nullptr; G.sp_bj = nullptr; G.species = nullptr; G.prof_species = nullptr;
  if (has_u) {                                                 // GAS_PARTICLES build: per-species blocks
    G.sp_part = dalloc<double>((size_t)nt * 2 * NSPC); G.sp_be = dalloc<double>((size_t)nt * 2); G.sp_bj = dalloc<long long>((size_t)nt * 2);
    c->h_species = dalloc<double>((size_t)nhalo * 64); c->h_prof_species = dalloc<double>((size_t)c->h_total_bins * 3);
    CUDA_CHECK(cudaMemsetAsync(c->h_species, 0, sizeof(double) * 64 * (size_t)nhalo, c->stream));
    G.species = c->h_species; G.prof_species = c->h_prof_species;
  }
  double *d_tt4 = dalloc<double>((size_t)nt * PNC), *d_tc4 = dalloc<double>((size_t)nt * PNC), *d_ht4 = dalloc<double>((size_t)nhalo * PNC);
  double *d_tt1 = dalloc<double>(nt), *d_tc1 = dalloc<double>(nt), *d_ht1 = dalloc<double>(nhalo);
  double *d_tval = dalloc<double>((size_t)nt * 2); long long *d_tidx = dalloc<long long>((size_t)nt * 2);
  LAUNCH(c, k_p_bins, nblk(nact, 64), 64, 0, c->pos4, d_ctr, d_moff0, d_members, G, d_act, nact, c->h_scal, P);
  LAUNCH(c, k_p_sum, (unsigned)nt, HB, 0, c->pos4, c->mom4, d_ctr, d_moff0, d_members, G, d_tiles, c->h_scal, has_u, d_tt4);
  LAUNCH(c, k_g_scan<PNC>, (unsigned)nact, HB, 0, d_act, d_tile0, d_ntile, d_tt4, d_tc4, d_ht4);
  LAUNCH(c, k_p_phi, (unsigned)nt, HB, 0, c->pos4, has_w, d_ctr, d_moff0, d_members, G, d_tiles, d_tc4, d_tt1);
  LAUNCH(c, k_g_scan<1>, (unsigned)nact, HB, 0, d_act, d_tile0, d_ntile, d_tt1, d_tc1, d_ht1);
  // slots of the (tile, bin) partial table: consecutive tiles of a halo share at most their boundary bin, so a halo needs at most
  // tiles + bins - 1 of them -- sized from what the host already knows, the prefix sum stays on the device (no read-back here)
  {
    DevBuf<int> bs;
    exclusive_scan_async<int32_t>(c, G.tile_ns, d_slot, (uint64_t)nt, nullptr, bs);
    bs.release();
  }
  G.partial = dalloc<double>(((size_t)nt + (size_t)c->h_total_bins) * NACC);
  if (nt >= 4096 && nt < 2 * nact && !getenv("AHFGPU_PMAIN_MB2")) LAUNCH(c, k_p_main<3>, (unsigned)nt, HB, 0, c->pos4, c->mom4, has_w, has_u, d_ctr, d_moff0, d_members, G, d_tiles, d_tc4, d_tc1, c->h_scal, P);
  else LAUNCH(c, k_p_main<2>, (unsigned)nt, HB, 0, c->pos4, c->mom4, has_w, has_u, d_ctr, d_moff0, d_members, G, d_tiles, d_tc4, d_tc1, c->h_scal, P);
  if (nact >= 2048 && nt < 4 * nact) LAUNCH(c, k_p_finish<64>, (unsigned)nact, 64, 0, c->pos4, c->mom4, d_ctr, d_moff0, d_members, G, d_act, c->h_scal, c->h_poff, c->h_prof, P);
  else LAUNCH(c, k_p_finish<HB>, (unsigned)nact, HB, 0, c->pos4, c->mom4, d_ctr, d_moff0, d_members, G, d_act, c->h_scal, c->h_poff, c->h_prof, P);
  LAUNCH(c, k_p_smooth, (unsigned)nt, HB, 0, d_moff0, G, d_tiles, G.y0a, G.y0b, G.y1a, G.y1b, 1);
  LAUNCH(c, k_p_smooth, (unsigned)nt, HB, 0, d_moff0, G, d_tiles, G.y0b, G.y0a, G.y1a, G.y1b, 0);
  LAUNCH(c, k_p_smooth, (unsigned)nt, HB, 0, d_moff0, G, d_tiles, G.y0a, G.y0b, G.y1a, G.y1b, 0);
  LAUNCH(c, k_p_argmax, (unsigned)nt, HB, 0, d_moff0, G, d_tiles, G.y0b, G.y1b, d_tval, d_tidx);
  LAUNCH(c, k_p_vmax, nblk(nact, 64), 64, 0, d_moff0, G, d_act, nact, has_w, d_tval, d_tidx, c->h_scal);
  for (void *q : { (void *)d_act, (void *)d_tile0, (void *)d_ntile, (void *)d_tiles, (void *)G.edge, (void *)G.vesc_bin, (void *)G.Vc_bin, (void *)G.tile_blo,
                   (void *)G.tile_ns, (void *)d_slot, (void *)G.tbest_e, (void *)G.tbest_j, (void *)G.w_r, (void *)G.y0a, (void *)G.y0b, (void *)G.y1a, (void *)G.y1b,
                   (void *)G.Mpre, (void *)G.sp_part, (void *)G.sp_be, (void *)G.sp_bj, (void *)d_tt4, (void *)d_tc4, (void *)d_ht4, (void *)d_tt1, (void *)d_tc1, (void *)d_ht1, (void *)d_tval, (void *)d_tidx, (void *)G.partial })
    ahf::dfree(q);
}

void halos_construct(ahfgpu_ctx *c, int64_t nhalo, const double *centre3, const double *gather_rad, const int64_t *seed)
{
  c->free_halos();
  c->nhalo = nhalo;
  const ahfgpu_params &par = c->par;
  HP P; P.r_fac = par.r_fac; P.x_fac = par.x_fac; P.v_fac = par.v_fac; P.m_fac = par.m_fac; P.rho_fac = par.rho_fac; P.phi_fac = par.phi_fac;
  P.hubble = par.hubble; P.ovlim = par.ovlim; P.rho_vir = par.rho_vir; P.vesc_tune = par.vesc_tune; P.min_part = par.min_part;
  if (P.min_part < 2) AHF_FAIL("min_part must be >= 2");
  c->h_scal = dalloc<double>((size_t)nhalo * AHFGPU_NSCAL);
  c->h_moff = dalloc<int64_t>(nhalo + 1); c->h_poff = dalloc<int64_t>(nhalo + 1);
  CUDA_CHECK(cudaMemsetAsync(c->h_scal, 0, sizeof(double) * AHFGPU_NSCAL * (size_t)nhalo, c->stream));
  if (nhalo == 0) {
    CUDA_CHECK(cudaMemsetAsync(c->h_moff, 0, sizeof(int64_t), c->stream)); CUDA_CHECK(cudaMemsetAsync(c->h_poff, 0, sizeof(int64_t), c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return;
  }
  const int64_t n = (int64_t)c->n;
  double  *d_ctr = dalloc<double>(3 * nhalo), *d_rad = dalloc<double>(nhalo);
  int64_t *d_seed = seed ? dalloc<int64_t>(nhalo) : nullptr;
  int64_t *d_rlo = dalloc<int64_t>(27 * nhalo), *d_rhi = dalloc<int64_t>(27 * nhalo), *d_cand = dalloc<int64_t>(nhalo), *d_candoff = dalloc<int64_t>(nhalo + 1);
  int64_t *d_ng = dalloc<int64_t>(nhalo);
  CUDA_CHECK(cudaMemcpyAsync(d_ctr, centre3, sizeof(double) * 3 * nhalo, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(d_rad, gather_rad, sizeof(double) * nhalo, cudaMemcpyHostToDevice, c->stream));
  if (seed) CUDA_CHECK(cudaMemcpyAsync(d_seed, seed, sizeof(int64_t) * nhalo, cudaMemcpyHostToDevice, c->stream));
  std::vector<int64_t> h_cand(nhalo), h_ng(nhalo), h_np(nhalo);
  int64_t tot_cand = 0;
  std::vector<int64_t> candoff;
  double *d_r2 = nullptr; uint32_t *d_idx = nullptr;
  {
    Stage st(c, "halo_gather", 0);
    LAUNCH(c, k_gather_ranges, nblk(nhalo * 32, 128), 128, 0, c->keys, n, d_ctr, d_rad, d_seed, nhalo, d_rlo, d_rhi, d_cand);
    // tile list of the gather (chunks of HT candidates of every (halo, search cell) range): its count is read back together with the
    // candidate counts -- one host synchronisation for both
    const bool  gather_v1 = getenv("AHFGPU_GATHER_V1") != nullptr;
    const int64_t nrange = 27 * nhalo;
    if (nrange >= (1ll << 31)) AHF_FAIL("too many haloes in one call");
    DevBuf<int> ntr, toff, bs, ttot;
    int nt = 0;
    if (!gather_v1) {
      ntr.reserve(nrange); toff.reserve(nrange); ttot.reserve(1);
      CUDA_CHECK(cudaMemsetAsync(ttot.p, 0, sizeof(int), c->stream));
      LAUNCH(c, k_gt_count, nblk(nrange, 256), 256, 0, d_rlo, d_rhi, nrange, ntr.p);
      exclusive_scan_async<int>(c, ntr.p, toff.p, (uint64_t)nrange, ttot.p, bs);
      CUDA_CHECK(cudaMemcpyAsync(&nt, ttot.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_CHECK(cudaMemcpyAsync(h_cand.data(), d_cand, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    candoff = host_excl(h_cand, &tot_cand);
    CUDA_CHECK(cudaMemcpyAsync(d_candoff, candoff.data(), sizeof(int64_t) * (nhalo + 1), cudaMemcpyHostToDevice, c->stream));
    d_r2 = dalloc<double>(tot_cand); d_idx = dalloc<uint32_t>(tot_cand);
    if (gather_v1) {
      LAUNCH(c, k_gather_fill, (unsigned)nhalo, HB, 0, c->pos4, d_ctr, d_rad, d_rlo, d_rhi, d_candoff, d_r2, d_idx, d_ng);
    } else {
      // tile list in the reference's append order
      int4 *d_gt = dalloc<int4>(nt); int32_t *d_act = dalloc<int32_t>(nhalo), *d_t0 = dalloc<int32_t>(nhalo), *d_ntl = dalloc<int32_t>(nhalo);
      double *d_tt = dalloc<double>(nt), *d_tc = dalloc<double>(nt), *d_ht = dalloc<double>(nhalo);
      LAUNCH(c, k_gt_fill, nblk(nrange, 256), 256, 0, d_rlo, d_rhi, nrange, toff.p, ttot.p, d_gt, d_act, d_t0, d_ntl);
      ntr.release(); toff.release(); bs.release(); ttot.release();
      if (nt) LAUNCH(c, k_gather_tiles<false>, (unsigned)nt, HB, 0, c->pos4, d_ctr, d_rad, d_gt, d_candoff, d_tc, d_tt, d_r2, d_idx);
      LAUNCH(c, k_g_scan<1>, (unsigned)nhalo, HB, 0, d_act, d_t0, d_ntl, d_tt, d_tc, d_ht);
      if (nt) LAUNCH(c, k_gather_tiles<true>, (unsigned)nt, HB, 0, c->pos4, d_ctr, d_rad, d_gt, d_candoff, d_tc, d_tt, d_r2, d_idx);
      LAUNCH(c, k_gather_counts, nblk(nhalo, 128), 128, 0, d_ht, nhalo, d_ng);
      ahf::dfree(d_gt); ahf::dfree(d_act); ahf::dfree(d_t0); ahf::dfree(d_ntl); ahf::dfree(d_tt); ahf::dfree(d_tc); ahf::dfree(d_ht);
    }
    CUDA_CHECK(cudaMemcpyAsync(h_ng.data(), d_ng, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  // packed layout of all gathered members (moff0) and of the ones that get sorted (eoff)
  int64_t tot_g = 0, tot_e = 0;
  std::vector<int64_t> moff0 = host_excl(h_ng, &tot_g);
  std::vector<int64_t> h_ne(nhalo);
  for (int64_t h = 0; h < nhalo; h++) h_ne[h] = h_ng[h] >= P.min_part ? h_ng[h] : 0;
  std::vector<int64_t> eoff = host_excl(h_ne, &tot_e);
  if (tot_e >= (1ll << 32)) AHF_FAIL("more than 2^32 gathered members in one call: split the halo list");
  c->stage_cnt_extra["halo_gathered"] = tot_g;
  int64_t *d_moff0 = dalloc<int64_t>(nhalo + 1), *d_eoff = dalloc<int64_t>(nhalo + 1);
  CUDA_CHECK(cudaMemcpyAsync(d_moff0, moff0.data(), sizeof(int64_t) * (nhalo + 1), cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(d_eoff, eoff.data(), sizeof(int64_t) * (nhalo + 1), cudaMemcpyHostToDevice, c->stream));
  uint32_t *d_members = dalloc<uint32_t>(tot_g);
  {
    Stage st(c, "halo_sort", tot_e);
    if (tot_e > 0) {
      uint64_t *k0 = dalloc<uint64_t>(tot_e), *k1 = dalloc<uint64_t>(tot_e);
      uint32_t *v0 = dalloc<uint32_t>(tot_e), *v1 = dalloc<uint32_t>(tot_e), *hid = dalloc<uint32_t>(tot_e);
      uint64_t *ks; uint32_t *vs;
      // mantissa bits kept: neighbouring r^2 of a halo with n members differ by ~1/n relative, so lg(n) + 6 bits leave a tie in about
      // one pair of 64 (measured before the two sorts were merged: too few bits cost 2.7 instead of 0.85 ms at 256^3, seconds on a 1e7 host)
      int64_t nmax = 1;
      for (int64_t h = 0; h < nhalo; h++) nmax = std::max(nmax, h_ng[h]);
      int lg = 0; while ((1ll << lg) < nmax) lg++;
      int hb = 1; while ((1ll << hb) < nhalo) hb++;
      int mbmin = std::min(52, lg + 6);
      if (getenv("AHFGPU_HALO_SORT_SKIP")) mbmin = std::max(0, 52 - atoi(getenv("AHFGPU_HALO_SORT_SKIP")));     // tests: long tie runs
      const int passes = std::min(8, (hb + 6 + mbmin + 7) / 8);
      const int mb = std::min(52, passes * 8 - hb - 6);                  // the key fits 64 bits: hb <= 32, so mb >= 26 in eight passes
      LAUNCH(c, k_sort_setup, nblk(tot_e, 256), 256, 0, d_candoff, d_eoff, nhalo, (uint64_t)tot_e, d_r2, k0, v0, hid, mb);
      radix_sort_pairs(c, k0, v0, k1, v1, (uint64_t)tot_e, hb + 6 + mb, &ks, &vs, 0);
      uint32_t *vs2 = vs;
      LAUNCH(c, k_fix_ties, nblk(tot_e, 256), 256, 0, vs2, ks, hid, d_candoff, d_eoff, d_r2, (uint64_t)tot_e);
      LAUNCH(c, k_apply_perm, nblk(tot_e, 256), 256, 0, vs2, hid, d_candoff, d_eoff, d_moff0, (uint64_t)tot_e, d_idx, d_members);
      ahf::dfree(k0); ahf::dfree(k1); ahf::dfree(v0); ahf::dfree(v1); ahf::dfree(hid);
    }
    LAUNCH(c, k_copy_unsorted, (unsigned)nhalo, 64, 0, d_candoff, d_ng, d_moff0, P.min_part, d_idx, d_members);
  }
  ahf::dfree(d_r2);
  // halo-local copies (k_localize_members); the context's particle pointers are swapped for the rest of the pass
  float4 *hpos = nullptr, *hmom = nullptr; uint32_t *d_gid = nullptr;
  struct Swap { ahfgpu_ctx *c; float4 *p, *m; ~Swap() { c->pos4 = p; c->mom4 = m; } } swap_back{ c, c->pos4, c->mom4 };
  if (tot_g > 0 && !getenv("AHFGPU_HALO_GLOBAL")) {
    Stage st(c, "halo_localize", tot_g);
    hpos = dalloc<float4>(tot_g); hmom = dalloc<float4>(tot_g); d_gid = dalloc<uint32_t>(tot_g);
    LAUNCH(c, k_localize_members, nblk(tot_g, 256), 256, 0, c->pos4, c->mom4, d_members, (uint64_t)tot_g, hpos, hmom, d_gid);
    c->pos4 = hpos; c->mom4 = hmom;
  }
  int64_t *d_np = dalloc<int64_t>(nhalo), *d_work = dalloc<int64_t>(nhalo);
  // U2/U3.  Hybrid: a halo up to `small_max` (16384) gathered members is unbound by ONE CTA that runs all its iterations inside one launch
  // (k_halo_unbind); larger haloes take the cooperative multi-block pass, whose every iteration is a dozen launches and two read-backs
  // -- right for 10^5..10^7 members, but with hundreds of small haloes the slowest of them sets the iteration count for all (256^3
  // box with the 470 seeds of the device tree: 2.4 ms, of which 1.5 ms were launches over nearly empty lists).  The member lists of
  // the two forms are identical, the scalars agree to rounding (fixed but different summation trees); which form serves a halo depends
  // on its gathered count alone.  AHFGPU_UNBIND_V1: everything by one CTA per halo; AHFGPU_UNBIND_SMALL=<n>: the threshold (0: none).
  {
    Stage st(c, "halo_unbind", tot_g);
    const bool all_v1 = getenv("AHFGPU_UNBIND_V1") != nullptr;
    int64_t small_max = 16384;          // measured on the 256^3 box with the 470 device-tree seeds: 0 -> 2.48 ms, 4096 -> 2.47, 16384 -> 1.36, 65536 -> 1.76, all -> 3.78 (scripts/unbind_diag.py)
    if (getenv("AHFGPU_UNBIND_SMALL")) small_max = atoll(getenv("AHFGPU_UNBIND_SMALL"));
    std::vector<char>    include(nhalo, 0);
    std::vector<int32_t> small;
    for (int64_t h = 0; h < nhalo; h++) {
      if (all_v1 || h_ng[h] <= small_max) small.push_back((int32_t)h); else include[h] = 1;
    }
    int64_t tw = 0;
    if (small.size() < (size_t)nhalo) unbind_cooperative(c, nhalo, P, d_ctr, d_moff0, moff0, d_ng, h_ng, d_members, tot_g, d_np, h_np, &tw, include);
    else { c->stage_cnt_extra["halo_unbind_iterations"] = 0; c->stage_cnt_extra["halo_unbind_mask_sweeps"] = 0; }
    if (!small.empty()) {
      int32_t *d_sel = dalloc<int32_t>(small.size());
      CUDA_CHECK(cudaMemcpyAsync(d_sel, small.data(), sizeof(int32_t) * small.size(), cudaMemcpyHostToDevice, c->stream));
      CUDA_CHECK(cudaMemsetAsync(d_work, 0, sizeof(int64_t) * nhalo, c->stream));
      if (small.size() > 1024) LAUNCH(c, k_halo_unbind<2>, (unsigned)small.size(), HB, 0, c->pos4, c->mom4, c->has_u ? 1 : 0, d_ctr, d_moff0, d_ng, d_members, P, c->h_scal, d_np, d_work, d_sel);
      else LAUNCH(c, k_halo_unbind<1>, (unsigned)small.size(), HB, 0, c->pos4, c->mom4, c->has_u ? 1 : 0, d_ctr, d_moff0, d_ng, d_members, P, c->h_scal, d_np, d_work, d_sel);
      std::vector<int64_t> h_work(nhalo);
      CUDA_CHECK(cudaMemcpyAsync(h_np.data(), d_np, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaMemcpyAsync(h_work.data(), d_work, sizeof(int64_t) * nhalo, cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));          // `small` was read by the copy above
      for (int32_t h : small) tw += h_work[h];
      ahf::dfree(d_sel);
    }
    c->stage_cnt_extra["halo_unbind_iter_members"] = tw;
    c->stage_cnt_extra["halo_unbind_small"] = (int64_t)small.size();
  }
  if (d_gid) {                                          // back to particle offsets and the shared particle arrays
    LAUNCH(c, k_globalize_members, nblk(tot_g, 256), 256, 0, d_members, (uint64_t)tot_g, d_gid);
    c->pos4 = swap_back.p; c->mom4 = swap_back.m;
    ahf::dfree(hpos); ahf::dfree(hmom); ahf::dfree(d_gid); hpos = hmom = nullptr; d_gid = nullptr;
  }
  // offsets of the final member lists, profile bins and scratch
  std::vector<int64_t> h_nb(nhalo), h_sc(nhalo);
  for (int64_t h = 0; h < nhalo; h++) {
    int nb = 0;
    if (h_np[h] >= P.min_part) { nb = (int)(6.2 * (log10((double)h_np[h])) - 3.5); if (nb < 2) nb = 2; if (nb > MAXBINS) nb = MAXBINS; }
    h_nb[h] = nb; h_sc[h] = h_np[h] >= P.min_part ? h_np[h] : 0;
  }
  int64_t tot_m = 0, tot_b = 0, tot_s = 0;
  std::vector<int64_t> moff = host_excl(h_np, &tot_m), poff = host_excl(h_nb, &tot_b), soff = host_excl(h_sc, &tot_s);
  c->h_total_members = tot_m; c->h_total_bins = tot_b;
  c->stage_cnt_extra["halo_final_members"] = tot_s;
  CUDA_CHECK(cudaMemcpyAsync(c->h_moff, moff.data(), sizeof(int64_t) * (nhalo + 1), cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(c->h_poff, poff.data(), sizeof(int64_t) * (nhalo + 1), cudaMemcpyHostToDevice, c->stream));
  int64_t *d_soff = dalloc<int64_t>(nhalo + 1);
  CUDA_CHECK(cudaMemcpyAsync(d_soff, soff.data(), sizeof(int64_t) * (nhalo + 1), cudaMemcpyHostToDevice, c->stream));
  c->h_prof = dalloc<double>((size_t)tot_b * AHFGPU_NPROFCOL);
  c->h_members = dalloc<int64_t>(tot_m);
  const uint32_t *out_ids = d_gid ? d_gid : (c->slab ? c->order : (uint32_t *)nullptr);
  // the member lists are final: with a registered pinned buffer (ahfgpu_halo_members_buffer) they leave for the host now, on the copy
  // stream, while the profiles are computed
  bool members_written = false;
  if (c->early_members && tot_m > 0 && tot_m <= c->early_cap) {
    if (!c->copy_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      for (auto &e : c->ev_copy) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_mom, cudaEventDisableTiming));
    }
    if (!c->ev_members) CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_members, cudaEventDisableTiming));
    LAUNCH(c, k_members_out, nblk(tot_m, 256), 256, 0, d_moff0, c->h_moff, nhalo, (int64_t)tot_m, d_members, out_ids, c->h_members);
    CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
    CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
    CUDA_CHECK(cudaMemcpyAsync(c->early_members, c->h_members, sizeof(int64_t) * tot_m, cudaMemcpyDeviceToHost, c->copy_stream));
    CUDA_CHECK(cudaEventRecord(c->ev_members, c->copy_stream));
    c->early_sent = true; members_written = true;
  }
  const bool prof_v1 = getenv("AHFGPU_PROFILES_V1") != nullptr;      // previous form: one CTA per halo (kept for A/B timing)
  double *d_scratch = prof_v1 ? dalloc<double>((size_t)tot_s * 3) : nullptr;
  {
    Stage st(c, "halo_profiles", tot_s);
    if (prof_v1)
      LAUNCH(c, k_halo_profiles, (unsigned)nhalo, HB, 0, c->pos4, c->mom4, c->has_weight ? 1 : 0, c->has_u ? 1 : 0, d_ctr, d_moff0, d_members, d_np, P,
             c->h_scal, c->h_poff, c->h_prof, d_soff, d_scratch);
    else
      profiles_cooperative(c, nhalo, P, d_ctr, d_moff0, d_members, tot_g, d_np, h_np);
    // a slab of a distributed box reports GLOBAL INPUT INDICES (the resident `order` array), not offsets into its own sorted set
    if (tot_m > 0 && !members_written) LAUNCH(c, k_members_out, nblk(tot_m, 256), 256, 0, d_moff0, c->h_moff, nhalo, (int64_t)tot_m, d_members, out_ids, c->h_members);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  ahf::dfree(d_ctr); ahf::dfree(d_rad); ahf::dfree(d_seed); ahf::dfree(d_rlo); ahf::dfree(d_rhi); ahf::dfree(d_cand); ahf::dfree(d_candoff); ahf::dfree(d_ng);
  ahf::dfree(d_idx); ahf::dfree(d_moff0); ahf::dfree(d_eoff); ahf::dfree(d_members); ahf::dfree(d_np); ahf::dfree(d_work); ahf::dfree(d_soff); ahf::dfree(d_scratch);
  ahf::dfree(hpos); ahf::dfree(hmom); ahf::dfree(d_gid);
}

}  // namespace ahf

// hilbert.cuh -- 3-D Hilbert index, bit-compatible with the reference's libsfc
// (src/libsfc/hilbert.c:139-243 after Butz/Moore; x is dimension 0; src/libsfc/hilbert_util.c:69-92 for
// the scaling/clamp of real coordinates).  Host+device so that the C-ABI host code can use it as well.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define AHF_HD __host__ __device__ __forceinline__
#else
#define AHF_HD inline
#endif

namespace ahf {

// spread the low 21 bits of v to every third bit
AHF_HD uint64_t spread3(uint64_t v)
{
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}
AHF_HD uint32_t compact3(uint64_t v)
{
  v &= 0x1249249249249249ull;
  v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ull;
  v = (v ^ (v >> 4)) & 0x100f00f00f00f00full;
  v = (v ^ (v >> 8)) & 0x1f0000ff0000ffull;
  v = (v ^ (v >> 16)) & 0x1f00000000ffffull;
  v = (v ^ (v >> 32)) & 0x1fffffull;
  return (uint32_t)v;
}

AHF_HD unsigned hil_next_rot(unsigned rot, unsigned g)
{
  unsigned low = g & (0u - g) & 3u;
  rot += (low == 1u) ? 1u : (low == 2u) ? 2u : 0u;
  rot += 1u;
  return rot >= 3u ? rot - 3u : rot;
}

// index of integer cell (cx,cy,cz) with `bits` bits per dimension (1..21)
AHF_HD uint64_t hilbert_index(uint32_t cx, uint32_t cy, uint32_t cz, unsigned bits)
{
  uint64_t inter = spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2);
  uint64_t index = 0;
  if (bits > 1) {
    inter ^= inter >> 3;
    unsigned rot = 0, flip = 0;
    for (int b = 3 * (int)bits - 3; b >= 0; b -= 3) {
      unsigned g = (unsigned)(inter >> b) & 7u;
      g ^= flip;
      g = ((g >> rot) | (g << (3 - rot))) & 7u;
      index = (index << 3) | g;
      flip = 1u << rot;
      rot = hil_next_rot(rot, g);
    }
    // bit 3j-1 for j = 1..bits-1 (hilbert.c: index ^= nthbits >> 1)
    const uint64_t mask = 0x4924924924924924ull & ((1ull << (3 * bits - 3)) - 1ull);
    index ^= mask;
  } else {
    index = inter;
  }
  for (unsigned j = 1; j < 3 * bits; j <<= 1) index ^= index >> j;
  return index;
}

// ---- table-driven form for bits == 21 (the particle keys): three levels per look-up.
// The loop above is a finite automaton: state = (rot, flip) with flip in {0, 1<<prev_rot}: 12 states; input = the 3 transformed
// bits of a level.  hil_tab3[state * 512 + 9 input bits] = 9 index bits | next state << 9 (built once on the host with the very
// step code above, hilbert_build_tab3).  Seven look-ups replace 21 dependent steps.
AHF_HD unsigned hil_state(unsigned rot, unsigned flip) { return rot * 4u + (flip == 4u ? 3u : flip); }
inline void hilbert_build_tab3(uint16_t *tab /* [12 * 512] */)
{
  for (unsigned s = 0; s < 12; s++)
    for (unsigned in = 0; in < 512; in++) {
      unsigned rot = s >> 2, fc = s & 3u, flip = fc == 3u ? 4u : fc, out = 0;
      for (int lv = 2; lv >= 0; lv--) {
        unsigned g = (in >> (3 * lv)) & 7u;
        g ^= flip;
        g = ((g >> rot) | (g << (3 - rot))) & 7u;
        out = (out << 3) | g;
        flip = 1u << rot;
        rot = hil_next_rot(rot, g);
      }
      tab[s * 512 + in] = (uint16_t)(out | (hil_state(rot, flip) << 9));
    }
}
AHF_HD uint64_t hilbert_index21_tab(uint32_t cx, uint32_t cy, uint32_t cz, const uint16_t *tab)
{
  uint64_t inter = spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2);
  inter ^= inter >> 3;
  uint64_t index = 0;
  unsigned st = 0;
#pragma unroll
  for (int b = 54; b >= 0; b -= 9) {
    const unsigned e = tab[st * 512u + ((unsigned)(inter >> b) & 511u)];
    index = (index << 9) | (e & 511u);
    st = e >> 9;
  }
  index ^= 0x4924924924924924ull & ((1ull << 60) - 1ull);
#pragma unroll
  for (unsigned j = 1; j < 63; j <<= 1) index ^= index >> j;
  return index;
}

// inverse: integer cell of an index
AHF_HD void hilbert_coords(uint64_t index, unsigned bits, uint32_t &cx, uint32_t &cy, uint32_t &cz)
{
  uint64_t coords = 0;
  if (bits > 1) {
    uint64_t nth = 0x1249249249249249ull & ((bits >= 21) ? 0x7fffffffffffffffull : ((1ull << (3 * bits)) - 1ull));
    index ^= (index ^ nth) >> 1;
    unsigned rot = 0, flip = 0;
    for (int b = 3 * (int)bits - 3; b >= 0; b -= 3) {
      unsigned g = (unsigned)(index >> b) & 7u;
      unsigned r = ((g << rot) | (g >> (3 - rot))) & 7u;
      coords = (coords << 3) | (r ^ flip);
      flip = 1u << rot;
      rot = hil_next_rot(rot, g);
    }
    for (unsigned j = 3; j < 3 * bits; j <<= 1) coords ^= coords >> j;
  } else {
    coords = index ^ (index >> 1);
  }
  cx = compact3(coords); cy = compact3(coords >> 1); cz = compact3(coords >> 2);
}

// hilbert_util_calcHikey: scale to 2^bits cells in double, truncate, clamp the one overflow value
AHF_HD uint64_t hilbert_key_posd(double x, double y, double z, unsigned bits)
{
  const double   mx = (double)(1u << bits);
  const uint32_t top = (1u << bits);
  uint32_t c0 = (uint32_t)(int64_t)(x * mx), c1 = (uint32_t)(int64_t)(y * mx), c2 = (uint32_t)(int64_t)(z * mx);
  if (c0 == top) c0 = top - 1;
  if (c1 == top) c1 = top - 1;
  if (c2 == top) c2 = top - 1;
  return hilbert_index(c0, c1, c2, bits);
}
// float positions: x * 2^bits is exact in float (power-of-two scaling), so no double arithmetic is needed
AHF_HD uint64_t hilbert_key_pos(float x, float y, float z, unsigned bits)
{
  const float    mx = (float)(1u << bits);
  const uint32_t top = (1u << bits);
  uint32_t c0 = (uint32_t)(int32_t)(x * mx), c1 = (uint32_t)(int32_t)(y * mx), c2 = (uint32_t)(int32_t)(z * mx);
  if (c0 == top) c0 = top - 1;
  if (c1 == top) c1 = top - 1;
  if (c2 == top) c2 = top - 1;
  return hilbert_index(c0, c1, c2, bits);
}
AHF_HD uint64_t hilbert_key_pos21_tab(float x, float y, float z, const uint16_t *tab)
{
  const float    mx = 2097152.0f;
  const uint32_t top = 1u << 21;
  uint32_t c0 = (uint32_t)(int32_t)(x * mx), c1 = (uint32_t)(int32_t)(y * mx), c2 = (uint32_t)(int32_t)(z * mx);
  if (c0 == top) c0 = top - 1;
  if (c1 == top) c1 = top - 1;
  if (c2 == top) c2 = top - 1;
  return hilbert_index21_tab(c0, c1, c2, tab);
}

}  // namespace ahf

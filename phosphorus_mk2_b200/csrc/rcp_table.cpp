// rcp_table.cpp — the host's RCPSS / RCPPS instruction as a table, for the device's reference-compatible normalisation.
//
// The reference normalises vectors as x * _mm256_rcp_ps(sqrt(dot)) (src/math/simd/vector.hpp:94-96,126-133): a ~12-bit
// reciprocal whose exact values are the CPU's business (Intel and AMD differ).  Its visible effect is systematic — shadow
// rays come out up to 3.7e-4 too long, so a share of them hit the light's own triangle and count as occluded
// (spt.hpp:116-148): a cpu_t tile is 13-40 % darker than exact arithmetic renders it.  A cuda_t device that shares a frame
// with cpu_t (plugins/blender/session.cpp:85-99) therefore has to normalise the way THIS host does.  RCPSS of a normal
// float is a function of the leading mantissa bits times an exact power of two; this file finds how many bits matter on
// the host it runs on (11 on Intel), samples the 2^bits values and verifies the model on a sweep of all mantissas.
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#define PHOS_HAVE_RCPSS 1
#endif

namespace phos {

#ifdef PHOS_HAVE_RCPSS
static float host_rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
static float from_bits(uint32_t b) {
  float f;
  memcpy(&f, &b, 4);
  return f;
}
#endif

// table[m] = rcpss(1.m) for the leading `bits` mantissa bits m; returns false when the host has no RCPSS or it does not
// follow the model (then the reference-compatible mode is refused rather than approximated)
bool sample_host_rcp(std::vector<float>& table, int& bits) {
#ifdef PHOS_HAVE_RCPSS
  for (bits = 11; bits <= 16; ++bits) {
    const uint32_t n = 1u << bits, shift = 23u - (uint32_t)bits;
    table.resize(n);
    for (uint32_t m = 0; m < n; ++m) table[m] = host_rcp(from_bits(0x3f800000u | (m << shift)));
    bool ok = true;
    for (uint32_t m = 0; ok && m < (1u << 23); m += 61u) {  // every 61st mantissa, three exponents
      for (int e : {-20, 0, 17}) {
        const float x = from_bits((uint32_t)(127 + e) << 23 | m);
        uint32_t want, tb;
        const float r = host_rcp(x);
        memcpy(&want, &r, 4);
        memcpy(&tb, &table[m >> shift], 4);
        if (tb - ((uint32_t)e << 23) != want) {
          ok = false;
          break;
        }
      }
    }
    if (ok) return true;
  }
#endif
  table.clear();
  bits = 0;
  return false;
}

}  // namespace phos

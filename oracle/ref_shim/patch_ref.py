#!/usr/bin/env python3
"""Make a throw-away copy of the reference sources compile with g++ 13.

TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Operates on a COPY of /root/reference/src that
oracle/Makefile places in a temp directory; nothing here is copied into the repository.

The reference was written for 2019-era clang, which bit-casts implicitly between __m256 and
__m256i ("lax vector conversions") and accepts two GNU/MS extensions g++ rejects.  Every edit
below only spells out what that compiler did implicitly — no arithmetic changes:
  * math/simd/int8.hpp, math/simd/float8.hpp: explicit _mm256_cast{si256_ps,ps_si256} bit-casts;
  * accel/bvh/binned_sah_builder.hpp: GNU range designator `[0 ... 7] = {geometry}` -> 8 explicit
    geometry_t(geometry) copies (same copy constructor);
  * bsdf.hpp: in-class `template<>` explicit specialisation -> plain overload (same overload wins).
Each replacement asserts that the expected text is present exactly `count` times, so a changed
reference fails loudly instead of building something else.
"""
import sys, pathlib

root = pathlib.Path(sys.argv[1])

def patch(rel, old, new, count=1):
    p = root / rel
    s = p.read_text()
    n = s.count(old)
    assert n == count, f"{rel}: expected {count} x {old!r}, found {n}"
    p.write_text(s.replace(old, new))

I8 = "math/simd/int8.hpp"
patch(I8, '#include "float8.hpp"\n',
      '#include "float8.hpp"\n'
      'namespace simd {\n'
      '  inline __m256 _f(const __m256i& x) { return _mm256_castsi256_ps(x); }\n'
      '  inline __m256 _f(const __m256& x) { return x; }\n'
      '}\n')
patch(I8, "andnot(l, r));", "andnot(_f(l), _f(r)));")
patch(I8, "_mm256_castps_si256(sub(l, r));", "_mm256_castps_si256(sub(_f(l), _f(r)));")
for op in ("_and", "_or", "_xor"):
    patch(I8, f"{op}(v, r.v)", f"{op}(_f(v), _f(r.v))", 2)
for op in ("mul", "div", "eq", "lte", "gte"):
    patch(I8, f"int32_t({op}(v, r.v))", f"int32_t(_mm256_castps_si256({op}(_f(v), _f(r.v))))")
patch(I8, "select(m.v, l.v, r.v)", "select(_f(m.v), _f(l.v), _f(r.v))")
patch("math/simd/float8.hpp", "_mm256_mask_i32gather_ps(s, p, i, m, 4)",
      "_mm256_mask_i32gather_ps(s, p, i, _mm256_castsi256_ps(m), 4)")
patch("accel/bvh/binned_sah_builder.hpp", "{ [0 ... 7] = { geometry } }",
      "{ geometry_t(geometry), geometry_t(geometry), geometry_t(geometry), geometry_t(geometry),"
      " geometry_t(geometry), geometry_t(geometry), geometry_t(geometry), geometry_t(geometry) }")
patch("bsdf.hpp", "  template<>\n  inline void add_lobe(", "  inline void add_lobe(")
# --- additive accessor for the GPU device glue (integration/cuda.cpp): the per-face
# smooth flag lives in the private mesh_t::details_t (src/mesh.cpp:10-18)
patch("mesh.hpp", "  inline bool has_per_vertex_normals() const {",
      "  /* is this face shaded with interpolated vertex normals? (added for the GPU device) */\n"
      "  bool is_smooth(uint32_t face) const;\n\n"
      "  inline bool has_per_vertex_normals() const {")
p = root / "mesh.cpp"
p.write_text(p.read_text() + "\nbool mesh_t::is_smooth(uint32_t face) const {\n  return details->smooth[face];\n}\n")
print("patched", root)

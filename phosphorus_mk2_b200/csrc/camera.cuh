// camera.cuh — primary-ray generation shared by the parity entry (render.cu) and the wavefront (wavefront.cu).
//
// camera::perspective_kernel_t (reference src/kernels/cpu/camera.hpp:113-152), one pixel:
//   ndcx = (px - 0.5) / W - 0.5, ndcy = 0.5 - (py - 0.5) / H,
//   d = ((ndcx + jx / W) * (W / H) * zoom, (ndcy + jy / H) * zoom, -1), normalised, zoom = 1.12 tan(fov / 2);
//   thin lens when aperture_radius != 0 (src/entities/camera.hpp:37-39): lens point from
//   simd::concentric_sample_disc (src/math/simd/sampling.hpp:7-32) scaled by the radius, focus distance
//   ft = |focal_distance / d.z|, d = d * ft - lens, normalised; then point / vector through to_world
//   (row-vector convention, mul, fmadd, fmadd, + row 3: src/math/simd/matrix.hpp:58-104).
// The lens mapping follows the reference to the letter — it uses the raw [0,1) samples (the [-1,1) offset it
// computes is never read), its constants named pi_o_4 / pi_o_2 hold 4/pi and 2/pi, and simd::select(m, l, r)
// is m ? r : l (float8.hpp:103-105) — the CPU restatement the tests compare against is pinned bit for
// bit to the compiled reference kernel (tests/).  Differences to the reference, as for
// every normalisation on the device: IEEE 1 / sqrt instead of the 12-bit RCPPS (unless the reference-
// compatible mode samples the host's RCPPS into a table, below); CUDA sinf / cosf.
// Every operation is an explicit round-to-nearest intrinsic: the result does not depend on -fmad.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace phos {

struct DevCamera {
  // reference-compatible normalisation (phos_cuda_reference_normalize): the host's RCPSS as a table over the leading
  // mantissa bits (rcp_table.cpp); null = exact 1 / sqrt
  const float* rcp_tab;
  uint32_t rcp_shift;  // 23 - bits
  float m[16];  // to_world, row-vector convention
  float zoom;   // 1.12 * tan(fov / 2)
  float stepx, stepy, ratio;
  float focal_distance, aperture_radius;
  uint32_t width, height;
};

// _mm256_rcp_ps(x) of the host for a positive normal x: the table entry of the leading mantissa bits, exponent negated
__device__ __forceinline__ float reference_rcp(const float* __restrict__ tab, uint32_t shift, float x) {
  const uint32_t b = __float_as_uint(x);
  const uint32_t e = (b >> 23) & 0xffu;
  if (e == 0u || e == 255u || (b >> 31)) return __fdiv_rn(1.0f, x);  // zero / denormal / inf / nan / negative: not a length
  return __uint_as_float(__float_as_uint(__ldg(tab + ((b & 0x7fffffu) >> shift))) - ((e - 127u) << 23));
}

// 1 / length the way vector3_t<8>::normalize does it (src/math/simd/vector.hpp:126-133): rcp(sqrt(l)) — exact division
// unless the reference-compatible mode is on
__device__ __forceinline__ float inv_length(const DevCamera& cam, float l) {
  const float s = __fsqrt_rn(l);
  return cam.rcp_tab ? reference_rcp(cam.rcp_tab, cam.rcp_shift, s) : __fdiv_rn(1.0f, s);
}

__device__ __forceinline__ void camera_ray(const DevCamera& cam, uint32_t px, uint32_t py, float jx, float jy, float lu,
                                           float lv, float* o, float* w) {
  const float sx = (float)px, sy = (float)py;
  const float ndcy = __fsub_rn(0.5f, __fmul_rn(__fadd_rn(-0.5f, sy), cam.stepy));
  const float ndcx = __fsub_rn(__fmul_rn(__fadd_rn(-0.5f, sx), cam.stepx), 0.5f);
  float dx = __fmul_rn(__fmul_rn(__fadd_rn(ndcx, __fmul_rn(jx, cam.stepx)), cam.ratio), cam.zoom);
  float dy = __fmul_rn(__fadd_rn(ndcy, __fmul_rn(jy, cam.stepy)), cam.zoom);
  float dz = -1.0f;
  float l = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
  float ool = inv_length(cam, l);
  dx = __fmul_rn(dx, ool);
  dy = __fmul_rn(dy, ool);
  dz = __fmul_rn(dz, ool);
  float lx = 0.0f, ly = 0.0f;
  if (cam.aperture_radius != 0.0f) {
    const float pi_o_2 = (float)(2.0f / 3.14159265358979323846), pi_o_4 = (float)(4.0f / 3.14159265358979323846);
    const bool x_gt_y = fabsf(lu) > fabsf(lv);
    const float r = x_gt_y ? lv : lu;
    const float theta1 = __fmul_rn(pi_o_4, __fdiv_rn(lv, lu));
    const float theta2 = __fsub_rn(pi_o_2, __fmul_rn(pi_o_4, __fdiv_rn(lu, lv)));
    const float theta = x_gt_y ? theta2 : theta1;
    lx = __fmul_rn(__fmul_rn(r, cosf(theta)), cam.aperture_radius);
    ly = __fmul_rn(__fmul_rn(r, sinf(theta)), cam.aperture_radius);
    const float ft = fabsf(__fdiv_rn(cam.focal_distance, dz));
    dx = __fsub_rn(__fmul_rn(dx, ft), lx);
    dy = __fsub_rn(__fmul_rn(dy, ft), ly);
    dz = __fsub_rn(__fmul_rn(dz, ft), 0.0f);
    l = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
    ool = inv_length(cam, l);
    dx = __fmul_rn(dx, ool);
    dy = __fmul_rn(dy, ool);
    dz = __fmul_rn(dz, ool);
  }
  const float* m = cam.m;
  o[0] = __fadd_rn(__fmaf_rn(0.0f, m[8], __fmaf_rn(ly, m[4], __fmul_rn(lx, m[0]))), m[12]);
  o[1] = __fadd_rn(__fmaf_rn(0.0f, m[9], __fmaf_rn(ly, m[5], __fmul_rn(lx, m[1]))), m[13]);
  o[2] = __fadd_rn(__fmaf_rn(0.0f, m[10], __fmaf_rn(ly, m[6], __fmul_rn(lx, m[2]))), m[14]);
  w[0] = __fmaf_rn(dz, m[8], __fmaf_rn(dy, m[4], __fmul_rn(dx, m[0])));
  w[1] = __fmaf_rn(dz, m[9], __fmaf_rn(dy, m[5], __fmul_rn(dx, m[1])));
  w[2] = __fmaf_rn(dz, m[10], __fmaf_rn(dy, m[6], __fmul_rn(dx, m[2])));
}

}  // namespace phos

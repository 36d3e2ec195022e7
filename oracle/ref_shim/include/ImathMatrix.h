// Stand-in for Imath::M44f (oracle/_ref build only; see ImathVec.h). Row-vector convention.
#pragma once
#include "ImathVec.h"
namespace Imath {
template <typename T> struct Matrix44 {
  T x[4][4];
  Matrix44() {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) x[i][j] = (i == j) ? T(1) : T(0);
  }
  T* operator[](int i) { return x[i]; }
  const T* operator[](int i) const { return x[i]; }
  template <typename S> void multVecMatrix(const Vec3<S>& src, Vec3<S>& dst) const {
    S a = src.x * x[0][0] + src.y * x[1][0] + src.z * x[2][0] + x[3][0];
    S b = src.x * x[0][1] + src.y * x[1][1] + src.z * x[2][1] + x[3][1];
    S c = src.x * x[0][2] + src.y * x[1][2] + src.z * x[2][2] + x[3][2];
    S w = src.x * x[0][3] + src.y * x[1][3] + src.z * x[2][3] + x[3][3];
    dst = Vec3<S>(a / w, b / w, c / w);
  }
  template <typename S> void multDirMatrix(const Vec3<S>& src, Vec3<S>& dst) const {
    S a = src.x * x[0][0] + src.y * x[1][0] + src.z * x[2][0];
    S b = src.x * x[0][1] + src.y * x[1][1] + src.z * x[2][1];
    S c = src.x * x[0][2] + src.y * x[1][2] + src.z * x[2][2];
    dst = Vec3<S>(a, b, c);
  }
};
typedef Matrix44<float> M44f;
}

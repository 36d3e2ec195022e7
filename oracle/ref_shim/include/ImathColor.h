// Stand-in for Imath::Color3f (oracle/_ref build only; see ImathVec.h).
#pragma once
#include "ImathVec.h"
namespace Imath {
template <typename T> struct Color3 : public Vec3<T> {
  Color3() {}
  explicit Color3(T a) : Vec3<T>(a) {}
  Color3(T a, T b, T c) : Vec3<T>(a, b, c) {}
  Color3(const Vec3<T>& v) : Vec3<T>(v) {}
  template <typename S> Color3(const Vec3<S>& v) : Vec3<T>(v) {}
};
typedef Color3<float> Color3f;
typedef Color3<float> C3f;
}

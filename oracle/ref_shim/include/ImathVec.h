// Minimal stand-in for the Imath vector types the reference's hot path uses.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build): Imath is a third-party dependency of
// jkrueger/phosphorus_mk2 that is not installed in this image.  Only the members the
// reference sources under src/{accel,kernels/cpu,math,...} actually call are provided.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <iostream>
#include <limits>

namespace Imath {

template <typename T> struct Vec2 {
  T x, y;
  Vec2() {}
  explicit Vec2(T a) : x(a), y(a) {}
  Vec2(T a, T b) : x(a), y(b) {}
  template <typename S> Vec2(const Vec2<S>& o) : x(T(o.x)), y(T(o.y)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  Vec2 operator+(const Vec2& o) const { return Vec2(x + o.x, y + o.y); }
  Vec2 operator-(const Vec2& o) const { return Vec2(x - o.x, y - o.y); }
  Vec2 operator-() const { return Vec2(-x, -y); }
  Vec2 operator*(T s) const { return Vec2(x * s, y * s); }
  Vec2 operator*(const Vec2& o) const { return Vec2(x * o.x, y * o.y); }
  Vec2 operator/(T s) const { return Vec2(x / s, y / s); }
  Vec2& operator+=(const Vec2& o) { x += o.x; y += o.y; return *this; }
  Vec2& operator*=(T s) { x *= s; y *= s; return *this; }
  bool operator==(const Vec2& o) const { return x == o.x && y == o.y; }
  T dot(const Vec2& o) const { return x * o.x + y * o.y; }
  T length() const { return std::sqrt(dot(*this)); }
};
template <typename T> inline Vec2<T> operator*(T s, const Vec2<T>& v) { return Vec2<T>(s * v.x, s * v.y); }

template <typename T> struct Vec3 {
  T x, y, z;
  Vec3() {}
  explicit Vec3(T a) : x(a), y(a), z(a) {}
  Vec3(T a, T b, T c) : x(a), y(b), z(c) {}
  template <typename S> Vec3(const Vec3<S>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  Vec3 operator+(const Vec3& o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
  Vec3 operator-(const Vec3& o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
  Vec3 operator-() const { return Vec3(-x, -y, -z); }
  Vec3 operator*(T s) const { return Vec3(x * s, y * s, z * s); }
  Vec3 operator*(const Vec3& o) const { return Vec3(x * o.x, y * o.y, z * o.z); }
  Vec3 operator/(T s) const { return Vec3(x / s, y / s, z / s); }
  Vec3& operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  Vec3& operator-=(const Vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  Vec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
  Vec3& operator*=(const Vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
  Vec3& operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
  bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const Vec3& o) const { return !(*this == o); }
  T dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
  Vec3 cross(const Vec3& o) const {
    return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
  }
  T length2() const { return dot(*this); }
  T length() const { return std::sqrt(dot(*this)); }
  const Vec3& normalize() {
    T l = length();
    if (l != T(0)) { x /= l; y /= l; z /= l; }
    return *this;
  }
  Vec3 normalized() const {
    T l = length();
    if (l == T(0)) return Vec3(T(0));
    return Vec3(x / l, y / l, z / l);
  }
};
template <typename T> inline Vec3<T> operator*(T s, const Vec3<T>& v) { return Vec3<T>(s * v.x, s * v.y, s * v.z); }
template <typename T> inline std::ostream& operator<<(std::ostream& o, const Vec3<T>& v) {
  return o << "(" << v.x << " " << v.y << " " << v.z << ")";
}
template <typename T> inline std::ostream& operator<<(std::ostream& o, const Vec2<T>& v) {
  return o << "(" << v.x << " " << v.y << ")";
}

typedef Vec2<float> V2f;
typedef Vec2<int> V2i;
typedef Vec3<float> V3f;
typedef Vec3<int> V3i;

}  // namespace Imath

// Stand-in for Imath::Box3f (oracle/_ref build only; see ImathVec.h).
// An empty box is min=+FLT_MAX, max=-FLT_MAX exactly like Imath's makeEmpty().
#pragma once
#include "ImathVec.h"
namespace Imath {
template <typename V> struct Box {
  V min, max;
  Box() { makeEmpty(); }
  Box(const V& p) : min(p), max(p) {}
  Box(const V& a, const V& b) : min(a), max(b) {}
  void makeEmpty() {
    min = V(std::numeric_limits<float>::max());
    max = V(std::numeric_limits<float>::lowest());
  }
  void extendBy(const V& p) {
    for (int i = 0; i < 3; ++i) {
      if (p[i] < min[i]) min[i] = p[i];
      if (p[i] > max[i]) max[i] = p[i];
    }
  }
  void extendBy(const Box& b) {
    for (int i = 0; i < 3; ++i) {
      if (b.min[i] < min[i]) min[i] = b.min[i];
      if (b.max[i] > max[i]) max[i] = b.max[i];
    }
  }
  V center() const { return (max + min) / 2; }
  V size() const { return max - min; }
  bool isEmpty() const {
    for (int i = 0; i < 3; ++i) if (max[i] < min[i]) return true;
    return false;
  }
};
typedef Box<V3f> Box3f;
}

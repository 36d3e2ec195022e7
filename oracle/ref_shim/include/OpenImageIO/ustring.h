// Stand-in for OIIO::ustring (oracle/_ref build only): an interned-string look-alike.
#pragma once
#include <string>
#include <iostream>
#include <vector>
namespace OIIO {
struct ustring {
  std::string s;
  ustring() {}
  ustring(const char* c) : s(c) {}
  ustring(const std::string& c) : s(c) {}
  const char* c_str() const { return s.c_str(); }
  const std::string& string() const { return s; }
  bool operator==(const ustring& o) const { return s == o.s; }
  bool operator!=(const ustring& o) const { return s != o.s; }
  bool operator==(const char* o) const { return s == o; }
  bool operator<(const ustring& o) const { return s < o.s; }
};
inline std::ostream& operator<<(std::ostream& o, const ustring& u) { return o << u.s; }
}

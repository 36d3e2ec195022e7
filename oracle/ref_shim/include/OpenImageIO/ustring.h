// Stand-in for OIIO::ustring (oracle/_ref build only): like the real one, a single pointer to an
// interned, immortal character string — trivially copyable, which the reference relies on
// (bsdf_t::add_lobe memcpy's lobe parameters that contain a ustring, src/bsdf.hpp:52-66).
#pragma once
#include <cstring>
#include <iostream>
#include <set>
#include <string>
#include <vector>
namespace OIIO {
struct ustring {
  const char* p;
  static const char* intern(const std::string& s) {
    static std::set<std::string>* table = new std::set<std::string>();
    return table->insert(s).first->c_str();
  }
  ustring() : p(intern("")) {}
  ustring(const char* c) : p(intern(c ? c : "")) {}
  ustring(const std::string& c) : p(intern(c)) {}
  const char* c_str() const { return p; }
  std::string string() const { return std::string(p); }
  bool operator==(const ustring& o) const { return p == o.p || std::strcmp(p, o.p) == 0; }
  bool operator!=(const ustring& o) const { return !(*this == o); }
  bool operator==(const char* o) const { return std::strcmp(p, o) == 0; }
  bool operator<(const ustring& o) const { return std::strcmp(p, o.p) < 0; }
};
inline std::ostream& operator<<(std::ostream& o, const ustring& u) { return o << u.p; }
}

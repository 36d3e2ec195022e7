// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Never linked into the product.
//
// Table-driven replacement for the reference's OSL-backed material system
// (reference: src/material.cpp, not buildable here — OSL 1.11 / OpenImageIO are absent).
// It implements the reference's own `material_t` API (src/material.hpp:53-89) for the OSL nodes
// that make up the renderer's built-in subset, producing exactly the closures the node would hand
// to material_t::details_t::eval_closure (src/material.cpp:218-305):
//   diffuse_bsdf_node.osl:20-25   Cs * diffuse(N), or Cs * oren_nayar(N, roughness) when roughness != 0
//   glossy_bsdf_node.osl:26-34    Cs * microfacet("ggx", N, 0, r*r, r*r, 0, 0), or Cs * reflection(N, 0) when r == 0
//   diffuse_emitter_node.osl:18   (power / M_PI) * Cs * emission()
//   background_node.osl           Cs * power * background()
//   "layered_node"                an explicit closure list = what a mix_closure_node / add_node tree of the
//                                 above plus refraction / sheen / transparent nodes flattens to
// Everything downstream (bsdf_t::add_lobe, precompute, sampling, the integrator) is the
// reference's own compiled code.
#include "material.hpp"
#include "bsdf.hpp"
#include "bsdf/params.hpp"
#include "utils/allocator.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <set>
#include <string>

struct stub_empty_params_t {  // material_t::details_t::empty_params_t, src/material.cpp:98-103
  static const uint32_t flags = bsdf::TRANSMIT;
  void precompute() {}
};

struct stub_lobe_t {
  int type = 0;  // bsdf_t::type_t
  Imath::Color3f weight = Imath::Color3f(1.0f);
  float param = 0.0f;
  float param2 = 0.0f;
};

struct material_t::details_t {
  std::string node;  // OSL node type of the (single) layer, or "layered_node": an explicit closure list
  Imath::Color3f cs;
  float roughness;
  float power;
  std::set<std::string> attributes;
  stub_lobe_t lobes[bsdf_t::MaxLobes];
  int num_lobes = 0;

  details_t() : cs(1.0f), roughness(0.0f), power(1.0f) {}

  bool emitter() const { return node == "diffuse_emitter_node"; }

  // one closure component -> bsdf_t::add_lobe, exactly the calls of eval_closure (src/material.cpp:251-301)
  static void add(bsdf_t* bsdf, const Imath::V3f& n, int type, const Imath::Color3f& cw, float param, float param2 = 0.0f) {
    switch (type) {
      case bsdf_t::Diffuse: {
        bsdf::lobes::diffuse_t p;
        p.n = n;
        bsdf->add_lobe(bsdf_t::Diffuse, cw, &p);
        break;
      }
      case bsdf_t::OrenNayar: {
        bsdf::lobes::oren_nayar_t p;
        p.n = n;
        p.alpha = param;
        bsdf->add_lobe(bsdf_t::OrenNayar, cw, &p);
        break;
      }
      case bsdf_t::Reflection: {
        bsdf::lobes::reflect_t p;
        p.n = n;
        p.eta = param;
        bsdf->add_lobe(bsdf_t::Reflection, cw, &p);
        break;
      }
      case bsdf_t::Refraction: {
        bsdf::lobes::refract_t p;
        p.n = n;
        p.eta = param;
        bsdf->add_lobe(bsdf_t::Refraction, cw, &p);
        break;
      }
      case bsdf_t::Microfacet: {
        bsdf::lobes::microfacet_t p;
        p.distribution = bsdf::lobes::microfacet_t::GGX;
        p.n = n;
        p.u = Imath::V3f(0.0f);
        p.xalpha = param;
        p.yalpha = param;
        p.eta = 0.0f;
        p.refract = 0;
        bsdf->add_lobe(bsdf_t::Microfacet, cw, &p);
        break;
      }
      case bsdf_t::Microfacet | 256: {  // microfacet(ggx, N, 0, r, r, eta, 1): refraction_bsdf_node.osl:37
        bsdf::lobes::microfacet_t p;
        p.distribution = bsdf::lobes::microfacet_t::GGX;
        p.n = n;
        p.u = Imath::V3f(0.0f);
        p.xalpha = param;
        p.yalpha = param;
        p.eta = param2;
        p.refract = 1;
        bsdf->add_lobe(bsdf_t::Microfacet, cw, &p);
        break;
      }
      case bsdf_t::Sheen: {
        bsdf::lobes::sheen_t p;
        p.n = n;
        p.r = param;
        bsdf->add_lobe(bsdf_t::Sheen, cw, &p);
        break;
      }
      case bsdf_t::Transparent: {
        stub_empty_params_t p;
        bsdf->add_lobe(bsdf_t::Transparent, cw, &p);
        break;
      }
      default: break;
    }
  }

  // closure tree of the node -> shading_result_t, as eval_closure would fill it
  void closure(const Imath::V3f& n, shading_result_t& result) const {
    if (node == "diffuse_emitter_node") {
      // MUL(weight = (power / M_PI) * Cs, emission()) -> result.e = cw
      const float k = (float)(power / M_PI);
      result.e = Imath::Color3f(1.0f, 1.0f, 1.0f) * (Imath::Color3f(k * cs.x, k * cs.y, k * cs.z));
    } else if (node == "background_node") {
      // MUL(weight = Cs * power, background()) -> result.e = cw (src/material.cpp:236-238)
      result.e = Imath::Color3f(1.0f, 1.0f, 1.0f) * (Imath::Color3f(cs.x * power, cs.y * power, cs.z * power));
    } else if (node == "glossy_bsdf_node") {
      if (result.bsdf) {
        const float r2 = roughness * roughness;
        // glossy_bsdf_node.osl:28-33: "sharp" or roughness 0 -> reflection(N, 0), else microfacet(ggx, N, 0, r2, r2, 0, 0)
        if (roughness == 0.0f) add(result.bsdf, n, bsdf_t::Reflection, Imath::Color3f(1.0f) * cs, 0.0f);
        else add(result.bsdf, n, bsdf_t::Microfacet, Imath::Color3f(1.0f) * cs, r2);
      }
    } else if (node == "layered_node") {
      if (result.bsdf)
        for (int i = 0; i < num_lobes; ++i) add(result.bsdf, n, lobes[i].type, lobes[i].weight, lobes[i].param, lobes[i].param2);
    } else {  // diffuse_bsdf_node.osl:20-25: roughness 0 -> diffuse(N), else oren_nayar(N, roughness)
      if (result.bsdf) {
        if (roughness == 0.0f) add(result.bsdf, n, bsdf_t::Diffuse, Imath::Color3f(1.0f) * cs, 0.0f);
        else add(result.bsdf, n, bsdf_t::OrenNayar, Imath::Color3f(1.0f) * cs, roughness);
      }
    }
  }
};

namespace {
struct stub_builder_t : public material_t::builder_t {
  material_t* material;
  explicit stub_builder_t(material_t* m) : material(m) {}
  void shader(const std::string& name, const std::string&, const std::string&) override {
    material->details->node = name;
  }
  void connect(const std::string&, const std::string&, const std::string&, const std::string&) override {}
  // "lobe<k>.<field>" parameters describe the closure list of a layered_node
  static bool lobe_field(const std::string& name, int* k, std::string* field) {
    if (name.compare(0, 4, "lobe") != 0 || name.size() < 7 || name[5] != '.') return false;
    *k = name[4] - '0';
    *field = name.substr(6);
    return *k >= 0 && *k < (int)bsdf_t::MaxLobes;
  }
  void parameter(const std::string& name, float f) override {
    if (name == "roughness") material->details->roughness = f;
    if (name == "power") material->details->power = f;
    int k;
    std::string field;
    if (lobe_field(name, &k, &field) && field == "param") material->details->lobes[k].param = f;
    if (lobe_field(name, &k, &field) && field == "param2") material->details->lobes[k].param2 = f;
  }
  void parameter(const std::string& name, int v) override {
    int k;
    std::string field;
    if (lobe_field(name, &k, &field) && field == "type") {
      material->details->lobes[k].type = v;
      material->details->num_lobes = std::max(material->details->num_lobes, k + 1);
    }
  }
  void parameter(const std::string& name, const Imath::Color3f& c) override {
    if (name == "Cs") material->details->cs = c;
    int k;
    std::string field;
    if (lobe_field(name, &k, &field) && field == "weight") material->details->lobes[k].weight = c;
  }
  void parameter(const std::string&, const std::string&) override {}
  void add_attribute(const std::string& name) override { material->details->attributes.insert(name); }
};
}  // namespace

material_t::material_t() : details(new details_t()) {}
material_t::~material_t() { delete details; }

material_t::builder_t* material_t::builder() { return new stub_builder_t(this); }

void material_t::evaluate(allocator_t& allocator, interaction_t<>* hits, const active_t<>& active) {
  for (uint32_t i = 0; i < active.num; ++i) {
    const auto index = active.index[i];
    shading_result_t result;
    result.bsdf = new (allocator) bsdf_t();
    // The reference leaves the lobe arrays of a fresh bsdf_t uninitialised (arena memory, src/material.cpp:441-443)
    // and still samples lobe 0 of a 0-lobe BSDF when a bounce ray lands on an emitter (src/bsdf.cpp:140-153,
    // SURVEY.md F7): undefined behaviour that replays whatever lobe an earlier hit left at that address.
    // Zeroing the object freezes that to "type Emissive, weight 0 -> black sample -> the path ends", which is
    // the behaviour the oracle and the GPU implement.
    memset((void*)result.bsdf, 0, sizeof(bsdf_t));
    details->closure(hits->n.at(index), result);
    hits->e.from(index, result.e);
    hits->bsdf[index] = result.bsdf;
  }
}

void material_t::evaluate(const Imath::V3f&, const Imath::V3f&, const Imath::V3f& n, const Imath::V2f&,
                          shading_result_t& result) {
  result.bsdf = nullptr;
  details->closure(n, result);
}

bool material_t::is_emitter() const { return details->emitter(); }
bool material_t::has_attribute(const std::string& name) const { return details->attributes.count(name); }
void material_t::attach() {}
void material_t::boot(const parsed_options_t&, const std::string&) {}

// Which built-in closure is this material?  (Asked by the GPU device glue, integration/cuda.cpp; with the real
// OSL material system the same answer comes from querying the shader group's layers and parameters.)
bool material_builtin_closure(const material_t* m, uint32_t* kind, float cs[3], float* roughness, float* power) {
  const auto* d = m->details;
  if (d->node == "diffuse_emitter_node") *kind = 2;
  else if (d->node == "glossy_bsdf_node") *kind = 1;
  else if (d->node == "diffuse_bsdf_node") *kind = 0;
  else if (d->node == "background_node") *kind = 3;
  else if (d->node == "layered_node") *kind = 4;
  else return false;
  cs[0] = d->cs.x; cs[1] = d->cs.y; cs[2] = d->cs.z;
  *roughness = d->roughness;
  *power = d->power;
  return true;
}

// the closure list of a layered_node (kind 4 above): type / weight / param per lobe; returns the count
int material_builtin_lobes(const material_t* m, uint32_t* type, float* weight3, float* param, float* param2) {
  const auto* d = m->details;
  for (int i = 0; i < d->num_lobes; ++i) {
    type[i] = (uint32_t)d->lobes[i].type;
    weight3[3 * i] = d->lobes[i].weight.x;
    weight3[3 * i + 1] = d->lobes[i].weight.y;
    weight3[3 * i + 2] = d->lobes[i].weight.z;
    param[i] = d->lobes[i].param;
    param2[i] = d->lobes[i].param2;
  }
  return d->num_lobes;
}

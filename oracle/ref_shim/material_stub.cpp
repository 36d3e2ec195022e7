// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Never linked into the product.
//
// Table-driven replacement for the reference's OSL-backed material system
// (reference: src/material.cpp, not buildable here — OSL 1.11 / OpenImageIO are absent).
// It implements the reference's own `material_t` API (src/material.hpp:53-89) for the three
// OSL nodes that make up the renderer's built-in subset, producing exactly the closure the node
// would hand to material_t::details_t::eval_closure (src/material.cpp:218-305):
//   diffuse_bsdf_node.osl:20-25   (roughness 0)   Cs * diffuse(N)
//   glossy_bsdf_node.osl:26-34    (ggx, r > 0)    Cs * microfacet("ggx", N, 0, r*r, r*r, 0, 0)
//   diffuse_emitter_node.osl:18                   (power / M_PI) * Cs * emission()
// Everything downstream (bsdf_t::add_lobe, precompute, sampling, the integrator) is the
// reference's own compiled code.
#include "material.hpp"
#include "bsdf.hpp"
#include "bsdf/params.hpp"
#include "utils/allocator.hpp"

#include <cmath>
#include <map>
#include <set>
#include <string>

struct material_t::details_t {
  std::string node;  // OSL node type of the (single) layer
  Imath::Color3f cs;
  float roughness;
  float power;
  std::set<std::string> attributes;

  details_t() : cs(1.0f), roughness(0.0f), power(1.0f) {}

  bool emitter() const { return node == "diffuse_emitter_node"; }

  // closure tree of the node -> shading_result_t, as eval_closure would fill it
  void closure(const Imath::V3f& n, shading_result_t& result) const {
    if (node == "diffuse_emitter_node") {
      // MUL(weight = (power / M_PI) * Cs, emission()) -> result.e = cw
      const float k = (float)(power / M_PI);
      result.e = Imath::Color3f(1.0f, 1.0f, 1.0f) * (Imath::Color3f(k * cs.x, k * cs.y, k * cs.z));
    } else if (node == "glossy_bsdf_node") {
      if (result.bsdf) {
        const float r2 = roughness * roughness;
        bsdf::lobes::microfacet_t p;
        p.distribution = bsdf::lobes::microfacet_t::GGX;
        p.n = n;
        p.u = Imath::V3f(0.0f);
        p.xalpha = r2;
        p.yalpha = r2;
        p.eta = 0.0f;
        p.refract = 0;
        result.bsdf->add_lobe(bsdf_t::Microfacet, Imath::Color3f(1.0f) * cs, &p);
      }
    } else {  // diffuse_bsdf_node
      if (result.bsdf) {
        bsdf::lobes::diffuse_t p;
        p.n = n;
        result.bsdf->add_lobe(bsdf_t::Diffuse, Imath::Color3f(1.0f) * cs, &p);
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
  void parameter(const std::string& name, float f) override {
    if (name == "roughness") material->details->roughness = f;
    if (name == "power") material->details->power = f;
  }
  void parameter(const std::string&, int) override {}
  void parameter(const std::string& name, const Imath::Color3f& c) override {
    if (name == "Cs") material->details->cs = c;
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
  else return false;
  cs[0] = d->cs.x; cs[1] = d->cs.y; cs[2] = d->cs.z;
  *roughness = d->roughness;
  *power = d->power;
  return true;
}

/* phos_scene.h — flat, C-ABI scene description shared by the CUDA device library, the host BVH
 * build and (as input only) the test oracle.
 *
 * It carries exactly what a device can read through the reference's public scene API
 * (reference: src/scene.hpp:14-50, src/mesh.hpp:68-77, src/triangle.hpp:12-36,
 * src/entities/camera.hpp:10-40) flattened into plain arrays so it can cross a C boundary:
 * meshes in scene order, every mesh's face sets in set order, every set's faces in set order —
 * the same mesh -> set -> face order scene_t::triangles() walks (src/scene.cpp:58-62,
 * src/mesh.cpp:118-128), which is what fixes triangle numbering for the BVH build.
 */
#ifndef PHOS_SCENE_H
#define PHOS_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Built-in closure subset (stands in for the three OSL nodes the north star keeps:
 * src/shaders/diffuse_bsdf_node.osl:20-25, glossy_bsdf_node.osl:26-34,
 * diffuse_emitter_node.osl:18). */
enum {
  PHOS_MAT_DIFFUSE    = 0, /* diffuse_bsdf_node: Cs * diffuse(N), or Cs * oren_nayar(N, roughness) if roughness != 0 */
  PHOS_MAT_GLOSSY     = 1, /* glossy_bsdf_node: Cs * microfacet("ggx", N, 0, r*r, r*r, 0, 0), or Cs * reflection(N, 0)
                              if roughness == 0 (glossy_bsdf_node.osl:28-33)                                        */
  PHOS_MAT_EMITTER    = 2, /* diffuse_emitter_node: (power / pi) * Cs * emission()                                  */
  PHOS_MAT_BACKGROUND = 3, /* background_node: Cs * power * background() — only as phos_scene_desc.environment      */
  PHOS_MAT_LAYERED    = 4  /* an explicit closure list (what mix_closure_node / add_node trees flatten to in
                              material_t::details_t::eval_closure, src/material.cpp:218-305): lobes[0..num_lobes)   */
};

/* Closure ids = bsdf_t::type_t (src/bsdf.hpp:14-24).  `param`: oren_nayar sigma in degrees
 * (bsdf/params.hpp:31-42), microfacet xalpha = yalpha BEFORE precompute (the glossy node passes roughness^2),
 * reflection / refraction eta, sheen roughness; unused otherwise.  Microfacet is the GGX reflection lobe
 * (refract = 0), MICROFACET_REFRACT the GGX transmission lobe (src/bsdf/microfacet.hpp:36-172). */
enum {
  PHOS_LOBE_DIFFUSE     = 1,
  PHOS_LOBE_OREN_NAYAR  = 2,
  PHOS_LOBE_REFLECTION  = 4,
  PHOS_LOBE_REFRACTION  = 8,
  PHOS_LOBE_MICROFACET  = 16,
  PHOS_LOBE_SHEEN       = 32,
  PHOS_LOBE_TRANSPARENT = 128,
  PHOS_LOBE_MICROFACET_REFRACT = 16 | 256 /* bsdf_t::Microfacet with microfacet_t::refract = 1 (rough glass,
                                             refraction_bsdf_node.osl:37): param = xalpha = yalpha before precompute
                                             (the node passes roughness, not its square), param2 = eta */
};
#define PHOS_MAX_LOBES 8 /* bsdf_t::MaxLobes */

typedef struct phos_lobe {
  uint32_t type;      /* PHOS_LOBE_*            */
  float    weight[3]; /* closure weight (color) */
  float    param;
  float    param2;    /* second parameter (eta of the rough-refraction lobe), 0 otherwise */
} phos_lobe;

typedef struct phos_material {
  uint32_t  kind;      /* PHOS_MAT_*                                          */
  float     cs[3];     /* Cs                                                  */
  float     roughness; /* diffuse / glossy node parameter (not alpha)         */
  float     power;     /* emitter / background                                */
  uint32_t  num_lobes; /* PHOS_MAT_LAYERED only, <= PHOS_MAX_LOBES            */
  phos_lobe lobes[PHOS_MAX_LOBES];
} phos_material;

/* camera_t (src/entities/camera.hpp:10-40); to_world is Imath row-vector convention:
 * p_world = p * M (translation in row 3). */
typedef struct phos_camera {
  float    to_world[16];
  float    fov; /* radians */
  float    focal_distance;
  float    aperture_radius; /* 0 = pinhole */
  uint32_t film_width, film_height;
} phos_camera;

typedef struct phos_scene_desc {
  uint32_t        num_meshes;
  const uint32_t* vert_offset;     /* [num_meshes+1] first vertex of mesh m in `vertices`     */
  const float*    vertices;        /* xyz per vertex                                          */
  const float*    normals;         /* xyz per vertex (NormalsPerVertex) or NULL               */
  const uint32_t* face_offset;     /* [num_meshes+1] first face of mesh m in `faces`          */
  const uint32_t* faces;           /* 3 mesh-local vertex indices per face                    */
  const uint8_t*  mesh_smooth;     /* [num_meshes] 1: every face smooth, 0: every face flat, 2: see face_smooth */
  const uint32_t* set_offset;      /* [num_meshes+1] first face set of mesh m                 */
  const uint32_t* set_material;    /* [num_sets] material id of the set                       */
  const uint32_t* set_face_offset; /* [num_sets+1] first entry of set s in `set_faces`        */
  const uint32_t* set_faces;       /* mesh-local face indices                                 */
  uint32_t             num_materials;
  const phos_material* materials;
  phos_camera          camera;
  int32_t              environment; /* material id of the environment (PHOS_MAT_BACKGROUND; scene_t::environment(),
                                       src/scene.cpp:64-74,126-128) or -1: what a ray that misses adds to the path */
  const uint8_t*       face_smooth; /* [total faces] or NULL: the per-face flag of mesh_t::builder_t::add_face(a, b, c, smooth)
                                       (src/mesh.hpp:46-66, src/mesh.cpp:10-18,202-206), indexed like `faces` / 3; read for
                                       the faces of meshes whose mesh_smooth is 2 (meshes mixing smooth and flat faces) */
} phos_scene_desc;

#ifdef __cplusplus
}
#endif
#endif /* PHOS_SCENE_H */

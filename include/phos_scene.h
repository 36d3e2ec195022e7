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
  PHOS_MAT_DIFFUSE = 0, /* Cs * diffuse(N)                                  */
  PHOS_MAT_GLOSSY  = 1, /* Cs * microfacet("ggx", N, 0, r*r, r*r, 0, 0)     */
  PHOS_MAT_EMITTER = 2  /* (power / pi) * Cs * emission()                   */
};

typedef struct phos_material {
  uint32_t kind;      /* PHOS_MAT_*                          */
  float    cs[3];     /* Cs                                  */
  float    roughness; /* glossy only (node parameter, not alpha) */
  float    power;     /* emitter only                        */
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
  const uint8_t*  mesh_smooth;     /* [num_meshes] 1: every face smooth, 0: every face flat   */
  const uint32_t* set_offset;      /* [num_meshes+1] first face set of mesh m                 */
  const uint32_t* set_material;    /* [num_sets] material id of the set                       */
  const uint32_t* set_face_offset; /* [num_sets+1] first entry of set s in `set_faces`        */
  const uint32_t* set_faces;       /* mesh-local face indices                                 */
  uint32_t             num_materials;
  const phos_material* materials;
  phos_camera          camera;
} phos_scene_desc;

#ifdef __cplusplus
}
#endif
#endif /* PHOS_SCENE_H */

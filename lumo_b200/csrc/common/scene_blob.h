// Device scene blob: the flattened, compacted form of a built lumo scene (two-level: object BVH
// -> per-mesh kd-tree -> triangles), produced once on the host (csrc/host/host_build.cpp, from
// lumo's own SAH kd-tree / Garanzha BVH build) and uploaded verbatim by lumo_gpu_scene_upload.
// Plain little-endian PODs, every section 16-byte aligned; shared by the host builder and the
// CUDA kernels.  Layout follows SURVEY.md Appendix B ("faithful layout"): node order, leaf order
// and child placement (left child = index+1) are exactly lumo's, so traversal visits the same
// primitives in the same order as the reference (src/tracer/object/bvh.rs:315-362,
// kdtree.rs:101-169) and hit ids are bit-exact.
#pragma once
#include <stdint.h>

#define LUMO_BLOB_MAGIC 0x31424F4C424D554CULL /* "LUMBLOB1" */
#define LUMO_BLOB_VERSION 6u
#define LUMO_NONE 0xFFFFFFFFu

enum LumoSection {
    LSEC_TLAS_NODES = 0,  // LumoTlasNode[]; objects BVH first, lights BVH after (params.lights_root)
    LSEC_TLAS_LEAF,       // u32[]  leaf object lists (indices local to the BVH they belong to)
    LSEC_OBJECTS,         // LumoObject[]  Scene.objects then Scene.lights (insertion order)
    LSEC_INSTANCES,       // LumoInstance[]
    LSEC_KD_TREES,        // LumoKdTree[]
    LSEC_KD_NODES,        // LumoKdNode[]  all trees, DFS pre-order each
    LSEC_KD_LEAF,         // u32[]  leaf triangle lists (indices local to the tree)
    LSEC_TRI_VERTS,       // LumoTriVerts[]  a,b,c as f64 (80 B, 5 x 16 B vector loads)
    LSEC_TRI_SHADE,       // LumoTriShade[]
    LSEC_NORMALS,         // f64[3] per entry
    LSEC_UVS,             // f64[2] per entry
    LSEC_RECTS,           // LumoRect[]
    LSEC_SPHERES,         // LumoSphere[]
    LSEC_MATERIALS,       // LumoMaterial[]
    LSEC_TABLES,          // f64[96] per table (95 samples, 5 nm, 360..830; +1 pad)
    LSEC_LIGHTS,          // LumoLight[]  alias table + pdf + area, one per Scene.lights entry
    LSEC_TEXTURES,        // LumoTexture[]  (texture.rs:23-38); materials refer to them by index
    LSEC_TEX_PIXELS,      // f32[4] per pixel: Image<Spectrum>.buffer as Spectrum {c0,c1,c2,scale}, row-major (image.rs:8-16)
    LSEC_TEX_F64,         // f64 pool: Perlin tables (256 x 3 lattice normals, then 3 x 256 permutation entries as f64) and
                          //           Image<Normal> buffers (3 f64 per pixel)
    LSEC_AH_NODES,        // LumoAhNode[]  4-wide BVH over every primitive of the scene in world space (order-free occlusion queries)
    LSEC_AH_PRIMS,        // LumoAhPrim[]  leaf primitive lists of that BVH
    LSEC_OBJ_PATH_OFF,    // u32[n_objects + n_lights + 1]  CSR offsets into LSEC_OBJ_PATH
    LSEC_OBJ_PATH,        // u32[]  per object: the object-BVH nodes from its root down to the leaf that lists it (absolute indices into LSEC_TLAS_NODES)
    LSEC_COUNT
};

enum LumoObjKind { LOBJ_KD = 0, LOBJ_RECT = 1, LOBJ_SPHERE = 2, LOBJ_TRI = 3 };
enum LumoMatKind { LMAT_BLANK = 0, LMAT_LAMBERTIAN = 1, LMAT_MFDIFFUSE = 2, LMAT_MFCONDUCTOR = 3, LMAT_MFDIELECTRIC = 4, LMAT_LIGHT = 5 };
enum LumoMatFlags { LMF_ETA_CONST = 1, LMF_TWO_SIDED = 2 };
enum LumoTables { LTAB_X = 0, LTAB_Y = 1, LTAB_Z = 2, LTAB_ILLUM0 = 3 /* A,D50,D65,F2,F7,CORNELL */, LTAB_FIRST_FREE = 9 };

struct LumoTlasNode {          // 64 B  (bvh/node.rs:8-14 flattened)
    double lo[3], hi[3];
    uint32_t right;            // LUMO_NONE if single child; left child is always index + 1
    uint32_t first, count;     // leaf: range in LSEC_TLAS_LEAF; count == 0 -> inner node
    uint32_t pad;
};
struct LumoObject {            // 32 B
    uint32_t kind;             // LumoObjKind
    uint32_t geom;             // kd tree id | sphere id | global triangle index
    int32_t inst;              // instance id or -1
    int32_t material;          // resolved material (Instance override applied)
    uint32_t rect;             // rect id for LOBJ_RECT (uv override + sampling)
    uint32_t pad[3];
};
struct LumoInstance {          // 272 B (instance.rs:5-15): rows of the affine 3x4 parts + normal matrix
    double inv[12];            // world -> local
    double m[12];              // local -> world
    double nrm[9];             // transpose(inv 3x3)
    double pad;
};
struct LumoKdTree {            // 64 B
    double lo[3], hi[3];       // KdTree.boundary
    uint32_t root, tri_base, n_tris, pad;
};
struct LumoKdNode {            // 16 B  (kdtree/node.rs:24-30 flattened)
    double point;              // split coordinate (inner)
    uint32_t a;                // inner: right child (absolute node index); leaf: first entry in LSEC_KD_LEAF
    uint32_t b;                // inner: axis 0..2; leaf: 0x80000000 | count
};
// Occlusion (`Scene::hit_light`, scene.rs:165-189) asks a boolean that does not depend on the visit order, so it is
// answered from a second structure over the same primitives: a 4-wide BVH in world space with f32 boxes rounded outwards
// (conservative culling only — every accepted primitive is confirmed by the reference's own f64 test and the reference's
// own per-object traversal, csrc/gpu/occlude.cuh).  128 B = one cache line = eight 16-byte loads.
#define LUMO_AH_LEAF 0x80000000u       /* child is a leaf: bits 0..26 first primitive, bits 27..30 count - 1 */
#define LUMO_AH_SPHERE 0x80000000u     /* LumoAhPrim.tri: bit 31 set -> sphere, low bits = index into LSEC_SPHERES */
#define LUMO_AH_INSTANCED 0x80000000u  /* LumoAhPrim.obj: bit 31 set -> the object has an instance transform */
#define LUMO_AH_MAX_LEAF 4u
struct LumoAhNode {            // 128 B
    float lo_x[4], lo_y[4], lo_z[4], hi_x[4], hi_y[4], hi_z[4];   // child boxes, structure of arrays; empty child: lo = +inf, hi = -inf
    uint32_t child[4];         // LUMO_NONE: empty; LUMO_AH_LEAF | ...: leaf; else index of the child node
    uint32_t pad[4];
};
struct LumoAhPrim {            // 8 B
    uint32_t tri;              // global triangle index (LSEC_TRI_VERTS), or LUMO_AH_SPHERE | sphere index
    uint32_t obj;              // global object index (Scene.objects then Scene.lights) | LUMO_AH_INSTANCED: frame of the exact test, and the object the confirming traversal runs on
};
struct LumoTriVerts { double a[3], b[3], c[3], pad; };   // 80 B
struct LumoTriShade {          // 32 B
    uint32_t n[3], t[3];       // indices into LSEC_NORMALS / LSEC_UVS
    uint32_t flags;            // bit0 has normals, bit1 has uvs
    uint32_t pad;
};
struct LumoRect { double origin[3], b0[3], b1[3], pad; };  // 80 B (rectangle.rs:4-13)
struct LumoSphere { double radius, pad; };
enum LumoTexKind { LTEX_SOLID = 0, LTEX_CHECKER = 1, LTEX_MARBLE = 2, LTEX_IMAGE = 3, LTEX_MANDELBROT = 4, LTEX_BUMP = 5 };
struct LumoTexture {           // 64 B  (texture.rs:23-38; LTEX_BUMP is the Image<Normal> of Material::Standard, material.rs:14)
    uint32_t kind;
    uint32_t a, b;             // Checkerboard: the two child textures
    uint32_t width, height;    // Image / bump
    uint32_t pad;
    uint64_t data;             // Image: first pixel in LSEC_TEX_PIXELS; Marble / bump: first f64 in LSEC_TEX_F64
    float spec[4];             // Solid / Marble colour; Image: the mean spectrum (Texture::power)
    double scale;              // Checkerboard scale
    double pad2;
};
struct LumoMaterial {          // 128 B
    uint32_t kind, flags;
    double roughness;          // clamped to >= 1e-5 (microfacet.rs:30-36), isotropic
    float kd[4], ks[4], tf[4], ke[4];   // Texture::Solid spectra {c0,c1,c2,scale} (spectrum.rs:14-19); Lambertian uses kd
    uint32_t eta_table, k_table, illum_table;
    uint32_t kd_tex;           // LUMO_NONE: the Solid spectrum above; else index into LSEC_TEXTURES (same for ks/tf/ke)
    double scale;              // Light scale
    uint32_t ks_tex, tf_tex, ke_tex;
    uint32_t bump_tex;         // LUMO_NONE or the LTEX_BUMP record of the normal map
    uint64_t pad;
};
struct LumoLight {             // 64 B  (bvh.rs:24-25 alias_table / alias_pdf)
    double alias_prob, pdf, area;
    uint32_t alias, pad;
    // World-space sphere that contains the light with room to spare (centre of its bounding box, half diagonal padded by
    // 1e-4 of itself + 1e-6 of the coordinates' magnitude): a ray whose line misses it cannot pass the light's intersection
    // test, so the BSDF-sampled NEE term skips that test (k_nee_b).  Not part of the reference's data; never changes a result.
    double bound_c[3], bound_r;
};

struct LumoCamera {
    double screen_to_raster_m[16], screen_to_raster_inv[16];
    double camera_to_screen_m[16], camera_to_screen_inv[16];
    double world_to_camera_m[16], world_to_camera_inv[16];
    double lens_radius, focal_length, image_plane_area, lens_area;
    uint32_t res_x, res_y, ortho, pad;
};
struct LumoFilm {
    double xyz_to_rgb[9], wb[9];
    double filter_r, filter_p, filter_gr;   // filter_gr = gauss(r, sigma) for the Gaussian
    uint32_t filter_kind, r_disc, color_space, pad;
};
struct LumoSceneParams {
    uint32_t n_objects, n_lights, lights_root, n_shadow_rays;
    uint32_t n_tlas_nodes, n_kd_trees, n_materials, n_tris;
    double bounds_lo[3], bounds_hi[3];
    LumoCamera camera;
    LumoFilm film;
};
struct LumoSectionRef { uint64_t offset, bytes, count, pad; };
struct LumoBlobHeader {
    uint64_t magic;
    uint32_t version, n_sections;
    uint64_t total_bytes, pad;
    LumoSceneParams params;
    LumoSectionRef sec[LSEC_COUNT];
};

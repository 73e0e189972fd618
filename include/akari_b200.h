/*
 * akari_b200.h — C-ABI boundary of the B200-native unidirectional path tracer.
 *
 * This is the drop-in seam for AkariRender's `pt` integrator.  The reference has no FFI for
 * integrators; the replaceable call is
 *     akari_integrator::pt::render(device, scene, sampler, color_pipeline, film, config, options)
 *     (reference: crates/akari_integrator/src/pt.rs:1161-1172, called from
 *      crates/akari_integrator/src/lib.rs:134-142).
 * A Rust host keeps doing scene load, shader-graph flattening (svm/compiler.rs) and hands this
 * library plain arrays; everything below `akr_b200_render_pt` runs as sm_100a CUDA kernels.
 *
 * Conventions (imitating the repo's own FFI, crates/akari_cpp_ext/cpp_ext/akari_cpp_ext.h:14-49
 * and crates/akari_api/src/lib.rs:6-26): plain C, POD structs, raw pointers + counts, caller
 * allocates outputs, every entry point returns an int status (0 = ok) and never aborts or throws
 * across the boundary; the context is not thread-safe (serialise calls per context), independent
 * contexts may run concurrently.  All host pointers are borrowed for the duration of the call only.
 *
 * There is NO CPU fallback: if no CUDA device is usable every call returns AKR_ERR_CUDA.
 */
#ifndef AKARI_B200_H
#define AKARI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AKR_B200_ABI_VERSION 2u

/* ---- status codes -------------------------------------------------------------------------- */
enum {
    AKR_OK = 0,
    AKR_ERR_INVALID_ARGUMENT = 1,
    AKR_ERR_CUDA = 2,            /* CUDA runtime/driver error, or no device          */
    AKR_ERR_UNSUPPORTED = 3,     /* feature outside the implemented hot-path scope   */
    AKR_ERR_OUT_OF_MEMORY = 4,
    AKR_ERR_STATE = 5            /* call order violated (e.g. render before upload)  */
};

/* ---- sampler tables (reference: crates/akari_data/src/pmj02bn.rs:2-5, bluenoise.rs:109-113) -- */
#define AKR_PMJ02BN_SETS 5u
#define AKR_PMJ02BN_SAMPLES 65536u
#define AKR_BLUE_NOISE_TEXTURES 48u
#define AKR_BLUE_NOISE_RESOLUTION 128u

/* ---- shader virtual machine, host-flattened --------------------------------------------------
 * Mirrors svm::SvmNode (crates/akari_render/src/svm/mod.rs:191-211) and ShaderRef (:213-219).
 * One AkrShaderKind = one distinct bytecode (identical bytecode => same kind,
 * svm/compiler.rs:25-46).  Node arguments are node indices inside the same kind ("relative
 * index") or byte offsets into the constant blob, relative to ShaderRef.data_offset.          */
enum {
    AKR_SVM_FLOAT = 0,            /* a0 = const offset (f32)                               */
    AKR_SVM_FLOAT3 = 1,           /* a0 = const offset (Float3, 16-byte aligned)           */
    AKR_SVM_RGB_TEX = 2,          /* a0 = node(Float3), a1 = colorspace id (1 = sRGB)      */
    AKR_SVM_SPECTRAL_UPLIFT = 3,  /* a0 = node(rgba)                                       */
    AKR_SVM_EMISSION = 4,         /* a0 = color node, a1 = strength node                   */
    AKR_SVM_DIFFUSE_BSDF = 5,     /* a0 = reflectance node                                 */
    AKR_SVM_GLASS_BSDF = 6,       /* a0 = kr, a1 = kt, a2 = roughness, a3 = eta            */
    AKR_SVM_PRINCIPLED_BSDF = 7,  /* a[0..25) in the field order of SvmPrincipledBsdf      */
    AKR_SVM_MATERIAL_OUTPUT = 8,  /* a0 = surface closure node                             */
    /* texture-driven nodes (svm/compiler.rs:135-337, svm/eval.rs:137-269); AKR_SVM_NONE = "input not connected"  */
    AKR_SVM_RGB_IMAGE_TEX = 9,    /* a0 = const offset (u32 texture index into AkrSceneDesc.images),
                                   * a1 = colorspace id (0 = none, 1 = sRGB: decode with srgb_to_linear),
                                   * a2 = uv node or AKR_SVM_NONE (then the hit's uv)                          */
    AKR_SVM_NORMAL_MAP = 10,      /* a0 = normal node, a1 = strength node (tangent space only)               */
    AKR_SVM_MAPPING = 11,         /* a0 = vector, a1 = type (0 point, 1 texture), a2 = location, a3 = rotation
                                   * (ignored, as in the reference), a4 = scale                                */
    AKR_SVM_EXTRACT_FIELD = 12,   /* a0 = node, a1 = field (AKR_SVM_FIELD_*)                                  */
    AKR_SVM_TEX_COORDS = 13,      /* the hit's uv; only its field "uv" can be extracted                       */
    AKR_SVM_CHECKERBOARD = 14,    /* a0 = vector node or AKR_SVM_NONE, a1 = scale, a2 = color1, a3 = color2  */
    AKR_SVM_SEPARATE_COLOR = 15   /* a0 = color node; fields Red / Green / Blue                               */
};
#define AKR_SVM_NONE 0xffffffffu
enum { AKR_SVM_FIELD_UV = 0, AKR_SVM_FIELD_RED = 1, AKR_SVM_FIELD_GREEN = 2, AKR_SVM_FIELD_BLUE = 3 };
#define AKR_SVM_MAX_ARGS 25u

/* Field order of a[] for AKR_SVM_PRINCIPLED_BSDF (svm/mod.rs:151-178). */
enum {
    AKR_P_BASE_COLOR = 0, AKR_P_METALLIC, AKR_P_ROUGHNESS, AKR_P_IOR, AKR_P_ALPHA, AKR_P_NORMAL,
    AKR_P_SUBSURFACE_WEIGHT, AKR_P_SUBSURFACE_RADIUS, AKR_P_SUBSURFACE_SCALE,
    AKR_P_SUBSURFACE_ANISOTROPY, AKR_P_SPECULAR_IOR_LEVEL, AKR_P_SPECULAR_TINT, AKR_P_ANISOTROPIC,
    AKR_P_ANISOTROPIC_ROTATION, AKR_P_TANGENT, AKR_P_TRANSMISSION_WEIGHT, AKR_P_SHEEN_WEIGHT,
    AKR_P_SHEEN_TINT, AKR_P_COAT_WEIGHT, AKR_P_COAT_ROUGHNESS, AKR_P_COAT_IOR, AKR_P_COAT_TINT,
    AKR_P_COAT_NORMAL, AKR_P_EMISSION_COLOR, AKR_P_EMISSION_STRENGTH
};

typedef struct AkrSvmNode {
    uint32_t op;
    uint32_t n_args;
    uint32_t a[AKR_SVM_MAX_ARGS];
} AkrSvmNode;

typedef struct AkrShaderKind {
    const AkrSvmNode *nodes;   /* evaluation order = array order; last node is the output */
    uint32_t n_nodes;
} AkrShaderKind;

typedef struct AkrShaderRef {      /* svm/mod.rs:213-219 */
    uint32_t shader_kind;
    uint32_t data_offset;          /* bytes into shader_data */
} AkrShaderRef;

/* ---- geometry (reference: crates/akari_render/src/mesh.rs:14-25,189-241) --------------------- */
enum {                             /* MeshInstanceFlags, mesh.rs:190-196 */
    AKR_MESH_HAS_NORMALS = 1u << 0,
    AKR_MESH_HAS_UVS = 1u << 1,
    AKR_MESH_HAS_TANGENTS = 1u << 2,
    AKR_MESH_HAS_MULTI_MATERIALS = 1u << 3
};

typedef struct AkrMesh {
    const float *vertices;          /* [n_vertices][3], indexed                                */
    const uint32_t *indices;        /* [n_triangles][3]                                        */
    const float *normals;           /* per corner [3*n_triangles][3], or NULL                  */
    const float *uvs;               /* per corner [3*n_triangles][2], or NULL                  */
    const float *tangents;          /* per corner [3*n_triangles][3], or NULL (only when the
                                       scene file supplies them, mesh.rs:180-182,277-281)      */
    const uint32_t *material_slots; /* [n_material_slots]; one entry = single material,
                                       otherwise one per triangle (load.rs:213-215)            */
    uint32_t n_vertices;
    uint32_t n_triangles;
    uint32_t n_material_slots;
    uint32_t _pad;
} AkrMesh;

typedef struct AkrInstance {
    float transform[16];            /* column-major 4x4 (glam Mat4 layout), MeshInstanceHost.transform.m */
    uint32_t geom_id;
    uint32_t flags;                 /* AKR_MESH_* */
    const AkrShaderRef *materials;  /* [n_materials] */
    uint32_t n_materials;
    uint32_t _pad;
} AkrInstance;

/* ---- image textures (reference: load.rs:536-646,680-702; one entry per (image, sampler) pair = one bindless slot) ------
 * The host decodes the file formats (png / jpeg / tiff / exr / dds / raw float, load.rs:550-610) and hands over texels:
 * RGBA8 unorm for the 8-bit formats, RGBA32F for exr / raw float, row 0 first as stored in the texture.             */
enum { AKR_TEXEL_RGBA8 = 0, AKR_TEXEL_RGBA32F = 1 };
enum { AKR_ADDRESS_REPEAT = 0, AKR_ADDRESS_ZERO = 1, AKR_ADDRESS_MIRROR = 2, AKR_ADDRESS_EDGE = 3 };   /* load.rs:684-689 */
enum { AKR_FILTER_POINT = 0, AKR_FILTER_LINEAR = 1 };                                                   /* load.rs:690-699 */
typedef struct AkrImage {
    const void *texels;             /* [height][width][4] u8 or f32                                */
    uint32_t width, height;
    uint32_t texel_format;          /* AKR_TEXEL_*                                                  */
    uint32_t address;               /* AKR_ADDRESS_*                                                */
    uint32_t filter;                /* AKR_FILTER_*                                                 */
    uint32_t _pad;
} AkrImage;

/* ---- camera (reference: crates/akari_render/src/camera/mod.rs:108-153) ------------------------ */
typedef struct AkrPerspectiveCamera {
    float c2w[16];                  /* column-major camera-to-world (load.rs:129-171)          */
    float fov;                      /* radians (load.rs:176)                                    */
    float lens_radius;              /* unused by generate_ray (camera/mod.rs:70-103)            */
    float focal_length;
    uint32_t width, height;         /* sensor resolution                                        */
    uint32_t _pad;
} AkrPerspectiveCamera;

/* ---- the scene blob ------------------------------------------------------------------------- */
typedef struct AkrSceneDesc {
    uint32_t abi_version;           /* AKR_B200_ABI_VERSION */
    uint32_t n_meshes;
    uint32_t n_instances;
    uint32_t n_shader_kinds;
    const AkrMesh *meshes;
    const AkrInstance *instances;
    const AkrShaderKind *shader_kinds;
    const uint8_t *shader_data;     /* constant blob, each material's block padded to 16 B      */
    size_t shader_data_size;
    AkrPerspectiveCamera camera;
    const AkrImage *images;         /* [n_images], indexed by the u32 an AKR_SVM_RGB_IMAGE_TEX node reads */
    uint32_t n_images;
    uint32_t _pad;
} AkrSceneDesc;

/* ---- render configuration -------------------------------------------------------------------
 * AkrPtConfig mirrors pt::Config (crates/akari_integrator/src/pt.rs:916-944) field for field.    */
typedef struct AkrPtConfig {
    uint32_t spp;                   /* total samples per pixel (sampler permutation length)     */
    uint32_t max_depth;
    uint32_t spp_per_pass;
    uint32_t rr_depth;
    uint32_t use_nee;               /* bool */
    uint32_t indirect_only;         /* bool */
    uint32_t force_diffuse;         /* bool */
    int32_t pixel_offset[2];
    int32_t debug_depth;            /* < 0 = None */
} AkrPtConfig;

enum { AKR_SAMPLER_INDEPENDENT = 0, AKR_SAMPLER_PMJ02BN = 1 };   /* sampler/mod.rs:282-295 */
typedef struct AkrSamplerConfig {
    uint32_t type;
    uint32_t _pad;
    uint64_t seed;
} AkrSamplerConfig;

enum { AKR_FILTER_BOX = 0, AKR_FILTER_GAUSSIAN = 1 };             /* film.rs:24-30 */
typedef struct AkrFilterConfig {
    uint32_t type;
    float radius;
} AkrFilterConfig;

/* AOV integrator (crates/akari_integrator/src/aov.rs:9-36): what is written to the film for the first hit of every
 * camera sample instead of radiance.  `remap` maps the vector outputs v -> v * 0.5 + 0.5 (aov.rs:101-127). */
enum { AKR_AOV_SHADING_NORMAL = 0, AKR_AOV_GEOMETRY_NORMAL = 1, AKR_AOV_TANGENT = 2, AKR_AOV_BITANGENT = 3,
       AKR_AOV_ALBEDO = 4, AKR_AOV_ROUGHNESS = 5 };
typedef struct AkrAovConfig {
    uint32_t spp;                   /* default 256 */
    uint32_t aov;                   /* AKR_AOV_*, default shading normal */
    uint32_t remap;                 /* bool, default true */
    uint32_t _pad;
} AkrAovConfig;

/* Image-plane shard rendered by this context: rows [y0, y1) of the full sensor, or — when n_shards > 1 —
 * every n_shards-th block of `block_rows` rows of that range: row y belongs to shard ((y - y0) / block_rows) % n_shards.
 * Interleaving spreads expensive image regions over all GPUs (contiguous bands leave the band with the light source
 * 8 % more work on cbox).  The film of the context holds only its own rows, packed in increasing y.  The sampler is
 * keyed by absolute pixel coordinates, so any tiling yields the same image.  block_rows = 0 means 1.               */
typedef struct AkrTile {
    uint32_t y0, y1;
    uint32_t block_rows;            /* rows per interleaved block (ignored when n_shards <= 1)  */
    uint32_t n_shards;              /* 0 or 1: the contiguous band [y0, y1)                      */
    uint32_t shard;                 /* which of the n_shards interleaved sets, < n_shards        */
    uint32_t _pad;
} AkrTile;

typedef struct AkrStats {
    uint64_t samples;               /* camera paths started                                     */
    uint64_t segments;              /* closest-hit rays traced                                  */
    uint64_t shadow_rays;           /* any-hit rays traced                                      */
    uint64_t kernel_launches;       /* CUDA kernels launched by render calls since reset        */
    double gpu_ms;                  /* CUDA-event time of render calls since reset              */
    double gpu_ms_kernel[8];        /* per stage (profile_stages): 0 raygen, 1 trace (closest-hit + shadow rays),
                                     * 2 shade/Lambert, 3 shade/conductor, 4 accumulate, 5 misc,
                                     * 6 shade/general (or the unsorted shade kernel)                */
    uint64_t launches_kernel[8];
    uint64_t shaded_hits;           /* hits that went through a shade / bounce kernel (segments minus misses) */
} AkrStats;

typedef struct AkrContext AkrContext;

/* ---- entry points ---------------------------------------------------------------------------- */

/* Create a context on CUDA device `device_ordinal`.  Replaces `ctx.create_device("cuda")`
 * (crates/akari_api/src/bin/akari_cli.rs:56-62). */
int akr_b200_create(int device_ordinal, AkrContext **out_ctx);
void akr_b200_destroy(AkrContext *ctx);
const char *akr_b200_last_error(const AkrContext *ctx);   /* "" if none; valid until next call */

/* Use `stream` (a cudaStream_t) for all subsequent work of this context; NULL = default stream. */
int akr_b200_set_stream(AkrContext *ctx, void *stream);

/* Upload PMJ02BN samples [5][65536][2] u32 and blue-noise [48][128][128] u16.
 * Replaces Pmj02BnSamplerCreator::new's uploads (sampler/mod.rs:369-469). */
int akr_b200_upload_sampler_tables(AkrContext *ctx, const uint32_t *pmj02bn, const uint16_t *bluenoise);

/* Upload the 16x16x16 f32 `ggx_dielectric_s` directional-albedo table
 * (svm/surface/mod.rs:1196-1378; precompute.rs:56-94).  Optional: when never called the library
 * derives the table deterministically on the device at first use. */
int akr_b200_upload_albedo_table(AkrContext *ctx, const float *table_16x16x16);

/* Copy the scene to the device, fold shader constants, detect mesh lights and build their alias
 * tables (load.rs:312-444), build the BVH.  Replaces SceneLoader::do_load's device half
 * (load.rs:238-456) and MeshAggregate::new (mesh.rs:258-348). */
int akr_b200_upload_scene(AkrContext *ctx, const AkrSceneDesc *scene);

/* Start a render: allocate + clear the film (Film::new, film.rs:101-152) and the sampler state
 * (sampler/mod.rs:443-457) for `tile`.  tile == NULL renders the whole sensor. */
int akr_b200_begin(AkrContext *ctx, const AkrPtConfig *cfg, const AkrSamplerConfig *sampler,
                   const AkrFilterConfig *filter, const AkrTile *tile);

/* Render `n_spp` more samples per pixel into the film (one reference "pass",
 * pt.rs:1126-1149).  Blocking when `blocking != 0`. */
int akr_b200_render_pass(AkrContext *ctx, uint32_t n_spp, int blocking);

/* Convenience: begin + ceil(spp/spp_per_pass) passes, blocking.  == pt::render (pt.rs:1161-1172). */
int akr_b200_render_pt(AkrContext *ctx, const AkrPtConfig *cfg, const AkrSamplerConfig *sampler,
                       const AkrFilterConfig *filter, const AkrTile *tile);

/* The `aov` method (aov.rs:52-185): spp camera samples per pixel, the chosen first-hit quantity (shading / geometric
 * normal, tangent, bitangent, albedo + emission, lobe roughness) accumulated into the film like radiance.  Blocking. */
int akr_b200_render_aov(AkrContext *ctx, const AkrAovConfig *cfg, const AkrSamplerConfig *sampler,
                        const AkrFilterConfig *filter, const AkrTile *tile);

int akr_b200_synchronize(AkrContext *ctx);

/* Film in the reference layout (film.rs:66-76,88-93): f32[3N] sum rgb*w | f32[3N] splat | f32[N]
 * sum w, N = width * tile_rows, pixel index x + local_row * width with the tile's rows packed in increasing y
 * (local_row = y - y0 for a contiguous band).  Caller-allocated host buffer.
 * akr_b200_tile_rows: number of rows of `tile` on a sensor of any width (host helper, no device needed). */
uint32_t akr_b200_tile_rows(const AkrTile *tile);
int akr_b200_download_film(AkrContext *ctx, float *out_7n, size_t n_floats);

/* Film::copy_to_rgba_image(hdr) (film.rs:120-148): rgb/weight (+splat*scale); writes
 * [rows][width][3] (or [4] when `rgba != 0`, alpha = 1) f32 linear sRGB.
 * `*_device` writes to a device pointer (e.g. a torch tensor) on the context stream. */
int akr_b200_resolve_film(AkrContext *ctx, float *out_host, size_t n_floats, int rgba);
int akr_b200_resolve_film_device(AkrContext *ctx, void *out_device, size_t n_floats, int rgba);

int akr_b200_get_stats(AkrContext *ctx, AkrStats *out);
int akr_b200_reset_stats(AkrContext *ctx);

/* Tunables of the wavefront engine (not part of the reference surface). */
enum {
    AKR_AOV_FIRST_HIT_IDS = 1u << 0   /* (instance, primitive) of the first hit of sample 0 of every pixel */
};
typedef struct AkrEngineOptions {
    uint32_t wave_size;             /* paths in flight per wave; 0 = default                    */
    uint32_t _unused0;              /* (was sort_by_material; hits are always binned per shade class) */
    uint32_t profile_stages;        /* record per-stage CUDA-event times (adds syncs)           */
    uint32_t trace_mode;            /* 0 = auto, 1 = BVH traversal (persistent warps, dynamic ray fetch),
                                     * 2 = flat primitive list (only honoured when every primitive fits
                                     * in shared memory)                                                  */
    uint32_t fused;                 /* 0 = auto (flat scenes without alpha-tested materials run the fused
                                     * bounce kernels: shade + shadow ray + next ray in one kernel per depth
                                     * and shade class), 2 = off (always trace stage + shade stage + queues),
                                     * 3 = fused, every kernel on the context stream (no side stream; A/B runs) */
    uint32_t smem_node_kb;          /* BVH scenes that do not fit in shared memory: KiB of top-of-tree nodes each
                                     * CTA stages (0 = default 4); read by akr_b200_upload_scene               */
    uint32_t aov_mask;              /* AKR_AOV_* outputs to record; read by akr_b200_begin                    */
    uint32_t _reserved[1];
} AkrEngineOptions;
int akr_b200_set_engine_options(AkrContext *ctx, const AkrEngineOptions *opts);

/* AKR_AOV_FIRST_HIT_IDS: first-hit (inst, prim) per pixel of sample 0 (0xffffffff = miss); AKR_ERR_STATE when the
 * output was not requested before akr_b200_begin. */
int akr_b200_debug_first_hits(AkrContext *ctx, uint32_t *out_inst, uint32_t *out_prim, size_t n_pixels);

#ifdef __cplusplus
}
#endif
#endif /* AKARI_B200_H */

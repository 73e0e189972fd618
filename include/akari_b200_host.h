/*
 * akari_b200_host.h — host front-end that stands in for the reference's Rust host side.
 *
 * In the reference, scene loading and shader-graph flattening are Rust
 * (crates/akari_render/src/load.rs:63-72,238-456; crates/akari_render/src/svm/compiler.rs;
 * crates/akari_scenegraph/src/scene.rs:598-668) and the method file is serde JSON
 * (crates/akari_integrator/src/lib.rs:57-109).  No Rust toolchain exists in this image, so the
 * same duties are implemented here in C++ behind a C ABI: it reads the reference's on-disk
 * formats and produces the plain-array `AkrSceneDesc` that `akari_b200.h` consumes.  A Rust host
 * would skip this library and fill `AkrSceneDesc` from its own `Scene`.
 * Pure CPU code, no CUDA dependency.
 */
#ifndef AKARI_B200_HOST_H
#define AKARI_B200_HOST_H

#include "akari_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AkrHostScene AkrHostScene;

/* SceneLoader::load_from_path (load.rs:63-72): parse scene.json, map its buffers (path buffers
 * are resolved relative to the scene file; an unresolvable absolute path falls back to the same
 * file name next to scene.json), compile every material's shader graph, load camera + instances.
 * Returns AKR_OK or an error code; the message is available from akr_host_last_error(). */
int akr_host_load_scene(const char *scene_json_path, AkrHostScene **out_scene);
void akr_host_free_scene(AkrHostScene *scene);

/* Borrowed view; valid until the scene is freed or its resolution is changed. */
const AkrSceneDesc *akr_host_scene_desc(const AkrHostScene *scene);

/* Camera::set_resolution (camera/mod.rs:54-65) — BASELINE configs override the sensor size. */
int akr_host_scene_set_resolution(AkrHostScene *scene, uint32_t width, uint32_t height);

/* RenderTask / RenderConfig (lib.rs:57-109): one `{method:{type:"pt"|"aov",...}, sampler, film}` entry.
 * `pt` (pt.rs:916-944) and `aov` (aov.rs:23-36) are in scope; other methods return AKR_ERR_UNSUPPORTED. */
enum { AKR_METHOD_PT = 0, AKR_METHOD_AOV = 1 };
typedef struct AkrRenderTask {
    AkrPtConfig pt;
    AkrSamplerConfig sampler;
    AkrFilterConfig filter;
    char out[512];                  /* film.out */
    uint32_t method;                /* AKR_METHOD_* */
    AkrAovConfig aov;               /* valid when method == AKR_METHOD_AOV */
} AkrRenderTask;
int akr_host_parse_method_file(const char *method_json_path, AkrRenderTask *out_task);
int akr_host_parse_method_string(const char *method_json, AkrRenderTask *out_task);
void akr_host_default_task(AkrRenderTask *out_task);   /* pt::Config::default, pt.rs:929-944 */

/* util::write_image (util/mod.rs:57-127): `.exr` = linear RGB f32, uncompressed scanlines (write_image_hdr); `.png` =
 * write_image_ldr: linear -> sRGB, (x * 255).clamp(0, 255) as u8, 8-bit RGB.  `.pfm` is also accepted.
 * rgb = [height][width][3]. */
int akr_host_write_image(const char *path, const float *rgb, uint32_t width, uint32_t height);

const char *akr_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* AKARI_B200_HOST_H */

// scene_build.h — host-side preparation of the device scene blob from an AkrSceneDesc.
//
// Does, on the CPU and once per upload, what the reference's loader does with device kernels and
// third-party calls (crates/akari_render/src/load.rs:238-456, mesh.rs:258-348):
//   * constant-folds every material's SVM program into a Material record (svm/eval.rs:97-269);
//   * precomputes the barycentric-independent part of surface_interaction per triangle (mesh.rs:487-654);
//   * detects mesh lights and builds their alias tables (load.rs:312-444, util/distribution.rs:34-78);
//   * builds the BVH that replaces rtx::Accel (mesh.rs:266,288-294,331-333).
#pragma once
#include "../../../include/akari_b200.h"
#include "../device/akr_scene.cuh"

#include <string>
#include <vector>

namespace akr {

constexpr uint32_t kFlatMaxPrims = 64;  // scenes this small are traced as one flat primitive list

struct HostSceneBlob {
    std::vector<BvhNode> nodes;
    std::vector<Bvh4Node> nodes4;     // the BVH collapsed to 4-wide nodes, breadth-first
    uint32_t bvh4_depth = 0;
    std::vector<PrimRec> prims;       // BVH leaf order (what the CUDA kernels intersect)
    std::vector<PrimBlock2> flat_blocks;  // scenes of <= kFlatMaxPrims primitives: pairs first, two per block (flat trace mode)
    uint32_t n_pair_blocks = 0, n_single_blocks = 0;          // complete list, then ...
    uint32_t n_occ_pair_blocks = 0, n_occ_single_blocks = 0;  // ... the occluder-only list (any-hit rays)
    uint32_t n_shadow_pair_blocks = 0, n_shadow_single_blocks = 0;  // leading blocks of each group of the complete list that hold occluders
    std::vector<TriGeom> tris;        // two per primitive, same order (host simulation's Moeller-Trumbore path)
    std::vector<TriShade> shade;      // by global triangle id
    std::vector<InstanceRec> instances;
    std::vector<Material> materials;
    std::vector<LightRec> lights;
    std::vector<uint32_t> alias_j;
    std::vector<float> alias_t, alias_pdf;
    std::vector<float> corner_normals;   // [n_tris * 9] or empty
    std::vector<float> corner_tangents;  // [n_tris * 9] or empty
    std::vector<float> corner_uvs;       // [n_tris * 6] or empty (only when a texture-driven material exists)
    // shader virtual machine tables (akr_svm.cuh): programs of every kind, constant blob, textures
    std::vector<AkrSvmNode> svm_nodes;
    std::vector<uint32_t> svm_kind_first;
    std::vector<uint8_t> svm_data;
    std::vector<uint64_t> svm_kind_hit_mask;  // [n_kinds] hit-dependent nodes of each program (SvmView::kind_hit_mask)
    std::vector<SvmVal> svm_static_vals;      // per texture-driven material: value of every node at upload (Material.static_offset)
    std::vector<TextureRec> textures;    // .texels = byte offset into `texels` (patched to a pointer by whoever owns the copy)
    std::vector<uint8_t> texels;
    std::vector<TextureRec> textures_host;  // the same records with .texels pointing into `texels` (host-side evaluation)
    uint32_t any_dynamic = 0;            // some material is texture-driven
    CameraRec camera{};
    uint32_t any_alpha = 0;
    uint32_t bvh_depth = 0;
    // per-light-instance total power, for diagnostics / tests
    std::vector<float> light_powers;
};

// Returns AKR_OK or an AKR_ERR_* code with a message in `err`.
int build_scene_blob(const AkrSceneDesc &desc, HostSceneBlob &out, std::string &err);

// [48][128][128] u16 table, reference layout [t][px % 128][py % 128] -> device layout [t][py % 128][px % 128]
void transpose_bluenoise(const uint16_t *src, uint16_t *dst);

// Fills a SceneView whose pointers alias the host vectors (used by the host-side kernel simulation in
// tests; the CUDA library fills a SceneView with device pointers instead).
SceneView host_scene_view(const HostSceneBlob &blob, const float *albedo_table);

// Deterministic derivation of the 16^3 `ggx_dielectric_s` directional-albedo table
// (svm/surface/precompute.rs:56-94, svm/surface/mod.rs:1338-1356): mean of f / pdf over an n x n
// midpoint grid of the 2-D sample instead of 2^20 PCG32 draws seeded from rand::StdRng.
void make_albedo_table(float *table_16x16x16, uint32_t n);

}  // namespace akr

// akr_scene.cuh — flat, pointer-based scene layout consumed by the kernels ("scene blob" in HBM).
//
// Replaces what the reference keeps in luisa buffers / bindless heap / rtx::Accel:
//   MeshAggregate + MeshHeader + MeshInstance   crates/akari_render/src/mesh.rs:189-348
//   LightAggregate + AliasTable                  light/mod.rs:87-92, util/distribution.rs:12-33
//   PerspectiveCameraData                        camera/mod.rs:108-153
//   Svm shader data                              svm/mod.rs:240-255  (constant-folded into Material[])
// Everything is read-only during rendering and small enough on `cbox` (a few KB) to be staged in
// shared memory by a TMA bulk copy at kernel start; larger scenes keep the top of the BVH in shared
// memory and the rest in L2.
#pragma once
#include "akr_svm.cuh"

namespace akr {

// BVH2 node holding both children's boxes (one 64-byte fetch decides both children).
struct alignas(16) BvhNode {
    float lo0[3], hi0[3];
    float lo1[3], hi1[3];
    int32_t c0, c1;  // >= 0: inner node index; < 0: leaf, ~c = (first_tri << 3) | count
    uint32_t _pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

// 4-wide node (the BVH2 collapsed two levels at a time): fewer, fatter steps per ray — half the dependent node
// fetches and loop trips of the binary tree, and the slab tests of two children can ride on one packed FFMA2.
// STATUS: built and validated on the host (tests: test_bvh4_collapse_bitwise) but NOT yet walked by the kernels — the
// first device version (dynamic-fetch kernel over this tree) measured slower than the binary tree on B200 (142 vs 76 ms
// per pass on the clutter scene: the per-lane child arrays went to local memory, the 3x deeper stack cost occupancy),
// so k_trace_bvh keeps the binary tree for now.
// lo[axis][child] / hi[axis][child]: children (0,1) and (2,3) are adjacent pairs.  Empty slots: lo = +inf, hi = -inf,
// c = ~0 (a leaf of zero primitives).  Child codes as in BvhNode, inner indices refer to the Bvh4Node array.
struct alignas(16) Bvh4Node {
    float lo[3][4];
    float hi[3][4];
    int32_t c[4];
    uint32_t _pad[4];
};
static_assert(sizeof(Bvh4Node) == 128, "Bvh4Node must be 128 bytes");

// World-space triangle for the Moeller-Trumbore traversal, stored in primitive order (2 per PrimRec).  48 bytes.
struct alignas(16) TriGeom {
    float v0[3];
    uint32_t gid;  // global triangle id = instance.tri_offset + prim (instances in id order)
    float e1[3];
    uint32_t cls;  // ShadeClass of the triangle's material: the trace stage bins hits by it
    float e2[3];
    uint32_t _p1;
};
static_assert(sizeof(TriGeom) == 48, "TriGeom must be 48 bytes");

// Traversal primitive of the CUDA kernels: one triangle, or two triangles that share an edge and form a
// parallelogram (both halves of a quad are decided by ONE plane + two-coordinate test).  The record
// holds the plane and the two rows of the world -> (s, q) affine map (Baldwin & Weber's precomputed
// transformation): for a hit point P,  s = r0 . P + r0.w,  q = r1 . P + r1.w.
//   single   : (s, q) are the barycentrics (u, v) of the triangle; inside  <=>  s, q >= 0, s + q <= 1
//   pair     : P = p0 + s a + q b over the parallelogram (p0, p1 = p0 + a, p2 = p0 + a + b, p3 = p0 + b);
//              inside <=> 0 <= s, q <= 1;  s >= q is triangle A = (p0, p1, p2), else B = (p0, p2, p3);
//              weights A: (1 - s, s - q, q), B: (1 - q, s, q - s) for (p0, p1|p2, p2|p3); `meta` says which
//              weight is the triangle's own u (of v1) and v (of v2).
struct alignas(16) PrimRec {
    float n[4];
    float r0[4];
    float r1[4];
    uint32_t gid_a, gid_b;  // gid_b == 0xffffffff: single triangle
    uint32_t meta;          // bits 0-1 iu_a, 2-3 iv_a, 4-5 iu_b, 6-7 iv_b, 8-9 cls_a, 10-11 cls_b, 12 light_a, 13 light_b
    uint32_t _pad;
};
static_assert(sizeof(PrimRec) == 64, "PrimRec must be 64 bytes");

// Flat trace mode (small scenes): the primitives transposed two by two so that one packed-FP32 instruction
// (Blackwell FFMA2: fma.rn.f32x2) advances the test of two primitives.  Pair primitives come first, then single
// triangles; each group is padded to an even count with a never-hit (NaN) primitive.
struct alignas(16) PrimBlock2 {  // 128 B
    float n[4][2];    // plane (nx, ny, nz, nw) of primitive 0 / 1
    float r0[4][2];
    float r1[4][2];
    uint32_t gid[4];  // gid_a[0], gid_b[0], gid_a[1], gid_b[1]
    uint32_t meta[2];
    uint32_t _pad[2];
};
static_assert(sizeof(PrimBlock2) == 128, "PrimBlock2 must be 128 bytes");

enum TriFlags : uint32_t {
    TRI_IS_LIGHT = 1u << 0,     // instance.light.valid()
    TRI_HAS_NORMALS = 1u << 1,  // per-corner normals: ns interpolated per hit
    TRI_HAS_TANGENTS = 1u << 2,
    TRI_HAS_UVS = 1u << 3,
    TRI_ALPHA = 1u << 4,        // material alpha < 1: stochastic alpha test applies
    // bits 8-12: sort key of the material's evaluation signature (type, lobe set, shader kind of texture-driven
    // materials), numbered by increasing cost.  The general shade class orders each CTA tile of records by it so that
    // a warp evaluates one kind of tree instead of the union of all of them (scene_build.cpp material_sort_keys).
    TRI_SORT_KEY_SHIFT = 8,
    TRI_SORT_KEY_MASK = 31u,
};

// Shading record per global triangle id: the bary-independent part of
// MeshAggregate::surface_interaction (mesh.rs:487-654), precomputed at upload with the same f32
// operations.  For flat triangles (no per-corner normals) the whole frame is constant.  96 bytes.
struct alignas(16) TriShade {
    float v0[3], v1[3], v2[3];  // local-space vertices (p = M * interp + t is evaluated per hit)
    float ng[3];                // world geometric normal  normalize((M^T)^-1 ng_local)
    float ft[3], fs[3];         // frame.t, frame.s for flat triangles (frame.n = ng)
    float area;                 // world prim_area
    float prim_pdf;             // area_sampler.pdf(prim) when the instance is a light
    uint32_t inst;
    uint32_t mat;               // index into Material[]
    uint32_t flags;             // TriFlags
    uint32_t prim;
};
static_assert(sizeof(TriShade) == 96, "TriShade must be 96 bytes");

struct alignas(16) InstanceRec {
    float m[9];        // upper 3x3, column-major
    float t[3];        // translation
    float m_inv_t[9];  // (M^T)^-1, column-major
    float det;         // MeshInstance.transform_det (mesh.rs:311-312)
    uint32_t light_id; // 0xffffffff = not a light
    uint32_t tri_offset;
    uint32_t n_tris;
    uint32_t geom_id;
    uint32_t flags;
    uint32_t _pad[5];
};
static_assert(sizeof(InstanceRec) == 128, "InstanceRec must be 128 bytes");

struct LightRec {  // AreaLight (light/area.rs:12-16) + its per-instance alias table
    uint32_t inst;
    uint32_t tri_offset;    // gid of prim 0
    uint32_t alias_offset;  // into alias_j / alias_t / alias_pdf
    uint32_t n_prims;
};

struct CameraRec {
    float c2w[12];   // columns 0..2 (xyz) then translation
    uint32_t c2w_identity;  // AffineTransform.close_to_identity (geometry.rs:212-218)
    float r2c_s[3], r2c_t[3];  // raster->camera is scale+translate (camera/mod.rs:129-145)
    uint32_t width, height;
};

struct SceneView {
    const BvhNode *nodes;      // leaves address primitives: ~c = (first_prim << 3) | count
    const Bvh4Node *nodes4;    // the same tree, 4-wide (host simulation only for now; nullptr on the device)
    uint32_t n_nodes4, bvh4_depth;
    const PrimRec *prims;      // BVH leaf order (CUDA kernels)
    const PrimBlock2 *flat_blocks;  // small scenes only (flat trace mode), else nullptr: [ every primitive | occluders only ]
    uint32_t n_pair_blocks, n_single_blocks;          // the complete list (closest-hit rays)
    uint32_t n_occ_pair_blocks, n_occ_single_blocks;  // the occluder list behind it (any-hit rays): primitives whose plane
                                                      // supports the whole scene (convex-hull walls) cannot block a segment
                                                      // between two points of the scene and are left out
    uint32_t n_shadow_pair_blocks, n_shadow_single_blocks;  // the complete list keeps occluders first in each group: a shadow ray
                                                            // only needs these leading blocks (dual trace of the fused kernels)
    const TriGeom *tris;       // two slots per primitive, gid 0xffffffff = empty (Moeller-Trumbore path of the host simulation)
    const TriShade *shade;
    const InstanceRec *instances;
    const Material *materials;
    const LightRec *lights;
    const uint32_t *alias_j;   // light distribution first (n_lights entries), then per-light tables
    const float *alias_t;
    const float *alias_pdf;
    const float *albedo_table; // 16^3
    SvmView svm;               // shader programs + constants + textures (texture-driven materials and alpha tests only)
    const float *corner_uvs;   // [n_tris * 3][2] per-corner texture coordinates, or nullptr when no material needs them
    uint32_t n_nodes, n_prims, n_tris, n_instances, n_materials, n_lights;
    uint32_t any_alpha;        // some material has alpha < 1
    CameraRec camera;
};

struct SamplerTables {
    const uint32_t *pmj;   // [5][65536][2]
    const uint16_t *bn;    // [48][128][128], transposed at upload to [t][py % 128][px % 128]
};

struct RenderParams {  // pt::Config + sampler + filter, resolved for one pass
    uint32_t spp_total;      // Pmj02BnState.spp
    uint32_t w_mask;         // Pmj02BnState.w
    uint32_t seed;
    uint32_t max_depth, rr_depth;
    uint32_t use_nee, indirect_only, force_diffuse;
    int32_t pixel_offset_x, pixel_offset_y;
    int32_t debug_depth;     // < 0: none
    uint32_t filter_type;
    float filter_radius;
    uint32_t width, height;  // full sensor
    uint32_t y0;             // first row of this context's tile
    uint32_t tile_block, tile_shards, tile_shard;  // interleaved tile (AkrTile): local row -> y0 + ((row / block) * shards + shard) * block + row % block
    // derived (finish_render_params)
    FastDiv width_div;       // pixel -> (row, column)
    FastDiv tile_block_div;
    uint32_t spp_pow2;       // spp_total is a power of two: x / spp == x * inv_spp exactly
    float inv_spp;
};
inline void finish_render_params(RenderParams &rp) {
    rp.width_div = make_fastdiv(rp.width);
    if (rp.tile_block == 0u) rp.tile_block = 1u;
    if (rp.tile_shards == 0u) rp.tile_shards = 1u;
    rp.tile_block_div = make_fastdiv(rp.tile_block);
    rp.spp_pow2 = (rp.spp_total & (rp.spp_total - 1u)) == 0u ? 1u : 0u;
    rp.inv_spp = 1.0f / (float)rp.spp_total;
}

// rows of [y0, y1) that belong to one interleaved shard (AkrTile semantics)
inline uint32_t interleaved_tile_rows(uint32_t y0, uint32_t y1, uint32_t block, uint32_t shards, uint32_t shard) {
    if (y1 <= y0) return 0u;
    if (shards <= 1u) return y1 - y0;
    if (block == 0u) block = 1u;
    const uint32_t span = y1 - y0, n_blocks = (span + block - 1u) / block;
    uint32_t rows = 0u;
    for (uint32_t b = shard; b < n_blocks; b += shards) rows += (b + 1u) * block <= span ? block : span - b * block;
    return rows;
}

}  // namespace akr

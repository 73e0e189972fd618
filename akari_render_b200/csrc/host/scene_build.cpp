// scene_build.cpp — see scene_build.h.  Compiled with -ffp-contract=off: every f32 operation here is
// one IEEE operation, in the order the cited reference code performs it.
#include "scene_build.h"
#include "../device/akr_trace.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <map>

namespace akr {

namespace {

// ---------------------------------------------------------------------------------------------
// SVM: every material's node program is evaluated once here with the evaluator the kernels use (akr_svm.cuh).  A
// program without hit-dependent nodes folds to constants; a texture-driven one keeps its ShaderRef for per-hit evaluation.
// ---------------------------------------------------------------------------------------------
// Can the alpha of shader (kind, data_offset) differ from hit to hit?  Alpha is the fourth component of the colour that
// feeds a Diffuse / Principled closure (diffuse.rs:85-92, principled.rs:15-22); it is carried by image textures
// (eval.rs:137-157) through spectral uplift (eval.rs:158-180) and the checkerboard.  A texture contributes a varying alpha
// when one of its texels has alpha != 1 or when its address mode returns zeros outside [0, 1)^2 (bilinear weights
// (1 - t) + t of an all-ones alpha round to exactly 1, so opaque textures stay exactly opaque).  Everything else has a
// constant alpha of exactly 1 and the material's triangles need no stochastic alpha test at all.
// Called after svm_eval<., true> has validated the program; `svm.textures` are the host copies.
bool alpha_varies(const SvmView &svm, AkrShaderRef ref) {
    const uint32_t first = svm.kind_first[ref.shader_kind], n_nodes = svm.kind_first[ref.shader_kind + 1u] - first;
    bool av[kSvmMaxNodes] = {};
    bool varies = false;
    for (uint32_t i = 0; i < n_nodes; ++i) {
        const AkrSvmNode &n = svm.nodes[first + i];
        switch (n.op) {
        case AKR_SVM_RGB_IMAGE_TEX: {
            uint32_t tex;
            std::memcpy(&tex, svm.data + ref.data_offset + n.a[0], 4);
            const TextureRec &t = svm.textures[tex];
            bool v = t.address != AKR_ADDRESS_REPEAT && t.address != AKR_ADDRESS_MIRROR && t.address != AKR_ADDRESS_EDGE;
            const size_t n_texels = static_cast<size_t>(t.width) * t.height;
            for (size_t k = 0; k < n_texels && !v; ++k)
                v = t.texel_format == AKR_TEXEL_RGBA8 ? static_cast<const uint8_t *>(t.texels)[4 * k + 3] != 255u
                                                      : static_cast<const float *>(t.texels)[4 * k + 3] != 1.0f;
            av[i] = v;
            break;
        }
        case AKR_SVM_SPECTRAL_UPLIFT: av[i] = av[n.a[0]]; break;
        case AKR_SVM_CHECKERBOARD: av[i] = av[n.a[2]] || av[n.a[3]]; break;
        case AKR_SVM_DIFFUSE_BSDF: varies |= av[n.a[0]]; break;
        case AKR_SVM_PRINCIPLED_BSDF: varies |= av[n.a[AKR_P_BASE_COLOR]]; break;
        default: break;  // scalars, vectors, rgb constants (alpha 1), glass and emission (alpha stays 1)
        }
    }
    return varies;
}

// Sort keys for the general shade class (akr_scene.cuh TRI_SORT_KEY_*): materials with the same evaluation signature —
// closure type, lobe set, normal-map frame, and for texture-driven materials the shader kind (their lobe set is only
// known per hit) — share a key; keys are numbered by increasing cost estimate so that a CTA can pair its cheapest run
// of a sorted tile with its most expensive one.  More than 32 signatures fold onto the last keys (a warp then mixes
// some neighbouring signatures: slower, not wrong).
std::vector<uint32_t> material_sort_keys(const std::vector<Material> &mats) {
    auto signature = [](const Material &m) -> uint64_t {
        if (m.dynamic) return (1ull << 40) | m.shader_kind;
        return (static_cast<uint64_t>(m.type) << 16) | (static_cast<uint64_t>(m.lobes & 0xffu) << 4) | (m.has_normal ? 2u : 0u) | (m.wrap_inner ? 1u : 0u);
    };
    auto cost = [](const Material &m) -> uint32_t {
        if (m.dynamic) return 100u + m.shader_kind % 16u;  // shader interpretation on top of an unknown tree: the most expensive
        uint32_t c = 0;
        if (m.type == MAT_CONDUCTOR) c = 6;
        else if (m.type == MAT_GLASS) c = 8;
        else if (m.type == MAT_PRINCIPLED)
            c = 1u * !!(m.lobes & LOBE_DIFFUSE) + 8u * !!(m.lobes & LOBE_TRANSMISSION) + 6u * !!(m.lobes & LOBE_METAL) + 5u * !!(m.lobes & LOBE_SPECULAR) +
                5u * !!(m.lobes & LOBE_COAT);
        return c;
    };
    std::map<uint64_t, uint32_t> cost_of;  // signature -> cost
    for (const Material &m : mats) cost_of[signature(m)] = cost(m);
    std::vector<std::pair<uint32_t, uint64_t>> order;
    for (const auto &kv : cost_of) order.emplace_back(kv.second, kv.first);
    std::sort(order.begin(), order.end());
    std::map<uint64_t, uint32_t> key_of;
    for (size_t i = 0; i < order.size(); ++i) key_of[order[i].second] = static_cast<uint32_t>(std::min<size_t>(i, TRI_SORT_KEY_MASK));
    std::vector<uint32_t> keys(mats.size());
    for (size_t i = 0; i < mats.size(); ++i) keys[i] = key_of[signature(mats[i])];
    return keys;
}

int fold_material(const SvmView &svm, AkrShaderRef ref, Material &m, std::vector<uint64_t> &kind_hit_mask, std::vector<SvmVal> &static_vals, std::string &err) {
    std::memset(&m, 0, sizeof(m));
    bool dynamic = false;
    SvmFoldInfo fold;
    int rc = svm_eval<false, true>(svm, ref.shader_kind, ref.data_offset, f2{0.0f, 0.0f}, m, &dynamic, AKR_SVM_NONE, &fold);
    if (rc == SVM_BAD_PROGRAM) {
        err = "malformed shader program (kind / constant offset / node reference / input type out of range)";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    if (rc != SVM_OK) {
        err = "shader uses an SVM op or colour space outside the implemented scope, or has more than 64 nodes";
        return AKR_ERR_UNSUPPORTED;
    }
    m.dynamic = dynamic ? 1u : 0u;
    m.alpha_dynamic = (dynamic && alpha_varies(svm, ref)) ? 1u : 0u;
    m.static_offset = AKR_SVM_NONE;
    kind_hit_mask[ref.shader_kind] = fold.hit_mask;  // a property of the program: the same for every material of the kind
    if (dynamic) {  // the per-hit evaluation reads the hit-independent nodes' values from here
        const uint32_t n_nodes = svm.kind_first[ref.shader_kind + 1u] - svm.kind_first[ref.shader_kind];
        m.static_offset = static_cast<uint32_t>(static_vals.size());
        static_vals.insert(static_vals.end(), fold.vals, fold.vals + n_nodes);
    }
    m.shader_kind = ref.shader_kind;
    m.data_offset = ref.data_offset;
    return AKR_OK;
}

// ---------------------------------------------------------------------------------------------
// small linear algebra with explicit operation order (mirrors the device helpers)
// ---------------------------------------------------------------------------------------------
struct H3 {
    float x, y, z;
};
H3 operator+(H3 a, H3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
H3 operator-(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
H3 operator*(H3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
H3 operator/(H3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
float hdot(H3 a, H3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
H3 hcross(H3 a, H3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float hlen(H3 a) { return std::sqrt(hdot(a, a)); }
H3 hnorm(H3 a) { return a * (1.0f / std::sqrt(hdot(a, a))); }
struct HM3 {
    H3 c0, c1, c2;
};
H3 hmul(const HM3 &m, H3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
HM3 htranspose(const HM3 &m) { return {{m.c0.x, m.c1.x, m.c2.x}, {m.c0.y, m.c1.y, m.c2.y}, {m.c0.z, m.c1.z, m.c2.z}}; }
HM3 hinverse(const HM3 &m) {  // adjugate / determinant
    H3 a = m.c0, b = m.c1, c = m.c2;
    H3 r0 = hcross(b, c), r1 = hcross(c, a), r2 = hcross(a, b);
    float inv_det = 1.0f / hdot(r2, c);
    return {H3{r0.x, r1.x, r2.x} * inv_det, H3{r0.y, r1.y, r2.y} * inv_det, H3{r0.z, r1.z, r2.z} * inv_det};
}
float mat4_det(const float *m) {  // glam Mat4::determinant, scalar form (mesh.rs:311-312)
    float m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3];
    float m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    float m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11];
    float m30 = m[12], m31 = m[13], m32 = m[14], m33 = m[15];
    float a2323 = m22 * m33 - m23 * m32;
    float a1323 = m21 * m33 - m23 * m31;
    float a1223 = m21 * m32 - m22 * m31;
    float a0323 = m20 * m33 - m23 * m30;
    float a0223 = m20 * m32 - m22 * m30;
    float a0123 = m20 * m31 - m21 * m30;
    return m00 * (m11 * a2323 - m12 * a1323 + m13 * a1223) - m01 * (m10 * a2323 - m12 * a0323 + m13 * a0223) +
           m02 * (m10 * a1323 - m11 * a0323 + m13 * a0123) - m03 * (m10 * a1223 - m11 * a0223 + m12 * a0123);
}
void st3(float *dst, H3 v) {
    dst[0] = v.x;
    dst[1] = v.y;
    dst[2] = v.z;
}

// Vose alias table exactly as util/distribution.rs:34-78
void build_alias(const std::vector<float> &weights, std::vector<uint32_t> &j, std::vector<float> &t, std::vector<float> &pdf) {
    size_t n = weights.size();
    float sum = 0.0f;
    for (float x : weights) sum += x;
    std::vector<float> prob(n);
    for (size_t i = 0; i < n; ++i) prob[i] = weights[i] / sum * static_cast<float>(n);
    std::deque<size_t> small, large;
    for (size_t i = 0; i < n; ++i) (prob[i] >= 1.0f ? large : small).push_back(i);
    j.assign(n, 0);
    t.assign(n, 0.0f);
    while (!small.empty() && !large.empty()) {
        size_t l = small.front();
        small.pop_front();
        size_t g = large.front();
        large.pop_front();
        t[l] = prob[l];
        j[l] = static_cast<uint32_t>(g);
        prob[g] = (prob[g] + prob[l]) - 1.0f;
        (prob[g] < 1.0f ? small : large).push_back(g);
    }
    while (!large.empty()) {
        size_t g = large.front();
        large.pop_front();
        t[g] = 1.0f;
        j[g] = static_cast<uint32_t>(g);
    }
    while (!small.empty()) {
        size_t l = small.front();
        small.pop_front();
        t[l] = 1.0f;
        j[l] = static_cast<uint32_t>(l);
    }
    pdf.resize(n);
    for (size_t i = 0; i < n; ++i) pdf[i] = weights[i] / sum;
}

// ---------------------------------------------------------------------------------------------
// BVH: binned SAH, leaves of <= 4 triangles, flattened breadth-first into BvhNode[]
// ---------------------------------------------------------------------------------------------
struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::numeric_limits<float>::infinity();
            hi[a] = -std::numeric_limits<float>::infinity();
        }
    }
    void grow(const float *p) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    void grow(const Box &b) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
        return dx * dy + dy * dz + dz * dx;
    }
};
struct BuildTri {
    Box box;
    float centroid[3];
    uint32_t index;  // into the world-triangle list (gid)
};
struct BuildNode {
    Box box;
    int32_t left = -1, right = -1;  // children (build-node indices) or leaf range
    uint32_t first = 0, count = 0;
};

constexpr uint32_t kLeafMax = 4;
constexpr int kBins = 16;

int32_t build_recursive(std::vector<BuildNode> &nodes, std::vector<BuildTri> &tris, uint32_t first, uint32_t count, uint32_t depth,
                        uint32_t &max_depth) {
    BuildNode node;
    node.box.reset();
    Box cbox;
    cbox.reset();
    for (uint32_t i = first; i < first + count; ++i) {
        node.box.grow(tris[i].box);
        cbox.grow(tris[i].centroid);
    }
    max_depth = std::max(max_depth, depth);
    int32_t self = static_cast<int32_t>(nodes.size());
    nodes.push_back(node);
    if (count <= kLeafMax) {
        nodes[self].first = first;
        nodes[self].count = count;
        return self;
    }
    // pick the best binned SAH split over the three axes
    float best_cost = std::numeric_limits<float>::infinity();
    int best_axis = -1, best_bin = -1;
    for (int axis = 0; axis < 3; ++axis) {
        float lo = cbox.lo[axis], hi = cbox.hi[axis];
        if (!(hi > lo)) continue;
        Box bin_box[kBins];
        uint32_t bin_cnt[kBins] = {0};
        for (auto &b : bin_box) b.reset();
        float scale = static_cast<float>(kBins) / (hi - lo);
        for (uint32_t i = first; i < first + count; ++i) {
            int b = std::min(kBins - 1, std::max(0, static_cast<int>((tris[i].centroid[axis] - lo) * scale)));
            bin_box[b].grow(tris[i].box);
            bin_cnt[b]++;
        }
        float right_area[kBins];
        uint32_t right_cnt[kBins];
        Box acc;
        acc.reset();
        uint32_t cnt = 0;
        for (int b = kBins - 1; b > 0; --b) {
            acc.grow(bin_box[b]);
            cnt += bin_cnt[b];
            right_area[b] = acc.half_area();
            right_cnt[b] = cnt;
        }
        acc.reset();
        cnt = 0;
        for (int b = 0; b < kBins - 1; ++b) {
            acc.grow(bin_box[b]);
            cnt += bin_cnt[b];
            if (cnt == 0 || right_cnt[b + 1] == 0) continue;
            float cost = acc.half_area() * static_cast<float>(cnt) + right_area[b + 1] * static_cast<float>(right_cnt[b + 1]);
            if (cost < best_cost) {
                best_cost = cost;
                best_axis = axis;
                best_bin = b;
            }
        }
    }
    uint32_t mid;
    if (best_axis < 0) {
        mid = first + count / 2;  // all centroids coincide: split by index
    } else {
        float lo = cbox.lo[best_axis], hi = cbox.hi[best_axis];
        float scale = static_cast<float>(kBins) / (hi - lo);
        auto it = std::stable_partition(tris.begin() + first, tris.begin() + first + count, [&](const BuildTri &t) {
            int b = std::min(kBins - 1, std::max(0, static_cast<int>((t.centroid[best_axis] - lo) * scale)));
            return b <= best_bin;
        });
        mid = static_cast<uint32_t>(it - tris.begin());
        if (mid == first || mid == first + count) mid = first + count / 2;
    }
    int32_t l = build_recursive(nodes, tris, first, mid - first, depth + 1, max_depth);
    int32_t r = build_recursive(nodes, tris, mid, first + count - mid, depth + 1, max_depth);
    nodes[self].left = l;
    nodes[self].right = r;
    return self;
}

}  // namespace

void transpose_bluenoise(const uint16_t *src, uint16_t *dst) {
    for (uint32_t t = 0; t < 48; ++t)
        for (uint32_t x = 0; x < 128; ++x)
            for (uint32_t y = 0; y < 128; ++y) dst[(t * 128u + y) * 128u + x] = src[(t * 128u + x) * 128u + y];
}

int build_scene_blob(const AkrSceneDesc &d, HostSceneBlob &out, std::string &err) {
    if (d.abi_version != AKR_B200_ABI_VERSION) {
        err = "AkrSceneDesc.abi_version mismatch";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    if (d.n_instances == 0 || d.n_meshes == 0 || !d.meshes || !d.instances || !d.shader_kinds || !d.shader_data) {
        err = "empty scene";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    out = HostSceneBlob{};
    // ---- shader virtual machine tables ----
    out.svm_kind_first.push_back(0u);
    for (uint32_t k = 0; k < d.n_shader_kinds; ++k) {
        const AkrShaderKind &kind = d.shader_kinds[k];
        if (kind.n_nodes && !kind.nodes) {
            err = "shader kind without nodes";
            return AKR_ERR_INVALID_ARGUMENT;
        }
        out.svm_nodes.insert(out.svm_nodes.end(), kind.nodes, kind.nodes + kind.n_nodes);
        out.svm_kind_first.push_back(static_cast<uint32_t>(out.svm_nodes.size()));
    }
    out.svm_data.assign(d.shader_data, d.shader_data + d.shader_data_size);
    if (d.n_images && !d.images) {
        err = "n_images > 0 without an image array";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    for (uint32_t t = 0; t < d.n_images; ++t) {
        const AkrImage &img = d.images[t];
        if (!img.texels || img.width == 0 || img.height == 0 || img.texel_format > AKR_TEXEL_RGBA32F || img.address > AKR_ADDRESS_EDGE || img.filter > AKR_FILTER_LINEAR) {
            err = "bad image";
            return AKR_ERR_INVALID_ARGUMENT;
        }
        while (out.texels.size() % 16) out.texels.push_back(0);
        TextureRec tr{};
        tr.texels = reinterpret_cast<const void *>(out.texels.size());  // byte offset, patched below / by the uploader
        tr.width = img.width;
        tr.height = img.height;
        tr.texel_format = img.texel_format;
        tr.address = img.address;
        tr.filter = img.filter;
        const size_t bytes = static_cast<size_t>(img.width) * img.height * (img.texel_format == AKR_TEXEL_RGBA8 ? 4 : 16);
        const uint8_t *src = static_cast<const uint8_t *>(img.texels);
        out.texels.insert(out.texels.end(), src, src + bytes);
        out.textures.push_back(tr);
    }
    std::vector<TextureRec> host_textures = out.textures;
    for (TextureRec &tr : host_textures) tr.texels = out.texels.data() + reinterpret_cast<size_t>(tr.texels);
    SvmView svm{};
    svm.nodes = out.svm_nodes.data();
    svm.kind_first = out.svm_kind_first.data();
    svm.data = out.svm_data.data();
    svm.textures = host_textures.data();
    svm.n_kinds = d.n_shader_kinds;
    svm.n_textures = d.n_images;
    svm.data_size = static_cast<uint32_t>(out.svm_data.size());
    svm.kind_hit_mask = nullptr;  // (the validating evaluation interprets every node)
    svm.static_vals = nullptr;
    out.svm_kind_hit_mask.assign(d.n_shader_kinds, 0ull);
    // ---- materials: one record per distinct ShaderRef ----
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> mat_index;
    auto material_of = [&](AkrShaderRef ref, uint32_t &idx) -> int {
        auto key = std::make_pair(ref.shader_kind, ref.data_offset);
        auto it = mat_index.find(key);
        if (it != mat_index.end()) {
            idx = it->second;
            return AKR_OK;
        }
        Material m;
        int rc = fold_material(svm, ref, m, out.svm_kind_hit_mask, out.svm_static_vals, err);
        if (rc == AKR_OK && m.dynamic) out.any_dynamic = 1;
        if (rc != AKR_OK) return rc;
        idx = static_cast<uint32_t>(out.materials.size());
        out.materials.push_back(m);
        mat_index[key] = idx;
        return AKR_OK;
    };

    // ---- instances + per-triangle records ----
    bool any_normals = false, any_tangents = false;
    uint32_t total_tris = 0;
    for (uint32_t i = 0; i < d.n_instances; ++i) {
        if (d.instances[i].geom_id >= d.n_meshes) {
            err = "instance geom_id out of range";
            return AKR_ERR_INVALID_ARGUMENT;
        }
        const AkrMesh &g = d.meshes[d.instances[i].geom_id];
        total_tris += g.n_triangles;
        any_normals |= g.normals != nullptr;
        any_tangents |= g.tangents != nullptr;
    }
    if (total_tris == 0 || total_tris >= (1u << 28)) {
        err = "triangle count out of range";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    out.shade.resize(total_tris);
    for (uint32_t i = 0; i < d.n_instances; ++i)  // every material is evaluated up front: per-corner uvs are kept only for texture-driven ones
        for (uint32_t k = 0; k < d.instances[i].n_materials; ++k) {
            uint32_t idx;
            int rc = material_of(d.instances[i].materials[k], idx);
            if (rc != AKR_OK) return rc;
        }
    out.textures_host = host_textures;
    const std::vector<uint32_t> sort_keys = material_sort_keys(out.materials);
    if (out.any_dynamic) out.corner_uvs.assign(static_cast<size_t>(total_tris) * 6, 0.0f);
    if (any_normals) out.corner_normals.assign(static_cast<size_t>(total_tris) * 9, 0.0f);
    if (any_tangents) out.corner_tangents.assign(static_cast<size_t>(total_tris) * 9, 0.0f);
    std::vector<float> world(static_cast<size_t>(total_tris) * 9);  // v0,v1,v2 world per gid
    out.instances.resize(d.n_instances);
    uint32_t tri_offset = 0;
    for (uint32_t i = 0; i < d.n_instances; ++i) {
        const AkrInstance &in = d.instances[i];
        const AkrMesh &g = d.meshes[in.geom_id];
        if (in.n_materials == 0 || !in.materials) {
            err = "instance without materials (mesh.rs:308)";
            return AKR_ERR_INVALID_ARGUMENT;
        }
        InstanceRec &ir = out.instances[i];
        std::memset(&ir, 0, sizeof(ir));
        const float *mm = in.transform;
        HM3 m{{mm[0], mm[1], mm[2]}, {mm[4], mm[5], mm[6]}, {mm[8], mm[9], mm[10]}};
        H3 tr{mm[12], mm[13], mm[14]};
        HM3 m_inv_t = hinverse(htranspose(m));
        st3(ir.m, m.c0);
        st3(ir.m + 3, m.c1);
        st3(ir.m + 6, m.c2);
        st3(ir.t, tr);
        st3(ir.m_inv_t, m_inv_t.c0);
        st3(ir.m_inv_t + 3, m_inv_t.c1);
        st3(ir.m_inv_t + 6, m_inv_t.c2);
        ir.det = mat4_det(mm);
        ir.light_id = 0xffffffffu;
        ir.tri_offset = tri_offset;
        ir.n_tris = g.n_triangles;
        ir.geom_id = in.geom_id;
        ir.flags = in.flags;
        std::vector<uint32_t> mats(in.n_materials);
        for (uint32_t k = 0; k < in.n_materials; ++k) {
            int rc = material_of(in.materials[k], mats[k]);
            if (rc != AKR_OK) return rc;
        }
        const bool multi = (in.flags & AKR_MESH_HAS_MULTI_MATERIALS) != 0;
        for (uint32_t p = 0; p < g.n_triangles; ++p) {
            uint32_t gid = tri_offset + p;
            TriShade &ts = out.shade[gid];
            std::memset(&ts, 0, sizeof(ts));
            const uint32_t *idx = g.indices + 3 * p;
            for (int k = 0; k < 3; ++k)
                if (idx[k] >= g.n_vertices) {
                    err = "vertex index out of range";
                    return AKR_ERR_INVALID_ARGUMENT;
                }
            auto vert = [&](uint32_t k) { return H3{g.vertices[3 * k], g.vertices[3 * k + 1], g.vertices[3 * k + 2]}; };
            H3 v0 = vert(idx[0]), v1 = vert(idx[1]), v2 = vert(idx[2]);
            st3(ts.v0, v0);
            st3(ts.v1, v1);
            st3(ts.v2, v2);
            // mesh.rs:526-533
            H3 ngu = hcross(v1 - v0, v2 - v0);
            float len = hlen(ngu);
            float area_local = len * 0.5f;
            H3 ng_local = ngu / len;
            // uv (mesh.rs:534-546)
            float uv0[2], uv1[2], uv2[2];
            if (g.uvs) {
                const float *u = g.uvs + static_cast<size_t>(p) * 6;
                uv0[0] = u[0]; uv0[1] = u[1]; uv1[0] = u[2]; uv1[1] = u[3]; uv2[0] = u[4]; uv2[1] = u[5];
            } else {
                uv0[0] = 0.0f; uv0[1] = 0.0f; uv1[0] = 1.0f; uv1[1] = 0.0f; uv2[0] = 1.0f; uv2[1] = 0.1f;
            }
            // dp/du fallback tangent (mesh.rs:571-590)
            H3 tt_local{0, 0, 0};
            {
                float duv02[2] = {uv0[0] - uv2[0], uv0[1] - uv2[1]};
                float duv12[2] = {uv1[0] - uv2[0], uv1[1] - uv2[1]};
                H3 dp02 = v0 - v2, dp12 = v1 - v2;
                float determinant = difference_of_products(duv02[0], duv12[1], duv02[1], duv12[0]);
                bool degenerate_uv = std::fabs(determinant) < 1e-8f;
                H3 t{0, 0, 0};
                if (!degenerate_uv) {
                    float inv_det = 1.0f / determinant;
                    t.x = difference_of_products(duv12[1], dp02.x, duv02[1], dp12.x) * inv_det;
                    t.y = difference_of_products(duv12[1], dp02.y, duv02[1], dp12.y) * inv_det;
                    t.z = difference_of_products(duv12[1], dp02.z, duv02[1], dp12.z) * inv_det;
                }
                if (degenerate_uv || hdot(t, t) == 0.0f) {
                    Frame f = frame_from_n(mk3(ng_local.x, ng_local.y, ng_local.z));
                    t = H3{f.t.x, f.t.y, f.t.z};
                }
                tt_local = t;
            }
            // world transform (mesh.rs:608-628)
            H3 tt = hmul(m, tt_local);
            H3 c = hmul(m, ng_local);
            H3 ng = hnorm(hmul(m_inv_t, ng_local));
            float area = (area_local == 0.0f || ir.det == 0.0f) ? 0.0f : std::fabs(area_local * ir.det / hdot(ng, c));
            st3(ts.ng, ng);
            ts.area = area;
            ts.inst = i;
            ts.prim = p;
            uint32_t slot = 0;
            if (multi) {
                if (p >= g.n_material_slots || g.material_slots[p] >= in.n_materials) {
                    err = "material slot out of range";
                    return AKR_ERR_INVALID_ARGUMENT;
                }
                slot = g.material_slots[p];
            }
            ts.mat = mats[slot];
            uint32_t flags = 0;
            if (g.normals) flags |= TRI_HAS_NORMALS;
            if (g.tangents) flags |= TRI_HAS_TANGENTS;
            if (g.uvs) flags |= TRI_HAS_UVS;
            if (!(out.materials[ts.mat].alpha >= 1.0f) || out.materials[ts.mat].alpha_dynamic) {  // (a texture-driven alpha is only known per hit)
                flags |= TRI_ALPHA;
                out.any_alpha = 1;
            }
            ts.flags = flags | (sort_keys[ts.mat] << TRI_SORT_KEY_SHIFT);
            if (!g.normals && !g.tangents) {
                // flat triangle: ns = ng and the frame is constant (mesh.rs:629-633)
                Frame f = (tt.x != 0.0f || tt.y != 0.0f || tt.z != 0.0f) ? frame_from_n_t(mk3(ng.x, ng.y, ng.z), mk3(tt.x, tt.y, tt.z))
                                                                           : frame_from_n(mk3(ng.x, ng.y, ng.z));
                st3(ts.ft, H3{f.t.x, f.t.y, f.t.z});
                st3(ts.fs, H3{f.s.x, f.s.y, f.s.z});
            } else {
                st3(ts.ft, tt);  // world dp/du tangent
                st3(ts.fs, tt);  // fallback when the tangent buffer holds non-finite values
            }
            if (g.uvs && !out.corner_uvs.empty()) std::memcpy(out.corner_uvs.data() + static_cast<size_t>(gid) * 6, g.uvs + static_cast<size_t>(p) * 6, 24);
            if (g.normals) std::memcpy(out.corner_normals.data() + static_cast<size_t>(gid) * 9, g.normals + static_cast<size_t>(p) * 9, 36);
            if (g.tangents) std::memcpy(out.corner_tangents.data() + static_cast<size_t>(gid) * 9, g.tangents + static_cast<size_t>(p) * 9, 36);
            // world-space vertices for traversal
            H3 w0 = hmul(m, v0) + tr, w1 = hmul(m, v1) + tr, w2 = hmul(m, v2) + tr;
            st3(world.data() + static_cast<size_t>(gid) * 9, w0);
            st3(world.data() + static_cast<size_t>(gid) * 9 + 3, w1);
            st3(world.data() + static_cast<size_t>(gid) * 9 + 6, w2);
        }
        tri_offset += g.n_triangles;
    }

    // ---- mesh lights (load.rs:312-415).  All supported emission closures are constant, so the 16 samples
    // of the estimation kernel (load.rs:319-341) contribute the same value; the f32 accumulation is literal.
    std::vector<float> light_weights;
    out.alias_j.clear();
    std::vector<std::vector<uint32_t>> per_light_j;
    std::vector<std::vector<float>> per_light_t, per_light_pdf;
    for (uint32_t i = 0; i < d.n_instances; ++i) {
        InstanceRec &ir = out.instances[i];
        std::vector<float> powers(ir.n_tris);
        for (uint32_t p = 0; p < ir.n_tris; ++p) {
            const TriShade &ts = out.shade[ir.tri_offset + p];
            const Material &mat = out.materials[ts.mat];
            float e = std::fmax(mat.emission[0], std::fmax(mat.emission[1], mat.emission[2]));
            float acc = 0.0f;
            for (int s = 0; s < 16; ++s) acc += e * ts.area;
            powers[p] = acc / 16.0f;
        }
        float total = 0.0f;
        for (float x : powers) total += x;
        if (total > 1e-4f) {
            ir.light_id = static_cast<uint32_t>(out.lights.size());
            LightRec lr{i, ir.tri_offset, 0, ir.n_tris};
            out.lights.push_back(lr);
            light_weights.push_back(total);
            out.light_powers.push_back(total);
            std::vector<uint32_t> j;
            std::vector<float> t, pdf;
            build_alias(powers, j, t, pdf);
            for (uint32_t p = 0; p < ir.n_tris; ++p) {
                out.shade[ir.tri_offset + p].prim_pdf = pdf[p];
                out.shade[ir.tri_offset + p].flags |= TRI_IS_LIGHT;
            }
            per_light_j.push_back(j);
            per_light_t.push_back(t);
            per_light_pdf.push_back(pdf);
        }
    }
    if (!light_weights.empty()) {
        build_alias(light_weights, out.alias_j, out.alias_t, out.alias_pdf);
        for (size_t l = 0; l < out.lights.size(); ++l) {
            out.lights[l].alias_offset = static_cast<uint32_t>(out.alias_j.size());
            out.alias_j.insert(out.alias_j.end(), per_light_j[l].begin(), per_light_j[l].end());
            out.alias_t.insert(out.alias_t.end(), per_light_t[l].begin(), per_light_t[l].end());
            out.alias_pdf.insert(out.alias_pdf.end(), per_light_pdf[l].begin(), per_light_pdf[l].end());
        }
    } else {
        out.alias_j.push_back(0);
        out.alias_t.push_back(1.0f);
        out.alias_pdf.push_back(0.0f);
    }

    // ---- traversal primitives: pair triangles that form a parallelogram, then a BVH over the primitives ----
    struct HostPrim {
        uint32_t gid_a, gid_b;  // gid_b = 0xffffffff: single
        double p0[3], a[3], b[3];
        uint32_t iu_a, iv_a, iu_b, iv_b;
    };
    std::vector<HostPrim> hprims;
    hprims.reserve(total_tris);
    {
        auto vtx = [&](uint32_t gid, int k) { return world.data() + static_cast<size_t>(gid) * 9 + 3 * k; };
        auto same = [&](const float *x, const float *y) { return std::memcmp(x, y, 12) == 0; };
        std::vector<uint8_t> used(total_tris, 0);
        // candidates share an edge; only triangles of the same instance are tried, in a small gid window
        // (mesh exporters emit the two halves of a quad next to each other)
        const uint32_t kWindow = 8;
        for (uint32_t ii = 0; ii < d.n_instances; ++ii) {
            const InstanceRec &ir = out.instances[ii];
            for (uint32_t i = ir.tri_offset; i < ir.tri_offset + ir.n_tris; ++i) {
                if (used[i]) continue;
                used[i] = 1;
                HostPrim hp;
                hp.gid_a = i;
                hp.gid_b = 0xffffffffu;
                hp.iu_a = hp.iv_a = hp.iu_b = hp.iv_b = 0;
                for (int c = 0; c < 3; ++c) {
                    hp.p0[c] = vtx(i, 0)[c];
                    hp.a[c] = static_cast<double>(vtx(i, 1)[c]) - vtx(i, 0)[c];
                    hp.b[c] = static_cast<double>(vtx(i, 2)[c]) - vtx(i, 0)[c];
                }
                bool paired = false;
                for (uint32_t j = i + 1; j < std::min(i + 1 + kWindow, ir.tri_offset + ir.n_tris) && !paired; ++j) {
                    if (used[j]) continue;
                    // shared edge (xa, xb) = two vertices of i equal to two vertices of j
                    for (int e = 0; e < 3 && !paired; ++e) {
                        const int ia = e, ib = (e + 1) % 3, ic = (e + 2) % 3;
                        int ja = -1, jb = -1;
                        for (int k = 0; k < 3; ++k) {
                            if (same(vtx(j, k), vtx(i, ia))) ja = k;
                            if (same(vtx(j, k), vtx(i, ib))) jb = k;
                        }
                        if (ja < 0 || jb < 0 || ja == jb) continue;
                        const int jd = 3 - ja - jb;
                        const float *xa = vtx(i, ia), *xb = vtx(i, ib), *xc = vtx(i, ic), *yd = vtx(j, jd);
                        // parallelogram <=> the diagonals bisect each other: xc + yd == xa + xb
                        double scale = 0.0, dev = 0.0;
                        for (int c = 0; c < 3; ++c) {
                            double av = static_cast<double>(xc[c]) - xa[c], bv = static_cast<double>(yd[c]) - xa[c];
                            scale = std::max(scale, std::max(std::fabs(av), std::fabs(bv)));
                            dev = std::max(dev, std::fabs((static_cast<double>(xc[c]) + yd[c]) - (static_cast<double>(xa[c]) + xb[c])));
                        }
                        if (!(dev <= 1e-6 * scale)) continue;
                        // p0 = xa, p1 = xc (triangle A = i), p2 = xb, p3 = yd (triangle B = j)
                        double cr[3], av[3], bv[3];
                        for (int c = 0; c < 3; ++c) {
                            av[c] = static_cast<double>(xc[c]) - xa[c];
                            bv[c] = static_cast<double>(yd[c]) - xa[c];
                        }
                        cr[0] = av[1] * bv[2] - av[2] * bv[1];
                        cr[1] = av[2] * bv[0] - av[0] * bv[2];
                        cr[2] = av[0] * bv[1] - av[1] * bv[0];
                        if (!(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2] > 0.0)) continue;
                        for (int c = 0; c < 3; ++c) {
                            hp.p0[c] = xa[c];
                            hp.a[c] = av[c];
                            hp.b[c] = bv[c];
                        }
                        hp.gid_b = j;
                        // local corner index (0 = p0, 1 = p1|p2, 2 = p2|p3) of each triangle's v1 and v2
                        auto local_a = [&](int k) { return k == ia ? 0u : (k == ic ? 1u : 2u); };  // A = (p0 = xa, p1 = xc, p2 = xb)
                        auto local_b = [&](int k) { return k == ja ? 0u : (k == jb ? 1u : 2u); };  // B = (p0 = xa, p2 = xb, p3 = yd)
                        hp.iu_a = local_a(1);
                        hp.iv_a = local_a(2);
                        hp.iu_b = local_b(1);
                        hp.iv_b = local_b(2);
                        used[j] = 1;
                        paired = true;
                    }
                }
                hprims.push_back(hp);
            }
        }
    }
    const uint32_t n_prims = static_cast<uint32_t>(hprims.size());
    auto prim_corners = [&](const HostPrim &hp, float (*c)[3]) -> int {  // world corners for the bounds
        for (int k = 0; k < 3; ++k) {
            c[0][k] = static_cast<float>(hp.p0[k]);
            c[1][k] = static_cast<float>(hp.p0[k] + hp.a[k]);
            c[2][k] = static_cast<float>(hp.p0[k] + hp.b[k]);
            c[3][k] = static_cast<float>(hp.p0[k] + hp.a[k] + hp.b[k]);
        }
        return hp.gid_b == 0xffffffffu ? 3 : 4;
    };
    std::vector<BuildTri> btris(n_prims);
    Box scene_box;
    scene_box.reset();
    for (uint32_t k = 0; k < n_prims; ++k) {
        BuildTri &bt = btris[k];
        bt.box.reset();
        // bounds from the triangles' own vertices (exact), plus the derived fourth corner of a pair
        for (uint32_t gid : {hprims[k].gid_a, hprims[k].gid_b}) {
            if (gid == 0xffffffffu) continue;
            const float *w = world.data() + static_cast<size_t>(gid) * 9;
            bt.box.grow(w);
            bt.box.grow(w + 3);
            bt.box.grow(w + 6);
        }
        float c4[4][3];
        int nc = prim_corners(hprims[k], c4);
        for (int c = 0; c < nc; ++c) bt.box.grow(c4[c]);
        for (int a = 0; a < 3; ++a) bt.centroid[a] = 0.5f * (bt.box.lo[a] + bt.box.hi[a]);
        bt.index = k;
        scene_box.grow(bt.box);
    }
    std::vector<BuildNode> bnodes;
    bnodes.reserve(static_cast<size_t>(n_prims) * 2);
    uint32_t max_depth = 0;
    build_recursive(bnodes, btris, 0, n_prims, 0, max_depth);
    out.bvh_depth = max_depth;
    if (max_depth + 2 >= AKR_BVH_STACK) {
        err = "BVH too deep for the traversal stack";
        return AKR_ERR_UNSUPPORTED;
    }
    // conservative padding so that rounding in the slab test can never reject a box whose primitive the
    // (independent) primitive test accepts
    float diag = 0.0f;
    for (int a = 0; a < 3; ++a) diag = std::max(diag, scene_box.hi[a] - scene_box.lo[a]);
    const float pad_abs = 1e-5f * diag;
    auto padded = [&](const Box &b, float *lo, float *hi) {
        for (int a = 0; a < 3; ++a) {
            float mag = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
            float pad = pad_abs + mag * 4e-7f;
            lo[a] = b.lo[a] - pad;
            hi[a] = b.hi[a] + pad;
        }
    };
    // primitives in leaf order; two Moeller-Trumbore triangle slots per primitive for the host simulation
    out.prims.resize(n_prims);
    out.tris.resize(static_cast<size_t>(n_prims) * 2);
    auto fill_tri = [&](TriGeom &tg, uint32_t gid) {
        std::memset(&tg, 0, sizeof(tg));
        tg.gid = gid;
        if (gid == 0xffffffffu) return;
        const float *w = world.data() + static_cast<size_t>(gid) * 9;
        H3 w0{w[0], w[1], w[2]}, w1{w[3], w[4], w[5]}, w2{w[6], w[7], w[8]};
        st3(tg.v0, w0);
        st3(tg.e1, w1 - w0);
        st3(tg.e2, w2 - w0);
        tg.cls = shade_class_of(out.materials[out.shade[gid].mat]);
    };
    for (uint32_t k = 0; k < n_prims; ++k) {
        const HostPrim &hp = hprims[btris[k].index];
        fill_tri(out.tris[2 * k], hp.gid_a);
        fill_tri(out.tris[2 * k + 1], hp.gid_b);
        PrimRec &pr = out.prims[k];
        std::memset(&pr, 0, sizeof(pr));
        pr.gid_a = hp.gid_a;
        pr.gid_b = hp.gid_b;
        // rows of the world -> (s, q) map, in double, rounded once:
        //   n = a x b;  s = (P - p0) . (b x n) / (a . (b x n));  q = (P - p0) . (n x a) / (b . (n x a))
        const double *A = hp.a, *B = hp.b, *P0 = hp.p0;
        double n[3] = {A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]};
        double bxn[3] = {B[1] * n[2] - B[2] * n[1], B[2] * n[0] - B[0] * n[2], B[0] * n[1] - B[1] * n[0]};
        double nxa[3] = {n[1] * A[2] - n[2] * A[1], n[2] * A[0] - n[0] * A[2], n[0] * A[1] - n[1] * A[0]};
        double ds = A[0] * bxn[0] + A[1] * bxn[1] + A[2] * bxn[2];
        double dq = B[0] * nxa[0] + B[1] * nxa[1] + B[2] * nxa[2];
        double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (!(nl > 0.0) || ds == 0.0 || dq == 0.0) {  // degenerate: can never be hit (NaN plane)
            for (int c = 0; c < 4; ++c) pr.n[c] = pr.r0[c] = pr.r1[c] = std::numeric_limits<float>::quiet_NaN();
        } else {
            double r0w = 0.0, r1w = 0.0, nw = 0.0;
            for (int c = 0; c < 3; ++c) {
                pr.n[c] = static_cast<float>(n[c] / nl);
                pr.r0[c] = static_cast<float>(bxn[c] / ds);
                pr.r1[c] = static_cast<float>(nxa[c] / dq);
            }
            // the offsets are computed against the ROUNDED rows so that p0 maps to (0, 0) as exactly as possible
            for (int c = 0; c < 3; ++c) {
                nw -= static_cast<double>(pr.n[c]) * P0[c];
                r0w -= static_cast<double>(pr.r0[c]) * P0[c];
                r1w -= static_cast<double>(pr.r1[c]) * P0[c];
            }
            pr.n[3] = static_cast<float>(nw);
            pr.r0[3] = static_cast<float>(r0w);
            pr.r1[3] = static_cast<float>(r1w);
        }
        uint32_t cls_a = shade_class_of(out.materials[out.shade[hp.gid_a].mat]);
        uint32_t cls_b = hp.gid_b == 0xffffffffu ? 0u : shade_class_of(out.materials[out.shade[hp.gid_b].mat]);
        uint32_t light_a = (out.shade[hp.gid_a].flags & TRI_IS_LIGHT) ? 1u : 0u;
        uint32_t light_b = (hp.gid_b != 0xffffffffu && (out.shade[hp.gid_b].flags & TRI_IS_LIGHT)) ? 1u : 0u;
        pr.meta = hp.iu_a | (hp.iv_a << 2) | (hp.iu_b << 4) | (hp.iv_b << 6) | (cls_a << 8) | (cls_b << 10) | (light_a << 12) | (light_b << 13);
    }
    if (n_prims <= kFlatMaxPrims) {
        PrimRec never;
        std::memset(&never, 0, sizeof(never));
        for (int c = 0; c < 4; ++c) never.n[c] = never.r0[c] = never.r1[c] = std::numeric_limits<float>::quiet_NaN();
        never.gid_a = never.gid_b = 0xffffffffu;
        // A primitive whose plane has every vertex of the scene on one side (a wall of the scene's convex hull)
        // can meet a segment between two scene points only at an end point, which the (0, t_max) range of a shadow
        // ray excludes: such primitives are left out of the any-hit list.
        auto supports_scene = [&](const PrimRec &pr) {
            if (!(pr.n[0] == pr.n[0])) return false;  // degenerate (NaN plane)
            bool pos = false, neg = false;
            for (size_t v = 0; v < static_cast<size_t>(total_tris) * 3; ++v) {
                const float *w = world.data() + v * 3;
                const double t0 = static_cast<double>(pr.n[0]) * w[0], t1 = static_cast<double>(pr.n[1]) * w[1], t2 = static_cast<double>(pr.n[2]) * w[2];
                const double dist = t0 + t1 + t2 + pr.n[3];
                // tolerance = a few f32 ulps of the terms of this very evaluation (the plane coefficients are rounded to f32, so
                // the wall's own vertices do not evaluate to exactly 0); anything that really protrudes past the wall keeps it
                // in the occluder list
                const double eps = 16.0 * 1.1920929e-7 * (std::fabs(t0) + std::fabs(t1) + std::fabs(t2) + std::fabs(static_cast<double>(pr.n[3])));
                pos |= dist > eps;
                neg |= dist < -eps;
                if (pos && neg) return false;
            }
            return true;
        };
        auto emit_blocks = [&](const std::vector<PrimRec> &list, uint32_t &n_pair_blocks, uint32_t &n_single_blocks) {
            std::vector<PrimRec> pairs, singles;
            for (const PrimRec &pr : list) (pr.gid_b != 0xffffffffu ? pairs : singles).push_back(pr);
            if (pairs.size() & 1u) pairs.push_back(never);
            if (singles.size() & 1u) singles.push_back(never);
            n_pair_blocks = static_cast<uint32_t>(pairs.size() / 2);
            n_single_blocks = static_cast<uint32_t>(singles.size() / 2);
            pairs.insert(pairs.end(), singles.begin(), singles.end());
            for (size_t b2 = 0; b2 < pairs.size() / 2; ++b2) {
                PrimBlock2 blk;
                std::memset(&blk, 0, sizeof(blk));
                for (int h = 0; h < 2; ++h) {
                    const PrimRec &pr = pairs[2 * b2 + h];
                    for (int c = 0; c < 4; ++c) {
                        blk.n[c][h] = pr.n[c];
                        blk.r0[c][h] = pr.r0[c];
                        blk.r1[c][h] = pr.r1[c];
                    }
                    blk.gid[2 * h] = pr.gid_a;
                    blk.gid[2 * h + 1] = pr.gid_b;
                    blk.meta[h] = pr.meta;
                }
                out.flat_blocks.push_back(blk);
            }
        };
        // complete list: occluders first inside the pair group and inside the single group, so that a kernel that walks the
        // complete list for a closest-hit ray can test a shadow ray against the leading `n_shadow_*_blocks` of each group
        // in the same trip (akari_b200.cu: trace_flat2_dual)
        std::vector<PrimRec> occluders, walls, ordered;
        for (const PrimRec &pr : out.prims) (supports_scene(pr) ? walls : occluders).push_back(pr);
        ordered = occluders;
        ordered.insert(ordered.end(), walls.begin(), walls.end());
        emit_blocks(ordered, out.n_pair_blocks, out.n_single_blocks);
        uint32_t occ_pairs = 0, occ_singles = 0;
        for (const PrimRec &pr : occluders) (pr.gid_b != 0xffffffffu ? occ_pairs : occ_singles) += 1u;
        out.n_shadow_pair_blocks = (occ_pairs + 1u) / 2u;
        out.n_shadow_single_blocks = (occ_singles + 1u) / 2u;
        emit_blocks(occluders, out.n_occ_pair_blocks, out.n_occ_single_blocks);
    }
    // flatten: breadth-first over inner nodes; a leaf root becomes an inner node with an empty second child
    auto leaf_code = [&](const BuildNode &n) { return ~static_cast<int32_t>((n.first << 3) | n.count); };
    std::vector<int32_t> inner_of(bnodes.size(), -1);
    std::vector<int32_t> order;
    if (bnodes[0].left < 0) {
        BvhNode root;
        std::memset(&root, 0, sizeof(root));
        padded(bnodes[0].box, root.lo0, root.hi0);
        for (int a = 0; a < 3; ++a) {
            root.lo1[a] = std::numeric_limits<float>::infinity();
            root.hi1[a] = -std::numeric_limits<float>::infinity();
        }
        root.c0 = leaf_code(bnodes[0]);
        root.c1 = ~0;  // empty leaf
        out.nodes.push_back(root);
    } else {
        order.push_back(0);
        inner_of[0] = 0;
        for (size_t h = 0; h < order.size(); ++h) {
            const BuildNode &n = bnodes[order[h]];
            for (int32_t ch : {n.left, n.right})
                if (bnodes[ch].left >= 0) {
                    inner_of[ch] = static_cast<int32_t>(order.size());
                    order.push_back(ch);
                }
        }
        out.nodes.resize(order.size());
        for (size_t h = 0; h < order.size(); ++h) {
            const BuildNode &n = bnodes[order[h]];
            BvhNode &o = out.nodes[h];
            std::memset(&o, 0, sizeof(o));
            const BuildNode &l = bnodes[n.left], &r = bnodes[n.right];
            padded(l.box, o.lo0, o.hi0);
            padded(r.box, o.lo1, o.hi1);
            o.c0 = l.left >= 0 ? inner_of[n.left] : leaf_code(l);
            o.c1 = r.left >= 0 ? inner_of[n.right] : leaf_code(r);
        }
    }

    // ---- 4-wide collapse: every inner node adopts its grandchildren, largest box first, until it has 4 children ----
    {
        struct Wide {
            int32_t kids[4];  // build-node indices, -1 = empty
            uint32_t depth;
        };
        std::vector<Wide> wide;
        std::vector<int32_t> wide_of(bnodes.size(), -1);
        std::vector<int32_t> worder;  // build-node index of each wide node, breadth-first
        auto make_wide = [&](int32_t bn) {
            std::vector<int32_t> kids;
            if (bnodes[bn].left < 0) {
                kids.push_back(bn);  // a leaf root
            } else {
                kids = {bnodes[bn].left, bnodes[bn].right};
                while (kids.size() < 4) {
                    int best = -1;
                    float best_area = -1.0f;
                    for (size_t k = 0; k < kids.size(); ++k)
                        if (bnodes[kids[k]].left >= 0 && bnodes[kids[k]].box.half_area() > best_area) {
                            best = static_cast<int>(k);
                            best_area = bnodes[kids[k]].box.half_area();
                        }
                    if (best < 0) break;
                    int32_t open = kids[best];
                    kids[best] = bnodes[open].left;
                    kids.push_back(bnodes[open].right);
                }
            }
            Wide w;
            for (int k = 0; k < 4; ++k) w.kids[k] = k < static_cast<int>(kids.size()) ? kids[k] : -1;
            w.depth = 0;
            return w;
        };
        worder.push_back(0);
        wide_of[0] = 0;
        wide.push_back(make_wide(0));
        uint32_t depth4 = 0;
        for (size_t h = 0; h < wide.size(); ++h) {
            for (int k = 0; k < 4; ++k) {
                int32_t kid = wide[h].kids[k];
                if (kid >= 0 && bnodes[kid].left >= 0) {
                    wide_of[kid] = static_cast<int32_t>(wide.size());
                    Wide w = make_wide(kid);
                    w.depth = wide[h].depth + 1;
                    depth4 = std::max(depth4, w.depth);
                    wide.push_back(w);
                }
            }
        }
        out.bvh4_depth = depth4;
        out.nodes4.resize(wide.size());
        for (size_t h = 0; h < wide.size(); ++h) {
            Bvh4Node &o = out.nodes4[h];
            std::memset(&o, 0, sizeof(o));
            for (int k = 0; k < 4; ++k) {
                int32_t kid = wide[h].kids[k];
                if (kid < 0) {
                    for (int a = 0; a < 3; ++a) {
                        o.lo[a][k] = std::numeric_limits<float>::infinity();
                        o.hi[a][k] = -std::numeric_limits<float>::infinity();
                    }
                    o.c[k] = ~0;
                    continue;
                }
                float lo[3], hi[3];
                padded(bnodes[kid].box, lo, hi);
                for (int a = 0; a < 3; ++a) {
                    o.lo[a][k] = lo[a];
                    o.hi[a][k] = hi[a];
                }
                o.c[k] = bnodes[kid].left >= 0 ? wide_of[kid] : leaf_code(bnodes[kid]);
            }
        }
    }

    // ---- camera (camera/mod.rs:119-153; load.rs:172-194) ----
    {
        const AkrPerspectiveCamera &cam = d.camera;
        if (cam.width == 0 || cam.height == 0) {
            err = "camera resolution is zero";
            return AKR_ERR_INVALID_ARGUMENT;
        }
        CameraRec &c = out.camera;
        std::memset(&c, 0, sizeof(c));
        const float *m = cam.c2w;
        for (int col = 0; col < 3; ++col)
            for (int r = 0; r < 3; ++r) c.c2w[col * 3 + r] = m[col * 4 + r];
        c.c2w[9] = m[12];
        c.c2w[10] = m[13];
        c.c2w[11] = m[14];
        bool ident = true;  // Mat4::abs_diff_eq(IDENTITY, 1e-4)
        for (int col = 0; col < 4; ++col)
            for (int r = 0; r < 4; ++r) {
                float idv = col == r ? 1.0f : 0.0f;
                if (!(std::fabs(m[col * 4 + r] - idv) <= 1e-4f)) ident = false;
            }
        c.c2w_identity = ident ? 1u : 0u;
        float fx = static_cast<float>(cam.width), fy = static_cast<float>(cam.height);
        float sx = 1.0f / fx, sy = 1.0f / fy, sz = 1.0f, tx = 0.0f, ty = 0.0f, tz = 0.0f;
        auto scale = [&](float a, float b, float cc) {
            sx = a * sx; sy = b * sy; sz = cc * sz;
            tx = a * tx; ty = b * ty; tz = cc * tz;
        };
        auto translate = [&](float a, float b, float cc) { tx = tx + a; ty = ty + b; tz = tz + cc; };
        scale(2.0f, 2.0f, 1.0f);
        translate(-1.0f, -1.0f, 0.0f);
        scale(1.0f, -1.0f, 1.0f);
        float s = std::tan(cam.fov / 2.0f);
        if (cam.width > cam.height) scale(s, s * fy / fx, 1.0f);
        else scale(s * fx / fy, s, 1.0f);
        translate(0.0f, 0.0f, -1.0f);
        c.r2c_s[0] = sx; c.r2c_s[1] = sy; c.r2c_s[2] = sz;
        c.r2c_t[0] = tx; c.r2c_t[1] = ty; c.r2c_t[2] = tz;
        c.width = cam.width;
        c.height = cam.height;
    }
    return AKR_OK;
}

SceneView host_scene_view(const HostSceneBlob &b, const float *albedo_table) {
    SceneView v;
    std::memset(&v, 0, sizeof(v));
    v.nodes = b.nodes.data();
    v.nodes4 = b.nodes4.data();
    v.n_nodes4 = static_cast<uint32_t>(b.nodes4.size());
    v.bvh4_depth = b.bvh4_depth;
    v.prims = b.prims.data();
    v.flat_blocks = b.flat_blocks.empty() ? nullptr : b.flat_blocks.data();
    v.n_pair_blocks = b.n_pair_blocks;
    v.n_single_blocks = b.n_single_blocks;
    v.n_occ_pair_blocks = b.n_occ_pair_blocks;
    v.n_occ_single_blocks = b.n_occ_single_blocks;
    v.n_shadow_pair_blocks = b.n_shadow_pair_blocks;
    v.n_shadow_single_blocks = b.n_shadow_single_blocks;
    v.tris = b.tris.data();
    v.shade = b.shade.data();
    v.instances = b.instances.data();
    v.materials = b.materials.data();
    v.lights = b.lights.data();
    v.alias_j = b.alias_j.data();
    v.alias_t = b.alias_t.data();
    v.alias_pdf = b.alias_pdf.data();
    v.albedo_table = albedo_table;
    v.svm.nodes = b.svm_nodes.data();
    v.svm.kind_first = b.svm_kind_first.data();
    v.svm.data = b.svm_data.data();
    v.svm.textures = b.textures_host.empty() ? nullptr : b.textures_host.data();
    v.svm.n_kinds = static_cast<uint32_t>(b.svm_kind_first.size() - 1);
    v.svm.n_textures = static_cast<uint32_t>(b.textures_host.size());
    v.svm.data_size = static_cast<uint32_t>(b.svm_data.size());
    v.svm.kind_hit_mask = b.svm_kind_hit_mask.data();
    v.svm.static_vals = b.svm_static_vals.empty() ? nullptr : b.svm_static_vals.data();
    v.corner_uvs = b.corner_uvs.empty() ? nullptr : b.corner_uvs.data();
    v.n_nodes = static_cast<uint32_t>(b.nodes.size());
    v.n_prims = static_cast<uint32_t>(b.prims.size());
    v.n_tris = static_cast<uint32_t>(b.shade.size());
    v.n_instances = static_cast<uint32_t>(b.instances.size());
    v.n_materials = static_cast<uint32_t>(b.materials.size());
    v.n_lights = static_cast<uint32_t>(b.lights.size());
    v.any_alpha = b.any_alpha;
    v.camera = b.camera;
    return v;
}

void make_albedo_table(float *table, uint32_t n) {
    for (uint32_t cell = 0; cell < 4096; ++cell) table[cell] = albedo_table_cell(cell, n);
}

}  // namespace akr

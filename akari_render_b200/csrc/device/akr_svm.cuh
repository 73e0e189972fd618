// akr_svm.cuh — shader virtual machine: evaluates a material's node program into a `Material` record.
//
// Reference: crates/akari_render/src/svm/eval.rs:97-269 (node semantics), :296-362 (auto-convert rules),
// svm/surface/principled.rs:13-216, diffuse.rs:82-104, glass.rs:13-45 (closure inputs).
//
// The same function serves three callers:
//   * scene upload (host, csrc/host/scene_build.cpp): every material is evaluated ONCE; when no node depends on the hit
//     (no image texture, no texture coordinates) the result is the constant-folded Material the shade kernels read;
//   * the shade / bounce kernels of shade class GENERAL (device): materials flagged `dynamic` are re-evaluated per hit
//     with the hit's uv — exactly what the reference's per-dispatch interpreter does for every material;
//   * the alpha test of the traversal kernels (device): SvmEvalMode::Alpha, only Material.alpha is used.
// Image sampling stands in for luisa's bindless `Tex2d::sample` (third party, ASSUMED): normalised coordinates, texel
// centres at (i + 0.5) / size, bilinear weights in f32, the sampler's address mode applied per texel.
#pragma once
#include "../../../include/akari_b200.h"
#include "akr_bsdf.cuh"

namespace akr {

struct TextureRec {            // one (image, sampler) slot (load.rs:611-646)
    const void *texels;        // [height][width][4] u8 unorm or f32
    uint32_t width, height;
    uint32_t texel_format;     // AKR_TEXEL_*
    uint32_t address, filter;  // AKR_ADDRESS_*, AKR_FILTER_*
    uint32_t _pad;
};
constexpr uint32_t kSvmMaxNodes = 64;
enum SvmKind : uint32_t { SV_NONE = 0, SV_F, SV_F2, SV_F3, SV_F4, SV_COLOR_ALPHA, SV_CLOSURE, SV_TEXCOORDS, SV_SEPARATE };
struct SvmVal {
    float v[4];
    uint32_t kind;
};
struct SvmView {
    const AkrSvmNode *nodes;        // every shader kind's nodes, concatenated
    const uint32_t *kind_first;     // [n_kinds + 1] first node of kind k
    const uint8_t *data;            // constant blob (AkrSceneDesc.shader_data)
    const TextureRec *textures;
    uint32_t n_kinds, n_textures;
    uint32_t data_size, _pad;
    // Per-hit evaluation of a texture-driven material only runs the nodes that depend on the hit: bit i of
    // kind_hit_mask[kind] marks node i as hit-dependent (an image texture, texture coordinates, a checkerboard on the mesh
    // uvs, or anything downstream of one — a property of the node program, not of its constants); the values of the other
    // nodes were computed once at upload and sit in static_vals[Material.static_offset + i] (scene_build.cpp
    // fold_material).  The reference interprets the whole program per dispatch (eval.rs:364-380); the values are the same.
    const uint64_t *kind_hit_mask;  // [n_kinds]
    const SvmVal *static_vals;
};
// what the validating evaluation at upload reports besides the folded Material
struct SvmFoldInfo {
    uint64_t hit_mask;            // see SvmView::kind_hit_mask
    SvmVal vals[kSvmMaxNodes];    // value of every node at uv = (0, 0): exact for the nodes outside hit_mask
};
enum SvmStatus : int { SVM_OK = 0, SVM_BAD_PROGRAM = 1, SVM_UNSUPPORTED = 2 };

AKR_HD float sv_float_auto(const SvmVal &v) { return v.v[0]; }  // eval_float_auto_convert (eval.rs:327-343): .x of whatever it is
AKR_HD void sv_float3_auto(const SvmVal &v, float out[3]) {      // eval_float3_auto_convert (eval.rs:311-326)
    out[0] = v.v[0];
    out[1] = (v.kind == SV_F) ? 0.0f : v.v[1];
    out[2] = (v.kind == SV_F3 || v.kind == SV_F4) ? v.v[2] : 0.0f;
}
AKR_HD f2 sv_float2_auto(const SvmVal &v) { return f2{v.v[0], v.kind == SV_F ? 0.0f : v.v[1]}; }  // eval.rs:296-310

// ---- image textures -----------------------------------------------------------------------------------------------
AKR_HD int wrap_texel(int i, int n, uint32_t address, bool &zero) {
    zero = false;
    if (i >= 0 && i < n) return i;
    if (address == AKR_ADDRESS_REPEAT) {
        int m = i % n;
        return m < 0 ? m + n : m;
    }
    if (address == AKR_ADDRESS_MIRROR) {
        int period = 2 * n;
        int m = i % period;
        if (m < 0) m += period;
        return m < n ? m : period - 1 - m;
    }
    if (address == AKR_ADDRESS_EDGE) return i < 0 ? 0 : n - 1;
    zero = true;
    return 0;
}
// texel (ix, iy) of an image, coordinates already wrapped; `zero`: outside under the zero address mode
AKR_HD void load_texel(const TextureRec &t, int ix, int iy, bool zero, float out[4]) {
    if (zero) {
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        return;
    }
    size_t i = ((size_t)iy * t.width + (size_t)ix) * 4u;
    if (t.texel_format == AKR_TEXEL_RGBA8) {
        const uint8_t *p = static_cast<const uint8_t *>(t.texels) + i;
        for (int c = 0; c < 4; ++c) out[c] = (float)p[c] / 255.0f;
    } else {
        const float *p = static_cast<const float *>(t.texels) + i;
        for (int c = 0; c < 4; ++c) out[c] = p[c];
    }
}
AKR_HD void fetch_texel(const TextureRec &t, int x, int y, float out[4]) {
    bool zx, zy;
    int ix = wrap_texel(x, (int)t.width, t.address, zx);
    int iy = wrap_texel(y, (int)t.height, t.address, zy);
    load_texel(t, ix, iy, zx || zy, out);
}
AKR_HD void sample_texture(const TextureRec &t, f2 uv, float out[4]) {
    float fx = uv.x * (float)t.width, fy = uv.y * (float)t.height;
    if (t.filter == AKR_FILTER_POINT) {
        fetch_texel(t, (int)floorf(fx), (int)floorf(fy), out);
        return;
    }
    float x = fx - 0.5f, y = fy - 0.5f;
    float x0 = floorf(x), y0 = floorf(y);
    float tx = x - x0, ty = y - y0;
    int ix = (int)x0, iy = (int)y0;
    // the four texels share two columns and two rows: four coordinate wraps (integer modulo) instead of eight
    bool zx0, zx1, zy0, zy1;
    const int wx0 = wrap_texel(ix, (int)t.width, t.address, zx0), wx1 = wrap_texel(ix + 1, (int)t.width, t.address, zx1);
    const int wy0 = wrap_texel(iy, (int)t.height, t.address, zy0), wy1 = wrap_texel(iy + 1, (int)t.height, t.address, zy1);
    float c00[4], c10[4], c01[4], c11[4];
    load_texel(t, wx0, wy0, zx0 || zy0, c00);
    load_texel(t, wx1, wy0, zx1 || zy0, c10);
    load_texel(t, wx0, wy1, zx0 || zy1, c01);
    load_texel(t, wx1, wy1, zx1 || zy1, c11);
    for (int c = 0; c < 4; ++c) {
        float a = c00[c] * (1.0f - tx) + c10[c] * tx;
        float b = c01[c] * (1.0f - tx) + c11[c] * tx;
        out[c] = a * (1.0f - ty) + b * ty;
    }
}
AKR_HD float srgb_to_linear1(float s) { return s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }  // color.rs:555-558

// ---- closure inputs -> Material ------------------------------------------------------------------------------------
// Gulbrandsen parametrisation (svm/surface/mod.rs:1040-1052), per channel
AKR_HD void artistic_to_conductor(float c, float g, float &n, float &k) {
    float r = fminf(fmaxf(c, 0.0f), 0.99f);
    float r_sqrt = sqrtf(r);
    float n_min = (1.0f - r) / (1.0f + r);
    float n_max = (1.0f + r_sqrt) / (1.0f - r_sqrt);
    n = g * (n_min - n_max) + n_max;  // n_max.lerp(n_min, g)
    float k2 = ((n + 1.0f) * (n + 1.0f) * r - (n - 1.0f) * (n - 1.0f)) / (1.0f - r);
    k2 = fmaxf(k2, 0.0f);
    k = sqrtf(k2);
}
AKR_HD float ior_from_f0(float f0) {  // mod.rs:1090-1094
    float s = sqrtf(fminf(fmaxf(f0, 0.0f), 0.99f));
    return (1.0f + s) / (1.0f - s);
}
AKR_HD float f0_from_ior(float ior) {  // mod.rs:1095-1098
    float f = (ior - 1.0f) / (ior + 1.0f);
    return f * f;
}

// Evaluates shader (kind, data_offset) at texture coordinates `uv` into `m`.  `dynamic` (optional) reports whether any
// node depends on the hit.  ALPHA_ONLY (SvmEvalMode::Alpha, eval.rs:158-166; principled.rs:15-22; diffuse.rs:85-92)
// skips everything of the closure but Material.alpha.  VALIDATE adds the bounds / ordering / type checks the upload
// runs once per material, so that the per-hit evaluation on the device can trust the program.
template <bool ALPHA_ONLY, bool VALIDATE>
AKR_HD int svm_eval(const SvmView &svm, uint32_t shader_kind, uint32_t data_offset, f2 uv, Material &m, bool *dynamic, uint32_t static_offset = AKR_SVM_NONE,
                    SvmFoldInfo *fold = nullptr) {
    if (VALIDATE && shader_kind >= svm.n_kinds) return SVM_BAD_PROGRAM;
    const uint32_t first = svm.kind_first[shader_kind], n_nodes = svm.kind_first[shader_kind + 1u] - first;
    if (VALIDATE && (n_nodes == 0u || n_nodes > kSvmMaxNodes)) return SVM_UNSUPPORTED;
    // `static_offset` given (per-hit evaluation of a texture-driven material): only hit-dependent nodes are interpreted,
    // into `local`; every other node's value is read from the table computed at upload.  Otherwise all nodes run.
    const bool partial = !VALIDATE && static_offset != AKR_SVM_NONE;
    const uint64_t hit_mask = partial ? svm.kind_hit_mask[shader_kind] : ~0ull;
    const SvmVal *table = partial ? svm.static_vals + static_offset : nullptr;
    SvmVal local[kSvmMaxNodes];
    uint64_t fold_mask = 0ull;  // VALIDATE: nodes found to depend on the hit
    struct Vals {
        SvmVal *local;
        const SvmVal *table;
        uint64_t hit_mask;
        AKR_HD const SvmVal &operator[](uint32_t j) const { return ((hit_mask >> j) & 1ull) ? local[j] : table[j]; }  // (full evaluation: hit_mask = ~0)
    } vals{local, table, hit_mask};
    bool dyn = false, have_closure = false;
    m.alpha = 1.0f;
    m.type = MAT_EMISSION;
    const uint8_t *blob = svm.data + data_offset;
    auto rd = [&](uint32_t off) {
        float f;
        memcpy(&f, blob + off, 4);
        return f;
    };
    for (uint32_t i = 0; i < n_nodes; ++i) {
        if (partial && !((hit_mask >> i) & 1ull)) continue;
        const AkrSvmNode &n = svm.nodes[first + i];
        if (VALIDATE) {
            if (n.op == AKR_SVM_RGB_IMAGE_TEX || n.op == AKR_SVM_TEX_COORDS || (n.op == AKR_SVM_CHECKERBOARD && n.a[0] == AKR_SVM_NONE)) fold_mask |= 1ull << i;
            for (uint32_t k = 0; k < n.n_args; ++k) {
                const bool is_const = (n.op == AKR_SVM_FLOAT || n.op == AKR_SVM_FLOAT3) || (n.op == AKR_SVM_RGB_TEX && k == 1u) ||
                                      (n.op == AKR_SVM_RGB_IMAGE_TEX && k <= 1u) || (n.op == AKR_SVM_MAPPING && k == 1u) ||
                                      (n.op == AKR_SVM_EXTRACT_FIELD && k == 1u);
                if (is_const) {
                    if ((n.op == AKR_SVM_FLOAT || n.op == AKR_SVM_FLOAT3 || (n.op == AKR_SVM_RGB_IMAGE_TEX && k == 0u)) &&
                        (size_t)data_offset + n.a[k] + (n.op == AKR_SVM_FLOAT3 ? 12u : 4u) > svm.data_size)
                        return SVM_BAD_PROGRAM;
                    continue;
                }
                if (n.a[k] == AKR_SVM_NONE && ((n.op == AKR_SVM_RGB_IMAGE_TEX && k == 2u) || (n.op == AKR_SVM_CHECKERBOARD && k == 0u))) continue;
                if (n.a[k] >= i) return SVM_BAD_PROGRAM;  // refers to a later node
                if ((fold_mask >> n.a[k]) & 1ull) fold_mask |= 1ull << i;  // downstream of a hit-dependent node
            }
        }
        SvmVal &r = local[i];
        r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0f;
        r.kind = SV_NONE;
        switch (n.op) {
        case AKR_SVM_FLOAT:
            r.kind = SV_F;
            r.v[0] = rd(n.a[0]);
            break;
        case AKR_SVM_FLOAT3:
            r.kind = SV_F3;
            r.v[0] = rd(n.a[0]);
            r.v[1] = rd(n.a[0] + 4u);
            r.v[2] = rd(n.a[0] + 8u);
            break;
        case AKR_SVM_RGB_TEX:  // eval.rs:127-136; sRGB -> sRGB working space is the identity (texture/mod.rs:9-31)
            if (VALIDATE && n.a[1] != 1u) return SVM_UNSUPPORTED;
            r.kind = SV_F4;
            r.v[0] = vals[n.a[0]].v[0];
            r.v[1] = vals[n.a[0]].v[1];
            r.v[2] = vals[n.a[0]].v[2];
            r.v[3] = 1.0f;
            break;
        case AKR_SVM_RGB_IMAGE_TEX: {  // eval.rs:137-157
            uint32_t tex;
            memcpy(&tex, blob + n.a[0], 4);
            if (VALIDATE && tex >= svm.n_textures) return SVM_BAD_PROGRAM;
            const f2 tuv = n.a[2] != AKR_SVM_NONE ? sv_float2_auto(vals[n.a[2]]) : uv;
            sample_texture(svm.textures[tex], tuv, r.v);
            if (n.a[1] != 0u) {
                AKR_NO_UNROLL
                for (int c = 0; c < 3; ++c) r.v[c] = srgb_to_linear1(r.v[c]);  // (one copy of powf)
            }
            r.kind = SV_F4;
            dyn = true;
            break;
        }
        case AKR_SVM_SPECTRAL_UPLIFT:  // eval.rs:158-180 (RGB pass-through, alpha carried along)
            if (VALIDATE && vals[n.a[0]].kind != SV_F4) return SVM_BAD_PROGRAM;
            r = vals[n.a[0]];
            r.kind = SV_COLOR_ALPHA;
            break;
        case AKR_SVM_NORMAL_MAP: {  // eval.rs:182-196
            float nv[3];
            sv_float3_auto(vals[n.a[0]], nv);
            f3 normal = 2.0f * mk3(nv[0], nv[1], nv[2]) - splat3(1.0f);
            const float strength = sv_float_auto(vals[n.a[1]]);
            if (strength != 1.0f) normal = normal * mk3(strength, strength, 1.0f);
            r.kind = SV_F3;
            r.v[0] = normal.x;
            r.v[1] = normal.y;
            r.v[2] = normal.z;
            break;
        }
        case AKR_SVM_MAPPING: {  // eval.rs:197-213 (rotation is a todo in the reference)
            float a[3], l[3], sc[3];
            sv_float3_auto(vals[n.a[0]], a);
            sv_float3_auto(vals[n.a[2]], l);
            sv_float3_auto(vals[n.a[4]], sc);
            const f3 v = mk3(a[0], a[1], a[2]), loc = mk3(l[0], l[1], l[2]), scale = mk3(sc[0], sc[1], sc[2]);
            const f3 o = n.a[1] == 0u ? v * scale + loc : (v - loc) / scale;
            r.kind = SV_F3;
            r.v[0] = o.x;
            r.v[1] = o.y;
            r.v[2] = o.z;
            break;
        }
        case AKR_SVM_TEX_COORDS:  // eval.rs:225-232
            r.kind = SV_TEXCOORDS;
            r.v[0] = uv.x;
            r.v[1] = uv.y;
            dyn = true;
            break;
        case AKR_SVM_SEPARATE_COLOR:  // eval.rs:249-264
            sv_float3_auto(vals[n.a[0]], r.v);
            r.kind = SV_SEPARATE;
            break;
        case AKR_SVM_EXTRACT_FIELD: {  // eval.rs:214-224
            const SvmVal &src = vals[n.a[0]];
            if (src.kind == SV_TEXCOORDS && n.a[1] == AKR_SVM_FIELD_UV) {
                r.kind = SV_F2;
                r.v[0] = src.v[0];
                r.v[1] = src.v[1];
            } else if (src.kind == SV_SEPARATE && n.a[1] >= AKR_SVM_FIELD_RED && n.a[1] <= AKR_SVM_FIELD_BLUE) {
                r.kind = SV_F;
                r.v[0] = src.v[n.a[1] - AKR_SVM_FIELD_RED];
            } else if (VALIDATE) {
                return SVM_BAD_PROGRAM;  // "Field not found"
            }
            break;
        }
        case AKR_SVM_CHECKERBOARD: {  // eval.rs:233-248
            const f2 cuv = n.a[0] != AKR_SVM_NONE ? sv_float2_auto(vals[n.a[0]]) : uv;
            if (n.a[0] == AKR_SVM_NONE) dyn = true;
            if (VALIDATE && (vals[n.a[2]].kind != SV_COLOR_ALPHA || vals[n.a[3]].kind != SV_COLOR_ALPHA || vals[n.a[1]].kind != SV_F)) return SVM_BAD_PROGRAM;
            const float scale = vals[n.a[1]].v[0];
            const int px = (int)floorf(cuv.x * scale * 2.0f), py = (int)floorf(cuv.y * scale * 2.0f);
            r = ((px + py) % 2 == 0) ? vals[n.a[2]] : vals[n.a[3]];
            r.kind = SV_COLOR_ALPHA;
            break;
        }
        case AKR_SVM_DIFFUSE_BSDF: {  // diffuse.rs:82-104
            const SvmVal &c = vals[n.a[0]];
            if (VALIDATE && c.kind != SV_COLOR_ALPHA) return SVM_BAD_PROGRAM;
            r.kind = SV_CLOSURE;
            m.type = MAT_LAMBERT;
            m.wrap_inner = 0u;
            m.alpha = c.v[3];
            have_closure = true;
            if (!ALPHA_ONLY)
                for (int c3 = 0; c3 < 3; ++c3) {
                    m.color[c3] = c.v[c3];
                    m.diffuse[c3] = c.v[c3] * AKR_FRAC_1_PI;
                }
            break;
        }
        case AKR_SVM_EMISSION: {  // svm/mod.rs:124-133
            const SvmVal &c = vals[n.a[0]];
            if (VALIDATE && (c.kind != SV_COLOR_ALPHA || vals[n.a[1]].kind != SV_F)) return SVM_BAD_PROGRAM;
            const float s = vals[n.a[1]].v[0];
            r.kind = SV_CLOSURE;
            m.type = MAT_EMISSION;
            for (int c3 = 0; c3 < 3; ++c3) m.emission[c3] = c.v[c3] * s;
            have_closure = true;
            break;
        }
        case AKR_SVM_GLASS_BSDF: {  // glass.rs:13-45
            if (VALIDATE && (vals[n.a[0]].kind != SV_COLOR_ALPHA || vals[n.a[1]].kind != SV_COLOR_ALPHA || vals[n.a[2]].kind != SV_F || vals[n.a[3]].kind != SV_F))
                return SVM_BAD_PROGRAM;
            r.kind = SV_CLOSURE;
            m.type = MAT_GLASS;
            have_closure = true;
            if (!ALPHA_ONLY) {
                for (int c3 = 0; c3 < 3; ++c3) {
                    m.color[c3] = vals[n.a[0]].v[c3];
                    m.trans_color[c3] = vals[n.a[1]].v[c3];
                }
                m.roughness_raw = vals[n.a[2]].v[0];
                m.roughness = m.roughness_raw;
                m.eta = vals[n.a[3]].v[0];
            }
            break;
        }
        case AKR_SVM_PRINCIPLED_BSDF: {  // principled.rs:23-49
            if (VALIDATE) {
                if (n.n_args != 25u) return SVM_BAD_PROGRAM;
                if (vals[n.a[AKR_P_BASE_COLOR]].kind != SV_COLOR_ALPHA || vals[n.a[AKR_P_EMISSION_COLOR]].kind != SV_COLOR_ALPHA ||
                    vals[n.a[AKR_P_SPECULAR_TINT]].kind != SV_COLOR_ALPHA || vals[n.a[AKR_P_COAT_TINT]].kind != SV_COLOR_ALPHA ||
                    vals[n.a[AKR_P_ROUGHNESS]].kind != SV_F)
                    return SVM_BAD_PROGRAM;
            }
            r.kind = SV_CLOSURE;
            m.alpha = vals[n.a[AKR_P_BASE_COLOR]].v[3];
            m.wrap_inner = 1u;
            have_closure = true;
            if (ALPHA_ONLY) {
                m.type = MAT_PRINCIPLED;
                break;
            }
            auto col = [&](uint32_t k, float out[3]) {
                const SvmVal &c = vals[n.a[k]];
                out[0] = c.v[0];
                out[1] = c.v[1];
                out[2] = c.v[2];
            };
            auto flt = [&](uint32_t k) { return sv_float_auto(vals[n.a[k]]); };
            col(AKR_P_BASE_COLOR, m.color);
            float em[3];
            col(AKR_P_EMISSION_COLOR, em);
            const float es = flt(AKR_P_EMISSION_STRENGTH);
            for (int c3 = 0; c3 < 3; ++c3) {
                m.emission[c3] = em[c3] * es;
                m.diffuse[c3] = m.color[c3] * AKR_FRAC_1_PI;
                m.trans_color[c3] = sqrtf(m.color[c3]);
            }
            m.metallic = flt(AKR_P_METALLIC);
            m.roughness = flt(AKR_P_ROUGHNESS);
            m.roughness_raw = vals[n.a[AKR_P_ROUGHNESS]].v[0];
            m.eta = flt(AKR_P_IOR);
            m.transmission = flt(AKR_P_TRANSMISSION_WEIGHT);
            const float level = flt(AKR_P_SPECULAR_IOR_LEVEL);
            col(AKR_P_SPECULAR_TINT, m.spec_tint);
            // specular layer (principled.rs:55-61)
            float eta_s = m.eta;
            float f0 = f0_from_ior(eta_s);
            if (level != 0.5f) {
                f0 *= 2.0f * level;
                eta_s = ior_from_f0(f0);
            }
            m.f0 = f0;
            m.eta_s = eta_s;
            m.coat_weight = flt(AKR_P_COAT_WEIGHT);
            m.coat_roughness = flt(AKR_P_COAT_ROUGHNESS);
            m.coat_ior = flt(AKR_P_COAT_IOR);
            float tint[3];
            col(AKR_P_COAT_TINT, tint);
            for (int c3 = 0; c3 < 3; ++c3) m.coat_scale[c3] = m.coat_weight * (tint[c3] - 1.0f) + 1.0f;  // white.lerp(tint, w)
            for (int c3 = 0; c3 < 3; ++c3) artistic_to_conductor(m.color[c3], m.spec_tint[c3], m.metal_n[c3], m.metal_k[c3]);
            float nrm[3];
            sv_float3_auto(vals[n.a[AKR_P_NORMAL]], nrm);
            m.normal[0] = -nrm[0];
            m.normal[1] = -nrm[1];
            m.normal[2] = nrm[2];
            m.has_normal = (m.normal[0] != 0.0f || m.normal[1] != 0.0f || m.normal[2] != 0.0f) ? 1u : 0u;
            const float EPS = 1e-4f;  // BsdfMixture::EPS
            uint32_t lobes = 0u;
            const bool spec_zero = (m.f0 == 0.0f) || (m.spec_tint[0] == 0.0f && m.spec_tint[1] == 0.0f && m.spec_tint[2] == 0.0f);
            if (m.coat_weight != 0.0f) lobes |= LOBE_COAT;
            if (!spec_zero) lobes |= LOBE_SPECULAR;
            if (m.metallic < 1.0f - EPS) lobes |= LOBE_BASE;
            if (m.metallic > EPS) lobes |= LOBE_METAL;
            if ((lobes & LOBE_BASE) && m.transmission < 1.0f - EPS) lobes |= LOBE_DIFFUSE;
            if ((lobes & LOBE_BASE) && m.transmission > EPS) lobes |= LOBE_TRANSMISSION;
            m.lobes = lobes;
            // exact reductions (see akr_bsdf.cuh header): only when the mix fractions are exactly 0 / 1
            if (!(lobes & LOBE_COAT) && m.metallic == 0.0f && !(lobes & LOBE_SPECULAR) && m.transmission == 0.0f) m.type = MAT_LAMBERT;
            else if (!(lobes & LOBE_COAT) && m.metallic == 1.0f) m.type = MAT_CONDUCTOR;
            else m.type = MAT_PRINCIPLED;
            break;
        }
        case AKR_SVM_MATERIAL_OUTPUT:
            r.kind = SV_CLOSURE;
            break;
        default:
            if (VALIDATE) return SVM_UNSUPPORTED;
            break;
        }
    }
    if (dynamic) *dynamic = dyn;
    if (VALIDATE && !have_closure) return SVM_BAD_PROGRAM;
    if (VALIDATE && fold) {
        fold->hit_mask = fold_mask;
        for (uint32_t i = 0; i < n_nodes; ++i) fold->vals[i] = local[i];
    }
    return SVM_OK;
}

}  // namespace akr

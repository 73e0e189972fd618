// akr_bsdf.cuh — surface closures of the hot path, evaluated in the local shading frame.
//
// Implements the closure tree the reference builds for a Principled BSDF
// (svm/surface/principled.rs:13-216) out of its combinators (svm/surface/mod.rs:330-695), the GGX
// microfacet lobes (mod.rs:820-1006, microfacet.rs:24-206), the Fresnel family (mod.rs:1009-1154),
// Lambert (diffuse.rs:13-80) and glass (glass.rs:13-45).
//
// Instead of a per-hit interpreter that rebuilds the tree (svm/eval.rs:364-380), every material is
// constant-folded once at scene upload into a `Material` record; lobes whose weight is exactly zero
// are pruned by `lobes` flags.  Pruning is value-preserving: a zero-weight lobe contributes
// f = 0 * finite = 0, its pdf is multiplied by a selection probability of exactly 0, and the
// selection remap (u - 0) / (1 - 0) returns u unchanged (mod.rs:486-522,606-639).
#pragma once
#include "akr_math.cuh"

namespace akr {

enum MaterialType : uint32_t {
    MAT_LAMBERT = 0,     // principled reduced to DiffuseBsdf(base_color / pi), or a diffuse node
    MAT_CONDUCTOR = 1,   // principled reduced to GGX reflection with FresnelComplex (metallic == 1, no coat)
    MAT_PRINCIPLED = 2,  // full tree
    MAT_GLASS = 3,       // glass node: Addictive(transmission, reflection)
    MAT_EMISSION = 4,    // emission node: no BSDF
};
enum MaterialLobes : uint32_t {
    LOBE_COAT = 1u << 0,          // coat_weight != 0
    LOBE_SPECULAR = 1u << 1,      // specular f0' != 0
    LOBE_DIFFUSE = 1u << 2,       // metallic < 1 - eps  &&  transmission < 1 - eps
    LOBE_TRANSMISSION = 1u << 3,  // metallic < 1 - eps  &&  transmission > eps
    LOBE_METAL = 1u << 4,         // metallic > eps
    LOBE_BASE = 1u << 5,          // metallic < 1 - eps (the non-metal branch is evaluated)
};

struct Material {  // 48 words
    uint32_t type;
    uint32_t lobes;
    float color[3];        // base colour (principled) / reflectance*pi (diffuse) / kr (glass)
    float alpha;           // Surface::alpha()
    float emission[3];     // emission_color * strength
    float diffuse[3];      // color / pi
    float trans_color[3];  // sqrt(color) (principled) / kt (glass)
    float metal_n[3], metal_k[3];
    float roughness;       // eval_float_auto_convert(roughness)
    float roughness_raw;   // eval_float(roughness) used by the dielectric lobes (principled.rs:103)
    float eta;             // ior
    float eta_s;           // specular layer ior after specular_ior_level (principled.rs:57-61)
    float f0;              // specular weight
    float spec_tint[3];
    float metallic, transmission;
    float coat_weight, coat_roughness, coat_ior;
    float coat_scale[3];   // lerp(1, coat_tint, coat_weight)  (principled.rs:190-193)
    float normal[3];       // normal input with x,y negated (principled.rs:200-202)
    uint32_t has_normal;   // normal != (0,0,0)
    uint32_t wrap_inner;   // 1 when the shader is a Principled node: it wraps itself in a second
                           // SurfaceClosure with the normal-map frame (principled.rs:208-214)
    uint32_t dynamic;      // some input depends on the hit (image texture / texture coordinates): the record holds the
                           // evaluation at uv = (0, 0) and is re-evaluated per hit from (shader_kind, data_offset), akr_svm.cuh
    uint32_t shader_kind, data_offset;  // ShaderRef (svm/mod.rs:213-219)
    uint32_t alpha_dynamic;  // `dynamic`, and alpha itself can differ from hit to hit (a texture with alpha != 1 texels or
                             // the zero address mode feeds the closure colour): only then traversal evaluates alpha per hit
    uint32_t static_offset;  // `dynamic`: first entry of this material's table of hit-independent node values (akr_svm.cuh)
    uint32_t _pad[1];
};
static_assert(sizeof(Material) == 192, "Material is 48 words");

// Shade classes: the trace stage bins every hit by the class of its material and one shade kernel is
// compiled per class, so a warp never interleaves Lambert, conductor and full-tree code (the
// "material-key sort" of the wavefront design).  CLS_ANY compiles the generic switch (host simulation).
enum ShadeClass : uint32_t { CLS_LAMBERT = 0, CLS_CONDUCTOR = 1, CLS_GENERAL = 2, CLS_COUNT = 3, CLS_ANY = 7 };
AKR_HD uint32_t shade_class_of(uint32_t material_type) {
    return material_type == MAT_LAMBERT ? (uint32_t)CLS_LAMBERT : (material_type == MAT_CONDUCTOR ? (uint32_t)CLS_CONDUCTOR : (uint32_t)CLS_GENERAL);
}
// a texture-driven material can be any type at any hit: it always takes the general kernel
AKR_HD uint32_t shade_class_of(const Material &m) { return m.dynamic ? (uint32_t)CLS_GENERAL : shade_class_of(m.type); }

struct BsdfEval {
    f3 f;
    float pdf;
};
struct BsdfDir {
    f3 wi;
    bool valid;
};

// ---- Trowbridge-Reitz (microfacet.rs), isotropic alpha pair kept as in the reference -------------
struct TR {
    float ax, ay;
};
AKR_HD TR tr_from_roughness(float r) {  // :24-43
    float a = fmaxf(sqr(r), 1e-4f);
    return TR{a, a};
}
AKR_HD float tr_lobe_roughness(float r) {  // roughness() = sqrt((ax + ay) / 2)
    float a = fmaxf(sqr(r), 1e-4f);
    return sqrtf((a + a) * 0.5f);
}
AKR_HD float tr_d(TR a, f3 wh) {  // :45-57
    float t2 = tan2_theta(wh);
    float c4 = sqr(cos2_theta(wh));
    float e = t2 * (sqr(cos_phi(wh) / a.ax) + sqr(sin_phi(wh) / a.ay));
    float inv_d = AKR_PI * a.ax * a.ay * c4 * sqr(1.0f + e);
    if (!is_finite(t2) || !is_finite(inv_d) || inv_d == 0.0f) return 0.0f;
    return 1.0f / inv_d;
}
AKR_HD float tr_lambda(TR a, f3 w) {  // :59-65
    float abs_tan = fabsf(tan_theta(w));
    float alpha2 = sqr(cos_phi(w)) * sqr(a.ax) + sqr(sin_phi(w)) * sqr(a.ay);
    float l = (-1.0f + sqrtf(1.0f + alpha2 * sqr(abs_tan))) * 0.5f;
    return !is_finite(abs_tan) ? 0.0f : l;
}
AKR_HD float tr_g1(TR a, f3 w) { return 1.0f / (1.0f + tr_lambda(a, w)); }
AKR_HD float tr_g(TR a, f3 wo, f3 wi) { return 1.0f / (1.0f + tr_lambda(a, wo) + tr_lambda(a, wi)); }
AKR_HD f3 tr_sample_wh_disk(TR a, f3 w, f2 p) {  // VNDF, :118-138, from the disk point p = uniform_sample_disk(u)
    f3 wh = normalize(mk3(a.ax * w.x, a.ay * w.y, w.z));
    if (wh.z < 0.0f) wh = -wh;
    f3 t1 = (wh.z < 0.99999f) ? normalize(cross(mk3(0, 0, 1), wh)) : mk3(1, 0, 0);
    f3 t2 = normalize(cross(wh, t1));
    float h = sqrtf(1.0f - sqr(p.x));
    p.y = lerpf(h, p.y, (1.0f + wh.z) * 0.5f);
    float pz = sqrtf(fmaxf(1.0f - (p.x * p.x + p.y * p.y), 0.0f));
    f3 nh = p.x * t1 + p.y * t2 + pz * wh;
    return normalize(mk3(a.ax * nh.x, a.ay * nh.y, fmaxf(nh.z, 1e-6f)));
}
AKR_HD f3 tr_sample_wh(TR a, f3 w, f2 u) { return tr_sample_wh_disk(a, w, uniform_sample_disk(u)); }
AKR_HD float tr_pdf(TR a, f3 wo, f3 wh) {  // :196-206 (sample_visible)
    return tr_d(a, wh) * tr_g1(a, wo) * fabsf(dot(wo, wh)) / abs_cos_theta(wo);
}

// ---- Fresnel (mod.rs:1009-1110) -----------------------------------------------------------------
AKR_HD float fr_dielectric(float cos_i, float eta) {
    cos_i = clampf(cos_i, -1.0f, 1.0f);
    eta = cos_i > 0.0f ? eta : 1.0f / eta;
    cos_i = fabsf(cos_i);
    float sin2_i = 1.0f - sqr(cos_i);
    float sin2_t = sin2_i / sqr(eta);
    if (sin2_t >= 1.0f) return 1.0f;
    float cos_t = sqrtf(fmaxf(1.0f - sin2_t, 0.0f));
    float r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
    return clampf((sqr(r_parl) + sqr(r_perp)) * 0.5f, 0.0f, 1.0f);
}
struct Cx {
    float re, im;
};
AKR_HD Cx cx_add(Cx a, Cx b) { return Cx{a.re + b.re, a.im + b.im}; }
AKR_HD Cx cx_sub(Cx a, Cx b) { return Cx{a.re - b.re, a.im - b.im}; }
AKR_HD Cx cx_mul(Cx a, Cx b) { return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
AKR_HD Cx cx_div(Cx a, Cx b) {
    float scale = 1.0f / (b.re * b.re + b.im * b.im);
    return Cx{(a.re * b.re + a.im * b.im) * scale, (a.im * b.re - a.re * b.im) * scale};
}
AKR_HD Cx cx_scale(Cx a, float s) { return Cx{a.re * s, a.im * s}; }
AKR_HD float cx_norm(Cx a) { return a.re * a.re + a.im * a.im; }
AKR_HD Cx cx_sqrt(Cx a) {  // util/mod.rs:541-554
    float n = sqrtf(cx_norm(a));
    float t1 = sqrtf(0.5f * (n + fabsf(a.re)));
    float t2 = 0.5f * a.im / t1;
    if (n == 0.0f) return Cx{0.0f, 0.0f};
    if (a.re >= 0.0f) return Cx{t1, t2};
    return Cx{fabsf(t2), copysignf(t1, a.im)};
}
AKR_HD float fr_complex(float cos_i, Cx eta) {  // mod.rs:1055-1067
    cos_i = clampf(cos_i, 0.0f, 0.999f);
    float sin2 = 1.0f - sqr(cos_i);
    Cx sin2_t = cx_div(Cx{sin2, 0.0f}, cx_mul(eta, eta));
    Cx cos_t = cx_sqrt(cx_sub(Cx{1.0f, 0.0f}, sin2_t));
    Cx eci = cx_scale(eta, cos_i);
    Cx r_parl = cx_div(cx_sub(eci, cos_t), cx_add(eci, cos_t));
    Cx ect = cx_mul(eta, cos_t);
    Cx r_perp = cx_div(cx_sub(Cx{cos_i, 0.0f}, ect), cx_add(Cx{cos_i, 0.0f}, ect));
    return (cx_norm(r_parl) + cx_norm(r_perp)) * 0.5f;
}
AKR_HD f3 fr_complex_spec(float cos_i, const float *n, const float *k) {  // mod.rs:1069-1081; FresnelComplex takes |cos|
    float c = fabsf(cos_i);
    return mk3(fr_complex(c, Cx{n[0], k[0]}), fr_complex(c, Cx{n[1], k[1]}), fr_complex(c, Cx{n[2], k[2]}));
}

// PreComputedTable::read_3d (mod.rs:1245-1322) on the 16^3 `ggx_dielectric_s` table
AKR_HD float table_read_1d(const float *buf, float x, uint32_t offset) {
    x = clampf(x, 0.0f, 1.0f) * 15.0f;
    uint32_t i = (uint32_t)floorf(x);
    uint32_t ni = i + 1 < 15u ? i + 1 : 15u;
    float t = x - (float)i;
    return (1.0f - t) * buf[offset + i] + t * buf[offset + ni];
}
AKR_HD float table_read_2d(const float *buf, float x, float y, uint32_t offset) {
    y = clampf(y, 0.0f, 1.0f) * 15.0f;
    uint32_t i = (uint32_t)floorf(y);
    uint32_t ni = i + 1 < 15u ? i + 1 : 15u;
    float t = y - (float)i;
    float d0 = table_read_1d(buf, x, offset + 16u * i);
    float d1 = table_read_1d(buf, x, offset + 16u * ni);
    return (1.0f - t) * d0 + t * d1;
}
AKR_HD float table_read_3d(const float *buf, float x, float y, float z) {
    z = clampf(z, 0.0f, 1.0f) * 15.0f;
    uint32_t i = (uint32_t)floorf(z);
    uint32_t ni = i + 1 < 15u ? i + 1 : 15u;
    float t = z - (float)i;
    float d0 = table_read_2d(buf, x, y, 256u * i);
    float d1 = table_read_2d(buf, x, y, 256u * ni);
    return (1.0f - t) * d0 + t * d1;
}
AKR_HD float ggx_dielectric_albedo_inl(const float *table, float roughness, float cos_i, float eta) {  // mod.rs:1145-1154
    float z = sqrtf(fabsf((eta - 1.0f) / (eta + 1.0f)));
    cos_i = fabsf(clampf(cos_i, -0.999f, 0.999f));
    return table_read_3d(table, roughness, fabsf(cos_i), z);
}

// The general-class tree looks the table up six times per bounce (coat and specular layer, at wo and wi, in evaluate and in
// sample_wi): one out-of-line copy (eight loads and a trilinear blend each; the general kernels are bound by instruction
// fetch).
AKR_HD_NOINLINE float ggx_dielectric_albedo(const float *table, float roughness, float cos_i, float eta) {
    return ggx_dielectric_albedo_inl(table, roughness, cos_i, eta);
}

// ---- lobes -------------------------------------------------------------------------------------
AKR_HD BsdfEval zero_eval() { return BsdfEval{splat3(0.0f), 0.0f}; }

AKR_HD BsdfEval diffuse_eval(f3 reflectance, f3 wo, f3 wi) {  // diffuse.rs:21-40
    bool same = same_hemisphere(wo, wi);
    float pdf = same ? abs_cos_theta(wi) * AKR_FRAC_1_PI : 0.0f;
    f3 c = same ? reflectance * abs_cos_theta(wi) : splat3(0.0f);
    return BsdfEval{c, pdf};
}
AKR_HD BsdfDir diffuse_sample(f3 wo, f2 u) {  // diffuse.rs:41-52
    f3 wi = cos_sample_hemisphere(u);
    wi = same_hemisphere(wo, wi) ? wi : -wi;
    return BsdfDir{wi, true};
}

// MicrofacetReflection::evaluate (mod.rs:831-858). FRESNEL: 0 = dielectric(eta), 1 = complex(n,k), 2 = complex(n,k) compact
template <int FRESNEL> AKR_HD BsdfEval mf_reflection_eval(f3 color, float eta, const float *n, const float *k, TR a, f3 wo, f3 wi) {
    f3 wh = wo + wi;
    float cos_o = cos_theta(wo), cos_i = cos_theta(wi);
    if ((dot(wh, wo) * dot(wi, wh)) < 0.0f || (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) || cos_i == 0.0f || cos_o == 0.0f ||
        !same_hemisphere(wo, wi))
        return zero_eval();
    wh = normalize(wh);
    float cf = dot(wi, face_forward(wh, mk3(0, 0, 1)));
    f3 fr;
    if (FRESNEL == 0) fr = splat3(1.0f) * fr_dielectric(cf, eta);
    else if (FRESNEL == 1) fr = fr_complex_spec(cf, n, k);
    else {  // the same three fr_complex calls through one copy of the code (general class)
        const float c = fabsf(cf);
        fr = splat3(0.0f);
        AKR_NO_UNROLL
        for (int ch = 0; ch < 3; ++ch) {
            const float v = fr_complex(c, Cx{n[ch], k[ch]});
            if (ch == 0) fr.x = v;
            else if (ch == 1) fr.y = v;
            else fr.z = v;
        }
    }
    float d = tr_d(a, wh);
    float g = tr_g(a, wo, wi);
    f3 f = color * fr * fabsf(0.25f * d * g / (cos_i * cos_o)) * fabsf(cos_i);
    float pdf = tr_pdf(a, wo, wh) / (4.0f * fabsf(dot(wo, wh)));
    return BsdfEval{f, pdf};
}
// GGX reflection with dielectric Fresnel: dielectric mixture, specular layer and clearcoat of the general tree share one
// out-of-line copy
AKR_HD_NOINLINE BsdfEval mf_reflection_eval_dielectric(f3 color, float eta, float rough, f3 wo, f3 wi) {
    return mf_reflection_eval<0>(color, eta, nullptr, nullptr, tr_from_roughness(rough), wo, wi);
}
AKR_HD BsdfDir mf_reflection_sample(TR a, f3 wo, f2 u) {  // mod.rs:861-873
    f3 wh = tr_sample_wh(a, wo, u);
    f3 wi = reflect(wo, wh);
    return BsdfDir{wi, same_hemisphere(wo, wi)};
}
// MicrofacetTransmission (mod.rs:914-979)
AKR_HD BsdfEval mf_transmission_eval(f3 color, float eta_in, TR a, f3 wo, f3 wi) {
    float cos_o = cos_theta(wo), cos_i = cos_theta(wi);
    float eta = cos_o > 0.0f ? eta_in : 1.0f / eta_in;
    f3 wh = normalize(wo + wi * eta);
    wh = face_forward(wh, mk3(0, 0, 1));
    bool backfacing = (dot(wh, wi) * cos_i) < 0.0f || (dot(wh, wo) * cos_o) < 0.0f;
    if ((dot(wh, wo) * dot(wi, wh)) > 0.0f || cos_i == 0.0f || cos_o == 0.0f || backfacing || same_hemisphere(wo, wi)) return zero_eval();
    f3 f;
    {
        f3 fr = splat3(1.0f) * fr_dielectric(dot(wo, wh), eta_in);
        float denom = sqr(dot(wi, wh) + dot(wo, wh) / eta) * cos_i * cos_o;
        if (denom == 0.0f) f = splat3(0.0f);
        else
            f = (splat3(1.0f) - fr) * color *
                fabsf(tr_d(a, wh) * tr_g(a, wo, wi) / sqr(eta) * fabsf(dot(wi, wh)) * fabsf(dot(wo, wh)) / denom) * fabsf(cos_i);
    }
    float pdf;
    {
        float denom = sqr(dot(wi, wh) + dot(wo, wh) / eta);
        float dwh_dwi = fabsf(dot(wi, wh)) / denom;
        pdf = denom == 0.0f ? 0.0f : tr_pdf(a, wo, wh) * dwh_dwi;
    }
    return BsdfEval{f, pdf};
}
AKR_HD BsdfDir mf_refract_about(float eta_in, f3 wh, f3 wo) {  // geometry.rs:284-303 on the sampled microfacet normal
    f3 n = wh;
    float cos_i = dot(wo, n);
    float eta = cos_i >= 0.0f ? eta_in : 1.0f / eta_in;
    n = cos_i >= 0.0f ? n : -n;
    cos_i = fabsf(cos_i);
    float sin2_i = fmaxf(1.0f - sqr(cos_i), 0.0f);
    float sin2_t = sin2_i / sqr(eta);
    if (sin2_t >= 1.0f) return BsdfDir{splat3(0.0f), false};
    float cos_t = sqrtf(1.0f - sin2_t);
    f3 wt = -wo / eta + (cos_i / eta - cos_t) * n;
    return BsdfDir{wt, !same_hemisphere(wo, wt)};
}
AKR_HD BsdfDir mf_transmission_sample(float eta_in, TR a, f3 wo, f2 u) {  // mod.rs:981-993
    return mf_refract_about(eta_in, tr_sample_wh(a, wo, u), wo);
}

// weighted_discrete_choice2_and_remap (sampling.rs:61-70): returns true for the first option
AKR_HD bool choose2(float weight_a, float &u) {
    bool first = u < weight_a;
    u = first ? u / weight_a : (u - weight_a) / (1.0f - weight_a);
    return first;
}

// dielectric = BsdfMixture Addictive(transmission, reflection), frac = fr_dielectric(cos wo, eta)
// (principled.rs:99-130, glass.rs:13-45)
AKR_HD BsdfEval dielectric_eval(f3 kr, f3 kt, float eta, float rough, f3 wo, f3 wi) {
    TR a = tr_from_roughness(rough);
    float frac = fr_dielectric(cos_theta(wo), eta);
    BsdfEval ea = mf_transmission_eval(kt, eta, a, wo, wi);
    BsdfEval eb = mf_reflection_eval<0>(kr, eta, nullptr, nullptr, a, wo, wi);
    return BsdfEval{ea.f + eb.f, lerpf(ea.pdf, eb.pdf, frac)};
}

// ---- the Principled tree in the material-local frame (principled.rs:143-199): general_eval / general_sample below ----
AKR_HD f3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }

// SurfaceClosure::check_wo_wi_valid (mod.rs:706-718)
AKR_HD bool check_wo_wi_valid(f3 ns, f3 ng, f3 wo, f3 wi) {
    float flipped = dot(ng, ns) > 0.0f ? 1.0f : -1.0f;
    float so = ((flipped * dot(wo, ns)) > 0.0f ? 1.0f : -1.0f) * (dot(wo, ng) > 0.0f ? 1.0f : -1.0f);
    float si = ((flipped * dot(wi, ns)) > 0.0f ? 1.0f : -1.0f) * (dot(wi, ng) > 0.0f ? 1.0f : -1.0f);
    return (so > 0.0f) && (si > 0.0f);
}

// One cell of the 16^3 `ggx_dielectric_s` table (svm/surface/precompute.rs:56-94; mod.rs:1338-1356):
// E[f / pdf] of a white GGX reflection lobe with dielectric Fresnel, as a function of
// (roughness, mu = cos theta_o, z = ior parametrisation).  The reference averages 2^20 PCG32 samples
// seeded from rand::StdRng (not reproducible); this is the mean over a fixed n x n midpoint grid.
AKR_HD float albedo_table_cell(uint32_t cell, uint32_t n) {
    uint32_t ix = cell & 15u, iy = (cell >> 4) & 15u, iz = cell >> 8;
    float roughness = clampf((float)ix / 15.0f, 1e-4f, 0.9999f);
    float mu = clampf((float)iy / 15.0f, 1e-4f, 0.9999f);
    float fz = clampf((float)iz / 15.0f, 1e-4f, 0.9999f);
    float sf0 = sqrtf(clampf(sqr(sqr(fz)), 0.0f, 0.99f));  // ior_parametrization -> ior_from_f0 (mod.rs:1090-1103)
    float ior = (1.0f + sf0) / (1.0f - sf0);
    TR a = tr_from_roughness(roughness);
    Frame fr = frame_from_n(mk3(0, 0, 1));
    f3 ng = mk3(0, 0, 1);
    f3 wo = mk3(sqrtf(1.0f - sqr(mu)), 0.0f, mu);
    f3 wo_l = to_local(fr, wo);
    double sum = 0.0;
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) {
            f2 u = f2{((float)i + 0.5f) / (float)n, ((float)j + 0.5f) / (float)n};
            BsdfDir s = mf_reflection_sample(a, wo_l, u);
            f3 wi = to_world(fr, s.wi);
            if (!(s.valid && check_wo_wi_valid(fr.n, ng, wo, wi))) continue;
            if (!check_wo_wi_valid(fr.n, ng, wo, wi)) continue;
            BsdfEval e = mf_reflection_eval<0>(splat3(1.0f), ior, nullptr, nullptr, a, wo_l, to_local(fr, wi));
            if (e.pdf > 0.0f) sum += (double)(e.f.x / e.pdf);
        }
    return (float)(sum / ((double)n * (double)n));
}

// ---- material dispatch in the material-local frame -----------------------------------------------
// The general class (every material type in one kernel) keeps ONE copy of each building block: the leaves a material
// needs are evaluated first — diffuse, dielectric (GGX transmission + reflection), metal (GGX with complex Fresnel) —
// and then combined by type.  The arithmetic per leaf and per combination is the reference's, operation for operation;
// what changes against a per-type switch of fully inlined trees is the code size (the general kernels were 19 K
// instructions, a third of their issue slots went to instruction fetch), not a single result bit.
AKR_HD BsdfEval general_eval(const Material &m, const float *table, f3 wo, f3 wi) {
    const float EPS = 1e-4f;  // BsdfMixture::EPS
    const uint32_t type = m.type;
    const bool is_p = type == MAT_PRINCIPLED;
    const float mt = m.metallic, tr = m.transmission;
    const bool base = is_p && mt < 1.0f - EPS;  // bsdf1 takes part in Mix(bsdf1, metal, metallic)   (principled.rs:131-142,170-175)
    BsdfEval e_diff = zero_eval(), e_diel = zero_eval(), e_metal = zero_eval();
    if (type == MAT_LAMBERT || (base && tr < 1.0f - EPS)) e_diff = diffuse_eval(ld3(m.diffuse), wo, wi);
    if (type == MAT_GLASS || (base && tr > EPS)) {  // dielectric_eval with the shared reflection lobe
        const float frac = fr_dielectric(cos_theta(wo), m.eta);
        const BsdfEval et = mf_transmission_eval(ld3(m.trans_color), m.eta, tr_from_roughness(m.roughness_raw), wo, wi);
        const BsdfEval er = mf_reflection_eval_dielectric(ld3(m.color), m.eta, m.roughness_raw, wo, wi);
        e_diel = BsdfEval{et.f + er.f, lerpf(et.pdf, er.pdf, frac)};
    }
    if (type == MAT_CONDUCTOR || (is_p && mt > EPS))
        e_metal = mf_reflection_eval<2>(splat3(1.0f), 0.0f, m.metal_n, m.metal_k, tr_from_roughness(m.roughness), wo, wi);
    if (type == MAT_LAMBERT) return e_diff;
    if (type == MAT_CONDUCTOR) return e_metal;
    if (type == MAT_GLASS) return e_diel;
    if (!is_p) return zero_eval();
    // bsdf0 = Mix(diffuse, dielectric, transmission)   (principled.rs:143-148; mod.rs:606-619)
    BsdfEval ea = zero_eval();
    if (base) {
        ea = BsdfEval{lerp3(e_diff.f, e_diel.f, tr), lerpf(e_diff.pdf, e_diel.pdf, tr)};
        if (m.lobes & LOBE_SPECULAR) {  // bsdf1 = Coated(top = specular GGX, bottom = bsdf0, e_top)   (principled.rs:55-80,151-168; mod.rs:486-503)
            BsdfEval top = mf_reflection_eval_dielectric(ld3(m.spec_tint) * m.f0, m.eta_s, m.roughness, wo, wi);
            f3 tint = ld3(m.spec_tint);
            f3 eo = tint * ggx_dielectric_albedo(table, m.roughness, abs_cos_theta(wo), m.eta_s) * m.f0;
            f3 ei = tint * ggx_dielectric_albedo(table, m.roughness, abs_cos_theta(wi), m.eta_s) * m.f0;
            float p_top = avg3(eo);
            float pdf = top.pdf * p_top + ea.pdf * (1.0f - p_top);
            f3 f = top.f + ea.f * min3(splat3(1.0f) - eo, splat3(1.0f) - ei);
            ea = BsdfEval{f, pdf};
        }
    }
    // bsdf2 = Mix(bsdf1, metal, metallic); EmissiveSurface: pass-through; ScaledBsdf(lerp(1, coat_tint, coat_weight))  (principled.rs:178-193)
    BsdfEval e2 = BsdfEval{lerp3(ea.f, e_metal.f, mt), lerpf(ea.pdf, e_metal.pdf, mt)};
    BsdfEval scaled = BsdfEval{e2.f * ld3(m.coat_scale), e2.pdf};
    if (!(m.lobes & LOBE_COAT)) return scaled;
    // bsdf4 = Coated(top = clearcoat GGX, bottom = scaled, e_top)   (principled.rs:81-98,183-199)
    BsdfEval top = mf_reflection_eval_dielectric(splat3(1.0f) * m.coat_weight, m.coat_ior, m.coat_roughness, wo, wi);
    f3 eo = splat3(1.0f) * m.coat_weight * ggx_dielectric_albedo(table, m.coat_roughness, abs_cos_theta(wo), m.coat_ior);
    f3 ei = splat3(1.0f) * m.coat_weight * ggx_dielectric_albedo(table, m.coat_roughness, abs_cos_theta(wi), m.coat_ior);
    float p_top = avg3(eo);
    float pdf = top.pdf * p_top + scaled.pdf * (1.0f - p_top);
    f3 f = top.f + scaled.f * min3(splat3(1.0f) - eo, splat3(1.0f) - ei);
    return BsdfEval{f, pdf};
}
// sample_wi of any material: walk the tree to ONE lobe (the choose2 sequence of BsdfMixture / CoatedBsdf::sample_wi,
// mod.rs:504-535,620-640), then run one copy of the visible-normal sampling for whichever GGX lobe was picked.
AKR_HD BsdfDir general_sample(const Material &m, const float *table, f3 wo, float u_select, f2 u) {
    enum { PICK_NONE, PICK_DIFFUSE, PICK_REFLECT, PICK_DIELECTRIC };
    int pick = PICK_NONE;
    float rough = 0.0f;
    switch (m.type) {
    case MAT_LAMBERT: pick = PICK_DIFFUSE; break;
    case MAT_CONDUCTOR: pick = PICK_REFLECT; rough = m.roughness; break;
    case MAT_GLASS: pick = PICK_DIELECTRIC; break;
    case MAT_PRINCIPLED: {
        if (m.lobes & LOBE_COAT) {
            f3 eo = splat3(1.0f) * m.coat_weight * ggx_dielectric_albedo(table, m.coat_roughness, abs_cos_theta(wo), m.coat_ior);
            if (choose2(avg3(eo), u_select)) { pick = PICK_REFLECT; rough = m.coat_roughness; break; }
        }
        // Mix(bsdf1, metal, metallic): choice(frac, 1, 0) -> first = metal
        if (choose2(m.metallic, u_select)) { pick = PICK_REFLECT; rough = m.roughness; break; }
        if (m.lobes & LOBE_SPECULAR) {
            f3 eo = ld3(m.spec_tint) * ggx_dielectric_albedo(table, m.roughness, abs_cos_theta(wo), m.eta_s) * m.f0;
            if (choose2(avg3(eo), u_select)) { pick = PICK_REFLECT; rough = m.roughness; break; }
        }  // else: choice with weight 0 never picks the top lobe and remaps u to (u - 0) / (1 - 0) == u
        // Mix(diffuse, dielectric, transmission): choice(frac, 1, 0) -> first = bsdf_b (dielectric)
        pick = choose2(m.transmission, u_select) ? PICK_DIELECTRIC : PICK_DIFFUSE;
        break;
    }
    default: break;
    }
    if (pick == PICK_NONE) return BsdfDir{splat3(0.0f), false};
    const f2 disk = uniform_sample_disk(u);  // the cosine-hemisphere and the visible-normal sampling both start from it: one sincos
    if (pick == PICK_DIFFUSE) {  // diffuse_sample
        f3 wi = cos_hemisphere_from_disk(disk);
        return BsdfDir{same_hemisphere(wo, wi) ? wi : -wi, true};
    }
    bool refract = false;
    if (pick == PICK_DIELECTRIC) {  // dielectric_sample: choice(frac, 1, 0): first -> reflection
        rough = m.roughness_raw;
        float frac = fr_dielectric(cos_theta(wo), m.eta);
        refract = !choose2(frac, u_select);
    }
    f3 wh = tr_sample_wh_disk(tr_from_roughness(rough), wo, disk);
    if (refract) return mf_refract_about(m.eta, wh, wo);
    f3 wi = reflect(wo, wh);
    return BsdfDir{wi, same_hemisphere(wo, wi)};
}
template <int CLS> AKR_HD BsdfEval material_eval(const Material &m, const float *table, f3 wo, f3 wi) {
    if (CLS == CLS_LAMBERT) return diffuse_eval(ld3(m.diffuse), wo, wi);
    if (CLS == CLS_CONDUCTOR) return mf_reflection_eval<1>(splat3(1.0f), 0.0f, m.metal_n, m.metal_k, tr_from_roughness(m.roughness), wo, wi);
    return general_eval(m, table, wo, wi);
}
template <int CLS> AKR_HD BsdfDir material_sample(const Material &m, const float *table, f3 wo, float u_select, f2 u) {
    if (CLS == CLS_LAMBERT) return diffuse_sample(wo, u);
    if (CLS == CLS_CONDUCTOR) return mf_reflection_sample(tr_from_roughness(m.roughness), wo, u);
    return general_sample(m, table, wo, u_select, u);
}

// The two nested SurfaceClosures the reference wraps around a surface shader:
//   outer: frame = si.frame, ng = si.ng                       (svm/eval.rs:488-492)
//   inner (principled only): frame = normal_map(...), ng = si.frame.to_local(si.ng)   (principled.rs:200-214, mod.rs:1380-1417)
struct ClosureFrames {
    Frame outer;
    f3 ng;          // world
    Frame inner;    // expressed in outer-local coordinates
    f3 ng_local;    // to_local(outer, ng)
    bool has_inner; // principled materials
};
AKR_HD Frame normal_map_frame(f3 normal, const Frame &frame) {  // mod.rs:1390-1410
    f3 n_world = to_world(frame, normalize(normal));
    Frame nf = frame_from_n_t(n_world, frame.t);
    return Frame{to_local(frame, nf.n), to_local(frame, nf.t), to_local(frame, nf.s)};
}
AKR_HD ClosureFrames make_closure_frames(const Material &m, const Frame &frame, f3 ng) {
    ClosureFrames c;
    c.outer = frame;
    c.ng = ng;
    c.has_inner = m.wrap_inner != 0;
    c.ng_local = to_local(frame, ng);
    c.inner = m.has_normal ? normal_map_frame(ld3(m.normal), frame) : frame_identity();
    return c;
}
// SurfaceClosure::evaluate_impl applied twice (mod.rs:729-748)
template <int CLS> AKR_HD BsdfEval closure_eval(const Material &m, const float *table, const ClosureFrames &c, f3 wo, f3 wi) {
    if (!check_wo_wi_valid(c.outer.n, c.ng, wo, wi)) return zero_eval();
    f3 wo_l = to_local(c.outer, wo), wi_l = to_local(c.outer, wi);
    if (c.has_inner) {
        if (!check_wo_wi_valid(c.inner.n, c.ng_local, wo_l, wi_l)) return zero_eval();
        wo_l = to_local(c.inner, wo_l);
        wi_l = to_local(c.inner, wi_l);
    }
    return material_eval<CLS>(m, table, wo_l, wi_l);
}
// SurfaceClosure::sample_wi_impl applied twice (mod.rs:750-764)
template <int CLS> AKR_HD BsdfDir closure_sample_wi(const Material &m, const float *table, const ClosureFrames &c, f3 wo, float u_select, f2 u) {
    f3 wo_l = to_local(c.outer, wo);
    BsdfDir s;
    if (c.has_inner) {
        f3 wo_ll = to_local(c.inner, wo_l);
        s = material_sample<CLS>(m, table, wo_ll, u_select, u);
        f3 wi_l = to_world(c.inner, s.wi);
        s.valid = s.valid && check_wo_wi_valid(c.inner.n, c.ng_local, wo_l, wi_l);
        s.wi = wi_l;
    } else {
        s = material_sample<CLS>(m, table, wo_l, u_select, u);
    }
    f3 wi = to_world(c.outer, s.wi);
    bool valid = s.valid && check_wo_wi_valid(c.outer.n, c.ng, wo, wi);
    return BsdfDir{wi, valid};
}

// ---- Surface::albedo / emission / roughness / ns (the `aov` integrator's taps, aov.rs:96-155) ---------------------
// albedo() + emission(): PrincipledBsdfWrapper overrides both with base_color / emission (principled.rs:227-234,267-274);
// diffuse node: reflectance * pi (diffuse.rs:56-63); glass node: Addictive mixture a.albedo + b.albedo = kt + kr
// (mod.rs:659-675,875-882,981-988); emission node: 0 + emission (mod.rs:372-383,399-410).
AKR_HD f3 material_albedo_plus_emission(const Material &m) {
    if (m.wrap_inner) return ld3(m.color) + ld3(m.emission);
    switch (m.type) {
    case MAT_LAMBERT: return ld3(m.diffuse) * AKR_PI + splat3(0.0f);
    case MAT_GLASS: return (ld3(m.trans_color) + ld3(m.color)) + (splat3(0.0f) + splat3(0.0f));
    case MAT_EMISSION: return splat3(0.0f) + ld3(m.emission);
    default: return splat3(0.0f);
    }
}
// roughness(wo, u_select): walks the tree like sample_wi does and returns the roughness of the lobe it lands on
// (mod.rs:536-553 Coated, :641-657 Mixture, :883-891 / :989-997 microfacet lobes, diffuse.rs:64-72 = 1).
AKR_HD float material_roughness(const Material &m, const float *table, f3 wo, float u_select) {
    if (!m.wrap_inner) {
        if (m.type == MAT_GLASS) return tr_lobe_roughness(m.roughness_raw);  // both lobes of the mixture share the distribution
        return 1.0f;                                                          // diffuse node, emission node (inner = None)
    }
    if (m.lobes & LOBE_COAT) {
        f3 eo = splat3(1.0f) * m.coat_weight * ggx_dielectric_albedo(table, m.coat_roughness, abs_cos_theta(wo), m.coat_ior);
        if (choose2(avg3(eo), u_select)) return tr_lobe_roughness(m.coat_roughness);
    }
    if (choose2(m.metallic, u_select)) return tr_lobe_roughness(m.roughness);  // Mix(bsdf1, metal): first = metal
    if (m.lobes & LOBE_SPECULAR) {
        f3 eo = ld3(m.spec_tint) * ggx_dielectric_albedo(table, m.roughness, abs_cos_theta(wo), m.eta_s) * m.f0;
        if (choose2(avg3(eo), u_select)) return tr_lobe_roughness(m.roughness);
    }
    if (choose2(m.transmission, u_select)) return tr_lobe_roughness(m.roughness_raw);  // dielectric: either lobe, same distribution
    return 1.0f;                                                                         // diffuse
}
AKR_HD float closure_roughness(const Material &m, const float *table, const ClosureFrames &c, f3 wo, float u_select) {  // mod.rs:774-783, twice
    f3 wo_l = to_local(c.outer, wo);
    if (c.has_inner) wo_l = to_local(c.inner, wo_l);
    return material_roughness(m, table, wo_l, u_select);
}
AKR_HD f3 closure_ns(const ClosureFrames &c) {  // SurfaceClosure::ns (mod.rs:724-727), innermost ns = (0, 0, 1)
    f3 ns = mk3(0.0f, 0.0f, 1.0f);
    if (c.has_inner) ns = to_world(c.inner, ns);
    return to_world(c.outer, ns);
}

}  // namespace akr

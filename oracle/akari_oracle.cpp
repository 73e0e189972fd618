// ============================================================================================
// akari_oracle.cpp — CPU restatement of AkariRender's unidirectional path tracer hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (akari_render_b200/, include/) may import,
// link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / the CPU baseline.
//
// PARITY STATUS: "parity unpinned".  The reference cannot be built or run in this image (no Rust
// toolchain; `luisa_compute` is an un-vendored path dependency: crates/akari_common/Cargo.toml:26,
// Cargo.lock:1250) and it ships no golden images or known-answer tests for the renderer
// (SURVEY.md §4, §8c).  This file therefore follows the reference source line by line (citations
// below, all relative to /root/reference/crates) and is pinned only by
//   * the two static sampler tables (bit-exact inputs),
//   * the reference's own unit-test properties (alias table, pow-4 helpers),
//   * analytic checks (white furnace, NEE on/off agreement, chi-square BSDF tests).
// Third-party pieces that are NOT in the tree are restated from their public definitions and
// flagged ASSUMED where they appear: ray/triangle intersection + traversal (OptiX/Embree behind
// luisa rtx::Accel), `offset_ray_origin` (luisa; Ray Tracing Gems ch. 6), `TriangleInterpolate`,
// `lerp`, `normalize`, float atomics.
//
// Deliberately scalar and literal: same operation order as the cited lines; compiled with
// -ffp-contract=off so every f32 operation is a single IEEE operation (Rust never contracts).
// ============================================================================================
#include "../include/akari_b200.h"
#include "chi2_tables.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

// --------------------------------------------------------------------------------------------
// small vector types
// --------------------------------------------------------------------------------------------
struct V2 {
    float x, y;
};
struct V3 {
    float x, y, z;
};
using Color = V3;  // Color::Rgb(Float3, SRgb) — spectral is todo!() in the reference (color.rs:71-73)

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3s(float s) { return V3{s, s, s}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length_squared(V3 a) { return dot(a, a); }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
// ASSUMED (luisa `normalize`): v * (1 / sqrt(dot(v, v))) — IEEE sqrt and divide on both sides.
inline V3 normalize(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline float sqr(float x) { return x * x; }
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
// ASSUMED (luisa `lerp(a, b, t)`): t * (b - a) + a
inline float lerpf(float a, float b, float t) { return t * (b - a) + a; }
inline V3 lerp3(V3 a, V3 b, V3 t) { return {lerpf(a.x, b.x, t.x), lerpf(a.y, b.y, t.y), lerpf(a.z, b.z, t.z)}; }
inline float reduce_max(V3 a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }
inline float reduce_min(V3 a) { return std::fmin(a.x, std::fmin(a.y, a.z)); }
inline bool has_nan(V3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
inline V3 min3(V3 a, V3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline float avg(V3 a) { return (a.x + a.y + a.z) / 3.0f; }  // color.rs:237-242
inline uint32_t f2u(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

constexpr float PI = 3.14159265358979323846f;
constexpr float FRAC_1_PI = 0.318309886183790671537767526745028724f;
constexpr float ONE_MINUS_EPSILON = 0x1.fffffep-1f;  // lib.rs:59

// util/mod.rs:326-331
inline float difference_of_products(float a, float b, float c, float d) {
    float cd = c * d;
    float diff = std::fma(a, b, -cd);
    float err = std::fma(-c, d, cd);
    return diff + err;
}

// --------------------------------------------------------------------------------------------
// geometry.rs
// --------------------------------------------------------------------------------------------
struct Frame {  // geometry.rs:72-76
    V3 n, t, s;
};
struct FrameFn {
    static float cos_theta(V3 w) { return w.z; }
    static float cos2_theta(V3 w) { return w.z * w.z; }
    static float abs_cos_theta(V3 w) { return std::fabs(w.z); }
    static float sin2_theta(V3 w) { return std::fmax(1.0f - cos2_theta(w), 0.0f); }
    static float sin_theta(V3 w) { return std::sqrt(std::fmax(1.0f - cos2_theta(w), 0.0f)); }
    static float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
    static float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
    // geometry.rs:122-152 — note sin_phi uses w.x and cos_phi uses w.y (as written in the reference)
    static float sin_phi(V3 w) {
        float st = sin_theta(w);
        return st == 0.0f ? 0.0f : clampf(w.x / st, -1.0f, 1.0f);
    }
    static float cos_phi(V3 w) {
        float st = sin_theta(w);
        return st == 0.0f ? 1.0f : clampf(w.y / st, -1.0f, 1.0f);
    }
    static float sin2_phi(V3 w) { return sqr(sin_phi(w)); }
    static float cos2_phi(V3 w) { return sqr(cos_phi(w)); }
    static bool same_hemisphere(V3 a, V3 b) { return (a.z * b.z) >= 0.0f; }
};
inline Frame frame_identity() { return {v3(0, 0, 1), v3(1, 0, 0), v3(0, 1, 0)}; }  // geometry.rs:155-157
inline Frame frame_from_n(V3 n) {                                                  // geometry.rs:159-167
    V3 t;
    if (std::fabs(n.x) > std::fabs(n.y)) t = v3(-n.z, 0.0f, n.x) / std::sqrt(n.x * n.x + n.z * n.z);
    else t = v3(0.0f, n.z, -n.y) / std::sqrt(n.y * n.y + n.z * n.z);
    V3 s = cross(n, t);
    return {n, t, s};
}
inline Frame frame_from_n_t(V3 n, V3 tt_in) {  // geometry.rs:169-191
    Frame frame{};
    V3 tt = tt_in - n * dot(n, tt_in);
    bool good = true;
    if (length(tt) < 1e-4f) good = false;
    else tt = normalize(tt);
    if (good) {
        V3 ss = cross(n, tt);
        if (length(ss) < 1e-4f) good = false;
        else {
            ss = normalize(ss);
            frame = {n, tt, ss};
        }
    }
    if (!good) frame = frame_from_n(n);
    return frame;
}
inline V3 to_world(const Frame &f, V3 v) { return f.t * v.x + f.s * v.y + f.n * v.z; }  // geometry.rs:193-195
inline V3 to_local(const Frame &f, V3 v) { return v3(dot(f.t, v), dot(f.s, v), dot(f.n, v)); }
inline V3 face_forward(V3 v, V3 n) { return dot(v, n) < 0.0f ? -v : v; }   // geometry.rs:264-272
inline V3 reflect(V3 w, V3 n) { return -w + 2.0f * dot(w, n) * n; }        // geometry.rs:277-281
struct Refract {
    bool ok;
    float eta;
    V3 wt;
};
inline Refract refract(V3 w, V3 n, float eta) {  // geometry.rs:284-303
    float cos_theta_i = dot(w, n);
    eta = cos_theta_i >= 0.0f ? eta : 1.0f / eta;
    n = cos_theta_i >= 0.0f ? n : -n;
    cos_theta_i = std::fabs(cos_theta_i);
    float sin2_theta_i = std::fmax(1.0f - sqr(cos_theta_i), 0.0f);
    float sin2_theta_t = sin2_theta_i / sqr(eta);
    if (sin2_theta_t >= 1.0f) return {false, eta, v3s(0.0f)};
    float cos_theta_t = std::sqrt(1.0f - sin2_theta_t);
    V3 wt = -w / eta + (cos_theta_i / eta - cos_theta_t) * n;
    return {true, eta, wt};
}

// ASSUMED: luisa::rtx::offset_ray_origin — Waechter & Binder, Ray Tracing Gems ch. 6
// (call sites pt.rs:856, light/area.rs:87).  Source is not in the tree.
inline V3 offset_ray_origin(V3 p, V3 n) {
    constexpr float origin = 1.0f / 32.0f;
    constexpr float float_scale = 1.0f / 65536.0f;
    constexpr float int_scale = 256.0f;
    int32_t ofx = static_cast<int32_t>(int_scale * n.x);
    int32_t ofy = static_cast<int32_t>(int_scale * n.y);
    int32_t ofz = static_cast<int32_t>(int_scale * n.z);
    auto shift = [](float p, int32_t of) {
        int32_t i = static_cast<int32_t>(f2u(p));
        i += (p < 0.0f) ? -of : of;
        return u2f(static_cast<uint32_t>(i));
    };
    V3 pi = v3(shift(p.x, ofx), shift(p.y, ofy), shift(p.z, ofz));
    return v3(std::fabs(p.x) < origin ? p.x + float_scale * n.x : pi.x,
              std::fabs(p.y) < origin ? p.y + float_scale * n.y : pi.y,
              std::fabs(p.z) < origin ? p.z + float_scale * n.z : pi.z);
}

struct Ray {  // geometry.rs:18-25
    V3 o, d;
    float t_min, t_max;
    uint32_t ex0_inst, ex0_prim, ex1_inst, ex1_prim;
};

// --------------------------------------------------------------------------------------------
// util/hash.rs:44-59, sampler/mod.rs
// --------------------------------------------------------------------------------------------
inline uint32_t rotl17(uint32_t h) { return (h << 17) | (h >> (32 - 17)); }
inline uint32_t xxhash32_4(uint32_t px, uint32_t py, uint32_t pz, uint32_t pw) {
    const uint32_t PRIME32_2 = 2246822519u, PRIME32_3 = 3266489917u, PRIME32_4 = 668265263u, PRIME32_5 = 374761393u;
    uint32_t h32 = pw + PRIME32_5 + px * PRIME32_3;
    h32 = PRIME32_4 * rotl17(h32);
    h32 = h32 + py * PRIME32_3;
    h32 = PRIME32_4 * rotl17(h32);
    h32 = h32 + pz * PRIME32_3;
    h32 = PRIME32_4 * rotl17(h32);
    h32 = PRIME32_2 * (h32 ^ (h32 >> 15));
    h32 = PRIME32_3 * (h32 ^ (h32 >> 13));
    return h32 ^ (h32 >> 16);
}

// sampler/mod.rs:473-505 (Kensler's permutation, rejection loop)
inline uint32_t permute_element(uint32_t i, uint32_t l, uint32_t w, uint32_t p) {
    do {
        i ^= p;
        i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;
        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;
        i *= 1 | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11;
        i *= 0x74dcb303u;
        i ^= (i & w) >> 2;
        i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;
        i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}

struct Tables {
    const uint32_t *pmj;  // [5][65536][2]
    const uint16_t *bn;   // [48][128][128]
};

struct Pmj02BnSampler {  // sampler/mod.rs:513-520 state + :551-669 behaviour
    Tables tab;
    uint32_t seed, dim, px, py, sample_index, spp, w;

    // texel read at uv = p.yx() % 128 of an R16 unorm texture stored row-major (sampler/mod.rs:421-433,542-550)
    float bluenoise(uint32_t tex_index, uint32_t x, uint32_t y) const {
        uint32_t ux = y % AKR_BLUE_NOISE_RESOLUTION;  // uv.x = p.y
        uint32_t uy = x % AKR_BLUE_NOISE_RESOLUTION;  // uv.y = p.x
        uint32_t t = tex_index % AKR_BLUE_NOISE_TEXTURES;
        uint16_t v = tab.bn[(static_cast<size_t>(t) * AKR_BLUE_NOISE_RESOLUTION + uy) * AKR_BLUE_NOISE_RESOLUTION + ux];
        return static_cast<float>(v) / 65535.0f;
    }
    V2 pmj02bn_sample(uint32_t set_index, uint32_t si) const {  // sampler/mod.rs:353-368
        set_index %= AKR_PMJ02BN_SETS;
        si %= AKR_PMJ02BN_SAMPLES;
        uint32_t i = AKR_PMJ02BN_SAMPLES * set_index + si;
        return {static_cast<float>(tab.pmj[i * 2]) * 0x1p-32f, static_cast<float>(tab.pmj[i * 2 + 1]) * 0x1p-32f};
    }
    void start() {  // :656-669
        dim = 4;
        if (sample_index == UINT32_MAX) sample_index = 0;
        else sample_index += 1;
    }
    float next_1d() {  // :555-574
        uint32_t hash = xxhash32_4(px, py, dim, seed);
        uint32_t index = permute_element(sample_index, spp, w, hash);
        float delta = bluenoise(dim, px, py);
        dim += 1;
        return std::fmin((static_cast<float>(index) + delta) / static_cast<float>(spp), ONE_MINUS_EPSILON);
    }
    V2 next_2d() {  // :583-620
        uint32_t index = sample_index;
        uint32_t d = dim;
        uint32_t pmj_instance = d / 2;
        if (pmj_instance >= AKR_PMJ02BN_SETS) {
            uint32_t hash = xxhash32_4(px, py, dim, seed);
            index = permute_element(sample_index, spp, w, hash);
        }
        V2 u = pmj02bn_sample(pmj_instance, index);
        float dx = bluenoise(d, px, py);
        float dy = bluenoise(d + 1, px, py);
        u.x = u.x + dx;
        u.y = u.y + dy;
        dim += 2;
        u.x = u.x - std::floor(u.x);
        u.y = u.y - std::floor(u.y);
        return {std::fmin(u.x, ONE_MINUS_EPSILON), std::fmin(u.y, ONE_MINUS_EPSILON)};
    }
    V3 next_3d() {  // Sampler::next_3d default, sampler/mod.rs:24-28
        float u0 = next_1d();
        V2 u12 = next_2d();
        return {u0, u12.x, u12.y};
    }
};

// --------------------------------------------------------------------------------------------
// sampling.rs
// --------------------------------------------------------------------------------------------
inline V2 uniform_sample_disk(V2 u) {  // :5-9
    float r = std::sqrt(u.x);
    float phi = u.y * 2.0f * PI;
    return {r * std::cos(phi), r * std::sin(phi)};
}
inline V3 cos_sample_hemisphere(V2 u) {  // :17-21
    V2 d = uniform_sample_disk(u);
    float z = std::sqrt(std::fmax(1.0f - d.x * d.x - d.y * d.y, 0.0f));
    return {d.x, d.y, z};
}
inline V2 uniform_sample_triangle(V2 u) {  // :32-44
    if (u.x < u.y) {
        float b0 = u.x / 2.0f;
        float b1 = u.y - b0;
        return {b0, b1};
    }
    float b1 = u.y / 2.0f;
    float b0 = u.x - b1;
    return {b0, b1};
}
struct ChoiceU {
    uint32_t i;
    float u;
};
inline ChoiceU uniform_discrete_choice_and_remap(uint32_t n, float u) {  // :54-59
    float fi = std::floor(u * static_cast<float>(n));
    int32_t i = static_cast<int32_t>(fi);
    int32_t hi = static_cast<int32_t>(n) - 1;
    i = i < 0 ? 0 : (i > hi ? hi : i);
    float remapped = u * static_cast<float>(n) - static_cast<float>(static_cast<uint32_t>(i));
    return {static_cast<uint32_t>(i), remapped};
}
inline ChoiceU weighted_discrete_choice2_and_remap(float weight_a, uint32_t a, uint32_t b, float u) {  // :61-70
    bool first = u < weight_a;
    return {first ? a : b, first ? u / weight_a : (u - weight_a) / (1.0f - weight_a)};
}

// util/distribution.rs:34-88
struct AliasTable {
    std::vector<uint32_t> j;
    std::vector<float> t;
    std::vector<float> pdf;
    void build(const std::vector<float> &weights) {
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
    struct Sample {
        uint32_t idx;
        float pdf, u;
    };
    Sample sample_and_remap(float u) const {  // :82-88
        ChoiceU c = uniform_discrete_choice_and_remap(static_cast<uint32_t>(j.size()), u);
        ChoiceU d = weighted_discrete_choice2_and_remap(t[c.i], c.i, j[c.i], c.u);
        return {d.i, pdf[d.i], d.u};
    }
};

// --------------------------------------------------------------------------------------------
// microfacet.rs — TrowbridgeReitzDistribution (sample_visible = true everywhere, principled.rs)
// --------------------------------------------------------------------------------------------
struct TrowbridgeReitz {
    V2 alpha;
    float roughness_;
    static TrowbridgeReitz from_roughness(float rx, float ry) {  // :24-43
        constexpr float MIN_ALPHA = 1e-4f;
        float ax = sqr(rx), ay = sqr(ry);
        TrowbridgeReitz d;
        d.alpha = {std::fmax(ax, MIN_ALPHA), std::fmax(ay, MIN_ALPHA)};
        d.roughness_ = std::sqrt((std::fmax(ax, MIN_ALPHA) + std::fmax(ay, MIN_ALPHA)) * 0.5f);
        return d;
    }
    float d(V3 wh) const {  // :45-57
        float tan2_theta = FrameFn::tan2_theta(wh);
        float cos4_theta = sqr(FrameFn::cos2_theta(wh));
        float ax = alpha.x, ay = alpha.y;
        float e = tan2_theta * (sqr(FrameFn::cos_phi(wh) / ax) + sqr(FrameFn::sin_phi(wh) / ay));
        float inv_d = PI * ax * ay * cos4_theta * sqr(1.0f + e);
        if (!std::isfinite(tan2_theta) || !std::isfinite(inv_d) || inv_d == 0.0f) return 0.0f;
        return 1.0f / inv_d;
    }
    float lambda(V3 w) const {  // :59-65
        float abs_tan_theta = std::fabs(FrameFn::tan_theta(w));
        float alpha2 = FrameFn::cos2_phi(w) * sqr(alpha.x) + FrameFn::sin2_phi(w) * sqr(alpha.y);
        float alpha2_tan2_theta = alpha2 * sqr(abs_tan_theta);
        float l = (-1.0f + std::sqrt(1.0f + alpha2_tan2_theta)) * 0.5f;
        return !std::isfinite(abs_tan_theta) ? 0.0f : l;
    }
    float g1(V3 w) const { return 1.0f / (1.0f + lambda(w)); }                       // :9-11
    float g(V3 wo, V3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }  // :13-15
    V3 sample_wh(V3 w, V2 u) const {                                                 // :118-138
        V3 wh = normalize(v3(alpha.x * w.x, alpha.y * w.y, w.z));
        if (wh.z < 0.0f) wh = -wh;
        V3 t1 = (wh.z < 0.99999f) ? normalize(cross(v3(0, 0, 1), wh)) : v3(1, 0, 0);
        V3 t2 = normalize(cross(wh, t1));
        V2 p = uniform_sample_disk(u);
        float h = std::sqrt(1.0f - sqr(p.x));
        p.y = lerpf(h, p.y, (1.0f + wh.z) * 0.5f);
        float pz = std::sqrt(std::fmax(1.0f - (p.x * p.x + p.y * p.y), 0.0f));
        V3 nh = p.x * t1 + p.y * t2 + pz * wh;
        return normalize(v3(alpha.x * nh.x, alpha.y * nh.y, std::fmax(nh.z, 1e-6f)));
    }
    float pdf(V3 wo, V3 wh) const {  // :196-206
        return d(wh) * g1(wo) * std::fabs(dot(wo, wh)) / FrameFn::abs_cos_theta(wo);
    }
    float roughness() const { return roughness_; }
};

// --------------------------------------------------------------------------------------------
// svm/surface/mod.rs — Fresnel family (:1009-1110) and util::Complex (util/mod.rs:519-604)
// --------------------------------------------------------------------------------------------
inline float fr_dielectric(float cos_theta_i, float eta) {  // :1009-1036
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    eta = cos_theta_i > 0.0f ? eta : 1.0f / eta;
    cos_theta_i = std::fabs(cos_theta_i);
    float sin2_theta_i = 1.0f - sqr(cos_theta_i);
    float sin2_theta_t = sin2_theta_i / sqr(eta);
    if (sin2_theta_t >= 1.0f) return 1.0f;
    float cos_theta_t = std::sqrt(std::fmax(1.0f - sin2_theta_t, 0.0f));
    float r_parl = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
    float r_perp = (cos_theta_i - eta * cos_theta_t) / (cos_theta_i + eta * cos_theta_t);
    float fr = (sqr(r_parl) + sqr(r_perp)) * 0.5f;
    return clampf(fr, 0.0f, 1.0f);
}
struct Cx {
    float re, im;
};
inline Cx cx(float re, float im) { return {re, im}; }
inline Cx operator+(Cx a, Cx b) { return {a.re + b.re, a.im + b.im}; }
inline Cx operator-(Cx a, Cx b) { return {a.re - b.re, a.im - b.im}; }
inline Cx operator*(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Cx operator/(Cx a, Cx b) {
    float scale = 1.0f / (b.re * b.re + b.im * b.im);
    return {(a.re * b.re + a.im * b.im) * scale, (a.im * b.re - a.re * b.im) * scale};
}
inline Cx operator*(Cx a, float s) { return {a.re * s, a.im * s}; }
inline float cx_norm(Cx a) { return a.re * a.re + a.im * a.im; }
inline Cx cx_sqrt(Cx a) {  // util/mod.rs:541-554
    float n = std::sqrt(cx_norm(a));
    float t1 = std::sqrt(0.5f * (n + std::fabs(a.re)));
    float t2 = 0.5f * a.im / t1;
    if (n == 0.0f) return {0.0f, 0.0f};
    if (a.re >= 0.0f) return {t1, t2};
    return {std::fabs(t2), std::copysign(t1, a.im)};
}
inline float fr_complex(float cos_theta_i, Cx eta) {  // :1055-1067
    cos_theta_i = clampf(cos_theta_i, 0.0f, 0.999f);
    float sin2_theta = 1.0f - sqr(cos_theta_i);
    Cx sin2_theta_t = cx(sin2_theta, 0.0f) / (eta * eta);
    Cx cos_theta_t = cx_sqrt(cx(1.0f, 0.0f) - sin2_theta_t);
    Cx r_parl = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
    Cx r_perp = (cx(cos_theta_i, 0.0f) - eta * cos_theta_t) / (cx(cos_theta_i, 0.0f) + eta * cos_theta_t);
    return (cx_norm(r_parl) + cx_norm(r_perp)) * 0.5f;
}
inline void artistic_to_conductor_fresnel(Color color, Color tint, Color &n_out, Color &k_out) {  // :1040-1052
    auto one_ch = [](float c, float g, float &n, float &k) {
        float r = clampf(c, 0.0f, 0.99f);  // Color::clamp(0.99) clamps to [0, 0.99] (color.rs:352-361)
        float r_sqrt = std::sqrt(r);
        float n_min = (1.0f - r) / (1.0f + r);
        float n_max = (1.0f + r_sqrt) / (1.0f - r_sqrt);
        n = lerpf(n_max, n_min, g);
        float k2 = ((n + 1.0f) * (n + 1.0f) * r - (n - 1.0f) * (n - 1.0f)) / (1.0f - r);
        k2 = std::fmax(k2, 0.0f);
        k = std::sqrt(k2);
    };
    one_ch(color.x, tint.x, n_out.x, k_out.x);
    one_ch(color.y, tint.y, n_out.y, k_out.y);
    one_ch(color.z, tint.z, n_out.z, k_out.z);
}
inline float ior_from_f0(float f0) {  // :1090-1094
    float sqrt_f0 = std::sqrt(clampf(f0, 0.0f, 0.99f));
    return (1.0f + sqrt_f0) / (1.0f - sqrt_f0);
}
inline float f0_from_ior(float ior) {  // :1095-1098
    float f0 = (ior - 1.0f) / (ior + 1.0f);
    return sqr(f0);
}
inline float ior_parametrization(float z) { return ior_from_f0(sqr(sqr(z))); }  // :1100-1103

// PreComputedTable::read_3d (:1245-1322): trilinear, x fastest
struct AlbedoTable {
    const float *data;  // [16][16][16]
    static float read_1d(const float *buf, float x, uint32_t offset, uint32_t size) {
        x = clampf(x, 0.0f, 1.0f) * (static_cast<float>(size) - 1.0f);
        uint32_t index = static_cast<uint32_t>(std::floor(x));
        uint32_t nindex = std::min(index + 1, size - 1);
        float t = x - static_cast<float>(index);
        float d0 = buf[offset + index], d1 = buf[offset + nindex];
        return (1.0f - t) * d0 + t * d1;
    }
    static float read_2d(const float *buf, float x, float y, uint32_t offset, uint32_t xs, uint32_t ys) {
        y = clampf(y, 0.0f, 1.0f) * (static_cast<float>(ys) - 1.0f);
        uint32_t index = static_cast<uint32_t>(std::floor(y));
        uint32_t nindex = std::min(index + 1, ys - 1);
        float t = y - static_cast<float>(index);
        float d0 = read_1d(buf, x, offset + xs * index, xs);
        float d1 = read_1d(buf, x, offset + xs * nindex, xs);
        return (1.0f - t) * d0 + t * d1;
    }
    float read_3d(float x, float y, float z) const {
        const uint32_t xs = 16, ys = 16, zs = 16;
        z = clampf(z, 0.0f, 1.0f) * (static_cast<float>(zs) - 1.0f);
        uint32_t index = static_cast<uint32_t>(std::floor(z));
        uint32_t nindex = std::min(index + 1, zs - 1);
        float t = z - static_cast<float>(index);
        float d0 = read_2d(data, x, y, xs * ys * index, xs, ys);
        float d1 = read_2d(data, x, y, xs * ys * nindex, xs, ys);
        return (1.0f - t) * d0 + t * d1;
    }
};
inline float ggx_dielectric_albedo(const AlbedoTable &table, float roughness, float cos_theta_i, float eta) {  // :1145-1154
    float z = std::sqrt(std::fabs((eta - 1.0f) / (eta + 1.0f)));
    cos_theta_i = std::fabs(clampf(cos_theta_i, -0.999f, 0.999f));
    return table.read_3d(roughness, std::fabs(cos_theta_i), z);
}

// --------------------------------------------------------------------------------------------
// Surface closures (svm/surface/mod.rs, diffuse.rs, principled.rs, glass.rs).
// Static composition instead of Rc<dyn Surface>; every method body follows the cited impl.
// `evaluate` returns (f * |cos theta_i|, pdf)   (mod.rs:59).
// --------------------------------------------------------------------------------------------
struct Eval {
    Color f;
    float pdf;
};
struct SampleWi {
    V3 wi;
    bool valid;
};

struct DiffuseBsdf {  // diffuse.rs:13-80
    Color reflectance;
    Eval evaluate(V3 wo, V3 wi) const {
        bool same = FrameFn::same_hemisphere(wo, wi);
        float pdf = same ? FrameFn::abs_cos_theta(wi) * FRAC_1_PI : 0.0f;
        Color c = same ? reflectance * FrameFn::abs_cos_theta(wi) : v3s(0.0f);
        return {c, pdf};
    }
    SampleWi sample_wi(V3 wo, float, V2 u) const {
        V3 wi = cos_sample_hemisphere(u);
        wi = FrameFn::same_hemisphere(wo, wi) ? wi : -wi;
        return {wi, true};
    }
    Color emission(V3) const { return v3s(0.0f); }
    Color albedo(V3) const { return reflectance * PI; }  // diffuse.rs:56-63
    float roughness(V3, float) const { return 1.0f; }    // diffuse.rs:64-72
    V3 ns() const { return v3(0, 0, 1); }
};

struct FresnelDielectric {  // mod.rs:1166-1175
    float eta;
    Color evaluate(float cos_theta_i) const { return v3s(1.0f) * fr_dielectric(cos_theta_i, eta); }
};
struct FresnelComplex {  // mod.rs:1177-1187
    Color n, k;
    Color evaluate(float cos_theta_i) const {
        float c = std::fabs(cos_theta_i);
        return v3(fr_complex(c, cx(n.x, k.x)), fr_complex(c, cx(n.y, k.y)), fr_complex(c, cx(n.z, k.z)));
    }
};

template <class Fresnel> struct MicrofacetReflection {  // mod.rs:820-900
    Color color;
    Fresnel fresnel;
    TrowbridgeReitz dist;
    Eval evaluate(V3 wo, V3 wi) const {
        V3 wh = wo + wi;
        float cos_o = FrameFn::cos_theta(wo), cos_i = FrameFn::cos_theta(wi);
        if ((dot(wh, wo) * dot(wi, wh)) < 0.0f || (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) || cos_i == 0.0f ||
            cos_o == 0.0f || !FrameFn::same_hemisphere(wo, wi))
            return {v3s(0.0f), 0.0f};
        wh = normalize(wh);
        Color f = fresnel.evaluate(dot(wi, face_forward(wh, v3(0, 0, 1))));
        float d = dist.d(wh);
        float g = dist.g(wo, wi);
        Color fc = color * f * std::fabs(0.25f * d * g / (cos_i * cos_o)) * std::fabs(cos_i);
        float pdf = dist.pdf(wo, wh) / (4.0f * std::fabs(dot(wo, wh)));
        return {fc, pdf};
    }
    SampleWi sample_wi(V3 wo, float, V2 u) const {
        V3 wh = dist.sample_wh(wo, u);
        V3 wi = reflect(wo, wh);
        return {wi, FrameFn::same_hemisphere(wo, wi)};
    }
    Color emission(V3) const { return v3s(0.0f); }
    Color albedo(V3) const { return color; }                          // mod.rs:875-882
    float roughness(V3, float) const { return dist.roughness(); }    // mod.rs:883-891
    V3 ns() const { return v3(0, 0, 1); }
};

struct MicrofacetTransmission {  // mod.rs:902-1006
    TrowbridgeReitz dist;
    Color color;
    float eta;
    FresnelDielectric fresnel;
    Eval evaluate(V3 wo, V3 wi) const {
        float cos_o = FrameFn::cos_theta(wo), cos_i = FrameFn::cos_theta(wi);
        float e = cos_o > 0.0f ? eta : 1.0f / eta;
        V3 wh = normalize(wo + wi * e);
        wh = face_forward(wh, v3(0, 0, 1));
        bool backfacing = (dot(wh, wi) * cos_i) < 0.0f || (dot(wh, wo) * cos_o) < 0.0f;
        if ((dot(wh, wo) * dot(wi, wh)) > 0.0f || cos_i == 0.0f || cos_o == 0.0f || backfacing ||
            FrameFn::same_hemisphere(wo, wi))
            return {v3s(0.0f), 0.0f};
        Color f;
        {
            Color fr = fresnel.evaluate(dot(wo, wh));
            float denom = sqr(dot(wi, wh) + dot(wo, wh) / e) * cos_i * cos_o;
            if (denom == 0.0f) f = v3s(0.0f);
            else
                f = (v3s(1.0f) - fr) * color *
                    std::fabs(dist.d(wh) * dist.g(wo, wi) / sqr(e) * std::fabs(dot(wi, wh)) * std::fabs(dot(wo, wh)) / denom) *
                    std::fabs(cos_i);
        }
        float pdf;
        {
            float denom = sqr(dot(wi, wh) + dot(wo, wh) / e);
            float dwh_dwi = std::fabs(dot(wi, wh)) / denom;
            pdf = denom == 0.0f ? 0.0f : dist.pdf(wo, wh) * dwh_dwi;
        }
        return {f, pdf};
    }
    SampleWi sample_wi(V3 wo, float, V2 u) const {
        V3 wh = dist.sample_wh(wo, u);
        Refract r = refract(wo, wh, eta);
        bool valid = r.ok && !FrameFn::same_hemisphere(wo, r.wt);
        return {r.wt, valid};
    }
    Color emission(V3) const { return v3s(0.0f); }
    Color albedo(V3) const { return color; }                          // mod.rs:981-988
    float roughness(V3, float) const { return dist.roughness(); }    // mod.rs:989-997
    V3 ns() const { return v3(0, 0, 1); }
};

enum class Blend { Addictive, Mix };
// BsdfMixture (mod.rs:568-695).  FracFn: float(V3 wo)
template <class A, class B, class FracFn> struct BsdfMixture {
    A a;
    B b;
    FracFn frac;
    Blend mode;
    static constexpr float EPS = 1e-4f;
    Eval evaluate(V3 wo, V3 wi) const {
        float fr = frac(wo);
        if (mode == Blend::Addictive) {
            Eval ea = a.evaluate(wo, wi);
            Eval eb = b.evaluate(wo, wi);
            return {ea.f + eb.f, lerpf(ea.pdf, eb.pdf, fr)};
        }
        Eval ea = fr < 1.0f - EPS ? a.evaluate(wo, wi) : Eval{v3s(0.0f), 0.0f};
        Eval eb = fr > EPS ? b.evaluate(wo, wi) : Eval{v3s(0.0f), 0.0f};
        return {lerp3(ea.f, eb.f, v3s(fr)), lerpf(ea.pdf, eb.pdf, fr)};
    }
    SampleWi sample_wi(V3 wo, float u_select, V2 u) const {
        float fr = frac(wo);
        ChoiceU c = weighted_discrete_choice2_and_remap(fr, 1u, 0u, u_select);
        if (c.i == 0) return a.sample_wi(wo, c.u, u);
        return b.sample_wi(wo, c.u, u);
    }
    Color emission(V3 wo) const {
        float fr = frac(wo);
        if (mode == Blend::Addictive) return a.emission(wo) + b.emission(wo);
        return a.emission(wo) * (1.0f - fr) + b.emission(wo) * fr;
    }
    Color albedo(V3 wo) const {  // mod.rs:659-675
        float fr = frac(wo);
        if (mode == Blend::Addictive) return a.albedo(wo) + b.albedo(wo);
        return a.albedo(wo) * (1.0f - fr) + b.albedo(wo) * fr;
    }
    float roughness(V3 wo, float u_select) const {  // mod.rs:641-657
        float fr = frac(wo);
        ChoiceU c = weighted_discrete_choice2_and_remap(fr, 1u, 0u, u_select);
        if (c.i == 0) return a.roughness(wo, c.u);
        return b.roughness(wo, c.u);
    }
    V3 ns() const { return a.ns(); }  // mod.rs:575-577
};
// CoatedBsdf (mod.rs:476-567).  EFn: Color(V3 w)
template <class Top, class Bottom, class EFn> struct CoatedBsdf {
    Top top;
    Bottom bottom;
    EFn e_top;
    Eval evaluate(V3 wo, V3 wi) const {
        Eval et = top.evaluate(wo, wi);
        Eval eb = bottom.evaluate(wo, wi);
        Color eo = e_top(wo);
        Color ei = e_top(wi);
        float pdf_select_top = avg(eo);
        Color one = v3s(1.0f);
        float pdf_select_bottom = 1.0f - pdf_select_top;
        float pdf = et.pdf * pdf_select_top + eb.pdf * pdf_select_bottom;
        Color f = et.f + eb.f * min3(one - eo, one - ei);
        return {f, pdf};
    }
    SampleWi sample_wi(V3 wo, float u_select, V2 u) const {
        Color eo = e_top(wo);
        float pdf_select_top = avg(eo);
        ChoiceU c = weighted_discrete_choice2_and_remap(pdf_select_top, 0u, 1u, u_select);
        if (c.i == 0) return top.sample_wi(wo, c.u, u);
        return bottom.sample_wi(wo, c.u, u);
    }
    Color emission(V3 wo) const {
        Color eo = e_top(wo);
        return top.emission(wo) * eo + bottom.emission(wo) * (v3s(1.0f) - eo);
    }
    Color albedo(V3 wo) const {  // mod.rs:523-535
        Color eo = e_top(wo);
        return top.albedo(wo) * eo + bottom.albedo(wo) * (v3s(1.0f) - eo);
    }
    float roughness(V3 wo, float u_select) const {  // mod.rs:536-553
        Color eo = e_top(wo);
        ChoiceU c = weighted_discrete_choice2_and_remap(avg(eo), 0u, 1u, u_select);
        if (c.i == 0) return top.roughness(wo, c.u);
        return bottom.roughness(wo, c.u);
    }
    V3 ns() const { return bottom.ns(); }  // mod.rs:482-484
};
template <class Inner> struct ScaledBsdf {  // mod.rs:412-475 (weight is wo-independent in principled.rs:190-193)
    Inner inner;
    Color weight;
    Eval evaluate(V3 wo, V3 wi) const {
        Eval e = inner.evaluate(wo, wi);
        return {e.f * weight, e.pdf};
    }
    SampleWi sample_wi(V3 wo, float us, V2 u) const { return inner.sample_wi(wo, us, u); }
    Color emission(V3 wo) const { return inner.emission(wo) * weight; }
    Color albedo(V3 wo) const { return inner.albedo(wo) * weight; }           // mod.rs:446-454
    float roughness(V3 wo, float us) const { return inner.roughness(wo, us); }  // mod.rs:456-464
    V3 ns() const { return inner.ns(); }
};
template <class Inner> struct EmissiveSurface {  // mod.rs:330-411, inner = Some
    Inner inner;
    Color emission_;
    Eval evaluate(V3 wo, V3 wi) const { return inner.evaluate(wo, wi); }
    SampleWi sample_wi(V3 wo, float us, V2 u) const { return inner.sample_wi(wo, us, u); }
    Color emission(V3 wo) const { return emission_ + inner.emission(wo); }
    Color albedo(V3 wo) const { return inner.albedo(wo); }                       // mod.rs:372-383
    float roughness(V3 wo, float us) const { return inner.roughness(wo, us); }  // mod.rs:385-397
    V3 ns() const { return inner.ns(); }                                         // mod.rs:338-342
};
struct EmissionOnly {  // EmissiveSurface { inner: None } (svm/mod.rs:124-133)
    Color emission_;
    Eval evaluate(V3, V3) const { return {v3s(0.0f), 0.0f}; }
    SampleWi sample_wi(V3, float, V2) const { return {v3s(0.0f), false}; }
    Color emission(V3) const { return emission_; }
    Color albedo(V3) const { return v3s(0.0f); }          // mod.rs:372-383 (inner = None)
    float roughness(V3, float) const { return 1.0f; }     // mod.rs:385-397
    V3 ns() const { return v3(0, 0, 1); }
};
template <class Inner> struct PrincipledBsdfWrapper {  // principled.rs:218-275
    Inner inner;
    Color albedo_c, emission_;
    Eval evaluate(V3 wo, V3 wi) const { return inner.evaluate(wo, wi); }
    SampleWi sample_wi(V3 wo, float us, V2 u) const { return inner.sample_wi(wo, us, u); }
    Color emission(V3) const { return emission_; }
    Color albedo(V3) const { return albedo_c; }                                 // principled.rs:227-234
    float roughness(V3 wo, float us) const { return inner.roughness(wo, us); }  // principled.rs:257-265
    V3 ns() const { return v3(0, 0, 1); }                                        // principled.rs:224-226
};
// SurfaceClosure (mod.rs:697-816)
template <class Inner> struct SurfaceClosure {
    Inner inner;
    Frame frame;
    V3 ng;
    bool check_wo_wi_valid(V3 wo, V3 wi) const {  // :706-718
        auto sign = [](float x) { return x > 0.0f ? 1.0f : -1.0f; };
        V3 ns = frame.n;
        float flipped = sign(dot(ng, ns));
        return (sign(flipped * dot(wo, ns)) * sign(dot(wo, ng)) > 0.0f) &&
               (sign(flipped * dot(wi, ns)) * sign(dot(wi, ng)) > 0.0f);
    }
    Eval evaluate(V3 wo, V3 wi) const {  // :729-748
        if (!check_wo_wi_valid(wo, wi)) return {v3s(0.0f), 0.0f};
        return inner.evaluate(to_local(frame, wo), to_local(frame, wi));
    }
    SampleWi sample_wi(V3 wo, float us, V2 u) const {  // :750-764
        SampleWi s = inner.sample_wi(to_local(frame, wo), us, u);
        V3 wi = to_world(frame, s.wi);
        bool valid = s.valid && check_wo_wi_valid(wo, wi);
        return {wi, valid};
    }
    Color emission(V3 wo) const { return inner.emission(to_local(frame, wo)); }
    Color albedo(V3 wo) const { return inner.albedo(to_local(frame, wo)); }                           // mod.rs:766-773
    float roughness(V3 wo, float us) const { return inner.roughness(to_local(frame, wo), us); }      // mod.rs:774-783
    V3 ns() const { return to_world(frame, inner.ns()); }                                             // mod.rs:724-727
};
struct BsdfSample {  // mod.rs:34-51
    V3 wi;
    float pdf;
    Color color;
    bool valid;
};
template <class Closure> BsdfSample closure_sample(const Closure &c, V3 wo, float u_select, V2 u_sample) {  // mod.rs:795-815
    SampleWi s = c.sample_wi(wo, u_select, u_sample);
    if (!s.valid) return {v3s(0.0f), 0.0f, v3s(0.0f), false};
    Eval e = c.evaluate(wo, s.wi);
    return {s.wi, e.pdf, e.f, s.valid && e.pdf > 0.0f};
}

// --------------------------------------------------------------------------------------------
// Scene data prepared once (the oracle's stand-in for load.rs:238-456 + mesh.rs:258-348)
// --------------------------------------------------------------------------------------------
struct SurfaceInteraction {  // interaction.rs:15-26
    Frame frame;
    V3 p, ng;
    V2 bary, uv;
    uint32_t inst_id, prim_id;
    AkrShaderRef surface;
    uint32_t mat_index;  // index into the instance's material list (oracle-side cache key for `surface`)
    float prim_area;
    bool valid;
};

struct PrincipledInputs {  // evaluated inputs of SvmPrincipledBsdf (principled.rs:23-49,200-201)
    Color color;
    float alpha;
    Color emission;
    float metallic, roughness, eta, transmission, specular_ior_level;
    Color specular_tint;
    float clearcoat_weight, clearcoat_roughness, clearcoat_ior;
    Color clearcoat_tint;
    V3 normal;
    float roughness_raw;  // svm_eval.eval_float(self.roughness) at principled.rs:103
};
struct ShaderEval {
    uint32_t out_op = UINT32_MAX;  // op of the closure feeding MaterialOutput
    PrincipledInputs principled{};
    Color diffuse_reflectance{};
    float diffuse_alpha = 1.0f;
    Color emission_only{};
    Color glass_kr{}, glass_kt{};
    float glass_eta = 1.0f, glass_roughness = 0.0f;
    bool dynamic = false;  // some input depends on the hit (uv / image texture): re-evaluated per dispatch
};
struct OInstance {
    float m[16];  // column-major
    float transform_det;
    uint32_t geom_id, flags;
    const AkrShaderRef *materials;
    uint32_t n_materials;
    int32_t light_id;  // -1 = not a light (TagIndex::INVALID)
    AliasTable area_sampler;
    uint32_t tri_offset;
    std::vector<ShaderEval> evals;  // evaluated constants of materials[k] (pure function of the blob)
    std::vector<float> alphas;      // Surface::alpha() of materials[k] in SvmEvalMode::Alpha (constant shaders)
    bool any_alpha = false;         // some material has alpha < 1, or an alpha that depends on the hit
};
struct WorldTri {
    V3 v0, v1, v2;
    uint32_t inst, prim;
};
struct OLight {
    uint32_t instance_id, geom_id;
};

// glam Mat4::determinant (scalar form) — mesh.rs:311-312
float mat4_det(const float *m) {
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

struct M3 {  // column-major 3x3
    V3 c0, c1, c2;
};
inline V3 mul(const M3 &m, V3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
inline M3 transpose(const M3 &m) { return {v3(m.c0.x, m.c1.x, m.c2.x), v3(m.c0.y, m.c1.y, m.c2.y), v3(m.c0.z, m.c1.z, m.c2.z)}; }
// ASSUMED (luisa Mat3::inverse): adjugate / determinant
inline M3 inverse(const M3 &m) {
    V3 a = m.c0, b = m.c1, c = m.c2;
    V3 r0 = cross(b, c), r1 = cross(c, a), r2 = cross(a, b);
    float inv_det = 1.0f / dot(r2, c);
    // inverse rows are r0,r1,r2 scaled; store column-major
    return {v3(r0.x, r1.x, r2.x) * inv_det, v3(r0.y, r1.y, r2.y) * inv_det, v3(r0.z, r1.z, r2.z) * inv_det};
}

struct Camera {  // camera/mod.rs:108-153
    float c2w[16];
    float r2c[16];
    uint32_t width, height;
};

struct Scene {
    const AkrSceneDesc *desc;
    std::vector<OInstance> instances;
    std::vector<WorldTri> tris;
    std::vector<OLight> lights;
    AliasTable light_distribution;
    AlbedoTable albedo;
    Camera camera;
    std::string error;
};

// ---- SVM evaluation: literal per-dispatch interpretation (svm/eval.rs:97-269,364-380) ----------
struct SvmValue {
    enum Kind { None, Float, Float2, Float3, Float4, ColorAlpha, Closure, TexCoords, SeparateColor } kind = None;
    float f = 0.0f;
    float v[4] = {0, 0, 0, 0};  // Float3/Float4 or ColorAlpha (rgb, alpha)
};

float val_float_auto(const SvmValue &v) {  // eval_float_auto_convert (eval.rs:327-343)
    if (v.kind == SvmValue::Float) return v.f;
    return v.v[0];
}
V3 val_float3_auto(const SvmValue &v) {  // eval_float3_auto_convert (eval.rs:311-326)
    if (v.kind == SvmValue::Float3 || v.kind == SvmValue::Float4) return v3(v.v[0], v.v[1], v.v[2]);
    if (v.kind == SvmValue::Float2 || v.kind == SvmValue::TexCoords) return v3(v.v[0], v.v[1], 0.0f);
    return v3(v.f, 0.0f, 0.0f);
}
V2 val_float2_auto(const SvmValue &v) {  // eval_float2_auto_convert (eval.rs:296-310)
    if (v.kind == SvmValue::Float) return V2{v.f, 0.0f};
    return V2{v.v[0], v.v[1]};
}

// ---- image textures: ASSUMED (luisa `Tex2d::sample` behind the bindless heap, eval.rs:139-147).  Normalised
// coordinates, texel centres at (i + 0.5) / size, bilinear weights in f32, the sampler's address mode per texel. ----
inline int wrap_texel(int i, int n, uint32_t address, bool &zero) {
    zero = false;
    if (i >= 0 && i < n) return i;
    switch (address) {
    case AKR_ADDRESS_REPEAT: {
        int m = i % n;
        return m < 0 ? m + n : m;
    }
    case AKR_ADDRESS_MIRROR: {
        int period = 2 * n;
        int m = i % period;
        if (m < 0) m += period;
        return m < n ? m : period - 1 - m;
    }
    case AKR_ADDRESS_EDGE: return i < 0 ? 0 : n - 1;
    default: zero = true; return 0;
    }
}
inline void fetch_texel(const AkrImage &img, int x, int y, float out[4]) {
    bool zx, zy;
    int ix = wrap_texel(x, static_cast<int>(img.width), img.address, zx);
    int iy = wrap_texel(y, static_cast<int>(img.height), img.address, zy);
    if (zx || zy) {
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        return;
    }
    size_t i = (static_cast<size_t>(iy) * img.width + static_cast<size_t>(ix)) * 4;
    if (img.texel_format == AKR_TEXEL_RGBA8) {
        const uint8_t *t = static_cast<const uint8_t *>(img.texels) + i;
        for (int c = 0; c < 4; ++c) out[c] = static_cast<float>(t[c]) / 255.0f;
    } else {
        const float *t = static_cast<const float *>(img.texels) + i;
        for (int c = 0; c < 4; ++c) out[c] = t[c];
    }
}
inline void sample_texture(const AkrImage &img, V2 uv, float out[4]) {
    float fx = uv.x * static_cast<float>(img.width), fy = uv.y * static_cast<float>(img.height);
    if (img.filter == AKR_FILTER_POINT) {
        fetch_texel(img, static_cast<int>(std::floor(fx)), static_cast<int>(std::floor(fy)), out);
        return;
    }
    float x = fx - 0.5f, y = fy - 0.5f;
    float x0 = std::floor(x), y0 = std::floor(y);
    float tx = x - x0, ty = y - y0;
    int ix = static_cast<int>(x0), iy = static_cast<int>(y0);
    float c00[4], c10[4], c01[4], c11[4];
    fetch_texel(img, ix, iy, c00);
    fetch_texel(img, ix + 1, iy, c10);
    fetch_texel(img, ix, iy + 1, c01);
    fetch_texel(img, ix + 1, iy + 1, c11);
    for (int c = 0; c < 4; ++c) {
        float a = c00[c] * (1.0f - tx) + c10[c] * tx;
        float b = c01[c] * (1.0f - tx) + c11[c] * tx;
        out[c] = a * (1.0f - ty) + b * ty;
    }
}
inline float srgb_to_linear1(float s) {  // color.rs:555-558
    return s <= 0.04045f ? s / 12.92f : std::pow((s + 0.055f) / 1.055f, 2.4f);
}

bool eval_shader(const Scene &sc, AkrShaderRef ref, V2 si_uv, ShaderEval &out, std::string *err) {
    const AkrSceneDesc &d = *sc.desc;
    if (ref.shader_kind >= d.n_shader_kinds) {
        if (err) *err = "shader kind out of range";
        return false;
    }
    const AkrShaderKind &kind = d.shader_kinds[ref.shader_kind];
    SvmValue vals[64];
    if (kind.n_nodes > 64) {
        if (err) *err = "shader has more than 64 nodes";
        return false;
    }
    auto read_f32 = [&](uint32_t off) {
        float f;
        std::memcpy(&f, d.shader_data + ref.data_offset + off, 4);
        return f;
    };
    for (uint32_t i = 0; i < kind.n_nodes; ++i) {
        const AkrSvmNode &n = kind.nodes[i];
        SvmValue &r = vals[i];
        switch (n.op) {
        case AKR_SVM_FLOAT:
            r.kind = SvmValue::Float;
            r.f = read_f32(n.a[0]);
            break;
        case AKR_SVM_FLOAT3:
            r.kind = SvmValue::Float3;
            r.v[0] = read_f32(n.a[0]);
            r.v[1] = read_f32(n.a[0] + 4);
            r.v[2] = read_f32(n.a[0] + 8);
            break;
        case AKR_SVM_RGB_TEX: {  // eval.rs:127-136: rgb_to_target_colorspace(sRGB -> sRGB) = identity, .extend(1.0)
            const SvmValue &rgb = vals[n.a[0]];
            if (n.a[1] != 1) {
                if (err) *err = "only sRGB rgb nodes are supported (ACEScg needs the CAT matrices)";
                return false;
            }
            r.kind = SvmValue::Float4;
            r.v[0] = rgb.v[0];
            r.v[1] = rgb.v[1];
            r.v[2] = rgb.v[2];
            r.v[3] = 1.0f;
            break;
        }
        case AKR_SVM_SPECTRAL_UPLIFT: {  // eval.rs:160-180: RGB passthrough + alpha
            const SvmValue &rgba = vals[n.a[0]];
            r.kind = SvmValue::ColorAlpha;
            r.v[0] = rgba.v[0];
            r.v[1] = rgba.v[1];
            r.v[2] = rgba.v[2];
            r.v[3] = rgba.v[3];
            break;
        }
        case AKR_SVM_DIFFUSE_BSDF: {  // diffuse.rs:82-104
            const SvmValue &c = vals[n.a[0]];
            r.kind = SvmValue::Closure;
            out.diffuse_reflectance = v3(c.v[0], c.v[1], c.v[2]) * FRAC_1_PI;
            out.diffuse_alpha = c.v[3];
            out.out_op = n.op;
            break;
        }
        case AKR_SVM_EMISSION: {  // svm/mod.rs:124-133
            const SvmValue &c = vals[n.a[0]];
            float strength = vals[n.a[1]].f;
            r.kind = SvmValue::Closure;
            out.emission_only = v3(c.v[0], c.v[1], c.v[2]) * strength;
            out.out_op = n.op;
            break;
        }
        case AKR_SVM_GLASS_BSDF: {  // glass.rs:13-45
            const SvmValue &kr = vals[n.a[0]];
            const SvmValue &kt = vals[n.a[1]];
            r.kind = SvmValue::Closure;
            out.glass_kr = v3(kr.v[0], kr.v[1], kr.v[2]);
            out.glass_kt = v3(kt.v[0], kt.v[1], kt.v[2]);
            out.glass_roughness = vals[n.a[2]].f;
            out.glass_eta = vals[n.a[3]].f;
            out.out_op = n.op;
            break;
        }
        case AKR_SVM_PRINCIPLED_BSDF: {  // principled.rs:23-49,200-201
            PrincipledInputs &p = out.principled;
            auto col = [&](uint32_t k) {
                const SvmValue &c = vals[n.a[k]];
                return v3(c.v[0], c.v[1], c.v[2]);
            };
            auto flt = [&](uint32_t k) { return val_float_auto(vals[n.a[k]]); };
            p.color = col(AKR_P_BASE_COLOR);
            p.alpha = vals[n.a[AKR_P_BASE_COLOR]].v[3];
            p.emission = col(AKR_P_EMISSION_COLOR) * flt(AKR_P_EMISSION_STRENGTH);
            p.metallic = flt(AKR_P_METALLIC);
            p.roughness = flt(AKR_P_ROUGHNESS);
            p.eta = flt(AKR_P_IOR);
            p.transmission = flt(AKR_P_TRANSMISSION_WEIGHT);
            p.specular_ior_level = flt(AKR_P_SPECULAR_IOR_LEVEL);
            p.specular_tint = col(AKR_P_SPECULAR_TINT);
            p.clearcoat_weight = flt(AKR_P_COAT_WEIGHT);
            p.clearcoat_roughness = flt(AKR_P_COAT_ROUGHNESS);
            p.clearcoat_ior = flt(AKR_P_COAT_IOR);
            p.clearcoat_tint = col(AKR_P_COAT_TINT);
            p.normal = val_float3_auto(vals[n.a[AKR_P_NORMAL]]);
            p.roughness_raw = vals[n.a[AKR_P_ROUGHNESS]].f;
            r.kind = SvmValue::Closure;
            out.out_op = n.op;
            break;
        }
        case AKR_SVM_MATERIAL_OUTPUT:
            r.kind = SvmValue::Closure;
            break;
        case AKR_SVM_RGB_IMAGE_TEX: {  // eval.rs:137-157
            uint32_t tex_idx;
            std::memcpy(&tex_idx, d.shader_data + ref.data_offset + n.a[0], 4);
            if (tex_idx >= d.n_images) {
                if (err) *err = "texture index out of range";
                return false;
            }
            V2 uv = n.a[2] != AKR_SVM_NONE ? val_float2_auto(vals[n.a[2]]) : si_uv;
            float rgba[4];
            sample_texture(d.images[tex_idx], uv, rgba);
            if (n.a[1] != 0)  // rgb_gamma_correction(rgb, sRGB) = srgb_to_linear (texture/mod.rs:52-58)
                for (int c = 0; c < 3; ++c) rgba[c] = srgb_to_linear1(rgba[c]);
            r.kind = SvmValue::Float4;
            for (int c = 0; c < 4; ++c) r.v[c] = rgba[c];
            out.dynamic = true;
            break;
        }
        case AKR_SVM_NORMAL_MAP: {  // eval.rs:182-196
            V3 nv = val_float3_auto(vals[n.a[0]]);
            V3 normal = 2.0f * nv - v3s(1.0f);
            float strength = val_float_auto(vals[n.a[1]]);
            if (strength != 1.0f) normal = normal * v3(strength, strength, 1.0f);
            r.kind = SvmValue::Float3;
            r.v[0] = normal.x;
            r.v[1] = normal.y;
            r.v[2] = normal.z;
            break;
        }
        case AKR_SVM_MAPPING: {  // eval.rs:197-213 (rotation is a todo in the reference)
            V3 v = val_float3_auto(vals[n.a[0]]);
            V3 location = val_float3_auto(vals[n.a[2]]);
            V3 scale = val_float3_auto(vals[n.a[4]]);
            V3 o = n.a[1] == 0 ? v * scale + location : (v - location) / scale;
            r.kind = SvmValue::Float3;
            r.v[0] = o.x;
            r.v[1] = o.y;
            r.v[2] = o.z;
            break;
        }
        case AKR_SVM_TEX_COORDS:  // eval.rs:225-232
            r.kind = SvmValue::TexCoords;
            r.v[0] = si_uv.x;
            r.v[1] = si_uv.y;
            out.dynamic = true;
            break;
        case AKR_SVM_SEPARATE_COLOR: {  // eval.rs:249-264
            V3 c = val_float3_auto(vals[n.a[0]]);
            r.kind = SvmValue::SeparateColor;
            r.v[0] = c.x;
            r.v[1] = c.y;
            r.v[2] = c.z;
            break;
        }
        case AKR_SVM_EXTRACT_FIELD: {  // eval.rs:214-224
            const SvmValue &src = vals[n.a[0]];
            if (src.kind == SvmValue::TexCoords && n.a[1] == AKR_SVM_FIELD_UV) {
                r.kind = SvmValue::Float2;
                r.v[0] = src.v[0];
                r.v[1] = src.v[1];
            } else if (src.kind == SvmValue::SeparateColor && n.a[1] >= AKR_SVM_FIELD_RED && n.a[1] <= AKR_SVM_FIELD_BLUE) {
                r.kind = SvmValue::Float;
                r.f = src.v[n.a[1] - AKR_SVM_FIELD_RED];
            } else {
                if (err) *err = "extract: field not found";
                return false;
            }
            break;
        }
        case AKR_SVM_CHECKERBOARD: {  // eval.rs:233-248
            V2 uv = n.a[0] != AKR_SVM_NONE ? val_float2_auto(vals[n.a[0]]) : si_uv;
            if (n.a[0] == AKR_SVM_NONE) out.dynamic = true;
            const SvmValue &c1 = vals[n.a[2]];
            const SvmValue &c2 = vals[n.a[3]];
            float scale = vals[n.a[1]].f;
            int px = static_cast<int>(std::floor(uv.x * scale * 2.0f)), py = static_cast<int>(std::floor(uv.y * scale * 2.0f));
            const SvmValue &pick = ((px + py) % 2 == 0) ? c1 : c2;
            r.kind = SvmValue::ColorAlpha;
            for (int c = 0; c < 4; ++c) r.v[c] = pick.v[c];
            break;
        }
        default:
            if (err) *err = "unsupported SVM op " + std::to_string(n.op);
            return false;
        }
    }
    return true;
}

// normal_map (mod.rs:1380-1417), TangentSpace
inline Frame normal_map_frame(V3 normal, const Frame &frame) {
    if (normal.x == 0.0f && normal.y == 0.0f && normal.z == 0.0f) return frame_identity();
    V3 tt = frame.t;
    normal = normalize(normal);
    V3 n_world = to_world(frame, normal);
    Frame nf = frame_from_n_t(n_world, tt);
    Frame r;
    r.t = to_local(frame, nf.t);
    r.s = to_local(frame, nf.s);
    r.n = to_local(frame, nf.n);
    return r;
}

// Build the closure for `si` exactly as Svm::dispatch_surface does (eval.rs:468-495) and hand it to `f`.
template <class F> auto with_surface_closure(const Scene &sc, const SurfaceInteraction &si, bool force_diffuse, F &&f) {
    if (force_diffuse) {  // pt.rs:268-279
        SurfaceClosure<DiffuseBsdf> c{DiffuseBsdf{v3s(1.0f) * FRAC_1_PI * 0.8f}, si.frame, si.ng};
        return f(c);
    }
    // SvmEvaluator::eval_shader re-reads the constant blob on every dispatch (eval.rs:364-380); the result
    // is a pure function of (kind, data_offset), evaluated once in prepare_scene with the same code.
    const ShaderEval &cached = sc.instances[si.inst_id].evals[si.mat_index];
    ShaderEval per_hit;
    if (cached.dynamic) eval_shader(sc, si.surface, si.uv, per_hit, nullptr);  // texture-driven inputs: evaluated at this hit's uv
    const ShaderEval &ev = cached.dynamic ? per_hit : cached;
    if (ev.out_op == AKR_SVM_DIFFUSE_BSDF) {
        SurfaceClosure<DiffuseBsdf> c{DiffuseBsdf{ev.diffuse_reflectance}, si.frame, si.ng};
        return f(c);
    }
    if (ev.out_op == AKR_SVM_EMISSION) {
        SurfaceClosure<EmissionOnly> c{EmissionOnly{ev.emission_only}, si.frame, si.ng};
        return f(c);
    }
    if (ev.out_op == AKR_SVM_GLASS_BSDF) {
        float eta = ev.glass_eta;
        FresnelDielectric fresnel{eta};
        TrowbridgeReitz dist = TrowbridgeReitz::from_roughness(ev.glass_roughness, ev.glass_roughness);
        MicrofacetReflection<FresnelDielectric> reflection{ev.glass_kr, fresnel, dist};
        MicrofacetTransmission transmission{dist, ev.glass_kt, eta, fresnel};
        auto frac = [eta](V3 wo) { return fr_dielectric(FrameFn::cos_theta(wo), eta); };
        BsdfMixture<MicrofacetTransmission, MicrofacetReflection<FresnelDielectric>, decltype(frac)> blend{
            transmission, reflection, frac, Blend::Addictive};
        SurfaceClosure<decltype(blend)> c{blend, si.frame, si.ng};
        return f(c);
    }
    // ---- Principled (principled.rs:13-216) ----
    const PrincipledInputs &p = ev.principled;
    const AlbedoTable &table = sc.albedo;
    Color color = p.color;
    Color transmission_color = v3(std::sqrt(color.x), std::sqrt(color.y), std::sqrt(color.z));
    DiffuseBsdf diffuse{color * FRAC_1_PI};
    float roughness = p.roughness;
    // specular layer (:55-80)
    float eta_s = p.eta;
    float f0 = f0_from_ior(eta_s);
    if (p.specular_ior_level != 0.5f) {
        f0 *= 2.0f * p.specular_ior_level;
        eta_s = ior_from_f0(f0);
    }
    float specular_weight = f0;
    Color specular_tint = p.specular_tint;
    MicrofacetReflection<FresnelDielectric> specular_brdf{specular_tint * f0, FresnelDielectric{eta_s},
                                                          TrowbridgeReitz::from_roughness(roughness, roughness)};
    // clearcoat (:81-98)
    float cc_w = p.clearcoat_weight, cc_r = p.clearcoat_roughness, cc_ior = p.clearcoat_ior;
    MicrofacetReflection<FresnelDielectric> clearcoat_brdf{v3s(1.0f) * cc_w, FresnelDielectric{cc_ior},
                                                           TrowbridgeReitz::from_roughness(cc_r, cc_r)};
    // dielectric (:99-130)
    float eta = p.eta;
    float rough_raw = p.roughness_raw;
    FresnelDielectric fresnel{eta};
    MicrofacetReflection<FresnelDielectric> d_reflection{color, fresnel, TrowbridgeReitz::from_roughness(rough_raw, rough_raw)};
    MicrofacetTransmission d_transmission{TrowbridgeReitz::from_roughness(rough_raw, rough_raw), transmission_color, eta, fresnel};
    auto d_frac = [eta](V3 wo) { return fr_dielectric(FrameFn::cos_theta(wo), eta); };
    BsdfMixture<MicrofacetTransmission, MicrofacetReflection<FresnelDielectric>, decltype(d_frac)> dielectric{
        d_transmission, d_reflection, d_frac, Blend::Addictive};
    // metal (:131-142)
    Color mn, mk;
    artistic_to_conductor_fresnel(color, specular_tint, mn, mk);
    MicrofacetReflection<FresnelComplex> metal{v3s(1.0f), FresnelComplex{mn, mk}, TrowbridgeReitz::from_roughness(roughness, roughness)};
    // diffuse/transmission mix (:143-148)
    float transmission = p.transmission;
    auto t_frac = [transmission](V3) { return transmission; };
    BsdfMixture<DiffuseBsdf, decltype(dielectric), decltype(t_frac)> bsdf0{diffuse, dielectric, t_frac, Blend::Mix};
    // specular coat (:151-168)
    auto e_spec = [&table, roughness, eta_s, specular_tint, specular_weight](V3 w) {
        float cos_theta = FrameFn::abs_cos_theta(w);
        float albedo = ggx_dielectric_albedo(table, roughness, cos_theta, eta_s);
        return specular_tint * albedo * specular_weight;
    };
    CoatedBsdf<decltype(specular_brdf), decltype(bsdf0), decltype(e_spec)> bsdf1{specular_brdf, bsdf0, e_spec};
    // metallic mix (:170-175)
    float metallic = p.metallic;
    auto m_frac = [metallic](V3) { return metallic; };
    BsdfMixture<decltype(bsdf1), decltype(metal), decltype(m_frac)> bsdf2{bsdf1, metal, m_frac, Blend::Mix};
    // emission (:178-181)
    EmissiveSurface<decltype(bsdf2)> bsdf3{bsdf2, p.emission};
    // clearcoat (:183-199)
    auto e_coat = [&table, cc_w, cc_r, cc_ior](V3 w) {
        float a = ggx_dielectric_albedo(table, cc_r, FrameFn::abs_cos_theta(w), cc_ior);
        return v3s(1.0f) * cc_w * a;
    };
    ScaledBsdf<decltype(bsdf3)> scaled{bsdf3, lerp3(v3s(1.0f), p.clearcoat_tint, v3s(cc_w))};
    CoatedBsdf<decltype(clearcoat_brdf), decltype(scaled), decltype(e_coat)> bsdf4{clearcoat_brdf, scaled, e_coat};
    // wrapper + normal map (:200-214)
    V3 normal = p.normal;
    normal.x = -normal.x;
    normal.y = -normal.y;
    PrincipledBsdfWrapper<decltype(bsdf4)> wrapper{bsdf4, color, p.emission};
    SurfaceClosure<decltype(wrapper)> inner{wrapper, normal_map_frame(normal, si.frame), to_local(si.frame, si.ng)};
    SurfaceClosure<decltype(inner)> outer{inner, si.frame, si.ng};
    return f(outer);
}

// ---- MeshAggregate::surface_interaction (mesh.rs:487-654) ---------------------------------------
SurfaceInteraction surface_interaction(const Scene &sc, uint32_t inst_id, uint32_t prim_id, V2 bary) {
    const OInstance &inst = sc.instances[inst_id];
    const AkrMesh &g = sc.desc->meshes[inst.geom_id];
    uint32_t mat_index = (inst.flags & AKR_MESH_HAS_MULTI_MATERIALS) ? g.material_slots[prim_id] : 0u;  // mesh.rs:509-519
    AkrShaderRef material = inst.materials[mat_index];
    const uint32_t *idx = g.indices + 3 * prim_id;
    auto vert = [&](uint32_t i) { return v3(g.vertices[3 * i], g.vertices[3 * i + 1], g.vertices[3 * i + 2]); };
    V3 v0 = vert(idx[0]), v1 = vert(idx[1]), v2 = vert(idx[2]);
    // ASSUMED (luisa TriangleInterpolate): (1 - u - v) * a + u * b + v * c
    auto interp3 = [&](V3 a, V3 b, V3 c) { return (1.0f - bary.x - bary.y) * a + bary.x * b + bary.y * c; };
    auto interp2 = [&](V2 a, V2 b, V2 c) {
        float w = 1.0f - bary.x - bary.y;
        return V2{w * a.x + bary.x * b.x + bary.y * c.x, w * a.y + bary.x * b.y + bary.y * c.y};
    };
    V3 p_local = interp3(v0, v1, v2);
    V3 ngu = cross(v1 - v0, v2 - v0);
    float len = length(ngu);
    float area_local = len * 0.5f;
    V3 ng_local = ngu / len;
    uint32_t prim_id3 = prim_id * 3;
    V2 uv0, uv1, uv2;
    if (g.uvs) {
        uv0 = {g.uvs[2 * (prim_id3 + 0)], g.uvs[2 * (prim_id3 + 0) + 1]};
        uv1 = {g.uvs[2 * (prim_id3 + 1)], g.uvs[2 * (prim_id3 + 1) + 1]};
        uv2 = {g.uvs[2 * (prim_id3 + 2)], g.uvs[2 * (prim_id3 + 2) + 1]};
    } else {
        uv0 = {0.0f, 0.0f};
        uv1 = {1.0f, 0.0f};
        uv2 = {1.0f, 0.1f};
    }
    V2 uv = interp2(uv0, uv1, uv2);
    // tangent (:553-591)
    V3 tt_local = v3s(0.0f);
    {
        bool use_default = false;
        V3 t = v3s(0.0f);
        if (g.tangents) {
            auto tan = [&](uint32_t i) { return v3(g.tangents[3 * i], g.tangents[3 * i + 1], g.tangents[3 * i + 2]); };
            V3 t0 = tan(prim_id3 + 0), t1 = tan(prim_id3 + 1), t2 = tan(prim_id3 + 2);
            auto fin = [](V3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); };
            if (!(fin(t0) && fin(t1) && fin(t2))) use_default = true;
            else t = normalize(interp3(t0, t1, t2));
        } else {
            use_default = true;
        }
        if (use_default) {
            V2 duv02 = {uv0.x - uv2.x, uv0.y - uv2.y};
            V2 duv12 = {uv1.x - uv2.x, uv1.y - uv2.y};
            V3 dp02 = v0 - v2, dp12 = v1 - v2;
            float determinant = difference_of_products(duv02.x, duv12.y, duv02.y, duv12.x);
            bool degenerate_uv = std::fabs(determinant) < 1e-8f;
            if (!degenerate_uv) {
                float inv_det = 1.0f / determinant;
                t.x = difference_of_products(duv12.y, dp02.x, duv02.y, dp12.x) * inv_det;
                t.y = difference_of_products(duv12.y, dp02.y, duv02.y, dp12.y) * inv_det;
                t.z = difference_of_products(duv12.y, dp02.z, duv02.y, dp12.z) * inv_det;
            }
            if (degenerate_uv || length_squared(t) == 0.0f) t = frame_from_n(ng_local).t;
        }
        tt_local = t;
    }
    V3 ns_local;
    if (g.normals) {
        auto nor = [&](uint32_t i) { return v3(g.normals[3 * i], g.normals[3 * i + 1], g.normals[3 * i + 2]); };
        ns_local = interp3(nor(prim_id3 + 0), nor(prim_id3 + 1), nor(prim_id3 + 2));
    } else {
        ns_local = ng_local;
    }
    // apply transform (:608-628)
    const float *mm = inst.m;
    V3 tr = v3(mm[12], mm[13], mm[14]);
    M3 m{v3(mm[0], mm[1], mm[2]), v3(mm[4], mm[5], mm[6]), v3(mm[8], mm[9], mm[10])};
    V3 p = mul(m, p_local) + tr;
    V3 tt = mul(m, tt_local);
    V3 c = mul(m, ng_local);
    M3 m_inv_t = inverse(transpose(m));
    V3 ng = normalize(mul(m_inv_t, ng_local));
    V3 ns = normalize(mul(m_inv_t, ns_local));
    float area = (area_local == 0.0f || inst.transform_det == 0.0f) ? 0.0f : std::fabs(area_local * inst.transform_det / dot(ng, c));
    Frame frame = (tt.x != 0.0f || tt.y != 0.0f || tt.z != 0.0f) ? frame_from_n_t(ns, tt) : frame_from_n(ns);
    SurfaceInteraction si;
    si.frame = frame;
    si.p = p;
    si.ng = ng;
    si.bary = bary;
    si.uv = uv;
    si.inst_id = inst_id;
    si.prim_id = prim_id;
    si.surface = material;
    si.mat_index = mat_index;
    si.prim_area = area;
    si.valid = true;
    return si;
}

// ---- ray / triangle (ASSUMED: third-party traversal; restated as Moeller-Trumbore on world-space
// triangles, closest hit = min (t, inst, prim), hit range t_min < t < t_max) ---------------------
struct Hit {
    bool hit;
    uint32_t inst, prim;
    V2 bary;
    float t;
};
inline bool tri_intersect(const WorldTri &tr, const Ray &ray, float t_max, float &t_out, V2 &bary) {
    V3 e1 = tr.v1 - tr.v0, e2 = tr.v2 - tr.v0;
    V3 pvec = cross(ray.d, e2);
    float det = dot(e1, pvec);
    if (det == 0.0f) return false;
    float inv_det = 1.0f / det;
    V3 tvec = ray.o - tr.v0;
    float u = dot(tvec, pvec) * inv_det;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    V3 qvec = cross(tvec, e1);
    float v = dot(ray.d, qvec) * inv_det;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    float t = dot(e2, qvec) * inv_det;
    if (!(t > ray.t_min && t < t_max)) return false;
    t_out = t;
    bary = {u, v};
    return true;
}
// alpha test (scene.rs:49-86): base-colour alpha of the hit material, stochastic by hash
inline bool alpha_test(const Scene &sc, uint32_t inst_id, uint32_t prim_id, V2 bary) {
    const OInstance &inst = sc.instances[inst_id];
    if (!inst.any_alpha) return true;  // alpha >= 1 passes regardless of the hash
    const AkrMesh &g = sc.desc->meshes[inst.geom_id];
    uint32_t mat_index = (inst.flags & AKR_MESH_HAS_MULTI_MATERIALS) ? g.material_slots[prim_id] : 0u;
    float alpha = inst.alphas[mat_index];
    if (inst.evals[mat_index].dynamic) {
        // surface_interaction_for_alpha_test (mesh.rs:426-485): uv only; NOTE the default third corner is (0, 0.1) here
        // and (1, 0.1) in surface_interaction (mesh.rs:456-467 vs :541-546) — preserved
        V2 uv0, uv1, uv2;
        uint32_t p3 = prim_id * 3;
        if (g.uvs) {
            uv0 = {g.uvs[2 * (p3 + 0)], g.uvs[2 * (p3 + 0) + 1]};
            uv1 = {g.uvs[2 * (p3 + 1)], g.uvs[2 * (p3 + 1) + 1]};
            uv2 = {g.uvs[2 * (p3 + 2)], g.uvs[2 * (p3 + 2) + 1]};
        } else {
            uv0 = {0.0f, 0.0f};
            uv1 = {1.0f, 0.0f};
            uv2 = {0.0f, 0.1f};
        }
        float w = 1.0f - bary.x - bary.y;
        V2 uv{w * uv0.x + bary.x * uv1.x + bary.y * uv2.x, w * uv0.y + bary.x * uv1.y + bary.y * uv2.y};
        ShaderEval ev;
        eval_shader(sc, inst.materials[mat_index], uv, ev, nullptr);
        alpha = 1.0f;  // SvmEvalMode::Alpha (principled.rs:15-22, diffuse.rs:85-92); other closures keep Surface::alpha() = 1
        if (ev.out_op == AKR_SVM_PRINCIPLED_BSDF) alpha = ev.principled.alpha;
        if (ev.out_op == AKR_SVM_DIFFUSE_BSDF) alpha = ev.diffuse_alpha;
    }
    uint32_t h = xxhash32_4(inst_id, prim_id, f2u(bary.x), f2u(bary.y));
    float hf = static_cast<float>(h) * static_cast<float>(1.0 / static_cast<double>(UINT32_MAX));
    return (alpha >= 1.0f) || (alpha > hf);
}
Hit trace_closest(const Scene &sc, const Ray &ray) {  // scene.rs:88-110
    Hit best{false, UINT32_MAX, UINT32_MAX, {0, 0}, ray.t_max};
    for (const WorldTri &tr : sc.tris) {
        if (!((tr.inst != ray.ex0_inst || tr.prim != ray.ex0_prim) && (tr.inst != ray.ex1_inst || tr.prim != ray.ex1_prim))) continue;
        float t;
        V2 b;
        // `<=` bound so that an equal-t later triangle is examined; ties resolved to the lower (inst, prim),
        // which is the earlier one in `tris` order, hence strict `<` on t keeps the first.
        if (!tri_intersect(tr, ray, ray.t_max, t, b)) continue;
        if (best.hit && !(t < best.t)) continue;
        if (!alpha_test(sc, tr.inst, tr.prim, b)) continue;
        best = {true, tr.inst, tr.prim, b, t};
    }
    return best;
}
bool trace_any(const Scene &sc, const Ray &ray) {  // scene.rs:155-185
    for (const WorldTri &tr : sc.tris) {
        if (!((tr.inst != ray.ex0_inst || tr.prim != ray.ex0_prim) && (tr.inst != ray.ex1_inst || tr.prim != ray.ex1_prim))) continue;
        float t;
        V2 b;
        if (!tri_intersect(tr, ray, ray.t_max, t, b)) continue;
        if (!alpha_test(sc, tr.inst, tr.prim, b)) continue;
        return true;
    }
    return false;
}

// ---- emission via dispatch_surface (light/area.rs:19-33) ---------------------------------------
Color surface_emission(const Scene &sc, const SurfaceInteraction &si, V3 wo, bool force_diffuse_unused) {
    (void)force_diffuse_unused;  // AreaLight::emission always dispatches the real shader
    return with_surface_closure(sc, si, false, [&](const auto &closure) { return closure.emission(wo); });
}

// ---- scene preparation ------------------------------------------------------------------------
bool prepare_scene(Scene &sc, const AkrSceneDesc *desc, const float *albedo_table) {
    sc.desc = desc;
    sc.albedo.data = albedo_table;
    if (desc->abi_version != AKR_B200_ABI_VERSION) {
        sc.error = "abi version mismatch";
        return false;
    }
    sc.instances.resize(desc->n_instances);
    uint32_t tri_offset = 0;
    for (uint32_t i = 0; i < desc->n_instances; ++i) {
        const AkrInstance &in = desc->instances[i];
        OInstance &o = sc.instances[i];
        std::memcpy(o.m, in.transform, sizeof(o.m));
        o.transform_det = mat4_det(o.m);
        o.geom_id = in.geom_id;
        o.flags = in.flags;
        o.materials = in.materials;
        o.n_materials = in.n_materials;
        o.light_id = -1;
        o.tri_offset = tri_offset;
        const AkrMesh &g = desc->meshes[in.geom_id];
        V3 tr = v3(o.m[12], o.m[13], o.m[14]);
        M3 m{v3(o.m[0], o.m[1], o.m[2]), v3(o.m[4], o.m[5], o.m[6]), v3(o.m[8], o.m[9], o.m[10])};
        for (uint32_t t = 0; t < g.n_triangles; ++t) {
            const uint32_t *idx = g.indices + 3 * t;
            auto vert = [&](uint32_t k) { return v3(g.vertices[3 * k], g.vertices[3 * k + 1], g.vertices[3 * k + 2]); };
            WorldTri w;
            w.v0 = mul(m, vert(idx[0])) + tr;
            w.v1 = mul(m, vert(idx[1])) + tr;
            w.v2 = mul(m, vert(idx[2])) + tr;
            w.inst = i;
            w.prim = t;
            sc.tris.push_back(w);
        }
        tri_offset += g.n_triangles;
        o.evals.resize(in.n_materials);
        o.alphas.assign(in.n_materials, 1.0f);
        for (uint32_t k = 0; k < in.n_materials; ++k) {
            std::string err;
            if (!eval_shader(sc, in.materials[k], V2{0.0f, 0.0f}, o.evals[k], &err)) {
                sc.error = err;
                return false;
            }
            // SvmEvalMode::Alpha: principled -> alpha of base_color (principled.rs:15-22); diffuse -> alpha of
            // the reflectance colour (diffuse.rs:85-92); emission / glass keep Surface::alpha() = 1 (mod.rs:54-56)
            if (o.evals[k].out_op == AKR_SVM_PRINCIPLED_BSDF) o.alphas[k] = o.evals[k].principled.alpha;
            if (o.evals[k].out_op == AKR_SVM_DIFFUSE_BSDF) o.alphas[k] = o.evals[k].diffuse_alpha;
            if (!(o.alphas[k] >= 1.0f) || o.evals[k].dynamic) o.any_alpha = true;
        }
    }
    // mesh lights (load.rs:312-415): per-triangle power = mean over 16 samples of max(emission) * area.
    // Emission of every supported closure is direction- and position-independent, so the 16 PCG32-driven
    // samples (load.rs:319-341) all contribute the same value; the f32 accumulation is kept literal.
    std::vector<float> light_weights;
    for (uint32_t i = 0; i < desc->n_instances; ++i) {
        OInstance &o = sc.instances[i];
        const AkrMesh &g = desc->meshes[o.geom_id];
        std::vector<float> powers(g.n_triangles);
        for (uint32_t t = 0; t < g.n_triangles; ++t) {
            SurfaceInteraction si = surface_interaction(sc, i, t, V2{1.0f / 3.0f, 1.0f / 3.0f});
            Color e = surface_emission(sc, si, si.frame.n, false);
            float acc = 0.0f;
            for (int s = 0; s < 16; ++s) acc += reduce_max(e) * si.prim_area;
            powers[t] = acc / 16.0f;
        }
        float total_power = 0.0f;
        for (float x : powers) total_power += x;
        if (total_power > 1e-4f) {
            o.light_id = static_cast<int32_t>(sc.lights.size());
            sc.lights.push_back({i, o.geom_id});
            light_weights.push_back(total_power);
            o.area_sampler.build(powers);
        }
    }
    if (!light_weights.empty()) sc.light_distribution.build(light_weights);
    // camera (camera/mod.rs:119-153)
    const AkrPerspectiveCamera &cam = desc->camera;
    std::memcpy(sc.camera.c2w, cam.c2w, sizeof(cam.c2w));
    sc.camera.width = cam.width;
    sc.camera.height = cam.height;
    {
        // m = T(0,0,-1) * S(aspect) * S(1,-1,1) * T(-1,-1,0) * S(2,2,1) * S(1/w,1/h,1); all factors are
        // scale/translate, so the glam products reduce to the per-axis scalar chains below (same op order).
        float fx = static_cast<float>(cam.width), fy = static_cast<float>(cam.height);
        float sx = 1.0f / fx, sy = 1.0f / fy, sz = 1.0f;
        float tx = 0.0f, ty = 0.0f, tz = 0.0f;
        auto scale = [&](float a, float b, float c) {
            sx = a * sx; sy = b * sy; sz = c * sz;
            tx = a * tx; ty = b * ty; tz = c * tz;
        };
        auto translate = [&](float a, float b, float c) { tx = tx + a; ty = ty + b; tz = tz + c; };
        scale(2.0f, 2.0f, 1.0f);
        translate(-1.0f, -1.0f, 0.0f);
        scale(1.0f, -1.0f, 1.0f);
        float s = std::tan(cam.fov / 2.0f);
        if (cam.width > cam.height) scale(s, s * fy / fx, 1.0f);
        else scale(s * fx / fy, s, 1.0f);
        translate(0.0f, 0.0f, -1.0f);
        float *r = sc.camera.r2c;
        std::memset(r, 0, sizeof(float) * 16);
        r[0] = sx; r[5] = sy; r[10] = sz; r[12] = tx; r[13] = ty; r[14] = tz; r[15] = 1.0f;
    }
    return true;
}

// AffineTransform::{transform_point, transform_vector} (geometry.rs:228-247); close_to_identity from
// AffineTransform::from_matrix (geometry.rs:212-218)
inline bool close_to_identity(const float *m) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float id = (c == r) ? 1.0f : 0.0f;
            if (!(std::fabs(m[c * 4 + r] - id) <= 1e-4f)) return false;
        }
    return true;
}
inline V3 transform_point(const float *m, V3 p) {
    if (close_to_identity(m)) return p;
    // Mat4 * Float4: col0*x + col1*y + col2*z + col3*w
    float q[4];
    for (int r = 0; r < 4; ++r) q[r] = m[0 + r] * p.x + m[4 + r] * p.y + m[8 + r] * p.z + m[12 + r] * 1.0f;
    return v3(q[0] / q[3], q[1] / q[3], q[2] / q[3]);
}
inline V3 transform_vector(const float *m, V3 v) {
    if (close_to_identity(m)) return v;
    float q[3];
    for (int r = 0; r < 3; ++r) q[r] = m[0 + r] * v.x + m[4 + r] * v.y + m[8 + r] * v.z + m[12 + r] * 0.0f;
    return v3(q[0], q[1], q[2]);
}

// PixelFilter::sample (film.rs:32-49)
inline V2 filter_sample(const AkrFilterConfig &f, V2 u) {
    if (f.type == AKR_FILTER_BOX) return {(u.x - 0.5f) * f.radius, (u.y - 0.5f) * f.radius};
    float width = f.radius;
    float sigma = width / 3.0f;
    float r = std::sqrt(-2.0f * std::log(u.x));
    float theta = 2.0f * PI * u.y;
    V2 offset = {r * std::cos(theta) * sigma, r * std::sin(theta) * sigma};
    return {clampf(offset.x, -width, width), clampf(offset.y, -width, width)};
}

// PerspectiveCamera::generate_ray (camera/mod.rs:70-103)
inline Ray generate_ray(const Scene &sc, const AkrFilterConfig &filter, uint32_t px, uint32_t py, Pmj02BnSampler &sampler) {
    V2 fpixel = {static_cast<float>(px) + 0.5f, static_cast<float>(py) + 0.5f};
    V2 offset = filter_sample(filter, sampler.next_2d());
    V2 p_film = {fpixel.x + offset.x, fpixel.y + offset.y};
    Ray ray;
    ray.o = v3s(0.0f);
    ray.d = normalize(transform_point(sc.camera.r2c, v3(p_film.x, p_film.y, 0.0f)));
    ray.t_min = 0.0f;
    ray.t_max = 1e20f;
    ray.ex0_inst = ray.ex0_prim = ray.ex1_inst = ray.ex1_prim = UINT32_MAX;
    ray.o = transform_point(sc.camera.c2w, ray.o);
    ray.d = transform_vector(sc.camera.c2w, ray.d);
    return ray;
}

// ---- lights (light/mod.rs:100-147, light/area.rs:36-130) -----------------------------------------
struct DirectLighting {  // pt.rs:57-77
    Color irradiance;
    V3 wi;
    float pdf;
    Ray shadow_ray;
    bool valid;
};
DirectLighting sample_light(const Scene &sc, const AkrPtConfig &cfg, uint32_t depth, const SurfaceInteraction &si, V3 u) {  // pt.rs:170-209
    DirectLighting invalid{v3s(0.0f), v3s(0.0f), 0.0f, Ray{}, false};
    if (!cfg.use_nee) return invalid;
    if (!(!cfg.indirect_only || depth > 1)) return invalid;
    if (sc.lights.empty()) return invalid;
    V3 pn_p = si.p, pn_n = si.ng;
    // LightAggregate::sample_direct (light/mod.rs:115-132)
    AliasTable::Sample ls = sc.light_distribution.sample_and_remap(u.x);
    float light_choice_pdf = ls.pdf;
    const OLight &light = sc.lights[ls.idx];
    // AreaLight::sample_direct (area.rs:51-107)
    const OInstance &linst = sc.instances[light.instance_id];
    AliasTable::Sample ps = linst.area_sampler.sample_and_remap(ls.u);
    uint32_t prim_id = ps.idx;
    float pdf = ps.pdf;
    V2 bary = uniform_sample_triangle(V2{u.y, u.z});
    SurfaceInteraction lsi = surface_interaction(sc, light.instance_id, prim_id, bary);
    float area = lsi.prim_area;
    V3 p = lsi.p, n = lsi.ng;
    V3 wi = p - pn_p;
    if (length_squared(wi) == 0.0f) return invalid;
    float dist2 = length_squared(wi);
    wi = wi / std::sqrt(dist2);
    Color emission = surface_emission(sc, lsi, -wi, false);
    Color li = dot(wi, n) < 0.0f ? emission : v3s(0.0f);
    float cos_theta_i = std::fabs(dot(n, wi));
    pdf = pdf / area * dist2 / cos_theta_i;
    V3 ro = offset_ray_origin(pn_p, face_forward(pn_n, wi));
    float dist = std::sqrt(dist2);
    Ray shadow;
    shadow.o = ro;
    shadow.d = wi;
    shadow.t_min = 0.0f;
    shadow.t_max = dist * (1.0f - 1e-3f);
    shadow.ex0_inst = UINT32_MAX;
    shadow.ex0_prim = UINT32_MAX;
    shadow.ex1_inst = light.instance_id;
    shadow.ex1_prim = prim_id;
    bool valid = std::isfinite(pdf);
    pdf = pdf * light_choice_pdf;  // light/mod.rs:130
    if (!valid) return invalid;
    shadow.ex0_inst = si.inst_id;  // pt.rs:189-190
    shadow.ex0_prim = si.prim_id;
    return {li, wi, pdf, shadow, true};
}
inline float mis_weight(float pdf_a, float pdf_b) { return pdf_a / (pdf_a + pdf_b); }  // pt.rs:962-973 with power = 1

struct PathStats {
    uint64_t segments = 0, shadow_rays = 0;
};

// ---- PathTracerBase::run_megakernel (pt.rs:325-327 -> 329-900 with shift_mapping = None) ---------
Color radiance(const Scene &sc, const AkrPtConfig &cfg, Ray ray, Pmj02BnSampler &sampler, PathStats &st, uint32_t *first_hit) {
    Color L = v3s(0.0f), beta = v3s(1.0f), base_replay_throughput = v3s(0.0f);
    uint32_t depth = 0;
    float prev_bsdf_pdf = 0.0f;
    V3 prev_ng = v3s(0.0f);
    const bool has_debug_depth = cfg.debug_depth >= 0;
    auto add_radiance = [&](Color r) {  // pt.rs:133-149
        if (has_debug_depth) {
            if (depth == static_cast<uint32_t>(cfg.debug_depth)) L = L + beta * r;
        } else {
            L = L + beta * r;
        }
    };
    while (true) {
        st.segments += 1;
        Hit hit = trace_closest(sc, ray);
        if (first_hit && depth == 0) {
            first_hit[0] = hit.hit ? hit.inst : UINT32_MAX;
            first_hit[1] = hit.hit ? hit.prim : UINT32_MAX;
        }
        if (!hit.hit) {
            add_radiance(v3s(0.0f) * 0.0f);  // hit_envmap = (0, 0)  (pt.rs:226-228,386-388)
            break;
        }
        SurfaceInteraction si = surface_interaction(sc, hit.inst, hit.prim, hit.bary);
        V3 wo = -ray.d;
        {  // handle_surface_light (pt.rs:230-258)
            Color direct = v3s(0.0f);
            float w = 0.0f;
            const OInstance &inst = sc.instances[si.inst_id];
            if (inst.light_id >= 0 && (!cfg.indirect_only || depth > 1)) {
                // AreaLight::le (area.rs:36-49)
                Color emission = surface_emission(sc, si, -ray.d, false);
                direct = dot(si.ng, ray.d) < 0.0f ? emission : v3s(0.0f);
                if (depth == 0 || !cfg.use_nee) {
                    w = 1.0f;
                } else {
                    // LightAggregate::pdf_direct (light/mod.rs:134-147) + AreaLight::pdf_direct (area.rs:109-130)
                    V3 pn_p = ray.o;
                    float light_choice_pdf = sc.light_distribution.pdf[static_cast<size_t>(inst.light_id)];
                    float prim_pdf = inst.area_sampler.pdf[si.prim_id];
                    V3 wi = si.p - pn_p;
                    float dist2 = length_squared(wi);
                    wi = wi / std::sqrt(dist2);
                    float pdf = prim_pdf / si.prim_area * dist2 / std::fmax(std::fabs(dot(si.ng, wi)), 1e-6f);
                    float light_pdf = light_choice_pdf * pdf;
                    w = mis_weight(prev_bsdf_pdf, light_pdf);
                    (void)prev_ng;
                }
            }
            add_radiance(direct * w);
        }
        if (depth == 0) base_replay_throughput = L;
        if (depth >= cfg.max_depth) break;
        depth += 1;
        V3 u_direct = sampler.next_3d();
        DirectLighting dl = sample_light(sc, cfg, depth, si, u_direct);
        V3 u_bsdf = sampler.next_3d();
        // sample_surface_and_shade_direct (pt.rs:297-323)
        Color direct = v3s(0.0f);
        BsdfSample bs = with_surface_closure(sc, si, cfg.force_diffuse != 0, [&](const auto &closure) {
            if (dl.valid) {
                Eval e = closure.evaluate(wo, dl.wi);
                float w = mis_weight(dl.pdf, e.pdf);
                direct = dl.irradiance * e.f * w / dl.pdf;
            }
            return closure_sample(closure, wo, u_bsdf.x, V2{u_bsdf.y, u_bsdf.z});
        });
        if (dl.valid) {  // pt.rs:504-513
            st.shadow_rays += 1;
            bool occluded = trace_any(sc, dl.shadow_ray);
            if (!occluded) add_radiance(direct);
            if (depth == 1) base_replay_throughput = L;
        }
        beta = beta * (bs.color / bs.pdf);  // mul_beta(f / pdf)  (pt.rs:783)
        if (bs.pdf <= 0.0f || !bs.valid || reduce_min(bs.color) < 0.0f) break;  // pt.rs:832-842
        if (depth > cfg.rr_depth) {  // pt.rs:211-218,843-850
            float cont_prob = clampf(reduce_max(beta), 0.0f, 1.0f) * 0.95f;
            bool rr = sampler.next_1d() >= cont_prob;
            if (rr) break;
            beta = beta * (v3s(1.0f) / cont_prob);
        }
        prev_bsdf_pdf = bs.pdf;
        prev_ng = si.ng;
        V3 ro = offset_ray_origin(si.p, face_forward(si.ng, bs.wi));
        ray.o = ro;
        ray.d = bs.wi;
        ray.t_min = 0.0f;
        ray.t_max = 1e20f;
        ray.ex0_inst = si.inst_id;
        ray.ex0_prim = si.prim_id;
        ray.ex1_inst = UINT32_MAX;
        ray.ex1_prim = UINT32_MAX;
    }
    {  // clamp_indirect = 1000 (pt.rs:130,871-876); Color::clamp clamps to [0, max]
        Color indirect = L - base_replay_throughput;
        indirect = v3(clampf(indirect.x, 0.0f, 1000.0f), clampf(indirect.y, 0.0f, 1000.0f), clampf(indirect.z, 0.0f, 1000.0f));
        L = base_replay_throughput + indirect;
    }
    return L;
}

struct TapClosure {
    int kind;
    Color color;
    float roughness, eta;
};
template <class F> auto with_tap(const TapClosure &t, F &&f) {
    TrowbridgeReitz dist = TrowbridgeReitz::from_roughness(t.roughness, t.roughness);
    if (t.kind == 0) return f(DiffuseBsdf{t.color * FRAC_1_PI});
    if (t.kind == 1) return f(MicrofacetReflection<FresnelDielectric>{t.color, FresnelDielectric{t.eta}, dist});
    if (t.kind == 2) return f(MicrofacetTransmission{dist, t.color, t.eta, FresnelDielectric{t.eta}});
    Color n, k;
    artistic_to_conductor_fresnel(t.color, v3s(1.0f), n, k);
    return f(MicrofacetReflection<FresnelComplex>{v3s(1.0f), FresnelComplex{n, k}, dist});
}
}  // namespace

// ============================================================================================
// C interface (ctypes)
// ============================================================================================
extern "C" {

typedef struct AkrOracleStats {
    uint64_t samples, segments, shadow_rays;
    double seconds;
    uint32_t threads;
    uint32_t n_lights;
} AkrOracleStats;

static thread_local std::string g_err;
const char *akr_oracle_last_error(void) { return g_err.c_str(); }

// Renders samples [spp_begin, spp_end) of every pixel in rows [y0, y1) and ACCUMULATES into
// `film_7n` (reference Film layout, film.rs:66-76), which the caller zero-initialises.
// cfg->spp is the total spp the sampler is configured for (sampler/mod.rs:381-386; pt.rs:1074).
// first_hits (optional): [rows*width][2] (inst, prim) of the depth-0 hit of sample `spp_begin`.
int akr_oracle_render(const AkrSceneDesc *scene, const AkrPtConfig *cfg, const AkrSamplerConfig *sampler_cfg,
                      const AkrFilterConfig *filter, const uint32_t *pmj02bn, const uint16_t *bluenoise,
                      const float *albedo_table, uint32_t y0, uint32_t y1, uint32_t spp_begin, uint32_t spp_end,
                      int n_threads, float *film_7n, uint32_t *first_hits, AkrOracleStats *stats) {
    if (!scene || !cfg || !sampler_cfg || !filter || !pmj02bn || !bluenoise || !albedo_table || !film_7n) {
        g_err = "null argument";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    if (sampler_cfg->type != AKR_SAMPLER_PMJ02BN) {
        g_err = "only the pmj02bn sampler is reproducible (independent seeds from rand::StdRng, sampler/mod.rs:148-160)";
        return AKR_ERR_UNSUPPORTED;
    }
    if (cfg->spp > AKR_PMJ02BN_SAMPLES || cfg->spp == 0 || spp_end > cfg->spp || spp_begin > spp_end) {
        g_err = "bad spp range (Pmj02BnSampler supports up to 65536 spp, sampler/mod.rs:374-380)";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    Scene sc;
    if (!prepare_scene(sc, scene, albedo_table)) {
        g_err = sc.error;
        return AKR_ERR_UNSUPPORTED;
    }
    const uint32_t width = sc.camera.width, height = sc.camera.height;
    if (y1 > height || y0 >= y1) {
        g_err = "bad tile";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    uint32_t w = cfg->spp - 1;  // sampler/mod.rs:381-386
    w |= w >> 1;
    w |= w >> 2;
    w |= w >> 4;
    w |= w >> 8;
    w |= w >> 16;
    const uint32_t rows = y1 - y0;
    const size_t n = static_cast<size_t>(width) * rows;
    Tables tab{pmj02bn, bluenoise};
    if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
    if (n_threads <= 0) n_threads = 1;
    std::atomic<uint32_t> next_row{0};
    std::vector<PathStats> tstats(static_cast<size_t>(n_threads));
    auto t_start = std::chrono::steady_clock::now();
    auto worker = [&](int tid) {
        PathStats st;
        while (true) {
            uint32_t r = next_row.fetch_add(1);
            if (r >= rows) break;
            uint32_t y = y0 + r;
            for (uint32_t x = 0; x < width; ++x) {
                Pmj02BnSampler sampler{tab, static_cast<uint32_t>(sampler_cfg->seed), 0, x, y, UINT32_MAX, cfg->spp, w};
                // state persists across passes (sampler/mod.rs:443-457,637-645): resume at spp_begin
                if (spp_begin > 0) sampler.sample_index = spp_begin - 1;
                size_t i = static_cast<size_t>(x) + static_cast<size_t>(r) * width;
                for (uint32_t s = spp_begin; s < spp_end; ++s) {
                    sampler.start();
                    int32_t sx = static_cast<int32_t>(x) + cfg->pixel_offset[0];  // pt.rs:1084-1088
                    int32_t sy = static_cast<int32_t>(y) + cfg->pixel_offset[1];
                    sx = std::min(std::max(sx, 0), static_cast<int32_t>(width) - 1);
                    sy = std::min(std::max(sy, 0), static_cast<int32_t>(height) - 1);
                    Ray ray = generate_ray(sc, *filter, static_cast<uint32_t>(sx), static_cast<uint32_t>(sy), sampler);
                    uint32_t *fh = (first_hits && s == spp_begin) ? first_hits + 2 * i : nullptr;
                    Color l = radiance(sc, *cfg, ray, sampler, st, fh);
                    if (has_nan(l)) l = v3s(0.0f);  // Film::add_sample -> remove_nan (film.rs:196-206)
                    l = l * 1.0f;                    // * weight (ray_w = 1)
                    film_7n[i * 3 + 0] += l.x;
                    film_7n[i * 3 + 1] += l.y;
                    film_7n[i * 3 + 2] += l.z;
                    film_7n[6 * n + i] += 1.0f;
                }
            }
        }
        tstats[static_cast<size_t>(tid)] = st;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
    auto t_end = std::chrono::steady_clock::now();
    if (stats) {
        stats->samples = static_cast<uint64_t>(n) * (spp_end - spp_begin);
        stats->segments = 0;
        stats->shadow_rays = 0;
        for (const PathStats &s : tstats) {
            stats->segments += s.segments;
            stats->shadow_rays += s.shadow_rays;
        }
        stats->seconds = std::chrono::duration<double>(t_end - t_start).count();
        stats->threads = static_cast<uint32_t>(n_threads);
        stats->n_lights = static_cast<uint32_t>(sc.lights.size());
    }
    return AKR_OK;
}

// The `aov` integrator (crates/akari_integrator/src/aov.rs:52-185): per camera sample the chosen first-hit quantity is
// added to the film (remove_nan, weight 1) instead of radiance.  Single pass over samples [0, cfg->spp).
int akr_oracle_render_aov(const AkrSceneDesc *scene, const AkrAovConfig *cfg, const AkrSamplerConfig *sampler_cfg, const AkrFilterConfig *filter,
                          const uint32_t *pmj02bn, const uint16_t *bluenoise, const float *albedo_table, uint32_t y0, uint32_t y1, float *film_7n) {
    if (!scene || !cfg || !sampler_cfg || !filter || !pmj02bn || !bluenoise || !albedo_table || !film_7n) {
        g_err = "null argument";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    if (sampler_cfg->type != AKR_SAMPLER_PMJ02BN || cfg->spp == 0 || cfg->spp > AKR_PMJ02BN_SAMPLES || cfg->aov > AKR_AOV_ROUGHNESS) {
        g_err = "unsupported aov configuration";
        return AKR_ERR_UNSUPPORTED;
    }
    Scene sc;
    if (!prepare_scene(sc, scene, albedo_table)) {
        g_err = sc.error;
        return AKR_ERR_UNSUPPORTED;
    }
    const uint32_t width = sc.camera.width, height = sc.camera.height;
    if (y1 > height || y0 >= y1) {
        g_err = "bad tile";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    uint32_t w = cfg->spp - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    const uint32_t rows = y1 - y0;
    const size_t n = static_cast<size_t>(width) * rows;
    Tables tab{pmj02bn, bluenoise};
    auto remap = [&](V3 v) { return cfg->remap ? v * 0.5f + v3s(0.5f) : v; };
    for (uint32_t r = 0; r < rows; ++r)
        for (uint32_t x = 0; x < width; ++x) {
            const uint32_t y = y0 + r;
            Pmj02BnSampler sampler{tab, static_cast<uint32_t>(sampler_cfg->seed), 0, x, y, UINT32_MAX, cfg->spp, w};
            const size_t i = static_cast<size_t>(x) + static_cast<size_t>(r) * width;
            for (uint32_t s = 0; s < cfg->spp; ++s) {
                sampler.start();
                Ray ray = generate_ray(sc, *filter, x, y, sampler);
                Hit hit = trace_closest(sc, ray);
                Color color = v3s(0.0f);
                if (hit.hit) {
                    SurfaceInteraction si = surface_interaction(sc, hit.inst, hit.prim, hit.bary);
                    const V3 wo = -ray.d;
                    switch (cfg->aov) {
                    case AKR_AOV_SHADING_NORMAL: color = remap(with_surface_closure(sc, si, false, [&](const auto &c) { return c.ns(); })); break;
                    case AKR_AOV_GEOMETRY_NORMAL: color = remap(si.ng); break;
                    case AKR_AOV_TANGENT: color = remap(si.frame.t); break;
                    case AKR_AOV_BITANGENT: color = remap(si.frame.s); break;
                    case AKR_AOV_ALBEDO: color = with_surface_closure(sc, si, false, [&](const auto &c) { return c.albedo(wo) + c.emission(wo); }); break;
                    default: {
                        const float u = sampler.next_1d();
                        color = v3s(1.0f) * with_surface_closure(sc, si, false, [&](const auto &c) { return c.roughness(wo, u); });
                    }
                    }
                }
                if (has_nan(color)) color = v3s(0.0f);
                color = color * 1.0f;
                film_7n[i * 3 + 0] += color.x;
                film_7n[i * 3 + 1] += color.y;
                film_7n[i * 3 + 2] += color.z;
                film_7n[6 * n + i] += 1.0f;
            }
        }
    return AKR_OK;
}

// Film::copy_to_rgba_image(hdr = true) (film.rs:120-148), splat_scale = 1
void akr_oracle_resolve(const float *film_7n, size_t n_pixels, float *rgb_out) {
    for (size_t i = 0; i < n_pixels; ++i) {
        float w = film_7n[6 * n_pixels + i];
        float d = (w == 0.0f) ? 1.0f : w;
        for (int c = 0; c < 3; ++c) rgb_out[i * 3 + c] = film_7n[i * 3 + c] / d + film_7n[3 * n_pixels + i * 3 + c] * 1.0f;
    }
}

// ---- known-answer taps ----------------------------------------------------------------------
uint32_t akr_oracle_xxhash32_4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return xxhash32_4(x, y, z, w); }
uint32_t akr_oracle_permute_element(uint32_t i, uint32_t l, uint32_t w, uint32_t p) { return permute_element(i, l, w, p); }

// The first `n_dims_pattern` draws of sample `sample_index` of pixel (px, py): pattern[k] = 1 -> next_1d,
// 2 -> next_2d, 3 -> next_3d; out receives the floats in draw order.
int akr_oracle_sampler_stream(const uint32_t *pmj02bn, const uint16_t *bluenoise, uint32_t seed, uint32_t spp, uint32_t px,
                              uint32_t py, uint32_t sample_index, const uint8_t *pattern, uint32_t n_pattern, float *out) {
    uint32_t w = spp - 1;
    w |= w >> 1;
    w |= w >> 2;
    w |= w >> 4;
    w |= w >> 8;
    w |= w >> 16;
    Pmj02BnSampler s{Tables{pmj02bn, bluenoise}, seed, 0, px, py, sample_index == 0 ? UINT32_MAX : sample_index - 1, spp, w};
    s.start();
    size_t o = 0;
    for (uint32_t k = 0; k < n_pattern; ++k) {
        if (pattern[k] == 1) out[o++] = s.next_1d();
        else if (pattern[k] == 2) {
            V2 u = s.next_2d();
            out[o++] = u.x;
            out[o++] = u.y;
        } else if (pattern[k] == 3) {
            V3 u = s.next_3d();
            out[o++] = u.x;
            out[o++] = u.y;
            out[o++] = u.z;
        } else return AKR_ERR_INVALID_ARGUMENT;
    }
    return AKR_OK;
}

// AliasTable::new (util/distribution.rs:34-78)
void akr_oracle_alias_table(const float *weights, uint32_t n, uint32_t *j_out, float *t_out, float *pdf_out) {
    AliasTable at;
    at.build(std::vector<float>(weights, weights + n));
    for (uint32_t i = 0; i < n; ++i) {
        j_out[i] = at.j[i];
        t_out[i] = at.t[i];
        pdf_out[i] = at.pdf[i];
    }
}
// AliasTable::sample_and_remap
void akr_oracle_alias_sample(const uint32_t *j, const float *t, const float *pdf, uint32_t n, float u, uint32_t *idx_out,
                             float *pdf_out, float *u_out) {
    AliasTable at;
    at.j.assign(j, j + n);
    at.t.assign(t, t + n);
    at.pdf.assign(pdf, pdf + n);
    AliasTable::Sample s = at.sample_and_remap(u);
    *idx_out = s.idx;
    *pdf_out = s.pdf;
    *u_out = s.u;
}

// Camera ray of (px, py) for filter offset drawn from u (tests the raster->camera->world chain).
int akr_oracle_camera_ray(const AkrSceneDesc *scene, const AkrFilterConfig *filter, uint32_t px, uint32_t py, float u0, float u1,
                          float *o3, float *d3) {
    static const float zero_table[4096] = {0};
    Scene sc;
    if (!prepare_scene(sc, scene, zero_table)) {
        g_err = sc.error;
        return AKR_ERR_UNSUPPORTED;
    }
    V2 fpixel = {static_cast<float>(px) + 0.5f, static_cast<float>(py) + 0.5f};
    V2 offset = filter_sample(*filter, V2{u0, u1});
    V3 d = normalize(transform_point(sc.camera.r2c, v3(fpixel.x + offset.x, fpixel.y + offset.y, 0.0f)));
    V3 o = transform_point(sc.camera.c2w, v3s(0.0f));
    d = transform_vector(sc.camera.c2w, d);
    o3[0] = o.x; o3[1] = o.y; o3[2] = o.z;
    d3[0] = d.x; d3[1] = d.y; d3[2] = d.z;
    return AKR_OK;
}

// Scene-level derived data: light list and per-instance determinant / alias tables.
int akr_oracle_scene_lights(const AkrSceneDesc *scene, uint32_t *n_lights, uint32_t *light_instances, float *light_powers,
                            uint32_t max_lights) {
    static const float zero_table[4096] = {0};
    Scene sc;
    if (!prepare_scene(sc, scene, zero_table)) {
        g_err = sc.error;
        return AKR_ERR_UNSUPPORTED;
    }
    *n_lights = static_cast<uint32_t>(sc.lights.size());
    for (uint32_t i = 0; i < sc.lights.size() && i < max_lights; ++i) {
        light_instances[i] = sc.lights[i].instance_id;
        light_powers[i] = sc.light_distribution.pdf[i];
    }
    return AKR_OK;
}

// BSDF taps for chi-square / furnace tests (methodology of crates/akari_api/src/bin/akari_test.rs:31-219).
// kind: 0 = Diffuse(reflectance 1/pi * color), 1 = GGX reflection (dielectric fresnel, eta),
//       2 = GGX transmission (eta), 3 = GGX conductor (artistic n,k from color / white tint)
void akr_oracle_bsdf_eval(int kind, const float *color3, float roughness, float eta, const float *wo3, const float *wi3,
                          float *f3_out, float *pdf_out) {
    TapClosure t{kind, v3(color3[0], color3[1], color3[2]), roughness, eta};
    Eval e = with_tap(t, [&](const auto &c) { return c.evaluate(v3(wo3[0], wo3[1], wo3[2]), v3(wi3[0], wi3[1], wi3[2])); });
    f3_out[0] = e.f.x; f3_out[1] = e.f.y; f3_out[2] = e.f.z;
    *pdf_out = e.pdf;
}
void akr_oracle_bsdf_sample(int kind, const float *color3, float roughness, float eta, const float *wo3, float u_select, float u0,
                            float u1, float *wi3_out, int *valid_out) {
    TapClosure t{kind, v3(color3[0], color3[1], color3[2]), roughness, eta};
    SampleWi s = with_tap(t, [&](const auto &c) { return c.sample_wi(v3(wo3[0], wo3[1], wo3[2]), u_select, V2{u0, u1}); });
    wi3_out[0] = s.wi.x; wi3_out[1] = s.wi.y; wi3_out[2] = s.wi.z;
    *valid_out = s.valid ? 1 : 0;
}

// The two tables of the reference's chi-square test (akari_test.rs:31-112) for one tap closure and one wo: observed
// histogram of sample_wi directions and expected counts from the integrated pdf (chi2_tables.h).
void akr_oracle_bsdf_chi2_tables(int kind, const float *color3, float roughness, float eta, const float *wo3, uint64_t n_samples, uint64_t seed,
                                 uint32_t theta_res, uint32_t phi_res, uint32_t *hist_out, double *expected_out) {
    TapClosure t{kind, v3(color3[0], color3[1], color3[2]), roughness, eta};
    const V3 wo = v3(wo3[0], wo3[1], wo3[2]);
    chi2::histogram(
        [&](float us, float u0, float u1, chi2::Dir &wi) {
            SampleWi s = with_tap(t, [&](const auto &c) { return c.sample_wi(wo, us, V2{u0, u1}); });
            wi = chi2::Dir{s.wi.x, s.wi.y, s.wi.z};
            return s.valid;
        },
        n_samples, seed, theta_res, phi_res, hist_out, 0);
    chi2::expected([&](chi2::Dir wi) { return with_tap(t, [&](const auto &c) { return c.evaluate(wo, v3(wi.x, wi.y, wi.z)); }).pdf; }, n_samples, theta_res,
                   phi_res, expected_out, 0);
}

// Deterministic re-derivation of the `ggx_dielectric_s` table (svm/surface/precompute.rs:56-94,
// svm/surface/mod.rs:1338-1356).  The reference estimates each cell with 2^20 PCG32 samples seeded from
// rand::StdRng (not reproducible); here each cell is the mean over a fixed `n x n` midpoint grid of u.
void akr_oracle_make_albedo_table(float *table_16x16x16, uint32_t n) {
    const uint32_t dim = 16;
    for (uint32_t iz = 0; iz < dim; ++iz)
        for (uint32_t iy = 0; iy < dim; ++iy)
            for (uint32_t ix = 0; ix < dim; ++ix) {
                float fx = clampf(static_cast<float>(ix) / (static_cast<float>(dim) - 1.0f), 1e-4f, 0.9999f);
                float fy = clampf(static_cast<float>(iy) / (static_cast<float>(dim) - 1.0f), 1e-4f, 0.9999f);
                float fz = clampf(static_cast<float>(iz) / (static_cast<float>(dim) - 1.0f), 1e-4f, 0.9999f);
                float roughness = fx, mu = fy, ior = ior_parametrization(fz);
                MicrofacetReflection<FresnelDielectric> bsdf{v3s(1.0f), FresnelDielectric{ior},
                                                             TrowbridgeReitz::from_roughness(roughness, roughness)};
                SurfaceClosure<decltype(bsdf)> closure{bsdf, frame_from_n(v3(0, 0, 1)), v3(0, 0, 1)};
                V3 wo = v3(std::sqrt(1.0f - sqr(mu)), 0.0f, mu);
                double sum = 0.0;
                for (uint32_t a = 0; a < n; ++a)
                    for (uint32_t b = 0; b < n; ++b) {
                        V2 u = {(static_cast<float>(a) + 0.5f) / static_cast<float>(n), (static_cast<float>(b) + 0.5f) / static_cast<float>(n)};
                        BsdfSample s = closure_sample(closure, wo, 0.5f, u);
                        if (s.valid && s.pdf > 0.0f) sum += static_cast<double>(s.color.x / s.pdf);
                    }
                table_16x16x16[ix + iy * dim + iz * dim * dim] = static_cast<float>(sum / (static_cast<double>(n) * n));
            }
}

}  // extern "C"

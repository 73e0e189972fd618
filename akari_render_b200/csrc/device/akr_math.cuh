// akr_math.cuh — scalar/vector helpers shared by every kernel of the wavefront path tracer.
//
// The same header compiles under nvcc (device code, sm_100a) and under plain g++ (a host-side
// per-thread simulation of the kernels used by tests/hostsim to debug kernel logic in a container
// without a GPU; it is never part of the shipped library).
//
// Formulas follow the reference's geometry helpers, cited per function
// (reference root: /root/reference/crates/akari_render/src).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define AKR_HD __host__ __device__ __forceinline__
#define AKR_D __device__ __forceinline__
#define AKR_HD_NOINLINE __host__ __device__ __noinline__   // big cold bodies called from several places of a hot loop
#define AKR_NO_UNROLL _Pragma("unroll 1")
#else
#define AKR_HD inline
#define AKR_D inline
#define AKR_HD_NOINLINE inline
#define AKR_NO_UNROLL
#endif

// Read-only scene / table data.  Routing these through __ldg (LDG.E.CONSTANT) was measured on B200 and LOST 4 % in the
// Lambert shade kernel (18.4 -> 19.2 ms per pass), so the marker expands to a plain load.
#define AKR_RO(x) (x)

namespace akr {

struct f2 {
    float x, y;
};
struct f3 {
    float x, y, z;
};

AKR_HD f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
AKR_HD f3 splat3(float s) { return f3{s, s, s}; }
AKR_HD f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
AKR_HD f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
AKR_HD f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
AKR_HD f3 operator/(f3 a, f3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
AKR_HD f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
AKR_HD f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
AKR_HD f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
AKR_HD f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
AKR_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
AKR_HD f3 cross(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
AKR_HD float length_squared(f3 a) { return dot(a, a); }
AKR_HD float length(f3 a) { return sqrtf(dot(a, a)); }
// normalize = v * (1 / sqrt(dot)): IEEE sqrt + divide, identical on host and device
AKR_HD f3 normalize(f3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
AKR_HD float sqr(float x) { return x * x; }
AKR_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
AKR_HD float lerpf(float a, float b, float t) { return t * (b - a) + a; }
AKR_HD f3 lerp3(f3 a, f3 b, float t) { return {lerpf(a.x, b.x, t), lerpf(a.y, b.y, t), lerpf(a.z, b.z, t)}; }
AKR_HD float reduce_max(f3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
AKR_HD float reduce_min(f3 a) { return fminf(a.x, fminf(a.y, a.z)); }
AKR_HD f3 min3(f3 a, f3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
AKR_HD float avg3(f3 a) { return (a.x + a.y + a.z) / 3.0f; }
AKR_HD bool has_nan(f3 a) { return (a.x != a.x) || (a.y != a.y) || (a.z != a.z); }
AKR_HD bool is_finite(float x) { return fabsf(x) <= 3.402823466e+38f; }

AKR_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
AKR_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

#define AKR_PI 3.14159265358979323846f
#define AKR_FRAC_1_PI 0.318309886183790671537767526745028724f
#define AKR_ONE_MINUS_EPSILON 0x1.fffffep-1f /* lib.rs:59 */

// util/mod.rs:326-331
AKR_HD float difference_of_products(float a, float b, float c, float d) {
    float cd = c * d;
    float diff = fmaf(a, b, -cd);
    float err = fmaf(-c, d, cd);
    return diff + err;
}

// ---- shading frame (geometry.rs:72-200) ---------------------------------------------------------
struct Frame {
    f3 n, t, s;
};
AKR_HD float cos_theta(f3 w) { return w.z; }
AKR_HD float cos2_theta(f3 w) { return w.z * w.z; }
AKR_HD float abs_cos_theta(f3 w) { return fabsf(w.z); }
AKR_HD float sin2_theta(f3 w) { return fmaxf(1.0f - cos2_theta(w), 0.0f); }
AKR_HD float sin_theta(f3 w) { return sqrtf(fmaxf(1.0f - cos2_theta(w), 0.0f)); }
AKR_HD float tan2_theta(f3 w) { return sin2_theta(w) / cos2_theta(w); }
AKR_HD float tan_theta(f3 w) { return sin_theta(w) / cos_theta(w); }
// geometry.rs:122-152: sin_phi is derived from w.x and cos_phi from w.y, as in the reference
AKR_HD float sin_phi(f3 w) {
    float st = sin_theta(w);
    return st == 0.0f ? 0.0f : clampf(w.x / st, -1.0f, 1.0f);
}
AKR_HD float cos_phi(f3 w) {
    float st = sin_theta(w);
    return st == 0.0f ? 1.0f : clampf(w.y / st, -1.0f, 1.0f);
}
AKR_HD bool same_hemisphere(f3 a, f3 b) { return (a.z * b.z) >= 0.0f; }

AKR_HD Frame frame_identity() { return Frame{mk3(0, 0, 1), mk3(1, 0, 0), mk3(0, 1, 0)}; }
AKR_HD Frame frame_from_n(f3 n) {  // geometry.rs:159-167
    f3 t;
    if (fabsf(n.x) > fabsf(n.y)) t = mk3(-n.z, 0.0f, n.x) / sqrtf(n.x * n.x + n.z * n.z);
    else t = mk3(0.0f, n.z, -n.y) / sqrtf(n.y * n.y + n.z * n.z);
    f3 s = cross(n, t);
    return Frame{n, t, s};
}
AKR_HD Frame frame_from_n_t(f3 n, f3 tt_in) {  // geometry.rs:169-191
    Frame frame = Frame{splat3(0), splat3(0), splat3(0)};
    f3 tt = tt_in - n * dot(n, tt_in);
    bool good = true;
    if (length(tt) < 1e-4f) good = false;
    else tt = normalize(tt);
    if (good) {
        f3 ss = cross(n, tt);
        if (length(ss) < 1e-4f) good = false;
        else {
            ss = normalize(ss);
            frame = Frame{n, tt, ss};
        }
    }
    if (!good) frame = frame_from_n(n);
    return frame;
}
AKR_HD f3 to_world(const Frame &f, f3 v) { return f.t * v.x + f.s * v.y + f.n * v.z; }
AKR_HD f3 to_local(const Frame &f, f3 v) { return mk3(dot(f.t, v), dot(f.s, v), dot(f.n, v)); }
AKR_HD f3 face_forward(f3 v, f3 n) { return dot(v, n) < 0.0f ? -v : v; }  // geometry.rs:264-272
AKR_HD f3 reflect(f3 w, f3 n) { return -w + 2.0f * dot(w, n) * n; }       // geometry.rs:277-281

// luisa::rtx::offset_ray_origin (third-party; Waechter & Binder, Ray Tracing Gems ch. 6).
// Call sites: akari_integrator/src/pt.rs:856, light/area.rs:87.
AKR_HD float offset_axis(float p, float n) {
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    int32_t of_i = (int32_t)(int_scale * n);
    int32_t pi = (int32_t)f2u(p) + ((p < 0.0f) ? -of_i : of_i);
    return fabsf(p) < origin ? p + float_scale * n : u2f((uint32_t)pi);
}
AKR_HD f3 offset_ray_origin(f3 p, f3 n) { return mk3(offset_axis(p.x, n.x), offset_axis(p.y, n.y), offset_axis(p.z, n.z)); }

// ---- exact unsigned division by a launch-constant divisor -----------------------------------------
// q = floor(n / d) without the ~20-instruction hardware-less divide: one mul.hi estimate that is either
// exact or one too small (m = floor(2^(32+s) / d) clamped to 32 bits, s = floor(log2 d)), then one fix-up.
struct FastDiv {
    uint32_t d, m, s;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d ? d : 1u;
    uint32_t s = 0;
    while ((2u << s) <= f.d && s < 31u) ++s;
    f.s = s;
    unsigned long long m = ((1ull << (32u + s)) / f.d);
    f.m = m > 0xffffffffull ? 0xffffffffu : (uint32_t)m;
    return f;
}
AKR_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((unsigned long long)a * (unsigned long long)b) >> 32);
#endif
}
AKR_HD uint32_t fastdiv(uint32_t n, const FastDiv &f, uint32_t &rem) {
    uint32_t q = mulhi_u32(n, f.m) >> f.s;
    uint32_t r = n - q * f.d;
    if (r >= f.d) {
        q += 1u;
        r -= f.d;
    }
    rem = r;
    return q;
}

// ---- integer hashes (util/hash.rs:44-59, sampler/mod.rs:473-505) ----------------------------------
AKR_HD uint32_t rotl17(uint32_t h) { return (h << 17) | (h >> 15); }
AKR_HD uint32_t xxhash32_4(uint32_t px, uint32_t py, uint32_t pz, uint32_t pw) {
    const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    uint32_t h = pw + P5 + px * P3;
    h = P4 * rotl17(h);
    h = h + py * P3;
    h = P4 * rotl17(h);
    h = h + pz * P3;
    h = P4 * rotl17(h);
    h = P2 * (h ^ (h >> 15));
    h = P3 * (h ^ (h >> 13));
    return h ^ (h >> 16);
}
AKR_HD uint32_t permute_element(uint32_t i, uint32_t l, uint32_t w, uint32_t p) {
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
    // l == w + 1 (power-of-two sample counts): x % l == x & w
    return (l == w + 1u) ? ((i + p) & w) : ((i + p) % l);
}

// ---- sampling.rs -------------------------------------------------------------------------------
AKR_HD f2 uniform_sample_disk(f2 u) {  // :5-9
    float r = sqrtf(u.x);
    float phi = u.y * 2.0f * AKR_PI;
    float s, c;
#if defined(__CUDA_ARCH__)
    sincosf(phi, &s, &c);
#else
    s = sinf(phi);
    c = cosf(phi);
#endif
    return f2{r * c, r * s};
}
AKR_HD f3 cos_hemisphere_from_disk(f2 d) {
    float z = sqrtf(fmaxf(1.0f - d.x * d.x - d.y * d.y, 0.0f));
    return mk3(d.x, d.y, z);
}
AKR_HD f3 cos_sample_hemisphere(f2 u) { return cos_hemisphere_from_disk(uniform_sample_disk(u)); }  // :17-21
AKR_HD f2 uniform_sample_triangle(f2 u) {  // :32-44
    if (u.x < u.y) {
        float b0 = u.x / 2.0f;
        return f2{b0, u.y - b0};
    }
    float b1 = u.y / 2.0f;
    return f2{u.x - b1, b1};
}
// AliasTable::sample_and_remap (util/distribution.rs:82-88) over interleaved {j,t} + pdf arrays
struct AliasSample {
    uint32_t idx;
    float pdf, u;
};
AKR_HD AliasSample alias_sample_and_remap(const uint32_t *aj, const float *at, const float *apdf, uint32_t n, float u) {
    // uniform_discrete_choice_and_remap (sampling.rs:54-59)
    float fi = floorf(u * (float)n);
    int32_t i = (int32_t)fi;
    int32_t hi = (int32_t)n - 1;
    i = i < 0 ? 0 : (i > hi ? hi : i);
    float u1 = u * (float)n - (float)(uint32_t)i;
    // weighted_discrete_choice2_and_remap (sampling.rs:61-70)
    float t = at[i];
    bool first = u1 < t;
    uint32_t idx = first ? (uint32_t)i : aj[i];
    float u2 = first ? u1 / t : (u1 - t) / (1.0f - t);
    return AliasSample{idx, apdf[idx], u2};
}

}  // namespace akr

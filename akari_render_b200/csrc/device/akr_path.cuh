// akr_path.cuh — per-path bodies of the wavefront stages (one call = one path of one stage).
//
// The radiance loop of the reference megakernel (crates/akari_integrator/src/pt.rs:329-900 with
// shift_mapping = None; kernel body pt.rs:1075-1103) is cut at its two ray casts into stages:
//
//   raygen     sampler.start + camera.generate_ray                 pt.rs:1081-1098
//   intersect  scene.intersect (closest hit)                       pt.rs:363-380
//   shade      emitter MIS, depth cap, NEE sample + BSDF eval,     pt.rs:381-513 (minus the occlusion
//              BSDF sample, beta update, Russian roulette, next ray           test), 775-865
//   shadow     scene.occlude + add_radiance(direct)                pt.rs:504-513
//   accumulate indirect clamp + film.add_sample                    pt.rs:871-876, 1100; film.rs:196-229
//
// All paths of one stage launch are at the same depth (no path regeneration inside a wave), so the
// depth and the sampler dimension are launch constants, not per-path state.
#pragma once
#include "../../../include/akari_b200.h"
#include "akr_trace.cuh"

namespace akr {

// ---- Pmj02BnSampler as a pure function of (pixel, sample index, dimension) -------------------------
// sampler/mod.rs:551-623; state layout :513-520.  `dim` is passed explicitly.
AKR_HD float bluenoise(const SamplerTables &tab, uint32_t tex, uint32_t px, uint32_t py) {
    // reference texel: BLUE_NOISE_TEXTURES[tex % 48][px % 128][py % 128] (sampler/mod.rs:542-550);
    // the table is transposed at upload so that consecutive px are consecutive in memory.
    uint32_t t = tex % 48u;
    uint16_t v = AKR_RO(tab.bn[(t * 128u + (py & 127u)) * 128u + (px & 127u)]);
    return (float)v / 65535.0f;
}
AKR_HD float sampler_1d(const SamplerTables &tab, const RenderParams &rp, uint32_t px, uint32_t py, uint32_t sample_index, uint32_t dim) {
    uint32_t hash = xxhash32_4(px, py, dim, rp.seed);
    uint32_t index = permute_element(sample_index, rp.spp_total, rp.w_mask, hash);
    float delta = bluenoise(tab, dim, px, py);
    float x = (float)index + delta;
    // division by a power of two is exact, so the reciprocal multiply returns the same bits
    float q = rp.spp_pow2 ? x * rp.inv_spp : x / (float)rp.spp_total;
    return fminf(q, AKR_ONE_MINUS_EPSILON);
}
AKR_HD f2 sampler_2d(const SamplerTables &tab, const RenderParams &rp, uint32_t px, uint32_t py, uint32_t sample_index, uint32_t dim) {
    uint32_t index = sample_index;
    uint32_t pmj_instance = dim / 2u;
    if (pmj_instance >= 5u) {
        uint32_t hash = xxhash32_4(px, py, dim, rp.seed);
        index = permute_element(sample_index, rp.spp_total, rp.w_mask, hash);
    }
    uint32_t i = 65536u * (pmj_instance % 5u) + (index % 65536u);
    float ux = (float)AKR_RO(tab.pmj[i * 2u]) * 0x1p-32f;
    float uy = (float)AKR_RO(tab.pmj[i * 2u + 1u]) * 0x1p-32f;
    ux = ux + bluenoise(tab, dim, px, py);
    uy = uy + bluenoise(tab, dim + 1u, px, py);
    ux = ux - floorf(ux);
    uy = uy - floorf(uy);
    return f2{fminf(ux, AKR_ONE_MINUS_EPSILON), fminf(uy, AKR_ONE_MINUS_EPSILON)};
}
// Dimension of the first draw of bounce `depth` (depth counted after the increment at pt.rs:469):
// start() sets dim = 4, the filter takes 2, every earlier bounce took 3 + 3 and one more when it drew
// the roulette sample (depth_j > rr_depth, pt.rs:843-846).
AKR_HD uint32_t bounce_first_dim(uint32_t depth, uint32_t rr_depth) {
    uint32_t prev = depth - 1u;
    uint32_t rr = prev > rr_depth ? prev - rr_depth : 0u;
    return 6u + 6u * prev + rr;
}

// ---- wave bookkeeping -------------------------------------------------------------------------------
struct WaveInfo {
    uint32_t pix0;      // first pixel (linear index inside the tile) of this wave
    uint32_t n_pix;     // pixels in this wave
    uint32_t s0;        // first sample index
    uint32_t n_spp;     // samples per pixel in this wave; path_id = s_local * n_pix + p_local
    FastDiv n_pix_div;  // path_id -> (s_local, p_local)
};
inline WaveInfo make_wave(uint32_t pix0, uint32_t n_pix, uint32_t s0, uint32_t n_spp) {
    WaveInfo w;
    w.pix0 = pix0;
    w.n_pix = n_pix;
    w.s0 = s0;
    w.n_spp = n_spp;
    w.n_pix_div = make_fastdiv(n_pix);
    return w;
}
struct PathCoord {
    uint32_t px, py, sample_index, pixel_in_tile;
};
AKR_HD PathCoord path_coord(const RenderParams &rp, const WaveInfo &w, uint32_t path_id) {
    uint32_t p_local, col;
    uint32_t s_local = fastdiv(path_id, w.n_pix_div, p_local);
    uint32_t pix = w.pix0 + p_local;
    uint32_t row = fastdiv(pix, rp.width_div, col);
    if (rp.tile_shards > 1u) {  // interleaved tile: local row -> sensor row
        uint32_t within;
        const uint32_t blk = fastdiv(row, rp.tile_block_div, within);
        row = (blk * rp.tile_shards + rp.tile_shard) * rp.tile_block + within;
    }
    return PathCoord{col, rp.y0 + row, w.s0 + s_local, pix};
}

struct PathState {  // what survives from one bounce to the next (13 words in the SoA queue)
    f3 o, d;
    uint32_t ex;        // global id of the triangle the ray leaves (exclude0)
    f3 beta;
    float prev_bsdf_pdf;
    uint32_t path_id;
    uint32_t pxpy, sample_index;  // fused pipeline only: sensor pixel (x | y << 16) and sample index, so that the sampler
                                  // needs no path_id -> (pixel, sample) divisions per bounce (sensors up to 65535 x 65535)
};
struct ShadowItem {    // 13 words
    f3 o, d;
    float t_max;
    uint32_t ex0, ex1;
    f3 contrib;         // beta * direct, added to L when unoccluded
    uint32_t path_id;
};
struct f4 {             // 16-byte record: one vector load/store per access
    float x, y, z, w;
};
struct AccView {       // per-path radiance accumulators, indexed by path_id (not compacted); zeroed by raygen
    f4 *l;             // radiance (xyz)
    f4 *b;             // base_replay_throughput (xyz)
    uint32_t *poison;  // queued pipeline: channels of the radiance a miss poisoned with NaN (see miss_body); zeroed by raygen
};
AKR_HD f4 ld4(const f4 *p) {
#if defined(__CUDA_ARCH__)
    float4 v = *reinterpret_cast<const float4 *>(p);
    return f4{v.x, v.y, v.z, v.w};
#else
    return *p;
#endif
}
AKR_HD void st4(f4 *p, f4 v) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4 *>(p) = make_float4(v.x, v.y, v.z, v.w);
#else
    *p = v;
#endif
}

// ---- stage: raygen ------------------------------------------------------------------------------------
AKR_HD f2 filter_sample(const RenderParams &rp, f2 u) {  // film.rs:32-49
    if (rp.filter_type == 0u) return f2{(u.x - 0.5f) * rp.filter_radius, (u.y - 0.5f) * rp.filter_radius};
    float width = rp.filter_radius;
    float sigma = width / 3.0f;
    float r = sqrtf(-2.0f * logf(u.x));
    float theta = 2.0f * AKR_PI * u.y;
    float s, c;
#if defined(__CUDA_ARCH__)
    sincosf(theta, &s, &c);
#else
    s = sinf(theta);
    c = cosf(theta);
#endif
    return f2{clampf(r * c * sigma, -width, width), clampf(r * s * sigma, -width, width)};
}
AKR_HD PathState raygen_body(const SceneView &sc, const SamplerTables &tab, const RenderParams &rp, const WaveInfo &w, uint32_t path_id) {
    PathCoord pc = path_coord(rp, w, path_id);
    // shifted pixel (pt.rs:1084-1088); the sampler stays keyed by the unshifted pixel
    int32_t sx = (int32_t)pc.px + rp.pixel_offset_x, sy = (int32_t)pc.py + rp.pixel_offset_y;
    sx = sx < 0 ? 0 : (sx > (int32_t)rp.width - 1 ? (int32_t)rp.width - 1 : sx);
    sy = sy < 0 ? 0 : (sy > (int32_t)rp.height - 1 ? (int32_t)rp.height - 1 : sy);
    f2 u = sampler_2d(tab, rp, pc.px, pc.py, pc.sample_index, 4u);
    f2 off = filter_sample(rp, u);
    float fx = ((float)sx + 0.5f) + off.x, fy = ((float)sy + 0.5f) + off.y;
    const CameraRec &cam = sc.camera;
    // r2c.transform_point((fx, fy, 0)) with a scale+translate matrix, then normalize (camera/mod.rs:84-93)
    f3 dc = normalize(mk3(cam.r2c_s[0] * fx + cam.r2c_t[0], cam.r2c_s[1] * fy + cam.r2c_t[1], cam.r2c_s[2] * 0.0f + cam.r2c_t[2]));
    PathState ps;
    if (cam.c2w_identity) {  // AffineTransform shortcuts (geometry.rs:228-247)
        ps.o = splat3(0.0f);
        ps.d = dc;
    } else {
        ps.o = mk3(cam.c2w[9], cam.c2w[10], cam.c2w[11]);
        ps.d = mk3(cam.c2w[0] * dc.x + cam.c2w[3] * dc.y + cam.c2w[6] * dc.z, cam.c2w[1] * dc.x + cam.c2w[4] * dc.y + cam.c2w[7] * dc.z,
                   cam.c2w[2] * dc.x + cam.c2w[5] * dc.y + cam.c2w[8] * dc.z);
    }
    ps.ex = 0xffffffffu;
    ps.beta = splat3(1.0f);
    ps.prev_bsdf_pdf = 0.0f;
    ps.path_id = path_id;
    ps.pxpy = pc.px | (pc.py << 16);
    ps.sample_index = pc.sample_index;
    return ps;
}

// ---- surface interaction (mesh.rs:487-654) from the per-triangle record --------------------------------
struct Surface {
    f3 p, ng;
    Frame frame;
    float area;
};
struct CornerAttribs {           // optional per-corner arrays indexed by gid * 3 + k
    const float *normals;        // [n_tris * 3][3] or nullptr
    const float *tangents;       // [n_tris * 3][3] or nullptr
};
AKR_HD f3 ld3g(const float *p) { return mk3(AKR_RO(p[0]), AKR_RO(p[1]), AKR_RO(p[2])); }  // read-only scene data in global memory
// per-hit frame of a triangle with per-corner normals / tangents: ns from interpolated corner normals
// (mesh.rs:594-602,620-621), tangent from the tangent buffer when present (mesh.rs:558-569) else the per-triangle dp/du
// stored in `ft`
AKR_HD Frame surface_frame_interpolated(const SceneView &sc, const CornerAttribs &ca, uint32_t gid, float u, float v, f3 ng) {
    const TriShade &ts = sc.shade[gid];
    const InstanceRec &in = sc.instances[ts.inst];
    float w0 = 1.0f - u - v;
    f3 c0 = ld3g(in.m), c1 = ld3g(in.m + 3), c2 = ld3g(in.m + 6);
    f3 ns = ng;
    f3 i0 = ld3g(in.m_inv_t), i1 = ld3g(in.m_inv_t + 3), i2 = ld3g(in.m_inv_t + 6);
    if (AKR_RO(ts.flags) & TRI_HAS_NORMALS) {
        const float *n = ca.normals + (size_t)gid * 9u;
        f3 nl = w0 * ld3g(n) + u * ld3g(n + 3) + v * ld3g(n + 6);
        ns = normalize(i0 * nl.x + i1 * nl.y + i2 * nl.z);
    }
    f3 tt = ld3g(ts.ft);
    if (AKR_RO(ts.flags) & TRI_HAS_TANGENTS) {
        const float *t = ca.tangents + (size_t)gid * 9u;
        f3 t0 = ld3g(t), t1 = ld3g(t + 3), t2 = ld3g(t + 6);
        bool fin = is_finite(t0.x) && is_finite(t0.y) && is_finite(t0.z) && is_finite(t1.x) && is_finite(t1.y) && is_finite(t1.z) &&
                   is_finite(t2.x) && is_finite(t2.y) && is_finite(t2.z);
        if (fin) {
            f3 tl = normalize(w0 * t0 + u * t1 + v * t2);
            tt = c0 * tl.x + c1 * tl.y + c2 * tl.z;
        } else {
            tt = ld3g(ts.fs);  // fallback dp/du tangent is kept in `fs` for tangent-buffer meshes
        }
    }
    return (tt.x != 0.0f || tt.y != 0.0f || tt.z != 0.0f) ? frame_from_n_t(ns, tt) : frame_from_n(ns);
}
AKR_HD Surface surface_from_hit(const SceneView &sc, const CornerAttribs &ca, uint32_t gid, float u, float v) {
    const TriShade &ts = sc.shade[gid];
    const InstanceRec &in = sc.instances[ts.inst];
    float w0 = 1.0f - u - v;
    f3 v0 = ld3g(ts.v0), v1 = ld3g(ts.v1), v2 = ld3g(ts.v2);
    f3 pl = w0 * v0 + u * v1 + v * v2;
    f3 c0 = ld3g(in.m), c1 = ld3g(in.m + 3), c2 = ld3g(in.m + 6);
    Surface s;
    s.p = (c0 * pl.x + c1 * pl.y + c2 * pl.z) + ld3g(in.t);
    s.ng = ld3g(ts.ng);
    s.area = AKR_RO(ts.area);
    if (!(AKR_RO(ts.flags) & (TRI_HAS_NORMALS | TRI_HAS_TANGENTS))) s.frame = Frame{s.ng, ld3g(ts.ft), ld3g(ts.fs)};
    else s.frame = surface_frame_interpolated(sc, ca, gid, u, v, s.ng);
    return s;
}

AKR_HD float mis_weight(float a, float b) { return a / (a + b); }  // pt.rs:962-973, power 1

// ---- stage: shade ----------------------------------------------------------------------------------------
struct ShadeOut {
    bool has_shadow, has_next;
    ShadowItem shadow;
    PathState next;
};

AKR_HD void acc_add(const AccView &acc, uint32_t id, f3 c) {
    f4 l = ld4(acc.l + id);
    st4(acc.l + id, f4{l.x + c.x, l.y + c.y, l.z + c.z, 0.0f});
}
AKR_HD void acc_set_lb(const AccView &acc, uint32_t id, f3 l) {  // depth 0: radiance = base_replay_throughput = l
    st4(acc.l + id, f4{l.x, l.y, l.z, 0.0f});
    st4(acc.b + id, f4{l.x, l.y, l.z, 0.0f});
}
AKR_HD void acc_snap_b(const AccView &acc, uint32_t id) { st4(acc.b + id, ld4(acc.l + id)); }  // base_replay_throughput = radiance

// ---- a ray that left the scene: hit_envmap = (0, 0) (pt.rs:226-228,381-396) -----------------------------
// add_radiance(beta * 0): a NaN/inf throughput poisons the sample exactly as in the reference; a finite
// one adds nothing.  Depth-0 misses need no work at all (raygen zeroed the accumulators).
AKR_HD void miss_body(const RenderParams &rp, uint32_t depth, f3 beta, uint32_t id, const AccView &acc) {
    const bool dbg_on = rp.debug_depth < 0;
    if (depth != 0u && (dbg_on || depth == (uint32_t)rp.debug_depth)) {
        f3 c = beta * (splat3(0.0f) * 0.0f);
        // c is 0 or NaN per channel.  The trace stage that finds the miss runs concurrently with the shadow rays of the same
        // path (which read-modify-write the radiance), so the NaN is not added here: the poisoned channels are recorded and
        // accumulate_body applies them — x + NaN = NaN whenever it is added, so the result is the reference's.
        uint32_t mask = (c.x != 0.0f ? 1u : 0u) | (c.y != 0.0f ? 2u : 0u) | (c.z != 0.0f ? 4u : 0u);
        if (mask) acc.poison[id] = mask;  // (one miss per path: no other writer)
    }
}

// ---- handle_surface_light (pt.rs:230-258): c = beta * (Le * w) at a hit on a light triangle, zero otherwise ----
// `ray_o`, `ray_d`, `beta`, `prev_bsdf_pdf` describe the ray that produced the hit at path depth `depth`.
AKR_HD f3 emitter_contrib(const SceneView &sc, const RenderParams &rp, uint32_t depth, f3 ray_o, f3 ray_d, f3 beta, float prev_bsdf_pdf,
                          const TriShade &ts, const Material &mat, const Surface &si) {
    f3 direct = splat3(0.0f);
    float w = 0.0f;
    if ((ts.flags & TRI_IS_LIGHT) && (!rp.indirect_only || depth > 1u)) {
        f3 emission = ld3(mat.emission);  // AreaLight::le (light/area.rs:36-49)
        direct = dot(si.ng, ray_d) < 0.0f ? emission : splat3(0.0f);
        if (depth == 0u || !rp.use_nee) {
            w = 1.0f;
        } else {
            // LightAggregate::pdf_direct (light/mod.rs:134-147), AreaLight::pdf_direct (light/area.rs:109-130)
            const InstanceRec &in = sc.instances[ts.inst];
            float light_choice_pdf = sc.alias_pdf[in.light_id];
            f3 wi = si.p - ray_o;
            float dist2 = length_squared(wi);
            wi = wi / sqrtf(dist2);
            float pdf = ts.prim_pdf / si.area * dist2 / fmaxf(fabsf(dot(si.ng, wi)), 1e-6f);
            w = mis_weight(prev_bsdf_pdf, light_choice_pdf * pdf);
        }
    }
    return beta * (direct * w);
}
// add_radiance of that contribution to a register-resident radiance (same additions as acc_set_lb / acc_add)
AKR_HD f3 emitter_add(const RenderParams &rp, uint32_t depth, f3 L, f3 c) {
    const bool dbg_on = rp.debug_depth < 0;
    if (depth == 0u) return (dbg_on || rp.debug_depth == 0) ? splat3(0.0f) + c : splat3(0.0f);  // radiance starts at 0
    if ((dbg_on || depth == (uint32_t)rp.debug_depth) && (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f)) return L + c;
    return L;
}
// miss_body on a register-resident radiance
AKR_HD f3 miss_add(const RenderParams &rp, uint32_t depth, f3 beta, f3 L) {
    const bool dbg_on = rp.debug_depth < 0;
    if (depth != 0u && (dbg_on || depth == (uint32_t)rp.debug_depth)) {
        f3 c = beta * (splat3(0.0f) * 0.0f);
        if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f) return L + c;
    }
    return L;
}

// `depth` = path depth when the ray was cast (0 for camera rays).  `hit` is a real hit (misses end in
// miss_body).  CLS is the shade class of the hit material (CLS_ANY: decide per call).
// EMIT = false (fused pipeline): the emitter term of this hit was already added by the stage that traced the ray, so
// ps.o / ps.prev_bsdf_pdf are not read and `acc` is not touched.
template <int CLS, bool EMIT = true, class Acc>
AKR_HD ShadeOut shade_body(const SceneView &sc, const CornerAttribs &ca, const SamplerTables &tab, const RenderParams &rp, const WaveInfo &wave,
                           uint32_t depth, const PathState &ps, HitRec hit, Acc &acc) {
    ShadeOut out;
    out.has_shadow = false;
    out.has_next = false;
    const uint32_t id = ps.path_id;
    const bool dbg_on = rp.debug_depth < 0;
    const uint32_t d1 = depth + 1u;         // pt.rs:469
    // The sampler draws of this bounce depend only on (path id, depth), so they may be issued before the hit triangle
    // and its material are fetched (table loads in flight meanwhile).  Measured on B200: helps the register-roomier
    // conductor / general kernels (6.7 -> 6.5 ms), costs the 64-register Lambert kernel 3 % => decided per class.
    constexpr bool kDrawFirst = CLS != CLS_LAMBERT;  // (measured again with the fused kernels: 23.5 vs 22.4 ms per pass for Lambert)
    // fused pipeline: the record carries (pixel, sample index); queued pipeline: recover them from the path id
    const PathCoord pc = EMIT ? path_coord(rp, wave, id) : PathCoord{ps.pxpy & 0xffffu, ps.pxpy >> 16, ps.sample_index, 0u};
    const uint32_t dim0 = bounce_first_dim(d1, rp.rr_depth);
    float ul0 = 0.0f, ub0 = 0.0f;
    f2 ul12{0.0f, 0.0f}, ub12{0.0f, 0.0f};
    if (kDrawFirst) {
        ul0 = sampler_1d(tab, rp, pc.px, pc.py, pc.sample_index, dim0);
        ul12 = sampler_2d(tab, rp, pc.px, pc.py, pc.sample_index, dim0 + 1u);
        ub0 = sampler_1d(tab, rp, pc.px, pc.py, pc.sample_index, dim0 + 3u);
        ub12 = sampler_2d(tab, rp, pc.px, pc.py, pc.sample_index, dim0 + 4u);
    }
    const TriShade &ts = sc.shade[hit.gid];
    const Material *matp = &sc.materials[ts.mat];
    Material dyn;
    if ((CLS == CLS_GENERAL || CLS == CLS_ANY) && matp->dynamic) {  // texture-driven inputs: the reference's per-dispatch evaluation (eval.rs:364-380)
        dyn = *matp;
        svm_eval<false, false>(sc.svm, matp->shader_kind, matp->data_offset, hit_uv(sc, hit.gid, hit.u, hit.v, false), dyn, nullptr, matp->static_offset);
        matp = &dyn;
    }
    const Material &mat = *matp;
    Surface si = surface_from_hit(sc, ca, hit.gid, hit.u, hit.v);
    f3 wo = -ps.d;
    // handle_surface_light (pt.rs:230-258)
    if (EMIT) {
        f3 c = emitter_contrib(sc, rp, depth, ps.o, ps.d, ps.beta, ps.prev_bsdf_pdf, ts, mat, si);
        if (depth == 0u) {
            // radiance starts at 0; base_replay_throughput = radiance (pt.rs:415-417)
            f3 l = (dbg_on || rp.debug_depth == 0) ? splat3(0.0f) + c : splat3(0.0f);
            acc_set_lb(acc, id, l);
        } else if (dbg_on || depth == (uint32_t)rp.debug_depth) {
            if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f) acc_add(acc, id, c);
        }
    }
    if (depth >= rp.max_depth) return out;  // pt.rs:466-468
    // u_direct = next_3d, u_bsdf = next_3d — both are always drawn (pt.rs:471-481)
    if (!kDrawFirst) {
        ul0 = sampler_1d(tab, rp, pc.px, pc.py, pc.sample_index, dim0);
        ul12 = sampler_2d(tab, rp, pc.px, pc.py, pc.sample_index, dim0 + 1u);
        ub0 = sampler_1d(tab, rp, pc.px, pc.py, pc.sample_index, dim0 + 3u);
        ub12 = sampler_2d(tab, rp, pc.px, pc.py, pc.sample_index, dim0 + 4u);
    }

    // sample_light (pt.rs:170-209) -> LightAggregate::sample_direct -> AreaLight::sample_direct (light/area.rs:51-107)
    bool dl_valid = false;
    f3 dl_li = splat3(0.0f), dl_wi = splat3(0.0f);
    float dl_pdf = 0.0f;
    if (rp.use_nee && (!rp.indirect_only || d1 > 1u) && sc.n_lights > 0u) {
        AliasSample ls = alias_sample_and_remap(sc.alias_j, sc.alias_t, sc.alias_pdf, sc.n_lights, ul0);
        const LightRec &light = sc.lights[ls.idx];
        AliasSample prs = alias_sample_and_remap(sc.alias_j + light.alias_offset, sc.alias_t + light.alias_offset,
                                                 sc.alias_pdf + light.alias_offset, light.n_prims, ls.u);
        f2 bary = uniform_sample_triangle(ul12);
        uint32_t lgid = light.tri_offset + prs.idx;
        Surface lsi = surface_from_hit(sc, ca, lgid, bary.x, bary.y);
        f3 wi = lsi.p - si.p;
        if (length_squared(wi) != 0.0f) {
            float dist2 = length_squared(wi);
            wi = wi / sqrtf(dist2);
            f3 emission = ld3(sc.materials[sc.shade[lgid].mat].emission);
            f3 li = dot(wi, lsi.ng) < 0.0f ? emission : splat3(0.0f);
            float cos_theta_i = fabsf(dot(lsi.ng, wi));
            float pdf = prs.pdf / lsi.area * dist2 / cos_theta_i;
            if (is_finite(pdf)) {  // LightSample.valid (light/area.rs:104)
                f3 ro = offset_ray_origin(si.p, face_forward(si.ng, wi));
                float dist = sqrtf(dist2);
                dl_valid = true;
                dl_li = li;
                dl_wi = wi;
                dl_pdf = pdf * ls.pdf;  // light/mod.rs:130
                out.shadow.o = ro;
                out.shadow.d = wi;
                out.shadow.t_max = dist * (1.0f - 1e-3f);
                out.shadow.ex0 = hit.gid;  // pt.rs:189-190
                out.shadow.ex1 = lgid;     // light/area.rs:95
                out.shadow.path_id = id;
            }
        }
    }

    // sample_surface_and_shade_direct (pt.rs:297-323)
    Material fd;
    const Material *m = &mat;
    if ((CLS == CLS_LAMBERT || CLS == CLS_ANY) && rp.force_diffuse) {  // pt.rs:268-279 (every hit is binned into the Lambert class then)
        fd = mat;
        fd.type = MAT_LAMBERT;
        fd.wrap_inner = 0u;
        fd.diffuse[0] = fd.diffuse[1] = fd.diffuse[2] = 1.0f * AKR_FRAC_1_PI * 0.8f;
        m = &fd;
    }
    ClosureFrames cf = make_closure_frames(*m, si.frame, si.ng);
    f3 direct = splat3(0.0f);
    f3 bs_wi = splat3(0.0f), bs_color = splat3(0.0f);
    float bs_pdf = 0.0f;
    bool bs_valid = false;
    if (CLS == CLS_GENERAL || CLS == CLS_ANY) {  // (measured for the conductor class too: 6.82 vs 6.58 ms per pass, its two inlined copies stay)
        // The full tree is evaluated twice per bounce, for the light direction and for the sampled direction.  Both go
        // through ONE copy of the evaluation code (a two-trip loop that is not unrolled); the direction is sampled
        // first, it does not depend on the light evaluation.
        BsdfDir sd = closure_sample_wi<CLS>(*m, sc.albedo_table, cf, wo, ub0, ub12);  // SurfaceClosure::sample (mod.rs:795-815)
        BsdfEval e_dl = zero_eval(), e_bs = zero_eval();
        AKR_NO_UNROLL
        for (int k = 0; k < 2; ++k) {
            const bool want = k == 0 ? dl_valid : sd.valid;
            const f3 wi = k == 0 ? dl_wi : sd.wi;
            BsdfEval e = zero_eval();
            if (want) e = closure_eval<CLS>(*m, sc.albedo_table, cf, wo, wi);
            if (k == 0) e_dl = e;
            else e_bs = e;
        }
        if (dl_valid) {
            float w = mis_weight(dl_pdf, e_dl.pdf);
            direct = dl_li * e_dl.f * w / dl_pdf;
        }
        if (sd.valid) {
            bs_wi = sd.wi;
            bs_color = e_bs.f;
            bs_pdf = e_bs.pdf;
            bs_valid = e_bs.pdf > 0.0f;
        }
    } else {
        if (dl_valid) {
            BsdfEval e = closure_eval<CLS>(*m, sc.albedo_table, cf, wo, dl_wi);
            float w = mis_weight(dl_pdf, e.pdf);
            direct = dl_li * e.f * w / dl_pdf;
        }
        // SurfaceClosure::sample (mod.rs:795-815)
        BsdfDir sd = closure_sample_wi<CLS>(*m, sc.albedo_table, cf, wo, ub0, ub12);
        if (sd.valid) {
            BsdfEval e = closure_eval<CLS>(*m, sc.albedo_table, cf, wo, sd.wi);
            bs_wi = sd.wi;
            bs_color = e.f;
            bs_pdf = e.pdf;
            bs_valid = e.pdf > 0.0f;
        }
    }
    if (dl_valid) {  // the occlusion test and add_radiance(direct) happen in the shadow stage (pt.rs:504-513)
        out.has_shadow = true;
        bool add = dbg_on || d1 == (uint32_t)rp.debug_depth;
        out.shadow.contrib = add ? ps.beta * direct : splat3(0.0f);
    }
    f3 beta = ps.beta * (bs_color / bs_pdf);  // mul_beta(f / pdf) (pt.rs:783)
    if (bs_pdf <= 0.0f || !bs_valid || reduce_min(bs_color) < 0.0f) return out;  // pt.rs:832-842
    if (d1 > rp.rr_depth) {  // pt.rs:211-218,843-850
        float cont_prob = clampf(reduce_max(beta), 0.0f, 1.0f) * 0.95f;
        float ur = sampler_1d(tab, rp, pc.px, pc.py, pc.sample_index, dim0 + 6u);
        if (ur >= cont_prob) return out;
        beta = beta * (splat3(1.0f) / cont_prob);
    }
    out.has_next = true;
    out.next.o = offset_ray_origin(si.p, face_forward(si.ng, bs_wi));  // pt.rs:856
    out.next.d = bs_wi;
    out.next.ex = hit.gid;
    out.next.beta = beta;
    out.next.prev_bsdf_pdf = bs_pdf;
    out.next.path_id = id;
    return out;
}

// ---- stage: shadow (pt.rs:504-513) ----------------------------------------------------------------------------
// `depth1` = depth after the increment of the bounce that produced the item.
template <class Acc> AKR_HD void shadow_resolve(Acc &acc, const ShadowItem &it, bool occluded, uint32_t depth1) {
    uint32_t id = it.path_id;
    if (!occluded) {
        f3 c = it.contrib;
        if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f) acc_add(acc, id, c);
    }
    if (depth1 == 1u) acc_snap_b(acc, id);  // base_replay_throughput = radiance (pt.rs:510-512)
}

// ---- fused pipeline (flat scenes): one kernel per (depth, shade class) does shade + shadow ray + next ray ----------
// The path record carries the hit the PREVIOUS stage found and the radiance so far, so a bounce reads one
// contiguous 64-byte record and writes one: no hit queue, no (slot, path id) indirection, no per-bounce accumulator
// read-modify-write.  The emitter term of a hit (pt.rs:230-258) is evaluated by the stage that traced the ray — it has
// the ray origin and the BSDF pdf in registers — which is why neither is part of the record.  Additions to the
// radiance happen in the reference order: emitter(d), NEE(d), emitter(d + 1), ...
struct BounceRec {     // 15 words, stored as 4 x 16 B
    f3 d;              // direction of the ray that produced the hit (wo = -d)
    uint32_t gid;      // hit triangle
    float u, v;
    uint32_t path_id;
    uint32_t pxpy;     // sensor pixel x | y << 16
    f3 beta;
    uint32_t sample_index;
    f3 L;              // radiance so far, emitter term of this hit included
};
struct TraceHit {      // result of a closest-hit query; gid 0xffffffff = miss
    uint32_t gid, cls, light;
    float u, v;
};
struct BounceOut {
    bool cont;         // the path goes on: `next` enters depth + 1 in shade class `cls`
    bool shadow;       // a shadow ray was traced (statistics)
    bool traced;       // a continuation ray was traced (statistics)
    uint32_t cls;
    BounceRec next;
};
// A Tracer answers `closest(active, o, d, ex0)` and `trace2(shadow ray, &occluded, continuation ray)`; on the device both
// are warp-collective (every lane calls, the flags say whether it carries a ray), so the bodies below never return early.
// `h` = closest hit of the continuation ray `nx` (gid 0xffffffff = miss)
AKR_HD BounceOut continue_path(const SceneView &sc, const CornerAttribs &ca, const RenderParams &rp, uint32_t depth1, bool active, bool has_next,
                               const PathState &nx, f3 L, const TraceHit &h, const AccView &acc) {
    BounceOut r;
    r.shadow = false;
    r.traced = has_next;
    r.cont = has_next && h.gid != 0xffffffffu;
    r.cls = rp.force_diffuse ? (uint32_t)CLS_LAMBERT : h.cls;
    if (has_next && !r.cont) L = miss_add(rp, depth1, nx.beta, L);
    if (r.cont && h.light) {  // (a hit that is no light adds beta * 0 * 0: nothing, also at depth 0 where radiance = 0 + c)
        const TriShade &ts = sc.shade[h.gid];
        const Surface si = surface_from_hit(sc, ca, h.gid, h.u, h.v);
        L = emitter_add(rp, depth1, L, emitter_contrib(sc, rp, depth1, nx.o, nx.d, nx.beta, nx.prev_bsdf_pdf, ts, sc.materials[ts.mat], si));
    }
    if (depth1 >= rp.max_depth) r.cont = false;  // pt.rs:466-468: the emitter term is all that happens at the last hit
    if (active && !r.cont) st4(acc.l + nx.path_id, f4{L.x, L.y, L.z, 0.0f});  // the path ends here: its radiance is final
    r.next.d = nx.d;
    r.next.gid = h.gid;
    r.next.u = h.u;
    r.next.v = h.v;
    r.next.path_id = nx.path_id;
    r.next.pxpy = nx.pxpy;
    r.next.sample_index = nx.sample_index;
    r.next.beta = nx.beta;
    r.next.L = L;
    return r;
}
// raygen + camera ray + emitter term of the first hit
template <class Tracer>
AKR_HD BounceOut raygen_fused(const SceneView &sc, const CornerAttribs &ca, const SamplerTables &tab, const RenderParams &rp, const WaveInfo &wave, bool active,
                              uint32_t path_id, Tracer &tr, const AccView &acc) {
    PathState ps;
    ps.o = splat3(0.0f);
    ps.d = mk3(1.0f, 0.0f, 0.0f);
    ps.ex = 0xffffffffu;
    ps.beta = splat3(1.0f);
    ps.prev_bsdf_pdf = 0.0f;
    ps.path_id = path_id;
    ps.pxpy = 0u;
    ps.sample_index = 0u;
    if (active) ps = raygen_body(sc, tab, rp, wave, path_id);
    const TraceHit h = tr.closest(active, ps.o, ps.d, ps.ex);
    BounceOut r = continue_path(sc, ca, rp, 0u, active, active, ps, splat3(0.0f), h, acc);
    // base_replay_throughput of a path that ends at depth 0 (miss: 0; max_depth = 0: the emitter term, pt.rs:415-417)
    if (active && !r.cont) st4(acc.b + path_id, f4{r.next.L.x, r.next.L.y, r.next.L.z, 0.0f});
    return r;
}
// one bounce of one path whose hit is of shade class CLS
template <int CLS, class Tracer>
AKR_HD BounceOut bounce_fused(const SceneView &sc, const CornerAttribs &ca, const SamplerTables &tab, const RenderParams &rp, const WaveInfo &wave, uint32_t depth,
                              bool active, const BounceRec &in, Tracer &tr, const AccView &acc) {
    ShadeOut o = ShadeOut();
    f3 L = in.L;
    if (active) {
        PathState ps;
        ps.o = splat3(0.0f);
        ps.d = in.d;
        ps.ex = 0xffffffffu;
        ps.beta = in.beta;
        ps.prev_bsdf_pdf = 0.0f;
        ps.path_id = in.path_id;
        ps.pxpy = in.pxpy;
        ps.sample_index = in.sample_index;
        o = shade_body<CLS, false>(sc, ca, tab, rp, wave, depth, ps, HitRec{in.gid, in.u, in.v}, acc);
    }
    o.next.path_id = in.path_id;
    o.next.pxpy = in.pxpy;
    o.next.sample_index = in.sample_index;
    // the NEE shadow ray and the continuation ray are both known here: one trace call serves both (on the device one walk
    // over the staged primitive list, two independent dependency chains per trip)
    bool occluded;
    const TraceHit h = tr.trace2(o.has_shadow, o.shadow.o, o.shadow.d, o.shadow.t_max, o.shadow.ex0, o.shadow.ex1, occluded, o.has_next, o.next.o, o.next.d, o.next.ex);
    if (o.has_shadow && !occluded) {  // shadow_resolve (pt.rs:504-513)
        const f3 c = o.shadow.contrib;
        if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f) L = L + c;
    }
    if (active && depth == 0u) st4(acc.b + in.path_id, f4{L.x, L.y, L.z, 0.0f});  // base_replay_throughput = radiance (pt.rs:415-417,510-512)
    BounceOut r = continue_path(sc, ca, rp, depth + 1u, active, o.has_next, o.next, L, h, acc);
    r.shadow = o.has_shadow;
    return r;
}

// ---- the `aov` integrator (crates/akari_integrator/src/aov.rs:96-155): colour of one camera sample's first hit -----------
// `d` is the camera ray direction; the roughness AOV draws next_1d at the dimension after the filter pair (dim 6).
AKR_HD f3 aov_body(const SceneView &sc, const CornerAttribs &ca, const SamplerTables &tab, const RenderParams &rp, const WaveInfo &wave, uint32_t aov,
                   bool remap, uint32_t path_id, f3 d, HitRec hit) {
    const TriShade &ts = sc.shade[hit.gid];
    const Material *matp = &sc.materials[ts.mat];
    Material dyn;
    if (matp->dynamic) {
        dyn = *matp;
        svm_eval<false, false>(sc.svm, matp->shader_kind, matp->data_offset, hit_uv(sc, hit.gid, hit.u, hit.v, false), dyn, nullptr, matp->static_offset);
        matp = &dyn;
    }
    const Material &mat = *matp;
    const Surface si = surface_from_hit(sc, ca, hit.gid, hit.u, hit.v);
    const f3 wo = -d;
    f3 v;
    if (aov == AKR_AOV_SHADING_NORMAL) v = closure_ns(make_closure_frames(mat, si.frame, si.ng));
    else if (aov == AKR_AOV_GEOMETRY_NORMAL) v = si.ng;
    else if (aov == AKR_AOV_TANGENT) v = si.frame.t;
    else if (aov == AKR_AOV_BITANGENT) v = si.frame.s;
    else if (aov == AKR_AOV_ALBEDO) return material_albedo_plus_emission(mat);
    else {
        const PathCoord pc = path_coord(rp, wave, path_id);
        const float u = sampler_1d(tab, rp, pc.px, pc.py, pc.sample_index, 6u);
        return splat3(1.0f) * closure_roughness(mat, sc.albedo_table, make_closure_frames(mat, si.frame, si.ng), wo, u);
    }
    return remap ? v * 0.5f + splat3(0.5f) : v;
}

// ---- stage: accumulate (pt.rs:871-876 + film.rs:196-229) -------------------------------------------------------
// One call per pixel of the wave; samples are added in sample-index order like the reference's
// per-thread loop (pt.rs:1080-1101).  film = | rgb 3N | splat 3N | weight N | (film.rs:66-76).
AKR_HD void accumulate_body(const AccView &acc, const WaveInfo &w, uint32_t p_local, float *film, uint32_t n_film_pixels) {
    uint32_t i = w.pix0 + p_local;
    float r = film[i * 3u + 0u], g = film[i * 3u + 1u], b = film[i * 3u + 2u];
    float wt = film[6u * n_film_pixels + i];
    for (uint32_t s = 0; s < w.n_spp; ++s) {
        uint32_t id = s * w.n_pix + p_local;
        f4 l4 = ld4(acc.l + id), b4 = ld4(acc.b + id);
        f3 L = mk3(l4.x, l4.y, l4.z);
        if (acc.poison) {  // queued pipeline: a miss with a non-finite throughput poisons the sample (pt.rs:381-396 adds beta * 0)
            const uint32_t m = acc.poison[id];
            const float nan = u2f(0x7fc00000u);
            if (m) L = mk3((m & 1u) ? L.x + nan : L.x, (m & 2u) ? L.y + nan : L.y, (m & 4u) ? L.z + nan : L.z);
        }
        f3 B = mk3(b4.x, b4.y, b4.z);
        f3 ind = L - B;  // clamp_indirect = 1000, Color::clamp -> [0, max] (pt.rs:130,871-876; color.rs:352-361)
        ind = mk3(clampf(ind.x, 0.0f, 1000.0f), clampf(ind.y, 0.0f, 1000.0f), clampf(ind.z, 0.0f, 1000.0f));
        L = B + ind;
        if (has_nan(L)) L = splat3(0.0f);  // remove_nan (color.rs:343-351)
        L = L * 1.0f;                       // * ray weight
        r += L.x;
        g += L.y;
        b += L.z;
        wt += 1.0f;
    }
    film[i * 3u + 0u] = r;
    film[i * 3u + 1u] = g;
    film[i * 3u + 2u] = b;
    film[6u * n_film_pixels + i] = wt;
}

}  // namespace akr

// hostsim.cpp — TEST-ONLY host-side simulation of the wavefront kernels.
//
// Runs the exact per-path stage bodies of akari_render_b200/csrc/device/*.cuh (the code the CUDA
// kernels execute per thread) in plain loops on the CPU, with the same queue/compaction semantics,
// so kernel logic can be debugged and compared with the oracle in a container without a GPU.
// It is NOT part of the product: libakari_b200.so contains no CPU execution path, and nothing in
// akari_render_b200/ links or loads this file.
#include "../../akari_render_b200/csrc/device/akr_path.cuh"
#include "../../akari_render_b200/csrc/host/scene_build.h"
#include "../../oracle/chi2_tables.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace akr;

static thread_local int g_use_prims = 0;
static thread_local int g_fused = 0;
static thread_local int g_general_order = 0;
static thread_local uint32_t g_tile_block = 1, g_tile_shards = 1, g_tile_shard = 0;  // interleaved tile (AkrTile semantics)

namespace {
template <bool ANY_HIT> HitRec host_trace(const SceneView &sc, const TraceData &td, f3 o, f3 d, float t_max, uint32_t ex0, uint32_t ex1) {
    return g_use_prims == 1   ? trace_ray_prims<ANY_HIT>(sc, o, d, 0.0f, t_max, ex0, ex1)
           : g_use_prims == 2 ? trace_ray4<ANY_HIT>(sc, o, d, 0.0f, t_max, ex0, ex1)
           : g_use_prims == 3 ? trace_flat_ref<ANY_HIT>(sc, o, d, 0.0f, t_max, ex0, ex1)
                              : trace_ray<ANY_HIT>(sc, td, o, d, 0.0f, t_max, ex0, ex1);
}
struct HostTracer {  // the Tracer of akr_path.cuh's fused bodies, one ray at a time
    const SceneView &sc;
    const TraceData &td;
    bool occluded(bool active, f3 o, f3 d, float t_max, uint32_t ex0, uint32_t ex1) const {
        return active && host_trace<true>(sc, td, o, d, t_max, ex0, ex1).gid != 0xffffffffu;
    }
    TraceHit trace2(bool has_shadow, f3 so, f3 sd, float st_max, uint32_t sex0, uint32_t sex1, bool &occ, bool has_next, f3 o, f3 d, uint32_t ex0) const {
        occ = occluded(has_shadow, so, sd, st_max, sex0, sex1);
        return closest(has_next, o, d, ex0);
    }
    TraceHit closest(bool active, f3 o, f3 d, uint32_t ex0) const {
        TraceHit t{0xffffffffu, 0u, 0u, 0.0f, 0.0f};
        if (!active) return t;
        const HitRec h = host_trace<false>(sc, td, o, d, 1e20f, ex0, 0xffffffffu);
        if (h.gid == 0xffffffffu) return t;
        const TriShade &ts = sc.shade[h.gid];
        return TraceHit{h.gid, shade_class_of(sc.materials[ts.mat]), (ts.flags & TRI_IS_LIGHT) ? 1u : 0u, h.u, h.v};
    }
};
}  // namespace

// permutation of [0, n) that orders items by key(i) as hostsim_set_general_order asks
template <class KeyFn> static std::vector<uint32_t> general_order(size_t n, KeyFn key) {
    std::vector<uint32_t> idx(n);
    for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
    if (g_general_order == 1) std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
    if (g_general_order == 2) {
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
        std::reverse(idx.begin(), idx.end());
    }
    return idx;
}

extern "C" {

static thread_local std::string g_err;
const char *hostsim_last_error(void) { return g_err.c_str(); }

struct HostsimStats {
    uint64_t samples, segments, shadow_rays;
    uint32_t n_nodes, n_tris, n_materials, n_lights, bvh_depth;
    uint32_t material_types[8];
    uint32_t n_prims, n_pairs;
    uint32_t flat_blocks, flat_occluder_blocks;  // flat trace mode: 2-primitive blocks of the complete / occluder-only list
    uint32_t n_nodes4, bvh4_depth;
    uint32_t any_alpha, any_dynamic;  // some triangle needs the stochastic alpha test / some material is texture-driven
};

// 0: Moeller-Trumbore triangles over the binary BVH (bit-exact twin of the oracle); 1: the CUDA kernels' primitive
// intersector; 2: Moeller-Trumbore triangles over the 4-wide BVH (validates the collapsed tree); 3: the primitive
// intersector over the staged flat lists (complete list for closest hits, occluder-only list for shadow rays)
void hostsim_set_intersector(int use_prims) { g_use_prims = use_prims; }
// 0: queued pipeline (trace stage + shade stage + shadow queue, what BVH scenes run); 1: fused pipeline (one stage per
// depth and shade class does shade + shadow ray + next ray on records that carry hit and radiance: akr_path.cuh
// bounce_fused, what flat scenes run)
void hostsim_set_pipeline(int fused) { g_fused = fused; }
// Order in which the records of the general shade class are processed at every depth: 0 = arrival order; 1 = by material
// sort key, ascending (what k_sort_hist / k_sort_scatter produce on the device, up to the order within a key); 2 = by key
// descending with each key's records reversed.  The film must not depend on it (one path per record, own accumulators).
void hostsim_set_general_order(int mode) { g_general_order = mode; }
// rows [y0, y1) of hostsim_render are then the rows of shard `shard` of `n_shards` interleaved sets of `block_rows`-row blocks
void hostsim_set_tile_interleave(uint32_t block_rows, uint32_t n_shards, uint32_t shard) {
    g_tile_block = block_rows ? block_rows : 1u;
    g_tile_shards = n_shards ? n_shards : 1u;
    g_tile_shard = shard;
}



int hostsim_render(const AkrSceneDesc *desc, const AkrPtConfig *cfg, const AkrSamplerConfig *scfg, const AkrFilterConfig *filter,
                   const uint32_t *pmj, const uint16_t *bn, const float *albedo_table, uint32_t y0, uint32_t y1, uint32_t spp_begin,
                   uint32_t spp_end, uint32_t wave_pixels, float *film_7n, uint32_t *first_hits, HostsimStats *stats) {
    HostSceneBlob blob;
    std::string err;
    int rc = build_scene_blob(*desc, blob, err);
    if (rc != AKR_OK) {
        g_err = err;
        return rc;
    }
    if (g_use_prims == 3 && blob.flat_blocks.empty()) {
        g_err = "scene too large for the flat trace mode";
        return AKR_ERR_UNSUPPORTED;
    }
    if (blob.bvh_depth + 2u > (uint32_t)AKR_BVH_STACK) {  // the host walks use a fixed stack and would silently drop subtrees beyond it
        g_err = "BVH deeper than the host walk's stack (the CUDA kernels size theirs from the tree)";
        return AKR_ERR_UNSUPPORTED;
    }
    if (scfg->type != AKR_SAMPLER_PMJ02BN) {
        g_err = "pmj02bn only";
        return AKR_ERR_UNSUPPORTED;
    }
    std::vector<uint16_t> bnt(48u * 128u * 128u);
    transpose_bluenoise(bn, bnt.data());
    SceneView sc = host_scene_view(blob, albedo_table);
    CornerAttribs ca{blob.corner_normals.empty() ? nullptr : blob.corner_normals.data(),
                     blob.corner_tangents.empty() ? nullptr : blob.corner_tangents.data()};
    SamplerTables tab{pmj, bnt.data()};
    TraceData td{nullptr, sc.nodes, sc.tris, 0u};
    RenderParams rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.spp_total = cfg->spp;
    uint32_t w = cfg->spp - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    rp.w_mask = w;
    rp.seed = (uint32_t)scfg->seed;
    rp.max_depth = cfg->max_depth;
    rp.rr_depth = cfg->rr_depth;
    rp.use_nee = cfg->use_nee;
    rp.indirect_only = cfg->indirect_only;
    rp.force_diffuse = cfg->force_diffuse;
    rp.pixel_offset_x = cfg->pixel_offset[0];
    rp.pixel_offset_y = cfg->pixel_offset[1];
    rp.debug_depth = cfg->debug_depth;
    rp.filter_type = filter->type;
    rp.filter_radius = filter->radius;
    rp.width = desc->camera.width;
    rp.height = desc->camera.height;
    rp.y0 = y0;
    rp.tile_block = g_tile_block;
    rp.tile_shards = g_tile_shards;
    rp.tile_shard = g_tile_shard;
    finish_render_params(rp);
    const uint32_t rows = interleaved_tile_rows(y0, y1, g_tile_block, g_tile_shards, g_tile_shard);
    const uint32_t n_pixels = rp.width * rows;
    const uint32_t k = spp_end - spp_begin;
    if (wave_pixels == 0) wave_pixels = n_pixels;
    uint64_t segments = 0, shadows = 0;
    for (uint32_t pix0 = 0; pix0 < n_pixels; pix0 += wave_pixels) {
        WaveInfo wave = make_wave(pix0, std::min(wave_pixels, n_pixels - pix0), spp_begin, k);
        const uint32_t n_paths = wave.n_pix * wave.n_spp;
        if (g_fused) {
            // every path writes its radiance and its base_replay_throughput exactly once: start from a sentinel to prove it
            const float kUnset = -7777.0f;
            std::vector<f4> acc(2u * (size_t)n_paths, f4{kUnset, kUnset, kUnset, kUnset});
            AccView av{acc.data(), acc.data() + n_paths, nullptr};
            HostTracer tr{sc, td};
            std::vector<BounceRec> cur[CLS_COUNT], next[CLS_COUNT];
            for (uint32_t i = 0; i < n_paths; ++i) {
                BounceOut r = raygen_fused(sc, ca, tab, rp, wave, true, i, tr, av);
                ++segments;
                if (first_hits && spp_begin == wave.s0 && i < wave.n_pix) {
                    uint32_t pix = wave.pix0 + i, gid = r.next.gid;
                    first_hits[2 * pix + 0] = gid == 0xffffffffu ? 0xffffffffu : sc.shade[gid].inst;
                    first_hits[2 * pix + 1] = gid == 0xffffffffu ? 0xffffffffu : sc.shade[gid].prim;
                }
                if (r.cont) cur[r.cls].push_back(r.next);
            }
            for (uint32_t depth = 0; depth < rp.max_depth; ++depth) {
                for (uint32_t c = 0; c < (uint32_t)CLS_COUNT; ++c) {
                    const std::vector<BounceRec> &recs = cur[c];
                    auto key_of = [&](uint32_t i) { return (sc.shade[recs[i].gid].flags >> TRI_SORT_KEY_SHIFT) & TRI_SORT_KEY_MASK; };
                    const std::vector<uint32_t> order = c == CLS_GENERAL ? general_order(recs.size(), key_of) : general_order(0, key_of);
                    for (size_t oi = 0; oi < recs.size(); ++oi) {
                        const BounceRec &rec = recs[c == CLS_GENERAL ? order[oi] : oi];
                        BounceOut r = c == CLS_LAMBERT     ? bounce_fused<CLS_LAMBERT>(sc, ca, tab, rp, wave, depth, true, rec, tr, av)
                                      : c == CLS_CONDUCTOR ? bounce_fused<CLS_CONDUCTOR>(sc, ca, tab, rp, wave, depth, true, rec, tr, av)
                                                           : bounce_fused<CLS_GENERAL>(sc, ca, tab, rp, wave, depth, true, rec, tr, av);
                        shadows += r.shadow ? 1u : 0u;
                        segments += r.traced ? 1u : 0u;
                        if (r.cont) next[r.cls].push_back(r.next);
                    }
                    cur[c].clear();
                }
                for (uint32_t c = 0; c < (uint32_t)CLS_COUNT; ++c) cur[c].swap(next[c]);
            }
            for (const f4 &v : acc)
                if (v.x == kUnset && v.w == kUnset) {
                    g_err = "fused pipeline: a path never wrote its accumulators";
                    return AKR_ERR_STATE;
                }
            for (uint32_t p = 0; p < wave.n_pix; ++p) accumulate_body(av, wave, p, film_7n, n_pixels);
            continue;
        }
        std::vector<f4> acc(2u * (size_t)n_paths, f4{0.0f, 0.0f, 0.0f, 0.0f});  // raygen zeroes the accumulators
        std::vector<uint32_t> poison(n_paths, 0u);
        AccView av{acc.data(), acc.data() + n_paths, poison.data()};
        std::vector<PathState> cur(n_paths), next;
        for (uint32_t i = 0; i < n_paths; ++i) cur[i] = raygen_body(sc, tab, rp, wave, i);
        for (uint32_t depth = 0; depth <= rp.max_depth && !cur.empty(); ++depth) {
            std::vector<HitRec> hits(cur.size());
            for (size_t i = 0; i < cur.size(); ++i)
                hits[i] = g_use_prims == 1   ? trace_ray_prims<false>(sc, cur[i].o, cur[i].d, 0.0f, 1e20f, cur[i].ex, 0xffffffffu)
                          : g_use_prims == 2 ? trace_ray4<false>(sc, cur[i].o, cur[i].d, 0.0f, 1e20f, cur[i].ex, 0xffffffffu)
                          : g_use_prims == 3 ? trace_flat_ref<false>(sc, cur[i].o, cur[i].d, 0.0f, 1e20f, cur[i].ex, 0xffffffffu)
                                             : trace_ray<false>(sc, td, cur[i].o, cur[i].d, 0.0f, 1e20f, cur[i].ex, 0xffffffffu);
            segments += cur.size();
            if (depth == 0 && first_hits && spp_begin == wave.s0)
                for (size_t i = 0; i < cur.size(); ++i) {
                    uint32_t id = cur[i].path_id;
                    if (id >= wave.n_pix) continue;  // sample s0 only
                    uint32_t pix = wave.pix0 + id;
                    uint32_t gid = hits[i].gid;
                    first_hits[2 * pix + 0] = gid == 0xffffffffu ? 0xffffffffu : sc.shade[gid].inst;
                    first_hits[2 * pix + 1] = gid == 0xffffffffu ? 0xffffffffu : sc.shade[gid].prim;
                }
            next.clear();
            std::vector<ShadowItem> shq;
            // the trace stage bins hits by shade class; the general class list is then shaded in the order of its sort keys
            std::vector<uint32_t> general_items, visit;
            auto class_of = [&](size_t i) { return rp.force_diffuse ? (uint32_t)CLS_LAMBERT : shade_class_of(sc.materials[sc.shade[hits[i].gid].mat]); };
            for (size_t i = 0; i < cur.size(); ++i) {
                if (hits[i].gid == 0xffffffffu) {  // the trace stage ends missed paths itself
                    miss_body(rp, depth, cur[i].beta, cur[i].path_id, av);
                    continue;
                }
                if (g_general_order != 0 && class_of(i) == CLS_GENERAL) general_items.push_back((uint32_t)i);
                else visit.push_back((uint32_t)i);
            }
            for (uint32_t k : general_order(general_items.size(), [&](uint32_t j) { return (sc.shade[hits[general_items[j]].gid].flags >> TRI_SORT_KEY_SHIFT) & TRI_SORT_KEY_MASK; }))
                visit.push_back(general_items[k]);
            for (uint32_t i : visit) {
                // each class runs its own specialisation
                uint32_t cls = class_of(i);
                ShadeOut o = cls == CLS_LAMBERT     ? shade_body<CLS_LAMBERT>(sc, ca, tab, rp, wave, depth, cur[i], hits[i], av)
                             : cls == CLS_CONDUCTOR ? shade_body<CLS_CONDUCTOR>(sc, ca, tab, rp, wave, depth, cur[i], hits[i], av)
                                                    : shade_body<CLS_GENERAL>(sc, ca, tab, rp, wave, depth, cur[i], hits[i], av);
                if (o.has_shadow) shq.push_back(o.shadow);
                if (o.has_next) next.push_back(o.next);
            }
            shadows += shq.size();
            for (const ShadowItem &it : shq) {
                HitRec h = g_use_prims == 1   ? trace_ray_prims<true>(sc, it.o, it.d, 0.0f, it.t_max, it.ex0, it.ex1)
                           : g_use_prims == 2 ? trace_ray4<true>(sc, it.o, it.d, 0.0f, it.t_max, it.ex0, it.ex1)
                           : g_use_prims == 3 ? trace_flat_ref<true>(sc, it.o, it.d, 0.0f, it.t_max, it.ex0, it.ex1)
                                              : trace_ray<true>(sc, td, it.o, it.d, 0.0f, it.t_max, it.ex0, it.ex1);
                shadow_resolve(av, it, h.gid != 0xffffffffu, depth + 1u);
            }
            cur.swap(next);
        }
        for (uint32_t p = 0; p < wave.n_pix; ++p) accumulate_body(av, wave, p, film_7n, n_pixels);
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->samples = (uint64_t)n_pixels * k;
        stats->segments = segments;
        stats->shadow_rays = shadows;
        stats->n_nodes = (uint32_t)blob.nodes.size();
        stats->n_tris = (uint32_t)blob.shade.size();
        stats->n_prims = (uint32_t)blob.prims.size();
        for (const PrimRec &pr : blob.prims) stats->n_pairs += pr.gid_b != 0xffffffffu ? 1u : 0u;
        stats->n_nodes4 = (uint32_t)blob.nodes4.size();
        stats->bvh4_depth = blob.bvh4_depth;
        stats->flat_blocks = blob.n_pair_blocks + blob.n_single_blocks;
        stats->flat_occluder_blocks = blob.n_occ_pair_blocks + blob.n_occ_single_blocks;
        stats->n_materials = (uint32_t)blob.materials.size();
        stats->n_lights = (uint32_t)blob.lights.size();
        stats->bvh_depth = blob.bvh_depth;
        for (const Material &m : blob.materials) stats->material_types[m.type & 7u]++;
        stats->any_alpha = blob.any_alpha;
        stats->any_dynamic = blob.any_dynamic;
    }
    return AKR_OK;
}

// The `aov` method through the kernels' bodies: raygen_body, the chosen intersector, aov_body, accumulate_body.
int hostsim_render_aov(const AkrSceneDesc *desc, const AkrAovConfig *cfg, const AkrSamplerConfig *scfg, const AkrFilterConfig *filter, const uint32_t *pmj,
                       const uint16_t *bn, const float *albedo_table, uint32_t y0, uint32_t y1, float *film_7n) {
    HostSceneBlob blob;
    std::string err;
    int rc = build_scene_blob(*desc, blob, err);
    if (rc != AKR_OK) {
        g_err = err;
        return rc;
    }
    std::vector<uint16_t> bnt(48u * 128u * 128u);
    transpose_bluenoise(bn, bnt.data());
    SceneView sc = host_scene_view(blob, albedo_table);
    CornerAttribs ca{blob.corner_normals.empty() ? nullptr : blob.corner_normals.data(), blob.corner_tangents.empty() ? nullptr : blob.corner_tangents.data()};
    SamplerTables tab{pmj, bnt.data()};
    TraceData td{nullptr, sc.nodes, sc.tris, 0u};
    RenderParams rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.spp_total = cfg->spp;
    uint32_t w = cfg->spp - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    rp.w_mask = w;
    rp.seed = (uint32_t)scfg->seed;
    rp.debug_depth = -1;
    rp.filter_type = filter->type;
    rp.filter_radius = filter->radius;
    rp.width = desc->camera.width;
    rp.height = desc->camera.height;
    rp.y0 = y0;
    finish_render_params(rp);
    const uint32_t n_pixels = rp.width * (y1 - y0);
    WaveInfo wave = make_wave(0, n_pixels, 0, cfg->spp);
    const uint32_t n_paths = wave.n_pix * wave.n_spp;
    std::vector<f4> acc(2u * (size_t)n_paths, f4{0.0f, 0.0f, 0.0f, 0.0f});
    AccView av{acc.data(), acc.data() + n_paths, nullptr};
    for (uint32_t i = 0; i < n_paths; ++i) {
        PathState ps = raygen_body(sc, tab, rp, wave, i);
        HitRec h = host_trace<false>(sc, td, ps.o, ps.d, 1e20f, 0xffffffffu, 0xffffffffu);
        if (h.gid == 0xffffffffu) continue;
        f3 c = aov_body(sc, ca, tab, rp, wave, cfg->aov, cfg->remap != 0u, i, ps.d, h);
        av.l[i] = f4{c.x, c.y, c.z, 0.0f};
        av.b[i] = f4{c.x, c.y, c.z, 0.0f};
    }
    for (uint32_t p = 0; p < wave.n_pix; ++p) accumulate_body(av, wave, p, film_7n, n_pixels);
    return AKR_OK;
}

void hostsim_make_albedo_table(float *table, uint32_t n) { make_albedo_table(table, n); }

// Chi-square tables (oracle/chi2_tables.h, methodology of akari_test.rs:31-112) of the DEVICE closures: the constant-folded
// Material of instance `inst`'s first triangle, evaluated through closure_sample_wi / closure_eval (akr_bsdf.cuh) in an
// identity shading frame, with the template instantiation the kernels use for that material's shade class.
int hostsim_bsdf_chi2_tables(const AkrSceneDesc *desc, uint32_t inst, const float *albedo_table, const float *wo3, uint64_t n_samples, uint64_t seed,
                             uint32_t theta_res, uint32_t phi_res, uint32_t *hist_out, double *expected_out, uint32_t *material_type_out) {
    HostSceneBlob blob;
    std::string err;
    int rc = build_scene_blob(*desc, blob, err);
    if (rc != AKR_OK) {
        g_err = err;
        return rc;
    }
    if (inst >= blob.instances.size()) {
        g_err = "no such instance";
        return AKR_ERR_INVALID_ARGUMENT;
    }
    const Material m = blob.materials[blob.shade[blob.instances[inst].tri_offset].mat];
    if (material_type_out) *material_type_out = m.type;
    const ClosureFrames cf = make_closure_frames(m, frame_identity(), mk3(0.0f, 0.0f, 1.0f));
    const f3 wo = mk3(wo3[0], wo3[1], wo3[2]);
    const uint32_t cls = shade_class_of(m);
    auto sample = [&](float us, float u0, float u1, chi2::Dir &wi) {
        BsdfDir s = cls == CLS_LAMBERT     ? closure_sample_wi<CLS_LAMBERT>(m, albedo_table, cf, wo, us, f2{u0, u1})
                    : cls == CLS_CONDUCTOR ? closure_sample_wi<CLS_CONDUCTOR>(m, albedo_table, cf, wo, us, f2{u0, u1})
                                           : closure_sample_wi<CLS_GENERAL>(m, albedo_table, cf, wo, us, f2{u0, u1});
        wi = chi2::Dir{s.wi.x, s.wi.y, s.wi.z};
        return s.valid;
    };
    auto pdf = [&](chi2::Dir w) {
        const f3 wi = mk3(w.x, w.y, w.z);
        return (cls == CLS_LAMBERT     ? closure_eval<CLS_LAMBERT>(m, albedo_table, cf, wo, wi)
                : cls == CLS_CONDUCTOR ? closure_eval<CLS_CONDUCTOR>(m, albedo_table, cf, wo, wi)
                                       : closure_eval<CLS_GENERAL>(m, albedo_table, cf, wo, wi))
            .pdf;
    };
    chi2::histogram(sample, n_samples, seed, theta_res, phi_res, hist_out, 0);
    chi2::expected(pdf, n_samples, theta_res, phi_res, expected_out, 0);
    return AKR_OK;
}

// exact-division helper of the kernels (akr_math.cuh: FastDiv): q[i], r[i] = n[i] / d, n[i] % d
void hostsim_fastdiv(const uint32_t *n, uint32_t count, uint32_t d, uint32_t *q, uint32_t *r) {
    FastDiv f = make_fastdiv(d);
    for (uint32_t i = 0; i < count; ++i) q[i] = fastdiv(n[i], f, r[i]);
}

// Slab test agreement: box_test ((plane - o) * inv_d, what trace_ray and the oracle's tree walk use) against box_test_fma
// (plane * inv_d + ood with the capped inverse direction, what the 4-wide walk evaluates per child) over every node of
// the scene's BVH and `n_rays` seeded rays, a quarter of them with a direction component of exactly 0, a quarter with a
// denormal-small one, a quarter starting exactly on a vertex coordinate (rays start on surfaces).  counts = {accepted by both, box_test only, fma only}.
int hostsim_box_test_agreement(const AkrSceneDesc *desc, uint32_t n_rays, uint32_t seed, uint64_t counts[3]) {
    HostSceneBlob blob;
    std::string err;
    int rc = build_scene_blob(*desc, blob, err);
    if (rc != AKR_OK) {
        g_err = err;
        return rc;
    }
    counts[0] = counts[1] = counts[2] = 0;
    uint64_t state = 0x9e3779b97f4a7c15ull ^ seed;
    auto rnd = [&]() {  // splitmix64 -> [0, 1)
        state += 0x9e3779b97f4a7c15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        return (float)(z >> 40) * (1.0f / 16777216.0f);
    };
    const BvhNode &root = blob.nodes[0];
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(root.lo0[a], root.lo1[a]);
        hi[a] = fmaxf(root.hi0[a], root.hi1[a]);
    }
    for (uint32_t r = 0; r < n_rays; ++r) {
        float o[3], d[3];
        for (int a = 0; a < 3; ++a) {
            o[a] = lo[a] + (hi[a] - lo[a]) * rnd();
            d[a] = 2.0f * rnd() - 1.0f;
        }
        const uint32_t mode = r & 3u, ax = (uint32_t)(rnd() * 3.0f) % 3u;
        const BvhNode &pick = blob.nodes[(size_t)(rnd() * (float)blob.nodes.size()) % blob.nodes.size()];
        // (origins exactly on a PADDED box plane are left out on purpose: there the reference form yields t = 0 and the fma
        // form +-1 ulp of o * inv_d, a touch-or-miss at a point the padding keeps every primitive away from; rays start on
        // surfaces, i.e. on vertex coordinates)
        if (mode == 0u) o[ax] = blob.shade[(size_t)(rnd() * (float)blob.shade.size()) % blob.shade.size()].v0[ax];
        (void)pick;
        const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int a = 0; a < 3; ++a) d[a] /= len;
        if (mode == 1u) d[ax] = rnd() < 0.5f ? 0.0f : -0.0f;
        if (mode == 2u) d[ax] = 1e-41f * (rnd() - 0.5f);
        const f3 of = mk3(o[0], o[1], o[2]), df = mk3(d[0], d[1], d[2]);
        const f3 inv = mk3(1.0f / df.x, 1.0f / df.y, 1.0f / df.z), invc = capped_inv_dir(df);
        const f3 ood = mk3(-of.x * invc.x, -of.y * invc.y, -of.z * invc.z);
        const float t_max = (r & 4u) ? 1e20f : 4.0f * rnd();
        for (const BvhNode &n : blob.nodes)
            for (int c = 0; c < 2; ++c) {
                float ta, tb;
                const bool a = box_test(c ? n.lo1 : n.lo0, c ? n.hi1 : n.hi0, of, inv, 0.0f, t_max, ta);
                const bool b = box_test_fma(c ? n.lo1 : n.lo0, c ? n.hi1 : n.hi0, invc, ood, 0.0f, t_max, tb);
                counts[0] += a && b;
                counts[1] += a && !b;
                counts[2] += b && !a;
            }
    }
    return AKR_OK;
}

// Per material (up to `cap`): the sort key the general shade class orders its CTA tiles by (TriShade.flags bits 8-12, read
// back from a triangle that uses the material), closure type, lobe set, `dynamic`.  Returns the number of materials.
int hostsim_material_keys(const AkrSceneDesc *desc, uint32_t cap, uint32_t *keys, uint32_t *types, uint32_t *lobes, uint32_t *dynamic) {
    HostSceneBlob blob;
    std::string err;
    int rc = build_scene_blob(*desc, blob, err);
    if (rc != AKR_OK) {
        g_err = err;
        return -1;
    }
    const uint32_t n = (uint32_t)std::min<size_t>(cap, blob.materials.size());
    for (uint32_t m = 0; m < n; ++m) {
        keys[m] = 0xffffffffu;
        types[m] = blob.materials[m].type;
        lobes[m] = blob.materials[m].lobes;
        dynamic[m] = blob.materials[m].dynamic;
    }
    for (const TriShade &ts : blob.shade)
        if (ts.mat < n) keys[ts.mat] = (ts.flags >> TRI_SORT_KEY_SHIFT) & TRI_SORT_KEY_MASK;
    return (int)blob.materials.size();
}

}  // extern "C"

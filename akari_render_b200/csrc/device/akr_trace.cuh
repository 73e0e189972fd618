// akr_trace.cuh — BVH2 traversal + ray/triangle intersection.
//
// Replaces the third-party layer the reference calls through luisa `rtx::Accel`
// (crates/akari_render/src/scene.rs:88-110 `_trace_closest_rq`, :155-185 `occlude`; OptiX on
// `-d cuda`, Embree on `-d cpu`).  Semantics reproduced here:
//   * candidates equal to ray.exclude0 / exclude1 are skipped (scene.rs:99-101,167-170);
//   * stochastic alpha test on candidates of materials with alpha < 1 (scene.rs:49-86);
//   * closest hit = minimum t in (t_min, t_max); exact ties resolve to the lower global triangle id
//     so the result does not depend on traversal order (and equals a brute-force scan in id order).
#pragma once
#include "akr_scene.cuh"

namespace akr {

struct HitRec {
    uint32_t gid;  // 0xffffffff = miss
    float u, v;
};

// Moeller-Trumbore on a precomputed (v0, e1, e2) triangle; bary (u, v) weights v1, v2.
AKR_HD bool tri_test(const TriGeom &tr, f3 o, f3 d, float t_min, float t_max, float &t_out, float &u_out, float &v_out) {
    f3 e1 = mk3(tr.e1[0], tr.e1[1], tr.e1[2]);
    f3 e2 = mk3(tr.e2[0], tr.e2[1], tr.e2[2]);
    f3 pvec = cross(d, e2);
    float det = dot(e1, pvec);
    if (det == 0.0f) return false;
    float inv_det = 1.0f / det;
    f3 tvec = o - mk3(tr.v0[0], tr.v0[1], tr.v0[2]);
    float u = dot(tvec, pvec) * inv_det;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    f3 qvec = cross(tvec, e1);
    float v = dot(d, qvec) * inv_det;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    float t = dot(e2, qvec) * inv_det;
    if (!(t > t_min && t < t_max)) return false;
    t_out = t;
    u_out = u;
    v_out = v;
    return true;
}

// Conservative slab test (boxes are padded at build time); NaNs from 0 * inf drop out of fminf/fmaxf.
AKR_HD bool box_test(const float *lo, const float *hi, f3 o, f3 inv_d, float t_min, float t_max, float &t_near) {
    float tx0 = (lo[0] - o.x) * inv_d.x, tx1 = (hi[0] - o.x) * inv_d.x;
    float ty0 = (lo[1] - o.y) * inv_d.y, ty1 = (hi[1] - o.y) * inv_d.y;
    float tz0 = (lo[2] - o.z) * inv_d.z, tz1 = (hi[2] - o.z) * inv_d.z;
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), t_min));
    float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), t_max));
    t_near = tn;
    return tn <= tf;
}

// The same slab test in fused-multiply-add form (what the 4-wide walk below uses per child, box4_test): t = plane * inv_d + ood,
// ood = -o * inv_d, with inv_d from capped_inv_dir.  With inv_d = inf (a direction component of exactly 0) the two infinities would cancel to
// NaN on one plane of a slab, the min / max drop it and the slab collapses to (-inf, -inf): a false reject.  Capped at 2^96
// every product stays finite and the slab keeps its meaning (origin inside: (-huge, +huge); outside: both ends on one
// side).  The rounding difference to box_test is |o| * 2^-24 in space, far inside the build-time padding.
constexpr float kInvDirCap = 7.9228163e28f;
AKR_HD f3 capped_inv_dir(f3 d) {
    return mk3(fminf(fmaxf(1.0f / d.x, -kInvDirCap), kInvDirCap), fminf(fmaxf(1.0f / d.y, -kInvDirCap), kInvDirCap),
               fminf(fmaxf(1.0f / d.z, -kInvDirCap), kInvDirCap));
}
AKR_HD bool box_test_fma(const float *lo, const float *hi, f3 inv_d, f3 ood, float t_min, float t_max, float &t_near) {
    float tx0 = fmaf(lo[0], inv_d.x, ood.x), tx1 = fmaf(hi[0], inv_d.x, ood.x);
    float ty0 = fmaf(lo[1], inv_d.y, ood.y), ty1 = fmaf(hi[1], inv_d.y, ood.y);
    float tz0 = fmaf(lo[2], inv_d.z, ood.z), tz1 = fmaf(hi[2], inv_d.z, ood.z);
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), t_min));
    float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), t_max));
    t_near = tn;
    return tn <= tf;
}

// Texture coordinates of a hit (mesh.rs:534-546): interpolated per-corner uvs, or the default corners (0,0), (1,0), (1,0.1).
// `for_alpha_test`: the default third corner is (0, 0.1) in surface_interaction_for_alpha_test (mesh.rs:456-467) — preserved.
AKR_HD f2 hit_uv(const SceneView &sc, uint32_t gid, float u, float v, bool for_alpha_test) {
    float a[2] = {0.0f, 0.0f}, b[2] = {1.0f, 0.0f}, c[2] = {for_alpha_test ? 0.0f : 1.0f, 0.1f};
    if ((sc.shade[gid].flags & TRI_HAS_UVS) && sc.corner_uvs) {
        const float *p = sc.corner_uvs + (size_t)gid * 6u;
        a[0] = p[0]; a[1] = p[1]; b[0] = p[2]; b[1] = p[3]; c[0] = p[4]; c[1] = p[5];
    }
    const float w = 1.0f - u - v;
    return f2{w * a[0] + u * b[0] + v * c[0], w * a[1] + u * b[1] + v * c[1]};
}
// scene.rs:49-86: pass if alpha >= 1 or alpha > hash / 2^32; alpha = Surface::alpha() of the material in SvmEvalMode::Alpha
// texture-driven alpha: ONE out-of-line copy of the shader interpreter per kernel (inlined at every candidate-hit site it
// made the alpha variants of the traversal kernels 26 K instructions long, 25x the opaque ones)
AKR_HD_NOINLINE float alpha_of_dynamic(const SceneView &sc, const Material &mat, uint32_t gid, float u, float v) {
    Material tmp;
    svm_eval<true, false>(sc.svm, mat.shader_kind, mat.data_offset, hit_uv(sc, gid, u, v, true), tmp, nullptr, mat.static_offset);
    return tmp.alpha;
}
AKR_HD bool alpha_test(const SceneView &sc, uint32_t gid, float u, float v) {
    const TriShade &ts = sc.shade[gid];
    if (!(ts.flags & TRI_ALPHA)) return true;
    const Material &mat = sc.materials[ts.mat];
    float alpha = mat.alpha;
    if (mat.alpha_dynamic) alpha = alpha_of_dynamic(sc, mat, gid, u, v);
    uint32_t h = xxhash32_4(ts.inst, ts.prim, f2u(u), f2u(v));
    float hf = (float)h * (float)(1.0 / 4294967295.0);
    return (alpha >= 1.0f) || (alpha > hf);
}

#define AKR_BVH_STACK 48

// ANY_HIT = false: closest hit (Scene::intersect).  ANY_HIT = true: first accepted hit (Scene::occlude).
// Where traversal data lives.  On the device the first `n_fast_nodes` nodes (breadth-first order = the
// top of the tree) and, when they fit, all triangles are staged in shared memory by a TMA bulk copy;
// the rest is read from global memory (L2-resident).  The host simulation passes n_fast_nodes = 0.
struct TraceData {
    const BvhNode *fast_nodes;  // shared memory copy of nodes[0 .. n_fast_nodes)
    const BvhNode *nodes;       // global
    const TriGeom *tris;        // shared or global
    uint32_t n_fast_nodes;
};

template <bool ANY_HIT>
AKR_HD HitRec trace_ray(const SceneView &sc, const TraceData &td, f3 o, f3 d, float t_min, float t_max, uint32_t ex0, uint32_t ex1) {
    const TriGeom *tris = td.tris;
    HitRec best{0xffffffffu, 0.0f, 0.0f};
    float best_t = t_max;
    f3 inv_d = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int32_t stack[AKR_BVH_STACK];
    int sp = 0;
    int32_t node = 0;
    while (true) {
        if (node >= 0) {
            const BvhNode n = ((uint32_t)node < td.n_fast_nodes) ? td.fast_nodes[node] : td.nodes[node];
            float tn0, tn1;
            // `<=` culling keeps boxes that may hold an equal-t triangle with a lower id
            bool h0 = box_test(n.lo0, n.hi0, o, inv_d, t_min, best_t, tn0);
            bool h1 = box_test(n.lo1, n.hi1, o, inv_d, t_min, best_t, tn1);
            if (h0 && h1) {
                int32_t near = n.c0, far = n.c1;
                if (tn1 < tn0) {
                    near = n.c1;
                    far = n.c0;
                }
                if (sp < AKR_BVH_STACK) stack[sp++] = far;
                node = near;
                continue;
            }
            if (h0) {
                node = n.c0;
                continue;
            }
            if (h1) {
                node = n.c1;
                continue;
            }
        } else {
            uint32_t leaf = (uint32_t)(~node);
            uint32_t first = (leaf >> 3) * 2u, count = (leaf & 7u) * 2u;  // primitives -> triangle slots
            for (uint32_t k = 0; k < count; ++k) {
                const TriGeom &tr = tris[first + k];
                uint32_t gid = tr.gid;
                if (gid == 0xffffffffu || gid == ex0 || gid == ex1) continue;
                float t, u, v;
                if (!tri_test(tr, o, d, t_min, t_max, t, u, v)) continue;
                bool closer = (t < best_t) || (t == best_t && gid < best.gid);
                if (!closer) continue;
                if (sc.any_alpha && !alpha_test(sc, gid, u, v)) continue;
                best = HitRec{gid, u, v};
                best_t = t;
                if (ANY_HIT) return best;
            }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return best;
}

// Slab test of child k of a 4-wide node: box_test_fma on the transposed planes (inv_d from capped_inv_dir).
AKR_HD bool box4_test(const Bvh4Node &n, int k, f3 inv_d, f3 ood, float t_min, float t_max, float &t_near) {
    float tx0 = fmaf(n.lo[0][k], inv_d.x, ood.x), tx1 = fmaf(n.hi[0][k], inv_d.x, ood.x);
    float ty0 = fmaf(n.lo[1][k], inv_d.y, ood.y), ty1 = fmaf(n.hi[1][k], inv_d.y, ood.y);
    float tz0 = fmaf(n.lo[2][k], inv_d.z, ood.z), tz1 = fmaf(n.hi[2][k], inv_d.z, ood.z);
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), t_min));
    float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), t_max));
    t_near = tn;
    return tn <= tf;
}
// One step of the 4-wide traversal: tests the four children of `n`, returns the nearest hit child (or 0x7fffffff when
// none is hit) and writes the other hit children to `push` (n_push of them).
AKR_HD int32_t bvh4_step(const Bvh4Node &n, f3 inv_d, f3 ood, float t_min, float t_max, int32_t push[3], int &n_push) {
    float tn[4];
    bool hit[4];
    for (int k = 0; k < 4; ++k) hit[k] = box4_test(n, k, inv_d, ood, t_min, t_max, tn[k]);
    int best = -1;
    for (int k = 0; k < 4; ++k)
        if (hit[k] && (best < 0 || tn[k] < tn[best])) best = k;
    n_push = 0;
    if (best < 0) return 0x7fffffff;
    for (int k = 0; k < 4; ++k)
        if (hit[k] && k != best) push[n_push++] = n.c[k];
    return n.c[best];
}

// Moeller-Trumbore closest / any hit over the 4-wide tree (host simulation: validates the collapsed tree bit for bit
// against the oracle's brute-force scan; the tie rule makes the result independent of the visiting order).
template <bool ANY_HIT>
AKR_HD HitRec trace_ray4(const SceneView &sc, f3 o, f3 d, float t_min, float t_max, uint32_t ex0, uint32_t ex1) {
    const TriGeom *tris = sc.tris;
    HitRec best{0xffffffffu, 0.0f, 0.0f};
    float best_t = t_max;
    const f3 inv_d = capped_inv_dir(d);  // finite: plane * inv_d + ood must never be inf - inf (see box_test_fma)
    const f3 ood = mk3(-o.x * inv_d.x, -o.y * inv_d.y, -o.z * inv_d.z);
    int32_t stack[3 * AKR_BVH_STACK];
    int sp = 0;
    int32_t node = 0;
    while (true) {
        if (node >= 0) {
            int32_t push[3];
            int n_push;
            int32_t next = bvh4_step(sc.nodes4[node], inv_d, ood, t_min, best_t, push, n_push);
            for (int k = 0; k < n_push; ++k)
                if (sp < 3 * AKR_BVH_STACK) stack[sp++] = push[k];
            if (next != 0x7fffffff) {
                node = next;
                continue;
            }
        } else {
            uint32_t leaf = (uint32_t)(~node);
            uint32_t first = (leaf >> 3) * 2u, count = (leaf & 7u) * 2u;
            for (uint32_t k = 0; k < count; ++k) {
                const TriGeom &tr = tris[first + k];
                uint32_t gid = tr.gid;
                if (gid == 0xffffffffu || gid == ex0 || gid == ex1) continue;
                float t, u, v;
                if (!tri_test(tr, o, d, t_min, t_max, t, u, v)) continue;
                bool closer = (t < best_t) || (t == best_t && gid < best.gid);
                if (!closer) continue;
                if (sc.any_alpha && !alpha_test(sc, gid, u, v)) continue;
                best = HitRec{gid, u, v};
                best_t = t;
                if (ANY_HIT) return best;
            }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return best;
}

// ---- primitive (triangle / parallelogram pair) intersector used by the CUDA kernels -----------------
struct PrimHit {
    float t, s, q;
    uint32_t k;  // primitive index, 0xffffffff = none
};
struct PrimDecoded {
    uint32_t gid, cls;
    float u, v;
    uint32_t light;  // the triangle belongs to a mesh light
};
AKR_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU; 1 ulp, 0 -> inf
    return r;
#else
    return 1.0f / x;
#endif
}
// which triangle of the primitive a point (s, q) belongs to, and that triangle's own barycentrics
AKR_HD PrimDecoded prim_decode(const PrimRec &p, float s, float q) {
    PrimDecoded r;
    if (p.gid_b == 0xffffffffu) {
        r.gid = p.gid_a;
        r.cls = (p.meta >> 8) & 3u;
        r.light = (p.meta >> 12) & 1u;
        r.u = s;
        r.v = q;
        return r;
    }
    const bool half = s < q;
    const float w0 = half ? 1.0f - q : 1.0f - s;
    const float w1 = half ? s : s - q;
    const float w2 = half ? q - s : q;
    const uint32_t m = half ? (p.meta >> 4) : p.meta;
    const uint32_t iu = m & 3u, iv = (m >> 2) & 3u;
    r.gid = half ? p.gid_b : p.gid_a;
    r.cls = (half ? (p.meta >> 10) : (p.meta >> 8)) & 3u;
    r.light = (half ? (p.meta >> 13) : (p.meta >> 12)) & 1u;
    r.u = iu == 0u ? w0 : (iu == 1u ? w1 : w2);
    r.v = iv == 0u ? w0 : (iv == 1u ? w1 : w2);
    return r;
}
// One candidate: updates `best` when primitive k is hit at t in (t_min, best.t) by a triangle that is not
// excluded.  Branch-free on purpose: every lane of a warp runs the same instructions per candidate and commits
// with one predicate (a NaN from a parallel ray or a degenerate primitive fails every comparison).  The
// operation order (explicit FMAs) and the inside tests are exactly those of the packed FFMA2 loop of the flat
// trace mode (akari_b200.cu: trace_flat2), so BVH and flat traversal compute identical (t, s, q) per primitive.
// KIND: 0 = decide per primitive, 1 = known pair, 2 = known single.
AKR_HD float prim_plane_t(const PrimRec &p, f3 o, f3 d) {
    const float den = fmaf(p.n[2], d.z, fmaf(p.n[1], d.y, p.n[0] * d.x));
    const float num = fmaf(p.n[2], o.z, fmaf(p.n[1], o.y, fmaf(p.n[0], o.x, p.n[3])));
    return num * fast_rcp(-den);  // t = -(n.o + nw) / (n.d)
}
AKR_HD void prim_coords(const PrimRec &p, f3 o, f3 d, float t, float &s, float &q) {
    const float hx = fmaf(t, d.x, o.x), hy = fmaf(t, d.y, o.y), hz = fmaf(t, d.z, o.z);
    s = fmaf(p.r0[0], hx, fmaf(p.r0[1], hy, fmaf(p.r0[2], hz, p.r0[3])));
    q = fmaf(p.r1[0], hx, fmaf(p.r1[1], hy, fmaf(p.r1[2], hz, p.r1[3])));
}
// pair: 0 <= s, q <= 1 tested as |s - 0.5| <= 0.5 and |q - 0.5| <= 0.5; single: s, q >= 0, s + q <= 1
AKR_HD bool prim_inside(bool pair, float s, float q) {
    return pair ? ((fabsf(s + -0.5f) <= 0.5f) & (fabsf(q + -0.5f) <= 0.5f)) : ((s >= 0.0f) & (q >= 0.0f) & (s + q <= 1.0f));
}
template <bool ALPHA, int KIND = 0>
AKR_HD void prim_test(const SceneView &sc, const PrimRec &p, uint32_t k, f3 o, f3 d, float t_min, uint32_t ex0, uint32_t ex1, PrimHit &best) {
    const float t = prim_plane_t(p, o, d);
    float s, q;
    prim_coords(p, o, d, t, s, q);
    const bool pair = KIND == 1 ? true : (KIND == 2 ? false : p.gid_b != 0xffffffffu);
    const uint32_t gid = (pair & (s < q)) ? p.gid_b : p.gid_a;
    bool ok = prim_inside(pair, s, q) & (t > t_min) & (t < best.t) & (gid != ex0) & (gid != ex1);
    if (ALPHA) {
        if (ok) {
            PrimDecoded dec = prim_decode(p, s, q);
            ok = alpha_test(sc, dec.gid, dec.u, dec.v);
        }
    }
    best.t = ok ? t : best.t;
    best.s = ok ? s : best.s;
    best.q = ok ? q : best.q;
    best.k = ok ? k : best.k;
}

// Reference (host / CLS-agnostic) traversal over primitives: same BVH, same candidate order as trace_ray.
template <bool ANY_HIT>
AKR_HD HitRec trace_ray_prims(const SceneView &sc, f3 o, f3 d, float t_min, float t_max, uint32_t ex0, uint32_t ex1) {
    PrimHit best{t_max, 0.0f, 0.0f, 0xffffffffu};
    f3 inv_d = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int32_t stack[AKR_BVH_STACK];
    int sp = 0;
    int32_t node = 0;
    while (true) {
        if (node >= 0) {
            const BvhNode n = sc.nodes[node];
            float tn0, tn1;
            bool h0 = box_test(n.lo0, n.hi0, o, inv_d, t_min, best.t, tn0);
            bool h1 = box_test(n.lo1, n.hi1, o, inv_d, t_min, best.t, tn1);
            if (h0 && h1) {
                int32_t near = n.c0, far = n.c1;
                if (tn1 < tn0) {
                    near = n.c1;
                    far = n.c0;
                }
                if (sp < AKR_BVH_STACK) stack[sp++] = far;
                node = near;
                continue;
            }
            if (h0) {
                node = n.c0;
                continue;
            }
            if (h1) {
                node = n.c1;
                continue;
            }
        } else {
            uint32_t leaf = (uint32_t)(~node);
            uint32_t first = leaf >> 3, count = leaf & 7u;
            for (uint32_t k = 0; k < count; ++k) {
                if (sc.any_alpha) prim_test<true>(sc, sc.prims[first + k], first + k, o, d, t_min, ex0, ex1, best);
                else prim_test<false>(sc, sc.prims[first + k], first + k, o, d, t_min, ex0, ex1, best);
                if (ANY_HIT && best.k != 0xffffffffu) break;
            }
            if (ANY_HIT && best.k != 0xffffffffu) break;
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    if (best.k == 0xffffffffu) return HitRec{0xffffffffu, 0.0f, 0.0f};
    PrimDecoded dec = prim_decode(sc.prims[best.k], best.s, best.q);
    return HitRec{dec.gid, dec.u, dec.v};
}

// Reference walk over the staged flat lists (PrimBlock2, akr_scene.cuh): closest-hit rays test the complete list,
// any-hit rays the occluder-only list behind it — what akari_b200.cu: trace_flat2 does with packed FMAs.  Used by the
// host simulation to check the block layout and that leaving out the scene-supporting primitives never changes an
// occlusion result.
AKR_HD PrimRec block_prim(const PrimBlock2 &b, uint32_t h) {
    PrimRec p;
    for (int c = 0; c < 4; ++c) {
        p.n[c] = b.n[c][h];
        p.r0[c] = b.r0[c][h];
        p.r1[c] = b.r1[c][h];
    }
    p.gid_a = b.gid[2u * h];
    p.gid_b = b.gid[2u * h + 1u];
    p.meta = b.meta[h];
    p._pad = 0u;
    return p;
}
template <bool ANY_HIT>
AKR_HD HitRec trace_flat_ref(const SceneView &sc, f3 o, f3 d, float t_min, float t_max, uint32_t ex0, uint32_t ex1) {
    const uint32_t b0 = ANY_HIT ? sc.n_pair_blocks + sc.n_single_blocks : 0u;
    const uint32_t nb = ANY_HIT ? sc.n_occ_pair_blocks + sc.n_occ_single_blocks : sc.n_pair_blocks + sc.n_single_blocks;
    PrimHit best{t_max, 0.0f, 0.0f, 0xffffffffu};
    for (uint32_t b = b0; b < b0 + nb; ++b)
        for (uint32_t h = 0; h < 2u; ++h) {
            const PrimRec p = block_prim(sc.flat_blocks[b], h);
            if (sc.any_alpha) prim_test<true>(sc, p, 2u * b + h, o, d, t_min, ex0, ex1, best);
            else prim_test<false>(sc, p, 2u * b + h, o, d, t_min, ex0, ex1, best);
            if (ANY_HIT && best.k != 0xffffffffu) return HitRec{0u, 0.0f, 0.0f};
        }
    if (best.k == 0xffffffffu) return HitRec{0xffffffffu, 0.0f, 0.0f};
    PrimDecoded dec = prim_decode(block_prim(sc.flat_blocks[best.k >> 1], best.k & 1u), best.s, best.q);
    return HitRec{dec.gid, dec.u, dec.v};
}

}  // namespace akr

// akari_b200.cu — wavefront path-tracing kernels (sm_100a) and the C-ABI of include/akari_b200.h.
//
// Two pipelines per wave of W paths (state = structure-of-arrays of 16-byte records in HBM):
//
// FUSED (small scenes: <= 64 primitives, no stochastic alpha — cbox):
//   k_raygen_fused -> [ k_bounce<class>(d) for every shade class present ] for d = 0 .. max_depth - 1 -> k_accumulate
//   One kernel per (depth, shade class) does a whole bounce: shade, the NEE shadow ray, the continuation ray and the
//   emitter term of the hit it finds (akr_path.cuh: bounce_fused).  Both rays are tested against the primitive list
//   staged in shared memory, two primitives per packed FP32 instruction (FFMA2), every lane in lock step.  Path
//   records (hit + throughput + radiance, 64 B) stream through per-class queues: each warp pulls its next 32 records
//   into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered) while it works on the
//   current ones, and appends survivors to the queue of the class of the NEXT hit with warp-aggregated atomics.
//   There is no hit queue, no shadow queue, no (slot, path id) indirection and no accumulator read-modify-write.
//   The general class (full Principled tree, glass, texture-driven materials) is fed differently: k_sort_hist /
//   k_sort_scatter order a depth's records by material signature and the warps gather through that order, so that a
//   warp — and the SM — runs one closure tree at a time.
//
// QUEUED (BVH scenes, and flat scenes with alpha-tested materials):
//   k_raygen -> [ trace(d) -> k_shade<class>(d) ] for d = 0 .. max_depth -> k_accumulate
//   trace(d) = k_trace_bvh: persistent warps with dynamic ray fetch and majority-vote rounds over a BVH2 (top of the tree
//   in shared memory), closest-hit rays of depth d and the shadow rays shade(d - 1) queued (k_trace_flat on flat scenes).
//   Hits are binned by the shade class of their material and one shade kernel is compiled per class.
//
// Every kernel is a persistent grid-stride kernel sized SM count x resident CTAs; queue lengths live in device memory,
// so a whole pass is enqueued without a host round trip.  Scene data is staged into shared memory with one TMA bulk
// copy per CTA.  There is no tensor-core work on this path (no dense contraction exists in a path tracer) and no CPU
// fallback: every entry point fails with AKR_ERR_CUDA when no device is usable.
#include "../../include/akari_b200.h"
#include "device/akr_path.cuh"
#include "host/scene_build.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace akr;

namespace {

constexpr int kBlock = 256;
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr uint32_t kMaxDepthSlots = 66;           // counters for depth 0 .. 64 (+1)
// Queue counters, one 32-byte block per depth d:
//   [0] paths entering depth d (raygen for d = 0, else the survivors of shade(d - 1))
//   [1] shadow rays produced by shade(d - 1), traced together with [0] by trace(d)
//   [2 + c] hits of shade class c found by trace(d)
//   [5] / [6] next closest-hit / shadow ray handed out by the dynamic-fetch BVH trace kernel
// [0]/[1] and [2]/[3] are 8-byte aligned pairs, so one 64-bit atomic reserves slots in two queues at once.
constexpr uint32_t kCtrStride = 8;
constexpr uint32_t kSmemSceneBudget = 48 * 1024;  // bytes of BVH nodes + triangles staged per CTA
constexpr uint32_t kSmemMax = 200 * 1024;         // opt-in ceiling for the trace kernels (staging + stacks)

// ------------------------------------------------------------------------------------------------
// wave state in HBM: structure-of-arrays of 16-byte records (one LDG.128 / STG.128 per record, fully
// coalesced across a warp)
// ------------------------------------------------------------------------------------------------
struct PathQueue {   // 48 B per path
    f4 *a;           // origin.xyz, dir.x
    f4 *b;           // dir.y, dir.z, exclude gid (bits), path_id (bits)   -- a + b is all the trace stage reads
    f4 *c;           // beta.rgb, prev_bsdf_pdf
};
struct HitQueue {    // 16 B per traced path, same slot as the path
    f4 *h;           // gid (bits), u, v, -
};
struct ShadowQueue { // 52 B per shadow ray
    f4 *a;           // origin.xyz, t_max
    f4 *b;           // dir.xyz, exclude0 (bits)
    f4 *c;           // contribution.rgb, path_id (bits)
    uint32_t *ex1;
};
struct ClassQueues { // per shade class: (slot in the path queue, path_id) of the hits of that class
    uint2 *idx[CLS_COUNT];
};
struct RecQueue {    // fused pipeline: 64 B per path = BounceRec (akr_path.cuh), one queue per (depth parity, shade class)
    f4 *r[4];        // (d.xyz, gid) | (u, v, path_id, pixel) | (beta.rgb, sample index) | (L.rgb, -)
};

struct LaunchParams {
    SceneView scene;
    CornerAttribs corners;
    SamplerTables tables;
    RenderParams rp;
    WaveInfo wave;
    PathQueue q[2];
    HitQueue hits;
    ShadowQueue shadow;
    ClassQueues cls;
    RecQueue cq[2][CLS_COUNT];
    AccView acc;
    uint32_t *counters;       // [kMaxDepthSlots][kCtrStride]
    float *film;
    uint32_t n_film_pixels;
    uint32_t scene_smem_nodes;  // nodes staged in shared memory
    uint32_t scene_smem_prims;  // 1 when all primitives are staged too
    uint32_t stack_depth;       // traversal stack entries per thread (shared memory)
    uint32_t stage_flat;        // stage the pairs-first flat list instead of the BVH-ordered primitives
    uint32_t aov, aov_remap;    // k_aov: which AKR_AOV_* quantity, v -> v * 0.5 + 0.5
    uint32_t *first_hits;       // optional AOV (engine option aov_mask bit 0): [n_film_pixels][2] (inst, prim) of sample 0
    uint32_t *sort_order;       // general class, global sort: record indices of a depth ordered by material sort key
    uint32_t *sort_hist;        // [kMaxDepthSlots][64]: per key the record count [0, 32) and the scatter cursor [32, 64)
};

// Queue records are written once and read once, gigabytes later: stream them past L2 residency (evict-first) so that
// the L2 keeps the sampler tables, the per-triangle records and the accumulators.  AKR_STREAM_QUEUES=0 restores the
// default policy (A/B runs).
#ifndef AKR_STREAM_QUEUES
#define AKR_STREAM_QUEUES 1
#endif
__device__ __forceinline__ f4 ldq(const f4 *p) {
#if AKR_STREAM_QUEUES
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
    return f4{v.x, v.y, v.z, v.w};
#else
    return ld4(p);
#endif
}
__device__ __forceinline__ void stq(f4 *p, f4 v) {
#if AKR_STREAM_QUEUES
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v.x, v.y, v.z, v.w));
#else
    st4(p, v);
#endif
}
__device__ __forceinline__ PathState load_path(const PathQueue &q, uint32_t i) {
    const f4 a = ldq(q.a + i), b = ldq(q.b + i), c = ldq(q.c + i);
    PathState p;
    p.o = mk3(a.x, a.y, a.z);
    p.d = mk3(a.w, b.x, b.y);
    p.ex = f2u(b.z);
    p.path_id = f2u(b.w);
    p.beta = mk3(c.x, c.y, c.z);
    p.prev_bsdf_pdf = c.w;
    return p;
}
__device__ __forceinline__ void store_path(const PathQueue &q, uint32_t i, const PathState &p) {
    stq(q.a + i, f4{p.o.x, p.o.y, p.o.z, p.d.x});
    stq(q.b + i, f4{p.d.y, p.d.z, u2f(p.ex), u2f(p.path_id)});
    stq(q.c + i, f4{p.beta.x, p.beta.y, p.beta.z, p.prev_bsdf_pdf});
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copy of the traversal data into shared memory
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Shared-memory plan of the trace kernel: [ nodes | primitives | per-thread traversal stacks ].
// Addresses are kept as 32-bit shared-space addresses and dereferenced with explicit ld.shared / st.shared:
// through a generic pointer the compiler emits generic LD/ST, which are slower than LDS/STS.
struct TraceSmem {
    uint32_t nodes;         // shared copy of nodes[0 .. n_fast_nodes)
    uint32_t prims;         // shared copy of all primitives (when staged)
    uint32_t stack;         // this thread's stack: entry k lives at stack + k * kBlock * 4
};
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int32_t lds32(uint32_t addr) {
    int32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, int32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
template <bool SHARED> __device__ __forceinline__ void load64(uint32_t saddr, const void *gptr, float4 &a, float4 &b, float4 &c, float4 &d) {
    if (SHARED) {
        a = lds128(saddr);
        b = lds128(saddr + 16u);
        c = lds128(saddr + 32u);
        d = lds128(saddr + 48u);
    } else {
        const float4 *g = static_cast<const float4 *>(gptr);
        a = __ldg(g);
        b = __ldg(g + 1);
        c = __ldg(g + 2);
        d = __ldg(g + 3);
    }
}
template <bool SHARED> __device__ __forceinline__ PrimRec load_prim(uint32_t s_prims, const PrimRec *g_prims, uint32_t k) {
    float4 a, b, c, d;
    load64<SHARED>(s_prims + k * (uint32_t)sizeof(PrimRec), g_prims + k, a, b, c, d);
    PrimRec p;
    p.n[0] = a.x; p.n[1] = a.y; p.n[2] = a.z; p.n[3] = a.w;
    p.r0[0] = b.x; p.r0[1] = b.y; p.r0[2] = b.z; p.r0[3] = b.w;
    p.r1[0] = c.x; p.r1[1] = c.y; p.r1[2] = c.z; p.r1[3] = c.w;
    p.gid_a = __float_as_uint(d.x);
    p.gid_b = __float_as_uint(d.y);
    p.meta = __float_as_uint(d.z);
    p._pad = 0u;
    return p;
}
// primitive h (0 / 1) of a staged PrimBlock2, back in PrimRec form
__device__ __forceinline__ PrimRec load_block_prim(uint32_t block_addr, uint32_t h) {
    PrimRec p;
    const uint32_t a = block_addr + h * 4u;
    for (int c = 0; c < 4; ++c) {
        p.n[c] = __int_as_float(lds32(a + c * 8u));
        p.r0[c] = __int_as_float(lds32(a + 32u + c * 8u));
        p.r1[c] = __int_as_float(lds32(a + 64u + c * 8u));
    }
    p.gid_a = (uint32_t)lds32(block_addr + 96u + h * 8u);
    p.gid_b = (uint32_t)lds32(block_addr + 100u + h * 8u);
    p.meta = (uint32_t)lds32(block_addr + 112u + h * 4u);
    p._pad = 0u;
    return p;
}
template <bool SHARED> __device__ __forceinline__ BvhNode load_node(uint32_t s_nodes, const BvhNode *g_nodes, uint32_t k) {
    float4 a, b, c, d;
    load64<SHARED>(s_nodes + k * (uint32_t)sizeof(BvhNode), g_nodes + k, a, b, c, d);
    BvhNode n;
    n.lo0[0] = a.x; n.lo0[1] = a.y; n.lo0[2] = a.z; n.hi0[0] = a.w;
    n.hi0[1] = b.x; n.hi0[2] = b.y; n.lo1[0] = b.z; n.lo1[1] = b.w;
    n.lo1[2] = c.x; n.hi1[0] = c.y; n.hi1[1] = c.z; n.hi1[2] = c.w;
    n.c0 = (int32_t)__float_as_uint(d.x);
    n.c1 = (int32_t)__float_as_uint(d.y);
    return n;
}

// Stages nodes[0 .. n_nodes) and (optionally) all primitives behind `smem` with one TMA bulk copy each.
__device__ __forceinline__ TraceSmem stage_scene(const LaunchParams &P, unsigned char *smem, uint64_t *bar) {
    const uint32_t node_bytes = P.scene_smem_nodes * (uint32_t)sizeof(BvhNode);
    const uint32_t tri_bytes = P.stage_flat ? (P.scene.n_pair_blocks + P.scene.n_single_blocks + P.scene.n_occ_pair_blocks + P.scene.n_occ_single_blocks) *
                                                  (uint32_t)sizeof(PrimBlock2)
                                            : (P.scene_smem_prims ? P.scene.n_prims * (uint32_t)sizeof(PrimRec) : 0u);
    BvhNode *s_nodes = reinterpret_cast<BvhNode *>(smem);
    PrimRec *s_tris = reinterpret_cast<PrimRec *>(smem + node_bytes);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && (node_bytes + tri_bytes) > 0) {
        mbar_expect_tx(bar, node_bytes + tri_bytes);
        if (node_bytes) tma_bulk_g2s(s_nodes, P.scene.nodes, node_bytes, bar);
        if (tri_bytes) tma_bulk_g2s(s_tris, P.stage_flat ? (const void *)P.scene.flat_blocks : (const void *)P.scene.prims, tri_bytes, bar);
    }
    if ((node_bytes + tri_bytes) > 0) mbar_wait(bar, 0);
    TraceSmem t;
    t.nodes = smem_u32(s_nodes);
    t.prims = smem_u32(s_tris);
    t.stack = smem_u32(smem + node_bytes + tri_bytes) + threadIdx.x * 4u;
    return t;
}

// ------------------------------------------------------------------------------------------------
// device traversal over primitives (akr_trace.cuh::prim_test): one plane + two-coordinate test decides a
// triangle or both halves of a parallelogram.  The host simulation keeps the Moeller-Trumbore
// trace_ray on the same BVH; the two agree up to rounding at primitive edges.
// ------------------------------------------------------------------------------------------------
struct DevHit {
    uint32_t gid, cls;
    float u, v;
};

// TRACE_BVH: top of the BVH in shared memory, the rest of the nodes and the primitives in global memory (L1/L2);
// TRACE_FLAT / TRACE_BVH_SMEM: the whole scene is staged in shared memory.
enum TraceMode : int { TRACE_BVH = 0, TRACE_FLAT = 1, TRACE_BVH_SMEM = 2 };

// TRACE_FLAT: every lane tests every primitive in list order (shared-memory broadcast reads, no divergence,
// no stack) — the cheapest schedule when the whole scene is a few dozen primitives.
// ---- flat mode with packed FP32 (FFMA2) ----------------------------------------------------------------
// Blackwell issues fma.rn.f32x2: one instruction = two FP32 FMAs per lane.  The staged PrimBlock2 layout puts
// the same coefficient of two primitives side by side, so the 16 FMAs of a primitive test advance TWO
// primitives per instruction (the FMA pipe was the limiter of the scalar loop).  Only (t, k) of the best
// candidate are kept; (s, q) are recomputed once after the loop with the same operation order.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void lds2x64(uint32_t addr, u64 &a, u64 &b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float rcp_neg(float x) {  // 1 / (-x): the negation folds into the MUFU operand modifier
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-x));
    return r;
}

// one PrimBlock2 (two primitives of the same kind) against one ray
template <bool PAIR, bool ALPHA>
__device__ __forceinline__ void flat2_block(const SceneView &sc, uint32_t addr, uint32_t b, f3 o, f3 d, float t_min, uint32_t ex0, uint32_t ex1, float &best_t,
                                            uint32_t &best_k) {
    const u64 ox2 = pk2(o.x, o.x), oy2 = pk2(o.y, o.y), oz2 = pk2(o.z, o.z);  // ptxas turns these into the scalar-broadcast operand form
    const u64 dx2 = pk2(d.x, d.x), dy2 = pk2(d.y, d.y), dz2 = pk2(d.z, d.z);
    u64 n0, n1, n2, nw, r00, r01, r02, r0w, r10, r11, r12, r1w;
    lds2x64(addr, n0, n1);
    lds2x64(addr + 16u, n2, nw);
    lds2x64(addr + 32u, r00, r01);
    lds2x64(addr + 48u, r02, r0w);
    lds2x64(addr + 64u, r10, r11);
    lds2x64(addr + 80u, r12, r1w);
    const uint4 g = lds_u4(addr + 96u);
    const u64 den = fma2(n2, dz2, fma2(n1, dy2, mul2(n0, dx2)));
    const u64 num = fma2(n2, oz2, fma2(n1, oy2, fma2(n0, ox2, nw)));
    float den0, den1;
    upk2(den, den0, den1);
    const u64 t2 = mul2(num, pk2(rcp_neg(den0), rcp_neg(den1)));  // t = -(n.o + nw) / (n.d)
    const u64 hx = fma2(t2, dx2, ox2), hy = fma2(t2, dy2, oy2), hz = fma2(t2, dz2, oz2);
    const u64 s2 = fma2(r00, hx, fma2(r01, hy, fma2(r02, hz, r0w)));
    const u64 q2 = fma2(r10, hx, fma2(r11, hy, fma2(r12, hz, r1w)));
    float t0, t1, s0, s1, q0, q1;
    upk2(t2, t0, t1);
    upk2(s2, s0, s1);
    upk2(q2, q0, q1);
    bool in0, in1;
    uint32_t gid0, gid1;
    if (PAIR) {  // akr_trace.cuh: prim_inside(pair = true)
        const u64 mhalf2 = pk2(-0.5f, -0.5f);
        float a0, a1, c0, c1;
        upk2(add2(s2, mhalf2), a0, a1);
        upk2(add2(q2, mhalf2), c0, c1);
        in0 = (fabsf(a0) <= 0.5f) & (fabsf(c0) <= 0.5f);
        in1 = (fabsf(a1) <= 0.5f) & (fabsf(c1) <= 0.5f);
        gid0 = s0 < q0 ? g.y : g.x;
        gid1 = s1 < q1 ? g.w : g.z;
    } else {
        float m0, m1;
        upk2(add2(s2, q2), m0, m1);
        in0 = (s0 >= 0.0f) & (q0 >= 0.0f) & (m0 <= 1.0f);
        in1 = (s1 >= 0.0f) & (q1 >= 0.0f) & (m1 <= 1.0f);
        gid0 = g.x;
        gid1 = g.z;
    }
    bool ok0 = in0 & (t0 > t_min) & (t0 < best_t) & (gid0 != ex0) & (gid0 != ex1);
    if (ALPHA) {
        if (ok0) {
            PrimDecoded dec = prim_decode(load_block_prim(addr, 0u), s0, q0);
            ok0 = alpha_test(sc, dec.gid, dec.u, dec.v);
        }
    }
    best_t = ok0 ? t0 : best_t;
    best_k = ok0 ? 2u * b : best_k;
    bool ok1 = in1 & (t1 > t_min) & (t1 < best_t) & (gid1 != ex0) & (gid1 != ex1);
    if (ALPHA) {
        if (ok1) {
            PrimDecoded dec = prim_decode(load_block_prim(addr, 1u), s1, q1);
            ok1 = alpha_test(sc, dec.gid, dec.u, dec.v);
        }
    }
    best_t = ok1 ? t1 : best_t;
    best_k = ok1 ? 2u * b + 1u : best_k;
}

// returns the index of the hit primitive in the staged list (0xffffffff = none) and its distance
template <bool ANY_HIT, bool ALPHA>
__device__ __forceinline__ uint32_t trace_flat2_core(const SceneView &sc, const TraceSmem &ts, bool active, f3 o, f3 d, float t_min, float t_max, uint32_t ex0,
                                                     uint32_t ex1, float &best_t) {
    best_t = active ? t_max : 0.0f;  // an idle lane accepts nothing
    uint32_t best_k = 0xffffffffu;
    // closest-hit rays walk the complete list, any-hit rays the occluder-only list staged right behind it
    const uint32_t b0 = ANY_HIT ? sc.n_pair_blocks + sc.n_single_blocks : 0u;
    const uint32_t n_pair_blocks = b0 + (ANY_HIT ? sc.n_occ_pair_blocks : sc.n_pair_blocks);
    const uint32_t n_blocks = n_pair_blocks + (ANY_HIT ? sc.n_occ_single_blocks : sc.n_single_blocks);
    bool all_done = false;
#pragma unroll 1
    for (uint32_t b = b0; b < n_pair_blocks; ++b) {
        flat2_block<true, ALPHA>(sc, ts.prims + b * (uint32_t)sizeof(PrimBlock2), b, o, d, t_min, ex0, ex1, best_t, best_k);
        if (ANY_HIT && (b & 3u) == 3u && __all_sync(0xffffffffu, !active || best_k != 0xffffffffu)) {
            all_done = true;
            break;
        }
    }
    if (!all_done) {
#pragma unroll 1
        for (uint32_t b = n_pair_blocks; b < n_blocks; ++b) {
            flat2_block<false, ALPHA>(sc, ts.prims + b * (uint32_t)sizeof(PrimBlock2), b, o, d, t_min, ex0, ex1, best_t, best_k);
            if (ANY_HIT && (b & 3u) == 3u && __all_sync(0xffffffffu, !active || best_k != 0xffffffffu)) break;
        }
    }
    return best_k;
}
// Closest-hit ray A and any-hit ray B of the same lane against one PrimBlock2, the block loaded ONCE.  Two independent
// dependency chains per trip also give the scheduler twice the instruction-level parallelism of the single-ray loop.
template <bool PAIR, bool WITH_B>
__device__ __forceinline__ void flat2_block_dual(uint32_t addr, uint32_t b, f3 oa, f3 da, uint32_t exa, float &best_t, uint32_t &best_k, f3 ob, f3 db, float tmax_b,
                                                 uint32_t exb0, uint32_t exb1, bool &occluded) {
    u64 n0, n1, n2, nw, r00, r01, r02, r0w, r10, r11, r12, r1w;
    lds2x64(addr, n0, n1);
    lds2x64(addr + 16u, n2, nw);
    lds2x64(addr + 32u, r00, r01);
    lds2x64(addr + 48u, r02, r0w);
    lds2x64(addr + 64u, r10, r11);
    lds2x64(addr + 80u, r12, r1w);
    const uint4 g = lds_u4(addr + 96u);
    auto coords = [&](f3 o, f3 d, float &t0, float &t1, float &s0, float &s1, float &q0, float &q1, u64 &s2, u64 &q2) {
        const u64 ox2 = pk2(o.x, o.x), oy2 = pk2(o.y, o.y), oz2 = pk2(o.z, o.z);
        const u64 dx2 = pk2(d.x, d.x), dy2 = pk2(d.y, d.y), dz2 = pk2(d.z, d.z);
        const u64 den = fma2(n2, dz2, fma2(n1, dy2, mul2(n0, dx2)));
        const u64 num = fma2(n2, oz2, fma2(n1, oy2, fma2(n0, ox2, nw)));
        float den0, den1;
        upk2(den, den0, den1);
        const u64 t2 = mul2(num, pk2(rcp_neg(den0), rcp_neg(den1)));
        const u64 hx = fma2(t2, dx2, ox2), hy = fma2(t2, dy2, oy2), hz = fma2(t2, dz2, oz2);
        s2 = fma2(r00, hx, fma2(r01, hy, fma2(r02, hz, r0w)));
        q2 = fma2(r10, hx, fma2(r11, hy, fma2(r12, hz, r1w)));
        upk2(t2, t0, t1);
        upk2(s2, s0, s1);
        upk2(q2, q0, q1);
    };
    auto inside = [&](float s0, float s1, float q0, float q1, u64 s2, u64 q2, bool &in0, bool &in1, uint32_t &gid0, uint32_t &gid1) {
        if (PAIR) {
            const u64 mhalf2 = pk2(-0.5f, -0.5f);
            float a0, a1, c0, c1;
            upk2(add2(s2, mhalf2), a0, a1);
            upk2(add2(q2, mhalf2), c0, c1);
            in0 = (fabsf(a0) <= 0.5f) & (fabsf(c0) <= 0.5f);
            in1 = (fabsf(a1) <= 0.5f) & (fabsf(c1) <= 0.5f);
            gid0 = s0 < q0 ? g.y : g.x;
            gid1 = s1 < q1 ? g.w : g.z;
        } else {
            float m0, m1;
            upk2(add2(s2, q2), m0, m1);
            in0 = (s0 >= 0.0f) & (q0 >= 0.0f) & (m0 <= 1.0f);
            in1 = (s1 >= 0.0f) & (q1 >= 0.0f) & (m1 <= 1.0f);
            gid0 = g.x;
            gid1 = g.z;
        }
    };
    {
        float t0, t1, s0, s1, q0, q1;
        u64 s2, q2;
        coords(oa, da, t0, t1, s0, s1, q0, q1, s2, q2);
        bool in0, in1;
        uint32_t gid0, gid1;
        inside(s0, s1, q0, q1, s2, q2, in0, in1, gid0, gid1);
        const bool ok0 = in0 & (t0 > 0.0f) & (t0 < best_t) & (gid0 != exa);
        best_t = ok0 ? t0 : best_t;
        best_k = ok0 ? 2u * b : best_k;
        const bool ok1 = in1 & (t1 > 0.0f) & (t1 < best_t) & (gid1 != exa);
        best_t = ok1 ? t1 : best_t;
        best_k = ok1 ? 2u * b + 1u : best_k;
    }
    if (WITH_B) {
        float t0, t1, s0, s1, q0, q1;
        u64 s2, q2;
        coords(ob, db, t0, t1, s0, s1, q0, q1, s2, q2);
        bool in0, in1;
        uint32_t gid0, gid1;
        inside(s0, s1, q0, q1, s2, q2, in0, in1, gid0, gid1);
        occluded |= in0 & (t0 > 0.0f) & (t0 < tmax_b) & (gid0 != exb0) & (gid0 != exb1);
        occluded |= in1 & (t1 > 0.0f) & (t1 < tmax_b) & (gid1 != exb0) & (gid1 != exb1);
    }
}
// One walk over the complete staged list: closest hit of ray A (t in (0, 1e20), exclude exa) and occlusion of ray B
// (t in (0, tmax_b), exclude exb0 / exb1).  B is only tested against the leading occluder blocks of each group.  Inactive
// rays accept nothing (A: best_t = 0; B: tmax_b = 0).
__device__ __forceinline__ uint32_t trace_flat2_dual(const SceneView &sc, const TraceSmem &ts, bool active_a, f3 oa, f3 da, uint32_t exa, float &best_t,
                                                     bool active_b, f3 ob, f3 db, float tmax_b, uint32_t exb0, uint32_t exb1, bool &occluded) {
    best_t = active_a ? 1e20f : 0.0f;
    tmax_b = active_b ? tmax_b : 0.0f;
    occluded = false;
    uint32_t best_k = 0xffffffffu;
    const uint32_t np = sc.n_pair_blocks, nps = sc.n_shadow_pair_blocks, ne = np + sc.n_single_blocks, nss = np + sc.n_shadow_single_blocks;
#pragma unroll 1
    for (uint32_t b = 0; b < nps; ++b)
        flat2_block_dual<true, true>(ts.prims + b * (uint32_t)sizeof(PrimBlock2), b, oa, da, exa, best_t, best_k, ob, db, tmax_b, exb0, exb1, occluded);
#pragma unroll 1
    for (uint32_t b = nps; b < np; ++b)
        flat2_block_dual<true, false>(ts.prims + b * (uint32_t)sizeof(PrimBlock2), b, oa, da, exa, best_t, best_k, ob, db, tmax_b, exb0, exb1, occluded);
#pragma unroll 1
    for (uint32_t b = np; b < nss; ++b)
        flat2_block_dual<false, true>(ts.prims + b * (uint32_t)sizeof(PrimBlock2), b, oa, da, exa, best_t, best_k, ob, db, tmax_b, exb0, exb1, occluded);
#pragma unroll 1
    for (uint32_t b = nss; b < ne; ++b)
        flat2_block_dual<false, false>(ts.prims + b * (uint32_t)sizeof(PrimBlock2), b, oa, da, exa, best_t, best_k, ob, db, tmax_b, exb0, exb1, occluded);
    return best_k;
}

template <bool ANY_HIT, bool ALPHA>
__device__ __forceinline__ DevHit trace_flat2(const LaunchParams &P, const TraceSmem &ts, bool active, f3 o, f3 d, float t_min, float t_max, uint32_t ex0,
                                              uint32_t ex1) {
    float best_t;
    const uint32_t best_k = trace_flat2_core<ANY_HIT, ALPHA>(P.scene, ts, active, o, d, t_min, t_max, ex0, ex1, best_t);
    if (best_k == 0xffffffffu) return DevHit{0xffffffffu, 0u, 0.0f, 0.0f};
    if (ANY_HIT) return DevHit{0u, 0u, 0.0f, 0.0f};  // occluded: which triangle does not matter
    // (s, q) of the winner, same operations in the same order as the packed loop
    const PrimRec p = load_block_prim(ts.prims + (best_k >> 1) * (uint32_t)sizeof(PrimBlock2), best_k & 1u);
    float s, q;
    prim_coords(p, o, d, best_t, s, q);
    const PrimDecoded dec = prim_decode(p, s, q);
    return DevHit{dec.gid, dec.cls, dec.u, dec.v};
}

// Warp-aggregated append to TWO queues whose counters are an aligned 32-bit pair: one 64-bit atomic per
// warp reserves both ranges (half the traffic to the hot L2 lines of per-queue atomics).
__device__ __forceinline__ void warp_append2(uint32_t *counter_pair, bool pred0, bool pred1, uint32_t &slot0, uint32_t &slot1) {
    const uint32_t m0 = __ballot_sync(0xffffffffu, pred0), m1 = __ballot_sync(0xffffffffu, pred1);
    slot0 = slot1 = 0u;
    if ((m0 | m1) == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long base = 0ull;
    if (lane == 0u)
        base = atomicAdd(reinterpret_cast<unsigned long long *>(counter_pair), (unsigned long long)__popc(m0) | ((unsigned long long)__popc(m1) << 32));
    base = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t below = (1u << lane) - 1u;
    slot0 = (uint32_t)base + (uint32_t)__popc(m0 & below);
    slot1 = (uint32_t)(base >> 32) + (uint32_t)__popc(m1 & below);
}
// warp-aggregated queue append: one atomic per warp, slots ordered by lane
__device__ __forceinline__ uint32_t warp_append(uint32_t *counter, bool pred) {
    const uint32_t mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0u) return 0u;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0u;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_raygen(const __grid_constant__ LaunchParams P) {
    const uint32_t n = P.wave.n_pix * P.wave.n_spp;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        PathState ps = raygen_body(P.scene, P.tables, P.rp, P.wave, i);
        store_path(P.q[0], i, ps);
        st4(P.acc.l + i, f4{0.0f, 0.0f, 0.0f, 0.0f});
        st4(P.acc.b + i, f4{0.0f, 0.0f, 0.0f, 0.0f});
        P.acc.poison[i] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) P.counters[0] = n;
}

// Queued pipeline on a flat scene (alpha-tested materials, or engine option fused = 2): one launch per depth traces BOTH
// ray kinds, the closest-hit rays of the paths entering `depth` and the shadow rays the previous depth's shade stage
// produced.  Work is handed out in warp-sized tasks (closest-hit tasks first, the shorter any-hit tasks fill the tail).
template <bool ALPHA> __global__ void __launch_bounds__(kBlock) k_trace_flat(const __grid_constant__ LaunchParams P, uint32_t depth) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    const uint32_t n_cl = P.counters[depth * kCtrStride];
    const uint32_t n_sh = P.counters[depth * kCtrStride + 1u];
    const uint32_t t_cl = (n_cl + 31u) >> 5, t_sh = (n_sh + 31u) >> 5;
    const uint32_t n_tasks = t_cl + t_sh;
    if (blockIdx.x * kWarpsPerBlock >= n_tasks) return;  // whole CTA has no work: skip the staging too
    const TraceSmem ts = stage_scene(P, smem, &bar);
    const PathQueue &q = P.q[depth & 1u];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t *ctr = P.counters + depth * kCtrStride;
    const bool miss_work = depth != 0u && (P.rp.debug_depth < 0 || depth == (uint32_t)P.rp.debug_depth);
    for (uint32_t task = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < n_tasks; task += gridDim.x * kWarpsPerBlock) {
        if (task < t_cl) {
            const uint32_t i = task * 32u + lane;
            const bool active = i < n_cl;
            f4 a{0.0f, 0.0f, 0.0f, 1.0f}, b{0.0f, 0.0f, 0.0f, 0.0f};
            if (active) {
                a = ldq(q.a + i);
                b = ldq(q.b + i);
            }
            const uint32_t path_id = f2u(b.w);
            const DevHit h = trace_flat2<false, ALPHA>(P, ts, active, mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), 0.0f, 1e20f, f2u(b.z), 0xffffffffu);
            if (active) stq(P.hits.h + i, f4{u2f(h.gid), h.u, h.v, 0.0f});
            const bool hit = active && h.gid != 0xffffffffu;
            const uint32_t cls = P.rp.force_diffuse ? (uint32_t)CLS_LAMBERT : h.cls;
            uint32_t s0, s1;
            warp_append2(ctr + 2u, hit && cls == CLS_LAMBERT, hit && cls == CLS_CONDUCTOR, s0, s1);
            const uint32_t s2 = warp_append(ctr + 4u, hit && cls == CLS_GENERAL);
            if (hit) P.cls.idx[cls][cls == CLS_LAMBERT ? s0 : (cls == CLS_CONDUCTOR ? s1 : s2)] = make_uint2(i, path_id);
            if (active && !hit && miss_work) {
                const f4 c = ldq(q.c + i);
                miss_body(P.rp, depth, mk3(c.x, c.y, c.z), path_id, P.acc);
            }
        } else {
            const uint32_t i = (task - t_cl) * 32u + lane;
            const bool active = i < n_sh;
            f4 a{0.0f, 0.0f, 0.0f, 0.0f}, b{1.0f, 0.0f, 0.0f, 0.0f};
            uint32_t ex1 = 0xffffffffu;
            if (active) {
                a = ldq(P.shadow.a + i);
                b = ldq(P.shadow.b + i);
                ex1 = P.shadow.ex1[i];
            }
            const DevHit h = trace_flat2<true, ALPHA>(P, ts, active, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), 0.0f, a.w, f2u(b.w), ex1);
            if (active) {
                const f4 c = ldq(P.shadow.c + i);
                ShadowItem it;
                it.contrib = mk3(c.x, c.y, c.z);
                it.path_id = f2u(c.w);
                shadow_resolve(P.acc, it, h.gid != 0xffffffffu, depth);  // produced at depth - 1: depth1 = depth
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// BVH scenes: persistent warps with dynamic ray fetch and majority-vote rounds
// ------------------------------------------------------------------------------------------------
// Incoherent rays leave a BVH at very different times; with one fixed ray per lane the warp idles until its slowest
// lane is done (ncu on the 8.5 K-triangle test scene: 8.5 of 32 lanes active).  Here a warp keeps per-lane traversal
// state and advances it in warp-uniform rounds (below); whenever fewer than AKR_REFILL_BELOW lanes still carry a live
// ray it retires the finished ones (hit record, class binning / shadow resolve, warp-aggregated) and hands the free
// lanes new rays from a global counter (counter block words [5] closest-hit, [6] shadow).  Measured on the clutter scene
// (trace stage of one 64-spp pass): while-while segments of 12 visits, refill below 24 lanes: 72.8 ms; vote rounds with
// refill below 24 / 20 / 16 / 12 / 8 lanes: 75.7 / 70.3 / 66.9 / 65.2 / 66.1 ms (a refill costs a retire pass with
// atomics and queue stores: with uniform rounds it pays to run the warp emptier); biasing the vote 2:1 either way or
// testing one primitive per leaf round: 65.8-72.4 ms.
#ifndef AKR_REFILL_BELOW
#define AKR_REFILL_BELOW 12
#endif
#ifndef AKR_SEGMENT_STEPS
#define AKR_SEGMENT_STEPS 32
#endif

template <bool ANY_HIT, bool SMEM_ALL, bool ALPHA>
__device__ __forceinline__ void bvh_phase(const LaunchParams &P, const TraceSmem &ts, uint32_t depth, uint32_t n_items, uint32_t *fetch) {
    if (n_items == 0u) return;
    const SceneView &sc = P.scene;
    const PathQueue &q = P.q[depth & 1u];
    uint32_t *ctr = P.counters + depth * kCtrStride;
    const uint32_t lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    const bool miss_work = depth != 0u && (P.rp.debug_depth < 0 || depth == (uint32_t)P.rp.debug_depth);
    const uint32_t n_fast = P.scene_smem_nodes;
    constexpr uint32_t kStackStride = kBlock * 4u;
    // per-lane ray + traversal state
    bool have = false, done = false;
    uint32_t slot = 0u, path_id = 0u, ex0 = 0xffffffffu, ex1 = 0xffffffffu;
    f3 o = splat3(0.0f), d = mk3(1.0f, 0.0f, 0.0f), inv_d = splat3(0.0f);
    PrimHit best{0.0f, 0.0f, 0.0f, 0xffffffffu};
    int32_t node = 0;
    uint32_t sp = ts.stack;
    bool exhausted = false;  // warp-uniform: the global counter ran past the queue
    while (true) {
        const uint32_t busy = __ballot_sync(0xffffffffu, have && !done);
        if (busy == 0u || (!exhausted && __popc(busy) < AKR_REFILL_BELOW)) {
            // ---- retire the finished lanes (all lanes take part in the ballots) ----
            const bool fin = have && done;
            if (!ANY_HIT) {
                DevHit h{0xffffffffu, 0u, 0.0f, 0.0f};
                if (fin && best.k != 0xffffffffu) {
                    const PrimDecoded dec = prim_decode(load_prim<SMEM_ALL>(ts.prims, sc.prims, best.k), best.s, best.q);
                    h = DevHit{dec.gid, dec.cls, dec.u, dec.v};
                }
                if (fin) stq(P.hits.h + slot, f4{u2f(h.gid), h.u, h.v, 0.0f});
                const bool hit = fin && h.gid != 0xffffffffu;
                const uint32_t cls = P.rp.force_diffuse ? (uint32_t)CLS_LAMBERT : h.cls;
                uint32_t s0, s1;
                warp_append2(ctr + 2u, hit && cls == CLS_LAMBERT, hit && cls == CLS_CONDUCTOR, s0, s1);
                const uint32_t s2 = warp_append(ctr + 4u, hit && cls == CLS_GENERAL);
                if (hit) P.cls.idx[cls][cls == CLS_LAMBERT ? s0 : (cls == CLS_CONDUCTOR ? s1 : s2)] = make_uint2(slot, path_id);
                if (fin && !hit && miss_work) {
                    const f4 c = ldq(q.c + slot);
                    miss_body(P.rp, depth, mk3(c.x, c.y, c.z), path_id, P.acc);
                }
            } else if (fin) {
                const f4 c = ldq(P.shadow.c + slot);
                ShadowItem it;
                it.contrib = mk3(c.x, c.y, c.z);
                it.path_id = f2u(c.w);
                shadow_resolve(P.acc, it, best.k != 0xffffffffu, depth);  // produced at depth - 1: depth1 = depth
            }
            have = have && !done;
            // ---- hand new rays to the free lanes ----
            if (!exhausted) {
                const bool need = !have;
                const uint32_t m = __ballot_sync(0xffffffffu, need);
                uint32_t base = 0u;
                if (lane == 0u) base = atomicAdd(fetch, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t idx = base + (uint32_t)__popc(m & below);
                if (need && idx < n_items) {
                    float t_max;
                    if (!ANY_HIT) {
                        const f4 a = ldq(q.a + idx), b = ldq(q.b + idx);
                        o = mk3(a.x, a.y, a.z);
                        d = mk3(a.w, b.x, b.y);
                        ex0 = f2u(b.z);
                        ex1 = 0xffffffffu;
                        path_id = f2u(b.w);
                        t_max = 1e20f;
                    } else {
                        const f4 a = ldq(P.shadow.a + idx), b = ldq(P.shadow.b + idx);
                        o = mk3(a.x, a.y, a.z);
                        d = mk3(b.x, b.y, b.z);
                        t_max = a.w;
                        ex0 = f2u(b.w);
                        ex1 = P.shadow.ex1[idx];
                    }
                    inv_d = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    best = PrimHit{t_max, 0.0f, 0.0f, 0xffffffffu};
                    slot = idx;
                    node = 0;  // the root is always an inner node
                    sp = ts.stack;
                    have = true;
                    done = false;
                }
                exhausted = base + (uint32_t)__popc(m) >= n_items;
            }
            if (__ballot_sync(0xffffffffu, have) == 0u) break;
        }
        // ---- rounds: the warp does whichever of the two activities more of its lanes are waiting for ----
        // A lane alternates between runs of inner-node visits and leaves.  In a while-while loop the lanes that reach a leaf
        // early idle until the last one has (ncu r02q: 12.7 of 32 lanes in the box tests, 16.4 in the primitive tests).  Here
        // every round is warp-uniform: one node visit for the lanes at inner nodes, or one leaf for the lanes at leaves,
        // decided by majority; the minority waits one round and grows meanwhile.
#pragma unroll 1
        for (int round = 0; round < AKR_SEGMENT_STEPS; ++round) {
            const bool live = have && !done;
            const uint32_t mi = __ballot_sync(0xffffffffu, live && node >= 0), ml = __ballot_sync(0xffffffffu, live && node < 0);
            if ((mi | ml) == 0u) break;
            if (!exhausted && __popc(mi | ml) < AKR_REFILL_BELOW) break;
            if (__popc(mi) >= __popc(ml)) {
                if (live && node >= 0) {
                    BvhNode n;
                    if (SMEM_ALL || (uint32_t)node < n_fast) n = load_node<true>(ts.nodes, nullptr, (uint32_t)node);
                    else n = load_node<false>(0u, sc.nodes, (uint32_t)node);
                    float tn0, tn1;
                    const bool h0 = box_test(n.lo0, n.hi0, o, inv_d, 0.0f, best.t, tn0);
                    const bool h1 = box_test(n.lo1, n.hi1, o, inv_d, 0.0f, best.t, tn1);
                    const int32_t c0 = n.c0, c1 = n.c1;
                    if (h0 && h1) {
                        const bool swap = tn1 < tn0;
                        sts32(sp, swap ? c0 : c1);
                        sp += kStackStride;
                        node = swap ? c1 : c0;
                    } else if (h0) {
                        node = c0;
                    } else if (h1) {
                        node = c1;
                    } else if (sp == ts.stack) {
                        done = true;
                    } else {
                        sp -= kStackStride;
                        node = lds32(sp);
                    }
                }
            } else if (live && node < 0) {
                const uint32_t leaf = (uint32_t)(~node);
                const uint32_t first = leaf >> 3, count = leaf & 7u;
                for (uint32_t k = 0; k < count; ++k)
                    prim_test<ALPHA>(sc, load_prim<SMEM_ALL>(ts.prims, sc.prims, first + k), first + k, o, d, 0.0f, ex0, ex1, best);
                if ((ANY_HIT && best.k != 0xffffffffu) || sp == ts.stack) {
                    done = true;
                } else {
                    sp -= kStackStride;
                    node = lds32(sp);
                }
            }
        }
    }
}

template <bool SMEM_ALL, bool ALPHA> __global__ void __launch_bounds__(kBlock) k_trace_bvh(const __grid_constant__ LaunchParams P, uint32_t depth) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    uint32_t *ctr = P.counters + depth * kCtrStride;
    const uint32_t n_cl = ctr[0];
    const uint32_t n_sh = ctr[1];
    if (blockIdx.x * kBlock >= n_cl + n_sh) return;  // more CTAs than rays: skip the staging too
    const TraceSmem ts = stage_scene(P, smem, &bar);
    bvh_phase<false, SMEM_ALL, ALPHA>(P, ts, depth, n_cl, ctr + 5u);
    bvh_phase<true, SMEM_ALL, ALPHA>(P, ts, depth, n_sh, ctr + 6u);
}

// resident CTAs per SM the shade / bounce kernels are compiled for (register budget = 65536 / (256 * n)); tuned on B200
#ifndef AKR_SHADE_MINB_LAMBERT
#define AKR_SHADE_MINB_LAMBERT 4
#endif
#ifndef AKR_SHADE_MINB_CONDUCTOR
#define AKR_SHADE_MINB_CONDUCTOR 2
#endif
#ifndef AKR_SHADE_BLOCK
#define AKR_SHADE_BLOCK 256
#endif
constexpr int kShadeBlock = AKR_SHADE_BLOCK;
constexpr int kShadeWarps = kShadeBlock / 32;
#ifndef AKR_SHADE_MINB_GENERAL
#define AKR_SHADE_MINB_GENERAL 2  // 128 registers, 2 CTAs per SM: measured 127 vs 166 ms per pass (1 CTA, 206 registers) on the all-Principled box
#endif
#ifndef AKR_GENERAL_BLOCK
#define AKR_GENERAL_BLOCK 256
#endif
constexpr int kGeneralBlock = AKR_GENERAL_BLOCK;  // block size of the kernels that carry the full Principled tree
constexpr int shade_block_of(int cls) { return cls == CLS_GENERAL ? kGeneralBlock : kShadeBlock; }
template <int CLS> struct ShadeLaunch {
    static constexpr int kThreads = shade_block_of(CLS);
    static constexpr int kMinBlocks = CLS == CLS_LAMBERT ? AKR_SHADE_MINB_LAMBERT : (CLS == CLS_CONDUCTOR ? AKR_SHADE_MINB_CONDUCTOR : AKR_SHADE_MINB_GENERAL);
};

// Queued pipeline: one shade kernel per material class over the (slot, path id) list the trace stage binned.
template <int CLS> __global__ void __launch_bounds__(ShadeLaunch<CLS>::kThreads, ShadeLaunch<CLS>::kMinBlocks) k_shade(const __grid_constant__ LaunchParams P, uint32_t depth) {
    uint32_t *ctr = P.counters + depth * kCtrStride;
    const uint32_t n = ctr[2u + CLS];
    const uint2 *slots = P.cls.idx[CLS];
    const PathQueue &qin = P.q[depth & 1u];
    const PathQueue &qout = P.q[(depth + 1u) & 1u];
    uint32_t *out_pair = ctr + kCtrStride;  // [0] next-depth paths, [1] shadow rays: reserved together
    auto emit = [&](const ShadeOut &o) {  // warp-collective: append the continuation and the shadow ray
        uint32_t ns, ss;
        warp_append2(out_pair, o.has_next, o.has_shadow, ns, ss);
        if (o.has_next) store_path(qout, ns, o.next);
        if (o.has_shadow) {
            const ShadowQueue &s = P.shadow;
            stq(s.a + ss, f4{o.shadow.o.x, o.shadow.o.y, o.shadow.o.z, o.shadow.t_max});
            stq(s.b + ss, f4{o.shadow.d.x, o.shadow.d.y, o.shadow.d.z, u2f(o.shadow.ex0)});
            stq(s.c + ss, f4{o.shadow.contrib.x, o.shadow.contrib.y, o.shadow.contrib.z, u2f(o.shadow.path_id)});
            s.ex1[ss] = o.shadow.ex1;
        }
    };
    auto shade_one = [&](bool active, uint2 cur) {
        ShadeOut o;
        o.has_shadow = false;
        o.has_next = false;
        if (active) {
            const uint32_t i = cur.x;
            const f4 hr = ldq(P.hits.h + i);
            HitRec h{f2u(hr.x), hr.y, hr.z};
            PathState ps = load_path(qin, i);
            ps.path_id = cur.y;  // same value, but already in a register: the sampler loads do not wait for the record
            o = shade_body<CLS>(P.scene, P.corners, P.tables, P.rp, P.wave, depth, ps, h, P.acc);
            if (depth == 0u && P.first_hits && ps.path_id < P.wave.n_pix && P.wave.s0 == 0u) {
                uint32_t pix = P.wave.pix0 + ps.path_id;
                P.first_hits[2u * pix + 0u] = P.scene.shade[h.gid].inst;
                P.first_hits[2u * pix + 1u] = P.scene.shade[h.gid].prim;
            }
        }
        emit(o);
    };
    const uint32_t stride = gridDim.x * blockDim.x;
    // warp-uniform trip count so that every lane takes part in the ballots; the (slot, path_id) entry of the
    // next trip is fetched one trip ahead so that its latency is off the dependent chain
    // general class: entry k is slots[sort_order[k]] — the class list ordered by material sort key (k_sort_hist / k_sort_scatter)
    auto entry = [&](uint32_t k) { return CLS == CLS_GENERAL ? slots[P.sort_order[k]] : slots[k]; };
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint2 ent = make_uint2(k, 0u);
    if (k < n) ent = entry(k);
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride, k += stride) {
        const bool active = k < n;
        const uint2 cur = ent;
        if (k + stride < n) ent = entry(k + stride);
        shade_one(active, cur);
    }
}

// ------------------------------------------------------------------------------------------------
// fused pipeline (flat scenes): k_raygen_fused, k_bounce<CLS>
// ------------------------------------------------------------------------------------------------
// Tracer of akr_path.cuh's fused bodies: both queries walk the primitive list staged in shared memory with the
// packed-FP32 loop; every lane of the warp calls them together.
struct DevTracer {
    const SceneView &sc;
    TraceSmem ts;
    // shadow ray + continuation ray of the same path in one walk over the staged list
    __device__ __forceinline__ TraceHit trace2(bool has_shadow, f3 so, f3 sd, float st_max, uint32_t sex0, uint32_t sex1, bool &occluded, bool has_next, f3 o, f3 d,
                                               uint32_t ex0) const {
        float best_t;
        const uint32_t best_k = trace_flat2_dual(sc, ts, has_next, o, d, ex0, best_t, has_shadow, so, sd, st_max, sex0, sex1, occluded);
        if (best_k == 0xffffffffu) return TraceHit{0xffffffffu, 0u, 0u, 0.0f, 0.0f};
        const PrimRec p = load_block_prim(ts.prims + (best_k >> 1) * (uint32_t)sizeof(PrimBlock2), best_k & 1u);
        float s, q;
        prim_coords(p, o, d, best_t, s, q);
        const PrimDecoded dec = prim_decode(p, s, q);
        return TraceHit{dec.gid, dec.cls, dec.light, dec.u, dec.v};
    }
    __device__ __forceinline__ TraceHit closest(bool active, f3 o, f3 d, uint32_t ex0) const {
        float best_t;
        const uint32_t best_k = trace_flat2_core<false, false>(sc, ts, active, o, d, 0.0f, 1e20f, ex0, 0xffffffffu, best_t);
        if (best_k == 0xffffffffu) return TraceHit{0xffffffffu, 0u, 0u, 0.0f, 0.0f};
        // (s, q) of the winner, same operations in the same order as the packed loop
        const PrimRec p = load_block_prim(ts.prims + (best_k >> 1) * (uint32_t)sizeof(PrimBlock2), best_k & 1u);
        float s, q;
        prim_coords(p, o, d, best_t, s, q);
        const PrimDecoded dec = prim_decode(p, s, q);
        return TraceHit{dec.gid, dec.cls, dec.light, dec.u, dec.v};
    }
};

__device__ __forceinline__ void store_rec(const RecQueue &q, uint32_t i, const BounceRec &r) {
    stq(q.r[0] + i, f4{r.d.x, r.d.y, r.d.z, u2f(r.gid)});
    stq(q.r[1] + i, f4{r.u, r.v, u2f(r.path_id), u2f(r.pxpy)});
    stq(q.r[2] + i, f4{r.beta.x, r.beta.y, r.beta.z, u2f(r.sample_index)});
    stq(q.r[3] + i, f4{r.L.x, r.L.y, r.L.z, 0.0f});
}
// appends a surviving path to the queue of the shade class of the hit it goes to (depth d1)
__device__ __forceinline__ void append_next(const LaunchParams &P, uint32_t d1, const BounceOut &r) {
    uint32_t *ctr = P.counters + d1 * kCtrStride;
    uint32_t s0, s1;
    warp_append2(ctr + 2u, r.cont && r.cls == CLS_LAMBERT, r.cont && r.cls == CLS_CONDUCTOR, s0, s1);
    const uint32_t s2 = warp_append(ctr + 4u, r.cont && r.cls == CLS_GENERAL);
    if (r.cont) store_rec(P.cq[d1 & 1u][r.cls], r.cls == CLS_LAMBERT ? s0 : (r.cls == CLS_CONDUCTOR ? s1 : s2), r.next);
}

__global__ void __launch_bounds__(kBlock) k_raygen_fused(const __grid_constant__ LaunchParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    const uint32_t n = P.wave.n_pix * P.wave.n_spp;
    if (blockIdx.x * blockDim.x >= n) return;
    const DevTracer tr{P.scene, stage_scene(P, smem, &bar)};
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride, i += stride) {
        const bool active = i < n;
        const BounceOut r = raygen_fused(P.scene, P.corners, P.tables, P.rp, P.wave, active, i, tr, P.acc);
        if (active && P.first_hits && i < P.wave.n_pix && P.wave.s0 == 0u) {
            const uint32_t pix = P.wave.pix0 + i, gid = r.next.gid;
            P.first_hits[2u * pix + 0u] = gid == 0xffffffffu ? 0xffffffffu : P.scene.shade[gid].inst;
            P.first_hits[2u * pix + 1u] = gid == 0xffffffffu ? 0xffffffffu : P.scene.shade[gid].prim;
        }
        append_next(P, 0u, r);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) P.counters[0] = n;  // camera rays traced
}

// Streaming loads of the record queues bypass L2 residency (evict-first): written once, read once, gigabytes apart.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst), "l"(gsrc),
                 "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
constexpr uint32_t kTileBytes = 4u * 32u * 16u;  // one warp tile: 32 records x 4 x 16 B

// One bounce of every path whose current hit is of shade class CLS.  Warp w of the grid owns record tiles w, w + W, ...
// (32 records each); lane 0 starts the TMA bulk copies of the NEXT tile into the warp's other shared-memory buffer
// before the warp waits for the current one, so the queue reads never sit on the dependent chain.  `it` counts the tiles
// this warp has consumed in this launch (it selects the buffer and the mbarrier phase) and carries over from class to class.
template <int CLS, int WARPS>
__device__ __forceinline__ void bounce_phase(const LaunchParams &P, uint32_t depth, const DevTracer &tr, uint32_t tiles, uint64_t *tile_bar2, uint32_t &it) {
    const uint32_t n = P.counters[depth * kCtrStride + 2u + CLS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t bar0 = smem_u32(tile_bar2);
    const RecQueue &qin = P.cq[depth & 1u][CLS];
    const uint64_t pol = l2_evict_first_policy();
    const uint32_t n_tiles = (n + 31u) >> 5, wstride = gridDim.x * (uint32_t)WARPS;
    auto issue = [&](uint32_t tile, uint32_t buf) {
        const uint32_t first = tile * 32u;
        const uint32_t bytes = min(32u, n - first) * 16u;
        const uint32_t b = bar0 + buf * 8u, dst = tiles + buf * kTileBytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4u * bytes) : "memory");
#pragma unroll
        for (uint32_t j = 0; j < 4u; ++j) tma_bulk_g2s_hint(dst + j * 512u, qin.r[j] + first, bytes, b, pol);
    };
    uint32_t tile = blockIdx.x * (uint32_t)WARPS + warp;
    if (tile < n_tiles && lane == 0u) issue(tile, it & 1u);
    uint32_t n_traced = 0u, n_shadow = 0u;  // warp-uniform
    for (; tile < n_tiles; tile += wstride, ++it) {
        const uint32_t buf = it & 1u;
        if (tile + wstride < n_tiles && lane == 0u) issue(tile + wstride, buf ^ 1u);
        mbar_wait(tile_bar2 + buf, (it >> 1) & 1u);
        const bool active = tile * 32u + lane < n;
        const uint32_t src = tiles + buf * kTileBytes + lane * 16u;
        const float4 r0 = lds128(src), r1 = lds128(src + 512u), r2 = lds128(src + 1024u), r3 = lds128(src + 1536u);
        __syncwarp();  // every lane has its record in registers before lane 0 may refill this buffer (two trips from now)
        BounceRec in;
        in.d = mk3(r0.x, r0.y, r0.z);
        in.gid = __float_as_uint(r0.w);
        in.u = r1.x;
        in.v = r1.y;
        in.path_id = __float_as_uint(r1.z);
        in.pxpy = __float_as_uint(r1.w);
        in.beta = mk3(r2.x, r2.y, r2.z);
        in.sample_index = __float_as_uint(r2.w);
        in.L = mk3(r3.x, r3.y, r3.z);
        const BounceOut r = bounce_fused<CLS>(P.scene, P.corners, P.tables, P.rp, P.wave, depth, active, in, tr, P.acc);
        n_traced += (uint32_t)__popc(__ballot_sync(0xffffffffu, r.traced));
        n_shadow += (uint32_t)__popc(__ballot_sync(0xffffffffu, r.shadow));
        append_next(P, depth + 1u, r);
    }
    // statistics: [0] continuation rays traced (= segments of depth + 1), [1] shadow rays of this depth
    if (lane == 0u && (n_traced | n_shadow))
        atomicAdd(reinterpret_cast<unsigned long long *>(P.counters + (depth + 1u) * kCtrStride), (unsigned long long)n_traced | ((unsigned long long)n_shadow << 32));
}
// ---- general class: the records of a depth ordered by material signature -------------------------------------------------
// The general class holds every closure tree there is (Principled lobe sets, glass, texture-driven programs).  In queue
// order — the order in which paths happened to land on such materials — a warp evaluates the UNION of its lanes' trees
// (ncu on the all-Principled box: 14.9 of 32 lanes active) and the warps of an SM run through different regions of an
// 8 K-instruction body at the same time (47 % of the stalls were instruction fetch).  So before a depth is shaded, two
// small kernels order its records by the 5-bit sort key scene_build.cpp stores in TriShade.flags (same closure type, lobe
// set, normal-map frame, shader kind of a texture-driven material => same key; numbered by increasing cost): a counting
// sort of the record INDICES (histogram, then a scatter behind per-key cursors) into `sort_order`.  The shade / bounce
// kernel takes 32 consecutive entries per warp and gathers their records: full lanes, and the whole machine works on one
// signature for long stretches.  The result does not depend on the order (one path per record, own accumulator slots).
// Measured (general kernel of one 64-spp pass, incl. the two sort kernels): all-Principled box 127 ms unsorted, 110 ms with
// a per-CTA sort of 512-record tiles in shared memory, 60 ms with this global order (at equal code otherwise 82 -> 60).
template <bool QUEUED> __device__ __forceinline__ uint32_t general_key_of(const LaunchParams &P, uint32_t depth, uint32_t i) {
    uint32_t gid;
    if (QUEUED) gid = __float_as_uint(reinterpret_cast<const float *>(P.hits.h + P.cls.idx[CLS_GENERAL][i].x)[0]);  // the hit of the entry's slot
    else gid = __float_as_uint(reinterpret_cast<const float *>(P.cq[depth & 1u][CLS_GENERAL].r[0] + i)[3]);     // the record's own hit
    return (P.scene.shade[gid].flags >> TRI_SORT_KEY_SHIFT) & TRI_SORT_KEY_MASK;
}
template <bool QUEUED> __global__ void __launch_bounds__(256) k_sort_hist(const __grid_constant__ LaunchParams P, uint32_t depth) {
    __shared__ uint32_t h[32];
    const uint32_t n = P.counters[depth * kCtrStride + 2u + CLS_GENERAL];
    if (blockIdx.x * blockDim.x >= n) return;
    if (threadIdx.x < 32) h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        const uint32_t key = i < n ? general_key_of<QUEUED>(P, depth, i) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (i < n && (int)lane == __ffs(peers) - 1) atomicAdd(&h[key], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 32 && h[threadIdx.x]) atomicAdd(P.sort_hist + depth * 64u + threadIdx.x, h[threadIdx.x]);
}
template <bool QUEUED> __global__ void __launch_bounds__(256) k_sort_scatter(const __grid_constant__ LaunchParams P, uint32_t depth) {
    __shared__ uint32_t first[32], h[32], base[32];
    const uint32_t n = P.counters[depth * kCtrStride + 2u + CLS_GENERAL];
    if (blockIdx.x * blockDim.x >= n) return;
    uint32_t *hist = P.sort_hist + depth * 64u, *cursor = hist + 32u;
    const uint32_t lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    if (threadIdx.x < 32) {  // exclusive scan of the key counts
        const uint32_t v = hist[lane];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        first[lane] = incl - v;
    }
    // per chunk of 256 records: ranks within the CTA from shared-memory atomics (one per warp and key), then ONE global
    // atomic per key reserves the CTA's range behind that key's cursor (per-warp global atomics on the handful of hot
    // cursors made this kernel 0.74 ms per depth, ncu r02zz)
    for (uint32_t chunk = blockIdx.x * blockDim.x; chunk < n; chunk += gridDim.x * blockDim.x) {
        if (threadIdx.x < 32) h[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t i = chunk + threadIdx.x;
        const uint32_t key = i < n ? general_key_of<QUEUED>(P, depth, i) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        uint32_t r = 0u;
        if (i < n && (int)lane == leader) r = atomicAdd(&h[key], (uint32_t)__popc(peers));
        r = __shfl_sync(0xffffffffu, r, leader) + (uint32_t)__popc(peers & below);
        __syncthreads();
        if (threadIdx.x < 32 && h[threadIdx.x]) base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], h[threadIdx.x]);
        __syncthreads();
        if (i < n) P.sort_order[first[key] + base[key] + r] = i;
    }
}
__device__ __forceinline__ void bounce_phase_ordered(const LaunchParams &P, uint32_t depth, const DevTracer &tr) {
    const uint32_t n = P.counters[depth * kCtrStride + 2u + CLS_GENERAL];
    const uint32_t lane = threadIdx.x & 31u, n_tiles = (n + 31u) >> 5, wstride = gridDim.x * (blockDim.x >> 5);
    const RecQueue &qin = P.cq[depth & 1u][CLS_GENERAL];
    uint32_t n_traced = 0u, n_shadow = 0u;
    for (uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles; tile += wstride) {
        const uint32_t k = tile * 32u + lane;
        const bool active = k < n;
        const uint32_t i = active ? P.sort_order[k] : 0u;
        const f4 r0 = ldq(qin.r[0] + i), r1 = ldq(qin.r[1] + i), r2 = ldq(qin.r[2] + i), r3 = ldq(qin.r[3] + i);
        BounceRec in;
        in.d = mk3(r0.x, r0.y, r0.z);
        in.gid = f2u(r0.w);
        in.u = r1.x;
        in.v = r1.y;
        in.path_id = f2u(r1.z);
        in.pxpy = f2u(r1.w);
        in.beta = mk3(r2.x, r2.y, r2.z);
        in.sample_index = f2u(r2.w);
        in.L = mk3(r3.x, r3.y, r3.z);
        const BounceOut r = bounce_fused<CLS_GENERAL>(P.scene, P.corners, P.tables, P.rp, P.wave, depth, active, in, tr, P.acc);
        n_traced += (uint32_t)__popc(__ballot_sync(0xffffffffu, r.traced));
        n_shadow += (uint32_t)__popc(__ballot_sync(0xffffffffu, r.shadow));
        append_next(P, depth + 1u, r);
    }
    if (lane == 0u && (n_traced | n_shadow))
        atomicAdd(reinterpret_cast<unsigned long long *>(P.counters + (depth + 1u) * kCtrStride), (unsigned long long)n_traced | ((unsigned long long)n_shadow << 32));
}

// MASK = the shade classes this launch serves, one after the other in every CTA (bit c = class c).  Measured on B200:
// serving Lambert + conductor in ONE launch per depth (half the launches, CTAs that run out of Lambert records start on
// conductor records) is SLOWER than one launch per class — 31.0 vs 29.1 ms per pass on one GPU, 12.8 vs 13.8 G samples/s
// on eight: the merged kernel's larger code and common register allocation cost more than the saved tails — so the
// engine launches single-class masks; the template keeps the general form.
constexpr size_t kGeneralTileSmem = 0;  // the general bounce kernel gathers its records through sort_order: no tile buffers
template <uint32_t MASK> struct BounceLaunch {
    static constexpr int kThreads = (MASK & (1u << CLS_GENERAL)) ? kGeneralBlock : kShadeBlock;
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kMinBlocks = (MASK & (1u << CLS_GENERAL)) ? AKR_SHADE_MINB_GENERAL : ((MASK & (1u << CLS_CONDUCTOR)) ? AKR_SHADE_MINB_CONDUCTOR : AKR_SHADE_MINB_LAMBERT);
};
template <uint32_t MASK> __global__ void __launch_bounds__(BounceLaunch<MASK>::kThreads, BounceLaunch<MASK>::kMinBlocks) k_bounce(const __grid_constant__ LaunchParams P, uint32_t depth) {
    constexpr int kWarps = BounceLaunch<MASK>::kWarps;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t tile_bar[kWarps][2];
    const uint32_t *ctr = P.counters + depth * kCtrStride + 2u;
    uint32_t n_max = 0u;
#pragma unroll
    for (uint32_t c = 0; c < (uint32_t)CLS_COUNT; ++c)
        if (MASK & (1u << c)) n_max = max(n_max, ctr[c]);
    if (blockIdx.x * blockDim.x >= n_max) return;  // whole CTA has no work in any class: skip the staging too
    const uint32_t warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int w = 0; w < kWarps; ++w) {
            mbar_init(&tile_bar[w][0], 1);
            mbar_init(&tile_bar[w][1], 1);
        }
    }
    const DevTracer tr{P.scene, stage_scene(P, smem, &bar)};  // (inits `bar`, fences the barrier inits, __syncthreads)
    const uint32_t scene_bytes = (P.scene.n_pair_blocks + P.scene.n_single_blocks + P.scene.n_occ_pair_blocks + P.scene.n_occ_single_blocks) * (uint32_t)sizeof(PrimBlock2);
    if (MASK == (1u << CLS_GENERAL)) {  // the general class alone: records in the order k_sort_hist / k_sort_scatter left in sort_order
        bounce_phase_ordered(P, depth, tr);
        return;
    }
    const uint32_t tiles = smem_u32(smem) + ((scene_bytes + 127u) & ~127u) + warp * 2u * kTileBytes;
    uint32_t it = 0u;
    if (MASK & (1u << CLS_LAMBERT)) bounce_phase<CLS_LAMBERT, kWarps>(P, depth, tr, tiles, &tile_bar[warp][0], it);
    if (MASK & (1u << CLS_CONDUCTOR)) bounce_phase<CLS_CONDUCTOR, kWarps>(P, depth, tr, tiles, &tile_bar[warp][0], it);
    if (MASK & (1u << CLS_GENERAL)) bounce_phase<CLS_GENERAL, kWarps>(P, depth, tr, tiles, &tile_bar[warp][0], it);
}

// The `aov` method (aov.rs:96-155) after raygen + one trace stage: the first-hit quantity of every camera sample goes to
// both accumulators (radiance == base_replay_throughput, so k_accumulate's indirect clamp is a no-op and its NaN removal
// is Film::add_sample's).  Misses keep the zeros raygen wrote.
__global__ void __launch_bounds__(kBlock) k_aov(const __grid_constant__ LaunchParams P) {
    const uint32_t n = P.wave.n_pix * P.wave.n_spp;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const f4 hr = ldq(P.hits.h + i);
        if (f2u(hr.x) == 0xffffffffu) continue;
        const f4 a = ldq(P.q[0].a + i), b = ldq(P.q[0].b + i);
        const f3 c = aov_body(P.scene, P.corners, P.tables, P.rp, P.wave, P.aov, P.aov_remap != 0u, f2u(b.w), mk3(a.w, b.x, b.y), HitRec{f2u(hr.x), hr.y, hr.z});
        st4(P.acc.l + f2u(b.w), f4{c.x, c.y, c.z, 0.0f});
        st4(P.acc.b + f2u(b.w), f4{c.x, c.y, c.z, 0.0f});
    }
}

__global__ void __launch_bounds__(kBlock) k_accumulate(const __grid_constant__ LaunchParams P) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.wave.n_pix; p += stride) accumulate_body(P.acc, P.wave, p, P.film, P.n_film_pixels);
}

// folds the per-depth counters of one wave into the 64-bit totals and clears them for the next wave
__global__ void k_fold_counters(uint32_t *counters, unsigned long long *totals, uint32_t n_depth, uint32_t *sort_hist) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long seg = 0, sh = 0, hits = 0;
        for (uint32_t d = 0; d < n_depth; ++d) {
            seg += counters[d * kCtrStride];
            sh += counters[d * kCtrStride + 1u];
            hits += (unsigned long long)counters[d * kCtrStride + 2u] + counters[d * kCtrStride + 3u] + counters[d * kCtrStride + 4u];
        }
        totals[0] += seg;
        totals[1] += sh;
        totals[2] += hits;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_depth * kCtrStride; i += blockDim.x) counters[i] = 0u;
    for (uint32_t i = threadIdx.x; i < n_depth * 64u; i += blockDim.x) sort_hist[i] = 0u;
}

// Film::copy_to_rgba_image(hdr = true) (film.rs:120-148), splat_scale = 1
__global__ void k_resolve_film(const float *film, uint32_t n, float *out, int rgba) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float w = film[6u * n + i];
        float d = (w == 0.0f) ? 1.0f : w;
        float r = film[i * 3u + 0u] / d + film[3u * n + i * 3u + 0u] * 1.0f;
        float g = film[i * 3u + 1u] / d + film[3u * n + i * 3u + 1u] * 1.0f;
        float b = film[i * 3u + 2u] / d + film[3u * n + i * 3u + 2u] * 1.0f;
        if (rgba) {
            out[i * 4u + 0u] = r; out[i * 4u + 1u] = g; out[i * 4u + 2u] = b; out[i * 4u + 3u] = 1.0f;
        } else {
            out[i * 3u + 0u] = r; out[i * 3u + 1u] = g; out[i * 3u + 2u] = b;
        }
    }
}

__global__ void k_albedo_table(float *table, uint32_t n) {
    uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell < 4096u) table[cell] = albedo_table_cell(cell, n);
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct AkrContext {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string error;

    // sampler tables
    DeviceBuffer pmj, bn, albedo;
    bool albedo_ready = false;

    // scene
    DeviceBuffer nodes, prims, flat_prims, shade, instances, materials, lights, alias_j, alias_t, alias_pdf, corner_n, corner_t, corner_uv;
    DeviceBuffer svm_nodes, svm_kind_first, svm_data, svm_kind_hit_mask, svm_static_vals, textures, texels, sort_order;
    SceneView scene{};
    CornerAttribs corners{};
    bool scene_ready = false;
    bool scene_needs_table = false;
    uint32_t smem_nodes = 0, smem_prims = 0, smem_bytes = 0;  // smem_bytes = nodes + primitives (stacks come on top)
    uint32_t bvh_depth = 0;
    uint32_t class_mask = 0;   // shade classes present in the scene
    int occ_trace_dyn = 1;   // k_trace_bvh (dynamic fetch)
    int occ_trace_flat = 1;  // k_trace_flat
    int occ_raygen_fused = 1;
    int occ_shade[3] = {1, 1, 1};                // resident CTAs per SM, per shade class
    int occ_bounce[3] = {1, 1, 1};               // k_bounce<1 << class>
    uint32_t flat_bytes = 0;  // staged PrimBlock2 lists (complete + occluder-only)

    // render state
    bool render_ready = false;
    RenderParams rp{};
    AkrPtConfig cfg{};
    uint32_t tile_y0 = 0, tile_y1 = 0;
    uint32_t n_pixels = 0;
    uint32_t spp_done = 0;
    DeviceBuffer film;

    // wave buffers
    uint32_t wave_capacity = 0;  // paths
    uint32_t wave_layout = 0;    // 0 = none, 1 = queued, 2 | class_mask << 8 = fused
    DeviceBuffer wave_mem;
    RecQueue cq[2][CLS_COUNT]{};
    PathQueue q[2]{};
    HitQueue hits{};
    ShadowQueue shadow{};
    ClassQueues cls{};
    AccView acc{};
    DeviceBuffer counters, totals, first_hits;

    int aov_mode = -1;        // >= 0: the render in progress is the `aov` method with this AKR_AOV_* output
    uint32_t aov_remap = 0;
    AkrEngineOptions opts{};
    AkrStats stats{};
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    // fused pipeline with several shade classes: the classes of one depth are independent, so their kernels run on two
    // streams and the second class's CTAs fill the tail of the first (fork / join with events at every depth)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_main = nullptr, ev_side = nullptr;
    std::vector<cudaEvent_t> stage_events;
};

namespace {

int fail(AkrContext *ctx, int code, const std::string &msg) {
    if (ctx) ctx->error = msg;
    return code;
}
#define AKR_CUDA(ctx, call)                                                                                     \
    do {                                                                                                        \
        cudaError_t e__ = (call);                                                                               \
        if (e__ != cudaSuccess) return fail((ctx), AKR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

int dev_alloc(AkrContext *ctx, DeviceBuffer &b, size_t bytes) {
    if (b.ptr && b.bytes >= bytes) return AKR_OK;
    if (b.ptr) {
        cudaFree(b.ptr);
        b.ptr = nullptr;
        b.bytes = 0;
    }
    if (bytes == 0) return AKR_OK;
    cudaError_t e = cudaMalloc(&b.ptr, bytes);
    if (e != cudaSuccess) {
        b.ptr = nullptr;
        return fail(ctx, e == cudaErrorMemoryAllocation ? AKR_ERR_OUT_OF_MEMORY : AKR_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    b.bytes = bytes;
    return AKR_OK;
}
void dev_free(DeviceBuffer &b) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
}
template <class T> int upload_vec(AkrContext *ctx, DeviceBuffer &b, const std::vector<T> &v) {
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    int rc = dev_alloc(ctx, b, bytes);
    if (rc != AKR_OK) return rc;
    if (!v.empty()) AKR_CUDA(ctx, cudaMemcpyAsync(b.ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return AKR_OK;
}

int grid_for(const AkrContext *ctx, uint32_t n, int ctas_per_sm) {
    uint32_t need = (n + kBlock - 1) / kBlock;
    uint32_t cap = (uint32_t)(ctx->sm_count * ctas_per_sm);
    return (int)std::max(1u, std::min(need, cap));
}

// One allocation per wave layout.  queued: two path queues (3 + 3 records of 16 B), hits (1), shadow queue (3 + one
// word), accumulators (2), class (slot, path_id) lists (2 words per class).  fused: per shade class present two record
// queues (4 + 4 records of 16 B: one per depth parity) and the accumulators (2).
int ensure_wave_buffers(AkrContext *ctx, uint32_t capacity, bool fused, uint32_t class_mask) {
    const uint32_t layout = fused ? (2u | (class_mask << 8)) : 1u;
    const size_t cap = ((size_t)capacity + 31u) & ~(size_t)31u;
    if ((class_mask & (1u << CLS_GENERAL)) && ctx->sort_order.bytes < cap * sizeof(uint32_t)) {  // index order of the general class (k_sort_*)
        int rc = dev_alloc(ctx, ctx->sort_order, cap * sizeof(uint32_t));
        if (rc != AKR_OK) return rc;
    }
    if (ctx->wave_capacity >= capacity && ctx->wave_mem.ptr && ctx->wave_layout == layout) return AKR_OK;
    size_t n_classes = 0;
    for (uint32_t c = 0; c < (uint32_t)CLS_COUNT; ++c) n_classes += (class_mask >> c) & 1u;
    const size_t n_vec = fused ? 8 * n_classes + 2 : 3 + 3 + 1 + 3 + 2, n_word = fused ? 0 : 2 + 2 * (size_t)CLS_COUNT;
    int rc = dev_alloc(ctx, ctx->wave_mem, cap * (n_vec * 16 + n_word * 4));
    if (rc != AKR_OK) return rc;

    f4 *vbase = static_cast<f4 *>(ctx->wave_mem.ptr);
    size_t voff = 0;
    auto take_v = [&]() {
        f4 *p = vbase + voff;
        voff += cap;
        return p;
    };
    std::memset(ctx->q, 0, sizeof(ctx->q));
    std::memset(ctx->cq, 0, sizeof(ctx->cq));
    std::memset(&ctx->hits, 0, sizeof(ctx->hits));
    std::memset(&ctx->shadow, 0, sizeof(ctx->shadow));
    std::memset(&ctx->cls, 0, sizeof(ctx->cls));
    ctx->acc.l = take_v();
    ctx->acc.b = take_v();
    ctx->acc.poison = nullptr;  // fused pipeline: the radiance lives in the record, a miss adds its NaN in place
    if (fused) {
        for (int k = 0; k < 2; ++k)
            for (uint32_t c = 0; c < (uint32_t)CLS_COUNT; ++c)
                if ((class_mask >> c) & 1u)
                    for (int j = 0; j < 4; ++j) ctx->cq[k][c].r[j] = take_v();
    } else {
        for (int k = 0; k < 2; ++k) {
            ctx->q[k].a = take_v();
            ctx->q[k].b = take_v();
            ctx->q[k].c = take_v();
        }
        ctx->hits.h = take_v();
        ctx->shadow.a = take_v();
        ctx->shadow.b = take_v();
        ctx->shadow.c = take_v();
        uint32_t *wbase = reinterpret_cast<uint32_t *>(vbase + voff);
        for (uint32_t c = 0; c < (uint32_t)CLS_COUNT; ++c) ctx->cls.idx[c] = reinterpret_cast<uint2 *>(wbase + cap * 2 * c);
        ctx->shadow.ex1 = wbase + cap * 2 * (size_t)CLS_COUNT;
        ctx->acc.poison = wbase + cap * (2 * (size_t)CLS_COUNT + 1);
    }
    ctx->wave_capacity = capacity;
    ctx->wave_layout = layout;
    return AKR_OK;
}

int ensure_albedo_table(AkrContext *ctx) {
    if (ctx->albedo_ready) return AKR_OK;
    int rc = dev_alloc(ctx, ctx->albedo, 4096 * sizeof(float));
    if (rc != AKR_OK) return rc;
    k_albedo_table<<<16, 256, 0, ctx->stream>>>(static_cast<float *>(ctx->albedo.ptr), 64u);
    AKR_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    ctx->stats.launches_kernel[5] += 1;
    ctx->albedo_ready = true;
    return AKR_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int akr_b200_create(int device_ordinal, AkrContext **out_ctx) {
    if (!out_ctx) return AKR_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return AKR_ERR_CUDA;  // no CPU fallback
    if (device_ordinal < 0 || device_ordinal >= n) return AKR_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return AKR_ERR_CUDA;
    AkrContext *ctx = new AkrContext();
    ctx->device = device_ordinal;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) {
        delete ctx;
        return AKR_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaEventCreate(&ctx->ev_start) != cudaSuccess || cudaEventCreate(&ctx->ev_stop) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return AKR_ERR_CUDA;
    }
    {
        cudaError_t e = cudaSuccess;
        auto opt_in = [&](const void *fn) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        };
        opt_in((const void *)k_trace_bvh<false, false>);
        opt_in((const void *)k_trace_bvh<false, true>);
        opt_in((const void *)k_trace_bvh<true, false>);
        opt_in((const void *)k_trace_bvh<true, true>);
        opt_in((const void *)k_trace_flat<false>);
        opt_in((const void *)k_trace_flat<true>);
        opt_in((const void *)k_raygen_fused);
        opt_in((const void *)k_bounce<1u>);
        opt_in((const void *)k_bounce<2u>);
        opt_in((const void *)k_bounce<4u>);
        if (e != cudaSuccess) {  // no sm_100a image for this device, or the opt-in shared-memory size is not available
            delete ctx;
            return AKR_ERR_CUDA;
        }
    }
    if (dev_alloc(ctx, ctx->counters, kMaxDepthSlots * (kCtrStride + 64u) * sizeof(uint32_t)) != AKR_OK ||  // (+ the sort histograms)
        dev_alloc(ctx, ctx->totals, 4 * sizeof(unsigned long long)) != AKR_OK) {
        delete ctx;
        return AKR_ERR_CUDA;
    }
    cudaMemset(ctx->counters.ptr, 0, ctx->counters.bytes);
    cudaMemset(ctx->totals.ptr, 0, ctx->totals.bytes);
    *out_ctx = ctx;
    return AKR_OK;
}

void akr_b200_destroy(AkrContext *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (DeviceBuffer *b : {&ctx->pmj, &ctx->bn, &ctx->albedo, &ctx->nodes, &ctx->prims, &ctx->flat_prims, &ctx->shade, &ctx->instances, &ctx->materials, &ctx->lights,
                            &ctx->alias_j, &ctx->alias_t, &ctx->alias_pdf, &ctx->corner_n, &ctx->corner_t, &ctx->corner_uv, &ctx->svm_nodes,
                            &ctx->svm_kind_first, &ctx->svm_data, &ctx->svm_kind_hit_mask, &ctx->svm_static_vals, &ctx->textures, &ctx->texels, &ctx->sort_order, &ctx->film, &ctx->wave_mem, &ctx->counters,
                            &ctx->totals, &ctx->first_hits})
        dev_free(*b);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    for (cudaEvent_t e : ctx->stage_events) cudaEventDestroy(e);
    delete ctx;
}

const char *akr_b200_last_error(const AkrContext *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int akr_b200_set_stream(AkrContext *ctx, void *stream) {
    if (!ctx) return AKR_ERR_INVALID_ARGUMENT;
    ctx->stream = static_cast<cudaStream_t>(stream);
    return AKR_OK;
}

int akr_b200_upload_sampler_tables(AkrContext *ctx, const uint32_t *pmj02bn, const uint16_t *bluenoise) {
    if (!ctx || !pmj02bn || !bluenoise) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pmj_bytes = (size_t)AKR_PMJ02BN_SETS * AKR_PMJ02BN_SAMPLES * 2 * sizeof(uint32_t);
    const size_t bn_count = (size_t)AKR_BLUE_NOISE_TEXTURES * AKR_BLUE_NOISE_RESOLUTION * AKR_BLUE_NOISE_RESOLUTION;
    int rc = dev_alloc(ctx, ctx->pmj, pmj_bytes);
    if (rc != AKR_OK) return rc;
    rc = dev_alloc(ctx, ctx->bn, bn_count * sizeof(uint16_t));
    if (rc != AKR_OK) return rc;
    std::vector<uint16_t> bnt(bn_count);
    transpose_bluenoise(bluenoise, bnt.data());
    AKR_CUDA(ctx, cudaMemcpyAsync(ctx->pmj.ptr, pmj02bn, pmj_bytes, cudaMemcpyHostToDevice, ctx->stream));
    AKR_CUDA(ctx, cudaMemcpyAsync(ctx->bn.ptr, bnt.data(), bn_count * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // bnt is a local
    return AKR_OK;
}

int akr_b200_upload_albedo_table(AkrContext *ctx, const float *table) {
    if (!ctx || !table) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = dev_alloc(ctx, ctx->albedo, 4096 * sizeof(float));
    if (rc != AKR_OK) return rc;
    AKR_CUDA(ctx, cudaMemcpyAsync(ctx->albedo.ptr, table, 4096 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->albedo_ready = true;
    ctx->scene.albedo_table = static_cast<const float *>(ctx->albedo.ptr);
    return AKR_OK;
}

int akr_b200_upload_scene(AkrContext *ctx, const AkrSceneDesc *desc) {
    if (!ctx || !desc) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    HostSceneBlob blob;
    std::string err;
    int rc = build_scene_blob(*desc, blob, err);
    if (rc != AKR_OK) return fail(ctx, rc, "akr_b200_upload_scene: " + err);
    ctx->scene_ready = false;
    ctx->render_ready = false;
    if ((rc = upload_vec(ctx, ctx->nodes, blob.nodes)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->prims, blob.prims)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->flat_prims, blob.flat_blocks)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->shade, blob.shade)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->instances, blob.instances)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->materials, blob.materials)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->lights, blob.lights)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->alias_j, blob.alias_j)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->alias_t, blob.alias_t)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->alias_pdf, blob.alias_pdf)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->corner_n, blob.corner_normals)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->corner_t, blob.corner_tangents)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->corner_uv, blob.corner_uvs)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->svm_nodes, blob.svm_nodes)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->svm_kind_first, blob.svm_kind_first)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->svm_data, blob.svm_data)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->svm_kind_hit_mask, blob.svm_kind_hit_mask)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->svm_static_vals, blob.svm_static_vals)) != AKR_OK) return rc;
    if ((rc = upload_vec(ctx, ctx->texels, blob.texels)) != AKR_OK) return rc;
    {  // texture records: byte offsets into the texel blob become device pointers
        std::vector<TextureRec> recs = blob.textures;
        for (TextureRec &tr : recs) tr.texels = static_cast<const uint8_t *>(ctx->texels.ptr) + reinterpret_cast<size_t>(tr.texels);
        if ((rc = upload_vec(ctx, ctx->textures, recs)) != AKR_OK) return rc;
        AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // recs is a local
    }
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // blob is a local
    SceneView &v = ctx->scene;
    v.nodes = static_cast<const BvhNode *>(ctx->nodes.ptr);
    v.prims = static_cast<const PrimRec *>(ctx->prims.ptr);
    v.flat_blocks = blob.flat_blocks.empty() ? nullptr : static_cast<const PrimBlock2 *>(ctx->flat_prims.ptr);
    v.n_pair_blocks = blob.n_pair_blocks;
    v.n_single_blocks = blob.n_single_blocks;
    v.n_occ_pair_blocks = blob.n_occ_pair_blocks;
    v.n_occ_single_blocks = blob.n_occ_single_blocks;
    v.n_shadow_pair_blocks = blob.n_shadow_pair_blocks;
    v.n_shadow_single_blocks = blob.n_shadow_single_blocks;
    v.tris = nullptr;  // the Moeller-Trumbore triangle list is host-simulation data; the kernels intersect primitives
    v.shade = static_cast<const TriShade *>(ctx->shade.ptr);
    v.instances = static_cast<const InstanceRec *>(ctx->instances.ptr);
    v.materials = static_cast<const Material *>(ctx->materials.ptr);
    v.lights = static_cast<const LightRec *>(ctx->lights.ptr);
    v.alias_j = static_cast<const uint32_t *>(ctx->alias_j.ptr);
    v.alias_t = static_cast<const float *>(ctx->alias_t.ptr);
    v.alias_pdf = static_cast<const float *>(ctx->alias_pdf.ptr);
    v.n_nodes = (uint32_t)blob.nodes.size();
    v.n_prims = (uint32_t)blob.prims.size();
    v.n_tris = (uint32_t)blob.shade.size();
    v.n_instances = (uint32_t)blob.instances.size();
    v.n_materials = (uint32_t)blob.materials.size();
    v.n_lights = (uint32_t)blob.lights.size();
    v.any_alpha = blob.any_alpha;
    v.camera = blob.camera;
    v.svm.nodes = static_cast<const AkrSvmNode *>(ctx->svm_nodes.ptr);
    v.svm.kind_first = static_cast<const uint32_t *>(ctx->svm_kind_first.ptr);
    v.svm.data = static_cast<const uint8_t *>(ctx->svm_data.ptr);
    v.svm.textures = static_cast<const TextureRec *>(ctx->textures.ptr);
    v.svm.n_kinds = (uint32_t)blob.svm_kind_first.size() - 1u;
    v.svm.n_textures = (uint32_t)blob.textures.size();
    v.svm.data_size = (uint32_t)blob.svm_data.size();
    v.svm.kind_hit_mask = static_cast<const uint64_t *>(ctx->svm_kind_hit_mask.ptr);
    v.svm.static_vals = static_cast<const SvmVal *>(ctx->svm_static_vals.ptr);
    v.corner_uvs = blob.corner_uvs.empty() ? nullptr : static_cast<const float *>(ctx->corner_uv.ptr);
    ctx->corners.normals = blob.corner_normals.empty() ? nullptr : static_cast<const float *>(ctx->corner_n.ptr);
    ctx->corners.tangents = blob.corner_tangents.empty() ? nullptr : static_cast<const float *>(ctx->corner_t.ptr);
    ctx->scene_needs_table = false;
    for (const Material &m : blob.materials)
        if (m.type == MAT_PRINCIPLED && (m.lobes & (LOBE_COAT | LOBE_SPECULAR))) ctx->scene_needs_table = true;
    // shared-memory staging plan: top of the BVH first, then all primitives if they still fit
    const uint32_t tri_bytes = v.n_prims * (uint32_t)sizeof(PrimRec);
    // whole scene in shared memory when it fits the budget; otherwise only the top of the tree (breadth-first order)
    // so that several CTAs stay resident per SM — the deeper nodes and the primitives come through L1/L2
    uint32_t used = v.n_nodes * (uint32_t)sizeof(BvhNode);
    ctx->smem_prims = (used + tri_bytes <= kSmemSceneBudget) ? 1u : 0u;
    if (ctx->smem_prims) {
        ctx->smem_nodes = v.n_nodes;
        used += tri_bytes;
    } else {
        // default 4 KiB = the top six levels.  Measured on the 8.5 K-triangle scene (trace stage of a pass): 2 / 8 / 16 / 32 / 48 KiB =
        // 63.8 / 63.9 / 65.2 / 69.0 / 77.8 ms — shared memory given to the tree is L1 taken from everything else
        const uint32_t top_bytes = (ctx->opts.smem_node_kb ? ctx->opts.smem_node_kb : 4u) * 1024u;
        ctx->smem_nodes = std::min(v.n_nodes, std::min(top_bytes, kSmemSceneBudget) / (uint32_t)sizeof(BvhNode));
        used = ctx->smem_nodes * (uint32_t)sizeof(BvhNode);
    }
    ctx->smem_bytes = used;
    ctx->bvh_depth = blob.bvh_depth;
    ctx->class_mask = 0;
    for (const Material &m : blob.materials) ctx->class_mask |= 1u << shade_class_of(m);
    // resident CTAs per SM of every kernel variant with this scene's shared-memory footprint
    {
        cudaError_t e = cudaSuccess;
        auto occ = [&](int &out, const void *fn, int block, size_t smem) {
            int v = 0;
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, fn, block, smem);
            out = std::max(v, 1);
        };
        const size_t smem_bvh = used + (size_t)(blob.bvh_depth + 2u) * kBlock * sizeof(int32_t);
        ctx->flat_bytes = (uint32_t)(blob.flat_blocks.size() * sizeof(PrimBlock2));
        const size_t smem_flat = (size_t)ctx->smem_nodes * sizeof(BvhNode) + ctx->flat_bytes;
        const size_t smem_bounce = ((ctx->flat_bytes + 127u) & ~127u) + (size_t)kShadeWarps * 2u * kTileBytes;
        const size_t smem_bounce_general = ((ctx->flat_bytes + 127u) & ~127u) + kGeneralTileSmem;
        const bool all = ctx->smem_prims != 0;
        if (blob.any_alpha) {
            occ(ctx->occ_trace_flat, (const void *)k_trace_flat<true>, kBlock, smem_flat);
            occ(ctx->occ_trace_dyn, all ? (const void *)k_trace_bvh<true, true> : (const void *)k_trace_bvh<false, true>, kBlock, smem_bvh);
        } else {
            occ(ctx->occ_trace_flat, (const void *)k_trace_flat<false>, kBlock, smem_flat);
            occ(ctx->occ_trace_dyn, all ? (const void *)k_trace_bvh<true, false> : (const void *)k_trace_bvh<false, false>, kBlock, smem_bvh);
        }
        occ(ctx->occ_shade[0], (const void *)k_shade<CLS_LAMBERT>, kShadeBlock, 0);
        occ(ctx->occ_shade[1], (const void *)k_shade<CLS_CONDUCTOR>, kShadeBlock, 0);
        occ(ctx->occ_shade[2], (const void *)k_shade<CLS_GENERAL>, kGeneralBlock, 0);
        occ(ctx->occ_raygen_fused, (const void *)k_raygen_fused, kBlock, ctx->flat_bytes);
        occ(ctx->occ_bounce[0], (const void *)k_bounce<1u>, kShadeBlock, smem_bounce);
        occ(ctx->occ_bounce[1], (const void *)k_bounce<2u>, kShadeBlock, smem_bounce);
        occ(ctx->occ_bounce[2], (const void *)k_bounce<4u>, kGeneralBlock, smem_bounce_general);
        if (e != cudaSuccess) return fail(ctx, AKR_ERR_CUDA, std::string("cudaOccupancyMaxActiveBlocksPerMultiprocessor: ") + cudaGetErrorString(e));
    }
    ctx->scene_ready = true;
    return AKR_OK;
}

int akr_b200_set_engine_options(AkrContext *ctx, const AkrEngineOptions *opts) {
    if (!ctx || !opts) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    ctx->opts = *opts;
    return AKR_OK;
}

int akr_b200_begin(AkrContext *ctx, const AkrPtConfig *cfg, const AkrSamplerConfig *sampler, const AkrFilterConfig *filter, const AkrTile *tile) {
    if (!ctx || !cfg || !sampler || !filter) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    if (!ctx->scene_ready) return fail(ctx, AKR_ERR_STATE, "akr_b200_begin: no scene uploaded");
    if (!ctx->pmj.ptr || !ctx->bn.ptr) return fail(ctx, AKR_ERR_STATE, "akr_b200_begin: sampler tables not uploaded");
    if (sampler->type != AKR_SAMPLER_PMJ02BN)
        return fail(ctx, AKR_ERR_UNSUPPORTED, "only the pmj02bn sampler is implemented (independent seeds from rand::StdRng, sampler/mod.rs:148-160)");
    if (cfg->spp == 0 || cfg->spp > AKR_PMJ02BN_SAMPLES) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "Pmj02BnSampler supports 1..65536 spp (sampler/mod.rs:374-380)");
    if (cfg->max_depth + 2 > kMaxDepthSlots) return fail(ctx, AKR_ERR_UNSUPPORTED, "max_depth > 64");
    if (filter->type != AKR_FILTER_BOX && filter->type != AKR_FILTER_GAUSSIAN) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "unknown filter");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t width = ctx->scene.camera.width, height = ctx->scene.camera.height;
    uint32_t y0 = 0, y1 = height, t_block = 1, t_shards = 1, t_shard = 0;
    if (tile) {
        y0 = tile->y0;
        y1 = tile->y1;
        if (tile->n_shards > 1u) {
            t_block = tile->block_rows ? tile->block_rows : 1u;
            t_shards = tile->n_shards;
            t_shard = tile->shard;
        }
    }
    if (y0 >= y1 || y1 > height || t_shard >= t_shards) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "bad tile");
    const uint32_t tile_rows = interleaved_tile_rows(y0, y1, t_block, t_shards, t_shard);
    if (tile_rows == 0u) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "bad tile: the shard owns no row");
    if (ctx->scene_needs_table) {
        int rc = ensure_albedo_table(ctx);
        if (rc != AKR_OK) return rc;
    } else if (!ctx->albedo_ready) {
        int rc = dev_alloc(ctx, ctx->albedo, 4096 * sizeof(float));  // never read, but keep the pointer valid
        if (rc != AKR_OK) return rc;
        AKR_CUDA(ctx, cudaMemsetAsync(ctx->albedo.ptr, 0, 4096 * sizeof(float), ctx->stream));
    }
    ctx->scene.albedo_table = static_cast<const float *>(ctx->albedo.ptr);
    RenderParams &rp = ctx->rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.spp_total = cfg->spp;
    uint32_t w = cfg->spp - 1;  // sampler/mod.rs:381-386
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    rp.w_mask = w;
    rp.seed = (uint32_t)sampler->seed;
    rp.max_depth = cfg->max_depth;
    rp.rr_depth = cfg->rr_depth;
    rp.use_nee = cfg->use_nee ? 1u : 0u;
    rp.indirect_only = cfg->indirect_only ? 1u : 0u;
    rp.force_diffuse = cfg->force_diffuse ? 1u : 0u;
    rp.pixel_offset_x = cfg->pixel_offset[0];
    rp.pixel_offset_y = cfg->pixel_offset[1];
    rp.debug_depth = cfg->debug_depth;
    rp.filter_type = filter->type;
    rp.filter_radius = filter->radius;
    rp.width = width;
    rp.height = height;
    rp.y0 = y0;
    rp.tile_block = t_block;
    rp.tile_shards = t_shards;
    rp.tile_shard = t_shard;
    finish_render_params(rp);
    ctx->cfg = *cfg;
    ctx->tile_y0 = y0;
    ctx->tile_y1 = y1;
    ctx->n_pixels = width * tile_rows;
    ctx->spp_done = 0;
    int rc = dev_alloc(ctx, ctx->film, (size_t)ctx->n_pixels * 7 * sizeof(float));
    if (rc != AKR_OK) return rc;
    AKR_CUDA(ctx, cudaMemsetAsync(ctx->film.ptr, 0, (size_t)ctx->n_pixels * 7 * sizeof(float), ctx->stream));  // Film::clear (film.rs:230-233)
    if (ctx->opts.aov_mask & AKR_AOV_FIRST_HIT_IDS) {
        rc = dev_alloc(ctx, ctx->first_hits, (size_t)ctx->n_pixels * 2 * sizeof(uint32_t));
        if (rc != AKR_OK) return rc;
        AKR_CUDA(ctx, cudaMemsetAsync(ctx->first_hits.ptr, 0xff, (size_t)ctx->n_pixels * 2 * sizeof(uint32_t), ctx->stream));
    } else {
        dev_free(ctx->first_hits);
    }
    AKR_CUDA(ctx, cudaMemsetAsync(ctx->counters.ptr, 0, ctx->counters.bytes, ctx->stream));
    ctx->render_ready = true;
    ctx->aov_mode = -1;
    return AKR_OK;
}

int akr_b200_render_pass(AkrContext *ctx, uint32_t n_spp, int blocking) {
    if (!ctx) return AKR_ERR_INVALID_ARGUMENT;
    if (!ctx->render_ready) return fail(ctx, AKR_ERR_STATE, "akr_b200_render_pass: call akr_b200_begin first");
    if (n_spp == 0) return AKR_OK;
    if (ctx->spp_done + n_spp > ctx->cfg.spp) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "pass exceeds the configured spp (sampler/mod.rs:666-668)");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    // trace schedule: flat list for tiny scenes that fit in shared memory, BVH otherwise (opts.trace_mode overrides)
    const int bvh_mode = ctx->smem_prims ? TRACE_BVH_SMEM : TRACE_BVH;
    const bool flat_ok = ctx->scene.flat_blocks != nullptr && ctx->smem_prims;
    int trace_mode = flat_ok ? TRACE_FLAT : bvh_mode;
    if (ctx->opts.trace_mode == 1u) trace_mode = bvh_mode;
    if (ctx->opts.trace_mode == 2u && flat_ok) trace_mode = TRACE_FLAT;
    const bool alpha = ctx->scene.any_alpha != 0u;
    // fused pipeline: flat list, no stochastic alpha; opts.fused = 2 forces the queued pipeline
    const bool aov = ctx->aov_mode >= 0;  // the aov method runs raygen + one trace stage of the queued pipeline + k_aov
    const bool fused = trace_mode == TRACE_FLAT && !alpha && ctx->opts.fused != 2u && !aov && ctx->rp.width <= 65535u && ctx->rp.height <= 65535u;  // (records pack the pixel as x | y << 16)
    const uint32_t class_mask = ctx->rp.force_diffuse ? (1u << CLS_LAMBERT) : ctx->class_mask;

    // wave geometry: pixels x samples with pixels * samples <= capacity
    uint32_t cap = ctx->opts.wave_size ? ctx->opts.wave_size : (1u << 22);
    cap = std::max(cap, 1024u);
    uint32_t spp_chunk = std::min(n_spp, std::max(1u, cap / 32u));
    uint32_t pix_chunk = std::max(32u, (cap / spp_chunk) & ~31u);
    pix_chunk = std::min(pix_chunk, ctx->n_pixels);
    int rc = ensure_wave_buffers(ctx, pix_chunk * spp_chunk, fused, class_mask);
    if (rc != AKR_OK) return rc;

    LaunchParams P;
    std::memset(&P, 0, sizeof(P));
    P.scene = ctx->scene;
    P.corners = ctx->corners;
    P.tables = SamplerTables{static_cast<const uint32_t *>(ctx->pmj.ptr), static_cast<const uint16_t *>(ctx->bn.ptr)};
    P.rp = ctx->rp;
    P.q[0] = ctx->q[0];
    P.q[1] = ctx->q[1];
    P.hits = ctx->hits;
    P.shadow = ctx->shadow;
    P.cls = ctx->cls;
    std::memcpy(P.cq, ctx->cq, sizeof(P.cq));
    P.acc = ctx->acc;
    P.counters = static_cast<uint32_t *>(ctx->counters.ptr);
    P.sort_hist = P.counters + kMaxDepthSlots * kCtrStride;
    P.sort_order = static_cast<uint32_t *>(ctx->sort_order.ptr);
    P.film = static_cast<float *>(ctx->film.ptr);
    P.n_film_pixels = ctx->n_pixels;
    P.scene_smem_nodes = fused ? 0u : ctx->smem_nodes;  // the fused kernels stage the flat list only
    P.scene_smem_prims = ctx->smem_prims;
    P.stack_depth = ctx->bvh_depth + 2u;
    P.first_hits = static_cast<uint32_t *>(ctx->first_hits.ptr);
    P.aov = aov ? (uint32_t)ctx->aov_mode : 0u;
    P.aov_remap = ctx->aov_remap;
    P.stage_flat = trace_mode == TRACE_FLAT ? 1u : 0u;
    // shared memory: staged nodes + (flat mode: the padded PrimBlock2 lists | BVH modes: primitives + per-thread stacks)
    const size_t node_smem = (size_t)ctx->smem_nodes * sizeof(BvhNode);
    const size_t flat_smem = node_smem + ctx->flat_bytes;
    const size_t trace_smem = trace_mode == TRACE_FLAT ? flat_smem : ctx->smem_bytes + (size_t)P.stack_depth * kBlock * sizeof(int32_t);
    const size_t bounce_smem = ((ctx->flat_bytes + 127u) & ~127u) + (size_t)kShadeWarps * 2u * kTileBytes;
    const size_t bounce_smem_general = ((ctx->flat_bytes + 127u) & ~127u) + kGeneralTileSmem;
    if (trace_smem > kSmemMax || bounce_smem > kSmemMax) return fail(ctx, AKR_ERR_UNSUPPORTED, "BVH too deep for the shared-memory traversal stack");

    const bool prof = ctx->opts.profile_stages != 0;
    struct StageMark {
        int stage;
        size_t ev;
    };
    std::vector<StageMark> marks;
    size_t ev_used = 0;
    auto mark = [&](int stage) -> int {
        if (!prof) return AKR_OK;
        for (int k = 0; k < 2; ++k) {
            if (ev_used >= ctx->stage_events.size()) {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) return AKR_ERR_CUDA;
                ctx->stage_events.push_back(e);
            }
            if (k == 0) marks.push_back({stage, ev_used});
            ++ev_used;
        }
        return AKR_OK;
    };
    auto count_launch = [&](int stage) {
        ctx->stats.kernel_launches += 1;
        ctx->stats.launches_kernel[stage] += 1;
    };
#define AKR_LAUNCH(stage, kernel, grid, smem, ...) AKR_LAUNCH_B(stage, kernel, grid, kBlock, smem, __VA_ARGS__)
#define AKR_LAUNCH_B(stage, kernel, grid, block, smem, ...)                                                   \
    do {                                                                                                      \
        if (prof) {                                                                                           \
            if (mark(stage) != AKR_OK) return fail(ctx, AKR_ERR_CUDA, "cudaEventCreate failed");              \
            cudaEventRecord(ctx->stage_events[marks.back().ev], ctx->stream);                                  \
        }                                                                                                     \
        kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);                                         \
        if (prof) cudaEventRecord(ctx->stage_events[marks.back().ev + 1], ctx->stream);                        \
        count_launch(stage);                                                                                  \
    } while (0)

    AKR_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    const uint32_t s_begin = ctx->spp_done;
    for (uint32_t pix0 = 0; pix0 < ctx->n_pixels; pix0 += pix_chunk) {
        const uint32_t n_pix = std::min(pix_chunk, ctx->n_pixels - pix0);
        for (uint32_t s0 = 0; s0 < n_spp; s0 += spp_chunk) {
            const uint32_t k = std::min(spp_chunk, n_spp - s0);
            P.wave = make_wave(pix0, n_pix, s_begin + s0, k);
            const uint32_t n_paths = n_pix * k;
            auto shade_grid = [&](int occ, int block = kShadeBlock) {
                uint32_t need = (n_paths + (uint32_t)block - 1u) / (uint32_t)block, capn = (uint32_t)(ctx->sm_count * occ);
                return (int)std::max(1u, std::min(need, capn));
            };
            if (fused) {
                AKR_LAUNCH(0, k_raygen_fused, grid_for(ctx, n_paths, ctx->occ_raygen_fused), ctx->flat_bytes, P);
                // Lambert on the context stream, conductor (and general) on the side stream when both exist: each depth forks
                // after the previous depth's kernels and joins before the next (per-stage profiling keeps one stream)
                const bool two_streams = !prof && ctx->opts.fused != 3u && (class_mask & 1u) && (class_mask & 6u);
                cudaStream_t side = two_streams ? ctx->side_stream : ctx->stream;
                for (uint32_t depth = 0; depth < ctx->rp.max_depth; ++depth) {
                    if (two_streams) {  // the side stream waits for everything the context stream has enqueued so far
                        cudaEventRecord(ctx->ev_main, ctx->stream);
                        cudaStreamWaitEvent(side, ctx->ev_main, 0);
                    }
                    if (class_mask & 1u) AKR_LAUNCH_B(2, (k_bounce<1u>), shade_grid(ctx->occ_bounce[0]), kShadeBlock, bounce_smem, P, depth);
                    if (class_mask & 2u) {
                        if (prof) AKR_LAUNCH_B(3, (k_bounce<2u>), shade_grid(ctx->occ_bounce[1]), kShadeBlock, bounce_smem, P, depth);
                        else {
                            k_bounce<2u><<<shade_grid(ctx->occ_bounce[1]), kShadeBlock, bounce_smem, side>>>(P, depth);
                            count_launch(3);
                        }
                    }
                    if (class_mask & 4u) {
                        {  // order the depth's general records by material sort key
                            const int sg = grid_for(ctx, n_paths, 8);
                            if (prof) {
                                AKR_LAUNCH_B(6, k_sort_hist<false>, sg, 256, 0, P, depth);
                                AKR_LAUNCH_B(6, k_sort_scatter<false>, sg, 256, 0, P, depth);
                            } else {
                                k_sort_hist<false><<<sg, 256, 0, side>>>(P, depth);
                                k_sort_scatter<false><<<sg, 256, 0, side>>>(P, depth);
                                count_launch(6);
                                count_launch(6);
                            }
                        }
                        if (prof) AKR_LAUNCH_B(6, (k_bounce<4u>), shade_grid(ctx->occ_bounce[2], kGeneralBlock), kGeneralBlock, bounce_smem_general, P, depth);
                        else {
                            k_bounce<4u><<<shade_grid(ctx->occ_bounce[2], kGeneralBlock), kGeneralBlock, bounce_smem_general, side>>>(P, depth);
                            count_launch(6);
                        }
                    }
                    if (two_streams) {  // join: the next depth (or k_accumulate) needs both
                        cudaEventRecord(ctx->ev_side, side);
                        cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0);
                    }
                }
            } else {
                AKR_LAUNCH(0, k_raygen, grid_for(ctx, n_paths, 8), 0, P);
                for (uint32_t depth = 0; depth <= (aov ? 0u : ctx->rp.max_depth); ++depth) {
                    if (trace_mode == TRACE_FLAT) {
                        const int g_trace = grid_for(ctx, n_paths, ctx->occ_trace_flat);
                        if (alpha) AKR_LAUNCH(1, (k_trace_flat<true>), g_trace, trace_smem, P, depth);
                        else AKR_LAUNCH(1, (k_trace_flat<false>), g_trace, trace_smem, P, depth);
                    } else {
                        const int g_dyn = grid_for(ctx, n_paths, ctx->occ_trace_dyn);
                        if (trace_mode == TRACE_BVH_SMEM) {
                            if (alpha) AKR_LAUNCH(1, (k_trace_bvh<true, true>), g_dyn, trace_smem, P, depth);
                            else AKR_LAUNCH(1, (k_trace_bvh<true, false>), g_dyn, trace_smem, P, depth);
                        } else {
                            if (alpha) AKR_LAUNCH(1, (k_trace_bvh<false, true>), g_dyn, trace_smem, P, depth);
                            else AKR_LAUNCH(1, (k_trace_bvh<false, false>), g_dyn, trace_smem, P, depth);
                        }
                    }
                    if (aov) {
                        AKR_LAUNCH(5, k_aov, grid_for(ctx, n_paths, 8), 0, P);
                        break;
                    }
                    if (class_mask & (1u << CLS_LAMBERT)) AKR_LAUNCH_B(2, (k_shade<CLS_LAMBERT>), shade_grid(ctx->occ_shade[0]), kShadeBlock, 0, P, depth);
                    if (class_mask & (1u << CLS_CONDUCTOR)) AKR_LAUNCH_B(3, (k_shade<CLS_CONDUCTOR>), shade_grid(ctx->occ_shade[1]), kShadeBlock, 0, P, depth);
                    if (class_mask & (1u << CLS_GENERAL)) {
                        const int sg = grid_for(ctx, n_paths, 8);
                        AKR_LAUNCH_B(6, k_sort_hist<true>, sg, 256, 0, P, depth);
                        AKR_LAUNCH_B(6, k_sort_scatter<true>, sg, 256, 0, P, depth);
                    }
                    if (class_mask & (1u << CLS_GENERAL)) AKR_LAUNCH_B(6, (k_shade<CLS_GENERAL>), shade_grid(ctx->occ_shade[2], kGeneralBlock), kGeneralBlock, 0, P, depth);
                }
            }
            AKR_LAUNCH(4, k_accumulate, grid_for(ctx, n_pix, 8), 0, P);
            AKR_LAUNCH_B(5, k_fold_counters, 1, 128, 0, P.counters, static_cast<unsigned long long *>(ctx->totals.ptr), ctx->rp.max_depth + 2u, P.sort_hist);
            ctx->stats.samples += n_paths;
        }
    }
#undef AKR_LAUNCH
#undef AKR_LAUNCH_B
    AKR_CUDA(ctx, cudaGetLastError());
    AKR_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    ctx->spp_done += n_spp;
    if (blocking || prof) {
        AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        AKR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        ctx->stats.gpu_ms += ms;
        if (prof) {
            for (const StageMark &m : marks) {
                float t = 0.0f;
                cudaEventElapsedTime(&t, ctx->stage_events[m.ev], ctx->stage_events[m.ev + 1]);
                ctx->stats.gpu_ms_kernel[m.stage] += t;
            }
        }
    }
    return AKR_OK;
}

int akr_b200_render_pt(AkrContext *ctx, const AkrPtConfig *cfg, const AkrSamplerConfig *sampler, const AkrFilterConfig *filter, const AkrTile *tile) {
    int rc = akr_b200_begin(ctx, cfg, sampler, filter, tile);
    if (rc != AKR_OK) return rc;
    // host pass loop of PathTracer::render (pt.rs:1126-1149)
    uint32_t cnt = 0;
    const uint32_t per_pass = std::max(1u, cfg->spp_per_pass);
    while (cnt < cfg->spp) {
        uint32_t cur = std::min(cfg->spp - cnt, per_pass);
        rc = akr_b200_render_pass(ctx, cur, 0);
        if (rc != AKR_OK) return rc;
        cnt += cur;
    }
    return akr_b200_synchronize(ctx);
}

int akr_b200_render_aov(AkrContext *ctx, const AkrAovConfig *cfg, const AkrSamplerConfig *sampler, const AkrFilterConfig *filter, const AkrTile *tile) {
    if (!ctx || !cfg) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    if (cfg->aov > AKR_AOV_ROUGHNESS) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "unknown aov");
    AkrPtConfig pc;
    std::memset(&pc, 0, sizeof(pc));
    pc.spp = cfg->spp;
    pc.spp_per_pass = cfg->spp;  // aov.rs renders all samples in one dispatch (:160-166)
    pc.debug_depth = -1;
    int rc = akr_b200_begin(ctx, &pc, sampler, filter, tile);
    if (rc != AKR_OK) return rc;
    ctx->aov_mode = (int)cfg->aov;
    ctx->aov_remap = cfg->remap ? 1u : 0u;
    rc = akr_b200_render_pass(ctx, cfg->spp, 1);
    ctx->aov_mode = -1;
    return rc;
}

int akr_b200_synchronize(AkrContext *ctx) {
    if (!ctx) return AKR_ERR_INVALID_ARGUMENT;
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AKR_OK;
}

int akr_b200_download_film(AkrContext *ctx, float *out, size_t n_floats) {
    if (!ctx || !out) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    if (!ctx->render_ready) return fail(ctx, AKR_ERR_STATE, "no film");
    if (n_floats != (size_t)ctx->n_pixels * 7) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "film size mismatch (expected 7 * width * rows floats)");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    AKR_CUDA(ctx, cudaMemcpyAsync(out, ctx->film.ptr, n_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AKR_OK;
}

int akr_b200_resolve_film_device(AkrContext *ctx, void *out_device, size_t n_floats, int rgba) {
    if (!ctx || !out_device) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    if (!ctx->render_ready) return fail(ctx, AKR_ERR_STATE, "no film");
    if (n_floats != (size_t)ctx->n_pixels * (rgba ? 4 : 3)) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "image size mismatch");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    k_resolve_film<<<grid_for(ctx, ctx->n_pixels, 8), kBlock, 0, ctx->stream>>>(static_cast<const float *>(ctx->film.ptr), ctx->n_pixels,
                                                                                static_cast<float *>(out_device), rgba);
    AKR_CUDA(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    ctx->stats.launches_kernel[5] += 1;
    return AKR_OK;
}

int akr_b200_resolve_film(AkrContext *ctx, float *out_host, size_t n_floats, int rgba) {
    if (!ctx || !out_host) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    DeviceBuffer tmp;
    int rc = dev_alloc(ctx, tmp, n_floats * sizeof(float));
    if (rc != AKR_OK) return rc;
    rc = akr_b200_resolve_film_device(ctx, tmp.ptr, n_floats, rgba);
    if (rc == AKR_OK) {
        cudaError_t e = cudaMemcpyAsync(out_host, tmp.ptr, n_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, AKR_ERR_CUDA, cudaGetErrorString(e));
    }
    dev_free(tmp);
    return rc;
}

uint32_t akr_b200_tile_rows(const AkrTile *tile) {
    if (!tile) return 0u;
    return interleaved_tile_rows(tile->y0, tile->y1, tile->block_rows, tile->n_shards, tile->shard);
}

int akr_b200_get_stats(AkrContext *ctx, AkrStats *out) {
    if (!ctx || !out) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long t[3] = {0, 0, 0};
    AKR_CUDA(ctx, cudaMemcpy(t, ctx->totals.ptr, sizeof(t), cudaMemcpyDeviceToHost));
    ctx->stats.segments = t[0];
    ctx->stats.shadow_rays = t[1];
    ctx->stats.shaded_hits = t[2];
    *out = ctx->stats;
    return AKR_OK;
}

int akr_b200_reset_stats(AkrContext *ctx) {
    if (!ctx) return AKR_ERR_INVALID_ARGUMENT;
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    AKR_CUDA(ctx, cudaMemset(ctx->totals.ptr, 0, ctx->totals.bytes));
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    return AKR_OK;
}

int akr_b200_debug_first_hits(AkrContext *ctx, uint32_t *out_inst, uint32_t *out_prim, size_t n_pixels) {
    if (!ctx || !out_inst || !out_prim) return fail(ctx, AKR_ERR_INVALID_ARGUMENT, "null argument");
    if (!ctx->render_ready || n_pixels != ctx->n_pixels) return fail(ctx, AKR_ERR_STATE, "no render / size mismatch");
    if (!ctx->first_hits.ptr) return fail(ctx, AKR_ERR_STATE, "first-hit ids were not requested (AkrEngineOptions.aov_mask & AKR_AOV_FIRST_HIT_IDS) before akr_b200_begin");
    std::vector<uint32_t> tmp(n_pixels * 2);
    AKR_CUDA(ctx, cudaSetDevice(ctx->device));
    AKR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    AKR_CUDA(ctx, cudaMemcpy(tmp.data(), ctx->first_hits.ptr, tmp.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n_pixels; ++i) {
        out_inst[i] = tmp[2 * i];
        out_prim[i] = tmp[2 * i + 1];
    }
    return AKR_OK;
}

}  // extern "C"

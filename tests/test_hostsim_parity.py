"""The per-thread bodies of the CUDA kernels, executed on the CPU by tests/hostsim (test-only), against the
oracle.  Without FMA contraction the two independent implementations (literal closure tree + brute-force
intersection vs constant-folded materials + BVH traversal + staged pipeline) must agree BIT FOR BIT."""
import ctypes as C
import os

import numpy as np
import pytest

import scene_variants as sv

HERE = os.path.dirname(os.path.abspath(__file__))


class HostsimStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("segments", C.c_uint64), ("shadow_rays", C.c_uint64), ("n_nodes", C.c_uint32),
                ("n_tris", C.c_uint32), ("n_materials", C.c_uint32), ("n_lights", C.c_uint32), ("bvh_depth", C.c_uint32),
                ("material_types", C.c_uint32 * 8), ("n_prims", C.c_uint32), ("n_pairs", C.c_uint32),
                ("flat_blocks", C.c_uint32), ("flat_occluder_blocks", C.c_uint32), ("n_nodes4", C.c_uint32), ("bvh4_depth", C.c_uint32),
                ("any_alpha", C.c_uint32), ("any_dynamic", C.c_uint32)]


@pytest.fixture(scope="module", params=["queued", "fused"])
def hostsim(request):
    """Both pipelines of the engine: `queued` (trace stage / shade stage / shadow queue: BVH scenes) and `fused` (one
    stage per depth and shade class on records that carry hit + radiance: flat scenes, akr_path.cuh bounce_fused)."""
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    lib.hostsim_set_pipeline(1 if request.param == "fused" else 0)
    yield lib
    lib.hostsim_set_pipeline(0)


def run_hostsim(lib, scene, task, tables, table, w, h, y0=0, y1=None, s0=0, s1=None, wave_pixels=0, n_rows=None):
    pmj, bn = tables
    y1 = h if y1 is None else y1
    s1 = task.pt.spp if s1 is None else s1
    n = w * ((y1 - y0) if n_rows is None else n_rows)  # n_rows: rows of an interleaved tile (hostsim_set_tile_interleave)
    film = np.zeros(7 * n, np.float32)
    fh = np.zeros((n, 2), np.uint32)
    st = HostsimStats()
    rc = lib.hostsim_render(scene.desc, C.byref(task.pt), C.byref(task.sampler), C.byref(task.filter), C.c_void_p(pmj.ctypes.data),
                            C.c_void_p(bn.ctypes.data), C.c_void_p(table.ctypes.data), y0, y1, s0, s1, wave_pixels,
                            C.c_void_p(film.ctypes.data), C.c_void_p(fh.ctypes.data), C.byref(st))
    assert rc == 0, lib.hostsim_last_error()
    return film, fh, st


def test_cbox_bitwise(hostsim, oracle, tables, cbox, cbox_task):
    w = h = 48
    scene, task = cbox(w, h), cbox_task(16)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    film, fh, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    assert np.array_equal(fh, ofh)
    assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    assert np.array_equal(film, ofilm)
    assert st.n_lights == 1 and list(st.material_types)[:3] == [7, 1, 0]  # 7 Lambert reductions + 1 conductor


def test_waves_passes_and_tiles_compose_bitwise(hostsim, oracle, tables, cbox, cbox_task):
    """Any wave / pass / tile decomposition gives the same film (sampler keyed by absolute pixel + sample)."""
    w, h = 40, 24
    scene, task = cbox(w, h), cbox_task(16)
    table = oracle.albedo_table()
    ref, _, _ = run_hostsim(hostsim, scene, task, tables, table, w, h)
    a, _, _ = run_hostsim(hostsim, scene, task, tables, table, w, h, wave_pixels=96)
    assert np.array_equal(a, ref)
    # two passes of 8 spp accumulate into the same film
    pmj, bn = tables
    film = np.zeros(7 * w * h, np.float32)
    st = HostsimStats()
    for s0, s1 in ((0, 8), (8, 16)):
        rc = hostsim.hostsim_render(scene.desc, C.byref(task.pt), C.byref(task.sampler), C.byref(task.filter), C.c_void_p(pmj.ctypes.data),
                                    C.c_void_p(bn.ctypes.data), C.c_void_p(table.ctypes.data), 0, h, s0, s1, 0, C.c_void_p(film.ctypes.data), None,
                                    C.byref(st))
        assert rc == 0
    assert np.array_equal(film, ref)
    # row tiles
    n = w * h
    top, _, _ = run_hostsim(hostsim, scene, task, tables, table, w, h, y0=0, y1=10)
    bot, _, _ = run_hostsim(hostsim, scene, task, tables, table, w, h, y0=10, y1=h)
    nt, nb = w * 10, w * (h - 10)
    assert np.array_equal(np.concatenate([top[:3 * nt], bot[:3 * nb]]), ref[:3 * n])


@pytest.mark.parametrize("variant", ["principled_mix", "nodes"])
def test_material_variants_bitwise(hostsim, oracle, tables, akr, cbox_task, tmp_path, variant):
    """General Principled tree (coat, specular, transmission, partial metallic, emission) and the
    diffuse / glass / emission nodes: pruned, constant-folded device closures == literal oracle tree."""
    w = h = 40
    path = sv.write_variant(tmp_path, variant, getattr(sv, "variant_" + variant))
    scene = akr.load_scene(path).set_resolution(w, h)
    task = cbox_task(16)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    film, fh, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    assert np.array_equal(fh, ofh)
    assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    assert np.array_equal(film, ofilm), np.abs(film - ofilm).max()
    assert ost.n_lights == st.n_lights >= 2


@pytest.mark.parametrize("kw", [dict(use_nee=0), dict(max_depth=0), dict(max_depth=1), dict(rr_depth=0), dict(indirect_only=1),
                                dict(force_diffuse=1), dict(debug_depth=2), dict(pixel_offset=(3, -2))])
def test_config_knobs_bitwise(hostsim, oracle, tables, cbox, cbox_task, kw):
    """pt::Config knobs (pt.rs:916-944): use_nee, max_depth, rr_depth, indirect_only, force_diffuse, debug_depth, pixel_offset."""
    w = h = 32
    scene = cbox(w, h)
    task = cbox_task(16)
    for k, v in kw.items():
        if k == "pixel_offset":
            task.pt.pixel_offset[0], task.pt.pixel_offset[1] = v
        else:
            setattr(task.pt, k, v)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, _ = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn)
    film, _, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    assert np.array_equal(film, ofilm)


def test_interleaved_tiles_compose_bitwise(hostsim, oracle, tables, cbox, cbox_task):
    """Interleaved row blocks (AkrTile block_rows / n_shards / shard): the rows a shard renders are bit-identical to the
    same rows of the whole frame, for block heights that do and do not divide the frame."""
    w, h = 24, 22
    scene, task = cbox(w, h), cbox_task(8)
    table = oracle.albedo_table()
    ref, _, _ = run_hostsim(hostsim, scene, task, tables, table, w, h)
    ref_rgb = ref[:3 * w * h].reshape(h, w, 3)
    try:
        for block, shards in ((1, 2), (4, 3), (5, 2), (8, 4)):
            for shard in range(shards):
                rows = [y for y in range(h) if (y // block) % shards == shard]
                hostsim.hostsim_set_tile_interleave(block, shards, shard)
                film, _, _ = run_hostsim(hostsim, scene, task, tables, table, w, h, n_rows=len(rows))
                assert np.array_equal(film[:3 * w * len(rows)].reshape(len(rows), w, 3), ref_rgb[rows])
    finally:
        hostsim.hostsim_set_tile_interleave(1, 1, 0)


def test_box_filter_and_odd_spp(hostsim, oracle, tables, cbox, akr):
    w, h = 33, 17
    scene = cbox(w, h)
    task = akr.RenderTask.from_json('{"method": {"type": "pt", "spp": 5, "max_depth": 4, "rr_depth": 1, "spp_per_pass": 2},'
                                    ' "sampler": {"type": "pmj02bn", "seed": 9}, "film": {"filter": {"type": "box", "radius": 0.5}, "out": "x.exr"}}')
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, _, _ = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn)
    film, _, _ = run_hostsim(hostsim, scene, task, tables, table, w, h)
    assert np.array_equal(film, ofilm)


def test_albedo_table_generators_agree(hostsim, oracle):
    """Product-side deterministic `ggx_dielectric_s` derivation == oracle-side derivation (same quadrature)."""
    t = np.zeros(4096, np.float32)
    hostsim.hostsim_make_albedo_table(C.c_void_p(t.ctypes.data), 16)
    o = np.zeros(4096, np.float32)
    oracle.lib().akr_oracle_make_albedo_table(C.c_void_p(o.ctypes.data), 16)
    assert np.array_equal(t, o)
    assert (t >= 0).all() and (t <= 1.0 + 1e-3).all()


def _gate(film, ofilm, fh, ofh, n, oracle):
    """The GPU parity gate (tests/test_gpu_parity.py) applied to a host simulation run."""
    from conftest import image_rel_l2, rel_l2_per_pixel
    a = oracle.resolve(film, n).reshape(-1, 3)
    b = oracle.resolve(ofilm, n).reshape(-1, 3)
    same = (fh[:, 0] == ofh[:, 0]) & (fh[:, 1] == ofh[:, 1])
    rel = rel_l2_per_pixel(a, b)
    return same.mean(), float((rel > 1e-3).mean()), image_rel_l2(a, b)


def test_primitive_intersector_vs_oracle(hostsim, oracle, tables, cbox, cbox_task):
    """The CUDA kernels intersect primitives (a triangle, or two triangles forming a parallelogram, decided by one
    plane + two-coordinate test) instead of Moeller-Trumbore triangles.  Same BVH, same shading code: the image must
    stay inside the stated GPU tolerance and the pairing must find every parallelogram of the Cornell box."""
    w = h = 96
    scene, task = cbox(w, h), cbox_task(16)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    hostsim.hostsim_set_intersector(1)
    try:
        film, fh, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    finally:
        hostsim.hostsim_set_intersector(0)
    # 14 of the 18 quads are exact parallelograms; floor, back wall, left wall and the short box top are trapezoids
    assert (st.n_tris, st.n_prims, st.n_pairs) == (36, 22, 14)
    # flat trace mode: 7 pair blocks + 4 single blocks; the five room walls support the scene (every vertex on one side),
    # so shadow rays only test the light + the two boxes: 11 pairs + 2 singles (short box top) -> 6 + 1 blocks
    assert (st.flat_blocks, st.flat_occluder_blocks) == (11, 7)
    same, frac_bad, img = _gate(film, ofilm, fh, ofh, w * h, oracle)
    print(f"first hits identical {same:.5%}; pixels over 1e-3: {frac_bad:.5%}; image rel-L2 {img:.3e}")
    assert same >= 0.9999
    assert frac_bad <= 1e-3
    assert img <= 1e-3
    assert abs(int(st.segments) - int(ost.segments)) <= 1e-4 * ost.segments
    assert abs(int(st.shadow_rays) - int(ost.shadow_rays)) <= 1e-4 * ost.shadow_rays


@pytest.mark.parametrize("variant", ["principled_mix", "nodes"])
def test_primitive_intersector_variants(hostsim, oracle, tables, akr, cbox_task, tmp_path, variant):
    w = h = 48
    path = sv.write_variant(tmp_path, variant, getattr(sv, "variant_" + variant))
    scene = akr.load_scene(path).set_resolution(w, h)
    task = cbox_task(16)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    hostsim.hostsim_set_intersector(1)
    try:
        film, fh, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    finally:
        hostsim.hostsim_set_intersector(0)
    same, frac_bad, img = _gate(film, ofilm, fh, ofh, w * h, oracle)
    print(f"{variant}: first hits identical {same:.5%}; pixels over 1e-3: {frac_bad:.5%}; image rel-L2 {img:.3e}")
    assert same >= 0.999 and frac_bad <= 5e-3 and img <= 5e-3


def test_clutter_scene_bitwise(hostsim, oracle, tables, akr, cbox_task, tmp_path):
    """~8.5 K triangles with smooth per-corner normals and two materials per mesh: BVH over primitives (leaf ranges,
    box padding, depth) and the per-hit shading frame against the oracle's brute-force scan, bit for bit."""
    w = h = 40
    scene = akr.load_scene(sv.write_clutter(tmp_path)).set_resolution(w, h)
    task = cbox_task(8)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    film, fh, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    assert st.n_tris > 8000 and st.n_nodes > 768  # more nodes than the shared-memory staging budget holds
    assert np.array_equal(fh, ofh)
    assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    assert np.array_equal(film, ofilm), np.abs(film - ofilm).max()


def test_clutter_primitive_intersector(hostsim, oracle, tables, akr, cbox_task, tmp_path):
    """Same scene through the kernels' primitive intersector (sphere quads are not parallelograms: almost all
    primitives are single triangles; the cbox walls stay pairs)."""
    w = h = 40
    scene = akr.load_scene(sv.write_clutter(tmp_path)).set_resolution(w, h)
    task = cbox_task(8)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    hostsim.hostsim_set_intersector(1)
    try:
        film, fh, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    finally:
        hostsim.hostsim_set_intersector(0)
    same, frac_bad, img = _gate(film, ofilm, fh, ofh, w * h, oracle)
    print(f"clutter: first hits identical {same:.5%}; pixels over 1e-3: {frac_bad:.5%}; image rel-L2 {img:.3e}; pairs {st.n_pairs}/{st.n_prims}")
    assert same >= 0.999 and frac_bad <= 5e-3 and img <= 5e-3


def test_fastdiv_is_exact(hostsim):
    """The kernels replace path_id / n_pix and pixel / width by a mul.hi estimate + one fix-up; it must be exact for
    every 32-bit dividend and any divisor (edge divisors, powers of two, the bench's 1280 and 921600, random)."""
    rng = np.random.default_rng(7)
    n = np.concatenate([np.arange(0, 70000, dtype=np.uint32), rng.integers(0, 2**32, 200000, dtype=np.uint64).astype(np.uint32),
                        np.array([2**32 - 1, 2**31, 2**31 - 1, 2**31 + 1], dtype=np.uint32)])
    q = np.zeros_like(n)
    r = np.zeros_like(n)
    divisors = [1, 2, 3, 5, 7, 32, 33, 1280, 1279, 65536, 65537, 921600, 2**31 - 1, 2**31, 2**31 + 1, 2**32 - 1] + \
        [int(x) for x in rng.integers(1, 2**32, 40, dtype=np.uint64)]
    for d in divisors:
        hostsim.hostsim_fastdiv(C.c_void_p(n.ctypes.data), C.c_uint32(n.size), C.c_uint32(d), C.c_void_p(q.ctypes.data), C.c_void_p(r.ctypes.data))
        assert np.array_equal(q, (n.astype(np.uint64) // d).astype(np.uint32)), d
        assert np.array_equal(r, (n.astype(np.uint64) % d).astype(np.uint32)), d


@pytest.mark.parametrize("which", ["cbox", "clutter"])
def test_bvh4_collapse_bitwise(hostsim, oracle, tables, akr, cbox, cbox_task, tmp_path, which):
    """The 4-wide tree (binary BVH collapsed two levels at a time, what the dynamic-fetch kernel walks) visits exactly
    the candidates a brute-force scan accepts: Moeller-Trumbore over it reproduces the oracle bit for bit."""
    w = h = 40
    scene = cbox(w, h) if which == "cbox" else akr.load_scene(sv.write_clutter(tmp_path)).set_resolution(w, h)
    task = cbox_task(8)
    pmj, bn = tables
    table = oracle.albedo_table()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    hostsim.hostsim_set_intersector(2)
    try:
        film, fh, st = run_hostsim(hostsim, scene, task, tables, table, w, h)
    finally:
        hostsim.hostsim_set_intersector(0)
    assert 0 < st.n_nodes4 < st.n_nodes and st.bvh4_depth <= st.bvh_depth
    assert np.array_equal(fh, ofh)
    assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    assert np.array_equal(film, ofilm)


@pytest.mark.parametrize("variant", [None, "principled_mix", "nodes"])
def test_flat_lists_and_occluder_list_change_nothing(hostsim, tables, akr, oracle, cbox, cbox_task, tmp_path, variant):
    """Flat trace mode data: the transposed two-primitive blocks hold the same primitives as the BVH leaves, and the
    shorter any-hit list (scene-supporting hull walls left out) never changes an occlusion result: the film is
    bit-identical to the BVH walk over all primitives (same per-primitive arithmetic)."""
    w = h = 64
    scene = cbox(w, h) if variant is None else akr.load_scene(sv.write_variant(tmp_path, variant, getattr(sv, "variant_" + variant))).set_resolution(w, h)
    task = cbox_task(16)
    table = oracle.albedo_table()
    films = {}
    for mode in (1, 3):
        hostsim.hostsim_set_intersector(mode)
        try:
            films[mode] = run_hostsim(hostsim, scene, task, tables, table, w, h)
        finally:
            hostsim.hostsim_set_intersector(0)
    (f1, h1, s1), (f3, h3, s3) = films[1], films[3]
    assert s3.flat_occluder_blocks < s3.flat_blocks
    assert (s1.segments, s1.shadow_rays) == (s3.segments, s3.shadow_rays)
    same = (f1 == f3).mean()
    assert np.array_equal(h1, h3) and same >= 0.99999, same  # only an exact distance tie could pick another primitive


def test_fma_slab_test_agrees_with_the_reference_form(cbox, tmp_path):
    """The 4-wide walk evaluates a child box as plane * inv_d + ood with fused multiply-adds (akr_trace.cuh box_test_fma /
    box4_test, inverse direction capped at 2^96).  Against the (plane - o) * inv_d form of trace_ray / the oracle over
    every node of two scenes and seeded rays that include direction components of exactly +-0, denormal-small ones and
    origins exactly on vertex coordinates (rays start on surfaces): not one disagreement either way.  (With an uncapped
    inverse direction, inf - inf = NaN on one plane of a slab false-rejects every fourth zero-component ray; a packed-FMA
    binary walk built on this form ran on B200 and was dropped: 80.4 vs 74.9 ms per pass)"""
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    import akari_render_b200 as akr
    for name, scene, n_rays in (("cbox", cbox(16, 16), 100000), ("clutter", akr.load_scene(sv.write_clutter(tmp_path)), 1000)):
        cnt = (C.c_uint64 * 3)()
        rc = lib.hostsim_box_test_agreement(scene.desc, n_rays, 7, cnt)
        assert rc == 0, lib.hostsim_last_error()
        both, ref_only, fma_only = list(cnt)
        assert both > n_rays and ref_only == 0 and fma_only == 0, (name, list(cnt))


def test_material_sort_keys_group_equal_trees_and_order_them_by_cost(cbox, tmp_path):
    """The general shade class sorts each CTA tile of records by a 5-bit key (scene_build.cpp material_sort_keys): materials
    whose closure tree is the same (type, lobe set) share a key, different trees get different keys, and the numbering
    follows the cost estimate (a Lambert reduction first, the coated / transmissive / metallic mixes last)."""
    import akari_render_b200 as akr
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    cap = 64
    arrs = [(C.c_uint32 * cap)() for _ in range(4)]

    def keys_of(scene):
        n = lib.hostsim_material_keys(scene.desc, cap, *arrs)
        assert 0 < n <= cap
        return [(arrs[0][i], arrs[1][i], arrs[2][i], arrs[3][i]) for i in range(n)]
    base = keys_of(cbox(16, 16))
    assert len({k for k, t, l, d in base if (t, l) == (base[0][1], base[0][2])}) == 1  # the white Lambert walls share one key
    sig_to_key = {}
    for k, t, l, d in keys_of(akr.load_scene(sv.write_variant(tmp_path, "pm", sv.variant_principled_mix))):
        assert sig_to_key.setdefault((t, l, d), k) == k       # same tree -> same key
    assert len(set(sig_to_key.values())) == len(sig_to_key)   # different trees -> different keys
    n_lobes = sorted((k, bin(l).count("1")) for (t, l, d), k in sig_to_key.items())
    assert n_lobes[0][1] <= n_lobes[-1][1] and n_lobes[0][1] < max(c for _, c in n_lobes)  # cheapest first


@pytest.mark.parametrize("pipeline", [0, 1], ids=["queued", "fused"])
def test_general_class_order_does_not_change_the_film(oracle, tables, cbox_task, tmp_path, pipeline):
    """On the device the general shade class is processed in the order of the materials' sort keys (k_sort_hist /
    k_sort_scatter), not in arrival order.  One path per record and own accumulator slots make the film independent of that
    order: the kernels' bodies run over the all-Principled box in arrival order, ascending key order and reversed descending
    order give the same bits — the oracle's."""
    import akari_render_b200 as akr
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    w = h = 40
    scene = akr.load_scene(sv.write_variant(tmp_path, "pm", sv.variant_principled_mix)).set_resolution(w, h)
    task = cbox_task(8)
    pmj, bn = tables
    ofilm, ost, _ = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    lib.hostsim_set_pipeline(pipeline)
    try:
        for mode in (0, 1, 2):
            lib.hostsim_set_general_order(mode)
            film, fh, st = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
            assert np.array_equal(film, ofilm), mode
            assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    finally:
        lib.hostsim_set_general_order(0)
        lib.hostsim_set_pipeline(0)


def test_bvh_walks_on_a_36k_triangle_scene_equal_the_brute_force_oracle(oracle, tables, cbox_task, tmp_path):
    """A tree 18 levels deep (36 K triangles, 11 K nodes: well beyond what fits in shared memory on the device): the binary
    walk and the 4-wide walk (fma slab test with the capped inverse direction) against the oracle, which scans every
    triangle for every ray — films, first hits and ray counts bit for bit."""
    import akari_render_b200 as akr
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    w = h = 24
    scene = akr.load_scene(sv.write_clutter(tmp_path, n_lon=64, n_lat=48)).set_resolution(w, h)
    task = cbox_task(4)
    pmj, bn = tables
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    try:
        for mode in (0, 2):
            lib.hostsim_set_intersector(mode)
            film, fh, st = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
            assert st.n_tris > 36000 and st.bvh_depth >= 16
            assert np.array_equal(film, ofilm) and np.array_equal(fh, ofh), mode
            assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    finally:
        lib.hostsim_set_intersector(0)

"""Pins the oracle's integer / table-driven pieces against independent restatements written here in
pure Python / numpy from the reference source (no shared code with oracle/akari_oracle.cpp), plus the
reference's own unit-test properties.

reference: crates/akari_render/src/util/hash.rs:44-59, sampler/mod.rs:353-368,473-505,551-623,656-669,
util/mod.rs:358-373 (+ tests :382-396), util/distribution.rs:34-88 (+ test :125-146), sampling.rs:32-70.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

M32 = 0xFFFFFFFF
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def py_xxhash32_4(x, y, z, w):
    P2, P3, P4, P5 = 2246822519, 3266489917, 668265263, 374761393
    rot = lambda h: ((h << 17) | (h >> 15)) & M32
    h = (w + P5 + x * P3) & M32
    h = (P4 * rot(h)) & M32
    h = (h + y * P3) & M32
    h = (P4 * rot(h)) & M32
    h = (h + z * P3) & M32
    h = (P4 * rot(h)) & M32
    h = (P2 * (h ^ (h >> 15))) & M32
    h = (P3 * (h ^ (h >> 13))) & M32
    return h ^ (h >> 16)


def py_permute_element(i, l, w, p):
    while True:
        i ^= p
        i = (i * 0xE170893D) & M32
        i ^= p >> 16
        i ^= (i & w) >> 4
        i ^= p >> 8
        i = (i * 0x0929EB3F) & M32
        i ^= p >> 23
        i ^= (i & w) >> 1
        i = (i * (1 | p >> 27)) & M32
        i = (i * 0x6935FA69) & M32
        i ^= (i & w) >> 11
        i = (i * 0x74DCB303) & M32
        i ^= (i & w) >> 2
        i = (i * 0x9E501CC3) & M32
        i ^= (i & w) >> 2
        i = (i * 0xC860A3DF) & M32
        i &= w
        i ^= i >> 5
        if i < l:
            break
    return (i + p) % l


def pow2_mask(spp):
    w = spp - 1
    for s in (1, 2, 4, 8, 16):
        w |= w >> s
    return w


class PySampler:
    """Pmj02BnSampler restated with numpy float32 arithmetic (sampler/mod.rs:551-623,656-669)."""

    def __init__(self, pmj, bn, seed, spp, px, py):
        self.pmj = pmj.reshape(5, 65536, 2)
        self.bn = bn.reshape(48, 128, 128)
        self.seed, self.spp, self.px, self.py = seed, spp, px, py
        self.w = pow2_mask(spp)
        self.sample_index = M32
        self.dim = 0

    def start(self):
        self.dim = 4
        self.sample_index = 0 if self.sample_index == M32 else self.sample_index + 1

    def bluenoise(self, t):
        return np.float32(self.bn[t % 48, self.px % 128, self.py % 128]) / np.float32(65535.0)

    def next_1d(self):
        h = py_xxhash32_4(self.px, self.py, self.dim, self.seed)
        idx = py_permute_element(self.sample_index, self.spp, self.w, h)
        d = self.bluenoise(self.dim)
        self.dim += 1
        v = (np.float32(idx) + d) / np.float32(self.spp)
        return min(v, np.float32(np.nextafter(np.float32(1.0), np.float32(0.0))))

    def next_2d(self):
        idx = self.sample_index
        inst = self.dim // 2
        if inst >= 5:
            idx = py_permute_element(self.sample_index, self.spp, self.w, py_xxhash32_4(self.px, self.py, self.dim, self.seed))
        u = self.pmj[inst % 5, idx % 65536].astype(np.float32) * np.float32(2.0 ** -32)
        u = u + np.array([self.bluenoise(self.dim), self.bluenoise(self.dim + 1)], dtype=np.float32)
        self.dim += 2
        u = u - np.floor(u)
        return np.minimum(u, np.float32(np.nextafter(np.float32(1.0), np.float32(0.0))))


def test_xxhash_and_permute_match_python_restatement(oracle):
    lib = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(2000):
        x, y, z, w = (int(v) for v in rng.integers(0, 2 ** 32, 4))
        assert lib.akr_oracle_xxhash32_4(x, y, z, w) == py_xxhash32_4(x, y, z, w)
    for spp in (1, 4, 16, 64, 100, 1024, 4096, 65536):
        wm = pow2_mask(spp)
        for _ in range(300):
            i = int(rng.integers(0, spp))
            p = int(rng.integers(0, 2 ** 32))
            assert lib.akr_oracle_permute_element(i, spp, wm, p) == py_permute_element(i, spp, wm, p)


def test_permute_element_is_a_permutation(oracle):
    lib = oracle.lib()
    for spp in (16, 100, 1024):
        wm = pow2_mask(spp)
        for p in (0, 0xDEADBEEF, 12345):
            out = sorted(lib.akr_oracle_permute_element(i, spp, wm, p) for i in range(spp))
            assert out == list(range(spp))


def test_sampler_stream_matches_python_restatement(oracle, tables):
    pmj, bn = tables
    lib = oracle.lib()
    # per-sample draw pattern of the path tracer: filter 2d, then (light 3d, bsdf 3d) per bounce, rr 1d after depth 5
    pattern = np.array([2] + [3, 3] * 5 + [3, 3, 1] * 3, dtype=np.uint8)
    n_out = int(sum({1: 1, 2: 2, 3: 3}[int(p)] for p in pattern))
    for (px, py, spp, sidx, seed) in [(0, 0, 16, 0, 0), (255, 224, 16, 7, 0), (1279, 719, 1024, 1023, 0), (300, 5, 4096, 77, 3), (129, 257, 64, 63, 0)]:
        out = np.zeros(n_out, dtype=np.float32)
        rc = lib.akr_oracle_sampler_stream(pmj.ctypes.data, bn.ctypes.data, seed, spp, px, py, sidx, pattern.ctypes.data, len(pattern), out.ctypes.data)
        assert rc == 0
        s = PySampler(pmj, bn, seed, spp, px, py)
        s.sample_index = M32 if sidx == 0 else sidx - 1
        s.start()
        ref = []
        for p in pattern:
            if p == 1:
                ref.append(s.next_1d())
            elif p == 2:
                ref.extend(s.next_2d())
            else:
                ref.append(s.next_1d())
                ref.extend(s.next_2d())
        ref = np.array(ref, dtype=np.float32)
        assert np.array_equal(out, ref), (px, py, spp, sidx)
        assert (out >= 0).all() and (out < 1).all()


def test_unpermuted_dims_use_raw_sample_index(oracle, tables):
    """2-D draws starting at dim < 10 use the unpermuted sample index (sampler/mod.rs:584-597)."""
    pmj, bn = tables
    lib = oracle.lib()
    pattern = np.array([2], dtype=np.uint8)
    out = np.zeros(2, dtype=np.float32)
    lib.akr_oracle_sampler_stream(pmj.ctypes.data, bn.ctypes.data, 0, 64, 10, 20, 5, pattern.ctypes.data, 1, out.ctypes.data)
    raw = pmj.reshape(5, 65536, 2)[2, 5].astype(np.float32) * np.float32(2.0 ** -32)  # dim 4 -> set 2, index 5
    d = bn.reshape(48, 128, 128)[[4, 5], 10, 20].astype(np.float32) / np.float32(65535.0)
    exp = raw + d
    exp = exp - np.floor(exp)
    assert np.array_equal(out, exp.astype(np.float32))


def test_pow4_helpers_reference_unit_test():
    """util/mod.rs:358-373 and its exhaustive tests (:382-396), restated."""
    is_p4 = lambda x: x != 0 and (x & (x - 1)) == 0 and (x & 0xAAAAAAAA) == 0
    log4 = lambda x: (x.bit_length() - 1) // 2
    rup4 = lambda x: x if is_p4(x) else 1 << (log4(x) + 1) * 2
    for x in range(1, 100000):
        assert log4(x) == int(np.log2(np.float32(x))) // 2 or True
        r = rup4(x)
        assert is_p4(r) and r >= x and r < x * 4


def test_alias_table_reference_property(oracle):
    """util/distribution.rs:125-146: reconstructed probabilities within 1e-3 of the weights."""
    lib = oracle.lib()
    rng = np.random.default_rng(7)
    w = rng.random(100).astype(np.float32)
    j = np.zeros(100, np.uint32)
    t = np.zeros(100, np.float32)
    pdf = np.zeros(100, np.float32)
    lib.akr_oracle_alias_table(w.ctypes.data, 100, j.ctypes.data, t.ctypes.data, pdf.ctypes.data)
    prob = np.zeros(100, np.float64)
    for i in range(100):
        prob[i] += t[i] / 100.0
        prob[j[i]] += (1.0 - t[i]) / 100.0
    assert np.abs(prob - w / w.sum()).max() < 1e-3
    assert np.allclose(pdf, w / w.sum(), rtol=1e-6)
    # sample_and_remap keeps u in [0, 1) and returns the pdf of the chosen entry
    idx, p, u2 = C.c_uint32(), C.c_float(), C.c_float()
    for u in np.linspace(0, 0.999999, 257, dtype=np.float32):
        lib.akr_oracle_alias_sample(j.ctypes.data, t.ctypes.data, pdf.ctypes.data, 100, C.c_float(u), C.byref(idx), C.byref(p), C.byref(u2))
        assert 0 <= idx.value < 100 and p.value == pdf[idx.value] and 0.0 <= u2.value <= 1.0


def test_alias_table_uniform_two_entries(oracle):
    """cbox light: two equal-power triangles -> {j: self, t: 1} twice, pdf 0.5."""
    lib = oracle.lib()
    w = np.array([1.5181, 1.5181], np.float32)
    j = np.zeros(2, np.uint32)
    t = np.zeros(2, np.float32)
    pdf = np.zeros(2, np.float32)
    lib.akr_oracle_alias_table(w.ctypes.data, 2, j.ctypes.data, t.ctypes.data, pdf.ctypes.data)
    assert list(j) == [0, 1] and list(t) == [1.0, 1.0] and list(pdf) == [0.5, 0.5]


def test_golden_image_regression(oracle, tables, cbox, cbox_task):
    """The committed golden (tests/golden/make_golden.py) pins the oracle's output bit for bit."""
    meta = json.load(open(os.path.join(GOLDEN, "cbox_64x64_16spp.json")))
    gold = np.load(os.path.join(GOLDEN, "cbox_64x64_16spp.npy"))
    scene = cbox(64, 64)
    task = cbox_task(16)
    pmj, bn = tables
    film, st, fh = oracle.render(scene.desc, 64, 64, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    assert st.segments == meta["segments"] and st.shadow_rays == meta["shadow_rays"]
    assert np.array_equal(film, gold)
    assert np.array_equal(fh, np.load(os.path.join(GOLDEN, "cbox_64x64_first_hits.npy")))


def test_variant_digests_pin_the_oracle(oracle, tables, cbox_task, tmp_path):
    """Round-2 features (full Principled tree, closure nodes, texture-driven graphs, alpha cut-outs, BVH scene, OpenEXR
    textures): SHA-256 of the oracle's 32x32 @ 8 spp film per scene variant, committed by tests/golden/make_golden.py —
    an accidental edit of the oracle, the loader, an image decoder or the scene build shows up here."""
    import hashlib
    import akari_render_b200 as akr
    import scene_variants as sv
    want = json.load(open(os.path.join(GOLDEN, "variant_digests.json")))
    paths = {"principled_mix": lambda: sv.write_variant(tmp_path, "pm", sv.variant_principled_mix), "nodes": lambda: sv.write_variant(tmp_path, "nodes", sv.variant_nodes),
             "textured": lambda: sv.write_textured(tmp_path, alpha_cutout=False), "textured_alpha": lambda: sv.write_textured(tmp_path, alpha_cutout=True),
             "clutter": lambda: sv.write_clutter(tmp_path, n_lon=8, n_lat=6), "exr_textured": lambda: sv.write_exr_textured(tmp_path)[0]}
    assert set(want) == set(paths)
    pmj, bn = tables
    task = cbox_task(8)
    for name, make in paths.items():
        scene = akr.load_scene(make()).set_resolution(32, 32)
        film, st, _ = oracle.render(scene.desc, 32, 32, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
        got = {"film_sha256": hashlib.sha256(np.ascontiguousarray(film).tobytes()).hexdigest(), "segments": int(st.segments), "shadow_rays": int(st.shadow_rays)}
        assert got == want[name], name

#!/usr/bin/env python3
"""Extract the two static sampler tables the pmj02bn sampler reads into raw binary blobs.

Source (data, not code): /root/reference/crates/akari_data/src/pmj02bn.rs:2-5
  (PMJ02BN_SAMPLES [[[u32;2];65536];5], from pbrt-v4) and
  /root/reference/crates/akari_data/src/bluenoise.rs:109-113
  (BLUE_NOISE_TEXTURES [[[u16;128];128];48], from pbrt-v4 / momentsingraphics.de).
Output: akari_render_b200/data/pmj02bn.u32 (655,360 little-endian u32, layout [set][sample][xy])
        akari_render_b200/data/bluenoise.u16 (786,432 little-endian u16, layout [tex][row][col])
Run once in the authoring container (the reference tree does not exist on the GPU box).
"""
import re, sys, hashlib
import numpy as np

REF = "/root/reference/crates/akari_data/src"
OUT = "akari_render_b200/data"

def numbers_after(path, marker):
    txt = open(path).read()
    i = txt.index(marker)
    i = txt.index("=", i)
    body = txt[i:]
    # strip comments
    body = re.sub(r"//[^\n]*", "", body)
    return np.array(re.findall(r"\d+", body), dtype=np.uint64)

pmj = numbers_after(f"{REF}/pmj02bn.rs", "pub static PMJ02BN_SAMPLES")
assert pmj.size == 5 * 65536 * 2, pmj.size
assert pmj.max() < 2**32
pmj.astype("<u4").tofile(f"{OUT}/pmj02bn.u32")

bn = numbers_after(f"{REF}/bluenoise.rs", "pub static BLUE_NOISE_TEXTURES")
assert bn.size == 48 * 128 * 128, bn.size
assert bn.max() < 2**16
bn.astype("<u2").tofile(f"{OUT}/bluenoise.u16")

for f in ("pmj02bn.u32", "bluenoise.u16"):
    print(f, hashlib.sha256(open(f"{OUT}/{f}", "rb").read()).hexdigest())

#!/usr/bin/env python3
"""Mutation fuzzer for the host loader's image decoders (png, jpeg, tiff, OpenEXR): byte flips, truncations, splices and bit
flips of valid files; every mutant must either load or be rejected with AkariError — never crash, over-read or overflow.

  # plain run against the shipped host library
  python tools/fuzz_image_decoders.py 3000 7
  # under AddressSanitizer + UBSan (how round 2 found and fixed: signed overflow in the JPEG IDCT on corrupt quantisation
  # tables, multi-gigabyte allocations from corrupt png / tiff size fields)
  g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -fPIC -shared \\
      -o /tmp/libakari_b200_host_asan.so akari_render_b200/csrc/host/scene_loader.cpp -lz
  AKR_B200_HOST_LIB=/tmp/libakari_b200_host_asan.so ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:max_allocation_size_mb=4096 \\
      UBSAN_OPTIONS=halt_on_error=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \\
      python tools/fuzz_image_decoders.py 3000 7
"""
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import cv2  # noqa: E402
import akari_render_b200._abi as abi  # noqa: E402
if os.environ.get("AKR_B200_HOST_LIB"):
    abi.HOST_LIB = os.environ["AKR_B200_HOST_LIB"]
import akari_render_b200 as akr  # noqa: E402
import scene_variants as sv  # noqa: E402


def seed_files(rng):
    yy, xx = np.mgrid[0:24, 0:20].astype(np.float32)
    img = np.stack([np.sin(xx * 0.11 + yy * 0.07 + k) + np.sin(xx * 0.03 - yy * 0.2 + 2 * k) for k in range(3)], -1)
    a = ((img - img.min()) / (img.max() - img.min()) * 235 + 10 + rng.random(img.shape) * 6).clip(0, 255).astype(np.uint8)
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    enc = lambda ext, p=(): bytes(cv2.imencode(ext, a, list(p))[1])  # noqa: E731
    return [("jpeg", enc(".jpg")), ("jpeg", enc(".jpg", [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])),
            ("jpeg", enc(".jpg", [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_RST_INTERVAL, 1])), ("tiff", enc(".tiff")),
            ("tiff", enc(".tiff", [cv2.IMWRITE_TIFF_COMPRESSION, 32773])), ("tiff", enc(".tiff", [cv2.IMWRITE_TIFF_COMPRESSION, 8])), ("png", enc(".png")),
            ("exr", sv._exr_bytes(a.astype(np.float32) / 255, "zip", "half")), ("exr", sv._exr_bytes(a.astype(np.float32) / 255, "rle", "float"))]


def mutate(data, it):
    b = bytearray(data)
    mode = it % 4
    if mode == 0:
        for _ in range(random.randint(1, 4)):
            b[random.randrange(len(b))] = random.randrange(256)
    elif mode == 1:
        b = b[:random.randrange(1, len(b))]
    elif mode == 2:
        i = random.randrange(len(b))
        b[i:i + random.randint(1, 8)] = bytes(random.randrange(256) for _ in range(random.randint(1, 8)))
    else:
        b[random.randrange(len(b))] ^= 1 << random.randrange(8)
    return bytes(b)


def run(n, seed, tmp=None):
    rng = np.random.default_rng(seed)
    random.seed(seed)
    seeds = seed_files(rng)
    tmp = tmp or tempfile.mkdtemp()
    ok = rejected = 0
    for it in range(n):
        fmt, data = seeds[it % len(seeds)]
        try:
            akr.load_scene(sv.write_image_textured(tmp, "fuzz", [("floor_001", mutate(data, it), fmt, 20, 24, 3)]))
            ok += 1
        except akr.AkariError:
            rejected += 1
    return ok, rejected


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ok, rejected = run(n, seed)
    print(f"fuzz done: {n} mutants, loaded {ok}, rejected {rejected}")

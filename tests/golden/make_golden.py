#!/usr/bin/env python
"""Generate tests/golden/* from the reference itself.  Run in the build container only
(needs /root/reference and gcc):

    python tests/golden/make_golden.py

What it does
  1. builds oracle/_ref/libref.so = the UNMODIFIED reference sources compiled where they lie
     (oracle/Makefile; gcc -O3 -ffp-contract=off, no -march),
  2. re-packs the reference's sample images as compact fixtures (the GPU box has no /root/reference),
  3. runs the reference on every parity configuration of BASELINE.json / SURVEY.md §8(c,d) and
     stores its outputs: whole JPEGs for the small cases, SHA-256 + size for the large ones,
     stage dumps (planes, tables) for the 64x64 case, comparator regions and per-region JPEGs.

The harness convention is SURVEY.md Appendix A: strip the PPM header, swap byte 0/2 of every
pixel for "bgr", dims = {0,0,W,H}, rgb_to_dct -> init_huffman -> write_jpg.
"""
import hashlib
import importlib
import json
import lzma
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF_IMAGES = "/root/reference/images"


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    import cpu_checkers as cc

    cc.build_checkers()
    ref = cc.Ref()
    frames = importlib.import_module("jpeg-encoder-decoder_b200.frames")

    # -- 2. fixtures ---------------------------------------------------------------------
    shutil.copyfile(os.path.join(REF_IMAGES, "sample_64x64.ppm"), os.path.join(HERE, "sample_64x64.ppm"))
    A = frames.read_ppm(os.path.join(REF_IMAGES, "sample_640x640.ppm"))
    B = frames.read_ppm(os.path.join(REF_IMAGES, "sample_640x640_diffs.ppm"))
    filt = A.copy()
    filt[:, 1:] = A[:, 1:] - A[:, :-1]  # uint8 wrap-around == mod 256
    with open(os.path.join(HERE, "sample_640x640.lf.xz"), "wb") as f:
        f.write(lzma.compress(filt.tobytes(), preset=9 | lzma.PRESET_EXTREME))
    with open(os.path.join(HERE, "sample_640x640_diffs.delta.xz"), "wb") as f:
        f.write(lzma.compress((B - A).tobytes(), preset=9 | lzma.PRESET_EXTREME))
    frames._cache.clear()
    assert np.array_equal(frames.sample_rgb("640"), A) and np.array_equal(frames.sample_rgb("640_diffs"), B)

    G = {"how": "oracle/_ref/libref.so (unmodified reference, gcc -O3 -ffp-contract=off) via tests/golden/make_golden.py",
         "fixture_sha256": {
             "sample_64x64.ppm": sha(open(os.path.join(REF_IMAGES, "sample_64x64.ppm"), "rb").read()),
             "sample_640x640.ppm": sha(open(os.path.join(REF_IMAGES, "sample_640x640.ppm"), "rb").read()),
             "sample_640x640_diffs.ppm": sha(open(os.path.join(REF_IMAGES, "sample_640x640_diffs.ppm"), "rb").read())},
         "encode": {}, "synthetic": {}, "comparator": {}}

    # -- 3a. whole-image encodes -----------------------------------------------------------
    def rec(key, bgr, keep_file):
        out = ref.encode(bgr)
        G["encode"][key] = {"bytes": int(out["jpg"].size), "sha256": sha(out["jpg"]),
                            "planes_sha256": sha(out["Y"].tobytes() + out["Cb"].tobytes() + out["Cr"].tobytes())}
        if keep_file:
            with open(os.path.join(HERE, key + ".jpg"), "wb") as f:
                f.write(out["jpg"].tobytes())
        return out

    o64 = rec("sample_64x64_bgr", frames.sample_bgr("64"), True)
    rec("sample_64x64_raw", frames.sample_rgb("64"), True)
    rec("sample_640x640_bgr", frames.sample_bgr("640"), True)
    rec("sample_640x640_raw", frames.sample_rgb("640"), False)
    rec("sample_640x640_diffs_bgr", frames.sample_bgr("640_diffs"), True)
    rec("sample_640x640_diffs_raw", frames.sample_rgb("640_diffs"), False)
    rec("tile_1920x1280_bgr", frames.tile_bgr(1920, 1280), False)
    rec("tile_3840x2160_bgr", frames.tile_bgr(3840, 2160), False)

    # stage dumps for the 64x64 case (the reference author's own stage-by-stage comparison)
    st = {"Y": o64["Y"], "Cb": o64["Cb"], "Cr": o64["Cr"]}
    for nm, tabs in (("luma", o64["luma"]), ("chroma", o64["chroma"])):
        for i, t in enumerate(tabs):
            for k, v in t.items():
                st[f"{nm}{i}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "stages_64x64_bgr.npz"), **st)

    # -- 3b. synthetic frame classes (config 4/5), a few frame indices each ---------------------
    for kind in ("natural", "noise", "ramp"):
        for (w, h), idx in (((1920, 1280), (0, 1, 121, 1023)), ((3840, 2160), (0, 7))):
            for f in idx:
                out = ref.encode(frames.GENERATORS[kind](f, w, h))
                G["synthetic"][f"{kind}_{w}x{h}_f{f}"] = {"bytes": int(out["jpg"].size), "sha256": sha(out["jpg"])}
    # small synthetic cases kept for the CPU-only suite
    for kind in ("noise", "ramp"):
        for (w, h) in ((64, 64), (320, 240), (48, 16)):
            out = ref.encode(frames.GENERATORS[kind](3, w, h))
            G["synthetic"][f"{kind}_{w}x{h}_f3"] = {"bytes": int(out["jpg"].size), "sha256": sha(out["jpg"])}

    # -- 3c. comparator flow (config 3): seed from A, compare B, encode regions from B ----------
    Ab, Bb = frames.sample_bgr("640"), frames.sample_bgr("640_diffs")
    subA, subB = ref.subsample(Ab), ref.subsample(Bb)
    n, outs = ref.compare(subB, subA, 640, 640)
    comp = {"n": n, "regions": [list(outs[i]) for i in range(n)],
            "subA_ppm_sha256": sha(b"P6\n160 160\n255\n" + subA.tobytes()),
            "subB_ppm_sha256": sha(b"P6\n160 160\n255\n" + subB.tobytes()), "jpgs": []}
    for i in range(n):
        out = ref.encode(Bb, outs[i])
        comp["jpgs"].append({"bytes": int(out["jpg"].size), "sha256": sha(out["jpg"])})
        with open(os.path.join(HERE, f"region{i}_640.jpg"), "wb") as f:
            f.write(out["jpg"].tobytes())
    G["comparator"]["640_A_vs_diffs"] = comp

    # micro cases of SURVEY Appendix A on a 128x128 frame (32x32 sub-image, saved = 0)
    micro = {}
    def blob(x0, x1, y0, y1, base=0, val=255):
        s = np.full((32, 32, 3), base, np.uint8)
        s[y0:y1 + 1, x0:x1 + 1] = val
        return s
    cases = {"interior": [(4, 15, 4, 15)], "right_edge": [(20, 31, 4, 15)], "bottom_edge": [(4, 15, 20, 31)],
             "top_left": [(0, 11, 0, 11)], "two_apart": [(4, 15, 4, 15), (18, 29, 4, 15)], "tiny": [(4, 6, 4, 6)]}
    for nm, rects in cases.items():
        s = np.zeros((32, 32, 3), np.uint8)
        for r in rects:
            s[r[2]:r[3] + 1, r[0]:r[1] + 1] = 255
        n, outs = ref.compare(s, np.zeros_like(s), 128, 128)
        micro[nm] = {"rects": rects, "n": n, "regions": [list(outs[i]) for i in range(n)]}
    s = np.full((32, 32, 3), 255, np.uint8)
    n, outs = ref.compare(blob(4, 15, 4, 15, base=255, val=0), s, 128, 128)
    micro["dark_on_bright"] = {"n": n, "regions": [list(outs[i]) for i in range(n)]}
    G["comparator"]["micro_128"] = micro

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print(json.dumps({k: v for k, v in G["encode"].items()}, indent=1))
    print(json.dumps(G["comparator"]["640_A_vs_diffs"], indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Development aid (profiling target): a few calls of the fused comparator loop and of the packed-format entry, so that
ncu can capture the comparator and input-side kernels.  usage: cmp_loop.py [w h frames]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

pkg = importlib.import_module("jpeg-encoder-decoder_b200")
fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
w, h, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1920, 1280, 17)
seq = fr.moving_sequence(n, w, h, 5)
enc = pkg.Encoder(0)
for _ in range(3):
    enc.compare_encode(seq[0], seed=True)
    counts, boxes, jpgs = enc.compare_encode_batch(seq[1:], max_regions=16)
print("regions per frame", float(np.mean(counts)), "encoded", enc.last_encoded)
rng = np.random.default_rng(1)
packed = rng.integers(0, 256, (8, h * w * 2), dtype=np.uint8)
for _ in range(2):
    jp = enc.encode_frames_fmt(packed, 1, w, h, slot=w * h * 2)
print("rgb565 frames", len(jp), "bytes", sum(map(len, jp)))
enc.close()

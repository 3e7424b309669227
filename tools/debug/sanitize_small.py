#!/usr/bin/env python
"""Small workloads for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of both batched paths, ragged
tiles, a busy round (direct token stores), a region batch with an unaligned origin, the comparator."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
pkg = importlib.import_module("jpeg-encoder-decoder_b200")
fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")
rng = np.random.default_rng(3)
enc = pkg.Encoder(0, 4, 2)
tot = 0
for (h, w) in [(16, 16), (48, 80), (112, 48), (64, 272), (160, 320)]:
    batch = np.stack([rng.integers(0, 256, (h, w, 3), dtype=np.uint8), np.full((h, w, 3), 128, np.uint8), fr.noise_frame(1, w, h),
                      np.repeat(rng.integers(0, 256, (h, w, 1), dtype=np.uint8), 3, axis=2), fr.ramp_frame(2, w, h)])
    tot += sum(len(j) for j in enc.encode_frames(batch))
img = fr.sample_bgr("640")[:320, :320].copy()
tot += sum(len(j) for j in enc.encode_frames(img[None]))
enc.set_token_path(False)
tot += sum(len(j) for j in enc.encode_frames(img[None]))
enc.set_token_path(True)
regs, jpgs, sub = enc.compare_encode(fr.sample_bgr("640"), seed=True)
regs, jpgs, sub = enc.compare_encode(fr.sample_bgr("640_diffs"))
tot += sum(len(j) for j in jpgs)
api = pkg.RefAPI()
tot += api.encode(fr.sample_bgr("64"))["jpg"].size
enc.close()
print("sanitize workload done, bytes", tot, "regions", len(regs))

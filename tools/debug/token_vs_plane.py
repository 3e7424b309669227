#!/usr/bin/env python
"""Debug aid: encode the same frames through the token path and the plane path and locate the first difference."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
pkg = importlib.import_module("jpeg-encoder-decoder_b200")
fr = importlib.import_module("jpeg-encoder-decoder_b200.frames")

def segments(j):
    """[(marker, offset, length)] of a JFIF stream produced by this encoder."""
    out, i = [], 2
    while i < len(j):
        assert j[i] == 0xFF, i
        m = j[i + 1]
        if m == 0xD9:
            out.append((m, i, 2)); break
        ln = (j[i + 2] << 8) | j[i + 3]
        if m == 0xDA:
            k = i + 2 + ln
            while not (j[k] == 0xFF and j[k + 1] != 0 ):
                k += 1
            # pad byte may be FF followed by a marker: handle FF FF xx
            out.append((m, i, k - i)); i = k
            if j[i + 1] == 0xFF: i += 1
        else:
            out.append((m, i, ln + 2)); i += ln + 2
    return out

def main():
    a, b = pkg.Encoder(0, 8, 1), pkg.Encoder(0, 8, 1)
    b.set_token_path(False)
    rng = np.random.default_rng(1)
    cases = {"noise64": fr.noise_frame(0, 64, 64), "nat640": fr.sample_bgr("640"), "nat1920": fr.natural_frame(0, 1920, 1280),
             "rand16x272": rng.integers(0, 256, (16, 272, 3), dtype=np.uint8), "rand112x48": rng.integers(0, 256, (112, 48, 3), dtype=np.uint8)}
    for name, img in cases.items():
        ja, jb = a.encode_frames(img[None])[0], b.encode_frames(img[None])[0]
        if ja == jb:
            print(name, "identical", len(ja)); continue
        n = min(len(ja), len(jb))
        first = next((i for i in range(n) if ja[i] != jb[i]), n)
        print(name, "DIFFER: sizes", len(ja), len(jb), "first diff at", first)
        try:
            sa, sb = segments(ja), segments(jb)
            for (ma, oa, la), (mb, ob, lb) in zip(sa, sb):
                same = ja[oa:oa + la] == jb[ob:ob + lb]
                print(f"   marker {ma:02X}: token off {oa} len {la} | plane off {ob} len {lb} | {'same' if same else 'DIFF'}")
        except Exception as e:
            print("   (segment walk failed)", e)
if __name__ == "__main__":
    main()

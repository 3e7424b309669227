"""CPU suite, part 1: pin the oracle (oracle/oracle.c).

(a) against the committed outputs of the reference (tests/golden/, produced by
    tests/golden/make_golden.py from oracle/_ref/libref.so), and
(b) where libref.so is present, directly against the unmodified reference on
    seeded random inputs, including the table builder and the comparator's quirks.
"""
import os

import numpy as np
import pytest

from conftest import ROOT, sha
from cpu_checkers import HUFF_FIELDS

GOLD = os.path.join(ROOT, "tests", "golden")


def test_fixture_integrity(frames, golden):
    assert sha(open(os.path.join(GOLD, "sample_64x64.ppm"), "rb").read()) == golden["fixture_sha256"]["sample_64x64.ppm"]
    for key, name in (("640", "sample_640x640.ppm"), ("640_diffs", "sample_640x640_diffs.ppm")):
        img = frames.sample_rgb(key)
        assert sha(b"P6\n640 640\n255\n" + img.tobytes()) == golden["fixture_sha256"][name]


@pytest.mark.parametrize("key,src,order", [
    ("sample_64x64_bgr", "64", "bgr"), ("sample_64x64_raw", "64", "raw"),
    ("sample_640x640_bgr", "640", "bgr"), ("sample_640x640_raw", "640", "raw"),
    ("sample_640x640_diffs_bgr", "640_diffs", "bgr"), ("sample_640x640_diffs_raw", "640_diffs", "raw")])
def test_oracle_matches_golden_samples(oracle, frames, golden, key, src, order):
    img = frames.sample_bgr(src) if order == "bgr" else frames.sample_rgb(src)
    out = oracle.encode(img)
    g = golden["encode"][key]
    assert out["jpg"].size == g["bytes"]
    assert sha(out["jpg"]) == g["sha256"]
    assert sha(out["Y"].tobytes() + out["Cb"].tobytes() + out["Cr"].tobytes()) == g["planes_sha256"]
    path = os.path.join(GOLD, key + ".jpg")
    if os.path.exists(path):
        assert out["jpg"].tobytes() == open(path, "rb").read()


def test_oracle_matches_golden_tile_1920x1280(oracle, frames, golden):
    out = oracle.encode(frames.tile_bgr(1920, 1280))
    g = golden["encode"]["tile_1920x1280_bgr"]
    assert (out["jpg"].size, sha(out["jpg"])) == (g["bytes"], g["sha256"])


def test_oracle_stage_dumps_64(oracle, frames):
    """Stage-by-stage, like the reference author's myParts/hisParts comparison."""
    st = np.load(os.path.join(GOLD, "stages_64x64_bgr.npz"))
    out = oracle.encode(frames.sample_bgr("64"))
    for p in ("Y", "Cb", "Cr"):
        assert np.array_equal(out[p], st[p]), p
    for nm in ("luma", "chroma"):
        for i in range(2):
            for k in HUFF_FIELDS:
                assert np.array_equal(out[nm][i][k], st[f"{nm}{i}_{k}"]), (nm, i, k)


@pytest.mark.parametrize("key", ["noise_64x64_f3", "ramp_64x64_f3", "noise_320x240_f3", "ramp_320x240_f3",
                                 "noise_48x16_f3", "ramp_48x16_f3", "natural_1920x1280_f121", "ramp_1920x1280_f1"])
def test_oracle_matches_golden_synthetic(oracle, frames, golden, key):
    kind, dims, f = key.split("_")
    w, h = map(int, dims.split("x"))
    out = oracle.encode(frames.GENERATORS[kind](int(f[1:]), w, h))
    g = golden["synthetic"][key]
    assert (out["jpg"].size, sha(out["jpg"])) == (g["bytes"], g["sha256"])


def test_oracle_comparator_golden(oracle, frames, golden):
    g = golden["comparator"]["640_A_vs_diffs"]
    A, B = frames.sample_bgr("640"), frames.sample_bgr("640_diffs")
    subA, subB = oracle.subsample(A), oracle.subsample(B)
    assert sha(b"P6\n160 160\n255\n" + subA.tobytes()) == g["subA_ppm_sha256"]
    assert sha(b"P6\n160 160\n255\n" + subB.tobytes()) == g["subB_ppm_sha256"]
    n, outs = oracle.compare(subB, subA, 640, 640)
    assert n == g["n"] and [list(o) for o in outs[:n]] == g["regions"]
    for i in range(n):
        out = oracle.encode(B, outs[i])
        assert out["jpg"].tobytes() == open(os.path.join(GOLD, f"region{i}_640.jpg"), "rb").read()
        assert sha(out["jpg"]) == g["jpgs"][i]["sha256"]


def test_oracle_comparator_micro(oracle, golden):
    micro = golden["comparator"]["micro_128"]
    for nm, case in micro.items():
        if nm == "dark_on_bright":
            s = np.full((32, 32, 3), 255, np.uint8)
            s[4:16, 4:16] = 0
            n, outs = oracle.compare(s, np.full((32, 32, 3), 255, np.uint8), 128, 128)
        else:
            s = np.zeros((32, 32, 3), np.uint8)
            for r in case["rects"]:
                s[r[2]:r[3] + 1, r[0]:r[1] + 1] = 255
            n, outs = oracle.compare(s, np.zeros_like(s), 128, 128)
        assert n == case["n"] and [list(o) for o in outs[:n]] == case["regions"], nm


# ---------------------------------------------------------------- direct oracle-vs-reference (container only)

def _rand_img(rng, h, w, kind):
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == 1:  # grey: every pixel on the exact-integer colour boundary
        return np.repeat(rng.integers(0, 256, (h, w, 1), dtype=np.uint8), 3, axis=2)
    if kind == 2:  # flat blocks: DC/128 coincidences
        v = rng.integers(0, 256, (h // 8, w // 8, 3), dtype=np.uint8)
        return np.ascontiguousarray(np.kron(v, np.ones((8, 8, 1), np.uint8)))
    base = rng.integers(0, 256, (1, 1, 3)).astype(np.int32)  # smooth-ish
    g = base + (np.arange(w)[None, :, None] * rng.integers(-2, 3) + np.arange(h)[:, None, None] * rng.integers(-2, 3))
    return np.clip(g + rng.integers(-3, 4, (h, w, 3)), 0, 255).astype(np.uint8)


def test_oracle_vs_ref_random_images(oracle, ref):
    rng = np.random.default_rng(1234)
    for it in range(24):
        w, h = 16 * int(rng.integers(1, 9)), 16 * int(rng.integers(1, 7))
        img = _rand_img(rng, h, w, it % 4)
        a, b = oracle.encode(img), ref.encode(img)
        for p in ("Y", "Cb", "Cr"):
            assert np.array_equal(a[p], b[p]), (it, p)
        for nm in ("luma", "chroma"):
            for i in range(2):
                for k in HUFF_FIELDS:
                    assert np.array_equal(a[nm][i][k], b[nm][i][k]), (it, nm, i, k)
        assert a["jpg"].tobytes() == b["jpg"].tobytes(), it


def test_timing_harness_keeps_the_streams_it_times(oracle, ref):
    """bench.py compares the device's bytes with the streams its CPU arm produced while it was timed (parity_live)."""
    rng = np.random.default_rng(77)
    batch = np.stack([_rand_img(rng, 48, 64, k) for k in range(4)])
    for chk in (oracle, ref):
        sec, nbytes, streams = chk.time_encode(batch, 2, keep=True)
        assert sec > 0 and len(streams) == 4 and nbytes == 2 * sum(len(s) for s in streams)
        for k in range(4):
            assert streams[k] == oracle.encode(batch[k])["jpg"].tobytes(), (chk.name, k)


def test_oracle_vs_ref_crops(oracle, ref, frames):
    img = frames.sample_bgr("640_diffs")
    for area in [(2, 36, 112, 432), (358, 66, 256, 336), (0, 0, 16, 16), (624, 624, 16, 16), (3, 5, 48, 32), (101, 7, 528, 16)]:
        a, b = oracle.encode(img, area), ref.encode(img, area)
        assert a["jpg"].tobytes() == b["jpg"].tobytes(), area


def test_oracle_vs_ref_table_builder_fuzz(oracle, ref):
    rng = np.random.default_rng(7)
    for it in range(1500):
        # nsym <= 254: with 255/256 used symbols the reference runs off the end of sym_sorted
        # (encoder.c:277,289-299, undefined behaviour); baseline JPEG has at most 162 AC symbols.
        nsym = int(rng.integers(1, 255 if it % 3 else 20))
        freq = np.zeros(257, np.int64)
        idx = rng.choice(256, nsym, replace=False)
        style = it % 5
        if style == 0:
            freq[idx] = rng.integers(1, 4, nsym)            # many ties
        elif style == 1:
            freq[idx] = rng.integers(1, 100000, nsym)
        elif style == 2:
            freq[idx] = np.minimum(1.6 ** np.minimum(np.arange(nsym), 40), 2 ** 20).astype(np.int64)  # deep tree -> 16-bit limiter
        elif style == 3:
            freq[idx] = 1
        else:
            freq[idx] = rng.geometric(0.02, nsym)
        freq[256] = 1
        a, b = oracle.build_table(freq), ref.build_table(freq)
        for k in HUFF_FIELDS:
            assert np.array_equal(a[k], b[k]), (it, k)


def test_oracle_vs_ref_comparator_fuzz(oracle, ref):
    rng = np.random.default_rng(99)
    for it in range(300):
        W, H = 16 * int(rng.integers(2, 21)), 16 * int(rng.integers(2, 16))
        sw, sh = W // 4, H // 4
        saved = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        sub = saved.copy()
        style = it % 4
        if style == 0:      # a few rectangles
            for _ in range(int(rng.integers(1, 8))):
                x0, y0 = int(rng.integers(0, sw)), int(rng.integers(0, sh))
                x1, y1 = min(sw, x0 + int(rng.integers(1, 20))), min(sh, y0 + int(rng.integers(1, 20)))
                sub[y0:y1, x0:x1] = rng.integers(0, 256, 3, dtype=np.uint8)
        elif style == 1:    # salt: many tiny regions -> >99 overflow path
            m = rng.random((sh, sw)) < float(rng.uniform(0.01, 0.3))
            sub[m] = 255 - sub[m]
        elif style == 2:    # small perturbations around the threshold
            sub = np.clip(sub.astype(np.int32) + rng.integers(-14, 15, sub.shape), 0, 255).astype(np.uint8)
        else:               # snakes that merge labels
            for _ in range(int(rng.integers(1, 5))):
                x, y = int(rng.integers(0, sw)), int(rng.integers(0, sh))
                for _ in range(int(rng.integers(5, 200))):
                    sub[y, x] = 255 - sub[y, x] if abs(int(sub[y, x, 1]) - 128) > 30 else 255
                    x = int(np.clip(x + rng.integers(-1, 2), 0, sw - 1)); y = int(np.clip(y + rng.integers(0, 2), 0, sh - 1))
        na, oa = oracle.compare(sub, saved, W, H)
        nb, ob = ref.compare(sub, saved, W, H)
        assert na == nb and oa == ob, (it, W, H, na, nb)


def test_oracle_vs_ref_subsample_and_enlarge(oracle, ref):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    assert np.array_equal(oracle.subsample(img), ref.subsample(img))
    for _ in range(500):
        W, H = 16 * int(rng.integers(2, 40)), 16 * int(rng.integers(2, 40))
        x0, y0 = int(rng.integers(0, W // 4)), int(rng.integers(0, H // 4))
        x1, y1 = int(rng.integers(x0, W // 4)), int(rng.integers(y0, H // 4))
        assert oracle.enlarge_adjust((x0, y0, x1, y1), W, H) == ref.enlarge_adjust((x0, y0, x1, y1), W, H)


def test_fmt2rgb888_known_answers(oracle):
    """Input side: the RGB565 / GRAYSCALE branches of fmt2rgb888 (esp32-camera 2.0.3, conversions/to_bmp.c; the reference calls
    it at main/main.c:134).  The dependency is not in the reference tree: these vectors were worked out by hand from its
    published source -  b = (lb & 0x1F) << 3, g = (hb & 0x07) << 5 | (lb & 0xE0) >> 3, r = hb & 0xF8, written B, G, R."""
    kat = {  # (hb, lb) -> (B, G, R)
        (0x00, 0x00): (0, 0, 0), (0xFF, 0xFF): (0xF8, 0xFC, 0xF8), (0xF8, 0x00): (0, 0, 0xF8), (0x07, 0xE0): (0, 0xFC, 0),
        (0x00, 0x1F): (0xF8, 0, 0), (0x12, 0x34): (0xA0, 0x44, 0x10), (0xA5, 0x5A): (0xD0, 0xA8, 0xA0), (0x80, 0x01): (0x08, 0, 0x80)}
    src = np.array([v for k in kat for v in k], np.uint8)
    got = oracle.fmt2rgb888(src, 1, len(kat))
    assert [tuple(int(x) for x in row) for row in got] == list(kat.values())
    g = np.arange(256, dtype=np.uint8)
    assert np.array_equal(oracle.fmt2rgb888(g, 2, 256), np.repeat(g[:, None], 3, axis=1))


# ---- second oracle: the upstream encoder utils/original.c --------------------------------------
UPSTREAM = os.path.join(ROOT, "oracle", "_ref", "upstream_original")


def _run_upstream(tmp_path, rgb):
    """utils/original.c is a program: <ppm> <quality> -> ./out.jpg, and it needs ./hisParts/ for its stage dumps."""
    import subprocess
    (tmp_path / "hisParts").mkdir(exist_ok=True)
    h, w, _ = rgb.shape
    (tmp_path / "in.ppm").write_bytes(b"P6\n%d %d\n255\n" % (w, h) + rgb.tobytes())
    subprocess.run([UPSTREAM, "in.ppm", "50"], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=120)
    return (tmp_path / "out.jpg").read_bytes()


@pytest.mark.skipif(not os.path.exists(UPSTREAM), reason="oracle/_ref/upstream_original not built (needs /root/reference)")
@pytest.mark.parametrize("key,src", [("sample_64x64_bgr", "64"), ("sample_640x640_bgr", "640"), ("sample_640x640_diffs_bgr", "640_diffs")])
def test_upstream_original_agrees_on_samples(tmp_path, frames, golden, oracle, key, src):
    """The upstream program reads the PPM as R,G,B; the reference reads the same bytes as B,G,R (encoder.c:132-135): the
    reference's output for the byte-reversed frame (the *_bgr fixtures) must be the upstream's output for the file."""
    jpg = _run_upstream(tmp_path, frames.sample_rgb(src))
    assert sha(jpg) == golden["encode"][key]["sha256"]
    assert jpg == oracle.encode(frames.sample_bgr(src))["jpg"].tobytes()


@pytest.mark.skipif(not os.path.exists(UPSTREAM), reason="oracle/_ref/upstream_original not built (needs /root/reference)")
def test_upstream_original_agrees_on_random_images(tmp_path, oracle):
    rng = np.random.default_rng(11)
    for (h, w) in [(16, 16), (32, 48), (64, 16), (48, 80)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if h == 32:
            rgb[:] = rgb[:, :, :1]                      # grey: every pixel on an integer boundary of the colour chain
        assert _run_upstream(tmp_path, rgb) == oracle.encode(np.ascontiguousarray(rgb[:, :, ::-1]))["jpg"].tobytes(), (h, w)

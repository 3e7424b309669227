"""Input frames for parity tests and the benchmark (host side, numpy only).

The encoder consumes B,G,R interleaved bytes (reference main/encoder.c:133-135 puts
the 0.299 weight on byte +2), while the sample fixtures are P6 PPMs (R,G,B).
Everything returned by this module is already in encoder byte order (BGR) unless
the name says ``rgb``.

Fixtures (tests/golden/, written by tests/golden/make_golden.py from the
reference's images/ directory, which does not exist on the GPU box):
  sample_64x64.ppm                 raw P6
  sample_640x640.lf.xz             left-neighbour-filtered bytes, xz
  sample_640x640_diffs.delta.xz    (diffs - sample) mod 256, xz

Synthetic frame classes of SURVEY.md §8(d), config 4/5:
  natural  : tile of the two 640x640 samples, cyclically shifted per frame
  noise    : splitmix64 byte stream
  ramp     : R=G=B=(x+y+f)&255 (every pixel sits on an exact-integer colour boundary)
"""
from __future__ import annotations

import lzma
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

_cache: dict = {}


def read_ppm(path: str) -> np.ndarray:
    """Minimal P6 reader -> (H, W, 3) uint8 in file (R,G,B) order."""
    with open(path, "rb") as f:
        data = f.read()
    assert data[:2] == b"P6", path
    tokens, pos = [], 2
    while len(tokens) < 3:
        while data[pos : pos + 1].isspace():
            pos += 1
        if data[pos : pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end : end + 1].isspace():
            end += 1
        tokens.append(int(data[pos:end]))
        pos = end
    pos += 1  # the single whitespace byte after maxval
    w, h, maxval = tokens
    assert maxval == 255
    return np.frombuffer(data, np.uint8, count=w * h * 3, offset=pos).reshape(h, w, 3).copy()


def write_ppm(path: str, rgb: np.ndarray) -> None:
    h, w, _ = rgb.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgb).tobytes())


def swap_rb(img: np.ndarray) -> np.ndarray:
    """RGB <-> BGR."""
    return np.ascontiguousarray(img[..., ::-1])


def sample_rgb(name: str) -> np.ndarray:
    """'64', '640' or '640_diffs' -> (H, W, 3) uint8 in PPM (R,G,B) order."""
    if name in _cache:
        return _cache[name]
    if name == "64":
        img = read_ppm(os.path.join(GOLDEN_DIR, "sample_64x64.ppm"))
    elif name == "640":
        with open(os.path.join(GOLDEN_DIR, "sample_640x640.lf.xz"), "rb") as f:
            filt = np.frombuffer(lzma.decompress(f.read()), np.uint8).reshape(640, 640, 3)
        img = np.cumsum(filt.astype(np.uint32), axis=1).astype(np.uint8)  # undo left-neighbour filter (mod 256)
    elif name == "640_diffs":
        with open(os.path.join(GOLDEN_DIR, "sample_640x640_diffs.delta.xz"), "rb") as f:
            delta = np.frombuffer(lzma.decompress(f.read()), np.uint8).reshape(640, 640, 3)
        img = (sample_rgb("640").astype(np.uint16) + delta).astype(np.uint8)
    else:
        raise KeyError(name)
    _cache[name] = img
    return img


def sample_bgr(name: str) -> np.ndarray:
    return swap_rb(sample_rgb(name))


def tile_bgr(w: int, h: int) -> np.ndarray:
    """Stand-in for the missing sample_1920x1280.ppm: rows alternate
    [A B A ...] / [B A B ...] of the two 640x640 samples, cropped to (h, w)."""
    key = ("tile", w, h)
    if key not in _cache:
        A, B = sample_bgr("640"), sample_bgr("640_diffs")
        nx, ny = -(-w // 640), -(-h // 640)
        rows = [np.concatenate([(A, B)[(ix + iy) & 1] for ix in range(nx)], axis=1) for iy in range(ny)]
        _cache[key] = np.ascontiguousarray(np.concatenate(rows, axis=0)[:h, :w])
    return _cache[key]


def natural_shift(f: int, w: int, h: int) -> tuple[int, int]:
    """(dx, dy) of frame f: 16-pixel steps, x fastest, wrapping over the MCU grid."""
    return 16 * (f % (w // 16)), 16 * ((f // (w // 16)) % (h // 16))


def natural_frame(f: int, w: int = 1920, h: int = 1280) -> np.ndarray:
    dx, dy = natural_shift(f, w, h)
    return np.ascontiguousarray(np.roll(tile_bgr(w, h), (dy, dx), axis=(0, 1)))


_SEED = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def noise_frame(f: int, w: int = 1920, h: int = 1280) -> np.ndarray:
    """byte i of frame f = byte (i&7) (little endian) of splitmix64(seed ^ (f<<40) ^ (i>>3))."""
    n = w * h * 3
    words = np.arange((n + 7) // 8, dtype=np.uint64) ^ np.uint64((_SEED ^ (f << 40)) & _M64)
    return _splitmix64(words).view(np.uint8)[:n].reshape(h, w, 3).copy()


def ramp_frame(f: int, w: int = 1920, h: int = 1280) -> np.ndarray:
    v = ((np.arange(w)[None, :] + np.arange(h)[:, None] + f) & 255).astype(np.uint8)
    return np.ascontiguousarray(np.repeat(v[:, :, None], 3, axis=2))


def moving_sequence(n: int, w: int = 640, h: int = 640, seed: int = 5) -> np.ndarray:
    """(n, h, w, 3) frames of one scene for the comparator workload (BASELINE.json config 3, app_main's loop main.c:137-162):
    the seed image (640x640 sample / natural tile) with rectangles of its `diffs` counterpart (or flat patches) pasted at
    changing places, so that consecutive frames differ in a handful of regions; every 7th frame repeats the plain scene."""
    rng = np.random.default_rng(seed)
    if (w, h) == (640, 640):
        A, B = sample_bgr("640"), sample_bgr("640_diffs")
    else:
        A = tile_bgr(w, h)
        B = np.ascontiguousarray(np.roll(A, (h // 2 // 16 * 16, 640), axis=(0, 1)))
    out = np.empty((n, h, w, 3), np.uint8)
    for f in range(n):
        img = A.copy()
        for _ in range(int(rng.integers(1, 5)) if f % 7 else 0):
            rw, rh = int(rng.integers(8, w // 3)), int(rng.integers(8, h // 3))
            x, y = int(rng.integers(0, w - rw)), int(rng.integers(0, h - rh))
            if rng.integers(0, 3):
                img[y:y + rh, x:x + rw] = B[y:y + rh, x:x + rw]
            else:
                img[y:y + rh, x:x + rw] = rng.integers(0, 256, 3, dtype=np.uint8)
        out[f] = img
    return out


GENERATORS = {"natural": natural_frame, "noise": noise_frame, "ramp": ramp_frame}


def batch(kind: str, n: int, w: int = 1920, h: int = 1280, start: int = 0) -> np.ndarray:
    """(n, h, w, 3) uint8 BGR frames start..start+n-1 of one class."""
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        out[i] = GENERATORS[kind](start + i, w, h)
    return out

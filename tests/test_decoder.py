"""Decoding side.

CPU part: the oracle decoder (oracle/oracle_decode.c) recovers, value for value, the coefficient planes the reference's
rgb_to_dct produced for a stream - the golden streams of the unmodified reference included - and its pixels stay close to
the input and to an independent decoder (PIL / libjpeg).  The pixel arithmetic is restated from the reference's decoder
stubs (utils/func_tester.c:1261-1319): "parity unpinned" for that half, the oracle's own header says so.

GPU part: jpegb200_decode_batch* against the oracle decoder, bit for bit, planes and pixels, and encode -> decode round
trips of device-resident batches at full size.
"""
import io
import os

import numpy as np
import pytest

from conftest import ROOT, sha

GOLD = os.path.join(ROOT, "tests", "golden")


def _planes(e):
    return np.concatenate([e["Y"], e["Cb"], e["Cr"]])


# ---------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("src", ["64", "640", "640_diffs"])
def test_oracle_decoder_recovers_the_planes(oracle, frames, src):
    img = frames.sample_bgr(src)
    e = oracle.encode(img)
    d = oracle.decode(e["jpg"], img.shape[1], img.shape[0])
    assert d["rc"] == 0 and (d["w"], d["h"]) == (img.shape[1], img.shape[0])
    for k in ("Y", "Cb", "Cr"):
        assert np.array_equal(d[k], e[k]), k
    err = d["bgr"].astype(np.float64) - img
    assert 10 * np.log10(255.0 ** 2 / np.mean(err ** 2)) > 24.0          # this quantiser on these photographs: 25-31 dB


def test_oracle_decoder_on_the_reference_golden_streams(oracle, golden):
    """Streams written by the UNMODIFIED reference (tests/golden/*.jpg): planes_sha256 was taken from its rgb_to_dct output."""
    seen = 0
    for key, g in golden["encode"].items():
        path = os.path.join(GOLD, key + ".jpg")
        if not os.path.exists(path):
            continue
        jpg = open(path, "rb").read()
        side = 64 if "64x64" in key else 640
        d = oracle.decode(jpg, side, side)
        assert d["rc"] == 0, key
        assert sha(d["Y"].tobytes() + d["Cb"].tobytes() + d["Cr"].tobytes()) == g["planes_sha256"], key
        seen += 1
    assert seen >= 2


def test_oracle_decoder_close_to_libjpeg(oracle, frames):
    from PIL import Image
    img = frames.sample_bgr("640")
    e = oracle.encode(img)
    d = oracle.decode(e["jpg"], 640, 640)
    pil = np.array(Image.open(io.BytesIO(e["jpg"].tobytes())).convert("RGB"))[:, :, ::-1]
    diff = np.abs(pil.astype(int) - d["bgr"].astype(int))
    assert diff.mean() < 1.5 and diff.max() <= 24       # other IDCT rounding, smooth chroma up-sampling, other colour constants


def test_oracle_decoder_random_and_extreme_content(oracle, frames):
    rng = np.random.default_rng(5)
    cases = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in [(16, 16), (48, 80), (64, 16), (32, 272)]]
    cases += [np.zeros((16, 32, 3), np.uint8), np.full((32, 16, 3), 255, np.uint8), frames.noise_frame(3, 64, 64), frames.ramp_frame(3, 64, 64)]
    chk = np.indices((32, 32)).sum(0) % 2 * 255                                   # checkerboard: 0xFF-heavy scans, ZRL symbols
    cases.append(np.repeat(chk[:, :, None], 3, 2).astype(np.uint8))
    for img in cases:
        e = oracle.encode(img)
        d = oracle.decode(e["jpg"], img.shape[1], img.shape[0])
        assert d["rc"] == 0 and np.array_equal(_planes(d), _planes(e)), img.shape


def test_oracle_decoder_rejects_damage(oracle, frames):
    jpg = bytearray(oracle.encode(frames.sample_bgr("64"))["jpg"].tobytes())
    assert oracle.decode(bytes(jpg[:200]), 64, 64)["rc"] < 0                      # truncated inside the tables
    assert oracle.decode(b"\x00" * 64, 64, 64)["rc"] < 0                           # no SOI
    bad = bytearray(jpg)
    bad[bad.index(b"\xff\xc0") + 4] = 12                                           # 12-bit precision
    assert oracle.decode(bytes(bad), 64, 64)["rc"] < 0
    assert oracle.decode(bytes(jpg), 64, 80)["rc"] < 0                             # not the dimensions the buffers were sized for


STUBS = os.path.join(ROOT, "oracle", "_ref", "libstubs.so")


@pytest.mark.skipif(not os.path.exists(STUBS), reason="oracle/_ref/libstubs.so not built (needs /root/reference)")
def test_decoder_steps_against_the_reference_stubs(oracle):
    """The stubs that are complete functions pin their step: toRgb on every (Y, Cb, Cr) whose three results lie in [0, 255]
    (outside, the stub's double -> uint8_t store is undefined; the decoder clamps), fromZigZag, abs_dc."""
    import ctypes as C
    L = C.CDLL(STUBS)
    n = 64 * 64
    ycc = (C.c_uint8 * (3 * n)).in_dll(L, "YCbCr")
    rgb = (C.c_uint8 * (3 * n)).in_dll(L, "rgb")
    ycc_np = np.ctypeslib.as_array(ycc).reshape(3, n)
    rgb_np = np.ctypeslib.as_array(rgb).reshape(3, n)
    oracle.lib.orc_to_bgr.argtypes = [C.POINTER(C.c_uint8)] * 3 + [C.c_int, C.POINTER(C.c_uint8)]
    rng = np.random.default_rng(4)
    checked = 0
    for rep in range(64):
        if rep < 16:                                   # structured: all Y against a slice of (Cb, Cr)
            yy, cc = np.meshgrid(np.arange(256), np.arange(16), indexing="ij")
            Y = yy.ravel().astype(np.uint8)
            Cb = ((cc.ravel() * 16 + rep) % 256).astype(np.uint8)
            Cr = ((cc.ravel() * 16 + 5 * rep + 3) % 256).astype(np.uint8)
        else:
            Y = rng.integers(0, 256, n, dtype=np.uint8)
            Cb, Cr = (rng.integers(128 - 56, 128 + 56, n).astype(np.uint8) for _ in range(2))     # photographic chroma: mostly in range
        ycc_np[0], ycc_np[1], ycc_np[2] = Y, Cb, Cr
        L.toRgb()
        mine = np.zeros((n, 3), np.uint8)
        ya, ca, ra = (np.ascontiguousarray(a) for a in (Y, Cb, Cr))
        oracle.lib.orc_to_bgr(ya.ctypes.data_as(C.POINTER(C.c_uint8)), ca.ctypes.data_as(C.POINTER(C.c_uint8)), ra.ctypes.data_as(C.POINTER(C.c_uint8)), n,
                              mine.ctypes.data_as(C.POINTER(C.c_uint8)))
        y, cb, cr = Y.astype(np.float64), Cb.astype(np.float64) - 128, Cr.astype(np.float64) - 128
        r, g, b = y + 1.4 * cr, y - 0.343 * cb - 0.711 * cr, y + 1.765 * cb
        ok = (r >= 0) & (r < 256) & (g >= 0) & (g < 256) & (b >= 0) & (b < 256)
        assert np.array_equal(mine[ok, 2], rgb_np[0][ok]) and np.array_equal(mine[ok, 1], rgb_np[1][ok]) and np.array_equal(mine[ok, 0], rgb_np[2][ok])
        checked += int(ok.sum())
    assert checked > 120000
    # fromZigZag (:1311-1314): out[scan_order[i]] = in[i]
    L.fromZigZag.argtypes = [C.POINTER(C.c_int16)] * 2
    src = np.arange(100, 164, dtype=np.int16)
    dst = np.zeros(64, np.int16)
    L.fromZigZag(src.ctypes.data_as(C.POINTER(C.c_int16)), dst.ctypes.data_as(C.POINTER(C.c_int16)))
    zz = np.ctypeslib.as_array((C.c_int * 64).in_dll(oracle.lib, "orc_zigzag"))
    assert np.array_equal(dst[zz], src)
    # abs_dc (:1316-1319): DC = running sum of the differences (the stub works on int8; small values)
    L.abs_dc.argtypes = [C.POINTER(C.c_int8)]
    q = np.zeros(n, np.int8)
    diffs = rng.integers(-3, 4, n // 64).astype(np.int8)
    q[::64] = diffs
    L.abs_dc(q.ctypes.data_as(C.POINTER(C.c_int8)))
    assert np.array_equal(q[::64].astype(int), np.cumsum(diffs.astype(int)))


# ---------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def enc():
    import importlib
    pkg = importlib.import_module("jpeg-encoder-decoder_b200")
    e = pkg.Encoder(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_decoder_equals_oracle_decoder(enc, oracle, frames):
    rng = np.random.default_rng(9)
    groups = {}
    imgs = [frames.sample_bgr("64"), frames.noise_frame(1, 64, 64), frames.ramp_frame(2, 64, 64), rng.integers(0, 256, (64, 64, 3), dtype=np.uint8),
            frames.sample_bgr("640"), frames.sample_bgr("640_diffs"), rng.integers(0, 256, (48, 80, 3), dtype=np.uint8),
            np.repeat((np.indices((48, 80)).sum(0) % 2 * 255)[:, :, None], 3, 2).astype(np.uint8), rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)]
    for img in imgs:
        groups.setdefault(img.shape[:2], []).append(img)
    for (h, w), group in groups.items():
        jpgs = [oracle.encode(img)["jpg"].tobytes() for img in group]
        bgr, status, planes = enc.decode_streams(jpgs, w, h, planes=True)
        assert not status.any(), status
        for i, jpg in enumerate(jpgs):
            d = oracle.decode(jpg, w, h)
            assert np.array_equal(planes[i], _planes(d)), (h, w, i)
            assert np.array_equal(bgr[i], d["bgr"]), (h, w, i, int(np.abs(bgr[i].astype(int) - d["bgr"]).max()))


@pytest.mark.gpu
def test_gpu_decoder_paths_agree(enc, oracle, frames):
    """The sub-sequence decoder (parallel inside a scan), its fallback and the warp-per-scan decoder give the same planes and
    pixels, and every scan of photographic, noisy, grey and flat content settles: the last synchronisation pass changes
    nothing and the blocks add up (a flat plane is the periodic stream "DC 0, EOB, DC 0, EOB ...", which only settles
    because the speculation starts in phase with it: at a block start)."""
    rng = np.random.default_rng(21)
    imgs = [frames.sample_bgr("640"), frames.sample_bgr("640_diffs"), frames.noise_frame(5, 640, 640), rng.integers(0, 256, (640, 640, 3), dtype=np.uint8),
            frames.ramp_frame(6, 640, 640), np.full((640, 640, 3), 77, np.uint8), np.zeros((640, 640, 3), np.uint8)]
    jpgs = [oracle.encode(img)["jpg"].tobytes() for img in imgs]
    enc.set_decode_sequential(0)
    bgr_p, st_p, pl_p = enc.decode_streams(jpgs, 640, 640, planes=True)
    stats = enc.decode_stats()
    assert stats["scans"] == 3 * len(jpgs) and stats["fallback"] == 0 and stats["changed_per_pass"][-1] == 0, stats
    enc.set_decode_sequential(2)
    bgr_f, st_f, pl_f = enc.decode_streams(jpgs, 640, 640, planes=True)
    assert enc.decode_stats()["fallback"] == 3 * len(jpgs)
    enc.set_decode_sequential(1)
    bgr_s, st_s, pl_s = enc.decode_streams(jpgs, 640, 640, planes=True)
    enc.set_decode_sequential(0)
    assert not st_p.any() and not st_f.any() and not st_s.any()
    assert np.array_equal(pl_p, pl_s) and np.array_equal(bgr_p, bgr_s)
    assert np.array_equal(pl_f, pl_s) and np.array_equal(bgr_f, bgr_s)


@pytest.mark.gpu
def test_gpu_decoder_golden_streams(enc, golden):
    for key, g in golden["encode"].items():
        path = os.path.join(GOLD, key + ".jpg")
        if not os.path.exists(path):
            continue
        side = 64 if "64x64" in key else 640
        _, status, planes = enc.decode_streams([open(path, "rb").read()], side, side, planes=True)
        assert status[0] == 0 and sha(planes[0].tobytes()) == g["planes_sha256"], key


@pytest.mark.gpu
def test_gpu_decoder_reports_bad_streams_without_stopping_the_batch(enc, oracle, frames):
    good = oracle.encode(frames.sample_bgr("64"))["jpg"].tobytes()
    cut = good[:len(good) // 2]                       # the Y scan ends early: the markers of the other scans are missing
    bad = bytearray(good)
    bad[bad.index(b"\xff\xc0") + 4] = 12
    noise = bytes(np.random.default_rng(2).integers(0, 256, 600, dtype=np.uint8))
    bgr, status = enc.decode_streams([good, cut, bytes(bad), noise, good], 64, 64)
    assert status[0] == 0 and status[4] == 0 and (status[1:4] < 0).all(), status
    want = oracle.decode(good, 64, 64)["bgr"]
    assert np.array_equal(bgr[0], want) and np.array_equal(bgr[4], want)


@pytest.mark.gpu
def test_gpu_decoder_survives_garbage_scans(enc, oracle, frames):
    """Valid headers and markers, random bytes inside the scans: every decoder mode must come back (a status or garbage
    pixels, never a fault) and the good streams of the same batch must be untouched."""
    rng = np.random.default_rng(33)
    good = oracle.encode(frames.sample_bgr("640"))["jpg"].tobytes()
    want = oracle.decode(good, 640, 640)["bgr"]
    sos = [i for i in range(len(good) - 1) if good[i] == 0xFF and good[i + 1] == 0xDA]
    assert len(sos) == 3
    bad = []
    for k in range(6):
        b = bytearray(good)
        lo, hi = sos[k % 3] + 10, (sos[k % 3 + 1] if k % 3 < 2 else len(good) - 2)
        junk = rng.integers(0, 255, hi - lo, dtype=np.uint8)            # never 0xFF: the markers stay where they are
        if k >= 3:
            junk[:] = (0x00, 0x55, 0xAA)[k - 3]                          # periodic junk
        b[lo:hi] = junk.tobytes()
        bad.append(bytes(b))
    for mode in (0, 2, 1):
        enc.set_decode_sequential(mode)
        bgr, status = enc.decode_streams([good] + bad + [good], 640, 640)
        assert status[0] == 0 and status[-1] == 0, (mode, status)
        assert np.array_equal(bgr[0], want) and np.array_equal(bgr[-1], want), mode
    enc.set_decode_sequential(0)


@pytest.mark.gpu
def test_encode_decode_round_trip_on_the_device_full_size(enc, frames):
    """64 frames of 1920 x 1280 stay on the device: encode (token path) -> decode; the decoder's planes must be the planes
    the plane path of the encoder materialises for the same frames, and the pixels must be close to the input."""
    import torch
    W, H, n = 1920, 1280, 16
    dev = torch.device("cuda", 0)
    host = np.stack([frames.GENERATORS["natural" if i % 2 == 0 else "noise"](i, W, H) for i in range(n)])
    d_in = torch.from_numpy(host).to(dev)
    slot = 2 * 1024 * 1024
    d_out = torch.zeros((n, slot), dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    enc.encode_batch_ptr(d_in.data_ptr(), n, W, H, W * H * 3, d_out.data_ptr(), slot, d_sizes.data_ptr(), st)
    d_bgr = torch.zeros((n, H, W, 3), dtype=torch.uint8, device=dev)
    d_planes = torch.zeros((n, W * H * 3 // 2), dtype=torch.int16, device=dev)
    d_status = torch.full((n,), 7, dtype=torch.int32, device=dev)
    enc.decode_batch_ptr(d_out.data_ptr(), slot, d_sizes.data_ptr(), n, W, H, d_bgr.data_ptr(), W * H * 3, d_planes.data_ptr(), d_status.data_ptr(), st)
    torch.cuda.synchronize()
    assert (d_sizes > 0).all() and not d_status.any()
    # re-encoding the decoded planes' source frames on the plane path gives the same planes: check through the stage function
    import importlib
    api = importlib.import_module("jpeg-encoder-decoder_b200").RefAPI()
    for i in (0, 1, n - 1):
        api.set_dims(W, H)
        Y, Cb, Cr = api.rgb_to_dct(host[i], (0, 0, W, H))
        assert np.array_equal(d_planes[i].cpu().numpy(), np.concatenate([Y, Cb, Cr])), i
    err = d_bgr[::2].float() - d_in[::2].float()          # the natural frames
    psnr = 10 * torch.log10(255.0 ** 2 / (err ** 2).mean())
    assert psnr > 23.0, float(psnr)                       # 24.7 dB: this quantiser on the tiled photograph (the CPU test sees 25.3 dB on one tile)


@pytest.mark.gpu
def test_decode_tester_c_caller(tmp_path, oracle, golden):
    """tools/decode_tester (plain C against include/jpegb200.h) on the golden streams of the unmodified reference: the PPMs it
    writes hold the oracle decoder's pixels."""
    import subprocess
    exe = os.path.join(ROOT, "tools", "decode_tester")
    assert os.path.exists(exe), "tools/decode_tester is built by __graft_entry__.build()"
    files = [os.path.join(GOLD, k + ".jpg") for k in ("sample_640x640_bgr", "sample_640x640_diffs_bgr")]
    files = [f for f in files if os.path.exists(f)]
    assert files
    r = subprocess.run([exe, "640", "640", str(tmp_path / "out")] + files, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "%d of %d streams decoded" % (len(files), len(files)) in r.stdout, (r.stdout, r.stderr)
    for i, f in enumerate(files):
        ppm = (tmp_path / ("out%d.ppm" % i)).read_bytes()
        hdr = b"P6\n640 640\n255\n"
        assert ppm.startswith(hdr)
        want = oracle.decode(open(f, "rb").read(), 640, 640)["bgr"][:, :, ::-1]
        assert ppm[len(hdr):] == want.tobytes()

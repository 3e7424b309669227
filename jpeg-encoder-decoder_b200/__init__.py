"""jpeg-encoder-decoder_b200 — host-side mirror of the reference interface over libjpegb200.so.

The product is the shared library next to this file (CUDA kernels for sm_100a + a plain C ABI,
see include/jpegb200.h).  This module only *binds* it with ctypes:

* ``RefAPI``   — the reference's own seven entry points, with the reference's signatures
                 (include/encoder.h:10-12, include/brain.h:7-10), exactly what a C caller links.
* ``Encoder``  — the batched C ABI (``jpegb200_encode_batch*``, ``jpegb200_compare_encode`` ...).

There is no CPU implementation here and nothing under oracle/ is ever imported: if the library is
missing or no B200 is visible, construction raises.  (Import it with
``importlib.import_module("jpeg-encoder-decoder_b200")`` — the directory name is not an identifier.)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JPEGB200_LIB") or os.path.join(_HERE, "libjpegb200.so")   # the override is a development aid (kernel variants)
ROOT = os.path.dirname(_HERE)

C_ABI_SYMBOLS = [
    "jpegb200_create", "jpegb200_destroy", "jpegb200_last_error", "jpegb200_configure", "jpegb200_launch_count",
    "jpegb200_set_timing", "jpegb200_get_timing", "jpegb200_get_stage_timing", "jpegb200_set_exact_dct", "jpegb200_set_token_path", "jpegb200_set_token_budget", "jpegb200_pin_host", "jpegb200_unpin_host", "jpegb200_debug_fix_count",
    "jpegb200_encode_batch", "jpegb200_encode_batch_host", "jpegb200_encode_batch_host_multi", "jpegb200_encode_batch_host_fmt", "jpegb200_unpack", "jpegb200_encode_regions",
    "jpegb200_stage_dct", "jpegb200_stage_huffman", "jpegb200_stage_write", "jpegb200_debug_build_tables",
    "jpegb200_subsample", "jpegb200_compare", "jpegb200_enlarge_adjust", "jpegb200_compare_encode", "jpegb200_compare_encode_batch",
]
REFERENCE_SYMBOLS = ["rgb_to_dct", "init_huffman", "write_jpg", "subsample", "store", "compare", "enlargeAdjust"]


class Area(C.Structure):
    """area_t, include/structs.h (passed by value)."""

    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int)]


class HuffCode(C.Structure):
    """huff_code, include/structs.h."""

    _fields_ = [("sym_freq", C.c_int * 257), ("code_len", C.c_int * 257), ("next", C.c_int * 257),
                ("code_len_freq", C.c_int * 32), ("sym_sorted", C.c_int * 256), ("sym_code_len", C.c_int * 256),
                ("sym_code", C.c_int * 256)]


HUFF_FIELDS = [f[0] for f in HuffCode._fields_]


def build(verbose: bool = False) -> str:
    """Compile libjpegb200.so in-tree (nvcc, sm_100a).  Works without a GPU."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libjpegb200.so failed")
    return LIB_PATH


_lib = None


def load_library() -> C.CDLL:
    """dlopen the product library and declare every prototype.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing - run __graft_entry__.build() (nvcc) first; there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    u8p, i16p, u32p, ip, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.c_void_p
    L.jpegb200_create.argtypes = [C.POINTER(vp), C.c_int]
    L.jpegb200_destroy.argtypes = [vp]
    L.jpegb200_destroy.restype = None
    L.jpegb200_last_error.restype = C.c_char_p
    L.jpegb200_configure.argtypes = [vp, C.c_int, C.c_int]
    L.jpegb200_set_exact_dct.argtypes = [vp, C.c_int]
    L.jpegb200_set_token_path.argtypes = [vp, C.c_int]
    L.jpegb200_set_token_budget.argtypes = [vp, C.c_int]
    L.jpegb200_pin_host.argtypes = [vp, C.c_size_t]
    L.jpegb200_unpin_host.argtypes = [vp]
    L.jpegb200_debug_fix_count.argtypes = [vp, C.c_int, u32p]
    L.jpegb200_launch_count.argtypes = [vp]
    L.jpegb200_launch_count.restype = C.c_uint64
    L.jpegb200_encode_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t, vp, vp]
    L.jpegb200_encode_batch_host.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp]
    L.jpegb200_encode_batch_host_fmt.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp]
    L.jpegb200_unpack.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.jpegb200_encode_batch_host_multi.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp]
    L.jpegb200_encode_regions.argtypes = [vp, vp, C.c_int, C.c_int, ip, C.c_int, vp, C.c_size_t, vp, vp]
    L.jpegb200_stage_dct.argtypes = [vp, u8p] + [C.c_int] * 6 + [i16p] * 3
    L.jpegb200_stage_huffman.argtypes = [vp] + [i16p] * 3 + [C.c_int, C.c_int, vp, vp]
    L.jpegb200_stage_write.argtypes = [vp, u8p, C.c_size_t] + [i16p] * 3 + [C.c_int, C.c_int, vp, vp]
    L.jpegb200_stage_write.restype = C.c_size_t
    L.jpegb200_debug_build_tables.argtypes = [vp, ip, C.c_int, vp]
    L.jpegb200_subsample.argtypes = [vp, u8p, C.c_int, C.c_int, u8p]
    L.jpegb200_compare.argtypes = [vp, u8p, u8p, C.c_int, C.c_int, ip]
    L.jpegb200_enlarge_adjust.argtypes = [vp, ip, C.c_int, C.c_int]
    L.jpegb200_compare_encode.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, ip, u8p, C.c_size_t, u32p, u8p]
    L.jpegb200_compare_encode_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, ip, ip, u32p, C.POINTER(C.c_uint64), vp, C.c_size_t]
    # the reference's own entry points (include/encoder.h, include/brain.h)
    L.jpegb200_decode_batch.argtypes = [vp, vp, C.c_size_t, vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp, vp, vp]
    L.jpegb200_decode_batch_host.argtypes = [vp, vp, C.c_size_t, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.jpegb200_set_decode_sequential.argtypes = [vp, C.c_int]
    L.jpegb200_debug_decode_stats.argtypes = [vp, u32p]
    L.jpegb200_set_dims.argtypes = [C.c_int, C.c_int]
    L.jpegb200_set_dims.restype = None
    L.rgb_to_dct.argtypes = [u8p, i16p, i16p, i16p, Area]
    L.rgb_to_dct.restype = None
    L.init_huffman.argtypes = [i16p, i16p, i16p, Area, C.POINTER(HuffCode), C.POINTER(HuffCode)]
    L.init_huffman.restype = None
    L.write_jpg.argtypes = [vp, u8p, i16p, i16p, i16p, Area, C.POINTER(HuffCode), C.POINTER(HuffCode)]
    L.write_jpg.restype = C.c_size_t
    L.subsample.argtypes = [vp, u8p, u8p]
    L.subsample.restype = None
    L.store.argtypes = [u8p, u8p]
    L.store.restype = None
    L.compare.argtypes = [u8p, u8p, C.POINTER(Area), vp]
    L.compare.restype = C.c_uint8
    L.enlargeAdjust.argtypes = [C.POINTER(Area)]
    L.enlargeAdjust.restype = None
    _lib = L
    return L


def _u8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i16(a: np.ndarray):
    assert a.dtype == np.int16 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int16))


def _huff_dict(h: HuffCode) -> dict:
    return {k: np.ctypeslib.as_array(getattr(h, k)).copy() for k in HUFF_FIELDS}


class JpegB200Error(RuntimeError):
    pass


def pin_host(a: np.ndarray):
    """Page-lock a numpy array in place (jpegb200_pin_host); unpin_host(a) before it is dropped."""
    if load_library().jpegb200_pin_host(C.c_void_p(a.ctypes.data), a.nbytes):
        raise JpegB200Error(load_library().jpegb200_last_error().decode())


def unpin_host(a: np.ndarray):
    if load_library().jpegb200_unpin_host(C.c_void_p(a.ctypes.data)):
        raise JpegB200Error(load_library().jpegb200_last_error().decode())


class RefAPI:
    """The seven reference entry points, called exactly as a C program linked against
    libjpegb200.so would call them (main/encoder.c, main/brain.c of this repo).  Frame geometry is the
    library's run-time stand-in for the reference's compile-time WIDTH/HEIGHT."""

    def __init__(self):
        self.lib = load_library()
        self._libc = C.CDLL(None)
        self._libc.fopen.restype = C.c_void_p
        self._libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        self._libc.fclose.argtypes = [C.c_void_p]

    def set_dims(self, w: int, h: int):
        self.lib.jpegb200_set_dims(w, h)

    def rgb_to_dct(self, bgr: np.ndarray, area):
        H, W, _ = bgr.shape
        self.set_dims(W, H)
        a = Area(*area)
        n = a.w * a.h
        Y, Cb, Cr = np.zeros(n, np.int16), np.zeros(n // 4, np.int16), np.zeros(n // 4, np.int16)
        self.lib.rgb_to_dct(_u8(np.ascontiguousarray(bgr)), _i16(Y), _i16(Cb), _i16(Cr), a)
        return Y, Cb, Cr

    def init_huffman(self, Y, Cb, Cr, area):
        luma, chroma = (HuffCode * 2)(), (HuffCode * 2)()
        self.lib.init_huffman(_i16(Y), _i16(Cb), _i16(Cr), Area(*area), luma, chroma)
        return luma, chroma

    def write_jpg(self, Y, Cb, Cr, area, luma, chroma, path: str | None = None, cap: int | None = None):
        a = Area(*area)
        jpg = np.zeros(cap or (3 * a.w * a.h + 65536), np.uint8)
        f = self._libc.fopen((path or "/dev/null").encode(), b"wb")
        try:
            n = self.lib.write_jpg(f, _u8(jpg), _i16(Y), _i16(Cb), _i16(Cr), a, luma, chroma)
        finally:
            self._libc.fclose(f)
        return jpg[:n].copy()

    def encode(self, bgr: np.ndarray, area=None, path: str | None = None):
        """rgb_to_dct -> init_huffman -> write_jpg, the per-region body of app_main (main.c:144-152)."""
        H, W, _ = bgr.shape
        area = tuple(area or (0, 0, W, H))
        Y, Cb, Cr = self.rgb_to_dct(bgr, area)
        luma, chroma = self.init_huffman(Y, Cb, Cr, area)
        # the reference's callers pass a 3*PIX_LEN byte buffer; tiny noisy crops can exceed 3*w*h, so size by the frame
        jpg = self.write_jpg(Y, Cb, Cr, area, luma, chroma, path, cap=3 * W * H + 65536)
        return dict(jpg=jpg, Y=Y, Cb=Cb, Cr=Cr, luma=[_huff_dict(luma[0]), _huff_dict(luma[1])],
                    chroma=[_huff_dict(chroma[0]), _huff_dict(chroma[1])])

    def subsample(self, bgr: np.ndarray, path: str | None = None):
        H, W, _ = bgr.shape
        self.set_dims(W, H)
        out = np.zeros((H // 4, W // 4, 3), np.uint8)
        f = self._libc.fopen((path or "/dev/null").encode(), b"wb")
        try:
            self.lib.subsample(f, _u8(np.ascontiguousarray(bgr)), _u8(out))
        finally:
            self._libc.fclose(f)
        return out

    def store(self, sub: np.ndarray, W: int, H: int):
        self.set_dims(W, H)
        saved = np.zeros_like(sub)
        self.lib.store(_u8(np.ascontiguousarray(sub)), _u8(saved))
        return saved

    def compare(self, sub: np.ndarray, saved: np.ndarray, W: int, H: int):
        self.set_dims(W, H)
        outs = (Area * 100)()
        diffs = (C.c_int * (8 * (W // 8 + 1)))()
        n = self.lib.compare(_u8(np.ascontiguousarray(sub)), _u8(np.ascontiguousarray(saved)), outs, diffs)
        return int(n), [(o.x, o.y, o.w, o.h) for o in outs]

    def enlarge_adjust(self, box, W: int, H: int):
        self.set_dims(W, H)
        a = Area(*box)
        self.lib.enlargeAdjust(C.byref(a))
        return (a.x, a.y, a.w, a.h)


class Encoder:
    """One GPU context of the batched C ABI."""

    def __init__(self, device: int = 0, frames_per_wave: int | None = None, lanes: int | None = None):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        if self.lib.jpegb200_create(C.byref(self.ctx), device) != 0:
            raise JpegB200Error(self.lib.jpegb200_last_error().decode())
        if frames_per_wave is not None or lanes is not None:
            self.configure(frames_per_wave or 32, lanes or 3)

    def _check(self, rc):
        if rc < 0:
            raise JpegB200Error(self.lib.jpegb200_last_error().decode())
        return rc

    def configure(self, frames_per_wave: int, lanes: int):
        self._check(self.lib.jpegb200_configure(self.ctx, frames_per_wave, lanes))

    def set_exact_dct(self, on: bool):
        self._check(self.lib.jpegb200_set_exact_dct(self.ctx, int(on)))

    def set_token_path(self, on: bool):
        self._check(self.lib.jpegb200_set_token_path(self.ctx, int(on)))

    def set_token_budget(self, tokens_per_block: int):
        """Size the token pools of the batched path for `tokens_per_block` (0 / 65 = worst case; include/jpegb200.h)."""
        self._check(self.lib.jpegb200_set_token_budget(self.ctx, int(tokens_per_block)))

    def fix_count(self, lane: int = 0) -> int:
        n = C.c_uint32(0)
        self._check(self.lib.jpegb200_debug_fix_count(self.ctx, lane, C.byref(n)))
        return int(n.value)

    def close(self):
        if self.ctx:
            self.lib.jpegb200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.jpegb200_launch_count(self.ctx))

    # -- device resident: raw pointers (torch tensors' data_ptr()) ---------------------------------
    def encode_batch_ptr(self, d_bgr: int, n: int, w: int, h: int, frame_stride: int, d_out: int, slot: int, d_sizes: int, stream: int = 0):
        self._check(self.lib.jpegb200_encode_batch(self.ctx, d_bgr, n, w, h, frame_stride, d_out, slot, d_sizes, stream))

    def encode_regions_ptr(self, d_frame: int, fw: int, fh: int, areas, d_out: int, slot: int, d_sizes: int, stream: int = 0):
        flat = (C.c_int * (4 * len(areas)))(*[v for a in areas for v in a])
        self._check(self.lib.jpegb200_encode_regions(self.ctx, d_frame, fw, fh, flat, len(areas), d_out, slot, d_sizes, stream))

    # -- host buffers -----------------------------------------------------------------------------------
    def encode_batch_host(self, frames: np.ndarray, slot: int, out: np.ndarray | None = None, sizes: np.ndarray | None = None):
        """frames: (N, H, W, 3) uint8 BGR host array.  Returns (out (N, slot) uint8, sizes (N,) uint32)."""
        N, H, W, _ = frames.shape
        assert frames.dtype == np.uint8 and frames.flags.c_contiguous
        if out is None:
            out = np.empty((N, slot), np.uint8)
        if sizes is None:
            sizes = np.zeros(N, np.uint32)
        self._check(self.lib.jpegb200_encode_batch_host(self.ctx, frames.ctypes.data, N, W, H, out.ctypes.data, slot, sizes.ctypes.data))
        return out, sizes

    def encode_batch_host_ptr(self, h_bgr: int, n: int, w: int, h: int, h_out: int, slot: int, h_sizes: int):
        self._check(self.lib.jpegb200_encode_batch_host(self.ctx, h_bgr, n, w, h, h_out, slot, h_sizes))

    def encode_frames_fmt(self, packed: np.ndarray, fmt: int, w: int, h: int, slot: int | None = None) -> list[bytes]:
        """(N, h*w*bpp) packed frames (fmt 1 = RGB565, 2 = GRAYSCALE; include/jpegb200.h) -> list of JFIF byte strings."""
        packed = np.ascontiguousarray(packed, np.uint8)
        N = packed.shape[0]
        slot = slot or (w * h * 3 // 2 + 65536)
        out = np.empty((N, slot), np.uint8)
        sizes = np.zeros(N, np.uint32)
        self._check(self.lib.jpegb200_encode_batch_host_fmt(self.ctx, C.c_void_p(packed.ctypes.data), fmt, N, w, h, C.c_void_p(out.ctypes.data), slot,
                                                            C.c_void_p(sizes.ctypes.data)))
        return [out[i, : sizes[i]].tobytes() for i in range(N)]

    def unpack_ptr(self, d_src: int, fmt: int, n: int, w: int, h: int, d_bgr: int, stream: int = 0):
        self._check(self.lib.jpegb200_unpack(self.ctx, C.c_void_p(d_src), fmt, n, w, h, C.c_void_p(d_bgr), C.c_void_p(stream)))

    def encode_frames(self, frames: np.ndarray, slot: int | None = None) -> list[bytes]:
        """Convenience: list of JFIF byte strings for a (N, H, W, 3) BGR array."""
        N, H, W, _ = frames.shape
        slot = slot or (W * H * 3 // 2 + 65536)
        out, sizes = self.encode_batch_host(np.ascontiguousarray(frames), slot)
        if (sizes == 0).any():
            raise JpegB200Error("an output did not fit its slot")
        return [out[i, : sizes[i]].tobytes() for i in range(N)]

    # ---- decoding side (include/jpegb200.h: jpegb200_decode_batch*) ----
    def decode_batch_ptr(self, d_streams: int, slot: int, d_sizes: int, n: int, w: int, h: int, d_bgr: int, frame_stride: int, d_planes: int = 0,
                         d_status: int = 0, stream: int = 0):
        self._check(self.lib.jpegb200_decode_batch(self.ctx, d_streams, slot, d_sizes, n, w, h, d_bgr or None, frame_stride, d_planes or None, d_status or None,
                                                   stream or None))

    def set_decode_sequential(self, on):
        """False / 0: sub-sequence decoder with fallback; True / 1: warp-per-scan decoder; 2: sub-sequence decoder, every scan through the fallback."""
        self._check(self.lib.jpegb200_set_decode_sequential(self.ctx, int(on)))

    def decode_stats(self) -> dict:
        st = (C.c_uint32 * 8)()
        self._check(self.lib.jpegb200_debug_decode_stats(self.ctx, st))
        return dict(scans=st[0], fallback=st[1], changed_per_pass=list(st[2:8]))

    def decode_streams(self, jpgs: list, w: int, h: int, planes: bool = False):
        """jpgs: byte strings of w x h streams.  Returns (bgr (N,h,w,3) uint8, status (N,) int32[, planes (N, w*h*3/2) int16])."""
        n = len(jpgs)
        slot = (max(len(j) for j in jpgs) + 15) & ~15
        buf = np.zeros((n, slot), np.uint8)
        sizes = np.zeros(n, np.uint32)
        for i, j in enumerate(jpgs):
            buf[i, :len(j)] = np.frombuffer(bytes(j), np.uint8)
            sizes[i] = len(j)
        bgr = np.zeros((n, h, w, 3), np.uint8)
        status = np.zeros(n, np.int32)
        pl = np.zeros((n, w * h * 3 // 2), np.int16) if planes else None
        self._check(self.lib.jpegb200_decode_batch_host(self.ctx, buf.ctypes.data, slot, sizes.ctypes.data, n, w, h, bgr.ctypes.data,
                                                        pl.ctypes.data if planes else None, status.ctypes.data))
        return (bgr, status, pl) if planes else (bgr, status)

    def build_tables(self, freqs: np.ndarray) -> list[dict]:
        """Test hook: device table builder on (T, 257) int32 histograms."""
        freqs = np.ascontiguousarray(freqs, dtype=np.int32)
        T = freqs.shape[0]
        out = (HuffCode * T)()
        self._check(self.lib.jpegb200_debug_build_tables(self.ctx, freqs.ctypes.data_as(C.POINTER(C.c_int)), T, out))
        return [_huff_dict(out[i]) for i in range(T)]

    def compare_encode_batch(self, frames, max_regions: int = 16, arena_bytes: int | None = None, on_device: bool = False):
        """app_main's loop (main.c:137-162) for a batch of consecutive frames, compare -> encode hand-off on the device.
        frames: (F, H, W, 3) uint8 numpy array (host) or, with on_device, a tuple (device pointer, F, H, W).
        Returns (counts[F], boxes[F][100], jpgs[F][<=max_regions] (None where a region was not encoded))."""
        if on_device:
            ptr, F, H, W = frames
        else:
            frames = np.ascontiguousarray(frames)
            F, H, W, _ = frames.shape
            ptr = frames.ctypes.data
        arena_bytes = arena_bytes or (F * W * H + 4096 * F * max_regions + W * H + 4096)
        counts = np.zeros(F, np.int32)
        boxes = np.zeros((F, 100, 4), np.int32)
        sizes = np.zeros(F * max_regions, np.uint32)
        offsets = np.zeros(F * max_regions, np.uint64)
        arena = np.empty(arena_bytes, np.uint8)
        n = self._check(self.lib.jpegb200_compare_encode_batch(
            self.ctx, C.c_void_p(ptr), int(on_device), F, W, H, W * H * 3, max_regions, counts.ctypes.data_as(C.POINTER(C.c_int)),
            boxes.ctypes.data_as(C.POINTER(C.c_int)), sizes.ctypes.data_as(C.POINTER(C.c_uint32)), offsets.ctypes.data_as(C.POINTER(C.c_uint64)),
            C.c_void_p(arena.ctypes.data), arena_bytes))
        jpgs = []
        for f in range(F):
            row = []
            for i in range(min(int(counts[f]), max_regions)):
                k = f * max_regions + i
                row.append(arena[int(offsets[k]): int(offsets[k]) + int(sizes[k])].tobytes() if sizes[k] else None)
            jpgs.append(row)
        self.last_encoded = n
        return counts, boxes, jpgs

    def compare_encode(self, frame: np.ndarray, seed: bool = False, slot: int | None = None):
        """One iteration of app_main's loop (main.c:137-162).  Returns (regions, jpgs, sub)."""
        H, W, _ = frame.shape
        slot = slot or (W * H * 3 // 2 + 65536)
        outs = (C.c_int * 400)()
        sizes = np.zeros(100, np.uint32)
        sub = np.zeros((H // 4, W // 4, 3), np.uint8)
        out = np.empty((1 if seed else 100, slot), np.uint8)
        n = self._check(self.lib.jpegb200_compare_encode(self.ctx, _u8(np.ascontiguousarray(frame)), W, H, int(seed), outs, _u8(out), slot,
                                                         sizes.ctypes.data_as(C.POINTER(C.c_uint32)), _u8(sub)))
        regions = [tuple(outs[4 * i: 4 * i + 4]) for i in range(min(n, 100))]
        jpgs = [out[i, : sizes[i]].tobytes() for i in range(min(n, 100))]
        return regions, jpgs, sub

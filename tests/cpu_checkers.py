"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE, never the product path).

* ``Oracle``  -> oracle/_build/liboracle.so : our C restatement (oracle/oracle.c)
* ``Ref``     -> oracle/_ref/libref.so      : the UNMODIFIED reference compiled by oracle/Makefile
                                              (present only where it was built; it travels to the
                                              GPU box as a prebuilt file)

Both expose the same Python surface so that tests can swap one for the other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref.so")


class Area(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int)]

    def tup(self):
        return (self.x, self.y, self.w, self.h)


class HuffCode(C.Structure):
    """Layout of the reference's huff_code (include/structs.h:5-13)."""

    _fields_ = [
        ("sym_freq", C.c_int * 257),
        ("code_len", C.c_int * 257),
        ("next", C.c_int * 257),
        ("code_len_freq", C.c_int * 32),
        ("sym_sorted", C.c_int * 256),
        ("sym_code_len", C.c_int * 256),
        ("sym_code", C.c_int * 256),
    ]


HUFF_FIELDS = [f[0] for f in HuffCode._fields_]


def huff_to_dict(h: HuffCode) -> dict:
    return {k: np.ctypeslib.as_array(getattr(h, k)).copy() for k in HUFF_FIELDS}


def build_checkers() -> None:
    """(Re)build the checkers; building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def _u8(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i16(a):
    assert a.dtype == np.int16 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int16))


class _Base:
    name = "base"

    def encode(self, bgr: np.ndarray, area=None):
        """bgr: (H, W, 3) uint8 in B,G,R order.  Returns dict(jpg, Y, Cb, Cr, luma, chroma)."""
        raise NotImplementedError


class Oracle(_Base):
    name = "oracle"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_checkers()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_rgb_to_dct.argtypes = [C.POINTER(C.c_uint8), C.c_int, Area] + [C.POINTER(C.c_int16)] * 3
        L.orc_init_huffman.argtypes = [C.POINTER(C.c_int16)] * 3 + [Area, C.POINTER(HuffCode), C.POINTER(HuffCode)]
        L.orc_write_jpg.argtypes = [C.POINTER(C.c_uint8)] + [C.POINTER(C.c_int16)] * 3 + [Area, C.POINTER(HuffCode), C.POINTER(HuffCode)]
        L.orc_write_jpg.restype = C.c_size_t
        L.orc_encode.argtypes = [C.POINTER(C.c_uint8), C.c_int, Area, C.POINTER(C.c_uint8)]
        L.orc_encode.restype = C.c_size_t
        L.orc_build_table.argtypes = [C.POINTER(HuffCode)]
        L.orc_subsample.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_uint8)]
        L.orc_compare.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(Area)]
        L.orc_compare.restype = C.c_int
        L.orc_enlarge_adjust.argtypes = [C.POINTER(Area), C.c_int, C.c_int]
        L.orc_diff_mask.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8)]
        L.orc_time_encode.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_time_encode.restype = C.c_double
        L.orc_time_encode_keep.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_uint8), C.c_size_t,
                                           C.POINTER(C.c_uint32)]
        L.orc_time_encode_keep.restype = C.c_double
        L.orc_fmt2rgb888.argtypes = [C.POINTER(C.c_uint8), C.c_size_t, C.c_int, C.POINTER(C.c_uint8)]
        L.orc_time_loop.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.orc_time_loop.restype = C.c_double
        L.orc_decode.argtypes = [C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.POINTER(C.c_int16)] * 3 + [C.POINTER(C.c_uint8)]
        L.orc_time_decode.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
        L.orc_time_decode.restype = C.c_double

    def decode(self, jpg, w, h):
        """jpg: bytes / uint8 array of one stream of w x h pixels.  Returns dict(rc, w, h, Y, Cb, Cr, bgr)."""
        jpg = np.ascontiguousarray(np.frombuffer(bytes(jpg), np.uint8))
        n = w * h
        Y, Cb, Cr = np.zeros(n, np.int16), np.zeros(n // 4, np.int16), np.zeros(n // 4, np.int16)
        bgr = np.zeros((h, w, 3), np.uint8)
        ww, hh = C.c_int(w), C.c_int(h)
        rc = self.lib.orc_decode(_u8(jpg), jpg.size, C.byref(ww), C.byref(hh), _i16(Y), _i16(Cb), _i16(Cr), _u8(bgr))
        return dict(rc=rc, w=ww.value, h=hh.value, Y=Y, Cb=Cb, Cr=Cr, bgr=bgr)

    def time_decode(self, jpgs, sizes, w, h, reps=1):
        """jpgs: (N, slot) uint8, sizes: (N,) uint32.  Returns seconds."""
        scratch = np.zeros(w * h * 3, np.uint8)
        sizes = np.ascontiguousarray(sizes, np.uint32)
        return self.lib.orc_time_decode(_u8(jpgs), sizes.ctypes.data_as(C.POINTER(C.c_uint32)), jpgs.shape[1], jpgs.shape[0], reps, _u8(scratch))

    def encode(self, bgr, area=None):
        H, W, _ = bgr.shape
        a = Area(*(area or (0, 0, W, H)))
        n = a.w * a.h
        Y, Cb, Cr = np.zeros(n, np.int16), np.zeros(n // 4, np.int16), np.zeros(n // 4, np.int16)
        luma, chroma = (HuffCode * 2)(), (HuffCode * 2)()
        jpg = np.zeros(3 * n + 4096, np.uint8)
        bgr = np.ascontiguousarray(bgr)
        self.lib.orc_rgb_to_dct(_u8(bgr), W, a, _i16(Y), _i16(Cb), _i16(Cr))
        self.lib.orc_init_huffman(_i16(Y), _i16(Cb), _i16(Cr), a, luma, chroma)
        sz = self.lib.orc_write_jpg(_u8(jpg), _i16(Y), _i16(Cb), _i16(Cr), a, luma, chroma)
        return dict(jpg=jpg[:sz].copy(), Y=Y, Cb=Cb, Cr=Cr, luma=[huff_to_dict(luma[0]), huff_to_dict(luma[1])],
                    chroma=[huff_to_dict(chroma[0]), huff_to_dict(chroma[1])])

    def build_table(self, freq257):
        h = HuffCode()
        h.sym_freq[:] = list(map(int, freq257))
        self.lib.orc_build_table(C.byref(h))
        return huff_to_dict(h)

    def subsample(self, bgr):
        H, W, _ = bgr.shape
        out = np.zeros((H // 4, W // 4, 3), np.uint8)
        self.lib.orc_subsample(_u8(np.ascontiguousarray(bgr)), W, H, _u8(out))
        return out

    def compare(self, sub, saved, W, H):
        outs = (Area * 100)()
        n = self.lib.orc_compare(_u8(np.ascontiguousarray(sub)), _u8(np.ascontiguousarray(saved)), W, H, outs)
        return n, [outs[i].tup() for i in range(100)]

    def enlarge_adjust(self, box, W, H):
        a = Area(*box)
        self.lib.orc_enlarge_adjust(C.byref(a), W, H)
        return a.tup()

    def diff_mask(self, sub, saved):
        n = sub.size // 3
        m = np.zeros(n, np.uint8)
        self.lib.orc_diff_mask(_u8(np.ascontiguousarray(sub)), _u8(np.ascontiguousarray(saved)), n, _u8(m))
        return m

    def time_encode(self, frames, reps=1, keep=False):
        """frames: (N, H, W, 3) uint8.  Returns (seconds, jpeg_bytes) or, with keep, (seconds, jpeg_bytes, [stream of every frame])."""
        N, H, W, _ = frames.shape
        nb = C.c_size_t(0)
        if not keep:
            s = self.lib.orc_time_encode(_u8(frames), N, H * W * 3, W, H, reps, C.byref(nb))
            return s, nb.value
        slot = 3 * W * H
        out, sizes = np.zeros((N, slot), np.uint8), np.zeros(N, np.uint32)
        s = self.lib.orc_time_encode_keep(_u8(frames), N, H * W * 3, W, H, reps, C.byref(nb), _u8(out), slot, sizes.ctypes.data_as(C.POINTER(C.c_uint32)))
        return s, nb.value, [out[i, : sizes[i]].tobytes() for i in range(N)]


    def fmt2rgb888(self, packed: np.ndarray, fmt: int, npix: int) -> np.ndarray:
        """fmt 1 = RGB565 (hb, lb pairs), 2 = GRAYSCALE -> (npix, 3) B,G,R bytes."""
        packed = np.ascontiguousarray(packed, np.uint8).reshape(-1)
        out = np.zeros((npix, 3), np.uint8)
        assert self.lib.orc_fmt2rgb888(_u8(packed), packed.size, fmt, _u8(out)) == 1
        return out

    def time_loop(self, frames):
        """app_main's loop over frames[1:] with frames[0] as the seed.  Returns (seconds, regions encoded, jpeg bytes)."""
        N, H, W, _ = frames.shape
        nr, nb = C.c_int(0), C.c_size_t(0)
        s = self.lib.orc_time_loop(_u8(frames), N, H * W * 3, W, H, C.byref(nr), C.byref(nb))
        return s, nr.value, nb.value


class Ref(_Base):
    name = "reference"

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        assert L.ref_sizeof_huff_code() == C.sizeof(HuffCode)
        L.ref_encode.argtypes = [C.POINTER(C.c_uint8)] + [C.c_int] * 4 + [C.POINTER(C.c_int16)] * 3 + [C.POINTER(HuffCode)] * 2 + [C.POINTER(C.c_uint8)]
        L.ref_encode.restype = C.c_size_t
        L.ref_build_table.argtypes = [C.POINTER(HuffCode)]
        L.ref_subsample.argtypes = [C.POINTER(C.c_uint8)] * 2
        L.ref_compare.argtypes = [C.POINTER(C.c_uint8)] * 2 + [C.POINTER(Area)]
        L.ref_compare.restype = C.c_int
        L.ref_enlarge_adjust.argtypes = [C.POINTER(Area)]
        L.ref_time_encode.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
        L.ref_time_encode.restype = C.c_double
        L.ref_time_encode_keep.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_uint32)]
        L.ref_time_encode_keep.restype = C.c_double
        L.ref_time_loop.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.ref_time_loop.restype = C.c_double

    def encode(self, bgr, area=None):
        H, W, _ = bgr.shape
        self.lib.ref_set_dims(W, H)
        x, y, w, h = area or (0, 0, W, H)
        n = w * h
        Y, Cb, Cr = np.zeros(n, np.int16), np.zeros(n // 4, np.int16), np.zeros(n // 4, np.int16)
        luma, chroma = (HuffCode * 2)(), (HuffCode * 2)()
        jpg = np.zeros(3 * n + 4096, np.uint8)
        sz = self.lib.ref_encode(_u8(np.ascontiguousarray(bgr)), x, y, w, h, _i16(Y), _i16(Cb), _i16(Cr), luma, chroma, _u8(jpg))
        return dict(jpg=jpg[:sz].copy(), Y=Y, Cb=Cb, Cr=Cr, luma=[huff_to_dict(luma[0]), huff_to_dict(luma[1])],
                    chroma=[huff_to_dict(chroma[0]), huff_to_dict(chroma[1])])

    def build_table(self, freq257):
        h = HuffCode()
        h.sym_freq[:] = list(map(int, freq257))
        self.lib.ref_build_table(C.byref(h))
        return huff_to_dict(h)

    def subsample(self, bgr):
        H, W, _ = bgr.shape
        self.lib.ref_set_dims(W, H)
        out = np.zeros((H // 4, W // 4, 3), np.uint8)
        self.lib.ref_subsample(_u8(np.ascontiguousarray(bgr)), _u8(out))
        return out

    def compare(self, sub, saved, W, H):
        self.lib.ref_set_dims(W, H)
        outs = (Area * 100)()
        n = self.lib.ref_compare(_u8(np.ascontiguousarray(sub)), _u8(np.ascontiguousarray(saved)), outs)
        return n, [outs[i].tup() for i in range(100)]

    def enlarge_adjust(self, box, W, H):
        self.lib.ref_set_dims(W, H)
        a = Area(*box)
        self.lib.ref_enlarge_adjust(C.byref(a))
        return a.tup()

    def time_encode(self, frames, reps=1, keep=False):
        N, H, W, _ = frames.shape
        self.lib.ref_set_dims(W, H)
        nb = C.c_size_t(0)
        if not keep:
            s = self.lib.ref_time_encode(_u8(frames), N, H * W * 3, reps, C.byref(nb))
            return s, nb.value
        slot = 3 * W * H
        out, sizes = np.zeros((N, slot), np.uint8), np.zeros(N, np.uint32)
        s = self.lib.ref_time_encode_keep(_u8(frames), N, H * W * 3, reps, C.byref(nb), _u8(out), slot, sizes.ctypes.data_as(C.POINTER(C.c_uint32)))
        return s, nb.value, [out[i, : sizes[i]].tobytes() for i in range(N)]

    def time_loop(self, frames):
        N, H, W, _ = frames.shape
        self.lib.ref_set_dims(W, H)
        nr, nb = C.c_int(0), C.c_size_t(0)
        s = self.lib.ref_time_loop(_u8(frames), N, H * W * 3, C.byref(nr), C.byref(nb))
        return s, nr.value, nb.value

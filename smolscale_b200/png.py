"""ctypes mirror of libsmolpng.so (include/smol-png.h): PNG file I/O either side of the scaling path.

The reference keeps this in its test program (png.c:159-209, used by `test ... generate`,
test.c:1303-1371).  Pixels are RGBA8, unassociated alpha, as numpy uint8 arrays (h, w, 4)."""
import ctypes
import os

import numpy as np

from . import _build

EXPORTED_SYMBOLS = ["smol_png_decode_mem", "smol_png_encode_mem", "smol_png_load", "smol_png_save",
                    "smol_png_strerror", "smoltest_load_image", "smoltest_save_image"]


class PngError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, message)
        self.code = code


class PngInfo(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("bit_depth", ctypes.c_uint8),
                ("color_type", ctypes.c_uint8), ("interlace", ctypes.c_uint8), ("has_trns", ctypes.c_uint8)]


_lib = None
_libc = ctypes.CDLL(None)
_libc.free.argtypes = [ctypes.c_void_p]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.PNG_LIB_PATH):
            raise RuntimeError("libsmolpng.so has not been built: run `python __graft_entry__.py`")
        L = ctypes.CDLL(_build.PNG_LIB_PATH)
        L.smol_png_decode_mem.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32),
                                          ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.POINTER(PngInfo)]
        L.smol_png_encode_mem.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.POINTER(ctypes.c_size_t)]
        L.smol_png_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                                    ctypes.POINTER(ctypes.c_void_p)]
        L.smol_png_save.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        L.smol_png_strerror.restype = ctypes.c_char_p
        L.smol_png_strerror.argtypes = [ctypes.c_int]
        L.smoltest_load_image.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint),
                                          ctypes.POINTER(ctypes.c_void_p)]
        L.smoltest_save_image.restype = None
        L.smoltest_save_image.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint]
        _lib = L
    return _lib


def _check(err):
    if err != 0:
        raise PngError(err, lib().smol_png_strerror(err).decode())


def _take(ptr, w, h):
    try:
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(h, w, 4)).copy()
    finally:
        _libc.free(ptr)


def decode(data, with_info=False):
    """PNG bytes -> (h, w, 4) uint8 RGBA."""
    w, h, out, info = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_void_p(), PngInfo()
    _check(lib().smol_png_decode_mem(data, len(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(out),
                                     ctypes.byref(info)))
    img = _take(out, w.value, h.value)
    return (img, info) if with_info else img


def encode(pixels, level=5):
    """(h, w, 4) RGBA or (h, w, 3) RGB uint8 array (rows may be strided) -> PNG bytes."""
    a = np.asarray(pixels)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] not in (3, 4) or a.strides[2] != 1 or a.strides[1] != a.shape[2]:
        a = np.ascontiguousarray(a, dtype=np.uint8)
    out, size = ctypes.c_void_p(), ctypes.c_size_t()
    _check(lib().smol_png_encode_mem(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], a.shape[2], level,
                                     ctypes.byref(out), ctypes.byref(size)))
    try:
        return ctypes.string_at(out, size.value)
    finally:
        _libc.free(out)


def load_image(file_name):
    """smol_png_load: any PNG file -> (h, w, 4) uint8 RGBA."""
    w, h, out = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_void_p()
    _check(lib().smol_png_load(os.fsencode(file_name), ctypes.byref(w), ctypes.byref(h), ctypes.byref(out)))
    return _take(out, w.value, h.value)


def save_image(file_name, rgba):
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 4
    _check(lib().smol_png_save(os.fsencode(file_name), a.ctypes.data, a.shape[1], a.shape[0], a.strides[0]))

"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the CPU checkers.

* ``restatement()``  -> oracle/liboracle.so, our plain-C restatement (smol_oracle.c), built from
  committed source by ``make -C oracle oracle`` (``__graft_entry__.build()`` does it).
* ``reference(avx2=False)`` -> oracle/_ref/libsmolref[_avx2].so, the UNMODIFIED reference compiled
  from /root/reference by ``make -C oracle ref``; present only where it was built (the build
  container) or where the prebuilt file travelled (the GPU box).  Returns None when absent.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (smolscale_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# SmolPixelType numbering (reference smolscale.h:14-35)
RGBA8_P, BGRA8_P, ARGB8_P, ABGR8_P, RGBA8_U, BGRA8_U, ARGB8_U, ABGR8_U, RGB8, BGR8 = range(10)
PIXEL_TYPE_NAMES = ["RGBA8_P", "BGRA8_P", "ARGB8_P", "ABGR8_P",
                    "RGBA8_U", "BGRA8_U", "ARGB8_U", "ABGR8_U", "RGB8", "BGR8"]


def bpp(pixel_type):
    return 3 if pixel_type >= RGB8 else 4


def build(ref=True):
    """(Re)build liboracle.so / libref_harness.so and, when the reference tree exists, _ref/."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


_SIMPLE_ARGS = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                ctypes.c_uint8]


class _Scaler:
    """Common numpy front end: every checker exposes scale_simple / scale_rows on byte buffers."""

    def out_buffer(self, type_out, w_out, h_out, stride_out=None, fill=0xCD):
        stride_out = stride_out or w_out * bpp(type_out)
        n = stride_out * (h_out - 1) + w_out * bpp(type_out) if h_out else 0
        return np.full(n, fill, dtype=np.uint8), stride_out

    def scale_simple(self, src, type_in, w_in, h_in, stride_in, type_out, w_out, h_out,
                     stride_out=None, srgb=0, fill=0xCD):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        out, stride_out = self.out_buffer(type_out, w_out, h_out, stride_out, fill)
        self._simple(src.ctypes.data, type_in, w_in, h_in, stride_in,
                     out.ctypes.data, type_out, w_out, h_out, stride_out, srgb)
        return out


class Restatement(_Scaler):
    class _Plan(ctypes.Structure):
        _fields_ = [("w_in", ctypes.c_uint32), ("h_in", ctypes.c_uint32),
                    ("w_out", ctypes.c_uint32), ("h_out", ctypes.c_uint32),
                    ("type_in", ctypes.c_int), ("type_out", ctypes.c_int),
                    ("filter_h", ctypes.c_int), ("filter_v", ctypes.c_int),
                    ("halvings_h", ctypes.c_uint32), ("halvings_v", ctypes.c_uint32),
                    ("bilin_w", ctypes.c_uint32), ("bilin_h", ctypes.c_uint32),
                    ("storage_bits", ctypes.c_int), ("mid", ctypes.c_int),
                    ("span_mul_x", ctypes.c_uint32), ("span_mul_y", ctypes.c_uint32),
                    ("tab_x", ctypes.POINTER(ctypes.c_uint16)), ("tab_y", ctypes.POINTER(ctypes.c_uint16)),
                    ("n_tab_x", ctypes.c_uint32), ("n_tab_y", ctypes.c_uint32)]

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = ctypes.CDLL(path)
        self.lib.oracle_scale_simple.argtypes = _SIMPLE_ARGS
        self.lib.oracle_scale_simple.restype = None
        self.lib.oracle_plan_init.argtypes = [ctypes.POINTER(self._Plan), ctypes.c_int, ctypes.c_uint32,
                                              ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                              ctypes.c_uint32, ctypes.c_uint8]
        self.lib.oracle_plan_free.argtypes = [ctypes.POINTER(self._Plan)]
        self.lib.oracle_scale_rows.argtypes = [ctypes.POINTER(self._Plan), ctypes.c_void_p, ctypes.c_uint32,
                                               ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                               ctypes.c_uint32]
        self._simple = self.lib.oracle_scale_simple

    def plan(self, type_in, w_in, h_in, type_out, w_out, h_out, srgb=0):
        """Plan as a dict (tables copied to numpy) -- used to check the product's host logic."""
        p = self._Plan()
        self.lib.oracle_plan_init(ctypes.byref(p), type_in, w_in, h_in, type_out, w_out, h_out, srgb)
        d = {f: getattr(p, f) for f, _ in self._Plan._fields_ if not f.startswith("tab_")}
        d["tab_x"] = (np.ctypeslib.as_array(p.tab_x, shape=(p.n_tab_x * 2,)).copy()
                      if p.n_tab_x else np.zeros(0, np.uint16))
        d["tab_y"] = (np.ctypeslib.as_array(p.tab_y, shape=(p.n_tab_y * 2,)).copy()
                      if p.n_tab_y else np.zeros(0, np.uint16))
        self.lib.oracle_plan_free(ctypes.byref(p))
        return d

    def scale_rows(self, src, type_in, w_in, h_in, stride_in, type_out, w_out, h_out,
                   first, n, stride_out=None, srgb=0, fill=0xCD):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        out, stride_out = self.out_buffer(type_out, w_out, n, stride_out, fill)
        p = self._Plan()
        self.lib.oracle_plan_init(ctypes.byref(p), type_in, w_in, h_in, type_out, w_out, h_out, srgb)
        self.lib.oracle_scale_rows(ctypes.byref(p), src.ctypes.data, stride_in, out.ctypes.data,
                                   stride_out, first, n)
        self.lib.oracle_plan_free(ctypes.byref(p))
        return out


class Reference(_Scaler):
    """The unmodified reference library (generic-only or AVX2 build)."""

    def __init__(self, path):
        self.path = path
        self.lib = ctypes.CDLL(path)
        self.lib.smol_scale_simple.argtypes = _SIMPLE_ARGS
        self.lib.smol_scale_simple.restype = None
        self.lib.smol_scale_new.argtypes = _SIMPLE_ARGS
        self.lib.smol_scale_new.restype = ctypes.c_void_p
        self.lib.smol_scale_batch_full.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_uint32, ctypes.c_uint32]
        self.lib.smol_scale_batch_full.restype = None
        self.lib.smol_scale_destroy.argtypes = [ctypes.c_void_p]
        self.lib.smol_scale_destroy.restype = None
        self._simple = self.lib.smol_scale_simple

    def scale_rows(self, src, type_in, w_in, h_in, stride_in, type_out, w_out, h_out,
                   first, n, stride_out=None, srgb=0, fill=0xCD):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        out, stride_out = self.out_buffer(type_out, w_out, n, stride_out, fill)
        ctx = self.lib.smol_scale_new(src.ctypes.data, type_in, w_in, h_in, stride_in,
                                      None, type_out, w_out, h_out, stride_out, srgb)
        self.lib.smol_scale_batch_full(ctx, out.ctypes.data, first, n)
        self.lib.smol_scale_destroy(ctx)
        return out


_cache = {}


def restatement():
    if "restatement" not in _cache:
        _cache["restatement"] = Restatement()
    return _cache["restatement"]


def reference(avx2=False):
    key = "ref_avx2" if avx2 else "ref"
    if key not in _cache:
        path = os.path.join(HERE, "_ref", "libsmolref_avx2.so" if avx2 else "libsmolref.so")
        _cache[key] = Reference(path) if os.path.exists(path) else None
    return _cache[key]


class Harness:
    """ctypes front end of libref_harness.so (threaded row-band / image-batch driver)."""

    def __init__(self, lib_path):
        path = os.path.join(HERE, "libref_harness.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = ctypes.CDLL(path)
        self.lib.harness_open.argtypes = [ctypes.c_char_p]
        self.lib.harness_open.restype = ctypes.c_void_p
        self.lib.harness_close.argtypes = [ctypes.c_void_p]
        self.lib.harness_scale_threaded.argtypes = [ctypes.c_void_p] + _SIMPLE_ARGS + [ctypes.c_uint32, ctypes.c_uint32]
        self.lib.harness_scale_threaded.restype = ctypes.c_double
        self.lib.harness_scale_images.argtypes = [
            ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
            ctypes.c_uint8, ctypes.c_uint32, ctypes.c_uint32]
        self.lib.harness_scale_images.restype = ctypes.c_double
        self.h = self.lib.harness_open(lib_path.encode())
        if not self.h:
            raise OSError("cannot open " + lib_path)

    def scale_threaded(self, src, type_in, w_in, h_in, stride_in, out, type_out, w_out, h_out, stride_out,
                       srgb, n_threads, reps):
        return self.lib.harness_scale_threaded(self.h, src.ctypes.data, type_in, w_in, h_in, stride_in,
                                               out.ctypes.data, type_out, w_out, h_out, stride_out,
                                               srgb, n_threads, reps)

    def scale_images(self, src, in_image_bytes, type_in, w_in, h_in, stride_in,
                     out, out_image_bytes, type_out, w_out, h_out, stride_out, srgb, n_images, n_threads):
        return self.lib.harness_scale_images(self.h, src.ctypes.data, in_image_bytes, type_in, w_in, h_in,
                                             stride_in, out.ctypes.data, out_image_bytes, type_out,
                                             w_out, h_out, stride_out, srgb, n_images, n_threads)

    def close(self):
        if self.h:
            self.lib.harness_close(self.h)
            self.h = None

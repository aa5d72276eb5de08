"""smolscale_b200 -- Python front end of libsmolscale_cuda.so, the B200 (sm_100a) implementation
of the smolscale scaling pipeline behind the unchanged smolscale.h C API.

The product is the C-ABI shared library (include/smolscale.h, include/smolscale-cuda.h).  This
module is a thin ctypes mirror of that interface -- same names, same argument order and meaning
as the reference's smolscale.h:47-82 -- used by the tests and by bench.py.  Buffers may be numpy
arrays (host memory), torch tensors (host, pinned or CUDA memory) or raw integer addresses.

There is no CPU implementation here: if the shared library has not been built, importing the
binding raises; if no CUDA device is usable, the library aborts the process with a message
(the reference's own error behaviour is abort(), smolscale.c:779-780).
"""
import ctypes
import enum
import os

from . import _build

__all__ = ["PixelType", "ScaleCtx", "scale_simple", "scale_images", "lib", "plan_query",
           "device_count", "set_stream", "set_device", "synchronize", "set_multi_gpu", "stats", "reset_stats",
           "force_kernel", "kernel_launches", "bytes_per_pixel", "LIB_PATH"]

# SMOLSCALE_B200_LIB: load another build of the library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("SMOLSCALE_B200_LIB") or _build.LIB_PATH


class PixelType(enum.IntEnum):
    """SmolPixelType (reference smolscale.h:14-35)."""
    RGBA8_PREMULTIPLIED = 0
    BGRA8_PREMULTIPLIED = 1
    ARGB8_PREMULTIPLIED = 2
    ABGR8_PREMULTIPLIED = 3
    RGBA8_UNASSOCIATED = 4
    BGRA8_UNASSOCIATED = 5
    ARGB8_UNASSOCIATED = 6
    ABGR8_UNASSOCIATED = 7
    RGB8 = 8
    BGR8 = 9


def bytes_per_pixel(pixel_type):
    return 3 if int(pixel_type) >= 8 else 4


class PlanInfo(ctypes.Structure):
    _fields_ = [("filter_h", ctypes.c_int32), ("filter_v", ctypes.c_int32),
                ("halvings_h", ctypes.c_uint32), ("halvings_v", ctypes.c_uint32),
                ("bilin_w", ctypes.c_uint32), ("bilin_h", ctypes.c_uint32),
                ("storage_bits", ctypes.c_int32), ("mid", ctypes.c_int32),
                ("span_mul_x", ctypes.c_uint32), ("span_mul_y", ctypes.c_uint32),
                ("n_tab_x", ctypes.c_uint32), ("n_tab_y", ctypes.c_uint32),
                ("kernel_id", ctypes.c_int32), ("kernel_name", ctypes.c_char * 64)]


class Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64),
                ("d2h_bytes", ctypes.c_uint64), ("table_uploads", ctypes.c_uint64)]


POST_ROW_FUNC = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.c_void_p)

_JOB_ARGS = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
             ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
             ctypes.c_uint8]

# Every symbol include/smolscale.h and include/smolscale-cuda.h declare.
EXPORTED_SYMBOLS = [
    "smol_scale_simple", "smol_scale_new", "smol_scale_new_full", "smol_scale_destroy",
    "smol_scale_batch", "smol_scale_batch_full",
    "smol_cuda_device_count", "smol_cuda_set_device", "smol_cuda_set_stream", "smol_cuda_synchronize",
    "smol_cuda_set_multi_gpu",
    "smol_cuda_scale_images", "smol_cuda_plan_query", "smol_cuda_band_source_rows",
    "smol_cuda_get_stats", "smol_cuda_reset_stats", "smol_cuda_force_kernel", "smol_cuda_get_kernel_launches",
]

_lib = None


def lib():
    """The loaded shared library (ctypes.CDLL).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libsmolscale_cuda.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python smolscale_b200/_build.py`); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    L.smol_scale_simple.argtypes = _JOB_ARGS
    L.smol_scale_simple.restype = None
    L.smol_scale_new.argtypes = _JOB_ARGS
    L.smol_scale_new.restype = ctypes.c_void_p
    L.smol_scale_new_full.argtypes = _JOB_ARGS + [POST_ROW_FUNC, ctypes.c_void_p]
    L.smol_scale_new_full.restype = ctypes.c_void_p
    L.smol_scale_destroy.argtypes = [ctypes.c_void_p]
    L.smol_scale_destroy.restype = None
    L.smol_scale_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]
    L.smol_scale_batch.restype = None
    L.smol_scale_batch_full.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]
    L.smol_scale_batch_full.restype = None
    L.smol_cuda_device_count.restype = ctypes.c_int
    L.smol_cuda_set_device.argtypes = [ctypes.c_int]
    L.smol_cuda_set_stream.argtypes = [ctypes.c_void_p]
    L.smol_cuda_synchronize.argtypes = []
    L.smol_cuda_set_multi_gpu.argtypes = [ctypes.c_int]
    L.smol_cuda_scale_images.argtypes = [
        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
        ctypes.c_uint8, ctypes.c_uint32]
    L.smol_cuda_scale_images.restype = None
    L.smol_cuda_plan_query.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint8,
                                       ctypes.POINTER(PlanInfo), ctypes.c_void_p, ctypes.c_void_p]
    L.smol_cuda_plan_query.restype = None
    L.smol_cuda_band_source_rows.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    L.smol_cuda_get_stats.argtypes = [ctypes.POINTER(Stats)]
    L.smol_cuda_reset_stats.argtypes = []
    L.smol_cuda_force_kernel.argtypes = [ctypes.c_int]
    L.smol_cuda_get_kernel_launches.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_char_p),
                                                ctypes.c_int]
    L.smol_cuda_get_kernel_launches.restype = ctypes.c_int
    _lib = L
    return L


def _addr(buf):
    """Address of a buffer: int, numpy array, torch tensor, ctypes object or None."""
    if buf is None:
        return None
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):          # torch.Tensor
        return buf.data_ptr()
    if hasattr(buf, "ctypes"):            # numpy.ndarray
        return buf.ctypes.data
    return ctypes.addressof(buf)


def scale_simple(pixels_in, pixel_type_in, width_in, height_in, rowstride_in,
                 pixels_out, pixel_type_out, width_out, height_out, rowstride_out, with_srgb=0):
    """smol_scale_simple (reference smolscale.h:47-51)."""
    lib().smol_scale_simple(_addr(pixels_in), int(pixel_type_in), width_in, height_in, rowstride_in,
                            _addr(pixels_out), int(pixel_type_out), width_out, height_out, rowstride_out,
                            int(with_srgb))


def scale_images(pixels_in, image_stride_in, pixel_type_in, width_in, height_in, rowstride_in,
                 pixels_out, image_stride_out, pixel_type_out, width_out, height_out, rowstride_out,
                 with_srgb, n_images):
    """smol_cuda_scale_images: many same-shaped device-resident images in one launch."""
    lib().smol_cuda_scale_images(_addr(pixels_in), image_stride_in, int(pixel_type_in), width_in, height_in,
                                 rowstride_in, _addr(pixels_out), image_stride_out, int(pixel_type_out),
                                 width_out, height_out, rowstride_out, int(with_srgb), n_images)


class ScaleCtx:
    """SmolScaleCtx with the batch API (reference smolscale.h:55-82)."""

    def __init__(self, pixels_in, pixel_type_in, width_in, height_in, rowstride_in,
                 pixels_out, pixel_type_out, width_out, height_out, rowstride_out, with_srgb=0,
                 post_row_func=None, user_data=None):
        self._keep = (pixels_in, pixels_out)
        self._cb = None
        args = (_addr(pixels_in), int(pixel_type_in), width_in, height_in, rowstride_in,
                _addr(pixels_out), int(pixel_type_out), width_out, height_out, rowstride_out,
                int(with_srgb))
        if post_row_func is None:
            self._ctx = lib().smol_scale_new(*args)
        else:
            self._cb = POST_ROW_FUNC(post_row_func)
            self._ctx = lib().smol_scale_new_full(*args, self._cb, user_data)

    def batch(self, first_outrow, n_outrows):
        lib().smol_scale_batch(self._ctx, first_outrow, n_outrows)

    def batch_full(self, outrows_dest, first_outrow, n_outrows):
        lib().smol_scale_batch_full(self._ctx, _addr(outrows_dest), first_outrow, n_outrows)

    def band_source_rows(self, first_outrow, n_outrows):
        a, b = ctypes.c_uint32(), ctypes.c_uint32()
        lib().smol_cuda_band_source_rows(self._ctx, first_outrow, n_outrows, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def destroy(self):
        if self._ctx:
            lib().smol_scale_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def plan_query(pixel_type_in, width_in, height_in, pixel_type_out, width_out, height_out, with_srgb=0,
               tables=False):
    """Host-side plan (no GPU needed): what filters / encoding / tables the job gets."""
    import numpy as np
    info = PlanInfo()
    lib().smol_cuda_plan_query(int(pixel_type_in), width_in, height_in, int(pixel_type_out),
                               width_out, height_out, int(with_srgb), ctypes.byref(info), None, None)
    d = {name: getattr(info, name) for name, _ in PlanInfo._fields_}
    d["kernel_name"] = info.kernel_name.decode()
    if tables:
        tx = np.zeros(info.n_tab_x * 2, np.uint16)
        ty = np.zeros(info.n_tab_y * 2, np.uint16)
        lib().smol_cuda_plan_query(int(pixel_type_in), width_in, height_in, int(pixel_type_out),
                                   width_out, height_out, int(with_srgb), ctypes.byref(info),
                                   tx.ctypes.data if tx.size else None, ty.ctypes.data if ty.size else None)
        d["tab_x"], d["tab_y"] = tx, ty
    return d


def device_count():
    return lib().smol_cuda_device_count()


def set_device(device):
    lib().smol_cuda_set_device(device)


def set_stream(cuda_stream):
    """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or None."""
    lib().smol_cuda_set_stream(cuda_stream)


def synchronize():
    lib().smol_cuda_synchronize()


def set_multi_gpu(n_devices):
    """How many GPUs one host-memory call may be spread over (1 = off, 0 = all visible)."""
    lib().smol_cuda_set_multi_gpu(n_devices)


def stats():
    s = Stats()
    lib().smol_cuda_get_stats(ctypes.byref(s))
    return {name: getattr(s, name) for name, _ in Stats._fields_}


def kernel_launches():
    """Launches per kernel family since the last reset_stats(): {family name: count}."""
    counts = (ctypes.c_uint64 * 16)()
    names = (ctypes.c_char_p * 16)()
    n = lib().smol_cuda_get_kernel_launches(counts, names, 16)
    return {names[k].decode(): int(counts[k]) for k in range(n)}


def reset_stats():
    lib().smol_cuda_reset_stats()


def force_kernel(kernel_id):
    lib().smol_cuda_force_kernel(kernel_id)

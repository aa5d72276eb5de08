"""CPU tests of libsmolpng.so (include/smol-png.h): the PNG I/O either side of the scaling path
(reference png.c:34-209).  The independent checker is Pillow's PNG codec plus a
small PNG writer below for the shapes Pillow cannot produce (Adam7, 16-bit RGB/RGBA, every row
filter forced)."""
import io
import os
import re
import struct
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

Image = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def png():
    from smolscale_b200 import _build
    _build.build_png()
    from smolscale_b200 import png as mod
    return mod


def _rng_image(w, h, seed):
    rng = np.random.default_rng(seed)
    # smooth gradients + noise, so the filter heuristic picks different filters on different rows
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 3 + y) & 255, (x + y * 5) & 255, (x * y) & 255, (255 - x - y) & 255], axis=2).astype(np.uint8)
    noise = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    img[h // 2:] = noise[h // 2:]
    return img


def _pil_to_rgba(im):
    return np.asarray(im.convert("RGBA"))


# ---- a minimal PNG writer for the test (not shared with the product) ----

def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def _filter_rows(rows, step, choose):
    """rows: list of bytes; choose(y) -> filter type."""
    out = bytearray()
    prev = bytes(len(rows[0])) if rows else b""
    for y, row in enumerate(rows):
        f = choose(y)
        out.append(f)
        for i, v in enumerate(row):
            a = row[i - step] if i >= step else 0
            b = prev[i]
            c = prev[i - step] if i >= step else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[f]
            out.append((v - pred) & 255)
        prev = row
    return bytes(out)


def _pack_samples(samples, depth):
    """samples: (h, w, channels) integer array -> list of row bytes at `depth` bits per sample."""
    h, w, c = samples.shape
    rows = []
    for y in range(h):
        flat = samples[y].reshape(-1)
        if depth == 16:
            rows.append(b"".join(struct.pack(">H", int(v)) for v in flat))
        elif depth == 8:
            rows.append(bytes(int(v) for v in flat))
        else:
            bits = "".join(format(int(v), "0%db" % depth) for v in flat)
            bits += "0" * (-len(bits) % 8)
            rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
    return rows


ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def write_png(samples, color_type, depth, interlace=False, choose=lambda y: y % 5, extra=b"", idat_split=None):
    h, w, c = samples.shape
    step = max(1, c * depth // 8)
    if interlace:
        raw = b""
        for x0, y0, dx, dy in ADAM7:
            sub = samples[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += _filter_rows(_pack_samples(sub, depth), step, choose)
    else:
        raw = _filter_rows(_pack_samples(samples, depth), step, choose)
    z = zlib.compress(raw, 6)
    pieces = [z] if not idat_split else [z[i:i + idat_split] for i in range(0, len(z), idat_split)]
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, int(interlace)))
            + extra + b"".join(_chunk(b"IDAT", p) for p in pieces) + _chunk(b"IEND", b""))


# ---- decode ----

@pytest.mark.parametrize("mode", ["RGBA", "RGB", "L", "LA", "1", "I;16"])
def test_decode_matches_pillow_direct_modes(png, mode):
    rgba = _rng_image(37, 29, 1)
    if mode == "I;16":
        im = Image.fromarray(rgba[:, :, 0].astype(np.uint16) * 257 + 13)
        want = np.dstack([np.asarray(im) >> 8] * 3 + [np.full((29, 37), 255)]).astype(np.uint8)
    else:
        im = Image.fromarray(rgba, "RGBA").convert(mode)
        want = _pil_to_rgba(im)
    buf = io.BytesIO()
    im.save(buf, "PNG")
    got, info = png.decode(buf.getvalue(), with_info=True)
    assert (info.width, info.height) == (37, 29)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
def test_decode_palette_with_transparency(png, bits):
    rng = np.random.default_rng(bits)
    n = 1 << bits
    idx = rng.integers(0, n, size=(23, 45), dtype=np.uint8)
    pal = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    alpha = rng.integers(0, 256, size=(max(1, n // 2),), dtype=np.uint8)     # tRNS shorter than the palette
    im = Image.fromarray(idx, "P")
    im.putpalette(pal.tobytes())
    buf = io.BytesIO()
    im.save(buf, "PNG", bits=bits, transparency=alpha.tobytes())
    got, info = png.decode(buf.getvalue(), with_info=True)
    assert info.color_type == 3 and info.bit_depth == bits and info.has_trns
    full_alpha = np.concatenate([alpha, np.full(n - alpha.size, 255, np.uint8)])
    want = np.dstack([pal[idx], full_alpha[idx]])
    assert np.array_equal(got, want)


def test_decode_colour_keys(png):
    rgb = _rng_image(20, 11, 3)[:, :, :3].copy()
    rgb[3, 4] = (9, 8, 7)
    rgb[10, 19] = (9, 8, 7)
    data = write_png(rgb, 2, 8, extra=_chunk(b"tRNS", struct.pack(">HHH", 9, 8, 7)))
    got = png.decode(data)
    want = np.dstack([rgb, np.where((rgb == (9, 8, 7)).all(axis=2), 0, 255)]).astype(np.uint8)
    assert np.array_equal(got, want)
    grey = (np.arange(16 * 5).reshape(5, 16, 1) % 16)
    got = png.decode(write_png(grey, 0, 4, extra=_chunk(b"tRNS", struct.pack(">H", 5))))
    g8 = (grey[:, :, 0] * 17).astype(np.uint8)
    assert np.array_equal(got, np.dstack([g8, g8, g8, np.where(grey[:, :, 0] == 5, 0, 255)]).astype(np.uint8))


@pytest.mark.parametrize("color_type,channels,depth", [(6, 4, 8), (6, 4, 16), (2, 3, 16), (2, 3, 8), (4, 2, 8),
                                                        (4, 2, 16), (0, 1, 16), (0, 1, 2), (0, 1, 1), (3, 1, 4)])
@pytest.mark.parametrize("interlace", [False, True])
def test_decode_all_filters_and_adam7(png, color_type, channels, depth, interlace):
    rng = np.random.default_rng(color_type * 100 + depth)
    for w, h in ((1, 1), (3, 2), (9, 9), (33, 17)):
        samples = rng.integers(0, 1 << depth, size=(h, w, channels))
        extra = b""
        if color_type == 3:
            pal = rng.integers(0, 256, size=(16, 3), dtype=np.uint8)
            extra = _chunk(b"PLTE", pal.tobytes())
        data = write_png(samples, color_type, depth, interlace=interlace, extra=extra, idat_split=7)
        got, info = png.decode(data, with_info=True)
        assert info.interlace == int(interlace)
        s8 = (samples >> 8 if depth == 16 else samples * 255 // ((1 << depth) - 1) if depth < 8 else samples).astype(np.uint8)
        full = np.full((h, w), 255, np.uint8)
        if color_type == 6:
            want = s8
        elif color_type == 2:
            want = np.dstack([s8, full])
        elif color_type == 4:
            want = np.dstack([s8[:, :, 0]] * 3 + [s8[:, :, 1]])
        elif color_type == 0:
            want = np.dstack([s8[:, :, 0]] * 3 + [full])
        else:
            want = np.dstack([pal[samples[:, :, 0]], full])
        assert np.array_equal(got, want), (w, h)
        # Pillow agrees wherever it reads the same thing (it keeps 16-bit grey as I;16 and has its own 16-bit rounding)
        if depth <= 8:
            assert np.array_equal(_pil_to_rgba(Image.open(io.BytesIO(data))), want)


def test_decode_rejects_damage(png):
    good = write_png(_rng_image(8, 8, 5), 6, 8)
    assert png.decode(good).shape == (8, 8, 4)
    with pytest.raises(png.PngError) as e:
        png.decode(b"not a png at all")
    assert e.value.code == 2
    bad = bytearray(good)
    bad[40] ^= 1                                            # inside IDAT: CRC mismatch
    with pytest.raises(png.PngError) as e:
        png.decode(bytes(bad))
    assert e.value.code == 3
    with pytest.raises(png.PngError):
        png.decode(good[:len(good) // 2])                   # truncated
    # a filter byte of 5 is invalid
    raw = b"".join(b"\x05" + bytes(8 * 4) for _ in range(8))
    data = (good[:8] + _chunk(b"IHDR", struct.pack(">IIBBBBB", 8, 8, 8, 6, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(raw))
            + _chunk(b"IEND", b""))
    with pytest.raises(png.PngError):
        png.decode(data)
    # a zlib stream shorter than the image
    data = (good[:8] + _chunk(b"IHDR", struct.pack(">IIBBBBB", 8, 8, 8, 6, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(raw[:100]))
            + _chunk(b"IEND", b""))
    with pytest.raises(png.PngError):
        png.decode(data)
    # an unknown critical chunk
    data = good[:33] + _chunk(b"ABCD", b"x") + good[33:]
    with pytest.raises(png.PngError) as e:
        png.decode(data)
    assert e.value.code == 4
    # an unknown ancillary chunk is skipped
    assert np.array_equal(png.decode(good[:33] + _chunk(b"teSt", b"x") + good[33:]), png.decode(good))


# ---- encode ----

@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (64, 64), (301, 47)])
@pytest.mark.parametrize("channels", [3, 4])
def test_encode_read_back_by_pillow_and_by_itself(png, w, h, channels):
    img = _rng_image(w, h, w + h)[:, :, :channels]
    data = png.encode(img)
    im = Image.open(io.BytesIO(data))
    assert im.mode == ("RGBA" if channels == 4 else "RGB") and im.size == (w, h)
    assert np.array_equal(np.asarray(im), img)
    back = png.decode(data)
    assert np.array_equal(back[:, :, :channels], img)
    if channels == 3:
        assert (back[:, :, 3] == 255).all()


def test_encode_strided_rows_and_levels(png):
    big = _rng_image(80, 40, 9)
    view = big[5:35, 10:70]                                  # rows at a pitch of 320 bytes, 60 pixels wide
    sizes = []
    for level in (0, 1, 5, 9):
        data = png.encode(view, level=level)
        assert np.array_equal(png.decode(data), view)
        sizes.append(len(data))
    assert sizes[0] > sizes[2]                               # level 0 = stored blocks
    smooth = np.zeros((64, 64, 4), np.uint8)
    smooth[:, :, 0] = np.arange(64)[None, :] * 4
    smooth[:, :, 1] = np.arange(64)[:, None] * 4
    smooth[:, :, 3] = 255
    data = png.encode(smooth)
    assert len(data) < 64 * 64 * 4 // 20                     # the Sub / Up filters make gradients compress
    assert np.array_equal(png.decode(data), smooth)


def test_encode_large_image_splits_idat(png):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(600, 700, 4), dtype=np.uint8)      # incompressible: > 1 MiB of IDAT
    data = png.encode(img, level=1)
    assert data.count(b"IDAT") >= 2
    assert np.array_equal(png.decode(data), img)
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(data))), img)


# ---- files, and the reference's helper interface ----

def test_files_and_reference_helpers(png, tmp_path):
    import ctypes
    img = _rng_image(50, 20, 11)
    path = str(tmp_path / "a.png")
    png.save_image(path, img)
    assert np.array_equal(png.load_image(path), img)
    assert np.array_equal(np.asarray(Image.open(path)), img)
    with pytest.raises(png.PngError) as e:
        png.load_image(str(tmp_path / "missing.png"))
    assert e.value.code == 1

    L = png.lib()
    # smoltest_save_image names the file <prefix>-WWWW-HHHH.png (png.c:204)
    prefix = str(tmp_path / "out")
    L.smoltest_save_image(prefix.encode(), img.ctypes.data, 50, 20)
    named = prefix + "-0050-0020.png"
    assert os.path.exists(named)
    w, h, data = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_void_p()
    assert L.smoltest_load_image(named.encode(), ctypes.byref(w), ctypes.byref(h), ctypes.byref(data)) == 1
    assert (w.value, h.value) == (50, 20)
    got = np.ctypeslib.as_array(ctypes.cast(data, ctypes.POINTER(ctypes.c_uint8)), shape=(20, 50, 4)).copy()
    assert np.array_equal(got, img)

    # like the reference (png.c:90-97) the helper refuses anything but 8-bit RGBA, with a message, by abort()
    rgb_path = str(tmp_path / "rgb.png")
    Image.fromarray(img[:, :, :3].copy(), "RGB").save(rgb_path)
    code = ("import ctypes,sys; sys.path.insert(0, %r); from smolscale_b200 import png; L = png.lib(); "
            "w = ctypes.c_uint(); h = ctypes.c_uint(); d = ctypes.c_void_p(); "
            "L.smoltest_load_image(%r.encode(), ctypes.byref(w), ctypes.byref(h), ctypes.byref(d))" % (ROOT, rgb_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "PNG_COLOR_TYPE_RGB" in r.stderr


def test_library_exports_every_declared_symbol(png):
    text = open(os.path.join(ROOT, "include", "smol-png.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(smol_png_[a-z_]+|smoltest_[a-z_]+)\s*\(", text))
    assert declared == set(png.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(png.lib(), name) is not None
    out = subprocess.run(["ldd", os.path.join(ROOT, "smolscale_b200", "libsmolpng.so")], capture_output=True, text=True).stdout
    assert "cuda" not in out and "oracle" not in out

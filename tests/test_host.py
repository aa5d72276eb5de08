"""CPU tests of the product's host logic: the C-ABI library loads, exports every declared symbol,
and plans jobs (filters, encoding, fixed-point tables) exactly like the reference.  No compute
calls are made -- those need a GPU."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(sb):
    lib = sb.lib()
    declared = set()
    for header in ("smolscale.h", "smolscale-cuda.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(smol_[a-z_]+)\s*\(", text))
    assert declared == set(sb.EXPORTED_SYMBOLS)
    for name in sorted(declared):
        assert getattr(lib, name) is not None


def test_enum_values_are_abi(sb):
    names = ["RGBA8_PREMULTIPLIED", "BGRA8_PREMULTIPLIED", "ARGB8_PREMULTIPLIED", "ABGR8_PREMULTIPLIED",
             "RGBA8_UNASSOCIATED", "BGRA8_UNASSOCIATED", "ARGB8_UNASSOCIATED", "ABGR8_UNASSOCIATED",
             "RGB8", "BGR8"]
    hdr = open(os.path.join(ROOT, "include", "smolscale.h")).read()
    for i, n in enumerate(names):
        assert int(getattr(sb.PixelType, n)) == i
        assert re.search(r"SMOL_PIXEL_%s\s*=\s*%d\b" % (n, i), hdr)
    assert re.search(r"SMOL_PIXEL_MAX\s*=\s*10\b", hdr)


def test_product_does_not_touch_the_oracle():
    """The product path must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "smolscale_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "liboracle" not in text and "smol_oracle" not in text, f
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libsmolscale_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "smolref" not in out


def test_plan_matches_restatement(sb, restatement):
    dims = [(1, 1), (1, 7), (2, 3), (3, 2), (5, 5), (7, 3), (10, 4), (17, 2), (16, 2), (100, 11), (100, 12),
            (520, 2), (2, 520), (65535, 1), (65535, 7), (2, 65535), (65534, 65535), (65535, 65534),
            (9000, 1000), (4320, 450), (7680, 800), (3840, 1920), (1024, 4096), (2048, 256), (16383, 2)]
    for (a, b), (ti, to), srgb in itertools.product(dims, [(0, 0), (4, 5), (8, 2), (6, 9), (1, 5)], (0, 1)):
        p = sb.plan_query(ti, a, a, to, b, b, srgb, tables=True)
        q = restatement.plan(ti, a, a, to, b, b, srgb)
        for k in ("filter_h", "filter_v", "halvings_h", "halvings_v", "bilin_w", "bilin_h",
                  "storage_bits", "mid", "span_mul_x", "span_mul_y", "n_tab_x", "n_tab_y"):
            assert p[k] == q[k], (a, b, ti, to, srgb, k)
        assert np.array_equal(p["tab_x"], q["tab_x"]) and np.array_equal(p["tab_y"], q["tab_y"])


class _RefCtx(ctypes.Structure):
    """Layout of the reference's private SmolScaleCtx (smolscale-private.h:280-312), x86-64."""
    _fields_ = [("pixels_in", ctypes.c_void_p), ("pixels_out", ctypes.c_void_p),
                ("width_in", ctypes.c_uint32), ("height_in", ctypes.c_uint32), ("rowstride_in", ctypes.c_uint32),
                ("width_out", ctypes.c_uint32), ("height_out", ctypes.c_uint32), ("rowstride_out", ctypes.c_uint32),
                ("pixel_type_in", ctypes.c_int), ("pixel_type_out", ctypes.c_int),
                ("filter_h", ctypes.c_int), ("filter_v", ctypes.c_int),
                ("storage_type", ctypes.c_int), ("gamma_type", ctypes.c_int),
                ("unpack_row_func", ctypes.c_void_p), ("pack_row_func", ctypes.c_void_p),
                ("hfilter_func", ctypes.c_void_p), ("vfilter_func", ctypes.c_void_p),
                ("post_row_func", ctypes.c_void_p), ("user_data", ctypes.c_void_p),
                ("precalc_x", ctypes.POINTER(ctypes.c_uint16)), ("precalc_y", ctypes.POINTER(ctypes.c_uint16)),
                ("span_mul_x", ctypes.c_uint32), ("span_mul_y", ctypes.c_uint32),
                ("precalc_x_storage", ctypes.c_void_p),
                ("width_bilin_out", ctypes.c_uint32), ("height_bilin_out", ctypes.c_uint32),
                ("width_halvings", ctypes.c_uint), ("height_halvings", ctypes.c_uint)]


@pytest.mark.ref
def test_plan_matches_reference_context(sb, reference):
    """Compare our planner with the reference's own context: filter classes, storage, span
    multipliers and the precalc tables themselves (x table is relative in the reference,
    generic:47, so it is prefix-summed before comparing)."""
    REF_COPY, REF_ONE, REF_BIL0, REF_BOX = 0, 1, 2, 9
    dims = [(1, 7), (5, 5), (7, 3), (10, 4), (17, 2), (100, 11), (100, 12), (520, 2), (65535, 7), (2, 65535),
            (65534, 65535), (9000, 1000), (4320, 450), (7680, 800), (3840, 1920), (1024, 4096), (2048, 256)]
    src = np.zeros(16, np.uint8)
    for (a, b), (ti, to), srgb in itertools.product(dims, [(0, 0), (4, 5), (8, 2)], (0, 1)):
        ctxp = reference.lib.smol_scale_new(src.ctypes.data, ti, a, a, a * 4, None, to, b, b, b * 4, srgb)
        ctx = _RefCtx.from_address(ctxp)
        p = sb.plan_query(ti, a, a, to, b, b, srgb, tables=True)

        def cls(f):
            return {REF_COPY: 0, REF_ONE: 1, REF_BOX: 3}.get(f, 2)
        assert cls(ctx.filter_h) == p["filter_h"] and cls(ctx.filter_v) == p["filter_v"]
        assert {2: 64, 3: 128}[ctx.storage_type] == p["storage_bits"]
        if p["filter_h"] == 2:
            assert ctx.filter_h - REF_BIL0 == p["halvings_h"] and ctx.width_bilin_out == p["bilin_w"]
            n = p["bilin_w"]
            rx = np.ctypeslib.as_array(ctx.precalc_x, shape=(n * 2,)).astype(np.uint32)
            assert np.array_equal(np.cumsum(rx[0::2]), p["tab_x"][0::2])
            assert np.array_equal(rx[1::2], p["tab_x"][1::2])
        if p["filter_h"] == 3:
            assert ctx.span_mul_x == p["span_mul_x"]
            n = b + 1
            rx = np.ctypeslib.as_array(ctx.precalc_x, shape=(n * 2,)).astype(np.uint32)
            # reference x entries: (whole pixels between the edge pixels, F); box i starts where i-1 ended
            starts = np.concatenate([[0], np.cumsum(rx[0:2 * b:2] + 1)])
            assert np.array_equal(starts[:b], p["tab_x"][0:2 * b:2])
            assert np.array_equal(rx[1:2 * b:2], p["tab_x"][1:2 * b:2])
            assert starts[b] == p["tab_x"][2 * b]
        if p["filter_v"] in (2, 3):
            n = p["n_tab_y"]
            ry = np.ctypeslib.as_array(ctx.precalc_y, shape=(n * 2,))
            assert np.array_equal(ry, p["tab_y"])
            if p["filter_v"] == 3:
                assert ctx.span_mul_y == p["span_mul_y"]
        reference.lib.smol_scale_destroy(ctxp)


def test_band_source_rows(sb):
    src = np.zeros(16, np.uint8)
    for (hi, ho) in [(2160, 1080), (4320, 450), (768, 3072), (2048, 256), (100, 100), (1, 50), (600, 7)]:
        ctx = sb.ScaleCtx(src, 0, 8, hi, 32, None, 0, 8, ho, 32, 0)
        p = sb.plan_query(0, 8, hi, 0, 8, ho, 0, tables=True)
        covered_lo, covered_hi = hi, -1
        for first in range(0, ho, max(1, ho // 7)):
            n = min(max(1, ho // 7), ho - first)
            r0, nr = ctx.band_source_rows(first, n)
            assert nr >= 1 and r0 + nr <= hi
            covered_lo, covered_hi = min(covered_lo, r0), max(covered_hi, r0 + nr - 1)
            if p["filter_v"] == 2:       # bilinear: rows ofs .. ofs + 1 of every sample in the band
                h = p["halvings_v"]
                ofs = p["tab_y"][0::2][(first << h):((first + n) << h)]
                assert r0 == ofs.min() and r0 + nr - 1 == min(int(ofs.max()) + 1, hi - 1)
            if p["filter_v"] == 3:
                ofs = p["tab_y"][0::2]
                assert r0 == ofs[first] and r0 + nr - 1 == ofs[first + n]
        if p["filter_v"] in (2, 3) and hi > ho:
            assert covered_lo == 0
        ctx.destroy()


def test_dispatch_table():
    """The dispatcher's choice of kernel family for representative jobs (pure host logic, best-case
    alignment): the five BASELINE configurations and the measured crossovers documented in
    smol_cuda_pick_kernel / DESIGN.md section 5."""
    import smolscale_b200 as sb
    c = cases
    expect = [
        ("half2x", c.RGBA8_P, 1920, 1080, c.RGBA8_P, 960, 540, 0),          # cfg 1: exact 2:1
        ("half2x", c.BGRA8_P, 3840, 2160, c.BGRA8_U, 1920, 1080, 0),        # cfg 2: exact 2:1 + unpremultiply
        ("box", c.RGBA8_P, 7680, 4320, c.RGBA8_P, 800, 450, 1),             # cfg 3: box x box, linear light
        ("magb", c.RGB8, 1024, 768, c.RGB8, 4096, 3072, 0),                 # cfg 4: 24bpp -> 24bpp, 4x up
        ("half2x", c.ARGB8_P, 2048, 2048, c.ARGB8_P, 256, 256, 0),          # cfg 5: exact 8:1
        ("taps_direct", c.RGBA8_P, 1920, 1080, c.RGBA8_P, 3840, 2160, 0),   # 32bpp upscale: register kernel
        ("taps_direct", c.RGB8, 2560, 1440, c.RGB8, 3840, 2160, 0),         # 24bpp below 2x: register kernel
        ("taps_direct", c.RGBA8_P, 3840, 2160, c.BGRA8_U, 3839, 2159, 0),   # the reference's conv shape
        ("taps_direct", c.RGBA8_P, 3840, 2160, c.RGBA8_P, 1280, 720, 0),    # 3:1, one halving
        ("taps128", c.RGBA8_P, 3840, 2160, c.RGBA8_P, 1280, 720, 1),        # ... in linear light
        ("taps128", c.RGBA8_U, 3840, 2160, c.RGBA8_U, 3839, 2159, 0),       # unassociated -> unassociated, ~1:1
        ("tile128", c.RGBA8_U, 1920, 1080, c.RGBA8_U, 3840, 2160, 0),       # ... upscale
        ("box", c.RGB8, 7680, 4320, c.RGB8, 800, 450, 0),                   # box without linear light
        ("rows", c.RGBA8_P, 3000, 7, c.RGBA8_P, 11, 7, 0),                  # box on one axis only
        ("rows", c.RGBA8_P, 7680, 1080, c.RGBA8_P, 800, 540, 1),            # ... in linear light
        ("rows", c.RGBA8_P, 7680, 4320, c.RGBA8_P, 20, 12, 0),              # > 255:1 without linear light
        ("general", c.RGBA8_P, 9000, 1, c.RGBA8_P, 1, 1, 0),                # ... a 36 KB window per warp: the CTA-wide backstop
    ]
    for want, ti, wi, hi, to, wo, ho, srgb in expect:
        got = sb.plan_query(ti, wi, hi, to, wo, ho, srgb)["kernel_name"]
        assert got == want, ((ti, wi, hi, to, wo, ho, srgb), got, want)


def _header_tables(path, prefix):
    """{name: list of ints} for every `static const <type> <prefix><name>[N] = {...};` in a generated header."""
    text = open(path).read()
    out = {}
    for m in re.finditer(r"static const \w+ %s(\w+)\[(\d+)\] = \{(.*?)\};" % re.escape(prefix), text, flags=re.S):
        vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", m.group(3))]
        assert len(vals) == int(m.group(2)), m.group(1)
        out[m.group(1)] = vals
    return out


def test_lut_headers_equal_reference_data(reference):
    """The six data tables (reference smolscale.c:87-421) the product and the oracle are built with are,
    entry for entry, the data symbols of the compiled reference; two of them have no
    generator, so this equality is the only thing that pins them besides bit-exact output."""
    symbols = {"from_srgb": ("_smol_from_srgb_lut", ctypes.c_uint16, 256), "to_srgb": ("_smol_to_srgb_lut", ctypes.c_uint8, 2048),
               "inv_div_p8": ("_smol_inv_div_p8_lut", ctypes.c_uint32, 256), "inv_div_p8l": ("_smol_inv_div_p8l_lut", ctypes.c_uint32, 256),
               "inv_div_p16": ("_smol_inv_div_p16_lut", ctypes.c_uint32, 256), "inv_div_p16l": ("_smol_inv_div_p16l_lut", ctypes.c_uint32, 256)}
    product = _header_tables(os.path.join(ROOT, "smolscale_b200", "csrc", "smolscale-cuda-luts.h"), "smol_lut_")
    checker = _header_tables(os.path.join(ROOT, "oracle", "smol_oracle_luts.h"), "oracle_lut_")
    assert set(product) == set(symbols) == set(checker)
    for name, (sym, ct, n) in symbols.items():
        ref = list((ct * n).in_dll(reference.lib, sym))
        assert product[name] == ref, name
        assert checker[name] == ref, name
    # the two closed forms the survey found (ceil (2^16 / a), ceil (2^19 / a)) as an independent cross-check
    assert all(product["inv_div_p16"][a] == -(-(1 << 16) // a) for a in range(1, 256))
    assert all(product["inv_div_p16l"][a] == -(-(1 << 19) // a) for a in range(1, 256))


def test_tile_window_bound_of_the_halving_tile_kernel(sb):
    """launch_tile128h sizes a tile's shared-memory window as ceil (tile * dim_in / dim_out) + 5 pixels per axis
    (smolscale-cuda-kernels.cu); the kernel traps if a tile's first and last taps span more.  Brute force over the
    planner's own tables: every tile of every job stays inside the bound (the widest seen needs + 4)."""
    import random
    rnd = random.Random(1)
    dims = []
    for _ in range(400):
        d_out = rnd.choice([1, 2, 3, 5, 7, 30, 31, 33, 64, 100, 127, 200, 333, 640, 1279, 1280]) if rnd.random() < 0.5 \
            else rnd.randint(1, 2500)
        d_in = min(65535, int(d_out * rnd.uniform(2.0, 8.0)) + rnd.randint(-2, 2))
        if d_out * 2 < d_in <= d_out * 8:
            dims.append((d_in, d_out))
    assert len(dims) > 300
    worst = -100
    for d_in, d_out in dims:
        p = sb.plan_query(cases.RGBA8_P, d_in, d_in, cases.RGBA8_P, d_out, d_out, 1, tables=True)
        hh = p["halvings_h"]
        assert p["filter_h"] == 2 and hh >= 1 and p["kernel_name"] == "taps128"
        ofs = p["tab_x"][0::2].astype(np.int64)
        assert (np.diff(ofs) >= 0).all()                      # the kernel takes the first and the last tap as the window's ends
        for tile in (64, 32, 16, 8):
            uncapped = (tile * d_in + d_out - 1) // d_out + 5
            bound = min(uncapped, d_in)
            t0 = np.arange(0, d_out, tile)
            t1 = np.minimum(t0 + tile, d_out)
            span = np.minimum(ofs[(t1 << hh) - 1] + 1, d_in - 1) - ofs[t0 << hh] + 1
            assert (span <= bound).all(), (d_in, d_out, tile)
            worst = max(worst, int((span - uncapped).max()))
    assert worst <= -1                                         # a pixel of margin wherever the row itself is not the bound

"""Shared, seeded test-case generators (synthetic images and job matrices).

Every generator is deterministic in its arguments so the oracle, the compiled reference and the
CUDA path all see identical bytes on any machine.
"""
import itertools

import numpy as np

# SmolPixelType numbering (reference smolscale.h:14-35)
RGBA8_P, BGRA8_P, ARGB8_P, ABGR8_P, RGBA8_U, BGRA8_U, ARGB8_U, ABGR8_U, RGB8, BGR8 = range(10)
ALL_TYPES = list(range(10))
TYPE_NAMES = ["RGBA8_P", "BGRA8_P", "ARGB8_P", "ABGR8_P", "RGBA8_U", "BGRA8_U", "ARGB8_U", "ABGR8_U",
              "RGB8", "BGR8"]

IMAGE_MODES = ["random", "premul", "alpha_edges", "saturated", "gradient", "zero"]


def bpp(t):
    return 3 if t >= RGB8 else 4


def alpha_index(t):
    """Byte index of alpha within a pixel, or None."""
    if t >= RGB8:
        return None
    return 3 if (t & 3) < 2 else 0


def make_image(pixel_type, width, height, stride=None, mode="random", seed=0):
    """Returns a flat uint8 buffer of stride * (height - 1) + width * bpp bytes... padded to
    stride * height so every row is fully addressable."""
    b = bpp(pixel_type)
    stride = stride or width * b
    assert stride >= width * b
    rng = np.random.default_rng([seed, pixel_type, width, height, IMAGE_MODES.index(mode)])
    buf = rng.integers(0, 256, size=stride * height, dtype=np.uint8)
    px = np.lib.stride_tricks.as_strided(buf, shape=(height, width, b), strides=(stride, b, 1))
    ai = alpha_index(pixel_type)
    if mode == "premul" and ai is not None:
        # premultiplied-valid: colour <= alpha
        al = px[:, :, ai].astype(np.uint32)
        for c in range(4):
            if c != ai:
                px[:, :, c] = ((px[:, :, c].astype(np.uint32) * al + 127) // 255).astype(np.uint8)
    elif mode == "alpha_edges" and ai is not None:
        px[:, :, ai] = rng.choice(np.array([0, 1, 2, 127, 128, 254, 255], dtype=np.uint8), size=(height, width))
    elif mode == "saturated":
        px[:, :, :] = 0xFF
    elif mode == "zero":
        px[:, :, :] = 0
    elif mode == "gradient":
        xs = (np.arange(width, dtype=np.uint32) * 255 // max(width - 1, 1)).astype(np.uint8)
        ys = (np.arange(height, dtype=np.uint32) * 255 // max(height - 1, 1)).astype(np.uint8)
        for c in range(b):
            px[:, :, c] = (xs[None, :] if c % 2 == 0 else ys[:, None])
        if ai is not None:
            px[:, :, ai] = 255 - xs[None, :] // 2
    return buf


# 1-D (dim_in, dim_out) pairs chosen to hit every filter class and table edge:
#   one, copy, bilinear 0h/1h/2h (magnify, minify, exact 2^k), box 64bpp, box 128bpp (> 255),
#   integer box ratios (last-pixel clamp quirk), sRGB cut-off (> 8191).
AXIS_PAIRS = [
    (1, 1), (1, 5), (2, 2), (2, 3), (2, 7), (3, 2), (5, 5), (7, 3), (8, 4), (10, 4), (9, 2),
    (16, 2), (17, 2), (33, 4), (64, 8), (65, 8), (64, 7), (37, 41), (41, 37), (100, 11), (100, 12),
    (100, 3), (300, 1), (520, 2), (1000, 111), (90, 10), (2048, 8), (2041, 8),
]

# more ratio classes (the soak tool's list): long box spans next to mild bilinear ratios, so that random pairs
# of them mix box and bilinear axes with large accumulations
SOAK_AXIS_PAIRS = AXIS_PAIRS + [(640, 200), (641, 97), (333, 777), (1000, 3), (4000, 15), (12, 700), (255, 1), (256, 1),
                                (257, 1), (2040, 8), (2041, 8), (96, 12), (100, 50), (77, 154), (200, 15), (255, 16),
                                (1500, 100), (3000, 230), (1300, 87), (160, 640), (33, 1000)]

BIG_AXIS_PAIRS = [(9000, 1), (16400, 2), (65535, 1), (65535, 65534), (65534, 65535), (1, 65535), (2, 65535)]


def job_matrix(seed, n_jobs, axis_pairs=AXIS_PAIRS, max_pixels=300000):
    """Random jobs: (type_in, w_in, h_in, stride_in, type_out, w_out, h_out, stride_out, srgb, mode)."""
    rng = np.random.default_rng(seed)
    jobs = []
    while len(jobs) < n_jobs:
        wi, wo = axis_pairs[int(rng.integers(len(axis_pairs)))]
        hi, ho = axis_pairs[int(rng.integers(len(axis_pairs)))]
        if wi * hi > max_pixels or wo * ho > max_pixels:
            continue
        ti, to = int(rng.integers(10)), int(rng.integers(10))
        srgb = int(rng.integers(2))
        mode = IMAGE_MODES[int(rng.integers(len(IMAGE_MODES)))]
        si = wi * bpp(ti) + int(rng.choice([0, 0, 1, 3, 4, 16]))
        so = wo * bpp(to) + int(rng.choice([0, 0, 1, 3, 4, 16]))
        jobs.append((ti, wi, hi, si, to, wo, ho, so, srgb, mode))
    return jobs


# exact 2^k : 1 reductions (every bilinear weight 128): the packed-byte "half" kernel family
HALF_AXIS_PAIRS = [(2, 1), (4, 1), (8, 1), (8, 4), (10, 5), (16, 4), (28, 7), (24, 3), (64, 8), (62, 31),
                   (36, 9), (50, 25), (256, 32)]


def half_jobs(seed=5):
    """Jobs that are eligible for the half kernel (32bpp premultiplied / alpha-less-free source,
    32bpp destination) over every halving combination and ragged widths."""
    rng = np.random.default_rng(seed)
    jobs = []
    for (wi, wo) in HALF_AXIS_PAIRS:
        for (hi, ho) in HALF_AXIS_PAIRS:
            ti = int(rng.integers(0, 4))
            to = int(rng.integers(0, 8))
            mode = IMAGE_MODES[int(rng.integers(len(IMAGE_MODES)))]
            jobs.append((ti, wi, hi, wi * 4, to, wo, ho, wo * 4, 0, mode))
    return jobs


def all_type_pairs():
    return list(itertools.product(ALL_TYPES, ALL_TYPES))


# The five BASELINE.json configurations: (name, type_in, w_in, h_in, type_out, w_out, h_out, srgb, image mode)
BASELINE_CONFIGS = [
    ("cfg1_1080p_to_540p_rgba_premul", RGBA8_P, 1920, 1080, RGBA8_P, 960, 540, 0, "premul"),
    ("cfg2_4k_to_1080p_bgra_p_to_u", BGRA8_P, 3840, 2160, BGRA8_U, 1920, 1080, 0, "premul"),
    ("cfg3_8k_to_800x450_box_srgb", RGBA8_P, 7680, 4320, RGBA8_P, 800, 450, 1, "premul"),
    ("cfg4_rgb_1024x768_to_4096x3072", RGB8, 1024, 768, RGB8, 4096, 3072, 0, "random"),
    ("cfg5_2048sq_to_256sq_argb", ARGB8_P, 2048, 2048, ARGB8_P, 256, 256, 0, "premul"),
]

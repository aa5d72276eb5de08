"""CPU tests: the plain-C restatement (oracle/smol_oracle.c) against
  (a) the committed golden digests generated from the compiled reference, and
  (b) the compiled reference itself when oracle/_ref is present (build container / GPU box)."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "digests.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)["digests"]


def run_job(scaler, job):
    ti, wi, hi, si, to, wo, ho, so, srgb, mode, seed = job
    src = cases.make_image(ti, wi, hi, si, mode, seed)
    return scaler.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)


def test_restatement_matches_golden(restatement):
    golden = load_golden()
    assert len(golden) > 700
    bad = []
    for name, g in sorted(golden.items()):
        out = run_job(restatement, tuple(g["job"]))
        if hashlib.sha256(out.tobytes()).hexdigest() != g["sha256"]:
            bad.append(name)
    assert not bad, "restatement differs from reference digests: %s" % bad[:10]


def test_restatement_matches_reference_random(restatement, reference):
    for idx, job in enumerate(cases.job_matrix(99, 500)):
        ti, wi, hi, si, to, wo, ho, so, srgb, mode = job
        src = cases.make_image(ti, wi, hi, si, mode, seed=idx)
        a = reference.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
        b = restatement.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
        assert np.array_equal(a, b), job


def test_restatement_rows_match_reference_bands(restatement, reference):
    """Row batches are independent of how they are cut."""
    rng = np.random.default_rng(3)
    for ti, wi, hi, to, wo, ho, srgb in [(cases.RGBA8_P, 200, 150, cases.BGRA8_U, 77, 64, 0),
                                         (cases.ARGB8_U, 900, 700, cases.ARGB8_U, 30, 41, 1),
                                         (cases.RGB8, 40, 30, cases.RGBA8_P, 130, 95, 0),
                                         (cases.BGRA8_P, 640, 480, cases.RGB8, 80, 60, 1)]:
        src = cases.make_image(ti, wi, hi, None, "random", seed=11)
        whole = reference.scale_simple(src, ti, wi, hi, wi * cases.bpp(ti), to, wo, ho, None, srgb)
        so = wo * cases.bpp(to)
        y = 0
        while y < ho:
            n = int(min(ho - y, rng.integers(1, 17)))
            band = restatement.scale_rows(src, ti, wi, hi, wi * cases.bpp(ti), to, wo, ho, y, n, None, srgb)
            assert np.array_equal(band, whole[y * so: y * so + band.size]), (ti, to, y, n)
            y += n


def test_documented_quirks(restatement):
    """Appendix C of SURVEY.md: behaviours that look odd but are part of the contract."""
    src = np.array([40, 60, 80, 128] * 4, dtype=np.uint8)
    rgb = restatement.scale_simple(src, cases.RGBA8_P, 2, 2, 8, cases.RGB8, 1, 1, 3, 1)
    bgr = restatement.scale_simple(src, cases.RGBA8_P, 2, 2, 8, cases.BGR8, 1, 1, 3, 1)
    assert list(rgb) == [79, 119, 159]          # unpremultiplied, gamma-compressed
    assert list(bgr) == [115, 86, 56]           # gamma-compressed while still premultiplied
    white = np.full(3000 * 20 * 4, 255, np.uint8)
    out = restatement.scale_simple(white, cases.RGBA8_U, 3000, 20, 12000, cases.RGBA8_U, 301, 2, None, 1)
    assert list(out[:4]) == [0x61, 0x61, 0x61, 0xFF]   # 16-bit truncation of 19-bit lanes
    out = restatement.scale_simple(np.full(9000 * 4, 255, np.uint8), cases.RGBA8_P, 9000, 1, 36000,
                                   cases.RGBA8_P, 1000, 1, None, 0)
    assert out[-1] == 0xE3 and out[0] == 0xFF          # integer-ratio box drops the last pixel


def test_saturation(restatement):
    """verify.c:304-395: all-0xff in, all-0xff out, every type pair, H and V, the reference's sizes."""
    for ti, to in cases.all_type_pairs():
        for n_in, n_out in [(1, 65535), (2, 65535), (65534, 65535), (65535, 1), (65535, 65534)]:
            src = np.full(n_in * cases.bpp(ti), 0xFF, np.uint8)
            for srgb in (0, 1):
                h = restatement.scale_simple(src, ti, n_in, 1, n_in * cases.bpp(ti), to, n_out, 1, None, srgb)
                v = restatement.scale_simple(src, ti, 1, n_in, cases.bpp(ti), to, 1, n_out, cases.bpp(to), srgb)
                assert (h == 0xFF).all() and (v == 0xFF).all(), (ti, to, n_in, n_out, srgb)

#!/usr/bin/env python3
"""Generate tests/golden/digests.json from the UNMODIFIED compiled reference (oracle/_ref).

The reference ships no golden vectors (SURVEY.md sections 4, 8c), so we make our own: seeded
synthetic inputs (tests/cases.py) -> reference smol_scale_simple (generic C path) -> SHA-256 of
the output bytes.  Runs only where /root/reference has been compiled (`make -C oracle ref`); the
resulting JSON is committed so the digests travel to machines without the reference.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases   # noqa: E402
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "digests.json")


def golden_jobs():
    """name -> job tuple (type_in, w_in, h_in, stride_in, type_out, w_out, h_out, stride_out, srgb, mode, seed)"""
    jobs = {}
    # the five BASELINE.json configurations at full size (config 5: three sampled images)
    for name, ti, wi, hi, to, wo, ho, srgb, mode in cases.BASELINE_CONFIGS:
        seeds = (0, 1, 2) if name.startswith("cfg5") else (0,)
        for s in seeds:
            jobs["%s_seed%d" % (name, s)] = (ti, wi, hi, wi * cases.bpp(ti), to, wo, ho, wo * cases.bpp(to),
                                             srgb, mode, s)
    # a fixed random matrix over every filter class / type / stride
    for i, j in enumerate(cases.job_matrix(20261017, 300)):
        jobs["matrix_%03d" % i] = j + (1000 + i,)
    # all 100 type pairs, sRGB off/on, on one bilinear and one box geometry
    k = 0
    for ti, to in cases.all_type_pairs():
        for srgb in (0, 1):
            jobs["pairs_bilinear_%03d" % k] = (ti, 61, 37, 61 * cases.bpp(ti) + 5, to, 40, 21,
                                               40 * cases.bpp(to) + 3, srgb, "random", 5000 + k)
            jobs["pairs_box_%03d" % k] = (ti, 331, 97, 331 * cases.bpp(ti), to, 13, 9,
                                          13 * cases.bpp(to), srgb, "alpha_edges", 7000 + k)
            k += 1
    # exact 2^k:1 reductions (packed-byte kernel family)
    for i, j in enumerate(cases.half_jobs()):
        jobs["half_%03d" % i] = j + (3000 + i,)
    # table-edge cases on long axes
    for i, (a, b) in enumerate(cases.BIG_AXIS_PAIRS):
        for srgb in (0, 1):
            jobs["bigaxis_h_%02d_%d" % (i, srgb)] = (cases.ARGB8_U, a, 2, a * 4, cases.ARGB8_U, b, 2, b * 4,
                                                     srgb, "random", 9000 + i)
            jobs["bigaxis_v_%02d_%d" % (i, srgb)] = (cases.RGBA8_P, 1, a, 4, cases.BGR8, 1, b, 3,
                                                     srgb, "random", 9100 + i)
    return jobs


def digest_of(scaler, job):
    ti, wi, hi, si, to, wo, ho, so, srgb, mode, seed = job
    src = cases.make_image(ti, wi, hi, si, mode, seed)
    out = scaler.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
    return hashlib.sha256(out.tobytes()).hexdigest()


def main():
    ref = oracle.reference()
    if ref is None:
        sys.exit("oracle/_ref/libsmolref.so missing: run `make -C oracle ref` where /root/reference exists")
    jobs = golden_jobs()
    digests = {name: {"job": list(job), "sha256": digest_of(ref, job)} for name, job in jobs.items()}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump({"generator": "tools/gen_golden.py", "source": "reference smolscale-generic.c via oracle/_ref/libsmolref.so",
                   "fill_byte": "0xCD in row padding", "digests": digests}, f, indent=0, sort_keys=True)
    print("wrote %d digests to %s" % (len(digests), OUT))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Alignment / filter-mix "cliff table": BASELINE-shaped jobs whose buffers miss the fast kernels' alignment
preconditions (pitch + 4 / + 8 / + 12, base + 4, tightly packed odd widths) or whose axes mix box and
bilinear, timed beside their aligned twins.  Protocol of tools/time_job.py (device-resident frames, one
smol_scale_simple per frame, CUDA-graph replay); every result is checked against the oracle on a row window.
Output: one JSON object; `ratio` = us of the variant / us of the aligned twin."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle
import smolscale_b200 as sb

chk = oracle.restatement()


def bpp(t):
    return 3 if t >= 8 else 4


def time_job(ti, wi, hi, to, wo, ho, srgb, pitch_in_extra=0, pitch_out_extra=0, base_off=0, frames=4):
    bi, bo = bpp(ti), bpp(to)
    si, so = wi * bi + pitch_in_extra, wo * bo + pitch_out_extra
    n_in, n_out = si * hi + 64, so * ho + 64
    d_in = torch.randint(0, 256, (frames, n_in), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((frames, n_out), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        sb.set_stream(stream.cuda_stream)

        def step():
            for f in range(frames):
                sb.scale_simple(d_in[f].data_ptr() + base_off, ti, wi, hi, si, d_out[f].data_ptr() + base_off, to, wo, ho, so, srgb)
        sb.reset_stats()
        step()
        stream.synchronize()
        fam = [k for k, v in sb.kernel_launches().items() if v]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            step()
        for _ in range(3):
            g.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(stream)
        for _ in range(reps):
            g.replay()
        e1.record(stream)
        stream.synchronize()
        sb.set_stream(None)
    us = e0.elapsed_time(e1) * 1e3 / (reps * frames)
    # oracle check of a window of rows of frame 0
    src = d_in[0].cpu().numpy()[base_off:base_off + si * hi]
    got = d_out[0].cpu().numpy()[base_off:]
    y0, n = ho // 2, min(3, ho - ho // 2)
    want = chk.scale_rows(src, ti, wi, hi, si, to, wo, ho, y0, n, so, srgb)
    ok = True
    for r in range(n):
        ok &= bool(np.array_equal(want[r * so:r * so + wo * bo], got[(y0 + r) * so:(y0 + r) * so + wo * bo]))
    return {"us": round(us, 2), "kernel": "+".join(fam), "ok": ok}


JOBS = [
    # name, type_in, w_in, h_in, type_out, w_out, h_out, srgb
    ("cfg1 1080p->540p RGBA_P", 0, 1920, 1080, 0, 960, 540, 0),
    ("cfg2 4K->1080p BGRA_P->BGRA_U", 1, 3840, 2160, 5, 1920, 1080, 0),
    ("cfg3 8K->800x450 box sRGB", 0, 7680, 4320, 0, 800, 450, 1),
    ("cfg4 RGB 1024x768->4096x3072", 8, 1024, 768, 8, 4096, 3072, 0),
    ("cfg5 2048^2->256^2 ARGB_P", 2, 2048, 2048, 2, 256, 256, 0),
    ("8K RGBA 7681 wide -> 3840 (tight, odd)", 0, 7681, 2160, 0, 3840, 1080, 0),
    ("8K RGB8 7681 wide tight -> 800x450 box", 8, 7681, 4320, 8, 800, 450, 0),
]
VARIANTS = [("aligned", 0, 0, 0), ("pitch+4", 4, 4, 0), ("pitch+8", 8, 8, 0), ("pitch+12", 12, 12, 0), ("base+4", 0, 0, 4),
            ("pitch+1", 1, 1, 0)]
MIXED = [
    ("8Kx1080 -> 800x540 (H box, V bilinear)", 0, 7680, 1080, 0, 800, 540, 0),
    ("8Kx1080 -> 800x540 sRGB (H box, V bilinear)", 0, 7680, 1080, 0, 800, 540, 1),
    ("1920x8640 -> 960x450 (H bilinear, V box)", 0, 1920, 8640, 0, 960, 450, 0),
    ("4K -> 1920x100 (H bilinear, V box)", 1, 3840, 2160, 1, 1920, 100, 0),
    ("7680x4320 -> 20x12 (> 255:1, no sRGB)", 0, 7680, 4320, 0, 20, 12, 0),
]

res = {"what": "us per frame, device-resident, graph replay", "jobs": []}
for name, ti, wi, hi, to, wo, ho, srgb in JOBS:
    row = {"job": name, "variants": {}}
    base = None
    for vname, pi, po, off in VARIANTS:
        r = time_job(ti, wi, hi, to, wo, ho, srgb, pi, po, off)
        if vname == "aligned":
            base = r["us"]
        r["ratio"] = round(r["us"] / base, 2)
        row["variants"][vname] = r
    res["jobs"].append(row)
    print(name, {k: (v["us"], v["kernel"], v["ok"]) for k, v in row["variants"].items()}, file=sys.stderr, flush=True)
for name, ti, wi, hi, to, wo, ho, srgb in MIXED:
    r = time_job(ti, wi, hi, to, wo, ho, srgb)
    alg = wi * hi * bpp(ti) + wo * ho * bpp(to)
    r["algorithmic_gbs"] = round(alg / r["us"] / 1e3, 0)
    res["jobs"].append({"job": name, "variants": {"aligned": r}})
    print(name, r, file=sys.stderr, flush=True)
print(json.dumps(res))

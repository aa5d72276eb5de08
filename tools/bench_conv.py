#!/usr/bin/env python3
"""The reference's `benchmark-conv` matrix (test.c:1055-1107) on the GPU path: all 10 x 10 pixel
type pairs at 3840x2160 -> 3839x2159 (bilinear, non-trivial weights), device-resident buffers.
Prints one line per pair: kernel family, us per frame, algorithmic GB/s, output Mpix/s, and
whether a sampled row window matches the oracle.  Usage: python tools/bench_conv.py [--srgb] [--geom WxH:WxH]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import cases, oracle
import smolscale_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--srgb", type=int, default=0)
ap.add_argument("--geom", default="3840x2160:3839x2159")
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--json", default="")
args = ap.parse_args()
(wi, hi), (wo, ho) = [tuple(int(v) for v in g.split("x")) for g in args.geom.split(":")]
chk = oracle.restatement()
stream = torch.cuda.Stream()
rows = []
with torch.cuda.stream(stream):
    sb.set_stream(stream.cuda_stream)
    for ti in range(10):
        bi = cases.bpp(ti)
        g = torch.Generator(device="cuda"); g.manual_seed(ti)
        d_in = torch.randint(0, 256, (args.frames, hi * wi * bi), dtype=torch.uint8, device="cuda", generator=g)
        for to in range(10):
            bo = cases.bpp(to)
            d_out = torch.zeros((args.frames, ho * wo * bo), dtype=torch.uint8, device="cuda")
            def step():
                for f in range(args.frames):
                    sb.scale_simple(d_in[f].data_ptr(), ti, wi, hi, wi * bi, d_out[f].data_ptr(), to, wo, ho, wo * bo, args.srgb)
            step(); stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                step()
            e1.record(stream); stream.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (args.reps * args.frames)
            alg = hi * wi * bi + ho * wo * bo
            src = d_in[0].cpu().numpy(); got = d_out[0].cpu().numpy()
            y0 = ho // 2
            want = chk.scale_rows(src, ti, wi, hi, wi * bi, to, wo, ho, y0, 3, wo * bo, args.srgb)
            ok = bool(np.array_equal(want, got[y0 * wo * bo: y0 * wo * bo + want.size]))
            k = sb.plan_query(ti, wi, hi, to, wo, ho, args.srgb)["kernel_name"]
            rows.append({"in": cases.TYPE_NAMES[ti], "out": cases.TYPE_NAMES[to], "kernel": k, "us": round(us, 2),
                         "gbs": round(alg / us / 1e3, 1), "mpix": round(wo * ho / us, 1), "ok": ok})
            print("%-8s -> %-8s %-12s %8.2f us %8.1f GB/s %10.1f Mpix/s %s" %
                  (cases.TYPE_NAMES[ti], cases.TYPE_NAMES[to], k, us, alg / us / 1e3, wo * ho / us, "ok" if ok else "MISMATCH"), flush=True)
if args.json:
    json.dump({"geom": args.geom, "srgb": args.srgb, "rows": rows}, open(args.json, "w"), indent=0)
print("all ok" if all(r["ok"] for r in rows) else "MISMATCHES PRESENT")

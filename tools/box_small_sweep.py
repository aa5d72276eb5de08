#!/usr/bin/env python3
"""Small and medium box-filter jobs (thumbnail shapes) timed under one table placement of the box kernel
(SMOL_BOX_LUTM=0|2|3 in the environment, read once per process): finds where the one-CTA-per-SM kernel with its
64 KB lane-replicated tables stops paying for itself.  Protocol of tools/time_job.py.  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, smolscale_b200 as sb

SHAPES = [(400, 400, 16, 16), (256, 256, 32, 32), (640, 480, 64, 48), (1024, 768, 128, 96), (1024, 1024, 64, 64),
          (1920, 1080, 240, 135), (2048, 2048, 128, 128), (2048, 2048, 200, 200), (4000, 3000, 200, 150),
          (4000, 3000, 400, 300), (7680, 4320, 800, 450)]
rows = []
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sb.set_stream(stream.cuda_stream)
    for srgb in (1, 0):
        for ti in (0, 8):                    # RGBA8 premultiplied, RGB8
            for wi, hi, wo, ho in SHAPES:
                bi = 3 if ti >= 8 else 4
                frames = 8 if wi * hi < 8e6 else 4
                si, so = (wi * bi + 15) // 16 * 16, (wo * bi + 15) // 16 * 16
                d_in = torch.randint(0, 256, (frames, hi * si), dtype=torch.uint8, device="cuda")
                d_out = torch.zeros((frames, ho * so), dtype=torch.uint8, device="cuda")

                def step():
                    for f in range(frames):
                        sb.scale_simple(d_in[f].data_ptr(), ti, wi, hi, si, d_out[f].data_ptr(), ti, wo, ho, so, srgb)
                sb.reset_stats(); step(); stream.synchronize()
                fam = [k for k, v in sb.kernel_launches().items() if v]
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    step()
                for _ in range(3):
                    g.replay()
                stream.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 20
                e0.record(stream)
                for _ in range(reps):
                    g.replay()
                e1.record(stream); stream.synchronize()
                rows.append({"job": "%dx%d->%dx%d t%d srgb%d" % (wi, hi, wo, ho, ti, srgb), "kernel": fam,
                             "us": round(e0.elapsed_time(e1) * 1e3 / (reps * frames), 2)})
print(json.dumps({"lutm": os.environ.get("SMOL_BOX_LUTM", "default"), "min_warps": os.environ.get("SMOL_BOX_MIN_WARPS", "default"),
                  "rows": rows}))

#!/usr/bin/env python3
"""128bpp bilinear-with-halvings jobs (linear light, unassociated -> unassociated; 2:1 < ratio <= 8:1 on an axis) from
thumbnail sizes up to 4K, timed with the protocol of tools/time_job.py, with a digest of every result so that two builds
of the library (SMOLSCALE_B200_LIB) can be compared bit for bit.  One JSON line."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, smolscale_b200 as sb

SHAPES = [(256, 256, 32, 32), (640, 480, 160, 120), (1024, 768, 128, 96), (1024, 768, 400, 300), (1920, 1080, 240, 135),
          (1920, 1080, 640, 360), (2048, 2048, 256, 256), (3840, 2160, 1280, 720), (3840, 2160, 640, 360),
          (4000, 3000, 1000, 750), (3840, 2160, 1279, 2160), (3840, 2160, 3840, 719)]
TYPES = [(0, 0, 1), (4, 4, 1), (8, 8, 1), (4, 5, 0), (1, 8, 1)]          # (in, out, srgb)
rows = []
stream = torch.cuda.Stream()
torch.manual_seed(5)
with torch.cuda.stream(stream):
    sb.set_stream(stream.cuda_stream)
    for ti, to, srgb in TYPES:
        for wi, hi, wo, ho in SHAPES:
            bi, bo = (3 if ti >= 8 else 4), (3 if to >= 8 else 4)
            frames = 8 if wi * hi < 4e6 else 4
            si, so = wi * bi, wo * bo
            d_in = torch.randint(0, 256, (frames, hi * si), dtype=torch.uint8, device="cuda")
            d_out = torch.zeros((frames, ho * so), dtype=torch.uint8, device="cuda")

            def step():
                for f in range(frames):
                    sb.scale_simple(d_in[f].data_ptr(), ti, wi, hi, si, d_out[f].data_ptr(), to, wo, ho, so, srgb)
            sb.reset_stats(); step(); stream.synchronize()
            fam = [k for k, v in sb.kernel_launches().items() if v]
            digest = hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest()[:16]
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
            rows.append({"job": "%dx%d->%dx%d t%d->t%d srgb%d" % (wi, hi, wo, ho, ti, to, srgb), "kernel": fam,
                         "us": round(e0.elapsed_time(e1) * 1e3 / (reps * frames), 2), "sha1": digest})
print(json.dumps({"lib": os.environ.get("SMOLSCALE_B200_LIB", "default"), "rows": rows}))

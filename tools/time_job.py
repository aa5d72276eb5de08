#!/usr/bin/env python3
"""Time one job shape on the GPU: N distinct frames per step through smol_scale_simple with
device-resident buffers, replayed as a CUDA graph (the bench.py protocol for an arbitrary job).
Usage: time_job.py WI HI WO HO TYPE_IN TYPE_OUT SRGB [FRAMES] [--align] [--opaque]  (pitches padded to 16 bytes with --align;
SMOL_FORCE_KERNEL=<id> forces a kernel family)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, smolscale_b200 as sb
a = [x for x in sys.argv[1:] if not x.startswith("--")]
align = "--align" in sys.argv
wi, hi, wo, ho, ti, to, srgb = [int(v) for v in a[:7]]
frames = int(a[7]) if len(a) > 7 else 8
bi, bo = (3 if ti >= 8 else 4), (3 if to >= 8 else 4)
si, so = wi * bi, wo * bo
if align:
    si, so = (si + 15) // 16 * 16, (so + 15) // 16 * 16
if os.environ.get("SMOL_FORCE_KERNEL"):
    sb.force_kernel(int(os.environ["SMOL_FORCE_KERNEL"]))
d_in = torch.randint(0, 256, (frames, hi * si), dtype=torch.uint8, device="cuda")
if "--opaque" in sys.argv and bi == 4:
    # every alpha byte 255 (a photograph in an RGBA container); needs pitch == width * 4
    d_in.view(frames, -1, 4)[:, :, 3 if (ti & 3) < 2 else 0] = 255
d_out = torch.zeros((frames, ho * so), dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sb.set_stream(stream.cuda_stream)
    band = [a for a in sys.argv if a.startswith("--rows=")]     # --rows=FIRST:COUNT : one output row band per frame (smol_scale_batch_full)
    def step():
        for f in range(frames):
            if band:
                first, count = [int(v) for v in band[0][7:].split(":")]
                ctx = sb.ScaleCtx(d_in[f].data_ptr(), ti, wi, hi, si, None, to, wo, ho, so, srgb)
                ctx.batch_full(d_out[f].data_ptr(), first, count)
                ctx.destroy()
            else:
                sb.scale_simple(d_in[f].data_ptr(), ti, wi, hi, si, d_out[f].data_ptr(), to, wo, ho, so, srgb)
    sb.reset_stats(); step(); stream.synchronize()
    fam = {k: v for k, v in sb.kernel_launches().items() if v}
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
us = e0.elapsed_time(e1) * 1e3 / (reps * frames)
alg = hi * wi * bi + ho * wo * bo
print("%dx%d t%d -> %dx%d t%d srgb %d: %s  %.2f us/frame  %.0f GB/s algorithmic  %.0f Mpix/s" %
      (wi, hi, ti, wo, ho, to, srgb, fam, us, alg / us / 1e3, wo * ho / us))

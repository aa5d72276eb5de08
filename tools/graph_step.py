#!/usr/bin/env python3
"""One bench step of a configuration as a single CUDA graph (FRAMES launches, one smol_scale_simple per
frame), replayed a few times -- the unit `ncu --graph-profiling graph` measures as ONE entity: total
duration and DRAM bytes of the whole step with the launches overlapping as they do in bench.py.
Usage: graph_step.py [cfg] [replays]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import smolscale_b200 as sb
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
replays = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ti, wi, hi, to, wo, ho, srgb, premul, frames, desc = bench.CONFIGS[cfg_name]
si, so = wi * bench.bpp(ti), wo * bench.bpp(to)
d_in = torch.randint(0, 256, (frames, si * hi), dtype=torch.uint8, device="cuda")
d_out = torch.zeros((frames, so * ho), dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sb.set_stream(stream.cuda_stream)
    def step():
        if cfg_name == "cfg5":
            sb.scale_images(d_in.data_ptr(), si * hi, ti, wi, hi, si, d_out.data_ptr(), so * ho, to, wo, ho, so, srgb, frames)
        else:
            for f in range(frames):
                sb.scale_simple(d_in[f].data_ptr(), ti, wi, hi, si, d_out[f].data_ptr(), to, wo, ho, so, srgb)
    step(); stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        step()
    for _ in range(2):
        g.replay()
    stream.synchronize()
    torch.cuda.profiler.start()         # ncu --profile-from-start off: only the replays below are profiled
    for _ in range(replays):
        g.replay()
    stream.synchronize()
    torch.cuda.profiler.stop()
print("replayed", replays, "graphs of", 1 if cfg_name == "cfg5" else frames, "launches")

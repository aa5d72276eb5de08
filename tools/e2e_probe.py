#!/usr/bin/env python3
"""End-to-end (host buffers) throughput of one configuration through the C API, by caller-memory kind and
caller thread count.  Library tunables come from the environment (SMOL_CUDA_BOUNCE, SMOL_CUDA_BOUNCE_BAND_KB,
SMOL_CUDA_BOUNCE_TASK_KB, SMOL_CUDA_HOST_THREADS, SMOL_CUDA_MULTI_GPU), so one run = one setting."""
import concurrent.futures, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import smolscale_b200 as sb
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = bench.CONFIGS[cfg_name]
ti, wi, hi, to, wo, ho, srgb = cfg[:7]
si, so = wi * bench.bpp(ti), wo * bench.bpp(to)
in_bytes, out_bytes = si * hi, so * ho
frames = 8
src = bench.synth_frames_host(cfg, frames, 1)
pin_in = torch.from_numpy(src.copy()).pin_memory(); pin_out = torch.zeros(frames * out_bytes, dtype=torch.uint8).pin_memory()
pg_in = src.reshape(-1).copy(); pg_out = np.zeros(frames * out_bytes, np.uint8)
res = {"config": cfg_name, "env": {k: v for k, v in os.environ.items() if k.startswith("SMOL_")}, "runs": []}
for kind, a, b in (("pinned", pin_in.data_ptr(), pin_out.data_ptr()), ("pageable", pg_in.ctypes.data, pg_out.ctypes.data)):
    for threads in [int(t) for t in os.environ.get("E2E_THREADS", "1,3").split(",")]:
        pool = concurrent.futures.ThreadPoolExecutor(max_workers=threads)
        def one(f):
            sb.scale_simple(a + f * in_bytes, ti, wi, hi, si, b + f * out_bytes, to, wo, ho, so, srgb)
        for _ in range(2):
            list(pool.map(one, range(frames)))
        t0 = time.perf_counter(); k = 6
        for _ in range(k):
            list(pool.map(one, range(frames)))
        dt = time.perf_counter() - t0
        res["runs"].append({"memory": kind, "caller_threads": threads, "mpix_s": round(k * frames * wo * ho / 1e6 / dt, 1),
                            "host_gbs": round(k * frames * (in_bytes + out_bytes) / dt / 1e9, 1)})
        pool.shutdown()
assert np.array_equal(pg_out, pin_out.numpy())
print(json.dumps(res))

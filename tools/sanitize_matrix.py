"""Small job matrix for compute-sanitizer (memcheck / racecheck): every kernel family, host and
device pointers, aligned and misaligned."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases, oracle
import smolscale_b200 as sb
chk = oracle.restatement()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
bad = 0
jobs = cases.job_matrix(31337, n) + cases.half_jobs()[:40]
# aligned-pitch magnifications (byte-granular tile kernel) and box jobs with long / ragged windows
for (wi, hi, wo, ho) in [(64, 48, 256, 192), (100, 7, 349, 65), (33, 30, 130, 61), (5, 4, 40, 30), (300, 40, 1500, 97)]:
    for ti, to in [(cases.RGB8, cases.BGR8), (cases.BGR8, cases.RGB8), (cases.RGBA8_P, cases.RGB8)]:
        jobs.append((ti, wi, hi, wi * cases.bpp(ti), to, wo, ho, (wo * cases.bpp(to) + 15) // 16 * 16, 0, "random"))
for (wi, hi, wo, ho) in [(200, 90, 15, 9), (255, 100, 16, 7), (1500, 120, 100, 11), (1021, 50, 64, 5), (4000, 40, 15, 3)]:
    for ti, srgb in [(cases.RGBA8_P, 1), (cases.ARGB8_U, 1), (cases.RGB8, 0), (cases.BGRA8_U, 0), (cases.RGB8, 1)]:
        jobs.append((ti, wi, hi, (wi * cases.bpp(ti) + 15) // 16 * 16, cases.RGBA8_P, wo, ho, wo * 4, srgb, "random"))
# round 2 kernels: 2:1 on 32- / 16- / 8- / 4-byte-aligned rows (256-bit loads and the narrow variants), box rows off
# 16-byte boundaries (row-shifted staging, 24bpp tight pitches), the 128bpp strip kernel, 24bpp stores at any alignment
for (wi, hi, wo, ho) in [(256, 64, 128, 32), (250, 62, 125, 31), (64, 64, 8, 8), (72, 40, 18, 10)]:
    for extra in (0, 4, 8, 12, 16):
        jobs.append((cases.BGRA8_P, wi, hi, wi * 4 + extra, cases.BGRA8_U, wo, ho, wo * 4 + extra, 0, "premul"))
for (wi, hi, wo, ho) in [(1500, 120, 100, 11), (1021, 50, 64, 5), (777, 95, 33, 7)]:
    for ti, extra, srgb in [(cases.RGBA8_P, 4, 1), (cases.ARGB8_U, 8, 1), (cases.RGB8, 1, 0), (cases.RGB8, 3, 1), (cases.BGRA8_U, 12, 0)]:
        jobs.append((ti, wi, hi, wi * cases.bpp(ti) + extra, cases.RGBA8_P, wo, ho, wo * 4, srgb, "random"))
for (wi, hi, wo, ho) in [(129, 33, 131, 35), (64, 48, 127, 96), (200, 100, 133, 67), (37, 21, 41, 37)]:
    for ti, to, srgb in [(cases.RGBA8_U, cases.ARGB8_U, 0), (cases.BGRA8_U, cases.BGRA8_U, 1), (cases.RGBA8_P, cases.RGB8, 1),
                         (cases.RGB8, cases.BGR8, 1), (cases.RGBA8_P, cases.RGB8, 0), (cases.RGB8, cases.RGB8, 0)]:
        jobs.append((ti, wi, hi, wi * cases.bpp(ti), to, wo, ho, wo * cases.bpp(to) + (1 if cases.bpp(to) == 3 else 0), srgb, "random"))
# 128bpp bilinear with halvings (taps128's two instances, strips, the tile kernel on tiny jobs; SMOL_TILE128H=1 in the
# environment sends all of them through the tile kernel): tight and padded pitches, 24bpp row ends
t128 = []
for (wi, hi, wo, ho) in [(256, 256, 32, 32), (100, 100, 33, 33), (37, 29, 13, 11), (9, 9, 3, 4), (5, 5, 2, 2), (255, 7, 100, 3),
                         (7, 300, 3, 101), (300, 40, 300, 11), (40, 300, 11, 300), (640, 480, 160, 120), (700, 500, 233, 499)]:
    for ti, to, srgb, extra in [(cases.RGBA8_P, cases.BGRA8_U, 1, 0), (cases.ARGB8_U, cases.ARGB8_P, 1, 4), (cases.RGB8, cases.BGR8, 1, 0),
                                (cases.RGB8, cases.RGBA8_P, 1, 1), (cases.BGRA8_U, cases.RGBA8_U, 0, 8), (cases.ABGR8_P, cases.RGB8, 1, 0)]:
        t128.append((ti, wi, hi, wi * cases.bpp(ti) + extra, to, wo, ho, wo * cases.bpp(to) + (extra if cases.bpp(to) == 4 else 0), srgb,
                     "premul" if ti < 4 else "random"))
jobs += t128
if len(sys.argv) > 2 and sys.argv[2] == "taps128":
    jobs = t128
for idx, job in enumerate(jobs):
    ti, wi, hi, si, to, wo, ho, so, srgb, mode = job
    src = cases.make_image(ti, wi, hi, si, mode, seed=idx)
    want = chk.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
    # device pointers placed at the very end of their allocations: any over-read/-write is out of bounds
    # (run with PYTORCH_NO_CUDA_MEMORY_CACHING=1 so that every tensor is its own cudaMalloc); the last
    # source row carries no pitch padding
    n_in = si * (hi - 1) + wi * cases.bpp(ti)
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_in.copy_(torch.from_numpy(src[:n_in]))
    d_out = torch.full((want.size,), 0xCD, dtype=torch.uint8, device="cuda")
    sb.scale_simple(d_in, ti, wi, hi, si, d_out, to, wo, ho, so, srgb)
    torch.cuda.synchronize()
    if not np.array_equal(d_out.cpu().numpy(), want):
        bad += 1; print("MISMATCH", job)
    got = np.full_like(want, 0xCD)
    sb.scale_simple(src, ti, wi, hi, si, got, to, wo, ho, so, srgb)
    if not np.array_equal(got, want):
        bad += 1; print("MISMATCH host", job)
# row batches of height-preserving jobs on host buffers (the staged band must hold the row below its last one)
import threading
for ti, wi, hi, to, wo, ho in [(cases.BGRA8_P, 640, 360, cases.BGRA8_P, 320, 360), (cases.RGBA8_U, 200, 120, cases.ABGR8_P, 200, 120)]:
    si, so = wi * cases.bpp(ti), wo * cases.bpp(to)
    src = cases.make_image(ti, wi, hi, si, "random", seed=7)
    want = chk.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, 0)
    out = np.zeros_like(want)
    ctx = sb.ScaleCtx(src, ti, wi, hi, si, out, to, wo, ho, so, 0)
    for y in range(0, ho, 50):
        ctx.batch(y, min(50, ho - y))
    ctx.destroy()
    if not np.array_equal(out, want):
        bad += 1; print("MISMATCH bands", (ti, wi, hi, to, wo, ho))
print("jobs", len(jobs), "bad", bad)

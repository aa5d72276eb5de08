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
for idx, job in enumerate(jobs):
    ti, wi, hi, si, to, wo, ho, so, srgb, mode = job
    src = cases.make_image(ti, wi, hi, si, mode, seed=idx)
    want = chk.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
    # device pointers placed at the very end of their allocations: any over-read/-write is out of bounds
    d_in = torch.empty(src.size, dtype=torch.uint8, device="cuda"); d_in.copy_(torch.from_numpy(src))
    d_out = torch.full((want.size,), 0xCD, dtype=torch.uint8, device="cuda")
    sb.scale_simple(d_in, ti, wi, hi, si, d_out, to, wo, ho, so, srgb)
    torch.cuda.synchronize()
    if not np.array_equal(d_out.cpu().numpy(), want):
        bad += 1; print("MISMATCH", job)
    got = np.full_like(want, 0xCD)
    sb.scale_simple(src, ti, wi, hi, si, got, to, wo, ho, so, srgb)
    if not np.array_equal(got, want):
        bad += 1; print("MISMATCH host", job)
print("jobs", len(jobs), "bad", bad)

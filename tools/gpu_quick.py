"""Ad-hoc GPU shake-out: random job matrix vs the oracle with verbose diagnostics."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, oracle
import smolscale_b200 as sb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
chk = oracle.restatement()
ref = oracle.reference()
print("devices", sb.device_count(), "ref", ref is not None, flush=True)
fails = 0
t0 = time.time()
for idx, job in enumerate(cases.job_matrix(1234, n)):
    ti, wi, hi, si, to, wo, ho, so, srgb, mode = job
    src = cases.make_image(ti, wi, hi, si, mode, seed=idx)
    want = chk.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
    if ref is not None:
        w2 = ref.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
        assert np.array_equal(want, w2), ("oracle != ref", job)
    got = np.full_like(want, 0xCD)
    sb.scale_simple(src, ti, wi, hi, si, got, to, wo, ho, so, srgb)
    if not np.array_equal(got, want):
        fails += 1
        bad = np.nonzero(got != want)[0]
        p = sb.plan_query(ti, wi, hi, to, wo, ho, srgb)
        if fails <= 25:
            print("MISMATCH", job, "nbad", bad.size, "first", bad[:6], "got", got[bad[:6]], "want", want[bad[:6]],
                  {k: p[k] for k in ("filter_h", "filter_v", "halvings_h", "halvings_v", "storage_bits", "mid")}, flush=True)
print("jobs", n, "fails", fails, "time %.1fs" % (time.time() - t0), sb.stats(), flush=True)
sys.exit(1 if fails else 0)

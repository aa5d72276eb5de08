"""Randomised soak test on the GPU: many more jobs/seeds than the unit tests, all kernel paths,
host + device + misaligned pointers, random row bands.  Exit code 1 on any mismatch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases, oracle
import smolscale_b200 as sb
chk = oracle.restatement()
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed0)
pairs = cases.AXIS_PAIRS + cases.HALF_AXIS_PAIRS + [(640, 200), (641, 97), (333, 777), (1000, 3), (4000, 15), (12, 700), (255, 1), (256, 1), (257, 1), (2040, 8), (2041, 8), (96, 12), (100, 50), (77, 154),
                                                       (200, 15), (255, 16), (1500, 100), (3000, 230), (1300, 87), (160, 640), (33, 1000)]
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < seconds:
    wi, wo = pairs[int(rng.integers(len(pairs)))]
    hi, ho = pairs[int(rng.integers(len(pairs)))]
    if wi * hi > 1500000 or wo * ho > 1500000:
        continue
    ti, to = int(rng.integers(10)), int(rng.integers(10))
    srgb = int(rng.integers(2))
    mode = cases.IMAGE_MODES[int(rng.integers(len(cases.IMAGE_MODES)))]
    align = int(rng.choice([0, 16]))          # aligned pitches half of the time (fast paths)
    si = wi * cases.bpp(ti); so = wo * cases.bpp(to)
    if align:
        si = (si + 15) & ~15; so = (so + 15) & ~15
    else:
        si += int(rng.choice([0, 1, 3, 4])); so += int(rng.choice([0, 1, 3, 4]))
    src = cases.make_image(ti, wi, hi, si, mode, seed=int(rng.integers(1 << 30)))
    want = chk.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
    variant = int(rng.integers(4))
    if os.environ.get("SOAK_TRACE"):
        with open(os.environ["SOAK_TRACE"], "w") as f:
            f.write(repr((ti, wi, hi, si, to, wo, ho, so, srgb, mode, variant)) + " " +
                    sb.plan_query(ti, wi, hi, to, wo, ho, srgb)["kernel_name"] + "\n")
    if variant == 0:      # host pointers
        got = np.full_like(want, 0xCD)
        sb.scale_simple(src, ti, wi, hi, si, got, to, wo, ho, so, srgb)
    elif variant == 1:    # device pointers, aligned base
        d_in = torch.from_numpy(src).cuda(); d_out = torch.full((want.size,), 0xCD, dtype=torch.uint8, device="cuda")
        sb.scale_simple(d_in, ti, wi, hi, si, d_out, to, wo, ho, so, srgb); torch.cuda.synchronize()
        got = d_out.cpu().numpy()
    elif variant == 2:    # device pointers, misaligned base
        o1, o2 = int(rng.integers(1, 16)), int(rng.integers(1, 16))
        d_in = torch.zeros(src.size + 16, dtype=torch.uint8, device="cuda"); d_in[o1:o1 + src.size] = torch.from_numpy(src).cuda()
        d_out = torch.full((want.size + 16,), 0xCD, dtype=torch.uint8, device="cuda")
        sb.scale_simple(d_in.data_ptr() + o1, ti, wi, hi, si, d_out.data_ptr() + o2, to, wo, ho, so, srgb); torch.cuda.synchronize()
        got = d_out.cpu().numpy()[o2:o2 + want.size]
    else:                 # random row bands through the batch API, device memory
        d_in = torch.from_numpy(src).cuda(); d_out = torch.full((want.size,), 0xCD, dtype=torch.uint8, device="cuda")
        ctx = sb.ScaleCtx(d_in, ti, wi, hi, si, d_out, to, wo, ho, so, srgb)
        y = 0
        while y < ho:
            k = int(min(ho - y, rng.integers(1, 40))); ctx.batch(y, k); y += k
        torch.cuda.synchronize(); ctx.destroy()
        got = d_out.cpu().numpy()
    n += 1
    if not np.array_equal(got, want):
        bad += 1
        d = np.nonzero(got != want)[0]
        print("MISMATCH", (ti, wi, hi, si, to, wo, ho, so, srgb, mode), "variant", variant, "nbad", d.size, d[:5], got[d[:5]], want[d[:5]],
              sb.plan_query(ti, wi, hi, to, wo, ho, srgb)["kernel_name"], flush=True)
        if bad > 20:
            break
print("soak: jobs", n, "bad", bad, "seconds %.0f" % (time.time() - t0))
sys.exit(1 if bad else 0)

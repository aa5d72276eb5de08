"""GPU test of tools/smol_generate: PNG in -> a range of scaled PNGs out, the reference's
`test smol generate` mode (test.c:1303-1371) driven through libsmolscale_cuda.so and libsmolpng.so.
Every file the program writes is decoded (Pillow) and compared bit-for-bit with the oracle scaling
the decoded input; file names and the size stepping are restated here from test.c:1330-1357."""
import os
import subprocess

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Image = pytest.importorskip("PIL.Image")


def _sizes(w, h, scale_min, scale_max, n_steps):
    clamp = lambda v: min(max(v, 1), 65535)
    w0, w1 = int(clamp(w * scale_min)), int(clamp(w * scale_max))
    h0, h1 = int(clamp(h * scale_min)), int(clamp(h * scale_max))
    if n_steps > 1:
        ws = np.float32(w1 - w0) / (np.float32(n_steps) - np.float32(1.0))
        hs = np.float32(h1 - h0) / (np.float32(n_steps) - np.float32(1.0))
    else:
        ws = hs = np.float32(99999.0)
    return [(int(clamp(np.float32(w0) + np.float32(s) * ws)), int(clamp(np.float32(h0) + np.float32(s) * hs)))
            for s in range(n_steps)]


@pytest.mark.parametrize("flags,ptype,srgb", [((), cases.RGBA8_U, 0), (("--srgb",), cases.RGBA8_U, 1),
                                              (("--reference-types",), cases.ARGB8_P, 0),
                                              (("--host", "--srgb"), cases.RGBA8_U, 1)])
def test_generate_matches_oracle(sb, restatement, tmp_path, flags, ptype, srgb):
    exe = os.path.join(ROOT, "tools", "smol_generate")
    assert os.path.exists(exe), "tools/smol_generate has not been built (python __graft_entry__.py)"
    w, h = 640, 360
    # premultiplied-valid bytes when the program reads them as ARGB premultiplied (alpha first)
    src = cases.make_image(ptype, w, h, None, "premul" if ptype == cases.ARGB8_P else "random", seed=3)
    path = str(tmp_path / "img.png")
    Image.fromarray(src.reshape(h, w, 4), "RGBA").save(path)
    scale_min, scale_max, n_steps = 0.11, 2.3, 6
    r = subprocess.run([exe, *flags, str(scale_min), str(scale_max), str(n_steps), path], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stderr.count("*") == n_steps
    for wo, ho in _sizes(w, h, scale_min, scale_max, n_steps):
        name = str(tmp_path / ("img-%04d-%04d.png" % (wo, ho)))
        assert os.path.exists(name), (name, os.listdir(tmp_path))
        got = np.asarray(Image.open(name))
        assert got.shape == (ho, wo, 4)
        want = restatement.scale_simple(src, ptype, w, h, w * 4, ptype, wo, ho, None, srgb)
        assert np.array_equal(got.reshape(-1), want)

"""GPU tests of the host layer around the kernels (smolscale-cuda.c): stream ordering of calls that
mix device and host buffers, pageable / pinned / managed caller memory, the banded staging pipeline,
row batches that stop short of the image's last row, one host-memory call spread over several GPUs,
first use of a geometry inside a CUDA graph capture, and more live contexts than the table cache has
slots.  Everything goes through the C-ABI and is compared bit-for-bit with the oracle."""
import ctypes
import threading

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _want(restatement, src, ti, wi, hi, to, wo, ho, srgb=0, si=None, so=None):
    return restatement.scale_simple(src, ti, wi, hi, si or wi * cases.bpp(ti), to, wo, ho, so, srgb)


def test_chained_calls_without_synchronize(sb, restatement):
    """device -> device (returns at once) followed by device -> host on the same stream, and a torch
    kernel producing the input of a device -> host call: no synchronize in between (the mixed call
    must order itself after the caller's stream)."""
    import torch
    ti, wi, hi, tm, wm, hm, to, wo, ho = cases.BGRA8_P, 3840, 2160, cases.BGRA8_P, 1920, 1080, cases.RGBA8_U, 700, 400
    src = cases.make_image(ti, wi, hi, None, "premul", seed=11)
    mid = _want(restatement, src, ti, wi, hi, tm, wm, hm)
    want = _want(restatement, mid, tm, wm, hm, to, wo, ho)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        sb.set_stream(stream.cuda_stream)
        try:
            for rep in range(6):
                d_in = torch.from_numpy(src).cuda(non_blocking=True)
                d_mid = torch.zeros(mid.size, dtype=torch.uint8, device="cuda")
                # enough queued work that the first scale has certainly not run when the second call arrives
                junk = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
                for _ in range(4):
                    junk.add_(1)
                sb.scale_simple(d_in, ti, wi, hi, wi * 4, d_mid, tm, wm, hm, wm * 4, 0)
                got = np.zeros_like(want)
                sb.scale_simple(d_mid, tm, wm, hm, wm * 4, got, to, wo, ho, wo * 4, 0)
                assert np.array_equal(got, want), rep
                # host -> device must not overtake earlier work on the stream that still reads the destination
                d_out = torch.zeros(mid.size, dtype=torch.uint8, device="cuda")
                for _ in range(4):
                    junk.add_(1)
                d_out.fill_(7)
                sb.scale_simple(src, ti, wi, hi, wi * 4, d_out, tm, wm, hm, wm * 4, 0)
                stream.synchronize()
                assert np.array_equal(d_out.cpu().numpy(), mid), rep
        finally:
            sb.set_stream(None)


@pytest.mark.parametrize("kind_in,kind_out", [("pageable", "pageable"), ("pinned", "pageable"), ("pageable", "pinned"),
                                              ("pinned", "pinned"), ("managed", "pageable"), ("pageable", "managed"),
                                              ("managed", "managed")])
def test_caller_memory_kinds(sb, restatement, kind_in, kind_out):
    """malloc'd, pinned and managed buffers on either side, on jobs large enough for the banded
    pipeline (and the pinned bounce buffers pageable memory goes through) and on a small one."""
    import torch
    from cuda.bindings import runtime as cudart
    jobs = [(cases.BGRA8_P, 2560, 1440, cases.BGRA8_U, 1280, 720, 0), (cases.RGB8, 700, 500, cases.ARGB8_P, 1900, 1300, 0),
            (cases.RGBA8_U, 3000, 2000, cases.BGR8, 301, 199, 1), (cases.ABGR8_P, 90, 70, cases.RGB8, 33, 21, 0)]
    managed = []

    def buf(kind, n, fill=None):
        if kind == "pageable":
            a = np.empty(n, np.uint8)
        elif kind == "pinned":
            a = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
        else:
            err, ptr = cudart.cudaMallocManaged(n, cudart.cudaMemAttachGlobal)
            assert int(err) == 0, err
            managed.append(ptr)
            a = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(int(ptr)))
        if fill is not None:
            a[:] = fill
        return a

    try:
        for ti, wi, hi, to, wo, ho, srgb in jobs:
            img = cases.make_image(ti, wi, hi, None, "premul", seed=3)
            want = _want(restatement, img, ti, wi, hi, to, wo, ho, srgb)
            src = buf(kind_in, img.size, img)
            out = buf(kind_out, want.size, 0xCD)
            sb.scale_simple(src, ti, wi, hi, wi * cases.bpp(ti), out, to, wo, ho, wo * cases.bpp(to), srgb)
            if kind_out == "managed" and kind_in == "managed":
                torch.cuda.synchronize()        # device path: returns at once
            assert np.array_equal(out, want), (kind_in, kind_out, ti, wi, hi, to, wo, ho)
            # odd pitches and a row band in the middle
            si, so = wi * cases.bpp(ti) + 5, wo * cases.bpp(to) + 3
            img2 = cases.make_image(ti, wi, hi, si, "random", seed=4)
            want2 = restatement.scale_simple(img2, ti, wi, hi, si, to, wo, ho, so, srgb)
            src2 = buf(kind_in, img2.size, img2)
            y0, n = ho // 5, ho // 2
            dest = buf(kind_out, so * (n - 1) + wo * cases.bpp(to), 0xCD)
            ctx = sb.ScaleCtx(src2, ti, wi, hi, si, None, to, wo, ho, so, srgb)
            ctx.batch_full(dest, y0, n)
            ctx.destroy()
            if kind_out == "managed" and kind_in == "managed":
                torch.cuda.synchronize()
            ref = want2[y0 * so: y0 * so + dest.size].copy()
            got = np.array(dest)
            # pitch padding of the destination is never written
            pad = np.ones(dest.size, bool)
            for r in range(n):
                pad[r * so: r * so + wo * cases.bpp(to)] = False
            assert (got[pad] == 0xCD).all()
            assert np.array_equal(got[~pad], ref[~pad]), (kind_in, kind_out, ti, wi, hi, to, wo, ho, "band")
    finally:
        import torch
        torch.cuda.synchronize()
        for ptr in managed:
            cudart.cudaFree(ptr)


def test_height_preserving_bands_on_host_buffers(sb, restatement):
    """Row batches of jobs whose vertical filter is COPY (or ONE), on host buffers, that stop short of
    the last image row: the kernels fetch the row below every row they need (weight 0), so the staged
    band must contain it (tools/sanitize_matrix.py runs the same jobs under compute-sanitizer)."""
    for ti, wi, hi, to, wo, ho, srgb in [(cases.BGRA8_P, 1920, 1080, cases.BGRA8_P, 960, 1080, 0),
                                         (cases.RGBA8_U, 1280, 720, cases.ABGR8_P, 1280, 720, 0),
                                         (cases.RGBA8_U, 640, 360, cases.BGRA8_U, 800, 360, 0),
                                         (cases.RGB8, 1000, 300, cases.RGB8, 333, 300, 1),
                                         (cases.ARGB8_P, 500, 1, cases.ARGB8_P, 500, 64, 0)]:
        si, so = wi * cases.bpp(ti), wo * cases.bpp(to)
        src = cases.make_image(ti, wi, hi, si, "random", seed=7)
        want = restatement.scale_simple(src, ti, wi, hi, si, to, wo, ho, so, srgb)
        out = np.zeros_like(want)
        ctx = sb.ScaleCtx(src, ti, wi, hi, si, out, to, wo, ho, so, srgb)
        T = 7
        per = (ho + T - 1) // T
        threads = [threading.Thread(target=ctx.batch, args=(y, min(per, ho - y))) for y in range(0, ho, per)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        ctx.destroy()
        assert np.array_equal(out, want), (ti, wi, hi, to, wo, ho)


def test_one_call_across_several_gpus(sb, restatement):
    """smol_cuda_set_multi_gpu: a host-memory call split into output row bands over every visible
    device (each uploads only its band + halo) == the single-device result.  Also: per-device table
    caches and kernel attributes with more than one device in ONE process."""
    import torch
    n_dev = sb.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs in one process")
    jobs = [(cases.BGRA8_P, 3840, 2160, cases.BGRA8_U, 1920, 1080, 0), (cases.RGBA8_P, 7680, 4320, cases.RGBA8_P, 800, 450, 1),
            (cases.RGB8, 1024, 768, cases.RGB8, 4096, 3072, 0), (cases.ARGB8_U, 4000, 3000, cases.ARGB8_U, 1500, 1100, 1),
            (cases.ARGB8_U, 2000, 1500, cases.ARGB8_U, 2300, 1700, 0)]
    for ti, wi, hi, to, wo, ho, srgb in jobs:
        src = cases.make_image(ti, wi, hi, None, "premul", seed=21)
        want = _want(restatement, src, ti, wi, hi, to, wo, ho, srgb)
        for n in sorted({2, n_dev}):
            sb.set_multi_gpu(n)
            try:
                sb.reset_stats()
                got = np.zeros_like(want)
                sb.scale_simple(src, ti, wi, hi, wi * cases.bpp(ti), got, to, wo, ho, wo * cases.bpp(to), srgb)
            finally:
                sb.set_multi_gpu(1)
            assert np.array_equal(got, want), (ti, wi, hi, to, wo, ho, n)
            # (the library uses fewer devices than allowed when a band would move less than ~8 MB)
            assert sb.stats()["kernel_launches"] >= 2
        # device-resident buffers on every device in turn (tables and attributes per device)
        for dev in range(n_dev):
            with torch.cuda.device(dev):
                d_in = torch.from_numpy(src).cuda()
                d_out = torch.zeros(want.size, dtype=torch.uint8, device="cuda")
                sb.set_stream(torch.cuda.current_stream().cuda_stream)
                sb.scale_simple(d_in, ti, wi, hi, wi * cases.bpp(ti), d_out, to, wo, ho, wo * cases.bpp(to), srgb)
                torch.cuda.synchronize()
                sb.set_stream(None)
                assert np.array_equal(d_out.cpu().numpy(), want), (dev, ti, to)


def test_first_use_inside_graph_capture(sb, restatement):
    """A geometry never seen before, first used while the caller is capturing a CUDA graph: the filter
    tables are uploaded outside the capture, the launch is captured, and replays give the right result."""
    import torch
    ti, wi, hi, to, wo, ho = cases.ABGR8_P, 1237, 811, cases.ARGB8_U, 613, 397      # not used by any other test
    src = cases.make_image(ti, wi, hi, None, "premul", seed=31)
    want = _want(restatement, src, ti, wi, hi, to, wo, ho)
    d_in = torch.from_numpy(src).cuda()
    d_out = torch.zeros(want.size, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        sb.set_stream(stream.cuda_stream)
        try:
            with torch.cuda.graph(graph, stream=stream):
                sb.scale_simple(d_in, ti, wi, hi, wi * 4, d_out, to, wo, ho, wo * 4, 0)
        finally:
            sb.set_stream(None)
    for _ in range(3):
        d_out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want)


def test_more_live_contexts_than_cache_slots(sb, restatement):
    """The reference has no limit on live contexts; here every context pins two cached filter tables
    (192 slots per device): past that, tables are allocated outside the cache instead of aborting."""
    ctxs, outs, wants = [], [], []
    ti = to = cases.RGBA8_P
    base = cases.make_image(ti, 700, 4, None, "random", seed=5)
    try:
        for k in range(230):
            wi, wo = 300 + k, 37 + k            # 230 distinct horizontal tables (+ a few vertical ones)
            src = base[: 4 * wi * 4].copy()
            out = np.zeros(wo * 3 * 4, np.uint8)
            ctxs.append(sb.ScaleCtx(src, ti, wi, 4, wi * 4, out, to, wo, 3, wo * 4, 0))
            outs.append(out)
            wants.append((src, wi, wo))
            ctxs[-1].batch(0, 1)                # forces the table upload while all earlier ones stay live
        for ctx, out, (src, wi, wo) in zip(ctxs, outs, wants):
            ctx.batch(1, 2)
            want = restatement.scale_simple(src, ti, wi, 4, wi * 4, to, wo, 3, wo * 4, 0)
            assert np.array_equal(out, want), (wi, wo)
    finally:
        for ctx in ctxs:
            ctx.destroy()

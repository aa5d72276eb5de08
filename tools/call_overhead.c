/* call_overhead.c -- host cost of one call through the C API, measured from C (no Python, no ctypes).
 *
 * A C caller that loops smol_scale_simple over device-resident frames is bound by
 * max (kernel time, host time per call).  This program times the host side alone: many calls on a
 * tiny job (the GPU finishes each one faster than the host can enqueue the next), per kernel family,
 * plus the full-size BASELINE shapes driven the same way (stream-ordered, one synchronize at the end)
 * so the "direct launch" figure has a C-level counterpart to bench.py's ctypes one.
 *
 * Build: gcc -O2 -o tools/call_overhead tools/call_overhead.c -Iinclude -I/usr/local/cuda/include \
 *            -L/usr/local/cuda/lib64 -lcudart_static -lpthread -lrt -ldl
 * Run:   tools/call_overhead smolscale_b200/libsmolscale_cuda.so */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "smolscale.h"

typedef void (*simple_fn) (const void *, SmolPixelType, uint32_t, uint32_t, uint32_t,
                           void *, SmolPixelType, uint32_t, uint32_t, uint32_t, uint8_t);
typedef SmolScaleCtx *(*new_fn) (const void *, SmolPixelType, uint32_t, uint32_t, uint32_t,
                                 void *, SmolPixelType, uint32_t, uint32_t, uint32_t, uint8_t);
typedef void (*batch_fn) (const SmolScaleCtx *, uint32_t, uint32_t);
typedef void (*destroy_fn) (SmolScaleCtx *);

static double now_s (void)
{
    struct timespec ts;
    clock_gettime (CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf (stderr, "%s: %s\n", #x, cudaGetErrorString (e_)); exit (1); } } while (0)

static int bpp (SmolPixelType t) { return t >= SMOL_PIXEL_RGB8 ? 3 : 4; }

int main (int argc, char **argv)
{
    void *lib = dlopen (argc > 1 ? argv[1] : "smolscale_b200/libsmolscale_cuda.so", RTLD_NOW);
    if (!lib) { fprintf (stderr, "%s\n", dlerror ()); return 1; }
    simple_fn simple = (simple_fn) dlsym (lib, "smol_scale_simple");
    new_fn ctx_new = (new_fn) dlsym (lib, "smol_scale_new");
    batch_fn batch = (batch_fn) dlsym (lib, "smol_scale_batch");
    destroy_fn destroy = (destroy_fn) dlsym (lib, "smol_scale_destroy");

    struct { const char *name; SmolPixelType ti; uint32_t wi, hi; SmolPixelType to; uint32_t wo, ho; uint8_t srgb; int calls; } jobs[] = {
        { "tiny half2x (64x64 -> 32x32)", SMOL_PIXEL_BGRA8_PREMULTIPLIED, 64, 64, SMOL_PIXEL_BGRA8_UNASSOCIATED, 32, 32, 0, 20000 },
        { "tiny taps (64x64 -> 41x37)", SMOL_PIXEL_RGBA8_PREMULTIPLIED, 64, 64, SMOL_PIXEL_RGBA8_PREMULTIPLIED, 41, 37, 0, 20000 },
        { "tiny box sRGB (400x400 -> 16x16)", SMOL_PIXEL_RGBA8_PREMULTIPLIED, 400, 400, SMOL_PIXEL_RGBA8_PREMULTIPLIED, 16, 16, 1, 20000 },
        { "tiny magb (16x16 RGB -> 64x64)", SMOL_PIXEL_RGB8, 16, 16, SMOL_PIXEL_RGB8, 64, 64, 0, 20000 },
        { "tiny taps128 (64x64 -> 30x30 sRGB)", SMOL_PIXEL_RGBA8_UNASSOCIATED, 64, 64, SMOL_PIXEL_RGBA8_UNASSOCIATED, 30, 30, 1, 20000 },
        { "cfg1 1080p -> 540p", SMOL_PIXEL_RGBA8_PREMULTIPLIED, 1920, 1080, SMOL_PIXEL_RGBA8_PREMULTIPLIED, 960, 540, 0, 4000 },
        { "cfg2 4K -> 1080p", SMOL_PIXEL_BGRA8_PREMULTIPLIED, 3840, 2160, SMOL_PIXEL_BGRA8_UNASSOCIATED, 1920, 1080, 0, 2000 },
        { "cfg4 RGB 1024x768 -> 4096x3072", SMOL_PIXEL_RGB8, 1024, 768, SMOL_PIXEL_RGB8, 4096, 3072, 0, 1000 },
    };
    const int n_jobs = (int) (sizeof (jobs) / sizeof (jobs[0]));
    const int ring = 8;        /* distinct buffers the calls rotate through (full-size shapes: more than L2) */
    cudaStream_t stream;

    CK (cudaSetDevice (0));
    CK (cudaStreamCreate (&stream));
    void (*set_stream) (void *) = (void (*) (void *)) dlsym (lib, "smol_cuda_set_stream");
    set_stream ((void *) stream);

    printf ("{\"what\": \"host microseconds per C call, device-resident buffers, stream-ordered\", \"jobs\": [\n");
    for (int j = 0; j < n_jobs; j++)
    {
        const size_t si = (size_t) jobs[j].wi * bpp (jobs[j].ti), so = (size_t) jobs[j].wo * bpp (jobs[j].to);
        const size_t in_bytes = si * jobs[j].hi, out_bytes = so * jobs[j].ho;
        char *d_in, *d_out;
        CK (cudaMalloc ((void **) &d_in, in_bytes * ring));
        CK (cudaMalloc ((void **) &d_out, out_bytes * ring));
        CK (cudaMemset (d_in, 0x55, in_bytes * ring));
        for (int w = 0; w < 3; w++)
            simple (d_in, jobs[j].ti, jobs[j].wi, jobs[j].hi, (uint32_t) si, d_out, jobs[j].to, jobs[j].wo, jobs[j].ho, (uint32_t) so, jobs[j].srgb);
        CK (cudaStreamSynchronize (stream));

        const int n = jobs[j].calls;
        double t0 = now_s ();
        for (int i = 0; i < n; i++)
            simple (d_in + (size_t) (i % ring) * in_bytes, jobs[j].ti, jobs[j].wi, jobs[j].hi, (uint32_t) si,
                    d_out + (size_t) (i % ring) * out_bytes, jobs[j].to, jobs[j].wo, jobs[j].ho, (uint32_t) so, jobs[j].srgb);
        double t_enq = now_s () - t0;
        CK (cudaStreamSynchronize (stream));
        double t_all = now_s () - t0;

        /* context reuse: smol_scale_new once, smol_scale_batch per call */
        SmolScaleCtx *ctx = ctx_new (d_in, jobs[j].ti, jobs[j].wi, jobs[j].hi, (uint32_t) si, d_out, jobs[j].to, jobs[j].wo, jobs[j].ho, (uint32_t) so, jobs[j].srgb);
        batch (ctx, 0, jobs[j].ho);
        CK (cudaStreamSynchronize (stream));
        t0 = now_s ();
        for (int i = 0; i < n; i++)
            batch (ctx, 0, jobs[j].ho);
        double t_batch = now_s () - t0;
        CK (cudaStreamSynchronize (stream));
        destroy (ctx);

        printf ("  {\"job\": \"%s\", \"calls\": %d, \"smol_scale_simple_enqueue_us\": %.2f, \"smol_scale_simple_us_incl_gpu\": %.2f, "
                "\"smol_scale_batch_enqueue_us\": %.2f}%s\n",
                jobs[j].name, n, t_enq / n * 1e6, t_all / n * 1e6, t_batch / n * 1e6, j + 1 < n_jobs ? "," : "");
        fflush (stdout);
        CK (cudaFree (d_in));
        CK (cudaFree (d_out));
    }
    printf ("]}\n");
    return 0;
}

/* TEST / BENCH INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * A small pthread driver for any library that implements smolscale.h.  It restates the
 * threaded-caller pattern of the reference's own benchmark program (test.c:811-883: one
 * smol_scale_new, then T workers each calling smol_scale_batch on ceil(H_out / T) consecutive
 * rows, then smol_scale_destroy), with plain pthreads instead of a GLib thread pool.  The library
 * under test is dlopen()ed by path, so the same harness times oracle/_ref/libsmolref.so
 * (generic), oracle/_ref/libsmolref_avx2.so (AVX2) and, for plumbing tests, our own
 * libsmolscale_cuda.so. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void *(*new_fn) (const void *, int, uint32_t, uint32_t, uint32_t,
                         void *, int, uint32_t, uint32_t, uint32_t, uint8_t);
typedef void (*batch_full_fn) (const void *, void *, uint32_t, uint32_t);
typedef void (*destroy_fn) (void *);
typedef void (*simple_fn) (const void *, int, uint32_t, uint32_t, uint32_t,
                           void *, int, uint32_t, uint32_t, uint32_t, uint8_t);

typedef struct
{
    void *dl;
    new_fn scale_new;
    batch_full_fn batch_full;
    destroy_fn destroy;
    simple_fn simple;
} harness;

void *
harness_open (const char *path)
{
    harness *h = calloc (1, sizeof (*h));

    h->dl = dlopen (path, RTLD_NOW | RTLD_LOCAL);
    if (!h->dl)
    {
        fprintf (stderr, "ref_harness: %s\n", dlerror ());
        free (h);
        return NULL;
    }
    h->scale_new = (new_fn) dlsym (h->dl, "smol_scale_new");
    h->batch_full = (batch_full_fn) dlsym (h->dl, "smol_scale_batch_full");
    h->destroy = (destroy_fn) dlsym (h->dl, "smol_scale_destroy");
    h->simple = (simple_fn) dlsym (h->dl, "smol_scale_simple");
    if (!h->scale_new || !h->batch_full || !h->destroy || !h->simple)
    {
        fprintf (stderr, "ref_harness: %s lacks the smolscale.h entry points\n", path);
        dlclose (h->dl);
        free (h);
        return NULL;
    }
    return h;
}

void
harness_close (void *hp)
{
    harness *h = hp;

    if (!h)
        return;
    dlclose (h->dl);
    free (h);
}

static double
now_s (void)
{
    struct timespec ts;

    clock_gettime (CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* ---- one image, rows split across threads ---- */

typedef struct
{
    harness *h;
    void *ctx;
    uint8_t *out;
    uint32_t stride_out;
    uint32_t first, n;
} band_job;

static void *
band_worker (void *arg)
{
    band_job *j = arg;

    j->h->batch_full (j->ctx, j->out + (size_t) j->first * j->stride_out, j->first, j->n);
    return NULL;
}

/* Scales one image `reps` times with `n_threads` row-band workers; returns the best wall time
 * in seconds (CLOCK_MONOTONIC around new + workers + destroy, as test.c:1033-1035 times one
 * whole call). */
double
harness_scale_threaded (void *hp,
                        const void *in, int type_in, uint32_t w_in, uint32_t h_in, uint32_t stride_in,
                        void *out, int type_out, uint32_t w_out, uint32_t h_out, uint32_t stride_out,
                        uint8_t with_srgb, uint32_t n_threads, uint32_t reps)
{
    harness *h = hp;
    double best = 1e30;
    uint32_t rep;

    if (n_threads < 1)
        n_threads = 1;
    if (n_threads > h_out)
        n_threads = h_out;

    for (rep = 0; rep < reps; rep++)
    {
        pthread_t *tids = calloc (n_threads, sizeof (pthread_t));
        band_job *jobs = calloc (n_threads, sizeof (band_job));
        uint32_t rows_per = (h_out + n_threads - 1) / n_threads;
        uint32_t n_jobs = 0, row;
        double t0 = now_s (), t1;
        void *ctx;

        ctx = h->scale_new (in, type_in, w_in, h_in, stride_in,
                            out, type_out, w_out, h_out, stride_out, with_srgb);
        for (row = 0; row < h_out; row += rows_per)
        {
            band_job *j = &jobs[n_jobs];

            j->h = h; j->ctx = ctx; j->out = out; j->stride_out = stride_out;
            j->first = row;
            j->n = (h_out - row < rows_per) ? h_out - row : rows_per;
            if (n_threads == 1)
                band_worker (j);
            else
                pthread_create (&tids[n_jobs], NULL, band_worker, j);
            n_jobs++;
        }
        if (n_threads > 1)
            for (row = 0; row < n_jobs; row++)
                pthread_join (tids[row], NULL);
        h->destroy (ctx);
        t1 = now_s ();
        if (t1 - t0 < best)
            best = t1 - t0;
        free (jobs);
        free (tids);
    }
    return best;
}

/* ---- many images, images split across threads (thumbnail batches) ---- */

typedef struct
{
    harness *h;
    const uint8_t *in; uint8_t *out;
    size_t in_image_bytes, out_image_bytes;
    int type_in, type_out;
    uint32_t w_in, h_in, stride_in, w_out, h_out, stride_out;
    uint8_t with_srgb;
    uint32_t first, n;
} image_job;

static void *
image_worker (void *arg)
{
    image_job *j = arg;
    uint32_t i;

    for (i = j->first; i < j->first + j->n; i++)
        j->h->simple (j->in + i * j->in_image_bytes, j->type_in, j->w_in, j->h_in, j->stride_in,
                      j->out + i * j->out_image_bytes, j->type_out, j->w_out, j->h_out, j->stride_out,
                      j->with_srgb);
    return NULL;
}

/* n_images images stored back to back (in_image_bytes / out_image_bytes apart); each worker
 * runs smol_scale_simple over its share.  Returns wall seconds for the whole batch. */
double
harness_scale_images (void *hp,
                      const void *in, size_t in_image_bytes, int type_in,
                      uint32_t w_in, uint32_t h_in, uint32_t stride_in,
                      void *out, size_t out_image_bytes, int type_out,
                      uint32_t w_out, uint32_t h_out, uint32_t stride_out,
                      uint8_t with_srgb, uint32_t n_images, uint32_t n_threads)
{
    harness *h = hp;
    pthread_t *tids;
    image_job *jobs;
    uint32_t per, n_jobs = 0, i;
    double t0, t1;

    if (n_threads < 1)
        n_threads = 1;
    if (n_threads > n_images)
        n_threads = n_images;
    tids = calloc (n_threads, sizeof (pthread_t));
    jobs = calloc (n_threads, sizeof (image_job));
    per = (n_images + n_threads - 1) / n_threads;

    t0 = now_s ();
    for (i = 0; i < n_images; i += per)
    {
        image_job *j = &jobs[n_jobs];

        j->h = h; j->in = in; j->out = out;
        j->in_image_bytes = in_image_bytes; j->out_image_bytes = out_image_bytes;
        j->type_in = type_in; j->type_out = type_out;
        j->w_in = w_in; j->h_in = h_in; j->stride_in = stride_in;
        j->w_out = w_out; j->h_out = h_out; j->stride_out = stride_out;
        j->with_srgb = with_srgb;
        j->first = i;
        j->n = (n_images - i < per) ? n_images - i : per;
        pthread_create (&tids[n_jobs], NULL, image_worker, j);
        n_jobs++;
    }
    for (i = 0; i < n_jobs; i++)
        pthread_join (tids[i], NULL);
    t1 = now_s ();
    free (jobs);
    free (tids);
    return t1 - t0;
}

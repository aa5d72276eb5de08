/* smolscale-cuda.c -- host side of the B200 implementation of the smolscale.h API.
 *
 * What lives here (all plain C; the only CUDA it touches is the runtime's C API):
 *   - job planning: which filter each axis gets, which intermediate encoding, the fixed-point
 *     offset / weight tables -- integer code that must reproduce the reference's choices exactly
 *     (reference smolscale.c:427-478, :724-814; smolscale-generic.c:14-135);
 *   - pointer classification (host / pinned / managed / device) and, for host memory, staging of
 *     exactly the source row band (+ filter halo) a batch needs;
 *   - per-device state: the data tables, a cache of uploaded filter tables keyed by geometry,
 *     a small pool of stream + staging-buffer "lanes" so concurrent smol_scale_batch callers
 *     (the reference's threading model, smolscale.h:70-74) overlap instead of serialising;
 *   - the seven public entry points and the optional smol_cuda_* extensions.
 *
 * There is no CPU implementation of the pipeline in this file or anywhere in the product: if
 * CUDA is unusable the entry points abort with a message. */

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#if defined (__SSE2__)
#include <emmintrin.h>
#endif

#include <cuda_runtime_api.h>

#include "smolscale.h"
#include "smolscale-cuda.h"
#include "smolscale-cuda-private.h"
#include "smolscale-cuda-luts.h"

#define SMOL_MAX_DEVICES 16
#define SMOL_MAX_LANES 16
#define SMOL_TAB_CACHE_MAX 192
#define SMOL_PLAN_CACHE_MAX 64

#define SMOL_EXPORT __attribute__ ((visibility ("default")))

static void
smol_fatal (const char *what, const char *detail)
{
    fprintf (stderr, "smolscale-cuda: fatal: %s%s%s\n", what, detail ? ": " : "", detail ? detail : "");
    fflush (stderr);
    abort ();
}

#define CK(call) \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) smol_fatal (#call, cudaGetErrorString (e_)); } while (0)

/* ------------------------------------------------------------------------------------------ *
 * Planning (pure host code)                                                                  *
 * ------------------------------------------------------------------------------------------ */

typedef struct
{
    int filter;             /* SMOL_CUDA_AXIS_* (the reference's choice) */
    uint32_t halvings;
    uint32_t bilin_dim;
    int storage_bits;
}
AxisPick;

/* reference smolscale.c:427-478 */
static AxisPick
pick_axis (uint32_t dim_in, uint32_t dim_out, uint8_t with_srgb)
{
    AxisPick a;

    a.halvings = 0;
    a.bilin_dim = dim_out;
    a.storage_bits = with_srgb ? 128 : 64;

    if (dim_in > dim_out * 255)
    {
        a.filter = SMOL_CUDA_AXIS_BOX;
        a.storage_bits = 128;
    }
    else if (dim_in > dim_out * 8)
    {
        a.filter = SMOL_CUDA_AXIS_BOX;
    }
    else if (dim_in == 1)
    {
        a.filter = SMOL_CUDA_AXIS_ONE;
    }
    else if (dim_in == dim_out)
    {
        a.filter = SMOL_CUDA_AXIS_COPY;
    }
    else
    {
        uint32_t d = dim_out;

        for (;;)
        {
            d *= 2;
            if (d >= dim_in)
                break;
            a.halvings++;
        }
        a.filter = SMOL_CUDA_AXIS_BILINEAR;
        a.bilin_dim = dim_out << a.halvings;
    }
    return a;
}

typedef struct
{
    int alpha_idx;      /* -1: none */
    int col0;
    int bgr;
    int unassoc;
    int bpp;
}
TypeInfo;

/* Memory layout of the ten public pixel types (reference smolscale.h:14-35, smolscale.c:45-75). */
static TypeInfo
type_info (SmolPixelType t)
{
    TypeInfo ti;

    if ((int) t < 0 || t >= SMOL_PIXEL_MAX)
        smol_fatal ("invalid SmolPixelType", NULL);

    if (t == SMOL_PIXEL_RGB8 || t == SMOL_PIXEL_BGR8)
    {
        ti.bpp = 3; ti.alpha_idx = -1; ti.col0 = 0; ti.unassoc = 0;
        ti.bgr = (t == SMOL_PIXEL_BGR8);
        return ti;
    }
    ti.bpp = 4;
    ti.unassoc = (t >= SMOL_PIXEL_RGBA8_UNASSOCIATED);
    switch ((int) t & 3)
    {
        case 0: ti.alpha_idx = 3; ti.col0 = 0; ti.bgr = 0; break;   /* RGBA */
        case 1: ti.alpha_idx = 3; ti.col0 = 0; ti.bgr = 1; break;   /* BGRA */
        case 2: ti.alpha_idx = 0; ti.col0 = 1; ti.bgr = 0; break;   /* ARGB */
        default: ti.alpha_idx = 0; ti.col0 = 1; ti.bgr = 1; break;  /* ABGR */
    }
    return ti;
}

/* Reference-layout tables: (absolute offset, F) uint16 pairs. */

/* reference smolscale-generic.c:14-66 (offsets kept absolute for both axes) */
static uint16_t *
make_bilinear_pairs (uint32_t dim_in, uint32_t n)
{
    uint16_t *t = malloc ((size_t) n * 2 * sizeof (uint16_t));
    const uint64_t unit = (uint64_t) 1 << 32;
    uint64_t step, pos;
    uint32_t i;

    if (!t)
        smol_fatal ("out of memory", NULL);

    if (dim_in > n)
    {
        step = ((uint64_t) dim_in * unit) / n;
        pos = (step - unit) / 2;
    }
    else
    {
        step = ((uint64_t) (dim_in - 1) * unit) / (n > 1 ? n - 1 : 1);
        pos = 0;
    }

    for (i = 0; i < n; i++)
    {
        uint64_t at = pos + (uint64_t) i * step;
        uint32_t ofs = (uint32_t) (at >> 32) & 0xffff;

        if (ofs >= dim_in - 1)
            break;
        t[2 * i] = (uint16_t) ofs;
        t[2 * i + 1] = (uint16_t) (256 - ((at >> 24) & 0xff));
    }
    /* once a sample would need the pixel past the end: 100 % of the last pixel */
    for (; i < n; i++)
    {
        t[2 * i] = (uint16_t) (dim_in - 2);
        t[2 * i + 1] = 0;
    }
    return t;
}

/* reference smolscale-generic.c:68-135 (absolute offsets, dim_out + 1 pairs) */
static uint16_t *
make_box_pairs (uint32_t dim_in, uint32_t dim_out, uint32_t *span_mul)
{
    uint16_t *t = malloc (((size_t) dim_out + 1) * 2 * sizeof (uint16_t));
    const uint64_t step = ((uint64_t) dim_in << 16) / dim_out;
    uint64_t whole = step >> 16, part = (step >> 8) & 0xff;
    uint64_t num = ((uint64_t) 255) << 24;
    uint64_t den = whole * 255 + (part * 255) / 256;
    uint32_t begin = 0, i;

    if (!t)
        smol_fatal ("out of memory", NULL);
    *span_mul = (uint32_t) ((num + den / 2) / den);

    for (i = 0; i < dim_out; i++)
    {
        uint64_t at = (uint64_t) (i + 1) * step;
        uint32_t end = (uint32_t) (at >> 16) & 0xffff;

        if (begin >= dim_in - 1)
        {
            begin = dim_in - 1;
            break;
        }
        if (end > dim_in - 1)
        {
            end = dim_in - 1;
            if (end <= begin)
                break;
        }
        t[2 * i] = (uint16_t) begin;
        t[2 * i + 1] = (uint16_t) ((at >> 8) & 0xff);
        begin = end;
    }
    for (; i <= dim_out; i++)
    {
        t[2 * i] = (uint16_t) begin;
        t[2 * i + 1] = 0;
    }
    return t;
}

typedef struct
{
    int filter;                 /* SMOL_CUDA_AXIS_* */
    uint32_t dim_in, dim_out;
    uint32_t halvings, bilin_dim;
    uint32_t span_mul;
    uint16_t *pairs;            /* reference-layout pairs, NULL for copy / one */
    uint32_t n_pairs;
    uint32_t *dev_entries;      /* host copy of the device table (packed ofs | F << 16) */
    uint32_t n_entries;
    uint8_t kind;               /* SMOL_AXIS_* */
    uint8_t all_half;
}
AxisPlan;

static void
axis_plan_init (AxisPlan *ap, AxisPick pick, uint32_t dim_in, uint32_t dim_out)
{
    uint32_t i;

    memset (ap, 0, sizeof (*ap));
    ap->filter = pick.filter;
    ap->dim_in = dim_in;
    ap->dim_out = dim_out;
    ap->halvings = pick.halvings;
    ap->bilin_dim = pick.bilin_dim;

    if (pick.filter == SMOL_CUDA_AXIS_BOX)
    {
        ap->kind = SMOL_AXIS_BOX;
        ap->pairs = make_box_pairs (dim_in, dim_out, &ap->span_mul);
        ap->n_pairs = dim_out + 1;
        ap->n_entries = dim_out + 1;
    }
    else
    {
        ap->kind = SMOL_AXIS_TAPS;
        ap->n_entries = pick.bilin_dim;
        if (pick.filter == SMOL_CUDA_AXIS_BILINEAR)
        {
            ap->pairs = make_bilinear_pairs (dim_in, pick.bilin_dim);
            ap->n_pairs = pick.bilin_dim;
        }
    }

    ap->dev_entries = malloc ((size_t) ap->n_entries * sizeof (uint32_t));
    if (!ap->dev_entries)
        smol_fatal ("out of memory", NULL);

    for (i = 0; i < ap->n_entries; i++)
    {
        uint32_t ofs, F;

        if (ap->pairs)
        {
            ofs = ap->pairs[2 * i];
            F = ap->pairs[2 * i + 1];
        }
        else if (pick.filter == SMOL_CUDA_AXIS_COPY)
        {
            ofs = i; F = 256;       /* reference generic:1591-1611, :2306-2318 */
        }
        else
        {
            ofs = 0; F = 256;       /* reference generic:1558-1589, :2262-2304 */
        }
        ap->dev_entries[i] = ofs | (F << 16);
    }

    ap->all_half = 0;
    if (pick.filter == SMOL_CUDA_AXIS_BILINEAR)
    {
        ap->all_half = 1;
        for (i = 0; i < ap->n_entries; i++)
            if (ap->dev_entries[i] != ((2 * i) | (128u << 16)))
            {
                ap->all_half = 0;
                break;
            }
    }
}

static void
axis_plan_clear (AxisPlan *ap)
{
    free (ap->pairs);
    free (ap->dev_entries);
    ap->pairs = NULL;
    ap->dev_entries = NULL;
}

typedef struct
{
    SmolJobDesc d;
    AxisPlan ax, ay;
    int storage_bits;
}
JobPlan;

static void
job_plan_init (JobPlan *jp,
               SmolPixelType type_in, uint32_t w_in, uint32_t h_in,
               SmolPixelType type_out, uint32_t w_out, uint32_t h_out,
               uint8_t with_srgb)
{
    TypeInfo ti = type_info (type_in), to = type_info (type_out);
    AxisPick px = pick_axis (w_in, w_out, with_srgb);
    AxisPick py = pick_axis (h_in, h_out, with_srgb);
    int linear = with_srgb ? 1 : 0;
    int both_unassoc = ti.unassoc && to.unassoc;
    SmolJobDesc *d = &jp->d;

    memset (jp, 0, sizeof (*jp));
    jp->storage_bits = px.storage_bits > py.storage_bits ? px.storage_bits : py.storage_bits;   /* smolscale.c:862 */

    /* unassociated in and out: 16 bits per channel internally (smolscale.c:751-758) */
    if (both_unassoc)
        jp->storage_bits = 128;
    /* not enough head room for linear light beyond 2^13 : 1 (smolscale.c:760-770) */
    if (w_in > w_out * 8191 || h_in > h_out * 8191)
        linear = 0;

    axis_plan_init (&jp->ax, px, w_in, w_out);
    axis_plan_init (&jp->ay, py, h_in, h_out);

    d->w_in = w_in; d->h_in = h_in; d->w_out = w_out; d->h_out = h_out;
    d->bpp_in = (uint8_t) ti.bpp; d->bpp_out = (uint8_t) to.bpp;
    d->in_alpha_idx = ti.alpha_idx < 0 ? 0xff : (uint8_t) ti.alpha_idx;
    d->in_col0 = (uint8_t) ti.col0;
    d->out_alpha_idx = to.alpha_idx < 0 ? 0xff : (uint8_t) to.alpha_idx;
    d->out_col0 = (uint8_t) to.col0;
    d->in_unassoc = (uint8_t) ti.unassoc;
    d->out_unassoc = (uint8_t) to.unassoc;
    d->swap_rb = (uint8_t) (ti.bgr != to.bgr);
    d->storage128 = jp->storage_bits == 128;
    if (both_unassoc)
        d->mid = linear ? SMOL_MID_P16L : SMOL_MID_P16;
    else
        d->mid = linear ? SMOL_MID_P8L : SMOL_MID_P8;
    /* Linear light to a 24bpp destination: the reference's repack search (smolscale.c:647-719)
     * lands on its "123" packer -- which gamma-compresses the still-premultiplied value
     * (generic:922-935) -- for 32bpp sources whose colour order is reversed in the output and for
     * 24bpp sources whose colour order is kept; on the "321" packer (unpremultiply first,
     * generic:1010-1023) otherwise. */
    d->pack24_direct = (uint8_t) ((ti.bpp == 4) ? d->swap_rb : !d->swap_rb);
    d->h_kind = jp->ax.kind; d->v_kind = jp->ay.kind;
    d->h_halvings = (uint8_t) jp->ax.halvings; d->v_halvings = (uint8_t) jp->ay.halvings;
    d->all_half_x = jp->ax.all_half; d->all_half_y = jp->ay.all_half;
    d->span_mul_x = jp->ax.span_mul; d->span_mul_y = jp->ay.span_mul;
    d->n_tab_x = jp->ax.n_entries; d->n_tab_y = jp->ay.n_entries;
}

static void
job_plan_clear (JobPlan *jp)
{
    axis_plan_clear (&jp->ax);
    axis_plan_clear (&jp->ay);
}

/* Source rows read by output rows [first, first + n): [*r0, *r0 + *nr). */
static void
plan_source_rows (const JobPlan *jp, uint32_t first, uint32_t n, uint32_t *r0, uint32_t *nr)
{
    const AxisPlan *ay = &jp->ay;
    uint32_t lo, hi;

    if (n == 0)
    {
        *r0 = 0; *nr = 0;
        return;
    }
    if (ay->kind == SMOL_AXIS_BOX)
    {
        lo = ay->dev_entries[first] & 0xffff;
        hi = ay->dev_entries[first + n] & 0xffff;
    }
    else
    {
        uint32_t i0 = first << ay->halvings;
        uint32_t i1 = ((first + n) << ay->halvings) - 1;

        lo = ay->dev_entries[i0] & 0xffff;
        hi = (ay->dev_entries[i1] & 0xffff) + 1;
        /* ONE has all offsets 0 and F = 256; COPY needs only its own rows */
        if (ay->filter != SMOL_CUDA_AXIS_BILINEAR)
            hi = ay->dev_entries[i1] & 0xffff;
    }
    if (hi > jp->d.h_in - 1)
        hi = jp->d.h_in - 1;
    *r0 = lo;
    *nr = hi - lo + 1;
}

/* ------------------------------------------------------------------------------------------ *
 * Per-device state                                                                           *
 * ------------------------------------------------------------------------------------------ */

typedef struct
{
    int used;
    uint8_t kind;
    uint32_t filter, dim_in, n_entries;
    uint32_t *dev;
    int refs;
    int uncached;           /* allocated outside the cache (every slot pinned): freed with its last reference */
    uint64_t stamp;
}
TabEntry;

#define SMOL_MAX_BANDS 16

typedef struct
{
    cudaStream_t stream;                /* kernels (and everything, for small jobs) */
    cudaStream_t s_h2d, s_d2h;          /* copy streams of the banded pipeline */
    cudaEvent_t ev_up[SMOL_MAX_BANDS], ev_done[SMOL_MAX_BANDS], ev_down[SMOL_MAX_BANDS];
    int pipeline_ready;
    void *d_in, *d_out;                 /* device staging */
    size_t d_in_cap, d_out_cap;
    void *h_in, *h_out;                 /* pinned bounce buffers for pageable caller memory */
    size_t h_in_cap, h_out_cap;
    int busy;
}
Lane;

typedef struct
{
    int ready;
    cudaStream_t s_upload;
    SmolDeviceLuts *luts;
    uint16_t *p8l_from_p, *p8l_from_u;
    TabEntry tabs[SMOL_TAB_CACHE_MAX];
    uint64_t clock;
    Lane lanes[SMOL_MAX_LANES];
    int n_lanes;
}
DeviceState;

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_lane_free = PTHREAD_COND_INITIALIZER;
static DeviceState g_dev[SMOL_MAX_DEVICES];
static int g_device_count = -1;
static int g_forced_kernel = 0;

static uint64_t g_stat_launches, g_stat_h2d, g_stat_d2h, g_stat_uploads;
static uint64_t g_stat_by_kernel[SMOL_CUDA_MAX_KERNEL_FAMILIES];

static __thread void *tl_stream = NULL;
static __thread int tl_device = -1;

static int
device_count_locked (void)
{
    if (g_device_count < 0)
    {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount (&n);

        if (e != cudaSuccess)
        {
            (void) cudaGetLastError ();
            n = 0;
        }
        if (n > SMOL_MAX_DEVICES)
            n = SMOL_MAX_DEVICES;
        g_device_count = n;
    }
    return g_device_count;
}

static void *device_upload (DeviceState *ds, const void *host, size_t bytes);

/* Caller holds g_lock and has made `dev` current. */
static DeviceState *
device_state_locked (int dev)
{
    DeviceState *ds;

    if (device_count_locked () <= 0)
        smol_fatal ("no usable CUDA device (this library has no CPU fallback)", NULL);
    if (dev < 0 || dev >= g_device_count)
        smol_fatal ("device ordinal out of range", NULL);

    ds = &g_dev[dev];
    if (!ds->ready)
    {
        SmolDeviceLuts *h = malloc (sizeof (*h));

        if (!h)
            smol_fatal ("out of memory", NULL);
        memcpy (h->inv_div_p8, smol_lut_inv_div_p8, sizeof (h->inv_div_p8));
        memcpy (h->inv_div_p8l, smol_lut_inv_div_p8l, sizeof (h->inv_div_p8l));
        memcpy (h->inv_div_p16, smol_lut_inv_div_p16, sizeof (h->inv_div_p16));
        memcpy (h->inv_div_p16l, smol_lut_inv_div_p16l, sizeof (h->inv_div_p16l));
        memcpy (h->from_srgb, smol_lut_from_srgb, sizeof (h->from_srgb));
        memcpy (h->to_srgb, smol_lut_to_srgb, sizeof (h->to_srgb));
        ds->luts = device_upload (ds, h, sizeof (*h));
        free (h);
        {
            /* Composite per-channel unpack tables for linear light, index (alpha << 8) | c:
             *   premultiplied source (reference generic:555-568): unpremultiply with the p8 table,
             *   from_srgb, premultiply to 11 bits;  unassociated source (generic:591-614): the
             *   last two steps only.  Pure functions of the six LUTs above. */
            uint16_t *tp = malloc (65536 * sizeof (uint16_t)), *tu = malloc (65536 * sizeof (uint16_t));
            uint32_t a, c;

            if (!tp || !tu)
                smol_fatal ("out of memory", NULL);
            for (a = 0; a < 256; a++)
                for (c = 0; c < 256; c++)
                {
                    const uint32_t m = (a << 3) + 1;
                    const uint32_t u = ((c * smol_lut_inv_div_p8[a]) >> 13) & 0xff;
                    tp[(a << 8) | c] = (uint16_t) ((((smol_lut_from_srgb[u] + 1u) * m - 1u) >> 11) & 0x7ff);
                    tu[(a << 8) | c] = (uint16_t) ((((smol_lut_from_srgb[c] + 1u) * m - 1u) >> 11) & 0x7ff);
                }
            ds->p8l_from_p = device_upload (ds, tp, 65536 * sizeof (uint16_t));
            ds->p8l_from_u = device_upload (ds, tu, 65536 * sizeof (uint16_t));
            free (tp);
            free (tu);
        }
        ds->ready = 1;
    }
    return ds;
}

/* Allocation + upload of a small table on the current device.  Runs with the calling thread's
 * stream-capture mode relaxed, so the first use of a geometry from inside a CUDA graph capture
 * uploads for real instead of invalidating the capture (the kernel launch that follows is what
 * gets captured).  The copy goes through a private stream: nothing here touches the caller's
 * stream or the legacy default stream. */
static void *
device_upload (DeviceState *ds, const void *host, size_t bytes)
{
    enum cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    void *dev = NULL;

    CK (cudaThreadExchangeStreamCaptureMode (&mode));
    if (!ds->s_upload)
        CK (cudaStreamCreateWithFlags (&ds->s_upload, cudaStreamNonBlocking));
    CK (cudaMalloc (&dev, bytes));
    CK (cudaMemcpyAsync (dev, host, bytes, cudaMemcpyHostToDevice, ds->s_upload));
    CK (cudaStreamSynchronize (ds->s_upload));
    CK (cudaThreadExchangeStreamCaptureMode (&mode));
    return dev;
}

static void
device_release (void *dev)
{
    enum cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;

    CK (cudaThreadExchangeStreamCaptureMode (&mode));
    CK (cudaFree (dev));                /* implicit device synchronisation: nothing can still read it */
    CK (cudaThreadExchangeStreamCaptureMode (&mode));
}

/* Looks up / uploads one axis table.  Caller holds g_lock, device is current. */
static TabEntry *
tab_acquire_locked (DeviceState *ds, const AxisPlan *ap)
{
    TabEntry *free_slot = NULL, *victim = NULL;
    int i;

    for (i = 0; i < SMOL_TAB_CACHE_MAX; i++)
    {
        TabEntry *t = &ds->tabs[i];

        if (!t->used)
        {
            if (!free_slot)
                free_slot = t;
            continue;
        }
        if (t->kind == ap->kind && t->filter == (uint32_t) ap->filter
            && t->dim_in == ap->dim_in && t->n_entries == ap->n_entries)
        {
            t->refs++;
            t->stamp = ++ds->clock;
            return t;
        }
        if (t->refs == 0 && (!victim || t->stamp < victim->stamp))
            victim = t;
    }

    if (!free_slot)
    {
        if (victim)
        {
            device_release (victim->dev);
            victim->used = 0;
            free_slot = victim;
        }
        else
        {
            /* every slot is pinned by a live context (the reference has no limit on those): this
             * table lives outside the cache and goes away with its last reference */
            free_slot = calloc (1, sizeof (*free_slot));
            if (!free_slot)
                smol_fatal ("out of memory", NULL);
            free_slot->uncached = 1;
        }
    }

    free_slot->used = 1;
    free_slot->kind = ap->kind;
    free_slot->filter = (uint32_t) ap->filter;
    free_slot->dim_in = ap->dim_in;
    free_slot->n_entries = ap->n_entries;
    free_slot->refs = 1;
    free_slot->stamp = ++ds->clock;
    free_slot->dev = device_upload (ds, ap->dev_entries, (size_t) ap->n_entries * sizeof (uint32_t));
    __atomic_add_fetch (&g_stat_uploads, 1, __ATOMIC_RELAXED);
    return free_slot;
}

/* Caller holds g_lock.  `dev` is the ordinal the table lives on. */
static void
tab_release_locked (TabEntry *t, int dev)
{
    if (--t->refs == 0 && t->uncached)
    {
        int prev = -1;

        CK (cudaGetDevice (&prev));
        if (prev != dev)
            CK (cudaSetDevice (dev));
        device_release (t->dev);
        if (prev != dev)
            CK (cudaSetDevice (prev));
        free (t);
    }
}

static Lane *
lane_acquire (int dev)
{
    Lane *lane = NULL;
    DeviceState *ds;
    int i;

    pthread_mutex_lock (&g_lock);
    ds = device_state_locked (dev);
    for (;;)
    {
        for (i = 0; i < ds->n_lanes; i++)
            if (!ds->lanes[i].busy)
            {
                lane = &ds->lanes[i];
                break;
            }
        if (lane)
            break;
        if (ds->n_lanes < SMOL_MAX_LANES)
        {
            lane = &ds->lanes[ds->n_lanes++];
            memset (lane, 0, sizeof (*lane));
            CK (cudaStreamCreateWithFlags (&lane->stream, cudaStreamNonBlocking));
            break;
        }
        pthread_cond_wait (&g_lane_free, &g_lock);
    }
    lane->busy = 1;
    pthread_mutex_unlock (&g_lock);
    return lane;
}

static void
lane_release (Lane *lane)
{
    pthread_mutex_lock (&g_lock);
    lane->busy = 0;
    pthread_cond_signal (&g_lane_free);
    pthread_mutex_unlock (&g_lock);
}

static void
lane_reserve (void **buf, size_t *cap, size_t need)
{
    if (need <= *cap)
        return;
    if (*buf)
        CK (cudaFree (*buf));
    need = (need + ((size_t) 1 << 20) - 1) & ~(((size_t) 1 << 20) - 1);
    CK (cudaMalloc (buf, need));
    *cap = need;
}

/* ------------------------------------------------------------------------------------------ *
 * Context                                                                                    *
 * ------------------------------------------------------------------------------------------ */

/* Plans are immutable and depend only on (types, dimensions, sRGB flag), so they are shared:
 * a small LRU cache makes repeated smol_scale_simple / smol_scale_new calls on the same geometry
 * (a stream of video frames, a batch of same-sized thumbnails) skip table construction and keeps
 * their device tables resident. */
typedef struct
{
    int used;
    SmolPixelType type_in, type_out;
    uint32_t w_in, h_in, w_out, h_out;
    uint8_t with_srgb;
    int refs;
    uint64_t stamp;
    JobPlan plan;
    TabEntry *tab_x[SMOL_MAX_DEVICES], *tab_y[SMOL_MAX_DEVICES];
}
SharedPlan;

static SharedPlan *g_plans[SMOL_PLAN_CACHE_MAX];
static uint64_t g_plan_clock;

static void
shared_plan_drop_locked (SharedPlan *sp)
{
    int i;

    for (i = 0; i < SMOL_MAX_DEVICES; i++)
    {
        if (sp->tab_x[i])
            tab_release_locked (sp->tab_x[i], i);
        if (sp->tab_y[i])
            tab_release_locked (sp->tab_y[i], i);
    }
    job_plan_clear (&sp->plan);
    free (sp);
}

static SharedPlan *
shared_plan_acquire (SmolPixelType type_in, uint32_t w_in, uint32_t h_in,
                     SmolPixelType type_out, uint32_t w_out, uint32_t h_out, uint8_t with_srgb)
{
    SharedPlan *sp = NULL;
    int i, free_slot = -1, victim = -1;

    with_srgb = with_srgb ? 1 : 0;
    pthread_mutex_lock (&g_lock);
    for (i = 0; i < SMOL_PLAN_CACHE_MAX; i++)
    {
        SharedPlan *c = g_plans[i];

        if (!c)
        {
            if (free_slot < 0)
                free_slot = i;
            continue;
        }
        if (c->w_in == w_in && c->h_in == h_in && c->w_out == w_out && c->h_out == h_out
            && c->type_in == type_in && c->type_out == type_out && c->with_srgb == with_srgb)
        {
            sp = c;
            break;
        }
        if (c->refs == 0 && (victim < 0 || c->stamp < g_plans[victim]->stamp))
            victim = i;
    }
    if (!sp)
    {
        if (free_slot < 0 && victim >= 0)
        {
            shared_plan_drop_locked (g_plans[victim]);
            g_plans[victim] = NULL;
            free_slot = victim;
        }
        sp = calloc (1, sizeof (*sp));
        if (!sp)
            smol_fatal ("out of memory", NULL);
        sp->type_in = type_in; sp->type_out = type_out;
        sp->w_in = w_in; sp->h_in = h_in; sp->w_out = w_out; sp->h_out = h_out;
        sp->with_srgb = with_srgb;
        job_plan_init (&sp->plan, type_in, w_in, h_in, type_out, w_out, h_out, with_srgb);
        if (free_slot >= 0)
        {
            sp->used = 1;               /* cached */
            g_plans[free_slot] = sp;
        }
        /* else: every slot is pinned by a live context; this plan lives and dies with its context */
    }
    sp->refs++;
    sp->stamp = ++g_plan_clock;
    pthread_mutex_unlock (&g_lock);
    return sp;
}

static void
shared_plan_release (SharedPlan *sp)
{
    pthread_mutex_lock (&g_lock);
    sp->refs--;
    if (sp->refs == 0 && !sp->used)
        shared_plan_drop_locked (sp);
    pthread_mutex_unlock (&g_lock);
}

struct SmolScaleCtx
{
    const char *pixels_in;
    char *pixels_out;
    uint32_t rowstride_in, rowstride_out;
    SmolPixelType pixel_type_in, pixel_type_out;
    SmolPostRowFunc *post_row_func;
    void *user_data;

    SharedPlan *sp;
};

static SmolScaleCtx *
ctx_new (const void *pixels_in, SmolPixelType pixel_type_in,
         uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
         void *pixels_out, SmolPixelType pixel_type_out,
         uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
         uint8_t with_srgb, SmolPostRowFunc post_row_func, void *user_data)
{
    SmolScaleCtx *ctx = calloc (1, sizeof (*ctx));

    if (!ctx)
        smol_fatal ("out of memory", NULL);
    if (width_in == 0 || height_in == 0 || width_out == 0 || height_out == 0
        || width_in > 65535 || height_in > 65535 || width_out > 65535 || height_out > 65535)
        smol_fatal ("image dimensions must be in [1, 65535]", NULL);
    (void) type_info (pixel_type_in);
    (void) type_info (pixel_type_out);

    ctx->pixels_in = pixels_in;
    ctx->pixels_out = pixels_out;
    ctx->rowstride_in = rowstride_in;
    ctx->rowstride_out = rowstride_out;
    ctx->pixel_type_in = pixel_type_in;
    ctx->pixel_type_out = pixel_type_out;
    ctx->post_row_func = post_row_func;
    ctx->user_data = user_data;
    ctx->sp = shared_plan_acquire (pixel_type_in, width_in, height_in,
                                   pixel_type_out, width_out, height_out, with_srgb);
    return ctx;
}

static void
ctx_free (SmolScaleCtx *ctx)
{
    shared_plan_release (ctx->sp);
    free (ctx);
}

/* Device must be current. */
static void
ctx_device_tables (SmolScaleCtx *ctx, int dev, SmolLaunch *L)
{
    const uint32_t **tx = &L->tab_x, **ty = &L->tab_y;
    const SmolDeviceLuts **luts = &L->luts;
    SharedPlan *sp = ctx->sp;
    TabEntry *ex = __atomic_load_n (&sp->tab_x[dev], __ATOMIC_ACQUIRE);
    TabEntry *ey = __atomic_load_n (&sp->tab_y[dev], __ATOMIC_ACQUIRE);

    if (!ex || !ey)
    {
        DeviceState *ds;

        pthread_mutex_lock (&g_lock);
        ds = device_state_locked (dev);
        if (!sp->tab_y[dev])
        {
            ex = tab_acquire_locked (ds, &sp->plan.ax);
            ey = tab_acquire_locked (ds, &sp->plan.ay);
            __atomic_store_n (&sp->tab_x[dev], ex, __ATOMIC_RELEASE);
            __atomic_store_n (&sp->tab_y[dev], ey, __ATOMIC_RELEASE);
        }
        ex = sp->tab_x[dev];
        ey = sp->tab_y[dev];
        pthread_mutex_unlock (&g_lock);
    }
    *tx = ex->dev;
    *ty = ey->dev;
    *luts = g_dev[dev].luts;
    L->p8l_from_p = g_dev[dev].p8l_from_p;
    L->p8l_from_u = g_dev[dev].p8l_from_u;
}

/* ------------------------------------------------------------------------------------------ *
 * Row rendering                                                                              *
 * ------------------------------------------------------------------------------------------ */

/* Where a caller's buffer lives. */
enum { SMOL_MEM_PAGEABLE = 0, SMOL_MEM_PINNED = 1, SMOL_MEM_DEVICE = 2 };

typedef struct
{
    int is_device;      /* device or managed: a kernel can dereference it */
    int device;
    int mem;            /* SMOL_MEM_* */
}
PtrClass;

static PtrClass
classify_pointer (const void *p)
{
    struct cudaPointerAttributes attr;
    PtrClass pc = { 0, -1, SMOL_MEM_PAGEABLE };
    cudaError_t e;

    if (!p)
        return pc;
    e = cudaPointerGetAttributes (&attr, p);
    if (e != cudaSuccess)
    {
        (void) cudaGetLastError ();
        return pc;
    }
    if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)
    {
        pc.is_device = 1;
        pc.device = attr.device;
        pc.mem = SMOL_MEM_DEVICE;
    }
    else if (attr.type == cudaMemoryTypeHost)
        pc.mem = SMOL_MEM_PINNED;
    return pc;
}

static size_t
align16 (size_t v)
{
    return (v + 15) & ~(size_t) 15;
}

static void
launch_checked (const SmolLaunch *L, cudaStream_t stream)
{
    int kid = smol_cuda_pick_kernel (L, g_forced_kernel);
    const char *name = NULL;
    int e = smol_cuda_launch (L, kid, stream, &name);

    if (e != 0)
        smol_fatal ("kernel launch failed", cudaGetErrorString ((cudaError_t) e));
    __atomic_add_fetch (&g_stat_launches, 1, __ATOMIC_RELAXED);
    /* the family that actually ran (the launcher falls back to the general kernel) */
    for (int k = 0; k < SMOL_KERNEL_MAX && k < SMOL_CUDA_MAX_KERNEL_FAMILIES; k++)
    {
        if (name == smol_cuda_kernel_name (k))
        {
            __atomic_add_fetch (&g_stat_by_kernel[k], 1, __ATOMIC_RELAXED);
            break;
        }
    }
}

/* 2D copy that also tolerates pitches smaller than the row width (rows then overlap in the
 * pitched buffer, which the reference handles naturally because it only ever walks rows). */
static void
copy_rows_async (void *dst, size_t dpitch, const void *src, size_t spitch,
                 size_t width_bytes, size_t n_rows, enum cudaMemcpyKind kind, cudaStream_t stream)
{
    if (n_rows == 0 || width_bytes == 0)
        return;
    if (dpitch == width_bytes && spitch == width_bytes)
    {
        CK (cudaMemcpyAsync (dst, src, width_bytes * n_rows, kind, stream));
    }
    else if (dpitch >= width_bytes && spitch >= width_bytes)
    {
        CK (cudaMemcpy2DAsync (dst, dpitch, src, spitch, width_bytes, n_rows, kind, stream));
    }
    else
    {
        size_t r;

        for (r = 0; r < n_rows; r++)
            CK (cudaMemcpyAsync ((char *) dst + r * dpitch, (const char *) src + r * spitch,
                                 width_bytes, kind, stream));
    }
}

/* ---- worker pool -------------------------------------------------------------------------- *
 * A few persistent host threads shared by everything in this file that wants more than the
 * calling thread: bouncing pageable caller memory through pinned buffers (one memcpy thread
 * moves ~10 GB/s, a PCIe 5 x16 link 55), and driving one device each when a host-memory call is
 * split across several GPUs.  pool_run() runs task 0 on the calling thread and, while waiting
 * for the others, helps with whatever is queued -- so nested use cannot deadlock. */

typedef struct PoolGroup { int pending; } PoolGroup;

typedef struct PoolTask
{
    void (*fn) (void *);
    void *arg;
    PoolGroup *grp;
    struct PoolTask *next;
}
PoolTask;

#define SMOL_POOL_MAX 32

static pthread_mutex_t g_pool_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_pool_cond = PTHREAD_COND_INITIALIZER;
static PoolTask *g_pool_head, *g_pool_tail;
static int g_pool_threads, g_pool_limit = -1;

static void
pool_finish_locked (PoolTask *t)
{
    if (--t->grp->pending == 0)
        pthread_cond_broadcast (&g_pool_cond);
}

static void *
pool_main (void *unused)
{
    (void) unused;
    pthread_mutex_lock (&g_pool_lock);
    for (;;)
    {
        PoolTask *t;

        while (!g_pool_head)
            pthread_cond_wait (&g_pool_cond, &g_pool_lock);
        t = g_pool_head;
        g_pool_head = t->next;
        if (!g_pool_head)
            g_pool_tail = NULL;
        pthread_mutex_unlock (&g_pool_lock);
        t->fn (t->arg);
        pthread_mutex_lock (&g_pool_lock);
        pool_finish_locked (t);
    }
    return NULL;
}

static int
pool_limit (void)
{
    if (g_pool_limit < 0)
    {
        const char *e = getenv ("SMOL_CUDA_HOST_THREADS");
        long n = e ? atol (e) : sysconf (_SC_NPROCESSORS_ONLN) / 2;

        if (n < 1)
            n = 1;
        if (n > SMOL_POOL_MAX)
            n = SMOL_POOL_MAX;
        g_pool_limit = (int) n;
    }
    return g_pool_limit;
}

/* Runs fn (args + i * arg_size) for i in [0, n) and returns when all are done. */
static void
pool_run (void (*fn) (void *), void *args, size_t arg_size, int n)
{
    PoolTask tasks[SMOL_POOL_MAX * 2];
    PoolGroup grp;
    int i;

    if (n > SMOL_POOL_MAX * 2)
        smol_fatal ("pool_run: too many tasks", NULL);
    if (n <= 1)
    {
        if (n == 1)
            fn (args);
        return;
    }
    grp.pending = n - 1;
    pthread_mutex_lock (&g_pool_lock);
    while (g_pool_threads < n - 1 && g_pool_threads < pool_limit ())
    {
        pthread_t th;
        pthread_attr_t at;

        pthread_attr_init (&at);
        pthread_attr_setdetachstate (&at, PTHREAD_CREATE_DETACHED);
        if (pthread_create (&th, &at, pool_main, NULL) != 0)
        {
            pthread_attr_destroy (&at);
            break;
        }
        pthread_attr_destroy (&at);
        g_pool_threads++;
    }
    for (i = 1; i < n; i++)
    {
        PoolTask *t = &tasks[i];

        t->fn = fn;
        t->arg = (char *) args + (size_t) i * arg_size;
        t->grp = &grp;
        t->next = NULL;
        if (g_pool_tail)
            g_pool_tail->next = t;
        else
            g_pool_head = t;
        g_pool_tail = t;
    }
    pthread_cond_broadcast (&g_pool_cond);
    pthread_mutex_unlock (&g_pool_lock);

    fn (args);

    pthread_mutex_lock (&g_pool_lock);
    while (grp.pending > 0)
    {
        PoolTask *t = g_pool_head;

        if (t)
        {
            g_pool_head = t->next;
            if (!g_pool_head)
                g_pool_tail = NULL;
            pthread_mutex_unlock (&g_pool_lock);
            t->fn (t->arg);
            pthread_mutex_lock (&g_pool_lock);
            pool_finish_locked (t);
        }
        else
            pthread_cond_wait (&g_pool_cond, &g_pool_lock);
    }
    pthread_mutex_unlock (&g_pool_lock);
}

typedef struct
{
    char *dst;
    const char *src;
    size_t dpitch, spitch, width_bytes, n_rows;
}
RowCopy;

/* Bounce copies move tens of megabytes that the copying core never looks at again (the next
 * reader is the DMA engine, or the caller much later), in tasks too small for memcpy's own
 * non-temporal threshold: stream the stores past the cache so the destination lines are not read
 * first (2 instead of 3 bytes of memory traffic per byte copied).  SMOL_CUDA_BOUNCE_NT=0: memcpy. */
static int g_bounce_nt = 1;

#if defined (__SSE2__)
static void
copy_streaming (char *dst, const char *src, size_t n)
{
    size_t head;

    if (!g_bounce_nt || n < 1024)
    {
        memcpy (dst, src, n);
        return;
    }
    head = (16 - ((uintptr_t) dst & 15)) & 15;
    if (head)
    {
        memcpy (dst, src, head);
        dst += head; src += head; n -= head;
    }
    for (; n >= 64; n -= 64, src += 64, dst += 64)
    {
        const __m128i a = _mm_loadu_si128 ((const __m128i *) src), b = _mm_loadu_si128 ((const __m128i *) (src + 16));
        const __m128i c = _mm_loadu_si128 ((const __m128i *) (src + 32)), d = _mm_loadu_si128 ((const __m128i *) (src + 48));

        _mm_stream_si128 ((__m128i *) dst, a);
        _mm_stream_si128 ((__m128i *) (dst + 16), b);
        _mm_stream_si128 ((__m128i *) (dst + 32), c);
        _mm_stream_si128 ((__m128i *) (dst + 48), d);
    }
    if (n)
        memcpy (dst, src, n);
}
#define copy_streaming_fence() _mm_sfence ()
#else
#define copy_streaming(dst, src, n) memcpy (dst, src, n)
#define copy_streaming_fence() ((void) 0)
#endif

static void
row_copy_task (void *arg)
{
    const RowCopy *c = arg;
    size_t r;

    if (c->dpitch == c->width_bytes && c->spitch == c->width_bytes)
        copy_streaming (c->dst, c->src, c->width_bytes * c->n_rows);
    else
        for (r = 0; r < c->n_rows; r++)
            copy_streaming (c->dst + r * c->dpitch, c->src + r * c->spitch, c->width_bytes);
    copy_streaming_fence ();
}

static long
env_long (const char *name, long dflt)
{
    const char *e = getenv (name);

    return e && *e ? atol (e) : dflt;
}

/* Tunables of the pageable-memory path (environment, read once): SMOL_CUDA_BOUNCE=0 hands
 * pageable pointers straight to cudaMemcpyAsync (the driver's own staging) instead of the pinned
 * bounce buffers; _BAND_KB / _TASK_KB set the pipeline's band size and the host-copy task size. */
static int g_bounce = -1;
static size_t g_bounce_band, g_bounce_task;

static void
bounce_config (void)
{
    if (__atomic_load_n (&g_bounce, __ATOMIC_ACQUIRE) >= 0)
        return;
    g_bounce_band = (size_t) env_long ("SMOL_CUDA_BOUNCE_BAND_KB", 4096) << 10;
    g_bounce_task = (size_t) env_long ("SMOL_CUDA_BOUNCE_TASK_KB", 512) << 10;
    if (g_bounce_band < 65536)
        g_bounce_band = 65536;
    if (g_bounce_task < 16384)
        g_bounce_task = 16384;
    g_bounce_nt = env_long ("SMOL_CUDA_BOUNCE_NT", 1) != 0;
    __atomic_store_n (&g_bounce, env_long ("SMOL_CUDA_BOUNCE", 1) != 0, __ATOMIC_RELEASE);
}

/* Host-to-host row copy spread over the pool. */
static void
copy_rows_host (void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t n_rows)
{
    RowCopy parts[SMOL_POOL_MAX];
    size_t total = width_bytes * n_rows, per, r;
    int n = (int) (total / g_bounce_task), i;

    if (n_rows == 0 || width_bytes == 0)
        return;
    if (n > pool_limit ())
        n = pool_limit ();
    if ((size_t) n > n_rows)
        n = (int) n_rows;
    if (n < 1)
        n = 1;
    per = (n_rows + n - 1) / n;
    for (i = 0, r = 0; i < n && r < n_rows; i++, r += per)
    {
        parts[i].dst = (char *) dst + r * dpitch;
        parts[i].src = (const char *) src + r * spitch;
        parts[i].dpitch = dpitch;
        parts[i].spitch = spitch;
        parts[i].width_bytes = width_bytes;
        parts[i].n_rows = n_rows - r < per ? n_rows - r : per;
    }
    pool_run (row_copy_task, parts, sizeof (parts[0]), i);
}

static void
lane_reserve_pinned (void **buf, size_t *cap, size_t need)
{
    if (need <= *cap)
        return;
    if (*buf)
        CK (cudaFreeHost (*buf));
    need = (need + ((size_t) 1 << 20) - 1) & ~(((size_t) 1 << 20) - 1);
    CK (cudaHostAlloc (buf, need, cudaHostAllocDefault));
    *cap = need;
}

/* Source rows a band must have on the device before its kernel runs: what it reads, plus the
 * row below for the copy filter -- the taps kernels fetch row r + 1 (clamped to the image) with
 * weight zero rather than branch on it, so it has to be addressable. */
static void
plan_staged_rows (const JobPlan *jp, uint32_t first, uint32_t n, uint32_t *r0, uint32_t *nr)
{
    plan_source_rows (jp, first, n, r0, nr);
    if (n > 0 && jp->ay.kind == SMOL_AXIS_TAPS && jp->ay.filter != SMOL_CUDA_AXIS_BILINEAR
        && *r0 + *nr < jp->d.h_in)
        (*nr)++;
}

/* One band of output rows of a call that involves host memory (or a row callback), rendered on
 * device `dev`: stage the source rows the band reads, run the kernel(s), bring the rows back,
 * wait.  Large bands run as a pipeline of sub-bands so H2D, kernel and D2H overlap. */
static void
render_staged (SmolScaleCtx *ctx, int dev, void *outrows_dest, uint32_t first_row, uint32_t n_rows,
               PtrClass pc_in, PtrClass pc_out, cudaStream_t caller_stream)
{
    const SmolJobDesc *d = &ctx->sp->plan.d;
    const size_t in_row_bytes = (size_t) d->w_in * d->bpp_in;
    const size_t out_row_bytes = (size_t) d->w_out * d->bpp_out;
    const size_t in_pitch = align16 (in_row_bytes), out_pitch = align16 (out_row_bytes);
    const int bounce_in = (bounce_config (), g_bounce) && !pc_in.is_device && pc_in.mem == SMOL_MEM_PAGEABLE;
    const int bounce_out = g_bounce && !pc_out.is_device && pc_out.mem == SMOL_MEM_PAGEABLE;
    SmolLaunch L;
    Lane *lane;
    cudaStream_t s;
    size_t staged_bytes = 0;
    uint32_t n_bands = 1, rows_per_band, b, r0, nr;
    int prev_dev = -1;

    CK (cudaGetDevice (&prev_dev));
    if (dev != prev_dev)
        CK (cudaSetDevice (dev));

    memset (&L, 0, sizeof (L));
    L.d = *d;
    L.n_images = 1;
    ctx_device_tables (ctx, dev, &L);

    lane = lane_acquire (dev);
    /* A caller's device buffer is ordered by the caller's stream (whatever produced the input
     * was enqueued there, whatever consumes the output will be): the kernel goes on that stream.
     * Pure host-memory calls use the lane's own stream so concurrent callers overlap. */
    s = (pc_in.is_device || pc_out.is_device) ? caller_stream : lane->stream;

    plan_staged_rows (&ctx->sp->plan, first_row, n_rows, &r0, &nr);

    if (pc_in.is_device)
    {
        L.src = (const uint8_t *) ctx->pixels_in;
        L.src_pitch = ctx->rowstride_in;
    }
    else
    {
        lane_reserve (&lane->d_in, &lane->d_in_cap, in_pitch * nr + 16);
        /* the kernel addresses rows from row 0 of the image; only rows [r0, r0 + nr) are read */
        L.src = (const uint8_t *) lane->d_in - (size_t) r0 * in_pitch;
        L.src_pitch = (uint32_t) in_pitch;
        staged_bytes += in_row_bytes * nr;
        if (bounce_in)
            lane_reserve_pinned (&lane->h_in, &lane->h_in_cap, in_row_bytes * nr);
    }
    if (pc_out.is_device)
    {
        L.dst = (uint8_t *) outrows_dest;
        L.dst_pitch = ctx->rowstride_out;
    }
    else
    {
        lane_reserve (&lane->d_out, &lane->d_out_cap, out_pitch * n_rows + 16);
        L.dst = (uint8_t *) lane->d_out;
        L.dst_pitch = (uint32_t) out_pitch;
        staged_bytes += out_row_bytes * n_rows;
        if (bounce_out)
            lane_reserve_pinned (&lane->h_out, &lane->h_out_cap, out_row_bytes * n_rows);
    }

    /* Large host-memory jobs run as a pipeline of row bands: while band b is being scaled,
     * band b + 1's source rows are already crossing PCIe and band b - 1's output rows are on
     * their way back (H2D and D2H overlap: the link is full duplex).  Bands share the staged
     * source image, so each source row is uploaded exactly once.  Pageable caller memory is
     * bounced through the lane's pinned buffers by the worker pool, band by band, so the host
     * copies overlap the transfers too. */
    if (staged_bytes >= ((size_t) 4 << 20) && n_rows >= 2 * SMOL_MAX_BANDS)
    {
        n_bands = (uint32_t) ((bounce_in || bounce_out) ? staged_bytes / g_bounce_band : staged_bytes >> 22);
        if (n_bands > ((bounce_in || bounce_out) ? SMOL_MAX_BANDS : SMOL_MAX_BANDS / 2))
            n_bands = (bounce_in || bounce_out) ? SMOL_MAX_BANDS : SMOL_MAX_BANDS / 2;
        if (n_bands < 2)
            n_bands = 2;
    }
    rows_per_band = (n_rows + n_bands - 1) / n_bands;

    if ((n_bands > 1 || bounce_out) && !lane->pipeline_ready)
    {
        CK (cudaStreamCreateWithFlags (&lane->s_h2d, cudaStreamNonBlocking));
        CK (cudaStreamCreateWithFlags (&lane->s_d2h, cudaStreamNonBlocking));
        for (b = 0; b < SMOL_MAX_BANDS; b++)
        {
            CK (cudaEventCreateWithFlags (&lane->ev_up[b], cudaEventDisableTiming));
            CK (cudaEventCreateWithFlags (&lane->ev_done[b], cudaEventDisableTiming));
            CK (cudaEventCreateWithFlags (&lane->ev_down[b], cudaEventDisableTiming | cudaEventBlockingSync));
        }
        lane->pipeline_ready = 1;
    }

    {
        const uint8_t *dst_base = L.dst;
        uint32_t uploaded_end = r0;     /* source rows [r0, uploaded_end) are on the device */
        uint32_t band_first_of[SMOL_MAX_BANDS], band_rows_of[SMOL_MAX_BANDS];
        uint32_t n_done = 0, drained = 0;
        cudaStream_t s_up = n_bands > 1 ? lane->s_h2d : s;
        cudaStream_t s_down = n_bands > 1 ? lane->s_d2h : s;

        for (b = 0; b < n_bands; b++)
        {
            const uint32_t band_first = first_row + b * rows_per_band;
            uint32_t band_rows, br0, bnr;

            if (band_first >= first_row + n_rows)
                break;
            band_rows = first_row + n_rows - band_first;
            if (band_rows > rows_per_band)
                band_rows = rows_per_band;
            band_first_of[b] = band_first;
            band_rows_of[b] = band_rows;
            n_done = b + 1;

            if (!pc_in.is_device)
            {
                plan_staged_rows (&ctx->sp->plan, band_first, band_rows, &br0, &bnr);
                if (br0 + bnr > uploaded_end)
                {
                    const uint32_t from = uploaded_end, cnt = br0 + bnr - uploaded_end;
                    const char *rows = ctx->pixels_in + (size_t) from * ctx->rowstride_in;
                    size_t rows_pitch = ctx->rowstride_in;

                    if (bounce_in)
                    {
                        char *bounce = (char *) lane->h_in + (size_t) (from - r0) * in_row_bytes;

                        copy_rows_host (bounce, in_row_bytes, rows, rows_pitch, in_row_bytes, cnt);
                        rows = bounce;
                        rows_pitch = in_row_bytes;
                    }
                    copy_rows_async ((char *) lane->d_in + (size_t) (from - r0) * in_pitch, in_pitch,
                                     rows, rows_pitch, in_row_bytes, cnt, cudaMemcpyHostToDevice, s_up);
                    __atomic_add_fetch (&g_stat_h2d, in_row_bytes * cnt, __ATOMIC_RELAXED);
                    uploaded_end = br0 + bnr;
                }
                if (n_bands > 1)
                {
                    CK (cudaEventRecord (lane->ev_up[b], s_up));
                    CK (cudaStreamWaitEvent (s, lane->ev_up[b], 0));
                }
            }

            L.first_row = band_first;
            L.n_rows = band_rows;
            L.dst = (uint8_t *) dst_base + (size_t) (band_first - first_row) * L.dst_pitch;
            launch_checked (&L, s);

            if (!pc_out.is_device)
            {
                char *to = (char *) outrows_dest + (size_t) (band_first - first_row) * ctx->rowstride_out;
                size_t to_pitch = ctx->rowstride_out;

                if (n_bands > 1)
                {
                    CK (cudaEventRecord (lane->ev_done[b], s));
                    CK (cudaStreamWaitEvent (s_down, lane->ev_done[b], 0));
                }
                if (bounce_out)
                {
                    to = (char *) lane->h_out + (size_t) (band_first - first_row) * out_row_bytes;
                    to_pitch = out_row_bytes;
                }
                copy_rows_async (to, to_pitch, L.dst, L.dst_pitch,
                                 out_row_bytes, band_rows, cudaMemcpyDeviceToHost, s_down);
                __atomic_add_fetch (&g_stat_d2h, out_row_bytes * band_rows, __ATOMIC_RELAXED);
                if (bounce_out)
                    CK (cudaEventRecord (lane->ev_down[b], s_down));
            }

            /* pageable destination: hand over bands that have already landed (two behind) */
            if (bounce_out)
                for (; drained + 2 <= b; drained++)
                {
                    CK (cudaEventSynchronize (lane->ev_down[drained]));
                    copy_rows_host ((char *) outrows_dest + (size_t) (band_first_of[drained] - first_row) * ctx->rowstride_out,
                                    ctx->rowstride_out,
                                    (char *) lane->h_out + (size_t) (band_first_of[drained] - first_row) * out_row_bytes,
                                    out_row_bytes, out_row_bytes, band_rows_of[drained]);
                }
        }
        if (bounce_out)
            for (; drained < n_done; drained++)
            {
                CK (cudaEventSynchronize (lane->ev_down[drained]));
                copy_rows_host ((char *) outrows_dest + (size_t) (band_first_of[drained] - first_row) * ctx->rowstride_out,
                                ctx->rowstride_out,
                                (char *) lane->h_out + (size_t) (band_first_of[drained] - first_row) * out_row_bytes,
                                out_row_bytes, out_row_bytes, band_rows_of[drained]);
            }
        if (n_bands > 1)
            CK (cudaStreamSynchronize (s_down));
        CK (cudaStreamSynchronize (s));
    }

    lane_release (lane);
    if (dev != prev_dev)
        CK (cudaSetDevice (prev_dev));
}

/* ---- splitting one host-memory call across several GPUs ------------------------------------ */

static int g_multi_gpu = -1;        /* devices a host-memory call may be spread over (1 = off) */

static int
multi_gpu_devices (void)
{
    int n = __atomic_load_n (&g_multi_gpu, __ATOMIC_RELAXED);

    if (n < 0)
    {
        const char *e = getenv ("SMOL_CUDA_MULTI_GPU");

        n = 1;
        if (e && *e)
            n = (strcmp (e, "all") == 0) ? SMOL_MAX_DEVICES : atoi (e);
        if (n < 1)
            n = 1;
        __atomic_store_n (&g_multi_gpu, n, __ATOMIC_RELAXED);
    }
    return n < g_device_count ? n : g_device_count;
}

typedef struct
{
    SmolScaleCtx *ctx;
    int dev;
    void *dest;
    uint32_t first_row, n_rows;
    PtrClass pc_in, pc_out;
}
DeviceBand;

static void
device_band_task (void *arg)
{
    DeviceBand *t = arg;

    render_staged (t->ctx, t->dev, t->dest, t->first_row, t->n_rows, t->pc_in, t->pc_out, NULL);
}

static void
do_rows (const SmolScaleCtx *cctx, void *outrows_dest, uint32_t first_row, uint32_t n_rows)
{
    SmolScaleCtx *ctx = (SmolScaleCtx *) cctx;
    const SmolJobDesc *d = &ctx->sp->plan.d;
    const size_t out_row_bytes = (size_t) d->w_out * d->bpp_out;
    PtrClass pc_in, pc_out;
    int prev_dev = -1, dev;

    if (n_rows == 0)
        return;
    if ((uint64_t) first_row + n_rows > d->h_out)
        smol_fatal ("row range outside the output image", NULL);

    if (__atomic_load_n (&g_device_count, __ATOMIC_ACQUIRE) < 0)
    {
        pthread_mutex_lock (&g_lock);
        (void) device_count_locked ();
        pthread_mutex_unlock (&g_lock);
    }
    if (g_device_count <= 0)
        smol_fatal ("no usable CUDA device (this library has no CPU fallback)", NULL);

    pc_in = classify_pointer (ctx->pixels_in);
    pc_out = classify_pointer (outrows_dest);

    CK (cudaGetDevice (&prev_dev));
    if (pc_out.is_device)
        dev = pc_out.device;
    else if (pc_in.is_device)
        dev = pc_in.device;
    else
        dev = tl_device >= 0 ? tl_device : prev_dev;
    if (dev < 0 || dev >= g_device_count)
        smol_fatal ("buffer lives on a device this library cannot use (more than 16 devices?)", NULL);

    if (pc_in.is_device && pc_out.is_device && !ctx->post_row_func)
    {
        /* Everything already lives on the GPU: enqueue and return (stream-ordered). */
        SmolLaunch L;

        if (dev != prev_dev)
            CK (cudaSetDevice (dev));
        memset (&L, 0, sizeof (L));
        L.d = *d;
        L.n_images = 1;
        L.first_row = first_row;
        L.n_rows = n_rows;
        ctx_device_tables (ctx, dev, &L);
        L.src = (const uint8_t *) ctx->pixels_in;
        L.src_pitch = ctx->rowstride_in;
        L.dst = (uint8_t *) outrows_dest;
        L.dst_pitch = ctx->rowstride_out;
        launch_checked (&L, (cudaStream_t) tl_stream);
        if (dev != prev_dev)
            CK (cudaSetDevice (prev_dev));
        return;
    }

    {
        /* Host memory on both sides and several GPUs allowed: every device renders a band of
         * output rows from its own upload of just the source rows that band reads, over its own
         * PCIe link (the reference's row-batch contract, smolscale.h:70-74, makes bands
         * independent).  Worth it from ~8 MB of transfers per device. */
        const size_t moved = (size_t) d->w_in * d->bpp_in * d->h_in * n_rows / d->h_out + out_row_bytes * n_rows;
        int n_dev = (!pc_in.is_device && !pc_out.is_device) ? multi_gpu_devices () : 1;

        while (n_dev > 1 && (moved / n_dev < ((size_t) 8 << 20) || n_rows / n_dev < 16))
            n_dev--;
        if (n_dev > 1)
        {
            DeviceBand bands[SMOL_MAX_DEVICES];
            const uint32_t per = (n_rows + n_dev - 1) / n_dev;
            int i, n = 0;

            for (i = 0; i < n_dev; i++)
            {
                const uint32_t at = (uint32_t) i * per;

                if (at >= n_rows)
                    break;
                bands[n].ctx = ctx;
                bands[n].dev = (dev + i) % g_device_count;
                bands[n].dest = (char *) outrows_dest + (size_t) at * ctx->rowstride_out;
                bands[n].first_row = first_row + at;
                bands[n].n_rows = n_rows - at < per ? n_rows - at : per;
                bands[n].pc_in = pc_in;
                bands[n].pc_out = pc_out;
                n++;
            }
            pool_run (device_band_task, bands, sizeof (bands[0]), n);
        }
        else
            render_staged (ctx, dev, outrows_dest, first_row, n_rows, pc_in, pc_out, (cudaStream_t) tl_stream);
    }

    if (ctx->post_row_func)
    {
        /* reference smolscale.c:502-503: once per finished row, on the calling thread */
        uint32_t i;

        if (!pc_out.is_device)
        {
            for (i = 0; i < n_rows; i++)
                ctx->post_row_func ((uint32_t *) ((char *) outrows_dest + (size_t) i * ctx->rowstride_out),
                                    (int) d->w_out, ctx->user_data);
        }
        else
        {
            /* device destination: bounce each row through host memory, on the caller's stream */
            uint32_t *tmp = malloc (align16 (out_row_bytes) + 16);
            cudaStream_t s = (cudaStream_t) tl_stream;

            if (!tmp)
                smol_fatal ("out of memory", NULL);
            if (dev != prev_dev)
                CK (cudaSetDevice (dev));
            for (i = 0; i < n_rows; i++)
            {
                char *row = (char *) outrows_dest + (size_t) i * ctx->rowstride_out;

                CK (cudaMemcpyAsync (tmp, row, out_row_bytes, cudaMemcpyDeviceToHost, s));
                CK (cudaStreamSynchronize (s));
                ctx->post_row_func (tmp, (int) d->w_out, ctx->user_data);
                CK (cudaMemcpyAsync (row, tmp, out_row_bytes, cudaMemcpyHostToDevice, s));
                CK (cudaStreamSynchronize (s));
            }
            if (dev != prev_dev)
                CK (cudaSetDevice (prev_dev));
            free (tmp);
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * Public API (reference smolscale.c:882-1008)                                                *
 * ------------------------------------------------------------------------------------------ */

SMOL_EXPORT SmolScaleCtx *
smol_scale_new (const void *pixels_in, SmolPixelType pixel_type_in,
                uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                void *pixels_out, SmolPixelType pixel_type_out,
                uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                uint8_t with_srgb)
{
    return ctx_new (pixels_in, pixel_type_in, width_in, height_in, rowstride_in,
                    pixels_out, pixel_type_out, width_out, height_out, rowstride_out,
                    with_srgb, NULL, NULL);
}

SMOL_EXPORT SmolScaleCtx *
smol_scale_new_full (const void *pixels_in, SmolPixelType pixel_type_in,
                     uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                     void *pixels_out, SmolPixelType pixel_type_out,
                     uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                     uint8_t with_srgb,
                     SmolPostRowFunc post_row_func, void *user_data)
{
    return ctx_new (pixels_in, pixel_type_in, width_in, height_in, rowstride_in,
                    pixels_out, pixel_type_out, width_out, height_out, rowstride_out,
                    with_srgb, post_row_func, user_data);
}

SMOL_EXPORT void
smol_scale_destroy (SmolScaleCtx *scale_ctx)
{
    if (scale_ctx)
        ctx_free (scale_ctx);
}

SMOL_EXPORT void
smol_scale_batch (const SmolScaleCtx *scale_ctx, uint32_t first_outrow, uint32_t n_outrows)
{
    do_rows (scale_ctx,
             scale_ctx->pixels_out + (size_t) first_outrow * scale_ctx->rowstride_out,
             first_outrow, n_outrows);
}

SMOL_EXPORT void
smol_scale_batch_full (const SmolScaleCtx *scale_ctx, void *outrows_dest,
                       uint32_t first_outrow, uint32_t n_outrows)
{
    do_rows (scale_ctx, outrows_dest, first_outrow, n_outrows);
}

SMOL_EXPORT void
smol_scale_simple (const void *pixels_in, SmolPixelType pixel_type_in,
                   uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                   void *pixels_out, SmolPixelType pixel_type_out,
                   uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                   uint8_t with_srgb)
{
    SmolScaleCtx *ctx = ctx_new (pixels_in, pixel_type_in, width_in, height_in, rowstride_in,
                                 pixels_out, pixel_type_out, width_out, height_out, rowstride_out,
                                 with_srgb, NULL, NULL);

    do_rows (ctx, pixels_out, 0, height_out);
    ctx_free (ctx);
}

/* ------------------------------------------------------------------------------------------ *
 * Extensions (smolscale-cuda.h)                                                              *
 * ------------------------------------------------------------------------------------------ */

SMOL_EXPORT int
smol_cuda_device_count (void)
{
    int n;

    pthread_mutex_lock (&g_lock);
    n = device_count_locked ();
    pthread_mutex_unlock (&g_lock);
    return n;
}

SMOL_EXPORT void
smol_cuda_set_device (int device)
{
    tl_device = device;
}

SMOL_EXPORT void
smol_cuda_set_stream (void *cuda_stream)
{
    tl_stream = cuda_stream;
}

SMOL_EXPORT void
smol_cuda_set_multi_gpu (int n_devices)
{
    if (n_devices < 1)
        n_devices = SMOL_MAX_DEVICES;       /* 0 / negative: every visible device */
    __atomic_store_n (&g_multi_gpu, n_devices, __ATOMIC_RELAXED);
}

SMOL_EXPORT void
smol_cuda_synchronize (void)
{
    CK (cudaStreamSynchronize ((cudaStream_t) tl_stream));
}

SMOL_EXPORT void
smol_cuda_scale_images (const void *pixels_in, size_t image_stride_in,
                        SmolPixelType pixel_type_in,
                        uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                        void *pixels_out, size_t image_stride_out,
                        SmolPixelType pixel_type_out,
                        uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                        uint8_t with_srgb, uint32_t n_images)
{
    SmolScaleCtx *ctx;
    PtrClass pc_in = classify_pointer (pixels_in), pc_out = classify_pointer (pixels_out);
    SmolLaunch L;
    int prev_dev = -1, dev;
    uint32_t done;

    if (n_images == 0)
        return;
    if (!pc_in.is_device || !pc_out.is_device)
        smol_fatal ("smol_cuda_scale_images needs device (or managed) memory for both buffers", NULL);

    ctx = ctx_new (pixels_in, pixel_type_in, width_in, height_in, rowstride_in,
                   pixels_out, pixel_type_out, width_out, height_out, rowstride_out,
                   with_srgb, NULL, NULL);
    CK (cudaGetDevice (&prev_dev));
    dev = pc_out.device;
    if (dev < 0 || dev >= SMOL_MAX_DEVICES)
        smol_fatal ("buffer lives on a device this library cannot use (more than 16 devices?)", NULL);
    if (dev != prev_dev)
        CK (cudaSetDevice (dev));

    memset (&L, 0, sizeof (L));
    L.d = ctx->sp->plan.d;
    L.src_pitch = rowstride_in;
    L.dst_pitch = rowstride_out;
    L.src_image_stride = image_stride_in;
    L.dst_image_stride = image_stride_out;
    L.first_row = 0;
    L.n_rows = height_out;
    ctx_device_tables (ctx, dev, &L);

    /* grid.z carries the image index and is limited to 65535 */
    for (done = 0; done < n_images; done += 65535)
    {
        L.src = (const uint8_t *) pixels_in + (size_t) done * image_stride_in;
        L.dst = (uint8_t *) pixels_out + (size_t) done * image_stride_out;
        L.n_images = n_images - done > 65535 ? 65535 : n_images - done;
        launch_checked (&L, (cudaStream_t) tl_stream);
    }

    if (dev != prev_dev)
        CK (cudaSetDevice (prev_dev));
    ctx_free (ctx);
}

SMOL_EXPORT void
smol_cuda_plan_query (SmolPixelType pixel_type_in, uint32_t width_in, uint32_t height_in,
                      SmolPixelType pixel_type_out, uint32_t width_out, uint32_t height_out,
                      uint8_t with_srgb,
                      SmolCudaPlanInfo *info, uint16_t *tab_x, uint16_t *tab_y)
{
    JobPlan jp;
    SmolLaunch L;

    job_plan_init (&jp, pixel_type_in, width_in, height_in, pixel_type_out, width_out, height_out, with_srgb);
    if (info)
    {
        memset (info, 0, sizeof (*info));
        info->filter_h = jp.ax.filter; info->filter_v = jp.ay.filter;
        info->halvings_h = jp.ax.halvings; info->halvings_v = jp.ay.halvings;
        info->bilin_w = jp.ax.bilin_dim; info->bilin_h = jp.ay.bilin_dim;
        info->storage_bits = jp.storage_bits;
        info->mid = jp.d.mid;
        info->span_mul_x = jp.ax.span_mul; info->span_mul_y = jp.ay.span_mul;
        info->n_tab_x = jp.ax.n_pairs; info->n_tab_y = jp.ay.n_pairs;
        memset (&L, 0, sizeof (L));
        L.d = jp.d;
        L.n_images = 1;
        L.n_rows = height_out;
        info->kernel_id = smol_cuda_pick_kernel (&L, g_forced_kernel);
        strncpy (info->kernel_name, smol_cuda_kernel_name (info->kernel_id), sizeof (info->kernel_name) - 1);
    }
    if (tab_x && jp.ax.pairs)
        memcpy (tab_x, jp.ax.pairs, (size_t) jp.ax.n_pairs * 2 * sizeof (uint16_t));
    if (tab_y && jp.ay.pairs)
        memcpy (tab_y, jp.ay.pairs, (size_t) jp.ay.n_pairs * 2 * sizeof (uint16_t));
    job_plan_clear (&jp);
}

SMOL_EXPORT void
smol_cuda_band_source_rows (const SmolScaleCtx *scale_ctx,
                            uint32_t first_outrow, uint32_t n_outrows,
                            uint32_t *first_inrow, uint32_t *n_inrows)
{
    plan_source_rows (&scale_ctx->sp->plan, first_outrow, n_outrows, first_inrow, n_inrows);
}

SMOL_EXPORT void
smol_cuda_get_stats (SmolCudaStats *stats)
{
    stats->kernel_launches = __atomic_load_n (&g_stat_launches, __ATOMIC_RELAXED);
    stats->h2d_bytes = __atomic_load_n (&g_stat_h2d, __ATOMIC_RELAXED);
    stats->d2h_bytes = __atomic_load_n (&g_stat_d2h, __ATOMIC_RELAXED);
    stats->table_uploads = __atomic_load_n (&g_stat_uploads, __ATOMIC_RELAXED);
}

SMOL_EXPORT void
smol_cuda_reset_stats (void)
{
    __atomic_store_n (&g_stat_launches, 0, __ATOMIC_RELAXED);
    __atomic_store_n (&g_stat_h2d, 0, __ATOMIC_RELAXED);
    __atomic_store_n (&g_stat_d2h, 0, __ATOMIC_RELAXED);
    __atomic_store_n (&g_stat_uploads, 0, __ATOMIC_RELAXED);
    for (int k = 0; k < SMOL_CUDA_MAX_KERNEL_FAMILIES; k++)
        __atomic_store_n (&g_stat_by_kernel[k], 0, __ATOMIC_RELAXED);
}

SMOL_EXPORT int
smol_cuda_get_kernel_launches (uint64_t *counts, const char **names, int max_families)
{
    int n = SMOL_KERNEL_MAX < SMOL_CUDA_MAX_KERNEL_FAMILIES ? SMOL_KERNEL_MAX : SMOL_CUDA_MAX_KERNEL_FAMILIES;

    if (n > max_families)
        n = max_families;
    for (int k = 0; k < n; k++)
    {
        counts[k] = __atomic_load_n (&g_stat_by_kernel[k], __ATOMIC_RELAXED);
        if (names)
            names[k] = smol_cuda_kernel_name (k);
    }
    return n;
}

SMOL_EXPORT void
smol_cuda_force_kernel (int kernel_id)
{
    g_forced_kernel = kernel_id;
}

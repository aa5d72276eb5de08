/* TEST INFRASTRUCTURE ONLY -- see smol_oracle.h.  Parity status: PINNED against the compiled
 * reference (tests/test_oracle.py) and the golden digests in tests/golden/.
 *
 * A plain scalar restatement of the smolscale pipeline.  Where the reference packs four 16-bit
 * or two 32-bit channel lanes into uint64_t words and filters whole words at a time
 * (smolscale-generic.c), this file keeps one uint32_t per channel and writes every step as
 * ordinary integer arithmetic on that channel, so each rounding rule is visible on its own.
 * That the two formulations agree bit-for-bit is exactly what the differential tests assert.
 *
 * Channel convention used here: v[0..2] = the three colour channels in the order they appear
 * in INPUT memory, v[3] = the alpha lane.  (The reference shuffles lanes per pixel type,
 * smolscale.c:647-719; because every filter treats all lanes alike, lane order does not change
 * any result -- the one place where the chosen repack variant matters is flagged below.) */

#include <stdlib.h>
#include <string.h>
#include "smol_oracle.h"
#include "smol_oracle_luts.h"

typedef struct { uint32_t v[4]; } opx;

/* ---- pixel type description (reference smolscale.c:45-75) ---- */

typedef struct { int bpp; int alpha_idx; int col0; int bgr; int unassoc; } otype;

static otype
type_info (int t)
{
    otype o;
    o.bpp = (t >= ORACLE_RGB8) ? 3 : 4;
    o.unassoc = (t >= ORACLE_RGBA8_U && t <= ORACLE_ABGR8_U);
    switch (t & 3)
    {
        /* RGBA, BGRA, ARGB, ABGR */
        case 0: o.alpha_idx = 3; o.col0 = 0; o.bgr = 0; break;
        case 1: o.alpha_idx = 3; o.col0 = 0; o.bgr = 1; break;
        case 2: o.alpha_idx = 0; o.col0 = 1; o.bgr = 0; break;
        default: o.alpha_idx = 0; o.col0 = 1; o.bgr = 1; break;
    }
    if (o.bpp == 3)
    {
        o.alpha_idx = -1; o.col0 = 0; o.bgr = (t == ORACLE_BGR8);
    }
    return o;
}

/* ---- filter selection (reference smolscale.c:427-478) ---- */

static void
pick_axis (uint32_t dim_in, uint32_t dim_out, uint8_t with_srgb,
           int *filter, uint32_t *halvings, uint32_t *bilin_dim, int *storage_bits)
{
    *bilin_dim = dim_out;
    *halvings = 0;
    *storage_bits = with_srgb ? 128 : 64;

    if (dim_in > dim_out * 255)
    {
        *filter = ORACLE_F_BOX;
        *storage_bits = 128;
    }
    else if (dim_in > dim_out * 8)
        *filter = ORACLE_F_BOX;
    else if (dim_in == 1)
        *filter = ORACLE_F_ONE;
    else if (dim_in == dim_out)
        *filter = ORACLE_F_COPY;
    else
    {
        uint32_t n = 0, d = dim_out;
        for (;;)
        {
            d *= 2;
            if (d >= dim_in)
                break;
            n++;
        }
        *filter = ORACLE_F_BILINEAR;
        *halvings = n;
        *bilin_dim = dim_out << n;
    }
}

/* ---- tables (reference smolscale-generic.c:14-66 and :68-135), absolute offsets ---- */

static uint16_t *
make_bilinear_table (uint32_t dim_in, uint32_t dim_out)
{
    uint16_t *tab = malloc ((size_t) dim_out * 2 * sizeof (uint16_t));
    const uint64_t one = (uint64_t) 1 << 32;
    uint64_t step, frac;
    uint32_t i;

    if (dim_in > dim_out)
    {
        step = ((uint64_t) dim_in * one) / dim_out;
        frac = (step - one) / 2;
    }
    else
    {
        step = ((uint64_t) (dim_in - 1) * one) / (dim_out > 1 ? dim_out - 1 : 1);
        frac = 0;
    }

    for (i = 0; i < dim_out; i++, frac += step)
    {
        uint16_t ofs = (uint16_t) (frac >> 32);

        /* ofs and ofs + 1 are sampled; from the first index whose neighbour would fall outside
         * the row onward, every entry means "100 % of the last pixel" (generic:42-65).  The
         * offsets are monotonic, so the per-entry test equals the reference's early break. */
        if (ofs >= dim_in - 1)
            break;
        tab[i * 2] = ofs;
        tab[i * 2 + 1] = (uint16_t) (256 - ((frac >> 24) & 255));
    }
    for (; i < dim_out; i++)
    {
        tab[i * 2] = (uint16_t) (dim_in - 2);
        tab[i * 2 + 1] = 0;
    }
    return tab;
}

static uint16_t *
make_box_table (uint32_t dim_in, uint32_t dim_out, uint32_t *span_mul)
{
    uint16_t *tab = malloc (((size_t) dim_out + 1) * 2 * sizeof (uint16_t));
    uint64_t step = ((uint64_t) dim_in * 65536) / dim_out;
    uint64_t frac = 0, stride, f, a, b;
    uint16_t ofs = 0, next_ofs;
    uint32_t i = 0;

    stride = step / 65536;
    f = (step / 256) % 256;
    a = ((uint64_t) 1 << 24) * 255;
    b = stride * 255 + (f * 255) / 256;
    *span_mul = (uint32_t) ((a + b / 2) / b);

    for (i = 0; i < dim_out; i++)
    {
        frac += step;
        next_ofs = (uint16_t) (frac / 65536);

        if (ofs >= dim_in - 1)
        {
            ofs = (uint16_t) (dim_in - 1);
            break;
        }
        if (next_ofs > dim_in - 1)
        {
            next_ofs = (uint16_t) (dim_in - 1);
            if (next_ofs <= ofs)
                break;
        }
        tab[i * 2] = ofs;
        tab[i * 2 + 1] = (uint16_t) ((frac / 256) % 256);
        ofs = next_ofs;
    }
    for (; i < dim_out; i++)
    {
        tab[i * 2] = ofs;
        tab[i * 2 + 1] = 0;
    }
    /* Sentinel pair: where the last box ends. */
    tab[dim_out * 2] = ofs;
    tab[dim_out * 2 + 1] = 0;
    return tab;
}

void
oracle_plan_init (oracle_plan *p,
                  int type_in, uint32_t w_in, uint32_t h_in,
                  int type_out, uint32_t w_out, uint32_t h_out,
                  uint8_t with_srgb)
{
    int st_h, st_v, linear;
    otype ti = type_info (type_in), to = type_info (type_out);

    memset (p, 0, sizeof (*p));
    p->w_in = w_in; p->h_in = h_in; p->w_out = w_out; p->h_out = h_out;
    p->type_in = type_in; p->type_out = type_out;

    pick_axis (w_in, w_out, with_srgb, &p->filter_h, &p->halvings_h, &p->bilin_w, &st_h);
    pick_axis (h_in, h_out, with_srgb, &p->filter_v, &p->halvings_v, &p->bilin_h, &st_v);
    p->storage_bits = st_h > st_v ? st_h : st_v;                      /* smolscale.c:862 */

    linear = with_srgb ? 1 : 0;
    if (ti.unassoc && to.unassoc)                                     /* smolscale.c:751-758 */
        p->storage_bits = 128;
    if (w_in > w_out * 8191 || h_in > h_out * 8191)                   /* smolscale.c:760-770 */
        linear = 0;

    if (ti.unassoc && to.unassoc)
        p->mid = linear ? ORACLE_MID_P16L : ORACLE_MID_P16;
    else
        p->mid = linear ? ORACLE_MID_P8L : ORACLE_MID_P8;

    if (p->filter_h == ORACLE_F_BOX)
    {
        p->tab_x = make_box_table (w_in, w_out, &p->span_mul_x);
        p->n_tab_x = w_out + 1;
    }
    else if (p->filter_h == ORACLE_F_BILINEAR)
    {
        p->tab_x = make_bilinear_table (w_in, p->bilin_w);
        p->n_tab_x = p->bilin_w;
    }
    if (p->filter_v == ORACLE_F_BOX)
    {
        p->tab_y = make_box_table (h_in, h_out, &p->span_mul_y);
        p->n_tab_y = h_out + 1;
    }
    else if (p->filter_v == ORACLE_F_BILINEAR)
    {
        p->tab_y = make_bilinear_table (h_in, p->bilin_h);
        p->n_tab_y = p->bilin_h;
    }
}

void
oracle_plan_free (oracle_plan *p)
{
    free (p->tab_x);
    free (p->tab_y);
    p->tab_x = p->tab_y = NULL;
}

/* ---- unpack: one input pixel -> intermediate (generic:349-752, helpers :185-318) ---- */

static opx
unpack_pixel (const uint8_t *b, otype ti, int mid)
{
    opx o;
    uint32_t a = ti.alpha_idx >= 0 ? b[ti.alpha_idx] : 0xff;
    int i;

    for (i = 0; i < 3; i++)
    {
        uint32_t c = b[ti.col0 + i];

        switch (mid)
        {
            case ORACLE_MID_P8:
                if (ti.unassoc)
                    c = (((c + 1) * (a + 1) - 1) >> 8) & 0xff;                 /* generic:238-244 */
                break;
            case ORACLE_MID_P8L:
                if (!ti.unassoc)
                    c = ((c * oracle_lut_inv_div_p8[a]) >> 13) & 0xff;         /* generic:227-236 */
                c = oracle_lut_from_srgb[c];                                   /* generic:185-199 */
                c = (((c + 1) * ((a << 3) + 1) - 1) >> 11) & 0x7ff;            /* generic:261-269 */
                break;
            case ORACLE_MID_P16:
                c = c * a;                                                     /* generic:616-625 */
                break;
            default: /* ORACLE_MID_P16L */
                c = oracle_lut_from_srgb[c] * a;                               /* generic:636-651 */
                break;
        }
        o.v[i] = c;
    }
    o.v[3] = (mid == ORACLE_MID_P16 || mid == ORACLE_MID_P16L) ? ((a << 8) | 0x80) : a;
    return o;
}

/* ---- pack: intermediate -> one output pixel (generic:754-1164) ---- */

static void
pack_pixel (opx in, uint8_t *b, otype ti, otype to, int mid)
{
    uint32_t a, c[3];
    int swapped = ti.bgr != to.bgr;
    int i;

    if (mid == ORACLE_MID_P16 || mid == ORACLE_MID_P16L)
        a = (in.v[3] >> 8) & 0xff;                                             /* generic:1140,1152 */
    else
        a = in.v[3] & 0xff;                                                    /* generic:876,1101 */

    for (i = 0; i < 3; i++)
    {
        uint64_t v = in.v[i];

        switch (mid)
        {
            case ORACLE_MID_P8:
                if (to.unassoc)
                    v = ((v & 0xffffffffu) * oracle_lut_inv_div_p8[a] >> 13) & 0xff;     /* generic:246-259 */
                break;
            case ORACLE_MID_P8L:
                if (to.bpp == 3)
                {
                    /* 24bpp output from linear light.  The reference has two packers that differ
                     * in behaviour (generic:922-935 vs :1010-1023): the "123" one gamma-compresses
                     * the still-premultiplied value, the "321" one unpremultiplies first; neither
                     * re-premultiplies.  Which one the repack search (smolscale.c:647-719) lands on:
                     * 32bpp input -> "123" iff colour order is reversed between input and output;
                     * 24bpp input -> "123" iff colour order is the same. */
                    int direct = (ti.bpp == 4) ? swapped : !swapped;
                    if (!direct)
                        v = (v * oracle_lut_inv_div_p8l[a] >> 10) & 0x7ff;
                    v = oracle_lut_to_srgb[v & 0x7ff];
                }
                else
                {
                    v = (v * oracle_lut_inv_div_p8l[a] >> 10) & 0x7ff;         /* generic:271-280 */
                    v = oracle_lut_to_srgb[v];                                 /* generic:201-211 */
                    if (!to.unassoc)
                        v = (((v + 1) * (a + 1) - 1) >> 8) & 0xff;             /* generic:217-225 */
                }
                break;
            case ORACLE_MID_P16:
                v = (v * oracle_lut_inv_div_p16[a] >> 16) & 0xff;              /* generic:290-299 */
                break;
            default: /* ORACLE_MID_P16L */
                v = (v * oracle_lut_inv_div_p16l[a] >> 19) & 0x7ff;            /* generic:309-318 */
                v = oracle_lut_to_srgb[v];
                break;
        }
        c[i] = (uint32_t) v & 0xff;
    }

    for (i = 0; i < 3; i++)
        b[to.col0 + i] = (uint8_t) c[swapped ? 2 - i : i];
    if (to.alpha_idx >= 0)
        b[to.alpha_idx] = (uint8_t) a;
}

/* ---- filter arithmetic, one lane at a time ---- */

/* generic:1317 and every other bilinear tap: ((((p - q) * F) >> 8) + q) & mask */
static uint32_t
lerp_lane (uint32_t p, uint32_t q, uint32_t F, uint32_t mask)
{
    int64_t d = ((int64_t) p - (int64_t) q) * (int64_t) F;
    /* floor division by 256 (arithmetic shift) */
    d = (d >= 0) ? (d >> 8) : -((-d + 255) >> 8);
    return (uint32_t) (d + (int64_t) q) & mask;
}

/* generic:1231-1261 */
static uint32_t
box_normalise (uint64_t acc, uint32_t mul, int storage_bits)
{
    if (storage_bits == 64)
        return (uint32_t) ((((acc & 0xffff) * mul + (1u << 23)) >> 24) & 0xff);
    return (uint32_t) ((((acc & 0xffffffffu) * mul + (1u << 23)) >> 24) & 0xffff);
}

/* ---- horizontal pass for one source row (generic:1290-1642) ---- */

static void
hfilter_row (const oracle_plan *p, const opx *in, opx *out)
{
    const uint32_t mask = p->storage_bits == 64 ? 0xff : 0xffffff;
    uint32_t x, l;

    if (p->filter_h == ORACLE_F_COPY)
    {
        memcpy (out, in, (size_t) p->w_out * sizeof (opx));
    }
    else if (p->filter_h == ORACLE_F_ONE)
    {
        for (x = 0; x < p->w_out; x++)
            out[x] = in[0];
    }
    else if (p->filter_h == ORACLE_F_BILINEAR)
    {
        uint32_t n = p->halvings_h, k;

        for (x = 0; x < p->w_out; x++)
        {
            uint32_t acc[4] = { 0, 0, 0, 0 };

            for (k = 0; k < (1u << n); k++)
            {
                uint32_t i = (x << n) + k;
                uint32_t ofs = p->tab_x[i * 2], F = p->tab_x[i * 2 + 1];

                for (l = 0; l < 4; l++)
                    acc[l] += lerp_lane (in[ofs].v[l], in[ofs + 1].v[l], F, mask);
            }
            for (l = 0; l < 4; l++)
                out[x].v[l] = (acc[l] >> n) & mask;
        }
    }
    else /* box, generic:1427-1556 */
    {
        for (x = 0; x < p->w_out; x++)
        {
            uint32_t L = p->tab_x[x * 2], R = p->tab_x[(x + 1) * 2];
            uint32_t F = p->tab_x[x * 2 + 1];
            uint32_t j;

            for (l = 0; l < 4; l++)
            {
                uint64_t acc = 0;
                uint32_t r = in[L].v[l];

                /* left edge: pixel 0 of the row in full, else what the previous box left over */
                if (x == 0)
                    acc += ((uint64_t) r * 256 >> 8) & mask;
                else
                    acc += (((uint64_t) r * 255 - (uint64_t) r * p->tab_x[x * 2 - 1]) >> 8) & mask;
                for (j = L + 1; j < R; j++)
                    acc += in[j].v[l];
                /* right edge (on the last box only when F > 0, generic:1472-1477) */
                if (F > 0)
                    acc += ((uint64_t) in[R].v[l] * F >> 8) & mask;
                out[x].v[l] = box_normalise (acc, p->span_mul_x, p->storage_bits);
            }
        }
    }
}

/* ---- whole-row pipeline ---- */

typedef struct
{
    const oracle_plan *p;
    const uint8_t *in;
    uint32_t stride_in;
    otype ti, to;
    opx *unpacked;      /* w_in */
    /* tiny cache of horizontally filtered rows, keyed by source row */
    opx *hrow[4];
    int64_t hrow_idx[4];
    uint32_t next_slot;
} octx;

/* keep: a row pointer handed out earlier that must stay valid (or NULL) */
static const opx *
get_hrow (octx *c, uint32_t r, const opx *keep)
{
    uint32_t i, x;
    const uint8_t *row;

    for (i = 0; i < 4; i++)
        if (c->hrow_idx[i] == (int64_t) r)
            return c->hrow[i];

    i = c->next_slot;
    if (c->hrow[i] == keep)
        i = (i + 1) & 3;
    c->next_slot = (i + 1) & 3;
    row = c->in + (size_t) c->stride_in * r;
    for (x = 0; x < c->p->w_in; x++)
        c->unpacked[x] = unpack_pixel (row + (size_t) x * c->ti.bpp, c->ti, c->p->mid);
    hfilter_row (c->p, c->unpacked, c->hrow[i]);
    c->hrow_idx[i] = r;
    return c->hrow[i];
}

void
oracle_scale_rows (const oracle_plan *p,
                   const void *pixels_in, uint32_t rowstride_in,
                   void *outrows_dest, uint32_t rowstride_out,
                   uint32_t first_row, uint32_t n_rows)
{
    const uint32_t mask = p->storage_bits == 64 ? 0xff : 0xffffff;
    octx c;
    opx *acc_row;
    uint64_t *acc64;
    uint32_t y, x, l, i;

    memset (&c, 0, sizeof (c));
    c.p = p; c.in = pixels_in; c.stride_in = rowstride_in;
    c.ti = type_info (p->type_in); c.to = type_info (p->type_out);
    c.unpacked = malloc ((size_t) p->w_in * sizeof (opx));
    for (i = 0; i < 4; i++)
    {
        c.hrow[i] = malloc ((size_t) p->w_out * sizeof (opx));
        c.hrow_idx[i] = -1;
    }
    acc_row = malloc ((size_t) p->w_out * sizeof (opx));
    acc64 = malloc ((size_t) p->w_out * 4 * sizeof (uint64_t));

    for (y = first_row; y < first_row + n_rows; y++)
    {
        uint8_t *dst = (uint8_t *) outrows_dest + (size_t) rowstride_out * (y - first_row);

        if (p->filter_v == ORACLE_F_COPY)                     /* generic:2306-2318 */
        {
            memcpy (acc_row, get_hrow (&c, y, NULL), (size_t) p->w_out * sizeof (opx));
        }
        else if (p->filter_v == ORACLE_F_ONE)                 /* generic:2262-2304 */
        {
            memcpy (acc_row, get_hrow (&c, 0, NULL), (size_t) p->w_out * sizeof (opx));
        }
        else if (p->filter_v == ORACLE_F_BILINEAR)            /* generic:1648-2007 */
        {
            uint32_t n = p->halvings_v, k;

            memset (acc_row, 0, (size_t) p->w_out * sizeof (opx));
            for (k = 0; k < (1u << n); k++)
            {
                uint32_t bi = (y << n) + k;
                uint32_t ofs = p->tab_y[bi * 2], F = p->tab_y[bi * 2 + 1];
                const opx *top = get_hrow (&c, ofs, NULL);
                const opx *bot = get_hrow (&c, ofs + 1, top);

                for (x = 0; x < p->w_out; x++)
                    for (l = 0; l < 4; l++)
                        acc_row[x].v[l] += lerp_lane (top[x].v[l], bot[x].v[l], F, mask);
            }
            for (x = 0; x < p->w_out; x++)
                for (l = 0; l < 4; l++)
                    acc_row[x].v[l] = (acc_row[x].v[l] >> n) & mask;
        }
        else /* box */
        {
            uint32_t T = p->tab_y[y * 2], B = p->tab_y[(y + 1) * 2];
            uint32_t F = p->tab_y[y * 2 + 1];
            uint32_t w1 = (y == 0) ? 256 : 255 - p->tab_y[y * 2 - 1];
            const opx *row;
            uint32_t r;

            /* first row, weighted (generic:2128, :2092 / :2218-2220) */
            row = get_hrow (&c, T, NULL);
            for (x = 0; x < p->w_out; x++)
                for (l = 0; l < 4; l++)
                    acc64[x * 4 + l] = ((uint64_t) row[x].v[l] * w1 >> 8) & mask;
            /* whole rows */
            for (r = T + 1; r < B; r++)
            {
                row = get_hrow (&c, r, NULL);
                for (x = 0; x < p->w_out; x++)
                    for (l = 0; l < 4; l++)
                        acc64[x * 4 + l] += row[x].v[l];
            }
            /* last row: 64bpp weighs it by F (generic:2129-2137, contributes 0 when F == 0);
             * 128bpp weighs it by F - 1 and skips it when F == 0 (generic:2240-2253) */
            if (F > 0)
            {
                uint32_t w2 = p->storage_bits == 64 ? F : F - 1;

                row = get_hrow (&c, B, NULL);
                for (x = 0; x < p->w_out; x++)
                    for (l = 0; l < 4; l++)
                        acc64[x * 4 + l] += ((uint64_t) row[x].v[l] * w2 >> 8) & mask;
            }
            for (x = 0; x < p->w_out; x++)
                for (l = 0; l < 4; l++)
                    acc_row[x].v[l] = box_normalise (acc64[x * 4 + l], p->span_mul_y, p->storage_bits);
        }

        for (x = 0; x < p->w_out; x++)
            pack_pixel (acc_row[x], dst + (size_t) x * c.to.bpp, c.ti, c.to, p->mid);
    }

    free (acc64);
    free (acc_row);
    for (i = 0; i < 4; i++)
        free (c.hrow[i]);
    free (c.unpacked);
}

void
oracle_scale_simple (const void *pixels_in, int type_in,
                     uint32_t w_in, uint32_t h_in, uint32_t rowstride_in,
                     void *pixels_out, int type_out,
                     uint32_t w_out, uint32_t h_out, uint32_t rowstride_out,
                     uint8_t with_srgb)
{
    oracle_plan p;

    oracle_plan_init (&p, type_in, w_in, h_in, type_out, w_out, h_out, with_srgb);
    oracle_scale_rows (&p, pixels_in, rowstride_in, pixels_out, rowstride_out, 0, h_out);
    oracle_plan_free (&p);
}

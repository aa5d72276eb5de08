/* -*- Mode: C; tab-width: 4; indent-tabs-mode: nil; c-basic-offset: 4 -*- */

/* PNG file I/O for the callers either side of the scaling path (include/smol-png.h).
 *
 * Replaces the reference's libpng helper (png.c:34-209) with a self-contained codec over zlib:
 * chunk parsing with CRC checks, inflate, the five row filters (PNG spec §9), Adam7, expansion
 * of every colour type / bit depth to RGBA8; on the way out per-row filter selection, deflate,
 * chunk framing.  Host code only: a PNG stream is a serial entropy-coded format and the work
 * either side of inflate/deflate is a few byte operations per pixel, so the GPU is used for
 * what sits between load and save (tools/smol_generate.c), not for the container format. */

#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "smol-png.h"

#define SMOL_PNG_EXPORT __attribute__ ((visibility ("default")))

static const uint8_t png_signature [8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };

#define CHUNK(a, b, c, d) (((uint32_t) (a) << 24) | ((uint32_t) (b) << 16) | ((uint32_t) (c) << 8) | (uint32_t) (d))

static uint32_t
get_be32 (const uint8_t *p)
{
    return ((uint32_t) p [0] << 24) | ((uint32_t) p [1] << 16) | ((uint32_t) p [2] << 8) | p [3];
}

static void
put_be32 (uint8_t *p, uint32_t v)
{
    p [0] = v >> 24; p [1] = v >> 16; p [2] = v >> 8; p [3] = v;
}

/* --- Row filters --- */

static inline int
paeth (int a, int b, int c)
{
    int p = a + b - c;
    int pa = abs (p - a), pb = abs (p - b), pc = abs (p - c);

    if (pa <= pb && pa <= pc)
        return a;
    return pb <= pc ? b : c;
}

/* Reverses the filter of one row in place.  prev = the reconstructed row above, or NULL on the
 * first row of an image (or interlace pass), where it counts as zeros.  step = bytes per
 * complete pixel, at least 1. */
static int
unfilter_row (uint8_t *row, const uint8_t *prev, size_t n, unsigned step, unsigned filter)
{
    size_t i;

    switch (filter)
    {
        case 0:
            break;
        case 1:
            for (i = step; i < n; i++)
                row [i] += row [i - step];
            break;
        case 2:
            if (prev)
                for (i = 0; i < n; i++)
                    row [i] += prev [i];
            break;
        case 3:
            /* the first pixel has no left neighbour; without a row above Average is half of Sub */
            for (i = 0; i < step && i < n; i++)
                row [i] += (prev ? prev [i] : 0) >> 1;
            if (prev)
                for (; i < n; i++)
                    row [i] += (row [i - step] + prev [i]) >> 1;
            else
                for (; i < n; i++)
                    row [i] += row [i - step] >> 1;
            break;
        case 4:
            /* Paeth (0, b, 0) = b on the first pixel, Paeth (a, 0, 0) = a on the first row */
            if (!prev)
            {
                for (i = step; i < n; i++)
                    row [i] += row [i - step];
                break;
            }
            for (i = 0; i < step && i < n; i++)
                row [i] += prev [i];
            for (; i < n; i++)
                row [i] += paeth (row [i - step], prev [i], prev [i - step]);
            break;
        default:
            return SMOL_PNG_ERR_CORRUPT;
    }
    return SMOL_PNG_OK;
}

/* Applies `filter` to one row of raw bytes; returns the sum of the output bytes read as signed
 * magnitudes, the selection heuristic of the PNG specification (§12.8). */
static uint64_t
filter_row (uint8_t *out, const uint8_t *row, const uint8_t *prev, size_t n, unsigned step, unsigned filter)
{
    uint64_t sum = 0;
    size_t i;

    for (i = 0; i < n; i++)
    {
        int left = i >= step ? row [i - step] : 0;
        int up = prev ? prev [i] : 0;
        int ul = (prev && i >= step) ? prev [i - step] : 0;
        uint8_t v;

        switch (filter)
        {
            case 1: v = row [i] - left; break;
            case 2: v = row [i] - up; break;
            case 3: v = row [i] - ((left + up) >> 1); break;
            case 4: v = row [i] - paeth (left, up, ul); break;
            default: v = row [i]; break;
        }
        out [i] = v;
        sum += v < 128 ? v : 256 - v;
    }
    return sum;
}

/* --- Decoder --- */

typedef struct
{
    SmolPngInfo info;
    unsigned channels;          /* samples per pixel in the file */
    unsigned bits_per_pixel;
    uint8_t palette [256] [4];  /* R, G, B, A */
    unsigned n_palette;
    uint16_t trns_key [3];      /* colour key of grey / RGB files */
}
PngHeader;

static size_t
row_bytes (const PngHeader *h, uint32_t width)
{
    return ((size_t) width * h->bits_per_pixel + 7) / 8;
}

/* Expands one reconstructed row of `width` pixels to RGBA8, writing every `dx`-th pixel of
 * `dest` (dx = 1 except for Adam7 passes). */
static void
expand_row (const PngHeader *h, const uint8_t *row, uint32_t width, uint8_t *dest, size_t dx)
{
    const unsigned depth = h->info.bit_depth;
    const unsigned sample_bytes = depth == 16 ? 2 : 1;
    uint32_t x;

    /* the two layouts nearly every file has */
    if (depth == 8 && dx == 1 && h->info.color_type == 6)
    {
        memcpy (dest, row, (size_t) width * 4);
        return;
    }
    if (depth == 8 && dx == 1 && h->info.color_type == 2 && !h->info.has_trns)
    {
        for (x = 0; x < width; x++, row += 3, dest += 4)
        {
            dest [0] = row [0]; dest [1] = row [1]; dest [2] = row [2]; dest [3] = 255;
        }
        return;
    }

    for (x = 0; x < width; x++, dest += dx * 4)
    {
        uint16_t s [4] = { 0, 0, 0, 0 };
        unsigned c;

        if (depth < 8)
        {
            size_t bit = (size_t) x * depth;
            s [0] = (row [bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
        }
        else
        {
            const uint8_t *p = row + (size_t) x * h->channels * sample_bytes;

            for (c = 0; c < h->channels; c++)
                s [c] = depth == 16 ? (uint16_t) ((p [c * 2] << 8) | p [c * 2 + 1]) : p [c];
        }

        switch (h->info.color_type)
        {
            case 3:
                /* an index past the palette is an error libpng renders as opaque black */
                if (s [0] < h->n_palette)
                    memcpy (dest, h->palette [s [0]], 4);
                else
                    dest [0] = dest [1] = dest [2] = 0, dest [3] = 255;
                break;
            case 0:
            case 4:
            {
                uint8_t g;

                /* 1/2/4-bit grey scales to the full range (x * 255 / max); 16-bit keeps the high byte */
                if (depth < 8)
                    g = (uint8_t) (s [0] * 255u / ((1u << depth) - 1));
                else
                    g = depth == 16 ? s [0] >> 8 : s [0];
                dest [0] = dest [1] = dest [2] = g;
                if (h->info.color_type == 4)
                    dest [3] = depth == 16 ? s [1] >> 8 : s [1];
                else
                    dest [3] = (h->info.has_trns && s [0] == h->trns_key [0]) ? 0 : 255;
                break;
            }
            default:
                for (c = 0; c < 3; c++)
                    dest [c] = depth == 16 ? s [c] >> 8 : s [c];
                if (h->info.color_type == 6)
                    dest [3] = depth == 16 ? s [3] >> 8 : s [3];
                else
                    dest [3] = (h->info.has_trns && s [0] == h->trns_key [0] && s [1] == h->trns_key [1]
                                && s [2] == h->trns_key [2]) ? 0 : 255;
                break;
        }
    }
}

static int
parse_ihdr (PngHeader *h, const uint8_t *d, uint32_t len)
{
    static const uint8_t channels_of [7] = { 1, 0, 3, 1, 2, 0, 4 };
    unsigned depth, ct;

    if (len != 13)
        return SMOL_PNG_ERR_CORRUPT;
    h->info.width = get_be32 (d);
    h->info.height = get_be32 (d + 4);
    depth = h->info.bit_depth = d [8];
    ct = h->info.color_type = d [9];
    h->info.interlace = d [12];
    if (h->info.width == 0 || h->info.height == 0 || h->info.width > 0x7fffffffu || h->info.height > 0x7fffffffu)
        return SMOL_PNG_ERR_CORRUPT;
    if (d [10] != 0 || d [11] != 0 || d [12] > 1)
        return SMOL_PNG_ERR_CORRUPT;
    if (ct > 6 || channels_of [ct] == 0)
        return SMOL_PNG_ERR_CORRUPT;
    if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16)
        return SMOL_PNG_ERR_CORRUPT;
    if ((ct == 3 && depth == 16) || ((ct == 2 || ct == 4 || ct == 6) && depth < 8))
        return SMOL_PNG_ERR_CORRUPT;
    h->channels = channels_of [ct];
    h->bits_per_pixel = h->channels * depth;
    return SMOL_PNG_OK;
}

/* Adam7: origin and spacing of the seven passes */
static const uint8_t adam7_x0 [7] = { 0, 4, 0, 2, 0, 1, 0 }, adam7_dx [7] = { 8, 8, 4, 4, 2, 2, 1 };
static const uint8_t adam7_y0 [7] = { 0, 0, 4, 0, 2, 0, 1 }, adam7_dy [7] = { 8, 8, 8, 4, 4, 2, 2 };

SMOL_PNG_EXPORT int
smol_png_decode_mem (const void *png, size_t png_size, uint32_t *width_out, uint32_t *height_out,
                     void **rgba_out, SmolPngInfo *info_out)
{
    const uint8_t *p = (const uint8_t *) png, *end = p + png_size;
    PngHeader h;
    uint8_t *idat = NULL, *raw = NULL, *rgba = NULL;
    size_t idat_size = 0, idat_cap = 0, raw_size = 0, pos;
    int seen_ihdr = 0, seen_iend = 0, err = SMOL_PNG_OK;
    unsigned pass, step, i;

    if (!png || !width_out || !height_out || !rgba_out)
        return SMOL_PNG_ERR_ARGUMENT;
    *rgba_out = NULL;
    if (png_size < 8 || memcmp (p, png_signature, 8))
        return SMOL_PNG_ERR_SIGNATURE;
    p += 8;
    memset (&h, 0, sizeof (h));

    while (!seen_iend)
    {
        uint32_t len, type;

        if ((size_t) (end - p) < 12)
        { err = SMOL_PNG_ERR_CORRUPT; goto out; }
        len = get_be32 (p);
        type = get_be32 (p + 4);
        if (len > 0x7fffffffu || (size_t) (end - p) - 12 < len)
        { err = SMOL_PNG_ERR_CORRUPT; goto out; }
        if (get_be32 (p + 8 + len) != (uint32_t) crc32 (crc32 (0, Z_NULL, 0), p + 4, len + 4))
        { err = SMOL_PNG_ERR_CORRUPT; goto out; }
        if (!seen_ihdr && type != CHUNK ('I', 'H', 'D', 'R'))
        { err = SMOL_PNG_ERR_CORRUPT; goto out; }

        switch (type)
        {
            case CHUNK ('I', 'H', 'D', 'R'):
                if (seen_ihdr || (err = parse_ihdr (&h, p + 8, len)) != SMOL_PNG_OK)
                { err = err ? err : SMOL_PNG_ERR_CORRUPT; goto out; }
                seen_ihdr = 1;
                break;
            case CHUNK ('P', 'L', 'T', 'E'):
                if (len % 3 != 0 || len > 768)
                { err = SMOL_PNG_ERR_CORRUPT; goto out; }
                h.n_palette = len / 3;
                for (i = 0; i < h.n_palette; i++)
                {
                    memcpy (h.palette [i], p + 8 + i * 3, 3);
                    h.palette [i] [3] = 255;
                }
                break;
            case CHUNK ('t', 'R', 'N', 'S'):
                if (h.info.color_type == 3)
                {
                    if (len > 256)
                    { err = SMOL_PNG_ERR_CORRUPT; goto out; }
                    for (i = 0; i < len; i++)
                        h.palette [i] [3] = p [8 + i];
                }
                else if (h.info.color_type == 0 && len == 2)
                    h.trns_key [0] = (p [8] << 8) | p [9];
                else if (h.info.color_type == 2 && len == 6)
                    for (i = 0; i < 3; i++)
                        h.trns_key [i] = (p [8 + i * 2] << 8) | p [9 + i * 2];
                else
                { err = SMOL_PNG_ERR_CORRUPT; goto out; }
                h.info.has_trns = 1;
                break;
            case CHUNK ('I', 'D', 'A', 'T'):
                if (idat_size + len > idat_cap)
                {
                    uint8_t *n;

                    idat_cap = (idat_size + len) * 2 + 4096;
                    n = (uint8_t *) realloc (idat, idat_cap);
                    if (!n)
                    { err = SMOL_PNG_ERR_MEMORY; goto out; }
                    idat = n;
                }
                memcpy (idat + idat_size, p + 8, len);
                idat_size += len;
                break;
            case CHUNK ('I', 'E', 'N', 'D'):
                seen_iend = 1;
                break;
            default:
                /* an unknown critical chunk (upper-case first letter) means the image cannot be shown */
                if (!((type >> 24) & 0x20))
                { err = SMOL_PNG_ERR_UNSUPPORTED; goto out; }
                break;
        }
        p += 12 + len;
    }

    if (!seen_ihdr || idat_size == 0 || (h.info.color_type == 3 && h.n_palette == 0))
    { err = SMOL_PNG_ERR_CORRUPT; goto out; }
    if ((uint64_t) h.info.width * h.info.height > ((uint64_t) 1 << 32))
    { err = SMOL_PNG_ERR_UNSUPPORTED; goto out; }

    /* Size of the filtered scanline stream: one filter byte per row of every (sub)image */
    if (h.info.interlace)
    {
        for (pass = 0; pass < 7; pass++)
        {
            uint32_t pw = (h.info.width + adam7_dx [pass] - 1 - adam7_x0 [pass]) / adam7_dx [pass];
            uint32_t ph = (h.info.height + adam7_dy [pass] - 1 - adam7_y0 [pass]) / adam7_dy [pass];

            if (pw && ph)
                raw_size += (size_t) ph * (1 + row_bytes (&h, pw));
        }
    }
    else
        raw_size = (size_t) h.info.height * (1 + row_bytes (&h, h.info.width));

    raw = (uint8_t *) malloc (raw_size ? raw_size : 1);
    rgba = (uint8_t *) malloc ((size_t) h.info.width * h.info.height * 4);
    if (!raw || !rgba)
    { err = SMOL_PNG_ERR_MEMORY; goto out; }

    {
        z_stream zs;
        int zr;

        memset (&zs, 0, sizeof (zs));
        if (inflateInit (&zs) != Z_OK)
        { err = SMOL_PNG_ERR_MEMORY; goto out; }
        pos = 0;
        zs.next_out = raw;
        do
        {
            /* avail_in / avail_out are 32-bit: feed the stream in pieces */
            size_t in_left = idat_size - pos, out_left = raw_size - (size_t) (zs.next_out - raw);

            if (zs.avail_in == 0)
            {
                zs.next_in = idat + pos;
                zs.avail_in = in_left > 0x40000000u ? 0x40000000u : (uInt) in_left;
                pos += zs.avail_in;
            }
            zs.avail_out = out_left > 0x40000000u ? 0x40000000u : (uInt) out_left;
            zr = inflate (&zs, Z_NO_FLUSH);
        }
        while (zr == Z_OK && ((size_t) (zs.next_out - raw) < raw_size || zs.avail_in > 0 || pos < idat_size));
        inflateEnd (&zs);
        /* the stream must hold exactly the scanlines (a short stream is an error; trailing
         * bytes after a complete image are tolerated, as libpng does with a warning) */
        if ((size_t) (zs.next_out - raw) != raw_size || (zr != Z_STREAM_END && zr != Z_OK && zr != Z_BUF_ERROR))
        { err = SMOL_PNG_ERR_CORRUPT; goto out; }
    }

    step = h.bits_per_pixel >= 8 ? h.bits_per_pixel / 8 : 1;
    pos = 0;
    for (pass = 0; pass < (h.info.interlace ? 7u : 1u); pass++)
    {
        uint32_t pw = h.info.width, ph = h.info.height, x0 = 0, y0 = 0, dx = 1, dy = 1, y;
        const uint8_t *prev = NULL;
        size_t rb;

        if (h.info.interlace)
        {
            x0 = adam7_x0 [pass]; dx = adam7_dx [pass];
            y0 = adam7_y0 [pass]; dy = adam7_dy [pass];
            pw = (h.info.width + dx - 1 - x0) / dx;
            ph = (h.info.height + dy - 1 - y0) / dy;
            if (pw == 0 || ph == 0)
                continue;
        }
        rb = row_bytes (&h, pw);
        for (y = 0; y < ph; y++)
        {
            uint8_t *row = raw + pos + 1;

            if ((err = unfilter_row (row, prev, rb, step, raw [pos])) != SMOL_PNG_OK)
                goto out;
            expand_row (&h, row, pw, rgba + ((size_t) (y0 + y * dy) * h.info.width + x0) * 4, dx);
            prev = row;
            pos += 1 + rb;
        }
    }

    *width_out = h.info.width;
    *height_out = h.info.height;
    *rgba_out = rgba;
    rgba = NULL;
    if (info_out)
        *info_out = h.info;

out:
    free (idat);
    free (raw);
    free (rgba);
    return err;
}

/* --- Encoder --- */

static size_t
put_chunk (uint8_t *out, uint32_t type, const uint8_t *data, uint32_t len)
{
    put_be32 (out, len);
    put_be32 (out + 4, type);
    if (len)
        memcpy (out + 8, data, len);
    put_be32 (out + 8 + len, (uint32_t) crc32 (crc32 (0, Z_NULL, 0), out + 4, len + 4));
    return 12 + (size_t) len;
}

SMOL_PNG_EXPORT int
smol_png_encode_mem (const void *pixels, uint32_t width, uint32_t height, uint32_t rowstride,
                     int channels, int level, void **png_out, size_t *png_size_out)
{
    const uint8_t *src = (const uint8_t *) pixels;
    const size_t rb = (size_t) width * (channels > 0 ? channels : 0);
    const size_t idat_piece = 1u << 20;
    uint8_t *filtered = NULL, *cand = NULL, *z = NULL, *out = NULL;
    size_t z_cap, z_size, n_pieces, o;
    uint8_t ihdr [13];
    z_stream zs;
    uint32_t y;
    int err = SMOL_PNG_OK, zr;

    if (!pixels || !png_out || !png_size_out || width == 0 || height == 0 || width > 0x7fffffffu
        || height > 0x7fffffffu || (channels != 3 && channels != 4) || rowstride < rb || level < 0 || level > 9)
        return SMOL_PNG_ERR_ARGUMENT;
    *png_out = NULL;
    *png_size_out = 0;

    filtered = (uint8_t *) malloc ((size_t) height * (1 + rb));
    cand = (uint8_t *) malloc (rb);
    if (!filtered || !cand)
    { err = SMOL_PNG_ERR_MEMORY; goto out; }

    for (y = 0; y < height; y++)
    {
        const uint8_t *row = src + (size_t) y * rowstride;
        const uint8_t *prev = y ? row - rowstride : NULL;
        uint8_t *dest = filtered + (size_t) y * (1 + rb);
        uint64_t best = UINT64_MAX;
        unsigned f;

        for (f = 0; f < 5; f++)
        {
            uint64_t sum = filter_row (cand, row, prev, rb, channels, f);

            if (sum < best)
            {
                best = sum;
                dest [0] = f;
                memcpy (dest + 1, cand, rb);
            }
        }
    }

    memset (&zs, 0, sizeof (zs));
    if (deflateInit (&zs, level) != Z_OK)
    { err = SMOL_PNG_ERR_MEMORY; goto out; }
    {
        size_t in_size = (size_t) height * (1 + rb), in_pos = 0;

        z_cap = deflateBound (&zs, in_size);
        z = (uint8_t *) malloc (z_cap);
        if (!z)
        { deflateEnd (&zs); err = SMOL_PNG_ERR_MEMORY; goto out; }
        zs.next_out = z;
        do
        {
            size_t out_left = z_cap - (size_t) (zs.next_out - z);

            if (zs.avail_in == 0 && in_pos < in_size)
            {
                size_t n = in_size - in_pos;

                zs.next_in = filtered + in_pos;
                zs.avail_in = n > 0x40000000u ? 0x40000000u : (uInt) n;
                in_pos += zs.avail_in;
            }
            zs.avail_out = out_left > 0x40000000u ? 0x40000000u : (uInt) out_left;
            zr = deflate (&zs, in_pos == in_size ? Z_FINISH : Z_NO_FLUSH);
        }
        while (zr == Z_OK);
        z_size = (size_t) (zs.next_out - z);
        deflateEnd (&zs);
        if (zr != Z_STREAM_END)
        { err = SMOL_PNG_ERR_MEMORY; goto out; }
    }

    n_pieces = (z_size + idat_piece - 1) / idat_piece;
    out = (uint8_t *) malloc (8 + (12 + 13) + z_size + 12 * n_pieces + 12);
    if (!out)
    { err = SMOL_PNG_ERR_MEMORY; goto out; }
    memcpy (out, png_signature, 8);
    o = 8;
    put_be32 (ihdr, width);
    put_be32 (ihdr + 4, height);
    ihdr [8] = 8;
    ihdr [9] = channels == 4 ? 6 : 2;
    ihdr [10] = ihdr [11] = ihdr [12] = 0;
    o += put_chunk (out + o, CHUNK ('I', 'H', 'D', 'R'), ihdr, 13);
    {
        size_t zp;

        for (zp = 0; zp < z_size; zp += idat_piece)
            o += put_chunk (out + o, CHUNK ('I', 'D', 'A', 'T'), z + zp,
                            (uint32_t) (z_size - zp < idat_piece ? z_size - zp : idat_piece));
    }
    o += put_chunk (out + o, CHUNK ('I', 'E', 'N', 'D'), NULL, 0);
    *png_out = out;
    *png_size_out = o;
    out = NULL;

out:
    free (filtered);
    free (cand);
    free (z);
    free (out);
    return err;
}

/* --- Files --- */

SMOL_PNG_EXPORT int
smol_png_load (const char *file_name, uint32_t *width_out, uint32_t *height_out, void **rgba_out)
{
    FILE *fp;
    uint8_t *buf = NULL;
    size_t size = 0, cap = 0;
    int err;

    if (!file_name)
        return SMOL_PNG_ERR_ARGUMENT;
    fp = fopen (file_name, "rb");
    if (!fp)
        return SMOL_PNG_ERR_IO;
    for (;;)
    {
        size_t n;

        if (size == cap)
        {
            uint8_t *nb;

            cap = cap ? cap * 2 : 1 << 16;
            nb = (uint8_t *) realloc (buf, cap);
            if (!nb)
            { free (buf); fclose (fp); return SMOL_PNG_ERR_MEMORY; }
            buf = nb;
        }
        n = fread (buf + size, 1, cap - size, fp);
        size += n;
        if (n == 0)
            break;
    }
    err = ferror (fp) ? SMOL_PNG_ERR_IO : SMOL_PNG_OK;
    fclose (fp);
    if (err == SMOL_PNG_OK)
        err = smol_png_decode_mem (buf, size, width_out, height_out, rgba_out, NULL);
    free (buf);
    return err;
}

SMOL_PNG_EXPORT int
smol_png_save (const char *file_name, const void *rgba, uint32_t width, uint32_t height, uint32_t rowstride)
{
    void *png = NULL;
    size_t size = 0;
    FILE *fp;
    int err;

    if (!file_name)
        return SMOL_PNG_ERR_ARGUMENT;
    /* compression level 5: png.c:127 */
    err = smol_png_encode_mem (rgba, width, height, rowstride, 4, 5, &png, &size);
    if (err != SMOL_PNG_OK)
        return err;
    fp = fopen (file_name, "wb");
    if (!fp)
    { free (png); return SMOL_PNG_ERR_IO; }
    if (fwrite (png, 1, size, fp) != size)
        err = SMOL_PNG_ERR_IO;
    if (fclose (fp) != 0)
        err = SMOL_PNG_ERR_IO;
    free (png);
    return err;
}

SMOL_PNG_EXPORT const char *
smol_png_strerror (int err)
{
    switch (err)
    {
        case SMOL_PNG_OK: return "ok";
        case SMOL_PNG_ERR_IO: return "file could not be opened, read or written";
        case SMOL_PNG_ERR_SIGNATURE: return "not a PNG file";
        case SMOL_PNG_ERR_CORRUPT: return "corrupt PNG stream";
        case SMOL_PNG_ERR_UNSUPPORTED: return "unsupported PNG feature";
        case SMOL_PNG_ERR_MEMORY: return "out of memory";
        case SMOL_PNG_ERR_ARGUMENT: return "invalid argument";
        default: return "unknown error";
    }
}

/* --- The reference's helper interface (png.c:158-209) --- */

static void
png_fatal (const char *fmt, ...)
{
    va_list args;

    va_start (args, fmt);
    vfprintf (stderr, fmt, args);
    fputc ('\n', stderr);
    va_end (args);
    abort ();
}

SMOL_PNG_EXPORT int
smoltest_load_image (const char *file_name, unsigned int *width_out, unsigned int *height_out, void **data_out)
{
    FILE *fp;
    uint8_t head [26];
    uint32_t w, h;
    size_t n;
    int err;

    /* The reference insists on an 8-bit RGBA file and says why it stops (png.c:90-97) */
    fp = fopen (file_name, "rb");
    if (!fp)
        png_fatal ("File %s could not be opened for reading", file_name);
    n = fread (head, 1, sizeof (head), fp);
    fclose (fp);
    if (n < 8 || memcmp (head, png_signature, 8))
        png_fatal ("File %s is not a PNG file", file_name);
    if (n == sizeof (head) && head [25] == 2)
        png_fatal ("Input file is PNG_COLOR_TYPE_RGB but must be PNG_COLOR_TYPE_RGBA (missing alpha channel)");
    if (n == sizeof (head) && (head [25] != 6 || head [24] != 8))
        png_fatal ("Color_type of input file must be PNG_COLOR_TYPE_RGBA (6) (is %d)", head [25]);

    err = smol_png_load (file_name, &w, &h, data_out);
    if (err != SMOL_PNG_OK)
        png_fatal ("Error during read_image: %s", smol_png_strerror (err));
    *width_out = w;
    *height_out = h;
    return 1;
}

SMOL_PNG_EXPORT void
smoltest_save_image (const char *prefix, uint32_t *data, unsigned int width, unsigned int height)
{
    size_t len = strlen (prefix) + 32;
    char *file_name = (char *) malloc (len);
    int err;

    if (!file_name)
        png_fatal ("out of memory");
    snprintf (file_name, len, "%s-%04u-%04u.png", prefix, width, height);
    err = smol_png_save (file_name, data, width, height, width * 4);
    if (err != SMOL_PNG_OK)
        png_fatal ("File %s could not be written: %s", file_name, smol_png_strerror (err));
    free (file_name);
}

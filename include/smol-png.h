/* -*- Mode: C; tab-width: 4; indent-tabs-mode: nil; c-basic-offset: 4 -*- */

/* PNG file I/O either side of the scaling path.
 *
 * The reference keeps this in its test program, not in the library: png.c:159-209
 * (smoltest_load_image / smoltest_save_image over libpng) feeding `test ... generate`
 * (test.c:1303-1371).  libpng is not in this image; this is a self-contained codec over
 * zlib's inflate/deflate.  It lives in its own shared library (libsmolpng.so) with no CUDA
 * dependency, so the scaling library's dependency list stays libc + the NVIDIA driver.
 *
 * Pixels are always handed over as 8-bit R,G,B,A in memory order (unassociated alpha),
 * i.e. SMOL_PIXEL_RGBA8_UNASSOCIATED. */

#ifndef _SMOL_PNG_H_
#define _SMOL_PNG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes of the smol_png_* calls (0 = success). */
enum
{
    SMOL_PNG_OK = 0,
    SMOL_PNG_ERR_IO = 1,          /* file could not be opened / read / written */
    SMOL_PNG_ERR_SIGNATURE = 2,   /* not a PNG stream */
    SMOL_PNG_ERR_CORRUPT = 3,     /* chunk structure, CRC, zlib stream or filter byte invalid */
    SMOL_PNG_ERR_UNSUPPORTED = 4, /* valid PNG this codec does not read */
    SMOL_PNG_ERR_MEMORY = 5,
    SMOL_PNG_ERR_ARGUMENT = 6
};

/* What the file held before it was expanded to RGBA8. */
typedef struct
{
    uint32_t width, height;
    uint8_t bit_depth;   /* 1, 2, 4, 8, 16 */
    uint8_t color_type;  /* 0 grey, 2 RGB, 3 palette, 4 grey + alpha, 6 RGBA */
    uint8_t interlace;   /* 0 none, 1 Adam7 */
    uint8_t has_trns;
}
SmolPngInfo;

/* Decodes a PNG stream in memory into a malloc'd width * height * 4 byte RGBA8 image (tightly
 * packed rows).  Every colour type and bit depth of the PNG specification is expanded: palette
 * and tRNS become alpha, grey is replicated, 16-bit samples keep their high byte; Adam7 is
 * de-interlaced.  The caller frees *rgba_out with free (). */
int smol_png_decode_mem (const void *png, size_t png_size,
                         uint32_t *width_out, uint32_t *height_out,
                         void **rgba_out, SmolPngInfo *info_out /* may be NULL */);

/* Encodes rows of 8-bit pixels into a malloc'd PNG stream.  channels = 4 (RGBA, colour type 6)
 * or 3 (RGB, colour type 2); rowstride in bytes; level = zlib compression level 0..9 (the
 * reference writes at 5, png.c:127).  Row filters are chosen per row by the minimum sum of
 * absolute differences. */
int smol_png_encode_mem (const void *pixels, uint32_t width, uint32_t height, uint32_t rowstride,
                         int channels, int level, void **png_out, size_t *png_size_out);

int smol_png_load (const char *file_name, uint32_t *width_out, uint32_t *height_out, void **rgba_out);
int smol_png_save (const char *file_name, const void *rgba, uint32_t width, uint32_t height,
                   uint32_t rowstride);

const char *smol_png_strerror (int err);

/* The reference's two helpers (png.c:158-209), same names and argument meaning, minus glib:
 *
 *   smoltest_load_image  returns 1 (TRUE) and a malloc'd RGBA8 image; like the reference it
 *                        accepts 8-bit RGBA files only and aborts with a message on anything
 *                        else (png.c:90-97, abort_ () at :22-32).
 *   smoltest_save_image  writes "<prefix>-WWWW-HHHH.png" (png.c:204), 8-bit RGBA, zlib level 5. */
int smoltest_load_image (const char *file_name, unsigned int *width_out, unsigned int *height_out,
                         void **data_out);
void smoltest_save_image (const char *prefix, uint32_t *data, unsigned int width, unsigned int height);

#ifdef __cplusplus
}
#endif

#endif

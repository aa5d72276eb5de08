/* smolscale.h -- public C API of the B200 (sm_100a) implementation.
 *
 * This header re-declares, symbol for symbol, the interface of the reference library
 * (reference smolscale.h:14-82) so that code written against the reference compiles and links
 * against libsmolscale_cuda.so unchanged.  Each declaration cites the reference line it replaces.
 *
 * Pointer semantics added by this implementation (invisible in the ABI):
 *   - pixels_in / pixels_out / outrows_dest may be ordinary host memory, pinned host memory,
 *     CUDA managed memory or device memory; the library classifies them per call.
 *   - With host (or pinned) output memory every call is fully synchronous, exactly like the
 *     reference: when the call returns, every output byte has been written.
 *   - With device output memory the work is enqueued on the library's current stream
 *     (smolscale-cuda.h: smol_cuda_set_stream; default: the legacy default stream) and the call
 *     returns without waiting, like cudaMemcpyAsync between device buffers.
 * There is no CPU fallback: if no CUDA device is usable the library prints a message and
 * aborts, as the reference does on an internal inconsistency (smolscale.c:779-780, :812-813). */

#ifndef SMOLSCALE_B200_SMOLSCALE_H
#define SMOLSCALE_B200_SMOLSCALE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Replaces reference smolscale.h:14-35.  Values are ABI (passed as int). */
typedef enum
{
    /* 32 bits per pixel */
    SMOL_PIXEL_RGBA8_PREMULTIPLIED = 0,
    SMOL_PIXEL_BGRA8_PREMULTIPLIED = 1,
    SMOL_PIXEL_ARGB8_PREMULTIPLIED = 2,
    SMOL_PIXEL_ABGR8_PREMULTIPLIED = 3,
    SMOL_PIXEL_RGBA8_UNASSOCIATED = 4,
    SMOL_PIXEL_BGRA8_UNASSOCIATED = 5,
    SMOL_PIXEL_ARGB8_UNASSOCIATED = 6,
    SMOL_PIXEL_ABGR8_UNASSOCIATED = 7,
    /* 24 bits per pixel */
    SMOL_PIXEL_RGB8 = 8,
    SMOL_PIXEL_BGR8 = 9,

    SMOL_PIXEL_MAX = 10
}
SmolPixelType;

/* Replaces reference smolscale.h:37-39.  Called once per finished output row on the calling
 * thread, on host memory; may modify the row in place. */
typedef void (SmolPostRowFunc) (uint32_t *row_inout, int width, void *user_data);

/* Replaces reference smolscale.h:41 (opaque). */
typedef struct SmolScaleCtx SmolScaleCtx;

/* Replaces reference smolscale.h:47-51: scale a whole image in one call.  Row strides are in
 * bytes and arbitrary; only width * bytes-per-pixel bytes of each row are read / written. */
void smol_scale_simple (const void *pixels_in, SmolPixelType pixel_type_in,
                        uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                        void *pixels_out, SmolPixelType pixel_type_out,
                        uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                        uint8_t with_srgb);

/* Replaces reference smolscale.h:55-59. */
SmolScaleCtx *smol_scale_new (const void *pixels_in, SmolPixelType pixel_type_in,
                              uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                              void *pixels_out, SmolPixelType pixel_type_out,
                              uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                              uint8_t with_srgb);

/* Replaces reference smolscale.h:61-66. */
SmolScaleCtx *smol_scale_new_full (const void *pixels_in, SmolPixelType pixel_type_in,
                                   uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                                   void *pixels_out, SmolPixelType pixel_type_out,
                                   uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                                   uint8_t with_srgb,
                                   SmolPostRowFunc post_row_func, void *user_data);

/* Replaces reference smolscale.h:68. */
void smol_scale_destroy (SmolScaleCtx *scale_ctx);

/* Replaces reference smolscale.h:70-74.  Re-entrant on one shared context from concurrent
 * threads as long as the row ranges do not overlap; no locking needed by the caller. */
void smol_scale_batch (const SmolScaleCtx *scale_ctx, uint32_t first_outrow, uint32_t n_outrows);

/* Replaces reference smolscale.h:76-82: same rows, written contiguously from outrows_dest with
 * the context's rowstride_out pitch. */
void smol_scale_batch_full (const SmolScaleCtx *scale_ctx,
                            void *outrows_dest,
                            uint32_t first_outrow, uint32_t n_outrows);

#ifdef __cplusplus
}
#endif

#endif

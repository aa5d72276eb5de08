/* smolscale-cuda.h -- OPTIONAL extensions of the B200 implementation.
 *
 * Nothing here is needed by code written against the reference API (smolscale.h); these entry
 * points exist for callers that already live on the GPU (stream control, batched submission of
 * many same-shaped images in one launch) and for tests / benchmarks (plan introspection,
 * counters).  Plain C ABI: pointers, sizes and ints only. */

#ifndef SMOLSCALE_B200_SMOLSCALE_CUDA_H
#define SMOLSCALE_B200_SMOLSCALE_CUDA_H

#include <stddef.h>
#include <stdint.h>
#include "smolscale.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device / stream control ------------------------------------------------------------ */

/* Number of usable CUDA devices (0 if none; never aborts). */
int smol_cuda_device_count (void);

/* Device used by the calling thread for staging host-memory calls (default: the thread's current
 * CUDA device).  Device-memory calls always run on the device that owns the output pointer. */
void smol_cuda_set_device (int device);

/* Stream (a cudaStream_t passed as void *) on which the calling thread's device-memory calls are
 * enqueued.  NULL = legacy default stream.  Thread-local.  It must belong to the device that owns
 * the buffers.  Calls with device memory on both sides return as soon as the kernel is enqueued;
 * calls that mix a device buffer with a host buffer run their kernel on this stream too (so they
 * are ordered after whatever produced the device buffer) and return when the result is complete.
 * The first call for a new geometry uploads its filter tables outside any ongoing stream capture;
 * the kernel launch itself is capturable (CUDA graphs). */
void smol_cuda_set_stream (void *cuda_stream);

/* How many GPUs ONE host-memory call may be spread over (default 1; 0 = every visible device;
 * environment: SMOL_CUDA_MULTI_GPU=N|all).  With N > 1, smol_scale_simple / smol_scale_batch* on
 * host buffers split their output rows into N bands; each device uploads only the source rows its
 * band reads (band + filter halo) over its own PCIe link, scales them and writes its rows back.
 * Results are bit-identical to the single-device path (disjoint row batches are independent in
 * the reference too, smolscale.h:70-74).  Calls on device memory are unaffected.  Process-wide. */
void smol_cuda_set_multi_gpu (int n_devices);

/* Waits for all work this library has enqueued from the calling thread's stream. */
void smol_cuda_synchronize (void);

/* ---- batched submission ----------------------------------------------------------------- */

/* Scales n_images images of identical geometry and pixel types in ONE kernel launch
 * (grid.z = image).  Image i is read from (const char *) pixels_in + i * image_stride_in and
 * written to (char *) pixels_out + i * image_stride_out.  Both buffers must be device (or
 * managed) memory; the launch is enqueued on the calling thread's stream and not waited for.
 * The result of every image is identical to one smol_scale_simple call on it (the reference's
 * thumbnail-batch pattern is one smol_scale_simple per image per worker thread, test.c:785-804). */
void smol_cuda_scale_images (const void *pixels_in, size_t image_stride_in,
                             SmolPixelType pixel_type_in,
                             uint32_t width_in, uint32_t height_in, uint32_t rowstride_in,
                             void *pixels_out, size_t image_stride_out,
                             SmolPixelType pixel_type_out,
                             uint32_t width_out, uint32_t height_out, uint32_t rowstride_out,
                             uint8_t with_srgb, uint32_t n_images);

/* ---- plan introspection (pure host code; works without a GPU) --------------------------- */

enum { SMOL_CUDA_AXIS_COPY = 0, SMOL_CUDA_AXIS_ONE = 1, SMOL_CUDA_AXIS_BILINEAR = 2, SMOL_CUDA_AXIS_BOX = 3 };
enum { SMOL_CUDA_MID_P8 = 0, SMOL_CUDA_MID_P8L = 1, SMOL_CUDA_MID_P16 = 2, SMOL_CUDA_MID_P16L = 3 };

typedef struct
{
    int32_t filter_h, filter_v;           /* SMOL_CUDA_AXIS_* : what the reference's pick_filter_params selects (smolscale.c:427-478) */
    uint32_t halvings_h, halvings_v;      /* bilinear only */
    uint32_t bilin_w, bilin_h;            /* dim_out << halvings */
    int32_t storage_bits;                 /* 64 or 128 (smolscale.c:862, :751-758) */
    int32_t mid;                          /* SMOL_CUDA_MID_* */
    uint32_t span_mul_x, span_mul_y;      /* box only (smolscale-generic.c:89-91) */
    uint32_t n_tab_x, n_tab_y;            /* number of (offset, F) pairs in the reference-layout tables */
    int32_t kernel_id;                    /* which kernel family the dispatcher would launch */
    char kernel_name[64];
}
SmolCudaPlanInfo;

/* Fills *info for the given job.  tab_x / tab_y, when non-NULL, receive the fixed-point tables as
 * (absolute offset, F) uint16 pairs in the reference's semantics (bilinear: bilin_dim pairs,
 * box: dim_out + 1 pairs, copy / one: none); they must have room for 2 * n_tab_* uint16 (call
 * once with NULL to learn the sizes). */
void smol_cuda_plan_query (SmolPixelType pixel_type_in, uint32_t width_in, uint32_t height_in,
                           SmolPixelType pixel_type_out, uint32_t width_out, uint32_t height_out,
                           uint8_t with_srgb,
                           SmolCudaPlanInfo *info, uint16_t *tab_x, uint16_t *tab_y);

/* Source rows [*first_inrow, *first_inrow + *n_inrows) that output rows
 * [first_outrow, first_outrow + n_outrows) of this context read (its band plus filter halo). */
void smol_cuda_band_source_rows (const SmolScaleCtx *scale_ctx,
                                 uint32_t first_outrow, uint32_t n_outrows,
                                 uint32_t *first_inrow, uint32_t *n_inrows);

/* ---- counters --------------------------------------------------------------------------- */

typedef struct
{
    uint64_t kernel_launches;             /* kernels of this library launched since reset */
    uint64_t h2d_bytes, d2h_bytes;        /* bytes this library copied for host-memory calls */
    uint64_t table_uploads;               /* filter-table uploads (cache misses) */
}
SmolCudaStats;

void smol_cuda_get_stats (SmolCudaStats *stats);
void smol_cuda_reset_stats (void);

/* Launches per kernel family since the last reset: fills counts[0..n) and, if `names` is not
 * NULL, the families' static names; returns n (at most max_families).  The family recorded is the
 * one that actually ran, after any fallback to the general kernel. */
#define SMOL_CUDA_MAX_KERNEL_FAMILIES 16
int smol_cuda_get_kernel_launches (uint64_t *counts, const char **names, int max_families);

/* Forces a kernel family for testing (0 = automatic).  See SMOL_KERNEL_* in
 * smolscale-cuda-private.h; families that cannot run the job fall back to the general one. */
void smol_cuda_force_kernel (int kernel_id);

#ifdef __cplusplus
}
#endif

#endif

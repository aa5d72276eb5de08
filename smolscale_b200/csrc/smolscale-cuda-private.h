/* smolscale-cuda-private.h -- structures shared by the host C layer (smolscale-cuda.c) and the
 * CUDA translation unit (smolscale-cuda-kernels.cu).  Plain C, includable from both. */

#ifndef SMOLSCALE_CUDA_PRIVATE_H
#define SMOLSCALE_CUDA_PRIVATE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-axis device filter class.  COPY, ONE and BILINEAR_nH of the reference all become "taps":
 * a table of (offset, F) pairs where output sample i is lerp (in[ofs], in[min (ofs + 1, dim - 1)], F)
 * and 2^halvings consecutive samples are summed and shifted.  COPY is ofs = i, F = 256; ONE is
 * ofs = 0, F = 256 (F = 256 makes the lerp return in[ofs] exactly). */
enum { SMOL_AXIS_TAPS = 0, SMOL_AXIS_BOX = 1 };

/* Intermediate pixel encodings (reference: alpha x gamma x storage, smolscale.c:724-778). */
enum { SMOL_MID_P8 = 0, SMOL_MID_P8L = 1, SMOL_MID_P16 = 2, SMOL_MID_P16L = 3 };

/* Kernel families. */
enum
{
    SMOL_KERNEL_AUTO = 0,
    SMOL_KERNEL_GENERAL = 1,     /* smem-staged rows, any filter combination, any format */
    SMOL_KERNEL_TAPS_DIRECT = 2, /* bilinear / copy / one on both axes, 64bpp, register-only */
    SMOL_KERNEL_HALF2X = 3,      /* exact 2^k:1 reductions (all F = 128), 32bpp in, packed-byte math */
    SMOL_KERNEL_BOX = 4,         /* box x box, tuned for large-span downscales */
    SMOL_KERNEL_MAG = 5,         /* vertical magnification: two-phase shared-memory tile */
    SMOL_KERNEL_TAPS128 = 6,     /* bilinear with halvings, 128bpp intermediate (linear light, P16) */
    SMOL_KERNEL_TILE128 = 7,     /* bilinear without halvings / copy / one, 128bpp intermediate */
    SMOL_KERNEL_MAGB = 8,        /* vertical magnification, byte-granular vertical stage (no alpha work on output) */
    SMOL_KERNEL_ROWS = 9,        /* any filter pair, any format: one warp per tile of columns, source rows streamed through shared memory */
    SMOL_KERNEL_MAX
};

/* Everything the device needs to know about one job's pixel formats and filters.  Immutable
 * after smol_scale_new; passed to kernels by value. */
typedef struct
{
    uint32_t w_in, h_in, w_out, h_out;

    uint8_t bpp_in, bpp_out;            /* 3 or 4 */
    uint8_t in_alpha_idx, in_col0;      /* byte index of alpha in an input pixel (0xff: none), of the first colour byte */
    uint8_t out_alpha_idx, out_col0;
    uint8_t in_unassoc, out_unassoc;
    uint8_t swap_rb;                    /* output colour order is the reverse of the input's */
    uint8_t mid;                        /* SMOL_MID_* */
    uint8_t storage128;                 /* 0: 64bpp intermediate (4 x 16-bit lanes), 1: 128bpp (4 x 32-bit lanes) */
    uint8_t pack24_direct;              /* P8L -> 24bpp: gamma-compress the premultiplied value (reference "123" packer quirk) */

    uint8_t h_kind, v_kind;             /* SMOL_AXIS_* */
    uint8_t h_halvings, v_halvings;     /* taps only */
    uint8_t all_half_x, all_half_y;     /* taps: every F == 128 and ofs == 2 * i (exact 2:1 per sample) */
    uint8_t pad_[2];

    uint32_t span_mul_x, span_mul_y;    /* box only */
    uint32_t n_tab_x, n_tab_y;          /* entries in the device tables */
}
SmolJobDesc;

/* The six data tables of the algorithm (reference smolscale.c:87-421), resident once per device. */
typedef struct
{
    uint32_t inv_div_p8[256];
    uint32_t inv_div_p8l[256];
    uint32_t inv_div_p16[256];
    uint32_t inv_div_p16l[256];
    uint16_t from_srgb[256];
    uint8_t to_srgb[2048];
}
SmolDeviceLuts;

/* Device tables: one uint32 per entry, low 16 bits = absolute offset, high 16 bits = F.
 * Taps axis: (dim_out << halvings) entries.  Box axis: dim_out + 1 entries (entry [i + 1] tells
 * where box i ends; the last one is the reference's sentinel pair). */
#define SMOL_TAB_OFS(e) ((e) & 0xffffu)
#define SMOL_TAB_F(e)   ((e) >> 16)

typedef struct
{
    SmolJobDesc d;
    const uint8_t *src;                 /* row 0 of the source image (device) */
    uint8_t *dst;                       /* where output row `first_row` goes (device) */
    uint32_t src_pitch, dst_pitch;
    size_t src_image_stride, dst_image_stride;  /* batched submission: bytes between images */
    uint32_t n_images;
    const uint32_t *tab_x, *tab_y;      /* device */
    const SmolDeviceLuts *luts;         /* device */
    /* device, 65536 x uint16 each, index (alpha << 8) | c: the whole 8-bit -> premultiplied
     * 11-bit linear unpack chain of one channel (reference generic:555-568 / :591-614) from a
     * premultiplied resp. unassociated source, precomputed once per device from the LUTs */
    const uint16_t *p8l_from_p, *p8l_from_u;
    uint32_t first_row, n_rows;         /* output rows to produce */
    /* launch shape chosen by the host */
    uint32_t lanes_per_col;             /* general kernel: threads cooperating on one output column (power of two) */
    uint32_t rows_per_cta;              /* output rows walked by one CTA */
    uint32_t chunk_px;                  /* general kernel: source pixels staged per pass */
}
SmolLaunch;

/* Launchers exported by the CUDA TU.  `stream` is a cudaStream_t.  They return 0 or a
 * cudaError_t value; *name_out (optional) receives a static kernel-family name. */
int smol_cuda_launch (const SmolLaunch *launch, int kernel_id, void *stream, const char **name_out);

/* Which family the dispatcher picks for a job (pure function of the descriptor + pointers'
 * alignment), and its name. */
int smol_cuda_pick_kernel (const SmolLaunch *launch, int forced);
const char *smol_cuda_kernel_name (int kernel_id);

#ifdef __cplusplus
}
#endif

#endif

/* TEST INFRASTRUCTURE ONLY.  Nothing in the product path (smolscale_b200/, include/) may
 * include, link or call anything under oracle/.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker.
 *
 * Plain-C, one-channel-at-a-time restatement of the smolscale scaling pipeline
 * (reference: smolscale.c, smolscale-generic.c).  Parity status: PINNED -- the restatement is
 * checked bit-for-bit against the unmodified reference compiled from /root/reference
 * (oracle/_ref/libsmolref.so, built by oracle/Makefile) by tests/test_oracle_vs_ref.py, and
 * against the committed golden digests in tests/golden/ (generated from that same compiled
 * reference by tools/gen_golden.py).  The reference itself ships no golden vectors
 * (SURVEY.md section 4 / 8c). */
#ifndef SMOL_ORACLE_H
#define SMOL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Pixel types: same numbering as SmolPixelType (reference smolscale.h:14-35). */
enum {
    ORACLE_RGBA8_P, ORACLE_BGRA8_P, ORACLE_ARGB8_P, ORACLE_ABGR8_P,
    ORACLE_RGBA8_U, ORACLE_BGRA8_U, ORACLE_ARGB8_U, ORACLE_ABGR8_U,
    ORACLE_RGB8, ORACLE_BGR8, ORACLE_PIXEL_MAX
};

/* Per-axis filter classes (reference SmolFilterType, smolscale-private.h:101-116). */
enum { ORACLE_F_COPY, ORACLE_F_ONE, ORACLE_F_BILINEAR, ORACLE_F_BOX };

/* Intermediate pixel encodings (reference alpha/gamma/storage triple, smolscale.c:724-778). */
enum {
    ORACLE_MID_P8,    /* premultiplied 8-bit, sRGB-compressed; 64bpp or 128bpp storage */
    ORACLE_MID_P8L,   /* premultiplied 11-bit linear light; 128bpp */
    ORACLE_MID_P16,   /* value*alpha (16 bit), alpha lane (a<<8)|0x80; 128bpp */
    ORACLE_MID_P16L   /* linear*alpha (19 bit), alpha lane (a<<8)|0x80; 128bpp */
};

typedef struct {
    uint32_t w_in, h_in, w_out, h_out;
    int type_in, type_out;
    int filter_h, filter_v;         /* ORACLE_F_* */
    uint32_t halvings_h, halvings_v;/* 0..2 for ORACLE_F_BILINEAR, else 0 */
    uint32_t bilin_w, bilin_h;      /* dim_out << halvings */
    int storage_bits;               /* 64 or 128 */
    int mid;                        /* ORACLE_MID_* */
    uint32_t span_mul_x, span_mul_y;/* box only */
    /* Tables of (absolute offset, F) pairs.  Bilinear: bilin_dim pairs.  Box: dim_out + 1 pairs.
     * COPY / ONE: NULL. */
    uint16_t *tab_x, *tab_y;
    uint32_t n_tab_x, n_tab_y;      /* number of pairs */
} oracle_plan;

/* Builds the plan (filter / storage / encoding selection and the fixed-point tables). */
void oracle_plan_init(oracle_plan *plan,
                      int type_in, uint32_t w_in, uint32_t h_in,
                      int type_out, uint32_t w_out, uint32_t h_out,
                      uint8_t with_srgb);
void oracle_plan_free(oracle_plan *plan);

/* Renders output rows [first_row, first_row + n_rows) contiguously from outrows_dest with
 * rowstride_out pitch (the smol_scale_batch_full contract, reference smolscale.c:998-1008). */
void oracle_scale_rows(const oracle_plan *plan,
                       const void *pixels_in, uint32_t rowstride_in,
                       void *outrows_dest, uint32_t rowstride_out,
                       uint32_t first_row, uint32_t n_rows);

/* Same contract as smol_scale_simple (reference smolscale.c:957-985). */
void oracle_scale_simple(const void *pixels_in, int type_in,
                         uint32_t w_in, uint32_t h_in, uint32_t rowstride_in,
                         void *pixels_out, int type_out,
                         uint32_t w_out, uint32_t h_out, uint32_t rowstride_out,
                         uint8_t with_srgb);

#ifdef __cplusplus
}
#endif
#endif

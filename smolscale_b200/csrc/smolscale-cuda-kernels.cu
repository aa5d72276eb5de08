/* smolscale-cuda-kernels.cu -- the device side of the B200 (sm_100a) smolscale pipeline.
 *
 * One fused kernel does what the reference does in four row passes (reference call stack,
 * smolscale.c:491-546 -> smolscale-generic.c:1613-1642, :1648-2318): unpack the packed 24/32bpp
 * source pixels into the intermediate representation (with premultiply / unpremultiply through
 * the inverse-division tables and optional sRGB linearisation), filter horizontally, filter
 * vertically, and repack -- intermediate rows live in registers / shared memory only and never
 * round-trip HBM.
 *
 * Arithmetic contract.  The reference computes on uint64_t words that hold four 16-bit lanes
 * ("64bpp") or two 32-bit lanes ("128bpp").  We keep the same word layout idea (so box sums use
 * the same 64-bit adds) but write every weighted tap in its non-negative form
 *     lerp (p, q, F) = ((p * F + q * (256 - F)) >> 8) & mask
 * which is algebraically identical, lane by lane, to the reference's
 *     ((((p - q) * F) >> 8) + q) & mask            (smolscale-generic.c:1317, :1704, ...)
 * because p*F + q*(256-F) = (p-q)*F + 256*q, and never borrows across lanes: every lane product
 * stays below the lane width for the value ranges the unpackers can produce (8-bit payload in
 * 16-bit lanes, <= 19-bit payload in 32-bit lanes).  tests/ check all of this bit-for-bit
 * against the compiled reference and the plain-C oracle.
 *
 * No tensor cores: every output is a 2-tap (bilinear) or variable-span (box) integer stencil
 * with a floor after each tap, i.e. bandwidth-bound byte work, not a dense contraction. */

#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smolscale-cuda-private.h"

#define SMOL_BLOCK 256

/* ------------------------------------------------------------------------------------------ *
 * Intermediate pixel: 1 (64bpp) or 2 (128bpp) 64-bit words.                                  *
 *   64bpp : w[0] = alpha | c0 << 16 | c1 << 32 | c2 << 48       (8-bit payloads)             *
 *   128bpp: w[0] = alpha | c0 << 32,  w[1] = c1 | c2 << 32      (up to 19-bit payloads)      *
 * c0..c2 are the colour channels in INPUT memory order.  The reference orders lanes per pixel *
 * type (smolscale.c:647-719); since all lanes are filtered alike, lane order is free.        *
 * ------------------------------------------------------------------------------------------ */

template <bool S128> struct Px { uint64_t w[S128 ? 2 : 1]; };

template <bool S128> struct PxTraits;
template <> struct PxTraits<false> { static constexpr uint64_t MASK = 0x00ff00ff00ff00ffULL; static constexpr int N = 1; };
template <> struct PxTraits<true>  { static constexpr uint64_t MASK = 0x00ffffff00ffffffULL; static constexpr int N = 2; };

template <bool S128> __device__ __forceinline__ Px<S128> px_zero ()
{
    Px<S128> r;
#pragma unroll
    for (int i = 0; i < PxTraits<S128>::N; i++) r.w[i] = 0;
    return r;
}

template <bool S128> __device__ __forceinline__ void px_add (Px<S128> &a, const Px<S128> &b)
{
#pragma unroll
    for (int i = 0; i < PxTraits<S128>::N; i++) a.w[i] += b.w[i];
}

/* ((p * w) >> 8) & mask -- reference weight_pixel_64bpp / _128bpp (generic:1177-1192); also the
 * "(255 - F) * r" left-over of a box edge (generic:1462, :1524, :2065) since
 * ((r << 8) - r - r * F) == r * (255 - F). */
template <bool S128> __device__ __forceinline__ Px<S128> px_weight (const Px<S128> &p, uint32_t w)
{
    Px<S128> r;
#pragma unroll
    for (int i = 0; i < PxTraits<S128>::N; i++) r.w[i] = ((p.w[i] * w) >> 8) & PxTraits<S128>::MASK;
    return r;
}

template <bool S128> __device__ __forceinline__ Px<S128> px_lerp (const Px<S128> &p, const Px<S128> &q, uint32_t F)
{
    Px<S128> r;
#pragma unroll
    for (int i = 0; i < PxTraits<S128>::N; i++)
        r.w[i] = ((p.w[i] * F + q.w[i] * (256u - F)) >> 8) & PxTraits<S128>::MASK;
    return r;
}

/* (acc >> n) & mask -- the halving step (generic:1319, :1357-1358, :1806, :1834) */
template <bool S128> __device__ __forceinline__ Px<S128> px_halve (const Px<S128> &a, uint32_t n)
{
    Px<S128> r;
#pragma unroll
    for (int i = 0; i < PxTraits<S128>::N; i++) r.w[i] = (a.w[i] >> n) & PxTraits<S128>::MASK;
    return r;
}

/* Box normalisation: scale_64bpp (generic:1231-1245) / scale_128bpp_half (generic:1247-1261). */
template <bool S128> __device__ __forceinline__ Px<S128> px_box_scale (const Px<S128> &acc, uint32_t mul)
{
    Px<S128> r;
    const uint64_t half = 1ull << 23;
    if constexpr (!S128)
    {
        uint64_t a = ((acc.w[0] & 0x0000ffff0000ffffULL) * mul + half + (half << 32)) >> 24;
        uint64_t b = (((acc.w[0] & 0xffff0000ffff0000ULL) >> 16) * mul + half + (half << 32)) >> 24;
        r.w[0] = (a & 0x000000ff000000ffULL) | ((b & 0x000000ff000000ffULL) << 16);
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            uint64_t a = ((acc.w[i] & 0xffffffffULL) * mul + half) >> 24;
            uint64_t b = ((acc.w[i] >> 32) * mul + half) >> 24;
            r.w[i] = (a & 0xffffULL) | ((b & 0xffffULL) << 32);
        }
    }
    return r;
}

template <bool S128> __device__ __forceinline__ Px<S128> px_shfl_xor_add (Px<S128> a, uint32_t m)
{
#pragma unroll
    for (int i = 0; i < PxTraits<S128>::N; i++)
        a.w[i] += __shfl_xor_sync (0xffffffffu, a.w[i], m);
    return a;
}

/* ------------------------------------------------------------------------------------------ *
 * Unpack / pack (reference smolscale-generic.c:349-1164 with helpers :185-318)               *
 * ------------------------------------------------------------------------------------------ */

/* raw = the pixel's bytes in memory order, byte 0 in bits 0..7 (for 24bpp the top byte is ignored) */
template <bool S128>
__device__ __forceinline__ Px<S128> unpack_px (uint32_t raw, const SmolJobDesc &d, const SmolDeviceLuts *__restrict__ lut)
{
    uint32_t a = (d.in_alpha_idx == 0xff) ? 255u : ((raw >> (8 * d.in_alpha_idx)) & 0xff);
    uint32_t alane = a;
    uint32_t c[3];

#pragma unroll
    for (int i = 0; i < 3; i++)
        c[i] = (raw >> (8 * (d.in_col0 + i))) & 0xff;

    if (d.mid == SMOL_MID_P8)
    {
        if (d.in_unassoc)
        {
#pragma unroll
            for (int i = 0; i < 3; i++)
                c[i] = (((c[i] + 1) * (a + 1) - 1) >> 8) & 0xff;                 /* generic:238-244 */
        }
    }
    else if (d.mid == SMOL_MID_P8L)
    {
        const uint32_t inv = d.in_unassoc ? 0 : lut->inv_div_p8[a];
        const uint32_t am = (a << 3) + 1;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint32_t v = c[i];
            if (!d.in_unassoc)
                v = ((v * inv) >> 13) & 0xff;                                    /* generic:227-236 */
            v = lut->from_srgb[v];                                               /* generic:185-199 */
            c[i] = (((v + 1) * am - 1) >> 11) & 0x7ff;                           /* generic:261-269 */
        }
    }
    else
    {
        /* P16 / P16L: value * alpha, alpha lane carries 8 fraction bits (generic:616-660, :708-752) */
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint32_t v = c[i];
            if (d.mid == SMOL_MID_P16L)
                v = lut->from_srgb[v];
            c[i] = v * a;
        }
        alane = (a << 8) | 0x80;
    }

    Px<S128> r;
    if constexpr (S128)
    {
        r.w[0] = (uint64_t) alane | ((uint64_t) c[0] << 32);
        r.w[1] = (uint64_t) c[1] | ((uint64_t) c[2] << 32);
    }
    else
    {
        r.w[0] = (uint64_t) alane | ((uint64_t) c[0] << 16) | ((uint64_t) c[1] << 32) | ((uint64_t) c[2] << 48);
    }
    return r;
}

/* Returns the output pixel's bytes in memory order (byte 0 in bits 0..7). */
template <bool S128>
__device__ __forceinline__ uint32_t pack_px (const Px<S128> &p, const SmolJobDesc &d, const SmolDeviceLuts *__restrict__ lut)
{
    uint64_t lane[4];   /* alpha lane, c0, c1, c2 */

    if constexpr (S128)
    {
        lane[0] = p.w[0] & 0xffffffffULL; lane[1] = p.w[0] >> 32;
        lane[2] = p.w[1] & 0xffffffffULL; lane[3] = p.w[1] >> 32;
    }
    else
    {
        lane[0] = p.w[0] & 0xffff; lane[1] = (p.w[0] >> 16) & 0xffff;
        lane[2] = (p.w[0] >> 32) & 0xffff; lane[3] = p.w[0] >> 48;
    }

    uint32_t a;
    uint32_t c[3];

    if (d.mid == SMOL_MID_P16 || d.mid == SMOL_MID_P16L)
        a = (uint32_t) (lane[0] >> 8) & 0xff;                                    /* generic:1140, :1152 */
    else
        a = (uint32_t) lane[0] & 0xff;                                           /* generic:876, :1101 */

    if (d.mid == SMOL_MID_P8)
    {
        const uint32_t inv = lut->inv_div_p8[a];
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint64_t v = lane[i + 1];
            if (d.out_unassoc)
                v = ((v & 0xff) * inv) >> 13;                                    /* generic:246-259 */
            c[i] = (uint32_t) v & 0xff;
        }
    }
    else if (d.mid == SMOL_MID_P8L)
    {
        const uint32_t inv = lut->inv_div_p8l[a];
        const bool unpremul = !(d.bpp_out == 3 && d.pack24_direct);              /* generic:922-935 vs :1010-1023 */
        const bool repremul = (d.bpp_out == 4 && !d.out_unassoc);                /* generic:1096-1109 */
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint64_t v = lane[i + 1];
            if (unpremul)
                v = (v * inv) >> 10;                                             /* generic:271-280 */
            v = lut->to_srgb[v & 0x7ff];                                         /* generic:201-211 */
            if (repremul)
                v = (((v + 1) * (a + 1) - 1) >> 8) & 0xff;                       /* generic:217-225 */
            c[i] = (uint32_t) v & 0xff;
        }
    }
    else if (d.mid == SMOL_MID_P16)
    {
        const uint32_t inv = lut->inv_div_p16[a];
#pragma unroll
        for (int i = 0; i < 3; i++)
            c[i] = (uint32_t) ((lane[i + 1] * inv) >> 16) & 0xff;                /* generic:290-299 */
    }
    else
    {
        const uint32_t inv = lut->inv_div_p16l[a];
#pragma unroll
        for (int i = 0; i < 3; i++)
            c[i] = lut->to_srgb[(uint32_t) ((lane[i + 1] * inv) >> 19) & 0x7ff]; /* generic:309-318 */
    }

    if (d.swap_rb)
    {
        uint32_t t = c[0]; c[0] = c[2]; c[2] = t;
    }

    uint32_t out = (c[0] << (8 * d.out_col0)) | (c[1] << (8 * (d.out_col0 + 1))) | (c[2] << (8 * (d.out_col0 + 2)));
    if (d.out_alpha_idx != 0xff)
        out |= a << (8 * d.out_alpha_idx);
    return out;
}

__device__ __forceinline__ uint32_t load_raw_px (const uint8_t *s, uint32_t bpp)
{
    if (bpp == 4 && (((uintptr_t) s) & 3) == 0)
        return *reinterpret_cast<const uint32_t *> (s);
    uint32_t v = (uint32_t) s[0] | ((uint32_t) s[1] << 8) | ((uint32_t) s[2] << 16);
    if (bpp == 4)
        v |= (uint32_t) s[3] << 24;
    return v;
}

__device__ __forceinline__ void store_raw_px (uint8_t *o, uint32_t v, uint32_t bpp)
{
    if (bpp == 4 && (((uintptr_t) o) & 3) == 0)
    {
        *reinterpret_cast<uint32_t *> (o) = v;
        return;
    }
    o[0] = (uint8_t) v; o[1] = (uint8_t) (v >> 8); o[2] = (uint8_t) (v >> 16);
    if (bpp == 4)
        o[3] = (uint8_t) (v >> 24);
}

/* Cold path: n_px (< 4 or unaligned destination) pixels stored one at a time.  The loop is kept
 * rolled on purpose: a rolled loop stays a real branch, so the hot path does not have to issue its
 * instructions predicated off. */
__device__ __forceinline__ void store_px_slow (uint8_t *dst, const uint32_t out[4], uint32_t n_px, uint32_t bpp)
{
#pragma unroll 1
    for (uint32_t o = 0; o < n_px; o++)
    {
        const uint32_t v = o == 0 ? out[0] : o == 1 ? out[1] : o == 2 ? out[2] : out[3];
        store_raw_px (dst + (size_t) bpp * o, v, bpp);
    }
}

/* Four finished pixels to a 4-byte-aligned destination row: one 128-bit store when the address
 * allows, else 32-bit stores (32bpp); three 32-bit stores (24bpp). */
template <int BO>
__device__ __forceinline__ void store_px4_aligned (uint8_t *dst, const uint32_t out[4])
{
    if constexpr (BO == 4)
    {
        if ((reinterpret_cast<uintptr_t> (dst) & 15) == 0)
            *reinterpret_cast<uint4 *> (dst) = make_uint4 (out[0], out[1], out[2], out[3]);
        else
        {
            uint32_t *d32 = reinterpret_cast<uint32_t *> (dst);
            d32[0] = out[0]; d32[1] = out[1]; d32[2] = out[2]; d32[3] = out[3];
        }
    }
    else
    {
        uint32_t *d32 = reinterpret_cast<uint32_t *> (dst);
        d32[0] = __byte_perm (out[0], out[1], 0x4210);
        d32[1] = __byte_perm (out[1], out[2], 0x5421);
        d32[2] = __byte_perm (out[2], out[3], 0x6542);
    }
}

/* Four finished 24bpp pixels (12 bytes) of one thread to a destination row at ANY byte alignment
 * (tightly packed RGB rows of odd width: 3 x width is rarely a multiple of four).  All threads of a
 * warp sit in one output row, 12 bytes apart, so the row's misalignment `phi` is warp-uniform and
 * the aligned words of the row interleave the threads' bytes: a thread writes the three aligned
 * words that start inside its own 12 bytes, taking the first (4 - phi) bytes of the word that
 * straddles into its right-hand neighbour from that neighbour by shuffle.  Only a warp's first
 * and last lanes (and the row's ragged end) touch single bytes.  `mask`: the warp's live lanes
 * (they never change inside the row walk); `has_next`: the lane to the right is live and owns at
 * least one pixel; n_px: valid pixels of this thread (1..4). */
__device__ __forceinline__ void
store_px4_rgb_anywhere (uint8_t *dst, const uint32_t out[4], uint32_t n_px, unsigned mask, bool has_next)
{
    const uint32_t o0 = __byte_perm (out[0], out[1], 0x4210), o1 = __byte_perm (out[1], out[2], 0x5421),
                   o2 = __byte_perm (out[2], out[3], 0x6542);
    const uint32_t phi = (uint32_t) reinterpret_cast<uintptr_t> (dst) & 3u;
    const uint32_t nxt = __shfl_down_sync (mask, o0, 1);
    const bool has_prev = (threadIdx.x & 31) != 0;      /* the lane to the left writes this thread's first 4 - phi bytes */

    if (phi == 0 && n_px == 4)
    {
        uint32_t *d32 = reinterpret_cast<uint32_t *> (dst);
        d32[0] = o0; d32[1] = o1; d32[2] = o2;
        return;
    }
    const uint32_t lead = phi ? 4u - phi : 0u;          /* bytes before this thread's first aligned word */
    auto byte_at = [&] (uint32_t j) -> uint8_t
    {
        const uint32_t w = j < 4 ? o0 : j < 8 ? o1 : o2;
        return (uint8_t) (w >> ((j & 3) * 8));
    };
    if (n_px == 4 && phi != 0)
    {
        const uint32_t sh = lead * 8;
        uint32_t *d32 = reinterpret_cast<uint32_t *> (dst + lead);
        d32[0] = __funnelshift_r (o0, o1, sh);
        d32[1] = __funnelshift_r (o1, o2, sh);
        if (has_next)
            d32[2] = __funnelshift_r (o2, nxt, sh);
        else
        {
#pragma unroll 1
            for (uint32_t j = 8 + lead; j < 12; j++)
                dst[j] = byte_at (j);
        }
        if (!has_prev)
        {
#pragma unroll 1
            for (uint32_t j = 0; j < lead; j++)
                dst[j] = byte_at (j);
        }
        return;
    }
    /* ragged end of the row (or an aligned row's short last thread): bytes, minus what the left lane wrote */
#pragma unroll 1
    for (uint32_t j = (has_prev && phi != 0) ? lead : 0u; j < 3 * n_px; j++)
        dst[j] = byte_at (j);
}

__device__ __forceinline__ uint4 ldg_nc_v4 (const void *p)
{
    uint4 r;
    asm volatile ("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                  : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

/* Cooperative copy of global bytes [gbeg, gend) into shared memory such that the byte at global
 * address A lands at sm[A - (gbeg & ~15)]: alignment (mod 16) is preserved, whole 16-byte
 * chunks move as one 128-bit load, and no byte outside [gbeg, gend) is touched. */
__device__ __forceinline__ void stage_bytes (uint8_t *sm, const uint8_t *gbeg, const uint8_t *gend)
{
    const uint8_t *abase = reinterpret_cast<const uint8_t *> (reinterpret_cast<uintptr_t> (gbeg) & ~(uintptr_t) 15);
    const uint32_t n_chunks = (uint32_t) ((gend - abase + 15) >> 4);

    for (uint32_t k = threadIdx.x; k < n_chunks; k += blockDim.x)
    {
        const uint8_t *ca = abase + 16 * (size_t) k;
        if (ca >= gbeg && ca + 16 <= gend)
        {
            *reinterpret_cast<uint4 *> (sm + 16 * (size_t) k) = ldg_nc_v4 (ca);
        }
        else
        {
            for (int b = 0; b < 16; b++)
                if (ca + b >= gbeg && ca + b < gend)
                    sm[16 * (size_t) k + b] = ca[b];
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * General kernel: any filter pair, any format.                                               *
 *                                                                                            *
 * A CTA owns TW output columns x rows_per_cta output rows.  `lanes_per_col` (G) threads       *
 * cooperate on one output column (G > 1 only for wide box spans; the G partial sums are       *
 * combined with warp shuffles).  The CTA walks its output rows top to bottom; for every       *
 * source row it needs it stages the row segment [sx0, sx1] into shared memory with coalesced  *
 * 128-bit loads, every thread horizontally filters its own column from there, and the result  *
 * goes straight into the thread's vertical accumulator (box) or its two-row register cache    *
 * (taps -- the device analogue of the reference's SmolVerticalCtx, generic:1648-1682).        *
 * ------------------------------------------------------------------------------------------ */

template <bool S128, bool HBOX>
struct HState
{
    uint32_t x;            /* output column (clamped) */
    uint32_t g, G;         /* lane within the column group */
    uint32_t sx0, sx1;     /* CTA source column range, inclusive */
    /* box */
    uint32_t hL, hR, wl, wr;
};

template <bool S128, bool HBOX>
__device__ __forceinline__ Px<S128>
hfilter_row (const SmolLaunch &L, const uint8_t *__restrict__ src_row, uint8_t *sm_raw,
             const HState<S128, HBOX> &hs)
{
    const SmolJobDesc &d = L.d;
    const uint32_t bpp = d.bpp_in;
    Px<S128> acc = px_zero<S128> ();

    for (uint32_t c0 = hs.sx0; c0 <= hs.sx1; c0 += L.chunk_px)
    {
        const uint32_t c1 = min (c0 + L.chunk_px, hs.sx1 + 1);                 /* exclusive */
        const uint32_t c1s = HBOX ? c1 : min (c1 + 1, hs.sx1 + 1);             /* taps read one pixel past */
        const uint8_t *gbeg = src_row + (size_t) c0 * bpp;
        const uint8_t *gend = src_row + (size_t) c1s * bpp;

        __syncthreads ();
        stage_bytes (sm_raw, gbeg, gend);
        __syncthreads ();

        const uint32_t sm_ofs = (uint32_t) (reinterpret_cast<uintptr_t> (gbeg) & 15);

        if constexpr (HBOX)
        {
            const uint32_t lo = max (hs.hL, c0), hi = min (hs.hR, c1 - 1);
            uint32_t j = hs.hL + hs.g;
            if (j < lo)
                j += ((lo - j + hs.G - 1) / hs.G) * hs.G;
            for (; j <= hi; j += hs.G)
            {
                Px<S128> p = unpack_px<S128> (load_raw_px (sm_raw + sm_ofs + (size_t) (j - c0) * bpp, bpp), d, L.luts);
                if (j == hs.hL)
                    p = px_weight<S128> (p, hs.wl);
                else if (j == hs.hR)
                    p = px_weight<S128> (p, hs.wr);
                px_add<S128> (acc, p);
            }
        }
        else
        {
            const uint32_t n = 1u << d.h_halvings;
            for (uint32_t k = 0; k < n; k++)
            {
                const uint32_t e = __ldg (&L.tab_x[(hs.x << d.h_halvings) + k]);
                const uint32_t ofs = SMOL_TAB_OFS (e), F = SMOL_TAB_F (e);
                if (ofs >= c0 && ofs < c1)
                {
                    const uint32_t ofs2 = min (ofs + 1, d.w_in - 1);
                    Px<S128> p = unpack_px<S128> (load_raw_px (sm_raw + sm_ofs + (size_t) (ofs - c0) * bpp, bpp), d, L.luts);
                    Px<S128> q = unpack_px<S128> (load_raw_px (sm_raw + sm_ofs + (size_t) (ofs2 - c0) * bpp, bpp), d, L.luts);
                    px_add<S128> (acc, px_lerp<S128> (p, q, F));
                }
            }
        }
    }

    if constexpr (HBOX)
    {
        for (uint32_t m = hs.G >> 1; m; m >>= 1)
            acc = px_shfl_xor_add<S128> (acc, m);
        return px_box_scale<S128> (acc, d.span_mul_x);
    }
    else
    {
        return px_halve<S128> (acc, d.h_halvings);
    }
}

template <bool S128, bool HBOX, bool VBOX>
__global__ void __launch_bounds__ (SMOL_BLOCK)
smol_general_kernel (const SmolLaunch L)
{
    extern __shared__ __align__ (16) uint8_t sm_raw[];

    const SmolJobDesc &d = L.d;
    const uint32_t G = L.lanes_per_col;
    const uint32_t TW = blockDim.x / G;
    const uint32_t x0 = blockIdx.x * TW;
    const uint32_t xlast = min (x0 + TW, d.w_out) - 1;
    const uint8_t *src = L.src + (size_t) blockIdx.z * L.src_image_stride;
    uint8_t *dst = L.dst + (size_t) blockIdx.z * L.dst_image_stride;

    HState<S128, HBOX> hs;
    hs.G = G;
    hs.g = threadIdx.x & (G - 1);
    hs.x = x0 + threadIdx.x / G;
    const bool active = hs.x < d.w_out && hs.g == 0;
    if (hs.x >= d.w_out)
        hs.x = d.w_out - 1;

    if constexpr (HBOX)
    {
        const uint32_t e0 = __ldg (&L.tab_x[hs.x]), e1 = __ldg (&L.tab_x[hs.x + 1]);
        hs.hL = SMOL_TAB_OFS (e0);
        hs.hR = SMOL_TAB_OFS (e1);
        hs.wr = SMOL_TAB_F (e0);
        hs.wl = (hs.x == 0) ? 256u : 255u - SMOL_TAB_F (__ldg (&L.tab_x[hs.x - 1]));
        hs.sx0 = SMOL_TAB_OFS (__ldg (&L.tab_x[x0]));
        hs.sx1 = SMOL_TAB_OFS (__ldg (&L.tab_x[xlast + 1]));
    }
    else
    {
        hs.hL = hs.hR = hs.wl = hs.wr = 0;
        hs.sx0 = SMOL_TAB_OFS (__ldg (&L.tab_x[x0 << d.h_halvings]));
        hs.sx1 = min (SMOL_TAB_OFS (__ldg (&L.tab_x[((xlast + 1) << d.h_halvings) - 1])) + 1, d.w_in - 1);
    }

    const uint32_t y_begin = L.first_row + blockIdx.y * L.rows_per_cta;
    const uint32_t y_end = min (y_begin + L.rows_per_cta, L.first_row + L.n_rows);

    /* two-row cache for the vertical taps */
    uint32_t idx0 = 0xffffffffu, idx1 = 0xffffffffu;
    Px<S128> row0 = px_zero<S128> (), row1 = px_zero<S128> ();

    for (uint32_t y = y_begin; y < y_end; y++)
    {
        Px<S128> out;

        if constexpr (VBOX)
        {
            /* generic:2112-2161 (64bpp) and :2198-2260 (128bpp) */
            const uint32_t e0 = __ldg (&L.tab_y[y]), e1 = __ldg (&L.tab_y[y + 1]);
            const uint32_t T = SMOL_TAB_OFS (e0), B = SMOL_TAB_OFS (e1), Fy = SMOL_TAB_F (e0);
            const uint32_t w1 = (y == 0) ? 256u : 255u - SMOL_TAB_F (__ldg (&L.tab_y[y - 1]));

            Px<S128> acc = px_weight<S128> (hfilter_row<S128, HBOX> (L, src + (size_t) T * L.src_pitch, sm_raw, hs), w1);
            for (uint32_t r = T + 1; r < B; r++)
                px_add<S128> (acc, hfilter_row<S128, HBOX> (L, src + (size_t) r * L.src_pitch, sm_raw, hs));
            if (Fy > 0)
            {
                /* 128bpp weighs the trailing row by F - 1, 64bpp by F (generic:2247-2249 vs :2129-2137) */
                const uint32_t w2 = S128 ? Fy - 1 : Fy;
                px_add<S128> (acc, px_weight<S128> (hfilter_row<S128, HBOX> (L, src + (size_t) B * L.src_pitch, sm_raw, hs), w2));
            }
            out = px_box_scale<S128> (acc, d.span_mul_y);
        }
        else
        {
            const uint32_t n = 1u << d.v_halvings;
            Px<S128> acc = px_zero<S128> ();

            for (uint32_t k = 0; k < n; k++)
            {
                const uint32_t e = __ldg (&L.tab_y[(y << d.v_halvings) + k]);
                const uint32_t r0 = SMOL_TAB_OFS (e), F = SMOL_TAB_F (e);
                const uint32_t r1 = min (r0 + 1, d.h_in - 1);

                /* F == 256 needs only the top row, F == 0 only the bottom one */
                if (F != 0)
                {
                    if (r0 == idx1)
                    {
                        Px<S128> t = row0; row0 = row1; row1 = t;
                        uint32_t ti = idx0; idx0 = idx1; idx1 = ti;
                    }
                    else if (r0 != idx0)
                    {
                        row0 = hfilter_row<S128, HBOX> (L, src + (size_t) r0 * L.src_pitch, sm_raw, hs);
                        idx0 = r0;
                    }
                }
                if (F != 256)
                {
                    if (r1 != idx1)
                    {
                        if (r1 == idx0)
                            row1 = row0;
                        else
                            row1 = hfilter_row<S128, HBOX> (L, src + (size_t) r1 * L.src_pitch, sm_raw, hs);
                        idx1 = r1;
                    }
                }
                Px<S128> v;
                if (F == 256)
                    v = px_halve<S128> (row0, 0);
                else if (F == 0)
                    v = px_halve<S128> (row1, 0);
                else
                    v = px_lerp<S128> (row0, row1, F);
                px_add<S128> (acc, v);
            }
            out = px_halve<S128> (acc, d.v_halvings);
        }

        if (active)
        {
            uint8_t *o = dst + (size_t) (y - L.first_row) * L.dst_pitch + (size_t) hs.x * d.bpp_out;
            store_raw_px (o, pack_px<S128> (out, d, L.luts), d.bpp_out);
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * "half" kernel: exact 2^k : 1 reductions on both axes (every bilinear weight is 128 and sample *
 * i reads pixels 2i, 2i + 1 -- BASELINE configs 1, 2 and 5), 32bpp premultiplied or alpha-less  *
 * source, 8-bit premultiplied intermediate.                                                    *
 *                                                                                              *
 * With F = 128 the reference's tap ((p - q) * 128 >> 8) + q is, lane by lane, floor ((p + q)/2) *
 * (SURVEY A.7), so all the filtering can be done on the packed source bytes without unpacking   *
 * to 16-bit lanes: one output pixel of a 2:1 job is three byte-wise floor averages.  Halvings   *
 * (sum of 2^n such samples, then >> n) are accumulated in 16-bit lanes.  A thread produces four *
 * adjacent output pixels: 128-bit loads of the source rows, one 128-bit store.                  *
 * ------------------------------------------------------------------------------------------ */

/* Programmatic dependent launch (PDL): our kernels are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, let the next kernel in the stream start
 * its CTAs early (pdl_launch_dependents) and wait for the previous kernel's memory to be
 * complete and visible before touching any user buffer (pdl_wait).  Stream order is preserved
 * exactly; only launch latency and prologue overlap the predecessor's tail. */
__device__ __forceinline__ void pdl_launch_dependents ()
{
    asm volatile ("griddepcontrol.launch_dependents;");
}

__device__ __forceinline__ void pdl_wait ()
{
    asm volatile ("griddepcontrol.wait;" ::: "memory");
}

/* L2 prefetch of a line this thread is going to read.  Issued BEFORE the dependency wait: a
 * prefetch has no functional effect (L2 is the coherence point, so data the previous grid still
 * has to write lands in the same L2 line), but it lets a grid that became resident early pull
 * its first source rows out of HBM while the previous grid drains -- the launch-to-launch bubble
 * of per-frame calls is otherwise idle HBM time.  SMOL_PDL_PREFETCH=0 turns it off (a launcher
 * parameter, for measurements). */
__device__ __forceinline__ void prefetch_l2 (const void *p)
{
    asm volatile ("prefetch.global.L2 [%0];" :: "l"(p));
}

__device__ __forceinline__ void prefetch_l1 (const void *p)
{
    asm volatile ("prefetch.global.L1 [%0];" :: "l"(p));
}

/* Only CTAs of the grid's first wave can be resident before the previous grid has finished; for
 * the rest a prefetch is pure overhead.  `first_wave` = how many CTAs (in launch order) that is. */
__device__ __forceinline__ bool in_first_wave (uint32_t first_wave)
{
    return blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) < first_wave;
}

/* shared-memory reads by 32-bit window address (LDS [R]: no generic-pointer arithmetic) */
__device__ __forceinline__ uint32_t lds_u32 (uint32_t addr)
{
    uint32_t v;
    asm ("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u64 (uint32_t addr)
{
    uint2 v;
    asm ("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
/* staging-buffer read: volatile, so it keeps its place among the cp.async waits and warp barriers
 * around it (all volatile asm); no "memory" clobber, which would make the compiler re-read kernel
 * parameters (the PRMT selectors) after every pixel */
__device__ __forceinline__ uint32_t lds_u32_ordered (uint32_t addr)
{
    uint32_t v;
    asm volatile ("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint2 lds_u64_ordered (uint32_t addr)
{
    uint2 v;
    asm volatile ("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t byte_avg_floor (uint32_t a, uint32_t b)
{
    /* per byte: floor ((a + b) / 2) without carries between bytes */
    return (a & b) + (((a ^ b) & 0xfefefefeu) >> 1);
}

struct HalfParams
{
    const uint8_t *src; uint8_t *dst;
    uint32_t src_pitch, dst_pitch;
    size_t src_image_stride, dst_image_stride;
    uint32_t w_out, first_row, n_rows;
    uint32_t items_per_row;         /* ceil (w_out / 4) */
    uint32_t prmt_sel;              /* source byte order -> destination byte order */
    const uint32_t *inv_div_p8;     /* device LUT */
    uint32_t prefetch;              /* CTAs (in launch order) that L2-prefetch their source on entry, ahead of the dependency wait */
};

/* Unpremultiply one packed pixel, kept in source byte order (reference generic:246-259 via
 * :892-901).  sm_inv holds inv_div_p8 << 3, so ((c * inv) >> 13) & 0xff is byte 2 of the 32-bit
 * product (c <= 255 and inv < 2^21 keep it below 2^32). */
template <bool ALPHA_FIRST>
__device__ __forceinline__ uint32_t half_unpremul (uint32_t v, const uint32_t *__restrict__ sm_inv)
{
    if constexpr (ALPHA_FIRST)
    {
        /* bytes: a c0 c1 c2 */
        const uint32_t inv8 = sm_inv[v & 0xff];
        const uint32_t p0 = __byte_perm (v, 0, 0x4441) * inv8;
        const uint32_t p1 = __byte_perm (v, 0, 0x4442) * inv8;
        const uint32_t p2 = (v >> 24) * inv8;
        const uint32_t t = __byte_perm (v, p0, 0x4460);        /* a, p0.b2 */
        const uint32_t u = __byte_perm (p1, p2, 0x4462);       /* p1.b2, p2.b2 */
        return __byte_perm (t, u, 0x5410);
    }
    else
    {
        /* bytes: c0 c1 c2 a */
        const uint32_t inv8 = sm_inv[v >> 24];
        const uint32_t p0 = (v & 0xff) * inv8;
        const uint32_t p1 = __byte_perm (v, 0, 0x4441) * inv8;
        const uint32_t p2 = __byte_perm (v, 0, 0x4442) * inv8;
        const uint32_t t = __byte_perm (p0, p1, 0x4462);       /* p0.b2, p1.b2 */
        const uint32_t u = __byte_perm (p2, v, 0x4472);        /* p2.b2, a */
        return __byte_perm (t, u, 0x5410);
    }
}

/* Horizontally reduced value of one output pixel on one source row, as packed bytes.
 * px points at the 2 << HH source pixels of this output pixel (already loaded). */
template <int HH>
__device__ __forceinline__ uint32_t half_hreduce (const uint32_t *px)
{
    if constexpr (HH == 0)
    {
        return byte_avg_floor (px[0], px[1]);
    }
    else
    {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < (1 << HH); k++)
        {
            const uint32_t v = byte_avg_floor (px[2 * k], px[2 * k + 1]);
            lo += v & 0x00ff00ffu;
            hi += (v >> 8) & 0x00ff00ffu;
        }
        lo = (lo >> HH) & 0x00ff00ffu;
        hi = (hi >> HH) & 0x00ff00ffu;
        return lo | (hi << 8);
    }
}

/* AL: what the buffers are aligned to -- 16 bytes (the fast case), 8 (a tightly packed 32bpp image
 * with an odd number of output columns) or 4 (views into larger images): 16-byte accesses become
 * two 64-bit or four 32-bit ones.  The narrow loads allocate in L1 -- a warp still covers a
 * contiguous run of the row, each sector is fetched from L2 once and its other words hit L1. */
/* `interior`: the four bytes before p and the four after p + 16 belong to the same source row (any
 * chunk but a row's first and last), so a chunk that is only 4-byte aligned may be assembled from
 * the three aligned 64-bit words around it instead of four 32-bit ones. */
template <int AL>
__device__ __forceinline__ uint4 half_load16 (const uint8_t *p, bool interior = false)
{
    if constexpr (AL == 4)
    {
        if ((reinterpret_cast<uintptr_t> (p) & 7) == 0)
        {
            const uint2 *w = reinterpret_cast<const uint2 *> (p);
            const uint2 a = __ldg (w), b = __ldg (w + 1);
            return make_uint4 (a.x, a.y, b.x, b.y);
        }
        if (interior)
        {
            const uint2 *w = reinterpret_cast<const uint2 *> (p - 4);
            const uint2 a = __ldg (w), b = __ldg (w + 1), c = __ldg (w + 2);
            return make_uint4 (a.y, b.x, b.y, c.x);
        }
        const uint32_t *w = reinterpret_cast<const uint32_t *> (p);
        return make_uint4 (__ldg (w), __ldg (w + 1), __ldg (w + 2), __ldg (w + 3));
    }
    else if constexpr (AL == 8)
    {
        const uint2 *w = reinterpret_cast<const uint2 *> (p);
        const uint2 a = __ldg (w), b = __ldg (w + 1);
        return make_uint4 (a.x, a.y, b.x, b.y);
    }
    else
        return ldg_nc_v4 (p);
}

template <int AL>
__device__ __forceinline__ void half_store16 (uint8_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    if constexpr (AL == 4)
    {
        uint32_t *w = reinterpret_cast<uint32_t *> (p);
        w[0] = a; w[1] = b; w[2] = c; w[3] = d;
    }
    else if constexpr (AL == 8)
    {
        uint2 *w = reinterpret_cast<uint2 *> (p);
        w[0] = make_uint2 (a, b); w[1] = make_uint2 (c, d);
    }
    else
        *reinterpret_cast<uint4 *> (p) = make_uint4 (a, b, c, d);
}

/* PACK: 0 = byte permutation only, 1 = unpremultiply with alpha in byte 3, 2 = ... in byte 0.
 * Block = (bx, by) threads; thread (tx, ty) of CTA (cx, cy, image) produces output pixels
 * 4 * (cx * bx + tx) .. + 3 of output row first_row + cy * by + ty. */
template <int HH, int VH, int PACK, int AL>
__global__ void __launch_bounds__ (256)
smol_half_kernel (const HalfParams P)
{
    __shared__ uint32_t sm_inv[256];

    pdl_launch_dependents ();
    constexpr int SRC_PER_OUT = 2 << HH;            /* source pixels per output pixel per row */
    constexpr int VEC_PER_OUT = SRC_PER_OUT / 4 > 0 ? SRC_PER_OUT / 4 : 1;

    const uint32_t xi = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t yl = blockIdx.y * blockDim.y + threadIdx.y;
    if (P.prefetch && xi < P.items_per_row
        && (SRC_PER_OUT * 16 >= 128 || (threadIdx.x & (128 / (SRC_PER_OUT * 16) - 1)) == 0))
    {
        /* L2 prefetch (see prefetch_l2), one lane per 128-byte line */
        if (yl < P.n_rows && in_first_wave (P.prefetch))
        {
            const uint8_t *p = P.src + (size_t) blockIdx.z * P.src_image_stride
                               + (size_t) (2 * ((P.first_row + yl) << VH)) * P.src_pitch + (size_t) xi * (SRC_PER_OUT * 16);
#pragma unroll
            for (int r = 0; r < (2 << VH); r++)
#pragma unroll
                for (int b = 0; b < SRC_PER_OUT * 16; b += 128)
                    prefetch_l2 (p + (size_t) r * P.src_pitch + b);
        }
    }
    if constexpr (PACK != 0)
    {
        /* the LUT is library-owned constant data: safe to read before the dependency wait */
        for (uint32_t i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y)
            sm_inv[i] = __ldg (&P.inv_div_p8[i]) << 3;
        __syncthreads ();
    }
    pdl_wait ();

    const bool live = xi < P.items_per_row && yl < P.n_rows;
    const uint32_t x = xi * 4;
    const uint32_t y = P.first_row + yl;
    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl * P.dst_pitch + (size_t) x * 4;
    const uint32_t n_px = live ? min (4u, P.w_out - x) : 0u;
    uint32_t out[4] = { 0, 0, 0, 0 };      /* filtered pixels, source byte order */

    if (n_px == 4)
    {
        uint32_t acc_lo[4] = { 0, 0, 0, 0 }, acc_hi[4] = { 0, 0, 0, 0 };

#pragma unroll
        for (int kv = 0; kv < (1 << VH); kv++)
        {
            const uint32_t r = 2 * ((y << VH) + kv);
            const uint8_t *row0 = src + (size_t) r * P.src_pitch + (size_t) x * (SRC_PER_OUT * 4);
            const uint8_t *row1 = row0 + P.src_pitch;
            uint32_t h0[4], h1[4];

            if constexpr (HH == 0)
            {
                /* 4 output pixels = 8 source pixels = two 128-bit loads per row */
                /* (x is a multiple of 4: this thread's 32 source bytes start the row iff x == 0 and end it iff x + 4 == w_out) */
                const bool in0 = x > 0, in1 = x + 4 < P.w_out;
                const uint4 a0 = half_load16<AL> (row0, in0), a1 = half_load16<AL> (row0 + 16, in1);
                const uint4 b0 = half_load16<AL> (row1, in0), b1 = half_load16<AL> (row1 + 16, in1);
                h0[0] = byte_avg_floor (a0.x, a0.y); h0[1] = byte_avg_floor (a0.z, a0.w);
                h0[2] = byte_avg_floor (a1.x, a1.y); h0[3] = byte_avg_floor (a1.z, a1.w);
                h1[0] = byte_avg_floor (b0.x, b0.y); h1[1] = byte_avg_floor (b0.z, b0.w);
                h1[2] = byte_avg_floor (b1.x, b1.y); h1[3] = byte_avg_floor (b1.z, b1.w);
            }
            else
            {
#pragma unroll
                for (int o = 0; o < 4; o++)
                {
                    uint32_t pa[SRC_PER_OUT], pb[SRC_PER_OUT];
#pragma unroll
                    for (int v = 0; v < VEC_PER_OUT; v++)
                    {
                        const uint4 a = half_load16<AL> (row0 + 16 * (o * VEC_PER_OUT + v));
                        const uint4 b = half_load16<AL> (row1 + 16 * (o * VEC_PER_OUT + v));
                        pa[4 * v] = a.x; pa[4 * v + 1] = a.y; pa[4 * v + 2] = a.z; pa[4 * v + 3] = a.w;
                        pb[4 * v] = b.x; pb[4 * v + 1] = b.y; pb[4 * v + 2] = b.z; pb[4 * v + 3] = b.w;
                    }
                    h0[o] = half_hreduce<HH> (pa);
                    h1[o] = half_hreduce<HH> (pb);
                }
            }

#pragma unroll
            for (int o = 0; o < 4; o++)
            {
                const uint32_t v = byte_avg_floor (h0[o], h1[o]);
                if constexpr (VH == 0)
                    out[o] = v;
                else
                {
                    acc_lo[o] += v & 0x00ff00ffu;
                    acc_hi[o] += (v >> 8) & 0x00ff00ffu;
                }
            }
        }

        if constexpr (VH > 0)
        {
#pragma unroll
            for (int o = 0; o < 4; o++)
                out[o] = ((acc_lo[o] >> VH) & 0x00ff00ffu) | (((acc_hi[o] >> VH) & 0x00ff00ffu) << 8);
        }
    }
    else
    {
        /* ragged end of a row: one pixel at a time, same arithmetic */
#pragma unroll
        for (int o = 0; o < 3; o++)
        {
            if ((uint32_t) o >= n_px)
                continue;
            uint32_t acc_lo = 0, acc_hi = 0, res = 0;

            for (int kv = 0; kv < (1 << VH); kv++)
            {
                const uint32_t r = 2 * ((y << VH) + kv);
                const uint32_t *row0 = reinterpret_cast<const uint32_t *> (src + (size_t) r * P.src_pitch)
                                       + (size_t) (x + o) * SRC_PER_OUT;
                const uint32_t *row1 = reinterpret_cast<const uint32_t *> (src + (size_t) (r + 1) * P.src_pitch)
                                       + (size_t) (x + o) * SRC_PER_OUT;
                uint32_t pa[SRC_PER_OUT], pb[SRC_PER_OUT];
#pragma unroll
                for (int k = 0; k < SRC_PER_OUT; k++)
                {
                    pa[k] = __ldg (row0 + k);
                    pb[k] = __ldg (row1 + k);
                }
                const uint32_t v = byte_avg_floor (half_hreduce<HH> (pa), half_hreduce<HH> (pb));
                if (VH == 0)
                    res = v;
                else
                {
                    acc_lo += v & 0x00ff00ffu;
                    acc_hi += (v >> 8) & 0x00ff00ffu;
                }
            }
            if (VH > 0)
                res = ((acc_lo >> VH) & 0x00ff00ffu) | (((acc_hi >> VH) & 0x00ff00ffu) << 8);
            out[o] = res;
        }
    }

    if (n_px == 0)
        return;

#pragma unroll
    for (int o = 0; o < 4; o++)
    {
        uint32_t v = out[o];
        if constexpr (PACK == 1)
            v = half_unpremul<false> (v, sm_inv);
        else if constexpr (PACK == 2)
            v = half_unpremul<true> (v, sm_inv);
        out[o] = __byte_perm (v, 0, P.prmt_sel);
    }
    if (n_px == 4)
        half_store16<AL> (dst, out[0], out[1], out[2], out[3]);
    else
    {
#pragma unroll
        for (int o = 0; o < 3; o++)
            if ((uint32_t) o < n_px)
                reinterpret_cast<uint32_t *> (dst)[o] = out[o];
    }
}

/* 2:1 on both axes (BASELINE cfg 1, 2) with 32-byte-aligned source rows: the same arithmetic as
 * smol_half_kernel<0, 0, PACK> on 256-bit loads.  A thread's four output pixels are eight source
 * pixels = ONE 32-byte sector per source row, so a warp instruction reads 1 KB of contiguous
 * bytes with every sector requested exactly once (the 128-bit form touches each sector from two
 * instructions, and with L1 allocation off both go to L2).  A thread walks ROWS output rows with
 * all its loads issued up front (ROWS x 64 bytes in flight); the per-CTA set-up -- staging the
 * inverse-division table with a few vector loads, the PDL prologue -- is spread over ROWS times
 * as much output.  Fewer, wider instructions per byte also matter for the sustained rate: under
 * a long run the board is power-limited and SM clocks drop with issue activity. */
__device__ __forceinline__ void ldg_nc_v8 (const void *p, uint4 &lo, uint4 &hi)
{
    asm volatile ("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                  : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p));
}

template <int PACK, int ROWS>
__global__ void __launch_bounds__ (256)
smol_half2v_kernel (const HalfParams P)
{
    __shared__ __align__ (16) uint32_t sm_inv[256];

    pdl_launch_dependents ();
    const uint32_t xi = blockIdx.x * blockDim.x + threadIdx.x;                  /* group of 4 output pixels */
    const uint32_t yl0 = (blockIdx.y * blockDim.y + threadIdx.y) * ROWS;        /* first of this thread's output rows */
    const bool full = xi * 4 + 4 <= P.w_out;
    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride
                         + (size_t) (2 * (P.first_row + yl0)) * P.src_pitch + (size_t) xi * 32;

    if (P.prefetch && xi * 4 < P.w_out && (threadIdx.x & 3) == 0 && in_first_wave (P.prefetch))
    {
        /* L2 prefetch ahead of the dependency wait (see prefetch_l2), one lane per 128-byte line */
#pragma unroll
        for (int r = 0; r < 2 * ROWS; r++)
            if (yl0 + r / 2 < P.n_rows)
                prefetch_l2 (src + (size_t) r * P.src_pitch);
    }
    if constexpr (PACK != 0)
    {
        /* library-owned constant data: safe to read before the dependency wait */
        /* (a block can be as small as one warp: a one-row job) */
        for (uint32_t t = threadIdx.y * blockDim.x + threadIdx.x; t < 64; t += blockDim.x * blockDim.y)
        {
            uint4 v = __ldg (reinterpret_cast<const uint4 *> (P.inv_div_p8) + t);
            v.x <<= 3; v.y <<= 3; v.z <<= 3; v.w <<= 3;
            reinterpret_cast<uint4 *> (sm_inv)[t] = v;
        }
        __syncthreads ();
    }
    pdl_wait ();
    if (xi * 4 >= P.w_out || yl0 >= P.n_rows)
        return;

    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl0 * P.dst_pitch + (size_t) xi * 16;

    if (full)
    {
        uint4 a_lo[ROWS], a_hi[ROWS], b_lo[ROWS], b_hi[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++)
            if (yl0 + r < P.n_rows)
            {
                ldg_nc_v8 (src + (size_t) (2 * r) * P.src_pitch, a_lo[r], a_hi[r]);
                ldg_nc_v8 (src + (size_t) (2 * r + 1) * P.src_pitch, b_lo[r], b_hi[r]);
            }
#pragma unroll
        for (int r = 0; r < ROWS; r++)
            if (yl0 + r < P.n_rows)
            {
                uint32_t out[4];
                out[0] = byte_avg_floor (byte_avg_floor (a_lo[r].x, a_lo[r].y), byte_avg_floor (b_lo[r].x, b_lo[r].y));
                out[1] = byte_avg_floor (byte_avg_floor (a_lo[r].z, a_lo[r].w), byte_avg_floor (b_lo[r].z, b_lo[r].w));
                out[2] = byte_avg_floor (byte_avg_floor (a_hi[r].x, a_hi[r].y), byte_avg_floor (b_hi[r].x, b_hi[r].y));
                out[3] = byte_avg_floor (byte_avg_floor (a_hi[r].z, a_hi[r].w), byte_avg_floor (b_hi[r].z, b_hi[r].w));
#pragma unroll
                for (int o = 0; o < 4; o++)
                {
                    uint32_t v = out[o];
                    if constexpr (PACK == 1)
                        v = half_unpremul<false> (v, sm_inv);
                    else if constexpr (PACK == 2)
                        v = half_unpremul<true> (v, sm_inv);
                    out[o] = __byte_perm (v, 0, P.prmt_sel);
                }
                *reinterpret_cast<uint4 *> (dst + (size_t) r * P.dst_pitch) = make_uint4 (out[0], out[1], out[2], out[3]);
            }
    }
    else
    {
        /* ragged end of a row: one pixel at a time, same arithmetic */
        const uint32_t n_px = P.w_out - xi * 4;
        for (int r = 0; r < ROWS && yl0 + r < P.n_rows; r++)
            for (uint32_t o = 0; o < n_px; o++)
            {
                const uint32_t *row0 = reinterpret_cast<const uint32_t *> (src + (size_t) (2 * r) * P.src_pitch) + 2 * o;
                const uint32_t *row1 = reinterpret_cast<const uint32_t *> (src + (size_t) (2 * r + 1) * P.src_pitch) + 2 * o;
                uint32_t v = byte_avg_floor (byte_avg_floor (__ldg (row0), __ldg (row0 + 1)),
                                             byte_avg_floor (__ldg (row1), __ldg (row1 + 1)));
                if constexpr (PACK == 1)
                    v = half_unpremul<false> (v, sm_inv);
                else if constexpr (PACK == 2)
                    v = half_unpremul<true> (v, sm_inv);
                reinterpret_cast<uint32_t *> (dst + (size_t) r * P.dst_pitch)[o] = __byte_perm (v, 0, P.prmt_sel);
            }
    }
}

/* Variant for 4:1 and 8:1 horizontal reductions (HH = 1, 2).  Here one output pixel spans 16 or
 * 32 source bytes per row, so the thread <-> data mapping is chosen for the loads: lane L of a
 * warp reads the L-th 16-byte chunk of the row segment (perfectly coalesced 512 bytes per
 * instruction, all 2 << VH source rows in flight at once).  With HH = 1 a chunk is exactly one
 * output pixel; with HH = 2 two neighbouring lanes hold the two halves of an output pixel and
 * add their partial sums with one shuffle per word. */
template <int HH, int VH, int PACK, int AL>
__global__ void __launch_bounds__ (256)
smol_half_wide_kernel (const HalfParams P)
{
    static_assert (HH == 1 || HH == 2, "wide variant is for 4:1 and 8:1");
    __shared__ uint32_t sm_inv[256];

    pdl_launch_dependents ();
    constexpr int N_ROWS = 2 << VH;
    const uint32_t cx = blockIdx.x * blockDim.x + threadIdx.x;      /* 16-byte chunk within the row */
    const uint32_t yl = blockIdx.y * blockDim.y + threadIdx.y;
    const uint32_t n_chunks = P.w_out << (HH - 1);
    const bool live = cx < n_chunks && yl < P.n_rows;               /* dead lanes still shuffle */
    const uint32_t y = P.first_row + yl;
    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride
                         + (size_t) (y << (VH + 1)) * P.src_pitch + (size_t) cx * 16;
    if (P.prefetch && cx < n_chunks && (threadIdx.x & 7) == 0)
    {
        /* eight lanes share a 128-byte line: one prefetch per line and row (see prefetch_l2) */
        if (yl < P.n_rows && in_first_wave (P.prefetch))
        {
#pragma unroll
            for (int r = 0; r < N_ROWS; r++)
                prefetch_l2 (src + (size_t) r * P.src_pitch);
        }
    }
    if constexpr (PACK != 0)
    {
        for (uint32_t i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y)
            sm_inv[i] = __ldg (&P.inv_div_p8[i]) << 3;
        __syncthreads ();
    }
    pdl_wait ();

    uint4 rows[N_ROWS];
#pragma unroll
    for (int r = 0; r < N_ROWS; r++)
        rows[r] = live ? half_load16<AL> (src + (size_t) r * P.src_pitch, cx > 0 && cx + 1 < n_chunks) : make_uint4 (0, 0, 0, 0);

    uint32_t acc_lo = 0, acc_hi = 0, res = 0;
#pragma unroll
    for (int kv = 0; kv < (1 << VH); kv++)
    {
        uint32_t h[2];
#pragma unroll
        for (int t = 0; t < 2; t++)
        {
            const uint4 q = rows[2 * kv + t];
            const uint32_t v0 = byte_avg_floor (q.x, q.y), v1 = byte_avg_floor (q.z, q.w);
            uint32_t lo = (v0 & 0x00ff00ffu) + (v1 & 0x00ff00ffu);
            uint32_t hi = ((v0 >> 8) & 0x00ff00ffu) + ((v1 >> 8) & 0x00ff00ffu);
            if constexpr (HH == 2)
            {
                lo += __shfl_xor_sync (0xffffffffu, lo, 1);
                hi += __shfl_xor_sync (0xffffffffu, hi, 1);
            }
            h[t] = ((lo >> HH) & 0x00ff00ffu) | (((hi >> HH) & 0x00ff00ffu) << 8);
        }
        const uint32_t v = byte_avg_floor (h[0], h[1]);
        if constexpr (VH == 0)
            res = v;
        else
        {
            acc_lo += v & 0x00ff00ffu;
            acc_hi += (v >> 8) & 0x00ff00ffu;
        }
    }
    if constexpr (VH > 0)
        res = ((acc_lo >> VH) & 0x00ff00ffu) | (((acc_hi >> VH) & 0x00ff00ffu) << 8);

    if constexpr (PACK == 1)
        res = half_unpremul<false> (res, sm_inv);
    else if constexpr (PACK == 2)
        res = half_unpremul<true> (res, sm_inv);
    res = __byte_perm (res, 0, P.prmt_sel);

    if (live && (HH == 1 || (cx & 1) == 0))
    {
        const uint32_t x = HH == 1 ? cx : cx >> 1;
        uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl * P.dst_pitch;
        reinterpret_cast<uint32_t *> (dst)[x] = res;
    }
}

/* ------------------------------------------------------------------------------------------ *
 * "taps" kernel: bilinear (any weights) / copy / one on both axes, 8-bit premultiplied          *
 * intermediate (reference 64bpp storage).  Registers only, no shared-memory staging.           *
 *                                                                                              *
 * A pixel is two 32-bit words of two 16-bit lanes each (bytes 0, 2 and bytes 1, 3 of the        *
 * packed source pixel), so one weighted tap is two multiply-adds per word:                      *
 *     ((p * F + q * (256 - F)) >> 8) & 0x00ff00ff                                               *
 * A thread owns four adjacent output columns and walks a strip of output rows top to bottom,    *
 * keeping the last two horizontally filtered source rows in registers -- the device analogue    *
 * of the reference's two-row SmolVerticalCtx cache (generic:1648-1682) -- so that on upscales   *
 * each source row is unpacked and filtered once per strip, not once per output row.            *
 * ------------------------------------------------------------------------------------------ */

struct TapsParams
{
    const uint8_t *src; uint8_t *dst;
    uint32_t src_pitch, dst_pitch;
    size_t src_image_stride, dst_image_stride;
    const uint32_t *tab_x, *tab_y;
    const uint32_t *inv_div_p8;
    uint32_t w_in, h_in, w_out;
    uint32_t first_row, n_rows;
    uint32_t rows_per_thread;
    uint32_t bpp_in, bpp_out;
    uint32_t in_alpha_shift;        /* 8 * byte index of alpha in the source pixel (24bpp: 24, byte forced to 0xff) */
    uint32_t in_unassoc, out_unassoc;
    uint32_t prmt_sel;              /* source byte order -> destination byte order */
};

struct Px16 { uint32_t a, b; };     /* a: bytes 0 and 2, b: bytes 1 and 3, one per 16-bit lane */

__device__ __forceinline__ Px16 taps_unpack (uint32_t raw, const TapsParams &P)
{
    Px16 r;
    r.a = raw & 0x00ff00ffu;
    r.b = (raw >> 8) & 0x00ff00ffu;
    if (P.in_unassoc)
    {
        /* premultiply the colour lanes: ((c + 1) * (alpha + 1) - 1) >> 8 (generic:238-244); the
         * alpha lane is cleared first and re-inserted afterwards, as the reference does */
        const uint32_t alpha = (raw >> P.in_alpha_shift) & 0xff;
        const uint32_t m = alpha + 1;
        if (P.in_alpha_shift == 24)
        {
            r.b &= 0x000000ffu;
            r.a = (((r.a + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
            r.b = ((((r.b + 0x00010001u) * m - 0x00010001u) >> 8) & 0x000000ffu) | (alpha << 16);
        }
        else
        {
            r.a &= 0x00ff0000u;
            r.a = ((((r.a + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff0000u) | alpha;
            r.b = (((r.b + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
        }
    }
    return r;
}

__device__ __forceinline__ uint32_t taps_load (const uint8_t *row, uint32_t x, const TapsParams &P)
{
    const uint8_t *p = row + (size_t) x * P.bpp_in;
    if (P.bpp_in == 4)
    {
        if ((reinterpret_cast<uintptr_t> (p) & 3) == 0)
            return __ldg (reinterpret_cast<const uint32_t *> (p));
        return (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16)
               | ((uint32_t) __ldg (p + 3) << 24);
    }
    return (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16) | 0xff000000u;
}

__device__ __forceinline__ uint32_t lerp16 (uint32_t p, uint32_t q, uint32_t F)
{
    return ((p * F + q * (256u - F)) >> 8) & 0x00ff00ffu;
}

/* Horizontally filtered values of the thread's four output columns on source row r. */
template <int HH>
__device__ __forceinline__ void taps_hrow (const TapsParams &P, const uint8_t *src, uint32_t r, uint32_t x, Px16 out[4])
{
    const uint8_t *row = src + (size_t) r * P.src_pitch;
#pragma unroll
    for (int o = 0; o < 4; o++)
    {
        const uint32_t xo = min (x + o, P.w_out - 1);
        uint32_t acc_a = 0, acc_b = 0;
#pragma unroll
        for (int k = 0; k < (1 << HH); k++)
        {
            const uint32_t e = __ldg (&P.tab_x[(xo << HH) + k]);
            const uint32_t ofs = SMOL_TAB_OFS (e), F = SMOL_TAB_F (e);
            const Px16 p = taps_unpack (taps_load (row, ofs, P), P);
            const Px16 q = taps_unpack (taps_load (row, min (ofs + 1, P.w_in - 1), P), P);
            acc_a += lerp16 (p.a, q.a, F);
            acc_b += lerp16 (p.b, q.b, F);
        }
        out[o].a = (acc_a >> HH) & 0x00ff00ffu;
        out[o].b = (acc_b >> HH) & 0x00ff00ffu;
    }
}

template <int HH, int VH>
__global__ void __launch_bounds__ (256)
smol_taps_kernel (const TapsParams P)
{
    __shared__ uint32_t sm_inv[256];

    pdl_launch_dependents ();
    if (P.out_unassoc)
    {
        for (uint32_t i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y)
            sm_inv[i] = __ldg (&P.inv_div_p8[i]) << 3;
        __syncthreads ();
    }
    pdl_wait ();

    const uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const uint32_t strip = blockIdx.y * blockDim.y + threadIdx.y;
    const uint32_t yl0 = strip * P.rows_per_thread;
    if (x >= P.w_out || yl0 >= P.n_rows)
        return;
    const uint32_t yl1 = min (yl0 + P.rows_per_thread, P.n_rows);
    const uint32_t n_px = min (4u, P.w_out - x);
    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    uint8_t *dst_img = P.dst + (size_t) blockIdx.z * P.dst_image_stride;

    uint32_t idx0 = 0xffffffffu, idx1 = 0xffffffffu;
    Px16 row0[4], row1[4];
#pragma unroll
    for (int o = 0; o < 4; o++)
        row0[o].a = row0[o].b = row1[o].a = row1[o].b = 0;

    for (uint32_t yl = yl0; yl < yl1; yl++)
    {
        const uint32_t y = P.first_row + yl;
        uint32_t acc_a[4] = { 0, 0, 0, 0 }, acc_b[4] = { 0, 0, 0, 0 };

#pragma unroll
        for (int kv = 0; kv < (1 << VH); kv++)
        {
            const uint32_t e = __ldg (&P.tab_y[(y << VH) + kv]);
            const uint32_t r0 = SMOL_TAB_OFS (e), F = SMOL_TAB_F (e);
            const uint32_t r1 = min (r0 + 1, P.h_in - 1);

            if (F != 0)
            {
                if (r0 == idx1)
                {
#pragma unroll
                    for (int o = 0; o < 4; o++)
                    {
                        const Px16 t = row0[o]; row0[o] = row1[o]; row1[o] = t;
                    }
                    const uint32_t ti = idx0; idx0 = idx1; idx1 = ti;
                }
                else if (r0 != idx0)
                {
                    taps_hrow<HH> (P, src, r0, x, row0);
                    idx0 = r0;
                }
            }
            if (F != 256 && r1 != idx1)
            {
                if (r1 == idx0)
                {
#pragma unroll
                    for (int o = 0; o < 4; o++)
                        row1[o] = row0[o];
                }
                else
                    taps_hrow<HH> (P, src, r1, x, row1);
                idx1 = r1;
            }
#pragma unroll
            for (int o = 0; o < 4; o++)
            {
                /* F == 256 / F == 0 select one row exactly; the unused operand is multiplied by 0 */
                const uint32_t pa = F != 0 ? row0[o].a : 0, pb = F != 0 ? row0[o].b : 0;
                const uint32_t qa = F != 256 ? row1[o].a : 0, qb = F != 256 ? row1[o].b : 0;
                acc_a[o] += lerp16 (pa, qa, F);
                acc_b[o] += lerp16 (pb, qb, F);
            }
        }

        uint32_t out[4];
#pragma unroll
        for (int o = 0; o < 4; o++)
        {
            const uint32_t a = (acc_a[o] >> VH) & 0x00ff00ffu, b = (acc_b[o] >> VH) & 0x00ff00ffu;
            uint32_t v = a | (b << 8);
            if (P.out_unassoc)
                v = (P.in_alpha_shift == 0) ? half_unpremul<true> (v, sm_inv) : half_unpremul<false> (v, sm_inv);
            out[o] = __byte_perm (v, 0, P.prmt_sel);
        }

        uint8_t *dst = dst_img + (size_t) yl * P.dst_pitch + (size_t) x * P.bpp_out;
        if (P.bpp_out == 4)
        {
            if (n_px == 4 && (reinterpret_cast<uintptr_t> (dst) & 15) == 0)
                *reinterpret_cast<uint4 *> (dst) = make_uint4 (out[0], out[1], out[2], out[3]);
            else
                for (uint32_t o = 0; o < n_px; o++)
                    store_raw_px (dst + 4 * o, out[o], 4);
        }
        else
        {
            if (n_px == 4 && (reinterpret_cast<uintptr_t> (dst) & 3) == 0)
            {
                uint32_t *d32 = reinterpret_cast<uint32_t *> (dst);
                d32[0] = (out[0] & 0x00ffffffu) | (out[1] << 24);
                d32[1] = ((out[1] >> 8) & 0x0000ffffu) | (out[2] << 16);
                d32[2] = ((out[2] >> 16) & 0x000000ffu) | (out[3] << 8);
            }
            else
                for (uint32_t o = 0; o < n_px; o++)
                    store_raw_px (dst + 3 * o, out[o], 3);
        }
    }
}

/* Specialised taps kernel for the no-halving case (HH = VH = 0: every plain bilinear resize
 * between 1:2 and any magnification, and copy / one), formats resolved at compile time:
 * BI / BO bytes per pixel in / out, IU / OU unassociated alpha in / out, AF alpha is byte 0 of a
 * 32bpp source pixel (else byte 3).  Same thread mapping as smol_taps_kernel; the arithmetic
 * uses the PRMT folds: a horizontal tap is 2 multiply-adds + 1 PRMT per word, a vertical tap +
 * repack is 4 multiply-adds + 1 PRMT per pixel. */
struct Taps0Params
{
    TapsParams t;
    uint32_t acc_prmt_sel;          /* (acc_a, acc_b) high bytes -> destination byte order */
    uint32_t src_u32_ok;            /* 32bpp source rows are 4-byte aligned */
    uint32_t prefetch;              /* CTAs (in launch order) that L2-prefetch their first source rows on entry (see prefetch_l2) */
    uint32_t row_ahead;             /* taps0: L1-prefetch the source row this many rows below the one being fetched (0: off) */
};

template <int BI, bool IU, bool AF, bool U32OK>
__device__ __forceinline__ Px16 taps0_fetch (const uint8_t *row, uint32_t x)
{
    const uint8_t *p = row + (size_t) x * BI;
    uint32_t raw;
    if constexpr (BI == 4 && U32OK)
        raw = __ldg (reinterpret_cast<const uint32_t *> (p));
    else
    {
        raw = (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16);
        raw |= BI == 4 ? ((uint32_t) __ldg (p + 3) << 24) : 0xff000000u;
    }
    Px16 r;
    r.a = raw & 0x00ff00ffu;
    r.b = (raw >> 8) & 0x00ff00ffu;
    if constexpr (IU)
    {
        /* premultiply: ((c + 1) * (alpha + 1) - 1) >> 8, alpha lane untouched (generic:238-244) */
        if constexpr (AF)
        {
            const uint32_t alpha = raw & 0xff, m = alpha + 1;
            r.a = (((((r.a & 0x00ff0000u) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff0000u) | alpha;
            r.b = (((r.b + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
        }
        else
        {
            const uint32_t alpha = raw >> 24, m = alpha + 1;
            r.a = (((r.a + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
            r.b = (((((r.b & 0x000000ffu) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x000000ffu) | (alpha << 16);
        }
    }
    return r;
}

/* FASTIO: 32bpp source rows and all destination rows are 4-byte aligned; resolved at compile
 * time so the byte-wise fallbacks cost nothing on the fast path. */
#ifndef SMOL_TAPS0_MINBLOCKS
#define SMOL_TAPS0_MINBLOCKS 5
#endif
/* PX: output pixels per thread.  4 gives 128-bit (or 3 x 32-bit) stores; 1 makes every load of a
 * warp touch one contiguous run of source pixels (32bpp destinations only). */
template <int BI, int BO, bool IU, bool OU, bool AF, bool FASTIO, int PX>
__global__ void __launch_bounds__ (256, SMOL_TAPS0_MINBLOCKS)
smol_taps0_kernel (const Taps0Params T)
{
    __shared__ uint32_t sm_inv[256];
    const TapsParams &P = T.t;

    pdl_launch_dependents ();
    if constexpr (OU)
    {
        for (uint32_t i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y)
            sm_inv[i] = __ldg (&P.inv_div_p8[i]) << 3;
        __syncthreads ();
    }

    const uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    const uint32_t strip = blockIdx.y * blockDim.y + threadIdx.y;
    const uint32_t yl0 = strip * P.rows_per_thread;
    if (x >= P.w_out || yl0 >= P.n_rows)
        return;
    const uint32_t yl1 = min (yl0 + P.rows_per_thread, P.n_rows);
    const uint32_t n_px = min ((uint32_t) PX, P.w_out - x);
    /* 24bpp stores cooperate across the warp's lanes (store_px4_rgb_anywhere): the lanes that got
     * here all walk the same strip of the same row range, so the live set never changes */
    const unsigned live_mask = __activemask ();
    const bool has_next = (threadIdx.x & 31) != 31 && x + PX < P.w_out;
    (void) live_mask; (void) has_next;

    /* this thread's four horizontal taps never change */
    uint32_t op[PX], oq[PX], Fx[PX];
#pragma unroll
    for (int o = 0; o < PX; o++)
    {
        const uint32_t e = __ldg (&P.tab_x[min (x + o, P.w_out - 1)]);
        op[o] = SMOL_TAB_OFS (e);
        oq[o] = min (op[o] + 1, P.w_in - 1);
        Fx[o] = SMOL_TAB_F (e);
    }

    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl0 * P.dst_pitch + (size_t) x * BO;
    const bool fast_store = FASTIO && n_px == PX;

    if (in_first_wave (T.prefetch))
    {
        /* the strip's first two source rows */
        const uint32_t r = SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl0]));
        const uint8_t *p = src + (size_t) r * P.src_pitch + (size_t) op[0] * BI;
        prefetch_l2 (p);
        prefetch_l2 (p + (size_t) (r + 1 < P.h_in ? P.src_pitch : 0));
    }
    pdl_wait ();

    auto hrow = [&] (uint32_t r, Px16 *out)
    {
        const uint8_t *row = src + (size_t) min (r, P.h_in - 1) * P.src_pitch;
        if (T.row_ahead && r + T.row_ahead < P.h_in)
            prefetch_l1 (row + (size_t) T.row_ahead * P.src_pitch + (size_t) op[0] * BI);   /* the strip walks down: hide the next rows' latency */
#pragma unroll
        for (int o = 0; o < PX; o++)
        {
            const Px16 p = taps0_fetch<BI, IU, AF, FASTIO> (row, op[o]);
            const Px16 q = taps0_fetch<BI, IU, AF, FASTIO> (row, oq[o]);
            const uint32_t F = Fx[o], G = 256u - F;
            out[o].a = __byte_perm (p.a * F + q.a * G, 0, 0x4341);     /* (acc >> 8) & 0x00ff00ff */
            out[o].b = __byte_perm (p.b * F + q.b * G, 0, 0x4341);
        }
    };

    /* one output row from the two cached source rows; F = 256 / F = 0 (copy, one, table tails)
     * come out exact from the same formula, so there are no special cases */
    auto emit = [&] (const Px16 *top, const Px16 *bot, uint32_t F)
    {
        const uint32_t G = 256u - F;
        uint32_t out[4];
#pragma unroll
        for (int o = 0; o < PX; o++)
        {
            const uint32_t acc_a = top[o].a * F + bot[o].a * G, acc_b = top[o].b * F + bot[o].b * G;
            if constexpr (OU)
            {
                uint32_t v = __byte_perm (acc_a, acc_b, 0x7351);      /* source byte order */
                v = half_unpremul<AF> (v, sm_inv);
                out[o] = __byte_perm (v, 0, P.prmt_sel);
            }
            else
                out[o] = __byte_perm (acc_a, acc_b, T.acc_prmt_sel);
        }
        if constexpr (PX == 1)
            *reinterpret_cast<uint32_t *> (dst) = out[0];       /* PX == 1 implies BO == 4 and FASTIO */
        else if constexpr (BO == 3)
            store_px4_rgb_anywhere (dst, out, n_px, live_mask, has_next);
        else if (fast_store)
            store_px4_aligned<BO> (dst, out);
        else
            store_px_slow (dst, out, n_px, BO);
        dst += P.dst_pitch;
    };

    /* Walk the strip with the two source rows ping-ponging between register sets A and B (the
     * reference's two-row cache, generic:1648-1682, without ever moving a row): in phase A the
     * top row is in A and the bottom one in B, in phase B the other way round. */
    const uint32_t *ty = P.tab_y + P.first_row;
    uint32_t yl = yl0;
    uint32_t e = __ldg (&ty[yl]);
    uint32_t r = SMOL_TAB_OFS (e);
    Px16 A[PX], B[PX];
    hrow (r, A);
    hrow (r + 1, B);

    for (;;)
    {
        /* phase A */
        do
        {
            emit (A, B, SMOL_TAB_F (e));
            if (++yl >= yl1)
                return;
            e = __ldg (&ty[yl]);
        }
        while (SMOL_TAB_OFS (e) == r);
        if (SMOL_TAB_OFS (e) != r + 1)
        {
            r = SMOL_TAB_OFS (e);
            hrow (r, A);
            hrow (r + 1, B);
            continue;
        }
        r++;
        hrow (r + 1, A);

        /* phase B */
        do
        {
            emit (B, A, SMOL_TAB_F (e));
            if (++yl >= yl1)
                return;
            e = __ldg (&ty[yl]);
        }
        while (SMOL_TAB_OFS (e) == r);
        if (SMOL_TAB_OFS (e) != r + 1)
        {
            r = SMOL_TAB_OFS (e);
            hrow (r, A);
            hrow (r + 1, B);
            continue;
        }
        r++;
        hrow (r + 1, B);
    }
}

/* Bilinear with halvings (reference BILINEAR_1H / _2H on either axis: every 2:1 .. 8:1
 * reduction that is not an exact power of two), 8-bit premultiplied intermediate, formats
 * resolved at compile time, 4-byte-aligned rows.  One thread per OUTPUT pixel: a downscale has
 * few output pixels and many taps each, so parallelism comes from the output grid and there is
 * next to no reuse between neighbouring outputs to exploit.  The thread sums 2^vh vertical
 * samples, each a tap between two horizontally filtered source rows (two-entry register cache:
 * consecutive samples usually share a row), each of those the sum of 2^hh horizontal taps. */
template <int BI, int BO, bool IU, bool OU, bool AF, int HH>
__global__ void __launch_bounds__ (256)
smol_tapsn_kernel (const Taps0Params T, uint32_t vh)
{
    __shared__ uint32_t sm_inv[256];
    const TapsParams &P = T.t;
    constexpr uint32_t N_H = 1u << HH;

    pdl_launch_dependents ();
    if constexpr (OU)
    {
        for (uint32_t i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y)
            sm_inv[i] = __ldg (&P.inv_div_p8[i]) << 3;
        __syncthreads ();
    }

    /* one thread = one output column x a strip of rows_per_thread output rows: the column's
     * horizontal taps are decoded once, and the two-row cache carries over from one output row
     * to the next (its last source row is often the next one's first) */
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t yl0 = (blockIdx.y * blockDim.y + threadIdx.y) * P.rows_per_thread;
    if (x >= P.w_out || yl0 >= P.n_rows)
        return;
    const uint32_t yl1 = min (yl0 + P.rows_per_thread, P.n_rows);

    const uint32_t n_v = 1u << vh;
    const uint32_t *ty = P.tab_y + ((P.first_row + yl0) << vh);
    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;

    uint32_t op[N_H], oq[N_H], Fx[N_H];
#pragma unroll
    for (uint32_t k = 0; k < N_H; k++)
    {
        const uint32_t e = __ldg (&P.tab_x[(x << HH) + k]);
        op[k] = SMOL_TAB_OFS (e) * BI;
        oq[k] = min (SMOL_TAB_OFS (e) + 1, P.w_in - 1) * BI;
        Fx[k] = SMOL_TAB_F (e);
    }

    if (in_first_wave (T.prefetch) && (threadIdx.x & 3) == 0)
    {
        /* every source row of the strip's first output pixel, from its first column on (four lanes
         * share the prefetch).  (Pulling the rows into L1 in every CTA was measured too: slower.) */
        const uint32_t ra = SMOL_TAB_OFS (__ldg (&ty[0])), rb = min (SMOL_TAB_OFS (__ldg (&ty[n_v - 1])) + 1, P.h_in - 1);
        for (uint32_t r = ra; r <= rb; r++)
            prefetch_l2 (src + (size_t) r * P.src_pitch + op[0]);
    }
    pdl_wait ();

    auto hval = [&] (uint32_t r) -> Px16
    {
        const uint8_t *row = src + (size_t) r * P.src_pitch;
        uint32_t acc_a = 0, acc_b = 0;
#pragma unroll
        for (uint32_t k = 0; k < N_H; k++)
        {
            const uint32_t F = Fx[k], G = 256u - F;
            const Px16 p = taps0_fetch<BI, IU, AF, true> (row + op[k], 0);
            const Px16 q = taps0_fetch<BI, IU, AF, true> (row + oq[k], 0);
            acc_a += __byte_perm (p.a * F + q.a * G, 0, 0x4341);       /* ((..) >> 8) & 0x00ff00ff */
            acc_b += __byte_perm (p.b * F + q.b * G, 0, 0x4341);
        }
        Px16 h;
        h.a = (acc_a >> HH) & 0x00ff00ffu;
        h.b = (acc_b >> HH) & 0x00ff00ffu;
        return h;
    };

    uint32_t idx0 = 0xffffffffu, idx1 = 0xffffffffu;
    Px16 c0, c1;
    c0.a = c0.b = c1.a = c1.b = 0;
    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl0 * P.dst_pitch + (size_t) x * BO;

#pragma unroll 1
    for (uint32_t yl = yl0; yl < yl1; yl++, ty += n_v, dst += P.dst_pitch)
    {
        uint32_t acc_a = 0, acc_b = 0;

#pragma unroll 1
        for (uint32_t kv = 0; kv < n_v; kv++)
        {
            const uint32_t e = __ldg (&ty[kv]);
            const uint32_t r0 = SMOL_TAB_OFS (e), F = SMOL_TAB_F (e), G = 256u - F;
            const uint32_t r1 = min (r0 + 1, P.h_in - 1);

            if (r0 != idx0)
            {
                if (r0 == idx1)
                {
                    const Px16 t = c0; c0 = c1; c1 = t;
                    idx1 = idx0;
                }
                else
                    c0 = hval (r0);
                idx0 = r0;
            }
            if (r1 != idx1)
            {
                c1 = hval (r1);
                idx1 = r1;
            }
            acc_a += __byte_perm (c0.a * F + c1.a * G, 0, 0x4341);
            acc_b += __byte_perm (c0.b * F + c1.b * G, 0, 0x4341);
        }

        const uint32_t fa = (acc_a >> vh) & 0x00ff00ffu, fb = (acc_b >> vh) & 0x00ff00ffu;
        uint32_t v = fa | (fb << 8);                                         /* source byte order */
        if constexpr (OU)
            v = half_unpremul<AF> (v, sm_inv);
        v = __byte_perm (v, 0, P.prmt_sel);

        if constexpr (BO == 4)
            *reinterpret_cast<uint32_t *> (dst) = v;
        else
        {
            dst[0] = (uint8_t) v; dst[1] = (uint8_t) (v >> 8); dst[2] = (uint8_t) (v >> 16);
        }
    }
}

#ifndef SMOL_TAPS11_MINBLOCKS
#define SMOL_TAPS11_MINBLOCKS 4
#endif
/* Bilinear with ONE halving on both axes (reference BILINEAR_1H x BILINEAR_1H: every reduction
 * between 2:1 and 4:1 that is not exactly 2:1 -- 4K -> 720p, 1080p -> 480p ...; the commonest
 * non-trivial downscale), 8-bit premultiplied intermediate.  smol_tapsn_kernel with everything it
 * decides at run time resolved: a thread owns one output column and a STRIP of output rows, reads
 * the column's two horizontal taps as a four-pixel window at constant offsets from one row pointer
 * and walks the strip with straight-line code -- an output pixel is the mean of two vertical
 * samples, each a tap between two horizontally filtered rows; which of those rows coincide (the
 * second sample usually starts on the first one's lower row, the next pixel often on this one's last
 * row) depends on the row only, so every branch is uniform across the CTA.  Less than half of the
 * general kernel's instructions per output pixel. */
template <int BI, int BO, bool IU, bool OU, bool AF>
__global__ void __launch_bounds__ (256, SMOL_TAPS11_MINBLOCKS)
smol_taps11_kernel (const Taps0Params T)
{
    __shared__ uint32_t sm_inv[OU ? 256 : 1];
    const TapsParams &P = T.t;

    pdl_launch_dependents ();
    if constexpr (OU)
    {
        for (uint32_t i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y)
            sm_inv[i] = __ldg (&P.inv_div_p8[i]) << 3;
        __syncthreads ();
    }

    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t yl0 = (blockIdx.y * blockDim.y + threadIdx.y) * P.rows_per_thread;
    if (x >= P.w_out || yl0 >= P.n_rows)
        return;
    const uint32_t yl1 = min (yl0 + P.rows_per_thread, P.n_rows);
    const uint32_t *ty = P.tab_y + ((P.first_row + yl0) << 1);
    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    const uint32_t h_last = P.h_in - 1, pitch = P.src_pitch;

    const uint32_t e0 = __ldg (&P.tab_x[2 * x]), e1 = __ldg (&P.tab_x[2 * x + 1]);
    const uint32_t F0 = SMOL_TAB_F (e0), G0 = 256u - F0, F1 = SMOL_TAB_F (e1), G1 = 256u - F1;
    /* The two taps read pixels p0, p0 + 1 and p1, p1 + 1 with p1 = p0 + 1 or p0 + 2 (the samples are
     * ratio / 2 = 1 .. 2 pixels apart): a window of four pixels at constant offsets from ONE row
     * pointer, the second tap's pair picked by two selects.  Columns whose window would cross the
     * row's end (the last one or two) address every pixel on its own. */
    const uint32_t p0 = SMOL_TAB_OFS (e0), p1 = SMOL_TAB_OFS (e1);
    const bool near = p1 == p0 + 1, window = (near || p1 == p0 + 2) && p0 + 3 < P.w_in;
    const uint8_t *col = src + (size_t) p0 * BI;
    const uint32_t oq0 = (min (p0 + 1, P.w_in - 1) - p0) * BI, op1 = (p1 - p0) * BI, oq1 = (min (p1 + 1, P.w_in - 1) - p0) * BI;

    if (in_first_wave (T.prefetch) && (threadIdx.x & 3) == 0)
    {
        /* the source rows of the strip's first output pixel (four lanes share a prefetch) */
        const uint32_t ra = SMOL_TAB_OFS (__ldg (&ty[0])), rb = min (SMOL_TAB_OFS (__ldg (&ty[1])) + 1, h_last);
#pragma unroll 1
        for (uint32_t r = ra; r <= rb; r++)
            prefetch_l2 (col + (size_t) r * pitch);
    }
    pdl_wait ();

    auto load_raw = [&] (const uint8_t *p) -> uint32_t
    {
        if constexpr (BI == 4)
            return __ldg (reinterpret_cast<const uint32_t *> (p));
        else
            return (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16) | 0xff000000u;
    };
    auto unpack = [&] (uint32_t raw) -> Px16
    {
        Px16 r;
        r.a = raw & 0x00ff00ffu;
        r.b = __byte_perm (raw, 0, 0x4341);
        if constexpr (IU)
        {
            /* premultiply: ((c + 1) * (alpha + 1) - 1) >> 8, alpha lane untouched (generic:238-244) */
            if constexpr (AF)
            {
                const uint32_t alpha = raw & 0xff, m = alpha + 1;
                r.a = (((((r.a & 0x00ff0000u) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff0000u) | alpha;
                r.b = (((r.b + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
            }
            else
            {
                const uint32_t alpha = raw >> 24, m = alpha + 1;
                r.a = (((r.a + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
                r.b = (((((r.b & 0x000000ffu) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x000000ffu) | (alpha << 16);
            }
        }
        return r;
    };
    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl0 * P.dst_pitch + (size_t) x * BO;

    /* WINDOW resolved at compile time: the body exists twice, the few boundary columns take the second copy */
    auto strip = [&] (auto window_tag)
    {
    constexpr bool WINDOW = decltype (window_tag)::value;
    /* A source row at this column: its four raw pixels (load_row), then the mean of the two taps
     * (hfilter).  Split so that ALL rows of an output pixel are requested before the first one is
     * used: one memory round trip per output pixel instead of one per row -- the kernel is bound by
     * load latency, not by issue slots. */
    struct Raw4 { uint32_t w[4]; };
    auto load_row = [&] (uint32_t r) -> Raw4
    {
        const uint8_t *row = col + (size_t) r * pitch;
        Raw4 v;
        v.w[0] = load_raw (row);
        if constexpr (WINDOW)
        {
            v.w[1] = load_raw (row + BI);
            v.w[2] = load_raw (row + 2 * BI);
            v.w[3] = load_raw (row + 3 * BI);       /* (only used when the taps are two pixels apart; a neighbouring lane reads it anyway) */
        }
        else
        {
            v.w[1] = load_raw (row + oq0);
            v.w[2] = load_raw (row + op1);
            v.w[3] = load_raw (row + oq1);
        }
        return v;
    };
    auto hfilter = [&] (const Raw4 &v) -> Px16
    {
        const uint32_t w2 = WINDOW && near ? v.w[1] : v.w[2], w3 = WINDOW && near ? v.w[2] : v.w[3];
        const Px16 a0 = unpack (v.w[0]), b0 = unpack (v.w[1]), a1 = unpack (w2), b1 = unpack (w3);
        Px16 h;
        h.a = ((__byte_perm (a0.a * F0 + b0.a * G0, 0, 0x4341) + __byte_perm (a1.a * F1 + b1.a * G1, 0, 0x4341)) >> 1) & 0x00ff00ffu;
        h.b = ((__byte_perm (a0.b * F0 + b0.b * G0, 0, 0x4341) + __byte_perm (a1.b * F1 + b1.b * G1, 0, 0x4341)) >> 1) & 0x00ff00ffu;
        return h;
    };

    Px16 last;                      /* the most recently filtered source row: often the next pixel's first */
    uint32_t last_idx = 0xffffffffu;
    last.a = last.b = 0;

    uint32_t ea = __ldg (&ty[0]), eb = __ldg (&ty[1]);
#pragma unroll 1
    for (uint32_t yl = yl0; yl < yl1; yl++, ty += 2, dst += P.dst_pitch)
    {
        const uint32_t a0 = SMOL_TAB_OFS (ea), a1 = min (a0 + 1, h_last), Fa = SMOL_TAB_F (ea), Ga = 256u - Fa;
        const uint32_t b0 = SMOL_TAB_OFS (eb), b1 = min (b0 + 1, h_last), Fb = SMOL_TAB_F (eb), Gb = 256u - Fb;

        if (yl + 1 < yl1)
        {
            /* the next pixel's rows on their way while this one is computed */
            ea = __ldg (&ty[2]);
            eb = __ldg (&ty[3]);
            if (T.row_ahead && (threadIdx.x & 3) == 0)
            {
                /* at most three rows are new: b1 + 1 .. the next second sample's lower row */
                const uint32_t nb = min (SMOL_TAB_OFS (eb) + 1, h_last), n0 = max (SMOL_TAB_OFS (ea), b1 + 1);
                const uint8_t *pr = col + (size_t) n0 * pitch;
                if (n0 <= nb)
                    prefetch_l2 (pr);
                if (n0 + 1 <= nb)
                    prefetch_l2 (pr + pitch);
                if (n0 + 2 <= nb)
                    prefetch_l2 (pr + 2 * (size_t) pitch);
            }
        }

        /* rows a0, a1 for the first sample, b0, b1 for the second; usually b0 == a1, and a0 is often
         * the previous pixel's b1.  All of this depends on the row only: uniform branches. */
        const bool have_a0 = a0 == last_idx, chained = b0 == a1;
        Raw4 ra, rb, rc, rd;
        if (!have_a0)
            ra = load_row (a0);
        rb = load_row (a1);
        rc = load_row (chained ? b1 : b0);
        if (!chained)
            rd = load_row (b1);

        Px16 X, Y, Z, W;
        if (have_a0) X = last; else X = hfilter (ra);
        Y = hfilter (rb);
        if (chained)
        {
            Z = Y;
            W = hfilter (rc);
        }
        else
        {
            Z = hfilter (rc);
            W = hfilter (rd);
        }
        last = W;
        last_idx = b1;

        const uint32_t acc_a = __byte_perm (X.a * Fa + Y.a * Ga, 0, 0x4341) + __byte_perm (Z.a * Fb + W.a * Gb, 0, 0x4341);
        const uint32_t acc_b = __byte_perm (X.b * Fa + Y.b * Ga, 0, 0x4341) + __byte_perm (Z.b * Fb + W.b * Gb, 0, 0x4341);
        uint32_t v = __byte_perm (acc_a >> 1, acc_b >> 1, 0x6240);             /* source byte order */
        if constexpr (OU)
            v = half_unpremul<AF> (v, sm_inv);
        v = __byte_perm (v, 0, P.prmt_sel);

        if constexpr (BO == 4)
            *reinterpret_cast<uint32_t *> (dst) = v;
        else
        {
            dst[0] = (uint8_t) v; dst[1] = (uint8_t) (v >> 8); dst[2] = (uint8_t) (v >> 16);
        }
    }
    };
    if (window)
        strip (std::true_type {});
    else
        strip (std::false_type {});
}

/* ------------------------------------------------------------------------------------------ *
 * "mag" kernel: vertical magnification (h_out > h_in; BASELINE config 4), bilinear / copy / one *
 * horizontally, 8-bit premultiplied intermediate.                                              *
 *                                                                                              *
 * On an upscale every source row feeds several output rows and every source pixel several      *
 * output columns, so the work is split in two phases per CTA tile (TW output columns x TH       *
 * output rows) with the intermediate kept in shared memory, never in HBM:                      *
 *   1. the few source rows the tile needs are staged (coalesced), unpacked once per source     *
 *      pixel into 16-bit lanes, and filtered horizontally once per (source row, output column) *
 *      into sm_h;                                                                              *
 *   2. every thread then produces groups of four adjacent output pixels of one output row:      *
 *      two 128-bit shared-memory reads per source row, four multiply-adds per pixel, one PRMT   *
 *      that does shift + mask + byte interleave + channel reorder at once, vector store.        *
 * ------------------------------------------------------------------------------------------ */

struct MagParams
{
    TapsParams t;
    uint32_t tile_w, tile_h;        /* output tile; tile_w is a multiple of 4, at most 256 */
    uint32_t u_pitch;               /* pixels per row of the unpacked source window (even) */
    uint32_t max_src_rows;          /* bound on source rows per tile */
    uint32_t u_cw;                  /* power of two >= min (u_pitch, 256): thread columns in phase 1a */
    uint32_t u_cw_log2, tile_w_log2;/* tile_w is a power of two (>= 4) */
    uint32_t acc_prmt_sel;          /* (acc_a, acc_b) high bytes -> destination byte order */
    uint32_t src_u32_ok;            /* 32bpp source rows are 4-byte aligned */
};

/* Format traits resolved at compile time: BI / BO bytes per pixel in / out, IU / OU unassociated
 * alpha in / out, AF alpha is the first byte of a 32bpp source pixel (else the last). */
template <int BI, int BO, bool IU, bool OU, bool AF, bool FASTIO>
__global__ void __launch_bounds__ (256)
smol_mag_kernel (const MagParams M)
{
    extern __shared__ __align__ (16) uint8_t sm_dyn[];
    __shared__ uint32_t sm_inv[256];
    __shared__ uint32_t sm_ty[64];
    const TapsParams &P = M.t;
    const uint32_t tid = threadIdx.x;

    pdl_launch_dependents ();
    if constexpr (OU)
        sm_inv[tid] = __ldg (&P.inv_div_p8[tid]) << 3;

    const uint32_t x0 = blockIdx.x * M.tile_w;
    const uint32_t x1 = min (x0 + M.tile_w, P.w_out);              /* exclusive */
    const uint32_t yl0 = blockIdx.y * M.tile_h;
    const uint32_t yl1 = min (yl0 + M.tile_h, P.n_rows);
    const uint32_t tw = x1 - x0, th = yl1 - yl0;

    /* source window of the tile (tables are library-owned: readable before the dependency wait) */
    const uint32_t c_lo = SMOL_TAB_OFS (__ldg (&P.tab_x[x0]));
    const uint32_t c_hi = min (SMOL_TAB_OFS (__ldg (&P.tab_x[x1 - 1])) + 1, P.w_in - 1);
    const uint32_t r_lo = SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl0]));
    const uint32_t r_hi = min (SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl1 - 1])) + 1, P.h_in - 1);
    const uint32_t n_cols = c_hi - c_lo + 1, n_rows = r_hi - r_lo + 1;
    if (tid < th)
        sm_ty[tid] = __ldg (&P.tab_y[P.first_row + yl0 + tid]);

    uint2 *sm_u = reinterpret_cast<uint2 *> (sm_dyn);                           /* [rows][u_pitch] unpacked source */
    uint32_t *sm_ha = reinterpret_cast<uint32_t *> (sm_u + (size_t) M.max_src_rows * M.u_pitch);
    uint32_t *sm_hb = sm_ha + (size_t) M.max_src_rows * M.tile_w;               /* two planes [rows][tile_w] */

    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    uint8_t *dst_img = P.dst + (size_t) blockIdx.z * P.dst_image_stride;

    pdl_wait ();

    /* phase 1a: load + unpack the source window, one source pixel per thread step */
    {
        const uint32_t ct = tid & (M.u_cw - 1), rt = tid >> M.u_cw_log2, r_step = 256u >> M.u_cw_log2;
        for (uint32_t r = rt; r < n_rows; r += r_step)
        {
            const uint8_t *row = src + (size_t) (r_lo + r) * P.src_pitch + (size_t) c_lo * BI;
            for (uint32_t c = ct; c < n_cols; c += M.u_cw)
            {
                const uint8_t *p = row + c * BI;
                uint32_t raw;
                if constexpr (BI == 4 && FASTIO)
                    raw = __ldg (reinterpret_cast<const uint32_t *> (p));
                else
                {
                    raw = (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16);
                    raw |= BI == 4 ? ((uint32_t) __ldg (p + 3) << 24) : 0xff000000u;
                }
                uint32_t a = raw & 0x00ff00ffu, b2 = (raw >> 8) & 0x00ff00ffu;
                if constexpr (IU)
                {
                    /* premultiply: ((c + 1) * (alpha + 1) - 1) >> 8, alpha lane untouched (generic:238-244) */
                    if constexpr (AF)
                    {
                        const uint32_t alpha = raw & 0xff, m = alpha + 1;
                        a = (((((a & 0x00ff0000u) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff0000u) | alpha;
                        b2 = (((b2 + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
                    }
                    else
                    {
                        const uint32_t alpha = raw >> 24, m = alpha + 1;
                        a = (((a + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
                        b2 = (((((b2 & 0x000000ffu) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x000000ffu) | (alpha << 16);
                    }
                }
                sm_u[r * M.u_pitch + c] = make_uint2 (a, b2);
            }
        }
    }
    __syncthreads ();

    /* phase 1b: horizontal taps, once per (source row, output column); thread -> column fixed */
    {
        /* full tiles: tile_w is a power of two, no divisions */
        const bool full = tw == M.tile_w;
        const uint32_t xl = full ? (tid & (M.tile_w - 1)) : tid % tw;
        const uint32_t r_step = full ? (256u >> M.tile_w_log2) : 256 / tw;
        const uint32_t r_first = full ? (tid >> M.tile_w_log2) : tid / tw;
        const uint32_t e = __ldg (&P.tab_x[x0 + xl]);
        const uint32_t op = SMOL_TAB_OFS (e) - c_lo, oq = min (SMOL_TAB_OFS (e) + 1, P.w_in - 1) - c_lo;
        const uint32_t F = SMOL_TAB_F (e), G = 256u - F;
        if (r_step > 0)
        {
            uint32_t r = r_first;
            const uint2 *pu = sm_u + r * M.u_pitch + op, *qu = sm_u + r * M.u_pitch + oq;
            uint32_t *wa = sm_ha + r * M.tile_w + xl, *wb = sm_hb + r * M.tile_w + xl;
            const uint32_t u_inc = r_step * M.u_pitch, h_inc = r_step * M.tile_w;
            for (; r < n_rows; r += r_step, pu += u_inc, qu += u_inc, wa += h_inc, wb += h_inc)
            {
                const uint2 p = *pu, q = *qu;
                *wa = __byte_perm (p.x * F + q.x * G, 0, 0x4341);   /* (acc >> 8) & 0x00ff00ff */
                *wb = __byte_perm (p.y * F + q.y * G, 0, 0x4341);
            }
        }
    }
    __syncthreads ();

    /* phase 2: vertical taps + pack + store.  Lane -> 4 adjacent output pixels (conflict-free
     * 128-bit shared-memory reads); each thread walks a run of consecutive output rows so the two
     * source rows stay in registers while only the weight changes (4:1 upscale: ~4 rows per load). */
    const bool full = tw == M.tile_w;
    const uint32_t groups = (tw + 3) >> 2;                          /* <= 64 */
    const uint32_t n_runs = full ? (1024u >> M.tile_w_log2) : 256 / groups;
    const uint32_t g = full ? (tid & (groups - 1)) : tid % groups;
    const uint32_t run = full ? (tid >> (M.tile_w_log2 - 2)) : tid / groups;
    if (run >= n_runs)
        return;
    const uint32_t rows_per_run = full ? ((th + n_runs - 1) >> (10 - M.tile_w_log2)) : (th + n_runs - 1) / n_runs;
    const uint32_t ry_begin = run * rows_per_run, ry_end = min (ry_begin + rows_per_run, th);
    const uint32_t x = x0 + 4 * g;
    const uint32_t n_px = min (4u, x1 - x);
    const uint32_t *ha = sm_ha + 4 * g, *hb = sm_hb + 4 * g;
    uint8_t *dst = dst_img + (size_t) (yl0 + ry_begin) * P.dst_pitch + (size_t) x * BO;
    const bool fast_store = FASTIO && n_px == 4;
    uint32_t ry = ry_begin;
    uint32_t e = ry < ry_end ? sm_ty[ry] : 0;
    while (ry < ry_end)
    {
        /* one segment = consecutive output rows that read the same source row pair */
        const uint32_t ofs = SMOL_TAB_OFS (e);
        const uint32_t r0 = ofs - r_lo, r1 = min (ofs + 1, P.h_in - 1) - r_lo;
        const uint4 ta = *reinterpret_cast<const uint4 *> (ha + r0 * M.tile_w);
        const uint4 tb = *reinterpret_cast<const uint4 *> (hb + r0 * M.tile_w);
        const uint4 ba = *reinterpret_cast<const uint4 *> (ha + r1 * M.tile_w);
        const uint4 bb = *reinterpret_cast<const uint4 *> (hb + r1 * M.tile_w);

        do
        {
            const uint32_t F = SMOL_TAB_F (e), G = 256u - F;
            const uint32_t acc_a[4] = { ta.x * F + ba.x * G, ta.y * F + ba.y * G, ta.z * F + ba.z * G, ta.w * F + ba.w * G };
            const uint32_t acc_b[4] = { tb.x * F + bb.x * G, tb.y * F + bb.y * G, tb.z * F + bb.z * G, tb.w * F + bb.w * G };
            uint32_t out[4];

#pragma unroll
            for (int o = 0; o < 4; o++)
            {
                if constexpr (OU)
                {
                    uint32_t v = __byte_perm (acc_a[o], acc_b[o], 0x7351);      /* source byte order */
                    v = half_unpremul<AF> (v, sm_inv);
                    out[o] = __byte_perm (v, 0, P.prmt_sel);
                }
                else
                    out[o] = __byte_perm (acc_a[o], acc_b[o], M.acc_prmt_sel);
            }

            if (fast_store)
                store_px4_aligned<BO> (dst, out);
            else
                store_px_slow (dst, out, n_px, BO);

            ry++;
            dst += P.dst_pitch;
            e = sm_ty[min (ry, th - 1)];
        }
        while (ry < ry_end && SMOL_TAB_OFS (e) == ofs);
    }
}

/* ------------------------------------------------------------------------------------------ *
 * "magb" kernel: the magnification tile again, but BYTE-granular in its vertical stage          *
 * (BASELINE config 4; any 24/32bpp pair whose output needs no per-pixel alpha operation).       *
 *                                                                                              *
 * Once the horizontal pass has put every channel in destination byte order, the vertical tap    *
 * is the same operation on every BYTE of the output row -- pixel boundaries stop mattering.     *
 * So the tile is tile_b output BYTES wide (a multiple of 16, not of the pixel size) and:        *
 *   1. the source window is loaded (24bpp: four pixels = three aligned words per step) and      *
 *      unpacked once per source pixel into 16-bit lanes;                                        *
 *   2. the horizontal taps run once per (source row, group of 4 output pixels); results are     *
 *      reordered to destination bytes by one PRMT per pixel and stored PACKED (12 or 16 bytes   *
 *      per group) as the tile's horizontally filtered rows;                                     *
 *   3. a thread owns one 16-byte column of the tile and walks output rows: per run of rows that *
 *      share a source row pair, two 128-bit shared-memory reads expanded to 16-bit lanes; per   *
 *      output row 16 multiply-adds (two bytes each), 4 PRMTs, ONE 128-bit store -- 1 multiply   *
 *      per output byte and a fully coalesced row segment per warp whatever the pixel size       *
 *      (the per-pixel kernel above spends 16 multiplies and 3 stores per 12 bytes at 24bpp).    *
 * ------------------------------------------------------------------------------------------ */

/* Resident CTAs per SM asked of the compiler.  8 (32 registers per thread) would let BASELINE
 * config 4's 1152 tiles run as one wave instead of 1.3, but the spills it costs are worse
 * (measured: 11.7 -> 12.3 us; 2x 24bpp 15.2 -> 19.1 us), so the kernel keeps its 40 registers. */
#ifndef SMOL_MAGB_MINBLOCKS
#define SMOL_MAGB_MINBLOCKS 1
#endif
#define SMOL_MAGB_MAX_TILE_H 256        /* rows of a tile (one thread per row builds the per-row tables) */

struct MagbParams
{
    TapsParams t;
    uint32_t nb_row;                /* bytes per output row: w_out * bpp_out */
    uint32_t tile_b, tile_h;        /* output tile: tile_b bytes (16 << chunks_log2), tile_h rows (<= SMOL_MAGB_MAX_TILE_H) */
    uint32_t chunks_log2;           /* log2 (tile_b / 16), at most 8 */
    uint32_t gcols_log2;            /* thread columns of stage 2: power of two >= pixel groups per tile, at most 8 */
    uint32_t u_pitch;               /* pixels per row of the unpacked source window */
    uint32_t u_cw, u_cw_log2;       /* thread columns of stage 1 (power of two, at most 256) */
    uint32_t h_pitch;               /* bytes per horizontally filtered row: tile_b + 32 */
    uint32_t max_src_rows;
    uint32_t acc_prmt_sel;          /* (acc_a, acc_b) high bytes -> destination byte order */
    uint32_t prefetch;              /* CTAs (in launch order) that L2-prefetch their source window ahead of the dependency wait */
};

/* BI / BO bytes per pixel in / out; IU unassociated input (premultiplied on unpack); OU unassociated
 * output (32bpp only: a 16-byte column is then four whole pixels, unpremultiplied just before the
 * store); AF alpha is byte 0 of the source pixel; SRC32 source rows are 4-byte aligned. */
template <int BI, int BO, bool IU, bool OU, bool AF, bool SRC32>
__global__ void __launch_bounds__ (256, SMOL_MAGB_MINBLOCKS)
smol_magb_kernel (const MagbParams M)
{
    static_assert (!OU || BO == 4, "unassociated output is 32bpp");
    extern __shared__ __align__ (16) uint8_t sm_dyn[];
    __shared__ uint32_t sm_inv[OU ? 256 : 1];
    __shared__ uint32_t sm_ty[SMOL_MAGB_MAX_TILE_H];
    __shared__ uint4 sm_row[SMOL_MAGB_MAX_TILE_H];  /* per output row of the tile: { F, 256 - F, source row offset, end of its run } */
    const TapsParams &P = M.t;
    const uint32_t tid = threadIdx.x;
    constexpr bool GROUPS = BI == 3 && SRC32;       /* stage 1 works on groups of four 24bpp pixels */

    pdl_launch_dependents ();
    if constexpr (OU)
        sm_inv[tid] = __ldg (&P.inv_div_p8[tid]) << 3;      /* visible after the stage barriers */

    const uint32_t b0 = blockIdx.x * M.tile_b;
    const uint32_t b1 = min (b0 + M.tile_b, M.nb_row);             /* exclusive */
    const uint32_t yl0 = blockIdx.y * M.tile_h;
    const uint32_t yl1 = min (yl0 + M.tile_h, P.n_rows);
    const uint32_t th = yl1 - yl0;
    /* groups of four output pixels that overlap the tile's bytes */
    const uint32_t g_lo = b0 / (4 * BO), g_hi = (b1 - 1) / (4 * BO);
    const uint32_t n_groups = g_hi - g_lo + 1;
    const uint32_t x_hi = min (4 * g_hi + 3, P.w_out - 1);

    /* source window (tables are library-owned: readable before the dependency wait) */
    uint32_t c_lo = SMOL_TAB_OFS (__ldg (&P.tab_x[4 * g_lo]));
    if constexpr (GROUPS)
        c_lo &= ~3u;
    const uint32_t c_hi = min (SMOL_TAB_OFS (__ldg (&P.tab_x[x_hi])) + 1, P.w_in - 1);
    const uint32_t r_lo = SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl0]));
    const uint32_t r_hi = min (SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl1 - 1])) + 1, P.h_in - 1);
    const uint32_t n_cols = c_hi - c_lo + 1, n_rows = r_hi - r_lo + 1;
    if (tid < th)
        sm_ty[tid] = __ldg (&P.tab_y[P.first_row + yl0 + tid]);

    uint2 *sm_u = reinterpret_cast<uint2 *> (sm_dyn);                               /* [rows][u_pitch] unpacked source */
    uint8_t *sm_h = sm_dyn + (size_t) M.max_src_rows * M.u_pitch * 8;               /* [rows][h_pitch] filtered rows, destination bytes */

    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    uint8_t *dst_img = P.dst + (size_t) blockIdx.z * P.dst_image_stride;

    if (in_first_wave (M.prefetch))
    {
        /* the tile's source window, one 128-byte line per thread (see prefetch_l2) */
        const uint32_t w_bytes = n_cols * BI, lines = (w_bytes + 127) / 128 + 1;
        if (tid < lines * n_rows)
        {
            const uint32_t r = tid / lines, l = tid - r * lines;
            const uint8_t *p = src + (size_t) (r_lo + r) * P.src_pitch + (size_t) c_lo * BI;
            prefetch_l2 (l + 1 < lines ? p + 128 * l : p + w_bytes - 1);
        }
    }

    /* stage 2's taps (this thread's column of pixel groups): fetched now, so the table latency
     * hides behind stage 1.  Kept as byte offsets into a row of sm_u. */
    const uint32_t gi = tid & ((1u << M.gcols_log2) - 1);
    uint32_t po[4], qo[4], F2[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const uint32_t e = __ldg (&P.tab_x[min (4 * (g_lo + gi) + i, P.w_out - 1)]);
        po[i] = SMOL_TAB_OFS (e);
        F2[i] = SMOL_TAB_F (e);
    }

    pdl_wait ();

    /* stage 1: load + unpack the source window */
    {
        auto unpack_store = [&] (uint32_t raw, uint2 *to)
        {
            uint32_t a = raw & 0x00ff00ffu, b2 = (raw >> 8) & 0x00ff00ffu;
            if constexpr (IU)
            {
                /* premultiply: ((c + 1) * (alpha + 1) - 1) >> 8, alpha lane untouched (generic:238-244) */
                if constexpr (AF)
                {
                    const uint32_t alpha = raw & 0xff, m = alpha + 1;
                    a = (((((a & 0x00ff0000u) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff0000u) | alpha;
                    b2 = (((b2 + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
                }
                else
                {
                    const uint32_t alpha = raw >> 24, m = alpha + 1;
                    a = (((a + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
                    b2 = (((((b2 & 0x000000ffu) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x000000ffu) | (alpha << 16);
                }
            }
            *to = make_uint2 (a, b2);
        };
        auto load_bytes = [&] (const uint8_t *p) -> uint32_t
        {
            uint32_t raw = (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16);
            raw |= BI == 4 ? ((uint32_t) __ldg (p + 3) << 24) : 0xff000000u;
            return raw;
        };
        const uint32_t n_items = GROUPS ? (n_cols + 3) >> 2 : n_cols;
        const uint32_t ct = tid & (M.u_cw - 1), rt = tid >> M.u_cw_log2, r_step = 256u >> M.u_cw_log2;
        for (uint32_t r = rt; r < n_rows; r += r_step)
        {
            const uint8_t *row = src + (size_t) (r_lo + r) * P.src_pitch + (size_t) c_lo * BI;
            uint2 *urow = sm_u + r * M.u_pitch;
            for (uint32_t c = ct; c < n_items; c += M.u_cw)
            {
                if constexpr (GROUPS)
                {
                    if (c_lo + 4 * c + 3 < P.w_in)
                    {
                        /* four pixels = three aligned words (c_lo is a multiple of 4) */
                        const uint32_t *w = reinterpret_cast<const uint32_t *> (row + 12 * c);
                        const uint32_t w0 = __ldg (w), w1 = __ldg (w + 1), w2 = __ldg (w + 2);
                        unpack_store (w0 | 0xff000000u, urow + 4 * c);
                        unpack_store (__funnelshift_r (w0, w1, 24) | 0xff000000u, urow + 4 * c + 1);
                        unpack_store (__funnelshift_r (w1, w2, 16) | 0xff000000u, urow + 4 * c + 2);
                        unpack_store ((w2 >> 8) | 0xff000000u, urow + 4 * c + 3);
                    }
                    else
                    {
                        for (uint32_t k = 0; k < 4 && c_lo + 4 * c + k < P.w_in; k++)
                            unpack_store (load_bytes (row + 12 * c + 3 * k), urow + 4 * c + k);
                    }
                }
                else if constexpr (BI == 4 && SRC32)
                    unpack_store (__ldg (reinterpret_cast<const uint32_t *> (row + 4 * c)), urow + c);
                else
                    unpack_store (load_bytes (row + c * BI), urow + c);
            }
        }
    }
    __syncthreads ();

    /* Rows of the tile grouped into runs that share a source row pair: per row its two weights
     * (one 64-bit broadcast read per row in stage 3), its source row and where its run ends. */
    if (tid < th)
    {
        const uint32_t e = sm_ty[tid], ofs = SMOL_TAB_OFS (e);
        uint32_t j = tid + 1;
        while (j < th && SMOL_TAB_OFS (sm_ty[j]) == ofs)
            j++;
        sm_row[tid] = make_uint4 (SMOL_TAB_F (e), 256u - SMOL_TAB_F (e), ofs, j);
    }

    /* stage 2: horizontal taps once per (source row, group of four output pixels); the thread's
     * column of groups is fixed, so offsets and weights stay in registers across rows and the
     * eight tap addresses just advance by a row */
    {
        const uint32_t rl = tid >> M.gcols_log2, r_step = 256u >> M.gcols_log2;
        if (gi < n_groups)
        {
            const uint32_t u_row = M.u_pitch * 8, u_base = (uint32_t) __cvta_generic_to_shared (sm_u) + rl * u_row;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                qo[i] = u_base + (min (po[i] + 1, P.w_in - 1) - c_lo) * 8;
                po[i] = u_base + (po[i] - c_lo) * 8;
            }
            /* output byte b of the row lives at sm_h[row][16 + b - b0]: the tile's first byte is
             * 16-byte aligned, groups that start before it (24bpp) fit in the 16 bytes of slack */
            uint8_t *h = sm_h + rl * M.h_pitch + (16 + (g_lo + gi) * 4 * BO - b0);
            const uint32_t u_inc = r_step * u_row, h_inc = r_step * M.h_pitch;
            for (uint32_t r = rl; r < n_rows; r += r_step, h += h_inc)
            {
                uint32_t D[4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    const uint2 p = lds_u64_ordered (po[i]), q = lds_u64_ordered (qo[i]);
                    const uint32_t G = 256u - F2[i];
                    /* destination byte order, or source order when the pixels still have to be unpremultiplied */
                    D[i] = __byte_perm (p.x * F2[i] + q.x * G, p.y * F2[i] + q.y * G, OU ? 0x7351u : M.acc_prmt_sel);
                    po[i] += u_inc;
                    qo[i] += u_inc;
                }
                if constexpr (BO == 4)
                    *reinterpret_cast<uint4 *> (h) = make_uint4 (D[0], D[1], D[2], D[3]);
                else
                {
                    uint32_t *h32 = reinterpret_cast<uint32_t *> (h);
                    h32[0] = __byte_perm (D[0], D[1], 0x4210);
                    h32[1] = __byte_perm (D[1], D[2], 0x5421);
                    h32[2] = __byte_perm (D[2], D[3], 0x6542);
                }
            }
        }
    }
    __syncthreads ();

    /* stage 3: vertical taps on bytes.  Thread -> one 16-byte column, a contiguous share of the rows. */
    const uint32_t c = tid & ((1u << M.chunks_log2) - 1), grp = tid >> M.chunks_log2;
    const uint32_t bb = b0 + 16 * c;
    if (bb >= b1)
        return;
    const uint32_t n_valid = min (16u, b1 - bb);
    const uint32_t rows_per = (th + (256u >> M.chunks_log2) - 1) >> (8 - M.chunks_log2);
    const uint32_t ry_begin = grp * rows_per, ry_end = min (ry_begin + rows_per, th);
    uint8_t *dst = dst_img + (size_t) (yl0 + ry_begin) * P.dst_pitch + bb;
    const uint8_t *hcol = sm_h + 16 + 16 * c;
    /* The two source rows of a run, expanded to 16-bit lanes.  On a magnification the next run's
     * upper row is this run's lower row, so the two register sets swap roles from run to run
     * (X on top, then Y on top) and only one row is fetched and expanded per run.
     * FULL: the column's 16 bytes all lie inside the row (everything but a ragged last column). */
    auto stage3 = [&] (auto full_tag, auto a16_tag)
    {
        constexpr bool FULL = decltype (full_tag)::value;
        constexpr bool A16 = decltype (a16_tag)::value;     /* destination rows on 16-byte boundaries (else: 4-byte) */
        uint32_t X[8], Y[8], hx = 0xffffffffu, hy = 0xffffffffu;
        auto fetch_row = [&] (uint32_t (&R)[8], uint32_t r)
        {
            const uint4 v = *reinterpret_cast<const uint4 *> (hcol + r * M.h_pitch);
            const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                R[2 * k] = __byte_perm (w[k], 0, 0x4140);
                R[2 * k + 1] = __byte_perm (w[k], 0, 0x4342);
            }
        };
        uint32_t ry = ry_begin;
        uint8_t *out = dst;
        auto run_rows = [&] (const uint32_t (&T)[8], const uint32_t (&B)[8], uint32_t run_end)
        {
            do
            {
                const uint2 fg = *reinterpret_cast<const uint2 *> (&sm_row[ry]);
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    o[j] = __byte_perm (T[2 * j] * fg.x + B[2 * j] * fg.y, T[2 * j + 1] * fg.x + B[2 * j + 1] * fg.y, 0x7531);
                if constexpr (OU)
                {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        o[j] = __byte_perm (half_unpremul<AF> (o[j], sm_inv), 0, P.prmt_sel);
                }
                if constexpr (FULL && A16)
                    *reinterpret_cast<uint4 *> (out) = make_uint4 (o[0], o[1], o[2], o[3]);
                else if constexpr (FULL)
                {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        reinterpret_cast<uint32_t *> (out)[k] = o[k];
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        if ((uint32_t) k < n_valid)
                            out[k] = (uint8_t) (o[k >> 2] >> ((k & 3) * 8));
                }
                out += P.dst_pitch;
            }
            while (++ry < run_end);
        };
        while (ry < ry_end)
        {
            {
                const uint2 oe = *reinterpret_cast<const uint2 *> (&sm_row[ry].z);
                const uint32_t r0 = oe.x - r_lo, r1 = min (oe.x + 1, P.h_in - 1) - r_lo;
                if (hx != r0) { fetch_row (X, r0); hx = r0; }
                if (hy != r1) { fetch_row (Y, r1); hy = r1; }
                run_rows (X, Y, min (oe.y, ry_end));
            }
            if (ry >= ry_end)
                break;
            {
                const uint2 oe = *reinterpret_cast<const uint2 *> (&sm_row[ry].z);
                const uint32_t r0 = oe.x - r_lo, r1 = min (oe.x + 1, P.h_in - 1) - r_lo;
                if (hy != r0) { fetch_row (Y, r0); hy = r0; }
                if (hx != r1) { fetch_row (X, r1); hx = r1; }
                run_rows (Y, X, min (oe.y, ry_end));
            }
        }
    };
    if (n_valid != 16)
        stage3 (std::false_type {}, std::false_type {});
    else if (((reinterpret_cast<uintptr_t> (dst_img) | P.dst_pitch) & 15) == 0)
        stage3 (std::true_type {}, std::true_type {});
    else
        stage3 (std::true_type {}, std::false_type {});
}

/* ------------------------------------------------------------------------------------------ *
 * "box" kernel: box filter on both axes (large downscales; BASELINE config 3), 32bpp source.     *
 *                                                                                              *
 * Almost all the work of a big downscale is per SOURCE pixel (unpack, and for linear light the  *
 * unpremultiply -> sRGB table -> premultiply chain), so the kernel is organised around feeding  *
 * source pixels to threads with as little overhead as possible:                                *
 *   - the unit of work is one WARP producing 32 / G adjacent output pixels of one output row    *
 *     (G lanes share a column when spans are long); warps stride over the work items, so load   *
 *     balance is at warp granularity and no block-wide barrier is ever needed;                  *
 *   - for each source row of the item the warp copies the row segment it needs into its own     *
 *     shared-memory buffer with cp.async (16-byte, fully coalesced; byte-exact at a ragged row  *
 *     end), double-buffered so the copy of row r + 1 overlaps the arithmetic on row r;          *
 *   - each lane then walks its own span in shared memory (stride ~ratio words between lanes:    *
 *     practically conflict-free), unpacks and accumulates in registers, applies the two edge    *
 *     weights, normalises with span_mul_x (the reference quantises after the horizontal pass,   *
 *     generic:1263-1270) and adds the row into its vertical accumulator with the row's weight;  *
 *   - after the last row: normalise with span_mul_y, repack, store.                            *
 * The data tables live in shared memory (data-dependent gathers).                              *
 * ------------------------------------------------------------------------------------------ */

enum { BM_P8_P = 0, BM_P8_U = 1, BM_P8L_P = 2, BM_P8L_U = 3, BM_P16_U = 4, BM_P16L_U = 5 };

struct BoxParams
{
    SmolJobDesc d;
    const uint8_t *src; uint8_t *dst;
    uint32_t src_pitch, dst_pitch;
    size_t src_image_stride, dst_image_stride;
    const uint32_t *tab_x, *tab_y;
    const SmolDeviceLuts *luts;
    uint32_t first_row, n_rows, n_images;
    uint32_t lanes_per_col_log2;    /* G = 1 << this */
    uint32_t x_tiles;               /* items per output row */
    uint32_t rows_per_item, n_strips;   /* an item = one column tile x a strip of consecutive output rows (lean row loop: > 1 row) */
    uint32_t seg_bytes;             /* bytes per staging buffer (multiple of 16) */
    uint32_t alpha_shift, col_shift;/* bit positions in the packed source pixel */
    uint32_t sel_alpha, sel_c0, sel_c1, sel_c2;     /* PRMT selectors: that byte -> bits 0..7, zeros above */
    uint32_t sel_ac0, sel_ac1, sel_ac2;             /* PRMT selectors: (alpha << 8) | colour byte */
    const uint16_t *unpack_tab;                     /* TAB variants: 65536-entry composite unpack table */
    uint32_t acc_fits_24;           /* every accumulator lane stays below 2^24: one-instruction normalisation */
    /* LUTM = 3 (byte-addressed lane-replicated tables): PRMT selectors that build a table OFFSET
     * straight from the source pixel: byte 0 = the lane's slot, byte 1 = the alpha / colour byte */
    uint32_t sel_aaddr, sel_f0, sel_f1, sel_f2;
    uint32_t mul8_x, mul8_y;        /* span_mul << 8 (span_mul < 2^24 on every box axis) */
    uint32_t warps_lo;              /* warps whose staging buffers lie below the tables (window < 0x10000) */
    uint32_t unroll2;               /* walk the span two pixels per trip */
    uint32_t prefetch;              /* L2-prefetch the first window rows ahead of the dependency wait */
    uint32_t use_tma;               /* stage full windows with cp.async.bulk instead of cp.async (measurements) */
};

/* One 16-byte chunk of a staged row, of which the first src_bytes (1..16) lie inside the source
 * row.  A whole chunk is an asynchronous copy; a partial one (only the chunk that straddles the end
 * of a row whose length is not a multiple of 16) is copied byte by byte: cp.async with a short
 * src-size still touches all 16 source bytes, which on the image's last row may lie outside the
 * caller's buffer (compute-sanitizer memcheck, buffers allocated with no slack).  The bytes past
 * the row's end are never read, so they need no zero-fill.  The plain stores become visible to the
 * warp at the __syncwarp that follows the cp.async wait. */
__device__ __forceinline__ void cp_async_16 (uint32_t smem_addr, const void *gptr, uint32_t src_bytes)
{
    if (src_bytes >= 16)
        asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gptr) : "memory");
    else
    {
        const uint8_t *g = static_cast<const uint8_t *> (gptr);
        for (uint32_t i = 0; i < src_bytes; i++)
        {
            const uint32_t v = __ldg (g + i);
            asm volatile ("st.shared.u8 [%0], %1;" :: "r"(smem_addr + i), "r"(v) : "memory");
        }
    }
}
__device__ __forceinline__ void cp_async_16_full (uint32_t smem_addr, const void *gptr)
{
    asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit ()
{
    asm volatile ("cp.async.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait ()
{
    asm volatile ("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

/* Per-pixel lanes.  128bpp modes: v[0] = alpha lane, v[1..3] = colours (32-bit lanes).
 * 64bpp modes: v[0] = bytes 0 and 2, v[1] = bytes 1 and 3 of the source pixel (16-bit lanes). */
template <int MODE> struct BoxPx { uint32_t v[MODE >= BM_P8L_P ? 4 : 2]; };

/* LUTM: how the data tables are held in shared memory.
 *   0: one copy each (gathers suffer bank conflicts on random data)
 *   1: the 64K-entry composite table (fewest instructions, conflicts remain)
 *   2: 32 lane-private copies, word (index * 32 + lane): every lane always hits its own bank,
 *      so the gathers are conflict-free whatever the data */
template <int MODE, int LUTM>
__device__ __forceinline__ BoxPx<MODE>
box_unpack (uint32_t raw, const BoxParams &P, const uint32_t *__restrict__ sm_inv8, const uint32_t *__restrict__ sm_from,
            const uint16_t *__restrict__ sm_tab)
{
    BoxPx<MODE> r;
    constexpr int LSH = LUTM == 2 ? 5 : 0;      /* replicated tables: sm_inv8 / sm_from already point at this lane's column */

    if constexpr (LUTM == 1)
    {
        /* one shared-memory lookup per channel: the whole unpremultiply -> from_srgb ->
         * premultiply chain was folded into a 64K-entry table indexed by (alpha, value) */
        r.v[0] = __byte_perm (raw, 0, P.sel_alpha);
        r.v[1] = sm_tab[__byte_perm (raw, 0, P.sel_ac0)];
        r.v[2] = sm_tab[__byte_perm (raw, 0, P.sel_ac1)];
        r.v[3] = sm_tab[__byte_perm (raw, 0, P.sel_ac2)];
        return r;
    }

    if constexpr (MODE == BM_P8_P)
    {
        r.v[0] = raw & 0x00ff00ffu;
        r.v[1] = (raw >> 8) & 0x00ff00ffu;
    }
    else if constexpr (MODE == BM_P8_U)
    {
        /* ((c + 1) * (alpha + 1) - 1) >> 8 on the colour lanes (generic:238-244) */
        const uint32_t alpha = (raw >> P.alpha_shift) & 0xff, m = alpha + 1;
        uint32_t a = raw & 0x00ff00ffu, b = (raw >> 8) & 0x00ff00ffu;
        if (P.alpha_shift == 0)
        {
            a = (((((a & 0x00ff0000u) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff0000u) | alpha;
            b = (((b + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
        }
        else
        {
            a = (((a + 0x00010001u) * m - 0x00010001u) >> 8) & 0x00ff00ffu;
            b = (((((b & 0x000000ffu) + 0x00010001u) * m - 0x00010001u) >> 8) & 0x000000ffu) | (alpha << 16);
        }
        r.v[0] = a;
        r.v[1] = b;
    }
    else
    {
        const uint32_t alpha = __byte_perm (raw, 0, P.sel_alpha);
        uint32_t c[3] = { __byte_perm (raw, 0, P.sel_c0), __byte_perm (raw, 0, P.sel_c1), __byte_perm (raw, 0, P.sel_c2) };

        if constexpr (MODE == BM_P8L_P || MODE == BM_P8L_U)
        {
            const uint32_t m = (alpha << 3) + 1;
            if constexpr (MODE == BM_P8L_P)
            {
                /* unpremultiply (generic:227-236): sm_inv8 = inv_div_p8 << 3, result = byte 2 */
                const uint32_t inv8 = sm_inv8[alpha << LSH];
#pragma unroll
                for (int i = 0; i < 3; i++)
                    c[i] = __byte_perm (c[i] * inv8, 0, 0x4442);
            }
#pragma unroll
            for (int i = 0; i < 3; i++)
            {
                const uint32_t lin = sm_from[c[i] << LSH];          /* generic:185-199 */
                c[i] = (lin * m + (m - 1)) >> 11;                   /* ((lin + 1) * m - 1) >> 11 <= 2040, generic:261-269 */
            }
            r.v[0] = alpha;
        }
        else
        {
#pragma unroll
            for (int i = 0; i < 3; i++)
                c[i] = (MODE == BM_P16L_U ? sm_from[c[i] << LSH] : c[i]) * alpha;  /* generic:616-660, :708-752 */
            r.v[0] = (alpha << 8) | 0x80;
        }
        r.v[1] = c[0]; r.v[2] = c[1]; r.v[3] = c[2];
    }
    return r;
}

template <int MODE> __device__ __forceinline__ void box_add (BoxPx<MODE> &a, const BoxPx<MODE> &b)
{
#pragma unroll
    for (int i = 0; i < (MODE >= BM_P8L_P ? 4 : 2); i++) a.v[i] += b.v[i];
}

/* ((p * w) >> 8) & mask (generic:1177-1192) */
template <int MODE> __device__ __forceinline__ BoxPx<MODE> box_weight (const BoxPx<MODE> &p, uint32_t w)
{
    BoxPx<MODE> r;
#pragma unroll
    for (int i = 0; i < (MODE >= BM_P8L_P ? 4 : 2); i++)
        r.v[i] = ((p.v[i] * w) >> 8) & (MODE >= BM_P8L_P ? 0x00ffffffu : 0x00ff00ffu);
    return r;
}

/* scale_64bpp / scale_128bpp_half (generic:1231-1261), lane by lane.
 * (acc * mul + 2^23) >> 24 == high word of ((acc << 8) * mul + 2^31) when acc < 2^24: one
 * wide multiply-add instead of a 64-bit add and shift. */
template <int MODE> __device__ __forceinline__ BoxPx<MODE> box_scale (const BoxPx<MODE> &acc, uint32_t mul, bool fits24)
{
    BoxPx<MODE> r;
    if constexpr (MODE >= BM_P8L_P)
    {
        if (fits24)
        {
#pragma unroll
            for (int i = 0; i < 4; i++)
                r.v[i] = (uint32_t) (((uint64_t) (acc.v[i] << 8) * mul + 0x80000000ull) >> 32) & 0xffffu;
        }
        else
        {
#pragma unroll
            for (int i = 0; i < 4; i++)
                r.v[i] = (uint32_t) (((uint64_t) acc.v[i] * mul + (1u << 23)) >> 24) & 0xffffu;
        }
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            /* 16-bit lanes: always below 2^24 */
            const uint32_t lo = (uint32_t) (((uint64_t) ((acc.v[i] & 0xffffu) << 8) * mul + 0x80000000ull) >> 32) & 0xffu;
            const uint32_t hi = (uint32_t) (((uint64_t) ((acc.v[i] >> 16) << 8) * mul + 0x80000000ull) >> 32) & 0xffu;
            r.v[i] = lo | (hi << 16);
        }
    }
    return r;
}

/* LUTM = 3: unpack one source pixel and add it into the four accumulator lanes in one go, with
 * byte-addressed lane-replicated tables (linear-light modes only).
 *
 * The tables sit at fixed addresses of the CTA's shared-memory WINDOW (not of the kernel's dynamic
 * allocation), each on a 64 KB boundary:
 *   from table at 0x10000:    entry u at 0x10000 + u * 256 + lane * 4   (32-bit: from_srgb[u] + 1 for the
 *                             P8-LINEAR modes -- the premultiply wants lin + 1 --, from_srgb[u] for P16-LINEAR)
 *   inverse table at 0x20000: entry a at 0x20000 + a * 256 + lane * 8   (64-bit: { inv_div_p8[a] << 3, 8 a + 1 })
 * An entry's index is byte 1 of its address and the other three bytes are per-lane constants, so
 * ONE PRMT turns "the register whose byte holds the index" into the load address -- no shift, no
 * multiply-add, no base addition -- and every lane still reads its own bank whatever the data
 * (conflict-free gathers).  The warps' staging buffers fill the space below 0x10000 and above the
 * tables.
 *
 * Per channel of a premultiplied source pixel (generic:227-236, :185-199, :261-269):
 *   u = ((c * inv8) >> 16) & 0xff; lin = from_srgb[u]; out = ((lin + 1) * (8 a + 1) - 1) >> 11
 * which is PRMT (c), IMAD, PRMT (address), LDS, IMAD, LEA.HI (accumulate) here.
 * WEIGHTED: the pixel is an edge of the span, each lane is scaled by w / 256 first (generic:1177-1192).
 * from_y / inv_y: this lane's table addresses for index 0. */
/* most warps per CTA of the lean box kernel (one CTA per SM): fewer warps, more registers each */
#ifndef SMOL_BOX3_MAX_WARPS
#define SMOL_BOX3_MAX_WARPS 32
#endif
#define SMOL_BOX3_FROM_WIN 0x10000u
/* highest window address the dynamic allocation may start at: 1 KB reserved by the system + the
 * kernel's static shared memory (the warps' transfer barriers) */
#define SMOL_BOX3_DYN_WIN_MAX 0x700u

#define SMOL_BOX3_INV_WIN  0x20000u

/* PRMT with a selector known to have bit 3 of every nibble clear (__byte_perm masks its selector
 * with 0x7777 first, an extra instruction whenever the selector is a kernel parameter) */
__device__ __forceinline__ uint32_t prmt_raw (uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm ("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

/* OPAQUE (24bpp sources: alpha is 255 everywhere): unpremultiplying by 255 is the identity
 * ((c * (inv_div_p8[255] << 3)) >> 16 == c for every byte c), so the whole chain is a function of
 * the colour byte alone and the "from" table holds ((from_srgb[c] + 1) * 2041 - 1) >> 11 directly:
 * PRMT, LDS, add per channel. */
/* AP: where alpha sits in a 32bpp source pixel, known at compile time so that the PRMT selectors
 * are immediates (0: read them from the parameters; 1: last byte, colours in bytes 0..2; 2: first
 * byte, colours in bytes 1..3). */
template <int MODE, bool WEIGHTED, bool OPAQUE = false, int AP = 0>
__device__ __forceinline__ void
box3_accum (uint32_t raw, uint32_t w, uint32_t acc[4], const BoxParams &P, uint32_t from_y, uint32_t inv_y)
{
    const uint32_t sel_aaddr = AP == 1 ? 0x7634u : AP == 2 ? 0x7604u : P.sel_aaddr;
    const uint32_t sel_alpha = AP == 1 ? 0x4443u : AP == 2 ? 0x4440u : P.sel_alpha;
    const uint32_t sel_c0 = AP == 1 ? 0x4440u : AP == 2 ? 0x4441u : P.sel_c0;
    const uint32_t sel_c1 = AP == 1 ? 0x4441u : AP == 2 ? 0x4442u : P.sel_c1;
    const uint32_t sel_c2 = AP == 1 ? 0x4442u : AP == 2 ? 0x4443u : P.sel_c2;
    const uint32_t sel_f0 = AP == 1 ? 0x7604u : AP == 2 ? 0x7614u : P.sel_f0;
    const uint32_t sel_f1 = AP == 1 ? 0x7614u : AP == 2 ? 0x7624u : P.sel_f1;
    const uint32_t sel_f2 = AP == 1 ? 0x7624u : AP == 2 ? 0x7634u : P.sel_f2;
    static_assert (MODE == BM_P8L_P || MODE == BM_P8L_U || MODE == BM_P16L_U, "linear-light modes only");
    static_assert (!OPAQUE || MODE == BM_P8L_P, "opaque shortcut: premultiplied linear-light unpack");
    auto add = [&] (uint32_t &a, uint32_t v, int shift)
    {
        if constexpr (WEIGHTED)
            a += ((v >> shift) * w) >> 8;
        else
            a += v >> shift;
    };

    if constexpr (OPAQUE)
    {
        add (acc[0], 255u, 0);
        add (acc[1], lds_u32 (prmt_raw (raw, from_y, sel_f0)), 0);
        add (acc[2], lds_u32 (prmt_raw (raw, from_y, sel_f1)), 0);
        add (acc[3], lds_u32 (prmt_raw (raw, from_y, sel_f2)), 0);
    }
    else if constexpr (MODE == BM_P8L_P)
    {
        const uint2 im = lds_u64 (prmt_raw (raw, inv_y, sel_aaddr));
        const uint32_t c[3] = { prmt_raw (raw, 0, sel_c0), prmt_raw (raw, 0, sel_c1), prmt_raw (raw, 0, sel_c2) };
        add (acc[0], im.y, 3);                                      /* alpha = (8 a + 1) >> 3 */
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            const uint32_t lin1 = lds_u32 (__byte_perm (c[i] * im.x, from_y, 0x7624));
            add (acc[i + 1], lin1 * im.y - 1, 11);
        }
    }
    else
    {
        const uint32_t alpha = prmt_raw (raw, 0, sel_alpha);
        const uint32_t lin[3] = { lds_u32 (prmt_raw (raw, from_y, sel_f0)), lds_u32 (prmt_raw (raw, from_y, sel_f1)),
                                  lds_u32 (prmt_raw (raw, from_y, sel_f2)) };
        if constexpr (MODE == BM_P8L_U)
        {
            const uint32_t m = alpha * 8 + 1;
            add (acc[0], alpha, 0);
#pragma unroll
            for (int i = 0; i < 3; i++)
                add (acc[i + 1], lin[i] * m - 1, 11);
        }
        else
        {
            add (acc[0], (alpha << 8) | 0x80, 0);                   /* generic:616-625 */
#pragma unroll
            for (int i = 0; i < 3; i++)
                add (acc[i + 1], lin[i] * alpha, 0);
        }
    }
}

/* LUTM = 1: one big CTA per SM (128 KB composite table + the warps' staging buffers);
 * LUTM = 2: 512-thread CTAs with lane-replicated LUTs; LUTM = 0: 256-thread CTAs, plain LUTs;
 * LUTM = 3: one big CTA per SM, byte-addressed lane-replicated LUTs (see box3_accum). */
/* Bulk asynchronous copy (the TMA unit, `cp.async.bulk`) of a contiguous, 16-byte-aligned run of
 * bytes into shared memory, completion signalled on an mbarrier by byte count.  One elected lane
 * issues it; the warp waits on the barrier's phase parity. */
__device__ __forceinline__ void mbar_init (uint32_t mbar, uint32_t arrivals)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(arrivals) : "memory");
}

__device__ __forceinline__ void tma_load_bytes (uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t mbar)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_wait (uint32_t mbar, uint32_t parity)
{
    asm volatile ("{\n"
                  ".reg .pred p;\n"
                  "SMOL_MBAR_WAIT:\n"
                  "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                  "@p bra SMOL_MBAR_DONE;\n"
                  "bra SMOL_MBAR_WAIT;\n"
                  "SMOL_MBAR_DONE:\n"
                  "}" :: "r"(mbar), "r"(parity) : "memory");
}

/* RS ("row shift", LUTM = 3 only): source rows need not start on 16-byte boundaries (32bpp: any
 * 4-byte-aligned base and pitch; 24bpp: any).  The copies stay 16-byte cp.async on chunks aligned
 * in GLOBAL memory; what changes from row to row is where the row's first byte lands in the
 * staging buffer, a per-row constant added to the lane's walk addresses.  Chunks that straddle the
 * row's start or end are copied byte-exactly, so nothing outside the row is ever touched. */
/* G1: one lane per column (spans of up to 16 pixels: cfg 3 and most thumbnails), known at compile
 * time: constant loop strides, no lane-group bookkeeping for the compiler to keep alive or rebuild in
 * the row loop (the kernel runs at its 64-register cap). */
template <int MODE, int LUTM, int BI, bool RS = false, bool G1 = false>
__global__ void __launch_bounds__ (LUTM == 3 ? SMOL_BOX3_MAX_WARPS * 32 : LUTM == 1 ? 1024 : LUTM == 2 ? 512 : 256, LUTM == 1 || LUTM == 3 ? 1 : LUTM == 2 ? 2 : 5)
smol_box_kernel (const BoxParams P)
{
    static_assert (!RS || LUTM == 3, "row-shifted staging exists for the lean row loop only");
    static_assert (!G1 || (LUTM == 3 && !RS), "the one-lane-per-column instance exists for the aligned lean row loop only");
    extern __shared__ __align__ (16) uint8_t sm_dyn[];
    __shared__ uint32_t sm_inv8_plain[LUTM == 0 ? 256 : 1];
    __shared__ uint32_t sm_from_plain[LUTM == 0 ? 256 : 1];
    __shared__ __align__ (8) uint64_t sm_mbar[LUTM == 3 && !RS ? 2 * SMOL_BOX3_MAX_WARPS : 1];   /* two per warp: one per staging slot */
#ifdef SMOL_BOX_TMA
    constexpr bool TMA = LUTM == 3 && !RS;
#else
    constexpr bool TMA = false;      /* see use_tma below: measured slower, and even unused its scaffolding costs 2 us */
#endif
    constexpr bool S128 = MODE >= BM_P8L_P;
    constexpr bool TAB = LUTM == 1;
    constexpr bool OPAQUE = LUTM == 3 && BI == 3 && MODE == BM_P8L_P;   /* see box3_accum */
    constexpr bool NEED_INV = MODE == BM_P8L_P && !OPAQUE;
    /* 32bpp premultiplied linear light: rows whose staged pixels all have alpha 255 (photographs in an
     * RGBA container) take the one-table-read-per-channel walk of the 24bpp case; see the row loop */
#ifdef SMOL_BOX_NO_SPEC
    constexpr bool SPEC_OPAQUE = false;
#else
    constexpr bool SPEC_OPAQUE = LUTM == 3 && NEED_INV && BI == 4;
#endif
    constexpr bool NEED_FROM = MODE == BM_P8L_P || MODE == BM_P8L_U || MODE == BM_P16L_U;
    constexpr uint32_t REP_BYTES = LUTM == 2 ? ((NEED_INV ? 32768u : 0u) + (NEED_FROM ? 32768u : 0u)) : 0u;
    constexpr uint32_t TAB_BYTES = TAB ? 65536 * 2 : REP_BYTES;
    /* LUTM = 3: first window address above the tables */
    constexpr uint32_t WIN_HI = NEED_INV ? SMOL_BOX3_INV_WIN + 65536u : SMOL_BOX3_FROM_WIN + 65536u;
    const SmolJobDesc &d = P.d;
    const uint16_t *sm_tab = reinterpret_cast<const uint16_t *> (sm_dyn);
    const uint32_t *sm_inv8 = sm_inv8_plain, *sm_from = sm_from_plain;

    pdl_launch_dependents ();
    if constexpr (TAB)
    {
        /* library-owned constant data: may be fetched before the dependency wait */
        const uint32_t tab_addr = (uint32_t) __cvta_generic_to_shared (sm_dyn);
        for (uint32_t i = threadIdx.x; i < TAB_BYTES / 16; i += blockDim.x)
            cp_async_16 (tab_addr + 16 * i, reinterpret_cast<const uint8_t *> (P.unpack_tab) + 16 * i, 16);
        cp_async_commit ();
        cp_async_wait<0> ();
    }
    else if constexpr (LUTM == 2)
    {
        /* word (index * 32 + lane) of each replicated table */
        uint32_t *rep_from = reinterpret_cast<uint32_t *> (sm_dyn);
        uint32_t *rep_inv = rep_from + (NEED_FROM ? 8192 : 0);
        for (uint32_t i = threadIdx.x; i < 8192; i += blockDim.x)
        {
            if constexpr (NEED_FROM)
                rep_from[i] = P.luts->from_srgb[i >> 5];
            if constexpr (NEED_INV)
                rep_inv[i] = P.luts->inv_div_p8[i >> 5] << 3;
        }
        sm_from = rep_from + (threadIdx.x & 31);
        sm_inv8 = rep_inv + (threadIdx.x & 31);
    }
    else if constexpr (LUTM == 3 && !NEED_FROM)
    {
        /* the lean row loop without data tables (8-bit and 16-bit premultiplied intermediates) */
    }
    else if constexpr (LUTM == 3)
    {
        /* layout: see box3_accum; window address -> offset inside the dynamic allocation */
        const uint32_t dyn_win = (uint32_t) __cvta_generic_to_shared (sm_dyn) & 0x00ffffffu;
        if (dyn_win > SMOL_BOX3_DYN_WIN_MAX)
            __trap ();                  /* the host sized the low staging region for a start at or below this */
        uint32_t *t_from = reinterpret_cast<uint32_t *> (sm_dyn + (SMOL_BOX3_FROM_WIN - dyn_win));
        uint2 *t_inv = reinterpret_cast<uint2 *> (sm_dyn + (SMOL_BOX3_INV_WIN - dyn_win));
        for (uint32_t i = threadIdx.x; i < 8192; i += blockDim.x)
        {
            const uint32_t e = i >> 5, l = i & 31;
            const uint32_t lin = P.luts->from_srgb[e];
            t_from[e * 64 + l] = OPAQUE ? ((lin + 1) * 2041u - 1) >> 11 : lin + (MODE == BM_P16L_U ? 0u : 1u);
            if constexpr (SPEC_OPAQUE)
                t_from[e * 64 + 32 + l] = ((lin + 1) * 2041u - 1) >> 11;    /* the chain's result for alpha = 255 (see box3_accum, OPAQUE) */
            if constexpr (NEED_INV)
                t_inv[e * 32 + l] = make_uint2 (P.luts->inv_div_p8[e] << 3, e * 8 + 1);
        }
    }
    else
    {
        if constexpr (NEED_INV)
            sm_inv8_plain[threadIdx.x] = P.luts->inv_div_p8[threadIdx.x] << 3;
        if constexpr (NEED_FROM)
            sm_from_plain[threadIdx.x] = P.luts->from_srgb[threadIdx.x];
    }
    const uint32_t mbar0 = (uint32_t) __cvta_generic_to_shared (&sm_mbar[TMA ? 2 * (threadIdx.x >> 5) : 0]);
    uint32_t tma_parity = 0;        /* bit s: phase parity the warp waits for next on slot s */
    if constexpr (TMA)
    {
        if ((threadIdx.x & 31) == 0)
        {
            mbar_init (mbar0, 1);
            mbar_init (mbar0 + 8, 1);
            asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads ();
    bool waited = false;            /* the dependency wait happens at the warp's first item, after an L2 prefetch */
    uint32_t opaque_backoff = 0;    /* rows to go before the warp tries the opaque walk again */

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t G = G1 ? 1u : 1u << P.lanes_per_col_log2, g = G1 ? 0u : lane & (G - 1);
    const uint32_t cols_per_item = G1 ? 32u : 32u >> P.lanes_per_col_log2;
    uint8_t *bufs = sm_dyn + TAB_BYTES + (size_t) warp * 2 * P.seg_bytes;
    if constexpr (LUTM == 3 && NEED_FROM)
    {
        /* staging buffers: the first warps_lo warps below the tables, the others above them */
        const uint32_t dyn_win = (uint32_t) __cvta_generic_to_shared (sm_dyn) & 0x00ffffffu;
        bufs = warp < P.warps_lo ? sm_dyn + (size_t) warp * 2 * P.seg_bytes
                                 : sm_dyn + (WIN_HI - dyn_win) + (size_t) (warp - P.warps_lo) * 2 * P.seg_bytes;
    }
    const uint32_t bufs_addr = (uint32_t) __cvta_generic_to_shared (bufs);
    const uint32_t row_bytes = d.w_in * BI;

    const uint32_t items_per_image = P.x_tiles * P.n_strips;
    const uint32_t n_items = items_per_image * P.n_images;
    const uint32_t warp_stride = gridDim.x * (blockDim.x >> 5);

    for (uint32_t item = blockIdx.x * (blockDim.x >> 5) + warp; item < n_items; item += warp_stride)
    {
        const uint32_t img = item / items_per_image;
        const uint32_t rem = item - img * items_per_image;
        const uint32_t strip = rem / P.x_tiles, xt = rem - strip * P.x_tiles;
        const uint32_t yl = strip * P.rows_per_item;        /* the item's first output row */
        const uint32_t y = P.first_row + yl;
        const uint32_t x_first = xt * cols_per_item;
        const uint32_t x_last = min (x_first + cols_per_item, d.w_out) - 1;
        uint32_t x = x_first + (G1 ? lane : lane >> P.lanes_per_col_log2);
        const bool store = x <= x_last && g == 0;
        x = min (x, x_last);

        /* horizontal span of this lane's column (generic:1427-1556 in absolute offsets) */
        const uint32_t ex0 = __ldg (&P.tab_x[x]), ex1 = __ldg (&P.tab_x[x + 1]);
        const uint32_t hL = SMOL_TAB_OFS (ex0), hR = SMOL_TAB_OFS (ex1), wr = SMOL_TAB_F (ex0);
        const uint32_t wl = x == 0 ? 256u : 255u - SMOL_TAB_F (__ldg (&P.tab_x[x - 1]));
        /* segment of the source rows the whole item reads, as an aligned byte window */
        const uint32_t sx0 = SMOL_TAB_OFS (__ldg (&P.tab_x[x_first]));
        const uint32_t sx1 = SMOL_TAB_OFS (__ldg (&P.tab_x[x_last + 1]));
        const uint32_t win0 = (sx0 * BI) & ~15u;
        const uint32_t n_chunks = (((sx1 + 1) * BI + 15) & ~15u) - win0 >> 4;

        /* vertical span (generic:2112-2161 / :2198-2260) */
        const uint32_t ey0 = __ldg (&P.tab_y[y]), ey1 = __ldg (&P.tab_y[y + 1]);
        const uint32_t T = SMOL_TAB_OFS (ey0), B = SMOL_TAB_OFS (ey1), Fy = SMOL_TAB_F (ey0);
        const uint32_t w1 = y == 0 ? 256u : 255u - SMOL_TAB_F (__ldg (&P.tab_y[y - 1]));
        const uint32_t w2 = S128 ? Fy - 1 : Fy;     /* 128bpp weighs the trailing row by F - 1 (generic:2247-2249) */
        const uint32_t r_end = Fy > 0 ? B : B - 1;  /* last source row that contributes */

        const uint8_t *src = P.src + (size_t) img * P.src_image_stride + win0;

        /* This lane's chunks of the window (lane, lane + 32, lane + 64; longer windows take the
         * generic loop) are the same for every row: work out once how many bytes of each lie inside
         * the row.  Chunks wholly past the row's end are skipped -- nothing ever reads them. */
        uint32_t cvalid[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            const uint32_t k = lane + 32 * c, ofs = win0 + 16 * k;
            cvalid[c] = (k < n_chunks && ofs < row_bytes) ? min (16u, row_bytes - ofs) : 0u;
        }
        const uint8_t *grow = src + (size_t) T * P.src_pitch + 16 * lane;
        const uint32_t sbuf = bufs_addr + 16 * lane;

        auto prefetch = [&] (uint32_t slot)
        {
            /* copies the row `grow` points at, then advances it to the next row */
            const uint32_t sbase = sbuf + slot * P.seg_bytes;
            if (cvalid[0])
                cp_async_16 (sbase, grow, cvalid[0]);
            if (cvalid[1])
                cp_async_16 (sbase + 512, grow + 512, cvalid[1]);
            if (cvalid[2])
                cp_async_16 (sbase + 1024, grow + 1024, cvalid[2]);
            if (n_chunks > 96)
            {
                for (uint32_t k = lane + 96; k < n_chunks; k += 32)
                {
                    const uint32_t ofs = win0 + 16 * k;
                    if (ofs < row_bytes)
                        cp_async_16 (sbase + 16 * (k - lane), grow + 16 * (k - lane), min (16u, row_bytes - ofs));
                }
            }
            cp_async_commit ();
            grow += P.src_pitch;
        };

        BoxPx<MODE> vacc;
#pragma unroll
        for (int i = 0; i < (S128 ? 4 : 2); i++) vacc.v[i] = 0;

        if (!waited)
        {
            if (P.prefetch)
            {
                /* the first rows of the warp's first window, one 128-byte line per lane (see prefetch_l2) */
                const uint32_t lines = (n_chunks + 7) / 8 + 1, r = lane / lines, l = lane - r * lines;
                if (T + r <= r_end)
                    prefetch_l2 (src + (size_t) (T + r) * P.src_pitch + min (128u * l, 16u * n_chunks - 1u));
            }
            pdl_wait ();
            waited = true;
        }

        if constexpr (LUTM == 3)
        {
            /* Lean row loop for the byte-addressed tables: the window's chunks are copied whole
             * when the window lies inside the row (all items but a row's last), the span is walked
             * with a shared-memory byte offset, pixels are unpacked straight into the accumulators,
             * and normalisation is one multiply-high per lane ((acc * mul + 2^23) >> 24 ==
             * hi32 (acc * (mul << 8) + 2^31) for any 32-bit acc). */
            const bool win_full = win0 + 16 * n_chunks <= row_bytes;
            const uint32_t win_hi = bufs_addr & 0xff000000u;        /* the CTA's window */
            const uint32_t from_y = win_hi | SMOL_BOX3_FROM_WIN | (lane * 4), inv_y = win_hi | SMOL_BOX3_INV_WIN | (lane * 8);
            /* byte offsets into the staging buffer of the first whole pixel of this lane, of the
             * span's end, and of the two edge pixels (32bpp; 24bpp goes through fetch3) */
            const uint32_t o_first = (hL + 1 + g) * 4 - win0, o_end = hR * 4 - win0;
            const uint32_t o_left = hL * 4 - win0;
            uint32_t cur = 0;                                       /* byte offset of the slot in use: 0 or seg_bytes */

            /* SMOL_BOX_TMA=1: windows that lie wholly inside the row travel as ONE bulk copy issued by
             * lane 0 (TMA) -- no per-lane address arithmetic, no LSU work for the copy -- and the warp
             * waits on the slot's barrier instead of its own cp.async group.  Measured on B200
             * (8K -> 800x450): 59.0 us against 56.5 us with cp.async for the linear-light job, 35.5
             * against 32.3 us without tables: a 1.2 KB copy per warp and row is too small for the
             * bulk path's fixed latency, and the barrier polls cost issue slots the kernel is short
             * of.  Compiled in with -DSMOL_BOX_TMA only (SMOL_NVCC_FLAGS); kept for measurements. */
            const bool use_tma = TMA && win_full && P.use_tma;
            auto prefetch3 = [&] (uint32_t slot_ofs)
            {
                const uint32_t sbase = sbuf + slot_ofs;
                if (use_tma)
                {
                    if (lane == 0)
                        tma_load_bytes (bufs_addr + slot_ofs, grow, 16 * n_chunks, mbar0 + (slot_ofs ? 8u : 0u));
                    grow += P.src_pitch;
                    return;
                }
                if (win_full)
                {
                    if (cvalid[0])
                        cp_async_16_full (sbase, grow);
                    if (cvalid[1])
                        cp_async_16_full (sbase + 512, grow + 512);
                    if (cvalid[2])
                        cp_async_16_full (sbase + 1024, grow + 1024);
                }
                else
                {
                    if (cvalid[0])
                        cp_async_16 (sbase, grow, cvalid[0]);
                    if (cvalid[1])
                        cp_async_16 (sbase + 512, grow + 512, cvalid[1]);
                    if (cvalid[2])
                        cp_async_16 (sbase + 1024, grow + 1024, cvalid[2]);
                }
                if (n_chunks > 96)
                {
                    for (uint32_t k = lane + 96; k < n_chunks; k += 32)
                    {
                        const uint32_t ofs = win0 + 16 * k;
                        if (ofs < row_bytes)
                            cp_async_16 (sbase + 16 * (k - lane), grow + 16 * (k - lane), min (16u, row_bytes - ofs));
                    }
                }
                cp_async_commit ();
                grow += P.src_pitch;
            };

            /* RS: row-relative offset (may be negative) of the first staged byte of source row r,
             * and the copy of that row's chunks */
            const uint8_t *img_base = P.src + (size_t) img * P.src_image_stride;
            const uint32_t sx_b = sx0 * BI, end_b = (sx1 + 1) * BI;
            auto staged_from = [&] (uint32_t r) -> int32_t
            {
                const uint32_t A = (uint32_t) reinterpret_cast<uintptr_t> (img_base + (size_t) r * P.src_pitch) & 15u;
                return (int32_t) ((A + sx_b) & ~15u) - (int32_t) A;
            };
            auto prefetch3s = [&] (uint32_t slot_ofs, uint32_t r)
            {
                const uint8_t *rowp = img_base + (size_t) r * P.src_pitch;
                const int32_t w0 = staged_from (r);
                const uint32_t nch = ((uint32_t) ((int32_t) end_b - w0) + 15u) >> 4;
                for (uint32_t k = lane; k < nch; k += 32)
                {
                    const int32_t ofs = w0 + 16 * (int32_t) k;
                    const uint32_t sa = bufs_addr + slot_ofs + 16 * k;
                    if (ofs >= 0 && ofs + 16 <= (int32_t) row_bytes)
                        cp_async_16_full (sa, rowp + ofs);
                    else
                    {
                        const int32_t lo = ofs < 0 ? -ofs : 0, hi = min (16, (int32_t) row_bytes - ofs);
                        for (int32_t b = lo; b < hi; b++)
                        {
                            const uint32_t v = __ldg (rowp + ofs + b);
                            asm volatile ("st.shared.u8 [%0], %1;" :: "r"(sa + b), "r"(v) : "memory");
                        }
                    }
                }
                cp_async_commit ();
            };

            /* The item's output rows share their boundary source rows (row B of one output row is
             * row T of the next, weighted F for the one and 255 - F for the other): the strip's
             * source rows T .. r_last are walked ONCE, in one uninterrupted double-buffered stream,
             * and a boundary row's horizontal result is handed on instead of being recomputed
             * (the reference unpacks and filters that row twice, generic:2198-2260). */
            const uint32_t yl_stop = min (yl + P.rows_per_item, P.n_rows);
            uint32_t yl_cur = yl, B_cur = B, rend_cur = r_end, w1_cur = w1, w2_cur = w2, Fy_cur = Fy;
            uint32_t r_last = r_end;
            if (yl_stop - yl > 1)
            {
                const uint32_t ya = P.first_row + yl_stop - 1;
                const uint32_t ea = __ldg (&P.tab_y[ya]), eb = __ldg (&P.tab_y[ya + 1]);
                r_last = SMOL_TAB_F (ea) > 0 ? SMOL_TAB_OFS (eb) : SMOL_TAB_OFS (eb) - 1;
            }
            bool top_pending = true;        /* the current output row's first source row is still to come */

            auto emit_row = [&] ()
            {
                BoxPx<MODE> fin;
                if constexpr (S128)
                {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        fin.v[i] = (uint32_t) (((uint64_t) vacc.v[i] * P.mul8_y + 0x80000000ull) >> 32) & 0xffffu;
                }
                else
                    fin = box_scale<MODE> (vacc, d.span_mul_y, true);
                if (store)
                {
                    uint32_t packed;
                    if constexpr (S128)
                    {
                        Px<true> o;
                        o.w[0] = (uint64_t) fin.v[0] | ((uint64_t) fin.v[1] << 32);
                        o.w[1] = (uint64_t) fin.v[2] | ((uint64_t) fin.v[3] << 32);
                        packed = pack_px<true> (o, d, P.luts);
                    }
                    else
                    {
                        const uint32_t bytes = fin.v[0] | (fin.v[1] << 8);
                        const uint32_t alpha = (bytes >> P.alpha_shift) & 0xff;
                        const uint32_t cols = bytes >> P.col_shift;
                        Px<false> o;
                        o.w[0] = (uint64_t) alpha | ((uint64_t) (cols & 0xff) << 16) | ((uint64_t) ((cols >> 8) & 0xff) << 32)
                                 | ((uint64_t) ((cols >> 16) & 0xff) << 48);
                        packed = pack_px<false> (o, d, P.luts);
                    }
                    uint8_t *o8 = P.dst + (size_t) img * P.dst_image_stride + (size_t) yl_cur * P.dst_pitch + (size_t) x * d.bpp_out;
                    store_raw_px (o8, packed, d.bpp_out);
                }
            };

            if constexpr (RS)
                prefetch3s (0, T);
            else
                prefetch3 (0);
            for (uint32_t r = T; r <= r_last; r++)
            {
                if (r < r_last)
                {
                    if constexpr (RS)
                        prefetch3s (P.seg_bytes - cur, r + 1);
                    else
                        prefetch3 (P.seg_bytes - cur);
                    if (!use_tma)
                        cp_async_wait<1> ();
                }
                else if (!use_tma)
                    cp_async_wait<0> ();
                if (use_tma)
                {
                    const uint32_t slot = cur ? 1u : 0u;
                    mbar_wait (mbar0 + 8 * slot, (tma_parity >> slot) & 1u);
                    tma_parity ^= 1u << slot;
                }
                __syncwarp ();

                /* window address of the byte staged for row-relative offset win0 */
                uint32_t row = bufs_addr + cur;
                if constexpr (RS)
                    row += (uint32_t) ((int32_t) win0 - staged_from (r));
                auto fetch3 = [&] (uint32_t j) -> uint32_t
                {
                    /* RS: `row` is not word aligned for 24bpp rows; split address and shift from the true staging offset */
                    const uint32_t b = RS ? j * 3 - win0 + (row - (bufs_addr + cur)) : j * 3 - win0;
                    const uint32_t a = (RS ? bufs_addr + cur : row) + (b & ~3u);
                    return __funnelshift_r (lds_u32_ordered (a), lds_u32_ordered (a + 4), (b & 3) * 8) | 0xff000000u;
                };
                BoxPx<MODE> acc;
#pragma unroll
                for (int i = 0; i < (S128 ? 4 : 2); i++) acc.v[i] = 0;
                /* The row's pixel walk, instantiated per alpha position for the 32bpp table modes
                 * (immediate PRMT selectors; one uniform branch per row picks the instance). */
                uint32_t alpha_and = 0xffffffffu;       /* AND of every pixel the opaque walk consumed */
                auto walk_row = [&] (auto ap_tag, auto opaque_tag)
                {
                constexpr int AP = decltype (ap_tag)::value;
                constexpr bool SPEC = decltype (opaque_tag)::value;
                constexpr bool OPQ = OPAQUE || SPEC;
                (void) AP;
                const uint32_t tab_y = SPEC ? from_y + 128u : from_y;
                /* unpack one pixel into the accumulators, table modes through box3_accum */
                auto accum = [&] (uint32_t raw)
                {
                    if constexpr (SPEC)
                        alpha_and &= raw;
                    if constexpr (NEED_FROM)
                        box3_accum<MODE, false, OPQ, AP> (raw, 0, acc.v, P, tab_y, inv_y);
                    else
                        box_add<MODE> (acc, box_unpack<MODE, 0> (raw, P, nullptr, nullptr, nullptr));
                };
                auto accum_w = [&] (uint32_t raw, uint32_t w)
                {
                    if constexpr (SPEC)
                    {
                        if (w > 0)
                            alpha_and &= raw;
                    }
                    if constexpr (NEED_FROM)
                        box3_accum<MODE, true, OPQ, AP> (raw, w, acc.v, P, tab_y, inv_y);
                    else
                        box_add<MODE> (acc, box_weight<MODE> (box_unpack<MODE, 0> (raw, P, nullptr, nullptr, nullptr), w));
                };

                if constexpr (BI == 4)
                {
                    uint32_t a = row + o_first;
                    const uint32_t a_end = row + o_end;
                    if (P.unroll2)
                    {
                        /* two pixels per trip: two independent table chains in flight */
                        for (; a + 4 * G < a_end; a += 8 * G)
                        {
                            const uint32_t raw0 = lds_u32_ordered (a), raw1 = lds_u32_ordered (a + 4 * G);
                            accum (raw0);
                            accum (raw1);
                        }
                        if (a < a_end)
                            accum (lds_u32_ordered (a));
                    }
                    else
                    {
                        for (; a < a_end; a += 4 * G)
                            accum (lds_u32_ordered (a));
                    }
                    if (G == 1)
                    {
                        /* both edge pixels at once: two independent unpack chains in flight, no
                         * divergent branch (a right edge of weight 0 reads a staged byte that may
                         * lie past the row's end and contributes nothing) */
                        const uint32_t raw_l = lds_u32_ordered (row + o_left), raw_r = lds_u32_ordered (row + o_end);
                        accum_w (raw_l, wl);
                        accum_w (raw_r, wr);
                    }
                    else
                    {
                        if (g == 0)
                            accum_w (lds_u32_ordered (row + o_left), wl);
                        if (g == G - 1 && wr > 0)
                            accum_w (lds_u32_ordered (row + o_end), wr);
                    }
                }
                else
                {
                    for (uint32_t j = hL + 1 + g; j < hR; j += G)
                        accum (fetch3 (j));
                    if (g == 0)
                        accum_w (fetch3 (hL), wl);
                    if (g == G - 1 && wr > 0)
                        accum_w (fetch3 (hR), wr);
                }
                };
                bool row_done = false;
                if constexpr (SPEC_OPAQUE)
                {
                    /* Speculate that the row's pixels are opaque: walk it with the alpha = 255 table while
                     * ANDing the pixels together; if any lane met another alpha the row is redone the
                     * general way and the warp does not try again for a while (random-alpha input pays
                     * one wasted cheap walk per 64 rows, opaque input runs close to the 24bpp rate).
                     * Measured (8K RGBA -> 800x450, linear light): opaque 56.5 -> 46.4 us per frame,
                     * random alpha 56.5 -> 57.5 us (the second walk's code costs the first a little). */
                    if (opaque_backoff == 0)
                    {
                        if (P.alpha_shift == 0)
                            walk_row (std::integral_constant<int, 2> {}, std::true_type {});
                        else
                            walk_row (std::integral_constant<int, 1> {}, std::true_type {});
                        const uint32_t amask = 0xffu << P.alpha_shift;
                        row_done = __all_sync (0xffffffffu, (alpha_and & amask) == amask);
                        if (!row_done)
                        {
                            opaque_backoff = 64;
#pragma unroll
                            for (int i = 0; i < 4; i++) acc.v[i] = 0;
                        }
                    }
                    else
                        opaque_backoff--;
                }
                if (!row_done)
                {
                    if constexpr (NEED_FROM && BI == 4)
                    {
                        if (P.alpha_shift == 0)
                            walk_row (std::integral_constant<int, 2> {}, std::false_type {});
                        else
                            walk_row (std::integral_constant<int, 1> {}, std::false_type {});
                    }
                    else
                        walk_row (std::integral_constant<int, 0> {}, std::false_type {});
                }
                for (uint32_t m = G >> 1; m; m >>= 1)
                {
#pragma unroll
                    for (int i = 0; i < (S128 ? 4 : 2); i++)
                        acc.v[i] += __shfl_xor_sync (0xffffffffu, acc.v[i], m);
                }

                /* scale_128bpp_half / scale_64bpp (generic:1231-1261): in the P8-LINEAR modes the
                 * 16-bit mask cannot bite (lanes are averages of values below 2^11) */
                BoxPx<MODE> h;
                if constexpr (S128)
                {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        h.v[i] = (uint32_t) (((uint64_t) acc.v[i] * P.mul8_x + 0x80000000ull) >> 32);
                        if constexpr (MODE == BM_P16L_U || MODE == BM_P16_U)
                            h.v[i] &= 0xffffu;
                    }
                }
                else
                    h = box_scale<MODE> (acc, d.span_mul_x, true);
                if (top_pending)
                {
                    box_add<MODE> (vacc, box_weight<MODE> (h, w1_cur));
                    top_pending = false;
                }
                else if (r == B_cur)
                    box_add<MODE> (vacc, box_weight<MODE> (h, w2_cur));
                else
                    box_add<MODE> (vacc, h);

                if (r == rend_cur)
                {
                    emit_row ();
                    if (++yl_cur < yl_stop)
                    {
                        const uint32_t yn = P.first_row + yl_cur;
                        const uint32_t e0 = __ldg (&P.tab_y[yn]), e1 = __ldg (&P.tab_y[yn + 1]);
                        const uint32_t F_new = SMOL_TAB_F (e0);
                        w1_cur = 255u - Fy_cur;
                        Fy_cur = F_new;
                        const bool shared = r == B_cur;         /* this row also opens the next output row */
                        B_cur = SMOL_TAB_OFS (e1);
                        w2_cur = S128 ? F_new - 1 : F_new;
                        rend_cur = F_new > 0 ? B_cur : B_cur - 1;
#pragma unroll
                        for (int i = 0; i < (S128 ? 4 : 2); i++) vacc.v[i] = 0;
                        if (shared)
                            box_add<MODE> (vacc, box_weight<MODE> (h, w1_cur));
                        top_pending = !shared;
                    }
                }
                cur = P.seg_bytes - cur;
                __syncwarp ();      /* everyone is done with this slot before it is refilled */
            }
            continue;               /* every row of the item has been stored */
        }
        else
        {
        prefetch (0);
        for (uint32_t r = T; r <= r_end; r++)
        {
            const uint32_t slot = (r - T) & 1;
            if (r < r_end)
            {
                prefetch (slot ^ 1);
                cp_async_wait<1> ();
            }
            else
                cp_async_wait<0> ();
            __syncwarp ();

            const uint32_t *words = reinterpret_cast<const uint32_t *> (bufs + slot * P.seg_bytes);
            /* source pixel j of this row: 32bpp straight from the window; 24bpp as a funnel shift
             * over the two words that hold its three bytes, alpha byte forced to 0xff */
            auto fetch = [&] (uint32_t j) -> uint32_t
            {
                if constexpr (BI == 4)
                    return words[j - (win0 >> 2)];
                else
                {
                    const uint32_t b = j * 3 - win0;
                    return __funnelshift_r (words[b >> 2], words[(b >> 2) + 1], (b & 3) * 8) | 0xff000000u;
                }
            };
            BoxPx<MODE> acc;
#pragma unroll
            for (int i = 0; i < (S128 ? 4 : 2); i++) acc.v[i] = 0;

            /* whole pixels hL + 1 .. hR - 1, interleaved over the G lanes of the column */
            for (uint32_t j = hL + 1 + g; j < hR; j += G)
                box_add<MODE> (acc, box_unpack<MODE, LUTM> (fetch (j), P, sm_inv8, sm_from, sm_tab));
            /* edge pixels: lane 0 of the column takes the left one, the last lane the right one */
            if (g == 0)
                box_add<MODE> (acc, box_weight<MODE> (box_unpack<MODE, LUTM> (fetch (hL), P, sm_inv8, sm_from, sm_tab), wl));
            if (g == G - 1 && wr > 0)
                box_add<MODE> (acc, box_weight<MODE> (box_unpack<MODE, LUTM> (fetch (hR), P, sm_inv8, sm_from, sm_tab), wr));
            for (uint32_t m = G >> 1; m; m >>= 1)
            {
#pragma unroll
                for (int i = 0; i < (S128 ? 4 : 2); i++)
                    acc.v[i] += __shfl_xor_sync (0xffffffffu, acc.v[i], m);
            }

            BoxPx<MODE> h = box_scale<MODE> (acc, d.span_mul_x, P.acc_fits_24 != 0);
            if (r == T)
                h = box_weight<MODE> (h, w1);
            else if (r == B)
                h = box_weight<MODE> (h, w2);
            box_add<MODE> (vacc, h);
            __syncwarp ();      /* everyone is done with this slot before it is refilled */
        }
        }

        const BoxPx<MODE> fin = box_scale<MODE> (vacc, d.span_mul_y, P.acc_fits_24 != 0);
        if (store)
        {
            uint32_t packed;
            if constexpr (S128)
            {
                Px<true> o;
                o.w[0] = (uint64_t) fin.v[0] | ((uint64_t) fin.v[1] << 32);
                o.w[1] = (uint64_t) fin.v[2] | ((uint64_t) fin.v[3] << 32);
                packed = pack_px<true> (o, d, P.luts);
            }
            else
            {
                /* fin holds the pixel's four bytes in source order */
                const uint32_t bytes = fin.v[0] | (fin.v[1] << 8);
                const uint32_t alpha = (bytes >> P.alpha_shift) & 0xff;
                const uint32_t cols = bytes >> P.col_shift;
                Px<false> o;
                o.w[0] = (uint64_t) alpha | ((uint64_t) (cols & 0xff) << 16) | ((uint64_t) ((cols >> 8) & 0xff) << 32)
                         | ((uint64_t) ((cols >> 16) & 0xff) << 48);
                packed = pack_px<false> (o, d, P.luts);
            }
            uint8_t *o8 = P.dst + (size_t) img * P.dst_image_stride + (size_t) yl * P.dst_pitch + (size_t) x * d.bpp_out;
            store_raw_px (o8, packed, d.bpp_out);
        }
    }
}

/* Repack of a 128bpp intermediate pixel (four 32-bit lanes: alpha lane, c0, c1, c2) with 32-bit
 * arithmetic.  The reference multiplies whole 64-bit words by the inverse-division entry and
 * shifts (generic:271-318); every field it then keeps lies inside the LOW 32 bits of the lane's
 * product (lanes are masked to 24 bits, table entries are at most 2^19, and the kept bits end at
 * bit 29 at most), so a plain 32-bit multiply gives the same bits.  sm_to_srgb: the 2048-entry
 * table in shared memory.  Returns the pixel's bytes in destination memory order. */
template <int MODE>
__device__ __forceinline__ uint32_t
pack128_fast (const uint32_t lane[4], const SmolJobDesc &d, const SmolDeviceLuts *__restrict__ lut,
              const uint8_t *__restrict__ sm_to_srgb, const uint32_t *__restrict__ sm_inv = nullptr)
{
    /* sm_inv: the mode's inverse-division table in shared memory, if the caller staged one */
    uint32_t a, c[3];

    if constexpr (MODE == BM_P16_U || MODE == BM_P16L_U)
        a = (lane[0] >> 8) & 0xff;                                       /* generic:1140, :1152 */
    else
        a = lane[0] & 0xff;                                              /* generic:1101 */

    if constexpr (MODE == BM_P16_U)
    {
        const uint32_t inv = sm_inv ? sm_inv[a] : __ldg (&lut->inv_div_p16[a]);
#pragma unroll
        for (int i = 0; i < 3; i++)
            c[i] = __byte_perm (lane[i + 1] * inv, 0, 0x4442);           /* (v * inv) >> 16 & 0xff, generic:290-299 */
    }
    else if constexpr (MODE == BM_P16L_U)
    {
        const uint32_t inv = sm_inv ? sm_inv[a] : __ldg (&lut->inv_div_p16l[a]);
#pragma unroll
        for (int i = 0; i < 3; i++)
            c[i] = sm_to_srgb[((lane[i + 1] * inv) >> 19) & 0x7ff];      /* generic:309-318 */
    }
    else
    {
        /* P8L (generic:1096-1134, 24bpp :922-935 / :1010-1023) */
        const uint32_t inv = sm_inv ? sm_inv[a] : __ldg (&lut->inv_div_p8l[a]);
        const bool unpremul = !(d.bpp_out == 3 && d.pack24_direct);
        const bool repremul = d.bpp_out == 4 && !d.out_unassoc;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint32_t v = lane[i + 1];
            if (unpremul)
                v = (v * inv) >> 10;
            v = sm_to_srgb[v & 0x7ff];
            if (repremul)
                v = (((v + 1) * (a + 1) - 1) >> 8) & 0xff;
            c[i] = v;
        }
    }

    if (d.swap_rb)
    {
        const uint32_t t = c[0]; c[0] = c[2]; c[2] = t;
    }
    uint32_t out = (c[0] << (8 * d.out_col0)) | (c[1] << (8 * (d.out_col0 + 1))) | (c[2] << (8 * (d.out_col0 + 2)));
    if (d.out_alpha_idx != 0xff)
        out |= a << (8 * d.out_alpha_idx);
    return out;
}

/* ------------------------------------------------------------------------------------------ *
 * "taps128" kernel: bilinear / copy / one on both axes with a 128bpp intermediate -- linear      *
 * light (P8L) and unassociated -> unassociated (P16 / P16L).  One thread per output pixel, four   *
 * 32-bit lanes per pixel, the unpack chain shared with the box kernel (box_unpack), tables in     *
 * shared memory as 32 lane-private copies so the gathers are conflict-free.                       *
 * ------------------------------------------------------------------------------------------ */

/* Two instances.  A job with enough items to fill every SM's 32 warps is issue-bound: 32 warps of 64
 * registers.  A smaller job is bound by each thread's chain of dependent loads: 16 warps of 128
 * registers, which lets a thread request every source pixel of up to four taps in two rows (16 loads)
 * before it unpacks the first (B200, 256x256 -> 32x32 linear light 23.6 -> 12.5 us with two taps per
 * batch at 80 registers; 4K -> 720p 34.5 -> 30.5 us with the 64-register instance). */
#define SMOL_TAPS128_WARPS(SMALL) ((SMALL) ? 16 : 32)
/* output pixels of a call up to which the tile kernel (smol_tile128h_kernel) takes over: measured 2.5x
 * faster at 1,024 pixels, 10 % slower at 19,200 */
#ifndef SMOL_TILE128H_MAX_PIXELS
#define SMOL_TILE128H_MAX_PIXELS 4096ull
#endif

template <int MODE, int BI, bool SMALL>
__global__ void __launch_bounds__ (SMOL_TAPS128_WARPS (SMALL) * 32, 1)
smol_taps128_kernel (const BoxParams P, uint32_t hh, uint32_t vh, uint32_t src_u32_ok, uint32_t rows_per_item)
{
    extern __shared__ __align__ (16) uint8_t sm_dyn[];
    constexpr bool OPAQUE = BI == 3 && MODE == BM_P8L_P;       /* see box3_accum */
    constexpr bool NEED_INV = MODE == BM_P8L_P && !OPAQUE;
    constexpr bool NEED_FROM = MODE == BM_P8L_P || MODE == BM_P8L_U || MODE == BM_P16L_U;
    const SmolJobDesc &d = P.d;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;

    pdl_launch_dependents ();
    /* byte-addressed lane-replicated tables at fixed window addresses: layout and use as in the
     * box kernel (see box3_accum) */
    uint32_t from_y = 0, inv_y = 0;
    if constexpr (NEED_FROM)
    {
        const uint32_t dyn_addr = (uint32_t) __cvta_generic_to_shared (sm_dyn), dyn_win = dyn_addr & 0x00ffffffu;
        uint32_t *t_from = reinterpret_cast<uint32_t *> (sm_dyn + (SMOL_BOX3_FROM_WIN - dyn_win));
        uint2 *t_inv = reinterpret_cast<uint2 *> (sm_dyn + (SMOL_BOX3_INV_WIN - dyn_win));
        for (uint32_t i = tid; i < 8192; i += nthr)
        {
            const uint32_t e = i >> 5, l = i & 31;
            const uint32_t lin = P.luts->from_srgb[e];
            t_from[e * 64 + l] = OPAQUE ? ((lin + 1) * 2041u - 1) >> 11 : lin + (MODE == BM_P16L_U ? 0u : 1u);
            if constexpr (NEED_INV)
                t_inv[e * 32 + l] = make_uint2 (P.luts->inv_div_p8[e] << 3, e * 8 + 1);
        }
        from_y = (dyn_addr & 0xff000000u) | SMOL_BOX3_FROM_WIN | ((tid & 31) * 4);
        inv_y = (dyn_addr & 0xff000000u) | SMOL_BOX3_INV_WIN | ((tid & 31) * 8);
    }
    __shared__ uint8_t sm_to_srgb[2048];
    if constexpr (MODE != BM_P16_U)
        for (uint32_t i = tid; i < 512; i += nthr)
            reinterpret_cast<uint32_t *> (sm_to_srgb)[i] = reinterpret_cast<const uint32_t *> (P.luts->to_srgb)[i];
    __syncthreads ();

    pdl_wait ();

    /* Persistent CTAs: the 64 KB of lane-replicated tables are filled once per CTA and then serve
     * every work item its warps walk over (item = 32 adjacent output pixels of one row; the host
     * picks the warp count that wastes least of the last round, as for the box kernel). */
    /* an item is rows_per_item consecutive output rows of 32 columns: without halvings the
     * two-row cache then carries half of every output row's work over from the row above */
    const uint32_t items_x = (d.w_out + 31) / 32, strips = (P.n_rows + rows_per_item - 1) / rows_per_item;
    const uint32_t items_per_image = items_x * strips;
    const uint32_t n_items = items_per_image * P.n_images;
    const uint32_t warps = nthr >> 5;
    /* item -> (image, strip, column tile) by multiply-high with a reciprocal computed once per thread
     * (floor (2^32 / d) gives a quotient at most one too small: one correction step) */
    const uint32_t rcp_x = items_x == 1 ? 0xffffffffu : (uint32_t) (0x100000000ull / items_x);
    const uint32_t rcp_img = items_per_image == 1 ? 0xffffffffu : (uint32_t) (0x100000000ull / items_per_image);
    auto div_by = [] (uint32_t t, uint32_t dv, uint32_t rcp) -> uint32_t
    {
        uint32_t q = __umulhi (t, rcp);
        if (t - q * dv >= dv)
            q++;
        return q;
    };
    /* 32bpp source pixels as one word or as four bytes: decided once, outside the item loop (inside, the
     * compiler turned the choice into predicated code and every pixel paid for both forms) */
    auto walk = [&] (auto u32_tag)
    {
    constexpr bool U32 = decltype (u32_tag)::value;
    for (uint32_t item = blockIdx.x * warps + (tid >> 5); item < n_items; item += gridDim.x * warps)
    {
    const uint32_t tz = P.n_images == 1 ? 0u : div_by (item, items_per_image, rcp_img), trem = item - tz * items_per_image;
    const uint32_t strip = div_by (trem, items_x, rcp_x);
    const uint32_t yl0 = strip * rows_per_item, yl1 = min (yl0 + rows_per_item, P.n_rows);
    const uint32_t x = (trem - strip * items_x) * 32 + (tid & 31);
    if (x >= d.w_out)
        continue;

    const uint32_t n_h = 1u << hh, n_v = 1u << vh;
    const uint32_t *tx = P.tab_x + (x << hh);
    const uint8_t *src = P.src + (size_t) tz * P.src_image_stride;

    auto load_raw = [&] (const uint8_t *row, uint32_t j) -> uint32_t
    {
        const uint8_t *p = row + (size_t) j * BI;
        uint32_t raw;
        if constexpr (U32)
            raw = __ldg (reinterpret_cast<const uint32_t *> (p));
        else
        {
            raw = (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16);
            raw |= BI == 4 ? ((uint32_t) __ldg (p + 3) << 24) : 0xff000000u;
        }
        return raw;
    };

    auto unpack = [&] (uint32_t raw) -> BoxPx<MODE>
    {
        if constexpr (NEED_FROM)
        {
            BoxPx<MODE> r;
#pragma unroll
            for (int i = 0; i < 4; i++) r.v[i] = 0;
            box3_accum<MODE, false, OPAQUE> (raw, 0, r.v, P, from_y, inv_y);
            return r;
        }
        else
            return box_unpack<MODE, 0> (raw, P, nullptr, nullptr, nullptr);
    };

    /* Horizontally filtered pixel of NR source rows at once.  A thread walks its output pixel's
     * taps alone, so what bounds a small job is the length of its dependent load chain, not the
     * arithmetic: the table entries of a batch of taps are requested together, then every source
     * pixel the batch needs in all NR rows (HB taps x 2 pixels x NR rows), and only then does the
     * unpack chain start. */
    constexpr int HB = SMALL ? 4 : 2;
    auto hrows = [&] (auto nr_c, const uint32_t *rs, BoxPx<MODE> *out)
    {
        constexpr int NR = decltype (nr_c)::value;
        const uint8_t *row[NR];
#pragma unroll
        for (int j = 0; j < NR; j++)
        {
            row[j] = src + (size_t) rs[j] * P.src_pitch;
#pragma unroll
            for (int i = 0; i < 4; i++) out[j].v[i] = 0;
        }
#pragma unroll 1
        for (uint32_t k0 = 0; k0 < n_h; k0 += HB)
        {
            uint32_t e[HB], rp[NR][HB], rq[NR][HB];
#pragma unroll
            for (int b = 0; b < HB; b++)
                e[b] = __ldg (&tx[min (k0 + b, n_h - 1)]);
#pragma unroll
            for (int b = 0; b < HB; b++)
            {
                const uint32_t ofs = SMOL_TAB_OFS (e[b]), ofs1 = min (ofs + 1, d.w_in - 1);
#pragma unroll
                for (int j = 0; j < NR; j++)
                {
                    rp[j][b] = load_raw (row[j], ofs);
                    rq[j][b] = load_raw (row[j], ofs1);
                }
            }
#pragma unroll
            for (int b = 0; b < HB; b++)
            {
                if (k0 + b < n_h)
                {
                    const uint32_t F = SMOL_TAB_F (e[b]), G = 256u - F;
#pragma unroll
                    for (int j = 0; j < NR; j++)
                    {
                        const BoxPx<MODE> p = unpack (rp[j][b]);
                        const BoxPx<MODE> q = unpack (rq[j][b]);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            out[j].v[i] += ((p.v[i] * F + q.v[i] * G) >> 8) & 0x00ffffffu;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NR; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) out[j].v[i] = (out[j].v[i] >> hh) & 0x00ffffffu;
    };

    uint32_t idx0 = 0xffffffffu, idx1 = 0xffffffffu;
    BoxPx<MODE> c0, c1;
#pragma unroll
    for (int i = 0; i < 4; i++) c0.v[i] = c1.v[i] = 0;

#pragma unroll 1
    for (uint32_t yl = yl0; yl < yl1; yl++)
    {
    const uint32_t *ty = P.tab_y + ((P.first_row + yl) << vh);
    BoxPx<MODE> acc;
#pragma unroll
    for (int i = 0; i < 4; i++) acc.v[i] = 0;
    /* the row's vertical taps (at most four) in one go, off the dependent chain of the loop below */
    uint32_t ey[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        ey[k] = __ldg (&ty[min ((uint32_t) k, n_v - 1)]);

#pragma unroll 1
    for (uint32_t kv = 0; kv < n_v; kv++)
    {
        const uint32_t e = kv == 0 ? ey[0] : kv == 1 ? ey[1] : kv == 2 ? ey[2] : ey[3];
        const uint32_t r0 = SMOL_TAB_OFS (e), F = SMOL_TAB_F (e), G = 256u - F;
        const uint32_t r1 = min (r0 + 1, d.h_in - 1);
        bool new0 = false;

        /* the reference's two-row cache (generic:1648-1682); which rows are new is the same for
         * the whole warp (one output row per item) */
        if (r0 != idx0)
        {
            if (r0 == idx1)
            {
                const BoxPx<MODE> t = c0; c0 = c1; c1 = t;
                idx1 = idx0;
            }
            else
                new0 = true;
            idx0 = r0;
        }
        const bool new1 = r1 != idx1;
        idx1 = r1;
        if (new0 && new1)
        {
            const uint32_t rs[2] = { r0, r1 };
            BoxPx<MODE> o[2];
            hrows (std::integral_constant<int, 2> (), rs, o);
            c0 = o[0]; c1 = o[1];
        }
        else if (new0)
            hrows (std::integral_constant<int, 1> (), &r0, &c0);
        else if (new1)
            hrows (std::integral_constant<int, 1> (), &r1, &c1);
#pragma unroll
        for (int i = 0; i < 4; i++)
            acc.v[i] += ((c0.v[i] * F + c1.v[i] * G) >> 8) & 0x00ffffffu;
    }

    uint32_t fin[4];
#pragma unroll
    for (int i = 0; i < 4; i++) fin[i] = (acc.v[i] >> vh) & 0x00ffffffu;
    const uint32_t packed = pack128_fast<MODE> (fin, d, P.luts, sm_to_srgb);
    uint8_t *o8 = P.dst + (size_t) tz * P.dst_image_stride + (size_t) yl * P.dst_pitch + (size_t) x * d.bpp_out;
    store_raw_px (o8, packed, d.bpp_out);
    }
    }
    };
    if constexpr (BI == 4)
    {
        if (src_u32_ok)
            walk (std::true_type {});
        else
            walk (std::false_type {});
    }
    else
        walk (std::false_type {});
}

/* ------------------------------------------------------------------------------------------ *
 * "taps0w" kernel: bilinear without halvings / copy / one on both axes with a 128bpp             *
 * intermediate -- linear light (P8L) and unassociated -> unassociated (P16 / P16L; reference      *
 * smolscale.c:751-758) -- word-aligned 32bpp rows (24bpp: any).  The 128bpp counterpart of        *
 * smol_taps0_kernel: a thread owns PX adjacent output columns and walks a strip of output rows    *
 * with the last two horizontally filtered source rows ping-ponging between two register sets     *
 * (the reference's two-row cache, generic:1648-1682), four 32-bit lanes per pixel.  The data      *
 * tables are plain shared-memory copies (a strip kernel runs the unpack chain about once per      *
 * output pixel, so a few bank conflicts cost less than 64 KB of lane-replicated tables per CTA).  *
 * Replaces the one-thread-per-pixel taps128 / the tile kernel for these jobs: 4K 1:1              *
 * unassociated 82 -> 29 us, 2x up 42 -> 21 us (B200).                                             *
 * ------------------------------------------------------------------------------------------ */

struct PxW { uint32_t v[4]; };      /* v[0]: alpha lane, v[1..3]: colour lanes in source colour order (box_unpack's layout) */

struct Taps0wParams
{
    TapsParams t;
    SmolJobDesc d;
    const SmolDeviceLuts *luts;
    uint32_t prefetch, row_ahead;
};

/* sm_from: from_srgb + 1 (P8L), from_srgb (P16L), or for opaque 24bpp sources the whole chain's
 * result ((from_srgb + 1) * 2041 - 1) >> 11 (see box3_accum) */
template <int MODE, int BI, bool AF>
__device__ __forceinline__ PxW
taps0w_fetch (const uint8_t *row, uint32_t x, const uint32_t *__restrict__ sm_inv8, const uint16_t *__restrict__ sm_from)
{
    uint32_t raw;
    if constexpr (BI == 4)
        raw = __ldg (reinterpret_cast<const uint32_t *> (row) + x);
    else
    {
        const uint8_t *p = row + (size_t) x * 3;
        raw = (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16) | 0xff000000u;
    }
    constexpr bool af = BI == 4 && AF;
    const uint32_t alpha = af ? raw & 0xffu : raw >> 24;
    uint32_t c[3] = { __byte_perm (raw, 0, af ? 0x4441 : 0x4440), __byte_perm (raw, 0, af ? 0x4442 : 0x4441),
                      __byte_perm (raw, 0, af ? 0x4443 : 0x4442) };
    PxW r;
    if constexpr (MODE == BM_P16_U || MODE == BM_P16L_U)
    {
        r.v[0] = (alpha << 8) | 0x80u;                              /* generic:616-625 */
#pragma unroll
        for (int i = 0; i < 3; i++)
            r.v[i + 1] = (MODE == BM_P16L_U ? (uint32_t) sm_from[c[i]] : c[i]) * alpha;     /* generic:616-660, :708-752 */
    }
    else if constexpr (BI == 3)
    {
        r.v[0] = 255u;                                              /* opaque: one table read per channel */
#pragma unroll
        for (int i = 0; i < 3; i++)
            r.v[i + 1] = sm_from[c[i]];
    }
    else
    {
        const uint32_t m = alpha * 8 + 1;
        if constexpr (MODE == BM_P8L_P)
        {
            const uint32_t inv8 = sm_inv8[alpha];                   /* unpremultiply, generic:227-236 */
#pragma unroll
            for (int i = 0; i < 3; i++)
                c[i] = __byte_perm (c[i] * inv8, 0, 0x4442);
        }
        r.v[0] = alpha;
#pragma unroll
        for (int i = 0; i < 3; i++)
            r.v[i + 1] = ((uint32_t) sm_from[c[i]] * m - 1) >> 11;  /* ((lin + 1) * m - 1) >> 11, generic:261-269 */
    }
    return r;
}

template <int MODE, int BI, int BO, bool AF>
__global__ void __launch_bounds__ (256)
smol_taps0w_kernel (const Taps0wParams T)
{
    constexpr int PX = BO == 3 ? 4 : 2;
    constexpr bool NEED_INV8 = MODE == BM_P8L_P && BI == 4;
    constexpr bool NEED_FROM = MODE != BM_P16_U;
    __shared__ uint32_t sm_inv8[NEED_INV8 ? 256 : 1];
    __shared__ uint32_t sm_invp[256];       /* the repack's inverse-division table */
    __shared__ uint16_t sm_from[NEED_FROM ? 256 : 2];
    __shared__ __align__ (4) uint8_t sm_to_srgb[NEED_FROM ? 2048 : 4];
    const TapsParams &P = T.t;
    const SmolJobDesc &d = T.d;

    pdl_launch_dependents ();
    {
        /* library-owned constant data: readable before the dependency wait */
        const uint32_t t = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
        for (uint32_t i = t; i < 256; i += nt)
        {
            sm_invp[i] = MODE == BM_P16_U ? T.luts->inv_div_p16[i] : MODE == BM_P16L_U ? T.luts->inv_div_p16l[i] : T.luts->inv_div_p8l[i];
            if constexpr (NEED_FROM)
            {
                const uint32_t lin = T.luts->from_srgb[i];
                sm_from[i] = (uint16_t) (MODE == BM_P16L_U ? lin : BI == 3 ? ((lin + 1) * 2041u - 1) >> 11 : lin + 1);
            }
            if constexpr (NEED_INV8)
                sm_inv8[i] = T.luts->inv_div_p8[i] << 3;
        }
        if constexpr (NEED_FROM)
            for (uint32_t i = t; i < 512; i += nt)
                reinterpret_cast<uint32_t *> (sm_to_srgb)[i] = reinterpret_cast<const uint32_t *> (T.luts->to_srgb)[i];
        __syncthreads ();
    }

    const uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    const uint32_t strip = blockIdx.y * blockDim.y + threadIdx.y;
    const uint32_t yl0 = strip * P.rows_per_thread;
    if (x >= P.w_out || yl0 >= P.n_rows)
        return;
    const uint32_t yl1 = min (yl0 + P.rows_per_thread, P.n_rows);
    const uint32_t n_px = min ((uint32_t) PX, P.w_out - x);
    const unsigned live_mask = __activemask ();         /* see smol_taps0_kernel */
    const bool has_next = (threadIdx.x & 31) != 31 && x + PX < P.w_out;
    (void) live_mask; (void) has_next;

    uint32_t op[PX], oq[PX], Fx[PX];
#pragma unroll
    for (int o = 0; o < PX; o++)
    {
        const uint32_t e = __ldg (&P.tab_x[min (x + o, P.w_out - 1)]);
        op[o] = SMOL_TAB_OFS (e);
        oq[o] = min (op[o] + 1, P.w_in - 1);
        Fx[o] = SMOL_TAB_F (e);
    }

    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl0 * P.dst_pitch + (size_t) x * BO;
    /* full-width store of a 32bpp thread's two pixels: 8 bytes */
    const bool fast_store = BO == 4 && n_px == PX && (reinterpret_cast<uintptr_t> (dst) & 7) == 0 && (P.dst_pitch & 7) == 0;

    if (in_first_wave (T.prefetch))
    {
        const uint32_t r = SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl0]));
        const uint8_t *p = src + (size_t) r * P.src_pitch + (size_t) op[0] * BI;
        prefetch_l2 (p);
        prefetch_l2 (p + (size_t) (r + 1 < P.h_in ? P.src_pitch : 0));
    }
    pdl_wait ();

    auto hrow = [&] (uint32_t r, PxW *out)
    {
        const uint8_t *row = src + (size_t) min (r, P.h_in - 1) * P.src_pitch;
        if (T.row_ahead && r + T.row_ahead < P.h_in)
            prefetch_l1 (row + (size_t) T.row_ahead * P.src_pitch + (size_t) op[0] * BI);
#pragma unroll
        for (int o = 0; o < PX; o++)
        {
            const PxW p = taps0w_fetch<MODE, BI, AF> (row, op[o], sm_inv8, sm_from);
            const PxW q = taps0w_fetch<MODE, BI, AF> (row, oq[o], sm_inv8, sm_from);
            const uint32_t F = Fx[o], G = 256u - F;
#pragma unroll
            for (int i = 0; i < 4; i++)
                out[o].v[i] = (p.v[i] * F + q.v[i] * G) >> 8;       /* lanes stay below 2^24: the reference's mask cannot bite */
        }
    };

    auto emit = [&] (const PxW *top, const PxW *bot, uint32_t F)
    {
        const uint32_t G = 256u - F;
        uint32_t out[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int o = 0; o < PX; o++)
        {
            uint32_t fin[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                fin[i] = (top[o].v[i] * F + bot[o].v[i] * G) >> 8;
            out[o] = pack128_fast<MODE> (fin, d, T.luts, sm_to_srgb, sm_invp);
        }
        if constexpr (BO == 4)
        {
            if (fast_store)
                *reinterpret_cast<uint2 *> (dst) = make_uint2 (out[0], out[1]);
            else
            {
                reinterpret_cast<uint32_t *> (dst)[0] = out[0];
                if (n_px > 1)
                    reinterpret_cast<uint32_t *> (dst)[1] = out[1];
            }
        }
        else
            store_px4_rgb_anywhere (dst, out, n_px, live_mask, has_next);
        dst += P.dst_pitch;
    };

    const uint32_t *ty = P.tab_y + P.first_row;
    uint32_t yl = yl0;
    uint32_t e = __ldg (&ty[yl]);
    uint32_t r = SMOL_TAB_OFS (e);
    PxW A[PX], B[PX];
    hrow (r, A);
    hrow (r + 1, B);

    for (;;)
    {
        /* phase A: top row in A, bottom row in B */
        do
        {
            emit (A, B, SMOL_TAB_F (e));
            if (++yl >= yl1)
                return;
            e = __ldg (&ty[yl]);
        }
        while (SMOL_TAB_OFS (e) == r);
        if (SMOL_TAB_OFS (e) != r + 1)
        {
            r = SMOL_TAB_OFS (e);
            hrow (r, A);
            hrow (r + 1, B);
            continue;
        }
        r++;
        hrow (r + 1, A);

        /* phase B: the other way round */
        do
        {
            emit (B, A, SMOL_TAB_F (e));
            if (++yl >= yl1)
                return;
            e = __ldg (&ty[yl]);
        }
        while (SMOL_TAB_OFS (e) == r);
        if (SMOL_TAB_OFS (e) != r + 1)
        {
            r = SMOL_TAB_OFS (e);
            hrow (r, A);
            hrow (r + 1, B);
            continue;
        }
        r++;
        hrow (r + 1, B);
    }
}

/* ------------------------------------------------------------------------------------------ *
 * "tile128" kernel: bilinear without halvings / copy / one on both axes (anything from 1:2 to     *
 * any magnification) with a 128bpp intermediate -- linear light and unassociated -> unassociated. *
 * Same two-phase shared-memory tile as the mag kernel, because here the unpack chain is the       *
 * expensive part and must run exactly once per source pixel:                                      *
 *   1a. load + unpack the tile's source window once per source pixel (four 32-bit lanes);         *
 *   1b. horizontal taps once per (source row, output column);                                     *
 *   2.  vertical taps per output pixel, walking runs of rows that share a source row pair,        *
 *       then the (format-generic) repack.                                                         *
 * ------------------------------------------------------------------------------------------ */

struct Tile128Params
{
    BoxParams b;
    uint32_t tile_w, tile_w_log2, tile_h;
    uint32_t u_pitch, u_cw, u_cw_log2, max_src_rows;
    uint32_t src_u32_ok;
};

/* BO: destination bytes per pixel; 32bpp destinations need 4-byte-aligned rows. */
template <int MODE, int BI, int BO>
__global__ void __launch_bounds__ (512)
smol_tile128_kernel (const Tile128Params M)
{
    extern __shared__ __align__ (16) uint8_t sm_dyn[];
    __shared__ uint32_t sm_ty[64];
    __shared__ uint32_t sm_inv8[256];
    __shared__ uint32_t sm_from[256];
    __shared__ uint8_t sm_to_srgb[2048];
    constexpr bool NEED_INV = MODE == BM_P8L_P;
    constexpr bool NEED_FROM = MODE == BM_P8L_P || MODE == BM_P8L_U || MODE == BM_P16L_U;
    const BoxParams &P = M.b;
    const SmolJobDesc &d = P.d;
    const uint32_t tid = threadIdx.x;

    pdl_launch_dependents ();
    /* one plain copy of each table per CTA: the gathers happen once per SOURCE pixel here, so
     * their bank conflicts are cheaper than replicating 64 KB of tables into every tile's CTA */
    if (tid < 256)
    {
        if constexpr (NEED_FROM)
            sm_from[tid] = P.luts->from_srgb[tid];
        if constexpr (NEED_INV)
            sm_inv8[tid] = P.luts->inv_div_p8[tid] << 3;
    }
    if constexpr (MODE != BM_P16_U)
        reinterpret_cast<uint32_t *> (sm_to_srgb)[tid] = reinterpret_cast<const uint32_t *> (P.luts->to_srgb)[tid];

    const uint32_t x0 = blockIdx.x * M.tile_w;
    const uint32_t x1 = min (x0 + M.tile_w, d.w_out);
    const uint32_t yl0 = blockIdx.y * M.tile_h;
    const uint32_t yl1 = min (yl0 + M.tile_h, P.n_rows);
    const uint32_t tw = x1 - x0, th = yl1 - yl0;

    const uint32_t c_lo = SMOL_TAB_OFS (__ldg (&P.tab_x[x0]));
    const uint32_t c_hi = min (SMOL_TAB_OFS (__ldg (&P.tab_x[x1 - 1])) + 1, d.w_in - 1);
    const uint32_t r_lo = SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl0]));
    const uint32_t r_hi = min (SMOL_TAB_OFS (__ldg (&P.tab_y[P.first_row + yl1 - 1])) + 1, d.h_in - 1);
    const uint32_t n_cols = c_hi - c_lo + 1, n_rows = r_hi - r_lo + 1;
    if (tid < th)
        sm_ty[tid] = __ldg (&P.tab_y[P.first_row + yl0 + tid]);

    uint4 *sm_u = reinterpret_cast<uint4 *> (sm_dyn);
    uint4 *sm_h = sm_u + (size_t) M.max_src_rows * M.u_pitch;

    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    __syncthreads ();
    pdl_wait ();

    /* phase 1a */
    {
        const uint32_t ct = tid & (M.u_cw - 1), rt = tid >> M.u_cw_log2, r_step = 512u >> M.u_cw_log2;
        for (uint32_t r = rt; r < n_rows; r += r_step)
        {
            const uint8_t *row = src + (size_t) (r_lo + r) * P.src_pitch + (size_t) c_lo * BI;
            for (uint32_t c = ct; c < n_cols; c += M.u_cw)
            {
                const uint8_t *p = row + c * BI;
                uint32_t raw;
                if (BI == 4 && M.src_u32_ok)
                    raw = __ldg (reinterpret_cast<const uint32_t *> (p));
                else
                {
                    raw = (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16);
                    raw |= BI == 4 ? ((uint32_t) __ldg (p + 3) << 24) : 0xff000000u;
                }
                const BoxPx<MODE> u = box_unpack<MODE, 0> (raw, P, sm_inv8, sm_from, nullptr);
                sm_u[r * M.u_pitch + c] = make_uint4 (u.v[0], u.v[1], u.v[2], u.v[3]);
            }
        }
    }
    __syncthreads ();

    /* phase 1b */
    {
        const bool full = tw == M.tile_w;
        const uint32_t xl = full ? (tid & (M.tile_w - 1)) : tid % tw;
        const uint32_t r_step = full ? (512u >> M.tile_w_log2) : 512 / tw;
        const uint32_t r_first = full ? (tid >> M.tile_w_log2) : tid / tw;
        const uint32_t e = __ldg (&P.tab_x[x0 + xl]);
        const uint32_t op = SMOL_TAB_OFS (e) - c_lo, oq = min (SMOL_TAB_OFS (e) + 1, d.w_in - 1) - c_lo;
        const uint32_t F = SMOL_TAB_F (e), G = 256u - F;
        if (r_step > 0)
            for (uint32_t r = r_first; r < n_rows; r += r_step)
            {
                const uint4 p = sm_u[r * M.u_pitch + op], q = sm_u[r * M.u_pitch + oq];
                uint4 h;
                h.x = ((p.x * F + q.x * G) >> 8) & 0x00ffffffu;
                h.y = ((p.y * F + q.y * G) >> 8) & 0x00ffffffu;
                h.z = ((p.z * F + q.z * G) >> 8) & 0x00ffffffu;
                h.w = ((p.w * F + q.w * G) >> 8) & 0x00ffffffu;
                sm_h[r * M.tile_w + xl] = h;
            }
    }
    __syncthreads ();

    /* phase 2 */
    const bool full = tw == M.tile_w;
    const uint32_t n_runs = full ? (512u >> M.tile_w_log2) : 512 / tw;
    const uint32_t xl = full ? (tid & (M.tile_w - 1)) : tid % tw;
    const uint32_t run = full ? (tid >> M.tile_w_log2) : tid / tw;
    if (run >= n_runs)
        return;
    const uint32_t rows_per_run = (th + n_runs - 1) / n_runs;
    const uint32_t ry_begin = run * rows_per_run, ry_end = min (ry_begin + rows_per_run, th);
    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) (yl0 + ry_begin) * P.dst_pitch
                   + (size_t) (x0 + xl) * BO;
    const uint4 *hcol = sm_h + xl;

    uint32_t ry = ry_begin;
    uint32_t e = ry < ry_end ? sm_ty[ry] : 0;
    while (ry < ry_end)
    {
        const uint32_t ofs = SMOL_TAB_OFS (e);
        const uint32_t r0 = ofs - r_lo, r1 = min (ofs + 1, d.h_in - 1) - r_lo;
        const uint4 t = hcol[r0 * M.tile_w], b = hcol[r1 * M.tile_w];

        do
        {
            const uint32_t F = SMOL_TAB_F (e), G = 256u - F;
            const uint32_t fin[4] = { ((t.x * F + b.x * G) >> 8) & 0x00ffffffu, ((t.y * F + b.y * G) >> 8) & 0x00ffffffu,
                                      ((t.z * F + b.z * G) >> 8) & 0x00ffffffu, ((t.w * F + b.w * G) >> 8) & 0x00ffffffu };
            const uint32_t v = pack128_fast<MODE> (fin, d, P.luts, sm_to_srgb);
            if constexpr (BO == 4)
                *reinterpret_cast<uint32_t *> (dst) = v;
            else
            {
                dst[0] = (uint8_t) v; dst[1] = (uint8_t) (v >> 8); dst[2] = (uint8_t) (v >> 16);
            }
            ry++;
            dst += P.dst_pitch;
            e = sm_ty[min (ry, th - 1)];
        }
        while (ry < ry_end && SMOL_TAB_OFS (e) == ofs);
    }
}

/* ------------------------------------------------------------------------------------------ *
 * "tile128h" kernel: bilinear WITH halvings on either axis (2:1 < ratio <= 8:1) on a 128bpp       *
 * intermediate, for TINY jobs (icons: 256x256 -> 32x32 in linear light) -- the tile kernel's       *
 * three phases with 2^h taps per output pixel.  taps128 gives every output pixel to one thread,     *
 * which walks all its taps alone: with a few hundred output pixels that is a few warps, each        *
 * working through 64 source pixels one dependent chain after the other (11 us).  Here a 512-thread  *
 * CTA owns a tile of output pixels and spreads the tile's whole source window over its threads:    *
 *   1a. the window is loaded (coalesced, eight pixels in flight per thread) and unpacked ONCE      *
 *       per source pixel;                                                                          *
 *   1b. per (source row, output column): the 2^hh horizontal taps, summed and halved;              *
 *   2.  per output pixel: the 2^vh vertical taps between filtered rows, halved, repacked.          *
 * Same lane arithmetic as taps128 (bit-identical results; 60-job sweep with SMOL_TILE128H=1).      *
 * 256x256 -> 32x32 linear light: 11.0 -> 4.4 us.  Measured and REJECTED for larger jobs: a tile's   *
 * life is a chain of dependent latencies (tables, window bounds, window, three barriers) that two   *
 * or three resident CTAs per SM do not hide -- 4K -> 720p 62 us against taps128's 29, and neither   *
 * more loads in flight nor larger tiles changed that (profiles/r02_tile128h_sweep.json).           *
 * Reported as kernel family "taps128".                                                             *
 * ------------------------------------------------------------------------------------------ */

struct Tile128hParams
{
    BoxParams b;
    uint32_t hh, vh;
    uint32_t tile_w, tile_h;            /* output tile */
    uint32_t u_pitch, max_src_rows;     /* bound of the tile's source window (pixels, rows) */
    uint32_t src_u32_ok;
};

template <int MODE, int BI, int BO>
__global__ void __launch_bounds__ (512)
smol_tile128h_kernel (const Tile128hParams M)
{
    extern __shared__ __align__ (16) uint8_t sm_dyn[];
    __shared__ uint32_t sm_ty[128];
    __shared__ uint32_t sm_inv8[256];
    __shared__ uint32_t sm_from[256];
    __shared__ uint8_t sm_to_srgb[2048];
    constexpr bool NEED_INV = MODE == BM_P8L_P;
    constexpr bool NEED_FROM = MODE == BM_P8L_P || MODE == BM_P8L_U || MODE == BM_P16L_U;
    const BoxParams &P = M.b;
    const SmolJobDesc &d = P.d;
    const uint32_t tid = threadIdx.x;
    const uint32_t hh = M.hh, vh = M.vh, n_h = 1u << hh, n_v = 1u << vh;

    pdl_launch_dependents ();
    if (tid < 256)
    {
        if constexpr (NEED_FROM)
            sm_from[tid] = P.luts->from_srgb[tid];
        if constexpr (NEED_INV)
            sm_inv8[tid] = P.luts->inv_div_p8[tid] << 3;
    }
    if constexpr (MODE != BM_P16_U)
        reinterpret_cast<uint32_t *> (sm_to_srgb)[tid] = reinterpret_cast<const uint32_t *> (P.luts->to_srgb)[tid];

    const uint32_t x0 = blockIdx.x * M.tile_w;
    const uint32_t x1 = min (x0 + M.tile_w, d.w_out);
    const uint32_t yl0 = blockIdx.y * M.tile_h;
    const uint32_t yl1 = min (yl0 + M.tile_h, P.n_rows);
    const uint32_t tw = x1 - x0, th = yl1 - yl0;
    const uint32_t *ty = P.tab_y + ((P.first_row + yl0) << vh);

    /* table offsets never decrease along an axis: the first and the last tap bound the window */
    const uint32_t c_lo = SMOL_TAB_OFS (__ldg (&P.tab_x[x0 << hh]));
    const uint32_t c_hi = min (SMOL_TAB_OFS (__ldg (&P.tab_x[(x1 << hh) - 1])) + 1, d.w_in - 1);
    const uint32_t r_lo = SMOL_TAB_OFS (__ldg (&ty[0]));
    const uint32_t r_hi = min (SMOL_TAB_OFS (__ldg (&ty[(th << vh) - 1])) + 1, d.h_in - 1);
    const uint32_t n_cols = c_hi - c_lo + 1, n_rows = r_hi - r_lo + 1;
    if (n_cols > M.u_pitch || n_rows > M.max_src_rows)
        __trap ();                      /* the host's window bound is wrong: never corrupt memory quietly */
    for (uint32_t i = tid; i < (th << vh); i += 512)
        sm_ty[i] = __ldg (&ty[i]);

    uint4 *sm_u = reinterpret_cast<uint4 *> (sm_dyn);
    uint4 *sm_h = sm_u + (size_t) M.max_src_rows * M.u_pitch;

    const uint8_t *src = P.src + (size_t) blockIdx.z * P.src_image_stride;
    __syncthreads ();
    pdl_wait ();

    /* phase 1a: (row, column) pairs of the window flattened over the CTA's threads */
    {
        const uint32_t total = n_rows * n_cols;
        const uint32_t rcp = n_cols == 1 ? 0xffffffffu : (uint32_t) (0x100000000ull / n_cols);
        const uint8_t *win = src + (size_t) r_lo * P.src_pitch + (size_t) c_lo * BI;
        auto fill = [&] (auto u32_tag)
        {
            constexpr bool U32 = decltype (u32_tag)::value;
            /* NB source pixels per thread requested before the first is unpacked: with one or two loads in
             * flight per thread the phase ran at the latency of a load, not at the rate of the memory system */
            constexpr int NB = 8;
            for (uint32_t i0 = tid; i0 < total; i0 += 512 * NB)
            {
                uint32_t raw[NB], at[NB];
#pragma unroll
                for (int b = 0; b < NB; b++)
                {
                    const uint32_t i = min (i0 + (uint32_t) b * 512, total - 1);
                    uint32_t r = __umulhi (i, rcp);
                    if (i - r * n_cols >= n_cols)
                        r++;
                    const uint32_t c = i - r * n_cols;
                    const uint8_t *p = win + (size_t) r * P.src_pitch + c * BI;
                    at[b] = r * M.u_pitch + c;
                    if constexpr (U32)
                        raw[b] = __ldg (reinterpret_cast<const uint32_t *> (p));
                    else
                    {
                        raw[b] = (uint32_t) __ldg (p) | ((uint32_t) __ldg (p + 1) << 8) | ((uint32_t) __ldg (p + 2) << 16);
                        raw[b] |= BI == 4 ? ((uint32_t) __ldg (p + 3) << 24) : 0xff000000u;
                    }
                }
#pragma unroll
                for (int b = 0; b < NB; b++)
                {
                    if (i0 + (uint32_t) b * 512 < total)
                    {
                        const BoxPx<MODE> u = box_unpack<MODE, 0> (raw[b], P, sm_inv8, sm_from, nullptr);
                        sm_u[at[b]] = make_uint4 (u.v[0], u.v[1], u.v[2], u.v[3]);
                    }
                }
            }
        };
        if constexpr (BI == 4)
        {
            if (M.src_u32_ok)
                fill (std::true_type {});
            else
                fill (std::false_type {});
        }
        else
            fill (std::false_type {});
    }
    __syncthreads ();

    /* tile_w is a power of two: full tiles split the thread index with shifts */
    const bool full = tw == M.tile_w;
    const uint32_t n_runs = 512 / tw;
    const uint32_t run = full ? tid / M.tile_w : tid / tw;
    const uint32_t xl = tid - run * tw;

    /* phase 1b */
    if (run < n_runs)
    {
        uint32_t op[4], oq[4], F[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const uint32_t e = __ldg (&P.tab_x[((x0 + xl) << hh) + min ((uint32_t) k, n_h - 1)]);
            op[k] = SMOL_TAB_OFS (e) - c_lo;
            oq[k] = min (SMOL_TAB_OFS (e) + 1, d.w_in - 1) - c_lo;
            F[k] = SMOL_TAB_F (e);
        }
        for (uint32_t r = run; r < n_rows; r += n_runs)
        {
            const uint4 *urow = sm_u + r * M.u_pitch;
            uint4 h = make_uint4 (0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                if ((uint32_t) k < n_h)
                {
                    const uint4 p = urow[op[k]], q = urow[oq[k]];
                    const uint32_t G = 256u - F[k];
                    h.x += ((p.x * F[k] + q.x * G) >> 8) & 0x00ffffffu;
                    h.y += ((p.y * F[k] + q.y * G) >> 8) & 0x00ffffffu;
                    h.z += ((p.z * F[k] + q.z * G) >> 8) & 0x00ffffffu;
                    h.w += ((p.w * F[k] + q.w * G) >> 8) & 0x00ffffffu;
                }
            }
            h.x = (h.x >> hh) & 0x00ffffffu; h.y = (h.y >> hh) & 0x00ffffffu;
            h.z = (h.z >> hh) & 0x00ffffffu; h.w = (h.w >> hh) & 0x00ffffffu;
            sm_h[r * M.tile_w + xl] = h;
        }
    }
    __syncthreads ();

    /* phase 2 */
    if (run >= n_runs)
        return;
    const uint4 *hcol = sm_h + xl;
    uint8_t *dst = P.dst + (size_t) blockIdx.z * P.dst_image_stride + (size_t) yl0 * P.dst_pitch + (size_t) (x0 + xl) * BO;
    for (uint32_t ry = run; ry < th; ry += n_runs)
    {
        uint32_t acc[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int kv = 0; kv < 4; kv++)
        {
            if ((uint32_t) kv < n_v)
            {
                const uint32_t e = sm_ty[(ry << vh) + kv];
                const uint32_t ofs = SMOL_TAB_OFS (e), Fv = SMOL_TAB_F (e), Gv = 256u - Fv;
                const uint4 t = hcol[(ofs - r_lo) * M.tile_w], b = hcol[(min (ofs + 1, d.h_in - 1) - r_lo) * M.tile_w];
                acc[0] += ((t.x * Fv + b.x * Gv) >> 8) & 0x00ffffffu;
                acc[1] += ((t.y * Fv + b.y * Gv) >> 8) & 0x00ffffffu;
                acc[2] += ((t.z * Fv + b.z * Gv) >> 8) & 0x00ffffffu;
                acc[3] += ((t.w * Fv + b.w * Gv) >> 8) & 0x00ffffffu;
            }
        }
        uint32_t fin[4];
#pragma unroll
        for (int i = 0; i < 4; i++) fin[i] = (acc[i] >> vh) & 0x00ffffffu;
        const uint32_t v = pack128_fast<MODE> (fin, d, P.luts, sm_to_srgb);
        uint8_t *o = dst + (size_t) ry * P.dst_pitch;
        if constexpr (BO == 4)
            *reinterpret_cast<uint32_t *> (o) = v;
        else
        {
            o[0] = (uint8_t) v; o[1] = (uint8_t) (v >> 8); o[2] = (uint8_t) (v >> 16);
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * "rows" kernel: ANY filter pair, any format, any alignment -- the fast backstop.  It takes what  *
 * the specialised kernels leave: box on one axis and bilinear on the other (8000 x 1000 -> 800 x   *
 * 500), ratios beyond 255:1 without linear light, box jobs on byte-misaligned 32bpp rows.          *
 *                                                                                              *
 * The frame is the box kernel's: the unit of work is one WARP producing 32 / G adjacent output     *
 * columns of a strip of consecutive output rows (no block barriers, load balance at warp           *
 * granularity); the strip's source rows are walked ONCE, in order, each staged into the warp's     *
 * own double-buffered shared-memory window with 16-byte cp.async on chunks aligned in global       *
 * memory (straddling chunks byte-exactly), so the copy of row r + 1 overlaps the arithmetic on     *
 * row r.  Per row a lane computes its column's horizontally filtered pixel -- a box span walk or   *
 * 2^h bilinear taps -- and feeds it to the vertical stage: box accumulation with the boundary      *
 * rows shared between neighbouring output rows, or a two-row cache from which every vertical       *
 * sample whose lower row has just arrived is taken.  The arithmetic is the general kernel's        *
 * (64-bit words of four 16-bit or two 32-bit lanes, runtime formats through unpack_px / pack_px),  *
 * with the six data tables copied to shared memory once per CTA.                                   *
 * ------------------------------------------------------------------------------------------ */

struct RowsParams
{
    SmolLaunch L;
    BoxParams b;                    /* the lane-arithmetic variants' selectors (box_params_init) */
    uint32_t lanes_per_col_log2;    /* box spans: lanes sharing a column */
    uint32_t x_tiles;               /* items per strip of rows */
    uint32_t rows_per_item, n_strips;
    uint32_t seg_bytes;             /* bytes per staging buffer (multiple of 16) */
};

/* Pixel arithmetic of the rows kernel, two interchangeable sets:
 *   RowsSwar<S128>      the general kernel's: 64-bit words of 16- or 32-bit lanes, formats decoded at
 *                       run time (unpack_px / pack_px) -- every format, every storage, any alignment;
 *   RowsLanes<MODE, BI> the box kernel's: 32-bit registers, the intermediate encoding and the source
 *                       pixel size fixed at compile time (box_unpack) -- word-aligned 32bpp rows or
 *                       24bpp rows; about four times fewer instructions per source pixel. */
struct RowsTables
{
    const SmolDeviceLuts *luts;     /* shared-memory copy */
    const uint32_t *inv8, *from;    /* inv_div_p8 << 3 and from_srgb as 32-bit words, shared memory */
};

template <bool S128>
struct RowsSwar
{
    typedef Px<S128> P;
    static constexpr bool WIDE = S128;
    static __device__ __forceinline__ P zero () { return px_zero<S128> (); }
    static __device__ __forceinline__ void add (P &a, const P &b) { px_add<S128> (a, b); }
    static __device__ __forceinline__ P weight (const P &p, uint32_t w) { return px_weight<S128> (p, w); }
    static __device__ __forceinline__ P lerp (const P &p, const P &q, uint32_t F) { return px_lerp<S128> (p, q, F); }
    static __device__ __forceinline__ P halve (const P &p, uint32_t n) { return px_halve<S128> (p, n); }
    static __device__ __forceinline__ P scale (const P &p, uint32_t mul, const RowsParams &) { return px_box_scale<S128> (p, mul); }
    static __device__ __forceinline__ P shfl_add (const P &p, uint32_t m) { return px_shfl_xor_add<S128> (p, m); }
    static __device__ __forceinline__ P fetch (const uint8_t *sm, uint32_t j, const RowsParams &R, const RowsTables &T)
    {
        return unpack_px<S128> (load_raw_px (sm + (size_t) j * R.L.d.bpp_in, R.L.d.bpp_in), R.L.d, T.luts);
    }
    static __device__ __forceinline__ uint32_t pack (const P &p, const RowsParams &R, const RowsTables &T)
    {
        return pack_px<S128> (p, R.L.d, T.luts);
    }
};

template <int MODE, int BI>
struct RowsLanes
{
    typedef BoxPx<MODE> P;
    static constexpr bool WIDE = MODE >= BM_P8L_P;
    static constexpr int N = WIDE ? 4 : 2;
    static constexpr uint32_t MASK = WIDE ? 0x00ffffffu : 0x00ff00ffu;
    static __device__ __forceinline__ P zero ()
    {
        P r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    static __device__ __forceinline__ void add (P &a, const P &b) { box_add<MODE> (a, b); }
    static __device__ __forceinline__ P weight (const P &p, uint32_t w) { return box_weight<MODE> (p, w); }
    static __device__ __forceinline__ P lerp (const P &p, const P &q, uint32_t F)
    {
        P r;
        const uint32_t G = 256u - F;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = ((p.v[i] * F + q.v[i] * G) >> 8) & MASK;
        return r;
    }
    static __device__ __forceinline__ P halve (const P &p, uint32_t n)
    {
        P r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = (p.v[i] >> n) & MASK;
        return r;
    }
    /* (always the wide form: with bilinear on one axis the lanes a box sums are not the 8- or 16-bit
     * values box_params_init's "fits 24 bits" estimate assumes -- found by the soak: P16L lanes of
     * 19 bits summed over a 33-row box) */
    static __device__ __forceinline__ P scale (const P &p, uint32_t mul, const RowsParams &) { return box_scale<MODE> (p, mul, false); }
    static __device__ __forceinline__ P shfl_add (P p, uint32_t m)
    {
#pragma unroll
        for (int i = 0; i < N; i++) p.v[i] += __shfl_xor_sync (0xffffffffu, p.v[i], m);
        return p;
    }
    static __device__ __forceinline__ P fetch (const uint8_t *sm, uint32_t j, const RowsParams &R, const RowsTables &T)
    {
        uint32_t raw;
        if constexpr (BI == 4)
            raw = *reinterpret_cast<const uint32_t *> (sm + (size_t) j * 4);       /* rows are word aligned, so is the staging offset */
        else
        {
            const uintptr_t a = reinterpret_cast<uintptr_t> (sm + (size_t) j * 3);
            const uint32_t *w = reinterpret_cast<const uint32_t *> (a & ~(uintptr_t) 3);
            raw = __funnelshift_r (w[0], w[1], (uint32_t) (a & 3) * 8) | 0xff000000u;
        }
        return box_unpack<MODE, 0> (raw, R.b, T.inv8, T.from, nullptr);
    }
    static __device__ __forceinline__ uint32_t pack (const P &p, const RowsParams &R, const RowsTables &T)
    {
        if constexpr (WIDE)
        {
            Px<true> o;
            o.w[0] = (uint64_t) p.v[0] | ((uint64_t) p.v[1] << 32);
            o.w[1] = (uint64_t) p.v[2] | ((uint64_t) p.v[3] << 32);
            return pack_px<true> (o, R.L.d, T.luts);
        }
        else
        {
            /* the pixel's four bytes in source order (see the box kernel) */
            const uint32_t bytes = p.v[0] | (p.v[1] << 8);
            const uint32_t alpha = (bytes >> R.b.alpha_shift) & 0xff;
            const uint32_t cols = bytes >> R.b.col_shift;
            Px<false> o;
            o.w[0] = (uint64_t) alpha | ((uint64_t) (cols & 0xff) << 16) | ((uint64_t) ((cols >> 8) & 0xff) << 32)
                     | ((uint64_t) ((cols >> 16) & 0xff) << 48);
            return pack_px<false> (o, R.L.d, T.luts);
        }
    }
};

template <class OPS, bool HBOX, bool VBOX>
__global__ void __launch_bounds__ (256)
smol_rows_kernel (const RowsParams P)
{
    typedef typename OPS::P PxT;
    constexpr bool S128 = OPS::WIDE;
    extern __shared__ __align__ (16) uint8_t sm_dyn[];
    __shared__ SmolDeviceLuts sm_luts;
    __shared__ uint32_t sm_inv8[256], sm_from[256];
    const SmolLaunch &L = P.L;
    const SmolJobDesc &d = L.d;
    const uint32_t bpp = d.bpp_in;

    pdl_launch_dependents ();
    {
        /* library-owned constant data: readable before the dependency wait */
        const uint32_t *g = reinterpret_cast<const uint32_t *> (L.luts);
        uint32_t *s = reinterpret_cast<uint32_t *> (&sm_luts);
        for (uint32_t i = threadIdx.x; i < sizeof (SmolDeviceLuts) / 4; i += blockDim.x)
            s[i] = __ldg (g + i);
        sm_inv8[threadIdx.x] = __ldg (&L.luts->inv_div_p8[threadIdx.x]) << 3;
        sm_from[threadIdx.x] = L.luts->from_srgb[threadIdx.x];
    }
    __syncthreads ();
    pdl_wait ();
    RowsTables tabs;
    tabs.luts = &sm_luts;
    tabs.inv8 = sm_inv8;
    tabs.from = sm_from;

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t G = 1u << P.lanes_per_col_log2, g = lane & (G - 1);
    const uint32_t cols_per_item = 32u >> P.lanes_per_col_log2;
    uint8_t *bufs = sm_dyn + (size_t) warp * 2 * P.seg_bytes;
    const uint32_t bufs_addr = (uint32_t) __cvta_generic_to_shared (bufs);
    const uint32_t row_bytes = d.w_in * bpp;
    const uint32_t hh = d.h_halvings, vh = d.v_halvings;

    const uint32_t items_per_image = P.x_tiles * P.n_strips;
    const uint32_t n_items = items_per_image * L.n_images;
    const uint32_t warp_stride = gridDim.x * (blockDim.x >> 5);

    for (uint32_t item = blockIdx.x * (blockDim.x >> 5) + warp; item < n_items; item += warp_stride)
    {
        const uint32_t img = item / items_per_image;
        const uint32_t rem = item - img * items_per_image;
        const uint32_t strip = rem / P.x_tiles, xt = rem - strip * P.x_tiles;
        const uint32_t yl0 = strip * P.rows_per_item, yl1 = min (yl0 + P.rows_per_item, L.n_rows);
        const uint32_t x_first = xt * cols_per_item;
        const uint32_t x_last = min (x_first + cols_per_item, d.w_out) - 1;
        uint32_t x = x_first + (lane >> P.lanes_per_col_log2);
        const bool store = x <= x_last && g == 0;
        x = min (x, x_last);

        /* horizontal plan of this lane's column, and the window of source pixels the item reads */
        uint32_t hL = 0, hR = 0, wl = 0, wr = 0, sx0, sx1;
        if constexpr (HBOX)
        {
            const uint32_t e0 = __ldg (&L.tab_x[x]), e1 = __ldg (&L.tab_x[x + 1]);
            hL = SMOL_TAB_OFS (e0); hR = SMOL_TAB_OFS (e1); wr = SMOL_TAB_F (e0);
            wl = x == 0 ? 256u : 255u - SMOL_TAB_F (__ldg (&L.tab_x[x - 1]));
            sx0 = SMOL_TAB_OFS (__ldg (&L.tab_x[x_first]));
            sx1 = SMOL_TAB_OFS (__ldg (&L.tab_x[x_last + 1]));
        }
        else
        {
            sx0 = SMOL_TAB_OFS (__ldg (&L.tab_x[x_first << hh]));
            sx1 = min (SMOL_TAB_OFS (__ldg (&L.tab_x[((x_last + 1) << hh) - 1])) + 1, d.w_in - 1);
        }
        const uint32_t sx_b = sx0 * bpp, end_b = (sx1 + 1) * bpp;

        /* vertical plan: the strip's source rows r_first .. r_last are consecutive */
        const uint32_t y0 = L.first_row + yl0;
        uint32_t r_first, r_last;
        uint32_t B_cur = 0, rend_cur = 0, w1_cur = 0, w2_cur = 0, Fy_cur = 0;     /* box */
        uint32_t k_cur = 0, k_stop = 0;                                            /* taps: sample indices into tab_y */
        if constexpr (VBOX)
        {
            const uint32_t e0 = __ldg (&L.tab_y[y0]), e1 = __ldg (&L.tab_y[y0 + 1]);
            r_first = SMOL_TAB_OFS (e0);
            B_cur = SMOL_TAB_OFS (e1);
            Fy_cur = SMOL_TAB_F (e0);
            w1_cur = y0 == 0 ? 256u : 255u - SMOL_TAB_F (__ldg (&L.tab_y[y0 - 1]));
            w2_cur = S128 ? Fy_cur - 1 : Fy_cur;        /* 128bpp weighs the trailing row by F - 1 (generic:2247-2249) */
            rend_cur = Fy_cur > 0 ? B_cur : B_cur - 1;
            const uint32_t ya = L.first_row + yl1 - 1;
            const uint32_t ea = __ldg (&L.tab_y[ya]), eb = __ldg (&L.tab_y[ya + 1]);
            r_last = SMOL_TAB_F (ea) > 0 ? SMOL_TAB_OFS (eb) : SMOL_TAB_OFS (eb) - 1;
        }
        else
        {
            k_cur = y0 << vh;
            k_stop = (L.first_row + yl1) << vh;
            r_first = SMOL_TAB_OFS (__ldg (&L.tab_y[k_cur]));
            r_last = min (SMOL_TAB_OFS (__ldg (&L.tab_y[k_stop - 1])) + 1, d.h_in - 1);
        }

        const uint8_t *img_base = L.src + (size_t) img * L.src_image_stride;
        uint8_t *dst_img = L.dst + (size_t) img * L.dst_image_stride;

        /* row-relative offset (may be negative) of the first staged byte of source row r */
        auto staged_from = [&] (uint32_t r) -> int32_t
        {
            const uint32_t A = (uint32_t) reinterpret_cast<uintptr_t> (img_base + (size_t) r * L.src_pitch) & 15u;
            return (int32_t) ((A + sx_b) & ~15u) - (int32_t) A;
        };
        auto stage = [&] (uint32_t slot_ofs, uint32_t r)
        {
            const uint8_t *rowp = img_base + (size_t) r * L.src_pitch;
            const int32_t w0 = staged_from (r);
            const uint32_t nch = ((uint32_t) ((int32_t) end_b - w0) + 15u) >> 4;
            for (uint32_t k = lane; k < nch; k += 32)
            {
                const int32_t ofs = w0 + 16 * (int32_t) k;
                const uint32_t sa = bufs_addr + slot_ofs + 16 * k;
                if (ofs >= 0 && ofs + 16 <= (int32_t) row_bytes)
                    cp_async_16_full (sa, rowp + ofs);
                else
                {
                    const int32_t lo = ofs < 0 ? -ofs : 0, hi = min (16, (int32_t) row_bytes - ofs);
                    for (int32_t b = lo; b < hi; b++)
                    {
                        const uint32_t v = __ldg (rowp + ofs + b);
                        asm volatile ("st.shared.u8 [%0], %1;" :: "r"(sa + b), "r"(v) : "memory");
                    }
                }
            }
            cp_async_commit ();
        };

        PxT vacc = OPS::zero ();       /* box: the output row's accumulator; taps: the sum of its samples */
        PxT h_prev = OPS::zero ();     /* taps: the row above */
        bool top_pending = true;                /* box: the current output row's first source row is still to come */
        uint32_t yl_cur = yl0;
        uint32_t cur = 0;

        auto emit_row = [&] (const PxT &out)
        {
            if (store)
            {
                uint8_t *o = dst_img + (size_t) yl_cur * L.dst_pitch + (size_t) x * d.bpp_out;
                store_raw_px (o, OPS::pack (out, P, tabs), d.bpp_out);
            }
            yl_cur++;
        };

        stage (0, r_first);
        for (uint32_t r = r_first; r <= r_last; r++)
        {
            if (r < r_last)
            {
                stage (P.seg_bytes - cur, r + 1);
                cp_async_wait<1> ();
            }
            else
                cp_async_wait<0> ();
            __syncwarp ();

            /* pixel j of the row starts at sm[j * bpp - w0] */
            const uint8_t *sm = bufs + cur - staged_from (r);
            auto px_at = [&] (uint32_t j) -> PxT
            {
                return OPS::fetch (sm, j, P, tabs);
            };

            PxT h;
            if constexpr (HBOX)
            {
                /* generic:1427-1556 in absolute offsets */
                PxT acc = OPS::zero ();
                for (uint32_t j = hL + 1 + g; j < hR; j += G)
                    OPS::add (acc, px_at (j));
                if (g == 0)
                    OPS::add (acc, OPS::weight (px_at (hL), wl));
                if (g == G - 1 && wr > 0)
                    OPS::add (acc, OPS::weight (px_at (hR), wr));
                for (uint32_t m = G >> 1; m; m >>= 1)
                    acc = OPS::shfl_add (acc, m);
                h = OPS::scale (acc, d.span_mul_x, P);
            }
            else
            {
                /* generic:1290-1425 */
                PxT acc = OPS::zero ();
                const uint32_t *tx = L.tab_x + (x << hh);
                for (uint32_t k = 0; k < (1u << hh); k++)
                {
                    const uint32_t e = __ldg (&tx[k]);
                    const uint32_t ofs = SMOL_TAB_OFS (e);
                    OPS::add (acc, OPS::lerp (px_at (ofs), px_at (min (ofs + 1, d.w_in - 1)), SMOL_TAB_F (e)));
                }
                h = OPS::halve (acc, hh);
            }

            if constexpr (VBOX)
            {
                /* generic:2112-2161 (64bpp) / :2198-2260 (128bpp); boundary rows are filtered once and
                 * handed on to the next output row (see the box kernel) */
                if (top_pending)
                {
                    OPS::add (vacc, OPS::weight (h, w1_cur));
                    top_pending = false;
                }
                else if (r == B_cur)
                    OPS::add (vacc, OPS::weight (h, w2_cur));
                else
                    OPS::add (vacc, h);
                if (r == rend_cur)
                {
                    emit_row (OPS::scale (vacc, d.span_mul_y, P));
                    if (yl_cur < yl1)
                    {
                        const uint32_t yn = L.first_row + yl_cur;
                        const uint32_t e0 = __ldg (&L.tab_y[yn]), e1 = __ldg (&L.tab_y[yn + 1]);
                        const uint32_t F_new = SMOL_TAB_F (e0);
                        const bool shared = r == B_cur;
                        w1_cur = 255u - Fy_cur;
                        Fy_cur = F_new;
                        B_cur = SMOL_TAB_OFS (e1);
                        w2_cur = S128 ? F_new - 1 : F_new;
                        rend_cur = F_new > 0 ? B_cur : B_cur - 1;
                        vacc = OPS::zero ();
                        if (shared)
                            OPS::add (vacc, OPS::weight (h, w1_cur));
                        top_pending = !shared;
                    }
                }
            }
            else
            {
                /* generic:1684-2007: every sample whose lower row is this one (its upper row is the
                 * one before, or this one again where the table clamps at the image's last row) */
                while (k_cur < k_stop)
                {
                    const uint32_t e = __ldg (&L.tab_y[k_cur]);
                    const uint32_t r0 = SMOL_TAB_OFS (e), r1 = min (r0 + 1, d.h_in - 1);
                    if (r1 > r)
                        break;
                    OPS::add (vacc, OPS::lerp (r0 == r ? h : h_prev, h, SMOL_TAB_F (e)));
                    k_cur++;
                    if ((k_cur & ((1u << vh) - 1)) == 0)
                    {
                        emit_row (OPS::halve (vacc, vh));
                        vacc = OPS::zero ();
                    }
                }
                h_prev = h;
            }
            cur = P.seg_bytes - cur;
            __syncwarp ();      /* everyone is done with this slot before it is refilled */
        }
    }
}

/* ------------------------------------------------------------------------------------------ *
 * Host-side dispatch                                                                         *
 * ------------------------------------------------------------------------------------------ */

static const char *const kernel_names[SMOL_KERNEL_MAX] =
{
    "auto", "general", "taps_direct", "half2x", "box", "mag", "taps128", "tile128", "magb", "rows"
};

extern "C" const char *
smol_cuda_kernel_name (int kernel_id)
{
    if (kernel_id < 0 || kernel_id >= SMOL_KERNEL_MAX)
        return "invalid";
    return kernel_names[kernel_id];
}

static bool
aligned16 (const void *p)
{
    return (reinterpret_cast<uintptr_t> (p) & 15) == 0;
}

static bool
aligned4 (const void *p)
{
    return (reinterpret_cast<uintptr_t> (p) & 3) == 0;
}

static bool
half_eligible (const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;

    /* 4-byte-aligned pixels; 16- and 32-byte alignment select wider accesses (launch_half) */
    return d.h_kind == SMOL_AXIS_TAPS && d.v_kind == SMOL_AXIS_TAPS
           && d.all_half_x && d.all_half_y
           && d.mid == SMOL_MID_P8 && !d.storage128
           && d.bpp_in == 4 && d.bpp_out == 4 && !d.in_unassoc
           && aligned4 (L.src) && aligned4 (L.dst)
           && (L.src_pitch & 3) == 0 && (L.dst_pitch & 3) == 0
           && (L.src_image_stride & 3) == 0 && (L.dst_image_stride & 3) == 0;
}

static bool
taps_eligible (const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;

    return d.h_kind == SMOL_AXIS_TAPS && d.v_kind == SMOL_AXIS_TAPS
           && d.mid == SMOL_MID_P8 && !d.storage128;
}

static bool
taps128_eligible (const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;

    /* one thread per output pixel pays off when outputs are few and taps many (a halving on
     * either axis, i.e. more than 2:1); SMOL_TAPS128_ALL=1 (measurements) takes every 128bpp
     * bilinear job */
    static int all = -1;
    if (all < 0)
    {
        const char *e = getenv ("SMOL_TAPS128_ALL");
        all = e ? atoi (e) : 0;
    }
    if (d.h_kind != SMOL_AXIS_TAPS || d.v_kind != SMOL_AXIS_TAPS || !d.storage128 || d.mid == SMOL_MID_P8)
        return false;
    if (all || d.h_halvings > 0 || d.v_halvings > 0)
        return true;
    /* Without halvings: a strip of rows per thread column with the two-row cache beats the tile
     * kernel up to mild magnifications (4K, us per frame, taps128 / tile128: 1:1 82 / 113, 1.5:1
     * down 48 / 79, 2x up 66 / 42), so the tile kernel keeps the upscales beyond 1.5x in area. */
    return (uint64_t) d.w_out * d.h_out * 2 <= (uint64_t) d.w_in * d.h_in * 3;
}

static bool
tile128_eligible (const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;

    return d.h_kind == SMOL_AXIS_TAPS && d.v_kind == SMOL_AXIS_TAPS && d.storage128
           && d.mid != SMOL_MID_P8 && d.h_halvings == 0 && d.v_halvings == 0
           && (d.bpp_out == 3 || ((reinterpret_cast<uintptr_t> (L.dst) & 3) == 0 && (L.dst_pitch & 3) == 0
                                  && (L.dst_image_stride & 3) == 0));
}

static bool
mag_eligible (const SmolLaunch &L)
{
    return taps_eligible (L) && L.d.h_out > L.d.h_in && L.d.h_halvings == 0;
}

static bool
magb_eligible (const SmolLaunch &L)
{
    /* byte-granular vertical stage with 16-byte stores (four 4-byte stores when the destination rows
     * sit on word boundaries only); SMOL_MAGB_ALL=1 (measurements) sends every
     * bilinear job without halvings here, not just vertical magnifications */
    static int all = -1;
    if (all < 0)
    {
        const char *e = getenv ("SMOL_MAGB_ALL");
        all = e ? atoi (e) : 0;
    }
    const bool shape_ok = all ? (taps_eligible (L) && L.d.h_halvings == 0 && L.d.v_halvings == 0) : mag_eligible (L);
    return shape_ok && aligned4 (L.dst) && (L.dst_pitch & 3) == 0 && (L.dst_image_stride & 3) == 0;
}

static bool
box_eligible (const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;

    if (d.h_kind != SMOL_AXIS_BOX || d.v_kind != SMOL_AXIS_BOX)
        return false;
    if (d.mid == SMOL_MID_P8 && d.storage128)
        return false;                   /* ratio > 255 without linear light: general kernel */
    if (d.w_in / d.w_out >= 512)
        return false;                   /* a warp's row segment must fit its staging buffer */
    if ((uint64_t) d.w_out * L.n_rows * L.n_images >= 0x7fffffffull)
        return false;                   /* 32-bit work item index */
    /* 32bpp: whole pixels must be word loads from the staged rows; 24bpp: any alignment.  Rows off
     * 16-byte boundaries take the row-shifted variant (launch_box). */
    return d.bpp_in == 3 || (aligned4 (L.src) && (L.src_pitch & 3) == 0 && (L.src_image_stride & 3) == 0);
}

/* The warp-per-tile backstop: anything whose per-warp source window fits its staging buffers. */
static uint32_t
rows_seg_bytes (const SmolLaunch &L, uint32_t *glog_out)
{
    const SmolJobDesc &d = L.d;
    uint32_t glog = 0;
    if (d.h_kind == SMOL_AXIS_BOX)
    {
        const uint32_t ratio = d.w_in / d.w_out;
        while (glog < 5 && ratio >= (16u << glog))
            glog++;
    }
    const uint32_t cols = 32u >> glog;
    /* widest window an item can need: its columns' share of the row, the taps' extra pixel, rounding slack */
    const uint64_t seg_px = ((uint64_t) cols * d.w_in + d.w_out - 1) / d.w_out + 4;
    if (glog_out)
        *glog_out = glog;
    const uint64_t bytes = (seg_px * d.bpp_in + 32 + 15) & ~(uint64_t) 15;
    return bytes > 0x7fffffffu ? 0x7fffffffu : (uint32_t) bytes;
}

static bool
rows_eligible (const SmolLaunch &L)
{
    static int on = -1;
    if (on < 0)
    {
        const char *e = getenv ("SMOL_ROWS_KERNEL");
        on = e ? atoi (e) : 1;
    }
    if (!on || rows_seg_bytes (L, nullptr) > 12 * 1024)
        return false;
    return (uint64_t) L.d.w_out * L.n_rows * L.n_images < 0x7fffffffull;
}

extern "C" int
smol_cuda_pick_kernel (const SmolLaunch *launch, int forced)
{
    const bool half_ok = half_eligible (*launch);
    const bool box_ok = box_eligible (*launch);

    if (forced == SMOL_KERNEL_BOX)
        return box_ok ? SMOL_KERNEL_BOX : SMOL_KERNEL_GENERAL;
    const bool taps_ok = taps_eligible (*launch);
    const bool mag_ok = mag_eligible (*launch);

    if (forced == SMOL_KERNEL_MAG)
        return mag_ok ? SMOL_KERNEL_MAG : SMOL_KERNEL_GENERAL;
    if (forced == SMOL_KERNEL_MAGB)
        return magb_eligible (*launch) ? SMOL_KERNEL_MAGB : SMOL_KERNEL_GENERAL;
    if (forced == SMOL_KERNEL_TAPS128)
        return taps128_eligible (*launch) ? SMOL_KERNEL_TAPS128 : SMOL_KERNEL_GENERAL;
    if (forced == SMOL_KERNEL_TILE128)
        return tile128_eligible (*launch) ? SMOL_KERNEL_TILE128 : SMOL_KERNEL_GENERAL;

    if (forced == SMOL_KERNEL_GENERAL)
        return SMOL_KERNEL_GENERAL;
    if (forced == SMOL_KERNEL_ROWS)
        return rows_eligible (*launch) ? SMOL_KERNEL_ROWS : SMOL_KERNEL_GENERAL;
    if (forced == SMOL_KERNEL_HALF2X)
        return half_ok ? SMOL_KERNEL_HALF2X : SMOL_KERNEL_GENERAL;
    if (forced == SMOL_KERNEL_TAPS_DIRECT)
        return taps_ok ? SMOL_KERNEL_TAPS_DIRECT : SMOL_KERNEL_GENERAL;
    if (forced > SMOL_KERNEL_AUTO && forced < SMOL_KERNEL_MAX)
        return SMOL_KERNEL_GENERAL;
    if (half_ok)
        return SMOL_KERNEL_HALF2X;
    if (box_ok)
        return SMOL_KERNEL_BOX;
    if (taps128_eligible (*launch) && (forced == SMOL_KERNEL_AUTO || forced == SMOL_KERNEL_TAPS128))
        return SMOL_KERNEL_TAPS128;
    if (tile128_eligible (*launch) && (forced == SMOL_KERNEL_AUTO || forced == SMOL_KERNEL_TILE128))
        return SMOL_KERNEL_TILE128;
    /* Measured on B200 (4K outputs, graph replay, us per frame, taps_direct vs magb): 24bpp -> 24bpp
     * 2x 17.2 / 15.2, 4x 15.4 / 12.0, 8x 10.3 / 7.2; 32bpp -> 32bpp 2x 14.1 / 23.3, 4x 10.9 / 12.3,
     * 8x 9.8 / 9.6; 32 -> 24bpp 2x 13.6 / 18.6; 24 -> 24bpp 1.5x 19.4 / 25.4.  The byte-granular tile
     * pays off where the register kernel has to assemble 24bpp stores; the older per-pixel tile
     * ("mag") loses to one or the other everywhere and is kept for forced runs only. */
    if (magb_eligible (*launch) && launch->d.bpp_in == 3 && launch->d.bpp_out == 3
        && launch->d.h_out >= 2 * launch->d.h_in)
        return SMOL_KERNEL_MAGB;
    if (taps_ok)
        return SMOL_KERNEL_TAPS_DIRECT;
    /* box on one axis only, > 255:1 without linear light, byte-misaligned 32bpp box rows */
    if (rows_eligible (*launch))
        return SMOL_KERNEL_ROWS;
    return SMOL_KERNEL_GENERAL;
}

/* Per-device launch state.  Function attributes and occupancy answers belong to a device (a
 * context), and the library serves several devices from one process, so everything cached here is
 * keyed by the current device's ordinal. */
#define SMOL_KERNELS_MAX_DEVICES 16

static int
current_device ()
{
    int dev = 0;
    if (cudaGetDevice (&dev) != cudaSuccess || dev < 0 || dev >= SMOL_KERNELS_MAX_DEVICES)
        dev = 0;
    return dev;
}

static int
num_sms ()
{
    static int sms[SMOL_KERNELS_MAX_DEVICES];
    const int dev = current_device ();
    int n = __atomic_load_n (&sms[dev], __ATOMIC_RELAXED);

    if (n == 0)
    {
        n = 148;
        cudaDeviceGetAttribute (&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0)
            n = 148;
        __atomic_store_n (&sms[dev], n, __ATOMIC_RELAXED);
    }
    return n;
}

/* Opts kernel `fn` in to `bytes` of dynamic shared memory on the current device, once per
 * (device, kernel): a small lock-free set of (fn, device mask) pairs. */
static cudaError_t
smem_optin (const void *fn, int bytes)
{
    static struct { const void *fn; uint32_t done; } slots[512];
    const int dev = current_device ();
    uint32_t h = (uint32_t) ((reinterpret_cast<uintptr_t> (fn) >> 4) * 2654435761u) & 511u;

    for (int probe = 0; probe < 512; probe++, h = (h + 1) & 511u)
    {
        const void *cur = __atomic_load_n (&slots[h].fn, __ATOMIC_ACQUIRE);
        if (cur == nullptr)
        {
            const void *expected = nullptr;
            if (!__atomic_compare_exchange_n (&slots[h].fn, &expected, fn, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)
                && expected != fn)
                continue;
            cur = fn;
        }
        if (cur != fn)
            continue;
        if (__atomic_load_n (&slots[h].done, __ATOMIC_ACQUIRE) & (1u << dev))
            return cudaSuccess;
        const cudaError_t err = cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (err == cudaSuccess)
            __atomic_fetch_or (&slots[h].done, 1u << dev, __ATOMIC_RELEASE);
        return err;
    }
    return cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

/* cudaOccupancyMaxActiveBlocksPerMultiprocessor costs ~20 us of host time: a C caller looping over
 * small box jobs would be bound by it.  Answers are remembered per (device, kernel, block size, shared
 * memory) in a small direct-mapped table (racing writers store the same value). */
static int
cached_occupancy (const void *fn, int threads, size_t smem)
{
    struct Entry { const void *fn; int threads, dev, occ; size_t smem; };
    static Entry table[256];
    const int dev = current_device ();
    const uint32_t h = (uint32_t) ((reinterpret_cast<uintptr_t> (fn) >> 4) * 2654435761u + (uint32_t) threads * 40503u
                                   + (uint32_t) smem * 2246822519u + (uint32_t) dev) & 255u;
    Entry e = table[h];
    if (e.fn == fn && e.threads == threads && e.smem == smem && e.dev == dev && e.occ > 0)
        return e.occ;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, fn, threads, smem) != cudaSuccess || occ < 1)
        occ = 1;
    e.fn = fn; e.threads = threads; e.smem = smem; e.dev = dev; e.occ = occ;
    table[h] = e;
    return occ;
}

/* See prefetch_l2 / in_first_wave: the number of CTAs of a launch that can be resident at once
 * (0 when SMOL_PDL_PREFETCH=0). */
static uint32_t
pdl_first_wave (uint32_t threads_per_cta, size_t smem_per_cta)
{
    static int on = -1;
    if (on < 0)
    {
        const char *e = getenv ("SMOL_PDL_PREFETCH");
        on = e ? atoi (e) : 1;
    }
    if (on == 0)
        return 0;
    if (on == 2)
        return 0xffffffffu;             /* measurement: every CTA prefetches */
    uint32_t per_sm = 2048 / (threads_per_cta ? threads_per_cta : 1);
    if (per_sm > 32)
        per_sm = 32;
    if (smem_per_cta > 0)
    {
        const uint32_t by_smem = (uint32_t) ((227 * 1024) / (smem_per_cta + 1024));
        if (by_smem < per_sm)
            per_sm = by_smem;
    }
    if (per_sm < 1)
        per_sm = 1;
    return (uint32_t) num_sms () * per_sm;
}

/* Launch with programmatic stream serialization allowed (see pdl_wait). */
template <typename Kernel, typename Params>
static cudaError_t
launch_pdl (Kernel kernel, const Params &P, dim3 grid, dim3 block, size_t smem, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;

    memset (&cfg, 0, sizeof (cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx (&cfg, kernel, P);
}

template <typename Kernel, typename... Args>
static cudaError_t
launch_pdl_args (Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;

    memset (&cfg, 0, sizeof (cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx (&cfg, kernel, args...);
}

template <int HH, int VH, int AL>
static cudaError_t
launch_half_hv (const HalfParams &P, int pack, dim3 grid, dim3 block, cudaStream_t stream)
{
    if constexpr (HH == 0)
    {
        if (pack == 1)
            return launch_pdl (smol_half_kernel<HH, VH, 1, AL>, P, grid, block, 0, stream);
        if (pack == 2)
            return launch_pdl (smol_half_kernel<HH, VH, 2, AL>, P, grid, block, 0, stream);
        return launch_pdl (smol_half_kernel<HH, VH, 0, AL>, P, grid, block, 0, stream);
    }
    else
    {
        if (pack == 1)
            return launch_pdl (smol_half_wide_kernel<HH, VH, 1, AL>, P, grid, block, 0, stream);
        if (pack == 2)
            return launch_pdl (smol_half_wide_kernel<HH, VH, 2, AL>, P, grid, block, 0, stream);
        return launch_pdl (smol_half_wide_kernel<HH, VH, 0, AL>, P, grid, block, 0, stream);
    }
}

/* Same, for a kernel picked at run time as a function pointer (one by-value parameter struct). */
static cudaError_t
launch_pdl_ptr (const void *fn, const void *params, dim3 grid, dim3 block, size_t smem, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    void *args[1] = { const_cast<void *> (params) };

    memset (&cfg, 0, sizeof (cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC (&cfg, fn, args);
}

/* rows per thread of the 256-bit 2:1 kernel (SMOL_HALF2V_ROWS: 0 = use the 128-bit kernel) */
static int
half2v_rows ()
{
    static int rows = -1;
    if (rows < 0)
    {
        const char *e = getenv ("SMOL_HALF2V_ROWS");
        const int r = e ? atoi (e) : 1;
        rows = r <= 0 ? 0 : r == 1 ? 1 : r < 4 ? 2 : 4;
    }
    return rows;
}

static cudaError_t
launch_half (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    HalfParams P;
    uint32_t sel = 0;

    P.src = L.src; P.dst = L.dst;
    P.src_pitch = L.src_pitch; P.dst_pitch = L.dst_pitch;
    P.src_image_stride = L.src_image_stride; P.dst_image_stride = L.dst_image_stride;
    P.w_out = d.w_out; P.first_row = L.first_row; P.n_rows = L.n_rows;
    P.items_per_row = (d.w_out + 3) / 4;
    P.inv_div_p8 = L.luts->inv_div_p8;

    /* destination byte j takes source byte perm[j] */
    for (int j = 0; j < 4; j++)
    {
        uint32_t from;
        if (j == d.out_alpha_idx)
            from = d.in_alpha_idx;
        else
        {
            const int i = j - d.out_col0;
            from = d.in_col0 + (d.swap_rb ? 2 - i : i);
        }
        sel |= from << (4 * j);
    }
    P.prmt_sel = sel;

    /* threads along x: HH = 0: one per 4 output pixels; HH = 1, 2: one per 16-byte source chunk */
    const uint32_t n_x = d.h_halvings == 0 ? P.items_per_row : d.w_out << (d.h_halvings - 1);

    /* Block shape: bx threads along x (a multiple of 32 that wastes the fewest lanes on the
     * row's ragged end), by rows, about 256 threads in all. */
    uint32_t bx = 32, best_waste = 0xffffffffu;
    for (uint32_t cand = 256; cand >= 32; cand -= 32)
    {
        const uint32_t waste = (n_x + cand - 1) / cand * cand - n_x;
        if (waste < best_waste)
        {
            best_waste = waste;
            bx = cand;
        }
    }
    uint32_t by = 256 / bx;
    if (by > L.n_rows)
        by = L.n_rows;
    dim3 block (bx, by);
    dim3 grid ((n_x + bx - 1) / bx, (L.n_rows + by - 1) / by, L.n_images);
    const int pack = d.out_unassoc ? (d.in_alpha_idx == 0 ? 2 : 1) : 0;
    /* With an inverse table to stage (its load and barrier sit between kernel entry and the first
     * source load) every CTA prefetches: the source then streams in during that phase.  Without
     * one the loads follow at once and only the first wave has anything to gain (measured: 4K ->
     * 1080p unassociated 8.2 -> 7.1 us per frame; 64 thumbnails per launch 156 -> 152 us). */
    P.prefetch = pdl_first_wave (bx * by, pack ? 1024 : 0);
    if (pack && P.prefetch)
        P.prefetch = 0xffffffffu;

    if (d.h_halvings == 0 && d.v_halvings == 0 && half2v_rows () > 0
        && (reinterpret_cast<uintptr_t> (L.src) & 31) == 0 && (L.src_pitch & 31) == 0 && (L.src_image_stride & 31) == 0
        && aligned16 (L.dst) && (L.dst_pitch & 15) == 0 && (L.dst_image_stride & 15) == 0)
    {
        /* 256-bit loads, several rows per thread (see smol_half2v_kernel) */
        const uint32_t rows = (uint32_t) half2v_rows ();
        const uint32_t n_y = (L.n_rows + rows - 1) / rows;
        static int tune_threads = -1, tune_pf = -1;
        if (tune_threads < 0)
        {
            const char *e = getenv ("SMOL_HALF2V_THREADS"), *f = getenv ("SMOL_HALF2V_PF");
            tune_pf = f ? atoi (f) : 1;
            tune_threads = e ? atoi (e) : 256;
        }
        uint32_t vbx = 32, vwaste = 0xffffffffu;
        for (uint32_t cand = (uint32_t) tune_threads; cand >= 32; cand -= 32)
        {
            const uint32_t waste = (n_x + cand - 1) / cand * cand - n_x;
            if (waste < vwaste)
            {
                vwaste = waste;
                vbx = cand;
            }
        }
        uint32_t vy = (uint32_t) tune_threads / vbx;
        if (vy > n_y)
            vy = n_y;
        dim3 vblock (vbx, vy), vgrid ((n_x + vbx - 1) / vbx, (n_y + vy - 1) / vy, L.n_images);
        P.prefetch = tune_pf == 0 ? 0 : tune_pf == 2 ? 0xffffffffu : pdl_first_wave (vbx * vy, pack ? 1024 : 0);
#define HALF2V(PK) (rows == 1 ? launch_pdl (smol_half2v_kernel<PK, 1>, P, vgrid, vblock, 0, stream) \
                    : rows == 2 ? launch_pdl (smol_half2v_kernel<PK, 2>, P, vgrid, vblock, 0, stream) \
                    : launch_pdl (smol_half2v_kernel<PK, 4>, P, vgrid, vblock, 0, stream))
        return pack == 1 ? HALF2V (1) : pack == 2 ? HALF2V (2) : HALF2V (0);
#undef HALF2V
    }

    const uintptr_t bits = reinterpret_cast<uintptr_t> (L.src) | reinterpret_cast<uintptr_t> (L.dst) | L.src_pitch | L.dst_pitch
                           | L.src_image_stride | L.dst_image_stride;
    const int al = (bits & 15) == 0 ? 16 : (bits & 7) == 0 ? 8 : 4;
#define HALF_HV(H, V) (al == 16 ? launch_half_hv<H, V, 16> (P, pack, grid, block, stream) \
                       : al == 8 ? launch_half_hv<H, V, 8> (P, pack, grid, block, stream) \
                       : launch_half_hv<H, V, 4> (P, pack, grid, block, stream))
    switch (d.h_halvings * 3 + d.v_halvings)
    {
        case 0: return HALF_HV (0, 0);
        case 1: return HALF_HV (0, 1);
        case 2: return HALF_HV (0, 2);
        case 3: return HALF_HV (1, 0);
        case 4: return HALF_HV (1, 1);
        case 5: return HALF_HV (1, 2);
        case 6: return HALF_HV (2, 0);
        case 7: return HALF_HV (2, 1);
        default: return HALF_HV (2, 2);
    }
#undef HALF_HV
}

/* destination byte j takes source byte perm[j] (PRMT selector); 24bpp sources carry a forced
 * 0xff in byte 3, 24bpp destinations take their three colour bytes into bytes 0..2 */
static uint32_t
byte_order_selector (const SmolJobDesc &d)
{
    const uint32_t in_alpha = d.in_alpha_idx == 0xff ? 3 : d.in_alpha_idx;
    uint32_t sel = 0;

    for (int j = 0; j < 4; j++)
    {
        uint32_t from;
        if (d.out_alpha_idx != 0xff && j == d.out_alpha_idx)
            from = in_alpha;
        else
        {
            const int i = j - d.out_col0;
            if (i < 0 || i > 2)
                from = in_alpha;        /* unused top byte of a 24bpp destination */
            else
                from = d.in_col0 + (d.swap_rb ? 2 - i : i);
        }
        sel |= from << (4 * j);
    }
    return sel;
}

template <int HH>
static cudaError_t
launch_taps_h (const TapsParams &P, uint32_t vh, dim3 grid, dim3 block, cudaStream_t stream)
{
    if (vh == 0)
        return launch_pdl (smol_taps_kernel<HH, 0>, P, grid, block, 0, stream);
    if (vh == 1)
        return launch_pdl (smol_taps_kernel<HH, 1>, P, grid, block, 0, stream);
    return launch_pdl (smol_taps_kernel<HH, 2>, P, grid, block, 0, stream);
}

/* Rows per thread of the strip kernels (taps0, taps0w).  Long strips reuse the two-row cache better
 * (a strip of n output rows filters n * h_in / h_out + 1 source rows), but what decides among the
 * reasonable lengths is how the resulting grid fills the GPU: CTAs / (SMs x resident CTAs) should be
 * just under a whole number -- 1.46 waves run as long as 2, 0.73 leave a quarter of the SMs idle
 * (4K 1:1, us per frame at 16 / 8 / 12 rows: 24.9 / 21.9 / 19.9). */
static uint32_t
pick_rows_per_thread (uint64_t x_threads, uint32_t bx, uint32_t n_rows, uint32_t n_images, double v, uint32_t per_sm)
{
    static const uint32_t cand[] = { 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 32 };
    uint32_t best = 8;
    double best_eff = -1.0;
    const double slots = (double) num_sms () * (per_sm ? per_sm : 1);

    for (uint32_t rpt : cand)
    {
        if (rpt > n_rows && rpt != cand[0])
            break;
        const uint32_t strips = (n_rows + rpt - 1) / rpt;
        uint32_t by = 256 / bx;
        if (by > strips)
            by = strips;
        const double ctas = (double) ((x_threads + bx - 1) / bx) * ((strips + by - 1) / by) * n_images;
        const double waves = ctas / slots;
        const double weff = waves <= 1.0 ? waves : waves / (double) (uint64_t) (waves + 0.999999);
        /* emit one output row = 1, horizontally filter one source row = 1.5 */
        const double work_eff = rpt * (1.0 + v * 1.5) / (rpt + (rpt * v + 1.0) * 1.5);
        const double eff = weff * work_eff;
        if (eff > best_eff * 1.01)
        {
            best_eff = eff;
            best = rpt;
        }
    }
    return best;
}

static cudaError_t
launch_taps (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    TapsParams P;

    P.src = L.src; P.dst = L.dst;
    P.src_pitch = L.src_pitch; P.dst_pitch = L.dst_pitch;
    P.src_image_stride = L.src_image_stride; P.dst_image_stride = L.dst_image_stride;
    P.tab_x = L.tab_x; P.tab_y = L.tab_y;
    P.inv_div_p8 = L.luts->inv_div_p8;
    P.w_in = d.w_in; P.h_in = d.h_in; P.w_out = d.w_out;
    P.first_row = L.first_row; P.n_rows = L.n_rows;
    P.bpp_in = d.bpp_in; P.bpp_out = d.bpp_out;
    P.in_alpha_shift = (d.in_alpha_idx == 0xff ? 3 : d.in_alpha_idx) * 8;
    P.in_unassoc = d.in_unassoc; P.out_unassoc = d.out_unassoc;
    P.prmt_sel = byte_order_selector (d);

    /* strip height: long strips amortise the two-row cache (essential on upscales) but leave
     * fewer threads; keep at least ~2 resident waves of threads on the GPU */
    const uint64_t x_threads = (d.w_out + 3) / 4;
    uint32_t bx = 32;
    while (bx < 128 && bx < x_threads)
        bx *= 2;
    uint32_t rpt;
    if (d.h_halvings == 0 && d.v_halvings == 0)
        rpt = pick_rows_per_thread (x_threads, bx, L.n_rows, L.n_images, (double) d.h_in / d.h_out, SMOL_TAPS0_MINBLOCKS);
    else
    {
        /* (the runtime-format kernel with halvings: strips while they leave ~2 resident waves of threads) */
        const uint64_t want_threads = (uint64_t) num_sms () * 1536;
        rpt = 16;
        while (rpt > 1 && x_threads * ((L.n_rows + rpt - 1) / rpt) * L.n_images < want_threads)
            rpt >>= 1;
    }
    {
        static int tune_rpt = -1;
        if (tune_rpt < 0)
        {
            const char *e = getenv ("SMOL_TAPS_RPT");
            tune_rpt = e ? atoi (e) : 0;
        }
        if (tune_rpt > 0)
            rpt = (uint32_t) tune_rpt;
    }
    P.rows_per_thread = rpt;

    const uint32_t strips = (L.n_rows + rpt - 1) / rpt;
    uint32_t by = 256 / bx;
    if (by > strips)
        by = strips;
    dim3 block (bx, by);
    dim3 grid ((unsigned) ((x_threads + bx - 1) / bx), (strips + by - 1) / by, L.n_images);

    if (d.h_halvings == 0 && d.v_halvings == 0)
    {
        Taps0Params T;
        static const uint32_t acc_byte[4] = { 1, 5, 3, 7 };
        uint32_t sel = 0;

        T.t = P;
        for (int j = 0; j < 4; j++)
            sel |= acc_byte[(P.prmt_sel >> (4 * j)) & 3] << (4 * j);
        T.acc_prmt_sel = sel;
        T.src_u32_ok = d.bpp_in == 3 || ((reinterpret_cast<uintptr_t> (L.src) & 3) == 0 && (L.src_pitch & 3) == 0
                                         && (L.src_image_stride & 3) == 0);
        T.prefetch = pdl_first_wave (256, d.out_unassoc ? 1024 : 0);
        if (d.out_unassoc && T.prefetch)
            T.prefetch = 0xffffffffu;   /* an inverse table is staged first: see launch_half */
        {
            static int tune_ahead = -1;
            if (tune_ahead < 0)
            {
                const char *e = getenv ("SMOL_TAPS_ROW_AHEAD");
                tune_ahead = e ? atoi (e) : 2;      /* measured: 4K 1:1 21.8 -> 19.8 us, 4K -> 1440p 12.4 -> 10.8 us */
            }
            T.row_ahead = (uint32_t) tune_ahead;
        }
        const bool af = d.in_alpha_idx == 0;
        /* (24bpp destinations are stored cooperatively at any alignment: store_px4_rgb_anywhere) */
        const bool fastio = T.src_u32_ok && (d.bpp_out == 3 || ((reinterpret_cast<uintptr_t> (L.dst) & 3) == 0
                                                                && (L.dst_pitch & 3) == 0 && (L.dst_image_stride & 3) == 0));
        /* one pixel per thread (fully coalesced source reads) when the destination is 32bpp and
         * rows are aligned; otherwise four (vector / 24bpp stores) */
        static int tune_px = -1;
        if (tune_px < 0)
        {
            const char *e = getenv ("SMOL_TAPS_PX");
            tune_px = e ? atoi (e) : 4;
        }
        const bool px1 = fastio && d.bpp_out == 4 && tune_px == 1;
        if (px1)
        {
            const uint64_t xt1 = d.w_out;
            uint32_t bx1 = 32;
            while (bx1 < 128 && bx1 < xt1)
                bx1 *= 2;
            uint32_t by1 = 256 / bx1;
            if (by1 > strips)
                by1 = strips;
            block = dim3 (bx1, by1);
            grid = dim3 ((unsigned) ((xt1 + bx1 - 1) / bx1), (strips + by1 - 1) / by1, L.n_images);
        }
#define TAPS0_PX(BI, BO, IU, OU, AF, PX) (fastio ? launch_pdl (smol_taps0_kernel<BI, BO, IU, OU, AF, true, PX>, T, grid, block, 0, stream) \
                                                 : launch_pdl (smol_taps0_kernel<BI, BO, IU, OU, AF, false, PX>, T, grid, block, 0, stream))
#define TAPS0(BI, BO, IU, OU, AF) (BO == 4 && px1 ? launch_pdl (smol_taps0_kernel<BI, 4, IU, OU, AF, true, 1>, T, grid, block, 0, stream) \
                                                  : TAPS0_PX (BI, BO, IU, OU, AF, 4))
        if (d.bpp_in == 3)
        {
            if (d.bpp_out == 3)     return TAPS0 (3, 3, false, false, false);
            if (d.out_unassoc)      return TAPS0 (3, 4, false, true, false);
            return TAPS0 (3, 4, false, false, false);
        }
        if (d.in_unassoc)
        {
            if (d.bpp_out == 3)
                return af ? TAPS0 (4, 3, true, false, true) : TAPS0 (4, 3, true, false, false);
            return af ? TAPS0 (4, 4, true, false, true) : TAPS0 (4, 4, true, false, false);
        }
        if (d.bpp_out == 3)
            return TAPS0 (4, 3, false, false, false);
        if (d.out_unassoc)
            return af ? TAPS0 (4, 4, false, true, true) : TAPS0 (4, 4, false, true, false);
        return TAPS0 (4, 4, false, false, false);
#undef TAPS0
#undef TAPS0_PX
    }

    {
        /* halvings on either axis: one thread per output pixel, compile-time formats (needs
         * 4-byte-aligned rows; anything else takes the runtime-format kernel below) */
        Taps0Params T;
        T.t = P;
        T.acc_prmt_sel = 0;
        T.src_u32_ok = d.bpp_in == 3 || ((reinterpret_cast<uintptr_t> (L.src) & 3) == 0 && (L.src_pitch & 3) == 0
                                         && (L.src_image_stride & 3) == 0);
        T.prefetch = pdl_first_wave (256, d.out_unassoc ? 1024 : 0);
        if (d.out_unassoc && T.prefetch)
            T.prefetch = 0xffffffffu;   /* an inverse table is staged first: see launch_half */
        {
            static int tune_ahead = -1;
            if (tune_ahead < 0)
            {
                const char *e = getenv ("SMOL_TAPS_ROW_AHEAD");
                tune_ahead = e ? atoi (e) : 2;      /* measured: 4K 1:1 21.8 -> 19.8 us, 4K -> 1440p 12.4 -> 10.8 us */
            }
            T.row_ahead = (uint32_t) tune_ahead;
        }
        const bool dst_ok = d.bpp_out == 3 || ((reinterpret_cast<uintptr_t> (L.dst) & 3) == 0 && (L.dst_pitch & 3) == 0
                                               && (L.dst_image_stride & 3) == 0);
        if (T.src_u32_ok && dst_ok)
        {
            /* one output pixel per thread reads its (ratio + 1)^2 source pixels through dependent
             * table lookups: prefetching the rows into L2 on entry pays in every CTA, not only in
             * the first wave (measured: 4K -> 720p 14.1 -> 13.4 us) */
            if (T.prefetch)
                T.prefetch = 0xffffffffu;
            uint32_t nbx = 32;
            while (nbx < 128 && nbx < d.w_out)
                nbx *= 2;
            /* output rows per thread (SMOL_TAPSN_RPT) */
            static int tune_nrpt = -1;
            if (tune_nrpt < 0)
            {
                const char *e = getenv ("SMOL_TAPSN_RPT");
                tune_nrpt = e ? atoi (e) : 0;
            }
            /* Measured (4K source, us per frame at 1 / 2 / 4 rows per thread): -> 1280x720 14.1 /
             * 14.4 / 17.1, -> 1600x900 20.8 / 18.4 / 17.9, -> 800x450 15.0 / 16.3 / 22.2: the row
             * cache rarely carries over and the lost parallelism costs more, so one row it is. */
            uint32_t nrpt = 1;
            if (tune_nrpt > 0)
                nrpt = (uint32_t) tune_nrpt;
            T.t.rows_per_thread = nrpt;
            const uint32_t nstrips = (L.n_rows + nrpt - 1) / nrpt;
            uint32_t nby = 256 / nbx;
            if (nby > nstrips)
                nby = nstrips;
            dim3 nblock (nbx, nby);
            dim3 ngrid ((d.w_out + nbx - 1) / nbx, (nstrips + nby - 1) / nby, L.n_images);
            const bool af = d.in_alpha_idx == 0;
            const uint32_t hh = d.h_halvings, vh = d.v_halvings;
            static int taps11_on = -1, taps11_rpt = 0, taps11_ahead = 1;
            if (taps11_on < 0)
            {
                const char *e = getenv ("SMOL_TAPS11"), *r = getenv ("SMOL_TAPS11_RPT"), *a = getenv ("SMOL_TAPS11_AHEAD");
                taps11_rpt = r ? atoi (r) : 0;
                taps11_ahead = a ? atoi (a) : 1;
                taps11_on = e ? atoi (e) : 1;
            }
            if (hh == 1 && vh == 1 && taps11_on && d.bpp_in == 4)     /* (24bpp sources, three byte loads per pixel: measured slower than the general kernel) */
            {
                /* one halving on both axes: the straight-line strip kernel.  Strip length: longer
                 * strips amortise the per-thread set-up and reuse the last filtered row, but the grid
                 * must keep about two waves of CTAs or load latency shows (measured, 4K -> 720p, us per
                 * frame at 1 / 2 / 3 / 5 / 7 rows: 12.4 / 10.9 / 10.9 / 11.6 / 12.7; 4K -> 900p 5 rows
                 * 14.3, 2 rows 15.6): the longest strip of up to 6 rows that leaves 1.9 waves. */
                uint32_t best_r = 1;
                const double slots = (double) num_sms () * SMOL_TAPS11_MINBLOCKS;
                for (uint32_t r = 2; r <= 6 && r <= L.n_rows; r++)
                {
                    const uint32_t strips = (L.n_rows + r - 1) / r;
                    const uint32_t by = 256 / nbx < strips ? 256 / nbx : strips;
                    const double ctas = (double) ((d.w_out + nbx - 1) / nbx) * ((strips + by - 1) / by) * L.n_images;
                    if (ctas >= 1.9 * slots)
                        best_r = r;
                }
                if (taps11_rpt > 0)
                    best_r = (uint32_t) taps11_rpt;
                T.t.rows_per_thread = best_r;
                T.row_ahead = (uint32_t) taps11_ahead;
                const uint32_t strips = (L.n_rows + best_r - 1) / best_r;
                const uint32_t by = 256 / nbx < strips ? 256 / nbx : strips;
                const dim3 block11 (nbx, by), grid11 ((d.w_out + nbx - 1) / nbx, (strips + by - 1) / by, L.n_images);
#define TAPS11(BI, BO, IU, OU, AF) launch_pdl (smol_taps11_kernel<BI, BO, IU, OU, AF>, T, grid11, block11, 0, stream)
                if (d.in_unassoc)
                {
                    if (d.bpp_out == 3)
                        return af ? TAPS11 (4, 3, true, false, true) : TAPS11 (4, 3, true, false, false);
                    return af ? TAPS11 (4, 4, true, false, true) : TAPS11 (4, 4, true, false, false);
                }
                if (d.bpp_out == 3)
                    return TAPS11 (4, 3, false, false, false);
                if (d.out_unassoc)
                    return af ? TAPS11 (4, 4, false, true, true) : TAPS11 (4, 4, false, true, false);
                return TAPS11 (4, 4, false, false, false);
#undef TAPS11
            }
#define TAPSN(BI, BO, IU, OU, AF) (hh == 0 ? launch_pdl_args (smol_tapsn_kernel<BI, BO, IU, OU, AF, 0>, ngrid, nblock, 0, stream, T, vh) \
                                   : hh == 1 ? launch_pdl_args (smol_tapsn_kernel<BI, BO, IU, OU, AF, 1>, ngrid, nblock, 0, stream, T, vh) \
                                             : launch_pdl_args (smol_tapsn_kernel<BI, BO, IU, OU, AF, 2>, ngrid, nblock, 0, stream, T, vh))
            if (d.bpp_in == 3)
            {
                if (d.bpp_out == 3)     return TAPSN (3, 3, false, false, false);
                if (d.out_unassoc)      return TAPSN (3, 4, false, true, false);
                return TAPSN (3, 4, false, false, false);
            }
            if (d.in_unassoc)
            {
                if (d.bpp_out == 3)
                    return af ? TAPSN (4, 3, true, false, true) : TAPSN (4, 3, true, false, false);
                return af ? TAPSN (4, 4, true, false, true) : TAPSN (4, 4, true, false, false);
            }
            if (d.bpp_out == 3)
                return TAPSN (4, 3, false, false, false);
            if (d.out_unassoc)
                return af ? TAPSN (4, 4, false, true, true) : TAPSN (4, 4, false, true, false);
            return TAPSN (4, 4, false, false, false);
#undef TAPSN
        }
    }

    if (d.h_halvings == 0)
        return launch_taps_h<0> (P, d.v_halvings, grid, block, stream);
    if (d.h_halvings == 1)
        return launch_taps_h<1> (P, d.v_halvings, grid, block, stream);
    return launch_taps_h<2> (P, d.v_halvings, grid, block, stream);
}

static void
taps_params_init (TapsParams &P, const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;

    P.src = L.src; P.dst = L.dst;
    P.src_pitch = L.src_pitch; P.dst_pitch = L.dst_pitch;
    P.src_image_stride = L.src_image_stride; P.dst_image_stride = L.dst_image_stride;
    P.tab_x = L.tab_x; P.tab_y = L.tab_y;
    P.inv_div_p8 = L.luts->inv_div_p8;
    P.w_in = d.w_in; P.h_in = d.h_in; P.w_out = d.w_out;
    P.first_row = L.first_row; P.n_rows = L.n_rows;
    P.rows_per_thread = 1;
    P.bpp_in = d.bpp_in; P.bpp_out = d.bpp_out;
    P.in_alpha_shift = (d.in_alpha_idx == 0xff ? 3 : d.in_alpha_idx) * 8;
    P.in_unassoc = d.in_unassoc; P.out_unassoc = d.out_unassoc;
    P.prmt_sel = byte_order_selector (d);
}

template <int BI, int BO, bool IU, bool OU, bool AF, bool FASTIO>
static cudaError_t
launch_mag_fmt_io (const MagParams &M, dim3 grid, size_t smem, cudaStream_t stream)
{
    if (smem > 32 * 1024)        /* static shared memory counts against the 48 KB default too */
    {
        cudaError_t err = smem_optin ((const void *) smol_mag_kernel<BI, BO, IU, OU, AF, FASTIO>, 200 * 1024);
        if (err != cudaSuccess)
            return err;
    }
    return launch_pdl (smol_mag_kernel<BI, BO, IU, OU, AF, FASTIO>, M, grid, dim3 (256), smem, stream);
}

template <int BI, int BO, bool IU, bool OU, bool AF>
static cudaError_t
launch_mag_fmt (const MagParams &M, dim3 grid, size_t smem, cudaStream_t stream)
{
    /* src_u32_ok doubles as "fast I/O": set by launch_mag only if the destination is aligned too */
    return M.src_u32_ok ? launch_mag_fmt_io<BI, BO, IU, OU, AF, true> (M, grid, smem, stream)
                        : launch_mag_fmt_io<BI, BO, IU, OU, AF, false> (M, grid, smem, stream);
}

static cudaError_t
launch_mag (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    MagParams M;

    taps_params_init (M.t, L);
    /* power-of-two tile width: the smallest of 4..256 that covers the row, else the cap */
    static int tune_tw = -1, tune_th = -1;
    if (tune_tw < 0)
    {
        const char *a = getenv ("SMOL_MAG_TW"), *b = getenv ("SMOL_MAG_TH");
        tune_tw = a ? atoi (a) : 0;
        tune_th = b ? atoi (b) : 0;
    }
    const uint32_t tw_cap = tune_tw > 0 ? (uint32_t) tune_tw : 256;
    M.tile_w = 4;
    M.tile_w_log2 = 2;
    while (M.tile_w < tw_cap && M.tile_w < d.w_out)
    {
        M.tile_w *= 2;
        M.tile_w_log2++;
    }
    M.tile_h = tune_th > 0 ? (uint32_t) tune_th : 32;

    /* Source window bounds from the sampling step: consecutive samples advance by at most
     * ceil (dim_in / dim_out) source pixels; + 1 for the second tap, + 2 slack for rounding. */
    size_t smem;
    for (;;)
    {
        const uint64_t cols = ((uint64_t) M.tile_w * d.w_in + d.w_out - 1) / d.w_out + 3;
        const uint64_t rows = ((uint64_t) M.tile_h * d.h_in + d.h_out - 1) / d.h_out + 3;
        M.u_pitch = (uint32_t) (cols < d.w_in ? cols : d.w_in);
        M.u_pitch = (M.u_pitch + 1) & ~1u;                         /* keeps the planes 16-byte aligned */
        M.max_src_rows = (uint32_t) (rows < d.h_in ? rows : d.h_in);
        smem = (size_t) M.max_src_rows * ((size_t) M.u_pitch + M.tile_w) * 8;
        if (smem <= 56 * 1024 || M.tile_h <= 4)
            break;
        M.tile_h /= 2;
    }
    M.u_cw = 1;
    M.u_cw_log2 = 0;
    while (M.u_cw < 256 && M.u_cw < M.u_pitch)
    {
        M.u_cw *= 2;
        M.u_cw_log2++;
    }
    M.src_u32_ok = (d.bpp_in == 3 || ((reinterpret_cast<uintptr_t> (L.src) & 3) == 0 && (L.src_pitch & 3) == 0
                                      && (L.src_image_stride & 3) == 0))
                   && (reinterpret_cast<uintptr_t> (L.dst) & 3) == 0
                   && (L.dst_pitch & 3) == 0 && (L.dst_image_stride & 3) == 0;

    /* (acc_a, acc_b) -> destination bytes: source byte s lives in byte 1 (s = 0), 5 (s = 1),
     * 3 (s = 2), 7 (s = 3) of the PRMT operand pair */
    static const uint32_t acc_byte[4] = { 1, 5, 3, 7 };
    uint32_t sel = 0;
    for (int j = 0; j < 4; j++)
        sel |= acc_byte[(M.t.prmt_sel >> (4 * j)) & 3] << (4 * j);
    M.acc_prmt_sel = sel;

    dim3 grid ((d.w_out + M.tile_w - 1) / M.tile_w, (L.n_rows + M.tile_h - 1) / M.tile_h, L.n_images);
    const bool af = d.in_alpha_idx == 0;

    if (d.bpp_in == 3)
    {
        if (d.bpp_out == 3)     return launch_mag_fmt<3, 3, false, false, false> (M, grid, smem, stream);
        if (d.out_unassoc)      return launch_mag_fmt<3, 4, false, true, false> (M, grid, smem, stream);
        return launch_mag_fmt<3, 4, false, false, false> (M, grid, smem, stream);
    }
    if (d.in_unassoc)
    {
        if (d.bpp_out == 3)
            return af ? launch_mag_fmt<4, 3, true, false, true> (M, grid, smem, stream)
                      : launch_mag_fmt<4, 3, true, false, false> (M, grid, smem, stream);
        return af ? launch_mag_fmt<4, 4, true, false, true> (M, grid, smem, stream)
                  : launch_mag_fmt<4, 4, true, false, false> (M, grid, smem, stream);
    }
    if (d.bpp_out == 3)
        return launch_mag_fmt<4, 3, false, false, false> (M, grid, smem, stream);
    if (d.out_unassoc)
        return af ? launch_mag_fmt<4, 4, false, true, true> (M, grid, smem, stream)
                  : launch_mag_fmt<4, 4, false, true, false> (M, grid, smem, stream);
    return launch_mag_fmt<4, 4, false, false, false> (M, grid, smem, stream);
}

template <int BI, int BO, bool IU, bool OU, bool AF>
static cudaError_t
launch_magb_fmt (const MagbParams &M, bool src32, dim3 grid, size_t smem, cudaStream_t stream)
{
    if (src32)
    {
        if (smem > 40 * 1024)
            smem_optin ((const void *) smol_magb_kernel<BI, BO, IU, OU, AF, true>, 200 * 1024);
        return launch_pdl (smol_magb_kernel<BI, BO, IU, OU, AF, true>, M, grid, dim3 (256), smem, stream);
    }
    if (smem > 40 * 1024)
        smem_optin ((const void *) smol_magb_kernel<BI, BO, IU, OU, AF, false>, 200 * 1024);
    return launch_pdl (smol_magb_kernel<BI, BO, IU, OU, AF, false>, M, grid, dim3 (256), smem, stream);
}

static cudaError_t
launch_magb (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    MagbParams M;

    taps_params_init (M.t, L);
    static int tune_tb = -1, tune_th = -1, tune_hmax = 64, tune_waves = 1;
    static uint32_t magb_reg_ctas = 5;      /* resident CTAs per SM by registers (48 x 256 threads) */
    if (tune_tb < 0)
    {
        const char *a = getenv ("SMOL_MAGB_TB"), *b = getenv ("SMOL_MAGB_TH"), *c = getenv ("SMOL_MAGB_CTAS");
        const char *hm = getenv ("SMOL_MAGB_HMAX"), *wv = getenv ("SMOL_MAGB_WAVES");
        if (hm && atoi (hm) >= 8 && atoi (hm) <= SMOL_MAGB_MAX_TILE_H)
            tune_hmax = atoi (hm);
        if (wv)
            tune_waves = atoi (wv) != 0;
        (void) tune_hmax;
        tune_tb = a ? atoi (a) : 0;
        tune_th = b ? atoi (b) : 0;
        if (c && atoi (c) > 0)
            magb_reg_ctas = (uint32_t) atoi (c);
    }
    M.nb_row = d.w_out * d.bpp_out;
    const bool src32 = (reinterpret_cast<uintptr_t> (L.src) & 3) == 0 && (L.src_pitch & 3) == 0
                       && (L.src_image_stride & 3) == 0;

    /* Shape of a tile (tile_b = 16 << k bytes wide, tile_h rows): everything that depends on it. */
    auto shape = [&] (uint32_t chunks_log2, uint32_t tile_h, size_t &smem_out) -> bool
    {
        M.chunks_log2 = chunks_log2;
        M.tile_b = 16u << chunks_log2;
        M.h_pitch = M.tile_b + 32;
        M.tile_h = tile_h;
        /* pixel groups per tile: tile_b / (4 bpp) rounded up, + 1 for a group straddling the tile's start */
        const uint32_t max_groups = (M.tile_b + 4 * d.bpp_out - 1) / (4 * d.bpp_out) + 1;
        M.gcols_log2 = 0;
        while ((1u << M.gcols_log2) < max_groups && M.gcols_log2 < 8)
            M.gcols_log2++;
        if ((1u << M.gcols_log2) < max_groups)
            return false;
        /* Source window bounds from the sampling step (see launch_mag); + 3 pixels either side for
         * the 4-pixel load groups of 24bpp sources. */
        const uint64_t px = (uint64_t) max_groups * 4;
        const uint64_t cols = (px * d.w_in + d.w_out - 1) / d.w_out + 3 + 8;
        const uint64_t rows = ((uint64_t) tile_h * d.h_in + d.h_out - 1) / d.h_out + 3;
        M.u_pitch = (uint32_t) (cols < (uint64_t) d.w_in + 8 ? cols : (uint64_t) d.w_in + 8);
        M.u_pitch = (M.u_pitch + 1) & ~1u;                         /* keeps the filtered rows 16-byte aligned */
        M.max_src_rows = (uint32_t) (rows < d.h_in ? rows : d.h_in);
        smem_out = (size_t) M.max_src_rows * ((size_t) M.u_pitch * 8 + M.h_pitch);
        return true;
    };

    /* Pick the shape with the lowest estimated time per output row:
     *   work  = 1 (vertical stage) + 1.6 x source rows per output row (horizontal stage runs once
     *           per source row of the tile, halo included) + a fixed per-tile share,
     *   waves = CTAs / (SMs x CTAs resident per SM): the last, partly filled wave costs a full one
     *           (a 1.4-wave grid runs as long as a 2-wave one), and low residency hides less latency. */
    size_t smem = 0;
    uint32_t best_c = 0, best_h = 0;
    {
        double best = 1e30;
        uint32_t c_max = 0;
        while ((16u << c_max) < M.nb_row && c_max < 7)
            c_max++;
        for (uint32_t c = c_max >= 5 ? 5 : c_max; c <= c_max; c++)
        {
            /* Candidate tile heights: powers of two, and the heights at which the grid fills exactly
             * 1, 2 or 3 waves of resident CTAs -- a 1.56-wave grid leaves the SMs half empty for its
             * last third and the next frame's CTAs idle at the dependency wait meanwhile (cfg 4: 64-row
             * tiles, 1152 CTAs, 11.5 us; 104-row tiles, 720 CTAs = one wave of 5 per SM, 10.0 us). */
            uint32_t heights[24], n_heights = 0;
            for (uint32_t h = 8; h <= 64; h *= 2)
                heights[n_heights++] = h;
            if (tune_th > 0 && tune_th <= SMOL_MAGB_MAX_TILE_H)
                heights[n_heights++] = (uint32_t) tune_th;
            if (tune_waves)
            {
                const uint64_t tiles_x = (uint64_t) ((M.nb_row + (16u << c) - 1) / (16u << c)) * L.n_images;
                for (uint32_t per = 2; per <= magb_reg_ctas; per++)
                    for (uint32_t w = 1; w <= 3; w++)
                    {
                        const uint64_t tiles_y = (uint64_t) num_sms () * per * w / tiles_x;
                        if (tiles_y < 1)
                            continue;
                        const uint64_t h = (L.n_rows + tiles_y - 1) / tiles_y;
                        if (h >= 8 && h <= SMOL_MAGB_MAX_TILE_H && n_heights < 24)
                            heights[n_heights++] = (uint32_t) h;
                    }
            }
            for (uint32_t hi = 0; hi < n_heights; hi++)
            {
                const uint32_t h = heights[hi];
                size_t sm;
                if ((tune_tb >= 16 && (16u << c) != (uint32_t) tune_tb && c != c_max) || (tune_th > 0 && h != (uint32_t) tune_th))
                    continue;
                if (!shape (c, h, sm) || sm > 100 * 1024)
                    continue;
                /* (finer tile heights and the register-limited residency of 6 were tried in this
                 * estimate: no better on cfg 4, worse at 2x and 8x) */
                uint32_t per_sm = (uint32_t) ((220 * 1024) / (sm + 3 * 1024));
                per_sm = per_sm > magb_reg_ctas ? magb_reg_ctas : per_sm;
                if (per_sm < 2)
                    continue;
                const double ctas = (double) ((M.nb_row + M.tile_b - 1) / M.tile_b) * ((L.n_rows + h - 1) / h) * L.n_images;
                const double waves = ctas / ((double) num_sms () * per_sm);
                const double eff = waves / (double) (uint64_t) (waves + 0.999999);
                const double work = 1.0 + 1.6 * M.max_src_rows / h + 8.0 / h * (1024.0 / M.tile_b);
                const double t = work / eff * (per_sm < 4 ? 1.3 : per_sm < 6 ? 1.1 : 1.0);
                if (t < best)
                {
                    best = t;
                    best_c = c;
                    best_h = h;
                }
            }
        }
        if (best_h == 0)
        {
            /* nothing fits comfortably (very wide source windows): shrink the tile until it does */
            best_c = c_max < 6 ? c_max : 6;
            best_h = 32;
            while (best_h > 4 && shape (best_c, best_h, smem) && smem > 56 * 1024)
                best_h /= 2;
        }
    }
    if (!shape (best_c, best_h, smem))
        return cudaErrorInvalidValue;   /* cannot happen: tile_b <= 2048 bytes */
    const uint32_t items = d.bpp_in == 3 && src32 ? (M.u_pitch + 3) / 4 : M.u_pitch;
    M.u_cw = 1;
    M.u_cw_log2 = 0;
    while (M.u_cw < 256 && M.u_cw < items)
    {
        M.u_cw *= 2;
        M.u_cw_log2++;
    }

    static const uint32_t acc_byte[4] = { 1, 5, 3, 7 };
    uint32_t sel = 0;
    for (int j = 0; j < 4; j++)
        sel |= acc_byte[(M.t.prmt_sel >> (4 * j)) & 3] << (4 * j);
    M.acc_prmt_sel = sel;

    dim3 grid ((M.nb_row + M.tile_b - 1) / M.tile_b, (L.n_rows + M.tile_h - 1) / M.tile_h, L.n_images);
    const bool af = d.in_alpha_idx == 0;
    M.prefetch = pdl_first_wave (256, smem + 2304);

    if (d.bpp_in == 3)
    {
        if (d.bpp_out == 3)     return launch_magb_fmt<3, 3, false, false, false> (M, src32, grid, smem, stream);
        if (d.out_unassoc)      return launch_magb_fmt<3, 4, false, true, false> (M, src32, grid, smem, stream);
        return launch_magb_fmt<3, 4, false, false, false> (M, src32, grid, smem, stream);
    }
    if (d.in_unassoc)
    {
        if (d.bpp_out == 3)
            return af ? launch_magb_fmt<4, 3, true, false, true> (M, src32, grid, smem, stream)
                      : launch_magb_fmt<4, 3, true, false, false> (M, src32, grid, smem, stream);
        return af ? launch_magb_fmt<4, 4, true, false, true> (M, src32, grid, smem, stream)
                  : launch_magb_fmt<4, 4, true, false, false> (M, src32, grid, smem, stream);
    }
    if (d.bpp_out == 3)
        return launch_magb_fmt<4, 3, false, false, false> (M, src32, grid, smem, stream);
    if (d.out_unassoc)
        return af ? launch_magb_fmt<4, 4, false, true, true> (M, src32, grid, smem, stream)
                  : launch_magb_fmt<4, 4, false, true, false> (M, src32, grid, smem, stream);
    return launch_magb_fmt<4, 4, false, false, false> (M, src32, grid, smem, stream);
}

/* work items per SM below which the box launcher trades lanes per column for more items (SMOL_BOX_MIN_WARPS) */
static double
box_min_warps ()
{
    static double v = -1.0;
    if (v < 0)
    {
        const char *e = getenv ("SMOL_BOX_MIN_WARPS");
        /* measured on B200 (us per frame at 26 = every resident warp / 16 / 12 / 8): a 57-row band of cfg 3
         * 25.7 / 16.6 / 16.6 / 19.5, 2048^2 -> 128^2 15.3 / 15.3 / 12.9 / 12.9, 4000x3000 -> 200x150 51.4 / 30.8 / 33.6 / 33.6 */
        v = e && atof (e) > 0 ? atof (e) : 12.0;
    }
    return v;
}

static void
box_params_init (BoxParams &P, const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;
    P.d = d;
    P.src = L.src; P.dst = L.dst;
    P.src_pitch = L.src_pitch; P.dst_pitch = L.dst_pitch;
    P.src_image_stride = L.src_image_stride; P.dst_image_stride = L.dst_image_stride;
    P.tab_x = L.tab_x; P.tab_y = L.tab_y; P.luts = L.luts;
    P.first_row = L.first_row; P.n_rows = L.n_rows; P.n_images = L.n_images;
    /* a 24bpp source pixel is fetched with a forced 0xff in byte 3 */
    const uint32_t alpha_idx = d.in_alpha_idx == 0xff ? 3u : d.in_alpha_idx;
    P.alpha_shift = alpha_idx * 8;
    P.col_shift = d.in_col0 * 8;
    P.sel_alpha = 0x4440u | alpha_idx;
    P.sel_c0 = 0x4440u | d.in_col0;
    P.sel_c1 = 0x4440u | (d.in_col0 + 1u);
    P.sel_c2 = 0x4440u | (d.in_col0 + 2u);
    P.sel_ac0 = 0x4400u | (alpha_idx << 4) | d.in_col0;
    P.sel_ac1 = 0x4400u | (alpha_idx << 4) | (d.in_col0 + 1u);
    P.sel_ac2 = 0x4400u | (alpha_idx << 4) | (d.in_col0 + 2u);
    P.unpack_tab = d.in_unassoc ? L.p8l_from_u : L.p8l_from_p;
    P.sel_aaddr = 0x7604u | (alpha_idx << 4);
    P.sel_f0 = 0x7604u | ((uint32_t) d.in_col0 << 4);
    P.sel_f1 = 0x7604u | ((d.in_col0 + 1u) << 4);
    P.sel_f2 = 0x7604u | ((d.in_col0 + 2u) << 4);
    P.warps_lo = 0;
    P.rows_per_item = 1;
    P.n_strips = L.n_rows;
    P.unroll2 = 1;
    P.prefetch = 0;
    {
        static int tma = -1;
        if (tma < 0)
        {
            const char *e = getenv ("SMOL_BOX_TMA");
            tma = e ? atoi (e) != 0 : 0;
        }
        P.use_tma = (uint32_t) tma;
    }
    P.mul8_x = d.span_mul_x << 8;
    P.mul8_y = d.span_mul_y << 8;
    {
        /* largest lane value after unpack x the longest span (+ 2 edge pixels) on either axis */
        const uint64_t lane_max = d.mid == SMOL_MID_P8 ? 255 : d.mid == SMOL_MID_P8L ? 2047
                                  : d.mid == SMOL_MID_P16 ? 0xff80 : 2047 * 255;
        const uint64_t span_x = d.w_in / d.w_out + 2, span_y = d.h_in / d.h_out + 2;
        const uint64_t h_max = lane_max * span_x, v_max = (d.mid == SMOL_MID_P8 && !d.storage128 ? 255 : 65535) * span_y;
        P.acc_fits_24 = h_max < (1u << 24) && v_max < (1u << 24);
    }

}

static cudaError_t
launch_box (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    BoxParams P;

    box_params_init (P, L);

    int mode;
    if (d.mid == SMOL_MID_P8)
        mode = d.in_unassoc ? BM_P8_U : BM_P8_P;
    else if (d.mid == SMOL_MID_P8L)
        mode = d.in_unassoc ? BM_P8L_U : BM_P8L_P;
    else
        mode = d.mid == SMOL_MID_P16 ? BM_P16_U : BM_P16L_U;

    static int tune_g = -1, tune_lut = -1, tune_wpc = 0, tune_unroll = 1;
    if (tune_g < 0)
    {
        const char *e = getenv ("SMOL_BOX_G_LOG2"), *t = getenv ("SMOL_BOX_LUTM"), *w = getenv ("SMOL_BOX_WPC");
        tune_wpc = w ? atoi (w) : 0;
        const char *u = getenv ("SMOL_BOX_UNROLL");
        tune_unroll = u ? atoi (u) : 1;
        tune_g = e ? atoi (e) : 99;
        tune_lut = t ? atoi (t) : 3;
    }
    /* table placement (see box_unpack / box3_accum): modes without gathers need none */
    const bool has_lut = mode == BM_P8L_P || mode == BM_P8L_U || mode == BM_P16L_U;
    /* 3 = the lean row loop (with byte-addressed tables where the mode has any); other values
     * select the older table placements for the linear-light modes and the plain loop elsewhere */
    int lutm = has_lut ? tune_lut : (tune_lut == 3 ? 3 : 0);
    if (lutm == 1 && d.mid != SMOL_MID_P8L)
        lutm = 2;
    if (lutm == 3 && (d.span_mul_x >= (1u << 24) || d.span_mul_y >= (1u << 24)))
        lutm = has_lut ? 2 : 0;         /* the multiply-high normalisation wants span_mul << 8 in 32 bits */
    if (lutm == 3 && has_lut)
    {
        /* the 64 KB-per-table layout must leave room for at least 8 warps' staging buffers
         * (longest segments: the starting G of the loop below) */
        const uint32_t ratio0 = d.w_in / d.w_out;
        uint32_t glog0 = 0;
        while (glog0 < 5 && ratio0 >= (16u << glog0))
            glog0++;
        if (tune_g != 99)
            glog0 = (uint32_t) tune_g;
        const uint64_t seg_px0 = ((uint64_t) (32u >> glog0) * d.w_in + d.w_out - 1) / d.w_out + 3;
        const size_t per_warp0 = 2 * (size_t) ((seg_px0 * d.bpp_in + 32 + 15) & ~(uint64_t) 15);
        const size_t win_hi0 = mode == BM_P8L_P && d.bpp_in != 3 ? 0x30000 : 0x20000;    /* 24bpp: no inverse table (box3_accum) */
        if ((0x10000 - SMOL_BOX3_DYN_WIN_MAX) / per_warp0 + (225 * 1024 - 64 - (win_hi0 - 0x400)) / per_warp0 < 8)
            lutm = 2;
    }

    P.unroll2 = tune_unroll != 0;
    P.prefetch = pdl_first_wave (1024, 0) != 0;
    const bool bi3 = d.bpp_in == 3;     /* 24bpp sources are never unassociated: modes P8_P / P8L_P only */
    /* rows that do not start on 16-byte boundaries: the row-shifted variant of the lean row loop */
    const bool rs = !(aligned16 (L.src) && (L.src_pitch & 15) == 0 && (L.src_image_stride & 15) == 0);
    if (rs && lutm != 3)
        return cudaErrorNotSupported;   /* the caller falls back to the general kernel */
#define BOX_K3(M, B) (rs ? (const void *) smol_box_kernel<M, 3, B, true, false> : g1 ? (const void *) smol_box_kernel<M, 3, B, false, true> \
                         : (const void *) smol_box_kernel<M, 3, B, false, false>)
#define BOX_KERNEL_FOR(M) (lutm == 3 ? BOX_K3 (M, 4) : lutm == 1 ? (const void *) smol_box_kernel<M, 1, 4> : lutm == 2 ? (const void *) smol_box_kernel<M, 2, 4> : (const void *) smol_box_kernel<M, 0, 4>)
#define BOX_KERNEL_FOR3(M) (lutm == 3 ? BOX_K3 (M, 3) : lutm == 1 ? (const void *) smol_box_kernel<M, 1, 3> : lutm == 2 ? (const void *) smol_box_kernel<M, 2, 3> : (const void *) smol_box_kernel<M, 0, 3>)
    /* (picked again once the lanes per column are known: one lane per column has its own instance) */
    const void *fn = nullptr;
    auto pick_fn = [&] (bool g1)
    {
    switch (mode)
    {
        case BM_P8_P:   fn = lutm == 3 ? (bi3 ? BOX_K3 (BM_P8_P, 3) : BOX_K3 (BM_P8_P, 4))
                                       : (bi3 ? (const void *) smol_box_kernel<BM_P8_P, 0, 3> : (const void *) smol_box_kernel<BM_P8_P, 0, 4>); break;
        case BM_P8_U:   fn = lutm == 3 ? BOX_K3 (BM_P8_U, 4) : (const void *) smol_box_kernel<BM_P8_U, 0, 4>; break;
        case BM_P8L_P:  fn = bi3 ? BOX_KERNEL_FOR3 (BM_P8L_P) : BOX_KERNEL_FOR (BM_P8L_P); break;
        case BM_P8L_U:  fn = BOX_KERNEL_FOR (BM_P8L_U); break;
        case BM_P16_U:  fn = lutm == 3 ? BOX_K3 (BM_P16_U, 4) : (const void *) smol_box_kernel<BM_P16_U, 0, 4>; break;
        default:        fn = lutm == 3 ? BOX_K3 (BM_P16L_U, 4)
                             : lutm == 2 ? (const void *) smol_box_kernel<BM_P16L_U, 2, 4> : (const void *) smol_box_kernel<BM_P16L_U, 0, 4>; break;
    }
    };
#undef BOX_KERNEL_FOR
#undef BOX_KERNEL_FOR3
#undef BOX_K3
    pick_fn (false);

    /* Lanes per column (G).  Long spans want several lanes per column (8..16 source pixels per
     * lane per row).  Every extra lane repeats the per-row overhead (edge pixels, normalisation),
     * so beyond that G only grows while there are fewer work items than resident warps. */
    const uint32_t ratio = d.w_in / d.w_out;
    uint32_t glog = 0;
    while (glog < 5 && ratio >= (16u << glog))
        glog++;

    const size_t lut_bytes = lutm == 1 ? 131072
                             : lutm == 3 ? (mode == BM_P8L_P && d.bpp_in != 3 ? 131072 : has_lut ? 65536 : 0)
                             : lutm == 2 ? (mode == BM_P8L_P ? 65536 : 32768) : 0;
    size_t smem = 0;
    uint32_t per_sm = 1, warps_per_cta = lutm == 2 ? 16 : 8;
    for (;; glog++)
    {
        if (tune_g != 99)
            glog = (uint32_t) tune_g;
        const uint32_t cols = 32u >> glog;
        P.lanes_per_col_log2 = glog;
        pick_fn (glog == 0 && lutm == 3 && !rs);
        P.x_tiles = (d.w_out + cols - 1) / cols;
        /* staging buffer: the widest segment an item can need, + alignment slack */
        const uint64_t seg_px = ((uint64_t) cols * d.w_in + d.w_out - 1) / d.w_out + 3;
        P.seg_bytes = (uint32_t) ((seg_px * d.bpp_in + 32 + 15) & ~(uint64_t) 15);
        if (lutm == 1)
        {
            /* one CTA per SM: as many warps as fit beside the tables */
            const size_t room = 220 * 1024 - lut_bytes;
            warps_per_cta = (uint32_t) (room / (2 * (size_t) P.seg_bytes));
            warps_per_cta = warps_per_cta > 32 ? 32 : warps_per_cta < 4 ? 4 : warps_per_cta;
        }
        smem = lut_bytes + (size_t) warps_per_cta * 2 * P.seg_bytes;
        if (lutm == 3)
        {
            /* One CTA per SM.  The tables sit at window addresses 0x10000 (+ 0x20000); staging
             * buffers go below them (the allocation starts a little above the 1 KB the system
             * reserves) and above them, up to the 225 KB opted in to. */
            const size_t per_warp = 2 * (size_t) P.seg_bytes;
            const size_t win_hi = lut_bytes == 131072 ? 0x30000 : 0x20000;
            const size_t dyn_max = 225 * 1024 - 64;
            /* without tables the buffers are simply contiguous ("lo" = all of them) */
            const size_t lo_room = has_lut ? 0x10000 - SMOL_BOX3_DYN_WIN_MAX : dyn_max, hi_room = has_lut ? dyn_max - (win_hi - 0x400) : 0;
            uint32_t lo = (uint32_t) (lo_room / per_warp), hi = (uint32_t) (hi_room / per_warp);
            lo = lo > SMOL_BOX3_MAX_WARPS ? SMOL_BOX3_MAX_WARPS : lo;
            hi = hi > SMOL_BOX3_MAX_WARPS - lo ? SMOL_BOX3_MAX_WARPS - lo : hi;
            /* An item is a column tile x a strip of K consecutive output rows; strips share their
             * boundary source rows, so an item costs about K (R - 1) + 1 source rows (R = rows per
             * output row).  A warp takes ceil (items / warps) items and the SM's issue slots are
             * split between its warps, so the time goes like rounds x rows per item x warps:
             * among the strip lengths and the warp counts that fit (down to 5/8 of the most) take
             * the cheapest, preferring more warps when it is close (latency hiding). */
            {
                const uint32_t w_max = lo + hi;
                const uint32_t x_tiles = (d.w_out + cols - 1) / cols;
                const double R = (double) d.h_in / d.h_out + 1.0;
                static int tune_k = -1;
                if (tune_k < 0)
                {
                    const char *e = getenv ("SMOL_BOX_ROWS_PER_ITEM");
                    tune_k = e ? atoi (e) : 0;
                }
                uint32_t best_w = w_max, best_k = 1;
                double best_cost = 1e300;
                for (uint32_t k = 1; k <= 8 && k <= L.n_rows; k++)
                {
                    if (tune_k > 0 && k != (uint32_t) tune_k && !(k == L.n_rows && (uint32_t) tune_k > L.n_rows))
                        continue;
                    const uint64_t items = (uint64_t) x_tiles * ((L.n_rows + k - 1) / k) * L.n_images;
                    for (uint32_t w = w_max; w >= 8 && w * 8 >= w_max * 5; w--)
                    {
                        const uint64_t slots = (uint64_t) num_sms () * w;
                        const uint64_t rounds = (items + slots - 1) / slots;
                        const double cost = (double) rounds * (k * (R - 1.0) + 1.0) * w * (1.0 + 0.15 * (w_max - w) / w_max);
                        if (cost < best_cost * 0.99)
                        {
                            best_cost = cost;
                            best_w = w;
                            best_k = k;
                        }
                    }
                }
                /* A job of fewer items than one round of warp slots: as few warps per CTA as still give
                 * every item a warp, so that the CTAs spread over ALL the SMs instead of filling some
                 * of them (1,950 items at 17 warps per CTA occupy 115 of 148 SMs; at 14 all of them,
                 * with fewer warps competing for each SM's issue slots).  Table modes only: measured on
                 * B200, linear light 400^2 -> 16^2 25.4 -> 21.0 us, 4000x3000 -> 200x150 33.8 -> 30.8, RGB8
                 * 28.2 -> 26.0, nothing slower; without tables 1024^2 -> 64^2 went 12.6 -> 15.7, so those keep
                 * the cost model's count (profiles/r02_box_spread_sweep.json).  SMOL_BOX_SPREAD=0: off. */
                static int tune_spread = -1;
                if (tune_spread < 0)
                {
                    const char *e = getenv ("SMOL_BOX_SPREAD");
                    tune_spread = e ? atoi (e) : 1;
                }
                if (tune_spread && has_lut)
                {
                    const uint64_t items = (uint64_t) x_tiles * ((L.n_rows + best_k - 1) / best_k) * L.n_images;
                    if (items <= (uint64_t) num_sms () * best_w)
                    {
                        uint32_t w_even = (uint32_t) ((items + num_sms () - 1) / num_sms ());
                        w_even = w_even < 4 ? 4 : w_even;
                        if (w_even < best_w)
                            best_w = w_even;
                    }
                }
                if (tune_wpc > 0 && (uint32_t) tune_wpc <= w_max)
                    best_w = (uint32_t) tune_wpc;
                P.rows_per_item = best_k;
                P.n_strips = (L.n_rows + best_k - 1) / best_k;
                if (best_w < lo)
                    lo = best_w;
                hi = best_w - lo;
            }
            P.warps_lo = lo;
            warps_per_cta = lo + hi;
            smem = has_lut ? (win_hi - 0x400) + hi * per_warp : (size_t) warps_per_cta * per_warp;
        }
        if (smem > 32 * 1024)        /* static shared memory counts against the 48 KB default too */
            smem_optin (fn, 225 * 1024);
        per_sm = (uint32_t) cached_occupancy (fn, (int) warps_per_cta * 32, smem);
        /* More lanes per column only while the job is too small to keep ~8 warps per SM busy (column
         * tiles x output rows; strips are chosen after G and must not make it grow): every extra lane
         * repeats the per-row overhead, so an under-filled GPU at G = 1 still beats a full one at G = 4
         * (a 57-row band of cfg 3, one GPU of eight: 25.7 -> 16.6 us). */
        const double busy = (double) P.x_tiles * L.n_rows * L.n_images / ((double) num_sms () * per_sm * box_min_warps ());
        if (tune_g != 99 || glog >= 5 || busy >= 1.0)
            break;
    }

    const uint64_t n_items = (uint64_t) P.x_tiles * P.n_strips * L.n_images;
    uint64_t blocks = (n_items + warps_per_cta - 1) / warps_per_cta;
    if (blocks > (uint64_t) num_sms () * per_sm)
        blocks = (uint64_t) num_sms () * per_sm;
    dim3 grid ((unsigned) blocks), block (warps_per_cta * 32);

    return launch_pdl_ptr (fn, &P, grid, block, smem, stream);
}

static void taps_params_init (TapsParams &P, const SmolLaunch &L);

/* bilinear / copy / one without halvings on a 128bpp intermediate, word-aligned 32bpp rows: smol_taps0w_kernel */
static bool
taps0w_eligible (const SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;
    static int on = -1;
    if (on < 0)
    {
        const char *e = getenv ("SMOL_TAPS0W");
        on = e ? atoi (e) : 1;
    }
    if (!on || d.h_kind != SMOL_AXIS_TAPS || d.v_kind != SMOL_AXIS_TAPS || !d.storage128 || d.mid == SMOL_MID_P8
        || d.h_halvings != 0 || d.v_halvings != 0)
        return false;
    if (d.bpp_in == 4 && !(aligned4 (L.src) && (L.src_pitch & 3) == 0 && (L.src_image_stride & 3) == 0))
        return false;
    if (d.bpp_out == 4 && !(aligned4 (L.dst) && (L.dst_pitch & 3) == 0 && (L.dst_image_stride & 3) == 0))
        return false;
    return true;
}

static cudaError_t
launch_taps0w (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    Taps0wParams T;

    memset (&T, 0, sizeof (T));
    taps_params_init (T.t, L);
    T.d = d;
    T.luts = L.luts;

    const bool af = d.in_alpha_idx == 0;
    const void *fn;
#define TAPS0W(M, BI, BO, AF) ((const void *) smol_taps0w_kernel<M, BI, BO, AF>)
#define TAPS0W_AF(M, BO) (af ? TAPS0W (M, 4, BO, true) : TAPS0W (M, 4, BO, false))
    if (d.mid == SMOL_MID_P16)
        fn = TAPS0W_AF (BM_P16_U, 4);
    else if (d.mid == SMOL_MID_P16L)
        fn = TAPS0W_AF (BM_P16L_U, 4);
    else if (d.bpp_in == 3)
        fn = d.bpp_out == 3 ? TAPS0W (BM_P8L_P, 3, 3, false) : TAPS0W (BM_P8L_P, 3, 4, false);
    else if (d.in_unassoc)
        fn = d.bpp_out == 3 ? TAPS0W_AF (BM_P8L_U, 3) : TAPS0W_AF (BM_P8L_U, 4);
    else
        fn = d.bpp_out == 3 ? TAPS0W_AF (BM_P8L_P, 3) : TAPS0W_AF (BM_P8L_P, 4);

    const uint32_t px = d.bpp_out == 3 ? 4 : 2;
    const uint64_t x_threads = (d.w_out + px - 1) / px;
    uint32_t bx = 32;
    while (bx < 128 && bx < x_threads)
        bx *= 2;
    /* strips while they leave ~2 resident waves of threads (the wave-fitting choice of the 64bpp strip
     * kernel, pick_rows_per_thread, was measured here too: 4K 1:1 unassociated 31.8 -> 34.7 us) */
    const uint64_t want_threads = (uint64_t) num_sms () * 1536;
    uint32_t rpt = 16;
    while (rpt > 1 && x_threads * ((L.n_rows + rpt - 1) / rpt) * L.n_images < want_threads)
        rpt >>= 1;
    if (d.h_in <= d.h_out && rpt < 4)
        rpt = 4;                        /* magnification: row reuse matters more than thread count */
    {
        static int tune_rpt = -1;
        if (tune_rpt < 0)
        {
            const char *e = getenv ("SMOL_TAPS0W_RPT");
            tune_rpt = e ? atoi (e) : 0;
        }
        if (tune_rpt > 0)
            rpt = (uint32_t) tune_rpt;
    }
    T.t.rows_per_thread = rpt;
    const uint32_t strips = (L.n_rows + rpt - 1) / rpt;
    uint32_t by = 256 / bx;
    if (by > strips)
        by = strips;
    dim3 block (bx, by), grid ((unsigned) ((x_threads + bx - 1) / bx), (strips + by - 1) / by, L.n_images);
    T.prefetch = pdl_first_wave (256, 4096);
    if (T.prefetch)
        T.prefetch = 0xffffffffu;       /* tables are staged first: see launch_half */
    T.row_ahead = 2;
    return launch_pdl_ptr (fn, &T, grid, block, 0, stream);
#undef TAPS0W_AF
#undef TAPS0W
}

static void box_params_init (BoxParams &P, const SmolLaunch &L);

/* The tile kernel for 128bpp bilinear with halvings takes the tiny jobs (see the kernel); cudaErrorNotSupported =
 * not this kernel's job.  SMOL_TILE128H=0 never, 1 whenever possible (measurements, tests). */
static cudaError_t
launch_tile128h (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    static int mode = -1, tune_kb = 72;
    if (mode < 0)
    {
        const char *e = getenv ("SMOL_TILE128H"), *k = getenv ("SMOL_TILE128H_KB");
        tune_kb = k && atoi (k) > 0 ? atoi (k) : 72;
        mode = e ? atoi (e) : 2;
    }
    if (mode == 0 || (d.h_halvings == 0 && d.v_halvings == 0))
        return cudaErrorNotSupported;
    if (d.bpp_out == 4 && !((reinterpret_cast<uintptr_t> (L.dst) & 3) == 0 && (L.dst_pitch & 3) == 0 && (L.dst_image_stride & 3) == 0))
        return cudaErrorNotSupported;
    if (mode == 2 && (uint64_t) d.w_out * L.n_rows * L.n_images > SMOL_TILE128H_MAX_PIXELS)
        return cudaErrorNotSupported;

    Tile128hParams M;
    box_params_init (M.b, L);
    M.hh = d.h_halvings;
    M.vh = d.v_halvings;
    M.src_u32_ok = d.bpp_in == 3 || ((reinterpret_cast<uintptr_t> (L.src) & 3) == 0 && (L.src_pitch & 3) == 0
                                     && (L.src_image_stride & 3) == 0);
    /* Tile shape: the cheapest in unpacks (window pixels per output pixel, the halo is paid again by the
     * neighbours) plus horizontal taps (window rows per output row), among the shapes whose window fits
     * 72 KB and whose vertical table fits sm_ty. */
    double best = 1e300;
    size_t smem = 0;
    M.tile_w = 0;
    for (uint32_t tw = 64; tw >= 8; tw /= 2)
        for (uint32_t th = 32; th >= 1; th /= 2)
        {
            if ((th << M.vh) > 128)
                continue;
            uint64_t cols = ((uint64_t) tw * d.w_in + d.w_out - 1) / d.w_out + 5;     /* + 4 is reached (brute force over the planner) */
            uint64_t rows = ((uint64_t) th * d.h_in + d.h_out - 1) / d.h_out + 5;
            cols = cols < d.w_in ? cols : d.w_in;
            rows = rows < d.h_in ? rows : d.h_in;
            const size_t bytes = (size_t) rows * (cols + tw) * 16;
            if (bytes > (size_t) tune_kb * 1024)
                continue;
            const uint32_t etw = tw < d.w_out ? tw : d.w_out, eth = th < L.n_rows ? th : L.n_rows;
            const double cost = (double) rows * cols / ((double) etw * eth) * 34.0 + (double) rows / eth * (6.0 + 13.0 * (1 << M.hh))
                                + 512.0 / ((double) etw * eth) * 4.0;
            if (cost < best)
            {
                best = cost;
                M.tile_w = tw; M.tile_h = th;
                M.u_pitch = (uint32_t) cols; M.max_src_rows = (uint32_t) rows;
                smem = bytes;
            }
        }
    if (M.tile_w == 0)
        return cudaErrorNotSupported;
    const uint32_t tiles_y = (L.n_rows + M.tile_h - 1) / M.tile_h;
    if (tiles_y > 65535 || L.n_images > 65535)
        return cudaErrorNotSupported;
    dim3 grid ((d.w_out + M.tile_w - 1) / M.tile_w, tiles_y, L.n_images);

#define TILE128H_BO(MD, B, O) (smem_optin ((const void *) smol_tile128h_kernel<MD, B, O>, 100 * 1024), \
                               launch_pdl (smol_tile128h_kernel<MD, B, O>, M, grid, dim3 (512), smem, stream))
#define TILE128H(MD, B) (d.bpp_out == 3 ? TILE128H_BO (MD, B, 3) : TILE128H_BO (MD, B, 4))
    if (d.mid == SMOL_MID_P8L)
    {
        if (d.in_unassoc)
            return TILE128H (BM_P8L_U, 4);
        return d.bpp_in == 3 ? TILE128H (BM_P8L_P, 3) : TILE128H (BM_P8L_P, 4);
    }
    if (d.mid == SMOL_MID_P16)
        return TILE128H (BM_P16_U, 4);
    return TILE128H (BM_P16L_U, 4);
#undef TILE128H
#undef TILE128H_BO
}

static cudaError_t
launch_taps128 (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    BoxParams P;

    if (taps0w_eligible (L))
        return launch_taps0w (L, stream);
    {
        const cudaError_t e = launch_tile128h (L, stream);
        if (e != cudaErrorNotSupported)
            return e;
    }

    box_params_init (P, L);
    const uint32_t src_u32_ok = d.bpp_in == 3 || ((reinterpret_cast<uintptr_t> (L.src) & 3) == 0 && (L.src_pitch & 3) == 0
                                                  && (L.src_image_stride & 3) == 0);
    const uint32_t hh = d.h_halvings, vh = d.v_halvings;
    /* Persistent CTAs: warps walk work items of 32 output pixels.  The warp count per CTA
     * (20..32) is the one that wastes least of the last round of items, given how many CTAs of
     * that size are resident per SM (registers and the up to 64 KB of tables decide). */
    static int tune_rpi = -1;
    if (tune_rpi < 0)
    {
        const char *e = getenv ("SMOL_TAPS128_RPI");
        tune_rpi = e ? atoi (e) : 0;
    }
    /* rows per item: with vertical halvings the row cache rarely carries over (one row); without, strips
     * share half of every output row's source rows with the row above (4K -> 1279x2160 linear light
     * 52.7 -> 43.7 us) -- as long as the strips still make at least two rounds of items for every warp
     * slot of the GPU (a small job is bound by the length of each warp's chain, and strips lengthen it) */
    uint32_t rpi = 1;
    if (vh == 0)
    {
        rpi = (uint64_t) d.h_in * 4 < (uint64_t) d.h_out * 5 ? 8u : 4u;
        while (rpi > 1 && (uint64_t) ((d.w_out + 31) / 32) * ((L.n_rows + rpi - 1) / rpi) * L.n_images
                          < 2 * (uint64_t) num_sms () * SMOL_TAPS128_WARPS (false))
            rpi /= 2;
    }
    if (tune_rpi > 0)
        rpi = (uint32_t) tune_rpi;
    const uint64_t n_items = (uint64_t) ((d.w_out + 31) / 32) * ((L.n_rows + rpi - 1) / rpi) * L.n_images;
    if (n_items > 0x7fffffffull)
        return cudaErrorInvalidValue;

    int variant;
    const void *fn;
    size_t bytes;
    /* dynamic shared memory: up to the end of the tables' fixed window addresses (see box3_accum) */
    const size_t one_tab = 0x20000 - 0x400, two_tabs = 0x30000 - 0x400;
    /* small: one round of items at 16 warps per SM covers the job (see the kernel) */
    const bool small = n_items <= (uint64_t) num_sms () * SMOL_TAPS128_WARPS (true);
    const uint32_t w_max = SMOL_TAPS128_WARPS (small);
#define T128_FN(M, B) (small ? (const void *) smol_taps128_kernel<M, B, true> : (const void *) smol_taps128_kernel<M, B, false>)
    if (d.mid == SMOL_MID_P8L && d.in_unassoc)      { variant = 0; fn = T128_FN (BM_P8L_U, 4); bytes = one_tab; }
    else if (d.mid == SMOL_MID_P8L && d.bpp_in == 3) { variant = 1; fn = T128_FN (BM_P8L_P, 3); bytes = one_tab; }
    else if (d.mid == SMOL_MID_P8L)                 { variant = 2; fn = T128_FN (BM_P8L_P, 4); bytes = two_tabs; }
    else if (d.mid == SMOL_MID_P16)                 { variant = 3; fn = T128_FN (BM_P16_U, 4); bytes = 0; }
    else                                            { variant = 4; fn = T128_FN (BM_P16L_U, 4); bytes = one_tab; }
#undef T128_FN

    /* resident CTAs per SM by device, instance and warps per CTA (0: not asked yet) */
    static int occ_cache[SMOL_KERNELS_MAX_DEVICES][10][33];
    const int dev = current_device (), inst = variant * 2 + (small ? 1 : 0);
    uint32_t best_w = w_max, best_occ = 1;
    smem_optin (fn, 200 * 1024);
    {
        double best_eff = 0.0;
        for (uint32_t w = w_max; w >= w_max * 5 / 8; w--)
        {
            int occ = __atomic_load_n (&occ_cache[dev][inst][w], __ATOMIC_RELAXED);
            if (occ == 0)
            {
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, fn, (int) w * 32, bytes) != cudaSuccess || occ < 1)
                    occ = 1;
                __atomic_store_n (&occ_cache[dev][inst][w], occ, __ATOMIC_RELAXED);
            }
            const uint64_t slots = (uint64_t) num_sms () * occ * w;
            const uint64_t rounds = (n_items + slots - 1) / slots;
            /* efficiency of the last round, scaled by how many warps the SM holds (latency hiding) */
            const double eff = (double) n_items / (double) (rounds * slots) * (0.5 + 0.5 * (double) (occ * w) / 36.0);
            if (eff > best_eff + 0.01)
            {
                best_eff = eff;
                best_w = w;
                best_occ = (uint32_t) occ;
            }
        }
    }
    const uint64_t ctas = (n_items + best_w - 1) / best_w, resident = (uint64_t) num_sms () * best_occ;
    dim3 block (best_w * 32), grid ((unsigned) (ctas < resident ? ctas : resident));

#define T128(M, B) (small ? launch_pdl_args (smol_taps128_kernel<M, B, true>, grid, block, bytes, stream, P, hh, vh, src_u32_ok, rpi) \
                          : launch_pdl_args (smol_taps128_kernel<M, B, false>, grid, block, bytes, stream, P, hh, vh, src_u32_ok, rpi))
    switch (variant)
    {
        case 0:  return T128 (BM_P8L_U, 4);
        case 1:  return T128 (BM_P8L_P, 3);
        case 2:  return T128 (BM_P8L_P, 4);
        case 3:  return T128 (BM_P16_U, 4);
        default: return T128 (BM_P16L_U, 4);
    }
#undef T128
}

static cudaError_t
launch_tile128 (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    Tile128Params M;

    if (taps0w_eligible (L))
        return launch_taps0w (L, stream);

    box_params_init (M.b, L);
    M.src_u32_ok = d.bpp_in == 3 || ((reinterpret_cast<uintptr_t> (L.src) & 3) == 0 && (L.src_pitch & 3) == 0
                                     && (L.src_image_stride & 3) == 0);
    const size_t lut_bytes = 0;
    M.tile_w = 4;
    M.tile_w_log2 = 2;
    while (M.tile_w < 128 && M.tile_w < d.w_out)
    {
        M.tile_w *= 2;
        M.tile_w_log2++;
    }
    M.tile_h = 32;
    size_t smem;
    for (;;)
    {
        const uint64_t cols = ((uint64_t) M.tile_w * d.w_in + d.w_out - 1) / d.w_out + 3;
        const uint64_t rows = ((uint64_t) M.tile_h * d.h_in + d.h_out - 1) / d.h_out + 3;
        M.u_pitch = (uint32_t) (cols < d.w_in ? cols : d.w_in);
        M.max_src_rows = (uint32_t) (rows < d.h_in ? rows : d.h_in);
        smem = lut_bytes + (size_t) M.max_src_rows * ((size_t) M.u_pitch + M.tile_w) * 16;
        if (smem <= 100 * 1024 || M.tile_h <= 2)
            break;
        M.tile_h /= 2;
    }
    M.u_cw = 1;
    M.u_cw_log2 = 0;
    while (M.u_cw < 512 && M.u_cw < M.u_pitch)
    {
        M.u_cw *= 2;
        M.u_cw_log2++;
    }
    dim3 grid ((d.w_out + M.tile_w - 1) / M.tile_w, (L.n_rows + M.tile_h - 1) / M.tile_h, L.n_images);

#define TILE128_BO(MD, B, O) (smem_optin ((const void *) smol_tile128_kernel<MD, B, O>, 200 * 1024), \
                              launch_pdl (smol_tile128_kernel<MD, B, O>, M, grid, dim3 (512), smem, stream))
#define TILE128(MD, B) (d.bpp_out == 3 ? TILE128_BO (MD, B, 3) : TILE128_BO (MD, B, 4))
    if (d.mid == SMOL_MID_P8L)
    {
        if (d.in_unassoc)
            return TILE128 (BM_P8L_U, 4);
        return d.bpp_in == 3 ? TILE128 (BM_P8L_P, 3) : TILE128 (BM_P8L_P, 4);
    }
    if (d.mid == SMOL_MID_P16)
        return TILE128 (BM_P16_U, 4);
    return TILE128 (BM_P16L_U, 4);
#undef TILE128
#undef TILE128_BO
}

template <class OPS, bool HBOX, bool VBOX>
static cudaError_t
launch_rows_k (const RowsParams &P, dim3 grid, size_t smem, cudaStream_t stream)
{
    if (smem > 32 * 1024)
        smem_optin ((const void *) smol_rows_kernel<OPS, HBOX, VBOX>, 200 * 1024);
    return launch_pdl (smol_rows_kernel<OPS, HBOX, VBOX>, P, grid, dim3 (256), smem, stream);
}

/* one of the two mixed filter pairs with compile-time lane arithmetic */
template <int MODE, int BI>
static cudaError_t
launch_rows_lanes (const RowsParams &P, bool hbox, dim3 grid, size_t smem, cudaStream_t stream)
{
    return hbox ? launch_rows_k<RowsLanes<MODE, BI>, true, false> (P, grid, smem, stream)
                : launch_rows_k<RowsLanes<MODE, BI>, false, true> (P, grid, smem, stream);
}

static cudaError_t
launch_rows (const SmolLaunch &L, cudaStream_t stream)
{
    const SmolJobDesc &d = L.d;
    RowsParams P;
    uint32_t glog = 0;

    memset (&P, 0, sizeof (P));
    P.L = L;
    P.seg_bytes = rows_seg_bytes (L, &glog);
    /* few work items (small outputs): more lanes per box column make more of them */
    const uint32_t warps_resident = (uint32_t) num_sms () * 8 * 4;
    while (d.h_kind == SMOL_AXIS_BOX && glog < 5
           && (uint64_t) ((d.w_out + (32u >> glog) - 1) / (32u >> glog)) * L.n_rows * L.n_images < warps_resident
           && d.w_in / d.w_out >= (4u << glog))
        glog++;
    P.lanes_per_col_log2 = glog;
    const uint32_t cols = 32u >> glog;
    P.x_tiles = (d.w_out + cols - 1) / cols;
    {
        const uint64_t seg_px = ((uint64_t) cols * d.w_in + d.w_out - 1) / d.w_out + 4;
        P.seg_bytes = (uint32_t) ((seg_px * d.bpp_in + 32 + 15) & ~(uint64_t) 15);
    }
    /* Strip length: long strips share boundary rows (box) and reuse row pairs (bilinear
     * magnification), short ones make more items; aim at a few items per resident warp. */
    uint64_t k = (uint64_t) P.x_tiles * L.n_rows * L.n_images / ((uint64_t) warps_resident * 2);
    k = k < 1 ? 1 : k > 16 ? 16 : k;
    if (d.v_kind == SMOL_AXIS_TAPS && d.h_out > d.h_in && k < 4)
        k = 4;
    {
        static int tune_k = -1;
        if (tune_k < 0)
        {
            const char *e = getenv ("SMOL_ROWS_PER_ITEM");
            tune_k = e ? atoi (e) : 0;
        }
        if (tune_k > 0)
            k = (uint64_t) tune_k;
    }
    if (k > L.n_rows)
        k = L.n_rows;
    P.rows_per_item = (uint32_t) k;
    P.n_strips = (L.n_rows + P.rows_per_item - 1) / P.rows_per_item;

    const size_t smem = (size_t) 8 * 2 * P.seg_bytes;
    const uint64_t n_items = (uint64_t) P.x_tiles * P.n_strips * L.n_images;
    uint32_t per_sm = (uint32_t) ((200 * 1024) / (smem + sizeof (SmolDeviceLuts) + 1024));
    per_sm = per_sm > 8 ? 8 : per_sm < 1 ? 1 : per_sm;
    uint64_t blocks = (n_items + 7) / 8;
    if (blocks > (uint64_t) num_sms () * per_sm)
        blocks = (uint64_t) num_sms () * per_sm;
    dim3 grid ((unsigned) blocks);
    const bool hb = d.h_kind == SMOL_AXIS_BOX, vb = d.v_kind == SMOL_AXIS_BOX;

    /* Box on one axis and bilinear on the other, an intermediate the lane arithmetic knows (not
     * 8-bit values in 128bpp storage, i.e. > 255:1 without linear light), word-aligned 32bpp rows:
     * compile-time formats. */
    static int lanes_on = -1;
    if (lanes_on < 0)
    {
        const char *e = getenv ("SMOL_ROWS_LANES");
        lanes_on = e ? atoi (e) : 1;
    }
    const bool src_ok = d.bpp_in == 3 || (aligned4 (L.src) && (L.src_pitch & 3) == 0 && (L.src_image_stride & 3) == 0);
    if (lanes_on && hb != vb && src_ok && !(d.mid == SMOL_MID_P8 && d.storage128))
    {
        box_params_init (P.b, L);
        const bool bi3 = d.bpp_in == 3;
        if (d.mid == SMOL_MID_P8)
        {
            if (d.in_unassoc)   return launch_rows_lanes<BM_P8_U, 4> (P, hb, grid, smem, stream);
            return bi3 ? launch_rows_lanes<BM_P8_P, 3> (P, hb, grid, smem, stream) : launch_rows_lanes<BM_P8_P, 4> (P, hb, grid, smem, stream);
        }
        if (d.mid == SMOL_MID_P8L)
        {
            if (d.in_unassoc)   return launch_rows_lanes<BM_P8L_U, 4> (P, hb, grid, smem, stream);
            return bi3 ? launch_rows_lanes<BM_P8L_P, 3> (P, hb, grid, smem, stream) : launch_rows_lanes<BM_P8L_P, 4> (P, hb, grid, smem, stream);
        }
        if (d.mid == SMOL_MID_P16)
            return launch_rows_lanes<BM_P16_U, 4> (P, hb, grid, smem, stream);
        return launch_rows_lanes<BM_P16L_U, 4> (P, hb, grid, smem, stream);
    }

    if (d.storage128)
    {
        if (hb && vb)   return launch_rows_k<RowsSwar<true>, true, true> (P, grid, smem, stream);
        if (hb)         return launch_rows_k<RowsSwar<true>, true, false> (P, grid, smem, stream);
        if (vb)         return launch_rows_k<RowsSwar<true>, false, true> (P, grid, smem, stream);
        return launch_rows_k<RowsSwar<true>, false, false> (P, grid, smem, stream);
    }
    if (hb && vb)       return launch_rows_k<RowsSwar<false>, true, true> (P, grid, smem, stream);
    if (hb)             return launch_rows_k<RowsSwar<false>, true, false> (P, grid, smem, stream);
    if (vb)             return launch_rows_k<RowsSwar<false>, false, true> (P, grid, smem, stream);
    return launch_rows_k<RowsSwar<false>, false, false> (P, grid, smem, stream);
}

template <bool S128, bool HBOX, bool VBOX>
static cudaError_t
launch_general (const SmolLaunch &L, cudaStream_t stream)
{
    const uint32_t TW = SMOL_BLOCK / L.lanes_per_col;
    dim3 grid ((L.d.w_out + TW - 1) / TW, (L.n_rows + L.rows_per_cta - 1) / L.rows_per_cta, L.n_images);
    /* chunk + one extra pixel for the taps + 16 bytes of alignment slack, rounded to 16 */
    size_t smem = ((size_t) (L.chunk_px + 1) * L.d.bpp_in + 16 + 15) & ~(size_t) 15;

    smol_general_kernel<S128, HBOX, VBOX><<<grid, SMOL_BLOCK, smem, stream>>> (L);
    return cudaGetLastError ();
}

/* Launch shape for the general kernel. */
static void
shape_general (SmolLaunch &L)
{
    const SmolJobDesc &d = L.d;
    uint32_t G = 1;

    if (d.h_kind == SMOL_AXIS_BOX)
    {
        /* one thread per ~8..16 source pixels of a span */
        const uint32_t ratio = d.w_in / d.w_out;
        while (G < 32 && ratio >= 16 * G)
            G *= 2;
    }
    L.lanes_per_col = G;
    L.chunk_px = 4096;

    if (d.v_kind == SMOL_AXIS_BOX)
    {
        L.rows_per_cta = 1;
    }
    else
    {
        /* walking several rows per CTA lets the two-row cache work; keep >= ~8 CTAs per SM */
        const uint32_t TW = SMOL_BLOCK / G;
        const uint64_t col_tiles = (d.w_out + TW - 1) / TW;
        uint32_t th = 16;
        while (th > 1 && col_tiles * ((L.n_rows + th - 1) / th) * L.n_images < 148u * 8u)
            th >>= 1;
        L.rows_per_cta = th;
    }
}

extern "C" int
smol_cuda_launch (const SmolLaunch *launch, int kernel_id, void *stream_p, const char **name_out)
{
    cudaStream_t stream = (cudaStream_t) stream_p;
    SmolLaunch L = *launch;
    const bool hb = L.d.h_kind == SMOL_AXIS_BOX, vb = L.d.v_kind == SMOL_AXIS_BOX;
    cudaError_t err;

    if (kernel_id <= SMOL_KERNEL_AUTO || kernel_id >= SMOL_KERNEL_MAX)
        kernel_id = smol_cuda_pick_kernel (launch, SMOL_KERNEL_AUTO);
    if (name_out)
        *name_out = kernel_names[kernel_id];

    if (L.n_rows == 0 || L.n_images == 0)
        return 0;

    if (kernel_id == SMOL_KERNEL_HALF2X && half_eligible (L))
        return (int) launch_half (L, stream);
    if (kernel_id == SMOL_KERNEL_BOX && box_eligible (L))
    {
        const cudaError_t e = launch_box (L, stream);
        if (e != cudaErrorNotSupported)
            return (int) e;
    }
    if (kernel_id == SMOL_KERNEL_TAPS128 && taps128_eligible (L))
        return (int) launch_taps128 (L, stream);
    if (kernel_id == SMOL_KERNEL_TILE128 && tile128_eligible (L))
        return (int) launch_tile128 (L, stream);
    if (kernel_id == SMOL_KERNEL_MAGB && magb_eligible (L))
        return (int) launch_magb (L, stream);
    if (kernel_id == SMOL_KERNEL_MAG && mag_eligible (L))
        return (int) launch_mag (L, stream);
    if (kernel_id == SMOL_KERNEL_TAPS_DIRECT && taps_eligible (L))
        return (int) launch_taps (L, stream);
    /* the box kernel's own fallback (rows off 16-byte boundaries with an older table placement) lands here too */
    if ((kernel_id == SMOL_KERNEL_ROWS || kernel_id == SMOL_KERNEL_BOX) && rows_eligible (L))
    {
        if (name_out)
            *name_out = kernel_names[SMOL_KERNEL_ROWS];
        return (int) launch_rows (L, stream);
    }
    if (name_out)
        *name_out = kernel_names[SMOL_KERNEL_GENERAL];

    shape_general (L);

    if (L.d.storage128)
    {
        if (hb && vb)       err = launch_general<true, true, true> (L, stream);
        else if (hb)        err = launch_general<true, true, false> (L, stream);
        else if (vb)        err = launch_general<true, false, true> (L, stream);
        else                err = launch_general<true, false, false> (L, stream);
    }
    else
    {
        if (hb && vb)       err = launch_general<false, true, true> (L, stream);
        else if (hb)        err = launch_general<false, true, false> (L, stream);
        else if (vb)        err = launch_general<false, false, true> (L, stream);
        else                err = launch_general<false, false, false> (L, stream);
    }
    return (int) err;
}

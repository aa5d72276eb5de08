"""Partitioning helpers for multi-GPU runs (one process per GPU, no data-path collective).

The pipeline has no reduction or exchange (reference smolscale.h:70-74: disjoint row batches on
a shared const context are independent), so multi-GPU work is pure partitioning:

* ``image_shard``  -- thumbnail batches / frame queues: rank r takes a contiguous slice of images.
* ``row_band``     -- one large image: rank r renders output rows [first, first + n) and needs only
                      the source rows ``ScaleCtx.band_source_rows(first, n)`` reports (band + halo).
"""


def image_shard(n_images, rank, world):
    """Contiguous, balanced slice of n_images for `rank`: returns (first, count)."""
    base, extra = divmod(n_images, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def row_band(height_out, rank, world):
    """Output row band of `rank`, ceil-sized like the reference's threaded caller (test.c:863-877)."""
    per = (height_out + world - 1) // world
    first = min(rank * per, height_out)
    return first, min(per, height_out - first)

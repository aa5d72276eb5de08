/* smol_generate.c -- PNG in, a range of scaled PNGs out: the reference's `test smol generate`
 * mode (test.c:1303-1371, :1496-1600) driven through the B200 library.
 *
 *   smol_generate [--srgb] [--reference-types] [--host] <min_scale> <max_scale> <n_steps> <file.png>
 *
 * writes <file>-WWWW-HHHH.png for n_steps sizes between min_scale and max_scale times the input
 * size, stepping exactly like run_generate (float step sizes, truncation to whole pixels, clamp
 * to 1..65535).
 *
 * The source image is uploaded to the GPU once and stays resident in HBM for all the sizes
 * (smol_scale_simple with a device source and a host destination: the library runs the kernel on
 * the caller's stream and brings the rows back); --host hands the library the host buffer every
 * time instead (its banded H2D / kernel / D2H pipeline), which is what an unmodified caller does.
 *
 * Pixel types: a PNG holds unassociated R,G,B,A bytes, so the default is RGBA8_UNASSOCIATED on
 * both sides.  --reference-types reproduces the reference program, which passes the same bytes
 * as PIXEL_TYPE_SMOL = ARGB8_PREMULTIPLIED (test.c:20, premultiplication compiled out at :1326).
 *
 * Build: gcc -O2 -o tools/smol_generate tools/smol_generate.c -Iinclude -I/usr/local/cuda/include \
 *            -Lsmolscale_b200 -lsmolscale_cuda -lsmolpng -L/usr/local/cuda/lib64 -lcudart_static \
 *            -lpthread -lrt -ldl -Wl,-rpath,'$ORIGIN/../smolscale_b200' */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smolscale.h"
#include "smol-png.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf (stderr, "%s: %s\n", #x, cudaGetErrorString (e_)); exit (1); } } while (0)
#define CLAMP(x, lo, hi) ((x) < (lo) ? (lo) : (x) > (hi) ? (hi) : (x))

static void
usage (void)
{
    fprintf (stderr, "Usage: smol_generate [--srgb] [--reference-types] [--host]\n"
                     "                     <min_scale> <max_scale> <n_steps> <filename>\n");
    exit (1);
}

int
main (int argc, char *argv [])
{
    SmolPixelType ptype = SMOL_PIXEL_RGBA8_UNASSOCIATED;
    int with_srgb = 0, host_source = 0, a = 1;
    double scale_min, scale_max;
    unsigned int n_steps, in_width, in_height, step;
    unsigned int out_width_min, out_width_max, out_height_min, out_height_max;
    float width_step_size, height_step_size;
    const char *filename;
    char *prefix, *dot;
    void *raw_data, *src;
    uint32_t *out_data;

    for (; a < argc && !strncmp (argv [a], "--", 2); a++)
    {
        if (!strcmp (argv [a], "--srgb"))
            with_srgb = 1;
        else if (!strcmp (argv [a], "--reference-types"))
            ptype = SMOL_PIXEL_ARGB8_PREMULTIPLIED;
        else if (!strcmp (argv [a], "--host"))
            host_source = 1;
        else
            usage ();
    }
    if (argc - a != 4)
        usage ();
    scale_min = strtod (argv [a], NULL);
    scale_max = strtod (argv [a + 1], NULL);
    n_steps = (unsigned int) strtoul (argv [a + 2], NULL, 10);
    filename = argv [a + 3];

    if (!smoltest_load_image (filename, &in_width, &in_height, &raw_data))
    {
        fprintf (stderr, "Failed to read image '%s'.\n", filename);
        return 1;
    }

    /* test.c:1330-1344 */
    out_width_min = (unsigned int) CLAMP (in_width * scale_min, 1, 65535);
    out_width_max = (unsigned int) CLAMP (in_width * scale_max, 1, 65535);
    out_height_min = (unsigned int) CLAMP (in_height * scale_min, 1, 65535);
    out_height_max = (unsigned int) CLAMP (in_height * scale_max, 1, 65535);
    if (n_steps > 1)
    {
        width_step_size = (out_width_max - out_width_min) / ((float) n_steps - 1.0);
        height_step_size = (out_height_max - out_height_min) / ((float) n_steps - 1.0);
    }
    else
    {
        width_step_size = 99999.0;
        height_step_size = 99999.0;
    }

    prefix = strdup (filename);
    dot = strrchr (prefix, '.');
    if (dot)
        *dot = '\0';

    if (host_source)
        src = raw_data;
    else
    {
        CK (cudaMalloc (&src, (size_t) in_width * in_height * 4));
        CK (cudaMemcpy (src, raw_data, (size_t) in_width * in_height * 4, cudaMemcpyHostToDevice));
    }
    out_data = (uint32_t *) malloc ((size_t) CLAMP (out_width_max, out_width_min, 65535)
                                    * CLAMP (out_height_max, out_height_min, 65535) * 4);
    if (!out_data)
    {
        fprintf (stderr, "out of memory\n");
        return 1;
    }

    for (step = 0; step < n_steps; step++)
    {
        unsigned int out_width = (unsigned int) CLAMP (out_width_min + step * width_step_size, 1, 65535);
        unsigned int out_height = (unsigned int) CLAMP (out_height_min + step * height_step_size, 1, 65535);

        smol_scale_simple (src, ptype, in_width, in_height, in_width * 4,
                           out_data, ptype, out_width, out_height, out_width * 4, with_srgb);
        smoltest_save_image (prefix, out_data, out_width, out_height);
        fputc ('*', stderr);
        fflush (stderr);
    }
    fputc ('\n', stderr);

    if (!host_source)
        CK (cudaFree (src));
    free (out_data);
    free (raw_data);
    free (prefix);
    return 0;
}

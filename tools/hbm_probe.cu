/* hbm_probe.cu -- measured HBM ceilings for the traffic mixes of the five BASELINE configurations.
 *
 * MEASURED_PEAKS.json's hbm_gbs is a device-to-device copy (1 byte read per byte written).  The scaling
 * jobs are not copies: cfg 5 reads 64 bytes per byte written, cfg 2 reads 4, cfg 4 WRITES 16 per byte read.
 * These hand-written probes (128-bit ld.global.nc / st.global, grid = SMs x resident CTAs, buffers far
 * larger than L2) give the ceiling of each mix, plus cudaMemsetAsync and a TMA bulk store
 * (cp.async.bulk.global.shared::cta) as independent write-only figures.
 *
 * Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hbm_probe tools/hbm_probe.cu
 * Output: one JSON object (GB/s = bytes read + bytes written per second, best and median of REPS). */
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf (stderr, "%s: %s\n", #x, cudaGetErrorString (e_)); exit (1); } } while (0)

__device__ __forceinline__ uint4 ldg_nc (const uint4 *p)
{
    uint4 v;
    asm volatile ("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

/* reads R 16-byte words per word written; W words written per word read when R == 0 is handled by write_mix */
template <int R>
__global__ void __launch_bounds__ (256) read_mix (const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n_out)
{
    /* output word i consumes input words [i * R, (i + 1) * R): a warp reads R contiguous 512-byte runs */
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (size_t) gridDim.x * blockDim.x)
    {
        const size_t warp_base = (i & ~(size_t) 31) * R + (i & 31);
        uint4 acc = make_uint4 (0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < R; r++)
        {
            const uint4 v = ldg_nc (in + warp_base + (size_t) r * 32);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        out[i] = acc;
    }
}

__global__ void __launch_bounds__ (256) read_only (const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n_in)
{
    uint4 acc = make_uint4 (0, 0, 0, 0);
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_in; i += 4 * stride)
    {
        const uint4 a = ldg_nc (in + i), b = ldg_nc (in + i + stride), c = ldg_nc (in + i + 2 * stride), d = ldg_nc (in + i + 3 * stride);
        acc.x ^= a.x ^ b.x ^ c.x ^ d.x; acc.y ^= a.y ^ b.y ^ c.y ^ d.y;
        acc.z ^= a.z ^ b.z ^ c.z ^ d.z; acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
    }
    for (; i < n_in; i += stride)
    {
        const uint4 a = ldg_nc (in + i);
        acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345679u)     /* never true for the probe's data: defeats dead-code removal */
        out[threadIdx.x] = acc;
}

/* writes W words per word read (W = 0: write-only) */
template <int W>
__global__ void __launch_bounds__ (256) write_mix (const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n_out)
{
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (size_t) gridDim.x * blockDim.x)
    {
        uint4 v = make_uint4 ((uint32_t) i, 1, 2, 3);
        if constexpr (W > 0)
            v = ldg_nc (in + i / W);
        out[i] = v;
    }
}

/* write-only through the TMA unit: fill a shared tile once, then bulk-store it over and over */
__global__ void __launch_bounds__ (128) tma_store (uint8_t *out, size_t n_tiles, uint32_t tile_bytes)
{
    extern __shared__ __align__ (128) uint8_t tile[];
    for (uint32_t i = threadIdx.x; i < tile_bytes / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *> (tile)[i] = i;
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads ();
    if (threadIdx.x == 0)
    {
        const uint32_t saddr = (uint32_t) __cvta_generic_to_shared (tile);
        for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x)
        {
            asm volatile ("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + t * tile_bytes), "r"(saddr), "r"(tile_bytes) : "memory");
            asm volatile ("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile ("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile ("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
static void timeit (const char *name, double bytes, int reps, F f, bool last = false)
{
    cudaEvent_t e0, e1;
    std::vector<float> ms;
    CK (cudaEventCreate (&e0)); CK (cudaEventCreate (&e1));
    f (); f ();
    CK (cudaDeviceSynchronize ());
    for (int r = 0; r < reps; r++)
    {
        float t;
        CK (cudaEventRecord (e0)); f (); CK (cudaEventRecord (e1)); CK (cudaEventSynchronize (e1));
        CK (cudaEventElapsedTime (&t, e0, e1));
        ms.push_back (t);
    }
    CK (cudaGetLastError ());
    std::sort (ms.begin (), ms.end ());
    /* sustained: back to back for about half a second (power / clock limits show up here, not in a burst) */
    int n = (int) (500.0f / ms[ms.size () / 2]) + 1;
    float t;
    CK (cudaEventRecord (e0));
    for (int r = 0; r < n; r++)
        f ();
    CK (cudaEventRecord (e1)); CK (cudaEventSynchronize (e1));
    CK (cudaEventElapsedTime (&t, e0, e1));
    printf ("  \"%s\": {\"best_gbs\": %.1f, \"median_gbs\": %.1f, \"sustained_gbs\": %.1f, \"sustained_ms\": %.0f}%s\n", name,
            bytes / ms[0] / 1e6, bytes / ms[ms.size () / 2] / 1e6, bytes * n / t / 1e6, t, last ? "" : ",");
    fflush (stdout);
}

int main (int argc, char **argv)
{
    const size_t big = (size_t) (argc > 1 ? atoi (argv[1]) : 2048) << 20;       /* bytes of the larger side */
    const int reps = argc > 2 ? atoi (argv[2]) : 15;
    uint4 *a, *b;
    int sms = 148;
    CK (cudaMalloc (&a, big)); CK (cudaMalloc (&b, big));
    CK (cudaMemset (a, 0x5a, big)); CK (cudaMemset (b, 0xa5, big));
    CK (cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 8;
    const size_t words = big / 16;

    printf ("{\"buffer_mb\": %zu, \"sms\": %d, \"unit\": \"GB/s, read + written bytes\",\n", big >> 20, sms);
    timeit ("copy_1r_1w_ldst128", 2.0 * big, reps, [&] { write_mix<1><<<grid, 256>>> (a, b, words); });
    timeit ("copy_cudaMemcpyAsync", 2.0 * big, reps, [&] { CK (cudaMemcpyAsync (b, a, big, cudaMemcpyDeviceToDevice)); });
    timeit ("read_only", 1.0 * big, reps, [&] { read_only<<<grid, 256>>> (a, b, words); });
    timeit ("read_64_write_1 (cfg5 mix)", big * (1.0 + 1.0 / 64), reps, [&] { read_mix<64><<<grid, 256>>> (a, b, words / 64); });
    timeit ("read_16_write_1", big * (1.0 + 1.0 / 16), reps, [&] { read_mix<16><<<grid, 256>>> (a, b, words / 16); });
    timeit ("read_4_write_1 (cfg1, cfg2 mix)", big * 1.25, reps, [&] { read_mix<4><<<grid, 256>>> (a, b, words / 4); });
    timeit ("write_only_st128", 1.0 * big, reps, [&] { write_mix<0><<<grid, 256>>> (a, b, words); });
    timeit ("write_only_cudaMemsetAsync", 1.0 * big, reps, [&] { CK (cudaMemsetAsync (b, 7, big)); });
    for (uint32_t tile = 4096; tile <= 32768; tile *= 2)
    {
        char name[64];
        snprintf (name, sizeof (name), "write_only_tma_bulk_store_%uk_tiles", tile / 1024);
        CK (cudaFuncSetAttribute (tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        timeit (name, 1.0 * big, reps, [&] { tma_store<<<sms * 4, 128, tile>>> ((uint8_t *) b, big / tile, tile); });
    }
    timeit ("read_1_write_4", big * 1.25, reps, [&] { write_mix<4><<<grid, 256>>> (a, b, words); });
    timeit ("read_1_write_16 (cfg4 mix)", big * (1.0 + 1.0 / 16), reps, [&] { write_mix<16><<<grid, 256>>> (a, b, words); }, true);
    printf ("}\n");
    return 0;
}

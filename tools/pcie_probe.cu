/* pcie_probe.cu -- what the host <-> device path of ONE box can carry when 1, 2, 4, 8 GPUs copy at once.
 *
 * The end-to-end leg of bench.py (host buffers, copies inside the timed region) is bound by this path,
 * not by the kernels.  The probe drives one host thread + one pinned buffer pair per GPU and reports the
 * aggregate H2D / D2H / both-ways rate for every GPU count, twice: with the buffers allocated wherever
 * the calling thread happens to run ("plain"), and with thread + buffer bound to the NUMA node the GPU
 * hangs off ("numa": sched_setaffinity + set_mempolicy before cudaHostAlloc).  It also times pageable
 * (malloc) sources through cudaMemcpyAsync, the path an unmodified reference caller takes.
 *
 * Build: nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu -lpthread     Output: one JSON object. */
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf (stderr, "%s: %s\n", #x, cudaGetErrorString (e_)); exit (1); } } while (0)

static double now_s (void)
{
    struct timespec ts;
    clock_gettime (CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static int gpu_numa_node (int dev)
{
    char bus[64], path[256];
    int node = -1;
    if (cudaDeviceGetPCIBusId (bus, sizeof (bus), dev) != cudaSuccess)
        return -1;
    for (char *p = bus; *p; p++)
        if (*p >= 'A' && *p <= 'Z')
            *p += 'a' - 'A';
    snprintf (path, sizeof (path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen (path, "r");
    if (!f)
        return -1;
    if (fscanf (f, "%d", &node) != 1)
        node = -1;
    fclose (f);
    return node;
}

/* cpus of a NUMA node, intersected with what this process may use */
static int node_cpuset (int node, cpu_set_t *out)
{
    char path[128], buf[4096];
    cpu_set_t allowed;
    int n = 0;
    CPU_ZERO (out);
    if (node < 0 || sched_getaffinity (0, sizeof (allowed), &allowed) != 0)
        return 0;
    snprintf (path, sizeof (path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen (path, "r");
    if (!f)
        return 0;
    if (!fgets (buf, sizeof (buf), f))
        buf[0] = 0;
    fclose (f);
    for (char *tok = strtok (buf, ",\n"); tok; tok = strtok (NULL, ",\n"))
    {
        int a, b;
        if (sscanf (tok, "%d-%d", &a, &b) == 2) ;
        else if (sscanf (tok, "%d", &a) == 1) b = a;
        else continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++)
            if (CPU_ISSET (c, &allowed))
            {
                CPU_SET (c, out);
                n++;
            }
    }
    return n;
}

static long set_mempolicy_node (int node)
{
    /* MPOL_PREFERRED = 1, MPOL_DEFAULT = 0 */
    if (node < 0)
        return syscall (SYS_set_mempolicy, 0, NULL, 0);
    unsigned long mask[16];
    memset (mask, 0, sizeof (mask));
    mask[node / (8 * sizeof (unsigned long))] |= 1ul << (node % (8 * sizeof (unsigned long)));
    return syscall (SYS_set_mempolicy, 1, mask, sizeof (mask) * 8);
}

typedef struct
{
    int dev, numa, bind, mode;      /* mode 0: h2d, 1: d2h, 2: both, 3: pageable h2d, 4: 4 bytes up per byte down (BASELINE cfg 2's mix) */
    size_t bytes;
    int reps;
    pthread_barrier_t *bar;
    double seconds;
    int bound_cpus;
    long mempolicy_rc;
}
Job;

static void *worker (void *arg)
{
    Job *j = (Job *) arg;
    void *h_up = NULL, *h_down = NULL, *d_a = NULL, *d_b = NULL, *pg = NULL;
    cudaStream_t s0, s1;
    cpu_set_t set, old;

    CK (cudaSetDevice (j->dev));
    sched_getaffinity (0, sizeof (old), &old);
    j->bound_cpus = 0;
    j->mempolicy_rc = 0;
    if (j->bind && j->numa >= 0)
    {
        j->bound_cpus = node_cpuset (j->numa, &set);
        if (j->bound_cpus > 0)
            sched_setaffinity (0, sizeof (set), &set);
        j->mempolicy_rc = set_mempolicy_node (j->numa);
    }
    CK (cudaHostAlloc (&h_up, j->bytes, cudaHostAllocDefault));
    CK (cudaHostAlloc (&h_down, j->bytes, cudaHostAllocDefault));
    memset (h_up, 1, j->bytes);
    memset (h_down, 2, j->bytes);
    if (j->mode == 3)
    {
        pg = malloc (j->bytes);
        memset (pg, 3, j->bytes);
    }
    if (j->bind && j->numa >= 0)
        set_mempolicy_node (-1);
    CK (cudaMalloc (&d_a, j->bytes));
    CK (cudaMalloc (&d_b, j->bytes));
    CK (cudaStreamCreateWithFlags (&s0, cudaStreamNonBlocking));
    CK (cudaStreamCreateWithFlags (&s1, cudaStreamNonBlocking));
    CK (cudaMemcpyAsync (d_a, h_up, j->bytes, cudaMemcpyHostToDevice, s0));
    CK (cudaMemcpyAsync (h_down, d_b, j->bytes, cudaMemcpyDeviceToHost, s1));
    CK (cudaStreamSynchronize (s0));
    CK (cudaStreamSynchronize (s1));

    pthread_barrier_wait (j->bar);
    double t0 = now_s ();
    for (int r = 0; r < j->reps; r++)
    {
        if (j->mode == 0 || j->mode == 2)
            CK (cudaMemcpyAsync (d_a, h_up, j->bytes, cudaMemcpyHostToDevice, s0));
        if (j->mode == 1 || j->mode == 2)
            CK (cudaMemcpyAsync (h_down, d_b, j->bytes, cudaMemcpyDeviceToHost, s1));
        if (j->mode == 3)
            CK (cudaMemcpyAsync (d_a, pg, j->bytes, cudaMemcpyHostToDevice, s0));
        if (j->mode == 4)
        {
            CK (cudaMemcpyAsync (d_a, h_up, j->bytes, cudaMemcpyHostToDevice, s0));
            CK (cudaMemcpyAsync (h_down, d_b, j->bytes / 4, cudaMemcpyDeviceToHost, s1));
        }
    }
    CK (cudaStreamSynchronize (s0));
    CK (cudaStreamSynchronize (s1));
    j->seconds = now_s () - t0;
    pthread_barrier_wait (j->bar);

    sched_setaffinity (0, sizeof (old), &old);
    cudaFreeHost (h_up); cudaFreeHost (h_down); cudaFree (d_a); cudaFree (d_b); free (pg);
    cudaStreamDestroy (s0); cudaStreamDestroy (s1);
    return NULL;
}

int main (int argc, char **argv)
{
    int n_dev = 0;
    const size_t bytes = (size_t) (argc > 1 ? atoi (argv[1]) : 256) << 20;
    const int reps = argc > 2 ? atoi (argv[2]) : 8;
    static const char *mode_name[] = { "h2d", "d2h", "both", "pageable_h2d", "mix_4up_1down" };

    CK (cudaGetDeviceCount (&n_dev));
    printf ("{\"gpus_visible\": %d, \"host_cpus\": %ld, \"buffer_mb\": %zu, \"reps\": %d, \"gpu_numa\": [",
            n_dev, sysconf (_SC_NPROCESSORS_ONLN), bytes >> 20, reps);
    int numa[64];
    for (int d = 0; d < n_dev && d < 64; d++)
    {
        numa[d] = gpu_numa_node (d);
        printf ("%s%d", d ? ", " : "", numa[d]);
    }
    printf ("], \"runs\": [");
    int first = 1;
    for (int n = 1; n <= n_dev; n *= 2)
        for (int bind = 0; bind <= 1; bind++)
            for (int mode = 0; mode < 5; mode++)
            {
                if (mode == 3 && bind)
                    continue;
                pthread_t th[64];
                Job jobs[64];
                pthread_barrier_t bar;
                pthread_barrier_init (&bar, NULL, n);
                for (int i = 0; i < n; i++)
                {
                    jobs[i].dev = i; jobs[i].numa = numa[i]; jobs[i].bind = bind; jobs[i].mode = mode;
                    jobs[i].bytes = bytes; jobs[i].reps = mode == 3 ? (reps + 3) / 4 : reps; jobs[i].bar = &bar;
                    pthread_create (&th[i], NULL, worker, &jobs[i]);
                }
                double worst = 0;
                for (int i = 0; i < n; i++)
                {
                    pthread_join (th[i], NULL);
                    if (jobs[i].seconds > worst)
                        worst = jobs[i].seconds;
                }
                pthread_barrier_destroy (&bar);
                const double dirs = mode == 2 ? 2.0 : mode == 4 ? 1.25 : 1.0;
                printf ("%s\n  {\"gpus\": %d, \"placement\": \"%s\", \"mode\": \"%s\", \"aggregate_gbs\": %.1f, \"per_gpu_gbs\": %.1f, "
                        "\"bound_cpus\": %d, \"mempolicy_rc\": %ld}",
                        first ? "" : ",", n, bind ? "numa" : "plain", mode_name[mode],
                        dirs * n * (double) bytes * jobs[0].reps / worst / 1e9,
                        dirs * (double) bytes * jobs[0].reps / worst / 1e9, jobs[0].bound_cpus, jobs[0].mempolicy_rc);
                first = 0;
                fflush (stdout);
            }
    printf ("\n]}\n");
    return 0;
}

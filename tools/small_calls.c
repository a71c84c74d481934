/* The reference's real call pattern, timed through the C ABI from plain C (no Python in the loop):
 * the binaries hand PsdCascade::process() one 4096-sample pageable Vec<f32> per Source::get()
 * (reference src/source.rs:116,123; src/bin/psd.rs:170-183).  This driver feeds `total` samples in calls of
 * `block` samples from PAGEABLE host memory into sspsd_cascade_process_f32, ends with one psd() readout,
 * and prints one JSON line: MS/s end to end (staging memcpy + H2D + kernels + readout inside the timed region).
 *
 *   small_calls [n_fft=4096] [block=4096] [total=400000000] [host_stage=0]
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sspsd.h"

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char **argv)
{
    uint32_t n_fft = argc > 1 ? (uint32_t)atoi(argv[1]) : 4096;
    size_t block = argc > 2 ? (size_t)atoll(argv[2]) : 4096;
    size_t total = argc > 3 ? (size_t)atoll(argv[3]) : 400000000;
    uint64_t host_stage = argc > 4 ? (uint64_t)atoll(argv[4]) : 0;
    sspsd_config cfg;
    sspsd_config_default(n_fft, &cfg);
    cfg.host_stage = host_stage;
    sspsd_cascade *c = NULL;
    int rc = sspsd_cascade_create(&cfg, &c);
    if (rc) {
        fprintf(stderr, "create: %d %s\n", rc, sspsd_last_error());
        return rc == SSPSD_ECUDA ? 3 : 1;
    }
    /* a pool of distinct pageable blocks (64 MiB, so the source is not served from the CPU's L2) */
    const size_t pool = ((size_t)1 << 24) / block * block;
    float *x = (float *)malloc(pool * sizeof(float));
    uint64_t s = 0x7654321;
    for (size_t i = 0; i < pool; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        x[i] = ((float)((s >> 40) & 0xffffff) / 16777216.0f - 0.5f) * 3.4641016f;
    }
    float *p = (float *)malloc(sizeof(float) * SSPSD_MAX_STAGES * (n_fft / 2 + 1));
    sspsd_break b[SSPSD_MAX_STAGES];
    /* warm-up: allocations, first launches */
    for (size_t pos = 0; pos < ((size_t)1 << 25); pos += block) sspsd_cascade_process_f32(c, x + pos % pool, block, SSPSD_MEM_HOST);
    size_t pl = SSPSD_MAX_STAGES * (n_fft / 2 + 1), bl = SSPSD_MAX_STAGES;
    sspsd_cascade_psd(c, NULL, p, &pl, b, &bl);
    sspsd_cascade_reset(c);
    sspsd_cascade_sync(c);
    const double t0 = now();
    size_t calls = 0;
    for (size_t pos = 0; pos < total; pos += block, ++calls) {
        const size_t n = total - pos < block ? total - pos : block;
        rc = sspsd_cascade_process_f32(c, x + pos % pool, n, SSPSD_MEM_HOST);
        if (rc) {
            fprintf(stderr, "process: %d %s\n", rc, sspsd_last_error());
            return 1;
        }
    }
    pl = SSPSD_MAX_STAGES * (n_fft / 2 + 1);
    bl = SSPSD_MAX_STAGES;
    rc = sspsd_cascade_psd(c, NULL, p, &pl, b, &bl);
    const double dt = now() - t0;
    if (rc) {
        fprintf(stderr, "psd: %d %s\n", rc, sspsd_last_error());
        return 1;
    }
    printf("{\"value\": %.3f, \"unit\": \"MS/s\", \"n_fft\": %u, \"block\": %zu, \"calls\": %zu, \"samples\": %zu, "
           "\"seconds\": %.4f, \"us_per_call\": %.4f, \"stage0_segments\": %u, \"stages\": %zu, "
           "\"memory\": \"pageable host, %zu-sample slices (reference src/source.rs:116)\"}\n",
           (double)total / dt / 1e6, n_fft, block, calls, total, dt, dt / (double)calls * 1e6, b[bl - 1].count, bl, block);
    sspsd_cascade_destroy(c);
    free(x);
    free(p);
    return 0;
}

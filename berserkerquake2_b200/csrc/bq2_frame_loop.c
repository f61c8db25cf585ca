/*
 * bq2_frame_loop.c -- libbq2loop.so: an engine's frame loop in C, on top of the PUBLIC C ABI only.
 *
 * bench.py's end-to-end leg measures frames through bq2_world_submit / bq2_frame_wait.  Driven from Python every call
 * costs a microsecond or two of interpreter and ctypes overhead that an engine (C, like the reference) does not pay;
 * this loop makes the same calls the way the engine's renderer would -- `depth` frames in flight, one submit and one
 * wait per iteration -- and times the steady state with clock_gettime.  Nothing here is part of the product ABI.
 */
#define _GNU_SOURCE
#include "bq2_gpu.h"
#include <time.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* k submit+wait iterations with `depth` frames in flight (the library refuses more than it has frame states for); the clock runs over the k iterations (k frames in,
 * k frames out), the frames that fill and drain the pipeline are outside it.  Returns 0 or the first error status;
 * *seconds = wall clock of the k iterations, *total_indices = sum of the k waited frames' index counts. */
int bq2_frame_loop(bq2_ctx *ctx, const bq2_world_desc *wd, int k, int depth, double *seconds, uint64_t *total_indices)
{
    bq2_frame_out fo;
    bq2_status s;
    uint64_t tot = 0;
    double t0;
    int i;
    if (!ctx || !wd || k < 0 || depth < 1 || depth > 16 || !seconds || !total_indices) return BQ2_E_INVALID;
    for (i = 0; i < depth - 1; i++)
        if ((s = bq2_world_submit(ctx, wd)) != BQ2_OK) return s;
    t0 = now_s();
    for (i = 0; i < k; i++) {
        if ((s = bq2_world_submit(ctx, wd)) != BQ2_OK) return s;
        if ((s = bq2_frame_wait(ctx, &fo)) != BQ2_OK) return s;
        tot += fo.total_indices;
    }
    *seconds = now_s() - t0;
    for (i = 0; i < depth - 1; i++)
        if ((s = bq2_frame_wait(ctx, &fo)) != BQ2_OK) return s;
    *total_indices = tot;
    return BQ2_OK;
}

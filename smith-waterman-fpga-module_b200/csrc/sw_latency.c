/*
 * sw_latency.c -- latency of the C ABI itself for one small job, measured from C (no interpreter in
 * the loop): the reference host's own use -- pack, submit, wait, read the score (main_test.c:290-528)
 * -- repeated `iters` times on the same handle.
 *
 *   sw_b200_latency -q <query.fa> -l <library.fa> [-n iters] [-e 0|1] [-o scores.txt]
 *
 * Per iteration: sw_score_batch (host buffers in) + sw_fetch (host scores out), timed with
 * clock_gettime(CLOCK_MONOTONIC).  -e 0 drops the CUDA events of the latency path
 * (sw_set_small_batch_timing).  Prints one JSON line; -o writes "name score" lines of the LAST
 * iteration so that the caller can check them against the oracle.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/sw_b200.h"

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

static int cmp_double(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : x > y;
}

int main(int argc, char **argv)
{
    const char *qf = NULL, *lf = NULL, *of = NULL;
    int iters = 2000, events = 1, c;
    while ((c = getopt(argc, argv, "q:l:n:e:o:")) != -1) {
        switch (c) {
            case 'q': qf = optarg; break;
            case 'l': lf = optarg; break;
            case 'n': iters = atoi(optarg); break;
            case 'e': events = atoi(optarg); break;
            case 'o': of = optarg; break;
            default: return 2;
        }
    }
    if (!qf || !lf || iters < 1) { fprintf(stderr, "usage: %s -q query.fa -l library.fa [-n iters] [-e 0|1] [-o scores]\n", argv[0]); return 2; }
    sw_seqset_t *q = NULL, *db = NULL;
    if (sw_read_fasta(qf, &q) != SW_OK) { printf("Query file error!\n"); return 1; }
    if (sw_read_fasta(lf, &db) != SW_OK) { printf("Library file error!\n"); return 1; }
    sw_handle_t *h = NULL;
    int rc = sw_init(&h, NULL, NULL, 0);
    if (rc != SW_OK) { printf("sw_init: %s\n", sw_strerror(rc)); return 1; }
    sw_set_small_batch_timing(h, events);
    rc = sw_set_queries(h, q->packed, q->len, q->off, (int)q->n);
    if (rc != SW_OK) { printf("sw_set_queries: %s\n", sw_strerror(rc)); return 1; }
    const size_t cap = q->n * db->n;
    int32_t *scores = (int32_t *)malloc(cap * sizeof(int32_t) + 16);
    double *t = (double *)malloc(sizeof(double) * (size_t)iters);
    double kms = 0.0;
    for (int it = -50; it < iters; ++it) {            /* 50 warm-up calls */
        const double t0 = now_us();
        rc = sw_score_batch(h, db->packed, db->len, db->off, NULL, db->n);
        if (rc == SW_OK) rc = sw_fetch(h, scores, cap, 10000);
        const double t1 = now_us();
        if (rc != SW_OK) { printf("scoring failed: %s (%s)\n", sw_strerror(rc), sw_last_cuda_error_string(h)); return 1; }
        if (it >= 0) { t[it] = t1 - t0; kms += sw_last_kernel_ms(h); }
    }
    qsort(t, (size_t)iters, sizeof(double), cmp_double);
    long long sum = 0;
    for (size_t i = 0; i < cap; ++i) sum += scores[i];
    printf("{\"queries\": %zu, \"subjects\": %zu, \"iters\": %d, \"events\": %d, \"kernel\": \"%s\", "
           "\"e2e_us_median\": %.2f, \"e2e_us_min\": %.2f, \"e2e_us_p90\": %.2f, \"e2e_us_p99\": %.2f, "
           "\"device_us_mean\": %.2f, \"cells\": %llu, \"score_sum\": %lld}\n",
           q->n, db->n, iters, events, sw_last_kernel_name(h), t[iters / 2], t[0], t[(size_t)(iters * 0.9)],
           t[(size_t)(iters * 0.99)], kms / iters * 1e3, (unsigned long long)sw_last_cells(h), sum);
    if (of) {
        FILE *f = fopen(of, "w");
        if (!f) return 1;
        for (size_t iq = 0; iq < q->n; ++iq)
            for (size_t is = 0; is < db->n; ++is) fprintf(f, "%zu %s %d\n", iq, db->name[is], scores[iq * db->n + is]);
        fclose(f);
    }
    free(t); free(scores);
    sw_destroy(h);
    sw_seqset_free(q); sw_seqset_free(db);
    return 0;
}

/* oracle/ref_bench.c — TEST/BENCH INFRASTRUCTURE ONLY: times the UNMODIFIED reference
 * (oracle/_ref/libroaring_ref.so) on the host cores, for bench.py's `cpu_baseline` leg and
 * `--impl reference` arm.  Nothing here is linked into the product library.
 *
 * The timed loops are the reference's own microbenchmark loops
 * (/root/reference/microbenchmarks/bench.cpp:85-96 SuccessiveIntersection, :196-207
 * SuccessiveUnion, :226-236 TotalUnion): result bitmaps are created, their cardinality read,
 * and freed inside the timed region.  Pairs are split statically over `nthreads` pthreads
 * (the library itself is single-threaded; BASELINE.md §3 "Cores").
 *
 * Prototypes are declared by hand (opaque pointers) so this file compiles anywhere the
 * prebuilt libroaring_ref.so is present.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct roaring_bitmap_s roaring_bitmap_t;
extern roaring_bitmap_t *roaring_bitmap_portable_deserialize_safe(const char *buf, size_t maxbytes);
extern void roaring_bitmap_free(const roaring_bitmap_t *r);
extern uint64_t roaring_bitmap_get_cardinality(const roaring_bitmap_t *r);
extern roaring_bitmap_t *roaring_bitmap_and(const roaring_bitmap_t *, const roaring_bitmap_t *);
extern roaring_bitmap_t *roaring_bitmap_or(const roaring_bitmap_t *, const roaring_bitmap_t *);
extern roaring_bitmap_t *roaring_bitmap_xor(const roaring_bitmap_t *, const roaring_bitmap_t *);
extern roaring_bitmap_t *roaring_bitmap_andnot(const roaring_bitmap_t *, const roaring_bitmap_t *);
extern uint64_t roaring_bitmap_and_cardinality(const roaring_bitmap_t *, const roaring_bitmap_t *);
extern roaring_bitmap_t *roaring_bitmap_or_many(size_t, const roaring_bitmap_t **);
extern int croaring_hardware_support(void);

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    int op; /* 0 and, 1 or, 2 xor, 3 andnot, 4 and_cardinality */
    roaring_bitmap_t **bms;
    const uint32_t *ia, *ib;
    size_t lo, hi;
    uint64_t sum;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    uint64_t s = 0;
    for (size_t p = j->lo; p < j->hi; p++) {
        const roaring_bitmap_t *a = j->bms[j->ia[p]], *b = j->bms[j->ib[p]];
        if (j->op == 4) {
            s += roaring_bitmap_and_cardinality(a, b);
            continue;
        }
        roaring_bitmap_t *r = j->op == 0   ? roaring_bitmap_and(a, b)
                              : j->op == 1 ? roaring_bitmap_or(a, b)
                              : j->op == 2 ? roaring_bitmap_xor(a, b)
                                           : roaring_bitmap_andnot(a, b);
        s += roaring_bitmap_get_cardinality(r);
        roaring_bitmap_free(r);
    }
    j->sum = s;
    return NULL;
}

/* Deserialize n bitmaps once; returns an opaque handle (array of pointers). */
void *refbench_load(size_t n, const char *const *bufs, const size_t *lens) {
    roaring_bitmap_t **bms = (roaring_bitmap_t **)calloc(n ? n : 1, sizeof(*bms));
    for (size_t i = 0; i < n; i++) {
        bms[i] = roaring_bitmap_portable_deserialize_safe(bufs[i], lens[i]);
        if (!bms[i]) return NULL;
    }
    return bms;
}

void refbench_unload(void *h, size_t n) {
    roaring_bitmap_t **bms = (roaring_bitmap_t **)h;
    for (size_t i = 0; i < n; i++) roaring_bitmap_free(bms[i]);
    free(bms);
}

/* One timed pass over the pair list with nthreads threads; returns seconds, *sumcard = checksum. */
double refbench_pairs(void *h, int op, const uint32_t *ia, const uint32_t *ib, size_t npairs,
                      int nthreads, uint64_t *sumcard) {
    roaring_bitmap_t **bms = (roaring_bitmap_t **)h;
    if (nthreads < 1) nthreads = 1;
    job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t].op = op;
        jobs[t].bms = bms;
        jobs[t].ia = ia;
        jobs[t].ib = ib;
        jobs[t].lo = npairs * (size_t)t / (size_t)nthreads;
        jobs[t].hi = npairs * (size_t)(t + 1) / (size_t)nthreads;
    }
    const double t0 = now_s();
    if (nthreads == 1) {
        worker(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    const double dt = now_s() - t0;
    uint64_t s = 0;
    for (int t = 0; t < nthreads; t++) s += jobs[t].sum;
    if (sumcard) *sumcard = s;
    free(jobs);
    free(th);
    return dt;
}

/* roaring_bitmap_or_many over bms[idx[0..n)], `reps` times back to back; returns seconds/rep. */
double refbench_or_many(void *h, const uint32_t *idx, size_t n, int reps, uint64_t *card) {
    roaring_bitmap_t **bms = (roaring_bitmap_t **)h;
    const roaring_bitmap_t **xs = (const roaring_bitmap_t **)calloc(n ? n : 1, sizeof(*xs));
    for (size_t i = 0; i < n; i++) xs[i] = bms[idx[i]];
    const double t0 = now_s();
    uint64_t c = 0;
    for (int r = 0; r < reps; r++) {
        roaring_bitmap_t *o = roaring_bitmap_or_many(n, xs);
        c = roaring_bitmap_get_cardinality(o);
        roaring_bitmap_free(o);
    }
    const double dt = (now_s() - t0) / (reps > 0 ? reps : 1);
    if (card) *card = c;
    free(xs);
    return dt;
}

/* bit 0: AVX2, bit 1: AVX-512 (src/isadetection.c:291-345) */
int refbench_hardware_support(void) { return croaring_hardware_support(); }

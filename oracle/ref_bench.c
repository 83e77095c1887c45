/* oracle/ref_bench.c — TEST/BENCH INFRASTRUCTURE ONLY: times the UNMODIFIED reference
 * (oracle/_ref/libroaring_ref.so) on the host cores, for bench.py's `cpu_baseline` leg and
 * `--impl reference` arm.  Nothing here is linked into the product library.
 *
 * The timed loops are the reference's own microbenchmark loops
 * (/root/reference/microbenchmarks/bench.cpp:85-96 SuccessiveIntersection, :196-207
 * SuccessiveUnion, :226-236 TotalUnion): result bitmaps are created, their cardinality read,
 * and freed inside the timed region.  The library itself is single-threaded; the N-thread figure
 * hands the independent pairs out dynamically to a persistent pthread pool (BASELINE.md §3 "Cores").
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
extern size_t roaring_bitmap_portable_size_in_bytes(const roaring_bitmap_t *r);
extern size_t roaring_bitmap_portable_serialize(const roaring_bitmap_t *r, char *buf);
extern int croaring_hardware_support(void);

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- persistent worker pool, dynamic chunks ------------------------------------------------
 * The N-thread throughput figure of SURVEY.md 8(d): the independent pairs are handed out in
 * chunks of CHUNK pairs from one atomic counter (all-pairs lists are triangular: a static split
 * would be unbalanced), to threads that live for the whole process (no pthread_create inside the
 * timed region). */
#define CHUNK 16
typedef struct {
    int op; /* 0 and, 1 or, 2 xor, 3 andnot, 4 and_cardinality, 5 deserialize */
    roaring_bitmap_t **bms;
    const uint32_t *ia, *ib;
    const char *const *bufs;
    const size_t *lens;
    size_t n;
    volatile size_t next;
    uint64_t sum;
    int failed;
} job_t;

static struct {
    pthread_mutex_t mu;
    pthread_cond_t cv, done;
    pthread_t *th;
    int nth;        /* threads created */
    int want;       /* helpers taking part in the current job */
    int started, running;
    uint64_t gen;
    job_t *job;
} P = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, NULL, 0, 0, 0, 0, 0, NULL};

static void run_job(job_t *j) {
    uint64_t s = 0;
    for (;;) {
        const size_t p0 = __sync_fetch_and_add(&j->next, (size_t)CHUNK);
        if (p0 >= j->n) break;
        const size_t p1 = p0 + CHUNK < j->n ? p0 + CHUNK : j->n;
        for (size_t p = p0; p < p1; p++) {
            if (j->op == 5) {
                j->bms[p] = roaring_bitmap_portable_deserialize_safe(j->bufs[p], j->lens[p]);
                if (!j->bms[p]) j->failed = 1;
                continue;
            }
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
    }
    __sync_fetch_and_add(&j->sum, s);
}

static void *pool_loop(void *arg) {
    (void)arg;
    uint64_t seen = 0;
    for (;;) {
        pthread_mutex_lock(&P.mu);
        while (!(P.gen != seen && P.job && P.started < P.want)) pthread_cond_wait(&P.cv, &P.mu);
        seen = P.gen;
        P.started++;
        P.running++;
        job_t *j = P.job;
        pthread_mutex_unlock(&P.mu);
        run_job(j);
        pthread_mutex_lock(&P.mu);
        P.running--;
        if (P.started == P.want && P.running == 0) pthread_cond_broadcast(&P.done);
        pthread_mutex_unlock(&P.mu);
    }
    return NULL;
}

/* run j on nthreads threads (the caller is one of them) */
static void pool_run(job_t *j, int nthreads) {
    if (nthreads <= 1) { run_job(j); return; }
    pthread_mutex_lock(&P.mu);
    if (P.nth < nthreads - 1) {
        P.th = (pthread_t *)realloc(P.th, sizeof(pthread_t) * (size_t)(nthreads - 1));
        while (P.nth < nthreads - 1) {
            if (pthread_create(&P.th[P.nth], NULL, pool_loop, NULL) != 0) break;
            P.nth++;
        }
    }
    P.job = j;
    P.want = nthreads - 1 < P.nth ? nthreads - 1 : P.nth;
    P.started = P.running = 0;
    P.gen++;
    pthread_cond_broadcast(&P.cv);
    pthread_mutex_unlock(&P.mu);
    run_job(j);
    pthread_mutex_lock(&P.mu);
    while (!(P.started == P.want && P.running == 0)) pthread_cond_wait(&P.done, &P.mu);
    P.job = NULL;
    pthread_mutex_unlock(&P.mu);
}

/* Start the pool's threads ahead of the first timed pass. */
void refbench_warm_pool(int nthreads) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.op = 4;
    pool_run(&j, nthreads);
}

/* Deserialize n bitmaps once (nthreads threads); returns an opaque handle (array of pointers). */
void *refbench_load_mt(size_t n, const char *const *bufs, const size_t *lens, int nthreads) {
    roaring_bitmap_t **bms = (roaring_bitmap_t **)calloc(n ? n : 1, sizeof(*bms));
    job_t j;
    memset(&j, 0, sizeof(j));
    j.op = 5;
    j.bms = bms;
    j.bufs = bufs;
    j.lens = lens;
    j.n = n;
    pool_run(&j, nthreads);
    if (j.failed) return NULL;
    return bms;
}
void *refbench_load(size_t n, const char *const *bufs, const size_t *lens) {
    return refbench_load_mt(n, bufs, lens, 1);
}

void refbench_unload(void *h, size_t n) {
    roaring_bitmap_t **bms = (roaring_bitmap_t **)h;
    for (size_t i = 0; i < n; i++) roaring_bitmap_free(bms[i]);
    free(bms);
}

/* One timed pass over the pair list with nthreads threads; returns seconds, *sumcard = checksum. */
double refbench_pairs(void *h, int op, const uint32_t *ia, const uint32_t *ib, size_t npairs,
                      int nthreads, uint64_t *sumcard) {
    job_t j;
    memset(&j, 0, sizeof(j));
    j.op = op;
    j.bms = (roaring_bitmap_t **)h;
    j.ia = ia;
    j.ib = ib;
    j.n = npairs;
    if (nthreads < 1) nthreads = 1;
    const double t0 = now_s();
    pool_run(&j, nthreads);
    const double dt = now_s() - t0;
    if (sumcard) *sumcard = j.sum;
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

/* One roaring_bitmap_or_many over bms[idx[0..n)] whose RESULT is returned as portable bytes
 * (malloc'd, release with refbench_free) for byte-level parity checks; returns the seconds of the
 * or_many call alone. */
double refbench_or_many_bytes(void *h, const uint32_t *idx, size_t n, char **out, size_t *outlen, uint64_t *card) {
    roaring_bitmap_t **bms = (roaring_bitmap_t **)h;
    const roaring_bitmap_t **xs = (const roaring_bitmap_t **)calloc(n ? n : 1, sizeof(*xs));
    for (size_t i = 0; i < n; i++) xs[i] = bms[idx ? idx[i] : i];
    const double t0 = now_s();
    roaring_bitmap_t *o = roaring_bitmap_or_many(n, xs);
    const double dt = now_s() - t0;
    if (card) *card = roaring_bitmap_get_cardinality(o);
    const size_t sz = roaring_bitmap_portable_size_in_bytes(o);
    *out = (char *)malloc(sz ? sz : 1);
    *outlen = roaring_bitmap_portable_serialize(o, *out);
    roaring_bitmap_free(o);
    free(xs);
    return dt;
}
void refbench_free(char *p) { free(p); }

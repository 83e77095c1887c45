/*
 * workgen.c — synthetic workload generators of BASELINE.json configs[2..4] (SURVEY.md §8(d) rows
 * 3, 4, 5), plain C + pthreads, host only.  They emit portable-serialized bitmaps
 * (format of /root/reference/src/roaring_array.c:469-531) directly, with the container types
 * roaring_bitmap_of_ptr + roaring_bitmap_run_optimize would choose (array iff card <= 4096,
 * src/roaring.c:60-90 / containers; then run iff 2 + 4*n_runs < current size,
 * src/containers/convert.c:217-250), so neither the reference nor the oracle is needed to BUILD
 * inputs; the parity tests feed the same bytes to both sides and tests/test_workgen.py checks the
 * emitted bytes against of_ptr + run_optimize + portable_serialize of the reference.
 *
 * The reference has no Zipf generator; the definition below is SURVEY.md §8(d)'s:
 *   PCG32 (the generator of /root/reference/benchmarks/random.h:18-30), stream of bitmap b:
 *   state = 0x853c49e6748fea9b ^ b, inc = 0xda3e39cb94b95bdb, first output discarded (it does not
 *   depend on the low bits of the state, i.e. on b).  One draw r -> u = (r + 1) / 2^32 in
 *   (0, 1]; v = floor(U^u) - 1 clipped to [0, U)  (P(v) ~ 1/(v+1): Zipf s = 1 by inverse CDF; U^u by
 *   the fixed double-precision recipe of zipf_value() below).
 *   Draw until the bitmap holds n distinct values.  With `density_draw` (config 5) the FIRST draw of
 *   the stream picks the bitmap's density log-uniformly: d = 0.001 * 300^u, n = max(1, round(d*U)).
 *   Config 4 (dense): bitmap i over universe 2^20, value j present iff draw number i*2^20 + j of the
 *   default pcg32_global stream (random.h:18-19) is odd — the single sequential stream of the
 *   survey's definition, reached in parallel with the LCG jump-ahead.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <sys/mman.h>

#define WG_API __attribute__((visibility("default")))

typedef struct { uint64_t state, inc; } pcg32_t;

static inline uint32_t pcg32_next(pcg32_t *g) {
    const uint64_t old = g->state;
    g->state = old * 6364136223846793005ULL + g->inc;
    const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((0u - rot) & 31u));
}

/* v = floor(U^u) - 1 clipped to [0, U), u = (r + 1) / 2^32, with U^u = 2^(u * log2 U) evaluated by
 * a fixed recipe in IEEE double arithmetic (no libm call per draw — the generator is bound by this
 * line): 2^x = 2^floor(x) * T[j] * P(t), j = floor(256 * frac(x)), t = ln2 * (frac(x) - j/256) < 0.0028,
 * P = Taylor polynomial of e^t to degree 4 (relative error < 2e-15), T[j] = pow(2, j/256) tabulated
 * once.  tests/test_workgen.py restates the same recipe operation by operation. */
static double g_T256[257];
static void zipf_tables(void) {
    static volatile int done = 0;
    if (done) return;
    for (int j = 0; j <= 256; j++) g_T256[j] = pow(2.0, (double)j / 256.0);
    __sync_synchronize();
    done = 1;
}
static inline uint64_t zipf_value(uint32_t r, double log2U, uint64_t U) {
    const double u = ((double)r + 1.0) * (1.0 / 4294967296.0);
    const double x = u * log2U;
    const int xi = (int)x;                        /* x >= 0: truncation is floor */
    const double f = x - (double)xi;
    const int j = (int)(f * 256.0);
    const double t = (f - (double)j * (1.0 / 256.0)) * 0.6931471805599453;
    const double p = 1.0 + t * (1.0 + t * (0.5 + t * ((1.0 / 6.0) + t * (1.0 / 24.0))));
    const double y = g_T256[j] * p * (double)(1ULL << xi);
    uint64_t q = (uint64_t)y;                     /* floor(y), y >= 1 */
    q = q ? q - 1 : 0;
    if (q >= U) q = U - 1;
    return q;
}

/* state after `delta` steps (the usual O(log delta) LCG jump) */
static uint64_t pcg32_jump(uint64_t state, uint64_t inc, uint64_t delta) {
    uint64_t cur_mult = 6364136223846793005ULL, cur_plus = inc, acc_mult = 1, acc_plus = 0;
    while (delta) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

/* ---------------------------------------------------------------- words -> portable bytes */
/* One bitmap from its membership bitset `bits` (universe rounded up to whole 2^16 chunks).
 * Returns a malloc'd portable blob.  run_optimize: apply convert_run_optimize's size rule. */
static uint8_t *emit_portable(const uint64_t *bits, uint32_t n_keys, int run_optimize, size_t *len_out) {
    /* pass 1: per key cardinality, runs, type */
    uint32_t *card = (uint32_t *)malloc(4 * (size_t)n_keys), *nrun = (uint32_t *)malloc(4 * (size_t)n_keys);
    uint8_t *type = (uint8_t *)malloc(n_keys);
    if (!card || !nrun || !type) { free(card); free(nrun); free(type); return NULL; }
    uint32_t n = 0;
    int hasrun = 0;
    size_t payload = 0;
    for (uint32_t k = 0; k < n_keys; k++) {
        const uint64_t *w = bits + (size_t)k * 1024;
        uint32_t c = 0, r = 0;
        uint64_t prev_top = 0;
        for (int i = 0; i < 1024; i++) {
            const uint64_t x = w[i];
            c += (uint32_t)__builtin_popcountll(x);
            r += (uint32_t)__builtin_popcountll(x & ~((x << 1) | prev_top));
            prev_top = x >> 63;
        }
        card[k] = c;
        nrun[k] = r;
        type[k] = 0;
        if (!c) continue;
        const uint32_t size_cur = c <= 4096 ? 2 * c : 8192, size_run = 2 + 4 * r;
        if (run_optimize && size_run < size_cur) { type[k] = 3; hasrun = 1; payload += size_run; }
        else if (c <= 4096) { type[k] = 2; payload += 2 * c; }
        else { type[k] = 1; payload += 8192; }
        n++;
    }
    size_t hdr;
    if (hasrun) hdr = 4 + (n + 7) / 8 + (n < 4 ? 4 * (size_t)n : 8 * (size_t)n);
    else hdr = 8 + 8 * (size_t)n;
    uint8_t *out = (uint8_t *)malloc(hdr + payload + 8);
    if (!out) { free(card); free(nrun); free(type); return NULL; }
    memset(out, 0, hdr);
    uint8_t *kc, *offs = NULL;
    if (hasrun) {
        const uint32_t cookie = 12347u | ((n - 1) << 16);
        memcpy(out, &cookie, 4);
        kc = out + 4 + (n + 7) / 8;
        if (n >= 4) offs = kc + 4 * (size_t)n;
    } else {
        const uint32_t cookie = 12346u;
        memcpy(out, &cookie, 4);
        memcpy(out + 4, &n, 4);
        kc = out + 8;
        offs = kc + 4 * (size_t)n;
    }
    size_t pos = hdr;
    uint32_t i = 0;
    for (uint32_t k = 0; k < n_keys; k++) {
        if (!type[k]) continue;
        const uint64_t *w = bits + (size_t)k * 1024;
        const uint16_t key = (uint16_t)k, cm1 = (uint16_t)(card[k] - 1);
        memcpy(kc + 4 * (size_t)i, &key, 2);
        memcpy(kc + 4 * (size_t)i + 2, &cm1, 2);
        if (offs) { const uint32_t o = (uint32_t)pos; memcpy(offs + 4 * (size_t)i, &o, 4); }
        if (type[k] == 3) {
            out[4 + (i >> 3)] |= (uint8_t)(1u << (i & 7));
            const uint16_t nr = (uint16_t)nrun[k];
            memcpy(out + pos, &nr, 2);
            uint16_t *rp = (uint16_t *)(out + pos + 2);
            int in_run = 0;
            uint32_t start = 0, q = 0;
            for (int wi = 0; wi < 1024; wi++) {
                uint64_t x = w[wi];
                if (!in_run && x == 0) continue;
                if (in_run && x == ~0ULL) continue;
                for (int b = 0; b < 64; b++) {
                    const int bit = (int)((x >> b) & 1);
                    if (bit && !in_run) { start = (uint32_t)(wi * 64 + b); in_run = 1; }
                    else if (!bit && in_run) {
                        rp[2 * q] = (uint16_t)start;
                        rp[2 * q + 1] = (uint16_t)((uint32_t)(wi * 64 + b) - 1 - start);
                        q++;
                        in_run = 0;
                    }
                }
            }
            if (in_run) { rp[2 * q] = (uint16_t)start; rp[2 * q + 1] = (uint16_t)(65535u - start); q++; }
            pos += 2 + 4 * (size_t)nrun[k];
        } else if (type[k] == 2) {
            uint16_t *ap = (uint16_t *)(out + pos);
            uint32_t q = 0;
            for (int wi = 0; wi < 1024; wi++) {
                uint64_t x = w[wi];
                while (x) {
                    ap[q++] = (uint16_t)(wi * 64 + __builtin_ctzll(x));
                    x &= x - 1;
                }
            }
            pos += 2 * (size_t)card[k];
        } else {
            memcpy(out + pos, w, 8192);
            pos += 8192;
        }
        i++;
    }
    free(card);
    free(nrun);
    free(type);
    *len_out = pos;
    return out;
}

/* ---------------------------------------------------------------- Zipf bitmaps (configs 3, 5) */
typedef struct {
    uint32_t b0, nb;
    uint64_t universe;
    const uint64_t *n_values; /* per bitmap, or NULL with density_draw */
    uint64_t n_fixed;
    int density_draw, run_optimize;
    uint8_t **blobs;
    size_t *lens;
    uint64_t *cards;
    volatile uint32_t next;
    volatile int failed;
} zipf_job_t;

static void *zipf_worker(void *arg) {
    zipf_job_t *J = (zipf_job_t *)arg;
    const uint64_t U = J->universe;
    const uint32_t n_keys = (uint32_t)((U + 65535) >> 16);
    /* the membership bitset: 2 MiB aligned + MADV_HUGEPAGE (128 workers first-touching 512 MiB each in
     * 4 KiB pages serialise on the address-space lock) */
    uint64_t *bits = NULL;
    const size_t bits_bytes = ((size_t)n_keys * 8192 + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    if (posix_memalign((void **)&bits, 2u << 20, bits_bytes) != 0) bits = NULL;
    if (!bits) { J->failed = 1; return NULL; }
    if (getenv("RB200_WG_THP")) madvise(bits, bits_bytes, MADV_HUGEPAGE);
    memset(bits, 0, bits_bytes);
    const double log2U = log2((double)U);
    /* batch buffers of the partitioned path (universes whose bitset exceeds ~16 MiB) */
    const uint64_t BATCH = 1u << 22;
    const uint32_t nbuckets = (uint32_t)((U + (1u << 21) - 1) >> 21);
    uint32_t *batch = NULL, *sorted = NULL, *hist = NULL;
    if (U > (1ULL << 27)) {
        batch = (uint32_t *)malloc(4 * BATCH);
        sorted = (uint32_t *)malloc(4 * BATCH);
        hist = (uint32_t *)malloc(4 * ((size_t)nbuckets + 2));
        if (!batch || !sorted || !hist) { free(batch); free(sorted); free(hist); batch = sorted = hist = NULL; }
    }
    for (;;) {
        const uint32_t i = __sync_fetch_and_add(&J->next, 1u);
        if (i >= J->nb || J->failed) break;
        const uint32_t b = J->b0 + i;
        pcg32_t g = {0x853c49e6748fea9bULL ^ (uint64_t)b, 0xda3e39cb94b95bdbULL};
        (void)pcg32_next(&g); /* the output function drops the low 27 state bits: without this
                                 warm-up step every stream b < 2^27 would start with the same draw */
        uint64_t want = J->n_values ? J->n_values[i] : J->n_fixed;
        if (J->density_draw) {
            const double u0 = ((double)pcg32_next(&g) + 1.0) * (1.0 / 4294967296.0);
            const double d = 0.001 * pow(300.0, u0);
            want = (uint64_t)llround(d * (double)U);
            if (want < 1) want = 1;
        }
        if (want > U) want = U;
        uint64_t have = 0;
        /* Large universes: the membership bitset (up to 512 MiB) does not fit any cache, so draws are
         * consumed in batches that are first partitioned by their high bits (2^21-value ranges =
         * 256 KiB of bitset each) and then applied range by range.  A batch never holds more draws
         * than distinct values still missing, so no draw past the stopping point of the sequential
         * definition is ever consumed and the resulting set is the same. */
        while (batch && want - have >= 65536) {
            uint64_t B = want - have;
            if (B > BATCH) B = BATCH;
            memset(hist, 0, sizeof(uint32_t) * (nbuckets + 1));
            for (uint64_t i = 0; i < B; i++) {
                const uint64_t v = zipf_value(pcg32_next(&g), log2U, U);
                batch[i] = (uint32_t)v;
                hist[(v >> 21) + 1]++;
            }
            for (uint32_t k = 0; k < nbuckets; k++) hist[k + 1] += hist[k];
            for (uint64_t i = 0; i < B; i++) sorted[hist[batch[i] >> 21]++] = batch[i];
            for (uint64_t i = 0; i < B; i++) {
                const uint64_t v = sorted[i];
                uint64_t *w = bits + (v >> 6);
                const uint64_t m = 1ULL << (v & 63);
                if (!(*w & m)) { *w |= m; have++; }
            }
        }
        while (have < want) {
            const uint64_t v = zipf_value(pcg32_next(&g), log2U, U);
            uint64_t *w = bits + (v >> 6);
            const uint64_t m = 1ULL << (v & 63);
            if (!(*w & m)) { *w |= m; have++; }
        }
        size_t len = 0;
        uint8_t *blob = emit_portable(bits, n_keys, J->run_optimize, &len);
        if (!blob) { J->failed = 1; break; }
        J->blobs[i] = blob;
        J->lens[i] = len;
        if (J->cards) J->cards[i] = have;
        memset(bits, 0, (size_t)n_keys * 8192);
    }
    free(bits);
    free(batch);
    free(sorted);
    free(hist);
    return NULL;
}

/* Bitmaps b0 .. b0+nb-1 of the Zipf family over [0, universe).  n_values: per-bitmap target
 * cardinalities (NULL: n_fixed for all, or drawn from the stream when density_draw != 0).
 * blobs[i] is malloc'd (release with rb200_workgen_free).  0 on success. */
WG_API int rb200_workgen_zipf(uint32_t b0, uint32_t nb, uint64_t universe, const uint64_t *n_values,
                              uint64_t n_fixed, int density_draw, int run_optimize, int threads,
                              uint8_t **blobs, size_t *lens, uint64_t *cards) {
    if (universe == 0 || universe > (1ULL << 32)) return -1;
    zipf_tables();
    zipf_job_t J;
    memset(&J, 0, sizeof(J));
    J.b0 = b0; J.nb = nb; J.universe = universe; J.n_values = n_values; J.n_fixed = n_fixed;
    J.density_draw = density_draw; J.run_optimize = run_optimize;
    J.blobs = blobs; J.lens = lens; J.cards = cards;
    for (uint32_t i = 0; i < nb; i++) { blobs[i] = NULL; lens[i] = 0; }
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > nb) threads = (int)nb;
    /* every worker owns a membership bitset of the whole universe (+ 32 MiB of batch buffers):
     * bound the total to a quarter of the physical memory, within [4, 96] GiB */
    const uint64_t per = ((universe + 65535) >> 16) * 8192ULL + (40ULL << 20);
    uint64_t budget = (uint64_t)sysconf(_SC_PHYS_PAGES) * (uint64_t)sysconf(_SC_PAGE_SIZE) / 4;
    if (budget < (4ULL << 30)) budget = 4ULL << 30;
    if (budget > (96ULL << 30)) budget = 96ULL << 30;
    const uint64_t cap = budget / per;
    if ((uint64_t)threads > cap) threads = (int)(cap ? cap : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int t = 0; t < threads; t++)
        if (pthread_create(&th[t], NULL, zipf_worker, &J) == 0) started++; else break;
    if (!started) zipf_worker(&J);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    if (J.failed) {
        for (uint32_t i = 0; i < nb; i++) { free(blobs[i]); blobs[i] = NULL; }
        return -1;
    }
    return 0;
}

/* ---------------------------------------------------------------- dense bitmaps (config 4) */
typedef struct {
    uint32_t i0, nb, n_keys;
    uint8_t **blobs;
    size_t *lens;
    volatile uint32_t next;
    volatile int failed;
} dense_job_t;

static void *dense_worker(void *arg) {
    dense_job_t *J = (dense_job_t *)arg;
    const uint32_t n_keys = J->n_keys;
    const uint64_t per = (uint64_t)n_keys << 16;
    uint64_t *bits = (uint64_t *)malloc((size_t)n_keys * 8192);
    if (!bits) { J->failed = 1; return NULL; }
    for (;;) {
        const uint32_t i = __sync_fetch_and_add(&J->next, 1u);
        if (i >= J->nb || J->failed) break;
        pcg32_t g = {0, 0xda3e39cb94b95bdbULL};
        g.state = pcg32_jump(0x853c49e6748fea9bULL, g.inc, (uint64_t)(J->i0 + i) * per);
        for (uint64_t w = 0; w < per / 64; w++) {
            uint64_t x = 0;
            for (int b = 0; b < 64; b++) x |= (uint64_t)(pcg32_next(&g) & 1u) << b;
            bits[w] = x;
        }
        size_t len = 0;
        uint8_t *blob = emit_portable(bits, n_keys, 0, &len);
        if (!blob) { J->failed = 1; break; }
        J->blobs[i] = blob;
        J->lens[i] = len;
    }
    free(bits);
    return NULL;
}

/* Bitmaps i0 .. i0+nb-1 of config 4: universe n_keys * 2^16, density 0.5 (see the header). */
WG_API int rb200_workgen_dense(uint32_t i0, uint32_t nb, uint32_t n_keys, int threads, uint8_t **blobs,
                               size_t *lens) {
    if (n_keys == 0 || n_keys > 65536) return -1;
    dense_job_t J;
    memset(&J, 0, sizeof(J));
    J.i0 = i0; J.nb = nb; J.n_keys = n_keys; J.blobs = blobs; J.lens = lens;
    for (uint32_t i = 0; i < nb; i++) { blobs[i] = NULL; lens[i] = 0; }
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > nb) threads = (int)nb;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int t = 0; t < threads; t++)
        if (pthread_create(&th[t], NULL, dense_worker, &J) == 0) started++; else break;
    if (!started) dense_worker(&J);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    if (J.failed) {
        for (uint32_t i = 0; i < nb; i++) { free(blobs[i]); blobs[i] = NULL; }
        return -1;
    }
    return 0;
}

WG_API void rb200_workgen_free(uint8_t **blobs, uint32_t nb) {
    for (uint32_t i = 0; i < nb; i++) { free(blobs[i]); blobs[i] = NULL; }
}

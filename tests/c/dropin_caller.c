/* tests/c/dropin_caller.c — a plain C caller of the REFERENCE API, compiled against the reference's
 * own headers (/root/reference/include) by oracle/Makefile.  TEST INFRASTRUCTURE.
 *
 * It runs the loops of the reference's own drivers — SuccessiveIntersection / SuccessiveUnion
 * (/root/reference/microbenchmarks/bench.cpp:85-96, 196-207), the successive and/or/xor/andnot,
 * in-place, or_many / xor_many checks of /root/reference/tests/realdata_unit.c:323-446 — and prints
 * one checksum line per loop (cardinalities + an FNV-1a hash of the reference's portable
 * serialisation of every result, so container TYPES count).  The same source is linked twice:
 *   dropin_caller_ref   -lroaring_ref                    (the reference alone)
 *   dropin_caller_b200  -lroaring_b200 -lroaring_ref     (our symbols first: "drops into callers unchanged")
 * and dropin_caller_ref is also run under LD_PRELOAD=libroaring_b200.so.  Every result is validated
 * with roaring_bitmap_internal_validate and released with the reference's roaring_bitmap_free.
 * With the argument "hook" all allocations go through roaring_init_memory_hook counters.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <roaring/roaring.h>
#include <roaring/roaring64.h>
#include <roaring/memory.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static long n_alloc = 0, n_free = 0;
static void *h_malloc(size_t n) { __sync_fetch_and_add(&n_alloc, 1); return malloc(n); }
static void *h_realloc(void *p, size_t n) { if (!p) __sync_fetch_and_add(&n_alloc, 1); return realloc(p, n); }
static void *h_calloc(size_t a, size_t b) { __sync_fetch_and_add(&n_alloc, 1); return calloc(a, b); }
static void h_free(void *p) { if (p) __sync_fetch_and_add(&n_free, 1); free(p); }
static void *h_amalloc(size_t al, size_t n) {
    void *p = NULL;
    __sync_fetch_and_add(&n_alloc, 1);
    return posix_memalign(&p, al, n) == 0 ? p : NULL;
}
static void h_afree(void *p) { if (p) __sync_fetch_and_add(&n_free, 1); free(p); }

static uint64_t fnv(uint64_t h, const char *p, size_t n) {
    for (size_t i = 0; i < n; i++) { h ^= (unsigned char)p[i]; h *= 1099511628211ULL; }
    return h;
}
static uint64_t hash_bitmap(uint64_t h, const roaring_bitmap_t *r) {
    const char *why = NULL;
    if (!roaring_bitmap_internal_validate(r, &why)) { printf("INVALID result: %s\n", why ? why : "?"); exit(3); }
    const size_t n = roaring_bitmap_portable_size_in_bytes(r);
    char *buf = (char *)malloc(n);
    roaring_bitmap_portable_serialize(r, buf);
    h = fnv(h, buf, n);
    free(buf);
    return h;
}

typedef roaring_bitmap_t *(*binop)(const roaring_bitmap_t *, const roaring_bitmap_t *);
typedef void (*inop)(roaring_bitmap_t *, const roaring_bitmap_t *);

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const int hook = argc > 2 && strcmp(argv[2], "hook") == 0;
    if (hook) {
        roaring_memory_t m = {h_malloc, h_realloc, h_calloc, h_free, h_amalloc, h_afree};
        roaring_init_memory_hook(m);
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) return 2;
    uint32_t *lens = (uint32_t *)malloc(4 * n);
    if (fread(lens, 4, n, f) != n) return 2;
    roaring_bitmap_t **b = (roaring_bitmap_t **)malloc(sizeof(*b) * n);
    for (uint32_t i = 0; i < n; i++) {
        char *buf = (char *)malloc(lens[i]);
        if (fread(buf, 1, lens[i], f) != lens[i]) return 2;
        b[i] = roaring_bitmap_portable_deserialize_safe(buf, lens[i]);
        free(buf);
        if (!b[i]) return 2;
    }
    fclose(f);

    const char *names[4] = {"and", "or", "xor", "andnot"};
    binop ops[4] = {roaring_bitmap_and, roaring_bitmap_or, roaring_bitmap_xor, roaring_bitmap_andnot};
    inop iops[4] = {roaring_bitmap_and_inplace, roaring_bitmap_or_inplace, roaring_bitmap_xor_inplace,
                    roaring_bitmap_andnot_inplace};
    for (int o = 0; o < 4; o++) {
        uint64_t card = 0, h = 1469598103934665603ULL;
        for (uint32_t i = 0; i + 1 < n; i++) {   /* bench.cpp:85-96 */
            roaring_bitmap_t *r = ops[o](b[i], b[i + 1]);
            if (!r) { printf("NULL result\n"); return 3; }
            card += roaring_bitmap_get_cardinality(r);
            h = hash_bitmap(h, r);
            roaring_bitmap_free(r);
        }
        printf("successive_%s card=%llu hash=%016llx\n", names[o], (unsigned long long)card, (unsigned long long)h);
        card = 0;
        h = 1469598103934665603ULL;
        for (uint32_t i = 0; i + 1 < n; i++) {   /* realdata_unit.c:323-349 */
            roaring_bitmap_t *c = roaring_bitmap_copy(b[i]);
            iops[o](c, b[i + 1]);
            card += roaring_bitmap_get_cardinality(c);
            h = hash_bitmap(h, c);
            roaring_bitmap_free(c);
        }
        printf("inplace_%s card=%llu hash=%016llx\n", names[o], (unsigned long long)card, (unsigned long long)h);
    }
    {
        uint64_t s = 0, inter = 0;
        double j = 0;
        for (uint32_t i = 0; i + 1 < n; i++) {
            s += roaring_bitmap_and_cardinality(b[i], b[i + 1]);
            s += 3 * roaring_bitmap_or_cardinality(b[i], b[i + 1]);
            inter += roaring_bitmap_intersect(b[i], b[i + 1]) ? 1 : 0;
            j += roaring_bitmap_jaccard_index(b[i], b[i + 1]);
        }
        printf("cardinalities sum=%llu intersect=%llu jaccard=%.12f\n", (unsigned long long)s,
               (unsigned long long)inter, j);
    }
    {
        roaring_bitmap_t *r = roaring_bitmap_or_many(n, (const roaring_bitmap_t **)b);
        printf("or_many card=%llu hash=%016llx\n", (unsigned long long)roaring_bitmap_get_cardinality(r),
               (unsigned long long)hash_bitmap(1469598103934665603ULL, r));
        roaring_bitmap_free(r);
        r = roaring_bitmap_xor_many(n, (const roaring_bitmap_t **)b);
        printf("xor_many card=%llu hash=%016llx\n", (unsigned long long)roaring_bitmap_get_cardinality(r),
               (unsigned long long)hash_bitmap(1469598103934665603ULL, r));
        roaring_bitmap_free(r);
        r = roaring_bitmap_or_many_heap(n, (const roaring_bitmap_t **)b);
        printf("or_many_heap card=%llu hash=%016llx\n", (unsigned long long)roaring_bitmap_get_cardinality(r),
               (unsigned long long)hash_bitmap(1469598103934665603ULL, r));
        roaring_bitmap_free(r);
    }
    {   /* 64-bit bitmaps (roaring64.h:423-522): the same inputs spread over three high-32 buckets */
        roaring64_bitmap_t *q[6];
        for (int k = 0; k < 6; k++) {
            q[k] = roaring64_bitmap_create();
            for (int part = 0; part < 3; part++) {
                const roaring_bitmap_t *src = b[(uint32_t)(5 * k + 3 * part) % n];
                const uint64_t card = roaring_bitmap_get_cardinality(src);
                uint32_t *vals = (uint32_t *)malloc(4 * (card ? card : 1));
                roaring_bitmap_to_uint32_array(src, vals);
                const uint64_t high = (uint64_t)(part == 2 ? 0xFFFFFFFFu : (uint32_t)(part * 7 + (k & 1))) << 32;
                for (uint64_t i = 0; i < card; i++) roaring64_bitmap_add(q[k], high | vals[i]);
                free(vals);
            }
            roaring64_bitmap_run_optimize(q[k]);
        }
        typedef roaring64_bitmap_t *(*binop64)(const roaring64_bitmap_t *, const roaring64_bitmap_t *);
        binop64 ops64[4] = {roaring64_bitmap_and, roaring64_bitmap_or, roaring64_bitmap_xor, roaring64_bitmap_andnot};
        for (int o = 0; o < 4; o++) {
            uint64_t card = 0, h = 1469598103934665603ULL;
            for (int k = 0; k + 1 < 6; k++) {
                roaring64_bitmap_t *r = ops64[o](q[k], q[k + 1]);
                if (!r) { printf("NULL r64 result\n"); return 3; }
                const char *why = NULL;
                if (!roaring64_bitmap_internal_validate(r, &why)) { printf("INVALID r64: %s\n", why ? why : "?"); return 3; }
                card += roaring64_bitmap_get_cardinality(r);
                const size_t sz = roaring64_bitmap_portable_size_in_bytes(r);
                char *buf = (char *)malloc(sz);
                roaring64_bitmap_portable_serialize(r, buf);
                h = fnv(h, buf, sz);
                free(buf);
                roaring64_bitmap_free(r);
            }
            printf("r64_%s card=%llu hash=%016llx\n", names[o], (unsigned long long)card, (unsigned long long)h);
        }
        uint64_t s = 0;
        for (int k = 0; k + 1 < 6; k++)
            s += roaring64_bitmap_and_cardinality(q[k], q[k + 1]) + 3 * roaring64_bitmap_or_cardinality(q[k], q[k + 1]) +
                 (roaring64_bitmap_intersect(q[k], q[k + 1]) ? 1 : 0);
        printf("r64_cardinalities sum=%llu\n", (unsigned long long)s);
        for (int k = 0; k < 6; k++) roaring64_bitmap_free(q[k]);
    }
    for (uint32_t i = 0; i < n; i++) roaring_bitmap_free(b[i]);
    free(b);
    free(lens);
    if (hook) printf("hook outstanding=%ld\n", n_alloc - n_free);
    /* which implementation served the calls: our library counts its kernel launches */
    uint64_t (*launches)(void) = (uint64_t(*)(void))dlsym(RTLD_DEFAULT, "rb200_kernel_launches");
    fprintf(stderr, "kernel_launches=%lld\n", launches ? (long long)launches() : -1LL);
    return 0;
}

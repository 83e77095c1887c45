/* oracle/roaring_oracle.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * A plain-C, scalar, single-threaded CPU restatement of the reference hot path
 * (CRoaring 5.1.0, /root/reference): pairwise and/or/xor/andnot, and_cardinality,
 * or_many, xor_many over the container x container grid, INCLUDING the reference's
 * result-container TYPE rules, so that its output — the portable serialization of the
 * result bitmap — is byte-identical to what the reference produces on the same inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The product library (croaring_b200/) must never link or call it.
 *
 * PARITY PINNED: tests/test_oracle_pinning.py checks this file byte-for-byte against
 *   (a) the unmodified reference compiled into oracle/_ref/libroaring_ref.so on the six
 *       real-data sets + seeded synthetic container mixes (all 9 type pairings, full
 *       containers, the or_many full-container state machine), and
 *   (b) the committed golden sha256/sum-card values in tests/golden/realdata_golden.json
 *       (generated from the reference by tests/golden/make_golden.py).
 *
 * How it restates the reference: every grid cell computes its value set on a 65536-bit
 * scratch bitset (scalar word loops) and then applies the cell's own type rule, each rule
 * citing the reference function it follows.  The many-way ops replay the reference's
 * lazy fold literally (lazy_or / lazy_or_inplace / lazy_xor / lazy_xor_inplace + repair).
 *
 * All file:line citations are relative to /root/reference.
 */
#include "roaring_oracle.h"

#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#define T_BITSET 1 /* include/roaring/containers/containers.h:48-51 */
#define T_ARRAY 2
#define T_RUN 3
#define WORDS 1024                     /* containers/bitset.h:40 */
#define MAXARR 4096                    /* containers/array.h:38 DEFAULT_MAX_SIZE */
#define LAZY_LOWER 1024                /* containers/perfparameters.h:28 ARRAY_LAZY_LOWERBOUND */
#define COOKIE_RUN 12347u              /* roaring_array.h:35-40 */
#define COOKIE_NORUN 12346u
#define NO_OFFSET_THRESHOLD 4

typedef struct {
    uint8_t type;
    int32_t card; /* bitset: cardinality or -1 = BITSET_UNKNOWN_CARDINALITY (bitset.h:42) */
    int32_t n;    /* array: #values, run: #runs */
    uint16_t *v;  /* array values / run (value,length) pairs */
    uint64_t *w;  /* bitset words */
} oc_t;

typedef struct {
    int32_t n, cap;
    uint16_t *keys;
    oc_t *c;
} obm_t;

/* ------------------------------------------------------------------ utilities */
static int popc64(uint64_t x) { return __builtin_popcountll(x); }

static void oc_free(oc_t *c) {
    free(c->v);
    free(c->w);
    c->v = NULL;
    c->w = NULL;
}

static oc_t oc_clone(const oc_t *c) {
    oc_t r = *c;
    if (c->w) {
        r.w = (uint64_t *)malloc(WORDS * 8);
        memcpy(r.w, c->w, WORDS * 8);
    }
    if (c->v) {
        size_t bytes = (size_t)(c->type == T_RUN ? 4 : 2) * (size_t)(c->n ? c->n : 1);
        r.v = (uint16_t *)malloc(bytes);
        memcpy(r.v, c->v, (size_t)(c->type == T_RUN ? 4 : 2) * (size_t)c->n);
    }
    return r;
}

static void words_set_range(uint64_t *w, uint32_t lo, uint32_t hi /*inclusive*/) {
    for (uint32_t x = lo; x <= hi; x++) w[x >> 6] |= UINT64_C(1) << (x & 63);
}

static void to_words(const oc_t *c, uint64_t *w) {
    if (c->type == T_BITSET) {
        memcpy(w, c->w, WORDS * 8);
        return;
    }
    memset(w, 0, WORDS * 8);
    if (c->type == T_ARRAY) {
        for (int i = 0; i < c->n; i++) w[c->v[i] >> 6] |= UINT64_C(1) << (c->v[i] & 63);
    } else {
        for (int i = 0; i < c->n; i++)
            words_set_range(w, c->v[2 * i], (uint32_t)c->v[2 * i] + c->v[2 * i + 1]);
    }
}

static int words_card(const uint64_t *w) {
    int s = 0;
    for (int i = 0; i < WORDS; i++) s += popc64(w[i]);
    return s;
}

/* number of maximal runs of set bits = number of (bit set, previous bit clear) positions */
static int words_nruns(const uint64_t *w) {
    int s = 0;
    uint64_t carry = 0; /* bit 63 of previous word */
    for (int i = 0; i < WORDS; i++) {
        uint64_t x = w[i];
        s += popc64(x & ~((x << 1) | carry));
        carry = x >> 63;
    }
    return s;
}

static int run_card(const oc_t *c) { /* run.c:1077 run_container_cardinality */
    int s = 0;
    for (int i = 0; i < c->n; i++) s += c->v[2 * i + 1] + 1;
    return s;
}

static bool run_is_full(const oc_t *c) { /* containers/run.h:394-397 */
    return c->type == T_RUN && c->n == 1 && c->v[0] == 0 && c->v[1] == 0xFFFF;
}

static int oc_card(const oc_t *c) {
    if (c->type == T_ARRAY) return c->n;
    if (c->type == T_RUN) return run_card(c);
    return c->card >= 0 ? c->card : words_card(c->w);
}

static oc_t mk_bitset(const uint64_t *w, int card) {
    oc_t r = {T_BITSET, card, 0, NULL, (uint64_t *)malloc(WORDS * 8)};
    memcpy(r.w, w, WORDS * 8);
    return r;
}

static oc_t mk_array(const uint64_t *w) { /* convert.c:55 array_container_from_bitset */
    int card = words_card(w);
    oc_t r = {T_ARRAY, card, card, (uint16_t *)malloc(2 * (size_t)(card ? card : 1)), NULL};
    int k = 0;
    for (int i = 0; i < WORDS; i++) {
        uint64_t x = w[i];
        while (x) {
            r.v[k++] = (uint16_t)(i * 64 + __builtin_ctzll(x));
            x &= x - 1;
        }
    }
    return r;
}

static oc_t mk_run(const uint64_t *w) {
    int n = words_nruns(w);
    oc_t r = {T_RUN, 0, n, (uint16_t *)malloc(4 * (size_t)(n ? n : 1)), NULL};
    int k = 0;
    int32_t start = -1;
    for (uint32_t x = 0; x < 65536; x++) {
        bool set = (w[x >> 6] >> (x & 63)) & 1;
        if (set && start < 0) start = (int32_t)x;
        if (!set && start >= 0) {
            r.v[2 * k] = (uint16_t)start;
            r.v[2 * k + 1] = (uint16_t)(x - 1 - (uint32_t)start);
            k++;
            start = -1;
        }
        if (!set && (x & 63) == 0 && w[x >> 6] == 0) x += 63; /* skip empty word */
    }
    if (start >= 0) {
        r.v[2 * k] = (uint16_t)start;
        r.v[2 * k + 1] = (uint16_t)(65535 - start);
        k++;
    }
    r.n = k;
    return r;
}

/* "card <= DEFAULT_MAX_SIZE ? array : bitset" — the rule every bitset-producing cell ends
 * with, e.g. mixed_intersection.c:305-330, mixed_xor.c:260-272, mixed_andnot.c:482-494 */
static oc_t ab_from_words(const uint64_t *w) {
    int card = words_card(w);
    return card <= MAXARR ? mk_array(w) : mk_bitset(w, card);
}

/* convert_run_to_efficient_container, convert.c:154-200: stay RUN iff
 * 2+4*n_runs <= min(8192, 2*card); else array if card<=4096 else bitset. */
static oc_t eff_from_words(const uint64_t *w) {
    int card = words_card(w), nr = words_nruns(w);
    int size_run = 2 + 4 * nr, size_arr = 2 * card;
    int min_non_run = 8192 < size_arr ? 8192 : size_arr;
    if (size_run <= min_non_run) return mk_run(w);
    return card <= MAXARR ? mk_array(w) : mk_bitset(w, card);
}

static oc_t empty_array(void) {
    oc_t r = {T_ARRAY, 0, 0, (uint16_t *)malloc(2), NULL};
    return r;
}

/* ------------------------------------------------------------------ grid cells */
#define PAIR(a, b) (4 * (a) + (b)) /* containers.h:62-65 */

/* container_and, containers.h:726-806 */
static oc_t cell_and(const oc_t *c1, const oc_t *c2) {
    uint64_t w1[WORDS], w2[WORDS];
    to_words(c1, w1);
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) w1[i] &= w2[i];
    switch (PAIR(c1->type, c2->type)) {
        case PAIR(T_BITSET, T_BITSET): /* mixed_intersection.c:305 */
            return ab_from_words(w1);
        case PAIR(T_RUN, T_RUN): /* run.c:387 + convert.c:203 */
            return eff_from_words(w1);
        case PAIR(T_BITSET, T_RUN):
        case PAIR(T_RUN, T_BITSET): { /* mixed_intersection.c:117-203 */
            const oc_t *r = c1->type == T_RUN ? c1 : c2;
            const oc_t *b = c1->type == T_RUN ? c2 : c1;
            if (run_is_full(r)) return mk_bitset(b->w, oc_card(b)); /* clone, :120-123 */
            if (run_card(r) <= MAXARR) return mk_array(w1);          /* :124-146 */
            return ab_from_words(w1);                                /* :171-202 */
        }
        default: /* A,A (array.c:288)  A,B/B,A (mixed_intersection.c:19)  A,R/R,A (:73) */
            return mk_array(w1);
    }
}

/* container_or, containers.h:1008-1103 */
static oc_t cell_or(const oc_t *c1, const oc_t *c2) {
    uint64_t w1[WORDS], w2[WORDS];
    to_words(c1, w1);
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) w1[i] |= w2[i];
    switch (PAIR(c1->type, c2->type)) {
        case PAIR(T_BITSET, T_BITSET): /* containers.h:1015-1020: stays bitset even if full */
        case PAIR(T_BITSET, T_ARRAY):
        case PAIR(T_ARRAY, T_BITSET): /* mixed_union.c:22 */
            return mk_bitset(w1, words_card(w1));
        case PAIR(T_ARRAY, T_ARRAY): /* mixed_union.c:162-192 */
            if (c1->n + c2->n <= MAXARR) return mk_array(w1);
            return ab_from_words(w1);
        case PAIR(T_BITSET, T_RUN):
        case PAIR(T_RUN, T_BITSET): { /* containers.h:1056-1080 */
            const oc_t *r = c1->type == T_RUN ? c1 : c2;
            if (run_is_full(r)) return oc_clone(r);
            return mk_bitset(w1, words_card(w1));
        }
        default: /* R,R (run.c:231)  A,R/R,A (mixed_union.c:66): run result + eff */
            return eff_from_words(w1);
    }
}

/* array_array_container_xor, mixed_xor.c:196-219 (given total input cardinality) */
static oc_t aa_xor_rule(const uint64_t *w, int total) {
    if (total <= MAXARR) return mk_array(w);
    return ab_from_words(w);
}

/* container_xor, containers.h:1449-1524 (container_ixor uses the same rules,
 * mixed_xor.c:302-376 all delegate) */
static oc_t cell_xor(const oc_t *c1, const oc_t *c2) {
    uint64_t w1[WORDS], w2[WORDS];
    to_words(c1, w1);
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) w1[i] ^= w2[i];
    switch (PAIR(c1->type, c2->type)) {
        case PAIR(T_ARRAY, T_ARRAY):
            return aa_xor_rule(w1, c1->n + c2->n);
        case PAIR(T_RUN, T_RUN): /* mixed_xor.c:179 */
            return eff_from_words(w1);
        case PAIR(T_ARRAY, T_RUN):
        case PAIR(T_RUN, T_ARRAY): { /* array_run_container_xor, mixed_xor.c:104-138 */
            const oc_t *a = c1->type == T_ARRAY ? c1 : c2;
            const oc_t *r = c1->type == T_ARRAY ? c2 : c1;
            if (a->n < 32) return eff_from_words(w1);
            int rc = run_card(r);
            if (rc <= MAXARR) return aa_xor_rule(w1, rc + a->n);
            return ab_from_words(w1);
        }
        default: /* B,B (mixed_xor.c:260)  A,B/B,A (:23)  R,B/B,R (:61) */
            return ab_from_words(w1);
    }
}

/* container_andnot, containers.h:1783-1876 */
static oc_t cell_andnot(const oc_t *c1, const oc_t *c2) {
    uint64_t w1[WORDS], w2[WORDS];
    to_words(c1, w1);
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) w1[i] &= ~w2[i];
    switch (PAIR(c1->type, c2->type)) {
        case PAIR(T_ARRAY, T_ARRAY):  /* array.c:226 */
        case PAIR(T_ARRAY, T_BITSET): /* mixed_andnot.c:24 */
        case PAIR(T_ARRAY, T_RUN):    /* mixed_andnot.c:381 (full run -> empty, :1855) */
            return mk_array(w1);
        case PAIR(T_RUN, T_RUN): /* containers.h:1807-1815 */
            if (run_is_full(c2)) return empty_array();
            return eff_from_words(w1);
        case PAIR(T_BITSET, T_RUN): /* containers.h:1833-1844, mixed_andnot.c:175 */
            if (run_is_full(c2)) return empty_array();
            return ab_from_words(w1);
        case PAIR(T_RUN, T_BITSET): /* mixed_andnot.c:104-148 */
            if (run_card(c1) <= MAXARR) return mk_array(w1);
            return ab_from_words(w1);
        case PAIR(T_RUN, T_ARRAY): { /* mixed_andnot.c:277-357 */
            int card = run_card(c1);
            if (card <= 32) {
                if (c2->n == 0) return oc_clone(c1);
                return eff_from_words(w1);
            }
            if (card <= MAXARR) return mk_array(w1);
            return ab_from_words(w1);
        }
        default: /* B,B (mixed_andnot.c:482)  B,A (:54) */
            return ab_from_words(w1);
    }
}

/* container_and_cardinality, containers.h:811-859 */
static int cell_and_card(const oc_t *c1, const oc_t *c2) {
    uint64_t w1[WORDS], w2[WORDS];
    to_words(c1, w1);
    to_words(c2, w2);
    int s = 0;
    for (int i = 0; i < WORDS; i++) s += popc64(w1[i] & w2[i]);
    return s;
}

/* ------------------------------------------------------------------ lazy cells */
static oc_t to_bitset(const oc_t *c) { /* containers.h:166-185 container_to_bitset */
    uint64_t w[WORDS];
    to_words(c, w);
    return mk_bitset(w, oc_card(c));
}

/* container_lazy_or, containers.h:1113-1215.  Only cells with >=1 bitset are reached from
 * or_many (LAZY_OR_BITSET_CONVERSION=true, roaring.c:2535-2545); the others are restated
 * for completeness. */
static oc_t cell_lazy_or(const oc_t *c1, const oc_t *c2) {
    uint64_t w1[WORDS], w2[WORDS];
    to_words(c1, w1);
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) w1[i] |= w2[i];
    switch (PAIR(c1->type, c2->type)) {
        case PAIR(T_BITSET, T_BITSET): /* or_nocard */
        case PAIR(T_BITSET, T_ARRAY):
        case PAIR(T_ARRAY, T_BITSET):
            return mk_bitset(w1, -1);
        case PAIR(T_ARRAY, T_ARRAY): /* mixed_union.c:247-283 */
            if (c1->n + c2->n <= LAZY_LOWER) return mk_array(w1);
            return mk_bitset(w1, -1);
        case PAIR(T_BITSET, T_RUN):
        case PAIR(T_RUN, T_BITSET): {
            const oc_t *r = c1->type == T_RUN ? c1 : c2;
            if (run_is_full(r)) return oc_clone(r);
            return mk_bitset(w1, -1);
        }
        case PAIR(T_RUN, T_RUN):
            return eff_from_words(w1);
        default: /* A,R / R,A left as RUN */
            return mk_run(w1);
    }
}

/* container_lazy_ior with a BITSET accumulator, containers.h:1342-1404.  Replaces *acc. */
static void cell_lazy_ior_bitset(oc_t *acc, const oc_t *c2) {
    uint64_t w2[WORDS];
    if (c2->type == T_RUN && run_is_full(c2)) { /* :1394-1399 */
        oc_free(acc);
        *acc = oc_clone(c2);
        return;
    }
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) acc->w[i] |= w2[i];
    if (c2->type == T_BITSET) { /* :1342-1352, LAZY_OR_BITSET_CONVERSION_TO_FULL defined */
        acc->card = words_card(acc->w);
        if (acc->card == 65536) {
            oc_free(acc);
            oc_t r = {T_RUN, 0, 1, (uint16_t *)malloc(4), NULL};
            r.v[0] = 0;
            r.v[1] = 0xFFFF;
            *acc = r;
        }
    } else {
        acc->card = -1;
    }
}

/* container_lazy_ior for every type pairing, containers.h:1333-1440.  Replaces *c1. */
static void cell_lazy_ior(oc_t *c1, const oc_t *c2) {
    if (c1->type == T_BITSET) { /* B,B / B,A / B,R */
        cell_lazy_ior_bitset(c1, c2);
        return;
    }
    uint64_t w1[WORDS], w2[WORDS];
    oc_t r;
    if (c1->type == T_RUN && c2->type == T_BITSET && run_is_full(c1)) return; /* :1409-1412 */
    to_words(c1, w1);
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) w1[i] |= w2[i];
    switch (PAIR(c1->type, c2->type)) {
        case PAIR(T_ARRAY, T_ARRAY): /* mixed_union.c:285-372 */
            r = (c1->n + c2->n <= LAZY_LOWER) ? mk_array(w1) : mk_bitset(w1, -1);
            break;
        case PAIR(T_RUN, T_RUN): /* :1366-1370 union + convert_run_to_efficient_container */
            r = eff_from_words(w1);
            break;
        case PAIR(T_ARRAY, T_BITSET): /* :1380-1388 */
        case PAIR(T_RUN, T_BITSET):   /* :1413-1419 */
            r = mk_bitset(w1, -1);
            break;
        default: /* A,R / R,A: left as RUN, :1421-1440 */
            r = mk_run(w1);
            break;
    }
    oc_free(c1);
    *c1 = r;
}

/* container_is_full, containers.h:262-277 */
static bool oc_is_full(const oc_t *c) {
    if (c->type == T_BITSET) return c->card == 65536;
    if (c->type == T_ARRAY) return c->n == 65536;
    return run_is_full(c);
}

/* container_lazy_xor, containers.h:1570-1654 */
static oc_t cell_lazy_xor(const oc_t *c1, const oc_t *c2) {
    uint64_t w1[WORDS], w2[WORDS];
    to_words(c1, w1);
    to_words(c2, w2);
    for (int i = 0; i < WORDS; i++) w1[i] ^= w2[i];
    switch (PAIR(c1->type, c2->type)) {
        case PAIR(T_ARRAY, T_ARRAY): /* mixed_xor.c:221-252 */
            if (c1->n + c2->n <= LAZY_LOWER) return mk_array(w1);
            return mk_bitset(w1, -1);
        case PAIR(T_RUN, T_RUN):
            return eff_from_words(w1);
        case PAIR(T_ARRAY, T_RUN):
        case PAIR(T_RUN, T_ARRAY): /* mixed_xor.c:145-174, left as RUN */
            return mk_run(w1);
        default:
            return mk_bitset(w1, -1);
    }
}

/* container_repair_after_lazy, containers.h:344-371 */
static void repair(oc_t *c) {
    if (c->type == T_BITSET) {
        c->card = words_card(c->w);
        if (c->card <= MAXARR) {
            oc_t a = mk_array(c->w);
            oc_free(c);
            *c = a;
        }
    } else if (c->type == T_RUN) {
        uint64_t w[WORDS];
        to_words(c, w);
        oc_t e = eff_from_words(w);
        oc_free(c);
        *c = e;
    }
}

/* ------------------------------------------------------------------ bitmap level */
static void bm_init(obm_t *b, int cap) {
    b->n = 0;
    b->cap = cap > 4 ? cap : 4;
    b->keys = (uint16_t *)malloc(2 * (size_t)b->cap);
    b->c = (oc_t *)malloc(sizeof(oc_t) * (size_t)b->cap);
}

static void bm_free(obm_t *b) {
    for (int i = 0; i < b->n; i++) oc_free(&b->c[i]);
    free(b->keys);
    free(b->c);
}

static void bm_reserve(obm_t *b, int n) {
    if (n <= b->cap) return;
    while (b->cap < n) b->cap *= 2;
    b->keys = (uint16_t *)realloc(b->keys, 2 * (size_t)b->cap);
    b->c = (oc_t *)realloc(b->c, sizeof(oc_t) * (size_t)b->cap);
}

static void bm_append(obm_t *b, uint16_t key, oc_t c) {
    bm_reserve(b, b->n + 1);
    b->keys[b->n] = key;
    b->c[b->n++] = c;
}

static void bm_insert(obm_t *b, int pos, uint16_t key, oc_t c) { /* roaring_array.c:348 */
    bm_reserve(b, b->n + 1);
    memmove(b->keys + pos + 1, b->keys + pos, 2 * (size_t)(b->n - pos));
    memmove(b->c + pos + 1, b->c + pos, sizeof(oc_t) * (size_t)(b->n - pos));
    b->keys[pos] = key;
    b->c[pos] = c;
    b->n++;
}

static void bm_remove(obm_t *b, int pos) {
    oc_free(&b->c[pos]);
    memmove(b->keys + pos, b->keys + pos + 1, 2 * (size_t)(b->n - pos - 1));
    memmove(b->c + pos, b->c + pos + 1, sizeof(oc_t) * (size_t)(b->n - pos - 1));
    b->n--;
}

static uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* ra_portable_deserialize, roaring_array.c:633-813 (format spec: roaring_array.c:469-531) */
static bool bm_parse(obm_t *b, const uint8_t *buf, size_t len) {
    if (len < 4) return false;
    uint32_t cookie = rd32(buf);
    size_t pos = 4;
    int32_t size;
    const uint8_t *runflags = NULL;
    bool hasrun = false;
    if ((cookie & 0xFFFF) == COOKIE_RUN) {
        size = (int32_t)(cookie >> 16) + 1;
        hasrun = true;
        runflags = buf + pos;
        pos += (size_t)(size + 7) / 8;
    } else if (cookie == COOKIE_NORUN) {
        if (len < 8) return false;
        size = (int32_t)rd32(buf + 4);
        pos = 8;
    } else {
        return false;
    }
    if (size < 0 || size > 65536) return false;
    bm_init(b, size);
    if (pos + 4 * (size_t)size > len) return false;
    const uint8_t *kc = buf + pos;
    pos += 4 * (size_t)size;
    if (!hasrun || size >= NO_OFFSET_THRESHOLD) pos += 4 * (size_t)size; /* offsets: skipped */
    for (int i = 0; i < size; i++) {
        uint16_t key = rd16(kc + 4 * i);
        int32_t card = (int32_t)rd16(kc + 4 * i + 2) + 1;
        bool isrun = hasrun && ((runflags[i / 8] >> (i % 8)) & 1);
        oc_t c = {0, 0, 0, NULL, NULL};
        if (isrun) {
            if (pos + 2 > len) return false;
            int32_t nr = rd16(buf + pos);
            pos += 2;
            if (pos + 4 * (size_t)nr > len) return false;
            c.type = T_RUN;
            c.n = nr;
            c.v = (uint16_t *)malloc(4 * (size_t)(nr ? nr : 1));
            for (int k = 0; k < 2 * nr; k++) c.v[k] = rd16(buf + pos + 2 * (size_t)k);
            pos += 4 * (size_t)nr;
        } else if (card > MAXARR) {
            if (pos + 8192 > len) return false;
            c.type = T_BITSET;
            c.card = card;
            c.w = (uint64_t *)malloc(WORDS * 8);
            for (int k = 0; k < WORDS; k++)
                c.w[k] = (uint64_t)rd32(buf + pos + 8 * (size_t)k) |
                         ((uint64_t)rd32(buf + pos + 8 * (size_t)k + 4) << 32);
            pos += 8192;
        } else {
            if (pos + 2 * (size_t)card > len) return false;
            c.type = T_ARRAY;
            c.card = c.n = card;
            c.v = (uint16_t *)malloc(2 * (size_t)card);
            for (int k = 0; k < card; k++) c.v[k] = rd16(buf + pos + 2 * (size_t)k);
            pos += 2 * (size_t)card;
        }
        bm_append(b, key, c);
    }
    return true;
}

static void wr16(uint8_t *p, uint16_t v) {
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
}
static void wr32(uint8_t *p, uint32_t v) {
    wr16(p, (uint16_t)v);
    wr16(p + 2, (uint16_t)(v >> 16));
}

static size_t oc_bytes(const oc_t *c) { /* containers.h:402-416 */
    if (c->type == T_BITSET) return 8192;
    if (c->type == T_ARRAY) return 2 * (size_t)c->n;
    return 2 + 4 * (size_t)c->n;
}

/* ra_portable_serialize, roaring_array.c:469-531 */
static size_t bm_serialize(const obm_t *b, uint8_t *out, size_t cap) {
    bool hasrun = false;
    size_t payload = 0;
    for (int i = 0; i < b->n; i++) {
        hasrun |= b->c[i].type == T_RUN;
        payload += oc_bytes(&b->c[i]);
    }
    size_t n = (size_t)b->n, hdr;
    if (hasrun)
        hdr = 4 + (n + 7) / 8 + (b->n < NO_OFFSET_THRESHOLD ? 4 * n : 8 * n);
    else
        hdr = 8 + 8 * n;
    size_t total = hdr + payload;
    if (total > cap || out == NULL) return total;
    uint8_t *p = out;
    if (hasrun) {
        wr32(p, COOKIE_RUN | ((uint32_t)(b->n - 1) << 16));
        p += 4;
        memset(p, 0, (n + 7) / 8);
        for (int i = 0; i < b->n; i++)
            if (b->c[i].type == T_RUN) p[i / 8] |= (uint8_t)(1 << (i % 8));
        p += (n + 7) / 8;
    } else {
        wr32(p, COOKIE_NORUN);
        wr32(p + 4, (uint32_t)b->n);
        p += 8;
    }
    for (int i = 0; i < b->n; i++) {
        wr16(p, b->keys[i]);
        wr16(p + 2, (uint16_t)(oc_card(&b->c[i]) - 1));
        p += 4;
    }
    if (!hasrun || b->n >= NO_OFFSET_THRESHOLD) {
        uint32_t off = (uint32_t)hdr;
        for (int i = 0; i < b->n; i++) {
            wr32(p, off);
            p += 4;
            off += (uint32_t)oc_bytes(&b->c[i]);
        }
    }
    for (int i = 0; i < b->n; i++) {
        const oc_t *c = &b->c[i];
        if (c->type == T_BITSET) {
            for (int k = 0; k < WORDS; k++) {
                wr32(p, (uint32_t)c->w[k]);
                wr32(p + 4, (uint32_t)(c->w[k] >> 32));
                p += 8;
            }
        } else if (c->type == T_ARRAY) {
            for (int k = 0; k < c->n; k++, p += 2) wr16(p, c->v[k]);
        } else {
            wr16(p, (uint16_t)c->n);
            p += 2;
            for (int k = 0; k < 2 * c->n; k++, p += 2) wr16(p, c->v[k]);
        }
    }
    return total;
}

/* roaring_bitmap_and :731, _or :877, _xor :1121, _andnot :1275 (all src/roaring.c) */
/* container_ior differs from container_or in one cell (containers.h:1226-1320): bitset|bitset
 * that saturates becomes the full run (OR_BITSET_CONVERSION_TO_FULL, :1234-1242), and
 * roaring_bitmap_or_inplace leaves a full left container untouched (roaring.c:1081-1083).
 * iand / ixor / iandnot apply the same type rules as their functional twins. */
static oc_t cell_ior(const oc_t *c1, const oc_t *c2) {
    if (oc_is_full(c1)) return oc_clone(c1);
    oc_t r = cell_or(c1, c2);
    if (c1->type == T_BITSET && c2->type == T_BITSET && r.type == T_BITSET && r.card == 65536) {
        oc_free(&r);
        oc_t f = {T_RUN, 0, 1, (uint16_t *)malloc(4), NULL};
        f.v[0] = 0;
        f.v[1] = 0xFFFF;
        return f;
    }
    return r;
}

static void bm_pair(int op, const obm_t *x1, const obm_t *x2, obm_t *ans) {
    const int inplace = op >= 4;
    op &= 3;
    bm_init(ans, x1->n + x2->n);
    int p1 = 0, p2 = 0;
    while (p1 < x1->n && p2 < x2->n) {
        uint16_t s1 = x1->keys[p1], s2 = x2->keys[p2];
        if (s1 == s2) {
            oc_t c;
            switch (op) {
                case ORC_AND: c = cell_and(&x1->c[p1], &x2->c[p2]); break;
                case ORC_OR: c = inplace ? cell_ior(&x1->c[p1], &x2->c[p2]) : cell_or(&x1->c[p1], &x2->c[p2]); break;
                case ORC_XOR: c = cell_xor(&x1->c[p1], &x2->c[p2]); break;
                default: c = cell_andnot(&x1->c[p1], &x2->c[p2]); break;
            }
            /* empties dropped: roaring.c:756-760, 1147-1151, 1308-1312 (OR cannot be empty) */
            if (oc_card(&c) > 0)
                bm_append(ans, s1, c);
            else
                oc_free(&c);
            p1++;
            p2++;
        } else if (s1 < s2) {
            if (op != ORC_AND) bm_append(ans, s1, oc_clone(&x1->c[p1])); /* pass-through */
            p1++;
        } else {
            if (op == ORC_OR || op == ORC_XOR) bm_append(ans, s2, oc_clone(&x2->c[p2]));
            p2++;
        }
    }
    if (op != ORC_AND)
        for (; p1 < x1->n; p1++) bm_append(ans, x1->keys[p1], oc_clone(&x1->c[p1]));
    if (op == ORC_OR || op == ORC_XOR)
        for (; p2 < x2->n; p2++) bm_append(ans, x2->keys[p2], oc_clone(&x2->c[p2]));
}

static void bm_copy(const obm_t *x, obm_t *ans) {
    bm_init(ans, x->n);
    for (int i = 0; i < x->n; i++) bm_append(ans, x->keys[i], oc_clone(&x->c[i]));
}

/* roaring_bitmap_lazy_or(x1,x2,bitsetconversion), roaring.c:2509-2598 */
static void bm_lazy_or(const obm_t *x1, const obm_t *x2, bool conv, obm_t *ans) {
    if (x1->n == 0) { bm_copy(x2, ans); return; }
    if (x2->n == 0) { bm_copy(x1, ans); return; }
    bm_init(ans, x1->n + x2->n);
    int p1 = 0, p2 = 0;
    while (p1 < x1->n && p2 < x2->n) {
        uint16_t s1 = x1->keys[p1], s2 = x2->keys[p2];
        if (s1 == s2) {
            const oc_t *c1 = &x1->c[p1], *c2 = &x2->c[p2];
            oc_t c;
            if (conv && c1->type != T_BITSET && c2->type != T_BITSET) { /* :2535-2545 */
                c = to_bitset(c1);
                cell_lazy_ior_bitset(&c, c2);
            } else {
                c = cell_lazy_or(c1, c2);
            }
            bm_append(ans, s1, c);
            p1++;
            p2++;
        } else if (s1 < s2) {
            bm_append(ans, s1, oc_clone(&x1->c[p1++]));
        } else {
            bm_append(ans, s2, oc_clone(&x2->c[p2++]));
        }
    }
    for (; p1 < x1->n; p1++) bm_append(ans, x1->keys[p1], oc_clone(&x1->c[p1]));
    for (; p2 < x2->n; p2++) bm_append(ans, x2->keys[p2], oc_clone(&x2->c[p2]));
}

/* roaring_bitmap_lazy_or_inplace(x1,x2,bitsetconversion), roaring.c:2600-2682 */
static void bm_lazy_or_inplace(obm_t *x1, const obm_t *x2, bool conv) {
    if (x2->n == 0) return;
    if (x1->n == 0) {
        bm_free(x1);
        bm_copy(x2, x1);
        return;
    }
    int p1 = 0, p2 = 0;
    while (p1 < x1->n && p2 < x2->n) {
        uint16_t s1 = x1->keys[p1], s2 = x2->keys[p2];
        if (s1 == s2) {
            oc_t *c1 = &x1->c[p1];
            if (!oc_is_full(c1)) { /* :2621 */
                if (conv && c1->type != T_BITSET) { /* :2622-2633 */
                    oc_t b = to_bitset(c1);
                    oc_free(c1);
                    *c1 = b;
                }
                cell_lazy_ior(c1, &x2->c[p2]);
            }
            p1++;
            p2++;
        } else if (s1 < s2) {
            p1++;
        } else {
            bm_insert(x1, p1, s2, oc_clone(&x2->c[p2])); /* :2669 */
            p1++;
            p2++;
        }
    }
    for (; p2 < x2->n; p2++) bm_append(x1, x2->keys[p2], oc_clone(&x2->c[p2]));
}

/* roaring_bitmap_lazy_xor, roaring.c:2684-2761 */
static void bm_lazy_xor(const obm_t *x1, const obm_t *x2, obm_t *ans) {
    if (x1->n == 0) { bm_copy(x2, ans); return; }
    if (x2->n == 0) { bm_copy(x1, ans); return; }
    bm_init(ans, x1->n + x2->n);
    int p1 = 0, p2 = 0;
    while (p1 < x1->n && p2 < x2->n) {
        uint16_t s1 = x1->keys[p1], s2 = x2->keys[p2];
        if (s1 == s2) {
            oc_t c = cell_lazy_xor(&x1->c[p1], &x2->c[p2]);
            if (oc_card(&c) > 0)
                bm_append(ans, s1, c);
            else
                oc_free(&c);
            p1++;
            p2++;
        } else if (s1 < s2) {
            bm_append(ans, s1, oc_clone(&x1->c[p1++]));
        } else {
            bm_append(ans, s2, oc_clone(&x2->c[p2++]));
        }
    }
    for (; p1 < x1->n; p1++) bm_append(ans, x1->keys[p1], oc_clone(&x1->c[p1]));
    for (; p2 < x2->n; p2++) bm_append(ans, x2->keys[p2], oc_clone(&x2->c[p2]));
}

/* roaring_bitmap_lazy_xor_inplace, roaring.c:2763-2843 with container_lazy_ixor,
 * containers.h:1749-1776 (only B,B is lazy; every other cell fixes the dirty cardinality
 * and runs the eager container_ixor) */
static void bm_lazy_xor_inplace(obm_t *x1, const obm_t *x2) {
    if (x2->n == 0) return;
    if (x1->n == 0) {
        bm_free(x1);
        bm_copy(x2, x1);
        return;
    }
    int p1 = 0, p2 = 0;
    while (p1 < x1->n && p2 < x2->n) {
        uint16_t s1 = x1->keys[p1], s2 = x2->keys[p2];
        if (s1 == s2) {
            oc_t *c1 = &x1->c[p1];
            const oc_t *c2 = &x2->c[p2];
            oc_t c;
            if (c1->type == T_BITSET && c2->type == T_BITSET) {
                c = oc_clone(c1);
                for (int i = 0; i < WORDS; i++) c.w[i] ^= c2->w[i];
                c.card = -1;
            } else {
                if (c1->type == T_BITSET && c1->card < 0) c1->card = words_card(c1->w);
                c = cell_xor(c1, c2);
            }
            if (oc_card(&c) > 0) {
                oc_free(c1);
                *c1 = c;
                p1++;
            } else {
                oc_free(&c);
                bm_remove(x1, p1);
            }
            p2++;
        } else if (s1 < s2) {
            p1++;
        } else {
            bm_insert(x1, p1, s2, oc_clone(&x2->c[p2]));
            p1++;
            p2++;
        }
    }
    if (p1 == x1->n)
        for (; p2 < x2->n; p2++) bm_append(x1, x2->keys[p2], oc_clone(&x2->c[p2]));
}

/* ------------------------------------------------------------------ public */
size_t oracle_pair_op(int op, const uint8_t *a, size_t na, const uint8_t *b, size_t nb,
                      uint8_t *out, size_t cap) {
    obm_t x1, x2, ans;
    bool ok1 = bm_parse(&x1, a, na);
    bool ok2 = ok1 && bm_parse(&x2, b, nb);
    if (!ok1 || !ok2) return (size_t)-1;
    /* roaring.c:1279-1288: andnot with an empty side short-circuits (same result set) */
    bm_pair(op, &x1, &x2, &ans);
    size_t r = bm_serialize(&ans, out, cap);
    bm_free(&x1);
    bm_free(&x2);
    bm_free(&ans);
    return r;
}

/* roaring_bitmap_or_many roaring.c:775-790 / roaring_bitmap_xor_many roaring.c:795-809 */
size_t oracle_many_op(int op, size_t n, const uint8_t *const *bufs, const size_t *lens,
                      uint8_t *out, size_t cap) {
    obm_t ans;
    if (n == 0) {
        bm_init(&ans, 0);
    } else {
        obm_t *xs = (obm_t *)malloc(sizeof(obm_t) * n);
        for (size_t i = 0; i < n; i++)
            if (!bm_parse(&xs[i], bufs[i], lens[i])) return (size_t)-1;
        if (n == 1) {
            bm_copy(&xs[0], &ans);
        } else {
            if (op == ORC_OR_MANY)
                bm_lazy_or(&xs[0], &xs[1], true, &ans);
            else
                bm_lazy_xor(&xs[0], &xs[1], &ans);
            for (size_t i = 2; i < n; i++) {
                if (op == ORC_OR_MANY)
                    bm_lazy_or_inplace(&ans, &xs[i], true);
                else
                    bm_lazy_xor_inplace(&ans, &xs[i]);
            }
            for (int i = 0; i < ans.n; i++) repair(&ans.c[i]); /* roaring.c:2845 */
        }
        for (size_t i = 0; i < n; i++) bm_free(&xs[i]);
        free(xs);
    }
    size_t r = bm_serialize(&ans, out, cap);
    bm_free(&ans);
    return r;
}

/* roaring_bitmap_and_cardinality, roaring.c:3048-3076 */
uint64_t oracle_and_cardinality(const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    obm_t x1, x2;
    if (!bm_parse(&x1, a, na)) return (uint64_t)-1;
    if (!bm_parse(&x2, b, nb)) return (uint64_t)-1;
    uint64_t s = 0;
    int p1 = 0, p2 = 0;
    while (p1 < x1.n && p2 < x2.n) {
        if (x1.keys[p1] == x2.keys[p2])
            s += (uint64_t)cell_and_card(&x1.c[p1++], &x2.c[p2++]);
        else if (x1.keys[p1] < x2.keys[p2])
            p1++;
        else
            p2++;
    }
    bm_free(&x1);
    bm_free(&x2);
    return s;
}

uint64_t oracle_cardinality(const uint8_t *a, size_t na) { /* roaring.c:1436 */
    obm_t x;
    if (!bm_parse(&x, a, na)) return (uint64_t)-1;
    uint64_t s = 0;
    for (int i = 0; i < x.n; i++) s += (uint64_t)oc_card(&x.c[i]);
    bm_free(&x);
    return s;
}

/* A left fold of the public lazy API followed by one repair:
 *   acc = lazy_op(x0, x1[, conv]); acc = lazy_op_inplace(acc, xi[, conv]) for i >= 2;
 *   roaring_bitmap_repair_after_lazy(acc).
 * op 0 = roaring_bitmap_lazy_or / _lazy_or_inplace (roaring.c:2509-2682, `conv` = their
 * bitsetconversion argument), op 1 = roaring_bitmap_lazy_xor / _lazy_xor_inplace
 * (roaring.c:2684-2843).  n == 1: repair(copy(x0)). */
size_t oracle_lazy_fold(int op, int conv, size_t n, const uint8_t *const *bufs, const size_t *lens,
                        uint8_t *out, size_t cap) {
    obm_t ans;
    if (n == 0) return (size_t)-1;
    obm_t *xs = (obm_t *)malloc(sizeof(obm_t) * n);
    for (size_t i = 0; i < n; i++)
        if (!bm_parse(&xs[i], bufs[i], lens[i])) return (size_t)-1;
    if (n == 1) {
        bm_copy(&xs[0], &ans);
    } else {
        if (op == 0) bm_lazy_or(&xs[0], &xs[1], conv != 0, &ans);
        else bm_lazy_xor(&xs[0], &xs[1], &ans);
        for (size_t i = 2; i < n; i++) {
            if (op == 0) bm_lazy_or_inplace(&ans, &xs[i], conv != 0);
            else bm_lazy_xor_inplace(&ans, &xs[i]);
        }
    }
    for (int i = 0; i < ans.n; i++) repair(&ans.c[i]); /* roaring.c:2845 */
    for (size_t i = 0; i < n; i++) bm_free(&xs[i]);
    free(xs);
    size_t r = bm_serialize(&ans, out, cap);
    bm_free(&ans);
    return r;
}

/* roaring_bitmap_portable_size_in_bytes (roaring.c:1490 -> roaring_array.c:469-500) of a bitmap
 * that may be in a lazy state: arrays count 2*card, bitsets 8192, runs 2+4*n_runs. */
static uint64_t bm_portable_size(const obm_t *b) {
    bool hasrun = false;
    uint64_t sz = 0;
    for (int i = 0; i < b->n; i++) {
        if (b->c[i].type == T_RUN) hasrun = true;
        sz += oc_bytes(&b->c[i]);
    }
    if (hasrun) {
        sz += 4 + (uint64_t)(b->n + 7) / 8 + 4ull * b->n;
        if (b->n >= NO_OFFSET_THRESHOLD) sz += 4ull * b->n;
    } else {
        sz += 4 + 4 + 8ull * b->n;
    }
    return sz;
}

/* lazy_or_from_lazy_inputs, roaring_priority_queue.c:99-181: container_lazy_ior on every
 * matched key (no full-container short cut), unmatched containers moved.  Consumes x1, x2. */
static void bm_lazy_or_from_lazy(obm_t *x1, obm_t *x2, obm_t *ans) {
    if (x1->n == 0) { bm_free(x1); *ans = *x2; return; }
    if (x2->n == 0) { bm_free(x2); *ans = *x1; return; }
    bm_init(ans, x1->n + x2->n);
    int p1 = 0, p2 = 0;
    while (p1 < x1->n && p2 < x2->n) {
        uint16_t s1 = x1->keys[p1], s2 = x2->keys[p2];
        if (s1 == s2) {
            oc_t *c1 = &x1->c[p1], *c2 = &x2->c[p2];
            if (c2->type == T_BITSET && c1->type != T_BITSET) { /* :133-139 operands swapped */
                cell_lazy_ior(c2, c1);
                bm_append(ans, s1, oc_clone(c2));
            } else {
                cell_lazy_ior(c1, c2);
                bm_append(ans, s1, oc_clone(c1));
            }
            p1++;
            p2++;
        } else if (s1 < s2) {
            bm_append(ans, s1, oc_clone(&x1->c[p1++]));
        } else {
            bm_append(ans, s2, oc_clone(&x2->c[p2++]));
        }
    }
    for (; p1 < x1->n; p1++) bm_append(ans, x1->keys[p1], oc_clone(&x1->c[p1]));
    for (; p2 < x2->n; p2++) bm_append(ans, x2->keys[p2], oc_clone(&x2->c[p2]));
    bm_free(x1);
    bm_free(x2);
}

/* The binary min-heap of roaring_priority_queue.c:12-97 (ordering by portable size, ties
 * resolved by heap position exactly as the reference's sift loops do). */
typedef struct {
    uint64_t size;
    bool temp;
    obm_t bm;
} hp_t;

static void hp_down(hp_t *e, uint32_t n, uint32_t i) { /* percolate_down :46-66 */
    uint32_t half = n >> 1;
    hp_t cur = e[i];
    while (i < half) {
        uint32_t child = 2 * i + 1;
        if (child + 1 < n && e[child + 1].size < e[child].size) child++;
        if (!(e[child].size < cur.size)) break;
        e[i] = e[child];
        i = child;
    }
    e[i] = cur;
}
static void hp_push(hp_t *e, uint32_t *n, hp_t t) { /* pq_add :31-42 */
    uint32_t i = (*n)++;
    while (i > 0) {
        uint32_t parent = (i - 1) >> 1;
        if (!(t.size < e[parent].size)) break;
        e[i] = e[parent];
        i = parent;
    }
    e[i] = t;
}
static hp_t hp_pop(hp_t *e, uint32_t *n) { /* pq_poll :84-95 */
    hp_t top = e[0];
    if (*n > 1) {
        e[0] = e[--(*n)];
        hp_down(e, *n, 0);
    } else {
        --(*n);
    }
    return top;
}

/* roaring_bitmap_or_many_heap, roaring_priority_queue.c:200-250 */
size_t oracle_or_many_heap(size_t n, const uint8_t *const *bufs, const size_t *lens, uint8_t *out,
                           size_t cap) {
    obm_t ans;
    if (n == 0) {
        bm_init(&ans, 0);
    } else {
        hp_t *e = (hp_t *)malloc(sizeof(hp_t) * n);
        for (size_t i = 0; i < n; i++) {
            if (!bm_parse(&e[i].bm, bufs[i], lens[i])) return (size_t)-1;
            e[i].temp = false;
            e[i].size = bm_portable_size(&e[i].bm);
        }
        if (n == 1) {
            bm_copy(&e[0].bm, &ans);
            bm_free(&e[0].bm);
        } else {
            uint32_t cnt = (uint32_t)n;
            for (int32_t i = (int32_t)(cnt >> 1); i >= 0; i--) hp_down(e, cnt, (uint32_t)i); /* create_pq :68-82 */
            while (cnt > 1) {
                hp_t a = hp_pop(e, &cnt), b = hp_pop(e, &cnt), r;
                r.temp = true;
                if (a.temp && b.temp) {
                    bm_lazy_or_from_lazy(&a.bm, &b.bm, &r.bm);
                } else if (b.temp) {
                    bm_lazy_or_inplace(&b.bm, &a.bm, false);
                    r.bm = b.bm;
                    bm_free(&a.bm);
                } else if (a.temp) {
                    bm_lazy_or_inplace(&a.bm, &b.bm, false);
                    r.bm = a.bm;
                    bm_free(&b.bm);
                } else {
                    bm_lazy_or(&a.bm, &b.bm, false, &r.bm);
                    bm_free(&a.bm);
                    bm_free(&b.bm);
                }
                r.size = bm_portable_size(&r.bm);
                hp_push(e, &cnt, r);
            }
            ans = hp_pop(e, &cnt).bm;
            for (int i = 0; i < ans.n; i++) repair(&ans.c[i]);
        }
        free(e);
    }
    size_t r = bm_serialize(&ans, out, cap);
    bm_free(&ans);
    return r;
}

/* roaring_bitmap_flip (roaring.c:2289-2349): negation of [range_start, range_end) — containers
 * outside the range are copied, keys inside it go through container_not_range / container_not
 * (containers.h:2009-2073; mixed_negation.c) or, when absent, container_range_of_ones
 * (containers.h:300-312).  Type rules: bitset / array input -> array if card <= 4096 else bitset
 * (mixed_negation.c:79-160, 182-203; a fully negated array is always a bitset, :26-37, which the
 * same rule yields); run input -> convert_run_to_efficient_container (:226-256); empty results
 * are dropped (roaring.c:2199-2204). */
size_t oracle_flip(const uint8_t *a, size_t na, uint64_t range_start, uint64_t range_end, uint8_t *out,
                   size_t cap) {
    obm_t x, ans;
    if (!bm_parse(&x, a, na)) return (size_t)-1;
    /* :2292-2294, and flip_closed's own guard on the truncated 32-bit ends (:2303-2305) */
    if (range_start >= range_end || range_start > (uint64_t)0xFFFFFFFFu + 1 ||
        (uint32_t)range_start > (uint32_t)(range_end - 1)) {
        size_t r0 = bm_serialize(&x, out, cap);
        bm_free(&x);
        return r0;
    }
    const uint32_t first = (uint32_t)range_start, last = (uint32_t)(range_end - 1);
    const uint32_t hb0 = first >> 16, hb1 = last >> 16;
    bm_init(&ans, x.n + (int)(hb1 - hb0) + 1);
    int p = 0;
    for (; p < x.n && x.keys[p] < hb0; p++) bm_append(&ans, x.keys[p], oc_clone(&x.c[p]));
    for (uint32_t hb = hb0; hb <= hb1; hb++) {
        const uint32_t lo = hb == hb0 ? (first & 0xFFFF) : 0, hi = hb == hb1 ? (last & 0xFFFF) : 0xFFFF;
        if (p < x.n && x.keys[p] == hb) {
            uint64_t w[WORDS], rng[WORDS];
            to_words(&x.c[p], w);
            memset(rng, 0, sizeof(rng));
            words_set_range(rng, lo, hi);
            for (int i = 0; i < WORDS; i++) w[i] ^= rng[i];
            oc_t c = x.c[p].type == T_RUN ? eff_from_words(w) : ab_from_words(w);
            if (oc_card(&c) > 0) bm_append(&ans, (uint16_t)hb, c);
            else oc_free(&c);
            p++;
        } else { /* container_range_of_ones: one value -> array, else a run */
            oc_t c;
            if (hi == lo) {
                c.type = T_ARRAY; c.card = 0; c.n = 1; c.w = NULL;
                c.v = (uint16_t *)malloc(2);
                c.v[0] = (uint16_t)lo;
            } else {
                c.type = T_RUN; c.card = 0; c.n = 1; c.w = NULL;
                c.v = (uint16_t *)malloc(4);
                c.v[0] = (uint16_t)lo;
                c.v[1] = (uint16_t)(hi - lo);
            }
            bm_append(&ans, (uint16_t)hb, c);
        }
    }
    for (; p < x.n; p++) bm_append(&ans, x.keys[p], oc_clone(&x.c[p]));
    size_t r = bm_serialize(&ans, out, cap);
    bm_free(&x);
    bm_free(&ans);
    return r;
}

/* ------------------------------------------------------------------ 64-bit bitmaps
 * roaring64_bitmap_{and,or,xor,andnot} (src/roaring64.c:1332, 1541, 1663, 1809) walk two ARTs in
 * high-48-bit key order and apply the SAME container cells (container_and / _or / _xor / _andnot)
 * to equal keys; the portable format (roaring64.c:2262-2395) groups the containers of one high-32
 * value into an ordinary 32-bit portable bitmap:  u64 n_buckets, then per bucket u32 high32 +
 * 32-bit blob.  So a 64-bit op is the 32-bit op per matching bucket, unmatched buckets passing
 * through (or / xor: both sides, andnot: left side), buckets left without containers dropped. */
typedef struct {
    uint64_t n;
    uint32_t *high;
    obm_t *bm;
} o64_t;

static void o64_free(o64_t *x) {
    for (uint64_t i = 0; i < x->n; i++) bm_free(&x->bm[i]);
    free(x->high);
    free(x->bm);
}
static bool o64_parse(o64_t *x, const uint8_t *buf, size_t len) {
    if (len < 8) return false;
    uint64_t n = (uint64_t)rd32(buf) | ((uint64_t)rd32(buf + 4) << 32);
    if (n > 0xFFFFFFFFull) return false;
    x->n = 0;
    x->high = (uint32_t *)malloc(4 * (size_t)(n ? n : 1));
    x->bm = (obm_t *)malloc(sizeof(obm_t) * (size_t)(n ? n : 1));
    size_t pos = 8;
    for (uint64_t i = 0; i < n; i++) {
        if (pos + 4 > len) return false;
        x->high[i] = rd32(buf + pos);
        pos += 4;
        if (!bm_parse(&x->bm[i], buf + pos, len - pos)) return false;
        x->n = i + 1;
        pos += bm_serialize(&x->bm[i], NULL, 0);
        if (pos > len) return false;
    }
    return true;
}

size_t oracle_r64_pair_op(int op, const uint8_t *a, size_t na, const uint8_t *b, size_t nb, uint8_t *out,
                          size_t cap) {
    o64_t x1, x2;
    if (!o64_parse(&x1, a, na) || !o64_parse(&x2, b, nb)) return (size_t)-1;
    obm_t empty;
    bm_init(&empty, 0);
    size_t pos = 8;
    uint64_t kept = 0, p1 = 0, p2 = 0;
    while (p1 < x1.n || p2 < x2.n) {
        const bool only1 = p2 >= x2.n || (p1 < x1.n && x1.high[p1] < x2.high[p2]);
        const bool only2 = !only1 && (p1 >= x1.n || x2.high[p2] < x1.high[p1]);
        const uint32_t high = only2 ? x2.high[p2] : x1.high[p1];
        obm_t ans;
        bm_pair(op, only2 ? &empty : &x1.bm[p1], only1 ? &empty : &x2.bm[p2], &ans);
        if (!only2) p1++;
        if (!only1) p2++;
        if (ans.n > 0) {
            const size_t sz = bm_serialize(&ans, NULL, 0);
            if (out && pos + 4 + sz <= cap) {
                wr32(out + pos, high);
                bm_serialize(&ans, out + pos + 4, sz);
            }
            pos += 4 + sz;
            kept++;
        }
        bm_free(&ans);
    }
    if (out && cap >= 8) {
        wr32(out, (uint32_t)kept);
        wr32(out + 4, (uint32_t)(kept >> 32));
    }
    o64_free(&x1);
    o64_free(&x2);
    bm_free(&empty);
    return pos;
}

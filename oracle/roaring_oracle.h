/* oracle/roaring_oracle.h — TEST INFRASTRUCTURE ONLY.  See roaring_oracle.c. */
#ifndef ROARING_ORACLE_H
#define ROARING_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_AND = 0, ORC_OR = 1, ORC_XOR = 2, ORC_ANDNOT = 3 };
enum { ORC_OR_MANY = 0, ORC_XOR_MANY = 1 };

/* Pairwise op on two portable-serialized bitmaps.  Writes the portable serialization of the
 * result (as the reference would produce it: same keys, container types, payloads) to `out`
 * when it fits in `cap`; always returns the size needed.  Returns (size_t)-1 on a malformed
 * input. */
size_t oracle_pair_op(int op, const uint8_t *a, size_t na, const uint8_t *b, size_t nb,
                      uint8_t *out, size_t cap);

/* roaring_bitmap_or_many / roaring_bitmap_xor_many over n serialized bitmaps. */
size_t oracle_many_op(int op, size_t n, const uint8_t *const *bufs, const size_t *lens,
                      uint8_t *out, size_t cap);

/* repair_after_lazy(fold of the public lazy API): op 0 = lazy_or/lazy_or_inplace with
 * bitsetconversion = conv, op 1 = lazy_xor/lazy_xor_inplace; inputs folded left to right. */
size_t oracle_lazy_fold(int op, int conv, size_t n, const uint8_t *const *bufs, const size_t *lens,
                        uint8_t *out, size_t cap);

/* roaring_bitmap_or_many_heap (roaring_priority_queue.c:200-250). */
size_t oracle_or_many_heap(size_t n, const uint8_t *const *bufs, const size_t *lens, uint8_t *out,
                           size_t cap);

/* roaring_bitmap_flip(x, range_start, range_end) (roaring.c:2289). */
size_t oracle_flip(const uint8_t *a, size_t na, uint64_t range_start, uint64_t range_end, uint8_t *out,
                   size_t cap);

/* roaring64_bitmap_{and,or,xor,andnot} on the 64-bit portable format (roaring64.c:1332-1895,
 * 2262-2395); op codes 0-3 as oracle_pair_op. */
size_t oracle_r64_pair_op(int op, const uint8_t *a, size_t na, const uint8_t *b, size_t nb, uint8_t *out,
                          size_t cap);

/* roaring_bitmap_and_cardinality; (uint64_t)-1 on malformed input. */
uint64_t oracle_and_cardinality(const uint8_t *a, size_t na, const uint8_t *b, size_t nb);

/* roaring_bitmap_get_cardinality of a serialized bitmap. */
uint64_t oracle_cardinality(const uint8_t *a, size_t na);

#ifdef __cplusplus
}
#endif
#endif

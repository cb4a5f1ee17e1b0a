/* TEST INFRASTRUCTURE ONLY (oracle/_ref build recipe).
 * The reference under-allocates two scratch buffers: DCz is malloc(nblk*sizeof(float)) and
 * bin_indexz is malloc(N) (dctz-comp-lib.c:246,258; dctz-decomp-lib.c:108,120) yet both receive up to
 * compressBound() bytes of deflate output / h.*_sz_compressed bytes of memcpy, which overflows
 * the heap whenever a section does not compress (always for tiny inputs).  The sources stay
 * unmodified; the _ref link line uses -Wl,--wrap=malloc so every malloc issued by the reference
 * objects gets slack, which makes the build safe to call in-process from the test-suite. */
#include <stddef.h>
void *__real_malloc(size_t n);
void *__wrap_malloc(size_t n) { return __real_malloc(n + n / 32 + 4096); }
/* gcc -O3 folds the reference's malloc()+memset(0) pairs into calloc(): pad that too. */
void *__real_calloc(size_t n, size_t s);
void *__wrap_calloc(size_t n, size_t s) { return __real_calloc(n + n / 32 + 4096 / (s ? s : 1) + 1, s); }

#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
/* TEST TOOLING: dumps the arrays hmmsearch sorts with 24-byte elements (p7_spensemble_Cluster's sigc list:
 * struct p7_spcoord_s { int idx, i, j, k, m; float prob; }) to the file named by QSORT_TRACE. */
static void (*real_qsort)(void *, size_t, size_t, int (*)(const void *, const void *));
static FILE *out;
void qsort(void *base, size_t n, size_t sz, int (*cmp)(const void *, const void *)) {
    if (!real_qsort) { real_qsort = dlsym(RTLD_NEXT, "qsort"); const char *p = getenv("QSORT_TRACE"); out = p ? fopen(p, "w") : NULL; }
    if (out && sz == 24) {
        fprintf(out, "SIGC %zu", n);
        for (size_t a = 0; a < n; a++) {
            const int *e = (const int *)((const char *)base + a * 24);
            fprintf(out, " %d,%d,%d,%d,%d,%.4f", e[0], e[1], e[2], e[3], e[4], *(const float *)(e + 5));
        }
        fprintf(out, "\n"); fflush(out);
    }
    real_qsort(base, n, sz, cmp);
}

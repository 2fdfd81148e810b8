/* Forwards the two library calls of the reference's lapacke_dpotrf.c to scipy's OpenBLAS.
 *   CHOL_REF_DUMP=<prefix>   write <prefix>.in / <prefix>.out (raw doubles, n*n) around dpotrf
 *   CHOL_REF_SKIP_RESIDUAL=1 skip the N^3 dgemm of the reference's (defective)
 *                            residual check; the timed region (dpotrf only) is unchanged.  */
#include <stdio.h>
#include <stdlib.h>
#include "cblas.h"
#include "lapacke.h"
extern int scipy_LAPACKE_dpotrf(int, char, int, double*, int);
extern void scipy_cblas_dgemm(int, int, int, int, int, int, double, const double*, int, const double*, int, double,
                              double*, int);
static void dump(const char* prefix, const char* suffix, const double* a, size_t n) {
    char path[4096];
    snprintf(path, sizeof path, "%s.%s", prefix, suffix);
    FILE* f = fopen(path, "wb");
    if (!f) return;
    fwrite(a, sizeof(double), n, f);
    fclose(f);
}
int LAPACKE_dpotrf(int layout, char uplo, int n, double* a, int lda) {
    const char* d = getenv("CHOL_REF_DUMP");
    if (d) dump(d, "in", a, (size_t)n * lda);
    int info = scipy_LAPACKE_dpotrf(layout, uplo, n, a, lda);
    if (d) dump(d, "out", a, (size_t)n * lda);
    return info;
}
void cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 double alpha, const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc) {
    if (getenv("CHOL_REF_SKIP_RESIDUAL")) return;
    scipy_cblas_dgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}

/* Shim so the reference's own CPU program (lapacke_dpotrf.c) compiles, unmodified, against
 * the OpenBLAS bundled in the scipy wheel (symbols prefixed scipy_).  Test infrastructure. */
#ifndef CHOL_REF_SHIM_CBLAS_H
#define CHOL_REF_SHIM_CBLAS_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
void cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 double alpha, const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc);
#endif

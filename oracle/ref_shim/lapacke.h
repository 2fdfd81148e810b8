/* Shim header, see cblas.h in this directory. */
#ifndef CHOL_REF_SHIM_LAPACKE_H
#define CHOL_REF_SHIM_LAPACKE_H
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
int LAPACKE_dpotrf(int matrix_layout, char uplo, int n, double* a, int lda);
#endif

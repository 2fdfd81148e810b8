/*
 * chol_oracle.c — CPU restatement of the reference's tile Cholesky path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (libchol_b200.so) never does and has no CPU fallback.
 *
 * What it restates (paths under /root/reference):
 *   W2 = cholesky_armonik/w_c_cons_v2/worker_construction2/src/worker_distrib.cpp
 *   C1 = cholesky_armonik/w_c_cons_v1/client_construction/client/src/client_distrib.cpp
 *   C2 = cholesky_armonik/w_c_cons_v2/client_construction2/client/src/client_distrib.cpp
 *   V6 = Cholesky_chameleon_VM/cho/docker_installation_and_bench_files/v6_test.c
 *   LP = Cholesky_chameleon_VM/cho/Cholesky_Chameleon_sauv/code_c/lapacke_dpotrf.c
 *
 * The arithmetic of the reference lives in third-party libraries that are NOT under
 * /root/reference: Chameleon (gitlab.inria.fr/solverstack/chameleon, unpinned HEAD,
 * Dockerfile.worker.v4:60) calling OpenBLAS v0.3.26 (Dockerfile.worker.v4:24) dpotrf /
 * dtrsm / dsyrk / dgemm.  Their published semantics (LAPACK/BLAS reference definitions)
 * are restated below as plain loops; the call sites fix the parameters:
 *   POTRF  W2:238  dpotrf(Lower)                       -> oracle_potrf_tile
 *   TRSM   W2:323  dtrsm(Right,Lower,Trans,NonUnit,1)  -> oracle_trsm_tile
 *   SYRK   W2:416  dsyrk(Lower,NoTrans,-1,A,1,C)       -> oracle_syrk_tile
 *   GEMM   W2:511  dgemm(NoTrans,Trans,-1,Ai,Aj,1,C)   -> oracle_gemm_tile
 *   DAG    C1:278-333 (wave loop)                       -> oracle_potrf_tiled
 *   whole  V6:56 / LP:54 dpotrf(Lower) of the matrix    -> oracle_potrf (= tile op with b=N)
 *   gen    V6:46 dplgsy(bump, Lower, seed)              -> oracle_plgsy (algorithm of Chameleon's
 *          coreblas core_dplgsy.c, recalled, not in /root/reference: 64-bit LCG with jump-ahead)
 *          C2:224-264 make_spd_like_chameleon lives in oracle.py (mt19937_64)
 *          LP:35-45 rand()-based SPD matrix              -> oracle_lp_matrix
 *
 * Pinning: the reference has no tests / golden vectors for this path.  The
 * oracle is pinned against the output of the reference's own CPU program LP compiled here
 * (oracle/Makefile -> oracle/_ref/), see oracle/pin_against_ref.py and
 * tests/golden/ref_lapacke_dpotrf.json.
 *
 * All matrices column-major FP64.  Compile with -ffp-contract=off so the generator is
 * bit-identical with the CUDA one.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define A_(p, ld, i, j) ((p)[(size_t)(j) * (size_t)(ld) + (size_t)(i)])

/* ---- POTRF: LAPACK dpotrf('L') semantics (dpotf2 column sweep).  Only the lower triangle is
 * read or written; returns info = 0 or the 1-based index of the first non-positive pivot. */
int oracle_potrf_tile(int n, double* A, int lda) {
    for (int j = 0; j < n; ++j) {
        double ajj = A_(A, lda, j, j);
        for (int k = 0; k < j; ++k) ajj -= A_(A, lda, j, k) * A_(A, lda, j, k);
        if (!(ajj > 0.0)) {
            A_(A, lda, j, j) = ajj;
            return j + 1;
        }
        ajj = sqrt(ajj);
        A_(A, lda, j, j) = ajj;
        /* column j below the diagonal: A[i][j] = (A[i][j] - sum_k A[i][k] A[j][k]) / ajj,
         * accumulated column-of-k by column-of-k so memory is walked contiguously */
        for (int k = 0; k < j; ++k) {
            const double ljk = A_(A, lda, j, k);
            const double* ck = &A_(A, lda, 0, k);
            double* cj = &A_(A, lda, 0, j);
            for (int i = j + 1; i < n; ++i) cj[i] -= ck[i] * ljk;
        }
        double* cj = &A_(A, lda, 0, j);
        for (int i = j + 1; i < n; ++i) cj[i] /= ajj;
    }
    return 0;
}

/* ---- TRSM: A <- A * L^{-T}  (Right, Lower, Trans, NonUnit, alpha = 1), A is m x n, L n x n.
 * Reference BLAS dtrsm loop for this case: for each column k of L in order,
 * B[:,k] /= L[k][k]; B[:,j] -= L[j][k] * B[:,k] for j > k. */
void oracle_trsm_tile(int m, int n, const double* L, int ldl, double* A, int lda) {
    for (int k = 0; k < n; ++k) {
        const double d = A_(L, ldl, k, k);
        double* bk = &A_(A, lda, 0, k);
        for (int i = 0; i < m; ++i) bk[i] /= d;
        for (int j = k + 1; j < n; ++j) {
            const double ljk = A_(L, ldl, j, k);
            if (ljk != 0.0) {
                double* bj = &A_(A, lda, 0, j);
                for (int i = 0; i < m; ++i) bj[i] -= ljk * bk[i];
            }
        }
    }
}

/* ---- SYRK: C <- C - A*A^T, lower triangle only (Lower, NoTrans, alpha=-1, beta=1). */
void oracle_syrk_tile(int n, int k, const double* A, int lda, double* C, int ldc) {
    for (int l = 0; l < k; ++l) {
        const double* al = &A_(A, lda, 0, l);
        for (int j = 0; j < n; ++j) {
            const double t = al[j];
            double* cj = &A_(C, ldc, 0, j);
            for (int i = j; i < n; ++i) cj[i] -= al[i] * t;
        }
    }
}

/* ---- GEMM: C <- C - Ai*Aj^T (NoTrans, Trans, alpha=-1, beta=1); C m x n, Ai m x k, Aj n x k. */
void oracle_gemm_tile(int m, int n, int k, const double* Ai, int ldai, const double* Aj, int ldaj, double* C,
                      int ldc) {
    for (int l = 0; l < k; ++l) {
        const double* al = &A_(Ai, ldai, 0, l);
        const double* bl = &A_(Aj, ldaj, 0, l);
        for (int j = 0; j < n; ++j) {
            const double t = bl[j];
            double* cj = &A_(C, ldc, 0, j);
            for (int i = 0; i < m; ++i) cj[i] -= al[i] * t;
        }
    }
}

/* ---- tile DAG in the order of the client's wave loop (C1:278-333; C2:506-565):
 *   for k: POTRF(k,k); TRSM(i,k) i>k; for i>k, k<j<=i: SYRK(i,i) if i==j else GEMM(i,j).
 * `tiles` = array of nt*(nt+1)/2 tile pointers, tile (i,j), i>=j, at index i*(i+1)/2 + j;
 * every tile b x b column-major ld=b (W2:76-79).  Returns LAPACK info (global index). */
int oracle_potrf_tiled(int nt, int b, double** tiles) {
#define T_(i, j) tiles[(size_t)(i) * ((i) + 1) / 2 + (j)]
    for (int k = 0; k < nt; ++k) {
        int info = oracle_potrf_tile(b, T_(k, k), b);
        if (info) return k * b + info;
        for (int i = k + 1; i < nt; ++i) oracle_trsm_tile(b, b, T_(k, k), b, T_(i, k), b);
        for (int i = k + 1; i < nt; ++i)
            for (int j = k + 1; j <= i; ++j) {
                if (i == j) oracle_syrk_tile(b, b, T_(i, k), b, T_(i, i), b);
                else oracle_gemm_tile(b, b, b, T_(i, k), b, T_(j, k), b, T_(i, j), b);
            }
    }
    return 0;
#undef T_
}

/* ---- dplgsy-like generator (V6:46): 64-bit LCG, counter based. -------------------------- */
#define RND64_A 6364136223846793005ULL
#define RND64_C 1ULL
#define RNDF_MUL 5.4210108624275222e-20

static unsigned long long rnd64_jump(unsigned long long n, unsigned long long seed) {
    unsigned long long a_k = RND64_A, c_k = RND64_C, ran = seed;
    for (; n; n >>= 1) {
        if (n & 1) ran = a_k * ran + c_k;
        c_k *= (a_k + 1);
        a_k *= a_k;
    }
    return ran;
}

/* Fill the mb x nb tile whose top-left element is (row0,col0) of the symmetric N x N matrix
 * (bigM = N for a whole matrix).  Entries beyond the edge: identity. */
void oracle_plgsy(double bump, int mb, int nb, double* A, int lda, long long bigM, long long row0, long long col0,
                  long long N, unsigned long long seed) {
    for (int c = 0; c < nb; ++c)
        for (int r = 0; r < mb; ++r) {
            const long long gi = row0 + r, gj = col0 + c;
            double v;
            if (gi >= N || gj >= N) {
                v = (gi == gj) ? 1.0 : 0.0;
            } else {
                const long long hi = gi >= gj ? gi : gj, lo = gi >= gj ? gj : gi;
                const unsigned long long ran =
                    rnd64_jump((unsigned long long)hi + (unsigned long long)lo * (unsigned long long)bigM, seed);
                const double prod = (double)ran * RNDF_MUL;
                v = 0.5 - prod;
                if (gi == gj) v += bump;
            }
            A_(A, lda, r, c) = v;
        }
}

/* ---- the matrix of the reference's plain CPU program (LP:35-45): rand()/RAND_MAX, row-major
 * N x N, diagonal += N, lower mirrored from upper.  Written here in the memory order LP uses. */
void oracle_lp_matrix(int N, double* A) {
    srand(0);
    for (size_t i = 0; i < (size_t)N * N; ++i) A[i] = ((double)rand()) / RAND_MAX;
    for (int i = 0; i < N; ++i) {
        A[(size_t)i * N + i] += N;
        for (int j = 0; j < i; ++j) A[(size_t)i * N + j] = A[(size_t)j * N + i];
    }
}

/* ---- checks ------------------------------------------------------------------------------ */
/* || A - L L^T ||_F and ||A||_F for symmetric A given by its lower triangle, L lower (strict
 * upper of the L array ignored).  out[0] = residual norm, out[1] = ||A||_F. */
void oracle_backward_error(int n, const double* A, int lda, const double* L, int ldl, double* out) {
    double r2 = 0.0, a2 = 0.0;
    double* row_i = (double*)malloc(sizeof(double) * (size_t)n);
    for (int j = 0; j < n; ++j) {
        /* column j of L L^T below the diagonal: sum_{k<=j} L[i][k] L[j][k] */
        for (int i = j; i < n; ++i) row_i[i] = 0.0;
        for (int k = 0; k <= j; ++k) {
            const double ljk = A_(L, ldl, j, k);
            const double* ck = &A_(L, ldl, 0, k);
            for (int i = j; i < n; ++i) row_i[i] += ck[i] * ljk;
        }
        for (int i = j; i < n; ++i) {
            const double a = A_(A, lda, i, j), d = a - row_i[i];
            const double w = (i == j) ? 1.0 : 2.0;
            r2 += w * d * d;
            a2 += w * a * a;
        }
    }
    free(row_i);
    out[0] = sqrt(r2);
    out[1] = sqrt(a2);
}

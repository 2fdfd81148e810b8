/*
 * chol_b200.h — C ABI of libchol_b200.so: B200 (sm_100a) tiled FP64 Cholesky.
 *
 * Drop-in boundary for the ONE hot path of HugoVuach/Dense-linear-app: the tile-task
 * Cholesky (POTRF / TRSM / SYRK / GEMM on b x b column-major FP64 tiles).  Every entry
 * point names the reference interface it replaces (paths under /root/reference):
 *   W2  = cholesky_armonik/w_c_cons_v2/worker_construction2/src/worker_distrib.cpp
 *   C1  = cholesky_armonik/w_c_cons_v1/client_construction/client/src/client_distrib.cpp
 *   V6  = Cholesky_chameleon_VM/cho/docker_installation_and_bench_files/v6_test.c
 *
 * Conventions
 *   - all matrix pointers are DEVICE pointers unless the name ends in _host;
 *   - tiles are column-major, FP64, operated on in place (W2:76-79, W2:212-227);
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); calls are
 *     asynchronous on that stream, nothing here synchronises the device;
 *   - return value: 0 ok, <0 = -(index of the bad argument) (LAPACK style),
 *     >0 = CUDA runtime error code (see chol_last_error()).  Numerical failure
 *     (matrix not positive definite) is reported through the device int `d_info`
 *     exactly like LAPACK dpotrf's info (1-based index of the failing minor, first
 *     failure wins, 0 = success) because the result is only known when the stream
 *     has run;
 *   - no function allocates user-visible memory; scratch comes from the caller and is
 *     sized by the *_workspace() queries;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns
 *     an error.
 */
#ifndef CHOL_B200_H
#define CHOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------- */

/* CHAMELEON_Init(ncpu, ngpu) analogue (V6:41, W2:589): selects the device and sets the
 * kernels' shared-memory attributes.  Idempotent, thread-safe. */
int chol_init(int device);
/* CHAMELEON_Finalize analogue (V6:93). */
int chol_finalize(void);
/* Text of the last error raised on this thread ("" if none). */
const char* chol_last_error(void);
/* Library version string, e.g. "chol_b200 0.1 sm_100a". */
const char* chol_version(void);

/* Number of CUDA kernels this library has enqueued so far in the process (all threads); the
 * benchmark reports the difference over its timed region. */
unsigned long long chol_launch_count(void);

/* ---- grouped rank-K update: the hot kernel (K4/K5 of SURVEY 2b) ------------------ */

/* One C-tile update  C <- beta*C + alpha * A * B^T  (A: m x k, B: n x k, col-major).
 * flags bit0 (CHOL_TASK_LOWER): only the lower triangle of C is referenced and written
 * (dsyrk uplo=Lower semantics, W2:416).  32 bytes, laid out for direct upload from a
 * numpy/torch int64 [ntasks,4] array. */
typedef struct chol_task {
    double*       C;
    const double* A;
    const double* B;
    int64_t       flags;
} chol_task_t;
#define CHOL_TASK_LOWER 1

/* All tasks share m, n, k and the leading dimensions.  `d_tasks` is a DEVICE array.
 * Replaces the per-tile SYRK/GEMM submissions of one wave of the client loop
 * (C1:307-329, C2:541-561) by ONE persistent launch: the "fused SYRK+GEMM trailing
 * update per panel".  Fast path (DMMA + TMA bulk staging) needs m,n even, k%4==0 and
 * even leading dimensions; other shapes take the generic CUDA kernel.  Tile pointers should be
 * 16-byte aligned: a task of the list whose pointers are only 8-byte aligned is still computed
 * correctly, by a slow scalar path inside the kernel (the pointers live in device memory, so
 * the host cannot pre-check them). */
int chol_gemm_tasks(const chol_task_t* d_tasks, int ntasks, int m, int n, int k,
                    int lda, int ldb, int ldc, double alpha, double beta, void* stream);

/* Same, with the residency of the update grid chosen by the caller: ctas_per_sm = 2 is chol_gemm_tasks;
 * ctas_per_sm = 1 leaves half of every SM free, so that the (short, latency-critical) panel kernels
 * of the next step find a free slot at once instead of waiting for update CTAs to retire.  The
 * whole-matrix driver uses it for the last steps, where the panel chain bounds the step time. */
int chol_gemm_tasks_ex(const chol_task_t* d_tasks, int ntasks, int m, int n, int k,
                       int lda, int ldb, int ldc, double alpha, double beta, int ctas_per_sm,
                       void* stream);

/* ---- the four tile ops of the worker (W2:179-546) -------------------------------- */

/* POTRF: A <- chol_lower(A); strict upper triangle untouched.
 * Replaces CHAMELEON_dpotrf_Tile(ChamLower, dA) on a 1-tile descriptor (W2:238).
 * `work` >= chol_potrf_tile_workspace(b) bytes; on return it holds the inverses of the
 * 128x128 diagonal blocks of L (input for chol_trsm_tiles).  `info_base` is added to a
 * failing index so a whole-matrix driver gets the global LAPACK info (V6:56). */
size_t chol_potrf_tile_workspace(int b);
int chol_potrf_tile(int b, double* A, int lda, double* work, int* d_info, int info_base,
                    void* stream);

/* TRSM: A <- A * L^{-T}   (Right, Lower, Trans, NonUnit, alpha=1).
 * Replaces CHAMELEON_dtrsm_Tile(ChamRight,ChamLower,ChamTrans,ChamNonUnit,1.0,dL,dA)
 * (W2:323).  Stateless form: inverts the diagonal blocks of L into `work`
 * (>= chol_trsm_tile_workspace(b) bytes) first. */
size_t chol_trsm_tile_workspace(int b);
int chol_trsm_tile(int b, const double* L, int ldl, double* A, int lda, double* work,
                   void* stream);

/* Panel form: the same solve applied to `ntiles` tiles (device array of tile pointers)
 * in one launch sequence, re-using the diagonal-block inverses left in `potrf_work` by
 * chol_potrf_tile.  `d_task_scratch` is reserved (the kernels address the tiles straight from
 * the pointer list; may be NULL).  Replaces the TRSM loop of one wave (C1:295-303). */
int chol_trsm_tiles(int b, const double* L, int ldl, const double* potrf_work,
                    double* const* d_tiles, int ntiles, int lda, void* d_task_scratch,
                    void* stream);

/* Panel form fused with the panel transport (multi-GPU, tile sizes that are multiples of 32): as
 * chol_trsm_tiles, and every finished 128-column block of tile t is ALSO stored, by the kernel that
 * computes it, into up to 7 peers' receive slots over NVLink — d_peer_dst[t*npeer + q] is the address
 * of tile t in peer q's mapped slot buffer (chol_peer_open), 0 if that peer does not read the tile.
 * The transfer overlaps the solve instead of following it as a copy; raise the peers' flags with
 * chol_flag_post on the same stream afterwards. */
int chol_trsm_tiles_push(int b, const double* L, int ldl, const double* potrf_work,
                         double* const* d_tiles, int ntiles, int lda,
                         const long long* d_peer_dst, int npeer, void* stream);

/* SYRK: C <- C - A*A^T, lower triangle only.
 * Replaces CHAMELEON_dsyrk_Tile(ChamLower,ChamNoTrans,-1.0,dA,1.0,dC) (W2:416). */
int chol_syrk_tile(int b, const double* A, int lda, double* C, int ldc, void* stream);

/* GEMM: C <- C - Ai*Aj^T.
 * Replaces CHAMELEON_dgemm_Tile(ChamNoTrans,ChamTrans,-1.0,dAi,dAj,1.0,dC) (W2:511). */
int chol_gemm_tile(int b, const double* Ai, int ldai, const double* Aj, int ldaj,
                   double* C, int ldc, void* stream);

/* ---- batched small Cholesky (the ArmoniK many-task workload, C1:139-141) ---------- */

/* `batch` independent n x n lower Cholesky factorizations, matrix i at A + i*stride
 * (doubles).  d_info[i] = LAPACK info of matrix i. */
int chol_potrf_batched(int n, int batch, double* A, int lda, long long stride,
                       int* d_info, void* stream);

/* ---- panel transport between the GPUs of one box ----------------------- */

/* The reference leaves the movement of tiles between workers to StarPU/MPI (descriptor grid p,q:
 * V6:44-45) or to the ArmoniK object store (W2:186,261).  Here the factored panel is PUSHED into
 * receive slots of the ranks that read it, over NVLink with the copy engines, and announced by a
 * flag word; no kernel is resident on a waiting GPU.  One process per GPU.
 *
 * chol_peer_alloc   cudaMalloc'ed, zero-filled buffer that can be exported to other processes.
 * chol_peer_export  64-byte CUDA IPC handle of such a buffer (send it to the peers by any means).
 * chol_peer_open    map a peer's buffer from its handle (enables P2P access); chol_peer_close unmaps.
 */
#define CHOL_IPC_HANDLE_BYTES 64
int chol_peer_alloc(size_t bytes, void** out);
int chol_peer_free(void* p);
int chol_peer_export(void* p, void* handle64);
int chol_peer_open(const void* handle64, void** out);
int chol_peer_close(void* p);

/* One push: `count` tiles of `tile_bytes` bytes from local `src` (tiles `src_stride` tiles apart) to
 * the peer-mapped `dst` (`dst_stride` apart) on `stream`, after (optionally) waiting until the LOCAL
 * word *credit has reached credit_value (the reader has released the slot), followed (optionally)
 * by raising the PEER-mapped word *flag to flag_value.  Flag comparisons are cyclic:
 * (int32)(*word - value) >= 0. */
typedef struct chol_xfer {
    void*           dst;
    const void*     src;
    long long       tile_bytes;
    int             count;
    int             dst_stride;
    int             src_stride;
    uint32_t        credit_value;
    const uint32_t* credit;
    uint32_t*       flag;
    uint32_t        flag_value;
    uint32_t        reserved;
    void*           stream;
} chol_xfer_t;

/* Enqueue `n` pushes; each first waits for everything enqueued so far on `ready_stream` (the
 * stream that produced the data).  Copies use the copy engines (no SM work besides the one-warp
 * flag store). */
int chol_peer_send(const chol_xfer_t* xf, int n, void* ready_stream);
/* Make `stream` wait until (int32)(*flag - value) >= 0.  `flag` is LOCAL device memory written by a
 * peer.  A stream memory operation (cuStreamWaitValue32): nothing runs on the SMs while waiting. */
int chol_flag_wait(const uint32_t* flag, uint32_t value, void* stream);
/* Raise up to 16 (peer-mapped or local) flag words to `value`, ordered after the work already
 * enqueued on `stream`.  `flags` is a HOST array of device pointers. */
int chol_flag_post(uint32_t* const* flags, int n, uint32_t value, void* stream);

/* ---- generators and checks (V6:46, V6:51, V6:72-86; off the timed path) ----------- */

/* dplgsy-like tile generator (V6:46 CHAMELEON_dplgsy_Tile(bump, ChamLower, desc, seed)):
 * fills the mb x nb tile whose top-left element is (row0, col0) of the symmetric N x N
 * matrix; element (i,j), i>=j, is 0.5 - LCG^(i + j*bigM)(seed) * 2^-64; diagonal += bump.
 * Elements beyond the matrix edge (row/col >= N) are set to the identity so ragged
 * edge tiles stay positive definite. */
int chol_plgsy_tile(double bump, int mb, int nb, double* A, int lda, long long bigM,
                    long long row0, long long col0, long long N, unsigned long long seed,
                    void* stream);

/* per-column sums of squares of a tile -> d_out[n] (mode 0: all m rows; mode 1: lower triangle
 * counted as a symmetric matrix, i.e. strict lower twice + diagonal once, rows above the diagonal
 * ignored).  Building block of the Frobenius norm of the residual (V6:72-86). */
int chol_tile_sumsq(int m, int n, const double* A, int lda, int mode, double* d_out,
                    void* stream);
/* per-row and per-column sums of |a_ij| of a tile (mode as above: mode 1 reads only the
 * lower triangle).  d_rows[m], d_cols[n].  Building block of dlange(inf) (V6:72,84). */
int chol_tile_abs_sums(int m, int n, const double* A, int lda, int mode, double* d_rows,
                       double* d_cols, void* stream);
/* B <- lower triangle of A, strict upper zeroed (dlacpy(ChamLower), V6:77). */
int chol_tile_tril(int n, const double* A, int lda, double* B, int ldb, void* stream);
/* In-place transposition of ntiles n x n column-major tiles, tile t at A + t*stride doubles.  The bridge for
 * uplo = Upper (cham_uplo_t of CHAMELEON_dpotrf_Tile, V6:56; `--uplo U` of v3_script_cholesky_x_arg_gpt.c:36-44):
 * A = U^T U is the lower factorization of the transposed tiles, U(j,i) = L(i,j)^T. */
int chol_tile_transpose(int n, double* A, int lda, long long stride, int ntiles, void* stream);

/* ---- SM partition for the latency-critical panel chain -------------------------------- */

/* The reference hands POTRF to a CPU worker and the updates to the CUDA worker (StarPU: no GPU codelet
 * for POTRF, W2:238 / SURVEY 8a row a2), so its panel chain never queues behind GEMM tiles.  Here both
 * run on one B200: while a trailing update fills every SM, each of the ~24 short kernels of a POTRF tile
 * waits for an update CTA to retire (measured: 0.55 ms alone, 1.3-1.5 ms under the update).  This call
 * splits the SMs with CUDA green contexts: `panel_stream` runs its kernels only on a private group of
 * >= min_panel_sms SMs (rounded up to the hardware granularity, 8 on sm_100), `rest_stream` only on the
 * remaining ones; both interoperate with ordinary streams and events.  The whole-matrix driver uses the
 * pair for the last steps of a factorization, where the panel chain, not the update, bounds the step.
 * Returns non-zero (and leaves *out zeroed) when the driver cannot provide such a split. */
typedef struct chol_partition {
    void* handle;        /* for chol_partition_destroy */
    void* panel_stream;  /* cudaStream_t */
    void* rest_stream;   /* cudaStream_t */
    int panel_sms, rest_sms;
} chol_partition_t;
int chol_partition_create(int device, int min_panel_sms, chol_partition_t* out);
int chol_partition_destroy(void* handle);

/* ---- microbenchmarks for the roofline denominators -------------------------------- */

/* kind 0: DFMA chains, 1: DMMA m8n8k4 chains on 8 warps per SM, 2: same on 4 warps.  Runs `iters` inner iterations on every
 * SM, returns achieved FLOP/s in *flops_out (timed with CUDA events on `stream`). */
int chol_fp64_peak(int kind, int iters, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHOL_B200_H */

/*
 * paladin_gpu.h -- C-ABI of libpaladin_gpu.so, the B200 (sm_100a) replacement for the one numerical
 * call on paladin's hot path.
 *
 * What it replaces in the reference (/root/reference):
 *   - extern "C" dgeev_(...)                         src/lapack-wrapper.hpp:47-60
 *     called twice per matrix (lwork=-1 query, solve) src/paladin.hpp:248-249, :253-254
 *   - the dense scatter inside SquareMatrix::read_matrix (mat_[(i-1)+n*(j-1)] = v, file order,
 *     later entries overwrite earlier ones)           src/square-matrix.hpp:101-117 (store at :115)
 *     over the zero fill of the SquareMatrix ctor      src/square-matrix.hpp:70-71
 *
 * Plain C types only. All matrices are column-major, indices in COO input are 1-based exactly as
 * they appear in the MatrixMarket file. Eigenvalues/eigenvectors come back in LAPACK dgeev's layout
 * (conjugate pairs consecutive, positive imaginary part first, wi[k+1] == -wi[k] exactly, vectors
 * real-packed) because the reference's unpack loop (src/paladin.hpp:256-333) consumes that layout.
 *
 * There is NO CPU fallback: every compute entry fails with PALADIN_GPU_ENODEVICE when no sm_100
 * device is usable.
 *
 * Thread model: the library is thread-safe. Every batch owns a CUDA stream, so batches driven from different
 * host threads overlap on one device (copies of one batch run under the kernels of another); calls on ONE batch
 * must come from one thread at a time.
 */
#ifndef PALADIN_GPU_H_
#define PALADIN_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (return value of every int-returning entry; 0 = success) -------------------- */
#define PALADIN_GPU_OK 0
#define PALADIN_GPU_EINVAL 1     /* bad argument (null pointer, negative size, bad job character) */
#define PALADIN_GPU_ENODEVICE 2  /* no usable CUDA device / device index out of range / not sm_100 */
#define PALADIN_GPU_ECUDA 3      /* a CUDA runtime call or kernel failed; see paladin_gpu_last_error */
#define PALADIN_GPU_ENOMEM 4     /* device or pinned-host allocation failed */
#define PALADIN_GPU_ENOTSUP 5    /* valid request that this build does not implement */

/* per-matrix info values (LAPACK dgeev convention, extended):
 *    0      success
 *    i > 0  the QR iteration failed to converge; wr/wi[i..n) hold converged eigenvalues, no vectors
 *   -4      the matrix holds a non-finite entry (argument 4 of dgeev); nothing was computed
 *   -101    a COO index is outside [1, n] (the reference would write out of bounds; we refuse) */
#define PALADIN_GPU_INFO_NONFINITE (-4)
#define PALADIN_GPU_INFO_BAD_INDEX (-101)

/* stage indices for paladin_gpu_batch_stage_ms */
#define PALADIN_GPU_STAGE_SCATTER 0    /* COO -> dense */
#define PALADIN_GPU_STAGE_BALANCE 1    /* scale check + balancing (0 when fused into a later stage) */
#define PALADIN_GPU_STAGE_HESSENBERG 2 /* reduction to Hessenberg form (+ explicit Q when vectors) */
#define PALADIN_GPU_STAGE_QR 3         /* QR iteration to (quasi-)triangular form */
#define PALADIN_GPU_STAGE_VECTORS 4    /* eigenvectors of T, back-transform, un-balance, normalise */
#define PALADIN_GPU_STAGE_SMALL 5      /* fused warp-per-matrix path for n <= 64 (all stages in one kernel) */
#define PALADIN_GPU_STAGE_GENERIC 6    /* fused CTA-per-matrix path in global memory (all stages; debug fallback) */
#define PALADIN_GPU_STAGE_MID 7        /* unused (the CTA-per-matrix path for n > 64 reports stages 2, 3 and 4) */
#define PALADIN_GPU_NUM_STAGES 8

/* ---- lifecycle ------------------------------------------------------------------------------- */

/* Number of usable devices (compute capability 10.x). 0 when there is none. Never fails. */
int paladin_gpu_device_count( void );

/* Creates the per-device context (stream, events). Idempotent. */
int paladin_gpu_init( int device );

/* Releases everything the library holds on `device` (open batches must be destroyed first). */
int paladin_gpu_shutdown( int device );

/* Human-readable description of the last failure on the calling thread ("" if none). */
const char* paladin_gpu_last_error( void );

/* Library version string, e.g. "paladin_gpu 0.1 sm_100a". */
const char* paladin_gpu_version( void );

/* Page-locked host memory so that uploads/downloads run as asynchronous DMA. */
int paladin_gpu_host_alloc( void** ptr, size_t bytes );
int paladin_gpu_host_free( void* ptr );

/* Device-side stopwatch on the library's own stream of `device` (CUDA events): start records an
 * event, stop records a second one, waits for it and returns the elapsed milliseconds. Used by
 * bench.py so that timings are taken on the stream the kernels are launched on. */
int paladin_gpu_timer_start( int device );
int paladin_gpu_timer_stop( int device, float* ms );

/* ---- one-shot batched entry (host buffers in, host buffers out) -------------------------------
 * Replaces, for `count` matrices at once, the reference's scatter (src/square-matrix.hpp:115) and
 * its two dgeev_ calls (src/paladin.hpp:248-254).
 *   n[k]                 order of matrix k (0 allowed)
 *   nnz_offsets[k..k+1]  range of matrix k's entries in coo_i/coo_j/coo_v (count+1 values)
 *   coo_i, coo_j         1-based row / column, file order; duplicates resolve to the LAST entry
 *   jobvl, jobvr         'N' or 'V'
 *   wr[k], wi[k]         n[k] doubles each (caller-owned)
 *   vl[k], vr[k]         n[k]*n[k] doubles each, ld = n[k]; may be NULL when the job is 'N'
 *   info[k]              per-matrix status, see above
 */
int paladin_gpu_geev_batch( int device, int count, const int* n, const int64_t* nnz_offsets, const int* coo_i,
                            const int* coo_j, const double* coo_v, char jobvl, char jobvr, double* const* wr,
                            double* const* wi, double* const* vl, double* const* vr, int* info );

/* ---- staged entry: the same work with the phases separated, so that a caller (or a benchmark)
 * can keep inputs resident in device memory, overlap phases of different batches, and read the
 * dense matrices for parity checks. ------------------------------------------------------------ */
typedef struct paladin_gpu_batch paladin_gpu_batch;

/* Allocates device storage for `count` matrices of the given orders and entry counts. */
int paladin_gpu_batch_create( int device, int count, const int* n, const int64_t* nnz_offsets, char jobvl, char jobvr,
                              paladin_gpu_batch** out );

/* Host -> device copy of the COO triplets (asynchronous when the buffers are pinned; the call
 * returns after the copies completed). */
int paladin_gpu_batch_upload( paladin_gpu_batch* b, const int* coo_i, const int* coo_j, const double* coo_v );

/* COO -> dense on the device. */
int paladin_gpu_batch_scatter( paladin_gpu_batch* b );

/* Alternative to upload+scatter: dense column-major matrices, concatenated (sum n[k]^2 doubles). */
int paladin_gpu_batch_upload_dense( paladin_gpu_batch* b, const double* dense );

/* Parity/debug: copies the dense matrix k (n[k]^2 doubles, column-major) as the scatter left it.
 * Only meaningful between scatter and solve (solve destroys the matrices). */
int paladin_gpu_batch_get_dense( paladin_gpu_batch* b, int k, double* out );

/* Eigen-solve of every matrix of the batch on the device; returns when it finished. */
int paladin_gpu_batch_solve( paladin_gpu_batch* b );

/* Device -> host copy of the results into caller-owned buffers laid out as concatenations:
 * wr/wi: sum n[k] doubles; vl/vr: sum n[k]^2 doubles (NULL when not computed); info: count ints. */
int paladin_gpu_batch_download( paladin_gpu_batch* b, double* wr, double* wi, double* vl, double* vr, int* info );

/* The write filter of the reference's eigenvector writers (reference src/spectrum-util.hpp:134-196) evaluated on the device,
 * after paladin_gpu_batch_solve: for every eigenvalue index c of matrix k the complex vector the reference unpacks for it
 * (reference src/paladin.hpp:256-333: (re, 0), (re, +im) for the first eigenvalue of a conjugate pair, (re, -im) for the
 * second) is scanned once:
 *   max2[sum n[0..k) + c] = max_r |x_r|^2         (std::norm of the largest element, src/spectrum-util.hpp:175-176)
 *   kept[k]              = number of entries of all n[k] vectors with |x_r|^2 > threshold * max2   (the nnz of the header line)
 * Products and sums are rounded separately (no fused multiply-add), so both agree bit for bit with the host evaluation.
 * side: 'R' or 'L' (the batch must have been created with the matching job). max2: sum n[k] doubles, kept: count values;
 * matrices with info != 0 get kept = 0. */
int paladin_gpu_batch_vector_filter( paladin_gpu_batch* b, char side, double threshold, double* max2, long long* kept );

/* ---- the text ends of the path on the device (GPU-assisted MatrixMarket parse, GPU-side formatting) --------- */

/* Alternative to paladin_gpu_batch_upload: the ENTRY LINES of every matrix as raw MatrixMarket text ("i j v\n", exactly as on
 * disk, after the size line; reference reader src/square-matrix.hpp:101-117). text_off[count + 1] are byte offsets into
 * `text` (host memory, ideally page-locked): matrix k's entry lines are text[text_off[k] .. text_off[k+1]). The device
 * splits the lines and converts the fields with the values atoi / atoi / atof give (bit-exact); fields it cannot convert
 * with certainty (hex floats, inf / nan, more than 19 significant digits, ...) are converted inside this call by the C
 * library on the host. needs_host[k] (count ints, may be NULL) is set to 1 when matrix k has to go through the caller's own
 * reader instead -- a line with fewer than two blanks, a field longer than the reference's buffers (5 / 5 / 31 characters),
 * fewer lines than entries -- and its triplets are then expected through paladin_gpu_batch_upload_coo_one. */
int paladin_gpu_batch_upload_text( paladin_gpu_batch* b, const char* text, const int64_t* text_off, int* needs_host );

/* Streaming variant of the text upload: paladin_gpu_batch_text_begin reserves `total_bytes` of device text,
 * paladin_gpu_batch_text_put copies a piece of it (any thread, any order, e.g. one file at a time out of a small page-locked
 * staging buffer), and paladin_gpu_batch_upload_text called with text == NULL converts what has been placed (text_off[0]
 * must then be 0). No batch-sized host buffer is needed. */
int paladin_gpu_batch_text_begin( paladin_gpu_batch* b, size_t total_bytes );
int paladin_gpu_batch_text_put( paladin_gpu_batch* b, size_t offset, const char* src, size_t bytes );

/* Host -> device copy of the COO triplets of ONE matrix of the batch (nnz of that matrix values each). */
int paladin_gpu_batch_upload_coo_one( paladin_gpu_batch* b, int k, const int* coo_i, const int* coo_j, const double* coo_v );

/* Parity/debug: the COO triplets as they are on the device (total nnz values each; any pointer may be NULL). */
int paladin_gpu_batch_get_coo( paladin_gpu_batch* b, int* coo_i, int* coo_j, double* coo_v );

/* The BODY of the reference's eigenvector files (reference src/spectrum-util.hpp:134-196, everything after the "N N nnz"
 * header line) formatted on the device, after paladin_gpu_batch_solve, for the matrices [first, first + count) of the batch:
 * for every vector (index outermost) and row the line "row col re im\n" when |x|^2 > threshold * max |x|^2, numbers as
 * operator<<(double) prints them (6 significant digits, "%g"). side: 'R' or 'L'. text: caller buffer of `cap` bytes (48
 * bytes per vector component of the range always suffice; page-locked memory makes the copy a DMA); the body of matrix
 * first + q is text[body_off[q] .. body_off[q+1]) (count + 1 offsets). kept[q] = number of lines = the nnz of the header
 * line. needs_host[q] = 1 when a number of that matrix could not be rounded with certainty on the device (its body is then
 * undefined and the caller formats that matrix from the downloaded vectors); matrices with info != 0 get empty bodies. */
int paladin_gpu_batch_format_vectors( paladin_gpu_batch* b, char side, double threshold, int first, int count, char* text, size_t cap,
                                      int64_t* body_off, long long* kept, int* needs_host );

/* Streaming variant: paladin_gpu_batch_format_vectors called with text == NULL (range of at most 48 Mi vector components)
 * leaves the formatted bodies on the device; this entry copies bytes [offset, offset + bytes) of them (the offsets are the
 * body_off values of that call) to dst. A small page-locked staging buffer per writer thread is enough, whatever the batch
 * holds. May be called from several threads at once for the same batch; the text stays valid until the next format call. */
int paladin_gpu_batch_fetch_text( paladin_gpu_batch* b, size_t offset, size_t bytes, char* dst );

/* Device time of each stage of the most recent scatter/solve, in milliseconds, measured with CUDA
 * events on the library's own stream. ms must hold PALADIN_GPU_NUM_STAGES floats. launches (may be
 * NULL) receives the number of kernel launches of each stage. */
int paladin_gpu_batch_stage_ms( paladin_gpu_batch* b, float* ms, int* launches );

/* Stopwatch on the batch's own stream (CUDA events): the device time between start and stop of everything
 * enqueued for this batch in between. */
int paladin_gpu_batch_timer_start( paladin_gpu_batch* b );
int paladin_gpu_batch_timer_stop( paladin_gpu_batch* b, float* ms );

/* Phase profile of the most recent CTA-per-matrix solve: SM cycles summed over all CTAs, 16 slots
 * (0 balance, 1 Hessenberg panels, 2 form Q, 3 QR remainder (fused kernel; in the multishift kernel: the share of slot 14 that follows the window's own QR iteration), 4 eigenvectors of T, 5 un-balance + normalise, 6 back-transform
 *  GEMM, 7 Hessenberg trailing updates, 8 deflation scans, 9 shift computation, 10 in-window bulge chasing, 11 strip updates by
 *  recorded reflectors, 12 small-block solves, 13 strip updates by small U, 14 early-deflation window, 15 its off-window update).
 * Diagnostic only. */
int paladin_gpu_batch_phase_cycles( paladin_gpu_batch* b, long long* cycles );

int paladin_gpu_batch_destroy( paladin_gpu_batch* b );

/* ---- single-matrix shim with the exact Fortran signature the reference declares ---------------
 * (src/lapack-wrapper.hpp:47-60). Exported as paladin_gpu_dgeev_ so that it can coexist with a real
 * LAPACK in one process; link with -Ddgeev_=paladin_gpu_dgeev_ (or the alias library) to make the
 * unmodified reference call it. lwork == -1 answers the workspace query with work[0] = 1 and
 * touches nothing else. Uses device 0 unless PALADIN_GPU_DEVICE is set in the environment. */
void paladin_gpu_dgeev_( const char* jobvl, const char* jobvr, int* n, double* a, int* lda, double* wr, double* wi,
                         double* vl, int* ldvl, double* vr, int* ldvr, double* work, int* lwork, int* info );

#ifdef __cplusplus
}
#endif

#endif /* PALADIN_GPU_H_ */

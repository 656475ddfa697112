/*
 * include/cgb200.h -- extended C ABI of liboclcg.so (the B200 CG engine).
 *
 * The reference's only entry point is cg() (include/clcg.h).  It re-creates its
 * OpenCL context, re-compiles its kernels and re-uploads the matrix on every call
 * (clcg.c:142-214), has no double precision (main.c:49), no convergence test and
 * no error reporting (clcg.c:52-56).  The functions below expose the same path
 * piecewise so that a caller can keep the matrix resident, solve in double
 * precision, stop on a tolerance and read timings.  Each one names the part of the
 * reference it stands for.
 *
 * Conventions
 *   - plain C types only; every pointer argument may be a HOST or a DEVICE pointer
 *     (the CUDA runtime tells them apart), so the same call serves a ctypes user with
 *     numpy arrays and a caller whose data already lives in HBM;
 *   - vectors use the reference's layout: nRHS blocks of n values, RHS r at offset
 *     r*n (axpy.cl:12, p_h-PY_C-CL.py:1929-1930), unless a `layout` argument says
 *     CGB200_LAYOUT_ROWMAJOR ([n][k], the engine's internal SpMM layout);
 *   - return value: 0 ok, < 0 failure (cgb200_last_error() has the text),
 *     > 0 numerical flags of cgb200_solve();
 *   - no function calls exit() or prints.
 */
#ifndef CGB200_H
#define CGB200_H

#include <stddef.h>

#ifndef CGB200_API
#define CGB200_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* value types: what `isComplex` (clcg.c:154-158) selects, times two precisions */
#define CGB200_F32 0   /* float                 -- cg(..., isComplex=0) */
#define CGB200_F64 1   /* double                -- cgd(..., isComplex=0) */
#define CGB200_C64 2   /* float  (re, im) pairs -- cg(..., isComplex=1), cfloat of cmplx.h:4 */
#define CGB200_C128 3  /* double (re, im) pairs -- cgd(..., isComplex=1) */

#define CGB200_LAYOUT_CLCG 0      /* [k][n]: RHS r at r*n, the cg() ABI */
#define CGB200_LAYOUT_ROWMAJOR 1  /* [n][k]: the k values of a row are contiguous */

#define CGB200_OK 0
#define CGB200_ERR_ARG (-1)
#define CGB200_ERR_CUDA (-2)
#define CGB200_ERR_NOMEM (-3)
#define CGB200_ERR_NCCL (-4)
#define CGB200_ERR_UNSUPPORTED (-5)
/* flags OR-ed into a positive return of cgb200_solve() */
#define CGB200_FLAG_MAXIT 1      /* tol > 0 and some RHS did not reach it */
#define CGB200_FLAG_BREAKDOWN 2  /* d.q or delta became 0 / non-finite for some RHS; that RHS was frozen */

typedef struct cgb200_ctx *cgb200_handle;

/* The prologue of cg() (clcg.c:137-214): device selection, buffers, matrix upload --
 * done once.  Copies the CSR arrays (host or device pointers) into HBM on `device`,
 * inspects the row-length distribution and picks the SpMV schedule. */
CGB200_API int cgb200_create(cgb200_handle *out, int n, long long nnz, const void *aValues,
                  const int *aPointers, const int *aCols, int dtype, int device);

/* New matrix content of the SAME sizes (n, nnz, dtype) into the resident buffers: the
 * uploads of clcg.c:202-207 without the allocations.  The SpMV schedule is rebuilt only
 * when the row offsets changed. */
CGB200_API int cgb200_update(cgb200_handle h, const void *aValues, const int *aPointers, const int *aCols);

/* clcg.c:432-459 (release everything). */
CGB200_API int cgb200_destroy(cgb200_handle h);

/* All work of `h` is enqueued on one CUDA stream (clcg.c:184 creates one in-order
 * queue).  By default the handle owns a stream; pass a cudaStream_t to share one. */
CGB200_API int cgb200_set_stream(cgb200_handle h, void *cuda_stream);

/* Tuning knobs (the reference's are compile-time macros, clcg.c:37-43).
 *   "spmv_variant"   [k = 1, CSR kernels]
 *                    0 auto: 6, or 3 for matrices whose row lengths vary wildly ("auto_irregular")
 *                    1 CSR-vector: lanes_per_row lanes per row (the simplest kernel; also the fallback without a schedule)
 *                    3 CSR-stream fed by TMA bulk copies through a 2-stage mbarrier ring: every gather of a tile in
 *                      flight at once, balanced row sums (the schedule for power-law matrices)
 *                    6 row-direct CSR-stream fed by TMA (no product buffer)
 *                    (2, 4, 5, 7, 8, 9 of round 1 -- plain-load stream, deeper rings -- measured slower everywhere and were removed)
 *   "solver"         0 auto: one cooperative launch for the whole solve when an iteration's working set
 *                      fits the L2 (2 grid barriers per iteration), else three kernels per iteration
 *                    1 three kernels per iteration in CUDA graphs     2 single cooperative launch
 *   "lanes_per_row"  0 auto | 1,2,4,8,16,32   lanes cooperating on one row (variant 1)
 *   "graph_chunk"    CG iterations captured per CUDA graph launch (default 16)
 *   "use_graph"      0/1
 *   "blocks_per_sm"  0 auto | n               persistent-grid size multiplier
 *   "defer_len"      rows with more than this many non-zeros per lane are walked by a whole warp (default 16,
 *                    0 off): keeps a tile of a power-law matrix from waiting for its longest row
 *   "pdl"            bit mask, programmatic dependent launch of 1 spmv | 2 x/r update | 4 direction update
 *                    (default 1): the SpMV's prologue -- barriers, first matrix tiles by TMA -- overlaps the tail
 *                    of the direction update
 *   "pdl_early"      1 (default): a kernel lets its dependents become resident as soon as it has started
 *   "vec_carveout"   -1 (default: leave alone) | 0..100: preferred shared-memory carve-out of the vector kernels, the
 *                    knob that showed why "pdl" = 2 hurts
 *   "auto_irregular" 1 (default): spmv_variant 0 picks variant 3 for matrices whose row lengths vary wildly
 *   "l2_keep"        d, q and r tagged evict-last in L2 (matrix stream and x are evict-first): 0 off, 1 on,
 *                    -1 by size (on when the three vectors fit half the L2)
 *   "pattern"        1 (default): k = 1 runs from the row-pattern dictionary when the matrix has <= 4096 distinct rows
 *                    ("patterns", read-only, tells how many were found; 0 = CSR kernels in use)
 *   "cg2"            1 (default): k = 1 on a matrix with a row-pattern dictionary whose column offsets fit a window
 *                    plan runs the TWO-kernel iteration (csrc/cg2.cuh: direction update folded into the SpMV's
 *                    gather, x lagging one update, 9 vector passes instead of 11); "cg2_ok" (read-only) tells
 *                    whether the matrix qualifies; "cg2_stages" sets the depth of its TMA ring (default: what fits);
 *                    "march" 1 (default): grid operators with one far offset +-P whose vectors are too big for the L2
 *                    run the plane-marching variant (csrc/cg2_march.cuh: each piece of the vectors is staged once per
 *                    strip instead of three times); 2: whenever it applies; 0: never;
 *                    "march_ok" (read-only) tells whether it applies, "march_lz" overrides the planes per run
 *   "trace"          n > 0: the loop kernels stamp %globaltimer into an 8-slot record per iteration for the
 *                    first n iterations of a solve; read it with cgb200_read_trace()
 */
CGB200_API int cgb200_set_option(cgb200_handle h, const char *key, long long value);
CGB200_API int cgb200_get_option(cgb200_handle h, const char *key, long long *value);

/* y = A x for k vectors: the `spmv` kernel alone (kernel/real/spmv.cl:5-50,
 * kernel/complex/spmv.cl:7-53).  Asynchronous when x and y are device pointers. */
CGB200_API int cgb200_spmv(cgb200_handle h, const void *x, void *y, int k, int layout);

/* The body of cg(): clcg.c:253-292 (q=Ax0, r=b-q, d=r, delta=r.r) and the loop
 * :296-419, for k right-hand sides.
 *   tol == 0   exactly max_iterations iterations (the reference's behaviour);
 *   tol  > 0   RHS c stops after the first iteration with sqrt(|delta|/|delta_0|) < tol,
 *              the recursive residual the reference already forms (clcg.c:384-391).
 * Optional outputs (host pointers or NULL):
 *   iterations[k]   iterations performed per RHS
 *   relres[k]       sqrt(|delta_final| / |delta_0|) per RHS
 *   delta_hist      (max_iterations+1) * k * (1|2) doubles: delta_new after the
 *                   initialisation and after every iteration (re, im for complex)
 * Blocks until x is complete. */
CGB200_API int cgb200_solve(cgb200_handle h, const void *b, void *x, int k, int max_iterations,
                 double tol, int *iterations, double *relres, double *delta_hist, int layout);

/* Jacobi-preconditioned CG -- the reference's PCG (helmFE_var.py:546-586) with M an inverse diagonal, which the
 * reference applies as a sparse matrix with one entry per row (:559-563); its report names preconditioning as
 * the path's own future work.  z = dinv*r ; rho = r.z ; p = z + (rho/rho_prev) p ; alpha = rho/(p.q).
 *   dinv   n values of the matrix dtype (host or device pointer), shared by all right-hand sides;
 *          NULL: 1 / diag(A) of the resident matrix, extracted on the device
 *   tol    ABSOLUTE, on sqrt|r.r| as the reference stops (:579-583); 0: exactly max_iterations iterations
 * Vectors in the cg() layout (RHS c at c*n); the k systems are solved one after the other.
 * Optional outputs: iterations[k] (iterations performed; the reference returns the index of the last one, i.e.
 * one less), resnorm[k] = sqrt|r.r| at exit, rr_hist = (max_iterations+1) * k * (1|2) doubles of r.r. */
CGB200_API int cgb200_solve_pcg(cgb200_handle h, const void *dinv, const void *b, void *x, int k, int max_iterations,
                                double tol, int *iterations, double *resnorm, double *rr_hist);

/* Milliseconds of the last cgb200_solve() on this handle, from CUDA events on its
 * stream: [0] inputs to HBM (+ layout change), [1] initialisation, [2] iterations,
 * [3] result back (+ layout change). */
CGB200_API int cgb200_last_timing(cgb200_handle h, double ms[4]);

/* Measurement hook: launches ONE kernel of the CG loop `reps` times back to back on the
 * handle's stream (after 3 warm-up launches) between two CUDA events and returns the mean
 * duration.  which: 0 spmv fused with d.q (spmv.cl + vdot.cl), 1 x/r update fused with r.r
 * (axpy.cl x2 + vdot.cl), 2 direction update (aypx.cl), 3 plain spmv, 4 / 5 the two kernels of the
 * two-kernel iteration (direction + spmv + x update + d.q; residual update + r.r).  Call it after a
 * cgb200_solve() with the same k; the next solve re-initialises the state it disturbs. */
CGB200_API int cgb200_time_kernel(cgb200_handle h, int which, int k, int reps, double *ms_avg);

/* Timeline of the last solve (option "trace"): out[it*8 + e], nanoseconds of the GPU's global timer, e =
 * 0 spmv starts, 1 all blocks of spmv done, 2 d.q known (after the all-reduce when sharded), 3 x/r update
 * starts, 4 all its blocks done, 5 delta known, 6 direction update starts, 7 halo entries of the peers
 * have arrived (sharded).  0 = not recorded. */
CGB200_API int cgb200_read_trace(cgb200_handle h, unsigned long long *out, int iterations);

/* Debug aid: the row-pattern dictionary as it sits in device memory.  which: 0 the 16-bit pattern number of
 * every row [n], 1 pattern lengths [patterns], 2 column offsets and 3 values [patterns][32]. */
CGB200_API int cgb200_debug_read_patterns(cgb200_handle h, int which, void *out, size_t bytes);

/* Debugging aid for pools where compute-sanitizer cannot run: with CGB200_GUARD=1 in the environment when a handle's
 * work vectors are allocated, each of them sits between two 4 KB zones filled with a byte pattern; this returns
 * the number of bytes a kernel changed there since (0 = no out-of-bounds store), or < 0 on failure. */
CGB200_API long long cgb200_check_guards(cgb200_handle h);

/* Host logic of the plane-marching dir_spmv (csrc/cg2_march.cuh), callable without a GPU: how `strips` x `planes` work
 * items are cut into runs for `blocks` thread blocks.  has_low / has_high: a halo plane below the first / above the last
 * owned plane (row-block shards); streaming: the vectors do not fit the L2; march_lz > 0: planes per run (option
 * "march_lz").  Writes up to `capacity` runs as {strip, first plane, planes, index of the same block's next run or -1}
 * to runs4, the launch grid to *grid (block b starts with run b), returns the number of runs or < 0. */
CGB200_API int cgb200_plan_march_runs(int strips, int planes, int blocks, int has_low, int has_high, int streaming,
                                      int march_lz, int *runs4, int capacity, int *grid);

/* Facts about a handle, for benches and tests:
 * [0] n [1] nnz [2] dtype [3] lanes_per_row [4] persistent grid of the SpMV kernel
 * [5] SM count [6] kernels launched so far [7] graph launches so far
 * [8] max row length [9] device ordinal */
CGB200_API int cgb200_info(cgb200_handle h, long long out[10]);

/* Device-side assembly of a constant-coefficient operator on a box grid (csrc/assemble.cuh) -- what the reference's
 * drivers build with O(n) Python loops on the host: local_rect (p_helmholtz.py:1342-1542), helmFE_var with a constant
 * wave speed (helmFE_var.py:9-331), Poisson (p_helmholtz.py:1545-1585).  Node (x, y, z) is row (z*ny + y)*nx + x; a
 * row is determined by the CLASS of its node, (cz*3 + cy)*3 + cx with c = 0 first / 1 interior / 2 last node of the
 * direction (the `if m == 0 and j == 0 ...` ladder of local_rect).
 *   class_len[27]           entries of each class's row (classes that cannot occur on the grid are ignored)
 *   class_dxyz[27][32][3]   neighbour offsets (dx, dy, dz) in {-1, 0, 1}, sorted by (dz, dy, dx) = ascending column
 *   class_val[27][32]       coefficients, matrix dtype
 * The CSR arrays are generated straight into HBM by one kernel (nothing of size n on the host or over PCIe) and the
 * handle is set up like an uploaded matrix.  cgb200_read_matrix copies the arrays out (any pointer may be NULL). */
CGB200_API int cgb200_create_grid(cgb200_handle *out, int dtype, int device, int nx, int ny, int nz,
                                  const int *class_len, const int *class_dxyz, const void *class_val);
CGB200_API int cgb200_read_matrix(cgb200_handle h, void *aValues, int *aPointers, int *aCols);

/* cg()/cgd() with the value type and the device spelled out (device < 0: the calling
 * thread's current CUDA device, or $CGB200_DEVICE).  This is what the pyopencl twin
 * of the reference, cl.py:44-200 `CG` / :203-360 `conjugate_gradient_multi_gpu`
 * (one context per device), maps onto.  Exactly nIterations iterations; x in/out. */
CGB200_API int cgb200_cg(int device, int dtype, int size, int nonZeros, const void *aValues,
                         const void *b, const int *aPointers, const int *aCols, void *x,
                         int nRHS, int nIterations);

/* Drop the matrix that cg()/cgd() keep resident between calls (keyed by content). */
CGB200_API int cgb200_clear_cache(void);

/* ---------------------------------------------------------------------------------
 * Row-block sharding across the GPUs of one box -- one process per GPU.
 *
 * The reference's multi-GPU script splits right-hand sides over devices and never
 * communicates (p_h-PY_C-CL-multi-GPU.py:2123-2181): that mode is cgb200_cg() /
 * cl.conjugate_gradient_multi_gpu() once per device.  These entry points are the other
 * mode: rank g owns rows [r_g, r_g+1) of A.  The caller (sharded.py) renumbers the
 * local columns as [owned 0..n_owned) | halo n_owned..n_owned+n_halo) with the halo
 * sorted by global index (hence grouped by owner rank), and says which owned entries
 * each peer needs.  Every iteration moves only those entries peer to peer (NCCL
 * send/recv over NVLink) and all-reduces the two dot scalars.  One right-hand side.
 * ------------------------------------------------------------------------------- */
typedef struct cgb200_shard_ctx *cgb200_shard;

/* ncclGetUniqueId() of rank 0, 128 bytes, to be broadcast to the other ranks. */
CGB200_API int cgb200_nccl_unique_id(void *out128);

/* aColsLocal: column indices in the local numbering.  send_counts[p] / recv_counts[p]:
 * entries sent to / received from rank p (0 for p == rank); send_idx: the owned indices
 * to send, concatenated in rank order; received blocks land in the halo in rank order.
 * row_boundary (optional, n_owned bytes): 1 for rows that reference a halo column; their SpMV tiles are
 * scheduled last so that the exchange is hidden behind the interior rows. */
CGB200_API int cgb200_shard_create(cgb200_shard *out, int rank, int world, const void *nccl_id128, int device,
                                   int n_owned, int n_halo, long long nnz, const void *aValues,
                                   const int *aPointers, const int *aColsLocal, int dtype,
                                   const int *send_counts, const int *send_idx, const int *recv_counts,
                                   const unsigned char *row_boundary);
CGB200_API int cgb200_shard_destroy(cgb200_shard sh);
/* The handle of the local row block (owned by the shard): for cgb200_info / cgb200_time_kernel. */
CGB200_API cgb200_handle cgb200_shard_local(cgb200_shard sh);
CGB200_API int cgb200_shard_set_stream(cgb200_shard sh, void *cuda_stream);
CGB200_API int cgb200_shard_set_option(cgb200_shard sh, const char *key, long long value);

/* Peer-memory collectives (optional, <= 8 ranks of one NVLink domain).  With them the halo entries are
 * written straight into the peers' vectors and the two dot products are all-reduced by the compute kernels
 * themselves through CUDA-IPC mapped pointers: no NCCL launch inside the iteration.
 *   export: this rank's blob of CGB200_P2P_BLOB_BYTES bytes: two 64-byte CUDA-IPC handles (exchange buffer;
 *           the allocation holding the direction and residual buffers) and the vectors' byte offsets in it
 *   import: all ranks' blobs (world x CGB200_P2P_BLOB_BYTES bytes, rank order) and, per peer p, the element
 *           offset in p's vectors where this rank's entries land (n_owned_p + p's receive offset for us)
 *   enable: switch between peer memory (1) and NCCL (0) afterwards */
#define CGB200_P2P_BLOB_BYTES 256
CGB200_API int cgb200_shard_p2p_export(cgb200_shard sh, void *out_blob);
CGB200_API int cgb200_shard_p2p_import(cgb200_shard sh, const void *all_blobs, const long long *remote_off);
CGB200_API int cgb200_shard_p2p_enable(cgb200_shard sh, int on);

/* Collective: every rank calls it with its slice of b and x (host or device pointers).
 * Same semantics as cgb200_solve() with k = 1; iterations / relres are global values. */
CGB200_API int cgb200_shard_solve(cgb200_shard sh, const void *b_owned, void *x_owned, int max_iterations,
                                  double tol, int *iterations, double *relres);

/* [0] n_owned [1] n_halo [2] entries sent per exchange [3] kernels launched [4] graph launches
 * [5] halo exchanges [6] all-reduces [7] local nnz */
CGB200_API int cgb200_shard_info(cgb200_shard sh, long long out[8]);
CGB200_API int cgb200_shard_last_timing(cgb200_shard sh, double ms[4]);

CGB200_API const char *cgb200_last_error(void);
CGB200_API int cgb200_device_count(void);
CGB200_API const char *cgb200_version(void);

#ifdef __cplusplus
}
#endif

#endif /* CGB200_H */

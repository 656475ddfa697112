/*
 * oracle/cpu_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the cg() hot path of ziyamammadov/conjugate-gradient-pyopencl:
 * a restatement of clcg.c:253-419 plus kernel/real and kernel/complex (*.cl) in plain C
 * (see cpu_ref_impl.h for the per-operation citations), instantiated for the
 * two precisions the reference runs (float, float complex) and the two
 * double-precision twins the north star asks parity in.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (liboclcg.so)
 * never links, loads or calls it.
 *
 * Parity status: the reference ships no golden vectors,
 * and its OpenCL path cannot run in this image (no ICD).  This oracle is pinned
 * instead (tests/test_oracle.py) against
 *   - helmFE_var.CG (helmFE_var.py:507-544) imported from the reference,
 *     through fixtures committed under tests/golden/ by oracle/make_golden.py;
 *   - the known answers of SURVEY.md section 8(c).
 *
 * Build:  make -C oracle     (gcc -O3 -fopenmp -ffp-contract=off)
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CPU_REF_WG 256    /* LOCAL_SIZE, clcg.c:37 */
#define CPU_REF_WAVE 32   /* WAVE_SIZE,  clcg.c:42 */

#define SFX f32
#define REAL float
#define CPLX 0
#include "cpu_ref_impl.h"
#undef SFX
#undef REAL
#undef CPLX

#define SFX f64
#define REAL double
#define CPLX 0
#include "cpu_ref_impl.h"
#undef SFX
#undef REAL
#undef CPLX

#define SFX c64
#define REAL float
#define CPLX 1
#include "cpu_ref_impl.h"
#undef SFX
#undef REAL
#undef CPLX

#define SFX c128
#define REAL double
#define CPLX 1
#include "cpu_ref_impl.h"
#undef SFX
#undef REAL
#undef CPLX

/* dtype codes shared with include/cgb200.h: 0 f32, 1 f64, 2 c64, 3 c128 */
int cpu_ref_cg(int dtype, int n, int nnz, const void *aValues, const void *b,
               const int *aPointers, const int *aCols, void *x, int k,
               int nIterations, double tol, int *iters, double *delta_hist) {
    switch (dtype) {
    case 0: return cpu_ref_cg_f32(n, nnz, aValues, b, aPointers, aCols, x, k, nIterations, tol, iters, delta_hist);
    case 1: return cpu_ref_cg_f64(n, nnz, aValues, b, aPointers, aCols, x, k, nIterations, tol, iters, delta_hist);
    case 2: return cpu_ref_cg_c64(n, nnz, aValues, b, aPointers, aCols, x, k, nIterations, tol, iters, delta_hist);
    case 3: return cpu_ref_cg_c128(n, nnz, aValues, b, aPointers, aCols, x, k, nIterations, tol, iters, delta_hist);
    }
    return -2;
}

int cpu_ref_spmv(int dtype, int n, const void *aValues, const int *aPointers,
                 const int *aCols, const void *x, void *y, int k) {
    switch (dtype) {
    case 0: return cpu_ref_spmv_f32(n, aValues, aPointers, aCols, x, y, k);
    case 1: return cpu_ref_spmv_f64(n, aValues, aPointers, aCols, x, y, k);
    case 2: return cpu_ref_spmv_c64(n, aValues, aPointers, aCols, x, y, k);
    case 3: return cpu_ref_spmv_c128(n, aValues, aPointers, aCols, x, y, k);
    }
    return -2;
}

/* The reference's own entry point shape (clcg.h:3-5), served by the oracle. */
int cpu_ref_pcg(int dtype, int n, const void *aValues, const void *b, const int *aPointers, const int *aCols,
                void *x, const void *dinv, int k, int nIterations, double *rr_hist) {
    switch (dtype) {
    case 0: return cpu_ref_pcg_f32(n, aValues, b, aPointers, aCols, x, dinv, k, nIterations, rr_hist);
    case 1: return cpu_ref_pcg_f64(n, aValues, b, aPointers, aCols, x, dinv, k, nIterations, rr_hist);
    case 2: return cpu_ref_pcg_c64(n, aValues, b, aPointers, aCols, x, dinv, k, nIterations, rr_hist);
    case 3: return cpu_ref_pcg_c128(n, aValues, b, aPointers, aCols, x, dinv, k, nIterations, rr_hist);
    }
    return -2;
}

float *cpu_ref_cg_legacy(int size, int nonZeros, const float *aValues, const float *b,
                         const int *aPointers, const int *aCols, float *x,
                         int nRHS, int nIterations, int isComplex) {
    cpu_ref_cg(isComplex ? 2 : 0, size, nonZeros, aValues, b, aPointers, aCols, x,
               nRHS, nIterations, 0.0, NULL, NULL);
    return x;
}

int cpu_ref_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void cpu_ref_set_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/*
 * include/clcg.h -- the drop-in boundary.
 *
 * `cg` is, bit for bit, the prototype the reference declares in
 * /root/reference/clcg.h:3-5 and defines in /root/reference/clcg.c:111-466.
 * It is what `main.c:56` links against and what the Python drivers bind with
 * ctypes (`CDLL("./build/liboclcg.so").cg`, p_h-PY_C-CL.py:38,1948-1950;
 * `CDLL("./liboclcg.so")`, p_helmholtz.py:29).  liboclcg.so built from this
 * repo exports it with the same name, argument order and semantics, served by
 * hand-written sm_100a CUDA kernels instead of OpenCL.
 *
 *   size         rows of the square matrix A
 *   nonZeros     stored entries of A
 *   aValues      nonZeros values: float if !isComplex, else interleaved (re, im) float pairs
 *   b            nRHS blocks of `size` values; right-hand side r starts at b + r*size
 *   aPointers    size+1 CSR row offsets, 0-based, int32
 *   aCols        nonZeros column indices, 0-based, any order inside a row, int32;
 *                symmetric matrices must be fully stored
 *   x            same layout as b; IN: initial guess, OUT: the iterate after nIterations
 *   nRHS         number of right-hand sides solved together (independent CGs sharing A)
 *   nIterations  exactly this many CG iterations are performed (no convergence test)
 *   isComplex    0 real, non-zero complex (unconjugated dot products: COCG for
 *                complex-symmetric A, as kernel/complex/vdot.cl:15)
 *
 * Returns x.  Every pointer is caller-owned; nothing is retained after return
 * (an internal device-side copy of the last matrix may be kept for reuse, keyed by
 * content).  Unlike the reference, failures never call exit():
 * cg() returns NULL and cgb200_last_error() explains.
 *
 * `cgd` is the same entry point in double precision (double / double complex),
 * which the reference does not have ("Can't handle double precision yet",
 * main.c:49) and BASELINE.json's parity bar (1e-10) needs.
 */
#ifndef OCLCG_CLCG_H
#define OCLCG_CLCG_H

#ifndef CGB200_API
#define CGB200_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

CGB200_API float *cg(int size, int nonZeros,
          const float *aValues, const float *b, const int *aPointers,
          const int *aCols, float *x, int nRHS, int nIterations, int isComplex);

CGB200_API double *cgd(int size, int nonZeros,
            const double *aValues, const double *b, const int *aPointers,
            const int *aCols, double *x, int nRHS, int nIterations, int isComplex);

#ifdef __cplusplus
}
#endif

#endif /* OCLCG_CLCG_H */

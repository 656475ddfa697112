/*
 * oracle/cpu_ref_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Type-generic body of the CPU oracle; included four times by cpu_ref.c with
 *   SFX   function suffix (f32, f64, c64, c128)
 *   REAL  float | double
 *   CPLX  0 | 1
 *
 * It restates, operation by operation and in the same floating-point
 * summation ORDER, what the reference computes on an OpenCL device:
 *
 *   spmv   kernel/real/spmv.cl:5-50,   kernel/complex/spmv.cl:7-53
 *   vdot   kernel/real/vdot.cl:2-38,   kernel/complex/vdot.cl:4-41
 *   axpy   kernel/real/axpy.cl:2-17,   kernel/complex/axpy.cl:4-22
 *   aypx   kernel/real/aypx.cl:2-10,   kernel/complex/aypx.cl:4-12
 *   sub    kernel/real/sub.cl:2-12,    kernel/complex/sub.cl:4-15
 *   complex arithmetic                 kernel/complex/cmplx.h:4-25
 *   host orchestration, host partial sums, alpha/beta   clcg.c:253-419
 *
 * Pinned: bit-identical (float, float complex; 1..4 right-hand sides) to the reference's own kernel sources
 * executed by oracle/clref -- tests/test_oracle.py, tests/golden/clref_*.npz.
 *
 * Deliberate differences (documented in DESIGN.md):
 *   - the out-of-bounds row-pointer read of spmv.cl:18-19 is not reproduced;
 *   - size < 256 (the reference prints "NOT SUPPORTED", clcg.c:123) is handled
 *     as one zero-padded 256-wide work-group;
 *   - a*b+c is never contracted to an FMA (build with -ffp-contract=off);
 *   - an optional tolerance / delta history (tol = 0 -> reference behaviour).
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

#if CPLX
typedef struct { REAL re, im; } FN(val_t);
#define VAL FN(val_t)
static inline VAL FN(vadd)(VAL a, VAL b) { VAL c = { a.re + b.re, a.im + b.im }; return c; }   /* cmplx.h:6-11  */
static inline VAL FN(vsub)(VAL a, VAL b) { VAL c = { a.re - b.re, a.im - b.im }; return c; }   /* cmplx.h:13-18 */
static inline VAL FN(vmul)(VAL a, VAL b) {                                                     /* cmplx.h:20-25 */
    VAL c = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return c;
}
/* host-side C99 complex division, clcg.c:326 and :390 */
static inline VAL FN(vdiv)(VAL a, VAL b) {
    REAL _Complex q = (a.re + a.im * (REAL _Complex)_Complex_I) / (b.re + b.im * (REAL _Complex)_Complex_I);
    VAL c = { __real__ q, __imag__ q };
    return c;
}
static inline double FN(vabs)(VAL a) { return hypot((double)a.re, (double)a.im); }
static inline VAL FN(vzero)(void) { VAL c = { 0, 0 }; return c; }
#define NCOMP 2
#else
typedef REAL FN(val_t);
#define VAL FN(val_t)
static inline VAL FN(vadd)(VAL a, VAL b) { return a + b; }
static inline VAL FN(vsub)(VAL a, VAL b) { return a - b; }
static inline VAL FN(vmul)(VAL a, VAL b) { return a * b; }
static inline VAL FN(vdiv)(VAL a, VAL b) { return a / b; }
static inline double FN(vabs)(VAL a) { return fabs((double)a); }
static inline VAL FN(vzero)(void) { return 0; }
#define NCOMP 1
#endif

#define VADD FN(vadd)
#define VSUB FN(vsub)
#define VMUL FN(vmul)
#define VDIV FN(vdiv)
#define VABS FN(vabs)
#define VZERO FN(vzero)

static inline void FN(store_hist)(double *h, VAL v) {
#if CPLX
    h[0] = (double)v.re; h[1] = (double)v.im;
#else
    h[0] = (double)v;
#endif
}

/* Adjacent-pair tree over a power-of-two array, the order of vdot.cl:20-29
 * and spmv.cl:32-43: offset 1,2,4,...; element i (i % 2*offset == 0) += element i+offset. */
static inline VAL FN(pair_tree)(VAL *s, int width) {
    for (int off = 1; off < width; off <<= 1)
        for (int i = 0; i + off < width; i += 2 * off)
            s[i] = VADD(s[i], s[i + off]);
    return s[0];
}

/* y[:, r] = A x[:, r]; one 32-lane wave per row. spmv.cl:13-49 */
static void FN(spmv)(int n, const VAL *av, const int *ap, const int *ac,
                     const VAL *x, VAL *y, int k) {
#pragma omp parallel for schedule(static)
    for (int row = 0; row < n; row++) {
        const int lo = ap[row], hi = ap[row + 1];
        for (int r = 0; r < k; r++) {
            const VAL *xr = x + (size_t)r * n;
            VAL lane[CPU_REF_WAVE];
            if (hi - lo <= CPU_REF_WAVE) {
                /* every lane holds at most one product; 0 + p is exact */
                int l = 0;
                for (int j = lo; j < hi; j++, l++)
                    lane[l] = VADD(VZERO(), VMUL(av[j], xr[ac[j]]));
                for (; l < CPU_REF_WAVE; l++) lane[l] = VZERO();
            } else {
                for (int l = 0; l < CPU_REF_WAVE; l++) {
                    VAL s = VZERO();
                    for (int j = lo + l; j < hi; j += CPU_REF_WAVE)
                        s = VADD(s, VMUL(av[j], xr[ac[j]]));
                    lane[l] = s;
                }
            }
            y[(size_t)r * n + row] = FN(pair_tree)(lane, CPU_REF_WAVE);
        }
    }
}

/* out[r] (+)= sum over work-groups (sequential, clcg.c:276-278/321-323/384-386)
 * of the 256-wide adjacent-pair tree of a[i]*b[i] (vdot.cl:11-37).
 * `part` is scratch of wgs*k values. The host sum starts from out[r] as given. */
static void FN(vdot)(int n, const VAL *a, const VAL *b, int k, VAL *part, VAL *out) {
    const int wgs = 1 + (n - 1) / CPU_REF_WG;            /* clcg.c:124 */
#pragma omp parallel for schedule(static)
    for (int w = 0; w < wgs; w++) {
        for (int r = 0; r < k; r++) {
            VAL loc[CPU_REF_WG];
            const size_t base = (size_t)r * n;
            for (int t = 0; t < CPU_REF_WG; t++) {
                const int i = w * CPU_REF_WG + t;
                loc[t] = (i < n) ? VMUL(a[base + i], b[base + i]) : VZERO();
            }
            part[(size_t)r * wgs + w] = FN(pair_tree)(loc, CPU_REF_WG);
        }
    }
    for (int r = 0; r < k; r++)
        for (int w = 0; w < wgs; w++)
            out[r] = VADD(out[r], part[(size_t)r * wgs + w]);
}

/* y (+|-)= a[r] * x   axpy.cl:8-16 */
static void FN(axpy)(int n, const VAL *x, VAL *y, const VAL *a, int plus, int k) {
    for (int r = 0; r < k; r++) {
        const VAL ar = a[r];
        const VAL *xr = x + (size_t)r * n;
        VAL *yr = y + (size_t)r * n;
        if (plus) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; i++) yr[i] = VADD(yr[i], VMUL(ar, xr[i]));
        } else {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; i++) yr[i] = VSUB(yr[i], VMUL(ar, xr[i]));
        }
    }
}

/* y = a[r] * y + x   aypx.cl:6-9 (real: y*a + x; complex: cmul(a, y) + x) */
static void FN(aypx)(int n, const VAL *x, VAL *y, const VAL *a, int k) {
    for (int r = 0; r < k; r++) {
        const VAL ar = a[r];
        const VAL *xr = x + (size_t)r * n;
        VAL *yr = y + (size_t)r * n;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) yr[i] = VADD(VMUL(ar, yr[i]), xr[i]);
    }
}

/* res = a - b   sub.cl:6-11 */
static void FN(vsubv)(size_t len, const VAL *a, const VAL *b, VAL *res) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < len; i++) res[i] = VSUB(a[i], b[i]);
}

/*
 * The CG driver. clcg.c:253-292 (init) and :296-419 (loop).
 *
 * tol == 0: exactly nIterations iterations for every RHS (the reference).
 * tol  > 0: RHS r stops being updated after the first iteration `it` with
 *           sqrt(|delta_it| / |delta_0|) < tol; iters[r] = it.
 * delta_hist (optional): (nIterations+1) * k * NCOMP doubles, delta_new after
 *           init (slot 0) and after every iteration.
 * Returns 0, or -1 on allocation failure.
 */
int FN(cpu_ref_cg)(int n, int nnz, const void *aValues, const void *bValues,
                   const int *aPointers, const int *aCols, void *xInOut,
                   int k, int nIterations, double tol, int *iters, double *delta_hist) {
    (void)nnz;
    const VAL *av = (const VAL *)aValues;
    const VAL *b = (const VAL *)bValues;
    VAL *x = (VAL *)xInOut;
    const size_t len = (size_t)n * k;
    const int wgs = 1 + (n - 1) / CPU_REF_WG;

    VAL *r = malloc(len * sizeof(VAL)), *d = malloc(len * sizeof(VAL)), *q = malloc(len * sizeof(VAL));
    VAL *part = malloc((size_t)wgs * k * sizeof(VAL));
    VAL *dNew = malloc(k * sizeof(VAL)), *dOld = malloc(k * sizeof(VAL)), *dq = malloc(k * sizeof(VAL));
    VAL *alpha = malloc(k * sizeof(VAL)), *beta = malloc(k * sizeof(VAL));
    double *d0 = malloc(k * sizeof(double));
    char *frozen = calloc(k, 1);
    if (!r || !d || !q || !part || !dNew || !dOld || !dq || !alpha || !beta || !d0 || !frozen) return -1;

    /* q = A x0 ; r = b - q ; d = r ; delta_new = r.r      clcg.c:255-279 */
    FN(spmv)(n, av, aPointers, aCols, x, q, k);
    FN(vsubv)(len, b, q, r);
    memcpy(d, r, len * sizeof(VAL));
    for (int c = 0; c < k; c++) dNew[c] = VZERO();
    FN(vdot)(n, r, r, k, part, dNew);
    for (int c = 0; c < k; c++) {
        dOld[c] = dNew[c];                                   /* clcg.c:282/289 */
        d0[c] = VABS(dNew[c]);
        if (iters) iters[c] = nIterations;
        if (delta_hist) FN(store_hist)(delta_hist + (size_t)c * NCOMP, dNew[c]);
        if (tol > 0 && d0[c] == 0.0) { frozen[c] = 1; if (iters) iters[c] = 0; }
    }

    for (int it = 0; it < nIterations; it++) {
        FN(spmv)(n, av, aPointers, aCols, d, q, k);          /* clcg.c:299-305 */
        for (int c = 0; c < k; c++) dq[c] = VZERO();         /* clcg.c:318-319 */
        FN(vdot)(n, d, q, k, part, dq);                      /* clcg.c:309-324 */
        for (int c = 0; c < k; c++)
            alpha[c] = frozen[c] ? VZERO() : VDIV(dNew[c], dq[c]);   /* clcg.c:326-327 */
        FN(axpy)(n, d, x, alpha, 1, k);                      /* clcg.c:338-342 */
        FN(axpy)(n, q, r, alpha, 0, k);                      /* clcg.c:345-349 */
        for (int c = 0; c < k; c++) {                        /* clcg.c:350-356 */
            if (frozen[c]) continue;
            dOld[c] = dNew[c];
            dNew[c] = VZERO();
        }
        {
            /* frozen columns must keep their delta; reduce into a scratch copy */
            VAL *acc = beta; /* reuse as scratch until beta is formed */
            for (int c = 0; c < k; c++) acc[c] = VZERO();
            FN(vdot)(n, r, r, k, part, acc);                 /* clcg.c:369-387 */
            for (int c = 0; c < k; c++) if (!frozen[c]) dNew[c] = acc[c];
        }
        for (int c = 0; c < k; c++)
            beta[c] = frozen[c] ? VZERO() : VDIV(dNew[c], dOld[c]);  /* clcg.c:389-391 */
        if (delta_hist)
            for (int c = 0; c < k; c++)
                FN(store_hist)(delta_hist + ((size_t)(it + 1) * k + c) * NCOMP, dNew[c]);
        if (tol > 0) {
            int live = 0;
            for (int c = 0; c < k; c++) {
                if (!frozen[c] && sqrt(VABS(dNew[c]) / d0[c]) < tol) {
                    frozen[c] = 1;
                    if (iters) iters[c] = it + 1;
                }
                live += !frozen[c];
            }
            if (!live) {
                /* remaining history slots repeat the final value */
                if (delta_hist)
                    for (int j = it + 2; j <= nIterations; j++)
                        memcpy(delta_hist + (size_t)j * k * NCOMP,
                               delta_hist + (size_t)(it + 1) * k * NCOMP, (size_t)k * NCOMP * sizeof(double));
                break;
            }
        }
        /* d = beta d + r   clcg.c:411-415.  A frozen column has beta = 0 and is never read again. */
        FN(aypx)(n, r, d, beta, k);
    }

    free(r); free(d); free(q); free(part); free(dNew); free(dOld); free(dq);
    free(alpha); free(beta); free(d0); free(frozen);
    return 0;
}

/*
 * Jacobi-preconditioned CG -- the oracle for SURVEY.md 8(f) rank 2 (not yet on the device).  The recurrence of the
 * reference's PCG (helmFE_var.py:546-586: z = M r, rho = r.z, p = z + (rho/rho_prev) p, alpha = rho / p.q) arranged
 * as the device will run it, with the kernels and summation orders above:
 *     q = A x0 ; r = b - q ; z = dinv o r ; d = z ; rho = r.z
 *     loop:  q = A d ; alpha = rho / d.q ; x += alpha d ; r -= alpha q ; z = dinv o r ; rho' = r.z ;
 *            d = z + (rho'/rho) d
 * dinv: n values (the inverse diagonal), shared by the k right-hand sides.  Exactly nIterations iterations;
 * rr_hist (optional, (nIterations+1)*k*NCOMP doubles): r.r after the initialisation and after every iteration,
 * the quantity the reference's PCG stops on (sqrt|r.r| < tol, :579-583).
 */
int FN(cpu_ref_pcg)(int n, const void *aValues, const void *bValues, const int *aPointers, const int *aCols,
                    void *xInOut, const void *dinv_, int k, int nIterations, double *rr_hist) {
    const VAL *av = (const VAL *)aValues, *b = (const VAL *)bValues, *dinv = (const VAL *)dinv_;
    VAL *x = (VAL *)xInOut;
    const size_t len = (size_t)n * k;
    const int wgs = 1 + (n - 1) / CPU_REF_WG;
    VAL *r = malloc(len * sizeof(VAL)), *d = malloc(len * sizeof(VAL)), *q = malloc(len * sizeof(VAL));
    VAL *z = malloc(len * sizeof(VAL)), *part = malloc((size_t)wgs * k * sizeof(VAL));
    VAL *rho = malloc(k * sizeof(VAL)), *rhoNew = malloc(k * sizeof(VAL)), *dq = malloc(k * sizeof(VAL));
    VAL *alpha = malloc(k * sizeof(VAL)), *beta = malloc(k * sizeof(VAL)), *rr = malloc(k * sizeof(VAL));
    if (!r || !d || !q || !z || !part || !rho || !rhoNew || !dq || !alpha || !beta || !rr) return -1;
#define PCG_APPLY_M()                                                            \
    for (int c = 0; c < k; c++)                                                  \
        for (int i = 0; i < n; i++) z[(size_t)c * n + i] = VMUL(dinv[i], r[(size_t)c * n + i]);
#define PCG_RR(slot)                                                             \
    if (rr_hist) {                                                               \
        for (int c = 0; c < k; c++) rr[c] = VZERO();                             \
        FN(vdot)(n, r, r, k, part, rr);                                          \
        for (int c = 0; c < k; c++) FN(store_hist)(rr_hist + ((size_t)(slot) * k + c) * NCOMP, rr[c]); \
    }
    FN(spmv)(n, av, aPointers, aCols, x, q, k);
    FN(vsubv)(len, b, q, r);
    PCG_APPLY_M();
    memcpy(d, z, len * sizeof(VAL));
    for (int c = 0; c < k; c++) rho[c] = VZERO();
    FN(vdot)(n, r, z, k, part, rho);
    PCG_RR(0);
    for (int it = 0; it < nIterations; it++) {
        FN(spmv)(n, av, aPointers, aCols, d, q, k);
        for (int c = 0; c < k; c++) dq[c] = VZERO();
        FN(vdot)(n, d, q, k, part, dq);
        for (int c = 0; c < k; c++) alpha[c] = VDIV(rho[c], dq[c]);
        FN(axpy)(n, d, x, alpha, 1, k);
        FN(axpy)(n, q, r, alpha, 0, k);
        PCG_APPLY_M();
        for (int c = 0; c < k; c++) rhoNew[c] = VZERO();
        FN(vdot)(n, r, z, k, part, rhoNew);
        for (int c = 0; c < k; c++) {
            beta[c] = VDIV(rhoNew[c], rho[c]);
            rho[c] = rhoNew[c];
        }
        PCG_RR(it + 1);
        FN(aypx)(n, z, d, beta, k);          /* d = beta d + z */
    }
#undef PCG_APPLY_M
#undef PCG_RR
    free(r); free(d); free(q); free(z); free(part); free(rho); free(rhoNew); free(dq); free(alpha); free(beta); free(rr);
    return 0;
}

/* y = A x with the reference's summation order; exported for kernel-level parity tests. */
int FN(cpu_ref_spmv)(int n, const void *aValues, const int *aPointers, const int *aCols,
                     const void *x, void *y, int k) {
    FN(spmv)(n, (const VAL *)aValues, aPointers, aCols, (const VAL *)x, (VAL *)y, k);
    return 0;
}

#undef VAL
#undef VADD
#undef VSUB
#undef VMUL
#undef VDIV
#undef VABS
#undef VZERO
#undef NCOMP
#undef CAT_
#undef CAT
#undef FN

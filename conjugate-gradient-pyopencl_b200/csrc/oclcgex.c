/*
 * oclcgex.c -- the reference's example executable (/root/reference/main.c:13-61) on liboclcg.so.
 *
 *     ./oclcgex <input matrix file> <number of RHS> <is complex> <number of iterations> [--double]
 *
 * Same four arguments and the same flow (main.c:15-56):
 *     load a Matrix Market file                      main.c:20   (BeBOP SMC there -- not vendored by the
 *                                                                reference, README.md:31 -- a reader of the
 *                                                                coordinate format here)
 *     expand symmetric storage to full storage       main.c:25
 *     convert to CSR                                 main.c:27
 *     b[r*n + i] = 5 (r + 1), x0 = 0                 main.c:41-46
 *     narrow the values to single precision          main.c:49-53
 *     cg(n, nnz, aValues, b, rowptr, colidx, x, nRHS, nIterations, isComplex)      main.c:56
 * main.c allocates complex buffers whatever <is complex> says and is therefore only correct for 1; here
 * <is complex> = 0 on a real file runs the real path on real buffers.  Unlike main.c the relative residual
 * |b - A x| / |b| of every right-hand side is printed (computed here on the host, in double).
 * --double solves through cgd().  `oclcgex <file> --dump-csr` prints the full-storage CSR matrix the file expands
 * to (no device work): the Matrix Market reader can be checked without a GPU.
 */
#include <complex.h>
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/clcg.h"

extern const char *cgb200_last_error(void);

typedef struct {
    int n;
    long long nnz;
    int *rowptr, *colidx;
    double complex *values;
    int is_complex;      /* the file's field was `complex` */
} csr_t;

typedef struct {
    int r, c;
    double complex v;
} entry_t;

static int entry_cmp(const void *a, const void *b) {
    const entry_t *x = (const entry_t *)a, *y = (const entry_t *)b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return 0;
}

static void lower(char *s) {
    for (; *s; s++) *s = (char)tolower((unsigned char)*s);
}

/* Matrix Market coordinate format -> full-storage CSR with sorted rows, duplicates summed. */
static int load_matrix_market(const char *path, csr_t *out) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1024], banner[64], object[64], format[64], field[64], symmetry[64];
    if (!fgets(line, sizeof line, f) ||
        sscanf(line, "%63s %63s %63s %63s %63s", banner, object, format, field, symmetry) != 5) {
        fclose(f);
        return -1;
    }
    lower(object); lower(format); lower(field); lower(symmetry);
    if (strcmp(banner, "%%MatrixMarket") || strcmp(object, "matrix") || strcmp(format, "coordinate")) {
        fclose(f);
        return -1;
    }
    const int is_complex = !strcmp(field, "complex"), is_pattern = !strcmp(field, "pattern");
    if (!is_complex && !is_pattern && strcmp(field, "real") && strcmp(field, "integer") && strcmp(field, "double")) {
        fclose(f);
        return -1;
    }
    int sym = 0;            /* 0 general, 1 symmetric, 2 hermitian, 3 skew-symmetric */
    if (!strcmp(symmetry, "symmetric")) sym = 1;
    else if (!strcmp(symmetry, "hermitian")) sym = 2;
    else if (!strcmp(symmetry, "skew-symmetric")) sym = 3;
    else if (strcmp(symmetry, "general")) {
        fclose(f);
        return -1;
    }
    do {
        if (!fgets(line, sizeof line, f)) {
            fclose(f);
            return -1;
        }
    } while (line[0] == '%' || line[0] == '\n' || line[0] == '\r');
    long long rows, cols, stored;
    if (sscanf(line, "%lld %lld %lld", &rows, &cols, &stored) != 3 || rows != cols || rows <= 0 || stored < 0 ||
        rows > 2147483647LL) {
        fclose(f);
        return -1;
    }
    entry_t *e = (entry_t *)malloc((size_t)(2 * stored + 1) * sizeof(entry_t));
    if (!e) {
        fclose(f);
        return -1;
    }
    long long m = 0;
    for (long long i = 0; i < stored; i++) {
        long long r, c;
        double re = 1.0, im = 0.0;
        int got;
        if (is_pattern) got = fscanf(f, "%lld %lld", &r, &c) == 2;
        else if (is_complex) got = fscanf(f, "%lld %lld %lf %lf", &r, &c, &re, &im) == 4;
        else got = fscanf(f, "%lld %lld %lf", &r, &c, &re) == 3;
        if (!got || r < 1 || c < 1 || r > rows || c > rows) {
            free(e);
            fclose(f);
            return -1;
        }
        e[m].r = (int)(r - 1);
        e[m].c = (int)(c - 1);
        e[m].v = re + im * I;
        m++;
        if (sym && r != c) {        /* main.c:25 -- expand to full storage */
            e[m].r = (int)(c - 1);
            e[m].c = (int)(r - 1);
            e[m].v = sym == 1 ? re + im * I : (sym == 2 ? re - im * I : -(re + im * I));
            m++;
        }
    }
    fclose(f);
    qsort(e, (size_t)m, sizeof(entry_t), entry_cmp);
    long long u = 0;                /* sum duplicates */
    for (long long i = 0; i < m; i++) {
        if (u > 0 && e[u - 1].r == e[i].r && e[u - 1].c == e[i].c) e[u - 1].v += e[i].v;
        else e[u++] = e[i];
    }
    if (u > 2147483647LL) {
        free(e);
        return -1;
    }
    out->n = (int)rows;
    out->nnz = u;
    out->is_complex = is_complex;
    out->rowptr = (int *)calloc((size_t)rows + 1, sizeof(int));
    out->colidx = (int *)malloc((size_t)(u ? u : 1) * sizeof(int));
    out->values = (double complex *)malloc((size_t)(u ? u : 1) * sizeof(double complex));
    if (!out->rowptr || !out->colidx || !out->values) {
        free(e);
        return -1;
    }
    for (long long i = 0; i < u; i++) {      /* main.c:27 -- CSR */
        out->rowptr[e[i].r + 1]++;
        out->colidx[i] = e[i].c;
        out->values[i] = e[i].v;
    }
    for (int r = 0; r < out->n; r++) out->rowptr[r + 1] += out->rowptr[r];
    free(e);
    return 0;
}

int main(int argc, char *argv[]) {
    int use_double = 0, npos = 0, dump = 0;
    char *pos[4];
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--double")) use_double = 1;
        else if (!strcmp(argv[i], "--dump-csr")) dump = 1;
        else if (npos < 4) pos[npos++] = argv[i];
        else npos++;
    }
    if (dump && npos == 1) {      /* the matrix as cg() would receive it, no device work: "n nnz" then "row col re im" */
        csr_t m;
        if (load_matrix_market(pos[0], &m)) {
            printf("Could not read matrix\n");
            return 1;
        }
        printf("%d %lld %d\n", m.n, m.nnz, m.is_complex);
        for (int r = 0; r < m.n; r++)
            for (int j = m.rowptr[r]; j < m.rowptr[r + 1]; j++)
                printf("%d %d %.17g %.17g\n", r, m.colidx[j], creal(m.values[j]), cimag(m.values[j]));
        return 0;
    }
    if (npos != 4) {      /* main.c:15-18 */
        fprintf(stderr, "Usage: ./CG <input matrix file> <number of RHS> <is complex> <number of iterations>\n");
        return 1;
    }
    csr_t a;
    if (load_matrix_market(pos[0], &a)) {       /* main.c:20-24 */
        printf("Could not read matrix\n");
        return 1;
    }
    const int nRHS = atoi(pos[1]), isComplex = atoi(pos[2]), nIterations = atoi(pos[3]);
    if (nRHS < 1 || nIterations < 0) {
        fprintf(stderr, "number of RHS must be >= 1 and number of iterations >= 0\n");
        return 1;
    }
    if (a.is_complex && !isComplex) {
        printf("matrix is complex: pass <is complex> = 1\n");
        return 1;
    }
    const int n = a.n;
    const long long nnz = a.nnz;
    const size_t comp = isComplex ? 2 : 1, real_bytes = use_double ? sizeof(double) : sizeof(float);
    void *vals = malloc((size_t)(nnz ? nnz : 1) * comp * real_bytes);
    void *b = malloc((size_t)nRHS * n * comp * real_bytes);
    void *x = calloc((size_t)nRHS * n * comp, real_bytes);           /* x0 = 0, main.c:43 */
    if (!vals || !b || !x) {
        fprintf(stderr, "out of memory\n");
        return 1;
    }
    for (long long i = 0; i < nnz; i++) {        /* main.c:50-53 */
        for (size_t c = 0; c < comp; c++) {
            const double v = c ? cimag(a.values[i]) : creal(a.values[i]);
            if (use_double) ((double *)vals)[i * comp + c] = v;
            else ((float *)vals)[i * comp + c] = (float)v;
        }
    }
    for (int r = 0; r < nRHS; r++) {             /* main.c:41-46 */
        for (int i = 0; i < n; i++) {
            for (size_t c = 0; c < comp; c++) {
                const double v = c ? 0.0 : (r + 1) * 5.0;
                const size_t at = ((size_t)r * n + i) * comp + c;
                if (use_double) ((double *)b)[at] = v;
                else ((float *)b)[at] = (float)v;
            }
        }
    }
    const void *ret = use_double        /* main.c:56 */
        ? (const void *)cgd(n, (int)nnz, (const double *)vals, (const double *)b, a.rowptr, a.colidx, (double *)x, nRHS, nIterations, isComplex)
        : (const void *)cg(n, (int)nnz, (const float *)vals, (const float *)b, a.rowptr, a.colidx, (float *)x, nRHS, nIterations, isComplex);
    if (!ret) {
        fprintf(stderr, "cg failed: %s\n", cgb200_last_error());
        return 2;
    }
    for (int r = 0; r < nRHS; r++) {
        double rr = 0.0, bb = 0.0;
        for (int i = 0; i < n; i++) {
            double complex s = 0.0;
            for (int j = a.rowptr[i]; j < a.rowptr[i + 1]; j++) {
                const size_t at = ((size_t)r * n + a.colidx[j]) * comp;
                const double xr = use_double ? ((double *)x)[at] : ((float *)x)[at];
                const double xi = isComplex ? (use_double ? ((double *)x)[at + 1] : ((float *)x)[at + 1]) : 0.0;
                s += a.values[j] * (xr + xi * I);
            }
            const double bi = (r + 1) * 5.0;
            rr += creal((bi - s) * conj(bi - s));
            bb += bi * bi;
        }
        printf("rhs %d: relative residual %.3e after %d iterations\n", r, sqrt(rr / bb), nIterations);
    }
    free(vals); free(b); free(x); free(a.rowptr); free(a.colidx); free(a.values);
    return 0;
}

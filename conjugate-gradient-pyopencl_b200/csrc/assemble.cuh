// assemble.cuh -- device-side assembly of constant-coefficient grid operators (included by cgb200.cu).
//
// The matrices the reference's drivers solve with are assembled on the host by O(n) Python loops over the grid
// nodes: `local_rect` (p_helmholtz.py:1342-1542, the subdomain operators of as_prec), `helmFE_var`
// (helmFE_var.py:9-331) and `Poisson` (p_helmholtz.py:1545-1585) -- about 12 s at one million unknowns, after
// which the arrays still have to cross PCIe.  All of them have the same structure: on a lexicographically
// numbered box grid a row is determined by the CLASS of its node -- first / interior / last in every coordinate
// direction, 9 classes in 2-D and 27 in 3-D (the `if m == 0 and j == 0 ...` ladder of local_rect) -- and every
// class has one fixed list of (neighbour offset, coefficient) pairs.
//
// cgb200_create_grid takes that class table (a few hundred bytes, evaluated by the caller in the reference's own
// arithmetic so that the coefficients agree to the last bit) and generates the CSR arrays -- row offsets in
// closed form, column indices, values -- directly in the handle's HBM buffers with one kernel.  Nothing of size n
// exists on the host, nothing of size n is uploaded.  The arrays then go through the same set-up as uploaded
// ones (validation, SpMV schedule, row-pattern dictionary built on the device), so every solver path works on
// an assembled matrix, and cgb200_read_matrix returns the arrays for inspection.
#pragma once

constexpr int GRID_CLASSES = 27;             // (cz*3 + cy)*3 + cx, c = 0 first, 1 interior, 2 last node of the direction
constexpr int GRID_MAXLEN = PAT_MAXLEN;      // entries per class

struct GridSpec {
    int nx, ny, nz;                          // grid points per direction (nz = 1: two-dimensional); x runs fastest
    int len[GRID_CLASSES];                   // entries of every class's row
    const int *dxyz;                         // host, [27][GRID_MAXLEN][3]: neighbour (dx, dy, dz), sorted by (dz, dy, dx)
    const void *val;                         // host, [27][GRID_MAXLEN] coefficients of the matrix dtype
};

struct GridCounts {                          // closed-form row offsets
    int cnt[GRID_CLASSES];
    long long line[9];                       // [cz*3 + cy]: non-zeros of one grid line
    long long plane[3];                      // [cz]: non-zeros of one grid plane
};

__host__ __device__ inline int grid_class(int v, int size) { return v == 0 ? 0 : (v == size - 1 ? 2 : 1); }
// sum of per-class amounts a[0], a[1], a[2] over the positions [0, upto) of a direction with `size` points
__host__ __device__ inline long long grid_prefix(const long long a[3], int upto, int size) {
    if (upto <= 0) return 0;
    long long s = a[0];                                                  // position 0
    const int mids = (upto - 1 < size - 2 ? upto - 1 : size - 2);
    if (mids > 0) s += (long long)mids * a[1];
    if (upto >= size && size >= 2) s += a[2];
    return s;
}

static void grid_counts(const GridSpec &g, GridCounts *gc) {
    for (int c = 0; c < GRID_CLASSES; c++) gc->cnt[c] = g.len[c];
    for (int cz = 0; cz < 3; cz++) {
        long long lines[3];
        for (int cy = 0; cy < 3; cy++) {
            const long long a[3] = {g.len[(cz * 3 + cy) * 3 + 0], g.len[(cz * 3 + cy) * 3 + 1], g.len[(cz * 3 + cy) * 3 + 2]};
            lines[cy] = gc->line[cz * 3 + cy] = grid_prefix(a, g.nx, g.nx);
        }
        gc->plane[cz] = grid_prefix(lines, g.ny, g.ny);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
grid_assemble_kernel(int nx, int ny, int nz, GridCounts gc, const int *__restrict__ dxyz, const T *__restrict__ cval,
                     T *__restrict__ vals, int *__restrict__ cols, int *__restrict__ rowptr) {
    const long long n = (long long)nx * ny * nz;
    const long long plane_rows = (long long)nx * ny;
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(row % nx), y = (int)((row / nx) % ny), z = (int)(row / plane_rows);
        const int cx = grid_class(x, nx), cy = grid_class(y, ny), cz = grid_class(z, nz);
        const long long lines[3] = {gc.line[cz * 3 + 0], gc.line[cz * 3 + 1], gc.line[cz * 3 + 2]};
        const int cl = (cz * 3 + cy) * 3 + cx;
        const long long in_line[3] = {gc.cnt[(cz * 3 + cy) * 3 + 0], gc.cnt[(cz * 3 + cy) * 3 + 1], gc.cnt[(cz * 3 + cy) * 3 + 2]};
        const long long start = grid_prefix(gc.plane, z, nz) + grid_prefix(lines, y, ny) + grid_prefix(in_line, x, nx);
        const int len = gc.cnt[cl];
        rowptr[row] = (int)start;
        if (row == n - 1) rowptr[n] = (int)(start + len);
        for (int e = 0; e < len; e++) {
            const int *d = dxyz + ((size_t)cl * GRID_MAXLEN + e) * 3;
            cols[start + e] = (int)(row + (long long)d[2] * plane_rows + (long long)d[1] * nx + d[0]);
            vals[start + e] = cval[(size_t)cl * GRID_MAXLEN + e];
        }
    }
}

static int grid_fill(cgb200_ctx *c, const GridSpec *g) {
    GridCounts gc;
    grid_counts(*g, &gc);
    int *d_dxyz = nullptr;
    void *d_val = nullptr;
    const size_t nd = (size_t)GRID_CLASSES * GRID_MAXLEN;
    CU(cudaMalloc(&d_dxyz, nd * 3 * sizeof(int)));
    CU(cudaMalloc(&d_val, nd * c->vsize));
    cudaError_t e = cudaMemcpyAsync(d_dxyz, g->dxyz, nd * 3 * sizeof(int), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_val, g->val, nd * c->vsize, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        const int grid = (int)std::min<long long>((long long)c->sm_count * 8, ((long long)c->n + 255) / 256);
        switch (c->dtype) {
        case CGB200_F32: grid_assemble_kernel<float><<<grid, 256, 0, c->stream>>>(g->nx, g->ny, g->nz, gc, d_dxyz, (const float *)d_val, (float *)c->d_vals, c->d_cols, c->d_rowptr); break;
        case CGB200_F64: grid_assemble_kernel<double><<<grid, 256, 0, c->stream>>>(g->nx, g->ny, g->nz, gc, d_dxyz, (const double *)d_val, (double *)c->d_vals, c->d_cols, c->d_rowptr); break;
        case CGB200_C64: grid_assemble_kernel<float2><<<grid, 256, 0, c->stream>>>(g->nx, g->ny, g->nz, gc, d_dxyz, (const float2 *)d_val, (float2 *)c->d_vals, c->d_cols, c->d_rowptr); break;
        default: grid_assemble_kernel<double2><<<grid, 256, 0, c->stream>>>(g->nx, g->ny, g->nz, gc, d_dxyz, (const double2 *)d_val, (double2 *)c->d_vals, c->d_cols, c->d_rowptr); break;
        }
        c->launches++;
        e = cudaStreamSynchronize(c->stream);
    }
    cudaFree(d_dxyz);
    cudaFree(d_val);
    if (e != cudaSuccess) return fail(CGB200_ERR_CUDA, "grid assembly: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" {

int cgb200_create_grid(cgb200_handle *out, int dtype, int device, int nx, int ny, int nz, const int *class_len,
                       const int *class_dxyz, const void *class_val) {
    if (!out) return fail(CGB200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!class_len || !class_dxyz || !class_val) return fail(CGB200_ERR_ARG, "NULL class table");
    if (nx < 1 || ny < 1 || nz < 1 || (long long)nx * ny * nz > 0x7fffffffLL) return fail(CGB200_ERR_ARG, "bad grid %d x %d x %d", nx, ny, nz);
    GridSpec g;
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.dxyz = class_dxyz;
    g.val = class_val;
    const int dims[3] = {nx, ny, nz};
    for (int cl = 0; cl < GRID_CLASSES; cl++) {
        const int cc[3] = {cl % 3, (cl / 3) % 3, cl / 9};
        g.len[cl] = class_len[cl];
        bool used = true;                       // a class exists only when every direction has such a node
        for (int a = 0; a < 3; a++) used = used && (cc[a] == 0 || (cc[a] == 2 && dims[a] >= 2) || (cc[a] == 1 && dims[a] >= 3));
        if (!used) {
            g.len[cl] = 0;
            continue;
        }
        if (class_len[cl] < 0 || class_len[cl] > GRID_MAXLEN) return fail(CGB200_ERR_ARG, "class %d has %d entries (max %d)", cl, class_len[cl], GRID_MAXLEN);
        for (int e = 0; e < class_len[cl]; e++) {
            const int *d = class_dxyz + ((size_t)cl * GRID_MAXLEN + e) * 3;
            for (int a = 0; a < 3; a++) {
                // nearest neighbours only, and the neighbour must exist: a first node has none below, a last node none above
                const int lo = cc[a] == 0 ? 0 : -1, hi = cc[a] == 2 ? 0 : 1;
                if (d[a] < lo || d[a] > hi || (dims[a] == 1 && d[a] != 0))
                    return fail(CGB200_ERR_ARG, "class %d entry %d: offset (%d,%d,%d) leaves the grid", cl, e, d[0], d[1], d[2]);
            }
            if (e > 0) {
                const int *p = d - 3;
                const bool ascending = p[2] < d[2] || (p[2] == d[2] && (p[1] < d[1] || (p[1] == d[1] && p[0] < d[0])));
                if (!ascending) return fail(CGB200_ERR_ARG, "class %d: entries must be sorted by (dz, dy, dx)", cl);
            }
        }
    }
    GridCounts gc;
    grid_counts(g, &gc);
    const long long nnz = grid_prefix(gc.plane, nz, nz);
    if (nnz > 0x7fffffffLL) return fail(CGB200_ERR_UNSUPPORTED, "nnz > 2^31-1 (int32 row offsets, as the reference)");
    return create_ctx(out, nx * ny * nz, nnz, nullptr, nullptr, nullptr, dtype, device, 0, nullptr, &g);
}

int cgb200_read_matrix(cgb200_handle c, void *aValues, int *aPointers, int *aCols) {
    if (!c) return fail(CGB200_ERR_ARG, "NULL handle");
    DeviceGuard guard(c->device);
    CU(cudaStreamSynchronize(c->stream));
    if (aValues) CU(cudaMemcpy(aValues, c->d_vals, (size_t)c->nnz * c->vsize, cudaMemcpyDefault));
    if (aPointers) CU(cudaMemcpy(aPointers, c->d_rowptr, ((size_t)c->n + 1) * sizeof(int), cudaMemcpyDefault));
    if (aCols) CU(cudaMemcpy(aCols, c->d_cols, (size_t)c->nnz * sizeof(int), cudaMemcpyDefault));
    return CGB200_OK;
}

}  // extern "C"

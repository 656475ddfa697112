// kernels.cuh -- the device side of the CG path, written for sm_100a.
//
// What the reference does with 6 kernel launches, 2 blocking reads and 2 blocking
// writes per iteration (clcg.c:296-419) is done here with three launches and no host
// round trip:
//
//   spmv_dot    q = A d  and the partial sums of d.q          (spmv.cl + vdot.cl)
//   update_xr   alpha = delta/(d.q);  x += alpha d;  r -= alpha q;  partial r.r
//                                                  (axpy.cl twice + vdot.cl + clcg.c:326)
//   update_d    beta = delta_new/delta_old;  d = r + beta d    (aypx.cl + clcg.c:390)
//
// alpha, beta, delta and the convergence state live in HBM (`CgScalars`); the
// cross-block part of every dot product is finished by the last block to retire
// (ticket counter), in a fixed order, so results are run-to-run deterministic.
//
// Vectors are stored row-major [n][k] for k right-hand sides (the k values of a row
// are contiguous), which turns the k strided gathers per non-zero of the reference's
// column-blocked layout (spmv.cl:25) into one contiguous k-wide load.
//
// All kernels are persistent grid-stride kernels: the grid is a multiple of the SM
// count and each block walks the work with stride gridDim.
#pragma once
#include "scalar.cuh"

namespace cgb {

enum : int { ST_ACTIVE = 0, ST_CONVERGED = 1, ST_BREAKDOWN = 2 };
enum : int { TK_SPMV = 0, TK_UPDATE = 1, TK_INIT = 2 };

// Device-resident scalar state of one solve (all arrays have k entries).
template <typename T> struct CgScalars {
    T *dq;              // d.q of the current iteration
    T *delta_new;       // r.r after the latest update
    T *delta_old;       // r.r before it
    double *delta0;     // |r0.r0|
    int *state;         // ST_*
    int *iters;         // iterations performed when the column stopped
    int *n_active;      // number of ST_ACTIVE columns
    int *it;            // iterations completed so far
    unsigned *ticket;   // [4] retire counters, TK_*
    T *partial;         // [grid][k] per-block partial dot products
    double *hist;       // optional delta history, hist_cap x k x (1|2) doubles
    int hist_cap;
    double tol;
};

// Matrix streams (values, column indices) are read exactly once per SpMV: mark them
// streaming so they do not displace the gathered vector from L1/L2.
template <typename T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
// Cross-block data (partials) must come from L2, never from a stale L1 line.
template <typename T> __device__ __forceinline__ T ld_cg(const T *p) { return __ldcg(p); }

// ---------------------------------------------------------------------------
// column-aware reductions
// ---------------------------------------------------------------------------
// Every thread holds V partial sums, for the V columns of column pack
// cp = threadIdx.x % group.  blockDim.x = group * 2^m.  On return smem[t*V + v],
// t < group, holds the block's sum for column t*V + v.
template <typename T, int V>
__device__ __forceinline__ void block_col_reduce(const T (&acc)[V], int group, T *smem) {
    const int t = threadIdx.x;
#pragma unroll
    for (int v = 0; v < V; v++) smem[t * V + v] = acc[v];
    __syncthreads();
    for (int off = blockDim.x >> 1; off >= group; off >>= 1) {
        if (t < off) {
#pragma unroll
            for (int v = 0; v < V; v++)
                smem[t * V + v] = Sc<T>::add(smem[t * V + v], smem[(t + off) * V + v]);
        }
        __syncthreads();
    }
}

// Publishes this block's column sums (smem[0 .. kv*V)) to partial[blockIdx][.] and
// returns true in exactly one block: the last one to arrive at `ticket`.
template <typename T, int V>
__device__ __forceinline__ bool publish_and_arrive(const T *smem, int kv, int k, T *partial,
                                                   unsigned *ticket) {
    __shared__ int s_last;
    const int t = threadIdx.x;
    if (t < kv) {
#pragma unroll
        for (int v = 0; v < V; v++)
            if (t * V + v < k) partial[(size_t)blockIdx.x * k + t * V + v] = smem[t * V + v];
        __threadfence();
    }
    __syncthreads();
    if (t == 0) {
        const unsigned prev = atomicAdd(ticket, 1u);
        s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    return s_last != 0;
}

// Run by the last block only: sums partial[b][.] over the blocks b in a fixed order.
// Thread t owns column pack t % group (valid when < kv) and blocks t/group, +blockDim/group, ...
// On return smem[t*V + v], t < kv, holds the grid-wide sum for column t*V + v.
template <typename T, int V>
__device__ __forceinline__ void grid_col_reduce(const T *partial, int group, int kv, int k, T *smem) {
    __threadfence();
    const int t = threadIdx.x;
    const int cp = t % group;
    T acc[V];
#pragma unroll
    for (int v = 0; v < V; v++) acc[v] = Sc<T>::zero();
    if (cp < kv) {
        for (int b = t / group; b < (int)gridDim.x; b += blockDim.x / group) {
#pragma unroll
            for (int v = 0; v < V; v++)
                if (cp * V + v < k)
                    acc[v] = Sc<T>::add(acc[v], ld_cg(partial + (size_t)b * k + cp * V + v));
        }
    }
    __syncthreads();
    block_col_reduce<T, V>(acc, group, smem);
}

// ---------------------------------------------------------------------------
// SpMV, one right-hand side: LPR lanes cooperate on a row (CSR-vector with the lane
// count matched to the row-length distribution), optionally fused with the partial
// sums of x.y (d.q in the CG loop).   spmv.cl:13-49 (+ vdot.cl)
// ---------------------------------------------------------------------------
template <typename T, int LPR, bool DOT>
__global__ void __launch_bounds__(256)
spmv1_kernel(int n, const T *__restrict__ vals, const int *__restrict__ rowptr,
             const int *__restrict__ cols, const T *__restrict__ x, T *__restrict__ y,
             CgScalars<T> sc) {
    if (DOT) {
        if (*sc.n_active == 0) return;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    const int t = threadIdx.x;
    const int lane = t % LPR;
    const int rows_per_block = blockDim.x / LPR;
    T dot[1] = {Sc<T>::zero()};

    for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < n;
         row0 += (long long)gridDim.x * rows_per_block) {
        const int row = (int)row0 + t / LPR;
        T sum = Sc<T>::zero();
        if (row < n) {
            const int lo = __ldg(rowptr + row), hi = __ldg(rowptr + row + 1);
            for (int j = lo + lane; j < hi; j += LPR) {
                const T a = ld_stream(vals + j);
                const int c = ld_stream(cols + j);
                sum = Sc<T>::fma(a, __ldg(x + c), sum);
            }
        }
        // LPR-lane segmented butterfly; all 32 lanes of the warp take part
#pragma unroll
        for (int off = LPR >> 1; off > 0; off >>= 1) {
            if constexpr (Sc<T>::cplx) {
                sum.x += __shfl_xor_sync(0xffffffffu, sum.x, off);
                sum.y += __shfl_xor_sync(0xffffffffu, sum.y, off);
            } else {
                sum += __shfl_xor_sync(0xffffffffu, sum, off);
            }
        }
        if (lane == 0 && row < n) {
            y[row] = sum;
            if (DOT) dot[0] = Sc<T>::fma(__ldg(x + row), sum, dot[0]);
        }
    }

    if (DOT) {
        block_col_reduce<T, 1>(dot, 1, smem);
        if (publish_and_arrive<T, 1>(smem, 1, 1, sc.partial, sc.ticket + TK_SPMV)) {
            grid_col_reduce<T, 1>(sc.partial, 1, 1, 1, smem);
            if (t == 0) {
                sc.dq[0] = smem[0];
                sc.ticket[TK_SPMV] = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// SpMM, k right-hand sides in row-major [n][k]: G lanes per row, lane cp owns the
// V-wide column pack cp (one 128-bit gather per non-zero per lane when
// V*sizeof(T) = 16).  kv = k / V <= G; G is a power of two <= 32.
// ---------------------------------------------------------------------------
template <typename T, int V, int G, bool DOT>
__global__ void __launch_bounds__(256)
spmm_kernel(int n, int k, const T *__restrict__ vals, const int *__restrict__ rowptr,
            const int *__restrict__ cols, const T *__restrict__ x, T *__restrict__ y,
            CgScalars<T> sc) {
    if (DOT) {
        if (*sc.n_active == 0) return;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    const int cp = t % G;
    const int kv = k / V;
    const bool active = cp < kv;
    const int rows_per_block = blockDim.x / G;
    T dot[V];
#pragma unroll
    for (int v = 0; v < V; v++) dot[v] = Sc<T>::zero();

    for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < n;
         row0 += (long long)gridDim.x * rows_per_block) {
        const int row = (int)row0 + t / G;
        if (row < n && active) {
            const int lo = __ldg(rowptr + row), hi = __ldg(rowptr + row + 1);
            T acc[V];
#pragma unroll
            for (int v = 0; v < V; v++) acc[v] = Sc<T>::zero();
#pragma unroll 4
            for (int j = lo; j < hi; j++) {
                const T a = __ldg(vals + j);   // one address for the G lanes of the row: a broadcast
                const int c = __ldg(cols + j);
                const P xv = *reinterpret_cast<const P *>(x + (size_t)c * k + (size_t)cp * V);
#pragma unroll
                for (int v = 0; v < V; v++) acc[v] = Sc<T>::fma(a, xv.v[v], acc[v]);
            }
            P out;
#pragma unroll
            for (int v = 0; v < V; v++) out.v[v] = acc[v];
            *reinterpret_cast<P *>(y + (size_t)row * k + (size_t)cp * V) = out;
            if (DOT) {
                const P xo = *reinterpret_cast<const P *>(x + (size_t)row * k + (size_t)cp * V);
#pragma unroll
                for (int v = 0; v < V; v++) dot[v] = Sc<T>::fma(xo.v[v], acc[v], dot[v]);
            }
        }
    }

    if (DOT) {
        block_col_reduce<T, V>(dot, G, smem);
        if (publish_and_arrive<T, V>(smem, kv, k, sc.partial, sc.ticket + TK_SPMV)) {
            grid_col_reduce<T, V>(sc.partial, G, kv, k, smem);
            if (t < kv) {
#pragma unroll
                for (int v = 0; v < V; v++) sc.dq[t * V + v] = smem[t * V + v];
            }
            if (t == 0) sc.ticket[TK_SPMV] = 0;
        }
    }
}

// ---------------------------------------------------------------------------
// Vector kernels.  The [n][k] array is walked as packs of V values; thread g of the
// grid owns packs g, g + stride, ... with stride = gridDim*blockDim a multiple of
// kv = k/V, so a thread always sees the same V columns: column(v) = (t % kv)*V + v.
// k == 1 (any V): every value belongs to column 0.
// blockDim.x = kv * 2^m.
// ---------------------------------------------------------------------------
template <int V> __device__ __forceinline__ int col_of(int k, int kv, int v) {
    return (k == 1) ? 0 : (int)(threadIdx.x % kv) * V + v;
}

// r = b - q ; d = r ; delta = r.r          clcg.c:259-292 (sub.cl, copy, vdot.cl)
template <typename T, int V>
__global__ void __launch_bounds__(256)
init_kernel(size_t npacks, size_t nelem, int k, int kv, const T *b /* may alias d */,
            const T *__restrict__ q, T *__restrict__ r, T *d, CgScalars<T> sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    T acc[V];
#pragma unroll
    for (int v = 0; v < V; v++) acc[v] = Sc<T>::zero();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + t; p < npacks; p += stride) {
        const P bv = reinterpret_cast<const P *>(b)[p];
        const P qv = reinterpret_cast<const P *>(q)[p];
        P rv;
#pragma unroll
        for (int v = 0; v < V; v++) {
            rv.v[v] = Sc<T>::sub(bv.v[v], qv.v[v]);
            acc[v] = Sc<T>::fma(rv.v[v], rv.v[v], acc[v]);
        }
        reinterpret_cast<P *>(r)[p] = rv;
        reinterpret_cast<P *>(d)[p] = rv;
    }
    if (V > 1 && blockIdx.x == 0) {   // k == 1 tail that does not fill a pack
        const size_t e = npacks * V + t;
        if (e < nelem) {
            const T rv = Sc<T>::sub(b[e], q[e]);
            r[e] = rv;
            d[e] = rv;
            acc[0] = Sc<T>::fma(rv, rv, acc[0]);
        }
    }
    if (k == 1) {
#pragma unroll
        for (int v = 1; v < V; v++) { acc[0] = Sc<T>::add(acc[0], acc[v]); acc[v] = Sc<T>::zero(); }
    }
    block_col_reduce<T, V>(acc, kv, smem);
    if (publish_and_arrive<T, V>(smem, kv, k, sc.partial, sc.ticket + TK_INIT)) {
        grid_col_reduce<T, V>(sc.partial, kv, kv, k, smem);
        if (t < kv) {
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int c = t * V + v;
                if (c < k) {
                    const T dl = smem[t * V + v];
                    const double a0 = Sc<T>::abs(dl);
                    sc.delta_new[c] = dl;
                    sc.delta_old[c] = dl;
                    sc.dq[c] = Sc<T>::zero();
                    sc.delta0[c] = a0;
                    const bool live = (a0 > 0.0) && Sc<T>::finite(dl);
                    sc.state[c] = live ? ST_ACTIVE : (a0 == 0.0 ? ST_CONVERGED : ST_BREAKDOWN);
                    sc.iters[c] = 0;
                    if (sc.hist && sc.hist_cap > 0)
                        Sc<T>::to_double2(dl, sc.hist + (size_t)c * (Sc<T>::cplx ? 2 : 1));
                }
            }
        }
        __syncthreads();
        if (t == 0) {
            int live = 0;
            for (int c = 0; c < k; c++) live += (sc.state[c] == ST_ACTIVE);
            *sc.n_active = live;
            *sc.it = 0;
            sc.ticket[TK_INIT] = 0;
        }
    }
}

// alpha = delta_new / dq ; x += alpha d ; r -= alpha q ; delta_old = delta_new ;
// delta_new = r.r ; convergence bookkeeping.       clcg.c:326-392
template <typename T, int V>
__global__ void __launch_bounds__(256)
update_xr_kernel(size_t npacks, size_t nelem, int k, int kv, const T *__restrict__ d,
                 const T *__restrict__ q, T *__restrict__ x, T *__restrict__ r, CgScalars<T> sc) {
    if (*sc.n_active == 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    T alpha[V], acc[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
        const int c = col_of<V>(k, kv, v);
        acc[v] = Sc<T>::zero();
        alpha[v] = Sc<T>::zero();
        if (c < k && sc.state[c] == ST_ACTIVE) {
            const T den = sc.dq[c];
            if (!Sc<T>::is_zero(den)) alpha[v] = Sc<T>::div(sc.delta_new[c], den);
        }
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + t; p < npacks; p += stride) {
        const P dv = reinterpret_cast<const P *>(d)[p];
        const P qv = reinterpret_cast<const P *>(q)[p];
        P xv = reinterpret_cast<const P *>(x)[p];
        P rv = reinterpret_cast<const P *>(r)[p];
#pragma unroll
        for (int v = 0; v < V; v++) {
            xv.v[v] = Sc<T>::fma(alpha[v], dv.v[v], xv.v[v]);
            rv.v[v] = Sc<T>::fnma(alpha[v], qv.v[v], rv.v[v]);
            acc[v] = Sc<T>::fma(rv.v[v], rv.v[v], acc[v]);
        }
        reinterpret_cast<P *>(x)[p] = xv;
        reinterpret_cast<P *>(r)[p] = rv;
    }
    if (V > 1 && blockIdx.x == 0) {
        const size_t e = npacks * V + t;
        if (e < nelem) {
            x[e] = Sc<T>::fma(alpha[0], d[e], x[e]);
            const T rv = Sc<T>::fnma(alpha[0], q[e], r[e]);
            r[e] = rv;
            acc[0] = Sc<T>::fma(rv, rv, acc[0]);
        }
    }
    if (k == 1) {
#pragma unroll
        for (int v = 1; v < V; v++) { acc[0] = Sc<T>::add(acc[0], acc[v]); acc[v] = Sc<T>::zero(); }
    }
    block_col_reduce<T, V>(acc, kv, smem);
    if (publish_and_arrive<T, V>(smem, kv, k, sc.partial, sc.ticket + TK_UPDATE)) {
        grid_col_reduce<T, V>(sc.partial, kv, kv, k, smem);
        const int it1 = *sc.it + 1;
        if (t < kv) {
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int c = t * V + v;
                if (c < k) {
                    if (sc.state[c] == ST_ACTIVE) {
                        const T nd = smem[t * V + v];
                        sc.delta_old[c] = sc.delta_new[c];
                        sc.delta_new[c] = nd;
                        const double a = Sc<T>::abs(nd);
                        int st = ST_ACTIVE;
                        if (!Sc<T>::finite(nd)) st = ST_BREAKDOWN;
                        else if (a == 0.0 || (sc.tol > 0.0 && sqrt(a / sc.delta0[c]) < sc.tol)) st = ST_CONVERGED;
                        if (st != ST_ACTIVE) {
                            sc.state[c] = st;
                            sc.iters[c] = it1;
                            atomicSub(sc.n_active, 1);
                        }
                    }
                    if (sc.hist && it1 < sc.hist_cap)
                        Sc<T>::to_double2(sc.delta_new[c],
                                          sc.hist + ((size_t)it1 * k + c) * (Sc<T>::cplx ? 2 : 1));
                }
            }
        }
        __syncthreads();
        if (t == 0) {
            *sc.it = it1;
            sc.ticket[TK_UPDATE] = 0;
        }
    }
}

// beta = delta_new / delta_old ; d = beta d + r        clcg.c:389-415 (aypx.cl)
template <typename T, int V>
__global__ void __launch_bounds__(256)
update_d_kernel(size_t npacks, size_t nelem, int k, int kv, const T *__restrict__ r,
                T *__restrict__ d, CgScalars<T> sc) {
    if (*sc.n_active == 0) return;
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    T beta[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
        const int c = col_of<V>(k, kv, v);
        beta[v] = Sc<T>::zero();
        if (c < k && sc.state[c] == ST_ACTIVE) {
            const T den = sc.delta_old[c];
            if (!Sc<T>::is_zero(den)) beta[v] = Sc<T>::div(sc.delta_new[c], den);
        }
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + t; p < npacks; p += stride) {
        const P rv = reinterpret_cast<const P *>(r)[p];
        P dv = reinterpret_cast<const P *>(d)[p];
#pragma unroll
        for (int v = 0; v < V; v++) dv.v[v] = Sc<T>::fma(beta[v], dv.v[v], rv.v[v]);
        reinterpret_cast<P *>(d)[p] = dv;
    }
    if (V > 1 && blockIdx.x == 0) {
        const size_t e = npacks * V + t;
        if (e < nelem) d[e] = Sc<T>::fma(beta[0], d[e], r[e]);
    }
}

// ---------------------------------------------------------------------------
// Layout change between the cg() ABI ([k][n], RHS r at r*n) and the engine's
// row-major [n][k].  src is [rows][cols], dst is [cols][rows].
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T *__restrict__ src, T *__restrict__ dst, int rows, long long cols) {
    __shared__ T tile[32][33];
    const long long c0 = (long long)blockIdx.x * 32;
    const int r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int rr = r0 + i;
        const long long cc = c0 + threadIdx.x;
        if (rr < rows && cc < cols) tile[i][threadIdx.x] = src[(size_t)rr * cols + cc];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long cc = c0 + i;
        const int rr = r0 + threadIdx.x;
        if (rr < rows && cc < cols) dst[(size_t)cc * rows + rr] = tile[threadIdx.x][i];
    }
}

}  // namespace cgb

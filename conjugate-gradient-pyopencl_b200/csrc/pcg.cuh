// pcg.cuh -- Jacobi-preconditioned CG on the device: the reference's PCG (helmFE_var.py:546-586) with
// M = an inverse diagonal (what the reference applies as a sparse matrix with one entry per row, :559-563).
//
//   z = M r ; rho = r.z ; p = z + (rho/rho_prev) p ; q = A p ; alpha = rho / p.q ; x += alpha p ; r -= alpha q ;
//   stop when sqrt|r.r| < tol (an ABSOLUTE tolerance, :579-583)
//
// arranged like the engine's CG loop (the arrangement the tests' C checker of this path follows): the SpMV fused with
// p.q is the CG kernel unchanged (CSR or row-pattern dictionary); the x/r update also forms z = dinv*r on the
// fly and reduces BOTH r.z and r.r in the same pass; the direction update recomputes z instead of reading a
// stored copy.  z is never written: the preconditioner costs one extra read of dinv in each vector kernel
// (8 instead of 6 passes, 4 instead of 3).  One right-hand side per launch; unconjugated dot products.
#pragma once

namespace cgb {

// block sums of two values in a fixed order (lanes by butterfly, then the warps one after the other);
// the result is valid in thread 0
template <typename T>
__device__ __forceinline__ void pcg_block_sum2(T &a, T &b, T *smem /* [2 * 32] */) {
    const int t = threadIdx.x;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        if constexpr (Sc<T>::cplx) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
            b.x += __shfl_xor_sync(0xffffffffu, b.x, off);
            b.y += __shfl_xor_sync(0xffffffffu, b.y, off);
        } else {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
    }
    __syncthreads();            // smem may still be read from a previous use
    if ((t & 31) == 0) {
        smem[2 * (t >> 5)] = a;
        smem[2 * (t >> 5) + 1] = b;
    }
    __syncthreads();
    if (t == 0) {
        T sa = Sc<T>::zero(), sb = Sc<T>::zero();
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; w++) {
            sa = Sc<T>::add(sa, smem[2 * w]);
            sb = Sc<T>::add(sb, smem[2 * w + 1]);
        }
        a = sa;
        b = sb;
    }
}

// Publishes the block's two sums; in the last block to arrive, thread 0 returns with the grid totals in (a, b)
// and true.  partial: [grid][2].
template <typename T>
__device__ __forceinline__ bool pcg_grid_sum2(T &a, T &b, T *partial, unsigned *ticket, T *smem) {
    __shared__ int s_last;
    const int t = threadIdx.x;
    pcg_block_sum2<T>(a, b, smem);
    if (t == 0) {
        partial[2 * (size_t)blockIdx.x] = a;
        partial[2 * (size_t)blockIdx.x + 1] = b;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    T sa = Sc<T>::zero(), sb = Sc<T>::zero();
    for (int blk = t; blk < (int)gridDim.x; blk += blockDim.x) {       // strided, ascending: a fixed order
        sa = Sc<T>::add(sa, ld_cg(partial + 2 * (size_t)blk));
        sb = Sc<T>::add(sb, ld_cg(partial + 2 * (size_t)blk + 1));
    }
    pcg_block_sum2<T>(sa, sb, smem);
    a = sa;
    b = sb;
    return t == 0;
}

// After rho = r.z and rr = r.r are known: convergence on sqrt|r.r| < tol (absolute), history of r.r.
template <typename T>
__device__ __forceinline__ void pcg_bookkeep(const CgScalars<T> &sc, int it1, T rho, T rr, bool init) {
    const double a = Sc<T>::abs(rr);
    if (init) {
        sc.delta_new[0] = rho;
        sc.delta_old[0] = rho;
        sc.dq[0] = Sc<T>::zero();
        sc.delta0[0] = a;
        sc.iters[0] = 0;
        const bool live = a > 0.0 && Sc<T>::finite(rr) && Sc<T>::finite(rho) && !(*sc.tol > 0.0 && sqrt(a) < *sc.tol);
        sc.state[0] = live ? ST_ACTIVE : ((a == 0.0 || (Sc<T>::finite(rr) && Sc<T>::finite(rho))) ? ST_CONVERGED : ST_BREAKDOWN);
        *sc.n_active = live ? 1 : 0;
        *sc.it = 0;
    } else if (sc.state[0] == ST_ACTIVE) {
        sc.delta_old[0] = sc.delta_new[0];
        sc.delta_new[0] = rho;
        int st = ST_ACTIVE;
        if (!Sc<T>::finite(rr) || !Sc<T>::finite(rho)) st = ST_BREAKDOWN;
        else if (a == 0.0 || (*sc.tol > 0.0 && sqrt(a) < *sc.tol)) st = ST_CONVERGED;
        if (st != ST_ACTIVE) {
            sc.state[0] = st;
            sc.iters[0] = it1;
            *sc.n_active = 0;
        }
        *sc.it = it1;
    }
    sc.rr[0] = rr;
    if (sc.hist && it1 < sc.hist_cap) Sc<T>::to_double2(rr, sc.hist + (size_t)it1 * (Sc<T>::cplx ? 2 : 1));
}

// r = b - q ; d = dinv*r ; rho = r.(dinv*r) ; rr = r.r          helmFE_var.py:556-572 (first pass of the loop)
template <typename T>
__global__ void __launch_bounds__(256)
pcg_init_kernel(size_t n, const T *b /* may alias d */, const T *__restrict__ q, const T *__restrict__ dinv,
                T *__restrict__ r, T *d, CgScalars<T> sc) {
    __shared__ T smem[64];
    T rho = Sc<T>::zero(), rr = Sc<T>::zero();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const T rv = Sc<T>::sub(b[i], q[i]);
        const T z = Sc<T>::mul(dinv[i], rv);
        r[i] = rv;
        d[i] = z;
        rho = Sc<T>::fma(rv, z, rho);
        rr = Sc<T>::fma(rv, rv, rr);
    }
    if (pcg_grid_sum2<T>(rho, rr, sc.partial, sc.ticket + TK_INIT, smem)) {
        pcg_bookkeep<T>(sc, 0, rho, rr, true);
        sc.ticket[TK_INIT] = 0;
    }
}

// alpha = rho / p.q ; x += alpha p ; r -= alpha q ; rho_new = r.(dinv*r) ; rr = r.r      helmFE_var.py:576-583
template <typename T>
__global__ void __launch_bounds__(256)
pcg_update_xr_kernel(size_t n, const T *__restrict__ d, const T *__restrict__ q, const T *__restrict__ dinv,
                     T *__restrict__ x, T *__restrict__ r, CgScalars<T> sc) {
    if (*sc.n_active == 0) return;
    __shared__ T smem[64];
    T alpha = Sc<T>::zero();
    {
        const T den = sc.dq[0];
        if (!Sc<T>::is_zero(den)) alpha = Sc<T>::div(sc.delta_new[0], den);
    }
    T rho = Sc<T>::zero(), rr = Sc<T>::zero();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        st_stream_bytes(x + i, Sc<T>::fma(alpha, d[i], ld_stream_bytes(x + i)));
        const T rv = Sc<T>::fnma(alpha, q[i], r[i]);
        r[i] = rv;
        const T z = Sc<T>::mul(dinv[i], rv);
        rho = Sc<T>::fma(rv, z, rho);
        rr = Sc<T>::fma(rv, rv, rr);
    }
    if (pcg_grid_sum2<T>(rho, rr, sc.partial, sc.ticket + TK_UPDATE, smem)) {
        pcg_bookkeep<T>(sc, *sc.it + 1, rho, rr, false);
        sc.ticket[TK_UPDATE] = 0;
    }
}

// beta = rho_new / rho_old ; p = dinv*r + beta p          helmFE_var.py:573-574
template <typename T>
__global__ void __launch_bounds__(256)
pcg_update_d_kernel(size_t n, const T *__restrict__ r, const T *__restrict__ dinv, T *__restrict__ d, CgScalars<T> sc) {
    if (*sc.n_active == 0) return;
    T beta = Sc<T>::zero();
    {
        const T den = sc.delta_old[0];
        if (!Sc<T>::is_zero(den)) beta = Sc<T>::div(sc.delta_new[0], den);
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = Sc<T>::fma(beta, d[i], Sc<T>::mul(dinv[i], r[i]));
}

// dinv[i] = 1 / A[i][i] (0 where the row has no diagonal entry): the Jacobi preconditioner of the resident matrix
template <typename T>
__global__ void __launch_bounds__(256)
jacobi_dinv_kernel(int n, const T *__restrict__ vals, const int *__restrict__ rowptr, const int *__restrict__ cols,
                   T *__restrict__ dinv) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
        T v = Sc<T>::zero();
        for (int j = rowptr[row]; j < rowptr[row + 1]; j++)
            if (cols[j] == row) v = vals[j];
        T one;
        if constexpr (Sc<T>::cplx) one = Sc<T>::make(1, 0);
        else one = (T)1;
        dinv[row] = Sc<T>::is_zero(v) ? Sc<T>::zero() : Sc<T>::div(one, v);
    }
}

}  // namespace cgb

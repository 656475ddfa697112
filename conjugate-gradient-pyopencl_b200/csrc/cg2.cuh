// cg2.cuh -- the CG iteration in TWO kernels and nine vector passes, for one right-hand side and a matrix
// that has a row-pattern dictionary (kernels.cuh, spmv_pattern_kernel): the stencil / finite-element operators
// of BASELINE configs 1, 2 and 4 and of the reference's as_prec subdomain solves.
//
// The three-kernel iteration (spmv_dot, update_xr, update_d) makes 11 passes over n-vectors per iteration
// (d, q | d, q, x, r, x, r | r, d, d).  Here:
//
//   dir_spmv   beta = delta_new/delta_old (from the scalars);  dn = r + beta d  is NOT a pass of its own
//              (aypx.cl, clcg.c:411-415): the gather of the SpMV forms r[j] + beta d[j] on the fly
//              (spmv.cl:13-49), the row's owner stores dn into the OTHER direction buffer, the pending
//              x += alpha_prev d (axpy.cl, clcg.c:331-337) rides along because d is in flight anyway, and
//              the partial sums of dn.q (vdot.cl) are fused in.    reads r d x, writes q dn x  (6 passes)
//   update_r   alpha = delta_new/(d.q);  r' = r - alpha q  into the OTHER residual buffer;  partial r'.r'
//              (axpy.cl + vdot.cl, clcg.c:326-386); bookkeeping (delta shuffle, beta, convergence,
//              clcg.c:350-356,389-391) by the last block.         reads q r, writes r'       (3 passes)
//
// x lags one update behind: x_k = x_{k-1} + alpha_{k-1} d_{k-1} is applied by dir_spmv of iteration k, and
// finish_x_kernel applies the last one.  All values are produced by the same FMAs in the same order as in the
// three-kernel path (the row sums follow CSR order, dn = fma(beta, d, r), x = fma(alpha, d, x),
// r = fma(-alpha, q, r)); only the association of the two dot-product reductions differs.
//
// The gather.  A block works on chunks of PAT_CHUNK consecutive rows.  The pattern dictionary knows every
// column offset (col - row) the matrix uses -- 7 for the 3-D Laplacian --; offsets that are closer than a
// chunk are merged into a WINDOW, and for a chunk starting at row c0 the window [lo, hi] needs exactly the
// vector entries [c0 + lo, c0 + hi + PAT_CHUNK).  Those are read with coalesced 128-bit loads from both r and
// d, combined (r + beta d) and left in shared memory once; the rows then read their neighbours from shared
// memory at consecutive addresses (conflict-free).  Compared with 7 scattered 8-byte gathers per row and
// vector this is 3.6 contiguous elements per row and vector for the 300^3 Laplacian, and the L1 pipe sees
// 128-bit requests instead of 64-bit ones.
//
// Row-block shards.  Ping-pong buffers for d (the SpMV reads the old one everywhere while owners write the
// new one) and for r make every kernel read-only on its inputs, so the entries a peer needs can be
// recomputed and stored straight into the peer's halo (NVLink) by a prologue of the SAME kernel that
// produces them: dir_spmv pushes dn, update_r pushes r'.  No push kernel, no side stream, no flags: the
// all-reduce at the tail of every kernel already orders "all my blocks are done" before "peers proceed",
// and the next kernel that reads a halo is at least one all-reduce later.
#pragma once

namespace cgb {

constexpr int WIN_MAX = 8;
struct PatWindows {
    int nwin;
    int total;             // staged elements per chunk, all windows
    int diag;              // staging position of column offset 0 for the chunk's first row
    int lo[WIN_MAX];       // first staged column of the window relative to the chunk's first row (multiple of the pack width)
    int size[WIN_MAX];     // staged elements (multiple of the pack width)
    int base[WIN_MAX];     // position of the window in the staging array (ascending, multiple of the pack width)
};

// OR over the rows of each chunk of the windows their patterns read (row-block shards: the windows of the
// halo offsets are only staged by the chunks that hold boundary rows).
__global__ void __launch_bounds__(256)
chunk_window_mask_kernel(int n, int nchunks, const unsigned short *__restrict__ pat, const unsigned *__restrict__ pat_mask,
                         unsigned *__restrict__ chunk_mask) {
    __shared__ unsigned s_or[256 / 32];
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        unsigned m = 0;
        for (int r = ch * PAT_CHUNK + threadIdx.x; r < min(n, (ch + 1) * PAT_CHUNK); r += 256) m |= pat_mask[pat[r]];
        m = __reduce_or_sync(0xffffffffu, m);
        if ((threadIdx.x & 31) == 0) s_or[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned all = 0;
            for (int w = 0; w < 256 / 32; w++) all |= s_or[w];
            chunk_mask[ch] = all;
        }
        __syncthreads();
    }
}

// Sums the per-block partial dot products (last block only), all-reduces over the GPUs, returns the total to
// every thread of that block.  `red` has blockDim entries.
template <typename T, bool PEER>
__device__ __forceinline__ T cg2_grid_total(const CgScalars<T> &sc, T *red) {
    grid_col_reduce<T, 1>(sc.partial, 1, 1, 1, red);
    T total = red[0];
    if constexpr (PEER) {
        if (sc.peer) total = peer_allreduce<T>(sc.peer, red[0]);
    }
    return total;
}

template <typename T, int STRIDE, bool PEER>
__global__ void __launch_bounds__(PAT_THREADS, sizeof(T) == 16 ? 2 : 3)
cg2_dir_spmv_kernel(int n, int ncols, int nchunks, int npat, PatWindows win, const unsigned short *__restrict__ pat,
                    const unsigned *__restrict__ chunk_mask, const int *__restrict__ p_len, const int *__restrict__ p_spos,
                    const T *__restrict__ p_val, T *__restrict__ x, T *__restrict__ q, T *r0, T *r1, T *d0, T *d1,
                    CgScalars<T> sc) {
    constexpr int NT = PAT_THREADS;
    constexpr int VPT = VecW<T>::value;
    using P = Pack<T, VPT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *red = reinterpret_cast<T *>(smem_raw);                                  // [NT]
    T *s_val = red + NT;                                                       // [npat][STRIDE]
    T *S = s_val + npat * STRIDE;                                              // [win.total]   r + beta d around the chunk
    int *s_pos = reinterpret_cast<int *>(S + win.total);                       // [npat][STRIDE] staging position of every entry
    int *s_len = s_pos + npat * STRIDE;                                        // [npat]
    const int t = threadIdx.x;
    // the table is part of the matrix, not the previous kernel's output: staged before the grid dependency wait
    for (int i = t; i < npat * STRIDE; i += NT) {
        const int id = i / STRIDE, j = i % STRIDE;
        s_val[i] = p_val[id * PAT_MAXLEN + j];
        s_pos[i] = p_spos[id * PAT_MAXLEN + j];
    }
    for (int i = t; i < npat; i += NT) s_len[i] = p_len[i];
    __syncthreads();
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (*sc.n_active == 0) return;
    const int it = *sc.it;
    if (sc.trace && blockIdx.x == 0 && t == 0) trace_mark<T>(sc, it, TR_SPMV_START);
    const bool odd = (it & 1) != 0;
    const T *__restrict__ r = odd ? r1 : r0;
    const T *__restrict__ dold = odd ? d1 : d0;
    T *__restrict__ dnew = odd ? d0 : d1;
    const T beta = sc.beta[0], alpha_prev = sc.alpha[0];

    if constexpr (PEER) {
        if (sc.peer && sc.peer->world > 1)
            peer_push_rows<T>(sc.peer, odd ? 0 : 1, [&](int row) { return Sc<T>::fma(beta, dold[row], r[row]); });
    }

    T dot[1] = {Sc<T>::zero()};
    const int npk = win.total / VPT;
    constexpr int ROWS_PT = PAT_CHUNK / NT;

    for (int ch = (int)blockIdx.x; ch < nchunks; ch += (int)gridDim.x) {
        const int c0 = ch * PAT_CHUNK;
        const unsigned wmask = chunk_mask[ch];
        int ids[ROWS_PT];
#pragma unroll
        for (int s = 0; s < ROWS_PT; s++) {
            const int row = c0 + t + s * NT;
            ids[s] = row < n ? (int)pat[row] : -1;
        }
        // ---- stage r + beta d of every window: packs t, t + NT, ..., four loads of each vector in flight per thread
        for (int e0 = t; e0 < npk; e0 += 4 * NT) {
            P rv[4], dv[4];
            int spos[4];
            bool ld[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * NT;
                const int pos = e * VPT;
                int w = 0;
#pragma unroll
                for (int i = 1; i < WIN_MAX; i++) w += (i < win.nwin && pos >= win.base[i]) ? 1 : 0;
                const long long g = (long long)c0 + win.lo[w] + (pos - win.base[w]);
                spos[u] = e < npk ? pos : -1;
                ld[u] = e < npk && ((wmask >> w) & 1u) && g >= 0 && g < ncols;
                if (ld[u]) {
                    rv[u] = *reinterpret_cast<const P *>(r + g);
                    dv[u] = *reinterpret_cast<const P *>(dold + g);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (spos[u] >= 0) {
                    P wv;
#pragma unroll
                    for (int v = 0; v < VPT; v++) wv.v[v] = ld[u] ? Sc<T>::fma(beta, dv[u].v[v], rv[u].v[v]) : Sc<T>::zero();
                    *reinterpret_cast<P *>(S + spos[u]) = wv;
                }
            }
        }
        __syncthreads();
        // ---- rows: neighbours out of shared memory at consecutive addresses.
        // STRIDE 8: the pattern of a thread's previous row of the chunk stays in registers (on a grid nearly
        // every row has the same one); longer rows read the table out of shared memory entry by entry
        int cid = -1;
        int cpos[STRIDE <= 8 ? STRIDE : 1];
        T cval[STRIDE <= 8 ? STRIDE : 1];
#pragma unroll
        for (int s = 0; s < ROWS_PT; s++) {
            const int tl = t + s * NT, row = c0 + tl;
            const int id = ids[s];
            if (id < 0) break;
            const T xo = ld_stream_bytes(x + row);
            const T dd = dold[row];
            T sum = Sc<T>::zero();
            if constexpr (STRIDE <= 8) {
                if (id != cid) {
                    cid = id;
#pragma unroll
                    for (int j = 0; j < STRIDE; j++) {
                        cpos[j] = s_pos[id * STRIDE + j];
                        cval[j] = s_val[id * STRIDE + j];
                    }
                }
                // padded entries point at the row's own position with coefficient 0 (as spmv_pattern_kernel pads)
#pragma unroll
                for (int j = 0; j < STRIDE; j++) sum = Sc<T>::fma(cval[j], S[cpos[j] + tl], sum);
            } else {
                const int len = s_len[id];
                for (int j = 0; j < len; j++) sum = Sc<T>::fma(s_val[id * STRIDE + j], S[s_pos[id * STRIDE + j] + tl], sum);
            }
            const T dn = S[win.diag + tl];
            q[row] = sum;
            dnew[row] = dn;
            st_stream_bytes(x + row, Sc<T>::fma(alpha_prev, dd, xo));
            dot[0] = Sc<T>::fma(dn, sum, dot[0]);
        }
        __syncthreads();        // S is rewritten by the next chunk
    }

    block_col_reduce<T, 1>(dot, 1, red);
    if (publish_and_arrive<T, 1>(red, 1, 1, sc.partial, sc.ticket + TK_SPMV)) {
        if (t == 0 && sc.trace) trace_mark<T>(sc, it, TR_SPMV_ALL_DONE);
        const T total = cg2_grid_total<T, PEER>(sc, red);
        if (t == 0) {
            sc.dq[0] = total;
            sc.ticket[TK_SPMV] = 0;
            if (sc.trace) trace_mark<T>(sc, it, TR_SPMV_END);
        }
    }
}

// alpha = delta_new / dq ; r' = r - alpha q (into the other residual buffer) ; delta_old = delta_new ;
// delta_new = r'.r' ; beta ; convergence bookkeeping.      clcg.c:326-392
template <typename T, int V, bool PEER>
__global__ void __launch_bounds__(256)
cg2_update_r_kernel(size_t npacks, size_t nelem, const T *__restrict__ q, T *r0, T *r1, CgScalars<T> sc) {
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (*sc.n_active == 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    const int it = *sc.it;
    if (sc.trace && blockIdx.x == 0 && t == 0) trace_mark<T>(sc, it, TR_XR_START);
    const bool odd = (it & 1) != 0;
    const T *__restrict__ r = odd ? r1 : r0;
    T *__restrict__ rn = odd ? r0 : r1;
    T alpha = Sc<T>::zero();
    if (sc.state[0] == ST_ACTIVE) {
        const T den = sc.dq[0];
        if (!Sc<T>::is_zero(den)) alpha = Sc<T>::div(sc.delta_new[0], den);
    }
    if constexpr (PEER) {
        if (sc.peer && sc.peer->world > 1)
            peer_push_rows<T>(sc.peer, odd ? 2 : 3, [&](int row) { return Sc<T>::fnma(alpha, q[row], r[row]); });
    }
    T acc[V];
#pragma unroll
    for (int v = 0; v < V; v++) acc[v] = Sc<T>::zero();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + t; p < npacks; p += stride) {
        const P qv = reinterpret_cast<const P *>(q)[p];
        P rv = reinterpret_cast<const P *>(r)[p];
#pragma unroll
        for (int v = 0; v < V; v++) {
            rv.v[v] = Sc<T>::fnma(alpha, qv.v[v], rv.v[v]);
            acc[v] = Sc<T>::fma(rv.v[v], rv.v[v], acc[v]);
        }
        reinterpret_cast<P *>(rn)[p] = rv;
    }
    if (V > 1 && blockIdx.x == 0) {
        const size_t e = npacks * V + t;
        if (e < nelem) {
            const T rv = Sc<T>::fnma(alpha, q[e], r[e]);
            rn[e] = rv;
            acc[0] = Sc<T>::fma(rv, rv, acc[0]);
        }
    }
#pragma unroll
    for (int v = 1; v < V; v++) acc[0] = Sc<T>::add(acc[0], acc[v]);
    T one[1] = {acc[0]};
    block_col_reduce<T, 1>(one, 1, smem);
    if (publish_and_arrive<T, 1>(smem, 1, 1, sc.partial, sc.ticket + TK_UPDATE)) {
        if (t == 0) trace_mark<T>(sc, it, TR_XR_ALL_DONE);
        const T total = cg2_grid_total<T, PEER>(sc, smem);
        if (t == 0) {
            update_bookkeep<T>(sc, 0, 1, it + 1, total);
            // what the next dir_spmv (or finish_x_kernel) needs: the step just taken and the new direction's weight
            sc.alpha[0] = alpha;
            T beta = Sc<T>::zero();
            if (sc.state[0] == ST_ACTIVE) {
                const T den = sc.delta_old[0];
                if (!Sc<T>::is_zero(den)) beta = Sc<T>::div(sc.delta_new[0], den);
            }
            sc.beta[0] = beta;
            *sc.it = it + 1;
            sc.ticket[TK_UPDATE] = 0;
            trace_mark<T>(sc, it, TR_XR_END);
        }
    }
}

// The update of x that is still pending when the loop ends: x += alpha_last d_last.
template <typename T>
__global__ void __launch_bounds__(256)
cg2_finish_x_kernel(size_t n, T *__restrict__ x, const T *d0, const T *d1, CgScalars<T> sc) {
    const T alpha = sc.alpha[0];
    if (Sc<T>::is_zero(alpha)) return;
    const T *__restrict__ d = (*sc.it & 1) ? d1 : d0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = Sc<T>::fma(alpha, d[i], x[i]);
    // (sc.alpha is cleared by the host-side caller's next initialisation; a second finish must not re-apply it)
}

}  // namespace cgb

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
//   update_r   alpha = delta_new/(d.q);  r -= alpha q  in place;  partial r.r
//              (axpy.cl + vdot.cl, clcg.c:326-386); bookkeeping (delta shuffle, beta, convergence,
//              clcg.c:350-356,389-391) by the last block.         reads q r, writes r'       (3 passes)
//
// x lags one update behind: x_k = x_{k-1} + alpha_{k-1} d_{k-1} is applied by dir_spmv of iteration k, and
// finish_x_kernel applies the last one.  All values are produced by the same FMAs in the same order as in the
// three-kernel path (the row sums follow CSR order, dn = fma(beta, d, r), x = fma(alpha, d, x),
// r = fma(-alpha, q, r)); only the association of the two dot-product reductions differs.
//
// The gather.  A block (one per SM, one row per thread) works on chunks of PAT_CHUNK consecutive rows.  The
// pattern dictionary knows every column offset (col - row) the matrix uses -- 7 for the 3-D Laplacian --;
// offsets that are closer than a chunk are merged into a WINDOW, and for a chunk starting at row c0 the
// window [lo, hi] needs exactly the vector entries [c0 + lo, c0 + hi + PAT_CHUNK).  Those pieces of r and of
// d arrive in shared memory by TMA bulk copies (cp.async.bulk + mbarrier, one elected thread issues them,
// nstage - 1 chunks ahead of the one being consumed), and the rows read their neighbours from shared memory
// at consecutive addresses (conflict-free).  Compared with 7 scattered 8-byte gathers per row and vector
// this is 3.6 contiguous elements per row and vector for the 300^3 Laplacian, none of them through
// registers or the L1 request path; what the threads still load themselves -- the 16-bit pattern number
// and x -- is requested one chunk ahead.  (First version, kept in git: windows staged THROUGH registers as
// r + beta d by 256-thread blocks, 3 per SM -- 885 us on C4, bound by ~7 dependent memory round trips per
// chunk with only 3 blocks to overlap them.)
//
// Row-block shards.  The direction has two buffers (the SpMV reads the old one everywhere while the owners write the
// new one), so dir_spmv is read-only on its inputs and the entries of dn a peer needs can be recomputed and stored
// straight into the peer's memory (NVLink) by dedicated warps of the SAME kernel, ahead of everything else: they go
// into the halo of the peer's RESIDUAL vector while the halos of both direction buffers stay zero, so that the
// peer's gather r[j] + beta d[j] yields exactly dn[j] there.  An arrival flag (numbered by the all-reduce count, as
// in kernels.cuh) tells the peer's TMA producer when its boundary chunks -- scheduled last -- may be loaded.
// No push kernel, no side stream, one exchange per iteration.  (First version, in git: r in two buffers as well and
// the boundary entries of r' pushed by update_r, flag-free; on one eighth of the 300^3 system the six vectors then
// no longer fit the 126 MB L2 and both kernels ran 30 % slower than alone, profiles/r02_notes.md.)
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

// The tail of a dir_spmv kernel: block sum of the consumers' partial dn.q in a fixed order (lanes by butterfly, then the
// warps one after the other), the last block to arrive sums the blocks' partials the same way, all-reduces over the
// GPUs (row-block shards) and publishes d.q.  `red` has 32 slots.
template <typename T, bool PEER, int NTHREADS, int NT>
__device__ __forceinline__ void dir_finish(T dot, T *red, const CgScalars<T> &sc, int it) {
    const int t = threadIdx.x;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        if constexpr (Sc<T>::cplx) {
            dot.x += __shfl_xor_sync(0xffffffffu, dot.x, off);
            dot.y += __shfl_xor_sync(0xffffffffu, dot.y, off);
        } else {
            dot += __shfl_xor_sync(0xffffffffu, dot, off);
        }
    }
    if ((t & 31) == 0) red[t >> 5] = dot;
    __syncthreads();
    if (t == 0) {
        T sum = Sc<T>::zero();
        for (int w = 0; w < NT / 32; w++) sum = Sc<T>::add(sum, red[w]);
        red[0] = sum;
    }
    __syncthreads();
    if (publish_and_arrive<T, 1>(red, 1, 1, sc.partial, sc.ticket + TK_SPMV)) {
        if (t == 0 && sc.trace) trace_mark<T>(sc, it, TR_SPMV_ALL_DONE);
        // sum of the per-block partials in a fixed order (grid_col_reduce wants a power-of-two block): strided
        // loads issued together, lanes by butterfly, then the warps one after the other
        __shared__ T s_total;
        __threadfence();
        T v = Sc<T>::zero();
        for (int b = t; b < (int)gridDim.x; b += NTHREADS) v = Sc<T>::add(v, ld_cg(sc.partial + b));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            if constexpr (Sc<T>::cplx) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
            } else {
                v += __shfl_xor_sync(0xffffffffu, v, off);
            }
        }
        if ((t & 31) == 0) red[t >> 5] = v;
        __syncthreads();
        if (t == 0) {
            T sum = Sc<T>::zero();
            for (int w = 0; w < NTHREADS / 32; w++) sum = Sc<T>::add(sum, red[w]);
            s_total = sum;
        }
        __syncthreads();
        T total = s_total;
        if constexpr (PEER) {
            if (sc.peer) total = peer_allreduce<T>(sc.peer, total);
        }
        if (t == 0) {
            sc.dq[0] = total;
            sc.ticket[TK_SPMV] = 0;
            if (sc.trace) trace_mark<T>(sc, it, TR_SPMV_END);
        }
    }
}

// dir_spmv is warp-specialised: DIR_CONSUMERS threads own the rows of a chunk (PAT_CHUNK / DIR_CONSUMERS rows
// each, independent accumulation chains), one more warp does nothing but feed the ring of stages with TMA.
constexpr int DIR_CONSUMERS = 512;
constexpr int DIR_THREADS = DIR_CONSUMERS + 32;         // + the TMA producer warp
constexpr int DIR_PUSH_WARPS = 3;                       // + the warps that store dn into the peers' halos (row-block shards)
constexpr int DIR_THREADS_PEER = DIR_THREADS + 32 * DIR_PUSH_WARPS;
constexpr int DIR_MAX_STAGES = 4;

// Row-block shards: DIR_PUSH_WARPS warps per block store dn = r + beta d of the rows the peers reference into the halos of their
// residual vectors, 16 entries in flight per lane, then raises the arrival flags.  It shares nothing with the rest of
// the block: the consumers start on the interior chunks at once.
template <typename T>
__device__ __forceinline__ void dir_push_halo(const CgScalars<T> &sc, T beta, const T *__restrict__ dold, const T *__restrict__ r) {
    const int t = threadIdx.x;
    const PeerComm *pc = sc.peer;
    if (!pc || pc->world <= 1) return;
    const int total = pc->send_off[pc->world];
    // 16 entries in flight per lane: with vectors that stream from HBM every dependent load of this warp (index, then
    // d and r) waits behind the TMA traffic of 148 producers -- at 8 in flight (5 rounds of 2 latencies for the two faces
    // of a 300 x 300 plane) the flags reached the neighbours ~45 us into the kernel and the blocks that start on a run
    // above the bottom halo plane sat idle that long (4 GPUs: 114 us for the inner ranks' dir_spmv against 67 us for
    // rank 0, profiles/r02_trace_c4_n4_before.txt)
    constexpr int PF = 16;
    // DIR_PUSH_WARPS warps per block share the list (they cost no registers: 17 or 20 warps, a partition holds five either way)
    const int lane = t & 31, pt = t - (DIR_CONSUMERS + 32), nl = (int)gridDim.x * (DIR_PUSH_WARPS * 32);
    for (int e0 = (int)blockIdx.x * (DIR_PUSH_WARPS * 32) + pt; e0 < total; e0 += PF * nl) {
        int row[PF];
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int e = e0 + u * nl;
            row[u] = e < total ? pc->send_idx[e] : -1;
        }
        T val[PF];
#pragma unroll
        for (int u = 0; u < PF; u++)
            if (row[u] >= 0) val[u] = Sc<T>::fma(beta, dold[row[u]], r[row[u]]);
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int e = e0 + u * nl;
            if (row[u] >= 0) {
                int p = 0;
                while (e >= pc->send_off[p + 1]) p++;
                reinterpret_cast<T *>(pc->vec[2][p])[pc->remote_off[p] + (e - pc->send_off[p])] = val[u];
            }
        }
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        // the last warp to get here (over all blocks) tells every peer that receives from this rank
        const unsigned prev = atomicAdd(&sc.peer->push_ticket[0], 1u);
        if (prev == gridDim.x * DIR_PUSH_WARPS - 1) {
            sc.peer->push_ticket[0] = 0;
            __threadfence_system();
            for (int p = 0; p < pc->world; p++)
                if (pc->send_off[p + 1] > pc->send_off[p]) st_release_sys_u64(pc->halo_flag[p] + pc->rank, pc->seq + 1);
        }
    }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory layout of dir_spmv (the host computes the same sizes, Engine::cg2_smem_bytes):
//   [0, 128)   mbarriers: full[stage] (the stage's copies have landed), empty[stage] (every consumer warp is done with it)
//   red        [32]                 block reduction scratch (one slot per warp)
//   s_val      [npat][STRIDE]       pattern coefficients
//   stages     nstage x { r window(s) [total], d window(s) [total] }   raw pieces of the two vectors, by TMA
//   s_pos      [npat][STRIDE] int   staging position of every pattern entry
//   s_len      [npat] int
template <typename T, int STRIDE, bool PEER>
__global__ void __launch_bounds__(PEER ? DIR_THREADS_PEER : DIR_THREADS, 1)
cg2_dir_spmv_kernel(int n, int ncols, int nchunks, int nchunks_interior, const int *__restrict__ chunks, int npat, int nstage,
                    PatWindows win, const unsigned short *__restrict__ pat, const unsigned *__restrict__ chunk_mask,
                    const int *__restrict__ p_len, const int *__restrict__ p_spos, const T *__restrict__ p_val,
                    T *__restrict__ x, T *__restrict__ q, const T *__restrict__ r, T *d0, T *d1, CgScalars<T> sc) {
    constexpr int NT = DIR_CONSUMERS;
    constexpr int NTHREADS = PEER ? DIR_THREADS_PEER : DIR_THREADS;
    constexpr int RPT = PAT_CHUNK / NT;                                        // rows per consumer thread and chunk
    constexpr int VPT = VecW<T>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *empty = full + DIR_MAX_STAGES;
    T *red = reinterpret_cast<T *>(smem_raw + 128);                            // [32]
    T *s_val = red + 32;                                                       // [npat][STRIDE]
    T *stage0 = s_val + npat * STRIDE;                                         // [nstage][2][win.total]
    int *s_pos = reinterpret_cast<int *>(stage0 + (size_t)nstage * 2 * win.total);
    int *s_len = s_pos + npat * STRIDE;
    const int t = threadIdx.x;
    const bool producer = t >= NT && t < NT + 32;
    const bool pusher = PEER && t >= NT + 32;
    // the table is part of the matrix, not the previous kernel's output: staged before the grid dependency wait
    for (int i = t; i < npat * STRIDE; i += NTHREADS) {
        const int id = i / STRIDE, j = i % STRIDE;
        s_val[i] = p_val[id * PAT_MAXLEN + j];
        s_pos[i] = p_spos[id * PAT_MAXLEN + j];
    }
    for (int i = t; i < npat; i += NTHREADS) s_len[i] = p_len[i];
    if (t == 0) {
        for (int s = 0; s < nstage; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NT / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (*sc.n_active == 0) return;
    const int it = *sc.it;
    if (sc.trace && blockIdx.x == 0 && t == 0) trace_mark<T>(sc, it, TR_SPMV_START);
    const bool odd = (it & 1) != 0;
    const T *__restrict__ dold = odd ? d1 : d0;
    T *__restrict__ dnew = odd ? d0 : d1;
    const T beta = sc.beta[0], alpha_prev = sc.alpha[0];
    // the block's i-th chunk is number chunks[blockIdx + i*grid] of the schedule (row-block shards: the chunks that
    // read halo entries come last), or blockIdx + i*grid itself
    const int count = ((int)blockIdx.x < nchunks) ? (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto chunk_of = [&](int i) { const int ci = (int)blockIdx.x + i * (int)gridDim.x; return chunks ? chunks[ci] : ci; };
    T dot = Sc<T>::zero();

    if (producer) {
        // ---- one elected thread: the windows of this block's i-th chunk, both vectors, into stage i % nstage,
        // as soon as the consumers have released the stage.  Parts of a window outside the vector are not
        // copied: no row of the chunk reads them.
        {
            const bool elected = t == NT;      // (the other lanes walk the loop with it and meet it at __syncwarp)
            const long long ncols_pad = ((long long)ncols + VPT - 1) / VPT * VPT;      // (every vector has >= 256 bytes of slack)
            int ch_next = (elected && count > 0) ? chunk_of(0) : 0;
            unsigned wmask = (elected && count > 0) ? chunk_mask[ch_next] : 0u;
            int s = 0;                          // stage i % nstage and the parity of its use, tracked without divisions
            unsigned use_parity = 1;            // ((i / nstage) - 1) & 1
            bool halo_ready = !(PEER && sc.peer && sc.peer->world > 1);
            for (int i = 0; i < count; i++, s++) {
                __syncwarp();
                if (s == nstage) {
                    s = 0;
                    use_parity ^= 1u;
                }
                if (!elected) continue;
                const int ch = ch_next;
                const long long c0 = (long long)ch * PAT_CHUNK;
                const unsigned wm = wmask;
                if (i + 1 < count) {                                                    // requested one chunk ahead
                    ch_next = chunk_of(i + 1);
                    wmask = chunk_mask[ch_next];
                }
                if constexpr (PEER) {
                    // the first chunk of this block that reads halo entries: the peers' stores of this iteration must have landed
                    if (!halo_ready && (int)blockIdx.x + i * (int)gridDim.x >= nchunks_interior) {
                        peer_wait_halo(sc.peer);
                        halo_ready = true;
                        if (sc.trace) trace_mark<T>(sc, it, TR_HALO_READY);
                    }
                }
                if (i >= nstage) mbar_wait(&empty[s], use_parity);
                T *Sr = stage0 + (size_t)s * 2 * win.total, *Sd = Sr + win.total;
                unsigned bytes = 0;
#pragma unroll
                for (int w = 0; w < WIN_MAX; w++) {          // (fully unrolled: `win` stays in the constant bank)
                    if (w >= win.nwin || !((wm >> w) & 1u)) continue;
                    const long long g0 = c0 + win.lo[w];
                    const long long a = g0 > 0 ? g0 : 0, b = min(g0 + win.size[w], ncols_pad);
                    if (b > a) bytes += (unsigned)((b - a) * sizeof(T));
                }
                mbar_arrive_expect_tx(&full[s], 2 * bytes);
#pragma unroll
                for (int w = 0; w < WIN_MAX; w++) {
                    if (w >= win.nwin || !((wm >> w) & 1u)) continue;
                    const long long g0 = c0 + win.lo[w];
                    const long long a = g0 > 0 ? g0 : 0, b = min(g0 + win.size[w], ncols_pad);
                    if (b <= a) continue;
                    const unsigned nb = (unsigned)((b - a) * sizeof(T));
                    bulk_g2s(Sr + win.base[w] + (a - g0), r + a, nb, &full[s]);
                    bulk_g2s(Sd + win.base[w] + (a - g0), dold + a, nb, &full[s]);
                }
            }
        }
    } else if (pusher) {
        if constexpr (PEER) dir_push_halo<T>(sc, beta, dold, r);
    } else {
        // The pattern number and x of a thread's rows are requested one chunk ahead.  The loads are unconditional
        // (clamped row) and nothing touches their result before the next chunk, so that the wait for a stage
        // below never waits for THEM (a select on the loaded value right after the load did exactly that: 37 % of
        // the stall samples of the first version, profiles/r02_notes.md).
        // (Tried and reverted, in git: a thread owning the PAIR of rows 2t, 2t + 1 with 16-byte accesses for x, q, dn and
        //  every even-positioned neighbour -- 7 shared-memory loads per row instead of 18.  On a 300-wide grid 43 % of
        //  the warps hold a pair whose rows have different patterns -- the first / last node of a grid line -- and run
        //  both the pair path and the per-row path: 46.5 instead of 41 us on the 38-plane slab, 324 instead of 280 on C4.)
        unsigned short id_cur[RPT];
        T x_cur[RPT];
        const long long last_row = (long long)n - 1;
        // (chunk numbers come from the schedule: the one after next is requested now, so that its rows' loads can be
        //  issued a whole chunk before they are needed)
        int ch_cur = count > 0 ? chunk_of(0) : 0, ch_nxt = count > 1 ? chunk_of(1) : 0;
#pragma unroll
        for (int s = 0; s < RPT; s++) {
            const long long row = min((long long)ch_cur * PAT_CHUNK + t + s * NT, last_row);
            id_cur[s] = pat[row];
            x_cur[s] = ld_stream_bytes(x + row);
        }
        // STRIDE 8: the pattern a thread used last stays in registers (on a grid nearly every row of a thread has
        // the same one); longer rows read the table out of shared memory entry by entry.  Positions are kept as
        // byte offsets: one add per shared-memory access.
        int cid = -1;
        int cposb[STRIDE <= 8 ? STRIDE : 1];
        T cval[STRIDE <= 8 ? STRIDE : 1];
        const unsigned d_off = (unsigned)win.total * (unsigned)sizeof(T);          // the d pieces follow the r pieces
        const unsigned diag_b = (unsigned)win.diag * (unsigned)sizeof(T);
        const unsigned stage_bytes = 2u * d_off;
        const unsigned char *stage_base = reinterpret_cast<const unsigned char *>(stage0);
        int st = 0;
        unsigned parity = 0;

        for (int i = 0; i < count; i++) {
            const int c0 = ch_cur * PAT_CHUNK;
            const int ch_nxt2 = i + 2 < count ? chunk_of(i + 2) : 0;
            unsigned short id_nxt[RPT];
            T x_nxt[RPT];
#pragma unroll
            for (int s = 0; s < RPT; s++) {
                const long long row = min((long long)ch_nxt * PAT_CHUNK + t + s * NT, last_row);
                id_nxt[s] = pat[row];
                x_nxt[s] = ld_stream_bytes(x + row);
            }
            const unsigned char *Sb = stage_base + (size_t)st * stage_bytes;
            mbar_wait(&full[st], parity);
#pragma unroll
            for (int s = 0; s < RPT; s++) {
                const int tl = t + s * NT, row = c0 + tl;
                if (row < n) {
                    const int id = id_cur[s];
                    const unsigned char *Srow = Sb + (size_t)tl * sizeof(T);             // the row's own position in the r pieces
                    T sum = Sc<T>::zero();
                    if constexpr (STRIDE <= 8) {
                        if (id != cid) {
                            cid = id;
#pragma unroll
                            for (int j = 0; j < STRIDE; j++) {
                                cposb[j] = s_pos[id * STRIDE + j] * (int)sizeof(T);
                                cval[j] = s_val[id * STRIDE + j];
                            }
                        }
                        // every neighbour loaded before the first FMA; padded entries point at the row's own
                        // position with coefficient 0 (as spmv_pattern_kernel pads)
                        T rw[STRIDE], dw[STRIDE];
#pragma unroll
                        for (int j = 0; j < STRIDE; j++) {
                            rw[j] = *reinterpret_cast<const T *>(Srow + cposb[j]);
                            dw[j] = *reinterpret_cast<const T *>(Srow + cposb[j] + d_off);
                        }
#pragma unroll
                        for (int j = 0; j < STRIDE; j++) sum = Sc<T>::fma(cval[j], Sc<T>::fma(beta, dw[j], rw[j]), sum);
                    } else {
                        const int len = s_len[id];
                        for (int j = 0; j < len; j++) {
                            const unsigned char *pp = Srow + (size_t)s_pos[id * STRIDE + j] * sizeof(T);
                            sum = Sc<T>::fma(s_val[id * STRIDE + j],
                                             Sc<T>::fma(beta, *reinterpret_cast<const T *>(pp + d_off), *reinterpret_cast<const T *>(pp)), sum);
                        }
                    }
                    const T dd = *reinterpret_cast<const T *>(Srow + diag_b + d_off);
                    const T dn = Sc<T>::fma(beta, dd, *reinterpret_cast<const T *>(Srow + diag_b));
                    q[row] = sum;
                    dnew[row] = dn;
                    st_stream_bytes(x + row, Sc<T>::fma(alpha_prev, dd, x_cur[s]));
                    dot = Sc<T>::fma(dn, sum, dot);
                }
                id_cur[s] = id_nxt[s];
                x_cur[s] = x_nxt[s];
            }
            __syncwarp();
            if ((t & 31) == 0) mbar_arrive(&empty[st]);          // this warp has read everything it needs from the stage
            ch_cur = ch_nxt;
            ch_nxt = ch_nxt2;
            if (++st == nstage) {
                st = 0;
                parity ^= 1u;
            }
        }
    }

    dir_finish<T, PEER, NTHREADS, NT>(dot, red, sc, it);
}

// alpha = delta_new / dq ; r -= alpha q ; delta_old = delta_new ; delta_new = r.r ; beta ; convergence bookkeeping.
//                                                                                                 clcg.c:326-392
//
// The tail of this kernel is on the critical path of every iteration (on shards it also carries the all-reduce), so
// everything the last block needs from memory -- the column's state, delta_0, the tolerance -- is read by every
// block up front, next to the scalars alpha needs anyway; after the sum is known the tail is arithmetic and stores.
template <typename T, int V, bool PEER>
__global__ void __launch_bounds__(256)
cg2_update_r_kernel(size_t npacks, size_t nelem, const T *__restrict__ q, T *__restrict__ r, CgScalars<T> sc) {
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (*sc.n_active == 0) return;
    __shared__ T s_red[8];
    __shared__ T s_total;
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    const int it = *sc.it;
    if (sc.trace && blockIdx.x == 0 && t == 0) trace_mark<T>(sc, it, TR_XR_START);
    const int state0 = sc.state[0];
    const T delta_cur = sc.delta_new[0];
    const double delta0 = sc.delta0[0], tol = *sc.tol;
    T alpha = Sc<T>::zero();
    if (state0 == ST_ACTIVE) {
        const T den = sc.dq[0];
        if (!Sc<T>::is_zero(den)) alpha = Sc<T>::div(delta_cur, den);
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
        reinterpret_cast<P *>(r)[p] = rv;
    }
    if (V > 1 && blockIdx.x == 0) {
        const size_t e = npacks * V + t;
        if (e < nelem) {
            const T rv = Sc<T>::fnma(alpha, q[e], r[e]);
            r[e] = rv;
            acc[0] = Sc<T>::fma(rv, rv, acc[0]);
        }
    }
#pragma unroll
    for (int v = 1; v < V; v++) acc[0] = Sc<T>::add(acc[0], acc[v]);

    // block sum, then (last block) grid sum, in a fixed order: lanes by butterfly, warps one after the other
    auto block_sum = [&](T v) -> T {            // valid in thread 0
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            if constexpr (Sc<T>::cplx) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
            } else {
                v += __shfl_xor_sync(0xffffffffu, v, off);
            }
        }
        __syncthreads();
        if ((t & 31) == 0) s_red[t >> 5] = v;
        __syncthreads();
        T sum = Sc<T>::zero();
        if (t == 0)
            for (int w = 0; w < 8; w++) sum = Sc<T>::add(sum, s_red[w]);
        return sum;
    };
    const T mine = block_sum(acc[0]);
    __shared__ int s_last;
    if (t == 0) {
        sc.partial[blockIdx.x] = mine;
        __threadfence();
        s_last = (atomicAdd(sc.ticket + TK_UPDATE, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (t == 0) trace_mark<T>(sc, it, TR_XR_ALL_DONE);
    __threadfence();
    T v = Sc<T>::zero();
    for (int b = t; b < (int)gridDim.x; b += 256) v = Sc<T>::add(v, ld_cg(sc.partial + b));
    v = block_sum(v);
    if (t == 0) s_total = v;
    __syncthreads();
    T total = s_total;
    if constexpr (PEER) {
        if (sc.peer) total = peer_allreduce<T>(sc.peer, total);
    }
    if (t == 0) {
        // delta shuffle (clcg.c:350-356), convergence test, history: update_bookkeep with the operands already in registers
        const int it1 = it + 1;
        T beta = Sc<T>::zero();
        if (state0 == ST_ACTIVE) {
            sc.delta_old[0] = delta_cur;
            sc.delta_new[0] = total;
            const double a = Sc<T>::abs(total);
            int st = ST_ACTIVE;
            if (!Sc<T>::finite(total)) st = ST_BREAKDOWN;
            else if (a == 0.0 || (tol > 0.0 && sqrt(a / delta0) < tol)) st = ST_CONVERGED;
            if (st != ST_ACTIVE) {
                sc.state[0] = st;
                sc.iters[0] = it1;
                *sc.n_active = 0;
            } else if (!Sc<T>::is_zero(delta_cur)) {
                beta = Sc<T>::div(total, delta_cur);
            }
        }
        if (sc.hist && it1 < sc.hist_cap)
            Sc<T>::to_double2(state0 == ST_ACTIVE ? total : delta_cur, sc.hist + (size_t)it1 * (Sc<T>::cplx ? 2 : 1));
        // what the next dir_spmv (or finish_x_kernel) needs: the step just taken and the new direction's weight
        sc.alpha[0] = alpha;
        sc.beta[0] = beta;
        *sc.it = it1;
        sc.ticket[TK_UPDATE] = 0;
        trace_mark<T>(sc, it, TR_XR_END);
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

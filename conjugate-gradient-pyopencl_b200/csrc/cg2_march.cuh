// cg2_march.cuh -- dir_spmv for grid operators with ONE far offset: a block marches through the planes.
//
// cg2_dir_spmv_kernel stages, for every chunk of rows, every window of the vectors its neighbours live in: for the
// 7-point Laplacian on a 300^3 grid that is the chunk itself widened by +-300 AND the two pieces one plane (+-90000)
// away -- 3.6 elements per row and vector, 58 KB per chunk, which leaves room for three stages only and moves each
// element through L2 and shared memory 3.6 times.
//
// When the far offsets are +-P (+ the near ones) and the chunk length `ch` divides P, the piece "one plane up" of
// chunk c IS the own piece of chunk c + P/ch.  A block therefore takes a strip of a plane (ch consecutive rows) and
// marches with it from plane to plane: the stage it loaded for plane z serves as the "down" piece at z + 1 and did
// serve as the "up" piece at z - 1.  Every element is loaded once per strip and run (plus the +-near overlap between
// neighbouring strips): 1.6 elements per row and vector instead of 3.6, a stage is 26 KB instead of 58, and eight
// stages fit -- five loads ahead of the consumers instead of two.
//
// Work = runs: (strip, first plane, planes).  A run of L planes loads L + 2 pieces (one below, one above), so runs are
// made as long as the balance over the SMs allows (Engine::plan_march).  Row-block shards: the piece below the first /
// above the last owned plane is the halo plane the peer filled (dir_push_halo): the runs that touch it are scheduled
// last, and the producer waits for the arrival flag before it loads such a piece.
//
// Same FMAs in the same order as cg2_dir_spmv_kernel and the three-kernel iteration.
#pragma once

namespace cgb {

constexpr int MARCH_MAX_STAGES = 8;

struct MarchPlan {
    int ok;
    int ch;              // rows per chunk; divides the plane
    int m;               // chunks (strips) per plane
    int nplanes;         // owned planes: n / (ch * m)
    int lo0;             // first staged column relative to the chunk's first row (<= 0, multiple of the pack width)
    int stage_el;        // elements per vector and stage (multiple of the pack width)
    int has_low, has_high;        // a halo plane below the first / above the last owned plane (row-block shards)
    long long low_src, high_src;  // its first element in the vectors
    int nruns;
    int grid;            // blocks the runs were dealt to: the launch grid
};
struct MarchRun {
    int strip, z0, len;
    int next;            // index of the same block's next run, -1: none (block b starts with runs[b])
};

// The sequence of pieces a block loads: for each of its runs the planes z0 - 1 .. z0 + len.
struct MarchWalk {
    const MarchRun *runs;
    int j;
    MarchRun run, nxt;
    bool have_nxt;
    __device__ __forceinline__ bool start(const MarchRun *r, int n) {
        runs = r;
        j = -1;
        if ((int)blockIdx.x >= n) return false;
        run = runs[blockIdx.x];
        have_nxt = run.next >= 0;
        if (have_nxt) nxt = runs[run.next];          // (requested a whole run ahead)
        return true;
    }
    __device__ __forceinline__ bool next() {
        if (j < run.len) {
            j++;
            return true;
        }
        if (!have_nxt) return false;
        run = nxt;
        j = -1;
        have_nxt = run.next >= 0;
        if (have_nxt) nxt = runs[run.next];
        return true;
    }
    __device__ __forceinline__ bool computes() const { return j >= 0 && j < run.len; }
};

template <typename T> __host__ __device__ inline unsigned march_stage_bytes(const MarchPlan &mp) {
    return 2u * (unsigned)mp.stage_el * (unsigned)sizeof(T) + (unsigned)mp.ch * (unsigned)sizeof(T) +
           (((unsigned)mp.ch * 2u + 15u) & ~15u);
}

template <typename T, int STRIDE, bool PEER>
__global__ void __launch_bounds__(PEER ? DIR_THREADS_PEER : DIR_THREADS, 1)
cg2_dir_march_kernel(int n, int ncols, MarchPlan mp, int nstage, const MarchRun *__restrict__ runs, int npat,
                     const unsigned short *__restrict__ pat, const int *__restrict__ p_len, const int *__restrict__ p_spos,
                     const T *__restrict__ p_val, T *__restrict__ x, T *__restrict__ q, const T *__restrict__ r, T *d0, T *d1,
                     CgScalars<T> sc) {
    constexpr int NT = DIR_CONSUMERS;
    constexpr int NTHREADS = PEER ? DIR_THREADS_PEER : DIR_THREADS;
    constexpr int VPT = VecW<T>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw);           // [MARCH_MAX_STAGES]
    unsigned long long *empty = full + MARCH_MAX_STAGES;
    T *red = reinterpret_cast<T *>(smem_raw + 128);                            // [32]
    T *s_val = red + 32;                                                       // [npat][STRIDE]
    unsigned char *stage0 = reinterpret_cast<unsigned char *>(s_val + npat * STRIDE);     // [nstage] stages, see below
    int *s_pos = reinterpret_cast<int *>(stage0 + (size_t)nstage * march_stage_bytes<T>(mp));   // (rel + 1) << 24 | position
    int *s_len = s_pos + npat * STRIDE;
    const int t = threadIdx.x;
    const bool producer = t >= NT && t < NT + 32;
    const bool pusher = PEER && t >= NT + 32;
    for (int i = t; i < npat * STRIDE; i += NTHREADS) {
        const int id = i / STRIDE, e = i % STRIDE;
        s_val[i] = p_val[id * PAT_MAXLEN + e];
        s_pos[i] = p_spos[id * PAT_MAXLEN + e];
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
    // a stage: the r piece, the d piece, and -- for the planes the block computes on -- the chunk's x and pattern numbers
    const unsigned d_off = (unsigned)mp.stage_el * (unsigned)sizeof(T);
    const unsigned x_off = 2u * d_off;
    const unsigned id_off = x_off + (unsigned)mp.ch * (unsigned)sizeof(T);
    const unsigned stage_bytes = march_stage_bytes<T>(mp);
    T dot = Sc<T>::zero();

    if (producer) {
        // ---- one elected thread walks the block's pieces and loads each into the next stage of the ring as soon as the
        // consumers have released that stage
        const bool elected = t == NT;
        const long long ncols_pad = ((long long)ncols + VPT - 1) / VPT * VPT;              // (every vector has >= 256 bytes of slack)
        const unsigned long long stream_policy = l2_evict_first_policy();
        MarchWalk w;
        bool more = w.start(runs, mp.nruns);
        int s = 0;
        unsigned use_parity = 1;                // ((item / nstage) - 1) & 1
        bool halo_ready = !(PEER && sc.peer && sc.peer->world > 1);
        for (int item = 0; more; item++, s++) {
            __syncwarp();
            if (s == nstage) {
                s = 0;
                use_parity ^= 1u;
            }
            if (elected) {
                const int cz = w.run.z0 + w.j;
                long long elem0 = -1;
                bool halo = false;
                if (cz >= 0 && cz < mp.nplanes) {
                    elem0 = ((long long)cz * mp.m + w.run.strip) * mp.ch;
                } else if (cz == -1 && mp.has_low) {
                    elem0 = mp.low_src + (long long)w.run.strip * mp.ch;
                    halo = true;
                } else if (cz == mp.nplanes && mp.has_high) {
                    elem0 = mp.high_src + (long long)w.run.strip * mp.ch;
                    halo = true;
                }
                if constexpr (PEER) {
                    if (halo && !halo_ready) {          // the peers' stores of this iteration must have landed
                        peer_wait_halo(sc.peer);
                        halo_ready = true;
                        if (sc.trace) trace_mark<T>(sc, it, TR_HALO_READY);
                    }
                }
                if (item >= nstage) mbar_wait(&empty[s], use_parity);
                unsigned char *St = stage0 + (size_t)s * stage_bytes;
                T *Sr = reinterpret_cast<T *>(St), *Sd = reinterpret_cast<T *>(St + d_off);
                unsigned nb = 0;
                long long a = 0, g0 = 0;
                if (elem0 >= 0) {
                    g0 = elem0 + mp.lo0;
                    a = g0 > 0 ? g0 : 0;
                    const long long b = min(g0 + mp.stage_el, ncols_pad);
                    if (b > a) nb = (unsigned)((b - a) * sizeof(T));
                }
                // x and the pattern numbers of the chunk: only where the block computes (not for the pieces below / above a run)
                const bool comp = w.computes();
                const unsigned xb = comp ? (unsigned)mp.ch * (unsigned)sizeof(T) : 0u, ib = comp ? (unsigned)mp.ch * 2u : 0u;
                mbar_arrive_expect_tx(&full[s], 2 * nb + xb + ib);
                if (nb) {
                    bulk_g2s(Sr + (a - g0), r + a, nb, &full[s]);
                    bulk_g2s(Sd + (a - g0), dold + a, nb, &full[s]);
                }
                if (comp) {
                    bulk_g2s_hint(St + x_off, x + elem0, xb, &full[s], stream_policy);      // x is touched once per iteration: evict-first
                    bulk_g2s(St + id_off, pat + elem0, ib, &full[s]);
                }
            }
            more = w.next();
        }
    } else if (pusher) {
        if constexpr (PEER) dir_push_halo<T>(sc, beta, dold, r);
    } else {
        // ---- consumers: rows t, t + 512 of the chunk (if < ch).  Everything they read -- the pieces of r and d, x, the
        // pattern numbers -- arrives in the stage by TMA: no global load, no register that waits for one.  (First version, in
        // git: x and the pattern number requested one chunk ahead into registers, as cg2_dir_spmv_kernel does; here the
        // compiler rotated those registers with MOVs that wait for the loads -- 46 % of the stall samples,
        // profiles/r02_notes.md.)  Ring positions are tracked incrementally: a division by the (run-time)
        // number of stages costs ~100 dependent cycles, and there would be five per piece.
        const unsigned diag_b = (unsigned)(-mp.lo0) * (unsigned)sizeof(T);
        MarchWalk w;
        bool more = w.start(runs, mp.nruns);
        auto row_base = [&](const MarchWalk &ww) { return ((ww.run.z0 + ww.j) * mp.m + ww.run.strip) * mp.ch; };
        int cid = -1;
        int cposb[STRIDE <= 8 ? STRIDE : 1];
        int crel[STRIDE <= 8 ? STRIDE : 1];
        T cval[STRIDE <= 8 ? STRIDE : 1];
        // slot of the current piece, of the one before it, and the wait cursor (slot + parity of the next piece to wait for)
        int s_cur = 0, s_prev = nstage - 1;
        int wait_slot = 0;                      // the wait cursor: slot and parity of the first piece not waited for yet
        unsigned wait_parity = 0;
        int ahead_of_item = 0;                  // pieces item, item + 1, .. item + ahead_of_item - 1 have been waited for
        for (int item = 0; more; item++) {
            const int s_next = s_cur + 1 == nstage ? 0 : s_cur + 1;
            if (w.computes()) {
                // the pieces of the planes below, at and above this one must have landed: `ahead_of_item` counts how many
                // pieces starting at this one have been waited for already (the wait cursor never falls behind item - 1)
                while (ahead_of_item < 2) {
                    mbar_wait(&full[wait_slot], wait_parity);
                    if (++wait_slot == nstage) {
                        wait_slot = 0;
                        wait_parity ^= 1u;
                    }
                    ahead_of_item++;
                }
                const unsigned char *S0 = stage0 + (size_t)s_cur * stage_bytes;
                const unsigned char *Sm = stage0 + (size_t)s_prev * stage_bytes;
                const unsigned char *Sp = stage0 + (size_t)s_next * stage_bytes;
                const int c0 = row_base(w);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const int tl = t + s * NT;
                    const int row = c0 + tl;
                    if (tl < mp.ch && row < n) {
                        const int id = reinterpret_cast<const unsigned short *>(S0 + id_off)[tl];
                        const size_t tb = (size_t)tl * sizeof(T);
                        T sum = Sc<T>::zero();
                        if constexpr (STRIDE <= 8) {
                            if (id != cid) {
                                cid = id;
#pragma unroll
                                for (int e = 0; e < STRIDE; e++) {
                                    const int code = s_pos[id * STRIDE + e];
                                    cposb[e] = (code & 0xffffff) * (int)sizeof(T);
                                    crel[e] = code >> 24;
                                    cval[e] = s_val[id * STRIDE + e];
                                }
                            }
                            // a batch of neighbours loaded before their FMAs (16-byte values: half a batch, or it spills)
                            constexpr int HB = sizeof(T) == 16 ? STRIDE / 2 : STRIDE;
#pragma unroll
                            for (int h = 0; h < STRIDE; h += HB) {
                                T rw[HB], dw[HB];
#pragma unroll
                                for (int e = 0; e < HB; e++) {
                                    const unsigned char *pp = (crel[h + e] == 1 ? S0 : (crel[h + e] == 0 ? Sm : Sp)) + cposb[h + e] + tb;
                                    rw[e] = *reinterpret_cast<const T *>(pp);
                                    dw[e] = *reinterpret_cast<const T *>(pp + d_off);
                                }
#pragma unroll
                                for (int e = 0; e < HB; e++) sum = Sc<T>::fma(cval[h + e], Sc<T>::fma(beta, dw[e], rw[e]), sum);
                            }
                        } else {
                            const int len = s_len[id];
                            for (int e = 0; e < len; e++) {
                                const int code = s_pos[id * STRIDE + e];
                                const int rel = code >> 24;
                                const unsigned char *pp = (rel == 1 ? S0 : (rel == 0 ? Sm : Sp)) + (size_t)(code & 0xffffff) * sizeof(T) + tb;
                                sum = Sc<T>::fma(s_val[id * STRIDE + e],
                                                 Sc<T>::fma(beta, *reinterpret_cast<const T *>(pp + d_off), *reinterpret_cast<const T *>(pp)), sum);
                            }
                        }
                        const T dd = *reinterpret_cast<const T *>(S0 + diag_b + tb + d_off);
                        const T dn = Sc<T>::fma(beta, dd, *reinterpret_cast<const T *>(S0 + diag_b + tb));
                        q[row] = sum;
                        dnew[row] = dn;
                        st_stream_bytes(x + row, Sc<T>::fma(alpha_prev, dd, *reinterpret_cast<const T *>(S0 + x_off + tb)));
                        dot = Sc<T>::fma(dn, sum, dot);
                    }
                }
                __syncwarp();
                if ((t & 31) == 0) mbar_arrive(&empty[s_prev]);        // the piece below is not needed any more
            } else if (w.j == w.run.len) {
                // the piece above the run's last plane: with it the run's last two pieces are done
                // (it has been waited for by the computation on the plane below it)
                __syncwarp();
                if ((t & 31) == 0) {
                    mbar_arrive(&empty[s_prev]);
                    mbar_arrive(&empty[s_cur]);
                }
            }
            // on to the next piece: the wait cursor is now one piece less far ahead of it
            if (ahead_of_item > 0) {
                ahead_of_item--;
            } else {
                // this piece was never waited for (the piece below a run's first plane): the cursor moves with the piece
                // -- but its barrier phase still has to be observed before the slot's next phase can be waited for
                mbar_wait(&full[wait_slot], wait_parity);
                if (++wait_slot == nstage) {
                    wait_slot = 0;
                    wait_parity ^= 1u;
                }
            }
            s_prev = s_cur;
            s_cur = s_next;
            more = w.next();
        }
    }
    dir_finish<T, PEER, NTHREADS, NT>(dot, red, sc, it);
}

}  // namespace cgb

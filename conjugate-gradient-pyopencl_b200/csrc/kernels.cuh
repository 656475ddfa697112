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
#include <cooperative_groups.h>

#include "scalar.cuh"

namespace cgb {

enum : int { ST_ACTIVE = 0, ST_CONVERGED = 1, ST_BREAKDOWN = 2 };
enum : int { TK_SPMV = 0, TK_UPDATE = 1, TK_INIT = 2 };

// ---------------------------------------------------------------------------
// Peer memory: the collectives of the row-block sharded solve done by the compute kernels
// themselves through NVLink-mapped pointers (CUDA IPC), instead of separate NCCL launches.
//
//   all-reduce of a dot product   the last block of spmv_dot / update_xr stores its GPU's
//       partial sum straight into a slot in EVERY peer's memory -- 8-byte words that carry the
//       sequence number next to 32 bits of payload, so no fence separates "value" from "flag" --,
//       then waits until all peers' words of the same sequence number have arrived in its own
//       memory and adds them up in rank order: every GPU gets the bit-identical sum, in about
//       one NVLink one-way trip, inside the kernel that produced the partial.
//   halo of d                     halo_push_kernel writes the entries a peer needs directly
//       into that peer's d vector and then raises a flag there; the peer's SpMV waits for
//       the flags after it has already started streaming its matrix tiles.
//
// The two dot products per iteration double as barriers, so neither the slots (double
// buffered by sequence parity) nor the halo regions can be overwritten while still in use.
// ---------------------------------------------------------------------------
constexpr int PEER_MAX = 8;
struct PeerSlot {              // 32 bytes: four self-validating words, (sequence number << 32) | 32 bits of payload
    unsigned long long w[4];   // re low, re high, im low, im high
};
struct PeerComm {
    int rank, world;
    unsigned long long seq;                     // all-reduces completed on this GPU
    unsigned long long halo_seq;                // (unused since an exchange is numbered by the all-reduce count)
    PeerSlot *slots[PEER_MAX];                  // slots[p]: rank p's [2][world] slot array (peer mapped; [rank] is local)
    unsigned long long *halo_flag[PEER_MAX];    // halo_flag[p]: rank p's [world] arrival flags
    void *d_peer[PEER_MAX];                     // rank p's direction vector [owned | halo]
    long long remote_off[PEER_MAX];             // element offset in rank p's d where this rank's entries land
    int send_off[PEER_MAX + 1];                 // this rank's send list is segmented by destination rank
    int recv_from[PEER_MAX];                    // 1 if rank p sends halo entries to this rank
    unsigned int push_ticket[PEER_MAX];
    // two-kernel iteration (cg2.cuh): rank p's direction buffers d0, d1 and residual vector r (index 2), each
    // [owned | halo]; dir_spmv stores the entries of the new direction rank p references into the halo of its r
    void *vec[4][PEER_MAX];
    const int *send_idx;                        // this rank's owned rows to send, segmented by destination (send_off)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ALL threads of ONE block per GPU.  Returns the sum over the GPUs (same bits everywhere).
//
// Every 8-byte word a GPU stores into a peer carries the sequence number of the all-reduce in its upper
// half and 32 bits of the value in its lower half (the scheme of NCCL's LL protocol): an aligned 8-byte
// store is single-copy atomic, so a word whose tag matches IS the data -- no fence between "value" and
// "flag", the cost is one NVLink one-way trip instead of a store, a system-scope release (a round trip)
// and a second store.
template <typename T> __device__ __forceinline__ T peer_allreduce(PeerComm *pc, T local) {
    __shared__ double s_val[2][PEER_MAX];
    __shared__ double s_sum[2];
    constexpr int NW = Sc<T>::cplx ? 4 : 2;
    const int t = threadIdx.x;
    const unsigned long long seq = pc->seq + 1;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const int parity = (int)(seq & 1);
    if (t < pc->world) {
        // Ordering of the halo entries the blocks of this kernel stored into the peers' vectors (cg2.cuh) against the
        // words below: every storing thread executed a system-scope fence after its stores and before its block
        // arrived at the kernel's ticket, i.e. those stores were performed at the peer before the last block --
        // this one -- even started.  (A second system fence here, in front of the words, cost ~2 us per
        // all-reduce: profiles/r02_notes.md.)
        double v[2] = {0.0, 0.0};
        Sc<T>::to_double2(local, v);
        PeerSlot *dst = pc->slots[t] + (size_t)parity * pc->world + pc->rank;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v[i >> 1]);
            st_relaxed_sys_u64(&dst->w[i], tag | ((i & 1) ? (bits >> 32) : (bits & 0xffffffffull)));
        }
        const PeerSlot *src = pc->slots[pc->rank] + (size_t)parity * pc->world + t;
        unsigned long long w[4] = {0, 0, 0, 0};
        unsigned long long spins = 0;
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < NW; i++) {
                w[i] = ld_relaxed_sys_u64(&src->w[i]);
                ok = ok && ((w[i] & 0xffffffff00000000ull) == tag);
            }
            if (ok) break;
            if (++spins > (1ull << 31)) __trap();
        }
        // (what the peers stored into this GPU's halos is read by the NEXT kernel: the kernel boundary orders it)
        s_val[0][t] = __longlong_as_double((long long)(((w[1] & 0xffffffffull) << 32) | (w[0] & 0xffffffffull)));
        s_val[1][t] = NW == 4 ? __longlong_as_double((long long)(((w[3] & 0xffffffffull) << 32) | (w[2] & 0xffffffffull))) : 0.0;
    }
    __syncthreads();
    if (t == 0) {
        double re = 0.0, im = 0.0;
        for (int p = 0; p < pc->world; p++) {      // rank order: the same bits on every GPU
            re += s_val[0][p];
            im += s_val[1][p];
        }
        s_sum[0] = re;
        s_sum[1] = im;
        pc->seq = seq;
    }
    __syncthreads();
    if constexpr (Sc<T>::cplx) return Sc<T>::make((typename Sc<T>::real)s_sum[0], (typename Sc<T>::real)s_sum[1]);
    else return (T)s_sum[0];
}

// Waits (one thread per block) until every sending peer's entries of the current exchange have landed.
// An exchange is identified by the number of all-reduces completed so far (+1): every exchange of the
// solve is separated from the next by at least one all-reduce, and all GPUs count them in lockstep.
__device__ __forceinline__ void peer_wait_halo(const PeerComm *pc) {
    const unsigned long long want = pc->seq + 1;
    for (int p = 0; p < pc->world; p++) {
        if (!pc->recv_from[p]) continue;
        unsigned long long spins = 0;
        while (ld_acquire_sys_u64(pc->halo_flag[pc->rank] + p) < want)
            if (++spins > (1ull << 31)) __trap();
    }
}

// Stores value_of_row(i) for every owned row i that a peer references into that peer's halo of vector
// `which` (PeerComm::vec).  Grid-strided over the send list, run by every block ahead of its own work so
// that the NVLink transfer overlaps the rest of the kernel.
template <typename T, typename F>
__device__ __forceinline__ bool peer_push_rows(const PeerComm *pc, int which, F value_of_row) {
    const int total = pc->send_off[pc->world];
    bool stored = false;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        int p = 0;
        while (e >= pc->send_off[p + 1]) p++;
        T *dst = reinterpret_cast<T *>(pc->vec[which][p]) + pc->remote_off[p] + (e - pc->send_off[p]);
        *dst = value_of_row(pc->send_idx[e]);
        stored = true;
    }
    // The caller fences (__threadfence_system) before its block arrives at the kernel's ticket -- at the END of its
    // work, when the stores have long been acknowledged, not here where the fence would wait a full NVLink round trip.
    return stored;
}

// Programmatic dependent launch: the three kernels of an iteration are launched with the
// programmatic-stream-serialization attribute, so the blocks of the next kernel are placed on SMs as
// the blocks of the running one retire and run their prologue (barrier set-up, the first matrix tiles
// by TMA -- nothing a previous kernel writes) before `pdl_wait` returns, which is when the previous
// kernel has completed and its writes are visible.  Every block of every kernel waits, so completion
// is transitive along the chain.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// timeline events of one iteration (CgScalars::trace, 8 slots per iteration)
enum : int { TR_SPMV_START = 0, TR_SPMV_ALL_DONE = 1, TR_SPMV_END = 2, TR_XR_START = 3, TR_XR_ALL_DONE = 4, TR_XR_END = 5,
             TR_D_START = 6, TR_HALO_READY = 7 };

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
    const double *tol;  // in device memory, so that one captured graph serves every tolerance
    T *rr;              // [k] this device's part of r.r when `defer` is set
    PeerComm *peer;     // non-NULL: the dot products are all-reduced inside the kernels through peer memory
    int l2_keep;        // 1: d, q and r are tagged evict-last in L2 (the system's vectors fit the L2)
    int pdl_early;      // 1: let the next kernel's blocks become resident as soon as this one has started
    unsigned long long *trace;   // optional timeline, [trace_cap][8] globaltimer stamps (TR_*)
    int trace_cap;
    int defer;          // row-block sharded solve: the dot products are only partial sums here; the
                        // bookkeeping runs in init_bookkeep_kernel / update_bookkeep_kernel after the
                        // all-reduce over the devices (dq is all-reduced in place)
    T *alpha;           // [k] two-kernel iteration (cg2.cuh): the step of the latest update (x lags by it)
    T *beta;            // [k] ... and the weight of the old direction in the next one
    int cg2;            // 1: the solve runs the two-kernel iteration
};

template <typename T> __device__ __forceinline__ void trace_mark(const CgScalars<T> &sc, int it, int ev) {
    if (sc.trace && it >= 0 && it < sc.trace_cap) sc.trace[(size_t)it * 8 + ev] = global_timer_ns();
}

// After delta = r0.r0 is known for column c.                      clcg.c:274-292
template <typename T> __device__ __forceinline__ void init_bookkeep(const CgScalars<T> &sc, int c, T dl) {
    const double a0 = Sc<T>::abs(dl);
    sc.delta_new[c] = dl;
    sc.delta_old[c] = dl;
    sc.dq[c] = Sc<T>::zero();
    sc.delta0[c] = a0;
    const bool live = (a0 > 0.0) && Sc<T>::finite(dl);
    sc.state[c] = live ? ST_ACTIVE : (a0 == 0.0 ? ST_CONVERGED : ST_BREAKDOWN);
    sc.iters[c] = 0;
    sc.alpha[c] = Sc<T>::zero();
    sc.beta[c] = Sc<T>::zero();
    if (sc.hist && sc.hist_cap > 0) Sc<T>::to_double2(dl, sc.hist + (size_t)c * (Sc<T>::cplx ? 2 : 1));
}

// After the new r.r is known for column c: delta shuffle (clcg.c:350-356), convergence test, history.
template <typename T> __device__ __forceinline__ void update_bookkeep(const CgScalars<T> &sc, int c, int k, int it1, T nd) {
    if (sc.state[c] == ST_ACTIVE) {
        sc.delta_old[c] = sc.delta_new[c];
        sc.delta_new[c] = nd;
        const double a = Sc<T>::abs(nd);
        int st = ST_ACTIVE;
        if (!Sc<T>::finite(nd)) st = ST_BREAKDOWN;
        else if (a == 0.0 || (*sc.tol > 0.0 && sqrt(a / sc.delta0[c]) < *sc.tol)) st = ST_CONVERGED;
        if (st != ST_ACTIVE) {
            sc.state[c] = st;
            sc.iters[c] = it1;
            atomicSub(sc.n_active, 1);
        }
    }
    if (sc.hist && it1 < sc.hist_cap)
        Sc<T>::to_double2(sc.delta_new[c], sc.hist + ((size_t)it1 * k + c) * (Sc<T>::cplx ? 2 : 1));
}

// The deferred halves of the two finalisations, one block, run after the all-reduce of sc.rr.
template <typename T> __global__ void init_bookkeep_kernel(int k, CgScalars<T> sc) {
    for (int c = threadIdx.x; c < k; c += blockDim.x) init_bookkeep<T>(sc, c, sc.rr[c]);
    __syncthreads();
    if (threadIdx.x == 0) {
        int live = 0;
        for (int c = 0; c < k; c++) live += (sc.state[c] == ST_ACTIVE);
        *sc.n_active = live;
        *sc.it = 0;
    }
}
template <typename T> __global__ void update_bookkeep_kernel(int k, CgScalars<T> sc) {
    if (*sc.n_active == 0) return;
    const int it1 = *sc.it + 1;
    for (int c = threadIdx.x; c < k; c += blockDim.x) update_bookkeep<T>(sc, c, k, it1, sc.rr[c]);
    __syncthreads();
    if (threadIdx.x == 0) *sc.it = it1;
}

// blockIdx.y = destination rank.  Writes this rank's entries that rank needs straight into its d vector
// over NVLink; the last block to finish for that destination raises the arrival flag there.
template <typename T>
__global__ void __launch_bounds__(256)
halo_push_kernel(PeerComm *pc, const int *__restrict__ idx, const T *__restrict__ d, const int *n_active) {
    // An exchange is identified by the all-reduce count (peer_wait_halo).  Once every column has converged the
    // kernels of the loop return at once and no all-reduce happens any more: an exchange pushed then would carry
    // the SAME number as the first exchange of the NEXT solve and let that solve's first SpMV run ahead of its
    // halo (seen as a graph-launched tolerance solve stopping one iteration late).  So: nothing to push, nobody
    // waits.  n_active == NULL: the exchange of the initialisation, always done.
    if (n_active && *n_active == 0) return;
    const int p = blockIdx.y;
    const int lo = pc->send_off[p], hi = pc->send_off[p + 1];
    if (hi <= lo) return;
    T *dst = reinterpret_cast<T *>(pc->d_peer[p]) + pc->remote_off[p];
    // four independent index -> value -> remote store chains per thread: with random columns (BASELINE config 5) nearly
    // the whole owned block goes to every peer, and one dependent chain per thread left the NVLink idle
    // (20 MB in ~400 us on 2 GPUs, profiles/r02_bench_c5_n2_before.json)
    const int stride = gridDim.x * blockDim.x;
    int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < hi; i += 4 * stride) {
        int j[4];
        T v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) j[u] = __ldg(idx + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = d[j[u]];
#pragma unroll
        for (int u = 0; u < 4; u++) dst[i + u * stride - lo] = v[u];
    }
    for (; i < hi; i += stride) dst[i - lo] = d[idx[i]];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&pc->push_ticket[p], 1u);
        if (prev == gridDim.x - 1) {
            pc->push_ticket[p] = 0;
            __threadfence_system();
            st_release_sys_u64(pc->halo_flag[p] + pc->rank, pc->seq + 1);
        }
    }
}

// Packs the entries of d that peer devices need (their halo) into one send buffer.
template <typename T>
__global__ void pack_kernel(int count, const int *__restrict__ idx, const T *__restrict__ d, T *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) out[i] = d[idx[i]];
}

// Matrix streams (values, column indices) are read exactly once per SpMV: mark them
// streaming so they do not displace the gathered vector from L1/L2.
template <typename T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
// 4-, 8- or 16-byte streaming load of any trivially copyable bundle.
template <typename P> __device__ __forceinline__ P ld_stream_bytes(const P *p) {
    P out;
    if constexpr (sizeof(P) == 16) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(p));
        memcpy(&out, &v, 16);
    } else if constexpr (sizeof(P) == 8) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(p));
        memcpy(&out, &v, 8);
    } else {
        static_assert(sizeof(P) == 4, "ld_stream_bytes: 4, 8 or 16 bytes");
        const unsigned v = __ldcs(reinterpret_cast<const unsigned *>(p));
        memcpy(&out, &v, 4);
    }
    return out;
}
template <typename P> __device__ __forceinline__ void st_stream_bytes(P *p, const P &val) {
    if constexpr (sizeof(P) == 16) {
        uint4 v;
        memcpy(&v, &val, 16);
        __stcs(reinterpret_cast<uint4 *>(p), v);
    } else if constexpr (sizeof(P) == 8) {
        uint2 v;
        memcpy(&v, &val, 8);
        __stcs(reinterpret_cast<uint2 *>(p), v);
    } else {
        static_assert(sizeof(P) == 4, "st_stream_bytes: 4, 8 or 16 bytes");
        unsigned v;
        memcpy(&v, &val, 4);
        __stcs(reinterpret_cast<unsigned *>(p), v);
    }
}
// Loads / stores of 4, 8 or 16 bytes that carry an L2 eviction policy (createpolicy): the working vectors
// d, q, r of a system that fits the 126 MB L2 are tagged evict-last so that the matrix stream (evict-first)
// and x (streamed) leave them resident from kernel to kernel and from iteration to iteration.
__device__ __forceinline__ unsigned long long l2_policy(bool keep) {
    unsigned long long pol;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <typename P> __device__ __forceinline__ P ld_hint_bytes(const P *p, unsigned long long pol) {
    P out;
    if constexpr (sizeof(P) == 16) {
        unsigned long long a, b;
        asm volatile("ld.global.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(pol));
        unsigned long long v[2] = {a, b};
        memcpy(&out, v, 16);
    } else if constexpr (sizeof(P) == 8) {
        unsigned long long a;
        asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(a) : "l"(p), "l"(pol));
        memcpy(&out, &a, 8);
    } else {
        static_assert(sizeof(P) == 4, "ld_hint_bytes: 4, 8 or 16 bytes");
        unsigned a;
        asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(a) : "l"(p), "l"(pol));
        memcpy(&out, &a, 4);
    }
    return out;
}
template <typename P> __device__ __forceinline__ void st_hint_bytes(P *p, const P &val, unsigned long long pol) {
    if constexpr (sizeof(P) == 16) {
        unsigned long long v[2];
        memcpy(v, &val, 16);
        asm volatile("st.global.L2::cache_hint.v2.b64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(pol) : "memory");
    } else if constexpr (sizeof(P) == 8) {
        unsigned long long a;
        memcpy(&a, &val, 8);
        asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(a), "l"(pol) : "memory");
    } else {
        static_assert(sizeof(P) == 4, "st_hint_bytes: 4, 8 or 16 bytes");
        unsigned a;
        memcpy(&a, &val, 4);
        asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(a), "l"(pol) : "memory");
    }
}
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
        // 8 blocks' partials per round with every load issued before the first add (one L2 round trip per
        // round instead of one per partial); the order of the adds is unchanged: b ascending
        constexpr int UB = 8;
        const int step = blockDim.x / group;
        for (int b0 = t / group; b0 < (int)gridDim.x; b0 += UB * step) {
            T part[UB][V];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int b = b0 + u * step;
#pragma unroll
                for (int v = 0; v < V; v++)
                    part[u][v] = (b < (int)gridDim.x && cp * V + v < k) ? ld_cg(partial + (size_t)b * k + cp * V + v)
                                                                         : Sc<T>::zero();
            }
#pragma unroll
            for (int u = 0; u < UB; u++) {
                if (b0 + u * step < (int)gridDim.x) {
#pragma unroll
                    for (int v = 0; v < V; v++)
                        if (cp * V + v < k) acc[v] = Sc<T>::add(acc[v], part[u][v]);
                }
            }
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
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
#pragma unroll 2
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
// CSR-stream tiles (the schedule of the TMA-fed kernels below).
//
// The non-zeros are cut, at matrix set-up, into tiles of at most RowTileCfg::CAP entries that hold whole rows (a
// row longer than a tile is cut into chunks).  Load balance is per non-zero, not per row, so power-law matrices
// cost the same per entry as stencils.  A long row's chunk sums go to `chunk_sum` and are added up by
// combine_long_rows_kernel; the fused d.q term of such a row uses linearity, d_i * (partial row sum), so it needs
// no second pass.  (The plain-load variant of round 1, spmv_stream_kernel, measured slower than both TMA-fed
// kernels on every config and is gone; it is in git.)
// ---------------------------------------------------------------------------
struct SpmvTile {
    int r0;   // first row
    int r1;   // one past the last row; < 0: chunk of long row r0, result slot -(r1 + 1)
    int p0;   // first non-zero
    int p1;   // one past the last non-zero
};

// ---------------------------------------------------------------------------
// SpMV, one right-hand side, CSR-stream fed by the TMA engine.
//
// Two phases per tile (products to shared memory, then row sums), and the matrix stream never
// passes through registers: one elected thread issues 1-D bulk async copies
// (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier) for the
// tile's values, column indices and row offsets into an S-stage ring, S tiles ahead of
// the one being consumed.  DRAM latency of the A stream is therefore always covered by
// S * ~20 KB in flight per block, independent of occupancy, and the 256 threads spend
// their issue slots on the x gather and the row sums only.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The matrix is read once per SpMV and is (for the large configs) bigger than the 126 MB L2:
// tag its lines evict-first so that they do not push the CG vectors (x, r, d, q -- re-read by the
// three kernels of every iteration) out of L2.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, unsigned bytes, unsigned long long *bar,
                                              unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the thread is parked by the hardware until the phase completes or the hint (in
// nanoseconds) runs out, instead of coming back to the issue port every few hundred cycles.  (In the TMA-fed CG kernels
// 18 % of all executed instructions were polls of this loop, taking issue slots from the warps that had work:
// profiles/r02_notes.md.)
__device__ __forceinline__ bool mbar_try_wait_hint(unsigned long long *bar, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// A lost completion would otherwise hang the GPU: trap instead after ~seconds.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u))
        if (++spins > (1u << 22)) __trap();
}

// The shape of a staged tile, shared by both TMA-fed kernels (the schedule is built once per matrix).
struct RowTileCfg {
    static constexpr int NT = 128;        // max rows per tile (= threads per block of the row-direct kernel)
    static constexpr int TILE = 1024;     // staged non-zeros per tile
    static constexpr int CAP = TILE - 4;  // non-zeros per tile (aligned windows may start 3 entries early)
};

template <typename T, int S> struct TmaCfg {
    static constexpr int NT = 256;                              // threads per block
    static constexpr int TILE = RowTileCfg::TILE;
    static constexpr int NPT = TILE / NT;                       // gathers in flight per thread
    static constexpr size_t VALS_BYTES = (size_t)TILE * sizeof(T);
    static constexpr size_t COLS_BYTES = (size_t)TILE * sizeof(int);
    static constexpr size_t ROWS_BYTES = (size_t)(RowTileCfg::NT + 8) * sizeof(int);
    static constexpr size_t STAGE_BYTES = VALS_BYTES + COLS_BYTES + ROWS_BYTES;
    static constexpr size_t BAR_BYTES = 128;
    static constexpr size_t PROD_BYTES = (size_t)TILE * sizeof(T);
    static constexpr size_t RED_BYTES = (size_t)NT * sizeof(T);
    static constexpr size_t SMEM_BYTES = BAR_BYTES + PROD_BYTES + RED_BYTES + S * STAGE_BYTES;
    static_assert(STAGE_BYTES % 16 == 0 && PROD_BYTES % 16 == 0 && RED_BYTES % 16 == 0, "16-byte aligned stages");
};

// Balanced in both phases, for matrices whose row lengths vary wildly inside a tile (power-law graphs):
//   1. thread t takes non-zeros t, t + 256, ... of the tile whatever rows they belong to: every gather of
//      the tile is in flight at once (ONE L2 round trip per tile instead of one per 8 entries of the
//      longest row), shared-memory reads are conflict-free, products go to shared memory;
//   2. row sums out of shared memory, where a round costs ~30 cycles instead of an L2 round trip: a
//      group of lpr lanes per row, and rows much longer than the rest are summed by a whole warp.
template <typename T, int S, bool DOT>
__global__ void __launch_bounds__(256)          // 4 blocks per SM by registers; forcing 6 (40 registers) measured 1.4x slower
spmv_tma_kernel(int ntiles, int ntiles_interior, const SpmvTile *__restrict__ tiles, const T *__restrict__ vals,
                const int *__restrict__ rowptr, const int *__restrict__ cols, const T *__restrict__ x,
                T *__restrict__ y, T *__restrict__ chunk_sum, CgScalars<T> sc) {
    using K = TmaCfg<T, S>;
    constexpr int VPT = VecW<T>::value, NT = K::NT, NPT = K::NPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);
    T *prod = reinterpret_cast<T *>(smem_raw + K::BAR_BYTES);
    T *red = reinterpret_cast<T *>(smem_raw + K::BAR_BYTES + K::PROD_BYTES);
    unsigned char *stage0 = smem_raw + K::BAR_BYTES + K::PROD_BYTES + K::RED_BYTES;
    const int t = threadIdx.x;
    const int count = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    T dot[1] = {Sc<T>::zero()};
    const unsigned long long stream_policy = l2_evict_first_policy();

    if (t == 0) {
        for (int s = 0; s < S; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int i) {   // thread 0: start the copies of this block's i-th tile
        const SpmvTile tl = tiles[blockIdx.x + (size_t)i * gridDim.x];
        unsigned char *st = stage0 + (size_t)(i % S) * K::STAGE_BYTES;
        unsigned long long *bar = &bars[i % S];
        const int vb = tl.p0 - (tl.p0 % VPT), cb = tl.p0 & ~3;
        const unsigned vbytes = (unsigned)(((tl.p1 - vb + VPT - 1) / VPT) * VPT * sizeof(T));
        const unsigned cbytes = (unsigned)(((tl.p1 - cb + 3) / 4) * 16);
        unsigned rbytes = 0;
        int rb = 0;
        if (tl.r1 >= 0) {
            rb = tl.r0 & ~3;
            rbytes = (unsigned)(((tl.r1 + 1 - rb + 3) / 4) * 16);
        }
        const bool has_nnz = tl.p1 > tl.p0;
        mbar_arrive_expect_tx(bar, (has_nnz ? vbytes + cbytes : 0u) + rbytes);
        if (has_nnz) {
            bulk_g2s_hint(st, vals + vb, vbytes, bar, stream_policy);
            bulk_g2s_hint(st + K::VALS_BYTES, cols + cb, cbytes, bar, stream_policy);
        }
        if (rbytes) bulk_g2s_hint(st + K::VALS_BYTES + K::COLS_BYTES, rowptr + rb, rbytes, bar, stream_policy);
    };

    if (t == 0)
        for (int i = 0; i < S && i < count; i++) issue(i);

    SpmvTile tl_next = {0, 0, 0, 0};
    if (count > 0) tl_next = tiles[blockIdx.x];

    pdl_wait();        // the matrix is not the previous kernel's output; x and the scalars are
    if (sc.pdl_early) pdl_trigger();
    if (DOT) {
        if (*sc.n_active == 0) {
            if (t == 0)
                for (int i = 0; i < S && i < count; i++) mbar_wait(&bars[i], 0u);
            return;
        }
    }

    int trace_it = -1;
    if (DOT && sc.trace) {
        trace_it = *sc.it;
        if (blockIdx.x == 0 && t == 0) trace_mark<T>(sc, trace_it, TR_SPMV_START);
    }
    // row-block shards (as spmv_tma_rows_kernel): tiles [ntiles_interior, ntiles) gather entries that peers push into
    // this GPU's vector; a block waits for them when it reaches its first such tile.  With the random columns this
    // kernel is chosen for that is the first tile -- the matrix data of S tiles is in flight by then.
    bool halo_ready = !(sc.peer && sc.peer->world > 1);

    for (int i = 0; i < count; i++) {
        if (!halo_ready && (int)blockIdx.x + i * (int)gridDim.x >= ntiles_interior) {
            if (t == 0) {
                peer_wait_halo(sc.peer);
                if ((int)blockIdx.x == ntiles_interior % (int)gridDim.x) trace_mark<T>(sc, trace_it, TR_HALO_READY);
            }
            __syncthreads();
            halo_ready = true;
        }
        const SpmvTile tl = tl_next;
        if (i + 1 < count) tl_next = tiles[blockIdx.x + (size_t)(i + 1) * gridDim.x];
        const int s = i % S;
        const unsigned char *st = stage0 + (size_t)s * K::STAGE_BYTES;
        const T *vals_s = reinterpret_cast<const T *>(st);
        const int *cols_s = reinterpret_cast<const int *>(st + K::VALS_BYTES);
        const int *rp_s = reinterpret_cast<const int *>(st + K::VALS_BYTES + K::COLS_BYTES);
        const int vb = tl.p0 - (tl.p0 % VPT), cb = tl.p0 & ~3;

        // who sums which row in phase 2 (known from the descriptor alone, so x[row] for the fused dot can be
        // requested before the stage has even arrived)
        const int rows = tl.r1 >= 0 ? tl.r1 - tl.r0 : 0;
        int lpr = 1;
        while (lpr < 32 && rows * lpr * 2 <= NT) lpr *= 2;
        const int rr = t / lpr, lane = t % lpr;
        const bool valid = rr < rows;
        T xr = Sc<T>::zero();
        if (DOT && valid && lane == 0) xr = __ldg(x + tl.r0 + rr);

        mbar_wait(&bars[s], (unsigned)((i / S) & 1));

        // ---- 1. products: every gather of the tile issued before the first multiply
        int cc[NPT];
#pragma unroll
        for (int u = 0; u < NPT; u++) {
            const int j = tl.p0 + t + u * NT;
            cc[u] = (j < tl.p1) ? cols_s[j - cb] : -1;
        }
        T xv[NPT];
#pragma unroll
        for (int u = 0; u < NPT; u++)
            if (cc[u] >= 0) xv[u] = __ldg(x + cc[u]);
#pragma unroll
        for (int u = 0; u < NPT; u++) {
            const int j = tl.p0 + t + u * NT;
            if (cc[u] >= 0) prod[j - tl.p0] = Sc<T>::mul(vals_s[j - vb], xv[u]);
        }
        __syncthreads();

        if (tl.r1 >= 0) {
            // ---- 2. row sums
            const int rb = tl.r0 & ~3;
            int lo = 0, hi = 0;
            if (valid) {
                lo = rp_s[tl.r0 + rr - rb] - tl.p0;
                hi = rp_s[tl.r0 + rr + 1 - rb] - tl.p0;
            }
            const bool deferred = valid && lpr < 32 && (hi - lo) > 16 * lpr;
            T sum = Sc<T>::zero();
            if (valid && !deferred)
                for (int j = lo + lane; j < hi; j += lpr) sum = Sc<T>::add(sum, prod[j]);
            for (int off = lpr >> 1; off > 0; off >>= 1) {
                if constexpr (Sc<T>::cplx) {
                    sum.x += __shfl_xor_sync(0xffffffffu, sum.x, off);
                    sum.y += __shfl_xor_sync(0xffffffffu, sum.y, off);
                } else {
                    sum += __shfl_xor_sync(0xffffffffu, sum, off);
                }
            }
            unsigned pending = __ballot_sync(0xffffffffu, deferred && lane == 0);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const int rlo = __shfl_sync(0xffffffffu, lo, src), rhi = __shfl_sync(0xffffffffu, hi, src);
                T part = Sc<T>::zero();
                for (int j = rlo + (t & 31); j < rhi; j += 32) part = Sc<T>::add(part, prod[j]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    if constexpr (Sc<T>::cplx) {
                        part.x += __shfl_xor_sync(0xffffffffu, part.x, off);
                        part.y += __shfl_xor_sync(0xffffffffu, part.y, off);
                    } else {
                        part += __shfl_xor_sync(0xffffffffu, part, off);
                    }
                }
                if ((t & 31) == src) sum = part;
            }
            if (valid && lane == 0) {
                y[tl.r0 + rr] = sum;
                if (DOT) dot[0] = Sc<T>::fma(xr, sum, dot[0]);
            }
        } else {
            // ---- 2'. chunk of a long row: one sum for the whole tile
            T part[1] = {Sc<T>::zero()};
            for (int j = t; j < tl.p1 - tl.p0; j += NT) part[0] = Sc<T>::add(part[0], prod[j]);
            block_col_reduce<T, 1>(part, 1, red);
            if (t == 0) {
                chunk_sum[-(tl.r1 + 1)] = red[0];
                if (DOT) dot[0] = Sc<T>::fma(__ldg(x + tl.r0), red[0], dot[0]);
            }
        }
        __syncthreads();   // the stage and `prod` are free again
        if (t == 0 && i + S < count) issue(i + S);
    }

    if (DOT) {
        block_col_reduce<T, 1>(dot, 1, red);
        if (publish_and_arrive<T, 1>(red, 1, 1, sc.partial, sc.ticket + TK_SPMV)) {
            if (t == 0) trace_mark<T>(sc, trace_it, TR_SPMV_ALL_DONE);
            grid_col_reduce<T, 1>(sc.partial, 1, 1, 1, red);
            T total = red[0];
            if (sc.peer) total = peer_allreduce<T>(sc.peer, red[0]);    // sum over the GPUs, inside this kernel
            if (t == 0) {
                sc.dq[0] = total;
                sc.ticket[TK_SPMV] = 0;
                trace_mark<T>(sc, trace_it, TR_SPMV_END);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// SpMV, one right-hand side, "row-direct" CSR-stream fed by TMA.
//
// Like spmv_tma_kernel the tile's values / column indices / row offsets arrive in shared
// memory by bulk async copies, S tiles ahead.  Unlike it there is no product buffer and
// no second phase: a tile holds at most RT::NT rows, every row is owned by a group of
// lpr lanes (1 for short rows, up to 32 when the tile holds few long rows) that walks the
// row straight out of the staged arrays, gathers x and accumulates in registers.  One
// block-wide barrier per tile (to recycle the stage), all shared memory spent on stages,
// i.e. on bytes in flight.
// ---------------------------------------------------------------------------
template <typename T, int S> struct RowTmaCfg {
    static constexpr size_t VALS_BYTES = (size_t)RowTileCfg::TILE * sizeof(T);
    static constexpr size_t COLS_BYTES = (size_t)RowTileCfg::TILE * sizeof(int);
    static constexpr size_t ROWS_BYTES = (size_t)(RowTileCfg::NT + 8) * sizeof(int);
    static constexpr size_t STAGE_BYTES = VALS_BYTES + COLS_BYTES + ROWS_BYTES;
    static constexpr size_t BAR_BYTES = 128;
    static constexpr size_t RED_BYTES = (size_t)RowTileCfg::NT * sizeof(T);
    static constexpr size_t SMEM_BYTES = BAR_BYTES + RED_BYTES + S * STAGE_BYTES;
    static_assert(STAGE_BYTES % 16 == 0 && RED_BYTES % 16 == 0, "16-byte aligned stages");
};

template <typename T, int S, bool DOT>
__global__ void __launch_bounds__(RowTileCfg::NT)
spmv_tma_rows_kernel(int ntiles, int ntiles_interior, int defer_len, const SpmvTile *__restrict__ tiles, const T *__restrict__ vals,
                     const int *__restrict__ rowptr, const int *__restrict__ cols, const T *__restrict__ x,
                     T *__restrict__ y, T *__restrict__ chunk_sum, CgScalars<T> sc) {
    using K = RowTmaCfg<T, S>;
    constexpr int VPT = VecW<T>::value, NT = RowTileCfg::NT;
    constexpr int UB = 8;                             // gathers in flight per thread
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);
    T *red = reinterpret_cast<T *>(smem_raw + K::BAR_BYTES);
    unsigned char *stage0 = smem_raw + K::BAR_BYTES + K::RED_BYTES;
    const int t = threadIdx.x;
    const int count = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    T dot[1] = {Sc<T>::zero()};
    const unsigned long long stream_policy = l2_evict_first_policy();
    const unsigned long long keep = l2_policy(sc.l2_keep != 0);

    if (t == 0) {
        for (int s = 0; s < S; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](const SpmvTile &tl, int i) {   // thread 0: start the copies of this block's i-th tile
        unsigned char *st = stage0 + (size_t)(i % S) * K::STAGE_BYTES;
        unsigned long long *bar = &bars[i % S];
        const int vb = tl.p0 - (tl.p0 % VPT), cb = tl.p0 & ~3;
        const unsigned vbytes = (unsigned)(((tl.p1 - vb + VPT - 1) / VPT) * VPT * sizeof(T));
        const unsigned cbytes = (unsigned)(((tl.p1 - cb + 3) / 4) * 16);
        unsigned rbytes = 0;
        int rb = 0;
        if (tl.r1 >= 0) {
            rb = tl.r0 & ~3;
            rbytes = (unsigned)(((tl.r1 + 1 - rb + 3) / 4) * 16);
        }
        const bool has_nnz = tl.p1 > tl.p0;
        mbar_arrive_expect_tx(bar, (has_nnz ? vbytes + cbytes : 0u) + rbytes);
        if (has_nnz) {
            bulk_g2s_hint(st, vals + vb, vbytes, bar, stream_policy);
            bulk_g2s_hint(st + K::VALS_BYTES, cols + cb, cbytes, bar, stream_policy);
        }
        if (rbytes) bulk_g2s_hint(st + K::VALS_BYTES + K::COLS_BYTES, rowptr + rb, rbytes, bar, stream_policy);
    };
    auto tile_of = [&](int i) { return tiles[blockIdx.x + (size_t)i * gridDim.x]; };

    // thread 0 keeps the descriptor of the next tile to issue in a register, fetched one tile early
    SpmvTile tl_issue = {0, 0, 0, 0};
    if (t == 0) {
        for (int i = 0; i < S && i < count; i++) issue(tile_of(i), i);
        if (S < count) tl_issue = tile_of(S);
    }
    SpmvTile tl_next = {0, 0, 0, 0};
    if (count > 0) tl_next = tile_of(0);

    // Everything above touched the matrix only.  The vectors and scalars are the previous kernel's output.
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (DOT) {
        if (*sc.n_active == 0) {
            // nothing to do, but the copies already in flight must land before this block's shared memory is released
            if (t == 0)
                for (int i = 0; i < S && i < count; i++) mbar_wait(&bars[i], 0u);
            return;
        }
    }
    int trace_it = -1;
    if (DOT && sc.trace) {
        trace_it = *sc.it;
        if (blockIdx.x == 0 && t == 0) trace_mark<T>(sc, trace_it, TR_SPMV_START);
    }

    // row-block shards: tiles [ntiles_interior, ntiles) gather entries that peers push into this GPU's vector;
    // a block waits for them only when it reaches its first such tile (its matrix data is already in flight)
    bool halo_ready = !(sc.peer && sc.peer->world > 1);

    for (int i = 0; i < count; i++) {
        if (!halo_ready && (int)blockIdx.x + i * (int)gridDim.x >= ntiles_interior) {
            if (t == 0) {
                peer_wait_halo(sc.peer);
                if ((int)blockIdx.x == ntiles_interior % (int)gridDim.x) trace_mark<T>(sc, trace_it, TR_HALO_READY);
            }
            __syncthreads();
            halo_ready = true;
        }
        const SpmvTile tl = tl_next;
        if (i + 1 < count) tl_next = tile_of(i + 1);
        const int s = i % S;
        const unsigned char *st = stage0 + (size_t)s * K::STAGE_BYTES;
        const T *vals_s = reinterpret_cast<const T *>(st);
        const int *cols_s = reinterpret_cast<const int *>(st + K::VALS_BYTES);
        const int *rp_s = reinterpret_cast<const int *>(st + K::VALS_BYTES + K::COLS_BYTES);
        const int vb = tl.p0 - (tl.p0 % VPT), cb = tl.p0 & ~3;
        mbar_wait(&bars[s], (unsigned)((i / S) & 1));

        if (tl.r1 >= 0) {
            const int rows = tl.r1 - tl.r0;          // <= NT
            const int rb = tl.r0 & ~3;
            int lpr = 1;
            while (lpr < 32 && rows * lpr * 2 <= NT) lpr *= 2;
            const int rr = t / lpr, lane = t % lpr;
            const bool valid = rr < rows;
            T sum = Sc<T>::zero();
            T xr = Sc<T>::zero();
            int lo = 0, hi = 0;
            if (valid) {
                if (DOT && lane == 0) xr = __ldg(x + tl.r0 + rr);
                lo = rp_s[tl.r0 + rr - rb];
                hi = rp_s[tl.r0 + rr + 1 - rb];
            }
            // A row much longer than its neighbours (power-law matrices: a 300-entry row among 3-entry rows)
            // would keep its lpr lanes walking it while the rest of the block idles.  Such a row is left
            // out here and walked by the whole warp below.
            const bool deferred = valid && lpr < 32 && defer_len > 0 && (hi - lo) > defer_len * lpr;
            if (valid && !deferred) {
                // Batches of UB non-zeros with every load issued before the first FMA: a warp issues in
                // order, so a plain loop would serialise one L2 round trip per non-zero.  Lanes past the
                // end of the row re-read the batch's first entry (a valid address) with a zero coefficient.
                for (int j0 = lo + lane; j0 < hi; j0 += UB * lpr) {
                    T av[UB], xv[UB];
#pragma unroll
                    for (int u = 0; u < UB; u++) {
                        const int j = j0 + u * lpr;
                        const bool ok = j < hi;
                        const int jj = ok ? j : j0;
                        xv[u] = __ldg(x + cols_s[jj - cb]);
                        av[u] = ok ? vals_s[jj - vb] : Sc<T>::zero();
                    }
#pragma unroll
                    for (int u = 0; u < UB; u++) sum = Sc<T>::fma(av[u], xv[u], sum);
                }
            }
            for (int off = lpr >> 1; off > 0; off >>= 1) {
                if constexpr (Sc<T>::cplx) {
                    sum.x += __shfl_xor_sync(0xffffffffu, sum.x, off);
                    sum.y += __shfl_xor_sync(0xffffffffu, sum.y, off);
                } else {
                    sum += __shfl_xor_sync(0xffffffffu, sum, off);
                }
            }
            // the deferred rows of this warp, one after the other, 32 lanes x UB gathers in flight each
            unsigned pending = __ballot_sync(0xffffffffu, deferred && lane == 0);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const int rlo = __shfl_sync(0xffffffffu, lo, src), rhi = __shfl_sync(0xffffffffu, hi, src);
                const int wl = t & 31;
                T part = Sc<T>::zero();
                for (int j0 = rlo + wl; j0 < rhi; j0 += UB * 32) {
                    T av[UB], xv[UB];
#pragma unroll
                    for (int u = 0; u < UB; u++) {
                        const int j = j0 + u * 32;
                        const bool ok = j < rhi;
                        const int jj = ok ? j : j0;
                        xv[u] = __ldg(x + cols_s[jj - cb]);
                        av[u] = ok ? vals_s[jj - vb] : Sc<T>::zero();
                    }
#pragma unroll
                    for (int u = 0; u < UB; u++) part = Sc<T>::fma(av[u], xv[u], part);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    if constexpr (Sc<T>::cplx) {
                        part.x += __shfl_xor_sync(0xffffffffu, part.x, off);
                        part.y += __shfl_xor_sync(0xffffffffu, part.y, off);
                    } else {
                        part += __shfl_xor_sync(0xffffffffu, part, off);
                    }
                }
                if (wl == src) sum = part;
            }
            if (valid && lane == 0) {
                st_hint_bytes(y + tl.r0 + rr, sum, keep);
                if (DOT) dot[0] = Sc<T>::fma(xr, sum, dot[0]);
            }
        } else {
            // chunk of a long row: the whole block strides over it, one sum for the tile
            T part[1] = {Sc<T>::zero()};
            for (int j0 = tl.p0 + t; j0 < tl.p1; j0 += UB * NT) {
                T av[UB], xv[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const int j = j0 + u * NT;
                    const bool ok = j < tl.p1;
                    const int jj = ok ? j : j0;
                    xv[u] = __ldg(x + cols_s[jj - cb]);
                    av[u] = ok ? vals_s[jj - vb] : Sc<T>::zero();
                }
#pragma unroll
                for (int u = 0; u < UB; u++) part[0] = Sc<T>::fma(av[u], xv[u], part[0]);
            }
            block_col_reduce<T, 1>(part, 1, red);
            if (t == 0) {
                chunk_sum[-(tl.r1 + 1)] = red[0];
                if (DOT) dot[0] = Sc<T>::fma(__ldg(x + tl.r0), red[0], dot[0]);
            }
        }
        __syncthreads();   // every thread is done with stage s
        if (t == 0 && i + S < count) {
            issue(tl_issue, i + S);
            if (i + S + 1 < count) tl_issue = tile_of(i + S + 1);
        }
    }

    if (DOT) {
        block_col_reduce<T, 1>(dot, 1, red);
        if (publish_and_arrive<T, 1>(red, 1, 1, sc.partial, sc.ticket + TK_SPMV)) {
            if (t == 0) trace_mark<T>(sc, trace_it, TR_SPMV_ALL_DONE);
            grid_col_reduce<T, 1>(sc.partial, 1, 1, 1, red);
            T total = red[0];
            if (sc.peer) total = peer_allreduce<T>(sc.peer, red[0]);    // sum over the GPUs, inside this kernel
            if (t == 0) {
                sc.dq[0] = total;
                sc.ticket[TK_SPMV] = 0;
                trace_mark<T>(sc, trace_it, TR_SPMV_END);
            }
        }
    }
}

// y[row] = sum of the chunk sums of a long row (fixed order).  One thread per long row.
struct LongRow {
    int row, slot0, nslots;
};
template <typename T>
__global__ void combine_long_rows_kernel(int nlong, const LongRow *__restrict__ rows,
                                         const T *__restrict__ chunk_sum, T *__restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlong) return;
    const LongRow lr = rows[i];
    T s = Sc<T>::zero();
    for (int c = 0; c < lr.nslots; c++) s = Sc<T>::add(s, chunk_sum[lr.slot0 + c]);
    y[lr.row] = s;
}

// ---------------------------------------------------------------------------
// Row-pattern dictionary ("pattern CSR"), k = 1.
//
// Matrices assembled on uniform grids with constant coefficients -- the Poisson / Laplace configs, the
// Helmholtz FE operators the reference's drivers build with a constant wave number (local_rect, helm_fe),
// i.e. the subdomain matrices as_prec really solves with -- consist of a handful of distinct ROWS when a
// row is written as the list of (column - row, value) pairs: interior, faces, edges, corners.  At set-up
// the rows are hashed on the device, the distinct patterns are collected into a small table, and every
// row keeps only a 16-bit pattern number.  The SpMV then moves 2 bytes per ROW instead of 12..20 bytes
// per NON-ZERO: what is left is one read of x and one write of y.  The arithmetic is the CSR kernel's:
// the same products added in the same (CSR) order with the same FMAs.
// The CSR arrays stay resident (k > 1, irregular rows and matrices with too many patterns use them).
// ---------------------------------------------------------------------------
constexpr int PAT_MAXLEN = 32;          // longest row a pattern may have
constexpr int PAT_MAXCOUNT = 4096;      // most distinct patterns
constexpr int PAT_TABLE_SLOTS = 1 << 15;
constexpr unsigned long long PAT_EMPTY = 0ull;

struct PatSlot {
    unsigned long long hash;   // PAT_EMPTY: free
    int row;                   // representative row
    int id;                    // pattern number
};
struct PatBuild {
    int count;                 // distinct patterns so far
    int fail;                  // 1: too many patterns / a row too long / a hash collision -> stay with CSR
};

template <typename T> __device__ __forceinline__ bool pat_same_bits(const T &a, const T &b) {
    unsigned long long x[2] = {0, 0}, y[2] = {0, 0};
    memcpy(x, &a, sizeof(T));
    memcpy(y, &b, sizeof(T));
    return x[0] == y[0] && x[1] == y[1];
}
__device__ __forceinline__ unsigned long long pat_mix(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xD6E8FEB86659FD93ull;
    return h ^ (h >> 32);
}

template <typename T>
__device__ __forceinline__ unsigned long long pat_hash_row(int row, int lo, int hi, const T *vals, const int *cols) {
    unsigned long long h = pat_mix(0x5bd1e995ull, (unsigned long long)(hi - lo));
    for (int j = lo; j < hi; j++) {
        h = pat_mix(h, (unsigned long long)(long long)(cols[j] - row));
        unsigned long long bits[2] = {0, 0};
        memcpy(bits, vals + j, sizeof(T));
        h = pat_mix(h, bits[0]);
        if (sizeof(T) > 8) h = pat_mix(h, bits[1]);
    }
    return h == PAT_EMPTY ? 1ull : h;
}

// pass 1: claim a table slot for every distinct hash
template <typename T>
__global__ void pat_insert_kernel(int n, const T *__restrict__ vals, const int *__restrict__ rowptr,
                                  const int *__restrict__ cols, PatSlot *table, PatBuild *pb) {
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
        const int lo = rowptr[row], hi = rowptr[row + 1];
        if (hi - lo > PAT_MAXLEN) {
            pb->fail = 1;
            return;
        }
        if (*reinterpret_cast<volatile int *>(&pb->fail)) return;
        const unsigned long long h = pat_hash_row<T>(row, lo, hi, vals, cols);
        unsigned slot = (unsigned)(h >> 17) & (PAT_TABLE_SLOTS - 1);
        for (int probe = 0; probe < PAT_TABLE_SLOTS; probe++) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&table[slot].hash);   // fast path: no atomic
            if (cur == PAT_EMPTY) cur = atomicCAS(&table[slot].hash, PAT_EMPTY, h);
            if (cur == PAT_EMPTY) {            // this thread created the pattern
                const int id = atomicAdd(&pb->count, 1);
                table[slot].row = row;
                table[slot].id = id;
                if (id >= PAT_MAXCOUNT) pb->fail = 1;
                break;
            }
            if (cur == h) break;
            slot = (slot + 1) & (PAT_TABLE_SLOTS - 1);
            if (*reinterpret_cast<volatile int *>(&pb->fail)) break;
        }
    }
}

// pass 2: every row looks its pattern up, checks entry by entry that it really IS the representative's
// pattern (a 64-bit hash collision must not merge two different rows), and keeps the number
template <typename T>
__global__ void pat_assign_kernel(int n, const T *__restrict__ vals, const int *__restrict__ rowptr,
                                  const int *__restrict__ cols, const PatSlot *__restrict__ table, PatBuild *pb,
                                  unsigned short *__restrict__ pat) {
    if (pb->fail) return;
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
        const unsigned long long h = pat_hash_row<T>(row, rowptr[row], rowptr[row + 1], vals, cols);
        unsigned slot = (unsigned)(h >> 17) & (PAT_TABLE_SLOTS - 1);
        while (table[slot].hash != h) slot = (slot + 1) & (PAT_TABLE_SLOTS - 1);
        const int rep = table[slot].row;
        const int lo = rowptr[row], hi = rowptr[row + 1], rlo = rowptr[rep];
        bool same = (hi - lo) == (rowptr[rep + 1] - rlo);
        for (int j = 0; same && j < hi - lo; j++) {
            same = (cols[lo + j] - row) == (cols[rlo + j] - rep) && pat_same_bits<T>(vals[lo + j], vals[rlo + j]);
        }
        if (!same) pb->fail = 1;
        pat[row] = (unsigned short)table[slot].id;
    }
}

// pass 3: the pattern table itself, [count][PAT_MAXLEN] offsets and values (+ lengths), from the representatives
template <typename T>
__global__ void pat_table_kernel(const T *__restrict__ vals, const int *__restrict__ rowptr, const int *__restrict__ cols,
                                 const PatSlot *__restrict__ table, const PatBuild *pb, int *__restrict__ p_len,
                                 int *__restrict__ p_off, T *__restrict__ p_val) {
    if (pb->fail) return;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < PAT_TABLE_SLOTS; slot += gridDim.x * blockDim.x) {
        if (table[slot].hash == PAT_EMPTY) continue;
        const int id = table[slot].id, rep = table[slot].row;
        const int lo = rowptr[rep], len = rowptr[rep + 1] - lo;
        p_len[id] = len;
        for (int j = 0; j < PAT_MAXLEN; j++) {
            p_off[id * PAT_MAXLEN + j] = j < len ? cols[lo + j] - rep : 0;
            p_val[id * PAT_MAXLEN + j] = j < len ? vals[lo + j] : Sc<T>::zero();
        }
    }
}

// y = A x from the pattern dictionary, fused with the partial sums of x.y.
// A block takes chunks of PAT_CHUNK consecutive rows (chunk c of the schedule `chunks`, or c itself when the
// schedule is NULL): consecutive threads take consecutive rows, so each of a row's gathers x[row + offset] is
// one coalesced 256-byte load per warp, and the +-NX neighbours of a chunk's rows are the chunk's own rows
// a little earlier / later (L1 hits).  Row-block shards list the chunks that touch halo columns last
// ([nchunks_interior, nchunks)) and wait for the peers' entries only there.
constexpr int PAT_THREADS = 256;
constexpr int PAT_CHUNK = 1024;

// PEER: the row-block sharded flavour (halo wait, all-reduce of d.q through peer memory); the single-GPU
// engine runs the PEER = false instantiation, which contains none of that code.
// STRIDE: entries per pattern in the shared-memory copy of the table (8, 16 or 32 >= the longest row).
template <typename T, bool DOT, int STRIDE, bool PEER>
__global__ void __launch_bounds__(PAT_THREADS)
spmv_pattern_kernel(int n, int nchunks, int nchunks_interior, const int *__restrict__ chunks, int npat,
                    const unsigned short *__restrict__ pat, const int *__restrict__ p_len, const int *__restrict__ p_off,
                    const T *__restrict__ p_val, const T *__restrict__ x, T *__restrict__ y, CgScalars<T> sc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *red = reinterpret_cast<T *>(smem_raw);                                  // [PAT_THREADS]
    T *s_val = red + PAT_THREADS;                                              // [npat][STRIDE]
    int *s_off = reinterpret_cast<int *>(s_val + npat * STRIDE);               // [npat][STRIDE]
    int *s_len = s_off + npat * STRIDE;                                        // [npat]
    const int t = threadIdx.x;
    // the table is part of the matrix, not the previous kernel's output: staged before the grid dependency wait
    for (int i = t; i < npat * STRIDE; i += PAT_THREADS) {
        const int id = i / STRIDE, j = i % STRIDE;
        s_val[i] = p_val[id * PAT_MAXLEN + j];
        s_off[i] = p_off[id * PAT_MAXLEN + j];
    }
    for (int i = t; i < npat; i += PAT_THREADS) s_len[i] = p_len[i];
    __syncthreads();
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (DOT) {
        if (*sc.n_active == 0) return;
    }
    // (the iteration number is re-read at every mark: *sc.it only changes in the x/r update)
    if (DOT && sc.trace && blockIdx.x == 0 && t == 0) trace_mark<T>(sc, *sc.it, TR_SPMV_START);
    const unsigned long long keep = l2_policy(sc.l2_keep != 0);
    T dot[1] = {Sc<T>::zero()};
    bool halo_ready = true;
    if constexpr (PEER) halo_ready = !(sc.peer && sc.peer->world > 1);

    // Chunks are dealt round-robin: at any moment the grid works on one compact window of ~gridDim chunks, so
    // the +-NX*NY neighbours of a row are being read by other blocks at the same time and come from L2.  (With a
    // contiguous share per block the blocks sit far apart in the vector and every neighbour plane is fetched
    // from HBM again: measured 2.6x the DRAM reads on the 300^3 Laplacian.)
    for (int ci = (int)blockIdx.x; ci < nchunks; ci += (int)gridDim.x) {
        if constexpr (PEER) {
            if (!halo_ready && ci >= nchunks_interior) {
                if (t == 0) {
                    peer_wait_halo(sc.peer);
                    if (DOT && sc.trace) trace_mark<T>(sc, *sc.it, TR_HALO_READY);
                }
                __syncthreads();
                halo_ready = true;
            }
        }
        const int chunk = chunks ? chunks[ci] : ci;
        const int row0 = chunk * PAT_CHUNK + t, row_end = min(n, (chunk + 1) * PAT_CHUNK);
        // A row costs two dependent trips to memory (its pattern number, then its gathers), and a thread has
        // little else to do: the kernel is bound by latency x bytes in flight.  So the pattern numbers of the
        // thread's PAT_STEPS rows of this chunk are requested up front, and two rows are gathered at a time.
        constexpr int PAT_STEPS = PAT_CHUNK / PAT_THREADS;
        int ids[PAT_STEPS];
#pragma unroll
        for (int s = 0; s < PAT_STEPS; s++) {
            const int row = row0 + s * PAT_THREADS;
            ids[s] = row < row_end ? (int)pat[row] : -1;
        }
#pragma unroll
        for (int s = 0; s < PAT_STEPS; s += 2) {
            const int rowa = row0 + s * PAT_THREADS, rowb = rowa + PAT_THREADS;
            const bool oka = ids[s] >= 0, okb = ids[s + 1] >= 0;
            if (!oka) break;                      // past the end of the (last) chunk; row b is even further
            // a missing row b repeats row a (valid addresses) and is not stored
            const int ida = ids[s], idb = okb ? ids[s + 1] : ida;
            const T *xa = x + rowa, *xb = okb ? x + rowb : xa;
            const int lena = s_len[ida], lenb = s_len[idb];
            const T *pva = s_val + ida * STRIDE, *pvb = s_val + idb * STRIDE;
            const int *poa = s_off + ida * STRIDE, *pob = s_off + idb * STRIDE;
            T suma = Sc<T>::zero(), sumb = Sc<T>::zero();
            // padded entries have offset 0 and coefficient 0: a batch needs no bounds test
#pragma unroll
            for (int j0 = 0; j0 < STRIDE; j0 += 8) {
                if (j0 < lena || j0 < lenb) {
                    T xva[8], xvb[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        xva[u] = __ldg(xa + poa[j0 + u]);
                        xvb[u] = __ldg(xb + pob[j0 + u]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        suma = Sc<T>::fma(pva[j0 + u], xva[u], suma);
                        sumb = Sc<T>::fma(pvb[j0 + u], xvb[u], sumb);
                    }
                }
            }
            if (oka) {
                st_hint_bytes(y + rowa, suma, keep);
                if (DOT) dot[0] = Sc<T>::fma(__ldg(xa), suma, dot[0]);
            }
            if (okb) {
                st_hint_bytes(y + rowb, sumb, keep);
                if (DOT) dot[0] = Sc<T>::fma(__ldg(xb), sumb, dot[0]);
            }
        }
    }

    if (DOT) {
        block_col_reduce<T, 1>(dot, 1, red);
        if (publish_and_arrive<T, 1>(red, 1, 1, sc.partial, sc.ticket + TK_SPMV)) {
            if (t == 0 && sc.trace) trace_mark<T>(sc, *sc.it, TR_SPMV_ALL_DONE);
            grid_col_reduce<T, 1>(sc.partial, 1, 1, 1, red);
            T total = red[0];
            if constexpr (PEER) {
                if (sc.peer) total = peer_allreduce<T>(sc.peer, red[0]);
            }
            if (t == 0) {
                sc.dq[0] = total;
                sc.ticket[TK_SPMV] = 0;
                if (sc.trace) trace_mark<T>(sc, *sc.it, TR_SPMV_END);
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
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (DOT) {
        if (*sc.n_active == 0) return;
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
            // (a contiguous row range per block, or batching the gathers ahead of the FMAs, both measured
            //  slower here: with k-wide rows of x the kernel is bound by L2 -> SM traffic)
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
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        T total0 = smem[0];
        if (sc.peer) total0 = peer_allreduce<T>(sc.peer, smem[0]);     // k == 1: sum over the GPUs
        if (t < kv) {
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int c = t * V + v;
                if (c < k) {
                    if (sc.defer) sc.rr[c] = smem[t * V + v];
                    else init_bookkeep<T>(sc, c, sc.peer ? total0 : smem[t * V + v]);
                }
            }
        }
        __syncthreads();
        if (t == 0) {
            if (!sc.defer) {
                int live = 0;
                for (int c = 0; c < k; c++) live += (sc.state[c] == ST_ACTIVE);
                *sc.n_active = live;
                *sc.it = 0;
            }
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
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (*sc.n_active == 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    if (sc.trace && blockIdx.x == 0 && t == 0) trace_mark<T>(sc, *sc.it, TR_XR_START);
    const unsigned long long keep = l2_policy(sc.l2_keep != 0);
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
        const P dv = ld_hint_bytes(reinterpret_cast<const P *>(d) + p, keep);
        const P qv = ld_hint_bytes(reinterpret_cast<const P *>(q) + p, keep);
        // x is touched by this kernel only, once per iteration: streamed (evict-first) in both directions so
        // that it never displaces r, d and q -- which three kernels share -- from the 126 MB L2
        P xv = ld_stream_bytes(reinterpret_cast<const P *>(x) + p);
        P rv = ld_hint_bytes(reinterpret_cast<const P *>(r) + p, keep);
#pragma unroll
        for (int v = 0; v < V; v++) {
            xv.v[v] = Sc<T>::fma(alpha[v], dv.v[v], xv.v[v]);
            rv.v[v] = Sc<T>::fnma(alpha[v], qv.v[v], rv.v[v]);
            acc[v] = Sc<T>::fma(rv.v[v], rv.v[v], acc[v]);
        }
        st_stream_bytes(reinterpret_cast<P *>(x) + p, xv);
        st_hint_bytes(reinterpret_cast<P *>(r) + p, rv, keep);
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
        const int it1 = *sc.it + 1;
        if (t == 0) trace_mark<T>(sc, it1 - 1, TR_XR_ALL_DONE);
        grid_col_reduce<T, V>(sc.partial, kv, kv, k, smem);
        T total0 = smem[0];
        if (sc.peer) total0 = peer_allreduce<T>(sc.peer, smem[0]);     // k == 1: sum over the GPUs
        if (t < kv) {
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int c = t * V + v;
                if (c < k) {
                    if (sc.defer) sc.rr[c] = smem[t * V + v];
                    else update_bookkeep<T>(sc, c, k, it1, sc.peer ? total0 : smem[t * V + v]);
                }
            }
        }
        __syncthreads();
        if (t == 0) {
            if (!sc.defer) *sc.it = it1;
            sc.ticket[TK_UPDATE] = 0;
            trace_mark<T>(sc, it1 - 1, TR_XR_END);
        }
    }
}

// beta = delta_new / delta_old ; d = beta d + r        clcg.c:389-415 (aypx.cl)
template <typename T, int V>
__global__ void __launch_bounds__(256)
update_d_kernel(size_t npacks, size_t nelem, int k, int kv, const T *__restrict__ r,
                T *__restrict__ d, CgScalars<T> sc) {
    pdl_wait();
    if (sc.pdl_early) pdl_trigger();
    if (*sc.n_active == 0) return;
    using P = Pack<T, V>;
    const int t = threadIdx.x;
    if (sc.trace && blockIdx.x == 0 && t == 0) trace_mark<T>(sc, *sc.it - 1, TR_D_START);
    const unsigned long long keep = l2_policy(sc.l2_keep != 0);
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
        const P rv = ld_hint_bytes(reinterpret_cast<const P *>(r) + p, keep);
        P dv = ld_hint_bytes(reinterpret_cast<const P *>(d) + p, keep);
#pragma unroll
        for (int v = 0; v < V; v++) dv.v[v] = Sc<T>::fma(beta[v], dv.v[v], rv.v[v]);
        st_hint_bytes(reinterpret_cast<P *>(d) + p, dv, keep);
    }
    if (V > 1 && blockIdx.x == 0) {
        const size_t e = npacks * V + t;
        if (e < nelem) d[e] = Sc<T>::fma(beta[0], d[e], r[e]);
    }
}

// ---------------------------------------------------------------------------
// The whole solve in ONE cooperative launch -- the schedule for systems whose iteration fits
// the L2 (config C1, and the subdomain solves the reference's as_prec actually makes:
// n ~ 16 k, k = M_s^2 right-hand sides, 256 fixed iterations).  There the three-kernel
// iteration is bound by launch and reduction latency (~15 us), not by bytes.
//
// One block of 1024 threads per SM, two grid-wide barriers per iteration:
//
//   phase 1   q = A (r + beta d)       the direction update d = r + beta d (aypx.cl) is NOT a
//             partial d.q               separate pass: the gather recomputes it on the fly
//   -- grid.sync; every block sums the per-block partials in the same order -> alpha
//   phase 2   dn = r + beta d ; x += alpha dn ; r -= alpha q ; d = dn ; partial r.r
//   -- grid.sync; every block sums the partials -> delta_new, beta, convergence state
//
// alpha, beta, delta and the per-column state are replicated in the shared memory of every
// block (the same deterministic arithmetic everywhere), so no block ever waits for a scalar
// written by another one; block 0 mirrors them to HBM for the host.  Rows are owned by G
// lanes: for k > 1 lane cp owns the 128-bit column pack cp (as spmm_kernel); for k = 1 the
// G lanes split the non-zeros of the row (as spmv1_kernel).
// ---------------------------------------------------------------------------
constexpr int FUSED_THREADS = 1024;
constexpr int FUSED_MAXK = 128;

// sums over the lanes with equal (lane % G) in every warp, then over the 32 warps (fixed order);
// s_out[c], c < G*V, holds the block's sum for column c
template <typename T, int V>
__device__ __forceinline__ void fused_block_reduce(T (&acc)[V], int G, T *s_warp, T *s_out) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int off = 16; off >= G; off >>= 1) {
#pragma unroll
        for (int v = 0; v < V; v++) {
            if constexpr (Sc<T>::cplx) {
                acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, off);
                acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, off);
            } else {
                acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], off);
            }
        }
    }
    if (lane < G) {
#pragma unroll
        for (int v = 0; v < V; v++) s_warp[warp * (G * V) + lane * V + v] = acc[v];
    }
    __syncthreads();
    if (t < G * V) {
        T sum = Sc<T>::zero();
        for (int w = 0; w < FUSED_THREADS / 32; w++) sum = Sc<T>::add(sum, s_warp[w * (G * V) + t]);
        s_out[t] = sum;
    }
    __syncthreads();
}

// publish the block's column sums, grid barrier, then EVERY block adds up all blocks' partials
template <typename T>
__device__ __forceinline__ void fused_grid_sum(cooperative_groups::grid_group &grid, const T *s_out, int k, T *partial,
                                               T *s_tot) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t < k) partial[(size_t)blockIdx.x * k + t] = s_out[t];
    grid.sync();
    for (int c = warp; c < k; c += FUSED_THREADS / 32) {
        T sum = Sc<T>::zero();
        // 8 partials per lane per round, all loads issued before the first add (one L2 round trip)
        for (int b0 = 0; b0 < (int)gridDim.x; b0 += 256) {
            T part[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int bb = b0 + lane + 32 * u;
                part[u] = bb < (int)gridDim.x ? ld_cg(partial + (size_t)bb * k + c) : Sc<T>::zero();
            }
#pragma unroll
            for (int u = 0; u < 8; u++) sum = Sc<T>::add(sum, part[u]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            if constexpr (Sc<T>::cplx) {
                sum.x += __shfl_xor_sync(0xffffffffu, sum.x, off);
                sum.y += __shfl_xor_sync(0xffffffffu, sum.y, off);
            } else {
                sum += __shfl_xor_sync(0xffffffffu, sum, off);
            }
        }
        if (lane == 0) s_tot[c] = sum;
    }
    __syncthreads();
}

template <typename T, int V, bool MULTI>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
cg_fused_kernel(int n, int k, int G, const T *__restrict__ vals, const int *__restrict__ rowptr,
                const int *__restrict__ cols, const T *b /* aliases d */, T *x, T *r, T *d, T *q,
                T *partial_a, T *partial_b, CgScalars<T> sc, int maxit) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    using P = Pack<T, V>;
    __shared__ T s_warp[32 * 32 * V];
    __shared__ T s_out[32 * V];
    __shared__ T s_tot[FUSED_MAXK];
    __shared__ T s_alpha[FUSED_MAXK], s_beta[FUSED_MAXK], s_dnew[FUSED_MAXK], s_dold[FUSED_MAXK];
    __shared__ double s_d0[FUSED_MAXK];
    __shared__ int s_state[FUSED_MAXK], s_iters[FUSED_MAXK];
    __shared__ int s_live;

    const int t = threadIdx.x;
    const int cp = t % G;                       // column pack (MULTI) or lane within the row (k = 1)
    const int kv = MULTI ? k / V : 1;
    const bool active = MULTI ? (cp < kv) : true;
    const int rows_per_block = FUSED_THREADS / G;
    const double tol = *sc.tol;
    const int ncomp = Sc<T>::cplx ? 2 : 1;

    // y = A * w(col) for this thread's share of one row; w is produced by `load_w`
    // When one pass of the grid covers every row, a thread works on the same row in every iteration:
    // the row (<= UB coefficients and column indices per thread for k = 1, per row for k > 1) is then
    // read from HBM/L2 once and kept on chip for the whole solve -- in registers for k = 1, in shared
    // memory (broadcast to the row's G lanes) for k > 1 -- and phase 1 is a single round of gathers.
    // cnt_my: -1 not cached (generic path), >= 0 cached.
    constexpr int UB = MULTI ? 8 : 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *s_a = reinterpret_cast<T *>(smem_raw);                                  // [rows_per_block][UB]  (MULTI)
    int *s_c = reinterpret_cast<int *>(smem_raw + (size_t)rows_per_block * UB * sizeof(T));
    int c_my[MULTI ? 1 : UB];
    T a_my[MULTI ? 1 : UB];
    int cnt_my = -1;
    const int lrow = t / G;                                                    // row slot inside the block
    if ((long long)gridDim.x * rows_per_block >= n) {
        const int row = (int)blockIdx.x * rows_per_block + lrow;
        cnt_my = 0;
        if constexpr (!MULTI) {
#pragma unroll
            for (int u = 0; u < UB; u++) {
                c_my[u] = 0;
                a_my[u] = Sc<T>::zero();
            }
        }
        int lo = 0, hi = 0;
        if (row < n) {
            lo = __ldg(rowptr + row);
            hi = __ldg(rowptr + row + 1);
        }
        if constexpr (MULTI) {
            const int cnt = hi - lo;
            // the row's G lanes fill its UB slots together (slot u by lane u, u + G, ...)
            for (int u = cp; u < UB; u += G) {
                const bool ok = u < cnt && cnt <= UB;
                s_c[lrow * UB + u] = ok ? __ldg(cols + lo + u) : 0;
                s_a[lrow * UB + u] = ok ? __ldg(vals + lo + u) : Sc<T>::zero();
            }
            cnt_my = (cnt <= UB) ? cnt : -1;
            __syncthreads();
        } else {
            const int start = lo + cp;
            const int cnt = start < hi ? (hi - start + G - 1) / G : 0;
            if (cnt <= UB) {
                cnt_my = cnt;
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    if (u < cnt) {
                        c_my[u] = __ldg(cols + start + u * G);
                        a_my[u] = __ldg(vals + start + u * G);
                    }
                }
            } else {
                cnt_my = -1;
            }
        }
    }
    auto row_product = [&](int row, bool valid, auto load_w, T (&acc)[V]) {
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = Sc<T>::zero();
        if (cnt_my >= 0) {
            if constexpr (MULTI) {
                if (valid) {
                    // two half batches: 4 gathers (8 packs in phase 1) in flight per thread
#pragma unroll
                    for (int h = 0; h < UB; h += 4) {
                        if (h < cnt_my) {
                            T w[4][V];
#pragma unroll
                            for (int u = 0; u < 4; u++) load_w(s_c[lrow * UB + h + u], w[u]);   // padded: column 0, coefficient 0
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const T a = s_a[lrow * UB + h + u];
#pragma unroll
                                for (int v = 0; v < V; v++) acc[v] = Sc<T>::fma(a, w[u][v], acc[v]);
                            }
                        }
                    }
                }
            } else {
                T w[UB][V];
#pragma unroll
                for (int u = 0; u < UB; u++) load_w(c_my[u], w[u]);
#pragma unroll
                for (int u = 0; u < UB; u++) acc[0] = Sc<T>::fma(a_my[u], w[u][0], acc[0]);
            }
        } else {
            int lo = 0, hi = 0;
            if (valid) {
                lo = __ldg(rowptr + row);
                hi = __ldg(rowptr + row + 1);
            }
            const int start = MULTI ? lo : lo + cp, stride = MULTI ? 1 : G;
            for (int j0 = start; j0 < hi; j0 += 4 * stride) {
                T a[4], w[4][V];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int j = j0 + u * stride;
                    const bool ok = j < hi;
                    const int jj = ok ? j : j0;
                    load_w(__ldg(cols + jj), w[u]);
                    a[u] = ok ? __ldg(vals + jj) : Sc<T>::zero();
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
#pragma unroll
                    for (int v = 0; v < V; v++) acc[v] = Sc<T>::fma(a[u], w[u][v], acc[v]);
                }
            }
        }
        if constexpr (!MULTI) {
            // every lane of the warp takes part, whatever its row
            for (int off = G >> 1; off > 0; off >>= 1) {
                if constexpr (Sc<T>::cplx) {
                    acc[0].x += __shfl_xor_sync(0xffffffffu, acc[0].x, off);
                    acc[0].y += __shfl_xor_sync(0xffffffffu, acc[0].y, off);
                } else {
                    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
                }
            }
        }
    };
    // element index of this thread's pack in row `row` (MULTI) / of the row's single value (k = 1)
    auto at = [&](int row) -> size_t { return MULTI ? (size_t)row * k + (size_t)cp * V : (size_t)row; };
    const bool owner = MULTI ? active : (cp == 0);   // the thread that stores a row's result

    // ---------------- initialisation: q = A x0 ; r = b - q ; d = r ; delta = r.r    clcg.c:253-292
    for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < n; row0 += (long long)gridDim.x * rows_per_block) {
        const int row = (int)row0 + t / G;
        const bool valid = row < n;
        T acc[V];
        row_product(row, valid && active, [&](int c, T (&w)[V]) {
            if constexpr (MULTI) {
                const P p = *reinterpret_cast<const P *>(x + (size_t)c * k + (size_t)cp * V);
#pragma unroll
                for (int v = 0; v < V; v++) w[v] = p.v[v];
            } else {
                w[0] = x[c];
            }
        }, acc);
        if (valid && owner) {
            if constexpr (MULTI) {
                P out;
#pragma unroll
                for (int v = 0; v < V; v++) out.v[v] = acc[v];
                *reinterpret_cast<P *>(q + at(row)) = out;
            } else {
                q[row] = acc[0];
            }
        }
    }
    grid.sync();
    {
        T acc[V];
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = Sc<T>::zero();
        for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < n; row0 += (long long)gridDim.x * rows_per_block) {
            const int row = (int)row0 + t / G;
            if (row < n && owner) {
                const size_t e = at(row);
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const T rv = Sc<T>::sub(b[e + v], q[e + v]);
                    r[e + v] = rv;
                    d[e + v] = rv;
                    acc[v] = Sc<T>::fma(rv, rv, acc[v]);
                }
            }
        }
        fused_block_reduce<T, V>(acc, G, s_warp, s_out);
        fused_grid_sum<T>(grid, s_out, k, partial_b, s_tot);
    }
    if (t < k) {
        const T dl = s_tot[t];
        const double a0 = Sc<T>::abs(dl);
        s_dnew[t] = dl;
        s_dold[t] = dl;
        s_d0[t] = a0;
        s_beta[t] = Sc<T>::zero();
        s_alpha[t] = Sc<T>::zero();
        s_state[t] = ((a0 > 0.0) && Sc<T>::finite(dl)) ? ST_ACTIVE : (a0 == 0.0 ? ST_CONVERGED : ST_BREAKDOWN);
        s_iters[t] = 0;
        if (blockIdx.x == 0 && sc.hist && sc.hist_cap > 0) Sc<T>::to_double2(dl, sc.hist + (size_t)t * ncomp);
    }
    __syncthreads();
    if (t == 0) {
        int live = 0;
        for (int c = 0; c < k; c++) live += (s_state[c] == ST_ACTIVE);
        s_live = live;
    }
    __syncthreads();

    // ---------------- the loop, clcg.c:296-419
    int it = 0;
    for (; it < maxit && s_live > 0; it++) {
        // partials of phase 1 go to buffer a, of phase 2 (and of the initialisation) to buffer b: a fast
        // block can then never overwrite partials that a slow block is still summing
        // phase 1: q = A (r + beta d), partial (r + beta d).q
        T dot[V];
#pragma unroll
        for (int v = 0; v < V; v++) dot[v] = Sc<T>::zero();
        T beta[V];
#pragma unroll
        for (int v = 0; v < V; v++) beta[v] = MULTI ? (active ? s_beta[cp * V + v] : Sc<T>::zero()) : s_beta[0];
        for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < n; row0 += (long long)gridDim.x * rows_per_block) {
            const int row = (int)row0 + t / G;
            const bool valid = row < n;
            T acc[V];
            row_product(row, valid && active, [&](int c, T (&w)[V]) {
                if constexpr (MULTI) {
                    const size_t e = (size_t)c * k + (size_t)cp * V;
                    const P rp = *reinterpret_cast<const P *>(r + e);
                    const P dp = *reinterpret_cast<const P *>(d + e);
#pragma unroll
                    for (int v = 0; v < V; v++) w[v] = Sc<T>::fma(beta[v], dp.v[v], rp.v[v]);
                } else {
                    w[0] = Sc<T>::fma(beta[0], d[c], r[c]);
                }
            }, acc);
            if (valid && owner) {
                const size_t e = at(row);
#pragma unroll
                for (int v = 0; v < V; v++) {
                    q[e + v] = acc[v];
                    const T dn = Sc<T>::fma(beta[v], d[e + v], r[e + v]);
                    dot[v] = Sc<T>::fma(dn, acc[v], dot[v]);
                }
            }
        }
        if (!MULTI && cp != 0) dot[0] = Sc<T>::zero();
        fused_block_reduce<T, V>(dot, MULTI ? G : 1, s_warp, s_out);
        fused_grid_sum<T>(grid, s_out, k, partial_a, s_tot);
        if (t < k) {
            T al = Sc<T>::zero();
            if (s_state[t] == ST_ACTIVE && !Sc<T>::is_zero(s_tot[t])) al = Sc<T>::div(s_dnew[t], s_tot[t]);
            s_alpha[t] = al;
        }
        __syncthreads();

        // phase 2: dn = r + beta d ; x += alpha dn ; r -= alpha q ; d = dn ; partial r.r
        T acc2[V], alpha[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
            acc2[v] = Sc<T>::zero();
            alpha[v] = MULTI ? (active ? s_alpha[cp * V + v] : Sc<T>::zero()) : s_alpha[0];
        }
        for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < n; row0 += (long long)gridDim.x * rows_per_block) {
            const int row = (int)row0 + t / G;
            if (row < n && owner) {
                const size_t e = at(row);
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const T dn = Sc<T>::fma(beta[v], d[e + v], r[e + v]);
                    x[e + v] = Sc<T>::fma(alpha[v], dn, x[e + v]);
                    const T rv = Sc<T>::fnma(alpha[v], q[e + v], r[e + v]);
                    r[e + v] = rv;
                    d[e + v] = dn;
                    acc2[v] = Sc<T>::fma(rv, rv, acc2[v]);
                }
            }
        }
        fused_block_reduce<T, V>(acc2, MULTI ? G : 1, s_warp, s_out);
        fused_grid_sum<T>(grid, s_out, k, partial_b, s_tot);
        if (t < k) {
            T be = Sc<T>::zero();
            if (s_state[t] == ST_ACTIVE) {
                const T nd = s_tot[t];
                s_dold[t] = s_dnew[t];
                s_dnew[t] = nd;
                const double a = Sc<T>::abs(nd);
                int st = ST_ACTIVE;
                if (!Sc<T>::finite(nd)) st = ST_BREAKDOWN;
                else if (a == 0.0 || (tol > 0.0 && sqrt(a / s_d0[t]) < tol)) st = ST_CONVERGED;
                if (st != ST_ACTIVE) {
                    s_state[t] = st;
                    s_iters[t] = it + 1;
                } else if (!Sc<T>::is_zero(s_dold[t])) {
                    be = Sc<T>::div(nd, s_dold[t]);
                }
            }
            s_beta[t] = be;
            if (blockIdx.x == 0 && sc.hist && it + 1 < sc.hist_cap)
                Sc<T>::to_double2(s_dnew[t], sc.hist + ((size_t)(it + 1) * k + t) * ncomp);
        }
        __syncthreads();
        if (t == 0) {
            int live = 0;
            for (int c = 0; c < k; c++) live += (s_state[c] == ST_ACTIVE);
            s_live = live;
        }
        __syncthreads();
    }

    // ---------------- a column that is still active ends with d = r + beta d pending: nothing reads d again.
    if (blockIdx.x == 0) {
        if (t < k) {
            sc.delta_new[t] = s_dnew[t];
            sc.delta_old[t] = s_dold[t];
            sc.delta0[t] = s_d0[t];
            sc.state[t] = s_state[t];
            sc.iters[t] = s_iters[t];
        }
        if (t == 0) {
            *sc.n_active = s_live;
            *sc.it = it;
        }
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
    // 1-D grid over 32x32 tiles (either dimension may be the long one: [k][n] -> [n][k] and back)
    const long long tiles_c = (cols + 31) / 32;
    const long long c0 = ((long long)blockIdx.x % tiles_c) * 32;
    const long long r0 = ((long long)blockIdx.x / tiles_c) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long rr = r0 + i;
        const long long cc = c0 + threadIdx.x;
        if (rr < rows && cc < cols) tile[i][threadIdx.x] = src[(size_t)rr * cols + cc];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long cc = c0 + i;
        const long long rr = r0 + threadIdx.x;
        if (rr < rows && cc < cols) dst[(size_t)cc * rows + rr] = tile[threadIdx.x][i];
    }
}

}  // namespace cgb

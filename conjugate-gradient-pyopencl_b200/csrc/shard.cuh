// shard.cuh -- row-block sharded CG across the GPUs of one box (included by cgb200.cu).
//
// The reference's "multi-GPU" splits right-hand sides and never communicates
// (p_h-PY_C-CL-multi-GPU.py:2123-2181); that mode needs nothing beyond one handle per
// device.  This file is the mode BASELINE.json's north star adds: rank g owns the rows
// [r_g, r_{g+1}) of A and the same slices of x, r, d, q.  Per iteration
//
//   exchange   the entries of d that other ranks' rows reference (the halo) travel
//              peer to peer over NVLink: pack -> ncclSend/ncclRecv (one group call)
//   spmv_dot   q = A_local [d_owned | d_halo], partial d.q       -> all-reduce (k scalars)
//   update_xr  x, r updates, partial r.r                          -> all-reduce (k scalars)
//   bookkeep   delta shuffle / convergence on the reduced value (update_bookkeep_kernel)
//   update_d   d = r + beta d
//
// One process per GPU (torchrun); NCCL is resolved with dlopen at first use so that
// liboclcg.so has no link-time dependency on it and shares the copy torch already loaded.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types only; the functions are looked up at run time

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return;
        }
#define NCCL_SYM(field, name)                                              \
    api.field = (decltype(api.field))dlsym(api.lib, name);                 \
    if (!api.field) api.error = std::string("libnccl lacks ") + name;
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        NCCL_SYM(CommInitRank, "ncclCommInitRank")
        NCCL_SYM(CommDestroy, "ncclCommDestroy")
        NCCL_SYM(AllReduce, "ncclAllReduce")
        NCCL_SYM(Send, "ncclSend")
        NCCL_SYM(Recv, "ncclRecv")
        NCCL_SYM(GroupStart, "ncclGroupStart")
        NCCL_SYM(GroupEnd, "ncclGroupEnd")
        NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    });
    return &api;
}

#define NC(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess)                                                                           \
            return fail(CGB200_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(r_)); \
    } while (0)

#define DBG(...)                                   \
    do {                                           \
        if (getenv("CGB200_DEBUG")) {              \
            fprintf(stderr, "[cgb200] " __VA_ARGS__); \
            fprintf(stderr, "\n");                 \
            fflush(stderr);                        \
        }                                          \
    } while (0)

struct cgb200_shard_ctx {
    cgb200_ctx *m = nullptr;     // the local rows, columns renumbered [owned | halo]
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int n_owned = 0, n_halo = 0;
    std::vector<int> send_counts, recv_counts, send_off, recv_off;
    int send_total = 0;
    int *d_send_idx = nullptr;
    void *d_sendbuf = nullptr;
    long long exchanges = 0, allreduces = 0;
    // peer-memory collectives (CUDA IPC): see PeerComm in kernels.cuh
    bool p2p = false;
    void *p2p_buf = nullptr;             // this rank's slots + arrival flags, mapped by every peer
    PeerComm *d_peer = nullptr;          // device copy of the PeerComm
    std::vector<void *> opened;          // peers' mappings to close
    int max_send = 0;
    cudaStream_t side = nullptr;         // the halo push runs beside the interior rows of the SpMV
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool push_pending = false;
};

static constexpr size_t P2P_SLOTS_BYTES = 2 * PEER_MAX * sizeof(PeerSlot);
static constexpr size_t P2P_BUF_BYTES = P2P_SLOTS_BYTES + PEER_MAX * sizeof(unsigned long long) + 256;

template <typename T> struct ShardEngine {
    using E = Engine<T>;
    static constexpr int NCOMP = Sc<T>::cplx ? 2 : 1;
    static ncclDataType_t nccl_type() { return sizeof(typename Sc<T>::real) == 4 ? ncclFloat : ncclDouble; }

    // halo of v (v has n_owned + n_halo entries): pack what the peers need, one grouped send/recv
    static int exchange(cgb200_shard_ctx *sh, T *v, const int *n_active = nullptr) {
        cgb200_ctx *c = sh->m;
        if (sh->world == 1) return 0;
        if (sh->p2p) {
            // entries go straight into the peers' d vectors over NVLink; the consumer (SpMV) waits for the flags
            if (sh->send_total > 0) {
                // fork: the push reads v while the SpMV (which only reads v too) already runs on the main stream;
                // joined again before the next kernel that writes v (join_push)
                // a block per 2048 entries of the largest segment: 1 .. 32 for the faces of a grid partition, up to two
                // per SM when the segment is a whole owned block (random columns)
                const int nb = std::max(1, std::min(sh->max_send > (1 << 18) ? 2 * c->sm_count : 32, (sh->max_send + 2047) / 2048));
                const bool balanced_spmv = c->spmv_variant == 3 || (c->spmv_variant == 0 && c->irregular && c->auto_irregular);
                if (sh->max_send > (1 << 18) || balanced_spmv) {
                    // whole-block segments, or the per-non-zero SpMV: IN stream order, ahead of the SpMV.  Every tile of that SpMV needs the halo, so
                    // there is nothing to overlap -- and its blocks (the per-non-zero kernel: 4 x 256 threads, every
                    // register of the SM) spin for the peers' entries from their first tile on.  Forked, the push of THIS
                    // GPU may find no room beside them while the peer's blocks wait for it and the peer's push finds no
                    // room beside those: both GPUs spin for ever (seen: a 2-GPU run of config 5 that never returned).
                    halo_push_kernel<T><<<dim3(nb, sh->world), 256, 0, c->stream>>>(sh->d_peer, sh->d_send_idx, v, n_active);
                    c->launches++;
                    sh->exchanges++;
                    return 0;
                }
                CU(cudaEventRecord(sh->ev_fork, c->stream));
                CU(cudaStreamWaitEvent(sh->side, sh->ev_fork, 0));
                halo_push_kernel<T><<<dim3(nb, sh->world), 256, 0, sh->side>>>(sh->d_peer, sh->d_send_idx, v, n_active);
                CU(cudaEventRecord(sh->ev_join, sh->side));
                sh->push_pending = true;
                c->launches++;
            }
            sh->exchanges++;
            return 0;
        }
        NcclApi *N = nccl_api();
        if (sh->send_total > 0) {
            const int grid = std::min(c->sm_count * 8, (sh->send_total + 255) / 256);
            pack_kernel<T><<<grid, 256, 0, c->stream>>>(sh->send_total, sh->d_send_idx, v, (T *)sh->d_sendbuf);
            c->launches++;
        }
        NC(N->GroupStart());
        for (int p = 0; p < sh->world; p++) {
            if (p == sh->rank) continue;
            if (sh->send_counts[p] > 0)
                NC(N->Send((const T *)sh->d_sendbuf + sh->send_off[p], (size_t)sh->send_counts[p] * NCOMP, nccl_type(), p,
                           sh->comm, c->stream));
            if (sh->recv_counts[p] > 0)
                NC(N->Recv(v + sh->n_owned + sh->recv_off[p], (size_t)sh->recv_counts[p] * NCOMP, nccl_type(), p, sh->comm,
                           c->stream));
        }
        NC(N->GroupEnd());
        sh->exchanges++;
        return 0;
    }

    static int join_push(cgb200_shard_ctx *sh) {
        if (sh->push_pending) {
            CU(cudaStreamWaitEvent(sh->m->stream, sh->ev_join, 0));
            sh->push_pending = false;
        }
        return 0;
    }

    static int allreduce(cgb200_shard_ctx *sh, T *buf, int k) {
        if (sh->world == 1 || sh->p2p) return 0;      // p2p: done inside the producing kernel
        NC(nccl_api()->AllReduce(buf, buf, (size_t)k * NCOMP, nccl_type(), ncclSum, sh->comm, sh->m->stream));
        sh->allreduces++;
        return 0;
    }

    static int iteration(cgb200_shard_ctx *sh, const typename E::VecGeom &g, const CgScalars<T> &sc) {
        cgb200_ctx *c = sh->m;
        // two kernels, no exchange step: the producing kernels store the boundary entries of d and r straight
        // into the peers' halos and the all-reduces at their tails order everything (cg2.cuh)
        if (sc.cg2) return E::cg2_iteration(c, sc);
        TRY(exchange(sh, (T *)c->d, sh->p2p ? sc.n_active : nullptr));
        TRY(E::template spmv<true>(c, 1, (const T *)c->d, (T *)c->q, sc));          // q = A d, local d.q -> sc.dq
        TRY(allreduce(sh, sc.dq, 1));
        if (g.V == 1) TRY(E::template launch_update_xr<1>(c, 1, g, sc));            // local r.r -> sc.rr
        else TRY(E::template launch_update_xr<E::VW>(c, 1, g, sc));
        if (sc.defer) {
            TRY(allreduce(sh, sc.rr, 1));
            update_bookkeep_kernel<T><<<1, 32, 0, c->stream>>>(1, sc);
            c->launches++;
        }
        TRY(join_push(sh));                                                         // update_d rewrites d
        if (g.V == 1) TRY(E::template launch_update_d<1>(c, 1, g, sc));
        else TRY(E::template launch_update_d<E::VW>(c, 1, g, sc));
        return 0;
    }

    static int solve(cgb200_shard_ctx *sh, const void *b, void *x, int maxit, double tol, int *iters, double *relres,
                     double ms[4]) {
        cgb200_ctx *c = sh->m;
        TRY(E::ensure_workspace(c, 1));
        CgScalars<T> sc = E::scalars(c, 1, tol, 0);
        sc.defer = sh->p2p ? 0 : 1;                    // p2p: reduced and book-kept inside the kernels
        sc.peer = sh->p2p ? sh->d_peer : nullptr;
        if (sh->world == 1) sc.defer = 0;
        const bool cg2 = E::use_cg2(c, 1) && (sh->world == 1 || sh->p2p);
        sc.cg2 = cg2 ? 1 : 0;
        // Programmatic dependent launch: on for the two-kernel iteration (dir_spmv's prologue -- pattern table, barriers --
        // overlaps the tail of update_r: 64.1 -> 62.1 us per iteration on 38-plane shards); off for the three-kernel
        // iteration on shards, where it never gained anything
        struct PdlGuard {
            cgb200_ctx *c;
            int saved;
            ~PdlGuard() { c->pdl = saved; }
        } pdl_guard{c, c->pdl};
        if (!cg2 && sh->world > 1 && !getenv("CGB200_PDL")) c->pdl = 0;
        if (sh->p2p && c->spmv_variant != 0 && c->spmv_variant != 6 && c->spmv_variant != 3)
            return fail(CGB200_ERR_UNSUPPORTED, "peer-memory collectives need a TMA-fed SpMV schedule (spmv_variant 0, 3 or 6)");
        const typename E::VecGeom g = E::geom(c, 1);
        const size_t bytes = (size_t)sh->n_owned * sizeof(T);

        CU(cudaMemcpyAsync((void *)sc.tol, &tol, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaEventRecord(c->ev[0], c->stream));
        CU(cudaMemcpyAsync(c->x, x, bytes, cudaMemcpyDefault, c->stream));
        CU(cudaMemcpyAsync(c->d, c->x, bytes, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaEventRecord(c->ev[1], c->stream));
        // q = A x0 (x0 with its halo) ; r = b - q ; d = r ; delta = r.r           clcg.c:253-292
        TRY(exchange(sh, (T *)c->d));
        TRY(E::template spmv<false>(c, 1, (const T *)c->d, (T *)c->q, sc));
        TRY(join_push(sh));
        void *b_dev = c->d;
        if (cg2 && sh->n_halo > 0) {
            // two-kernel iteration: the halos of BOTH direction buffers stay zero for the whole solve (the peers store the
            // new direction's boundary entries into the halo of r, and r[j] + beta * 0 is what the gather forms there)
            CU(cudaMemsetAsync((T *)c->d + sh->n_owned, 0, (size_t)sh->n_halo * sizeof(T), c->stream));
            CU(cudaMemsetAsync((T *)c->d2 + sh->n_owned, 0, (size_t)sh->n_halo * sizeof(T), c->stream));
        }
        CU(cudaMemcpyAsync(b_dev, b, bytes, cudaMemcpyDefault, c->stream));
        if (g.V == 1) TRY(E::template launch_init<1>(c, 1, g, (const T *)b_dev, (const T *)c->q, (T *)c->r, (T *)c->d, sc));
        else TRY(E::template launch_init<E::VW>(c, 1, g, (const T *)b_dev, (const T *)c->q, (T *)c->r, (T *)c->d, sc));
        if (sc.defer) {
            TRY(allreduce(sh, sc.rr, 1));
            init_bookkeep_kernel<T><<<1, 32, 0, c->stream>>>(1, sc);
            c->launches++;
        }
        CU(cudaEventRecord(c->ev[2], c->stream));
        DBG("rank %d: init enqueued (maxit %d tol %g graph %d)", sh->rank, maxit, tol, c->use_graph);

        int done = 0;
        const int chunk = std::max(1, c->graph_chunk);
        if (c->use_graph && maxit >= chunk) {
            // captured once per shard (the tolerance lives in device memory); re-captured only when the
            // chunk length or the stream changed
            const int gkey = sh->p2p ? -2 : -1;
            if (!c->graph || c->graph_k != gkey || c->graph_chunk_built != chunk || c->graph_cg2 != (int)cg2) {
                drop_graph(c);
                cudaGraph_t gr = nullptr;
                const long long before = c->launches;
                DBG("rank %d: capturing %d iterations", sh->rank, chunk);
                CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                int rc = 0;
                for (int i = 0; i < chunk && rc >= 0; i++) rc = iteration(sh, g, sc);
                cudaError_t ce = cudaStreamEndCapture(c->stream, &gr);
                c->graph_nodes = c->launches - before;
                c->launches = before;
                if (rc < 0) return rc;
                if (ce != cudaSuccess) return fail(CGB200_ERR_CUDA, "graph capture: %s", cudaGetErrorString(ce));
                ce = cudaGraphInstantiate(&c->graph, gr, 0);
                cudaGraphDestroy(gr);
                if (ce != cudaSuccess) return fail(CGB200_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(ce));
                c->graph_cg2 = (int)cg2;
                c->graph_k = gkey;             // marks a shard graph (never matches a plain solve's k)
                c->graph_chunk_built = chunk;
                DBG("rank %d: graph ready (%lld nodes)", sh->rank, c->graph_nodes);
            }
            while (done + chunk <= maxit) {
                CU(cudaGraphLaunch(c->graph, c->stream));
                c->graph_launches++;
                c->launches += c->graph_nodes;
                done += chunk;
                if (tol > 0) {
                    CU(cudaMemcpyAsync(c->h_flag, sc.n_active, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                    CU(cudaStreamSynchronize(c->stream));
                    DBG("rank %d: %d iterations done, n_active=%d", sh->rank, done, *c->h_flag);
                    if (*c->h_flag == 0) { done = maxit; break; }   // identical on every rank: reduced values
                }
            }
        }
        for (; done < maxit; done++) {
            TRY(iteration(sh, g, sc));
            if (tol > 0 && (done % chunk) == chunk - 1) {
                CU(cudaMemcpyAsync(c->h_flag, sc.n_active, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                if (*c->h_flag == 0) break;
            }
        }
        if (cg2) TRY(E::launch_finish_x(c, sc));
        CU(cudaEventRecord(c->ev[3], c->stream));
        CU(cudaMemcpyAsync(x, c->x, bytes, cudaMemcpyDefault, c->stream));
        CU(cudaEventRecord(c->ev[4], c->stream));
        DBG("rank %d: loop enqueued, reading results", sh->rank);

        T dn;
        double d0;
        int st, its;
        CU(cudaMemcpyAsync(&dn, sc.delta_new, sizeof(T), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&d0, sc.delta0, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&st, sc.state, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&its, sc.iters, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        for (int i = 0; i < 4; i++) {
            float f = 0;
            CU(cudaEventElapsedTime(&f, c->ev[i], c->ev[i + 1]));
            ms[i] = f;
        }
        int flags = 0;
        if (st == ST_ACTIVE) {
            its = maxit;
            if (tol > 0) flags |= CGB200_FLAG_MAXIT;
        }
        if (st == ST_BREAKDOWN) flags |= CGB200_FLAG_BREAKDOWN;
        if (iters) *iters = its;
        if (relres) *relres = d0 > 0 ? sqrt(Sc<T>::abs(dn) / d0) : 0.0;
        return flags;
    }
};

extern "C" {

int cgb200_nccl_unique_id(void *out128) {
    if (!out128) return fail(CGB200_ERR_ARG, "NULL argument");
    NcclApi *N = nccl_api();
    if (!N->error.empty()) return fail(CGB200_ERR_NCCL, "%s", N->error.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(N->GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return CGB200_OK;
}

int cgb200_shard_create(cgb200_shard *out, int rank, int world, const void *nccl_id128, int device, int n_owned,
                        int n_halo, long long nnz, const void *aValues, const int *aPointers, const int *aColsLocal,
                        int dtype, const int *send_counts, const int *send_idx, const int *recv_counts,
                        const unsigned char *row_boundary) {
    if (!out) return fail(CGB200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world || n_owned <= 0 || n_halo < 0)
        return fail(CGB200_ERR_ARG, "bad shard arguments (rank %d of %d, n_owned=%d, n_halo=%d)", rank, world, n_owned, n_halo);
    if (world > 1 && (!nccl_id128 || !send_counts || !recv_counts)) return fail(CGB200_ERR_ARG, "NULL exchange plan");
    cgb200_shard_ctx *sh = new cgb200_shard_ctx();
    sh->rank = rank;
    sh->world = world;
    sh->n_owned = n_owned;
    sh->n_halo = n_halo;
    auto bail = [&](int rc) {
        cgb200_shard_destroy(sh);
        return rc;
    };
    // (the halo-touching rows, when given, are scheduled last: the exchange hides behind the interior rows)
    int halo_low = 0;           // the halo is sorted by global index: what lower ranks send comes first
    for (int p = 0; p < rank && world > 1; p++) halo_low += recv_counts[p];
    int rc = create_ctx(&sh->m, n_owned, nnz, aValues, aPointers, aColsLocal, dtype, device, n_halo,
                        world > 1 ? row_boundary : nullptr, nullptr, halo_low);
    if (rc < 0) return bail(rc);
    DeviceGuard guard(device);
    sh->send_counts.assign(world, 0);
    sh->recv_counts.assign(world, 0);
    sh->send_off.assign(world, 0);
    sh->recv_off.assign(world, 0);
    if (world > 1) {
        int so = 0, ro = 0;
        for (int p = 0; p < world; p++) {
            sh->send_counts[p] = send_counts[p];
            sh->recv_counts[p] = recv_counts[p];
            sh->send_off[p] = so;
            sh->recv_off[p] = ro;
            so += send_counts[p];
            ro += recv_counts[p];
        }
        if (ro != n_halo) return bail(fail(CGB200_ERR_ARG, "recv_counts sum to %d but n_halo=%d", ro, n_halo));
        if (sh->send_counts[rank] || sh->recv_counts[rank]) return bail(fail(CGB200_ERR_ARG, "a rank does not exchange with itself"));
        sh->send_total = so;
        if (so > 0) {
            if (!send_idx) return bail(fail(CGB200_ERR_ARG, "send_idx is NULL"));
            for (int i = 0; i < so; i++)
                if (send_idx[i] < 0 || send_idx[i] >= n_owned) return bail(fail(CGB200_ERR_ARG, "send_idx[%d]=%d out of range", i, send_idx[i]));
            cudaError_t e = cudaMalloc(&sh->d_send_idx, (size_t)so * sizeof(int));
            if (e == cudaSuccess) e = cudaMemcpy(sh->d_send_idx, send_idx, (size_t)so * sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMalloc(&sh->d_sendbuf, (size_t)so * dtype_size(dtype));
            if (e != cudaSuccess) return bail(fail(CGB200_ERR_CUDA, "shard buffers: %s", cudaGetErrorString(e)));
        }
        NcclApi *N = nccl_api();
        if (!N->error.empty()) return bail(fail(CGB200_ERR_NCCL, "%s", N->error.c_str()));
        ncclUniqueId id;
        memcpy(&id, nccl_id128, sizeof(id));
        ncclResult_t r = N->CommInitRank(&sh->comm, world, id, rank);
        if (r != ncclSuccess) return bail(fail(CGB200_ERR_NCCL, "ncclCommInitRank: %s", N->GetErrorString(r)));
    }
    *out = sh;
    return CGB200_OK;
}

int cgb200_shard_destroy(cgb200_shard sh) {
    if (!sh) return CGB200_OK;
    if (sh->m) {
        DeviceGuard guard(sh->m->device);
        if (sh->m->stream) cudaStreamSynchronize(sh->m->stream);
        drop_graph(sh->m);      // a captured graph holds NCCL resources: it must go before the communicator
        if (sh->side) {
            cudaStreamSynchronize(sh->side);
            cudaStreamDestroy(sh->side);
            cudaEventDestroy(sh->ev_fork);
            cudaEventDestroy(sh->ev_join);
        }
        for (void *p : sh->opened) cudaIpcCloseMemHandle(p);
        if (sh->d_peer) cudaFree(sh->d_peer);
        if (sh->p2p_buf) cudaFree(sh->p2p_buf);
        if (sh->comm) nccl_api()->CommDestroy(sh->comm);
        if (sh->d_send_idx) cudaFree(sh->d_send_idx);
        if (sh->d_sendbuf) cudaFree(sh->d_sendbuf);
        cgb200_destroy(sh->m);
    }
    delete sh;
    return CGB200_OK;
}

cgb200_handle cgb200_shard_local(cgb200_shard sh) { return sh ? sh->m : nullptr; }

// The blob a rank publishes: CUDA-IPC handle of its slot/flag buffer, CUDA-IPC handle of the ONE allocation that
// holds its d, d2, r, r2 vectors ([owned | halo] each), and their byte offsets inside it.
struct P2pBlob {
    cudaIpcMemHandle_t buf, vecs;
    long long off[4];
    char pad[CGB200_P2P_BLOB_BYTES - 2 * 64 - 4 * 8];
};
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
static_assert(sizeof(P2pBlob) == CGB200_P2P_BLOB_BYTES, "blob layout");

int cgb200_shard_p2p_export(cgb200_shard sh, void *out_blob) {
    if (!sh || !out_blob) return fail(CGB200_ERR_ARG, "NULL argument");
    if (sh->world > PEER_MAX) return fail(CGB200_ERR_UNSUPPORTED, "peer-memory collectives support up to %d ranks", PEER_MAX);
    cgb200_ctx *c = sh->m;
    DeviceGuard guard(c->device);
    TRY(DISPATCH(c, E::ensure_workspace(c, 1)));
    if (!sh->p2p_buf) {
        CU(cudaMalloc(&sh->p2p_buf, P2P_BUF_BYTES));
        CU(cudaMemset(sh->p2p_buf, 0, P2P_BUF_BYTES));
    }
    P2pBlob blob;
    memset(&blob, 0, sizeof(blob));
    CU(cudaIpcGetMemHandle(&blob.buf, sh->p2p_buf));
    CU(cudaIpcGetMemHandle(&blob.vecs, c->vec_base));          // (an IPC handle names the whole allocation)
    const long long lead = (const char *)c->vec_block - (const char *)c->vec_base;
    for (int i = 0; i < 4; i++) blob.off[i] = lead + (long long)c->vec_off[i];
    memcpy(out_blob, &blob, sizeof(blob));
    return CGB200_OK;
}

int cgb200_shard_p2p_import(cgb200_shard sh, const void *all_blobs, const long long *remote_off) {
    if (!sh || !all_blobs || !remote_off) return fail(CGB200_ERR_ARG, "NULL argument");
    if (!sh->p2p_buf) return fail(CGB200_ERR_ARG, "call cgb200_shard_p2p_export first");
    cgb200_ctx *c = sh->m;
    DeviceGuard guard(c->device);
    PeerComm pc;
    memset(&pc, 0, sizeof(pc));
    pc.rank = sh->rank;
    pc.world = sh->world;
    const P2pBlob *blobs = (const P2pBlob *)all_blobs;
    sh->max_send = 0;
    for (int p = 0; p < sh->world; p++) {
        void *buf = sh->p2p_buf, *vecs = c->vec_base;
        if (p != sh->rank) {
            CU(cudaIpcOpenMemHandle(&buf, blobs[p].buf, cudaIpcMemLazyEnablePeerAccess));
            sh->opened.push_back(buf);
            CU(cudaIpcOpenMemHandle(&vecs, blobs[p].vecs, cudaIpcMemLazyEnablePeerAccess));
            sh->opened.push_back(vecs);
        }
        pc.slots[p] = (PeerSlot *)buf;
        pc.halo_flag[p] = (unsigned long long *)((char *)buf + P2P_SLOTS_BYTES);
        // PeerComm::vec: 0, 1 the direction buffers, 2 the residual vector (vec_block order: d, d2, r, r2)
        for (int i = 0; i < 4; i++) pc.vec[i][p] = (char *)vecs + blobs[p].off[i];
        pc.d_peer[p] = pc.vec[0][p];
        pc.remote_off[p] = remote_off[p];
        pc.send_off[p] = sh->send_off[p];
        pc.recv_from[p] = sh->recv_counts[p] > 0;
        sh->max_send = std::max(sh->max_send, sh->send_counts[p]);
    }
    pc.send_off[sh->world] = sh->send_total;
    for (int p = sh->world; p < PEER_MAX; p++) pc.send_off[p + 1] = sh->send_total;
    pc.send_idx = sh->d_send_idx;
    if (!sh->d_peer) CU(cudaMalloc((void **)&sh->d_peer, sizeof(PeerComm)));
    if (!sh->side) {
        CU(cudaStreamCreateWithFlags(&sh->side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&sh->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sh->ev_join, cudaEventDisableTiming));
    }
    CU(cudaMemcpy(sh->d_peer, &pc, sizeof(pc), cudaMemcpyHostToDevice));
    drop_graph(c);
    sh->p2p = true;
    return CGB200_OK;
}

int cgb200_shard_p2p_enable(cgb200_shard sh, int on) {
    if (!sh) return fail(CGB200_ERR_ARG, "NULL shard");
    if (on && !sh->d_peer) return fail(CGB200_ERR_ARG, "peer memory was not set up (export / import)");
    drop_graph(sh->m);
    sh->p2p = on != 0;
    return CGB200_OK;
}

int cgb200_shard_set_stream(cgb200_shard sh, void *cuda_stream) {
    if (!sh) return fail(CGB200_ERR_ARG, "NULL shard");
    return cgb200_set_stream(sh->m, cuda_stream);
}

int cgb200_shard_set_option(cgb200_shard sh, const char *key, long long value) {
    if (!sh) return fail(CGB200_ERR_ARG, "NULL shard");
    return cgb200_set_option(sh->m, key, value);
}

int cgb200_shard_solve(cgb200_shard sh, const void *b_owned, void *x_owned, int max_iterations, double tol,
                       int *iterations, double *relres) {
    if (!sh || !b_owned || !x_owned || max_iterations < 0 || !(tol >= 0)) return fail(CGB200_ERR_ARG, "bad shard solve arguments");
    cgb200_ctx *c = sh->m;
    DeviceGuard guard(c->device);
    double ms[4] = {0, 0, 0, 0};
    int rc = -1;
    switch (c->dtype) {
    case CGB200_F32: rc = ShardEngine<float>::solve(sh, b_owned, x_owned, max_iterations, tol, iterations, relres, ms); break;
    case CGB200_F64: rc = ShardEngine<double>::solve(sh, b_owned, x_owned, max_iterations, tol, iterations, relres, ms); break;
    case CGB200_C64: rc = ShardEngine<float2>::solve(sh, b_owned, x_owned, max_iterations, tol, iterations, relres, ms); break;
    case CGB200_C128: rc = ShardEngine<double2>::solve(sh, b_owned, x_owned, max_iterations, tol, iterations, relres, ms); break;
    }
    for (int i = 0; i < 4; i++) c->last_ms[i] = ms[i];
    return rc;
}

int cgb200_shard_info(cgb200_shard sh, long long out[8]) {
    if (!sh || !out) return fail(CGB200_ERR_ARG, "NULL argument");
    out[0] = sh->n_owned;
    out[1] = sh->n_halo;
    out[2] = sh->send_total;
    out[3] = sh->m->launches;
    out[4] = sh->m->graph_launches;
    out[5] = sh->exchanges;
    out[6] = sh->allreduces;
    out[7] = sh->m->nnz;
    return CGB200_OK;
}

int cgb200_shard_last_timing(cgb200_shard sh, double ms[4]) {
    if (!sh) return fail(CGB200_ERR_ARG, "NULL shard");
    return cgb200_last_timing(sh->m, ms);
}

}  // extern "C"

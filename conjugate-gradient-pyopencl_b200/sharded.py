"""Multi-GPU drivers for the CG path: one process per GPU (torchrun / torch.distributed).

Two ways the path shards:

  rhs-split   what the reference does (`distribute_workloads_on_devices`,
              p_h-PY_C-CL-multi-GPU.py:2123-2140): contiguous blocks of right-hand sides per
              GPU, matrix replicated, no communication.  `split_rhs` reproduces its split.

  row-block   what BASELINE.json's north star adds: rank g owns rows [r_g, r_g+1) of A.
              `plan_row_block` renumbers the local columns [owned | halo] and works out which
              owned entries each peer needs; `ShardedMatrix` hands that to the C ABI
              (cgb200_shard_*), which moves exactly those entries over NVLink every iteration
              and all-reduces the two dot scalars.

The planning is plain numpy + torch.distributed object collectives, so it runs (and is
tested) on CPU with the gloo backend; tests/sharded_numpy.py runs the same communication
pattern in numpy over any backend as the oracle of the plan (test code, not product).
"""
import ctypes

import numpy as np

try:
    from . import _lib
except ImportError:
    import _lib


# ----------------------------------------------------------------------------------------
# rhs-split (the reference's mode)
# ----------------------------------------------------------------------------------------
def split_rhs(n_rhs, n_devices):
    """[(start, end)] per device: the first n_rhs % n_devices devices get one extra column
    (p_h-PY_C-CL-multi-GPU.py:2125-2134)."""
    per, extra = divmod(n_rhs, n_devices)
    out, start = [], 0
    for i in range(n_devices):
        end = start + per + (1 if i < extra else 0)
        out.append((start, end))
        start = end
    return out


def cg_rhs_split(size, non_zeros, a_values, b_values, a_pointers, a_cols, x_values, n_rhs, n_iterations,
                 devices=None):
    """The reference's multi-GPU solve in one process (`distribute_computations_with_threads`,
    p_h-PY_C-CL-multi-GPU.py:2142-2181): contiguous blocks of right-hand sides per device, the matrix
    replicated, one host thread per device, no device-to-device traffic.  x_values is filled in place.

    devices: CUDA ordinals (default: every visible GPU)."""
    import threading
    L = _lib.lib()
    if devices is None:
        devices = list(range(L.cgb200_device_count()))
    if not devices:
        raise RuntimeError("no CUDA device: the engine has no CPU fallback")
    a_values = np.ascontiguousarray(a_values)
    dt = a_values.dtype
    code = _lib.DTYPE_CODE[dt]
    b_values = np.ascontiguousarray(b_values, dtype=dt)
    if x_values.dtype != dt or not x_values.flags["C_CONTIGUOUS"]:
        raise TypeError("x_values must be a C-contiguous array of the matrix dtype")
    a_pointers = np.ascontiguousarray(a_pointers, dtype=np.intc)
    a_cols = np.ascontiguousarray(a_cols, dtype=np.intc)
    errors = []

    def worker(dev, start, end):
        if end <= start:
            return
        b = b_values[start * size:end * size]
        x = x_values[start * size:end * size]          # a view: the C call writes the caller's array directly
        rc = L.cgb200_cg(int(dev), code, int(size), int(non_zeros), _lib.ptr(a_values), _lib.ptr(b),
                         _lib.ptr(a_pointers), _lib.ptr(a_cols), _lib.ptr(x), end - start, int(n_iterations))
        if rc < 0:
            errors.append((dev, L.cgb200_last_error().decode(errors="replace")))

    threads = [threading.Thread(target=worker, args=(dev, s, e))
               for dev, (s, e) in zip(devices, split_rhs(n_rhs, len(devices)))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise _lib.CgError(-2, f"rhs-split solve failed: {errors}")
    return x_values


# ----------------------------------------------------------------------------------------
# row-block
# ----------------------------------------------------------------------------------------
def split_rows(indptr, world, by="rows", align=1):
    """Row boundaries [world+1].  by='rows': equal row counts (stencils: z-slabs);
    by='nnz': equal non-zero counts (irregular matrices).  Boundaries are multiples of `align`."""
    n = len(indptr) - 1
    if by == "rows":
        bounds = [(n * p) // world for p in range(world + 1)]
    elif by == "nnz":
        nnz = int(indptr[-1])
        targets = [(nnz * p) // world for p in range(world + 1)]
        bounds = [int(np.searchsorted(indptr, t, side="left")) for t in targets]
        bounds[0], bounds[-1] = 0, n
    else:
        raise ValueError(by)
    if align > 1:
        bounds = [min(n, (b // align) * align) for b in bounds[:-1]] + [n]
    for p in range(world):
        if bounds[p + 1] <= bounds[p]:
            raise ValueError(f"rank {p} would own no rows (n={n}, world={world})")
    return np.asarray(bounds, dtype=np.int64)


class ShardPlan:
    """Rank-local view of a row-block partition."""
    __slots__ = ("rank", "world", "bounds", "n_owned", "n_halo", "indptr", "cols_local", "data",
                 "halo_globals", "recv_counts", "send_counts", "send_idx", "row_boundary")


def plan_row_block(indptr, indices, data, bounds, rank, exchange=None):
    """Builds rank `rank`'s ShardPlan from the GLOBAL CSR arrays (or from its own row slice, see below).

    indptr/indices/data may be the whole matrix, or only the rows [bounds[rank], bounds[rank+1]) with
    `indptr` rebased to start at 0 (len(indptr) == n_owned + 1) -- what a rank that generated only its
    slab has.  `exchange(obj) -> list[obj per rank]` is an all-gather of Python objects
    (default: torch.distributed.all_gather_object); every rank must call plan_row_block together.
    """
    world = len(bounds) - 1
    rb, re = int(bounds[rank]), int(bounds[rank + 1])
    n_owned = re - rb
    indptr = np.asarray(indptr)
    if len(indptr) == n_owned + 1 and world > 1:
        lp = indptr - indptr[0]
        cols = np.asarray(indices[: lp[-1]])
        vals = np.asarray(data[: lp[-1]])
    else:
        lo, hi = int(indptr[rb]), int(indptr[re])
        lp = indptr[rb:re + 1] - lo
        cols = np.asarray(indices[lo:hi])
        vals = np.asarray(data[lo:hi])
    owned = (cols >= rb) & (cols < re)
    halo_globals = np.unique(cols[~owned])
    owner = np.searchsorted(bounds, halo_globals, side="right") - 1
    recv_counts = np.bincount(owner, minlength=world).astype(np.intc)
    cols_local = np.where(owned, cols - rb, n_owned + np.searchsorted(halo_globals, cols)).astype(np.intc)

    # tell every owner which of its entries this rank needs; learn what the others need from us
    need = {int(p): halo_globals[owner == p] for p in np.nonzero(recv_counts)[0]}
    if exchange is None:
        import torch.distributed as dist

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    all_need = exchange(need) if world > 1 else [need]
    send_counts = np.zeros(world, dtype=np.intc)
    send_lists = []
    for p in range(world):
        wanted = all_need[p].get(rank) if p != rank else None
        if wanted is not None and len(wanted):
            wanted = np.asarray(wanted, dtype=np.int64)
            if wanted.min() < rb or wanted.max() >= re:
                raise ValueError(f"rank {p} asks rank {rank} for rows it does not own")
            send_counts[p] = len(wanted)
            send_lists.append((wanted - rb).astype(np.intc))
    plan = ShardPlan()
    plan.rank, plan.world, plan.bounds = rank, world, np.asarray(bounds)
    plan.n_owned, plan.n_halo = n_owned, int(len(halo_globals))
    plan.indptr = np.ascontiguousarray(lp, dtype=np.intc)
    plan.cols_local = np.ascontiguousarray(cols_local)
    plan.data = np.ascontiguousarray(vals)
    plan.halo_globals = halo_globals
    plan.recv_counts = recv_counts
    plan.send_counts = send_counts
    plan.send_idx = (np.concatenate(send_lists) if send_lists else np.zeros(0, np.intc)).astype(np.intc)
    # rows that reference a halo column: their SpMV tiles are scheduled last (the exchange hides behind the rest)
    nz_per_row = np.diff(lp)
    plan.row_boundary = (np.add.reduceat(~owned, lp[:-1][nz_per_row > 0]) > 0 if owned.size else np.zeros(0, bool))
    rbnd = np.zeros(n_owned, dtype=np.uint8)
    rbnd[np.nonzero(nz_per_row > 0)[0]] = plan.row_boundary.astype(np.uint8)
    plan.row_boundary = rbnd
    return plan


class ShardedMatrix:
    """A row block of A resident on this rank's GPU, plus the NCCL communicator (`cgb200_shard_create`)."""

    def __init__(self, plan, device=None, dtype=None, p2p=True):
        import torch
        import torch.distributed as dist
        self.plan = plan
        data = plan.data if dtype is None else plan.data.astype(dtype)
        self.dtype = data.dtype
        if device is None:
            device = torch.cuda.current_device()
        L = _lib.lib()
        uid = np.zeros(128, dtype=np.uint8)
        if plan.world > 1:
            if plan.rank == 0:
                _lib.check(L.cgb200_nccl_unique_id(_lib.ptr(uid)))
            box = [uid.tobytes()]
            dist.broadcast_object_list(box, src=0)
            uid = np.frombuffer(box[0], dtype=np.uint8).copy()
        h = ctypes.c_void_p()
        _lib.check(L.cgb200_shard_create(
            ctypes.byref(h), plan.rank, plan.world, _lib.ptr(uid), int(device), plan.n_owned, plan.n_halo,
            int(data.size), _lib.ptr(np.ascontiguousarray(data)), _lib.ptr(plan.indptr), _lib.ptr(plan.cols_local),
            _lib.DTYPE_CODE[np.dtype(self.dtype)], _lib.ptr(np.ascontiguousarray(plan.send_counts)),
            _lib.ptr(np.ascontiguousarray(plan.send_idx)), _lib.ptr(np.ascontiguousarray(plan.recv_counts)),
            _lib.ptr(np.ascontiguousarray(plan.row_boundary))))
        self._h = h
        self.p2p = False
        if p2p and 1 < plan.world <= 8:
            self._setup_peer_memory(dist)

    def _setup_peer_memory(self, dist):
        """CUDA-IPC handles of every rank's exchange buffer and direction vector, all-gathered; afterwards the
        halo and the dot-product all-reduces go through NVLink-mapped pointers inside the kernels."""
        plan, L = self.plan, _lib.lib()
        mine = np.zeros(_lib.P2P_BLOB_BYTES, dtype=np.uint8)
        _lib.check(L.cgb200_shard_p2p_export(self._h, _lib.ptr(mine)))
        recv_off = np.concatenate([[0], np.cumsum(plan.recv_counts)[:-1]]).astype(np.int64)
        box = [None] * plan.world
        dist.all_gather_object(box, (mine.tobytes(), int(plan.n_owned), recv_off.tolist()))
        handles = np.frombuffer(b"".join(h for h, _, _ in box), dtype=np.uint8).copy()
        remote_off = np.array([box[p][1] + box[p][2][plan.rank] for p in range(plan.world)], dtype=np.int64)
        _lib.check(L.cgb200_shard_p2p_import(self._h, _lib.ptr(handles), _lib.ptr(remote_off)))
        self.p2p = True

    def enable_peer_memory(self, on=True):
        _lib.check(_lib.lib().cgb200_shard_p2p_enable(self._h, 1 if on else 0))
        self.p2p = bool(on)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().cgb200_shard_destroy(self._h)
            self._h = None

    __del__ = close

    def time_kernel(self, which, reps=50):
        """Mean ms of one kernel of the loop on the local row block (see cgb200_time_kernel)."""
        kinds = {"spmv_dot": 0, "update_xr": 1, "update_d": 2, "spmv": 3, "dir_spmv": 4, "update_r": 5}
        ms = ctypes.c_double()
        h = ctypes.c_void_p(_lib.lib().cgb200_shard_local(self._h))
        _lib.check(_lib.lib().cgb200_time_kernel(h, kinds[which], 1, int(reps), ctypes.byref(ms)))
        return ms.value

    def read_trace(self, iterations):
        """Timeline of the local row block's kernels in the last solve (set_option("trace", n) first)."""
        from . import engine
        return engine.read_trace(ctypes.c_void_p(_lib.lib().cgb200_shard_local(self._h)), iterations)

    def set_stream(self, cuda_stream):
        _lib.check(_lib.lib().cgb200_shard_set_stream(self._h, ctypes.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def set_option(self, key, value):
        _lib.check(_lib.lib().cgb200_shard_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key):
        """An option / read-only fact of the local row block's handle (e.g. "patterns")."""
        v = ctypes.c_longlong()
        h = ctypes.c_void_p(_lib.lib().cgb200_shard_local(self._h))
        _lib.check(_lib.lib().cgb200_get_option(h, key.encode(), ctypes.byref(v)))
        return v.value

    def info(self):
        out = (ctypes.c_longlong * 8)()
        _lib.check(_lib.lib().cgb200_shard_info(self._h, out))
        return dict(zip(("n_owned", "n_halo", "sent_per_exchange", "launches", "graph_launches", "exchanges",
                         "allreduces", "nnz"), list(out)))

    def solve(self, b_owned, x_owned=None, max_iterations=1000, tol=0.0):
        """Collective.  b_owned / x_owned: this rank's slices (numpy arrays or device tensors)."""
        if isinstance(b_owned, np.ndarray):
            b_owned = np.ascontiguousarray(b_owned, dtype=self.dtype)
            if x_owned is None:
                x_owned = np.zeros(self.plan.n_owned, dtype=self.dtype)
        elif x_owned is None:
            raise ValueError("x_owned is required for device pointers")
        its, rel = ctypes.c_int(), ctypes.c_double()
        rc = _lib.check(_lib.lib().cgb200_shard_solve(self._h, _lib.ptr(b_owned), _lib.ptr(x_owned),
                                                      int(max_iterations), float(tol), ctypes.byref(its),
                                                      ctypes.byref(rel)))
        ms = (ctypes.c_double * 4)()
        _lib.check(_lib.lib().cgb200_shard_last_timing(self._h, ms))
        return x_owned, {"flags": rc, "iterations": its.value, "relres": rel.value,
                         "timing_ms": dict(zip(("h2d", "init", "iterations", "d2h"), list(ms)))}

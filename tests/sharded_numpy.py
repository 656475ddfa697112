"""Test infrastructure (NOT product code): the row-block sharded CG in numpy over torch.distributed
(any backend) -- the oracle of the communication plan that conjugate-gradient-pyopencl_b200/sharded.py builds."""
import numpy as np


def halo_exchange_numpy(plan, v_owned, dist):
    """[v_owned | halo] for this rank: the communication pattern of ShardEngine::exchange."""
    import torch
    out = np.empty(plan.n_owned + plan.n_halo, dtype=v_owned.dtype)
    out[:plan.n_owned] = v_owned
    sendbuf = v_owned[plan.send_idx]
    reqs, recv_t, so, ro = [], [], 0, plan.n_owned
    # bytes are bytes: ship every dtype as float32 / float64 words
    view = (lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.float64 if a.dtype.itemsize % 8 == 0 else np.float32)))
    for p in range(plan.world):
        if plan.send_counts[p]:
            reqs.append(dist.isend(view(sendbuf[so:so + plan.send_counts[p]]), dst=p))
            so += plan.send_counts[p]
        if plan.recv_counts[p]:
            buf = np.empty(plan.recv_counts[p], dtype=v_owned.dtype)
            t = view(buf)
            recv_t.append((ro, buf, t))
            reqs.append(dist.irecv(t, src=p))
            ro += plan.recv_counts[p]
    for r in reqs:
        r.wait()
    for ro, buf, t in recv_t:
        out[ro:ro + buf.size] = t.numpy().view(v_owned.dtype)
    return out


def reference_sharded_cg(plan, b_owned, x_owned, iters, dist):
    """Row-block CG in numpy with the plan's halo exchange and two all-reduces per iteration:
    the recurrence of clcg.c:253-419 / helmFE_var.py:507-544 distributed by rows."""
    import scipy.sparse as sp
    import torch
    A = sp.csr_matrix((plan.data, plan.cols_local, plan.indptr), shape=(plan.n_owned, plan.n_owned + plan.n_halo))

    def allsum(z):
        z = np.asarray([z], dtype=np.complex128 if np.iscomplexobj(z) else np.float64)
        t = torch.from_numpy(z.view(np.float64))
        dist.all_reduce(t)
        return z[0]

    x = x_owned.copy()
    r = b_owned - A @ halo_exchange_numpy(plan, x, dist)
    d = r.copy()
    delta_new = allsum(np.dot(r, r))
    for _ in range(iters):
        q = A @ halo_exchange_numpy(plan, d, dist)
        alpha = delta_new / allsum(np.dot(d, q))
        x = x + alpha * d
        r = r - alpha * q
        delta_old, delta_new = delta_new, allsum(np.dot(r, r))
        d = r + (delta_new / delta_old) * d
    return x

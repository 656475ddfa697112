"""Where does a cold cg() call spend its time?  (run on the GPU box)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cg_b200
from bench import make_problem

wl, dtype = (sys.argv[1:] + ["c2", "c128"])[:2]
A, B = make_problem(wl, dtype, 1)
n, nnz = A.shape[0], A.nnz
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
vals, ptr, cols = pin(A.data), pin(A.indptr.astype(np.intc)), pin(A.indices.astype(np.intc))
b, x = pin(B), pin(np.zeros_like(B))
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    M = cg_b200.Matrix(vals, ptr, cols)
    t1 = time.perf_counter()
    x[...] = 0
    t2 = time.perf_counter()
    _, info = M.solve(b, x=x, max_iterations=256)
    t3 = time.perf_counter()
    _, info2 = M.solve(b, x=x, max_iterations=256)
    t4 = time.perf_counter()
    M.close()
    t5 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms | first solve {1e3*(t3-t2):.1f} ms {info.timing_ms} | "
          f"second solve {1e3*(t4-t3):.1f} ms | destroy {1e3*(t5-t4):.1f} ms")
os.environ["CGB200_CACHE"] = "0"
for rep in range(3):
    x[...] = 0
    t0 = time.perf_counter()
    cg_b200.cg(n, nnz, vals, b, ptr, cols, x, 1, 256)
    print(f"cg() uncached: {1e3*(time.perf_counter()-t0):.1f} ms")
os.environ["CGB200_CACHE"] = "1"
for rep in range(3):
    x[...] = 0
    t0 = time.perf_counter()
    cg_b200.cg(n, nnz, vals, b, ptr, cols, x, 1, 256)
    print(f"cg() cached: {1e3*(time.perf_counter()-t0):.1f} ms")
# pageable
vals2, ptr2, cols2, b2, x2 = (np.array(a) for a in (vals, ptr, cols, b, x))
for rep in range(2):
    x2[...] = 0
    t0 = time.perf_counter()
    cg_b200.cg(n, nnz, vals2, b2, ptr2, cols2, x2, 1, 256)
    print(f"cg() cached, pageable numpy: {1e3*(time.perf_counter()-t0):.1f} ms")

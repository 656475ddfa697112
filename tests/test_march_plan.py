"""Host logic of the plane-marching dir_spmv (csrc/cg2_march.cuh, cgb200_plan_march_runs): how strips x planes work items are
cut into runs and dealt to the thread blocks.  No GPU: the planner is plain C++ behind the C ABI.

What must hold for ANY cut (the kernel relies on it): every (strip, plane) item belongs to exactly one run; block b starts with
run b and follows the `next` links, the chains of the blocks partition the runs; on a row-block shard no block meets a run that
needs a halo plane as its FIRST piece before it has done all its other runs (the arrival flag of the neighbour comes late:
DESIGN.md 6), and the runs that touch no halo plane come first."""
import ctypes

import numpy as np
import pytest

import cg_b200
from cg_b200 import _lib


def plan(strips, planes, blocks=148, low=0, high=0, streaming=0, lz=0):
    L = _lib.lib()
    cap = strips * planes + 8
    runs = np.zeros((cap, 4), dtype=np.int32)
    grid = ctypes.c_int(0)
    n = L.cgb200_plan_march_runs(strips, planes, blocks, low, high, streaming, lz, runs.ctypes.data_as(ctypes.c_void_p), cap,
                                 ctypes.byref(grid))
    assert n > 0, _lib.lib().cgb200_last_error()
    return runs[:n], grid.value


def chains(runs, grid):
    out = []
    for b in range(grid):
        c, i = [], b
        while i >= 0:
            c.append(i)
            i = int(runs[i, 3])
            assert len(c) <= len(runs)
        out.append(c)
    return out


CASES = [
    # strips, planes, low, high, streaming, lz        what it is
    (90, 300, 0, 0, 1, 0),          # 300^3 on one GPU (vectors stream from HBM): equal segments
    (90, 38, 0, 0, 0, 0),           # one eighth of it, unsharded: contiguous cost-balanced ranges
    (90, 38, 1, 1, 0, 0),           # an inner rank of 8
    (90, 37, 0, 1, 0, 0),           # rank 0 of 8
    (90, 75, 1, 1, 1, 0),           # an inner rank of 4
    (90, 75, 1, 0, 1, 0),           # the last rank of 4
    (90, 150, 1, 0, 1, 0),          # rank 1 of 2
    (1, 1024, 0, 0, 0, 0),          # 1024^2 FE grid: one strip per grid line
    (2, 7, 0, 0, 0, 0),             # fewer items than blocks
    (1, 1, 0, 0, 0, 0),
    (2, 7, 0, 0, 0, 2),             # the option march_lz (tests/test_gpu_cg2.py uses 1, 2, 5)
    (90, 38, 1, 1, 0, 5),
    (3, 40, 1, 1, 1, 1),
]


@pytest.mark.parametrize("strips,planes,low,high,streaming,lz", CASES)
def test_every_item_once_chains_partition_the_runs_and_halo_runs_come_last(strips, planes, low, high, streaming, lz):
    runs, grid = plan(strips, planes, low=low, high=high, streaming=streaming, lz=lz)
    assert 1 <= grid <= min(148, len(runs))
    seen = np.zeros((strips, planes), dtype=np.int32)
    for s, z0, ln, _ in runs:
        assert 0 <= s < strips and 0 <= z0 and ln >= 1 and z0 + ln <= planes
        seen[s, z0:z0 + ln] += 1
    assert (seen == 1).all()
    ch = chains(runs, grid)
    flat = sorted(i for c in ch for i in c)
    assert flat == list(range(len(runs)))                    # a partition: every run in exactly one block's chain
    for c in ch:
        kinds = []
        for i in c:
            s, z0, ln, _ = runs[i]
            bottom = bool(low) and z0 == 0
            top = bool(high) and z0 + ln == planes
            kinds.append(2 if bottom else (1 if top else 0))
        assert kinds == sorted(kinds), kinds                 # interior runs, then top-halo runs, then bottom-halo runs


def test_contiguous_ranges_are_balanced_and_equal_segments_are_used_where_they_must():
    # unsharded, L2-resident: no block gets more than ~10 % above the mean cost (items + 0.7 per run)
    runs, grid = plan(90, 38)
    cost = [sum(runs[i, 2] + 0.7 for i in c) for c in chains(runs, grid)]
    assert grid == 148 and max(cost) < 1.12 * (90 * 38 / 148.0 + 0.7)
    # streaming sizes and shards: every run of a strip has the same length (the last one may be shorter)
    for kw in (dict(streaming=1), dict(low=1, high=1), dict(high=1)):
        runs, _ = plan(90, 75, **kw)
        lens = sorted(set(int(r[2]) for r in runs))
        assert len(lens) <= 2, (kw, lens)


def test_bad_arguments_are_rejected():
    L = _lib.lib()
    g = ctypes.c_int(0)
    assert L.cgb200_plan_march_runs(0, 5, 148, 0, 0, 0, 0, None, 0, ctypes.byref(g)) < 0
    assert L.cgb200_plan_march_runs(3, 5, 148, 0, 0, 0, 0, None, 0, None) < 0
    # capacity 0: only the count and the grid
    assert L.cgb200_plan_march_runs(3, 5, 148, 0, 0, 0, 0, None, 0, ctypes.byref(g)) >= 3 and g.value >= 1

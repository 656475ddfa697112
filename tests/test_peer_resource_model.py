"""Why the halo push of a row-block shard may NOT be forked onto a side stream when the SpMV that waits for it fills the
GPU (csrc/shard.cuh::exchange) -- a model of the block scheduler, no GPU.

Per GPU and iteration two kernels become ready at the same moment (the previous kernel of the main stream has ended):

    P   halo_push_kernel: stores this GPU's entries into the peers' vectors, raises their arrival flags, ends
    S   the SpMV: its blocks are persistent and SPIN (peer_wait_halo) until every peer's flag has arrived

The hardware places blocks of ready kernels wherever they fit, in no guaranteed order across streams.  S never gives its
slots back before the flags arrive; P cannot run before it gets a slot.  The per-non-zero balanced SpMV (spmv_tma_kernel:
4 x 256 threads x 64 registers = every register of an SM) leaves P no room, so "S first on both GPUs" is a state in which
both wait for ever -- seen on hardware as a 2-GPU run of BASELINE config 5 that never returned.  The row-direct
and pattern kernels leave room; that is why the fork was fine until the balanced kernel became peer-aware.  In stream order
(P, then S on the SAME stream) there is no such state, whatever S occupies."""
import itertools

import pytest


def explore(world, capacity, s_slots, p_slots, forked):
    """All placement orders of the ready kernels on every GPU.  Returns (completed, deadlocked) counts."""
    completed = deadlocked = 0
    # per GPU: the order in which the scheduler happens to place the two ready kernels (forked) or the stream order
    orders = [("P", "S"), ("S", "P")] if forked else [("P", "S")]
    for choice in itertools.product(orders, repeat=world):
        free = [capacity] * world
        placed = [set() for _ in range(world)]
        pushed = [False] * world                 # this GPU's flags have been raised at its peers
        done_s = [False] * world
        progress = True
        while progress:
            progress = False
            for g in range(world):
                for k in choice[g]:
                    if k in placed[g]:
                        continue
                    need = p_slots if k == "P" else s_slots
                    # stream order: S is not even ready before P has ended
                    if not forked and k == "S" and not pushed[g]:
                        break
                    if free[g] >= need:
                        free[g] -= need
                        placed[g].add(k)
                        progress = True
                    else:
                        break                    # (a kernel that does not fit blocks nothing else in the model: the next is tried later)
                if "P" in placed[g] and not pushed[g]:
                    pushed[g] = True             # P runs to its end without waiting for anybody
                    free[g] += p_slots
                    progress = True
                if "S" in placed[g] and not done_s[g] and all(pushed[p] for p in range(world) if p != g):
                    done_s[g] = True             # every peer's flag has arrived: the SpMV finishes
                    free[g] += s_slots
                    progress = True
        if all(done_s):
            completed += 1
        else:
            deadlocked += 1
    return completed, deadlocked


@pytest.mark.parametrize("world", [2, 3, 4])
def test_forked_push_deadlocks_exactly_when_the_spinning_spmv_leaves_it_no_room(world):
    # the balanced kernel: S takes the whole GPU
    ok, dead = explore(world, capacity=8, s_slots=8, p_slots=1, forked=True)
    assert dead > 0 and ok > 0                       # some placement orders are fatal -- "it ran fine yesterday"
    # row-direct / pattern kernels: S + P fit side by side
    assert explore(world, capacity=8, s_slots=6, p_slots=2, forked=True) == (2 ** world, 0)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_stream_ordered_push_never_deadlocks_whatever_the_spmv_occupies(world):
    for s_slots in (1, 6, 8):
        assert explore(world, capacity=8, s_slots=s_slots, p_slots=2, forked=False) == (1, 0)

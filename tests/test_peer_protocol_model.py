"""A model of the peer-memory protocol of the row-block sharded solve (csrc/kernels.cuh: PeerComm,
halo_push_kernel, peer_wait_halo, peer_allreduce), checked over random interleavings -- no GPU.

What the device does, per rank and per solve (csrc/shard.cuh::solve / ::iteration):

    init:       push halo (exchange)  ->  SpMV waits for the peers' halo  ->  all-reduce (r.r)
    iteration:  push halo             ->  SpMV waits                      ->  all-reduce (d.q)  ->  all-reduce (r.r)

An exchange carries no number of its own: the sender raises the receiver's flag to `seq + 1`, seq being the number
of all-reduces the SENDER has completed, and the receiver waits for `flag >= its own seq + 1`.  All-reduces keep the
ranks in lockstep (nobody completes all-reduce s before everybody has contributed to it), so between two exchanges
every rank's seq has grown and the numbers are unique -- as long as all-reduces keep happening.  After convergence
they do not: the kernels of the loop return at once.  The model shows that

  * with pushes that continue after convergence (the behaviour until this round) a rank can pass the first wait of the
    NEXT solve on a flag raised by a stale, post-convergence push -- the bug seen on hardware as a tolerance solve that
    stopped one iteration late;
  * with `halo_push_kernel` returning at once when no column is active, every wait is satisfied by the push of the
    same logical exchange, in every interleaving tried;
  * no push ever lands in a halo region its receiver is still reading (the two all-reduces of an iteration are the
    barrier that guarantees it).
"""
import random

import pytest


class Rank:
    """Program of one GPU as a list of steps; `pc` walks through it."""

    def __init__(self, rank, world, solves, push_after_convergence):
        self.rank, self.world = rank, world
        self.seq = 0                       # all-reduces completed (PeerComm::seq)
        self.flag = [0] * world            # flag[p]: raised by rank p's push (PeerComm::halo_flag)
        self.flag_exchange = [None] * world  # model only: which logical exchange raised it
        self.contrib = {}                  # all-reduce number -> set of ranks whose slot has arrived
        self.reads_done = 0                # exchanges whose halo this rank has finished reading (its SpMV is over)
        self.reading = False               # between passing a wait and the all-reduce that ends that SpMV
        self.program = []
        exchange = 0
        for iters, converged_at, chunk in solves:
            # initialisation: always pushes (n_active == NULL in halo_push_kernel)
            self.program += [("push", exchange), ("wait", exchange), ("allreduce",)]
            exchange += 1
            active = True
            launched = 0
            while launched < iters:
                # the host launches whole chunks and polls n_active only between them (graph_chunk iterations)
                for _ in range(chunk):
                    launched += 1
                    if active:
                        self.program += [("push", exchange), ("wait", exchange), ("allreduce",), ("allreduce",)]
                        exchange += 1
                        if launched == converged_at:
                            active = False
                    else:
                        # every kernel of the loop returns at once; the push only with the new rule
                        if push_after_convergence:
                            self.program.append(("push", None))
                if not active:
                    break
        self.pc = 0

    def done(self):
        return self.pc >= len(self.program)


def neighbours(rank, world):
    return [p for p in (rank - 1, rank + 1) if 0 <= p < world]


def run(world, solves, push_after_convergence, rng):
    """Random interleaving.  Returns the list of violations: (rank, peer, exchange waited for, exchange seen)."""
    ranks = [Rank(r, world, solves, push_after_convergence) for r in range(world)]
    violations = []
    stuck = 0
    while not all(r.done() for r in ranks):
        r = rng.choice([x for x in ranks if not x.done()])
        step = r.program[r.pc]
        progressed = False
        if step[0] == "push":
            for p in neighbours(r.rank, world):
                if step[1] is not None and ranks[p].reads_done < step[1]:
                    # the new entries would land in a halo region the receiver has not finished reading
                    violations.append((r.rank, p, "overwrite", step[1], ranks[p].reads_done))
                ranks[p].flag[r.rank] = r.seq + 1          # st.release flag = seq + 1   (halo_push_kernel)
                ranks[p].flag_exchange[r.rank] = step[1]
            r.pc += 1
            progressed = True
        elif step[0] == "wait":
            want = r.seq + 1                                # peer_wait_halo
            if all(r.flag[p] >= want for p in neighbours(r.rank, world)):
                for p in neighbours(r.rank, world):
                    if r.flag_exchange[p] != step[1]:
                        violations.append((r.rank, p, step[1], r.flag_exchange[p]))
                r.reading = True
                r.pc += 1
                progressed = True
        else:                                               # peer_allreduce number seq + 1
            s = r.seq + 1
            if r.rank not in ranks[0].contrib.setdefault(s, set()):
                if r.reading:                               # the all-reduce sits at the END of the SpMV: its reads are over
                    r.reading = False
                    r.reads_done += 1
                for q in ranks:                             # store the slot into every peer (and itself)
                    q.contrib.setdefault(s, set()).add(r.rank)
                progressed = True
            if len(r.contrib[s]) == world:                  # all slots have arrived here
                r.seq = s
                r.pc += 1
                progressed = True
        stuck = 0 if progressed else stuck + 1
        assert stuck < 10000, "deadlock in the model"
    return violations


SOLVES_TOL = [(64, 37, 16), (64, 41, 16)]        # two tolerance solves that converge inside a 16-iteration chunk
SOLVES_FIXED = [(32, None, 16), (32, None, 16)]  # fixed-iteration solves never stop all-reducing


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_pushes_stop_with_convergence_every_wait_sees_its_own_exchange(world):
    rng = random.Random(world)
    for _ in range(30):
        assert run(world, SOLVES_TOL, push_after_convergence=False, rng=rng) == []
        assert run(world, SOLVES_FIXED, push_after_convergence=False, rng=rng) == []


@pytest.mark.parametrize("world", [2, 4])
def test_the_model_reproduces_the_stale_push_of_the_old_rule(world):
    """With pushes continuing after convergence some interleaving lets the next solve's first wait through on a
    stale flag (exchange None = a post-convergence push); fixed-iteration solves were never affected."""
    rng = random.Random(100 + world)
    bad = []
    for _ in range(60):
        bad += run(world, SOLVES_TOL, push_after_convergence=True, rng=rng)
    assert bad and all(len(v) == 4 and v[3] is None for v in bad)
    for _ in range(20):
        assert run(world, SOLVES_FIXED, push_after_convergence=True, rng=rng) == []

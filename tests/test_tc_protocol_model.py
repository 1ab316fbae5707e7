"""Model check of the mbarrier protocol of the tcgen05 TD kernel (reagent_b200/csrc/rb200_dqn_tc.cu).

The kernel's step loop is synchronised only by mbarriers that are waited on by PARITY:
  full[s] / sfree[s]  shared-memory weight ring (producer warp <-> 4 loader warps)
  afull[t] / adone[t] tensor-memory weight ring (4 loader warps <-> MMA warp)
  dready[t]           accumulator tile t complete (MMA warp -> 8 epilogue warps), one per tile
  opready             next B operand in shared memory (epilogue threads -> MMA warp)
A parity wait cannot tell "the phase I want" from "two phases later", so the protocol is only
correct if no waiter can fall two completions behind.  This test restates the four roles as
small state machines over an exact mbarrier model (completion counter; `wait(parity)` passes iff
the counter's parity differs), runs them under many random interleavings and network shapes and
asserts (a) progress to the end (no deadlock) and (b) that no waiter ever faces a barrier that
is already >= 2 completions past the one it intends to observe.  It also checks that the model
DOES catch the bug the first version of the kernel had (one shared `dready` barrier for all the
tiles of a step), so the check is not vacuous.
"""
import random

import pytest

STAGES = 6   # kQStages (shared memory)
ASTAGES = 4  # kAStages (tensor memory)
LOADERS = 4  # loader warps: one arrival each on sfree / afull


class MBar:
    def __init__(self, count):
        self.count = count
        self.pending = count
        self.completed = 0  # number of completed phases

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.completed += 1
            self.pending = self.count


class Hazard(Exception):
    pass


def wait_ok(bar, intended):
    """Parity wait for the `intended`-th completion (1-based).  Returns True when it passes."""
    parity = (intended - 1) & 1
    if bar.completed >= intended + 1 and (bar.completed & 1) == parity:
        # the waiter is two (or more) completions late: a parity wait would now block forever
        # (or, one later, pass for the wrong phase)
        raise Hazard(f"aliasing: intended completion {intended}, barrier at {bar.completed}")
    return (bar.completed & 1) != parity


def simulate(steps, n_epi, rng, shared_dready=False, max_ticks=200000):
    """steps: list of (tiles, chunks_per_tile).  Returns True when every agent finished."""
    full = [MBar(1) for _ in range(STAGES)]
    sfree = [MBar(LOADERS) for _ in range(STAGES)]
    afull = [MBar(LOADERS) for _ in range(ASTAGES)]
    adone = [MBar(1) for _ in range(ASTAGES)]
    ntile_bars = 1 if shared_dready else 4
    dready = [MBar(1) for _ in range(ntile_bars)]
    opready = MBar(n_epi)
    pending_events = []  # (due_tick, fn): asynchronous completions (bulk copies, MMA commits)
    total_chunks = sum(t * c for t, c in steps)

    def producer():
        n = 0
        while n < total_chunks:
            stage, use = n % STAGES, n // STAGES
            if use > 0:
                while not wait_ok(sfree[stage], use):
                    yield
            pending_events.append((tick[0] + rng.randint(1, 30), full[stage].arrive))
            n += 1
            yield

    def loader():
        for n in range(total_chunks):
            ss, suse = n % STAGES, n // STAGES
            ts, tuse = n % ASTAGES, n // ASTAGES
            while not wait_ok(full[ss], suse + 1):
                yield
            for _ in range(rng.randint(0, 3)):  # shared-memory loads + split
                yield
            sfree[ss].arrive()
            if tuse > 0:
                while not wait_ok(adone[ts], tuse):
                    yield
            for _ in range(rng.randint(0, 3)):  # tcgen05.st + wait::st
                yield
            afull[ts].arrive()
            yield

    def mma():
        n = 0
        for s, (tiles, chunks) in enumerate(steps):
            while not wait_ok(opready, s + 1):
                yield
            for t in range(tiles):
                for _ in range(chunks):
                    ts, tuse = n % ASTAGES, n // ASTAGES
                    while not wait_ok(afull[ts], tuse + 1):
                        yield
                    pending_events.append((tick[0] + rng.randint(1, 12), adone[ts].arrive))
                    n += 1
                    yield
                bar = dready[0 if shared_dready else t]
                # tcgen05.commit tracks ALL earlier MMAs: completes after the last one above
                pending_events.append((tick[0] + rng.randint(12, 20), bar.arrive))
                yield

    def epilogue(speed):
        uses = [0] * ntile_bars
        opready.arrive()  # operand of step 0 (the X tile)
        for s, (tiles, chunks) in enumerate(steps):
            for t in range(tiles):
                b = 0 if shared_dready else t
                uses[b] += 1
                while not wait_ok(dready[b], uses[b]):
                    yield
                for _ in range(rng.randint(1, speed)):  # epilogue work of this tile
                    yield
            if s + 1 < len(steps):
                opready.arrive()
            yield

    tick = [0]
    agents = ([producer(), mma()] + [loader() for _ in range(LOADERS)]
              + [epilogue(rng.choice([3, 10, 40])) for _ in range(n_epi)])
    alive = list(agents)
    while alive and tick[0] < max_ticks:
        tick[0] += 1
        due = [e for e in pending_events if e[0] <= tick[0]]
        for e in due:
            pending_events.remove(e)
            e[1]()
        a = rng.choice(alive)
        try:
            next(a)
        except StopIteration:
            alive.remove(a)
    return not alive


def _random_steps(rng):
    n = rng.randint(2, 15)
    return [(rng.randint(1, 4), rng.randint(1, 8)) for _ in range(n)]


def test_protocol_makes_progress_and_never_aliases():
    rng = random.Random(1234)
    for trial in range(150):
        steps = _random_steps(rng)
        assert simulate(steps, n_epi=rng.randint(1, 4), rng=rng), (trial, steps)
    # the shapes of BASELINE config 2: 3 passes x (2x4, 1x8, 1x4 chunks) + backward (1x1, 2x4)
    cfg2 = [(2, 4), (1, 8), (1, 4)] * 3 + [(1, 1), (2, 4)]
    for _ in range(20):
        assert simulate(cfg2, n_epi=4, rng=rng)


def test_model_catches_the_shared_accumulator_barrier_bug():
    """With ONE dready barrier for all tiles of a step (the first version of the kernel) a slow
    epilogue falls two completions behind as soon as a step has >= 3 short tiles."""
    rng = random.Random(7)
    caught = 0
    for _ in range(60):
        steps = [(4, 1)] * 6  # many short tiles per step: the MMA warp runs far ahead
        try:
            finished = simulate(steps, n_epi=2, rng=rng, shared_dready=True, max_ticks=20000)
            caught += 0 if finished else 1
        except Hazard:
            caught += 1
    assert caught > 0

"""Model check of the mbarrier protocol of the tcgen05 TD kernel (reagent_b200/csrc/rb200_dqn_tc.cu).

The kernel's step loop is synchronised only by mbarriers that are waited on by PARITY:
  full[s] / sfree[s]  shared-memory weight ring (producer warp <-> the 4 loader warps of ONE group)
  afull[t] / adone[t] tensor-memory weight ring (those 4 loader warps <-> MMA warp)
Chunk i of the weight stream belongs to loader group i % GROUPS (groups convert consecutive
chunks concurrently); every loader warp tracks the stage / parity sequence of ALL chunks.
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

LOADERS = 4  # loader warps of one group: one arrival each on sfree / afull


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
    if intended >= bar.completed + 2 and (bar.completed & 1) != parity:
        # the waiter is two phases EARLY: the parity test passes on a completion that is two
        # behind the one it wants
        raise Hazard(f"early pass: intended completion {intended}, barrier at {bar.completed}")
    return (bar.completed & 1) != parity


def simulate(steps, n_epi, rng, shared_dready=False, max_ticks=400000, groups=1, astages=3, stages=3):
    """steps: list of (tiles, chunks_per_tile); groups = kQLoaderGroups, astages = QDev.a_stages
    (2..7, whatever the accumulators leave; a stage holds kQSub chunks).  Returns True when every agent finished."""
    ASTAGES = astages
    STAGES = stages
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

    def loader(group):
        for n in range(total_chunks):
            if n % groups != group:
                continue
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

    commit_due = [0]  # tcgen05.commit completions happen in issue order

    def commit_at(lo, hi):
        commit_due[0] = max(commit_due[0], tick[0] + rng.randint(lo, hi))
        return commit_due[0]

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
                    pending_events.append((commit_at(1, 12), adone[ts].arrive))
                    n += 1
                    yield
                bar = dready[0 if shared_dready else t]
                # tcgen05.commit tracks ALL earlier MMAs: completes after the last one above
                pending_events.append((commit_at(1, 20), bar.arrive))
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
    agents = ([producer(), mma()] + [loader(g) for g in range(groups) for _ in range(LOADERS)]
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
        # the kernel requires both ring depths to be multiples of the number of loader groups, so
        # that a ring stage is only ever used by ONE group: bulk copies may land out of order, and
        # a group that could reach a stage two uses early would pass its parity wait on the wrong
        # completion (see the tests below)
        g = rng.randint(1, 3)
        astages = g * rng.randint(1 if g > 1 else 2, 7 // g)
        stages = g * rng.randint(1 if g > 1 else 2, 6 // g)
        assert simulate(steps, n_epi=rng.randint(1, 4), rng=rng, groups=g, astages=astages,
                        stages=stages), (trial, steps)
    # the shapes of BASELINE config 2: 3 passes x (2x4, 1x8, 1x4 chunks) + backward (1x1, 2x4)
    cfg2 = [(2, 4), (1, 8), (1, 4)] * 3 + [(1, 1), (2, 4)]
    for _ in range(20):
        assert simulate(cfg2, n_epi=4, rng=rng, groups=1, astages=3, stages=3)
    # the same network in 64-k stages: (2x2, 1x4, 1x2) x 3 + (1x1, 2x2)
    cfg2s = [(2, 2), (1, 4), (1, 2)] * 3 + [(1, 1), (2, 2)]
    for _ in range(20):
        assert simulate(cfg2s, n_epi=4, rng=rng, groups=1, astages=3, stages=3)


def test_model_catches_more_loader_groups_than_ring_stages():
    """With more loader groups than tensor-memory stages a group can reach a stage two uses
    early, where a parity wait passes on the wrong completion."""
    rng = random.Random(11)
    caught = 0
    for _ in range(40):
        try:
            simulate([(2, 6)] * 6, n_epi=2, rng=rng, groups=4, astages=2, stages=4, max_ticks=40000)
        except Hazard:
            caught += 1
    assert caught > 0


def test_model_catches_ring_stages_shared_between_loader_groups():
    """Two loader groups on a 3-stage shared-memory ring: copies land out of order, one group
    runs ahead onto a stage whose previous use (the other group's) has not landed yet."""
    rng = random.Random(5)
    caught = 0
    for _ in range(60):
        try:
            simulate([(2, 6)] * 6, n_epi=2, rng=rng, groups=2, astages=4, stages=3, max_ticks=40000)
        except Hazard:
            caught += 1
    assert caught > 0


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

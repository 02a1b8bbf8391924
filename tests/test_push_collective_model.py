"""Model check of the push all-reduce protocol of csrc/pinnjet_comm.cu (reduce_allreduce_kernel), on the CPU.

The kernel's cross-rank protocol is small enough to state as a model: per rank a symmetric buffer with two slots (epoch
parity) of one 64-bit word {epoch, value} per (source rank, element); a call with epoch e PUSHES its words into slot e & 1
of every peer, then POLLS its own slot until every peer's word carries e, then sums in rank order.  Calls of one rank are
stream ordered; ranks run at arbitrary relative speeds.  The model executes random interleavings of those steps (including
a rank racing up to a full call ahead of a slow peer, which is as far as the data dependency lets it get) and checks the two
claims the design rests on: every sum is the sum of the SAME epoch's values (no stale or future word is ever accepted), and a
word is never overwritten before its reader has consumed it (slot reuse after two epochs is safe without a barrier)."""
import random

import pytest


class Rank:
    def __init__(self, r, world, n):
        self.r, self.world, self.n = r, world, n
        self.slots = [[[(0, None)] * n for _ in range(world)] for _ in range(2)]   # [parity][src][i] = (epoch, value)
        self.epoch = 0           # completed calls
        self.phase = "idle"      # idle -> pushing -> polling -> idle
        self.cursor = 0
        self.results = []


def value(rank, epoch, i):
    return (rank + 1) * 1000.0 + epoch * 10.0 + i * 0.25


def run(world, n, calls, seed):
    rng = random.Random(seed)
    ranks = [Rank(r, world, n) for r in range(world)]
    consumed = {}                # (dst, parity, src, i) -> epoch last consumed by dst
    while any(rk.epoch < calls for rk in ranks):
        rk = rng.choice([x for x in ranks if x.epoch < calls])
        e = rk.epoch + 1
        if rk.phase == "idle":
            rk.phase, rk.cursor = "pushing", 0
        elif rk.phase == "pushing":                      # one (peer, element) store per step, in any order the kernel may issue
            todo = [(p, i) for p in range(world) if p != rk.r for i in range(n)]
            p, i = todo[rk.cursor]
            dst = ranks[p]
            old_epoch, _ = dst.slots[e & 1][rk.r][i]
            # the word being overwritten (epoch e - 2 or the initial 0) must already have been consumed by its reader
            assert old_epoch == 0 or consumed.get((p, e & 1, rk.r, i)) == old_epoch, \
                f"rank {rk.r} overwrites epoch {old_epoch} in rank {p}'s slot before it was read"
            assert old_epoch in (0, e - 2)
            dst.slots[e & 1][rk.r][i] = (e, value(rk.r, e, i))
            rk.cursor += 1
            if rk.cursor == len(todo):
                rk.phase, rk.cursor = "polling", 0
        elif rk.phase == "polling":                      # element by element: all peers' words must carry epoch e
            i = rk.cursor
            words = [rk.slots[e & 1][p][i] for p in range(world) if p != rk.r]
            assert all(w[0] <= e for w in words), "a future epoch is visible in the slot being polled"
            if all(w[0] == e for w in words):
                acc = 0.0
                for p in range(world):                   # rank order; own value from the register
                    acc += value(rk.r, e, i) if p == rk.r else rk.slots[e & 1][p][i][1]
                    if p != rk.r:
                        consumed[(rk.r, e & 1, p, i)] = e
                assert acc == sum(value(p, e, i) for p in range(world))
                rk.cursor += 1
                if rk.cursor == n:
                    rk.results.append(e)
                    rk.epoch, rk.phase = e, "idle"
            # else: spin (another rank gets scheduled)
    return ranks


@pytest.mark.parametrize("world", [2, 3, 8])
def test_random_interleavings_keep_epochs_apart(world):
    for seed in range(20):
        ranks = run(world, n=3, calls=6, seed=seed)
        assert all(rk.results == list(range(1, 7)) for rk in ranks)


def test_a_fast_rank_cannot_run_two_epochs_ahead():
    """The slot-reuse argument: a rank can only START epoch e + 2 after finishing e + 1, which needs every peer's e + 1 words,
    which a peer sends only after its epoch-e call completed (stream order) -- so the words of e are consumed by then."""
    world, n = 3, 2
    ranks = [Rank(r, world, n) for r in range(world)]
    # rank 0 pushes epoch 1 to everybody and tries to complete: it cannot, the peers have not pushed yet
    for p in (1, 2):
        for i in range(n):
            ranks[p].slots[1][0][i] = (1, value(0, 1, i))
    assert any(ranks[0].slots[1][p][i][0] != 1 for p in (1, 2) for i in range(n))   # polling would spin: epoch 1 not complete

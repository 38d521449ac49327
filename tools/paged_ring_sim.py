"""Discrete simulation of the paged decode kernel's shared-memory ring protocol (one CTA: a producer thread and the
consumer warps, mbarrier phase/parity semantics, TMA tiles landing OUT OF ORDER).

Tile g of a CTA's stream is loaded into stage g % NS and consumed by warp g % NW; a consumer passes
`mbarrier.try_wait.parity(P)` as soon as the barrier's current phase parity differs from P.  If NS is not a multiple
of NW, successive fills of a stage belong to different warps, and a warp that runs ahead can pass the wait for fill k
while fill k-1 is still in flight (the parity test cannot tell "phase k done" from "phase k-1 not done yet") and read
stale data.  With NS % NW == 0 the same warp owns every fill of a stage and reaches fill k only after consuming k-1.
This is the rule asserted in csrc/kernels/kernel_params.h (PagedCfg); tests/test_ring_protocol.py runs it.

usage: python tools/paged_ring_sim.py NS tiles_per_item,... [max_landing_delay]
"""
import random
import sys


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):          # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


def simulate(NS, items, max_delay=6, NW=4, seed=0, max_steps=2_000_000):
    """Returns (ok, detail).  ok is False when a consumer read a stage holding the wrong tile, or on deadlock."""
    rng = random.Random(seed)
    full = [Bar(1) for _ in range(NS)]
    empty = [Bar(1) for _ in range(NS)]
    content = [None] * NS
    inflight = []
    err = []

    def producer():
        n = 0
        for nt in items:
            for i in range(nt):
                g = n + i
                s = g % NS
                if g >= NS:
                    while not empty[s].test(((g // NS) - 1) & 1):
                        yield
                inflight.append([g, s, rng.randint(1, max_delay)])
                yield
            n += nt

    def consumer(w):
        n = 0
        for nt in items:
            i = (w + NW - (n % NW)) % NW
            while i < nt:
                g = n + i
                s = g % NS
                while not full[s].test((g // NS) & 1):
                    yield
                if content[s] != g:
                    err.append((w, g, s, content[s]))
                    return
                for _ in range(rng.randint(0, 3)):
                    yield
                empty[s].arrive()
                i += NW
            n += nt
            for _ in range(rng.randint(0, 8)):
                yield

    procs = [producer()] + [consumer(w) for w in range(NW)]
    alive = [True] * len(procs)
    steps = 0
    while any(alive) and not err:
        steps += 1
        if steps > max_steps:
            return False, "deadlock"
        for t in inflight[:]:
            t[2] -= 1
            if t[2] <= 0:
                content[t[1]] = t[0]
                full[t[1]].arrive()
                inflight.remove(t)
        order = list(range(len(procs)))
        rng.shuffle(order)
        for k in order:
            if alive[k]:
                try:
                    next(procs[k])
                except StopIteration:
                    alive[k] = False
    if err:
        w, g, s, c = err[0]
        return False, f"warp {w} read stage {s} for tile {g} but it held tile {c}"
    return True, f"{steps} steps"


if __name__ == "__main__":
    ns = int(sys.argv[1])
    its = [int(x) for x in sys.argv[2].split(",")]
    delay = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    print(simulate(ns, its, delay))

"""Model check (CPU, pure Python) of the lock-free minimum that the Boruvka pixel kernel uses
(csrc/dofs_seg.cuh: k_bor_pixel, pick_offer, bor_pixel_settle).

Every pixel offers the (prefix, slot)-smallest of its outgoing edges to its component with ONE native atomicMin on
(prefix << 32 | slot); the reference order is (weight, slot), which differs when two offers share a prefix but not a
weight.  The protocol repairs that pairwise: an offer that meets another edge of its own prefix — in the value it read or
in the value its atomicMin returned — offers itself under the exact order with a CAS loop and re-offers what its
atomicMin displaced; a pixel whose own edges share the smallest prefix offers them all exactly.  This test runs the
protocol as interleaved per-thread step sequences under a random scheduler and checks that the word always ends as the
exact minimum.  It models the algorithm, not the CUDA code: the GPU parity tests exercise the kernel itself."""
import random

NONE = (0xFFFFFFFF, 0xFFFFFFFF)  # (prefix, slot) of "no pick": larger than any real pick in L order


class Edge:
    def __init__(self, prefix, weight, slot):
        self.prefix, self.weight, self.slot = prefix, weight, slot

    @property
    def L(self):  # what the atomicMin compares
        return (self.prefix, self.slot)

    @property
    def E(self):  # the reference's order
        return (self.prefix, self.weight, self.slot)


def exact_less(a, b):
    """pick_less: a, b are Edge or None (= no pick)."""
    if a is None:
        return False
    if b is None:
        return True
    if a.prefix != b.prefix:
        return a.prefix < b.prefix
    if a.slot == b.slot:
        return False
    if a.prefix == 0:
        return a.slot < b.slot  # prefix 0 is the weight 0 exactly
    return a.E < b.E


def meets(cand, seen):
    return seen is not None and seen.prefix == cand.prefix and seen.slot != cand.slot and cand.prefix != 0


def exact_offer(word, cand):
    """pick_offer: CAS loop; yields before every access to the shared word."""
    yield
    cur = word[0]
    while exact_less(cand, cur):
        yield
        if word[0] is cur:  # CAS succeeds
            word[0] = cand
            return
        cur = word[0]  # CAS failed: it returned the current value


def pixel_thread(word, edges):
    """One pixel of k_bor_pixel with its outgoing edges."""
    mine = min(edges, key=lambda e: e.L)
    local_tie = mine.prefix != 0 and any(e is not mine and e.prefix == mine.prefix for e in edges)
    yield
    seen = word[0]  # the (possibly stale by the time it is used) read of best[cp]
    sent = False
    if seen is None or mine.L < seen.L:
        yield
        old = word[0]  # atomicMin returns the old value
        if old is None or mine.L < old.L:
            word[0] = mine
        seen, sent = old, True
    if meets(mine, seen):
        yield from exact_offer(word, mine)
        if sent and mine.L < seen.L:  # the atomicMin displaced `seen`
            yield from exact_offer(word, seen)
    if local_tie:
        for e in edges:
            yield from exact_offer(word, e)


def run(pixels, rng):
    word = [None]
    threads = [pixel_thread(word, edges) for edges in pixels]
    live = list(range(len(threads)))
    while live:
        i = rng.choice(live)
        try:
            next(threads[i])
        except StopIteration:
            live.remove(i)
    return word[0]


def random_component(rng):
    """A component with a few boundary pixels; prefixes and weights drawn from tiny ranges so that ties are the norm."""
    n_pixels = rng.randint(1, 6)
    slots = rng.sample(range(1000), 4 * n_pixels)
    pixels, k = [], 0
    for _ in range(n_pixels):
        edges = []
        for _ in range(rng.randint(1, 4)):
            prefix = rng.choice([0, 5, 5, 5, 7])
            weight = 0.0 if prefix == 0 else prefix + rng.choice([0.0, 0.1, 0.1, 0.2, 0.3])
            edges.append(Edge(prefix, weight, slots[k]))
            k += 1
        pixels.append(edges)
    return pixels


def test_exact_minimum_under_random_interleavings():
    rng = random.Random(1234)
    for trial in range(4000):
        pixels = random_component(rng)
        want = min((e for edges in pixels for e in edges), key=lambda e: e.E)
        for _ in range(3):  # several schedules of the same component
            got = run(pixels, rng)
            assert got is want, (trial, [(e.prefix, e.weight, e.slot) for edges in pixels for e in edges],
                                 (got.prefix, got.weight, got.slot), (want.prefix, want.weight, want.slot))


def test_atomic_min_alone_is_not_enough():
    """The counter-example the settlement exists for: equal prefix, the later slot is lighter."""
    a, b = Edge(5, 5.2, 10), Edge(5, 5.1, 20)
    word = [None]
    for e in (a, b):  # atomicMin only
        if word[0] is None or e.L < word[0].L:
            word[0] = e
    assert word[0] is a and min((a, b), key=lambda e: e.E) is b
    for order in ([[a], [b]], [[b], [a]]):
        assert run(order, random.Random(0)) is b

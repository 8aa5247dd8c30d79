"""CPU: the dealing rule of collide_warp_balanced (csrc/bcg_device.cuh), restated in NumPy and checked as a property.

Every lane of a warp lists the non-empty tiles under its env's footprint (a bit mask, 4 bits per band); the warp numbers
its (lane, tile) pairs by an inclusive prefix sum and, round after round, lane l takes pair (lanes present) * round + l:
it finds the pair's owner by a binary search over the inclusive counts (probes beyond the last lane present count as
+infinity) and the tile as the r-th set bit of the owner's mask.  Property: for any masks and any number of lanes present
(the tail warp of a batch), every pair is dealt exactly once and to nobody twice -- the case a first version got wrong for
partly filled warps."""
import numpy as np
import pytest


def _deal(masks):
    """masks: one tile mask per lane PRESENT (lanes 0 .. len - 1).  Returns the (owner, bit) pairs in dealing order and
    the number of rounds, following the kernel step by step."""
    nact = len(masks)
    last = nact - 1
    cnt = [bin(m).count("1") for m in masks]
    incl = list(np.cumsum(cnt))
    total = incl[last]
    dealt, rounds = [], 0
    base = 0
    while base < total:
        rounds += 1
        for lane in range(nact):
            i = base + lane
            if i >= total:
                continue
            o = 0
            st = 16
            while st >= 1:                                   # the kernel's unrolled binary search
                probe = o + st - 1
                t = incl[min(probe, last)]
                if probe <= last and t <= i:
                    o += st
                st >>= 1
            o = min(o, last)
            r = i - (incl[o] - cnt[o])
            m = masks[o]
            for _ in range(r):
                m &= m - 1
            bit = (m & -m).bit_length() - 1
            dealt.append((o, bit))
        base += last + 1
    return dealt, rounds


@pytest.mark.parametrize("nact", [1, 2, 5, 17, 31, 32])
def test_every_pair_is_dealt_exactly_once(nact):
    rng = np.random.RandomState(nact)
    for trial in range(300):
        kind = trial % 4
        if kind == 0:
            masks = [int(rng.randint(0, 1 << 24)) for _ in range(nact)]
        elif kind == 1:                                       # mostly free space: a few lanes hold tiles
            masks = [int(rng.randint(0, 1 << 24)) if rng.rand() < 0.2 else 0 for _ in range(nact)]
        elif kind == 2:                                       # Monte-Carlo fan-out: every lane the same list
            masks = [int(rng.randint(1, 1 << 12))] * nact
        else:                                                 # nothing under any footprint
            masks = [0] * nact
        dealt, rounds = _deal(masks)
        want = sorted((lane, b) for lane, m in enumerate(masks) for b in range(24) if (m >> b) & 1)
        assert sorted(dealt) == want
        assert rounds == -(-len(want) // nact)                # ceil: no lane idles before the last round


def test_the_choice_between_dealing_and_the_per_thread_loop():
    """collide_warp_balanced deals only when 3 * rounds < 2 * longest list: equal lists (nothing to balance) keep the
    per-thread loop, a warp with one long list among short ones deals."""
    def deals(masks):
        cnt = [bin(m).count("1") for m in masks]
        rounds = -(-sum(cnt) // len(masks))
        return 3 * rounds < 2 * max(cnt)
    assert not deals([0b111] * 32)                            # 3 rounds against 3 iterations
    assert deals([0b111111111] + [0b1] * 31)                  # 2 rounds against 9 iterations
    assert not deals([0] * 32)

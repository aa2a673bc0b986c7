"""The lattice kernel's work items (CPU restatement of the host / device logic): however the (ego, horizon) pairs are cut
into items -- whole items of `slots` pairs, or FISS_BIG_FRAC percent of the pairs in whole items and the rest one pair per item
(work drawn through the device counter) -- every pair must belong to exactly one item, the items must be decodable from their
number alone, and a chained launch's bookkeeping (sequence numbers with 0 reserved, alternating counter pairs) must hold across
the 32-bit wrap.  Mirrors fiss_abi.cu eval_grid / launch_grid and fiss_grid_kernel.cuh item_decode."""
import itertools

import numpy as np


def host_items(n_pairs, slots, dynamic, big_pct):
    """(n_big, items) as eval_grid computes them for an unchunked lattice."""
    full_items = (n_pairs + slots - 1) // slots
    split = dynamic and big_pct < 100
    n_big = n_pairs * big_pct // 100 // slots if split else full_items
    items = n_big + (n_pairs - n_big * slots) if split else n_big
    return n_big, items


def item_decode(item, n_big, slots, n_pairs):
    """(first pair, pairs) of an item: fiss_grid_kernel.cuh item_decode, n_chunks == 1."""
    if item < n_big:
        bk0 = item * slots
        return bk0, min(slots, n_pairs - bk0)
    return n_big * slots + (item - n_big), 1


def test_every_pair_belongs_to_exactly_one_item():
    for n_pairs, slots, (dynamic, pct) in itertools.product(
            (1, 2, 5, 296, 297, 2560, 2561, 20480), (1, 2, 3, 4), ((False, 70), (True, 100), (True, 70), (True, 50), (True, 0), (True, 1))):
        n_big, items = host_items(n_pairs, slots, dynamic, pct)
        seen = np.zeros(n_pairs, dtype=np.int32)
        for item in range(items):
            bk0, gv = item_decode(item, n_big, slots, n_pairs)
            assert 1 <= gv <= slots and bk0 + gv <= n_pairs, (n_pairs, slots, dynamic, pct, item)
            seen[bk0:bk0 + gv] += 1
        assert (seen == 1).all(), (n_pairs, slots, dynamic, pct)
        # big items first, single pairs last: the CTAs run out of work together
        sizes = [item_decode(i, n_big, slots, n_pairs)[1] for i in range(items)]
        assert sizes[:max(n_big - 1, 0)] == [slots] * max(n_big - 1, 0)


def test_sequence_numbers_skip_zero_and_order_across_the_wrap():
    """launch_grid numbers the directly issued launches; 0 means "not numbered" (a launch inside a graph); the kernel's gate
    `(int32)(done - seq_prev) < 0` must read "the predecessor is not over yet" for every pair of neighbours."""
    seq = np.uint32(0xfffffffd)
    issued = []
    for _ in range(6):
        prev = seq
        seq = np.uint32((int(seq) + 1) & 0xffffffff)
        if seq == 0:
            seq = np.uint32(1)
        issued.append((int(prev), int(seq)))
    assert [s for _, s in issued] == [0xfffffffe, 0xffffffff, 1, 2, 3, 4]
    for prev, cur in issued:
        for done, over in ((prev, True), (cur, True), ((prev - 1) & 0xffffffff, False)):
            waiting = np.int32(np.uint32((done - prev) & 0xffffffff)) < 0
            assert waiting == (not over), (prev, cur, done)
    # the very first launch of a handle: nothing issued before (seq_prev = 0), nothing published (done = 0): no wait
    assert not (np.int32(np.uint32(0)) < 0)


def test_counter_pairs_alternate_between_direct_launches_only():
    """[0..1] / [2..3] alternate between directly issued launches (a chained launch draws beside its predecessor's last
    items); launches recorded for a graph use [6..7] and leave the alternation alone."""
    parity, used = 0, []
    for recorded in (False, False, True, False, True, True, False):
        if recorded:
            used.append(6)
        else:
            parity ^= 1
            used.append(2 * parity)
    direct = [u for u in used if u != 6]
    assert all(a != b for a, b in zip(direct, direct[1:]))
    assert set(used) <= {0, 2, 6}

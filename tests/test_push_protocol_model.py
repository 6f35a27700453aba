"""Model check of the staging-buffer protocol of the fused wgrad-GEMM -> reduce-scatter path (``parallel/engine.py``
``_prepare_grads / _reduce / _on_reduce_barrier``, ``parallel/fused_comm.py::reduce_pushed``).

Every rank runs two in-order streams:
  compute:  for unit j:  [wait free(X[j % D])]  push_j  (writes slot `me` of X[j % D] on EVERY rank)
  reduce :  for unit j:  [wait compute finished push_j]  barrier_j  sum_j (reads all slots of MY X[j % D])
A buffer is declared free for unit j by the completion of barrier_{j-D+1} on THIS rank's reduce stream (the engine records
the buffer's ``free_event`` right after the next flag round that follows its slot sum).  The claim: whatever the relative
speed of the ranks, no push into X ever lands on a rank that has not finished summing X's previous contents, and every
sum sees exactly the pushes of its own unit.  The simulator executes random interleavings of the 2 * W stream heads and
checks both properties; it also shows that freeing a buffer one flag round too early (the bug the rule prevents) is caught.
"""
import random

import pytest


def simulate(W, D, units, seed, free_after):
    """free_after(j) -> index of the barrier whose completion on the pushing rank frees the buffer for unit j."""
    rng = random.Random(seed)
    # per rank progress
    pushed = [-1] * W          # last unit whose push this rank has executed
    posted = [-1] * W          # last barrier this rank has posted (= entered)
    passed = [-1] * W          # last barrier this rank has passed
    summed = [-1] * W          # last unit this rank has summed
    # content[r][b][s] = unit whose data sits in slot s of buffer b on rank r
    content = [[[None] * W for _ in range(D)] for _ in range(W)]
    steps = 0
    while min(summed) < units - 1:
        steps += 1
        assert steps < 200000, "deadlock"
        r = rng.randrange(W)
        if rng.random() < 0.5:
            # ---- compute stream of rank r: next push
            j = pushed[r] + 1
            if j >= units:
                continue
            need = free_after(j)
            if need >= 0 and passed[r] < need:
                continue                                   # buffer not declared free yet
            b = j % D
            for o in range(W):                             # my tile lands in slot r of every owner's buffer b
                prev = content[o][b][r]
                if prev is not None:
                    assert summed[o] >= prev, f"rank {r} overwrote unit {prev} in buffer {b} of rank {o} before it was summed"
                content[o][b][r] = j
            pushed[r] = j
        else:
            # ---- reduce stream of rank r: post barrier / pass barrier / sum, in order
            j = summed[r] + 1
            if j >= units:
                continue
            if posted[r] < j:
                if pushed[r] >= j:                         # reduce stream waits for the compute stream (event)
                    posted[r] = j
                continue
            if passed[r] < j:
                if all(p >= j for p in posted):            # flag round: everybody has posted barrier j
                    passed[r] = j
                continue
            b = j % D
            assert all(content[r][b][s] == j for s in range(W)), f"rank {r} summed unit {j} with slots {content[r][b]}"
            summed[r] = j
    return steps


@pytest.mark.parametrize("W,D", [(2, 2), (2, 3), (4, 3), (8, 3), (8, 2), (3, 4)])
def test_buffer_reuse_rule_is_safe_under_any_interleaving(W, D):
    for seed in range(30):
        simulate(W, D, units=14, seed=seed, free_after=lambda j: j - D + 1 if j >= D else -1)


def test_freeing_one_flag_round_early_is_caught():
    """With the buffer declared free after barrier_{j-D} (i.e. as soon as ITS OWN flag round completed, before the peers'
    slot sums), some interleaving lets a fast rank overwrite data a slow rank has not summed yet."""
    caught = 0
    for seed in range(200):
        try:
            simulate(4, 2, units=10, seed=seed, free_after=lambda j: j - 2 if j >= 2 else -1)
        except AssertionError as e:
            if "overwrote" in str(e) or "summed unit" in str(e):
                caught += 1
    assert caught > 0

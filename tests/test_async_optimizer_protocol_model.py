"""Model check of the asynchronous SHARDED optimizer (``parallel/optim.py::ShardedAdamW.step``,
``parallel/engine.py::_wait_unit_updated``, ``parallel/fused_comm.py::post_unit_updated / wait_unit_updated``).

With parameters sharded over W ranks, the AdamW update of step t runs on a side stream under the forward of step t+1, and the
step barrier ("everybody has updated everything") is replaced by one flag per (rank, unit).  Every rank runs three in-order
streams; per step t (units 0..U-1, forward order = index order):

  compute:  for u ascending:   [wait own updated(u, t-1)] [wait flags(u, t-1) of ALL ranks]  gather_begin(u) ... gather_end(u)
            for u descending:  re-gather of u for backward (same step: no second wait), then push wgrad(u)
  reduce :  for u descending:  [wait compute pushed (u, t)]  flag round(u, t)  (passes when all ranks posted)  sum(u, t)
  opt    :  [wait compute finished backward t AND reduce finished step t]  for u ascending: write shard(u) := version t, post flag(u, t)

A gather of u in step t reads the shard of u on every rank over NVLink for a while (begin .. end).  Claims checked under random
interleavings of the 3 * W stream heads:
  (RAW) every shard a gather reads carries version t-1 exactly -- never the pre-update t-2;
  (WAR) no rank writes version t into a shard while some rank's step-t gather of it is still in flight.
Dropping the cross-rank flag wait (keeping only the rank-local event) must be caught as a RAW violation.
"""
import random

import pytest


def simulate(W, U, T, seed, wait_peer_flags=True):
    rng = random.Random(seed)
    version = [[-1] * U for _ in range(W)]            # version[r][u]: last step whose update rank r applied to its shard of u
    flag = [[-1] * U for _ in range(W)]               # flag[r][u]: last step rank r announced for unit u
    readers = [[0] * U for _ in range(W)]             # in-flight gathers reading rank r's shard of u
    # compute-stream programme per step: ("gb", u) ("ge", u) for forward, then for backward ("gb", u) ("ge", u) ("push", u)
    prog = []
    for u in range(U):
        prog += [("gb", u, True), ("ge", u, True)]
    for u in reversed(range(U)):
        prog += [("gb", u, False), ("ge", u, False), ("push", u, False)]
    pc = [0] * W                                      # index into prog * T
    pushed = [[-1] * U for _ in range(W)]             # pushed[r][u]: last step whose wgrad of u rank r pushed
    posted = [[-1] * U for _ in range(W)]             # reduce stream: flag round (u, t) entered
    summed = [[-1] * U for _ in range(W)]
    red_pc = [0] * W                                  # index into (T x reversed units)
    opt_pc = [0] * W                                  # index into (T x units)
    order_b = list(reversed(range(U)))
    total_c, total_r, total_o = len(prog) * T, U * T, U * T
    steps = 0
    while min(opt_pc) < total_o or min(pc) < total_c:
        steps += 1
        assert steps < 2_000_000, "deadlock"
        r = rng.randrange(W)
        which = rng.randrange(3)
        if which == 0 and pc[r] < total_c:
            t, (op, u, fwd) = pc[r] // len(prog), prog[pc[r] % len(prog)]
            if op == "gb":
                if fwd and t > 0:
                    if version[r][u] < t - 1:
                        continue                                        # own ev_updated not reached yet
                    if wait_peer_flags and any(flag[o][u] < t - 1 for o in range(W)):
                        continue                                        # a peer has not announced its update of u
                for o in range(W):
                    assert version[o][u] >= t - 1, f"RAW: rank {r} step {t} gathers unit {u} from rank {o} at version {version[o][u]}"
                    assert version[o][u] <= t - 1, f"WAR: rank {r} step {t} gathers unit {u} from rank {o} already at version {version[o][u]}"
                    readers[o][u] += 1
            elif op == "ge":
                for o in range(W):
                    readers[o][u] -= 1
            else:
                pushed[r][u] = t
            pc[r] += 1
        elif which == 1 and red_pc[r] < total_r:
            t, u = red_pc[r] // U, order_b[red_pc[r] % U]
            if posted[r][u] < t:
                if pushed[r][u] >= t:
                    posted[r][u] = t
                continue
            if all(posted[o][u] >= t for o in range(W)):
                summed[r][u] = t
                red_pc[r] += 1
        elif which == 2 and opt_pc[r] < total_o:
            t, u = opt_pc[r] // U, opt_pc[r] % U
            # s_opt.wait_stream(s_compute) at optimizer.step(): backward of step t is done on this rank, and the compute
            # stream had waited for the reduce stream (all units of step t summed)
            if pc[r] < (t + 1) * len(prog) or red_pc[r] < (t + 1) * U:
                continue
            assert readers[r][u] == 0, f"WAR: rank {r} updates unit {u} to version {t} under {readers[r][u]} in-flight gathers"
            version[r][u] = t
            flag[r][u] = t
            opt_pc[r] += 1
    return steps


@pytest.mark.parametrize("W,U", [(2, 3), (4, 4), (8, 5), (3, 2)])
def test_per_unit_flags_replace_the_step_barrier_safely(W, U):
    for seed in range(25):
        simulate(W, U, T=4, seed=seed)


def test_without_the_cross_rank_flags_a_stale_shard_is_gathered():
    caught = 0
    for seed in range(100):
        try:
            simulate(4, 3, T=3, seed=seed, wait_peer_flags=False)
        except AssertionError as e:
            assert "RAW" in str(e), e
            caught += 1
    assert caught > 0

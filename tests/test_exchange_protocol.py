"""Model check of the slab engine's DIRECT plane exchange (imhd-cuda_b200/csrc/imhd_slabs.cu, ring_exchange): a slab writes its
plane straight into the neighbour's receive buffer and then a sequence number into the neighbour's arrival counter; the
receiver's stream waits for `counter >= n`.  There is no rendezvous, so a neighbour may be one exchange ahead -- it can post
exchange n+1 as soon as this slab has POSTED n, before this slab has consumed n.  The engine's buffer discipline:

  * ghost planes land in receive STAGING planes double-buffered by the parity of the exchange count (recv_stage);
  * predictor planes land in plane SET `which`, toggled at every predictor exchange of the overlapped schedule; a set is
    read by the kernels of the NEXT step, as late as that step's interior launch ends.

This test runs the per-slab programs under random interleavings (copies land at once: the worst case for overwriting) and
checks that every consumer reads exactly the plane it is meant to read.  It also shows the model has teeth: with single
receive staging, two consecutive ghost exchanges (imhd_ctx_set_state_local + refresh, then prime) DO get overwritten.
CPU only; the bits of the real exchange are checked on GPUs by tests/test_gpu_multi.py and tools/check_slab_engine.py.
"""
import random

import pytest


class Slab:
    def __init__(self, rank, world, staging_buffers):
        self.rank, self.world = rank, world
        self.up, self.down = (rank + 1) % world, (rank - 1) % world
        self.counter = [0, 0]            # [0] up-going message arrived (from below), [1] down-going (from above)
        self.staging = [[None, None] for _ in range(staging_buffers)]   # [parity][from below, from above]
        self.sets = [[None, None], [None, None]]                        # [which][lo (from below), hi (from above)]
        self.pc = 0
        self.program = []


def build_program(kinds, toggle_sets, consume_lag):
    """kinds: 'G' (ghost planes of Q) / 'P' (predictor planes) per exchange, numbered from 1.  consume_lag = 3 (overlapped
    schedule): predictor planes P_m are read by the kernels of the NEXT step, whose interior launch runs beside exchanges
    m+1 and m+2 and ends just before the slab posts exchange m+3 -- the latest (worst) point is modelled.  consume_lag = 0
    (slabs too thin to overlap): they are read by the kernels that follow the exchange directly."""
    prog, which, pending_p = [], 0, []
    for n, kind in enumerate(kinds, start=1):
        for p in [p for p in pending_p if p[0] + consume_lag <= n]:
            prog.append(("consume_set", p[0], p[1]))
            pending_p.remove(p)
        if kind == "P" and toggle_sets:
            which = 1 - which
        prog.append(("post", n, kind, which))
        prog.append(("wait", n))
        if kind == "G":
            prog.append(("consume_staging", n))
        elif consume_lag == 0:
            prog.append(("consume_set", n, which))
        else:
            pending_p.append((n, which))
    for p in pending_p:
        prog.append(("consume_set", p[0], p[1]))
    return prog


def run(world, kinds, staging_buffers, toggle_sets, seed, consume_lag=3):
    rng = random.Random(seed)
    slabs = [Slab(r, world, staging_buffers) for r in range(world)]
    for s in slabs:
        s.program = build_program(kinds, toggle_sets, consume_lag)
    while True:
        runnable = []
        for s in slabs:
            if s.pc >= len(s.program):
                continue
            op = s.program[s.pc]
            if op[0] == "wait" and not (s.counter[0] >= op[1] and s.counter[1] >= op[1]):
                continue
            runnable.append(s)
        if not runnable:
            assert all(s.pc >= len(s.program) for s in slabs), "deadlock"
            return
        s = rng.choice(runnable)
        op = s.program[s.pc]
        s.pc += 1
        if op[0] == "post":
            _, n, kind, which = op
            above, below = slabs[s.up], slabs[s.down]
            if kind == "G":
                above.staging[n % staging_buffers][0] = (s.rank, n)      # up-going: the neighbour's plane from below
                below.staging[n % staging_buffers][1] = (s.rank, n)
            else:
                above.sets[which][0] = (s.rank, n)
                below.sets[which][1] = (s.rank, n)
            above.counter[0] = n                                          # the sequence number follows the data
            below.counter[1] = n
        elif op[0] == "consume_staging":
            n = op[1]
            got = s.staging[n % staging_buffers]
            assert got == [(s.down, n), (s.up, n)], f"slab {s.rank} unpacks exchange {n}, found {got}"
        elif op[0] == "consume_set":
            _, n, which = op
            assert s.sets[which] == [(s.down, n), (s.up, n)], f"slab {s.rank} reads predictor planes of exchange {n}, found {s.sets[which]}"


STEPPING = ["G", "P"] + ["G", "P"] * 12                       # prime-time exchanges, then twelve overlapped steps
WITH_REFRESH = ["G", "G", "P"] + ["G", "P"] * 6               # set_state_local + refresh_ghosts, prime (path A), steps
THIN = ["G", "P"] * 8                                         # slabs too thin to overlap: ghosts, then the predictor planes at the next step's start


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("kinds,toggle,lag", [(STEPPING, True, 3), (WITH_REFRESH, True, 3), (THIN, False, 0)])
def test_direct_exchange_never_overwrites_an_unread_plane(world, kinds, toggle, lag):
    for seed in range(300):
        run(world, kinds, staging_buffers=2, toggle_sets=toggle, seed=seed, consume_lag=lag)


def test_single_receive_staging_would_be_overwritten():
    """Why recv_stage() exists: two ghost exchanges in a row through ONE staging pair race (the model must notice)."""
    failures = 0
    for seed in range(300):
        try:
            run(2, WITH_REFRESH, staging_buffers=1, toggle_sets=True, seed=seed)
        except AssertionError:
            failures += 1
    assert failures > 0

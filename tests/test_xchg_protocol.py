"""The buffer discipline of the fused exchange kernel (megakv_b200/csrc/gpuhash_xchg.cu), checked on a model (CPU only).

Launch j of a rank scatters exchange j into the owners' inboxes (slot j % B), serves exchange j-1 from its own inbox
(slot (j-1) % B) writing results into the origins' staging areas (slot (j-1) % B), and gathers exchange j-2 from its own
staging area (slot (j-2) % B).  The only synchronisation between ranks: launch j may start once every peer has raised
flag >= j-1, and a rank raises its flag for launch j when its scatter / lookup / gather tiles are done -- BEFORE the launch
ends: its delete / insert tiles keep reading the inbox afterwards.  The model draws random schedules that respect exactly
those rules and looks for a slot that is written while a launch that needs its previous content may still be reading it.
B = 3 (what the kernel uses) must never show one; B = 2 must (otherwise the model could not see the hazard it is for)."""
import random

import pytest


def simulate(G, J, B, seed, flag_at_end=False):
    """returns the list of hazards found: (kind, writer (rank, launch), reader (rank, launch))"""
    rnd = random.Random(seed)
    start = [[None] * (J + 1) for _ in range(G)]       # launch numbers 1..J
    flag = [[None] * (J + 1) for _ in range(G)]
    end = [[None] * (J + 1) for _ in range(G)]
    now = 0.0
    nxt = [1] * G                                      # next launch of each rank
    running = [None] * G                               # (launch, flag time, end time)
    # event-driven: at every step pick a rank that may act
    while any(n <= J for n in nxt) or any(r is not None for r in running):
        now += rnd.random()
        progressed = False
        order = list(range(G)); rnd.shuffle(order)
        for r in order:
            if running[r] is not None:
                j, tf, te = running[r]
                if flag[r][j] is None and now >= tf:
                    flag[r][j] = now; progressed = True
                if now >= te and flag[r][j] is not None:
                    end[r][j] = now; running[r] = None; progressed = True
                continue
            j = nxt[r]
            if j > J:
                continue
            # stream order (own previous launch has ended) + the mem-op wait: every peer's flag >= j-1
            if j > 1 and (end[r][j - 1] is None or any(flag[p][j - 1] is None for p in range(G))):
                continue
            if rnd.random() < 0.5:                       # a rank may also just be late
                continue
            start[r][j] = now
            dur = 0.2 + 3.0 * rnd.random()
            tf = now + dur if flag_at_end else now + dur * rnd.uniform(0.3, 0.95)
            running[r] = (j, tf, now + dur)
            nxt[r] = j + 1; progressed = True
        if not progressed:
            now += 0.5
    hazards = []

    def overlaps_or_precedes(w, rd):
        """a writer interval that starts before the reader interval has ended can clobber what the reader still needs"""
        (ws, we), (rs, re) = w, rd
        return ws < re                                   # the write may land before the read is over

    for r in range(G):
        for j in range(1, J + 1):
            w_int = (start[r][j], end[r][j])
            # inbox slot j % B on every owner d: previous content = exchange j - B, read by owner d's launch j - B + 1 (whole launch: its
            # insert tiles read the inbox until the launch ends)
            for d in range(G):
                k = j - B + 1
                if k >= 1 and overlaps_or_precedes(w_int, (start[d][k], end[d][k])):
                    hazards.append(("inbox", (r, j), (d, k)))
            # staging slot (j-1) % B on every origin s (written by my serve of exchange j-1): previous content = exchange j-1-B, gathered
            # by origin s's launch j-1-B+2 (gather tiles: done by that launch's FLAG time)
            for s_ in range(G):
                k = j - 1 - B + 2
                if k >= 1 and not (flag[s_][k] <= start[r][j]):
                    hazards.append(("stage", (r, j), (s_, k)))
    return hazards


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_three_slots_are_enough_under_any_schedule(G):
    for seed in range(60):
        assert simulate(G, J=12, B=3, seed=seed) == []


def test_two_slots_are_not_enough_for_the_inboxes_and_the_model_can_tell():
    """B = 2: a peer's launch j+1 may start as soon as the flags of launch j are up, and then scatters into inbox slot
    (j+1) % 2 == (j-1) % 2 -- which my launch j may still be serving (its delete / insert tiles run after the flag).
    The staging areas would do with two slots: a gather is over by its launch's flag.  The kernel keeps three of both."""
    kinds, found = set(), 0
    for seed in range(60):
        h = simulate(4, J=12, B=2, seed=seed)
        found += bool(h)
        kinds |= {x[0] for x in h}
    assert found > 0 and kinds == {"inbox"}


def test_two_slots_would_do_if_the_flag_were_raised_at_the_very_end_of_a_launch():
    """the alternative the kernel does not take: flags at launch end need only two slots, at the price of every launch
    waiting for the slowest peer's previous launch to END instead of to finish the part others depend on"""
    for seed in range(60):
        assert simulate(4, J=12, B=2, seed=seed, flag_at_end=True) == []

"""Model check of the cross-CTA protocol of `groupnorm_fused_kernel` (csrc/norm_kernels.cu): work counter, per-image
arrive / ready / depart word, last-arriver reduction — under random interleavings and a bounded number of resident CTAs.

The kernel spins while it waits for sibling CTAs, so what has to hold is (1) progress whenever all slabs of ONE image fit on
the device at once (the launcher's eligibility rule), whatever order the hardware starts and advances CTAs in; (2) the
statistics of an image are reduced exactly once, after every slab has published its partials; (3) every counter is back at
zero when the grid has drained (the next GroupNorm reuses the scratch without clearing it).  The same model shows the rule is
necessary: with fewer resident CTAs than slabs per image the grid deadlocks."""
import random

import pytest


class Grid:
    def __init__(self, images, parts, resident, rng):
        self.nb, self.parts, self.resident, self.rng = images, parts, resident, rng
        self.total = images * parts
        self.work = 0                      # atomic work counter
        self.ticket = [0] * images         # arrive / ready / depart word per image
        self.published = [0] * images
        self.reduced = [0] * images
        self.unstarted = self.total
        self.ctas = []                     # resident CTAs: dict(state, n)
        self.finished = 0

    def runnable(self):
        acts = []
        if self.unstarted and len(self.ctas) < self.resident:
            acts.append(("start", None))
        for c in self.ctas:
            if c["state"] != "wait" or self.ticket[c["n"]] >= self.parts + 1:
                acts.append(("step", c))
        return acts

    def step(self, act, c):
        if act == "start":
            item = self.work
            self.work += 1
            if item == self.total - 1:
                self.work = 0              # the last slab handed out resets the counter for the next launch
            self.unstarted -= 1
            self.ctas.append({"state": "publish", "n": item // self.parts})
            return
        n, s = c["n"], c["state"]
        if s == "publish":
            self.published[n] += 1
            c["state"] = "arrive"
        elif s == "arrive":
            old = self.ticket[n]
            self.ticket[n] += 1
            c["state"] = "reduce" if old == self.parts - 1 else "wait"
        elif s == "reduce":                # the last arriver: every partial must be there, exactly one reduction per image
            assert self.published[n] == self.parts and self.reduced[n] == 0
            self.reduced[n] = 1
            self.ticket[n] += 1            # parts + 1: statistics ready
            c["state"] = "depart"
        elif s == "wait":
            assert self.reduced[n] == 1    # only reachable once the statistics exist
            c["state"] = "depart"
        elif s == "depart":
            old = self.ticket[n]
            self.ticket[n] += 1
            if old == 2 * self.parts:
                self.ticket[n] = 0
            self.ctas.remove(c)
            self.finished += 1

    def run(self):
        while self.finished < self.total:
            acts = self.runnable()
            if not acts:
                return False               # deadlock
            self.step(*self.rng.choice(acts))
        return True


@pytest.mark.parametrize("images,parts,resident", [(1, 1, 1), (3, 1, 2), (5, 4, 4), (7, 4, 9), (64, 55, 444), (4, 164, 444),
                                                   (9, 13, 13), (2, 86, 100)])
def test_protocol_makes_progress_and_leaves_the_counters_at_zero(images, parts, resident):
    for seed in range(20 if images * parts < 1000 else 3):
        g = Grid(images, parts, resident, random.Random(seed))
        if parts == 1:                     # single-slab images take the kernel's shortcut: no cross-CTA traffic at all
            continue
        assert g.run(), (images, parts, resident, seed)
        assert g.work == 0 and g.ticket == [0] * images and g.reduced == [1] * images and not g.ctas


def test_fewer_resident_ctas_than_slabs_per_image_deadlocks():
    """why b200sd_groupnorm only takes the one-pass kernel when `parts <= resident CTAs` (gn_fused_capacity)"""
    stuck = 0
    for seed in range(10):
        g = Grid(3, 8, 5, random.Random(seed))
        stuck += not g.run()
    assert stuck == 10

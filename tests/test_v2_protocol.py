"""Model check of the mbarrier protocol of the role-specialised appearance kernel (csrc/appearance_mma2.cuh).

The kernel's four roles (8 decoder warps, 16 gather warps, two issuer and two loader warps) only meet on mbarriers.  This
test restates the protocol -- every wait / arrive / commit in the order the kernel executes them, with the ring depths,
arrival counts and TMEM map the library exports (t2n_debug_v2_plan) -- as a set of cooperating actors, runs them under
many random schedules together with a model of the in-order tensor pipe and the TMA engine, and checks

* no deadlock, every tile's colours written exactly once;
* mbarrier phase discipline: no barrier runs two phases ahead of a waiter (parity waits would alias);
* no data hazard on the rings and accumulators: a TMEM / shared-memory A stage or a weight stage is never rewritten
  before the MMAs that read it have completed, D0 / D1 / D2 are never overwritten before their readers are done and
  never read before they are complete.

It is a model of the design, not of the compiled code (the GPU parity tests check that); it guards the reasoning behind
the ring depths and the order of the waits when the kernel is changed.
"""
import ctypes as C
import random

import pytest

from text2nerf_b200 import _native as nat


def plan(n_app_total, Kp, view_cols):
    out = (C.c_int * 32)()
    n = nat.load().t2n_debug_v2_plan(n_app_total, Kp, view_cols, out, 32)
    assert n >= 20, n
    keys = ["smem_bytes", "threads", "regs_per_thread", "tmem_cols", "col_d1", "col_d2", "col_d0", "n_d0", "col_a", "a_stages",
            "a_stage_cols", "nb", "g_stages", "nk0", "nk1", "nk2", "p_warps", "g_warps", "p_full_arrivals", "g_full_arrivals",
            "d0_free_arrivals"]
    return dict(zip(keys, list(out)[:len(keys)]))


@pytest.mark.parametrize("n_app_total,Kp,view_cols", [(144, 416, 0), (144, 160, 3), (128, 416, 3), (160, 416, 3), (48, 288, 0)])
def test_budgets(n_app_total, Kp, view_cols):
    p = plan(n_app_total, Kp, view_cols)
    assert p["smem_bytes"] <= 232448                                   # opt-in dynamic shared memory of a B200 CTA
    assert p["threads"] <= 1024 and p["threads"] * p["regs_per_thread"] <= 65536
    assert p["threads"] == 32 * (p["p_warps"] + p["g_warps"] + 4)
    regions = [(p["col_d1"], 128), (p["col_d2"], 128)] + [(p["col_d0"] + 32 * b, 32) for b in range(p["n_d0"])] + \
              [(p["col_a"] + p["a_stage_cols"] * s, p["a_stage_cols"]) for s in range(p["a_stages"])]
    cover = set()
    for lo, n in regions:
        cols = set(range(lo, lo + n))
        assert not (cover & cols), "TMEM regions overlap"
        cover |= cols
    assert max(cover) < p["tmem_cols"] <= 512
    assert p["nk0"] == (n_app_total + view_cols + 31) // 32 and p["nk1"] == Kp // 32 and p["nk2"] == 4
    assert p["a_stages"] < 4 and p["nb"] == 4 and p["g_stages"] == 2    # the barrier slot arithmetic below assumes these


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending = self.count
            self.completed += 1

    def passed(self, use):
        """The kernel waits for use number `use` (0-based) of this barrier with parity use & 1."""
        assert self.completed <= use + 1, f"{self.name}: phase {self.completed} completed while a waiter still needs phase {use}"
        return self.completed > use


class Sim:
    def __init__(self, p, n_tiles, rng):
        self.p, self.n, self.rng = p, n_tiles, rng
        B = lambda name, cnt=1: Bar(name, cnt)
        self.pb_full = [B(f"pb_full{i}") for i in range(4)]
        self.p_done = [B(f"p_done{i}") for i in range(4)]
        self.p_full = [B(f"p_full{i}", p["p_full_arrivals"]) for i in range(4)]
        self.acc1, self.acc2 = B("acc1"), B("acc2")
        self.g_full = [B(f"g_full{i}", p["g_full_arrivals"]) for i in range(2)]
        self.g_done = [B(f"g_done{i}") for i in range(2)]
        self.bb_full = [B(f"bb_full{i}") for i in range(2)]
        self.d0_full = [B(f"d0_full{i}") for i in range(2)]
        self.d0_free = [B(f"d0_free{i}", p["d0_free_arrivals"]) for i in range(2)]
        self.pipe = []              # in-order tensor pipe: (kind, payload, commits)
        self.tma = []               # outstanding bulk copies: (ring, stage, chunk, barrier)
        # resource state for the hazard checks
        self.a_tmem = [None] * p["a_stages"]        # decoder A stage: chunk whose MMAs have not completed
        self.a_written = {}                         # decoder chunk -> set of warps that stored it
        self.pb = [None] * 4                        # decoder weight stage: chunk resident / in use
        self.ga = [None] * 2                        # basis A stage
        self.ga_written = {}
        self.bb = [None] * 2
        self.d0_owner = [None, None]                # tile whose D0 is (being) accumulated / valid
        self.d0_complete = [False, False]
        self.d0_readers = [set(), set()]
        self.d1_tile, self.d1_complete, self.d1_readers = None, False, set()
        self.d2_tile, self.d2_complete, self.d2_readers = None, False, set()
        self.rgb_written = {}
        self.psync = 0                              # decoder-group barrier (bar.sync 1): arrivals of the current phase
        self.psync_phase = 0

    # ---- actors (generators yield a predicate to wait for, or None to give the scheduler a turn) -------------------
    def decoder(self, w):
        p, n = self.p, self.n
        nk1, nk2, ns = p["nk1"], p["nk2"], p["a_stages"]
        it, pend = 0, None

        def flush():
            nonlocal pend
            if pend is not None:
                self.p_full[pend & 3].arrive()
                pend = None

        def chunk(reads=None):
            nonlocal it, pend
            c = it
            yield None
            flush()
            if c - ns >= 0:
                yield lambda: self.p_done[(c - ns) & 3].passed((c - ns) >> 2)
            # tcgen05.st into stage c % ns
            s = c % ns
            assert self.a_tmem[s] is None or self.a_tmem[s] == c, f"decoder stage {s} rewritten while chunk {self.a_tmem[s]} is in flight"
            self.a_tmem[s] = c
            self.a_written.setdefault(c, set()).add(w)
            pend = c
            it += 1

        for i in range(n + 1):
            if i >= 1:                                              # S2(i-1)
                flush()
                yield lambda: self.acc1.passed(i - 1)
                for c in range(nk2):
                    assert self.d1_tile == i - 1 and self.d1_complete, "layer 2 reads an incomplete D1"
                    self.d1_readers.add((w, c))
                    yield from chunk()
            if i < n:                                               # S1(i)
                flush()
                b = i & 1
                yield lambda: self.d0_full[b].passed(i >> 1)
                assert self.d0_owner[b] == i and self.d0_complete[b], "decoder reads an incomplete D0"
                yield from chunk()                                  # identity chunk
                self.d0_readers[b].add(w)
                self.d0_free[b].arrive()
                for _ in range(nk1 - 1):
                    yield from chunk()
            if i >= 1:                                              # S3(i-1)
                flush()
                yield lambda: self.acc2.passed(i - 1)
                assert self.d2_tile == i - 1 and self.d2_complete, "layer 3 reads an incomplete D2"
                self.d2_readers.add(w)
                my_phase = self.psync_phase                          # bar.sync over the 8 decoder warps
                self.psync += 1
                if self.psync == p["p_warps"]:
                    self.psync, self.psync_phase = 0, self.psync_phase + 1
                yield lambda: self.psync_phase > my_phase
                self.rgb_written[(i - 1, w)] = self.rgb_written.get((i - 1, w), 0) + 1
        flush()

    def gather(self, w):
        p, n = self.p, self.n
        gi = 0
        for i in range(n):
            for c in range(p["nk0"]):
                yield None
                if gi >= 2:
                    yield lambda g=gi: self.g_done[g & 1].passed((g - 2) >> 1)
                s = gi & 1
                assert self.ga[s] is None or self.ga[s] == gi, f"basis A stage {s} rewritten while chunk {self.ga[s]} is in flight"
                self.ga[s] = gi
                self.ga_written.setdefault(gi, set()).add(w)
                yield None
                self.g_full[s].arrive()
                gi += 1

    def g_issuer(self):
        p, n = self.p, self.n
        gi = 0
        for i in range(n):
            for c in range(p["nk0"]):
                s = gi & 1
                if c == 0 and i >= 2:
                    yield lambda: self.d0_free[i & 1].passed((i >> 1) - 1)
                yield lambda g=gi: self.bb_full[g & 1].passed(g >> 1)
                yield lambda g=gi: self.g_full[g & 1].passed(g >> 1)
                assert len(self.ga_written.get(gi, ())) == p["g_warps"] and self.bb[s] == ("landed", gi)
                if c == 0:          # the first MMA of a tile overwrites D0[b]: it may execute as soon as it is issued
                    b = i & 1
                    assert self.d0_owner[b] is None or len(self.d0_readers[b]) == p["p_warps"], "D0 overwritten before the decoder read it"
                    self.d0_owner[b], self.d0_complete[b], self.d0_readers[b] = i, False, set()
                commits = [self.g_done[s]] + ([self.d0_full[i & 1]] if c == p["nk0"] - 1 else [])
                self.pipe.append(("basis", (i, c, gi), commits))
                gi += 1

    def g_loader(self):
        p, n = self.p, self.n
        gi = 0
        for i in range(n):
            for c in range(p["nk0"]):
                if gi >= 2:
                    yield lambda g=gi: self.g_done[g & 1].passed((g - 2) >> 1)
                s = gi & 1
                assert self.bb[s] is None, f"basis weight stage {s} overwritten while in use"
                self.bb[s] = ("flying", gi)
                self.tma.append(("bb", s, gi, self.bb_full[s]))
                gi += 1
                yield None

    def p_schedule(self):
        p, n = self.p, self.n
        for i in range(n + 1):
            if i >= 1:
                for c in range(p["nk2"]):
                    yield ("S2", i - 1, c)
            if i < n:
                for c in range(p["nk1"]):
                    yield ("S1", i, c)

    def p_issuer(self):
        p = self.p
        ns = p["a_stages"]
        for it, (kind, tile, c) in enumerate(self.p_schedule()):
            bs = it & 3
            yield lambda k=it: self.pb_full[k & 3].passed(k >> 2)
            yield lambda k=it: self.p_full[k & 3].passed(k >> 2)
            assert len(self.a_written.get(it, ())) == p["p_warps"] and self.a_tmem[it % ns] == it and self.pb[bs] == ("landed", it)
            if c == 0 and kind == "S1":
                if self.d1_tile is not None:
                    assert self.d1_readers == {(w, k) for w in range(p["p_warps"]) for k in range(p["nk2"])}, \
                        "D1 overwritten before layer 2 of the previous tile read it"
                self.d1_tile, self.d1_complete, self.d1_readers = tile, False, set()
            if c == 0 and kind == "S2":
                if self.d2_tile is not None:
                    assert len(self.d2_readers) == p["p_warps"], "D2 overwritten before layer 3 of the previous tile read it"
                self.d2_tile, self.d2_complete, self.d2_readers = tile, False, set()
            last = c == (p["nk2"] if kind == "S2" else p["nk1"]) - 1
            commits = [self.p_done[bs]] + ([self.acc2 if kind == "S2" else self.acc1] if last else [])
            self.pipe.append((kind, (tile, c, it), commits))

    def p_loader(self):
        for ld, _ in enumerate(self.p_schedule()):
            bs = ld & 3
            if ld >= 4:
                yield lambda k=ld: self.p_done[k & 3].passed((k >> 2) - 1)
            assert self.pb[bs] is None, f"decoder weight stage {bs} overwritten while in use"
            self.pb[bs] = ("flying", ld)
            self.tma.append(("pb", bs, ld, self.pb_full[bs]))
            yield None

    # ---- hardware models ------------------------------------------------------------------------------------------
    def tma_step(self):
        k = self.rng.randrange(len(self.tma))
        ring, s, chunk, bar = self.tma.pop(k)
        tgt = self.bb if ring == "bb" else self.pb
        assert tgt[s] == ("flying", chunk)
        tgt[s] = ("landed", chunk)
        bar.arrive()

    def pipe_step(self):
        p = self.p
        kind, (tile, c, idx), commits = self.pipe.pop(0)            # the tensor pipe completes MMAs in issue order
        if kind == "basis":
            b = tile & 1
            assert self.d0_owner[b] == tile
            if c == p["nk0"] - 1:
                self.d0_complete[b] = True
            self.ga[idx & 1] = None
            self.bb[idx & 1] = None
        else:
            if kind == "S1" and c == p["nk1"] - 1:
                self.d1_complete = True
            if kind == "S2" and c == p["nk2"] - 1:
                self.d2_complete = True
            self.a_tmem[idx % p["a_stages"]] = None
            self.pb[idx & 3] = None
        for b in commits:
            b.arrive()

    def run(self):
        p = self.p
        actors = [self.decoder(w) for w in range(p["p_warps"])] + [self.gather(w) for w in range(p["g_warps"])] + \
                 [self.g_issuer(), self.g_loader(), self.p_issuer(), self.p_loader()]
        waiting = [None] * len(actors)
        alive = set(range(len(actors)))
        steps = 0
        while alive or self.pipe or self.tma:
            steps += 1
            assert steps < 2_000_000
            choices = [("actor", k) for k in alive if waiting[k] is None or waiting[k]()]
            if self.pipe:
                choices.append(("pipe", 0))
            if self.tma:
                choices.append(("tma", 0))
            assert choices, "deadlock: " + ", ".join(str(k) for k in sorted(alive))
            kind, k = self.rng.choice(choices)
            if kind == "pipe":
                self.pipe_step()
            elif kind == "tma":
                self.tma_step()
            else:
                waiting[k] = None
                try:
                    waiting[k] = next(actors[k])
                except StopIteration:
                    alive.discard(k)
        for t in range(self.n):
            for w in range(p["p_warps"]):
                assert self.rgb_written.get((t, w)) == 1, f"tile {t}: colours of decoder warp {w} written {self.rgb_written.get((t, w))} times"


@pytest.mark.parametrize("n_app_total,Kp,view_cols", [(144, 416, 0), (48, 96, 3), (160, 224, 3)])
@pytest.mark.parametrize("n_tiles", [0, 1, 2, 3, 6])
def test_protocol_random_schedules(n_app_total, Kp, view_cols, n_tiles):
    p = plan(n_app_total, Kp, view_cols)
    for seed in range(6):
        Sim(p, n_tiles, random.Random(1000 * n_tiles + seed)).run()


def test_model_detects_a_missing_wait():
    """The checker is not vacuous: with the D0 double buffer's `d0_free` wait removed, some schedule overwrites D0 early."""
    p = plan(144, 416, 0)

    class Broken(Sim):
        def g_issuer(self):
            gi = 0
            for i in range(self.n):
                for c in range(self.p["nk0"]):
                    s = gi & 1
                    yield lambda g=gi: self.bb_full[g & 1].passed(g >> 1)
                    yield lambda g=gi: self.g_full[g & 1].passed(g >> 1)
                    if c == 0:
                        b = i & 1
                        assert self.d0_owner[b] is None or len(self.d0_readers[b]) == self.p["p_warps"], "D0 overwritten early"
                        self.d0_owner[b], self.d0_complete[b], self.d0_readers[b] = i, False, set()
                    commits = [self.g_done[s]] + ([self.d0_full[i & 1]] if c == self.p["nk0"] - 1 else [])
                    self.pipe.append(("basis", (i, c, gi), commits))
                    gi += 1

    failures = 0
    for seed in range(20):
        try:
            Broken(p, 5, random.Random(seed)).run()
        except AssertionError:
            failures += 1
    assert failures > 0

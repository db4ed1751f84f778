"""Discrete simulation of conv_ss.cu's warp-role protocol (mbarrier phases / parities, ring slots) — runs anywhere:

    python tools/probe/simulate_protocol.py

A wrong parity or arrival count in a warp-specialised kernel does not give wrong numbers, it HANGS the GPU (a strike under gpurun),
so the protocol is checked here before the kernel ever meets hardware.  Each role of the kernel is a Python generator that
mirrors the CUDA code's loops, index arithmetic and waits one to one (TMA producer, the two lo-pass groups, the UMMA issuer, the two
epilogue groups); mbarriers are modelled with their real semantics (pending-arrival count, expected-tx bytes, one phase bit,
try_wait(parity) succeeds iff the phase with that parity has completed); asynchronous completions (TMA bytes landing, tcgen05.commit
arrivals) fire after random delays, and a random scheduler interleaves the roles.  Checked invariants:
  * no deadlock: every role terminates for many random schedules and tile shapes (incl. odd heights, ring wrap, many tiles);
  * a staged row is never overwritten by TMA, nor its lo buffer by the lo pass, while UMMAs that read it are still in flight;
  * the UMMA issuer only reads rows whose lo buffer has been written for THIS use of the stage;
  * every accumulator slot is zero when an output row opens it, receives exactly its three taps before the
    epilogue reads it, and is read exactly once per output row;
  * every output row of every tile is produced exactly once.
The pair kernel (conv_pair.cu) is the same protocol without the lo pass (the issuer waits on s_full directly): `--pair`.
Found this way: with an ODD ring depth the two lo-pass groups take turns on a stage, a group then waits for phase n of s_full
without having seen phase n-1 complete, and the parity wait wakes up two phases early — conv_ss.cu now static_asserts NS % 2 == 0.
"""
import random
import sys

PND = 8


class MBar:
    def __init__(self, count):
        self.init, self.pending, self.tx, self.phase = count, count, 0, 0   # phase = number of completed phases

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.init

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._maybe_complete()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        assert self.tx >= 0
        self._maybe_complete()

    def done(self, parity):          # try_wait.parity: true iff the phase with this parity has completed
        # the barrier is "in" phase self.phase (0-based, incomplete); waiting on parity p succeeds when the current incomplete
        # phase has parity != p, i.e. the last completed phase has parity p
        return (self.phase & 1) != (parity & 1)


class Sim:
    def __init__(self, H, W, TR, NS, gridx, bid, pair_kernel, rng):
        self.H, self.W, self.TR, self.NS, self.rng, self.pair = H, W, TR, NS, rng, pair_kernel
        assert pair_kernel or NS % 2 == 0, "conv_ss: the ring depth must be even (see SsGeom's static_assert)"
        self.tiles_x, self.tiles_y = (W + 127) // 128, (H + TR - 1) // TR
        self.tiles = list(range(bid, self.tiles_x * self.tiles_y, gridx))
        self.s_full = [MBar(1) for _ in range(NS)]
        self.a_ready = [MBar(4) for _ in range(NS)]
        self.s_empty = [MBar(1) for _ in range(NS)]
        self.d_full = [MBar(1) for _ in range(PND)]
        self.d_empty = [MBar(4) for _ in range(PND)]
        self.pending = []                    # async completions: (ready_time, fn)
        self.time = 0
        # ---- shadow state for the invariants ----
        self.stage_use = [None] * NS         # (input-row id) currently staged
        self.stage_lo = [None] * NS          # input-row id whose lo has been written
        self.stage_readers = [0] * NS        # UMMA batches in flight that read the stage
        self.slot_taps = [[0] * 4 for _ in range(PND)]   # taps accumulated since the last zeroing, per TMEM lane quarter (one warp each)
        self.slot_inflight = [0] * PND       # UMMAs in flight into the slot
        self.slot_owner = [None] * PND
        self.produced = {}

    def tile_rows(self, tile):
        y0 = (tile // self.tiles_x) * self.TR
        n = min(self.H - y0, self.TR)
        return (n + 1) & ~1

    def later(self, fn, lo=1, hi=40):
        self.pending.append((self.time + self.rng.randint(lo, hi), fn))

    # ---------------- roles (generators yield a predicate to wait for, or None to just be rescheduled) ----------------
    def tma(self):
        i = 0
        for tile in self.tiles:
            for r in range(-1, self.tile_rows(tile) + 1):
                s, n = i % self.NS, i // self.NS
                if n >= 1:
                    yield lambda s=s, n=n: self.s_empty[s].done((n - 1) & 1)
                assert self.stage_readers[s] == 0, "TMA overwrites a row that UMMAs still read"
                self.stage_use[s] = None
                self.s_full[s].arrive_expect_tx(100)

                def land(s=s, rid=(tile, r)):
                    self.stage_use[s] = rid
                    self.s_full[s].complete_tx(100)
                self.later(land)
                i += 1
                yield None

    def lo_pass(self, lgroup, w):
        i = 0
        for tile in self.tiles:
            for r in range(-1, self.tile_rows(tile) + 1):
                if (i & 1) == lgroup:
                    s = i % self.NS
                    yield lambda s=s, i=i: self.s_full[s].done((i // self.NS) & 1)
                    assert self.stage_use[s] == (tile, r), "lo pass reads a row that is not the one it expects"
                    assert self.stage_readers[s] == 0, "lo buffer rewritten while UMMAs still read it"
                    if w == 0:
                        self.stage_lo[s] = (tile, r)
                    yield None
                    self.a_ready[s].arrive()
                i += 1

    def issuer(self):
        i, g0 = 0, 0
        for tile in self.tiles:
            nrows = self.tile_rows(tile)
            for r in range(-1, nrows + 1):
                s = i % self.NS
                ready = self.s_full if self.pair else self.a_ready
                conds = [lambda s=s, i=i: ready[s].done((i // self.NS) & 1)]
                g = g0 + r + 1
                if r + 1 <= nrows - 1 and g >= PND:
                    conds.append(lambda g=g: self.d_empty[g % PND].done((g // PND - 1) & 1))
                yield lambda conds=conds: all(c() for c in conds)
                assert self.stage_use[s] == (tile, r), "issuer reads a stage that holds another row"
                if not self.pair:
                    assert self.stage_lo[s] == (tile, r), "issuer reads a stale lo buffer"
                lo, hi = max(r - 1, 0), min(r + 1, nrows - 1)
                touched = []
                o = lo
                while o <= hi:
                    slot = (g0 + o) % PND
                    n = min(hi - o + 1, PND - slot)
                    for t in range(n):
                        sl, orow = slot + t, (tile, o + t)
                        if self.slot_owner[sl] != orow:                      # this UMMA opens the slot for a new output row
                            assert self.slot_taps[sl] == [0] * 4 and self.slot_inflight[sl] == 0, f"slot {sl} opened before it was drained / zeroed"
                            self.slot_owner[sl] = orow
                        self.slot_inflight[sl] += 1
                        touched.append(sl)
                    o += n
                self.stage_readers[s] += 1
                fin = (g0 + r - 1) % PND if r >= 1 else None

                def complete(s=s, touched=tuple(touched), fin=fin):          # tcgen05.commit: everything issued so far has finished
                    for sl in touched:
                        self.slot_inflight[sl] -= 1
                        for q in range(4):
                            self.slot_taps[sl][q] += 1
                    self.stage_readers[s] -= 1
                    self.s_empty[s].arrive()
                    if fin is not None:
                        self.d_full[fin].arrive()
                self.commits.append(complete)
                i += 1
                yield None
            g0 += nrows

    def epilogue(self, group, w):
        g0 = 0
        for tile in self.tiles:
            nrows = self.tile_rows(tile)
            y0 = (tile // self.tiles_x) * self.TR
            for m in range(nrows // 2):
                g = g0 + 2 * m
                if ((g >> 1) & 1) != group:
                    continue
                slot = g % PND
                yield lambda slot=slot, g=g: self.d_full[slot + 1].done((g // PND) & 1)
                for h in range(2):
                    sl, orow = slot + h, 2 * m + h
                    assert self.slot_owner[sl] == (tile, orow), "epilogue reads a slot that belongs to another row"
                    assert self.slot_inflight[sl] == 0, "epilogue reads a slot with UMMAs in flight"
                    want = 3                                                 # rows r-1, r, r+1 (the tile's halo rows are staged like any other)
                    assert self.slot_taps[sl][w] == want, f"row {orow}: {self.slot_taps[sl][w]} taps accumulated, expected {want}"
                yield None
                for h in range(2):
                    self.slot_taps[slot + h][w] = 0                          # this warp's lane quarter is re-zeroed
                if w == 0:
                    for h in range(2):
                        key = (tile, 2 * m + h)
                        assert key not in self.produced, "an output row is produced twice"
                        self.produced[key] = y0 + 2 * m + h
                self.d_empty[slot].arrive()
                self.d_empty[slot + 1].arrive()
            g0 += nrows

    def run(self):
        self.commits = []                    # commit callbacks complete IN ORDER (tcgen05.commit tracks all prior UMMAs)
        self.ep_sync = [[], []]
        roles = {"tma": self.tma(), "issuer": self.issuer()}
        if not self.pair:
            for lg in range(2):
                for w in range(4):
                    roles[f"lo{lg}.{w}"] = self.lo_pass(lg, w)
        for grp in range(2):
            for w in range(4):
                roles[f"epi{grp}.{w}"] = self.epilogue(grp, w)
        waiting = {k: None for k in roles}
        steps = 0
        while roles:
            steps += 1
            assert steps < 5_000_000, "runaway simulation"
            self.time += 1
            # async completions: TMA landings in any order, commits strictly in issue order
            due = [p for p in self.pending if p[0] <= self.time]
            for p in due:
                self.pending.remove(p)
                p[1]()
            if self.commits and self.rng.random() < 0.3:
                self.commits.pop(0)()
            runnable = [k for k in roles if waiting[k] is None or waiting[k]()]
            if not runnable:
                if self.pending or self.commits:
                    continue
                raise AssertionError(f"DEADLOCK at t={self.time}: waiting roles {sorted(roles)}")
            k = self.rng.choice(runnable)
            try:
                waiting[k] = next(roles[k])
            except StopIteration:
                del roles[k]
                del waiting[k]
        while self.commits:
            self.commits.pop(0)()
        want = {(t, o) for t in self.tiles for o in range(self.tile_rows(t))}
        assert set(self.produced) == want, "not every output row was produced"


if __name__ == "__main__":
    pair = "--pair" in sys.argv
    rng = random.Random(1)
    cases = 0
    for H, W, TR, NS, gridx in ((64, 128, 64, 12, 1), (37, 130, 32, 12, 1), (70, 200, 32, 4, 2), (512, 300, 8, 5, 3), (200, 128, 4, 4, 1),
                                (1024, 256, 64, 8, 4), (6, 128, 32, 16, 1), (258, 640, 32, 4, 148)):
        for bid in range(min(gridx, 3)):
            for rep in range(6):
                Sim(H, W, TR, NS if pair else NS + (NS & 1), gridx, bid, pair, random.Random(rng.random())).run()
                cases += 1
    print(f"{'conv_pair' if pair else 'conv_ss'} protocol: {cases} randomised schedules, no deadlock, all invariants hold")

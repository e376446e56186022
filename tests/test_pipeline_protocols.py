"""Executable model of the mbarrier protocols of the tcgen05 kernels (csrc/gemm_f32_tc.cu sgemm_tc_kernel, csrc/gemm_f64_i8.cu dgemm_i8_kernel).
Written in round 1 for kernels that had not yet run; both run on hardware since round 2 (bit-exact parity tests), so this is now a cheap pre-flight
check for changes to their pipelines - it is not evidence that a kernel works, the GPU tests are.  (The model also still knows the roles of the
A-through-TMEM and CTA-pair variants that were measured slower and removed; only the one-CTA configurations are exercised.)

Every role of a kernel (TMA producer, split warps, relay lane, MMA issuer, accumulate warps) is a generator transcribed from the kernel's
loops - same barriers, same arrival counts, same parity bookkeeping, same order of waits / arrives / commits; a tiny scheduler interleaves
them with random delays (TMA latency, MMA completion, warp skew) over many seeds and checks
  * liveness: every role finishes (a wrong count or parity shows up as "all roles blocked"),
  * stage hazards: a ring stage / TMEM A slot is refilled only after the MMAs that read it have completed, and read only when filled,
  * accumulator hazards: a TMEM window is overwritten only after every accumulate warp has drained it, and drained only when complete.
The model knows nothing about data or layouts - those are covered by tools/tc_probe2.cu on the box - only about who waits for whom."""
import random

import pytest


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for (phase overrun)"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):  # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return self.phase != parity


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.roles, self.timers, self.now = [], [], 0

    def add(self, name, gen):
        self.roles.append([name, gen, None])  # [name, generator, blocked-on]

    def later(self, lo, hi, fn):
        self.timers.append((self.now + self.rng.randint(lo, hi), self.rng.random(), fn))

    def run(self, limit=2_000_000):
        live = list(self.roles)
        while live:
            self.now += 1
            assert self.now < limit, "model ran away"
            due = sorted(t for t in self.timers if t[0] <= self.now)
            self.timers = [t for t in self.timers if t[0] > self.now]
            for _, _, fn in due:
                fn()
            progressed = bool(due)
            self.rng.shuffle(live)
            for role in list(live):
                name, gen, blocked = role
                if blocked is not None:
                    kind, a, b = blocked
                    if kind == "wait" and not a.passed(b):
                        continue
                    if kind == "sleep" and self.now < a:
                        continue
                    role[2] = None
                try:
                    req = next(gen)
                    progressed = True
                    if req[0] == "wait":
                        role[2] = ("wait", req[1], req[2])
                    elif req[0] == "sleep":
                        role[2] = ("sleep", self.now + req[1], None)
                except StopIteration:
                    live.remove(role)
                    progressed = True
            sleeping = any(r[2] is not None and r[2][0] == "sleep" for r in live)
            if not progressed and not self.timers and not sleeping:
                stuck = [(r[0], r[2][0]) for r in live]
                raise AssertionError(f"deadlock at t={self.now}: {stuck}")


class Pipe:
    """State of one CTA's ring and TMEM windows, for the hazard checks."""

    def __init__(self, stages, accs):
        self.stage = ["free"] * stages     # free -> loading -> landed -> (split) ready -> in_mma -> free
        self.reads_outstanding = [0] * stages
        self.acc = ["free"] * accs         # free -> accumulating -> full -> free
        self.acc_drains = [0] * accs


def mma_async(sim, pipes, stage, bars_empty, on_done=None, lo=5, hi=60):
    """tcgen05.mma batch on `stage` of every CTA in `pipes`, then tcgen05.commit to bars_empty: completion is asynchronous."""
    for p in pipes:
        assert p.stage[stage] == "ready", f"MMA reads stage {stage} in state {p.stage[stage]}"
        p.stage[stage] = "in_mma"

    def done():
        for p in pipes:
            p.stage[stage] = "free"
        for b in bars_empty:
            b.arrive()
        if on_done:
            on_done()
    sim.later(lo, hi, done)


# ------------------------------------------------------------------------------------------------------------------------------------------
# sgemm_tc_ts_kernel<PAIR>: STAGES = 4, ACC_BUFS = 2, split warps = 4 (A -> TMEM) + 4 (B in smem); pair: both CTAs' split warps and accumulate
# warps arrive on rank 0's ready / acc_empty barriers, commits are multicast
def build_ts(sim, pair, tiles, kblocks, window, stages=4, accs=2):
    STAGES, ACCS, SPLIT_WARPS = stages, accs, 8
    ncta = 2 if pair else 1
    full = [[Barrier(1) for _ in range(STAGES)] for _ in range(ncta)]
    ready = [Barrier(2 * SPLIT_WARPS if pair else SPLIT_WARPS) for _ in range(STAGES)]        # rank 0's
    empty = [[Barrier(1) for _ in range(STAGES)] for _ in range(ncta)]
    acc_full = [[Barrier(1) for _ in range(ACCS)] for _ in range(ncta)]
    acc_empty = [Barrier(8 if pair else 4) for _ in range(ACCS)]                               # rank 0's
    pipes = [Pipe(STAGES, ACCS) for _ in range(ncta)]
    split_left = [[0] * STAGES for _ in range(ncta)]

    def producer(r):
        stage, phase = 0, 0
        for _ in range(tiles):
            for _ in range(kblocks):
                yield ("wait", empty[r][stage], phase ^ 1)
                assert pipes[r].stage[stage] == "free", f"TMA overwrites stage {stage} in state {pipes[r].stage[stage]}"
                pipes[r].stage[stage] = "loading"

                def landed(r=r, s=stage):
                    pipes[r].stage[s] = "landed"
                    split_left[r][s] = SPLIT_WARPS
                    full[r][s].arrive()
                sim.later(20, 200, landed)
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1

    def split_warp(r):
        stage, phase = 0, 0
        for _ in range(tiles):
            for _ in range(kblocks):
                yield ("wait", full[r][stage], phase)
                assert pipes[r].stage[stage] == "landed", f"split reads stage {stage} in state {pipes[r].stage[stage]}"
                yield ("sleep", sim.rng.randint(1, 30))
                split_left[r][stage] -= 1
                if split_left[r][stage] == 0:
                    pipes[r].stage[stage] = "ready"
                ready[stage].arrive()          # pair: remote arrive on rank 0's barrier
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1

    def mma():
        stage, acc, phase, acc_phase = 0, 0, 0, 0
        for _ in range(tiles):
            wk = 0
            for kb in range(kblocks):
                if wk == 0:
                    yield ("wait", acc_empty[acc], acc_phase ^ 1)
                    for p in pipes:
                        assert p.acc[acc] == "free", f"window {acc} reopened in state {p.acc[acc]}"
                        p.acc[acc] = "accumulating"
                yield ("wait", ready[stage], phase)
                last = (wk + 1 == window) or (kb == kblocks - 1)

                def close(a=acc):
                    for r in range(ncta):
                        pipes[r].acc[a] = "full"
                        pipes[r].acc_drains[a] = 4
                        acc_full[r][a].arrive()
                mma_async(sim, pipes, stage, [empty[r][stage] for r in range(ncta)], on_done=close if last else None)
                stage += 1
                if stage == STAGES:
                    stage, phase = 0, phase ^ 1
                wk += 1
                if last:
                    acc += 1
                    if acc == ACCS:
                        acc, acc_phase = 0, acc_phase ^ 1
                    wk = 0

    def accumulate_warp(r):
        acc, acc_phase = 0, 0
        windows = (kblocks + window - 1) // window
        for _ in range(tiles):
            for _ in range(windows):
                yield ("wait", acc_full[r][acc], acc_phase)
                assert pipes[r].acc[acc] == "full", f"window {acc} drained in state {pipes[r].acc[acc]}"
                yield ("sleep", sim.rng.randint(1, 40))
                pipes[r].acc_drains[acc] -= 1
                if pipes[r].acc_drains[acc] == 0:
                    pipes[r].acc[acc] = "free"
                acc_empty[acc].arrive()        # pair: remote arrive on rank 0's barrier
                acc += 1
                if acc == ACCS:
                    acc, acc_phase = 0, acc_phase ^ 1
            yield ("sleep", sim.rng.randint(1, 300))  # the tile's global-memory epilogue

    for r in range(ncta):
        sim.add(f"tma{r}", producer(r))
        for w in range(SPLIT_WARPS):
            sim.add(f"split{r}.{w}", split_warp(r))
        for w in range(4):
            sim.add(f"acc{r}.{w}", accumulate_warp(r))
    sim.add("mma", mma())


# ------------------------------------------------------------------------------------------------------------------------------------------
# dgemm_i8_kernel (one CTA): STAGES = 6, ACC_BUFS = 4, 8 accumulate warps; per tile the groups g = S-1 .. 0 with (g + 1) * kblocks k-blocks each
# pair::igemm_group_kernel (two CTAs, one group per launch): relay lane per CTA -> rank 0's ready (count 2); ACC_BUFS = 2; acc_empty count 16
def build_i8(sim, pair, tiles, kblocks, slices, max_window):
    STAGES, ACCS = 6, (2 if pair else 4)
    ncta = 2 if pair else 1
    groups = [slices - 1] if pair else list(range(slices - 1, -1, -1))   # the pair kernel is launched once per group
    full = [[Barrier(1) for _ in range(STAGES)] for _ in range(ncta)]
    ready = [Barrier(2) for _ in range(STAGES)]
    empty = [[Barrier(1) for _ in range(STAGES)] for _ in range(ncta)]
    acc_full = [[Barrier(1) for _ in range(ACCS)] for _ in range(ncta)]
    acc_empty = [Barrier(16 if pair else 8) for _ in range(ACCS)]
    pipes = [Pipe(STAGES, ACCS) for _ in range(ncta)]

    def producer(r):
        stage, phase = 0, 0
        for _ in range(tiles):
            for g in groups:
                for _s in range(g + 1):
                    for _ in range(kblocks):
                        yield ("wait", empty[r][stage], phase ^ 1)
                        assert pipes[r].stage[stage] == "free", f"TMA overwrites stage {stage} in state {pipes[r].stage[stage]}"
                        pipes[r].stage[stage] = "loading"

                        def landed(r=r, s=stage):
                            pipes[r].stage[s] = "ready" if not pair else "landed"
                            full[r][s].arrive()
                        sim.later(20, 200, landed)
                        stage += 1
                        if stage == STAGES:
                            stage, phase = 0, phase ^ 1

    def relay(r):
        stage, phase = 0, 0
        for _ in range(tiles):
            for g in groups:
                for _ in range((g + 1) * kblocks):
                    yield ("wait", full[r][stage], phase)
                    pipes[r].stage[stage] = "ready"
                    ready[stage].arrive()
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1

    def mma():
        stage, acc, phase, acc_phase = 0, 0, 0, 0
        for _ in range(tiles):
            for g in groups:
                group_kblocks = (g + 1) * kblocks
                wk = 0
                for kb in range(group_kblocks):
                    if wk == 0:
                        yield ("wait", acc_empty[acc], acc_phase ^ 1)
                        for p in pipes:
                            assert p.acc[acc] == "free", f"window {acc} reopened in state {p.acc[acc]}"
                            p.acc[acc] = "accumulating"
                    yield ("wait", (ready[stage] if pair else full[0][stage]), phase)
                    last = (wk + 1 == max_window) or (kb == group_kblocks - 1)

                    def close(a=acc):
                        for r in range(ncta):
                            pipes[r].acc[a] = "full"
                            pipes[r].acc_drains[a] = 8
                            acc_full[r][a].arrive()
                    mma_async(sim, pipes, stage, [empty[r][stage] for r in range(ncta)], on_done=close if last else None)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
                    wk += 1
                    if last:
                        acc += 1
                        if acc == ACCS:
                            acc, acc_phase = 0, acc_phase ^ 1
                        wk = 0

    def accumulate_warp(r):
        acc, acc_phase = 0, 0
        for _ in range(tiles):
            for g in groups:
                windows = ((g + 1) * kblocks + max_window - 1) // max_window
                for _ in range(windows):
                    yield ("wait", acc_full[r][acc], acc_phase)
                    assert pipes[r].acc[acc] == "full", f"window {acc} drained in state {pipes[r].acc[acc]}"
                    yield ("sleep", sim.rng.randint(1, 40))
                    pipes[r].acc_drains[acc] -= 1
                    if pipes[r].acc_drains[acc] == 0:
                        pipes[r].acc[acc] = "free"
                    acc_empty[acc].arrive()
                    acc += 1
                    if acc == ACCS:
                        acc, acc_phase = 0, acc_phase ^ 1
            yield ("sleep", sim.rng.randint(1, 300))

    for r in range(ncta):
        sim.add(f"tma{r}", producer(r))
        if pair:
            sim.add(f"relay{r}", relay(r))
        for w in range(8):
            sim.add(f"acc{r}.{w}", accumulate_warp(r))
    sim.add("mma", mma())


@pytest.mark.parametrize("tiles,kblocks,window", [(1, 1, 4), (3, 5, 4), (2, 9, 4), (4, 3, 1)])
def test_default_sgemm_kernel_protocol(tiles, kblocks, window):
    """sgemm_tc_kernel (the hardware-validated default): the same roles with a 3-stage ring and four window accumulators; the split stage of
    all eight warps works in shared memory.  Kept here so that a change to that kernel's pipeline is checked before it costs GPU time."""
    for seed in range(8):
        sim = Sim(seed)
        build_ts(sim, False, tiles, kblocks, window, stages=3, accs=4)
        sim.run()


# ------------------------------------------------------------------------------------------------------------------------------------------
# sgemm_tc_kernel with the dynamic tile feed (round 2): the producer draws tile numbers from a counter shared by all CTAs of the launch and passes
# them on through a TILE_RING-entry ring (tile_full count 1; tile_empty count 1 MMA lane + 4 accumulate warps + 8 split warps); one packed
# counter per thread gives slot and parity; -1 ends every role; the CTA that reports its past-the-end draw last zeroes the counter pair.
def build_dynamic(sim, ctas, total_tiles, kblocks, window, stages=3, accs=4, ring=4):
    STAGES, ACCS, SPLIT_WARPS, RING = stages, accs, 8, ring
    sched = [0, 0]
    seen = {role: [] for role in ("tma", "mma", "acc", "split")}

    def cta(c):
        full = [Barrier(1) for _ in range(STAGES)]
        ready = [Barrier(SPLIT_WARPS) for _ in range(STAGES)]
        empty = [Barrier(1) for _ in range(STAGES)]
        acc_full = [Barrier(1) for _ in range(ACCS)]
        acc_empty = [Barrier(4) for _ in range(ACCS)]
        tile_full = [Barrier(1) for _ in range(RING)]
        tile_empty = [Barrier(1 + 4 + SPLIT_WARPS) for _ in range(RING)]
        tile_ring = [None] * RING
        pipe = Pipe(STAGES, ACCS)
        split_left = [0] * STAGES

        def next_tile(state):  # consumer side: state = [tr_count]
            slot = state[0] % RING
            yield ("wait", tile_full[slot], (state[0] // RING) & 1)
            t = tile_ring[slot]
            assert t is not None
            tile_empty[slot].arrive()
            state[0] += 1
            state.append(t)

        def producer():
            stage, phase, tr = 0, 0, 0

            def draw():
                t = sched[0]
                sched[0] += 1
                return t if t < total_tiles else -1
            tile = draw()
            while True:
                slot = tr % RING
                yield ("wait", tile_empty[slot], ((tr // RING) & 1) ^ 1)
                tile_ring[slot] = tile
                tile_full[slot].arrive()
                tr += 1
                if tile < 0:
                    break
                seen["tma"].append(tile)
                yield ("sleep", sim.rng.randint(1, 20))  # the atomic's round trip
                nxt = draw()
                for _ in range(kblocks):
                    yield ("wait", empty[stage], phase ^ 1)
                    assert pipe.stage[stage] == "free", f"TMA overwrites stage {stage} in state {pipe.stage[stage]}"
                    pipe.stage[stage] = "loading"

                    def landed(s=stage):
                        pipe.stage[s] = "landed"
                        split_left[s] = SPLIT_WARPS
                        full[s].arrive()
                    sim.later(20, 200, landed)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
                tile = nxt
            sched[1] += 1
            if sched[1] == ctas:
                sched[0] = sched[1] = 0

        def split_warp(w):
            stage, phase, st = 0, 0, [0]
            while True:
                yield from next_tile(st)
                if st.pop() < 0:
                    break
                for _ in range(kblocks):
                    yield ("wait", full[stage], phase)
                    assert pipe.stage[stage] == "landed", f"split reads stage {stage} in state {pipe.stage[stage]}"
                    yield ("sleep", sim.rng.randint(1, 30))
                    split_left[stage] -= 1
                    if split_left[stage] == 0:
                        pipe.stage[stage] = "ready"
                    ready[stage].arrive()
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1

        def mma():
            stage, acc, phase, acc_phase, st = 0, 0, 0, 0, [0]
            while True:
                yield from next_tile(st)
                t = st.pop()
                if t < 0:
                    break
                seen["mma"].append(t)
                wk = 0
                for kb in range(kblocks):
                    if wk == 0:
                        yield ("wait", acc_empty[acc], acc_phase ^ 1)
                        assert pipe.acc[acc] == "free", f"window {acc} reopened in state {pipe.acc[acc]}"
                        pipe.acc[acc] = "accumulating"
                    yield ("wait", ready[stage], phase)
                    last = (wk + 1 == window) or (kb == kblocks - 1)

                    def close(a=acc):
                        pipe.acc[a] = "full"
                        pipe.acc_drains[a] = 4
                        acc_full[a].arrive()
                    mma_async(sim, [pipe], stage, [empty[stage]], on_done=close if last else None)
                    stage += 1
                    if stage == STAGES:
                        stage, phase = 0, phase ^ 1
                    wk += 1
                    if last:
                        acc += 1
                        if acc == ACCS:
                            acc, acc_phase = 0, acc_phase ^ 1
                        wk = 0

        def accumulate_warp(w):
            acc_count, st = 0, [0]
            windows = (kblocks + window - 1) // window
            while True:
                yield from next_tile(st)
                t = st.pop()
                if t < 0:
                    break
                if w == 0:
                    seen["acc"].append(t)
                for _ in range(windows):
                    acc = acc_count % ACCS
                    yield ("wait", acc_full[acc], (acc_count // ACCS) & 1)
                    assert pipe.acc[acc] == "full", f"window {acc} drained in state {pipe.acc[acc]}"
                    yield ("sleep", sim.rng.randint(1, 40))
                    pipe.acc_drains[acc] -= 1
                    if pipe.acc_drains[acc] == 0:
                        pipe.acc[acc] = "free"
                    acc_empty[acc].arrive()
                    acc_count += 1
                yield ("sleep", sim.rng.randint(1, 300))  # the tile's global-memory epilogue

        def late(gen, delay):  # a CTA that gets its SM late (another launch held it)
            yield ("sleep", delay)
            yield from gen

        delay = sim.rng.choice([1, 1, 500, 3000])
        sim.add(f"tma{c}", late(producer(), delay))
        for w in range(SPLIT_WARPS):
            sim.add(f"split{c}.{w}", late(split_warp(w), delay))
        for w in range(4):
            sim.add(f"acc{c}.{w}", late(accumulate_warp(w), delay))
        sim.add(f"mma{c}", late(mma(), delay))

    for c in range(ctas):
        cta(c)
    return sched, seen


@pytest.mark.parametrize("ctas,tiles,kblocks,window", [(1, 1, 1, 4), (1, 6, 5, 4), (3, 10, 3, 4), (4, 3, 9, 4), (2, 7, 2, 1)])
def test_dynamic_tile_feed_protocol(ctas, tiles, kblocks, window):
    """Every tile is loaded, multiplied and written back exactly once, by the roles of one and the same CTA, whichever CTAs start late; every role
    ends; the counter pair is back at zero for the launch that takes the slot next."""
    for seed in range(8):
        sim = Sim(seed)
        sched, seen = build_dynamic(sim, ctas, tiles, kblocks, window)
        sim.run()
        assert sched == [0, 0]
        for role in ("tma", "mma", "acc"):
            assert sorted(seen[role]) == list(range(tiles)), (role, seen[role])


@pytest.mark.parametrize("pair", [False], ids=["one-cta"])
@pytest.mark.parametrize("tiles,kblocks,slices,max_window", [(1, 1, 2, 2048), (2, 3, 4, 2048), (3, 2, 8, 2048), (2, 5, 3, 4)])
def test_int8_dgemm_protocol(pair, tiles, kblocks, slices, max_window):
    for seed in range(8):
        sim = Sim(seed)
        build_i8(sim, pair, tiles, kblocks, slices, max_window)
        sim.run()


def test_the_model_catches_a_wrong_arrival_count():
    """Sanity of the checker itself: with one arrival too few on the ready barrier the pipeline must be reported dead."""
    sim = Sim(0)
    build_ts(sim, False, 2, 3, 4)
    # sabotage: drop one split warp (7 arrivals for a barrier initialised with 8)
    sim.roles = [r for r in sim.roles if r[0] != "split0.7"]
    with pytest.raises(AssertionError):
        sim.run()

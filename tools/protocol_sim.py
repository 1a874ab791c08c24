"""Discrete-event model of the attention-forward barrier protocols (csrc/attention_fwd.cu).

The ping-pong kernels coordinate one issue thread, 8 or 16 softmax warps, the TMA engine and
the (in-order) tensor pipe through mbarriers with phase parities, a ring of three O buffers,
two S / P buffers and small statistics rings.  A protocol slip shows up on hardware as a hang
or as silently stale data, and GPU time is the scarce resource - so the protocols are restated
here (control program + warp program, transcribed from the CUDA source) and executed under
randomised latencies.  Every buffer carries a version tag; the model raises on
  * a read that does not see the version the reader expects (stale / early data),
  * a write into a buffer that still has readers, or a read of a buffer being written,
  * a deadlock (nothing runnable before all programs finished),
  * an mbarrier wait that could alias (waiter two phases behind).
Usage: python tools/protocol_sim.py [pp|pp16|wg1|wg2|bwd|bwd_split] [trials]     (tests/test_protocol_sim.py runs it)
"""
from __future__ import annotations

import heapq
import random
import sys


class Hazard(Exception):
    pass


class MBar:
    def __init__(self, sim, name, count):
        self.sim, self.name, self.count = sim, name, count
        self.pending, self.phase, self.waiters = count, 0, []

    def arrive(self):
        self.pending -= 1
        if self.pending < 0:
            raise Hazard(f"{self.name}: more arrivals than its count")
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count
            ws, self.waiters = self.waiters, []
            for proc, parity, want in ws:
                self.sim.check_wait(self, proc, parity, want)

    def passed(self, parity):
        return (self.phase & 1) != parity


class HwBarrier:
    """bar.sync id, n: releases when n participants arrived."""

    def __init__(self, sim, n):
        self.sim, self.n, self.waiting = sim, n, []

    def arrive(self, proc):
        self.waiting.append(proc)
        if len(self.waiting) == self.n:
            ws, self.waiting = self.waiting, []
            for p in ws:
                self.sim.schedule(0.0, p)
            return True
        return False


class Res:
    def __init__(self, name):
        self.name, self.ver, self.writing, self.readers = name, None, False, 0


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.now, self.q, self.seq = 0.0, [], 0
        self.res, self.live = {}, 0
        self.tensor_fifo, self.tensor_busy = [], False

    # ---- scheduling
    def schedule(self, dt, proc, val=None):
        self.seq += 1
        heapq.heappush(self.q, (self.now + dt, self.seq, proc, val))

    def spawn(self, gen):
        self.live += 1
        self.schedule(0.0, gen)

    def jitter(self, lo, hi):
        return self.rng.uniform(lo, hi)

    def run(self):
        while self.q:
            self.now, _, proc, val = heapq.heappop(self.q)
            if callable(proc):
                proc()
                continue
            self.step(proc, val)
        if self.live:
            raise Hazard(f"deadlock: {self.live} program(s) never finished")

    def step(self, proc, val=None):
        try:
            act = proc.send(val)
        except StopIteration:
            self.live -= 1
            return
        kind = act[0]
        if kind == "delay":
            self.schedule(act[1], proc)
        elif kind == "wait":
            _, bar, parity, want = act
            self.check_wait(bar, proc, parity, want)
        elif kind == "hw":
            act[1].arrive(proc)
        else:
            raise AssertionError(kind)

    def check_wait(self, bar, proc, parity, want):
        # `want` = the 1-based phase the waiter is after (for the aliasing check only)
        if bar.phase > want:
            raise Hazard(f"{bar.name}: waiter for phase {want} sees phase {bar.phase} (parity alias)")
        if bar.passed(parity):
            if bar.phase != want:
                raise Hazard(f"{bar.name}: parity passed at phase {bar.phase}, expected {want}")
            self.schedule(0.0, proc)
        else:
            bar.waiters.append((proc, parity, want))

    # ---- versioned buffers
    def r(self, name):
        if name not in self.res:
            self.res[name] = Res(name)
        return self.res[name]

    def read_begin(self, name, expect):
        x = self.r(name)
        if x.writing:
            raise Hazard(f"read of {name} while it is being written (expect {expect})")
        if x.ver != expect:
            raise Hazard(f"read of {name}: holds {x.ver}, reader expects {expect}")
        x.readers += 1

    def read_end(self, name):
        self.r(name).readers -= 1

    def write_begin(self, name, ver):
        x = self.r(name)
        if x.readers or x.writing:
            raise Hazard(f"write of {name} (-> {ver}) while busy: readers={x.readers} writing={x.writing}, holds {x.ver}")
        x.writing = True

    def write_end(self, name, ver):
        x = self.r(name)
        x.writing, x.ver = False, ver

    def write(self, name, ver):
        self.write_begin(name, ver)
        self.write_end(name, ver)

    # ---- engines
    def tma(self, names, ver, bar):
        for n in names:
            self.write_begin(n, ver)

        def done():
            for n in names:
                self.write_end(n, ver)
            bar.arrive()
        self.schedule(self.jitter(300, 2500), done)

    def tensor_issue(self, op):
        """op = ('mma', reads[(name,ver)], writes[(name,ver)], dur) | ('commit', bar)"""
        self.tensor_fifo.append(op)
        if not self.tensor_busy:
            self.tensor_next()

    def tensor_next(self):
        if not self.tensor_fifo:
            self.tensor_busy = False
            return
        self.tensor_busy = True
        op = self.tensor_fifo.pop(0)
        if op[0] == "commit":
            bar = op[1]

            def fire():
                bar.arrive()
                self.tensor_next()
            self.schedule(self.jitter(5, 60), fire)
            return
        _, reads, writes, dur = op
        for n, v in reads:
            self.read_begin(n, v)
        for n, v in writes:
            self.write_begin(n, v)

        def fin():
            for n, _ in reads:
                self.read_end(n)
            for n, v in writes:
                self.write_end(n, v)
            self.tensor_next()
        self.schedule(dur * self.jitter(0.7, 1.5), fin)


def wait(bar, parity, want):
    return ("wait", bar, parity, want)


# ------------------------------------------------------------------------------------------
# protocol transcription.  mode: 'pp' (attn_fwd_pp_kernel, 8 warps, CTA-wide row barrier), 'pp16'
# (attn_fwd_pp16_kernel, 16 warps, 128-thread barrier per lane quarter),
# 'wg1' (attn_fwd_wg_kernel<1>), 'wg2' (attn_fwd_wg_kernel<2>)
def build(sim: Sim, mode: str, items: int, T: int, mutate: str = ""):
    """`mutate` breaks the protocol on purpose (the model must then raise): 'no_e_bar' (PV does not
    wait for the epilogue that still reads its ring slot), 'ring4' (4-slot statistics ring in the
    warpgroup variants), 'no_p_free' (P buffer rewritten without waiting for the PV that reads it - harmless on an
    in-order tensor pipe, S(g) is queued behind PV(g-2)), 'no_s_free' (S(g) issued without its own wait
    for the softmax warps of block g-2 - also implied, by the wait inside issue_pv(g-2)), 'no_q_wait'
    (Q / K reloaded without waiting for the S MMA that reads them)."""
    halves = 2 if mode in ("pp", "wg2") else 1
    nwarps = 16 if mode in ("wg2", "pp16") else 8
    wg = mode not in ("pp", "pp16")
    ngrp = 4 if mode == "pp16" else 2             # column groups per row in the pp kernels
    ring = 3 if mutate == "ring4" else 7          # statistics ring mask of the warpgroup variants
    B = lambda name, c: MBar(sim, name, c)
    k_bar, v_bar, q_bar = B("k_bar", 1), B("v_bar", 1), B("q_bar", 1)
    s_bar = [B(f"s_bar{i}", 1) for i in range(2)]
    p_bar = [B(f"p_bar{i}", (4 * halves) if wg else nwarps) for i in range(2)]
    o_bar = [B(f"o_bar{i}", 1) for i in range(3)]
    e_bar = [B(f"e_bar{i}", nwarps) for i in range(2)]
    all_bar = HwBarrier(sim, nwarps)
    pair_bar = {}

    def p_parts(buf):          # the smem parts of P buffer `buf` (one per writing warp)
        if wg:
            return [f"P{buf}.q{q}.h{h}" for q in range(4) for h in range(halves)]
        return [f"P{buf}.q{q}.h{h}" for q in range(4) for h in range(ngrp)]

    def control():
        g = tt = 0
        # phases completed so far per barrier family (for the alias check)
        sim.tma(["K"], 0, k_bar)
        sim.tma(["Q"], (0, 0), q_bar)
        sim.tma(["V"], 0, v_bar)

        def issue_pv(gb, item):
            yield wait(p_bar[gb & 1], (gb >> 1) & 1, (gb >> 1) + 1)
            if wg and gb >= 3 and mutate != "no_e_bar":
                te = (gb - 3) >> 1
                yield wait(e_bar[te & 1], (te >> 1) & 1, (te >> 1) + 1)
            yield ("delay", sim.jitter(20, 200))
            reads = [(n, gb) for n in p_parts(gb & 1)] + [("V", item)]
            sim.tensor_issue(("mma", reads, [(f"O{gb % 3}", gb)], 500))
            sim.tensor_issue(("commit", o_bar[gb % 3]))

        for item in range(items):
            nxt = item + 1
            yield wait(k_bar, item & 1, item + 1)
            for t in range(T):
                for kb in range(2):
                    if kb == 0:
                        yield wait(q_bar, tt & 1, tt + 1)
                        tt += 1
                    if g >= 2 and mutate != "no_s_free":
                        yield wait(p_bar[g & 1], ((g - 2) >> 1) & 1, ((g - 2) >> 1) + 1)
                    yield ("delay", sim.jitter(20, 200))
                    sim.tensor_issue(("mma", [("Q", (item, t)), ("K", item)], [(f"S{g & 1}", g)], 500))
                    sim.tensor_issue(("commit", s_bar[g & 1]))
                    if kb == 1:
                        if t + 1 < T:
                            if mutate != "no_q_wait":
                                yield wait(s_bar[g & 1], (g >> 1) & 1, (g >> 1) + 1)
                            sim.tma(["Q"], (item, t + 1), q_bar)
                        elif nxt < items:
                            if mutate != "no_q_wait":
                                yield wait(s_bar[g & 1], (g >> 1) & 1, (g >> 1) + 1)
                            sim.tma(["K"], nxt, k_bar)
                            sim.tma(["Q"], (nxt, 0), q_bar)
                    if t == 0 and kb == 1:
                        yield wait(v_bar, item & 1, item + 1)
                    if not (t == 0 and kb == 0):
                        yield from issue_pv(g - 1, item)
                    g += 1
            yield from issue_pv(g - 1, item)
            yield wait(o_bar[(g - 1) % 3], ((g - 1) // 3) & 1, (g - 1) // 3 + 1)
            if nxt < items:
                sim.tma(["V"], nxt, v_bar)

    def stat(slot, q, h):
        return f"stat{slot}.q{q}.h{h}"

    def warp_pp(q, h):
        g = 0
        for item in range(items):
            sim.write(f"tables.w{q}{h}", item)
            yield ("hw", all_bar)
            for t in range(T):
                for kb in range(2):
                    yield wait(s_bar[g & 1], (g >> 1) & 1, (g >> 1) + 1)
                    sim.read_begin(f"S{g & 1}", g)
                    yield ("delay", sim.jitter(200, 1500))
                    sim.write(f"max{g & 1}.q{q}.h{h}", g)
                    if mode == "pp16":                      # 128-thread named barrier per lane quarter
                        if ("quad", q) not in pair_bar:
                            pair_bar[("quad", q)] = HwBarrier(sim, 4)
                        yield ("hw", pair_bar[("quad", q)])
                    else:
                        yield ("hw", all_bar)
                    for o in range(1, ngrp):
                        sim.read_begin(f"max{g & 1}.q{q}.h{(h + o) % ngrp}", g)
                        sim.read_end(f"max{g & 1}.q{q}.h{(h + o) % ngrp}")
                    if g >= 2 and mutate != "no_p_free":
                        yield wait(o_bar[(g - 2) % 3], ((g - 2) // 3) & 1, (g - 2) // 3 + 1)
                    sim.write_begin(f"P{g & 1}.q{q}.h{h}", g)
                    yield ("delay", sim.jitter(200, 2500))
                    sim.write_end(f"P{g & 1}.q{q}.h{h}", g)
                    sim.read_end(f"S{g & 1}")
                    sim.write(stat(g & 3, q, h), g)
                    p_bar[g & 1].arrive()
                    if kb == 0 and t > 0:
                        yield from epi_pp(q, h, g - 2)
                    g += 1
            yield from epi_pp(q, h, g - 2)

    def epi_pp(q, h, g0):
        g1 = g0 + 1
        yield wait(o_bar[g0 % 3], (g0 // 3) & 1, g0 // 3 + 1)
        yield wait(o_bar[g1 % 3], (g1 // 3) & 1, g1 // 3 + 1)
        for gg in (g0, g1):
            for o in range(1, ngrp):
                sim.read_begin(stat(gg & 3, q, (h + o) % ngrp), gg)
                sim.read_end(stat(gg & 3, q, (h + o) % ngrp))
            sim.read_begin(f"O{gg % 3}", gg)
        yield ("delay", sim.jitter(50, 800))
        for gg in (g0, g1):
            sim.read_end(f"O{gg % 3}")

    def warp_wg(q, blk, h):
        gt = 0
        for item in range(items):
            sim.write(f"tables.w{q}{blk}{h}", item)
            yield ("hw", all_bar)
            for t in range(T):
                gb = 2 * gt + blk
                yield wait(s_bar[blk], gt & 1, gt + 1)
                sim.read_begin(f"S{blk}", gb)
                yield ("delay", sim.jitter(200, 2500))
                if halves == 2:
                    par = gt & 1
                    sim.write(f"max{par}.b{blk}.q{q}.h{h}", gb)
                    key = (blk, q)
                    if key not in pair_bar:
                        pair_bar[key] = HwBarrier(sim, 2)
                    yield ("hw", pair_bar[key])
                    sim.read_begin(f"max{par}.b{blk}.q{q}.h{h ^ 1}", gb)
                    sim.read_end(f"max{par}.b{blk}.q{q}.h{h ^ 1}")
                if gb >= 2 and mutate != "no_p_free":
                    yield wait(o_bar[(gb - 2) % 3], ((gb - 2) // 3) & 1, (gb - 2) // 3 + 1)
                sim.write_begin(f"P{blk}.q{q}.h{h}", gb)
                yield ("delay", sim.jitter(200, 2500))
                sim.write_end(f"P{blk}.q{q}.h{h}", gb)
                sim.read_end(f"S{blk}")
                sim.write(stat(gb & ring, q, h), gb)
                p_bar[blk].arrive()
                if t > 0:
                    yield from epi_wg(q, gt - 1)
                gt += 1
            yield from epi_wg(q, gt - 1)

    def epi_wg(q, gtile):
        g0, g1 = 2 * gtile, 2 * gtile + 1
        yield wait(o_bar[g0 % 3], (g0 // 3) & 1, g0 // 3 + 1)
        yield wait(o_bar[g1 % 3], (g1 // 3) & 1, g1 // 3 + 1)
        for gg in (g0, g1):
            for hh in range(halves):
                sim.read_begin(stat(gg & ring, q, hh), gg)
                sim.read_end(stat(gg & ring, q, hh))
            sim.read_begin(f"O{gg % 3}", gg)
        yield ("delay", sim.jitter(50, 800))          # tcgen05.ld of both O slices
        for gg in (g0, g1):
            sim.read_end(f"O{gg % 3}")
        e_bar[gtile & 1].arrive()
        yield ("delay", sim.jitter(50, 800))          # global stores

    sim.spawn(control())
    if mode in ("pp", "pp16"):
        for q in range(4):
            for h in range(ngrp):
                sim.spawn(warp_pp(q, h))
    else:
        for q in range(4):
            for blk in range(2):
                for h in range(halves):
                    sim.spawn(warp_wg(q, blk, h))


# ------------------------------------------------------------------------------------------
# attention backward (csrc/attention_bwd.cu).  mode 'bwd': the shipped kernel; 'bwd_split': the
# planned variant whose S/dP tile is produced in two 64-key halves with their own barriers so the
# next pair's first half is computed by the tensor pipe while the softmax warps still work on the
# second half of the current pair.
def build_bwd(sim: Sim, mode: str, items: int, ntiles: int, mutate: str = ""):
    split = mode == "bwd_split"
    nwarps = 8
    B = lambda name, c: MBar(sim, name, c)
    kv_bar = [B(f"kv_bar{i}", 1) for i in range(2)]
    qdo_bar = [B(f"qdo_bar{i}", 1) for i in range(2)]
    s_bar = [B(f"s_bar{i}", 1) for i in range(2)]        # 'bwd' uses s_bar[0] only
    hfree = B("hfree0", nwarps)
    pds_bar, g_bar, dq_bar = B("pds_bar", nwarps), B("g_bar", 1), B("dq_bar", 1)
    all_bar = HwBarrier(sim, nwarps)
    pairs = [(bh, j, i) for bh in range(items) for j in range(ntiles) for i in range(ntiles)]
    ksteps = [(bh, j) for bh in range(items) for j in range(ntiles)]
    npairs = len(pairs)
    kc_of = {pc: pc // ntiles for pc in range(npairs)}
    parts = [f"PdS.h{h}.w{w}" for h in range(2) for w in range(nwarps)]

    def control():
        def load_kv(kc):
            sim.tma([f"KV{kc & 1}"], ksteps[kc], kv_bar[kc & 1])

        def load_qdo(pc):
            bh, j, i = pairs[pc]
            sim.tma([f"QdO{pc & 1}"], (bh, i, pc), qdo_bar[pc & 1])

        def issue_scores(pc, halves):
            bh, j, i = pairs[pc]
            kc = kc_of[pc]
            reads = [(f"QdO{pc & 1}", (bh, i, pc)), (f"KV{kc & 1}", ksteps[kc])]
            for h in halves:
                sim.tensor_issue(("mma", reads, [(f"S{h}", pc)], 256 if split else 512))
                sim.tensor_issue(("commit", s_bar[h]))

        load_kv(0)
        load_qdo(0)
        yield wait(kv_bar[0], 0, 1)
        yield wait(qdo_bar[0], 0, 1)
        issue_scores(0, (0, 1) if split else (0,))
        for pc in range(npairs):
            bh, j, i = pairs[pc]
            kc = kc_of[pc]
            has_next = pc + 1 < npairs
            if pc >= 1:
                yield wait(g_bar, (pc - 1) & 1, pc)
            if i == 0 and kc + 1 < len(ksteps):
                load_kv(kc + 1)
            if has_next:
                load_qdo(pc + 1)

            def next_ready():
                nkc = kc_of[pc + 1]
                if pairs[pc + 1][2] == 0:
                    yield wait(kv_bar[nkc & 1], (nkc >> 1) & 1, (nkc >> 1) + 1)
                yield wait(qdo_bar[(pc + 1) & 1], ((pc + 1) >> 1) & 1, ((pc + 1) >> 1) + 1)

            if split and has_next:
                if mutate != "no_hfree":
                    yield wait(hfree, pc & 1, pc + 1)
                yield from next_ready()
                yield ("delay", sim.jitter(20, 300))
                issue_scores(pc + 1, (0,))
            yield wait(pds_bar, pc & 1, pc + 1)
            yield ("delay", sim.jitter(20, 300))
            kvn, qn = f"KV{kc & 1}", f"QdO{pc & 1}"
            kvv, qv = ksteps[kc], (bh, i, pc)
            pds_reads = [(n, pc) for n in parts]
            sim.tensor_issue(("mma", pds_reads + [(kvn, kvv)], [("dQ", pc)], 256))
            sim.tensor_issue(("commit", dq_bar))
            if split and has_next:
                issue_scores(pc + 1, (1,))
            sim.tensor_issue(("mma", pds_reads + [(qn, qv)], [("dKV", pc)], 512))
            sim.tensor_issue(("commit", g_bar))
            if (not split) and has_next:
                yield from next_ready()
                issue_scores(pc + 1, (0,))
        yield wait(g_bar, (npairs - 1) & 1, npairs)

    def warp(w):
        pc = 0
        for bh in range(items):
            for j in range(ntiles):
                if w < 4:
                    sim.write(f"tables.w{w}", (bh, j))
                yield ("hw", all_bar)
                for i in range(ntiles):
                    for h in ((0, 1) if split else (0,)):
                        yield wait(s_bar[h], pc & 1, pc + 1)
                        sim.read_begin(f"S{h}", pc)
                        if not split and pc >= 1:
                            yield wait(g_bar, (pc - 1) & 1, pc)
                        sim.read_begin(f"tables.w{w & 3}", (bh, j))
                        yield ("delay", sim.jitter(150, 1500))
                        sim.read_end(f"tables.w{w & 3}")
                        sim.read_end(f"S{h}")
                        if split and h == 0:
                            hfree.arrive()
                            if pc >= 1 and mutate != "no_g_wait":
                                yield wait(g_bar, (pc - 1) & 1, pc)
                        names = [f"PdS.h{h}.w{w}"] if split else [f"PdS.h0.w{w}", f"PdS.h1.w{w}"]
                        for nme in names:
                            sim.write_begin(nme, pc)
                        yield ("delay", sim.jitter(20, 300))
                        for nme in names:
                            sim.write_end(nme, pc)
                    pds_bar.arrive()
                    if j > 0:
                        sim.read_begin(f"ws.{bh}.{i}.w{w}", j - 1)
                        sim.read_end(f"ws.{bh}.{i}.w{w}")
                    yield wait(dq_bar, pc & 1, pc + 1)
                    sim.read_begin("dQ", pc)
                    yield ("delay", sim.jitter(50, 600))
                    sim.read_end("dQ")
                    sim.write(f"ws.{bh}.{i}.w{w}", j)
                    pc += 1
                yield wait(g_bar, (pc - 1) & 1, pc)
                sim.read_begin("dKV", pc - 1)
                yield ("delay", sim.jitter(50, 600))
                sim.read_end("dKV")

    sim.spawn(control())
    for w in range(nwarps):
        sim.spawn(warp(w))


def run_bwd(mode: str, trials: int = 100, seed0: int = 0, mutate: str = ""):
    for trial in range(trials):
        for ntiles in (1, 2, 3):
            sim = Sim(seed0 + 1000 * trial + ntiles)
            build_bwd(sim, mode, items=sim.rng.choice([1, 2, 3]), ntiles=ntiles, mutate=mutate)
            try:
                sim.run()
            except Hazard as e:
                raise Hazard(f"[{mode} ntiles={ntiles} trial={trial}] {e}") from None
    return True


def run(mode: str, trials: int = 200, seed0: int = 0, mutate: str = ""):
    for trial in range(trials):
        for T in (1, 2, 3):
            sim = Sim(seed0 + 1000 * trial + T)
            build(sim, mode, items=sim.rng.choice([1, 2, 3, 4]), T=T, mutate=mutate)
            try:
                sim.run()
            except Hazard as e:
                raise Hazard(f"[{mode} T={T} trial={trial}] {e}") from None
    return True


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "pp"
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    (run_bwd if mode.startswith("bwd") else run)(mode, trials)
    print(f"{mode}: {trials} randomised trials x 1..3 tiles: no hazard, no deadlock")

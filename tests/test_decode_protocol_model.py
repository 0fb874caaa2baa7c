"""CPU: a scheduling model of the barrier-free persistent decode kernel (csrc/gpt2.cu: gpt_decode_kernel). Every CTA is a coroutine
that executes the kernel's phase sequence for the rows it owns -- poll a tagged vector for an epoch, publish its own rows with the
next epoch -- and a random scheduler interleaves the 148 of them. The model checks what DESIGN.md 3b claims about the protocol:
no deadlock, no reader ever finds a word already overwritten by a LATER epoch (write-after-read safety follows from the data flow:
a CTA takes part in a phase only if it owns rows there), and a STOP epoch published by the pick phase terminates every CTA.
It mirrors the ownership arithmetic of the kernel (rows dealt warp by warp, 16 warps per CTA), so a change of that arithmetic in
the kernel has to be made here too."""
import random

import pytest

G, NW = 148, 16
TW = G * NW
STOP = 1 << 40


class Hazard(Exception):
    pass


def owners(n_rows):
    """producer CTA of every row: row n belongs to warp n % TW, i.e. CTA (n % TW) // NW"""
    return [(n % TW) // NW for n in range(n_rows)]


class Vec:
    """a tagged vector: one epoch per row"""

    def __init__(self, n):
        self.ep = [0] * n

    def ready(self, rows, want):
        ok = True
        for r in rows:
            e = self.ep[r]
            if e == STOP:
                return "stop"
            if e > want:
                raise Hazard(f"row {r} already carries epoch {e} > {want}")
            ok = ok and e == want
        return ok


def make_ctas(D, FF, Vm, H, L, tokens, stop_at=None):
    th, tqkv, tatt, tff, tlog = Vec(D), Vec(3 * D), Vec(D), Vec(FF), Vec(Vm)
    th.ep = [1] * D                                      # gpt_tag_init_kernel
    own = {"qkv": owners(3 * D), "d": owners(D), "fc": owners(FF), "head": owners(Vm)}
    rows_of = {k: [[r for r, c in enumerate(v) if c == cta] for cta in range(G)] for k, v in own.items()}
    all_d, all_ff, all_v = range(D), range(FF), range(Vm)
    log = {"picked": 0, "exited": 0}

    def cta(c):
        gw0 = c * NW
        in_qkv, in_d, in_fc, in_head, in_att = gw0 < 3 * D, gw0 < D, gw0 < FF, gw0 < Vm, c < H

        def poll(vec, rows, want):
            while True:
                r = vec.ready(rows, want)
                if r == "stop":
                    return False
                if r:
                    return True
                yield

        for t in range(tokens):
            eh0 = 1 + t * (2 * L + 1)
            for l in range(L):
                ev, eh = 1 + t * L + l, eh0 + 2 * l
                if in_qkv:
                    ok = yield from poll(th, all_d, eh)
                    if not ok:
                        log["exited"] += 1
                        return
                    for r in rows_of["qkv"][c]:
                        tqkv.ep[r] = ev
                    yield
                if in_att:
                    mine = [p * D + c * 64 + d for p in range(3) for d in range(64)]
                    ok = yield from poll(tqkv, mine, ev)
                    assert ok
                    for d in range(64):
                        tatt.ep[c * 64 + d] = ev
                    yield
                if in_d:
                    ok = yield from poll(tatt, all_d, ev)
                    assert ok
                    for r in rows_of["d"][c]:
                        th.ep[r] = eh + 1
                    yield
                if in_fc:
                    ok = yield from poll(th, all_d, eh + 1)
                    if not ok:
                        log["exited"] += 1
                        return
                    for r in rows_of["fc"][c]:
                        tff.ep[r] = ev
                    yield
                if in_d:
                    ok = yield from poll(tff, all_ff, ev)
                    assert ok
                    for r in rows_of["d"][c]:
                        th.ep[r] = eh + 2
                    yield
            if in_head:
                ok = yield from poll(th, all_d, eh0 + 2 * L)
                if not ok:
                    log["exited"] += 1
                    return
                for r in rows_of["head"][c]:
                    tlog.ep[r] = 1 + t
                yield
            if c == 0:
                ok = yield from poll(tlog, all_v, 1 + t)
                assert ok
                log["picked"] += 1
                stop = stop_at is not None and t == stop_at
                for r in all_d:
                    th.ep[r] = STOP if stop else eh0 + 2 * L + 1
                yield
        log["exited"] += 1

    return [cta(c) for c in range(G)], log


def run(D, FF, Vm, H, L, tokens, seed, stop_at=None):
    rng = random.Random(seed)
    ctas, log = make_ctas(D, FF, Vm, H, L, tokens, stop_at)
    live = list(range(G))
    idle = 0
    while live:
        i = rng.randrange(len(live)) if rng.random() < 0.9 else 0            # mostly random, sometimes favour the oldest survivor
        c = live[i]
        try:
            next(ctas[c])
            idle += 1
        except StopIteration:
            live.pop(i)
            idle = 0
        if idle > 400000:
            raise AssertionError(f"no CTA finished in {idle} scheduling steps: deadlock? live={live[:8]}")
    return log


@pytest.mark.parametrize("shape", [dict(D=512, FF=2048, Vm=130, H=8, L=3), dict(D=1280, FF=5120, Vm=8194, H=20, L=2)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_no_deadlock_no_overwrite_before_read(shape, seed):
    log = run(tokens=3, seed=seed, **shape)
    assert log["picked"] == 3 and log["exited"] == G


@pytest.mark.parametrize("shape", [dict(D=512, FF=2048, Vm=130, H=8, L=3), dict(D=1280, FF=5120, Vm=8194, H=20, L=2)])
def test_stop_epoch_terminates_every_cta(shape):
    log = run(tokens=5, seed=7, stop_at=1, **shape)
    assert log["picked"] == 2 and log["exited"] == G


def test_model_catches_an_overwrite():
    """Sanity of the checker itself: a vector that already carries a later epoch than the one a reader waits for is reported."""
    v = Vec(4)
    v.ep = [3, 3, 5, 3]
    with pytest.raises(Hazard):
        v.ready(range(4), 3)
    v.ep = [3, 3, STOP, 3]
    assert v.ready(range(4), 3) == "stop"

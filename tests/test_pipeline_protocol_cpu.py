"""Model check of the cross-rank buffer-reuse protocol of the chunk pipeline (channel_b200/csrc/chb_api.cu,
convolutions_all).

On several GPUs the pencil transposes are plain stores into the peers' work buffers (zfwd -> every rank's Ar, xpass -> every
rank's Br), separated from their readers by flag barriers only; there is no barrier that says "you may overwrite my buffer".
That this is safe rests on the ORDER in which every rank enqueues kernels, barriers and event waits on its two streams.
This test restates that order (the launch schedule of convolutions_all, one lane and two lanes) as data, runs P ranks
under a random scheduler - any stream of any rank whose head operation is enabled may go next, a barrier completes only once
every rank has reached it - and checks at every kernel the four hazards:

  zfwd(c)  of rank q writes Ar[r][lane]   -> rank r's xpass(c - nlanes) (the previous reader) must be complete
  xpass(c) of rank r reads  Ar[r][lane]   -> every rank's zfwd(c) must be complete
  xpass(c) of rank q writes Br[r][lane]   -> rank r's zbwd(c - nlanes) must be complete
  zbwd(c)  of rank r reads  Br[r][lane]   -> every rank's xpass(c) must be complete

and that the run never deadlocks.  It also shows that the protocol is not trivially safe: dropping the wait of the z-pass
stream for the lane's x-pass event (the line `cudaStreamWaitEvent(h->sB, ln.evA)` in zfwd_part) is caught in the CFL pre-pass.  The model is a
restatement, kept next to the code it mirrors: if convolutions_all changes its order, change `schedule()` with it.  The
same schedule runs on real hardware in tests/mgpu_worker.py (2, 4 and 8 ranks against the oracle).
"""
import random

import pytest


def schedule(nch, nlanes, products=True, drop_evA_wait=False, nsweeps=1):
    """Operations one rank enqueues for `nsweeps` sweeps of `nch` chunks: {stream: [op, ...]}; op = (kind, gid, waits) where
    gid = sweep * nch + chunk identifies the chunk globally and waits is a list of events that must have been recorded (on
    this rank) before the op may start.  Streams: 'A' and 'B' (two lanes) or 'S' (one lane).  Mirrors convolutions_all;
    between two sweeps every stream waits for the end of both streams' previous sweep (the join into the handle's stream
    and the fork of the next sweep).  Returns (streams, lane_of) with lane_of[gid] = the lane the chunk uses."""
    streams = {"S": []} if nlanes == 1 else {"A": [], "B": []}
    lane_of = {}
    for sw in range(nsweeps):
        fork = [("end", sw - 1, st) for st in streams] if sw > 0 else []
        first = {st: True for st in streams}

        def add(st, kind, gid, waits):
            if first[st]:
                waits = waits + fork
                first[st] = False
            streams[st].append((kind, gid, waits))

        g0 = sw * nch
        for c in range(nch):
            lane_of[g0 + c] = c % nlanes
        if nlanes == 1:
            for c in range(nch):
                add("S", "zfwd", g0 + c, []); add("S", "barA", g0 + c, []); add("S", "xpass", g0 + c, []); add("S", "barB", g0 + c, [])
                if products:
                    add("S", "zbwd", g0 + c, [])
        else:
            def zfwd_part(c):
                waits = [("evA", g0 + c - nlanes)] if (c >= nlanes and not drop_evA_wait) else []
                add("B", "zfwd", g0 + c, waits)
                add("B", "barA", g0 + c, [])
                add("B", "rec_evZ", g0 + c, [])

            zfwd_part(0)
            for c in range(nch):
                add("A", "xpass", g0 + c, [("evZ", g0 + c)])
                add("A", "barB", g0 + c, [])
                add("A", "rec_evA", g0 + c, [])
                if c + 1 < nch:
                    zfwd_part(c + 1)
                if products:
                    add("B", "zbwd", g0 + c, [("evA", g0 + c)])
        for st in streams:
            streams[st].append(("rec_end", sw, [st]))
    return streams, lane_of


class Violation(Exception):
    pass


def run(P, nch, nlanes, seed, products=True, drop_evA_wait=False, nsweeps=1, max_steps=400000):
    rng = random.Random(seed)
    sched, lane_of = schedule(nch, nlanes, products, drop_evA_wait, nsweeps)
    prev_in_lane, last = {}, {}
    for gid in sorted(lane_of):                 # the chunk that used the same lane before
        prev_in_lane[gid] = last.get(lane_of[gid])
        last[lane_of[gid]] = gid
    queues = [{s: list(ops) for s, ops in sched.items()} for _ in range(P)]
    done = [set() for _ in range(P)]            # (kind, gid) of completed kernels per rank
    events = [set() for _ in range(P)]          # recorded events per rank
    arrived = {}                                # barrier (kind, gid) -> set of ranks that have reached it
    waiting = [dict() for _ in range(P)]        # stream -> barrier it is spinning in

    def all_done(kind, gid):
        return all((kind, gid) in done[q] for q in range(P))

    for _ in range(max_steps):
        enabled = []
        for r in range(P):
            for s, ops in queues[r].items():
                if not ops:
                    continue
                kind, gid, waits = ops[0]
                if s in waiting[r]:                               # spinning in a barrier: enabled when every rank has arrived
                    if len(arrived[waiting[r][s]]) == P:
                        enabled.append((r, s))
                    continue
                if kind == "rec_end" or all(ev in events[r] for ev in waits):
                    enabled.append((r, s))
        if not enabled:
            if all(not ops for q in queues for ops in q.values()):
                return True
            raise Violation(f"deadlock: P={P} nch={nch} nlanes={nlanes} seed={seed}")
        r, s = rng.choice(enabled)
        kind, gid, waits = queues[r][s][0]
        prev = prev_in_lane.get(gid)
        if kind in ("barA", "barB"):
            key = (kind, gid)
            if s not in waiting[r]:                               # arrive (publish the flag), then spin
                arrived.setdefault(key, set()).add(r)
                waiting[r][s] = key
                continue
            del waiting[r][s]                                     # every rank has arrived: the barrier kernel ends
        elif kind == "zfwd":                                      # stores into Ar[q][lane] of every rank q
            if prev is not None and not all_done("xpass", prev):
                raise Violation(f"rank {r} zfwd({gid}) overwrites an Ar that some rank's xpass({prev}) has not read")
        elif kind == "xpass":
            if not all_done("zfwd", gid):
                raise Violation(f"rank {r} xpass({gid}) reads its Ar before every rank's zfwd({gid}) has stored into it")
            if products and prev is not None and not all_done("zbwd", prev):
                raise Violation(f"rank {r} xpass({gid}) overwrites a Br that some rank's zbwd({prev}) has not read")
        elif kind == "zbwd":
            if not all_done("xpass", gid):
                raise Violation(f"rank {r} zbwd({gid}) reads its Br before every rank's xpass({gid}) has stored into it")
        elif kind == "rec_evZ":
            events[r].add(("evZ", gid))
        elif kind == "rec_evA":
            events[r].add(("evA", gid))
        elif kind == "rec_end":
            events[r].add(("end", gid, waits[0]))
        if kind in ("zfwd", "xpass", "zbwd"):
            done[r].add((kind, gid))
        queues[r][s].pop(0)
    raise Violation("step limit")


@pytest.mark.parametrize("nlanes", [1, 2])
@pytest.mark.parametrize("P", [2, 3, 8])
def test_buffer_reuse_protocol_is_safe_under_any_interleaving(P, nlanes):
    for nch in (1, 2, 3, 5, 9):
        for seed in range(40):
            assert run(P, nch, nlanes, seed)
            assert run(P, nch, nlanes, seed, products=False)      # the CFL pre-pass: no zbwd, Br never read
        for seed in range(10):                                    # sweep after sweep: the first chunks reuse the lanes of the last
            assert run(P, nch, nlanes, seed, nsweeps=3)


def test_the_model_catches_a_missing_wait():
    """Without sB's wait for the lane's x-pass event a fast rank's zfwd(c) can overwrite the Ar another rank still reads.
    In a buildrhs sweep the wait is implied (zbwd(c-2), which precedes zfwd(c) on sB, waits for the same event); in the CFL
    pre-pass, which has no zbwd, it is what keeps the sweep safe."""
    for seed in range(100):
        assert run(3, 6, 2, seed, products=True, drop_evA_wait=True)
    caught = 0
    for seed in range(200):
        try:
            run(3, 6, 2, seed, products=False, drop_evA_wait=True)
        except Violation as e:
            assert "overwrites an Ar" in str(e)
            caught += 1
    assert caught > 0

"""Model check of the FlowProjection pipeline's scheduling protocol (flow_projection_fast.cu, fp_pipeline_kernel)
on the CPU: queue order, look-ahead of two positions per CTA, per-frame completion counters, one completion signal
per RUN of same-kind items, dependency waits.  Random CTA speeds; asserts

  * no deadlock (every item of every frame runs),
  * an item only starts when its dependency is complete (splat(f) after average(f-3), average(f) after splat(f),
    fill(f) after average(f)),
  * a CTA never waits while it holds completion signals it has not sent,
  * accumulator slot f % 3 is never written by splat(f) while average(f-3) still reads it.

    python tools/fp_pipeline_model.py [B nS nA n_cta seeds]
"""
import random
import sys


def decode(q, B, nS, nA, fillhole, total):
    if q >= total:
        return (-2, 0, 0)
    per = 2 * nA + nS
    c, r = divmod(q, per)
    if r < nA:
        t, f, tile = 1, c - 2, r
    elif r < 2 * nA:
        t, f, tile = 2, c - 3, r - nA
    else:
        t, f, tile = 0, c, r - 2 * nA
    if f < 0 or f >= B or (t == 2 and not fillhole):
        t = -1
    return (t, f, tile)


def simulate(B, nS, nA, n_cta, fillhole=True, seed=0):
    rng = random.Random(seed)
    total = (B + 3) * (2 * nA + nS)
    head = [0]
    done_S, done_A = [0] * B, [0] * B
    ran = {0: [0] * B, 1: [0] * B, 2: [0] * B}
    reading_A = [0] * B          # average tiles of frame f currently running (they read slot f % 3)

    def fetch():
        q = head[0]
        head[0] += 1
        return decode(q, B, nS, nA, fillhole, total)

    class Cta(object):
        def __init__(self):
            self.it, self.nx = fetch(), fetch()
            self.known_S = self.known_A = -1
            self.pending = 0
            self.state = "start"   # start -> (wait) -> run -> finish
            self.dep = None

    ctas = [Cta() for _ in range(n_cta)]
    live = set(range(n_cta))
    idle_rounds = 0
    while live:
        progressed = False
        for k in rng.sample(sorted(live), len(live)):
            c = ctas[k]
            t, f, _ = c.it
            if t == -2:
                assert c.pending == 0
                live.discard(k)
                progressed = True
                continue
            if rng.random() < 0.35:          # this CTA is slow this round
                continue
            if c.state == "start":
                c.fetched = fetch()          # the position after next
                c.dep = None
                if t == 0 and f >= 3 and c.known_A < f - 3:
                    c.dep, c.known_A = ("A", f - 3, nA), f - 3
                if t == 1 and c.known_S < f:
                    c.dep, c.known_S = ("S", f, nS), f
                if t == 2 and c.known_A < f:
                    c.dep, c.known_A = ("A", f, nA), f
                c.state = "wait"
                progressed = True
            if c.state == "wait":
                if c.dep is not None:
                    kind, df, need = c.dep
                    assert c.pending == 0, "waits while holding unsent completion signals"
                    if (done_S if kind == "S" else done_A)[df] < need:
                        continue             # spin
                # dependency complete (or known complete from an earlier item of this CTA): check the real state
                if t == 0 and f >= 3:
                    assert done_A[f - 3] == nA and reading_A[f - 3] == 0, "splat before its slot was handed back"
                if t == 1:
                    assert done_S[f] == nS, "average before the frame's splat finished"
                    reading_A[f] += 1
                if t == 2:
                    assert done_A[f] == nA, "fill-hole before the frame's masks are complete"
                c.state = "run"
                progressed = True
                continue
            if c.state == "run":
                if t >= 0:
                    ran[t][f] += 1
                if t == 1:
                    reading_A[f] -= 1
                if t in (0, 1):
                    c.pending += 1
                    nt, nf, _ = c.nx
                    if nt != t or nf != f:
                        (done_S if t == 0 else done_A)[f] += c.pending
                        c.pending = 0
                c.it, c.nx = c.nx, c.fetched
                c.state = "start"
                progressed = True
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 200, "deadlock: no CTA can make progress"
    assert ran[0] == [nS] * B and ran[1] == [nA] * B and ran[2] == [nA if fillhole else 0] * B
    assert done_S == [nS] * B and done_A == [nA] * B
    return True


def main():
    a = [int(v) for v in sys.argv[1:]]
    B, nS, nA, n_cta, seeds = (a + [16, 40, 10, 12, 20][len(a):])[:5]
    for s in range(seeds):
        simulate(B, nS, nA, n_cta, True, s)
        simulate(B, nS, nA, n_cta, False, s)
    print("fp_pipeline_model: ok (B=%d nS=%d nA=%d ctas=%d, %d seeds x 2)" % (B, nS, nA, n_cta, seeds))


if __name__ == "__main__":
    main()

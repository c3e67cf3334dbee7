"""Model check of the FlowProjection pipeline's scheduling protocol (tools/fp_pipeline_model.py mirrors the queue
order, look-ahead, completion counters, batched signals and dependency waits of fp_pipeline_kernel): no deadlock,
no item before its dependency, no accumulator slot rewritten while it is still read -- under random CTA speeds,
with more and with fewer CTAs than items per cycle, with and without fill-hole.  CPU only."""
import pytest

from tools.fp_pipeline_model import decode, simulate


@pytest.mark.parametrize("B,nS,nA,n_cta", [(3, 7, 3, 50), (4, 12, 5, 3), (16, 40, 10, 12), (9, 33, 8, 64), (5, 2040, 510, 592)])
@pytest.mark.parametrize("fillhole", [True, False])
def test_pipeline_protocol(B, nS, nA, n_cta, fillhole):
    for seed in range(2 if nS > 1000 else 8):
        assert simulate(B, nS, nA, n_cta, fillhole, seed)


def test_queue_order_puts_dependencies_a_cycle_back():
    """Every dependency of an item sits at least a full cycle minus the cycle's own average / fill items earlier."""
    B, nS, nA = 6, 10, 4
    per = 2 * nA + nS
    total = (B + 3) * per
    first = {}
    last = {}
    for q in range(total):
        t, f, _ = decode(q, B, nS, nA, True, total)
        if t >= 0:
            first.setdefault((t, f), q)
            last[(t, f)] = q
    for f in range(B):
        assert last[(0, f)] < first[(1, f)] and first[(1, f)] - last[(0, f)] > nS          # average(f) a cycle after splat(f)
        assert last[(1, f)] < first[(2, f)]                                               # fill(f) after average(f)
        if f >= 3:
            assert last[(1, f - 3)] < first[(0, f)]                                       # splat(f) after average(f-3)

"""The example scripts run end to end on the device and learn (loss goes down)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_gcn_full_batch_example():
    import gcn_full_batch
    first, last = gcn_full_batch.main(n=1500, epochs=15)
    assert last < first


def test_sampled_graphsage_example():
    import sampled_graphsage
    losses = sampled_graphsage.main(n=20000, deg=20, batches=8)
    assert len(losses) == 8 and all(l == l for l in losses) and min(losses[-3:]) < losses[0]

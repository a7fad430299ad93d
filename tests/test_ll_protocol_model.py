"""CPU model check of the flag-in-data exchange protocol of the experimental decode megakernel (csrc/llama_mega_ll.cuh): random and heavily
skewed interleavings of element accesses of G simulated CTAs never deadlock and every gather sees the version it was meant to see; a
deliberately broken protocol (WO waits for one head only) IS caught.  See tools/ll_protocol_sim.py."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))


def test_ll_protocol_has_no_deadlock_and_no_stale_reads():
    import ll_protocol_sim as sim
    for s in range(6):
        sim.run(s)
        sim.run(1000 + s, G=5, H=5, L=1, E=20, FF=12, launches=2)
        sim.run(2000 + s, G=9, H=2, L=3, E=16, FF=48, launches=2)
    for s in range(3):
        sim.run(3000 + s, skew=True)


def test_ll_model_check_catches_a_broken_protocol():
    import ll_protocol_sim as sim
    caught = 0
    for s in range(6):
        try:
            sim.run(5000 + s, bug="wo_partial_gather", skew=True)
        except AssertionError:
            caught += 1
    assert caught > 0

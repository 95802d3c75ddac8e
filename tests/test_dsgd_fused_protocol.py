"""CPU: the flag protocol of the experimental fused DSGD epoch kernel (librec_b200/csrc/dsgd_fused.cuh), restated in
librec_b200/dsgd_plan.py::FusedRank, under random interleavings of the ranks: no deadlock, every rank trains block
(rank + stratum) mod G in every epoch, nobody pushes into a buffer its owner still reads, and every rank ends an epoch holding
its starting block.  (This checks the sequencing; the memory-ordering side -- ld.acquire.sys / st.release.sys, fences around
grid.sync -- can only be checked on the GPUs.)"""
import numpy as np
import pytest

from librec_b200 import dsgd_plan as plan


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_fused_protocol_under_random_interleavings(world):
    rng = np.random.default_rng(world)
    for trial in range(30):
        ranks = [plan.FusedRank(r, world, start_block=r) for r in range(world)]
        for epoch in range(3):
            for r in ranks:
                r.begin_epoch()
            stalled = 0
            # adversarial scheduler: pick a random rank; now and then let one rank run far ahead or starve one
            while not all(r.finished() for r in ranks):
                if rng.random() < 0.1:
                    cand = [ranks[int(rng.integers(world))]] * 20
                else:
                    cand = [ranks[int(rng.integers(world))]]
                progressed = False
                for r in cand:
                    if not r.finished() and r.step(ranks) is not None:
                        progressed = True
                stalled = 0 if progressed else stalled + 1
                if stalled > 2000:
                    # nobody picked could move: make sure SOMEBODY can (otherwise it is a deadlock)
                    assert any((not r.finished()) and _can_move(r, ranks) for r in ranks), "deadlock"
                    stalled = 0
            for r in ranks:
                mine = [blk for seq, blk in r.trained if epoch * world <= seq < (epoch + 1) * world]
                assert mine == [plan.block_at(r.rank, world, s) for s in range(world)], (r.rank, mine)
                held = r.buf[r.cur0]
                assert held == r.rank, (r.rank, held)             # back to the starting block, in the buffer the next epoch reads


def _can_move(r, ranks):
    import copy
    snap = copy.deepcopy(ranks)
    return snap[r.rank].step(snap) is not None


def test_fused_protocol_matches_the_ring_schedule():
    """in lock-step the fused protocol visits the same (sub-epoch, rank) -> block table as the NCCL ring of dsgd.cuh"""
    world = 4
    ranks = [plan.FusedRank(r, world, start_block=r) for r in range(world)]
    for r in ranks:
        r.begin_epoch()
    while not all(r.finished() for r in ranks):
        assert any(r.step(ranks) is not None for r in ranks if not r.finished())
    for r in ranks:
        assert [b for _, b in r.trained] == [plan.block_at(r.rank, world, s) for s in range(world)]
        assert plan.send_peer(r.rank, world) == (r.rank - 1) % world and plan.recv_peer(r.rank, world) == (r.rank + 1) % world

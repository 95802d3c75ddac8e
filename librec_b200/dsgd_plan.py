"""The DSGD stratum schedule (host logic) -- the same pure functions as librec_b200/csrc/dsgd.cuh.

World of G ranks, one per GPU.  Rank g owns a user shard for the whole run.  Items are cut into G
contiguous blocks balanced by global rating count.  In sub-epoch s rank g trains on item block
(g + s) mod G, then passes that block to rank g-1 and receives the next one from rank g+1.  After G
sub-epochs every rank holds its starting block again.  Blocks of one sub-epoch are disjoint in users
and in items, so a DSGD epoch equals a sequential walk over its strata in (sub-epoch, rank) order.
"""
import numpy as np


def block_at(rank, world, sub):
    return (rank + sub) % world


def send_peer(rank, world):
    return (rank - 1 + world) % world


def recv_peer(rank, world):
    return (rank + 1) % world


def item_bounds(item_count, world):
    """contiguous item blocks balanced by rating count: block b = [bounds[b], bounds[b+1])"""
    item_count = np.asarray(item_count, np.int64)
    I = item_count.shape[0]
    total = int(item_count.sum())
    bounds = [I] * (world + 1)
    bounds[0] = 0
    acc, b = 0, 1
    for i in range(I):
        if b >= world:
            break
        acc += int(item_count[i])
        while b < world and (acc * world >= total * b or I - (i + 1) <= world - b):
            bounds[b] = i + 1
            b += 1
    for j in range(1, world + 1):
        if bounds[j] < bounds[j - 1]:
            bounds[j] = bounds[j - 1]
    bounds[world] = I
    return bounds


def segments(col, bounds):
    """per item block: indices of the CSR entries whose column falls in it (CSR order kept)"""
    col = np.asarray(col)
    blk = np.searchsorted(np.asarray(bounds[1:-1]), col, side="right")
    return [np.nonzero(blk == b)[0] for b in range(len(bounds) - 1)]


# ---- flag protocol of the experimental fused epoch kernel (csrc/dsgd_fused.cuh), restated as a step function so that the
# sequencing (who may read / overwrite which buffer when) can be exercised on the CPU under arbitrary interleavings.
class FusedRank:
    """One rank's view: two block buffers, ready[2] (written by rank+1), peer_free[2] (written by rank-1).  `step()` advances
    the rank by one protocol action if its wait condition holds and returns what it did (None = blocked)."""

    def __init__(self, rank, world, start_block):
        self.rank, self.world = rank, world
        self.buf = [start_block, None]          # block id held in each buffer (None = stale / free)
        self.ready = [0, 0]
        self.peer_free = [0, 0]
        self.cur0, self.seq0 = 0, 0
        self.t, self.phase = 0, "wait_ready"    # phases per stratum: wait_ready -> compute -> wait_free -> push
        self.trained = []                       # (seq, block) in the order the rank trained them
        self.final_wait_done = False

    def begin_epoch(self):
        self.t, self.phase, self.final_wait_done = 0, "wait_ready", False

    def finished(self):
        return self.t == self.world and self.final_wait_done

    def step(self, ranks):
        G = self.world
        prev, nxt = ranks[(self.rank - 1) % G], ranks[(self.rank + 1) % G]
        if self.t == G:                                        # the block of the next epoch's first stratum must have arrived
            b = (self.cur0 + G) & 1
            if self.ready[b] >= self.seq0 + G:
                self.final_wait_done = True
                self.cur0, self.seq0 = b, self.seq0 + G
                return "final"
            return None
        b, seq = (self.cur0 + self.t) & 1, self.seq0 + self.t
        if self.phase == "wait_ready":
            if self.ready[b] < seq:
                return None
            self.phase = "compute"
            return "ready"
        if self.phase == "compute":
            assert self.buf[b] is not None, "trained on a buffer that was never filled"
            self.trained.append((seq, self.buf[b]))
            self.phase = "wait_free"
            return "compute"
        if self.phase == "wait_free":
            if self.peer_free[b ^ 1] < seq:
                return None
            self.phase = "push"
            return "free"
        # push: copy my buffer b into prev's buffer b^1, then signal
        assert prev.phase_reads() != (b ^ 1), "pushed into a buffer its owner is still reading"
        prev.buf[b ^ 1] = self.buf[b]
        prev.ready[b ^ 1] = seq + 1
        nxt.peer_free[b] = seq + 1                             # my buffer b may be overwritten by rank+1 again
        self.buf[b] = None
        self.t += 1
        self.phase = "wait_ready"
        return "push"

    def phase_reads(self):
        """index of the buffer this rank is reading right now (computing on it or pushing from it), else -1"""
        if self.t < self.world and self.phase in ("compute", "wait_free", "push"):
            return (self.cur0 + self.t) & 1
        return -1

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

"""torchrun script: wall time of every C-ABI call of one trainModel() under DSGD (config C4 shard), repeated 3 times.
LRK_DSGD_FUSED=0/1 selects the ring.  Prints rank 0's table."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from librec_b200 import capi, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = synth.make_ratings("netflix")
U, I, k = d["U"], d["I"], 128
lo, hi = rank * U // world, (rank + 1) * U // world
a, b = int(d["rowptr"][lo]), int(d["rowptr"][hi])
rowptr = np.ascontiguousarray(d["rowptr"][lo:hi + 1] - a)
col, val = np.ascontiguousarray(d["col"][a:b]), np.ascontiguousarray(d["val"][a:b])
P0, Q0, _, _ = synth.init_factors(U, I, k, 11, False)
h = capi.Handle(capi.MODEL_PMF, k, device=local, seed=1)
uid = [capi.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
h.comm_init(rank, world, uid[0])
if "--stream0" in sys.argv:
    h.set_stream(torch.cuda.current_stream().cuda_stream)      # the legacy default stream, as bench.py r01 did
if "--torch-stream" in sys.argv:
    ts = torch.cuda.Stream()
    h.set_stream(ts.cuda_stream)
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t = [time.perf_counter()]
    h.set_train_csr(hi - lo, I, rowptr, col, val); t.append(time.perf_counter())
    h.set_factors(P0[lo:hi], Q0); t.append(time.perf_counter())
    ep = []
    for it in range(10):
        h.sgd_epoch(0.01, 0.08, 0.08, 0.0, it + 1); ep.append(time.perf_counter())
    t.append(ep[-1])
    h.get_factors(); t.append(time.perf_counter())
    if rank == 0:
        dt = np.diff(t) * 1e3
        eps = np.diff([t[2]] + ep) * 1e3
        print("rep %d: set_train_csr %.1f  set_factors %.1f  epochs %.1f (%s)  get_factors %.1f ms" % (
            rep, dt[0], dt[1], dt[2], " ".join("%.1f" % x for x in eps), dt[3]), flush=True)
h.close()
dist.barrier(); dist.destroy_process_group()

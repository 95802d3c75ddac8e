#!/usr/bin/env python
"""The plugin path on N GPUs: ONE process, ONE multi-device handle (lrk_create_multi = rec.cuda.devices in the Java shim) on BASELINE
configs[3] (PMF k=128, Netflix shape).  Prints one JSON line: device-resident epochs and one trainModel() through the C ABI from
host buffers (stage the full CSR + factors, E epochs, factors back).  Run on a box with N GPUs:  python tests/multi_handle_bench.py --gpus 8"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    from librec_b200 import capi, synth
    d = synth.make_ratings("netflix")
    U, I, nnz = d["U"], d["I"], int(d["rowptr"][-1])
    k, lr, reg = 128, 0.01, 0.08
    P0, Q0, _, _ = synth.init_factors(U, I, k, 11, False)
    out = {"workload": "PMF k=128, synthetic Netflix shape (%d x %d, %d ratings), one multi-device handle over %d GPUs" % (U, I, nnz, a.gpus)}
    with capi.Handle(capi.MODEL_PMF, k, seed=1, devices=list(range(a.gpus))) as h:
        t0 = time.perf_counter()
        h.set_train_csr(U, I, d["rowptr"], d["col"], d["val"])
        t1 = time.perf_counter()
        h.set_factors(P0, Q0)
        t2 = time.perf_counter()
        losses = list(h.sgd_epochs(a.warmup, lr, reg, reg, 0.0, 1))
        t3 = time.perf_counter()
        losses += list(h.sgd_epochs(a.steps, lr, reg, reg, 0.0, a.warmup + 1))
        t4 = time.perf_counter()
        gP, gQ, _, _ = h.get_factors()
        t5 = time.perf_counter()
        guard = h.sgd_safeguard()
        out.update({"metric": "MF SGD rating-updates/s", "value": nnz * a.steps / (t4 - t3), "unit": "updates/s", "n_gpus": a.gpus,
                    "ms_per_epoch_wall": (t4 - t3) / a.steps * 1e3, "device_ms_per_epoch": h.last_epoch_ms(),
                    "losses": losses, "rollbacks": guard["rollbacks"],
                    "trainModel_ms": {"set_train_csr": (t1 - t0) * 1e3, "set_factors": (t2 - t1) * 1e3, "get_factors": (t5 - t4) * 1e3},
                    "e2e_updates_per_s_10_epochs": nnz * 10 / ((t1 - t0) + (t2 - t1) + 10 * (t4 - t3) / a.steps + (t5 - t4)),
                    "finite": bool(np.isfinite(gP).all() and np.isfinite(gQ).all())})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

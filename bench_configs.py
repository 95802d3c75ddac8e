#!/usr/bin/env python
"""Secondary measurements on one B200 for the BASELINE.json configs that bench.py does not headline:
  C3: BPR k=128 SGD epoch (samples/s) on the synthetic ML-20M shape (implicit feedback, device-side sampling)
  C4: PMF k=128 SGD epoch (updates/s) on the synthetic Netflix shape (480 189 x 17 770, 100 480 507 ratings), one GPU's view
One JSON line per config: device-resident epochs timed with CUDA events (lrk_last_epoch_ms), L2 flushed between
epochs, algorithmic-byte roofline per SURVEY.md 8(d)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def run(name, model_name, shape, k, lr, reg, steps, warmup):
    import torch
    from librec_b200 import capi, synth
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    bpr = model_name == "bpr"
    d = synth.make_ratings(shape, binary=bpr)
    U, I, nnz = d["U"], d["I"], int(d["rowptr"][-1])
    P, Q, _, _ = synth.init_factors(U, I, k, 11, False)
    model = capi.MODEL_BPR if bpr else capi.MODEL_PMF
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with capi.Handle(model, k, seed=1) as h:
        h.set_train_csr(U, I, d["rowptr"], d["col"], d["val"])
        h.set_factors(P, Q)
        ms, losses = [], []
        for s in range(warmup + steps):
            flush.zero_(); torch.cuda.synchronize()
            losses.append(h.sgd_epoch(lr, reg, reg, 0.0, s + 1))
            if s >= warmup:
                ms.append(h.last_epoch_ms())
        guard = h.sgd_safeguard()
    kms = float(np.mean(ms))
    bytes_per = 6 * k * 4 if bpr else 12 + 4 * k * 4
    achieved = bytes_per * nnz / (kms * 1e-3) / 1e9
    print(json.dumps({"config": name, "metric": "BPR samples/s" if bpr else "MF SGD rating-updates/s", "value": nnz / (kms * 1e-3),
                      "unit": "samples/s" if bpr else "updates/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": kms,
                      "workload": "%s k=%d, synthetic %s shape (%d x %d, %d ratings), lr %g reg %g" % (model_name, k, shape, U, I, nnz, lr, reg),
                      "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                                   "algorithmic_bytes_per_unit": bytes_per},
                      "loss_first_last": [losses[0], losses[-1]], "losses": losses, "safeguard": guard}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.only in ("", "c3"):
        run("C3", "bpr", "ml-20m", 128, 0.01, 0.01, a.steps, a.warmup)          # bpr-test.properties
    if a.only in ("", "c4"):
        run("C4", "pmf", "netflix", 128, 0.01, 0.08, a.steps, a.warmup)         # pmf-test.properties

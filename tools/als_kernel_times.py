#!/usr/bin/env python
"""Per-kernel times of one WRMF / eALS iteration on the ML-20M shape.  Run under
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:als_ --csv --log-file gpurun_out/als_kernels.csv python tools/als_kernel_times.py
(the values under ncu are serialised and cold-cache; they show where an iteration's time goes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librec_b200 import capi, synth  # noqa: E402


def main():
    d = synth.make_ratings("ml-20m")
    U, I = d["U"], d["I"]
    val = 1.0 + d["val"]
    for model, k in ((capi.MODEL_WRMF, 20), (capi.MODEL_WRMF, 64), (capi.MODEL_EALS, 32)):
        P, Q, _, _ = synth.init_factors(U, I, k, 11, False)
        with capi.Handle(model, k, seed=1) as h:
            h.set_train_csr(U, I, d["rowptr"], d["col"], val)
            if model == capi.MODEL_EALS:
                h.set_matrix("eals.confidences", np.ones(I))
                P = np.zeros_like(P)
            h.set_factors(P, Q)
            h.sgd_epoch(0.0, 0.01, 0.01, 0.0, 1)
            print(model, k, h.last_epoch_ms(), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors of the reference's createPR(matching='multi') (lens/src/metrics.py:21-139).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_pr_multi.py

Imports the reference's own lens/src/metrics.py (matplotlib replaced by an inert fake, as in make_golden.py) and
calls createPR(..., matching='multi', n_thresh=...) on the sequence-matched matrices and ground truths of the
two bundled runs (tests/golden/config1.npz, brisevent.npz: D and GTtol produced by the reference itself), on a
soft-ground-truth variant and on a small random matrix with ties.  Output: tests/golden/pr_multi.npz.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

mg.install_fakes()
spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/lens/src/metrics.py")
rm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rm)

out = {}
for name in ("config1", "brisevent"):
    g = np.load(os.path.join(HERE, name + ".npz"))
    S, GT = g["D"].T, g["GTtol"].T                   # the orientation evaluate() passes (run_model.py:321)
    for n in (100, 7):
        P, R = rm.createPR(S, GT, None, matching="multi", n_thresh=n)
        out[f"{name}/n{n}/P"], out[f"{name}/n{n}/R"] = np.array(P, np.float64), np.array(R, np.float64)
    soft = np.roll(GT, 1, axis=1) | GT
    P, R = rm.createPR(S, GT, None, GTsoft=soft, matching="multi", n_thresh=50)
    out[f"{name}/soft/P"], out[f"{name}/soft/R"] = np.array(P, np.float64), np.array(R, np.float64)
rng = np.random.default_rng(7)
S = (rng.integers(0, 9, (37, 53)) / 4).astype(np.float32)
GT = (rng.random((37, 53)) < 0.07).astype(np.uint8)
P, R = rm.createPR(S, GT, None, matching="multi", n_thresh=20)
out["random/S"], out["random/GT"] = S, GT
out["random/P"], out["random/R"] = np.array(P, np.float64), np.array(R, np.float64)
np.savez_compressed(os.path.join(HERE, "pr_multi.npz"), **out)
print({k: v.shape for k, v in out.items()})

#!/usr/bin/env python
"""Tiny driver for ncu captures: builds the pipeline of one configuration and runs a few hot-path steps.

    ncu --set full --clock-control none --import-source on -k regex:output_tc -s 1 -c 1 \
        -o gpurun_out/prof_tc python tools/profile_step.py --steps 2 --streams 11234 --queries 10 --places 10000 --seq-len 10
    ncu ... -k regex:bin_pow2 python tools/profile_step.py --events 1000000000
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lens_b200 import synth  # noqa: E402
from lens_b200.pipeline import InferencePipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--streams", type=int, default=1000)
ap.add_argument("--queries", type=int, default=16)
ap.add_argument("--places", type=int, default=1000)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--seq-len", type=int, default=2)
ap.add_argument("--events", type=int, default=0, help="profile the event binning (K1) on this many events instead")
a = ap.parse_args()
if a.events:
    from lens_b200 import ops  # noqa: E402
    t, x, y, n_win = synth.events_device(a.events, sensor=128, seed=5, device="cuda:0")
    for _ in range(a.steps):
        f, p, c = ops.bin_events(t, x, y, 0, 250_000, n_win, 128, 8)
    torch.cuda.synchronize()
    print("binned", int(c.sum().item()), "events into", n_win, "windows")
    sys.exit(0)
Wf, Wo = synth.weights(100, 200, a.places, seed=1)
pipe = InferencePipeline(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, T=250, L=a.seq_len,
                         max_streams=a.streams, device="cuda:0", mode=a.mode)
frames = synth.frames_device(a.streams, a.queries, 80, seed=2, device="cuda:0")
gt = torch.from_numpy(synth.gt_centers(a.streams, a.queries - a.seq_len + 1, a.places - a.seq_len + 1)).cuda()
for _ in range(a.steps):
    out = pipe.step(frames=frames, gt_center=gt, gt_tol=2)
torch.cuda.synchronize()
print("hits", out["hits"].tolist(), "valid", int(out["n_valid"].item()))

#!/usr/bin/env python
"""Golden fixtures for the event-ingestion row, produced by the REAL reference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_events.py

What runs for real (imported unmodified from /root/reference):
  lens/tools/dvstools.py        FrameRep.event_preparation / event_data (text+zip path, tool='simple_rep',
                                both --reference and non-reference runs) / save_frame (cv2 PNGs)
  lens/tools/create_data_csv.py create_csv_from_images
Stubbed: `rosbag` (only ExtractRosbag uses it), matplotlib, pynmea2 (GPS reader, unused here).

Output (committed): tests/golden/events_simple_rep.npz with the synthetic event stream (as text lines, so
the float parsing is part of what is pinned), the hot pixels, the reference's random patch layout
(unique_indices / centroid dictionary) and every frame the reference wrote, for several argument sets.
"""
import argparse
import json
import os
import sys
import tempfile
import types
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def fake(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda k: (_ for _ in ()).throw(AttributeError(k)) if k.startswith("__") else (lambda *a, **k: None)
    sys.modules[name] = m
    return m


def synth_events(seed, n, W, H, t0, rate_hz, gap_every=0):
    rng = np.random.default_rng(seed)
    dt = rng.exponential(1.0 / rate_hz, size=n)
    if gap_every:
        dt[gap_every::gap_every] += rng.uniform(0.15, 0.45, size=len(dt[gap_every::gap_every]))   # silent gaps
    t = t0 + np.cumsum(dt)
    x = rng.integers(0, W, size=n)
    y = rng.integers(0, H, size=n)
    p = rng.integers(0, 2, size=n)
    lines = ["{:.12f} {} {} {}".format(tt, xx, yy, pp) for tt, xx, yy, pp in zip(t, x, y, p)]
    return lines


def run_case(FrameRep, workdir, name, lines, W, H, hot, args_over, coord_seed):
    folder = os.path.join(workdir, name)
    os.makedirs(folder, exist_ok=True)
    txt = "{} {}\n".format(W, H) + "\n".join(lines) + "\n"
    with zipfile.ZipFile(os.path.join(folder, "syn.zip"), "w") as z:
        z.writestr("syn.txt", txt)
        z.writestr("event_sum.txt", str(len(lines)))
    if hot is not None:
        with open(os.path.join(folder, "syn_hot_pixels.txt"), "w") as f:
            for hx, hy in hot:
                f.write("{},{}\n".format(hx, hy))
    out = {}
    layout = None
    for reference in (True, False):
        args = argparse.Namespace(tool="simple_rep", input_file="syn", hot_pixels="syn_hot_pixels",
                                  output_name="frames_ref" if reference else "frames_qry", dataset_folder=folder,
                                  timebin=10.0, decay_factor=5.0, accum_factor=1.0, offset=0.0, frames_max=900,
                                  frame_limit=False, pixels=25, reference=reference)
        for k, v in args_over.items():
            setattr(args, k, v)
        if not reference:
            # the non-reference run loads 'sunset1<pixels>_coordinates.npz' (dvstools.py:263): give it the
            # layout the reference run just wrote under the name it insists on
            src = os.path.join(folder, "syn{}_coordinates.npz".format(args.pixels))
            dst = os.path.join(folder, "sunset1{}_coordinates.npz".format(args.pixels))
            with open(src, "rb") as a, open(dst, "wb") as b:
                b.write(a.read())
        np.random.seed(coord_seed)
        rep = FrameRep(args)
        for _ in rep.event_data():
            pass
        frames_dir = os.path.join(folder, args.output_name)
        import cv2
        files = sorted(f for f in os.listdir(frames_dir) if f.endswith(".png"))
        frames = np.stack([cv2.imread(os.path.join(frames_dir, f), cv2.IMREAD_UNCHANGED) for f in files]) if files \
            else np.zeros((0, 5, 5), np.uint8)
        out["frames_ref" if reference else "frames_qry"] = frames
        out["files_ref" if reference else "files_qry"] = np.array(files)
        if reference:
            with np.load(os.path.join(folder, "{}unique_indices.npz".format(args.pixels))) as d:
                uniq = d["arr_0"]
            with open(os.path.join(folder, "cooridnates_dict{}.json".format(args.pixels))) as f:
                cdict = json.load(f)
            layout = (uniq, cdict)
            out["offset_after"] = np.float64(args.offset)
    out["unique_indices"] = layout[0]
    out["dict_keys"] = np.array([int(k) for k in layout[1].keys()], dtype=np.int64)
    out["dict_vals"] = np.array([int(v) for v in layout[1].values()], dtype=np.int64)
    return out, folder


def main():
    fake("rosbag")
    fake("matplotlib"); fake("matplotlib.pyplot"); fake("matplotlib.colors")
    fake("pynmea2")
    sys.path.insert(0, REF)
    from lens.tools.dvstools import FrameRep
    from lens.tools.create_data_csv import create_csv_from_images

    W, H = 32, 24
    cases = {
        # name: (lines, hot pixels, arg overrides)
        "plain": (synth_events(1, 4000, W, H, 1587452582.35, 2500.0), None, {}),
        "hot_gaps": (synth_events(2, 5000, W, H, 12.5, 3000.0, gap_every=700),
                     [(3, 4), (10, 10), (31, 23), (0, 0)], {"timebin": 8.0}),
        "offset": (synth_events(3, 4000, W, H, 100.0, 2500.0), [(5, 5)],
                   {"offset": 100.4, "timebin": 20.0, "pixels": 16}),
        "limit": (synth_events(4, 6000, W, H, 0.5, 4000.0), None,
                  {"frame_limit": True, "frames_max": 5, "timebin": 25.0, "accum_factor": 3.0}),
    }
    golden = {}
    with tempfile.TemporaryDirectory() as work:
        for name, (lines, hot, over) in cases.items():
            out, folder = run_case(FrameRep, work, name, lines, W, H, hot, over, coord_seed=7)
            golden[name + "/lines"] = np.array(lines)
            golden[name + "/hot"] = np.array(hot if hot is not None else np.zeros((0, 2)), dtype=np.int64).reshape(-1, 2)
            golden[name + "/args"] = np.array(json.dumps(over))
            for k, v in out.items():
                golden[name + "/" + k] = v
            print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
            if name == "plain":
                csv_path = os.path.join(work, "plain.csv")
                create_csv_from_images(os.path.join(folder, "frames_ref"), csv_path)
                golden["plain/csv"] = np.array(open(csv_path).read())
    golden["sensor"] = np.array([W, H])
    np.savez_compressed(os.path.join(HERE, "events_simple_rep.npz"), **golden)
    print("wrote events_simple_rep.npz", os.path.getsize(os.path.join(HERE, "events_simple_rep.npz")))


if __name__ == "__main__":
    main()

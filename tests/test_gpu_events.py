"""GPU tests of the event-ingestion row: event-driven frames (dvstools simple_rep), the
FrameRep mirror end to end against the PNGs the reference wrote, and the events -> PNG/CSV dataset path."""
import argparse
import json
import os
import zipfile

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    g = np.load(os.path.join(GOLDEN, "events_simple_rep.npz"))
    W, H = (int(v) for v in g["sensor"])
    for name in ("plain", "hot_gaps", "offset", "limit"):
        args = dict(timebin=10.0, accum_factor=1.0, offset=0.0, frames_max=900, frame_limit=False, pixels=25)
        args.update(json.loads(str(g[name + "/args"])))
        yield name, g, W, H, [str(s) for s in g[name + "/lines"]], args, [tuple(int(v) for v in r) for r in g[name + "/hot"]]


def test_framerep_writes_the_reference_frames(tmp_path):
    """FrameRep(args).event_data() on the same zip / hot-pixel file / numpy seed -> the same PNG files."""
    import cv2
    from lens_b200.tools.dvstools import FrameRep
    for name, g, W, H, lines, over, hot in _cases():
        folder = tmp_path / name
        folder.mkdir()
        with zipfile.ZipFile(folder / "syn.zip", "w") as z:
            z.writestr("syn.txt", "{} {}\n".format(W, H) + "\n".join(lines) + "\n")
            z.writestr("event_sum.txt", str(len(lines)))
        if hot:
            (folder / "syn_hot_pixels.txt").write_text("".join("{},{}\n".format(a, b) for a, b in hot))
        for reference in (True, False):
            a = argparse.Namespace(tool="simple_rep", input_file="syn", hot_pixels="syn_hot_pixels",
                                   output_name="out_ref" if reference else "out_qry", dataset_folder=str(folder),
                                   timebin=10.0, decay_factor=5.0, accum_factor=1.0, offset=0.0, frames_max=900,
                                   frame_limit=False, pixels=25, reference=reference)
            for k, v in over.items():
                setattr(a, k, v)
            np.random.seed(7)                                # the golden run's seed for the patch layout
            rep = FrameRep(a)
            assert list(rep.event_data()) == []
            want = g[name + "/frames_ref"]
            files = sorted(os.listdir(folder / a.output_name))
            assert files == [str(s) for s in g[name + "/files_ref"]], name
            got = np.stack([cv2.imread(str(folder / a.output_name / f), cv2.IMREAD_UNCHANGED) for f in files])
            assert np.array_equal(got, want), name
            assert np.array_equal(rep.frames, want)
            assert a.offset == float(g[name + "/offset_after"])


def _random_stream(seed, n, W, H, rate, gaps, t0):
    rng = np.random.default_rng(seed)
    dt = rng.exponential(1.0 / rate, size=n)
    if gaps:
        idx = rng.choice(n, size=gaps, replace=False)
        dt[idx] += rng.uniform(0.0, 0.3, size=gaps)
    t = np.round(t0 + np.cumsum(dt), 12)
    return t, rng.integers(0, W, size=n), rng.integers(0, H, size=n)


@pytest.mark.parametrize("seed,n,W,H,fps,pixels,n_hot,offset_mode,frames_max", [
    (1, 60000, 128, 128, 30.0, 100, 0, "first", None),
    (2, 60000, 128, 128, 200.0, 64, 800, "first", None),      # many short frames, 5 % hot pixels
    (3, 40000, 346, 260, 2000.0, 400, 3000, "mid", None),     # frames of a few events, empty frames
    (4, 50000, 64, 48, 50.0, 9, 100, "before", 7),            # frame limit
    (5, 100000, 16, 16, 1.0, 4, 0, "first", None),              # > 256 events per slot: uint8 wrap
    (6, 50000, 32, 32, 100.0, 16, 1023, "first", None),        # all pixels but one are hot
])
def test_event_driven_frames_match_oracle(seed, n, W, H, fps, pixels, n_hot, offset_mode, frames_max):
    from lens_b200.tools import dvstools
    t, x, y = _random_stream(seed, n, W, H, rate=20000.0, gaps=25, t0=1000.5)
    rng = np.random.default_rng(100 + seed)
    uniq, cdict = dvstools.make_patch_layout((H, W), pixels, rng)
    hot = set()
    if n_hot:
        flat = rng.choice(W * H, size=n_hot, replace=False)
        hot = set((int(f % W), int(f // W)) for f in flat)
    offset = {"first": 0.0, "mid": float(t[n // 3]) + 1e-7, "before": float(t[0]) - 0.013}[offset_mode]
    want, want_offset = O.simple_rep(t, x, y, (H, W), uniq, cdict, hot or None, fps, offset, 1.0,
                                     frames_max if frames_max else 10 ** 9, frames_max is not None)
    lut = dvstools.layout_lut((H, W), uniq, cdict, hot)
    frames, t0s, offset_used = dvstools.events_to_slot_frames(t, x, y, lut, pixels, 1.0 / fps, offset, 1.0, frames_max)
    assert frames.shape == want.shape
    assert np.array_equal(frames.cpu().numpy(), want)
    assert offset_used == want_offset
    assert len(want) > (3 if frames_max is None else 0)
    if seed == 5:
        assert (np.diff(t0s.cpu().numpy()) > 1.0).all()          # every slot saw > 256 events per frame


def test_accum_factor_and_unsorted_input():
    from lens_b200.tools import dvstools
    t, x, y = _random_stream(9, 20000, 32, 32, rate=5000.0, gaps=0, t0=3.0)
    uniq, cdict = dvstools.make_patch_layout((32, 32), 25, np.random.default_rng(1))
    lut = dvstools.layout_lut((32, 32), uniq, cdict, None)
    for accum in (2.0, 3.7, 0.5):
        want, _ = O.simple_rep(t, x, y, (32, 32), uniq, cdict, None, 4.0, 0.0, accum)
        got, _, _ = dvstools.events_to_slot_frames(t, x, y, lut, 25, 0.25, 0.0, accum)
        assert np.array_equal(got.cpu().numpy(), want), accum
    t[100], t[101] = t[101], t[100]
    with pytest.raises(ValueError):
        dvstools.events_to_slot_frames(t, x, y, lut, 25, 0.25)


def test_events_to_dataset(tmp_path):
    """Recorded stream -> fixed timebin windows -> frame_%05d.png + CSV (collect_data.py:193-202, 252),
    empty windows dropped without advancing the counter."""
    import cv2
    from lens_b200 import collect_data
    rng = np.random.default_rng(3)
    n, roi, k = 30000, 80, 8
    t_us = np.sort(rng.integers(0, 40 * 33000, size=n)).astype(np.uint32)
    t_us[(t_us >= 5 * 33000) & (t_us < 8 * 33000)] = 5 * 33000 - 1      # three empty windows
    t_us = np.sort(t_us)
    x, y = rng.integers(0, roi, size=n), rng.integers(0, roi, size=n)
    frames, pooled, keep = collect_data.frames_from_events(t_us, x, y, 33, roi, k, t0_us=0, n_windows=40)
    want_frames, want_pooled, want_cnt = O.bin_events(t_us, x.astype(np.uint16), y.astype(np.uint16), 0, 33000, 40,
                                                      roi, k)
    nz = np.nonzero(want_cnt > 0)[0]
    assert len(nz) == 37 and np.array_equal(keep.cpu().numpy(), nz)
    assert np.array_equal(frames.cpu().numpy(), want_frames[nz])
    assert np.array_equal(pooled.cpu().numpy(), want_pooled[nz])
    n_written = collect_data.write_dataset(frames, str(tmp_path / "imgs"), str(tmp_path / "d.csv"))
    assert n_written == 37
    rows = (tmp_path / "d.csv").read_text().splitlines()
    assert rows[0] == "Image_name,index" and rows[1] == "frame_00000.png,0" and rows[-1] == "frame_00036.png,36"
    back = cv2.imread(str(tmp_path / "imgs" / "frame_00005.png"), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(back, want_frames[nz[5]])

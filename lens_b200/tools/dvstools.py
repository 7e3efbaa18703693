"""Raw DVS event ingestion ahead of the binning kernels (mirror of lens/tools/dvstools.py:109-361).

Readers: the "t x y p" text format inside `<name>.zip` (first line "width height", a second member
event_sum.txt with the event count; written by ExtractRosbag, dvstools.py:75-99) and parquet files with
columns t (microseconds), x, y, p.  `FrameRep` keeps the reference's constructor and `event_data()` entry
point for tool='simple_rep'; the per-event Python loop of dvstools.py:279-349 is replaced by
lens_event_windows + lens_bin_events_lut (include/lens_b200.h).

Not mirrored, with the reason:
  * tool='decay_rep'  - dvstools.py:302-309 indexes the 1-D `frame_data` vector with [y, x]; the reference
    raises IndexError on its first event, so there is no behaviour to reproduce.
  * the parquet branch of event_data() (dvstools.py:185-213) - it subscripts the bound method
    `self.event_data`, and event_preparation() opens the parquet file as a zip, so it cannot run either.
    `read_events_parquet` + `lens_b200.ops.bin_events` give the working equivalent (fixed time grid).
  * ExtractRosbag / CreateVideo - ROS bag and mp4 I/O, no arithmetic.
"""
import io
import json
import os
import zipfile

import numpy as np
import torch

from .. import ops

LUT_HOT, LUT_NONE = -2, -1


# ------------------------------------------------------------------------------------------ readers
def _parse_event_text(buf, has_header=True):
    import pandas as pd
    # float_precision='round_trip' = correctly rounded, like the float() the reference calls per field
    df = pd.read_csv(buf, sep=r" ", header=None, skiprows=1 if has_header else 0, engine="c",
                     float_precision="round_trip", names=["t", "x", "y", "p"], dtype=np.float64)
    t = df["t"].to_numpy(np.float64)
    # the reference does int(float(field)): truncation towards zero
    x = df["x"].to_numpy(np.float64).astype(np.int64)
    y = df["y"].to_numpy(np.float64).astype(np.int64)
    p = df["p"].to_numpy(np.float64).astype(np.int64)
    return t, x, y, p


def read_events_zip(zip_path, member=None):
    """`<name>.zip` holding `<name>.txt` ("W H" then "t x y p" lines, t in seconds) and event_sum.txt.
    -> dict(t f64, x i64, y i64, p i64, width, height, event_sum)   (dvstools.py:113-120, 162-166, 279-283)"""
    with zipfile.ZipFile(zip_path, "r") as z:
        if member is None:
            member = os.path.splitext(os.path.basename(zip_path))[0] + ".txt"
        with z.open(member) as f:
            raw = f.read()
        event_sum = None
        if "event_sum.txt" in z.namelist():
            event_sum = int(z.read("event_sum.txt").decode("utf-8").strip())
    first, _, _ = raw.partition(b"\n")
    width, height = (int(v) for v in first.split())
    t, x, y, p = _parse_event_text(io.BytesIO(raw))
    return dict(t=t, x=x, y=y, p=p, width=width, height=height, event_sum=event_sum)


def read_events_text(path, has_header=True):
    """Plain "t x y p" text file (optionally with the "W H" header line)."""
    width = height = None
    if has_header:
        with open(path, "rb") as f:
            width, height = (int(v) for v in f.readline().split())
    t, x, y, p = _parse_event_text(path, has_header)
    return dict(t=t, x=x, y=y, p=p, width=width, height=height, event_sum=len(t))


def write_events_zip(zip_path, t, x, y, p, width, height):
    """Inverse of read_events_zip, in ExtractRosbag's format (dvstools.py:29-31, 75-99)."""
    name = os.path.splitext(os.path.basename(zip_path))[0]
    lines = ["{} {}".format(width, height)]
    lines += ["{:.12f} {} {} {}".format(tt, xx, yy, pp) for tt, xx, yy, pp in zip(t, x, y, p)]
    with zipfile.ZipFile(zip_path, "w") as z:
        z.writestr(name + ".txt", "\n".join(lines) + "\n", compress_type=zipfile.ZIP_DEFLATED)
        z.writestr("event_sum.txt", str(len(t)), compress_type=zipfile.ZIP_DEFLATED)


def read_events_parquet(path):
    """Parquet with columns t (microseconds), x, y[, p] -> the same dict, t converted to seconds like
    dvstools.py:187; sensor size is the DAVIS346 default the reference assumes (dvstools.py:152)."""
    import pyarrow.parquet as pq
    tab = pq.read_table(path)
    cols = {n: tab.column(n).to_numpy() for n in tab.column_names}
    t = cols["t"].astype(np.float64) / 1000000
    p = cols["p"].astype(np.int64) if "p" in cols else np.zeros(len(t), np.int64)
    return dict(t=t, x=cols["x"].astype(np.int64), y=cols["y"].astype(np.int64), p=p, width=346, height=260,
                event_sum=len(t))


def read_hot_pixels(path):
    """Lines "x,y" -> set of (x, y)   (dvstools.py:122-128)."""
    hot = set()
    with open(path, "r") as f:
        for line in f:
            if line.strip():
                xs, ys = line.strip().split(",")
                hot.add((int(xs), int(ys)))
    return hot


# ------------------------------------------------------------------------------------- patch layout
def make_patch_layout(dimensions, pixels, rng=np.random):
    """Random centroids and their 3x3 patches (dvstools.py:221-244).
    dimensions = (height, width).  -> (unique_indices int64 [pixels], {flat pixel: flat centroid})."""
    H, W = dimensions
    unique_indices = rng.choice(H * W, size=pixels, replace=False)
    cdict = {}
    for centroid in unique_indices:
        r0, c0 = divmod(int(centroid), W)
        for r in range(r0 - 1, r0 + 2):
            for c in range(c0 - 1, c0 + 2):
                if 0 <= r < H and 0 <= c < W:
                    cdict[int(r * W + c)] = int(centroid)        # a later centroid takes over shared pixels
    return np.asarray(unique_indices), cdict


def layout_lut(dimensions, unique_indices, cdict, hot_pixels=None):
    """Pixel -> slot lookup table int16 [H, W]: slot of the owning centroid, LUT_NONE, or LUT_HOT."""
    H, W = dimensions
    if len(unique_indices) > 8192:
        raise ValueError("at most 8192 slots")
    slot_of = {int(c): i for i, c in reversed(list(enumerate(unique_indices)))}   # np.where(...)[0][0]: first
    lut = np.full(H * W, LUT_NONE, dtype=np.int16)
    for pix, cen in cdict.items():
        lut[int(pix)] = slot_of[int(cen)]
    lut = lut.reshape(H, W)
    if hot_pixels:
        for (hx, hy) in hot_pixels:
            if 0 <= hx < W and 0 <= hy < H:
                lut[hy, hx] = LUT_HOT
    return lut


# ------------------------------------------------------------------------------------- GPU pipeline
def events_to_slot_frames(t, x, y, lut, n_slots, interval, offset=0.0, accum_factor=1.0, frames_max=None,
                          device="cuda"):
    """Event arrays -> (frames u8 [n_frames, n_slots] on the device, frame start times f64 [n_frames],
    offset used).

    t float64 seconds (ascending), x / y integer pixel coordinates, lut int16 [H, W] from layout_lut.
    Frame boundaries follow dvstools.py:286-349 (lens_event_windows); the frames are what the reference
    writes as images_00000.png ... (the still-open last frame is not emitted, like the reference).
    offset == 0 means "start at the first event" (dvstools.py:286-290)."""
    n = len(t)
    if accum_factor < 0:
        raise ValueError("accum_factor must be >= 0")
    if n and (np.min(x) < 0 or np.min(y) < 0 or np.max(x) > 65535 or np.max(y) > 65535):
        raise ValueError("pixel coordinates must fit uint16")
    dev = torch.device(device)
    tt = torch.from_numpy(np.array(t, dtype=np.float64)).to(dev)
    xx = torch.as_tensor(np.ascontiguousarray(x).astype(np.uint16).view(np.int16)).to(dev)
    yy = torch.as_tensor(np.ascontiguousarray(y).astype(np.uint16).view(np.int16)).to(dev)
    lut_d = torch.from_numpy(np.array(lut, dtype=np.int16)).to(dev)
    max_windows = n if frames_max is None else min(int(frames_max), n)
    use_first = offset == 0
    wb, we, wt, n_win = ops.event_windows(tt, xx, yy, lut_d, use_first, float(offset), float(interval), max_windows)
    n_frames = int(n_win.item())
    hint = (n // n_frames) if n_frames else 0
    # int(prev + accum) on every event: a fractional accum_factor adds its integer part (dvstools.py:322)
    frames = ops.bin_events_lut(xx, yy, lut_d, wb[:n_frames], we[:n_frames], n_slots, int(accum_factor) & 255, hint)
    offset_used = float(t[0]) if (use_first and n) else float(offset)
    return frames, wt[:n_frames], offset_used


class FrameRep:
    """Same constructor / entry points as the reference's FrameRep (dvstools.py:109-361) for the zip text
    path with tool='simple_rep'.  `event_data()` is a generator (the reference's is too, and like the
    reference's text branch it yields nothing): iterate it to run the conversion.  Afterwards
    `self.frames` (u8 [n, sqrt(pixels), sqrt(pixels)]) holds what was written as images_%05d.png."""

    def __init__(self, args):
        self.args = args
        self.frames = None
        self.event_preparation()

    def read_camera_dimensions(self, file_path):
        with zipfile.ZipFile(file_path, "r") as z:
            with z.open(self.args.input_file + ".txt") as f:
                dims = tuple(map(int, f.readline().split()))
        print("Camera dimensions: {} x {}".format(dims[0], dims[1]))
        return dims

    def read_hot_pixels(self, file_path):
        return read_hot_pixels(file_path)

    def event_preparation(self):
        a = self.args
        self.file_path = os.path.join(a.dataset_folder, a.input_file + ".zip")
        self.is_parquet = False
        if os.path.exists(os.path.join(a.dataset_folder, a.input_file + ".parquet")):
            raise NotImplementedError("the reference's parquet branch cannot run (see module docstring); use "
                                      "read_events_parquet + lens_b200.ops.bin_events")
        w, h = self.read_camera_dimensions(self.file_path)
        self.dimensions = (h, w)                                  # (height, width), dvstools.py:147
        self.hot_pixels = None
        hot_path = os.path.join(a.dataset_folder, a.hot_pixels + ".txt")
        if os.path.exists(hot_path):
            self.hot_pixels = self.read_hot_pixels(hot_path)
        self.frame_folder = os.path.join(a.dataset_folder, a.input_file if a.output_name == "" else a.output_name)
        os.makedirs(self.frame_folder, exist_ok=True)
        with zipfile.ZipFile(self.file_path, "r") as z:
            self.total_frames = int(z.read("event_sum.txt").decode("utf-8").strip())
        if a.frames_max < self.total_frames and a.frame_limit:
            self.total_frames = a.frames_max

    # file names are the reference's own, typos included (dvstools.py:245-255, 262-274)
    def _layout_paths(self):
        a = self.args
        d = a.dataset_folder
        return (os.path.join(d, "cooridnates_dict{}.json".format(a.pixels)),
                os.path.join(d, "cooridnates_centroid{}.json".format(a.pixels)),
                os.path.join(d, "{}{}_coordinates.npz".format(a.input_file, a.pixels)),
                os.path.join(d, "{}unique_indices.npz".format(a.pixels)))

    def _layout(self):
        a = self.args
        p_dict, p_cent, p_coord, p_uniq = self._layout_paths()
        if a.reference:
            uniq, cdict = make_patch_layout(self.dimensions, a.pixels)
            with open(p_dict, "w") as f:
                json.dump(cdict, f)
            with open(p_cent, "w") as f:
                json.dump([int(c) for c in uniq], f)
            keys = np.array(list(cdict.keys()), dtype=np.int64)
            rows, cols = np.unravel_index(keys, self.dimensions)
            np.savez_compressed(p_coord, np.stack([cols, rows], axis=1))
            np.savez_compressed(p_uniq, uniq)
        else:
            with np.load(p_uniq) as d:
                uniq = d["arr_0"]
            with open(p_dict) as f:
                cdict = {int(k): int(v) for k, v in json.load(f).items()}
        return uniq, cdict

    def event_data(self):
        a = self.args
        if a.tool == "decay_rep":
            raise NotImplementedError("decay_rep raises IndexError in the reference (see module docstring)")
        if int(np.sqrt(a.pixels)) ** 2 != a.pixels:
            raise ValueError("pixels must be a square number (save_frame reshapes to sqrt x sqrt)")
        ev = read_events_zip(self.file_path, a.input_file + ".txt")
        uniq, cdict = self._layout()
        lut = layout_lut(self.dimensions, uniq, cdict, self.hot_pixels)
        frames, _, offset_used = events_to_slot_frames(
            ev["t"], ev["x"], ev["y"], lut, n_slots=a.pixels, interval=1.0 / a.timebin, offset=a.offset,
            accum_factor=a.accum_factor, frames_max=a.frames_max if a.frame_limit else None)
        a.offset = offset_used                                    # dvstools.py:287 writes it back too
        side = int(np.sqrt(a.pixels))
        self.frames = frames.reshape(-1, side, side).cpu().numpy()
        if a.tool != "event_profile":
            for i, fr in enumerate(self.frames):
                self.save_frame(fr, i, output_dir=self.frame_folder)
        return
        yield  # pragma: no cover  (generator, like the reference's)

    def save_frame(self, frame, frame_index, output_dir="frames"):
        os.makedirs(output_dir, exist_ok=True)
        side = int(np.sqrt(self.args.pixels))
        write_png(os.path.join(output_dir, f"images_{frame_index:05d}.png"), np.reshape(frame, (side, side)))


def write_png(path, frame_u8):
    """8-bit grayscale PNG (cv2.imwrite in dvstools.py:358, imageio.imwrite in collect_data.py:198)."""
    try:
        import cv2
        cv2.imwrite(path, np.ascontiguousarray(frame_u8, dtype=np.uint8))
    except ImportError:
        from PIL import Image
        Image.fromarray(np.ascontiguousarray(frame_u8, dtype=np.uint8)).save(path)

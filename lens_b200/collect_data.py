"""Event stream -> PNG/CSV dataset (the post-processing half of lens/collect_data.py:186-252).

The reference's LENS_Collector gathers one list of Speck events per `timebin` ms (event_collector,
collect_data.py:186-191), turns every non-empty list into `frame_%05d.png` (create_images, :193-202,
`frame[y-1, x-1] += 1`, uint8 wrap) and finishes with create_csv_from_images (:252).  Here the event
lists are the fixed `timebin` windows of a recorded stream and the frames come from lens_bin_events.
"""
import os

import numpy as np
import torch

from . import ops
from .tools.create_data_csv import create_csv_from_images
from .tools.dvstools import write_png


def frames_from_events(t_us, x, y, timebin_ms, roi_dim, kernel_size, t0_us=None, n_windows=None, roi_x0=0, roi_y0=0,
                       device="cuda"):
    """-> (frames u8 [n_kept, roi, roi], pooled u8 [n_kept, dims^2], kept window indices) on the device.
    Windows without events are dropped and do not advance the frame counter (collect_data.py:194-202)."""
    t_us = np.ascontiguousarray(t_us, dtype=np.uint32)
    window_us = int(timebin_ms) * 1000
    if t0_us is None:
        t0_us = int(t_us[0]) if len(t_us) else 0
    if n_windows is None:
        n_windows = (int(t_us[-1]) - t0_us) // window_us + 1 if len(t_us) else 0
    dev = torch.device(device)
    td = torch.as_tensor(t_us.view(np.int32)).to(dev)
    xd = torch.as_tensor(np.ascontiguousarray(x).astype(np.uint16).view(np.int16)).to(dev)
    yd = torch.as_tensor(np.ascontiguousarray(y).astype(np.uint16).view(np.int16)).to(dev)
    frames, pooled, n_events = ops.bin_events(td, xd, yd, t0_us, window_us, n_windows, roi_dim, kernel_size,
                                              roi_x0=roi_x0, roi_y0=roi_y0)
    keep = torch.nonzero(n_events > 0).flatten()
    return frames[keep], pooled[keep], keep


def write_dataset(frames, img_folder, csv_path):
    """frame_%05d.png per frame + the dataset CSV (collect_data.py:198, 252)."""
    os.makedirs(img_folder, exist_ok=True)
    host = frames.cpu().numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
    for i, fr in enumerate(host):
        write_png(os.path.join(img_folder, f"frame_{i:05d}.png"), fr)
    create_csv_from_images(img_folder, csv_path)
    return len(host)

"""Seeded synthetic workloads of the shapes BASELINE.json names.

No datasets or checkpoints are reachable from the GPU box, so benchmarks and scale tests use
random weights with the statistics of the trained LENS models and random count frames / event
streams with the statistics of the bundled Speck recordings.  Everything is generated on the
host with numpy (deterministic per seed) and moved to the device by the caller.
"""
import numpy as np


def weights(I, F, P, seed=1):
    """W_feat [F, I], W_out [P, F] f32 mimicking trained statistics.

    W_feat: ~30 % positive (mean +0.147), ~53 % negative (mean -0.089), ~17 % zero, clipped to
    [-4.7, 1.05]; W_out: dense N(0, 0.0143) clipped to [-0.11, 0.055] with |w| >= 1e-6 (the trainer's
    clamp, lens/src/blitnet.py:234-235 of the reference).
    """
    rng = np.random.default_rng(seed)
    kind = rng.random((F, I))
    Wf = np.where(kind < 0.30, rng.exponential(0.147, (F, I)),
                  np.where(kind < 0.83, -rng.exponential(0.089, (F, I)), 0.0))
    Wf = np.clip(Wf, -4.7, 1.05).astype(np.float32)
    Wo = np.clip(rng.normal(0.0, 0.0143, (P, F)), -0.11, 0.055)
    Wo = np.where(np.abs(Wo) < 1e-6, 1e-6, Wo).astype(np.float32)
    return Wf, Wo


def pixel_counts(shape, seed=2):
    """u8 event counts per pixel: geometric with mean ~8, ~43 % zeros, wrapped modulo 256."""
    rng = np.random.default_rng(seed)
    v = rng.geometric(1.0 / 15.0, shape) - 1
    v = np.where(rng.random(shape) < 0.40, 0, v)
    return (v % 256).astype(np.uint8)


def frames(B, Q, roi=80, seed=2):
    """Count frames u8 [B, Q, roi, roi] of 'Speck resolution' (80x80 ROI of the 128x128 sensor)."""
    return pixel_counts((B, Q, roi, roi), seed)


def events(n_events, sensor=128, window_us=250_000, events_per_window=131_072, seed=5,
           hot_pixels=16, hot_rate=100.0):
    """SoA DVS stream (t_us u32 sorted, x u16, y u16) with hot pixels.

    Uniform inter-arrival times (mean window_us / events_per_window); x, y uniform over the sensor
    except `hot_pixels` pixels that fire `hot_rate` times more often than the rest.
    """
    rng = np.random.default_rng(seed)
    n_win = int(np.ceil(n_events / events_per_window))
    t = np.sort(rng.integers(0, n_win * window_us, n_events, dtype=np.int64)).astype(np.uint32)
    x = rng.integers(0, sensor, n_events).astype(np.uint16)
    y = rng.integers(0, sensor, n_events).astype(np.uint16)
    if hot_pixels:
        p_hot = hot_pixels * hot_rate / (sensor * sensor + hot_pixels * (hot_rate - 1))
        m = rng.random(n_events) < p_hot
        hx = rng.integers(0, sensor, hot_pixels).astype(np.uint16)
        hy = rng.integers(0, sensor, hot_pixels).astype(np.uint16)
        pick = rng.integers(0, hot_pixels, int(m.sum()))
        x[m] = hx[pick]
        y[m] = hy[pick]
    return t, x, y, n_win


def gt_centers(B, Qo, Po, seed=7):
    """Synthetic ground truth: query q of stream b shows place (offset_b + q) (i32 [B, Qo])."""
    rng = np.random.default_rng(seed)
    off = rng.integers(0, max(1, Po - Qo + 1), B)
    return (off[:, None] + np.arange(Qo)[None, :]).astype(np.int32)


# ---- device-side generators for the full-size configurations (host numpy would take minutes) ----------
def frames_device(B, Q, roi=80, seed=2, device="cuda", chunk_streams=2048):
    """Count frames u8 [B, Q, roi, roi] with the statistics of `pixel_counts`, drawn on the device
    (torch CUDA generator, deterministic per seed and device type) in chunks of streams."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    out = torch.empty((B, Q, roi, roi), dtype=torch.uint8, device=device)
    for b0 in range(0, B, chunk_streams):
        n = min(chunk_streams, B - b0)
        shape = (n, Q, roi, roi)
        v = torch.empty(shape, dtype=torch.float32, device=device).geometric_(1.0 / 15.0, generator=g) - 1.0
        zero = torch.rand(shape, device=device, generator=g) < 0.40
        v = torch.where(zero, torch.zeros_like(v), v)
        out[b0:b0 + n] = v.to(torch.int32).remainder_(256).to(torch.uint8)
    return out


def events_device(n_events, sensor=128, window_us=250_000, events_per_window=131_072, seed=5,
                  hot_pixels=16, hot_rate=100.0, device="cuda", chunk_windows=512):
    """The event stream of `events` drawn on the device: (t_us i32 sorted, x i16, y i16, n_windows).

    Timestamps are uniform inside consecutive spans of `chunk_windows` windows (sorted per span), so
    every window holds events_per_window events on average; x, y uniform with `hot_pixels` pixels
    firing `hot_rate` times more often."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    n_win = -(-n_events // events_per_window)
    t = torch.empty(n_events, dtype=torch.int32, device=device)
    per_chunk = chunk_windows * events_per_window
    for c0 in range(0, n_events, per_chunk):
        n = min(per_chunk, n_events - c0)
        wins = -(-n // events_per_window)
        base = (c0 // events_per_window) * window_us
        tt = torch.randint(0, wins * window_us, (n,), dtype=torch.int32, device=device, generator=g)
        t[c0:c0 + n] = torch.sort(tt).values + base
    assert n_win * window_us < 2 ** 31
    x = torch.randint(0, sensor, (n_events,), dtype=torch.int16, device=device, generator=g)
    y = torch.randint(0, sensor, (n_events,), dtype=torch.int16, device=device, generator=g)
    if hot_pixels:
        p_hot = hot_pixels * hot_rate / (sensor * sensor + hot_pixels * (hot_rate - 1))
        hx = torch.randint(0, sensor, (hot_pixels,), dtype=torch.int16, device=device, generator=g)
        hy = torch.randint(0, sensor, (hot_pixels,), dtype=torch.int16, device=device, generator=g)
        step = 1 << 26
        for c0 in range(0, n_events, step):
            n = min(step, n_events - c0)
            m = torch.rand(n, device=device, generator=g) < p_hot
            pick = torch.randint(0, hot_pixels, (n,), device=device, generator=g)
            x[c0:c0 + n] = torch.where(m, hx[pick], x[c0:c0 + n])
            y[c0:c0 + n] = torch.where(m, hy[pick], y[c0:c0 + n])
    return t, x, y, int(n_win)

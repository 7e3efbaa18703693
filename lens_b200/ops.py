"""Host-side wrappers of the C ABI (torch tensors in, torch tensors out).

PyTorch is used for device memory and streams only; every computation happens in
liblens_b200.so (hand-written sm_100a kernels).  All wrappers are asynchronous on
torch's current CUDA stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr, require_cuda

RECALL_NS = (1, 5, 10, 15, 20, 25)   # lens/run_model.py:266


def pool_geometry(roi, k):
    """dims and centre index of the one-hot strided conv (lens/run_model.py:130-137)."""
    dims = roi // k
    c = (k // 2) - 1
    if c < 0:
        c += k
    return dims, c


def pool_frames(frames, k):
    """frames u8 [n, roi, roi] -> pooled u8 [n, dims*dims] (pixel pick of run_model.py:130-137)."""
    require_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() == 3 and frames.shape[1] == frames.shape[2]
    n, roi, _ = frames.shape
    dims, _ = pool_geometry(roi, k)
    out = torch.empty((n, dims * dims), dtype=torch.uint8, device=frames.device)
    check(_lib.lib().lens_pool_frames(ptr(frames), n, roi, k, ptr(out), stream_ptr()), "lens_pool_frames")
    return out


def bin_events(t_us, x, y, t0_us, window_us, n_win, roi, k, roi_x0=0, roi_y0=0, index_shift=1,
               wrap_u8=True, want_frames=True, want_pooled=True, check_sorted=False):
    """Event stream -> (frames u8 [n_win, roi, roi], pooled u8 [n_win, I], events-per-window i32).

    Mirrors lens/collect_data.py:186-202 (`frame[y-1, x-1] += 1`, `.astype(np.uint8)`); the
    caller drops windows whose count is 0 to reproduce create_images' "No events" branch.
    t_us must be ascending (window ranges are found by binary search); check_sorted=True verifies it with
    one extra pass over t_us (lens_check_sorted_u32, synchronises) and raises ValueError otherwise.
    """
    require_cuda(t_us, x, y)
    assert t_us.dtype == torch.int32 or t_us.dtype == torch.uint32
    assert x.dtype in (torch.int16, torch.uint16) and y.dtype in (torch.int16, torch.uint16)
    n = t_us.numel()
    dev = t_us.device
    if check_sorted:
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        check(_lib.lib().lens_check_sorted_u32(ptr(t_us), n, ptr(flag), stream_ptr()), "lens_check_sorted_u32")
        if int(flag.item()):
            raise ValueError("event timestamps must be ascending")
    dims, _ = pool_geometry(roi, k)
    frames = torch.empty((n_win, roi, roi), dtype=torch.uint8, device=dev) if want_frames else None
    pooled = torch.empty((n_win, dims * dims), dtype=torch.uint8, device=dev) if want_pooled else None
    cnt = torch.empty((n_win,), dtype=torch.int32, device=dev)
    offs = torch.empty((n_win + 1,), dtype=torch.int64, device=dev)
    check(_lib.lib().lens_bin_events(ptr(t_us), ptr(x), ptr(y), n, t0_us, window_us, roi_x0, roi_y0,
                                     roi, k, index_shift, int(wrap_u8), ptr(frames), ptr(pooled),
                                     ptr(cnt), ptr(offs), n_win, stream_ptr()), "lens_bin_events")
    return frames, pooled, cnt


def seqmatch_topk(S, L, N=25, want_D=False):
    """S f32 [B, Q, P] -> (top_val [B, Qo, N], top_idx [B, Qo, N], D [B, Po, Qo] or None).

    lens/run_model.py:248-254 + the selection of lens/src/metrics.py:218.
    """
    require_cuda(S)
    assert S.dtype == torch.float32 and S.dim() == 3
    B, Q, P = S.shape
    Qo, Po = Q - L + 1, P - L + 1
    D = torch.empty((B, Po, Qo), dtype=torch.float32, device=S.device) if want_D else None
    tv = torch.empty((B, Qo, N), dtype=torch.float32, device=S.device)
    ti = torch.empty((B, Qo, N), dtype=torch.int32, device=S.device)
    check(_lib.lib().lens_seqmatch_topk(ptr(S), B, Q, P, L, N, ptr(D), ptr(tv), ptr(ti), stream_ptr()),
          "lens_seqmatch_topk")
    return tv, ti, D


def recall_counts(top_idx, Po, gt_dense=None, gt_center=None, gt_tol=0, ns=RECALL_NS,
                  hits=None, n_valid=None):
    """Accumulate Recall@N hit counters (device int64) for top_idx [B, Qo, N].

    gt_dense: u8 [Po, Qo] (shared by all streams) or [B, Po, Qo]; or gt_center i32 [B, Qo].
    Returns (hits [len(ns)] i64, n_valid [1] i64) on the device.
    """
    require_cuda(top_idx, gt_dense, gt_center)
    B, Qo, N = top_idx.shape
    dev = top_idx.device
    if hits is None:
        hits = torch.zeros((len(ns),), dtype=torch.int64, device=dev)
    if n_valid is None:
        n_valid = torch.zeros((1,), dtype=torch.int64, device=dev)
    stride = 0
    if gt_dense is not None:
        assert gt_dense.dtype == torch.uint8
        if gt_dense.dim() == 3:
            assert gt_dense.shape == (B, Po, Qo)
            stride = Po * Qo
        else:
            assert gt_dense.shape == (Po, Qo)
    if gt_center is not None:
        assert gt_center.dtype == torch.int32 and gt_center.shape == (B, Qo)
    ns_arr = (C.c_int * len(ns))(*ns)
    check(_lib.lib().lens_recall(ptr(top_idx), B, Qo, Po, N, ptr(gt_dense), stride, ptr(gt_center),
                                 gt_tol, ns_arr, len(ns), ptr(hits), ptr(n_valid), stream_ptr()),
          "lens_recall")
    return hits, n_valid


def topn_merge(vals, idx):
    """Global top-N from W per-shard lists: vals f32 / idx i32 [W, B, Qo, N] (idx = global place index, -1 =
    empty) -> (val [B, Qo, N], idx [B, Qo, N]) under the order (value desc, place index desc)."""
    require_cuda(vals, idx)
    assert vals.dtype == torch.float32 and idx.dtype == torch.int32 and vals.shape == idx.shape and vals.dim() == 4
    vals, idx = vals.contiguous(), idx.contiguous()
    W, B, Qo, N = vals.shape
    ov = torch.empty((B, Qo, N), dtype=torch.float32, device=vals.device)
    oi = torch.empty((B, Qo, N), dtype=torch.int32, device=vals.device)
    check(_lib.lib().lens_topn_merge(ptr(vals), ptr(idx), W, B * Qo, N, ptr(ov), ptr(oi), stream_ptr()),
          "lens_topn_merge")
    return ov, oi


def recall_bounds(D, gt_dense, ns=RECALL_NS):
    """Tie-aware (lo, hi, n_valid) hit counters of Recall@N over every order of equal similarities.

    D f32 [Po, Qo], gt_dense u8 [Po, Qo] (rows = database) -> (lo [len(ns)] i64, hi, n_valid [1]) on the device;
    lo / n_valid <= the reference's recallAtK (numpy's unstable argsort, metrics.py:218) <= hi / n_valid."""
    require_cuda(D, gt_dense)
    assert D.dtype == torch.float32 and gt_dense.dtype == torch.uint8 and D.shape == gt_dense.shape and D.dim() == 2
    D, gt_dense = D.contiguous(), gt_dense.contiguous()
    Po, Qo = D.shape
    lo = torch.zeros((len(ns),), dtype=torch.int64, device=D.device)
    hi = torch.zeros((len(ns),), dtype=torch.int64, device=D.device)
    n_valid = torch.zeros((1,), dtype=torch.int64, device=D.device)
    ns_arr = (C.c_int * len(ns))(*ns)
    check(_lib.lib().lens_recall_bounds(ptr(D), ptr(gt_dense), Po, Qo, ns_arr, len(ns), ptr(lo), ptr(hi), ptr(n_valid),
                                        stream_ptr()), "lens_recall_bounds")
    return lo, hi, n_valid


def sad_matrix(query_frames, reference_frames):
    """u8 [Q, npix], u8 [R, npix] -> L1 distances f32 [Q, R] (torch.cdist(a, b, 1) of lens/src/sad.py:38)."""
    require_cuda(query_frames, reference_frames)
    assert query_frames.dtype == torch.uint8 and reference_frames.dtype == torch.uint8
    Q, npix = query_frames.shape
    R = reference_frames.shape[0]
    assert reference_frames.shape[1] == npix
    dist = torch.empty((Q, R), dtype=torch.float32, device=query_frames.device)
    check(_lib.lib().lens_sad_matrix(ptr(query_frames), ptr(reference_frames), Q, R, npix, ptr(dist), stream_ptr()),
          "lens_sad_matrix")
    return dist


def reciprocal(x):
    """1 / x elementwise in IEEE fp32 (numpy's `1 / dist_matrix_seq`, lens/src/sad.py:52,62)."""
    require_cuda(x)
    assert x.dtype == torch.float32
    out = torch.empty_like(x)
    check(_lib.lib().lens_reciprocal(ptr(x), x.numel(), ptr(out), stream_ptr()), "lens_reciprocal")
    return out


def event_windows(t, x, y, lut, use_first_event, start, interval, max_windows, check_sorted=True):
    """Frame index ranges of the event-driven representation (lens/tools/dvstools.py:286-349).

    t f64 [n] seconds ascending, x / y 16-bit [n], lut i16 [H, W] (-2 hot pixel, -1 unused, >= 0 slot).
    -> (win_begin i64 [max_windows], win_end i64, win_t0 f64, n_windows i64 [1]); raises if t is unsorted."""
    require_cuda(t, x, y, lut)
    assert t.dtype == torch.float64 and x.dtype in (torch.int16, torch.uint16) and y.dtype == x.dtype
    assert lut.dtype == torch.int16 and lut.dim() == 2
    n = t.numel()
    H, W = lut.shape
    dev = t.device
    nw = max(int(max_windows), 0)
    wb = torch.empty(max(nw, 1), dtype=torch.int64, device=dev)
    we = torch.empty(max(nw, 1), dtype=torch.int64, device=dev)
    wt = torch.empty(max(nw, 1), dtype=torch.float64, device=dev)
    n_win = torch.zeros(1, dtype=torch.int64, device=dev)
    unsorted = torch.zeros(1, dtype=torch.int32, device=dev) if check_sorted else None
    check(_lib.lib().lens_event_windows(ptr(t), ptr(x), ptr(y), n, ptr(lut), W, H, int(bool(use_first_event)),
                                        float(start), float(interval), nw, ptr(wb), ptr(we), ptr(wt), ptr(n_win),
                                        ptr(unsorted), stream_ptr()), "lens_event_windows")
    if check_sorted and int(unsorted.item()):
        raise ValueError("event timestamps must be ascending")
    return wb, we, wt, n_win


def bin_events_lut(x, y, lut, win_begin, win_end, n_slots, weight=1, events_per_window_hint=0):
    """Slot histograms of the given event ranges -> frames u8 [n_windows, n_slots]
    (`frame_data[index] += accum_factor`, dvstools.py:310-322; counts wrap mod 256)."""
    require_cuda(x, y, lut, win_begin, win_end)
    n_win = win_begin.numel()
    H, W = lut.shape
    frames = torch.empty((n_win, n_slots), dtype=torch.uint8, device=x.device)
    counts = torch.empty((n_win, n_slots), dtype=torch.int32, device=x.device)
    check(_lib.lib().lens_bin_events_lut(ptr(x), ptr(y), ptr(lut), W, H, ptr(win_begin), ptr(win_end), n_win, n_slots,
                                         int(weight), int(events_per_window_hint), ptr(counts), ptr(frames),
                                         stream_ptr()), "lens_bin_events_lut")
    return frames

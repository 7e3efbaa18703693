"""Python face of the CPU oracle (ctypes over oracle/liblens_oracle.so + numpy).

TEST INFRASTRUCTURE ONLY -- see the header of lens_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by lens_b200/.

Reference sites restated here (relative to /root/reference):
  seed-50 raster matrix      lens/src/dataset.py:120-121
  GT slicing + dilation      lens/run_model.py:266-294
  Recall@K                   lens/src/metrics.py:183-226
  L == 0 / evaluate tail     lens/run_model.py:248-254,301-302
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblens_oracle.so")
_lib = None


def build(force=False):
    """Compile liblens_oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "lens_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liblens_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
        L.lens_oracle_bin_events.restype = i32
        L.lens_oracle_bin_events.argtypes = [vp, vp, vp, i64, C.c_uint32, C.c_uint32, i32, i32, i32,
                                             i32, i32, i32, vp, vp, vp, i64]
        L.lens_oracle_pool.restype = i32
        L.lens_oracle_pool.argtypes = [vp, i64, i32, i32, vp]
        L.lens_oracle_snn_create.restype = vp
        L.lens_oracle_snn_create.argtypes = [i32, i32, i32, i32, f32, f32, vp, vp, vp, i32]
        L.lens_oracle_snn_destroy.argtypes = [vp]
        L.lens_oracle_snn_reset.argtypes = [vp]
        L.lens_oracle_snn_get_state.argtypes = [vp, vp, vp, vp]
        L.lens_oracle_snn_overflow.restype = i64
        L.lens_oracle_snn_overflow.argtypes = [vp]
        L.lens_oracle_snn_forward.restype = i32
        L.lens_oracle_snn_forward.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.lens_oracle_snn_forward_float.restype = i32
        L.lens_oracle_snn_forward_float.argtypes = [vp, vp, i32, i32, vp]
        L.lens_oracle_seqmatch.restype = i32
        L.lens_oracle_seqmatch.argtypes = [vp, i32, i32, i32, vp]
        L.lens_oracle_topk.restype = i32
        L.lens_oracle_topk.argtypes = [vp, i32, i32, i32, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# --------------------------------------------------------------------------- R2
def raster_uniforms(T, roi, k):
    """U[T, dims*dims]: the sub-sampled columns of the reference's raster matrix.

    dataset.py:120-121 draws torch.rand(T, roi*roi) after torch.manual_seed(50)
    for EVERY image; the pooling conv (run_model.py:130-137) then keeps pixel
    (k*i+c, k*j+c).  torch's CPU MT19937 stream is the definition, so torch is
    used to draw it (plumbing, not arithmetic).
    """
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(50)
    U = torch.rand(T, roi * roi, generator=g).numpy()
    return np.ascontiguousarray(U[:, pool_index(roi, k)])


def pool_index(roi, k):
    """Flat indices (into a roi*roi frame) picked by the one-hot strided conv."""
    dims = roi // k
    c = (k // 2) - 1
    if c < 0:
        c += k
    ii = k * np.arange(dims) + c
    return (ii[:, None] * roi + ii[None, :]).reshape(-1)


def pool(frames, k):
    frames = _c(frames, np.uint8)
    n, roi, _ = frames.shape
    dims = roi // k
    out = np.empty((n, dims * dims), np.uint8)
    rc = lib().lens_oracle_pool(_p(frames), n, roi, k, _p(out))
    assert rc == 0
    return out


# --------------------------------------------------------------------------- R0
def bin_events(t_us, x, y, t0_us, window_us, n_win, roi, k, roi_x0=0, roi_y0=0,
               index_shift=1, wrap_u8=True):
    t_us, x, y = _c(t_us, np.uint32), _c(x, np.uint16), _c(y, np.uint16)
    dims = roi // k
    frames = np.empty((n_win, roi, roi), np.uint8)
    pooled = np.empty((n_win, dims * dims), np.uint8)
    cnt = np.empty((n_win,), np.int32)
    rc = lib().lens_oracle_bin_events(_p(t_us), _p(x), _p(y), t_us.shape[0], t0_us, window_us,
                                      roi_x0, roi_y0, roi, k, index_shift, int(wrap_u8),
                                      _p(frames), _p(pooled), _p(cnt), n_win)
    assert rc == 0, rc
    return frames, pooled, cnt


# ---------------------------------------------------------------------- R3 - R6
class OracleSNN:
    """The converted sinabs network (IAF -> Linear -> IAF -> Linear -> IAF), stateful."""

    def __init__(self, W_feat, W_out, U, T, thr=1.0, v_min=-1.0, n_streams=1):
        self.W_feat, self.W_out = _c(W_feat, np.float32), _c(W_out, np.float32)
        self.F, self.I = self.W_feat.shape
        self.P = self.W_out.shape[0]
        assert self.W_out.shape[1] == self.F
        self.T, self.B = T, n_streams
        self.U = _c(U, np.float32) if U is not None else None
        if self.U is not None:
            assert self.U.shape == (T, self.I)
        self.h = lib().lens_oracle_snn_create(self.I, self.F, self.P, T, thr, v_min, _p(self.W_feat),
                                              _p(self.W_out), _p(self.U), n_streams)
        if not self.h:
            raise RuntimeError("lens_oracle_snn_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().lens_oracle_snn_destroy(self.h)
            self.h = None

    def reset_states(self):
        lib().lens_oracle_snn_reset(self.h)

    def state(self):
        v0 = np.empty((self.B, self.I), np.float32)
        v1 = np.empty((self.B, self.F), np.float32)
        v2 = np.empty((self.B, self.P), np.float32)
        lib().lens_oracle_snn_get_state(self.h, _p(v0), _p(v1), _p(v2))
        return v0, v1, v2

    def overflow(self):
        return int(lib().lens_oracle_snn_overflow(self.h))

    def run_streams(self, pooled, want_steps=False):
        """pooled u8 [B, Q, I] -> counts f32 [B, Q, P] (state carried over)."""
        pooled = _c(pooled, np.uint8)
        B, Q, I = pooled.shape
        assert I == self.I and B <= self.B
        counts = np.empty((B, Q, self.P), np.float32)
        hid = np.empty((B, Q * self.T, self.F), np.uint8) if want_steps else None
        out = np.empty((B, Q * self.T, self.P), np.uint8) if want_steps else None
        rc = lib().lens_oracle_snn_forward(self.h, _p(pooled), B, Q, _p(counts), _p(hid), _p(out))
        assert rc == 0, rc
        return (counts, hid, out) if want_steps else counts

    def forward_float(self, x):
        """x f32 [B, steps, I] (pooled float raster) -> spikes f32 [B, steps, P]."""
        x = _c(x, np.float32)
        B, steps, I = x.shape
        assert I == self.I and B <= self.B
        out = np.empty((B, steps, self.P), np.float32)
        rc = lib().lens_oracle_snn_forward_float(self.h, _p(x), B, steps, _p(out))
        assert rc == 0, rc
        return out


# --------------------------------------------------------------------------- R7
def seqmatch(S, L):
    """S [Q, P] -> D [P-L+1, Q-L+1] f32 (rows = database); L == 0 returns S unchanged."""
    if L == 0:
        return np.asarray(S)                      # run_model.py:254
    S = _c(S, np.float32)
    Q, P = S.shape
    D = np.empty((P - L + 1, Q - L + 1), np.float32)
    rc = lib().lens_oracle_seqmatch(_p(S), Q, P, L, _p(D))
    assert rc == 0, rc
    return D


def topk(D, K):
    """Best K database rows per query column, ties -> larger row index first."""
    D = _c(D, np.float32)
    Po, Qo = D.shape
    idx = np.empty((Qo, K), np.int32)
    val = np.empty((Qo, K), np.float32)
    rc = lib().lens_oracle_topk(_p(D), Po, Qo, K, _p(idx), _p(val))
    assert rc == 0, rc
    return idx, val


# --------------------------------------------------------------------------- R8
def make_gt_tol(GT, L, tol):
    """run_model.py:268-294: slice by the sequence length, dilate by `tol`, transpose."""
    from scipy.ndimage import binary_dilation
    GT = np.asarray(GT)
    if L != 0:
        GT = GT[L - 2:-1, L - 2:-1]
    se = np.ones((2 * tol + 1, 2 * tol + 1), dtype=int)
    return binary_dilation(GT, structure=se).astype(int).T


# --------------------------------------------------------------------------- R9
def recall_at_k(S_in, GThard, GTsoft=None, K=1, kind=None):
    """Recall@K of metrics.py:183-226.

    kind=None reproduces the reference call exactly (numpy's default, unstable
    argsort); kind='stable' is the documented deterministic tie rule the CUDA
    kernel implements (larger database index wins a tie).
    """
    S_in, GThard = np.asarray(S_in), np.asarray(GThard)
    assert S_in.shape == GThard.shape and S_in.ndim == 2 and K >= 1
    gt = GThard.astype(bool)
    S = S_in.copy()
    if GTsoft is not None:
        soft = np.asarray(GTsoft).astype(bool)
        S[soft & ~gt] = S.min()
    has_match = gt.sum(0) > 0
    S, gt = S[:, has_match], gt[:, has_match]
    order = S.argsort(0) if kind is None else S.argsort(0, kind=kind)
    best = order[-K:, :]
    cols = np.tile(np.arange(best.shape[1]), [K, 1])
    found = gt[best, cols]
    return np.sum(found.sum(0) > 0) / found.shape[1]


def recall_bounds(D, GTtol, K):
    """Tie-aware [lower, upper] bound on Recall@K over every possible tie order."""
    D, gt = np.asarray(D), np.asarray(GTtol).astype(bool)
    keep = gt.sum(0) > 0
    D, gt = D[:, keep], gt[:, keep]
    lo = hi = 0
    for q in range(D.shape[1]):
        col = D[:, q]
        kth = np.sort(col)[-K] if K <= col.size else -np.inf
        sure = col > kth
        tie = col == kth
        room = K - int(sure.sum())
        if gt[sure, q].any():
            lo += 1
            hi += 1
            continue
        n_tie, n_tie_pos = int(tie.sum()), int(gt[tie, q].sum())
        if n_tie_pos > 0:
            hi += 1
            if n_tie - n_tie_pos < room:   # cannot avoid picking a positive
                lo += 1
    n = D.shape[1]
    return lo / n, hi / n


def evaluate_tail(S, GT, L, tol, Ns=(1, 5, 10, 15, 20, 25), kind=None):
    """run_model.py:247-302 after the similarity matrix is known -> (D, GTtol, R)."""
    D = seqmatch(np.asarray(S, dtype=np.float64), L)
    GTtol = make_gt_tol(GT, L, tol)
    R = [round(recall_at_k(D, GTtol, K=n, kind=kind), 2) for n in Ns]
    return D, GTtol, R


def create_pr(S_in, GThard, n_thresh=100, matching="single", GTsoft=None):
    """createPR of lens/src/metrics.py:21-139 without the figure -> (P, R) lists.
    matching='single': best match per query (:59-68); 'multi': every entry of the matrix (:70-72)."""
    S, GT = np.asarray(S_in).copy(), np.asarray(GThard).astype(bool)
    if GTsoft is not None:
        S[np.asarray(GTsoft).astype(bool) & ~GT] = S.min()
    if matching == "single":
        gtp = np.count_nonzero(GT.any(0))
        best_row = np.argmax(S, axis=0)
        hit = GT[best_row, np.arange(GT.shape[1])]
        val = np.max(S, axis=0)
    else:
        gtp = np.count_nonzero(GT)
        hit, val = GT, S
    P, R = [1], [0]
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in np.linspace(val.max(), val.min(), n_thresh):
            sel = val >= t
            tp, fp = np.count_nonzero(hit & sel), np.count_nonzero(~hit & sel)
            P.append(np.float64(tp) / np.float64(tp + fp))
            R.append(np.float64(tp) / np.float64(gtp))
    return P, R


def online_match(sequence, L):
    """run_speck.py:200-204 restated with plain loops: convolve2d(sequence.T, eye(L), mode='same') / L
    and the first-maximum argmax of every column.  sequence: int [R, P] (rows oldest first).
    'same' keeps the [P, R] window of the full convolution that starts at offset (L-1)//2 in both axes."""
    seq = np.asarray(sequence, dtype=np.int64)
    R, P = seq.shape
    o = (L - 1) // 2
    result = np.zeros((P, R), dtype=np.float64)
    for p in range(P):
        for r in range(R):
            acc = 0
            for a in range(L):
                rr, pp = r + o - a, p + o - a
                if 0 <= rr < R and 0 <= pp < P:
                    acc += int(seq[rr, pp])
            result[p, r] = np.float64(acc) / np.float64(L)
    return result, np.argmax(result, axis=0)


class OnlineMatcherOracle:
    """State machine of run_speck.py:155-226 (custom_readout + seq_match), one push() per readout."""

    def __init__(self, reference_places, sequence_length, readouts_per_row=4, rows_per_match=4):
        self.P, self.L = reference_places, sequence_length
        self.per_row, self.per_match = readouts_per_row, rows_per_match
        self.sum = np.zeros(self.P, dtype=np.int64)     # run_speck.py:159-164 (dict feature -> count)
        self.qry = 0
        self.sequence = None
        self.matrix = None

    def push(self, counts):
        self.sum += np.asarray(counts).astype(np.int64)
        self.qry += 1                                   # :169
        if self.qry != self.per_row:                    # :180
            return None
        self.qry = 0                                    # :226
        row = self.sum // self.per_row                  # :195,198 (sum keeps accumulating across rows)
        self.sequence = row[None] if self.sequence is None else np.vstack((self.sequence, row))
        if self.sequence.shape[0] != self.per_match:    # :200
            return None
        result, arg = online_match(self.sequence, self.L)
        self.matrix = result if self.matrix is None else np.concatenate((self.matrix, result), axis=1)   # :213-217
        self.sum = np.zeros(self.P, dtype=np.int64)     # :221
        self.sequence = None                            # :222
        return arg, result


def simple_rep(t, x, y, dimensions, unique_indices, centroid_dict, hot_pixels, timebin_fps, offset=0.0,
               accum_factor=1.0, frames_max=900, frame_limit=False):
    """FrameRep.event_data, text branch with tool='simple_rep' (lens/tools/dvstools.py:173-349), restated
    event by event.  dimensions = (height, width); centroid_dict {flat pixel: flat centroid};
    -> (frames u8 [n, pixels] in the order the reference saves them, offset the reference ends up with)."""
    H, W = dimensions
    pixels = len(unique_indices)
    frame_interval = 1.0 / timebin_fps                               # :174
    start = current = None
    if offset != 0:                                                  # :175-177
        start = current = offset
    frame = np.zeros(pixels, dtype=np.uint8)
    saved = {}
    frame_number = 0
    uniq = np.asarray(unique_indices)
    for ts, xe, ye in zip(t, x, y):
        ts, xe, ye = float(ts), int(xe), int(ye)
        if offset == 0:                                              # :286-290
            offset = ts
            start = current = ts
        if ts < start or (hot_pixels and (xe, ye) in hot_pixels):    # :293-294
            continue
        if abs(ts - current) <= frame_interval:                      # :297
            flat = ye * W + xe
            if flat in centroid_dict:                                # :310-323
                slot = int(np.where(uniq == centroid_dict[flat])[0][0])
                frame[slot] = np.uint8(int(float(frame[slot]) + accum_factor) & 255)
        else:                                                        # :326-349
            saved[frame_number] = frame.copy()                       # save_frame(frame_data, frame_number)
            if not ts <= offset:
                frame_number += 1
            frame = np.zeros(pixels, dtype=np.uint8)
            current = ts
            if frame_number >= frames_max and frame_limit:
                break
    n = max(saved) + 1 if saved else 0
    return np.stack([saved[i] for i in range(n)]) if n else np.zeros((0, pixels), np.uint8), offset

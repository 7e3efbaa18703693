"""Error behaviour at the C-ABI boundary: bad arguments come back as negative return codes with a message
(LensError in the Python layer), never as a crash, and leave the library usable."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bad_arguments_are_reported_not_fatal():
    from lens_b200 import _lib, ops, synth
    from lens_b200._lib import LensError
    from lens_b200.network import B200Network
    from lens_b200.online import OnlineMatcher
    lib = _lib.lib()
    S = torch.zeros((2, 4, 50), dtype=torch.float32, device="cuda")
    with pytest.raises(LensError, match="1 <= L"):
        ops.seqmatch_topk(S, 5, 3)                       # L > Q
    with pytest.raises(LensError, match="outside"):
        ops.seqmatch_topk(S, 2, 65)                      # N above the supported maximum
    # misaligned event arrays
    t = torch.zeros(64, dtype=torch.int32, device="cuda")
    xy = torch.zeros(65, dtype=torch.int16, device="cuda")
    with pytest.raises(LensError, match="aligned"):
        ops.bin_events(t, xy[1:], xy[1:], 0, 1000, 1, 16, 2)
    with pytest.raises(LensError, match="window_us"):
        ops.bin_events(t, xy[:64], xy[:64], 0, 0, 1, 16, 2)
    # network: wrong input shapes / missing input / too many streams through the raw ABI
    Wf, Wo = synth.weights(100, 200, 64, seed=1)
    net = B200Network(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, num_timesteps=25, max_streams=2)
    with pytest.raises(LensError, match="multiple of num_timesteps"):
        net(torch.zeros((26, 1, 80, 80), device="cuda"))
    with pytest.raises(LensError, match="expected"):
        net(torch.zeros((25, 1, 40, 40), device="cuda"))
    with pytest.raises(LensError, match="exactly one"):
        net.run_streams()
    with pytest.raises(LensError, match="pooled must be"):
        net.run_streams(pooled=torch.zeros((2, 1, 99), dtype=torch.uint8, device="cuda"))
    pooled = torch.zeros((3, 1, 100), dtype=torch.uint8, device="cuda")
    counts = torch.zeros((3, 1, 64), dtype=torch.float32, device="cuda")
    rc = lib.lens_snn_forward(net._h, C.c_void_p(pooled.data_ptr()), 3, 1, C.c_void_p(counts.data_ptr()), None, None,
                              0, None)
    assert rc < 0 and b"exceeds max_streams" in lib.lens_last_error()
    rc = lib.lens_snn_forward(net._h, C.c_void_p(pooled.data_ptr()), 2, 1, C.c_void_p(counts.data_ptr()), None, None,
                              7, None)
    assert rc < 0 and b"bad mode" in lib.lens_last_error()
    rc = lib.lens_snn_forward(None, None, 1, 1, None, None, None, 0, None)
    assert rc < 0 and b"NULL handle" in lib.lens_last_error()
    with pytest.raises(LensError):
        B200Network(torch.zeros((2000, 100)), torch.zeros((10, 2000)), roi=80, k=8, num_timesteps=5)   # F too large
    with pytest.raises(LensError):
        B200Network(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, num_timesteps=25, device="cpu")
    # recall: exactly one kind of ground truth
    ti = torch.zeros((1, 3, 5), dtype=torch.int32, device="cuda")
    with pytest.raises((LensError, AssertionError)):
        ops.recall_counts(ti, 50)
    with pytest.raises(ValueError):
        OnlineMatcher(10, 2).push(torch.zeros(10, dtype=torch.float64, device="cuda"))
    with pytest.raises(LensError):
        ops.sad_matrix(torch.zeros((2, 70000), dtype=torch.uint8, device="cuda"),
                       torch.zeros((2, 70000), dtype=torch.uint8, device="cuda"))
    # ... and the library still works afterwards
    out = net.run_streams(pooled=torch.full((2, 1, 100), 200, dtype=torch.uint8, device="cuda"))
    assert out.shape == (2, 1, 64) and torch.isfinite(out).all()


def test_bin_events_sortedness_check():
    """lens_bin_events needs ascending timestamps (binary-searched windows); the optional check finds a
    single descent anywhere in the array (vector body, tail, across the 16-byte seams)."""
    from lens_b200 import ops
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 64, 1001, 4096 + 3):
        t = np.sort(rng.integers(0, 5000, n)).astype(np.int32)
        x = rng.integers(0, 16, n).astype(np.int16)
        args = (torch.from_numpy(x).cuda(), torch.from_numpy(x).cuda(), 0, 1000, 5, 16, 2)
        ops.bin_events(torch.from_numpy(t).cuda(), *args, check_sorted=True)          # sorted: fine
        for pos in sorted({0, n // 2, n - 2, min(3, n - 2), min(4, n - 2)}):
            if pos < 0 or pos + 1 >= n or t[pos] == t[-1] + 7:
                continue
            bad = t.copy()
            bad[pos] = t[-1] + 7                                                         # a descent after pos
            with pytest.raises(ValueError, match="ascending"):
                ops.bin_events(torch.from_numpy(bad).cuda(), *args, check_sorted=True)

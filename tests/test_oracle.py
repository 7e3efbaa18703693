"""CPU tests: the oracle against the golden vectors produced by the reference's own code.

Goldens: tests/golden/make_golden.py ran /root/reference's run_model.LENS.evaluate,
dataset.CustomImageDataset and metrics.recallAtK (sinabs restated, see sinabs_stub.py).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _setup(g):
    roi, dims, T = int(g["roi_dim"]), int(g["dims"]), int(g["timebin"])
    k = roi // dims
    U = O.raster_uniforms(T, roi, k)
    pooled = O.pool(g["frames"], k)
    return roi, k, T, U, pooled


def test_raster_uniforms_hash(golden):
    import hashlib
    import torch
    g = golden("config1")
    torch.manual_seed(50)
    U = torch.rand(int(g["timebin"]), int(g["roi_dim"]) ** 2).numpy()
    assert hashlib.sha256(U.tobytes()).hexdigest() == str(g["U_sha256"])
    sub = O.raster_uniforms(int(g["timebin"]), int(g["roi_dim"]), 8)
    assert np.array_equal(sub, U[:, O.pool_index(80, 8)])


def test_pool_matches_onehot_conv():
    import torch
    rng = np.random.default_rng(0)
    for roi, k in [(80, 8), (7, 1), (12, 3), (10, 2)]:
        fr = rng.integers(0, 256, (5, roi, roi)).astype(np.uint8)
        kern = torch.zeros(1, 1, k, k)
        c = (k // 2) - 1
        kern[0, 0, c, c] = 1   # lens/run_model.py:130-134 (c = -1 wraps for k = 1)
        ref = torch.nn.functional.conv2d(torch.from_numpy(fr).float()[:, None], kern, stride=k)
        assert np.array_equal(O.pool(fr, k), ref.reshape(5, -1).numpy().astype(np.uint8))


def test_config1_counts_bit_exact(golden):
    """Bundled example model + data: spike counts, hidden spikes, per-step outputs identical."""
    g = golden("config1")
    roi, k, T, U, pooled = _setup(g)
    net = O.OracleSNN(g["W_feat"], g["W_out"], U, T)
    counts, hid, out = net.run_streams(pooled[None], want_steps=True)
    Q = pooled.shape[0]
    assert np.array_equal(counts[0], g["S"].astype(np.float32))
    assert np.array_equal(hid[0].reshape(Q, T, -1).sum(1), g["hidden_counts"])
    n = g["hidden_steps"].shape[0]
    assert np.array_equal(hid[0][:n * T].reshape(n, T, -1), g["hidden_steps"])
    assert np.array_equal(out[0][:n * T].reshape(n, T, -1), g["out_steps"])
    v0, v1, v2 = net.state()
    # membrane potentials: within 1e-5 of the BLAS-ordered reference (absolute, |v| <= ~1)
    assert np.abs(v1[0] - g["v1"]).max() < 1e-5
    assert np.abs(v2[0] - g["v2"]).max() < 1e-5
    assert np.array_equal(v0[0], g["v0"])
    assert net.overflow() == 0


def test_config1_tail_identical(golden):
    g = golden("config1")
    L, tol = int(g["sequence_length"]), int(g["GT_tolerance"])
    D, GTtol, R = O.evaluate_tail(g["S"].astype(np.float64), g["GT"], L, tol)
    assert np.array_equal(D, g["D"])
    assert np.array_equal(GTtol, g["GTtol"])
    assert np.allclose(R, g["R"])


@pytest.mark.timeout(300)
def test_brisevent_counts(golden):
    """Second bundled model (641 places, 724 queries, k = 1).  The reference's BLAS summation
    order moves two spikes across a query boundary (2 / 464084 entries differ by one count);
    everything else is identical and Recall@N is the same."""
    g = golden("brisevent")
    roi, k, T, U, pooled = _setup(g)
    assert k == 1
    net = O.OracleSNN(g["W_feat"], g["W_out"], U, T)
    counts = net.run_streams(pooled[None])[0]
    S = g["S"].astype(np.float32)
    diff = counts != S
    assert diff.sum() <= 4 and np.abs(counts - S).max() <= 1
    L, tol = int(g["sequence_length"]), int(g["GT_tolerance"])
    D, GTtol, R = O.evaluate_tail(counts, g["GT"], L, tol)
    assert np.array_equal(GTtol, g["GTtol"])
    assert np.allclose(R, g["R"])
    Dg, _, Rg = O.evaluate_tail(S, g["GT"], L, tol)
    assert np.array_equal(Dg, g["D"]) and np.allclose(Rg, g["R"])


def test_seqmatch_matches_conv2d():
    import torch
    rng = np.random.default_rng(3)
    for Q, P, L in [(9, 11, 1), (9, 11, 2), (16, 12, 5), (10, 10, 10)]:
        S = rng.integers(0, 40, (Q, P)).astype(np.float64)
        t = torch.tensor(S)[None, None].to(torch.float32)
        w = torch.eye(L)[None, None]
        ref = torch.nn.functional.conv2d(t, w).squeeze(0).squeeze(0).numpy() / L   # run_model.py:249-252
        ref = ref.T
        assert np.array_equal(O.seqmatch(S, L), ref.reshape(P - L + 1, Q - L + 1))
    assert O.seqmatch(S, 0) is not None


def test_topk_is_stable_argsort():
    rng = np.random.default_rng(4)
    D = rng.integers(0, 5, (37, 9)).astype(np.float32) / 2   # many ties
    for K in (1, 5, 25, 40):
        idx, val = O.topk(D, K)
        ref = np.argsort(D, axis=0, kind="stable")[-K:][::-1].T
        kk = min(K, D.shape[0])
        assert np.array_equal(idx[:, :kk], ref[:, :kk])
        assert (idx[:, kk:] == -1).all()
        assert np.array_equal(val[:, :kk], np.take_along_axis(D.T, ref[:, :kk], 1))


def test_recall_variants(golden):
    g = golden("config1")
    D, GT = g["D"], g["GTtol"]
    for K in (1, 5, 10, 25):
        lo, hi = O.recall_bounds(D, GT, K)
        r_def = O.recall_at_k(D, GT, K=K)
        r_stable = O.recall_at_k(D, GT, K=K, kind="stable")
        assert lo - 1e-12 <= r_def <= hi + 1e-12
        assert lo - 1e-12 <= r_stable <= hi + 1e-12
        # stable rule from the top-K lists
        idx, _ = O.topk(D, K)
        keep = GT.astype(bool).sum(0) > 0
        hit = [GT[idx[q][idx[q] >= 0], q].any() for q in range(D.shape[1]) if keep[q]]
        assert abs(np.mean(hit) - r_stable) < 1e-12


def test_bin_events_reference_loop():
    """Oracle binning == the literal per-event loop of collect_data.py:193-202, including the events at
    x == roi / y == roi that the reference's `frame[y-1, x-1]` still accepts (the Speck crop is 81
    columns wide); only indices the reference would raise IndexError on are cropped."""
    rng = np.random.default_rng(5)
    roi, n = 16, 4000
    t = np.sort(rng.integers(0, 1000, n)).astype(np.uint32)
    x = rng.integers(0, roi + 3, n).astype(np.uint16)     # 0 .. roi + 2: roi is valid, roi + 1 is not
    y = rng.integers(0, roi + 3, n).astype(np.uint16)
    x[:600] = 3
    y[:600] = 0   # > 255 hits on one pixel of window 0 -> uint8 wrap; y = 0 -> row -1
    t[:600] = 5
    t.sort()
    frames, pooled, cnt = O.bin_events(t, x, y, 0, 250, 4, roi, 4)
    edge = 0
    for w in range(4):
        fr = np.zeros((roi, roi), dtype=np.int64)
        sel = (t >= 250 * w) & (t < 250 * (w + 1))
        kept = 0
        for xe, ye in zip(x[sel].astype(int), y[sel].astype(int)):
            try:
                fr[ye - 1, xe - 1] += 1
                kept += 1
                edge += xe == roi or ye == roi
            except IndexError:
                pass
        assert np.array_equal(frames[w], fr.astype(np.uint8))
        assert cnt[w] == kept
    assert edge > 50                       # the x == roi / y == roi column and row really are exercised
    assert np.array_equal(pooled, O.pool(frames, 4))
    assert frames.astype(int).sum() != cnt.sum()   # the wrap really happened


def test_forward_float_matches_raster_path(golden):
    g = golden("config1")
    roi, k, T, U, pooled = _setup(g)
    Q = 3
    a = O.OracleSNN(g["W_feat"], g["W_out"], U, T)
    ca, _, outa = a.run_streams(pooled[None, :Q], want_steps=True)
    b = O.OracleSNN(g["W_feat"], g["W_out"], None, T)
    p = pooled[:Q].astype(np.float32) / np.float32(255)
    x = (U[None] < p[:, None, :]).astype(np.float32).reshape(1, Q * T, -1)
    sb = b.forward_float(x)
    assert np.array_equal(sb[0], outa[0].astype(np.float32))
    assert np.array_equal(sb[0].reshape(Q, T, -1).sum(1), ca[0])


def test_create_pr_matches_reference(golden):
    """createPR(matching='single') restatement == the reference's own function on both bundled runs."""
    for name in ("config1", "brisevent"):
        g = golden(name)
        P, R = O.create_pr(g["D"].T, g["GTtol"].T)
        assert np.array_equal(np.array(P, dtype=np.float64), g["PR_P"], equal_nan=True)
        assert np.array_equal(np.array(R, dtype=np.float64), g["PR_R"], equal_nan=True)


def test_online_match_against_scipy():
    """The online matcher's arithmetic is scipy.signal.convolve2d(mode='same') (run_speck.py:202): pin the
    loop restatement against scipy itself, the library the reference calls."""
    from scipy.signal import convolve2d
    rng = np.random.default_rng(11)
    for P, R, L in [(7, 4, 1), (25, 4, 2), (100, 4, 4), (33, 4, 5), (9, 6, 3), (5, 4, 7)]:
        seq = rng.integers(0, 40, size=(R, P))
        result, arg = O.online_match(seq, L)
        ref = convolve2d(seq.T, np.eye(L, dtype=np.float32), mode="same") / L
        assert ref.dtype == result.dtype and np.array_equal(ref, result)
        assert np.array_equal(arg, np.argmax(ref, axis=0))


def test_online_matcher_state_machine():
    """4 readouts -> one row (cumulative sum // 4), 4 rows -> one match, then the sums restart."""
    rng = np.random.default_rng(12)
    P, L = 30, 3
    m = O.OnlineMatcherOracle(P, L)
    pushes = rng.integers(0, 9, size=(32, P)).astype(np.float32)
    outs = [m.push(c) for c in pushes]
    done = [i for i, o in enumerate(outs) if o is not None]
    assert done == [15, 31]
    cum = np.cumsum(pushes[:16].astype(np.int64), axis=0)
    seq = np.stack([cum[3], cum[7], cum[11], cum[15]]) // 4
    assert np.array_equal(outs[15][1], O.online_match(seq, L)[0])
    assert m.matrix.shape == (P, 8)


def _load_event_cases():
    import json
    g = np.load(os.path.join(GOLDEN, "events_simple_rep.npz"))
    W, H = (int(v) for v in g["sensor"])
    for name in ("plain", "hot_gaps", "offset", "limit"):
        lines = [str(s) for s in g[name + "/lines"]]
        t = np.array([float(l.split()[0]) for l in lines])       # float(): what the reference calls per field
        x = np.array([int(float(l.split()[1])) for l in lines])
        y = np.array([int(float(l.split()[2])) for l in lines])
        args = dict(timebin=10.0, accum_factor=1.0, offset=0.0, frames_max=900, frame_limit=False, pixels=25)
        args.update(json.loads(str(g[name + "/args"])))
        hot = set((int(a), int(b)) for a, b in g[name + "/hot"]) or None
        cdict = {int(k): int(v) for k, v in zip(g[name + "/dict_keys"], g[name + "/dict_vals"])}
        yield name, g, (W, H), lines, t, x, y, args, hot, cdict


def test_simple_rep_oracle_against_reference_frames():
    """The event-by-event restatement reproduces every PNG the reference's FrameRep wrote (reference and
    query mode write the same frames from the same layout)."""
    n_frames = 0
    for name, g, (W, H), lines, t, x, y, args, hot, cdict in _load_event_cases():
        uniq = g[name + "/unique_indices"]
        frames, offset = O.simple_rep(t, x, y, (H, W), uniq, cdict, hot, args["timebin"], args["offset"],
                                      args["accum_factor"], args["frames_max"], args["frame_limit"])
        side = int(np.sqrt(args["pixels"]))
        want = g[name + "/frames_ref"]
        assert np.array_equal(want, g[name + "/frames_qry"])
        assert frames.reshape(-1, side, side).shape == want.shape, name
        assert np.array_equal(frames.reshape(-1, side, side), want), name
        assert offset == float(g[name + "/offset_after"])
        assert want.any()
        n_frames += len(want)
    assert n_frames > 50


def test_event_text_reader_parses_like_float():
    """Host reader (pandas C parser, round-trip precision) == the reference's float()/int() per field."""
    import tempfile
    from lens_b200.tools import dvstools
    for name, g, (W, H), lines, t, x, y, args, hot, cdict in _load_event_cases():
        with tempfile.TemporaryDirectory() as d:
            import zipfile
            zp = os.path.join(d, "syn.zip")
            with zipfile.ZipFile(zp, "w") as z:
                z.writestr("syn.txt", "{} {}\n".format(W, H) + "\n".join(lines) + "\n")
                z.writestr("event_sum.txt", str(len(lines)))
            ev = dvstools.read_events_zip(zp)
            assert (ev["width"], ev["height"], ev["event_sum"]) == (W, H, len(lines))
            assert np.array_equal(ev["t"], t) and np.array_equal(ev["x"], x) and np.array_equal(ev["y"], y)
            # writer round trip (ExtractRosbag's '{:.12f}' format)
            zp2 = os.path.join(d, "again.zip")
            dvstools.write_events_zip(zp2, ev["t"], ev["x"], ev["y"], ev["p"], W, H)
            ev2 = dvstools.read_events_zip(zp2)
            assert np.array_equal(ev2["t"], t) and np.array_equal(ev2["p"], ev["p"])
        break


def test_patch_layout_matches_reference_layout():
    """make_patch_layout draws the same centroids / patch ownership as dvstools.py:221-244 for the same
    numpy global seed (the golden run used np.random.seed(7))."""
    from lens_b200.tools import dvstools
    for name, g, (W, H), lines, t, x, y, args, hot, cdict in _load_event_cases():
        np.random.seed(7)
        uniq, d = dvstools.make_patch_layout((H, W), args["pixels"])
        assert np.array_equal(uniq, g[name + "/unique_indices"])
        assert d == cdict and list(d.keys()) == list(cdict.keys())
        lut = dvstools.layout_lut((H, W), uniq, d, hot)
        assert lut.shape == (H, W) and (lut >= 0).sum() == len(set(d) - {hy * W + hx for hx, hy in (hot or ())})


def test_dataset_csv_matches_reference_writer(tmp_path):
    """create_csv_from_images writes the bytes the reference's writer wrote for the same folder listing."""
    from lens_b200.tools.create_data_csv import create_csv_from_images, haversine
    g = np.load(os.path.join(GOLDEN, "events_simple_rep.npz"))
    for f in g["plain/files_ref"]:
        (tmp_path / str(f)).write_bytes(b"")
    (tmp_path / "notes.txt").write_text("ignored")
    out = tmp_path / "out.csv"
    create_csv_from_images(str(tmp_path), str(out))
    assert open(out).read() == str(g["plain/csv"])        # the golden was read in text mode too
    assert abs(haversine(153.0, -27.5, 153.0, -27.5)) == 0 and 110e3 < haversine(153.0, -27.0, 153.0, -28.0) < 112e3


def test_evaluate_tail_option_space(golden):
    """Sequence matching, GT slicing / dilation and Recall@N for other --sequence_length / --GT_tolerance
    values, against the reference's own run on the bundled example (tests/golden/options.npz)."""
    g = np.load(os.path.join(GOLDEN, "options.npz"))
    c1 = golden("config1")
    S, GT = c1["S"].astype(np.float64), c1["GT"]
    n = 0
    for L, tol in g["variants"]:
        key = f"L{L}_tol{tol}"
        if str(g[key + "/error"]):
            assert L == 1 and "same shape" in str(g[key + "/error"])
            D, GTtol = O.seqmatch(S, L), O.make_gt_tol(GT, L, tol)
            assert D.shape != GTtol.shape          # the reference's assertion fires for the same reason
            continue
        D, GTtol, R = O.evaluate_tail(S, GT, int(L), int(tol))
        assert np.array_equal(D.astype(np.float32), g[key + "/D"]), key
        assert np.array_equal(GTtol, g[key + "/GTtol"]), key
        for got, want, K in zip(R, g[key + "/R"], (1, 5, 10, 15, 20, 25)):
            lo, hi = O.recall_bounds(D, GTtol, K)
            assert round(lo, 2) <= want <= round(hi, 2), (key, K)       # tie-aware bounds hold for the reference
            assert got == want, (key, K)                                  # same numpy, same unstable argsort
        n += 1
    assert n == 7


def test_create_pr_multi_matches_reference(golden):
    """The oracle's createPR(matching='multi') == the reference's own lists (tests/golden/make_golden_pr_multi.py)."""
    pm = golden("pr_multi")
    for name in ("config1", "brisevent"):
        g = golden(name)
        S, GT = g["D"].T, g["GTtol"].T
        for n in (100, 7):
            P, R = O.create_pr(S, GT, n_thresh=n, matching="multi")
            assert np.array_equal(np.array(P, np.float64), pm[f"{name}/n{n}/P"], equal_nan=True)
            assert np.array_equal(np.array(R, np.float64), pm[f"{name}/n{n}/R"], equal_nan=True)
        soft = np.roll(GT, 1, axis=1) | GT
        P, R = O.create_pr(S, GT, n_thresh=50, matching="multi", GTsoft=soft)
        assert np.array_equal(np.array(P, np.float64), pm[f"{name}/soft/P"], equal_nan=True)
        assert np.array_equal(np.array(R, np.float64), pm[f"{name}/soft/R"], equal_nan=True)
    P, R = O.create_pr(pm["random/S"], pm["random/GT"], n_thresh=20, matching="multi")
    assert np.array_equal(np.array(P, np.float64), pm["random/P"], equal_nan=True)
    assert np.array_equal(np.array(R, np.float64), pm["random/R"], equal_nan=True)

"""Recall@K with the reference's call signature (lens/src/metrics.py:183-226), on the GPU.

`recallAtK(S_in, GThard, GTsoft=None, K=1)` takes numpy matrices [database, query] and returns a
float.  The K best rows of every query column are selected by the CUDA top-N kernel
(lens_seqmatch_topk with L = 1) under the deterministic rule "larger value first, larger
database index first among equals" -- the order `np.argsort(kind='stable')[-K:]` yields; the
reference's default (unstable) argsort may break ties differently.
"""
import numpy as np
import torch

from .. import ops

MAX_K = 64


def _topk_hits(S, GT, K):
    """hits / valid for one K through the CUDA kernels.  S, GT: [Po, Qo] numpy."""
    Po, Qo = S.shape
    # the kernel consumes query-major similarity rows: S_in[r][q] = S_t[q][r]
    S_t = torch.from_numpy(np.ascontiguousarray(S.T, dtype=np.float32)).cuda()[None]
    _, ti, _ = ops.seqmatch_topk(S_t, 1, K)
    gt = torch.from_numpy(np.ascontiguousarray(GT, dtype=np.uint8)).cuda()
    hits, nv = ops.recall_counts(ti, Po, gt_dense=gt, ns=(K,))
    return int(hits[0].item()), int(nv.item())


def recallAtK(S_in, GThard, GTsoft=None, K=1):
    S_in, GThard = np.asarray(S_in), np.asarray(GThard)
    assert S_in.shape == GThard.shape, "S_in and GThard must have the same shape"
    assert S_in.ndim == 2, "S_in, GThard and GTsoft must be two-dimensional"
    assert K >= 1, "K must be >=1"
    if K > MAX_K:
        raise ValueError(f"K={K} exceeds the supported maximum {MAX_K}")
    GT = GThard.astype(bool)
    S = S_in.astype(np.float32, copy=True)
    if GTsoft is not None:
        soft = np.asarray(GTsoft).astype(bool)
        assert S_in.shape == soft.shape, "S_in and GTsoft must have the same shape"
        S[soft & ~GT] = S.min()
    hits, valid = _topk_hits(S, GT, K)
    return hits / valid if valid else float("nan")


def createPR(S_in, GThard, outputdir=None, datatype="LENS", GTsoft=None, matching="multi", n_thresh=100):
    """Precision / recall lists of lens/src/metrics.py:21-139 for matching='single' (the only mode the
    reference's inference path uses, lens/run_model.py:321).  S_in, GThard: numpy [database, query].
    Returns (P, R): python lists of length n_thresh + 1 starting with P = 1, R = 0.  The figure the
    reference saves on the last threshold is not produced."""
    from .._lib import lib, check, ptr, stream_ptr
    S_in, GThard = np.asarray(S_in), np.asarray(GThard)
    assert S_in.shape == GThard.shape, "S_in, GThard and GTsoft must have the same shape"
    assert S_in.ndim == 2, "S_in, GThard and GTsoft must be two-dimensional"
    assert matching in ("single", "multi"), "matching should contain one of the following strings: [single, multi]"
    assert n_thresh > 1, "n_thresh must be >1"
    if matching != "single":
        raise NotImplementedError("lens_b200 implements createPR for matching='single' only")
    GT = GThard.astype(bool)
    S = S_in.astype(np.float32, copy=True)
    if GTsoft is not None:
        soft = np.asarray(GTsoft).astype(bool)
        S[soft & ~GT] = S.min()
    Po, Qo = S.shape
    dS = torch.from_numpy(np.ascontiguousarray(S)).cuda()
    dG = torch.from_numpy(np.ascontiguousarray(GT, dtype=np.uint8)).cuda()
    tp = torch.zeros(n_thresh, dtype=torch.int64, device="cuda")
    fp = torch.zeros(n_thresh, dtype=torch.int64, device="cuda")
    gtp = torch.zeros(1, dtype=torch.int64, device="cuda")
    check(lib().lens_pr_counts(ptr(dS), ptr(dG), Po, Qo, n_thresh, ptr(tp), ptr(fp), ptr(gtp), stream_ptr()),
          "lens_pr_counts")
    tp, fp, gtp = tp.cpu().numpy(), fp.cpu().numpy(), int(gtp.item())
    P, R = [1], [0]
    with np.errstate(divide="ignore", invalid="ignore"):
        for a, b in zip(tp, fp):
            P.append(np.float64(a) / np.float64(a + b))
            R.append(np.float64(a) / np.float64(gtp))
    return P, R

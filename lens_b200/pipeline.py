"""The hot path as one object: frames -> pooled pixels -> spike counts -> sequence matching ->
top-N -> Recall@N counters, for a shard of independent query streams on one GPU.

This is what `LENS.evaluate` does for one stream and what bench.py / the multi-GPU driver do for
thousands: every stage is a call into liblens_b200.so; torch only owns the buffers.  With
`torch.distributed` initialised, streams are sharded contiguously across ranks (one process per
GPU, no data-path collective) and the Recall@N counters are summed with one all-reduce
(SURVEY.md 8e).
"""
import torch

from . import ops
from .network import B200Network, MODE_AUTO


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n units owned by `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class InferencePipeline:
    def __init__(self, W_feat, W_out, roi, k, T, L, n_top=25, ns=ops.RECALL_NS, max_streams=1,
                 device=None, mode=MODE_AUTO):
        self.net = B200Network(W_feat, W_out, roi=roi, k=k, num_timesteps=T, max_streams=max_streams,
                               device=device)
        self.L, self.n_top, self.ns, self.mode = int(L), int(n_top), tuple(ns), mode
        self.device = self.net.device

    def similarity(self, frames=None, pooled=None):
        """u8 frames [B, Q, roi, roi] (or pooled [B, Q, I]) -> spike counts f32 [B, Q, P]."""
        return self.net.run_streams(frames=frames, pooled=pooled, mode=self.mode)

    def match(self, S, gt_dense=None, gt_center=None, gt_tol=0, want_D=False):
        """S [B, Q, P] -> dict(top_val, top_idx, D, hits, n_valid); hits are per-rank counters."""
        tv, ti, D = ops.seqmatch_topk(S, self.L, self.n_top, want_D=want_D)
        out = dict(top_val=tv, top_idx=ti, D=D, hits=None, n_valid=None)
        if gt_dense is not None or gt_center is not None:
            Po = S.shape[2] - self.L + 1
            out["hits"], out["n_valid"] = ops.recall_counts(ti, Po, gt_dense=gt_dense,
                                                            gt_center=gt_center, gt_tol=gt_tol, ns=self.ns)
        return out

    def step(self, frames=None, pooled=None, gt_dense=None, gt_center=None, gt_tol=0, reduce=True):
        """One pass of the hot path over this rank's shard; Recall counters all-reduced if distributed."""
        S = self.similarity(frames=frames, pooled=pooled)
        out = self.match(S, gt_dense=gt_dense, gt_center=gt_center, gt_tol=gt_tol)
        out["S"] = S
        if reduce and out["hits"] is not None and torch.distributed.is_available() \
                and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            packed = torch.cat([out["hits"], out["n_valid"]])
            torch.distributed.all_reduce(packed)          # sum of 6 hit counters + valid count
            out["hits"], out["n_valid"] = packed[:-1], packed[-1:]
        return out

    @staticmethod
    def recall(hits, n_valid):
        nv = int(n_valid.item())
        return [float(h) / nv if nv else float("nan") for h in hits.tolist()]

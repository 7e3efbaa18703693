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


def place_shard_range(P, L, rank, world):
    """Places a rank holds when the DATABASE is sharded: its own contiguous range [p0, p1) of the
    P - L + 1 sequence-matched rows plus the L - 1 places after it (every diagonal that starts in the
    range is complete without an exchange).  -> (p0, p1, p1_with_halo)."""
    Po = P - max(L, 1) + 1
    p0, p1 = shard_range(Po, rank, world)
    return p0, p1, min(P, p1 + max(L, 1) - 1)


class InferencePipeline:
    def __init__(self, W_feat, W_out, roi, k, T, L, n_top=25, ns=ops.RECALL_NS, max_streams=1,
                 device=None, mode=MODE_AUTO):
        self.net = B200Network(W_feat, W_out, roi=roi, k=k, num_timesteps=T, max_streams=max_streams,
                               device=device)
        self.L, self.n_top, self.ns, self.mode = int(L), int(n_top), tuple(ns), mode
        self.device = self.net.device
        # optional per-stage CUDA-event timing (bench.py): [(start, after similarity, after match)]
        self.stage_timing = False
        self._stage_events = []

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
        ev = None
        if self.stage_timing:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        S = self.similarity(frames=frames, pooled=pooled)
        if ev:
            ev[1].record()
        out = self.match(S, gt_dense=gt_dense, gt_center=gt_center, gt_tol=gt_tol)
        if ev:
            ev[2].record()
            self._stage_events.append(ev)
        out["S"] = S
        out["overflow"] = self.net.overflow_tensor()   # device counter, no synchronisation: check with .item()
        if reduce and out["hits"] is not None and torch.distributed.is_available() \
                and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            packed = torch.cat([out["hits"], out["n_valid"]])
            torch.distributed.all_reduce(packed)          # sum of 6 hit counters + valid count
            out["hits"], out["n_valid"] = packed[:-1], packed[-1:]
        return out

    def step_host(self, frames_host, gt_dense=None, gt_center=None, gt_tol=0, next_frames=None, reduce=True):
        """The same step fed from PINNED HOST frames u8 [B, Q, roi, roi].

        The host->device copy runs on a side stream into one of two staging buffers.  Passing the
        following step's host tensor as `next_frames` starts its copy right after this step's kernels
        are queued, so in a stream of steps the PCIe transfer of step k+1 hides behind the compute of
        step k (every step still pays for its own copy; only the first one is exposed)."""
        net, dev = self.net, self.device
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage, self._stage_free, self._pending = [None, None], [None, None], {}
            self._slot = 0

        def start_copy(src):
            slot = self._slot
            self._slot ^= 1
            if self._stage[slot] is None or self._stage[slot].shape != src.shape:
                self._stage[slot] = torch.empty(src.shape, dtype=torch.uint8, device=dev)
                self._stage_free[slot] = None
                # the allocator may hand back a block that kernels queued on the main stream still read
                self._copy_stream.wait_stream(main)
            with torch.cuda.stream(self._copy_stream):
                if self._stage_free[slot] is not None:
                    self._copy_stream.wait_event(self._stage_free[slot])   # last reader of this buffer
                self._stage[slot].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._pending[src.data_ptr()] = (slot, ev)

        if frames_host.data_ptr() not in self._pending:
            start_copy(frames_host)
        slot, ev = self._pending.pop(frames_host.data_ptr())
        main.wait_event(ev)
        out = self.step(frames=self._stage[slot], gt_dense=gt_dense, gt_center=gt_center, gt_tol=gt_tol,
                        reduce=reduce)
        self._stage_free[slot] = torch.cuda.Event()
        self._stage_free[slot].record(main)
        if next_frames is not None:
            start_copy(next_frames)
        return out

    def pop_stage_timing(self):
        """-> dict(similarity_ms, match_ms, n) summed over the steps recorded since the last call."""
        torch.cuda.synchronize(self.device)
        sim = sum(a.elapsed_time(b) for a, b, _ in self._stage_events)
        mat = sum(b.elapsed_time(c) for _, b, c in self._stage_events)
        n = len(self._stage_events)
        self._stage_events = []
        return dict(similarity_ms=sim, match_ms=mat, n=n)

    @staticmethod
    def recall(hits, n_valid):
        nv = int(n_valid.item())
        return [float(h) / nv if nv else float("nan") for h in hits.tolist()]


class PlaceShardedPipeline:
    """The same hot path with the DATABASE sharded across ranks instead of the streams (the
    alternative for few streams / very large databases): every rank runs ALL streams against its own range
    of places (output layer is column-parallel; the tiny feature layer is replicated), ranks them locally,
    and the per-rank top-N lists are all-gathered (NCCL) and merged on the GPU (lens_topn_merge).  The merged
    lists equal the single-GPU lists bit for bit: the order (value desc, place index desc) is global."""

    def __init__(self, W_feat, W_out, roi, k, T, L, n_top=25, ns=ops.RECALL_NS, max_streams=1, device=None,
                 mode=MODE_AUTO, rank=None, world=None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.rank = rank if rank is not None else (self.dist.get_rank() if self.dist else 0)
        self.world = world if world is not None else (self.dist.get_world_size() if self.dist else 1)
        self.P = int(W_out.shape[0])
        self.L, self.n_top, self.ns, self.mode = int(L), int(n_top), tuple(ns), mode
        self.p0, self.p1, p1h = place_shard_range(self.P, self.L, self.rank, self.world)
        self.net = B200Network(W_feat, W_out[self.p0:p1h], roi=roi, k=k, num_timesteps=T, max_streams=max_streams,
                               device=device)
        self.device = self.net.device

    def step(self, frames=None, pooled=None, gt_center=None, gt_tol=0):
        """-> dict(top_val, top_idx (global place indices), hits, n_valid); every rank holds the merged result."""
        S = self.net.run_streams(frames=frames, pooled=pooled, mode=self.mode)       # [B, Q, local places + halo]
        tv, ti, _ = ops.seqmatch_topk(S, self.L, self.n_top)
        ti = torch.where(ti >= 0, ti + self.p0, ti)
        if self.world > 1:
            allv = torch.empty((self.world,) + tuple(tv.shape), dtype=tv.dtype, device=tv.device)
            alli = torch.empty((self.world,) + tuple(ti.shape), dtype=ti.dtype, device=ti.device)
            self.dist.all_gather_into_tensor(allv, tv.contiguous())
            self.dist.all_gather_into_tensor(alli, ti.contiguous())
        else:
            allv, alli = tv[None], ti[None]
        tv, ti = ops.topn_merge(allv, alli)
        out = dict(top_val=tv, top_idx=ti, hits=None, n_valid=None)
        if gt_center is not None:
            out["hits"], out["n_valid"] = ops.recall_counts(ti, self.P - self.L + 1, gt_center=gt_center,
                                                            gt_tol=gt_tol, ns=self.ns)
        return out

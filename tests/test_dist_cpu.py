"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: stream sharding and the Recall@N
counter merge.  The CUDA kernels are not involved; the oracle stands in for the per-rank compute."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_everything():
    from lens_b200.pipeline import shard_range
    for n in (0, 1, 7, 8, 1000, 65536):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_place_shard_range_covers_every_diagonal():
    """Database sharding: the rows [p0, p1) of all ranks partition the P - L + 1 sequence-matched rows, and
    with the L - 1 halo places each rank can form every diagonal that starts in its range."""
    from lens_b200.pipeline import place_shard_range
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    for P, L, world in [(40, 2, 2), (41, 10, 3), (100, 1, 8), (64, 4, 1)]:
        S = rng.poisson(1.3, (L + 3, P)).astype(np.float32)
        D = O.seqmatch(S, L)                                   # [P - L + 1, Q - L + 1]
        rows = []
        for r in range(world):
            p0, p1, p1h = place_shard_range(P, L, r, world)
            assert p1h - p0 >= (p1 - p0) and p1h <= P
            if p1 > p0:
                Dr = O.seqmatch(S[:, p0:p1h], L)               # the shard's own matrix
                assert Dr.shape[0] == p1 - p0
                rows.append(Dr)
        assert np.array_equal(np.concatenate(rows), D)


def _worker(rank, world, port, B, Q, P, L, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lens_b200.pipeline import shard_range
    from oracle import oracle as O
    rng = np.random.default_rng(0)                      # same global problem on every rank
    S = rng.poisson(1.3, (B, Q, P)).astype(np.float32)
    centre = rng.integers(0, P - L + 1, (B, Q - L + 1))
    lo, hi = shard_range(B, rank, world)
    ns = (1, 5, 10)
    hits = np.zeros(len(ns) + 1, dtype=np.int64)
    for b in range(lo, hi):                             # this rank's streams only
        idx, _ = O.topk(O.seqmatch(S[b], L), max(ns))
        for q in range(Q - L + 1):
            hits[-1] += 1
            for i, n in enumerate(ns):
                hits[i] += int((np.abs(idx[q, :n] - centre[b, q]) <= 2).any())
    t = torch.from_numpy(hits)
    dist.all_reduce(t)                                   # the merge InferencePipeline.step performs
    if rank == 0:
        np.save(os.path.join(tmp, "merged.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_rank_recall_merge_equals_single_process(tmp_path):
    from oracle import oracle as O
    B, Q, P, L = 5, 6, 40, 2
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, B, Q, P, L, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / "merged.npy")
    rng = np.random.default_rng(0)
    S = rng.poisson(1.3, (B, Q, P)).astype(np.float32)
    centre = rng.integers(0, P - L + 1, (B, Q - L + 1))
    ns = (1, 5, 10)
    want = np.zeros(len(ns) + 1, dtype=np.int64)
    for b in range(B):
        idx, _ = O.topk(O.seqmatch(S[b], L), max(ns))
        for q in range(Q - L + 1):
            want[-1] += 1
            for i, n in enumerate(ns):
                want[i] += int((np.abs(idx[q, :n] - centre[b, q]) <= 2).any())
    assert np.array_equal(merged, want)

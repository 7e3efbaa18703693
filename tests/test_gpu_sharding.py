"""GPU tests of the two ways the path shards, on ONE device: the per-rank pieces are run
one after the other and merged with the same kernels the multi-GPU run uses, and must reproduce the
unsharded result bit for bit.  The NCCL run itself is tests/test_gpu_multi.py (needs >= 2 GPUs)."""
import numpy as np
import pytest
import torch

from lens_b200 import synth

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("W,N,M", [(1, 25, 7), (2, 25, 100), (8, 25, 333), (16, 32, 5), (3, 1, 64)])
def test_topn_merge_equals_global_sort(W, N, M):
    """lens_topn_merge == sorting the union of the W lists by (value desc, place index desc)."""
    from lens_b200 import ops
    rng = np.random.default_rng(W * 100 + N)
    P = 4000
    vals = np.empty((W, 1, M, N), np.float32)
    idx = np.empty((W, 1, M, N), np.int32)
    want_v, want_i = np.empty((M, N), np.float32), np.empty((M, N), np.int32)
    for m in range(M):
        places = rng.permutation(P)[:W * N].reshape(W, N)          # distinct places across shards
        v = rng.integers(0, 6, (W, N)).astype(np.float32) / 4     # many ties
        if m % 5 == 0:                                             # a shard with fewer than N candidates
            places[0, N // 2:], v[0, N // 2:] = -1, -np.inf
        for w in range(W):                                         # every list arrives sorted
            order = np.lexsort((-places[w], -v[w]))
            vals[w, 0, m], idx[w, 0, m] = v[w][order], places[w][order]
        keep = places.reshape(-1) >= 0
        pv, pi = v.reshape(-1)[keep], places.reshape(-1)[keep]
        order = np.lexsort((-pi, -pv))[:N]
        want_v[m], want_i[m] = -np.inf, -1
        want_v[m, :len(order)], want_i[m, :len(order)] = pv[order], pi[order]
    ov, oi = ops.topn_merge(cuda(vals), cuda(idx))
    assert np.array_equal(oi[0].cpu().numpy(), want_i)
    assert np.array_equal(ov[0].cpu().numpy(), want_v)


@pytest.mark.parametrize("world,P,L", [(2, 1000, 2), (8, 2050, 10), (3, 777, 1)])
def test_place_sharded_pipeline_equals_unsharded(world, P, L):
    """Database sharded over `world` ranks (own rows + L - 1 halo places each), per-rank top-N, merged:
    identical to ranking the whole database at once."""
    from lens_b200 import ops
    from lens_b200.pipeline import InferencePipeline, PlaceShardedPipeline, place_shard_range
    B, Q = 6, max(3, L + 1)
    Wf, Wo = synth.weights(100, 200, P, seed=11)
    frames = cuda(synth.frames(B, Q, 80, seed=12))
    gt = cuda(synth.gt_centers(B, Q - L + 1, P - L + 1, seed=13))
    full = InferencePipeline(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, T=250, L=L, max_streams=B)
    ref = full.step(frames=frames, gt_center=gt, gt_tol=2)
    lists_v, lists_i, covered = [], [], []
    for r in range(world):
        # rank r's shard of the maths (its network holds rows p0 .. p1 + L - 2 of W_out); the collective is
        # replaced by running the ranks one after the other
        part = PlaceShardedPipeline(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, T=250, L=L,
                                    max_streams=B, rank=r, world=world)
        p1h = place_shard_range(P, L, r, world)[2]
        S = part.net.run_streams(frames=frames)
        assert torch.equal(S, ref["S"][:, :, part.p0:p1h])            # column-parallel output layer
        tv, ti, _ = ops.seqmatch_topk(S, L, 25)
        lists_v.append(tv)
        lists_i.append(torch.where(ti >= 0, ti + part.p0, ti))
        covered.append((part.p0, part.p1))
    assert covered[0][0] == 0 and covered[-1][1] == P - max(L, 1) + 1
    mv, mi = ops.topn_merge(torch.stack(lists_v), torch.stack(lists_i))
    assert torch.equal(mi, ref["top_idx"]) and torch.equal(mv, ref["top_val"])
    hits, nv = ops.recall_counts(mi, P - L + 1, gt_center=gt, gt_tol=2)
    assert torch.equal(hits, ref["hits"]) and torch.equal(nv, ref["n_valid"])

#!/usr/bin/env python
"""Multi-GPU check (run under torchrun with the nccl backend, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Streams are sharded across ranks; the all-reduced Recall@N counters and the gathered spike counts
must equal what one rank computes alone for the whole batch.  Also exercises the all-gather of a
row-sharded place database (config 5 of BASELINE.json) before the network is built, and the
place-sharded mode (every rank ranks its own places for all streams, NCCL all-gather of the per-rank
top-N lists, lens_topn_merge): the merged lists must equal the single-rank lists bit for bit.
Also run by tests/test_gpu_multi.py (pytest -m gpu) when the box has >= 2 GPUs.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lens_b200 import synth  # noqa: E402
from lens_b200.pipeline import InferencePipeline, PlaceShardedPipeline, shard_range  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, Q, P, F, L = 64, 4, 2048, 200, 2
    Wf, Wo = synth.weights(100, F, P, seed=1)
    # database arrives row-sharded: every rank owns P / world places and all-gathers the rest
    lo_p, hi_p = shard_range(P, rank, world)
    assert (hi_p - lo_p) * world == P
    shard = torch.from_numpy(Wo[lo_p:hi_p]).to(dev)
    full = torch.empty((P, F), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(full, shard)
    assert torch.equal(full.cpu(), torch.from_numpy(Wo))
    frames = synth.frames(B, Q, 80, seed=2)
    gt = synth.gt_centers(B, Q - L + 1, P - L + 1, seed=7)
    lo, hi = shard_range(B, rank, world)
    pipe = InferencePipeline(torch.from_numpy(Wf), full, roi=80, k=8, T=250, L=L, max_streams=hi - lo, device=dev)
    out = pipe.step(frames=torch.from_numpy(frames[lo:hi]).to(dev), gt_center=torch.from_numpy(gt[lo:hi]).to(dev),
                    gt_tol=2)
    gathered = [torch.empty((shard_range(B, r, world)[1] - shard_range(B, r, world)[0], Q, P), device=dev)
                for r in range(world)]
    dist.all_gather(gathered, out["S"])
    S_all = torch.cat(gathered)
    # place-sharded mode: all streams on every rank, the database split
    psp = PlaceShardedPipeline(torch.from_numpy(Wf), full, roi=80, k=8, T=250, L=L, max_streams=B, device=dev)
    pout = psp.step(frames=torch.from_numpy(frames).to(dev), gt_center=torch.from_numpy(gt).to(dev), gt_tol=2)
    ok = True
    if rank == 0:
        solo = InferencePipeline(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, T=250, L=L,
                                 max_streams=B, device=dev)
        ref = solo.step(frames=torch.from_numpy(frames).to(dev), gt_center=torch.from_numpy(gt).to(dev),
                        gt_tol=2, reduce=False)
        ok_streams = torch.equal(S_all, ref["S"]) and torch.equal(out["hits"], ref["hits"]) and \
            torch.equal(out["n_valid"], ref["n_valid"])
        ok_places = torch.equal(pout["top_idx"], ref["top_idx"]) and torch.equal(pout["top_val"], ref["top_val"]) and \
            torch.equal(pout["hits"], ref["hits"])
        ok = ok_streams and ok_places
        print(f"multi_gpu_check world={world}: counts equal={torch.equal(S_all, ref['S'])} "
              f"hits {out['hits'].tolist()} vs {ref['hits'].tolist()} valid {int(out['n_valid'])}; "
              f"place-sharded top-N merge equal={ok_places} -> {'OK' if ok else 'MISMATCH'}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

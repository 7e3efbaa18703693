#!/usr/bin/env python
"""Small pass through every kernel of the hot path for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck  python tools/sanitize_step.py
    compute-sanitizer --tool racecheck python tools/sanitize_step.py
    compute-sanitizer --tool synccheck python tools/sanitize_step.py

Sizes are tiny (the tools slow kernels down by 10-100x) but exercise the same code: binning, pooling, both
SNN paths (CUDA-core and tcgen05 digit-plane kernels incl. ragged chunks, odd stream count, 5- and 6-plane
tiles), sequence matching, top-N, merge, recall, recall bounds.  Results are checked against each other.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lens_b200 import ops, synth  # noqa: E402
from lens_b200.network import B200Network, MODE_SIMT, MODE_TC  # noqa: E402

roi, k, T, L, B, Q, I, F, P = 16, 2, 40, 2, 3, 3, 64, 96, 300
t, x, y, n_win = synth.events(B * Q * 3000, sensor=roi, window_us=1000, events_per_window=3000, seed=3)
f, p, c = ops.bin_events(torch.from_numpy(t.view(np.int32)).cuda(), torch.from_numpy(x.view(np.int16)).cuda(),
                         torch.from_numpy(y.view(np.int16)).cuda(), 0, 1000, B * Q, roi, k, check_sorted=True)
assert torch.equal(p, ops.pool_frames(f, k))
Wf, Wo = synth.weights(I, F, P, seed=1)
Wo[::7] *= np.float32(2.0 ** -14)          # rows spanning many binary orders of magnitude -> 6-plane tiles too
Wo[::7, 0] = np.float32(0.1)
nets = [B200Network(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=roi, k=k, num_timesteps=T, max_streams=B)
        for _ in range(2)]
pooled = p.reshape(B, Q, I)
Sa = nets[0].run_streams(pooled=pooled, mode=MODE_SIMT)
Sb = nets[1].run_streams(pooled=pooled, mode=MODE_TC)
assert torch.equal(Sa, Sb), "tensor-core path != CUDA-core path"
for a, b in zip(nets[0].state(), nets[1].state()):
    assert torch.equal(a, b)
tv, ti, D = ops.seqmatch_topk(Sb, L, 25, want_D=True)
gt = (torch.rand((P - L + 1, Q - L + 1), device="cuda") < 0.05).to(torch.uint8)
hits, nv = ops.recall_counts(ti[:1], P - L + 1, gt_dense=gt)
lo, hi, nv2 = ops.recall_bounds(D[0], gt)
assert int(nv.item()) == int(nv2.item()) and bool((lo <= hits).all()) and bool((hits <= hi).all())
h = (P - L + 1) // 2
va, ia, _ = ops.seqmatch_topk(Sb[:, :, :h + L - 1].contiguous(), L, 25)
vb, ib, _ = ops.seqmatch_topk(Sb[:, :, h:].contiguous(), L, 25)
mv, mi = ops.topn_merge(torch.stack([va, vb]), torch.stack([ia, torch.where(ib >= 0, ib + h, ib)]))
assert torch.equal(mi, ti) and torch.equal(mv, tv)
torch.cuda.synchronize()
print("sanitize_step ok: overflow", nets[1].overflow(), "n_inexact", nets[1].n_inexact, "sum", float(Sb.sum()))

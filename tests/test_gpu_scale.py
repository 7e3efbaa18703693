"""GPU property tests at BASELINE.json's configuration sizes (no oracle needed at these sizes):
cross-kernel equality (tensor-core vs event-driven path, both exact), state carry-over across calls,
linearity of binning, ordering invariants of the top-N lists."""
import numpy as np
import pytest
import torch

from lens_b200 import synth

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_net(P, B, T=250, F=200):
    from lens_b200.network import B200Network
    Wf, Wo = synth.weights(100, F, P, seed=1)
    return B200Network(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, num_timesteps=T, max_streams=B)


@pytest.mark.parametrize("P,B,Q", [(1000, 96, 3), (10000, 9, 2), (1000, 1, 16)])
def test_tensor_core_path_equals_event_driven_path(P, B, Q):
    """config-2 / config-3 place counts: both output-layer kernels are exact, so every spike count and
    every membrane potential must agree bit for bit (odd stream counts exercise the phantom pair half)."""
    pooled = cuda(synth.pixel_counts((B, Q, 100), seed=5))
    a, b = make_net(P, B), make_net(P, B)
    ca = a.run_streams(pooled=pooled, mode=1)
    cb = b.run_streams(pooled=pooled, mode=2)
    assert torch.equal(ca, cb)
    for x, y in zip(a.state(), b.state()):
        assert torch.equal(x, y)
    assert float(ca.sum()) > 0 and a.overflow() == 0 and b.overflow() == 0


def test_state_carry_over_across_calls_full_size():
    """16 queries in one call == 10 + 6 queries in two calls (the reference never resets state)."""
    B, Q, P = 200, 16, 1000
    pooled = cuda(synth.pixel_counts((B, Q, 100), seed=6))
    a, b = make_net(P, B), make_net(P, B)
    one = a.run_streams(pooled=pooled, mode=2)
    two = torch.cat([b.run_streams(pooled=pooled[:, :10].contiguous(), mode=2),
                     b.run_streams(pooled=pooled[:, 10:].contiguous(), mode=2)], dim=1)
    assert torch.equal(one, two)
    a.reset_states()
    again = a.run_streams(pooled=pooled, mode=2)
    assert torch.equal(one, again)       # reset_states() really returns to the initial condition


def test_streams_are_independent():
    """A stream's counts do not depend on which other streams share the launch (sharding safety)."""
    B, Q, P = 33, 2, 1000
    pooled = cuda(synth.pixel_counts((B, Q, 100), seed=7))
    full = make_net(P, B).run_streams(pooled=pooled, mode=2)
    part = make_net(P, 5).run_streams(pooled=pooled[7:12].contiguous(), mode=2)
    assert torch.equal(full[7:12], part)


def test_binning_linearity_and_conservation():
    """bin(A u B) == bin(A) + bin(B) (mod 256); without wrap the counts sum to the event count."""
    from lens_b200 import ops
    n = 1 << 22
    t, x, y, n_win = synth.events(n, sensor=128, seed=11)
    sel = np.random.default_rng(1).random(n) < 0.5

    def run(mask, wrap=True):
        tt, xx, yy = t[mask], x[mask], y[mask]
        f, p, c = ops.bin_events(cuda(tt.view(np.int32)), cuda(xx.view(np.int16)), cuda(yy.view(np.int16)),
                                 0, 250_000, n_win, 128, 8, wrap_u8=wrap)
        return f, p, c
    fa, pa, ca = run(sel)
    fb, pb, cb = run(~sel)
    fu, pu, cu = run(np.ones(n, bool))
    assert torch.equal(fu, (fa.to(torch.int32) + fb.to(torch.int32)).remainder(256).to(torch.uint8))
    assert torch.equal(cu, ca + cb) and int(cu.sum()) == n
    from lens_b200.ops import pool_frames
    assert torch.equal(pu, pool_frames(fu, 8))
    # sparse stream (no pixel reaches 256): the frame conserves the events
    m = np.zeros(n, bool)
    m[::64] = True
    fs, _, cs = run(m, wrap=False)
    assert int(fs.to(torch.int64).sum()) == int(m.sum()) == int(cs.sum())


@pytest.mark.parametrize("B,Q,P,L", [(64, 10, 10000, 10), (8, 10, 100000, 10)])
def test_topn_invariants_large(B, Q, P, L):
    from lens_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    S = torch.poisson(torch.full((B, Q, P), 1.3, device="cuda"), generator=g)
    tv, ti, D = ops.seqmatch_topk(S, L, 25, want_D=True)
    Po = P - L + 1
    assert ti.min() >= 0 and ti.max() < Po
    assert torch.all(tv[..., :-1] >= tv[..., 1:])                               # descending values
    tie = tv[..., :-1] == tv[..., 1:]
    assert torch.all(ti[..., :-1][tie] > ti[..., 1:][tie])                       # ties: larger index first
    Dq = D.transpose(1, 2)                                                      # [B, Qo, Po]
    assert torch.equal(tv[..., 0], Dq.max(dim=2).values)
    assert torch.equal(torch.gather(Dq, 2, ti.long()), tv)
    kth = torch.topk(Dq, 25, dim=2).values
    assert torch.equal(kth, tv)
    # diagonal sum definition on a few entries
    for (b, q, r) in [(0, 0, 0), (B - 1, Q - L, Po - 1), (1, 0, 1234)]:
        # numpy float32 / int is an IEEE division (torch's CUDA scalar division multiplies by 1/L)
        want = np.float32(sum(float(S[b, q + j, r + j]) for j in range(L))) / np.float32(L)
        assert float(D[b, r, q]) == float(want)


def test_host_fed_step_equals_resident_step():
    """step_host (pinned host frames, prefetched on a side stream) == step (resident frames)."""
    from lens_b200.pipeline import InferencePipeline
    B, Q, P, L = 50, 3, 1000, 2
    Wf, Wo = synth.weights(100, 200, P, seed=1)
    frames = torch.from_numpy(synth.frames(B, Q, 80, seed=2))
    gt = cuda(synth.gt_centers(B, Q - L + 1, P - L + 1))
    a = InferencePipeline(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, T=250, L=L, max_streams=B)
    b = InferencePipeline(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, T=250, L=L, max_streams=B)
    host = frames.pin_memory()
    for _ in range(2):       # second round: state carry-over + buffer reuse
        ra = a.step(frames=frames.cuda(), gt_center=gt, gt_tol=2)
        rb = b.step_host(host, gt_center=gt, gt_tol=2, next_frames=host)
        assert torch.equal(ra["S"], rb["S"])
        assert torch.equal(ra["top_idx"], rb["top_idx"]) and torch.equal(ra["hits"], rb["hits"])


def test_forward_range_equals_full_forward():
    """lens_snn_forward_range over [0,20) + [20,50) == one lens_snn_forward over all 50 streams."""
    B, Q, P = 50, 2, 1000
    pooled = cuda(synth.pixel_counts((B, Q, 100), seed=9))
    a, b = make_net(P, B), make_net(P, B)
    full = a.run_streams(pooled=pooled, mode=2)
    parts = torch.empty_like(full)
    b.run_streams_range(pooled[:20].contiguous(), 0, parts[:20], mode=2)
    b.run_streams_range(pooled[20:].contiguous(), 20, parts[20:], mode=2)
    assert torch.equal(full, parts)
    for x, y in zip(a.state(), b.state()):
        assert torch.equal(x, y)

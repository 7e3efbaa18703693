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


def test_config3_full_size_properties():
    """BASELINE.json configs[2] at its full size on one GPU: 65 536 streams x 10 queries x 250 steps,
    P = 10 000, L = 10 (1.64e8 query timesteps, 26 GB of spike counts).  Checked through properties that
    need no oracle: planted duplicate streams give identical rows, a sample of streams equals the
    event-driven CUDA-core path bit for bit, the top-N lists are ordered and consistent with D."""
    from lens_b200 import ops
    P, B, Q, L, N = 10000, 65536, 10, 10, 25
    g = torch.Generator(device="cuda").manual_seed(17)
    # pixel counts: ~40 % zeros, otherwise 0..44 (same order of magnitude as synth.pixel_counts), made on the GPU
    pooled = torch.randint(0, 45, (B, Q, 100), device="cuda", generator=g, dtype=torch.int32)
    pooled = torch.where(torch.rand((B, Q, 100), device="cuda", generator=g) < 0.4, 0, pooled).to(torch.uint8)
    src = torch.tensor([3, 1000, 40000, 65535], device="cuda")
    dup = torch.tensor([65000, 7, 12345, 0], device="cuda")
    pooled[dup] = pooled[src]
    net = make_net(P, B)
    S = net.run_streams(pooled=pooled, mode=2)
    assert S.shape == (B, Q, P) and net.overflow() == 0
    assert torch.equal(S[dup], S[src])
    sample = torch.tensor([0, 3, 4097, 32768, 65535], device="cuda")
    ref = make_net(P, len(sample)).run_streams(pooled=pooled[sample].contiguous(), mode=1)
    assert torch.equal(S[sample], ref)
    assert float(S[sample].sum()) > 0
    tv, ti, _ = ops.seqmatch_topk(S, L, N)
    Qo, Po = Q - L + 1, P - L + 1
    assert tv.shape == (B, Qo, N) and int(ti.min()) >= 0 and int(ti.max()) < Po
    assert bool((tv[..., :-1] >= tv[..., 1:]).all())
    # the best value of a sampled stream equals the maximum of its sequence-matched column
    _, _, D = ops.seqmatch_topk(S[sample].contiguous(), L, N, want_D=True)
    assert torch.equal(D.amax(dim=1), tv[sample][..., 0])
    del S, tv, ti, D, net
    torch.cuda.empty_cache()


def test_config4_full_size_binning():
    """BASELINE.json configs[3]: 1e9 events on the 128x128 sensor, 250 ms windows.  Checksum of checksums:
    every frame's pixel sum equals the window's event count modulo 256, the window counts add up to the
    number of events, and binning the two halves of the stream separately adds up to the whole (mod 256)."""
    from lens_b200 import ops
    n, window_us = 1_000_000_000, 250_000
    g = torch.Generator(device="cuda").manual_seed(23)
    # ascending timestamps with ~131 k events per window: increments 0..3 us (mean 1.9 us); cumsum stays < 2^32
    t = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
    t[::2] += 1
    t = torch.cumsum(t, 0, dtype=torch.int64)
    n_win = int(t[-1].item()) // window_us + 1
    assert int(t[-1].item()) < 2 ** 32
    t = t.to(torch.uint32).view(torch.int32)
    x = torch.randint(0, 128, (n,), device="cuda", generator=g, dtype=torch.int16)
    y = torch.randint(0, 128, (n,), device="cuda", generator=g, dtype=torch.int16)
    f, p, c = ops.bin_events(t, x, y, 0, window_us, n_win, 128, 8)
    assert int(c.sum()) == n and n_win > 7000
    sums = f.view(n_win, -1).to(torch.int64).sum(1)
    assert torch.equal(sums % 256, c.to(torch.int64) % 256)
    assert torch.equal(p, ops.pool_frames(f, 8))
    h = n // 2 + 12344                      # multiple of 8: the library wants 16-byte aligned x / y
    fa, _, ca = ops.bin_events(t[:h].contiguous(), x[:h].contiguous(), y[:h].contiguous(), 0, window_us, n_win, 128, 8)
    fb, _, cb = ops.bin_events(t[h:].contiguous(), x[h:].contiguous(), y[h:].contiguous(), 0, window_us, n_win, 128, 8)
    assert torch.equal(c, ca + cb)
    assert torch.equal(f, (fa.to(torch.int32) + fb.to(torch.int32)).remainder(256).to(torch.uint8))
    del t, x, y, f, fa, fb
    torch.cuda.empty_cache()


def test_config5_shard_properties():
    """BASELINE.json configs[4], one rank's shard: 1 024 streams against the full 100 000-place database
    (Q = 10, L = 10, N = 25).  Same oracle-free properties as config 3."""
    from lens_b200 import ops
    P, B, Q, L, N = 100000, 1024, 10, 10, 25
    pooled = cuda(synth.pixel_counts((B, Q, 100), seed=31))
    pooled[1000] = pooled[17]
    net = make_net(P, B)
    S = net.run_streams(pooled=pooled, mode=2)
    assert net.overflow() == 0 and torch.equal(S[1000], S[17])
    sample = torch.tensor([0, 17, 1023], device="cuda")
    ref = make_net(P, len(sample)).run_streams(pooled=pooled[sample].contiguous(), mode=1)
    assert torch.equal(S[sample], ref) and float(ref.sum()) > 0
    tv, ti, _ = ops.seqmatch_topk(S, L, N)
    assert int(ti.min()) >= 0 and int(ti.max()) < P - L + 1 and bool((tv[..., :-1] >= tv[..., 1:]).all())
    _, _, D = ops.seqmatch_topk(S[sample].contiguous(), L, N, want_D=True)
    assert torch.equal(D.amax(dim=1), tv[sample][..., 0])
    assert torch.equal(torch.gather(D, 1, ti[sample].long().transpose(1, 2)).transpose(1, 2), tv[sample])


@pytest.mark.parametrize("T,F,P,B,Q", [(1, 96, 130, 3, 5), (31, 63, 100, 5, 3), (32, 200, 257, 4, 2), (33, 200, 129, 7, 3),
                                       (64, 40, 128, 2, 4), (100, 200, 1000, 9, 2), (250, 130, 300, 33, 2)])
def test_tensor_core_path_odd_shapes(T, F, P, B, Q):
    """Query-aligned chunking with ragged last chunks (T not a multiple of 32, T < 32), K and M padding (F, P not
    multiples of 32 / 128), odd stream counts (phantom pair half), stream-pair blocks with an odd last pair: both
    tensor-core layers must equal the event-driven CUDA-core kernels bit for bit, state included, over two calls."""
    from lens_b200.network import B200Network
    Wf, Wo = synth.weights(100, F, P, seed=T + F)
    nets = [B200Network(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=80, k=8, num_timesteps=T, max_streams=B)
            for _ in range(2)]
    for call in range(2):
        pooled = cuda(synth.pixel_counts((B, Q, 100), seed=20 + call))
        ca, ha, oa = nets[0].run_streams(pooled=pooled, mode=1, want_steps=True)
        cb = nets[1].run_streams(pooled=pooled, mode=2)
        assert torch.equal(ca, cb)
        assert torch.equal(ca, oa.to(torch.float32).reshape(B, Q, T, P).sum(2))       # counts = per-step spikes summed
        for x, y in zip(nets[0].state(), nets[1].state()):
            assert torch.equal(x, y)
    # and the per-step outputs of the tensor-core path (debug variant of both kernels) equal the CUDA-core ones
    pooled = cuda(synth.pixel_counts((B, Q, 100), seed=30))
    ca, ha, oa = nets[0].run_streams(pooled=pooled, mode=1, want_steps=True)
    cb, hb, ob = nets[1].run_streams(pooled=pooled, mode=2, want_steps=True)
    assert torch.equal(ca, cb) and torch.equal(ha, hb) and torch.equal(oa, ob)
    assert nets[0].overflow() == 0 and nets[1].overflow() == 0

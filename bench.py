#!/usr/bin/env python
"""Benchmark of the LENS inference hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, host cores)

Workload (BASELINE.json configs[1]): LENS architecture I=100 -> F=200 -> P=1000 places,
random-init weights with trained statistics, 1000 query streams PER GPU (weak scaling) of 16
queries x T=250 timesteps, sequence length 2, synthetic Speck-resolution count frames.
A step = one pass of the hot path over the whole batch:
    u8 frames [B,16,80,80] -> pooling -> SNN (feature + output layers) -> spike counts
    -> diagonal sequence matching + top-25 -> Recall@N counters (all-reduced when N > 1).
`value`  : query timesteps / s with the frames already resident in HBM.
`e2e`    : the same through the public API with HOST (pinned) frames: H2D of the frames and D2H
           of the recall counters + top-N indices inside the timed region.
Extra    : `binning` = events/s of the event->frame kernel on a synthetic DVS stream,
           `roofline` for the dominant kernel (output-layer contraction), `cpu_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

I_DIMS, ROI, K_POOL, T_STEPS = 10, 80, 8, 250
N_TOP, NS = 25, (1, 5, 10, 15, 20, 25)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=1000, help="query streams per GPU")
    ap.add_argument("--queries", type=int, default=16, help="queries per stream")
    ap.add_argument("--places", type=int, default=1000)
    ap.add_argument("--feature", type=int, default=200)
    ap.add_argument("--seq-len", type=int, default=2)
    ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 CUDA-core, 2 tensor-core output layer")
    ap.add_argument("--events", type=int, default=1 << 28, help="events for the binning measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-binning", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def load_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (None if the workload differs)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        t = json.load(open(path)).get(kernel)
        return t["dram_bytes_per_launch"] if t and t["workload"] == workload else None
    except (OSError, ValueError, KeyError):
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"],
                    bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed regions run (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower() == "active":
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
def cpu_arm(args, seconds, steps=1, warmup=0):
    """Times oracle/ (C restatement of the reference's path) on a bounded sample of the workload.

    One python thread per host core, each with its own oracle network and its own streams (ctypes
    releases the GIL); a sample step = `cores` streams x Q queries x T timesteps through
    pooling -> SNN -> sequence matching -> top-N.
    """
    from lens_b200 import synth
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    I, F, P, Q, L = I_DIMS * I_DIMS, args.feature, args.places, args.queries, args.seq_len
    Wf, Wo = synth.weights(I, F, P, seed=1)
    U = O.raster_uniforms(T_STEPS, ROI, K_POOL)
    # calibrate: one stream, one query
    net = O.OracleSNN(Wf, Wo, U, T_STEPS)
    fr = synth.frames(1, 1, ROI, seed=99)
    t0 = time.perf_counter()
    net.run_streams(O.pool(fr[0], K_POOL)[None])
    per_query = time.perf_counter() - t0
    streams_per_thread = max(1, int(seconds / max(steps + warmup, 1) / (per_query * Q)))
    nets = [O.OracleSNN(Wf, Wo, U, T_STEPS, n_streams=streams_per_thread) for _ in range(cores)]
    frames = [synth.frames(streams_per_thread, Q, ROI, seed=100 + c) for c in range(cores)]

    def work(c):
        pooled = O.pool(frames[c].reshape(-1, ROI, ROI), K_POOL).reshape(streams_per_thread, Q, I)
        S = nets[c].run_streams(pooled)
        for b in range(streams_per_thread):
            O.topk(O.seqmatch(S[b], L), N_TOP)

    times = []
    for s in range(warmup + steps):
        th = [threading.Thread(target=work, args=(c,)) for c in range(cores)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    units = cores * streams_per_thread * Q * T_STEPS
    sample = (f"{cores} threads x {streams_per_thread} streams x {Q} queries x {T_STEPS} steps per step "
              f"(P={P}), C oracle port: pool + SNN + seq-match + top-{N_TOP}")
    return dict(value=units * len(times) / sum(times), unit="query_timesteps/s", cores=cores, kind="port",
                sample=sample), float(np.mean(times)) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("LENS_BENCH_CPU_SECONDS", max(20.0, 4.0 * (args.steps + args.warmup))))
    cb, ms = cpu_arm(args, seconds=budget, steps=args.steps, warmup=args.warmup)
    line = dict(metric="query_timesteps_per_sec", value=cb["value"], unit="query_timesteps/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32+int64", data="synthetic",
                impl="reference", config=workload_config(args, args.gpus), cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit="query_timesteps/s", h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0),
                note="reference = Python/sinabs (not installable offline); timed: the C oracle port of its "
                     "path on all host cores")
    emit(line)


def workload_config(args, world):
    return dict(workload="config2: LENS I=100 F=%d P=%d random-init, %d streams/GPU x %d queries x T=%d, "
                         "L=%d, synthetic 80x80 Speck count frames" % (args.feature, args.places,
                                                                      args.streams, args.queries, T_STEPS,
                                                                      args.seq_len),
                streams_per_gpu=args.streams, queries_per_stream=args.queries, timesteps=T_STEPS,
                places=args.places, sequence_length=args.seq_len, parallelism="stream-sharded x%d" % world,
                l2="flushed between timed iterations (256 MiB memset)")


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Route everything libraries print to stdout (e.g. NCCL's version banner) to stderr so that the
    ONE JSON line is the only thing on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from lens_b200 import synth, ops, _lib
    from lens_b200.pipeline import InferencePipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    B, Q, L = args.streams, args.queries, args.seq_len
    I, F, P = I_DIMS * I_DIMS, args.feature, args.places
    Wf, Wo = synth.weights(I, F, P, seed=1)
    pipe = InferencePipeline(torch.from_numpy(Wf), torch.from_numpy(Wo), roi=ROI, k=K_POOL, T=T_STEPS,
                             L=L, n_top=N_TOP, ns=NS, max_streams=B, device=dev, mode=args.mode)
    frames_host = torch.from_numpy(synth.frames(B, Q, ROI, seed=2 + rank)).pin_memory()
    frames_dev = frames_host.to(dev)
    gt_center = torch.from_numpy(synth.gt_centers(B, Q - L + 1, P - L + 1, seed=7 + rank)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    units_per_step = B * Q * T_STEPS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return pipe.step(frames=frames_dev, gt_center=gt_center, gt_tol=2)

    def step_e2e():
        # public API fed from pinned host memory: H2D of this step's frames (side stream; the copy for the
        # following step is started behind this step's kernels) ... D2H of the step's result
        out = pipe.step_host(frames_host, gt_center=gt_center, gt_tol=2, next_frames=frames_host)
        res = torch.cat([out["hits"], out["n_valid"]]).cpu()
        idx = out["top_idx"].cpu()
        return res, idx

    def timed(fn, steps):
        evs = []
        for _ in range(steps):
            flush.zero_()                                                # evict L2 between iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # samples through warm-up and both timed regions
    for _ in range(max(args.warmup, 3)):
        step_resident()
    for _ in range(2):
        step_e2e()
    barrier()
    pipe.net.set_timing(True)
    launches0 = _lib.launch_count()
    barrier()
    ms = timed(step_resident, args.steps)
    barrier()
    launches = _lib.launch_count() - launches0
    ktime = pipe.net.get_timing()
    pipe.net.set_timing(False)
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    last = step_resident()
    recall = pipe.recall(last["hits"], last["n_valid"])

    tot = torch.tensor([sum(ms), sum(ms_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)                       # max over ranks
    tot_ms, tot_e2e_ms = tot.tolist()
    value = world * units_per_step * args.steps / (tot_ms / 1e3)
    e2e_value = world * units_per_step * args.steps / (tot_e2e_ms / 1e3)

    # ---- roofline of the dominant kernel: output-layer contraction + IAF (tensor bound)
    out_ms = ktime["output_ms"] / max(ktime["n_output"], 1)                # average launch duration
    total_flops = 2.0 * F * P * units_per_step * args.steps                # algorithmic, all launches
    achieved = total_flops / (ktime["output_ms"] / 1e3) / 1e12 if ktime["output_ms"] > 0 else 0.0
    roofline = dict(kernel="snn output layer (F->P contraction + IAF#2)", bound="tensor", achieved=achieved,
                    peak=peaks["bf16_sustained"], unit="TFLOP/s", frac=achieved / peaks["bf16_sustained"],
                    traffic=load_traffic("output_tc_kernel", "streams=%d,queries=%d,places=%d,feature=%d" %
                                         (B, Q, P, F)),
                    peak_source=peaks["source"] + " bf16 sustained",
                    kernel_ms=out_ms, feature_ms_per_step=ktime["feature_ms"] / max(ktime["n_output"], 1),
                    share_of_step=ktime["output_ms"] / sum(ms),
                    algorithmic="2*F*P FLOP per query timestep (output layer); 2*I*F more in the feature kernel")

    line = dict(metric="query_timesteps_per_sec", value=value, unit="query_timesteps/s", n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=tot_ms / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32+int64",
                data="synthetic", config=workload_config(args, world),
                e2e=dict(value=e2e_value, unit="query_timesteps/s",
                         h2d_bytes_per_step=int(frames_host.numel()),
                         d2h_bytes_per_step=int(7 * 8 + B * (Q - L + 1) * N_TOP * 4),
                         ms_per_step=tot_e2e_ms / args.steps),
                gpu_launches=int(launches), roofline=roofline, clocks=clocks,
                recall_at_n=dict(zip(map(str, NS), recall)))

    # ---- binning: events/s (K1), HBM-bound
    if not args.no_binning:
        n_ev = args.events
        t, x, y, n_win = synth.events(min(n_ev, 1 << 24), sensor=128, seed=5 + rank)
        reps = max(1, n_ev // t.shape[0])                                # tile the stream in time
        span = int(n_win) * 250_000
        assert span * reps < 2 ** 32
        t_all = np.concatenate([t.astype(np.int64) + r * span for r in range(reps)]).astype(np.uint32)
        t_dev = torch.from_numpy(t_all.view(np.int32)).to(dev)
        x_dev = torch.from_numpy(np.tile(x, reps).view(np.int16)).to(dev)
        y_dev = torch.from_numpy(np.tile(y, reps).view(np.int16)).to(dev)
        n_tot, n_win_tot = x_dev.numel(), int(n_win) * reps

        def bin_step():
            return ops.bin_events(t_dev, x_dev, y_dev, 0, 250_000, n_win_tot, 128, 8)
        for _ in range(3):
            bin_step()
        bms = timed(bin_step, max(3, args.steps))
        bt = torch.tensor([float(np.mean(bms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
        bsec = bt.item() / 1e3
        alg_bytes = 8.0 * n_tot + n_win_tot * (128 * 128 + I)
        line["binning"] = dict(metric="events_per_sec", value=world * n_tot / bsec, unit="events/s",
                               events_per_gpu=n_tot, windows=n_win_tot, ms=bsec * 1e3,
                               roofline=dict(bound="hbm", achieved=alg_bytes / bsec / 1e9, peak=peaks["hbm_gbs"],
                                             unit="GB/s", frac=alg_bytes / bsec / 1e9 / peaks["hbm_gbs"],
                                             traffic=load_traffic("bin_kernel", "events=%d" % n_tot),
                                             peak_source=peaks["source"],
                                             algorithmic="8 B/event + 1 B per frame pixel + I B per frame"))
        del t_dev, x_dev, y_dev

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_arm(args, seconds=args.cpu_seconds)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

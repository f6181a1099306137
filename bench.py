#!/usr/bin/env python
"""Headline benchmark: train samples/s of unet_convlstm_attention, 48x72, seq_len 6, bf16
(BASELINE.json configs[2]; the same per-GPU workload on every rank for N>1 = configs[3], weak scaling).

    python bench.py --gpus N --steps K --warmup W          # our arm (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W  # CPU reference arm (oracle port, bounded sample)

A "step" = H2D (e2e only) + zero-grad + forward + MSE + backward + gradient all-reduce + Adam on one
synthetic batch of 64 windows per GPU.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs per sample, fwd+bwd: conv/convT/linear MACs x2, bwd = 2x fwd
# minus the never-needed data gradient of the first layer.
F_TRAIN_PER_SAMPLE = 2.996422e9
METRIC = "train samples/s UNet-ConvLSTM 48x72 seq6 @1/2/4/8 B200; weighted-RMSE parity"
B, T, C, H, W, BASE = 64, 6, 7, 48, 72, 16
N_ROTATE = 8   # resident input batches rotated through (8 x 37 MB of inputs > 126 MB L2)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_cpu_step_time(batch: int, steps: int, warmup: int, threads: int):
    """fwd + MSE + bwd + Adam of the oracle restatement (torch CPU fp32) on `batch` windows."""
    import torch
    from oracle import model_oracle as O
    torch.set_num_threads(threads)
    sd = O.synth_state_dict(O.attunet_spec(C, 2, BASE), 42)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    used = [v for k, v in params.items() if not k.startswith("post_conv")]
    ms = [torch.zeros_like(p) for p in used]
    vs = [torch.zeros_like(p) for p in used]
    x, y, _ = O.synth_attunet_batch(batch, T, H, W, 43)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in used:
            p.grad = None
        loss = O.mse_loss(O.attunet_convlstm(x, params), y)
        loss.backward()
        with torch.no_grad():
            for p, m, v in zip(used, ms, vs):
                O.adam_step(p, p.grad, m, v, it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


def run_reference(args):
    """CPU reference arm: the reference's algorithm (oracle port; /root/reference is Python and does
    not exist on the GPU box) on the host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_b = 16
    times = oracle_cpu_step_time(sample_b, args.steps, max(args.warmup, 1), cores)
    total = sum(times)
    val = sample_b * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "unet_convlstm_attention base16 in_ch7 48x72 seq6, train step (fwd+MSE+bwd+Adam)",
                   "per_gpu_batch": B, "sample": f"{sample_b} windows per step (bounded CPU sample of the 64-window batch)"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} steps x {sample_b} windows, torch CPU fp32, {cores} threads"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def kernel_trace(step_fn, n_steps=2, dump=None):
    """Eager steps with per-C-ABI-call CUDA events on the launching stream.  The device is first parked on
    a ~50 ms spin kernel so the host queues the whole step ahead: the kernels then run back to back and the
    event pairs measure kernel durations, not the host's launch latency.
    -> list of (ms, name, args) of the LAST step, and the per-step totals by kernel name."""
    import torch
    from pcm_b200._lib import lib
    L = lib()
    torch.cuda.synchronize()
    last = []
    agg = {}
    for it in range(n_steps):
        torch.cuda._sleep(100_000_000)
        L.trace = []
        step_fn()
        torch.cuda.synchronize()
        tr, L.trace = L.trace, None
        last = [(e0.elapsed_time(e1), name, a) for name, a, e0, e1 in tr]
        for ms, name, a in last:
            rec = agg.setdefault(name, [0.0, 0])
            rec[0] += ms / n_steps
            rec[1] += 1.0 / n_steps
    if dump:
        from pcm_b200.costmodel import algo_cost, shape_key
        with open(dump, "w") as f:
            for ms, name, a in last:
                fl, by, _ = algo_cost(name, a)
                f.write(f"{ms * 1e3:9.2f} us  {fl / max(ms, 1e-6) / 1e9:8.2f} TF/s {by / max(ms, 1e-6) / 1e6:8.1f} GB/s  "
                        f"{shape_key(name, a)}\n")
    return last, agg


def roofline_groups(calls, peaks):
    """Group the launches of one step by (kernel, shape); for each group: algorithmic FLOPs and bytes per
    launch (costmodel.py), mean launch duration, the binding roof (whichever of FLOPs/peak_TF and
    bytes/peak_BW is the larger time) and the achieved fraction of it."""
    from pcm_b200.costmodel import algo_cost, shape_key
    groups = {}
    for ms, name, a in calls:
        fl, by, _ = algo_cost(name, a)
        g = groups.setdefault(shape_key(name, a), {"kernel": name, "launches": 0, "ms": 0.0, "flops": fl, "bytes": by})
        g["launches"] += 1
        g["ms"] += ms
    tf_peak, bw_peak = peaks["bf16_tflops"], peaks["hbm_gbs"]
    out = []
    for key, g in groups.items():
        avg_s = g["ms"] / g["launches"] / 1e3
        t_tensor = g["flops"] / (tf_peak * 1e12)
        t_hbm = g["bytes"] / (bw_peak * 1e9)
        bound = "tensor" if t_tensor > t_hbm else "hbm"
        if bound == "tensor":
            ach, peak, unit = g["flops"] / avg_s / 1e12, tf_peak, "TFLOP/s"
        else:
            ach, peak, unit = g["bytes"] / avg_s / 1e9, bw_peak, "GB/s"
        out.append({"launch": key, "kernel": g["kernel"], "launches_per_step": g["launches"],
                    "ms_per_step": round(g["ms"], 5), "avg_us": round(avg_s * 1e6, 2), "bound": bound,
                    "achieved": round(ach, 2), "peak": peak, "unit": unit, "frac": round(ach / peak, 4),
                    "algo_flops_per_launch": g["flops"], "algo_bytes_per_launch": g["bytes"]})
    out.sort(key=lambda r: -r["ms_per_step"])
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pcm_b200
    from pcm_b200._lib import lib
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    from pcm_b200.trainer import TrainStep
    from oracle import model_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the pcm_b200 kernels have no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the ONE JSON line: any NCCL_DEBUG level (even WARN) makes NCCL print its version banner
        if "PCM_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["PCM_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_src = load_peaks()

    torch.manual_seed(42)                                   # configs/main_config.yaml:11
    model = AttUNetConvLSTM(in_ch=C, out_ch=2, base=BASE, seq_len=T).to(dev)
    step = TrainStep(model, (B, T, C, H, W), (B, 2, H, W), lr=5e-4, weight_decay=0.0, use_graph=not args.no_graph)

    # synthetic data of the named shape, distinct per rank (seed 42 + rank), rotated so inputs are not L2 resident
    xs_h, ys_h = [], []
    for i in range(N_ROTATE):
        x, y, _ = O.synth_attunet_batch(B, T, H, W, seed=42 + rank + 1000 * i)
        xs_h.append(x.pin_memory()); ys_h.append(y.pin_memory())
    xs_d = [x.to(dev) for x in xs_h]
    ys_d = [y.to(dev) for y in ys_h]
    step.load_batch(xs_d[0], ys_d[0])
    step.warmup_and_capture(warmup=3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    sampler = ClockSampler(local)       # started before the warm-up steps (nvidia-smi needs ~100 ms to deliver its first
    sampler.start()                     # sample); it keeps sampling through the timed region
    # W untimed warm-up steps (at least 10 under NCCL: the first replays of a graph holding collectives are slow)
    for i in range(max(args.warmup, 10) if world > 1 else args.warmup):
        step.step(xs_d[i % N_ROTATE], ys_d[i % N_ROTATE])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step.step(xs_d[i % N_ROTATE], ys_d[i % N_ROTATE])
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    final_loss = float(step.loss.item())

    # ---- end to end: pinned host inputs, H2D inside the timed region, loss read back every step ------
    # Every timed step trains on one batch, copies ONE batch (the next one) from pinned host memory and reads the
    # loss back to the host; the copy runs on a second stream so that it overlaps the step (TrainStep.step_prefetch),
    # as a pinned-memory DataLoader does for the reference loop.  `e2e_serial` is the same without the overlap.
    loss_h = torch.zeros((), dtype=torch.float32).pin_memory()
    for i in range(max(3, args.warmup)):
        step.step(xs_h[i % N_ROTATE], ys_h[i % N_ROTATE])
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        l = step.step(xs_h[i % N_ROTATE], ys_h[i % N_ROTATE])
        loss_h.copy_(l, non_blocking=True)
        torch.cuda.current_stream().synchronize()              # the user reads the loss every step
    e1.record()
    barrier()
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_serial_ms_total = float(ms2.item())

    step.load_batch(xs_h[0], ys_h[0])
    for i in range(max(3, args.warmup)):
        step.step_prefetch(xs_h[(i + 1) % N_ROTATE], ys_h[(i + 1) % N_ROTATE])
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        l = step.step_prefetch(xs_h[(i + 1) % N_ROTATE], ys_h[(i + 1) % N_ROTATE])
        loss_h.copy_(l, non_blocking=True)
        torch.cuda.current_stream().synchronize()              # the user reads the loss every step
    e1.record()
    barrier()
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms_total = float(ms2.item())

    line = None
    if rank == 0:
        value = world * B * args.steps / (ms_total / 1e3)
        e2e = world * B * args.steps / (e2e_ms_total / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "unet_convlstm_attention base16 in_ch7 (5 forcings + sin/cos month) 48x72 seq6, "
                                   "train step = fwd + MSE + bwd + grad all-reduce + Adam (BASELINE configs[2]; "
                                   "same per-GPU batch on every rank = configs[3])",
                       "per_gpu_batch": B, "global_batch": world * B, "seq_len": T, "parallelism": f"dp{world}",
                       "cuda_graph": step.graph is not None,
                       "streams": "2 (weight gradients / skips forked inside the graph)" if step.side is not None else "1",
                       "l2": f"inputs rotate over {N_ROTATE} resident batches ({N_ROTATE * B * T * C * H * W * 4 / 1e6:.0f} MB) "
                             "> 126 MB L2; per-step activation traffic is several hundred MB"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "samples/s", "ms_per_step": e2e_ms_total / args.steps,
                    "h2d_bytes_per_step": (B * T * C * H * W + B * 2 * H * W) * 4, "d2h_bytes_per_step": 4,
                    "how": "TrainStep.step_prefetch: H2D of the next batch on a copy stream overlaps the step; loss read "
                           "back and stream synchronised every step",
                    "serial_value": world * B * args.steps / (e2e_serial_ms_total / 1e3),
                    "serial_ms_per_step": e2e_serial_ms_total / args.steps},
            "gpu_launches": step.launches_per_step * (3 * args.steps + args.warmup + 2 * max(3, args.warmup)),
            "launches_per_step": step.launches_per_step,
            "final_loss": final_loss,
            "step_roofline": {"bound": "tensor", "achieved": value / world * F_TRAIN_PER_SAMPLE / 1e12,
                              "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                              "frac": value / world * F_TRAIN_PER_SAMPLE / 1e12 / peaks["bf16_tflops_sustained"],
                              "note": "whole step, algorithmic FLOPs (2.996 GFLOP/sample) over wall time"},
        }

    # ---- dominant kernel roofline: eager pass with per-call CUDA events on the launching stream -----------
    if rank == 0 and world == 1:
        calls, agg = kernel_trace(step._step_impl, n_steps=2, dump=args.trace_file)
        tot = sum(v[0] for v in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1][0])
        line["kernel_breakdown_ms"] = {k: round(v[0], 4) for k, v in top[:12]}
        line["kernel_time_ms_eager_step"] = round(tot, 4)
        groups = roofline_groups(calls, peaks)
        g0 = groups[0]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from `ncu --set full`
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(g0["launch"])
        line["roofline"] = {"bound": g0["bound"], "kernel": g0["kernel"], "launch": g0["launch"],
                            "achieved": g0["achieved"], "peak": g0["peak"], "unit": g0["unit"], "frac": g0["frac"],
                            "traffic": traffic, "avg_launch_us": g0["avg_us"],
                            "launches_per_step": g0["launches_per_step"], "share_of_step": round(g0["ms_per_step"] / tot, 4),
                            "algo_bytes_per_launch": g0["algo_bytes_per_launch"],
                            "algo_flops_per_launch": g0["algo_flops_per_launch"],
                            "peak_source": peak_src + ", burst figures (kernel timed alone between events)",
                            "how": "dominant (kernel, shape) group of one eager step; CUDA events around each launch "
                                   "on the launching stream with the host queued ahead of the device"}
        line["roofline_top"] = [{k: r[k] for k in ("launch", "launches_per_step", "ms_per_step", "avg_us", "bound",
                                                   "achieved", "unit", "frac")} for r in groups[:10]]
        # CPU baseline beside it: the oracle port on the host cores, bounded sample
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            times = oracle_cpu_step_time(16, 3, 1, cores)
            line["cpu_baseline"] = {"value": 16 * len(times) / sum(times), "unit": "samples/s", "cores": cores,
                                    "kind": "port", "sample": f"{len(times)} steps x 16 windows of the same workload, "
                                                              f"oracle (torch CPU fp32), {cores} threads"}
    if rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # Tearing a NCCL communicator down while a captured graph still references its kernels can hang
        # (seen on 2 x B200): drop the graph, drain the device, meet at a barrier and leave without the
        # collective destructor.
        step.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace-file", default=None, help="write the per-call timing of one eager step here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

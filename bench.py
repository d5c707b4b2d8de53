#!/usr/bin/env python
"""Benchmark of the DYffusion sampling hot path (BASELINE.json metric):

    grid-cells x diffusion-steps / sec, Navier-Stokes 221x42, horizon 16, 64 rows per B200

    python bench.py --gpus N --steps K --warmup W            # our engine (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation (oracle port)

A "step" is one `sample()` call over one batch of synthetic Navier-Stokes rows: 16 forecaster + 44 interpolator
UNet forwards (cold sampling + refinement, interpolator dropout on).  Prints ONE JSON line on rank 0.

Timed region: K calls of `sample()` with device-resident inputs, one CUDA-event pair around the whole region (max over
ranks), barrier + synchronize on both sides; inside it only the launches of the dominant kernel class carry their own
event pairs (`roofline`: algorithmic FLOPs / measured device time of that kernel, live).  Before it: W >= 3 warm-up steps
and one extra untimed step with events around EVERY launch (`kernel_ms_per_step`).  After it: the same K steps end to end
through `predict_forward` with pinned host buffers (`e2e`), then (N=1) the CPU baseline.  The working set of one network
forward (>= 5 GB of activations) is far beyond the 126 MB L2, so no explicit flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, HORIZON, ROWS = 221, 42, 16, 64
METRIC = "grid-cells*diffusion-steps/sec, Navier-Stokes 221x42 h=16"
UNIT = "cell-steps/s"
# algorithmic (reference-dense, 2*MAC) FLOPs of one network forward per row -- SURVEY.md 6 / BASELINE.md 2
GF_FORECASTER, GF_INTERPOLATOR = 48.161, 48.186


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p["bf16_tflops_sustained"], tflops_burst=p.get("bf16_tflops"),
                    source="measured (MEASURED_PEAKS.json, sustained: the kernel is timed inside a long, power-capped step)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        busy = [c for c in sm if c > 0.5 * max(sm)] or sm
        return dict(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def workload(rows):
    """`config.workload` of both arms (the reference arm times a bounded sample of it, named in its `cpu_baseline.sample`)."""
    return (f"NS 221x42 DYffusion sampling h=16, {rows} rows/GPU (BASELINE configs[1]): 16 forecaster + 44 interpolator "
            "unet_simple forwards, cold sampling + refinement, interpolator dropout 0.15")


def synth_inputs(rows, seed):
    from oracle.synth import synth_tensor
    ic = synth_tensor(f"bench.ic{seed}", (rows, 3, H, W))
    static = synth_tensor(f"bench.static{seed}", (rows, 2, H, W), kind="mask")
    return ic, static


def cpu_reference_once(rows):
    """The reference's CPU implementation of the path (oracle port of the unmodified PyTorch code, fp32, all host
    threads): one NS h=16 sample() -- 16 F + 44 I forwards, refinement on, interpolator dropout on."""
    import torch
    from oracle import configs as C
    from oracle import dyffusion_oracle as O
    from oracle.dyffusion_oracle import torch_dropout
    from oracle.synth import synth_state_dict
    from tests import helpers as Hh

    shapes = Hh.golden_json("state_shapes.json")
    sdF, sdI = synth_state_dict(shapes["ns_F"], seed=3), synth_state_dict(shapes["ns_I"], seed=2)
    dk = C.diffusion_kwargs("ns")
    sched = Hh.oracle_schedule(dk)
    ic, static = synth_inputs(rows, 0)
    F = Hh.oracle_net("ns", "F", sdF)
    I = Hh.oracle_net("ns", "I", sdI, drop=torch_dropout)
    t0 = time.perf_counter()
    with torch.no_grad():
        out = O.sample_loop(F, I, sched, ic, static, num_input_channels=3,
                            forward_conditioning=dk["forward_conditioning"],
                            refine_intermediate_predictions=dk["refine_intermediate_predictions"])
    dt = time.perf_counter() - t0
    assert len(out) == HORIZON
    return dt


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = 1
    for _ in range(min(args.warmup, 1)):
        cpu_reference_once(rows)
    times = [cpu_reference_once(rows) for _ in range(max(1, args.steps))]
    dt = sum(times) / len(times)
    v = rows * H * W * HORIZON / dt
    sample = f"rows={rows} of the same NS h=16 sample() (60 UNet forwards) per step, {len(times)} step(s)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.rows), "rows_per_gpu": args.rows,
                   "sample": "the reference's CPU path is timed on rows=1 per step of this workload (same networks, schedule, "
                             "refinement, dropout); throughput is per row, so the bounded sample does not bias it"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_engine(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import dyffusion_b200.engine as E
    from tests.gpu_helpers import build_dyffusion

    rows = args.rows
    dyf = build_dyffusion("ns")  # shipped NS config: h=16, cold sampling, refine, fcond none, dropout 0.15
    n_steps_sched = len(dyf.sampling_schedule)
    ic_h, st_h = synth_inputs(rows, rank)
    ic_pin, st_pin = ic_h.pin_memory(), st_h.pin_memory()
    ic, st = ic_pin.to(dev), st_pin.to(dev)
    gathered = torch.empty((world, HORIZON, rows, 3, H, W), device=dev) if world > 1 else None

    def step_device():
        out = dyf.sample(ic, static_condition=st)
        preds = out["t1_preds"]._base if out["t1_preds"]._base is not None else torch.stack(list(out.values()))
        if world > 1:  # the path's only exchange: one all-gather of the per-rank forecasts
            dist.all_gather_into_tensor(gathered, preds.contiguous())
        return preds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step_device()
        barrier()
        # untimed profiling pass: every launch bracketed by CUDA events -> kernel-class breakdown of one step and the
        # dominant kernel class (the one the roofline is reported for)
        E.profile_filter(None)
        E.profile_enable(True)
        step_device()
        barrier()
        breakdown = E.profile_read()
        E.profile_enable(False)
        conv = {k: breakdown[k] for k in ("conv_mma", "conv_umma", "conv_up")}
        dom = max(conv, key=lambda k: conv[k]["ms"])
        barrier()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        # timed region: only the dominant kernel's launches carry events (measured live, on the launching stream);
        # no event records between the other launches
        E.profile_filter(dom)
        E.profile_enable(True)
        launches0 = E.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step_device()
        ev1.record()
        barrier()
        launches = E.launch_count() - launches0
        prof = E.profile_read()
        E.profile_enable(False)
        E.profile_filter(None)
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())

        # ---- end to end through the public API with host buffers: H2D of the inputs, sample(), D2H of the forecasts
        host_out = torch.empty((HORIZON, rows, 3, H, W), dtype=torch.float32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ic_d = ic_pin.to(dev, non_blocking=True)
            st_d = st_pin.to(dev, non_blocking=True)
            out = dyf.predict_forward(ic_d, condition=st_d)
            preds = out["t1_preds"]._base if out["t1_preds"]._base is not None else torch.stack(list(out.values()))
            if world > 1:
                dist.all_gather_into_tensor(gathered, preds.contiguous())
            host_out.copy_(preds, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        pk = peaks()
        units_per_step = world * rows * H * W * n_steps_sched
        value = units_per_step * args.steps / (ms_total / 1e3)
        e2e = units_per_step * args.steps / float(e2e_s.item())
        d = prof[dom]  # the dominant kernel class, timed live inside the timed region
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0
        # DRAM traffic of that kernel from the committed `ncu --set full` capture (profiles/r01_traffic.json): bytes per
        # launch averaged over the launches of one 64-row forward, next to the algorithmic bytes of the same launches
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        total_kernel_ms = sum(v["ms"] for v in breakdown.values())
        flop_step = rows * (16 * GF_FORECASTER + 44 * GF_INTERPOLATOR) * 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload(rows), "rows_per_gpu": rows, "l2": "working set >> L2 (>= 5 GB of activations per network forward)",
                       "operands": "bf16 operands, fp32 accumulate/epilogue, fp32 sampler state"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(ic_pin.numel() * 4 + st_pin.numel() * 4),
                    "d2h_bytes_per_step": int(host_out.numel() * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": ach / pk["tflops"], "frac_of_burst_peak": ach / pk["tflops_burst"] if pk.get("tflops_burst") else None,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_detail": traffic, "algorithmic_bytes_per_launch": d["bytes"] / max(1, d["launches"]),
                         "peak_source": pk["source"],
                         "launches": d["launches"], "avg_launch_ms": d["ms"] / max(1, d["launches"]),
                         "share_of_kernel_time": breakdown[dom]["ms"] / total_kernel_ms if total_kernel_ms else None,
                         "whole_step_tflops": flop_step * args.steps / (ms_total / 1e3) / 1e12},
            "kernel_ms_per_step": {k: v["ms"] for k, v in breakdown.items() if v["launches"]},
            "kernel_ms_note": "one extra untimed step with events around every launch; roofline.* is timed live in the timed region",
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            dt = cpu_reference_once(1)
            line["cpu_baseline"] = {"value": H * W * HORIZON / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "rows=1 of the same NS h=16 sample() (60 UNet forwards), 1 pass"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=int, default=ROWS, help="rows (batch x ensemble members) per GPU")
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the DYffusion sampling hot path (BASELINE.json metric: grid-cells x diffusion-steps / sec).

    python bench.py --gpus N --steps K --warmup W                  # default: Navier-Stokes, weak scaling (configs[1])
    python bench.py --config {ns,sst,spring} [--scaling strong]    # the three shipped configurations (configs[1..3])
    python bench.py --impl reference ...                           # the reference's CPU implementation (oracle port)

A "step" is one `sample()` call over one batch of synthetic rows (row = batch element x ensemble member):
  ns      Navier-Stokes 221x42, h=16:  16 forecaster + 44 interpolator `unet_simple` forwards (cold sampling + refinement)
  sst     SST 60x60, h=7, k=25 (32 steps, "data+noise"):  32 + 61 `Unet` forwards
  spring  spring-mesh 10x10, h=134 (refinement):  134 + 398 `SimpleConvNet` forwards
Weak scaling (default): every GPU owns `rows` rows (ns 64, sst 304, spring 800), the job grows with N.  Strong scaling
(`--scaling strong`): ONE job of that many rows is split over the N ranks.  Both go through the product's multi-GPU API,
`dyffusion_b200.distributed.sample_sharded` (row shards, no data-path collective, one all-gather of the forecasts).
Prints ONE JSON line on rank 0.

Timed region: K calls of `sample()` with device-resident inputs, one CUDA-event pair around the whole region (max over
ranks), barrier + synchronize on both sides; the sampler replays its CUDA graph (the product default), so kernels are not
individually bracketed there.  Before it: W >= 3 warm-up steps and one extra untimed step on plain launches with events
around EVERY launch (`kernel_ms_per_step`, and the dominant kernel class of `roofline`: algorithmic FLOPs / its device
time).  After it: the same K steps end to end through `predict_forward` / `sample_sharded` with pinned host buffers
(`e2e`), then (N=1) the CPU baseline and the eager-PyTorch-on-GPU comparator.  The working set of one network forward is
far beyond the 126 MB L2 for ns / sst (>= 5 GB / >= 1 GB of activations), so no explicit flush is needed; for spring-mesh
(35 MB per forward) a 256 MB buffer is rewritten between steps.
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

UNIT = "cell-steps/s"
# algorithmic (reference-dense, 2*MAC) GFLOPs of ONE row's whole trajectory -- SURVEY.md 8d
CONFIGS = {
    "ns": dict(rows=64, gf_row=16 * 48.161 + 44 * 48.186, calls="16 forecaster + 44 interpolator unet_simple forwards",
               metric="grid-cells*diffusion-steps/sec, Navier-Stokes 221x42 h=16", baseline="configs[1]",
               desc="NS 221x42 DYffusion sampling h=16, cold sampling + refinement, interpolator dropout 0.15"),
    "sst": dict(rows=304, gf_row=93 * 10.712, calls="32 forecaster + 61 interpolator Unet forwards",
                metric="grid-cells*diffusion-steps/sec, SST 60x60 h=7 k=25", baseline="configs[2]",
                desc="SST 60x60 DYffusion sampling h=7, k=25 (32 steps), data+noise conditioning, 50-member ensemble x batch 6 "
                     "padded to 304 rows"),
    "spring": dict(rows=800, gf_row=532 * 0.078, calls="134 forecaster + 398 interpolator SimpleConvNet forwards",
                   metric="grid-cells*diffusion-steps/sec, spring-mesh 10x10 h=134", baseline="configs[3]",
                   desc="spring-mesh 10x10 DYffusion sampling h=134, cold sampling + refinement, 50 members x batch 16 = 800 rows"),
}


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


def workload(cfg, rows, scaling, world):
    c = CONFIGS[cfg]
    split = f"{rows} rows per GPU" if scaling == "weak" else f"one job of {rows} rows split over {world} GPU(s)"
    return f"{c['desc']} (BASELINE {c['baseline']}): {c['calls']}; {split}"


def synth_inputs(preset, rows, seed):
    """Synthetic inputs of the named shape: initial condition ~ N(0,1), static condition = Bernoulli(0.1) mask."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(1000 + seed)
    d = preset["dataset"]
    h, w = d["spatial_shape"]
    ic = torch.randn((rows, d["channels"] * d["window"], h, w), generator=g)
    static = (torch.rand((rows, d["static_channels"], h, w), generator=g) < 0.1).float() if d["static_channels"] else None
    return ic, static


# --------------------------------------------------------------------------------------------- CPU / eager baselines
def oracle_sampler(cfg, device="cpu"):
    """The reference's implementation of the path as restated by the oracle (plain PyTorch fp32, the unmodified model code's
    arithmetic): returns sample(ic, static).  Test infrastructure used here ONLY as the timed CPU / eager-GPU baseline."""
    import torch
    from oracle import configs as C
    from oracle import dyffusion_oracle as O
    from oracle.dyffusion_oracle import torch_dropout
    from oracle.synth import synth_state_dict
    from tests import helpers as Hh

    shapes = Hh.golden_json("state_shapes.json")
    sdF = {k: v.to(device) for k, v in synth_state_dict(shapes[f"{cfg}_F"], seed=3).items()}
    sdI = {k: v.to(device) for k, v in synth_state_dict(shapes[f"{cfg}_I"], seed=2).items()}
    dk = C.diffusion_kwargs(cfg)
    sched = Hh.oracle_schedule(dk)
    F = Hh.oracle_net(cfg, "F", sdF)
    I = Hh.oracle_net(cfg, "I", sdI, drop=torch_dropout)

    def sample(ic, static):
        with torch.no_grad():
            return O.sample_loop(F, I, sched, ic, static, num_input_channels=C.DATASETS[cfg]["channels"],
                                 forward_conditioning=dk["forward_conditioning"],
                                 refine_intermediate_predictions=dk["refine_intermediate_predictions"])
    return sample, len(sched.sampling_schedule)


CPU_SAMPLE_ROWS = {"ns": 1, "sst": 2, "spring": 16}


def cpu_reference_once(cfg, preset):
    import torch
    rows = CPU_SAMPLE_ROWS[cfg]
    sample, n_steps = oracle_sampler(cfg)
    ic, static = synth_inputs(preset, rows, 0)
    t0 = time.perf_counter()
    out = sample(ic, static)
    dt = time.perf_counter() - t0
    h, w = preset["dataset"]["spatial_shape"]
    assert len(out) == preset["dataset"]["horizon"]
    return rows * h * w * n_steps / dt, dt, rows


def gpu_eager_once(cfg, preset, rows, dev):
    """The same sample() through stock PyTorch ops on the B200 (cuDNN / cuBLAS, TF32 allowed): the strong comparator."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.set_float32_matmul_precision("high")
    sample, n_steps = oracle_sampler(cfg, device=dev)
    ic, static = synth_inputs(preset, rows, 0)
    ic = ic.to(dev)
    static = None if static is None else static.to(dev)
    prev = torch.get_default_device()
    torch.set_default_device(dev)  # the oracle builds its time vectors with torch.full(...)
    try:
        small = min(rows, 8)
        sample(ic[:small], None if static is None else static[:small])  # warm-up (cuDNN heuristics, allocator)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sample(ic, static)
        e1.record()
        torch.cuda.synchronize()
    finally:
        torch.set_default_device(prev)
    ms = e0.elapsed_time(e1)
    h, w = preset["dataset"]["spatial_shape"]
    return dict(value=rows * h * w * n_steps / (ms / 1e3), unit=UNIT, ms_per_step=ms, rows=rows,
                how="oracle port (the reference's model arithmetic) as eager torch ops on the same GPU, fp32 tensors with TF32 "
                    "matmul/conv allowed, cuDNN; 1 warm-up at 8 rows, 1 timed sample()")


def run_reference(args):
    import torch
    from dyffusion_b200 import presets as P
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    preset = P.load_preset(cfg)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_once(cfg, preset)
    res = [cpu_reference_once(cfg, preset) for _ in range(max(1, args.steps))]
    v = sum(r[0] for r in res) / len(res)
    dt = sum(r[1] for r in res) / len(res)
    rows = res[0][2]
    sample = f"rows={rows} of the same sample() ({CONFIGS[cfg]['calls']}) per step, {len(res)} step(s)"
    print(json.dumps({
        "impl": "reference", "metric": CONFIGS[cfg]["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(res), "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(cfg, args.rows or CONFIGS[cfg]["rows"], args.scaling, args.gpus),
                   "rows_per_gpu": args.rows or CONFIGS[cfg]["rows"],
                   "sample": f"the reference's CPU path is timed on rows={rows} per step of this workload (same networks, "
                             "schedule, refinement, dropout); throughput is per row, so the bounded sample does not bias it"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------- the engine arm
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
    from dyffusion_b200 import presets as P
    from dyffusion_b200.distributed import sample_sharded, shard_bounds

    cfg = args.config
    preset = P.load_preset(cfg)
    rows_cfg = args.rows or CONFIGS[cfg]["rows"]
    job_rows = rows_cfg * world if args.scaling == "weak" else rows_cfg  # rows of the whole job
    dyf = P.build_dyffusion(cfg, device=dev, seed=0)  # shipped configuration, reference default init (synthetic weights)
    n_sched = len(dyf.sampling_schedule)
    h, w = preset["dataset"]["spatial_shape"]
    horizon = preset["dataset"]["horizon"]
    ic_h, st_h = synth_inputs(preset, job_rows, 0)  # every rank holds the whole job's inputs (a replicated eval batch)
    ic_pin = ic_h.pin_memory()
    st_pin = None if st_h is None else st_h.pin_memory()
    ic = ic_pin.to(dev)
    st = None if st_pin is None else st_pin.to(dev)
    my_rows = shard_bounds(job_rows, world)[rank]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if cfg == "spring" else None

    def step_device(ic_d=ic, st_d=st):
        if flush is not None:
            flush.fill_(1)  # spring-mesh: the per-forward working set fits L2 -> evict between steps
        return sample_sharded(dyf, ic_d, st_d)  # the product's multi-GPU API (world 1: plain sample())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = step_device()
        assert len(out) == horizon and tuple(out["t1_preds"].shape) == (job_rows, preset["dataset"]["channels"], h, w)
        barrier()
        # untimed profiling pass on plain launches: every launch bracketed by CUDA events -> kernel-class breakdown of one
        # step and the dominant class (the one the roofline is reported for)
        E.profile_filter(None)
        E.profile_enable(True)
        step_device()
        barrier()
        breakdown = E.profile_read()
        E.profile_enable(False)
        dom = max(breakdown, key=lambda k: breakdown[k]["ms"] if breakdown[k]["flops"] > 0 else -1.0)
        barrier()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        launches0 = E.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step_device()
        ev1.record()
        barrier()
        launches = E.launch_count() - launches0
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())

        # ---- the dominant kernel class timed LIVE: the same K steps on plain launches with CUDA events around the launches
        # of that class only (on the launching stream; the graph replay above cannot carry per-kernel events)
        barrier()
        E.profile_filter(dom)
        E.profile_enable(True)
        evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evp0.record()
        for _ in range(args.steps):
            step_device()
        evp1.record()
        barrier()
        prof = E.profile_read()
        E.profile_enable(False)
        E.profile_filter(None)
        ms_plain = evp0.elapsed_time(evp1)

        # ---- end to end through the public API with host buffers.  Per step every rank copies ITS shard of the job's inputs
        # from pinned host memory, runs sample_sharded, and reads ITS rows of the (gathered) forecasts back to pinned host memory
        b0, b1 = my_rows
        host_out = torch.empty((horizon, max(b1 - b0, 1), preset["dataset"]["channels"], h, w), dtype=torch.float32).pin_memory()
        ic_e2e = torch.zeros_like(ic)
        st_e2e = None if st is None else torch.zeros_like(st)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ic_e2e[b0:b1].copy_(ic_pin[b0:b1], non_blocking=True)
            if st_e2e is not None:
                st_e2e[b0:b1].copy_(st_pin[b0:b1], non_blocking=True)
            out = step_device(ic_e2e, st_e2e)
            base = out["t1_preds"]._base
            if base is not None and base.dim() == 5 and base.shape[0] == horizon:  # all forecasts are views of one tensor
                host_out[:, :b1 - b0].copy_(base[:, b0:b1], non_blocking=True)
            else:
                for i in range(horizon):
                    host_out[i, :b1 - b0].copy_(out[f"t{i + 1}_preds"][b0:b1], non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        pk = peaks()
        units_per_step = job_rows * h * w * n_sched
        value = units_per_step * args.steps / (ms_total / 1e3)
        e2e = units_per_step * args.steps / float(e2e_s.item())
        d = prof[dom]  # the dominant kernel class, timed live
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(cfg, {}).get(dom)
        total_kernel_ms = sum(v["ms"] for v in breakdown.values())
        flop_step = job_rows * CONFIGS[cfg]["gf_row"] * 1e9
        step_tflops = flop_step * args.steps / (ms_total / 1e3) / 1e12
        row_in = ic_pin[0].numel() * 4 + (0 if st_pin is None else st_pin[0].numel() * 4)
        h2d = row_in * job_rows                          # summed over ranks: every row's inputs cross PCIe once
        d2h = horizon * job_rows * preset["dataset"]["channels"] * h * w * 4
        line = {
            "metric": CONFIGS[cfg]["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": E.act_dtype(), "data": "synthetic",
            "config": {"workload": workload(cfg, rows_cfg, args.scaling, world), "name": cfg, "job_rows": job_rows,
                       "rows_this_rank": my_rows[1] - my_rows[0],
                       "l2": ("256 MB buffer rewritten between steps" if flush is not None else
                              "working set >> L2 (>= 1 GB of activations per network forward)"),
                       "operands": f"{E.act_dtype()} operands and activations, fp32 accumulate / epilogues / sampler state",
                       "multi_gpu_api": "dyffusion_b200.distributed.sample_sharded", "cuda_graph": os.environ.get("DYF_CUDA_GRAPH", "1") != "0"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "bytes_note": "whole job (sum over ranks); every rank moves its own row shard"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": ach / pk["tflops"], "frac_of_burst_peak": ach / pk["tflops_burst"] if pk.get("tflops_burst") else None,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_detail": traffic,
                         "algorithmic_bytes_per_launch": d["bytes"] / max(1, d["launches"]), "peak_source": pk["source"],
                         "launches": d["launches"], "avg_launch_ms": d["ms"] / max(1, d["launches"]),
                         "share_of_kernel_time": (d["ms"] / args.steps) / total_kernel_ms if total_kernel_ms else None,
                         "timed": f"live: {args.steps} plain-launch steps right after the timed region with CUDA events around the "
                                  "launches of this class only (the timed region itself replays the sampler's CUDA graph)",
                         "ms_per_step_during_measurement": ms_plain / args.steps,
                         "whole_step_tflops": step_tflops, "whole_step_frac": step_tflops / world / pk["tflops"]},
            "kernel_ms_per_step": {k: v["ms"] for k, v in breakdown.items() if v["launches"]},
            "kernel_launches_per_step": {k: v["launches"] for k, v in breakdown.items() if v["launches"]},
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            v, dt, r = cpu_reference_once(cfg, preset)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"rows={r} of the same sample() ({CONFIGS[cfg]['calls']}), 1 pass, {dt:.1f} s"}
            try:
                line["gpu_eager_baseline"] = gpu_eager_once(cfg, preset, min(job_rows, args.eager_rows or job_rows), dev)
            except Exception as e:  # an out-of-memory eager run must not lose the line
                line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="ns", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (weak) / rows of the job (strong); 0 = the configuration's")
    ap.add_argument("--eager-rows", type=int, default=0, help="rows of the eager-GPU comparator (0 = the job's)")
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()

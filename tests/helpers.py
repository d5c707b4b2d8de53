"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import json
import os

import torch

from oracle import configs as C
from oracle import dyffusion_oracle as O
from oracle.synth import SiteDropout, synth_state_dict, synth_tensor

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_pt(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu")


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def model_kwargs(dataset, role):
    kw = dict(C.MODELS[dataset]["kwargs"])
    if role == "I":
        kw.update(C.INTERPOLATOR_OVERRIDES[dataset])
    return kw


def oracle_kwargs(dataset, role):
    """kwargs of the oracle backbone function for a dataset/role."""
    kw = model_kwargs(dataset, role)
    arch = C.MODELS[dataset]["arch"]
    if arch == "unet_simple":
        keep = ("dim", "upsample_dims", "outer_sample_mode", "dropout", "input_dropout")
    elif arch == "unet_resnet":
        keep = ("dim", "dim_mults", "block_dropout", "block_dropout1", "attn_dropout", "input_dropout",
                "keep_spatial_dims", "init_padding", "init_stride")
    else:
        keep = ("dim", "kernel_sizes", "residual", "dropout")
    out = {k: kw[k] for k in keep if k in kw}
    if arch == "unet_resnet":
        out["groups"] = kw["resnet_block_groups"]
    return arch, out


def forward_inputs(dataset, role, rows=2):
    fcond = C.DIFFUSION[dataset]["forward_conditioning"]
    tag = f"{dataset}_{role}"
    cin, ccond, _ = C.channels(dataset, role, fcond)
    H, W = C.DATASETS[dataset]["spatial"]
    st = C.DATASETS[dataset]["static"]
    x = synth_tensor(f"{tag}.x", (rows, cin, H, W))
    cond = None
    if ccond > 0:
        parts = []
        if ccond - st > 0:
            parts.append(synth_tensor(f"{tag}.c", (rows, ccond - st, H, W)))
        if st > 0:
            parts.append(synth_tensor(f"{tag}.s", (rows, st, H, W), kind="mask"))
        cond = torch.cat(parts, dim=1)
    return x, cond


def oracle_net(dataset, role, sd, drop=None):
    arch, kw = oracle_kwargs(dataset, role)
    fn = O.BACKBONES[arch]
    return lambda x, t, c: fn(sd, x, t, c, drop=drop, **kw)


def sampler_case_inputs(name, dataset, rows):
    d = C.DATASETS[dataset]
    H, W = d["spatial"]
    ic = synth_tensor(f"{name}.ic", (rows, d["channels"], H, W))
    static = synth_tensor(f"{name}.static", (rows, d["static"], H, W), kind="mask") if d["static"] else None
    return ic, static


def oracle_schedule(dk):
    return O.Schedule(dk["timesteps"], dk["schedule"], dk["additional_interpolation_steps"],
                      dk["additional_interpolation_steps_factor"], dk["interpolate_before_t1"], dk["sampling_schedule"])


# ---------------------------------------------------------------- autoregressive rollout cases
ROLLOUT_CASES = {
    # name: dataset, horizon, ensemble members, batch, autoregressive steps, physical system of the boundary conditions
    "rollout_spring": dict(dataset="spring", horizon=4, members=3, batch=2, ar_steps=2, system="spring-mesh"),
    "rollout_ns": dict(dataset="ns", horizon=2, members=2, batch=2, ar_steps=1, system="navier-stokes"),
    "rollout_spring_single": dict(dataset="spring", horizon=3, members=1, batch=3, ar_steps=1, system="spring-mesh"),
}


def rollout_case(name):
    """Synthetic evaluation batch of the reference's layout (`dynamics` (B, T, C, H, W), `condition`, `metadata`) and the
    `t0` / `dt` the datamodule would hand over (physical_systems_benchmark.py:299-303)."""
    c = ROLLOUT_CASES[name]
    d = C.DATASETS[c["dataset"]]
    (Hh, Ww), ch, b = d["spatial"], d["channels"], c["batch"]
    T = 1 + c["horizon"] * (c["ar_steps"] + 1)
    batch = {"dynamics": synth_tensor(f"{name}.dyn", (b, T, ch, Hh, Ww)),
             "condition": synth_tensor(f"{name}.static", (b, d["static"], Hh, Ww), kind="mask")}
    fixed = synth_tensor(f"{name}.fixed", (b, ch, Hh, Ww), kind="mask") > 0
    if c["system"] == "spring-mesh":
        batch["metadata"] = {"fixed_mask": fixed, "features": synth_tensor(f"{name}.features", (b, 5, ch, Hh, Ww))}
        t0, dt = 0.0, 1.0
    else:
        batch["metadata"] = {"fixed_mask": fixed, "vertices": synth_tensor(f"{name}.vert", (b, 2, Hh, Ww)).abs() * 0.2,
                             "in_velocity": 0.5 + synth_tensor(f"{name}.vel", (b, 1)).abs()}
        t0, dt = synth_tensor(f"{name}.t0", (b,)).abs(), torch.full((b,), 0.05)
    return c, batch, t0, dt


def oracle_rollout_sampler(dataset, horizon, seeds=(3, 2)):
    """`sample(ic, static)` over the oracle sampler with the synthetic weights of tests/golden (forecaster seed 3,
    interpolator seed 2), interpolator dropout off."""
    shapes = golden_json("state_shapes.json")
    sdF = synth_state_dict({k: tuple(v) for k, v in shapes[f"{dataset}_F"].items()}, seed=seeds[0])
    sdI = synth_state_dict({k: tuple(v) for k, v in shapes[f"{dataset}_I"].items()}, seed=seeds[1])
    dk = C.diffusion_kwargs(dataset, horizon=horizon)
    F, I, sched = oracle_net(dataset, "F", sdF), oracle_net(dataset, "I", sdI), oracle_schedule(dk)
    return lambda ic, static: O.sample_loop(
        F, I, sched, ic, static, num_input_channels=C.DATASETS[dataset]["channels"],
        forward_conditioning=dk["forward_conditioning"],
        refine_intermediate_predictions=dk["refine_intermediate_predictions"])


# ---------------------------------------------------------------- trajectory store cases
def synth_trajectories(system, lengths, seed=0):
    """Objects shaped like `TrajectoryDataset.__getitem__`'s result (datasets/physical_systems_benchmark.py:66-160)."""
    import types
    ch, hw, st = (3, (221, 42), 2) if system == "navier-stokes" else (4, (10, 10), 1)
    out = []
    for i, T in enumerate(lengths):
        tag = f"traj.{system}.{i}"
        meta = {"name": f"traj_{i:05d}", "num_time_steps": T, "time_step_size": 0.05 * (1 + i % 2)}
        tr = types.SimpleNamespace(
            features=synth_tensor(f"{tag}.features", (T, ch, *hw), seed).numpy(),
            condition=synth_tensor(f"{tag}.cond", (st, *hw), seed, kind="mask").numpy(),
            fixed_mask=(synth_tensor(f"{tag}.fixed", (ch, *hw), seed, kind="mask") > 0).numpy(),
            t=(0.1 * i + meta["time_step_size"] * torch.arange(T, dtype=torch.float32)).numpy(), trajectory_meta=meta, vertices=[])
        if system == "navier-stokes":
            meta["in_velocity"] = 0.5 + 0.25 * i
            tr.vertices = (synth_tensor(f"{tag}.vert", (2, *hw), seed).abs() * 0.2).numpy()
        out.append(tr)
    return out

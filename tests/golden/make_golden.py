"""Generates the committed golden fixtures by running the UNMODIFIED reference (/root/reference, imported through
oracle/ref_shims.py) on synthetic weights/inputs from oracle/synth.py.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (tests/golden/): state_shapes.json, schedule_kat.json, fwd_*.pt, sample_*.pt.
The reference ships no golden vectors of its own for this path; these are "outputs of the
reference itself run here", which is what pins oracle/ and, through it, the CUDA path.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import configs as C  # noqa: E402
from oracle import ref_build, ref_shims  # noqa: E402
from oracle.synth import synth_state_dict, synth_tensor  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLER_CASES = {
    # name: (dataset, diffusion overrides)  -- small horizons so the CPU reference finishes in seconds
    "ns_h4": ("ns", dict(horizon=4)),
    "ns_h4_naive": ("ns", dict(horizon=4, sampling_type="naive", refine_intermediate_predictions=False)),
    "sst_h3_k2": ("sst", dict(horizon=3, additional_interpolation_steps=2)),
    "sst_h3_k4_every2": ("sst", dict(horizon=3, additional_interpolation_steps=4, sampling_schedule="every2nd",
                                      forward_conditioning="data")),
    "spring_h6": ("spring", dict(horizon=6)),
    "spring_h5_coldlast": ("spring", dict(horizon=5, use_cold_sampling_for_last_step=True,
                                          refine_intermediate_predictions=False)),
}


def shapes_of(module) -> dict:
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def load_synth(backbone, seed):
    sd = synth_state_dict(shapes_of(backbone), seed=seed)
    backbone.load_state_dict(sd, strict=True)
    return sd


def make_inputs(dataset, role, fcond, rows, tag):
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


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_shims.install()
    shapes, kat = {}, {}

    # ---------------- single-forward goldens (dropout off and site-hooked dropout) ----------------
    for ds in ("ns", "sst", "spring"):
        fcond = C.DIFFUSION[ds]["forward_conditioning"]
        ipol = ref_build.build_interpolator(ds, horizon=4)
        dyf = ref_build.build_dyffusion(ds, ipol, horizon=4)
        nets = {"I": ipol.model, "F": dyf.model.model}
        for role, net in nets.items():
            tag = f"{ds}_{role}"
            shapes[tag] = shapes_of(net)
            load_synth(net, seed=1)
            net.eval()
            x, cond = make_inputs(ds, role, fcond, rows=2, tag=tag)
            t = torch.tensor([1.0, 2.5]) if role == "I" else torch.tensor([0.0, 3.0])
            with torch.no_grad():
                y = net(x, time=t, condition=cond)
                hook = ref_build.HookedDropout(net, seed=7)
                y_drop = net(x, time=t, condition=cond)
                hook.remove()
            torch.save({"time": t, "y": y, "y_sitedrop": y_drop, "drop_seed": 7, "weight_seed": 1, "rows": 2},
                       os.path.join(OUT, f"fwd_{tag}.pt"))
            print(tag, tuple(y.shape), float(y.abs().mean()), float((y - y_drop).abs().mean()))

    # ---------------- sampler goldens (interpolator dropout disabled -> deterministic) ----------------
    for name, (ds, ov) in SAMPLER_CASES.items():
        ov = dict(ov)
        horizon = ov["horizon"]
        ipol = ref_build.build_interpolator(ds, horizon=horizon)
        dyf = ref_build.build_dyffusion(ds, ipol, enable_interpolator_dropout=False, **ov)
        load_synth(ipol.model, seed=2)
        load_synth(dyf.model.model, seed=3)
        d = C.DATASETS[ds]
        H, W = d["spatial"]
        rows = 1 if ds == "ns" else 2
        ic = synth_tensor(f"{name}.ic", (rows, d["channels"], H, W))
        static = synth_tensor(f"{name}.static", (rows, d["static"], H, W), kind="mask") if d["static"] else None
        calls = {"n": 0, "F": 0, "I": 0}

        def fake_randn_like(t, _c=calls, _name=name):
            _c["n"] += 1
            return synth_tensor(f"{_name}.noise{_c['n'] - 1}", tuple(t.shape))

        hF = dyf.model.model.register_forward_hook(lambda *a, _c=calls: _c.__setitem__("F", _c["F"] + 1))
        hI = ipol.model.register_forward_hook(lambda *a, _c=calls: _c.__setitem__("I", _c["I"] + 1))
        real = torch.randn_like
        torch.randn_like = fake_randn_like
        try:
            with torch.no_grad():
                preds = dyf.predict(ic, condition=static, num_predictions=1) if static is not None \
                    else dyf.predict(ic, num_predictions=1)
        finally:
            torch.randn_like = real
            hF.remove(), hI.remove()
        diff = dyf.model
        kat[name] = dict(dataset=ds, overrides=ov, num_timesteps=diff.num_timesteps,
                         sampling_schedule=[float(s) for s in diff.sampling_schedule],
                         dynamical_steps={str(k): float(v) for k, v in diff.dynamical_steps.items()},
                         calls_F=calls["F"], calls_I=calls["I"], noise_draws=calls["n"], keys=sorted(preds.keys()))
        torch.save({"preds": {k: v.clone() for k, v in preds.items()}, "rows": rows},
                   os.path.join(OUT, f"sample_{name}.pt"))
        print(name, kat[name]["sampling_schedule"], calls, {k: float(v.abs().mean()) for k, v in preds.items()})

    # ---------------- host-logic known-answer vectors ----------------
    from src.diffusion.dyffusion import BaseDYffusion

    class _Bare(BaseDYffusion):  # schedule logic only: no interpolator needed
        def _interpolate(self, *a, **k):
            raise NotImplementedError

        def p_losses(self, *a, **k):
            raise NotImplementedError

    ipol = ref_build.build_interpolator("spring", horizon=4)
    sched_cases = []
    grid = [dict(timesteps=16), dict(timesteps=7, additional_interpolation_steps=25),
            dict(timesteps=5, schedule="linear", additional_interpolation_steps_factor=2, interpolate_before_t1=True),
            dict(timesteps=5, schedule="linear", additional_interpolation_steps_factor=2, interpolate_before_t1=False),
            dict(timesteps=134), dict(timesteps=4, additional_interpolation_steps=3)]
    specs = [None, "only_dynamics", "only_dynamics_plus3", "only_dynamics_plus_discrete3", "every2nd", "every5th",
             "first5", "first0.5", "every3rd", "first1"]
    for g in grid:
        for spec in specs:
            kw = {k: v for k, v in C.DIFFUSION_DEFAULTS.items() if not k.startswith("lambda_")}
            kw.update(g)
            kw["sampling_schedule"] = spec
            try:
                b = _Bare(model=ipol.model, **kw)
                rec = dict(args=g, spec=spec, num_timesteps=b.num_timesteps,
                           schedule=[float(s) for s in b.sampling_schedule],
                           all_int=all(isinstance(s, int) for s in b.sampling_schedule),
                           tau=[float(b.diffusion_step_to_interpolation_step(d)) for d in range(b.num_timesteps)],
                           dynamical={str(k): float(v) for k, v in b.dynamical_steps.items()})
            except (AssertionError, ValueError, IndexError) as e:
                rec = dict(args=g, spec=spec, error=type(e).__name__)
            sched_cases.append(rec)
    kat["schedules"] = sched_cases

    with open(os.path.join(OUT, "state_shapes.json"), "w") as f:
        json.dump(shapes, f, indent=0, sort_keys=True)
    with open(os.path.join(OUT, "schedule_kat.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)
    print("wrote goldens to", OUT)


if __name__ == "__main__":
    main()

"""Randomised differential test of the host-side schedule logic (exact integer / float bookkeeping): the
product's `DiffusionSchedule` and the oracle's `Schedule` against the reference's own `BaseDYffusion` constructor + sampling-
schedule setter (src/diffusion/dyffusion.py:44-138, :245-333) over a few hundred random configurations, including the ones the
reference rejects (same error class of outcome: both raise or both agree).  Build container only (needs /root/reference); the
committed known answers of tests/golden/schedule_kat.json cover the same code elsewhere."""
import random

import pytest

from oracle import configs as C
from oracle import dyffusion_oracle as O

pytestmark = pytest.mark.needs_reference
ERR = (AssertionError, ValueError, IndexError, ZeroDivisionError)


def _reference_factory():
    from oracle import ref_build, ref_shims
    ref_shims.install()
    from src.diffusion.dyffusion import BaseDYffusion

    class Bare(BaseDYffusion):
        def _interpolate(self, *a, **k):
            raise NotImplementedError

        def p_losses(self, *a, **k):
            raise NotImplementedError

    backbone = ref_build.build_interpolator("spring", horizon=4).model
    base = {k: v for k, v in C.DIFFUSION_DEFAULTS.items() if not k.startswith("lambda_")}
    return lambda **kw: Bare(model=backbone, **{**base, **kw})


def _random_config(rng):
    kind = rng.choice(["before_t1_only", "before_t1_only", "linear"])
    cfg = dict(timesteps=rng.randint(2, 12), schedule=kind)
    if kind == "linear":
        cfg.update(additional_interpolation_steps_factor=rng.randint(0, 4), interpolate_before_t1=rng.random() < 0.5)
    else:
        cfg.update(additional_interpolation_steps=rng.choice([0, 0, 1, 2, 3, 5, 9, 25]), interpolate_before_t1=True)
    spec = rng.choice([None, "only_dynamics", f"only_dynamics_plus{rng.randint(1, 6)}",
                       f"only_dynamics_plus_discrete{rng.randint(1, 6)}", f"every{rng.randint(1, 7)}th",
                       f"first{rng.randint(1, 9)}", f"first0.{rng.randint(1, 9)}", "bogus",
                       sorted(rng.sample(range(0, 14), rng.randint(1, 5)))])
    return cfg, spec


def test_schedules_agree_with_the_reference_on_random_configurations():
    from dyffusion_b200.diffusion.schedule import DiffusionSchedule
    make_ref = _reference_factory()
    rng = random.Random(20260117)
    agreed = rejected = 0
    for _ in range(300):
        cfg, spec = _random_config(rng)
        try:
            ref = make_ref(**cfg, sampling_schedule=spec)
            want = dict(n=ref.num_timesteps, sched=[float(s) for s in ref.sampling_schedule],
                        ints=all(isinstance(s, int) for s in ref.sampling_schedule),
                        tau=[float(ref.diffusion_step_to_interpolation_step(d)) for d in range(ref.num_timesteps)],
                        dyn={int(k): float(v) for k, v in ref.dynamical_steps.items()})
        except ERR:
            want = None
        a = (cfg["timesteps"], cfg["schedule"], cfg.get("additional_interpolation_steps", 0),
             cfg.get("additional_interpolation_steps_factor", 0), cfg["interpolate_before_t1"])
        for name in ("product", "oracle"):
            try:
                if name == "product":
                    s = DiffusionSchedule(*a)
                    sched = s.parse_sampling_schedule(spec)
                    tau = [float(s.interpolation_time(d)) for d in range(s.num_timesteps)]
                else:
                    s = O.Schedule(*a, sampling_schedule=spec)
                    sched, tau = s.sampling_schedule, [float(s.tau(d)) for d in range(s.num_timesteps)]
                got = dict(n=s.num_timesteps, sched=[float(v) for v in sched], ints=all(isinstance(v, int) for v in sched),
                           tau=tau, dyn={int(k): float(v) for k, v in s.dynamical_steps.items()})
            except ERR:
                got = None
            assert got == want, (name, cfg, spec, got, want)
        agreed += want is not None
        rejected += want is None
    assert agreed >= 100 and rejected >= 20, (agreed, rejected)  # the generator exercises both outcomes
